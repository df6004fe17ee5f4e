"""load_params / load_data with the reference's signatures (sub_modules/load_modules.py:5-20).

``load_params`` reads the three YAML files a prepared simulation directory holds (params_prepare.yaml,
params_main.yaml -- the unchanged schema of default_params/*/params_main.yaml --, lbls.yaml) into plain dicts.
``load_data`` takes any object with the reference's ``HdfOperator`` methods (``load_train_data`` / ``load_test_data``);
h5py is not installed in this image, so the HDF5 reader itself is the caller's.  mu_train / mu_test are zero.
"""
import numpy as np
import yaml


def load_params(params_path="../data_input"):
    out = []
    for name in ("params_main.yaml", "params_prepare.yaml", "lbls.yaml"):
        with open(f"{params_path}/{name}") as file:
            out.append(yaml.safe_load(file))
    return tuple(out)  # params_main, params_prepare, lbls


def load_data(lbls, vnames, hdf_operator):
    r_train, f_train = hdf_operator.load_train_data(lbls["train"], vnames["train"])
    r_test, f_test = hdf_operator.load_test_data(lbls["test"], vnames["test"])
    mu_train = [np.zeros_like(f) for f in f_train]
    mu_test = [np.zeros_like(f) for f in f_test]
    return r_test, mu_test, r_train, mu_train, f_train
