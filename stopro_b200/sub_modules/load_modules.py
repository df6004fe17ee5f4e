"""load_params / load_data with the reference's signatures (sub_modules/load_modules.py:5-20).

``load_params`` reads the three YAML files a prepared simulation directory holds -- params_main.yaml (the unchanged
schema of default_params/*/params_main.yaml), params_prepare.yaml and lbls.yaml -- into plain dicts and returns them
in the reference's order (main, prepare, labels).  ``load_data`` works with any object that offers the reference
``HdfOperator``'s two readers (h5py is not installed in this image, so the HDF5 reader itself stays with the caller)
and supplies the zero prior means the scripts pass around.
"""
import os

import numpy as np
import yaml

_FILES = ("params_main.yaml", "params_prepare.yaml", "lbls.yaml")


def load_params(params_path="../data_input"):
    loaded = []
    for name in _FILES:
        with open(os.path.join(params_path, name)) as handle:
            loaded.append(yaml.safe_load(handle))
    return tuple(loaded)


def _zero_means(values):
    return [np.zeros_like(v) for v in values]


def load_data(lbls, vnames, hdf_operator):
    """-> (r_test, mu_test, r_train, mu_train, f_train), the argument order of the reference's scripts."""
    train = hdf_operator.load_train_data(lbls["train"], vnames["train"])
    test = hdf_operator.load_test_data(lbls["test"], vnames["test"])
    return test[0], _zero_means(test[1]), train[0], _zero_means(train[1]), train[1]
