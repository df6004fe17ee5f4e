import os
import sys

import pytest

# tests/test_gpu_dist.py runs several ranks (3 streams each, with spinning flag waits) inside one process: every stream
# needs its own hardware queue or a wait in one stream stalls an unrelated one (the default is 8 queues per process)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# a dead-lock of that single-GPU emulation must fail a test in seconds, not after the production time-outs (10 s / 300 s)
os.environ.setdefault("PIGP_WAIT_TIMEOUT_S", "6")
os.environ.setdefault("PIGP_BARRIER_TIMEOUT_S", "20")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def oracle_for(cfg, backend="closed"):
    """The CPU oracle (oracle.gp_ref.GPRef) configured like a stopro_b200.synthetic configuration."""
    from oracle.gp_ref import GPRef

    kw = cfg["model_kwargs"]
    gp = GPRef(cfg["model"], kernel_form=cfg["kernel"]["kernel_form"], dim=cfg["kernel"]["input_dim"],
               lbox=kw.get("lbox"), index_optimize_noise=kw.get("index_optimize_noise"), backend=backend,
               kernel_type=cfg["kernel"].get("kernel_type", "se"))
    gp.set_constants(cfg["r_test"], cfg["mu_test"], cfg["r_train"], cfg["delta_y"], cfg["eps"])
    return gp


@pytest.fixture(scope="session")
def cuda_device():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")
