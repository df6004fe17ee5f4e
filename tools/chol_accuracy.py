"""Backward error of the tile Cholesky (pigp_potrf_lower) against cuSOLVER (torch.linalg.cholesky) and LAPACK on the
ill-conditioned C3 training matrix (reference generator's points, eps = 1e-6, cond ~ 1e10), plus the NLL each factor gives
against the long-double truth of the fixture."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import scipy.linalg as sl
import torch

from conftest import oracle_for
from stopro_b200 import _lib, synthetic

name = sys.argv[1] if len(sys.argv) > 1 else "ref_c3_sinusoidal"
path = os.path.join(ROOT, "tests", "golden", name + ".npz")
g = np.load(path)
cfg = synthetic.from_golden(path)
ref = oracle_for(cfg)
th, y, eps = g["theta"], cfg["delta_y"], cfg["eps"]
S = ref.training_sigma(th, cfg["r_train"], eps)
n = len(y)
npad = (n + 127) // 128 * 128
A = np.eye(npad)
A[:n, :n] = S
tr = float(g["truth_nll"])
dev = torch.device("cuda:0")
lib = _lib.lib()


def report(label, L):
    Ld = np.tril(L)
    bw = np.linalg.norm(Ld @ Ld.T - A) / np.linalg.norm(A)
    v = sl.solve_triangular(Ld[:n, :n], y, lower=True)
    nll = 0.5 * v @ v + np.sum(np.log(np.diag(Ld[:n, :n]))) + 0.5 * n * np.log(2 * np.pi)
    print(f"{label:28s} backward error {bw:.2e}   NLL rel err vs truth {abs(nll - tr) / abs(tr):.2e}")


report("LAPACK (numpy)", np.linalg.cholesky(A))
At = torch.as_tensor(A, device=dev)
report("cuSOLVER (torch)", torch.linalg.cholesky(At).cpu().numpy())
buf = At.clone().contiguous()
invd = torch.empty(npad // 128, 128, 128, dtype=torch.float64, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
_lib.check(lib.pigp_potrf_lower(buf.data_ptr(), npad, npad, 0, invd.data_ptr(), info.data_ptr(), None))
torch.cuda.synchronize()
report("pigp_potrf_lower", buf.cpu().numpy())
# accuracy of the inverse diagonal tiles: |W L_kk - I|
Lg = np.tril(buf.cpu().numpy())
W = invd.cpu().numpy()
res = max(np.abs(np.tril(W[k]) @ Lg[k * 128:(k + 1) * 128, k * 128:(k + 1) * 128] - np.eye(128)).max() for k in range(npad // 128))
conds = [np.linalg.cond(Lg[k * 128:(k + 1) * 128, k * 128:(k + 1) * 128]) for k in range(npad // 128)]
print(f"inverse diagonal tiles: max |W L - I| = {res:.2e}; cond(L_kk) up to {max(conds):.2e}")
# the whole path on the same inputs
gp = synthetic.make_model(cfg)
args = (cfg["r_train"], y, eps)
gp.set_constants(*args, only_training=True)
nll = gp.trainingFunction_all(th, *args)
Sg = gp.training_sigma(th, cfg["r_train"], eps)
print(f"GPU path NLL rel err vs truth {abs(nll - tr) / abs(tr):.2e}; |Sigma_gpu - Sigma_oracle| / max = {np.abs(Sg - S).max() / np.abs(S).max():.2e}")
Ag = np.eye(npad)
Ag[:n, :n] = Sg
report("LAPACK on the GPU's Sigma", np.linalg.cholesky(Ag))
