"""Secondary on-GPU comparator: cuSOLVER/cuBLAS FP64 through torch on the same B200 --
torch.linalg.cholesky (potrf), torch.cholesky_inverse (potri), vs pigp_potrf_lower / pigp_potri_lower."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stopro_b200 import _lib

lib = _lib.lib()
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 20096
torch.manual_seed(0)
X = torch.randn(n, 512, dtype=torch.float64, device=dev)
S = X @ X.t()
S.diagonal().add_(float(n))
del X


def timeit(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


buf = torch.empty_like(S)
invd = torch.empty(n // 128, 128, 128, dtype=torch.float64, device=dev)
W = torch.empty_like(S)


def ours_potrf():
    buf.copy_(S)
    _lib.check(lib.pigp_potrf_lower(buf.data_ptr(), n, n, 0, invd.data_ptr(), None, None))


def ours_potri():
    _lib.check(lib.pigp_potri_lower(buf.data_ptr(), n, n, invd.data_ptr(), W.data_ptr(), buf.data_ptr(), None))


t_copy = timeit(lambda: buf.copy_(S))
t_potrf = timeit(ours_potrf) - t_copy
ours_potrf()
t_potri = timeit(ours_potri, reps=1)
print(f"n={n}: ours potrf {t_potrf:.2f} ms ({n**3 / 3 / t_potrf * 1e-9:.2f} TF), ours potri {t_potri:.2f} ms ({2 * n**3 / 3 / t_potri * 1e-9:.2f} TF)")
del W
L = None
t_cs = timeit(lambda: torch.linalg.cholesky(S))
L = torch.linalg.cholesky(S)
t_ci = timeit(lambda: torch.cholesky_inverse(L), reps=1)
print(f"n={n}: cusolver potrf {t_cs:.2f} ms ({n**3 / 3 / t_cs * 1e-9:.2f} TF), torch.cholesky_inverse {t_ci:.2f} ms ({2 * n**3 / 3 / t_ci * 1e-9:.2f} TF)")
