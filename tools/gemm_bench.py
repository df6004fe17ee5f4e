"""Time pigp_dgemm (the DMMA kernel) on square / panel shapes against cuBLAS DGEMM (torch.matmul)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stopro_b200 import _lib

lib = _lib.lib()
dev = torch.device("cuda:0")
shapes = [(8192, 8192, 8192, 0), (8192, 8192, 8192, 1), (10112, 4992, 4992, 0), (16384, 128, 128, 0), (4096, 4096, 512, 1),
          (2048, 2048, 2048, 0), (1024, 1024, 1024, 0)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in sys.argv[1].split(","))]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for (M, N, K, lower) in shapes:
    A = torch.randn(M, K, dtype=torch.float64, device=dev)
    B = torch.randn(N, K, dtype=torch.float64, device=dev)
    Cm = torch.zeros(M, N, dtype=torch.float64, device=dev)

    def ours():
        _lib.check(lib.pigp_dgemm(M, N, K, 1.0, A.data_ptr(), K, 1, B.data_ptr(), K, 1, 0.0, Cm.data_ptr(), N, lower, None))

    def cublas():
        torch.matmul(A, B.t(), out=Cm)

    res = {}
    for name, fn in (("ours", ours), ("cublas", cublas)):
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
        fl = 2.0 * M * N * K * (0.5 * (1 + 128.0 / N) if (lower and name == "ours") else 1.0)
        res[name] = (best, fl / best * 1e-9)
    print(f"M={M} N={N} K={K} lower={lower}: ours {res['ours'][0]:.3f} ms {res['ours'][1]:.2f} TF | cublas {res['cublas'][0]:.3f} ms {res['cublas'][1]:.2f} TF")
