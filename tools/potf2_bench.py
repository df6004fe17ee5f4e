"""Time k_potf2 (128 x 128 factor + inverse) in isolation and print its phase stamps (cycles)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from stopro_b200 import _lib

lib = _lib.lib()
dev = torch.device("cuda:0")
torch.manual_seed(0)
X = torch.randn(128, 512, dtype=torch.float64, device=dev)
S = X @ X.t() + 128 * torch.eye(128, dtype=torch.float64, device=dev)
invd = torch.empty(128 * 128, dtype=torch.float64, device=dev)
info = torch.zeros(1, dtype=torch.int32, device=dev)
bufs = [S.clone() for _ in range(64)]
for b in bufs[:4]:
    _lib.check(lib.pigp_potrf_lower(b.data_ptr(), 128, 128, 0, invd.data_ptr(), info.data_ptr(), None))
torch.cuda.synchronize()
L = torch.linalg.cholesky(S)
print("max |L - ref| =", (torch.tril(bufs[0]) - L).abs().max().item(),
      " max |invd L - I| =", (invd.view(128, 128) @ L - torch.eye(128, dtype=torch.float64, device=dev)).abs().max().item())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for b in bufs[4:]:
    _lib.check(lib.pigp_potrf_lower(b.data_ptr(), 128, 128, 0, invd.data_ptr(), info.data_ptr(), None))
e1.record()
torch.cuda.synchronize()
print(f"k_potf2 (+memset info): {e0.elapsed_time(e1) / 60 * 1e3:.1f} us per call, back to back")
st = torch.zeros(32, dtype=torch.int64, device=dev)
_lib.check(lib.pigp_debug_potf2_stamps(st.data_ptr()))
for b in bufs[5:9]:  # the last (warm instruction cache) call's stamps are kept
    _lib.check(lib.pigp_potrf_lower(b.copy_(S).data_ptr(), 128, 128, 0, invd.data_ptr(), info.data_ptr(), None))
torch.cuda.synchronize()
_lib.check(lib.pigp_debug_potf2_stamps(None))
t = st.cpu().tolist()
names = ["load", "potrf32[0]", "panel[0]", "trail[0]+potrf32[1]", "panel[1]", "trail[1]+potrf32[2]", "panel[2]",
         "trail[2]+potrf32[3]", "(loop exit)", "store L", "store diagonal inverse blocks", "(end)"]
for i, nme in enumerate(names):
    print(f"  {nme:22s} {t[i + 1] - t[i]:8d} cycles")
print(f"  total                  {t[12] - t[0]:8d} cycles")
