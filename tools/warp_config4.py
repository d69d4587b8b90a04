"""BASELINE configs[3]: fused warp + L1 on the full 320x240 grid, B = 64 (standalone launcher for ncu)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unsuperviseddeephomographyral2018_b200 import _lib
B, Hh, W, nb = 64, 240, 320, 4
g = torch.Generator(device="cuda").manual_seed(5)
src = [torch.randn(B, Hh, W, 1, device="cuda", generator=g) for _ in range(nb)]
tgt = [torch.randn(B, Hh, W, 1, device="cuda", generator=g) for _ in range(nb)]
pts = torch.tensor([[96., 56., 224., 56., 224., 184., 96., 184.]], device="cuda").repeat(B, 1).contiguous()
hh = (torch.rand(B, 8, device="cuda", generator=g) * 20 - 10).contiguous()
Hm = torch.empty(B, 9, device="cuda"); sums = torch.zeros(8, device="cuda", dtype=torch.float64)
p = lambda t: ctypes.c_void_p(t.data_ptr())
_lib.check(_lib.lib.udh_dlt_fwd(p(pts), p(hh), p(Hm), B, None), "dlt")
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 8):
    _lib.check(_lib.lib.udh_warp_loss_fwd_ex(p(src[i % nb]), 1, Hh, W, p(Hm), p(tgt[i % nb]), None, 0, W, Hh, None, p(sums), 0, B, None), "warp")
torch.cuda.synchronize()
print("l1 =", sums[0].item() / (B * Hh * W))
