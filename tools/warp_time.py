"""BASELINE configs[3] (fused warp + L1 on the full 320x240 grid, B = 64) and the train-step window (128x128, B = 128):
CUDA-event time per launch of udh_warp_loss_fwd_ex, inputs rotating through > L2 worth of buffers."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unsuperviseddeephomographyral2018_b200 import _lib
p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
g = torch.Generator(device="cuda").manual_seed(5)


def run(B, Hh, W, pw, ph, all_sums, windowed, nb=6, iters=60):
    src = [torch.randn(B, Hh, W, 1, device="cuda", generator=g) for _ in range(nb)]
    tgt = [torch.randn(B, ph, pw, 1, device="cuda", generator=g) for _ in range(nb)]
    pts = torch.tensor([[96., 56., 224., 56., 224., 184., 96., 184.]], device="cuda").repeat(B, 1).contiguous()
    hh = (torch.rand(B, 8, device="cuda", generator=g) * 20 - 10).contiguous()
    Hm = torch.empty(B, 9, device="cuda"); sums = torch.zeros(8, device="cuda", dtype=torch.float64)
    idx = (torch.full((B,), 56 * W + 96, device="cuda", dtype=torch.int32)) if windowed else None
    _lib.check(_lib.lib.udh_dlt_fwd(p(pts), p(hh), p(Hm), B, None), "dlt")
    def go(i):
        _lib.check(_lib.lib.udh_warp_loss_fwd_ex(p(src[i % nb]), 1, Hh, W, p(Hm), p(tgt[i % nb]), p(idx), 1 if windowed else 0, pw, ph, None,
                                                 p(sums), all_sums, B, None), "warp")
    for i in range(10):
        go(i)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for i in range(iters):
        go(i)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    byts = B * ph * pw * 4 * 2 if not windowed else B * (ph * pw * 4 * 2 + 36)
    print("B=%d grid %dx%d window %dx%d all_sums=%d: %.2f us/launch, %.0f GB/s algorithmic (%.1f MB), l1=%.5f"
          % (B, W, Hh, pw, ph, all_sums, us, byts / us / 1e3, byts / 1e6, sums[0].item() / (B * ph * pw * (iters + 10))))


run(64, 240, 320, 320, 240, 0, False)
run(64, 240, 320, 320, 240, 1, False)
run(128, 240, 320, 128, 128, 1, True)
run(128, 240, 320, 128, 128, 0, True)
