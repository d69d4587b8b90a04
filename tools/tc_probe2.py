"""CTA-pair (cta_group::2) tcgen05 probe (csrc/probes/tc_probe2.cu): correctness of the paired MMA / multicast commit / remote TMA
completion, and cycles per MMA of a long MMA stream with one CTA (M = 128) versus a pair (M = 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
print("probe2: importing torch", flush=True)
import torch
from unsuperviseddeephomographyral2018_b200 import _lib
lib = _lib.load_probes()      # libudh_probe.so (csrc/probes/)
g = torch.Generator(device="cuda").manual_seed(3)
A = torch.randn(256, 64, device="cuda", generator=g).bfloat16().contiguous()
for N in (64, 128, 256):
    B = torch.randn(N, 64, device="cuda", generator=g).bfloat16().contiguous()
    ref = A.float() @ B.float().t()
    for pair, remote in ((0, 0), (1, 0), (1, 1)):
        out = torch.full((256, N), float("nan"), device="cuda")
        cyc = torch.zeros(2, device="cuda", dtype=torch.int64)
        print("probe2: N=%d pair=%d remote_tma=%d" % (N, pair, remote), flush=True)
        rc = lib.udh_debug_umma2_probe(A.data_ptr(), B.data_ptr(), out.data_ptr(), cyc.data_ptr(), N, pair, remote, 1, 1, None)
        assert rc == 0, lib.udh_last_error()
        torch.cuda.synchronize()
        want = ref if pair else torch.cat([ref[:128], ref[:128]])
        print("   max|err| = %.3g" % (out - want).abs().max().item(), flush=True)
        if remote:
            continue
        reps = 2048
        for nacc in (1, 2, 4):
            if nacc * N > 512:
                continue
            rc = lib.udh_debug_umma2_probe(A.data_ptr(), B.data_ptr(), out.data_ptr(), cyc.data_ptr(), N, pair, remote, reps, nacc, None)
            assert rc == 0, lib.udh_last_error()
            torch.cuda.synchronize()
            c = cyc.cpu().tolist()
            per = c[0] / (4.0 * reps)
            macs = (256 if pair else 128) * N * 16
            print("   nacc=%d: %.1f cycles / MMA, %.0f MAC/clk per SM (peak 4096)" % (nacc, per, macs / per / (2 if pair else 1)), flush=True)
