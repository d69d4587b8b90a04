"""Hardware probe of the tcgen05 plumbing (see csrc/probes/tc_probe.cu): prints max errors per mode / shift."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
print("probe: importing torch", flush=True)
import torch
print("probe: torch imported, cuda", torch.cuda.is_available(), flush=True)
from unsuperviseddeephomographyral2018_b200 import _lib
lib = _lib.load_probes()      # libudh_probe.so (csrc/probes/)

def run(A, B, mode, bo):
    print("probe: launching mode", mode, "bo", bo, flush=True)
    out = torch.full((128, 512), float("nan"), device="cuda")
    rc = lib.udh_debug_umma_probe(A.data_ptr(), A.shape[0], B.data_ptr(), B.shape[0], out.data_ptr(), mode, bo, None)
    assert rc == 0, lib.udh_last_error()
    torch.cuda.synchronize()
    return out

torch.manual_seed(0)
g = torch.Generator(device="cuda").manual_seed(1)
# ---- mode 0: K-major, row-shifted A
A = torch.randn(144, 64, device="cuda", generator=g).bfloat16().contiguous()
B = torch.randn(64, 64, device="cuda", generator=g).bfloat16().contiguous()
for bo in (1, 0):
    out = run(A, B, 0, bo)
    errs = []
    for s in range(8):
        ref = A[s:s + 128].float() @ B.float().t()
        errs.append((out[:, 64 * s:64 * s + 64] - ref).abs().max().item())
    print("mode0 K-major  base_offset=%d  max|err| per row shift:" % bo, ["%.3g" % e for e in errs])
# ---- mode 1: MN-major, G two 64-wide blocks, X shifted
G = torch.randn(128, 128, device="cuda", generator=g).bfloat16()          # [px][co]
Gb = torch.cat([G[:, :64], G[:, 64:]], dim=0).contiguous()                # [2*128][64] block layout
X = torch.randn(144, 64, device="cuda", generator=g).bfloat16().contiguous()
for bo in (1, 0):
    out = run(Gb, X, 1, bo)
    errs = []
    for s in range(8):
        ref = G.float().t() @ X[s:s + 128].float()                        # [co=128][ci=64]
        errs.append((out[:, 64 * s:64 * s + 64] - ref).abs().max().item())
    print("mode1 MN-major base_offset=%d  max|err| per px shift: " % bo, ["%.3g" % e for e in errs])
# ---- mode 2: M=64 accumulator layout
out = run(A, B, 2, 0)
ref = A[:64].float() @ B.float().t()
o = out[:, :64]
print("mode2 M=64: finite lanes:", torch.isfinite(o).all(dim=1).nonzero().flatten().tolist()[:80])
for name, rows in (("lanes 0..63", list(range(64))), ("16 per subpartition", [32 * (i // 16) + i % 16 for i in range(64)])):
    print("   hypothesis %-22s max|err| = %.3g" % (name, (o[rows] - ref).abs().max().item()))
