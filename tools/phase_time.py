"""Per-phase CUDA-event times of the train step (library brackets, udh_prof_*) for one numeric mode on one GPU.
usage: python tools/phase_time.py [numeric=bf16x3] [B=128] [steps=10]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unsuperviseddeephomographyral2018_b200 import _lib, engine, synthetic

numeric = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
eng = engine.HomographyEngine(B, numeric=numeric, seed=0, loss_type="h_loss", lr=5e-4)
bs = [synthetic.make_batch(B, seed=i) for i in range(3)]
for i in range(3):
    eng.train_step(bs[i % 3])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(K):
    eng.train_step(bs[i % 3])
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / K
print("%s B=%d: %.3f ms/step, %.1f k pairs/s" % (numeric, B, ms, B / ms))
_lib.lib.udh_prof_enable(1); _lib.lib.udh_prof_reset()
for i in range(K):
    eng.train_step(bs[i % 3])
torch.cuda.synchronize()
ph = _lib.prof_read_all()
_lib.lib.udh_prof_enable(0)
tot = 0.0
for k, (t, n) in sorted(ph.items(), key=lambda kv: -kv[1][0]):
    print("  %-16s %8.4f ms/step  (%d launches-brackets/step)" % (k, t / K, n // K))
    tot += t / K
print("  sum of phases %.3f ms" % tot)
