"""A few bf16x3 train steps at the headline size (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from unsuperviseddeephomographyral2018_b200 import engine, synthetic
numeric = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = engine.HomographyEngine(128, numeric=numeric, seed=0, loss_type="h_loss", lr=5e-4)
b = synthetic.make_batch(128, seed=1)
for _ in range(n):
    eng.train_step(b)
torch.cuda.synchronize()
print("done")
