import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from oracle import oracle as O
from unsuperviseddeephomographyral2018_b200 import engine, params
def rel(a, b): return (np.linalg.norm(a.double() - b.double()) / max(np.linalg.norm(b.double()), 1e-30)).item() if isinstance(a, torch.Tensor) else 0
seed, B = 0, 4
batch = {k: (v.cuda().contiguous() if isinstance(v, torch.Tensor) else v) for k, v in O.make_batch(seed, B).items()}
for lt in ("h_loss",):
    e32 = engine.HomographyEngine(B, seed=seed, numeric="fp32", loss_type=lt, lr=5e-4)
    e16 = engine.HomographyEngine(B, seed=seed, numeric="bf16", loss_type=lt, lr=5e-4)
    o32 = e32.forward(batch, train=True, dropout_seed=123); o16 = e16.forward(batch, train=True, dropout_seed=123)
    print("pred32", o32["pred_h4p"][0].tolist()); print("pred16", o16["pred_h4p"][0].tolist())
    e32.backward(batch, o32); e16.backward(batch, o16)
    specs = params.param_specs(); g32, g16 = e32.grads.cpu(), e16.grads.cpu()
    for name, s in specs.items():
        a, b = g16[s.offset:s.offset + s.size], g32[s.offset:s.offset + s.size]
        cos = (a.double() @ b.double() / (a.double().norm() * b.double().norm() + 1e-30)).item()
        print("%-40s rel %.4f  cos %.5f  |g32| %.3e |g16| %.3e" % (name, rel(a, b), cos, b.norm().item(), a.norm().item()))
    # repeat bf16 backward to see run-to-run noise
    e16.grads.zero_(); e16.backward(batch, o16); g16b = e16.grads.cpu()
    s = specs["model/conv_block1/conv1/weights"]
    print("bf16 run-to-run conv1_1 rel:", rel(g16b[s.offset:s.offset + s.size], g16[s.offset:s.offset + s.size]))
