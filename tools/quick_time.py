"""Scratch timing of the step phases on one GPU (CUDA events). Not the bench; used while developing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import oracle as O
from unsuperviseddeephomographyral2018_b200 import engine

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
numeric = sys.argv[2] if len(sys.argv) > 2 else "fp32"
loss = sys.argv[3] if len(sys.argv) > 3 else "h_loss"
b4 = O.make_batch(0, 4)
batch = {k: v.repeat(B // 4, *([1] * (v.dim() - 1))).cuda().contiguous() for k, v in b4.items() if isinstance(v, torch.Tensor) and k != "H_gt"}
eng = engine.HomographyEngine(B, seed=0, numeric=numeric, loss_type=loss, lr=5e-4)

def timeit(fn, n=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

out = eng.forward(batch, train=True)
print("fwd(train)  ms", timeit(lambda: eng.forward(batch, train=True)))
def bwd():
    eng.grads.zero_(); eng.backward(batch, out)
print("bwd         ms", timeit(bwd))
print("adam        ms", timeit(lambda: eng.update()))
t = timeit(lambda: eng.train_step(batch))
print("train_step  ms", t, "pairs/s", B / t * 1e3)
print(eng.losses_dict(eng.forward(batch, train=False)))
