import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
    if os.path.exists(f): print(f, open(f).read().strip())
from oracle import oracle as O
from unsuperviseddeephomographyral2018_b200 import params as P
specs = P.param_specs(); flat = torch.tensor(P.init_flat(0)); batch = O.make_batch(0, 4)
for nt in (8, 16, 32, 64):
    torch.set_num_threads(nt)
    m, v = torch.zeros_like(flat), torch.zeros_like(flat)
    ts = []
    for i in range(2):
        t0 = time.perf_counter(); O.train_step(flat, m, v, 0, batch, specs, loss_type="h_loss", lr=5e-4); ts.append(time.perf_counter() - t0)
    print("threads", nt, "B=4 train step s:", ["%.2f" % t for t in ts], "pairs/s %.2f" % (4 / ts[-1]))
