"""Worker of tests/test_gpu_dp.py (launched with torchrun, one process per GPU): drives the repo's HomographyEngine
data-parallel and checks Row G (code/utils/utils.py:380-403, code/homography_CNN_synthetic.py:199-207,277-284)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.distributed as dist

from unsuperviseddeephomographyral2018_b200 import engine as en, params as P, synthetic


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    numeric = sys.argv[1]
    out_path = sys.argv[2]
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pg = dist.group.WORLD
    B = 4
    res = {}
    flat = P.init_flat_large(0)
    eng = en.HomographyEngine(B, numeric=numeric, seed=None, loss_type="h_loss", lr=5e-4, process_group=pg, world_size=world)
    eng.load_flat(flat)
    ref = en.HomographyEngine(B, numeric=numeric, seed=None, loss_type="h_loss", lr=5e-4)      # single-GPU engine, same rank-local data
    ref.load_flat(flat)
    ref.dropout_seed = eng.dropout_seed
    # per-rank dropout streams differ
    seeds = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(seeds, torch.tensor([eng.dropout_seed & 0x7FFFFFFFFFFFFFFF], dtype=torch.int64, device="cuda"))
    res["dropout_seeds_distinct"] = len(set(int(s.item()) for s in seeds)) == world
    # ---- (1) the reduced gradient of the overlapped two-phase path == mean over ranks of the single-GPU gradients
    batch = synthetic.make_batch(B, seed=100 + rank)
    out = eng.forward(batch, train=True, dropout_seed=777 + rank)
    eng.backward(batch, out)                # head phase + allreduce of the fc slice on the comm stream under the conv backward
    eng.allreduce_grads()                   # conv slice + join
    torch.cuda.synchronize()
    g_dp = eng.grads.clone() / world
    o2 = ref.forward(batch, train=True, dropout_seed=777 + rank)
    ref.backward(batch, o2)
    g_local = ref.grads.clone()
    g_sum = g_local.clone()
    dist.all_reduce(g_sum)                  # plain one-shot allreduce as the oracle of the schedule
    g_mean = g_sum / world
    head, convs = en.dp_slices(eng.specs)
    res["rel_err_fc_slice"] = rel(g_dp[head], g_mean[head])
    res["rel_err_conv_slice"] = rel(g_dp[convs], g_mean[convs])
    res["ranks_differ_locally"] = rel(g_local, g_mean) > 1e-3           # the test is not vacuous: local gradients differ
    # ---- (2) three optimiser steps: parameters stay bit-identical across ranks, and move
    eng.grads.zero_(); ref.grads.zero_()
    p0 = eng.params.clone()
    for i in range(3):
        eng.train_step(synthetic.make_batch(B, seed=1000 * rank + i))
    torch.cuda.synchronize()
    mine = eng.params.clone()
    root = mine.clone()
    dist.broadcast(root, 0)
    same = torch.tensor([1.0 if torch.equal(root, mine) else 0.0], device="cuda")
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    res["params_bit_identical_across_ranks"] = bool(same.item() == 1.0)
    res["params_moved"] = float((mine - p0).abs().max())
    res["global_step"] = eng.global_step
    res["default_path_is_nccl"] = eng._mc is None
    # ---- (3) one DP step == single-GPU Adam on the averaged gradient (1/N folded into grad_scale)
    e1 = en.HomographyEngine(B, numeric=numeric, seed=None, loss_type="h_loss", lr=5e-4, process_group=pg, world_size=world); e1.load_flat(flat)
    e2 = en.HomographyEngine(B, numeric=numeric, seed=None, loss_type="h_loss", lr=5e-4); e2.load_flat(flat)
    e2.dropout_seed = e1.dropout_seed
    b = synthetic.make_batch(B, seed=300 + rank)
    e1.train_step(b)
    o = e2.forward(b, train=True); e2.backward(b, o)
    dist.all_reduce(e2.grads); e2.grads /= world
    e2.update()
    torch.cuda.synchronize()
    big = g_mean.abs() > 0            # everywhere
    upd1, upd2 = e1.params - torch.tensor(flat, device="cuda"), e2.params - torch.tensor(flat, device="cuda")
    # TF-Adam's first step is ~lr*sign(g): compare where the gradient is not rounding noise
    gm = e2.adam_m.abs()
    sel = gm > 1e-3 * gm.max()
    res["dp_step_vs_manual_max_update_diff_over_lr"] = float((upd1 - upd2)[sel].abs().max() / 5e-4)
    # ---- (4) the multicast path (switch-side sum, sharded Adam, multicast weights) against the NCCL path, three steps each.
    # The mode is opt-in and needs NVSwitch multicast memory: a box without it reports that instead of failing the suite.
    e3 = en.HomographyEngine(B, numeric=numeric, seed=None, loss_type="h_loss", lr=5e-4, process_group=pg, world_size=world); e3.load_flat(flat)
    res["nccl_engine_is_nccl"] = e3._mc is None
    os.environ["UDH_DP_MODE"] = "multicast"
    e4, why = None, ""
    try:
        e4 = en.HomographyEngine(B, numeric=numeric, seed=None, loss_type="h_loss", lr=5e-4, process_group=pg, world_size=world)
    except Exception as e:                                    # symmetric memory / multicast not available here
        why = "%s: %s" % (type(e).__name__, e)
    os.environ.pop("UDH_DP_MODE")
    ok = torch.tensor([1.0 if e4 is not None and e4._mc is not None else 0.0], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)                # all ranks take the same branch
    res["multicast_path"] = bool(ok.item() == 1.0)
    if not res["multicast_path"]:
        res["multicast_unavailable"] = why[:300]
    else:
        e4.load_flat(flat)
        e4.dropout_seed = e3.dropout_seed
        for i in range(3):
            bb = synthetic.make_batch(B, seed=500 + 10 * rank + i)
            e3.train_step(bb); e4.train_step(bb)
        torch.cuda.synchronize()
        # Adam's m, v of fc1's weights are sharded on the multicast path: after the gather every rank holds all of them
        e4.sync_optimizer_state()
        mm = e4.adam_m.clone(); mr = mm.clone(); dist.broadcast(mr, 0)
        same = torch.tensor([1.0 if torch.equal(mr, mm) else 0.0], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        res["adam_m_identical_after_sync"] = bool(same.item() == 1.0)
        w = e4.specs["model/fc1/fc1/weights"]
        cnt = 1
        for d_ in w.shape:
            cnt *= int(d_)
        per = -(-cnt // (4 * world)) * 4
        res["adam_m_fc1_nonzero_fraction_per_shard"] = [float((mm[w.offset + r * per:min(w.offset + (r + 1) * per, w.offset + cnt)].abs() > 0).float().mean())
                                                        for r in range(world)]
        mine4 = e4.params.clone(); root4 = mine4.clone(); dist.broadcast(root4, 0)
        same = torch.tensor([1.0 if torch.equal(root4, mine4) else 0.0], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        res["multicast_params_bit_identical_across_ranks"] = bool(same.item() == 1.0)
        d = (e3.params - e4.params).abs()
        res["multicast_vs_nccl_max_param_diff_over_lr"] = float(d.max() / 5e-4)
        res["multicast_vs_nccl_rel_l2_of_update"] = rel(e4.params - torch.tensor(flat, device="cuda"), e3.params - torch.tensor(flat, device="cuda"))
        if e4._mirror is not None:                       # the replicas' tensor-core limbs follow the fp32 master
            mp, mb, mcnt, _ = e4._mirror
            off = mp - e4.ws.data_ptr()
            planes = e4.ws[off:off + mcnt * (4 if numeric == "bf16x3" else 2)].view(torch.bfloat16).view(-1, mcnt).float()
            res["mirror_matches_master"] = rel(planes.sum(0), e4.params[mb:mb + mcnt]) < (1e-4 if numeric == "bf16x3" else 1e-2)
    if rank == 0:
        with open(out_path, "w") as f:
            json.dump(res, f)
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
