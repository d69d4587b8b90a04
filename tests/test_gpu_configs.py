"""GPU tests of the BASELINE.json configurations that had no parity test in round 1:
  configs[0]  `homography_CNN_synthetic.py --mode test`, batch 4 (the reference's CPU-runnable plumbing case): the printed result
              table against the oracle's test metrics on the same batches;
  configs[2]  inference-only CNN + DLT forward at batch 512: sampled rows against the CPU oracle (the oracle cannot run 512
              samples in seconds; rows are independent, so 8 rows spread over the CTA / tile boundaries pin the batch)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from oracle import oracle as O                                               # noqa: E402


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


@pytest.mark.parametrize("numeric", ["bf16x3", "fp32"])
def test_config2_inference_forward_B512_sampled_rows_vs_oracle(numeric):
    _need_gpu()
    from unsuperviseddeephomographyral2018_b200 import engine, params, synthetic, ops
    B = 512
    flat = params.init_flat_large(1)
    eng = engine.HomographyEngine(B, seed=None, numeric=numeric); eng.load_flat(flat)
    batch = synthetic.make_batch(B, seed=2024)
    out = eng.eval_step(batch)
    h = out["pred_h4p"].cpu().double()
    H = out["H_mat"].cpu().double()
    rows = [0, 1, 63, 64, 127, 128, 300, 511]
    p = params.unflatten(torch.tensor(flat), params.param_specs())
    x = torch.cat([batch["I1_aug"][rows], batch["I2_aug"][rows]], dim=3).cpu()
    ref = O.vgg_forward(p, x, None).double()
    scale = ref.abs().max().item()
    assert scale > 10.0
    err = (h[rows] - ref).abs().max().item() / scale
    print("configs[2] B=512 %s: max rel err of pred_h4p on 8 sampled rows %.2e (|pred| up to %.1f px)" % (numeric, err, scale))
    assert err <= (3e-5 if numeric == "bf16x3" else 1e-5)
    Href = O.solve_dlt(batch["pts1"][rows].cpu().double(), h[rows])         # DLT of the engine's own prediction, fp64
    hs = Href.abs().amax(dim=(1, 2), keepdim=True)
    assert ((H[rows] - Href).abs() / hs).max().item() < 2e-4


def test_config0_cli_test_mode_table_vs_oracle(tmp_path):
    """`--mode test --batch_size 4 --synthetic 16` (3 epochs = 12 iterations): the printed table must equal the oracle's
    bounded_h_loss / l1_loss / failure rate on the very same batches and weights."""
    _need_gpu()
    from unsuperviseddeephomographyral2018_b200 import dataloader as dl, params, synthetic
    seed, nb, bs = 3, 16, 4
    model_dir, log_dir, res_dir = str(tmp_path / "m"), str(tmp_path / "l"), str(tmp_path / "r")
    # a checkpoint with the large-output weights (the CLI restores <model_dir>/<prefix>/<model_name>-<step>.pt)
    flat = params.init_flat_large(seed)
    ck_dir = os.path.join(model_dir, "l1_loss_normalize")
    os.makedirs(ck_dir)
    z = torch.zeros(len(flat))
    torch.save(dict(params=torch.tensor(flat), adam_m=z, adam_v=z, global_step=7, patch_size=128), os.path.join(ck_dir, "model.ckpt-7.pt"))
    cmd = [sys.executable, os.path.join(ROOT, "homography_CNN_synthetic.py"), "--mode", "test", "--batch_size", str(bs), "--synthetic", str(nb),
           "--num_gpus", "1", "--numeric", "bf16x3", "--seed", str(seed), "--do_augment", "0", "--model_dir", model_dir, "--log_dir", log_dir,
           "--results_dir", res_dir]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = r.stdout.splitlines()
    i = next(k for k, ln in enumerate(lines) if ln.startswith("|Steps"))
    step, h_loss, l1_loss, fail_pct = [float(v) for v in lines[i + 1].split()]
    assert "no checkpoint" not in r.stdout and int(step) == 3 * (nb // bs) - 1
    # the same batches through the oracle
    dparams = dl.dataloader_params(data_path="", filenames_file=None, pts1_file=None, gt_file=None, mode='test', batch_size=bs, img_h=240,
                                   img_w=320, patch_size=128, augment_list=['normalize'], do_augment=0.0)
    loader = dl.Dataloader(dparams, shuffle=True, synthetic_pairs=nb, seed=seed * 97 + 12345, device="cuda")
    p = params.unflatten(torch.tensor(flat), params.param_specs())
    hs, l1s, fails = [], [], 0.0
    for _ in range(3 * (nb // bs)):
        b = {k: (v.cpu() if isinstance(v, torch.Tensor) else v) for k, v in loader.next_batch().items()}
        # the fast generator hands over the window ORIGIN of each sample; the oracle gathers with the full index table
        x0 = b["pts1"][:, 0].long(); y0 = b["pts1"][:, 1].long()
        yy, xx = torch.meshgrid(torch.arange(128), torch.arange(128), indexing="ij")
        b["patch_indices"] = ((yy[None] + y0[:, None, None]) * 320 + (xx[None] + x0[:, None, None])).reshape(bs, -1).to(torch.int32)
        out = O.forward(p, b, None, mode="test")
        hs.append(float(out["bounded_h_loss"])); l1s.append(float(out["l1_loss"])); fails += float(out["num_fail"])
    assert abs(h_loss - np.mean(hs)) <= 1e-3, (h_loss, np.mean(hs))
    assert abs(l1_loss - np.mean(l1s)) <= 2e-4 * np.mean(l1s) + 1e-6, (l1_loss, np.mean(l1s))
    assert abs(fail_pct - 100.0 * fails / len(hs) / bs) < 1e-9
    assert re.search(r"Percentile Values", r.stdout)


def test_fused_prep_kernel_augmentation_vs_oracle():
    """udh_prep_inputs_u8_ex (augment on the raw 0..255 values + normalise + gray + patch gather, one kernel) against the
    oracle's restatement of code/dataloader.py:163-177,203-227,323-375, joint and disjoint parameter tables; and the fast
    on-device generator's pairs are self-consistent (warp with H_gt reproduces I2 up to the uint8 cast)."""
    _need_gpu()
    from unsuperviseddeephomographyral2018_b200 import synthetic, ops
    B = 5
    sb = synthetic.make_batch_fast(B, seed=11)
    assert sb["pts1"][:, 0].min() >= 45 and sb["pts1"][:, 0].max() <= 147 and sb["pts1"][:, 1].min() >= 45 and sb["pts1"][:, 1].max() <= 67
    assert sb["gt"].abs().max() <= 45 and torch.equal(sb["gt"], sb["gt"].round()) and sb["gt"].abs().max() >= 30
    assert sb["I_u8"].float().std() > 20 and sb["I_u8"].float().mean() > 80 and sb["I_u8"].float().mean() < 175
    # the second image is the reference's generator step: uint8(transformer(I, M^-1 H_gt M)) (gen_synthetic_data.py:56-64)
    Hgt = ops.dlt_forward(sb["pts1"], sb["gt"])
    M = torch.tensor([[160.0, 0., 160.0], [0., 120.0, 120.0], [0., 0., 1.]], device="cuda")
    theta = (torch.linalg.inv(M) @ Hgt @ M).contiguous()
    Ip_ref, _ = ops.transformer(sb["I_u8"].float().contiguous(), theta, (240, 320))
    dI = (sb["I_prime_u8"].int() - Ip_ref.clamp(0, 255).to(torch.uint8).int()).abs()
    assert (dI > 1).float().mean().item() < 1e-3 and dI.float().mean().item() < 0.05     # theta is rounded differently: isolated +-1 levels
    # self-consistency: warping I with H_gt reproduces the I2 patch up to the uint8 truncation of I' (< 1.5 grey levels)
    _, sums = ops.warp_loss_forward(sb["I_aug"], Hgt, sb["I2_aug"], sb["patch_indices"], 128, 128, want_pred=False)
    print("mean |warp(I, H_gt) - I2| = %.4f normalised units" % (sums[0].item() / (B * 128 * 128)))
    assert sums[0].item() / (B * 128 * 128) < 1.5 / 69.0
    rng = np.random.default_rng(5)
    for joint in (True, False):
        aug = np.zeros((B, 11), np.float32)
        aug[:, 0] = [1, 0, 1, 1, 1]
        a = np.concatenate([rng.uniform(0.8, 1.2, (B, 1)), rng.uniform(0.5, 2.0, (B, 1)), rng.uniform(0.8, 1.2, (B, 3))], 1)
        b = a if joint else np.concatenate([rng.uniform(0.8, 1.2, (B, 1)), rng.uniform(0.5, 2.0, (B, 1)), rng.uniform(0.8, 1.2, (B, 3))], 1)
        aug[:, 1:6], aug[:, 6:11] = a, b
        got = synthetic.prep_u8(sb["I_u8"], sb["I_prime_u8"], sb["pts1"], torch.tensor(aug).cuda(), 128, want_rgb=True)
        ref = O.prep_inputs(sb["I_u8"].cpu(), sb["I_prime_u8"].cpu(), sb["pts1"].cpu(), torch.tensor(aug), 128)
        assert (got["I_aug_rgb"].cpu() - ref["I_aug"]).abs().max() <= 2e-4          # powf vs torch.pow: a few ulp of 255^1.2
        assert (got["I_aug"].cpu()[..., 0] - ref["I_aug"].mean(dim=3)).abs().max() <= 2e-4
        for k in ("I1", "I2", "I1_aug", "I2_aug"):
            assert (got[k].cpu() - ref[k]).abs().max() <= 2e-4, k
        assert torch.equal(got["patch_indices"].cpu(), ref["patch_indices"][:, 0])
        # sample 1 is not augmented: identical to the plain path, bit for bit
        assert torch.equal(got["I1_aug"][1], got["I1"][1]) and torch.equal(got["I2_aug"][1], got["I2"][1])
        # an augmented sample really changed, and saturates like the reference (x ** gamma on 0..255 clips at 255)
        assert (got["I1_aug"][0] - got["I1"][0]).abs().max() > 0.05


def test_real_data_path_gtless_finetune_and_correspondence_test(tmp_path):
    """homography_CNN_real.py end to end on the synthetic stand-in at the real-data geometry (142x190 images): a synthetic
    model is saved as a TensorFlow checkpoint, --finetune restores it with the step reset and trains WITHOUT ground truth
    (l1_loss), the new checkpoint lands in --save_model_dir, and --mode test prints the correspondence-metric table."""
    _need_gpu()
    from unsuperviseddeephomographyral2018_b200 import engine, params, tf_checkpoint as tfc
    load_dir, save_dir = str(tmp_path / "syn"), str(tmp_path / "real")
    eng = engine.HomographyEngine(2, seed=4, numeric="fp32")
    eng.global_step = 999
    os.makedirs(os.path.join(load_dir, "l1_loss_normalize"))
    prefix = os.path.join(load_dir, "l1_loss_normalize", "model.ckpt-999")
    eng.save_tf_checkpoint(prefix)
    tfc.update_checkpoint_state(os.path.dirname(prefix), prefix)
    common = ["--num_gpus", "1", "--synthetic", "64", "--batch_size", "8", "--seed", "4", "--load_model_dir", load_dir + "/", "--save_model_dir", save_dir + "/",
              "--log_dir", str(tmp_path / "log") + "/", "--results_dir", str(tmp_path / "res") + "/"]
    r = subprocess.run([sys.executable, os.path.join(ROOT, "homography_CNN_real.py"), "--mode", "train", "--max_iterations", "12"] + common,
                       capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "Finetune from" in r.stdout
    assert not re.search(r"\bh_loss\b", [ln for ln in r.stdout.splitlines() if ln.startswith("Train")][-1])     # gt-less: no h_loss
    ck = tfc.latest_checkpoint(os.path.join(save_dir, "l1_loss_normalize"))
    assert ck and ck.endswith("model.ckpt-11")                                 # the step was reset to 0 by --finetune
    v = tfc.read_checkpoint(ck)
    assert int(v["Variable"]) == 12
    w0 = eng.named_parameters()["model/conv_block4/conv2/weights"].cpu().numpy()
    assert np.abs(v["model/conv_block4/conv2/weights"] - w0).max() > 0          # it trained (from the restored weights) ...
    assert np.abs(v["model/conv_block4/conv2/weights"] - w0).max() < 12 * 1.1e-4   # ... by at most lr per step
    r = subprocess.run([sys.executable, os.path.join(ROOT, "homography_CNN_real.py"), "--mode", "test", "--do_augment", "0"] + common,
                       capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    lines = r.stdout.splitlines()
    i = next(k for k, ln in enumerate(lines) if ln.startswith("|Steps"))
    step, h_loss, l1_loss, fail_pct = [float(x) for x in lines[i + 1].split()]
    assert int(step) == 7 and 0 < h_loss < 30 and 0 < l1_loss < 2 and 0 <= fail_pct <= 100
