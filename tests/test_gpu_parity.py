"""GPU parity tests (-m gpu): libudh's CUDA path, called through the C ABI, against the CPU oracle and the committed
golden vectors (tests/golden/, produced from the reference's own NumPy transformer / Aux_M* / cv2).

Tolerances (floating point; the reference computes in fp32):
  DLT H            : |dH| <= 2e-4 * max|H| on the golden set, <= 1e-3 worst case / 3e-5 median over 1000 random
                     samples and <= 4x the error of fp32 LAPACK (fp32 LU of a cond~5e5 system); corners reproject < 1e-2 px
  warp pred_I2     : |d| <= 2e-4 abs on >= 99.9 % of pixels (bilinear of unit-variance data, fp32 coordinates of
                     magnitude ~300 px; isolated pixels that straddle a clip boundary may differ), mean |d| <= 2e-5
  photometric loss : 1e-4 relative
  CNN activations  : 1e-5 * max|layer| per layer, pred_h4p 1e-5 * max|pred_h4p| (fp32 mode), on the LARGE-OUTPUT weights
                     (params.init_flat_large: |pred_h4p| of tens of pixels — on plain Xavier weights pred_h4p is ~0.01 px and
                     any pixel-unit tolerance is vacuous)
  gradients        : relative L2 error <= 2e-3 per tensor (fp32 atomics, different summation order; a ReLU / arg-max within
                     rounding of its threshold flips under any fp32 evaluation order, see tests/test_gpu_x3.py)
  mean corner error: |MCE_cuda - MCE_oracle| <= 1e-3 px  (BASELINE.json north_star), pinned both against the true gt and
                     against a synthetic gt a few px from the prediction (so that no sample hits the identity bound)
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O                                               # noqa: E402


@pytest.fixture(scope="module")
def udh():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import unsuperviseddeephomographyral2018_b200 as pkg
    from unsuperviseddeephomographyral2018_b200 import _lib, engine, ops, params
    _lib.require_device()

    class NS:
        pass
    ns = NS()
    ns.lib, ns.ops, ns.engine, ns.params, ns.pkg = _lib, ops, engine, params, pkg
    return ns


def dev(batch):
    return {k: (v.cuda().contiguous() if isinstance(v, torch.Tensor) else v) for k, v in batch.items()}


def rel_l2(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)


# ------------------------------------------------------------------------------------------------ DLT
def test_dlt_forward_golden(udh, golden_dir):
    g = np.load(os.path.join(golden_dir, "dlt_golden.npz"))
    pts1 = torch.tensor(g["pts1"], dtype=torch.float32).cuda()
    h4p = torch.tensor(g["h4p"], dtype=torch.float32).cuda()
    H = udh.ops.dlt_forward(pts1, h4p).cpu().double().numpy()
    Href = O.solve_dlt(pts1.cpu().double(), h4p.cpu().double()).numpy()      # same fp32-rounded inputs, fp64 solve
    scale = np.abs(Href).max(axis=(1, 2), keepdims=True)
    assert (np.abs(H - Href) / scale).max() < 2e-4
    assert (np.abs(H[:16] - g["H_cv2"]) / scale[:16]).max() < 2e-4           # integer-valued half == cv2 golden
    assert np.all(H[:, 2, 2] == 1.0)
    # corners map pts1 -> pts2 to well under a pixel hundredth
    p = np.concatenate([g["pts1"].reshape(-1, 4, 2), np.ones((32, 4, 1))], 2)
    q = np.einsum("bij,bkj->bki", H, p)
    err = np.abs(q[..., :2] / q[..., 2:] - (pts1.cpu().double().numpy() + h4p.cpu().double().numpy()).reshape(-1, 4, 2)).max()
    assert err < 5e-3


def test_dlt_edge_cases(udh):
    # B = 1, B not a multiple of the 4 warps per CTA, identity (h4p = 0), large batch
    for B in (1, 3, 5, 1000):
        rng = np.random.default_rng(B)
        x0 = rng.integers(45, 148, size=B); y0 = rng.integers(45, 68, size=B)
        pts1 = np.stack([x0, y0, x0 + 128, y0, x0 + 128, y0 + 128, x0, y0 + 128], 1).astype(np.float32)
        h4p = rng.uniform(-45, 45, size=(B, 8)).astype(np.float32)
        h4p[0] = 0
        H = udh.ops.dlt_forward(torch.tensor(pts1).cuda(), torch.tensor(h4p).cuda()).cpu().double()
        Href = O.solve_dlt(torch.tensor(pts1).double(), torch.tensor(h4p).double())
        scale = Href.abs().amax(dim=(1, 2), keepdim=True)
        err = ((H - Href).abs() / scale).amax(dim=(1, 2))
        # fp32 LU of a cond ~5e5 system: compare with what fp32 LAPACK (the reference's tf.matrix_solve class of
        # algorithm) achieves on the same inputs, and bound the worst case over 1000 random samples
        err32 = ((O.solve_dlt(torch.tensor(pts1), torch.tensor(h4p)).double() - Href).abs() / scale).amax(dim=(1, 2))
        assert err.max() < 1e-3 and err.median() < 3e-5
        assert err.max() <= 4 * err32.max() + 1e-5, (err.max(), err32.max())
        # what the warp consumes: corners reproject within 1e-2 px
        p = torch.cat([torch.tensor(pts1).double().reshape(B, 4, 2), torch.ones(B, 4, 1, dtype=torch.float64)], 2)
        q = torch.einsum("bij,bkj->bki", H, p)
        assert (q[..., :2] / q[..., 2:] - torch.tensor(pts1 + h4p).double().reshape(B, 4, 2)).abs().max() < 1e-2
        assert (H[0] - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-5


def test_dlt_backward(udh):
    rng = np.random.default_rng(11)
    B = 37
    x0 = rng.integers(45, 148, size=B); y0 = rng.integers(45, 68, size=B)
    pts1 = torch.tensor(np.stack([x0, y0, x0 + 128, y0, x0 + 128, y0 + 128, x0, y0 + 128], 1).astype(np.float32))
    h4p = torch.tensor(rng.uniform(-45, 45, size=(B, 8)).astype(np.float32))
    dH = torch.tensor(rng.normal(size=(B, 3, 3)).astype(np.float32))
    h64 = h4p.double().requires_grad_(True)
    H64 = O.solve_dlt(pts1.double(), h64)
    (H64 * dH.double()).sum().backward()
    Hc = udh.ops.dlt_forward(pts1.cuda(), h4p.cuda())
    got = udh.ops.dlt_backward(pts1.cuda(), h4p.cuda(), Hc, dH.cuda().contiguous()).cpu().double()
    assert rel_l2(got, h64.grad) < 2e-3
    # through the autograd.Function
    hc = h4p.cuda().requires_grad_(True)
    (udh.ops.solve_dlt(pts1.cuda(), hc) * dH.cuda()).sum().backward()
    assert rel_l2(hc.grad.cpu(), h64.grad) < 2e-3


# ------------------------------------------------------------------------------------------------ warp + losses
def _check_pred(pred, ref, frac=0.999, tol=2e-4, mean_tol=2e-5, cond=None):
    """cond = sum_k |w_k I_k| from the oracle: where taps cancel (out-of-range samples, near-horizon pixels of a strongly
    projective H) an fp32 result is only defined up to ~eps*cond, so the tolerance grows with it."""
    d = np.abs(np.asarray(pred, dtype=np.float64) - np.asarray(ref, dtype=np.float64))
    t = tol if cond is None else tol + 1e-6 * np.asarray(cond, dtype=np.float64)
    assert (d <= t).mean() >= frac, "only %.5f of pixels within tolerance (max %g)" % ((d <= t).mean(), d.max())
    if cond is None:
        assert d.mean() <= mean_tol, d.mean()
    else:
        assert np.median(d) <= mean_tol and (d / t).mean() <= 0.2


def test_warp_window_vs_reference_numpy_twin(udh, golden_dir):
    g = np.load(os.path.join(golden_dir, "warp_golden.npz"))
    batch = O.make_batch(int(g["full_seed"]), 2)
    b = dev(batch)
    H = batch["H_gt"].float().cuda().contiguous()
    pred, sums = udh.ops.warp_loss_forward(b["I_aug"], H, b["I2_aug"], b["patch_indices"], 128, 128)
    _check_pred(pred.cpu().numpy()[..., 0], g["full_window"])


def test_transformer_operator_vs_reference_numpy_twin(udh, golden_dir):
    g = np.load(os.path.join(golden_dir, "warp_golden.npz"))
    U = torch.tensor(g["small_img"], dtype=torch.float32).cuda()
    theta = torch.tensor(g["small_theta"], dtype=torch.float32).cuda()
    out, _ = udh.ops.transformer(U, theta, (24, 32))
    ref, _ = O.transformer(U.cpu(), theta.cpu(), (24, 32))                    # fp32 oracle on the same fp32 inputs
    d = (out.cpu() - ref).abs()
    assert (d <= 2e-2).float().mean() > 0.995 and d.mean() < 1e-3            # 0..255 data: 2e-2 abs = 1e-4 relative
    d64 = np.abs(out.cpu().double().numpy() - g["small_out"])
    assert (d64 <= 5e-2).mean() > 0.99


@pytest.mark.parametrize("seed,B", [(0, 2), (1, 2), (4, 5)])
def test_warp_and_photometric_losses_vs_oracle(udh, seed, B):
    batch = O.make_batch(seed, B)
    b = dev(batch)
    h4p = batch["gt"] + torch.tensor(np.random.default_rng(seed).normal(0, 2.0, size=(B, 8)).astype(np.float32))
    H = O.solve_dlt(batch["pts1"], h4p)
    ref_pred, cond = O.transform(batch["I_aug"], H, batch["patch_indices"], 128, return_cond=True)
    ref = O.losses(h4p, batch["gt"], ref_pred, batch["I2_aug"])
    Hc = H.cuda().contiguous()
    pred, sums = udh.ops.warp_loss_forward(b["I_aug"], Hc, b["I2_aug"], b["patch_indices"], 128, 128)
    _check_pred(pred.cpu().numpy(), ref_pred.numpy(), cond=cond.numpy())
    pl = udh.ops.photo_losses(pred, b["I2_aug"], sums, 128, 128, B).cpu().numpy()
    L = udh.lib
    # seed 1 holds a strongly projective H whose horizon (t_s = 0) crosses the window: a handful of pixels there are
    # numerically undefined in fp32 (see _check_pred) and dominate the reductions -> only finiteness is checked
    if float(cond.max()) >= 1e2:
        assert np.isfinite(pl[:5]).all()
        return
    rtol = 1e-4
    for name, slot in (("rec_loss", L.L_REC), ("ssim_loss", L.L_SSIM), ("l1_loss", L.L_L1), ("l1_smooth_loss", L.L_L1_SMOOTH),
                       ("ncc_loss", L.L_NCC)):
        assert abs(pl[slot] - ref[name].item()) <= rtol * abs(ref[name].item()) + 1e-6, (name, pl[slot], ref[name].item())


def test_warp_gray_input_and_full_grid(udh):
    """C = 1 path and config 4 (window = whole 240x320 grid, patch_indices = None)."""
    batch = O.make_batch(9, 2, window=(320, 240, 0, 0))
    gray = batch["I_aug"].mean(dim=3, keepdim=True).contiguous()
    H = batch["H_gt"].float()
    ref_pred, cond = O.transform(gray, H, batch["patch_indices"], 240, 320, return_cond=True)
    ref_l1 = (ref_pred - batch["I2_aug"]).abs().mean().item()
    pred, sums = udh.ops.warp_loss_forward(gray.cuda(), H.cuda().contiguous(), batch["I2_aug"].cuda().contiguous(), None, 320, 240)
    _check_pred(pred.cpu().numpy(), ref_pred.numpy(), cond=cond.numpy())
    oob = ref_pred.numpy() == 0.0                                            # exact cancellation outside the source image
    assert oob.mean() > 0.05 and (np.abs(pred.cpu().numpy()[oob]) <= 1e-6 * np.maximum(cond.numpy()[oob], 1.0)).all()
    l1 = sums[0].item() / (2 * 240 * 320)
    assert abs(l1 - ref_l1) <= 1e-4 * ref_l1


@pytest.mark.parametrize("loss_name", ["l1_loss", "rec_loss", "l1_smooth_loss", "ncc_loss", "ssim_loss"])
def test_warp_loss_backward_vs_fp64_autograd(udh, loss_name):
    B = 3
    batch = O.make_batch(21, B, dtype=torch.float64)
    h4p = batch["gt"] + torch.tensor(np.random.default_rng(3).normal(0, 1.5, size=(B, 8)))
    H = O.solve_dlt(batch["pts1"], h4p).detach().requires_grad_(True)
    pred = O.transform(batch["I_aug"], H, batch["patch_indices"], 128)
    O.losses(h4p, None, pred, batch["I2_aug"])[loss_name].backward()
    lt = {"l1_loss": udh.lib.LOSS_L1, "rec_loss": udh.lib.LOSS_REC, "l1_smooth_loss": udh.lib.LOSS_L1_SMOOTH, "ncc_loss": udh.lib.LOSS_NCC,
          "ssim_loss": udh.lib.LOSS_CUSTOM}[loss_name]
    b = dev({k: (v.float() if isinstance(v, torch.Tensor) and v.dtype == torch.float64 else v) for k, v in batch.items()})
    Hc = H.detach().float().cuda().contiguous()
    predc, sums = udh.ops.warp_loss_forward(b["I_aug"], Hc, b["I2_aug"], b["patch_indices"], 128, 128, want_pred=True)
    dpm = udh.ops.ssim_backward(predc, b["I2_aug"], 128, 128) if loss_name == "ssim_loss" else None
    if dpm is not None:                                                     # the per-pixel SSIM gradient itself vs fp64 autograd
        pr = pred.detach().clone().requires_grad_(True)
        O.ssim_map(pr, batch["I2_aug"]).mean().backward()
        assert rel_l2(dpm.cpu().double().reshape(-1), pr.grad.reshape(-1)) < 2e-3
    dH = udh.ops.warp_loss_backward(b["I_aug"], Hc, b["I2_aug"], b["patch_indices"], 128, 128, lt, sums, dpred=dpm).cpu().double()
    g = H.grad.clone(); g[:, 2, 2] = 0; dH[:, 2, 2] = 0                     # h33 is a constant downstream
    for i in range(B):
        assert rel_l2(dH[i], g[i]) < 5e-3, (i, dH[i], g[i])
    # and the chain down to h4p
    dh = udh.ops.dlt_backward(b["pts1"], h4p.float().cuda().contiguous(), Hc, dH.float().cuda().contiguous()).cpu().double()
    h2 = h4p.clone().requires_grad_(True)
    pred2 = O.transform(batch["I_aug"], O.solve_dlt(batch["pts1"], h2), batch["patch_indices"], 128)
    O.losses(h2, None, pred2, batch["I2_aug"])[loss_name].backward()
    assert rel_l2(dh, h2.grad) < 1e-2


# ------------------------------------------------------------------------------------------------ regressor
def _engine(udh, B, seed, **kw):
    """Engine on the large-output parity weights of `seed` (dropout seeds as HomographyEngine(seed=seed) would use)."""
    eng = udh.engine.HomographyEngine(B, seed=None, **kw)
    eng.load_flat(udh.params.init_flat_large(seed))
    eng.dropout_seed = 0x5EED0000 + seed
    return eng


def _oracle_params(udh, seed, dtype=torch.float32):
    flat = torch.tensor(udh.params.init_flat_large(seed)).to(dtype)
    return flat, udh.params.unflatten(flat, udh.params.param_specs())


@pytest.mark.parametrize("seed,B", [(0, 2), (1, 3)])
def test_cnn_forward_fp32_vs_oracle(udh, seed, B):
    eng = _engine(udh, B, seed)
    batch = O.make_batch(seed, B)
    out = eng.forward(dev(batch), train=False)
    flat, params = _oracle_params(udh, seed)
    x = torch.cat([batch["I1_aug"], batch["I2_aug"]], dim=3)
    ref, acts = O.vgg_forward(params, x, None, return_acts=True)
    names = list(acts.keys())   # conv1_1, conv1_2, pool1, conv2_1, conv2_2, pool2, ...
    order = {0: "model/conv_block1/conv1", 1: "model/conv_block1/conv2", 8: "pool1", 2: "model/conv_block2/conv1",
             3: "model/conv_block2/conv2", 9: "pool2", 4: "model/conv_block3/conv1", 5: "model/conv_block3/conv2", 10: "pool3",
             6: "model/conv_block4/conv1", 7: "model/conv_block4/conv2"}
    for layer, name in order.items():
        a = acts[name].permute(0, 2, 3, 1).contiguous().numpy()              # oracle conv acts are NCHW
        got = eng.activation(layer).cpu().numpy().reshape(a.shape)
        assert np.abs(got - a).max() <= 1e-5 * np.abs(a).max(), (name, np.abs(got - a).max() / np.abs(a).max())
    f1 = acts["fc1"].numpy()
    assert np.abs(eng.activation(11).cpu().numpy().reshape(B, 1024) - f1).max() <= 1e-5 * np.abs(f1).max()
    assert np.abs(ref.numpy()).max() > 10.0                                  # the fixture is not vacuous
    assert np.abs(out["pred_h4p"].cpu().numpy() - ref.numpy()).max() <= 1e-5 * np.abs(ref.numpy()).max()


def test_e2e_golden_and_mean_corner_error(udh, golden_dir):
    g = np.load(os.path.join(golden_dir, "e2e_golden.npz"))
    for seed in (0, 1):
        eng = _engine(udh, 2, seed)
        batch = O.make_batch(seed, 2)
        db = dev(batch)
        out = eng.forward(db, train=False)
        d = eng.losses_dict(out)
        ps = np.abs(g["s%d_pred_h4p" % seed]).max()
        assert ps > 10.0
        assert np.abs(out["pred_h4p"].cpu().numpy() - g["s%d_pred_h4p" % seed]).max() <= 1e-5 * ps
        Hs = np.abs(g["s%d_H_mat" % seed]).max()
        assert np.abs(out["H_mat"].cpu().numpy() - g["s%d_H_mat" % seed]).max() < 5e-4 * Hs
        _check_pred(out["pred_I2"].cpu().numpy(), g["s%d_pred_I2" % seed], tol=1e-3, mean_tol=1e-4)
        # mean corner error (reference definition: bounded_h_loss; also h_loss) within 1e-3 px
        assert abs(d["bounded_h_loss"] - float(g["s%d_bounded_h_loss" % seed])) <= 1e-3
        assert abs(d["h_loss"] - float(g["s%d_h_loss" % seed])) <= 1e-3
        assert d["num_fail"] == float(g["s%d_num_fail" % seed])
        for k in ("rec_loss", "ssim_loss", "l1_loss", "l1_smooth_loss", "ncc_loss"):
            assert abs(d[k] - float(g["s%d_%s" % (seed, k)])) <= 2e-4 * abs(float(g["s%d_%s" % (seed, k)])) + 1e-6, k
        per = out["batch_h_loss"].cpu().numpy()
        assert np.abs(per - g["s%d_batch_h_loss" % seed]).max() <= 1e-3
        # the same metric against a gt a few px from the prediction: no sample hits the identity bound, so
        # bounded_h_loss really is a function of the regressor's output
        db["gt"] = torch.tensor(g["s%d_gt_metric" % seed]).cuda()
        out = eng.forward(db, train=False)
        d = eng.losses_dict(out)
        assert d["num_fail"] == 0.0 == float(g["s%d_num_fail_m" % seed])
        assert abs(d["bounded_h_loss"] - float(g["s%d_bounded_h_loss_m" % seed])) <= 1e-3
        assert abs(d["h_loss"] - float(g["s%d_h_loss_m" % seed])) <= 1e-3
        assert np.abs(out["batch_h_loss"].cpu().numpy() - g["s%d_batch_h_loss_m" % seed]).max() <= 1e-3


@pytest.mark.parametrize("loss_type,lr", [("h_loss", 5e-4), ("l1_loss", 1e-4)])
def test_train_step_gradients_and_adam_vs_oracle(udh, loss_type, lr):
    seed, B = 0, 2
    specs = udh.params.param_specs()
    eng = _engine(udh, B, seed, loss_type=loss_type, lr=lr)
    batch = O.make_batch(seed, B)
    db = dev(batch)
    # train-mode forward draws the dropout masks on the device; the oracle reuses exactly those masks
    out = eng.forward(db, train=True)
    m1, m2 = eng.dropout_masks()
    keep = (m1.cpu().float(), m2.cpu().float())
    assert 0.45 < keep[0].mean() < 0.55 and 0.4 < keep[1].mean() < 0.6
    eng.backward(db, out)
    flat, _ = _oracle_params(udh, seed)
    newp, m, v, ref_out, g = O.train_step(flat, torch.zeros_like(flat), torch.zeros_like(flat), 0, batch, specs,
                                          loss_type=loss_type, lr=lr, keep_masks=keep)
    assert np.abs(out["pred_h4p"].cpu().numpy() - ref_out["pred_h4p"].numpy()).max() <= 1e-5 * ref_out["pred_h4p"].abs().max().item()
    got = eng.grads.cpu()
    for name, s in specs.items():
        a, r = got[s.offset:s.offset + s.size], g[s.offset:s.offset + s.size]
        # h_loss: the CNN backward alone.  l1_loss: the chain adds sign(pred-I2) flips at |d| ~ fp32 noise and an
        # ill-conditioned (cond ~5e5) transposed DLT solve, which the fp32 oracle itself only resolves to ~1e-2
        # flip-limited on the large-output weights (see the module docstring): measured 2e-4 .. 5e-3
        assert rel_l2(a, r) < (1e-2 if loss_type == "h_loss" else 2e-2), (name, rel_l2(a, r))
    eng.update()
    # TF-Adam's first step moves every touched weight by ~lr * sign(g): compare where the gradient is not ~0
    upd, ref_upd = (eng.params.cpu() - flat), (newp - flat)
    big = g.abs() > 1e-3 * g.abs().max()
    assert (upd[big] - ref_upd[big]).abs().max() <= 0.02 * lr
    assert eng.grads.abs().max().item() == 0.0                                # zero_grad fused into the update
    assert eng.global_step == 1


def test_adam_kernel_vs_oracle(udh):
    n = 4096 + 32
    rng = np.random.default_rng(0)
    p = torch.tensor(rng.normal(size=n).astype(np.float32)); g = torch.tensor(rng.normal(size=n).astype(np.float32))
    m = torch.tensor(rng.normal(size=n).astype(np.float32) * 0.1); v = torch.tensor(rng.uniform(0, 1, size=n).astype(np.float32))
    import math
    t, lr = 7, 3e-4
    alpha = lr * math.sqrt(1 - 0.999 ** t) / (1 - 0.9 ** t)
    rp, rm, rv = O.adam_step(p.double(), g.double() * 0.5, m.double(), v.double(), t, lr)
    pc, gc, mc, vc = p.cuda(), g.cuda(), m.cuda(), v.cuda()
    udh.ops.adam_step(pc, gc, mc, vc, alpha, grad_scale=0.5, zero_grad=False)
    assert (pc.cpu().double() - rp).abs().max() < 1e-6 and (mc.cpu().double() - rm).abs().max() < 1e-6
    assert (vc.cpu().double() - rv).abs().max() < 1e-6 and torch.equal(gc.cpu(), g)


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_properties_B128(udh):
    """BASELINE config sizes (B = 128): properties that need no full-size oracle run."""
    B = 128
    batch = O.make_batch(100, 4)
    rep = lambda t: t.repeat(B // 4, *([1] * (t.dim() - 1))).cuda().contiguous()
    db = {k: rep(v) for k, v in batch.items() if isinstance(v, torch.Tensor) and k != "H_gt"}
    eng = _engine(udh, B, 0)
    out = eng.forward(db, train=False)
    h = out["pred_h4p"]
    hs = h.abs().max().item()
    assert hs > 10.0
    # (1) batch invariance: replicated samples give identical predictions, and equal the B = 4 run
    # (fc1 is a split-K SGEMM with fp32 atomics: replicas agree to rounding, not bitwise)
    assert (h[:4] - h[4:8]).abs().max().item() < 2e-6 * hs and (h[:4] - h[-4:]).abs().max().item() < 2e-6 * hs
    eng4 = _engine(udh, 4, 0)
    out4 = eng4.forward(dev(batch), train=False)
    assert (out4["pred_h4p"] - h[:4]).abs().max().item() < 2e-6 * hs
    # (2) warp with the ground-truth homography reproduces I2 up to the uint8 cast of I'
    Hgt = udh.ops.dlt_forward(db["pts1"], db["gt"])
    pred, sums = udh.ops.warp_loss_forward(db["I_aug"], Hgt, db["I2_aug"], db["patch_indices"], 128, 128)
    assert sums[0].item() / (B * 128 * 128) < 1.0 / 69.0
    # (3) means are replication invariant
    d4, d128 = eng4.losses_dict(out4), eng.losses_dict(out)
    for k in ("l1_loss", "rec_loss", "ssim_loss", "l1_smooth_loss", "h_loss", "bounded_h_loss"):
        assert abs(d4[k] - d128[k]) <= 1e-5 * max(1.0, abs(d4[k])), k
    assert d128["num_fail"] == d4["num_fail"] * (B // 4)


# ------------------------------------------------------------------------------------------------ device-side inputs
def test_device_input_pipeline_and_synthetic_generator(udh):
    """udh_prep_inputs_u8 (normalise / gray / patch gather from uint8) == the oracle's post-dataloader tensors, and the
    on-device synthetic generator produces self-consistent pairs (warp with H_gt reproduces I2)."""
    import ctypes
    from unsuperviseddeephomographyral2018_b200 import synthetic, trainer
    batch = O.make_batch(12, 3)
    I_u8 = torch.tensor(batch["I_u8"]).cuda(); Ip_u8 = torch.tensor(batch["I_prime_u8"]).cuda()
    pts1 = batch["pts1"].cuda().contiguous()
    I_aug = torch.empty(3, 240, 320, 3, device="cuda"); I1 = torch.empty(3, 128, 128, device="cuda"); I2 = torch.empty_like(I1)
    origin = torch.empty(3, device="cuda", dtype=torch.int32)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    assert udh.lib.lib.udh_prep_inputs_u8(p(I_u8), p(Ip_u8), p(pts1), p(I_aug), p(I1), p(I2), p(origin), 3, 240, 320, 128, None) == 0
    torch.cuda.synchronize()
    assert (I_aug.cpu() - batch["I_aug"]).abs().max() < 1e-6
    assert (I1.cpu() - batch["I1_aug"][..., 0]).abs().max() < 1e-6 and (I2.cpu() - batch["I2_aug"][..., 0]).abs().max() < 1e-6
    assert torch.equal(origin.cpu(), batch["patch_indices"][:, 0])
    # on-device generator: same distribution rules, self-consistent pair
    sb = synthetic.make_batch(8, seed=3)
    assert sb["pts1"][:, 0].min() >= 45 and sb["pts1"][:, 0].max() <= 147 and sb["pts1"][:, 1].min() >= 45 and sb["pts1"][:, 1].max() <= 67
    assert sb["gt"].abs().max() <= 45 and torch.equal(sb["gt"], sb["gt"].round())
    Hgt = udh.ops.dlt_forward(sb["pts1"], sb["gt"])
    _, sums = udh.ops.warp_loss_forward(sb["I_aug"], Hgt, sb["I2_aug"], sb["patch_indices"], 128, 128, want_pred=False)
    assert sums[0].item() / (8 * 128 * 128) < 1.0 / 69.0
    # host stepper, both input variants, agree with a direct engine step on the same data
    eng = udh.engine.HomographyEngine(8, seed=0, loss_type="h_loss", lr=5e-4)
    st = trainer.HostStepper(eng)
    st.step(trainer.pin_batch(sb)); r_f32 = st.flush()
    eng2 = udh.engine.HomographyEngine(8, seed=0, loss_type="h_loss", lr=5e-4)
    st2 = trainer.HostStepper(eng2)
    st2.step_u8(trainer.pin_batch_u8(sb["I_u8"], sb["I_prime_u8"], sb["pts1"], sb["gt"])); r_u8 = st2.flush()
    eng3 = udh.engine.HomographyEngine(8, seed=0, loss_type="h_loss", lr=5e-4)
    d = eng3.losses_dict(eng3.train_step(sb))
    for k in ("h_loss", "l1_loss", "rec_loss"):
        assert abs(r_f32[k] - d[k]) <= 1e-5 * max(1, abs(d[k])) and abs(r_u8[k] - d[k]) <= 1e-4 * max(1, abs(d[k])), k
    assert st2.h2d_bytes < st.h2d_bytes / 2


def test_one_call_step_equals_separate_calls(udh):
    """udh_step_forward_backward (one C call per step) == forward() + backward() through the individual entry points."""
    for loss_type in ("h_loss", "l1_loss", "ssim_loss", "ncc_loss", "rec_loss", "l1_smooth_loss"):
        batch = dev(O.make_batch(2, 3))
        e1 = udh.engine.HomographyEngine(3, seed=1, loss_type=loss_type, lr=5e-4)
        e2 = udh.engine.HomographyEngine(3, seed=1, loss_type=loss_type, lr=5e-4)
        o1 = e1.forward(batch, train=True); e1.backward(batch, o1); g1 = e1.grads.clone(); d1 = e1.losses_dict(o1); e1.update()
        o2 = e2.train_step(batch); d2 = e2.losses_dict(o2)
        assert (o1["pred_h4p"] - o2["pred_h4p"]).abs().max().item() < 1e-6
        for k in d1:
            assert abs(d1[k] - d2[k]) <= 1e-5 * max(1.0, abs(d1[k])), k
        # TF-Adam's first step is ~lr*sign(g): compare where g is not rounding noise (fp32 atomics reorder between runs)
        big = g1.abs() > 1e-3 * g1.abs().max()
        assert (e1.params - e2.params)[big].abs().max().item() <= 0.02 * 5e-4 and e2.global_step == 1
        assert rel_l2(e2.adam_m.cpu(), e1.adam_m.cpu()) < 1e-4
        ev = e2.losses_dict(e2.eval_step(batch))
        assert abs(ev["l1_loss"] - e2.losses_dict(e2.forward(batch, train=False))["l1_loss"]) < 1e-6
