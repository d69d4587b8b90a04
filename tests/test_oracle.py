"""CPU tests: the oracle against the committed golden vectors (tests/golden/, made by make_golden.py from the
reference's own NumPy transformer / Aux_M* literals / cv2.getPerspectiveTransform) and against itself."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O
from unsuperviseddeephomographyral2018_b200 import params as P


def test_dlt_matches_reference_aux_and_cv2(golden_dir):
    g = np.load(os.path.join(golden_dir, "dlt_golden.npz"))
    pts1, h4p = torch.tensor(g["pts1"]), torch.tensor(g["h4p"])
    A, b = O.dlt_system(pts1, pts1 + h4p)
    assert np.array_equal(A.numpy(), g["A_ref"]) and np.array_equal(b.numpy(), g["b_ref"])
    assert (A[:, 0, 0] == 0).all()                      # pivoting is mandatory
    H = O.solve_dlt(pts1, h4p).numpy()
    assert np.abs(H - g["H_ref"]).max() < 1e-9
    assert np.abs(H[:16] - g["H_cv2"]).max() < 1e-9
    # fp32 restatement stays close to fp64 (reference arithmetic is fp32)
    H32 = O.solve_dlt(pts1.float(), h4p.float()).double().numpy()
    assert np.abs(H32 - g["H_ref"]).max() / np.abs(g["H_ref"]).max() < 1e-4
    # H maps pts1 -> pts2
    p = torch.cat([pts1.reshape(-1, 4, 2), torch.ones(32, 4, 1, dtype=torch.float64)], 2)
    q = torch.einsum("bij,bkj->bki", torch.tensor(H), p)
    assert (q[..., :2] / q[..., 2:] - (pts1 + h4p).reshape(-1, 4, 2)).abs().max() < 1e-8


def test_warp_matches_reference_numpy_twin(golden_dir):
    g = np.load(os.path.join(golden_dir, "warp_golden.npz"))
    out, cond = O.transformer(torch.tensor(g["small_img"]), torch.tensor(g["small_theta"]), (24, 32))
    assert np.abs(out.numpy() - g["small_out"]).max() < 1e-9
    assert (np.abs(g["small_out"]) < 1e-9).mean() > 0.01          # the set does exercise out-of-range (~0) samples
    batch = O.make_batch(int(g["full_seed"]), 2, dtype=torch.float64)
    assert int(batch["I_u8"].astype(np.int64).sum()) == int(g["full_I_u8_crc"]), "synthetic input generator drifted"
    win = O.transform(batch["I_aug"], batch["H_gt"], batch["patch_indices"], 128).numpy()[..., 0]
    assert np.abs(win - g["full_window"]).max() < 1e-9
    cf = O.warp_closed_form(batch["I_aug"], batch["H_gt"], batch["pts1"][:, 0].numpy(), batch["pts1"][:, 1].numpy(), 128, 128)
    assert np.abs(cf.numpy()[..., 0] - g["full_window"]).max() < 1e-4


def test_warp_with_gt_homography_reproduces_I2():
    """End-to-end chain of SURVEY §8a-W: DLT(pts1, gt) -> warp -> window == I2 patch up to the uint8 cast of I'."""
    batch = O.make_batch(3, 2)
    H = O.solve_dlt(batch["pts1"], batch["gt"])
    pred = O.transform(batch["I_aug"], H, batch["patch_indices"], 128)
    d = (pred - batch["I2_aug"]).abs()
    assert d.mean() < 1.0 / 69.0                                     # < 1 grey level after normalisation


def test_e2e_golden_regression(golden_dir):
    g = np.load(os.path.join(golden_dir, "e2e_golden.npz"))
    specs = P.param_specs()
    assert P.num_parameters(specs) == 34192264
    for seed in (0, 1):
        flat = torch.tensor(P.init_flat_large(seed))
        batch = O.make_batch(seed, 2)
        dig = np.array([batch["I_aug"].double().sum().item(), batch["I2_aug"].double().sum().item(),
                        batch["pts1"].double().sum().item(), batch["gt"].double().sum().item()])
        assert np.allclose(dig, g["s%d_input_digest" % seed], rtol=1e-12)
        out = O.forward(P.unflatten(flat, specs), batch, None, mode="test")
        for k in ("pred_h4p", "H_mat", "h_loss", "l1_loss", "rec_loss", "ssim_loss", "l1_smooth_loss", "ncc_loss",
                  "bounded_h_loss", "num_fail"):
            ref = g["s%d_%s" % (seed, k)]
            assert np.allclose(out[k].detach().numpy(), ref, rtol=2e-4, atol=2e-5), k
        assert np.abs(out["pred_I2"].detach().numpy() - g["s%d_pred_I2" % seed]).max() < 1e-3
        assert np.abs(g["s%d_pred_h4p" % seed]).max() > 10.0              # large-output fixture: pixel tolerances are not vacuous
        tm = O.test_metrics(out["pred_h4p"].detach(), torch.tensor(g["s%d_gt_metric" % seed]))
        assert float(tm["num_fail"]) == 0.0
        assert abs(float(tm["bounded_h_loss"]) - float(g["s%d_bounded_h_loss_m" % seed])) < 1e-4


def test_second_restatement_of_rows_C_L_O_agrees_in_fp64():
    """oracle.py (torch ops) against oracle/numpy_ref.py (plain NumPy loops, written independently from the cited reference
    lines and the documented TF semantics): regressor forward incl. dropout masks, all six losses, test metrics, and three
    TF-1 Adam steps, in fp64.  Also pins Adam against a hand-worked vector of the documented ApplyAdam formula."""
    from oracle import numpy_ref as NR
    rng = np.random.default_rng(3)
    Pz, B = 16, 3                                                           # 16x16 patches: conv4 is 2x2x128, fc1 has 512 inputs
    params = {}
    for (blk, c, cin, cout) in [(1, 1, 2, 64), (1, 2, 64, 64), (2, 1, 64, 64), (2, 2, 64, 64), (3, 1, 64, 128), (3, 2, 128, 128),
                                (4, 1, 128, 128), (4, 2, 128, 128)]:
        s = "model/conv_block%d/conv%d" % (blk, c)
        params[s + "/weights"] = rng.normal(0, (2.0 / (9 * cin)) ** 0.5, size=(3, 3, cin, cout))
        params[s + "/biases"] = rng.normal(0, 0.1, size=cout)
    params["model/fc1/fc1/weights"] = rng.normal(0, 0.05, size=(2 * 2 * 128, 1024)); params["model/fc1/fc1/biases"] = rng.normal(0, 0.1, size=1024)
    params["model/fc2/fc2/weights"] = rng.normal(0, 0.5, size=(1024, 8)); params["model/fc2/fc2/biases"] = rng.normal(0, 5.0, size=8)
    x = rng.normal(size=(B, Pz, Pz, 2))
    keep = (rng.integers(0, 2, size=(B, 2, 2, 128)).astype(np.float64), rng.integers(0, 2, size=(B, 1024)).astype(np.float64))
    tp = {k: torch.tensor(v) for k, v in params.items()}
    for km in (None, keep):
        a = O.vgg_forward(tp, torch.tensor(x), None if km is None else (torch.tensor(km[0]), torch.tensor(km[1]))).numpy()
        b = NR.vgg_forward(params, x, km)
        assert np.abs(a).max() > 1.0 and np.abs(a - b).max() <= 1e-11 * np.abs(b).max()
    pred, gt = rng.normal(0, 20, size=(B, 8)), rng.integers(-45, 46, size=(B, 8)).astype(np.float64)
    pI, I2 = rng.normal(size=(B, 32, 32, 1)), rng.normal(size=(B, 32, 32, 1))
    la = O.losses(torch.tensor(pred), torch.tensor(gt), torch.tensor(pI), torch.tensor(I2))
    lb = NR.losses(pred, gt, pI, I2)
    for k, v in lb.items():
        assert abs(float(la[k]) - v) <= 1e-12 * max(1.0, abs(v)), k
    ta, tb = O.test_metrics(torch.tensor(pred), torch.tensor(gt)), NR.test_metrics(pred, gt)
    assert np.abs(ta["batch_h_loss"].numpy() - tb["batch_h_loss"]).max() < 1e-12
    assert float(ta["num_fail"]) == tb["num_fail"] and abs(float(ta["bounded_h_loss"]) - tb["bounded_h_loss"]) < 1e-12
    # three Adam steps, two restatements
    p = rng.normal(size=50); m = np.zeros(50); v = np.zeros(50)
    tp_, tm_, tv_ = torch.tensor(p), torch.zeros(50, dtype=torch.float64), torch.zeros(50, dtype=torch.float64)
    for t in (1, 2, 3):
        gr = rng.normal(size=50)
        p, m, v = NR.adam_tf1(p, gr, m, v, t, 5e-4)
        tp_, tm_, tv_ = O.adam_step(tp_, torch.tensor(gr), tm_, tv_, t, 5e-4)
        assert np.abs(tp_.numpy() - p).max() < 1e-15 and np.abs(tv_.numpy() - v).max() < 1e-18
    # hand-worked: p0 = 1, g = (0.5, -0.25, 1.0), lr = 0.1 (documented ApplyAdam formula, eps outside the bias correction)
    #   t=1: m = 0.05, v = 2.5e-4, lr_t = 0.1*sqrt(0.001)/0.1            -> p = 1 - 0.0316227766*0.05/(0.0158113883+1e-8)
    pw, mw, vw = np.array([1.0]), np.zeros(1), np.zeros(1)
    hand = []
    for t, gr in ((1, 0.5), (2, -0.25), (3, 1.0)):
        pw, mw, vw = NR.adam_tf1(pw, np.array([gr]), mw, vw, t, 0.1)
        hand.append(float(pw[0]))
    m1, v1 = 0.05, 2.5e-4
    p1 = 1.0 - (0.1 * np.sqrt(1 - 0.999) / (1 - 0.9)) * m1 / (np.sqrt(v1) + 1e-8)
    m2, v2 = 0.9 * m1 + 0.1 * -0.25, 0.999 * v1 + 0.001 * 0.0625
    p2 = p1 - (0.1 * np.sqrt(1 - 0.999 ** 2) / (1 - 0.81)) * m2 / (np.sqrt(v2) + 1e-8)
    m3, v3 = 0.9 * m2 + 0.1 * 1.0, 0.999 * v2 + 0.001 * 1.0
    p3 = p2 - (0.1 * np.sqrt(1 - 0.999 ** 3) / (1 - 0.729)) * m3 / (np.sqrt(v3) + 1e-8)
    assert np.allclose(hand, [p1, p2, p3], rtol=0, atol=1e-15)
    assert abs(p1 - 0.9000000632) < 1e-9                                    # ~ p0 - lr*sign(g): TF-Adam's first step


def test_schedule_and_adam():
    assert O.decay_steps(1e-4, 0.9e-4) == 58117 and O.decay_steps(5e-4, 0.9e-4) == 3570
    assert O.learning_rate(58116, 1e-4, 0.9e-4) == 1e-4
    assert abs(O.learning_rate(58117, 1e-4, 0.9e-4) - 0.96e-4) < 1e-18
    p, g = torch.tensor([1.0, -2.0]), torch.tensor([0.5, -0.25])
    p1, m, v = O.adam_step(p, g, torch.zeros(2), torch.zeros(2), 1, 1e-3)
    # first TF-Adam step moves every coordinate by ~lr against the gradient sign
    assert torch.allclose(p1, p - 1e-3 * torch.sign(g), atol=1e-6)


def test_fp64_gradient_chain():
    """DLT -> warp -> L1 is differentiable w.r.t. h4p and agrees with central differences."""
    batch = O.make_batch(5, 1, dtype=torch.float64)
    h = (batch["gt"] + 0.37).clone().requires_grad_(True)

    def f(hh):
        Hm = O.solve_dlt(batch["pts1"], hh)
        return (O.transform(batch["I_aug"], Hm, batch["patch_indices"], 128) - batch["I2_aug"]).abs().mean()
    f(h).backward()
    num = torch.zeros(8, dtype=torch.float64)
    for i in range(8):
        e = torch.zeros(1, 8, dtype=torch.float64); e[0, i] = 1e-5
        num[i] = (f(h.detach() + e) - f(h.detach() - e)) / 2e-5
    assert (h.grad[0] - num).abs().max() / num.abs().max() < 1e-4


def test_losses_and_metrics_hand_computed():
    """Row L on values small enough to do by hand (homography_model.py:136-166,274-296)."""
    pred = torch.tensor([0.0, 2.0, 0.5, -3.0, 0.0, 0.0, 0.0, 0.0, 0.0]).reshape(1, 3, 3, 1)    # 3x3: SSIM needs one window
    tgt = torch.zeros(1, 3, 3, 1)
    d = O.losses(None, None, pred, tgt)
    assert "h_loss" not in d                                                   # gt-less operation (the real-data trainer)
    assert abs(float(d["l1_loss"]) - (0 + 2 + 0.5 + 3) / 9) < 1e-7
    assert abs(float(d["rec_loss"]) - np.sqrt((0 + 4 + 0.25 + 9) / 9)) < 1e-6
    # Huber, delta = 1: 0.5 d^2 below 1, d - 0.5 above
    assert abs(float(d["l1_smooth_loss"]) - (0 + 1.5 + 0.125 + 2.5) / 9) < 1e-7
    # NCC = || x/||x|| - y/||y|| ||_2 over the whole batch tensor: identical directions -> 0, opposite -> 2
    x = torch.tensor([3.0, 4.0]).reshape(1, 1, 2, 1)
    assert float(O.ncc_loss(x, 2 * x)) < 1e-6 and abs(float(O.ncc_loss(x, -x)) - 2.0) < 1e-6
    # SSIM term of a constant image against itself is 0, against its negative it saturates at 1 after the clip
    c = torch.full((1, 5, 5, 1), 0.7, dtype=torch.float64)       # fp64: the variance terms cancel exactly
    assert float(O.ssim_map(c, c).max()) < 1e-6 and O.ssim_map(c, c).shape == (1, 1, 3, 3)
    assert abs(float(O.ssim_map(c, -c).mean()) - min(1.0, (1 - (-2 * 0.49 + 1e-4) / (2 * 0.49 + 1e-4)) / 2)) < 1e-5
    # test-mode metric: sample 0 beats the identity bound, sample 1 does not (>=) and is replaced by the bound
    gt = torch.tensor([[3.0] * 8, [1.0] * 8])
    ph = torch.tensor([[2.0] * 8, [3.0] * 8])
    m = O.test_metrics(ph, gt)
    assert torch.allclose(m["batch_h_loss"], torch.tensor([1.0, 2.0])) and torch.allclose(m["h_loss_identity"], torch.tensor([3.0, 1.0]))
    assert float(m["num_fail"]) == 1.0 and abs(float(m["bounded_h_loss"]) - (1.0 + 1.0) / 2) < 1e-7
    assert abs(float(m["ace"]) - (np.sqrt(2.0) + 2 * np.sqrt(2.0)) / 2) < 1e-6
    assert abs(float(O.losses(ph, gt, pred, tgt)["h_loss"]) - np.sqrt((8 * 1.0 + 8 * 4.0) / 16)) < 1e-6


def test_transformer_grid_and_border_quirks():
    """Row W quirks the CUDA kernels reproduce (utils/tf_spatial_transformer.py:97-139,162-177,230-240): with the identity the
    output pixel j samples x = j*W/(W-1) (linspace(-1,1,W) endpoint mismatch), so column 0 is exact, the last column falls on
    x = W and comes out as 0 (both taps clip to the same pixel with cancelling weights), and a vanishing t_s takes the 1e-6
    epsilon instead of dividing by zero."""
    Hh, W = 4, 5
    img = torch.arange(Hh * W, dtype=torch.float64).reshape(1, Hh, W, 1) + 1.0
    eye = torch.eye(3, dtype=torch.float64).reshape(1, 9)
    out, _ = O.transformer(img, eye, (Hh, W))
    assert torch.allclose(out[0, 0, 0, 0], img[0, 0, 0, 0])
    xs = torch.arange(W, dtype=torch.float64) * W / (W - 1)
    for j in range(W - 1):                               # interior columns: plain bilinear at x = j*W/(W-1) on row 0
        xj = float(xs[j]); x0 = int(np.floor(xj)); f = xj - x0
        want = (1 - f) * img[0, 0, x0, 0] + f * img[0, 0, min(x0 + 1, W - 1), 0] if x0 + 1 <= W - 1 else None
        if want is not None:
            assert abs(float(out[0, 0, j, 0]) - float(want)) < 1e-9
    assert abs(float(out[0, 0, W - 1, 0])) < 1e-9 and abs(float(out[0, Hh - 1, 0, 0])) < 1e-9     # x = W / y = Hh: the "black border"
    # t_s == 0 everywhere: the reference adds 1e-6 where |t_s| < 1e-7 -> finite output (zeros after the clip-cancellation)
    theta = torch.tensor([[1.0, 0, 0, 0, 1, 0, 0, 0, 0]], dtype=torch.float64)
    out0, _ = O.transformer(img, theta, (Hh, W))
    assert torch.isfinite(out0).all()


def test_average_grads_is_the_tower_mean():
    """utils/utils.py:380-403 — what allreduce(sum) / N reproduces."""
    t0 = [torch.tensor([1.0, 2.0]), torch.tensor([[4.0]])]
    t1 = [torch.tensor([3.0, 6.0]), torch.tensor([[0.0]])]
    avg = O.average_grads([t0, t1])
    assert torch.equal(avg[0], torch.tensor([2.0, 4.0])) and torch.equal(avg[1], torch.tensor([[2.0]]))
