"""Generate the golden fixtures under tests/golden/ — run in the BUILD container only
(`python tests/golden/make_golden.py`), where /root/reference is mounted.  Nothing here runs on the GPU box.

What is pinned against the REFERENCE'S OWN CODE (imported / exec'd from /root/reference, never copied):
  dlt_golden.npz   A, b assembled with the reference's Aux_M* literals (code/utils/utils.py:11-122, exec'd from
                   the file text because importing utils.py needs TensorFlow and a tty) exactly as
                   code/homography_model.py:223-238 does, and H from cv2.getPerspectiveTransform
                   (the reference's own ground-truth routine, code/utils/gen_synthetic_data.py:56).
  warp_golden.npz  outputs of the reference's NumPy spatial transformer
                   (code/utils/numpy_spatial_transformer.py:12-132, imported with stub matplotlib/skimage modules;
                   float output, i.e. before the uint8 cast at :131) on seeded images / homographies.
What is a regression pin of the oracle itself (no runnable reference exists for it — TF1 graph):
  e2e_golden.npz   fp32 oracle outputs (pred_h4p, H, pred_I2, six losses, test metrics, one Adam step digest)
                   on make_batch(seed) inputs and params.init_flat_large(seed) weights (the large-output parity recipe:
                   |pred_h4p| is tens of pixels, so a 1e-3 px / 3e-5-relative tolerance actually constrains the regressor).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/code"

from oracle import oracle as O                                            # noqa: E402
from unsuperviseddeephomographyral2018_b200 import params as P            # noqa: E402


def load_reference_numpy_transformer():
    for name in ("matplotlib", "matplotlib.pyplot", "skimage", "skimage.io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["skimage"].io = sys.modules["skimage.io"]
    sys.path.insert(0, os.path.join(REF, "utils"))
    import numpy_spatial_transformer as nst
    return nst


def load_reference_aux():
    lines = open(os.path.join(REF, "utils", "utils.py")).read().split("\n")
    ns = {"np": np}
    exec("\n".join(lines[9:123]), ns)                                      # the Aux_M* literal block only
    return ns


def reference_dlt_system(aux, pts1, pts2):
    """A, b exactly as code/homography_model.py:223-238 builds them, in fp64 NumPy."""
    p1 = pts1.reshape(8, 1); p2 = pts2.reshape(8, 1)
    A1 = aux["Aux_M1"] @ p1; A2 = aux["Aux_M2"] @ p1; A3 = aux["Aux_M3"]
    A4 = aux["Aux_M4"] @ p1; A5 = aux["Aux_M5"] @ p1; A6 = aux["Aux_M6"]
    A7 = (aux["Aux_M71"] @ p2) * (aux["Aux_M72"] @ p1)
    A8 = (aux["Aux_M71"] @ p2) * (aux["Aux_M8"] @ p1)
    A = np.stack([a.reshape(8) for a in (A1, A2, A3, A4, A5, A6, A7, A8)], axis=0).T
    b = aux["Aux_Mb"] @ p2
    return A, b


def main():
    import cv2
    rng = np.random.default_rng(2018)
    aux = load_reference_aux()
    nst = load_reference_numpy_transformer()

    # ---------------- DLT ----------------
    n = 32
    x0 = rng.integers(45, 148, size=n); y0 = rng.integers(45, 68, size=n)
    pts1 = np.stack([x0, y0, x0 + 128, y0, x0 + 128, y0 + 128, x0, y0 + 128], 1).astype(np.float64)
    h4p = rng.integers(-45, 46, size=(n, 8)).astype(np.float64)
    h4p[n // 2:] += rng.normal(0, 2.0, size=(n - n // 2, 8))
    A_ref = np.zeros((n, 8, 8)); b_ref = np.zeros((n, 8, 1)); H_cv = np.zeros((n, 3, 3)); H_ref = np.zeros((n, 3, 3))
    for i in range(n):
        A_ref[i], b_ref[i] = reference_dlt_system(aux, pts1[i], pts1[i] + h4p[i])
        H_ref[i] = np.append(np.linalg.solve(A_ref[i], b_ref[i]).reshape(8), 1.0).reshape(3, 3)
        H_cv[i] = cv2.getPerspectiveTransform(pts1[i].reshape(4, 2).astype(np.float32),
                                              (pts1[i] + h4p[i]).reshape(4, 2).astype(np.float32))
    A_o, b_o = O.dlt_system(torch.tensor(pts1), torch.tensor(pts1 + h4p))
    H_o = O.solve_dlt(torch.tensor(pts1), torch.tensor(h4p)).numpy()
    assert np.array_equal(A_o.numpy(), A_ref) and np.array_equal(b_o.numpy(), b_ref), "oracle A,b != reference Aux assembly"
    print("DLT  oracle vs reference-Aux solve  max|dH| = %.3e" % np.abs(H_o - H_ref).max())
    # cv2 needs float32 points; the first half of the set is integer-valued so that is exact
    print("DLT  oracle vs cv2.getPerspectiveTransform (integer pts) max|dH| = %.3e" % np.abs(H_o[:n // 2] - H_cv[:n // 2]).max())
    assert np.abs(H_o - H_ref).max() < 1e-9 and np.abs(H_o[:n // 2] - H_cv[:n // 2]).max() < 1e-9
    np.savez_compressed(os.path.join(HERE, "dlt_golden.npz"), pts1=pts1, h4p=h4p, A_ref=A_ref, b_ref=b_ref,
                        H_ref=H_ref, H_cv2=H_cv[:n // 2])

    # ---------------- warp vs the reference NumPy twin ----------------
    cases = {}
    # (a) small multi-channel images with strong homographies (plenty of out-of-range samples)
    Hs, Ws = 24, 32
    img_s = rng.uniform(0, 255, size=(4, Hs, Ws, 3))
    # M float32 and inv(M) float32, exactly as numpy_spatial_transformer.py:138-141 / homography_model.py:63-70
    M_s = np.array([[Ws / 2.0, 0, Ws / 2.0], [0, Hs / 2.0, Hs / 2.0], [0, 0, 1.0]]).astype(np.float32)
    th_s, out_s = [], []
    for i in range(4):
        c = np.array([[4, 3], [26, 3], [26, 20], [4, 20]], dtype=np.float32)
        d = (c + rng.uniform(-6, 6, size=(4, 2))).astype(np.float32)
        Hpix = cv2.getPerspectiveTransform(c, d).astype(np.float64)
        theta = np.linalg.inv(M_s) @ Hpix @ M_s
        grid = nst._meshgrid(Hs, Ws)
        T = theta @ grid
        # per channel: the twin's 3-D branch relies on an old-NumPy expand_dims leniency (numpy_spatial_transformer.py:86-91)
        out = np.stack([nst._interpolate(img_s[i][..., c], T[0] / T[2], T[1] / T[2], [Hs, Ws]).reshape(Hs, Ws)
                        for c in range(3)], axis=2)
        th_s.append(theta); out_s.append(out)
    th_s, out_s = np.stack(th_s), np.stack(out_s)
    o_s, _ = O.transformer(torch.tensor(img_s), torch.tensor(th_s), (Hs, Ws))
    d_small = np.abs(o_s.numpy() - out_s).max()
    print("warp oracle vs reference NumPy twin (24x32x3, 4 cases)  max|d| = %.3e" % d_small)
    assert d_small < 1e-9
    cases.update(small_img=img_s, small_theta=th_s, small_out=out_s)

    # (b) full 240x320 batch from make_batch: H = H_gt, patch window of the twin's output vs oracle.transform
    batch = O.make_batch(7, 2, dtype=torch.float64)
    Hh, W = 240, 320
    M = np.array([[W / 2.0, 0, W / 2.0], [0, Hh / 2.0, Hh / 2.0], [0, 0, 1.0]]).astype(np.float32)
    win = []
    for b in range(2):
        theta = np.linalg.inv(M) @ batch["H_gt"][b].numpy() @ M
        grid = nst._meshgrid(Hh, W)
        T = theta @ grid
        Ib = batch["I_aug"][b].numpy()
        full = np.stack([nst._interpolate(Ib[..., c], T[0] / T[2], T[1] / T[2], [Hh, W]).reshape(Hh, W)
                         for c in range(3)], axis=2)
        win.append(full.mean(axis=2).reshape(-1)[batch["patch_indices"][b].numpy()].reshape(128, 128))
    win = np.stack(win)
    o_w = O.transform(batch["I_aug"], batch["H_gt"], batch["patch_indices"], 128).numpy()[..., 0]
    o_c = O.warp_closed_form(batch["I_aug"], batch["H_gt"], batch["pts1"][:, 0].numpy(), batch["pts1"][:, 1].numpy(), 128, 128).numpy()[..., 0]
    print("warp oracle.transform vs reference NumPy twin (240x320 window)  max|d| = %.3e" % np.abs(o_w - win).max())
    print("warp closed form      vs reference NumPy twin (240x320 window)  max|d| = %.3e" % np.abs(o_c - win).max())
    assert np.abs(o_w - win).max() < 1e-9 and np.abs(o_c - win).max() < 1e-4   # closed form drops the fp32 inv(M) rounding
    cases.update(full_seed=np.array(7), full_window=win.astype(np.float64),
                 full_I_u8_crc=np.array(int(batch["I_u8"].astype(np.int64).sum())))
    np.savez_compressed(os.path.join(HERE, "warp_golden.npz"), **cases)

    # ---------------- end-to-end fp32 oracle pin ----------------
    torch.manual_seed(0)
    specs = P.param_specs()
    e2e = {}
    for seed in (0, 1):
        B = 2
        flat = torch.tensor(P.init_flat_large(seed))      # |pred_h4p| of tens of px: pixel-unit tolerances are not vacuous
        batch = O.make_batch(seed, B)
        params = P.unflatten(flat, specs)
        out = O.forward(params, batch, None, mode="test")
        for k in ("pred_h4p", "H_mat", "pred_I2", "h_loss", "rec_loss", "ssim_loss", "l1_loss", "l1_smooth_loss",
                  "ncc_loss", "bounded_h_loss", "num_fail", "batch_h_loss"):
            e2e["s%d_%s" % (seed, k)] = out[k].detach().numpy()
        # Mean-corner-error pin that depends on the prediction: with random weights every sample "fails" against the true
        # gt (bounded_h_loss then equals the identity error whatever the regressor outputs), so the metric is ALSO pinned
        # against a synthetic gt within a few px of the oracle's prediction (gt enters the metrics only, not the network).
        gt_m = torch.tensor(np.round(out["pred_h4p"].detach().numpy() + np.random.default_rng(seed + 31).normal(0, 4.0, size=(B, 8))).astype(np.float32))
        lm = O.losses(out["pred_h4p"].detach(), gt_m, out["pred_I2"].detach(), batch["I2_aug"])
        tm = O.test_metrics(out["pred_h4p"].detach(), gt_m)
        e2e["s%d_gt_metric" % seed] = gt_m.numpy()
        e2e["s%d_h_loss_m" % seed] = lm["h_loss"].numpy()
        for k in ("bounded_h_loss", "num_fail", "batch_h_loss"):
            e2e["s%d_%s_m" % (seed, k)] = tm[k].numpy()
        assert float(tm["num_fail"]) == 0.0
        for lt in ("h_loss", "l1_loss"):
            newp, m, v, _, g = O.train_step(flat, torch.zeros_like(flat), torch.zeros_like(flat), 0, batch, specs,
                                            loss_type=lt, lr=5e-4 if lt == "h_loss" else 1e-4)
            # digest: per-tensor gradient L2 norms and a strided sample of the update
            e2e["s%d_%s_gradnorm" % (seed, lt)] = np.array([g[s.offset:s.offset + s.size].norm().item() for s in specs.values()])
            e2e["s%d_%s_dp_sample" % (seed, lt)] = (newp - flat)[::100003].numpy()
            e2e["s%d_%s_g_sample" % (seed, lt)] = g[::100003].numpy()
        e2e["s%d_input_digest" % seed] = np.array([batch["I_aug"].double().sum().item(), batch["I2_aug"].double().sum().item(),
                                                   batch["pts1"].double().sum().item(), batch["gt"].double().sum().item()])
    np.savez_compressed(os.path.join(HERE, "e2e_golden.npz"), **e2e)
    for f in ("dlt_golden.npz", "warp_golden.npz", "e2e_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
