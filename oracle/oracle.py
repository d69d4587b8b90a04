"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.

A line-by-line CPU restatement (PyTorch-CPU / NumPy, fp32 or fp64) of the reference's
unsupervised-homography hot path.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline`
/ `--impl reference` legs of `bench.py` may import this module; the product package
(`unsuperviseddeephomographyral2018_b200/`) never does and has no CPU fallback.

PARITY PINNING.  The arithmetic of this path lives in TensorFlow 1.x (`tensorflow-gpu==1.4.1`,
/root/reference/requirements.txt:28) + `tensorflow.contrib.slim`, which is not vendored under
/root/reference and cannot be installed here (no TF wheel, Python 3.12, no network), and the reference
ships no tests or golden vectors for it.  What pins this oracle instead (tests/golden/make_golden.py,
run in the build container where /root/reference exists, outputs committed under tests/golden/):
  * the warp is checked against the reference's own importable NumPy twin
    (code/utils/numpy_spatial_transformer.py:12-132) — fp64 agreement ~1e-11;
  * the DLT system is checked against A, b assembled from the reference's `Aux_M*` literals
    (code/utils/utils.py:11-122, exec'd from the file text) and against
    `cv2.getPerspectiveTransform`, which the reference itself uses as the same function
    (code/utils/gen_synthetic_data.py:56);
  * CNN / losses / Adam follow the cited lines with stock torch-CPU ops; the TF graph itself could not be
    executed, so for those rows parity is "restated, not pinned by a reference run" (see DESIGN.md).

Every function cites the reference lines it follows.  All tensors are NHWC like the reference.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# Row D — solve_DLT (code/homography_model.py:169-250, code/utils/utils.py:11-122)
# ----------------------------------------------------------------------------------------------

def dlt_system(pts1, pts2):
    """A [B,8,8], b [B,8,1] of the inhomogeneous 4-point DLT with h33 = 1.

    homography_model.py:223-238: columns A1..A8 are selector products of pts1 (x_i, y_i) and
    pts2 (u_i, v_i).  Per corner i:
        row 2i   = [ 0, 0, 0, -x, -y, -1,  v*x,  v*y | -v ]
        row 2i+1 = [ x, y, 1,  0,  0,  0, -u*x, -u*y |  u ]
    (A7 = (M71.pts2)*(M72.pts1): row 2i -> v*x, row 2i+1 -> u*(-x);  b = Mb.pts2: [-v, u]).
    """
    B = pts1.shape[0]
    p1 = pts1.reshape(B, 4, 2)
    p2 = pts2.reshape(B, 4, 2)
    x, y = p1[..., 0], p1[..., 1]
    u, v = p2[..., 0], p2[..., 1]
    z = torch.zeros_like(x)
    o = torch.ones_like(x)
    r0 = torch.stack([z, z, z, -x, -y, -o, v * x, v * y], dim=-1)      # rows 0,2,4,6
    r1 = torch.stack([x, y, o, z, z, z, -u * x, -u * y], dim=-1)       # rows 1,3,5,7
    A = torch.stack([r0, r1], dim=2).reshape(B, 8, 8)
    b = torch.stack([-v, u], dim=2).reshape(B, 8, 1)
    return A, b


def solve_dlt(pts1, h4p):
    """H [B,3,3] mapping pts1 -> pts1 + h4p in pixel units (homography_model.py:174-176,242-250).
    `tf.matrix_solve` = LU with partial pivoting; torch.linalg.solve is LAPACK getrf/getrs (same algorithm)."""
    pts2 = pts1 + h4p
    A, b = dlt_system(pts1, pts2)
    h8 = torch.linalg.solve(A, b)                                     # [B,8,1]
    ones = torch.ones(h8.shape[0], 1, 1, dtype=h8.dtype)
    return torch.cat([h8, ones], dim=1).reshape(-1, 3, 3)


# ----------------------------------------------------------------------------------------------
# Row W — transformer (code/utils/tf_spatial_transformer.py:18-251) and HomographyModel.transform
# ----------------------------------------------------------------------------------------------

def _linspace(start, stop, num, dtype):
    """TF LinSpace kernel: step = (stop-start)/(num-1); out[i] = start + step*i, all in T
    (tf_spatial_transformer.py:162-165 use tf.linspace(-1.0, 1.0, n))."""
    step = torch.tensor((stop - start), dtype=dtype) / torch.tensor(num - 1, dtype=dtype)
    i = torch.arange(num, dtype=dtype)
    return torch.tensor(start, dtype=dtype) + step * i


def meshgrid(height, width, dtype):
    """tf_spatial_transformer.py:141-180 with scale_h=True: rows x_t, y_t, ones -> [3, H*W]."""
    x_t = _linspace(-1.0, 1.0, width, dtype).reshape(1, width).expand(height, width)
    y_t = _linspace(-1.0, 1.0, height, dtype).reshape(height, 1).expand(height, width)
    return torch.stack([x_t.reshape(-1), y_t.reshape(-1), torch.ones(height * width, dtype=dtype)], dim=0)


def interpolate(im, x, y, return_cond=False):
    """tf_spatial_transformer.py:76-139.  im [B,H,W,C]; x, y flat [B*H_out*W_out] in [-1,1] units.
    return_cond: also return sum_k |w_k I_k| — the magnitude that cancels for out-of-range samples; fp32 results are
    only defined up to ~eps * that (tests scale their tolerance with it)."""
    B, H, W, C = im.shape
    dt = im.dtype
    n_per = x.numel() // B
    x = (x + 1.0) * float(W) / 2.0                                   # :97
    y = (y + 1.0) * float(H) / 2.0                                   # :98
    x0 = torch.floor(x).to(torch.int64)                              # :101
    x1 = x0 + 1
    y0 = torch.floor(y).to(torch.int64)
    y1 = y0 + 1
    x0 = x0.clamp(0, W - 1)                                          # :106-109
    x1 = x1.clamp(0, W - 1)
    y0 = y0.clamp(0, H - 1)
    y1 = y1.clamp(0, H - 1)
    base = (torch.arange(B, dtype=torch.int64) * (H * W)).repeat_interleave(n_per)   # :112 (_repeat)
    im_flat = im.reshape(-1, C)
    Ia = im_flat[base + y0 * W + x0]                                 # :114-127
    Ib = im_flat[base + y1 * W + x0]
    Ic = im_flat[base + y0 * W + x1]
    Id = im_flat[base + y1 * W + x1]
    x0f, x1f, y0f, y1f = x0.to(dt), x1.to(dt), y0.to(dt), y1.to(dt)  # weights from the CLIPPED corners
    wa = ((x1f - x) * (y1f - y)).unsqueeze(1)                        # :134-137
    wb = ((x1f - x) * (y - y0f)).unsqueeze(1)
    wc = ((x - x0f) * (y1f - y)).unsqueeze(1)
    wd = ((x - x0f) * (y - y0f)).unsqueeze(1)
    out = ((wa * Ia + wb * Ib) + wc * Ic) + wd * Id                  # :138 add_n
    if return_cond:
        return out, (wa * Ia).abs() + (wb * Ib).abs() + (wc * Ic).abs() + (wd * Id).abs()
    return out


def transformer(U, theta, out_size, return_cond=False):
    """tf_spatial_transformer.py:182-251.  U [B,H,W,C], theta [B,3,3] (normalised H'), out_size (H_out,W_out).
    Returns (output [B,H_out,W_out,C], condition)."""
    B, H, W, C = U.shape
    dt = U.dtype
    oh, ow = out_size
    grid = meshgrid(oh, ow, dt).unsqueeze(0).expand(B, 3, oh * ow)   # :206-210
    T_g = torch.matmul(theta.reshape(-1, 3, 3).to(dt), grid)         # :213
    x_s, y_s, t_s = T_g[:, 0, :].reshape(-1), T_g[:, 1, :].reshape(-1), T_g[:, 2, :].reshape(-1)
    small = torch.tensor(1e-7, dtype=dt)
    smallers = 1e-6 * (1.0 - (t_s.abs() >= small).to(dt))           # :230-231
    t_s = t_s + smallers                                             # :234
    condition = (t_s.abs() > small).to(dt).sum()                     # :235
    if return_cond:
        out, cond = interpolate(U, x_s / t_s, y_s / t_s, True)
        return out.reshape(B, oh, ow, C), cond.reshape(B, oh, ow, C)
    out = interpolate(U, x_s / t_s, y_s / t_s)                       # :239-242
    return out.reshape(B, oh, ow, C), condition


def norm_matrices(img_w, img_h, dtype):
    """homography_model.py:63-72: M (float32 literal) and its inverse (np.linalg.inv of the fp32 M)."""
    M = np.array([[img_w / 2.0, 0., img_w / 2.0], [0., img_h / 2.0, img_h / 2.0], [0., 0., 1.]]).astype(np.float32)
    M_inv = np.linalg.inv(M)
    return torch.tensor(M, dtype=dtype), torch.tensor(M_inv.astype(np.float32), dtype=dtype)


def transform(I, H_mat, patch_indices, patch_size, patch_w=None, return_cond=False):
    """homography_model.py:252-269: H' = M^-1 H M, full-grid warp of I [B,Hh,W,C], channel mean,
    flat gather of `patch_indices` [B,P*P] (+ b*Hh*W, :74-76) -> pred_I2 [B,P,P,1]."""
    B, Hh, W, C = I.shape
    M, M_inv = norm_matrices(W, Hh, I.dtype)
    Hn = torch.matmul(torch.matmul(M_inv.expand(B, 3, 3), H_mat), M.expand(B, 3, 3))     # :254
    idx = patch_indices.reshape(B, -1).to(torch.int64) + (torch.arange(B, dtype=torch.int64) * (Hh * W)).unsqueeze(1)
    if return_cond:
        warped, cond = transformer(I, Hn, (Hh, W), True)
        shape = (B, patch_size, patch_w or patch_size, 1)
        return warped.mean(dim=3).reshape(-1)[idx.reshape(-1)].reshape(shape), cond.mean(dim=3).reshape(-1)[idx.reshape(-1)].reshape(shape)
    warped, _ = transformer(I, Hn, (Hh, W))                                              # :257
    gray = warped.mean(dim=3).reshape(-1)                                                # :263-264
    return gray[idx.reshape(-1)].reshape(B, patch_size, patch_w or patch_size, 1)       # :267-269


def warp_closed_form(I, H_mat, x0, y0, pw, ph):
    """Window-only closed form of row W: out(i,j) = bilinear(I, H·(j·W/(W-1), i·Hh/(Hh-1), 1))
    with the reference's clip-then-weight rule.  Used only as an independent cross-check of `transform`."""
    B, Hh, W, C = I.shape
    dt = I.dtype
    outs = []
    for b in range(B):
        jj = (torch.arange(pw, dtype=dt) + float(x0[b])) * (W / (W - 1.0))
        ii = (torch.arange(ph, dtype=dt) + float(y0[b])) * (Hh / (Hh - 1.0))
        X, Y = jj.reshape(1, pw).expand(ph, pw), ii.reshape(ph, 1).expand(ph, pw)
        Hb = H_mat[b]
        den = Hb[2, 0] * X + Hb[2, 1] * Y + Hb[2, 2]
        sx = (Hb[0, 0] * X + Hb[0, 1] * Y + Hb[0, 2]) / den
        sy = (Hb[1, 0] * X + Hb[1, 1] * Y + Hb[1, 2]) / den
        fx0 = torch.floor(sx).to(torch.int64); fy0 = torch.floor(sy).to(torch.int64)
        cx0, cx1 = fx0.clamp(0, W - 1), (fx0 + 1).clamp(0, W - 1)
        cy0, cy1 = fy0.clamp(0, Hh - 1), (fy0 + 1).clamp(0, Hh - 1)
        img = I[b].mean(dim=2)
        wa = (cx1.to(dt) - sx) * (cy1.to(dt) - sy); wb = (cx1.to(dt) - sx) * (sy - cy0.to(dt))
        wc = (sx - cx0.to(dt)) * (cy1.to(dt) - sy); wd = (sx - cx0.to(dt)) * (sy - cy0.to(dt))
        outs.append(wa * img[cy0, cx0] + wb * img[cy1, cx0] + wc * img[cy0, cx1] + wd * img[cy1, cx1])
    return torch.stack(outs).unsqueeze(-1)


# ----------------------------------------------------------------------------------------------
# Row C — _vgg (code/homography_model.py:88-133)
# ----------------------------------------------------------------------------------------------

CONV_SCOPES = ["model/conv_block%d/conv%d" % (b, c) for b in (1, 2, 3, 4) for c in (1, 2)]


def vgg_forward(params, x, keep_masks=None, return_acts=False):
    """params: dict checkpoint-name -> tensor (HWIO conv kernels, [in,out] fc kernels).
    x: [B,P,P,2] NHWC.  keep_masks: None (test mode, keep_prob=1) or (mask_conv4 [B,P/8,P/8,128], mask_fc1 [B,1024])
    of {0,1}; kept activations are scaled by 1/keep_prob = 2 (slim.dropout, :119-121,129).
    conv = explicit zero pad 1 + VALID 3x3 s1 + bias + ReLU (:88-95); maxpool 2x2 s2 VALID after blocks 1-3 (:102-116)."""
    acts = OrderedDict()
    h = x.permute(0, 3, 1, 2)
    for i, scope in enumerate(CONV_SCOPES):
        w = params[scope + "/weights"].permute(3, 2, 0, 1)            # HWIO -> OIHW
        h = F.relu(F.conv2d(F.pad(h, (1, 1, 1, 1)), w, params[scope + "/biases"]))
        acts[scope] = h
        if i in (1, 3, 5):
            h = F.max_pool2d(h, 2, 2)
            acts["pool%d" % (i // 2 + 1)] = h
    h = h.permute(0, 2, 3, 1)                                         # NHWC for slim.flatten (:124)
    if keep_masks is not None:
        h = h * keep_masks[0].to(h.dtype) * 2.0
    flat = h.reshape(h.shape[0], -1)
    fc1 = F.relu(flat @ params["model/fc1/fc1/weights"] + params["model/fc1/fc1/biases"])     # :128
    acts["fc1"] = fc1
    if keep_masks is not None:
        fc1 = fc1 * keep_masks[1].to(h.dtype) * 2.0
    out = fc1 @ params["model/fc2/fc2/weights"] + params["model/fc2/fc2/biases"]              # :131
    return (out, acts) if return_acts else out


# ----------------------------------------------------------------------------------------------
# Row L — build_losses (code/homography_model.py:136-166,271-352)
# ----------------------------------------------------------------------------------------------

def l1_smooth_loss(x, y):
    d = (x - y).abs()                                                 # :136-139
    return torch.where(d < 1, 0.5 * d * d, d - 0.5).mean()


def ssim_map(x, y, size=3):
    """:141-158; x, y [B,P,P,1] -> clip((1-SSIM)/2, 0, 1) on the VALID 3x3 avg-pool grid."""
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    xn, yn = x.permute(0, 3, 1, 2), y.permute(0, 3, 1, 2)
    mu_x, mu_y = F.avg_pool2d(xn, size, 1), F.avg_pool2d(yn, size, 1)
    sig_x = F.avg_pool2d(xn * xn, size, 1) - mu_x ** 2
    sig_y = F.avg_pool2d(yn * yn, size, 1) - mu_y ** 2
    sig_xy = F.avg_pool2d(xn * yn, size, 1) - mu_x * mu_y
    n = (2 * mu_x * mu_y + C1) * (2 * sig_xy + C2)
    d = (mu_x ** 2 + mu_y ** 2 + C1) * (sig_x + sig_y + C2)
    return ((1 - n / d) / 2).clamp(0, 1)


def ncc_loss(x, y):
    lx, ly = x.square().sum().sqrt(), y.square().sum().sqrt()         # :161-166
    return (x / lx - y / ly).square().sum().sqrt()


def losses(pred_h4p, gt, pred_I2, I2):
    """All six scalars of build_losses (:285-352; which one carries gradient is the caller's choice)."""
    out = OrderedDict()
    if gt is not None:
        out["h_loss"] = (pred_h4p - gt).square().mean().sqrt()        # :288
    out["rec_loss"] = (pred_I2 - I2).square().mean().sqrt()           # :291
    out["ssim_loss"] = ssim_map(pred_I2, I2).mean()                   # :292-293
    out["l1_loss"] = (pred_I2 - I2).abs().mean()                      # :294 / :328
    out["l1_smooth_loss"] = l1_smooth_loss(pred_I2, I2)               # :295
    out["ncc_loss"] = ncc_loss(I2, pred_I2)                           # :296
    return out


def test_metrics(pred_h4p, gt):
    """Test-mode metric (:274-281): per-sample RMSE over the 8 corner coordinates, identity bound, failures."""
    batch_h_loss = (pred_h4p - gt).square().mean(dim=1).sqrt()
    h_loss_identity = gt.square().mean(dim=1).sqrt()
    is_failure = (batch_h_loss >= h_loss_identity).to(pred_h4p.dtype)
    num_fail = is_failure.sum()
    bounded = (batch_h_loss * (1 - is_failure) + is_failure * h_loss_identity).mean()
    ace = (pred_h4p - gt).reshape(-1, 4, 2).square().sum(dim=2).sqrt().mean()   # literature's mean corner distance
    return OrderedDict(batch_h_loss=batch_h_loss, h_loss_identity=h_loss_identity, num_fail=num_fail,
                       bounded_h_loss=bounded, ace=ace)


# ----------------------------------------------------------------------------------------------
# Row O / G — schedule, Adam, gradient mean (code/homography_CNN_synthetic.py:161-183,277-278)
# ----------------------------------------------------------------------------------------------

def decay_steps(lr, min_lr, num_total_steps=150000, decay_rate=0.96):
    return int((math.log(decay_rate) * num_total_steps) / math.log(min_lr * 1.0 / lr))      # :166,169


def learning_rate(step, lr, min_lr, num_total_steps=150000, decay_rate=0.96):
    """tf.train.exponential_decay(..., staircase=True) (:169)."""
    return lr * decay_rate ** (step // decay_steps(lr, min_lr, num_total_steps, decay_rate))


def adam_step(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """TF-1 AdamOptimizer (:183): alpha_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= alpha_t*m/(sqrt(v)+eps).  t is 1-based."""
    m = beta1 * m + (1 - beta1) * g
    v = beta2 * v + (1 - beta2) * g * g
    alpha = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    return p - alpha * m / (v.sqrt() + eps), m, v


def average_grads(tower_grads):
    """utils/utils.py:380-403: per-variable mean over towers."""
    return [torch.stack(gs, 0).mean(0) for gs in zip(*tower_grads)]


# ----------------------------------------------------------------------------------------------
# Whole model (HomographyModel.__init__ order: build_model -> solve_DLT -> transform -> build_losses)
# ----------------------------------------------------------------------------------------------

def forward(params, batch, keep_masks=None, mode="train"):
    """batch: dict with I1_aug, I2_aug [B,P,P,1], I_aug [B,Hh,W,3], pts1 [B,8], gt [B,8] or None,
    patch_indices [B,P*P] int.  Returns dict of the model's result attributes."""
    P = batch["I1_aug"].shape[1]
    x = torch.cat([batch["I1_aug"], batch["I2_aug"]], dim=3)          # :359
    pred_h4p = vgg_forward(params, x, keep_masks)
    H_mat = solve_dlt(batch["pts1"], pred_h4p)
    pred_I2 = transform(batch["I_aug"], H_mat, batch["patch_indices"], P)
    out = OrderedDict(pred_h4p=pred_h4p, H_mat=H_mat, pred_I2=pred_I2)
    out.update(losses(pred_h4p, batch.get("gt"), pred_I2, batch["I2_aug"]))
    if mode == "test" and batch.get("gt") is not None:
        out.update(test_metrics(pred_h4p, batch["gt"]))
    return out


def train_step(flat_params, m, v, step, batch, specs, loss_type="h_loss", lr=1e-4, min_lr=0.9e-4,
               keep_masks=None, grad_scale=1.0):
    """One optimiser step on a flat fp32/fp64 parameter vector (layout: package params.param_specs).
    `step` is the 0-based global step before the update.  Returns (new_flat, m, v, outputs, flat_grad)."""
    flat = flat_params.clone().requires_grad_(True)
    params = OrderedDict((n, flat[s.offset:s.offset + s.size].reshape(s.shape)) for n, s in specs.items())
    out = forward(params, batch, keep_masks, mode="train")
    out[loss_type].backward()
    g = flat.grad.detach() * grad_scale
    lr_t = learning_rate(step, lr, min_lr)
    with torch.no_grad():
        new_p, m, v = adam_step(flat.detach(), g, m, v, step + 1, lr_t)
    return new_p, m, v, OrderedDict((k, val.detach()) for k, val in out.items()), g


# ----------------------------------------------------------------------------------------------
# Synthetic inputs (distribution of code/utils/gen_synthetic_data.py:40-68, normalisation and
# patch indices of code/dataloader.py:99-100,172-177,203-227)
# ----------------------------------------------------------------------------------------------

MEAN_I = np.array([118.93, 113.97, 102.60], dtype=np.float32)
STD_I = np.array([69.85, 68.81, 72.45], dtype=np.float32)


def _texture(rng, B, Hh, W):
    """Band-limited random texture, uint8 [B,Hh,W,3] (stand-in for MS-COCO, which is not available offline)."""
    from scipy.ndimage import gaussian_filter
    n = rng.uniform(0, 1, size=(B, Hh, W, 3)).astype(np.float32)
    n = gaussian_filter(n, sigma=(0, 2.0, 2.0, 0.6))
    lo = n.min(axis=(1, 2, 3), keepdims=True); hi = n.max(axis=(1, 2, 3), keepdims=True)
    return np.clip((n - lo) / (hi - lo) * 255.0, 0, 255).astype(np.uint8)


def make_batch(seed, B, img_h=240, img_w=320, patch=128, rho=45, dtype=torch.float32, window=None):
    """Seeded post-dataloader tensors.  `window=(pw,ph,x0,y0)` overrides the patch window (config 4: whole image)."""
    rng = np.random.default_rng(seed)
    I_u8 = _texture(rng, B, img_h, img_w)
    x0 = rng.integers(rho, img_w - rho - patch + 1, size=B)           # gen_synthetic_data.py:42
    y0 = rng.integers(rho, img_h - rho - patch + 1, size=B)           # :43
    pts1 = np.stack([x0, y0, x0 + patch, y0, x0 + patch, y0 + patch, x0, y0 + patch], axis=1).astype(np.float32)  # :46-50
    gt = rng.integers(-rho, rho + 1, size=(B, 8)).astype(np.float32)  # :52-53
    H_gt = solve_dlt(torch.tensor(pts1, dtype=torch.float64), torch.tensor(gt, dtype=torch.float64))
    # I' = reference warp of I with H_gt (numpy_spatial_transformer.py:135-146: theta = M^-1 inv(H_inverse) M = M^-1 H M),
    # cast to uint8 like numpy_spatial_transformer.py:131
    M, M_inv = norm_matrices(img_w, img_h, torch.float64)
    theta = M_inv @ H_gt @ M
    Ip, _ = transformer(torch.tensor(I_u8, dtype=torch.float64), theta, (img_h, img_w))
    Ip_u8 = Ip.numpy().astype(np.uint8)
    I_n = (I_u8.astype(np.float32) - MEAN_I) / STD_I                  # dataloader.py:172-177 (both with I's stats)
    Ip_n = (Ip_u8.astype(np.float32) - MEAN_I) / STD_I
    if window is None:
        pw = ph = patch; wx0, wy0 = x0, y0
    else:
        pw, ph = window[0], window[1]
        wx0 = np.full(B, window[2]); wy0 = np.full(B, window[3])
    yy, xx = np.meshgrid(np.arange(ph), np.arange(pw), indexing="ij")
    idx = ((yy[None] + wy0[:, None, None]) * img_w + (xx[None] + wx0[:, None, None])).reshape(B, -1).astype(np.int32)  # dataloader.py:203-207
    gray_I, gray_Ip = I_n.mean(axis=3).reshape(B, -1), Ip_n.mean(axis=3).reshape(B, -1)
    I1 = np.take_along_axis(gray_I, idx.astype(np.int64), axis=1).reshape(B, ph, pw, 1)
    I2 = np.take_along_axis(gray_Ip, idx.astype(np.int64), axis=1).reshape(B, ph, pw, 1)
    t = lambda a: torch.tensor(a, dtype=dtype)
    return OrderedDict(I1=t(I1), I2=t(I2), I1_aug=t(I1), I2_aug=t(I2), I_aug=t(I_n), I_prime_aug=t(Ip_n),
                       pts1=t(pts1), gt=t(gt), patch_indices=torch.tensor(idx), I_u8=I_u8, I_prime_u8=Ip_u8,
                       H_gt=H_gt)


# ----------------------------------------------------------------------------------------------
# Input pipeline after JPEG decode (code/dataloader.py:163-177,203-227,323-375)
# ----------------------------------------------------------------------------------------------

def augment_image(img, gamma, brightness, colors, min_val=0.0, max_val=255.0):
    """One image of joint_/disjoint_augment_image_pair (:323-375): img ** gamma on the RAW 0..255 float values, then
    brightness, then the per-channel colour image, then clip_by_value."""
    x = img ** gamma
    x = x * brightness
    x = x * colors.reshape(1, 1, 3)
    return x.clamp(min_val, max_val)


def prep_inputs(I_u8, Ip_u8, pts1, aug=None, patch=128):
    """Post-dataloader tensors from decoded uint8 images (fp32, as the reference): optional augmentation (aug [B,11] =
    {on, gamma, brightness, colour RGB of I, the same five of I'}), normalisation of BOTH images with I's statistics
    (:173-177), gray = channel mean, patch gather at (x0, y0) = pts1[0:2] (:203-227)."""
    I = torch.as_tensor(I_u8).to(torch.float32); Ip = torch.as_tensor(Ip_u8).to(torch.float32)
    B, Hh, W, _ = I.shape
    mean, std = torch.tensor(MEAN_I), torch.tensor(STD_I)
    Ia, Ipa = I.clone(), Ip.clone()
    if aug is not None:
        for b in range(B):
            if float(aug[b, 0]) != 0.0:
                Ia[b] = augment_image(I[b], aug[b, 1], aug[b, 2], aug[b, 3:6])
                Ipa[b] = augment_image(Ip[b], aug[b, 6], aug[b, 7], aug[b, 8:11])
    norm = lambda t: (t - mean) / std
    I_n, Ip_n, Ia_n, Ipa_n = norm(I), norm(Ip), norm(Ia), norm(Ipa)
    x0 = pts1[:, 0].long(); y0 = pts1[:, 1].long()
    yy, xx = torch.meshgrid(torch.arange(patch), torch.arange(patch), indexing="ij")
    idx = ((yy[None] + y0[:, None, None]) * W + (xx[None] + x0[:, None, None])).reshape(B, -1)
    g = lambda t: torch.gather(t.mean(dim=3).reshape(B, -1), 1, idx).reshape(B, patch, patch, 1)
    return OrderedDict(I1=g(I_n), I2=g(Ip_n), I1_aug=g(Ia_n), I2_aug=g(Ipa_n), I_aug=Ia_n, I_prime_aug=Ipa_n,
                       patch_indices=idx.to(torch.int32))
