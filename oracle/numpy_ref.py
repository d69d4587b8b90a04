"""TEST INFRASTRUCTURE — second, independent restatement of rows C / L / O in plain NumPy fp64 (explicit loops over the
3x3 taps, no torch ops), used only by tests/test_oracle.py to cross-check oracle.py.  TensorFlow 1.x cannot run in this
environment, so rows C / L / O of the oracle cannot be pinned by a reference run ("parity unpinned" for them);
this file removes the single-implementation risk: two restatements written from the cited reference lines and from the
documented TF op semantics must agree to fp64 rounding.

Reference lines: code/homography_model.py:88-133 (_conv2d: explicit zero pad 1 then VALID 3x3, slim defaults bias + ReLU;
_maxpool2d 2x2 stride 2 VALID; fc1 ReLU, fc2 linear; NHWC flatten), :136-166 and :285-296 (losses), :274-283 (test
metrics); TF-1 AdamOptimizer as documented in tensorflow/python/training/adam.py (lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t),
m_t = b1 m + (1 - b1) g, v_t = b2 v + (1 - b2) g^2, var -= lr_t * m_t / (sqrt(v_t) + eps)).
"""
import numpy as np


def conv3x3_relu(x, w, b):
    """x [B,H,W,Cin], w HWIO [3,3,Cin,Cout], b [Cout] -> relu(conv) [B,H,W,Cout]; zero pad 1, stride 1."""
    B, H, W, Cin = x.shape
    xp = np.zeros((B, H + 2, W + 2, Cin), dtype=np.float64)
    xp[:, 1:H + 1, 1:W + 1] = x
    out = np.zeros((B, H, W, w.shape[3]), dtype=np.float64)
    for ky in range(3):
        for kx in range(3):
            out += xp[:, ky:ky + H, kx:kx + W, :] @ w[ky, kx]
    return np.maximum(out + b, 0.0)


def maxpool2x2(x):
    B, H, W, C = x.shape
    return x.reshape(B, H // 2, 2, W // 2, 2, C).max(axis=(2, 4))


def vgg_forward(params, x, keep_masks=None):
    """params: name -> ndarray (TF-Slim names, as oracle.vgg_forward).  x [B,P,P,2]."""
    h = np.asarray(x, dtype=np.float64)
    i = 0
    for blk in (1, 2, 3, 4):
        for c in (1, 2):
            s = "model/conv_block%d/conv%d" % (blk, c)
            h = conv3x3_relu(h, np.asarray(params[s + "/weights"], np.float64), np.asarray(params[s + "/biases"], np.float64))
            i += 1
        if blk < 4:
            h = maxpool2x2(h)
    if keep_masks is not None:
        h = h * np.asarray(keep_masks[0], np.float64) * 2.0
    flat = h.reshape(h.shape[0], -1)                                       # NHWC flatten
    fc1 = np.maximum(flat @ np.asarray(params["model/fc1/fc1/weights"], np.float64) + np.asarray(params["model/fc1/fc1/biases"], np.float64), 0.0)
    if keep_masks is not None:
        fc1 = fc1 * np.asarray(keep_masks[1], np.float64) * 2.0
    return fc1 @ np.asarray(params["model/fc2/fc2/weights"], np.float64) + np.asarray(params["model/fc2/fc2/biases"], np.float64)


def _avg3(a):
    """3x3 VALID average pooling of [B,P,P]."""
    P = a.shape[1]
    out = np.zeros((a.shape[0], P - 2, P - 2))
    for dy in range(3):
        for dx in range(3):
            out += a[:, dy:dy + P - 2, dx:dx + P - 2]
    return out / 9.0


def losses(pred_h4p, gt, pred_I2, I2):
    x = np.asarray(pred_I2, np.float64)[..., 0]
    y = np.asarray(I2, np.float64)[..., 0]
    d = x - y
    out = {}
    if gt is not None:
        out["h_loss"] = np.sqrt(np.mean((np.asarray(pred_h4p, np.float64) - np.asarray(gt, np.float64)) ** 2))
    out["rec_loss"] = np.sqrt(np.mean(d ** 2))
    mu_x, mu_y = _avg3(x), _avg3(y)
    sx, sy, sxy = _avg3(x * x) - mu_x ** 2, _avg3(y * y) - mu_y ** 2, _avg3(x * y) - mu_x * mu_y
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim = ((2 * mu_x * mu_y + C1) * (2 * sxy + C2)) / ((mu_x ** 2 + mu_y ** 2 + C1) * (sx + sy + C2))
    out["ssim_loss"] = np.mean(np.clip((1 - ssim) / 2, 0, 1))
    out["l1_loss"] = np.mean(np.abs(d))
    a = np.abs(d)
    out["l1_smooth_loss"] = np.mean(np.where(a < 1, 0.5 * a * a, a - 0.5))
    out["ncc_loss"] = np.sqrt(np.sum((y / np.sqrt(np.sum(y * y)) - x / np.sqrt(np.sum(x * x))) ** 2))
    return out


def test_metrics(pred_h4p, gt):
    p, g = np.asarray(pred_h4p, np.float64), np.asarray(gt, np.float64)
    bh = np.sqrt(np.mean((p - g) ** 2, axis=1))
    ident = np.sqrt(np.mean(g ** 2, axis=1))
    fail = (bh >= ident).astype(np.float64)
    return dict(batch_h_loss=bh, num_fail=fail.sum(), bounded_h_loss=np.mean(bh * (1 - fail) + fail * ident))


def adam_tf1(p, g, m, v, t, lr, b1=0.9, b2=0.999, eps=1e-8):
    lr_t = lr * np.sqrt(1.0 - b2 ** t) / (1.0 - b1 ** t)
    m = b1 * m + (1.0 - b1) * g
    v = b2 * v + (1.0 - b2) * g * g
    return p - lr_t * m / (np.sqrt(v) + eps), m, v
