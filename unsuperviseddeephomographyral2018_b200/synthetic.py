"""On-device synthetic pair generation.

Follows the reference's offline generator code/utils/gen_synthetic_data.py:40-68 — random patch corner in
[rho, W-rho-P] x [rho, Hh-rho-P], corner perturbation U{-rho..rho}^8, H = 4-point homography, I' = warp of I with the
reference's spatial-transformer convention, uint8 cast — and the dataloader's post-processing
(code/dataloader.py:99-100,172-177,203-227: normalise both images with I's statistics, gray = channel mean, patch
gather by patch_indices).  MS-COCO is not available offline, so I is a seeded band-limited random texture.
The homography and the warp run through libudh (udh_dlt_fwd, udh_transformer_fwd); torch only moves bytes.
"""
import ctypes

import torch
import torch.nn.functional as F

from . import ops
from ._lib import check, lib

MEAN_I = (118.93, 113.97, 102.60)
STD_I = (69.85, 68.81, 72.45)


def _texture(gen, B, Hh, W, device):
    """Seeded multi-octave random texture, uint8 NHWC: random grids at 1/1 ... 1/32 resolution, bicubically upsampled and
    summed with amplitude proportional to their scale (roughly the 1/f spectrum of natural images), so photometric losses
    see structure at the scale of the rho = 45 px displacements as well as fine detail."""
    img = torch.zeros(B, 3, Hh, W, device=device)
    for o in range(6):
        s = 2 ** o
        h, w = max(2, -(-Hh // s) + 1), max(2, -(-W // s) + 1)
        g = torch.rand(B, 3, h, w, device=device, generator=gen) - 0.5
        if o == 0:
            g = F.avg_pool2d(F.pad(g, (1, 1, 1, 1), mode="reflect"), 3, 1)[..., :Hh, :W]
        else:
            g = F.interpolate(g, size=(h * s, w * s), mode="bicubic", align_corners=False)[..., :Hh, :W]
        img = img + g * float(s) ** 0.9
    lo = img.amin(dim=(1, 2, 3), keepdim=True); hi = img.amax(dim=(1, 2, 3), keepdim=True)
    return ((img - lo) / (hi - lo) * 255.0).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous()   # NHWC uint8


def normalise(img_u8):
    mean = torch.tensor(MEAN_I, device=img_u8.device, dtype=torch.float32)
    std = torch.tensor(STD_I, device=img_u8.device, dtype=torch.float32)
    return ((img_u8.to(torch.float32) - mean) / std).contiguous()


def make_batch(B, seed=0, img_h=240, img_w=320, patch=128, rho=45, device="cuda"):
    """Post-dataloader tensors of one batch, on the device (same keys as the reference Dataloader's *_batch attributes)."""
    gen = torch.Generator(device=device); gen.manual_seed(int(seed))
    I_u8 = _texture(gen, B, img_h, img_w, device)
    x0 = torch.randint(rho, img_w - rho - patch + 1, (B,), device=device, generator=gen)
    y0 = torch.randint(rho, img_h - rho - patch + 1, (B,), device=device, generator=gen)
    pts1 = torch.stack([x0, y0, x0 + patch, y0, x0 + patch, y0 + patch, x0, y0 + patch], dim=1).to(torch.float32).contiguous()
    gt = torch.randint(-rho, rho + 1, (B, 8), device=device, generator=gen).to(torch.float32).contiguous()
    H_gt = ops.dlt_forward(pts1, gt)
    # theta = M^-1 H M (numpy_spatial_transformer.py:135-146 with H = inv(H_inverse))
    M = torch.tensor([[img_w / 2.0, 0., img_w / 2.0], [0., img_h / 2.0, img_h / 2.0], [0., 0., 1.]], device=device)
    theta = (torch.linalg.inv(M) @ H_gt @ M).contiguous()
    Ip, _ = ops.transformer(I_u8.to(torch.float32).contiguous(), theta, (img_h, img_w))
    Ip_u8 = Ip.clamp(0, 255).to(torch.uint8)                         # numpy_spatial_transformer.py:131 (uint8 cast)
    I_n, Ip_n = normalise(I_u8), normalise(Ip_u8)
    yy, xx = torch.meshgrid(torch.arange(patch, device=device), torch.arange(patch, device=device), indexing="ij")
    idx = ((yy[None] + y0[:, None, None]) * img_w + (xx[None] + x0[:, None, None])).reshape(B, -1)
    gray_I, gray_Ip = I_n.mean(dim=3).reshape(B, -1), Ip_n.mean(dim=3).reshape(B, -1)
    I1 = torch.gather(gray_I, 1, idx).reshape(B, patch, patch, 1).contiguous()
    I2 = torch.gather(gray_Ip, 1, idx).reshape(B, patch, patch, 1).contiguous()
    return dict(I1=I1, I2=I2, I1_aug=I1, I2_aug=I2, I_aug=I_n, I_prime_aug=Ip_n, pts1=pts1, gt=gt,
                patch_indices=idx.to(torch.int32).contiguous(), I_u8=I_u8, I_prime_u8=Ip_u8)


def draw_augmentation(B, do_augment, mode, generator, device):
    """Per-sample parameters of the reference's photometric augmentation (code/dataloader.py:163-169,323-375) as the
    [B,11] table udh_prep_inputs_u8_ex takes: {on, gamma, brightness, colour RGB for I, then the same five for I'}.
    A sample is augmented when u > 1 - do_augment (:167,169); train = JOINT (one draw for the pair), test = DISJOINT."""
    u = torch.rand(B, device=device, generator=generator)
    def five():
        g = torch.rand(B, 5, device=device, generator=generator)
        lo = torch.tensor([0.8, 0.5, 0.8, 0.8, 0.8], device=device); hi = torch.tensor([1.2, 2.0, 1.2, 1.2, 1.2], device=device)
        return lo + g * (hi - lo)
    a = five()
    b = a if mode == 'train' else five()
    return torch.cat([(u > (1.0 - do_augment)).float().unsqueeze(1), a, b], dim=1).contiguous()


def make_batch_fast(B, seed=0, img_h=240, img_w=320, patch=128, rho=45, device="cuda", do_augment=0.0, mode='train', want_rgb=False):
    """The same post-dataloader tensors as make_batch, produced by four kernel launches (udh_synth_scene_u8, udh_dlt_fwd,
    udh_warp_image_u8, udh_prep_inputs_u8_ex) instead of ~40 eager torch ops, with the reference's photometric augmentation
    (joint in train, disjoint in test) applied on the device when do_augment > 0.  `I_aug` is the GRAY warp source
    [B,Hh,W,1] (the channel mean commutes with the warp's linear sampling); want_rgb=True also returns the 3-channel tensor
    of the reference contract as `I_aug_rgb`.  `patch_indices` holds the window origin of each sample ([B] int32): the
    kernels only ever read the first gathered index (dataloader.py:203-207)."""
    dev = torch.device(device)
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    st = ops._stream()
    I_u8 = torch.empty(B, img_h, img_w, 3, device=dev, dtype=torch.uint8)
    Ip_u8 = torch.empty_like(I_u8)
    pts1 = torch.empty(B, 8, device=dev); gt = torch.empty(B, 8, device=dev)
    check(lib.udh_synth_scene_u8(p(I_u8), p(pts1), p(gt), B, img_h, img_w, patch, rho, int(seed) & 0xFFFFFFFFFFFFFFFF, st), "udh_synth_scene_u8")
    H_gt = ops.dlt_forward(pts1, gt)
    check(lib.udh_warp_image_u8(p(I_u8), p(H_gt), p(Ip_u8), B, img_h, img_w, st), "udh_warp_image_u8")
    aug = None
    if do_augment > 0:
        gen = torch.Generator(device=dev); gen.manual_seed(int(seed) * 7919 + 13)
        aug = draw_augmentation(B, do_augment, mode, gen, dev)
    out = prep_u8(I_u8, Ip_u8, pts1, aug, patch, want_rgb=want_rgb)
    out.update(gt=gt, I_u8=I_u8, I_prime_u8=Ip_u8, aug=aug)
    return out


def prep_u8(I_u8, Ip_u8, pts1, aug, patch=128, want_rgb=False, want_plain=True):
    """udh_prep_inputs_u8_ex on device tensors: augment + normalise + gray + patch gather in one pass."""
    B, img_h, img_w = I_u8.shape[0], I_u8.shape[1], I_u8.shape[2]
    dev = I_u8.device
    p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    I_gray = torch.empty(B, img_h, img_w, 1, device=dev)
    I_rgb = torch.empty(B, img_h, img_w, 3, device=dev) if want_rgb else None
    I1a = torch.empty(B, patch, patch, 1, device=dev); I2a = torch.empty_like(I1a)
    I1 = torch.empty_like(I1a) if want_plain else None
    I2 = torch.empty_like(I1a) if want_plain else None
    origin = torch.empty(B, device=dev, dtype=torch.int32)
    check(lib.udh_prep_inputs_u8_ex(p(I_u8), p(Ip_u8), p(pts1), p(aug), p(I_gray), p(I_rgb), p(I1), p(I2), p(I1a), p(I2a), p(origin),
                                    B, img_h, img_w, patch, ops._stream()), "udh_prep_inputs_u8_ex")
    out = dict(I1=I1 if want_plain else I1a, I2=I2 if want_plain else I2a, I1_aug=I1a, I2_aug=I2a, I_aug=I_gray, I_prime_aug=None,
               pts1=pts1, patch_indices=origin)
    if want_rgb:
        out["I_aug_rgb"] = I_rgb
    return out
