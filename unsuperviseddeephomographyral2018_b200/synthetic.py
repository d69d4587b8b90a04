"""On-device synthetic pair generation.

Follows the reference's offline generator code/utils/gen_synthetic_data.py:40-68 — random patch corner in
[rho, W-rho-P] x [rho, Hh-rho-P], corner perturbation U{-rho..rho}^8, H = 4-point homography, I' = warp of I with the
reference's spatial-transformer convention, uint8 cast — and the dataloader's post-processing
(code/dataloader.py:99-100,172-177,203-227: normalise both images with I's statistics, gray = channel mean, patch
gather by patch_indices).  MS-COCO is not available offline, so I is a seeded band-limited random texture.
The homography and the warp run through libudh (udh_dlt_fwd, udh_transformer_fwd); torch only moves bytes.
"""
import torch
import torch.nn.functional as F

from . import ops

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
