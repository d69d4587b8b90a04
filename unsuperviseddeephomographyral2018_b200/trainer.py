"""Host-facing step API: the call a user of the reference's training loop makes once per step
(`sess.run([apply_grad_opt, losses...])`, code/homography_CNN_synthetic.py:335-345) with HOST tensors in, scalars out.

Inputs arrive in (preferably pinned) host memory as the reference's post-dataloader tensors; they are staged to the
device on a dedicated copy stream into a ring of device slots so the H2D copy of step k+1 overlaps the kernels of
step k, and the step's scalar results come back through one small pinned D2H copy per step.
"""
from collections import OrderedDict

import torch

from . import _lib
from .engine import HomographyEngine

_INPUT_KEYS = ("I1_aug", "I2_aug", "I_aug", "pts1", "gt")
_RESULT_NAMES = ("h_loss", "bounded_h_loss", "num_fail", "ace", "rec_loss", "ssim_loss", "l1_loss", "l1_smooth_loss", "ncc_loss")


def pin_batch(batch):
    """Copy a batch (CPU or CUDA tensors) into pinned host memory, keeping only what a step consumes."""
    out = {}
    for k in _INPUT_KEYS:
        if batch.get(k) is not None:
            out[k] = batch[k].detach().to("cpu").contiguous().pin_memory()
    pi = batch.get("patch_indices")
    if pi is not None:
        # the kernels only need the window origin = first gathered index of each sample (dataloader.py:203-207)
        out["patch_origin"] = pi.detach().to("cpu")[:, 0].contiguous().to(torch.int32).pin_memory()
    return out


def pin_batch_u8(I_u8, I_prime_u8, pts1, gt):
    """Host batch for HostStepper.step_u8: decoded uint8 images [B,Hh,W,3] + pts1/gt [B,8] in pinned memory."""
    f = lambda t, dt: torch.as_tensor(t).detach().to("cpu").to(dt).contiguous().pin_memory()
    return dict(I_u8=f(I_u8, torch.uint8), I_prime_u8=f(I_prime_u8, torch.uint8), pts1=f(pts1, torch.float32), gt=f(gt, torch.float32))


class HostStepper(object):
    def __init__(self, engine: HomographyEngine, depth=2):
        self.eng = engine
        self.depth = depth
        dev = engine.device
        B, P, Hh, W = engine.B, engine.Pz, engine.img_h, engine.img_w
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.slots = []
        for _ in range(depth):
            self.slots.append(dict(
                I1_aug=torch.empty(B, P, P, 1, device=dev), I2_aug=torch.empty(B, P, P, 1, device=dev),
                I_aug=torch.empty(B, Hh, W, 3, device=dev), pts1=torch.empty(B, 8, device=dev), gt=torch.empty(B, 8, device=dev),
                patch_indices=torch.empty(B, device=dev, dtype=torch.int32)))
        self.ready = [torch.cuda.Event() for _ in range(depth)]
        self.free = [torch.cuda.Event() for _ in range(depth)]
        self.done = [torch.cuda.Event() for _ in range(depth)]
        self.host_results = [torch.zeros(_lib.NMETRICS + _lib.NLOSSES, dtype=torch.float32).pin_memory() for _ in range(depth)]
        self.u8_slots = None
        self.i = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self._pending = None

    def _results(self, slot):
        self.done[slot].synchronize()
        r = self.host_results[slot].tolist()
        m, pl = r[:_lib.NMETRICS], r[_lib.NMETRICS:]
        return OrderedDict(zip(_RESULT_NAMES, m + [pl[_lib.L_REC], pl[_lib.L_SSIM], pl[_lib.L_L1], pl[_lib.L_L1_SMOOTH], pl[_lib.L_NCC]]))

    def step(self, host_batch, train=True):
        """Enqueue one step on `host_batch` (from pin_batch); returns the scalar results of the PREVIOUS step
        (None on the first call) so the host never stalls the pipeline; call flush() for the last one."""
        slot = self.i % self.depth
        dst = self.slots[slot]
        with torch.cuda.stream(self.copy_stream):
            if self.i >= self.depth:
                self.copy_stream.wait_event(self.free[slot])
            nbytes = 0
            for k in _INPUT_KEYS:
                dst[k].copy_(host_batch[k], non_blocking=True)
                nbytes += host_batch[k].numel() * host_batch[k].element_size()
            dst["patch_indices"].copy_(host_batch["patch_origin"], non_blocking=True)
            nbytes += host_batch["patch_origin"].numel() * 4
            self.ready[slot].record(self.copy_stream)
        self.h2d_bytes = nbytes
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready[slot])
        out = self.eng.train_step(dst) if train else self.eng.eval_step(dst)
        self.free[slot].record(cur)
        res = torch.cat([out["h4p_metrics"], out["photo_losses"]])
        self.host_results[slot].copy_(res, non_blocking=True)
        self.d2h_bytes = res.numel() * 4
        self.done[slot].record(cur)
        prev = self._pending
        self._pending = slot
        self.i += 1
        return self._results(prev) if prev is not None else None

    def step_u8(self, host_batch, train=True):
        """Same as step(), but the host hands over the DECODED uint8 images (what the reference dataloader holds before
        normalising): augmentation (host_batch["aug_dev"], optional [B,11] device tensor), normalisation, gray conversion
        and patch gather run on the device in one kernel (udh_prep_inputs_u8_ex)."""
        import ctypes
        from ._lib import check, lib
        eng = self.eng
        if self.u8_slots is None:
            B, Hh, W = eng.B, eng.img_h, eng.img_w
            self.u8_slots = [dict(I_u8=torch.empty(B, Hh, W, 3, device=eng.device, dtype=torch.uint8),
                                  I_prime_u8=torch.empty(B, Hh, W, 3, device=eng.device, dtype=torch.uint8),
                                  I_gray=torch.empty(B, Hh, W, 1, device=eng.device)) for _ in range(self.depth)]
        slot = self.i % self.depth
        dst, u8 = self.slots[slot], self.u8_slots[slot]
        with torch.cuda.stream(self.copy_stream):
            if self.i >= self.depth:
                self.copy_stream.wait_event(self.free[slot])
            u8["I_u8"].copy_(host_batch["I_u8"], non_blocking=True)
            u8["I_prime_u8"].copy_(host_batch["I_prime_u8"], non_blocking=True)
            dst["pts1"].copy_(host_batch["pts1"], non_blocking=True)
            dst["gt"].copy_(host_batch["gt"], non_blocking=True)
            self.ready[slot].record(self.copy_stream)
        self.h2d_bytes = 2 * host_batch["I_u8"].numel() + 2 * host_batch["pts1"].numel() * 4
        cur = torch.cuda.current_stream()
        cur.wait_event(self.ready[slot])
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
        # one fused pass: (augment,) normalise, gray, patch gather; the warp reads the GRAY plane (a third of the bytes of the
        # fp32 3-channel tensor, which is never materialised on this path)
        check(lib.udh_prep_inputs_u8_ex(p(u8["I_u8"]), p(u8["I_prime_u8"]), p(dst["pts1"]), p(host_batch.get("aug_dev")), p(u8["I_gray"]), None,
                                        None, None, p(dst["I1_aug"]), p(dst["I2_aug"]), p(dst["patch_indices"]), eng.B, eng.img_h,
                                        eng.img_w, eng.Pz, ctypes.c_void_p(cur.cuda_stream)), "udh_prep_inputs_u8_ex")
        dst = dict(dst, I_aug=u8["I_gray"])
        out = eng.train_step(dst) if train else eng.eval_step(dst)
        self.free[slot].record(cur)
        res = torch.cat([out["h4p_metrics"], out["photo_losses"]])
        self.host_results[slot].copy_(res, non_blocking=True)
        self.d2h_bytes = res.numel() * 4
        self.done[slot].record(cur)
        prev = self._pending
        self._pending = slot
        self.i += 1
        return self._results(prev) if prev is not None else None

    def flush(self):
        if self._pending is None:
            return None
        r = self._results(self._pending)
        self._pending = None
        return r
