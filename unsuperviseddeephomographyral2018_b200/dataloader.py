"""Input pipeline with the reference Dataloader's contract (code/dataloader.py:11-22,76-235).

Produces, per batch, the nine post-dataloader tensors the model consumes (`*_batch` attributes / dict keys):
I1, I2, I1_aug, I2_aug [B,P,P,1]; I (=I_aug), I_prime (=I_prime_aug) [B,Hh,W,3]; pts1, gt [B,8]; patch_indices [B,P*P].
JPEG decode and list parsing run on the host (as in the reference: queue-runner threads); photometric augmentation,
normalisation, gray conversion and patch gather run on the device.  `synthetic=N` replaces the on-disk set by N
on-device generated pairs (synthetic.make_batch) — MS-COCO is not available offline.
"""
import os
from collections import namedtuple

import numpy as np
import torch

from . import synthetic

dataloader_params = namedtuple('parameters',
                               'data_path,'
                               'filenames_file,'
                               'pts1_file,'
                               'gt_file,'
                               'mode,'
                               'batch_size,'
                               'img_h,'
                               'img_w,'
                               'patch_size,'
                               'augment_list,'
                               'do_augment,')


def count_text_lines(path):
    """utils/utils.py:358-362."""
    with open(path, 'r') as f:
        return len(f.readlines())


def read_img_and_gt(filenames_file, pts1_file, gt_file):
    """dataloader.py:49-72: list lines "<I name> <I' name>", pts1 / gt = 8 floats per line (np.savetxt)."""
    with open(pts1_file, 'r') as f:
        pts1 = np.array([ln.split() for ln in f.read().strip().split("\n")]).astype('float64')
    with open(filenames_file, 'r') as f:
        names = [ln.split() for ln in f.read().strip().split("\n")]
    if not gt_file:
        return names, pts1, None
    with open(gt_file, 'r') as f:
        gt = np.array([ln.split() for ln in f.read().strip().split("\n")]).astype('float64')
    return names, pts1, gt


class Dataloader(object):
    def __init__(self, params, shuffle=True, synthetic_pairs=0, seed=0, device="cuda", rho=45):
        self.params = params
        self.mode = params.mode
        self.shuffle = shuffle
        self.device = torch.device(device)
        self.synthetic_pairs = int(synthetic_pairs)
        self.seed = seed
        self.rho = rho
        self._rng = np.random.default_rng(seed)
        self._step = 0
        if not self.synthetic_pairs:
            self.names, self.pts1, self.gt = read_img_and_gt(params.filenames_file, params.pts1_file, params.gt_file)
            self.num_samples = len(self.names)
            self._order = np.arange(self.num_samples)
            self._cursor = self.num_samples
        else:
            self.num_samples = self.synthetic_pairs

    # ---- on-disk path -------------------------------------------------------------------------------------------
    def _read_pair(self, idx):
        import cv2
        name = self.names[idx][1] if len(self.names[idx]) > 1 else self.names[idx][0]   # dataloader.py:143-144 (column 1 for both)
        imgs = []
        for sub in ("I", "I_prime"):
            im = cv2.imread(os.path.join(self.params.data_path, sub, name), cv2.IMREAD_COLOR)
            if im is None:
                raise IOError("cannot read %s" % os.path.join(self.params.data_path, sub, name))
            im = cv2.cvtColor(im, cv2.COLOR_BGR2RGB)
            if im.shape[0] != self.params.img_h or im.shape[1] != self.params.img_w:
                im = cv2.resize(im, (self.params.img_w, self.params.img_h), interpolation=cv2.INTER_AREA)    # dataloader.py:245
            imgs.append(im)
        return imgs

    def _next_indices(self, B):
        out = []
        while len(out) < B:
            if self._cursor >= self.num_samples:
                if self.shuffle:
                    self._rng.shuffle(self._order)
                self._cursor = 0
            out.append(self._order[self._cursor]); self._cursor += 1
        return np.array(out)

    def _disk_batch(self):
        p = self.params
        B, P, Hh, W = p.batch_size, p.patch_size, p.img_h, p.img_w
        idx = self._next_indices(B)
        I = np.empty((B, Hh, W, 3), np.uint8); Ip = np.empty_like(I)
        for j, i in enumerate(idx):
            I[j], Ip[j] = self._read_pair(i)
        dev = self.device
        I_u8 = torch.from_numpy(I).to(dev); Ip_u8 = torch.from_numpy(Ip).to(dev)
        if 'normalize' not in p.augment_list:
            raise NotImplementedError("augment_list without 'normalize' is not supported (the reference default is ['normalize'])")
        # augmentation parameters on the host RNG (reproducible order), applied by the fused device kernel
        aug = np.zeros((B, 11), np.float32)
        aug[:, 0] = self._rng.uniform(0, 1, size=B) > (1 - p.do_augment)                          # dataloader.py:163-169
        draw = lambda: np.concatenate([[self._rng.uniform(0.8, 1.2), self._rng.uniform(0.5, 2.0)], self._rng.uniform(0.8, 1.2, size=3)])
        for j in range(B):
            a = draw()
            aug[j, 1:6] = a
            aug[j, 6:11] = a if self.mode == 'train' else draw()                                   # joint (train) / disjoint (test)
        pts1 = torch.tensor(self.pts1[idx], dtype=torch.float32, device=dev)
        gt = torch.tensor(self.gt[idx], dtype=torch.float32, device=dev) if self.gt is not None else None
        out = synthetic.prep_u8(I_u8, Ip_u8, pts1, torch.tensor(aug, device=dev), P, want_rgb=True)
        out["gt"] = gt
        out["I_aug"] = out.pop("I_aug_rgb")                        # the reference contract: [B,Hh,W,3] normalised (augmented) I
        return out

    # ---- public ---------------------------------------------------------------------------------------------------
    def next_batch(self):
        p = self.params
        if self.synthetic_pairs:
            # on-device generator + fused augment / normalise / gray / crop (four kernel launches per batch); photometric
            # augmentation as the reference: probability do_augment, joint in train, disjoint in test
            b = synthetic.make_batch_fast(p.batch_size, seed=self.seed * 1000003 + self._step, img_h=p.img_h, img_w=p.img_w,
                                          patch=p.patch_size, rho=self.rho, device=self.device, do_augment=float(p.do_augment),
                                          mode=self.mode)
        else:
            b = self._disk_batch()
        self._step += 1
        # the reference's attribute names
        self.I1_batch, self.I2_batch, self.I1_aug_batch, self.I2_aug_batch = b["I1"], b["I2"], b["I1_aug"], b["I2_aug"]
        self.I_batch, self.I_prime_batch = b["I_aug"], b["I_prime_aug"]
        self.pts1_batch, self.gt_batch, self.patch_indices_batch = b["pts1"], b["gt"], b["patch_indices"]
        return b
