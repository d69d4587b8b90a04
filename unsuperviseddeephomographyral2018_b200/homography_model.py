"""Drop-in mirror of the reference's model object (code/homography_model.py:12-22,43-86).

    HomographyModel(args, I1, I2, I1_aug, I2_aug, I_aug, I_prime_aug, h4p, gt, patch_indices,
                    reuse_variables=None, model_index=0)

Same constructor parameters and result attributes as the TF1 class; tensors are eager CUDA tensors instead of graph
nodes and every attribute is computed by libudh's CUDA kernels through HomographyEngine.  `gt` may be None (the
reference silently skips h_loss then, homography_model.py:287-290).  Variables are shared between model instances the
way `reuse_variables=True` shares them in the reference: through a per-process engine registry keyed by shape.
"""
from collections import namedtuple

import torch

from . import _lib
from .engine import HomographyEngine

homography_model_params = namedtuple('parameters',
                                     'mode,'
                                     'batch_size,'
                                     'patch_size,'
                                     'img_w,'
                                     'img_h,'
                                     'loss_type,'
                                     'use_batch_norm,'
                                     'augment_list,'
                                     'leftright_consistent_weight,')

_ENGINES = {}


def get_engine(params, numeric="fp32", seed=0, reuse=None, **kw):
    key = (params.batch_size, params.patch_size, params.img_h, params.img_w, numeric)
    if reuse and key in _ENGINES:
        return _ENGINES[key]
    eng = HomographyEngine(params.batch_size, params.patch_size, params.img_h, params.img_w, numeric=numeric, seed=seed,
                           loss_type=params.loss_type, **kw)
    _ENGINES[key] = eng
    return eng


class HomographyModel(object):
    def __init__(self, args, I1, I2, I1_aug, I2_aug, I_aug, I_prime_aug, h4p, gt, patch_indices, reuse_variables=None,
                 model_index=0, engine=None, numeric="fp32", dropout_seed=None):
        if getattr(args, "use_batch_norm", False):
            raise _lib.UdhError("use_batch_norm is out of scope (the reference passes is_training as the BN decay, "
                                "homography_model.py:93); run with --use_batch_norm False")
        self.params = args
        self.mode = args.mode
        self.is_training = self.mode == 'train'
        self.I1, self.I2, self.I1_aug, self.I2_aug = I1, I2, I1_aug, I2_aug
        self.I, self.I_prime = I_aug, I_prime_aug            # homography_model.py:54-55
        self.pts_1, self.gt = h4p, gt
        self.patch_indices = patch_indices
        self.reuse_variables = reuse_variables
        self.model_collection = ['model_' + str(model_index)]
        self.engine = engine if engine is not None else get_engine(args, numeric=numeric, reuse=reuse_variables)
        self.engine.loss_type = args.loss_type
        batch = dict(I1_aug=I1_aug.contiguous(), I2_aug=I2_aug.contiguous(), I_aug=I_aug.contiguous(), pts1=h4p.contiguous(),
                     gt=None if gt is None else gt.contiguous(), patch_indices=patch_indices.contiguous())
        self._batch = batch
        # build_model -> solve_DLT -> transform -> build_losses (homography_model.py:81-85), eagerly
        out = self.engine.forward(batch, train=self.is_training, dropout_seed=dropout_seed)
        self._out = out
        self.pred_h4p, self.H_mat, self.pred_I2 = out["pred_h4p"], out["H_mat"], out["pred_I2"]
        pl = out["photo_losses"]
        self.rec_loss, self.ssim_loss, self.l1_loss = pl[_lib.L_REC], pl[_lib.L_SSIM], pl[_lib.L_L1]
        self.l1_smooth_loss, self.ncc_loss = pl[_lib.L_L1_SMOOTH], pl[_lib.L_NCC]
        if gt is not None:
            m = out["h4p_metrics"]
            self.h_loss = m[_lib.M_H_LOSS]
            if self.mode == 'test':                          # homography_model.py:274-281
                self.bounded_h_loss, self.num_fail = m[_lib.M_BOUNDED_H_LOSS], m[_lib.M_NUM_FAIL]
                self.batch_h_loss = out["batch_h_loss"]
            self.ace = m[_lib.M_ACE]

    def selected_loss(self):
        return getattr(self, self.params.loss_type)

    def compute_gradients(self):
        """opt_step.compute_gradients(selected loss) (homography_CNN_synthetic.py:258-269): fills engine.grads."""
        self.engine.backward(self._batch, self._out)
        return self.engine.named_gradients()

    def apply_gradients(self):
        """get_average_grads + apply_gradients (homography_CNN_synthetic.py:277-278)."""
        self.engine.allreduce_grads()
        return self.engine.update()
