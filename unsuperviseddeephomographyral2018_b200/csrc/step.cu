// One C call per training / evaluation step: the whole reference step graph (code/homography_CNN_synthetic.py:229-278,345)
// — regressor forward, h4p losses, DLT, fused warp + photometric diagnostics, backward of the selected loss — enqueued on
// one stream from C, so the host side costs one FFI call instead of ~20 (the reference does one sess.run per step).
#include <string.h>

#include "common.cuh"

extern "C" int udh_step_forward_backward(const udh_step_args* a, int phase, void* stream) {
  UDH_REQUIRE(a, "udh_step_forward_backward: null args");
  UDH_REQUIRE(phase == UDH_STEP_ALL || phase == UDH_STEP_FWD_HEAD || phase == UDH_STEP_CONVS || phase == UDH_STEP_FWD_ONLY,
              "udh_step_forward_backward: bad phase %d", phase);
  cudaStream_t st = udh::as_stream(stream);
  int rc;
#define STEP_TRY(call) do { rc = (call); if (rc != UDH_OK) return rc; } while (0)
  const int B = a->B, P = a->P;
  if (phase != UDH_STEP_CONVS) {
    UDH_REQUIRE(a->params && a->I1 && a->I2 && a->I_aug && a->pts1 && a->h4p && a->H && a->sums && a->photo_losses && a->ws,
                "udh_step_forward_backward: null pointer in args");
    STEP_TRY(udh_cnn_fwd_ex(a->params, a->I1, a->I2, a->h4p, a->ws, a->ws_bytes, B, P, a->train, a->seed, a->numeric_mode, a->fwd_flags, stream));
    const bool want_dpred = a->train && a->loss_type == UDH_STEP_LOSS_H && phase != UDH_STEP_FWD_ONLY;
    if (a->gt) STEP_TRY(udh_h4p_loss(a->h4p, a->gt, B, a->h4p_metrics, a->per_sample, want_dpred ? a->dh4p : nullptr, stream));
    STEP_TRY(udh_dlt_fwd(a->pts1, a->h4p, a->H, B, stream));
    UDH_CUDA(cudaMemsetAsync(a->sums, 0, sizeof(double) * UDH_NSUMS, st));
    STEP_TRY(udh_warp_loss_fwd(a->I_aug, a->C, a->img_h, a->img_w, a->H, a->I2, a->patch_indices, a->idx_stride, P, P, a->pred_I2,
                               a->sums, B, stream));
    if (a->pred_I2) STEP_TRY(udh_ssim_fwd(a->pred_I2, a->I2, P, P, a->sums, B, stream));
    STEP_TRY(udh_photo_losses_finalize(a->sums, (double)B * P * P, a->pred_I2 ? (double)B * (P - 2) * (P - 2) : 0.0, a->photo_losses, stream));
    if (phase == UDH_STEP_FWD_ONLY) return UDH_OK;
    UDH_REQUIRE(a->grads && a->dh4p, "udh_step_forward_backward: backward needs grads and dh4p buffers");
    if (a->loss_type != UDH_STEP_LOSS_H) {
      UDH_REQUIRE(a->dH && a->scratch, "udh_step_forward_backward: photometric backward needs dH and scratch");
      int lt;
      switch (a->loss_type) {
        case UDH_STEP_LOSS_L1: lt = UDH_LOSS_L1; break;
        case UDH_STEP_LOSS_REC: lt = UDH_LOSS_REC; break;
        case UDH_STEP_LOSS_L1_SMOOTH: lt = UDH_LOSS_L1_SMOOTH; break;
        case UDH_STEP_LOSS_NCC: lt = UDH_LOSS_NCC; break;
        case UDH_STEP_LOSS_SSIM: lt = UDH_LOSS_CUSTOM; break;
        default: udh::set_error("udh_step_forward_backward: unknown loss_type %d", a->loss_type); return UDH_EINVAL;
      }
      if (lt == UDH_LOSS_CUSTOM) {
        UDH_REQUIRE(a->pred_I2 && a->dpred_map, "udh_step_forward_backward: ssim_loss needs pred_I2 and dpred_map");
        STEP_TRY(udh_ssim_bwd(a->pred_I2, a->I2, P, P, a->dpred_map, B, stream));
      }
      STEP_TRY(udh_warp_loss_bwd_ex(a->I_aug, a->C, a->img_h, a->img_w, a->H, a->I2, a->patch_indices, a->idx_stride, P, P, lt, a->sums,
                                    lt == UDH_LOSS_CUSTOM ? a->dpred_map : nullptr, 1.0f, a->dH, a->scratch, B, stream));
      STEP_TRY(udh_dlt_bwd(a->pts1, a->h4p, a->H, a->dH, a->dh4p, B, stream));
    } else {
      UDH_REQUIRE(a->gt, "udh_step_forward_backward: h_loss needs gt");
    }
    STEP_TRY(udh_cnn_bwd_phase(a->params, a->I1, a->I2, a->dh4p, a->grads, a->ws, a->ws_bytes, B, P, a->train, a->numeric_mode,
                               phase == UDH_STEP_ALL ? UDH_BWD_ALL : UDH_BWD_HEAD, stream));
    return UDH_OK;
  }
  return udh_cnn_bwd_phase(a->params, a->I1, a->I2, a->dh4p, a->grads, a->ws, a->ws_bytes, B, P, a->train, a->numeric_mode, UDH_BWD_CONVS,
                           stream);
#undef STEP_TRY
}
