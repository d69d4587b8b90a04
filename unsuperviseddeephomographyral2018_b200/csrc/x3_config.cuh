// Limb formats of the two-limb mode (UDH_NUMERIC_BF16X3): forward operands (activations, forward weights, fc1's weight
// mirror) and backward operands (gradients, mirrored conv weights).  kind::f16 MMAs take fp16 or bf16 on either side.
#pragma once
#include "tc_common.cuh"

namespace udh {
#ifdef UDH_X3_FWD_FP16
constexpr int kX3Fwd = tc::kFmtF16;
#else
constexpr int kX3Fwd = tc::kFmtBF16;
#endif
constexpr int kX3Grad = tc::kFmtBF16;
}  // namespace udh
