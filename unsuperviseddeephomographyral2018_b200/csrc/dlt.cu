// Batched 4-point DLT (Row D): one warp per sample, the 8x9 augmented system [A|b] lives in registers
// (lane r < 8 owns row r), LU with partial pivoting and back substitution run on warp shuffles.
// Reference: code/homography_model.py:169-250 (assembly with Aux_M*, utils/utils.py:11-122; tf.matrix_solve).
// Latency-bound (~100 B, ~1 kFLOP per sample): no tensor cores, no shared memory traffic.
// I/O is fp32 like the reference; the elimination itself runs in fp64 registers: cond(A) ~ 5e5 in pixel units, so
// an fp32 LU carries ~1e-4 relative noise that depends on the LU variant (LAPACK vs cuSOLVER vs this one) — there
// is no canonical fp32 answer to match, and fp64 costs nothing in a latency-bound kernel.
#include "common.cuh"

namespace udh {

// Solve the 8x8 system whose row `lane` (< 8) is a[0..7] | a[8].  All 32 lanes call; on return every lane holds
// the full solution x[0..7].  Partial pivoting picks the first row of maximal |a[k]| (LAPACK isamax rule).
__device__ __forceinline__ void lu_solve_warp(double (&a)[9], double (&x)[8], int lane) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    double best = (lane >= k && lane < 8) ? fabs(a[k]) : -1.0;
    int idx = lane;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      double ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
    }
    const int piv = __shfl_sync(0xffffffffu, idx, 0);     // lanes 0..7 agree after the xor-butterfly
    const int src = (lane == k) ? piv : ((lane == piv) ? k : lane);
#pragma unroll
    for (int c = k; c < 9; ++c) a[c] = __shfl_sync(0xffffffffu, a[c], src);   // row swap k <-> piv
    double pk[9];
#pragma unroll
    for (int c = k; c < 9; ++c) pk[c] = __shfl_sync(0xffffffffu, a[c], k);
    if (lane > k && lane < 8) {
      const double f = a[k] / pk[k];
#pragma unroll
      for (int c = k + 1; c < 9; ++c) a[c] = fma(-f, pk[c], a[c]);
      a[k] = 0.0;
    }
  }
#pragma unroll
  for (int k = 7; k >= 0; --k) {
    const double xk = __shfl_sync(0xffffffffu, a[8] / a[k], k);
    x[k] = xk;
    if (lane < k) a[8] = fma(-a[k], xk, a[8]);
  }
}

// Row `r` of [A|b] (homography_model.py:223-238): corner i = r/2, (x,y) = pts1, (u,v) = pts1 + h4p
//   even r: [0,0,0,-x,-y,-1, v*x, v*y | -v]      odd r: [x,y,1,0,0,0,-u*x,-u*y | u]
__device__ __forceinline__ void build_row(const float* __restrict__ p1, const float* __restrict__ h, int r, double (&a)[9]) {
  const int i = r >> 1;
  const double x = p1[2 * i], y = p1[2 * i + 1];
  const double u = (double)(p1[2 * i] + h[2 * i]), v = (double)(p1[2 * i + 1] + h[2 * i + 1]);   // pts2 is an fp32 tensor (:176)
  if ((r & 1) == 0) {
    a[0] = 0.; a[1] = 0.; a[2] = 0.; a[3] = -x; a[4] = -y; a[5] = -1.; a[6] = v * x; a[7] = v * y; a[8] = -v;
  } else {
    a[0] = x; a[1] = y; a[2] = 1.; a[3] = 0.; a[4] = 0.; a[5] = 0.; a[6] = -u * x; a[7] = -u * y; a[8] = u;
  }
}

__global__ void __launch_bounds__(128) dlt_fwd_kernel(const float* __restrict__ pts1, const float* __restrict__ h4p,
                                                      float* __restrict__ H, int B) {
  pdl_wait(); pdl_trigger();   // launched through launch_chain (common.cuh)
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;                                        // warp-uniform
  double a[9], x[8];
  if (lane < 8) build_row(pts1 + (size_t)b * 8, h4p + (size_t)b * 8, lane, a);
  else {
#pragma unroll
    for (int c = 0; c < 9; ++c) a[c] = 0.;
  }
  lu_solve_warp(a, x, lane);
  float out = 1.0f;                                          // h33 = 1 (homography_model.py:247-250)
#pragma unroll
  for (int k = 0; k < 8; ++k) if (lane == k) out = (float)x[k];
  if (lane < 9) H[(size_t)b * 9 + lane] = out;
}

// Backward (TF autodiff of matrix_solve + the selector products): lambda = A^-T g with g = dH[0:8];
// dA = -lambda h^T, db = lambda; only A columns 6,7 and b depend on pts2 = pts1 + h4p:
//   dL/du_i =  lambda[2i+1] * (h6 x_i + h7 y_i + 1),   dL/dv_i = -lambda[2i] * (h6 x_i + h7 y_i + 1).
__global__ void __launch_bounds__(128) dlt_bwd_kernel(const float* __restrict__ pts1, const float* __restrict__ h4p,
                                                      const float* __restrict__ H, const float* __restrict__ dH,
                                                      float* __restrict__ dh4p, int B) {
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const float* p1 = pts1 + (size_t)b * 8;
  const float* h = h4p + (size_t)b * 8;
  double a[9], lam[8];
#pragma unroll
  for (int c = 0; c < 9; ++c) a[c] = 0.;
  if (lane < 8) {
    // row `lane` of A^T = column `lane` of A
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double x = p1[2 * i], y = p1[2 * i + 1];
      const double u = (double)(p1[2 * i] + h[2 * i]), v = (double)(p1[2 * i + 1] + h[2 * i + 1]);
      double e0, e1;                                         // entries at rows 2i, 2i+1 of column `lane`
      switch (lane) {
        case 0: e0 = 0.; e1 = x; break;
        case 1: e0 = 0.; e1 = y; break;
        case 2: e0 = 0.; e1 = 1.; break;
        case 3: e0 = -x; e1 = 0.; break;
        case 4: e0 = -y; e1 = 0.; break;
        case 5: e0 = -1.; e1 = 0.; break;
        case 6: e0 = v * x; e1 = -u * x; break;
        default: e0 = v * y; e1 = -u * y; break;
      }
      a[2 * i] = e0; a[2 * i + 1] = e1;
    }
    a[8] = dH[(size_t)b * 9 + lane];
  }
  lu_solve_warp(a, lam, lane);
  const double h6 = H[(size_t)b * 9 + 6], h7 = H[(size_t)b * 9 + 7];
  float out = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double s = fma(h6, (double)p1[2 * i], fma(h7, (double)p1[2 * i + 1], 1.0));
    if (lane == 2 * i) out = (float)(lam[2 * i + 1] * s);    // d/du_i
    if (lane == 2 * i + 1) out = (float)(-lam[2 * i] * s);   // d/dv_i
  }
  if (lane < 8) dh4p[(size_t)b * 8 + lane] = out;
}

}  // namespace udh

extern "C" int udh_dlt_fwd(const float* pts1, const float* h4p, float* H, int B, void* stream) {
  UDH_REQUIRE(pts1 && h4p && H && B >= 0, "udh_dlt_fwd: null pointer or negative batch");
  if (B == 0) return UDH_OK;
  udh::ProfScope ps(udh::PROF_DLT, udh::as_stream(stream));
  udh::launch_chain(udh::dlt_fwd_kernel, dim3((B + 3) / 4), dim3(128), 0, udh::as_stream(stream), pts1, h4p, H, B);
  return udh::check_launch("udh_dlt_fwd");
}

extern "C" int udh_dlt_bwd(const float* pts1, const float* h4p, const float* H, const float* dH, float* dh4p, int B,
                           void* stream) {
  UDH_REQUIRE(pts1 && h4p && H && dH && dh4p && B >= 0, "udh_dlt_bwd: null pointer or negative batch");
  if (B == 0) return UDH_OK;
  udh::ProfScope ps(udh::PROF_DLT, udh::as_stream(stream));
  udh::dlt_bwd_kernel<<<(B + 3) / 4, 128, 0, udh::as_stream(stream)>>>(pts1, h4p, H, dH, dh4p, B);
  return udh::check_launch("udh_dlt_bwd");
}
