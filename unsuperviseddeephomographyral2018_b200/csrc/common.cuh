// Shared helpers of libudh (error reporting, launch checks, warp/block reductions).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/udh.h"

namespace udh {

void set_error(const char* fmt, ...);

// ---- in-library instrumentation (bench.py reads it): kernel-launch counter and per-phase CUDA-event timers ----
enum ProfTag {
  PROF_CONV_FWD0 = 0,    // +layer (0..7)
  PROF_CONV_DGRAD0 = 8,  // +layer
  PROF_CONV_WGRAD0 = 16, // +layer
  PROF_POOL_FWD = 24, PROF_POOL_BWD, PROF_FC_FWD, PROF_FC_BWD, PROF_DLT, PROF_WARP_FWD, PROF_WARP_BWD, PROF_SSIM,
  PROF_H4P_LOSS, PROF_ADAM, PROF_ELTWISE, PROF_TC_PREP, PROF_NUM_TAGS
};
extern unsigned long long g_launches;
extern int g_sm_reserve;      // SMs the persistent tensor-core kernels leave free (udh_set_sm_reserve)
// CTAs a persistent one-CTA-per-SM kernel should launch on the current device
inline int persistent_ctas() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n = sms - g_sm_reserve;
  return n > 0 ? n : 1;
}
void prof_begin(int tag, cudaStream_t st);
void prof_end(int tag, cudaStream_t st);
struct ProfScope {
  int tag; cudaStream_t st;
  ProfScope(int t, cudaStream_t s) : tag(t), st(s) { prof_begin(tag, st); }
  ~ProfScope() { prof_end(tag, st); }
};

inline int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return UDH_ECUDA;
  }
  return UDH_OK;
}

#define UDH_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      udh::set_error(__VA_ARGS__);    \
      return UDH_EINVAL;              \
    }                                 \
  } while (0)

#define UDH_CUDA(call)                                                      \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess) {                                               \
      udh::set_error("%s failed: %s", #call, cudaGetErrorString(e__));      \
      return UDH_ECUDA;                                                     \
    }                                                                       \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of NV values per thread; result valid in thread 0.  `scratch` holds NV * 32 elements of T.
template <typename T, int NV>
__device__ __forceinline__ void block_sum(T (&v)[NV], T* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) scratch[i * 32 + warp] = v[i];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      T x = lane < nwarp ? scratch[i * 32 + lane] : T(0);
      v[i] = warp_sum(x);
    }
  }
}

}  // namespace udh
