// conv1_1 (2 -> 64 channels, K = 18) for the bf16 numeric mode: too thin for the tensor pipe (K = 18 < one MMA k-step
// of useful work), so it runs on CUDA cores in fp32 with a layout chosen for the HBM side: one warp walks one image row,
// each lane owns two output channels, so every load/store of the 64-channel bf16 stream is one coalesced 128-byte line.
//   forward : I1, I2 fp32 planes -> relu(conv + bias) written into the zero-bordered bf16 stream [B][H+2][W+2][64]
//   wgrad   : dW[tap][ci][co] += sum I_ci(shifted) * G,  db[co] += sum G,  G read from its bf16 stream
// Reference: code/homography_model.py:88-95,108-109 (conv_block1/conv1) and its TF autodiff.
#include <cuda_bf16.h>

#include "cnn_kernels.cuh"

namespace udh {
namespace {

constexpr int ROWS = 8;          // image rows per CTA tile (one per warp)

// stage the (ROWS+2) x (W+2) halo of both planes, zero padded: tile[ci][r][c]
__device__ __forceinline__ void load_planes(float* tile, const float* __restrict__ I1, const float* __restrict__ I2, int n, int y0,
                                            int H, int W) {
  const int pitch = W + 2, plane = (ROWS + 2) * pitch;
  for (int i = threadIdx.x; i < 2 * plane; i += blockDim.x) {
    const int ci = i / plane, rem = i - ci * plane;
    const int r = rem / pitch, c = rem - r * pitch;
    const int gy = y0 - 1 + r, gx = c - 1;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = __ldg((ci ? I2 : I1) + ((size_t)n * H + gy) * W + gx);
    tile[i] = v;
  }
}

// spread the low 16 bits of x to the even bit positions
__device__ __forceinline__ uint32_t spread16(uint32_t x) {
  x &= 0xFFFFu;
  x = (x | (x << 8)) & 0x00FF00FFu;
  x = (x | (x << 4)) & 0x0F0F0F0Fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}

__global__ void __launch_bounds__(256) conv1_fwd_kernel(const float* __restrict__ I1, const float* __restrict__ I2,
                                                        const float* __restrict__ w, const float* __restrict__ bias,
                                                        __nv_bfloat16* __restrict__ out, uint32_t* __restrict__ mask_out, int B, int H,
                                                        int W) {
  extern __shared__ float tile[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pitch = W + 2, plane = (ROWS + 2) * pitch;
  float wr[9][2][2];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(w + (t * 2 + ci) * 64 + 2 * lane));
      wr[t][ci][0] = v.x; wr[t][ci][1] = v.y;
    }
  const float2 bv = __ldg(reinterpret_cast<const float2*>(bias + 2 * lane));
  const int tiles_per_img = H / ROWS;
  for (int tix = blockIdx.x; tix < B * tiles_per_img; tix += gridDim.x) {
    const int n = tix / tiles_per_img, y0 = (tix - n * tiles_per_img) * ROWS;
    __syncthreads();
    load_planes(tile, I1, I2, n, y0, H, W);
    __syncthreads();
    const int y = y0 + warp;
    float xw[2][3][3];
#pragma unroll
    for (int ci = 0; ci < 2; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        xw[ci][ky][1] = tile[ci * plane + (warp + ky) * pitch + 0];
        xw[ci][ky][2] = tile[ci * plane + (warp + ky) * pitch + 1];
      }
    __nv_bfloat16* orow = out + (((size_t)n * (H + 2) + y + 1) * (W + 2) + 1) * 64 + 2 * lane;
    uint32_t keep_e = 0, keep_o = 0;
#pragma unroll 4
    for (int x = 0; x < W; ++x) {
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          xw[ci][ky][0] = xw[ci][ky][1]; xw[ci][ky][1] = xw[ci][ky][2];
          xw[ci][ky][2] = tile[ci * plane + (warp + ky) * pitch + x + 2];
        }
      float a0 = bv.x, a1 = bv.y;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int ci = 0; ci < 2; ++ci) {
            a0 = fmaf(xw[ci][ky][kx], wr[ky * 3 + kx][ci][0], a0);
            a1 = fmaf(xw[ci][ky][kx], wr[ky * 3 + kx][ci][1], a1);
          }
      const __nv_bfloat162 p = __floats2bfloat162_rn(fmaxf(a0, 0.f), fmaxf(a1, 0.f));
      *reinterpret_cast<__nv_bfloat162*>(orow + (size_t)x * 64) = p;
      if (mask_out) {
        // 1-bit ReLU mask, channel c -> bit c%32 of word c/32 (lane l owns channels 2l, 2l+1).  Lane x%32 keeps the two
        // ballots of position x; every 32 positions each lane packs and stores one position: a coalesced 256-byte store.
        const uint32_t e = __ballot_sync(0xffffffffu, __low2float(p) > 0.f), o = __ballot_sync(0xffffffffu, __high2float(p) > 0.f);
        if ((x & 31) == lane) { keep_e = e; keep_o = o; }
        if ((x & 31) == 31) {
          const size_t q = ((size_t)n * (H + 2) + y + 1) * (W + 2) + (x - 31 + lane) + 1;
          uint2 mw;
          mw.x = spread16(keep_e) | (spread16(keep_o) << 1);
          mw.y = spread16(keep_e >> 16) | (spread16(keep_o >> 16) << 1);
          *reinterpret_cast<uint2*>(mask_out + q * 2) = mw;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) conv1_wgrad_kernel(const float* __restrict__ I1, const float* __restrict__ I2,
                                                          const __nv_bfloat16* __restrict__ G, float* __restrict__ dW,
                                                          float* __restrict__ db, int B, int H, int W) {
  extern __shared__ float tile[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pitch = W + 2, plane = (ROWS + 2) * pitch;
  float acc[9][2][2], bacc[2] = {0.f, 0.f};
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) { acc[t][ci][0] = 0.f; acc[t][ci][1] = 0.f; }
  const int tiles_per_img = H / ROWS;
  for (int tix = blockIdx.x; tix < B * tiles_per_img; tix += gridDim.x) {
    const int n = tix / tiles_per_img, y0 = (tix - n * tiles_per_img) * ROWS;
    __syncthreads();
    load_planes(tile, I1, I2, n, y0, H, W);
    __syncthreads();
    const int y = y0 + warp;
    float xw[2][3][3];
#pragma unroll
    for (int ci = 0; ci < 2; ++ci)
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        xw[ci][ky][1] = tile[ci * plane + (warp + ky) * pitch + 0];
        xw[ci][ky][2] = tile[ci * plane + (warp + ky) * pitch + 1];
      }
    const __nv_bfloat16* grow = G + (((size_t)n * (H + 2) + y + 1) * (W + 2) + 1) * 64 + 2 * lane;
#pragma unroll 4
    for (int x = 0; x < W; ++x) {
#pragma unroll
      for (int ci = 0; ci < 2; ++ci)
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
          xw[ci][ky][0] = xw[ci][ky][1]; xw[ci][ky][1] = xw[ci][ky][2];
          xw[ci][ky][2] = tile[ci * plane + (warp + ky) * pitch + x + 2];
        }
      const __nv_bfloat162 gp = *reinterpret_cast<const __nv_bfloat162*>(grow + (size_t)x * 64);
      const float g0 = __low2float(gp), g1 = __high2float(gp);
      bacc[0] += g0; bacc[1] += g1;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx)
#pragma unroll
          for (int ci = 0; ci < 2; ++ci) {
            acc[ky * 3 + kx][ci][0] = fmaf(xw[ci][ky][kx], g0, acc[ky * 3 + kx][ci][0]);
            acc[ky * 3 + kx][ci][1] = fmaf(xw[ci][ky][kx], g1, acc[ky * 3 + kx][ci][1]);
          }
    }
  }
  // cross-warp reduction in smem (reuse the tile), then one atomic per output per CTA
  __syncthreads();
  float* red = tile;                                   // [8 warps][38][64]: 36 weight sums + 2 bias sums per lane pair
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
      red[(warp * 19 + t * 2 + ci) * 64 + 2 * lane] = acc[t][ci][0];
      red[(warp * 19 + t * 2 + ci) * 64 + 2 * lane + 1] = acc[t][ci][1];
    }
  red[(warp * 19 + 18) * 64 + 2 * lane] = bacc[0];
  red[(warp * 19 + 18) * 64 + 2 * lane + 1] = bacc[1];
  __syncthreads();
  for (int i = threadIdx.x; i < 19 * 64; i += blockDim.x) {
    float s = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) s += red[wv * 19 * 64 + i];
    if (i < 18 * 64) atomicAdd(dW + i, s); else atomicAdd(db + (i - 18 * 64), s);
  }
}

}  // namespace

int conv1_fwd_bf16(const float* I1, const float* I2, const float* w, const float* bias, __nv_bfloat16* out_pad, uint32_t* mask_out,
                   int B, int H, int W, cudaStream_t st) {
  UDH_REQUIRE(H % ROWS == 0 && W % 32 == 0, "conv1_fwd_bf16: unsupported shape");
  const size_t smem = (size_t)2 * (ROWS + 2) * (W + 2) * sizeof(float);
  const int tiles = B * (H / ROWS);
  conv1_fwd_kernel<<<tiles < 148 * 8 ? tiles : 148 * 8, 256, smem, st>>>(I1, I2, w, bias, out_pad, mask_out, B, H, W);
  return check_launch("conv1_fwd_bf16");
}

int conv1_wgrad_bf16(const float* I1, const float* I2, const __nv_bfloat16* G_pad, float* dW, float* db, int B, int H, int W,
                     cudaStream_t st) {
  UDH_REQUIRE(H % ROWS == 0 && W >= 8, "conv1_wgrad_bf16: unsupported shape");
  size_t smem = (size_t)2 * (ROWS + 2) * (W + 2) * sizeof(float);
  const size_t red = (size_t)8 * 19 * 64 * sizeof(float);
  if (red > smem) smem = red;
  UDH_CUDA(cudaFuncSetAttribute(conv1_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int tiles = B * (H / ROWS);
  conv1_wgrad_kernel<<<tiles < 148 * 4 ? tiles : 148 * 4, 256, smem, st>>>(I1, I2, G_pad, dW, db, B, H, W);
  return check_launch("conv1_wgrad_bf16");
}

}  // namespace udh
