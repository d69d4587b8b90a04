// tcgen05 GEMM for the fully connected head (fc1: 32768 -> 1024), bf16 operands, fp32 TMEM accumulation.
//   D[M][N] (+)= A . B over K, tiles of 128 x 256, k-blocks of 64, 4-stage TMA ring, two TMEM accumulator sets.
// Each operand is either K-major (contraction index contiguous in memory: tile = [rows][64 k], box (64, rows)) or
// MN-major (row/column index contiguous: tile = 64-wide column blocks of [64 k][64 mn]), which covers
//   forward   y  = x . W        A = x [B][F] K-major,      B = W [F][N] MN-major,   split-K, atomic accumulation
//   dgrad     dx = dy . W^T     A = dy [B][N] K-major,     B = W [F][N] as [n=f][k] K-major
//   wgrad     dW = x^T . dy     A = x [B][F] MN-major,     B = dy [B][N] MN-major   (K = batch)
// with no transposed copies.  warp 0: TMA | warp 1: MMA | warp 2: TMEM alloc | warps 4-7: epilogue.
#pragma once
#include "tc_common.cuh"

namespace udh {
namespace tc {

struct GemmGeom {
  int M, N;            // logical output size (rows beyond M / cols beyond N are not stored)
  int m_tiles, n_tiles, k_splits;
  int kb_per_split;    // k-blocks (of 64) per split
  int64_t ldc;
};

constexpr int kGemmStages = 4;
constexpr int kGemmStageBytes = 16384 + 32768;   // A 128 x 64, B 256 x 64 (bf16)

constexpr int kGemmEpiBytes = 2 * 4 * 32 * 128;  // TMA-store staging of the non-atomic epilogue: 2 x (32 rows x 32 floats) per warp

// ATOMIC: split-K partial tiles are added with 16-byte vector reductions.  Otherwise the tile is stored through a swizzled
// shared-memory staging block and TMA (full 128-byte lines; rows / columns beyond M / N are clipped by the tensor map).
template <bool A_MN, bool B_MN, bool ATOMIC>
__global__ void __launch_bounds__(256, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
               const GemmGeom g, float* __restrict__ C) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sEpi = base + (size_t)kGemmStages * kGemmStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEpi + (ATOMIC ? 0 : kGemmEpiBytes));
  uint64_t* full = bars;                       // [stages]
  uint64_t* empty = bars + kGemmStages;        // [stages]
  uint64_t* t_full = bars + 2 * kGemmStages;   // [2]
  uint64_t* t_empty = t_full + 2;              // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kGemmStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 128); }
    fence_barrier_init();
  }
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int num_tiles = g.m_tiles * g.n_tiles * g.k_splits;
  pdl_wait();                  // predecessor grid complete: global memory may be touched from here on
  pdl_trigger();
  const int my_tiles = (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  auto decode = [&](int tile, int& mt, int& nt, int& ks) {
    ks = tile % g.k_splits; tile /= g.k_splits;
    nt = tile % g.n_tiles;
    mt = tile / g.n_tiles;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int i = 0; i < my_tiles; ++i) {
        int mt, nt, ks;
        decode((int)blockIdx.x + i * (int)gridDim.x, mt, nt, ks);
        for (int kb = 0; kb < g.kb_per_split; ++kb, ++cnt) {
          const int s = cnt % kGemmStages;
          const int k0 = (ks * g.kb_per_split + kb) * 64;
          mbar_wait(&empty[s], ((cnt / kGemmStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full[s], kGemmStageBytes);
          uint8_t* sa = base + (size_t)s * kGemmStageBytes;
          uint8_t* sb = sa + 16384;
          if (A_MN) {
            for (int j = 0; j < 2; ++j) tma_load_2d(sa + j * 8192, &tmA, mt * 128 + j * 64, k0, &full[s]);
          } else {
            tma_load_2d(sa, &tmA, k0, mt * 128, &full[s]);
          }
          if (B_MN) {
            for (int j = 0; j < 4; ++j) tma_load_2d(sb + j * 8192, &tmB, nt * 256 + j * 64, k0, &full[s]);
          } else {
            tma_load_2d(sb, &tmB, k0, nt * 256, &full[s]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // whole warp converged; one elected lane issues
    constexpr uint32_t idesc = make_idesc_bf16(128, 256, A_MN ? 1 : 0, B_MN ? 1 : 0);
    uint32_t cnt = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int b = i & 1;
      mbar_wait(&t_empty[b], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      for (int kb = 0; kb < g.kb_per_split; ++kb, ++cnt) {
        const int s = cnt % kGemmStages;
        mbar_wait(&full[s], (cnt / kGemmStages) & 1);
        tc_fence_after();
        const uint32_t a_addr = smem_u32(base + (size_t)s * kGemmStageBytes), b_addr = a_addr + 16384;
        const uint32_t a_lo = desc_lo(a_addr, A_MN ? 8192 : 16), b_lo = desc_lo(b_addr, B_MN ? 8192 : 16);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16(tmem_base + (uint32_t)(b * 256), desc_from_lo(a_lo + k * (A_MN ? 128 : 2)), desc_from_lo(b_lo + k * (B_MN ? 128 : 2)),
                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&empty[s]);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(&t_full[b]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    uint8_t* stg0 = sEpi + ew * 8192;            // two staging blocks per warp: chunk c fills one while the TMA store
    const int sw = lane & 7;                      // of chunk c-1 still reads the other
    int stores_in_flight = 0;
    for (int i = 0; i < my_tiles; ++i) {
      const int b = i & 1;
      int mt, nt, ks;
      decode((int)blockIdx.x + i * (int)gridDim.x, mt, nt, ks);
      mbar_wait(&t_full[b], (i >> 1) & 1);
      tc_fence_after();
      const int m = mt * 128 + ew * 32 + lane;
      float* crow = C + (int64_t)m * g.ldc + nt * 256;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(b * 256 + c * 32), v);
        const int n0 = nt * 256 + c * 32;
        if (ATOMIC) {
          if (m < g.M) {
            if (n0 + 32 <= g.N) {
#pragma unroll
              for (int j = 0; j < 8; ++j) red_add_v4(crow + c * 32 + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (n0 + j < g.N) atomicAdd(crow + c * 32 + j, v[j]);
            }
          }
        } else {
          uint8_t* stg = stg0 + (c & 1) * 4096;
          const uint32_t my_row = smem_u32(stg) + lane * 128;
          if (stores_in_flight >= 2) { if (lane == 0) bulk_wait_read1(); __syncwarp(); stores_in_flight = 1; }
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(my_row + (uint32_t)((j4 ^ sw) << 4)), "f"(v[4 * j4]), "f"(v[4 * j4 + 1]),
                         "f"(v[4 * j4 + 2]), "f"(v[4 * j4 + 3]) : "memory");
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {                          // (always commit, possibly an empty group, so the group count stays uniform)
            if (n0 < g.N && mt * 128 + ew * 32 < g.M) tma_store_2d(&tmC, stg, n0, mt * 128 + ew * 32);
            bulk_commit();
          }
          ++stores_in_flight;
        }
      }
      tc_fence_before();
      mbar_arrive(&t_empty[b]);
    }
    if (!ATOMIC) { if (lane == 0) bulk_wait0(); __syncwarp(); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace udh
