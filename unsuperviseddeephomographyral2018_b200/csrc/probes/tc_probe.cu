// Hardware probe for the tcgen05 plumbing (debug entry point, used by tools/tc_probe.py and a GPU test):
// one CTA, TMA(SW128) -> smem -> tcgen05.mma -> TMEM -> global dump of all 128 lanes x 512 columns.
//   mode 0: K-major A [a_rows][64] and B [64][64]; for s in 0..7: D_s = A[s:s+128] . B^T at TMEM column 64*s.
//           The A descriptor starts s rows (s*128 bytes) into the swizzle atom; use_bo selects base_offset = s or 0.
//   mode 1: MN-major operands: G two blocks [128 px][64] (LBO = 16 KiB apart), X [a_rows px][64];
//           D_s[co][ci] = sum_px G[px][co] * X[px+s][ci], M = 128, N = 64, K = 128 px (8 MMAs of K = 16).
//   mode 2: M = 64 accumulator layout: D = A[0:64] . B^T, K-major, dumped raw.
#include "../tc_common.cuh"

namespace udh {
namespace tc {

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                    float* __restrict__ out, int mode, int use_bo, int a_bytes, int b_bytes) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;
  uint8_t* sB = base + ((a_bytes + 1023) & ~1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_tma, (uint32_t)(a_bytes + b_bytes));
    tma_load_2d(sA, &tmA, 0, 0, &bar_tma);
    tma_load_2d(sB, &tmB, 0, 0, &bar_tma);
    mbar_wait(&bar_tma, 0);
    tc_fence_after();
    const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
    if (mode == 0) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      for (int s = 0; s < 8; ++s)
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = make_smem_desc(a_addr + s * 128 + k * 32, 16, 1024, use_bo ? s : 0);
          const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024, 0);
          umma_bf16(tmem_base + s * 64, ad, bd, idesc, k > 0);
        }
    } else if (mode == 1) {
      const uint32_t idesc = make_idesc_bf16(128, 64, 1, 1);
      for (int s = 0; s < 8; ++s)
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ad = make_smem_desc(a_addr + kk * 2048, 16384, 1024, 0);
          const uint64_t bd = make_smem_desc(b_addr + s * 128 + kk * 2048, 16384, 1024, use_bo ? s : 0);
          umma_bf16(tmem_base + s * 64, ad, bd, idesc, kk > 0);
        }
    } else {
      const uint32_t idesc = make_idesc_bf16(64, 64, 0, 0);
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = make_smem_desc(a_addr + k * 32, 16, 1024, 0);
        const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024, 0);
        umma_bf16(tmem_base, ad, bd, idesc, k > 0);
      }
    }
    umma_commit(&bar_mma);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int chunk = 0; chunk < 16; ++chunk) {
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + chunk * 32, v);
    float* dst = out + (size_t)(warp * 32 + lane) * 512 + chunk * 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) dst[j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

}  // namespace tc
}  // namespace udh

// A: [a_rows][64] bf16 row-major, B: [b_rows][64] bf16 row-major (device), out: float[128*512] (device).
extern "C" int udh_debug_umma_probe(const void* A, int a_rows, const void* B, int b_rows, float* out, int mode, int use_bo,
                                    void* stream) {
  using namespace udh;
  UDH_REQUIRE(A && B && out && a_rows > 0 && a_rows <= 256 && b_rows > 0 && b_rows <= 256, "udh_debug_umma_probe: bad arguments");
  CUtensorMap tmA, tmB;
  uint64_t dimsA[2] = {64, (uint64_t)a_rows}, strA[2] = {2, 128};
  uint32_t boxA[2] = {64, (uint32_t)a_rows};
  uint64_t dimsB[2] = {64, (uint64_t)b_rows}, strB[2] = {2, 128};
  uint32_t boxB[2] = {64, (uint32_t)b_rows};
  int rc = tc::make_tmap_bf16(&tmA, A, 2, dimsA, strA, boxA);
  if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmB, B, 2, dimsB, strB, boxB);
  if (rc) return rc;
  const int a_bytes = a_rows * 128, b_bytes = b_rows * 128;
  const int smem = 1024 + ((a_bytes + 1023) & ~1023) + ((b_bytes + 1023) & ~1023);
  UDH_CUDA(cudaFuncSetAttribute(tc::probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc::probe_kernel<<<1, 128, smem, as_stream(stream)>>>(tmA, tmB, out, mode, use_bo, a_bytes, b_bytes);
  return check_launch("udh_debug_umma_probe");
}
