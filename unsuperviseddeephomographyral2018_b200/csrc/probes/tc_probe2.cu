// Hardware probe for the CTA-pair (cta_group::2) tcgen05 plumbing (debug entry point, tools/tc_probe2.py):
// a 2-CTA cluster computes D[256][N] = A[256][64] . B[N][64]^T with ONE stream of M = 256 MMAs issued by the leader CTA.
//   CTA r stages rows [128r, 128r+128) of A and rows [N/2 r, N/2 r + N/2) of B in its own shared memory (same offsets in
//   both CTAs), its TMEM receives rows [128r, 128r+128) x all N columns.
//   remote_tma = 0: each CTA waits for its own TMA loads, a cluster barrier publishes them to the leader;
//   remote_tma = 1: the peer's loads complete on the LEADER's mbarrier (cp.async.bulk.tensor ... cta_group::2 with a
//                   mapa-translated barrier address + a remote expect_tx), the pattern a pipelined kernel needs.
//   The MMA sequence (4 k-steps) is repeated `reps` times; the leader reports cycles per MMA, so the same kernel measures
//   the shared-memory operand bandwidth relief of pairing (tools/tc_probe2.py compares with cta_group::1, pair = 0).
#include "../tc_common.cuh"

namespace udh {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_remote_arrive_expect_tx(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cluster.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* m, int c0, int c1, uint32_t bar_cluster_addr) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at the same offset in every CTA of `mask` once the pair's MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128)
probe2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ out,
              unsigned long long* __restrict__ cycles, int N, int pair, int remote_tma, int reps, int nacc) {
  extern __shared__ uint8_t raw[];
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = base;                       // [128][64] bf16
  uint8_t* sB = base + 16384;               // pair: [N/2][64]; single: [N][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int b_rows = pair ? N / 2 : N;
  const uint32_t my_bytes = 16384u + (uint32_t)b_rows * 128u;

  if (threadIdx.x == 0) {
    mbar_init(&bar_tma, (pair && remote_tma && rank == 0) ? 2 : 1);
    mbar_init(&bar_mma, 1);
    fence_barrier_init();
  }
  if (warp == 0) { if (pair) tmem_alloc_pair(&tmem_base_s, 512); else tmem_alloc(&tmem_base_s, 512); }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                        // barriers of both CTAs initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (threadIdx.x == 0) {
    const int a_row0 = pair ? (int)rank * 128 : 0;
    const int b_row0 = pair ? (int)rank * b_rows : 0;
    if (pair && remote_tma) {
      const uint32_t leader_bar = mapa_u32(smem_u32(&bar_tma), 0);
      if (rank == 0) mbar_arrive_expect_tx(&bar_tma, my_bytes);
      else mbar_remote_arrive_expect_tx(leader_bar, my_bytes);
      tma_load_2d_pair(sA, &tmA, 0, a_row0, leader_bar);
      tma_load_2d_pair(sB, &tmB, 0, b_row0, leader_bar);
    } else {
      mbar_arrive_expect_tx(&bar_tma, my_bytes);
      tma_load_2d(sA, &tmA, 0, a_row0, &bar_tma);
      tma_load_2d(sB, &tmB, 0, b_row0, &bar_tma);
      mbar_wait(&bar_tma, 0);
    }
  }
  __syncthreads();
  if (!(pair && remote_tma)) cluster_sync_all();    // publish "my operands are in shared memory" to the leader

  if (threadIdx.x == 0 && (!pair || rank == 0)) {
    if (pair && remote_tma) mbar_wait(&bar_tma, 0);
    tc_fence_after();
    const uint32_t a_lo = desc_lo(smem_u32(sA), 16), b_lo = desc_lo(smem_u32(sB), 16);
    const uint32_t idesc = make_idesc_bf16(pair ? 256 : 128, N, 0, 0);
    const long long t0 = clock64();
    // nacc > 1: consecutive MMAs rotate over nacc accumulators (timing only; nacc == 1 leaves the checked product in
    // accumulator 0).  Descriptors are loop-invariant registers so the issue loop is four back-to-back MMAs.
    uint64_t ad[4], bd[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { ad[k] = desc_from_lo(a_lo + k * 2); bd[k] = desc_from_lo(b_lo + k * 2); }
    uint32_t dk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) dk[k] = tmem_base + (uint32_t)((k % nacc) * N);
    if (pair) {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16_pair(dk[k], ad[k], bd[k], idesc, k >= nacc ? 1u : 0u);
#pragma unroll 1
      for (int r = 1; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_pair(dk[k], ad[k], bd[k], idesc, 1u);
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_bf16(dk[k], ad[k], bd[k], idesc, k >= nacc ? 1u : 0u);
#pragma unroll 1
      for (int r = 1; r < reps; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(dk[k], ad[k], bd[k], idesc, 1u);
      }
    }
    if (pair) umma_commit_pair(&bar_mma, 3); else umma_commit(&bar_mma);
    mbar_wait(&bar_mma, 0);
    const long long t1 = clock64();
    if (cycles) cycles[rank] = (unsigned long long)(t1 - t0);
  }
  __syncwarp();
  mbar_wait(&bar_mma, 0);                    // pair: the leader's multicast commit arrives on both CTAs' barriers
  tc_fence_after();
  const int row0 = pair ? (int)rank * 128 : (int)rank * 128;      // single mode: both CTAs compute the same product
  for (int chunk = 0; chunk < N / 32; ++chunk) {
    float v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + chunk * 32, v);
    float* dst = out + (size_t)(row0 + warp * 32 + lane) * N + chunk * 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) dst[j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                        // the peer's shared memory / TMEM stay alive until every MMA has retired
  if (warp == 0) { if (pair) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512); }
}

}  // namespace tc
}  // namespace udh

// A: [256][64] bf16 row-major (pair) / [128][64] (single), B: [N][64] bf16 row-major, out: float [256][N] (device),
// cycles: 2 x uint64 (device, nullable): clock cycles of the whole MMA stream as seen by each issuing CTA.
extern "C" int udh_debug_umma2_probe(const void* A, const void* B, float* out, unsigned long long* cycles, int N, int pair,
                                     int remote_tma, int reps, int nacc, void* stream) {
  using namespace udh;
  UDH_REQUIRE(A && B && out && (N == 64 || N == 128 || N == 256) && reps >= 1 && (nacc == 1 || nacc == 2 || nacc == 4) && nacc * N <= 512, "udh_debug_umma2_probe: bad arguments");
  CUtensorMap tmA, tmB;
  uint64_t dimsA[2] = {64, 256}, strA[2] = {2, 128};
  uint32_t boxA[2] = {64, 128};
  const int b_rows = pair ? N / 2 : N;
  uint64_t dimsB[2] = {64, (uint64_t)N}, strB[2] = {2, 128};
  uint32_t boxB[2] = {64, (uint32_t)b_rows};
  int rc = tc::make_tmap_bf16(&tmA, A, 2, dimsA, strA, boxA);
  if (rc) return rc;
  rc = tc::make_tmap_bf16(&tmB, B, 2, dimsB, strB, boxB);
  if (rc) return rc;
  const int smem = 1024 + 16384 + N * 128;
  UDH_CUDA(cudaFuncSetAttribute(tc::probe2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  tc::probe2_kernel<<<2, 128, smem, as_stream(stream)>>>(tmA, tmB, out, cycles, N, pair, remote_tma, reps, nacc);
  return check_launch("udh_debug_umma2_probe");
}
