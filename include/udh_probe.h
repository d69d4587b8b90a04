/*
 * libudh_probe.so - hardware probes used while developing the tcgen05 / TMA kernels of libudh (NOT part of the product
 * library).  Sources: unsuperviseddeephomographyral2018_b200/csrc/probes/; drivers: tools/tc_probe.py, tools/tc_probe2.py;
 * recorded outputs: profiles/r1_tc_probe.log, profiles/r1_tc_probe2.log.  Same conventions as include/udh.h.
 */
#ifndef UDH_PROBE_H_
#define UDH_PROBE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* TMA(SWIZZLE_128B) -> smem -> tcgen05.mma -> TMEM with descriptor starts shifted by whole 128-byte rows; K-major and
 * MN-major operands; M = 64 accumulator layout (csrc/probes/tc_probe.cu). */
int udh_debug_umma_probe(const void* A, int a_rows, const void* B, int b_rows, float* out, int mode, int use_bo, void* stream);
/* CTA-pair (cta_group::2) probe, csrc/probes/tc_probe2.cu: D[256][N] = A[256][64] . B[N][64]^T on a 2-CTA cluster; cycles[2]. */
int udh_debug_umma2_probe(const void* A, const void* B, float* out, unsigned long long* cycles, int N, int pair,
                          int remote_tma, int reps, int nacc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UDH_PROBE_H_ */
