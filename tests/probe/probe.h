/* Test-only C entry point of tests/probe/libdepth_b200_probe.so (tests/test_umma_probe_gpu.py). */
#ifndef DEPTH_B200_PROBE_H
#define DEPTH_B200_PROBE_H
#include <cuda_runtime.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* One CTA, one tcgen05.mma chain with host-specified shared-memory descriptors; dumps TMEM lanes 0..127 x
 * ncols_dump fp32 columns to `out`.  A / B are row-major bf16 matrices loaded by TMA in boxes of
 * (box_rows x box_cols), boxes laid out consecutively along the column axis.  adesc / bdesc are HOST arrays
 * {lbo_bytes, sbo_bytes, layout_type, k_advance_bytes, start_offset_bytes[, base_offset (adesc only)]}.  */
int dp_umma_probe(const void* A, int a_rows, int a_cols, int a_box_rows, int a_box_cols, const void* B, int b_rows,
                  int b_cols, int b_box_rows, int b_box_cols, int M, int N, int nk, int a_mn_major, int b_mn_major,
                  const uint32_t* adesc, const uint32_t* bdesc, float* out, int ncols_dump, cudaStream_t stream);
#ifdef __cplusplus
}
#endif
#endif
