/*
 * diffpose_b200_diag.h -- DIAGNOSTIC entry points of libdiffpose_b200.so, kept apart from the product ABI
 * (include/diffpose_b200.h): the "UMMA lab" used by tests/test_gpu_umma_lab.py and tools/mma_rate.py to pin the
 * tcgen05 operand layouts against numpy, and the hand-over trace used by tools/phase_trace.py.  Nothing here replaces a
 * reference interface.
 */
#ifndef DIFFPOSE_B200_DIAG_H
#define DIFFPOSE_B200_DIAG_H

#include "diffpose_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Diagnostic "UMMA lab": copies `smem_image` (device pointer, image_bytes % 16 == 0) to shared memory offset 0, issues the
 * listed tcgen05.mma.kind::f16 instructions in order (descriptor fields in bytes, offsets relative to the image start;
 * SWIZZLE_NONE canonical layouts; idesc = the 32-bit instruction descriptor), then writes TMEM lanes 0..127, columns
 * 0..ncols-1 to tmem_out[128][ncols] (device, fp32).  Synchronises the stream.  The GPU tests use it to pin every operand
 * flavour the tensor-core engine relies on against numpy. */
typedef struct dp_mma_op {
  unsigned a_off, a_lbo, a_sbo;   /* A operand: start, leading-dimension byte offset, stride-dimension byte offset */
  unsigned b_off, b_lbo, b_sbo;   /* B operand */
  unsigned idesc;                 /* instruction descriptor (formats, majors, N>>3 at bit 17, M>>4 at bit 24)    */
  unsigned tmem_col;              /* first accumulator column                                                     */
  unsigned accumulate;            /* bit 0: 0 D = A*B, 1 D += A*B; bit 1 (dp_selftest_umma_ts): A is in TMEM at column a_off */
} dp_mma_op;
int dp_selftest_umma(const void* smem_image, int image_bytes, const dp_mma_op* ops_host, int n_ops, float* tmem_out, int ncols,
                     void* stream);
/* Same, after preloading tensor memory: tmem_image is a device array [128 lanes][tmem_ncols] of 32-bit words written to
 * columns tmem_col0.. (tmem_ncols % 8 == 0); ops with accumulate bit 1 read their A operand from TMEM (two fp16 per word). */
int dp_selftest_umma_ts(const void* smem_image, int image_bytes, const void* tmem_image, int tmem_col0, int tmem_ncols,
                        const dp_mma_op* ops_host, int n_ops, float* tmem_out, int ncols, void* stream);

/* SM cycles of the last dp_selftest_umma[_ts] launch: out2[0] = to issue all MMAs and the commit (one thread),
 * out2[1] = from the first issue until the committed mbarrier was observed (host array of 2). */
int dp_selftest_cycles(long long* out2);

/* Diagnostic: while dev_buf is set (device pointer, `capacity` 64-bit slots; NULL/0 switches it off), thread 0 of CTA 0 of
 * the DP_ENGINE_TCG kernel stores (clock64() << 1) | kind at every hand-over between the compute warps and the MMA issuer
 * (kind 0: "operands ready" is about to be signalled, kind 1: "accumulator ready" was observed) in program order in the
 * first half of the buffer; the issuer stores clock64() before/after each of its waits in the second half.  Used by
 * tools/phase_trace.py to attribute the per-layer time to the individual epilogues and MMA groups. */
int dp_set_trace(dp_handle h, long long* dev_buf, int capacity);

#ifdef __cplusplus
}
#endif
#endif /* DIFFPOSE_B200_DIAG_H */
