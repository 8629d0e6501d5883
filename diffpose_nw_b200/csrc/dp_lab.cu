// UMMA lab (diagnostic): runs a caller-described list of tcgen05.mma instructions over a caller-built shared-memory / TMEM
// image and dumps tensor memory, so that the GPU tests can pin every operand flavour the tensor-core engines rely on
// (K-major / MN-major A and B, custom leading-dimension offsets, TMEM-resident A, N = 16..128) against numpy
// (tests/test_gpu_umma_lab.py), and tools/mma_rate.py can measure issue rates.  Declared in include/diffpose_b200_diag.h.
#include <cuda_fp16.h>
#include "dp_internal.h"
#include "dp_sm100.cuh"

namespace dp {

namespace {

using namespace sm100;

__device__ long long g_lab_cycles[2];   // last lab launch: cycles to issue all MMAs, cycles until the commit was observed

// ---- UMMA lab: run a caller-described list of tcgen05.mma instructions over a caller-built shared-memory image and
// dump TMEM.  The GPU tests use it to pin down every operand flavour the engine relies on (K-major / MN-major A and B,
// custom leading-dimension offsets, N = 32/64/96/128) against numpy.
__global__ void __launch_bounds__(128, 1) tc_lab_kernel(const uint8_t* __restrict__ image, int image_bytes, const dp_mma_op* __restrict__ ops,
                                                        int n_ops, float* __restrict__ out, int ncols, const uint32_t* __restrict__ tmem_image,
                                                        int tmem_col0, int tmem_ncols) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sbase = smem_u32(smem);
  const int off_bar = (image_bytes + 15) / 16 * 16;
  const uint32_t done = sbase + off_bar;
  if (tid == 0) { mbar_init(done, 1); fence_mbar_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc(sbase + off_bar + 16, 512);
  for (int i = tid; i < image_bytes / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(image)[i];
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<uint32_t*>(smem + off_bar + 16);
  // clear the dumped columns so that untouched accumulators read as zero
  for (int c = 0; c < ncols; c += 8) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(tmem_base + ((uint32_t)(warp * 32) << 16) + c), "r"(0) : "memory");
  }
  // optional TMEM preload (A operands that live in tensor memory): lane = row, tmem_ncols 32-bit columns from tmem_col0
  for (int c = 0; c < tmem_ncols; c += 8) {
    const uint32_t* src = tmem_image + (size_t)(warp * 32 + (tid & 31)) * tmem_ncols + c;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(tmem_base + ((uint32_t)(warp * 32) << 16) + tmem_col0 + c),
                 "r"(src[0]), "r"(src[1]), "r"(src[2]), "r"(src[3]), "r"(src[4]), "r"(src[5]), "r"(src[6]), "r"(src[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  long long t_issue0 = 0, t_issue1 = 0;
  if (tid == 0) {
    tc_fence_after();
    // a single op with tmem_col >= 0x10000 is a rate measurement: it is issued (tmem_col >> 16) times from registers
    int reps = 1;
    dp_mma_op first = ops[0];
    if (n_ops == 1 && first.tmem_col >= 0x10000u) { reps = (int)(first.tmem_col >> 16); first.tmem_col &= 0xFFFFu; }
    if (reps > 1) {
      // rate measurement: descriptors live in registers, the loop body is the MMA alone
      const uint64_t ad = make_desc(sbase + first.a_off, first.a_lbo, first.a_sbo), bd = make_desc(sbase + first.b_off, first.b_lbo, first.b_sbo);
      const uint32_t d = tmem_base + first.tmem_col, ta = tmem_base + first.a_off, idc = first.idesc;
      const bool ts = (first.accumulate & 2u) != 0;
      t_issue0 = clock64();
      if (ts) {
#pragma unroll 8
        for (int i = 0; i < reps; ++i) umma_f16_ts(d, ta, bd, idc, 1u);
      } else {
#pragma unroll 8
        for (int i = 0; i < reps; ++i) umma_f16(d, ad, bd, idc, 1u);
      }
    } else {
      t_issue0 = clock64();
      for (int i = 0; i < n_ops; ++i) {
        const dp_mma_op o = (n_ops == 1) ? first : ops[i];
        if (o.accumulate & 2u)
          umma_f16_ts(tmem_base + o.tmem_col, tmem_base + o.a_off, make_desc(sbase + o.b_off, o.b_lbo, o.b_sbo), o.idesc, o.accumulate & 1u);
        else
          umma_f16(tmem_base + o.tmem_col, make_desc(sbase + o.a_off, o.a_lbo, o.a_sbo), make_desc(sbase + o.b_off, o.b_lbo, o.b_sbo), o.idesc, o.accumulate);
      }
    }
    umma_commit(done);
    t_issue1 = clock64();
  }
  __syncwarp();
  mbar_wait(done, 0);
  tc_fence_after();
  if (tid == 0) { g_lab_cycles[0] = t_issue1 - t_issue0; g_lab_cycles[1] = clock64() - t_issue0; }
  const int row = warp * 32 + (tid & 31);
  for (int c = 0; c < ncols; c += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
    for (int i = 0; i < 16; ++i) out[row * ncols + c + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

}  // namespace

static long long g_lab_cycles_host[2] = {0, 0};
void tc_lab_cycles(long long* out2) { out2[0] = g_lab_cycles_host[0]; out2[1] = g_lab_cycles_host[1]; }

// Diagnostic entry point behind dp_selftest_umma (see include/diffpose_b200.h).
int tc_lab(const void* image_dev, int image_bytes, const dp_mma_op* ops_host, int n_ops, float* out_dev, int ncols,
           const void* tmem_image_dev, int tmem_col0, int tmem_ncols, cudaStream_t s) {
  if (tmem_image_dev == nullptr) { tmem_col0 = 0; tmem_ncols = 0; }
  if (tmem_ncols < 0 || tmem_ncols % 8 || tmem_col0 < 0 || tmem_col0 + tmem_ncols > 512) {
    set_error("dp_selftest_umma_ts: the TMEM image must be a multiple of 8 columns inside the 512 allocated ones");
    return DP_ERR_INVALID;
  }
  if (image_bytes <= 0 || image_bytes % 16 || image_bytes > 200 * 1024 || n_ops <= 0 || n_ops > 256 || ncols <= 0 || ncols > 512 || ncols % 16) {
    set_error("dp_selftest_umma: image must be a multiple of 16 B (<= 200 KiB), 1..256 ops, ncols a multiple of 16 (<= 512)");
    return DP_ERR_INVALID;
  }
  dp_mma_op* ops_dev = nullptr;
  DP_CUDA(cudaMalloc(reinterpret_cast<void**>(&ops_dev), n_ops * sizeof(dp_mma_op)));
  cudaError_t e = cudaMemcpyAsync(ops_dev, ops_host, n_ops * sizeof(dp_mma_op), cudaMemcpyHostToDevice, s);
  const int smem = (image_bytes + 15) / 16 * 16 + 64;
  if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_lab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (e == cudaSuccess) {
    tc_lab_kernel<<<1, 128, smem, s>>>(static_cast<const uint8_t*>(image_dev), image_bytes, ops_dev, n_ops, out_dev, ncols,
                                       static_cast<const uint32_t*>(tmem_image_dev), tmem_col0, tmem_ncols);
    count_launch();
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  if (e == cudaSuccess) e = cudaMemcpyFromSymbol(g_lab_cycles_host, g_lab_cycles, sizeof(g_lab_cycles_host));
  cudaFree(ops_dev);
  if (e != cudaSuccess) { set_error(std::string("dp_selftest_umma: ") + cudaGetErrorString(e)); return DP_ERR_CUDA; }
  return DP_OK;
}

}  // namespace dp
