// Evaluation kernels that follow the sampler: hypothesis mean, MPJPE and Procrustes-aligned MPJPE.
//   hypothesis mean   runners/diffpose_frame.py:382
//   mpjpe             common/loss.py:7-13 (after root-centring, runners/diffpose_frame.py:384-386)
//   p_mpjpe           common/loss.py:25-64 == common/utils.py:155-187 (numpy float64 SVD in the reference;
//                     here one warp per pose: fp32 shuffle reductions for the centring / cross-covariance, a one-sided
//                     Jacobi SVD of the 3x3 matrix in fp32 registers, fp64 for the partial sums over poses)
#include "dp_internal.h"
#include "dp_metrics_dev.cuh"

namespace dp {
namespace {

using metric::NP;

__global__ void hyp_mean_kernel(const float* __restrict__ x, float* __restrict__ out, long n_pose, int n_hyp,
                                int row_floats) {
  const long total = n_pose * row_floats;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int h = 0; h < n_hyp; ++h) s += x[(size_t)h * total + i];
    out[i] = s / (float)n_hyp;
  }
}

// One WARP per pose: lane j < 17 owns joint j of the prediction and the target; every per-pose sum is a warp shuffle
// reduction, and the 3x3 SVD -- a register-only Jacobi iteration -- runs redundantly on all lanes (no divergence, no
// local-memory arrays).  A block of 8 warps handles 8 poses, so 1024 poses already
// spread over 128 CTAs (the first version ran one thread per pose with 34x3 doubles of local memory on 8 CTAs: 44 us
// per 1024 poses, half a sampler launch).
constexpr int kMetricWarps = 8;
__global__ void __launch_bounds__(kMetricWarps * 32) metrics_kernel(const float* __restrict__ pred, int ps, int po, const float* __restrict__ gt, long n,
                                                                    double* __restrict__ sums, float* __restrict__ per_pose) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // programmatic dependent launch (the evaluation loop alternates sampler and metrics kernels on one stream): this grid may
  // have been launched while the sampler that produces `pred` was draining -- wait for its results, and let the next
  // sampler launch begin its own set-up right away
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  double e1s = 0.0, e2s = 0.0, cnt = 0.0;      // this warp's partial sums (identical on every lane)
  for (long i = (long)blockIdx.x * kMetricWarps + warp; i < n; i += (long)gridDim.x * kMetricWarps) {
    const int j = lane < NP ? lane : 0;
    const float* p = pred + (size_t)i * NP * ps + po + j * ps;
    const float* g = gt + (size_t)i * NP * 3 + j * 3;
    float pv[3], gv[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { pv[c] = p[c]; gv[c] = g[c]; }
    float e1f, e2f;
    metric::pose_errors(pv, gv, lane, e1f, e2f);
    const double e1 = (double)e1f, e2 = (double)e2f;
    e1s += e1; e2s += e2; cnt += 1.0;
    if (per_pose && lane == 0) { per_pose[2 * i] = (float)e1; per_pose[2 * i + 1] = (float)e2; }
  }
  // block reduction -> one atomic per block and quantity
  __shared__ double red[3][kMetricWarps];
  if (lane == 0) { red[0][warp] = e1s; red[1][warp] = e2s; red[2][warp] = cnt; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int w = 0; w < kMetricWarps; ++w) s += red[threadIdx.x][w];
    if (s != 0.0) atomicAdd(sums + threadIdx.x, s);
  }
}

// [uv | xyz - xyz[root]]: the glue between the lifter and the sampler for the engines whose forward kernel does not
// write it directly (runners/diffpose_frame.py:337-343 with the intended out-of-place root-centring)
__global__ void lift_glue_kernel(const float* __restrict__ uv, const float* __restrict__ xyz, float* __restrict__ out, long n, int n_pts,
                                 int c_in, int c_out) {
  const int wd = c_in + c_out;
  const long total = n * n_pts * wd;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const long row = i / wd;
    const int c = (int)(i - row * wd);
    const long pose = row / n_pts;
    out[i] = c < c_in ? uv[row * c_in + c] : __fsub_rn(xyz[row * c_out + (c - c_in)], xyz[pose * n_pts * c_out + (c - c_in)]);
  }
}

}  // namespace

int metrics_launch(const float* pred, int pred_stride, int pred_offset, const float* gt, long n, int n_pts,
                   double* sums, float* per_pose, cudaStream_t s) {
  (void)n_pts;
  long grid = (n + kMetricWarps - 1) / kMetricWarps;
  if (grid > 148 * 8) grid = 148 * 8;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(kMetricWarps * 32); cfg.dynamicSmemBytes = 0; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  DP_CUDA(cudaLaunchKernelEx(&cfg, metrics_kernel, pred, pred_stride, pred_offset, gt, n, sums, per_pose));
  count_launch();
  DP_CUDA(cudaGetLastError());
  return DP_OK;
}

int hyp_mean_launch(const float* x, float* out, long n_pose, int n_hyp, int row_floats, cudaStream_t s) {
  const long total = n_pose * row_floats;
  long grid = (total + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  hyp_mean_kernel<<<(unsigned)grid, 256, 0, s>>>(x, out, n_pose, n_hyp, row_floats);
  count_launch();
  DP_CUDA(cudaGetLastError());
  return DP_OK;
}

int lift_glue_launch(const float* uv, const float* xyz, float* out, long n, int n_pts, int c_in, int c_out, cudaStream_t s) {
  const long total = n * n_pts * (c_in + c_out);
  long grid = (total + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  lift_glue_kernel<<<(unsigned)grid, 256, 0, s>>>(uv, xyz, out, n, n_pts, c_in, c_out);
  count_launch();
  DP_CUDA(cudaGetLastError());
  return DP_OK;
}

}  // namespace dp
