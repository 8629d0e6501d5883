// Internal declarations shared by the translation units of libdiffpose_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>
#include "../../include/diffpose_b200.h"
#include "../../include/diffpose_b200_diag.h"

namespace dp {

constexpr int kMaxLayers = 16;
constexpr int kMaxPts = 17;       // the kernels are written for the 17-joint Human3.6M skeleton
constexpr int kMaxCoord = 8;      // c_in / c_out upper bound (uvxyz = 5)

struct Dims {
  int n_pts, c_in, c_out, hid, n_layer, n_head, has_temb;
};

// fp32 weights of one GraAttenLayer + _ResChebGC(_diff) pair, every matrix stored [K][N] row-major
// (K = input feature, N = output feature) so that a warp reads consecutive output features.
struct LayerW {
  const float *ln0_a, *ln0_b;   // atten_layers.l.sublayer.0.norm
  const float *wqkv, *bqkv;     // [hid][3hid] = linears.0|1|2 side by side
  const float *wo, *bo;         // [hid][hid]   linears.3
  const float *ln1_a, *ln1_b;   // sublayer.1.norm
  const float *lhat;            // [n_pts][n_pts]  D A_hat D  (GraFormer.py:174-178)
  const float *w1, *b1;         // [hid][2hid]  feed_forward.gconv1.fc
  const float *w2, *b2;         // [2hid][hid]  feed_forward.gconv2.fc
  const float *wc1, *bc1;       // [3hid][hid]  gconv_layers.l.gconv1.gconv  (k = cheb_order*hid + c)
  const float *wc2, *bc2;       // [3hid][hid]  gconv_layers.l.gconv2.gconv
  const float *wt, *bt;         // [4hid][hid]  gconv_layers.l.temb_proj (has_temb)
};

struct Weights {
  const float *win, *bin;       // [3c_in][hid]
  const float *wout, *bout;     // [3hid][c_out]
  const float *t1, *t2;         // Chebyshev T1 = L, T2 = 2L^2 - I, [n_pts][n_pts]
  const float *t1m, *t2m;       // the same matrices factored as diag(t?s) * t?m with fp16-exact t?m when the rows allow it
  const float *t1s, *t2s;       //   (row scales [n_pts]); used by the tensor-core engine, see integerise_rows in dp_api.cu
  const float *wd0, *bd0;       // [hid][4hid]  temb.dense.0
  const float *wd1, *bd1;       // [4hid][4hid] temb.dense.1
  LayerW layer[kMaxLayers];
};

void set_error(const std::string& msg);
void count_launch(int n = 1);

#define DP_CUDA(expr)                                                                              \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      dp::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                           \
      return DP_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

#define DP_REQUIRE(cond, msg)                                                                      \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      dp::set_error(std::string(msg));                                                             \
      return DP_ERR_INVALID;                                                                       \
    }                                                                                              \
  } while (0)

#define DP_TRY(expr)              \
  do {                            \
    int _rc = (expr);             \
    if (_rc != DP_OK) return _rc; \
  } while (0)

// DDIM step scalars travel as a by-value kernel argument (no host->device copy, graph friendly).
constexpr int kMaxInlineSteps = 64;
struct StepsArg {
  dp_step s[kMaxInlineSteps];
};

struct TcxPack;  // defined in dp_tcx.cu
struct Tc2Pack;  // defined in dp_tc2.cu

}  // namespace dp

// The opaque handle.
struct dp_model {
  dp::Dims d{};
  int device = 0;
  int sm_count = 0;
  int engine = DP_ENGINE_AUTO;
  bool packed = false;
  long n_params = 0;

  float* blob = nullptr;          // packed fp32 weights
  size_t blob_floats = 0;
  dp::Weights hw{};               // host copy of the device pointers
  dp::Weights* dw = nullptr;      // the same struct in device memory

  float* temb = nullptr;          // [n_rows][n_layer][hid] time-embedding table scratch
  size_t temb_cap = 0;            // capacity in floats
  std::vector<float> temb_t;      // timesteps the table currently holds (sampler schedule cache; empty = invalid)
  dp_step* steps = nullptr;       // device copy of the step scalars, only used when n_steps > kMaxInlineSteps
  size_t steps_cap = 0;
  std::vector<dp_step> steps_host;   // what `steps` currently holds (a long schedule is uploaded once, not per call)
  float* hyp_scratch = nullptr;   // [n_pose*n_hyp, n_pts, c]: hypothesis mean of the engines that do not fuse it (fp32, tcx)
  size_t hyp_cap = 0;
  float* lift_scratch = nullptr;  // [n, n_pts, c_out]: dp_lift on the engines that do not fuse the glue
  size_t lift_cap = 0;

  dp::TcxPack* tcx = nullptr;     // split-precision tensor-core engine state (hi/lo fp16 weight blocks, packed lazily)
  dp::Tc2Pack* tc2 = nullptr;     // default tensor-core engine state
  long last_launch[6] = {0, 0, 0, 0, 0, 0};
  long long* trace = nullptr;     // caller-owned device buffer for the hand-over timestamps (dp_set_trace)
  int trace_cap = 0;
};

namespace dp {
// dp_simt.cu
// time-embedding table [n_t][n_layer][hid]: t comes from t_dev[i*t_stride] or, when t_dev is NULL, from inl.s[i].t
int simt_temb(dp_model* m, const float* t_dev, int t_stride, const StepsArg* inl, long n_t, cudaStream_t s);
int simt_forward(dp_model* m, const float* x, const float* t, const unsigned char* mask, float* out, long n,
                 cudaStream_t s);
int simt_sample(dp_model* m, const float* x_in, int x_is_repeated, float* x_out, long n_pose, int n_hyp,
                const dp_step* steps_dev, const StepsArg* inl, int n_steps, const float* noise,
                const unsigned char* mask, cudaStream_t s);
int ensure_capacity(float** p, size_t* cap, size_t need_floats);
// dp_tcx.cu
bool tcx_supported(const Dims& d);
void tcx_invalidate(dp_model* m);   // after dp_pack: the split weights are rebuilt on the engine's next use
void tcx_free(dp_model* m);
int tcx_sample(dp_model* m, const float* x_in, int x_is_repeated, float* x_out, long n_pose, int n_hyp,
               const dp_step* steps_dev, const StepsArg* inl, int n_steps, const float* noise,
               const unsigned char* mask, cudaStream_t s);
int tcx_forward(dp_model* m, const float* x, const float* t, const unsigned char* mask, float* out, long n, int emit_uvxyz, cudaStream_t s);
// dp_lab.cu
int tc_lab(const void* image_dev, int image_bytes, const dp_mma_op* ops_host, int n_ops, float* out_dev, int ncols,
           const void* tmem_image_dev, int tmem_col0, int tmem_ncols, cudaStream_t s);
void tc_lab_cycles(long long* out2);
// dp_tc2.cu
bool tc2_supported(const Dims& d);
int tc2_forward(dp_model* m, const float* x, const float* t, const unsigned char* mask, float* out, long n, cudaStream_t s);
int tc2_pack(dp_model* m, cudaStream_t s);
// after simt_temb filled m->temb for a sampler schedule: the same embeddings as GC2 bias blocks of the tcg engine
int tc2_tau(dp_model* m, int n_steps, cudaStream_t s);
void tc2_free(dp_model* m);
// mean_over_hyp: x_out is [n_pose, n_pts, c], the hypothesis mean fused into the kernel's final store
// gt / sums (optional): fused evaluation tail, see dp_sample_eval
int tc2_sample(dp_model* m, const float* x_in, int x_is_repeated, float* x_out, long n_pose, int n_hyp,
               const dp_step* steps_dev, const StepsArg* inl, int n_steps, const float* noise,
               const unsigned char* mask, int mean_over_hyp, const float* gt, double* sums, cudaStream_t s);
// dp_metrics.cu
int metrics_launch(const float* pred, int pred_stride, int pred_offset, const float* gt, long n, int n_pts,
                   double* sums, float* per_pose, cudaStream_t s);
int hyp_mean_launch(const float* x, float* out, long n_pose, int n_hyp, int row_floats, cudaStream_t s);
// out[n][n_pts][c_in + c_out] = [uv | xyz - xyz[root]] (runners/diffpose_frame.py:337-343 with out-of-place root-centring)
int lift_glue_launch(const float* uv, const float* xyz, float* out, long n, int n_pts, int c_in, int c_out, cudaStream_t s);
}  // namespace dp
