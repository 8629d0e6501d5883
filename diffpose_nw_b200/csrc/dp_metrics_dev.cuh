// Per-pose MPJPE / P-MPJPE on one warp, shared by the stand-alone metrics kernel (dp_metrics.cu) and the sampler kernel's
// fused evaluation tail (dp_tc2.cu).
//   mpjpe    common/loss.py:7-13 (after root-centring, runners/diffpose_frame.py:384-386)
//   p_mpjpe  common/loss.py:25-64 == common/utils.py:155-187 (numpy float64 SVD in the reference)
#pragma once
#include <cuda_runtime.h>

namespace dp {
namespace metric {

constexpr int NP = 17;

template <typename T>
__device__ __forceinline__ void rot_cols(T a[3][3], int p, int q, T c, T s) {
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    const T ap = a[r][p], aq = a[r][q];
    a[r][p] = c * ap - s * aq;
    a[r][q] = s * ap + c * aq;
  }
}

// H = U diag(sv) V^T with sv sorted descending; U, V orthogonal (U completed by a cross product when rank < 3).
// One-sided Jacobi, entirely in registers.  T = float: H is the cross-covariance of two unit-norm centred point sets
// (entries <= 1), so an fp32 decomposition leaves the rotation accurate to ~2e-7 -- 1e-7 m on a pose -- against the 1e-6 m
// the golden vectors are compared at; the sums that feed H and the final partial sums stay fp64.
template <typename T>
__device__ __forceinline__ void svd3(const T h[3][3], T u[3][3], T sv[3], T v[3][3]) {
  // fp32: a rotation is skipped once the two columns are orthogonal to 1e-7 (cosine); Jacobi converges quadratically
  // (1e-1 -> 1e-2 -> 1e-4 -> 1e-8), so 6 sweeps always suffice for a 3x3 and the usual exit is the `off == 0` test after 4-5
  // (a 6e-8 threshold sat inside the rounding noise of the dot products: all 12 sweeps ran, 7 of the kernel's 10 us)
  const T kConv = sizeof(T) == 4 ? (T)1e-7 : (T)1e-15, kFloor = sizeof(T) == 4 ? (T)1e-37 : (T)1e-300, kRank = sizeof(T) == 4 ? (T)1e-6 : (T)1e-14;
  const int kSweeps = sizeof(T) == 4 ? 6 : 30;
  T a[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { a[i][j] = h[i][j]; v[i][j] = (i == j) ? (T)1 : (T)0; }
  for (int sweep = 0; sweep < kSweeps; ++sweep) {
    T off = 0;
#pragma unroll
    for (int pair = 0; pair < 3; ++pair) {
      const int p = pair == 2 ? 1 : 0, q = pair == 0 ? 1 : 2;
      T alpha = 0, beta = 0, gamma = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r) { alpha += a[r][p] * a[r][p]; beta += a[r][q] * a[r][q]; gamma += a[r][p] * a[r][q]; }
      const T lim = kConv * sqrt(alpha * beta);
      if (fabs(gamma) > lim && fabs(gamma) > kFloor) {
        off += fabs(gamma);
        const T zeta = (beta - alpha) / ((T)2 * gamma);
        const T t = (zeta >= 0 ? (T)1 : (T)-1) / (fabs(zeta) + sqrt((T)1 + zeta * zeta));
        const T c = (T)1 / sqrt((T)1 + t * t), s = c * t;
        rot_cols<T>(a, p, q, c, s);
        rot_cols<T>(v, p, q, c, s);
      }
    }
    if (off == (T)0) break;
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) sv[j] = sqrt(a[0][j] * a[0][j] + a[1][j] * a[1][j] + a[2][j] * a[2][j]);
  // sort columns by descending singular value (3-element network)
#pragma unroll
  for (int pass = 0; pass < 3; ++pass) {
    const int p = pass == 1 ? 1 : 0, q = p + 1;
    if (sv[p] < sv[q]) {
      T t = sv[p]; sv[p] = sv[q]; sv[q] = t;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        t = a[r][p]; a[r][p] = a[r][q]; a[r][q] = t;
        t = v[r][p]; v[r][p] = v[r][q]; v[r][q] = t;
      }
    }
  }
  const T tiny = kRank * (sv[0] > 0 ? sv[0] : (T)1);
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (sv[j] > tiny) {
#pragma unroll
      for (int r = 0; r < 3; ++r) u[r][j] = a[r][j] / sv[j];
    }
  }
  if (sv[1] <= tiny) {  // rank <= 1: any unit vector orthogonal to u0
    T ax = fabs(u[0][0]), ay = fabs(u[1][0]), az = fabs(u[2][0]);
    const int ei = (ax <= ay && ax <= az) ? 0 : (ay <= az ? 1 : 2);      // (selects, no indexed array: the callers must stay stack free)
    const T e0 = ei == 0 ? (T)1 : (T)0, e1 = ei == 1 ? (T)1 : (T)0, e2 = ei == 2 ? (T)1 : (T)0;
    T w0 = u[1][0] * e2 - u[2][0] * e1, w1 = u[2][0] * e0 - u[0][0] * e2, w2 = u[0][0] * e1 - u[1][0] * e0;
    const T n = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    u[0][1] = w0 / n; u[1][1] = w1 / n; u[2][1] = w2 / n;
  }
  if (sv[2] <= tiny) {
    u[0][2] = u[1][0] * u[2][1] - u[2][0] * u[1][1];
    u[1][2] = u[2][0] * u[0][1] - u[0][0] * u[2][1];
    u[2][2] = u[0][0] * u[1][1] - u[1][0] * u[0][1];
  }
}

template <typename T>
__device__ __forceinline__ T det3(const T m[3][3]) {
  return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
         m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp, lane j < 17 holds joint j of the prediction (pv) and of the target (gv), xyz in metres, NOT yet root-centred
// (lanes >= 17 pass joint 0 again; they take part in the shuffles only).  Returns on every lane the pose's MPJPE (e1) and
// Procrustes-aligned MPJPE (e2).  Both inputs are root-centred out of place first (runners/diffpose_frame.py:384-385, intended
// semantics).  Every per-pose sum is a warp shuffle reduction, the 3x3 SVD -- a register-only Jacobi iteration -- runs
// redundantly on all lanes (no divergence, no local-memory arrays).
__device__ __forceinline__ void pose_errors(float pv[3], float gv[3], int lane, float& e1_out, float& e2_out) {
  float d2 = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    pv[c] -= __shfl_sync(0xffffffffu, pv[c], 0);
    gv[c] -= __shfl_sync(0xffffffffu, gv[c], 0);
    const float df = pv[c] - gv[c];
    d2 += df * df;
  }
  const bool live = lane < NP;
  // mpjpe: mean over joints of the fp32 joint distance
  float acc = live ? sqrtf(d2) : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  e1_out = acc / (float)NP;
  // p_mpjpe (common/loss.py:25-64): X = target, Y = prediction.  All of it in fp32 registers + warp shuffles: coordinates
  // are ~0.3 m, so fp32 centring / norms / cross-covariance are good to ~1e-7 and the aligned joint errors to ~1e-7 m,
  // against the 1e-6 m the golden vectors are compared at and the 0.05 mm of the north-star tolerance; fp64 only for the
  // sums over poses below.  (The first warp-per-pose version kept fp64 throughout: 15 dependent fp64 divisions / square
  // roots and 18 two-word shuffle reductions per pose made it 14 us per 1024 poses.)
  float X[3], Y[3], mx[3], my[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) { X[c] = live ? gv[c] : 0.f; Y[c] = live ? pv[c] : 0.f; }
#pragma unroll
  for (int c = 0; c < 3; ++c) { mx[c] = warp_sum_f(X[c]) * (1.0f / NP); my[c] = warp_sum_f(Y[c]) * (1.0f / NP); }
  float a0[3], b0[3], nx = 0.f, ny = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    a0[c] = live ? X[c] - mx[c] : 0.f;
    b0[c] = live ? Y[c] - my[c] : 0.f;
    nx = fmaf(a0[c], a0[c], nx); ny = fmaf(b0[c], b0[c], ny);
  }
  nx = sqrtf(warp_sum_f(nx)); ny = sqrtf(warp_sum_f(ny));
  const float inx = 1.0f / nx, iny = 1.0f / ny;
#pragma unroll
  for (int c = 0; c < 3; ++c) { a0[c] *= inx; b0[c] *= iny; }
  float h[3][3];                                         // H = X0^T Y0 of the normalised, centred point sets
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) h[a][b] = warp_sum_f(a0[a] * b0[b]);
  float u[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}, v[3][3], sv[3];
  svd3<float>(h, u, sv, v);
  float r[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) r[a][b] = v[a][0] * u[b][0] + v[a][1] * u[b][1] + v[a][2] * u[b][2];
  const float dt = det3<float>(r);
  const float sg = dt > 0 ? 1.0f : (dt < 0 ? -1.0f : 0.0f);
#pragma unroll
  for (int a = 0; a < 3; ++a) v[a][2] *= sg;
  sv[2] *= sg;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) r[a][b] = v[a][0] * u[b][0] + v[a][1] * u[b][1] + v[a][2] * u[b][2];
  const float scale = (sv[0] + sv[1] + sv[2]) * nx * iny;
  float e = 0.f;
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    // aligned = scale * Y R + t with t = muX - scale * muY R   <=>   scale * (Y - muY) R + muX
    const float al = scale * ny * (b0[0] * r[0][b] + b0[1] * r[1][b] + b0[2] * r[2][b]) + mx[b];
    const float df = al - X[b];
    e = fmaf(df, df, e);
  }
  e2_out = warp_sum_f(live ? sqrtf(e) : 0.f) * (1.0f / NP);
}

}  // namespace metric
}  // namespace dp
