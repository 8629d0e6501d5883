// Tensor-core engine: persistent sm_100a kernel, one 128-row tile (7 poses x 17 joints) per CTA, every DDIM step
// and every layer executed without leaving the SM.
//
//   * dense projections (QKV, out-proj, GraphNet fc1/fc2, Chebyshev convs: 11.0 of the 12.3 MMAC per pose-forward)
//     run as tcgen05.mma kind::f16 (fp16 operands = 11-bit significand like TF32, fp32 accumulation in TMEM),
//     M=128, N=96, K=16 per instruction, operands in shared memory in the canonical K-major no-swizzle layout;
//   * biases ride in the MMA: a constant-one K slab of A times a (hi, lo) fp16 bias slab of the weight block;
//   * weights stream L2 -> shared memory as 21.5 KB pre-packed blocks through a 4-stage ring filled by the TMA
//     bulk-copy engine (cp.async.bulk + mbarrier complete_tx), issued by a dedicated producer warp;
//   * the 17x17 structures (Chebyshev T1/T2 gather, learnable-adjacency aggregation, per-head attention with
//     softmax), LayerNorm, residual stream and the DDIM update stay fp32 on the CUDA cores, in shared memory.
//
// Reference semantics: see the list at the top of dp_simt.cu (same functions, same file:line).
#include <cuda_fp16.h>
#include <cmath>
#include "dp_internal.h"
#include "dp_sm100.cuh"

namespace dp {

namespace {

using namespace sm100;

constexpr int NP = 17;
constexpr int TM = 128;              // tile rows = UMMA M
constexpr int TP = 7;                // poses per tile
constexpr int TR = TP * NP;          // 119 valid rows
constexpr int H = 96;
constexpr int XLD = 100;             // fp32 row stride (floats): thread-per-row float4 access is conflict free
constexpr int NSTAGE = 4;
constexpr int WK = 112;              // weight block K extent: 96 weights + 16 (bias slab; k=96 hi, k=97 lo)
constexpr int W_LBO = 12 * 128;      // bytes between K-adjacent 8x8 core matrices of a weight block
constexpr int W_SBO = 128;           // bytes between N-adjacent core matrices
constexpr int WBLK_BYTES = (WK / 8) * W_LBO;        // 21504
constexpr int A_LBO = 16 * 128 + 16; // 2064: +16 B skews consecutive K chunks across banks
constexpr int A_SBO = 128;
constexpr int ABLK_BYTES = 12 * A_LBO;              // 24768
constexpr int ONES_BYTES = 2 * A_LBO;               // 4128
constexpr int BLOCKS_PER_LAYER = 14;
constexpr int NNB = 9;               // max |2-hop neighbourhood| in the H36M tree (support of T2 = 2L^2 - I)
constexpr int kComputeThreads = 256;
constexpr int kThreads = kComputeThreads + 32;
constexpr int TMEM_COLS = 512;

// shared memory map (bytes)
constexpr int OFF_X = 0;                                   // fp32 residual stream [128][100]
constexpr int OFF_A = OFF_X + TM * XLD * 4;                // 51200: three fp16 operand blocks; fp32 scratch aliases them
constexpr int OFF_ONES = OFF_A + 3 * ABLK_BYTES;           // 125504
constexpr int OFF_W = (OFF_ONES + ONES_BYTES + 127) / 128 * 128;  // 129664
constexpr int OFF_XT = OFF_W + NSTAGE * WBLK_BYTES;        // 215744: x_t [128][8] fp32
constexpr int OFF_EP = OFF_XT + TM * 8 * 4;                // 219840: eps [128][8] fp32
constexpr int al16(int x) { return (x + 15) / 16 * 16; }
constexpr int OFF_NBI = OFF_EP + TM * 8 * 4;               // neighbour index  [17][9] int
constexpr int OFF_NBC = al16(OFF_NBI + NP * NNB * 4);      // neighbour coeffs [17][9] float2 (T1, T2)
constexpr int OFF_LH = al16(OFF_NBC + NP * NNB * 8);       // Lhat [17][17]
constexpr int OFF_TE = al16(OFF_LH + NP * NP * 4);         // temb of the current (step, layer) [96]
constexpr int OFF_MASK = OFF_TE + H * 4;                   // key mask [32]
constexpr int OFF_BAR = OFF_MASK + 128;                    // mbarriers: full[4], empty[4], mma_done
constexpr int OFF_TMEM = OFF_BAR + 128;
constexpr int SMEM_BYTES = OFF_TMEM + 16;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(OFF_W % 128 == 0 && OFF_A % 128 == 0 && OFF_ONES % 16 == 0 && OFF_BAR % 16 == 0 && OFF_NBC % 16 == 0 && OFF_XT % 16 == 0, "alignment");
static_assert(TR * XLD * 4 <= 2 * ABLK_BYTES, "fp32 scratch rows must not reach operand block 2");

// ------------------------------------------------------------------------------------------------ engine-specific helpers
// (mbarrier / TMA / tcgen05 wrappers: dp_sm100.cuh)
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
// Instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 (bit 4), A=B=f16 (0), K-major both, N>>3 at 17, M>>4 at 24
constexpr uint32_t kIdescN96 = (1u << 4) | ((96u >> 3) << 17) | ((128u >> 4) << 24);

// byte offset of the 16-byte chunk holding elements (row, 8*kc .. 8*kc+7) inside an fp16 operand block
__device__ __forceinline__ uint32_t a_chunk(int row, int kc) { return kc * A_LBO + (row >> 3) * A_SBO + (row & 7) * 16; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct TcArgs {
  const Weights* w;          // fp32 blob (LayerNorm, Lhat, biases added on CUDA cores, in/out convolutions)
  const uint8_t* wpack;      // fp16 weight blocks [n_layer][14][21504 B]
  int n_layer;
  const float* x_in;
  int x_is_repeated;
  float* out;
  long n_rows, n_pose;
  int n_steps;
  const float* temb;         // [n_steps][n_layer][96]
  const float* noise;
  const unsigned char* mask;
  const dp_step* steps_dev;
};

// One GEMM = a few weight blocks.  Issued by a single thread; D columns / accumulate flags per block.
struct Pipe {
  uint32_t full0, empty0, done;   // smem addresses of the barriers
  uint32_t stage, phase;          // weight ring position (thread 0 and the producer keep their own copy)
  uint32_t done_phase;            // every compute thread tracks the parity of the "GEMM finished" barrier
};

__device__ __forceinline__ void issue_block(Pipe& p, uint32_t smem_base, uint32_t tmem_d, int a_blk, bool accumulate, bool bias) {
  mbar_wait(p.full0 + 8 * p.stage, p.phase);
  tc_fence_after();
  const uint32_t wa = smem_base + OFF_W + p.stage * WBLK_BYTES;
  const uint32_t aa = smem_base + OFF_A + a_blk * ABLK_BYTES;
#pragma unroll
  for (int ks = 0; ks < 6; ++ks)
    umma_f16(tmem_d, make_desc(aa + ks * 2 * A_LBO, A_LBO, A_SBO), make_desc(wa + ks * 2 * W_LBO, W_LBO, W_SBO), kIdescN96,
             (accumulate || ks > 0) ? 1u : 0u);
  if (bias)
    umma_f16(tmem_d, make_desc(smem_base + OFF_ONES, A_LBO, A_SBO), make_desc(wa + 12 * W_LBO, W_LBO, W_SBO), kIdescN96, 1u);
  umma_commit(p.empty0 + 8 * p.stage);   // the stage is free once these MMAs have read it
  if (++p.stage == NSTAGE) { p.stage = 0; p.phase ^= 1; }
}

// Epilogue plumbing: warp w reads TMEM lanes 32*(w&3).. (its rows) and the column half (w>>2).
enum EpiKind { EPI_F16 = 0, EPI_RELU_F16 = 1, EPI_XADD = 2, EPI_XADD_RELU = 3, EPI_F32 = 4, EPI_RELU_TEMB_F16 = 5 };

template <int NCOLS, int KIND>
__device__ __forceinline__ void epilogue(uint8_t* smem, uint32_t tmem_base, const float* __restrict__ temb) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = (warp & 3) * 32 + lane;
  const int c0 = (warp >> 2) * (NCOLS / 2);
  const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  float* X = reinterpret_cast<float*>(smem + OFF_X) + row * XLD;
  float* Z = reinterpret_cast<float*>(smem + OFF_A) + row * XLD;
#pragma unroll
  for (int cc = 0; cc < NCOLS / 2; cc += 16) {
    const int c = c0 + cc;
    float v[16];
    tmem_ld16(taddr + c, v);
    if (KIND == EPI_RELU_F16 || KIND == EPI_XADD_RELU || KIND == EPI_RELU_TEMB_F16) {
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
    }
    if (KIND == EPI_RELU_TEMB_F16) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(temb + c + i);
        v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
      }
    }
    if (KIND == EPI_F16 || KIND == EPI_RELU_F16 || KIND == EPI_RELU_TEMB_F16) {
      const int blk = c / H, kc = (c % H) >> 3;
      uint8_t* dst = smem + OFF_A + blk * ABLK_BYTES;
      *reinterpret_cast<uint4*>(dst + a_chunk(row, kc)) = pack8(v);
      *reinterpret_cast<uint4*>(dst + a_chunk(row, kc + 1)) = pack8(v + 8);
    } else if (KIND == EPI_XADD || KIND == EPI_XADD_RELU) {
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        float4 x = *reinterpret_cast<float4*>(X + c + i);
        x.x += v[i]; x.y += v[i + 1]; x.z += v[i + 2]; x.w += v[i + 3];
        *reinterpret_cast<float4*>(X + c + i) = x;
      }
    } else {  // EPI_F32 -> scratch Z
#pragma unroll
      for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(Z + c + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
    }
  }
}

// after the A operand was written with ordinary stores: make it visible to the tensor core (async proxy), then sync
__device__ __forceinline__ void publish_operand() {
  fence_async_smem();
  tc_fence_before();
  bar_compute();
}
__device__ __forceinline__ void wait_gemm(Pipe& p) {
  mbar_wait(p.done, p.done_phase);
  p.done_phase ^= 1;
  tc_fence_after();
}

// LayerNorm of the residual stream (GraFormer.py:67-70): lanes (2r, 2r+1) own the two 48-channel halves of row r.
// TO_F16: write the fp16 operand block `blk` (all 128 rows); else write fp32 rows (stride XLD) at OFF_A (valid rows).
template <bool TO_F16>
__device__ __forceinline__ void layer_norm_tile(uint8_t* smem, int blk, const float* __restrict__ ga, const float* __restrict__ gb) {
  const int row = threadIdx.x >> 1, hh = threadIdx.x & 1;
  const float* xr = reinterpret_cast<const float*>(smem + OFF_X) + row * XLD + hh * 48;
  float v[48];
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    const float4 u = *reinterpret_cast<const float4*>(xr + 4 * q);
    v[4 * q] = u.x; v[4 * q + 1] = u.y; v[4 * q + 2] = u.z; v[4 * q + 3] = u.w;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 48; ++i) s += v[i];
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  const float mean = s * (1.0f / (float)H);
  float q2 = 0.f;
#pragma unroll
  for (int i = 0; i < 48; ++i) { v[i] -= mean; q2 = fmaf(v[i], v[i], q2); }
  q2 += __shfl_xor_sync(0xffffffffu, q2, 1);
  const float inv = 1.0f / (sqrtf(q2 * (1.0f / (float)(H - 1))) + 1e-6f);
  if (!TO_F16 && row >= TR) return;
#pragma unroll
  for (int q = 0; q < 12; ++q) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(ga + hh * 48) + q), b = __ldg(reinterpret_cast<const float4*>(gb + hh * 48) + q);
    v[4 * q] = fmaf(a.x * inv, v[4 * q], b.x); v[4 * q + 1] = fmaf(a.y * inv, v[4 * q + 1], b.y);
    v[4 * q + 2] = fmaf(a.z * inv, v[4 * q + 2], b.z); v[4 * q + 3] = fmaf(a.w * inv, v[4 * q + 3], b.w);
  }
  if (TO_F16) {
#pragma unroll
    for (int c = 0; c < 6; ++c) *reinterpret_cast<uint4*>(smem + OFF_A + blk * ABLK_BYTES + a_chunk(row, hh * 6 + c)) = pack8(v + 8 * c);
  } else {
    float* y = reinterpret_cast<float*>(smem + OFF_A) + row * XLD + hh * 48;
#pragma unroll
    for (int q = 0; q < 12; ++q) *reinterpret_cast<float4*>(y + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
  }
}

// Multi-head attention over the joints of each pose (GraFormer.py:99-113), fp16 q/k/v in operand blocks 0/1/2,
// output written in place of q.  One thread per (head, pose, query joint).
__device__ __forceinline__ void attention_tile(uint8_t* smem, int npose) {
  const float* maskf = reinterpret_cast<const float*>(smem + OFF_MASK);
  const uint8_t* Q = smem + OFF_A;
  const uint8_t* K = Q + ABLK_BYTES;
  const uint8_t* V = K + ABLK_BYTES;
  const float scale = 1.0f / sqrtf(24.0f);
  const int per_head = npose * NP;
  for (int task = threadIdx.x; task < 4 * per_head; task += kComputeThreads) {
    const int h = task / per_head, rr = task - h * per_head;   // rr = pose*17 + joint = tile row
    const int p = rr / NP;
    float q[24];
#pragma unroll
    for (int c = 0; c < 3; ++c) unpack8(*reinterpret_cast<const uint4*>(Q + a_chunk(rr, 3 * h + c)), q + 8 * c);
    float sc[NP];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int rj = p * NP + j;
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float kv[8];
        unpack8(*reinterpret_cast<const uint4*>(K + a_chunk(rj, 3 * h + c)), kv);
#pragma unroll
        for (int e = 0; e < 8; ++e) s = fmaf(q[8 * c + e], kv[e], s);
      }
      s = s * scale;
      if (maskf[j] == 0.f) s = -1e9f;
      sc[j] = s;
      mx = fmaxf(mx, s);
    }
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < NP; ++j) { sc[j] = __expf(sc[j] - mx); sum += sc[j]; }
    const float inv = 1.0f / sum;
    float o[24];
#pragma unroll
    for (int e = 0; e < 24; ++e) o[e] = 0.f;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int rj = p * NP + j;
      const float pj = sc[j] * inv;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float vv[8];
        unpack8(*reinterpret_cast<const uint4*>(V + a_chunk(rj, 3 * h + c)), vv);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[8 * c + e] = fmaf(pj, vv[e], o[8 * c + e]);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) *reinterpret_cast<uint4*>(smem + OFF_A + a_chunk(rr, 3 * h + c)) = pack8(o + 8 * c);
  }
}

// out[i] = sum_j Lhat[i][j] Y[j]  over the pose of row i (GraFormer.py:174-186), Y fp32 scratch -> fp16 block `blk`
__device__ __forceinline__ void lhat_to_operand(uint8_t* smem, int blk) {
  const int row = threadIdx.x & 127;
  if (row >= TR) {
    for (int kc = threadIdx.x >> 7; kc < 12; kc += 2) *reinterpret_cast<uint4*>(smem + OFF_A + blk * ABLK_BYTES + a_chunk(row, kc)) = make_uint4(0, 0, 0, 0);
    return;
  }
  const float* Y = reinterpret_cast<const float*>(smem + OFF_A);
  const float* lh = reinterpret_cast<const float*>(smem + OFF_LH);
  const int p = row / NP, i = row - p * NP;
  float co[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) co[j] = lh[i * NP + j];
  for (int kc = threadIdx.x >> 7; kc < 12; kc += 2) {
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const float* src = Y + (p * NP + j) * XLD + kc * 8;
      const float4 u0 = *reinterpret_cast<const float4*>(src), u1 = *reinterpret_cast<const float4*>(src + 4);
      acc[0] = fmaf(co[j], u0.x, acc[0]); acc[1] = fmaf(co[j], u0.y, acc[1]); acc[2] = fmaf(co[j], u0.z, acc[2]); acc[3] = fmaf(co[j], u0.w, acc[3]);
      acc[4] = fmaf(co[j], u1.x, acc[4]); acc[5] = fmaf(co[j], u1.y, acc[5]); acc[6] = fmaf(co[j], u1.z, acc[6]); acc[7] = fmaf(co[j], u1.w, acc[7]);
    }
    *reinterpret_cast<uint4*>(smem + OFF_A + blk * ABLK_BYTES + a_chunk(row, kc)) = pack8(acc);
  }
}

// X[i] += sum_j Lhat[i][j] Z[j] + b2   (second LAM_Gconv with fc2 commuted in front of the aggregation)
__device__ __forceinline__ void lhat_residual(uint8_t* smem, const float* __restrict__ b2) {
  const int row = threadIdx.x & 127;
  if (row >= TR) return;
  const float* Z = reinterpret_cast<const float*>(smem + OFF_A);
  float* X = reinterpret_cast<float*>(smem + OFF_X) + row * XLD;
  const float* lh = reinterpret_cast<const float*>(smem + OFF_LH);
  const int p = row / NP, i = row - p * NP;
  float co[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) co[j] = lh[i * NP + j];
  for (int kc = threadIdx.x >> 7; kc < 12; kc += 2) {
    float acc[8];
    const float4 bb0 = __ldg(reinterpret_cast<const float4*>(b2 + kc * 8)), bb1 = __ldg(reinterpret_cast<const float4*>(b2 + kc * 8 + 4));
    acc[0] = bb0.x; acc[1] = bb0.y; acc[2] = bb0.z; acc[3] = bb0.w; acc[4] = bb1.x; acc[5] = bb1.y; acc[6] = bb1.z; acc[7] = bb1.w;
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const float* src = Z + (p * NP + j) * XLD + kc * 8;
      const float4 u0 = *reinterpret_cast<const float4*>(src), u1 = *reinterpret_cast<const float4*>(src + 4);
      acc[0] = fmaf(co[j], u0.x, acc[0]); acc[1] = fmaf(co[j], u0.y, acc[1]); acc[2] = fmaf(co[j], u0.z, acc[2]); acc[3] = fmaf(co[j], u0.w, acc[3]);
      acc[4] = fmaf(co[j], u1.x, acc[4]); acc[5] = fmaf(co[j], u1.y, acc[5]); acc[6] = fmaf(co[j], u1.z, acc[6]); acc[7] = fmaf(co[j], u1.w, acc[7]);
    }
    float4 x0 = *reinterpret_cast<float4*>(X + kc * 8), x1 = *reinterpret_cast<float4*>(X + kc * 8 + 4);
    x0.x += acc[0]; x0.y += acc[1]; x0.z += acc[2]; x0.w += acc[3]; x1.x += acc[4]; x1.y += acc[5]; x1.z += acc[6]; x1.w += acc[7];
    *reinterpret_cast<float4*>(X + kc * 8) = x0;
    *reinterpret_cast<float4*>(X + kc * 8 + 4) = x1;
  }
}

// Chebyshev input panel [V | T1 V | T2 V] (ChebConv.py:74-112) as three fp16 operand blocks.
// FROM_X: V is the fp32 residual stream (block 0 is written too); else V is the fp16 block 0 (hidden activation).
template <bool FROM_X>
__device__ __forceinline__ void cheb_concat_tile(uint8_t* smem) {
  const int row = threadIdx.x & 127;
  uint8_t* A0 = smem + OFF_A;
  if (row >= TR) {
    for (int kc = threadIdx.x >> 7; kc < 12; kc += 2) {
      const uint4 z = make_uint4(0, 0, 0, 0);
      if (FROM_X) *reinterpret_cast<uint4*>(A0 + a_chunk(row, kc)) = z;
      *reinterpret_cast<uint4*>(A0 + ABLK_BYTES + a_chunk(row, kc)) = z;
      *reinterpret_cast<uint4*>(A0 + 2 * ABLK_BYTES + a_chunk(row, kc)) = z;
    }
    return;
  }
  const int* nbi = reinterpret_cast<const int*>(smem + OFF_NBI);
  const float2* nbc = reinterpret_cast<const float2*>(smem + OFF_NBC);
  const float* X = reinterpret_cast<const float*>(smem + OFF_X);
  const int p = row / NP, i = row - p * NP;
  int nj[NNB];
  float2 nc[NNB];
#pragma unroll
  for (int n = 0; n < NNB; ++n) { nj[n] = p * NP + nbi[i * NNB + n]; nc[n] = nbc[i * NNB + n]; }
  for (int kc = threadIdx.x >> 7; kc < 12; kc += 2) {
    float a1[8], a2[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { a1[e] = 0.f; a2[e] = 0.f; }
#pragma unroll
    for (int n = 0; n < NNB; ++n) {
      float u[8];
      if (FROM_X) {
        const float* src = X + nj[n] * XLD + kc * 8;
        const float4 u0 = *reinterpret_cast<const float4*>(src), u1 = *reinterpret_cast<const float4*>(src + 4);
        u[0] = u0.x; u[1] = u0.y; u[2] = u0.z; u[3] = u0.w; u[4] = u1.x; u[5] = u1.y; u[6] = u1.z; u[7] = u1.w;
      } else {
        unpack8(*reinterpret_cast<const uint4*>(A0 + a_chunk(nj[n], kc)), u);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) { a1[e] = fmaf(nc[n].x, u[e], a1[e]); a2[e] = fmaf(nc[n].y, u[e], a2[e]); }
    }
    if (FROM_X) {
      const float* src = X + row * XLD + kc * 8;
      float u[8];
      const float4 u0 = *reinterpret_cast<const float4*>(src), u1 = *reinterpret_cast<const float4*>(src + 4);
      u[0] = u0.x; u[1] = u0.y; u[2] = u0.z; u[3] = u0.w; u[4] = u1.x; u[5] = u1.y; u[6] = u1.z; u[7] = u1.w;
      *reinterpret_cast<uint4*>(A0 + a_chunk(row, kc)) = pack8(u);
    }
    *reinterpret_cast<uint4*>(A0 + ABLK_BYTES + a_chunk(row, kc)) = pack8(a1);
    *reinterpret_cast<uint4*>(A0 + 2 * ABLK_BYTES + a_chunk(row, kc)) = pack8(a2);
  }
}

__global__ void __launch_bounds__(kThreads, 1) tc_kernel(TcArgs a, StepsArg inl) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  float* X = reinterpret_cast<float*>(smem + OFF_X);
  float* xt = reinterpret_cast<float*>(smem + OFF_XT);
  float* ep = reinterpret_cast<float*>(smem + OFF_EP);
  float* lh = reinterpret_cast<float*>(smem + OFF_LH);
  float* maskf = reinterpret_cast<float*>(smem + OFF_MASK);
  int* nbi = reinterpret_cast<int*>(smem + OFF_NBI);
  float2* nbc = reinterpret_cast<float2*>(smem + OFF_NBC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  const Weights& w = *a.w;

  Pipe pp;
  pp.full0 = sbase + OFF_BAR; pp.empty0 = sbase + OFF_BAR + 32; pp.done = sbase + OFF_BAR + 64;
  pp.stage = 0; pp.phase = 0; pp.done_phase = 0;

  // ---------------------------------------------------------------- one-time setup
  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(pp.full0 + 8 * s, 1); mbar_init(pp.empty0 + 8 * s, 1); }
    mbar_init(pp.done, 1);
    fence_mbar_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(sbase + OFF_TMEM, TMEM_COLS);
  // constant-one K slab: element (row, 0) = (row, 1) = 1, the other 14 of the 16 columns are 0
  for (int i = tid; i < 2 * TM; i += kThreads) {
    const int row = i & 127, kc = i >> 7;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (kc == 0) v.x = pack2(1.0f, 1.0f);
    *reinterpret_cast<uint4*>(smem + OFF_ONES + a_chunk(row, kc)) = v;
  }
  if (tid < 32) maskf[tid] = (tid < NP && a.mask && a.mask[tid] == 0) ? 0.f : 1.f;
  if (tid < NP) {
    // neighbour list of joint tid: columns where T1 or T2 is non-zero, padded with (self, 0, 0)
    int n = 0;
    for (int j = 0; j < NP; ++j) {
      const float c1 = __ldg(w.t1 + tid * NP + j), c2 = __ldg(w.t2 + tid * NP + j);
      if ((c1 != 0.f || c2 != 0.f) && n < NNB) { nbi[tid * NNB + n] = j; nbc[tid * NNB + n] = make_float2(c1, c2); ++n; }
    }
    for (; n < NNB; ++n) { nbi[tid * NNB + n] = tid; nbc[tid * NNB + n] = make_float2(0.f, 0.f); }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const long n_tiles = (a.n_rows + TP - 1) / TP;
  const int L = a.n_layer;

  if (warp == 8) {
    // ---------------------------------------------------------------- weight producer (TMA bulk copies)
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int step = 0; step < a.n_steps; ++step)
          for (int blk = 0; blk < L * BLOCKS_PER_LAYER; ++blk) {
            mbar_wait_sleep(pp.empty0 + 8 * stage, phase ^ 1);
            mbar_expect_tx(pp.full0 + 8 * stage, WBLK_BYTES);
            bulk_g2s(sbase + OFF_W + stage * WBLK_BYTES, a.wpack + (size_t)blk * WBLK_BYTES, WBLK_BYTES, pp.full0 + 8 * stage);
            if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
          }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------- compute warps
    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const long g0 = tile * TP;
      const int npose = (int)min((long)TP, a.n_rows - g0);
      const int R = npose * NP;
      for (int idx = tid; idx < TM * 8; idx += kComputeThreads) {
        const int r = idx >> 3, c = idx & 7;
        float v = 0.f;
        if (r < R && c < 5) {
          const long g = g0 + r / NP;
          const long src = a.x_is_repeated ? g : (g % a.n_pose);
          v = a.x_in[(src * NP + (r % NP)) * 5 + c];
        }
        xt[idx] = v;
      }
      bar_compute();

      for (int step = 0; step < a.n_steps; ++step) {
        // ---- input ChebConv (K = 15): fp32 on the CUDA cores.  B[row][0:15] = [x | T1 x | T2 x] in the eps scratch rows
        float* Bin = reinterpret_cast<float*>(smem + OFF_A);   // [128][16]
        for (int idx = tid; idx < TM * 5; idx += kComputeThreads) {
          const int r = idx / 5, c = idx - r * 5;
          float v0 = 0.f, v1 = 0.f, v2 = 0.f;
          if (r < TR) {
            const int p = r / NP, i = r - p * NP;
            v0 = xt[r * 8 + c];
#pragma unroll
            for (int n = 0; n < NNB; ++n) {
              const float u = xt[(p * NP + nbi[i * NNB + n]) * 8 + c];
              const float2 cf = nbc[i * NNB + n];
              v1 = fmaf(cf.x, u, v1);
              v2 = fmaf(cf.y, u, v2);
            }
          }
          Bin[r * 16 + c] = v0; Bin[r * 16 + 5 + c] = v1; Bin[r * 16 + 10 + c] = v2;
        }
        bar_compute();
        {
          const int row = tid & 127, hh = tid >> 7;
          float acc[48];
#pragma unroll
          for (int g = 0; g < 12; ++g) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(w.bin + hh * 48) + g);
            acc[4 * g] = b4.x; acc[4 * g + 1] = b4.y; acc[4 * g + 2] = b4.z; acc[4 * g + 3] = b4.w;
          }
          for (int k = 0; k < 15; ++k) {
            const float bv = Bin[row * 16 + k];
#pragma unroll
            for (int g = 0; g < 12; ++g) {
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(w.win + k * H + hh * 48) + g);
              acc[4 * g] = fmaf(bv, w4.x, acc[4 * g]); acc[4 * g + 1] = fmaf(bv, w4.y, acc[4 * g + 1]);
              acc[4 * g + 2] = fmaf(bv, w4.z, acc[4 * g + 2]); acc[4 * g + 3] = fmaf(bv, w4.w, acc[4 * g + 3]);
            }
          }
#pragma unroll
          for (int g = 0; g < 12; ++g)
            *reinterpret_cast<float4*>(X + row * XLD + hh * 48 + 4 * g) = make_float4(acc[4 * g], acc[4 * g + 1], acc[4 * g + 2], acc[4 * g + 3]);
        }
        bar_compute();

        for (int l = 0; l < L; ++l) {
          const LayerW& Lw = w.layer[l];
          for (int i = tid; i < NP * NP; i += kComputeThreads) lh[i] = __ldg(Lw.lhat + i);
          if (tid < H) reinterpret_cast<float*>(smem + OFF_TE)[tid] = __ldg(a.temb + ((size_t)step * L + l) * H + tid);
          // ======== x = x + attn(LN0(x))
          layer_norm_tile<true>(smem, 2, Lw.ln0_a, Lw.ln0_b);
          publish_operand();
          if (tid == 0) {
            tc_fence_after();
            issue_block(pp, sbase, tmem_base + 0, 2, false, true);     // Q
            issue_block(pp, sbase, tmem_base + 96, 2, false, true);    // K
            issue_block(pp, sbase, tmem_base + 192, 2, false, true);   // V
            umma_commit(pp.done);
          }
          wait_gemm(pp);
          epilogue<288, EPI_F16>(smem, tmem_base, nullptr);
          tc_fence_before();
          bar_compute();
          attention_tile(smem, TP);
          publish_operand();
          if (tid == 0) {
            tc_fence_after();
            issue_block(pp, sbase, tmem_base, 0, false, true);         // out projection
            umma_commit(pp.done);
          }
          wait_gemm(pp);
          epilogue<96, EPI_XADD>(smem, tmem_base, nullptr);
          tc_fence_before();
          bar_compute();
          // ======== x = x + GraphNet(LN1(x))
          layer_norm_tile<false>(smem, 0, Lw.ln1_a, Lw.ln1_b);
          bar_compute();
          lhat_to_operand(smem, 2);
          publish_operand();
          if (tid == 0) {
            tc_fence_after();
            issue_block(pp, sbase, tmem_base + 0, 2, false, true);     // fc1, outputs 0..95
            issue_block(pp, sbase, tmem_base + 96, 2, false, true);    // fc1, outputs 96..191
            umma_commit(pp.done);
          }
          wait_gemm(pp);
          epilogue<192, EPI_RELU_F16>(smem, tmem_base, nullptr);
          publish_operand();
          if (tid == 0) {
            tc_fence_after();
            issue_block(pp, sbase, tmem_base, 0, false, false);        // fc2, inputs 0..95
            issue_block(pp, sbase, tmem_base, 1, true, false);         // fc2, inputs 96..191
            umma_commit(pp.done);
          }
          wait_gemm(pp);
          epilogue<96, EPI_F32>(smem, tmem_base, nullptr);
          tc_fence_before();
          bar_compute();
          lhat_residual(smem, Lw.b2);
          bar_compute();
          // ======== x = x + GC2(GC1(x) + temb)
          cheb_concat_tile<true>(smem);
          publish_operand();
          if (tid == 0) {
            tc_fence_after();
            issue_block(pp, sbase, tmem_base, 0, false, true);
            issue_block(pp, sbase, tmem_base, 1, true, false);
            issue_block(pp, sbase, tmem_base, 2, true, false);
            umma_commit(pp.done);
          }
          wait_gemm(pp);
          epilogue<96, EPI_RELU_TEMB_F16>(smem, tmem_base, reinterpret_cast<const float*>(smem + OFF_TE));
          tc_fence_before();
          bar_compute();
          cheb_concat_tile<false>(smem);
          publish_operand();
          if (tid == 0) {
            tc_fence_after();
            issue_block(pp, sbase, tmem_base, 0, false, true);
            issue_block(pp, sbase, tmem_base, 1, true, false);
            issue_block(pp, sbase, tmem_base, 2, true, false);
            umma_commit(pp.done);
          }
          wait_gemm(pp);
          epilogue<96, EPI_XADD_RELU>(smem, tmem_base, nullptr);
          tc_fence_before();
          bar_compute();
        }

        // ---- output ChebConv (N = 5): U_k = X Wout_k on the CUDA cores, then eps = b + U0 + T1 U1 + T2 U2
        {
          float* U = reinterpret_cast<float*>(smem + OFF_A);   // [2][128][16]
          const int row = tid & 127, hh = tid >> 7;
          float acc[15];
#pragma unroll
          for (int i = 0; i < 15; ++i) acc[i] = 0.f;
          for (int cq = 0; cq < 12; ++cq) {
            const float4 xv = *reinterpret_cast<const float4*>(X + row * XLD + hh * 48 + cq * 4);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int c = hh * 48 + cq * 4 + e;
#pragma unroll
              for (int k3 = 0; k3 < 3; ++k3)
#pragma unroll
                for (int n = 0; n < 5; ++n) acc[k3 * 5 + n] = fmaf(xs[e], __ldg(w.wout + (k3 * H + c) * 5 + n), acc[k3 * 5 + n]);
            }
          }
#pragma unroll
          for (int i = 0; i < 15; ++i) U[(hh * TM + row) * 16 + i] = acc[i];
        }
        bar_compute();
        {
          const float* U = reinterpret_cast<const float*>(smem + OFF_A);
          for (int idx = tid; idx < TR * 5; idx += kComputeThreads) {
            const int r = idx / 5, n = idx - r * 5;
            const int p = r / NP, i = r - p * NP;
            float v = __ldg(w.bout + n) + U[r * 16 + n] + U[(TM + r) * 16 + n];
#pragma unroll
            for (int q = 0; q < NNB; ++q) {
              const int rj = p * NP + nbi[i * NNB + q];
              const float2 cf = nbc[i * NNB + q];
              v = fmaf(cf.x, U[rj * 16 + 5 + n] + U[(TM + rj) * 16 + 5 + n], v);
              v = fmaf(cf.y, U[rj * 16 + 10 + n] + U[(TM + rj) * 16 + 10 + n], v);
            }
            ep[r * 8 + n] = v;
          }
        }
        bar_compute();
        // ---- DDIM update (common/utils_diff.py:59-65), same operation order, no FMA contraction
        {
          const dp_step st = a.steps_dev ? a.steps_dev[step] : inl.s[step];
          for (int idx = tid; idx < R * 5; idx += kComputeThreads) {
            const int r = idx / 5, c = idx - r * 5;
            const float et = ep[r * 8 + c], xv = xt[r * 8 + c];
            const float x0 = __fdiv_rn(__fsub_rn(xv, __fmul_rn(et, st.sqrt_1m_at)), st.sqrt_at);
            float nx = __fmul_rn(st.sqrt_an, x0);
            if (a.noise) {
              const float z = a.noise[((size_t)step * a.n_rows + g0) * NP * 5 + idx];
              nx = __fadd_rn(nx, __fmul_rn(st.c1, z));
            }
            xt[r * 8 + c] = __fadd_rn(nx, __fmul_rn(st.c2, et));
          }
        }
        bar_compute();
      }
      for (int idx = tid; idx < R * 5; idx += kComputeThreads) a.out[(size_t)g0 * NP * 5 + idx] = xt[(idx / 5) * 8 + idx % 5];
      bar_compute();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, TMEM_COLS); }
}

// fp32 [K][N] panels of the fp32 blob -> fp16 weight block in the canonical K-major no-swizzle UMMA layout.
// block element (n, k): n in [0,96) output feature, k in [0,112): k < 96 weight W[k0+k][n0+n]; k = 96/97 bias hi/lo.
__global__ void tc_pack_block_kernel(uint8_t* __restrict__ dst, const float* __restrict__ W, int ldw, int k0, int n0,
                                     const float* __restrict__ bias) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 96 * WK; idx += gridDim.x * blockDim.x) {
    const int n = idx / WK, k = idx - n * WK;
    float v = 0.f;
    if (k < 96) v = W[(size_t)(k0 + k) * ldw + n0 + n];
    else if (bias != nullptr && k == 96) v = __half2float(__float2half_rn(bias[n0 + n]));
    else if (bias != nullptr && k == 97) { const float b = bias[n0 + n]; v = b - __half2float(__float2half_rn(b)); }
    const size_t off = (size_t)(k >> 3) * W_LBO + (size_t)(n >> 3) * W_SBO + (n & 7) * 16 + (k & 7) * 2;
    *reinterpret_cast<__half*>(dst + off) = __float2half_rn(v);
  }
}

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

struct TcPack {
  uint8_t* blocks = nullptr;   // [n_layer][14][WBLK_BYTES]
  size_t bytes = 0;
};

bool tc_supported(const Dims& d) {
  return d.hid == 96 && d.n_head == 4 && d.n_pts == 17 && d.c_in == 5 && d.c_out == 5 && d.has_temb == 1;
}

void tc_free(dp_model* m) {
  if (m->tc) {
    if (m->tc->blocks) cudaFree(m->tc->blocks);
    delete m->tc;
    m->tc = nullptr;
  }
}

static int pack_block(uint8_t* dst, const float* W, int ldw, int k0, int n0, const float* bias, cudaStream_t s) {
  tc_pack_block_kernel<<<12, 256, 0, s>>>(dst, W, ldw, k0, n0, bias);
  count_launch();
  DP_CUDA(cudaGetLastError());
  return DP_OK;
}

int tc_pack(dp_model* m, cudaStream_t s) {
  const Dims& d = m->d;
  if (!m->tc) m->tc = new TcPack();
  const size_t need = (size_t)d.n_layer * BLOCKS_PER_LAYER * WBLK_BYTES;
  if (m->tc->bytes < need) {
    if (m->tc->blocks) cudaFree(m->tc->blocks);
    m->tc->blocks = nullptr; m->tc->bytes = 0;
    DP_CUDA(cudaMalloc(reinterpret_cast<void**>(&m->tc->blocks), need));
    m->tc->bytes = need;
  }
  for (int l = 0; l < d.n_layer; ++l) {
    const LayerW& L = m->hw.layer[l];
    uint8_t* b = m->tc->blocks + (size_t)l * BLOCKS_PER_LAYER * WBLK_BYTES;
    int i = 0;
    // consumption order of the kernel: q, k, v, o, fc1 (two output halves), fc2 (two input halves), cheb1 x3, cheb2 x3
    for (int part = 0; part < 3; ++part) DP_TRY(pack_block(b + (size_t)(i++) * WBLK_BYTES, L.wqkv, 3 * H, 0, part * H, L.bqkv, s));
    DP_TRY(pack_block(b + (size_t)(i++) * WBLK_BYTES, L.wo, H, 0, 0, L.bo, s));
    for (int part = 0; part < 2; ++part) DP_TRY(pack_block(b + (size_t)(i++) * WBLK_BYTES, L.w1, 2 * H, 0, part * H, L.b1, s));
    for (int part = 0; part < 2; ++part) DP_TRY(pack_block(b + (size_t)(i++) * WBLK_BYTES, L.w2, H, part * H, 0, nullptr, s));
    for (int part = 0; part < 3; ++part) DP_TRY(pack_block(b + (size_t)(i++) * WBLK_BYTES, L.wc1, H, part * H, 0, part == 0 ? L.bc1 : nullptr, s));
    for (int part = 0; part < 3; ++part) DP_TRY(pack_block(b + (size_t)(i++) * WBLK_BYTES, L.wc2, H, part * H, 0, part == 0 ? L.bc2 : nullptr, s));
  }
  return DP_OK;
}

int tc_sample(dp_model* m, const float* x_in, int x_is_repeated, float* x_out, long n_pose, int n_hyp,
              const dp_step* steps_dev, const StepsArg* inl, int n_steps, const float* noise,
              const unsigned char* mask, cudaStream_t s) {
  if (!m->tc || !m->tc->blocks) { set_error("tensor-core engine: weights are not packed"); return DP_ERR_STATE; }
  static bool configured[64] = {};          // function attributes are per device
  bool& done = configured[m->device & 63];
  if (!done) {
    DP_CUDA(cudaFuncSetAttribute(tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    done = true;
  }
  TcArgs a{};
  a.w = m->dw; a.wpack = m->tc->blocks; a.n_layer = m->d.n_layer; a.x_in = x_in; a.x_is_repeated = x_is_repeated; a.out = x_out;
  a.n_rows = n_pose * n_hyp; a.n_pose = n_pose; a.n_steps = n_steps; a.temb = m->temb; a.noise = noise; a.mask = mask;
  a.steps_dev = steps_dev;
  const long n_tiles = (a.n_rows + TP - 1) / TP;
  const int grid = (int)(n_tiles < m->sm_count ? n_tiles : m->sm_count);
  tc_kernel<<<grid, kThreads, SMEM_BYTES, s>>>(a, *inl);
  count_launch();
  DP_CUDA(cudaGetLastError());
  m->last_launch[0] = grid; m->last_launch[1] = kThreads; m->last_launch[2] = SMEM_BYTES;
  m->last_launch[3] = TP; m->last_launch[4] = DP_ENGINE_TC; m->last_launch[5] = n_tiles;
  return DP_OK;
}

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
