/*
 * diffpose_b200.h -- C ABI of libdiffpose_b200.so: the B200 (sm_100a) implementation of DiffPose's
 * frame-based reverse-diffusion sampling path (GCNdiff denoiser + DDIM loop).
 *
 * The reference (nwicakson/diffpose-nw) is pure Python/PyTorch and has no FFI layer; its boundary for this
 * path is three Python call signatures.  Each entry point below names the reference interface it replaces
 * (paths relative to the reference repository root).  INTEGRATION.md shows the ctypes stub a reference
 * maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success, a negative dp_status otherwise; dp_last_error() returns a
 *     thread-local, human readable description of the last failure.  Nothing here calls abort().
 *   - all tensor pointers are DEVICE pointers to contiguous fp32 unless the name ends in _host.
 *   - the library borrows caller memory for the duration of a call; it owns only its packed weights and
 *     scratch, released by dp_destroy().
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Calls only enqueue work;
 *     they never synchronise the device.
 *   - one host thread per handle; one process per GPU.
 */
#ifndef DIFFPOSE_B200_H
#define DIFFPOSE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dp_model* dp_handle;

enum dp_status {
  DP_OK = 0,
  DP_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  DP_ERR_CUDA = -2,        /* a CUDA runtime call failed       */
  DP_ERR_STATE = -3,       /* e.g. dp_forward before dp_pack   */
  DP_ERR_UNSUPPORTED = -4  /* the requested engine cannot run this configuration */
};

/* Which kernel family executes the denoiser. */
enum dp_engine {
  DP_ENGINE_AUTO = 0,  /* DP_ENGINE_TCG when the configuration allows it, else fp32                 */
  DP_ENGINE_FP32 = 1,  /* fp32 FMA persistent kernel: every contraction in fp32 (bit-level reference) */
  DP_ENGINE_TC = 2,    /* first tcgen05 persistent kernel, kept for A/B measurements: fp16 operands (11-bit
                          significand, same as TF32), fp32 accumulation in TMEM, graph operators and attention
                          on the CUDA cores; hid_dim=96, n_head=4, n_pts=17 only                        */
  DP_ENGINE_TCG = 3    /* second-generation tcgen05 kernel (the default): every contraction on the tensor cores
                          (projections, 17x17 graph operators, attention, in/out convolutions), residual stream
                          in TMEM, dedicated MMA-issuer warp; same limits                                */
};

/* One DDIM step.  The scalars are evaluated by the caller with the reference's own fp32 tensor ops
 * (common/utils_diff.py:55-64) so that no re-derivation on the device can drift from them. */
typedef struct dp_step {
  float t;          /* timestep value fed to the denoiser (common/utils_diff.py:53,58)           */
  float sqrt_at;    /* sqrt(abar_t)                         (:59)                                 */
  float sqrt_1m_at; /* sqrt(1 - abar_t)                     (:59)                                 */
  float sqrt_an;    /* sqrt(abar_next)                      (:65)                                 */
  float c1;         /* eta * sqrt((1-at/an)(1-an)/(1-at))   (:61-63)                              */
  float c2;         /* sqrt((1-an) - c1^2)                  (:64)                                 */
} dp_step;

/* Replaces the constructors GCNdiff(adj, config) (models/gcndiff.py:56-98) and GCNpose(adj, config)
 * (models/gcnpose.py:56-98).  n_pts = config.model.n_pts, (c_in, c_out) = config.model.coords_dim,
 * hid = hid_dim, n_layer = num_layer, n_head = n_head; has_temb = 1 for GCNdiff, 0 for GCNpose. */
int dp_create(dp_handle* out, int n_pts, int c_in, int c_out, int hid, int n_layer, int n_head, int has_temb);

/* Number of fp32 values dp_pack expects (the state_dict flattened in the canonical order below). */
long dp_param_count(dp_handle h);

/* Replaces model.load_state_dict(states[0]) (runners/diffpose_frame.py:131-132) + the per-call graph
 * algebra of ChebConv.get_laplacian/cheb_polynomial (models/ChebConv.py:90-130) and
 * LAM_Gconv.laplacian_batch (models/GraFormer.py:174-178), which are evaluated once here.
 *
 * params: device pointer to the parameters, each tensor flattened in PyTorch (row-major) layout and
 * concatenated in this order (names are the reference state_dict keys):
 *   gconv_input.weight [3,1,c_in,hid], gconv_input.bias [hid]
 *   for l in 0..n_layer-1:
 *     gconv_layers.l.gconv1.gconv.weight [3,1,hid,hid], .bias [hid]
 *     gconv_layers.l.gconv2.gconv.weight [3,1,hid,hid], .bias [hid]
 *     (has_temb only) gconv_layers.l.temb_proj.weight [hid,4hid], .bias [hid]
 *     atten_layers.l.self_attn.linears.{0,1,2,3}.weight [hid,hid] each followed by its .bias [hid]
 *     atten_layers.l.feed_forward.A_hat [n_pts,n_pts]
 *     atten_layers.l.feed_forward.gconv1.fc.weight [2hid,hid], .bias [2hid]
 *     atten_layers.l.feed_forward.gconv2.fc.weight [hid,2hid], .bias [hid]
 *     atten_layers.l.sublayer.0.norm.a_2 [hid], .b_2 [hid], atten_layers.l.sublayer.1.norm.a_2, .b_2
 *   gconv_output.weight [3,1,hid,c_out], gconv_output.bias [c_out]
 *   (has_temb only) temb.dense.0.weight [4hid,hid], .bias [4hid], temb.dense.1.weight [4hid,4hid], .bias [4hid]
 * adj_host: host pointer, [n_pts*n_pts] fp32, the row-normalised adjacency (models/ChebConv.py:36-48). */
int dp_pack(dp_handle h, const float* params, long n_floats, const float* adj_host, void* stream);

/* Select the engine for subsequent dp_forward/dp_sample calls (default DP_ENGINE_AUTO). */
int dp_set_engine(dp_handle h, int engine);
/* Engine that a call with the current setting would use (DP_ENGINE_FP32, DP_ENGINE_TC or DP_ENGINE_TCG). */
int dp_get_engine(dp_handle h);

/* Replaces GCNdiff.forward(x, mask, t, cemd) (models/gcndiff.py:101-113; call sites
 * common/utils_diff.py:58, runners/diffpose_frame.py:225) and GCNpose.forward(x, mask)
 * (models/gcnpose.py:101-113; call site runners/diffpose_frame.py:337).
 *   x [n,n_pts,c_in], t [n] (per-sample timesteps; ignored/NULL when has_temb = 0),
 *   mask: [n_pts] bytes on the device (non-zero = key visible) or NULL (= all visible),
 *   out [n,n_pts,c_out]. */
int dp_forward(dp_handle h, const float* x, const float* t, const unsigned char* mask, float* out, long n, void* stream);

/* Replaces generalized_steps(x, src_mask, seq, model, b, eta=...) (common/utils_diff.py:46-67) plus the
 * hypothesis handling around it (runners/diffpose_frame.py:342 `.repeat(test_times,1,1)` and :382
 * `mean(reshape(test_times,-1,17,5),0)`), in ONE persistent launch for all T steps.
 *   x_in      [n_rows_in,n_pts,c] with n_rows_in = n_pose*n_hyp if x_is_repeated else n_pose
 *             (the library reads pose b for every hypothesis h when x_is_repeated = 0);
 *   x_out     [n_pose*n_hyp,n_pts,c] hypothesis-major (index h*n_pose+b) when mean_over_hyp = 0,
 *             [n_pose,n_pts,c] when mean_over_hyp = 1;
 *   steps_host T entries in execution order (largest t first);
 *   noise     [T,n_pose*n_hyp,n_pts,c] host-drawn N(0,1) replacing randn_like (:65), or NULL (term skipped;
 *             exact when every c1 = 0, i.e. eta = 0);
 *   mask      as in dp_forward. */
int dp_sample(dp_handle h, const float* x_in, int x_is_repeated, float* x_out, long n_pose, int n_hyp,
              const dp_step* steps_host, int n_steps, const float* noise, const unsigned char* mask,
              int mean_over_hyp, void* stream);

/* Replaces mpjpe (common/loss.py:7-13) and p_mpjpe (common/loss.py:25-64, common/utils.py:155-187) as used
 * at runners/diffpose_frame.py:382-387: root-centres both inputs out of place, then accumulates
 * sums[0] += sum_pose mean_joint |pred-gt|, sums[1] += sum_pose P-MPJPE(pose), sums[2] += n.
 *   pred [n,n_pts,pred_stride] with xyz at column pred_offset (2 for uvxyz, 0 for xyz), gt [n,n_pts,3],
 *   sums: 3 doubles on the device (caller zeroes them; per-rank partials feed one all-reduce),
 *   per_pose: optional [n,2] fp32 device output (mpjpe, p_mpjpe per pose) or NULL. */
int dp_metrics(const float* pred, int pred_stride, int pred_offset, const float* gt, long n, int n_pts,
               double* sums, float* per_pose, void* stream);

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

/* Number of kernels this library has launched in this process (bench.py reports it as gpu_launches). */
long dp_launch_count(void);

/* Description of the kernel the last dp_sample/dp_forward launched: grid, block, dynamic smem bytes,
 * poses per tile, engine.  out6 is a host array of 6 longs. */
int dp_last_launch_info(dp_handle h, long* out6);

const char* dp_last_error(void);
const char* dp_version(void);
void dp_destroy(dp_handle h);

#ifdef __cplusplus
}
#endif
#endif /* DIFFPOSE_B200_H */
