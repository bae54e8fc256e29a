/*
 * node_b200.h - C ABI of the B200-native dopri5 / ODE-Net hot path.
 *
 * Drop-in boundary for fabiocarrara/neural-ode-features: the reference binds this path by
 * Python import (`from torchdiffeq import odeint_adjoint, odeint`, reference model.py:3); the
 * package neural-ode-features_b200/torchdiffeq re-exports those two names and reaches the
 * kernels below through ctypes.  Plain pointers and sizes only - no torch types.
 *
 * All pointers are DEVICE pointers unless named host_*.  `stream` is a cudaStream_t passed
 * as void* (torch.cuda.current_stream().cuda_stream).  Every entry point only ENQUEUES work
 * and returns a cudaError_t-compatible int (0 = success); none of them synchronises.
 *
 * File:line citations are relative to /root/reference/torchdiffeq/torchdiffeq/_impl/ unless
 * they start with model.py.
 */
#ifndef NODE_B200_H_
#define NODE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NODE_B200_ABI_VERSION 1
#define NODE_MAX_SEG 8      /* members of a tuple state (adjoint uses 4: y, adj_y, adj_t, adj_p) */
#define NODE_MAX_TRACE 192  /* attempted steps recorded for parity tests */

/* dtype tags for the generic-callable route */
#define NODE_F32 0
#define NODE_F64 1

/* status bits, raised as AssertionError by the Python boundary after the solve */
#define NODE_ST_DT_UNDERFLOW 1  /* dopri5.py:100  t0 + dt > t0 */
#define NODE_ST_NONFINITE    2  /* dopri5.py:102  non-finite values in state */
#define NODE_ST_MAX_STEPS    4  /* dopri5.py:89   max_num_steps exceeded */
#define NODE_ST_INTERP_RANGE 8  /* interp.py:58   t0 <= t <= t1 */
#define NODE_ST_WATCHDOG     16 /* an in-kernel barrier wait timed out (never expected) */

/*
 * Device-resident controller block: the whole adaptive-step state machine of
 * dopri5.py:77-122 + misc.py:84-170 lives here so that no host round trip is needed per
 * step.  The host reads it back once per solve (fused route) or once per attempted step
 * (generic-callable route, where the callable itself is host code).
 */
typedef struct node_ctl {
  /* float64 time axis (solvers.py:28, dopri5.py:72-75) */
  double t0, t1;          /* last ACCEPTED interval; outputs are interpolated inside it */
  double dt;              /* size of the next attempt */
  double t_attempt;       /* start of the attempt being / about to be evaluated (== t1) */
  double dt_attempt;      /* its size */
  double h0;              /* misc.py:128-131 first guess, kept in the state dtype */
  double d1max;           /* misc.py:126 max over tuple members, kept for INIT_B */
  double ratio[NODE_MAX_SEG]; /* misc.py:155-156 mean squared error ratio of the last attempt */
  double rtol[NODE_MAX_SEG], atol[NODE_MAX_SEG];
  double safety, ifactor, dfactor; /* as built by dopri5.py:72-74: rounded through torch's default dtype */
  double expo;            /* misc.py:168 exponent 1/order, same rounding */
  /* the same times rounded to the state dtype (rk_common.py:45-46, interp.py:54-56) */
  double ts64[7];         /* [0] attempt start s, [1..6] s + alpha_i*h, state dtype f64 */
  double h64;
  float  ts32[7];
  float  h32;
  float  h0_32, pad0;
  /* dense output bookkeeping (dopri5.py:85-92) */
  double it_t0, it_t1;    /* interval the interpolant belongs to, float64 */
  double it_h64;  float it_h32; int32_t it_cur;  /* its dt in state dtype; buffer that holds y0,f0 */
  int32_t out_lo, out_hi; /* output indices [lo,hi) to emit from the step just accepted */
  int32_t next_out, n_out;
  /* counters */
  int32_t n_attempt, n_accept, n_reject, nfe;
  int32_t steps_this_advance, max_num_steps;
  int32_t status, done, cur, accepted_last;
  int32_t n_seg, dtype, tsign, reserved;
  int64_t seg_numel[NODE_MAX_SEG]; /* GLOBAL element count per member (all shards) */
  /* trace of attempted steps */
  double tr_t[NODE_MAX_TRACE], tr_dt[NODE_MAX_TRACE], tr_ratio[NODE_MAX_TRACE];
  int32_t tr_acc[NODE_MAX_TRACE];
} node_ctl_t;

int node_b200_abi_version(void);
/* sizeof(node_ctl_t) and byte offsets of the fields the Python side reads, in the order
 * documented in node_b200/native.py (CTL_FIELDS). Host function. */
int node_b200_ctl_layout(int64_t* host_out, int capacity);

/* ---- generic-callable route: building blocks (func is evaluated by the caller) ---------- */

/* Zero-initialise the block and set tolerances / options (dopri5.py:60-75).
 * host arrays rtol/atol have n_seg entries; seg_numel are GLOBAL counts. safety/ifactor/dfactor/expo
 * are passed as the reference builds them: `torch.tensor(x)` in the DEFAULT dtype widened to float64
 * (misc.py:37-44), i.e. (double)(float)0.9 etc. when the default dtype is float32. */
int node_b200_ctl_init(node_ctl_t* ctl, int dtype, int n_seg, const double* host_rtol, const double* host_atol,
                       const int64_t* host_seg_numel, double safety, double ifactor, double dfactor, double expo,
                       int max_num_steps, int n_out, int tsign, void* stream);

/* K2 - stage combination  out = y0 + sum_j (h*c_j)*k_j  (rk_common.py:49-51, misc.py:22-25).
 * `which` selects the coefficient row: 0..5 = beta row of stage i+1, 6 = C_MID (dopri5.py:33-42),
 * 7 = initial-step probe y0 + h0*f0 (misc.py:133). h is read from ctl: the current attempt's for rows 0..5,
 * the last ACCEPTED step's for row 6 (call it after the controller), h0 for row 7. n_k pointers in ks. Row 6 only feeds the
 * dense output: it returns at once unless the controller scheduled outputs for the accepted step (ctl.out_hi > ctl.out_lo). */
int node_b200_rk_stage_combine(const node_ctl_t* ctl, int dtype, int which, void* out, const void* y0,
                               const void* const* host_ks, int n_k, int64_t numel, void* stream);

/* K3 - error estimate + per-member sum of squared ratios (rk_common.py:60, misc.py:146-157);
 * never materialises y1_error. seg_off/seg_len (host arrays, elements) describe the members of
 * the flat state. Writes partials[seg][block] (double) and sets a non-finite flag for |y0|.
 * host_ks = k1..k7. */
int node_b200_rk_error_norm(const node_ctl_t* ctl, int dtype, const void* y0, const void* y1,
                            const void* const* host_ks, const int64_t* host_seg_off, const int64_t* host_seg_len,
                            int n_seg, double* partials, int* nonfinite_flag, void* stream);

/* K5 - initial step norms (misc.py:123-126,136): mode 0 writes sum((y0/scale)^2), sum((f0/scale)^2);
 * mode 1 writes sum(((f1-f0)/scale)^2). partials layout [seg][2][block]. */
int node_b200_init_norms(const node_ctl_t* ctl, int dtype, int mode, const void* y0, const void* f0, const void* f1,
                         const int64_t* host_seg_off, const int64_t* host_seg_len, int n_seg,
                         double* partials, void* stream);

/* Fold per-block partials into sums[seg][2] in a fixed order (deterministic); this is the
 * buffer a multi-GPU run all-reduces before the controller consumes it. */
int node_b200_reduce_partials(const double* partials, int n_rows, double* sums, void* stream);

/* K6 - device-side controller. mode 0: INIT_A (misc.py:123-131), 1: INIT_B (misc.py:136-143,
 * dopri5.py:80-83), 2: STEP (dopri5.py:109-121, misc.py:160-170, dopri5.py:88-92 output
 * scheduling), 3: FIXED FIRST STEP (dopri5.py:81-82, options['first_step'] given: dt = sums[0], no
 * probe evaluation). t_out: the requested times as float64 [n_out] (already sign-flipped if the
 * caller integrates backwards, misc.py:184-187). */
int node_b200_controller(node_ctl_t* ctl, int mode, const double* sums, const int* nonfinite_flag,
                         const double* t_out, void* stream);

/* K4 - dense output (dopri5.py:39-45, interp.py:5-65) for output indices [ctl.out_lo, ctl.out_hi):
 * out[(idx*out_stride) + e] for e < numel. No-op when the range is empty. use_ctl_cur != 0:
 * (y0,y1) and (f0,f1) are ping-pong buffers and ctl.it_cur says which one held the start of
 * the accepted step (fused route); 0: they are taken as named. */
int node_b200_interp_eval(const node_ctl_t* ctl, int dtype, const double* t_out, void* out, int64_t out_stride,
                          const void* y0, const void* y1, const void* ymid, const void* f0, const void* f1,
                          int64_t numel, int use_ctl_cur, void* stream);

/* ---- fused route: recognised ODEfunc (model.py:326-348), fp32 NCHW ---------------------- */

/* Bytes of workspace the fused solver needs for a [N,C,H,W] shard. Host function. */
int64_t node_b200_fused_workspace_bytes(int N, int C, int H, int W);

/* Pre-arrange the live parameters for the kernels: fp16 (and, for the first engine, TF32) hi/lo split weight tiles in the
 * UMMA K-major 128B-swizzled layout, the position-dependent time map Tmap[c,h,w]
 *  and packed GroupNorm affine terms. Pointers are the reference's own
 * nn.Parameters (model.py:329-334): conv weights [C, C+1, 3, 3] with the time plane at
 * input channel 0, biases [C], GroupNorm weight/bias [C]. */
int node_b200_fused_prepare(void* workspace, int C, int H, int W,
                            const float* conv1_w, const float* conv1_b, const float* conv2_w, const float* conv2_b,
                            const float* gn1_w, const float* gn1_b, const float* gn2_w, const float* gn2_b,
                            const float* gn3_w, const float* gn3_b, float eps, void* stream);

/* One evaluation k = tsign * ODEfunc(tsign * t, y) on a [N,C,H,W] tensor (model.py:339-348);
 * conv_mode 0 = tcgen05 fp16 operand split (fp32 contract, the step engine), 1 = single fp16 product, 2 = SIMT fp32 FFMA,
 * 3 = tcgen05 3xTF32 (fp32 contract, first engine), 4 = single TF32 product. */
int node_b200_odefunc_forward(void* workspace, const float* y, float t, float tsign, float* k,
                              int N, int C, int H, int W, int conv_mode, void* stream);

/* Whole forward solve of odeint(func, y0, t, rtol, atol, method='dopri5') for the recognised
 * ODEfunc (odeint.py:20-76, solvers.py:25-33, dopri5.py:60-122): enqueues f0, the initial-step
 * probe, and `n_steps_enqueue` attempted steps (each: fused 6-stage step kernel, controller,
 * dense output); steps after ctl.done are no-ops. Call again with first_call=0 to enqueue more
 * attempts if the host finds ctl.done == 0. out: [T, N, C, H, W]; out[0] = y0 (solvers.py:27).
 * global_numel = N_global*C*H*W (differs from the shard's when the batch is sharded).
 * t_out is a HOST array of T float64 times, strictly increasing; tsign = -1 integrates the
 * time-reversed system f'(t, y) = -f(-t, y) (misc.py:184-187; the caller negates t). */
int node_b200_fused_solve(void* workspace, const float* y0, const double* t_out, int T, double rtol, double atol,
                          int N, int C, int H, int W, int64_t global_numel, float* out, int conv_mode,
                          int tsign, int first_call, int n_steps_enqueue, void* stream);

/* Multi-GPU stepping: phase 0 = f0 + INIT_A norms, 1 = INIT_A controller + probe + INIT_B norms,
 * 2 = INIT_B controller, 3 = step kernel + norms, 4 = STEP controller + dense output. Between
 * phases {0,1,3} and the next one the caller all-reduces node_b200_fused_sums(workspace). */
int node_b200_fused_phase(void* workspace, int phase, const float* y0, const double* t_out, int T, double rtol,
                          double atol, int N, int C, int H, int W, int64_t global_numel, float* out,
                          int conv_mode, int tsign, void* stream);
double* node_b200_fused_sums(void* workspace);       /* device pointer, 2*NODE_MAX_SEG doubles */
node_ctl_t* node_b200_fused_ctl(void* workspace);    /* device pointer to the controller block */

/* ---- adjoint: native vector-Jacobian product of the recognised ODEfunc (adjoint.py:32-55) ---- */

/* Bytes of scratch the VJP needs for a [N,C,H,W] shard (convolution inputs and output gradients
 * for the weight-gradient GEMM, per-CTA partials). Host function; < 0 when the shape is not served. */
int64_t node_b200_vjp_workspace_bytes(int N, int C, int H, int W);

/* One evaluation of the adjoint's augmented dynamics (adjoint.py:32-55) for the recognised ODEfunc,
 * replacing `func(t, y)` + `torch.autograd.grad(f, (t, y) + params, -adj_y)`:
 *   f_out      = s * ODEfunc(s*t, y)                      [N,C,H,W]
 *   vjp_y      = s * d<f, -adj_y>/dy                      [N,C,H,W]
 *   vjp_t      = s * d<f, -adj_y>/dt                      1 float
 *   vjp_params = s * d<f, -adj_y>/dparams                 75,392 floats in func.parameters() order
 *                (norm1.w, norm1.b, conv1.W[C,C+1,3,3], conv1.b, norm2.*, conv2.*, norm3.*; misc.py:5-7)
 * s = tsign (+1, or -1 for the reversed-time wrapper of misc.py:184-187); t_dev is a DEVICE scalar (the
 * solver's stage time lives in the controller block). `workspace` is the fused workspace whose parameters
 * were prepared with node_b200_fused_prepare. Enqueues three kernels: fused VJP (tcgen05 forward + data
 * gradients), weight-gradient GEMM (tcgen05), deterministic fold. */
int node_b200_odefunc_vjp(void* workspace, void* vjp_workspace, const float* y, const float* adj_y,
                          const float* t_dev, float tsign, float* f_out, float* vjp_y, float* vjp_t,
                          float* vjp_params, int N, int C, int H, int W, void* stream);

/* The weight-gradient GEMM alone: per-CTA partials of dW[co, ci, tap] = sum GC[n,co,h,w] * IN[n,ci,h+dy,w+dx]
 * for both convolutions (r = convolution inputs, gc = gradients at the convolution outputs, [N,C,H,W]). */
int node_b200_wgrad(void* vjp_workspace, const float* r1, const float* gc1, const float* r2, const float* gc2,
                    int N, int C, int H, int W, void* stream);

/* Test aid: device pointers inside the VJP workspace (0..3 = R1, R2, GC1, GC2; 4 = weight-gradient partials
 * [splits][2][9][64][80]). */
void* node_b200_vjp_buffer(void* vjp_workspace, int which, int N, int C, int H, int W);

/* Callers of the hot path: y = relu?(GroupNorm(x)) for a contiguous NCHW fp32 tensor [N, C, HW] with
 * `groups` groups - replaces nn.GroupNorm (model.py:268-271) followed by nn.ReLU in the downsamplers
 * (model.py:119-178) and the classifier head (model.py:231-250) by one pass (1 read + 1 write). Two-pass mean /
 * biased variance like native_group_norm. Returns cudaErrorInvalidValue for cells larger than 4096 floats
 * (the caller then keeps its own GroupNorm). x == y (in place) is allowed. */
int node_b200_groupnorm_relu(const float* x, float* y, const float* gamma, const float* beta, int64_t N, int C,
                             int groups, int HW, float eps, int relu, void* stream);

/* Callers of the hot path: the tail of the reference's ResBlock (model.py:156-178),
 *     out = conv2(relu(norm2(x))) + shortcut,  conv2 = Conv2d(64, 64, 3, 1, 1, bias=False), norm2 = GroupNorm(32, 64),
 * as one tcgen05 kernel (fp32 contract by fp16 operand splitting, like the ODE-Net step engine). x, shortcut, out are
 * contiguous NCHW fp32 [N,64,H,W]; supported maps: 15x15, 8x8 (CIFAR), 13x13, 7x7 (MNIST) - workspace_bytes returns 0
 * otherwise and the caller keeps its own ops. prepare() packs the live conv weight [64,64,3,3] and the operand scales
 * into the workspace (once per parameter version). With next_gn_w / next_gn_b (the affine parameters of the FOLLOWING
 * block's norm1, same eps; NULL to disable) the kernel writes relu(GroupNorm(out)) instead of out - the following block
 * consumes nothing else when it has a projection shortcut (model.py:167-172). */
int64_t node_b200_resconv_workspace_bytes(int C, int H, int W);
int node_b200_resconv_prepare(void* workspace, int C, int H, int W, const float* conv_w, const float* gn_w, const float* gn_b,
                              void* stream);
int node_b200_resconv_forward(void* workspace, const float* x, const float* shortcut, float* out, const float* gn_w,
                              const float* gn_b, const float* next_gn_w, const float* next_gn_b, int N, int C, int H, int W,
                              float eps, void* stream);

/* Callers of the hot path: the head of the reference's strided ResBlock (model.py:156-178) on the already
 * normalised activation a = relu(norm1(x)), contiguous NCHW fp32 [N,64,HI,WI]:
 *     c_out  = conv1(a)       conv1      = Conv2d(64, 64, 3, stride 2, padding 1, bias=False)   [N,64,HO,WO]
 *     sc_out = downsample(a)  downsample = Conv2d(64, 64, 1, stride 2, bias=False)              [N,64,HO,WO]
 * in one tcgen05 kernel (parity planes: every stride-2 tap is a stride-1 tap of one plane). Supported inputs: 30x30,
 * 15x15 (CIFAR), 26x26, 13x13 (MNIST); workspace_bytes returns 0 otherwise. prepare() takes the two live weights and the
 * affine parameters of norm1 (they bound the activation for the fp16 operand split). */
int64_t node_b200_convs2_workspace_bytes(int C, int HI, int WI);
int node_b200_convs2_prepare(void* workspace, int C, int HI, int WI, const float* conv_w, const float* down_w,
                             const float* gn_w, const float* gn_b, void* stream);
int node_b200_convs2_forward(void* workspace, const float* act, float* c_out, float* sc_out, int N, int C, int HI, int WI,
                             void* stream);

/* Callers of the hot path: the stem of the reference's downsamplers (model.py:119-178),
 *     out = relu(GroupNorm(32, 64)(Conv2d(CIN, 64, 3, 1)(x)))          x [N,CIN,HIN,WIN] -> out [N,64,HIN-2,WIN-2]
 * in one pass (one CTA per image, one thread per output pixel, fp32 FFMA, two-pass GroupNorm statistics). Served:
 * (CIN,HIN,WIN) = (3,32,32) CIFAR and (1,28,28) MNIST; cudaErrorInvalidValue otherwise (the caller keeps its own ops). */
int node_b200_stem_gn_relu(const float* x, const float* conv_w, const float* conv_b, const float* gn_w, const float* gn_b,
                           float* out, int N, int CIN, int HIN, int WIN, float eps, void* stream);

/* Callers of the hot path: the reference's FCClassifier (model.py:231-250) - GroupNorm(32, 64) -> ReLU ->
 * global average pool -> Linear(64, n_out) on a contiguous NCHW fp32 [N,64,HW] tensor in one pass; lin_w == NULL stops
 * after the pool (feature mode, model.py:39-40) and writes [N,64]. No dropout (inference). */
int node_b200_head(const float* x, const float* gn_w, const float* gn_b, const float* lin_w, const float* lin_b, float* out,
                   int64_t N, int C, int HW, int n_out, float eps, void* stream);

/* SURVEY 8f-4 - the edges after the feature extractor, on device-resident buffers.
 * feature_normalize: the reference's retrieval normalisation (evaluate.py:326)
 *     out = features / (np.linalg.norm(features, axis=-2, keepdims=True) + 1e-7)
 * for features [planes, N, D] fp32 (planes = tol x T of the HDF5 layout features[tol, T, N, D], evaluate.py:88-94): the norm
 * runs over the N SAMPLES of a plane, one value per feature dimension (norms [planes, D], also returned). out may alias
 * features.
 * retrieval_scores: scores = queries . db^T (evaluate.py:339), queries [nq, D], db [ns, D], scores [nq, ns], fp32 FFMA. */
int node_b200_feature_normalize(const float* features, float* out, float* norms, int64_t planes, int64_t N, int D, void* stream);
int node_b200_retrieval_scores(const float* queries, const float* db, float* scores, int64_t nq, int64_t ns, int D, void* stream);

/* SURVEY 8f-3, backward of the callers (model.py:119-178 and 231-250 under autograd; train.py:40-58). All tensors contiguous
 * NCHW fp32 on the device.
 * groupnorm_relu_backward: y = relu?(GroupNorm(groups, C)(x)) with C = 2 * groups; given grad_out = dL/dy writes grad_in = dL/dx
 *   (2 reads + 1 write), grad_gamma[C], grad_beta[C]; partials = scratch of 2 * N * C floats (per-image sums, folded in a fixed
 *   order). Statistics are recomputed (two-pass) - nothing is saved by the forward.
 * absmax: *out_bits = bit pattern of max |v| (the power-of-two operand scale of a gradient tensor in conv_wgrad).
 * plane_split / plane_merge: planes[pr][pc][n][c][i][j] = a[n][c][2i+pr][2j+pc] (zero beyond the map), planes
 *   [4][N][C][HO][WO] with HO = (HI-1)/2+1 - a stride-2 3x3 convolution is the sum of four stride-1 convolutions on them - and
 *   the inverse scatter for the gradients.
 * conv3x3_prepare / conv3x3_forward: out = conv2d(x, W, stride 1, padding 1) (+ addend) for a SIGNED x [N,64,H,W] and an
 *   ordinary weight [64,64,3,3] on the tcgen05 engine of resconv_forward (fp16 operand split, the operand scale found per
 *   super-tile in the kernel): the data gradient of a 3x3 convolution when W is its flipped, transposed kernel. Workspace of
 *   resconv_workspace_bytes; maps 15x15, 8x8, 13x13, 7x7.
 * conv_wgrad: dw[p][co][ci][tap] = sum_{n,pos} grads[p][n,co,pos] * inputs[p][n,ci,pos+tap] for npairs <= 6 (input, gradient)
 *   pairs in one launch of the adjoint's weight-gradient GEMM; input_scales[p] / grad_max_bits[p] are DEVICE scalars (the
 *   power-of-two scale of the input operand; the bit pattern of max |grad|, see absmax). Maps 15x15, 8x8 (CIFAR), 13x13, 7x7 (MNIST).
 * stem_backward: gradients of out = relu(GroupNorm(32,64)(Conv2d(CIN,64,3,1)(x))) with respect to conv weight [64*CIN*9],
 *   conv bias [64], gamma [64], beta [64] - written in that order to grads - without materialising the conv output
 *   (recomputed per image). x itself receives no gradient. */
int node_b200_groupnorm_relu_backward(const float* x, const float* grad_out, float* grad_in, const float* gamma, const float* beta,
                                      float* partials, float* grad_gamma, float* grad_beta, int64_t N, int C, int groups, int HW,
                                      float eps, int relu, void* stream);
int node_b200_absmax(const float* v, int64_t n, unsigned* out_bits, void* stream);
int node_b200_plane_split(const float* a, float* planes, int64_t N, int C, int HI, int WI, void* stream);
int node_b200_plane_merge(const float* gplanes, float* ga, int64_t N, int C, int HI, int WI, void* stream);
int node_b200_conv3x3_prepare(void* workspace, int C, int H, int W, const float* conv_w, void* stream);
int node_b200_conv3x3_forward(void* workspace, const float* x, const float* addend, float* out, int N, int C, int H, int W,
                              void* stream);
int64_t node_b200_conv_wgrad_workspace_bytes(int npairs);
int node_b200_conv_wgrad(void* workspace, int npairs, const float* const* inputs, const float* const* grads,
                         const float* const* input_scales, const unsigned* const* grad_max_bits, float* dw, int N, int C, int H,
                         int W, void* stream);
int64_t node_b200_stem_backward_workspace_bytes(int CIN);
int node_b200_stem_backward(const float* x, const float* conv_w, const float* conv_b, const float* gn_w, const float* gn_b,
                            const float* grad_out, void* workspace, float* grads, int N, int CIN, int HIN, int WIN, float eps,
                            void* stream);
/* byte offset of the operand-scale block (float[8]; [0] = activation scale) inside a resconv / convs2 workspace */
int64_t node_b200_resconv_scal_offset(void);
int64_t node_b200_convs2_scal_offset(void);

/* all-reduce of the solver's float64 sums over NVLink peer memory, fused into the kernel that folds the per-CTA
 * partials (csrc/peer_reduce.cu) - no host-launched collective between an attempted step and its controller. One process per
 * GPU: peer_alloc returns the 64-byte CUDA IPC handle of this process' exchange buffer; the host side gathers the handles of
 * all ranks (torch.distributed) and passes them, rank-major, to peer_open; from then on every node_b200_fused_phase /
 * node_b200_fused_solve call reduces its sums over the peers (all ranks must issue the same call sequence). peer_close returns
 * to single-GPU behaviour. fold_reduce is the building block itself: sums[row] = sum_b partials[row][b], then summed over ranks. */
int node_b200_peer_alloc(void* handle_out_64_bytes);
int node_b200_peer_open(int world, int rank, const void* handles);
int node_b200_peer_close(void);
int node_b200_peer_world(void);
int node_b200_fold_reduce(const double* partials, int nblocks, double* sums, int nrows, void* stream);

/* One attempted step of odeint_adjoint's backward integration (adjoint.py:23-102 via dopri5.py:94-122) for the fused ODE-Net
 * dynamics, enqueued by ONE call: 6 x (stage combination of the 4-member augmented state, augmented dynamics through the native
 * VJP kernels), error norm, fold, controller. `bufs` are the generic route's 11 state rows of row_elems floats (Y0 Y1 F0 F1
 * K2..K6 YMID YI), members (y, adj_y, adj_t, adj_params) at seg_off; `cur` = index of the current y/f pair; ts32_offset_bytes =
 * offset of the controller's fp32 stage times inside *ctl (node_b200_ctl_layout). The caller reads the controller block back
 * once per attempt, as before. */
int node_b200_adjoint_step(void* ctl, float* bufs, int64_t row_elems, int cur, const int64_t* host_seg_off,
                           const int64_t* host_seg_len, int n_seg, void* workspace, void* vjp_workspace, float tsign,
                           int64_t ts32_offset_bytes, int N, int C, int H, int W, double* partials, double* sums,
                           int* nonfinite_flag, const double* t_out, void* stream);

/* One INTERVAL of odeint_adjoint's backward integration (adjoint.py:77-97: odeint of the augmented system over [t_i, t_{i-1}];
 * SURVEY 8b `node_dopri5_adjoint_backward`) as ONE call without a host read: f0 (dopri5.py:78), the initial-step probe and
 * norms (misc.py:84-143) - or the fixed first step when first_step_given != 0, sums[0] holding the constant - and then the loop
 * `while next_t > t1: step` (dopri5.py:88) as a CUDA graph WHILE node whose condition the device-resident controller sets
 * (csrc/adjoint_solve.cu). Arguments as node_b200_adjoint_step; the start state is in row 0 of `bufs`, ctl is initialised by
 * node_b200_ctl_init, t_out = device float64 [n_out] (sign-flipped for reversed spans), out = [n_out][row_elems] floats: out[0] =
 * the start state, out[j] = the dense output at t_out[j]. The attempt always runs rows (Y0, F0) -> (Y1, F1) and a commit kernel
 * copies the accepted pair back, so ctl.cur is not used. The loop graphs are cached per argument set (the caller keeps the
 * buffers alive and at the same addresses to reuse them); adjoint_solve_reset drops them. The caller reads ctl back ONCE,
 * after the call, for status / counters / trace. */
int node_b200_adjoint_solve(void* ctl, float* bufs, int64_t row_elems, const int64_t* host_seg_off, const int64_t* host_seg_len,
                            int n_seg, void* workspace, void* vjp_workspace, void* vjp_workspace2, float tsign,
                            int64_t ts32_offset_bytes, int N, int C, int H, int W, double* partials, double* sums, int* nonfinite_flag,
                            const double* t_out, float* out, int first_step_given, void* stream);
/* vjp_workspace2 (a second node_b200_vjp_workspace_bytes buffer, or null): when given, the weight-gradient GEMM + fold of stage i
 * run on a side branch of the loop graph under stage i + 1's k_vjp (small batches: the interval is a dependent chain).
 * odefunc_vjp_split: node_b200_odefunc_vjp with that second part on side_stream, ordered after the first by fork_event. */
int node_b200_odefunc_vjp_split(void* workspace, void* vjp_workspace, const float* y, const float* adj_y, const float* t_dev,
                                float tsign, float* f_out, float* vjp_y, float* vjp_t, float* vjp_params, int N, int C, int H, int W,
                                void* stream, void* side_stream, void* fork_event);
int node_b200_adjoint_solve_reset(void);

/* Linear combinations with device-resident coefficients - the autograd nodes of the unrolled route (node_b200/unrolled.py; the
 * reference records `y + sum((h*c_j)*k_j)` (misc.py:22-25), the Hermite fit (interp.py:5-35) and the interpolant (interp.py:54-65)
 * as one ATen multiply + one ATen add per term). dtype NODE_F32 / NODE_F64, n <= 7 sources of numel elements, coef = n values of
 * that dtype on the device. Same rounding as the reference's op sequence (every product and addition rounds separately, left to
 * right from 0, then base + sum).
 * lincomb: out = base + sum_j coef[j] * src_j (base may be null); lincomb_scale: dst_j = coef[j] * g (null dst skipped);
 * lincomb_dots: dots[j] = sum_e src_j[e] * g[e] (float64 accumulation, fixed order; partial = lincomb_scratch_doubles() doubles). */
int node_b200_lincomb(int dtype, void* out, const void* base, const void* const* host_srcs, const void* coef, int n, int64_t numel,
                      void* stream);
int node_b200_lincomb_scale(int dtype, void* const* host_dsts, const void* g, const void* coef, int n, int64_t numel, void* stream);
int node_b200_lincomb_dots(int dtype, const void* const* host_srcs, const void* g, int n, int64_t numel, double* partial, void* dots,
                           void* stream);
int64_t node_b200_lincomb_scratch_doubles(void);

/* Wide dynamics (n_filters = 128, 192, 256: the paper's CIFAR setting, reproduce.sh:21): ODEfunc.forward (model.py:339-348) as
 * 64-channel blocks on the tcgen05 engine instead of cuDNN.
 * conv3x3_forward_strided = conv3x3_forward on a 64-channel block of a wider tensor: x / (addend, out) point at the block's
 *   first channel and in_image_stride / out_image_stride are the element distances between images (C_total * H * W);
 *   addend == out accumulates the input-channel blocks of one output block.
 * groupnorm_relu_ex: y = post * relu?(GroupNorm(x + add_bias[c] + tsign * t * add_tmap[c][pix])) - the GroupNorm after a
 *   ConcatConv2d whose time channel is folded into b + t * Tmap (t: device scalar). L = (C / groups) * HW must be a multiple
 *   of 4 and <= 4096. */
int node_b200_conv3x3_forward_strided(void* workspace, const float* x, const float* addend, float* out, int N, int C, int H, int W,
                                      int64_t in_image_stride, int64_t out_image_stride, void* stream);
int node_b200_groupnorm_relu_ex(const float* x, float* y, const float* gamma, const float* beta, const float* add_bias,
                                const float* add_tmap, const float* t_dev, float tsign, float post, int64_t N, int C, int groups,
                                int HW, float eps, int relu, void* stream);

/* Backward pieces of the WIDE dynamics' augmented system (adjoint.py:32-55 at n_filters = 128 / 192 / 256; csrc/wide_vjp.cu):
 * groupnorm_backward_ex: (grad_in, grad_gamma[C], grad_beta[C]) of y = relu?(GroupNorm(x + add_bias[c] + tsign * t * add_tmap[c][pix]))
 *   for any C / groups with (C / groups) * HW <= 4096; partials = 2 * N * C floats of scratch (per-image dgamma / dbeta, folded in
 *   a fixed order). The ReLU mask is recomputed from x.
 * batch_colsum: out[j] = sum_n v[n][j] (float64 accumulation, fixed order) - the bias / time-channel gradients' sum over the batch.
 * pow2_scale: scale = 2^floor(log2(16384 / max)) from the bit pattern of max |v| (node_b200_absmax): the operand scale
 *   node_b200_conv_wgrad expects for a non-negative activation. */
int node_b200_groupnorm_backward_ex(const float* x, const float* grad_out, float* grad_in, const float* gamma, const float* beta,
                                    const float* add_bias, const float* add_tmap, const float* t_dev, float tsign, float* partials,
                                    float* grad_gamma, float* grad_beta, int64_t N, int C, int groups, int HW, float eps, int relu,
                                    void* stream);
int node_b200_batch_colsum(const float* v, float* out, int64_t N, int64_t cols, void* stream);
int node_b200_pow2_scale(const unsigned* max_bits, float* scale, void* stream);

/* wide_vjp: ONE evaluation of the adjoint's augmented dynamics (adjoint.py:32-55) of a wide ODEfunc by one call - forward keeping the
 * activations, GroupNorm backward x 3, data gradients (wide8 implicit GEMM on a raw operand image at 8x8 / C = 128, 256; block
 * convolutions otherwise), weight gradients (64-channel block GEMMs), bias / time-channel gradients, vjp_t. p = host array of 38
 * device pointers, dims = host array {N, C, H, W, block workspace stride, use_wide8}: indices documented in csrc/wide_vjp.cu.
 * Outputs are tsign * (f, vjp_y, vjp_t, vjp_params)(tsign * t) with cotangent -adj_y (misc.py:184-187 for reversed spans).
 * wide8_raw_operand: operand image of a SIGNED fp32 tensor at a scale found from max|x| (max_bits from node_b200_absmax); the
 * following node_b200_wide8_conv(workspace, which, ...) uses it. */
int node_b200_wide_vjp(void* const* host_ptrs, const int64_t* host_dims, float tsign, void* stream);
int node_b200_wide8_raw_operand(void* workspace, int which, const float* x, void* operand, const unsigned* max_bits, int N, int C,
                                void* stream);

/* wide_odefunc: one evaluation out = s * ODEfunc(s * t, y) of a C = 64 * nb model by ONE call (the sequence above); block_ws =
 * [2 convs][nb][nb] conv3x3 workspaces (conv3x3_prepare on W[64co:64co+64, 1+64ci:1+64ci+64]) block_ws_stride bytes apart,
 * bias / tmap = the convolutions' biases [C] and folded time maps [C,H,W], tmp_a / tmp_c = [N,C,H,W] scratch. */
int node_b200_wide_odefunc(void* block_ws, int64_t block_ws_stride, const float* y, float* out, float* tmp_a, float* tmp_c,
                           const float* g1w, const float* g1b, const float* g2w, const float* g2b, const float* g3w, const float* g3b,
                           const float* bias1, const float* tmap1, const float* bias2, const float* tmap2, const float* t_dev,
                           float tsign, int N, int C, int H, int W, void* stream);

/* wide_conv_blocks: out[:, 64o:64o+64] = sum_i conv3x3(x[:, 64i:64i+64], block (o, i)) - the nb^2 block launches of ONE wide
 * convolution by one call; block_ws = its nb x nb prepared workspaces (conv3x3_prepare), block_ws_stride bytes apart. */
int node_b200_wide_conv_blocks(void* block_ws, int64_t block_ws_stride, const float* x, float* out, int N, int C, int H, int W,
                               void* stream);

/* wide8: the wide dynamics on 8x8 maps (C = 128 / 256) as ONE TMA-fed tcgen05 implicit GEMM over all channels per convolution
 * (csrc/wide8_engine.cu) instead of (C/64)^2 block launches. ODEfunc.forward (model.py:339-348) =
 *   gn_operand(0, y) -> conv(0) -> gn_operand(1, c1 + b1 + t*Tmap1) -> conv(1) -> groupnorm_relu_ex(c2 + b2 + t*Tmap2).
 * prepare: packs W1[:, 1:], W2[:, 1:] ([C, C+1, 3, 3] ConcatConv2d weights, model.py:313-323) as fp16 hi/lo weight tiles at a
 *   power-of-two scale, the folded time maps and the operand scales (from |gamma|, |beta| of norm1 / norm2).
 * operand = node_b200_wide8_operand_bytes(N, C) bytes, ZERO-FILLED ONCE by the caller (the kernels never write the halo entries).
 * gn_operand(which): operand <- scale * relu(GroupNorm_32(x (+ add_bias + tsign * t * Tmap1 when which == 1))) as fp16 hi + lo.
 * conv(which): out[N, C, 8, 8] <- conv3x3(operand, W_which[:, 1:]) (no bias, no time channel).
 * odefunc: the five launches above; out = tsign * f(tsign * t, y). watchdog: 1 if a bounded barrier wait expired (synchronises). */
int64_t node_b200_wide8_workspace_bytes(int C, int H, int W);
int64_t node_b200_wide8_operand_bytes(int64_t N, int C);
int node_b200_wide8_prepare(void* workspace, int C, int H, int W, const float* conv1_w, const float* conv2_w, const float* g1w,
                            const float* g1b, const float* g2w, const float* g2b, void* stream);
int node_b200_wide8_gn_operand(void* workspace, int which, const float* x, void* operand, const float* gamma, const float* beta,
                               const float* add_bias, const float* t_dev, float tsign, int N, int C, void* stream);
int node_b200_wide8_conv(void* workspace, int which, const void* operand, float* out, int N, int C, void* stream);
int node_b200_wide8_watchdog(void* workspace, int C, void* stream);
int node_b200_wide8_odefunc(void* workspace, const float* y, float* out, void* operand, float* tmp_c, const float* g1w,
                            const float* g1b, const float* g2w, const float* g2b, const float* g3w, const float* g3b,
                            const float* bias1, const float* bias2, const float* t_dev, float tsign, int N, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NODE_B200_H_ */
