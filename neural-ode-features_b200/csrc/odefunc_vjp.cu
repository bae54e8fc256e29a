// K7 - native vector-Jacobian product of the ODE-Net dynamics for the adjoint's augmented system
// (reference adjoint.py:32-55): parameter preparation of the data-gradient weight tiles, shape dispatch of the
// fused VJP kernel (vjp_engine.cuh) and the weight-gradient GEMM (wgrad_engine.cuh), the deterministic fold of
// their per-CTA partials into the flat parameter gradient, and the C entry points.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdlib>
#include "fused_common.cuh"

namespace node {

constexpr int kNParam = 2 * (kC * (kC + 1) * 9 + kC) + 6 * kC;   // 75,392 (misc.py:5-7 order)

struct VjpWs {
  float* R[2]; float* GC[2];
  float* chan_part; double* t_part; float* wpart; unsigned* gc_max;
};

static int64_t vjp_ws_layout(void* base, int N, int C, int H, int W, VjpWs* out) {
  const int64_t E = (int64_t)N * C * H * W;
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t r = o; o = align_up(o + bytes, 1024); return r; };
  int64_t o_t[4];
  for (int i = 0; i < 4; ++i) o_t[i] = take(E * 4);
  const int64_t o_chan = take((int64_t)kMaxGrid * 384 * 4);
  const int64_t o_tp = take((int64_t)kMaxGrid * 8);
  const int64_t o_wp = take((int64_t)kWgSplits * 2 * 9 * 64 * kWgCols * 4);
  const int64_t o_gm = take(16);
  if (out != nullptr) {
    char* b = (char*)base;
    out->R[0] = (float*)(b + o_t[0]); out->R[1] = (float*)(b + o_t[1]);
    out->GC[0] = (float*)(b + o_t[2]); out->GC[1] = (float*)(b + o_t[3]);
    out->chan_part = (float*)(b + o_chan); out->t_part = (double*)(b + o_tp); out->wpart = (float*)(b + o_wp);
    out->gc_max = (unsigned*)(b + o_gm);
  }
  return o;
}

// Data-gradient weight tiles (sets 2, 3 of w16): dL/dr[q, ci] = sum_tap' sum_co GC[q + off(tap'), co] * W[co, ci+1, 8 - tap'],
// i.e. the forward implicit GEMM with "cout" = ci, "cin" = co and flipped taps. fp16 hi (rows 0..63) / lo (rows 64..127) of
// w * scal[2 + cv] (the forward tiles' power-of-two weight scale), 64 K-values per 128-byte row, SWIZZLE_128B image.
__global__ void k_prepare_dgrad_tiles(FusedWs w, const float* c1w, const float* c2w) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const float* cw[2] = {c1w, c2w};
  for (int i = tid; i < 2 * 9 * 128 * 64; i += nth) {
    int r = i;
    const int co = r % 64; r /= 64;          // K index
    const int row = r % 128; r /= 128;
    const int tap = r % 9; r /= 9;
    const int cv = r;
    const int ci = row & 63;
    const float v = cw[cv][((int64_t)co * (kC + 1) + ci + 1) * 9 + (8 - tap)] * w.scal[2 + cv];
    const __half hi = __float2half_rn(v);
    const __half val = row < 64 ? hi : __float2half_rn(v - __half2float(hi));
    const int chunk = (co >> 3) ^ (row & 7);
    const int64_t dst = ((int64_t)((2 + cv) * 9 + tap) * 128 + row) * 64 + chunk * 8 + (co & 7);
    w.w16[dst] = *reinterpret_cast<const uint16_t*>(&val);
  }
}

// CTA-pair half tiles of the adjoint pair engine (vjp8_engine.cuh): set 0, 1 = conv1, conv2 forward; set 2, 3 = their data
// gradients. CTA c of a pair holds the output channels [32c, 32c+32): rows 0..31 their hi parts, rows 32..63 their lo parts, so that
// the three products of the split (a_hi*w_hi, a_lo*w_hi, a_hi*w_lo; N = 64 each) land in the same 64 accumulator columns.
__global__ void k_prepare_pair3_tiles(FusedWs w, const float* c1w, const float* c2w) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const float* cw[2] = {c1w, c2w};
  for (int i = tid; i < 4 * 9 * 2 * 64 * 64; i += nth) {
    int r = i;
    const int k = r % 64; r /= 64;
    const int row = r % 64; r /= 64;
    const int cta = r % 2; r /= 2;
    const int tap = r % 9; r /= 9;
    const int set = r, cv = set & 1;
    const int n = 32 * cta + (row & 31);
    const float v = (set < 2 ? cw[cv][((int64_t)n * (kC + 1) + k + 1) * 9 + tap]
                             : cw[cv][((int64_t)k * (kC + 1) + n + 1) * 9 + (8 - tap)]) * w.scal[2 + cv];
    const __half hi = __float2half_rn(v);
    const __half val = row < 32 ? hi : __float2half_rn(v - __half2float(hi));
    const int chunk = (k >> 3) ^ (row & 7);
    const int64_t dst = (((int64_t)(set * 9 + tap) * 2 + cta) * 64 + row) * 64 + chunk * 8 + (k & 7);
    w.w16q[dst] = *reinterpret_cast<const uint16_t*>(&val);
  }
}

int launch_prepare_dgrad(const FusedWs& w, const float* c1w, const float* c2w, cudaStream_t st) {
  k_prepare_dgrad_tiles<<<148, 256, 0, st>>>(w, c1w, c2w);
  NODE_CUDA_OK(cudaGetLastError());
  k_prepare_pair3_tiles<<<148, 256, 0, st>>>(w, c1w, c2w);
  return (int)cudaGetLastError();
}

// Fold the per-CTA partials in a fixed order into (vjp_t, flat vjp_params) in func.parameters() order:
// norm1.w, norm1.b, conv1.W [64,65,3,3], conv1.b, norm2.w, norm2.b, conv2.W, conv2.b, norm3.w, norm3.b.
__global__ void k_vjp_finalize(const float* __restrict__ wpart, int nsplit, const float* __restrict__ chan_part, int ngrid,
                               const double* __restrict__ t_part, const float* __restrict__ t_dev, float tsign,
                               float* __restrict__ out_t, float* __restrict__ out_p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const float t_eff = tsign * t_dev[0];
  constexpr int kWsz = kC * (kC + 1) * 9, kBlock = 2 * kC + kWsz + kC;   // one (norm, conv) block: 37,632
  if (i == kNParam) {
    double s = 0.0;
    for (int b = 0; b < ngrid; ++b) s += t_part[b];
    out_t[0] = (float)s * tsign;
    return;
  }
  if (i > kNParam) return;
  const int blk = i / kBlock, r = i % kBlock;       // blk 0, 1: (norm, conv) pairs; blk 2: norm3 only
  float v = 0.f;
  if (r < 2 * kC) {                                  // GroupNorm affine gradients
    const int q = blk * 2 + r / kC, c = r % kC;
    for (int b = 0; b < ngrid; ++b) v += chan_part[(size_t)b * 384 + q * 64 + c];
  } else if (r < 2 * kC + kWsz) {                    // conv weight [co][ci1][tap]
    const int j = r - 2 * kC;
    const int co = j / ((kC + 1) * 9), rem = j % ((kC + 1) * 9), ci1 = rem / 9, tap = rem % 9;
    const int col = ci1 == 0 ? 64 : ci1 - 1;
    for (int s = 0; s < nsplit; ++s) v += wpart[(((size_t)(s * 2 + blk) * 9 + tap) * 64 + co) * kWgCols + col];
    if (ci1 == 0) v *= t_eff;                        // the time plane is t*ones (model.py:321)
  } else {                                           // conv bias: the ones column at the centre tap
    const int co = r - 2 * kC - kWsz;
    for (int s = 0; s < nsplit; ++s) v += wpart[(((size_t)(s * 2 + blk) * 9 + 4) * 64 + co) * kWgCols + 64];
  }
  out_p[i] = v * tsign;
}

// per-shape translation units (vjp_shape_HxW.cu)
int launch_vjp_8x8(const VjpArgs& a, cudaStream_t st);
int launch_vjp8_dense(const VjpArgs& a, cudaStream_t st, int* grid_out);      // vjp8.cu: dense 8x8 tiling on CTA pairs
constexpr int kVjp8MinBatch = 445;        // below, the strip engine's one super-tile per CTA is the shorter chain (as for k_step8)
int launch_vjp_7x7(const VjpArgs& a, cudaStream_t st);
int launch_vjp_6x6(const VjpArgs& a, cudaStream_t st);
int launch_vjp_14x14(const VjpArgs& a, cudaStream_t st);
int launch_vjp_16x16(const VjpArgs& a, cudaStream_t st);
int launch_wgrad_8x8(const WgradArgs& a, cudaStream_t st);
int launch_wgrad_7x7(const WgradArgs& a, cudaStream_t st);
int launch_wgrad_6x6(const WgradArgs& a, cudaStream_t st);
int launch_wgrad_14x14(const WgradArgs& a, cudaStream_t st);
int launch_wgrad_16x16(const WgradArgs& a, cudaStream_t st);
int launch_wgrad_15x15(const WgradArgs& a, cudaStream_t st);      // callers only (wgrad_shape_15x15.cu)
int launch_wgrad_13x13(const WgradArgs& a, cudaStream_t st);      // callers only, MNIST (wgrad_shape_13x13.cu)
int launch_wgrad8_dense(const WgradArgs& a, bool ones, cudaStream_t st);      // wgrad8.cu
// NODE_B200_WGRAD8=0: the strip-tiled k_wgrad<8,8> instead of the dense 8x8 kernel (kept as the second implementation the tests compare)
static bool wgrad8_enabled() { const char* e = getenv("NODE_B200_WGRAD8"); return !(e != nullptr && e[0] == '0'); }

}  // namespace node

using namespace node;

extern "C" int64_t node_b200_vjp_workspace_bytes(int N, int C, int H, int W) {
  Geo g;
  if (!make_geo(N, C, H, W, &g) || !step_engine_supports(H, W)) return -1;
  return vjp_ws_layout(nullptr, N, C, H, W, nullptr);
}

// Device pointers of the VJP workspace's buffers (0..3: R1, R2, GC1, GC2; 4: weight-gradient partials) - test aid.
extern "C" void* node_b200_vjp_buffer(void* vjp_workspace, int which, int N, int C, int H, int W) {
  VjpWs v;
  vjp_ws_layout(vjp_workspace, N, C, H, W, &v);
  switch (which) {
    case 0: return v.R[0];
    case 1: return v.R[1];
    case 2: return v.GC[0];
    case 3: return v.GC[1];
    case 4: return v.wpart;
    default: return nullptr;
  }
}

// operand scales of the fused workspace the last node_b200_odefunc_vjp call ran on (node_b200_wgrad is always called from there)
static const float* g_wgrad_scal = nullptr;

extern "C" int node_b200_wgrad(void* vjp_workspace, const float* r1, const float* gc1, const float* r2, const float* gc2,
                               int N, int C, int H, int W, void* stream) {
  WgradArgs a{};
  if (!make_geo(N, C, H, W, &a.g) || !step_engine_supports(H, W)) return (int)cudaErrorInvalidValue;
  VjpWs v;
  vjp_ws_layout(vjp_workspace, N, C, H, W, &v);
  a.R[0] = r1; a.R[1] = r2; a.GC[0] = gc1; a.GC[1] = gc2; a.part = v.wpart;
  a.gc_max[0] = v.gc_max; a.gc_max[1] = v.gc_max + 1; a.scal[0] = g_wgrad_scal; a.scal[1] = g_wgrad_scal + 1; a.ncv = 2;
  cudaStream_t st = (cudaStream_t)stream;
  if (H == 8 && W == 8) return wgrad8_enabled() ? launch_wgrad8_dense(a, true, st) : launch_wgrad_8x8(a, st);
  if (H == 7 && W == 7) return launch_wgrad_7x7(a, st);
  if (H == 6 && W == 6) return launch_wgrad_6x6(a, st);
  if (H == 14 && W == 14) return launch_wgrad_14x14(a, st);
  if (H == 16 && W == 16) return launch_wgrad_16x16(a, st);
  return (int)cudaErrorInvalidValue;
}

// The evaluation in two parts, so that a caller may run the second one on a side stream: (1) k_vjp - forward recompute, data
// gradients, the operands of the weight gradient; (2) k_wgrad + the fold into (vjp_t, vjp_params), which nothing but the stage
// combination of the parameter adjoint reads (adjoint.py:35: augmented_dynamics never looks at adj_t / adj_params).
struct VjpLaunch { VjpWs v; float tsign; bool use_dense; int grid_dense; int gs; };

static int vjp_part1(void* workspace, void* vjp_workspace, const float* y, const float* adj_y, const float* t_dev, float tsign,
                     float* f_out, float* vjp_y, int N, int C, int H, int W, cudaStream_t st, VjpLaunch* L) {
  VjpArgs a{};
  if (!make_geo(N, C, H, W, &a.g) || !step_engine_supports(H, W)) return (int)cudaErrorInvalidValue;
  ws_layout(workspace, N, C, H, W, &a.w);
  VjpWs& v = L->v;
  vjp_ws_layout(vjp_workspace, N, C, H, W, &v);
  a.y = y; a.adj = adj_y; a.f_out = f_out; a.vy_out = vjp_y;
  a.R[0] = v.R[0]; a.R[1] = v.R[1]; a.GC[0] = v.GC[0]; a.GC[1] = v.GC[1];
  a.chan_part = v.chan_part; a.t_part = v.t_part; a.t_dev = t_dev; a.tsign = tsign < 0 ? -1.f : 1.f; a.eps = 1e-5f;
  a.gc_max = v.gc_max;
  NODE_CUDA_OK(cudaMemsetAsync(v.gc_max, 0, 16, st));
  g_wgrad_scal = a.w.scal;
  int rc, grid_dense = 0;
  const char* dense = getenv("NODE_B200_VJP8");           // "0" / "1": force the strip-tiled / the dense pair engine
  const bool use_dense = H == 8 && W == 8 && (dense != nullptr ? dense[0] != '0' : N >= kVjp8MinBatch);
  if (use_dense) rc = launch_vjp8_dense(a, st, &grid_dense);
  else if (H == 8 && W == 8) rc = launch_vjp_8x8(a, st);
  else if (H == 7 && W == 7) rc = launch_vjp_7x7(a, st);
  else if (H == 6 && W == 6) rc = launch_vjp_6x6(a, st);
  else if (H == 14 && W == 14) rc = launch_vjp_14x14(a, st);
  else rc = launch_vjp_16x16(a, st);
  L->tsign = a.tsign; L->use_dense = use_dense; L->grid_dense = grid_dense; L->gs = a.g.gs;
  return rc;
}

static int vjp_part2(void* vjp_workspace, const VjpLaunch& L, const float* t_dev, float* vjp_t, float* vjp_params, int N, int C, int H,
                     int W, cudaStream_t st) {
  const VjpWs& v = L.v;
  int rc = node_b200_wgrad(vjp_workspace, v.R[0], v.GC[0], v.R[1], v.GC[1], N, C, H, W, (void*)st);
  if (rc != 0) return rc;
  // grid sizes the two kernels used (same rules as their launchers)
  const int per = strip_images(H, W);
  const int NST = (N + per - 1) / per;
  const int NSTV = (N + L.gs - 1) / L.gs;                  // k_vjp's super-tiles hold gs images (strip_gs)
  const int nst_vjp = L.use_dense ? L.grid_dense : (NSTV < kMaxGrid ? NSTV : kMaxGrid);
  const int nsplit = (H == 8 && W == 8 && wgrad8_enabled()) ? wgrad8_splits(N, 2) : (NST < kWgSplits ? NST : kWgSplits);
  k_vjp_finalize<<<(kNParam + 1 + 255) / 256, 256, 0, st>>>(v.wpart, nsplit, v.chan_part, nst_vjp, v.t_part, t_dev, L.tsign,
                                                              vjp_t, vjp_params);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_odefunc_vjp(void* workspace, void* vjp_workspace, const float* y, const float* adj_y,
                                     const float* t_dev, float tsign, float* f_out, float* vjp_y, float* vjp_t,
                                     float* vjp_params, int N, int C, int H, int W, void* stream) {
  VjpLaunch L;
  const int rc = vjp_part1(workspace, vjp_workspace, y, adj_y, t_dev, tsign, f_out, vjp_y, N, C, H, W, (cudaStream_t)stream, &L);
  if (rc != 0) return rc;
  return vjp_part2(vjp_workspace, L, t_dev, vjp_t, vjp_params, N, C, H, W, (cudaStream_t)stream);
}

// The same evaluation with its second part on `side_stream`, ordered after the first by `fork` (an event the caller owns; inside a
// stream capture the pair becomes a graph edge). The caller joins `side_stream` before anything reads vjp_t / vjp_params or
// rewrites this vjp_workspace.
extern "C" int node_b200_odefunc_vjp_split(void* workspace, void* vjp_workspace, const float* y, const float* adj_y, const float* t_dev,
                                           float tsign, float* f_out, float* vjp_y, float* vjp_t, float* vjp_params, int N, int C, int H,
                                           int W, void* stream, void* side_stream, void* fork_event) {
  VjpLaunch L;
  const int rc = vjp_part1(workspace, vjp_workspace, y, adj_y, t_dev, tsign, f_out, vjp_y, N, C, H, W, (cudaStream_t)stream, &L);
  if (rc != 0) return rc;
  NODE_CUDA_OK(cudaEventRecord((cudaEvent_t)fork_event, (cudaStream_t)stream));
  NODE_CUDA_OK(cudaStreamWaitEvent((cudaStream_t)side_stream, (cudaEvent_t)fork_event, 0));
  return vjp_part2(vjp_workspace, L, t_dev, vjp_t, vjp_params, N, C, H, W, (cudaStream_t)side_stream);
}

// ---- weight gradients of the callers' 3x3 stride-1 convolutions (model.py:119-178 under autograd) ------------------
// dW[p][co][ci][tap] = sum_{n,pos} GC_p[n,co,pos] * R_p[n,ci,pos + tap] for up to kWgMaxPairs (input, output-gradient) pairs in one
// k_wgrad launch (the stride-2 convolutions arrive as four stride-1 pairs on parity planes, caller_ops.py). Fixed-order fold.
namespace node {
__global__ void k_wgrad_fold(const float* __restrict__ wpart, int nsplit, int ncv, float* __restrict__ dw) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncv * kC * kC * 9) return;
  const int p = i / (kC * kC * 9), j = i % (kC * kC * 9);
  const int co = j / (kC * 9), ci = (j / 9) % kC, tap = j % 9;
  float v = 0.f;
  for (int s = 0; s < nsplit; ++s) v += wpart[(((size_t)(s * ncv + p) * 9 + tap) * 64 + co) * kWgCols + ci];
  dw[i] = v;
}
}  // namespace node

extern "C" int64_t node_b200_conv_wgrad_workspace_bytes(int npairs) {
  if (npairs < 1 || npairs > kWgMaxPairs) return 0;
  return (int64_t)wgrad_splits(1 << 30, npairs) * npairs * 9 * 64 * kWgCols * 4;
}

extern "C" int node_b200_conv_wgrad(void* workspace, int npairs, const float* const* inputs, const float* const* grads,
                                    const float* const* input_scales, const unsigned* const* grad_max_bits, float* dw,
                                    int N, int C, int H, int W, void* stream) {
  WgradArgs a{};
  if (npairs < 1 || npairs > kWgMaxPairs || N < 1 || C != kC) return (int)cudaErrorInvalidValue;
  if (!((H == 8 && W == 8) || (H == 15 && W == 15) || (H == 13 && W == 13) || (H == 7 && W == 7))) return (int)cudaErrorInvalidValue;
  if (!make_geo(N, C, H, W, &a.g)) return (int)cudaErrorInvalidValue;
  for (int p = 0; p < npairs; ++p) { a.R[p] = inputs[p]; a.GC[p] = grads[p]; a.scal[p] = input_scales[p]; a.gc_max[p] = grad_max_bits[p]; }
  a.part = (float*)workspace; a.ncv = npairs;
  cudaStream_t st = (cudaStream_t)stream;
  const bool dense = H == 8 && wgrad8_enabled();
  const int rc = dense ? launch_wgrad8_dense(a, false, st)
                       : (H == 8 ? launch_wgrad_8x8(a, st) : (H == 15 ? launch_wgrad_15x15(a, st) : (H == 13 ? launch_wgrad_13x13(a, st) : launch_wgrad_7x7(a, st))));
  if (rc != 0) return rc;
  const int per = strip_images(H, W);
  const int NST = (N + per - 1) / per;
  const int nsplit = dense ? wgrad8_splits(N, npairs) : wgrad_splits(NST, npairs);
  k_wgrad_fold<<<(npairs * kC * kC * 9 + 255) / 256, 256, 0, st>>>(a.part, nsplit, npairs, dw);
  return (int)cudaGetLastError();
}

// ---- one attempted step of the adjoint's backward integration as ONE call (SURVEY 8b `node_dopri5_adjoint_backward`, the
// per-step half of it; reference adjoint.py:23-102 through dopri5.py:94-122) ------------------------------------------------------
// The augmented state (y, adj_y, adj_t, adj_params) lives in the generic route's flat buffers (rows of `row_elems` floats:
// Y0 Y1 F0 F1 K2..K6 YMID YI, members at seg_off). The host loop issued 6 x (stage combination, augmented dynamics) + error norm
// + fold + controller as ~20 Python -> C calls per attempt; at the reference's batch of 128 those calls, not the kernels, were
// the step. Same kernels, same order, same arithmetic - enqueued from here.
extern "C" int node_b200_adjoint_step(void* ctl_v, float* bufs, int64_t row_elems, int cur, const int64_t* seg_off,
                                      const int64_t* seg_len, int n_seg, void* workspace, void* vjp_workspace, float tsign,
                                      int64_t ts32_offset_bytes, int N, int C, int H, int W, double* partials, double* sums,
                                      int* nonfinite_flag, const double* t_out, void* stream) {
  if (n_seg != 4) return (int)cudaErrorInvalidValue;
  node_ctl_t* ctl = (node_ctl_t*)ctl_v;
  enum { Y0 = 0, F0 = 2, K2 = 4, YI = 10 };
  auto row = [&](int i) { return bufs + (size_t)i * row_elems; };
  const int y = Y0 + cur, f = F0 + cur, yn = Y0 + (cur ^ 1), fn = F0 + (cur ^ 1);
  const int ks[7] = {f, K2, K2 + 1, K2 + 2, K2 + 3, K2 + 4, fn};
  const void* kp[7];
  for (int i = 0; i < 7; ++i) kp[i] = row(ks[i]);
  const float* ts32 = (const float*)((const char*)ctl_v + ts32_offset_bytes);
  for (int i = 0; i < 6; ++i) {
    float* dst = row(i < 5 ? YI : yn);
    NODE_CUDA_OK((cudaError_t)node_b200_rk_stage_combine(ctl, NODE_F32, i, dst, row(y), kp, i + 1, row_elems, stream));
    float* out = row(ks[i + 1]);
    NODE_CUDA_OK((cudaError_t)node_b200_odefunc_vjp(workspace, vjp_workspace, dst + seg_off[0], dst + seg_off[1], ts32 + (i + 1), tsign,
                                                    out + seg_off[0], out + seg_off[1], out + seg_off[2], out + seg_off[3], N, C, H, W,
                                                    stream));
  }
  NODE_CUDA_OK((cudaError_t)node_b200_rk_error_norm(ctl, NODE_F32, row(y), row(yn), kp, seg_off, seg_len, n_seg, partials,
                                                    nonfinite_flag, stream));
  NODE_CUDA_OK((cudaError_t)node_b200_reduce_partials(partials, 2 * n_seg, sums, stream));
  return node_b200_controller(ctl, 2, sums, nonfinite_flag, t_out, stream);
}
