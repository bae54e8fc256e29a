// C ABI of the fused ResBlock tail (resconv_engine.cuh): parameter preparation and shape dispatch.
#include <cuda_fp16.h>
#include "resconv_engine.cuh"
#include "convs2_engine.cuh"

namespace node {

int launch_resconv_15x15(const ResConvArgs& a, cudaStream_t st);
int launch_resconv_8x8(const ResConvArgs& a, cudaStream_t st);
int launch_resconv_13x13(const ResConvArgs& a, cudaStream_t st);
int launch_resconv_7x7(const ResConvArgs& a, cudaStream_t st);
int launch_resconv_raw_15x15(const ResConvArgs& a, cudaStream_t st);
int launch_resconv_raw_8x8(const ResConvArgs& a, cudaStream_t st);
int launch_resconv_raw_13x13(const ResConvArgs& a, cudaStream_t st);
int launch_resconv_raw_7x7(const ResConvArgs& a, cudaStream_t st);

struct ResConvWs { uint16_t* w16; float* scal; };

static int64_t resconv_layout(void* base, ResConvWs* out) {
  const int64_t o_w16 = 0, o_scal = (int64_t)9 * kW16TileBytes;
  if (out != nullptr) { out->w16 = (uint16_t*)((char*)base + o_w16); out->scal = (float*)((char*)base + o_scal); }
  return o_scal + 1024;
}

// Power-of-two operand scales (exact), same bounds as the step engine's (odefunc_step.cu): |relu(GN(x))| <=
// max|gamma| * sqrt(n) + max|beta| with n elements per GroupNorm cell.
__global__ void k_resconv_scales(ResConvWs w, int HW, const float* cw, const float* gw, const float* gb) {
  __shared__ float red[3][256];
  const int tid = threadIdx.x;
  float mw = 0.f, mg = 0.f, mb = 0.f;
  for (int i = tid; i < kC * kC * 9; i += 256) mw = fmaxf(mw, fabsf(cw[i]));
  if (gw != nullptr)
    for (int i = tid; i < kC; i += 256) { mg = fmaxf(mg, fabsf(gw[i])); mb = fmaxf(mb, fabsf(gb[i])); }
  red[0][tid] = mw; red[1][tid] = mg; red[2][tid] = mb;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) for (int r = 0; r < 3; ++r) red[r][tid] = fmaxf(red[r][tid], red[r][tid + s]);
    __syncthreads();
  }
  if (tid == 0) {
    const float bound_a = red[1][0] * sqrtf((float)(kCpg * HW)) + red[2][0];      // no GroupNorm given (raw operand): scale 1
    int ea = bound_a > 0.f ? (int)floorf(log2f(32768.0f / bound_a)) : 0;
    int ew = red[0][0] > 0.f ? (int)floorf(log2f(16384.0f / red[0][0])) : 0;
    ea = max(-24, min(24, ea)); ew = max(-24, min(24, ew));
    w.scal[0] = exp2f((float)ea); w.scal[1] = exp2f((float)ew); w.scal[2] = exp2f((float)(-ea - ew));
  }
}

// rows 0..63 = hi part of cout, 64..127 = lo part; 64 cin halves per row (128 B), SW128; weight layout [cout][cin][3][3]
__global__ void k_resconv_tiles(ResConvWs w, const float* cw) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < 9 * 128 * 64; i += nth) {
    int r = i;
    const int cin = r % 64; r /= 64;
    const int row = r % 128; r /= 128;
    const int tap = r;
    const int co = row & 63;
    const float v = cw[((int64_t)co * kC + cin) * 9 + tap] * w.scal[1];
    const __half hi = __float2half_rn(v);
    const __half val = row < 64 ? hi : __float2half_rn(v - __half2float(hi));
    const int chunk = (cin >> 3) ^ (row & 7);
    w.w16[((int64_t)tap * 128 + row) * 64 + chunk * 8 + (cin & 7)] = *reinterpret_cast<const uint16_t*>(&val);
  }
}

// ---- strided ResBlock head (convs2_engine.cuh) ----------------------------------------------------------------------
int launch_convs2_30x30(const ConvS2Args& a, cudaStream_t st);
int launch_convs2_15x15(const ConvS2Args& a, cudaStream_t st);
int launch_convs2_26x26(const ConvS2Args& a, cudaStream_t st);
int launch_convs2_13x13(const ConvS2Args& a, cudaStream_t st);

static int64_t convs2_layout(void* base, ResConvWs* out) {
  const int64_t o_scal = (int64_t)kS2Tiles * kW16TileBytes;
  if (out != nullptr) { out->w16 = (uint16_t*)base; out->scal = (float*)((char*)base + o_scal); }
  return o_scal + 1024;
}

// scal[0] activation scale, [1] / [4] weight scales of conv1 / shortcut, [2] / [3] their inverse products
__global__ void k_convs2_scales(ResConvWs w, int HWin, const float* cw, const float* dw, const float* gw, const float* gb) {
  __shared__ float red[4][256];
  const int tid = threadIdx.x;
  float mw = 0.f, md = 0.f, mg = 0.f, mb = 0.f;
  for (int i = tid; i < kC * kC * 9; i += 256) mw = fmaxf(mw, fabsf(cw[i]));
  for (int i = tid; i < kC * kC; i += 256) md = fmaxf(md, fabsf(dw[i]));
  for (int i = tid; i < kC; i += 256) { mg = fmaxf(mg, fabsf(gw[i])); mb = fmaxf(mb, fabsf(gb[i])); }
  red[0][tid] = mw; red[1][tid] = md; red[2][tid] = mg; red[3][tid] = mb;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) for (int r = 0; r < 4; ++r) red[r][tid] = fmaxf(red[r][tid], red[r][tid + s]);
    __syncthreads();
  }
  if (tid == 0) {
    const float bound_a = red[2][0] * sqrtf((float)(kCpg * HWin)) + red[3][0];
    int ea = bound_a > 0.f ? (int)floorf(log2f(32768.0f / bound_a)) : 0;
    int ew = red[0][0] > 0.f ? (int)floorf(log2f(16384.0f / red[0][0])) : 0;
    int ed = red[1][0] > 0.f ? (int)floorf(log2f(16384.0f / red[1][0])) : 0;
    ea = max(-24, min(24, ea)); ew = max(-24, min(24, ew)); ed = max(-24, min(24, ed));
    w.scal[0] = exp2f((float)ea); w.scal[1] = exp2f((float)ew); w.scal[4] = exp2f((float)ed);
    w.scal[2] = exp2f((float)(-ea - ew)); w.scal[3] = exp2f((float)(-ea - ed));
  }
}

// tiles in the kernel's issue order: 3x3 taps 0, 2, 6, 8, 1, 7, 3, 5, 4, then the 1x1 shortcut
__global__ void k_convs2_tiles(ResConvWs w, const float* cw, const float* dw) {
  const int order[kS2Tiles] = {0, 2, 6, 8, 1, 7, 3, 5, 4, -1};
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = tid; i < kS2Tiles * 128 * 64; i += nth) {
    int r = i;
    const int cin = r % 64; r /= 64;
    const int row = r % 128; r /= 128;
    const int t = r;
    const int co = row & 63, tap = order[t];
    const float v = tap >= 0 ? cw[((int64_t)co * kC + cin) * 9 + tap] * w.scal[1] : dw[(int64_t)co * kC + cin] * w.scal[4];
    const __half hi = __float2half_rn(v);
    const __half val = row < 64 ? hi : __float2half_rn(v - __half2float(hi));
    const int chunk = (cin >> 3) ^ (row & 7);
    w.w16[((int64_t)t * 128 + row) * 64 + chunk * 8 + (cin & 7)] = *reinterpret_cast<const uint16_t*>(&val);
  }
}

}  // namespace node

using namespace node;

extern "C" int64_t node_b200_convs2_workspace_bytes(int C, int HI, int WI) {
  if (C != kC) return 0;
  const bool ok = (HI == 30 && WI == 30) || (HI == 15 && WI == 15) || (HI == 26 && WI == 26) || (HI == 13 && WI == 13);
  return ok ? convs2_layout(nullptr, nullptr) : 0;
}

extern "C" int node_b200_convs2_prepare(void* workspace, int C, int HI, int WI, const float* conv_w, const float* down_w,
                                        const float* gn_w, const float* gn_b, void* stream) {
  if (node_b200_convs2_workspace_bytes(C, HI, WI) <= 0) return (int)cudaErrorInvalidValue;
  ResConvWs w; convs2_layout(workspace, &w);
  cudaStream_t st = (cudaStream_t)stream;
  k_convs2_scales<<<1, 256, 0, st>>>(w, HI * WI, conv_w, down_w, gn_w, gn_b);
  NODE_CUDA_OK(cudaGetLastError());
  k_convs2_tiles<<<148, 256, 0, st>>>(w, conv_w, down_w);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_convs2_forward(void* workspace, const float* act, float* c_out, float* sc_out, int N, int C, int HI,
                                        int WI, void* stream) {
  if (N < 1 || node_b200_convs2_workspace_bytes(C, HI, WI) <= 0) return (int)cudaErrorInvalidValue;
  ResConvWs w; convs2_layout(workspace, &w);
  ConvS2Args a{};
  a.w16 = w.w16; a.scal = w.scal; a.act = act; a.c_out = c_out; a.sc_out = sc_out; a.N = N;
  cudaStream_t st = (cudaStream_t)stream;
  if (HI == 30) return launch_convs2_30x30(a, st);
  if (HI == 15) return launch_convs2_15x15(a, st);
  if (HI == 26) return launch_convs2_26x26(a, st);
  return launch_convs2_13x13(a, st);
}

extern "C" int64_t node_b200_resconv_workspace_bytes(int C, int H, int W) {
  if (C != kC) return 0;
  const bool ok = (H == 15 && W == 15) || (H == 8 && W == 8) || (H == 13 && W == 13) || (H == 7 && W == 7);
  return ok ? resconv_layout(nullptr, nullptr) : 0;
}

extern "C" int node_b200_resconv_prepare(void* workspace, int C, int H, int W, const float* conv_w, const float* gn_w,
                                         const float* gn_b, void* stream) {
  if (node_b200_resconv_workspace_bytes(C, H, W) <= 0) return (int)cudaErrorInvalidValue;
  ResConvWs w; resconv_layout(workspace, &w);
  cudaStream_t st = (cudaStream_t)stream;
  k_resconv_scales<<<1, 256, 0, st>>>(w, H * W, conv_w, gn_w, gn_b);
  NODE_CUDA_OK(cudaGetLastError());
  k_resconv_tiles<<<148, 256, 0, st>>>(w, conv_w);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_resconv_forward(void* workspace, const float* x, const float* shortcut, float* out, const float* gn_w,
                                         const float* gn_b, const float* next_gn_w, const float* next_gn_b, int N, int C, int H,
                                         int W, float eps, void* stream) {
  if (N < 1 || node_b200_resconv_workspace_bytes(C, H, W) <= 0) return (int)cudaErrorInvalidValue;
  ResConvWs w; resconv_layout(workspace, &w);
  ResConvArgs a{};
  a.w16 = w.w16; a.scal = w.scal; a.gamma = gn_w; a.beta = gn_b; a.gamma_next = next_gn_w; a.beta_next = next_gn_w != nullptr ? next_gn_b : nullptr; a.x = x; a.shortcut = shortcut; a.out = out; a.N = N; a.eps = eps;
  cudaStream_t st = (cudaStream_t)stream;
  if (H == 15) return launch_resconv_15x15(a, st);
  if (H == 8) return launch_resconv_8x8(a, st);
  if (H == 13) return launch_resconv_13x13(a, st);
  return launch_resconv_7x7(a, st);
}

// Raw 3x3 stride-1 convolution of a signed tensor (the data gradients of the callers' convolutions, caller_ops.py):
// out = conv(x, W) (+ addend). prepare() packs an ordinary [64,64,3,3] weight - the caller passes the flipped, transposed
// kernel for a data gradient - and its power-of-two scale; the operand scale is found per super-tile inside the kernel.
extern "C" int node_b200_conv3x3_prepare(void* workspace, int C, int H, int W, const float* conv_w, void* stream) {
  if (node_b200_resconv_workspace_bytes(C, H, W) <= 0) return (int)cudaErrorInvalidValue;
  ResConvWs w; resconv_layout(workspace, &w);
  cudaStream_t st = (cudaStream_t)stream;
  k_resconv_scales<<<1, 256, 0, st>>>(w, H * W, conv_w, nullptr, nullptr);
  NODE_CUDA_OK(cudaGetLastError());
  k_resconv_tiles<<<148, 256, 0, st>>>(w, conv_w);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_conv3x3_forward(void* workspace, const float* x, const float* addend, float* out, int N, int C, int H, int W,
                                         void* stream) {
  return node_b200_conv3x3_forward_strided(workspace, x, addend, out, N, C, H, W, 0, 0, stream);
}

extern "C" int node_b200_conv3x3_forward_strided(void* workspace, const float* x, const float* addend, float* out, int N, int C, int H,
                                                 int W, int64_t in_image_stride, int64_t out_image_stride, void* stream) {
  if (N < 1 || node_b200_resconv_workspace_bytes(C, H, W) <= 0) return (int)cudaErrorInvalidValue;
  ResConvWs w; resconv_layout(workspace, &w);
  ResConvArgs a{};
  a.in_stride = in_image_stride; a.out_stride = out_image_stride;
  a.w16 = w.w16; a.scal = w.scal; a.x = x; a.shortcut = addend; a.out = out; a.N = N; a.eps = 0.f;
  cudaStream_t st = (cudaStream_t)stream;
  if (H == 15) return launch_resconv_raw_15x15(a, st);
  if (H == 8) return launch_resconv_raw_8x8(a, st);
  if (H == 13) return launch_resconv_raw_13x13(a, st);
  return launch_resconv_raw_7x7(a, st);
}

extern "C" int64_t node_b200_resconv_scal_offset(void) { return (int64_t)9 * kW16TileBytes; }
extern "C" int64_t node_b200_convs2_scal_offset(void) { return (int64_t)kS2Tiles * kW16TileBytes; }

// One evaluation of a wide ODEfunc (C = 64 * nb; node_b200/wide.py, model.py:339-348) enqueued by ONE call: GN1 -> ReLU,
// nb x nb block convolutions, GN2 (+ folded time channel) -> ReLU, nb x nb block convolutions, GN3 (+ folded time channel).
// block_ws = [2 layers][nb co][nb ci] prepared conv3x3 workspaces, block_ws_stride bytes apart; tmp_a / tmp_c = [N,C,H,W] scratch.
extern "C" int node_b200_wide_odefunc(void* block_ws, int64_t block_ws_stride, const float* y, float* out, float* tmp_a, float* tmp_c,
                                      const float* g1w, const float* g1b, const float* g2w, const float* g2b, const float* g3w,
                                      const float* g3b, const float* bias1, const float* tmap1, const float* bias2,
                                      const float* tmap2, const float* t_dev, float tsign, int N, int C, int H, int W, void* stream) {
  if (C % kC != 0 || C <= kC || N < 1) return (int)cudaErrorInvalidValue;
  const int nb = C / kC, HW = H * W;
  const int64_t stride = (int64_t)C * HW;
  auto convs = [&](int layer, const float* a, float* c) -> int {
    for (int co = 0; co < nb; ++co)
      for (int ci = 0; ci < nb; ++ci) {
        char* ws = (char*)block_ws + ((int64_t)(layer * nb + co) * nb + ci) * block_ws_stride;
        float* dst = c + (int64_t)co * kC * HW;
        const int rc = node_b200_conv3x3_forward_strided(ws, a + (int64_t)ci * kC * HW, ci ? dst : nullptr, dst, N, kC, H, W, stride, stride, stream);
        if (rc != 0) return rc;
      }
    return 0;
  };
  NODE_CUDA_OK((cudaError_t)node_b200_groupnorm_relu(y, tmp_a, g1w, g1b, N, C, 32, HW, 1e-5f, 1, stream));
  NODE_CUDA_OK((cudaError_t)convs(0, tmp_a, tmp_c));
  NODE_CUDA_OK((cudaError_t)node_b200_groupnorm_relu_ex(tmp_c, tmp_a, g2w, g2b, bias1, tmap1, t_dev, tsign, 1.0f, N, C, 32, HW, 1e-5f, 1, stream));
  NODE_CUDA_OK((cudaError_t)convs(1, tmp_a, tmp_c));
  return node_b200_groupnorm_relu_ex(tmp_c, out, g3w, g3b, bias2, tmap2, t_dev, tsign, tsign < 0 ? -1.0f : 1.0f, N, C, 32, HW, 1e-5f, 0, stream);
}

// out[:, 64o:64o+64] = sum_i conv3x3(x[:, 64i:64i+64], block (o, i)) for a [N, C, H, W] tensor, C = 64 * nb: one call for the
// nb^2 launches of one wide convolution; block_ws points at the nb x nb prepared workspaces of this convolution (forward weights,
// or the transposed / flipped ones of the data gradient, node_b200/wide.py).
extern "C" int node_b200_wide_conv_blocks(void* block_ws, int64_t block_ws_stride, const float* x, float* out, int N, int C, int H, int W,
                                          void* stream) {
  if (C % kC != 0 || C <= kC || N < 1) return (int)cudaErrorInvalidValue;
  const int nb = C / kC, HW = H * W;
  const int64_t stride = (int64_t)C * HW;
  for (int o = 0; o < nb; ++o)
    for (int i = 0; i < nb; ++i) {
      char* ws = (char*)block_ws + ((int64_t)o * nb + i) * block_ws_stride;
      float* dst = out + (int64_t)o * kC * HW;
      const int rc = node_b200_conv3x3_forward_strided(ws, x + (int64_t)i * kC * HW, i ? dst : nullptr, dst, N, kC, H, W, stride, stride, stream);
      if (rc != 0) return rc;
    }
  return 0;
}
