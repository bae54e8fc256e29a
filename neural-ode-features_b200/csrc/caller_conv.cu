// C ABI of the fused ResBlock tail (resconv_engine.cuh): parameter preparation and shape dispatch.
#include <cuda_fp16.h>
#include "resconv_engine.cuh"

namespace node {

int launch_resconv_15x15(const ResConvArgs& a, cudaStream_t st);
int launch_resconv_8x8(const ResConvArgs& a, cudaStream_t st);
int launch_resconv_13x13(const ResConvArgs& a, cudaStream_t st);
int launch_resconv_7x7(const ResConvArgs& a, cudaStream_t st);

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
  for (int i = tid; i < kC; i += 256) { mg = fmaxf(mg, fabsf(gw[i])); mb = fmaxf(mb, fabsf(gb[i])); }
  red[0][tid] = mw; red[1][tid] = mg; red[2][tid] = mb;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (tid < s) for (int r = 0; r < 3; ++r) red[r][tid] = fmaxf(red[r][tid], red[r][tid + s]);
    __syncthreads();
  }
  if (tid == 0) {
    const float bound_a = red[1][0] * sqrtf((float)(kCpg * HW)) + red[2][0];
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

}  // namespace node

using namespace node;

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
                                         const float* gn_b, int N, int C, int H, int W, float eps, void* stream) {
  if (N < 1 || node_b200_resconv_workspace_bytes(C, H, W) <= 0) return (int)cudaErrorInvalidValue;
  ResConvWs w; resconv_layout(workspace, &w);
  ResConvArgs a{};
  a.w16 = w.w16; a.scal = w.scal; a.gamma = gn_w; a.beta = gn_b; a.x = x; a.shortcut = shortcut; a.out = out; a.N = N; a.eps = eps;
  cudaStream_t st = (cudaStream_t)stream;
  if (H == 15) return launch_resconv_15x15(a, st);
  if (H == 8) return launch_resconv_8x8(a, st);
  if (H == 13) return launch_resconv_13x13(a, st);
  return launch_resconv_7x7(a, st);
}
