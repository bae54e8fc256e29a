// f16-split engine of the fused ODE-Net route: parameter preparation and shape dispatch.
// The kernel itself is the template in step_engine.cuh, instantiated per shape in step_shape_*.cu.
#include <cuda_fp16.h>
#include <cstdlib>
#include "fused_common.cuh"

namespace node {

// ---- parameter preparation for the f16 engine ------------------------------------------------------
// scal[0..1]: activation scales of conv1 / conv2 inputs, scal[2..3]: weight scales, scal[4..5]: 1/(sa*sw).
// Powers of two (exact): |relu(GN(x))| <= max|gamma| * sqrt(n) + max|beta| with n = 2*HW elements per
// GroupNorm cell bounds the activations, so a*sa stays below 2^15 (fp16 max 65504) and the low parts
// stay out of the fp16 subnormal range for everything that matters.
__global__ void k_prepare16_scales(FusedWs w, int HW, const float* c1w, const float* c2w, const float* g1w, const float* g1b,
                                   const float* g2w, const float* g2b) {
  __shared__ float red[4][256];
  const float* cw[2] = {c1w, c2w};
  const float* gw[2] = {g1w, g2w};
  const float* gb[2] = {g1b, g2b};
  const int tid = threadIdx.x;
  for (int cv = 0; cv < 2; ++cv) {
    float mw = 0.f, mg = 0.f, mb = 0.f;
    for (int i = tid; i < kC * (kC + 1) * 9; i += 256) mw = fmaxf(mw, fabsf(cw[cv][i]));
    for (int i = tid; i < kC; i += 256) { mg = fmaxf(mg, fabsf(gw[cv][i])); mb = fmaxf(mb, fabsf(gb[cv][i])); }
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
      w.scal[cv] = exp2f((float)ea);
      w.scal[2 + cv] = exp2f((float)ew);
      w.scal[4 + cv] = exp2f((float)(-ea - ew));
    }
    __syncthreads();
  }
}

__global__ void k_prepare16_tiles(FusedWs w, int H, int W, const float* c1w, const float* c2w) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const float* cw[2] = {c1w, c2w};
  // forward tiles: rows 0..63 = hi part of cout, 64..127 = lo part; 64 cin halves per row (128 B), SW128
  for (int i = tid; i < 2 * 9 * 128 * 64; i += nth) {
    int r = i;
    const int cin = r % 64; r /= 64;
    const int row = r % 128; r /= 128;
    const int tap = r % 9; r /= 9;
    const int cv = r;
    const int co = row & 63;
    const float v = cw[cv][((int64_t)co * (kC + 1) + cin + 1) * 9 + tap] * w.scal[2 + cv];
    const __half hi = __float2half_rn(v);
    const __half val = row < 64 ? hi : __float2half_rn(v - __half2float(hi));
    const int chunk = (cin >> 3) ^ (row & 7);
    const int64_t dst = ((int64_t)(cv * 9 + tap) * 128 + row) * 64 + chunk * 8 + (cin & 7);
    w.w16[dst] = *reinterpret_cast<const uint16_t*>(&val);
  }
  // CTA-pair half tiles (tcgen05.mma.cta_group::2: each CTA of the pair holds half of the B rows). Row order chosen so that
  // all three products of the split land in the right accumulator columns (see step8_engine.cuh / tools/pair_test.cu):
  //   CTA 0: rows 0..31 = w_hi[0..31],  rows 32..63 = w_lo[32..63];   CTA 1: rows 0..31 = w_hi[32..63], rows 32..63 = w_lo[0..31]
  for (int i = tid; i < 2 * 9 * 2 * 64 * 64; i += nth) {
    int r = i;
    const int cin = r % 64; r /= 64;
    const int row = r % 64; r /= 64;
    const int cta = r % 2; r /= 2;
    const int tap = r % 9; r /= 9;
    const int cv = r;
    const bool is_hi = row < 32;
    const int co = cta == 0 ? row : (is_hi ? 32 + row : row - 32);
    const float v = cw[cv][((int64_t)co * (kC + 1) + cin + 1) * 9 + tap] * w.scal[2 + cv];
    const __half hi = __float2half_rn(v);
    const __half val = is_hi ? hi : __float2half_rn(v - __half2float(hi));
    const int chunk = (cin >> 3) ^ (row & 7);
    const int64_t dst = (((int64_t)(cv * 9 + tap) * 2 + cta) * 64 + row) * 64 + chunk * 8 + (cin & 7);
    w.w16p[dst] = *reinterpret_cast<const uint16_t*>(&val);
  }
  // Tmap border classes: class = 3*rowclass + colclass, 0 = first, 1 = interior, 2 = last row / column
  for (int i = tid; i < 2 * 9 * kC; i += nth) {
    const int co = i % kC, cls = (i / kC) % 9, cv = i / (9 * kC);
    const int rc = cls / 3, cc = cls % 3;
    float s = 0.f;
    for (int tap = 0; tap < 9; ++tap) {
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      const bool in = !(dy < 0 && rc == 0) && !(dy > 0 && rc == 2) && !(dx < 0 && cc == 0) && !(dx > 0 && cc == 2);
      if (in) s += cw[cv][((int64_t)co * (kC + 1)) * 9 + tap];
    }
    w.tmapc[i] = s;
  }
  (void)H; (void)W;
}

int launch_prepare16(const FusedWs& w, int H, int W, const float* c1w, const float* c2w, const float* g1w, const float* g1b,
                     const float* g2w, const float* g2b, cudaStream_t st) {
  k_prepare16_scales<<<1, 256, 0, st>>>(w, H * W, c1w, c2w, g1w, g1b, g2w, g2b);
  NODE_CUDA_OK(cudaGetLastError());
  k_prepare16_tiles<<<148, 256, 0, st>>>(w, H, W, c1w, c2w);
  return (int)cudaGetLastError();
}

int launch_step_8x8(const FusedArgs& a, cudaStream_t st);
int launch_step8_dense(const FusedArgs& a, cudaStream_t st);
constexpr int kStep8MinBatch = 445;      // up to 444 images the strip engine runs its one-slot variant, one super-tile per CTA      // step8.cu: dense 8x8 tiling, software-pipelined chain
int launch_step_7x7(const FusedArgs& a, cudaStream_t st);
int launch_step_6x6(const FusedArgs& a, cudaStream_t st);
int launch_step_14x14(const FusedArgs& a, cudaStream_t st);
int launch_step_16x16(const FusedArgs& a, cudaStream_t st);

bool step_engine_supports(int H, int W) {
  return (H == 8 && W == 8) || (H == 7 && W == 7) || (H == 6 && W == 6) || (H == 14 && W == 14) || (H == 16 && W == 16);
}

int launch_step_engine(const FusedArgs& a, cudaStream_t st) {
  const int H = a.g.H, W = a.g.W;
  if (H == 8 && W == 8) {
    const char* dense = getenv("NODE_B200_STEP8");                 // "0" / "1": force the strip-tiled / the dense pair engine
    if (dense != nullptr && dense[0] == '0') return launch_step_8x8(a, st);
    if (dense != nullptr && dense[0] == '1') return launch_step8_dense(a, st);
    // Below kStep8MinBatch images a launch is one latency chain per CTA whatever the engine; the strip engine's chain is
    // shorter there (3 images per CTA on more SMs, no pair handshakes; 1.08 vs 1.25 ms per solve at batch 128): tools/engine_sweep.py.
    if (a.g.N < kStep8MinBatch) return launch_step_8x8(a, st);
    return launch_step8_dense(a, st);
  }
  if (H == 7 && W == 7) return launch_step_7x7(a, st);
  if (H == 6 && W == 6) return launch_step_6x6(a, st);
  if (H == 14 && W == 14) return launch_step_14x14(a, st);
  if (H == 16 && W == 16) return launch_step_16x16(a, st);
  return (int)cudaErrorInvalidValue;
}

}  // namespace node
