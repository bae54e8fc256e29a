// Wide dynamics on 8x8 maps (n_filters = 128 / 256: the paper's CIFAR setting, reference reproduce.sh:21, model.py:326-348) as a
// TMA-fed tcgen05 implicit GEMM over ALL channels: one evaluation of ODEfunc.forward (model.py:339-348) is
//
//     k_wide_gn<operand>  a1 = relu(GN1(y))                        -> operand image (fp16 hi + lo)
//     k_wide_conv         c1 = conv3x3(a1, W1[:, 1:])              N = C accumulator columns, K = 9 * C
//     k_wide_gn<operand>  a2 = relu(GN2(c1 + b1 + t * Tmap1))      (time channel of ConcatConv2d folded, model.py:320-323)
//     k_wide_conv         c2 = conv3x3(a2, W2[:, 1:])
//     k_wide_gn<fp32>     k  = s * GN3(c2 + b2 + t * Tmap2)
//
// The block path of caller_conv.cu (node_b200_wide_odefunc) ran a convolution as (C/64)^2 launches of the 64-channel engine,
// each converting its input block again and read-modify-writing its output block. Here the GroupNorm pass writes the activation
// ONCE in the layout the tensor core reads - the dense 8x8 tiling of step8_engine.cuh (a row-slot = the 8 pixels of an image row
// + one zero entry, rows of two images interleaved, SBO = 144 B, so a 3x3 tap is a byte offset of the descriptor) - split into
// fp16 hi + lo parts at a power-of-two scale (fp32 contract: a_hi*w_hi + a_lo*w_hi + a_hi*w_lo accumulate in fp32). The
// convolution kernel is then a pure producer / issuer / epilogue pipeline:
//   * warp 0 streams operand stages (4 images x 32 channels x hi/lo = 41.9 KB, one bulk copy, 2-deep ring) and weight tiles
//     ((stage, tap): C rows x 32 channels x hi/lo = 128*C bytes, one bulk copy, 4-deep ring) through the TMA engine;
//   * warp 1 issues 12 tcgen05.mma (M = 128, N = C, K = 16) per weight tile: 2 M tiles (2 images each) x 2 K steps x 3
//     products; a weight tile is read from L2 once per 4 images;
//   * 8 epilogue warps drain the 2 x C accumulator columns to [N, C, 8, 8] fp32.
// A super-tile (4 images) is 9 * C / 32 weight tiles = 864 MMAs at C = 256 (110k clocks at the f16 rate); persistent CTAs, one
// per SM, take super-tiles round-robin.
#include <cuda_fp16.h>
#include <cstdlib>
#include "node_common.cuh"
#include "ptx.cuh"

namespace node { namespace w8 {

constexpr int kSlotB = 144;                       // bytes of a row-slot in one k-chunk (8 pixels + 1 zero entry, 8 halves each)
constexpr int kChunkSlots = 36;                   // [2 zero][tile 0: 16][2 zero][tile 1: 16]; the next chunk's zeros close it
constexpr int kLBO = kChunkSlots * kSlotB;        // 5184 B between k-chunks
constexpr int kStageChunks = 4;                   // k-chunks (8 channels) per operand stage
constexpr int kPartB = kStageChunks * kLBO;       // hi or lo part of one stage
constexpr int kLead = kSlotB, kTail = 2 * kSlotB;
constexpr int kStageB = kLead + 2 * kPartB + kTail;    // 41,904 B
constexpr int kARing = 2, kBRing = 4;
constexpr int kThreads = 320;                     // producer warp, issuer warp, 8 epilogue warps
constexpr int kImgs = 4;

__host__ __device__ constexpr uint32_t btile_bytes(int C) { return 128u * (uint32_t)C; }     // 2 parts x 4 chunks x C rows x 16 B
__host__ __device__ constexpr size_t conv_smem(int C) { return 128 + (size_t)kBRing * btile_bytes(C) + (size_t)kARing * kStageB + 128; }
__host__ __device__ constexpr uint32_t idesc(int n) { return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

// workspace: [scal 64 floats][tmap 2*C*64 floats][w16: 2 convs x (C/32 stages) x 9 taps x btile]
struct Ws { float* scal; unsigned* mx; float* tmap; uint8_t* w16; };
static int64_t ws_layout(void* base, int C, Ws* out) {
  const int64_t o_scal = 0, o_tmap = 1024, o_w16 = o_tmap + (int64_t)2 * C * 64 * 4;
  const int64_t total = o_w16 + (int64_t)2 * (C / 32) * 9 * btile_bytes(C);
  if (out != nullptr) {
    char* b = (char*)base;
    out->scal = (float*)(b + o_scal); out->mx = (unsigned*)(b + o_scal + 256); out->tmap = (float*)(b + o_tmap); out->w16 = (uint8_t*)(b + o_w16);
  }
  return total;
}

// ---- prepare ---------------------------------------------------------------------------------------------------------------
// mx[0], mx[1] = bit patterns of max |W1[:, 1:]|, max |W2[:, 1:]|; mx[2..5] = max |gamma1|, |beta1|, |gamma2|, |beta2|
__global__ void k_wide_absmax(Ws w, int C, const float* w1, const float* w2, const float* g1w, const float* g1b, const float* g2w, const float* g2b) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  const int per = (C + 1) * 9;
  float m[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = tid; i < C * per; i += nth) {
    if (i % per >= 9) { m[0] = fmaxf(m[0], fabsf(w1[i])); m[1] = fmaxf(m[1], fabsf(w2[i])); }
  }
  for (int i = tid; i < C; i += nth) {
    m[2] = fmaxf(m[2], fabsf(g1w[i])); m[3] = fmaxf(m[3], fabsf(g1b[i])); m[4] = fmaxf(m[4], fabsf(g2w[i])); m[5] = fmaxf(m[5], fabsf(g2b[i]));
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m[r] = fmaxf(m[r], __shfl_xor_sync(0xffffffffu, m[r], o));
    if ((threadIdx.x & 31) == 0 && m[r] > 0.f) atomicMax(w.mx + r, __float_as_uint(m[r]));
  }
}

// scal: [0] sa1, [1] sw1, [2] 1/(sa1*sw1), [3] sa2, [4] sw2, [5] 1/(sa2*sw2) - powers of two; |relu(GN(x))| <= |gamma| sqrt(L) + |beta|
__global__ void k_wide_scales(Ws w, int C) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float L = (float)((C / 32) * 64);
  for (int cv = 0; cv < 2; ++cv) {
    const float mw = __uint_as_float(w.mx[cv]), mg = __uint_as_float(w.mx[2 + 2 * cv]), mb = __uint_as_float(w.mx[3 + 2 * cv]);
    const float bound = mg * sqrtf(L) + mb;
    int ea = bound > 0.f ? (int)floorf(log2f(32768.0f / bound)) : 0;
    int ew = mw > 0.f ? (int)floorf(log2f(16384.0f / mw)) : 0;
    ea = max(-24, min(24, ea)); ew = max(-24, min(24, ew));
    w.scal[3 * cv] = exp2f((float)ea); w.scal[3 * cv + 1] = exp2f((float)ew); w.scal[3 * cv + 2] = exp2f((float)(-ea - ew));
  }
}

// weight image [conv][stage][tap][part hi/lo][chunk 4][co C][8 halves] (K-major core matrices of 8 rows x 16 B: SBO = 128 B,
// LBO = 16 * C) and the folded time maps Tmap[conv][c][y][x] = sum of the in-bounds taps of W[c, 0] (model.py:320-323)
__global__ void k_wide_tiles(Ws w, int C, const float* w1, const float* w2) {
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const int S = C / 32;
  const int64_t per_conv = (int64_t)S * 9 * 2 * 4 * C * 8;
  __half* img = reinterpret_cast<__half*>(w.w16);
  for (int64_t i = tid; i < 2 * per_conv; i += nth) {
    int64_t r = i;
    const int e = (int)(r % 8); r /= 8;
    const int co = (int)(r % C); r /= C;
    const int cj = (int)(r % 4); r /= 4;
    const int part = (int)(r % 2); r /= 2;
    const int tap = (int)(r % 9); r /= 9;
    const int s = (int)(r % S); r /= S;
    const int cv = (int)r;
    const int ci = 32 * s + 8 * cj + e;
    const float v = (cv ? w2 : w1)[((int64_t)co * (C + 1) + 1 + ci) * 9 + tap] * w.scal[3 * cv + 1];
    const __half hi = __float2half_rn(v);
    img[i] = part == 0 ? hi : __float2half_rn(v - __half2float(hi));
  }
  for (int64_t i = tid; i < (int64_t)2 * C * 64; i += nth) {
    const int pix = (int)(i % 64), c = (int)((i / 64) % C), cv = (int)(i / (64 * C));
    const int y = pix >> 3, x = pix & 7;
    const float* wt = (cv ? w2 : w1) + (int64_t)c * (C + 1) * 9;
    float s = 0.f;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int yy = y + ky - 1, xx = x + kx - 1;
        if (yy >= 0 && yy < 8 && xx >= 0 && xx < 8) s += wt[ky * 3 + kx];
      }
    w.tmap[i] = s;
  }
}

// ---- GroupNorm -> ReLU -> operand image -------------------------------------------------------------------------------------
// One block per (image, k-chunk of 8 channels): thread = (channel c, pixel quad q). CPG = channels per group (C / 32): 8 or 4.
__device__ __forceinline__ size_t operand_entry(int img, int S, int kc, int part, int pix) {
  const int st = img >> 2, il = img & 3, tile = il >> 1, which = il & 1;
  const int slot = 2 + tile * 18 + 2 * (pix >> 3) + which;
  return ((size_t)st * S + (kc >> 2)) * kStageB + kLead + (size_t)part * kPartB + (size_t)(kc & 3) * kLBO + (size_t)slot * kSlotB + (size_t)(pix & 7) * 16;
}

// One block per (image, stage of 4 k-chunks = 32 channels): thread = (channel c of each chunk, pixel quad q), 4 independent
// 128-bit loads in flight per thread. OPERAND: write scale * relu(GN(.)) as fp16 hi / lo entries of the operand image; otherwise
// write post * GN(.) as fp32 [N, C, 8, 8] (norm3).
template <int CPG, bool OPERAND>
__global__ void __launch_bounds__(128) k_wide_gn(const float* __restrict__ x, uint8_t* __restrict__ a16, float* __restrict__ y,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  const float* __restrict__ add_bias, const float* __restrict__ add_tmap,
                                                  const float* __restrict__ t_dev, float tsign, const float* __restrict__ scale, float post,
                                                  int C, float eps) {
  constexpr int WPG = CPG / 2;                    // warps per GroupNorm group
  __shared__ float red[2][4][4];
  __shared__ __align__(16) __half tile[OPERAND ? 4 : 1][2][64][8];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = C >> 5;
  const int img = blockIdx.x / S, stg = blockIdx.x % S;
  const int c = tid >> 4, q = tid & 15;
  float v[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 v4 = __ldg(reinterpret_cast<const float4*>(x + ((size_t)img * C + 32 * stg + 8 * j + c) * 64 + 4 * q));
    v[j][0] = v4.x; v[j][1] = v4.y; v[j][2] = v4.z; v[j][3] = v4.w;
  }
  if (add_bias != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float b = __ldg(add_bias + 32 * stg + 8 * j + c);
#pragma unroll
      for (int e = 0; e < 4; ++e) v[j][e] += b;
    }
  }
  if (add_tmap != nullptr) {
    const float t = tsign * __ldg(t_dev);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 m = __ldg(reinterpret_cast<const float4*>(add_tmap + (size_t)(32 * stg + 8 * j + c) * 64 + 4 * q));
      v[j][0] = fmaf(t, m.x, v[j][0]); v[j][1] = fmaf(t, m.y, v[j][1]); v[j][2] = fmaf(t, m.z, v[j][2]); v[j][3] = fmaf(t, m.w, v[j][3]);
    }
  }
  const int wg0 = (warp / WPG) * WPG;
  constexpr float inv_n = 1.0f / (float)(CPG * 64);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float s = warp_sum((v[j][0] + v[j][1]) + (v[j][2] + v[j][3]));
    if (lane == 0) red[0][j][warp] = s;
  }
  __syncthreads();
  float mean[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < WPG; ++i) tot += red[0][j][wg0 + i];
    mean[j] = tot * inv_n;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float q2 = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) { v[j][e] -= mean[j]; q2 = fmaf(v[j][e], v[j][e], q2); }
    q2 = warp_sum(q2);
    if (lane == 0) red[1][j][warp] = q2;
  }
  __syncthreads();
  const float sa = OPERAND ? __ldg(scale) : post;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float tot = 0.f;
#pragma unroll
    for (int i = 0; i < WPG; ++i) tot += red[1][j][wg0 + i];
    const float rstd = 1.0f / sqrtf(tot * inv_n + eps);
    const int ch = 32 * stg + 8 * j + c;
    const float ga = rstd * __ldg(gamma + ch), be = __ldg(beta + ch);
    if constexpr (OPERAND) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float r = fmaxf(fmaf(v[j][e], ga, be), 0.f) * sa;
        const __half hi = __float2half_rn(r);
        tile[j][0][4 * q + e][c] = hi;
        tile[j][1][4 * q + e][c] = __float2half_rn(r - __half2float(hi));
      }
    } else {
      float4 o;
      o.x = fmaf(v[j][0], ga, be) * sa; o.y = fmaf(v[j][1], ga, be) * sa; o.z = fmaf(v[j][2], ga, be) * sa; o.w = fmaf(v[j][3], ga, be) * sa;
      *reinterpret_cast<float4*>(y + ((size_t)img * C + ch) * 64 + 4 * q) = o;
    }
  }
  if constexpr (OPERAND) {
    __syncthreads();
    const int part = tid >> 6, pix = tid & 63;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 e = *reinterpret_cast<const uint4*>(&tile[j][part][pix][0]);
      *reinterpret_cast<uint4*>(a16 + operand_entry(img, S, 4 * stg + j, part, pix)) = e;
    }
  }
}

// Operand mode without a transposition: thread = (pixel, half of the stage's 32 channels) owns 16 channels of ONE pixel = two
// complete 16-byte operand entries (hi and lo): 16 warp-coalesced scalar loads in flight, statistics by warp sums + one exchange
// between the two warps of a half, four 16-byte stores.
template <int CPG>
__global__ void __launch_bounds__(128) k_wide_gn_op(const float* __restrict__ x, uint8_t* __restrict__ a16, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, const float* __restrict__ add_bias,
                                                     const float* __restrict__ add_tmap, const float* __restrict__ t_dev, float tsign,
                                                     const float* __restrict__ scale, int C, float eps) {
  constexpr int NG = 16 / CPG;                    // GroupNorm groups per thread
  __shared__ float red[2][4][NG];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = C >> 5;
  const int img = blockIdx.x / S, stg = blockIdx.x % S;
  const int px = tid & 63, h = tid >> 6, ch0 = 32 * stg + 16 * h;
  const float* xp = x + ((size_t)img * C + ch0) * 64 + px;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __ldg(xp + i * 64);
  if (add_bias != nullptr) {
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += __ldg(add_bias + ch0 + i);
  }
  if (add_tmap != nullptr) {
    const float t = tsign * __ldg(t_dev);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = fmaf(t, __ldg(add_tmap + (size_t)(ch0 + i) * 64 + px), v[i]);
  }
  constexpr float inv_n = 1.0f / (float)(CPG * 64);
  float mean[NG], rstd[NG];
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < CPG; ++i) s += v[g * CPG + i];
    s = warp_sum(s);
    if (lane == 0) red[0][warp][g] = s;
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    mean[g] = (red[0][2 * h][g] + red[0][2 * h + 1][g]) * inv_n;
    float q2 = 0.f;
#pragma unroll
    for (int i = 0; i < CPG; ++i) { v[g * CPG + i] -= mean[g]; q2 = fmaf(v[g * CPG + i], v[g * CPG + i], q2); }
    q2 = warp_sum(q2);
    if (lane == 0) red[1][warp][g] = q2;
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < NG; ++g) rstd[g] = 1.0f / sqrtf((red[1][2 * h][g] + red[1][2 * h + 1][g]) * inv_n + eps);
  const float sa = __ldg(scale);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float r[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int i = 8 * k + 2 * e + u;
        r[u] = fmaxf(fmaf(v[i], rstd[i / CPG] * __ldg(gamma + ch0 + i), __ldg(beta + ch0 + i)), 0.f) * sa;
      }
      const __half2 hh = __floats2half2_rn(r[0], r[1]);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(r[0] - hf.x, r[1] - hf.y);
      hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(a16 + operand_entry(img, S, 4 * stg + 2 * h + k, 0, px)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a16 + operand_entry(img, S, 4 * stg + 2 * h + k, 1, px)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- raw operand (data gradients of the adjoint: a SIGNED fp32 tensor, no GroupNorm) ------------------------------------------
// scal[3 * which] = sa = 2^floor(log2(16384 / max|x|)), scal[3 * which + 2] = 1 / (sa * sw): the operand scale is found from the
// tensor itself (gradients have no a-priori bound), the weight scale sw stays the prepared one.
__global__ void k_wide_dyn_scale(Ws w, int which, const unsigned* __restrict__ max_bits) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float m = __uint_as_float(*max_bits);
  int e = m > 0.f ? (int)floorf(log2f(16384.0f / m)) : 0;
  e = max(-40, min(40, e));
  const float sa = exp2f((float)e);
  w.scal[3 * which] = sa;
  w.scal[3 * which + 2] = 1.0f / (sa * w.scal[3 * which + 1]);
}

// Same mapping as k_wide_gn_op: thread = (pixel, half of the stage's 32 channels) writes two complete hi / lo operand entries.
__global__ void __launch_bounds__(128) k_wide_raw_op(const float* __restrict__ x, uint8_t* __restrict__ a16, const float* __restrict__ scale, int C) {
  const int tid = threadIdx.x;
  const int S = C >> 5;
  const int img = blockIdx.x / S, stg = blockIdx.x % S;
  const int px = tid & 63, h = tid >> 6, ch0 = 32 * stg + 16 * h;
  const float* xp = x + ((size_t)img * C + ch0) * 64 + px;
  float v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __ldg(xp + i * 64);
  const float sa = __ldg(scale);
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float r0 = v[8 * k + 2 * e] * sa, r1 = v[8 * k + 2 * e + 1] * sa;
      const __half2 hh = __floats2half2_rn(r0, r1);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
      hi[e] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    *reinterpret_cast<uint4*>(a16 + operand_entry(img, S, 4 * stg + 2 * h + k, 0, px)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(a16 + operand_entry(img, S, 4 * stg + 2 * h + k, 1, px)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

// ---- the convolution ----------------------------------------------------------------------------------------------------------
struct ConvArgs {
  const uint8_t* a16;       // operand image [super-tile][stage][kStageB]
  const uint8_t* w16;       // weight image of this convolution [stage][tap][btile]
  float* out;               // [N, C, 8, 8]
  const float* inv;         // device scalar 1 / (sa * sw)
  unsigned* watchdog;       // set to 1 when a bounded barrier wait expired
  int N;
};

// CN = output channels (accumulator columns) per work unit: CN == C is the large-batch configuration (a weight tile is one bulk
// copy, read once per 4 images); CN < C splits the output channels of a super-tile over C / CN CTAs so that a small batch (the
// reference's 128: 32 super-tiles) still fills the 148 SMs - a unit's weight tile is then the CN rows of each of the 8 (part,
// chunk) slabs (8 bulk copies into a dense [part][chunk][CN rows] image, LBO = 16 * CN).
template <int C, int CN>
__global__ void __launch_bounds__(kThreads, 1) k_wide_conv(const ConvArgs a) {
  constexpr int S = C / 32;
  constexpr int NSPLIT = C / CN;
  constexpr uint32_t kBTile = btile_bytes(CN);
  constexpr uint32_t kBTileFull = btile_bytes(C);
  constexpr uint32_t kCols = 2 * CN;              // two M tiles of CN accumulator columns
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t s0 = ptx::smem_u32(smem_raw);
  const uint32_t al = (s0 + 127u) & ~127u;
  uint8_t* base = smem_raw + (al - s0);
  const uint32_t bring = al, aring = al + kBRing * kBTile, bars = aring + kARing * kStageB;
  const uint32_t bar_afull = bars, bar_afree = bars + 16, bar_bfull = bars + 32, bar_bfree = bars + 64, bar_accfull = bars + 96, bar_accfree = bars + 104;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + (bars - al) + 112);
  if (tid == 0) {
    for (int i = 0; i < kARing; ++i) { ptx::mbar_init(bar_afull + 8 * i, 1); ptx::mbar_init(bar_afree + 8 * i, 1); }
    for (int i = 0; i < kBRing; ++i) { ptx::mbar_init(bar_bfull + 8 * i, 1); ptx::mbar_init(bar_bfree + 8 * i, 1); }
    ptx::mbar_init(bar_accfull, 1); ptx::mbar_init(bar_accfree, 8);
    ptx::fence_mbar_init();
  }
  if (warp == 0) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), kCols);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int NST = (a.N + kImgs - 1) / kImgs;
  const int NU = NST * NSPLIT;                    // work units: (super-tile, output-channel split)
  bool timeout = false;

  if (warp == 0) {
    // ---- producer: operand stages and weight tiles through the TMA engine
    const bool lead = ptx::elect_one();
    uint32_t ai = 0, bi = 0;
#pragma unroll 1
    for (int u = blockIdx.x; u < NU; u += gridDim.x) {
      const int st = u / NSPLIT, co0 = (u % NSPLIT) * CN;
#pragma unroll 1
      for (int s = 0; s < S; ++s, ++ai) {
        const uint32_t as = ai % kARing;
        if (ai >= (uint32_t)kARing && !timeout && !ptx::mbar_wait(bar_afree + 8 * as, ((ai / kARing) - 1) & 1)) timeout = true;
        if (lead) {
          ptx::mbar_expect_tx(bar_afull + 8 * as, kStageB);
          ptx::bulk_g2s(aring + as * kStageB, a.a16 + ((size_t)st * S + s) * kStageB, kStageB, bar_afull + 8 * as);
        }
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap, ++bi) {
          const uint32_t bs = bi % kBRing;
          if (bi >= (uint32_t)kBRing && !timeout && !ptx::mbar_wait(bar_bfree + 8 * bs, ((bi / kBRing) - 1) & 1)) timeout = true;
          if (lead) {
            ptx::mbar_expect_tx(bar_bfull + 8 * bs, kBTile);
            const uint8_t* src = a.w16 + ((size_t)s * 9 + tap) * kBTileFull;
            if (NSPLIT == 1) {
              ptx::bulk_g2s(bring + bs * kBTile, src, kBTile, bar_bfull + 8 * bs);
            } else {
#pragma unroll
              for (int slab = 0; slab < 8; ++slab)        // (part, chunk) slabs of C rows x 16 B: rows co0 .. co0 + CN
                ptx::bulk_g2s(bring + bs * kBTile + slab * (CN * 16), src + ((size_t)slab * C + co0) * 16, CN * 16, bar_bfull + 8 * bs);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- issuer: 12 MMAs per weight tile (2 M tiles x 2 K steps x 3 products), all into the tile's accumulator
    const bool lead = ptx::elect_one();
    constexpr uint32_t a_hiw = ((uint32_t)kSlotB >> 4) | (1u << 14);              // SBO = 144 B, descriptor version 1
    constexpr uint32_t b_hiw = (128u >> 4) | (1u << 14);                          // SBO = 128 B
    constexpr uint32_t kId = idesc(CN);
    auto pack = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
    uint32_t ai = 0, bi = 0, nacc = 0;
#pragma unroll 1
    for (int u = blockIdx.x; u < NU; u += gridDim.x, ++nacc) {
      if (nacc > 0 && !timeout && !ptx::mbar_wait(bar_accfree, (nacc - 1) & 1)) timeout = true;
      ptx::tc_fence_after();
#pragma unroll 1
      for (int s = 0; s < S; ++s, ++ai) {
        const uint32_t as = ai % kARing;
        if (!timeout && !ptx::mbar_wait(bar_afull + 8 * as, (ai / kARing) & 1)) timeout = true;
        const uint32_t abase = aring + as * kStageB + kLead + 2 * kSlotB;           // tile 0, row-slot 0, chunk 0, hi part
        const uint32_t a_lo0 = ((abase & 0x3FFFFu) >> 4) | (((uint32_t)kLBO >> 4) << 16);
#pragma unroll 1
        for (int tap = 0; tap < 9; ++tap, ++bi) {
          const uint32_t bs = bi % kBRing;
          if (!timeout && !ptx::mbar_wait(bar_bfull + 8 * bs, (bi / kBRing) & 1)) timeout = true;
          ptx::tc_fence_after();
          const int off = (tap / 3 - 1) * 2 * kSlotB + (tap % 3 - 1) * 16;
          const uint32_t a_tap = a_lo0 + (uint32_t)(off >> 4);                      // arithmetic shift: never borrows into the LBO field
          const uint32_t b_lo0 = (((bring + bs * kBTile) & 0x3FFFFu) >> 4) | (((16u * CN) >> 4) << 16);
          if (lead) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const uint32_t d = tmem + (uint32_t)(mt * CN);
#pragma unroll
              for (int ks = 0; ks < 2; ++ks) {
                const uint32_t a_w = a_tap + (uint32_t)((mt * 18 * kSlotB + 2 * ks * kLBO) >> 4);
                const uint32_t b_w = b_lo0 + (uint32_t)((2 * ks * 16 * CN) >> 4);
                const uint64_t a_hi = pack(a_w, a_hiw), a_lo = pack(a_w + (uint32_t)(kPartB >> 4), a_hiw);
                const uint64_t b_hi = pack(b_w, b_hiw), b_lo = pack(b_w + (uint32_t)((4 * 16 * CN) >> 4), b_hiw);
                ptx::mma_f16_ss(d, a_hi, b_hi, kId, (s == 0 && tap == 0 && ks == 0) ? 0u : 1u);
                ptx::mma_f16_ss(d, a_lo, b_hi, kId, 1u);
                ptx::mma_f16_ss(d, a_hi, b_lo, kId, 1u);
              }
            }
            ptx::tc_commit(bar_bfree + 8 * bs);
          }
          __syncwarp();
        }
        if (lead) ptx::tc_commit(bar_afree + 8 * as);
        __syncwarp();
      }
      if (lead) ptx::tc_commit(bar_accfull);
      __syncwarp();
    }
  } else {
    // ---- epilogue: TMEM lane = position of the M tile, column = output channel
    const int q = warp & 3, mt = (warp - 2) >> 2;
    const int r = 32 * q + lane, g = r >> 3;
    const int pix = (g >> 1) * 8 + (r & 7);
    const float inv = __ldg(a.inv);
    uint32_t nacc = 0;
#pragma unroll 1
    for (int u = blockIdx.x; u < NU; u += gridDim.x, ++nacc) {
      const int st = u / NSPLIT, co0 = (u % NSPLIT) * CN;
      if (!timeout && !ptx::mbar_wait_relaxed(bar_accfull, nacc & 1)) timeout = true;
      ptx::tc_fence_after();
      const int img = st * kImgs + mt * 2 + (g & 1);
      const bool valid = img < a.N;
      float* o = a.out + ((size_t)(valid ? img : 0) * C + co0) * 64 + pix;
      const uint32_t t0 = tmem + ((uint32_t)(32 * q) << 16) + (uint32_t)(mt * CN);
#pragma unroll 1
      for (int b = 0; b < CN / 32; ++b) {
        uint32_t v[32];
        ptx::tmem_ld32(t0 + 32u * b, v);
        ptx::tc_wait_ld();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 32; ++j) o[(size_t)(32 * b + j) * 64] = __uint_as_float(v[j]) * inv;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_accfree);
    }
  }
  if (timeout && a.watchdog != nullptr) *a.watchdog = 1u;
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) ptx::tmem_dealloc(tmem, kCols);
}

template <int C, int CN>
static int launch_conv_split(const ConvArgs& a, cudaStream_t st) {
  constexpr size_t smem = conv_smem(CN);
  static_assert(smem <= 227 * 1024, "shared memory budget");
  NODE_SET_SMEM_ONCE((k_wide_conv<C, CN>), smem);
  const int NU = ((a.N + kImgs - 1) / kImgs) * (C / CN);
  const int grid = NU < 148 ? NU : 148;
  k_wide_conv<C, CN><<<grid, kThreads, smem, st>>>(a);
  return (int)cudaGetLastError();
}

// small batches: split the output channels until the units fill the SMs (NODE_B200_WIDE8_CN = 0 / 64 / 128 overrides)
template <int C>
static int launch_conv(const ConvArgs& a, cudaStream_t st) {
  const int NST = (a.N + kImgs - 1) / kImgs;
  int cn = C;
  if (NST * 2 <= 148) cn = C / 2;
  if (NST * 4 <= 148 + 20 && C >= 256) cn = C / 4;
  const char* e = getenv("NODE_B200_WIDE8_CN");
  if (e != nullptr && atoi(e) > 0) cn = atoi(e);
  if (cn == 64 && C >= 128) return launch_conv_split<C, 64>(a, st);
  if (cn == 128 && C >= 256) return launch_conv_split<C, 128>(a, st);
  return launch_conv_split<C, C>(a, st);
}

static int launch_gn(const float* x, uint8_t* a16, float* y, const float* gamma, const float* beta, const float* add_bias, const float* add_tmap,
                     const float* t_dev, float tsign, const float* scale, float post, int N, int C, cudaStream_t st) {
  const unsigned grid = (unsigned)((int64_t)N * (C / 32));
  if (a16 != nullptr) {
    if (C == 256) k_wide_gn_op<8><<<grid, 128, 0, st>>>(x, a16, gamma, beta, add_bias, add_tmap, t_dev, tsign, scale, C, 1e-5f);
    else k_wide_gn_op<4><<<grid, 128, 0, st>>>(x, a16, gamma, beta, add_bias, add_tmap, t_dev, tsign, scale, C, 1e-5f);
  } else {
    if (C == 256) k_wide_gn<8, false><<<grid, 128, 0, st>>>(x, nullptr, y, gamma, beta, add_bias, add_tmap, t_dev, tsign, nullptr, post, C, 1e-5f);
    else k_wide_gn<4, false><<<grid, 128, 0, st>>>(x, nullptr, y, gamma, beta, add_bias, add_tmap, t_dev, tsign, nullptr, post, C, 1e-5f);
  }
  return (int)cudaGetLastError();
}

}}  // namespace node::w8

using namespace node;

static bool wide8_ok(int C, int H, int W) { return (C == 128 || C == 256) && H == 8 && W == 8; }

extern "C" int64_t node_b200_wide8_workspace_bytes(int C, int H, int W) {
  return wide8_ok(C, H, W) ? w8::ws_layout(nullptr, C, nullptr) : 0;
}

extern "C" int64_t node_b200_wide8_operand_bytes(int64_t N, int C) {
  if (N < 1 || (C != 128 && C != 256)) return 0;
  return ((N + w8::kImgs - 1) / w8::kImgs) * (int64_t)(C / 32) * w8::kStageB;
}

extern "C" int node_b200_wide8_prepare(void* workspace, int C, int H, int W, const float* conv1_w, const float* conv2_w, const float* g1w,
                                       const float* g1b, const float* g2w, const float* g2b, void* stream) {
  if (!wide8_ok(C, H, W)) return (int)cudaErrorInvalidValue;
  w8::Ws w; w8::ws_layout(workspace, C, &w);
  cudaStream_t st = (cudaStream_t)stream;
  NODE_CUDA_OK(cudaMemsetAsync(w.scal, 0, 1024, st));
  w8::k_wide_absmax<<<148, 256, 0, st>>>(w, C, conv1_w, conv2_w, g1w, g1b, g2w, g2b);
  NODE_CUDA_OK(cudaGetLastError());
  w8::k_wide_scales<<<1, 32, 0, st>>>(w, C);
  NODE_CUDA_OK(cudaGetLastError());
  w8::k_wide_tiles<<<296, 256, 0, st>>>(w, C, conv1_w, conv2_w);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_wide8_gn_operand(void* workspace, int which, const float* x, void* operand, const float* gamma, const float* beta,
                                          const float* add_bias, const float* t_dev, float tsign, int N, int C, void* stream) {
  if (N < 1 || (C != 128 && C != 256) || which < 0 || which > 1) return (int)cudaErrorInvalidValue;
  w8::Ws w; w8::ws_layout(workspace, C, &w);
  const float* tmap = (add_bias != nullptr && which == 1) ? w.tmap : nullptr;           // GN2 follows conv1 (Tmap1)
  return w8::launch_gn(x, (uint8_t*)operand, nullptr, gamma, beta, add_bias, tmap, t_dev, tsign, w.scal + 3 * which, 1.f, N, C, (cudaStream_t)stream);
}

extern "C" int node_b200_wide8_raw_operand(void* workspace, int which, const float* x, void* operand, const unsigned* max_bits, int N, int C,
                                           void* stream) {
  if (N < 1 || (C != 128 && C != 256) || which < 0 || which > 1) return (int)cudaErrorInvalidValue;
  w8::Ws w; w8::ws_layout(workspace, C, &w);
  cudaStream_t st = (cudaStream_t)stream;
  w8::k_wide_dyn_scale<<<1, 32, 0, st>>>(w, which, max_bits);
  NODE_CUDA_OK(cudaGetLastError());
  w8::k_wide_raw_op<<<(unsigned)((int64_t)N * (C / 32)), 128, 0, st>>>(x, (uint8_t*)operand, w.scal + 3 * which, C);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_wide8_conv(void* workspace, int which, const void* operand, float* out, int N, int C, void* stream) {
  if (N < 1 || (C != 128 && C != 256) || which < 0 || which > 1) return (int)cudaErrorInvalidValue;
  w8::Ws w; w8::ws_layout(workspace, C, &w);
  w8::ConvArgs a{};
  a.a16 = (const uint8_t*)operand; a.w16 = w.w16 + (size_t)which * (C / 32) * 9 * w8::btile_bytes(C); a.out = out; a.inv = w.scal + 3 * which + 2;
  a.watchdog = w.mx + 8; a.N = N;
  return C == 256 ? w8::launch_conv<256>(a, (cudaStream_t)stream) : w8::launch_conv<128>(a, (cudaStream_t)stream);
}

extern "C" int node_b200_wide8_watchdog(void* workspace, int C, void* stream) {
  w8::Ws w; w8::ws_layout(workspace, C, &w);
  unsigned v = 0;
  NODE_CUDA_OK(cudaMemcpyAsync(&v, w.mx + 8, 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  NODE_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  return (int)v;
}

extern "C" int node_b200_wide8_odefunc(void* workspace, const float* y, float* out, void* operand, float* tmp_c, const float* g1w,
                                       const float* g1b, const float* g2w, const float* g2b, const float* g3w, const float* g3b,
                                       const float* bias1, const float* bias2, const float* t_dev, float tsign, int N, int C, void* stream) {
  if (N < 1 || (C != 128 && C != 256)) return (int)cudaErrorInvalidValue;
  w8::Ws w; w8::ws_layout(workspace, C, &w);
  NODE_CUDA_OK((cudaError_t)node_b200_wide8_gn_operand(workspace, 0, y, operand, g1w, g1b, nullptr, t_dev, tsign, N, C, stream));
  NODE_CUDA_OK((cudaError_t)node_b200_wide8_conv(workspace, 0, operand, tmp_c, N, C, stream));
  NODE_CUDA_OK((cudaError_t)node_b200_wide8_gn_operand(workspace, 1, tmp_c, operand, g2w, g2b, bias1, t_dev, tsign, N, C, stream));
  NODE_CUDA_OK((cudaError_t)node_b200_wide8_conv(workspace, 1, operand, tmp_c, N, C, stream));
  return w8::launch_gn(tmp_c, nullptr, out, g3w, g3b, bias2, w.tmap + (size_t)C * 64, t_dev, tsign, nullptr, tsign < 0 ? -1.0f : 1.0f, N, C,
                       (cudaStream_t)stream);
}
