// f16-split engine of the fused ODE-Net route (reference model.py:326-348 inside rk_common.py:49-52).
//
// Same contract as odefunc_fused.cu (one launch = f0 / initial-step probe / all six dopri5 stages of
// an attempted step / one plain evaluation), different mapping to the SM:
//
//   * one worker THREAD owns one position of a zero-padded image strip and keeps its 64 channels in
//     registers from the Runge-Kutta stage input to k_i: stage combination, the three GroupNorms,
//     both ReLUs and both convolution epilogues never leave the register file; GroupNorm statistics
//     are warp transpose-reductions (+ one shared-memory fold across the warps of an image);
//   * the 3x3 convolution is an implicit GEMM whose A operand the tensor core reads STRAIGHT from
//     shared memory (SS-mode tcgen05.mma): the normalised activation is written once per conv as a
//     K-major, un-swizzled image [k-chunk][position][8 halves] in which images are separated by a
//     shared zero column / zero row, so a conv tap is nothing but a row offset in the A descriptor -
//     no per-tap staging by the threads (that staging bound the previous engine, profiles/r01a);
//   * fp32 contract by operand splitting in FP16 (same 11-bit significands as TF32, twice the MMA
//     rate and half the bytes): a = a_hi + a_lo, w = w_hi + w_lo after exact power-of-two scaling,
//     D = a_hi*[w_hi ; w_lo] (one N=128 MMA) + a_lo*w_hi (one N=64 MMA), fp32 accumulation in TMEM;
//   * weight tiles (16 KB per tap: hi and lo halves, UMMA SWIZZLE_128B image) stream L2 -> shared
//     through the TMA engine into a 4-deep ring shared by two independent worker slots: while the
//     tensor core runs the taps of one slot the other slot's threads do epilogue / GroupNorm work.
#pragma once
#include <cuda_fp16.h>
#include <cstdlib>
#include "fused_common.cuh"

namespace node {

#ifndef NODE_KNW
#define NODE_KNW 4
#endif
constexpr int kNW = NODE_KNW;             // weight ring depth (taps)
constexpr int kWGap = 2;                  // a ring slot is refilled once the tap two back has retired
constexpr int kTbFloats = 16 * 9 * 4;     // per slot: bias + t*Tmap of ONE convolution, [4-channel block][border class][4 channels]
constexpr float kGnIllCond = 16.0f;       // one-pass GroupNorm moments are redone in two passes when mean^2 > 16 var

template <int H_, int W_>
struct Tile {
  static constexpr int H = H_, W = W_, HW = H_ * W_;
  static constexpr int Wp = W_ + 1;                 // row pitch: one shared zero column
  static constexpr int IS = (H_ + 1) * Wp;          // image stride: one shared zero row
  static constexpr int SPAN = (H_ - 1) * Wp + W_;   // positions from an image's first to its last pixel
  static constexpr int MT = SPAN <= 256 ? 2 : (SPAN + 127) / 128;   // 128-row M tiles per super-tile
  static constexpr int P = MT * 128;                // positions = worker threads of a slot
  static constexpr int G = (P - SPAN) / IS + 1;     // images per super-tile
  static constexpr int HALO = Wp + 1;
  static constexpr int R = P + 2 * HALO;            // rows of the A image
  static constexpr int LBO = R * 16;                // bytes between k-chunks (8 halves) of the A image
  static constexpr int A_PART = 8 * LBO;            // hi or lo part
  static constexpr int NWARP = P / 32;
  static_assert(IS >= 32, "a warp may straddle at most two images");
  static_assert(MT * 128 <= 512, "accumulators exceed tensor memory");
};

__host__ __device__ constexpr size_t step_smem_bytes(int A_PART, int NSLOT, int NWARP, int G) {
  return 1024 + (size_t)kNW * kW16TileBytes + (size_t)NSLOT * 2 * A_PART + (size_t)NSLOT * NWARP * 64 * 4 +
         (size_t)NSLOT * G * 32 * 8 + (size_t)NSLOT * G * 32 * 16 + 3 * 32 * 16 + (size_t)NSLOT * kTbFloats * 4 + 2 * 9 * 64 * 4 + 2 * 64 * 4 + 64 * 4 + 32 * 8 +
         16 * 8 + 64;
}

// Tuning aid: when enabled (node_b200_step_debug), CTA 0 records clock64() stamps of every conv job of each slot:
// [slot][job][0..3] = A image published, turn acquired, MMAs issued, accumulators complete.
#ifdef NODE_STEP_DEBUG
static __device__ long long g_step_dbg[2 * 256 * 4];
static __device__ long long g_step_dbg2[2 * 256 * 2];   // clocks the leader waited for weights to land / for ring slots to free
static __device__ long long g_step_phase[2 * 16];       // per slot: clocks thread 32 of CTA 0 spent in each phase of the main loop
#define NODE_STAMP(i) do { if (rec_ph) { const long long now_ = clock64(); g_step_phase[me.slot * 16 + (i)] += now_ - last_ph; last_ph = now_; } } while (0)
#else
#define NODE_STAMP(i) do { } while (0)
#endif

struct StepSmem {
  uint32_t wring;        // shared address of the weight ring
  uint32_t abase;        // shared address of slot 0's A image (hi part)
  float* part;           // [NSLOT][NWARP][2][32] warp partials of the GroupNorm reductions (k_vjp: [NWARP][2][16])
  float2* stat;          // [NSLOT][G][32] (mean, rstd)
  float4* aff;           // [NSLOT][G][32] (a0, a1, b0, b1): GN(x) = a*x + b for the two channels of a group (k_step)
  float4* tb;            // [NSLOT][16][9] bias + t*Tmap of the convolution whose epilogue the slot runs next (k_step)
  uint32_t* illcond;     // [NSLOT] set by a fold thread when a one-pass variance is ill-conditioned (k_step)
  float4* gnp;           // [3][32] (gamma0, gamma1, beta0, beta1) per group
  float* bias;           // [2][64]
  float* tmapc;          // [2][9][64]
  float* coef;           // [8][8] h * coefficient table
  double* scratch;       // 32 doubles
  uint32_t bar_wfull, bar_wfree, bar_turn, bar_acc;
  volatile uint32_t* ring;   // [0] weight tiles requested, [1] taps consumed (owned by whoever holds the turn)
  uint32_t* tmem_slot;
};

__device__ __forceinline__ void slot_sync(int slot, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(nthreads) : "memory");
}

// Sum 16 per-lane values across the warp; lanes L and L^16 return the total of u[L & 15] (16 shuffles).
__device__ __forceinline__ float xreduce16(const float (&u)[16], int lane) {
  float b[8], c[4], d[2];
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? u[i] : u[i + 8], keep = up ? u[i + 8] : u[i];
      b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? b[i] : b[i + 4], keep = up ? b[i + 4] : b[i];
      c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = lane & 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? c[i] : c[i + 2], keep = up ? c[i + 2] : c[i];
      d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
  }
  const bool up = lane & 1;
  const float send = up ? d[0] : d[1], keep = up ? d[1] : d[0];
  const float r = keep + __shfl_xor_sync(0xffffffffu, send, 1);
  return r + __shfl_xor_sync(0xffffffffu, r, 16);
}

// Sum 32 per-lane values across the warp; lane L returns the total of u[L] (31 shuffles).
__device__ __forceinline__ float xreduce32(const float (&u)[32], int lane) {
  float a[16];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float send = up ? u[i] : u[i + 16], keep = up ? u[i + 16] : u[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  float b[8], c[4], d[2];
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? a[i] : a[i + 8], keep = up ? a[i + 8] : a[i];
      b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? b[i] : b[i + 4], keep = up ? b[i + 4] : b[i];
      c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = lane & 2;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? c[i] : c[i + 2], keep = up ? c[i + 2] : c[i];
      d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
  }
  const bool up = lane & 1;
  const float send = up ? d[0] : d[1], keep = up ? d[1] : d[0];
  return keep + __shfl_xor_sync(0xffffffffu, send, 1);
}

// Per-thread constants of a worker.
struct Who {
  int slot, wt, warp, lane;      // slot, position in the strip, warp within the slot
  int img_l, cls;                // image within the super-tile, Tmap border class
  bool inimg;                    // the position is a pixel (not padding)
  bool straddle, isB;            // the warp holds two images; this lane belongs to the second
  int pix;                       // h*W + w
};

// All per-pixel work runs on HALF of the channels at a time (32 channels = 16 GroupNorm groups): GroupNorm
// cells never couple the halves, and 32 live values per thread leave room for two worker slots per SM.

// One GroupNorm reduction round over 16 groups: val(j) summed over every pixel of the image ->
// fin(&stat[img][16*hb + j], total). A warp that straddles two images reduces twice.
template <class T, class Val, class Fin>
__device__ __forceinline__ void gn_reduce(const StepSmem& sm, const Who& me, int hb, Val val, Fin fin) {
  float* part = sm.part + (me.slot * T::NWARP + me.warp) * 32;
  {
    float u[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) u[j] = val(j);
    if (!me.straddle) {
      const float r = xreduce16(u, me.lane);
      if (me.lane < 16) part[me.lane] = r;
    } else {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = me.isB ? 0.f : u[j];
      const float ra = xreduce16(v, me.lane);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = me.isB ? u[j] : 0.f;
      const float rb = xreduce16(v, me.lane);
      if (me.lane < 16) { part[me.lane] = ra; part[16 + me.lane] = rb; }
    }
  }
  slot_sync(me.slot, T::P);
  if (me.wt < T::G * 16) {
    const int img = me.wt >> 4, g = me.wt & 15;
    const float* pp = sm.part + me.slot * T::NWARP * 32;
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < T::NWARP; ++w) {
      const int ia = (w * 32) / T::IS, ib = (w * 32 + 31) / T::IS;
      if (ia == img) tot += pp[w * 32 + g];
      if (ib != ia && ib == img) tot += pp[w * 32 + 16 + g];
    }
    fin(sm.stat + (me.slot * T::G + img) * 32 + 16 * hb + g, tot);
  }
  slot_sync(me.slot, T::P);
}

// GroupNorm statistics of x (channels [32*hb, 32*hb+32) of this thread's pixel) over the whole image: two
// passes (mean, then squared deviations) like the reference's native_group_norm; leaves (mean, rstd) in sm.stat.
template <class T>
__device__ __forceinline__ void gn_stats(const StepSmem& sm, const Who& me, int hb, const float (&x)[32], bool valid, float eps) {
  constexpr float inv_n = 1.0f / (float)(kCpg * T::HW);
  gn_reduce<T>(sm, me, hb, [&](int j) { return x[2 * j] + x[2 * j + 1]; },          // x is zero on padding
               [&](float2* dst, float tot) { dst->x = tot * inv_n; });
  const float2* st = sm.stat + (me.slot * T::G + min(me.img_l, T::G - 1)) * 32 + 16 * hb;
  gn_reduce<T>(sm, me, hb,
               [&](int j) {
                 const float m = st[j].x;
                 const float d0 = x[2 * j] - m, d1 = x[2 * j + 1] - m;
                 return valid ? fmaf(d0, d0, d1 * d1) : 0.f;
               },
               [&](float2* dst, float tot) { dst->y = 1.0f / sqrtf(tot * inv_n + eps); });
}

// k_step's GroupNorm statistics: ONE reduction round. Every thread contributes, for each of its 16 groups, the sum and the
// sum of squares of its two channels; the fold threads turn the totals into the affine form GN(x) = a*x + b of
// GroupNorm `n` (sm.aff). var = E[x^2] - mean^2 loses digits when |mean| >> std, so a fold thread that sees
// mean^2 > kGnIllCond * var raises the slot's flag and the whole slot repeats the statistics with the two-pass scheme
// of the reference's native_group_norm (gn_reduce with squared deviations) - rare, slot-uniform, deterministic.
template <class T>
__device__ __forceinline__ void gn_affine(const StepSmem& sm, const Who& me, int hb, int n, const float (&x)[32], bool valid, float eps,
                                          float post = 1.0f) {
  constexpr float inv_n = 1.0f / (float)(kCpg * T::HW);
  float* part = sm.part + (me.slot * T::NWARP + me.warp) * 64;
#pragma unroll
  for (int q = 0; q < 2; ++q) {              // q = 0: sums, q = 1: sums of squares (x is zero on padding)
    float u[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      u[j] = q == 0 ? x[2 * j] + x[2 * j + 1] : fmaf(x[2 * j], x[2 * j], x[2 * j + 1] * x[2 * j + 1]);
    if (!me.straddle) {
      const float r = xreduce16(u, me.lane);
      if (me.lane < 16) part[16 * q + me.lane] = r;
    } else {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = me.isB ? 0.f : u[j];
      const float ra = xreduce16(v, me.lane);
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = me.isB ? u[j] : 0.f;
      const float rb = xreduce16(v, me.lane);
      if (me.lane < 16) { part[16 * q + me.lane] = ra; part[32 + 16 * q + me.lane] = rb; }
    }
  }
  slot_sync(me.slot, T::P);
  const bool folder = me.wt < T::G * 16;
  const int fimg = me.wt >> 4, fg = me.wt & 15;
  auto publish = [&](float mean, float var) {
    const float rstd = 1.0f / sqrtf(var + eps);
    const float4 p = sm.gnp[n * 32 + 16 * hb + fg];
    const float a0 = rstd * p.x, a1 = rstd * p.y;      // `post` (a power of two or +-1: exact) is the operand scale / time sign
    sm.aff[(me.slot * T::G + fimg) * 32 + 16 * hb + fg] = make_float4(a0 * post, a1 * post, (p.z - a0 * mean) * post, (p.w - a1 * mean) * post);
  };
  if (folder) {
    const float* pp = sm.part + me.slot * T::NWARP * 64;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < T::NWARP; ++w) {
      const int ia = (w * 32) / T::IS, ib = (w * 32 + 31) / T::IS;
      if (ia == fimg) { s1 += pp[w * 64 + fg]; s2 += pp[w * 64 + 16 + fg]; }
      if (ib != ia && ib == fimg) { s1 += pp[w * 64 + 32 + fg]; s2 += pp[w * 64 + 48 + fg]; }
    }
    const float mean = s1 * inv_n;
    const float var = fmaxf(fmaf(-mean, mean, s2 * inv_n), 0.f);
    sm.stat[(me.slot * T::G + fimg) * 32 + 16 * hb + fg].x = mean;
    if (mean * mean > kGnIllCond * var) sm.illcond[me.slot] = 1u;
    publish(mean, var);
  }
  slot_sync(me.slot, T::P);
  if (*reinterpret_cast<volatile uint32_t*>(sm.illcond + me.slot) != 0u) {           // slot-uniform: read after the barrier, cleared behind the next one
    // the one-pass means are accurate (plain sums); every cell of the slot gets the two-pass variance
    const float2* st = sm.stat + (me.slot * T::G + min(me.img_l, T::G - 1)) * 32 + 16 * hb;
    float* p32 = sm.part + (me.slot * T::NWARP + me.warp) * 64;
    {
      float u[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float m = st[j].x;
        const float d0 = x[2 * j] - m, d1 = x[2 * j + 1] - m;
        u[j] = valid ? fmaf(d0, d0, d1 * d1) : 0.f;
      }
      if (!me.straddle) {
        const float r = xreduce16(u, me.lane);
        if (me.lane < 16) p32[me.lane] = r;
      } else {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = me.isB ? 0.f : u[j];
        const float ra = xreduce16(v, me.lane);
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = me.isB ? u[j] : 0.f;
        const float rb = xreduce16(v, me.lane);
        if (me.lane < 16) { p32[me.lane] = ra; p32[32 + me.lane] = rb; }
      }
    }
    slot_sync(me.slot, T::P);
    if (me.wt == 0) sm.illcond[me.slot] = 0u;
    if (folder) {
      const float* pp = sm.part + me.slot * T::NWARP * 64;
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < T::NWARP; ++w) {
        const int ia = (w * 32) / T::IS, ib = (w * 32 + 31) / T::IS;
        if (ia == fimg) tot += pp[w * 64 + fg];
        if (ib != ia && ib == fimg) tot += pp[w * 64 + 32 + fg];
      }
      publish(sm.stat[(me.slot * T::G + fimg) * 32 + 16 * hb + fg].x, tot * inv_n);
    }
    slot_sync(me.slot, T::P);
  }
}

// relu(a*x + b) * scale (sm.aff of this thread's image) split into fp16 hi + lo and written into this position's row
// of the A image (k-chunks [4*hb, 4*hb+4)).
template <class T>
__device__ __forceinline__ void affine_to_A(const StepSmem& sm, const Who& me, int hb, const float (&x)[32], bool valid, bool split) {
  const float4* af = sm.aff + (me.slot * T::G + min(me.img_l, T::G - 1)) * 32 + 16 * hb;
  const uint32_t row = sm.abase + me.slot * 2 * T::A_PART + (T::HALO + me.wt) * 16 + 4 * hb * T::LBO;
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = kc * 4 + j;
      const float4 p = af[g];
      const float r0 = fmaxf(fmaf(x[2 * g], p.x, p.z), 0.f);         // the operand scale is folded into p (gn_affine `post`)
      const float r1 = fmaxf(fmaf(x[2 * g + 1], p.y, p.w), 0.f);
      const __half2 h = __floats2half2_rn(r0, r1);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    if (valid) {
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + kc * T::LBO), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
      if (split)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + T::A_PART + kc * T::LBO), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
    }
  }
}

// relu(GN(x)) * scale split into fp16 hi + lo and written into this position's row of the A image
// (k-chunks [4*hb, 4*hb+4)).
template <class T>
__device__ __forceinline__ void gn_apply_to_A(const StepSmem& sm, const Who& me, int hb, const float (&x)[32], int n, float scale,
                                              bool valid, bool split) {
  const float2* st = sm.stat + (me.slot * T::G + min(me.img_l, T::G - 1)) * 32 + 16 * hb;
  const float4* gp = sm.gnp + n * 32 + 16 * hb;
  const uint32_t row = sm.abase + me.slot * 2 * T::A_PART + (T::HALO + me.wt) * 16 + 4 * hb * T::LBO;
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = kc * 4 + j;
      const float2 s = st[g];
      const float4 p = gp[g];
      const float a0 = s.y * p.x, a1 = s.y * p.y;
      const float b0 = p.z - a0 * s.x, b1 = p.w - a1 * s.x;
      const float r0 = fmaxf(fmaf(x[2 * g], a0, b0), 0.f) * scale;
      const float r1 = fmaxf(fmaf(x[2 * g + 1], a1, b1), 0.f) * scale;
      const __half2 h = __floats2half2_rn(r0, r1);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    if (valid) {
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + kc * T::LBO), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
      if (split)
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + T::A_PART + kc * T::LBO), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
    }
  }
}

constexpr uint32_t kIdF16N128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);   // kind::f16, fp16 x fp16 -> fp32, M128
constexpr uint32_t kIdF16N64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// Global conv-job schedule of a CTA: jobs alternate between the slots while both have super-tiles left.
struct Jobs {
  uint32_t jobs_full, jobs;     // jobs in rounds where every slot works / all jobs
  uint32_t single = 0;          // 1: every job is convolution 0 (k_resconv); 0: a slot alternates conv1, conv2 (k_step)
  __device__ __forceinline__ int slot_of(uint32_t j, int nslot) const { return j < jobs_full ? (int)(j % nslot) : 0; }
  __device__ __forceinline__ uint32_t conv_of(uint32_t j, int nslot) const {
    return single ? 0u : (j < jobs_full ? (j / nslot) & 1u : (j - jobs_full) & 1u);
  }
};

// The weight tiles of a CTA form one global sequence: tile i = tap i % 9 of conv job i / 9, living in ring slot
// i % kNW. Two warps of the slot whose turn it is drive a conv job, decoupled from each other:
//   * the LEADER (warp 0) only waits for tiles to land (bar_wfull), issues the tcgen05.mma stream of each tap and
//     commits it to bar_wfree - it never waits for a tap to retire, so the tensor core's queue stays fed;
//   * the PRODUCER (warp 1) refills the ring: tile i is requested as soon as tile i - kNW has retired. The producer
//     of job j requests tiles 9j + kWAhead .. 9j + kWAhead + 8 (the first kWAhead tiles of the next job included),
//     so nothing but the job index travels between the slots.
// Both run with warp-uniform arguments (descriptors live in uniform registers); the asynchronous instructions
// themselves are issued by one elected lane.
#ifndef NODE_KWAHEAD
#define NODE_KWAHEAD 3
#endif
constexpr uint32_t kWAhead = NODE_KWAHEAD;      // tiles requested before a job's first tap is issued

template <int NSLOT>
__device__ __forceinline__ void request_tile(const StepSmem& sm, const Jobs& jb, const uint16_t* __restrict__ w16, uint32_t i) {
  const uint32_t slot = i % kNW, cv = jb.conv_of(i / 9, NSLOT), tp = i % 9;
  ptx::mbar_expect_tx(sm.bar_wfull + 8 * slot, kW16TileBytes);
  ptx::bulk_g2s(sm.wring + slot * kW16TileBytes, (const char*)w16 + (size_t)(cv * 9 + tp) * kW16TileBytes, kW16TileBytes,
                sm.bar_wfull + 8 * slot);
}

template <class T, int NSLOT>
__device__ __forceinline__ void produce_conv_job(const StepSmem& sm, const Jobs& jb, const uint16_t* __restrict__ w16, int s, uint32_t job,
                                                 uint32_t nth, bool& timeout) {
  const bool lead = ptx::elect_one();
  // the turn also tells that every tap of the previous job has been issued, hence (bar_wfull gating) that the
  // barrier phases this job waits on cannot alias older ones
  if (!timeout && !ptx::mbar_wait(sm.bar_turn + 8 * s, nth & 1)) timeout = true;
  const uint32_t total = jb.jobs * 9;
#pragma unroll 1
  for (uint32_t i = job * 9 + kWAhead; i < job * 9 + kWAhead + 9 && i < total; ++i) {
    const uint32_t slot = i % kNW;
    if (i >= (uint32_t)kNW && !timeout && !ptx::mbar_wait(sm.bar_wfree + 8 * slot, ((i / kNW) - 1) & 1)) timeout = true;
    if (lead) request_tile<NSLOT>(sm, jb, w16, i);
  }
  __syncwarp();
}

template <class T, int NSLOT>
__device__ __forceinline__ void issue_conv_job(const StepSmem& sm, const Jobs& jb, uint32_t tmem, int s, uint32_t job, uint32_t nth,
                                               bool split, bool& timeout, int mt_used) {
  const bool lead = ptx::elect_one();
  if (!timeout && !ptx::mbar_wait(sm.bar_turn + 8 * s, nth & 1)) timeout = true;   // my slot's nth turn
#ifdef NODE_STEP_DEBUG
  if (lead && blockIdx.x == 0 && nth < 256) g_step_dbg[(s * 256 + nth) * 4 + 1] = clock64();
#endif
  ptx::tc_fence_after();
  // Descriptors as (low word + compile-time constant, constant high word): the start-address field is the low 14 bits
  // (shared memory < 256 KiB, 16-byte units), so a tap, an M tile, a K step or the lo part is ONE 32-bit add on the low
  // word - the leader shares its scheduler with busy worker warps, every instruction it does not issue shortens the job.
  const uint32_t abase = sm.abase + s * 2 * T::A_PART + T::HALO * 16;
  const uint32_t a_lo0 = ((abase & 0x3FFFFu) >> 4) | (((uint32_t)T::LBO >> 4) << 16);
  constexpr uint32_t a_hiw = (128u >> 4) | (1u << 14);                      // SBO = 128 B, descriptor version 1
  constexpr uint32_t b_hiw = (1024u >> 4) | (1u << 14) | (2u << 29);        // SBO = 1024 B, version 1, SWIZZLE_128B
  auto pack = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const uint32_t tile = job * 9 + tap, slot = tile % kNW;
#ifdef NODE_STEP_DEBUG
    const long long q1 = clock64();
#endif
    if (!timeout && !ptx::mbar_wait(sm.bar_wfull + 8 * slot, (tile / kNW) & 1)) timeout = true;
#ifdef NODE_STEP_DEBUG
    if (lead && blockIdx.x == 0 && nth < 256) g_step_dbg2[(s * 256 + nth) * 2 + 0] += clock64() - q1;
#endif
    ptx::tc_fence_after();
    const int off = (tap / 3 - 1) * T::Wp + (tap % 3 - 1);
    const uint32_t a_tap = a_lo0 + (uint32_t)off;                            // never borrows: the image starts HALO rows in
    const uint32_t b_lo0 = ((sm.wring + slot * kW16TileBytes) & 0x3FFFFu) >> 4;
#ifdef NODE_STEP_DEBUG
    const long long q2 = clock64();
#endif
    if (lead) {
#pragma unroll
      for (int mt = 0; mt < T::MT; ++mt) {
        if (mt >= mt_used) continue;           // an M tile that holds no valid image (a batch-1 solve fills one of two): nothing to multiply
        const uint32_t d = tmem + (uint32_t)((s * T::MT + mt) * 128);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t a_hi = pack(a_tap + (uint32_t)((mt * 128 * 16 + 2 * ks * T::LBO) >> 4), a_hiw);
          const uint64_t bk = pack(b_lo0 + (uint32_t)((ks * 32) >> 4), b_hiw);
          const uint32_t first = (tap == 0 && ks == 0) ? 0u : 1u;
          if (split) {
            const uint64_t a_lo = pack(a_tap + (uint32_t)((mt * 128 * 16 + 2 * ks * T::LBO + T::A_PART) >> 4), a_hiw);
            ptx::mma_f16_ss(d, a_hi, bk, kIdF16N128, first);   // a_hi * [w_hi ; w_lo] -> columns [0,64) and [64,128)
            ptx::mma_f16_ss(d, a_lo, bk, kIdF16N64, 1u);       // a_lo * w_hi         -> columns [0,64)
          } else {
            ptx::mma_f16_ss(d, a_hi, bk, kIdF16N64, first);
          }
        }
      }
      ptx::tc_commit(sm.bar_wfree + 8 * slot);
    }
#ifdef NODE_STEP_DEBUG
    if (lead && blockIdx.x == 0) g_step_phase[s * 16 + 13] += clock64() - q2;
#endif
  }
  if (lead) {
    ptx::tc_commit(sm.bar_acc + 8 * s);
    if (job + 1 < jb.jobs) ptx::mbar_arrive(sm.bar_turn + 8 * jb.slot_of(job + 1, NSLOT));
  }
  __syncwarp();
}

// Publish the A image and run the conv job on the tensor core; returns when the accumulators are complete.
template <class T, int NSLOT>
__device__ __forceinline__ void conv_run(const StepSmem& sm, const Who& me, const Jobs& jb, const uint16_t* __restrict__ w16,
                                         uint32_t tmem, uint32_t& njob, uint32_t nfull, bool& timeout, bool split, int mt_used = 1 << 20) {
  const uint32_t job = njob < nfull ? njob * NSLOT + me.slot : jb.jobs_full + (njob - nfull);   // global index of my slot's njob-th job
  ptx::fence_proxy_async();          // my rows of the A image -> visible to the tensor core
  ptx::tc_fence_before();            // my tcgen05.ld of the previous accumulators are done
  slot_sync(me.slot, T::P);
#ifdef NODE_STEP_DEBUG
  const bool rec = blockIdx.x == 0 && me.wt == 0 && njob < 256;
  if (rec) g_step_dbg[(me.slot * 256 + njob) * 4 + 0] = clock64();
#endif
  const int wu = __shfl_sync(0xffffffffu, me.warp, 0);     // shuffles: tell the compiler these values are warp-uniform
  if (wu == 0)
    issue_conv_job<T, NSLOT>(sm, jb, __shfl_sync(0xffffffffu, tmem, 0), __shfl_sync(0xffffffffu, me.slot, 0),
                             __shfl_sync(0xffffffffu, job, 0), __shfl_sync(0xffffffffu, njob, 0), split, timeout,
                             __shfl_sync(0xffffffffu, mt_used, 0));
  else if (wu == 1)
    produce_conv_job<T, NSLOT>(sm, jb, w16, __shfl_sync(0xffffffffu, me.slot, 0), __shfl_sync(0xffffffffu, job, 0),
                               __shfl_sync(0xffffffffu, njob, 0), timeout);
#ifdef NODE_STEP_DEBUG
  if (rec) g_step_dbg[(me.slot * 256 + njob) * 4 + 2] = clock64();
#endif
  if (!timeout && !ptx::mbar_wait_relaxed(sm.bar_acc + 8 * me.slot, njob & 1)) timeout = true;
#ifdef NODE_STEP_DEBUG
  if (rec) g_step_dbg[(me.slot * 256 + njob) * 4 + 3] = clock64();
#endif
  ++njob;
  ptx::tc_fence_after();
}

// x <- acc/scale + (bias + t*Tmap) for output channels [32*hb, 32*hb+32) of this thread's position.
template <class T, bool TB = true>
__device__ __forceinline__ void conv_read(const StepSmem& sm, const Who& me, int hb, float (&x)[32], uint32_t tmem, int cv,
                                          float inv_scale, bool split, bool valid) {
  const uint32_t taddr = tmem + ((uint32_t)((me.warp & 3) * 32) << 16) + (uint32_t)((me.slot * T::MT + (me.wt >> 7)) * 128 + 32 * hb);
  const float4* tb = TB ? sm.tb + (me.slot * 16 + 8 * hb) * 9 + me.cls : nullptr;
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 8) {
    uint32_t v0[8], v1[8];
    ptx::tmem_ld8(taddr + c0, v0);
    if (split) ptx::tmem_ld8(taddr + 64 + c0, v1);
    const float4 e0 = TB ? tb[(c0 >> 2) * 9] : make_float4(0.f, 0.f, 0.f, 0.f), e1 = TB ? tb[((c0 >> 2) + 1) * 9] : make_float4(0.f, 0.f, 0.f, 0.f);
    const float ex[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    ptx::tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float acc = __uint_as_float(v0[j]);
      if (split) acc += __uint_as_float(v1[j]);
      x[c0 + j] = valid ? fmaf(acc, inv_scale, ex[j]) : 0.f;
    }
  }
}

// bias + t*Tmap of convolution `cv` at time t (the slot's threads). Written before the conv job whose epilogue reads it;
// the GroupNorm barriers after the previous epilogue separate it from that epilogue's reads.
template <class T>
__device__ __forceinline__ void make_tb(const StepSmem& sm, const Who& me, int cv, float t) {
  float* tb = reinterpret_cast<float*>(sm.tb) + me.slot * kTbFloats;
  for (int i = me.wt; i < 9 * 64; i += T::P) {
    const int c = i & 63, cls = i >> 6;
    tb[((c >> 2) * 9 + cls) * 4 + (c & 3)] = fmaf(t, sm.tmapc[cv * 9 * 64 + i], sm.bias[cv * 64 + c]);
  }
}

// x = y + sum_j (h*c_j) k_j over NK sources (rk_common.py:49-51), reference rounding and order, for 32 channels
// starting at `p0 = goff + 32*hb*HW`.
template <int HW, int NK>
__device__ __forceinline__ void stage_in(float (&x)[32], const float* __restrict__ y, const float* const (&src)[6],
                                         const float (&hc)[6], float* __restrict__ ynew, size_t p0, bool valid) {
  using A = Arith<float>;
  // Padding threads read image 0's data (p0 = their pixel offset only) and discard it: no divergent loads.
  // Software pipeline over batches of B channels: the loads of batch b+1 are in flight while batch b is
  // combined; the warp-level fences stop the assembler from hoisting every load to the top (register budget).
  // B is sized so that about 48 values per thread are in flight whatever the number of sources.
  constexpr int B = NK <= 2 ? 8 : 4;
  float yv[2][B], kv[2][NK][B];
  auto load = [&](int b, int c0) {
#pragma unroll
    for (int i = 0; i < B; ++i) {
      yv[b][i] = ptx::ldg_ordered(y + p0 + (size_t)(c0 + i) * HW);
#pragma unroll
      for (int j = 0; j < NK; ++j) kv[b][j][i] = ptx::ldg_ordered(src[j] + p0 + (size_t)(c0 + i) * HW);
    }
  };
  load(0, 0);
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += B) {
    const int b = (c0 / B) & 1;
    if (c0 + B < 32) load(b ^ 1, c0 + B);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < B; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NK; ++j) s = A::add(s, A::mul(hc[j], kv[b][j][i]));
      const float r = A::add(yv[b][i], s);
      if (ynew != nullptr && valid) ynew[p0 + (size_t)(c0 + i) * HW] = r;
      x[c0 + i] = valid ? r : 0.f;
    }
    __syncwarp();
  }
}

template <int H_, int W_, int NSLOT>
__global__ void __launch_bounds__(NSLOT * Tile<H_, W_>::P, 1) k_step(const FusedArgs a) {
  using T = Tile<H_, W_>;
  using A = Arith<float>;
  constexpr int HW = T::HW, P = T::P;
  extern __shared__ uint8_t smem_raw[];
  const FusedWs& w = a.w;
  node_ctl_t* ctl = w.ctl;
  const int tid = threadIdx.x;

  if ((a.mode == MODE_STEP || a.mode == MODE_PROBE) && ctl->done) return;   // uniform

  // ---- carve shared memory
  StepSmem sm;
  {
    const uint32_t s0 = ptx::smem_u32(smem_raw);
    const uint32_t al = (s0 + 1023u) & ~1023u;
    uint8_t* base = smem_raw + (al - s0);
    size_t o = 0;
    sm.wring = al; o += (size_t)kNW * kW16TileBytes;
    sm.abase = al + (uint32_t)o; o += (size_t)NSLOT * 2 * T::A_PART;
    sm.part = reinterpret_cast<float*>(base + o); o += (size_t)NSLOT * T::NWARP * 64 * 4;
    sm.stat = reinterpret_cast<float2*>(base + o); o += (size_t)NSLOT * T::G * 32 * 8;
    sm.aff = reinterpret_cast<float4*>(base + o); o += (size_t)NSLOT * T::G * 32 * 16;
    sm.gnp = reinterpret_cast<float4*>(base + o); o += 3 * 32 * 16;
    sm.tb = reinterpret_cast<float4*>(base + o); o += (size_t)NSLOT * kTbFloats * 4;
    sm.bias = reinterpret_cast<float*>(base + o); o += 2 * 64 * 4;
    sm.tmapc = reinterpret_cast<float*>(base + o); o += 2 * 9 * 64 * 4;
    sm.coef = reinterpret_cast<float*>(base + o); o += 64 * 4;
    sm.scratch = reinterpret_cast<double*>(base + o); o += 32 * 8;
    sm.bar_wfull = al + (uint32_t)o; o += 8 * kNW;
    sm.bar_wfree = al + (uint32_t)o; o += 8 * kNW;
    sm.bar_turn = al + (uint32_t)o; o += 8 * 2;
    sm.bar_acc = al + (uint32_t)o; o += 8 * 2;
    sm.ring = reinterpret_cast<volatile uint32_t*>(base + o); o += 16;
    sm.tmem_slot = reinterpret_cast<uint32_t*>(base + o); o += 8;
    sm.illcond = reinterpret_cast<uint32_t*>(base + o);
    // zero the A images once: padding rows / columns are never written again
    uint4* az = reinterpret_cast<uint4*>(base + (size_t)kNW * kW16TileBytes);
    for (int i = tid; i < NSLOT * 2 * T::A_PART / 16; i += blockDim.x) az[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = tid; i < 3 * 32; i += blockDim.x) {
    const int n = i / 32, g = i % 32;
    sm.gnp[i] = make_float4(w.gn[(2 * n) * kC + 2 * g], w.gn[(2 * n) * kC + 2 * g + 1], w.gn[(2 * n + 1) * kC + 2 * g],
                            w.gn[(2 * n + 1) * kC + 2 * g + 1]);
  }
  if (tid < 2) sm.illcond[tid] = 0u;
  for (int i = tid; i < 2 * 64; i += blockDim.x) sm.bias[i] = w.bias[i];
  for (int i = tid; i < 2 * 9 * 64; i += blockDim.x) sm.tmapc[i] = w.tmapc[i];
  const float h = a.mode == MODE_STEP ? ctl->h32 : (a.mode == MODE_PROBE ? ctl->h0_32 : 0.f);
  if (tid < 64) {   // rows 0..5 stage betas, 6 = C_MID, 7 = C_ERR (misc.py:22-25: (h*c)*k)
    const int r = tid >> 3, j = tid & 7;
    double c = 0.0;
    if (r < 7) c = j < 7 ? kCoef(r, j) : 0.0; else c = j < 7 ? kCErr(j) : 0.0;
    sm.coef[tid] = A::mul(h, (float)c);
  }
  if (tid == 0) {
    for (int i = 0; i < kNW; ++i) { ptx::mbar_init(sm.bar_wfull + 8 * i, 1); ptx::mbar_init(sm.bar_wfree + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(sm.bar_turn + 8 * i, 1); ptx::mbar_init(sm.bar_acc + 8 * i, 1); }
    ptx::fence_mbar_init();
    ptx::mbar_arrive(sm.bar_turn);          // the first conv job belongs to slot 0
  }
  if (tid < 32) ptx::tmem_alloc(ptx::smem_u32(sm.tmem_slot), kTmemCols);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *sm.tmem_slot;

  // ---- super-tile schedule: stream u = (cta, slot) takes super-tiles u, u + stride, ...
  const int gs = a.g.gs;                                // images per super-tile for this batch (<= T::G)
  const int NST = (a.g.N + gs - 1) / gs;
  const int stride = gridDim.x * NSLOT;
  const int nevals = a.mode == MODE_STEP ? 6 : 1;
  const bool split = a.conv_mode == CONV_F16X3;
  bool timeout = false;
  double acc0 = 0.0, acc1 = 0.0;
  bool bad = false;
  Jobs jb;
  uint32_t nfull;                 // conv jobs of a slot inside the rounds where every slot works
  {
    int nst[2] = {0, 0};
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int u = blockIdx.x * NSLOT + s;
      nst[s] = u < NST ? (NST - u + stride - 1) / stride : 0;
    }
    nfull = (uint32_t)nst[NSLOT - 1] * nevals * 2;
    jb.jobs_full = nfull * NSLOT;
    jb.jobs = (uint32_t)(nst[0] + (NSLOT > 1 ? nst[1] : 0)) * nevals * 2;
  }
  if (tid == 0)                     // the first tiles of the weight sequence; every later one is requested by a producer warp
    for (uint32_t i = 0; i < kWAhead && i < jb.jobs * 9; ++i) request_tile<NSLOT>(sm, jb, w.w16, i);

  {
    // ===== every thread is a worker: one position of its slot's strip =====
    Who me;
    me.slot = tid / P; me.wt = tid % P; me.warp = me.wt >> 5; me.lane = tid & 31;
    me.img_l = me.wt / T::IS;
    {
      const int r = me.wt % T::IS, hh = r / T::Wp, ww = r % T::Wp;
      me.inimg = me.img_l < T::G && hh < T::H && ww < T::W;
      me.pix = hh * T::W + ww;
      me.cls = (hh == 0 ? 0 : (hh == T::H - 1 ? 2 : 1)) * 3 + (ww == 0 ? 0 : (ww == T::W - 1 ? 2 : 1));
      if (!me.inimg) me.cls = 4;
      const int ia = (me.warp * 32) / T::IS, ib = (me.warp * 32 + 31) / T::IS;
      me.straddle = ia != ib;
      me.isB = me.img_l != ia;
    }
    const int cur = a.mode == MODE_STEP ? ctl->cur : 0;
    const float rtol = (float)ctl->rtol[0], atol = (float)ctl->atol[0];
    uint32_t njob = 0;
    const int u0 = blockIdx.x * NSLOT + me.slot;
#ifdef NODE_STEP_DEBUG
    const bool rec_ph = blockIdx.x == 0 && me.wt == 32 && a.mode == MODE_STEP;
    long long last_ph = clock64();
#endif

#pragma unroll 1
    for (int st = u0; st < NST; st += stride) {
      const int img = st * gs + me.img_l;
      const bool valid = me.inimg && me.img_l < gs && img < a.g.N;
      const size_t goff = valid ? (size_t)img * kC * HW + me.pix : (size_t)(me.inimg ? me.pix : 0);
      const int nvalid = min(gs, a.g.N - st * gs);                          // images of this super-tile (CTA-slot uniform)
      const int mt_used = min(T::MT, (nvalid * T::IS + 127) / 128);         // M tiles that hold one
      float x[32];

#pragma unroll 1
      for (int ev = 0; ev < nevals; ++ev) {
        const float t_state = a.mode == MODE_STEP ? ctl->ts32[ev + 1] : (a.mode == MODE_PROBE ? ctl->ts32[1] : a.t_explicit);
        const float t = a.tsign * t_state;                  // reversed-time wrapper (misc.py:184-187)
        make_tb<T>(sm, me, 0, t);
        NODE_STAMP(0);

        // ---- stage input (rk_common.py:49-51) -> GN1 -> ReLU -> A image of conv1
#pragma unroll 1
        for (int hb = 0; hb < 2; ++hb) {
          const size_t p0 = goff + (size_t)(32 * hb) * HW;
          float* const Ycur = w.Y[cur];
          const float* const Fcur = w.F[cur];
          if (a.mode == MODE_F0 || a.mode == MODE_EVAL) {
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] = ptx::ldg_ordered(a.y_in + p0 + (size_t)c * HW);   // all 32 loads in flight
            if (a.mode == MODE_F0 && valid) {
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                Ycur[p0 + (size_t)c * HW] = x[c];
                if (a.out0 != nullptr) a.out0[p0 + (size_t)c * HW] = x[c];
              }
            }
            if (!valid) {
#pragma unroll
              for (int c = 0; c < 32; ++c) x[c] = 0.f;
            }
          } else {
            // sources in reference order k1, k2, ... with the zero coefficient beta_62 dropped
            float* ynew = (a.mode == MODE_STEP && ev == 5) ? w.Y[cur ^ 1] : nullptr;
            const float* const cf = sm.coef + ev * 8;
            if (a.mode == MODE_PROBE) {                               // y0 + h0*f0 (misc.py:133)
              const float* src[6] = {Fcur, Fcur, Fcur, Fcur, Fcur, Fcur};
              const float hc[6] = {h, 0.f, 0.f, 0.f, 0.f, 0.f};
              stage_in<HW, 1>(x, Ycur, src, hc, nullptr, p0, valid);
            } else if (ev == 0) {
              const float* src[6] = {Fcur, Fcur, Fcur, Fcur, Fcur, Fcur};
              const float hc[6] = {cf[0], 0.f, 0.f, 0.f, 0.f, 0.f};
              stage_in<HW, 1>(x, Ycur, src, hc, ynew, p0, valid);
            } else if (ev == 1) {
              const float* src[6] = {Fcur, w.K[0], Fcur, Fcur, Fcur, Fcur};
              const float hc[6] = {cf[0], cf[1], 0.f, 0.f, 0.f, 0.f};
              stage_in<HW, 2>(x, Ycur, src, hc, ynew, p0, valid);
            } else if (ev == 2) {
              const float* src[6] = {Fcur, w.K[0], w.K[1], Fcur, Fcur, Fcur};
              const float hc[6] = {cf[0], cf[1], cf[2], 0.f, 0.f, 0.f};
              stage_in<HW, 3>(x, Ycur, src, hc, ynew, p0, valid);
            } else if (ev == 3) {
              const float* src[6] = {Fcur, w.K[0], w.K[1], w.K[2], Fcur, Fcur};
              const float hc[6] = {cf[0], cf[1], cf[2], cf[3], 0.f, 0.f};
              stage_in<HW, 4>(x, Ycur, src, hc, ynew, p0, valid);
            } else if (ev == 4) {
              const float* src[6] = {Fcur, w.K[0], w.K[1], w.K[2], w.K[3], Fcur};
              const float hc[6] = {cf[0], cf[1], cf[2], cf[3], cf[4], 0.f};
              stage_in<HW, 5>(x, Ycur, src, hc, ynew, p0, valid);
            } else {
              const float* src[6] = {Fcur, w.K[1], w.K[2], w.K[3], w.K[4], Fcur};
              const float hc[6] = {cf[0], cf[2], cf[3], cf[4], cf[5], 0.f};
              stage_in<HW, 5>(x, Ycur, src, hc, ynew, p0, valid);
            }
          }
          NODE_STAMP(1);
          gn_affine<T>(sm, me, hb, 0, x, valid, a.eps, w.scal[0]);
          NODE_STAMP(2);
          affine_to_A<T>(sm, me, hb, x, valid, split);
          NODE_STAMP(3);
        }
        conv_run<T, NSLOT>(sm, me, jb, w.w16, tmem, njob, nfull, timeout, split, mt_used);
        NODE_STAMP(4);

        // ---- conv1 epilogue -> GN2 -> ReLU -> A image of conv2 (model.py:343-346)
#pragma unroll 1
        for (int hb = 0; hb < 2; ++hb) {
          conv_read<T>(sm, me, hb, x, tmem, 0, w.scal[4], split, valid);
          NODE_STAMP(5);
          gn_affine<T>(sm, me, hb, 1, x, valid, a.eps, w.scal[1]);
          NODE_STAMP(6);
          affine_to_A<T>(sm, me, hb, x, valid, split);
          NODE_STAMP(7);
        }
        make_tb<T>(sm, me, 1, t);
        conv_run<T, NSLOT>(sm, me, jb, w.w16, tmem, njob, nfull, timeout, split, mt_used);
        NODE_STAMP(8);

        // ---- conv2 epilogue -> GN3 -> k_{ev+2} (model.py:346-348), and the norms that feed the controller
#pragma unroll 1
        for (int hb = 0; hb < 2; ++hb) {
          conv_read<T>(sm, me, hb, x, tmem, 1, w.scal[5], split, valid);
          NODE_STAMP(9);
          gn_affine<T>(sm, me, hb, 2, x, valid, a.eps, a.tsign);
          NODE_STAMP(10);
          {
            const float4* af = sm.aff + (me.slot * T::G + min(me.img_l, T::G - 1)) * 32 + 16 * hb;
#pragma unroll
            for (int g = 0; g < 16; ++g) {
              const float4 p = af[g];
              x[2 * g] = fmaf(x[2 * g], p.x, p.z);              // the time sign is folded into p
              x[2 * g + 1] = fmaf(x[2 * g + 1], p.y, p.w);
            }
          }
          if (!valid) continue;
          const size_t p0 = goff + (size_t)(32 * hb) * HW;
          float* kdst = a.mode == MODE_STEP ? (ev < 5 ? w.K[ev] : w.F[cur ^ 1]) : (a.mode == MODE_F0 ? w.F[cur] : (a.mode == MODE_EVAL ? a.k_out : nullptr));
          if (kdst != nullptr) {
#pragma unroll
            for (int c = 0; c < 32; ++c) kdst[p0 + (size_t)c * HW] = x[c];
          }
          NODE_STAMP(11);
          if (a.mode == MODE_F0) {               // misc.py:121-126
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float y = a.y_in[p0 + (size_t)c * HW];
              const float scale = A::add(atol, A::mul(fabsf(y), rtol));
              const float uu = A::div(y, scale), vv = A::div(x[c], scale);
              acc0 += (double)A::mul(uu, uu);
              acc1 += (double)A::mul(vv, vv);
            }
          } else if (a.mode == MODE_PROBE) {     // misc.py:136
            const float* const Ycur = w.Y[cur];
            const float* const Fcur = w.F[cur];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float y = Ycur[p0 + (size_t)c * HW];
              const float scale = A::add(atol, A::mul(fabsf(y), rtol));
              const float uu = A::div(A::sub(x[c], Fcur[p0 + (size_t)c * HW]), scale);
              acc0 += (double)A::mul(uu, uu);
            }
          } else if (a.mode == MODE_STEP && ev == 5) {      // rk_common.py:60, misc.py:146-157, dopri5.py:39-42
            const float* ce = sm.coef + 7 * 8;
            const float* cm = sm.coef + 6 * 8;
            const float* const Ycur = w.Y[cur];
            const float* const Ynew = w.Y[cur ^ 1];
            const float* const Fcur = w.F[cur];
            float part = 0.f;
            // software pipeline over pairs of channels (same scheme as stage_in)
            float ld[2][2][6];
            auto load = [&](int b, int c) {
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const size_t o = p0 + (size_t)(c + i) * HW;
                ld[b][i][0] = ptx::ldg_ordered(Ycur + o); ld[b][i][1] = ptx::ldg_ordered(Ynew + o);
                ld[b][i][2] = ptx::ldg_ordered(Fcur + o); ld[b][i][3] = ptx::ldg_ordered(w.K[1] + o);
                ld[b][i][4] = ptx::ldg_ordered(w.K[2] + o); ld[b][i][5] = ptx::ldg_ordered(w.K[3] + o);
              }
            };
            load(0, 0);
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              const int b = (c >> 1) & 1;
              const float k6a = ptx::ldg_ordered(w.K[4] + p0 + (size_t)c * HW), k6b = ptx::ldg_ordered(w.K[4] + p0 + (size_t)(c + 1) * HW);
              if (c + 2 < 32) load(b ^ 1, c + 2);
              __syncwarp(__activemask());   // padding lanes are not here
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                const size_t o = p0 + (size_t)(c + i) * HW;
                const float y0 = ld[b][i][0], y1 = ld[b][i][1];
                const float kk[7] = {ld[b][i][2], 0.f, ld[b][i][3], ld[b][i][4], ld[b][i][5], i == 0 ? k6a : k6b, x[c + i]};
                float e = 0.f, md = 0.f;
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                  if (j == 1) continue;
                  e = A::add(e, A::mul(ce[j], kk[j]));
                  md = A::add(md, A::mul(cm[j], kk[j]));
                }
                bad |= !isfinite(y0);
                const float tol = A::add(atol, A::mul(rtol, A::max(fabsf(y0), fabsf(y1))));
                const float qv = A::div(e, tol);
                part += A::mul(qv, qv);
                w.YMID[o] = A::add(y0, md);
              }
              __syncwarp(__activemask());   // padding lanes are not here
            }
            acc0 += (double)part;
            NODE_STAMP(12);
          }
        }
      }
    }
  }

  if (a.mode != MODE_EVAL) {
    if (bad) atomicOr(w.nonfinite, 1);
    const double r0 = block_sum(acc0, sm.scratch);
    const double r1 = block_sum(acc1, sm.scratch);
    if (tid == 0) {
      w.partials[blockIdx.x] = r0;
      w.partials[kPartialBlocksF + blockIdx.x] = r1;
    }
  }
  if (timeout) atomicOr(&ctl->status, NODE_ST_WATCHDOG);
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, kTmemCols);
}

// ---- launch ----------------------------------------------------------------------------------------
template <int H_, int W_, int NSLOT>
static int launch_step_shape(const FusedArgs& a, cudaStream_t st) {
  using T = Tile<H_, W_>;
  constexpr size_t smem = step_smem_bytes(T::A_PART, NSLOT, T::NWARP, T::G);
  static_assert(smem <= 227 * 1024, "shared memory budget");
  NODE_SET_SMEM_ONCE((k_step<H_, W_, NSLOT>), smem);
  const int NST = (a.g.N + a.g.gs - 1) / a.g.gs;
  int grid = (NST + NSLOT - 1) / NSLOT;
  if (grid > kMaxGrid) grid = kMaxGrid;
  k_step<H_, W_, NSLOT><<<grid, NSLOT * T::P, smem, st>>>(a);
  return (int)cudaGetLastError();
}

static_assert(step_smem_bytes(Tile<8, 8>::A_PART, 2, Tile<8, 8>::NWARP, Tile<8, 8>::G) <= 227 * 1024,
              "the two-slot variant of the headline shape must fit in shared memory");

template <int H_, int W_>
static int launch_step_slots(const FusedArgs& a, cudaStream_t st) {
  using T = Tile<H_, W_>;
  constexpr bool two = step_smem_bytes(T::A_PART, 2, T::NWARP, T::G) <= 227 * 1024 && 2 * T::MT * 128 <= 512;
  if constexpr (two) {
    const int NST = (a.g.N + a.g.gs - 1) / a.g.gs;
    static const char* force = getenv("NODE_B200_SLOTS");          // "1": always one slot (tuning aid)
    if (NST > kMaxGrid && !(force != nullptr && force[0] == '1')) return launch_step_shape<H_, W_, 2>(a, st);
  }
  return launch_step_shape<H_, W_, 1>(a, st);
}

}  // namespace node

// One translation unit per feature-map shape (they compile in parallel): step_shape_HxW.cu
#define NODE_STEP_SHAPE_TU(H, W) \
  namespace node { int launch_step_##H##x##W(const FusedArgs& a, cudaStream_t st) { return launch_step_slots<H, W>(a, st); } }
#if defined(NODE_STEP_DEBUG) && defined(NODE_STEP_DEBUG_EXPORT)
extern "C" int node_b200_step_debug_read(long long* host, int n) {
  return (int)cudaMemcpyFromSymbol(host, node::g_step_dbg, sizeof(long long) * (n < 2048 ? n : 2048));
}
extern "C" int node_b200_step_phase_read(long long* host, int clear) {
  int rc = (int)cudaMemcpyFromSymbol(host, node::g_step_phase, sizeof(long long) * 32);
  if (clear) { static long long z[32]; rc |= (int)cudaMemcpyToSymbol(node::g_step_phase, z, sizeof(z)); }
  return rc;
}
extern "C" int node_b200_step_debug_read2(long long* host, int n, int clear) {
  int rc = (int)cudaMemcpyFromSymbol(host, node::g_step_dbg2, sizeof(long long) * (n < 1024 ? n : 1024));
  if (clear) { static long long z[1024]; rc |= (int)cudaMemcpyToSymbol(node::g_step_dbg2, z, sizeof(z)); }
  return rc;
}
#endif
