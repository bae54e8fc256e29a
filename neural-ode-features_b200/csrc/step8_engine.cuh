// Dense 8x8 engine of the fused ODE-Net route (reference model.py:326-348 inside rk_common.py:49-52): the headline shape
// (CIFAR `residual`, state [N,64,8,8]). Same contract as k_step (step_engine.cuh): one launch = f0 / initial-step probe /
// all six dopri5 stages of an attempted step / one plain evaluation. What differs is the mapping to the SM:
//
//   * DENSE TILING. A 128-row M tile holds exactly two images: a row-slot is the 8 pixels of one image row plus one zero
//     entry (9 x 16 B = 144 B per k-chunk of 8 channels), the rows of the two images are interleaved (A0 B0 A1 B1 ...) and
//     two zero slots separate the tiles. With the A descriptor's stride-byte-offset = 144 a 3x3 tap (dy, dx) is the byte
//     offset dy*288 + dx*16: no im2col, no per-tap staging, and NO padding rows in M (the strip tiling of step_engine.cuh
//     multiplies 25 % zeros at 8x8). tools/sbo_test.cu checks the layout against a scalar reference on the device.
//   * CTA PAIRS (tcgen05.mma.cta_group::2, cluster of two CTAs on one TPC). One MMA covers 256 rows - the 128-row tile of
//     EACH CTA - and each CTA holds only HALF of the weight rows of a tap (8 KB instead of 16 KB): the shared-memory
//     operand traffic that bound the single-CTA SS-mode stream drops below the math time, the weight ring is 6 taps deep
//     in the same 48 KB, and the L2 -> SM weight stream halves. The pair's leader CTA issues; the peer's workers signal
//     "A image ready" on the leader's mbarrier through DSMEM, a relay warp of the peer forwards "weight half landed",
//     and tcgen05.commit multicasts completion to both CTAs. tools/pair_test.cu is the stand-alone check of this protocol
//     and of the column bookkeeping of the fp16 split under the N split.
//   * ONE SOFTWARE-PIPELINED CHAIN PER CTA instead of two independent slots: all 512 worker threads work on super-tile
//     S0 (4 images, M = 256), publish its A image and move straight on to super-tile S1 while a dedicated warp issues
//     S0's tcgen05.mma stream; the threads only come back to S0 when they have finished S1's phase.
//   * TWO THREADS PER POSITION (32 channels each) where tensor memory is read, and a QUAD mapping (4 pixels x 8 channels
//     per thread, 128-bit global accesses, GroupNorm statistics inside a half warp: no shared memory, no barrier)
//     everywhere else; the conv2 accumulators are transposed into the quad mapping through the idle A image.
// fp32 contract by FP16 operand splitting as in step_engine.cuh (a_hi*[w_hi;w_lo] N=128 + a_lo*w_hi N=64).
#pragma once
#include <cuda_fp16.h>
#include <cstdlib>
#include "step_engine.cuh"

namespace node { namespace s8 {

constexpr int kSlotB = 144;                       // bytes of a row-slot in one k-chunk
constexpr int kChunkSlots = 36;                   // [2 zero][tile 0: 16][2 zero][tile 1: 16]; the next chunk's zeros close it
constexpr int kLBO = kChunkSlots * kSlotB;        // 5184 B between k-chunks
constexpr int kAPart = 8 * kLBO;                  // hi or lo part of one super-tile image
constexpr int kLead = kSlotB, kTail = 2 * kSlotB;
constexpr int kVBytes = kLead + 2 * kAPart + kTail;   // one virtual slot: 83,376 B
constexpr int kRing = 6;                          // weight ring depth (taps)
constexpr int kHalfTile = kW16TileBytes / 2;      // one CTA's half of a tap's weight rows
constexpr int kWorkers = 512, kThreads = kWorkers + 128;   // + one aux warpgroup: warp 16 issues the MMAs, warp 17 streams weights
constexpr int kWorkerRegs = 112, kAuxRegs = 32;            // setmaxnreg moves registers INSIDE the CTA's launch allocation (640 x 96): the aux
                                                           // warpgroup frees 128*(96-32) = 8192, the workers take 512*(112-96) = 8192
constexpr int kImgs = 4;                          // images per super-tile

// Tuning aid (-DNODE_STEP8_DEBUG): CTA 0 accumulates clock64() deltas per phase for worker thread 0 ([0..15]), the
// issuer ([16..23]) and the producer ([24..31]); read back with node_b200_step8_phase_read.
#ifdef NODE_STEP8_DEBUG
static __device__ long long g_s8_phase[32];
// accumulate in a private local array (registers: the indices are literals), flush once at the end
#define S8_STAMP(i) do { if (rec_ph) { const long long now_ = clock64(); ph_acc[(i) & 15] += now_ - last_ph; last_ph = now_; } } while (0)
#define S8_FLUSH(base) do { if (rec_ph) { for (int i_ = 0; i_ < 16; ++i_) g_s8_phase[(base) + i_] += ph_acc[i_]; } } while (0)
static __shared__ long long s8_dbg_shared[16];
#else
#define S8_STAMP(i) do { } while (0)
#define S8_FLUSH(base) do { } while (0)
#endif

struct Smem {
  uint32_t wring, abase;           // shared addresses (ring 1024-aligned)
  float* part;                     // [16 warps][2][32] warp partials of the GroupNorm reductions
  float4* aff;                     // [4 images][32 groups] GN(x) = a*x + b
  float* mean;                     // [4][32] one-pass means (two-pass fallback)
  float4* gnp;                     // [3][32] (gamma0, gamma1, beta0, beta1)
  float4* tm4;                     // [2 conv][16 channel quads][9 border classes]
  float4* bias4;                   // [2][16]
  float* coef;                     // [8][8] h * coefficient
  double* scratch;                 // 32 doubles
  uint32_t* illcond;               // [4] per 128-thread group
  uint32_t bar_wfull, bar_wfree, bar_wpeer, bar_ready, bar_acc;    // wpeer / ready are used in the leader CTA only
  uint32_t* tmem_slot;
};

constexpr size_t smem_bytes() {
  return 1024 + (size_t)kRing * kHalfTile + 2 * (size_t)kVBytes + 16 * 64 * 4 + 4 * 32 * 16 + 4 * 32 * 4 + 3 * 32 * 16 +
         2 * 16 * 9 * 16 + 2 * 16 * 16 + 64 * 4 + 32 * 8 + 16 + 8 * (3 * kRing + 4) + 16;
}
static_assert(smem_bytes() <= 227 * 1024, "shared memory budget");

// Super-tile / conv-job schedule of a CTA, shared by the workers and the aux warp: rounds of (S0, S1); in every round, for
// every evaluation: conv1 of each active virtual slot, then conv2 of each.
// Virtual slot 1 runs `lag` evaluations behind slot 0 (3 of the 6 stages of an attempted step): the k tensors an image
// needs grow from 2 to 7 over the stages of a step, and with every CTA at the same stage the live set of the 1184 images
// in flight (132 MB at the last stage) exceeded the L2 - de-phased, the two slots of a CTA peak at different times (104 MB).
struct Sched {
  int rounds, rounds2;     // super-tiles of virtual slot 0 / of virtual slot 1 (rounds2 <= rounds)
  int nevals, lag;
  int convs = 2;           // conv jobs per evaluation and slot (the adjoint engine, vjp8_engine.cuh, has 4)
  __device__ __forceinline__ uint32_t jobs() const { return (uint32_t)(rounds + rounds2) * (uint32_t)nevals * (uint32_t)convs; }
  __device__ __forceinline__ int iters() const { const int a = rounds * nevals, b = rounds2 > 0 ? rounds2 * nevals + lag : 0; return a > b ? a : b; }
  __device__ __forceinline__ bool active(int it, int v) const {
    return v == 0 ? it < rounds * nevals : (it >= lag && it - lag < rounds2 * nevals);
  }
};

struct JobIter {        // conv jobs in issue order: for every iteration, conv1 of each active slot, then conv2 of each, ...
  int it = 0, cv = 0, v = 0;
  __device__ __forceinline__ void next(const Sched& s) {
    const int n = s.iters();
    do {
      if (++v == 2) { v = 0; if (++cv == s.convs) { cv = 0; ++it; } }
    } while (it < n && !s.active(it, v));
  }
};

__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); }

// Per-thread constants of a worker in the POSITION mapping (accumulator epilogues).
struct Pos {
  int h, mt, wq, lane, grp, warp;      // channel half, M tile, lane quarter, 128-thread group = h*2 + mt
  int il, pix, cls;                    // image within the super-tile, h*8 + w, Tmap border class
  uint32_t arow;                       // byte offset of this position's entry inside a virtual slot (chunk 4h, hi part)
  uint32_t tcol;                       // tensor-memory address offset: lane quarter + mt*128 + 32h
};

// Sum 16 per-lane values over the 16 lanes that share lane bit 3 (one image of the tile): 15 shuffles; the lane with bits
// (b4, b2, b1, b0) returns the total of u[8*b4 + 4*b2 + 2*b1 + b0].
__device__ __forceinline__ float xreduce16_img(const float (&u)[16], int lane) {
  float b[8], c[4], d[2];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? u[i] : u[i + 8], keep = up ? u[i + 8] : u[i];
      b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
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
__device__ __forceinline__ int xreduce16_slot(int lane) {      // (image bit)*16 + group the lane holds after xreduce16_img
  return ((lane >> 3) & 1) * 16 + ((lane >> 4) & 1) * 8 + ((lane >> 2) & 1) * 4 + (lane & 3);
}

// GroupNorm statistics of x (32 channels = 16 groups of this thread's position) in the position mapping: one reduction
// round (sums and sums of squares), folded across the four warps of the thread's 128-thread group; the fold threads publish
// GN(x)*post = a*x + b (sm.aff). A cell whose one-pass variance is ill-conditioned (mean^2 > kGnIllCond var) raises the
// group's flag and the group repeats the statistics with the two-pass scheme of native_group_norm.
__device__ __forceinline__ void gn_affine_pos(const Smem& sm, const Pos& me, int n, const float (&x)[32], float eps, float post) {
  constexpr float inv_n = 1.0f / (float)(kCpg * 64);
  float* part = sm.part + me.warp * 64;
  const int slot = xreduce16_slot(me.lane);
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float u[16];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      u[j] = q == 0 ? x[2 * j] + x[2 * j + 1] : fmaf(x[2 * j], x[2 * j], x[2 * j + 1] * x[2 * j + 1]);
    part[32 * q + slot] = xreduce16_img(u, me.lane);
  }
  group_sync(me.grp);
  const bool folder = me.wq == 0;                  // lane = (image bit)*16 + group
  const int fimg = me.mt * 2 + (me.lane >> 4), fg = me.lane & 15;
  auto publish = [&](float mean, float var) {
    const float rstd = 1.0f / sqrtf(var + eps);
    const float4 p = sm.gnp[n * 32 + 16 * me.h + fg];
    const float a0 = rstd * p.x, a1 = rstd * p.y;
    sm.aff[fimg * 32 + 16 * me.h + fg] = make_float4(a0 * post, a1 * post, (p.z - a0 * mean) * post, (p.w - a1 * mean) * post);
  };
  if (folder) {
    const float* pp = sm.part + (me.warp & ~3) * 64;
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < 4; ++w) { s1 += pp[w * 64 + me.lane]; s2 += pp[w * 64 + 32 + me.lane]; }
    const float mean = s1 * inv_n;
    const float var = fmaxf(fmaf(-mean, mean, s2 * inv_n), 0.f);
    sm.mean[fimg * 32 + 16 * me.h + fg] = mean;
    if (mean * mean > kGnIllCond * var) sm.illcond[me.grp] = 1u;
    publish(mean, var);
  }
  group_sync(me.grp);
  if (*reinterpret_cast<volatile uint32_t*>(sm.illcond + me.grp) != 0u) {      // group-uniform: read after the barrier
    const float* mp = sm.mean + me.il * 32 + 16 * me.h;
    float u[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float m = mp[j];
      const float d0 = x[2 * j] - m, d1 = x[2 * j + 1] - m;
      u[j] = fmaf(d0, d0, d1 * d1);
    }
    part[slot] = xreduce16_img(u, me.lane);
    group_sync(me.grp);
    if (me.wq == 0 && me.lane == 0) sm.illcond[me.grp] = 0u;
    if (folder) {
      const float* pp = sm.part + (me.warp & ~3) * 64;
      float tot = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) tot += pp[w * 64 + me.lane];
      publish(sm.mean[fimg * 32 + 16 * me.h + fg], tot * inv_n);
    }
    group_sync(me.grp);
  }
}

// relu(a*x + b) (operand scale folded into a, b) split into fp16 hi + lo -> this position's entries of k-chunks 4h..4h+3.
__device__ __forceinline__ void affine_to_A_pos(const Smem& sm, const Pos& me, uint32_t vbase, const float (&x)[32], bool split) {
  const float4* af = sm.aff + me.il * 32 + 16 * me.h;
  const uint32_t row = vbase + me.arow;
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = kc * 4 + j;
      const float4 p = af[g];
      const float r0 = fmaxf(fmaf(x[2 * g], p.x, p.z), 0.f);
      const float r1 = fmaxf(fmaf(x[2 * g + 1], p.y, p.w), 0.f);
      const __half2 hh = __floats2half2_rn(r0, r1);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
      hi[j] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[j] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + kc * kLBO), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
    if (split)
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + kAPart + kc * kLBO), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
  }
}

// x <- acc/scale + bias + t*Tmap for output channels [32h, 32h+32) of this thread's position.
__device__ __forceinline__ void conv_read_pos(const Smem& sm, const Pos& me, float (&x)[32], uint32_t tmem, int v, int cv, float inv_scale,
                                              float t, bool split) {
  // columns of a pair accumulator: [hi 0-31 (+ a_lo*w_hi) | lo 32-63 (+ a_lo*w_hi 32-63) | hi 32-63 | lo 0-31]: output channel n is
  // column n + column (n < 32 ? 96 + n : 32 + n); without the split, channel n is column n.
  const uint32_t tbase = tmem + me.tcol + (uint32_t)(v * 256);
  const uint32_t taddr = split ? tbase + (me.h ? 64u : 0u) : tbase + (uint32_t)(32 * me.h);
  const uint32_t taddr2 = tbase + (me.h ? 32u : 96u);
  const float4* tm = sm.tm4 + (cv * 16 + 8 * me.h) * 9 + me.cls;
  const float4* bs = sm.bias4 + cv * 16 + 8 * me.h;
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 8) {
    uint32_t v0[8], v1[8];
    ptx::tmem_ld8(taddr + c0, v0);
    if (split) ptx::tmem_ld8(taddr2 + c0, v1);
    const float4 m0 = tm[(c0 >> 2) * 9], m1 = tm[((c0 >> 2) + 1) * 9], b0 = bs[c0 >> 2], b1 = bs[(c0 >> 2) + 1];
    const float ex[8] = {fmaf(t, m0.x, b0.x), fmaf(t, m0.y, b0.y), fmaf(t, m0.z, b0.z), fmaf(t, m0.w, b0.w),
                         fmaf(t, m1.x, b1.x), fmaf(t, m1.y, b1.y), fmaf(t, m1.z, b1.z), fmaf(t, m1.w, b1.w)};
    ptx::tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float acc = __uint_as_float(v0[j]);
      if (split) acc += __uint_as_float(v1[j]);
      x[c0 + j] = fmaf(acc, inv_scale, ex[j]);
    }
  }
}

// Per-thread constants of a worker in the QUAD mapping (stage combination -> GroupNorm 1 -> A image): the warp owns one
// k-chunk (8 channels) of one M tile (two images); a lane owns 4 consecutive pixels of one image row.
struct Quad {
  int kc, mt, il, lane;          // k-chunk (channels 8kc..8kc+7), M tile, image within the super-tile
  int pix;                       // first pixel: row*8 + 4*hcol
  uint32_t arow;                 // byte offset of the first pixel's entry inside a virtual slot (chunk kc, hi part)
};

// x[c][j] = y + sum_k (h*c_k) k_k for 8 channels x 4 pixels (rk_common.py:49-51), reference rounding and order.
template <int NK>
__device__ __forceinline__ void stage_in_quad(float (&x)[8][4], const float* __restrict__ y, const float* const (&src)[6],
                                              const float (&hc)[6], float* __restrict__ ynew, size_t p0, bool valid, uint64_t pol_ld,
                                              uint64_t pol_st) {
  using A = Arith<float>;
  constexpr int B = NK <= 1 ? 4 : 2;       // channels per batch (one or two GroupNorm groups): 8-12 128-bit loads in flight
#pragma unroll
  for (int c0 = 0; c0 < 8; c0 += B) {
    float4 yv[B], kv[NK][B];
#pragma unroll
    for (int i = 0; i < B; ++i) {
      yv[i] = ptx::ldg128_hint(y + p0 + (size_t)(c0 + i) * 64, pol_ld);
#pragma unroll
      for (int j = 0; j < NK; ++j) kv[j][i] = ptx::ldg128_hint(src[j] + p0 + (size_t)(c0 + i) * 64, pol_ld);
    }
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const float ya[4] = {yv[i].x, yv[i].y, yv[i].z, yv[i].w};
      float r[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < NK; ++j) {
          const float kk = e == 0 ? kv[j][i].x : (e == 1 ? kv[j][i].y : (e == 2 ? kv[j][i].z : kv[j][i].w));
          s = A::add(s, A::mul(hc[j], kk));
        }
        r[e] = A::add(ya[e], s);
        x[c0 + i][e] = valid ? r[e] : 0.f;
      }
      if (ynew != nullptr && valid) ptx::stg128_hint(ynew + p0 + (size_t)(c0 + i) * 64, make_float4(r[0], r[1], r[2], r[3]), pol_st);
    }
  }
}

// Sum 8 per-lane values over the 16 lanes that share lane bit 3; every lane returns all 8 totals of ITS image (16 shuffles).
__device__ __forceinline__ void allreduce8_img(float (&u)[8], int lane) {
  float b[4], c[2];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? u[i] : u[i + 4], keep = up ? u[i + 4] : u[i];
      b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? b[i] : b[i + 2], keep = up ? b[i + 2] : b[i];
      c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  const bool up = lane & 2;
  const float send = up ? c[0] : c[1], keep = up ? c[1] : c[0];
  float d = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  d += __shfl_xor_sync(0xffffffffu, d, 1);
  // the lane with bits (b4, b2, b1) holds the total of u[4*b4 + 2*b2 + b1]
  const int img = lane & 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = __shfl_sync(0xffffffffu, d, img | ((i >> 2) << 4) | (((i >> 1) & 1) << 2) | ((i & 1) << 1));
}

// GroupNorm statistics + affine form of 4 groups (8 channels x 4 pixels per lane, 64 pixels per image inside the warp):
// af[g] = (a0, a1, b0, b1) with GN(x)*post = a*x + b. No shared memory, no barrier.
__device__ __forceinline__ void gn_affine_quad(const Smem& sm, const Quad& me, int n, const float (&x)[8][4], float eps, float post,
                                               float4 (&af)[4]) {
  constexpr float inv_n = 1.0f / (float)(kCpg * 64);
  float u[8];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      s1 += x[2 * g][e] + x[2 * g + 1][e];
      s2 = fmaf(x[2 * g][e], x[2 * g][e], fmaf(x[2 * g + 1][e], x[2 * g + 1][e], s2));
    }
    u[g] = s1; u[4 + g] = s2;
  }
  allreduce8_img(u, me.lane);
  float mean[4], var[4];
  bool ill = false;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    mean[g] = u[g] * inv_n;
    var[g] = fmaxf(fmaf(-mean[g], mean[g], u[4 + g] * inv_n), 0.f);
    ill |= mean[g] * mean[g] > kGnIllCond * var[g];
  }
  if (__any_sync(0xffffffffu, ill)) {              // warp-uniform two-pass variance (native_group_norm's scheme)
    float d[8];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float d0 = x[2 * g][e] - mean[g], d1 = x[2 * g + 1][e] - mean[g];
        s = fmaf(d0, d0, fmaf(d1, d1, s));
      }
      d[g] = s; d[4 + g] = 0.f;
    }
    allreduce8_img(d, me.lane);
#pragma unroll
    for (int g = 0; g < 4; ++g) var[g] = d[g] * inv_n;
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    const float rstd = 1.0f / sqrtf(var[g] + eps);
    const float4 p = sm.gnp[n * 32 + 4 * me.kc + g];
    const float a0 = rstd * p.x, a1 = rstd * p.y;
    af[g] = make_float4(a0 * post, a1 * post, (p.z - a0 * mean[g]) * post, (p.w - a1 * mean[g]) * post);
  }
}

// relu(a*x + b) split into fp16 hi + lo -> the 4 pixel entries (16 B each: the chunk's 8 channels) of this lane.
__device__ __forceinline__ void affine_to_A_quad(const Quad& me, uint32_t vbase, const float (&x)[8][4], const float4 (&af)[4], bool split) {
  const uint32_t row = vbase + me.arow;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float r0 = fmaxf(fmaf(x[2 * g][e], af[g].x, af[g].z), 0.f);
      const float r1 = fmaxf(fmaf(x[2 * g + 1][e], af[g].y, af[g].w), 0.f);
      const __half2 hh = __floats2half2_rn(r0, r1);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
      hi[g] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[g] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + e * 16), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
    if (split)
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + kAPart + e * 16), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
  }
}

__device__ __forceinline__ void request_tile(const Smem& sm, const uint16_t* __restrict__ w16p, uint32_t slot, int cv, int tap, int cta) {
  ptx::mbar_expect_tx(sm.bar_wfull + 8 * slot, kHalfTile);
  ptx::bulk_g2s(sm.wring + slot * kHalfTile, (const char*)w16p + (size_t)((cv * 9 + tap) * 2 + cta) * kHalfTile, kHalfTile,
                sm.bar_wfull + 8 * slot);
}

// The aux warps. Warp 17 of EACH CTA streams that CTA's half tiles (tile i = tap i % 9 of conv job i / 9, ring slot
// i % kRing, requested as soon as tile i - kRing has retired: the retirement is multicast to both CTAs). Warp 16 of the PEER
// relays "my half of tile i has landed" to the leader's mbarrier. Warp 16 of the LEADER waits for a virtual slot's A image
// (32 warp arrivals: both CTAs), for both halves of a tile, and issues the tcgen05.mma.cta_group::2 stream of every conv job
// of the pair in schedule order. Warp-uniform control flow, descriptors in uniform registers, asynchronous instructions
// by one elected lane.
__device__ __forceinline__ void producer_loop(const Smem& sm, const Sched& sc, const uint16_t* __restrict__ w16p, int cta, bool& timeout) {
  const bool lead = ptx::elect_one();
  const uint32_t total = sc.jobs() * 9u;
  JobIter pit;
  int ptap = 0;
#ifdef NODE_STEP8_DEBUG
  const bool rec_ph = blockIdx.x == 0 && lead;
  long long last_ph = clock64();
  long long ph_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
#pragma unroll 1
  for (uint32_t i = 0; i < total; ++i) {
    const uint32_t slot = i % kRing;
    S8_STAMP(8);
    if (i >= (uint32_t)kRing && !timeout && !ptx::mbar_wait(sm.bar_wfree + 8 * slot, ((i / kRing) - 1) & 1)) timeout = true;
    S8_STAMP(9);
    if (lead) request_tile(sm, w16p, slot, pit.cv, ptap, cta);
    if (++ptap == 9) { ptap = 0; pit.next(sc); }
  }
  __syncwarp();
#ifdef NODE_STEP8_DEBUG
  if (rec_ph) { g_s8_phase[24] += ph_acc[8]; g_s8_phase[25] += ph_acc[9]; }
#endif
}

__device__ __forceinline__ void relay_loop(const Smem& sm, const Sched& sc, bool& timeout) {
  // One lane per ring slot, each with its own chain tile -> wait -> signal: a single chain forwards one tile per DSMEM
  // round trip (~1.6k clocks measured), slower than the tensor cores consume them.
  const uint32_t total = sc.jobs() * 9u;
  const uint32_t lane = threadIdx.x & 31u;
  if (lane < (uint32_t)kRing) {
    const uint32_t peer = ptx::mapa(sm.bar_wpeer + 8 * lane, 0);
#pragma unroll 1
    for (uint32_t i = lane; i < total; i += kRing) {
      if (!timeout && !ptx::mbar_wait(sm.bar_wfull + 8 * lane, (i / kRing) & 1)) timeout = true;
      ptx::mbar_arrive_cluster_relaxed(peer);
    }
  }
  __syncwarp();
}

constexpr uint32_t kIdF16N128M256 = (1u << 4) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
constexpr uint32_t kIdF16N64M256 = (1u << 4) | ((64u >> 3) << 17) | ((256u >> 4) << 24);

__device__ __forceinline__ void issuer_loop(const Smem& sm, const Sched& sc, uint32_t tmem, bool split, bool& timeout) {
  const bool lead = ptx::elect_one();
  JobIter it;
  constexpr uint32_t a_hiw = ((uint32_t)kSlotB >> 4) | (1u << 14);           // SBO = 144 B, descriptor version 1
  constexpr uint32_t b_hiw = (1024u >> 4) | (1u << 14) | (2u << 29);          // SBO = 1024 B, version 1, SWIZZLE_128B
  auto pack = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
  uint32_t nready[2] = {0u, 0u};
  uint32_t tile = 0;
  const uint32_t njobs = sc.jobs();
#ifdef NODE_STEP8_DEBUG
  const bool rec_ph = blockIdx.x == 0 && lead;
  long long last_ph = clock64();
  long long ph_acc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif
#pragma unroll 1
  for (uint32_t job = 0; job < njobs; ++job) {
    const int v = it.v;
    S8_STAMP(16);
    if (!timeout && !ptx::mbar_wait_cluster(sm.bar_ready + 8 * v, nready[v] & 1)) timeout = true;
    S8_STAMP(17);
    ++nready[v];
    ptx::tc_fence_after();
    const uint32_t abase = sm.abase + (uint32_t)v * kVBytes + kLead + 2 * kSlotB;       // tile 0, slot 0, chunk 0, hi part
    const uint32_t a_lo0 = ((abase & 0x3FFFFu) >> 4) | (((uint32_t)kLBO >> 4) << 16);
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap, ++tile) {
      const uint32_t slot = tile % kRing;
      S8_STAMP(18);
      if (!timeout && !ptx::mbar_wait(sm.bar_wfull + 8 * slot, (tile / kRing) & 1)) timeout = true;
      S8_STAMP(19);
      if (!timeout && !ptx::mbar_wait_cluster(sm.bar_wpeer + 8 * slot, (tile / kRing) & 1)) timeout = true;
      S8_STAMP(20);
      ptx::tc_fence_after();
      const int off = (tap / 3 - 1) * 2 * kSlotB + (tap % 3 - 1) * 16;
      const uint32_t a_tap = a_lo0 + (uint32_t)(off >> 4);           // arithmetic shift: off may be negative, never borrows
      const uint32_t b_lo0 = ((sm.wring + slot * kHalfTile) & 0x3FFFFu) >> 4;
      if (lead) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t d = tmem + (uint32_t)(v * 256 + mt * 128);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = pack(a_tap + (uint32_t)((mt * 18 * kSlotB + 2 * ks * kLBO) >> 4), a_hiw);
            const uint64_t bk = pack(b_lo0 + (uint32_t)((ks * 32) >> 4), b_hiw);
            const uint32_t first = (tap == 0 && ks == 0) ? 0u : 1u;
            if (split) {
              const uint64_t a_lo = pack(a_tap + (uint32_t)((mt * 18 * kSlotB + 2 * ks * kLBO + kAPart) >> 4), a_hiw);
              ptx::mma2_f16_ss(d, a_hi, bk, kIdF16N128M256, first);
              ptx::mma2_f16_ss(d, a_lo, bk, kIdF16N64M256, 1u);
            } else {
              ptx::mma2_f16_ss(d, a_hi, bk, kIdF16N64M256, first);
            }
          }
        }
        ptx::tc_commit_pair(sm.bar_wfree + 8 * slot, 3);
      }
    }
    if (lead) ptx::tc_commit_pair(sm.bar_acc + 8 * v, 3);
    __syncwarp();
    it.next(sc);
  }
  S8_FLUSH(16);
}

#ifndef NODE_STEP8_HELPERS_ONLY      // vjp8_engine.cuh reuses the helpers above
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_step8(const FusedArgs a) {
  using A = Arith<float>;
  constexpr int HW = 64;
  extern __shared__ uint8_t smem_raw[];
  const FusedWs& w = a.w;
  node_ctl_t* ctl = w.ctl;
  const int tid = threadIdx.x;

  if ((a.mode == MODE_STEP || a.mode == MODE_PROBE) && ctl->done) return;   // uniform over the grid
  const int cta = (int)ptx::cluster_ctarank();

  Smem sm;
  {
    const uint32_t s0 = ptx::smem_u32(smem_raw);
    const uint32_t al = (s0 + 1023u) & ~1023u;
    uint8_t* base = smem_raw + (al - s0);
    size_t o = 0;
    sm.wring = al; o += (size_t)kRing * kHalfTile;
    sm.abase = al + (uint32_t)o;
    uint4* az = reinterpret_cast<uint4*>(base + o);
    o += 2 * (size_t)kVBytes;
    sm.part = reinterpret_cast<float*>(base + o); o += 16 * 64 * 4;
    sm.aff = reinterpret_cast<float4*>(base + o); o += 4 * 32 * 16;
    sm.gnp = reinterpret_cast<float4*>(base + o); o += 3 * 32 * 16;
    sm.tm4 = reinterpret_cast<float4*>(base + o); o += 2 * 16 * 9 * 16;
    sm.bias4 = reinterpret_cast<float4*>(base + o); o += 2 * 16 * 16;
    sm.mean = reinterpret_cast<float*>(base + o); o += 4 * 32 * 4;
    sm.coef = reinterpret_cast<float*>(base + o); o += 64 * 4;
    sm.scratch = reinterpret_cast<double*>(base + o); o += 32 * 8;
    sm.bar_wfull = al + (uint32_t)o; o += 8 * kRing;
    sm.bar_wfree = al + (uint32_t)o; o += 8 * kRing;
    sm.bar_wpeer = al + (uint32_t)o; o += 8 * kRing;
    sm.bar_ready = al + (uint32_t)o; o += 8 * 2;
    sm.bar_acc = al + (uint32_t)o; o += 8 * 2;
    sm.illcond = reinterpret_cast<uint32_t*>(base + o); o += 16;
    sm.tmem_slot = reinterpret_cast<uint32_t*>(base + o);
    // zero the A images once: zero entries / zero slots are never written again
    for (int i = tid; i < 2 * kVBytes / 16; i += kThreads) az[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  for (int i = tid; i < 3 * 32; i += kThreads) {
    const int n = i / 32, g = i % 32;
    sm.gnp[i] = make_float4(w.gn[(2 * n) * kC + 2 * g], w.gn[(2 * n) * kC + 2 * g + 1], w.gn[(2 * n + 1) * kC + 2 * g],
                            w.gn[(2 * n + 1) * kC + 2 * g + 1]);
  }
  if (tid < 4) sm.illcond[tid] = 0u;
  for (int i = tid; i < 2 * 16; i += kThreads) {
    const int cv = i / 16, q = i % 16;
    sm.bias4[i] = make_float4(w.bias[cv * 64 + 4 * q], w.bias[cv * 64 + 4 * q + 1], w.bias[cv * 64 + 4 * q + 2], w.bias[cv * 64 + 4 * q + 3]);
  }
  for (int i = tid; i < 2 * 16 * 9; i += kThreads) {
    const int cls = i % 9, q = (i / 9) % 16, cv = i / (9 * 16);
    const float* tmc = w.tmapc + (cv * 9 + cls) * 64 + 4 * q;
    sm.tm4[i] = make_float4(tmc[0], tmc[1], tmc[2], tmc[3]);
  }
  const float h = a.mode == MODE_STEP ? ctl->h32 : (a.mode == MODE_PROBE ? ctl->h0_32 : 0.f);
  if (tid < 64) {   // rows 0..5 stage betas, 6 = C_MID, 7 = C_ERR (misc.py:22-25: (h*c)*k)
    const int r = tid >> 3, j = tid & 7;
    double c = 0.0;
    if (r < 7) c = j < 7 ? kCoef(r, j) : 0.0; else c = j < 7 ? kCErr(j) : 0.0;
    sm.coef[tid] = A::mul(h, (float)c);
  }
  if (tid == 0) {
    for (int i = 0; i < kRing; ++i) { ptx::mbar_init(sm.bar_wfull + 8 * i, 1); ptx::mbar_init(sm.bar_wfree + 8 * i, 1); ptx::mbar_init(sm.bar_wpeer + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(sm.bar_ready + 8 * i, 32); ptx::mbar_init(sm.bar_acc + 8 * i, 1); }
    ptx::fence_mbar_init();
  }
  if (tid < 32) ptx::tmem_alloc_pair(ptx::smem_u32(sm.tmem_slot), kTmemCols);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();            // both CTAs' barriers exist before anybody arrives through DSMEM
  ptx::tc_fence_after();
  const uint32_t tmem = *sm.tmem_slot;

  // ---- schedule: the pair p takes super-tiles 4p + 2v + cta + r * (2 * grid) for its two virtual slots v, rounds r. Both CTAs of
  // a pair run the SAME job sequence (the MMAs are issued for the pair): a slot is active in a round when the LEADER's
  // super-tile exists; a peer super-tile beyond the batch is computed on zeros and never stored.
  const int NST = (a.g.N + kImgs - 1) / kImgs;
  const int stride = gridDim.x * 2;
  const int unit0 = ((int)blockIdx.x >> 1) * 4;
  Sched sc;
  sc.nevals = a.mode == MODE_STEP ? 6 : 1;
  sc.lag = a.mode == MODE_STEP ? ((a.nw >> 24) & 7) : 0;      // NODE_B200_STEP8_LAG (tuning; default 0)
  sc.rounds = unit0 < NST ? (NST - unit0 + stride - 1) / stride : 0;
  sc.rounds2 = unit0 + 2 < NST ? (NST - (unit0 + 2) + stride - 1) / stride : 0;
  // De-phase the pairs: every CTA runs the same phase sequence, and in lockstep all 148 SMs would hit the L2-bound stage
  // combination (and then the tensor-bound phases) at the same time. A start offset of a fraction of an iteration per pair
  // group spreads the L2 bursts (a.nw >> 1 = offset unit in units of 256 ns; 0 = off).
  if (a.mode == MODE_STEP && ((a.nw >> 1) & 0xFFFFF) != 0) {
    const unsigned grp = (blockIdx.x >> 1) & 3u;
    for (unsigned i = 0; i < grp * (unsigned)((a.nw >> 1) & 0xFFFFF); ++i) __nanosleep(256);
  }
  const bool split = a.conv_mode == CONV_F16X3;
  bool timeout = false;
  double acc0 = 0.0, acc1 = 0.0;
  bool bad = false;

  if (tid >= kWorkers) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kAuxRegs));
    const int aw = __shfl_sync(0xffffffffu, (tid - kWorkers) >> 5, 0);
    if (aw == 0) { if (cta == 0) issuer_loop(sm, sc, tmem, split, timeout); else relay_loop(sm, sc, timeout); }
    else if (aw == 1) producer_loop(sm, sc, w.w16p, cta, timeout);
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kWorkerRegs));
    Pos me;
    me.warp = tid >> 5; me.lane = tid & 31;
    me.h = tid >> 8; me.mt = (tid >> 7) & 1; me.wq = me.warp & 3; me.grp = me.h * 2 + me.mt;
    {
      const int pl = tid & 127, s = pl >> 3, e = pl & 7, row = s >> 1;
      me.il = me.mt * 2 + (s & 1);
      me.pix = row * 8 + e;
      me.cls = (row == 0 ? 0 : (row == 7 ? 2 : 1)) * 3 + (e == 0 ? 0 : (e == 7 ? 2 : 1));
      me.arow = (uint32_t)(kLead + (2 + me.mt * 18 + s) * kSlotB + e * 16 + 4 * me.h * kLBO);
      me.tcol = ((uint32_t)(me.wq * 32) << 16) + (uint32_t)(me.mt * 128);
    }
    Quad qd;
    qd.lane = me.lane; qd.mt = me.mt; qd.kc = 4 * me.h + me.wq;
    {
      const int b3 = (me.lane >> 3) & 1, row = 4 * (me.lane >> 4) + ((me.lane & 7) >> 1), hcol = me.lane & 1;
      qd.il = qd.mt * 2 + b3;
      qd.pix = row * 8 + 4 * hcol;
      qd.arow = (uint32_t)(kLead + qd.kc * kLBO + (2 + qd.mt * 18 + 2 * row + b3) * kSlotB + hcol * 64);
    }
    const int cur = a.mode == MODE_STEP ? ctl->cur : 0;
    const float rtol = (float)ctl->rtol[0], atol = (float)ctl->atol[0];
    // L2 residency of the step's working set (a.nw bit 30 = NODE_B200_STEP8_L2HINT=0 switches the hints off): y, f and the k tensors
    // of the images in flight are kept (evict_last) until the error norm has read them for the last time (evict_first), as are
    // the outputs nothing reads again within this launch.
    const bool hints = a.mode == MODE_STEP && !((a.nw >> 30) & 1);
    const uint64_t pol_keep = hints ? ptx::policy_evict_last() : ptx::policy_evict_normal();
    const uint64_t pol_drop = hints ? ptx::policy_evict_first() : ptx::policy_evict_normal();
    float* const Ycur = w.Y[cur];
    const float* const Fcur = w.F[cur];
    uint32_t nacc[2] = {0u, 0u};
#ifdef NODE_STEP8_DEBUG
    const bool rec_ph = blockIdx.x == 0 && tid == 0 && a.mode == MODE_STEP;
    long long last_ph = clock64();
    long long* const ph_acc = s8_dbg_shared;
    if (tid < 16) s8_dbg_shared[tid] = 0;
    __syncwarp();
#endif
    const uint32_t ready0 = ptx::mapa(sm.bar_ready, 0);          // the LEADER's barrier (DSMEM address; also valid in the leader itself)
    auto publish = [&](int v) {
      ptx::fence_proxy_async();          // my entries of the A image -> visible to the tensor cores of the pair
      ptx::tc_fence_before();            // my tcgen05.ld of the previous accumulators are done
      __syncwarp();
      if (me.lane == 0) ptx::mbar_arrive_cluster(ready0 + 8 * v);
    };
    auto wait_acc = [&](int v) {
      if (!timeout && !ptx::mbar_wait_relaxed(sm.bar_acc + 8 * v, nacc[v] & 1)) timeout = true;
      ++nacc[v];
      ptx::tc_fence_after();
    };

    // L2 prefetch by ONE worker thread, a phase or two ahead of the use: the k tensors of the ~1200 images in flight do
    // not survive in the L2 from one stage of a step to the next (live set ~130 MB, reuse distance one iteration), so without
    // it about half of the stage-combination loads are DRAM round trips inside a latency-bound phase.
    auto prefetch_inputs = [&](int v, int k, bool for_error_norm) {      // inputs of slot v's k-th evaluation
      if (a.mode != MODE_STEP || tid != 0 || k < 0 || (a.nw & 1)) return;      // a.nw bit 0: NODE_B200_STEP8_PREFETCH=0
      const int r = k / sc.nevals, e = k - r * sc.nevals;
      if (r >= (v == 0 ? sc.rounds : sc.rounds2)) return;
      const int img0 = (unit0 + 2 * v + cta + r * stride) * kImgs;
      int n = a.g.N - img0; n = n > kImgs ? kImgs : n;
      if (n <= 0) return;
      const size_t off = (size_t)img0 * kC * HW;
      const uint32_t bytes = (uint32_t)n * kC * HW * 4;
      ptx::bulk_prefetch_l2(Ycur + off, bytes);
      ptx::bulk_prefetch_l2(Fcur + off, bytes);
      if (for_error_norm) {
        for (int j = 1; j < 5; ++j) ptx::bulk_prefetch_l2(w.K[j] + off, bytes);
      } else {
        for (int j = (e == 5 ? 1 : 0); j < e && j < 5; ++j) ptx::bulk_prefetch_l2(w.K[j] + off, bytes);
      }
    };
    const int n_iters = sc.iters();
    constexpr int nv = 2;
#pragma unroll 1
    for (int it = 0; it < n_iters; ++it) {
      {

        // ---- W1: stage input (rk_common.py:49-51) -> GN1 -> ReLU -> A image of conv1, quad mapping
#pragma unroll 1
        for (int v = 0; v < nv; ++v) {
          if (!sc.active(it, v)) continue;
          const int kk_ = it - (v ? sc.lag : 0), r = kk_ / sc.nevals, ev = kk_ - r * sc.nevals;
          const int st = unit0 + 2 * v + cta + r * stride;
          const int img = st * kImgs + qd.il;
          const bool valid = img < a.g.N;
          const size_t p0 = (valid ? (size_t)img * kC * HW : (size_t)0) + (size_t)(8 * qd.kc) * HW + qd.pix;
          if (v == 0) prefetch_inputs(1, it - sc.lag, false);          // slot 1's stage combination comes right after this one
          float x[8][4];
          if (a.mode == MODE_F0 || a.mode == MODE_EVAL) {
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 v4 = ptx::ldg128_ordered(a.y_in + p0 + (size_t)c * HW);
              if (a.mode == MODE_F0 && valid) {
                *reinterpret_cast<float4*>(Ycur + p0 + (size_t)c * HW) = v4;
                if (a.out0 != nullptr) *reinterpret_cast<float4*>(a.out0 + p0 + (size_t)c * HW) = v4;
              }
              x[c][0] = valid ? v4.x : 0.f; x[c][1] = valid ? v4.y : 0.f; x[c][2] = valid ? v4.z : 0.f; x[c][3] = valid ? v4.w : 0.f;
            }
          } else {
            // sources in reference order k1, k2, ... with the zero coefficient beta_62 dropped
            float* ynew = (a.mode == MODE_STEP && ev == 5) ? w.Y[cur ^ 1] : nullptr;
            const float* const cf = sm.coef + ev * 8;
            if (a.mode == MODE_PROBE) {                               // y0 + h0*f0 (misc.py:133)
              const float* src[6] = {Fcur, Fcur, Fcur, Fcur, Fcur, Fcur};
              const float hc[6] = {h, 0.f, 0.f, 0.f, 0.f, 0.f};
              stage_in_quad<1>(x, Ycur, src, hc, nullptr, p0, valid, pol_keep, pol_keep);
            } else if (ev == 0) {
              const float* src[6] = {Fcur, Fcur, Fcur, Fcur, Fcur, Fcur};
              const float hc[6] = {cf[0], 0.f, 0.f, 0.f, 0.f, 0.f};
              stage_in_quad<1>(x, Ycur, src, hc, ynew, p0, valid, pol_keep, pol_keep);
            } else if (ev == 1) {
              const float* src[6] = {Fcur, w.K[0], Fcur, Fcur, Fcur, Fcur};
              const float hc[6] = {cf[0], cf[1], 0.f, 0.f, 0.f, 0.f};
              stage_in_quad<2>(x, Ycur, src, hc, ynew, p0, valid, pol_keep, pol_keep);
            } else if (ev == 2) {
              const float* src[6] = {Fcur, w.K[0], w.K[1], Fcur, Fcur, Fcur};
              const float hc[6] = {cf[0], cf[1], cf[2], 0.f, 0.f, 0.f};
              stage_in_quad<3>(x, Ycur, src, hc, ynew, p0, valid, pol_keep, pol_keep);
            } else if (ev == 3) {
              const float* src[6] = {Fcur, w.K[0], w.K[1], w.K[2], Fcur, Fcur};
              const float hc[6] = {cf[0], cf[1], cf[2], cf[3], 0.f, 0.f};
              stage_in_quad<4>(x, Ycur, src, hc, ynew, p0, valid, pol_keep, pol_keep);
            } else if (ev == 4) {
              const float* src[6] = {Fcur, w.K[0], w.K[1], w.K[2], w.K[3], Fcur};
              const float hc[6] = {cf[0], cf[1], cf[2], cf[3], cf[4], 0.f};
              stage_in_quad<5>(x, Ycur, src, hc, ynew, p0, valid, pol_keep, pol_keep);
            } else {
              const float* src[6] = {Fcur, w.K[1], w.K[2], w.K[3], w.K[4], Fcur};
              const float hc[6] = {cf[0], cf[2], cf[3], cf[4], cf[5], 0.f};
              stage_in_quad<5>(x, Ycur, src, hc, ynew, p0, valid, pol_keep, pol_keep);
            }
          }
          S8_STAMP(0);
          float4 af[4];
          gn_affine_quad(sm, qd, 0, x, a.eps, w.scal[0], af);
          S8_STAMP(1);
          affine_to_A_quad(qd, sm.abase + (uint32_t)v * kVBytes, x, af, split);
          publish(v);
          S8_STAMP(2);
        }

        // ---- W2: conv1 epilogue -> GN2 -> ReLU -> A image of conv2 (model.py:343-346), position mapping
#pragma unroll 1
        for (int v = 0; v < nv; ++v) {
          if (!sc.active(it, v)) continue;
          const int kk_ = it - (v ? sc.lag : 0), r = kk_ / sc.nevals, ev = kk_ - r * sc.nevals;
          const float t = a.tsign * (a.mode == MODE_STEP ? ctl->ts32[ev + 1] : (a.mode == MODE_PROBE ? ctl->ts32[1] : a.t_explicit));   // misc.py:184-187
          S8_STAMP(15);
          wait_acc(v);
          S8_STAMP(3);
          if (ev == 5) prefetch_inputs(v, kk_, true);                  // the error norm re-reads y, k1, k3..k6 two phases from now
          float x[32];
          conv_read_pos(sm, me, x, tmem, v, 0, w.scal[4], t, split);
          S8_STAMP(4);
          gn_affine_pos(sm, me, 1, x, a.eps, w.scal[1]);
          S8_STAMP(5);
          affine_to_A_pos(sm, me, sm.abase + (uint32_t)v * kVBytes, x, split);
          publish(v);
          S8_STAMP(6);
        }

        // ---- W3: conv2 epilogue -> GN3 -> k_{ev+2} (model.py:346-348), and the norms that feed the controller.
        // The accumulators are read in the position mapping (tensor-memory lanes), staged as fp32 through the data entries
        // of the now idle A image (a position's 16 entries of 16 B hold exactly its 64 channels; chunk kc <- channels
        // 8kc..8kc+7, so a quad warp reads back exactly the entries it overwrites in its next W1) and everything else -
        // GroupNorm 3, the k store, error norm / dense-output mid-point - runs in the quad mapping with 128-bit accesses.
#pragma unroll 1
        for (int v = 0; v < nv; ++v) {
          if (!sc.active(it, v)) continue;
          const int kk_ = it - (v ? sc.lag : 0), r = kk_ / sc.nevals, ev = kk_ - r * sc.nevals;
          const float t = a.tsign * (a.mode == MODE_STEP ? ctl->ts32[ev + 1] : (a.mode == MODE_PROBE ? ctl->ts32[1] : a.t_explicit));   // misc.py:184-187
          S8_STAMP(15);
          wait_acc(v);
          S8_STAMP(7);
          if (v == 1 || !sc.active(it, 1)) prefetch_inputs(0, it + 1, false);      // slot 0's next stage combination
          const uint32_t vbase = sm.abase + (uint32_t)v * kVBytes;
          {
            float x[32];
            conv_read_pos(sm, me, x, tmem, v, 1, w.scal[5], t, split);
            const uint32_t row = vbase + me.arow;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row + j * kLBO), "f"(x[8 * j]), "f"(x[8 * j + 1]), "f"(x[8 * j + 2]), "f"(x[8 * j + 3]) : "memory");
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row + kAPart + j * kLBO), "f"(x[8 * j + 4]), "f"(x[8 * j + 5]), "f"(x[8 * j + 6]), "f"(x[8 * j + 7]) : "memory");
            }
          }
          S8_STAMP(8);
          group_sync(me.grp);
          const int st = unit0 + 2 * v + cta + r * stride;
          const int img = st * kImgs + qd.il;
          const bool valid = img < a.g.N;
          float x[8][4];
          {
            const uint32_t row = vbase + qd.arow;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              float4 lo4, hi4;
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(lo4.x), "=f"(lo4.y), "=f"(lo4.z), "=f"(lo4.w) : "r"(row + e * 16));
              asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(hi4.x), "=f"(hi4.y), "=f"(hi4.z), "=f"(hi4.w) : "r"(row + kAPart + e * 16));
              x[0][e] = valid ? lo4.x : 0.f; x[1][e] = valid ? lo4.y : 0.f; x[2][e] = valid ? lo4.z : 0.f; x[3][e] = valid ? lo4.w : 0.f;
              x[4][e] = valid ? hi4.x : 0.f; x[5][e] = valid ? hi4.y : 0.f; x[6][e] = valid ? hi4.z : 0.f; x[7][e] = valid ? hi4.w : 0.f;
            }
          }
          float4 af[4];
          gn_affine_quad(sm, qd, 2, x, a.eps, a.tsign, af);        // the time sign is folded into (a, b)
          S8_STAMP(9);
          if (!valid) continue;
          const size_t p0 = (size_t)img * kC * HW + (size_t)(8 * qd.kc) * HW + qd.pix;
          float* kdst = a.mode == MODE_STEP ? (ev < 5 ? w.K[ev] : w.F[cur ^ 1]) : (a.mode == MODE_F0 ? w.F[cur] : (a.mode == MODE_EVAL ? a.k_out : nullptr));
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 p = af[c >> 1];
            const float aa = (c & 1) ? p.y : p.x, bb = (c & 1) ? p.w : p.z;
#pragma unroll
            for (int e = 0; e < 4; ++e) x[c][e] = fmaf(x[c][e], aa, bb);
            if (kdst != nullptr) ptx::stg128_hint(kdst + p0 + (size_t)c * HW, make_float4(x[c][0], x[c][1], x[c][2], x[c][3]), (a.mode == MODE_STEP && ev < 5) ? pol_keep : pol_drop);
          }
          S8_STAMP(10);
          if (a.mode == MODE_F0) {               // misc.py:121-126
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 y4 = ptx::ldg128_ordered(a.y_in + p0 + (size_t)c * HW);
              const float ya[4] = {y4.x, y4.y, y4.z, y4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float scale = A::add(atol, A::mul(fabsf(ya[e]), rtol));
                const float uu = A::div(ya[e], scale), vv = A::div(x[c][e], scale);
                acc0 += (double)A::mul(uu, uu);
                acc1 += (double)A::mul(vv, vv);
              }
            }
          } else if (a.mode == MODE_PROBE) {     // misc.py:136
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const float4 y4 = ptx::ldg128_ordered(Ycur + p0 + (size_t)c * HW), f4 = ptx::ldg128_ordered(Fcur + p0 + (size_t)c * HW);
              const float ya[4] = {y4.x, y4.y, y4.z, y4.w}, fa[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float scale = A::add(atol, A::mul(fabsf(ya[e]), rtol));
                const float uu = A::div(A::sub(x[c][e], fa[e]), scale);
                acc0 += (double)A::mul(uu, uu);
              }
            }
          } else if (a.mode == MODE_STEP && ev == 5) {      // rk_common.py:60, misc.py:146-157, dopri5.py:39-42
            const float* ce = sm.coef + 7 * 8;
            const float* cm = sm.coef + 6 * 8;
            const float* const Ynew = w.Y[cur ^ 1];
            const float* const srcs[7] = {Ycur, Ynew, Fcur, w.K[1], w.K[2], w.K[3], w.K[4]};
            float part = 0.f;
            float4 ld[2][7];                     // two channels (14 x 128 bits) in flight per batch
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              const int b = c & 1;
              if (b == 0) {
#pragma unroll
                for (int j = 0; j < 7; ++j) { ld[0][j] = ptx::ldg128_hint(srcs[j] + p0 + (size_t)c * HW, pol_drop); ld[1][j] = ptx::ldg128_hint(srcs[j] + p0 + (size_t)(c + 1) * HW, pol_drop); }
              }
              float mid[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                auto el = [&](const float4& q) { return e == 0 ? q.x : (e == 1 ? q.y : (e == 2 ? q.z : q.w)); };
                const float y0 = el(ld[b][0]), y1 = el(ld[b][1]);
                const float kk[7] = {el(ld[b][2]), 0.f, el(ld[b][3]), el(ld[b][4]), el(ld[b][5]), el(ld[b][6]), x[c][e]};
                float er = 0.f, md = 0.f;
#pragma unroll
                for (int j = 0; j < 7; ++j) {
                  if (j == 1) continue;
                  er = A::add(er, A::mul(ce[j], kk[j]));
                  md = A::add(md, A::mul(cm[j], kk[j]));
                }
                bad |= !isfinite(y0);
                const float tol = A::add(atol, A::mul(rtol, A::max(fabsf(y0), fabsf(y1))));
                const float qv = A::div(er, tol);
                part += A::mul(qv, qv);
                mid[e] = A::add(y0, md);
              }
              ptx::stg128_hint(w.YMID + p0 + (size_t)c * HW, make_float4(mid[0], mid[1], mid[2], mid[3]), pol_drop);
            }
            acc0 += (double)part;
            S8_STAMP(11);
          }
        }
      }
    }
    S8_FLUSH(0);
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
  ptx::cluster_sync_all();            // neither CTA's tensor memory / barriers go away while the pair still uses them
  if (tid < 32) ptx::tmem_dealloc_pair(tmem, kTmemCols);
}

static int launch_step8(const FusedArgs& a_in, cudaStream_t st) {
  constexpr size_t smem = smem_bytes();
  FusedArgs a = a_in;
  static const char* pf = getenv("NODE_B200_STEP8_PREFETCH");
  static const char* ph = getenv("NODE_B200_STEP8_DEPHASE");      // start offset per pair group in units of 256 ns (default 0: measured no gain)
  const char* l2 = getenv("NODE_B200_STEP8_L2HINT");
  const char* lg = getenv("NODE_B200_STEP8_LAG");
  a.nw = ((pf != nullptr && pf[0] == '0') ? 1 : 0) | (((ph != nullptr ? atoi(ph) : 0) & 0xFFFFF) << 1) | ((l2 != nullptr && l2[0] == '0') ? (1 << 30) : 0) | (((lg != nullptr ? atoi(lg) : 0) & 7) << 24);
  NODE_SET_SMEM_ONCE(k_step8, smem);
  const int NST = (a.g.N + kImgs - 1) / kImgs;
  int grid = 2 * ((NST + 3) / 4);                  // CTA pairs: 16 images per pair and round
  if (grid > kMaxGrid) grid = kMaxGrid;
  k_step8<<<grid, kThreads, smem, st>>>(a);
  return (int)cudaGetLastError();
}
#endif  // NODE_STEP8_HELPERS_ONLY

}}  // namespace node::s8

#if defined(NODE_STEP8_DEBUG) && !defined(NODE_STEP8_HELPERS_ONLY)
extern "C" int node_b200_step8_phase_read(long long* host, int clear) {
  int rc = (int)cudaMemcpyFromSymbol(host, node::s8::g_s8_phase, sizeof(long long) * 32);
  if (clear) { static long long z[32]; rc |= (int)cudaMemcpyToSymbol(node::s8::g_s8_phase, z, sizeof(z)); }
  return rc;
}
#endif
