// K7a on the dense 8x8 pair engine - the adjoint's augmented dynamics (reference adjoint.py:32-55 applied to model.py:339-348)
// for the headline shape [N,64,8,8]. Same contract as k_vjp (vjp_engine.cuh): per image
//     f = s*ODEfunc(s*t, y),   vjp_y = s*d<f,-a>/dy,   R1, R2, GC1, GC2 (operands of the weight-gradient GEMM, K7b),
//     per-CTA partials of the six GroupNorm affine gradients and of vjp_t.
// What differs is the mapping, which is the one of step8_engine.cuh:
//   * dense tiling (a 128-row M tile = two images, taps as descriptor offsets with SBO = 144), CTA pairs issuing
//     tcgen05.mma.cta_group::2 with each CTA holding half of a tap's weight rows, a dedicated issuer / weight-producer /
//     relay warp, 512 worker threads;
//   * TWO SOFTWARE-PIPELINED SUPER-TILES per CTA: the four conv jobs of an evaluation (conv1, conv2, and the two data-gradient
//     convolutions with flipped, transposed weights) of one super-tile run on the tensor cores while the workers do the
//     GroupNorm forward / backward phase of the other. The one-slot k_vjp had nothing to overlap its MMAs with.
//   * the QUAD mapping (4 pixels x 8 channels per thread) for everything but the tensor-memory reads: 128-bit global
//     accesses for y, a, f, vjp_y, R and GC, GroupNorm statistics and their backward reductions inside a half warp. The
//     accumulators are read in the position mapping and transposed through the idle A image.
//   * c1 stays in tensor memory from conv1 until the backward of GroupNorm 2 has consumed it: every accumulator is 64
//     columns wide (three N = 64 products a_hi*w_hi + a_lo*w_hi + a_hi*w_lo into the same columns), so the c1 and c2 / gradient
//     accumulators of both M tiles of both super-tiles fill the 512 columns exactly.
// Gradient operands are fp16 hi/lo after a power-of-two scale PER IMAGE (gradients have no a-priori bound).
#pragma once
#define NODE_STEP8_HELPERS_ONLY
#include "step8_engine.cuh"

namespace node { namespace v8 {

using s8::kSlotB; using s8::kLBO; using s8::kAPart; using s8::kVBytes; using s8::kLead; using s8::kRing; using s8::kHalfTile;
using s8::kWorkers; using s8::kThreads; using s8::kWorkerRegs; using s8::kAuxRegs; using s8::kImgs;
using s8::Pos; using s8::Quad; using s8::Sched; using s8::JobIter; using s8::group_sync;

constexpr int kW16qSets = 4;                       // conv1, conv2, dgrad of conv1, dgrad of conv2 (FusedWs::w16q)
constexpr uint32_t kColC1 = 0, kColC2 = 64;        // accumulator columns inside an M tile's 128

struct Smem8 {
  s8::Smem s;              // ring, A images, gnp, tm4, bias4, scratch, barriers (part / aff / mean / coef unused)
  float2* stat;            // [2 slots][2 norms][4 images][32 groups] (mean, rstd) of GroupNorm 1, 2 (forward -> backward)
  float* gmx;              // [2 slots][2 uses][2 M tiles][8 warps][2 images] max |g| per warp and image
  float* chacc;            // [16 warps][3 norms][16]: dgamma of the warp's 8 channels, then dbeta
};

constexpr size_t smem_bytes() {
  return 1024 + (size_t)kRing * kHalfTile + 2 * (size_t)kVBytes + 3 * 32 * 16 + 2 * 16 * 9 * 16 + 2 * 16 * 16 + 32 * 8 +
         2 * 2 * 4 * 32 * 8 + 2 * 2 * 2 * 8 * 2 * 4 + 16 * 3 * 16 * 4 + 8 * (3 * kRing + 4) + 16;
}
static_assert(smem_bytes() <= 227 * 1024, "shared memory budget");

// ---- tensor memory -> position mapping -> (A image as fp32 scratch) -> quad mapping -------------------------------------
// x <- acc*mul (+ bias + t*Tmap when cv >= 0) for channels [32h, 32h+32) of this thread's position.
__device__ __forceinline__ void acc_read_pos(const s8::Smem& sm, const Pos& me, float (&x)[32], uint32_t tmem, int v, uint32_t col, int cv,
                                             float mul, float t) {
  const uint32_t taddr = tmem + me.tcol + (uint32_t)(v * 256) + col + (uint32_t)(32 * me.h);
  const int cvi = cv < 0 ? 0 : cv;
  const float4* tm = sm.tm4 + (cvi * 16 + 8 * me.h) * 9 + me.cls;
  const float4* bs = sm.bias4 + cvi * 16 + 8 * me.h;
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 16) {
    uint32_t r[16];
    ptx::tmem_ld16(taddr + c0, r);
    float ex[16];
    if (cv >= 0) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 m = tm[((c0 >> 2) + q) * 9], b = bs[(c0 >> 2) + q];
        ex[4 * q] = fmaf(t, m.x, b.x); ex[4 * q + 1] = fmaf(t, m.y, b.y); ex[4 * q + 2] = fmaf(t, m.z, b.z); ex[4 * q + 3] = fmaf(t, m.w, b.w);
      }
    }
    ptx::tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) x[c0 + j] = cv >= 0 ? fmaf(__uint_as_float(r[j]), mul, ex[j]) : __uint_as_float(r[j]) * mul;
  }
}

// a position's 16 entries of 16 B hold exactly its 64 channels as fp32: chunk kc <- channels 8kc..8kc+3 (hi part), 8kc+4..8kc+7 (lo part)
__device__ __forceinline__ void stage_pos(const Pos& me, uint32_t vbase, const float (&x)[32]) {
  const uint32_t row = vbase + me.arow;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row + j * kLBO), "f"(x[8 * j]), "f"(x[8 * j + 1]), "f"(x[8 * j + 2]), "f"(x[8 * j + 3]) : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row + kAPart + j * kLBO), "f"(x[8 * j + 4]), "f"(x[8 * j + 5]), "f"(x[8 * j + 6]), "f"(x[8 * j + 7]) : "memory");
  }
}
__device__ __forceinline__ void unstage_quad(const Quad& qd, uint32_t vbase, float (&x)[8][4], float mul) {
  const uint32_t row = vbase + qd.arow;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float4 lo4, hi4;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(lo4.x), "=f"(lo4.y), "=f"(lo4.z), "=f"(lo4.w) : "r"(row + e * 16));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(hi4.x), "=f"(hi4.y), "=f"(hi4.z), "=f"(hi4.w) : "r"(row + kAPart + e * 16));
    x[0][e] = lo4.x * mul; x[1][e] = lo4.y * mul; x[2][e] = lo4.z * mul; x[3][e] = lo4.w * mul;
    x[4][e] = hi4.x * mul; x[5][e] = hi4.y * mul; x[6][e] = hi4.z * mul; x[7][e] = hi4.w * mul;
  }
}

// ---- GroupNorm in the quad mapping ---------------------------------------------------------------------------------------
// (mean, rstd) of the 4 groups of this warp's k-chunk for the lane's image (one-pass moments, warp-uniform two-pass fallback as
// in s8::gn_affine_quad / native_group_norm).
__device__ __forceinline__ void gn_stats_quad(const Quad& me, const float (&x)[8][4], float eps, float (&mean)[4], float (&rstd)[4]) {
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
  s8::allreduce8_img(u, me.lane);
  float var[4];
  bool ill = false;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    mean[g] = u[g] * inv_n;
    var[g] = fmaxf(fmaf(-mean[g], mean[g], u[4 + g] * inv_n), 0.f);
    ill |= mean[g] * mean[g] > kGnIllCond * var[g];
  }
  if (__any_sync(0xffffffffu, ill)) {
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
    s8::allreduce8_img(d, me.lane);
#pragma unroll
    for (int g = 0; g < 4; ++g) var[g] = d[g] * inv_n;
  }
#pragma unroll
  for (int g = 0; g < 4; ++g) rstd[g] = 1.0f / sqrtf(var[g] + eps);
}

// GN(x)*post = a*x + b per channel pair of group g: (a0, a1, b0, b1)
__device__ __forceinline__ float4 gn_aff(float mean, float rstd, const float4 p, float post) {
  const float a0 = rstd * p.x, a1 = rstd * p.y;
  return make_float4(a0 * post, a1 * post, (p.z - a0 * mean) * post, (p.w - a1 * mean) * post);
}

__device__ __forceinline__ void stat_store(const Smem8& sm, const Quad& me, int v, int n, const float (&mean)[4], const float (&rstd)[4]) {
  if ((me.lane & 0x17) == 0) {                  // one lane per image of the warp's M tile (lanes 0 and 8)
    float2* dst = sm.stat + ((v * 2 + n) * 4 + me.il) * 32 + 4 * me.kc;
#pragma unroll
    for (int g = 0; g < 4; ++g) dst[g] = make_float2(mean[g], rstd[g]);
  }
}
__device__ __forceinline__ void stat_load(const Smem8& sm, const Quad& me, int v, int n, float (&mean)[4], float (&rstd)[4]) {
  const float2* src = sm.stat + ((v * 2 + n) * 4 + me.il) * 32 + 4 * me.kc;
#pragma unroll
  for (int g = 0; g < 4; ++g) { const float2 s = src[g]; mean[g] = s.x; rstd[g] = s.y; }
}

// Sum 16 per-lane values over the whole warp: the lane with bits (b4, b3, b2, b1) returns the total of u[8 b4 + 4 b3 + 2 b2 + b1].
__device__ __forceinline__ float xreduce16_warp(const float (&u)[16], int lane) {
  float b[8], c[4], d[2];
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float send = up ? u[i] : u[i + 8], keep = up ? u[i + 8] : u[i]; b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16); }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float send = up ? b[i] : b[i + 4], keep = up ? b[i + 4] : b[i]; c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) { const float send = up ? c[i] : c[i + 2], keep = up ? c[i + 2] : c[i]; d[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
  }
  const bool up = lane & 2;
  const float send = up ? d[0] : d[1], keep = up ? d[1] : d[0];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

// Backward of one GroupNorm (oracle/odefunc_port.py:_gn_bwd) in the quad mapping. in: x = the norm's input, g = gradient at its
// output (already masked by the ReLU that follows, if any); out: g <- gradient at the norm's input. Accumulates dgamma / dbeta.
__device__ __forceinline__ void gn_backward_quad(const Smem8& sm, const Quad& me, int warp, int n, const float (&x)[8][4], float (&g)[8][4],
                                                 const float (&mean)[4], const float (&rstd)[4]) {
  constexpr float inv_m = 1.0f / (float)(kCpg * 64);
  float u[16], s[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float dg = 0.f, db = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float xh = (x[c][e] - mean[c >> 1]) * rstd[c >> 1];
      dg = fmaf(g[c][e], xh, dg);
      db += g[c][e];
    }
    u[c] = dg; u[8 + c] = db;
  }
  float4 p[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    p[q] = sm.s.gnp[n * 32 + 4 * me.kc + q];
    s[q] = u[8 + 2 * q] * p[q].x + u[8 + 2 * q + 1] * p[q].y;          // S1 = sum_cell g*gamma
    s[4 + q] = u[2 * q] * p[q].x + u[2 * q + 1] * p[q].y;              // S2 = sum_cell g*gamma*xhat
  }
  {
    const float r = xreduce16_warp(u, me.lane);
    if ((me.lane & 1) == 0) sm.chacc[(warp * 3 + n) * 16 + (me.lane >> 1)] += r;
  }
  s8::allreduce8_img(s, me.lane);
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int q = c >> 1;
    const float gam = (c & 1) ? p[q].y : p[q].x;
    const float m1 = s[q] * inv_m, m2 = s[4 + q] * inv_m;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float xh = (x[c][e] - mean[q]) * rstd[q];
      g[c][e] = rstd[q] * (g[c][e] * gam - m1 - xh * m2);
    }
  }
}

// relu(a*x + b) (the operand scale folded into a, b): unscaled value -> r_out (NCHW fp32, operand of the weight gradient), fp16 hi/lo
// -> the lane's 4 pixel entries of the A image. mask = false: no ReLU (never used for activations).
__device__ __forceinline__ void act_to_A_quad(const Quad& me, uint32_t vbase, const float (&x)[8][4], const float4 (&af)[4], float inv_post,
                                              float* __restrict__ r_out, size_t p0, bool valid) {
  const uint32_t row = vbase + me.arow;
  float r[8][4];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 a = af[c >> 1];
    const float aa = (c & 1) ? a.y : a.x, bb = (c & 1) ? a.w : a.z;
#pragma unroll
    for (int e = 0; e < 4; ++e) r[c][e] = valid ? fmaxf(fmaf(x[c][e], aa, bb), 0.f) : 0.f;
    if (valid) *reinterpret_cast<float4*>(r_out + p0 + (size_t)c * 64) = make_float4(r[c][0] * inv_post, r[c][1] * inv_post, r[c][2] * inv_post, r[c][3] * inv_post);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __half2 hh = __floats2half2_rn(r[2 * q][e], r[2 * q + 1][e]);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(r[2 * q][e] - hf.x, r[2 * q + 1][e] - hf.y);
      hi[q] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[q] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + e * 16), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + kAPart + e * 16), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
  }
}

// g*scale split into fp16 hi/lo -> the lane's entries of the A image (gradient operand of a data-gradient convolution)
__device__ __forceinline__ void grad_to_A_quad(const Quad& me, uint32_t vbase, const float (&g)[8][4], float scale) {
  const uint32_t row = vbase + me.arow;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float v0 = g[2 * q][e] * scale, v1 = g[2 * q + 1][e] * scale;
      const __half2 hh = __floats2half2_rn(v0, v1);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
      hi[q] = *reinterpret_cast<const uint32_t*>(&hh);
      lo[q] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + e * 16), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + kAPart + e * 16), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
  }
}

// Power-of-two scale that brings max|g| of the lane's IMAGE just below 2^14. The per-warp maxima of use `use` (0: GC2, 1: GC1) of
// slot v stay in shared memory: the phase that reads the data-gradient accumulators back derives the same scale from them.
__device__ __forceinline__ float* gmx_cell(const Smem8& sm, int v, int use, int mt) { return sm.gmx + ((v * 2 + use) * 2 + mt) * 16; }
__device__ __forceinline__ float scale_from(const float* cell, int b3, bool inverse) {
  float m = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) m = fmaxf(m, cell[w * 2 + b3]);
  int e = 0;
  if (m > 0.f && m < 3.0e38f) {
    int ex;
    (void)frexpf(m, &ex);                         // m = f * 2^ex, f in [0.5, 1)
    e = 14 - ex;
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
  }
  return exp2f((float)(inverse ? -e : e));
}
__device__ __forceinline__ float grad_scale_img(const Smem8& sm, const Quad& me, int v, int use, const float (&g)[8][4], float& run_max) {
  float m = 0.f;
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int e = 0; e < 4; ++e) m = fmaxf(m, fabsf(g[c][e]));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 16));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
  m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
  run_max = fmaxf(run_max, m);
  float* cell = gmx_cell(sm, v, use, me.mt);
  const int b3 = (me.lane >> 3) & 1;
  if ((me.lane & 0x17) == 0) cell[me.kc * 2 + b3] = m;
  asm volatile("bar.sync %0, 256;" ::"r"(5 + me.mt) : "memory");       // the 8 warps of this M tile
  return scale_from(cell, b3, false);
}

// ---- aux warps ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int set_of(int cv) { return cv == 2 ? 3 : (cv == 3 ? 2 : cv); }    // conv1, conv2, dgrad(conv2), dgrad(conv1)

__device__ __forceinline__ void producer_loop(const s8::Smem& sm, const Sched& sc, const uint16_t* __restrict__ w16q, int cta, bool& timeout) {
  const bool lead = ptx::elect_one();
  const uint32_t total = sc.jobs() * 9u;
  JobIter pit;
  int ptap = 0;
#pragma unroll 1
  for (uint32_t i = 0; i < total; ++i) {
    const uint32_t slot = i % kRing;
    if (i >= (uint32_t)kRing && !timeout && !ptx::mbar_wait(sm.bar_wfree + 8 * slot, ((i / kRing) - 1) & 1)) timeout = true;
    if (lead) s8::request_tile(sm, w16q, slot, set_of(pit.cv), ptap, cta);
    if (++ptap == 9) { ptap = 0; pit.next(sc); }
  }
  __syncwarp();
}

constexpr uint32_t kIdF16N64M256 = (1u << 4) | ((64u >> 3) << 17) | ((256u >> 4) << 24);

__device__ __forceinline__ void issuer_loop(const s8::Smem& sm, const Sched& sc, uint32_t tmem, bool& timeout) {
  const bool lead = ptx::elect_one();
  JobIter it;
  constexpr uint32_t a_hiw = ((uint32_t)kSlotB >> 4) | (1u << 14);           // SBO = 144 B, descriptor version 1
  constexpr uint32_t b_hiw = (1024u >> 4) | (1u << 14) | (2u << 29);          // SBO = 1024 B, version 1, SWIZZLE_128B
  auto pack = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
  uint32_t tile = 0;
  const uint32_t njobs = sc.jobs();
#pragma unroll 1
  for (uint32_t job = 0; job < njobs; ++job) {
    const int v = it.v, cv = it.cv;
    // a slot's A image is published four times per super-tile (once per conv job): the phase parity is the job's parity
    if (!timeout && !ptx::mbar_wait_cluster(sm.bar_ready + 8 * v, (uint32_t)cv & 1u)) timeout = true;
    ptx::tc_fence_after();
    const uint32_t abase = sm.abase + (uint32_t)v * kVBytes + kLead + 2 * kSlotB;       // tile 0, slot 0, chunk 0, hi part
    const uint32_t a_lo0 = ((abase & 0x3FFFFu) >> 4) | (((uint32_t)kLBO >> 4) << 16);
    const uint32_t dcol = (uint32_t)(v * 256) + (cv == 0 ? kColC1 : kColC2);
#pragma unroll 1
    for (int tap = 0; tap < 9; ++tap, ++tile) {
      const uint32_t slot = tile % kRing;
      if (!timeout && !ptx::mbar_wait(sm.bar_wfull + 8 * slot, (tile / kRing) & 1)) timeout = true;
      if (!timeout && !ptx::mbar_wait_cluster(sm.bar_wpeer + 8 * slot, (tile / kRing) & 1)) timeout = true;
      ptx::tc_fence_after();
      const int off = (tap / 3 - 1) * 2 * kSlotB + (tap % 3 - 1) * 16;
      const uint32_t a_tap = a_lo0 + (uint32_t)(off >> 4);
      const uint32_t b_lo0 = ((sm.wring + slot * kHalfTile) & 0x3FFFFu) >> 4;
      if (lead) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          const uint32_t d = tmem + dcol + (uint32_t)(mt * 128);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t a_hi = pack(a_tap + (uint32_t)((mt * 18 * kSlotB + 2 * ks * kLBO) >> 4), a_hiw);
            const uint64_t a_lo = pack(a_tap + (uint32_t)((mt * 18 * kSlotB + 2 * ks * kLBO + kAPart) >> 4), a_hiw);
            const uint64_t b_hi = pack(b_lo0 + (uint32_t)((ks * 32) >> 4), b_hiw);                 // rows 0..31 of each CTA's half tile: w_hi
            const uint64_t b_lo = pack(b_lo0 + (uint32_t)((32 * 128 + ks * 32) >> 4), b_hiw);      // rows 32..63: w_lo
            ptx::mma2_f16_ss(d, a_hi, b_hi, kIdF16N64M256, (tap == 0 && ks == 0) ? 0u : 1u);
            ptx::mma2_f16_ss(d, a_lo, b_hi, kIdF16N64M256, 1u);
            ptx::mma2_f16_ss(d, a_hi, b_lo, kIdF16N64M256, 1u);
          }
        }
        ptx::tc_commit_pair(sm.bar_wfree + 8 * slot, 3);
      }
    }
    if (lead) ptx::tc_commit_pair(sm.bar_acc + 8 * v, 3);
    __syncwarp();
    it.next(sc);
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) k_vjp8(const VjpArgs a) {
  constexpr int HW = 64;
  extern __shared__ uint8_t smem_raw[];
  const FusedWs& w = a.w;
  const int tid = threadIdx.x;
  const int cta = (int)ptx::cluster_ctarank();

  Smem8 sm;
  {
    const uint32_t s0 = ptx::smem_u32(smem_raw);
    const uint32_t al = (s0 + 1023u) & ~1023u;
    uint8_t* base = smem_raw + (al - s0);
    size_t o = 0;
    sm.s.wring = al; o += (size_t)kRing * kHalfTile;
    sm.s.abase = al + (uint32_t)o;
    uint4* az = reinterpret_cast<uint4*>(base + o);
    o += 2 * (size_t)kVBytes;
    sm.s.gnp = reinterpret_cast<float4*>(base + o); o += 3 * 32 * 16;
    sm.s.tm4 = reinterpret_cast<float4*>(base + o); o += 2 * 16 * 9 * 16;
    sm.s.bias4 = reinterpret_cast<float4*>(base + o); o += 2 * 16 * 16;
    sm.s.scratch = reinterpret_cast<double*>(base + o); o += 32 * 8;
    sm.stat = reinterpret_cast<float2*>(base + o); o += 2 * 2 * 4 * 32 * 8;
    sm.gmx = reinterpret_cast<float*>(base + o); o += 2 * 2 * 2 * 8 * 2 * 4;
    sm.chacc = reinterpret_cast<float*>(base + o); o += 16 * 3 * 16 * 4;
    sm.s.bar_wfull = al + (uint32_t)o; o += 8 * kRing;
    sm.s.bar_wfree = al + (uint32_t)o; o += 8 * kRing;
    sm.s.bar_wpeer = al + (uint32_t)o; o += 8 * kRing;
    sm.s.bar_ready = al + (uint32_t)o; o += 8 * 2;
    sm.s.bar_acc = al + (uint32_t)o; o += 8 * 2;
    sm.s.tmem_slot = reinterpret_cast<uint32_t*>(base + o);
    sm.s.part = nullptr; sm.s.aff = nullptr; sm.s.mean = nullptr; sm.s.coef = nullptr; sm.s.illcond = nullptr;
    for (int i = tid; i < 2 * kVBytes / 16; i += kThreads) az[i] = make_uint4(0u, 0u, 0u, 0u);     // zero entries / slots are never written again
    for (int i = tid; i < 16 * 3 * 16; i += kThreads) sm.chacc[i] = 0.f;
    for (int i = tid; i < 2 * 2 * 2 * 8 * 2; i += kThreads) sm.gmx[i] = 0.f;
  }
  for (int i = tid; i < 3 * 32; i += kThreads) {
    const int n = i / 32, g = i % 32;
    sm.s.gnp[i] = make_float4(w.gn[(2 * n) * kC + 2 * g], w.gn[(2 * n) * kC + 2 * g + 1], w.gn[(2 * n + 1) * kC + 2 * g],
                              w.gn[(2 * n + 1) * kC + 2 * g + 1]);
  }
  for (int i = tid; i < 2 * 16; i += kThreads) {
    const int cv = i / 16, q = i % 16;
    sm.s.bias4[i] = make_float4(w.bias[cv * 64 + 4 * q], w.bias[cv * 64 + 4 * q + 1], w.bias[cv * 64 + 4 * q + 2], w.bias[cv * 64 + 4 * q + 3]);
  }
  for (int i = tid; i < 2 * 16 * 9; i += kThreads) {
    const int cls = i % 9, q = (i / 9) % 16, cv = i / (9 * 16);
    const float* tmc = w.tmapc + (cv * 9 + cls) * 64 + 4 * q;
    sm.s.tm4[i] = make_float4(tmc[0], tmc[1], tmc[2], tmc[3]);
  }
  if (tid == 0) {
    for (int i = 0; i < kRing; ++i) { ptx::mbar_init(sm.s.bar_wfull + 8 * i, 1); ptx::mbar_init(sm.s.bar_wfree + 8 * i, 1); ptx::mbar_init(sm.s.bar_wpeer + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(sm.s.bar_ready + 8 * i, 32); ptx::mbar_init(sm.s.bar_acc + 8 * i, 1); }
    ptx::fence_mbar_init();
  }
  if (tid < 32) ptx::tmem_alloc_pair(ptx::smem_u32(sm.s.tmem_slot), kTmemCols);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem = *sm.s.tmem_slot;

  // schedule (as k_step8): pair p takes super-tiles 4p + 2v + cta + r*(2*grid); both CTAs of a pair run the same job sequence
  const int NST = (a.g.N + kImgs - 1) / kImgs;
  const int stride = gridDim.x * 2;
  const int unit0 = ((int)blockIdx.x >> 1) * 4;
  Sched sc;
  sc.nevals = 1; sc.lag = 0; sc.convs = 4;
  sc.rounds = unit0 < NST ? (NST - unit0 + stride - 1) / stride : 0;
  sc.rounds2 = unit0 + 2 < NST ? (NST - (unit0 + 2) + stride - 1) / stride : 0;
  bool timeout = false;
  float tacc = 0.f;
  float gc_max[2] = {0.f, 0.f};

  if (tid >= kWorkers) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kAuxRegs));
    const int aw = __shfl_sync(0xffffffffu, (tid - kWorkers) >> 5, 0);
    if (aw == 0) { if (cta == 0) v8::issuer_loop(sm.s, sc, tmem, timeout); else s8::relay_loop(sm.s, sc, timeout); }
    else if (aw == 1) v8::producer_loop(sm.s, sc, w.w16q, cta, timeout);
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
    int qcls[4];                                  // Tmap border class of the lane's 4 pixels
    {
      const int b3 = (me.lane >> 3) & 1, row = 4 * (me.lane >> 4) + ((me.lane & 7) >> 1), hcol = me.lane & 1;
      qd.il = qd.mt * 2 + b3;
      qd.pix = row * 8 + 4 * hcol;
      qd.arow = (uint32_t)(kLead + qd.kc * kLBO + (2 + qd.mt * 18 + 2 * row + b3) * kSlotB + hcol * 64);
      const int rc = row == 0 ? 0 : (row == 7 ? 2 : 1);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const int col = 4 * hcol + e; qcls[e] = rc * 3 + (col == 0 ? 0 : (col == 7 ? 2 : 1)); }
    }
    const float t = a.tsign * a.t_dev[0];         // reversed-time wrapper (misc.py:184-187)
    const float inv_sw1 = 1.0f / w.scal[2], inv_sw2 = 1.0f / w.scal[3];
    const float post1 = w.scal[0], post2 = w.scal[1];
    const uint32_t ready0 = ptx::mapa(sm.s.bar_ready, 0);
    auto publish = [&](int v) {
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (me.lane == 0) ptx::mbar_arrive_cluster(ready0 + 8 * v);
    };
    auto wait_acc = [&](int v, int k) {           // k-th conv job of the slot's current super-tile
      if (!timeout && !ptx::mbar_wait_relaxed(sm.s.bar_acc + 8 * v, (uint32_t)k & 1u)) timeout = true;
      ptx::tc_fence_after();
    };
    // sum g * Tmap(conv cv) over the lane's 8 channels x 4 pixels (vjp_t)
    auto time_grad = [&](int cv, const float (&g)[8][4]) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 m0 = sm.s.tm4[(cv * 16 + 2 * qd.kc) * 9 + qcls[e]], m1 = sm.s.tm4[(cv * 16 + 2 * qd.kc + 1) * 9 + qcls[e]];
        s = fmaf(g[0][e], m0.x, s); s = fmaf(g[1][e], m0.y, s); s = fmaf(g[2][e], m0.z, s); s = fmaf(g[3][e], m0.w, s);
        s = fmaf(g[4][e], m1.x, s); s = fmaf(g[5][e], m1.y, s); s = fmaf(g[6][e], m1.z, s); s = fmaf(g[7][e], m1.w, s);
      }
      tacc += s;
    };
    const int b3 = (me.lane >> 3) & 1;

#pragma unroll 1
    for (int r = 0; r < sc.rounds; ++r) {
      const int nv = r < sc.rounds2 ? 2 : 1;
      // ---- P1: y -> GN1 -> ReLU -> R1, A image of conv1 (model.py:341-343)
#pragma unroll 1
      for (int v = 0; v < nv; ++v) {
        const int img = (unit0 + 2 * v + cta + r * stride) * kImgs + qd.il;
        const bool valid = img < a.g.N;
        const size_t p0 = (valid ? (size_t)img * kC * HW : (size_t)0) + (size_t)(8 * qd.kc) * HW + qd.pix;
        float x[8][4];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v4 = *reinterpret_cast<const float4*>(a.y + p0 + (size_t)c * HW);
          x[c][0] = valid ? v4.x : 0.f; x[c][1] = valid ? v4.y : 0.f; x[c][2] = valid ? v4.z : 0.f; x[c][3] = valid ? v4.w : 0.f;
        }
        float mean[4], rstd[4];
        gn_stats_quad(qd, x, a.eps, mean, rstd);
        stat_store(sm, qd, v, 0, mean, rstd);
        float4 af[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) af[q] = gn_aff(mean[q], rstd[q], sm.s.gnp[0 * 32 + 4 * qd.kc + q], post1);
        act_to_A_quad(qd, sm.s.abase + (uint32_t)v * kVBytes, x, af, 1.0f / post1, a.R[0], p0, valid);
        publish(v);
      }
      // ---- P2: c1 -> GN2 -> ReLU -> R2, A image of conv2 (model.py:344-346)
#pragma unroll 1
      for (int v = 0; v < nv; ++v) {
        const int img = (unit0 + 2 * v + cta + r * stride) * kImgs + qd.il;
        const bool valid = img < a.g.N;
        const size_t p0 = (valid ? (size_t)img * kC * HW : (size_t)0) + (size_t)(8 * qd.kc) * HW + qd.pix;
        const uint32_t vbase = sm.s.abase + (uint32_t)v * kVBytes;
        wait_acc(v, 0);
        {
          float xp[32];
          acc_read_pos(sm.s, me, xp, tmem, v, kColC1, 0, w.scal[4], t);
          stage_pos(me, vbase, xp);
        }
        group_sync(me.grp);
        float x[8][4];
        unstage_quad(qd, vbase, x, 1.0f);
        if (!valid) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int e = 0; e < 4; ++e) x[c][e] = 0.f;
        }
        float mean[4], rstd[4];
        gn_stats_quad(qd, x, a.eps, mean, rstd);
        stat_store(sm, qd, v, 1, mean, rstd);
        float4 af[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) af[q] = gn_aff(mean[q], rstd[q], sm.s.gnp[1 * 32 + 4 * qd.kc + q], post2);
        act_to_A_quad(qd, vbase, x, af, 1.0f / post2, a.R[1], p0, valid);
        publish(v);
      }
      // ---- P3: c2 -> GN3 = f; backward of GN3 with cotangent -a (adjoint.py:43) -> GC2, gradient operand of dgrad(conv2)
#pragma unroll 1
      for (int v = 0; v < nv; ++v) {
        const int img = (unit0 + 2 * v + cta + r * stride) * kImgs + qd.il;
        const bool valid = img < a.g.N;
        const size_t p0 = (valid ? (size_t)img * kC * HW : (size_t)0) + (size_t)(8 * qd.kc) * HW + qd.pix;
        const uint32_t vbase = sm.s.abase + (uint32_t)v * kVBytes;
        wait_acc(v, 1);
        {
          float xp[32];
          acc_read_pos(sm.s, me, xp, tmem, v, kColC2, 1, w.scal[5], t);
          stage_pos(me, vbase, xp);
        }
        group_sync(me.grp);
        float x[8][4], g[8][4];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v4 = *reinterpret_cast<const float4*>(a.adj + p0 + (size_t)c * HW);
          g[c][0] = valid ? -v4.x : 0.f; g[c][1] = valid ? -v4.y : 0.f; g[c][2] = valid ? -v4.z : 0.f; g[c][3] = valid ? -v4.w : 0.f;
        }
        unstage_quad(qd, vbase, x, 1.0f);
        if (!valid) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int e = 0; e < 4; ++e) x[c][e] = 0.f;
        }
        float mean[4], rstd[4];
        gn_stats_quad(qd, x, a.eps, mean, rstd);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 af = gn_aff(mean[c >> 1], rstd[c >> 1], sm.s.gnp[2 * 32 + 4 * qd.kc + (c >> 1)], a.tsign);
            const float aa = (c & 1) ? af.y : af.x, bb = (c & 1) ? af.w : af.z;
            *reinterpret_cast<float4*>(a.f_out + p0 + (size_t)c * HW) =
                make_float4(fmaf(x[c][0], aa, bb), fmaf(x[c][1], aa, bb), fmaf(x[c][2], aa, bb), fmaf(x[c][3], aa, bb));
          }
        }
        gn_backward_quad(sm, qd, me.warp, 2, x, g, mean, rstd);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(a.GC[1] + p0 + (size_t)c * HW) = make_float4(g[c][0], g[c][1], g[c][2], g[c][3]);
          time_grad(1, g);
        }
        const float scale = grad_scale_img(sm, qd, v, 0, g, gc_max[1]);
        grad_to_A_quad(qd, vbase, g, scale);
        publish(v);
      }
      // ---- P4: dL/dr2 -> ReLU mask -> backward of GN2 -> GC1, gradient operand of dgrad(conv1)
#pragma unroll 1
      for (int v = 0; v < nv; ++v) {
        const int img = (unit0 + 2 * v + cta + r * stride) * kImgs + qd.il;
        const bool valid = img < a.g.N;
        const size_t p0 = (valid ? (size_t)img * kC * HW : (size_t)0) + (size_t)(8 * qd.kc) * HW + qd.pix;
        const uint32_t vbase = sm.s.abase + (uint32_t)v * kVBytes;
        wait_acc(v, 2);
        const float mul2 = scale_from(gmx_cell(sm, v, 0, qd.mt), b3, true) * inv_sw2;
        float x[8][4], g[8][4];
        {
          float xp[32];
          acc_read_pos(sm.s, me, xp, tmem, v, kColC2, -1, 1.0f, 0.f);
          stage_pos(me, vbase, xp);
        }
        group_sync(me.grp);
        unstage_quad(qd, vbase, g, mul2);
        group_sync(me.grp);                        // every quad read of the gradient is done before c1 lands in the same entries
        {
          float xp[32];
          acc_read_pos(sm.s, me, xp, tmem, v, kColC1, 0, w.scal[4], t);
          stage_pos(me, vbase, xp);
        }
        group_sync(me.grp);
        unstage_quad(qd, vbase, x, 1.0f);
        float mean[4], rstd[4];
        stat_load(sm, qd, v, 1, mean, rstd);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 af = gn_aff(mean[c >> 1], rstd[c >> 1], sm.s.gnp[1 * 32 + 4 * qd.kc + (c >> 1)], post2);
          const float aa = (c & 1) ? af.y : af.x, bb = (c & 1) ? af.w : af.z;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (!valid || !(fmaf(x[c][e], aa, bb) > 0.f)) g[c][e] = 0.f;
        }
        if (!valid) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int e = 0; e < 4; ++e) x[c][e] = 0.f;
        }
        gn_backward_quad(sm, qd, me.warp, 1, x, g, mean, rstd);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<float4*>(a.GC[0] + p0 + (size_t)c * HW) = make_float4(g[c][0], g[c][1], g[c][2], g[c][3]);
          time_grad(0, g);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
#pragma unroll
            for (int e = 0; e < 4; ++e) g[c][e] = 0.f;
        }
        const float scale = grad_scale_img(sm, qd, v, 1, g, gc_max[0]);
        grad_to_A_quad(qd, vbase, g, scale);
        publish(v);
      }
      // ---- P5: dL/dr1 -> ReLU mask -> backward of GN1 -> vjp_y
#pragma unroll 1
      for (int v = 0; v < nv; ++v) {
        const int img = (unit0 + 2 * v + cta + r * stride) * kImgs + qd.il;
        const bool valid = img < a.g.N;
        const size_t p0 = (valid ? (size_t)img * kC * HW : (size_t)0) + (size_t)(8 * qd.kc) * HW + qd.pix;
        const uint32_t vbase = sm.s.abase + (uint32_t)v * kVBytes;
        float x[8][4], g[8][4];
#pragma unroll
        for (int c = 0; c < 8; ++c) {               // y again: requested before the wait for the last conv job
          const float4 v4 = *reinterpret_cast<const float4*>(a.y + p0 + (size_t)c * HW);
          x[c][0] = valid ? v4.x : 0.f; x[c][1] = valid ? v4.y : 0.f; x[c][2] = valid ? v4.z : 0.f; x[c][3] = valid ? v4.w : 0.f;
        }
        wait_acc(v, 3);
        const float mul1 = scale_from(gmx_cell(sm, v, 1, qd.mt), b3, true) * inv_sw1;
        {
          float xp[32];
          acc_read_pos(sm.s, me, xp, tmem, v, kColC2, -1, 1.0f, 0.f);
          stage_pos(me, vbase, xp);
        }
        group_sync(me.grp);
        unstage_quad(qd, vbase, g, mul1);
        float mean[4], rstd[4];
        stat_load(sm, qd, v, 0, mean, rstd);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 af = gn_aff(mean[c >> 1], rstd[c >> 1], sm.s.gnp[0 * 32 + 4 * qd.kc + (c >> 1)], post1);
          const float aa = (c & 1) ? af.y : af.x, bb = (c & 1) ? af.w : af.z;
#pragma unroll
          for (int e = 0; e < 4; ++e)
            if (!valid || !(fmaf(x[c][e], aa, bb) > 0.f)) g[c][e] = 0.f;
        }
        gn_backward_quad(sm, qd, me.warp, 0, x, g, mean, rstd);
        if (valid) {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(a.vy_out + p0 + (size_t)c * HW) =
                make_float4(g[c][0] * a.tsign, g[c][1] * a.tsign, g[c][2] * a.tsign, g[c][3] * a.tsign);
        }
        // the A image of slot v is rewritten by the next round's P1: its last readers are this thread's own unstage above
        // (same entries) and dgrad(conv1)'s MMAs (complete: wait_acc); the position-mapped staging writes of the next round
        // come after the next P1's publish. Nothing to wait for.
      }
    }
  }

  // ---- per-CTA partials: GroupNorm affine gradients (fold the two M tiles), vjp_t, max |GC| (operand scale of k_wgrad)
  __syncthreads();
  for (int i = tid; i < 6 * 64; i += kThreads) {
    const int q = i >> 6, c = i & 63;              // q = norm*2 + {gamma, beta}
    const int kc = c >> 3, hh = kc >> 2, wq = kc & 3, n = q >> 1, kind = q & 1;
    float tot = 0.f;
    for (int mt = 0; mt < 2; ++mt) tot += sm.chacc[((hh * 8 + mt * 4 + wq) * 3 + n) * 16 + kind * 8 + (c & 7)];
    a.chan_part[(size_t)blockIdx.x * 384 + i] = tot;
  }
  const double tsum = block_sum((double)tacc, sm.s.scratch);
  if (tid == 0) a.t_part[blockIdx.x] = tsum;
#pragma unroll
  for (int q = 0; q < 2; ++q) {                    // non-negative floats order like their bit patterns
    float m = gc_max[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0 && m > 0.f && m < 3.0e38f) atomicMax(a.gc_max + q, __float_as_uint(m));
  }
  if (timeout) atomicOr(&w.ctl->status, NODE_ST_WATCHDOG);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();
  if (tid < 32) ptx::tmem_dealloc_pair(tmem, kTmemCols);
}

// returns the grid (number of CTAs whose partials k_vjp_finalize folds) through *grid_out
static int launch_vjp8(const VjpArgs& a, cudaStream_t st, int* grid_out) {
  constexpr size_t smem = smem_bytes();
  NODE_SET_SMEM_ONCE(k_vjp8, smem);
  const int NST = (a.g.N + kImgs - 1) / kImgs;
  int grid = 2 * ((NST + 3) / 4);                  // CTA pairs: 16 images per pair and round
  if (grid > kMaxGrid) grid = kMaxGrid;
  *grid_out = grid;
  k_vjp8<<<grid, kThreads, smem, st>>>(a);
  return (int)cudaGetLastError();
}

}}  // namespace node::v8
