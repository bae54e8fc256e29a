// K7a - fused vector-Jacobian product of the ODE-Net dynamics (reference adjoint.py:32-55 applied to
// model.py:339-348): one launch evaluates, for every image,
//     f      = s * ODEfunc(s*t, y)
//     vjp_y  = s * d<f, -a>/dy        (a = adj_y; the cotangent -a is adjoint.py:43)
// and leaves everything the parameter / time gradients need:
//     R1 = relu(GN1(y)), R2 = relu(GN2(c1))      the convolution inputs
//     GC1, GC2 = dL/dc1, dL/dc2                   the gradients at the convolution outputs
//     per-CTA partial sums of the six GroupNorm affine gradients and of vjp_t = sum GC*Tmap.
// The weight gradient itself (a GEMM whose reduction dimension is the batch) is K7b, wgrad.cuh.
//
// Mapping: the step engine's (step_engine.cuh) thread-per-position tiling with ONE worker slot. The two
// forward convolutions keep their fp32 accumulators (c1, c2) in TENSOR MEMORY until the backward pass has
// consumed them, so nothing but y and a is read from HBM: 2*MT*64 columns hold c1 and c2, the two data-gradient
// convolutions (the same implicit GEMM with flipped, transposed weight tiles) reuse c2's columns.
// Forward operands are the fp16 hi/lo split of the step engine; gradient operands are split into bf16 hi/lo
// (gradients have no a-priori bound, bf16 keeps fp32's exponent range): g*w ~ g_hi*w_hi + g_lo*w_hi + g_hi*w_lo,
// 2^-16 relative, fp32 accumulation.
#pragma once
#include <cuda_bf16.h>
#include "step_engine.cuh"

namespace node {

constexpr uint32_t kIdBF16N64 = kIdF16N64 | (1u << 7) | (1u << 10);   // kind::f16 with bf16 A and B

constexpr int kTmPitch = 65;   // floats per (conv, border class) row of the time map in shared memory: lanes of different
                               // classes then hit different banks (a pitch of 64 makes every class collide)

struct VjpSmem {
  StepSmem s;            // wring, abase, part, stat (set 0), gnp, bias, tmapc, scratch, barriers of the step engine
  float2* stat3;         // [3][G][32] (mean, rstd) of GN1, GN2, GN3
  float* gsum;           // [G][32]: per image, S1 of 16 groups then S2 of 16 groups (current half)
  float* part64;         // [NWARP][64]
  float* chacc;          // [NWARP][12][32] per-lane channel accumulators: [(norm*2 + {gamma,beta})*2 + half]
};

// 32 values per thread summed over every pixel of the thread's image -> dst[img][32].
template <class T>
__device__ __forceinline__ void reduce32_img(const VjpSmem& sm, const Who& me, const float (&u)[32]) {
  float* pw = sm.part64 + me.warp * 64;
  if (!me.straddle) {
    const float r = xreduce32(u, me.lane);
    pw[me.lane] = r;
  } else {
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = me.isB ? 0.f : u[j];
    const float ra = xreduce32(v, me.lane);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = me.isB ? u[j] : 0.f;
    const float rb = xreduce32(v, me.lane);
    pw[me.lane] = ra; pw[32 + me.lane] = rb;
  }
  slot_sync(0, T::P);
  if (me.wt < T::G * 32) {
    const int img = me.wt >> 5, q = me.wt & 31;
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < T::NWARP; ++w) {
      const int ia = (w * 32) / T::IS, ib = (w * 32 + 31) / T::IS;
      if (ia == img) tot += sm.part64[w * 64 + q];
      if (ib != ia && ib == img) tot += sm.part64[w * 64 + 32 + q];
    }
    sm.gsum[img * 32 + q] = tot;
  }
  slot_sync(0, T::P);
}

// Per-channel sums over the warp's positions, accumulated in the lane's private shared-memory cell.
__device__ __forceinline__ void chan_accumulate(const VjpSmem& sm, const Who& me, int q, const float (&u)[32]) {
  const float r = xreduce32(u, me.lane);
  sm.chacc[(me.warp * 12 + q) * 32 + me.lane] += r;
}

// relu(GN(x)) for 32 channels: fp16 hi/lo image rows for the tensor core (scaled) + the plain value to HBM.
template <class T>
__device__ __forceinline__ void act_to_A(const VjpSmem& sm, const Who& me, int hb, const float (&x)[32], int n, float scale,
                                         bool valid, float* __restrict__ r_out, size_t p0) {
  const float2* st = sm.stat3 + (n * T::G + min(me.img_l, T::G - 1)) * 32 + 16 * hb;
  const float4* gp = sm.s.gnp + n * 32 + 16 * hb;
  const uint32_t row = sm.s.abase + (T::HALO + me.wt) * 16 + 4 * hb * T::LBO;
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
      const float v0 = fmaxf(fmaf(x[2 * g], a0, b0), 0.f), v1 = fmaxf(fmaf(x[2 * g + 1], a1, b1), 0.f);
      if (valid) { r_out[p0 + (size_t)(2 * g) * T::HW] = v0; r_out[p0 + (size_t)(2 * g + 1) * T::HW] = v1; }
      const float r0 = v0 * scale, r1 = v1 * scale;
      const __half2 h = __floats2half2_rn(r0, r1);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    if (valid) {
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + kc * T::LBO), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + T::A_PART + kc * T::LBO), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
    }
  }
}

// Gradient operands of the two data-gradient convolutions. Gradients have no a-priori bound, so their fp16 hi/lo split
// (2^-22 relative, like the forward operands; the bf16 split used before was 2^-16 and dominated the adjoint's end-to-end
// error) needs a scale from the data: the 32 values of a half are first STAGED as fp32 in this position's own entries of
// the A image (16 entries x 16 B = its 64 channels; chunk c <- channels 8c..8c+3 in the hi part, 8c+4..8c+7 in the lo part)
// while every thread tracks max|g|; after both halves the CTA agrees on a power-of-two scale for the super-tile and every
// thread converts ITS OWN entries in place, chunk by chunk.
template <class T>
__device__ __forceinline__ void grad_stage(const VjpSmem& sm, const Who& me, int hb, const float (&g)[32], bool valid, float& gmax) {
  const uint32_t row = sm.s.abase + (T::HALO + me.wt) * 16 + 4 * hb * T::LBO;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
#pragma unroll
    for (int i = 0; i < 8; ++i) gmax = fmaxf(gmax, fabsf(g[8 * j + i]));
    if (valid) {
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row + j * T::LBO), "f"(g[8 * j]), "f"(g[8 * j + 1]), "f"(g[8 * j + 2]), "f"(g[8 * j + 3]) : "memory");
      asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row + T::A_PART + j * T::LBO), "f"(g[8 * j + 4]), "f"(g[8 * j + 5]), "f"(g[8 * j + 6]), "f"(g[8 * j + 7]) : "memory");
    }
  }
}

// Power-of-two scale that brings max|g| of the super-tile just below 2^14; returns the scale, leaves 1/scale in inv.
template <class T>
__device__ __forceinline__ float grad_scale(const VjpSmem& sm, const Who& me, float gmax, float& inv) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) gmax = fmaxf(gmax, __shfl_xor_sync(0xffffffffu, gmax, o));
  float* red = sm.part64;                       // [NWARP] (free between the GroupNorm reductions)
  slot_sync(0, T::P);
  if (me.lane == 0) red[me.warp] = gmax;
  slot_sync(0, T::P);
  float m = 0.f;
#pragma unroll
  for (int w = 0; w < T::NWARP; ++w) m = fmaxf(m, red[w]);
  int e = 0;
  if (m > 0.f && m < 3.0e38f) {
    int ex;
    (void)frexpf(m, &ex);                       // m = f * 2^ex, f in [0.5, 1)
    e = 14 - ex;
    e = e > 100 ? 100 : (e < -100 ? -100 : e);
  }
  inv = exp2f((float)-e);
  return exp2f((float)e);
}

template <class T>
__device__ __forceinline__ void grad_finalize_A(const VjpSmem& sm, const Who& me, float scale, bool valid) {
  if (!valid) return;
  const uint32_t row = sm.s.abase + (T::HALO + me.wt) * 16;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    float a[8];
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a[0]), "=f"(a[1]), "=f"(a[2]), "=f"(a[3]) : "r"(row + c * T::LBO));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a[4]), "=f"(a[5]), "=f"(a[6]), "=f"(a[7]) : "r"(row + T::A_PART + c * T::LBO));
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v0 = a[2 * j] * scale, v1 = a[2 * j + 1] * scale;
      const __half2 h = __floats2half2_rn(v0, v1);
      const float2 hf = __half22float2(h);
      const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
      hi[j] = *reinterpret_cast<const uint32_t*>(&h);
      lo[j] = *reinterpret_cast<const uint32_t*>(&l);
    }
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + c * T::LBO), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + T::A_PART + c * T::LBO), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
  }
}

// One conv job = 9 taps of MT x 4 k-steps x 3 MMAs (hi*hi, lo*hi, hi*lo) into `dcol`. As in the step engine the weight
// tiles form one sequence (tile i = tap i % 9 of job i / 9; set of job j: conv1, conv2, dgrad2, dgrad1), the leader
// (warp 0) only waits for tiles to land and issues, the producer (warp 1) requests tile i once tile i - kNW has retired.
__device__ __forceinline__ int vjp_set_of(uint32_t job) { const uint32_t k = job & 3u; return k == 0 ? 0 : (k == 1 ? 1 : (k == 2 ? 3 : 2)); }

__device__ __forceinline__ void vjp_request_tile(const VjpSmem& sm, const uint16_t* __restrict__ w16, uint32_t i) {
  const uint32_t slot = i % kNW, set = (uint32_t)vjp_set_of(i / 9), tp = i % 9;
  ptx::mbar_expect_tx(sm.s.bar_wfull + 8 * slot, kW16TileBytes);
  ptx::bulk_g2s(sm.s.wring + slot * kW16TileBytes, (const char*)w16 + (size_t)(set * 9 + tp) * kW16TileBytes, kW16TileBytes,
                sm.s.bar_wfull + 8 * slot);
}

__device__ __forceinline__ void vjp_produce_job(const VjpSmem& sm, const uint16_t* __restrict__ w16, uint32_t job, uint32_t total,
                                                bool& timeout) {
  const bool lead = ptx::elect_one();
#pragma unroll 1
  for (uint32_t i = job * 9 + kWAhead; i < job * 9 + kWAhead + 9 && i < total; ++i) {
    const uint32_t slot = i % kNW;
    if (i >= (uint32_t)kNW && !timeout && !ptx::mbar_wait(sm.s.bar_wfree + 8 * slot, ((i / kNW) - 1) & 1)) timeout = true;
    if (lead) vjp_request_tile(sm, w16, i);
  }
  __syncwarp();
}

template <class T>
__device__ __forceinline__ void vjp_issue_job(const VjpSmem& sm, uint32_t tmem, uint32_t dcol, uint32_t idesc, uint32_t job,
                                              bool& timeout, int mt_used) {
  // whole first warp, warp-uniform arguments, asynchronous instructions by one elected lane (see step_engine.cuh)
  const bool lead = ptx::elect_one();
  ptx::tc_fence_after();
  const uint32_t abase = sm.s.abase + T::HALO * 16;
  const uint32_t a_lo0 = ((abase & 0x3FFFFu) >> 4) | (((uint32_t)T::LBO >> 4) << 16);
  constexpr uint32_t a_hiw = (128u >> 4) | (1u << 14), b_hiw = (1024u >> 4) | (1u << 14) | (2u << 29);
  auto pack = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
#pragma unroll 1
  for (int tap = 0; tap < 9; ++tap) {
    const uint32_t tile = job * 9 + tap, slot = tile % kNW;
    if (!timeout && !ptx::mbar_wait(sm.s.bar_wfull + 8 * slot, (tile / kNW) & 1)) timeout = true;
    ptx::tc_fence_after();
    const int off = (tap / 3 - 1) * T::Wp + (tap % 3 - 1);
    const uint32_t a_tap = a_lo0 + (uint32_t)off;
    const uint32_t b_hi0 = ((sm.s.wring + slot * kW16TileBytes) & 0x3FFFFu) >> 4, b_lo0 = b_hi0 + (uint32_t)((64 * 128) >> 4);
    if (lead) {
#pragma unroll
      for (int mt = 0; mt < T::MT; ++mt) {
        if (mt >= mt_used) continue;           // M tile without a valid image (batch 1: one of two)
        const uint32_t d = tmem + dcol + (uint32_t)(mt * 64);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t a_hi = pack(a_tap + (uint32_t)((mt * 128 * 16 + 2 * ks * T::LBO) >> 4), a_hiw);
          const uint64_t a_lo = pack(a_tap + (uint32_t)((mt * 128 * 16 + 2 * ks * T::LBO + T::A_PART) >> 4), a_hiw);
          const uint32_t kadv = (uint32_t)((ks * 32) >> 4);
          ptx::mma_f16_ss(d, a_hi, pack(b_hi0 + kadv, b_hiw), idesc, (tap == 0 && ks == 0) ? 0u : 1u);
          ptx::mma_f16_ss(d, a_lo, pack(b_hi0 + kadv, b_hiw), idesc, 1u);
          ptx::mma_f16_ss(d, a_hi, pack(b_lo0 + kadv, b_hiw), idesc, 1u);
        }
      }
      ptx::tc_commit(sm.s.bar_wfree + 8 * slot);
    }
  }
  if (lead) ptx::tc_commit(sm.s.bar_acc);
  __syncwarp();
}

template <class T>
__device__ __forceinline__ void vjp_conv_run(const VjpSmem& sm, const Who& me, uint32_t total, const uint16_t* __restrict__ w16,
                                             uint32_t tmem, uint32_t dcol, uint32_t idesc, uint32_t& njob, bool& timeout, int mt_used) {
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  slot_sync(0, T::P);
  const int wu = __shfl_sync(0xffffffffu, me.warp, 0);     // shuffles: tell the compiler these values are warp-uniform
  if (wu == 0)
    vjp_issue_job<T>(sm, __shfl_sync(0xffffffffu, tmem, 0), __shfl_sync(0xffffffffu, dcol, 0), __shfl_sync(0xffffffffu, idesc, 0),
                     __shfl_sync(0xffffffffu, njob, 0), timeout, __shfl_sync(0xffffffffu, mt_used, 0));
  else if (wu == 1)
    vjp_produce_job(sm, w16, __shfl_sync(0xffffffffu, njob, 0), __shfl_sync(0xffffffffu, total, 0), timeout);
  if (!timeout && !ptx::mbar_wait_relaxed(sm.s.bar_acc, njob & 1)) timeout = true;
  ++njob;
  ptx::tc_fence_after();
}

// 32 accumulator columns of this thread's row -> x = acc*mul (+ bias + t*Tmap when cv >= 0).
template <class T>
__device__ __forceinline__ void vjp_tmem_read(const VjpSmem& sm, const Who& me, int hb, float (&x)[32], uint32_t tmem, uint32_t col,
                                              int cv, float mul, float t, bool valid) {
  const uint32_t taddr = tmem + ((uint32_t)((me.warp & 3) * 32) << 16) + col + (uint32_t)((me.wt >> 7) * 64 + 32 * hb);
#pragma unroll
  for (int c0 = 0; c0 < 32; c0 += 16) {
    uint32_t v[16];
    ptx::tmem_ld16(taddr + c0, v);
    ptx::tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float acc = __uint_as_float(v[j]) * mul;
      if (cv >= 0) {
        const float extra = fmaf(t, sm.s.tmapc[(cv * 9 + me.cls) * kTmPitch + 32 * hb + c0 + j], sm.s.bias[cv * 64 + 32 * hb + c0 + j]);
        x[c0 + j] = valid ? acc + extra : 0.f;
      } else {
        x[c0 + j] = valid ? acc : 0.f;
      }
    }
  }
}

// Backward of one GroupNorm for 32 channels of this thread's position (oracle/odefunc_port.py:_gn_bwd):
// in: x = the norm's input, g = gradient at its output (already masked by the ReLU that follows, if any);
// out: g <- gradient at the norm's input. Accumulates dgamma / dbeta of norm n.
template <class T>
__device__ __forceinline__ void gn_backward(const VjpSmem& sm, const Who& me, int hb, int n, const float (&x)[32], float (&g)[32]) {
  constexpr float inv_m = 1.0f / (float)(kCpg * T::HW);
  const float2* st = sm.stat3 + (n * T::G + min(me.img_l, T::G - 1)) * 32 + 16 * hb;
  const float4* gp = sm.s.gnp + n * 32 + 16 * hb;
  float u[32];
  // dgamma_c = sum g*xhat, dbeta_c = sum g (over every image and pixel)
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 s = st[j];
    u[2 * j] = g[2 * j] * ((x[2 * j] - s.x) * s.y);
    u[2 * j + 1] = g[2 * j + 1] * ((x[2 * j + 1] - s.x) * s.y);
  }
  chan_accumulate(sm, me, (n * 2 + 0) * 2 + hb, u);
  chan_accumulate(sm, me, (n * 2 + 1) * 2 + hb, g);
  // S1 = sum_cell g*gamma, S2 = sum_cell g*gamma*xhat
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 s = st[j];
    const float4 p = gp[j];
    const float d0 = g[2 * j] * p.x, d1 = g[2 * j + 1] * p.y;
    u[j] = d0 + d1;
    u[16 + j] = d0 * ((x[2 * j] - s.x) * s.y) + d1 * ((x[2 * j + 1] - s.x) * s.y);
  }
  reduce32_img<T>(sm, me, u);
  const float* gs = sm.gsum + min(me.img_l, T::G - 1) * 32;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float2 s = st[j];
    const float4 p = gp[j];
    const float m1 = gs[j] * inv_m, m2 = gs[16 + j] * inv_m;
    const float xh0 = (x[2 * j] - s.x) * s.y, xh1 = (x[2 * j + 1] - s.x) * s.y;
    g[2 * j] = s.y * (g[2 * j] * p.x - m1 - xh0 * m2);
    g[2 * j + 1] = s.y * (g[2 * j + 1] * p.y - m1 - xh1 * m2);
  }
}

template <int H_, int W_>
__global__ void __launch_bounds__(Tile<H_, W_>::P, 1) k_vjp(const VjpArgs a) {
  using T = Tile<H_, W_>;
  constexpr int HW = T::HW, P = T::P;
  extern __shared__ uint8_t smem_raw[];
  const FusedWs& w = a.w;
  const int tid = threadIdx.x;

  VjpSmem sm;
  {
    const uint32_t s0 = ptx::smem_u32(smem_raw);
    const uint32_t al = (s0 + 1023u) & ~1023u;
    uint8_t* base = smem_raw + (al - s0);
    size_t o = 0;
    sm.s.wring = al; o += (size_t)kNW * kW16TileBytes;
    sm.s.abase = al + (uint32_t)o; o += (size_t)2 * T::A_PART;
    sm.s.part = reinterpret_cast<float*>(base + o); o += (size_t)T::NWARP * 32 * 4;
    sm.part64 = reinterpret_cast<float*>(base + o); o += (size_t)T::NWARP * 64 * 4;
    sm.stat3 = reinterpret_cast<float2*>(base + o); o += (size_t)3 * T::G * 32 * 8;
    sm.gsum = reinterpret_cast<float*>(base + o); o += (size_t)T::G * 32 * 4;
    sm.chacc = reinterpret_cast<float*>(base + o); o += (size_t)T::NWARP * 12 * 32 * 4;
    sm.s.gnp = reinterpret_cast<float4*>(base + o); o += 3 * 32 * 16;
    sm.s.bias = reinterpret_cast<float*>(base + o); o += 2 * 64 * 4;
    sm.s.tmapc = reinterpret_cast<float*>(base + o); o += 2 * 9 * kTmPitch * 4 + 8;
    sm.s.scratch = reinterpret_cast<double*>(base + o); o += 32 * 8;
    sm.s.bar_wfull = al + (uint32_t)o; o += 8 * kNW;
    sm.s.bar_wfree = al + (uint32_t)o; o += 8 * kNW;
    sm.s.bar_acc = al + (uint32_t)o; o += 8;
    sm.s.tmem_slot = reinterpret_cast<uint32_t*>(base + o);
    sm.s.stat = sm.stat3;
    uint4* az = reinterpret_cast<uint4*>(base + (size_t)kNW * kW16TileBytes);
    for (int i = tid; i < 2 * T::A_PART / 16; i += blockDim.x) az[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = tid; i < T::NWARP * 12 * 32; i += blockDim.x) sm.chacc[i] = 0.f;
  }
  for (int i = tid; i < 3 * 32; i += blockDim.x) {
    const int n = i / 32, g = i % 32;
    sm.s.gnp[i] = make_float4(w.gn[(2 * n) * kC + 2 * g], w.gn[(2 * n) * kC + 2 * g + 1], w.gn[(2 * n + 1) * kC + 2 * g],
                              w.gn[(2 * n + 1) * kC + 2 * g + 1]);
  }
  for (int i = tid; i < 2 * 64; i += blockDim.x) sm.s.bias[i] = w.bias[i];
  for (int i = tid; i < 2 * 9 * 64; i += blockDim.x) sm.s.tmapc[(i >> 6) * kTmPitch + (i & 63)] = w.tmapc[i];
  if (tid == 0) {
    for (int i = 0; i < kNW; ++i) { ptx::mbar_init(sm.s.bar_wfull + 8 * i, 1); ptx::mbar_init(sm.s.bar_wfree + 8 * i, 1); }
    ptx::mbar_init(sm.s.bar_acc, 1);
    ptx::fence_mbar_init();
  }
  if (tid < 32) ptx::tmem_alloc(ptx::smem_u32(sm.s.tmem_slot), kTmemCols);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *sm.s.tmem_slot;
  constexpr uint32_t kColC1 = 0, kColC2 = T::MT * 64;

  const int gs = a.g.gs;                                // images per super-tile for this batch (<= T::G)
  const int NST = (a.g.N + gs - 1) / gs;
  const int nst = (int)blockIdx.x < NST ? (NST - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t total = (uint32_t)nst * 4u * 9u;
  if (tid == 0)                     // the first tiles of the weight sequence; later ones are requested by the producer warp
    for (uint32_t i = 0; i < kWAhead && i < total; ++i) vjp_request_tile(sm, w.w16, i);
  bool timeout = false;
  uint32_t njob = 0;
  float tacc = 0.f;
  float gc_max[2] = {0.f, 0.f};                  // max |GC1|, |GC2| over this CTA's images (operand scale of k_wgrad)
  const float inv_sw1 = 1.0f / w.scal[2], inv_sw2 = 1.0f / w.scal[3];

  Who me;
  me.slot = 0; me.wt = tid; me.warp = tid >> 5; me.lane = tid & 31;
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
  const float t = a.tsign * a.t_dev[0];      // reversed-time wrapper (misc.py:184-187)
  StepSmem st1 = sm.s, st2 = sm.s, st3 = sm.s;
  st1.stat = sm.stat3; st2.stat = sm.stat3 + T::G * 32; st3.stat = sm.stat3 + 2 * T::G * 32;

#pragma unroll 1
  for (int st = blockIdx.x; st < NST; st += gridDim.x) {
    const int img = st * gs + me.img_l;
    const bool valid = me.inimg && me.img_l < gs && img < a.g.N;
    const size_t goff = valid ? (size_t)img * kC * HW + me.pix : (size_t)(me.inimg ? me.pix : 0);
    const int mt_used = min(T::MT, (min(gs, a.g.N - st * gs) * T::IS + 127) / 128);        // M tiles that hold a valid image
    float x[32], g[32];

    // ---- forward: y -> GN1 -> ReLU -> conv1 (model.py:341-343)
#pragma unroll 1
    for (int hb = 0; hb < 2; ++hb) {
      const size_t p0 = goff + (size_t)(32 * hb) * HW;
#pragma unroll
      for (int c = 0; c < 32; ++c) { const float v = a.y[p0 + (size_t)c * HW]; x[c] = valid ? v : 0.f; }
      gn_stats<T>(st1, me, hb, x, valid, a.eps);
      act_to_A<T>(sm, me, hb, x, 0, w.scal[0], valid, a.R[0], p0);
    }
    vjp_conv_run<T>(sm, me, total, w.w16, tmem, kColC1, kIdF16N64, njob, timeout, mt_used);
    // ---- c1 -> GN2 -> ReLU -> conv2 (model.py:344-346)
#pragma unroll 1
    for (int hb = 0; hb < 2; ++hb) {
      const size_t p0 = goff + (size_t)(32 * hb) * HW;
      vjp_tmem_read<T>(sm, me, hb, x, tmem, kColC1, 0, w.scal[4], t, valid);
      gn_stats<T>(st2, me, hb, x, valid, a.eps);
      act_to_A<T>(sm, me, hb, x, 1, w.scal[1], valid, a.R[1], p0);
    }
    vjp_conv_run<T>(sm, me, total, w.w16, tmem, kColC2, kIdF16N64, njob, timeout, mt_used);
    // ---- c2 -> GN3 = f; backward of GN3 with cotangent -a (adjoint.py:43)
    float gmax = 0.f, ginv = 1.f;
#pragma unroll 1
    for (int hb = 0; hb < 2; ++hb) {
      const size_t p0 = goff + (size_t)(32 * hb) * HW;
      vjp_tmem_read<T>(sm, me, hb, x, tmem, kColC2, 1, w.scal[5], t, valid);
      gn_stats<T>(st3, me, hb, x, valid, a.eps);
      {
        const float2* stt = st3.stat + min(me.img_l, T::G - 1) * 32 + 16 * hb;
        const float4* gp = sm.s.gnp + 2 * 32 + 16 * hb;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 s = stt[j];
          const float4 p = gp[j];
          const float a0 = s.y * p.x, a1 = s.y * p.y;
          const float b0 = p.z - a0 * s.x, b1 = p.w - a1 * s.x;
          if (valid) {
            a.f_out[p0 + (size_t)(2 * j) * HW] = fmaf(x[2 * j], a0, b0) * a.tsign;
            a.f_out[p0 + (size_t)(2 * j + 1) * HW] = fmaf(x[2 * j + 1], a1, b1) * a.tsign;
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) { const float v = a.adj[p0 + (size_t)c * HW]; g[c] = valid ? -v : 0.f; }
      gn_backward<T>(sm, me, hb, 2, x, g);
      const float* tm = sm.s.tmapc + (9 + me.cls) * kTmPitch + 32 * hb;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (valid) { a.GC[1][p0 + (size_t)c * HW] = g[c]; tacc = fmaf(g[c], tm[c], tacc); }
      }
      grad_stage<T>(sm, me, hb, g, valid, gmax);
    }
    gc_max[1] = fmaxf(gc_max[1], gmax);
    grad_finalize_A<T>(sm, me, grad_scale<T>(sm, me, gmax, ginv), valid);
    vjp_conv_run<T>(sm, me, total, w.w16, tmem, kColC2, kIdF16N64, njob, timeout, mt_used);      // dL/dr2 over c2's columns
    const float mul2 = ginv * inv_sw2;
    gmax = 0.f;
    // ---- ReLU mask of GN2's output, backward of GN2
#pragma unroll 1
    for (int hb = 0; hb < 2; ++hb) {
      const size_t p0 = goff + (size_t)(32 * hb) * HW;
      vjp_tmem_read<T>(sm, me, hb, g, tmem, kColC2, -1, mul2, 0.f, valid);
      vjp_tmem_read<T>(sm, me, hb, x, tmem, kColC1, 0, w.scal[4], t, valid);
      {
        const float2* stt = st2.stat + min(me.img_l, T::G - 1) * 32 + 16 * hb;
        const float4* gp = sm.s.gnp + 1 * 32 + 16 * hb;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 s = stt[j];
          const float4 p = gp[j];
          const float a0 = s.y * p.x, a1 = s.y * p.y;
          const float b0 = p.z - a0 * s.x, b1 = p.w - a1 * s.x;
          if (!(fmaf(x[2 * j], a0, b0) > 0.f)) g[2 * j] = 0.f;
          if (!(fmaf(x[2 * j + 1], a1, b1) > 0.f)) g[2 * j + 1] = 0.f;
        }
      }
      gn_backward<T>(sm, me, hb, 1, x, g);
      const float* tm = sm.s.tmapc + me.cls * kTmPitch + 32 * hb;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (valid) { a.GC[0][p0 + (size_t)c * HW] = g[c]; tacc = fmaf(g[c], tm[c], tacc); }
      }
      grad_stage<T>(sm, me, hb, g, valid, gmax);
    }
    gc_max[0] = fmaxf(gc_max[0], gmax);
    grad_finalize_A<T>(sm, me, grad_scale<T>(sm, me, gmax, ginv), valid);
    vjp_conv_run<T>(sm, me, total, w.w16, tmem, kColC2, kIdF16N64, njob, timeout, mt_used);      // dL/dr1
    const float mul1 = ginv * inv_sw1;
    // ---- ReLU mask of GN1's output, backward of GN1 -> vjp_y
#pragma unroll 1
    for (int hb = 0; hb < 2; ++hb) {
      const size_t p0 = goff + (size_t)(32 * hb) * HW;
      vjp_tmem_read<T>(sm, me, hb, g, tmem, kColC2, -1, mul1, 0.f, valid);
#pragma unroll
      for (int c = 0; c < 32; ++c) { const float v = a.y[p0 + (size_t)c * HW]; x[c] = valid ? v : 0.f; }
      {
        const float2* stt = st1.stat + min(me.img_l, T::G - 1) * 32 + 16 * hb;
        const float4* gp = sm.s.gnp + 16 * hb;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 s = stt[j];
          const float4 p = gp[j];
          const float a0 = s.y * p.x, a1 = s.y * p.y;
          const float b0 = p.z - a0 * s.x, b1 = p.w - a1 * s.x;
          if (!(fmaf(x[2 * j], a0, b0) > 0.f)) g[2 * j] = 0.f;
          if (!(fmaf(x[2 * j + 1], a1, b1) > 0.f)) g[2 * j + 1] = 0.f;
        }
      }
      gn_backward<T>(sm, me, hb, 0, x, g);
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (valid) a.vy_out[p0 + (size_t)c * HW] = g[c] * a.tsign;
      }
    }
  }

  // ---- per-CTA partials: channel sums (fold the warps) and the time gradient
  __syncthreads();
  for (int i = tid; i < 12 * 32; i += blockDim.x) {
    const int q = i >> 5, lane = i & 31;
    float tot = 0.f;
    for (int wp = 0; wp < T::NWARP; ++wp) tot += sm.chacc[(wp * 12 + q) * 32 + lane];
    // q = (norm*2 + kind)*2 + half  ->  [norm*2 + kind][64]
    a.chan_part[(size_t)blockIdx.x * 384 + (q >> 1) * 64 + (q & 1) * 32 + lane] = tot;
  }
  const double tsum = block_sum((double)tacc, sm.s.scratch);
  if (tid == 0) a.t_part[blockIdx.x] = tsum;
#pragma unroll
  for (int q = 0; q < 2; ++q) {                  // non-negative floats order like their bit patterns
    float m = gc_max[q];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0 && m > 0.f && m < 3.0e38f) atomicMax(a.gc_max + q, __float_as_uint(m));
  }
  if (timeout) atomicOr(&w.ctl->status, NODE_ST_WATCHDOG);
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, kTmemCols);
}

template <int H_, int W_>
constexpr size_t vjp_smem_bytes() {
  using T = Tile<H_, W_>;
  return 1024 + (size_t)kNW * kW16TileBytes + (size_t)2 * T::A_PART + (size_t)T::NWARP * 32 * 4 + (size_t)T::NWARP * 64 * 4 +
         (size_t)3 * T::G * 32 * 8 + (size_t)T::G * 32 * 4 + (size_t)T::NWARP * 12 * 32 * 4 + 3 * 32 * 16 + 2 * 64 * 4 +
         2 * 9 * kTmPitch * 4 + 8 + 32 * 8 + 16 * kNW + 8 + 64;
}

template <int H_, int W_>
static int launch_vjp_shape(const VjpArgs& a, cudaStream_t st) {
  using T = Tile<H_, W_>;
  constexpr size_t smem = vjp_smem_bytes<H_, W_>();
  static_assert(smem <= 227 * 1024, "shared memory budget");
  static_assert(2 * T::MT * 64 <= kTmemCols, "tensor memory budget");
  static_assert(T::G == strip_images(H_, W_), "strip_images out of sync");
  NODE_SET_SMEM_ONCE((k_vjp<H_, W_>), smem);
  const int NST = (a.g.N + a.g.gs - 1) / a.g.gs;
  const int grid = NST < kMaxGrid ? NST : kMaxGrid;
  k_vjp<H_, W_><<<grid, T::P, smem, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace node
