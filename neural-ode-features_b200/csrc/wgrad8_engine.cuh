// K7b on the dense 8x8 tiling - the weight-gradient GEMM of wgrad_engine.cuh (reference adjoint.py:41-44 ->
// convolution_backward) for [N,64,8,8] operands, restructured around what bound k_wgrad<8,8> (0.33 ms per adjoint evaluation at
// batch 4736, the top kernel of the training step):
//   * DENSE K. The reduction dimension is "every position of every image"; a stage is the 144 entries of two images in the
//     row-interleaved layout of step8_engine.cuh (row-slot = 8 pixels + one zero entry, rows of the two images interleaved, zero
//     slots as halo): 9 K-steps of 16 per stage = 4.5 per image against 5.33 on the padded strips, and a 3x3 tap is still nothing
//     but a row offset (dy*18 + dx entries) in the B descriptor.
//   * STAGING OVERLAPS THE MMAs. Two operand buffers; 256 stager threads (quad mapping: 8 channels x 4 pixels, 128-bit loads,
//     16 st.shared.v4 per operand) fill buffer b^1 while a dedicated issuer warp streams the 90 tcgen05.mma of buffer b -
//     k_wgrad staged, issued and waited in sequence with every thread.
//   * the ones plane (time-channel and bias gradients) only where it is wanted: N = 80 for the adjoint, N = 64 for the callers.
// Operands: M = 128 = [GC_hi ; GC_lo], N = [IN_hi | ones] then IN_lo into the same columns, both MN-major fp16 after exact
// power-of-two scaling; accumulators of the CTA's taps (taps split 5 + 4 over two CTA groups) stay in tensor memory across all
// stages; per-CTA partials in k_wgrad's layout, so k_vjp_finalize / k_wgrad_fold are unchanged.
#pragma once
#include "wgrad_engine.cuh"

namespace node { namespace w8 {

constexpr int kEntries = 144;                        // K extent of a stage: 16 row-slots x 9 entries (two images)
constexpr int kHalo = 20;                            // zero entries before and after the stage (taps reach +-19 entries)
constexpr int kGStride = kEntries * 16;              // bytes between 8-channel chunks of the GC image
constexpr int kRStride = (kEntries + 2 * kHalo) * 16;
constexpr int kGChunks = 16, kRChunks = 18;          // GC hi 8 + lo 8; IN hi 8 + ones + zero + lo 8
constexpr int kGBytes = kGChunks * kGStride, kRBytes = kRChunks * kRStride;
constexpr int kStageBytes = kGBytes + kRBytes;       // 88,704 B
constexpr int kStagers = 256, kThreads = kStagers + 32;
constexpr size_t kSmem = 1024 + 2 * (size_t)kStageBytes + 64;
static_assert(kSmem <= 227 * 1024, "shared memory budget");
static_assert(128 * (kWgCols + 1) * 4 <= kStageBytes, "drain staging");

template <bool ONES>
__global__ void __launch_bounds__(kThreads, 1) k_wgrad8(const WgradArgs a) {
  constexpr int HW = 64;
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x;
  const int split = blockIdx.x, tg = blockIdx.y, cv = blockIdx.z;
  const int tap0 = tg == 0 ? 0 : 5, ntap = tg == 0 ? 5 : 4;

  const uint32_t s0 = ptx::smem_u32(smem_raw);
  const uint32_t al = (s0 + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (al - s0);
  const uint32_t bar_full = al + 2 * kStageBytes, bar_free = bar_full + 16, bar_done = bar_full + 32;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + 2 * (size_t)kStageBytes + 48);
  {
    uint4* z = reinterpret_cast<uint4*>(base);
    for (int i = tid; i < 2 * kStageBytes / 16; i += kThreads) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  if (ONES) {      // the ones plane: channel 0 of chunk 8 = fp16 1.0 on every pixel entry of both buffers (never rewritten)
    for (int i = tid; i < 2 * 128; i += kThreads) {
      const int b = i >> 7, p = i & 127, slot = p >> 3, x = p & 7;
      *reinterpret_cast<uint4*>(base + (size_t)b * kStageBytes + kGBytes + 8 * kRStride + (kHalo + slot * 9 + x) * 16) =
          make_uint4(0x00003C00u, 0u, 0u, 0u);
    }
  }
  if (tid == 0) {
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(bar_full + 8 * i, kStagers / 32); ptx::mbar_init(bar_free + 8 * i, 1); }
    ptx::mbar_init(bar_done, 1);
    ptx::fence_mbar_init();
  }
  if (tid < 32) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), kTmemCols);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const float s_r = *a.scal[cv];
  float s_g = 1.f;
  {
    const float m = __uint_as_float(*a.gc_max[cv]);
    if (m > 0.f && m < 3.0e38f) {
      int ex;
      (void)frexpf(m, &ex);
      int e = 14 - ex;
      e = e > 100 ? 100 : (e < -100 ? -100 : e);
      s_g = exp2f((float)e);
    }
  }
  const int NST = (a.g.N + 1) / 2;                    // stages: two images each
  bool timeout = false;
  uint32_t nstage = 0;
  for (int st = split; st < NST; st += a.nsplit) ++nstage;

  if (tid < kStagers) {
    // quad mapping: image i of the stage, 8-channel chunk c8, image row r, pixels 4hx .. 4hx+3
    const int i = tid >> 7, c8 = (tid >> 4) & 7, r = (tid >> 1) & 7, hx = tid & 1;
    const uint32_t entry0 = (uint32_t)((2 * r + i) * 9 + 4 * hx);
    const float* __restrict__ Rsrc = a.R[cv];
    const float* __restrict__ Gsrc = a.GC[cv];
    // The operands stream from HBM (310 MB per evaluation at batch 4736) and a thread has one stage in flight: the NEXT stage's
    // 16 quads are requested into registers before this stage is converted, so their latency hides behind the conversion and
    // the wait for the buffer.
    float4 nx[2][8];
    auto request = [&](int st2) {
      const int img2 = 2 * st2 + i;
      const size_t q0 = ((st2 < NST && img2 < a.g.N) ? (size_t)img2 * kC * HW : (size_t)0) + (size_t)(8 * c8) * HW + r * 8 + 4 * hx;
#pragma unroll
      for (int c = 0; c < 8; ++c) { nx[0][c] = ptx::ldg128_ordered(Gsrc + q0 + (size_t)c * HW); nx[1][c] = ptx::ldg128_ordered(Rsrc + q0 + (size_t)c * HW); }
    };
    request(split);
    uint32_t it = 0;
#pragma unroll 1
    for (int st = split; st < NST; st += a.nsplit, ++it) {
      const uint32_t b = it & 1u;
      const bool valid = 2 * st + i < a.g.N;
      float4 cur[2][8];
#pragma unroll
      for (int c = 0; c < 8; ++c) { cur[0][c] = nx[0][c]; cur[1][c] = nx[1][c]; }
      request(st + a.nsplit);
      if (it >= 2u && !timeout && !ptx::mbar_wait(bar_free + 8 * b, ((it >> 1) - 1u) & 1u)) timeout = true;
      const uint32_t gbuf = al + b * kStageBytes, rbuf = gbuf + kGBytes;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const float sc = valid ? (half == 0 ? s_g : s_r) : 0.f;
        const float4 (&v)[8] = cur[half];
        const uint32_t row = half == 0 ? gbuf + (uint32_t)c8 * kGStride + entry0 * 16
                                       : rbuf + (uint32_t)c8 * kRStride + (kHalo + entry0) * 16;
        const uint32_t lo_off = half == 0 ? 8u * kGStride : 10u * kRStride;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint32_t hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 qa = v[2 * j], qb = v[2 * j + 1];
            const float v0 = (e == 0 ? qa.x : (e == 1 ? qa.y : (e == 2 ? qa.z : qa.w))) * sc;
            const float v1 = (e == 0 ? qb.x : (e == 1 ? qb.y : (e == 2 ? qb.z : qb.w))) * sc;
            const __half2 h = __floats2half2_rn(v0, v1);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
            hi[j] = *reinterpret_cast<const uint32_t*>(&h);
            lo[j] = *reinterpret_cast<const uint32_t*>(&l);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + e * 16), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + lo_off + e * 16), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        }
      }
      ptx::fence_proxy_async();
      __syncwarp();
      if ((tid & 31) == 0) ptx::mbar_arrive(bar_full + 8 * b);
    }
  } else {
    // issuer warp: warp-uniform control flow, one elected lane issues
    const bool lead = ptx::elect_one();
    constexpr uint32_t idesc_hi = wg_idesc(ONES ? kWgCols : 64), idesc_lo = wg_idesc(64);
    // descriptors: only the 14-bit start-address field (bytes >> 4) changes from MMA to MMA - one 32-bit add on the low word
    // (the single issuing thread is instruction-latency bound: rebuilding three descriptors per K-step cost 3x the MMA time)
    auto pack = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
    uint32_t a_lo0[2], bh_lo0[2], bl_lo0[2], a_hiw, b_hiw;
    {
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        const uint32_t gbuf = al + b * kStageBytes, rbuf = gbuf + kGBytes;
        const uint64_t da = ptx::make_desc_nosw(gbuf, 128u, (uint32_t)kGStride);
        const uint64_t dh = ptx::make_desc_nosw(rbuf + kHalo * 16, 128u, (uint32_t)kRStride);
        const uint64_t dl = ptx::make_desc_nosw(rbuf + kHalo * 16 + 10u * (uint32_t)kRStride, 128u, (uint32_t)kRStride);
        a_lo0[b] = (uint32_t)da; bh_lo0[b] = (uint32_t)dh; bl_lo0[b] = (uint32_t)dl;
        a_hiw = (uint32_t)(da >> 32); b_hiw = (uint32_t)(dh >> 32);
      }
    }
#pragma unroll 1
    for (uint32_t it = 0; it < nstage; ++it) {
      const uint32_t b = it & 1u;
      if (!timeout && !ptx::mbar_wait(bar_full + 8 * b, (it >> 1) & 1u)) timeout = true;
      ptx::tc_fence_after();
      const uint32_t al0 = b ? a_lo0[1] : a_lo0[0], bh0 = b ? bh_lo0[1] : bh_lo0[0], bl0 = b ? bl_lo0[1] : bl_lo0[0];
#pragma unroll 1
      for (int tp = 0; tp < ntap; ++tp) {
        const int tap = tap0 + tp;
        const int off = (tap / 3 - 1) * 18 + (tap % 3 - 1);       // entries = 16-byte units: added to the address field directly
        const uint32_t d = tmem + (uint32_t)(tp * kWgCols);
        if (lead) {
#pragma unroll
          for (int kk = 0; kk < kEntries / 16; ++kk) {
            const uint64_t adesc = pack(al0 + (uint32_t)(kk * 16), a_hiw);
            const uint64_t b_hi = pack(bh0 + (uint32_t)(kk * 16 + off), b_hiw);
            const uint64_t b_lo = pack(bl0 + (uint32_t)(kk * 16 + off), b_hiw);
            ptx::mma_f16_ss(d, adesc, b_hi, idesc_hi, (it == 0u && kk == 0) ? 0u : 1u);
            ptx::mma_f16_ss(d, adesc, b_lo, idesc_lo, 1u);
          }
        }
      }
      if (lead) {
        ptx::tc_commit(bar_free + 8 * b);
        if (it + 1u == nstage) ptx::tc_commit(bar_done);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (nstage > 0u && !ptx::mbar_wait_relaxed(bar_done, 0u)) timeout = true;
  ptx::tc_fence_after();

  const float inv_g = 1.0f / s_g, inv_gr = inv_g / s_r;
  // ---- drain: rows 0-63 = GC_hi^T * IN, rows 64-127 = GC_lo^T * IN; their sum is this CTA's partial
  float* stage = reinterpret_cast<float*>(base);          // [128][kWgCols + 1] fp32 over the (now free) operand buffers
  const int warp = tid >> 5, lane = tid & 31;
  constexpr int ncol = ONES ? kWgCols : 64;
#pragma unroll 1
  for (int tp = 0; tp < ntap; ++tp) {
    __syncthreads();
    if (warp < 4) {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tp * kWgCols);
#pragma unroll
      for (int c0 = 0; c0 < ncol; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(taddr + c0, v);
        ptx::tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) stage[(warp * 32 + lane) * (kWgCols + 1) + c0 + j] = nstage == 0u ? 0.f : __uint_as_float(v[j]);
      }
    }
    __syncthreads();
    float* dst = a.part + ((size_t)(split * a.ncv + cv) * 9 + (tap0 + tp)) * 64 * kWgCols;
    for (int i = tid; i < 64 * kWgCols; i += kThreads) {
      const int co = i / kWgCols, n = i % kWgCols;
      dst[i] = n < ncol ? (stage[co * (kWgCols + 1) + n] + stage[(64 + co) * (kWgCols + 1) + n]) * (n < 64 ? inv_gr : inv_g) : 0.f;
    }
  }
  (void)timeout;
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, kTmemCols);
}

static int launch_wgrad8(WgradArgs a, bool ones, cudaStream_t st) {
  a.nsplit = wgrad8_splits(a.g.N, a.ncv);
  if (ones) {
    NODE_SET_SMEM_ONCE((k_wgrad8<true>), kSmem);
    k_wgrad8<true><<<dim3(a.nsplit, 2, a.ncv), kThreads, kSmem, st>>>(a);
  } else {
    NODE_SET_SMEM_ONCE((k_wgrad8<false>), kSmem);
    k_wgrad8<false><<<dim3(a.nsplit, 2, a.ncv), kThreads, kSmem, st>>>(a);
  }
  return (int)cudaGetLastError();
}

}}  // namespace node::w8
