// K7b - weight gradient of the two ConcatConv2d layers (reference adjoint.py:41-44 -> convolution_backward):
//     dW[co, ci, tap] = sum_{n, h, w} GC[n, co, h, w] * IN[n, ci, h + dy, w + dx]        (zero padded)
// with IN = [t*ones, relu(GN(.))] (model.py:320-323).  A GEMM whose REDUCTION dimension is the batch:
// M = co (64), N = ci (64 + the ones plane), K = every position of every image.
//
// Mapping: the step engine's zero-padded position strips again, now as the K dimension. Both operands are
// staged once per super-tile as fp16 hi/lo images (power-of-two scaled) [8-channel chunk][position][8 values] - which is exactly the
// MN-major no-swizzle UMMA layout (8 K-rows x 16 bytes per core matrix) - so a convolution tap is nothing but a
// row offset in the B descriptor, like in the forward engine. One tcgen05.mma has M = 128 = [GC_hi ; GC_lo]
// (16 chunks), K = 16 positions; B = [IN_hi | ones | 0] (N = 80) and B = IN_lo (N = 64) accumulate into the same
// 80 fp32 columns per tap, which stay resident in tensor memory across ALL super-tiles of the CTA.
// 9 taps x 80 columns exceed the 512 columns of tensor memory, so the taps are split over two CTA groups
// (taps 0-4 and 5-8); grid = splits x 2 tap groups x 2 convolutions. The column of the ones plane yields the
// time-channel weight gradient (dW[:, 0, tap] / t) and, at the centre tap, the bias gradient.
// Per-CTA results go to a partial buffer that k_vjp_finalize folds in a fixed order (deterministic).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include "step_engine.cuh"

namespace node {


// kind::f16, fp16 x fp16 -> fp32, A and B MN-major, M = 128
__host__ __device__ constexpr uint32_t wg_idesc(int n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

template <int H_, int W_>
struct WgTile {
  using T = Tile<H_, W_>;
  static constexpr int GCH = 16, RCH = 18;                 // chunks: GC hi 8 + lo 8; IN hi 8 + ones + zero + lo 8
  static constexpr int G_STRIDE = T::P * 16;               // bytes between chunks of the GC image (no halo: A rows never shift)
  static constexpr int R_STRIDE = T::R * 16;               // bytes between chunks of the IN image (halo rows on both sides)
  static constexpr size_t smem = 1024 + (size_t)GCH * G_STRIDE + (size_t)RCH * R_STRIDE + 64;
};

template <int H_, int W_>
__global__ void __launch_bounds__(Tile<H_, W_>::P, 1) k_wgrad(const WgradArgs a) {
  using T = Tile<H_, W_>;
  using WT = WgTile<H_, W_>;
  constexpr int HW = T::HW, P = T::P;
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x;
  const int split = blockIdx.x, tg = blockIdx.y, cv = blockIdx.z;
  const int tap0 = tg == 0 ? 0 : 5, ntap = tg == 0 ? 5 : 4;

  const uint32_t s0 = ptx::smem_u32(smem_raw);
  const uint32_t al = (s0 + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (al - s0);
  const uint32_t gimg = al, rimg = al + WT::GCH * WT::G_STRIDE;
  const uint32_t bar = rimg + WT::RCH * WT::R_STRIDE;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + (size_t)WT::GCH * WT::G_STRIDE + (size_t)WT::RCH * WT::R_STRIDE + 16);
  {
    uint4* z = reinterpret_cast<uint4*>(base);
    const int n16 = (WT::GCH * WT::G_STRIDE + WT::RCH * WT::R_STRIDE) / 16;
    for (int i = tid; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (tid < 32) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), kTmemCols);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler

  // this thread's position in the strip
  const int wt = tid, img_l = wt / T::IS;
  const int r = wt % T::IS, hh = r / T::Wp, ww = r % T::Wp;
  const bool inimg = img_l < T::G && hh < T::H && ww < T::W;
  const int pix = hh * T::W + ww;
  const uint32_t grow = gimg + (uint32_t)wt * 16, rrow = rimg + (uint32_t)(T::HALO + wt) * 16;
  const float* __restrict__ Rsrc = a.R[cv];
  const float* __restrict__ Gsrc = a.GC[cv];
  // descriptor strides of the MN-major no-swizzle layout (confirmed on B200): the "leading" field is the byte
  // distance between groups of 8 positions (K), the "stride" field the distance between 8-channel chunks (M / N)
  constexpr uint32_t g_lbo = 128u, g_sbo = (uint32_t)WT::G_STRIDE;
  constexpr uint32_t r_lbo = 128u, r_sbo = (uint32_t)WT::R_STRIDE;

  // Operand scales (powers of two, exact): the activations are bounded a priori (the forward engine's scale), the gradients by
  // the batch maximum k_vjp has just measured; fp16 hi/lo of the scaled values = 2^-22 relative (the bf16 split was 2^-16).
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
  const int NST = (a.g.N + T::G - 1) / T::G;
  bool timeout = false;
  uint32_t it = 0;
  // The GC half of the NEXT super-tile is requested into registers before the wait for this super-tile's MMAs, so its
  // latency hides behind them (the shared-memory images themselves are single-buffered).
  float pf[64];
  auto prefetch = [&](int st2) {
    if (st2 >= NST) return;
    const int img2 = st2 * T::G + img_l;
    const size_t g2 = (inimg && img2 < a.g.N) ? (size_t)img2 * kC * HW + pix : 0;
#pragma unroll
    for (int c = 0; c < 64; ++c) pf[c] = ptx::ldg_ordered(Gsrc + g2 + (size_t)c * HW);
  };
  prefetch(split);
#pragma unroll 1
  for (int st = split; st < NST; st += a.nsplit, ++it) {
    const int img = st * T::G + img_l;
    const bool valid = inimg && img < a.g.N;
    const size_t goff = valid ? (size_t)img * kC * HW + pix : 0;
    // ---- stage this position's 64 + 64 channels as bf16 hi/lo rows (zeros on padding)
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      const float* __restrict__ src = half == 0 ? Gsrc : Rsrc;
      const uint32_t row = half == 0 ? grow : rrow;
      const uint32_t cstride = half == 0 ? (uint32_t)WT::G_STRIDE : (uint32_t)WT::R_STRIDE;
      const uint32_t lo_chunk0 = half == 0 ? 8u : 10u;
      const float sc = half == 0 ? s_g : s_r;
      // all 64 loads of this half in flight at once (the kernel has one CTA per SM and registers to spare; staging
      // is latency-bound and never overlaps the MMAs of the same super-tile)
      float v[64];
      if (half == 0) {
#pragma unroll
        for (int c = 0; c < 64; ++c) v[c] = pf[c];
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) v[c] = ptx::ldg_ordered(src + goff + (size_t)c * HW);
      }
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float v0 = valid ? v[kc * 8 + 2 * j] * sc : 0.f, v1 = valid ? v[kc * 8 + 2 * j + 1] * sc : 0.f;
          const __half2 h = __floats2half2_rn(v0, v1);
          const float2 hf = __half22float2(h);
          const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
          hi[j] = *reinterpret_cast<const uint32_t*>(&h);
          lo[j] = *reinterpret_cast<const uint32_t*>(&l);
        }
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + kc * cstride), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + (lo_chunk0 + kc) * cstride), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
      }
    }
    {   // the ones plane: channel 0 of chunk 8 (fp16 1.0 = 0x3C00), zero elsewhere
      const uint32_t one = valid ? 0x00003C00u : 0u;
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(rrow + 8u * (uint32_t)WT::R_STRIDE), "r"(one), "r"(0u), "r"(0u), "r"(0u) : "memory");
    }
    ptx::fence_proxy_async();
    __syncthreads();
    if (__shfl_sync(0xffffffffu, tid >> 5, 0) == 0) {      // first warp, warp-uniform; one elected lane issues
      const bool lead = ptx::elect_one();
      ptx::tc_fence_after();
      // only the start-address field (bytes >> 4, the low 14 bits) of a descriptor changes from MMA to MMA: one 32-bit add on
      // the low word instead of three descriptor constructions per K-step (the issuing thread is instruction-latency bound)
      const uint64_t da = ptx::make_desc_nosw(gimg, g_lbo, g_sbo);
      const uint64_t dh = ptx::make_desc_nosw(rimg + (uint32_t)T::HALO * 16, r_lbo, r_sbo);
      const uint64_t dl = ptx::make_desc_nosw(rimg + (uint32_t)T::HALO * 16 + 10u * (uint32_t)WT::R_STRIDE, r_lbo, r_sbo);
      const uint32_t a_hiw = (uint32_t)(da >> 32), b_hiw = (uint32_t)(dh >> 32);
      auto pack = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
#pragma unroll 1
      for (int tp = 0; tp < ntap; ++tp) {
        const int tap = tap0 + tp;
        const int off = (tap / 3 - 1) * T::Wp + (tap % 3 - 1);       // rows = 16-byte units of the address field
        const uint32_t d = tmem + (uint32_t)(tp * kWgCols);
        if (lead) {
#pragma unroll 4
          for (int k0 = 0; k0 < P; k0 += 16) {
            const uint64_t adesc = pack((uint32_t)da + (uint32_t)k0, a_hiw);
            const uint64_t b_hi = pack((uint32_t)dh + (uint32_t)(k0 + off), b_hiw);
            const uint64_t b_lo = pack((uint32_t)dl + (uint32_t)(k0 + off), b_hiw);
            ptx::mma_f16_ss(d, adesc, b_hi, wg_idesc(kWgCols), (it == 0 && k0 == 0) ? 0u : 1u);
            ptx::mma_f16_ss(d, adesc, b_lo, wg_idesc(64), 1u);
          }
        }
      }
      if (lead) ptx::tc_commit(bar);
      __syncwarp();
    }
    prefetch(st + a.nsplit);
    if (!timeout && !ptx::mbar_wait_relaxed(bar, it & 1)) timeout = true;   // operands free again / accumulators current
    ptx::tc_fence_after();
  }

  const float inv_g = 1.0f / s_g, inv_gr = inv_g / s_r;
  // ---- drain: rows 0-63 = GC_hi^T * IN, rows 64-127 = GC_lo^T * IN; their sum is this CTA's partial
  float* stage = reinterpret_cast<float*>(base);          // [128][kWgCols + 1] fp32 over the (now free) operand images
  const int warp = tid >> 5, lane = tid & 31;
#pragma unroll 1
  for (int tp = 0; tp < ntap; ++tp) {
    __syncthreads();
    if (warp < 4) {
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(tp * kWgCols);
#pragma unroll
      for (int c0 = 0; c0 < kWgCols; c0 += 16) {
        uint32_t v[16];
        ptx::tmem_ld16(taddr + c0, v);
        ptx::tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) stage[(warp * 32 + lane) * (kWgCols + 1) + c0 + j] = it == 0 ? 0.f : __uint_as_float(v[j]);
      }
    }
    __syncthreads();
    float* dst = a.part + ((size_t)(split * a.ncv + cv) * 9 + (tap0 + tp)) * 64 * kWgCols;
    for (int i = tid; i < 64 * kWgCols; i += blockDim.x) {
      const int co = i / kWgCols, n = i % kWgCols;
      dst[i] = (stage[co * (kWgCols + 1) + n] + stage[(64 + co) * (kWgCols + 1) + n]) * (n < 64 ? inv_gr : inv_g);   // the ones column carries no activation scale
    }
  }
  (void)timeout;
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, kTmemCols);
}

template <int H_, int W_>
static int launch_wgrad_shape(WgradArgs a, cudaStream_t st) {
  using T = Tile<H_, W_>;
  using WT = WgTile<H_, W_>;
  static_assert(WT::smem <= 227 * 1024, "shared memory budget");
  static_assert(128 * (kWgCols + 1) * 4 <= WT::GCH * WT::G_STRIDE + WT::RCH * WT::R_STRIDE, "drain staging");
  NODE_SET_SMEM_ONCE((k_wgrad<H_, W_>), WT::smem);
  const int NST = (a.g.N + T::G - 1) / T::G;
  a.nsplit = wgrad_splits(NST, a.ncv);
  k_wgrad<H_, W_><<<dim3(a.nsplit, 2, a.ncv), T::P, WT::smem, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace node

// One translation unit per feature-map shape: vjp_shape_HxW.cu
#define NODE_VJP_SHAPE_TU(H, W) \
  namespace node { \
  int launch_vjp_##H##x##W(const VjpArgs& a, cudaStream_t st) { return launch_vjp_shape<H, W>(a, st); } \
  int launch_wgrad_##H##x##W(const WgradArgs& a, cudaStream_t st) { return launch_wgrad_shape<H, W>(a, st); } }

// weight gradients only (the callers' convolutions): wgrad_shape_HxW.cu
#define NODE_WGRAD_SHAPE_TU(H, W) \
  namespace node { int launch_wgrad_##H##x##W(const WgradArgs& a, cudaStream_t st) { return launch_wgrad_shape<H, W>(a, st); } }
