// the tail of the reference's ResBlock (model.py:156-178),
//     out = conv2(relu(norm2(c))) + shortcut          conv2 = Conv2d(64, 64, 3, 1, 1, bias=False), norm2 = GroupNorm(32, 64)
// as ONE kernel on the step engine's machinery (step_engine.cuh): position strips, one thread per position with its
// channels in registers, GroupNorm statistics by warp transpose-reductions, the activation written once as the K-major
// A image of an SS-mode tcgen05.mma implicit GEMM in which a tap is a descriptor row offset, fp32 contract by fp16
// operand splitting, weight tiles streamed by a producer warp. ATen runs this as GroupNorm (2 kernels) + ReLU + cuDNN
// fp32 SIMT convolution + add: 5 launches, 4 extra passes over the tensor.
#pragma once
#include "step_engine.cuh"

namespace node {

struct ResConvArgs {
  const uint16_t* w16;       // [9 taps][128 rows][64 halves] weight tiles (hi rows 0..63, lo rows 64..127), SW128 image
  const float* scal;         // [0] activation scale, [1] weight scale, [2] 1 / (sa * sw)
  const float* gamma; const float* beta;
  const float* gamma_next; const float* beta_next;   // optional: the NEXT block's norm1 - out = relu(norm1_next(conv + shortcut))
  const float* x; const float* shortcut; float* out;
  int N; float eps;
  int64_t in_stride, out_stride;   // elements between images of x / of shortcut and out (0 = 64 * H * W): channel blocks of wider tensors
};

__host__ __device__ constexpr size_t resconv_smem_bytes(int A_PART, int NSLOT, int NWARP, int G) {
  return 1024 + (size_t)kNW * kW16TileBytes + (size_t)NSLOT * 2 * A_PART + (size_t)NSLOT * NWARP * 64 * 4 +
         (size_t)NSLOT * G * 32 * 8 + (size_t)NSLOT * G * 32 * 16 + 2 * 32 * 16 + 16 * 8 + 64;
}

// x * scale (a power of two) split into fp16 hi + lo -> this position's row of the A image (k-chunks [4*hb, 4*hb+4)): the RAW
// operand of the data-gradient mode (no GroupNorm, no ReLU: gradients are signed).
template <class T>
__device__ __forceinline__ void raw_to_A(const StepSmem& sm, const Who& me, int hb, const float (&x)[32], float scale, bool valid) {
  const uint32_t row = sm.abase + me.slot * 2 * T::A_PART + (T::HALO + me.wt) * 16 + 4 * hb * T::LBO;
#pragma unroll
  for (int kc = 0; kc < 4; ++kc) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float r0 = x[8 * kc + 2 * j] * scale, r1 = x[8 * kc + 2 * j + 1] * scale;
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

// RAW = false: out = conv(relu(GN(x))) + shortcut (the ResBlock tail). RAW = true: out = conv(x) (+ shortcut when given) for an
// arbitrary signed input - the data gradient of a 3x3 stride-1 convolution when the weight tiles hold the flipped, transposed
// kernel (caller_ops.py); the operand scale is the power of two that brings the super-tile's |max| just below 2^15.
template <int H_, int W_, int NSLOT, bool RAW = false>
__global__ void __launch_bounds__(NSLOT * Tile<H_, W_>::P, 1) k_resconv(const ResConvArgs a) {
  using T = Tile<H_, W_>;
  constexpr int HW = T::HW, P = T::P;
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x;
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
    sm.gnp = reinterpret_cast<float4*>(base + o); o += 2 * 32 * 16;
    sm.tb = nullptr; sm.bias = nullptr; sm.tmapc = nullptr; sm.coef = nullptr; sm.scratch = nullptr; sm.ring = nullptr;
    sm.bar_wfull = al + (uint32_t)o; o += 8 * kNW;
    sm.bar_wfree = al + (uint32_t)o; o += 8 * kNW;
    sm.bar_turn = al + (uint32_t)o; o += 8 * 2;
    sm.bar_acc = al + (uint32_t)o; o += 8 * 2;
    sm.tmem_slot = reinterpret_cast<uint32_t*>(base + o); o += 8;
    sm.illcond = reinterpret_cast<uint32_t*>(base + o);
    uint4* az = reinterpret_cast<uint4*>(base + (size_t)kNW * kW16TileBytes);   // padding rows / columns stay zero
    for (int i = tid; i < NSLOT * 2 * T::A_PART / 16; i += blockDim.x) az[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  const bool norm_next = a.gamma_next != nullptr;
  for (int g = tid; g < (RAW ? 0 : 32); g += blockDim.x) {
    sm.gnp[g] = make_float4(a.gamma[2 * g], a.gamma[2 * g + 1], a.beta[2 * g], a.beta[2 * g + 1]);
    if (norm_next) sm.gnp[32 + g] = make_float4(a.gamma_next[2 * g], a.gamma_next[2 * g + 1], a.beta_next[2 * g], a.beta_next[2 * g + 1]);
  }
  if (tid < 2) sm.illcond[tid] = 0u;
  if (tid == 0) {
    for (int i = 0; i < kNW; ++i) { ptx::mbar_init(sm.bar_wfull + 8 * i, 1); ptx::mbar_init(sm.bar_wfree + 8 * i, 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(sm.bar_turn + 8 * i, 1); ptx::mbar_init(sm.bar_acc + 8 * i, 1); }
    ptx::fence_mbar_init();
    ptx::mbar_arrive(sm.bar_turn);
  }
  if (tid < 32) ptx::tmem_alloc(ptx::smem_u32(sm.tmem_slot), kTmemCols);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *sm.tmem_slot;

  const int NST = (a.N + T::G - 1) / T::G;
  const int stride = gridDim.x * NSLOT;
  bool timeout = false;
  Jobs jb;
  uint32_t nfull;
  {
    int nst[2] = {0, 0};
#pragma unroll
    for (int s = 0; s < NSLOT; ++s) {
      const int u = blockIdx.x * NSLOT + s;
      nst[s] = u < NST ? (NST - u + stride - 1) / stride : 0;
    }
    nfull = (uint32_t)nst[NSLOT - 1];
    jb.jobs_full = nfull * NSLOT;
    jb.jobs = (uint32_t)(nst[0] + (NSLOT > 1 ? nst[1] : 0));
    jb.single = 1;
  }
  if (tid == 0)
    for (uint32_t i = 0; i < kWAhead && i < jb.jobs * 9; ++i) request_tile<NSLOT>(sm, jb, a.w16, i);

  Who me;
  me.slot = tid / P; me.wt = tid % P; me.warp = me.wt >> 5; me.lane = tid & 31;
  me.img_l = me.wt / T::IS;
  {
    const int r = me.wt % T::IS, hh = r / T::Wp, ww = r % T::Wp;
    me.inimg = me.img_l < T::G && hh < T::H && ww < T::W;
    me.pix = hh * T::W + ww;
    me.cls = 4;
    const int ia = (me.warp * 32) / T::IS, ib = (me.warp * 32 + 31) / T::IS;
    me.straddle = ia != ib;
    me.isB = me.img_l != ia;
  }
  const float sa = a.scal[0], inv = a.scal[2];
  const size_t istr = a.in_stride > 0 ? (size_t)a.in_stride : (size_t)kC * HW, ostr = a.out_stride > 0 ? (size_t)a.out_stride : (size_t)kC * HW;
  uint32_t njob = 0;
#pragma unroll 1
  for (int st = blockIdx.x * NSLOT + me.slot; st < NST; st += stride) {
    const int img = st * T::G + me.img_l;
    const bool valid = me.inimg && img < a.N;
    const size_t goff = valid ? (size_t)img * istr + me.pix : (size_t)(me.inimg ? me.pix : 0);
    const size_t ooff = valid ? (size_t)img * ostr + me.pix : (size_t)(me.inimg ? me.pix : 0);
    float x[32];
    float inv_job = inv;
    if constexpr (RAW) {
      float mx = 0.f;
#pragma unroll 1
      for (int hb = 0; hb < 2; ++hb) {
        const size_t p0 = goff + (size_t)(32 * hb) * HW;
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = ptx::ldg_ordered(a.x + p0 + (size_t)c * HW);
#pragma unroll
        for (int c = 0; c < 32; ++c) mx = fmaxf(mx, valid ? fabsf(x[c]) : 0.f);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float* part = sm.part + me.slot * T::NWARP * 64;
      if (me.lane == 0) part[me.warp] = mx;
      slot_sync(me.slot, T::P);
      mx = 0.f;
#pragma unroll
      for (int w = 0; w < T::NWARP; ++w) mx = fmaxf(mx, part[w]);
      slot_sync(me.slot, T::P);
      float s_job = 1.f;
      if (mx > 0.f && mx < 3.0e38f) {
        int ex;
        (void)frexpf(mx, &ex);
        int e = 14 - ex;
        e = e > 100 ? 100 : (e < -100 ? -100 : e);
        s_job = exp2f((float)e);
      }
      inv_job = inv * sa / s_job;                    // scal[2] = 1 / (sa * sw): the activation scale of the workspace is not used here
#pragma unroll 1
      for (int hb = 0; hb < 2; ++hb) {
        const size_t p0 = goff + (size_t)(32 * hb) * HW;
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = ptx::ldg_ordered(a.x + p0 + (size_t)c * HW);   // second read: L1 / L2
        raw_to_A<T>(sm, me, hb, x, s_job, valid);
      }
    } else {
#pragma unroll 1
      for (int hb = 0; hb < 2; ++hb) {
        const size_t p0 = goff + (size_t)(32 * hb) * HW;
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = ptx::ldg_ordered(a.x + p0 + (size_t)c * HW);   // all 32 loads in flight
        if (!valid) {
#pragma unroll
          for (int c = 0; c < 32; ++c) x[c] = 0.f;
        }
        gn_affine<T>(sm, me, hb, 0, x, valid, a.eps, sa);
        affine_to_A<T>(sm, me, hb, x, valid, true);
      }
    }
    conv_run<T, NSLOT>(sm, me, jb, a.w16, tmem, njob, nfull, timeout, true);
#pragma unroll 1
    for (int hb = 0; hb < 2; ++hb) {
      conv_read<T, false>(sm, me, hb, x, tmem, 0, inv_job, true, valid);
      const size_t p0 = ooff + (size_t)(32 * hb) * HW;
      if (!RAW || a.shortcut != nullptr) {
        float sc[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) sc[c] = ptx::ldg_ordered(a.shortcut + p0 + (size_t)c * HW);   // padding threads: image 0, discarded
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = valid ? x[c] + sc[c] : 0.f;
      }
      if (!RAW && norm_next) {           // kernel-uniform: the next block's GroupNorm -> ReLU on the block output (model.py:167)
        gn_affine<T>(sm, me, hb, 1, x, valid, a.eps);
        const float4* af = sm.aff + (me.slot * T::G + min(me.img_l, T::G - 1)) * 32 + 16 * hb;
#pragma unroll
        for (int g = 0; g < 16; ++g) {
          const float4 p = af[g];
          x[2 * g] = fmaxf(fmaf(x[2 * g], p.x, p.z), 0.f);
          x[2 * g + 1] = fmaxf(fmaf(x[2 * g + 1], p.y, p.w), 0.f);
        }
      }
      if (valid) {
#pragma unroll
        for (int c = 0; c < 32; ++c) a.out[p0 + (size_t)c * HW] = x[c];
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, kTmemCols);
  (void)timeout;
}

template <int H_, int W_, int NSLOT, bool RAW>
static int launch_resconv_slots(const ResConvArgs& a, cudaStream_t st) {
  using T = Tile<H_, W_>;
  constexpr size_t smem = resconv_smem_bytes(T::A_PART, NSLOT, T::NWARP, T::G);
  static_assert(smem <= 227 * 1024, "shared memory budget");
  NODE_SET_SMEM_ONCE((k_resconv<H_, W_, NSLOT, RAW>), smem);
  const int NST = (a.N + T::G - 1) / T::G;
  int grid = (NST + NSLOT - 1) / NSLOT;
  if (grid > kMaxGrid) grid = kMaxGrid;
  k_resconv<H_, W_, NSLOT, RAW><<<grid, NSLOT * T::P, smem, st>>>(a);
  return (int)cudaGetLastError();
}

template <int H_, int W_, bool RAW = false>
static int launch_resconv_shape(const ResConvArgs& a, cudaStream_t st) {
  using T = Tile<H_, W_>;
  constexpr bool two = resconv_smem_bytes(T::A_PART, 2, T::NWARP, T::G) <= 227 * 1024 && 2 * T::MT * 128 <= 512;
  if constexpr (two) {
    const int NST = (a.N + T::G - 1) / T::G;
    if (NST > kMaxGrid) return launch_resconv_slots<H_, W_, 2, RAW>(a, st);
  }
  return launch_resconv_slots<H_, W_, 1, RAW>(a, st);
}

}  // namespace node

#define NODE_RESCONV_SHAPE_TU(H, W) \
  namespace node { int launch_resconv_##H##x##W(const ResConvArgs& a, cudaStream_t st) { return launch_resconv_shape<H, W>(a, st); } \
                   int launch_resconv_raw_##H##x##W(const ResConvArgs& a, cudaStream_t st) { return launch_resconv_shape<H, W, true>(a, st); } }
