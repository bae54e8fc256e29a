// the head of the reference's ResBlock (model.py:156-178) for its strided instances,
//     c  = conv1(a)        conv1      = Conv2d(64, 64, 3, stride 2, padding 1, bias=False)
//     sc = downsample(a)   downsample = Conv2d(64, 64, 1, stride 2, bias=False)           a = relu(norm1(x))
// as one tcgen05 kernel. A stride-2 3x3 convolution reads, for output pixel (i, j), the inputs (2i+dy, 2j+dx): split the
// input into its four PARITY PLANES p[pr][pc](i, j) = a(2i+pr, 2j+pc) and every tap becomes a stride-1 tap of one
// plane - (dy, dx) = (2 bi + pr, 2 bj + pc) with block offsets bi, bj in {-1, 0}. Each plane is laid out exactly like the
// step engine's position strips over the OUTPUT positions (zero column / zero row shared between rows / images), so a
// tap is again nothing but a row offset in the A descriptor; the 1x1 stride-2 shortcut is one more tap on plane (0, 0)
// with its own weights and its own accumulator columns.
// Four planes x (hi, lo) x 64 channels do not fit in shared memory at once: two passes of two planes each accumulate
// into the same tensor-memory columns (pass 1: planes (1,1), (1,0) = 6 taps; pass 2: planes (0,1), (0,0) = 3 taps +
// shortcut). fp32 contract by fp16 operand splitting as in step_engine.cuh; weight tiles stream through the same ring.
#pragma once
#include "step_engine.cuh"

namespace node {

constexpr int kS2Tiles = 10;       // weight tiles of a job in issue order (see k_convs2_tiles)

struct ConvS2Args {
  const uint16_t* w16;       // [10][128 rows][64 halves]
  const float* scal;         // [0] activation scale, [2] 1/(sa*sw1), [3] 1/(sa*swd)
  const float* act; float* c_out; float* sc_out;
  int N;
};

__host__ __device__ constexpr size_t convs2_smem_bytes(int A_PART) {
  return 1024 + (size_t)kNW * kW16TileBytes + (size_t)4 * A_PART + 16 * 8 + 64;
}

// x (already relu'd, >= 0) * scale split into fp16 hi + lo -> row `row` of a plane image, k-chunks [4*hb, 4*hb+4)
template <class T>
__device__ __forceinline__ void raw_to_A(uint32_t row, const float (&x)[32], float scale) {
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
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + kc * T::LBO), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(row + T::A_PART + kc * T::LBO), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
  }
}

template <int HO, int WO, int HI, int WI>
__global__ void __launch_bounds__(Tile<HO, WO>::P, 1) k_convs2(const ConvS2Args a) {
  using T = Tile<HO, WO>;
  static_assert(T::MT == 2, "accumulators: 2 M tiles x (128 + 128) columns");
  constexpr int HWO = HO * WO, HWI = HI * WI, P = T::P;
  extern __shared__ uint8_t smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t s0 = ptx::smem_u32(smem_raw);
  const uint32_t al = (s0 + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (al - s0);
  size_t o = 0;
  const uint32_t wring = al; o += (size_t)kNW * kW16TileBytes;
  const uint32_t planes = al + (uint32_t)o; o += (size_t)4 * T::A_PART;      // plane slot s: [hi | lo] at s * 2 * A_PART
  const uint32_t bar_wfull = al + (uint32_t)o; o += 8 * kNW;
  const uint32_t bar_wfree = al + (uint32_t)o; o += 8 * kNW;
  const uint32_t bar_pass = al + (uint32_t)o; o += 8;
  const uint32_t bar_acc = al + (uint32_t)o; o += 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base + o);
  {
    uint4* az = reinterpret_cast<uint4*>(base + (size_t)kNW * kW16TileBytes);   // padding rows / columns stay zero
    for (int i = tid; i < 4 * T::A_PART / 16; i += blockDim.x) az[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  if (tid == 0) {
    for (int i = 0; i < kNW; ++i) { ptx::mbar_init(bar_wfull + 8 * i, 1); ptx::mbar_init(bar_wfree + 8 * i, 1); }
    ptx::mbar_init(bar_pass, 1); ptx::mbar_init(bar_acc, 1);
    ptx::fence_mbar_init();
  }
  if (tid < 32) ptx::tmem_alloc(ptx::smem_u32(tmem_slot), kTmemCols);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int NST = (a.N + T::G - 1) / T::G;
  const int njobs = (int)blockIdx.x < NST ? (NST - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const uint32_t total = (uint32_t)njobs * kS2Tiles;
  auto request = [&](uint32_t i) {
    const uint32_t slot = i % kNW;
    ptx::mbar_expect_tx(bar_wfull + 8 * slot, kW16TileBytes);
    ptx::bulk_g2s(wring + slot * kW16TileBytes, (const char*)a.w16 + (size_t)(i % kS2Tiles) * kW16TileBytes, kW16TileBytes,
                  bar_wfull + 8 * slot);
  };
  if (tid == 0)
    for (uint32_t i = 0; i < kWAhead && i < total; ++i) request(i);

  // this thread's output position
  const int img_l = tid / T::IS, rr = tid % T::IS, oi = rr / T::Wp, oj = rr % T::Wp;
  const bool inimg = img_l < T::G && oi < HO && oj < WO;
  const float sa = a.scal[0];
  bool timeout = false;
  // (plane slot, pr, pc) of the two passes and their taps (bi, bj) in weight-tile order
  // pass 0: slot 0 = plane (1,1): (-1,-1) (-1,0) (0,-1) (0,0); slot 1 = plane (1,0): (-1,0) (0,0)
  // pass 1: slot 0 = plane (0,1): (0,-1) (0,0);               slot 1 = plane (0,0): (0,0) and the shortcut (0,0)
  // Even widths: the first 32 channels of the NEXT staging unit (next pass, or pass 0 of the next image) are requested
  // right after a pass' MMAs have been issued, so their DRAM latency hides behind the MMAs / the epilogue.
  float2 pf[32];
  bool pf_inb = false;
  auto prefetch = [&](int jn2, int pass2) {
    if (jn2 >= njobs) return;
    const int img2 = (blockIdx.x + jn2 * gridDim.x) * T::G + img_l;
    const int ii2 = 2 * oi + (pass2 == 0 ? 1 : 0);
    pf_inb = inimg && img2 < a.N && ii2 < HI;
    const size_t g2 = pf_inb ? ((size_t)img2 * kC * HWI + (size_t)ii2 * WI + 2 * oj) : 0;
#pragma unroll
    for (int c = 0; c < 32; ++c)
      asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(pf[c].x), "=f"(pf[c].y) : "l"(a.act + g2 + (size_t)c * HWI));
  };
  if constexpr (WI % 2 == 0) prefetch(0, 0);
#pragma unroll 1
  for (int jn = 0; jn < njobs; ++jn) {
    const int st = blockIdx.x + jn * gridDim.x;
    const int img = st * T::G + img_l;
    const bool valid = inimg && img < a.N;
    const uint32_t job = (uint32_t)jn;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      // ---- stage the two planes of this pass (loads first, then wait until the previous pass' MMAs have read the images)
      if (pass == 1 && !timeout && !ptx::mbar_wait_relaxed(bar_pass, job & 1)) timeout = true;
      if constexpr (WI % 2 == 0) {
        // even width: the two column parities of a row pair are adjacent floats - one 8-byte load feeds both planes
        const int pr = pass == 0 ? 1 : 0;
        const int ii = 2 * oi + pr;
        const bool inb = valid && ii < HI;
        const size_t g0 = inb ? ((size_t)img * kC * HWI + (size_t)ii * WI + 2 * oj) : 0;
        const uint32_t row = planes + (uint32_t)(T::HALO + tid) * 16;
#pragma unroll 1
        for (int hb = 0; hb < 2; ++hb) {
          float x0[32], x1[32];
          if (hb == 0) {
#pragma unroll
            for (int c = 0; c < 32; ++c) { x0[c] = pf_inb ? pf[c].x : 0.f; x1[c] = pf_inb ? pf[c].y : 0.f; }
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              float2 v;
              asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(a.act + g0 + (size_t)(32 * hb + c) * HWI));
              x0[c] = inb ? v.x : 0.f; x1[c] = inb ? v.y : 0.f;
            }
          }
          if (valid) {
            raw_to_A<T>(row + 4 * hb * T::LBO, x1, sa);                          // plane slot 0: pc = 1
            raw_to_A<T>(row + 2 * T::A_PART + 4 * hb * T::LBO, x0, sa);          // plane slot 1: pc = 0
          }
        }
      } else {
#pragma unroll 1
        for (int ps = 0; ps < 2; ++ps) {
          const int pr = pass == 0 ? 1 : 0, pc = ps == 0 ? 1 : 0;
          const int ii = 2 * oi + pr, ij = 2 * oj + pc;
          const bool inb = valid && ii < HI && ij < WI;
          const size_t g0 = inb ? ((size_t)img * kC * HWI + (size_t)ii * WI + ij) : 0;
          const uint32_t row = planes + (uint32_t)ps * 2 * T::A_PART + (uint32_t)(T::HALO + tid) * 16;
#pragma unroll 1
          for (int hb = 0; hb < 2; ++hb) {
            float x[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] = ptx::ldg_ordered(a.act + g0 + (size_t)(32 * hb + c) * HWI);
            if (!inb) {
#pragma unroll
              for (int c = 0; c < 32; ++c) x[c] = 0.f;
            }
            if (valid) raw_to_A<T>(row + 4 * hb * T::LBO, x, sa);
          }
        }
      }
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncthreads();
      // ---- leader: the MMA stream of this pass; producer: refill the ring
      if (__shfl_sync(0xffffffffu, warp, 0) == 0) {
        const bool lead = ptx::elect_one();
        ptx::tc_fence_after();
        const int ntile = pass == 0 ? 6 : 4;
#pragma unroll 1
        for (int q = 0; q < ntile; ++q) {
          const int lt = pass == 0 ? q : 6 + q;                       // local tile index
          const uint32_t tile = job * kS2Tiles + lt, slot = tile % kNW;
          if (!timeout && !ptx::mbar_wait(bar_wfull + 8 * slot, (tile / kNW) & 1)) timeout = true;
          ptx::tc_fence_after();
          // tap geometry
          int ps, bi, bj;
          if (pass == 0) { ps = q < 4 ? 0 : 1; bi = q < 4 ? (q < 2 ? -1 : 0) : (q == 4 ? -1 : 0); bj = q < 4 ? ((q & 1) ? 0 : -1) : 0; }
          else { ps = q < 2 ? 0 : 1; bi = 0; bj = q == 0 ? -1 : 0; }
          const bool shortcut = lt == 9;
          const int off = bi * T::Wp + bj;
          // descriptors as low-word adds (see issue_conv_job in step_engine.cuh)
          constexpr uint32_t a_hiw = (128u >> 4) | (1u << 14), b_hiw = (1024u >> 4) | (1u << 14) | (2u << 29);
          auto pack = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
          const uint32_t a_tap = ((((planes + (uint32_t)ps * 2 * T::A_PART + (uint32_t)(T::HALO * 16)) & 0x3FFFFu) >> 4) |
                                  (((uint32_t)T::LBO >> 4) << 16)) + (uint32_t)off;
          const uint32_t b_lo0 = ((wring + slot * kW16TileBytes) & 0x3FFFFu) >> 4;
          if (lead) {
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const uint32_t d = tmem + (uint32_t)(mt * 256 + (shortcut ? 128 : 0));
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                const uint64_t a_hi = pack(a_tap + (uint32_t)((mt * 128 * 16 + 2 * ks * T::LBO) >> 4), a_hiw);
                const uint64_t a_lo = pack(a_tap + (uint32_t)((mt * 128 * 16 + 2 * ks * T::LBO + T::A_PART) >> 4), a_hiw);
                const uint64_t bk = pack(b_lo0 + (uint32_t)((ks * 32) >> 4), b_hiw);
                const uint32_t first = ((lt == 0 || shortcut) && ks == 0) ? 0u : 1u;
                ptx::mma_f16_ss(d, a_hi, bk, kIdF16N128, first);
                ptx::mma_f16_ss(d, a_lo, bk, kIdF16N64, 1u);
              }
            }
            ptx::tc_commit(bar_wfree + 8 * slot);
          }
        }
        if (lead) ptx::tc_commit(pass == 0 ? bar_pass : bar_acc);
        __syncwarp();
      } else if (__shfl_sync(0xffffffffu, warp, 0) == 1) {
        const bool lead = ptx::elect_one();
        // pass 0: local tiles 3..9 (their predecessors in the ring belong to pass 0 or the previous job);
        // pass 1: the next job's first tiles (predecessors: this pass)
        const uint32_t lo_t = job * kS2Tiles + (pass == 0 ? kWAhead : (uint32_t)kS2Tiles);
        const uint32_t hi_t = job * kS2Tiles + (pass == 0 ? (uint32_t)kS2Tiles : (uint32_t)kS2Tiles + kWAhead);
#pragma unroll 1
        for (uint32_t i = lo_t; i < hi_t && i < total; ++i) {
          const uint32_t slot = i % kNW;
          if (i >= (uint32_t)kNW && !timeout && !ptx::mbar_wait(bar_wfree + 8 * slot, ((i / kNW) - 1) & 1)) timeout = true;
          if (lead) request(i);
        }
        __syncwarp();
      }
      if constexpr (WI % 2 == 0) prefetch(pass == 0 ? jn : jn + 1, pass == 0 ? 1 : 0);
    }
    // ---- epilogue
    if (!timeout && !ptx::mbar_wait_relaxed(bar_acc, job & 1)) timeout = true;
    ptx::tc_fence_after();
    {
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((tid >> 7) * 256);
      const size_t gout = valid ? ((size_t)img * kC * HWO + (size_t)oi * WO + oj) : 0;
#pragma unroll 1
      for (int part = 0; part < 2; ++part) {           // 0: conv1 -> c_out, 1: shortcut -> sc_out
        float* dst = part == 0 ? a.c_out : a.sc_out;
        const float inv = part == 0 ? a.scal[2] : a.scal[3];
#pragma unroll 1
        for (int c0 = 0; c0 < 64; c0 += 8) {
          uint32_t v0[8], v1[8];
          ptx::tmem_ld8(taddr + part * 128 + c0, v0);
          ptx::tmem_ld8(taddr + part * 128 + 64 + c0, v1);
          ptx::tc_wait_ld();
          if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[gout + (size_t)(c0 + j) * HWO] = (__uint_as_float(v0[j]) + __uint_as_float(v1[j])) * inv;
          }
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();       // accumulators and plane images are free for the next job
    ptx::tc_fence_after();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, kTmemCols);
}

template <int HO, int WO, int HI, int WI>
static int launch_convs2_shape(const ConvS2Args& a, cudaStream_t st) {
  using T = Tile<HO, WO>;
  constexpr size_t smem = convs2_smem_bytes(T::A_PART);
  static_assert(smem <= 227 * 1024, "shared memory budget");
  NODE_SET_SMEM_ONCE((k_convs2<HO, WO, HI, WI>), smem);
  const int NST = (a.N + T::G - 1) / T::G;
  const int grid = NST < kMaxGrid ? NST : kMaxGrid;
  k_convs2<HO, WO, HI, WI><<<grid, T::P, smem, st>>>(a);
  return (int)cudaGetLastError();
}

}  // namespace node

#define NODE_CONVS2_SHAPE_TU(HO, WO, HI, WI) \
  namespace node { int launch_convs2_##HI##x##WI(const ConvS2Args& a, cudaStream_t st) { return launch_convs2_shape<HO, WO, HI, WI>(a, st); } }
