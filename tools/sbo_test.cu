// Tuning aid (not part of the product): is a K-major un-swizzled A descriptor with a stride byte offset that is NOT a
// multiple of 128 legal, and how fast is it? Layout under test = the dense 8x8 tiling of step8_engine.cuh:
//   row-slot = 8 pixels of one image row + one zero entry (9 x 16 B = 144 B = SBO), the rows of two images interleaved
//   (A0 B0 A1 B1 ...), two zero slots between M tiles; a 3x3 tap (dy, dx) is the byte offset dy*288 + dx*16.
// Checks D[m][n] = sum_k A[nbr(m, tap)][k] B[n][k] for all 9 taps against a scalar reference and times a whole conv job
// (9 taps x 2 M tiles x 4 K steps x {N128, N64}).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-ode-features_b200/csrc -I include \
//        tools/sbo_test.cu -o tools/sbo_test
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "node_b200.h"
#include "ptx.cuh"

using namespace node;

constexpr int kSlot = 144, kSlotsPerChunk = 36, kLBO = kSlotsPerChunk * kSlot;   // 5184
constexpr int kAPart = 8 * kLBO;
constexpr int kLead = kSlot;                       // one zero slot in front of the image
constexpr int kABytes = kLead + 2 * kAPart + 2 * kSlot;
constexpr int kBTile = 128 * 128;
constexpr uint32_t kIdN128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdN64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

// logical activation of (tile mt, image i, row r, col c, channel k); zero outside the image
__device__ __forceinline__ float a_val(int mt, int img, int r, int c, int k) {
  if (r < 0 || r > 7 || c < 0 || c > 7) return 0.f;
  return (float)((((mt * 2 + img) * 64 + r * 8 + c) * 7 + k * 3) % 17 - 8) * 0.125f;
}
__device__ __forceinline__ float b_val(int n, int k) { return (float)(((n * 5 + k) % 13) - 6) * 0.125f; }

struct Res { long long clk; float maxerr; int bad; };

__global__ void __launch_bounds__(256, 1) k_sbo(int reps, Res* res) {
  extern __shared__ uint8_t raw[];
  const uint32_t s0 = ptx::smem_u32(raw);
  const uint32_t al = (s0 + 1023u) & ~1023u;
  uint8_t* base = raw + (al - s0);
  uint8_t* B = base;                       // 4 weight tiles, 1024-aligned
  uint8_t* A = base + 4 * kBTile;
  const uint32_t sB = al, sA = al + 4 * kBTile;
  uint32_t* misc = reinterpret_cast<uint32_t*>(base + 4 * kBTile + ((kABytes + 127) & ~127));
  const uint32_t bar = sA + ((kABytes + 127) & ~127) + 64;
  const int tid = threadIdx.x;
  for (int i = tid; i < kABytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(A)[i] = 0u;
  __syncthreads();
  // thread = position p of the 256-row super-tile
  {
    const int mt = tid >> 7, p = tid & 127, s = p >> 3, e = p & 7, img = s & 1, r = s >> 1;
    for (int k = 0; k < 64; ++k) {
      const size_t o = (size_t)kLead + (size_t)(k >> 3) * kLBO + (size_t)(2 + mt * 18 + s) * kSlot + e * 16 + (k & 7) * 2;
      const __half v = __float2half(a_val(mt, img, r, e, k));
      *reinterpret_cast<__half*>(A + o) = v;
      *reinterpret_cast<__half*>(A + kAPart + o) = v;
    }
  }
  for (int t = 0; t < 4; ++t)
    for (int i = tid; i < 128 * 64; i += blockDim.x) {
      const int n = i / 64, k = i % 64;
      const size_t o = (size_t)n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
      *reinterpret_cast<__half*>(B + t * kBTile + o) = __float2half(b_val(n, k));
    }
  if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (tid < 32) ptx::tmem_alloc(ptx::smem_u32(misc), 512);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *misc;

  auto issue_tap = [&](int tap, int bt, bool first_tap, bool do_split) {
    const uint64_t b0 = ptx::make_desc_sw128(sB + bt * kBTile);
    const int off = (tap / 3 - 1) * 2 * kSlot + (tap % 3 - 1) * 16;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const uint32_t d = tmem + mt * 128;
      const uint32_t st = sA + kLead + (2 + mt * 18) * kSlot + off;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t a_hi = ptx::make_desc_nosw(st + 2 * ks * kLBO, kLBO, kSlot);
        const uint64_t a_lo = ptx::make_desc_nosw(st + kAPart + 2 * ks * kLBO, kLBO, kSlot);
        const uint64_t bk = b0 + (uint64_t)((ks * 32) >> 4);
        const uint32_t first = (first_tap && ks == 0) ? 0u : 1u;
        ptx::mma_f16_ss(d, a_hi, bk, kIdN128, first);
        if (do_split) ptx::mma_f16_ss(d, a_lo, bk, kIdN64, 1u);
      }
    }
  };

  float maxerr = 0.f; int bad = 0; uint32_t phase = 0;
  for (int tap = 0; tap < 9; ++tap) {
    if (tid == 0) { issue_tap(tap, tap & 3, true, false); ptx::tc_commit(bar); }
    ptx::mbar_wait(bar, phase & 1); ++phase;
    ptx::tc_fence_after();
    {
      const int mt = tid >> 7, p = tid & 127, s = p >> 3, e = p & 7, img = s & 1, r = s >> 1;
      const int dy = tap / 3 - 1, dx = tap % 3 - 1;
      for (int c0 = 0; c0 < 128; c0 += 8) {
        uint32_t v[8];
        ptx::tmem_ld8(tmem + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + mt * 128 + c0, v);
        ptx::tc_wait_ld();
        for (int j = 0; j < 8; ++j) {
          float ref = 0.f;
          for (int k = 0; k < 64; ++k) ref += a_val(mt, img, r + dy, e + dx, k) * b_val(c0 + j, k);
          const float err = fabsf(__uint_as_float(v[j]) - ref);
          if (err > maxerr) maxerr = err;
          if (err > 1e-3f) ++bad;
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
  }
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int t = 0; t < 9; ++t) issue_tap(t, t & 3, t == 0, true);
      ptx::tc_commit(bar);
      ptx::mbar_wait(bar, phase & 1); ++phase;
    }
    t1 = clock64();
  }
  __shared__ float serr[256]; __shared__ int sbad[256];
  serr[tid] = maxerr; sbad[tid] = bad;
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < 256; ++i) { serr[0] = fmaxf(serr[0], serr[i]); sbad[0] += sbad[i]; }
    res[blockIdx.x].clk = (t1 - t0) / reps; res[blockIdx.x].maxerr = serr[0]; res[blockIdx.x].bad = sbad[0];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  const size_t smem = 1024 + 4 * kBTile + ((kABytes + 127) & ~127) + 256;
  Res* res; cudaMalloc(&res, sizeof(Res) * 148);
  cudaFuncSetAttribute(k_sbo, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_sbo<<<148, 256, smem>>>(20, res);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("sbo_test: %s\n", cudaGetErrorString(e)); return 1; }
  Res h[148]; cudaMemcpy(h, res, sizeof(h), cudaMemcpyDeviceToHost);
  long long mn = h[0].clk, mx = h[0].clk; float me = 0; int bad = 0;
  for (int i = 0; i < 148; ++i) { mn = h[i].clk < mn ? h[i].clk : mn; mx = h[i].clk > mx ? h[i].clk : mx; me = h[i].maxerr > me ? h[i].maxerr : me; bad += h[i].bad; }
  printf("dense 8x8 layout, SBO=144: split conv job clk min %lld max %lld (math floor %d) | check maxerr %.3g bad %d  smem %zu\n", mn, mx,
         9 * 2 * 4 * 96, me, bad, smem);
  return bad != 0;
}
