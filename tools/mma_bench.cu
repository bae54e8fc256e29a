// Tuning aid (not part of the product): rate and correctness of the SS-mode tcgen05.mma stream of one conv job
// (9 taps x MT x 4 K-steps x {N128, N64}) for different shared-memory layouts of the A image.
//   variant 0: K-major un-swizzled A ([k-chunk][row][16 B]), tap = row offset in the descriptor (engine of r01b)
//   variant 1: K-major SWIZZLE_128B A ([row][128 B], chunk ^= row & 7 on ABSOLUTE rows), tap = row offset,
//              descriptor base_offset field = 0
//   variant 2: same as 1 with base_offset = (start >> 7) & 7
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-ode-features_b200/csrc -I include \
//        tools/mma_bench.cu -o tools/mma_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "node_b200.h"
#include "ptx.cuh"

using namespace node;

constexpr int kRows = 256, kHalo = 16, kR = kRows + 2 * kHalo;        // rows of the A image
constexpr int kLBO = kR * 16;
constexpr int kAPart = kR * 128;
constexpr int kBTile = 128 * 128;
constexpr uint32_t kIdN128 = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);
constexpr uint32_t kIdN64 = (1u << 4) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ float a_val(int row, int k) { return (float)(((row * 7 + k * 3) % 17) - 8) * 0.125f; }
__device__ __forceinline__ float b_val(int n, int k) { return (float)(((n * 5 + k) % 13) - 6) * 0.125f; }

__device__ __forceinline__ uint64_t desc_sw128_bo(uint32_t saddr, uint32_t bo) {
  return ptx::make_desc_sw128(saddr) | ((uint64_t)(bo & 7) << 49);
}

struct Res { long long clk; float maxerr; int bad; };

template <int variant, int split>
__global__ void __launch_bounds__(512, 1) k_bench(int mt_count_, int reps, int load, const int* offs_g, Res* res, const float* gbuf) {
  constexpr int mt_count = 2; (void)mt_count_;
  __shared__ int offs[9];
  if (threadIdx.x < 9) offs[threadIdx.x] = offs_g[threadIdx.x];
  __syncthreads();
  extern __shared__ uint8_t raw[];
  const uint32_t s0 = ptx::smem_u32(raw);
  const uint32_t al = (s0 + 1023u) & ~1023u;
  uint8_t* base = raw + (al - s0);
  const uint32_t sA = al, sB = al + 2 * kAPart;
  uint8_t* A = base; uint8_t* B = base + 2 * kAPart;
  uint32_t* misc = reinterpret_cast<uint32_t*>(base + 2 * kAPart + 4 * kBTile);
  const uint32_t bar = sB + 4 * kBTile + 64;
  const int tid = threadIdx.x;
  // fill A (part 0 = hi, part 1 = hi again: values only matter for part 0 in the check) and B
  for (int i = tid; i < kR * 64; i += blockDim.x) {
    const int row = i / 64, k = i % 64;
    const __half v = __float2half(a_val(row, k));
    size_t o;
    if (variant == 0) o = (size_t)(k >> 3) * kLBO + (size_t)row * 16 + (k & 7) * 2;
    else o = (size_t)row * 128 + ((((k >> 3) ^ (row & 7))) << 4) + (k & 7) * 2;
    *reinterpret_cast<__half*>(A + o) = v;
    *reinterpret_cast<__half*>(A + kAPart + o) = v;
  }
  for (int t = 0; t < 4; ++t)
    for (int i = tid; i < 128 * 64; i += blockDim.x) {
      const int n = i / 64, k = i % 64;
      const size_t o = (size_t)n * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
      *reinterpret_cast<__half*>(B + t * kBTile + o) = __float2half(b_val(n, k));
    }
  if (tid == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); misc[4] = 0; }
  if (tid < 32) ptx::tmem_alloc(ptx::smem_u32(misc), 512);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *misc;

  auto issue_tap = [&](int off, int bt, bool first_tap, bool do_split) {
    const uint64_t b0 = ptx::make_desc_sw128(sB + bt * kBTile);
#pragma unroll
    for (int mt = 0; mt < mt_count; ++mt) {
      const uint32_t d = tmem + mt * 128;
      const int row0 = kHalo + mt * 128 + off;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint64_t a_hi, a_lo;
        if (variant == 0) {
          a_hi = ptx::make_desc_nosw(sA + row0 * 16 + 2 * ks * kLBO, kLBO, 128);
          a_lo = ptx::make_desc_nosw(sA + kAPart + row0 * 16 + 2 * ks * kLBO, kLBO, 128);
        } else {
          const uint32_t st = sA + row0 * 128 + ks * 32;
          const uint32_t bo = variant == 2 ? ((st >> 7) & 7) : 0;
          a_hi = desc_sw128_bo(st, bo);
          a_lo = desc_sw128_bo(st + kAPart, bo);
        }
        const uint64_t bk = b0 + (uint64_t)((ks * 32) >> 4);
        const uint32_t first = (first_tap && ks == 0) ? 0u : 1u;
        if (do_split) {
          ptx::mma_f16_ss(d, a_hi, bk, kIdN128, first);
          ptx::mma_f16_ss(d, a_lo, bk, kIdN64, 1u);
        } else {
          ptx::mma_f16_ss(d, a_hi, bk, kIdN128, first);
        }
      }
    }
  };

  // ---- correctness: one tap per offset, N128 only, check D[m][n] = sum_k A[m+off][k] B[n][k]
  float maxerr = 0.f; int bad = 0; uint32_t phase = 0;
  for (int t = 0; t < 9; ++t) {
    if (tid == 0) { issue_tap(offs[t], t & 3, true, false); ptx::tc_commit(bar); }
    ptx::mbar_wait(bar, phase & 1); ++phase;
    ptx::tc_fence_after();
    if (tid < 128) {
      for (int mt = 0; mt < mt_count; ++mt) {
        const int m = mt * 128 + tid;
        for (int c0 = 0; c0 < 128; c0 += 8) {
          uint32_t v[8];
          ptx::tmem_ld8(tmem + ((uint32_t)((tid >> 5) * 32) << 16) + mt * 128 + c0, v);
          ptx::tc_wait_ld();
          for (int j = 0; j < 8; ++j) {
            float ref = 0.f;
            for (int k = 0; k < 64; ++k) ref += a_val(kHalo + m + offs[t], k) * b_val(c0 + j, k);
            const float e = fabsf(__uint_as_float(v[j]) - ref);
            if (e > maxerr) maxerr = e;
            if (e > 1e-3f) ++bad;
          }
        }
      }
    }
    ptx::tc_fence_before();
    __syncthreads();
  }

  // ---- rate: reps jobs of 9 taps
  __syncthreads();
  long long t0 = 0, t1 = 0;
  volatile float* spam = reinterpret_cast<volatile float*>(base);
  float sink = 0.f;
  if (tid == 0) {
    t0 = clock64();
    int o[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) o[t] = offs[t];
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int t = 0; t < 9; ++t) issue_tap(o[t], t & 3, t == 0, split != 0);
      ptx::tc_commit(bar);
      ptx::mbar_wait(bar, phase & 1); ++phase;
    }
    t1 = clock64();
    misc[4] = 1;
  } else if (load && tid >= 32) {
    // interference from the other warps while the MMAs run:
    // 1 = conflict-free LDS.128 stream, 2 = 8-way conflicting LDS.32, 3 = STS.128 stream, 4 = tcgen05.ld of the accumulators,
    // 5 = FFMA only (issue-slot pressure), 6 = global loads (L2 hits)
    volatile uint32_t* flag = misc + 4;
    const uint32_t lane_addr = al + 2 * kAPart + 4 * kBTile + 1024 + (tid & 31) * 16 + (tid >> 5) * 512;   // scratch after misc
    float f0 = 1.f, f1 = 2.f, f2 = 3.f, f3 = 4.f;
    while (*flag == 0) {
      if (load == 1) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint32_t a, b, c, d;
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(lane_addr));
          sink += __uint_as_float(a ^ b ^ c ^ d);
        }
      } else if (load == 2) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          uint32_t a;
          asm volatile("ld.shared.b32 %0, [%1];" : "=r"(a) : "r"(al + 2 * kAPart + 4 * kBTile + 1024 + ((tid & 31) & 7) * 128 + ((tid & 31) >> 3) * 4));
          sink += __uint_as_float(a);
        }
      } else if (load == 3) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
          asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(lane_addr), "r"(u) : "memory");
      } else if (load == 4) {
        if (tid < 160) {
          uint32_t v[8];
          ptx::tmem_ld8(tmem + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + 256 + (tid & 7) * 8, v);
          ptx::tc_wait_ld();
          sink += __uint_as_float(v[0] ^ v[7]);
        }
      } else if (load == 5) {
#pragma unroll
        for (int u = 0; u < 32; ++u) { f0 = fmaf(f0, 1.0001f, f1); f1 = fmaf(f1, 0.9999f, f2); f2 = fmaf(f2, 1.0002f, f3); f3 = fmaf(f3, 0.9998f, f0); }
      } else {
#pragma unroll
        for (int u = 0; u < 4; ++u) sink += ptx::ldg_ordered(gbuf + ((tid * 32 + u * 8192 + (int)sink) & 1048575));
      }
    }
    sink += f0 + f1 + f2 + f3;
  }
  __syncthreads();
  if (sink == 123.456f) misc[5] = 1;
  // reduce error over the CTA
  __shared__ float serr[512]; __shared__ int sbad[512];
  serr[tid] = maxerr; sbad[tid] = bad;
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < 512; ++i) { serr[0] = fmaxf(serr[0], serr[i]); sbad[0] += sbad[i]; }
    res[blockIdx.x].clk = (t1 - t0) / reps; res[blockIdx.x].maxerr = serr[0]; res[blockIdx.x].bad = sbad[0];
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  const size_t smem = 1024 + 2 * kAPart + 4 * kBTile + 1024 + 16 * 512;
  float* gbuf; cudaMalloc(&gbuf, 4 << 20); cudaMemset(gbuf, 0, 4 << 20);
  Res* res; cudaMalloc(&res, sizeof(Res) * 148);
  int* offs; cudaMalloc(&offs, 9 * sizeof(int));
  const int sets[4][9] = {{-8, -8, -8, 0, 0, 0, 8, 8, 8}, {-10, -9, -8, -1, 0, 1, 8, 9, 10}, {-9, -8, -7, -1, 0, 1, 7, 8, 9}, {0, 0, 0, 0, 0, 0, 0, 0, 0}};
  const char* setname[4] = {"aligned8", "pitch9", "pitch8", "zero"};
  printf("floor per job (MT=2): split %d clk, hi-only %d clk\n", 9 * 2 * 4 * 96, 9 * 2 * 4 * 64);
  for (int variant = 0; variant < 2; ++variant)
    for (int s = 1; s < 2; ++s)
      for (int split = 0; split < 2; ++split)
        for (int load = 0; load < 7; ++load) {
          cudaMemcpy(offs, sets[s], sizeof(int) * 9, cudaMemcpyHostToDevice);
          auto go = [&](auto kern) { cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); kern<<<148, 512, smem>>>(2, 20, load, offs, res, gbuf); };
          if (variant == 0 && split == 0) go(k_bench<0, 0>);
          if (variant == 0 && split == 1) go(k_bench<0, 1>);
          if (variant == 1 && split == 0) go(k_bench<1, 0>);
          if (variant == 1 && split == 1) go(k_bench<1, 1>);
          if (variant == 2 && split == 0) go(k_bench<2, 0>);
          if (variant == 2 && split == 1) go(k_bench<2, 1>);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("variant %d set %s: %s\n", variant, setname[s], cudaGetErrorString(e)); return 1; }
          Res h[148]; cudaMemcpy(h, res, sizeof(h), cudaMemcpyDeviceToHost);
          long long mn = h[0].clk, mx = h[0].clk; float me = 0; int bad = 0;
          for (int i = 0; i < 148; ++i) { mn = h[i].clk < mn ? h[i].clk : mn; mx = h[i].clk > mx ? h[i].clk : mx; me = h[i].maxerr > me ? h[i].maxerr : me; bad += h[i].bad; }
          printf("variant %d offsets %-9s split %d smem-load %d : clk/job min %6lld max %6lld  | check maxerr %.3g bad %d\n", variant,
                 setname[s], split, load, mn, mx, me, bad);
        }
  return 0;
}
