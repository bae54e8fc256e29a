// Tuning aid (not part of the product): feasibility of driving the dense 8x8 conv job with tcgen05.mma.cta_group::2 on a
// CTA pair - A images written by ordinary stores in EACH CTA (generic proxy), the weight tile split between the two CTAs'
// shared memories, the peer CTA's "A image ready" arriving on the leader's mbarrier through DSMEM, completion multicast
// to both CTAs. Layout of the weight half-tiles (64 rows x 64 cin halves, SWIZZLE_128B) so that all three products of
// the fp16 split land in the right accumulator columns:
//   CTA 0: rows 0..31 = w_hi[0..31],  rows 32..63 = w_lo[32..63]
//   CTA 1: rows 0..31 = w_hi[32..63], rows 32..63 = w_lo[0..31]
//   MMA 1 (N = 128): a_hi x [CTA0 rows | CTA1 rows] -> columns [hi 0-31 | lo 32-63 | hi 32-63 | lo 0-31]
//   MMA 2 (N = 64):  a_lo x [CTA0 rows 0..31 | CTA1 rows 0..31] = a_lo x w_hi[0..63] -> columns 0..63
//   output channel n = column n + column (n < 32 ? 96 + n : 32 + n)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-ode-features_b200/csrc -I include \
//        tools/pair_test.cu -o tools/pair_test
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "node_b200.h"
#include "ptx.cuh"

using namespace node;

constexpr int kSlot = 144, kLBO = 36 * kSlot, kAPart = 8 * kLBO, kLead = kSlot;
constexpr int kABytes = kLead + 2 * kAPart + 2 * kSlot;
constexpr int kBHalf = 64 * 128;                                         // one CTA's half of a tap's weight tile
constexpr uint32_t kIdN128M256 = (1u << 4) | ((128u >> 3) << 17) | ((256u >> 4) << 24);
constexpr uint32_t kIdN64M256 = (1u << 4) | ((64u >> 3) << 17) | ((256u >> 4) << 24);

__device__ __forceinline__ float a_hi_val(int cta, int mt, int img, int r, int c, int k) {
  if (r < 0 || r > 7 || c < 0 || c > 7) return 0.f;
  return (float)(((((cta * 2 + mt) * 2 + img) * 64 + r * 8 + c) * 7 + k * 3) % 17 - 8) * 0.125f;
}
__device__ __forceinline__ float a_lo_val(int cta, int mt, int img, int r, int c, int k) {
  if (r < 0 || r > 7 || c < 0 || c > 7) return 0.f;
  return (float)(((((cta * 2 + mt) * 2 + img) * 64 + r * 8 + c) * 5 + k) % 11 - 5) * 0.0625f;
}
__device__ __forceinline__ float w_hi_val(int tap, int n, int k) { return (float)(((n * 5 + k + tap) % 13) - 6) * 0.125f; }
__device__ __forceinline__ float w_lo_val(int tap, int n, int k) { return (float)(((n * 3 + k * 7 + tap) % 9) - 4) * 0.03125f; }

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t local, uint32_t cta) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(cta)); return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return true;
  }
  return false;
}
__device__ __forceinline__ void mma2_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit2(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

struct Res { long long clk; float maxerr; int bad; int timeout; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(256, 1) k_pair(int reps, Res* res) {
  extern __shared__ uint8_t raw[];
  const uint32_t s0 = ptx::smem_u32(raw);
  const uint32_t al = (s0 + 1023u) & ~1023u;
  uint8_t* base = raw + (al - s0);
  uint8_t* B = base;                                        // 9 half-tiles, 1024-aligned
  uint8_t* A = base + 9 * kBHalf;
  const uint32_t sB = al, sA = al + 9 * kBHalf;
  const int aoff = 9 * kBHalf + ((kABytes + 127) & ~127);
  uint32_t* misc = reinterpret_cast<uint32_t*>(base + aoff);
  const uint32_t bar_ready = al + aoff + 64, bar_acc = al + aoff + 72;
  const int tid = threadIdx.x;
  const int cta = (int)cluster_ctarank();
  for (int i = tid; i < kABytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(A)[i] = 0u;
  __syncthreads();
  {
    const int mt = tid >> 7, p = tid & 127, s = p >> 3, e = p & 7, img = s & 1, r = s >> 1;
    for (int k = 0; k < 64; ++k) {
      const size_t o = (size_t)kLead + (size_t)(k >> 3) * kLBO + (size_t)(2 + mt * 18 + s) * kSlot + e * 16 + (k & 7) * 2;
      *reinterpret_cast<__half*>(A + o) = __float2half(a_hi_val(cta, mt, img, r, e, k));
      *reinterpret_cast<__half*>(A + kAPart + o) = __float2half(a_lo_val(cta, mt, img, r, e, k));
    }
  }
  for (int t = 0; t < 9; ++t)
    for (int i = tid; i < 64 * 64; i += blockDim.x) {
      const int row = i / 64, k = i % 64;
      float v;
      if (cta == 0) v = row < 32 ? w_hi_val(t, row, k) : w_lo_val(t, row, k);
      else v = row < 32 ? w_hi_val(t, 32 + row, k) : w_lo_val(t, row - 32, k);
      const size_t o = (size_t)row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2;
      *reinterpret_cast<__half*>(B + t * kBHalf + o) = __float2half(v);
    }
  if (tid == 0) { ptx::mbar_init(bar_ready, 2 * 8); ptx::mbar_init(bar_acc, 1); ptx::fence_mbar_init(); }
  if (tid < 32) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(misc)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();                      // both CTAs' barriers are initialised before anyone arrives remotely
  ptx::tc_fence_after();
  const uint32_t tmem = *misc;

  constexpr uint32_t a_hiw = ((uint32_t)kSlot >> 4) | (1u << 14);
  constexpr uint32_t b_hiw = (1024u >> 4) | (1u << 14) | (2u << 29);
  auto pack = [](uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | (uint64_t)lo; };
  auto issue_job = [&]() {
    const uint32_t a_lo0 = (((sA + kLead + 2 * kSlot) & 0x3FFFFu) >> 4) | (((uint32_t)kLBO >> 4) << 16);
    for (int tap = 0; tap < 9; ++tap) {
      const int off = (tap / 3 - 1) * 2 * kSlot + (tap % 3 - 1) * 16;
      const uint32_t a_tap = a_lo0 + (uint32_t)(off >> 4);
      const uint32_t b_lo0 = ((sB + tap * kBHalf) & 0x3FFFFu) >> 4;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const uint32_t d = tmem + (uint32_t)(mt * 128);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t a_hi = pack(a_tap + (uint32_t)((mt * 18 * kSlot + 2 * ks * kLBO) >> 4), a_hiw);
          const uint64_t a_lo = pack(a_tap + (uint32_t)((mt * 18 * kSlot + 2 * ks * kLBO + kAPart) >> 4), a_hiw);
          const uint64_t bk = pack(b_lo0 + (uint32_t)((ks * 32) >> 4), b_hiw);
          mma2_f16_ss(d, a_hi, bk, kIdN128M256, (tap == 0 && ks == 0) ? 0u : 1u);
          mma2_f16_ss(d, a_lo, bk, kIdN64M256, 1u);
        }
      }
    }
  };

  // "A image ready": one arrive per warp of BOTH CTAs on the leader's barrier
  int timeout = 0;
  uint32_t phase = 0;
  const uint32_t leader_ready = mapa(bar_ready, 0);
  __syncwarp();
  if ((tid & 31) == 0) mbar_arrive_remote(leader_ready);
  if (cta == 0 && tid == 0) {
    if (!mbar_wait_cluster(bar_ready, 0)) timeout = 1;
    ptx::tc_fence_after();
    issue_job();
    commit2(bar_acc, 3);
  }
  if (!ptx::mbar_wait(bar_acc, phase & 1)) timeout |= 2;
  ++phase;
  ptx::tc_fence_after();

  float maxerr = 0.f; int bad = 0;
  {
    const int mt = tid >> 7, p = tid & 127, s = p >> 3, e = p & 7, img = s & 1, r = s >> 1;
    for (int c0 = 0; c0 < 64; c0 += 8) {
      uint32_t v0[8], v1[8];
      const uint32_t lane = (uint32_t)(((tid >> 5) & 3) * 32) << 16;
      ptx::tmem_ld8(tmem + lane + mt * 128 + c0, v0);
      ptx::tmem_ld8(tmem + lane + mt * 128 + (c0 < 32 ? 96 + c0 : 32 + c0), v1);
      ptx::tc_wait_ld();
      for (int j = 0; j < 8; ++j) {
        const int n = c0 + j;
        float ref = 0.f;
        for (int tap = 0; tap < 9; ++tap) {
          const int dy = tap / 3 - 1, dx = tap % 3 - 1;
          for (int k = 0; k < 64; ++k) {
            const float ah = a_hi_val(cta, mt, img, r + dy, e + dx, k), alo = a_lo_val(cta, mt, img, r + dy, e + dx, k);
            ref += ah * w_hi_val(tap, n, k) + ah * w_lo_val(tap, n, k) + alo * w_hi_val(tap, n, k);
          }
        }
        const float got = __uint_as_float(v0[j]) + __uint_as_float(v1[j]);
        const float err = fabsf(got - ref);
        if (err > maxerr) maxerr = err;
        if (err > 2e-2f) ++bad;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();

  // ---- rate: reps jobs back to back (operands unchanged)
  long long t0 = 0, t1 = 0;
  if (cta == 0 && tid == 0) {
    t0 = clock64();
    for (int r = 0; r < reps; ++r) issue_job();
    commit2(bar_acc, 3);
  }
  if (!ptx::mbar_wait(bar_acc, phase & 1)) timeout |= 4;
  if (cta == 0 && tid == 0) t1 = clock64();
  ++phase;

  __shared__ float serr[256]; __shared__ int sbad[256]; __shared__ int sto[256];
  serr[tid] = maxerr; sbad[tid] = bad; sto[tid] = timeout;
  __syncthreads();
  if (tid == 0) {
    for (int i = 1; i < 256; ++i) { serr[0] = fmaxf(serr[0], serr[i]); sbad[0] += sbad[i]; sto[0] |= sto[i]; }
    res[blockIdx.x].clk = reps > 0 ? (t1 - t0) / reps : 0; res[blockIdx.x].maxerr = serr[0]; res[blockIdx.x].bad = sbad[0]; res[blockIdx.x].timeout = sto[0];
  }
  ptx::tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
  const size_t smem = 1024 + 9 * kBHalf + ((kABytes + 127) & ~127) + 256;
  Res* res; cudaMalloc(&res, sizeof(Res) * 148); cudaMemset(res, 0, sizeof(Res) * 148);
  cudaFuncSetAttribute(k_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_pair<<<148, 256, smem>>>(20, res);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("pair_test: %s\n", cudaGetErrorString(e)); return 1; }
  Res h[148]; cudaMemcpy(h, res, sizeof(h), cudaMemcpyDeviceToHost);
  long long mn = 1ll << 60, mx = 0; float me = 0; int bad = 0, to = 0;
  for (int i = 0; i < 148; ++i) {
    if (i % 2 == 0) { mn = h[i].clk < mn ? h[i].clk : mn; mx = h[i].clk > mx ? h[i].clk : mx; }
    me = h[i].maxerr > me ? h[i].maxerr : me; bad += h[i].bad; to |= h[i].timeout;
  }
  printf("cta_group::2 dense 8x8 conv job (pair = 8 images): clk/job min %lld max %lld (math floor %d) | check maxerr %.3g bad %d timeout %d  smem %zu\n",
         mn, mx, 9 * 2 * 4 * 96, me, bad, to, smem);
  return bad != 0 || to != 0;
}
