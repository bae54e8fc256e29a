// Tuning aid (not part of the product): how far can one thread run ahead of the tensor core? Issues n back-to-back
// tcgen05.mma (M128 N128 K16, fp16, SS mode, operands resident in shared memory) and reports the clocks until the
// issuing thread is free again and until the MMAs have completed. A flat "issue" column means the instructions are queued.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-ode-features_b200/csrc -I include \
//        tools/queue_bench.cu -o tools/queue_bench
#include <cstdio>
#include <cuda_runtime.h>
#include "node_b200.h"
#include "step_engine.cuh"

using namespace node;

__global__ void __launch_bounds__(128, 1) k_queue(int n, long long* out) {
  extern __shared__ uint8_t raw[];
  const uint32_t s0 = ptx::smem_u32(raw);
  const uint32_t al = (s0 + 1023u) & ~1023u;
  uint8_t* base = raw + (al - s0);
  const uint32_t sA = al, sB = al + 64 * 1024, bar = al + 96 * 1024;
  uint32_t* slot = reinterpret_cast<uint32_t*>(base + 96 * 1024 + 64);
  for (int i = threadIdx.x; i < 96 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(base)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { ptx::mbar_init(bar, 1); ptx::fence_mbar_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc(ptx::smem_u32(slot), 512);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint64_t a = ptx::make_desc_nosw(sA, 4096, 128), b = ptx::make_desc_sw128(sB);
    for (int rep = 0; rep < 3; ++rep) {           // last repetition is reported
      const long long t0 = clock64();
      for (int i = 0; i < n; ++i) ptx::mma_f16_ss(tmem, a, b, kIdF16N128, i ? 1u : 0u);
      const long long t1 = clock64();
      ptx::tc_commit(bar);
      ptx::mbar_wait(bar, rep & 1);
      const long long t2 = clock64();
      out[0] = t1 - t0; out[1] = t2 - t0;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  long long* out; cudaMalloc(&out, 16);
  const int smem = 100 * 1024;
  cudaFuncSetAttribute(k_queue, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  printf("    n   issue clk   done clk   (64 clk of math per MMA)\n");
  for (int n : {1, 2, 4, 8, 12, 16, 24, 32, 48, 64, 96, 128, 192, 256}) {
    k_queue<<<1, 128, smem>>>(n, out);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("%5d %10lld %10lld\n", n, h[0], h[1]);
  }
  return 0;
}
