// Tuning aid (not part of the product): the step engine's own conv-job issue routine (issue_conv_job: TMA weight ring +
// tcgen05.mma stream + commits) run in isolation, one CTA per SM, to separate the cost of the issue path from the
// worker phases. Prints clocks per conv job.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-ode-features_b200/csrc -I include \
//        tools/issue_bench.cu -o tools/issue_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "node_b200.h"
#include "step_engine.cuh"

using namespace node;
using T = Tile<8, 8>;

__global__ void __launch_bounds__(512, 1) k_issue(const uint16_t* w16, int reps, int split, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  StepSmem sm;
  const uint32_t s0 = ptx::smem_u32(smem_raw);
  const uint32_t al = (s0 + 1023u) & ~1023u;
  uint8_t* base = smem_raw + (al - s0);
  size_t o = 0;
  sm.wring = al; o += (size_t)kNW * kW16TileBytes;
  sm.abase = al + (uint32_t)o; o += (size_t)2 * T::A_PART;
  sm.bar_wfull = al + (uint32_t)o; o += 8 * kNW;
  sm.bar_wfree = al + (uint32_t)o; o += 8 * kNW;
  sm.bar_turn = al + (uint32_t)o; o += 8 * 2;
  sm.bar_acc = al + (uint32_t)o; o += 8 * 2;
  sm.ring = reinterpret_cast<volatile uint32_t*>(base + o); o += 16;
  sm.tmem_slot = reinterpret_cast<uint32_t*>(base + o);
  const int tid = threadIdx.x;
  uint4* az = reinterpret_cast<uint4*>(base + (size_t)kNW * kW16TileBytes);
  for (int i = tid; i < 2 * T::A_PART / 16; i += blockDim.x) az[i] = make_uint4(0u, 0u, 0u, 0u);
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
  Jobs jb; jb.jobs = (uint32_t)reps; jb.jobs_full = (uint32_t)reps;
  bool timeout = false;
  if (tid == 0)
    for (uint32_t i = 0; i < kWAhead && i < jb.jobs * 9; ++i) request_tile<1>(sm, jb, w16, i);
  __syncthreads();
  if (tid >= 32 && tid < 64) {
    for (int j = 0; j < reps; ++j) {
      produce_conv_job<T, 1>(sm, jb, w16, 0, (uint32_t)j, (uint32_t)j, timeout);
      if (!ptx::mbar_wait_relaxed(sm.bar_acc, j & 1)) timeout = true;
    }
  }
  if (tid < 32) {
    const long long t0 = clock64();
    for (int j = 0; j < reps; ++j) {
      issue_conv_job<T, 1>(sm, jb, tmem, 0, (uint32_t)j, (uint32_t)j, split != 0, timeout);
      if (!ptx::mbar_wait(sm.bar_acc, j & 1)) timeout = true;
      ptx::tc_fence_after();
    }
    const long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = timeout ? -1 : (t1 - t0) / reps;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (tid < 32) ptx::tmem_dealloc(tmem, kTmemCols);
}

int main() {
  const size_t smem = 1024 + (size_t)kNW * kW16TileBytes + 2 * T::A_PART + 512;
  printf("kNW %d kWAhead %u\n", kNW, kWAhead);
  uint16_t* w16; cudaMalloc(&w16, (size_t)kW16Sets * 9 * kW16TileBytes); cudaMemset(w16, 0, (size_t)kW16Sets * 9 * kW16TileBytes);
  long long* out; cudaMalloc(&out, 148 * sizeof(long long));
  cudaFuncSetAttribute(k_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int split = 0; split < 2; ++split)
    for (int grid : {1, 148}) {
      k_issue<<<grid, 512, smem>>>(w16, 40, split, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long h[148]; cudaMemcpy(h, out, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
      long long mn = h[0], mx = h[0];
      for (int i = 0; i < grid; ++i) { mn = h[i] < mn ? h[i] : mn; mx = h[i] > mx ? h[i] : mx; }
      printf("issue_conv_job split %d grid %3d : clk/job min %6lld max %6lld (floor %d)\n", split, grid, mn, mx, split ? 6912 : 4608);
    }
  return 0;
}
