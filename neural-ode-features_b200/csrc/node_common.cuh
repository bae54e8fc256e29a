// Shared device helpers for the dopri5 hot path (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/node_b200.h"

#define NODE_CUDA_OK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) return (int)e__; } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: remember per device whether this kernel has it
// (a process that moves a model to cuda:1 would otherwise launch with the 48 KB default there).
#define NODE_SET_SMEM_ONCE(kernel, bytes)                                                                         \
  do {                                                                                                            \
    static bool done__[64] = {};                                                                                  \
    int dev__ = 0;                                                                                                \
    NODE_CUDA_OK(cudaGetDevice(&dev__));                                                                          \
    if (dev__ < 0 || dev__ >= 64 || !done__[dev__]) {                                                             \
      NODE_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));      \
      if (dev__ >= 0 && dev__ < 64) done__[dev__] = true;                                                         \
    }                                                                                                             \
  } while (0)

namespace node {

// Dormand-Prince / Shampine tableau (reference dopri5.py:11-36), evaluated in double exactly
// as the Python literals are, then rounded to the state dtype at the point of use.
__host__ __device__ constexpr double kAlpha(int i) {
  return i == 0 ? 1.0 / 5 : i == 1 ? 3.0 / 10 : i == 2 ? 4.0 / 5 : i == 3 ? 8.0 / 9 : 1.0;
}
// rows 0..5 = beta of stages 2..7; row 6 = C_MID; row 7 = probe (single coefficient 1)
__host__ __device__ constexpr double kCoef(int row, int j) {
  switch (row) {
    case 0: return j == 0 ? 1.0 / 5 : 0.0;
    case 1: return j == 0 ? 3.0 / 40 : j == 1 ? 9.0 / 40 : 0.0;
    case 2: return j == 0 ? 44.0 / 45 : j == 1 ? -56.0 / 15 : j == 2 ? 32.0 / 9 : 0.0;
    case 3: return j == 0 ? 19372.0 / 6561 : j == 1 ? -25360.0 / 2187 : j == 2 ? 64448.0 / 6561
                 : j == 3 ? -212.0 / 729 : 0.0;
    case 4: return j == 0 ? 9017.0 / 3168 : j == 1 ? -355.0 / 33 : j == 2 ? 46732.0 / 5247
                 : j == 3 ? 49.0 / 176 : j == 4 ? -5103.0 / 18656 : 0.0;
    case 5: return j == 0 ? 35.0 / 384 : j == 1 ? 0.0 : j == 2 ? 500.0 / 1113 : j == 3 ? 125.0 / 192
                 : j == 4 ? -2187.0 / 6784 : j == 5 ? 11.0 / 84 : 0.0;
    case 6: return j == 0 ? 6025192743.0 / 30085553152.0 / 2 : j == 1 ? 0.0
                 : j == 2 ? 51252292925.0 / 65400821598.0 / 2 : j == 3 ? -2691868925.0 / 45128329728.0 / 2
                 : j == 4 ? 187940372067.0 / 1594534317056.0 / 2 : j == 5 ? -1776094331.0 / 19743644256.0 / 2
                 : 11237099.0 / 235043384.0 / 2;
    default: return j == 0 ? 1.0 : 0.0;
  }
}
__host__ __device__ constexpr double kCErr(int j) {
  return j == 0 ? 35.0 / 384 - 1951.0 / 21600 : j == 1 ? 0.0 : j == 2 ? 500.0 / 1113 - 22642.0 / 50085
       : j == 3 ? 125.0 / 192 - 451.0 / 720 : j == 4 ? -2187.0 / 6784 - -12231.0 / 42400
       : j == 5 ? 11.0 / 84 - 649.0 / 6300 : -1.0 / 60.0;
}
__host__ __device__ constexpr int kRowLen(int row) { return row < 6 ? row + 1 : row == 6 ? 7 : 1; }

// Round-to-nearest arithmetic that the compiler may not contract into FMAs: the reference
// evaluates every elementwise op as a separate ATen kernel, so mul and add round separately.
template <typename T> struct Arith;
template <> struct Arith<float> {
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float div(float a, float b) { return __fdiv_rn(a, b); }
  static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
  static __device__ __forceinline__ float max(float a, float b) { return (a != a || b != b) ? a + b : fmaxf(a, b); }
};
template <> struct Arith<double> {
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double div(double a, double b) { return __ddiv_rn(a, b); }
  static __device__ __forceinline__ double abs(double a) { return fabs(a); }
  static __device__ __forceinline__ double max(double a, double b) { return (a != a || b != b) ? a + b : fmax(a, b); }
};

template <typename T> __device__ __forceinline__ T ctl_h(const node_ctl_t* c);
template <> __device__ __forceinline__ float ctl_h<float>(const node_ctl_t* c) { return c->h32; }
template <> __device__ __forceinline__ double ctl_h<double>(const node_ctl_t* c) { return c->h64; }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of a double; result valid in thread 0. `scratch` holds >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double r = 0.0;
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    r = lane < nw ? scratch[lane] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

}  // namespace node
