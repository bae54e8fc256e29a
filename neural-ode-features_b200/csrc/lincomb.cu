// Linear combinations with DEVICE-resident coefficients, as autograd nodes of the unrolled route (node_b200/unrolled.py):
// the reference records `y + sum([(h*c_j) * k_j])` (misc.py:22-25), the Hermite fit (interp.py:5-35) and the interpolant
// (interp.py:54-65) as one ATen multiply and one ATen add per term; here a combination is ONE pass with the same rounding (every
// product and every addition rounds separately, left to right from 0, then base + sum), and its backward is two passes:
//   node_b200_lincomb        out = base + sum_j coef[j] * src_j          (base may be null)
//   node_b200_lincomb_scale  dst_j = coef[j] * g                          (gradients of the sources)
//   node_b200_lincomb_dots   dot[j] = sum_e src_j[e] * g[e]               (gradients of the coefficients; float64, fixed order)
// fp32 or fp64, up to 7 sources, numel elements each (scalar tail handled), coefficients in the tensors' dtype.
#include "node_common.cuh"

namespace node {

constexpr int kLcThreads = 256;
constexpr int kLcBlocks = 296;
struct LcPtrs { const void* p[7]; };
struct LcOut { void* p[7]; };

template <typename T>
__global__ void __launch_bounds__(kLcThreads) k_lincomb(T* __restrict__ out, const T* __restrict__ base, LcPtrs srcs, const T* __restrict__ coef,
                                                        int n, int64_t numel) {
  using A = Arith<T>;
  T c[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) c[j] = j < n ? coef[j] : (T)0;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += nthr) {
    T acc = (T)0;
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (j < n) acc = A::add(acc, A::mul(c[j], reinterpret_cast<const T*>(srcs.p[j])[i]));
    out[i] = base != nullptr ? A::add(base[i], acc) : acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(kLcThreads) k_lincomb_scale(LcOut dst, const T* __restrict__ g, const T* __restrict__ coef, int n, int64_t numel) {
  using A = Arith<T>;
  T c[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) c[j] = j < n ? coef[j] : (T)0;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += nthr) {
    const T gv = g[i];
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (j < n && dst.p[j] != nullptr) reinterpret_cast<T*>(dst.p[j])[i] = A::mul(c[j], gv);
  }
}

// partial[j][block] = sum over the block's elements of src_j * g (float64)
template <typename T>
__global__ void __launch_bounds__(kLcThreads) k_lincomb_dots(LcPtrs srcs, const T* __restrict__ g, int n, int64_t numel, double* __restrict__ partial) {
  __shared__ double scratch[32];
  double acc[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) acc[j] = 0.0;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < numel; i += nthr) {
    const double gv = (double)g[i];
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (j < n) acc[j] += (double)reinterpret_cast<const T*>(srcs.p[j])[i] * gv;
  }
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    if (j < n) {
      const double r = block_sum(acc[j], scratch);
      if (threadIdx.x == 0) partial[(size_t)j * kLcBlocks + blockIdx.x] = r;
    }
  }
}

template <typename T>
__global__ void k_lincomb_fold(const double* __restrict__ partial, int nblocks, T* __restrict__ dots) {
  const int j = blockIdx.x;
  double v = 0.0;
  for (int b = threadIdx.x; b < nblocks; b += 32) v += partial[(size_t)j * kLcBlocks + b];
  v = warp_sum(v);
  if (threadIdx.x == 0) dots[j] = (T)v;
}

static int lc_grid(int64_t numel) {
  int64_t b = (numel + kLcThreads - 1) / kLcThreads;
  return (int)(b < 1 ? 1 : (b > kLcBlocks ? kLcBlocks : b));
}

}  // namespace node

using namespace node;

extern "C" int node_b200_lincomb(int dtype, void* out, const void* base, const void* const* srcs, const void* coef, int n, int64_t numel,
                                 void* stream) {
  if (n < 1 || n > 7 || numel < 1) return (int)cudaErrorInvalidValue;
  LcPtrs p{};
  for (int j = 0; j < n; ++j) p.p[j] = srcs[j];
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == NODE_F32) k_lincomb<float><<<lc_grid(numel) * 4 > 148 * 8 ? 148 * 8 : lc_grid(numel) * 4, kLcThreads, 0, st>>>((float*)out, (const float*)base, p, (const float*)coef, n, numel);
  else k_lincomb<double><<<lc_grid(numel) * 4 > 148 * 8 ? 148 * 8 : lc_grid(numel) * 4, kLcThreads, 0, st>>>((double*)out, (const double*)base, p, (const double*)coef, n, numel);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_lincomb_scale(int dtype, void* const* dsts, const void* g, const void* coef, int n, int64_t numel, void* stream) {
  if (n < 1 || n > 7 || numel < 1) return (int)cudaErrorInvalidValue;
  LcOut p{};
  for (int j = 0; j < n; ++j) p.p[j] = dsts[j];
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = lc_grid(numel) * 4 > 148 * 8 ? 148 * 8 : lc_grid(numel) * 4;
  if (dtype == NODE_F32) k_lincomb_scale<float><<<grid, kLcThreads, 0, st>>>(p, (const float*)g, (const float*)coef, n, numel);
  else k_lincomb_scale<double><<<grid, kLcThreads, 0, st>>>(p, (const double*)g, (const double*)coef, n, numel);
  return (int)cudaGetLastError();
}

// partial: kLcBlocks * 7 doubles of scratch; dots: n values of the tensors' dtype
extern "C" int node_b200_lincomb_dots(int dtype, const void* const* srcs, const void* g, int n, int64_t numel, double* partial, void* dots,
                                      void* stream) {
  if (n < 1 || n > 7 || numel < 1) return (int)cudaErrorInvalidValue;
  LcPtrs p{};
  for (int j = 0; j < n; ++j) p.p[j] = srcs[j];
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = lc_grid(numel);
  if (dtype == NODE_F32) {
    k_lincomb_dots<float><<<grid, kLcThreads, 0, st>>>(p, (const float*)g, n, numel, partial);
    k_lincomb_fold<float><<<n, 32, 0, st>>>(partial, grid, (float*)dots);
  } else {
    k_lincomb_dots<double><<<grid, kLcThreads, 0, st>>>(p, (const double*)g, n, numel, partial);
    k_lincomb_fold<double><<<n, 32, 0, st>>>(partial, grid, (double*)dots);
  }
  return (int)cudaGetLastError();
}

extern "C" int64_t node_b200_lincomb_scratch_doubles(void) { return (int64_t)kLcBlocks * 7; }
