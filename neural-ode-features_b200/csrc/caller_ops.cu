// SURVEY 8(f3), first step of widening into the callers of the hot path: the GroupNorm -> ReLU pairs of the
// downsamplers and of the classifier head (reference model.py:119-178, 231-250, 268-271) as ONE memory-bound pass.
// ATen runs them as three kernels (row moments, affine apply, in-place ReLU: 3 reads + 2 writes of the tensor);
// here one CTA owns one GroupNorm cell (channels-per-group x H*W contiguous floats of an NCHW tensor), keeps it in
// registers, computes the two-pass mean / biased variance of native_group_norm and writes relu(gamma*xhat + beta):
// 1 read + 1 write, the HBM floor.
#include "node_common.cuh"

namespace node {

constexpr int kGnThreads = 128;

__device__ __forceinline__ float block_sum_f(float v, float* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kGnThreads / 32; ++w) r += scratch[w];
  return r;
}

// VPT float4 vectors (VEC = 4) or scalars (VEC = 1) per thread; L = floats per cell <= kGnThreads * VPT * VEC.
template <int VPT, int VEC>
__global__ void __launch_bounds__(kGnThreads) k_groupnorm_relu(const float* __restrict__ x, float* __restrict__ y,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                int groups, int cpg, int HW, float eps, int relu) {
  __shared__ float scratch[kGnThreads / 32];
  const int L = cpg * HW;
  const size_t base = (size_t)blockIdx.x * L;
  const int g = blockIdx.x % groups;
  float v[VPT][VEC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int e = (i * kGnThreads + threadIdx.x) * VEC;
    if (e < L) {
      if (VEC == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(x + base + e));
        v[i][0] = q.x; v[i][1 % VEC] = q.y; v[i][2 % VEC] = q.z; v[i][3 % VEC] = q.w;
      } else {
        v[i][0] = __ldg(x + base + e);
      }
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) v[i][j] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) s += v[i][j];
  }
  const float inv_n = 1.0f / (float)L;
  const float mean = block_sum_f(s, scratch) * inv_n;
  float q2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int e = (i * kGnThreads + threadIdx.x) * VEC;
    if (e < L) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) { const float d = v[i][j] - mean; q2 = fmaf(d, d, q2); }
    }
  }
  const float rstd = 1.0f / sqrtf(block_sum_f(q2, scratch) * inv_n + eps);
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int e = (i * kGnThreads + threadIdx.x) * VEC;
    if (e < L) {
      float o[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int c = g * cpg + (e + j) / HW;
        const float a = rstd * __ldg(gamma + c);
        float r = fmaf(v[i][j] - mean, a, __ldg(beta + c));
        if (relu) r = fmaxf(r, 0.f);
        o[j] = r;
      }
      if (VEC == 4) *reinterpret_cast<float4*>(y + base + e) = make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
      else y[base + e] = o[0];
    }
  }
}

template <int VPT, int VEC>
static int launch_gn(const float* x, float* y, const float* gamma, const float* beta, int64_t cells, int groups, int cpg, int HW,
                     float eps, int relu, cudaStream_t st) {
  k_groupnorm_relu<VPT, VEC><<<(unsigned)cells, kGnThreads, 0, st>>>(x, y, gamma, beta, groups, cpg, HW, eps, relu);
  return (int)cudaGetLastError();
}

}  // namespace node

extern "C" int node_b200_groupnorm_relu(const float* x, float* y, const float* gamma, const float* beta, int64_t N, int C,
                                        int groups, int HW, float eps, int relu, void* stream) {
  using namespace node;
  if (N < 1 || C < 1 || groups < 1 || C % groups != 0 || HW < 1) return (int)cudaErrorInvalidValue;
  const int cpg = C / groups;
  const int64_t L = (int64_t)cpg * HW, cells = N * groups;
  if (cells > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = L % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)y % 16 == 0);
  if (vec) {
    const int64_t nv = L / 4;
    if (nv <= kGnThreads) return launch_gn<1, 4>(x, y, gamma, beta, cells, groups, cpg, HW, eps, relu, st);
    if (nv <= 2 * kGnThreads) return launch_gn<2, 4>(x, y, gamma, beta, cells, groups, cpg, HW, eps, relu, st);
    if (nv <= 4 * kGnThreads) return launch_gn<4, 4>(x, y, gamma, beta, cells, groups, cpg, HW, eps, relu, st);
    if (nv <= 8 * kGnThreads) return launch_gn<8, 4>(x, y, gamma, beta, cells, groups, cpg, HW, eps, relu, st);
    return (int)cudaErrorInvalidValue;      // cell larger than 4096 floats: the caller keeps its own GroupNorm
  }
  if (L <= kGnThreads) return launch_gn<1, 1>(x, y, gamma, beta, cells, groups, cpg, HW, eps, relu, st);
  if (L <= 4 * kGnThreads) return launch_gn<4, 1>(x, y, gamma, beta, cells, groups, cpg, HW, eps, relu, st);
  if (L <= 16 * kGnThreads) return launch_gn<16, 1>(x, y, gamma, beta, cells, groups, cpg, HW, eps, relu, st);
  if (L <= 32 * kGnThreads) return launch_gn<32, 1>(x, y, gamma, beta, cells, groups, cpg, HW, eps, relu, st);
  return (int)cudaErrorInvalidValue;
}
