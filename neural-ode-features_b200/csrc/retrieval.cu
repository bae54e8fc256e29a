// SURVEY 8f-4 - the edges after the feature extractor (reference evaluate.py:88-94 feature layout, :326-339 retrieval):
//     features[tol, T, N, D]  ->  features / (||features||_{axis=-2} + 1e-7)  ->  scores = queries . db^T
// The reference normalises along axis -2 - over the SAMPLES, one norm per (tol, T, feature dimension) (evaluate.py:326,
// `np.linalg.norm(features, axis=-2, keepdims=True)`) - and scores every query against every sample with one fp32
// matrix product (evaluate.py:339). Both are restated here for device-resident feature buffers: the scores matrix of
// the 10,000-image test set is 400 MB and is produced at the rate the HBM takes it (fp32 FFMA, K = 64; bound by the
// N x N write, not by the 12.8 GFLOP).
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/node_b200.h"

namespace node {

// norms[p][d] = sqrt(sum_n f[p][n][d]^2): one CTA per plane p = (tol, T); float64 accumulation, fixed order.
__global__ void __launch_bounds__(256) k_feature_colnorm(const float* __restrict__ f, float* __restrict__ norms, int64_t N, int D) {
  __shared__ double red[256];
  const float* fp = f + (size_t)blockIdx.x * (size_t)N * D;
  for (int d0 = 0; d0 < D; d0 += 64) {
    const int d = d0 + (threadIdx.x & 63), part = threadIdx.x >> 6;
    double s = 0.0;
    if (d < D)
      for (int64_t n = part; n < N; n += 4) { const float v = fp[(size_t)n * D + d]; s += (double)v * (double)v; }
    red[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x < 64 && d < D)
      norms[(size_t)blockIdx.x * D + d] = sqrtf((float)(red[threadIdx.x] + red[threadIdx.x + 64] + red[threadIdx.x + 128] + red[threadIdx.x + 192]));
    __syncthreads();
  }
}

// out = f / (norm + 1e-7f), fp32 like numpy (the python scalar takes the array's dtype)
__global__ void __launch_bounds__(256) k_feature_scale(const float* __restrict__ f, const float* __restrict__ norms, float* __restrict__ out,
                                                       int64_t per_plane, int D, int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / per_plane;
    const int d = (int)(i % D);
    out[i] = __fdiv_rn(f[i], __fadd_rn(norms[p * D + d], 1e-7f));
  }
}

// scores[q][s] = sum_d Q[q][d] * DB[s][d]: 128 x 128 tile per CTA, 8 x 8 outputs per thread, K in chunks of 32.
constexpr int kTile = 128, kKC = 32, kPad = 4;
__global__ void __launch_bounds__(256) k_scores(const float* __restrict__ Q, const float* __restrict__ DB, float* __restrict__ S,
                                                int64_t nq, int64_t ns, int D) {
  __shared__ __align__(16) float qs[kKC][kTile + kPad];
  __shared__ __align__(16) float ds[kKC][kTile + kPad];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t q0 = (int64_t)blockIdx.y * kTile, s0 = (int64_t)blockIdx.x * kTile;
  const bool dvec = (D & 3) == 0;      // rows 16-byte aligned
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < D; k0 += kKC) {
    // stage both tiles transposed: [k][row]; every thread moves 4 + 4 chunks of 4 k-values
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int idx = tid + 256 * r, row = idx >> 3, kq = (idx & 7) * 4;
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
      if (dvec && k0 + kq + 3 < D) {
        if (q0 + row < nq) a = *reinterpret_cast<const float4*>(Q + (size_t)(q0 + row) * D + k0 + kq);
        if (s0 + row < ns) b = *reinterpret_cast<const float4*>(DB + (size_t)(s0 + row) * D + k0 + kq);
      } else {
        float av[4] = {0.f, 0.f, 0.f, 0.f}, bv[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < 4; ++j)
          if (k0 + kq + j < D) {
            if (q0 + row < nq) av[j] = Q[(size_t)(q0 + row) * D + k0 + kq + j];
            if (s0 + row < ns) bv[j] = DB[(size_t)(s0 + row) * D + k0 + kq + j];
          }
        a = make_float4(av[0], av[1], av[2], av[3]);
        b = make_float4(bv[0], bv[1], bv[2], bv[3]);
      }
      qs[kq][row] = a.x; qs[kq + 1][row] = a.y; qs[kq + 2][row] = a.z; qs[kq + 3][row] = a.w;
      ds[kq][row] = b.x; ds[kq + 1][row] = b.y; ds[kq + 2][row] = b.z; ds[kq + 3][row] = b.w;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kKC; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&qs[k][ty * 8]), a1 = *reinterpret_cast<const float4*>(&qs[k][ty * 8 + 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&ds[k][tx * 8]), b1 = *reinterpret_cast<const float4*>(&ds[k][tx * 8 + 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool vec = (ns & 3) == 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t q = q0 + ty * 8 + i;
    if (q >= nq) continue;
    float* dst = S + (size_t)q * ns + s0 + tx * 8;
    if (vec && s0 + tx * 8 + 7 < ns) {
      __stcs(reinterpret_cast<float4*>(dst), make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));        // streaming: written once, read by the host
      __stcs(reinterpret_cast<float4*>(dst) + 1, make_float4(acc[i][4], acc[i][5], acc[i][6], acc[i][7]));
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (s0 + tx * 8 + j < ns) dst[j] = acc[i][j];
    }
  }
}

}  // namespace node

extern "C" int node_b200_feature_normalize(const float* features, float* out, float* norms, int64_t planes, int64_t N, int D, void* stream) {
  if (planes <= 0 || N <= 0 || D <= 0) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  node::k_feature_colnorm<<<(unsigned)planes, 256, 0, st>>>(features, norms, N, D);
  const int64_t total = planes * N * D;
  int64_t blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  node::k_feature_scale<<<(unsigned)blocks, 256, 0, st>>>(features, norms, out, N * D, D, total);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_retrieval_scores(const float* queries, const float* db, float* scores, int64_t nq, int64_t ns, int D, void* stream) {
  if (nq <= 0 || ns <= 0 || D <= 0) return (int)cudaErrorInvalidValue;
  dim3 grid((unsigned)((ns + node::kTile - 1) / node::kTile), (unsigned)((nq + node::kTile - 1) / node::kTile));
  node::k_scores<<<grid, 256, 0, (cudaStream_t)stream>>>(queries, db, scores, nq, ns, D);
  return (int)cudaGetLastError();
}
