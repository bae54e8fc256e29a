// Pieces of the adjoint's augmented dynamics (adjoint.py:32-55) for WIDE ODE-Net dynamics (n_filters = 128, 192, 256: the
// paper's CIFAR training setting, reproduce.sh:21-25 with --adjoint) that the 64-filter kernels do not cover:
//   * node_b200_groupnorm_backward_ex - backward of y = relu?(GroupNorm_32(x + bias[c] + t * Tmap[c][pix])) for any number of
//     channels per group (the caller kernels' backward is specialised to 2 channels per group), with the folded time channel
//     of ConcatConv2d (model.py:320-323) added on the fly;
//   * node_b200_batch_colsum          - S[j] = sum_n g[n][j] in a fixed order (float64), for the bias / time-channel gradients;
//   * node_b200_pow2_scale            - the power-of-two operand scale of a non-negative activation from its |max| bit pattern.
// The convolutions of the backward pass (data gradients = conv3x3 with transposed / flipped weights, weight gradients) run on
// the existing tcgen05 engines, block by block (node_b200/wide.py).
#include "node_common.cuh"

namespace node {

constexpr int kWvThreads = 128;

__device__ __forceinline__ float wv_block_sum(float v, float* scratch) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < kWvThreads / 32; ++i) r += scratch[i];
  return r;
}

// One CTA per (image, group) cell of L = cpg * HW floats (L <= 4096, kept in shared memory).
// in:  x (+ bias + t * tmap), g = dL/d(output); out: gx = dL/dx, pgamma / pbeta [N][C] per-image partials of dL/dgamma, dL/dbeta.
__global__ void __launch_bounds__(kWvThreads) k_gn_backward_ex(const float* __restrict__ x, const float* __restrict__ g,
                                                                float* __restrict__ gx, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, const float* __restrict__ add_bias,
                                                                const float* __restrict__ add_tmap, const float* __restrict__ t_dev,
                                                                float tsign, float* __restrict__ pgamma, float* __restrict__ pbeta,
                                                                int groups, int cpg, int HW, float eps, int relu) {
  __shared__ float xs[4096];
  __shared__ float ds[4096];
  __shared__ float scratch[kWvThreads / 32];
  const int L = cpg * HW;
  const size_t base = (size_t)blockIdx.x * L;
  const int grp = blockIdx.x % groups;
  const size_t n = blockIdx.x / groups;
  const float t = add_tmap != nullptr ? tsign * __ldg(t_dev) : 0.f;
  float s = 0.f;
  for (int e = threadIdx.x; e < L; e += kWvThreads) {
    const int c = grp * cpg + e / HW, pix = e % HW;
    float v = __ldg(x + base + e);
    if (add_bias != nullptr) v += __ldg(add_bias + c);
    if (add_tmap != nullptr) v = fmaf(t, __ldg(add_tmap + (size_t)c * HW + pix), v);
    xs[e] = v;
    s += v;
  }
  const float inv_n = 1.0f / (float)L;
  const float mean = wv_block_sum(s, scratch) * inv_n;
  float q2 = 0.f;
  for (int e = threadIdx.x; e < L; e += kWvThreads) { const float d = xs[e] - mean; q2 = fmaf(d, d, q2); }
  const float rstd = 1.0f / sqrtf(wv_block_sum(q2, scratch) * inv_n + eps);
  // xs <- xhat, ds <- g masked by the ReLU; S1 = sum g*gamma, S2 = sum g*gamma*xhat
  float s1 = 0.f, s2 = 0.f;
  for (int e = threadIdx.x; e < L; e += kWvThreads) {
    const int c = grp * cpg + e / HW;
    const float ga = __ldg(gamma + c);
    const float xh = (xs[e] - mean) * rstd;
    float d = __ldg(g + base + e);
    if (relu && !(fmaf(xh, ga, __ldg(beta + c)) > 0.f)) d = 0.f;
    xs[e] = xh;
    ds[e] = d;
    s1 = fmaf(d, ga, s1);
    s2 = fmaf(d * ga, xh, s2);
  }
  const float m1 = wv_block_sum(s1, scratch) * inv_n;
  const float m2 = wv_block_sum(s2, scratch) * inv_n;
  for (int e = threadIdx.x; e < L; e += kWvThreads) {
    const int c = grp * cpg + e / HW;
    gx[base + e] = rstd * (ds[e] * __ldg(gamma + c) - m1 - xs[e] * m2);
  }
  // per-channel partials: warp w takes channels w, w + 4, ...
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int cl = warp; cl < cpg; cl += kWvThreads / 32) {
    float a = 0.f, b = 0.f;
    for (int p = lane; p < HW; p += 32) { const float d = ds[cl * HW + p]; a = fmaf(d, xs[cl * HW + p], a); b += d; }
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { pgamma[n * (size_t)groups * cpg + grp * cpg + cl] = a; pbeta[n * (size_t)groups * cpg + grp * cpg + cl] = b; }
  }
}

// out[j] = sum_n v[n][j], j < cols: a CTA = 32 columns x 8 row lanes (a warp reads 32 consecutive columns of one image), float64
// accumulation, the 8 lane sums folded in a fixed order (deterministic)
__global__ void __launch_bounds__(256) k_batch_colsum(const float* __restrict__ v, float* __restrict__ out, int64_t N, int64_t cols) {
  __shared__ double red[8][32];
  const int lane = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int64_t j = (int64_t)blockIdx.x * 32 + lane;
  double s = 0.0;
  if (j < cols) {
    int64_t n = rl;
    for (; n + 24 < N; n += 32) {
      const float a = __ldg(v + n * cols + j), b = __ldg(v + (n + 8) * cols + j), c = __ldg(v + (n + 16) * cols + j),
                  d = __ldg(v + (n + 24) * cols + j);
      s += (double)a; s += (double)b; s += (double)c; s += (double)d;
    }
    for (; n < N; n += 8) s += (double)__ldg(v + n * cols + j);
  }
  red[rl][lane] = s;
  __syncthreads();
  if (rl == 0 && j < cols) {
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot += red[i][lane];
    out[j] = (float)tot;
  }
}

// scale = 2^floor(log2(16384 / max)) clamped to [2^-24, 2^24]; 1 when max == 0
__global__ void k_pow2_scale(const unsigned* __restrict__ max_bits, float* __restrict__ scale) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float m = __uint_as_float(*max_bits);
  int e = m > 0.f ? (int)floorf(log2f(16384.0f / m)) : 0;
  e = max(-24, min(24, e));
  *scale = exp2f((float)e);
}

}  // namespace node

extern "C" int node_b200_groupnorm_backward_ex(const float* x, const float* grad_out, float* grad_in, const float* gamma,
                                               const float* beta, const float* add_bias, const float* add_tmap, const float* t_dev,
                                               float tsign, float* partials, float* grad_gamma, float* grad_beta, int64_t N, int C,
                                               int groups, int HW, float eps, int relu, void* stream) {
  using namespace node;
  if (N < 1 || C < 1 || groups < 1 || C % groups != 0 || HW < 1) return (int)cudaErrorInvalidValue;
  const int cpg = C / groups;
  if ((int64_t)cpg * HW > 4096 || N * groups > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  float* pg = partials;
  float* pb = partials + N * C;
  k_gn_backward_ex<<<(unsigned)(N * groups), kWvThreads, 0, st>>>(x, grad_out, grad_in, gamma, beta, add_bias, add_tmap, t_dev, tsign, pg, pb,
                                                                  groups, cpg, HW, eps, relu);
  NODE_CUDA_OK(cudaGetLastError());
  k_batch_colsum<<<(C + 31) / 32, 256, 0, st>>>(pg, grad_gamma, N, C);
  k_batch_colsum<<<(C + 31) / 32, 256, 0, st>>>(pb, grad_beta, N, C);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_batch_colsum(const float* v, float* out, int64_t N, int64_t cols, void* stream) {
  if (N < 1 || cols < 1) return (int)cudaErrorInvalidValue;
  node::k_batch_colsum<<<(unsigned)((cols + 31) / 32), 256, 0, (cudaStream_t)stream>>>(v, out, N, cols);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_pow2_scale(const unsigned* max_bits, float* scale, void* stream) {
  node::k_pow2_scale<<<1, 32, 0, (cudaStream_t)stream>>>(max_bits, scale);
  return (int)cudaGetLastError();
}
