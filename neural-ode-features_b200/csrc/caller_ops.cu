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

// y = post * relu?(GN(x + bias[c] + t * tmap[c][pix])): the GroupNorm that follows a ConcatConv2d whose time channel has been
// folded into a position-dependent bias (model.py:320-323; conv(cat([t*1, x])) = conv(x, W[:,1:]) + b + t*Tmap) - used by the
// wide (C = 128, 256, ...) dynamics, whose convolutions run as 64-channel blocks and carry no bias of their own.
template <int VPT, int VEC>
__global__ void __launch_bounds__(kGnThreads) k_groupnorm_relu_ex(const float* __restrict__ x, float* __restrict__ y,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   const float* __restrict__ add_bias, const float* __restrict__ add_tmap,
                                                                   const float* __restrict__ t_dev, float tsign, float post,
                                                                   int groups, int cpg, int HW, float eps, int relu) {
  __shared__ float scratch[kGnThreads / 32];
  const int L = cpg * HW;
  const size_t base = (size_t)blockIdx.x * L;
  const int g = blockIdx.x % groups;
  const float t = add_tmap != nullptr ? tsign * __ldg(t_dev) : 0.f;
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
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int c = g * cpg + (e + j) / HW, pix = (e + j) % HW;
        if (add_bias != nullptr) v[i][j] += __ldg(add_bias + c);
        if (add_tmap != nullptr) v[i][j] = fmaf(t, __ldg(add_tmap + (size_t)c * HW + pix), v[i][j]);
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
        o[j] = r * post;
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

extern "C" int node_b200_groupnorm_relu_ex(const float* x, float* y, const float* gamma, const float* beta, const float* add_bias,
                                           const float* add_tmap, const float* t_dev, float tsign, float post, int64_t N, int C,
                                           int groups, int HW, float eps, int relu, void* stream) {
  using namespace node;
  if (N < 1 || C < 1 || groups < 1 || C % groups != 0 || HW < 1 || (add_tmap != nullptr && t_dev == nullptr)) return (int)cudaErrorInvalidValue;
  const int cpg = C / groups;
  const int64_t L = (int64_t)cpg * HW, cells = N * groups;
  if (cells > 0x7fffffffLL || L % 4 != 0 || L > 8 * 4 * kGnThreads || ((uintptr_t)x % 16) || ((uintptr_t)y % 16)) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nv = L / 4;
#define NODE_GN_EX(VPT) k_groupnorm_relu_ex<VPT, 4><<<(unsigned)cells, kGnThreads, 0, st>>>(x, y, gamma, beta, add_bias, add_tmap, t_dev, tsign, post, groups, cpg, HW, eps, relu)
  if (nv <= kGnThreads) NODE_GN_EX(1);
  else if (nv <= 2 * kGnThreads) NODE_GN_EX(2);
  else if (nv <= 4 * kGnThreads) NODE_GN_EX(4);
  else NODE_GN_EX(8);
#undef NODE_GN_EX
  return (int)cudaGetLastError();
}

// ---- stem: conv0 (Conv2d(CIN, 64, 3, 1), bias) -> GroupNorm(32, 64) -> ReLU in one pass ------------------------------------
// The first layer of the reference's ResDownsample / ConvDownsample (model.py:119-178) followed by the first GroupNorm:
// cuDNN + ATen run it as convolution, bias add, row moments, affine apply, ReLU = 5 passes over the [N,64,HO,WO] tensor;
// here one CTA of 256 threads owns one image (4 images resident per SM), a thread 4 output pixels: the 64 output channels
// are produced 8 at a time (4 GroupNorm cells) from the shared-memory copy of the image, their statistics are block
// reductions, and only relu(GN(conv(x))) is written - 1 pass. Plain fp32 FFMA (K = 27: no tensor-core shape).
namespace node {

// Sum NV per-thread values over the CTA: warp shuffles, one partial per warp, warp 0 folds the partials, everyone reads
// the totals (3 barriers; the fold costs NV loads per thread instead of NV x #warps).
template <int NV>
__device__ __forceinline__ void block_sum_n(float (&v)[NV], float* scratch, float* total, int nwarp) {
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();                                   // previous totals / partials have been consumed
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) scratch[warp * NV + i] = v[i];
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float r = lane < nwarp ? scratch[lane * NV + i] : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
      if (lane == 0) total[i] = r;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = total[i];
}

constexpr int kStemThreads = 256;

template <int CIN, int HIN, int WIN>
__global__ void __launch_bounds__(kStemThreads, 4)
k_stem_gn_relu(const float* __restrict__ x, const float* __restrict__ cw, const float* __restrict__ cb, const float* __restrict__ gamma,
               const float* __restrict__ beta, float* __restrict__ out, float eps) {
  // 256 threads per image, 4 images resident per SM (their barrier phases interleave); a thread owns PPT output pixels
  constexpr int HO = HIN - 2, WO = WIN - 2, NPIX = HO * WO, K = CIN * 9, NT = kStemThreads, PPT = (NPIX + NT - 1) / NT;
  __shared__ float s_in[CIN * HIN * WIN];
  __shared__ __align__(16) float s_w[64 * K];        // [pass of 8 channels][k][8 channels]
  __shared__ float s_b[64], s_g[64], s_be[64];
  __shared__ float scratch[(NT / 32) * 8], total[8];
  const int tid = threadIdx.x, n = blockIdx.x;
  for (int i = tid; i < CIN * HIN * WIN; i += NT) s_in[i] = x[(size_t)n * CIN * HIN * WIN + i];
  for (int i = tid; i < 64 * K; i += NT) {
    const int c = i / K, k = i % K;
    s_w[((c >> 3) * K + k) * 8 + (c & 7)] = cw[i];
  }
  if (tid < 64) { s_b[tid] = cb[tid]; s_g[tid] = gamma[tid]; s_be[tid] = beta[tid]; }
  __syncthreads();
  int base[PPT];
  bool valid[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int p = tid + q * NT;
    valid[q] = p < NPIX;
    const int pp = valid[q] ? p : 0;
    base[q] = (pp / WO) * WIN + pp % WO;
  }
  constexpr float inv_n = 1.0f / (float)(2 * NPIX);
#pragma unroll 1
  for (int ps = 0; ps < 8; ++ps) {
    float acc[PPT][8];
#pragma unroll
    for (int q = 0; q < PPT; ++q)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[q][j] = s_b[8 * ps + j];
    const float4* w4 = reinterpret_cast<const float4*>(s_w + (size_t)ps * K * 8);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float4 a = w4[2 * k], b = w4[2 * k + 1];
      const int off = (k / 9) * HIN * WIN + ((k % 9) / 3) * WIN + k % 3;
#pragma unroll
      for (int q = 0; q < PPT; ++q) {
        const float v = s_in[base[q] + off];
        acc[q][0] = fmaf(v, a.x, acc[q][0]); acc[q][1] = fmaf(v, a.y, acc[q][1]); acc[q][2] = fmaf(v, a.z, acc[q][2]); acc[q][3] = fmaf(v, a.w, acc[q][3]);
        acc[q][4] = fmaf(v, b.x, acc[q][4]); acc[q][5] = fmaf(v, b.y, acc[q][5]); acc[q][6] = fmaf(v, b.z, acc[q][6]); acc[q][7] = fmaf(v, b.w, acc[q][7]);
      }
    }
    // one round: sums and sums of squares of the 4 GroupNorm cells; E[x^2] - mean^2 is redone in two passes when a cell
    // is ill-conditioned (CTA-uniform decision: every thread sees the same totals)
    float s[8];
#pragma unroll
    for (int g = 0; g < 8; ++g) s[g] = 0.f;
#pragma unroll
    for (int q = 0; q < PPT; ++q)
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (valid[q]) {
          s[g] += acc[q][2 * g] + acc[q][2 * g + 1];
          s[4 + g] += fmaf(acc[q][2 * g], acc[q][2 * g], acc[q][2 * g + 1] * acc[q][2 * g + 1]);
        }
      }
    block_sum_n<8>(s, scratch, total, NT / 32);
    bool ill = false;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      s[g] *= inv_n;                                                  // mean
      s[4 + g] = fmaxf(fmaf(-s[g], s[g], s[4 + g] * inv_n), 0.f);     // variance
      ill |= s[g] * s[g] > 16.0f * s[4 + g];
    }
    if (ill) {
      float qq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < PPT; ++q)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float d0 = acc[q][2 * g] - s[g], d1 = acc[q][2 * g + 1] - s[g];
          if (valid[q]) qq[g] += fmaf(d0, d0, d1 * d1);
        }
      block_sum_n<4>(qq, scratch, total, NT / 32);
#pragma unroll
      for (int g = 0; g < 4; ++g) s[4 + g] = qq[g] * inv_n;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float rstd = 1.0f / sqrtf(s[4 + g] + eps);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int c = 8 * ps + 2 * g + h;
        const float ga = rstd * s_g[c], be = s_be[c];
#pragma unroll
        for (int q = 0; q < PPT; ++q)
          if (valid[q]) out[((size_t)n * 64 + c) * NPIX + tid + q * NT] = fmaxf(fmaf(acc[q][2 * g + h] - s[g], ga, be), 0.f);
      }
    }
  }
}

}  // namespace node

extern "C" int node_b200_stem_gn_relu(const float* x, const float* conv_w, const float* conv_b, const float* gn_w, const float* gn_b,
                                      float* out, int N, int CIN, int HIN, int WIN, float eps, void* stream) {
  using namespace node;
  if (N < 1) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  if (CIN == 3 && HIN == 32 && WIN == 32) {
    k_stem_gn_relu<3, 32, 32><<<N, kStemThreads, 0, st>>>(x, conv_w, conv_b, gn_w, gn_b, out, eps);
  } else if (CIN == 1 && HIN == 28 && WIN == 28) {
    k_stem_gn_relu<1, 28, 28><<<N, kStemThreads, 0, st>>>(x, conv_w, conv_b, gn_w, gn_b, out, eps);
  } else {
    return (int)cudaErrorInvalidValue;      // the caller keeps its own ops
  }
  return (int)cudaGetLastError();
}

// ---- classifier head: GroupNorm(32, 64) -> ReLU -> global average pool (-> Linear(64, n_out)) in one pass ----------------
// The reference's FCClassifier (model.py:231-250), also applied per output time in feature mode (model.py:39-40). One
// 256-thread CTA per image: thread t owns channel t / 4 and a quarter of its pixels, the 8 threads of a GroupNorm cell
// and the 4 threads of a channel are neighbours in a warp (shuffles only), two-pass statistics, the 64 pooled values
// meet in shared memory for the optional linear layer. Reads the [N,64,HW] tensor once, writes [N,64] or [N,n_out].
namespace node {

__global__ void __launch_bounds__(256) k_head(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                              const float* __restrict__ lw, const float* __restrict__ lb, float* __restrict__ out, int HW,
                                              int n_out, float eps) {
  __shared__ float pooled[64];
  const int tid = threadIdx.x, c = tid >> 2, q = tid & 3;
  const float* xc = x + ((size_t)blockIdx.x * 64 + c) * HW;
  float s = 0.f;
  for (int p = q; p < HW; p += 4) s += xc[p];
  s += __shfl_xor_sync(0xffffffffu, s, 1); s += __shfl_xor_sync(0xffffffffu, s, 2); s += __shfl_xor_sync(0xffffffffu, s, 4);
  const float inv_n = 1.0f / (float)(2 * HW);
  const float mean = s * inv_n;
  float d2 = 0.f;
  for (int p = q; p < HW; p += 4) { const float d = xc[p] - mean; d2 = fmaf(d, d, d2); }      // second pass: L1 / L2 hits
  d2 += __shfl_xor_sync(0xffffffffu, d2, 1); d2 += __shfl_xor_sync(0xffffffffu, d2, 2); d2 += __shfl_xor_sync(0xffffffffu, d2, 4);
  const float rstd = 1.0f / sqrtf(d2 * inv_n + eps);
  const float a = rstd * gamma[c], b = beta[c] - mean * a;
  float acc = 0.f;
  for (int p = q; p < HW; p += 4) acc += fmaxf(fmaf(xc[p], a, b), 0.f);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1); acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  const float avg = acc / (float)HW;
  if (lw == nullptr) {
    if (q == 0) out[(size_t)blockIdx.x * 64 + c] = avg;
    return;
  }
  if (q == 0) pooled[c] = avg;
  __syncthreads();
  if (tid < n_out) {
    float r = lb != nullptr ? lb[tid] : 0.f;
#pragma unroll 8
    for (int k = 0; k < 64; ++k) r = fmaf(pooled[k], lw[tid * 64 + k], r);
    out[(size_t)blockIdx.x * n_out + tid] = r;
  }
}

}  // namespace node

extern "C" int node_b200_head(const float* x, const float* gn_w, const float* gn_b, const float* lin_w, const float* lin_b, float* out,
                              int64_t N, int C, int HW, int n_out, float eps, void* stream) {
  if (N < 1 || N > 0x7fffffffLL || C != 64 || HW < 1 || (lin_w != nullptr && (n_out < 1 || n_out > 256))) return (int)cudaErrorInvalidValue;
  node::k_head<<<(unsigned)N, 256, 0, (cudaStream_t)stream>>>(x, gn_w, gn_b, lin_w, lin_b, out, HW, n_out, eps);
  return (int)cudaGetLastError();
}
