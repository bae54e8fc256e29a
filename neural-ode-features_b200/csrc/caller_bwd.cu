// SURVEY 8(f3), backward of the callers (reference model.py:119-178 under autograd, train.py:40-58): the memory-bound
// pieces. ATen runs the backward of GroupNorm -> ReLU as five kernels (threshold_backward, row moments kept from the
// forward, internal gradients, fused params, the elementwise apply, gamma/beta column sums: >= 4 reads + 2 writes of
// the tensor); here one CTA owns one GroupNorm cell, keeps x and the incoming gradient in registers, recomputes the
// statistics (two-pass, like native_group_norm) and writes the input gradient: 2 reads + 1 write.
//
// Also here: the parity-plane split / merge that turns the stride-2 convolutions of the ResBlock head into stride-1
// convolutions for the backward engines (caller_ops.py), the |max| reduction that scales a gradient tensor for the
// fp16 operand split, and the column sums of the per-image gamma / beta partials (fixed order: deterministic).
#include "node_common.cuh"

namespace node {

constexpr int kBwThreads = 128;

template <int NV>
__device__ __forceinline__ void bw_block_sum(float (&v)[NV], float* scratch) {
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) scratch[warp * NV + i] = v[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < kBwThreads / 32; ++w) r += scratch[w * NV + i];
    v[i] = r;
  }
}

// y = relu?(GN(x)); given gy = dL/dy:  gx = rstd * (gamma * g' - mean(gamma * g') - xhat * mean(gamma * g' * xhat)),
// g' = gy * [y > 0]; per-image partials pgamma[n][c] = sum g' * xhat, pbeta[n][c] = sum g'.
// VPT float4 vectors per thread; CPG channels per group (cell = CPG * HW floats, HW % 4 == 0 or VEC = 1).
template <int VPT, int VEC, int CPG>
__global__ void __launch_bounds__(kBwThreads) k_gn_relu_bwd(const float* __restrict__ x, const float* __restrict__ gy,
                                                             float* __restrict__ gx, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float* __restrict__ pgamma,
                                                             float* __restrict__ pbeta, int groups, int HW, float eps, int relu) {
  __shared__ float scratch[(kBwThreads / 32) * (2 + 2 * CPG)];
  constexpr int cpg = CPG;
  const int L = cpg * HW;
  const size_t base = (size_t)blockIdx.x * L;
  const int g = blockIdx.x % groups;
  const int C = groups * cpg;
  float v[VPT][VEC], d[VPT][VEC];
  float s[1] = {0.f};
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int e = (i * kBwThreads + threadIdx.x) * VEC;
    if (e < L) {
      if (VEC == 4) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(x + base + e));
        const float4 r = __ldg(reinterpret_cast<const float4*>(gy + base + e));
        v[i][0] = q.x; v[i][1 % VEC] = q.y; v[i][2 % VEC] = q.z; v[i][3 % VEC] = q.w;
        d[i][0] = r.x; d[i][1 % VEC] = r.y; d[i][2 % VEC] = r.z; d[i][3 % VEC] = r.w;
      } else {
        v[i][0] = __ldg(x + base + e);
        d[i][0] = __ldg(gy + base + e);
      }
    } else {
#pragma unroll
      for (int j = 0; j < VEC; ++j) { v[i][j] = 0.f; d[i][j] = 0.f; }
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) s[0] += v[i][j];
  }
  const float inv_n = 1.0f / (float)L;
  bw_block_sum<1>(s, scratch);
  const float mean = s[0] * inv_n;
  float q2[1] = {0.f};
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int e = (i * kBwThreads + threadIdx.x) * VEC;
    if (e < L) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) { const float t = v[i][j] - mean; q2[0] = fmaf(t, t, q2[0]); }
    }
  }
  bw_block_sum<1>(q2, scratch);
  const float rstd = 1.0f / sqrtf(q2[0] * inv_n + eps);
  float gam[CPG], bet[CPG];
#pragma unroll
  for (int c = 0; c < CPG; ++c) { gam[c] = __ldg(gamma + g * cpg + c); bet[c] = __ldg(beta + g * cpg + c); }
  // acc: [0] sum gamma g', [1] sum gamma g' xhat, [2 + c] sum g' xhat (channel c), [2 + CPG + c] sum g' (channel c)
  float acc[2 + 2 * CPG];
#pragma unroll
  for (int i = 0; i < 2 + 2 * CPG; ++i) acc[i] = 0.f;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int e = (i * kBwThreads + threadIdx.x) * VEC;
    if (e < L) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int cl = (e + j) / HW;
        float gm = gam[0], bt = bet[0];
#pragma unroll
        for (int c = 1; c < CPG; ++c) if (cl == c) { gm = gam[c]; bt = bet[c]; }
        const float xh = (v[i][j] - mean) * rstd;
        const float yv = fmaf(xh, gm, bt);
        const float gp = (relu && !(yv > 0.f)) ? 0.f : d[i][j];
        v[i][j] = xh;
        d[i][j] = gp * gm;
        acc[0] += d[i][j];
        acc[1] = fmaf(d[i][j], xh, acc[1]);
#pragma unroll
        for (int c = 0; c < CPG; ++c) if (cl == c) { acc[2 + c] = fmaf(gp, xh, acc[2 + c]); acc[2 + CPG + c] += gp; }
      }
    }
  }
  bw_block_sum<2 + 2 * CPG>(acc, scratch);
  const float m1 = acc[0] * inv_n, m2 = acc[1] * inv_n;
#pragma unroll
  for (int i = 0; i < VPT; ++i) {
    const int e = (i * kBwThreads + threadIdx.x) * VEC;
    if (e < L) {
      float o[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) o[j] = rstd * (d[i][j] - m1 - v[i][j] * m2);
      if (VEC == 4) *reinterpret_cast<float4*>(gx + base + e) = make_float4(o[0], o[1 % VEC], o[2 % VEC], o[3 % VEC]);
      else gx[base + e] = o[0];
    }
  }
  if (threadIdx.x < CPG) {
    const size_t n = blockIdx.x / groups;
    pgamma[n * C + g * cpg + threadIdx.x] = acc[2 + threadIdx.x];
    pbeta[n * C + g * cpg + threadIdx.x] = acc[2 + CPG + threadIdx.x];
  }
}

// out[c] = sum_n part[n][c] for `nvec` stacked [N][C] arrays: one CTA per (array, channel), float64, fixed order.
__global__ void __launch_bounds__(256) k_colsum(const float* __restrict__ part, float* __restrict__ out, int64_t N, int C) {
  __shared__ double red[256];
  const int c = blockIdx.x % C;
  const float* p = part + (size_t)(blockIdx.x / C) * (size_t)N * C;
  double s = 0.0;
  for (int64_t n = threadIdx.x; n < N; n += 256) s += (double)p[(size_t)n * C + c];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (float)red[0];
}

// max |v| over a tensor as the bit pattern of a non-negative float (ordered like unsigned integers); *out zeroed by the caller
__global__ void __launch_bounds__(256) k_absmax(const float* __restrict__ v, int64_t n, unsigned* __restrict__ out) {
  float m = 0.f;
  const int64_t n4 = n / 4;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 q = __ldg(reinterpret_cast<const float4*>(v) + i);
    m = fmaxf(m, fmaxf(fmaxf(fabsf(q.x), fabsf(q.y)), fmaxf(fabsf(q.z), fabsf(q.w))));
  }
  if (blockIdx.x == 0)
    for (int64_t i = n4 * 4 + threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, fabsf(v[i]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(out, __float_as_uint(m));      // NaN never wins: the scale stays finite
}

// planes[pr][pc][n][c][i][j] = a[n][c][2i + pr][2j + pc] (zero beyond the map): one thread per output element
__global__ void __launch_bounds__(256) k_plane_split(const float* __restrict__ a, float* __restrict__ planes, int64_t NC, int HI, int WI,
                                                     int HO, int WO) {
  const int64_t per = NC * HO * WO, total = 4 * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i / per);
    const int64_t r = i - (int64_t)p * per;
    const int64_t nc = r / (HO * WO);
    const int ij = (int)(r - nc * (HO * WO)), ii = ij / WO, jj = ij - ii * WO;
    const int y = 2 * ii + (p >> 1), xx = 2 * jj + (p & 1);
    planes[i] = (y < HI && xx < WI) ? __ldg(a + (size_t)nc * HI * WI + (size_t)y * WI + xx) : 0.f;
  }
}

// ga[n][c][y][x] = gplanes[y & 1][x & 1][n][c][y >> 1][x >> 1]
__global__ void __launch_bounds__(256) k_plane_merge(const float* __restrict__ gplanes, float* __restrict__ ga, int64_t NC, int HI, int WI,
                                                     int HO, int WO) {
  const int64_t total = NC * HI * WI, per = NC * HO * WO;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t nc = i / (HI * WI);
    const int yx = (int)(i - nc * (HI * WI)), y = yx / WI, xx = yx - y * WI;
    const int p = ((y & 1) << 1) | (xx & 1);
    ga[i] = __ldg(gplanes + (size_t)p * per + (size_t)nc * HO * WO + (size_t)(y >> 1) * WO + (xx >> 1));
  }
}

// even WI: a thread moves the two neighbours (2jj, 2jj+1) of a row - an 8-byte access on the full map, 4-byte accesses on the two
// column planes of the row's parity; blockIdx.x walks the (n, c) maps, so all index arithmetic is 32-bit and per map
__global__ void __launch_bounds__(256) k_plane_split2(const float* __restrict__ a, float* __restrict__ planes, int64_t NC, int HI, int WI,
                                                      int HO, int WO) {
  const size_t per = (size_t)NC * HO * WO;
  const int cells = HI * WO;
  for (int64_t nc = blockIdx.x; nc < NC; nc += gridDim.x) {
    const float2* src = reinterpret_cast<const float2*>(a + (size_t)nc * HI * WI);
    float* dst = planes + (size_t)nc * HO * WO;
    for (int i = threadIdx.x; i < cells; i += 256) {
      const int y = i / WO, jj = i - y * WO;
      const float2 v = __ldcs(src + i);
      float* p0 = dst + (size_t)((y & 1) << 1) * per + (y >> 1) * WO + jj;
      p0[0] = v.x;
      p0[per] = v.y;
    }
  }
}
__global__ void __launch_bounds__(256) k_plane_merge2(const float* __restrict__ gplanes, float* __restrict__ ga, int64_t NC, int HI, int WI,
                                                      int HO, int WO) {
  const size_t per = (size_t)NC * HO * WO;
  const int cells = HI * WO;
  for (int64_t nc = blockIdx.x; nc < NC; nc += gridDim.x) {
    float2* dst = reinterpret_cast<float2*>(ga + (size_t)nc * HI * WI);
    const float* src = gplanes + (size_t)nc * HO * WO;
    for (int i = threadIdx.x; i < cells; i += 256) {
      const int y = i / WO, jj = i - y * WO;
      const float* p0 = src + (size_t)((y & 1) << 1) * per + (y >> 1) * WO + jj;
      dst[i] = make_float2(__ldcs(p0), __ldcs(p0 + per));
    }
  }
}

template <int VPT, int VEC>
static int launch_gn_bwd(const float* x, const float* gy, float* gx, const float* gamma, const float* beta, float* pg, float* pb,
                         int64_t cells, int groups, int HW, float eps, int relu, cudaStream_t st) {
  k_gn_relu_bwd<VPT, VEC, 2><<<(unsigned)cells, kBwThreads, 0, st>>>(x, gy, gx, gamma, beta, pg, pb, groups, HW, eps, relu);
  return (int)cudaGetLastError();
}

}  // namespace node

extern "C" int node_b200_groupnorm_relu_backward(const float* x, const float* grad_out, float* grad_in, const float* gamma,
                                                 const float* beta, float* partials, float* grad_gamma, float* grad_beta, int64_t N,
                                                 int C, int groups, int HW, float eps, int relu, void* stream) {
  using namespace node;
  if (N < 1 || C < 1 || groups < 1 || C != 2 * groups || HW < 1) return (int)cudaErrorInvalidValue;    // 2 channels per group (GroupNorm(32, 64))
  const int64_t L = 2 * (int64_t)HW, cells = N * groups;
  if (cells > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  float* pg = partials;
  float* pb = partials + (size_t)N * C;
  const bool vec = HW % 4 == 0 && ((uintptr_t)x % 16 == 0) && ((uintptr_t)grad_out % 16 == 0) && ((uintptr_t)grad_in % 16 == 0);
  int rc;
  if (vec) {
    const int64_t nv = L / 4;
    if (nv <= kBwThreads) rc = launch_gn_bwd<1, 4>(x, grad_out, grad_in, gamma, beta, pg, pb, cells, groups, HW, eps, relu, st);
    else if (nv <= 2 * kBwThreads) rc = launch_gn_bwd<2, 4>(x, grad_out, grad_in, gamma, beta, pg, pb, cells, groups, HW, eps, relu, st);
    else if (nv <= 4 * kBwThreads) rc = launch_gn_bwd<4, 4>(x, grad_out, grad_in, gamma, beta, pg, pb, cells, groups, HW, eps, relu, st);
    else return (int)cudaErrorInvalidValue;
  } else {
    if (L <= kBwThreads) rc = launch_gn_bwd<1, 1>(x, grad_out, grad_in, gamma, beta, pg, pb, cells, groups, HW, eps, relu, st);
    else if (L <= 4 * kBwThreads) rc = launch_gn_bwd<4, 1>(x, grad_out, grad_in, gamma, beta, pg, pb, cells, groups, HW, eps, relu, st);
    else if (L <= 16 * kBwThreads) rc = launch_gn_bwd<16, 1>(x, grad_out, grad_in, gamma, beta, pg, pb, cells, groups, HW, eps, relu, st);
    else return (int)cudaErrorInvalidValue;
  }
  if (rc != 0) return rc;
  // grad_gamma then grad_beta: the two [N][C] partial arrays are stacked, the outputs must be too (or separate calls)
  k_colsum<<<C, 256, 0, st>>>(pg, grad_gamma, N, C);
  k_colsum<<<C, 256, 0, st>>>(pb, grad_beta, N, C);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_absmax(const float* v, int64_t n, unsigned* out_bits, void* stream) {
  if (n < 1) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  NODE_CUDA_OK(cudaMemsetAsync(out_bits, 0, sizeof(unsigned), st));
  int64_t blocks = (n / 4 + 255) / 256;
  blocks = blocks < 1 ? 1 : (blocks > 148 * 8 ? 148 * 8 : blocks);
  node::k_absmax<<<(unsigned)blocks, 256, 0, st>>>(v, n, out_bits);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_plane_split(const float* a, float* planes, int64_t N, int C, int HI, int WI, void* stream) {
  if (N < 1 || C < 1 || HI < 1 || WI < 1) return (int)cudaErrorInvalidValue;
  const int HO = (HI - 1) / 2 + 1, WO = (WI - 1) / 2 + 1;
  const int64_t total = 4 * N * C * HO * WO;
  int64_t blocks = (total + 255) / 256;
  blocks = blocks > 148 * 32 ? 148 * 32 : blocks;
  if ((WI & 1) == 0 && ((uintptr_t)a % 8 == 0) && HO * 2 == HI)
    node::k_plane_split2<<<(unsigned)(N * C < 148 * 64 ? N * C : 148 * 64), 256, 0, (cudaStream_t)stream>>>(a, planes, N * C, HI, WI, HO, WO);
  else
    node::k_plane_split<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a, planes, N * C, HI, WI, HO, WO);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_plane_merge(const float* gplanes, float* ga, int64_t N, int C, int HI, int WI, void* stream) {
  if (N < 1 || C < 1 || HI < 1 || WI < 1) return (int)cudaErrorInvalidValue;
  const int HO = (HI - 1) / 2 + 1, WO = (WI - 1) / 2 + 1;
  const int64_t total = N * C * HI * WI;
  int64_t blocks = (total + 255) / 256;
  blocks = blocks > 148 * 32 ? 148 * 32 : blocks;
  if ((WI & 1) == 0 && ((uintptr_t)ga % 8 == 0) && HO * 2 == HI)
    node::k_plane_merge2<<<(unsigned)(N * C < 148 * 64 ? N * C : 148 * 64), 256, 0, (cudaStream_t)stream>>>(gplanes, ga, N * C, HI, WI, HO, WO);
  else
    node::k_plane_merge<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(gplanes, ga, N * C, HI, WI, HO, WO);
  return (int)cudaGetLastError();
}

// ---- stem backward: conv0 (Conv2d(CIN, 64, 3, 1), bias) -> GroupNorm(32, 64) -> ReLU (model.py:119-178 under autograd) --------
// Given go = dL/d relu(GN(conv0(x))) the kernel produces the gradients of conv0.weight / conv0.bias / gamma / beta and never
// materialises the [N,64,HO,WO] conv output or its gradient: like the forward stem kernel (caller_ops.cu) a CTA recomputes
// the convolution of ONE image 8 channels at a time from the shared-memory copy of the image, redoes the GroupNorm
// statistics, forms gc = dL/d conv0(x) in registers, parks the 8 x NPIX tile in shared memory and lets thread (c, k)
// accumulate dW[c][k] += sum_p gc[c][p] * x[p + off(k)] over it. CTAs are persistent over images and keep their weight
// gradient partials in registers; a second kernel folds the per-CTA partials in a fixed order (deterministic).
namespace node {

constexpr int kSbThreads = 256;

template <int CIN, int HIN, int WIN>
struct StemBwdGeom {
  static constexpr int PITCH = 35;
  static constexpr int CS = HIN * PITCH + (9 - (HIN * PITCH) % 32 + 32) % 32;
  static constexpr int K = CIN * 9, NPIX = (HIN - 2) * (WIN - 2);
  static constexpr int smem = (2 * 64 * K + CIN * CS + 8 * NPIX) * 4;
  static_assert(WIN <= PITCH, "row pitch");
};

template <int NV>
__device__ __forceinline__ void sb_block_sum(float (&v)[NV], float* scratch, float* total) {
#pragma unroll
  for (int i = 0; i < NV; ++i)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0)
#pragma unroll
    for (int i = 0; i < NV; ++i) scratch[warp * NV + i] = v[i];
  __syncthreads();
  if (threadIdx.x < NV) {
    float r = 0.f;
#pragma unroll
    for (int w = 0; w < kSbThreads / 32; ++w) r += scratch[w * NV + threadIdx.x];
    total[threadIdx.x] = r;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = total[i];
}

// partial layout per CTA: [64 * K] dW, [64] dbias, [64] dgamma, [64] dbeta
template <int CIN, int HIN, int WIN>
__global__ void __launch_bounds__(kSbThreads, 2)
k_stem_bwd(const float* __restrict__ x, const float* __restrict__ cw, const float* __restrict__ cb, const float* __restrict__ gamma,
           const float* __restrict__ beta, const float* __restrict__ go, float* __restrict__ part, int N, float eps) {
  constexpr int HO = HIN - 2, WO = WIN - 2, NPIX = HO * WO, K = CIN * 9, NT = kSbThreads, PPT = (NPIX + NT - 1) / NT;
  constexpr int PSTRIDE = 64 * K + 3 * 64;
  // shared-memory image: row pitch 35 and channel stride = 9 (mod 32) put the 27 patch elements (ci, ky, kx) of the
  // weight-gradient threads of a warp on 27 different banks (pitch 32 / stride 1024 would fold them onto 3)
  constexpr int PITCH = StemBwdGeom<CIN, HIN, WIN>::PITCH, CS = StemBwdGeom<CIN, HIN, WIN>::CS;
  static_assert(8 * K <= NT, "one thread per (channel of the pass, k)");
  extern __shared__ __align__(16) float sb_dyn[];
  float* s_w = sb_dyn;                               // [pass of 8 channels][k][8 channels]
  float* s_in = s_w + 64 * K;                        // the image
  float* s_gc = s_in + CIN * CS;                     // gc tile of the current pass
  float* s_dw = s_gc + 8 * NPIX;                     // [64][K] weight-gradient partial of this CTA
  __shared__ float s_b[64], s_g[64], s_be[64];
  __shared__ float scratch[(NT / 32) * 32], total[32], s_var[4];
  const int tid = threadIdx.x;
  for (int i = tid; i < 64 * K; i += NT) {
    const int c = i / K, k = i % K;
    s_w[((c >> 3) * K + k) * 8 + (c & 7)] = cw[i];
  }
  if (tid < 64) { s_b[tid] = cb[tid]; s_g[tid] = gamma[tid]; s_be[tid] = beta[tid]; }
  int base[PPT];
  bool valid[PPT];
#pragma unroll
  for (int q = 0; q < PPT; ++q) {
    const int p = tid + q * NT;
    valid[q] = p < NPIX;
    const int pp = valid[q] ? p : 0;
    base[q] = (pp / WO) * PITCH + pp % WO;
  }
  // thread (wc, wk) of the weight-gradient phase: channel wc of the pass, patch element wk
  const int wc = tid / K, wk = tid % K;
  const bool wthread = tid < 8 * K;
  const int woff = (wk / 9) * CS + ((wk % 9) / 3) * PITCH + wk % 3;
  float dchan[3] = {0.f, 0.f, 0.f};                  // threads < 64: dbias, dgamma, dbeta of channel tid
  for (int i = tid; i < 64 * K; i += NT) s_dw[i] = 0.f;      // every (channel, k) entry is owned by one thread: no atomics
  constexpr float inv_n = 1.0f / (float)(2 * NPIX);
#pragma unroll 1
  for (int n = blockIdx.x; n < N; n += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < CIN * HIN * WIN; i += NT) {
      const int c = i / (HIN * WIN), rem = i % (HIN * WIN);
      s_in[c * CS + (rem / WIN) * PITCH + rem % WIN] = x[(size_t)n * CIN * HIN * WIN + i];
    }
    __syncthreads();
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
        const int off = (k / 9) * CS + ((k % 9) / 3) * PITCH + k % 3;
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
          const float v = s_in[base[q] + off];
          acc[q][0] = fmaf(v, a.x, acc[q][0]); acc[q][1] = fmaf(v, a.y, acc[q][1]); acc[q][2] = fmaf(v, a.z, acc[q][2]); acc[q][3] = fmaf(v, a.w, acc[q][3]);
          acc[q][4] = fmaf(v, b.x, acc[q][4]); acc[q][5] = fmaf(v, b.y, acc[q][5]); acc[q][6] = fmaf(v, b.z, acc[q][6]); acc[q][7] = fmaf(v, b.w, acc[q][7]);
        }
      }
      // statistics of the 4 cells: means, then two-pass variances (native_group_norm)
      float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int q = 0; q < PPT; ++q)
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (valid[q]) s[g] += acc[q][2 * g] + acc[q][2 * g + 1];
      sb_block_sum<4>(s, scratch, total);
      float mean[4], rstd[4];
      float qq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int g = 0; g < 4; ++g) mean[g] = s[g] * inv_n;
#pragma unroll
      for (int q = 0; q < PPT; ++q)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const float d0 = acc[q][2 * g] - mean[g], d1 = acc[q][2 * g + 1] - mean[g];
          if (valid[q]) qq[g] += fmaf(d0, d0, d1 * d1);
        }
      sb_block_sum<4>(qq, scratch, total);
#pragma unroll
      for (int g = 0; g < 4; ++g) rstd[g] = 1.0f / sqrtf(qq[g] * inv_n + eps);
      if (tid < 4) s_var[tid] = total[tid] * inv_n;      // total = the squared-deviation sums just folded
      // incoming gradient through the ReLU; r: [g] sum gamma g', [4 + g] sum gamma g' xhat, [8 + j] sum g' xhat, [16 + j] sum g', [24 + j] sum xhat
      float r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = 0.f;
      float gp[PPT][8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = 8 * ps + j, g = j >> 1;
        const float gm = s_g[c], bt = s_be[c];
#pragma unroll
        for (int q = 0; q < PPT; ++q) {
          float gv = 0.f, xh = 0.f;
          if (valid[q]) {
            xh = (acc[q][j] - mean[g]) * rstd[g];
            const float yv = fmaf(xh, gm, bt);
            const float gin = __ldg(go + ((size_t)n * 64 + c) * NPIX + tid + q * NT);
            gv = yv > 0.f ? gin : 0.f;
          }
          acc[q][j] = xh;
          gp[q][j] = gv * gm;
          r[g] += gp[q][j];
          r[4 + g] = fmaf(gp[q][j], xh, r[4 + g]);
          r[8 + j] = fmaf(gv, xh, r[8 + j]);
          r[16 + j] += gv;
          r[24 + j] += xh;
        }
      }
      sb_block_sum<32>(r, scratch, total);
      // gc = rstd * (gamma g' - m1 - xhat m2) -> shared tile [channel of the pass][pixel]
      __syncthreads();                               // the previous pass' tile has been consumed
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int g = j >> 1;
        const float m1 = r[g] * inv_n, m2 = r[4 + g] * inv_n;
#pragma unroll
        for (int q = 0; q < PPT; ++q)
          if (valid[q]) s_gc[j * NPIX + tid + q * NT] = rstd[g] * (gp[q][j] - m1 - acc[q][j] * m2);
      }
      if (tid >= 8 * ps && tid < 8 * ps + 8) {       // thread c = tid owns the sums of channel c; `total` = the block sums above
        const int j = tid - 8 * ps, g = j >> 1;
        const float m1 = total[g] * inv_n, m2 = total[4 + g] * inv_n;
        const float rs = 1.0f / sqrtf(s_var[g] + eps);
        // dbias = sum gc = rstd (gamma sum g' - npix m1 - m2 sum xhat)
        dchan[0] += rs * (s_g[tid] * total[16 + j] - (float)NPIX * m1 - m2 * total[24 + j]);
        dchan[1] += total[8 + j];
        dchan[2] += total[16 + j];
      }
      __syncthreads();
      if (wthread) {
        const float* gt = s_gc + wc * NPIX;
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 2
        for (int y = 0; y < HO; ++y) {
          const float* gi = gt + y * WO;
          const float* xi = s_in + woff + y * PITCH;
#pragma unroll 6
          for (int xx = 0; xx + 1 < WO; xx += 2) { a0 = fmaf(gi[xx], xi[xx], a0); a1 = fmaf(gi[xx + 1], xi[xx + 1], a1); }
          if (WO & 1) a0 = fmaf(gi[WO - 1], xi[WO - 1], a0);
        }
        s_dw[(8 * ps + wc) * K + wk] += a0 + a1;
      }
    }
  }
  float* pp = part + (size_t)blockIdx.x * PSTRIDE;
  __syncthreads();
  for (int i = tid; i < 64 * K; i += NT) pp[i] = s_dw[i];
  if (tid < 64) { pp[64 * K + tid] = dchan[0]; pp[64 * K + 64 + tid] = dchan[1]; pp[64 * K + 128 + tid] = dchan[2]; }
}

__global__ void __launch_bounds__(256) k_stem_bwd_fold(const float* __restrict__ part, int nblocks, int pstride, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pstride) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += (double)part[(size_t)b * pstride + i];
  out[i] = (float)s;
}

}  // namespace node

extern "C" int64_t node_b200_stem_backward_workspace_bytes(int CIN) {
  return (int64_t)(148 * 2) * (64 * CIN * 9 + 3 * 64) * 4;
}

extern "C" int node_b200_stem_backward(const float* x, const float* conv_w, const float* conv_b, const float* gn_w, const float* gn_b,
                                       const float* grad_out, void* workspace, float* grads, int N, int CIN, int HIN, int WIN,
                                       float eps, void* stream) {
  using namespace node;
  if (N < 1) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = N < 148 * 2 ? N : 148 * 2;
  const int pstride = 64 * CIN * 9 + 3 * 64;
  float* part = (float*)workspace;
  if (CIN == 3 && HIN == 32 && WIN == 32) {
    constexpr int smem = StemBwdGeom<3, 32, 32>::smem;
    NODE_SET_SMEM_ONCE((k_stem_bwd<3, 32, 32>), smem);
    k_stem_bwd<3, 32, 32><<<grid, kSbThreads, smem, st>>>(x, conv_w, conv_b, gn_w, gn_b, grad_out, part, N, eps);
  } else if (CIN == 1 && HIN == 28 && WIN == 28) {
    constexpr int smem = StemBwdGeom<1, 28, 28>::smem;
    NODE_SET_SMEM_ONCE((k_stem_bwd<1, 28, 28>), smem);
    k_stem_bwd<1, 28, 28><<<grid, kSbThreads, smem, st>>>(x, conv_w, conv_b, gn_w, gn_b, grad_out, part, N, eps);
  } else {
    return (int)cudaErrorInvalidValue;
  }
  NODE_CUDA_OK(cudaGetLastError());
  k_stem_bwd_fold<<<(pstride + 255) / 256, 256, 0, st>>>(part, grid, pstride, grads);
  return (int)cudaGetLastError();
}
