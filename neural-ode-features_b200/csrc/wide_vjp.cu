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
// in:  x (+ bias + t * tmap), gscale * g = dL/d(output); out: gx = dL/dx, pgamma / pbeta [N][C] per-image partials of dL/dgamma, dL/dbeta.
__global__ void __launch_bounds__(kWvThreads) k_gn_backward_ex(const float* __restrict__ x, const float* __restrict__ g,
                                                                float* __restrict__ gx, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, const float* __restrict__ add_bias,
                                                                const float* __restrict__ add_tmap, const float* __restrict__ t_dev,
                                                                float tsign, float gscale, float* __restrict__ pgamma,
                                                                float* __restrict__ pbeta, int groups, int cpg, int HW, float eps, int relu) {
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
    float d = __ldg(g + base + e) * gscale;
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

// [N, C, HW] -> [C / 64][N][64][HW]: dense 64-channel blocks for the weight-gradient GEMM (T = float4 when HW % 4 == 0)
template <typename T>
__global__ void __launch_bounds__(256) k_to_blocks(const T* __restrict__ x, T* __restrict__ out, int64_t N, int C, int HWv) {
  const int64_t total = N * C * HWv, nthr = (int64_t)gridDim.x * blockDim.x;
  const int64_t per_img = (int64_t)C * HWv, per_blk = (int64_t)64 * HWv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += nthr) {
    const int64_t n = i / per_img, r = i % per_img;
    const int64_t b = r / per_blk, q = r % per_blk;
    out[(b * N + n) * per_blk + q] = __ldg(x + i);
  }
}

// dW[co][1 + ci][tap] <- pair (co / 64, ci / 64) of the block weight gradients [nb * nb][64][64][9]
__global__ void __launch_bounds__(256) k_assemble_dw(const float* __restrict__ pairs, float* __restrict__ w_out, int C) {
  const int64_t total = (int64_t)C * C * 9;
  const int nb = C / 64;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % 9), ci = (int)((i / 9) % C), co = (int)(i / (9 * (int64_t)C));
    const int64_t src = ((((int64_t)(co / 64) * nb + ci / 64) * 64 + (co % 64)) * 64 + (ci % 64)) * 9 + tap;
    w_out[((int64_t)co * (C + 1) + 1 + ci) * 9 + tap] = __ldg(pairs + src);
  }
}

// One warp per channel c: b_out[c] = sum_pix S, dW[c][0][tap] = t * sum_pix S * valid(tap, pix) (the folded time channel of
// ConcatConv2d, model.py:320-323), vt_part[c] = sum_pix S * Tmap (this convolution's share of dL/dt through the time channel)
__global__ void __launch_bounds__(32) k_time_bias_grads(const float* __restrict__ S, const float* __restrict__ tmap, const float* __restrict__ t_dev,
                                                         float tsign, int C, int H, int W, float* __restrict__ b_out,
                                                         float* __restrict__ w_out, double* __restrict__ vt_part) {
  const int c = blockIdx.x, lane = threadIdx.x, HW = H * W;
  double acc[11];
#pragma unroll
  for (int i = 0; i < 11; ++i) acc[i] = 0.0;
  for (int p = lane; p < HW; p += 32) {
    const float s = __ldg(S + (size_t)c * HW + p);
    const int y = p / W, x = p % W;
    acc[9] += (double)s;
    acc[10] += (double)s * (double)__ldg(tmap + (size_t)c * HW + p);
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int yy = y + tap / 3 - 1, xx = x + tap % 3 - 1;
      if (yy >= 0 && yy < H && xx >= 0 && xx < W) acc[tap] += (double)s;
    }
  }
#pragma unroll
  for (int i = 0; i < 11; ++i) acc[i] = warp_sum(acc[i]);
  if (lane == 0) {
    const float t = tsign * __ldg(t_dev);
    for (int tap = 0; tap < 9; ++tap) w_out[(size_t)c * (C + 1) * 9 + tap] = t * (float)acc[tap];
    b_out[c] = (float)acc[9];
    vt_part[c] = acc[10];
  }
}

__global__ void __launch_bounds__(32) k_sum_vt(const double* __restrict__ part, int n, float* __restrict__ out) {
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) s += part[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) *out = (float)s;
}

}  // namespace node

static int gn_backward_ex(const float* x, const float* grad_out, float gscale, float* grad_in, const float* gamma, const float* beta,
                          const float* add_bias, const float* add_tmap, const float* t_dev, float tsign, float* partials,
                          float* grad_gamma, float* grad_beta, int64_t N, int C, int groups, int HW, float eps, int relu, cudaStream_t st) {
  using namespace node;
  if (N < 1 || C < 1 || groups < 1 || C % groups != 0 || HW < 1) return (int)cudaErrorInvalidValue;
  const int cpg = C / groups;
  if ((int64_t)cpg * HW > 4096 || N * groups > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
  float* pg = partials;
  float* pb = partials + N * C;
  k_gn_backward_ex<<<(unsigned)(N * groups), kWvThreads, 0, st>>>(x, grad_out, grad_in, gamma, beta, add_bias, add_tmap, t_dev, tsign, gscale,
                                                                  pg, pb, groups, cpg, HW, eps, relu);
  NODE_CUDA_OK(cudaGetLastError());
  k_batch_colsum<<<(C + 31) / 32, 256, 0, st>>>(pg, grad_gamma, N, C);
  k_batch_colsum<<<(C + 31) / 32, 256, 0, st>>>(pb, grad_beta, N, C);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_groupnorm_backward_ex(const float* x, const float* grad_out, float* grad_in, const float* gamma,
                                               const float* beta, const float* add_bias, const float* add_tmap, const float* t_dev,
                                               float tsign, float* partials, float* grad_gamma, float* grad_beta, int64_t N, int C,
                                               int groups, int HW, float eps, int relu, void* stream) {
  return gn_backward_ex(x, grad_out, 1.0f, grad_in, gamma, beta, add_bias, add_tmap, t_dev, tsign, partials, grad_gamma, grad_beta, N, C,
                        groups, HW, eps, relu, (cudaStream_t)stream);
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

// ---- one evaluation of the wide augmented dynamics by ONE call ---------------------------------------------------------------------
// p[] (device pointers unless noted): 0 y, 1 adj_y, 2 t_dev, 3 f_out, 4 vjp_y, 5 vjp_t, 6 vjp_params (flat, func.parameters() order),
// 7..12 norm1.w norm1.b norm2.w norm2.b norm3.w norm3.b, 13 conv1.bias, 14 conv2.bias, 15 Tmap1, 16 Tmap2 ([C, H, W] each),
// 17 block workspaces forward [2][nb][nb] (block path), 18 block workspaces data gradient [2][nb][nb], 19 wide8 workspace forward,
// 20 wide8 workspace data gradient, 21 wide8 operand image, 22 a1, 23 c1, 24 a2, 25 c2, 26 gc2, 27 gr, 28 gc1 ([N, C, H, W] each),
// 29 ab, 30 gb ([nb][N][64][H][W]), 31 GroupNorm partials (2 N C floats), 32 S (C H W floats), 33 max bits (2 unsigned), 34 scale
// (1 float), 35 wgrad workspace (node_b200_conv_wgrad_workspace_bytes(6)), 36 block weight gradients (nb^2 x 64 x 64 x 9 floats),
// 37 vt partials (2 C doubles).  dims: N, C, H, W, block workspace stride (bytes), use_wide8.
extern "C" int node_b200_wide_vjp(void* const* p, const int64_t* dims, float tsign, void* stream) {
  using namespace node;
  const int N = (int)dims[0], C = (int)dims[1], H = (int)dims[2], W = (int)dims[3];
  const int64_t bstride = dims[4];
  const bool w8 = dims[5] != 0;
  if (N < 1 || C % 64 != 0 || C <= 64) return (int)cudaErrorInvalidValue;
  const int nb = C / 64, HW = H * W;
  const int64_t E = (int64_t)N * C * HW;
  cudaStream_t st = (cudaStream_t)stream;
  auto F = [&](int i) { return (float*)p[i]; };
  const float* t_dev = F(2);
  const float s = tsign < 0 ? -1.0f : 1.0f;
  // flat parameter gradient: norm1.w, norm1.b, conv1.W, conv1.b, norm2.w, norm2.b, conv2.W, conv2.b, norm3.w, norm3.b (misc.py:5-7)
  float* vp = F(6);
  const int64_t nw = (int64_t)C * (C + 1) * 9;
  float* g1w = vp; float* g1b = g1w + C; float* w1 = g1b + C; float* b1 = w1 + nw;
  float* g2w = b1 + C; float* g2b = g2w + C; float* w2 = g2b + C; float* b2 = w2 + nw;
  float* g3w = b2 + C; float* g3b = g3w + C;
  double* vt_part = (double*)p[37];

  auto gn = [&](const float* x, float* y, int ni, const float* bias, const float* tmap, float post, int relu) {
    return node_b200_groupnorm_relu_ex(x, y, F(7 + 2 * ni), F(8 + 2 * ni), bias, tmap, t_dev, tsign, post, N, C, 32, HW, 1e-5f, relu, stream);
  };
  auto conv_fwd = [&](int li, const float* act_in_raw, int ni, const float* bias, const float* act, float* out) -> int {
    if (w8) {       // operand image straight from the GroupNorm input (fused GroupNorm -> ReLU -> fp16 split), one implicit GEMM
      NODE_CUDA_OK((cudaError_t)node_b200_wide8_gn_operand(p[19], li, act_in_raw, p[21], F(7 + 2 * ni), F(8 + 2 * ni), bias, t_dev, tsign, N, C, stream));
      return node_b200_wide8_conv(p[19], li, p[21], out, N, C, stream);
    }
    return node_b200_wide_conv_blocks((char*)p[17] + (int64_t)li * nb * nb * bstride, bstride, act, out, N, C, H, W, stream);
  };
  auto conv_dgrad = [&](int li, const float* gc, float* out) -> int {
    if (w8) {
      NODE_CUDA_OK((cudaError_t)node_b200_wide8_raw_operand(p[20], li, gc, p[21], (const unsigned*)p[33] + 1, N, C, stream));
      return node_b200_wide8_conv(p[20], li, p[21], out, N, C, stream);
    }
    return node_b200_wide_conv_blocks((char*)p[18] + (int64_t)li * nb * nb * bstride, bstride, gc, out, N, C, H, W, stream);
  };
  auto param_grads = [&](int li, const float* act, const float* gc, float* w_out, float* b_out) -> int {
    const int grid = (int)((E / 4 + 255) / 256 < 148 * 8 ? (E / 4 + 255) / 256 : 148 * 8);
    if (HW % 4 == 0) {
      k_to_blocks<float4><<<grid, 256, 0, st>>>((const float4*)act, (float4*)p[29], N, C, HW / 4);
      k_to_blocks<float4><<<grid, 256, 0, st>>>((const float4*)gc, (float4*)p[30], N, C, HW / 4);
    } else {
      k_to_blocks<float><<<grid, 256, 0, st>>>(act, F(29), N, C, HW);
      k_to_blocks<float><<<grid, 256, 0, st>>>(gc, F(30), N, C, HW);
    }
    NODE_CUDA_OK(cudaGetLastError());
    NODE_CUDA_OK((cudaError_t)node_b200_absmax(act, E, (unsigned*)p[33], stream));
    NODE_CUDA_OK((cudaError_t)node_b200_pow2_scale((const unsigned*)p[33], F(34), stream));
    NODE_CUDA_OK((cudaError_t)node_b200_absmax(gc, E, (unsigned*)p[33] + 1, stream));      // also the data gradient's operand scale
    const int64_t blk = (int64_t)N * 64 * HW;
    for (int k = 0; k < nb * nb; k += 6) {
      const int n = nb * nb - k < 6 ? nb * nb - k : 6;
      const float* in[6]; const float* gr[6]; const float* sc[6]; const unsigned* mb[6];
      for (int j = 0; j < n; ++j) {
        const int o = (k + j) / nb, i = (k + j) % nb;
        in[j] = F(29) + i * blk; gr[j] = F(30) + o * blk; sc[j] = F(34); mb[j] = (const unsigned*)p[33] + 1;
      }
      NODE_CUDA_OK((cudaError_t)node_b200_conv_wgrad(p[35], n, in, gr, sc, mb, F(36) + (int64_t)k * 64 * 64 * 9, N, 64, H, W, stream));
    }
    k_assemble_dw<<<148 * 4, 256, 0, st>>>(F(36), w_out, C);
    NODE_CUDA_OK((cudaError_t)node_b200_batch_colsum(gc, F(32), N, (int64_t)C * HW, stream));
    k_time_bias_grads<<<C, 32, 0, st>>>(F(32), F(15 + li), t_dev, tsign, C, H, W, b_out, w_out, vt_part + (int64_t)li * C);
    return (int)cudaGetLastError();
  };

  // forward, keeping the activations (model.py:339-348)
  NODE_CUDA_OK((cudaError_t)gn(F(0), F(22), 0, nullptr, nullptr, 1.0f, 1));
  NODE_CUDA_OK((cudaError_t)conv_fwd(0, F(0), 0, nullptr, F(22), F(23)));
  NODE_CUDA_OK((cudaError_t)gn(F(23), F(24), 1, F(13), F(15), 1.0f, 1));
  NODE_CUDA_OK((cudaError_t)conv_fwd(1, F(23), 1, F(13), F(24), F(25)));
  NODE_CUDA_OK((cudaError_t)gn(F(25), F(3), 2, F(14), F(16), s, 0));
  // backward with cotangent -tsign * adj_y on ODEfunc's output (adjoint.py:40-49; misc.py:186 negates the reversed system)
  NODE_CUDA_OK((cudaError_t)gn_backward_ex(F(25), F(1), -s, F(26), F(11), F(12), F(14), F(16), t_dev, tsign, F(31), g3w, g3b, N, C, 32, HW, 1e-5f, 0, st));
  NODE_CUDA_OK((cudaError_t)param_grads(1, F(24), F(26), w2, b2));
  NODE_CUDA_OK((cudaError_t)conv_dgrad(1, F(26), F(27)));
  NODE_CUDA_OK((cudaError_t)gn_backward_ex(F(23), F(27), 1.0f, F(28), F(9), F(10), F(13), F(15), t_dev, tsign, F(31), g2w, g2b, N, C, 32, HW, 1e-5f, 1, st));
  NODE_CUDA_OK((cudaError_t)param_grads(0, F(22), F(28), w1, b1));
  NODE_CUDA_OK((cudaError_t)conv_dgrad(0, F(28), F(27)));
  NODE_CUDA_OK((cudaError_t)gn_backward_ex(F(0), F(27), 1.0f, F(4), F(7), F(8), nullptr, nullptr, t_dev, tsign, F(31), g1w, g1b, N, C, 32, HW, 1e-5f, 1, st));
  k_sum_vt<<<1, 32, 0, st>>>(vt_part, 2 * C, F(5));
  return (int)cudaGetLastError();
}
