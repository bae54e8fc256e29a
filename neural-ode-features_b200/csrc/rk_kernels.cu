// K2-K6: Runge-Kutta stage combination, error norm, initial-step norms, dense output and the
// device-side step-size controller of the dopri5 hot path.  Memory-bound kernels: 128-bit
// vectorised, coalesced, grid sized in multiples of the SM count, warp-shuffle reductions,
// deterministic two-level sums in float64.
//
// Reference behaviour restated (paths under /root/reference/torchdiffeq/torchdiffeq/_impl/):
//   rk_common.py:22-61, misc.py:22-30,71-76,84-170, dopri5.py:39-45,77-122, interp.py:5-65.
#include "node_common.cuh"

namespace node {

constexpr int kThreads = 256;
constexpr int kNormThreads = 512;     // the read-only norm kernel wants more bytes in flight per SM than the grid cap of 296 CTAs gives at 256
constexpr int kPartialBlocks = 296;  // 2 x 148 SMs

struct KPtrs { const void* p[7]; };
struct Segs { int64_t off[NODE_MAX_SEG]; int64_t len[NODE_MAX_SEG]; };

template <typename T> struct Vec;
template <> struct Vec<float> { using type = float4; static constexpr int n = 4; };
template <> struct Vec<double> { using type = double2; static constexpr int n = 2; };

template <typename T> __device__ __forceinline__ void vload(const T* p, T (&v)[Vec<T>::n]);
template <> __device__ __forceinline__ void vload<float>(const float* p, float (&v)[4]) {
  const float4 q = *reinterpret_cast<const float4*>(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
}
template <> __device__ __forceinline__ void vload<double>(const double* p, double (&v)[2]) {
  const double2 q = *reinterpret_cast<const double2*>(p); v[0] = q.x; v[1] = q.y;
}
template <typename T> __device__ __forceinline__ void vstore(T* p, const T (&v)[Vec<T>::n]);
template <> __device__ __forceinline__ void vstore<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void vstore<double>(double* p, const double (&v)[2]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}

// ---- K2: out = y0 + sum_j (h*c_j)*k_j ------------------------------------------------------
// ROW is a template parameter: the coefficient row, its non-zero entries and therefore the set of source tensors
// are compile-time, so every 128-bit load of an iteration (y0 + up to 6 k's, two vectors each) is issued before the
// first use - the kernel is bound by HBM, not by load latency.
template <typename T, int ROW>
__global__ void __launch_bounds__(kThreads) k_stage_combine(const node_ctl_t* __restrict__ ctl, T* __restrict__ out,
                                                            const T* __restrict__ y0, KPtrs ks, int64_t n) {
  using A = Arith<T>;
  constexpr int V = Vec<T>::n;
  constexpr int U = 2;                      // vectors per thread per iteration
  // the mid-point only feeds the dense output: nothing to do unless the controller scheduled outputs for the accepted step
  // (callers that decide on the host launch row 6 only then; node_b200_adjoint_solve launches it every attempt)
  if (ROW == 6 && ctl->out_hi <= ctl->out_lo) return;
  // row 7: probe step h0; row 6 (dense-output mid-point) runs AFTER the controller accepted the step, when
  // ctl->h* already hold the next attempt's size, so it takes the accepted step's h saved in it_h*.
  const T h = ROW == 7 ? (sizeof(T) == 4 ? (T)ctl->h0_32 : (T)ctl->h0)
            : ROW == 6 ? (sizeof(T) == 4 ? (T)ctl->it_h32 : (T)ctl->it_h64) : ctl_h<T>(ctl);
  T hc[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) hc[j] = A::mul(h, (T)kCoef(ROW, j));
  const int64_t nvec = n / V;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (int64_t i0 = tid; i0 < nvec; i0 += nthr * U) {
    T kv[U][7][V], y[U][V];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * nthr;
      if (i < nvec) {
        vload<T>(y0 + i * V, y[u]);
#pragma unroll
        for (int j = 0; j < 7; ++j)
          if (j < kRowLen(ROW) && kCoef(ROW, j) != 0.0) vload<T>(reinterpret_cast<const T*>(ks.p[j]) + i * V, kv[u][j]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * nthr;
      if (i < nvec) {
        T acc[V];
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = (T)0;
#pragma unroll
        for (int j = 0; j < 7; ++j) {        // reference order, zero coefficients skipped (adding +-0 changes nothing)
          if (j < kRowLen(ROW) && kCoef(ROW, j) != 0.0) {
#pragma unroll
            for (int e = 0; e < V; ++e) acc[e] = A::add(acc[e], A::mul(hc[j], kv[u][j][e]));
          }
        }
#pragma unroll
        for (int e = 0; e < V; ++e) acc[e] = A::add(y[u][e], acc[e]);
        vstore<T>(out + i * V, acc);
      }
    }
  }
  // scalar tail
  for (int64_t i = nvec * V + tid; i < n; i += nthr) {
    T acc = (T)0;
#pragma unroll
    for (int j = 0; j < 7; ++j)
      if (j < kRowLen(ROW) && kCoef(ROW, j) != 0.0) acc = A::add(acc, A::mul(hc[j], reinterpret_cast<const T*>(ks.p[j])[i]));
    out[i] = A::add(y0[i], acc);
  }
}

template <typename T>
static void launch_stage_combine(int row, int grid, cudaStream_t st, const node_ctl_t* ctl, T* out, const T* y0, const KPtrs& kp, int64_t n) {
  switch (row) {
    case 0: k_stage_combine<T, 0><<<grid, kThreads, 0, st>>>(ctl, out, y0, kp, n); break;
    case 1: k_stage_combine<T, 1><<<grid, kThreads, 0, st>>>(ctl, out, y0, kp, n); break;
    case 2: k_stage_combine<T, 2><<<grid, kThreads, 0, st>>>(ctl, out, y0, kp, n); break;
    case 3: k_stage_combine<T, 3><<<grid, kThreads, 0, st>>>(ctl, out, y0, kp, n); break;
    case 4: k_stage_combine<T, 4><<<grid, kThreads, 0, st>>>(ctl, out, y0, kp, n); break;
    case 5: k_stage_combine<T, 5><<<grid, kThreads, 0, st>>>(ctl, out, y0, kp, n); break;
    case 6: k_stage_combine<T, 6><<<grid, kThreads, 0, st>>>(ctl, out, y0, kp, n); break;
    default: k_stage_combine<T, 7><<<grid, kThreads, 0, st>>>(ctl, out, y0, kp, n); break;
  }
}

// ---- K3: error estimate + per-member sum of squared ratios -----------------------------------
template <typename T>
__global__ void __launch_bounds__(kNormThreads) k_error_norm(const node_ctl_t* __restrict__ ctl, const T* __restrict__ y0,
                                                         const T* __restrict__ y1, KPtrs ks, Segs segs,
                                                         double* __restrict__ partials, int* __restrict__ nonfinite) {
  using A = Arith<T>;
  constexpr int V = Vec<T>::n;
  __shared__ double scratch[32];
  const int seg = blockIdx.y;
  const T h = ctl_h<T>(ctl);
  const T rtol = (T)ctl->rtol[seg], atol = (T)ctl->atol[seg];
  T hc[6];
  const int src[6] = {0, 2, 3, 4, 5, 6};
#pragma unroll
  for (int j = 0; j < 6; ++j) hc[j] = A::mul(h, (T)kCErr(src[j]));
  const int64_t base = segs.off[seg], len = segs.len[seg];
  const int64_t nvec = len / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  bool bad = false;
  auto one = [&](T a, T b, T e) {
    bad |= !isfinite(a);  // inf or nan in |y0| (dopri5.py:102)
    const T tol = A::add(atol, A::mul(rtol, A::max(A::abs(a), A::abs(b))));
    const T q = A::div(e, tol);
    acc += (double)A::mul(q, q);
  };
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    T e[V], kv[V], a[V], b[V];
#pragma unroll
    for (int u = 0; u < V; ++u) e[u] = (T)0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      vload<T>(reinterpret_cast<const T*>(ks.p[src[j]]) + base + i * V, kv);
#pragma unroll
      for (int u = 0; u < V; ++u) e[u] = A::add(e[u], A::mul(hc[j], kv[u]));
    }
    vload<T>(y0 + base + i * V, a);
    vload<T>(y1 + base + i * V, b);
#pragma unroll
    for (int u = 0; u < V; ++u) one(a[u], b[u], e[u]);
  }
  for (int64_t i = nvec * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    T e = (T)0;
#pragma unroll
    for (int j = 0; j < 6; ++j) e = A::add(e, A::mul(hc[j], reinterpret_cast<const T*>(ks.p[src[j]])[base + i]));
    one(y0[base + i], y1[base + i], e);
  }
  if (bad) atomicOr(nonfinite, 1);
  const double s = block_sum(acc, scratch);
  if (threadIdx.x == 0) {
    partials[(int64_t)(seg * 2 + 0) * kPartialBlocks + blockIdx.x] = s;
    partials[(int64_t)(seg * 2 + 1) * kPartialBlocks + blockIdx.x] = 0.0;
  }
}

// ---- K5: initial-step norms -------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) k_init_norms(const node_ctl_t* __restrict__ ctl, int mode, const T* __restrict__ y0,
                                                         const T* __restrict__ f0, const T* __restrict__ f1, Segs segs,
                                                         double* __restrict__ partials) {
  using A = Arith<T>;
  __shared__ double scratch[32];
  const int seg = blockIdx.y;
  // misc.py:121-123 uses rtol[0] / atol[0] for every member (dopri5.py:80 passes only those)
  const T rtol = (T)ctl->rtol[0], atol = (T)ctl->atol[0];
  const int64_t base = segs.off[seg], len = segs.len[seg];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += stride) {
    const T y = y0[base + i];
    const T scale = A::add(atol, A::mul(A::abs(y), rtol));
    if (mode == 0) {
      const T a = A::div(y, scale), b = A::div(f0[base + i], scale);
      s0 += (double)A::mul(a, a);
      s1 += (double)A::mul(b, b);
    } else {
      const T a = A::div(A::sub(f1[base + i], f0[base + i]), scale);
      s0 += (double)A::mul(a, a);
    }
  }
  const double r0 = block_sum(s0, scratch);
  const double r1 = block_sum(s1, scratch);
  if (threadIdx.x == 0) {
    partials[(int64_t)(seg * 2 + 0) * kPartialBlocks + blockIdx.x] = r0;
    partials[(int64_t)(seg * 2 + 1) * kPartialBlocks + blockIdx.x] = r1;
  }
}

// ---- fixed-order fold of the per-block partials ----------------------------------------------
__global__ void k_reduce_partials(const double* __restrict__ partials, int n_rows, double* __restrict__ sums) {
  const int row = blockIdx.x;
  if (row >= n_rows) return;
  double v = 0.0;
  for (int b = threadIdx.x; b < kPartialBlocks; b += 32) v += partials[(int64_t)row * kPartialBlocks + b];
  v = warp_sum(v);
  if (threadIdx.x == 0) sums[row] = v;
}

// ---- K6: controller ---------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T tsqrt(T x);
template <> __device__ __forceinline__ float tsqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double tsqrt<double>(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ T tpow(T x, T y);
template <> __device__ __forceinline__ float tpow<float>(float x, float y) { return powf(x, y); }
template <> __device__ __forceinline__ double tpow<double>(double x, double y) { return pow(x, y); }

// Rounds the next attempt's times to the state dtype (rk_common.py:45-49) or finishes the solve.
template <typename T>
__device__ void prepare_attempt(node_ctl_t* c) {
  using A = Arith<T>;
  if (c->next_out >= c->n_out) { c->done = 1; return; }
  if (c->steps_this_advance >= c->max_num_steps) { c->status |= NODE_ST_MAX_STEPS; c->done = 1; return; }
  c->t_attempt = c->t1;
  c->dt_attempt = c->dt;
  if (!(c->t_attempt + c->dt_attempt > c->t_attempt)) { c->status |= NODE_ST_DT_UNDERFLOW; c->done = 1; return; }
  const T s = (T)c->t_attempt, h = (T)c->dt_attempt;
  T ts[7];
  ts[0] = s;
  for (int i = 0; i < 6; ++i) ts[i + 1] = A::add(s, A::mul((T)kAlpha(i), h));
  for (int i = 0; i < 7; ++i) { c->ts64[i] = (double)ts[i]; c->ts32[i] = (float)ts[i]; }
  c->h64 = (double)h;
  c->h32 = (float)h;
}

template <typename T>
__device__ void controller_impl(node_ctl_t* c, int mode, const double* sums, const int* nonfinite, const double* t_out) {
  using A = Arith<T>;
  const int ns = c->n_seg;
  if (c->done) { c->out_lo = c->out_hi = c->next_out; return; }
  if (mode == 0) {
    // misc.py:123-131: d0, d1 per member; h0 = 0.01 * max(d0/d1) unless either norm is tiny
    T d0max = (T)0, d1max = (T)0, qmax = (T)0;
    for (int s = 0; s < ns; ++s) {
      const T rootn = (T)sqrt((double)c->seg_numel[s]);
      const T d0 = A::div((T)sqrt(sums[s * 2 + 0]), rootn);
      const T d1 = A::div((T)sqrt(sums[s * 2 + 1]), rootn);
      const T q = A::div(d0, d1);
      if (s == 0 || d0 > d0max) d0max = d0;
      if (s == 0 || d1 > d1max) d1max = d1;
      if (s == 0 || q > qmax) qmax = q;
    }
    T h0;
    if ((double)d0max < 1e-5 || (double)d1max < 1e-5) h0 = (T)1e-6;
    else h0 = A::mul((T)0.01, qmax);
    c->h0 = (double)h0;
    c->h0_32 = (float)h0;
    c->d1max = (double)d1max;
    // probe time t0 + h0 in the state dtype (misc.py:134)
    const T t0 = (T)t_out[0];
    c->ts64[0] = (double)t0; c->ts32[0] = (float)t0;
    c->ts64[1] = (double)A::add(t0, h0); c->ts32[1] = (float)A::add(t0, h0);
    c->nfe += 2;  // f0 (dopri5.py:78) and the probe (misc.py:134)
    return;
  }
  if (mode == 1) {
    // misc.py:136-143
    const T h0 = (T)c->h0;
    T d2max = (T)0;
    for (int s = 0; s < ns; ++s) {
      const T rootn = (T)sqrt((double)c->seg_numel[s]);
      const T d2 = A::div(A::div((T)sqrt(sums[s * 2 + 0]), rootn), h0);
      if (s == 0 || d2 > d2max) d2max = d2;
    }
    const T d1max = (T)c->d1max;
    T h1;
    if ((double)d1max <= 1e-15 && (double)d2max <= 1e-15) {
      const T a = (T)1e-6, b = A::mul(h0, (T)1e-3);
      h1 = a > b ? a : b;
    } else {
      const T m = d2max > d1max ? d2max : d1max;
      h1 = tpow<T>(A::div((T)0.01, m), (T)(1.0 / 5.0));
    }
    const T h100 = A::mul((T)100, h0);
    const T first = h100 < h1 ? h100 : h1;
    c->dt = (double)first;
    c->t0 = c->t1 = t_out[0];
    c->next_out = 1;
    c->out_lo = c->out_hi = 1;
    c->steps_this_advance = 0;
    prepare_attempt<T>(c);
    return;
  }
  if (mode == 3) {
    // options['first_step'] given (dopri5.py:81-82): the first step is the constant 0.01 whatever the value, and the
    // initial-step probe is never evaluated; sums[0] carries the constant as the reference rounds it.
    c->dt = sums[0];
    c->t0 = c->t1 = t_out[0];
    c->next_out = 1;
    c->out_lo = c->out_hi = 1;
    c->steps_this_advance = 0;
    c->nfe += 1;  // f0 only (dopri5.py:78)
    prepare_attempt<T>(c);
    return;
  }
  // mode 2: dopri5.py:109-121
  if (nonfinite != nullptr && *nonfinite) { c->status |= NODE_ST_NONFINITE; c->done = 1; c->out_lo = c->out_hi = c->next_out; return; }
  bool accept = true;
  T r = (T)0;
  for (int s = 0; s < ns; ++s) {
    const T m = (T)(sums[s * 2 + 0] / (double)c->seg_numel[s]);
    c->ratio[s] = (double)m;
    if (!(m <= (T)1)) accept = false;
    if (s == 0 || m > r) r = m;
  }
  const double dt = c->dt_attempt;
  double dt_next;
  if (r == (T)0) {
    dt_next = dt * c->ifactor;
  } else {
    const double dfac = (r < (T)1) ? 1.0 : c->dfactor;
    const double e = (double)tsqrt<T>(r);
    const double lo = 1.0 / c->ifactor, hi = 1.0 / dfac;
    double f = pow(e, c->expo) / c->safety;
    f = (f != f) ? f : fmax(lo, fmin(f, hi));
    dt_next = dt / f;
  }
  const int a = c->n_attempt;
  if (a < NODE_MAX_TRACE) { c->tr_t[a] = c->t_attempt; c->tr_dt[a] = dt; c->tr_ratio[a] = (double)r; c->tr_acc[a] = accept ? 1 : 0; }
  c->n_attempt = a + 1;
  c->nfe += 6;
  c->steps_this_advance += 1;
  c->accepted_last = accept ? 1 : 0;
  c->out_lo = c->out_hi = c->next_out;
  if (accept) {
    c->it_t0 = c->t_attempt;
    c->it_t1 = c->t_attempt + dt;
    c->it_h64 = c->h64;
    c->it_h32 = c->h32;
    c->it_cur = c->cur;
    c->t0 = c->it_t0;
    c->t1 = c->it_t1;
    c->cur ^= 1;
    c->n_accept += 1;
    int nx = c->next_out;
    while (nx < c->n_out && !(t_out[nx] > c->t1)) {
      const T a0 = (T)c->it_t0, a1 = (T)c->it_t1, tt = (T)t_out[nx];
      if (!((a0 <= tt) && (tt <= a1))) c->status |= NODE_ST_INTERP_RANGE;
      ++nx;
    }
    c->out_hi = nx;
    if (nx > c->next_out) c->steps_this_advance = 0;
    c->next_out = nx;
    if (c->status & NODE_ST_INTERP_RANGE) { c->done = 1; }
  } else {
    c->t0 = c->t_attempt;
    c->n_reject += 1;
  }
  c->dt = dt_next;
  if (!c->done) prepare_attempt<T>(c);
}

__global__ void k_controller(node_ctl_t* c, int mode, const double* sums, const int* nonfinite, const double* t_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (c->dtype == NODE_F64) controller_impl<double>(c, mode, sums, nonfinite, t_out);
  else controller_impl<float>(c, mode, sums, nonfinite, t_out);
}

// ---- K4: dense output ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) k_interp(const node_ctl_t* __restrict__ ctl, const double* __restrict__ t_out,
                                                     T* __restrict__ out, int64_t out_stride, const T* ya,
                                                     const T* yb, const T* __restrict__ ym, const T* fa,
                                                     const T* fb, int64_t n, int use_ctl_cur) {
  using A = Arith<T>;
  const int lo = ctl->out_lo, hi = ctl->out_hi;
  if (hi <= lo) return;
  // fused route: (ya,yb)/(fa,fb) are the ping-pong buffers and ctl->it_cur names the one that
  // held the accepted step's start; generic route: ya=y0, yb=y1 as given.
  const bool swap = use_ctl_cur && ctl->it_cur == 1;
  const T* __restrict__ y0 = swap ? yb : ya;
  const T* __restrict__ y1 = swap ? ya : yb;
  const T* __restrict__ f0 = swap ? fb : fa;
  const T* __restrict__ f1 = swap ? fa : fb;
  const T dt = sizeof(T) == 4 ? (T)ctl->it_h32 : (T)ctl->it_h64;
  const T t0 = (T)ctl->it_t0, t1 = (T)ctl->it_t1;
  const T m2dt = A::mul((T)-2, dt), p2dt = A::mul((T)2, dt), p5dt = A::mul((T)5, dt), m3dt = A::mul((T)-3, dt),
          m4dt = A::mul((T)-4, dt);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const T F0 = f0[i], F1 = f1[i], Y0 = y0[i], Y1 = y1[i], YM = ym[i];
    auto dot5 = [&](T c0, T c1, T c2, T c3, T c4) {
      T acc = A::add((T)0, A::mul(c0, F0));
      acc = A::add(acc, A::mul(c1, F1));
      acc = A::add(acc, A::mul(c2, Y0));
      acc = A::add(acc, A::mul(c3, Y1));
      return A::add(acc, A::mul(c4, YM));
    };
    const T ca = dot5(m2dt, p2dt, (T)-8, (T)-8, (T)16);
    const T cb = dot5(p5dt, m3dt, (T)18, (T)14, (T)-32);
    const T cc = dot5(m4dt, dt, (T)-11, (T)-5, (T)16);
    const T cd = A::mul(dt, F0);
    for (int idx = lo; idx < hi; ++idx) {
      const T x = A::div(A::sub((T)t_out[idx], t0), A::sub(t1, t0));
      const T x2 = A::mul(x, x), x3 = A::mul(x2, x), x4 = A::mul(x3, x);
      T acc = A::add((T)0, A::mul(ca, x4));
      acc = A::add(acc, A::mul(cb, x3));
      acc = A::add(acc, A::mul(cc, x2));
      acc = A::add(acc, A::mul(cd, x));
      acc = A::add(acc, A::mul(Y0, (T)1));
      out[(int64_t)idx * out_stride + i] = acc;
    }
  }
}

// fp32, 128-bit vectorised variant of k_interp (same arithmetic, same order): 4 elements per thread so that 5 x 16 bytes
// are in flight per thread, x and its powers hoisted out of the element loop. Used when numel % 4 == 0 and all
// pointers / strides are 16-byte aligned (always the case on the fused route).
__global__ void __launch_bounds__(kThreads) k_interp_v4(const node_ctl_t* __restrict__ ctl, const double* __restrict__ t_out,
                                                        float* __restrict__ out, int64_t out_stride, const float* ya,
                                                        const float* yb, const float* __restrict__ ym, const float* fa,
                                                        const float* fb, int64_t n4, int use_ctl_cur) {
  using A = Arith<float>;
  const int lo = ctl->out_lo, hi = ctl->out_hi;
  if (hi <= lo) return;
  const bool swap = use_ctl_cur && ctl->it_cur == 1;
  const float4* __restrict__ y0 = reinterpret_cast<const float4*>(swap ? yb : ya);
  const float4* __restrict__ y1 = reinterpret_cast<const float4*>(swap ? ya : yb);
  const float4* __restrict__ f0 = reinterpret_cast<const float4*>(swap ? fb : fa);
  const float4* __restrict__ f1 = reinterpret_cast<const float4*>(swap ? fa : fb);
  const float4* __restrict__ ym4 = reinterpret_cast<const float4*>(ym);
  const float dt = (float)ctl->it_h32;
  const float t0 = (float)ctl->it_t0, t1 = (float)ctl->it_t1;
  const float m2dt = A::mul(-2.f, dt), p2dt = A::mul(2.f, dt), p5dt = A::mul(5.f, dt), m3dt = A::mul(-3.f, dt), m4dt = A::mul(-4.f, dt);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 F0 = f0[i], F1 = f1[i], Y0 = y0[i], Y1 = y1[i], YM = ym4[i];
    const float vF0[4] = {F0.x, F0.y, F0.z, F0.w}, vF1[4] = {F1.x, F1.y, F1.z, F1.w}, vY0[4] = {Y0.x, Y0.y, Y0.z, Y0.w},
                vY1[4] = {Y1.x, Y1.y, Y1.z, Y1.w}, vYM[4] = {YM.x, YM.y, YM.z, YM.w};
    float ca[4], cb[4], cc[4], cd[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      auto dot5 = [&](float c0, float c1, float c2, float c3, float c4) {
        float acc = A::add(0.f, A::mul(c0, vF0[e]));
        acc = A::add(acc, A::mul(c1, vF1[e]));
        acc = A::add(acc, A::mul(c2, vY0[e]));
        acc = A::add(acc, A::mul(c3, vY1[e]));
        return A::add(acc, A::mul(c4, vYM[e]));
      };
      ca[e] = dot5(m2dt, p2dt, -8.f, -8.f, 16.f);
      cb[e] = dot5(p5dt, m3dt, 18.f, 14.f, -32.f);
      cc[e] = dot5(m4dt, dt, -11.f, -5.f, 16.f);
      cd[e] = A::mul(dt, vF0[e]);
    }
    for (int idx = lo; idx < hi; ++idx) {
      const float x = A::div(A::sub((float)t_out[idx], t0), A::sub(t1, t0));
      const float x2 = A::mul(x, x), x3 = A::mul(x2, x), x4 = A::mul(x3, x);
      float r[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float acc = A::add(0.f, A::mul(ca[e], x4));
        acc = A::add(acc, A::mul(cb[e], x3));
        acc = A::add(acc, A::mul(cc[e], x2));
        acc = A::add(acc, A::mul(cd[e], x));
        r[e] = A::add(acc, A::mul(vY0[e], 1.f));
      }
      reinterpret_cast<float4*>(out + (int64_t)idx * out_stride)[i] = make_float4(r[0], r[1], r[2], r[3]);
    }
  }
}

// Per-member tolerances and element counts travel as kernel arguments (by value): nothing is read from host memory
// after the call returns, so the launch sequence of a solve can be captured in a CUDA graph and replayed.
struct CtlSegs { double rtol[NODE_MAX_SEG], atol[NODE_MAX_SEG]; int64_t numel[NODE_MAX_SEG]; };

__global__ void k_ctl_init(node_ctl_t* c, int dtype, int n_seg, double safety, double ifactor, double dfactor, double expo,
                           int max_num_steps, int n_out, int tsign, CtlSegs segs) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  for (int i = 0; i < NODE_MAX_SEG; ++i) { c->rtol[i] = segs.rtol[i]; c->atol[i] = segs.atol[i]; c->seg_numel[i] = segs.numel[i]; }
  c->dtype = dtype; c->n_seg = n_seg; c->safety = safety; c->ifactor = ifactor; c->dfactor = dfactor; c->expo = expo;
  c->max_num_steps = max_num_steps; c->n_out = n_out; c->tsign = tsign; c->next_out = 1;
}

// elementwise kernels without per-block partials: up to 8 resident blocks per SM
static int grid_wide(int64_t n_items) {
  int64_t b = (n_items + kThreads - 1) / kThreads;
  if (b < 1) b = 1;
  if (b > 148 * 8) b = 148 * 8;
  return (int)b;
}

static int grid_for(int64_t n_items) {
  int64_t b = (n_items + kThreads - 1) / kThreads;
  if (b < 1) b = 1;
  if (b > kPartialBlocks) b = kPartialBlocks;
  return (int)b;
}

}  // namespace node

using namespace node;

extern "C" int node_b200_abi_version(void) { return NODE_B200_ABI_VERSION; }

extern "C" int node_b200_ctl_layout(int64_t* o, int cap) {
  const int64_t v[] = {
      (int64_t)sizeof(node_ctl_t), kPartialBlocks, NODE_MAX_SEG, NODE_MAX_TRACE,
      offsetof(node_ctl_t, t0), offsetof(node_ctl_t, t1), offsetof(node_ctl_t, dt), offsetof(node_ctl_t, ratio),
      offsetof(node_ctl_t, ts64), offsetof(node_ctl_t, ts32), offsetof(node_ctl_t, h64), offsetof(node_ctl_t, h32),
      offsetof(node_ctl_t, out_lo), offsetof(node_ctl_t, out_hi), offsetof(node_ctl_t, next_out),
      offsetof(node_ctl_t, n_attempt), offsetof(node_ctl_t, n_accept), offsetof(node_ctl_t, n_reject),
      offsetof(node_ctl_t, nfe), offsetof(node_ctl_t, status), offsetof(node_ctl_t, done), offsetof(node_ctl_t, cur),
      offsetof(node_ctl_t, accepted_last), offsetof(node_ctl_t, tr_t), offsetof(node_ctl_t, tr_dt),
      offsetof(node_ctl_t, tr_ratio), offsetof(node_ctl_t, tr_acc), offsetof(node_ctl_t, h0), offsetof(node_ctl_t, h0_32),
      offsetof(node_ctl_t, it_t0), offsetof(node_ctl_t, it_t1), offsetof(node_ctl_t, it_h64), offsetof(node_ctl_t, it_h32)};
  const int n = (int)(sizeof(v) / sizeof(v[0]));
  for (int i = 0; i < n && i < cap; ++i) o[i] = v[i];
  return n;
}

extern "C" int node_b200_ctl_init(node_ctl_t* ctl, int dtype, int n_seg, const double* rtol, const double* atol,
                                  const int64_t* seg_numel, double safety, double ifactor, double dfactor, double expo,
                                  int max_num_steps, int n_out, int tsign, void* stream) {
  if (n_seg < 1 || n_seg > NODE_MAX_SEG) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  NODE_CUDA_OK(cudaMemsetAsync(ctl, 0, sizeof(node_ctl_t), st));
  CtlSegs segs;
  for (int i = 0; i < NODE_MAX_SEG; ++i) {
    segs.rtol[i] = i < n_seg ? rtol[i] : 0.0; segs.atol[i] = i < n_seg ? atol[i] : 0.0; segs.numel[i] = i < n_seg ? seg_numel[i] : 0;
  }
  k_ctl_init<<<1, 32, 0, st>>>(ctl, dtype, n_seg, safety, ifactor, dfactor, expo, max_num_steps, n_out, tsign, segs);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_rk_stage_combine(const node_ctl_t* ctl, int dtype, int which, void* out, const void* y0,
                                          const void* const* ks, int n_k, int64_t numel, void* stream) {
  if (which < 0 || which > 7 || n_k < kRowLen(which)) return (int)cudaErrorInvalidValue;
  KPtrs kp{};
  for (int i = 0; i < 7; ++i) kp.p[i] = i < n_k ? ks[i] : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == NODE_F32) {
    launch_stage_combine<float>(which, grid_wide(numel / 8 + 1), st, ctl, (float*)out, (const float*)y0, kp, numel);
  } else {
    launch_stage_combine<double>(which, grid_wide(numel / 4 + 1), st, ctl, (double*)out, (const double*)y0, kp, numel);
  }
  return (int)cudaGetLastError();
}

static Segs make_segs(const int64_t* off, const int64_t* len, int n) {
  Segs s{};
  for (int i = 0; i < n; ++i) { s.off[i] = off[i]; s.len[i] = len[i]; }
  return s;
}

extern "C" int node_b200_rk_error_norm(const node_ctl_t* ctl, int dtype, const void* y0, const void* y1,
                                       const void* const* ks, const int64_t* seg_off, const int64_t* seg_len, int n_seg,
                                       double* partials, int* nonfinite_flag, void* stream) {
  if (n_seg < 1 || n_seg > NODE_MAX_SEG) return (int)cudaErrorInvalidValue;
  KPtrs kp{};
  for (int i = 0; i < 7; ++i) kp.p[i] = ks[i];
  cudaStream_t st = (cudaStream_t)stream;
  NODE_CUDA_OK(cudaMemsetAsync(partials, 0, sizeof(double) * 2 * NODE_MAX_SEG * kPartialBlocks, st));
  NODE_CUDA_OK(cudaMemsetAsync(nonfinite_flag, 0, sizeof(int), st));
  int64_t mx = 1;
  for (int i = 0; i < n_seg; ++i) mx = seg_len[i] > mx ? seg_len[i] : mx;
  const Segs segs = make_segs(seg_off, seg_len, n_seg);
  if (dtype == NODE_F32) {
    dim3 g(grid_for(mx / 4 + 1), n_seg);
    k_error_norm<float><<<g, kNormThreads, 0, st>>>(ctl, (const float*)y0, (const float*)y1, kp, segs, partials, nonfinite_flag);
  } else {
    dim3 g(grid_for(mx / 2 + 1), n_seg);
    k_error_norm<double><<<g, kNormThreads, 0, st>>>(ctl, (const double*)y0, (const double*)y1, kp, segs, partials, nonfinite_flag);
  }
  return (int)cudaGetLastError();
}

extern "C" int node_b200_init_norms(const node_ctl_t* ctl, int dtype, int mode, const void* y0, const void* f0, const void* f1,
                                    const int64_t* seg_off, const int64_t* seg_len, int n_seg, double* partials, void* stream) {
  if (n_seg < 1 || n_seg > NODE_MAX_SEG) return (int)cudaErrorInvalidValue;
  cudaStream_t st = (cudaStream_t)stream;
  NODE_CUDA_OK(cudaMemsetAsync(partials, 0, sizeof(double) * 2 * NODE_MAX_SEG * kPartialBlocks, st));
  int64_t mx = 1;
  for (int i = 0; i < n_seg; ++i) mx = seg_len[i] > mx ? seg_len[i] : mx;
  const Segs segs = make_segs(seg_off, seg_len, n_seg);
  dim3 g(grid_for(mx), n_seg);
  if (dtype == NODE_F32) k_init_norms<float><<<g, kThreads, 0, st>>>(ctl, mode, (const float*)y0, (const float*)f0, (const float*)f1, segs, partials);
  else k_init_norms<double><<<g, kThreads, 0, st>>>(ctl, mode, (const double*)y0, (const double*)f0, (const double*)f1, segs, partials);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_reduce_partials(const double* partials, int n_rows, double* sums, void* stream) {
  k_reduce_partials<<<n_rows, 32, 0, (cudaStream_t)stream>>>(partials, n_rows, sums);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_controller(node_ctl_t* ctl, int mode, const double* sums, const int* nonfinite_flag,
                                    const double* t_out, void* stream) {
  k_controller<<<1, 32, 0, (cudaStream_t)stream>>>(ctl, mode, sums, nonfinite_flag, t_out);
  return (int)cudaGetLastError();
}

extern "C" int node_b200_interp_eval(const node_ctl_t* ctl, int dtype, const double* t_out, void* out, int64_t out_stride,
                                     const void* y0, const void* y1, const void* ymid, const void* f0, const void* f1,
                                     int64_t numel, int use_ctl_cur, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(numel);
  auto al16 = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
  if (dtype == NODE_F32 && numel % 4 == 0 && out_stride % 4 == 0 && al16(out) && al16(y0) && al16(y1) && al16(ymid) && al16(f0) && al16(f1)) {
    int64_t b = (numel / 4 + kThreads - 1) / kThreads;
    if (b > 148 * 8) b = 148 * 8;
    k_interp_v4<<<(int)b, kThreads, 0, st>>>(ctl, t_out, (float*)out, out_stride, (const float*)y0, (const float*)y1,
                                             (const float*)ymid, (const float*)f0, (const float*)f1, numel / 4, use_ctl_cur);
    return (int)cudaGetLastError();
  }
  if (dtype == NODE_F32)
    k_interp<float><<<g, kThreads, 0, st>>>(ctl, t_out, (float*)out, out_stride, (const float*)y0, (const float*)y1,
                                            (const float*)ymid, (const float*)f0, (const float*)f1, numel, use_ctl_cur);
  else
    k_interp<double><<<g, kThreads, 0, st>>>(ctl, t_out, (double*)out, out_stride, (const double*)y0, (const double*)y1,
                                             (const double*)ymid, (const double*)f0, (const double*)f1, numel, use_ctl_cur);
  return (int)cudaGetLastError();
}
