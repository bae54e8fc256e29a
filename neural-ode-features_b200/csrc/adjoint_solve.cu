// node_b200_adjoint_solve - one interval of odeint_adjoint's backward integration (adjoint.py:77-97: odeint of the augmented
// system from t[i] to t[i-1]; dopri5.py:77-122 underneath) as ONE call with NO host read inside: f0, the initial-step probe
// and its two norms are enqueued directly, and the adaptive loop `while next_t > t1: step` (dopri5.py:88) is a CUDA graph
// WHILE node whose condition the last kernel of the body sets from the device-resident controller block
// (cudaGraphSetConditional) - the accept / reject decision, the step size and the end of the integration never leave the GPU.
//
// The body is the kernel sequence of node_b200_adjoint_step (csrc/odefunc_vjp.cu) with the ping-pong of the (y, f) pair
// removed: an attempt always goes from rows (Y0, F0) to rows (Y1, F1), and a commit kernel copies the new pair over the old one
// when the controller accepted - the pointers of the captured launches are therefore the same for every iteration.
#include <cstring>
#include <mutex>
#include <vector>
#include "node_common.cuh"

extern "C" int node_b200_odefunc_vjp(void* workspace, void* vjp_workspace, const float* y, const float* adj_y, const float* t_dev,
                                     float tsign, float* f_out, float* vjp_y, float* vjp_t, float* vjp_params, int N, int C,
                                     int H, int W, void* stream);
extern "C" int node_b200_odefunc_vjp_split(void* workspace, void* vjp_workspace, const float* y, const float* adj_y, const float* t_dev,
                                           float tsign, float* f_out, float* vjp_y, float* vjp_t, float* vjp_params, int N, int C, int H,
                                           int W, void* stream, void* side_stream, void* fork_event);

namespace node {

__global__ void k_adjoint_t0(node_ctl_t* c, const double* t_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  c->ts64[0] = (double)(float)t_out[0];
  c->ts32[0] = (float)t_out[0];
}

// rows (Y1, F1) -> (Y0, F0) when the last attempt was accepted (dopri5.py:114-116: y0 <- y1, f0 <- f1)
__global__ void __launch_bounds__(256) k_adjoint_commit(const node_ctl_t* __restrict__ c, float4* __restrict__ y0,
                                                        const float4* __restrict__ y1, float4* __restrict__ f0,
                                                        const float4* __restrict__ f1, int64_t nvec) {
  if (!c->accepted_last) return;
  const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * nvec; i += nthr) {
    if (i < nvec) y0[i] = y1[i];
    else f0[i - nvec] = f1[i - nvec];
  }
}

__global__ void k_adjoint_continue(cudaGraphConditionalHandle h, const node_ctl_t* c) {
  if (threadIdx.x == 0 && blockIdx.x == 0) cudaGraphSetConditional(h, c->done ? 0u : 1u);
}

struct AdjointSolveKey {
  void *ctl, *bufs, *workspace, *vjp_workspace, *vjp_workspace2, *partials, *sums, *flag, *t_out, *out;
  int64_t row_elems, ts32_off, seg_off[4], seg_len[4];
  float tsign;
  int N, C, H, W, device;
};

struct AdjointSolveGraph { AdjointSolveKey key; cudaGraph_t graph; cudaGraphExec_t exec; };

static std::mutex g_adjoint_mutex;
static std::vector<AdjointSolveGraph> g_adjoint_graphs;
constexpr size_t kMaxAdjointGraphs = 16;

}  // namespace node

using namespace node;

// Side stream + events of the overlapped body (per device, created once; inside the capture they become graph edges)
struct AdjointSide { cudaStream_t stream; cudaEvent_t fork[6], done[6], join; bool ok; };
static AdjointSide g_adjoint_side[64];

static AdjointSide* adjoint_side(int device) {
  if (device < 0 || device >= 64) return nullptr;
  AdjointSide& s = g_adjoint_side[device];
  if (!s.ok) {
    if (cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    for (int i = 0; i < 6; ++i)
      if (cudaEventCreateWithFlags(&s.fork[i], cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&s.done[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    s.ok = true;
  }
  return &s;
}

// the body of the loop: one attempted step, rows (Y0, F0) -> (Y1, F1), controller, dense output, commit
static int adjoint_enqueue_attempt(const AdjointSolveKey& k, cudaStream_t st) {
  node_ctl_t* ctl = (node_ctl_t*)k.ctl;
  float* bufs = (float*)k.bufs;
  enum { Y0 = 0, Y1 = 1, F0 = 2, F1 = 3, K2 = 4, YMID = 9, YI = 10 };
  auto row = [&](int i) { return bufs + (size_t)i * k.row_elems; };
  const int ks[7] = {F0, K2, K2 + 1, K2 + 2, K2 + 3, K2 + 4, F1};
  const void* kp[7];
  for (int i = 0; i < 7; ++i) kp[i] = row(ks[i]);
  const float* ts32 = (const float*)((const char*)k.ctl + k.ts32_off);
  AdjointSide* side = k.vjp_workspace2 != nullptr ? adjoint_side(k.device) : nullptr;
  if (side != nullptr) {
    // Small batches (a dependent chain of ~100 us evaluations): the weight-gradient GEMM and the fold into (vjp_t, vjp_params) of
    // stage i run on a side branch under stage i + 1's k_vjp - nothing on the main branch reads the parameter / time adjoints
    // (adjoint.py:35) until the error norm. Stage combinations on the main branch cover the (y, adj_y) members only; the
    // (adj_t, adj_params) members are combined once, for y1, on the side branch. Two VJP workspaces alternate, so stage i + 1's
    // k_vjp does not overwrite the operands stage i's GEMM is reading.
    const int64_t n_state = k.seg_off[2], n_par = k.row_elems - k.seg_off[2];
    const void* kpar[7];
    for (int i = 0; i < 7; ++i) kpar[i] = row(ks[i]) + n_state;
    void* vws[2] = {k.vjp_workspace, k.vjp_workspace2};
    for (int i = 0; i < 6; ++i) {
      float* dst = row(i < 5 ? YI : Y1);
      NODE_CUDA_OK((cudaError_t)node_b200_rk_stage_combine(ctl, NODE_F32, i, dst, row(Y0), kp, i + 1, n_state, st));
      if (i >= 2) NODE_CUDA_OK(cudaStreamWaitEvent(st, side->done[i - 2], 0));      // this workspace's previous GEMM has read its operands
      float* o = row(ks[i + 1]);
      NODE_CUDA_OK((cudaError_t)node_b200_odefunc_vjp_split(k.workspace, vws[i & 1], dst + k.seg_off[0], dst + k.seg_off[1], ts32 + (i + 1),
                                                            k.tsign, o + k.seg_off[0], o + k.seg_off[1], o + k.seg_off[2], o + k.seg_off[3],
                                                            k.N, k.C, k.H, k.W, st, side->stream, side->fork[i]));
      NODE_CUDA_OK(cudaEventRecord(side->done[i], side->stream));
    }
    NODE_CUDA_OK((cudaError_t)node_b200_rk_stage_combine(ctl, NODE_F32, 5, row(Y1) + n_state, row(Y0) + n_state, kpar, 6, n_par, side->stream));
    NODE_CUDA_OK(cudaEventRecord(side->join, side->stream));
    NODE_CUDA_OK(cudaStreamWaitEvent(st, side->join, 0));
  } else {
    for (int i = 0; i < 6; ++i) {
      float* dst = row(i < 5 ? YI : Y1);
      NODE_CUDA_OK((cudaError_t)node_b200_rk_stage_combine(ctl, NODE_F32, i, dst, row(Y0), kp, i + 1, k.row_elems, st));
      float* o = row(ks[i + 1]);
      NODE_CUDA_OK((cudaError_t)node_b200_odefunc_vjp(k.workspace, k.vjp_workspace, dst + k.seg_off[0], dst + k.seg_off[1], ts32 + (i + 1),
                                                      k.tsign, o + k.seg_off[0], o + k.seg_off[1], o + k.seg_off[2], o + k.seg_off[3],
                                                      k.N, k.C, k.H, k.W, st));
    }
  }
  NODE_CUDA_OK((cudaError_t)node_b200_rk_error_norm(ctl, NODE_F32, row(Y0), row(Y1), kp, k.seg_off, k.seg_len, 4, (double*)k.partials,
                                                    (int*)k.flag, st));
  NODE_CUDA_OK((cudaError_t)node_b200_reduce_partials((const double*)k.partials, 8, (double*)k.sums, st));
  NODE_CUDA_OK((cudaError_t)node_b200_controller(ctl, 2, (const double*)k.sums, (const int*)k.flag, (const double*)k.t_out, st));
  // dense output of the accepted step (dopri5.py:39-45, 92): both kernels return at once unless the controller scheduled outputs
  NODE_CUDA_OK((cudaError_t)node_b200_rk_stage_combine(ctl, NODE_F32, 6, row(YMID), row(Y0), kp, 7, k.row_elems, st));
  NODE_CUDA_OK((cudaError_t)node_b200_interp_eval(ctl, NODE_F32, (const double*)k.t_out, k.out, k.row_elems, row(Y0), row(Y1), row(YMID),
                                                  row(F0), row(F1), k.row_elems, 0, st));
  const int64_t nvec = k.row_elems / 4;
  int64_t blocks = (2 * nvec + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_adjoint_commit<<<(int)blocks, 256, 0, st>>>(ctl, (float4*)row(Y0), (const float4*)row(Y1), (float4*)row(F0), (const float4*)row(F1), nvec);
  return (int)cudaGetLastError();
}

static int adjoint_build_graph(const AdjointSolveKey& k, AdjointSolveGraph* out) {
  cudaGraph_t g = nullptr;
  NODE_CUDA_OK(cudaGraphCreate(&g, 0));
  cudaGraphConditionalHandle handle;
  cudaError_t e = cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return (int)e; }
  cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
  p.conditional.handle = handle;
  p.conditional.type = cudaGraphCondTypeWhile;
  p.conditional.size = 1;
  cudaGraphNode_t node;
  e = cudaGraphAddNode(&node, g, nullptr, 0, &p);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return (int)e; }
  cudaGraph_t body = p.conditional.phGraph_out[0];
  cudaStream_t cs = nullptr;
  e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking);
  if (e != cudaSuccess) { cudaGraphDestroy(g); return (int)e; }
  e = cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
  int rc = (int)e;
  if (e == cudaSuccess) {
    rc = adjoint_enqueue_attempt(k, cs);
    if (rc == 0) {
      k_adjoint_continue<<<1, 32, 0, cs>>>(handle, (const node_ctl_t*)k.ctl);
      rc = (int)cudaGetLastError();
    }
    cudaGraph_t captured = nullptr;
    e = cudaStreamEndCapture(cs, &captured);
    if (rc == 0) rc = (int)e;
  }
  cudaStreamDestroy(cs);
  cudaGraphExec_t exec = nullptr;
  if (rc == 0) rc = (int)cudaGraphInstantiate(&exec, g, 0);
  if (rc != 0) { cudaGraphDestroy(g); cudaGetLastError(); return rc; }
  out->key = k; out->graph = g; out->exec = exec;
  return 0;
}

extern "C" int node_b200_adjoint_solve(void* ctl_v, float* bufs, int64_t row_elems, const int64_t* seg_off, const int64_t* seg_len,
                                       int n_seg, void* workspace, void* vjp_workspace, void* vjp_workspace2, float tsign,
                                       int64_t ts32_offset_bytes, int N, int C, int H, int W, double* partials, double* sums,
                                       int* nonfinite_flag, const double* t_out, float* out, int first_step_given, void* stream) {
  if (n_seg != 4 || row_elems % 4 != 0) return (int)cudaErrorInvalidValue;
  AdjointSolveKey k;
  memset(&k, 0, sizeof(k));
  k.ctl = ctl_v; k.bufs = bufs; k.workspace = workspace; k.vjp_workspace = vjp_workspace; k.vjp_workspace2 = vjp_workspace2;
  k.partials = partials; k.sums = sums;
  k.flag = nonfinite_flag; k.t_out = (void*)t_out; k.out = out; k.row_elems = row_elems; k.ts32_off = ts32_offset_bytes;
  for (int i = 0; i < 4; ++i) { k.seg_off[i] = seg_off[i]; k.seg_len[i] = seg_len[i]; }
  k.tsign = tsign; k.N = N; k.C = C; k.H = H; k.W = W;
  NODE_CUDA_OK(cudaGetDevice(&k.device));
  cudaStream_t st = (cudaStream_t)stream;
  node_ctl_t* ctl = (node_ctl_t*)ctl_v;
  enum { Y0 = 0, F0 = 2, K2 = 4, YI = 10 };
  auto row = [&](int i) { return bufs + (size_t)i * row_elems; };
  const float* ts32 = (const float*)((const char*)ctl_v + ts32_offset_bytes);
  auto eval = [&](int src, int dst, int ti) {
    return node_b200_odefunc_vjp(workspace, vjp_workspace, row(src) + seg_off[0], row(src) + seg_off[1], ts32 + ti, tsign,
                                 row(dst) + seg_off[0], row(dst) + seg_off[1], row(dst) + seg_off[2], row(dst) + seg_off[3], N, C, H, W,
                                 stream);
  };
  // dopri5.py:77-83 before_integrate: f0, first step
  k_adjoint_t0<<<1, 32, 0, st>>>(ctl, t_out);
  NODE_CUDA_OK(cudaGetLastError());
  NODE_CUDA_OK((cudaError_t)eval(Y0, F0, 0));
  if (!first_step_given) {
    NODE_CUDA_OK((cudaError_t)node_b200_init_norms(ctl, NODE_F32, 0, row(Y0), row(F0), nullptr, seg_off, seg_len, 4, partials, stream));
    NODE_CUDA_OK((cudaError_t)node_b200_reduce_partials(partials, 8, sums, stream));
    NODE_CUDA_OK((cudaError_t)node_b200_controller(ctl, 0, sums, nullptr, t_out, stream));
    const void* kp0[1] = {row(F0)};
    NODE_CUDA_OK((cudaError_t)node_b200_rk_stage_combine(ctl, NODE_F32, 7, row(YI), row(Y0), kp0, 1, row_elems, stream));
    NODE_CUDA_OK((cudaError_t)eval(YI, K2, 1));                                 // misc.py:134
    NODE_CUDA_OK((cudaError_t)node_b200_init_norms(ctl, NODE_F32, 1, row(Y0), row(F0), row(K2), seg_off, seg_len, 4, partials, stream));
    NODE_CUDA_OK((cudaError_t)node_b200_reduce_partials(partials, 8, sums, stream));
    NODE_CUDA_OK((cudaError_t)node_b200_controller(ctl, 1, sums, nullptr, t_out, stream));
  } else {
    NODE_CUDA_OK((cudaError_t)node_b200_controller(ctl, 3, sums, nullptr, t_out, stream));   // sums[0] = 0.01 as the caller rounded it
  }
  NODE_CUDA_OK(cudaMemcpyAsync(out, row(Y0), (size_t)row_elems * 4, cudaMemcpyDeviceToDevice, st));   // solvers.py:26 solution[0] = y0
  // dopri5.py:88 `while next_t > t1: step` on the device
  cudaGraphExec_t exec = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_adjoint_mutex);
    for (auto& g : g_adjoint_graphs)
      if (memcmp(&g.key, &k, sizeof(k)) == 0) { exec = g.exec; break; }
    if (exec == nullptr) {
      if (g_adjoint_graphs.size() >= kMaxAdjointGraphs) {
        NODE_CUDA_OK(cudaDeviceSynchronize());            // no launch of an evicted graph may still be running
        for (auto& g : g_adjoint_graphs) { cudaGraphExecDestroy(g.exec); cudaGraphDestroy(g.graph); }
        g_adjoint_graphs.clear();
      }
      AdjointSolveGraph ng;
      const int rc = adjoint_build_graph(k, &ng);
      if (rc != 0) return rc;
      g_adjoint_graphs.push_back(ng);
      exec = ng.exec;
    }
  }
  return (int)cudaGraphLaunch(exec, st);
}

// Drops every cached loop graph (the buffers they point at are about to be released).
extern "C" int node_b200_adjoint_solve_reset(void) {
  std::lock_guard<std::mutex> lock(g_adjoint_mutex);
  if (!g_adjoint_graphs.empty()) NODE_CUDA_OK(cudaDeviceSynchronize());
  for (auto& g : g_adjoint_graphs) { cudaGraphExecDestroy(g.exec); cudaGraphDestroy(g.graph); }
  g_adjoint_graphs.clear();
  return 0;
}
