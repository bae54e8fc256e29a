// the batch-global error norm of a SHARDED solve without a host-launched collective.
// Every attempted step needs the sum over the GPUs of <= 16 float64 partial sums (misc.py:123-136, 146-157 are means over the
// whole batch). Round 1 folded the per-CTA partials on the device, called ncclAllReduce on 128 bytes from the host and then
// launched the controller: 7 collectives per forward, each a launch gap plus NCCL's small-message latency, and no CUDA graph.
// Here the fold kernel itself exchanges the sums over NVLink peer memory: one thread per peer stores this rank's sums and a
// sequence flag into the peer's exchange buffer (plain P2P stores, system-scope fence), the kernel then polls its OWN buffer
// until every rank's flag carries the current sequence number and adds the ranks' values in rank order - bit-identical on
// every GPU, so all controllers take the same decision. The whole solve stays one enqueued (and capturable) launch sequence.
//
// Exchange buffer of a rank (cudaMalloc, shared with the peers through CUDA IPC handles that the host side passes around
// with torch.distributed): [0] sequence counter (local), [1 + p*8 + r] flag of rank r for parity p, then values
// [p][r][16] doubles. Two parities: a rank one step ahead never overwrites values a slower peer still reads (it cannot be two
// steps ahead - the step in between needs that peer's flag).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "node_common.cuh"
#include "peer_reduce.cuh"

namespace node {

static PeerCtx g_peer{};                 // one process per GPU: a single context
static void* g_peer_mine = nullptr;
static void* g_peer_opened[kPeerMaxWorld] = {};

const PeerCtx& peer_ctx() { return g_peer; }

__global__ void __launch_bounds__(64) k_fold_reduce(const double* __restrict__ partials, int nblocks, double* __restrict__ sums, int nrows,
                                                    PeerCtx pc, int* status, int status_bit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int row = warp; row < nrows; row += 2) {
    double v = 0.0;
    for (int b = lane; b < nblocks; b += 32) v += partials[(int64_t)row * nblocks + b];
    v = warp_sum(v);
    if (lane == 0) sums[row] = v;
  }
  __syncthreads();
  if (pc.world <= 1) return;
  __shared__ unsigned long long s_seq;
  unsigned long long* mine = reinterpret_cast<unsigned long long*>(pc.buf[pc.rank]);
  if (threadIdx.x == 0) { s_seq = mine[0] + 1ull; mine[0] = s_seq; }
  __syncthreads();
  const unsigned long long seq = s_seq;
  const int p = (int)(seq & 1ull);
  if (threadIdx.x < pc.world) {                       // thread r delivers to rank r (its own rank included)
    char* dst = reinterpret_cast<char*>(pc.buf[threadIdx.x]);
    volatile double* vals = reinterpret_cast<volatile double*>(dst + kPeerValsOff) + (p * kPeerMaxWorld + pc.rank) * kPeerMaxRows;
    for (int i = 0; i < nrows; ++i) vals[i] = sums[i];
    __threadfence_system();
    volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(dst + kPeerFlagsOff) + p * kPeerMaxWorld + pc.rank;
    *flag = seq;
  }
  __syncthreads();
  if (threadIdx.x < pc.world) {                       // thread r waits for rank r's flag in MY buffer
    volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(reinterpret_cast<char*>(mine) + kPeerFlagsOff) +
                                        p * kPeerMaxWorld + threadIdx.x;
    const long long t0 = clock64();
    bool ok = true;
    while (*flag != seq) {
      if (clock64() - t0 > 20000000000ll) { ok = false; break; }      // ~10 s: a peer died; do not hang the GPU
      __nanosleep(64);
    }
    if (!ok && status != nullptr) atomicOr(status, status_bit);
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x < nrows) {
    const volatile double* vals = reinterpret_cast<const volatile double*>(reinterpret_cast<char*>(mine) + kPeerValsOff) + p * kPeerMaxWorld * kPeerMaxRows;
    double s = 0.0;
    for (int r = 0; r < pc.world; ++r) s += vals[r * kPeerMaxRows + threadIdx.x];
    sums[threadIdx.x] = s;
  }
}

int launch_fold_reduce(const double* partials, int nblocks, double* sums, int nrows, int* status, cudaStream_t st) {
  if (nrows < 1 || nrows > kPeerMaxRows) return (int)cudaErrorInvalidValue;
  k_fold_reduce<<<1, 64, 0, st>>>(partials, nblocks, sums, nrows, g_peer, status, NODE_ST_WATCHDOG);
  return (int)cudaGetLastError();
}

}  // namespace node

using namespace node;

extern "C" int node_b200_peer_alloc(void* handle_out_64_bytes) {
  if (g_peer_mine == nullptr) {
    NODE_CUDA_OK(cudaMalloc(&g_peer_mine, kPeerBytes));
    NODE_CUDA_OK(cudaMemset(g_peer_mine, 0, kPeerBytes));
  }
  cudaIpcMemHandle_t h;
  NODE_CUDA_OK(cudaIpcGetMemHandle(&h, g_peer_mine));
  static_assert(sizeof(h) == 64, "CUDA IPC handle size");
  memcpy(handle_out_64_bytes, &h, sizeof(h));
  return 0;
}

extern "C" int node_b200_peer_open(int world, int rank, const void* handles) {
  if (world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || g_peer_mine == nullptr) return (int)cudaErrorInvalidValue;
  PeerCtx pc{};
  pc.world = world; pc.rank = rank;
  for (int r = 0; r < world; ++r) {
    if (r == rank) { pc.buf[r] = g_peer_mine; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + 64 * r, sizeof(h));
    void* p = nullptr;
    NODE_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g_peer_opened[r] = p;
    pc.buf[r] = p;
  }
  NODE_CUDA_OK(cudaMemset(g_peer_mine, 0, kPeerBytes));
  NODE_CUDA_OK(cudaDeviceSynchronize());
  g_peer = pc;
  return 0;
}

extern "C" int node_b200_peer_close(void) {
  g_peer = PeerCtx{};
  for (int r = 0; r < kPeerMaxWorld; ++r)
    if (g_peer_opened[r] != nullptr) { cudaIpcCloseMemHandle(g_peer_opened[r]); g_peer_opened[r] = nullptr; }
  return 0;
}

extern "C" int node_b200_peer_world(void) { return g_peer.world; }

/* fold `nrows` rows of `nblocks` float64 partials and, when peers are configured, all-reduce the rows over them */
extern "C" int node_b200_fold_reduce(const double* partials, int nblocks, double* sums, int nrows, void* stream) {
  return launch_fold_reduce(partials, nblocks, sums, nrows, nullptr, (cudaStream_t)stream);
}
