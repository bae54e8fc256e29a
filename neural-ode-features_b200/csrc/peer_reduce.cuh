// Peer-memory all-reduce of the solver's float64 sums (peer_reduce.cu).
#pragma once
#include <cuda_runtime.h>

namespace node {

constexpr int kPeerMaxWorld = 8, kPeerMaxRows = 16;
constexpr int kPeerFlagsOff = 8, kPeerValsOff = 8 + 2 * kPeerMaxWorld * 8;
constexpr int kPeerBytes = 4096;
static_assert(kPeerValsOff + 2 * kPeerMaxWorld * kPeerMaxRows * 8 <= kPeerBytes, "exchange buffer");

struct PeerCtx {
  int world, rank;
  void* buf[kPeerMaxWorld];      // every rank's exchange buffer, addressable from this GPU (CUDA IPC over NVLink)
};

const PeerCtx& peer_ctx();
// sums[row] = sum_b partials[row][b] for row < nrows, then (world > 1) summed over the ranks in rank order
int launch_fold_reduce(const double* partials, int nblocks, double* sums, int nrows, int* status, cudaStream_t st);

}  // namespace node
