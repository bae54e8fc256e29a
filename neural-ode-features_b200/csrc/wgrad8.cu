// Dense 8x8 weight-gradient GEMM (wgrad8_engine.cuh): the adjoint's K7b and the callers' 8x8 convolutions.
#include "wgrad8_engine.cuh"
namespace node {
int launch_wgrad8_dense(const WgradArgs& a, bool ones, cudaStream_t st) { return w8::launch_wgrad8(a, ones, st); }
}
