// Dense 8x8 adjoint engine (vjp8_engine.cuh): its own translation unit so that it compiles in parallel with the strip shapes.
#include "vjp8_engine.cuh"
namespace node { int launch_vjp8_dense(const VjpArgs& a, cudaStream_t st, int* grid_out) { return v8::launch_vjp8(a, st, grid_out); } }
