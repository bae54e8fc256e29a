// Dense 8x8 engine (step8_engine.cuh): its own translation unit so that it compiles in parallel with the strip shapes.
#include "step8_engine.cuh"
namespace node { int launch_step8_dense(const FusedArgs& a, cudaStream_t st) { return s8::launch_step8(a, st); } }
