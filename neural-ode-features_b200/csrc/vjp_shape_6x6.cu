// Adjoint kernels (K7) for 6x6 feature maps: fused VJP + weight-gradient GEMM.
#include "vjp_engine.cuh"
#include "wgrad_engine.cuh"
NODE_VJP_SHAPE_TU(6, 6)
