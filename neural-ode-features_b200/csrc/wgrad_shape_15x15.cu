// Weight-gradient GEMM on 15x15 maps (the callers' ResBlock convolutions).
#include "wgrad_engine.cuh"
NODE_WGRAD_SHAPE_TU(15, 15)
