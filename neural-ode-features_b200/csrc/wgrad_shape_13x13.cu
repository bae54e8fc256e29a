// Weight-gradient GEMM on 13x13 maps (the callers' ResBlock convolutions on MNIST).
#include "wgrad_engine.cuh"
NODE_WGRAD_SHAPE_TU(13, 13)
