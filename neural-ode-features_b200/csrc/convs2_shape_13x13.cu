#include "convs2_engine.cuh"
NODE_CONVS2_SHAPE_TU(7, 7, 13, 13)
