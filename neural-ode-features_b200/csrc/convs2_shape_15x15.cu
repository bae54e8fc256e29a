#include "convs2_engine.cuh"
NODE_CONVS2_SHAPE_TU(8, 8, 15, 15)
