#include "convs2_engine.cuh"
NODE_CONVS2_SHAPE_TU(15, 15, 30, 30)
