#include "convs2_engine.cuh"
NODE_CONVS2_SHAPE_TU(13, 13, 26, 26)
