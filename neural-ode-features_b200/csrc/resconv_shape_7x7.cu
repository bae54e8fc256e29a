#include "resconv_engine.cuh"
NODE_RESCONV_SHAPE_TU(7, 7)
