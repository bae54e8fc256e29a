#include "step_engine.cuh"
NODE_STEP_SHAPE_TU(6, 6)
