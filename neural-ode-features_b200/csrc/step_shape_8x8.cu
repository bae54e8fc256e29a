// NODE_STEP_DEBUG (tuning aid, off by default: build with -DNODE_STEP_DEBUG) exports its reader from this shape only
#define NODE_STEP_DEBUG_EXPORT
#include "step_engine.cuh"
NODE_STEP_SHAPE_TU(8, 8)
