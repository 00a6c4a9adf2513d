#define NINT_T float
#include "nint_tile_inst.cuh"
