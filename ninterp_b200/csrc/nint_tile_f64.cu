#define NINT_T double
#include "nint_tile_inst.cuh"
