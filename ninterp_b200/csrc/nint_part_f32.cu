#define NINT_T float
#include "nint_part_inst.cuh"
