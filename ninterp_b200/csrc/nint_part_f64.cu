#define NINT_T double
#include "nint_part_inst.cuh"
