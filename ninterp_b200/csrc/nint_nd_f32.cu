// InterpND kernels, T = float
#define NINT_T float
#include "nint_inst_nd.cuh"
