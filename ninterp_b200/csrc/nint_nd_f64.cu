// InterpND kernels, T = double
#define NINT_T double
#include "nint_inst_nd.cuh"
