// Interp1D/2D/3D kernels, T = double
#define NINT_T double
#include "nint_inst_fixed.cuh"
