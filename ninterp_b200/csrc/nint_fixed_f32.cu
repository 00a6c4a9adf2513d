// Interp1D/2D/3D kernels, T = float
#define NINT_T float
#include "nint_inst_fixed.cuh"
