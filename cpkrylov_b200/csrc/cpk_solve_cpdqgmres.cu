// Translation unit of the persistent cpdqgmres kernel (kernels/cpdqgmres.m of the reference):
// compiled on its own so the six solvers build in parallel.
#include "cpk_solvers.cuh"
CPK_DEFINE_SOLVER_TU(5, cpdqgmres)
