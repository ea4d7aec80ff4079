// Translation unit of the persistent cpgmres kernel (kernels/cpgmres.m of the reference):
// compiled on its own so the six solvers build in parallel.
#include "cpk_solvers.cuh"
CPK_DEFINE_SOLVER_TU(4, cpgmres)
