// view_variable.hpp — launcher of the device restatement of ViewVariable::get (VariableConvertor.cpp:754-872), see view_variable.cu
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

#include "physics.cuh"

namespace sdg {

// cons [npts][NV], grad [npts][NV*D] (conserved gradient; nullptr for Euler models), eps [npts] (artificial viscosity; nullptr without it),
// variable = ViewVariableEnum value (src/Utils/Enum.cpp)
void launchViewVariable(int D, const PhysParams& P, int variable, size_t npts, const double* cons, const double* grad, const double* eps, double* out,
                        cudaStream_t stream);

}  // namespace sdg
