// launchers.hpp — kernel selection tables of the tensor path, compiled in their own translation units (euler_launch.cu,
// ns_launch.cu) so that the template instantiations build in parallel with the C ABI.
#pragma once
#include <cuda_runtime.h>

#include "tensor_kernels.cuh"

namespace sdg {

using StageFn = void (*)(const StageArgs&, int nBlocks, cudaStream_t);

// Euler stage kernel for (D, N = p + 1): sets the chunk size K the kernel is compiled for
StageFn pickEulerFn(int D, int N, bool affine, int ph, int& K);
// NS gradient pass + residual pass
void pickNsFn(int D, int N, bool affine, int ph, StageFn& grad, StageFn& stage, int& K);

// trace-based line kernels for P3 hexahedra (nsl_kernels.cuh): U -> TU, boundary virtual traces, pass G, pass R (visc = false: the
// same residual pass without viscous terms, i.e. an Euler stage through published traces)
using BoundaryFn = void (*)(const StageArgs&, const int4* bndRec, int nBnd, cudaStream_t);
struct LineFns { StageFn trace = nullptr, grad = nullptr, stage = nullptr; BoundaryFn boundary = nullptr; };
void pickNslFns(bool affine, int ph, bool visc, bool gather, LineFns& out, int& K);   // gather (inviscid): no published traces, partners interpolated from their nodal states

}  // namespace sdg
