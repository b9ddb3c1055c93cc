// nsl_launch.cu — instantiations and selection of the trace-based line kernels for P3 hexahedra (nsl_kernels.cuh).
#include <stdexcept>

#include "dev_util.cuh"
#include "launchers.hpp"
#include "nsl_kernels.cuh"

namespace sdg {

namespace {

void launchNslTrace(const StageArgs& a, int nBlocks, cudaStream_t s) { nslTraceKernel<<<nBlocks, 128, sizeof(double) * kLK * 5 * 64, s>>>(a); }

template <bool AFFINE>
void launchNslBoundary(const StageArgs& a, const int4* rec, int nBnd, cudaStream_t s) {
  if (nBnd > 0) nslBoundaryKernel<AFFINE><<<(nBnd * 16 + 127) / 128, 128, 0, s>>>(a, rec, nBnd);
}

template <bool AFFINE>
void launchNslGrad(const StageArgs& a, int nBlocks, cudaStream_t s) {
  using L = NslGradLayout<AFFINE>;
  static std::atomic<unsigned long long> configured{0};
  if (firstUseOnThisDevice(configured)) CUDA_OK(cudaFuncSetAttribute(nslGradKernel<AFFINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
  nslGradKernel<AFFINE><<<nBlocks, 128, L::bytes, s>>>(a);
}

template <bool AFFINE, int PH, bool VISC, bool GATHER>
void launchNslStage(const StageArgs& a, int nBlocks, cudaStream_t s) {
  using L = NslStageLayout<GATHER, NslStageUsm<VISC, GATHER>::value>;
  static std::atomic<unsigned long long> configured{0};
  if (firstUseOnThisDevice(configured)) CUDA_OK(cudaFuncSetAttribute(nslStageKernel<AFFINE, PH, VISC, GATHER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
  nslStageKernel<AFFINE, PH, VISC, GATHER><<<nBlocks, 128, L::bytes, s>>>(a);
}

template <bool AFFINE, bool VISC, bool GATHER>
StageFn pickStage(int ph) { return ph ? launchNslStage<AFFINE, 1, VISC, GATHER> : launchNslStage<AFFINE, 0, VISC, GATHER>; }

}  // namespace

void pickNslFns(bool affine, int ph, bool visc, bool gather, LineFns& out, int& K) {
  K = kLK;
  out.trace = launchNslTrace;
  out.boundary = affine ? launchNslBoundary<true> : launchNslBoundary<false>;
  out.grad = affine ? launchNslGrad<true> : launchNslGrad<false>;
  if (visc && gather) throw std::runtime_error("the gathering residual pass is inviscid");
  if (affine) out.stage = visc ? pickStage<true, true, false>(ph) : gather ? pickStage<true, false, true>(ph) : pickStage<true, false, false>(ph);
  else out.stage = visc ? pickStage<false, true, false>(ph) : gather ? pickStage<false, false, true>(ph) : pickStage<false, false, false>(ph);
}

}  // namespace sdg
