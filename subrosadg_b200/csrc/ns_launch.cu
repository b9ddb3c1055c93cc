// ns_launch.cu — instantiations and selection of the Navier-Stokes pass kernels (ns_kernels.cuh).
#include <cstdlib>
#include <stdexcept>

#include "dev_util.cuh"
#include "launchers.hpp"
#include "ns_kernels.cuh"

namespace sdg {

namespace {
template <int D, int N, int K, bool AFFINE>
void launchNsGrad(const StageArgs& a, int nBlocks, cudaStream_t s) {
  using L = NsLayout<D, N, K, AFFINE, false>;
  static std::atomic<unsigned long long> configured{0};
  if (firstUseOnThisDevice(configured)) CUDA_OK(cudaFuncSetAttribute(nsGradKernel<D, N, K, AFFINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
  nsGradKernel<D, N, K, AFFINE><<<nBlocks, kThreads, L::bytes, s>>>(a);
}
// threads per block of the NS residual pass: one face point per thread in the face phase where the register budget allows
// (P3 hexahedra: 2x2x1 brick = 20 faces x 16 points = 320 face points, 256 nodes)
template <int D, int N> struct NsThreadsOf { static constexpr int TH = kThreads; };
#ifndef SDG_NSR_TH34
#define SDG_NSR_TH34 256
#endif
template <> struct NsThreadsOf<3, 4> { static constexpr int TH = SDG_NSR_TH34; };
template <int D, int N, int K, bool AFFINE, int PH>
void launchNsStage(const StageArgs& a, int nBlocks, cudaStream_t s) {
  constexpr int TH = NsThreadsOf<D, N>::TH;
  using L = NsLayout<D, N, K, AFFINE, true, TH>;
  static std::atomic<unsigned long long> configured{0};
  if (firstUseOnThisDevice(configured)) CUDA_OK(cudaFuncSetAttribute(nsStageKernel<D, N, K, AFFINE, PH, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
  nsStageKernel<D, N, K, AFFINE, PH, TH><<<nBlocks, TH, L::bytes, s>>>(a);
}
template <int D, int N> struct NsChunkOf;
template <> struct NsChunkOf<1, 2> { static constexpr int K = 64; };
template <> struct NsChunkOf<1, 3> { static constexpr int K = 64; };
template <> struct NsChunkOf<1, 4> { static constexpr int K = 64; };
template <> struct NsChunkOf<1, 5> { static constexpr int K = 32; };
template <> struct NsChunkOf<1, 6> { static constexpr int K = 32; };
template <> struct NsChunkOf<2, 2> { static constexpr int K = 32; };
template <> struct NsChunkOf<2, 3> { static constexpr int K = 16; };
template <> struct NsChunkOf<2, 4> { static constexpr int K = 16; };
template <> struct NsChunkOf<2, 5> { static constexpr int K = 8; };
template <> struct NsChunkOf<2, 6> { static constexpr int K = 4; };
template <> struct NsChunkOf<3, 2> { static constexpr int K = 16; };
template <> struct NsChunkOf<3, 3> { static constexpr int K = 8; };
#ifndef SDG_NS34_K
#define SDG_NS34_K 4
#endif
template <> struct NsChunkOf<3, 4> { static constexpr int K = SDG_NS34_K; };
template <int D, int N>
void pickNs(bool affine, int ph, StageFn& grad, StageFn& stage, int& K) {
  constexpr int KK = NsChunkOf<D, N>::K;
  K = KK;
  if (affine) { grad = launchNsGrad<D, N, KK, true>; stage = ph ? launchNsStage<D, N, KK, true, 1> : launchNsStage<D, N, KK, true, 0>; }
  else { grad = launchNsGrad<D, N, KK, false>; stage = ph ? launchNsStage<D, N, KK, false, 1> : launchNsStage<D, N, KK, false, 0>; }
}
template <> struct NsChunkOf<3, 5> { static constexpr int K = 2; };
template <> struct NsChunkOf<3, 6> { static constexpr int K = 1; };
}  // namespace

void pickNsFn(int D, int N, bool affine, int ph, StageFn& grad, StageFn& stage, int& K) {
  if (D == 1 && N == 2) return pickNs<1, 2>(affine, ph, grad, stage, K);
  if (D == 1 && N == 3) return pickNs<1, 3>(affine, ph, grad, stage, K);
  if (D == 1 && N == 4) return pickNs<1, 4>(affine, ph, grad, stage, K);
  if (D == 1 && N == 5) return pickNs<1, 5>(affine, ph, grad, stage, K);
  if (D == 1 && N == 6) return pickNs<1, 6>(affine, ph, grad, stage, K);
  if (D == 2 && N == 2) return pickNs<2, 2>(affine, ph, grad, stage, K);
  if (D == 2 && N == 3) return pickNs<2, 3>(affine, ph, grad, stage, K);
  if (D == 2 && N == 4) return pickNs<2, 4>(affine, ph, grad, stage, K);
  if (D == 3 && N == 2) return pickNs<3, 2>(affine, ph, grad, stage, K);
  if (D == 3 && N == 3) return pickNs<3, 3>(affine, ph, grad, stage, K);
  if (D == 3 && N == 4) return pickNs<3, 4>(affine, ph, grad, stage, K);
  if (D == 2 && N == 5) return pickNs<2, 5>(affine, ph, grad, stage, K);
  if (D == 2 && N == 6) return pickNs<2, 6>(affine, ph, grad, stage, K);
  if (D == 3 && N == 5) return pickNs<3, 5>(affine, ph, grad, stage, K);
  if (D == 3 && N == 6) return pickNs<3, 6>(affine, ph, grad, stage, K);
  throw std::runtime_error("device path implements line/quadrangle/hexahedron blocks with p = 1..5");
}


}  // namespace sdg
