// euler_launch.cu — instantiations and selection of the Euler stage kernels (tensor_kernels.cuh, line_kernels.cuh).
#include <cstdlib>
#include <stdexcept>

#include "dev_util.cuh"
#include "launchers.hpp"
#include "line_kernels.cuh"

namespace sdg {

namespace {


template <int D, int N, int K, bool AFFINE, int PH>
void launchEuler(const StageArgs& a, int nBlocks, cudaStream_t s) {
  using L = Layout<D, N, K>;
  static std::atomic<unsigned long long> configured{0};
  if (firstUseOnThisDevice(configured)) CUDA_OK(cudaFuncSetAttribute(eulerStageKernel<D, N, K, AFFINE, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
  eulerStageKernel<D, N, K, AFFINE, PH><<<nBlocks, kThreads, L::bytes, s>>>(a);
}

template <int N, int K, bool AFFINE, int PH>
void launchEulerLine(const StageArgs& a, int nBlocks, cudaStream_t s) {
  using L = LineLayout<N, K>;
  static std::atomic<unsigned long long> configured{0};
  if (firstUseOnThisDevice(configured)) CUDA_OK(cudaFuncSetAttribute(eulerLineKernel<N, K, AFFINE, PH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::bytes));
  eulerLineKernel<N, K, AFFINE, PH><<<nBlocks, L::THREADS, L::bytes, s>>>(a);
}

// chunk sizes: bricks of 2^D .. elements, sized so that two blocks fit an SM
template <int D, int N> struct ChunkOf;
template <> struct ChunkOf<1, 2> { static constexpr int K = 64; };
template <> struct ChunkOf<1, 3> { static constexpr int K = 64; };
template <> struct ChunkOf<1, 4> { static constexpr int K = 64; };
template <> struct ChunkOf<1, 5> { static constexpr int K = 32; };
template <> struct ChunkOf<1, 6> { static constexpr int K = 32; };
template <> struct ChunkOf<2, 2> { static constexpr int K = 64; };
template <> struct ChunkOf<2, 3> { static constexpr int K = 32; };
template <> struct ChunkOf<2, 4> { static constexpr int K = 16; };
template <> struct ChunkOf<3, 2> { static constexpr int K = 32; };
template <> struct ChunkOf<3, 3> { static constexpr int K = 8; };
template <> struct ChunkOf<3, 4> { static constexpr int K = 8; };
template <> struct ChunkOf<2, 5> { static constexpr int K = 16; };
template <> struct ChunkOf<2, 6> { static constexpr int K = 8; };
template <> struct ChunkOf<3, 5> { static constexpr int K = 2; };
template <> struct ChunkOf<3, 6> { static constexpr int K = 1; };

template <int D, int N>
StageFn pickEuler(bool affine, int ph) {
  constexpr int K = ChunkOf<D, N>::K;
  if (affine) return ph ? launchEuler<D, N, K, true, 1> : launchEuler<D, N, K, true, 0>;
  return ph ? launchEuler<D, N, K, false, 1> : launchEuler<D, N, K, false, 0>;
}
}  // namespace

StageFn pickEulerFn(int D, int N, bool affine, int ph, int& K) {
  if (D == 1 && N == 2) { K = ChunkOf<1, 2>::K; return pickEuler<1, 2>(affine, ph); }
  if (D == 1 && N == 3) { K = ChunkOf<1, 3>::K; return pickEuler<1, 3>(affine, ph); }
  if (D == 1 && N == 4) { K = ChunkOf<1, 4>::K; return pickEuler<1, 4>(affine, ph); }
  if (D == 1 && N == 5) { K = ChunkOf<1, 5>::K; return pickEuler<1, 5>(affine, ph); }
  if (D == 1 && N == 6) { K = ChunkOf<1, 6>::K; return pickEuler<1, 6>(affine, ph); }
  if (D == 2 && N == 2) { K = ChunkOf<2, 2>::K; return pickEuler<2, 2>(affine, ph); }
  if (D == 2 && N == 3) { K = ChunkOf<2, 3>::K; return pickEuler<2, 3>(affine, ph); }
  if (D == 2 && N == 4) { K = ChunkOf<2, 4>::K; return pickEuler<2, 4>(affine, ph); }
  if (D == 3 && N == 2) { K = ChunkOf<3, 2>::K; return pickEuler<3, 2>(affine, ph); }
  if (D == 3 && N == 3) { K = ChunkOf<3, 3>::K; return pickEuler<3, 3>(affine, ph); }
  if (D == 3 && N == 4) {
    K = ChunkOf<3, 4>::K;
    if (getenv("SDG_NODE_KERNEL")) return pickEuler<3, 4>(affine, ph);   // A/B switch: node-per-thread kernel of tensor_kernels.cuh
    constexpr int KK = ChunkOf<3, 4>::K;
    if (affine) return ph ? launchEulerLine<4, KK, true, 1> : launchEulerLine<4, KK, true, 0>;
    return ph ? launchEulerLine<4, KK, false, 1> : launchEulerLine<4, KK, false, 0>;
  }
  if (D == 2 && N == 5) { K = ChunkOf<2, 5>::K; return pickEuler<2, 5>(affine, ph); }
  if (D == 2 && N == 6) { K = ChunkOf<2, 6>::K; return pickEuler<2, 6>(affine, ph); }
  if (D == 3 && N == 5) { K = ChunkOf<3, 5>::K; return pickEuler<3, 5>(affine, ph); }
  if (D == 3 && N == 6) { K = ChunkOf<3, 6>::K; return pickEuler<3, 6>(affine, ph); }
  throw std::runtime_error("device path implements line/quadrangle/hexahedron blocks with p = 1..5");
}


}  // namespace sdg
