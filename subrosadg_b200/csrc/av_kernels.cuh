// av_kernels.cuh — ShockCapturingEnum::ArtificialViscosity, the once-per-step part (Solver::calculateArtificialViscosity,
// src/Solver/SpatialDiscrete.cpp:37-192): smoothness indicator per element, maximum over the elements sharing a corner node, node values
// back to the elements' corners.  The per-stage part (eps * grad(U) in the volume and face fluxes) lives in the stage kernels.
#pragma once
#include <cuda_runtime.h>

namespace sdg {

// calculateElementArtificialViscosity (:37-87), collocation form: the density at the nodes IS variable_density_all_order; the part carried by
// the modes above order P-1 is H u with H = Phi[:, high] Phi^-1[high, :] (built on the host).  One warp per element.
static __global__ void avIndicatorKernel(const double* __restrict__ U, const double* __restrict__ H, const double* __restrict__ geoE, const double* __restrict__ invjw,
                                         const double* __restrict__ wq, int affine, int REC, int DD, int n, int NV, int NN, int p, double tol, double empTol,
                                         double factor, const double* __restrict__ radius, double* __restrict__ avE) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
  for (int e = warp; e < n; e += nWarps) {
    const double* u = U + (size_t)e * NV * NN;   // field 0: density
    double num = 0.0, den = 0.0;
    for (int q = lane; q < NN; q += 32) {
      double high = 0.0;
      for (int r = 0; r < NN; r++) high += H[(size_t)q * NN + r] * u[r];
      const double w = affine ? geoE[(size_t)e * REC + DD] * wq[q] : 1.0 / invjw[(size_t)e * NN + q];
      num += high * (high * w); den += u[q] * (u[q] * w);
    }
    for (int o = 16; o > 0; o >>= 1) { num += __shfl_xor_sync(0xffffffffu, num, o); den += __shfl_xor_sync(0xffffffffu, den, o); }
    if (lane == 0) {
      const double shock = log10(num / den);   // http://persson.berkeley.edu/pub/persson13transient_shocks.pdf
      const double full = factor * (radius[e] / p);
      double val;
      if (shock < tol - empTol) val = 0.0;
      else if (shock > tol + empTol) val = full;
      else val = full * (1.0 + sin(3.14159265358979323846 * (shock - tol) / (2.0 * empTol))) / 2.0;
      avE[e] = val;
    }
  }
}

// maxElementArtificialViscosity (:89-108).  The values are >= 0, so the maximum of the doubles is the maximum of their bit patterns:
// an integer atomicMax, independent of the order of arrival.
static __global__ void avNodeMaxKernel(const double* __restrict__ avE, const int* __restrict__ tags, int n, int NB, unsigned long long* __restrict__ node) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * NB; i += gridDim.x * blockDim.x) {
    const double v = avE[i / NB];
    if (v > 0.0) atomicMax(node + tags[i], (unsigned long long)__double_as_longlong(v));
  }
}

// storeElementArtificialViscosity (:110-122)
static __global__ void avStoreKernel(const double* __restrict__ node, const int* __restrict__ tags, int n, int NB, double* __restrict__ avElem) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * NB; i += gridDim.x * blockDim.x) avElem[i] = node[tags[i]];
}

// artificial_viscosity_ at the volume nodes, caller element order: out[e][q] = sum_k tabQ[q][k] avElem[perm[e]][k]
static __global__ void avAtNodesKernel(const double* __restrict__ avElem, const double* __restrict__ tabQ, const int* __restrict__ perm, int n, int NN, int NB,
                                       double* __restrict__ out) {
  const size_t total = (size_t)n * NN;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(i / NN), q = (int)(i - (size_t)e * NN);
    const int pos = perm ? perm[e] : e;
    double s = 0.0;
    for (int k = 0; k < NB; k++) s += tabQ[q * NB + k] * avElem[(size_t)pos * NB + k];
    out[i] = s;
  }
}

}  // namespace sdg
