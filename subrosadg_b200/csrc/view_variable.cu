// view_variable.cu — ViewVariable::get (src/Solver/VariableConvertor.cpp:754-872) on the device: the scalar fields the reference's VTU
// writer derives from the conserved variables (and, for Navier-Stokes, from the primitive gradients) — here evaluated at the volume
// quadrature points of the resident state, so that a monitor or an in-situ writer needs no host pass over the raw files.
// The switch of the reference falls through where a variable does not exist for the equation set (Entropy of an incompressible model ->
// Vorticity; Vorticity of an Euler model -> ArtificialViscosity; the directional vorticity / heat-flux entries of an Euler model -> 0);
// the same chain is kept.
#include <cuda_runtime.h>

#include <stdexcept>

#include "dev_util.cuh"
#include "physics.cuh"
#include "view_variable.hpp"

namespace sdg {

namespace {

template <int D>
__global__ void viewVariableKernel(PhysParams P, int variable, size_t npts, const double* __restrict__ cons, const double* __restrict__ grad,
                                   const double* __restrict__ eps, double* __restrict__ out) {
  constexpr int NV = D + 2, NC = D + 3, G = NV * D;
  const Phys<0> ph(P);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += (size_t)gridDim.x * blockDim.x) {
    double c[NV], comp[NC], gp[G];
    for (int k = 0; k < NV; k++) c[k] = cons[i * NV + k];
    compFromCons<D>(ph, c, comp);
    const double rho = comp[0], p = comp[D + 2];
    if (grad != nullptr) {
      double g[G];
      for (int k = 0; k < G; k++) g[k] = grad[i * G + k];
      primGradFromConsGrad<D>(ph, c, comp, g, gp);   // VariableGradient::calculatePrimitiveFromConserved, RawBinary.cpp:228-232
    } else {
      for (int k = 0; k < G; k++) gp[k] = 0.0;
    }
    const bool ns = P.ns != 0;
    auto dUdX = [&](int comp_, int dir) { return gp[(1 + comp_) * D + dir]; };   // d(velocity comp_) / d(x_dir)
    double r = 0.0;
    int v = variable;
    for (;;) {   // the fall-through chain of the reference's switch
      if (v == 0) { r = rho; break; }
      if (v == 1) { r = sqrt(vsq<D>(comp)); break; }
      if (v == 2) { r = ph.TFromE(comp[D + 1]); break; }
      if (v == 3) { r = p; break; }
      if (v == 4) { r = ph.ideal() ? sqrt(P.gamma * p / rho) : P.c0; break; }
      if (v == 5) { r = sqrt(vsq<D>(comp)) / (ph.ideal() ? sqrt(P.gamma * p / rho) : P.c0); break; }
      if (v == 6) { if (P.compressible) { r = p / pow(rho, 1.4); break; } v = 7; continue; }   // kSpecificHeatRatio, PhysicalModel.cpp:148-150
      if (v == 7) {
        if (ns && D == 2) { r = dUdX(1, 0) - dUdX(0, 1); break; }
        if (ns && D == 3) {
          const double a = dUdX(2 % D, 1 % D) - dUdX(1 % D, 2 % D), b = dUdX(0, 2 % D) - dUdX(2 % D, 0), cc = dUdX(1 % D, 0) - dUdX(0, 1 % D);
          r = sqrt(a * a + b * b + cc * cc); break;
        }
        v = 9; continue;
      }
      if (v == 9) { r = eps != nullptr ? eps[i] : 0.0; break; }
      if (v >= 10 && v <= 12) { r = comp[1 + (v - 10)]; break; }
      if (v >= 13 && v <= 15) { r = comp[1 + (v - 13)] / (ph.ideal() ? sqrt(P.gamma * p / rho) : P.c0); break; }
      if (v == 16 && ns) { r = dUdX(2 % D, 1 % D) - dUdX(1 % D, 2 % D); break; }
      if (v >= 16 && v <= 17 && ns) { r = dUdX(0, 2 % D) - dUdX(2 % D, 0); break; }
      if (v >= 16 && v <= 18 && ns) { r = dUdX(1 % D, 0) - dUdX(0, 1 % D); break; }
      if (v >= 16 && v <= 19 && ns) { r = gp[(D + 1) * D + 0]; break; }
      if (v >= 16 && v <= 20 && ns) { r = gp[(D + 1) * D + 1 % D]; break; }
      if (v >= 16 && v <= 21 && ns) { r = gp[(D + 1) * D + 2 % D]; break; }
      r = 0.0; break;   // HeatFlux (8) and everything an Euler model falls through to: default
    }
    out[i] = r;
  }
}

}  // namespace

void launchViewVariable(int D, const PhysParams& P, int variable, size_t npts, const double* cons, const double* grad, const double* eps, double* out,
                        cudaStream_t stream) {
  if (variable < 0 || variable > 21) throw std::runtime_error("ViewVariableEnum value out of range");
  const int needDim = (variable == 11 || variable == 14) ? 2 : (variable == 12 || variable == 15) ? 3 : 1;
  if (D < needDim) throw std::runtime_error("this view variable does not exist in this dimension");
  if (P.ns && D < 3 && (variable == 16 || variable == 17 || variable == 21)) throw std::runtime_error("this view variable needs three dimensions");
  if (P.ns && D < 2 && (variable == 18 || variable == 20)) throw std::runtime_error("this view variable needs two dimensions");
  if (npts == 0) return;
  const int blocks = (int)std::min<size_t>((npts + 255) / 256, 148 * 16);
  if (D == 1) viewVariableKernel<1><<<blocks, 256, 0, stream>>>(P, variable, npts, cons, grad, eps, out);
  else if (D == 2) viewVariableKernel<2><<<blocks, 256, 0, stream>>>(P, variable, npts, cons, grad, eps, out);
  else viewVariableKernel<3><<<blocks, 256, 0, stream>>>(P, variable, npts, cons, grad, eps, out);
  CUDA_OK(cudaGetLastError());
}

}  // namespace sdg
