// physics_debug.cu — sdg_debug_physics: the pointwise device functions of physics.cuh evaluated on caller-supplied points (parity hook).
// Same entry-point shape as the reference driver oracle/ref_physics.cpp and the oracle's orc_physics, so that one golden file
// (tests/golden/reference_physics.json, generated from the reference's own sources) checks all three.  Needs a CUDA device.
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "../../include/subrosadg_b200.h"
#include "dev_util.cuh"
#include "physics.cuh"
#include "view_variable.hpp"
#include <vector>

namespace sdg {

// what: 0 Riemann flux, 1 boundary face point, 2 viscous terms, 3 conversions / raw flux / source (layouts: reference_physics.json "layout");
// what 4 (view variables) runs the product's viewVariableKernel instead, see sdg_debug_physics
template <int D, int PH>
__global__ void physicsDebugKernel(PhysParams P, int what, int bc, int n, const double* __restrict__ in, double* __restrict__ out) {
  constexpr int NV = D + 2, NC = D + 3, G = NV * D;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Phys<PH> ph(P);
  if (what == 0) {
    const double* a = in + (size_t)i * (D + 2 * NV);
    double nrm[D], consL[NV], consR[NV], compL[NC], compR[NC], F[NV];
    for (int d = 0; d < D; d++) nrm[d] = a[d];
    for (int k = 0; k < NV; k++) { consL[k] = a[D + k]; consR[k] = a[D + NV + k]; }
    const double irL = compFromCons<D>(ph, consL, compL), irR = compFromCons<D>(ph, consR, compR);
    convFlux<D>(ph, nrm, consL, compL, irL, consR, compR, irR, F);
    for (int k = 0; k < NV; k++) out[(size_t)i * NV + k] = F[k];
  } else if (what == 1) {
    const int ni = D + 2 * NV + G, no = NC + 3 * NV + (P.ns ? NC + NV : 0);
    const double* a = in + (size_t)i * ni; double* o = out + (size_t)i * no;
    double nrm[D], consL[NV], compL[NC], prim[NV], compR[NC], b[NC], vol[NV], itf[NV], Fn[NV];
    for (int d = 0; d < D; d++) nrm[d] = a[d];
    for (int k = 0; k < NV; k++) { consL[k] = a[D + k]; prim[k] = a[D + NV + k]; }
    compFromCons<D>(ph, consL, compL);
    compFromPrim<D>(ph, prim, compR);                       // boundary_dummy_variable_, InitialCondition.cpp:118-149
    bcBoundaryVariable<D>(ph, bc, nrm, compL, compR, b);
    bcBoundaryGradientVariable<D>(ph, bc, nrm, consL, compL, compR, vol, itf);
    convNormalFlux<D>(ph, nrm, b, Fn);
    for (int k = 0; k < NC; k++) o[k] = b[k];
    for (int k = 0; k < NV; k++) { o[NC + k] = vol[k]; o[NC + NV + k] = itf[k]; o[NC + 2 * NV + k] = Fn[k]; }
    if (P.ns) {
      double g[G], pL[G], gb[G], va[NV], vb[NV];
      for (int k = 0; k < G; k++) g[k] = a[D + 2 * NV + k];
      primGradFromConsGrad<D>(ph, consL, compL, g, pL);
      if (bcIsWall(bc)) for (int k = 0; k < NC; k++) compL[k] = b[k];
      for (int k = 0; k < G; k++) gb[k] = pL[k];
      if (bc == kAdiabaticSlipWall || bc == kAdiabaticNonSlipWall) for (int d = 0; d < D; d++) gb[(D + 1) * D + d] = 0.0;
      viscNormalFlux<D>(ph, nrm, compL, pL, va);
      viscNormalFlux<D>(ph, nrm, b, gb, vb);
      for (int k = 0; k < NC; k++) o[NC + 3 * NV + k] = compL[k];
      for (int k = 0; k < NV; k++) o[2 * NC + 3 * NV + k] = (va[k] + vb[k]) / 2.0;
    }
  } else if (what == 2) {
    const int ni = D + NV + G, no = 2 * G + NV;
    const double* a = in + (size_t)i * ni; double* o = out + (size_t)i * no;
    double nrm[D], cons[NV], comp[NC], g[G], gp[G], F[G], Fn[NV];
    for (int d = 0; d < D; d++) nrm[d] = a[d];
    for (int k = 0; k < NV; k++) cons[k] = a[D + k];
    for (int k = 0; k < G; k++) g[k] = a[D + NV + k];
    compFromCons<D>(ph, cons, comp);
    primGradFromConsGrad<D>(ph, cons, comp, g, gp);
    viscRawFlux<D>(ph, comp, gp, F);
    viscNormalFlux<D>(ph, nrm, comp, gp, Fn);
    for (int k = 0; k < G; k++) { o[k] = gp[k]; o[G + k] = F[k]; }
    for (int k = 0; k < NV; k++) o[2 * G + k] = Fn[k];
  } else {
    const int no = NC + NV + G + NV;
    double* o = out + (size_t)i * no;
    double cons[NV], comp[NC], F[G];
    for (int k = 0; k < NV; k++) cons[k] = in[(size_t)i * NV + k];
    compFromCons<D>(ph, cons, comp);
    for (int k = 0; k < NC; k++) o[k] = comp[k];
    o[NC] = comp[0];
    for (int d = 0; d < D; d++) o[NC + 1 + d] = comp[1 + d];
    o[NC + D + 1] = ph.TFromE(comp[D + 1]);
    convRawFlux<D>(ph, comp, F);
    for (int k = 0; k < G; k++) o[NC + NV + k] = F[k];
    for (int k = 0; k < NV; k++) o[NC + NV + G + k] = 0.0;
    if (P.source == kBoussinesq && D >= 2) o[NC + NV + G + D] = boussinesqSource<D>(ph, comp);
  }
}

}  // namespace sdg

using namespace sdg;

extern "C" int sdg_debug_physics(const int32_t* cfg, const double* params, int32_t what, int32_t bc, int32_t n, const double* in, double* out) {
  static thread_local std::string err;
  try {
    const int D = cfg[0];
    if (D < 1 || D > 3 || what < 0 || what > 4 || n < 0) throw std::runtime_error("sdg_debug_physics: bad arguments");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) throw std::runtime_error("sdg_debug_physics: no CUDA device — this library has no CPU path");
    PhysParams P{};
    P.model = cfg[1]; P.eos = cfg[2]; P.transport = cfg[3]; P.conv = cfg[4]; P.source = cfg[5];
    P.compressible = (P.model == kCompresibleEuler || P.model == kCompresibleNS) ? 1 : 0;
    P.ns = (P.model == kCompresibleNS || P.model == kIncompresibleNS) ? 1 : 0;
    P.visc = P.ns ? kBR2 : kViscNone;
    P.cp = params[0]; P.cv = params[1]; P.icv = 1.0 / params[1]; P.gamma = 1.4; P.kg = 0.5 * (1.4 + 1.0) / 1.4;
    P.mu0 = params[2]; P.k0 = params[0] * params[2] / 0.71; P.c0 = params[3]; P.rho0 = params[4]; P.padd = 0.01 * params[4] * params[3] * params[3];
    P.beta = params[5]; P.tref = params[6];
    const int NV = D + 2, NC = D + 3, G = NV * D;
    if (what == 4) {
      // the 22 ViewVariableEnum values at caller-supplied points through the product's viewVariableKernel (view_variable.cu):
      // in = cons[NV], conserved gradient[G], artificial viscosity; 0 where the variable names a direction the dimension lacks
      std::vector<double> hc((size_t)n * NV), hg((size_t)n * G), he((size_t)n), col((size_t)n);
      for (int i = 0; i < n; i++) {
        const double* a = in + (size_t)i * (NV + G + 1);
        for (int k = 0; k < NV; k++) hc[(size_t)i * NV + k] = a[k];
        for (int k = 0; k < G; k++) hg[(size_t)i * G + k] = a[NV + k];
        he[i] = a[NV + G];
      }
      DevBuf<double> dc, dg, de, dr;
      dc.alloc(hc.size() + 1); dg.alloc(hg.size() + 1); de.alloc(he.size() + 1); dr.alloc((size_t)n + 1);
      CUDA_OK(cudaMemcpy(dc.p, hc.data(), hc.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(dg.p, hg.data(), hg.size() * sizeof(double), cudaMemcpyHostToDevice));
      CUDA_OK(cudaMemcpy(de.p, he.data(), he.size() * sizeof(double), cudaMemcpyHostToDevice));
      for (int w = 0; w < 22; w++) {
        const bool needs3 = w == 12 || w == 15 || w == 16 || w == 17 || w == 21, needs2 = w == 11 || w == 14 || w == 18 || w == 20;
        if ((needs3 && D < 3) || (needs2 && D < 2)) { for (int i = 0; i < n; i++) out[(size_t)i * 22 + w] = 0.0; continue; }
        if (n > 0) launchViewVariable(D, P, w, (size_t)n, dc.p, P.ns ? dg.p : nullptr, de.p, dr.p, nullptr);
        CUDA_OK(cudaGetLastError());
        CUDA_OK(cudaMemcpy(col.data(), dr.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n; i++) out[(size_t)i * 22 + w] = col[i];
      }
      return 0;
    }
    const int ni = what == 0 ? D + 2 * NV : what == 1 ? D + 2 * NV + G : what == 2 ? D + NV + G : NV;
    const int no = what == 0 ? NV : what == 1 ? NC + 3 * NV + (P.ns ? NC + NV : 0) : what == 2 ? 2 * G + NV : NC + NV + G + NV;
    if (what == 2 && !P.ns) throw std::runtime_error("sdg_debug_physics: viscous terms need a Navier-Stokes model");
    DevBuf<double> din, dout;
    din.alloc((size_t)std::max(n, 1) * ni); dout.alloc((size_t)std::max(n, 1) * no);
    CUDA_OK(cudaMemcpy(din.p, in, sizeof(double) * (size_t)n * ni, cudaMemcpyHostToDevice));
    const int blocks = (n + 63) / 64;
    // PH = 1 is the compile-time specialisation (compressible, ideal gas, HLLC) of the benchmark configurations: exercise it where it applies
    const bool ph1 = P.compressible && P.eos == kIdealGas && P.conv == kHLLC;
    if (n > 0) {
      if (D == 1) { if (ph1) physicsDebugKernel<1, 1><<<blocks, 64>>>(P, what, bc, n, din.p, dout.p); else physicsDebugKernel<1, 0><<<blocks, 64>>>(P, what, bc, n, din.p, dout.p); }
      else if (D == 2) { if (ph1) physicsDebugKernel<2, 1><<<blocks, 64>>>(P, what, bc, n, din.p, dout.p); else physicsDebugKernel<2, 0><<<blocks, 64>>>(P, what, bc, n, din.p, dout.p); }
      else { if (ph1) physicsDebugKernel<3, 1><<<blocks, 64>>>(P, what, bc, n, din.p, dout.p); else physicsDebugKernel<3, 0><<<blocks, 64>>>(P, what, bc, n, din.p, dout.p); }
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpy(out, dout.p, sizeof(double) * (size_t)n * no, cudaMemcpyDeviceToHost));
  } catch (const std::exception& ex) {
    err = ex.what();
    sdg::setLastError(err.c_str());
    return 1;
  }
  return 0;
}
