// mixed_path.cu — sm_100a kernels + host sequencing of the dense-operator path (triangle / quadrangle blocks in one mesh).
//
// One RK stage = the reference's sweeps (TimeIntegration.cpp:326-350) regrouped into at most four launches per element type:
//   mxFaceKernel<0>   G2  face traces, volume / interface gradient fluxes            (SpatialDiscrete.cpp:844-968)      [NS only]
//   mxGradElemKernel  G1+G3+G4  U_q, gradient residuals, M^-1, BR1 / BR2 lifts       (:294-322,1034-1068, TimeIntegration.cpp:200-228) [NS only]
//   mxFaceKernel<1>   R2  traces (+ lifted gradient traces), Riemann / boundary flux  (:633-842)
//   mxElemKernel      R1+R3+R4+K  volume fluxes, residual, M^-1, RK update, norm      (:194-266,1016-1032, TimeIntegration.cpp:181-198,279-324)
// The Euler gradient sweeps of the reference are dead work without artificial viscosity and are skipped.
// Face fluxes go to per-(element, local face, point) slots — the reference's own race-free rule — so there are no atomics.
// Meshes on this path are the small hybrid meshes of configs 2/3 (1e3-1e4 elements): launch-latency bound, operators stay
// in L1/L2; the HBM-roofline work of the repository is the tensor path.
#include "mixed_path.hpp"
#include "av_kernels.cuh"
#include "view_variable.hpp"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace sdg {

namespace {

constexpr int kD = 2, kNV = 4, kG = 8;
constexpr int kElemThreads = 128;

struct MxType {
  int n, Nb, Nq, Nf, Nqf, Naq, normOff;
  const double *Phi, *dPhi, *PhiF, *Proj, *mt, *jw, *Minv, *minEdge;
  double *U, *Ulast, *R, *A, *RM, *AGv, *AGi, *Gvol, *Gtot, *Gf;
  // artificial viscosity: corner values [n][NB], order-1 nodal basis at the volume / face points, mesh data, scratch
  int NB, nbLow;
  const double *avElem, *NodalQ, *NodalF, *radius;
  const int* tags;
  double* avE;
};
struct MxFaces {
  int nInt, nBnd, Nqf;
  const int *le, *lt, *lf, *re, *rt, *rf, *bc;
  const double *nrm, *fjw, *dummy;
};

}  // namespace

struct MixedSolver::Args {
  MxType t[2];      // slot 0 triangle, slot 1 quadrangle
  MxFaces F;
  PhysParams phys;
  double aLast, aCur, bdt;
  double* normPartial;
  int mode;         // 0 RK update, 1 parity hook (R and R M^-1 written, state untouched)
};

namespace {

using Args = MixedSolver::Args;

__device__ __forceinline__ int slotOf(int type) { return type == kTriangle ? 0 : 1; }

// AdjacencyElementVariable::get, VariableConvertor.cpp:432-485: trace = U Phi_f[row]^T
__device__ __forceinline__ void traceOf(const double* __restrict__ U, const double* __restrict__ phi, int Nb, double* cons) {
  for (int v = 0; v < kNV; v++) cons[v] = 0.0;
  for (int b = 0; b < Nb; b++) {
    const double f = phi[b];
    for (int v = 0; v < kNV; v++) cons[v] = fma(U[b * kNV + v], f, cons[v]);
  }
}
// AdjacencyElementVariableGradient::get<kViscousFlux>, VariableConvertor.cpp:640-723: BR1 total gradient, BR2 volume + this face's lift
__device__ __forceinline__ void gradTraceOf(const MxType& T, int visc, int e, int f, const double* __restrict__ phi, double* g) {
  for (int r = 0; r < kG; r++) g[r] = 0.0;
  const double* C0 = (visc == kBR1 ? T.Gtot : T.Gvol) + (size_t)e * T.Nb * kG;
  const double* C1 = visc == kBR2 ? T.Gf + ((size_t)e * T.Nf + f) * T.Nb * kG : nullptr;
  for (int b = 0; b < T.Nb; b++) {
    const double ph = phi[b];
    for (int r = 0; r < kG; r++) g[r] = fma(C0[b * kG + r] + (C1 ? C1[b * kG + r] : 0.0), ph, g[r]);
  }
}

// eps at a point: order-1 nodal basis row * corner values (SpatialDiscrete.cpp:213-214, 529-631)
__device__ __forceinline__ double mxEps(const MxType& T, const double* __restrict__ row, int e) {
  double s = 0.0;
  for (int k = 0; k < T.NB; k++) s = fma(row[k], T.avElem[(size_t)e * T.NB + k], s);
  return s;
}
// eps * grad(U_conserved) . n with the VOLUME gradient (calculateArtificialViscousNormalFlux, ViscousFlux.cpp:126-136)
// epsRow: the face row the viscosity is taken at -- for the right side of an interior face the reference passes the LEFT loop index j
// (right_quadrature_node_artificial_viscosity(j), SpatialDiscrete.cpp:714-719), not the matching point its gradient column uses
__device__ __forceinline__ void mxAvNormalFlux(const MxType& T, int e, int row, int epsRow, const double* n, double* out) {
  double g[kG];
  for (int r = 0; r < kG; r++) g[r] = 0.0;
  const double* C = T.Gvol + (size_t)e * T.Nb * kG;
  const double* phi = T.PhiF + (size_t)row * T.Nb;
  for (int b = 0; b < T.Nb; b++) { const double ph = phi[b]; for (int r = 0; r < kG; r++) g[r] = fma(C[b * kG + r], ph, g[r]); }
  const double eps = mxEps(T, T.NodalF + (size_t)epsRow * T.NB, e);
  for (int v = 0; v < kNV; v++) out[v] = eps * (g[v * kD] * n[0] + g[v * kD + 1] * n[1]);
}

template <int PASS>
__global__ void __launch_bounds__(128) mxFaceKernel(const __grid_constant__ Args A) {
  const Phys<0> ph(A.phys);
  const MxFaces& F = A.F;
  const int Nqf = F.Nqf, total = (F.nInt + F.nBnd) * Nqf;
  const bool ns = A.phys.ns != 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int face = i / Nqf, j = i - face * Nqf;
    const bool interior = face < F.nInt;
    const MxType& TL = A.t[slotOf(F.lt[face])];
    const int eL = F.le[face], fL = F.lf[face], rowL = fL * Nqf + j;
    const double n[2] = {F.nrm[(size_t)i * 2], F.nrm[(size_t)i * 2 + 1]};
    const double w = F.fjw[i];
    double consL[kNV], compL[kD + 3];
    traceOf(TL.U + (size_t)eL * TL.Nb * kNV, TL.PhiF + (size_t)rowL * TL.Nb, TL.Nb, consL);
    const double irL = compFromCons<kD>(ph, consL, compL);
    if (interior) {
      const MxType& TR = A.t[slotOf(F.rt[face])];
      const int eR = F.re[face], fR = F.rf[face], rowR = fR * Nqf + (Nqf - 1 - j);   // line faces: reversal, SimulationControl.cpp:387-403
      double consR[kNV], compR[kD + 3];
      traceOf(TR.U + (size_t)eR * TR.Nb * kNV, TR.PhiF + (size_t)rowR * TR.Nb, TR.Nb, consR);
      const double irR = compFromCons<kD>(ph, consR, compR);
      if constexpr (PASS == 0) {
        double* aL = TL.AGv + ((size_t)eL * TL.Naq + rowL) * kG; double* aR = TR.AGv + ((size_t)eR * TR.Naq + rowR) * kG;
        double* bL = TL.AGi + ((size_t)eL * TL.Naq + rowL) * kG; double* bR = TR.AGi + ((size_t)eR * TR.Naq + rowR) * kG;
        for (int v = 0; v < kNV; v++) for (int c = 0; c < kD; c++) {
          const double t = n[c] * (consL[v] + consR[v]) / 2.0 * w;      // calculateVolumeGardientFlux, ViscousFlux.cpp:33-43
          aL[v * kD + c] = t; aR[v * kD + c] = -t;
          const double s = n[c] * (consR[v] - consL[v]) / 2.0 * w;      // calculateInterfaceGardientFlux, :46-56 (same sign both sides)
          bL[v * kD + c] = s; bR[v * kD + c] = s;
        }
      } else {
        double Fc[kNV];
        convFlux<kD>(ph, n, consL, compL, irL, consR, compR, irR, Fc);
        if (ns) {  // calculateViscousFlux, ViscousFlux.cpp:139-153
          double g[kG], gp[kG], a[kNV], b[kNV];
          gradTraceOf(TL, A.phys.visc, eL, fL, TL.PhiF + (size_t)rowL * TL.Nb, g);
          primGradFromConsGrad<kD>(ph, consL, compL, g, gp); viscNormalFlux<kD>(ph, n, compL, gp, a);
          gradTraceOf(TR, A.phys.visc, eR, fR, TR.PhiF + (size_t)rowR * TR.Nb, g);
          primGradFromConsGrad<kD>(ph, consR, compR, g, gp); viscNormalFlux<kD>(ph, n, compR, gp, b);
          for (int v = 0; v < kNV; v++) Fc[v] -= (a[v] + b[v]) / 2.0;
        }
        if (A.phys.av) {  // calculateArtificialViscousFlux, ViscousFlux.cpp:172-186
          double a[kNV], b[kNV];
          mxAvNormalFlux(TL, eL, rowL, rowL, n, a); mxAvNormalFlux(TR, eR, rowR, fR * Nqf + j, n, b);
          for (int v = 0; v < kNV; v++) Fc[v] -= (a[v] + b[v]) / 2.0;
        }
        double* aL = TL.A + ((size_t)eL * TL.Naq + rowL) * kNV; double* aR = TR.A + ((size_t)eR * TR.Naq + rowR) * kNV;
        for (int v = 0; v < kNV; v++) { aL[v] = Fc[v] * w; aR[v] = -Fc[v] * w; }
      }
    } else {
      const int bc = F.bc[face];
      const double* dm = F.dummy + (size_t)(i - F.nInt * Nqf) * (kD + 3);
      double R[kD + 3];
      for (int k = 0; k < kD + 3; k++) R[k] = dm[k];
      if constexpr (PASS == 0) {
        double vol[kNV], itf[kNV];
        bcBoundaryGradientVariable<kD>(ph, bc, n, consL, compL, R, vol, itf);
        double* aL = TL.AGv + ((size_t)eL * TL.Naq + rowL) * kG; double* bL = TL.AGi + ((size_t)eL * TL.Naq + rowL) * kG;
        for (int v = 0; v < kNV; v++) for (int c = 0; c < kD; c++) { aL[v * kD + c] = n[c] * vol[v] * w; bL[v * kD + c] = n[c] * itf[v] * w; }
      } else {
        double b[kD + 3], Fc[kNV];
        bcBoundaryVariable<kD>(ph, bc, n, compL, R, b);
        convNormalFlux<kD>(ph, n, b, Fc);   // SpatialDiscrete.cpp:802-803: normal flux of the boundary state, no Riemann solve
        if (ns) {  // modifyBoundaryVariable (BoundaryCondition.cpp:299-307,443-452,490-501,535-546) + averaged viscous flux
          double g[kG], gp[kG], gb[kG], a[kNV], c2[kNV], cl[kD + 3];
          gradTraceOf(TL, A.phys.visc, eL, fL, TL.PhiF + (size_t)rowL * TL.Nb, g);
          primGradFromConsGrad<kD>(ph, consL, compL, g, gp);
          const bool wall = bcIsWall(bc);
          for (int k = 0; k < kD + 3; k++) cl[k] = wall ? b[k] : compL[k];
          for (int k = 0; k < kG; k++) gb[k] = gp[k];
          if (bc == kAdiabaticSlipWall || bc == kAdiabaticNonSlipWall) for (int d = 0; d < kD; d++) gb[(kD + 1) * kD + d] = 0.0;
          viscNormalFlux<kD>(ph, n, cl, gp, a); viscNormalFlux<kD>(ph, n, b, gb, c2);
          for (int v = 0; v < kNV; v++) Fc[v] -= (a[v] + c2[v]) / 2.0;
        }
        if (A.phys.av) {  // boundary faces: the interior side alone (SpatialDiscrete.cpp:813-819)
          double a[kNV];
          mxAvNormalFlux(TL, eL, rowL, rowL, n, a);
          for (int v = 0; v < kNV; v++) Fc[v] -= a[v];
        }
        double* aL = TL.A + ((size_t)eL * TL.Naq + rowL) * kNV;
        for (int v = 0; v < kNV; v++) aL[v] = Fc[v] * w;
      }
    }
  }
}

// G1 + G3 + G4: one WARP per element, kElemThreads / 32 elements per thread block (the per-element tiles are a few hundred
// doubles: a whole thread block per element left most of its threads idle)
__global__ void __launch_bounds__(kElemThreads) mxGradElemKernel(const __grid_constant__ Args A, int slot) {
  const MxType& T = A.t[slot];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int e = blockIdx.x * (kElemThreads / 32) + wib, Nb = T.Nb, Nq = T.Nq, Naq = T.Naq;
  if (e >= T.n) return;   // whole warp
  extern __shared__ double sm[];
  double* sU = sm + (size_t)wib * (Nb * kNV + Nq * kNV + Nb * kG); double* sUq = sU + Nb * kNV; double* sRg = sUq + Nq * kNV;
  for (int k = lane; k < Nb * kNV; k += 32) sU[k] = T.U[(size_t)e * Nb * kNV + k];
  __syncwarp();
  for (int k = lane; k < Nq * kNV; k += 32) {
    const int q = k / kNV, v = k - q * kNV;
    double s = 0.0;
    for (int b = 0; b < Nb; b++) s = fma(sU[b * kNV + v], T.Phi[q * Nb + b], s);
    sUq[k] = s;
  }
  __syncwarp();
  const int nOut = Nb * kG;   // <= 128: at most 4 outputs per lane
  const double* Mi = T.Minv + (size_t)e * Nb * Nb;
  const double* av = T.AGv + (size_t)e * Naq * kG;
  const double* mt = T.mt + (size_t)e * Nq * 4;
  // Rg_vol = A_vol Phi_f - Q_vol grad Phi, SpatialDiscrete.cpp:1034-1068
  for (int k = lane; k < nOut; k += 32) {
    const int b0 = k / kG, r = k - b0 * kG, v = r / kD, c = r - v * kD;
    double s = 0.0;
    for (int aq = 0; aq < Naq; aq++) s = fma(av[aq * kG + r], T.PhiF[aq * Nb + b0], s);
    for (int q = 0; q < Nq; q++) for (int dd = 0; dd < kD; dd++) s = fma(-sUq[q * kNV + v] * mt[q * 4 + dd * kD + c], T.dPhi[(q * 2 + dd) * Nb + b0], s);
    sRg[k] = s;
  }
  __syncwarp();
  double tot[4] = {0.0, 0.0, 0.0, 0.0};
  for (int k = lane, m = 0; k < nOut; k += 32, m++) {  // G_vol = Rg_vol M^-1, TimeIntegration.cpp:200-228
    const int b0 = k / kG, r = k - b0 * kG;
    double s = 0.0;
    for (int b = 0; b < Nb; b++) s = fma(sRg[b * kG + r], Mi[b * Nb + b0], s);
    T.Gvol[(size_t)e * Nb * kG + k] = s; tot[m] = s;
  }
  const bool br2 = A.phys.visc == kBR2;
  const int nl = br2 ? T.Nf : 1;
  const double* ai = T.AGi + (size_t)e * Naq * kG;
  for (int f = 0; f < nl; f++) {
    __syncwarp();
    const int lo = br2 ? f * T.Nqf : 0, hi = br2 ? lo + T.Nqf : Naq;   // BR1: A_int Phi_f; BR2: per face A_int[:, f] Phi_f[f, :]
    for (int k = lane; k < nOut; k += 32) {
      const int b0 = k / kG, r = k - b0 * kG;
      double s = 0.0;
      for (int aq = lo; aq < hi; aq++) s = fma(ai[aq * kG + r], T.PhiF[aq * Nb + b0], s);
      sRg[k] = s;
    }
    __syncwarp();
    for (int k = lane, m = 0; k < nOut; k += 32, m++) {
      const int b0 = k / kG, r = k - b0 * kG;
      double s = 0.0;
      for (int b = 0; b < Nb; b++) s = fma(sRg[b * kG + r], Mi[b * Nb + b0], s);
      if (br2) T.Gf[((size_t)e * T.Nf + f) * Nb * kG + k] = s;
      tot[m] += s;
    }
  }
  for (int k = lane, m = 0; k < nOut; k += 32, m++) T.Gtot[(size_t)e * Nb * kG + k] = tot[m];
}

// R1 + R3 + R4 + K: one warp per element, kElemThreads / 32 elements per thread block
__global__ void __launch_bounds__(kElemThreads) mxElemKernel(const __grid_constant__ Args A, int slot) {
  const MxType& T = A.t[slot];
  const Phys<0> ph(A.phys);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int e = blockIdx.x * (kElemThreads / 32) + wib, Nb = T.Nb, Nq = T.Nq, Naq = T.Naq;
  if (e >= T.n) return;   // whole warp
  const bool ns = A.phys.ns != 0, src = A.phys.source != kSourceNone, av = A.phys.av != 0;
  extern __shared__ double sm[];
  const int per = Nb * kNV + Nb * kG + Nq * kD * kNV + Nq * kNV + Nb * kNV;
  double* sU = sm + (size_t)wib * per; double* sG = sU + Nb * kNV; double* sQ = sG + Nb * kG; double* sS = sQ + Nq * kD * kNV; double* sR = sS + Nq * kNV;
  for (int k = lane; k < Nb * kNV; k += 32) sU[k] = T.U[(size_t)e * Nb * kNV + k];
  if (ns) for (int k = lane; k < Nb * kG; k += 32) sG[k] = T.Gtot[(size_t)e * Nb * kG + k];
  if (av) for (int k = lane; k < Nb * kG; k += 32) sG[k] = T.Gvol[(size_t)e * Nb * kG + k];   // Euler + artificial viscosity: the volume gradient
  __syncwarp();
  for (int q = lane; q < Nq; q += 32) {  // calculateElementQuadrature, SpatialDiscrete.cpp:194-266
    double cons[kNV] = {0, 0, 0, 0}, comp[kD + 3], Fc[kD * kNV];
    for (int b = 0; b < Nb; b++) { const double f = T.Phi[q * Nb + b]; for (int v = 0; v < kNV; v++) cons[v] = fma(sU[b * kNV + v], f, cons[v]); }
    compFromCons<kD>(ph, cons, comp);
    convRawFlux<kD>(ph, comp, Fc);
    if (ns) {
      double g[kG], gp[kG], Fv[kD * kNV];
      for (int r = 0; r < kG; r++) g[r] = 0.0;
      for (int b = 0; b < Nb; b++) { const double f = T.Phi[q * Nb + b]; for (int r = 0; r < kG; r++) g[r] = fma(sG[b * kG + r], f, g[r]); }
      primGradFromConsGrad<kD>(ph, cons, comp, g, gp);
      viscRawFlux<kD>(ph, comp, gp, Fv);
      for (int k = 0; k < kD * kNV; k++) Fc[k] -= Fv[k];
    }
    if (av) {  // calculateArtificialViscousRawFlux, ViscousFlux.cpp:105-113
      double g[kG];
      for (int r = 0; r < kG; r++) g[r] = 0.0;
      for (int b = 0; b < Nb; b++) { const double f = T.Phi[q * Nb + b]; for (int r = 0; r < kG; r++) g[r] = fma(sG[b * kG + r], f, g[r]); }
      const double eps = mxEps(T, T.NodalQ + (size_t)q * T.NB, e);
      for (int k = 0; k < kD * kNV; k++) Fc[k] -= eps * g[k];
    }
    const double* mt = T.mt + ((size_t)e * Nq + q) * 4;
    for (int dd = 0; dd < kD; dd++) for (int k = 0; k < kNV; k++) {
      double t = 0.0;
      for (int c = 0; c < kD; c++) t = fma(Fc[k * kD + c], mt[dd * kD + c], t);
      sQ[(q * kD + dd) * kNV + k] = t;
    }
    if (src) {
      const double jw = T.jw[(size_t)e * Nq + q];
      for (int k = 0; k < kNV; k++) sS[q * kNV + k] = 0.0;
      sS[q * kNV + kD] = boussinesqSource<kD>(ph, comp) * jw;   // SourceTerm.cpp:29-58, SpatialDiscrete.cpp:254-262
    }
  }
  __syncwarp();
  const double* a = T.A + (size_t)e * Naq * kNV;
  for (int k = lane; k < Nb * kNV; k += 32) {  // calculateElementResidual, SpatialDiscrete.cpp:1016-1032
    const int b = k / kNV, v = k - b * kNV;
    double s = 0.0;
    for (int qd = 0; qd < Nq * kD; qd++) s = fma(sQ[qd * kNV + v], T.dPhi[qd * Nb + b], s);
    for (int aq = 0; aq < Naq; aq++) s = fma(-a[aq * kNV + v], T.PhiF[aq * Nb + b], s);
    if (src) for (int q = 0; q < Nq; q++) s = fma(sS[q * kNV + v], T.Phi[q * Nb + b], s);
    sR[k] = s;
    if (A.mode == 1) T.R[(size_t)e * Nb * kNV + k] = s;
  }
  __syncwarp();
  const double* Mi = T.Minv + (size_t)e * Nb * Nb;
  for (int k = lane; k < Nb * kNV; k += 32) {  // updateElementBasisFunctionCoefficient, TimeIntegration.cpp:181-198
    const int b0 = k / kNV, v = k - b0 * kNV;
    double s = 0.0;
    for (int b = 0; b < Nb; b++) s = fma(sR[b * kNV + v], Mi[b * Nb + b0], s);
    const size_t at = (size_t)e * Nb * kNV + k;
    if (A.mode == 1) T.RM[at] = s;
    else T.U[at] = A.aCur * sU[k] + A.aLast * T.Ulast[at] + A.bdt * s;
  }
  if (A.normPartial) {  // calculateElementRelativeError, TimeIntegration.cpp:279-298: mean_q |R Phi^T|
    for (int k = lane; k < Nq * kNV; k += 32) {
      const int q = k / kNV, v = k - q * kNV;
      double s = 0.0;
      for (int b = 0; b < Nb; b++) s = fma(sR[b * kNV + v], T.Phi[q * Nb + b], s);
      sQ[k] = fabs(s);
    }
    __syncwarp();
    if (lane < kNV) {
      double s = 0.0;
      for (int q = 0; q < Nq; q++) s += sQ[q * kNV + lane];
      A.normPartial[(size_t)(T.normOff + e) * kNV + lane] = s / Nq;
    }
  }
}

// calculateElementArtificialViscosity (SpatialDiscrete.cpp:37-87) in the modal representation: one warp per element
__global__ void mxAvIndicatorKernel(const __grid_constant__ Args A, int slot, int p, double tol, double empTol, double factor) {
  const MxType& T = A.t[slot];
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nWarps = (gridDim.x * blockDim.x) >> 5;
  for (int e = warp; e < T.n; e += nWarps) {
    const double* U = T.U + (size_t)e * T.Nb * kNV;
    double num = 0.0, den = 0.0;
    for (int q = lane; q < T.Nq; q += 32) {
      double all = 0.0, high = 0.0;
      for (int b = 0; b < T.Nb; b++) { const double t = T.Phi[q * T.Nb + b] * U[b * kNV]; all += t; if (b >= T.nbLow) high += t; }
      const double w = T.jw[(size_t)e * T.Nq + q];
      num += high * (high * w); den += all * (all * w);
    }
    for (int o = 16; o > 0; o >>= 1) { num += __shfl_xor_sync(0xffffffffu, num, o); den += __shfl_xor_sync(0xffffffffu, den, o); }
    if (lane == 0) {
      const double shock = log10(num / den);
      const double full = factor * (T.radius[e] / p);
      double val;
      if (shock < tol - empTol) val = 0.0;
      else if (shock > tol + empTol) val = full;
      else val = full * (1.0 + sin(3.14159265358979323846 * (shock - tol) / (2.0 * empTol))) / 2.0;
      T.avE[e] = val;
    }
  }
}

// out[q][c] = sum_b C[e][b][c] Phi[q][b]  (C = U with ncol 4, or a gradient field with ncol 8)
__global__ void mxToQuadratureKernel(const double* __restrict__ C, const double* __restrict__ Phi, int n, int Nb, int Nq, int ncol, double* __restrict__ out) {
  const size_t total = (size_t)n * Nq * ncol;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % ncol); const size_t eq = i / ncol; const int q = (int)(eq % Nq); const size_t e = eq / Nq;
    double s = 0.0;
    for (int b = 0; b < Nb; b++) s = fma(C[(e * Nb + b) * ncol + c], Phi[q * Nb + b], s);
    out[i] = s;
  }
}

// Solver::initializeSolver, InitialCondition.cpp:85-116: primitive at the quadrature points -> conserved -> U = Uq Phi (Phi^T Phi)^-1
__global__ void mxProjectKernel(const double* __restrict__ prim, const double* __restrict__ Proj, int n, int Nb, int Nq, PhysParams P, double* __restrict__ U) {
  const size_t total = (size_t)n * Nb;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i / Nb; const int b = (int)(i - e * Nb);
    double acc[kNV] = {0, 0, 0, 0};
    for (int q = 0; q < Nq; q++) {
      const double* p = prim + (e * Nq + q) * kNV;
      const double rho = p[0], v2 = p[1] * p[1] + p[2] * p[2], ei = P.cv * p[kD + 1];
      const double cons[kNV] = {rho, rho * p[1], rho * p[2], P.compressible ? rho * (ei + 0.5 * v2) : rho * ei};   // VariableConvertor.cpp:341-366
      const double f = Proj[b * Nq + q];
      for (int v = 0; v < kNV; v++) acc[v] = fma(cons[v], f, acc[v]);
    }
    for (int v = 0; v < kNV; v++) U[i * kNV + v] = acc[v];
  }
}

// boundary_dummy_variable_ (InitialCondition.cpp:118-149): primitive [nBnd][Nqf][NV] -> computational [nBnd][Nqf][D+3]
__global__ void mxBoundaryKernel(const double* __restrict__ prim, int total, PhysParams P, double* __restrict__ dummy) {
  const Phys<0> ph(P);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    double pr[kNV], comp[kD + 3];
    for (int k = 0; k < kNV; k++) pr[k] = prim[(size_t)i * kNV + k];
    compFromPrim<kD>(ph, pr, comp);
    for (int k = 0; k < kD + 3; k++) dummy[(size_t)i * (kD + 3) + k] = comp[k];
  }
}

// calculateElementDeltaTime, TimeIntegration.cpp:104-131
__global__ void mxDeltaTimeKernel(const double* __restrict__ U, const double* __restrict__ Phi, const double* __restrict__ minEdge, int n, int Nb, int Nq, int p,
                                  double cfl, PhysParams P, double* __restrict__ partial) {
  const Phys<0> ph(P);
  double best = 1.7976931348623157e308;
  const size_t total = (size_t)n * Nq;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t e = i / Nq; const int q = (int)(i - e * Nq);
    double cons[kNV] = {0, 0, 0, 0}, comp[kD + 3];
    for (int b = 0; b < Nb; b++) { const double f = Phi[q * Nb + b]; for (int v = 0; v < kNV; v++) cons[v] = fma(U[(e * Nb + b) * kNV + v], f, cons[v]); }
    compFromCons<kD>(ph, cons, comp);
    const double sr = sqrt(vsq<kD>(comp)) + ph.sound(comp[0], comp[kD + 2]);
    const double dt = cfl * minEdge[e] / (sr * (p + 1.0) * (p + 1.0));
    best = dt < best ? dt : best;
  }
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_down_sync(0xffffffffu, best, o); best = t < best ? t : best; }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) { for (int w = 1; w < (int)(blockDim.x >> 5); w++) best = red[w] < best ? red[w] : best; partial[blockIdx.x] = best; }
}

// deterministic sum over elements of the per-element norm partials: out[v] = sum_e partial[e][v]
__global__ void mxNormReduceKernel(const double* __restrict__ partial, int n, double* __restrict__ out) {
  __shared__ double red[32];
  const int v = blockIdx.x;
  double s = 0.0;
  for (int c = threadIdx.x; c < n; c += blockDim.x) s += partial[(size_t)c * kNV + v];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w]; out[v] = t; }
}

int elemBlocks(int n) { return (n + kElemThreads / 32 - 1) / (kElemThreads / 32); }
size_t elemSmemBytes(const MixedTable& T) { return (kElemThreads / 32) * sizeof(double) * ((size_t)T.Nb * kNV + (size_t)T.Nb * kG + (size_t)T.Nq * kD * kNV + (size_t)T.Nq * kNV + (size_t)T.Nb * kNV); }
size_t gradSmemBytes(const MixedTable& T) { return (kElemThreads / 32) * sizeof(double) * ((size_t)T.Nb * kNV + (size_t)T.Nq * kNV + (size_t)T.Nb * kG); }

}  // namespace

// ---- host side ---------------------------------------------------------------------------------------------------------------
MixedSolver::MixedSolver(int p, const PhysParams& phys, int nStages, const double (*rkc)[3], cudaStream_t stream, bool hasDevice, int device)
    : p_(p), nStages_(nStages), device_(device), phys_(phys), stream_(stream), hasDevice_(hasDevice) {
  std::memcpy(rkc_, rkc, sizeof(rkc_));
  if (p < 1 || p > 3) throw std::runtime_error("dense-operator path: polynomial order 1..3");
}

void MixedSolver::needDevice() const { if (!hasDevice_) throw std::runtime_error("no CUDA device bound to this context (plan-only context); the product has no CPU path"); }
MixedBlock& MixedSolver::block(int type) { if (!hasType(type)) throw std::runtime_error("no element block of this type"); return *blk_[type]; }

void MixedSolver::addBlock(int type, int n, int nGhost, int g, const double* coords) {
  if (type != kTriangle && type != kQuadrangle) throw std::runtime_error("dense-operator path: triangle and quadrangle blocks only (tetrahedra / pyramids are not implemented)");
  if (nGhost != 0) throw std::runtime_error("dense-operator path runs on one GPU (no ghost elements); partitioned runs need single-type quadrangle / hexahedron meshes");
  if (blk_[type]) throw std::runtime_error("element block of this type already set");
  auto B = std::make_unique<MixedBlock>();
  B->type = type; B->n = n; B->g = g;
  B->T.build(type, p_, g);
  B->X.assign(coords, coords + (size_t)n * B->T.nn * kD);
  blk_[type] = std::move(B);
}

static void elementGeometry(MixedBlock& B) {
  const MixedTable& T = B.T;
  const int Nq = T.Nq, nn = T.nn, Nb = T.Nb, n = B.n;
  B.xq.assign((size_t)n * Nq * 2, 0.0); B.jw.assign((size_t)n * Nq, 0.0); B.mt.assign((size_t)n * Nq * 4, 0.0);
  B.Minv.assign((size_t)n * Nb * Nb, 0.0); B.minEdge.assign(n, 0.0);
  bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
  for (int e = 0; e < n; e++) {
    const double* X = &B.X[(size_t)e * nn * 2];
    for (int q = 0; q < Nq; q++) {  // getElementJacobian, Geometry.cpp:44-67: Jt(k,l) = d x_l / d xi_k
      double Jt[4] = {0, 0, 0, 0}, inv[4], x[2] = {0, 0};
      for (int m = 0; m < nn; m++) {
        const double N = T.GN[(size_t)q * nn + m];
        for (int l = 0; l < 2; l++) x[l] += N * X[m * 2 + l];
        for (int k = 0; k < 2; k++) { const double dN = T.dGN[((size_t)q * 2 + k) * nn + m]; for (int l = 0; l < 2; l++) Jt[k * 2 + l] += dN * X[m * 2 + l]; }
      }
      const double det = invertSmall(2, Jt, inv);
      if (!(det > 0.0)) bad = true;
      const double w = det * T.wq[q];
      B.xq[((size_t)e * Nq + q) * 2] = x[0]; B.xq[((size_t)e * Nq + q) * 2 + 1] = x[1];
      B.jw[(size_t)e * Nq + q] = w;
      double* mt = &B.mt[((size_t)e * Nq + q) * 4];   // (Jt^-1).reshaped() * detJ * w: index c + 2 d  <-  inv(c, d)
      for (int c = 0; c < 2; c++) for (int dd = 0; dd < 2; dd++) mt[dd * 2 + c] = inv[c * 2 + dd] * w;
    }
    // calculateElementLocalMassMatrixInverse, Geometry.cpp:88-100: M = Phi^T diag(detJ w) Phi
    std::vector<long double> M((size_t)Nb * Nb), I;
    for (int a = 0; a < Nb; a++) for (int b = a; b < Nb; b++) {
      long double s = 0;
      for (int q = 0; q < Nq; q++) s += (long double)T.Phi[(size_t)q * Nb + a] * (long double)B.jw[(size_t)e * Nq + q] * (long double)T.Phi[(size_t)q * Nb + b];
      M[(size_t)a * Nb + b] = s; M[(size_t)b * Nb + a] = s;
    }
    MixedTable::invertLong(M, I, Nb);
    for (int k = 0; k < Nb * Nb; k++) B.Minv[(size_t)e * Nb * Nb + k] = (double)I[k];
    // getElementQuality "minEdge", Geometry.cpp:29-42
    const int nc = T.type == kTriangle ? 3 : 4;
    double me = 1e300;
    for (int k = 0; k < nc; k++) { const int a = k, b = (k + 1) % nc; const double dx = X[a * 2] - X[b * 2], dy = X[a * 2 + 1] - X[b * 2 + 1]; me = std::min(me, std::sqrt(dx * dx + dy * dy)); }
    B.minEdge[e] = me;
  }
  if (bad) throw std::runtime_error("non-positive Jacobian determinant");
}

// getAdjacencyElementJacobian + calculateNormalVector, Geometry.cpp:69-86,114-129: from the LEFT parent's map
void MixedSolver::faceGeometry() {
  const int nf = F_.nInt + F_.nBnd, Nqf = p_ + 1;
  xf_.assign((size_t)nf * Nqf * 2, 0.0); nrm_.assign((size_t)nf * Nqf * 2, 0.0); fjw_.assign((size_t)nf * Nqf, 0.0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nf; i++) {
    const MixedBlock& B = *blk_[F_.lt[i]]; const MixedTable& T = B.T;
    const int e = F_.le[i], f = F_.lf[i], nn = T.nn;
    const double* X = &B.X[(size_t)e * nn * 2];
    for (int j = 0; j < Nqf; j++) {
      const int row = f * Nqf + j;
      double Jt[4] = {0, 0, 0, 0}, x[2] = {0, 0};
      for (int m = 0; m < nn; m++) {
        const double N = T.GNf[(size_t)row * nn + m];
        for (int l = 0; l < 2; l++) x[l] += N * X[m * 2 + l];
        for (int k = 0; k < 2; k++) { const double dN = T.dGNf[((size_t)row * 2 + k) * nn + m]; for (int l = 0; l < 2; l++) Jt[k * 2 + l] += dN * X[m * 2 + l]; }
      }
      double t[2] = {0, 0};
      for (int k = 0; k < 2; k++) for (int l = 0; l < 2; l++) t[l] += T.ftan[(size_t)f * 2 + k] * Jt[k * 2 + l];
      const double scale = std::sqrt(t[0] * t[0] + t[1] * t[1]);
      const size_t at = (size_t)i * Nqf + j;
      xf_[at * 2] = x[0]; xf_[at * 2 + 1] = x[1];
      nrm_[at * 2] = t[1] / scale; nrm_[at * 2 + 1] = -t[0] / scale;
      fjw_[at] = scale * T.wf[j];
    }
  }
}

void MixedSolver::finalize() {
  if (finalized_) throw std::runtime_error("context already finalized");
  const int nf = F_.nInt + F_.nBnd;
  for (int i = 0; i < nf; i++) for (int s = 0; s < (i < F_.nInt ? 2 : 1); s++) {
    const int t = s ? F_.rt[i] : F_.lt[i], e = s ? F_.re[i] : F_.le[i], f = s ? F_.rf[i] : F_.lf[i];
    if (!hasType(t) || e < 0 || e >= blk_[t]->n || f < 0 || f >= blk_[t]->T.Nf) throw std::runtime_error("face record out of range");
    if (i >= F_.nInt && (F_.bc[i] < 0 || F_.bc[i] > 5)) throw std::runtime_error("boundary face without a boundary condition (Periodic is not a boundary type)");
  }
  for (auto& b : blk_) if (b) elementGeometry(*b);
  faceGeometry();
  if (hasDevice_) {
    CUDA_OK(cudaSetDevice(device_));
    const bool ns = phys_.ns != 0;
    for (auto& bp : blk_) if (bp) {
      MixedBlock& B = *bp; const MixedTable& T = B.T;
      B.dPhi_.upload(T.Phi, stream_); B.dDPhi.upload(T.dPhi, stream_); B.dPhiF.upload(T.PhiF, stream_); B.dProj.upload(T.Proj, stream_);
      B.dMt.upload(B.mt, stream_); B.dJw.upload(B.jw, stream_); B.dMinv.upload(B.Minv, stream_); B.dMinEdge.upload(B.minEdge, stream_);
      const size_t ns4 = (size_t)B.n * T.Nb * kNV, ng = (size_t)B.n * T.Nb * kG;
      B.U.alloc(ns4); B.Ulast.alloc(ns4); B.R.alloc(ns4); B.RM.alloc(ns4); B.A.alloc((size_t)B.n * T.Naq * kNV);
      B.U.zero(stream_); B.Ulast.zero(stream_); B.R.zero(stream_); B.RM.zero(stream_); B.A.zero(stream_);
      if (ns || phys_.av) {
        B.AGv.alloc((size_t)B.n * T.Naq * kG); B.AGi.alloc((size_t)B.n * T.Naq * kG); B.Gvol.alloc(ng); B.Gtot.alloc(ng);
        B.AGv.zero(stream_); B.AGi.zero(stream_); B.Gvol.zero(stream_); B.Gtot.zero(stream_);
        if (phys_.visc == kBR2) { B.Gf.alloc(ng * T.Nf); B.Gf.zero(stream_); }
      }
      if (phys_.av) {
        if ((int)B.tags.size() != B.n * T.nbasic || (int)B.radius.size() != B.n) throw std::runtime_error("artificial viscosity needs sdg_set_element_nodes for every block before sdg_finalize");
        B.dTags.upload(B.tags, stream_); B.dRadius.upload(B.radius, stream_);
        B.dNodalQ.upload(T.NodalQ, stream_); B.dNodalF.upload(T.NodalF, stream_);
        B.dAvElem.alloc((size_t)B.n * T.nbasic); B.dAvElem.zero(stream_); B.dAvE.alloc((size_t)B.n); B.dAvE.zero(stream_);
      }
      const size_t sm = std::max(elemSmemBytes(T), gradSmemBytes(T));
      if (sm > 48 * 1024) throw std::runtime_error("internal: element tile exceeds the default shared-memory window");
    }
    dLe.upload(F_.le, stream_); dLt.upload(F_.lt, stream_); dLf.upload(F_.lf, stream_); dRe.upload(F_.re, stream_); dRt.upload(F_.rt, stream_);
    dRf.upload(F_.rf, stream_); dBc.upload(F_.bc, stream_);
    dNrm.upload(nrm_, stream_); dFjw.upload(fjw_, stream_);
    dDummy.alloc((size_t)std::max(F_.nBnd, 1) * (p_ + 1) * (kD + 3)); dDummy.zero(stream_);
    normPartial.alloc((size_t)totalElements() * kNV); normPartial.zero(stream_); normOut.alloc(8); dtPartial.alloc(1024);
    if (phys_.av) { avNode_.alloc((size_t)std::max(avNodes_, 1)); avNode_.zero(stream_); }
    CUDA_OK(cudaStreamSynchronize(stream_));
  }
  finalized_ = true;
}

void MixedSolver::sizes(int type, int32_t* out) const {
  if (!hasType(type)) throw std::runtime_error("no element block of this type");
  const MixedTable& T = blk_[type]->T;
  out[0] = blk_[type]->n; out[1] = T.Nb; out[2] = T.Nq; out[3] = T.Nf; out[4] = T.Naq; out[5] = T.nn; out[6] = T.Nqf; out[7] = kNV;
}

void MixedSolver::quadratureCoordinates(int type, double* xq) const {
  if (!hasType(type)) throw std::runtime_error("no element block of this type");
  MixedBlock tmp; tmp.type = type; tmp.n = blk_[type]->n; tmp.g = blk_[type]->g; tmp.T = blk_[type]->T; tmp.X = blk_[type]->X;
  if (!blk_[type]->xq.empty()) { std::memcpy(xq, blk_[type]->xq.data(), blk_[type]->xq.size() * sizeof(double)); return; }
  elementGeometry(tmp);
  std::memcpy(xq, tmp.xq.data(), tmp.xq.size() * sizeof(double));
}

void MixedSolver::boundaryQuadratureCoordinates(double* xb) {
  if (xf_.empty()) faceGeometry();
  const int Nqf = p_ + 1;
  std::memcpy(xb, xf_.data() + (size_t)F_.nInt * Nqf * 2, (size_t)F_.nBnd * Nqf * 2 * sizeof(double));
}

void MixedSolver::fill(Args& a) {
  std::memset(&a, 0, sizeof(a));
  int off = 0;
  for (int type : {(int)kTriangle, (int)kQuadrangle}) {
    MxType& t = a.t[type == kTriangle ? 0 : 1];
    if (!blk_[type]) continue;
    MixedBlock& B = *blk_[type]; const MixedTable& T = B.T;
    t.n = B.n; t.Nb = T.Nb; t.Nq = T.Nq; t.Nf = T.Nf; t.Nqf = T.Nqf; t.Naq = T.Naq; t.normOff = off; off += B.n;
    t.Phi = B.dPhi_.p; t.dPhi = B.dDPhi.p; t.PhiF = B.dPhiF.p; t.Proj = B.dProj.p; t.mt = B.dMt.p; t.jw = B.dJw.p; t.Minv = B.dMinv.p; t.minEdge = B.dMinEdge.p;
    t.U = B.U.p; t.Ulast = B.Ulast.p; t.R = B.R.p; t.A = B.A.p; t.RM = B.RM.p;
    t.AGv = B.AGv.p; t.AGi = B.AGi.p; t.Gvol = B.Gvol.p; t.Gtot = B.Gtot.p; t.Gf = B.Gf.p;
    t.NB = T.nbasic; t.nbLow = p_ == 1 ? 0 : mixedNumBasis(type, p_ - 1);
    t.avElem = B.dAvElem.p; t.NodalQ = B.dNodalQ.p; t.NodalF = B.dNodalF.p; t.radius = B.dRadius.p; t.tags = B.dTags.p; t.avE = B.dAvE.p;
  }
  a.F.nInt = F_.nInt; a.F.nBnd = F_.nBnd; a.F.Nqf = p_ + 1;
  a.F.le = dLe.p; a.F.lt = dLt.p; a.F.lf = dLf.p; a.F.re = dRe.p; a.F.rt = dRt.p; a.F.rf = dRf.p; a.F.bc = dBc.p;
  a.F.nrm = dNrm.p; a.F.fjw = dFjw.p; a.F.dummy = dDummy.p;
  a.phys = phys_;
  a.aLast = 0.0; a.aCur = 1.0; a.bdt = 0.0; a.normPartial = nullptr; a.mode = 0;
}

// one residual evaluation (+ update when a.mode == 0)
void MixedSolver::evalResidual(Args& a, int mode, bool wantNorm) {
  a.mode = mode; a.normPartial = wantNorm ? normPartial.p : nullptr;
  const int nfp = (F_.nInt + F_.nBnd) * (p_ + 1);
  const int fb = std::max(1, (nfp + 127) / 128);
  if (phys_.ns || phys_.av) {
    mxFaceKernel<0><<<fb, 128, 0, stream_>>>(a); launches++;
    for (int type : {(int)kTriangle, (int)kQuadrangle}) if (blk_[type]) {
      mxGradElemKernel<<<elemBlocks(blk_[type]->n), kElemThreads, gradSmemBytes(blk_[type]->T), stream_>>>(a, type == kTriangle ? 0 : 1); launches++;
    }
  }
  mxFaceKernel<1><<<fb, 128, 0, stream_>>>(a); launches++;
  for (int type : {(int)kTriangle, (int)kQuadrangle}) if (blk_[type]) {
    mxElemKernel<<<elemBlocks(blk_[type]->n), kElemThreads, elemSmemBytes(blk_[type]->T), stream_>>>(a, type == kTriangle ? 0 : 1); launches++;
  }
  CUDA_OK(cudaGetLastError());
}

void MixedSolver::step(double dt, int nSteps, double* relErr, float* ms) {
  needDevice();
  CUDA_OK(cudaSetDevice(device_));
  Args a; fill(a);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ms) { CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1)); CUDA_OK(cudaStreamSynchronize(stream_)); CUDA_OK(cudaEventRecord(e0, stream_)); }
  auto oneStep = [&]() {
    if (nStages_ > 1) for (auto& b : blk_) if (b) CUDA_OK(cudaMemcpyAsync(b->Ulast.p, b->U.p, b->U.n * sizeof(double), cudaMemcpyDeviceToDevice, stream_));  // copyElementBasisFunctionCoefficient, TimeIntegration.cpp:70-78
    if (phys_.av) avUpdate();   // once per step, before the stages (TimeIntegration.cpp:336-338)
    for (int s = 0; s < nStages_; s++) {
      a.aLast = s == 0 ? 0.0 : rkc_[s][0]; a.aCur = s == 0 ? 1.0 : rkc_[s][1]; a.bdt = rkc_[s][2] * dt;
      evalResidual(a, 0, s == nStages_ - 1);
    }
  };
  // these meshes are launch bound (6 launches per stage): one step is captured as a CUDA graph and replayed
  int it = 0;
  if (nSteps >= 4 && !getenv("SDG_NO_GRAPH")) {
    if (!graphWarm_) { oneStep(); it++; graphWarm_ = true; }
    if (stepGraph_ == nullptr || graphDt_ != dt) {
      if (stepGraph_) { cudaGraphExecDestroy(stepGraph_); stepGraph_ = nullptr; }
      cudaGraph_t g = nullptr;
      const int64_t l0 = launches;
      CUDA_OK(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal));
      oneStep();
      CUDA_OK(cudaStreamEndCapture(stream_, &g));
      launchesPerStep_ = launches - l0; launches = l0;
      CUDA_OK(cudaGraphInstantiate(&stepGraph_, g, 0));
      cudaGraphDestroy(g);
      graphDt_ = dt;
    }
    for (; it < nSteps; it++) { CUDA_OK(cudaGraphLaunch(stepGraph_, stream_)); launches += launchesPerStep_; }
  }
  for (; it < nSteps; it++) oneStep();
  if (nSteps > 0) gradFromStep_ = true;
  if (ms) { CUDA_OK(cudaEventRecord(e1, stream_)); CUDA_OK(cudaEventSynchronize(e1)); CUDA_OK(cudaEventElapsedTime(ms, e0, e1)); cudaEventDestroy(e0); cudaEventDestroy(e1); }
  if (relErr) {
    const int ne = totalElements();
    // Solver::calculateRelativeError (TimeIntegration.cpp:300-324) hands the SAME vector to every element type and each
    // calculateElementRelativeError ASSIGNS its sum to it (:294-297): on a mixed mesh the value is the sum over the LAST element type
    // only (quadrangles), divided by the number of ALL elements.  Reproduced as is (pinned by tests/golden/reference_sweeps.json).
    const int lastType = blk_[kQuadrangle] ? kQuadrangle : kTriangle;
    const int off = (lastType == kQuadrangle && blk_[kTriangle]) ? blk_[kTriangle]->n : 0;
    mxNormReduceKernel<<<kNV, 256, 0, stream_>>>(normPartial.p + (size_t)off * kNV, blk_[lastType]->n, normOut.p); launches++;
    double h[kNV];
    CUDA_OK(cudaMemcpyAsync(h, normOut.p, sizeof(h), cudaMemcpyDeviceToHost, stream_));
    CUDA_OK(cudaStreamSynchronize(stream_));
    for (int v = 0; v < kNV; v++) relErr[v] = h[v] / ne;   // TimeIntegration.cpp:323
  } else CUDA_OK(cudaStreamSynchronize(stream_));
}

void MixedSolver::residual(int type, double* Rmodal, double* rhsq) {
  needDevice(); gradFromStep_ = false;
  CUDA_OK(cudaSetDevice(device_));
  MixedBlock& B = block(type); const MixedTable& T = B.T;
  Args a; fill(a);
  if (phys_.av) avUpdate();   // parity hook: the viscosity of the CURRENT state
  evalResidual(a, 1, false);
  const size_t nd = (size_t)B.n * T.Nb * kNV;
  if (Rmodal) CUDA_OK(cudaMemcpyAsync(Rmodal, B.R.p, nd * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  if (rhsq) {
    DevBuf<double> tmp; tmp.alloc((size_t)B.n * T.Nq * kNV);
    mxToQuadratureKernel<<<148 * 4, 256, 0, stream_>>>(B.RM.p, B.dPhi_.p, B.n, T.Nb, T.Nq, kNV, tmp.p); launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(rhsq, tmp.p, tmp.n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
    CUDA_OK(cudaStreamSynchronize(stream_));
  }
  CUDA_OK(cudaStreamSynchronize(stream_));
}

void MixedSolver::setStateFromPrimitive(int type, const double* prim) {
  needDevice(); gradFromStep_ = false;
  CUDA_OK(cudaSetDevice(device_));
  MixedBlock& B = block(type); const MixedTable& T = B.T;
  DevBuf<double> tmp; tmp.alloc((size_t)B.n * T.Nq * kNV);
  CUDA_OK(cudaMemcpyAsync(tmp.p, prim, tmp.n * sizeof(double), cudaMemcpyHostToDevice, stream_));
  const int blocks = (int)std::min<size_t>(((size_t)B.n * T.Nb + 127) / 128, 148 * 16);
  mxProjectKernel<<<blocks, 128, 0, stream_>>>(tmp.p, B.dProj.p, B.n, T.Nb, T.Nq, phys_, B.U.p); launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(stream_));
}

void MixedSolver::setBoundaryPrimitive(const double* prim) {
  needDevice();
  CUDA_OK(cudaSetDevice(device_));
  const int total = F_.nBnd * (p_ + 1);
  if (total == 0) return;
  DevBuf<double> tmp; tmp.alloc((size_t)total * kNV);
  CUDA_OK(cudaMemcpyAsync(tmp.p, prim, tmp.n * sizeof(double), cudaMemcpyHostToDevice, stream_));
  mxBoundaryKernel<<<(total + 127) / 128, 128, 0, stream_>>>(tmp.p, total, phys_, dDummy.p); launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(stream_));
}

void MixedSolver::setState(int type, const double* U) {
  needDevice(); gradFromStep_ = false; CUDA_OK(cudaSetDevice(device_));
  MixedBlock& B = block(type);
  CUDA_OK(cudaMemcpyAsync(B.U.p, U, B.U.n * sizeof(double), cudaMemcpyHostToDevice, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
}
void MixedSolver::getState(int type, double* U) {
  needDevice(); CUDA_OK(cudaSetDevice(device_));
  MixedBlock& B = block(type);
  CUDA_OK(cudaMemcpyAsync(U, B.U.p, B.U.n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
}
void MixedSolver::setStateDevice(int type, const void* U) {
  needDevice(); gradFromStep_ = false; CUDA_OK(cudaSetDevice(device_));
  MixedBlock& B = block(type);
  CUDA_OK(cudaMemcpyAsync(B.U.p, U, B.U.n * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
}
void MixedSolver::getStateDevice(int type, void* U) {
  needDevice(); CUDA_OK(cudaSetDevice(device_));
  MixedBlock& B = block(type);
  CUDA_OK(cudaMemcpyAsync(U, B.U.p, B.U.n * sizeof(double), cudaMemcpyDeviceToDevice, stream_));
}

void MixedSolver::stateAtQuadrature(int type, double* Uq) {
  needDevice(); CUDA_OK(cudaSetDevice(device_));
  MixedBlock& B = block(type); const MixedTable& T = B.T;
  DevBuf<double> tmp; tmp.alloc((size_t)B.n * T.Nq * kNV);
  mxToQuadratureKernel<<<148 * 4, 256, 0, stream_>>>(B.U.p, B.dPhi_.p, B.n, T.Nb, T.Nq, kNV, tmp.p); launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(Uq, tmp.p, tmp.n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
}

void MixedSolver::gradientAtQuadrature(int type, double* Gq) {
  needDevice(); CUDA_OK(cudaSetDevice(device_));
  if (!phys_.ns) throw std::runtime_error("gradient state exists for Navier-Stokes models only");
  MixedBlock& B = block(type); const MixedTable& T = B.T;
  Args a; fill(a);
  const int nfp = (F_.nInt + F_.nBnd) * (p_ + 1);
  mxFaceKernel<0><<<std::max(1, (nfp + 127) / 128), 128, 0, stream_>>>(a); launches++;
  for (int t : {(int)kTriangle, (int)kQuadrangle}) if (blk_[t]) { mxGradElemKernel<<<elemBlocks(blk_[t]->n), kElemThreads, gradSmemBytes(blk_[t]->T), stream_>>>(a, t == kTriangle ? 0 : 1); launches++; }
  DevBuf<double> tmp; tmp.alloc((size_t)B.n * T.Nq * kG);
  mxToQuadratureKernel<<<148 * 4, 256, 0, stream_>>>(B.Gtot.p, B.dPhi_.p, B.n, T.Nb, T.Nq, kG, tmp.p); launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(Gq, tmp.p, tmp.n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
}

// gradient coefficients of the CURRENT state in every block (Gvol, Gtot, Gf); after a step the getters keep what the last stage left (gradFromStep_)
void MixedSolver::refreshGradient() {
  if (!phys_.ns) throw std::runtime_error("gradient state exists for Navier-Stokes models only");
  Args a; fill(a);
  const int nfp = (F_.nInt + F_.nBnd) * (p_ + 1);
  mxFaceKernel<0><<<std::max(1, (nfp + 127) / 128), 128, 0, stream_>>>(a); launches++;
  for (int t : {(int)kTriangle, (int)kQuadrangle}) if (blk_[t]) { mxGradElemKernel<<<elemBlocks(blk_[t]->n), kElemThreads, gradSmemBytes(blk_[t]->T), stream_>>>(a, t == kTriangle ? 0 : 1); launches++; }
  CUDA_OK(cudaGetLastError());
}

// RawBinary.cpp:75-88: variable_gradient_basis_function_coefficient_, [n][Nb][Nv*D]
void MixedSolver::gradientState(int type, double* G) {
  needDevice(); CUDA_OK(cudaSetDevice(device_));
  if (!gradFromStep_) refreshGradient();
  MixedBlock& B = block(type);
  CUDA_OK(cudaMemcpyAsync(G, B.Gtot.p, B.Gtot.n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
}

// RawBinary.cpp:89-154: per boundary face the parent's gradient block (BR1 total; BR2 volume part + the lift of that face)
static __global__ void mxBoundaryGradientKernel(const double* __restrict__ base, const double* __restrict__ lift, const int* __restrict__ rec /* {e, f, offset} */,
                                                int nRec, int Nf, int len, double* __restrict__ out) {
  const int r = blockIdx.x;
  if (r >= nRec) return;
  const int e = rec[3 * r], f = rec[3 * r + 1]; double* dst = out + rec[3 * r + 2];
  const double* a = base + (size_t)e * len;
  const double* b = lift ? lift + ((size_t)e * Nf + f) * len : nullptr;
  for (int k = threadIdx.x; k < len; k += blockDim.x) dst[k] = a[k] + (b ? b[k] : 0.0);
}
void MixedSolver::boundaryGradientState(double* Gb) {
  needDevice(); CUDA_OK(cudaSetDevice(device_));
  if (!gradFromStep_) refreshGradient();
  if (F_.nBnd == 0) return;
  std::vector<int> rec[7]; size_t at = 0;
  for (int b = 0; b < F_.nBnd; b++) {
    const int i = F_.nInt + b, t = F_.lt[i];
    rec[t].push_back(F_.le[i]); rec[t].push_back(F_.lf[i]); rec[t].push_back((int)at);
    at += (size_t)block(t).T.Nb * kG;
  }
  DevBuf<double> out; out.alloc(at);
  DevBuf<int> d;
  for (int t = 0; t < 7; t++) if (!rec[t].empty()) {
    MixedBlock& B = block(t);
    d.alloc(rec[t].size());
    CUDA_OK(cudaMemcpyAsync(d.p, rec[t].data(), rec[t].size() * sizeof(int), cudaMemcpyHostToDevice, stream_));
    const int nRec = (int)(rec[t].size() / 3);
    const bool br2 = phys_.visc == kBR2;
    mxBoundaryGradientKernel<<<nRec, 64, 0, stream_>>>(br2 ? B.Gvol.p : B.Gtot.p, br2 ? B.Gf.p : nullptr, d.p, nRec, B.T.Nf, B.T.Nb * kG, out.p); launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(stream_));
  }
  CUDA_OK(cudaMemcpyAsync(Gb, out.p, at * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
}

void MixedSolver::setArtificialViscosity(double empiricalTolerance, double factor, int nodeNumber) {
  if (finalized_) throw std::runtime_error("sdg_set_artificial_viscosity must precede sdg_finalize");
  phys_.av = 1; avTol_ = empiricalTolerance; avFactor_ = factor; avNodes_ = nodeNumber;
}
void MixedSolver::setElementNodes(int type, const int32_t* nodeTag, const double* innerRadius) {
  if (finalized_) throw std::runtime_error("sdg_set_element_nodes must precede sdg_finalize");
  MixedBlock& B = block(type);
  B.tags.assign(nodeTag, nodeTag + (size_t)B.n * B.T.nbasic);
  B.radius.assign(innerRadius, innerRadius + B.n);
  for (int t : B.tags) if (t < 0 || t >= avNodes_) throw std::runtime_error("node tag out of range");
}
// Solver::calculateArtificialViscosity, SpatialDiscrete.cpp:124-192
void MixedSolver::avUpdate() {
  static const double kTol[5] = {0.0, -1.20411998266, -1.90848501888, -2.40823996531, -2.79588001734};   // SimulationControl.cpp:892-893
  Args a; fill(a);
  for (int type : {(int)kTriangle, (int)kQuadrangle}) if (blk_[type]) {
    mxAvIndicatorKernel<<<std::min((blk_[type]->n + 7) / 8, 148 * 8), 256, 0, stream_>>>(a, type == kTriangle ? 0 : 1, p_, kTol[p_ - 1], avTol_, avFactor_); launches++;
  }
  CUDA_OK(cudaMemsetAsync(avNode_.p, 0, avNode_.n * sizeof(double), stream_));
  for (int type : {(int)kTriangle, (int)kQuadrangle}) if (blk_[type]) {
    MixedBlock& B = *blk_[type]; const int NB = B.T.nbasic;
    avNodeMaxKernel<<<std::min((B.n * NB + 255) / 256, 148 * 8), 256, 0, stream_>>>(B.dAvE.p, B.dTags.p, B.n, NB, reinterpret_cast<unsigned long long*>(avNode_.p)); launches++;
  }
  for (int type : {(int)kTriangle, (int)kQuadrangle}) if (blk_[type]) {
    MixedBlock& B = *blk_[type]; const int NB = B.T.nbasic;
    avStoreKernel<<<std::min((B.n * NB + 255) / 256, 148 * 8), 256, 0, stream_>>>(avNode_.p, B.dTags.p, B.n, NB, B.dAvElem.p); launches++;
  }
  CUDA_OK(cudaGetLastError());
}
void MixedSolver::updateArtificialViscosity() {
  needDevice(); CUDA_OK(cudaSetDevice(device_));
  avUpdate();
  CUDA_OK(cudaStreamSynchronize(stream_));
}
void MixedSolver::nodeArtificialViscosity(double* out) {
  needDevice(); CUDA_OK(cudaSetDevice(device_));
  CUDA_OK(cudaMemcpyAsync(out, avNode_.p, (size_t)avNodes_ * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
}
void MixedSolver::elementArtificialViscosity(int type, double* out) {
  needDevice(); CUDA_OK(cudaSetDevice(device_));
  MixedBlock& B = block(type);
  CUDA_OK(cudaMemcpyAsync(out, B.dAvElem.p, B.dAvElem.n * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
}

void MixedSolver::viewVariable(int type, int variable, double* out) {
  needDevice(); CUDA_OK(cudaSetDevice(device_));
  MixedBlock& B = block(type); const MixedTable& T = B.T;
  const size_t npts = (size_t)B.n * T.Nq;
  DevBuf<double> cons, grad, eps, res;
  cons.alloc(npts * kNV); res.alloc(npts);
  mxToQuadratureKernel<<<148 * 4, 256, 0, stream_>>>(B.U.p, B.dPhi_.p, B.n, T.Nb, T.Nq, kNV, cons.p); launches++;
  if (phys_.ns) {
    if (!gradFromStep_) refreshGradient();
    grad.alloc(npts * kG);
    mxToQuadratureKernel<<<148 * 4, 256, 0, stream_>>>(B.Gtot.p, B.dPhi_.p, B.n, T.Nb, T.Nq, kG, grad.p); launches++;
  }
  if (phys_.av) {
    eps.alloc(npts);
    avAtNodesKernel<<<148 * 4, 256, 0, stream_>>>(B.dAvElem.p, B.dNodalQ.p, nullptr, B.n, T.Nq, T.nbasic, eps.p); launches++;
  }
  launchViewVariable(kD, phys_, variable, npts, cons.p, phys_.ns ? grad.p : nullptr, phys_.av ? eps.p : nullptr, res.p, stream_); launches++;
  CUDA_OK(cudaMemcpyAsync(out, res.p, npts * sizeof(double), cudaMemcpyDeviceToHost, stream_));
  CUDA_OK(cudaStreamSynchronize(stream_));
}

double MixedSolver::computeDt(double cfl) {
  needDevice(); CUDA_OK(cudaSetDevice(device_));
  double best = 1.7976931348623157e308;
  for (auto& bp : blk_) if (bp) {
    MixedBlock& B = *bp; const MixedTable& T = B.T;
    const int blocks = (int)std::min<size_t>(((size_t)B.n * T.Nq + 255) / 256, 1024);
    mxDeltaTimeKernel<<<blocks, 256, 0, stream_>>>(B.U.p, B.dPhi_.p, B.dMinEdge.p, B.n, T.Nb, T.Nq, p_, cfl, phys_, dtPartial.p); launches++;
    CUDA_OK(cudaGetLastError());
    std::vector<double> h(blocks);
    CUDA_OK(cudaMemcpyAsync(h.data(), dtPartial.p, sizeof(double) * blocks, cudaMemcpyDeviceToHost, stream_));
    CUDA_OK(cudaStreamSynchronize(stream_));
    for (double v : h) best = std::min(best, v);
  }
  return best;
}

}  // namespace sdg
