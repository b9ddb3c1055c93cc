// physics.cuh — pointwise physics of the DG stage on the device (sm_100a), fp64.
//
// Restates what SubrosaDG evaluates per quadrature point:
//   PhysicalModel.cpp:26-165 (thermodynamics, EOS, transport), VariableConvertor.cpp:291-381,574-620 (variable sets and
//   primitive gradients), ConvectiveFlux.cpp:28-439 (raw/normal flux, Central, Lax-Friedrichs, HLLC, Roe, Exact),
//   ViscousFlux.cpp:59-153, BoundaryCondition.cpp:79-547, SourceTerm.cpp:29-58.
// "Computational" variables = (rho, u[D], e, p); gradients are stored with row index var*D + dir.
#pragma once
#include <cuda_runtime.h>

namespace sdg {

enum { kCompresibleEuler = 0, kCompresibleNS = 1, kIncompresibleEuler = 2, kIncompresibleNS = 3 };
enum { kIdealGas = 0, kWeakCompressibleFluid = 1 };
enum { kTransportNone = 0, kTransportConstant = 1, kTransportSutherland = 2 };
enum { kCentral = 0, kLaxFriedrichs = 1, kHLLC = 2, kRoe = 3, kExact = 4 };
enum { kViscNone = 0, kBR1 = 1, kBR2 = 2 };
enum { kSourceNone = 0, kBoussinesq = 1 };
enum { kRiemannFarfield = 0, kVelocityInflow = 1, kPressureOutflow = 2, kIsoThermalNonSlipWall = 3, kAdiabaticSlipWall = 4,
       kAdiabaticNonSlipWall = 5, kPeriodic = 6 };
enum { kForwardEuler = 0, kHeunRK2 = 1, kSSPRK3 = 2 };

// Reciprocal / square root for the hot loops: hardware seed (MUFU.RCP64H / RSQ64H, >= 20 good bits) + two Newton steps = full
// double precision to within an ulp, without the special-case branches of the IEEE-exact library sequences (the stage kernels
// are bound by the FP64 pipe and by dependent-issue latency, profiles/r01_euler_line_ncu.md).  NaN and sign propagate as usual;
// arguments here are densities, pressures and wave-speed differences (never denormal).
__device__ __forceinline__ double frcp(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);
  return fma(y, e, y);
}
__device__ __forceinline__ double fsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double h = 0.5 * x;
  y = y * fma(-h * y, y, 1.5);
  y = y * fma(-h * y, y, 1.5);
  double s = x * y;
  s = fma(0.5 * y, fma(-s, s, x), s);
  return x == 0.0 ? 0.0 : s;
}

struct PhysParams {
  int model, eos, transport, conv, visc, source;
  int compressible, ns;
  int av, pad_;   // ShockCapturingEnum::ArtificialViscosity (Euler models): the gradient pass runs and the viscous terms are eps * grad(U_conserved)
  double cp, cv, icv, kg, gamma, mu0, k0, c0, rho0, padd, beta, tref;
};

// PH = 1: compile-time specialisation CompresibleEuler/NS + IdealGas + HLLC (the benchmark configurations);
// PH = 0: every switch is read from PhysParams at run time (uniform branches).
template <int PH>
struct Phys {
  const PhysParams& P;
  __device__ __forceinline__ explicit Phys(const PhysParams& p) : P(p) {}
  __device__ __forceinline__ bool comp() const { if constexpr (PH == 1) return true; else return P.compressible != 0; }
  __device__ __forceinline__ bool ideal() const { if constexpr (PH == 1) return true; else return P.eos == kIdealGas; }
  __device__ __forceinline__ int conv() const { if constexpr (PH == 1) return kHLLC; else return P.conv; }
  __device__ __forceinline__ double pressure(double rho, double e) const {  // PhysicalModel.cpp:47-49,68-72
    return ideal() ? (P.gamma - 1.0) * rho * e : P.c0 * P.c0 * (rho - P.rho0) + P.padd;
  }
  __device__ __forceinline__ double sound(double rho, double p) const {  // :51-54,74-77
    return ideal() ? fsqrt(P.gamma * p * frcp(rho)) : P.c0;
  }
  __device__ __forceinline__ double TFromE(double e) const { return e * P.icv; }
  __device__ __forceinline__ double sutherland(double T) const {  // :100-122
    const double Ts = 110.4 / 273.15;
    return sqrt(T * T * T) * (1.0 + Ts) / (T + Ts);
  }
  __device__ __forceinline__ double mu(double T) const { return P.transport == kTransportSutherland ? P.mu0 * sutherland(T) : P.mu0; }
  __device__ __forceinline__ double kappa(double T) const { return P.transport == kTransportSutherland ? P.k0 * sutherland(T) : P.k0; }
};

template <int D>
__device__ __forceinline__ double vsq(const double* c) { double s = 0; for (int d = 0; d < D; d++) s += c[1 + d] * c[1 + d]; return s; }
template <int D>
__device__ __forceinline__ double dotn(const double* c, const double* n) { double s = 0; for (int d = 0; d < D; d++) s += c[1 + d] * n[d]; return s; }

// Variable::calculateComputationalFromConserved, VariableConvertor.cpp:315-339
template <int D, int PH>
__device__ __forceinline__ double compFromCons(const Phys<PH>& ph, const double* cons, double* comp) {
  const double rho = cons[0];
  const double ir = frcp(rho);
  comp[0] = rho;
#pragma unroll
  for (int d = 0; d < D; d++) comp[1 + d] = cons[1 + d] * ir;
  double e = cons[D + 1] * ir;
  if (ph.comp()) e -= 0.5 * vsq<D>(comp);
  comp[D + 1] = e;
  comp[D + 2] = ph.pressure(rho, e);
  return ir;
}
// calculateConservedFromComputational, :291-313
template <int D, int PH>
__device__ __forceinline__ void consFromComp(const Phys<PH>& ph, const double* comp, double* cons) {
  const double rho = comp[0];
  cons[0] = rho;
#pragma unroll
  for (int d = 0; d < D; d++) cons[1 + d] = rho * comp[1 + d];
  cons[D + 1] = ph.comp() ? rho * (comp[D + 1] + 0.5 * vsq<D>(comp)) : rho * comp[D + 1];
}
// primitive (rho, u, T) -> computational, :368-381
template <int D, int PH>
__device__ __forceinline__ void compFromPrim(const Phys<PH>& ph, const double* prim, double* comp) {
  comp[0] = prim[0];
#pragma unroll
  for (int d = 0; d < D; d++) comp[1 + d] = prim[1 + d];
  const double e = ph.P.cv * prim[D + 1];
  comp[D + 1] = e;
  comp[D + 2] = ph.pressure(prim[0], e);
}

// calculateConvectiveRawFlux, ConvectiveFlux.cpp:28-57: F[v*D+d]
template <int D, int PH>
__device__ __forceinline__ void convRawFlux(const Phys<PH>& ph, const double* comp, double* F) {
  const double rho = comp[0], p = comp[D + 2];
#pragma unroll
  for (int d = 0; d < D; d++) F[d] = rho * comp[1 + d];
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int d = 0; d < D; d++) F[(1 + c) * D + d] = rho * comp[1 + d] * comp[1 + c] + (c == d ? p : 0.0);
  const double h = ph.comp() ? rho * (comp[D + 1] + 0.5 * vsq<D>(comp)) + p : rho * comp[D + 1];
#pragma unroll
  for (int d = 0; d < D; d++) F[(D + 1) * D + d] = h * comp[1 + d];
}
// calculateConvectiveNormalFlux, :60-91
template <int D, int PH>
__device__ __forceinline__ void convNormalFlux(const Phys<PH>& ph, const double* n, const double* comp, double* Fn) {
  const double rho = comp[0], p = comp[D + 2];
  const double un = dotn<D>(comp, n);
  Fn[0] = rho * un;
#pragma unroll
  for (int d = 0; d < D; d++) Fn[1 + d] = rho * un * comp[1 + d] + p * n[d];
  Fn[D + 1] = ph.comp() ? (rho * (comp[D + 1] + 0.5 * vsq<D>(comp)) + p) * un : rho * comp[D + 1] * un;
}

// calculateConvectiveHLLCFlux, :137-238 (pressure estimate without the 1/2 on the velocity jump, p* in the star energy).
// Restated branch-free (the reference's early returns S_L >= 0 -> F_L, S_R <= 0 -> F_R become selects, so a warp never
// diverges) and with the wave-speed factor folded into the sound speed:
//   c_K q_K = sqrt(g p_K/rho_K) sqrt(1 + (g+1)/(2g) (p*/p_K - 1)) = sqrt(g/rho_K (p_K + (g+1)/(2g) max(p* - p_K, 0)))
// which is the same number up to round-off and costs one square root instead of two divisions and two square roots.
// irL/irR = 1/rho of the two states (already known from the conserved -> computational conversion).
template <int D, int PH>
__device__ __forceinline__ void hllcFlux(const Phys<PH>& ph, const double* n, const double* consL, const double* compL, double irL,
                                         const double* consR, const double* compR, double irR, double* F) {
  constexpr int NV = D + 2;
  const double g = ph.P.gamma;
  const double rL = compL[0], rR = compR[0], pL = compL[D + 2], pR = compR[D + 2];
  const double unL = dotn<D>(compL, n), unR = dotn<D>(compR, n);
  const double cL = fsqrt(g * pL * irL), cR = fsqrt(g * pR * irR);
  const double ps = fmax(0.0, 0.5 * (pL + pR) - (unR - unL) * (0.5 * (rL + rR)) * (0.5 * (cL + cR)));
  const double kg = ph.P.kg;   // (gamma + 1) / (2 gamma)
  const double SL = unL - fsqrt(g * irL * (pL + kg * fmax(ps - pL, 0.0)));
  const double SR = unR + fsqrt(g * irR * (pR + kg * fmax(ps - pR, 0.0)));
  const double mL = rL * (SL - unL), mR = rR * (SR - unR);
  const double Ss = (pR - pL + mL * unL - mR * unR) * frcp(mL - mR);
  const bool left = SL >= 0.0 ? true : (SR <= 0.0 ? false : Ss >= 0.0);
  const bool pure = SL >= 0.0 || SR <= 0.0;
  const double S = left ? SL : SR, un = left ? unL : unR, p = left ? pL : pR, m = left ? mL : mR;
  double comp[D + 3], cons[NV];
#pragma unroll
  for (int k = 0; k < D + 3; k++) comp[k] = left ? compL[k] : compR[k];
#pragma unroll
  for (int v = 0; v < NV; v++) cons[v] = left ? consL[v] : consR[v];
  double FK[NV];
  convNormalFlux<D>(ph, n, comp, FK);
  const double inv = frcp(S - Ss);
  double Us[NV];
  Us[0] = m * inv;
#pragma unroll
  for (int d = 0; d < D; d++) Us[1 + d] = (m * comp[1 + d] + (ps - p) * n[d]) * inv;
  Us[D + 1] = (m * (comp[D + 1] + 0.5 * vsq<D>(comp)) - p * un + ps * Ss) * inv;
#pragma unroll
  for (int v = 0; v < NV; v++) F[v] = pure ? FK[v] : FK[v] + S * (Us[v] - cons[v]);
}

// calculateConvectiveFlux dispatch, :417-439
template <int D, int PH>
__device__ __forceinline__ void convFlux(const Phys<PH>& ph, const double* n, const double* consL, const double* compL, double irL,
                                         const double* consR, const double* compR, double irR, double* F) {
  constexpr int NV = D + 2;
  const int kind = ph.conv();
  if (kind == kHLLC) { hllcFlux<D>(ph, n, consL, compL, irL, consR, compR, irR, F); return; }
  if constexpr (PH == 0) {
    double FL[NV], FR[NV];
    if (kind == kCentral) {  // :94-104
      convNormalFlux<D>(ph, n, compL, FL); convNormalFlux<D>(ph, n, compR, FR);
      for (int v = 0; v < NV; v++) F[v] = 0.5 * (FL[v] + FR[v]);
    } else if (kind == kLaxFriedrichs) {  // :107-134
      convNormalFlux<D>(ph, n, compL, FL); convNormalFlux<D>(ph, n, compR, FR);
      const double sL = fabs(dotn<D>(compL, n)) + ph.sound(compL[0], compL[D + 2]);
      const double sR = fabs(dotn<D>(compR, n)) + ph.sound(compR[0], compR[D + 2]);
      const double sr = fmax(sL, sR);
      for (int v = 0; v < NV; v++) F[v] = 0.5 * ((FL[v] + FR[v]) - sr * (consR[v] - consL[v]));
    } else if (kind == kRoe) {  // :241-350
      const double g = ph.P.gamma;
      convNormalFlux<D>(ph, n, compL, FL); convNormalFlux<D>(ph, n, compR, FR);
      const double sL = sqrt(compL[0]), sR = sqrt(compR[0]), ss = sL + sR;
      const double rho = sqrt(compL[0] * compR[0]);
      double u[D], q2 = 0;
      for (int d = 0; d < D; d++) { u[d] = (sL * compL[1 + d] + sR * compR[1 + d]) / ss; q2 += u[d] * u[d]; }
      const double HL = compL[D + 1] * g + 0.5 * vsq<D>(compL), HR = compR[D + 1] * g + 0.5 * vsq<D>(compR);
      const double H = (sL * HL + sR * HR) / ss;
      const double e = (H - 0.5 * q2) / g;
      const double p = ph.pressure(rho, e);
      double un = 0; for (int d = 0; d < D; d++) un += u[d] * n[d];
      const double c = ph.sound(rho, p);
      double dc[D + 3]; for (int k = 0; k < D + 3; k++) dc[k] = compR[k] - compL[k];
      double dun = 0; for (int d = 0; d < D; d++) dun += dc[1 + d] * n[d];
      const double hd = c / 20.0;  // Harten entropy fix on u -+ c only
      const double lm = fabs(un - c) > hd ? fabs(un - c) : ((un - c) * (un - c) + hd * hd) / (2.0 * hd);
      const double lp = fabs(un + c) > hd ? fabs(un + c) : ((un + c) * (un + c) + hd * hd) / (2.0 * hd);
      double sum[NV]; for (int v = 0; v < NV; v++) sum[v] = 0.0;
      { const double f = lm * (dc[D + 2] - rho * c * dun) / (2.0 * c * c);
        sum[0] += f; for (int d = 0; d < D; d++) sum[1 + d] += f * (u[d] - c * n[d]); sum[D + 1] += f * (H - c * un); }
      { const double f = fabs(un) * (dc[0] - dc[D + 2] / (c * c));
        sum[0] += f; for (int d = 0; d < D; d++) sum[1 + d] += f * u[d]; sum[D + 1] += f * (0.5 * q2); }
      if (D >= 2) {
        const double f = fabs(un) * rho;
        double udu = 0; for (int d = 0; d < D; d++) udu += u[d] * dc[1 + d];
        for (int d = 0; d < D; d++) sum[1 + d] += f * (dc[1 + d] - dun * n[d]);
        sum[D + 1] += f * (udu - un * dun);
      }
      { const double f = lp * (dc[D + 2] + rho * c * dun) / (2.0 * c * c);
        sum[0] += f; for (int d = 0; d < D; d++) sum[1 + d] += f * (u[d] + c * n[d]); sum[D + 1] += f * (H + c * un); }
      for (int v = 0; v < NV; v++) F[v] = 0.5 * ((FL[v] + FR[v]) - sum[v]);
    } else {  // kExact, :353-414 (weakly compressible: constant sound speed)
      const double c = ph.sound(0.0, 0.0);
      const double unL = dotn<D>(compL, n), unR = dotn<D>(compR, n);
      const double rho = sqrt(compL[0] * compR[0] * exp((unL - unR) / c));
      const double un = 0.5 * (unL + unR) + log(compL[0] / compR[0]) * c * 0.5;
      const double* S = un < 0.0 ? compR : compL;
      const double unS = un < 0.0 ? unR : unL;
      double x[D + 3];
      const double e = S[D + 1] * S[0] / rho;
      x[0] = rho;
      for (int d = 0; d < D; d++) x[1 + d] = S[1 + d] + (un - unS) * n[d];
      x[D + 1] = e; x[D + 2] = ph.pressure(rho, e);
      convNormalFlux<D>(ph, n, x, F);
    }
  }
}

// VariableGradient::calculatePrimitiveFromConserved, VariableConvertor.cpp:574-620 (gradient of rho, u, T)
template <int D, int PH>
__device__ __forceinline__ void primGradFromConsGrad(const Phys<PH>& ph, const double* cons, const double* comp, const double* gc, double* gp) {
  const double ir = frcp(comp[0]);
#pragma unroll
  for (int d = 0; d < D; d++) gp[d] = gc[d];
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int r = 0; r < D; r++) gp[(1 + c) * D + r] = (gc[(1 + c) * D + r] - gc[r] * comp[1 + c]) * ir;
  const double E = cons[D + 1] * ir;
  const double icv = ph.P.icv;
#pragma unroll
  for (int r = 0; r < D; r++) {
    double ge = (gc[(D + 1) * D + r] - gc[r] * E) * ir;
    if (ph.comp()) { double s = 0; for (int c = 0; c < D; c++) s += gp[(1 + c) * D + r] * comp[1 + c]; ge -= s; }
    gp[(D + 1) * D + r] = ge * icv;
  }
}
// calculateViscousRawFlux, ViscousFlux.cpp:59-103: F[v*D+d]
template <int D, int PH>
__device__ __forceinline__ void viscRawFlux(const Phys<PH>& ph, const double* comp, const double* gp, double* F) {
#pragma unroll
  for (int d = 0; d < D; d++) F[d] = 0.0;
  const double T = ph.TFromE(comp[D + 1]);
  const double mu = ph.mu(T), k = ph.kappa(T);
  double tr = 0;
#pragma unroll
  for (int d = 0; d < D; d++) tr += gp[(1 + d) * D + d];
  double tau[D][D];
#pragma unroll
  for (int r = 0; r < D; r++)
#pragma unroll
    for (int c = 0; c < D; c++) tau[r][c] = mu * (gp[(1 + c) * D + r] + gp[(1 + r) * D + c]) - (r == c ? 2.0 / 3.0 * mu * tr : 0.0);
#pragma unroll
  for (int c = 0; c < D; c++)
#pragma unroll
    for (int r = 0; r < D; r++) F[(1 + c) * D + r] = tau[r][c];
#pragma unroll
  for (int r = 0; r < D; r++) {
    double s = 0;
    if (ph.comp()) for (int c = 0; c < D; c++) s += tau[r][c] * comp[1 + c];
    F[(D + 1) * D + r] = s + k * gp[(D + 1) * D + r];
  }
}
template <int D, int PH>
__device__ __forceinline__ void viscNormalFlux(const Phys<PH>& ph, const double* n, const double* comp, const double* gp, double* Fn) {  // :116-124
  double F[D * (D + 2)];
  viscRawFlux<D>(ph, comp, gp, F);
#pragma unroll
  for (int v = 0; v < D + 2; v++) { double s = 0; for (int d = 0; d < D; d++) s += F[v * D + d] * n[d]; Fn[v] = s; }
}

__device__ __forceinline__ bool bcIsWall(int bc) { return bc == kIsoThermalNonSlipWall || bc == kAdiabaticSlipWall || bc == kAdiabaticNonSlipWall; }

// BoundaryConditionImpl<...>::calculateBoundaryVariable, BoundaryCondition.cpp:79-547: computational column of the boundary
// state from the interior trace L and the user-supplied state R (boundary_dummy_variable_).
template <int D, int PH>
__device__ void bcBoundaryVariable(const Phys<PH>& ph, int bc, const double* n, const double* L, const double* R, double* b) {
  constexpr int NC = D + 3;
  switch (bc) {
    case kRiemannFarfield: {  // :82-285
      const double unL = dotn<D>(L, n), unR = dotn<D>(R, n);
      const double cL = ph.sound(L[0], L[D + 2]);
      const double mach = unL / cL;
      if (fabs(mach) > 1.0) { const double* s = mach < 0.0 ? R : L; for (int k = 0; k < NC; k++) b[k] = s[k]; return; }
      const double* S = mach < 0.0 ? R : L;
      const double unS = mach < 0.0 ? unR : unL;
      if (ph.comp()) {
        const double g = ph.P.gamma;
        const double Rm = unR - 2.0 * ph.sound(R[0], R[D + 2]) / (g - 1.0);
        const double Rp = unL + 2.0 * cL / (g - 1.0);
        const double bun = 0.5 * (Rm + Rp);
        const double c = (g - 1.0) * (Rp - Rm) / 4.0;
        const double s = S[D + 2] / pow(S[0], g);  // calculateEntropyFromDensityPressure, PhysicalModel.cpp:148-150
        const double rho = pow(c * c / (g * s), 1.0 / (g - 1.0));
        const double p = rho * c * c / g;
        b[0] = rho;
        for (int d = 0; d < D; d++) b[1 + d] = S[1 + d] + (bun - unS) * n[d];
        b[D + 1] = p / ((g - 1.0) * rho); b[D + 2] = p;
      } else {
        const double c = ph.sound(0.0, 0.0);
        const double rho = sqrt(L[0] * R[0] * exp((unL - unR) / c));
        const double bun = 0.5 * (unL + unR) + log(L[0] / R[0]) * c * 0.5;
        const double e = S[D + 1] * S[0] / rho;
        b[0] = rho;
        for (int d = 0; d < D; d++) b[1 + d] = S[1 + d] + (bun - unS) * n[d];
        b[D + 1] = e; b[D + 2] = ph.pressure(rho, e);
      }
      return;
    }
    case kVelocityInflow: {  // :313-332
      const double mach = dotn<D>(L, n) / ph.sound(L[0], L[D + 2]);
      for (int k = 0; k < NC; k++) b[k] = R[k];
      if (mach > -1.0) b[D + 2] = L[D + 2];
      return;
    }
    case kPressureOutflow: {  // :360-379
      const double mach = dotn<D>(L, n) / ph.sound(L[0], L[D + 2]);
      for (int k = 0; k < NC; k++) b[k] = L[k];
      if (mach < 1.0) b[D + 2] = R[D + 2];
      return;
    }
    case kIsoThermalNonSlipWall: {  // :407-424
      b[0] = L[0];
      for (int d = 0; d < D; d++) b[1 + d] = R[1 + d];
      b[D + 1] = R[D + 1];
      b[D + 2] = ph.pressure(L[0], R[D + 1]);
      return;
    }
    case kAdiabaticSlipWall: {  // :458-471
      for (int k = 0; k < NC; k++) b[k] = L[k];
      const double un = dotn<D>(L, n);
      for (int d = 0; d < D; d++) b[1 + d] = L[1 + d] - un * n[d];
      return;
    }
    default: {  // kAdiabaticNonSlipWall, :507-516
      for (int k = 0; k < NC; k++) b[k] = L[k];
      for (int d = 0; d < D; d++) b[1 + d] = R[1 + d];
      return;
    }
  }
}

// calculateBoundaryGradientVariable (:287-297 and the wall variants :426-441,473-488,518-533): conserved state entering
// the volume-gradient flux and the interface-gradient (lifting) flux of a boundary face.
template <int D, int PH>
__device__ __forceinline__ void bcBoundaryGradientVariable(const Phys<PH>& ph, int bc, const double* n, const double* consL, const double* compL,
                                                           const double* compR, double* volCons, double* intCons) {
  constexpr int NV = D + 2;
  if (!bcIsWall(bc)) {
    for (int v = 0; v < NV; v++) { volCons[v] = consL[v]; intCons[v] = 0.0; }
    return;
  }
  double b[D + 3], bc_cons[NV];
  bcBoundaryVariable<D>(ph, bc, n, compL, compR, b);
  consFromComp<D>(ph, b, bc_cons);
  for (int v = 0; v < NV; v++) { volCons[v] = bc_cons[v]; intCons[v] = bc_cons[v] - consL[v]; }
}

// SourceTermBase<Boussinesq>::calculateSourceTerm, SourceTerm.cpp:29-58 (unit gravity along the last axis)
template <int D, int PH>
__device__ __forceinline__ double boussinesqSource(const Phys<PH>& ph, const double* comp) {
  return comp[0] * ph.P.beta * (ph.TFromE(comp[D + 1]) - ph.P.tref);
}

}  // namespace sdg
