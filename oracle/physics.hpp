// oracle/physics.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle).  Pointwise physics restated from
//   src/Solver/PhysicalModel.cpp, VariableConvertor.cpp:204-421,574-620, ConvectiveFlux.cpp, ViscousFlux.cpp,
//   BoundaryCondition.cpp:79-547, SourceTerm.cpp:29-58 of SubrosaDG.  "Restate the code, not Toro."
#pragma once
#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace orc {

enum EquationModel { kCompresibleEuler = 0, kCompresibleNS = 1, kIncompresibleEuler = 2, kIncompresibleNS = 3 };  // Enum.cpp:56-64
enum Eos { kIdealGas = 0, kWeakCompressibleFluid = 1 };                                                          // :125-128
enum Transport { kTransportNone = 0, kTransportConstant = 1, kTransportSutherland = 2 };                          // :130-134
enum ConvFlux { kCentral = 0, kLaxFriedrichs = 1, kHLLC = 2, kRoe = 3, kExact = 4 };                              // :107-113
enum ViscFlux { kViscNone = 0, kBR1 = 1, kBR2 = 2 };                                                              // :115-119
enum SourceKind { kSourceNone = 0, kBoussinesq = 1 };                                                             // :71-74
enum BcType { kRiemannFarfield = 0, kVelocityInflow = 1, kPressureOutflow = 2, kIsoThermalNonSlipWall = 3,
              kAdiabaticSlipWall = 4, kAdiabaticNonSlipWall = 5, kPeriodic = 6 };                                 // :92-100
enum TimeScheme { kForwardEuler = 0, kHeunRK2 = 1, kSSPRK3 = 2 };                                                 // :136-140

constexpr int kMaxV = 5;   // conserved variables (D+2)
constexpr int kMaxC = 6;   // computational variables (rho, u[D], e, p)
constexpr int kMaxD = 3;

struct Phys {
  int D = 2, Nv = 4;
  int model = kCompresibleEuler, eos = kIdealGas, transport = kTransportNone, conv = kHLLC, visc = kViscNone, source = kSourceNone;
  double cp = 2.5, cv = 25.0 / 14.0;
  double gamma = 1.4;  // EquationOfState<IdealGas>::kSpecificHeatRatio, PhysicalModel.cpp:45
  double mu0 = 0.0, k0 = 0.0;
  double c0 = 1.0, rho0 = 1.0, padd = 0.0;
  double beta = 0.0, Tref = 0.0;
  bool comp() const { return model == kCompresibleEuler || model == kCompresibleNS; }
  bool ns() const { return model == kCompresibleNS || model == kIncompresibleNS; }

  // PhysicalModel.cpp:31-37
  double eFromT(double T) const { return cv * T; }
  double TFromE(double e) const { return e / cv; }
  // :47-54, :68-77
  double pressure(double rho, double e) const { return eos == kIdealGas ? (gamma - 1.0) * rho * e : c0 * c0 * (rho - rho0) + padd; }
  double sound(double rho, double p) const { return eos == kIdealGas ? std::sqrt(gamma * p / rho) : c0; }
  double entropy(double rho, double p) const { return p / std::pow(rho, gamma); }  // :148-150
  // :93-122
  double sutherland(double T) const { const double Ts = 110.4 / 273.15; return std::sqrt(T * T * T) * (1.0 + Ts) / (T + Ts); }
  double mu(double T) const { return transport == kTransportSutherland ? mu0 * sutherland(T) : mu0; }
  double kappa(double T) const { return transport == kTransportSutherland ? k0 * sutherland(T) : k0; }
};

// One column of Variable<SC,N> (VariableConvertor.cpp:204-421): conserved (Nv), computational (Nv+1: rho,u,e,p),
// primitive (Nv: rho,u,T).
struct Var {
  double cons[kMaxV], comp[kMaxC], prim[kMaxV];
};

inline double vsq(const Phys& P, const double* comp) { double s = 0; for (int d = 0; d < P.D; d++) s += comp[1 + d] * comp[1 + d]; return s; }

inline void compFromCons(const Phys& P, Var& v) {  // :315-339
  const int D = P.D;
  const double rho = v.cons[0];
  v.comp[0] = rho;
  for (int d = 0; d < D; d++) v.comp[1 + d] = v.cons[1 + d] / rho;
  double e;
  if (P.comp()) e = v.cons[D + 1] / rho - vsq(P, v.comp) / 2.0; else e = v.cons[D + 1] / rho;
  v.comp[D + 1] = e;
  v.comp[D + 2] = P.pressure(rho, e);
}
inline void consFromComp(const Phys& P, Var& v) {  // :291-313
  const int D = P.D;
  const double rho = v.comp[0];
  v.cons[0] = rho;
  for (int d = 0; d < D; d++) v.cons[1 + d] = rho * v.comp[1 + d];
  if (P.comp()) v.cons[D + 1] = rho * (v.comp[D + 1] + vsq(P, v.comp) / 2.0); else v.cons[D + 1] = rho * v.comp[D + 1];
}
inline void consFromPrim(const Phys& P, Var& v) {  // :341-366 (also fills the computational velocity)
  const int D = P.D;
  const double rho = v.prim[0];
  v.cons[0] = rho;
  for (int d = 0; d < D; d++) { v.cons[1 + d] = rho * v.prim[1 + d]; v.comp[1 + d] = v.prim[1 + d]; }
  if (P.comp()) v.cons[D + 1] = rho * (P.eFromT(v.prim[D + 1]) + vsq(P, v.comp) / 2.0); else v.cons[D + 1] = rho * P.eFromT(v.prim[D + 1]);
}
inline void compFromPrim(const Phys& P, Var& v) {  // :368-381
  const int D = P.D;
  v.comp[0] = v.prim[0];
  for (int d = 0; d < D; d++) v.comp[1 + d] = v.prim[1 + d];
  const double e = P.eFromT(v.prim[D + 1]);
  v.comp[D + 1] = e;
  v.comp[D + 2] = P.pressure(v.prim[0], e);
}

// calculateConvectiveRawFlux, ConvectiveFlux.cpp:28-57.  F is D x Nv column-major: F[v*D + d].
inline void convRawFlux(const Phys& P, const double* comp, double* F) {
  const int D = P.D;
  const double rho = comp[0], p = comp[D + 2];
  for (int d = 0; d < D; d++) F[0 * D + d] = rho * comp[1 + d];
  for (int c = 0; c < D; c++) for (int d = 0; d < D; d++) F[(1 + c) * D + d] = rho * comp[1 + d] * comp[1 + c] + (c == d ? p : 0.0);
  if (P.comp()) { const double E = comp[D + 1] + vsq(P, comp) / 2.0; for (int d = 0; d < D; d++) F[(D + 1) * D + d] = (rho * E + p) * comp[1 + d]; }
  else for (int d = 0; d < D; d++) F[(D + 1) * D + d] = rho * comp[D + 1] * comp[1 + d];
}
// calculateConvectiveNormalFlux, :60-91
inline void convNormalFlux(const Phys& P, const double* n, const double* comp, double* Fn) {
  const int D = P.D;
  const double rho = comp[0], p = comp[D + 2];
  double un = 0; for (int d = 0; d < D; d++) un += comp[1 + d] * n[d];
  Fn[0] = rho * un;
  for (int d = 0; d < D; d++) Fn[1 + d] = rho * un * comp[1 + d] + p * n[d];
  if (P.comp()) Fn[D + 1] = (rho * (comp[D + 1] + vsq(P, comp) / 2.0) + p) * un; else Fn[D + 1] = rho * comp[D + 1] * un;
}

// calculateConvectiveFlux dispatch, :417-439.  L,R carry cons and comp.
inline void convFlux(const Phys& P, const double* n, const Var& L, const Var& R, double* F) {
  const int D = P.D, Nv = P.Nv;
  double FL[kMaxV], FR[kMaxV];
  auto dotn = [&](const double* c) { double s = 0; for (int d = 0; d < D; d++) s += c[1 + d] * n[d]; return s; };
  switch (P.conv) {
    case kCentral: {  // :94-104
      convNormalFlux(P, n, L.comp, FL); convNormalFlux(P, n, R.comp, FR);
      for (int v = 0; v < Nv; v++) F[v] = (FL[v] + FR[v]) / 2.0;
      return;
    }
    case kLaxFriedrichs: {  // :107-134
      convNormalFlux(P, n, L.comp, FL); convNormalFlux(P, n, R.comp, FR);
      const double unL = dotn(L.comp), unR = dotn(R.comp);
      const double cL = P.sound(L.comp[0], L.comp[D + 2]), cR = P.sound(R.comp[0], R.comp[D + 2]);
      const double sr = std::max(std::fabs(unL) + cL, std::fabs(unR) + cR);
      for (int v = 0; v < Nv; v++) F[v] = ((FL[v] + FR[v]) - sr * (R.cons[v] - L.cons[v])) / 2.0;
      return;
    }
    case kHLLC: {  // :137-238
      if (P.eos != kIdealGas) throw std::runtime_error("oracle: HLLC needs the ideal-gas EOS (reference uses kSpecificHeatRatio)");
      const double g = P.gamma;
      const double rL = L.comp[0], rR = R.comp[0], pL = L.comp[D + 2], pR = R.comp[D + 2];
      const double unL = dotn(L.comp), unR = dotn(R.comp);
      const double cL = P.sound(rL, pL), cR = P.sound(rR, pR);
      const double rbar = (rL + rR) / 2.0, cbar = (cL + cR) / 2.0;
      const double ps = std::max(0.0, (pL + pR) / 2.0 - (unR - unL) * rbar * cbar);  // NOTE: no 1/2 on the jump term
      const double SL = unL - cL * (ps <= pL ? 1.0 : std::sqrt(1.0 + (g + 1.0) * (ps / pL - 1.0) / 2.0 / g));
      if (SL >= 0.0) { convNormalFlux(P, n, L.comp, F); return; }
      const double SR = unR + cR * (ps <= pR ? 1.0 : std::sqrt(1.0 + (g + 1.0) * (ps / pR - 1.0) / 2.0 / g));
      if (SR <= 0.0) { convNormalFlux(P, n, R.comp, F); return; }
      const double Ss = (pR - pL + rL * unL * (SL - unL) - rR * unR * (SR - unR)) / (rL * (SL - unL) - rR * (SR - unR));
      double Us[kMaxV];
      if (Ss >= 0.0) {
        convNormalFlux(P, n, L.comp, FL);
        Us[0] = rL * (SL - unL) / (SL - Ss);
        for (int d = 0; d < D; d++) Us[1 + d] = ((SL - unL) * rL * L.comp[1 + d] + (ps - pL) * n[d]) / (SL - Ss);
        Us[D + 1] = ((SL - unL) * rL * (L.comp[D + 1] + vsq(P, L.comp) / 2.0) - pL * unL + ps * Ss) / (SL - Ss);
        for (int v = 0; v < Nv; v++) F[v] = FL[v] + SL * (Us[v] - L.cons[v]);
      } else {
        convNormalFlux(P, n, R.comp, FR);
        Us[0] = rR * (SR - unR) / (SR - Ss);
        for (int d = 0; d < D; d++) Us[1 + d] = ((SR - unR) * rR * R.comp[1 + d] + (ps - pR) * n[d]) / (SR - Ss);
        Us[D + 1] = ((SR - unR) * rR * (R.comp[D + 1] + vsq(P, R.comp) / 2.0) - pR * unR + ps * Ss) / (SR - Ss);
        for (int v = 0; v < Nv; v++) F[v] = FR[v] + SR * (Us[v] - R.cons[v]);
      }
      return;
    }
    case kRoe: {  // :241-350
      if (P.eos != kIdealGas) throw std::runtime_error("oracle: Roe needs the ideal-gas EOS");
      const double g = P.gamma;
      convNormalFlux(P, n, L.comp, FL); convNormalFlux(P, n, R.comp, FR);
      const double sL = std::sqrt(L.comp[0]), sR = std::sqrt(R.comp[0]), ss = sL + sR;
      const double rho = std::sqrt(L.comp[0] * R.comp[0]);
      double u[kMaxD], q2 = 0;
      for (int d = 0; d < D; d++) { u[d] = (sL * L.comp[1 + d] + sR * R.comp[1 + d]) / ss; q2 += u[d] * u[d]; }
      const double HL = L.comp[D + 1] * g + vsq(P, L.comp) / 2.0, HR = R.comp[D + 1] * g + vsq(P, R.comp) / 2.0;
      const double H = (sL * HL + sR * HR) / ss;
      const double e = (H - q2 / 2.0) / g;
      const double p = P.pressure(rho, e);
      double un = 0; for (int d = 0; d < D; d++) un += u[d] * n[d];
      const double c = P.sound(rho, p);
      double dc[kMaxC]; for (int k = 0; k < D + 3; k++) dc[k] = R.comp[k] - L.comp[k];
      double dun = 0; for (int d = 0; d < D; d++) dun += dc[1 + d] * n[d];
      const double hd = c / 20.0;
      const double lm = std::fabs(un - c) > hd ? std::fabs(un - c) : ((un - c) * (un - c) + hd * hd) / (2.0 * hd);
      const double lp = std::fabs(un + c) > hd ? std::fabs(un + c) : ((un + c) * (un + c) + hd * hd) / (2.0 * hd);
      double sum[kMaxV] = {0, 0, 0, 0, 0};
      {  // column 0
        const double f = lm * (dc[D + 2] - rho * c * dun) / (2.0 * c * c);
        sum[0] += f; for (int d = 0; d < D; d++) sum[1 + d] += f * (u[d] - c * n[d]); sum[D + 1] += f * (H - c * un);
      }
      {  // column 1
        const double f = std::fabs(un) * (dc[0] - dc[D + 2] / (c * c));
        sum[0] += f; for (int d = 0; d < D; d++) sum[1 + d] += f * u[d]; sum[D + 1] += f * (q2 / 2.0);
      }
      if (D >= 2) {  // column 2
        const double f = std::fabs(un) * rho;
        double udu = 0; for (int d = 0; d < D; d++) udu += u[d] * dc[1 + d];
        for (int d = 0; d < D; d++) sum[1 + d] += f * (dc[1 + d] - dun * n[d]);
        sum[D + 1] += f * (udu - un * dun);
      }
      {  // column D+1
        const double f = lp * (dc[D + 2] + rho * c * dun) / (2.0 * c * c);
        sum[0] += f; for (int d = 0; d < D; d++) sum[1 + d] += f * (u[d] + c * n[d]); sum[D + 1] += f * (H + c * un);
      }
      for (int v = 0; v < Nv; v++) F[v] = ((FL[v] + FR[v]) - sum[v]) / 2.0;
      return;
    }
    case kExact: {  // :353-414
      const double c = P.sound(0.0, 0.0);
      const double unL = dotn(L.comp), unR = dotn(R.comp);
      const double rho = std::sqrt(L.comp[0] * R.comp[0] * std::exp((unL - unR) / c));
      const double un = (unL + unR) / 2.0 + std::log(L.comp[0] / R.comp[0]) * c / 2.0;
      const Var& S = un < 0.0 ? R : L;
      const double unS = un < 0.0 ? unR : unL;
      double x[kMaxC];
      const double e = S.comp[D + 1] * S.comp[0] / rho;
      x[0] = rho;
      for (int d = 0; d < D; d++) x[1 + d] = S.comp[1 + d] + (un - unS) * n[d];
      x[D + 1] = e; x[D + 2] = P.pressure(rho, e);
      convNormalFlux(P, n, x, F);
      return;
    }
  }
  throw std::runtime_error("oracle: unknown convective flux");
}

// VariableGradient::calculatePrimitiveFromConserved, VariableConvertor.cpp:574-620.  Gradient rows = var*D + dir.
inline void primGradFromConsGrad(const Phys& P, const Var& v, const double* gc, double* gp) {
  const int D = P.D;
  const double rho = v.comp[0];
  for (int d = 0; d < D; d++) gp[d] = gc[d];
  // velocity_gradient(dir r, comp c) = (d_r m_c - d_r rho * u_c)/rho, stored at (1+c)*D + r
  for (int c = 0; c < D; c++) for (int r = 0; r < D; r++) gp[(1 + c) * D + r] = (gc[(1 + c) * D + r] - gc[r] * v.comp[1 + c]) / rho;
  double ge[kMaxD];
  if (P.comp()) {
    const double E = v.cons[D + 1] / rho;
    for (int r = 0; r < D; r++) {
      double s = 0; for (int c = 0; c < D; c++) s += gp[(1 + c) * D + r] * v.comp[1 + c];
      ge[r] = (gc[(D + 1) * D + r] - gc[r] * E) / rho - s;
    }
  } else {
    const double e = v.cons[D + 1] / rho;
    for (int r = 0; r < D; r++) ge[r] = (gc[(D + 1) * D + r] - gc[r] * e) / rho;
  }
  for (int r = 0; r < D; r++) gp[(D + 1) * D + r] = P.TFromE(ge[r]);
}

// calculateViscousRawFlux, ViscousFlux.cpp:59-103.  F is D x Nv column-major.
inline void viscRawFlux(const Phys& P, const double* comp, const double* gp, double* F) {
  const int D = P.D;
  for (int d = 0; d < D; d++) F[d] = 0.0;
  const double T = P.TFromE(comp[D + 1]);
  const double mu = P.mu(T), k = P.kappa(T);
  double tr = 0; for (int d = 0; d < D; d++) tr += gp[(1 + d) * D + d];
  double tau[kMaxD][kMaxD];
  // velocity_gradient(r,c) = gp[(1+c)*D + r]; stress = mu (G + G^T) - 2/3 mu tr I
  for (int r = 0; r < D; r++) for (int c = 0; c < D; c++)
    tau[r][c] = mu * (gp[(1 + c) * D + r] + gp[(1 + r) * D + c]) - (r == c ? 2.0 / 3.0 * mu * tr : 0.0);
  for (int c = 0; c < D; c++) for (int r = 0; r < D; r++) F[(1 + c) * D + r] = tau[r][c];
  for (int r = 0; r < D; r++) {
    double s = 0;
    if (P.comp()) for (int c = 0; c < D; c++) s += tau[r][c] * comp[1 + c];
    F[(D + 1) * D + r] = s + k * gp[(D + 1) * D + r];
  }
}
inline void viscNormalFlux(const Phys& P, const double* n, const double* comp, const double* gp, double* Fn) {  // :116-124
  double F[kMaxD * kMaxV];
  viscRawFlux(P, comp, gp, F);
  for (int v = 0; v < P.Nv; v++) { double s = 0; for (int d = 0; d < P.D; d++) s += F[v * P.D + d] * n[d]; Fn[v] = s; }
}

// BoundaryConditionImpl<...>::calculateBoundaryVariable, BoundaryCondition.cpp:79-547.  Produces the COMPUTATIONAL
// column of the boundary state from the interior (left) state and the user-supplied dummy (right) state.
inline void bcBoundaryVariable(const Phys& P, int bc, const double* n, const Var& L, const Var& R, double* b) {
  const int D = P.D, Nc = D + 3;
  auto dotn = [&](const double* c) { double s = 0; for (int d = 0; d < D; d++) s += c[1 + d] * n[d]; return s; };
  switch (bc) {
    case kRiemannFarfield: {  // :82-285
      const double un = dotn(L.comp);
      const double mach = un / P.sound(L.comp[0], L.comp[D + 2]);
      if (std::fabs(mach) > 1.0) {
        const double* s = mach < 0.0 ? R.comp : L.comp;
        for (int k = 0; k < Nc; k++) b[k] = s[k];
        return;
      }
      const bool inflow = mach < 0.0;
      const Var& S = inflow ? R : L;  // state supplying entropy / tangential velocity
      if (P.comp()) {
        const double g = P.gamma;
        const double Rm = dotn(R.comp) - 2.0 * P.sound(R.comp[0], R.comp[D + 2]) / (g - 1.0);
        const double Rp = dotn(L.comp) + 2.0 * P.sound(L.comp[0], L.comp[D + 2]) / (g - 1.0);
        const double bun = (Rm + Rp) / 2.0;
        const double c = (g - 1.0) * (Rp - Rm) / 4.0;
        const double s = P.entropy(S.comp[0], S.comp[D + 2]);
        const double rho = std::pow(c * c / (g * s), 1.0 / (g - 1.0));
        const double p = rho * c * c / g;
        const double e = p / ((g - 1.0) * rho);
        const double unS = dotn(S.comp);
        b[0] = rho;
        for (int d = 0; d < D; d++) b[1 + d] = S.comp[1 + d] + (bun - unS) * n[d];
        b[D + 1] = e; b[D + 2] = p;
      } else {
        const double c = P.sound(0.0, 0.0);
        const double unL = dotn(L.comp), unR = dotn(R.comp);
        const double rho = std::sqrt(L.comp[0] * R.comp[0] * std::exp((unL - unR) / c));
        const double bun = (unL + unR) / 2.0 + std::log(L.comp[0] / R.comp[0]) * c / 2.0;
        const double e = S.comp[D + 1] * S.comp[0] / rho;
        const double unS = dotn(S.comp);
        b[0] = rho;
        for (int d = 0; d < D; d++) b[1 + d] = S.comp[1 + d] + (bun - unS) * n[d];
        b[D + 1] = e; b[D + 2] = P.pressure(rho, e);
      }
      return;
    }
    case kVelocityInflow: {  // :313-332
      const double mach = dotn(L.comp) / P.sound(L.comp[0], L.comp[D + 2]);
      for (int k = 0; k < Nc; k++) b[k] = R.comp[k];
      if (mach > -1.0) b[D + 2] = L.comp[D + 2];
      return;
    }
    case kPressureOutflow: {  // :360-379
      const double mach = dotn(L.comp) / P.sound(L.comp[0], L.comp[D + 2]);
      for (int k = 0; k < Nc; k++) b[k] = L.comp[k];
      if (mach < 1.0) b[D + 2] = R.comp[D + 2];
      return;
    }
    case kIsoThermalNonSlipWall: {  // :407-424
      b[0] = L.comp[0];
      for (int d = 0; d < D; d++) b[1 + d] = R.comp[1 + d];
      b[D + 1] = R.comp[D + 1];
      b[D + 2] = P.pressure(L.comp[0], R.comp[D + 1]);
      return;
    }
    case kAdiabaticSlipWall: {  // :458-471
      for (int k = 0; k < Nc; k++) b[k] = L.comp[k];
      const double un = dotn(L.comp);
      for (int d = 0; d < D; d++) b[1 + d] = L.comp[1 + d] - un * n[d];
      return;
    }
    case kAdiabaticNonSlipWall: {  // :507-516
      for (int k = 0; k < Nc; k++) b[k] = L.comp[k];
      for (int d = 0; d < D; d++) b[1 + d] = R.comp[1 + d];
      return;
    }
  }
  throw std::runtime_error("oracle: boundary condition type has no boundary-state function (Periodic?)");
}
inline bool bcIsWall(int bc) { return bc == kIsoThermalNonSlipWall || bc == kAdiabaticSlipWall || bc == kAdiabaticNonSlipWall; }

// calculateBoundaryGradientVariable (:287-297 and the wall variants :426-441, :473-488, :518-533)
inline void bcBoundaryGradientVariable(const Phys& P, int bc, const double* n, const Var& L, const Var& R, double* volCons, double* intCons) {
  if (!bcIsWall(bc)) {
    for (int v = 0; v < P.Nv; v++) { volCons[v] = L.cons[v]; intCons[v] = 0.0; }
    return;
  }
  Var b;
  bcBoundaryVariable(P, bc, n, L, R, b.comp);
  consFromComp(P, b);
  for (int v = 0; v < P.Nv; v++) { volCons[v] = b.cons[v]; intCons[v] = b.cons[v] - L.cons[v]; }
}

// SourceTermBase<Boussinesq>::calculateSourceTerm, SourceTerm.cpp:29-58 (gravity = 1 along the last axis)
inline void sourceTerm(const Phys& P, const double* comp, double* S) {
  for (int v = 0; v < P.Nv; v++) S[v] = 0.0;
  if (P.source == kBoussinesq && P.D >= 2) S[P.D] = comp[0] * P.beta * (P.TFromE(comp[P.D + 1]) - P.Tref) * 1.0;
}

}  // namespace orc
