// ref_physics.cpp — C entry point over the REFERENCE'S OWN pointwise physics, compiled from the sources where they lie under
// /root/reference (oracle/Makefile, target `ref`, output oracle/_ref/libref_physics.so).  TEST INFRASTRUCTURE ONLY: it generates
// tests/golden/reference_physics.json (tests/golden/make_reference_physics.py) and is never loaded by the product.
//
// What is the reference's code here and what is not: every arithmetic statement that produces a golden number lives in
// src/Solver/{VariableConvertor,ConvectiveFlux,ViscousFlux,BoundaryCondition,PhysicalModel,SourceTerm}.cpp and is compiled unmodified,
// except for the one g++-rejected construct that oracle/ref_patch.py folds (VariableConvertor.cpp:228-231).  This file only fills the
// reference's Variable / VariableGradient objects, calls the reference's functions in the order of its face loops
// (SpatialDiscrete.cpp:701-745, 792-839), and copies results out.  Eigen, magic_enum, oneTBB, SYCL, dbg-macro and Gmsh are replaced by
// the declaration-level stand-ins of oracle/ref_shim/ (the container has none of them); Eigen's small fixed-size arithmetic is
// restated there as plain loops in expression order, compiled with -ffp-contract=off.
#include <cstdint>
#include <cstring>
#include <type_traits>

#include "Solver/BoundaryCondition.cpp"
#include "Solver/ConvectiveFlux.cpp"
#include "Solver/SourceTerm.cpp"
#include "Solver/ViscousFlux.cpp"

using namespace SubrosaDG;

namespace {

template <int D> struct DimOf;
template <> struct DimOf<1> { static constexpr DimensionEnum v = DimensionEnum::D1; static constexpr MeshModelEnum m = MeshModelEnum::Line; };
template <> struct DimOf<2> { static constexpr DimensionEnum v = DimensionEnum::D2; static constexpr MeshModelEnum m = MeshModelEnum::Quadrangle; };
template <> struct DimOf<3> { static constexpr DimensionEnum v = DimensionEnum::D3; static constexpr MeshModelEnum m = MeshModelEnum::Hexahedron; };

template <int D, SourceTermEnum S>
using SolveC = SolveControl<DimOf<D>::v, PolynomialOrderEnum::P3, BoundaryTimeEnum::Steady, S>;
template <int D>
using NumC = NumericalControl<DimOf<D>::m, ShockCapturingEnum::None, LimiterEnum::None, InitialConditionEnum::Function, TimeIntegrationEnum::SSPRK3>;

struct Params { double cp, cv, mu, c0, rho0, beta, t_ref; };

template <typename SC>
void setModel(PhysicalModel<SC>& pm, const Params& p) {
  pm.thermodynamic_model_.specific_heat_constant_pressure = p.cp;
  pm.thermodynamic_model_.specific_heat_constant_volume = p.cv;
  if constexpr (SC::kEquationOfState == EquationOfStateEnum::WeakCompressibleFluid) {
    pm.equation_of_state_.reference_sound_speed = p.c0;
    pm.equation_of_state_.reference_density = p.rho0;
    pm.equation_of_state_.calculatePressureAdditionFromSoundSpeedDensity();   // System::setEquationOfState, SystemControl.cpp:88-96
  }
  if constexpr (SC::kTransportModel != TransportModelEnum::None) {
    pm.transport_model_.dynamic_viscosity = p.mu;
    pm.calculateThermalConductivityFromDynamicViscosity();                    // System::setTransportModel, SystemControl.cpp:98-103
  }
}

// Face-side objects hold kN = 2 points (the value sits in column 0, column 1 repeats it): with one point per side the reference's two
// calculateViscousFlux overloads (ViscousFlux.cpp:139-170: interior face <N, N>, boundary face <N, 1>) would be ambiguous.
constexpr int kN = 2;

template <typename SC>
void fillConserved(Variable<SC, kN>& v, const double* cons, const PhysicalModel<SC>& pm) {
  for (int c = 0; c < kN; c++) for (int k = 0; k < SC::kConservedVariableNumber; k++) v.conserved_(k, c) = cons[k];
  v.calculateComputationalFromConserved(pm);
}

template <BoundaryConditionEnum B, typename SC>
void boundaryPoint(const PhysicalModel<SC>& pm, const Eigen::Vector<Real, SC::kDimension>& n, Variable<SC, kN>& left, const Variable<SC, kN>& dummy,
                   const double* gradIn, double* out) {
  constexpr int D = SC::kDimension, NV = SC::kConservedVariableNumber, NC = SC::kComputationalVariableNumber;
  using Impl = BoundaryConditionImpl<SC, B>;
  Variable<SC, 1> b, vol, itf;
  Impl::template calculateBoundaryVariable<kN>(pm, n, left, dummy, b, 0);
  Impl::template calculateBoundaryGradientVariable<kN>(pm, n, left, dummy, vol, itf, 0);
  Flux<SC> conv;
  calculateConvectiveNormalFlux(n, b, conv.result_, 0);                                    // SpatialDiscrete.cpp:802-803
  int o = 0;
  for (int k = 0; k < NC; k++) out[o++] = b.computational_(k, 0);
  for (int k = 0; k < NV; k++) out[o++] = vol.conserved_(k, 0);
  for (int k = 0; k < NV; k++) out[o++] = itf.conserved_(k, 0);
  for (int k = 0; k < NV; k++) out[o++] = conv.result_.normal_variable_(k);
  if constexpr (IsNS<SC::kEquationModel>) {
    VariableGradient<SC, kN> lg;
    VariableGradient<SC, 1> bg;
    for (int c = 0; c < kN; c++) for (int k = 0; k < NV * D; k++) lg.conserved_(k, c) = gradIn[k];
    lg.calculatePrimitiveFromConserved(pm, left);                                          // :779-784, from the unmodified interior trace
    Impl::template modifyBoundaryVariable<kN>(left, lg, b, bg, 0);                          // :804-808
    Flux<SC> visc;
    calculateViscousFlux(pm, n, left, lg, b, bg, visc, 0, 0);                               // :809-812
    for (int k = 0; k < NC; k++) out[o++] = left.computational_(k, 0);                     // interior state after modifyBoundaryVariable
    for (int k = 0; k < NV; k++) out[o++] = visc.result_.normal_variable_(k);
  }
}

template <typename SC>
int run(const Params& p, int what, int bc, int n, const double* in, double* out) {
  constexpr int D = SC::kDimension, NV = SC::kConservedVariableNumber, NC = SC::kComputationalVariableNumber;
  PhysicalModel<SC> pm;
  setModel(pm, p);
  for (int i = 0; i < n; i++) {
    if (what == 0) {   // Riemann flux of the configured ConvectiveFluxEnum: in = normal[D], consL[NV], consR[NV]; out = flux[NV]
      const double* a = in + (size_t)i * (D + 2 * NV);
      Eigen::Vector<Real, D> nv;
      for (int d = 0; d < D; d++) nv(d) = a[d];
      Variable<SC, kN> L, R;
      fillConserved(L, a + D, pm); fillConserved(R, a + D + NV, pm);
      Flux<SC> f;
      calculateConvectiveFlux(pm, nv, L, R, f, 0, 0);
      for (int k = 0; k < NV; k++) out[(size_t)i * NV + k] = f.result_.normal_variable_(k);
    } else if (what == 1) {
      // boundary face point: in = normal[D], consL[NV], user primitive (rho, u, T)[NV], conserved gradient trace [NV*D] (row var*D+dir);
      // out = boundary computational state[NC], volume- / interface-gradient conserved states[NV each], convective boundary flux[NV],
      //       NS only: interior computational state after modifyBoundaryVariable[NC], averaged viscous flux[NV]
      const int ni = D + 2 * NV + NV * D, no = NC + 3 * NV + (IsNS<SC::kEquationModel> ? NC + NV : 0);
      const double* a = in + (size_t)i * ni;
      Eigen::Vector<Real, D> nv;
      for (int d = 0; d < D; d++) nv(d) = a[d];
      Variable<SC, kN> L, dummy;
      fillConserved(L, a + D, pm);
      for (int c = 0; c < kN; c++) for (int k = 0; k < NV; k++) dummy.primitive_(k, c) = a[D + NV + k];
      dummy.calculateConservedFromPrimitive(pm);                                            // InitialCondition.cpp:118-149
      dummy.calculateComputationalFromPrimitive(pm);
      double* o = out + (size_t)i * no;
      const double* g = a + D + 2 * NV;
      switch (static_cast<BoundaryConditionEnum>(bc)) {
        case BoundaryConditionEnum::RiemannFarfield: boundaryPoint<BoundaryConditionEnum::RiemannFarfield>(pm, nv, L, dummy, g, o); break;
        case BoundaryConditionEnum::VelocityInflow: boundaryPoint<BoundaryConditionEnum::VelocityInflow>(pm, nv, L, dummy, g, o); break;
        case BoundaryConditionEnum::PressureOutflow: boundaryPoint<BoundaryConditionEnum::PressureOutflow>(pm, nv, L, dummy, g, o); break;
        case BoundaryConditionEnum::IsoThermalNonSlipWall: boundaryPoint<BoundaryConditionEnum::IsoThermalNonSlipWall>(pm, nv, L, dummy, g, o); break;
        case BoundaryConditionEnum::AdiabaticSlipWall: boundaryPoint<BoundaryConditionEnum::AdiabaticSlipWall>(pm, nv, L, dummy, g, o); break;
        case BoundaryConditionEnum::AdiabaticNonSlipWall: boundaryPoint<BoundaryConditionEnum::AdiabaticNonSlipWall>(pm, nv, L, dummy, g, o); break;
        default: return 2;
      }
    } else if (what == 2) {
      // viscous terms at one point: in = normal[D], cons[NV], conserved gradient[NV*D]; out = primitive gradient[NV*D], raw viscous flux
      // [NV*D] (row var*D+dir), normal viscous flux[NV]
      if constexpr (IsNS<SC::kEquationModel>) {
        const int ni = D + NV + NV * D, no = 2 * NV * D + NV;
        const double* a = in + (size_t)i * ni;
        Eigen::Vector<Real, D> nv;
        for (int d = 0; d < D; d++) nv(d) = a[d];
        Variable<SC, kN> V;
        fillConserved(V, a + D, pm);
        VariableGradient<SC, kN> G;
        for (int c = 0; c < kN; c++) for (int k = 0; k < NV * D; k++) G.conserved_(k, c) = a[D + NV + k];
        G.calculatePrimitiveFromConserved(pm, V);
        FluxVariable<SC> raw;
        calculateViscousRawFlux(pm, V, G, raw, 0);
        FluxNormalVariable<SC> nf;
        calculateViscousNormalFlux(pm, nv, V, G, nf, 0);
        double* o = out + (size_t)i * no;
        for (int k = 0; k < NV * D; k++) o[k] = G.primitive_(k, 0);
        for (int v = 0; v < NV; v++) for (int d = 0; d < D; d++) o[NV * D + v * D + d] = raw.variable_(d, v);
        for (int k = 0; k < NV; k++) o[2 * NV * D + k] = nf.normal_variable_(k);
      } else {
        return 3;
      }
    } else if (what == 3) {
      // raw convective flux and source term at one point: in = cons[NV]; out = computational[NC], primitive[NV], raw flux[NV*D], source[NV]
      const int no = NC + NV + NV * D + NV;
      Variable<SC, kN> V;
      fillConserved(V, in + (size_t)i * NV, pm);
      V.calculatePrimitiveFromConserved(pm);
      FluxVariable<SC> raw;
      calculateConvectiveRawFlux(V, raw, 0);
      double* o = out + (size_t)i * no;
      for (int k = 0; k < NC; k++) o[k] = V.computational_(k, 0);
      for (int k = 0; k < NV; k++) o[NC + k] = V.primitive_(k, 0);
      for (int v = 0; v < NV; v++) for (int d = 0; d < D; d++) o[NC + NV + v * D + d] = raw.variable_(d, v);
      for (int k = 0; k < NV; k++) o[NC + NV + NV * D + k] = 0.0;
      if constexpr (SC::kSourceTerm == SourceTermEnum::Boussinesq) {
        SourceTerm<SC> st;
        st.thermal_expansion_coefficient = p.beta; st.reference_temperature = p.t_ref;
        FluxNormalVariable<SC> s;
        st.calculateSourceTerm(pm, V, s, 0);
        for (int k = 0; k < NV; k++) o[NC + NV + NV * D + k] = s.normal_variable_(k);
      }
    } else if (what == 4) {
      // ViewVariable::get (VariableConvertor.cpp:754-872) as ElementViewSolver::calcluateElementViewVariable fills it (RawBinary.cpp:193-240):
      // in = cons[NV], conserved gradient[NV*D] (row var*D+dir), artificial viscosity; out = the 22 ViewVariableEnum values (0 where the
      // variable names a direction the dimension does not have: the reference never asks for those)
      using ET = std::conditional_t<D == 1, LineTrait<3>, std::conditional_t<D == 2, QuadrangleTrait<3>, HexahedronTrait<3>>>;
      constexpr int NB = ET::kBasisFunctionNumber;
      const int ni = NV + NV * D + 1, no = 22;
      const double* a = in + (size_t)i * ni;
      ViewVariable<ET, SC> vv;
      for (int c = 0; c < NB; c++) for (int k = 0; k < NV; k++) vv.variable_.conserved_(k, c) = a[k];
      vv.variable_.calculateComputationalFromConserved(pm);
      if constexpr (IsNS<SC::kEquationModel>) {
        for (int c = 0; c < NB; c++) for (int k = 0; k < NV * D; k++) vv.variable_gradient_.conserved_(k, c) = a[NV + k];
        vv.variable_gradient_.calculatePrimitiveFromConserved(pm, vv.variable_);            // RawBinary.cpp:228-232
      }
      for (int c = 0; c < NB; c++) vv.artificial_viscosity_(c) = a[NV + NV * D];
      double* o = out + (size_t)i * no;
      for (int w = 0; w < no; w++) {
        const bool needs3 = w == 12 || w == 15 || w == 16 || w == 17 || w == 21, needs2 = w == 11 || w == 14 || w == 18 || w == 20;
        o[w] = ((needs3 && D < 3) || (needs2 && D < 2)) ? 0.0 : vv.get(pm, static_cast<ViewVariableEnum>(w), 0);
      }
    } else {
      return 4;
    }
  }
  return 0;
}

constexpr auto TC = ThermodynamicModelEnum::Constant;
constexpr auto IG = EquationOfStateEnum::IdealGas;
constexpr auto WC = EquationOfStateEnum::WeakCompressibleFluid;

template <int D, SourceTermEnum S>
int dispatchModel(int model, int eos, int transport, int conv, const Params& p, int what, int bc, int n, const double* in, double* out) {
  using SV = SolveC<D, S>;
  using NC = NumC<D>;
#define RUN(...) return run<SimulationControl<SV, NC, __VA_ARGS__>>(p, what, bc, n, in, out)
  if (model == 0 && eos == 0) {
    if (conv == 0) RUN(CompresibleEulerVariable<TC, IG, ConvectiveFluxEnum::Central>);
    if (conv == 1) RUN(CompresibleEulerVariable<TC, IG, ConvectiveFluxEnum::LaxFriedrichs>);
    if (conv == 2) RUN(CompresibleEulerVariable<TC, IG, ConvectiveFluxEnum::HLLC>);
    if (conv == 3) RUN(CompresibleEulerVariable<TC, IG, ConvectiveFluxEnum::Roe>);
  }
  if (model == 2 && eos == 1) {
    if (conv == 0) RUN(IncompresibleEulerVariable<TC, WC, ConvectiveFluxEnum::Central>);
    if (conv == 1) RUN(IncompresibleEulerVariable<TC, WC, ConvectiveFluxEnum::LaxFriedrichs>);
    if (conv == 4) RUN(IncompresibleEulerVariable<TC, WC, ConvectiveFluxEnum::Exact>);
  }
  if (model == 1 && eos == 0 && conv == 2) {
    if (transport == 1) RUN(CompresibleNSVariable<TC, IG, TransportModelEnum::Constant, ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>);
    if (transport == 2) RUN(CompresibleNSVariable<TC, IG, TransportModelEnum::Sutherland, ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>);
  }
  if (model == 3 && eos == 1 && conv == 4 && transport == 1)
    RUN(IncompresibleNSVariable<TC, WC, TransportModelEnum::Constant, ConvectiveFluxEnum::Exact, ViscousFluxEnum::BR2>);
#undef RUN
  return 1;
}

}  // namespace

extern "C" {

// cfg = {dim, model, eos, transport, conv_flux, source} with the integer values of src/Utils/Enum.cpp; params = {cp, cv, mu, c0, rho0,
// beta, t_ref}.  Returns 0, or 1 for a configuration this driver does not instantiate.
int ref_physics(const int32_t* cfg, const double* params, int what, int bc, int n, const double* in, double* out) {
  Params p;
  std::memcpy(&p, params, sizeof(p));
  const int dim = cfg[0], model = cfg[1], eos = cfg[2], transport = cfg[3], conv = cfg[4], source = cfg[5];
  if (source == 1) {
    if (dim == 2) return dispatchModel<2, SourceTermEnum::Boussinesq>(model, eos, transport, conv, p, what, bc, n, in, out);
    if (dim == 3) return dispatchModel<3, SourceTermEnum::Boussinesq>(model, eos, transport, conv, p, what, bc, n, in, out);
    return 1;
  }
  if (dim == 1) return dispatchModel<1, SourceTermEnum::None>(model, eos, transport, conv, p, what, bc, n, in, out);
  if (dim == 2) return dispatchModel<2, SourceTermEnum::None>(model, eos, transport, conv, p, what, bc, n, in, out);
  if (dim == 3) return dispatchModel<3, SourceTermEnum::None>(model, eos, transport, conv, p, what, bc, n, in, out);
  return 1;
}

}  // extern "C"
