// ref_sweeps.cpp — the REFERENCE'S OWN solver driven from hand-filled mesh objects.  TEST INFRASTRUCTURE ONLY: it generates
// tests/golden/reference_sweeps.json (tests/golden/make_reference_sweeps.py) and is never loaded by the product.
//
// What runs here is the reference's code, compiled from where it lies under /root/reference/src (oracle/Makefile, target `ref`):
//   * ElementBasisFunction / AdjacencyElementBasisFunction / ElementQuadrature constructors (src/Mesh/BasisFunction.cpp, Quadrature.cpp):
//     they assemble modal_value_, modal_gradient_value_, modal_adjacency_value_, modal_least_squares_inverse_, nodal_* from the three Gmsh
//     calls they make — answered by oracle/ref_run/gmsh.h from the repository's restatement of Gmsh's tables;
//   * Solver<SC>::initializeSolver (InitialCondition.cpp:85-186), calculateDeltaTime (TimeIntegration.cpp:104-179) and stepSolver
//     (TimeIntegration.cpp:326-350: all eight sweeps of SpatialDiscrete.cpp, the RK update, the relative error) — unmodified.
// What is NOT the reference's: the mesh reader and the geometry (src/Mesh/ReadControl.cpp, Adjacency.cpp, Geometry.cpp need a live Gmsh
// model).  The Mesh<SC> object is filled from arrays: adjacency records as the C ABI takes them, and per element / face the geometric
// factors computed by the oracle (quadrature coordinates, detJ w, (J^T)^-1 detJ w, M^-1, minimum edge, normals, |J| w).  The golden
// numbers therefore pin the ASSEMBLY around the physics — projection, sweeps, scatter, RK, norm — given the same geometric factors.
// Eigen / oneTBB / magic_enum / zstd / vtu11 are the stand-ins of oracle/ref_shim (plain-loop arithmetic, -ffp-contract=off).
#include <cstdint>
#include <cstring>
#include <string>

#include "Mesh/BasisFunction.cpp"
#include "Mesh/Quadrature.cpp"
#include "Mesh/ReadControl.cpp"
#include "Solver/BoundaryCondition.cpp"
#include "Solver/InitialCondition.cpp"
#include "Solver/SolveControl.cpp"
#include "Solver/SourceTerm.cpp"
#include "Solver/SpatialDiscrete.cpp"
#include "Solver/TimeIntegration.cpp"
#include "View/RawBinary.cpp"   // Solver<SC>::writeRawBinary + RawBinaryCompress::write (:42-57, :75-191); linked against the system's libzstd.so.1

using namespace SubrosaDG;

namespace {

struct Params {
  double cp, cv, mu, amp, vel[3]; double jump_width, jump_radius, av_tolerance, av_factor;
  double c0, rho0, beta, t_ref;   // EquationOfState<WeakCompressibleFluid> (PhysicalModel.cpp:57-78), SourceTermBase<Boussinesq> (SourceTerm.cpp:30-33)
  int weak;                       // 1: the weakly compressible field below (incompressible examples' variable set)
  double time_rate;               // BoundaryTimeEnum::TimeVarying: the boundary velocity grows like 1 + time_rate * t
};
Params g_params;
thread_local std::string g_error;
std::string g_raw_path;   // non-empty: the reference's writeRawBinary writes its .zst file there after the last step

// the analytic fields of tests/cases.py::ic_perturbed_freestream / bc_freestream (free stream rho = 1.4, T = 1, velocity vel, times a
// smooth perturbation 1 + amp sin(pi x) cos(pi y) [cos(pi z)]; the boundary callback returns the unperturbed free stream)
// shock-capturing cases: a steep density / pressure jump, fluid at rest — across the oblique plane sum_d x_d / sqrt(D) = 0.5 sqrt(D)
// (jump_radius == 0) or across the circle |x| = jump_radius; tests/test_gpu_av.py::jump_ic / radial_jump_ic
template <int D>
Eigen::Vector<Real, D + 2> jumpAt(const Eigen::Vector<Real, D>& x) {
  double a = 0.0;
  if (g_params.jump_radius > 0.0) {
    for (int d = 0; d < D; d++) a += x(d) * x(d);
    a = (std::sqrt(a) - g_params.jump_radius) / g_params.jump_width;
  } else {
    const double c = 1.0 / std::sqrt(static_cast<double>(D));
    for (int d = 0; d < D; d++) a += x(d) * c;
    a = (a - 0.5 * c * D) / g_params.jump_width;
  }
  const double s = std::tanh(a);
  const double rho = g_params.jump_radius > 0.0 ? 0.75 - 0.25 * s : 0.5625 - 0.4375 * s;
  const double p = g_params.jump_radius > 0.0 ? 0.75 - 0.25 * s : 0.55 - 0.45 * s;
  Eigen::Vector<Real, D + 2> q;
  q(0) = rho;
  for (int d = 0; d < D; d++) q(1 + d) = 0.0;
  q(D + 1) = 1.4 * p / rho;
  return q;
}

template <int D>
Eigen::Vector<Real, D + 2> fieldAt(const Eigen::Vector<Real, D>& x, const double amp) {
  if (g_params.jump_width > 0.0) return jumpAt<D>(x);
  double s = std::sin(kPi * x(0));
  if constexpr (D >= 2) s *= std::cos(kPi * x(1));
  if constexpr (D >= 3) s *= std::cos(kPi * x(2));
  const double g = 1.0 + amp * s;
  Eigen::Vector<Real, D + 2> p;
  if (g_params.weak) {   // density close to the reference density (p = c0^2 (rho - rho0) + p_add), smooth velocity, temperature field for the buoyancy
    p(0) = g_params.rho0 * (1.0 + 0.01 * amp * s);
    for (int d = 0; d < D; d++) p(1 + d) = g_params.vel[d] * g;
    p(D + 1) = 1.0 + 2.0 * amp * s;
    return p;
  }
  p(0) = 1.4 * g;
  for (int d = 0; d < D; d++) p(1 + d) = g_params.vel[d] * g + 0.0 * s;
  p(D + 1) = 1.0 * g;
  return p;
}

}  // namespace

template <typename SimulationControl>
inline Eigen::Vector<Real, SimulationControl::kPrimitiveVariableNumber> SubrosaDG::InitialCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate) const {
  return fieldAt<SimulationControl::kDimension>(coordinate, g_params.amp);
}
template <typename SimulationControl>
inline Eigen::Vector<Real, SimulationControl::kPrimitiveVariableNumber> SubrosaDG::BoundaryCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate, [[maybe_unused]] const Isize gmsh_physical_index) const {
  return fieldAt<SimulationControl::kDimension>(coordinate, 0.0);
}

// BoundaryTimeEnum::TimeVarying: Solver::updateBoundaryVariable calls this overload with t = iteration_ * delta_time_ at the start of every
// stepSolver (BoundaryCondition.cpp:29-74, TimeIntegration.cpp:332-334)
template <typename SimulationControl>
inline Eigen::Vector<Real, SimulationControl::kPrimitiveVariableNumber> SubrosaDG::BoundaryCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate, const Real time, [[maybe_unused]] const Isize gmsh_physical_index) const {
  Eigen::Vector<Real, SimulationControl::kPrimitiveVariableNumber> p = fieldAt<SimulationControl::kDimension>(coordinate, 0.0);
  for (int d = 0; d < SimulationControl::kDimension; d++) p(1 + d) = p(1 + d) * (1.0 + g_params.time_rate * time);
  return p;
}

namespace {

// per element type: what the oracle's geometry getters hand out (oracle/__init__.py: element_geometry 0..4)
struct BlockIn { int type, n; const double *xq, *jw, *mt, *minv, *min_edge; double* coef_out; const int32_t* node_tag; const double* inner_radius; };
struct FacesIn {
  int n_int, n_bnd;
  const int32_t *le, *lt, *lf, *re, *rt, *rf, *rot, *bc, *phys;
  const double *xf, *nrm, *jw;
};

template <typename ElementTrait, int D>
void fillElementMesh(ElementMesh<ElementTrait>& em, const BlockIn& b) {
  constexpr int Nq = ElementTrait::kQuadratureNumber, Nb = ElementTrait::kBasisFunctionNumber;
  em.number_ = b.n;
  em.element_.resize(b.n);
  for (Isize i = 0; i < b.n; i++) {
    auto& e = em.element_(i);
    for (int q = 0; q < Nq; q++) {
      for (int d = 0; d < D; d++) e.quadrature_node_coordinate_(d, q) = b.xq[((std::size_t)i * Nq + q) * D + d];
      e.jacobian_determinant_mutiply_weight_(q) = b.jw[(std::size_t)i * Nq + q];
      for (int k = 0; k < D * D; k++) e.jacobian_transpose_inverse_mutiply_deteminate_and_weight_(k, q) = b.mt[((std::size_t)i * Nq + q) * D * D + k];
    }
    for (int r = 0; r < Nb; r++) for (int c = 0; c < Nb; c++) e.local_mass_matrix_inverse_(r, c) = b.minv[((std::size_t)i * Nb + c) * Nb + r];   // column-major
    e.minimum_edge_ = b.min_edge[i];
    e.inner_radius_ = b.inner_radius ? b.inner_radius[i] : 0.0;
    if (b.node_tag) for (int k = 0; k < ElementTrait::kBasicNodeNumber; k++) e.node_tag_(k) = b.node_tag[(std::size_t)i * ElementTrait::kBasicNodeNumber + k] + 1;   // Gmsh tags are 1-based
  }
}

template <typename AdjacencyElementTrait, int D, int P>
void fillAdjacencyMesh(AdjacencyElementMesh<AdjacencyElementTrait>& am, const FacesIn& f) {
  constexpr int Nqf = AdjacencyElementTrait::kQuadratureNumber;
  static const int line[5] = {1, 8, 26, 27, 28}, tri[5] = {2, 9, 21, 23, 25}, quad[5] = {3, 10, 36, 37, 38}, hex[5] = {5, 12, 92, 93, 94};
  auto gmshType = [&](int t) { return t == 1 ? line[P - 1] : t == 2 ? tri[P - 1] : t == 3 ? quad[P - 1] : hex[P - 1]; };   // SimulationControl.cpp:26-31
  const int nf = f.n_int + f.n_bnd;
  am.interior_number_ = f.n_int; am.boundary_number_ = f.n_bnd;
  am.element_.resize(nf);
  for (Isize i = 0; i < nf; i++) {
    auto& a = am.element_(i);
    const bool interior = i < f.n_int;
    a.parent_index_each_type_(0) = f.le[i]; a.adjacency_sequence_in_parent_(0) = f.lf[i]; a.parent_gmsh_type_number_(0) = gmshType(f.lt[i]);
    a.parent_index_each_type_(1) = interior ? f.re[i] : 0; a.adjacency_sequence_in_parent_(1) = interior ? f.rf[i] : 0;
    a.parent_gmsh_type_number_(1) = interior ? gmshType(f.rt[i]) : 0;
    a.adjacency_right_rotation_ = interior ? f.rot[i] : 0;
    a.boundary_condition_type_ = static_cast<BoundaryConditionEnum>(f.bc[i]);
    a.gmsh_physical_index_ = f.phys[i];
    for (int j = 0; j < Nqf; j++) {
      for (int d = 0; d < D; d++) {
        a.quadrature_node_coordinate_(d, j) = f.xf[((std::size_t)i * Nqf + j) * D + d];
        a.normal_vector_(d, j) = f.nrm[((std::size_t)i * Nqf + j) * D + d];
      }
      a.jacobian_determinant_mutiply_weight_(j) = f.jw[(std::size_t)i * Nqf + j];
    }
  }
}

template <typename SC>
int runCase(const BlockIn* blocks, int n_blocks, const FacesIn& faces, int nsteps, double cfl, double dt_in, double* relerr_out, double* dt_out, int node_number,
            double* node_av_out) {
  constexpr int D = SC::kDimension, P = SC::kPolynomialOrder;
  auto mesh_p = std::make_unique<Mesh<SC>>();   // constructors assemble the reference-element tables (through the Gmsh stand-in)
  Mesh<SC>& mesh = *mesh_p;
  PhysicalModel<SC> physical_model;
  physical_model.thermodynamic_model_.specific_heat_constant_pressure = g_params.cp;
  physical_model.thermodynamic_model_.specific_heat_constant_volume = g_params.cv;
  if constexpr (SC::kEquationOfState == EquationOfStateEnum::WeakCompressibleFluid) {   // System::setEquationOfState, SystemControl.cpp
    physical_model.equation_of_state_.reference_sound_speed = g_params.c0;
    physical_model.equation_of_state_.reference_density = g_params.rho0;
    physical_model.equation_of_state_.calculatePressureAdditionFromSoundSpeedDensity();
  }
  if constexpr (SC::kTransportModel != TransportModelEnum::None) {
    physical_model.transport_model_.dynamic_viscosity = g_params.mu;
    physical_model.calculateThermalConductivityFromDynamicViscosity();
  }
  mesh.element_number_ = 0;
  for (int k = 0; k < n_blocks; k++) {
    const BlockIn& b = blocks[k];
    mesh.element_number_ += b.n;
    if constexpr (D == 1) { fillElementMesh<LineTrait<P>, D>(mesh.line_, b); }
    else if constexpr (D == 2) {
      if constexpr (HasTriangle<SC::kMeshModel>) { if (b.type == 2) fillElementMesh<TriangleTrait<P>, D>(mesh.triangle_, b); }
      if constexpr (HasQuadrangle<SC::kMeshModel>) { if (b.type == 3) fillElementMesh<QuadrangleTrait<P>, D>(mesh.quadrangle_, b); }
    } else { fillElementMesh<HexahedronTrait<P>, D>(mesh.hexahedron_, b); }
  }
  if constexpr (D == 1) fillAdjacencyMesh<AdjacencyPointTrait<P>, D, P>(mesh.point_, faces);
  else if constexpr (D == 2) fillAdjacencyMesh<AdjacencyLineTrait<P>, D, P>(mesh.line_, faces);
  else fillAdjacencyMesh<AdjacencyQuadrangleTrait<P>, D, P>(mesh.quadrangle_, faces);
  mesh.node_number_ = node_number;

  BoundaryCondition<SC> boundary_condition;
  InitialCondition<SC> initial_condition;
  SourceTerm<SC> source_term;
  if constexpr (SC::kSourceTerm == SourceTermEnum::Boussinesq) {   // System::setSourceTerm
    source_term.thermal_expansion_coefficient = g_params.beta;
    source_term.reference_temperature = g_params.t_ref;
  }
  TimeIntegration<SC> time_integration;
  auto solver_p = std::make_unique<Solver<SC>>();
  Solver<SC>& solver = *solver_p;
  solver.empirical_tolerance_ = g_params.av_tolerance;          // System::setArtificialViscosity, SystemControl.cpp:105-108
  solver.artificial_viscosity_factor_ = g_params.av_factor;
  solver.initializeSolver(mesh, physical_model, boundary_condition, initial_condition);
  time_integration.courant_friedrichs_lewy_number_ = cfl;
  time_integration.delta_time_ = dt_in;
  if (dt_in == 0.0) solver.calculateDeltaTime(mesh, physical_model, time_integration);   // System::solve, SystemControl.cpp:163-165
  *dt_out = time_integration.delta_time_;
  for (int i = 1; i <= nsteps; i++) {
    solver.stepSolver(mesh, source_term, physical_model, boundary_condition, time_integration);
    time_integration.iteration_ = i;
  }
  if (!g_raw_path.empty()) {   // System::solve at an output step, SystemControl.cpp:178-183
    solver.writeRawBinary(mesh, g_raw_path);
    solver.write_raw_binary_future_.get();
  }
  for (int v = 0; v < SC::kConservedVariableNumber; v++) relerr_out[v] = solver.relative_error_(v);
  if (node_av_out) for (int k = 0; k < node_number; k++) node_av_out[k] = solver.node_artificial_viscosity_(k);
  for (int k = 0; k < n_blocks; k++) {
    const BlockIn& b = blocks[k];
    auto copyOut = [&](const auto& element_solver) {
      using Coef = std::remove_cvref_t<decltype(element_solver.element_(0).variable_basis_function_coefficient_)>;
      for (Isize i = 0; i < b.n; i++) std::memcpy(b.coef_out + (std::size_t)i * Coef::size(), element_solver.element_(i).variable_basis_function_coefficient_.data(), sizeof(double) * Coef::size());
    };
    if constexpr (D == 1) copyOut(solver.line_);
    else if constexpr (D == 2) {
      if constexpr (HasTriangle<SC::kMeshModel>) { if (b.type == 2) copyOut(solver.triangle_); }
      if constexpr (HasQuadrangle<SC::kMeshModel>) { if (b.type == 3) copyOut(solver.quadrangle_); }
    } else copyOut(solver.hexahedron_);
  }
  return 0;
}

template <DimensionEnum D, PolynomialOrderEnum P, MeshModelEnum M, TimeIntegrationEnum RK, typename Variable, ShockCapturingEnum S = ShockCapturingEnum::None,
          SourceTermEnum Src = SourceTermEnum::None, BoundaryTimeEnum BT = BoundaryTimeEnum::Steady>
using Control = SimulationControl<SolveControl<D, P, BT, Src>,
                                  NumericalControl<M, S, LimiterEnum::None, InitialConditionEnum::Function, RK>, Variable>;
template <ConvectiveFluxEnum F>
using IncEuler = IncompresibleEulerVariable<ThermodynamicModelEnum::Constant, EquationOfStateEnum::WeakCompressibleFluid, F>;
template <ConvectiveFluxEnum F, ViscousFluxEnum V>
using IncNS = IncompresibleNSVariable<ThermodynamicModelEnum::Constant, EquationOfStateEnum::WeakCompressibleFluid, TransportModelEnum::Constant, F, V>;
template <ConvectiveFluxEnum F>
using Euler = CompresibleEulerVariable<ThermodynamicModelEnum::Constant, EquationOfStateEnum::IdealGas, F>;
template <TransportModelEnum T, ConvectiveFluxEnum F, ViscousFluxEnum V>
using NS = CompresibleNSVariable<ThermodynamicModelEnum::Constant, EquationOfStateEnum::IdealGas, T, F, V>;

}  // namespace

extern "C" {

const char* ref_sweeps_error() { return g_error.c_str(); }
void ref_sweeps_set_raw_path(const char* path) { g_raw_path = path ? path : ""; }

// case_id selects one of the compiled control types (tests/golden/make_reference_sweeps.py lists them with their meshes);
// params = {cp, cv, mu, amp, vel[3], jump_width, jump_radius, av_tolerance, av_factor, c0, rho0, beta, t_ref, weak, time_rate}; node tags (0-based) / inner radii for the shock cases; blocks / faces: see BlockIn / FacesIn (faces in the order and meaning of sdg_set_faces)
int ref_sweeps(int case_id, const double* params, int n_blocks, const int32_t* types, const int32_t* counts, const double* const* xq, const double* const* jw,
               const double* const* mt, const double* const* minv, const double* const* min_edge, int n_int, int n_bnd, const int32_t* const* face_int /* 9 arrays */,
               const double* xf, const double* nrm, const double* fjw, int nsteps, double cfl, double dt_in, double* const* coef_out, double* relerr_out,
               double* dt_out, int node_number, const int32_t* const* node_tag, const double* const* inner_radius, double* node_av_out) {
  try {
    g_params = Params{params[0], params[1], params[2], params[3], {params[4], params[5], params[6]}, params[7], params[8], params[9], params[10],
                      params[11], params[12], params[13], params[14], params[15] != 0.0, params[16]};
    BlockIn blocks[4];
    for (int k = 0; k < n_blocks; k++) blocks[k] = BlockIn{types[k], counts[k], xq[k], jw[k], mt[k], minv[k], min_edge[k], coef_out[k], node_tag ? node_tag[k] : nullptr, inner_radius ? inner_radius[k] : nullptr};
    const FacesIn F{n_int, n_bnd, face_int[0], face_int[1], face_int[2], face_int[3], face_int[4], face_int[5], face_int[6], face_int[7], face_int[8], xf, nrm, fjw};
    using enum DimensionEnum; using enum PolynomialOrderEnum; using enum MeshModelEnum; using enum TimeIntegrationEnum;
    switch (case_id) {
      case 0: return runCase<Control<D2, P2, Quadrangle, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 1: return runCase<Control<D2, P3, Quadrangle, SSPRK3, NS<TransportModelEnum::Sutherland, ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 2: return runCase<Control<D2, P2, Quadrangle, HeunRK2, NS<TransportModelEnum::Constant, ConvectiveFluxEnum::Roe, ViscousFluxEnum::BR1>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 3: return runCase<Control<D1, P3, Line, ForwardEuler, Euler<ConvectiveFluxEnum::LaxFriedrichs>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 4: return runCase<Control<D3, P2, Hexahedron, SSPRK3, NS<TransportModelEnum::Constant, ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 5: return runCase<Control<D2, P2, Triangle, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 6: return runCase<Control<D2, P3, TriangleQuadrangle, SSPRK3, NS<TransportModelEnum::Sutherland, ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 7: return runCase<Control<D3, P3, Hexahedron, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      // ShockCapturingEnum::ArtificialViscosity: sod_1d_ceuler / sedovblast_2d_ceuler / explosion_2d_ceuler / cylinder_2d_ceuler control types
      case 8: return runCase<Control<D1, P2, Line, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>, ShockCapturingEnum::ArtificialViscosity>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 9: return runCase<Control<D2, P3, Quadrangle, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>, ShockCapturingEnum::ArtificialViscosity>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 10: return runCase<Control<D2, P2, Triangle, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>, ShockCapturingEnum::ArtificialViscosity>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 11: return runCase<Control<D2, P3, TriangleQuadrangle, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>, ShockCapturingEnum::ArtificialViscosity>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 12: return runCase<Control<D3, P2, Hexahedron, SSPRK3, Euler<ConvectiveFluxEnum::Roe>, ShockCapturingEnum::ArtificialViscosity>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      // the incompressible (weakly compressible) examples: lidcavity_2d / cylinder_2d / kovasznay_2d_incns, thermalcavity_2d_incns (Boussinesq),
      // shearlayer_2d_inceuler, lidcavity_3d / cylinder_3d / square_3d_incns, thermalcavity_3d_incns (sphere_3d_incns without the source), taylorvortex_2d_incns
      case 13: return runCase<Control<D2, P3, Quadrangle, SSPRK3, IncNS<ConvectiveFluxEnum::LaxFriedrichs, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 14: return runCase<Control<D2, P1, Quadrangle, SSPRK3, IncNS<ConvectiveFluxEnum::Exact, ViscousFluxEnum::BR2>, ShockCapturingEnum::None, SourceTermEnum::Boussinesq>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 15: return runCase<Control<D2, P1, Quadrangle, SSPRK3, IncEuler<ConvectiveFluxEnum::LaxFriedrichs>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 16: return runCase<Control<D3, P1, Hexahedron, SSPRK3, IncNS<ConvectiveFluxEnum::LaxFriedrichs, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 17: return runCase<Control<D3, P3, Hexahedron, SSPRK3, IncNS<ConvectiveFluxEnum::Exact, ViscousFluxEnum::BR2>, ShockCapturingEnum::None, SourceTermEnum::Boussinesq>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 18: return runCase<Control<D2, P4, Quadrangle, SSPRK3, IncNS<ConvectiveFluxEnum::Exact, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      // the remaining compressible control types of examples/: sphere_3d_cns (the north-star kernel family: P3 hexahedra, NS, BR2), blasius_3d / delta_3d_cns,
      // rae2822_2d_cns (P5), khinstability_2d_ceuler (P5 + artificial viscosity), sod_1d / shuosher_1d_ceuler (P3 + artificial viscosity)
      case 19: return runCase<Control<D3, P3, Hexahedron, SSPRK3, NS<TransportModelEnum::Constant, ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 20: return runCase<Control<D3, P1, Hexahedron, SSPRK3, NS<TransportModelEnum::Constant, ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 21: return runCase<Control<D2, P5, Quadrangle, SSPRK3, NS<TransportModelEnum::Sutherland, ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 22: return runCase<Control<D2, P5, Quadrangle, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>, ShockCapturingEnum::ArtificialViscosity>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      case 23: return runCase<Control<D1, P3, Line, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>, ShockCapturingEnum::ArtificialViscosity>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      // BoundaryTimeEnum::TimeVarying (the alternative control type of cylinder_2d_incns.cpp:31-41): boundary values re-evaluated before every step
      case 24: return runCase<Control<D2, P2, Quadrangle, SSPRK3, IncNS<ConvectiveFluxEnum::LaxFriedrichs, ViscousFluxEnum::BR2>, ShockCapturingEnum::None, SourceTermEnum::None,
                                      BoundaryTimeEnum::TimeVarying>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out, node_number, node_av_out);
      default: throw std::runtime_error("ref_sweeps: unknown case");
    }
  } catch (const std::exception& ex) {
    g_error = ex.what();
    return 1;
  }
}

}  // extern "C"
