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

using namespace SubrosaDG;

namespace {

struct Params { double cp, cv, mu, amp, vel[3]; };
Params g_params;
thread_local std::string g_error;

// the analytic fields of tests/cases.py::ic_perturbed_freestream / bc_freestream (free stream rho = 1.4, T = 1, velocity vel, times a
// smooth perturbation 1 + amp sin(pi x) cos(pi y) [cos(pi z)]; the boundary callback returns the unperturbed free stream)
template <int D>
Eigen::Vector<Real, D + 2> fieldAt(const Eigen::Vector<Real, D>& x, const double amp) {
  double s = std::sin(kPi * x(0));
  if constexpr (D >= 2) s *= std::cos(kPi * x(1));
  if constexpr (D >= 3) s *= std::cos(kPi * x(2));
  const double g = 1.0 + amp * s;
  Eigen::Vector<Real, D + 2> p;
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

namespace {

// per element type: what the oracle's geometry getters hand out (oracle/__init__.py: element_geometry 0..4)
struct BlockIn { int type, n; const double *xq, *jw, *mt, *minv, *min_edge; double* coef_out; };
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
    e.inner_radius_ = 0.0;
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
int runCase(const BlockIn* blocks, int n_blocks, const FacesIn& faces, int nsteps, double cfl, double dt_in, double* relerr_out, double* dt_out) {
  constexpr int D = SC::kDimension, P = SC::kPolynomialOrder;
  auto mesh_p = std::make_unique<Mesh<SC>>();   // constructors assemble the reference-element tables (through the Gmsh stand-in)
  Mesh<SC>& mesh = *mesh_p;
  PhysicalModel<SC> physical_model;
  physical_model.thermodynamic_model_.specific_heat_constant_pressure = g_params.cp;
  physical_model.thermodynamic_model_.specific_heat_constant_volume = g_params.cv;
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
  mesh.node_number_ = 1;

  BoundaryCondition<SC> boundary_condition;
  InitialCondition<SC> initial_condition;
  SourceTerm<SC> source_term;
  TimeIntegration<SC> time_integration;
  auto solver_p = std::make_unique<Solver<SC>>();
  Solver<SC>& solver = *solver_p;
  solver.initializeSolver(mesh, physical_model, boundary_condition, initial_condition);
  time_integration.courant_friedrichs_lewy_number_ = cfl;
  time_integration.delta_time_ = dt_in;
  if (dt_in == 0.0) solver.calculateDeltaTime(mesh, physical_model, time_integration);   // System::solve, SystemControl.cpp:163-165
  *dt_out = time_integration.delta_time_;
  for (int i = 1; i <= nsteps; i++) {
    solver.stepSolver(mesh, source_term, physical_model, boundary_condition, time_integration);
    time_integration.iteration_ = i;
  }
  for (int v = 0; v < SC::kConservedVariableNumber; v++) relerr_out[v] = solver.relative_error_(v);
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

template <DimensionEnum D, PolynomialOrderEnum P, MeshModelEnum M, TimeIntegrationEnum RK, typename Variable>
using Control = SimulationControl<SolveControl<D, P, BoundaryTimeEnum::Steady, SourceTermEnum::None>,
                                  NumericalControl<M, ShockCapturingEnum::None, LimiterEnum::None, InitialConditionEnum::Function, RK>, Variable>;
template <ConvectiveFluxEnum F>
using Euler = CompresibleEulerVariable<ThermodynamicModelEnum::Constant, EquationOfStateEnum::IdealGas, F>;
template <TransportModelEnum T, ConvectiveFluxEnum F, ViscousFluxEnum V>
using NS = CompresibleNSVariable<ThermodynamicModelEnum::Constant, EquationOfStateEnum::IdealGas, T, F, V>;

}  // namespace

extern "C" {

const char* ref_sweeps_error() { return g_error.c_str(); }

// case_id selects one of the compiled control types (tests/golden/make_reference_sweeps.py lists them with their meshes);
// params = {cp, cv, mu, amp, vel[3]}; blocks / faces: see BlockIn / FacesIn (faces in the order and meaning of sdg_set_faces)
int ref_sweeps(int case_id, const double* params, int n_blocks, const int32_t* types, const int32_t* counts, const double* const* xq, const double* const* jw,
               const double* const* mt, const double* const* minv, const double* const* min_edge, int n_int, int n_bnd, const int32_t* const* face_int /* 9 arrays */,
               const double* xf, const double* nrm, const double* fjw, int nsteps, double cfl, double dt_in, double* const* coef_out, double* relerr_out,
               double* dt_out) {
  try {
    g_params = Params{params[0], params[1], params[2], params[3], {params[4], params[5], params[6]}};
    BlockIn blocks[4];
    for (int k = 0; k < n_blocks; k++) blocks[k] = BlockIn{types[k], counts[k], xq[k], jw[k], mt[k], minv[k], min_edge[k], coef_out[k]};
    const FacesIn F{n_int, n_bnd, face_int[0], face_int[1], face_int[2], face_int[3], face_int[4], face_int[5], face_int[6], face_int[7], face_int[8], xf, nrm, fjw};
    using enum DimensionEnum; using enum PolynomialOrderEnum; using enum MeshModelEnum; using enum TimeIntegrationEnum;
    switch (case_id) {
      case 0: return runCase<Control<D2, P2, Quadrangle, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out);
      case 1: return runCase<Control<D2, P3, Quadrangle, SSPRK3, NS<TransportModelEnum::Sutherland, ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out);
      case 2: return runCase<Control<D2, P2, Quadrangle, HeunRK2, NS<TransportModelEnum::Constant, ConvectiveFluxEnum::Roe, ViscousFluxEnum::BR1>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out);
      case 3: return runCase<Control<D1, P3, Line, ForwardEuler, Euler<ConvectiveFluxEnum::LaxFriedrichs>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out);
      case 4: return runCase<Control<D3, P2, Hexahedron, SSPRK3, NS<TransportModelEnum::Constant, ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out);
      case 5: return runCase<Control<D2, P2, Triangle, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out);
      case 6: return runCase<Control<D2, P3, TriangleQuadrangle, SSPRK3, NS<TransportModelEnum::Sutherland, ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out);
      case 7: return runCase<Control<D3, P3, Hexahedron, SSPRK3, Euler<ConvectiveFluxEnum::HLLC>>>(blocks, n_blocks, F, nsteps, cfl, dt_in, relerr_out, dt_out);
      default: throw std::runtime_error("ref_sweeps: unknown case");
    }
  } catch (const std::exception& ex) {
    g_error = ex.what();
    return 1;
  }
}

}  // extern "C"
