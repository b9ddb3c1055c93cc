// SubrosaDG.hpp — header-only C++ host side of the B200 path: SubrosaDG's template configuration surface over the C ABI.
//
// The reference is a header-only C++23 library whose user programs (examples/*.cpp) declare
//     using SimulationControl = SubrosaDG::SimulationControl<SolveControl<…>, NumericalControl<…>, …Variable<…>>;
// specialise InitialCondition<SC>::calculatePrimitiveFromCoordinate / BoundaryCondition<SC>::calculatePrimitiveFromCoordinate
// and drive a System<SC> (src/Utils/SystemControl.cpp:55-231).  This header reproduces that surface — same namespace, enum
// names and VALUES (src/Utils/Enum.cpp:22-233), same control templates and constexpr member names
// (src/Solver/SimulationControl.cpp:1197-1279), same System setters, same Solver<SC> members
// (src/Solver/SolveControl.cpp:327-436) — for builds without icpx / Eigen / Gmsh, and forwards the hot path to
// libsubrosadg_b200.so (include/subrosadg_b200.h).  What is NOT here: Gmsh meshing (generateMesh) — meshes come from the
// in-code producers below or from a flat file written by subrosadg_b200.mesh.write_flat — and the View/VTU output.
//
// If Eigen is available, define SUBROSA_DG_B200_USE_EIGEN before including this header; otherwise a minimal
// Eigen::Vector<T,N> stand-in with the members the example callbacks use (brace init, x() y() z(), operator[], Zero(),
// norm()) is provided.
#ifndef SUBROSA_DG_B200_HPP_
#define SUBROSA_DG_B200_HPP_

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <filesystem>
#include <fstream>
#include <initializer_list>
#include <map>
#include <numbers>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

extern "C" {
#include "../subrosadg_b200.h"
}

#ifdef SUBROSA_DG_B200_USE_EIGEN
#include <Eigen/Core>
#else
namespace Eigen {
template <typename T, int N>
struct Vector {
  std::array<T, N> v{};
  Vector() = default;
  Vector(std::initializer_list<T> init) { int i = 0; for (const T& x : init) { if (i < N) v[static_cast<std::size_t>(i++)] = x; } }
  static Vector Zero() { return Vector{}; }
  T& operator[](int i) { return v[static_cast<std::size_t>(i)]; }
  const T& operator[](int i) const { return v[static_cast<std::size_t>(i)]; }
  T& operator()(int i) { return v[static_cast<std::size_t>(i)]; }
  const T& operator()(int i) const { return v[static_cast<std::size_t>(i)]; }
  const T& x() const { return v[0]; }
  const T& y() const { static_assert(N >= 2); return v[1]; }
  const T& z() const { static_assert(N >= 3); return v[2]; }
  T squaredNorm() const { T s{}; for (const T& a : v) s += a * a; return s; }
  T norm() const { return std::sqrt(squaredNorm()); }
  const T* data() const { return v.data(); }
  T* data() { return v.data(); }
};
}  // namespace Eigen
#endif

namespace SubrosaDG {

// src/Utils/BasicDataType.cpp:32-39, src/Utils/Constant.cpp
using Real = double;
using Isize = int;
using Usize = unsigned;
inline constexpr Real kPi{std::numbers::pi_v<Real>};
inline constexpr Real operator""_r(long double x) { return static_cast<Real>(x); }
inline constexpr Real operator""_deg(long double x) { return static_cast<Real>(x) * kPi / 180.0; }

// src/Utils/Enum.cpp:22-233 (names and values)
enum class DimensionEnum { D1 = 1, D2, D3 };
enum class ElementEnum { Point, Line, Triangle, Quadrangle, Tetrahedron, Pyramid, Hexahedron };
enum class MeshModelEnum { Line, Triangle, Quadrangle, TriangleQuadrangle, Tetrahedron, Hexahedron, TetrahedronPyramidHexahedron };
enum class PolynomialOrderEnum { P1 = 1, P2, P3, P4, P5 };
enum class EquationModelEnum { CompresibleEuler, CompresibleNS, IncompresibleEuler, IncompresibleNS, CompresibleRANS, IdealMHD, ViscousMHD };
enum class SourceTermEnum { None, Boussinesq };
enum class ShockCapturingEnum { None, ArtificialViscosity };
enum class LimiterEnum { None, PositivityPreserving };
enum class InitialConditionEnum { Function, SpecificFile, LastStep };
enum class BoundaryConditionEnum { RiemannFarfield, VelocityInflow, PressureOutflow, IsoThermalNonSlipWall, AdiabaticSlipWall, AdiabaticNonSlipWall, Periodic };
enum class BoundaryTimeEnum { Steady, TimeVarying };
enum class ConvectiveFluxEnum { Central, LaxFriedrichs, HLLC, Roe, Exact };
enum class ViscousFluxEnum { None, BR1, BR2 };
enum class ThermodynamicModelEnum { Constant };
enum class EquationOfStateEnum { IdealGas, WeakCompressibleFluid };
enum class TransportModelEnum { None, Constant, Sutherland };
enum class TimeIntegrationEnum { ForwardEuler, HeunRK2, SSPRK3 };
enum class ViewVariableEnum { Density, Velocity, Temperature, Pressure, SoundSpeed, MachNumber, Entropy, Vorticity, HeatFlux, ArtificialViscosity,
                              VelocityX, VelocityY, VelocityZ, MachNumberX, MachNumberY, MachNumberZ, VorticityX, VorticityY, VorticityZ,
                              HeatFluxX, HeatFluxY, HeatFluxZ };

// src/Solver/SimulationControl.cpp:1197-1279
template <DimensionEnum Dimension, PolynomialOrderEnum PolynomialOrder, BoundaryTimeEnum BoundaryTimeType, SourceTermEnum SourceTermType>
struct SolveControl {
  inline static constexpr int kDimension{static_cast<int>(Dimension)};
  inline static constexpr int kPolynomialOrder{static_cast<int>(PolynomialOrder)};
  inline static constexpr BoundaryTimeEnum kBoundaryTime{BoundaryTimeType};
  inline static constexpr SourceTermEnum kSourceTerm{SourceTermType};
};
template <MeshModelEnum MeshModelType, ShockCapturingEnum ShockCapturingType, LimiterEnum LimiterType, InitialConditionEnum InitialConditionType,
          TimeIntegrationEnum TimeIntegrationType>
struct NumericalControl {
  inline static constexpr MeshModelEnum kMeshModel{MeshModelType};
  inline static constexpr InitialConditionEnum kInitialCondition{InitialConditionType};
  inline static constexpr ShockCapturingEnum kShockCapturing{ShockCapturingType};
  inline static constexpr LimiterEnum kLimiter{LimiterType};
  inline static constexpr TimeIntegrationEnum kTimeIntegration{TimeIntegrationType};
};
template <ThermodynamicModelEnum ThermodynamicModelType, EquationOfStateEnum EquationOfStateType, ConvectiveFluxEnum ConvectiveFluxType>
struct CompresibleEulerVariable {
  inline static constexpr EquationModelEnum kEquationModel{EquationModelEnum::CompresibleEuler};
  inline static constexpr ThermodynamicModelEnum kThermodynamicModel{ThermodynamicModelType};
  inline static constexpr EquationOfStateEnum kEquationOfState{EquationOfStateType};
  inline static constexpr TransportModelEnum kTransportModel{TransportModelEnum::None};
  inline static constexpr ConvectiveFluxEnum kConvectiveFlux{ConvectiveFluxType};
  inline static constexpr ViscousFluxEnum kViscousFlux{ViscousFluxEnum::None};
};
template <ThermodynamicModelEnum ThermodynamicModelType, EquationOfStateEnum EquationOfStateType, TransportModelEnum TransportModelType,
          ConvectiveFluxEnum ConvectiveFluxType, ViscousFluxEnum ViscousFluxType>
struct CompresibleNSVariable {
  inline static constexpr EquationModelEnum kEquationModel{EquationModelEnum::CompresibleNS};
  inline static constexpr ThermodynamicModelEnum kThermodynamicModel{ThermodynamicModelType};
  inline static constexpr EquationOfStateEnum kEquationOfState{EquationOfStateType};
  inline static constexpr TransportModelEnum kTransportModel{TransportModelType};
  inline static constexpr ConvectiveFluxEnum kConvectiveFlux{ConvectiveFluxType};
  inline static constexpr ViscousFluxEnum kViscousFlux{ViscousFluxType};
};
template <ThermodynamicModelEnum ThermodynamicModelType, EquationOfStateEnum EquationOfStateType, ConvectiveFluxEnum ConvectiveFluxType>
struct IncompresibleEulerVariable {
  inline static constexpr EquationModelEnum kEquationModel{EquationModelEnum::IncompresibleEuler};
  inline static constexpr ThermodynamicModelEnum kThermodynamicModel{ThermodynamicModelType};
  inline static constexpr EquationOfStateEnum kEquationOfState{EquationOfStateType};
  inline static constexpr TransportModelEnum kTransportModel{TransportModelEnum::None};
  inline static constexpr ConvectiveFluxEnum kConvectiveFlux{ConvectiveFluxType};
  inline static constexpr ViscousFluxEnum kViscousFlux{ViscousFluxEnum::None};
};
template <ThermodynamicModelEnum ThermodynamicModelType, EquationOfStateEnum EquationOfStateType, TransportModelEnum TransportModelType,
          ConvectiveFluxEnum ConvectiveFluxType, ViscousFluxEnum ViscousFluxType>
struct IncompresibleNSVariable {
  inline static constexpr EquationModelEnum kEquationModel{EquationModelEnum::IncompresibleNS};
  inline static constexpr ThermodynamicModelEnum kThermodynamicModel{ThermodynamicModelType};
  inline static constexpr EquationOfStateEnum kEquationOfState{EquationOfStateType};
  inline static constexpr TransportModelEnum kTransportModel{TransportModelType};
  inline static constexpr ConvectiveFluxEnum kConvectiveFlux{ConvectiveFluxType};
  inline static constexpr ViscousFluxEnum kViscousFlux{ViscousFluxType};
};
template <typename SolveControlT, typename NumericalControlT, typename EquationVariable>
struct SimulationControl : SolveControlT, NumericalControlT, EquationVariable {
  // getConservedVariableNumber / getComputationalVariableNumber / getPrimitiveVariableNumber, SimulationControl.cpp:1143-1195
  inline static constexpr int kConservedVariableNumber{SolveControlT::kDimension + 2};
  inline static constexpr int kComputationalVariableNumber{SolveControlT::kDimension + 3};
  inline static constexpr int kPrimitiveVariableNumber{SolveControlT::kDimension + 2};
};

// ---- host mesh: what Mesh<SC>::readMeshElement leaves behind (src/Mesh/ReadControl.cpp:60-155), flattened -------------------
struct MeshData {
  int dim{0};
  struct Block { int type{0}; int geom_order{1}; int n{0}; int nn{0}; std::vector<double> coords; };
  std::vector<Block> blocks;
  int n_int{0}, n_bnd{0};
  std::vector<int32_t> le, lt, lf, re, rt, rf, rot, bc, phys;   // AdjacencyElementMesh records, interior faces first

  // flat file written by subrosadg_b200.mesh.write_flat (little endian): magic "SDGM", dim, nblocks, {type, g, n, nn, coords}, n_int, n_bnd, 9 arrays
  static MeshData readFlat(const std::filesystem::path& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("cannot open mesh file " + path.string());
    auto rd = [&](void* p, std::size_t n) { f.read(static_cast<char*>(p), static_cast<std::streamsize>(n)); if (!f) throw std::runtime_error("truncated mesh file"); };
    char magic[4]; rd(magic, 4);
    if (std::string_view(magic, 4) != "SDGM") throw std::runtime_error("not a flat SubrosaDG-b200 mesh file");
    MeshData m; int32_t nb = 0, d = 0; rd(&d, 4); rd(&nb, 4); m.dim = d;
    for (int b = 0; b < nb; b++) {
      int32_t h[4]; rd(h, 16);
      Block blk; blk.type = h[0]; blk.geom_order = h[1]; blk.n = h[2]; blk.nn = h[3];
      blk.coords.resize(static_cast<std::size_t>(blk.n) * static_cast<std::size_t>(blk.nn) * static_cast<std::size_t>(m.dim));
      rd(blk.coords.data(), blk.coords.size() * sizeof(double));
      m.blocks.push_back(std::move(blk));
    }
    int32_t nf[2]; rd(nf, 8); m.n_int = nf[0]; m.n_bnd = nf[1];
    const std::size_t n = static_cast<std::size_t>(m.n_int + m.n_bnd);
    for (auto* a : {&m.le, &m.lt, &m.lf, &m.re, &m.rt, &m.rf, &m.rot, &m.bc, &m.phys}) { a->resize(n); rd(a->data(), n * 4); }
    return m;
  }
};

// In-code stand-in for the generateMesh() of examples/periodic_{2,3}d_ceuler.cpp: fully periodic [lo,hi]^dim box of n^dim
// quadrangles / hexahedra, lexicographic numbering (x fastest), gmsh corner order, first-encounter face order with the low
// face as the master ("left") side of each periodic pair — identical arrays to subrosadg_b200.mesh.periodic_box_fast.
inline MeshData makePeriodicBox(int dim, int n, double lo = 0.0, double hi = 2.0) {
  if (dim != 2 && dim != 3) throw std::runtime_error("makePeriodicBox: dim must be 2 or 3");
  if (n < 3) throw std::runtime_error("makePeriodicBox: need at least 3 cells per periodic direction");
  MeshData m; m.dim = dim;
  MeshData::Block b; b.type = dim == 2 ? 3 : 6; b.geom_order = 1; b.nn = dim == 2 ? 4 : 8;
  int ne = 1; for (int a = 0; a < dim; a++) ne *= n;
  b.n = ne;
  static const int cq[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
  static const int ch[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
  const double h = (hi - lo) / n;
  b.coords.resize(static_cast<std::size_t>(ne) * static_cast<std::size_t>(b.nn) * static_cast<std::size_t>(dim));
  auto idx = [&](int e, int a) { int s = 1; for (int k = 0; k < a; k++) s *= n; return (e / s) % n; };
  for (int e = 0; e < ne; e++)
    for (int c = 0; c < b.nn; c++)
      for (int a = 0; a < dim; a++) {
        const int off = dim == 2 ? cq[c][a] : ch[c][a];
        b.coords[(static_cast<std::size_t>(e) * static_cast<std::size_t>(b.nn) + static_cast<std::size_t>(c)) * static_cast<std::size_t>(dim) + static_cast<std::size_t>(a)] =
            lo + (static_cast<double>(idx(e, a) + off) / n) * (hi - lo) + 0.0 * h;
      }
  m.blocks.push_back(std::move(b));
  // local faces on the low / high side of each axis (getElementPerAdjacencyNodeIndex, SimulationControl.cpp:177-216)
  const int low2[2] = {3, 0}, high2[2] = {1, 2}, low3[3] = {2, 1, 0}, high3[3] = {3, 4, 5}, rot3[3] = {0, 1, 0};
  struct Rec { long key; int le, lf, re, rf, rot; };
  std::vector<Rec> recs;
  for (int a = 0; a < dim; a++) {
    int s = 1; for (int k = 0; k < a; k++) s *= n;
    const int lowf = dim == 2 ? low2[a] : low3[a], highf = dim == 2 ? high2[a] : high3[a], r = dim == 2 ? 0 : rot3[a];
    for (int e = 0; e < ne; e++) {
      const int i = idx(e, a);
      if (i < n - 1) recs.push_back({static_cast<long>(e) * 8 + highf, e, highf, e + s, lowf, r});
      if (i == 0) recs.push_back({static_cast<long>(e) * 8 + lowf, e, lowf, e + (n - 1) * s, highf, r});
    }
  }
  std::stable_sort(recs.begin(), recs.end(), [](const Rec& x, const Rec& y) { return x.key < y.key; });
  m.n_int = static_cast<int>(recs.size()); m.n_bnd = 0;
  const int t = dim == 2 ? 3 : 6;
  for (const Rec& r : recs) {
    m.le.push_back(r.le); m.lt.push_back(t); m.lf.push_back(r.lf); m.re.push_back(r.re); m.rt.push_back(t); m.rf.push_back(r.rf);
    m.rot.push_back(r.rot); m.bc.push_back(static_cast<int>(BoundaryConditionEnum::Periodic)); m.phys.push_back(0);
  }
  return m;
}

// ---- user callbacks (specialised in every example exactly like in the reference) -------------------------------------------
template <typename SimulationControl>
struct InitialCondition {   // src/Solver/InitialCondition.cpp:36-40
  [[nodiscard]] inline Eigen::Vector<Real, SimulationControl::kPrimitiveVariableNumber> calculatePrimitiveFromCoordinate(
      const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate) const;
};
template <typename SimulationControl>
struct BoundaryCondition {   // src/Solver/BoundaryCondition.cpp:573-579
  [[nodiscard]] inline Eigen::Vector<Real, SimulationControl::kPrimitiveVariableNumber> calculatePrimitiveFromCoordinate(
      const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate, Isize gmsh_physical_index) const;
  [[nodiscard]] inline Eigen::Vector<Real, SimulationControl::kPrimitiveVariableNumber> calculatePrimitiveFromCoordinate(
      const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate, Real time, Isize gmsh_physical_index) const;
};

struct PhysicalModelData {   // src/Solver/PhysicalModel.cpp:26-123 (values set through the System setters)
  Real specific_heat_constant_pressure{2.5}, specific_heat_constant_volume{25.0 / 14.0};
  Real dynamic_viscosity{0.0};
  Real reference_sound_speed{1.0}, reference_density{1.0};
};
struct SourceTermData { Real thermal_expansion_coefficient{0.0}, reference_temperature{0.0}; };   // SourceTerm.cpp:29-58
struct TimeIntegrationData {   // src/Solver/TimeIntegration.cpp:45-65 + SolveControl
  int iteration_start_{0}, iteration_end_{0}, iteration_{0};
  Real courant_friedrichs_lewy_number_{0.0}, delta_time_{0.0};
};

// ---- Solver<SC>: the drop-in seam (src/Solver/SolveControl.cpp:327-436) over the C ABI ---------------------------------------------
template <typename SimulationControl>
struct Solver {
  inline static constexpr int kNv{SimulationControl::kConservedVariableNumber};
  std::array<Real, static_cast<std::size_t>(kNv)> relative_error_{};   // SolveControl.cpp:300
  sdg_ctx* ctx_{nullptr};
  std::vector<int> types_;
  std::vector<double> boundary_coordinate_;
  std::vector<int32_t> boundary_physical_;
  int n_bnd_{0}, nqf_{0};

  Solver() = default;
  Solver(const Solver&) = delete;
  Solver& operator=(const Solver&) = delete;
  ~Solver() { if (ctx_ != nullptr) sdg_destroy(ctx_); }

  static void check(int rc) { if (rc != 0) throw std::runtime_error(std::string("subrosadg_b200: ") + sdg_last_error()); }

  inline void createContext(const MeshData& mesh, const PhysicalModelData& pm, const SourceTermData& st, int device) {
    sdg_config cfg{};
    cfg.dim = SimulationControl::kDimension; cfg.p = SimulationControl::kPolynomialOrder;
    cfg.model = static_cast<int>(SimulationControl::kEquationModel);
    cfg.eos = static_cast<int>(SimulationControl::kEquationOfState);
    cfg.transport = static_cast<int>(SimulationControl::kTransportModel);
    cfg.conv_flux = static_cast<int>(SimulationControl::kConvectiveFlux);
    cfg.visc_flux = static_cast<int>(SimulationControl::kViscousFlux);
    cfg.source = static_cast<int>(SimulationControl::kSourceTerm);
    cfg.rk = static_cast<int>(SimulationControl::kTimeIntegration);
    cfg.device = device; cfg.chunk = 0; cfg.reorder = 1;
    cfg.cp = pm.specific_heat_constant_pressure; cfg.cv = pm.specific_heat_constant_volume; cfg.mu = pm.dynamic_viscosity;
    cfg.c0 = pm.reference_sound_speed; cfg.rho0 = pm.reference_density;
    cfg.beta = st.thermal_expansion_coefficient; cfg.t_ref = st.reference_temperature;
    check(sdg_create(&cfg, &ctx_));
    for (const auto& b : mesh.blocks) { check(sdg_add_elements(ctx_, b.type, b.n, 0, b.geom_order, b.coords.data())); types_.push_back(b.type); }
    check(sdg_set_faces(ctx_, mesh.n_int, mesh.n_bnd, mesh.le.data(), mesh.lt.data(), mesh.lf.data(), mesh.re.data(), mesh.rt.data(), mesh.rf.data(),
                        mesh.rot.data(), mesh.bc.data(), mesh.phys.data()));
    check(sdg_finalize(ctx_));
    n_bnd_ = mesh.n_bnd;
    boundary_physical_.assign(mesh.phys.begin() + mesh.n_int, mesh.phys.end());
  }

  // Solver::initializeSolver, SolveControl.cpp:377-380 / InitialCondition.cpp:151-186
  inline void initializeSolver(const MeshData& mesh, const PhysicalModelData& physical_model, const SourceTermData& source_term,
                               const BoundaryCondition<SimulationControl>& boundary_condition,
                               const InitialCondition<SimulationControl>& initial_condition, int device = 0) {
    constexpr int D = SimulationControl::kDimension, NP = SimulationControl::kPrimitiveVariableNumber;
    createContext(mesh, physical_model, source_term, device);
    for (int t : types_) {
      int32_t sz[8]; check(sdg_sizes(ctx_, t, sz));
      const std::size_t npt = static_cast<std::size_t>(sz[0]) * static_cast<std::size_t>(sz[2]);
      std::vector<double> xq(npt * D), prim(npt * NP);
      check(sdg_get_quadrature_coordinates(ctx_, t, xq.data()));
      for (std::size_t i = 0; i < npt; i++) {
        Eigen::Vector<Real, D> x; for (int d = 0; d < D; d++) x[d] = xq[i * D + static_cast<std::size_t>(d)];
        const auto p = initial_condition.calculatePrimitiveFromCoordinate(x);
        for (int k = 0; k < NP; k++) prim[i * NP + static_cast<std::size_t>(k)] = p[k];
      }
      check(sdg_set_state_from_primitive(ctx_, t, prim.data()));
      nqf_ = sz[6];
    }
    if (n_bnd_ > 0) {
      boundary_coordinate_.resize(static_cast<std::size_t>(n_bnd_) * static_cast<std::size_t>(nqf_) * D);
      check(sdg_get_boundary_quadrature_coordinates(ctx_, boundary_coordinate_.data()));
      updateBoundaryVariable(boundary_condition, 0.0, false);
    }
  }

  // Solver::updateBoundaryVariable, BoundaryCondition.cpp:29-74
  inline void updateBoundaryVariable(const BoundaryCondition<SimulationControl>& boundary_condition, Real time, bool time_varying) {
    constexpr int D = SimulationControl::kDimension, NP = SimulationControl::kPrimitiveVariableNumber;
    const std::size_t npt = static_cast<std::size_t>(n_bnd_) * static_cast<std::size_t>(nqf_);
    std::vector<double> prim(npt * NP);
    for (std::size_t i = 0; i < npt; i++) {
      Eigen::Vector<Real, D> x; for (int d = 0; d < D; d++) x[d] = boundary_coordinate_[i * D + static_cast<std::size_t>(d)];
      const Isize phys = boundary_physical_[i / static_cast<std::size_t>(nqf_)];
      Eigen::Vector<Real, NP> p;
      if constexpr (SimulationControl::kBoundaryTime == BoundaryTimeEnum::TimeVarying) p = boundary_condition.calculatePrimitiveFromCoordinate(x, time, phys);
      else p = boundary_condition.calculatePrimitiveFromCoordinate(x, phys);
      for (int k = 0; k < NP; k++) prim[i * NP + static_cast<std::size_t>(k)] = p[k];
    }
    static_cast<void>(time_varying);
    check(sdg_set_boundary_primitive(ctx_, prim.data()));
  }

  // Solver::calculateDeltaTime, SolveControl.cpp:389-391
  inline void calculateDeltaTime(TimeIntegrationData& time_integration) {
    check(sdg_compute_dt(ctx_, time_integration.courant_friedrichs_lewy_number_, &time_integration.delta_time_));
  }

  // Solver::stepSolver, SolveControl.cpp:427-431 / TimeIntegration.cpp:326-350
  inline void stepSolver(const BoundaryCondition<SimulationControl>& boundary_condition, const TimeIntegrationData& time_integration) {
    if constexpr (SimulationControl::kBoundaryTime == BoundaryTimeEnum::TimeVarying) {
      if (n_bnd_ > 0) updateBoundaryVariable(boundary_condition, time_integration.iteration_ * time_integration.delta_time_, true);
    }
    check(sdg_step(ctx_, time_integration.delta_time_, 1, relative_error_.data()));
  }

  // payload of Solver::writeRawBinary (RawBinary.cpp:75-88): modal coefficients [n][Nb][Nv] of one element type
  inline std::vector<double> getCoefficient(int type) const {
    int32_t sz[8]; check(sdg_sizes(ctx_, type, sz));
    std::vector<double> u(static_cast<std::size_t>(sz[0]) * static_cast<std::size_t>(sz[1]) * static_cast<std::size_t>(sz[7]));
    check(sdg_get_state(ctx_, type, u.data()));
    return u;
  }
  inline std::vector<double> getStateAtQuadrature(int type) const {
    int32_t sz[8]; check(sdg_sizes(ctx_, type, sz));
    std::vector<double> u(static_cast<std::size_t>(sz[0]) * static_cast<std::size_t>(sz[2]) * static_cast<std::size_t>(sz[7]));
    check(sdg_get_state_at_quadrature(ctx_, type, u.data()));
    return u;
  }
  inline std::vector<double> getQuadratureCoordinate(int type) const {
    int32_t sz[8]; check(sdg_sizes(ctx_, type, sz));
    std::vector<double> x(static_cast<std::size_t>(sz[0]) * static_cast<std::size_t>(sz[2]) * static_cast<std::size_t>(SimulationControl::kDimension));
    check(sdg_get_quadrature_coordinates(ctx_, type, x.data()));
    return x;
  }
};

// ---- System<SC>: src/Utils/SystemControl.cpp:55-231 -------------------------------------------------------------------------------------
template <typename SimulationControl>
struct System {
  MeshData mesh_;
  PhysicalModelData physical_model_;
  SourceTermData source_term_;
  BoundaryCondition<SimulationControl> boundary_condition_;
  InitialCondition<SimulationControl> initial_condition_;
  TimeIntegrationData time_integration_;
  Solver<SimulationControl> solver_;
  std::map<Isize, BoundaryConditionEnum> physical_boundary_;
  std::filesystem::path output_directory_;
  std::string output_file_name_prefix_;
  int io_interval_{0};
  int device_{0};
  bool print_{true};

  // the reference takes (mesh_file_path, generateMesh); Gmsh is not available, so the producer returns the flat mesh directly
  inline void setMesh(MeshData mesh) { mesh_ = std::move(mesh); }
  inline void setMesh(const std::filesystem::path& flat_mesh_file) { mesh_ = MeshData::readFlat(flat_mesh_file); }

  template <BoundaryConditionEnum BoundaryConditionType>
  inline void addBoundaryCondition(const Isize physical_index) { physical_boundary_[physical_index] = BoundaryConditionType; }

  template <ThermodynamicModelEnum ThermodynamicModelType>
    requires(ThermodynamicModelType == ThermodynamicModelEnum::Constant)
  inline void setThermodynamicModel(const Real specific_heat_constant_pressure, const Real specific_heat_constant_volume) {
    physical_model_.specific_heat_constant_pressure = specific_heat_constant_pressure;
    physical_model_.specific_heat_constant_volume = specific_heat_constant_volume;
  }
  template <EquationOfStateEnum EquationOfStateType>
    requires(EquationOfStateType == EquationOfStateEnum::WeakCompressibleFluid)
  inline void setEquationOfState(const Real reference_sound_speed, const Real reference_density) {
    physical_model_.reference_sound_speed = reference_sound_speed; physical_model_.reference_density = reference_density;
  }
  template <TransportModelEnum TransportModelType>
    requires(TransportModelType == TransportModelEnum::Constant || TransportModelType == TransportModelEnum::Sutherland)
  inline void setTransportModel(const Real dynamic_viscosity) { physical_model_.dynamic_viscosity = dynamic_viscosity; }
  template <SourceTermEnum SourceTermType>
    requires(SourceTermType == SourceTermEnum::Boussinesq)
  inline void setSourceTerm(const Real thermal_expansion_coefficient, const Real reference_temperature) {
    source_term_.thermal_expansion_coefficient = thermal_expansion_coefficient; source_term_.reference_temperature = reference_temperature;
  }
  inline void setTimeIntegration(const Real courant_friedrichs_lewy_number, const std::pair<int, int> iteration_range) {
    time_integration_.iteration_start_ = iteration_range.first; time_integration_.iteration_end_ = iteration_range.second;
    time_integration_.courant_friedrichs_lewy_number_ = courant_friedrichs_lewy_number;
  }
  inline void setDeltaTime(const Real delta_time) { time_integration_.delta_time_ = delta_time; }
  inline void setViewConfig(const std::filesystem::path& output_directory, const std::string_view output_file_name_prefix, const int io_interval = -1) {
    output_directory_ = output_directory; output_file_name_prefix_ = std::string(output_file_name_prefix); io_interval_ = io_interval;
  }
  inline void addViewVariable(const std::vector<ViewVariableEnum>&) {}   // View/VTU output is out of scope (host post-processing)
  inline void setDevice(int device) { device_ = device; }

  // System::synchronize, SystemControl.cpp:142-157: boundary types onto the face records
  inline void synchronize() {
    for (int i = mesh_.n_int; i < mesh_.n_int + mesh_.n_bnd; i++) {
      const auto it = physical_boundary_.find(mesh_.phys[static_cast<std::size_t>(i)]);
      if (it == physical_boundary_.end()) throw std::runtime_error("boundary face without addBoundaryCondition for its physical index");
      mesh_.bc[static_cast<std::size_t>(i)] = static_cast<int>(it->second);
    }
  }

  // System::solve, SystemControl.cpp:159-195
  inline void solve() {
    solver_.initializeSolver(mesh_, physical_model_, source_term_, boundary_condition_, initial_condition_, device_);
    if (time_integration_.delta_time_ == 0.0) solver_.calculateDeltaTime(time_integration_);
    for (int i = time_integration_.iteration_start_ + 1; i <= time_integration_.iteration_end_; i++) {
      solver_.stepSolver(boundary_condition_, time_integration_);
      time_integration_.iteration_ = i;   // after the step, as SystemControl.cpp:175-177: step i sees t = (i - 1) dt
      bool all_nan = true;
      for (Real e : solver_.relative_error_) all_nan = all_nan && std::isnan(e);
      if (print_ && (i == time_integration_.iteration_end_ || i % std::max(1, io_interval_ > 0 ? io_interval_ : time_integration_.iteration_end_) == 0)) {
        std::printf("%13.5e", time_integration_.delta_time_ * i);   // error.txt line, CommandLine.cpp:129-133
        for (Real e : solver_.relative_error_) std::printf(" |%13.5e", e);
        std::printf("\n");
      }
      if (all_nan) { time_integration_.iteration_end_ = i; break; }   // SystemControl.cpp:185-191
    }
  }
  inline void view() {}   // out of scope
};

}  // namespace SubrosaDG

using SubrosaDG::operator""_r;
using SubrosaDG::operator""_deg;

#endif  // SUBROSA_DG_B200_HPP_
