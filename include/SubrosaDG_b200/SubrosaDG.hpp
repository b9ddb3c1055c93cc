// SubrosaDG.hpp — header-only C++ host side of the B200 path: SubrosaDG's template configuration surface over the C ABI.
//
// The reference is a header-only C++23 library whose user programs (examples/*.cpp) declare
//     using SimulationControl = SubrosaDG::SimulationControl<SolveControl<…>, NumericalControl<…>, …Variable<…>>;
// specialise InitialCondition<SC>::calculatePrimitiveFromCoordinate / BoundaryCondition<SC>::calculatePrimitiveFromCoordinate
// and drive a System<SC> (src/Utils/SystemControl.cpp:55-231).  This header reproduces that surface — same namespace, enum
// names and VALUES (src/Utils/Enum.cpp:22-233), same control templates and constexpr member names
// (src/Solver/SimulationControl.cpp:1197-1279), same System setters, and Solver<SC> with the reference's member SIGNATURES
// (src/Solver/SolveControl.cpp:290-300,377-435: initializeSolver(mesh, physical_model, boundary_condition, initial_condition),
// calculateDeltaTime(mesh, physical_model, time_integration), stepSolver(mesh, source_term, physical_model, boundary_condition,
// time_integration), writeRawBinary(mesh, path)) and fields (relative_error_, error_finout_, raw_binary_ss_,
// write_raw_binary_future_, node_artificial_viscosity_) — for builds without icpx / Eigen / Gmsh — and forwards the hot path to
// libsubrosadg_b200.so (include/subrosadg_b200.h).  The raw/<prefix>_<step>.zst files are the reference's own container
// (src/View/RawBinary.cpp:42-74) and payload (:75-191); InitialConditionEnum::LastStep / SpecificFile read them back
// (src/Solver/InitialCondition.cpp:41-80).  What is NOT here: Gmsh meshing (generateMesh) — meshes come from the in-code producers
// below or from a flat file written by subrosadg_b200.mesh.write_flat — and the View/VTU output.
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
#include <cstring>
#include <filesystem>
#include <fstream>
#include <functional>
#include <future>
#include <initializer_list>
#include <map>
#include <numbers>
#include <sstream>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

#include <dlfcn.h>

extern "C" {
#include "../subrosadg_b200.h"
}

#ifdef SUBROSA_DG_B200_USE_EIGEN
#include <Eigen/Core>
#else
namespace Eigen {
inline constexpr int Dynamic = -1;
template <typename T, int N>
struct Vector;
template <typename T>
struct Vector<T, Dynamic> {   // what Solver::node_artificial_viscosity_ needs
  std::vector<T> v;
  void resize(long n) { v.resize(static_cast<std::size_t>(n)); }
  void setZero() { std::fill(v.begin(), v.end(), T{}); }
  [[nodiscard]] long size() const { return static_cast<long>(v.size()); }
  T& operator()(long i) { return v[static_cast<std::size_t>(i)]; }
  const T& operator()(long i) const { return v[static_cast<std::size_t>(i)]; }
  const T* data() const { return v.data(); }
  T* data() { return v.data(); }
};
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
  Isize node_number_{0};      // Mesh::node_number_ (ReadControl.cpp:248): length of Solver::node_artificial_viscosity_ in the raw files
  Isize element_number_{0};

  // the flat format keeps coordinates per element, not node tags: nodes = distinct coordinate tuples (the copies of a periodic pair
  // stay distinct, as in the Gmsh mesh)
  std::vector<std::array<double, 3>> node_point_;   // the distinct nodes, sorted: tag = position (0-based)
  void countNodes() {
    std::vector<std::array<double, 3>>& pts = node_point_;
    pts.clear();
    element_number_ = 0;
    for (const Block& b : blocks) {
      element_number_ += b.n;
      const std::size_t np = static_cast<std::size_t>(b.n) * static_cast<std::size_t>(b.nn);
      for (std::size_t i = 0; i < np; i++) pts.push_back(point(b, i));
    }
    std::sort(pts.begin(), pts.end());
    pts.erase(std::unique(pts.begin(), pts.end()), pts.end());
    node_number_ = static_cast<Isize>(pts.size());
  }
  [[nodiscard]] std::array<double, 3> point(const Block& b, std::size_t node) const {
    std::array<double, 3> x{0.0, 0.0, 0.0};
    for (int d = 0; d < dim; d++) x[static_cast<std::size_t>(d)] = b.coords[node * static_cast<std::size_t>(dim) + static_cast<std::size_t>(d)];
    return x;
  }
  [[nodiscard]] static int basicNodeNumber(int type) { return type == 1 ? 2 : type == 2 ? 3 : type == 3 ? 4 : 8; }   // kBasicNodeNumber: the corners lead the gmsh node order
  // PerElementMesh::node_tag_ of the corner nodes (ReadControl.cpp:64), 0-based, [n][kBasicNodeNumber]
  [[nodiscard]] std::vector<int32_t> nodeTags(const Block& b) const {
    const int nb = basicNodeNumber(b.type);
    std::vector<int32_t> tags(static_cast<std::size_t>(b.n) * static_cast<std::size_t>(nb));
    for (int e = 0; e < b.n; e++)
      for (int k = 0; k < nb; k++) {
        const auto x = point(b, static_cast<std::size_t>(e) * static_cast<std::size_t>(b.nn) + static_cast<std::size_t>(k));
        tags[static_cast<std::size_t>(e) * static_cast<std::size_t>(nb) + static_cast<std::size_t>(k)] =
            static_cast<int32_t>(std::lower_bound(node_point_.begin(), node_point_.end(), x) - node_point_.begin());
      }
    return tags;
  }
  // smallest circle tangent to three consecutive edge lines of a planar quadrangle (Gmsh MQuadrangle::getInnerRadius, restated)
  [[nodiscard]] static double quadInnerRadius(const std::array<std::array<double, 2>, 4>& P) {
    auto unit = [](double x, double y) { const double l = std::sqrt(x * x + y * y); return std::array<double, 2>{x / l, y / l}; };
    double best = 1.0e300;
    for (std::size_t i = 0; i < 4; i++) {
      const auto &A = P[i], &B = P[(i + 1) % 4], &prev = P[(i + 3) % 4], &nxt = P[(i + 2) % 4];
      const auto a1 = unit(prev[0] - A[0], prev[1] - A[1]), a2 = unit(B[0] - A[0], B[1] - A[1]);
      const auto b1 = unit(A[0] - B[0], A[1] - B[1]), b2 = unit(nxt[0] - B[0], nxt[1] - B[1]);
      const double dA[2] = {a1[0] + a2[0], a1[1] + a2[1]}, dB[2] = {b1[0] + b2[0], b1[1] + b2[1]};
      const double den = dA[0] * dB[1] - dA[1] * dB[0];
      if (std::abs(den) < 1.0e-300) continue;
      const double ab[2] = {B[0] - A[0], B[1] - A[1]};
      const double t = (ab[0] * dB[1] - ab[1] * dB[0]) / den;
      const double c[2] = {t * dA[0], t * dA[1]};   // centre - A
      best = std::min(best, std::abs(ab[0] * c[1] - ab[1] * c[0]) / std::sqrt(ab[0] * ab[0] + ab[1] * ab[1]));
    }
    return best;
  }
  // PerElementMesh::inner_radius_ = gmsh "innerRadius" quality (Geometry.cpp:31-41), restated: line = half length, triangle = inscribed
  // circle, quadrangle = quadInnerRadius, hexahedron = minimum over its faces
  [[nodiscard]] std::vector<double> innerRadius(const Block& b) const {
    std::vector<double> r(static_cast<std::size_t>(b.n));
    static const int hexFace[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {0, 4, 7, 3}, {1, 2, 6, 5}, {2, 3, 7, 6}, {4, 5, 6, 7}};
    for (int e = 0; e < b.n; e++) {
      auto X = [&](int k) { return point(b, static_cast<std::size_t>(e) * static_cast<std::size_t>(b.nn) + static_cast<std::size_t>(k)); };
      auto dist = [](const std::array<double, 3>& p, const std::array<double, 3>& q) { return std::sqrt((p[0] - q[0]) * (p[0] - q[0]) + (p[1] - q[1]) * (p[1] - q[1]) + (p[2] - q[2]) * (p[2] - q[2])); };
      double v = 0.0;
      if (b.type == 1) {
        v = 0.5 * dist(X(0), X(1));
      } else if (b.type == 2) {
        const double a = dist(X(0), X(1)), c = dist(X(1), X(2)), d = dist(X(2), X(0)), k = 0.5 * (a + c + d);
        v = std::sqrt(k * (k - a) * (k - c) * (k - d)) / k;
      } else if (b.type == 3) {
        v = quadInnerRadius({{{X(0)[0], X(0)[1]}, {X(1)[0], X(1)[1]}, {X(2)[0], X(2)[1]}, {X(3)[0], X(3)[1]}}});
      } else {
        v = 1.0e300;
        for (const auto& f : hexFace) {
          const auto p0 = X(f[0]), p1 = X(f[1]), p2 = X(f[2]), p3 = X(f[3]);
          const double d1[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]}, d2[3] = {p3[0] - p1[0], p3[1] - p1[1], p3[2] - p1[2]};
          double n[3] = {d1[1] * d2[2] - d1[2] * d2[1], d1[2] * d2[0] - d1[0] * d2[2], d1[0] * d2[1] - d1[1] * d2[0]};
          const double nl = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
          for (double& c : n) c /= nl;
          double e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
          const double el = std::sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
          for (double& c : e1) c /= el;
          const double e2[3] = {n[1] * e1[2] - n[2] * e1[1], n[2] * e1[0] - n[0] * e1[2], n[0] * e1[1] - n[1] * e1[0]};
          std::array<std::array<double, 2>, 4> P{};
          const std::array<double, 3> q[4] = {p0, p1, p2, p3};
          for (std::size_t m = 0; m < 4; m++) {
            const double w[3] = {q[m][0] - p0[0], q[m][1] - p0[1], q[m][2] - p0[2]};
            P[m] = {w[0] * e1[0] + w[1] * e1[1] + w[2] * e1[2], w[0] * e2[0] + w[1] * e2[1] + w[2] * e2[2]};
          }
          v = std::min(v, quadInnerRadius(P));
        }
      }
      r[static_cast<std::size_t>(e)] = v;
    }
    return r;
  }
  void writeFlat(const std::filesystem::path& path) const {
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    if (!f) throw std::runtime_error("cannot open mesh file " + path.string());
    auto wr = [&](const void* p, std::size_t n) { f.write(static_cast<const char*>(p), static_cast<std::streamsize>(n)); };
    const int32_t d = dim, nb = static_cast<int32_t>(blocks.size());
    wr("SDGM", 4); wr(&d, 4); wr(&nb, 4);
    for (const Block& b : blocks) { const int32_t h[4] = {b.type, b.geom_order, b.n, b.nn}; wr(h, 16); wr(b.coords.data(), b.coords.size() * sizeof(double)); }
    const int32_t nf[2] = {n_int, n_bnd}; wr(nf, 8);
    for (const auto* a : {&le, &lt, &lf, &re, &rt, &rf, &rot, &bc, &phys}) wr(a->data(), a->size() * 4);
  }

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
    m.countNodes();
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
  m.countNodes();
  return m;
}

// Mesh<SC> (src/Mesh/ReadControl.cpp:158-300): the flattened records above behind the reference's class name
template <typename SimulationControl>
struct Mesh : MeshData {
  inline void initializeMesh(const std::filesystem::path& mesh_file_path) { static_cast<MeshData&>(*this) = MeshData::readFlat(mesh_file_path); }
};

// ---- physical model / source term / time integration: the reference's parameter carriers ------------------------------------------
// src/Solver/PhysicalModel.cpp:22-123 — the parameters are `inline static` members set through the System setters.  The pointwise
// functions that use them (equation of state, Sutherland's law ...) run on the device (subrosadg_b200/csrc/physics.cuh).
template <ThermodynamicModelEnum ThermodynamicModelType>
struct ThermodynamicModel;
template <>
struct ThermodynamicModel<ThermodynamicModelEnum::Constant> {
  inline static Real specific_heat_constant_pressure{2.5};
  inline static Real specific_heat_constant_volume{25.0 / 14.0};
};
template <EquationOfStateEnum EquationOfStateType>
struct EquationOfState;
template <>
struct EquationOfState<EquationOfStateEnum::IdealGas> {
  inline static constexpr Real kSpecificHeatRatio = 1.4;
};
template <>
struct EquationOfState<EquationOfStateEnum::WeakCompressibleFluid> {
  inline static Real reference_sound_speed{1.0};
  inline static Real reference_density{1.0};
  inline static Real reference_pressure_addition{0.01};
  inline void calculatePressureAdditionFromSoundSpeedDensity() {
    reference_pressure_addition = 0.01 * reference_density * reference_sound_speed * reference_sound_speed;
  }
};
template <TransportModelEnum TransportModelType>
struct TransportModel {   // Constant and Sutherland carry the same two parameters
  inline static Real dynamic_viscosity{0.0};
  inline static Real thermal_conductivity{0.0};
  inline static constexpr Real kPrandtlNumber = 0.71;
};
template <>
struct TransportModel<TransportModelEnum::None> {};
template <typename SimulationControl>
struct PhysicalModel {
  ThermodynamicModel<SimulationControl::kThermodynamicModel> thermodynamic_model_;
  EquationOfState<SimulationControl::kEquationOfState> equation_of_state_;
  TransportModel<SimulationControl::kTransportModel> transport_model_;
  inline void calculateThermalConductivityFromDynamicViscosity() {   // PhysicalModel.cpp (k = cp mu / Pr)
    if constexpr (SimulationControl::kTransportModel != TransportModelEnum::None) {
      transport_model_.thermal_conductivity =
          thermodynamic_model_.specific_heat_constant_pressure * transport_model_.dynamic_viscosity / transport_model_.kPrandtlNumber;
    }
  }
};

// src/Solver/SourceTerm.cpp:25-58
template <typename SimulationControl, SourceTermEnum SourceTermType>
struct SourceTermBase {};
template <typename SimulationControl>
struct SourceTermBase<SimulationControl, SourceTermEnum::Boussinesq> {
  inline static constexpr Real kGravity = 1.0;
  inline static Real thermal_expansion_coefficient{0.0};
  inline static Real reference_temperature{0.0};
};
template <typename SimulationControl>
struct SourceTerm : SourceTermBase<SimulationControl, SimulationControl::kSourceTerm> {};

// src/Solver/TimeIntegration.cpp:31-65
struct TimeIntegrationBase {
  int iteration_start_{0};
  int iteration_end_{0};
  int iteration_{0};
  Real courant_friedrichs_lewy_number_{0.0};
  Real delta_time_{0.0};
};
template <typename SimulationControl>
struct TimeIntegration : TimeIntegrationBase {
  inline static constexpr int kStep{SimulationControl::kTimeIntegration == TimeIntegrationEnum::ForwardEuler ? 1
                                    : SimulationControl::kTimeIntegration == TimeIntegrationEnum::HeunRK2    ? 2
                                                                                                             : 3};
};

// ---- the raw/<prefix>_<step>.zst container, src/View/RawBinary.cpp:42-74 --------------------------------------------------------------
// [ std::size_t ZSTD_compressBound(payload size) ][ one zstd frame of the payload, level 1 ].  The reader takes the 8-byte header as the
// destination capacity of ZSTD_decompress (it is >= the payload size) and the rest of the file as the frame.
// The reference links libzstd; this header binds the same four functions from the libzstd.so.1 of the system at run time (no zstd
// headers are needed to build).  Without the library the writer emits a standard zstd frame of RAW blocks (RFC 8878 3.1.1: magic,
// frame header with the 8-byte content size, 3-byte block headers) that libzstd decodes like any other frame, and the reader decodes
// frames made of raw / RLE blocks (its own files) and refuses compressed blocks.
struct RawBinaryCompress {
  struct Zstd {
    std::size_t (*compressBound)(std::size_t){nullptr};
    std::size_t (*compress)(void*, std::size_t, const void*, std::size_t, int){nullptr};
    std::size_t (*decompress)(void*, std::size_t, const void*, std::size_t){nullptr};
    unsigned (*isError)(std::size_t){nullptr};
    [[nodiscard]] bool ok() const { return compressBound != nullptr && compress != nullptr && decompress != nullptr && isError != nullptr; }
  };
  inline static bool use_system_zstd{true};   // tests switch the library off to cover the self-contained frame writer / reader
  inline static const Zstd& zstd() {
    static const Zstd z = [] {
      Zstd r;
      if (void* h = dlopen("libzstd.so.1", RTLD_NOW | RTLD_LOCAL)) {
        r.compressBound = reinterpret_cast<decltype(r.compressBound)>(dlsym(h, "ZSTD_compressBound"));
        r.compress = reinterpret_cast<decltype(r.compress)>(dlsym(h, "ZSTD_compress"));
        r.decompress = reinterpret_cast<decltype(r.decompress)>(dlsym(h, "ZSTD_decompress"));
        r.isError = reinterpret_cast<decltype(r.isError)>(dlsym(h, "ZSTD_isError"));
      }
      return r;
    }();
    return z;
  }
  // ZSTD_COMPRESSBOUND of zstd.h: srcSize + (srcSize >> 8) + (srcSize < 128 KB ? (128 KB - srcSize) >> 11 : 0)
  [[nodiscard]] inline static std::size_t compressBound(std::size_t n) {
    return n + (n >> 8) + (n < (std::size_t{128} << 10) ? ((std::size_t{128} << 10) - n) >> 11 : 0);
  }
  inline static constexpr std::size_t kBlock{std::size_t{1} << 17};   // Block_Maximum_Size
  [[nodiscard]] inline static std::string rawFrame(const std::string& src) {
    std::string out;
    out.reserve(src.size() + 14 + 3 * (src.size() / kBlock + 1));
    const unsigned char head[6] = {0x28, 0xB5, 0x2F, 0xFD, 0xC0 /* 8-byte Frame_Content_Size, no checksum, no dictionary */, 0x38 /* window 128 KB */};
    out.append(reinterpret_cast<const char*>(head), 6);
    const std::uint64_t n = src.size();
    for (int b = 0; b < 8; b++) out.push_back(static_cast<char>((n >> (8 * b)) & 0xFF));
    std::size_t at = 0;
    do {
      const std::size_t len = std::min(kBlock, src.size() - at);
      const bool last = at + len == src.size();
      const std::uint32_t h = static_cast<std::uint32_t>(len << 3) | (last ? 1U : 0U);   // Block_Type 0 = Raw_Block
      for (int b = 0; b < 3; b++) out.push_back(static_cast<char>((h >> (8 * b)) & 0xFF));
      out.append(src, at, len);
      at += len;
    } while (at < src.size());
    return out;
  }
  [[nodiscard]] inline static std::string decodeRawFrame(const std::string& in) {
    auto u8 = [&](std::size_t i) { if (i >= in.size()) throw std::runtime_error("truncated zstd frame"); return static_cast<unsigned>(static_cast<unsigned char>(in[i])); };
    if (in.size() < 6 || u8(0) != 0x28 || u8(1) != 0xB5 || u8(2) != 0x2F || u8(3) != 0xFD) throw std::runtime_error("not a zstd frame");
    const unsigned fhd = u8(4);
    const bool single = ((fhd >> 5) & 1U) != 0U, checksum = ((fhd >> 2) & 1U) != 0U;
    const unsigned fcs = fhd >> 6, did = fhd & 3U;
    std::size_t at = 5 + (single ? 0 : 1) + (did == 3 ? 4 : did);
    at += fcs == 0 ? (single ? 1 : 0) : (std::size_t{1} << fcs);
    std::string out;
    for (;;) {
      const std::uint32_t h = u8(at) | (u8(at + 1) << 8) | (u8(at + 2) << 16);
      at += 3;
      const std::size_t len = h >> 3;
      const unsigned type = (h >> 1) & 3U;
      if (type == 0) { if (at + len > in.size()) throw std::runtime_error("truncated zstd frame"); out.append(in, at, len); at += len; }
      else if (type == 1) { out.append(len, in[at]); at += 1; }
      else throw std::runtime_error("compressed zstd block: libzstd.so.1 is needed to read this file");
      if ((h & 1U) != 0U) break;
    }
    static_cast<void>(checksum);
    return out;
  }

  inline static void write(const std::filesystem::path& raw_binary_path, std::stringstream& raw_binary_ss) {
    raw_binary_ss.seekg(0, std::ios::beg);
    raw_binary_ss.seekp(0, std::ios::beg);
    const std::string payload = raw_binary_ss.str();
    std::ofstream fout(raw_binary_path, std::ios::binary | std::ios::trunc);
    if (!fout) throw std::runtime_error("cannot open " + raw_binary_path.string());
    const Zstd& z = zstd();
    const bool lib = use_system_zstd && z.ok();
    const std::size_t bound = lib ? z.compressBound(payload.size()) : compressBound(payload.size());
    fout.write(reinterpret_cast<const char*>(&bound), static_cast<std::streamsize>(sizeof(std::size_t)));
    if (lib) {
      std::vector<char> compressed(bound);
      const std::size_t actual = z.compress(compressed.data(), bound, payload.data(), payload.size(), 1);
      if (z.isError(actual) != 0U) throw std::runtime_error("ZSTD_compress failed");
      fout.write(compressed.data(), static_cast<std::streamsize>(actual));
    } else {
      const std::string frame = rawFrame(payload);
      fout.write(frame.data(), static_cast<std::streamsize>(frame.size()));
    }
  }

  inline static void read(const std::filesystem::path& raw_binary_path, std::stringstream& raw_binary_ss) {
    raw_binary_ss.seekg(0, std::ios::beg);
    raw_binary_ss.seekp(0, std::ios::beg);
    std::ifstream fin(raw_binary_path, std::ios::binary);
    if (!fin) throw std::runtime_error("cannot open " + raw_binary_path.string());
    std::size_t capacity = 0;
    fin.read(reinterpret_cast<char*>(&capacity), static_cast<std::streamsize>(sizeof(std::size_t)));
    std::string frame((std::istreambuf_iterator<char>(fin)), std::istreambuf_iterator<char>());
    if (!fin.eof() && !fin) throw std::runtime_error("cannot read " + raw_binary_path.string());
    const Zstd& z = zstd();
    if (use_system_zstd && z.ok()) {
      std::string out(capacity, '\0');
      const std::size_t n = z.decompress(out.data(), capacity, frame.data(), frame.size());
      if (z.isError(n) != 0U) throw std::runtime_error("ZSTD_decompress failed on " + raw_binary_path.string());
      out.resize(n);
      raw_binary_ss << out;
    } else {
      raw_binary_ss << decodeRawFrame(frame);
    }
  }
};

// ---- user callbacks (specialised in every example exactly like in the reference) -------------------------------------------
template <typename SimulationControl>
struct InitialCondition {   // src/Solver/InitialCondition.cpp:32-40
  std::filesystem::path raw_binary_path_;
  std::stringstream raw_binary_ss_;
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

// number of H1-Legendre basis functions of an element type at order p (getElementBasisFunctionNumber, src/Mesh/BasisFunction.cpp)
[[nodiscard]] inline constexpr int getElementBasisFunctionNumber(int type, int p) {
  switch (static_cast<ElementEnum>(type)) {
    case ElementEnum::Line: return p + 1;
    case ElementEnum::Triangle: return (p + 1) * (p + 2) / 2;
    case ElementEnum::Quadrangle: return (p + 1) * (p + 1);
    case ElementEnum::Hexahedron: return (p + 1) * (p + 1) * (p + 1);
    default: return 0;
  }
}

// ---- Solver<SC>: the drop-in seam (src/Solver/SolveControl.cpp:290-300,327-436) over the C ABI ---------------------------------------
template <typename SimulationControl>
struct SolverBase {   // SolveControl.cpp:290-302
  Real empirical_tolerance_{0.0};
  Real artificial_viscosity_factor_{1.0};
  std::stringstream raw_binary_ss_;
  std::fstream error_finout_;
  std::future<void> write_raw_binary_future_;
  Eigen::Vector<Real, SimulationControl::kConservedVariableNumber> relative_error_{
      Eigen::Vector<Real, SimulationControl::kConservedVariableNumber>::Zero()};
  Eigen::Vector<Real, Eigen::Dynamic> node_artificial_viscosity_;
};

template <typename SimulationControl>
struct Solver : SolverBase<SimulationControl> {
  inline static constexpr int kNv{SimulationControl::kConservedVariableNumber};
  // what stands in for the reference's per-type ElementSolver / AdjacencyElementSolver members: the device context and its block list
  sdg_ctx* ctx_{nullptr};
  int device_{0};
  std::vector<int> types_;   // ascending ElementEnum, the order of writeRawBinary
  std::vector<double> boundary_coordinate_;

  Solver() = default;
  Solver(const Solver&) = delete;
  Solver& operator=(const Solver&) = delete;
  ~Solver() {
    if (this->write_raw_binary_future_.valid()) this->write_raw_binary_future_.wait();
    if (ctx_ != nullptr) sdg_destroy(ctx_);
  }

  static void check(int rc) { if (rc != 0) throw std::runtime_error(std::string("subrosadg_b200: ") + sdg_last_error()); }

  struct Sizes { int n, Nb, Nq, Nf, Nqf, Nv; };
  [[nodiscard]] inline Sizes sizes(int type) const {
    int32_t sz[8]; check(sdg_sizes(ctx_, type, sz));
    return Sizes{sz[0], sz[1], sz[2], sz[4], sz[6], sz[7]};
  }

  inline void createContext(const Mesh<SimulationControl>& mesh, const PhysicalModel<SimulationControl>& physical_model) {
    if constexpr (SimulationControl::kLimiter != LimiterEnum::None) {
      throw std::runtime_error("subrosadg_b200: LimiterEnum::PositivityPreserving is not built on the B200 path");
    }
    sdg_config cfg{};
    cfg.dim = SimulationControl::kDimension; cfg.p = SimulationControl::kPolynomialOrder;
    cfg.model = static_cast<int>(SimulationControl::kEquationModel);
    cfg.eos = static_cast<int>(SimulationControl::kEquationOfState);
    cfg.transport = static_cast<int>(SimulationControl::kTransportModel);
    cfg.conv_flux = static_cast<int>(SimulationControl::kConvectiveFlux);
    cfg.visc_flux = static_cast<int>(SimulationControl::kViscousFlux);
    cfg.source = static_cast<int>(SimulationControl::kSourceTerm);
    cfg.rk = static_cast<int>(SimulationControl::kTimeIntegration);
    cfg.device = device_; cfg.chunk = 0; cfg.reorder = 1;
    cfg.cp = physical_model.thermodynamic_model_.specific_heat_constant_pressure;
    cfg.cv = physical_model.thermodynamic_model_.specific_heat_constant_volume;
    cfg.c0 = 1.0; cfg.rho0 = 1.0;
    if constexpr (SimulationControl::kTransportModel != TransportModelEnum::None) cfg.mu = physical_model.transport_model_.dynamic_viscosity;
    if constexpr (SimulationControl::kEquationOfState == EquationOfStateEnum::WeakCompressibleFluid) {
      cfg.c0 = physical_model.equation_of_state_.reference_sound_speed; cfg.rho0 = physical_model.equation_of_state_.reference_density;
    }
    if constexpr (SimulationControl::kSourceTerm == SourceTermEnum::Boussinesq) {   // `inline static` in the reference as well (SourceTerm.cpp:31-33)
      cfg.beta = SourceTerm<SimulationControl>::thermal_expansion_coefficient; cfg.t_ref = SourceTerm<SimulationControl>::reference_temperature;
    }
    check(sdg_create(&cfg, &ctx_));
    constexpr bool kAV = SimulationControl::kShockCapturing == ShockCapturingEnum::ArtificialViscosity;
    if constexpr (kAV) {   // System::setArtificialViscosity (SystemControl.cpp:105-108) -> empirical_tolerance_, artificial_viscosity_factor_
      check(sdg_set_artificial_viscosity(ctx_, this->empirical_tolerance_, this->artificial_viscosity_factor_, static_cast<int32_t>(mesh.node_number_)));
    }
    for (const auto& b : mesh.blocks) {
      check(sdg_add_elements(ctx_, b.type, b.n, 0, b.geom_order, b.coords.data()));
      types_.push_back(b.type);
      if constexpr (kAV) {   // mesh data of Solver::calculateArtificialViscosity: node_tag_ of the corners, inner_radius_
        const std::vector<int32_t> tags = mesh.nodeTags(b);
        const std::vector<double> radius = mesh.innerRadius(b);
        check(sdg_set_element_nodes(ctx_, b.type, tags.data(), radius.data()));
      }
    }
    std::sort(types_.begin(), types_.end());
    check(sdg_set_faces(ctx_, mesh.n_int, mesh.n_bnd, mesh.le.data(), mesh.lt.data(), mesh.lf.data(), mesh.re.data(), mesh.rt.data(), mesh.rf.data(),
                        mesh.rot.data(), mesh.bc.data(), mesh.phys.data()));
    check(sdg_finalize(ctx_));
  }

  // Solver::initializeSolver, SolveControl.cpp:377-380 / InitialCondition.cpp:151-186 (Function: :85-116; LastStep / SpecificFile: :41-80)
  inline void initializeSolver(const Mesh<SimulationControl>& mesh, const PhysicalModel<SimulationControl>& physical_model,
                               const BoundaryCondition<SimulationControl>& boundary_condition,
                               InitialCondition<SimulationControl>& initial_condition) {
    constexpr int D = SimulationControl::kDimension, NP = SimulationControl::kPrimitiveVariableNumber;
    constexpr bool kNS = SimulationControl::kViscousFlux != ViscousFluxEnum::None;
    this->node_artificial_viscosity_.resize(mesh.node_number_);
    this->node_artificial_viscosity_.setZero();
    createContext(mesh, physical_model);
    for (int t : types_) {
      const Sizes s = sizes(t);
      if constexpr (SimulationControl::kInitialCondition == InitialConditionEnum::Function) {
        const std::size_t npt = static_cast<std::size_t>(s.n) * static_cast<std::size_t>(s.Nq);
        std::vector<double> xq(npt * D), prim(npt * NP);
        check(sdg_get_quadrature_coordinates(ctx_, t, xq.data()));
        for (std::size_t i = 0; i < npt; i++) {
          Eigen::Vector<Real, D> x; for (int d = 0; d < D; d++) x[d] = xq[i * D + static_cast<std::size_t>(d)];
          const auto p = initial_condition.calculatePrimitiveFromCoordinate(x);
          for (int k = 0; k < NP; k++) prim[i * NP + static_cast<std::size_t>(k)] = p[k];
        }
        check(sdg_set_state_from_primitive(ctx_, t, prim.data()));
      } else {
        // LastStep: the file holds this order's coefficients; SpecificFile: those of a run at order P-1, which are the leading
        // columns of the hierarchical basis at order P (the remaining ones start from zero)
        const int nb_file = SimulationControl::kInitialCondition == InitialConditionEnum::LastStep
                                ? s.Nb : getElementBasisFunctionNumber(t, SimulationControl::kPolynomialOrder - 1);
        const std::size_t row = static_cast<std::size_t>(s.Nb) * kNv, row_file = static_cast<std::size_t>(nb_file) * kNv;
        std::vector<double> u(static_cast<std::size_t>(s.n) * row, 0.0), skip(row_file * D);
        for (int e = 0; e < s.n; e++) {
          initial_condition.raw_binary_ss_.read(reinterpret_cast<char*>(u.data() + static_cast<std::size_t>(e) * row),
                                                static_cast<std::streamsize>(row_file * sizeof(double)));
          if constexpr (kNS) initial_condition.raw_binary_ss_.read(reinterpret_cast<char*>(skip.data()), static_cast<std::streamsize>(skip.size() * sizeof(double)));
          if (!initial_condition.raw_binary_ss_) throw std::runtime_error("raw binary initial condition is shorter than the mesh needs");
        }
        check(sdg_set_state(ctx_, t, u.data()));
      }
    }
    if (mesh.n_bnd > 0) {
      const int nqf = sizes(mesh.lt[static_cast<std::size_t>(mesh.n_int)]).Nqf;
      boundary_coordinate_.resize(static_cast<std::size_t>(mesh.n_bnd) * static_cast<std::size_t>(nqf) * D);
      check(sdg_get_boundary_quadrature_coordinates(ctx_, boundary_coordinate_.data()));
      TimeIntegration<SimulationControl> at_start;
      updateBoundaryVariable(mesh, physical_model, boundary_condition, at_start);
    }
  }

  // Solver::updateBoundaryVariable, SolveControl.cpp:382-385 / BoundaryCondition.cpp:29-74: the primitive boundary values at the
  // boundary quadrature points, at t = iteration_ * delta_time_ for BoundaryTimeEnum::TimeVarying
  inline void updateBoundaryVariable(const Mesh<SimulationControl>& mesh, [[maybe_unused]] const PhysicalModel<SimulationControl>& physical_model,
                                     const BoundaryCondition<SimulationControl>& boundary_condition,
                                     const TimeIntegration<SimulationControl>& time_integration) {
    constexpr int D = SimulationControl::kDimension, NP = SimulationControl::kPrimitiveVariableNumber;
    if (mesh.n_bnd == 0) return;
    const std::size_t npt = boundary_coordinate_.size() / D, nqf = npt / static_cast<std::size_t>(mesh.n_bnd);
    std::vector<double> prim(npt * NP);
    [[maybe_unused]] const Real time = static_cast<Real>(time_integration.iteration_) * time_integration.delta_time_;
    for (std::size_t i = 0; i < npt; i++) {
      Eigen::Vector<Real, D> x; for (int d = 0; d < D; d++) x[d] = boundary_coordinate_[i * D + static_cast<std::size_t>(d)];
      const Isize phys = mesh.phys[static_cast<std::size_t>(mesh.n_int) + i / nqf];
      Eigen::Vector<Real, NP> p;
      if constexpr (SimulationControl::kBoundaryTime == BoundaryTimeEnum::TimeVarying) p = boundary_condition.calculatePrimitiveFromCoordinate(x, time, phys);
      else p = boundary_condition.calculatePrimitiveFromCoordinate(x, phys);
      for (int k = 0; k < NP; k++) prim[i * NP + static_cast<std::size_t>(k)] = p[k];
    }
    check(sdg_set_boundary_primitive(ctx_, prim.data()));
  }

  // Solver::calculateDeltaTime, SolveControl.cpp:389-391 / TimeIntegration.cpp:104-179
  inline void calculateDeltaTime([[maybe_unused]] const Mesh<SimulationControl>& mesh, [[maybe_unused]] const PhysicalModel<SimulationControl>& physical_model,
                                 TimeIntegration<SimulationControl>& time_integration) {
    check(sdg_compute_dt(ctx_, time_integration.courant_friedrichs_lewy_number_, &time_integration.delta_time_));
  }

  // Solver::stepSolver, SolveControl.cpp:427-431 / TimeIntegration.cpp:326-350
  inline void stepSolver(const Mesh<SimulationControl>& mesh, [[maybe_unused]] const SourceTerm<SimulationControl>& source_term,
                         const PhysicalModel<SimulationControl>& physical_model, const BoundaryCondition<SimulationControl>& boundary_condition,
                         const TimeIntegration<SimulationControl>& time_integration) {
    if constexpr (SimulationControl::kBoundaryTime == BoundaryTimeEnum::TimeVarying) {
      updateBoundaryVariable(mesh, physical_model, boundary_condition, time_integration);
    }
    check(sdg_step(ctx_, time_integration.delta_time_, 1, this->relative_error_.data()));
  }

  // The same step for a caller that keeps the modal coefficients [n][Nb][Nv] of block `type` in HOST memory, as the reference's
  // variable_basis_function_coefficient_ does (sdg_step_host: upload, stages and download streamed; coefficient_in and coefficient_out may
  // be the same array; bit-identical to sdg_set_state -> stepSolver -> sdg_get_state).
  inline void stepSolverHost(const Mesh<SimulationControl>& mesh, const PhysicalModel<SimulationControl>& physical_model,
                             const BoundaryCondition<SimulationControl>& boundary_condition, const TimeIntegration<SimulationControl>& time_integration,
                             int type, const double* coefficient_in, double* coefficient_out) {
    if constexpr (SimulationControl::kBoundaryTime == BoundaryTimeEnum::TimeVarying) {
      updateBoundaryVariable(mesh, physical_model, boundary_condition, time_integration);
    }
    check(sdg_step_host(ctx_, type, time_integration.delta_time_, coefficient_in, coefficient_out, this->relative_error_.data()));
  }

  // Solver::writeRawBinary, RawBinary.cpp:156-191: per element type (ascending ElementEnum) and element the modal coefficients
  // [Nb][Nv] and, for Navier-Stokes, the gradient coefficients [Nb][Nv*D] (:75-88); per boundary face (face order) the same two
  // blocks of its parent, the gradient being BR1: total, BR2: volume part + the lift of that face (:89-154); node_number_ reals of
  // node artificial viscosity.  The stream is compressed and written by a std::async task, joined before the next write.
  inline void writeRawBinary(const Mesh<SimulationControl>& mesh, const std::filesystem::path& raw_binary_path) {
    constexpr int D = SimulationControl::kDimension;
    constexpr bool kNS = SimulationControl::kViscousFlux != ViscousFluxEnum::None;
    if (this->write_raw_binary_future_.valid()) this->write_raw_binary_future_.get();   // the task owns raw_binary_ss_ until it is done
    std::map<int, std::vector<double>> U, G;
    std::map<int, Sizes> S;
    for (int t : types_) {
      const Sizes s = sizes(t); S[t] = s;
      U[t].resize(static_cast<std::size_t>(s.n) * static_cast<std::size_t>(s.Nb) * kNv);
      check(sdg_get_state(ctx_, t, U[t].data()));
      if constexpr (kNS) {
        G[t].resize(U[t].size() * D);
        check(sdg_get_gradient_state(ctx_, t, G[t].data()));
      }
    }
    std::stringstream& ss = this->raw_binary_ss_;
    auto put = [&](const double* p, std::size_t n) { ss.write(reinterpret_cast<const char*>(p), static_cast<std::streamsize>(n * sizeof(double))); };
    for (int t : types_) {
      const std::size_t row = static_cast<std::size_t>(S[t].Nb) * kNv;
      for (int e = 0; e < S[t].n; e++) {
        put(U[t].data() + static_cast<std::size_t>(e) * row, row);
        if constexpr (kNS) put(G[t].data() + static_cast<std::size_t>(e) * row * D, row * D);
      }
    }
    if (mesh.n_bnd > 0) {
      std::vector<double> Gb;
      if constexpr (kNS) {
        std::size_t n = 0;
        for (int i = mesh.n_int; i < mesh.n_int + mesh.n_bnd; i++) n += static_cast<std::size_t>(S[mesh.lt[static_cast<std::size_t>(i)]].Nb) * kNv * D;
        Gb.resize(n);
        check(sdg_get_boundary_gradient_state(ctx_, Gb.data()));
      }
      std::size_t at = 0;
      for (int i = mesh.n_int; i < mesh.n_int + mesh.n_bnd; i++) {
        const int t = mesh.lt[static_cast<std::size_t>(i)], e = mesh.le[static_cast<std::size_t>(i)];
        const std::size_t row = static_cast<std::size_t>(S[t].Nb) * kNv;
        put(U[t].data() + static_cast<std::size_t>(e) * row, row);
        if constexpr (kNS) { put(Gb.data() + at, row * D); at += row * D; }
      }
    }
    if constexpr (SimulationControl::kShockCapturing == ShockCapturingEnum::ArtificialViscosity) {
      check(sdg_get_node_artificial_viscosity(ctx_, this->node_artificial_viscosity_.data()));   // as of the last step (zero before the first)
    }
    put(this->node_artificial_viscosity_.data(), static_cast<std::size_t>(mesh.node_number_));
    this->write_raw_binary_future_ = std::async(std::launch::async, RawBinaryCompress::write, raw_binary_path, std::ref(this->raw_binary_ss_));
  }

  // views of the device state for drivers and tests
  inline std::vector<double> getCoefficient(int type) const {
    const Sizes s = sizes(type);
    std::vector<double> u(static_cast<std::size_t>(s.n) * static_cast<std::size_t>(s.Nb) * static_cast<std::size_t>(s.Nv));
    check(sdg_get_state(ctx_, type, u.data()));
    return u;
  }
  inline std::vector<double> getStateAtQuadrature(int type) const {
    const Sizes s = sizes(type);
    std::vector<double> u(static_cast<std::size_t>(s.n) * static_cast<std::size_t>(s.Nq) * static_cast<std::size_t>(s.Nv));
    check(sdg_get_state_at_quadrature(ctx_, type, u.data()));
    return u;
  }
  inline std::vector<double> getQuadratureCoordinate(int type) const {
    const Sizes s = sizes(type);
    std::vector<double> x(static_cast<std::size_t>(s.n) * static_cast<std::size_t>(s.Nq) * static_cast<std::size_t>(SimulationControl::kDimension));
    check(sdg_get_quadrature_coordinates(ctx_, type, x.data()));
    return x;
  }
};

// ---- what System::solve needs of View / CommandLine (src/View/IOControl.cpp:318-337, src/View/CommandLine.cpp:70-140) ------------------
template <typename SimulationControl>
struct View {
  std::filesystem::path output_directory_;
  std::string output_file_name_prefix_;
  int io_interval_{0};
  int iteration_order_{0};
  std::vector<ViewVariableEnum> variable_type_;

  inline void initializeSolverFinout(const bool delete_dir, std::fstream& error_finout) {
    const std::filesystem::path raw_output_directory = output_directory_ / "raw";
    std::ios::openmode open_mode = std::ios::in | std::ios::out;
    if (delete_dir && SimulationControl::kInitialCondition != InitialConditionEnum::LastStep) {
      std::filesystem::remove_all(raw_output_directory);
      open_mode |= std::ios::trunc;
    }
    std::filesystem::create_directories(raw_output_directory);
    if (!std::filesystem::exists(output_directory_ / "error.txt")) open_mode |= std::ios::trunc;
    error_finout.open((output_directory_ / "error.txt").string(), open_mode);
  }
  inline void finalizeSolverFinout(std::fstream& error_finout) { error_finout.close(); }
  [[nodiscard]] inline std::filesystem::path rawBinaryPath(int step) const {
    return output_directory_ / ("raw/" + output_file_name_prefix_ + "_" + std::to_string(step) + ".zst");
  }
};

template <typename SimulationControl>
struct CommandLine {
  bool is_open_{true};
  Real delta_time_{0.0};

  [[nodiscard]] inline static std::string centred(const std::string& s) {   // std::format's {:^13}
    const std::size_t w = 13, pad = s.size() < w ? w - s.size() : 0;
    return std::string(pad / 2, ' ') + s + std::string(pad - pad / 2, ' ');
  }
  [[nodiscard]] inline static std::string number(Real x) { char b[32]; std::snprintf(b, sizeof b, "%.5e", x); return centred(b); }
  [[nodiscard]] inline std::string getVariableList() const {
    constexpr bool kIncompressible = SimulationControl::kEquationModel == EquationModelEnum::IncompresibleEuler ||
                                     SimulationControl::kEquationModel == EquationModelEnum::IncompresibleNS;
    static const char* const mom[3] = {"rho*u", "rho*v", "rho*w"};
    std::string line = "|" + centred("Time") + "|" + centred("rho") + "|";
    for (int d = 0; d < SimulationControl::kDimension; d++) line += centred(mom[d]) + "|";
    return line + centred(kIncompressible ? "rho*e" : "rho*E") + "|";
  }
  [[nodiscard]] inline std::string getLineInformation(const Real time_value,
                                                      const Eigen::Vector<Real, SimulationControl::kConservedVariableNumber>& error) const {
    std::string line = "|" + number(time_value) + "|";
    for (int v = 0; v < SimulationControl::kConservedVariableNumber; v++) line += number(error(v)) + "|";
    return line;
  }
  inline void initializeSolver(const TimeIntegration<SimulationControl>& time_integration, std::fstream& error_finout) {
    delta_time_ = time_integration.delta_time_;
    if constexpr (SimulationControl::kInitialCondition != InitialConditionEnum::LastStep) {
      error_finout << getVariableList() << '\n'
                   << getLineInformation(0.0, Eigen::Vector<Real, SimulationControl::kConservedVariableNumber>::Zero()) << '\n';
    } else {   // keep the header and the lines up to iteration_start_, continue after them
      error_finout.seekg(0, std::ios::beg);
      std::string line;
      for (int i = 0; i < time_integration.iteration_start_ + 2 && std::getline(error_finout, line); i++) {}
      error_finout.clear();
      error_finout.seekp(error_finout.tellg());
    }
  }
  inline void updateSolver(const int step, const Eigen::Vector<Real, SimulationControl::kConservedVariableNumber>& new_error, std::fstream& error_finout) {
    error_finout << getLineInformation(static_cast<Real>(step) * delta_time_, new_error) << '\n';
  }
};

// ---- System<SC>: src/Utils/SystemControl.cpp:55-231 -------------------------------------------------------------------------------------
template <typename SimulationControl>
struct System {
  Mesh<SimulationControl> mesh_;
  SourceTerm<SimulationControl> source_term_;
  PhysicalModel<SimulationControl> physical_model_;
  BoundaryCondition<SimulationControl> boundary_condition_;
  InitialCondition<SimulationControl> initial_condition_;
  TimeIntegration<SimulationControl> time_integration_;
  Solver<SimulationControl> solver_;
  View<SimulationControl> view_;
  CommandLine<SimulationControl> command_line_;
  std::map<Isize, BoundaryConditionEnum> physical_boundary_;

  // SystemControl.cpp:60-66: the mesh generator writes the file, the mesh is read from it.  Gmsh is not available, so the file is the
  // flat format (MeshData::writeFlat / subrosadg_b200.mesh.write_flat) and the generator fills it from an in-code producer.
  inline void setMesh(const std::filesystem::path& mesh_file_path,
                      const std::function<void(const std::filesystem::path& mesh_file_path)>& generate_mesh_function) {
    if constexpr (SimulationControl::kInitialCondition != InitialConditionEnum::LastStep) generate_mesh_function(mesh_file_path);
    mesh_.initializeMesh(mesh_file_path);
  }
  inline void setMesh(MeshData mesh) { static_cast<MeshData&>(mesh_) = std::move(mesh); mesh_.countNodes(); }
  inline void setMesh(const std::filesystem::path& flat_mesh_file) { mesh_.initializeMesh(flat_mesh_file); }

  template <SourceTermEnum SourceTermType>
    requires(SourceTermType == SourceTermEnum::Boussinesq)
  inline void setSourceTerm(const Real thermal_expansion_coefficient, const Real reference_temperature) {
    source_term_.thermal_expansion_coefficient = thermal_expansion_coefficient;
    source_term_.reference_temperature = reference_temperature;
  }
  template <InitialConditionEnum InitialConditionType>
    requires(InitialConditionType == InitialConditionEnum::SpecificFile)
  inline void addInitialCondition(const std::filesystem::path& initial_condition_file) { initial_condition_.raw_binary_path_ = initial_condition_file; }
  template <BoundaryConditionEnum BoundaryConditionType>
  inline void addBoundaryCondition(const Isize physical_index) { physical_boundary_[physical_index] = BoundaryConditionType; }
  template <ThermodynamicModelEnum ThermodynamicModelType>
    requires(ThermodynamicModelType == ThermodynamicModelEnum::Constant)
  inline void setThermodynamicModel(const Real specific_heat_constant_pressure, const Real specific_heat_constant_volume) {
    physical_model_.thermodynamic_model_.specific_heat_constant_pressure = specific_heat_constant_pressure;
    physical_model_.thermodynamic_model_.specific_heat_constant_volume = specific_heat_constant_volume;
  }
  template <EquationOfStateEnum EquationOfStateType>
    requires(EquationOfStateType == EquationOfStateEnum::WeakCompressibleFluid)
  inline void setEquationOfState(const Real reference_sound_speed, const Real reference_density) {
    physical_model_.equation_of_state_.reference_sound_speed = reference_sound_speed;
    physical_model_.equation_of_state_.reference_density = reference_density;
    physical_model_.equation_of_state_.calculatePressureAdditionFromSoundSpeedDensity();
  }
  template <TransportModelEnum TransportModelType>
    requires(TransportModelType == TransportModelEnum::Constant || TransportModelType == TransportModelEnum::Sutherland)
  inline void setTransportModel(const Real dynamic_viscosity) {
    physical_model_.transport_model_.dynamic_viscosity = dynamic_viscosity;
    physical_model_.calculateThermalConductivityFromDynamicViscosity();
  }
  inline void setArtificialViscosity(const Real empirical_tolerance, const Real artificial_viscosity_factor = 1.0) {
    solver_.empirical_tolerance_ = empirical_tolerance; solver_.artificial_viscosity_factor_ = artificial_viscosity_factor;
  }
  inline void setTimeIntegration(const Real courant_friedrichs_lewy_number, const std::pair<int, int> iteration_range = {0, 0}) {
    if (iteration_range.first == 0 && iteration_range.second == 0) {
      std::printf("\nSet time integration end number: ");
      if (std::scanf("%d", &time_integration_.iteration_end_) != 1) time_integration_.iteration_end_ = 0;
    } else {
      time_integration_.iteration_start_ = iteration_range.first; time_integration_.iteration_end_ = iteration_range.second;
    }
    time_integration_.courant_friedrichs_lewy_number_ = courant_friedrichs_lewy_number;
  }
  inline void setDeltaTime(const Real delta_time) { time_integration_.delta_time_ = delta_time; }
  inline void setViewConfig(const std::filesystem::path& output_directory, const std::string_view output_file_name_prefix, const int io_interval = 0) {
    if (io_interval == 0) {
      std::printf("Set view interval: ");
      if (std::scanf("%d", &view_.io_interval_) != 1 || view_.io_interval_ == -1) view_.io_interval_ = time_integration_.iteration_end_;
    } else if (io_interval == -1) {
      view_.io_interval_ = time_integration_.iteration_end_;
    } else {
      view_.io_interval_ = io_interval;
    }
    view_.iteration_order_ = static_cast<int>(std::log10(std::max(1, time_integration_.iteration_end_)) + 1);
    view_.output_directory_ = output_directory; view_.output_file_name_prefix_ = std::string(output_file_name_prefix);
  }
  inline void addViewVariable(const std::vector<ViewVariableEnum>& view_variable) { view_.variable_type_ = view_variable; }
  inline void setDevice(int device) { solver_.device_ = device; }   // B200 path only: which GPU the context binds

  // System::synchronize, SystemControl.cpp:142-157: boundary types onto the face records, raw-binary initial conditions read
  inline void synchronize() {
    for (int i = mesh_.n_int; i < mesh_.n_int + mesh_.n_bnd; i++) {
      const auto it = physical_boundary_.find(mesh_.phys[static_cast<std::size_t>(i)]);
      if (it == physical_boundary_.end()) throw std::runtime_error("boundary face without addBoundaryCondition for its physical index");
      mesh_.bc[static_cast<std::size_t>(i)] = static_cast<int>(it->second);
    }
    if constexpr (SimulationControl::kInitialCondition == InitialConditionEnum::SpecificFile) {
      RawBinaryCompress::read(initial_condition_.raw_binary_path_, initial_condition_.raw_binary_ss_);
    } else if constexpr (SimulationControl::kInitialCondition == InitialConditionEnum::LastStep) {
      initial_condition_.raw_binary_path_ = view_.rawBinaryPath(time_integration_.iteration_start_);
      RawBinaryCompress::read(initial_condition_.raw_binary_path_, initial_condition_.raw_binary_ss_);
    }
  }

  // System::solve, SystemControl.cpp:159-195
  inline void solve(const bool delete_dir = true) {
    view_.initializeSolverFinout(delete_dir, solver_.error_finout_);
    solver_.initializeSolver(mesh_, physical_model_, boundary_condition_, initial_condition_);
    if (time_integration_.delta_time_ == 0.0) solver_.calculateDeltaTime(mesh_, physical_model_, time_integration_);
    if constexpr (SimulationControl::kInitialCondition != InitialConditionEnum::LastStep) {
      solver_.writeRawBinary(mesh_, view_.rawBinaryPath(0));
    } else {
      solver_.write_raw_binary_future_ = std::async(std::launch::async, []() {});
    }
    command_line_.initializeSolver(time_integration_, solver_.error_finout_);
    const int io = std::max(1, view_.io_interval_);
    for (int i = time_integration_.iteration_start_ + 1; i <= time_integration_.iteration_end_; i++) {
      solver_.stepSolver(mesh_, source_term_, physical_model_, boundary_condition_, time_integration_);
      time_integration_.iteration_ = i;   // after the step, as SystemControl.cpp:175-177: step i sees t = (i - 1) dt
      if (i % io == 0) {
        solver_.write_raw_binary_future_.get();
        solver_.writeRawBinary(mesh_, view_.rawBinaryPath(i));
      }
      command_line_.updateSolver(i, solver_.relative_error_, solver_.error_finout_);
      bool all_nan = true;
      for (int v = 0; v < SimulationControl::kConservedVariableNumber; v++) all_nan = all_nan && std::isnan(solver_.relative_error_(v));
      if (command_line_.is_open_ && (i == time_integration_.iteration_end_ || i % io == 0 || all_nan)) {
        std::printf("%s\n", command_line_.getLineInformation(time_integration_.delta_time_ * i, solver_.relative_error_).c_str());
      }
      if (all_nan) {   // SystemControl.cpp:185-191
        if (view_.io_interval_ == time_integration_.iteration_end_) view_.io_interval_ = i;
        time_integration_.iteration_end_ = i;
        break;
      }
    }
    if (solver_.write_raw_binary_future_.valid()) solver_.write_raw_binary_future_.get();
    view_.finalizeSolverFinout(solver_.error_finout_);
  }
  inline void view([[maybe_unused]] const bool delete_dir = true) {}   // VTU output: out of scope (host post-processing of the raw files)
};

}  // namespace SubrosaDG

using SubrosaDG::operator""_r;
using SubrosaDG::operator""_deg;

#endif  // SUBROSA_DG_B200_HPP_
