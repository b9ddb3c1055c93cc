/**
 * SolveControlB200.cpp — the reference-side binding of libsubrosadg_b200.so: a replacement for Solver<SimulationControl>
 * (src/Solver/SolveControl.cpp:327-436) written against the REFERENCE'S OWN types (Mesh<SC>, PhysicalModel<SC>, BoundaryCondition<SC>,
 * InitialCondition<SC>, SourceTerm<SC>, TimeIntegration<SC>, SolverBase<SC>, RawBinaryCompress).  A maintainer drops this file into
 * src/Solver/, includes it from src/SubrosaDG.cpp after View/RawBinary.cpp, and declares
 *     SolverB200<SimulationControl> solver_;            // instead of Solver<SimulationControl> solver_ (src/Utils/SystemControl.cpp:52)
 * System<SC>::solve() (SystemControl.cpp:159-195) is unchanged: it only calls initializeSolver / calculateDeltaTime / stepSolver /
 * writeRawBinary and reads relative_error_, error_finout_ and write_raw_binary_future_, all of which exist here with the same
 * signatures.  View<SC> keeps reading the raw files with its own ViewSolver (the payload is the reference's, RawBinary.cpp:75-191).
 *
 * This file is compiled by tests/test_integration_binding.py against the headers under /root/reference/src (with the declaration-level
 * stand-ins of oracle/ref_shim for Eigen, Gmsh, oneTBB, magic_enum, zstd), for every MeshModelEnum the B200 path supports.
 */
#ifndef SUBROSA_DG_SOLVE_CONTROL_B200_CPP_
#define SUBROSA_DG_SOLVE_CONTROL_B200_CPP_

#include <algorithm>
#include <filesystem>
#include <functional>
#include <future>
#include <stdexcept>
#include <string>
#include <vector>

#include "Mesh/ReadControl.cpp"
#include "Solver/BoundaryCondition.cpp"
#include "Solver/InitialCondition.cpp"
#include "Solver/PhysicalModel.cpp"
#include "Solver/SimulationControl.cpp"
#include "Solver/SolveControl.cpp"
#include "Solver/SourceTerm.cpp"
#include "Solver/TimeIntegration.cpp"
#include "Utils/BasicDataType.cpp"
#include "Utils/Concept.cpp"
#include "Utils/Enum.cpp"
#include "View/RawBinary.cpp"

extern "C" {
#include "subrosadg_b200.h"
}

namespace SubrosaDG {

template <typename SimulationControl>
struct SolverB200 : SolverBase<SimulationControl> {   // relative_error_, error_finout_, raw_binary_ss_, write_raw_binary_future_, ... (:290-302)
  inline static constexpr int kD{SimulationControl::kDimension};
  inline static constexpr int kNv{SimulationControl::kConservedVariableNumber};
  inline static constexpr int kP{SimulationControl::kPolynomialOrder};
  inline static constexpr MeshModelEnum kModel{SimulationControl::kMeshModel};
  static_assert(kModel == MeshModelEnum::Line || kModel == MeshModelEnum::Triangle || kModel == MeshModelEnum::Quadrangle ||
                    kModel == MeshModelEnum::TriangleQuadrangle || kModel == MeshModelEnum::Hexahedron,
                "the B200 path has no tetrahedron / pyramid kernels");
  static_assert(SimulationControl::kLimiter == LimiterEnum::None, "the positivity limiter is not built on the B200 path");
  static_assert(SimulationControl::kShockCapturing == ShockCapturingEnum::None || IsEuler<SimulationControl::kEquationModel>,
                "artificial viscosity is built for the Euler models");
  inline static constexpr bool kAV{SimulationControl::kShockCapturing == ShockCapturingEnum::ArtificialViscosity};

  sdg_ctx* ctx_{nullptr};
  int device_{0};
  std::vector<double> boundary_coordinate_;      // [n_bnd][Nqf][D], the order of the adjacency element meshes
  std::vector<Isize> boundary_physical_index_;   // [n_bnd]

  SolverB200() = default;
  SolverB200(const SolverB200&) = delete;
  SolverB200& operator=(const SolverB200&) = delete;
  ~SolverB200() {
    if (this->write_raw_binary_future_.valid()) this->write_raw_binary_future_.wait();
    if (ctx_ != nullptr) sdg_destroy(ctx_);
  }

  static void check(const int rc) {
    if (rc != 0) throw std::runtime_error(std::string("subrosadg_b200: ") + sdg_last_error());
  }

  // ---- the reference's per-type members, visited in ascending ElementEnum order (the order of writeRawBinary, RawBinary.cpp:156-191) ----
  template <typename Function>
  inline static void forEachElementMesh(const Mesh<SimulationControl>& mesh, Function&& function) {
    if constexpr (kD == 1) {
      function(mesh.line_, LineTrait<kP>{});
    } else if constexpr (kD == 2) {
      if constexpr (HasTriangle<kModel>) function(mesh.triangle_, TriangleTrait<kP>{});
      if constexpr (HasQuadrangle<kModel>) function(mesh.quadrangle_, QuadrangleTrait<kP>{});
    } else {
      function(mesh.hexahedron_, HexahedronTrait<kP>{});
    }
  }
  // the (D-1)-dimensional adjacency mesh of the supported models: points, lines or quadrangles
  inline static const auto& adjacencyMesh(const Mesh<SimulationControl>& mesh) {
    if constexpr (kD == 1) return mesh.point_;
    else if constexpr (kD == 2) return mesh.line_;
    else return mesh.quadrangle_;
  }
  // parent_gmsh_type_number_ (ReadControl.cpp:79) -> ElementEnum value, the type key of the C ABI
  [[nodiscard]] inline static int elementEnumOfGmshType(const Isize gmsh_type_number) {
    if constexpr (kD == 1) return magic_enum::enum_integer(ElementEnum::Line);
    if constexpr (kD == 2) {
      return gmsh_type_number == TriangleTrait<kP>::kGmshTypeNumber ? magic_enum::enum_integer(ElementEnum::Triangle)
                                                                   : magic_enum::enum_integer(ElementEnum::Quadrangle);
    }
    return magic_enum::enum_integer(ElementEnum::Hexahedron);
  }

  struct Sizes { int n, Nb, Nq, Nqf; };
  [[nodiscard]] inline Sizes sizes(const int type) const {
    int32_t s[8];
    check(sdg_sizes(ctx_, type, s));
    return Sizes{s[0], s[1], s[2], s[6]};
  }

  // Solver::initializeSolver, SolveControl.cpp:377-380 / InitialCondition.cpp:151-186
  inline void initializeSolver(const Mesh<SimulationControl>& mesh, const PhysicalModel<SimulationControl>& physical_model,
                               const BoundaryCondition<SimulationControl>& boundary_condition,
                               InitialCondition<SimulationControl>& initial_condition) {
    this->node_artificial_viscosity_.resize(mesh.node_number_);
    this->node_artificial_viscosity_.setZero();
    sdg_config cfg{};
    cfg.dim = kD;
    cfg.p = kP;
    cfg.model = magic_enum::enum_integer(SimulationControl::kEquationModel);
    cfg.eos = magic_enum::enum_integer(SimulationControl::kEquationOfState);
    cfg.transport = magic_enum::enum_integer(SimulationControl::kTransportModel);
    cfg.conv_flux = magic_enum::enum_integer(SimulationControl::kConvectiveFlux);
    if constexpr (IsNS<SimulationControl::kEquationModel>) cfg.visc_flux = magic_enum::enum_integer(SimulationControl::kViscousFlux);   // the Euler variable sets have no kViscousFlux
    cfg.source = magic_enum::enum_integer(SimulationControl::kSourceTerm);
    cfg.rk = magic_enum::enum_integer(SimulationControl::kTimeIntegration);
    cfg.device = device_;
    cfg.reorder = 1;
    cfg.cp = physical_model.thermodynamic_model_.specific_heat_constant_pressure;   // PhysicalModel.cpp:28-29
    cfg.cv = physical_model.thermodynamic_model_.specific_heat_constant_volume;
    cfg.c0 = 1.0;
    cfg.rho0 = 1.0;
    if constexpr (SimulationControl::kTransportModel != TransportModelEnum::None) {
      cfg.mu = physical_model.transport_model_.dynamic_viscosity;   // :88-89; the library derives k = cp mu / Pr like :118-122
    }
    if constexpr (SimulationControl::kEquationOfState == EquationOfStateEnum::WeakCompressibleFluid) {
      cfg.c0 = physical_model.equation_of_state_.reference_sound_speed;
      cfg.rho0 = physical_model.equation_of_state_.reference_density;
    }
    if constexpr (SimulationControl::kSourceTerm == SourceTermEnum::Boussinesq) {   // inline static, SourceTerm.cpp:31-33
      cfg.beta = SourceTerm<SimulationControl>::thermal_expansion_coefficient;
      cfg.t_ref = SourceTerm<SimulationControl>::reference_temperature;
    }
    check(sdg_create(&cfg, &ctx_));
    if constexpr (kAV) {   // System::setArtificialViscosity has filled the two SolverBase fields (SystemControl.cpp:105-108)
      check(sdg_set_artificial_viscosity(ctx_, this->empirical_tolerance_, this->artificial_viscosity_factor_, static_cast<int32_t>(mesh.node_number_)));
    }
    // one block per element type: node coordinates in Gmsh node order (PerElementMesh::node_coordinate_, ReadControl.cpp:86-88)
    forEachElementMesh(mesh, [&]<typename ElementTrait>(const ElementMesh<ElementTrait>& element_mesh, ElementTrait) {
      std::vector<double> x(static_cast<std::size_t>(element_mesh.number_) * ElementTrait::kAllNodeNumber * kD);
      std::size_t at = 0;
      for (Isize i = 0; i < element_mesh.number_; i++) {
        for (Isize k = 0; k < ElementTrait::kAllNodeNumber; k++) {
          for (Isize d = 0; d < kD; d++) x[at++] = element_mesh.element_(i).node_coordinate_(d, k);
        }
      }
      check(sdg_add_elements(ctx_, magic_enum::enum_integer(ElementTrait::kElementType), static_cast<int32_t>(element_mesh.number_), 0, kP, x.data()));
      if constexpr (kAV) {   // what Solver::calculateArtificialViscosity reads of the mesh (SpatialDiscrete.cpp:75-78,96-99): node_tag_ (1-based) of the corners, inner_radius_
        std::vector<int32_t> tags(static_cast<std::size_t>(element_mesh.number_) * ElementTrait::kBasicNodeNumber);
        std::vector<double> radius(static_cast<std::size_t>(element_mesh.number_));
        for (Isize i = 0; i < element_mesh.number_; i++) {
          for (Isize j = 0; j < ElementTrait::kBasicNodeNumber; j++) {
            tags[static_cast<std::size_t>(i * ElementTrait::kBasicNodeNumber + j)] = static_cast<int32_t>(element_mesh.element_(i).node_tag_(j) - 1);
          }
          radius[static_cast<std::size_t>(i)] = element_mesh.element_(i).inner_radius_;
        }
        check(sdg_set_element_nodes(ctx_, magic_enum::enum_integer(ElementTrait::kElementType), tags.data(), radius.data()));
      }
    });
    // AdjacencyElementMesh records (ReadControl.cpp:72-83), interior faces first, then boundary faces
    const auto& adjacency = adjacencyMesh(mesh);
    const Isize n_int = adjacency.interior_number_, n_bnd = adjacency.boundary_number_;
    std::vector<int32_t> le, lt, lf, re, rt, rf, rot, bc, phys;
    for (Isize i = 0; i < n_int + n_bnd; i++) {
      const auto& face = adjacency.element_(i);
      const bool interior = i < n_int;
      le.push_back(static_cast<int32_t>(face.parent_index_each_type_(0)));
      lt.push_back(elementEnumOfGmshType(face.parent_gmsh_type_number_(0)));
      lf.push_back(static_cast<int32_t>(face.adjacency_sequence_in_parent_(0)));
      re.push_back(interior ? static_cast<int32_t>(face.parent_index_each_type_(1)) : -1);
      rt.push_back(interior ? elementEnumOfGmshType(face.parent_gmsh_type_number_(1)) : 0);
      rf.push_back(interior ? static_cast<int32_t>(face.adjacency_sequence_in_parent_(1)) : 0);
      rot.push_back(interior ? static_cast<int32_t>(face.adjacency_right_rotation_) : 0);
      bc.push_back(interior ? magic_enum::enum_integer(BoundaryConditionEnum::Periodic) : magic_enum::enum_integer(face.boundary_condition_type_));
      phys.push_back(static_cast<int32_t>(face.gmsh_physical_index_));
      if (!interior) boundary_physical_index_.push_back(face.gmsh_physical_index_);
    }
    check(sdg_set_faces(ctx_, static_cast<int32_t>(n_int), static_cast<int32_t>(n_bnd), le.data(), lt.data(), lf.data(), re.data(), rt.data(),
                        rf.data(), rot.data(), bc.data(), phys.data()));
    check(sdg_finalize(ctx_));
    // the user callbacks stay on the host (InitialCondition.cpp:85-116); raw-binary initial conditions go in as modal coefficients (:41-80)
    forEachElementMesh(mesh, [&]<typename ElementTrait>(const ElementMesh<ElementTrait>& element_mesh, ElementTrait) {
      const int type = magic_enum::enum_integer(ElementTrait::kElementType);
      const Sizes s = sizes(type);
      if constexpr (SimulationControl::kInitialCondition == InitialConditionEnum::Function) {
        const std::size_t npt = static_cast<std::size_t>(s.n) * static_cast<std::size_t>(s.Nq);
        std::vector<double> xq(npt * kD), prim(npt * SimulationControl::kPrimitiveVariableNumber);
        check(sdg_get_quadrature_coordinates(ctx_, type, xq.data()));
        for (std::size_t i = 0; i < npt; i++) {
          Eigen::Vector<Real, kD> coordinate;
          for (Isize d = 0; d < kD; d++) coordinate(d) = xq[i * kD + static_cast<std::size_t>(d)];
          const Eigen::Vector<Real, SimulationControl::kPrimitiveVariableNumber> primitive =
              initial_condition.calculatePrimitiveFromCoordinate(coordinate);
          for (Isize k = 0; k < SimulationControl::kPrimitiveVariableNumber; k++) {
            prim[i * SimulationControl::kPrimitiveVariableNumber + static_cast<std::size_t>(k)] = primitive(k);
          }
        }
        check(sdg_set_state_from_primitive(ctx_, type, prim.data()));
      } else {
        Eigen::Array<Eigen::Matrix<Real, kNv, ElementTrait::kBasisFunctionNumber>, Eigen::Dynamic, 1> coefficient;
        coefficient.resize(element_mesh.number_);
        initial_condition.getVariableBasisFunctionCoefficient(element_mesh, coefficient);   // the reference's own reader
        std::vector<double> u(static_cast<std::size_t>(s.n) * static_cast<std::size_t>(s.Nb) * kNv);
        for (Isize i = 0; i < element_mesh.number_; i++) {
          std::copy_n(coefficient(i).data(), s.Nb * kNv, u.data() + static_cast<std::size_t>(i) * static_cast<std::size_t>(s.Nb) * kNv);
        }
        check(sdg_set_state(ctx_, type, u.data()));
      }
    });
    if (n_bnd > 0) {
      const int nqf = sizes(lt[static_cast<std::size_t>(n_int)]).Nqf;
      boundary_coordinate_.resize(static_cast<std::size_t>(n_bnd) * static_cast<std::size_t>(nqf) * kD);
      check(sdg_get_boundary_quadrature_coordinates(ctx_, boundary_coordinate_.data()));
      TimeIntegration<SimulationControl> at_start;
      updateBoundaryVariable(mesh, physical_model, boundary_condition, at_start);
    }
  }

  // Solver::updateBoundaryVariable, SolveControl.cpp:382-385 / BoundaryCondition.cpp:29-74
  inline void updateBoundaryVariable([[maybe_unused]] const Mesh<SimulationControl>& mesh,
                                     [[maybe_unused]] const PhysicalModel<SimulationControl>& physical_model,
                                     const BoundaryCondition<SimulationControl>& boundary_condition,
                                     const TimeIntegration<SimulationControl>& time_integration) {
    constexpr int kNp = SimulationControl::kPrimitiveVariableNumber;
    if (boundary_physical_index_.empty()) return;
    const std::size_t npt = boundary_coordinate_.size() / kD, nqf = npt / boundary_physical_index_.size();
    std::vector<double> prim(npt * kNp);
    for (std::size_t i = 0; i < npt; i++) {
      Eigen::Vector<Real, kD> coordinate;
      for (Isize d = 0; d < kD; d++) coordinate(d) = boundary_coordinate_[i * kD + static_cast<std::size_t>(d)];
      Eigen::Vector<Real, kNp> primitive;
      if constexpr (SimulationControl::kBoundaryTime == BoundaryTimeEnum::TimeVarying) {
        primitive = boundary_condition.calculatePrimitiveFromCoordinate(
            coordinate, static_cast<Real>(time_integration.iteration_) * time_integration.delta_time_, boundary_physical_index_[i / nqf]);
      } else {
        primitive = boundary_condition.calculatePrimitiveFromCoordinate(coordinate, boundary_physical_index_[i / nqf]);
      }
      for (Isize k = 0; k < kNp; k++) prim[i * kNp + static_cast<std::size_t>(k)] = primitive(k);
    }
    check(sdg_set_boundary_primitive(ctx_, prim.data()));
  }

  // Solver::calculateDeltaTime, SolveControl.cpp:389-391 / TimeIntegration.cpp:104-179
  inline void calculateDeltaTime([[maybe_unused]] const Mesh<SimulationControl>& mesh,
                                 [[maybe_unused]] const PhysicalModel<SimulationControl>& physical_model,
                                 TimeIntegration<SimulationControl>& time_integration) {
    check(sdg_compute_dt(ctx_, time_integration.courant_friedrichs_lewy_number_, &time_integration.delta_time_));
  }

  // Solver::stepSolver, SolveControl.cpp:427-431 / TimeIntegration.cpp:326-350
  inline void stepSolver(const Mesh<SimulationControl>& mesh, [[maybe_unused]] const SourceTerm<SimulationControl>& source_term,
                         const PhysicalModel<SimulationControl>& physical_model,
                         const BoundaryCondition<SimulationControl>& boundary_condition,
                         const TimeIntegration<SimulationControl>& time_integration) {
    if constexpr (SimulationControl::kBoundaryTime == BoundaryTimeEnum::TimeVarying) {
      updateBoundaryVariable(mesh, physical_model, boundary_condition, time_integration);
    }
    check(sdg_step(ctx_, time_integration.delta_time_, 1, this->relative_error_.data()));   // NaNs propagate (SystemControl.cpp:185-191)
  }

  // Solver::writeRawBinary, SolveControl.cpp:435 / RawBinary.cpp:75-191: device -> host at output steps only, the reference's payload
  // order, the reference's own RawBinaryCompress::write on the reference's own std::async task
  inline void writeRawBinary(const Mesh<SimulationControl>& mesh, const std::filesystem::path& raw_binary_path) {
    constexpr bool kNS = IsNS<SimulationControl::kEquationModel>;
    std::vector<std::vector<double>> state(7), gradient(7);
    std::vector<int> basis_function_number(7, 0);
    auto put = [&](const double* p, const std::size_t n) {
      this->raw_binary_ss_.write(reinterpret_cast<const char*>(p), static_cast<std::streamsize>(n * kRealSize));
    };
    forEachElementMesh(mesh, [&]<typename ElementTrait>(const ElementMesh<ElementTrait>&, ElementTrait) {
      const int type = magic_enum::enum_integer(ElementTrait::kElementType);
      const Sizes s = sizes(type);
      const std::size_t row = static_cast<std::size_t>(s.Nb) * kNv;
      basis_function_number[static_cast<std::size_t>(type)] = s.Nb;
      std::vector<double>& u = state[static_cast<std::size_t>(type)];
      std::vector<double>& g = gradient[static_cast<std::size_t>(type)];
      u.resize(static_cast<std::size_t>(s.n) * row);
      check(sdg_get_state(ctx_, type, u.data()));                       // [n][Nb][Nv] = Eigen::Matrix<Real, Nv, Nb>, column major
      if constexpr (kNS) {
        g.resize(u.size() * kD);
        check(sdg_get_gradient_state(ctx_, type, g.data()));            // [n][Nb][Nv*D]
      }
      for (int i = 0; i < s.n; i++) {                                   // ElementSolver::writeElementRawBinary, RawBinary.cpp:75-88
        put(u.data() + static_cast<std::size_t>(i) * row, row);
        if constexpr (kNS) put(g.data() + static_cast<std::size_t>(i) * row * kD, row * kD);
      }
    });
    const auto& adjacency = adjacencyMesh(mesh);                        // writeBoundaryAdjacencyElementRawBinary, :89-154
    if (adjacency.boundary_number_ > 0) {
      std::vector<double> boundary_gradient;
      if constexpr (kNS) {
        std::size_t n = 0;
        for (Isize i = adjacency.interior_number_; i < adjacency.interior_number_ + adjacency.boundary_number_; i++) {
          const int type = elementEnumOfGmshType(adjacency.element_(i).parent_gmsh_type_number_(0));
          n += static_cast<std::size_t>(basis_function_number[static_cast<std::size_t>(type)]) * kNv * kD;
        }
        boundary_gradient.resize(n);
        check(sdg_get_boundary_gradient_state(ctx_, boundary_gradient.data()));   // BR1: total, BR2: volume + this face's lift
      }
      std::size_t at = 0;
      for (Isize i = adjacency.interior_number_; i < adjacency.interior_number_ + adjacency.boundary_number_; i++) {
        const int type = elementEnumOfGmshType(adjacency.element_(i).parent_gmsh_type_number_(0));
        const std::size_t row = static_cast<std::size_t>(basis_function_number[static_cast<std::size_t>(type)]) * kNv;
        put(state[static_cast<std::size_t>(type)].data() + static_cast<std::size_t>(adjacency.element_(i).parent_index_each_type_(0)) * row, row);
        if constexpr (kNS) {
          put(boundary_gradient.data() + at, row * kD);
          at += row * kD;
        }
      }
    }
    if constexpr (kAV) check(sdg_get_node_artificial_viscosity(ctx_, this->node_artificial_viscosity_.data()));
    this->raw_binary_ss_.write(reinterpret_cast<const char*>(this->node_artificial_viscosity_.data()), mesh.node_number_ * kRealSize);
    this->write_raw_binary_future_ =
        std::async(std::launch::async, RawBinaryCompress::write, raw_binary_path, std::ref(this->raw_binary_ss_));
  }
};

}  // namespace SubrosaDG

#endif  // SUBROSA_DG_SOLVE_CONTROL_B200_CPP_
