/**
 * periodic_2d_ceuler — BASELINE.json configs[0] on the B200 path.
 *
 * Mirrors /root/reference/examples/periodic_2d_ceuler.cpp line for line where the surface allows: same SimulationControl
 * typedef (:19-26), same InitialCondition / BoundaryCondition specialisations (:28-44), same System setter sequence (:49-59).
 * Differences: the include, and generateMesh() (gmsh transfinite 10x10 recombined square with periodic sides, :64-96) is
 * written with the in-code producer makePeriodicBox into the flat mesh format (Gmsh is not available here); setTimeIntegration /
 * setViewConfig get their iteration count explicitly because the reference reads it from std::cin.  Output (raw/<prefix>_<step>.zst, error.txt)
 * goes to build/out/periodic_2d_ceuler under the working directory.
 *
 * usage: periodic_2d_ceuler [iterations=100] [state_out.bin]
 */
#include "SubrosaDG_b200/SubrosaDG.hpp"

#include <cstdlib>
#include <iostream>

inline const std::string kExampleName{"periodic_2d_ceuler"};

inline const std::filesystem::path kExampleDirectory{std::filesystem::path("build/out") / kExampleName};

void generateMesh(const std::filesystem::path& mesh_file_path);

using SimulationControl = SubrosaDG::SimulationControl<SubrosaDG::SolveControl<SubrosaDG::DimensionEnum::D2,
    SubrosaDG::PolynomialOrderEnum::P3, SubrosaDG::BoundaryTimeEnum::Steady, SubrosaDG::SourceTermEnum::None>,
    SubrosaDG::NumericalControl<SubrosaDG::MeshModelEnum::Quadrangle, SubrosaDG::ShockCapturingEnum::None,
        SubrosaDG::LimiterEnum::None, SubrosaDG::InitialConditionEnum::Function, SubrosaDG::TimeIntegrationEnum::SSPRK3>,
    SubrosaDG::CompresibleEulerVariable<SubrosaDG::ThermodynamicModelEnum::Constant, SubrosaDG::EquationOfStateEnum::IdealGas,
        SubrosaDG::ConvectiveFluxEnum::HLLC>>;

template <typename SC>
using Primitive = Eigen::Vector<SubrosaDG::Real, SC::kPrimitiveVariableNumber>;

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::InitialCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate) const {
  return Primitive<SimulationControl>{1.0_r + 0.2_r * std::sin(SubrosaDG::kPi * (coordinate.x() + coordinate.y())), 0.7_r,
      0.3_r,
      1.4_r / (1.0_r + 0.2_r * std::sin(SubrosaDG::kPi * (coordinate.x() + coordinate.y())))};
}

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::BoundaryCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    [[maybe_unused]] const Eigen::Vector<SubrosaDG::Real, SimulationControl::kDimension>& coordinate,
    [[maybe_unused]] const SubrosaDG::Isize gmsh_physical_index) const {
  return Primitive<SimulationControl>::Zero();
}

int main(int argc, char* argv[]) {
  const int iterations = argc > 1 ? std::atoi(argv[1]) : 100;
  SubrosaDG::System<SimulationControl> system;
  system.setMesh(kExampleDirectory / "periodic_2d_ceuler.sdgm", generateMesh);
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::Periodic>(1);
  system.setThermodynamicModel<SimulationControl::kThermodynamicModel>(2.5_r, 25.0_r / 14.0_r);
  system.setTimeIntegration(1.0_r, {0, iterations});
  system.setDeltaTime(1.0e-03_r);
  system.setViewConfig(kExampleDirectory, kExampleName, -1);
  system.addViewVariable({SubrosaDG::ViewVariableEnum::Density, SubrosaDG::ViewVariableEnum::Velocity,
      SubrosaDG::ViewVariableEnum::Pressure});
  system.synchronize();
  system.solve();
  system.view();

  // known answer: the density wave is advected unchanged, rho(x, t) = 1 + 0.2 sin(pi (x + y - (u + v) t)), p = 1
  const int type = static_cast<int>(SubrosaDG::ElementEnum::Quadrangle);
  const std::vector<double> x = system.solver_.getQuadratureCoordinate(type);
  const std::vector<double> u = system.solver_.getStateAtQuadrature(type);
  const double t = system.time_integration_.delta_time_ * system.time_integration_.iteration_end_;
  double err = 0.0, ref = 0.0;
  const std::size_t npt = x.size() / 2;
  for (std::size_t i = 0; i < npt; i++) {
    const double exact = 1.0 + 0.2 * std::sin(SubrosaDG::kPi * (x[2 * i] + x[2 * i + 1] - 1.0 * t));
    err += (u[4 * i] - exact) * (u[4 * i] - exact);
    ref += exact * exact;
  }
  std::cout << "density rel-L2 error against the exact travelling wave at t = " << t << ": " << std::sqrt(err / ref) << "\n";
  if (argc > 2) {
    std::ofstream f(argv[2], std::ios::binary);
    f.write(reinterpret_cast<const char*>(u.data()), static_cast<std::streamsize>(u.size() * sizeof(double)));
  }
  return EXIT_SUCCESS;
}

// the reference meshes the [0,2]^2 square with a 10 x 10 transfinite recombined Gmsh grid with periodic sides (:64-96); here the same
// grid comes from the in-code producer, through the mesh file like in the reference
void generateMesh(const std::filesystem::path& mesh_file_path) {
  std::filesystem::create_directories(mesh_file_path.parent_path());
  SubrosaDG::makePeriodicBox(SimulationControl::kDimension, 10, 0.0, 2.0).writeFlat(mesh_file_path);
}
