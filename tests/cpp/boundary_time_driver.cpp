// Pins the time a BoundaryTimeEnum::TimeVarying callback sees: step i runs with t = (i - 1) * delta_time_ because System::solve assigns
// iteration_ after stepSolver (SystemControl.cpp:175-177, BoundaryCondition.cpp:29-74).  usage: boundary_time_driver MESH.sdgm OUT_DIR STEPS [host]
#include "SubrosaDG_b200/SubrosaDG.hpp"

#include <cstdlib>
#include <iostream>
#include <string>

using SimulationControl = SubrosaDG::SimulationControl<SubrosaDG::SolveControl<SubrosaDG::DimensionEnum::D2,
    SubrosaDG::PolynomialOrderEnum::P2, SubrosaDG::BoundaryTimeEnum::TimeVarying, SubrosaDG::SourceTermEnum::None>,
    SubrosaDG::NumericalControl<SubrosaDG::MeshModelEnum::Quadrangle, SubrosaDG::ShockCapturingEnum::None,
        SubrosaDG::LimiterEnum::None, SubrosaDG::InitialConditionEnum::Function, SubrosaDG::TimeIntegrationEnum::SSPRK3>,
    SubrosaDG::CompresibleEulerVariable<SubrosaDG::ThermodynamicModelEnum::Constant, SubrosaDG::EquationOfStateEnum::IdealGas,
        SubrosaDG::ConvectiveFluxEnum::HLLC>>;

template <typename SC>
using Primitive = Eigen::Vector<SubrosaDG::Real, SC::kPrimitiveVariableNumber>;

inline std::vector<double> seen_times;

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::InitialCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    [[maybe_unused]] const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate) const {
  return Primitive<SimulationControl>{1.4_r, 0.3_r, 0.1_r, 1.0_r};
}

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::BoundaryCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    [[maybe_unused]] const Eigen::Vector<SubrosaDG::Real, SimulationControl::kDimension>& coordinate, const SubrosaDG::Real time,
    [[maybe_unused]] const SubrosaDG::Isize gmsh_physical_index) const {
  if (seen_times.empty() || seen_times.back() != time) seen_times.push_back(time);
  return Primitive<SimulationControl>{1.4_r, 0.3_r * (1.0_r + 5.0_r * time), 0.1_r, 1.0_r};
}

int main(int argc, char* argv[]) {
  if (argc < 4) return 2;
  SubrosaDG::System<SimulationControl> system;
  system.command_line_.is_open_ = false;
  system.setMesh(std::filesystem::path(argv[1]));
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::RiemannFarfield>(1);
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::RiemannFarfield>(2);
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::RiemannFarfield>(3);
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::RiemannFarfield>(4);
  system.setThermodynamicModel<SimulationControl::kThermodynamicModel>(2.5_r, 25.0_r / 14.0_r);
  system.setTimeIntegration(1.0_r, {0, std::atoi(argv[3])});
  system.setDeltaTime(2.0e-03_r);
  system.setViewConfig(std::filesystem::path(argv[2]), "bt", -1);
  system.synchronize();
  system.solve();
  std::cout << "times";
  for (double t : seen_times) std::cout << " " << t;
  std::cout << "\n";
  if (argc > 4 && std::string(argv[4]) == "host") {
    // one more step on coefficients held in host memory (Solver::stepSolverHost over sdg_step_host), in place
    const int quad = static_cast<int>(SubrosaDG::ElementEnum::Quadrangle);
    std::vector<double> c = system.solver_.getCoefficient(quad);
    system.solver_.stepSolverHost(system.mesh_, system.physical_model_, system.boundary_condition_, system.time_integration_, quad, c.data(), c.data());
    std::ofstream g(std::filesystem::path(argv[2]) / "coefficient_host.bin", std::ios::binary);
    g.write(reinterpret_cast<const char*>(c.data()), static_cast<std::streamsize>(c.size() * sizeof(double)));
  }
  const std::vector<double> u = system.solver_.getStateAtQuadrature(static_cast<int>(SubrosaDG::ElementEnum::Quadrangle));
  std::ofstream f(std::filesystem::path(argv[2]) / "state.bin", std::ios::binary);
  f.write(reinterpret_cast<const char*>(u.data()), static_cast<std::streamsize>(u.size() * sizeof(double)));
  return EXIT_SUCCESS;
}
