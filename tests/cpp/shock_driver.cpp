// sod_1d_ceuler-style run through System<SC> with ShockCapturingEnum::ArtificialViscosity (examples/sod_1d_ceuler.cpp:19-60 of the reference:
// Line mesh, Riemann far field at both ends, setArtificialViscosity).  usage: shock_driver MESH.sdgm OUT_DIR STEPS
#include "SubrosaDG_b200/SubrosaDG.hpp"

#include <cstdlib>
#include <iostream>

using SimulationControl = SubrosaDG::SimulationControl<SubrosaDG::SolveControl<SubrosaDG::DimensionEnum::D1,
    SubrosaDG::PolynomialOrderEnum::P2, SubrosaDG::BoundaryTimeEnum::Steady, SubrosaDG::SourceTermEnum::None>,
    SubrosaDG::NumericalControl<SubrosaDG::MeshModelEnum::Line, SubrosaDG::ShockCapturingEnum::ArtificialViscosity,
        SubrosaDG::LimiterEnum::None, SubrosaDG::InitialConditionEnum::Function, SubrosaDG::TimeIntegrationEnum::SSPRK3>,
    SubrosaDG::CompresibleEulerVariable<SubrosaDG::ThermodynamicModelEnum::Constant, SubrosaDG::EquationOfStateEnum::IdealGas,
        SubrosaDG::ConvectiveFluxEnum::HLLC>>;

template <typename SC>
using Primitive = Eigen::Vector<SubrosaDG::Real, SC::kPrimitiveVariableNumber>;

// density / pressure steps 1 -> 0.125, 1 -> 0.1 smoothed inside one cell: IEEE operations only, so that the Python mirror reproduces the bits
inline Primitive<SimulationControl> sodState(const SubrosaDG::Real x) {
  SubrosaDG::Real s = (x - 0.51_r) / 0.01_r;
  s = s < -1.0_r ? -1.0_r : (s > 1.0_r ? 1.0_r : s);
  const SubrosaDG::Real rho = 0.5625_r - 0.4375_r * s, p = 0.55_r - 0.45_r * s;
  return Primitive<SimulationControl>{rho, 0.0_r, 1.4_r * p / rho};
}

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::InitialCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate) const {
  return sodState(coordinate.x());
}

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::BoundaryCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    const Eigen::Vector<SubrosaDG::Real, SimulationControl::kDimension>& coordinate,
    [[maybe_unused]] const SubrosaDG::Isize gmsh_physical_index) const {
  return sodState(coordinate.x());
}

int main(int argc, char* argv[]) {
  if (argc < 4) return 2;
  SubrosaDG::System<SimulationControl> system;
  system.command_line_.is_open_ = false;
  system.setMesh(std::filesystem::path(argv[1]));
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::RiemannFarfield>(1);
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::RiemannFarfield>(2);
  system.setThermodynamicModel<SimulationControl::kThermodynamicModel>(2.5_r, 25.0_r / 14.0_r);
  system.setArtificialViscosity(0.5_r);
  system.setTimeIntegration(0.1_r, {0, std::atoi(argv[3])});
  system.setViewConfig(std::filesystem::path(argv[2]), "sod", -1);
  system.synchronize();
  system.solve();
  const std::vector<double> u = system.solver_.getStateAtQuadrature(static_cast<int>(SubrosaDG::ElementEnum::Line));
  std::ofstream f(std::filesystem::path(argv[2]) / "state.bin", std::ios::binary);
  f.write(reinterpret_cast<const char*>(u.data()), static_cast<std::streamsize>(u.size() * sizeof(double)));
  std::cout << "node_number " << system.mesh_.node_number_ << " delta_time " << system.time_integration_.delta_time_ << "\n";
  return EXIT_SUCCESS;
}
