// Restart paths of System<SC> over the raw files (InitialConditionEnum::Function / LastStep / SpecificFile, SystemControl.cpp:142-195,
// InitialCondition.cpp:41-80) on a periodic Navier-Stokes case (BR1, so that the files carry gradient blocks the readers must skip).
// build: -DIC_KIND=Function|LastStep|SpecificFile -DPOLY=P2|P3
// usage: restart_driver OUT_DIR START END IO_INTERVAL COEFFICIENT_OUT [SPECIFIC_FILE]
#include "SubrosaDG_b200/SubrosaDG.hpp"

#include <cstdlib>
#include <iostream>

#ifndef IC_KIND
#define IC_KIND Function
#endif
#ifndef POLY
#define POLY P3
#endif

using SimulationControl = SubrosaDG::SimulationControl<SubrosaDG::SolveControl<SubrosaDG::DimensionEnum::D2,
    SubrosaDG::PolynomialOrderEnum::POLY, SubrosaDG::BoundaryTimeEnum::Steady, SubrosaDG::SourceTermEnum::None>,
    SubrosaDG::NumericalControl<SubrosaDG::MeshModelEnum::Quadrangle, SubrosaDG::ShockCapturingEnum::None,
        SubrosaDG::LimiterEnum::None, SubrosaDG::InitialConditionEnum::IC_KIND, SubrosaDG::TimeIntegrationEnum::SSPRK3>,
    SubrosaDG::CompresibleNSVariable<SubrosaDG::ThermodynamicModelEnum::Constant, SubrosaDG::EquationOfStateEnum::IdealGas,
        SubrosaDG::TransportModelEnum::Constant, SubrosaDG::ConvectiveFluxEnum::HLLC, SubrosaDG::ViscousFluxEnum::BR1>>;

template <typename SC>
using Primitive = Eigen::Vector<SubrosaDG::Real, SC::kPrimitiveVariableNumber>;

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::InitialCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate) const {
  const Real rho = 1.0_r + 0.2_r * std::sin(SubrosaDG::kPi * (coordinate.x() + coordinate.y()));
  return Primitive<SimulationControl>{rho, 0.7_r, 0.3_r, 1.4_r / rho};
}

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::BoundaryCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    [[maybe_unused]] const Eigen::Vector<SubrosaDG::Real, SimulationControl::kDimension>& coordinate,
    [[maybe_unused]] const SubrosaDG::Isize gmsh_physical_index) const {
  return Primitive<SimulationControl>::Zero();
}

int main(int argc, char* argv[]) {
  if (argc < 6) return 2;
  const std::filesystem::path out(argv[1]);
  SubrosaDG::System<SimulationControl> system;
  system.command_line_.is_open_ = false;
  system.setMesh(out / "mesh.sdgm", [](const std::filesystem::path& p) {
    std::filesystem::create_directories(p.parent_path());
    SubrosaDG::makePeriodicBox(2, 6, 0.0, 2.0).writeFlat(p);
  });
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::Periodic>(1);
  system.setThermodynamicModel<SimulationControl::kThermodynamicModel>(2.5_r, 25.0_r / 14.0_r);
  system.setTransportModel<SimulationControl::kTransportModel>(2.0e-03_r);
  system.setTimeIntegration(1.0_r, {std::atoi(argv[2]), std::atoi(argv[3])});
  system.setDeltaTime(1.0e-03_r);
  system.setViewConfig(out, "run", std::atoi(argv[4]));
  [&](auto& sys) {   // a generic lambda, so that the discarded branch is not instantiated
    using SC = SimulationControl;
    if constexpr (SC::kInitialCondition == SubrosaDG::InitialConditionEnum::SpecificFile) {
      sys.template addInitialCondition<SC::kInitialCondition>(std::filesystem::path(argv[6]));
    }
  }(system);
  system.synchronize();
  system.solve();
  const std::vector<double> u = system.solver_.getCoefficient(static_cast<int>(SubrosaDG::ElementEnum::Quadrangle));
  std::ofstream f(argv[5], std::ios::binary);
  f.write(reinterpret_cast<const char*>(u.data()), static_cast<std::streamsize>(u.size() * sizeof(double)));
  std::cout << "node_number " << system.mesh_.node_number_ << " iteration " << system.time_integration_.iteration_ << "\n";
  return EXIT_SUCCESS;
}
