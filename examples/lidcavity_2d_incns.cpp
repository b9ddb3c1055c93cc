/**
 * lidcavity_2d_incns — the lid-driven cavity of a weakly compressible fluid (reference example lidcavity_2d_incns) on the B200 path.
 *
 * Mirrors /root/reference/examples/lidcavity_2d_incns.cpp where the surface allows: same SimulationControl typedef, same InitialCondition /
 * BoundaryCondition specialisations, same System setter sequence and values.  Differences: the include, and generateMesh() (Gmsh is not
 * available here) is replaced by a flat mesh file written by the in-code producer `python -m subrosadg_b200.mesh lidcavity_2d <file>`; the
 * iteration count is an argument because the reference reads it from std::cin.
 *
 * usage: lidcavity_2d_incns mesh.sdgm [iterations=10] [state_out_prefix]   (state_out_prefix.<ElementEnum>.bin = conserved variables at the volume
 *        quadrature points of each element type, [n][Nq][Nv])
 */
#include "SubrosaDG_b200/SubrosaDG.hpp"

#include <cstdlib>
#include <iostream>

inline const std::string kExampleName{"lidcavity_2d_incns"};

using SimulationControl = SubrosaDG::SimulationControl<SubrosaDG::SolveControl<SubrosaDG::DimensionEnum::D2,
    SubrosaDG::PolynomialOrderEnum::P3, SubrosaDG::BoundaryTimeEnum::Steady, SubrosaDG::SourceTermEnum::None>,
    SubrosaDG::NumericalControl<SubrosaDG::MeshModelEnum::Quadrangle, SubrosaDG::ShockCapturingEnum::None,
        SubrosaDG::LimiterEnum::None, SubrosaDG::InitialConditionEnum::Function, SubrosaDG::TimeIntegrationEnum::SSPRK3>,
    SubrosaDG::IncompresibleNSVariable<SubrosaDG::ThermodynamicModelEnum::Constant, SubrosaDG::EquationOfStateEnum::WeakCompressibleFluid,
        SubrosaDG::TransportModelEnum::Constant, SubrosaDG::ConvectiveFluxEnum::LaxFriedrichs, SubrosaDG::ViscousFluxEnum::BR2>>;

template <typename SC>
using Primitive = Eigen::Vector<SubrosaDG::Real, SC::kPrimitiveVariableNumber>;

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::InitialCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    [[maybe_unused]] const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate) const {
  return Primitive<SimulationControl>{1.0_r, 0.0_r, 0.0_r, 1.0_r};
}

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::BoundaryCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    [[maybe_unused]] const Eigen::Vector<SubrosaDG::Real, SimulationControl::kDimension>& coordinate,
    const SubrosaDG::Isize gmsh_physical_index) const {
  if (gmsh_physical_index == 1) {
    return Primitive<SimulationControl>{1.0_r, 0.0_r, 0.0_r, 1.0_r};
  }
  if (gmsh_physical_index == 2) {
    return Primitive<SimulationControl>{1.0_r, 1.0_r, 0.0_r, 1.0_r};
  }
  return Primitive<SimulationControl>::Zero();
}

int main(int argc, char* argv[]) {
  if (argc < 2) {
    std::cerr << "usage: " << kExampleName << " mesh.sdgm [iterations] [state_out_prefix]\n";
    return EXIT_FAILURE;
  }
  const int iterations = argc > 2 ? std::atoi(argv[2]) : 10;
  SubrosaDG::System<SimulationControl> system;
  system.setMesh(std::filesystem::path(argv[1]));
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::AdiabaticNonSlipWall>(1);
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::AdiabaticNonSlipWall>(2);
  system.setThermodynamicModel<SimulationControl::kThermodynamicModel>(1.0_r, 1.0_r);
  system.setEquationOfState<SimulationControl::kEquationOfState>(10.0_r, 1.0_r);
  system.setTransportModel<SimulationControl::kTransportModel>(1.0_r * 1.0_r * 1.0_r / 5000.0_r);
  system.setTimeIntegration(1.0_r, {0, iterations});
  system.setViewConfig("build/out/" + kExampleName, kExampleName, -1);
  system.addViewVariable({SubrosaDG::ViewVariableEnum::Density, SubrosaDG::ViewVariableEnum::Velocity,
      SubrosaDG::ViewVariableEnum::Pressure, SubrosaDG::ViewVariableEnum::Temperature,
      SubrosaDG::ViewVariableEnum::MachNumber, SubrosaDG::ViewVariableEnum::Vorticity});
  system.synchronize();
  system.solve();
  system.view();
  if (argc > 3) {
    for (int type : system.solver_.types_) {
      const std::vector<double> u = system.solver_.getStateAtQuadrature(type);
      std::ofstream f(std::string(argv[3]) + "." + std::to_string(type) + ".bin", std::ios::binary);
      f.write(reinterpret_cast<const char*>(u.data()), static_cast<std::streamsize>(u.size() * sizeof(double)));
    }
  }
  std::cout << "delta_time " << system.time_integration_.delta_time_ << "\n";
  return EXIT_SUCCESS;
}
