/**
 * thermalcavity_2d_incns — the differentially heated cavity with the Boussinesq source (reference example thermalcavity_2d_incns; P1 quadrangles: seven-point rule) on the B200 path.
 *
 * Mirrors /root/reference/examples/thermalcavity_2d_incns.cpp where the surface allows: same SimulationControl typedef, same InitialCondition /
 * BoundaryCondition specialisations, same System setter sequence and values.  Differences: the include, and generateMesh() (Gmsh is not
 * available here) is replaced by a flat mesh file written by the in-code producer `python -m subrosadg_b200.mesh thermalcavity_2d <file>`; the
 * iteration count is an argument because the reference reads it from std::cin.
 *
 * usage: thermalcavity_2d_incns mesh.sdgm [iterations=10] [state_out_prefix]   (state_out_prefix.<ElementEnum>.bin = conserved variables at the volume
 *        quadrature points of each element type, [n][Nq][Nv])
 */
#include "SubrosaDG_b200/SubrosaDG.hpp"

#include <cstdlib>
#include <iostream>

inline const std::string kExampleName{"thermalcavity_2d_incns"};

using SimulationControl = SubrosaDG::SimulationControl<SubrosaDG::SolveControl<SubrosaDG::DimensionEnum::D2,
    SubrosaDG::PolynomialOrderEnum::P1, SubrosaDG::BoundaryTimeEnum::Steady, SubrosaDG::SourceTermEnum::Boussinesq>,
    SubrosaDG::NumericalControl<SubrosaDG::MeshModelEnum::Quadrangle, SubrosaDG::ShockCapturingEnum::None,
        SubrosaDG::LimiterEnum::None, SubrosaDG::InitialConditionEnum::Function, SubrosaDG::TimeIntegrationEnum::SSPRK3>,
    SubrosaDG::IncompresibleNSVariable<SubrosaDG::ThermodynamicModelEnum::Constant, SubrosaDG::EquationOfStateEnum::WeakCompressibleFluid,
        SubrosaDG::TransportModelEnum::Constant, SubrosaDG::ConvectiveFluxEnum::Exact, SubrosaDG::ViscousFluxEnum::BR2>>;

template <typename SC>
using Primitive = Eigen::Vector<SubrosaDG::Real, SC::kPrimitiveVariableNumber>;

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::InitialCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    [[maybe_unused]] const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate) const {
  return Primitive<SimulationControl>{1.0_r, 0.0_r, 0.0_r, 0.5_r};
}

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::BoundaryCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    [[maybe_unused]] const Eigen::Vector<SubrosaDG::Real, SimulationControl::kDimension>& coordinate,
    const SubrosaDG::Isize gmsh_physical_index) const {
  if (gmsh_physical_index == 1) {
    return Primitive<SimulationControl>{1.0_r, 0.0_r, 0.0_r, 0.5_r};
  }
  if (gmsh_physical_index == 2) {
    return Primitive<SimulationControl>{1.0_r, 0.0_r, 0.0_r, 0.0_r};
  }
  if (gmsh_physical_index == 3) {
    return Primitive<SimulationControl>{1.0_r, 0.0_r, 0.0_r, 1.0_r};
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
  system.setSourceTerm<SimulationControl::kSourceTerm>(1.0_r, 0.5_r);
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::AdiabaticNonSlipWall>(1);
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::IsoThermalNonSlipWall>(2);
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::IsoThermalNonSlipWall>(3);
  system.setThermodynamicModel<SimulationControl::kThermodynamicModel>(1.0_r, 1.0_r);
  // Ra = 1e6 , c0 = 3.0 ; Ra = 1e7 , c0 = 5.0
  system.setEquationOfState<SimulationControl::kEquationOfState>(3.0_r, 1.0_r);
  system.setTransportModel<SimulationControl::kTransportModel>(std::sqrt(0.71_r / 1e6_r));
  system.setTimeIntegration(0.5_r, {0, iterations});
  system.setViewConfig("build/out/" + kExampleName, kExampleName, -1);
  system.addViewVariable({SubrosaDG::ViewVariableEnum::Density, SubrosaDG::ViewVariableEnum::Velocity,
      SubrosaDG::ViewVariableEnum::Pressure, SubrosaDG::ViewVariableEnum::Temperature,
      SubrosaDG::ViewVariableEnum::MachNumber, SubrosaDG::ViewVariableEnum::Vorticity,
      SubrosaDG::ViewVariableEnum::HeatFlux});
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
