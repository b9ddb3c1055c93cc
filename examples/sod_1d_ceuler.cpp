/**
 * sod_1d_ceuler — the Sod shock tube with artificial viscosity (reference example sod_1d_ceuler) on the B200 path.
 *
 * Mirrors /root/reference/examples/sod_1d_ceuler.cpp where the surface allows: same SimulationControl typedef, same InitialCondition /
 * BoundaryCondition specialisations, same System setter sequence and values.  Differences: the include, and generateMesh() (Gmsh is not
 * available here) is replaced by a flat mesh file written by the in-code producer `python -m subrosadg_b200.mesh sod_1d <file>`; the
 * iteration count is an argument because the reference reads it from std::cin.
 *
 * usage: sod_1d_ceuler mesh.sdgm [iterations=10] [state_out_prefix]   (state_out_prefix.<ElementEnum>.bin = conserved variables at the volume
 *        quadrature points of each element type, [n][Nq][Nv])
 */
#include "SubrosaDG_b200/SubrosaDG.hpp"

#include <cstdlib>
#include <iostream>

inline const std::string kExampleName{"sod_1d_ceuler"};

using SimulationControl = SubrosaDG::SimulationControl<SubrosaDG::SolveControl<SubrosaDG::DimensionEnum::D1,
    SubrosaDG::PolynomialOrderEnum::P3, SubrosaDG::BoundaryTimeEnum::Steady, SubrosaDG::SourceTermEnum::None>,
    SubrosaDG::NumericalControl<SubrosaDG::MeshModelEnum::Line, SubrosaDG::ShockCapturingEnum::ArtificialViscosity,
        SubrosaDG::LimiterEnum::None, SubrosaDG::InitialConditionEnum::Function, SubrosaDG::TimeIntegrationEnum::SSPRK3>,
    SubrosaDG::CompresibleEulerVariable<SubrosaDG::ThermodynamicModelEnum::Constant, SubrosaDG::EquationOfStateEnum::IdealGas,
        SubrosaDG::ConvectiveFluxEnum::HLLC>>;

template <typename SC>
using Primitive = Eigen::Vector<SubrosaDG::Real, SC::kPrimitiveVariableNumber>;

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::InitialCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    [[maybe_unused]] const Eigen::Vector<Real, SimulationControl::kDimension>& coordinate) const {
  return Primitive<SimulationControl>{coordinate.x() <= 0.5_r ? 1.0_r : 0.125_r, coordinate.x() <= 0.5_r ? 0.75_r : 0.0_r,
                                      coordinate.x() <= 0.5_r ? 1.4_r : 0.8_r * 1.4_r};
}

template <typename SimulationControl>
inline Primitive<SimulationControl> SubrosaDG::BoundaryCondition<SimulationControl>::calculatePrimitiveFromCoordinate(
    [[maybe_unused]] const Eigen::Vector<SubrosaDG::Real, SimulationControl::kDimension>& coordinate,
    const SubrosaDG::Isize gmsh_physical_index) const {
  if (gmsh_physical_index == 1) {
    return Primitive<SimulationControl>{1.0_r, 0.75_r, 1.4_r};
  }
  if (gmsh_physical_index == 2) {
    return Primitive<SimulationControl>{0.125_r, 0.0_r, 0.8_r * 1.4_r};
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
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::RiemannFarfield>(1);
  system.addBoundaryCondition<SubrosaDG::BoundaryConditionEnum::RiemannFarfield>(2);
  system.setThermodynamicModel<SimulationControl::kThermodynamicModel>(2.5_r, 25.0_r / 14.0_r);
  system.setArtificialViscosity(0.5_r);
  system.setTimeIntegration(0.001_r, {0, iterations});
  system.setViewConfig("build/out/" + kExampleName, kExampleName, -1);
  system.addViewVariable({SubrosaDG::ViewVariableEnum::Density, SubrosaDG::ViewVariableEnum::Velocity,
      SubrosaDG::ViewVariableEnum::Pressure, SubrosaDG::ViewVariableEnum::MachNumber,
      SubrosaDG::ViewVariableEnum::ArtificialViscosity});
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
