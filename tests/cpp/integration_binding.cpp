// Compiles integration/SolveControlB200.cpp — the reference-side binding INTEGRATION.md describes — against the reference's own
// headers, and instantiates every member for the mesh models / equation sets of BASELINE.json's five configs.
#include "SolveControlB200.cpp"

using namespace SubrosaDG;

template <DimensionEnum D, MeshModelEnum M, typename Variable, BoundaryTimeEnum T = BoundaryTimeEnum::Steady,
          InitialConditionEnum I = InitialConditionEnum::Function>
using Control = SimulationControl<SolveControl<D, PolynomialOrderEnum::P3, T, SourceTermEnum::None>,
                                  NumericalControl<M, ShockCapturingEnum::None, LimiterEnum::None, I, TimeIntegrationEnum::SSPRK3>, Variable>;
using EulerVariable = CompresibleEulerVariable<ThermodynamicModelEnum::Constant, EquationOfStateEnum::IdealGas, ConvectiveFluxEnum::HLLC>;
using NSVariable = CompresibleNSVariable<ThermodynamicModelEnum::Constant, EquationOfStateEnum::IdealGas, TransportModelEnum::Sutherland,
                                         ConvectiveFluxEnum::HLLC, ViscousFluxEnum::BR2>;

using Periodic2d = Control<DimensionEnum::D2, MeshModelEnum::Quadrangle, EulerVariable>;                                   // configs[0]
using Periodic3d = Control<DimensionEnum::D3, MeshModelEnum::Hexahedron, EulerVariable>;                                   // configs[1]
using Karman2d = Control<DimensionEnum::D2, MeshModelEnum::TriangleQuadrangle, NSVariable>;                                // configs[2]
using Naca2d = Control<DimensionEnum::D2, MeshModelEnum::Triangle, EulerVariable>;                                         // configs[3]
using Sphere3d = Control<DimensionEnum::D3, MeshModelEnum::Hexahedron, NSVariable, BoundaryTimeEnum::TimeVarying>;         // configs[4]
using Restart2d = Control<DimensionEnum::D2, MeshModelEnum::Quadrangle, NSVariable, BoundaryTimeEnum::Steady, InitialConditionEnum::LastStep>;
// the shock examples: sod_1d_ceuler (Line) and cylinder_2d_ceuler (TriangleQuadrangle) with ShockCapturingEnum::ArtificialViscosity
template <DimensionEnum D, MeshModelEnum M>
using ShockControl = SimulationControl<SolveControl<D, PolynomialOrderEnum::P3, BoundaryTimeEnum::Steady, SourceTermEnum::None>,
                                       NumericalControl<M, ShockCapturingEnum::ArtificialViscosity, LimiterEnum::None, InitialConditionEnum::Function,
                                                        TimeIntegrationEnum::SSPRK3>, EulerVariable>;
using Sod1d = ShockControl<DimensionEnum::D1, MeshModelEnum::Line>;
using Cylinder2d = ShockControl<DimensionEnum::D2, MeshModelEnum::TriangleQuadrangle>;

template struct SubrosaDG::SolverB200<Periodic2d>;
template struct SubrosaDG::SolverB200<Periodic3d>;
template struct SubrosaDG::SolverB200<Karman2d>;
template struct SubrosaDG::SolverB200<Naca2d>;
template struct SubrosaDG::SolverB200<Sphere3d>;
template struct SubrosaDG::SolverB200<Restart2d>;
template struct SubrosaDG::SolverB200<Sod1d>;
template struct SubrosaDG::SolverB200<Cylinder2d>;

// the call sequence of System<SC>::solve (SystemControl.cpp:159-195) against the replacement
template <typename SC>
void solveLikeSystem(Mesh<SC>& mesh, SourceTerm<SC>& source_term, PhysicalModel<SC>& physical_model, BoundaryCondition<SC>& boundary_condition,
                     InitialCondition<SC>& initial_condition, TimeIntegration<SC>& time_integration, SolverB200<SC>& solver) {
  solver.initializeSolver(mesh, physical_model, boundary_condition, initial_condition);
  if (time_integration.delta_time_ == 0.0_r) solver.calculateDeltaTime(mesh, physical_model, time_integration);
  solver.writeRawBinary(mesh, "raw/x_0.zst");
  for (int i = time_integration.iteration_start_ + 1; i <= time_integration.iteration_end_; i++) {
    solver.stepSolver(mesh, source_term, physical_model, boundary_condition, time_integration);
    time_integration.iteration_ = i;
    solver.write_raw_binary_future_.get();
    if (std::isnan(solver.relative_error_(0))) break;
  }
  solver.error_finout_.close();
}
template void solveLikeSystem<Sphere3d>(Mesh<Sphere3d>&, SourceTerm<Sphere3d>&, PhysicalModel<Sphere3d>&, BoundaryCondition<Sphere3d>&,
                                        InitialCondition<Sphere3d>&, TimeIntegration<Sphere3d>&, SolverB200<Sphere3d>&);
