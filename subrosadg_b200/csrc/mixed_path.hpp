// mixed_path.hpp — dense-operator device path for meshes that contain triangles (hybrid triangle/quadrangle meshes of
// examples/karmanvortex_2d_cns.cpp and the naca0012 hybrid variant of BASELINE.json) — interface used by sdg_api.cu.
//
// The tensor path (tensor_kernels.cuh / line_kernels.cuh / ns_kernels.cuh) needs Nq == Nb and one element type; this path
// keeps the reference's own representation instead: modal coefficients [n][Nb][Nv] per element type
// (SolveControl.cpp:45-58), dense per-type operators (BasisFunction.cpp:136-230), dense per-element M^-1
// (Geometry.cpp:88-100), one slot per (element, local face, face point) for the face fluxes (SpatialDiscrete.cpp:738-744).
// Kernels: mixed_path.cu.  No CPU fallback: every compute entry needs the CUDA device of the owning context.
#pragma once
#include <cstdint>
#include <memory>

#include "dev_util.cuh"
#include "host_plan.hpp"
#include "mixed_tables.hpp"
#include "physics.cuh"

namespace sdg {

struct MixedBlock {
  int type = 0, n = 0, g = 1;
  MixedTable T;
  std::vector<double> X;                               // [n][nn][2]
  std::vector<double> xq, jw, mt, Minv, minEdge;       // Geometry.cpp:29-100
  // device
  DevBuf<double> dPhi_, dDPhi, dPhiF, dProj, dMt, dJw, dMinv, dMinEdge;
  DevBuf<double> U, Ulast, R, A, RM;                   // state, residual, face-flux slots, (R M^-1) scratch of the parity hook
  DevBuf<double> AGv, AGi, Gvol, Gtot, Gf;             // NS: gradient face slots, volume / total / per-face lifted gradient coefficients
  // artificial viscosity: node_tag_ of the corner nodes (0-based), inner_radius_, variable_artificial_viscosity_ [n][nbasic]
  std::vector<int> tags; std::vector<double> radius;
  DevBuf<int> dTags; DevBuf<double> dRadius, dAvElem, dAvE, dNodalQ, dNodalF;
};

class MixedSolver {
 public:
  MixedSolver(int p, const PhysParams& phys, int nStages, const double (*rkc)[3], cudaStream_t stream, bool hasDevice, int device);
  void addBlock(int type, int n, int nGhost, int g, const double* coords);
  ~MixedSolver() { if (stepGraph_) cudaGraphExecDestroy(stepGraph_); }
  void setFaces(const FaceInput& F) { F_ = F; }
  void finalize();
  bool hasType(int type) const { return type >= 0 && type < 7 && blk_[type] != nullptr; }
  void sizes(int type, int32_t* out) const;
  void quadratureCoordinates(int type, double* xq) const;
  void boundaryQuadratureCoordinates(double* xb);
  void setStateFromPrimitive(int type, const double* prim);
  void setBoundaryPrimitive(const double* prim);
  void setState(int type, const double* U);
  void getState(int type, double* U);
  void setStateDevice(int type, const void* U);
  void getStateDevice(int type, void* U);
  void stateAtQuadrature(int type, double* Uq);
  void gradientAtQuadrature(int type, double* Gq);
  void gradientState(int type, double* G);
  void boundaryGradientState(double* Gb);
  void refreshGradient();
  // ShockCapturingEnum::ArtificialViscosity (SpatialDiscrete.cpp:37-192)
  void setArtificialViscosity(double empiricalTolerance, double factor, int nodeNumber);
  void setElementNodes(int type, const int32_t* nodeTag, const double* innerRadius);
  void updateArtificialViscosity();
  void nodeArtificialViscosity(double* out);
  void elementArtificialViscosity(int type, double* out);
  void viewVariable(int type, int variable, double* out);   // ViewVariable::get at the quadrature points, VariableConvertor.cpp:754-872
  double computeDt(double cfl);
  void step(double dt, int nSteps, double* relErr, float* ms);
  void residual(int type, double* Rmodal, double* rhsq);
  int64_t launches = 0;
  // host-plan diagnostics (CPU tests): what = 100*type + {0 Phi[Nq][Nb], 1 gradPhi[Nq][2][Nb], 2 Phi_f[Naq][Nb], 3 projection[Nb][Nq],
  // 4 detJ w[n][Nq], 5 (J^T)^-1 detJ w[n][Nq][4], 6 M^-1[n][Nb][Nb], 7 minEdge[n]}; 90 face normals[nf][Nqf][2], 91 |J| w[nf][Nqf]
  const std::vector<double>& debugArray(int what) const {
    if (what == 90) return nrm_;
    if (what == 91) return fjw_;
    const int type = what / 100, id = what % 100;
    if (!hasType(type)) throw std::runtime_error("no element block of this type");
    const MixedBlock& B = *blk_[type];
    switch (id) {
      case 0: return B.T.Phi; case 1: return B.T.dPhi; case 2: return B.T.PhiF; case 3: return B.T.Proj;
      case 4: return B.jw; case 5: return B.mt; case 6: return B.Minv; case 7: return B.minEdge;
    }
    throw std::runtime_error("bad diagnostics id");
  }
  int totalElements() const { int s = 0; for (auto& b : blk_) if (b) s += b->n; return s; }
 
 struct Args;   // kernel parameter block (mixed_path.cu)

 private:
  void fill(Args& a);
  void faceGeometry();
  void evalResidual(Args& a, int mode, bool wantNorm);
  MixedBlock& block(int type);
  void needDevice() const;

  int p_, nStages_, device_;
  PhysParams phys_;
  double rkc_[3][3];
  cudaStream_t stream_;
  bool hasDevice_, finalized_ = false;
  // true after a step: Gvol / Gtot / Gf hold what the LAST RK STAGE computed (the gradient of that stage's input) -- what the reference writes
  // to its raw files (Solver::writeRawBinary after stepSolver); false after a state setter: the getters evaluate the current state first
  bool gradFromStep_ = false;
  std::unique_ptr<MixedBlock> blk_[7];
  FaceInput F_;
  std::vector<double> xf_, nrm_, fjw_;
  DevBuf<int> dLe, dLt, dLf, dRe, dRt, dRf, dBc;
  DevBuf<double> dNrm, dFjw, dDummy, normPartial, normOut, dtPartial;
  double avTol_ = 0.0, avFactor_ = 1.0; int avNodes_ = 0; DevBuf<double> avNode_;
  void avUpdate();
  cudaGraphExec_t stepGraph_ = nullptr; double graphDt_ = 0.0; bool graphWarm_ = false; int64_t launchesPerStep_ = 0;
};

}  // namespace sdg
