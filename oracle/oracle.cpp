// oracle/oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement ("port") of SubrosaDG's DG residual evaluation + explicit SSP-RK time stepping, i.e. everything
// below Solver<SC>::stepSolver (src/Solver/TimeIntegration.cpp:326-350), in the reference's own shape: AoS-like
// per-element storage, dense per-element operators, dense per-element inverse mass matrix, eight sweeps per RK stage
// (four gradient sweeps that also run for Euler, TimeIntegration.cpp:339-348), face -> element scatter.  OpenMP
// `parallel for` replaces tbb::parallel_for over the same ranges.  It doubles as the timed CPU baseline of bench.py.
//
// PARITY, what is pinned and what is not.  The whole reference cannot be built here (icpx/SYCL, oneTBB, Eigen, Gmsh 4.13.1 SDK,
// magic_enum, dbg-macro, zstd are absent) and it ships no tests or golden vectors.  PINNED AGAINST THE REFERENCE'S OWN CODE: the
// pointwise physics (physics.hpp: variable conversions, the five Riemann fluxes, the six boundary conditions with their gradient states
// and modifyBoundaryVariable, primitive gradients, viscous fluxes, Sutherland, Boussinesq) — `make ref` compiles the reference's
// src/Solver/{VariableConvertor,ConvectiveFlux,ViscousFlux,BoundaryCondition,PhysicalModel,SourceTerm}.cpp where they lie (ref_physics.cpp
// + the stand-in headers of ref_shim/), tests/golden/reference_physics.json holds their outputs and tests/test_reference_physics.py checks
// orc_physics against every vector to 1e-13; and the integer tables embedded in the reference (tests/golden/reference_tables.json).
// ALSO PINNED AGAINST THE REFERENCE'S OWN CODE: the assembly around the physics — ref_sweeps.cpp runs the reference's initializeSolver,
// calculateDeltaTime and stepSolver (all sweeps, RK update, relative error) on hand-filled Mesh<SC> objects; tests/golden/
// reference_sweeps.json holds 8 control types and tests/test_reference_sweeps.py checks this restatement against them to 1e-12.
// STILL UNPINNED w.r.t. a reference binary: the Gmsh-provided inputs (quadrature, H1Legendre / Lagrange basis values, Jacobians, the
// "innerRadius" quality) — restated in tables.hpp / elementGeometry and pinned only by exactness properties, free-stream preservation,
// the exact travelling wave, SSP-RK order and the analytic viscous decay (tests/test_oracle_*.py).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
#include <omp.h>

#include <cstdint>
#include <cstring>
#include <memory>
#include <string>

#include "physics.hpp"
#include "tables.hpp"

namespace orc {

// C(m x n) = alpha * A(m x k) * B(k x n) + beta * C, column-major
static inline void gemm(int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
                        double* C, int ldc) {
  for (int j = 0; j < n; j++) {
    double* c = C + (size_t)j * ldc;
    if (beta == 0.0) for (int i = 0; i < m; i++) c[i] = 0.0;
    else if (beta != 1.0) for (int i = 0; i < m; i++) c[i] *= beta;
    for (int l = 0; l < k; l++) {
      const double b = alpha * B[l + (size_t)j * ldb];
      const double* a = A + (size_t)l * lda;
      for (int i = 0; i < m; i++) c[i] += a[i] * b;
    }
  }
}
// C(m x n) = alpha * A(m x k) * B(n x k)^T + beta * C
static inline void gemmNT(int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
                          double* C, int ldc) {
  for (int j = 0; j < n; j++) {
    double* c = C + (size_t)j * ldc;
    if (beta == 0.0) for (int i = 0; i < m; i++) c[i] = 0.0;
    else if (beta != 1.0) for (int i = 0; i < m; i++) c[i] *= beta;
    for (int l = 0; l < k; l++) {
      const double b = alpha * B[j + (size_t)l * ldb];
      const double* a = A + (size_t)l * lda;
      for (int i = 0; i < m; i++) c[i] += a[i] * b;
    }
  }
}

// C(m x n) = alpha * A(m x k) * B(k x n) + beta * C with extended-precision accumulation (accurate mode: keeps the
// ill-conditioned modal M^-1 / least-squares products at round-off of the RESULT, so that the checker is not the noisier side)
template <class TB>
static inline void gemmLD(int m, int n, int k, double alpha, const double* A, int lda, const TB* B, int ldb, double beta,
                          double* C, int ldc) {
  for (int j = 0; j < n; j++) for (int i = 0; i < m; i++) {
    long double s = 0.0L;
    for (int l = 0; l < k; l++) s += (long double)A[i + (size_t)l * lda] * (long double)B[l + (size_t)j * ldb];
    C[i + (size_t)j * ldc] = (double)((long double)alpha * s + (beta == 0.0 ? 0.0L : (long double)beta * (long double)C[i + (size_t)j * ldc]));
  }
}

// Per element type: ElementBasisFunction + ElementQuadrature (src/Mesh/BasisFunction.cpp:136-230, Quadrature.cpp:36-52)
struct ElemTable {
  int type = 0, D = 0, p = 0, g = 1, Nb = 0, Nq = 0, Nf = 0, Naq = 0, nn = 0;
  std::vector<int> off, nqf;
  Quadrature quad, fquad;
  std::vector<long double> LSinvL;  // extended-precision copy of LSinv (accurate mode)
  std::vector<double> Phi, dPhi, PhiF, LSinv;   // modal_value_, modal_gradient_value_ (row q*D+d), modal_adjacency_value_
  std::vector<double> GN, dGN;                  // geometry Lagrange basis at volume points: Nq x nn, (Nq*D) x nn
  std::vector<double> GNf, dGNf;                // at face points (parent coordinates): Naq x nn, (Naq*D) x nn
  std::vector<double> ftan;                     // per face: d(xi)/d(s_a), Nf x (D-1) x D
  std::vector<double> fxi;                      // face points in parent reference coordinates, Naq x 3
  int nbasic = 0;                               // kBasicNodeNumber: corner nodes (the first nodes of the gmsh order)
  std::vector<double> NodalQ, NodalF;           // nodal_value_ (Nq x nbasic), nodal_adjacency_value_ (Naq x nbasic): order-1 Lagrange basis, BasisFunction.cpp:149-208

  void build(int type_, int p_, int g_) {
    type = type_; p = p_; g = g_; D = elemDim(type);
    Nb = numNodes(type, p);  // getElementBasisFunctionNumber :243-266 (same counts as nodes for line/tri/quad/hex)
    Nf = numFaces(type);
    quad = makeQuadrature(type, 2 * p);           // getElementQuadratureOrder :275-278
    fquad = makeQuadrature(faceType(type), 2 * p + 1);  // getAdjacencyElementQuadratureOrder :280-283
    Nq = quad.n;
    off.assign(Nf + 1, 0); nqf.assign(Nf, fquad.n);
    for (int f = 0; f < Nf; f++) off[f + 1] = off[f] + nqf[f];  // getElementAccumulateAdjacencyQuadratureNumber :368-379
    Naq = off[Nf];
    ModalBasis mb(type, p);
    LagrangeBasis lb(type, g);
    LagrangeBasis l1(type, 1);
    nn = lb.n; nbasic = l1.n;
    NodalQ.assign((size_t)Nq * nbasic, 0); NodalF.assign((size_t)Naq * nbasic, 0);
    std::vector<double> val; std::vector<std::array<double, 3>> grad;
    Phi.assign((size_t)Nq * Nb, 0); dPhi.assign((size_t)Nq * D * Nb, 0);
    GN.assign((size_t)Nq * nn, 0); dGN.assign((size_t)Nq * D * nn, 0);
    for (int q = 0; q < Nq; q++) {
      const double* x = &quad.pts[3 * q];
      mb.eval(x[0], x[1], x[2], val, grad);
      for (int b = 0; b < Nb; b++) { Phi[(size_t)b * Nq + q] = val[b]; for (int d = 0; d < D; d++) dPhi[(size_t)b * Nq * D + q * D + d] = grad[b][d]; }
      l1.eval(x[0], x[1], x[2], val, grad);
      for (int b = 0; b < nbasic; b++) NodalQ[(size_t)b * Nq + q] = val[b];
      lb.eval(x[0], x[1], x[2], val, grad);
      for (int b = 0; b < nn; b++) { GN[(size_t)b * Nq + q] = val[b]; for (int d = 0; d < D; d++) dGN[(size_t)b * Nq * D + q * D + d] = grad[b][d]; }
    }
    // modal_least_squares_inverse_ = (Phi^T Phi)^-1, BasisFunction.cpp:217
    LSinv.assign((size_t)Nb * Nb, 0);
    for (int a = 0; a < Nb; a++) for (int b = 0; b < Nb; b++) { long double s = 0; for (int q = 0; q < Nq; q++) s += (long double)Phi[(size_t)a * Nq + q] * (long double)Phi[(size_t)b * Nq + q]; LSinv[(size_t)b * Nb + a] = (double)s; }
    invertInPlace(LSinv, Nb, &LSinvL);
    // face tables: parent basis at face points through the P1 map of the face corners, BasisFunction.cpp:76-111,149-197
    PhiF.assign((size_t)Naq * Nb, 0); GNf.assign((size_t)Naq * nn, 0); dGNf.assign((size_t)Naq * D * nn, 0);
    ftan.assign((size_t)Nf * std::max(D - 1, 1) * D, 0); fxi.assign((size_t)Naq * 3, 0);
    const int ft = faceType(type);
    LagrangeBasis fl(ft, 1);
    for (int f = 0; f < Nf; f++) {
      std::vector<int> fc = faceCorners(type, f);
      for (int j = 0; j < nqf[f]; j++) {
        const double* s = &fquad.pts[3 * j];
        std::vector<double> fv; std::vector<std::array<double, 3>> fg;
        fl.eval(s[0], s[1], s[2], fv, fg);
        double xi[3] = {0, 0, 0};
        for (size_t m = 0; m < fc.size(); m++) { auto c = cornerCoord(type, fc[m]); for (int k = 0; k < 3; k++) xi[k] += fv[m] * c[k]; }
        if (j == 0) for (int a = 0; a < D - 1; a++) for (int k = 0; k < D; k++) { double t = 0; for (size_t m = 0; m < fc.size(); m++) t += fg[m][a] * cornerCoord(type, fc[m])[k]; ftan[((size_t)f * std::max(D - 1, 1) + a) * D + k] = t; }
        const int row = off[f] + j;
        for (int k = 0; k < 3; k++) fxi[(size_t)row * 3 + k] = xi[k];
        mb.eval(xi[0], xi[1], xi[2], val, grad);
        for (int b = 0; b < Nb; b++) PhiF[(size_t)b * Naq + row] = val[b];
        l1.eval(xi[0], xi[1], xi[2], val, grad);
        for (int b = 0; b < nbasic; b++) NodalF[(size_t)b * Naq + row] = val[b];
        lb.eval(xi[0], xi[1], xi[2], val, grad);
        for (int b = 0; b < nn; b++) { GNf[(size_t)b * Naq + row] = val[b]; for (int d = 0; d < D; d++) dGNf[(size_t)b * Naq * D + row * D + d] = grad[b][d]; }
      }
    }
  }
};

// ElementMesh + ElementSolver for one element type (src/Mesh/ReadControl.cpp:60-155, src/Solver/SolveControl.cpp:45-221)
struct ElemBlock {
  int type = 0, n = 0;
  ElemTable tab;
  std::vector<double> X;                         // node_coordinate_: n x nn x D (gmsh node order)
  std::vector<long double> MinvL;  // extended-precision copy of Minv (accurate mode only)
  std::vector<double> xq, jw, mt, Minv, minEdge; // quadrature_node_coordinate_, detJ*w, (J^T)^-1 detJ w, M^-1, minimum_edge_
  // solver state (column-major per element, rows = variables)
  std::vector<double> coef, coefLast, vq, vaq, res, sq;
  std::vector<double> gvq, gvaq, gresVol, gcoefVol;   // PerElementVolumeGradientSolver
  std::vector<double> gcoef, giaq, gires, gicoef;     // NS: total gradient, interface quadrature, per-face (BR2) / single (BR1) lifts
  // ShockCapturingEnum::ArtificialViscosity: mesh data (node_tag_ of the corner nodes, 0-based; inner_radius_) and
  // variable_artificial_viscosity_ (n x nbasic), SolveControl.cpp:132
  std::vector<int> nodeTag;
  std::vector<double> innerRadius, avElem;
};

// AdjacencyElementMesh + AdjacencyElementSolver (ReadControl.cpp:72-83, SolveControl.cpp:223-289)
struct FaceSet {
  int ftype = 0, nInt = 0, nBnd = 0, Nqf = 0;
  std::vector<int> elem[2], etype[2], lface[2], rot, bc, phys;
  std::vector<double> xf, nrm, jw;   // quadrature_node_coordinate_, normal_vector_, |J| w   (per face x Nqf)
  std::vector<Var> dummy;            // boundary_dummy_variable_: nBnd x Nqf
  std::vector<double> dummyPrim;     // as uploaded
};

struct Oracle {
  Phys P;
  int p = 1, rk = kSSPRK3;
  bool deadGradient = true;  // run G1-G4 for Euler like the reference does
  bool accurate = false;     // extended-precision accumulation in the M^-1 / least-squares products (checker mode)
  std::unique_ptr<ElemBlock> blk[7];
  FaceSet F;
  bool finalized = false;
  double relErr[kMaxV] = {0, 0, 0, 0, 0};
  bool av = false;           // ShockCapturingEnum::ArtificialViscosity
  double avTol = 0.0, avFactor = 1.0;   // Solver::empirical_tolerance_, artificial_viscosity_factor_ (SolveControl.cpp:293-294)
  std::vector<double> nodeAV;           // Solver::node_artificial_viscosity_
  int totalElems() const { int s = 0; for (auto& b : blk) if (b) s += b->n; return s; }
};

static thread_local std::string g_err;

// ---- geometry (src/Mesh/Geometry.cpp) -------------------------------------------------------------------------------
static double det_inv(int D, const double* Jt /*row-major k,l*/, double* inv /*row-major*/) {
  if (D == 1) { inv[0] = 1.0 / Jt[0]; return Jt[0]; }
  if (D == 2) {
    const double det = Jt[0] * Jt[3] - Jt[1] * Jt[2];
    inv[0] = Jt[3] / det; inv[1] = -Jt[1] / det; inv[2] = -Jt[2] / det; inv[3] = Jt[0] / det;
    return det;
  }
  const double a = Jt[0], b = Jt[1], c = Jt[2], d = Jt[3], e = Jt[4], f = Jt[5], g = Jt[6], h = Jt[7], i = Jt[8];
  const double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
  const double det = a * A + b * B + c * C;
  inv[0] = A / det; inv[1] = -(b * i - c * h) / det; inv[2] = (b * f - c * e) / det;
  inv[3] = B / det; inv[4] = (a * i - c * g) / det; inv[5] = -(a * f - c * d) / det;
  inv[6] = C / det; inv[7] = -(a * h - b * g) / det; inv[8] = (a * e - b * d) / det;
  return det;
}

static void elementGeometry(ElemBlock& B, bool accurate) {
  const ElemTable& T = B.tab;
  const int D = T.D, Nq = T.Nq, nn = T.nn, Nb = T.Nb;
  B.xq.assign((size_t)B.n * Nq * D, 0); B.jw.assign((size_t)B.n * Nq, 0); B.mt.assign((size_t)B.n * Nq * D * D, 0);
  B.Minv.assign((size_t)B.n * Nb * Nb, 0); B.minEdge.assign(B.n, 0);
  if (accurate) B.MinvL.assign((size_t)B.n * Nb * Nb, 0.0L);
  bool bad = false;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < B.n; e++) {
    const double* X = &B.X[(size_t)e * nn * D];
    for (int q = 0; q < Nq; q++) {
      // getElementJacobian, Geometry.cpp:44-67: jacobian_transpose(k,l) = d x_l / d xi_k
      double Jt[9] = {0}, inv[9];
      for (int m = 0; m < nn; m++) {
        const double N = T.GN[(size_t)m * Nq + q];
        for (int l = 0; l < D; l++) B.xq[((size_t)e * Nq + q) * D + l] += N * X[m * D + l];
        for (int k = 0; k < D; k++) { const double dN = T.dGN[(size_t)m * Nq * D + q * D + k]; for (int l = 0; l < D; l++) Jt[k * D + l] += dN * X[m * D + l]; }
      }
      const double det = det_inv(D, Jt, inv);
      if (!(det > 0.0)) bad = true;
      const double w = det * T.quad.wts[q];
      B.jw[(size_t)e * Nq + q] = w;
      // column = Jt.inverse().reshaped() (column-major) * detJ*w : index c + D*d'  <-  inv(c,d')
      double* mt = &B.mt[((size_t)e * Nq + q) * D * D];
      for (int c = 0; c < D; c++) for (int dd = 0; dd < D; dd++) mt[dd * D + c] = inv[c * D + dd] * w;
    }
    // calculateElementLocalMassMatrixInverse, Geometry.cpp:88-100: M = Phi^T diag(detJ w) Phi
    std::vector<double> M((size_t)Nb * Nb);
    for (int a = 0; a < Nb; a++) for (int b = a; b < Nb; b++) {
      long double s = 0; for (int q = 0; q < Nq; q++) s += (long double)T.Phi[(size_t)a * Nq + q] * (long double)B.jw[(size_t)e * Nq + q] * (long double)T.Phi[(size_t)b * Nq + q];
      M[(size_t)b * Nb + a] = (double)s; M[(size_t)a * Nb + b] = (double)s;
    }
    std::vector<long double> ML;
    invertInPlace(M, Nb, B.MinvL.empty() ? nullptr : &ML);
    std::memcpy(&B.Minv[(size_t)e * Nb * Nb], M.data(), sizeof(double) * Nb * Nb);
    if (!B.MinvL.empty()) std::copy(ML.begin(), ML.end(), B.MinvL.begin() + (size_t)e * Nb * Nb);
    // getElementQuality "minEdge", Geometry.cpp:29-42: shortest straight distance between the end vertices of an edge
    double me = 1e300;
    auto dist = [&](int a, int b) { double s = 0; for (int l = 0; l < D; l++) { double d = X[a * D + l] - X[b * D + l]; s += d * d; } return std::sqrt(s); };
    if (T.type == kLine) me = dist(0, 1);
    else if (T.type == kTriangle) for (int k = 0; k < 3; k++) me = std::min(me, dist(k, (k + 1) % 3));
    else if (T.type == kQuadrangle) for (int k = 0; k < 4; k++) me = std::min(me, dist(k, (k + 1) % 4));
    else if (T.type == kHexahedron) { static const int E[12][2] = {{0, 1}, {0, 3}, {0, 4}, {1, 2}, {1, 5}, {2, 3}, {2, 6}, {3, 7}, {4, 5}, {4, 7}, {5, 6}, {6, 7}}; for (auto& ed : E) me = std::min(me, dist(ed[0], ed[1])); }
    B.minEdge[e] = me;
  }
  if (bad) throw std::runtime_error("oracle: non-positive Jacobian determinant");
}

// getAdjacencyElementJacobian + calculateAdjacencyElementNormalVector, Geometry.cpp:69-86,102-169.  The face element is
// the restriction of the LEFT parent's mapping (its nodes are the parent's face nodes), so tangents come from the parent.
static void faceGeometry(Oracle& O) {
  FaceSet& F = O.F;
  const int D = O.P.D, nf = F.nInt + F.nBnd, Nqf = F.Nqf;
  F.xf.assign((size_t)nf * Nqf * D, 0); F.nrm.assign((size_t)nf * Nqf * D, 0); F.jw.assign((size_t)nf * Nqf, 0);
#pragma omp parallel for schedule(static)
  for (int i = 0; i < nf; i++) {
    const ElemBlock& B = *O.blk[F.etype[0][i]];
    const ElemTable& T = B.tab;
    const int e = F.elem[0][i], f = F.lface[0][i], nn = T.nn, Naq = T.Naq;
    const double* X = &B.X[(size_t)e * nn * D];
    for (int j = 0; j < Nqf; j++) {
      const int row = T.off[f] + j;
      double Jt[9] = {0};
      double* x = &F.xf[((size_t)i * Nqf + j) * D];
      for (int m = 0; m < nn; m++) {
        const double N = T.GNf[(size_t)m * Naq + row];
        for (int l = 0; l < D; l++) x[l] += N * X[m * D + l];
        for (int k = 0; k < D; k++) { const double dN = T.dGNf[(size_t)m * Naq * D + row * D + k]; for (int l = 0; l < D; l++) Jt[k * D + l] += dN * X[m * D + l]; }
      }
      double* nv = &F.nrm[((size_t)i * Nqf + j) * D];
      double scale = 1.0;
      if (D == 1) {  // :102-112
        nv[0] = f == 0 ? -1.0 : 1.0;
      } else if (D == 2) {  // :114-129  normal = (t_y, -t_x)/|t|
        const double* ts = &T.ftan[((size_t)f * 1 + 0) * D];
        double t[2] = {0, 0};
        for (int k = 0; k < 2; k++) for (int l = 0; l < 2; l++) t[l] += ts[k] * Jt[k * 2 + l];
        scale = std::sqrt(t[0] * t[0] + t[1] * t[1]);
        nv[0] = t[1] / scale; nv[1] = -t[0] / scale;
      } else {  // :131-148  normal = (d_s x) x (d_t x) normalised
        const double* ts = &T.ftan[((size_t)f * 2 + 0) * D];
        const double* tt = &T.ftan[((size_t)f * 2 + 1) * D];
        double a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
        for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) { a[l] += ts[k] * Jt[k * 3 + l]; b[l] += tt[k] * Jt[k * 3 + l]; }
        double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
        scale = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        for (int l = 0; l < 3; l++) nv[l] = c[l] / scale;
      }
      F.jw[(size_t)i * Nqf + j] = scale * T.fquad.wts[j];
    }
  }
}

// ---- solver helpers -------------------------------------------------------------------------------------------------
struct Sizes { int D, Nv, G, Nb, Nq, Naq, Nf; };
static inline Sizes sizes(const Oracle& O, const ElemBlock& B) { return {O.P.D, O.P.Nv, O.P.Nv * O.P.D, B.tab.Nb, B.tab.Nq, B.tab.Naq, B.tab.Nf}; }
static inline int nLift(const Oracle& O, const ElemBlock& B) { return O.P.visc == kBR2 ? B.tab.Nf : 1; }

static void allocSolver(Oracle& O, ElemBlock& B) {
  const Sizes s = sizes(O, B);
  const size_t n = B.n;
  B.coef.assign(n * s.Nv * s.Nb, 0); B.coefLast = B.coef;
  B.vq.assign(n * s.Nv * s.Nq * s.D, 0); B.vaq.assign(n * s.Nv * s.Naq, 0); B.res.assign(n * s.Nv * s.Nb, 0);
  B.gvq.assign(n * s.G * s.Nq * s.D, 0); B.gvaq.assign(n * s.G * s.Naq, 0); B.gresVol.assign(n * s.G * s.Nb, 0); B.gcoefVol.assign(n * s.G * s.Nb, 0);
  if (O.P.ns()) {
    B.gcoef.assign(n * s.G * s.Nb, 0); B.giaq.assign(n * s.G * s.Naq, 0);
    B.gires.assign(n * nLift(O, B) * s.G * s.Nb, 0); B.gicoef = B.gires;
  }
  if (O.P.source != kSourceNone) B.sq.assign(n * s.Nv * s.Nq, 0);
}

// AdjacencyElementVariable::get, VariableConvertor.cpp:432-485: cons (Nv x Nqf) = U * Phi_f[face]^T, then comp.
static inline void faceTrace(const Oracle& O, int type, int e, int f, Var* out) {
  const ElemBlock& B = *O.blk[type]; const ElemTable& T = B.tab;
  const int Nv = O.P.Nv, Nb = T.Nb, Naq = T.Naq, nq = T.nqf[f];
  const double* U = &B.coef[(size_t)e * Nv * Nb];
  for (int j = 0; j < nq; j++) {
    const int row = T.off[f] + j;
    for (int v = 0; v < Nv; v++) out[j].cons[v] = 0.0;
    for (int b = 0; b < Nb; b++) { const double ph = T.PhiF[(size_t)b * Naq + row]; for (int v = 0; v < Nv; v++) out[j].cons[v] += U[b * Nv + v] * ph; }
  }
}
// AdjacencyElementVariableGradient::get<kViscousFlux>, VariableConvertor.cpp:640-723
static inline void faceGradTrace(const Oracle& O, int type, int e, int f, double* out /* G x nq */) {
  const ElemBlock& B = *O.blk[type]; const ElemTable& T = B.tab;
  const int G = O.P.Nv * O.P.D, Nb = T.Nb, Naq = T.Naq, nq = T.nqf[f];
  std::vector<double> tmp;
  const double* C;
  if (O.P.visc == kBR1) C = &B.gcoef[(size_t)e * G * Nb];
  else if (O.P.visc == kBR2) {
    tmp.resize((size_t)G * Nb);
    const double* a = &B.gcoefVol[(size_t)e * G * Nb];
    const double* b = &B.gicoef[((size_t)e * T.Nf + f) * G * Nb];
    for (int k = 0; k < G * Nb; k++) tmp[k] = a[k] + b[k];
    C = tmp.data();
  } else C = &B.gcoefVol[(size_t)e * G * Nb];
  for (int j = 0; j < nq; j++) {
    const int row = T.off[f] + j;
    for (int r = 0; r < G; r++) out[(size_t)j * G + r] = 0.0;
    for (int b = 0; b < Nb; b++) { const double ph = T.PhiF[(size_t)b * Naq + row]; for (int r = 0; r < G; r++) out[(size_t)j * G + r] += C[b * G + r] * ph; }
  }
}

// ---- the eight sweeps of one RK stage (TimeIntegration.cpp:339-348) --------------------------------------------------
// G1 calculateElementGardientQuadrature, SpatialDiscrete.cpp:294-322
static void sweepG1(Oracle& O, ElemBlock& B) {
  const Sizes s = sizes(O, B); const ElemTable& T = B.tab;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < B.n; e++) {
    std::vector<double> uq((size_t)s.Nv * s.Nq);
    gemmNT(s.Nv, s.Nq, s.Nb, 1.0, &B.coef[(size_t)e * s.Nv * s.Nb], s.Nv, T.Phi.data(), s.Nq, 0.0, uq.data(), s.Nv);
    double* out = &B.gvq[(size_t)e * s.G * s.Nq * s.D];
    for (int q = 0; q < s.Nq; q++) {
      const double* mt = &B.mt[((size_t)e * s.Nq + q) * s.D * s.D];
      for (int v = 0; v < s.Nv; v++) for (int c = 0; c < s.D; c++) for (int dd = 0; dd < s.D; dd++)
        out[(size_t)(q * s.D + dd) * s.G + v * s.D + c] = uq[(size_t)q * s.Nv + v] * mt[dd * s.D + c];
    }
  }
}
// G2 calculateInterior/BoundaryAdjacencyElementGardientQuadrature, SpatialDiscrete.cpp:844-968
static void sweepG2(Oracle& O) {
  FaceSet& F = O.F; const Phys& P = O.P;
  const int D = P.D, Nv = P.Nv, G = Nv * D, Nqf = F.Nqf;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < F.nInt; i++) {
    std::vector<Var> L(Nqf), R(Nqf);
    const int tL = F.etype[0][i], tR = F.etype[1][i], eL = F.elem[0][i], eR = F.elem[1][i], fL = F.lface[0][i], fR = F.lface[1][i];
    faceTrace(O, tL, eL, fL, L.data()); faceTrace(O, tR, eR, fR, R.data());
    const std::vector<int> seq = faceQuadratureSequence(F.ftype, O.p, F.rot[i]);
    ElemBlock& BL = *O.blk[tL]; ElemBlock& BR = *O.blk[tR];
    const int offL = BL.tab.off[fL], offR = BR.tab.off[fR];
    for (int j = 0; j < Nqf; j++) {
      const double* n = &F.nrm[((size_t)i * Nqf + j) * D]; const double w = F.jw[(size_t)i * Nqf + j];
      const int jr = seq[j];
      double* aL = &BL.gvaq[((size_t)eL * BL.tab.Naq + offL + j) * G];
      double* aR = &BR.gvaq[((size_t)eR * BR.tab.Naq + offR + jr) * G];
      for (int v = 0; v < Nv; v++) for (int c = 0; c < D; c++) {  // calculateVolumeGardientFlux, ViscousFlux.cpp:33-43
        const double t = n[c] * (L[j].cons[v] + R[jr].cons[v]) / 2.0 * w;
        aL[v * D + c] = t; aR[v * D + c] = -t;
      }
      if (P.ns()) {  // calculateInterfaceGardientFlux, ViscousFlux.cpp:46-56 — same sign on both sides (:899-906)
        double* bL = &BL.giaq[((size_t)eL * BL.tab.Naq + offL + j) * G];
        double* bR = &BR.giaq[((size_t)eR * BR.tab.Naq + offR + jr) * G];
        for (int v = 0; v < Nv; v++) for (int c = 0; c < D; c++) {
          const double t = n[c] * (R[jr].cons[v] - L[j].cons[v]) / 2.0 * w;
          bL[v * D + c] = t; bR[v * D + c] = t;
        }
      }
    }
  }
#pragma omp parallel for schedule(static)
  for (int i = F.nInt; i < F.nInt + F.nBnd; i++) {
    std::vector<Var> L(Nqf);
    const int tL = F.etype[0][i], eL = F.elem[0][i], fL = F.lface[0][i];
    faceTrace(O, tL, eL, fL, L.data());
    for (int j = 0; j < Nqf; j++) compFromCons(P, L[j]);
    ElemBlock& BL = *O.blk[tL];
    const int offL = BL.tab.off[fL];
    for (int j = 0; j < Nqf; j++) {
      const double* n = &F.nrm[((size_t)i * Nqf + j) * D]; const double w = F.jw[(size_t)i * Nqf + j];
      double vol[kMaxV], itf[kMaxV];
      bcBoundaryGradientVariable(P, F.bc[i], n, L[j], F.dummy[(size_t)(i - F.nInt) * Nqf + j], vol, itf);
      double* aL = &BL.gvaq[((size_t)eL * BL.tab.Naq + offL + j) * G];
      for (int v = 0; v < Nv; v++) for (int c = 0; c < D; c++) aL[v * D + c] = n[c] * vol[v] * w;  // calculateGardientRawFlux :26-30
      if (P.ns()) {
        double* bL = &BL.giaq[((size_t)eL * BL.tab.Naq + offL + j) * G];
        for (int v = 0; v < Nv; v++) for (int c = 0; c < D; c++) bL[v * D + c] = n[c] * itf[v] * w;
      }
    }
  }
}
// G3 calculateElementGardientResidual, SpatialDiscrete.cpp:1034-1068
static void sweepG3(Oracle& O, ElemBlock& B) {
  const Sizes s = sizes(O, B); const ElemTable& T = B.tab;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < B.n; e++) {
    double* r = &B.gresVol[(size_t)e * s.G * s.Nb];
    gemm(s.G, s.Nb, s.Naq, 1.0, &B.gvaq[(size_t)e * s.G * s.Naq], s.G, T.PhiF.data(), s.Naq, 0.0, r, s.G);
    gemm(s.G, s.Nb, s.Nq * s.D, -1.0, &B.gvq[(size_t)e * s.G * s.Nq * s.D], s.G, T.dPhi.data(), s.Nq * s.D, 1.0, r, s.G);
    if (O.P.ns()) {
      if (O.P.visc == kBR1) {
        gemm(s.G, s.Nb, s.Naq, 1.0, &B.giaq[(size_t)e * s.G * s.Naq], s.G, T.PhiF.data(), s.Naq, 0.0, &B.gires[(size_t)e * s.G * s.Nb], s.G);
      } else if (O.P.visc == kBR2) {
        for (int f = 0; f < s.Nf; f++)
          gemm(s.G, s.Nb, T.nqf[f], 1.0, &B.giaq[((size_t)e * s.Naq + T.off[f]) * s.G], s.G, T.PhiF.data() + T.off[f], s.Naq, 0.0,
               &B.gires[((size_t)e * s.Nf + f) * s.G * s.Nb], s.G);
      }
    }
  }
}
// G4 updateElementGardientBasisFunctionCoefficient, TimeIntegration.cpp:200-228
static void sweepG4(Oracle& O, ElemBlock& B) {
  const Sizes s = sizes(O, B);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < B.n; e++) {
    const double* Mi = &B.Minv[(size_t)e * s.Nb * s.Nb];
    double* gv = &B.gcoefVol[(size_t)e * s.G * s.Nb];
    gemm(s.G, s.Nb, s.Nb, 1.0, &B.gresVol[(size_t)e * s.G * s.Nb], s.G, Mi, s.Nb, 0.0, gv, s.G);
    if (O.P.ns()) {
      double* gt = &B.gcoef[(size_t)e * s.G * s.Nb];
      std::memcpy(gt, gv, sizeof(double) * s.G * s.Nb);
      const int nl = nLift(O, B);
      for (int f = 0; f < nl; f++) {
        double* gi = &B.gicoef[((size_t)e * nl + f) * s.G * s.Nb];
        gemm(s.G, s.Nb, s.Nb, 1.0, &B.gires[((size_t)e * nl + f) * s.G * s.Nb], s.G, Mi, s.Nb, 0.0, gi, s.G);
        for (int k = 0; k < s.G * s.Nb; k++) gt[k] += gi[k];
      }
    }
  }
}
// R1 calculateElementQuadrature, SpatialDiscrete.cpp:194-266
static void sweepR1(Oracle& O, ElemBlock& B) {
  const Sizes s = sizes(O, B); const ElemTable& T = B.tab; const Phys& P = O.P;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < B.n; e++) {
    std::vector<double> uq((size_t)s.Nv * s.Nq), gq;
    gemmNT(s.Nv, s.Nq, s.Nb, 1.0, &B.coef[(size_t)e * s.Nv * s.Nb], s.Nv, T.Phi.data(), s.Nq, 0.0, uq.data(), s.Nv);
    if (P.ns()) {  // ElementVariableGradient::get<kViscousFlux>: total gradient coefficients, VariableConvertor.cpp:623-638
      gq.resize((size_t)s.G * s.Nq);
      gemmNT(s.G, s.Nq, s.Nb, 1.0, &B.gcoef[(size_t)e * s.G * s.Nb], s.G, T.Phi.data(), s.Nq, 0.0, gq.data(), s.G);
    }
    std::vector<double> gvolq;
    if (O.av) {  // ElementVariableGradient::get<ViscousFluxEnum::None>: the volume gradient coefficients
      gvolq.resize((size_t)s.G * s.Nq);
      gemmNT(s.G, s.Nq, s.Nb, 1.0, &B.gcoefVol[(size_t)e * s.G * s.Nb], s.G, T.Phi.data(), s.Nq, 0.0, gvolq.data(), s.G);
    }
    double* out = &B.vq[(size_t)e * s.Nv * s.Nq * s.D];
    for (int q = 0; q < s.Nq; q++) {
      Var v;
      for (int k = 0; k < s.Nv; k++) v.cons[k] = uq[(size_t)q * s.Nv + k];
      compFromCons(P, v);
      double Fc[kMaxD * kMaxV], Fv[kMaxD * kMaxV];
      convRawFlux(P, v.comp, Fc);
      if (P.ns()) {
        double gp[kMaxD * kMaxV];
        primGradFromConsGrad(P, v, &gq[(size_t)q * s.G], gp);
        viscRawFlux(P, v.comp, gp, Fv);
        for (int k = 0; k < s.D * s.Nv; k++) Fc[k] -= Fv[k];
      }
      if (O.av) {  // calculateArtificialViscousRawFlux, ViscousFlux.cpp:105-113: eps(q) * volume gradient of the conserved variables (SpatialDiscrete.cpp:210-228,249-253)
        double eps = 0.0;
        for (int k = 0; k < T.nbasic; k++) eps += T.NodalQ[(size_t)k * s.Nq + q] * B.avElem[(size_t)e * T.nbasic + k];
        for (int k = 0; k < s.D * s.Nv; k++) Fc[k] -= eps * gvolq[(size_t)q * s.G + k];
      }
      const double* mt = &B.mt[((size_t)e * s.Nq + q) * s.D * s.D];
      // flux^T (Nv x D) * Mt (D x D)
      for (int dd = 0; dd < s.D; dd++) for (int k = 0; k < s.Nv; k++) {
        double t = 0; for (int c = 0; c < s.D; c++) t += Fc[k * s.D + c] * mt[dd * s.D + c];
        out[(size_t)(q * s.D + dd) * s.Nv + k] = t;
      }
      if (P.source != kSourceNone) {
        double S[kMaxV]; sourceTerm(P, v.comp, S);
        for (int k = 0; k < s.Nv; k++) B.sq[((size_t)e * s.Nq + q) * s.Nv + k] = S[k] * B.jw[(size_t)e * s.Nq + q];
      }
    }
  }
}
// AdjacencyElementVariableGradient::get<ViscousFluxEnum::None> (volume gradient trace) and calculateAdjacencyElementArtificialViscosity
// (SpatialDiscrete.cpp:529-631): eps at the face points of local face f = nodal_adjacency_value_ rows of that face * corner values
static inline void faceVolGradTrace(const Oracle& O, int type, int e, int f, double* out /* G x nq */, double* eps /* nq */) {
  const ElemBlock& B = *O.blk[type]; const ElemTable& T = B.tab;
  const int G = O.P.Nv * O.P.D, Nb = T.Nb, Naq = T.Naq, nq = T.nqf[f];
  const double* C = &B.gcoefVol[(size_t)e * G * Nb];
  for (int j = 0; j < nq; j++) {
    const int row = T.off[f] + j;
    for (int r = 0; r < G; r++) out[(size_t)j * G + r] = 0.0;
    for (int b = 0; b < Nb; b++) { const double ph = T.PhiF[(size_t)b * Naq + row]; for (int r = 0; r < G; r++) out[(size_t)j * G + r] += C[b * G + r] * ph; }
    eps[j] = 0.0;
    for (int k = 0; k < T.nbasic; k++) eps[j] += T.NodalF[(size_t)k * Naq + row] * B.avElem[(size_t)e * T.nbasic + k];
  }
}
// R2 calculateInterior/BoundaryAdjacencyElementQuadrature, SpatialDiscrete.cpp:633-842
static void sweepR2(Oracle& O) {
  FaceSet& F = O.F; const Phys& P = O.P;
  const int D = P.D, Nv = P.Nv, G = Nv * D, Nqf = F.Nqf;
#pragma omp parallel for schedule(static)
  for (int i = 0; i < F.nInt; i++) {
    std::vector<Var> L(Nqf), R(Nqf);
    std::vector<double> gL, gR, pL, pR;
    const int tL = F.etype[0][i], tR = F.etype[1][i], eL = F.elem[0][i], eR = F.elem[1][i], fL = F.lface[0][i], fR = F.lface[1][i];
    faceTrace(O, tL, eL, fL, L.data()); faceTrace(O, tR, eR, fR, R.data());
    for (int j = 0; j < Nqf; j++) { compFromCons(P, L[j]); compFromCons(P, R[j]); }
    if (P.ns()) {
      gL.resize((size_t)G * Nqf); gR.resize((size_t)G * Nqf); pL.resize((size_t)G * Nqf); pR.resize((size_t)G * Nqf);
      faceGradTrace(O, tL, eL, fL, gL.data()); faceGradTrace(O, tR, eR, fR, gR.data());
      for (int j = 0; j < Nqf; j++) { primGradFromConsGrad(P, L[j], &gL[(size_t)j * G], &pL[(size_t)j * G]); primGradFromConsGrad(P, R[j], &gR[(size_t)j * G], &pR[(size_t)j * G]); }
    }
    std::vector<double> avgL, avgR, epsL, epsR;
    if (O.av) {
      avgL.resize((size_t)G * Nqf); avgR.resize((size_t)G * Nqf); epsL.resize(Nqf); epsR.resize(Nqf);
      faceVolGradTrace(O, tL, eL, fL, avgL.data(), epsL.data()); faceVolGradTrace(O, tR, eR, fR, avgR.data(), epsR.data());
    }
    const std::vector<int> seq = faceQuadratureSequence(F.ftype, O.p, F.rot[i]);
    ElemBlock& BL = *O.blk[tL]; ElemBlock& BR = *O.blk[tR];
    const int offL = BL.tab.off[fL], offR = BR.tab.off[fR];
    for (int j = 0; j < Nqf; j++) {
      const double* n = &F.nrm[((size_t)i * Nqf + j) * D]; const double w = F.jw[(size_t)i * Nqf + j];
      const int jr = seq[j];
      double Fc[kMaxV];
      convFlux(P, n, L[j], R[jr], Fc);
      if (P.ns()) {  // calculateViscousFlux, ViscousFlux.cpp:139-153
        double a[kMaxV], b[kMaxV];
        viscNormalFlux(P, n, L[j].comp, &pL[(size_t)j * G], a);
        viscNormalFlux(P, n, R[jr].comp, &pR[(size_t)jr * G], b);
        for (int v = 0; v < Nv; v++) Fc[v] -= (a[v] + b[v]) / 2.0;
      }
      if (O.av) {  // calculateArtificialViscousFlux, ViscousFlux.cpp:172-186: average of eps * grad(U) . n of both sides.  The reference
        // passes right_quadrature_node_artificial_viscosity(j) (SpatialDiscrete.cpp:714-719): the right VISCOSITY is taken at the right
        // element's own point j, not at the matching point sequence[j] its gradient column uses -- restated as it is
        for (int v = 0; v < Nv; v++) {
          double a = 0.0, b = 0.0;
          for (int c = 0; c < D; c++) { a += epsL[j] * avgL[(size_t)j * G + v * D + c] * n[c]; b += epsR[j] * avgR[(size_t)jr * G + v * D + c] * n[c]; }
          Fc[v] -= (a + b) / 2.0;
        }
      }
      double* aL = &BL.vaq[((size_t)eL * BL.tab.Naq + offL + j) * Nv];
      double* aR = &BR.vaq[((size_t)eR * BR.tab.Naq + offR + jr) * Nv];
      for (int v = 0; v < Nv; v++) { aL[v] = Fc[v] * w; aR[v] = -Fc[v] * w; }
    }
  }
#pragma omp parallel for schedule(static)
  for (int i = F.nInt; i < F.nInt + F.nBnd; i++) {
    std::vector<Var> L(Nqf);
    std::vector<double> gL, pL;
    const int tL = F.etype[0][i], eL = F.elem[0][i], fL = F.lface[0][i];
    faceTrace(O, tL, eL, fL, L.data());
    for (int j = 0; j < Nqf; j++) compFromCons(P, L[j]);
    if (P.ns()) {
      gL.resize((size_t)G * Nqf); pL.resize((size_t)G * Nqf);
      faceGradTrace(O, tL, eL, fL, gL.data());
      for (int j = 0; j < Nqf; j++) primGradFromConsGrad(P, L[j], &gL[(size_t)j * G], &pL[(size_t)j * G]);
    }
    std::vector<double> avgL, epsL;
    if (O.av) { avgL.resize((size_t)G * Nqf); epsL.resize(Nqf); faceVolGradTrace(O, tL, eL, fL, avgL.data(), epsL.data()); }
    ElemBlock& BL = *O.blk[tL];
    const int offL = BL.tab.off[fL];
    for (int j = 0; j < Nqf; j++) {
      const double* n = &F.nrm[((size_t)i * Nqf + j) * D]; const double w = F.jw[(size_t)i * Nqf + j];
      Var b;
      bcBoundaryVariable(P, F.bc[i], n, L[j], F.dummy[(size_t)(i - F.nInt) * Nqf + j], b.comp);
      double Fc[kMaxV];
      convNormalFlux(P, n, b.comp, Fc);  // SpatialDiscrete.cpp:802-803: flux of the boundary state, no Riemann solve
      if (P.ns()) {
        // modifyBoundaryVariable (BoundaryCondition.cpp:299-307,443-452,490-501,535-546): walls overwrite the interior
        // computational column; boundary gradient = interior primitive gradient (adiabatic: zero temperature gradient)
        double gb[kMaxD * kMaxV];
        if (bcIsWall(F.bc[i])) for (int k = 0; k < D + 3; k++) L[j].comp[k] = b.comp[k];
        for (int k = 0; k < G; k++) gb[k] = pL[(size_t)j * G + k];
        if (F.bc[i] == kAdiabaticSlipWall || F.bc[i] == kAdiabaticNonSlipWall) for (int d = 0; d < D; d++) gb[(D + 1) * D + d] = 0.0;
        double a[kMaxV], c[kMaxV];
        viscNormalFlux(P, n, L[j].comp, &pL[(size_t)j * G], a);
        viscNormalFlux(P, n, b.comp, gb, c);
        for (int v = 0; v < Nv; v++) Fc[v] -= (a[v] + c[v]) / 2.0;
      }
      if (O.av) {  // boundary faces: the interior side alone, calculateArtificialViscousNormalFlux (SpatialDiscrete.cpp:813-819)
        for (int v = 0; v < Nv; v++) {
          double a = 0.0;
          for (int c = 0; c < D; c++) a += epsL[j] * avgL[(size_t)j * G + v * D + c] * n[c];
          Fc[v] -= a;
        }
      }
      double* aL = &BL.vaq[((size_t)eL * BL.tab.Naq + offL + j) * Nv];
      for (int v = 0; v < Nv; v++) aL[v] = Fc[v] * w;
    }
  }
}
// R3 calculateElementResidual, SpatialDiscrete.cpp:1016-1032
static void sweepR3(Oracle& O, ElemBlock& B) {
  const Sizes s = sizes(O, B); const ElemTable& T = B.tab;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < B.n; e++) {
    double* r = &B.res[(size_t)e * s.Nv * s.Nb];
    gemm(s.Nv, s.Nb, s.Nq * s.D, 1.0, &B.vq[(size_t)e * s.Nv * s.Nq * s.D], s.Nv, T.dPhi.data(), s.Nq * s.D, 0.0, r, s.Nv);
    gemm(s.Nv, s.Nb, s.Naq, -1.0, &B.vaq[(size_t)e * s.Nv * s.Naq], s.Nv, T.PhiF.data(), s.Naq, 1.0, r, s.Nv);
    if (O.P.source != kSourceNone) gemm(s.Nv, s.Nb, s.Nq, 1.0, &B.sq[(size_t)e * s.Nv * s.Nq], s.Nv, T.Phi.data(), s.Nq, 1.0, r, s.Nv);
  }
}
// TimeIntegrationData<...>::kStepCoefficients, TimeIntegration.cpp:45-65
static void rkTable(int scheme, int& nstage, double c[3][3]) {
  const double FE[1][3] = {{1.0, 0.0, 1.0}};
  const double H2[2][3] = {{1.0, 0.0, 1.0}, {0.5, 0.5, 0.5}};
  const double S3[3][3] = {{1.0, 0.0, 1.0}, {3.0 / 4.0, 1.0 / 4.0, 1.0 / 4.0}, {1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0}};
  if (scheme == kForwardEuler) { nstage = 1; std::memcpy(c, FE, sizeof(FE)); }
  else if (scheme == kHeunRK2) { nstage = 2; std::memcpy(c, H2, sizeof(H2)); }
  else { nstage = 3; std::memcpy(c, S3, sizeof(S3)); }
}
// R4 updateElementBasisFunctionCoefficient, TimeIntegration.cpp:181-198
static void sweepR4(Oracle& O, ElemBlock& B, const double* c, double dt) {
  const Sizes s = sizes(O, B);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < B.n; e++) {
    double* U = &B.coef[(size_t)e * s.Nv * s.Nb];
    const double* UL = &B.coefLast[(size_t)e * s.Nv * s.Nb];
    for (int k = 0; k < s.Nv * s.Nb; k++) U[k] *= c[1];
    for (int k = 0; k < s.Nv * s.Nb; k++) U[k] += c[0] * UL[k];
    if (O.accurate) gemmLD(s.Nv, s.Nb, s.Nb, c[2] * dt, &B.res[(size_t)e * s.Nv * s.Nb], s.Nv, &B.MinvL[(size_t)e * s.Nb * s.Nb], s.Nb, 1.0, U, s.Nv);
    else gemm(s.Nv, s.Nb, s.Nb, c[2] * dt, &B.res[(size_t)e * s.Nv * s.Nb], s.Nv, &B.Minv[(size_t)e * s.Nb * s.Nb], s.Nb, 1.0, U, s.Nv);
  }
}

// Solver::calculateArtificialViscosity, SpatialDiscrete.cpp:37-192: per element the Persson-Peraire smoothness indicator of the density
// (energy of the modes above order P-1 against the energy of all modes), a constant, zero or sine-ramped viscosity per element, the
// maximum over the elements sharing each corner node, and the node values copied back to the elements' corners.
static void artificialViscosity(Oracle& O) {
  static const double kTol[5] = {0.0, -1.20411998266, -1.90848501888, -2.40823996531, -2.79588001734};   // SimulationControl.cpp:892-893
  const double tol = kTol[O.p - 1];
  const double kPi = 3.14159265358979323846;
  for (auto& bp : O.blk) if (bp) {
    ElemBlock& B = *bp; const Sizes s = sizes(O, B); const ElemTable& T = B.tab;
    if ((int)B.nodeTag.size() != B.n * T.nbasic || (int)B.innerRadius.size() != B.n) throw std::runtime_error("oracle: artificial viscosity needs orc_set_element_nodes for every block");
    const int nbLow = O.p == 1 ? 0 : numNodes(B.type, O.p - 1);   // getElementBasisFunctionNumber<type, P - 1>; P1: every mode is "high"
    B.avElem.assign((size_t)B.n * T.nbasic, 0.0);
#pragma omp parallel for schedule(static)
    for (int e = 0; e < B.n; e++) {
      const double* U = &B.coef[(size_t)e * s.Nv * s.Nb];
      double num = 0.0, den = 0.0;
      for (int q = 0; q < s.Nq; q++) {
        double all = 0.0, high = 0.0;
        for (int b = 0; b < s.Nb; b++) { const double t = T.Phi[(size_t)b * s.Nq + q] * U[b * s.Nv]; all += t; if (b >= nbLow) high += t; }
        const double w = B.jw[(size_t)e * s.Nq + q];
        num += high * (high * w); den += all * (all * w);
      }
      const double shock = std::log10(num / den);
      const double full = O.avFactor * (B.innerRadius[e] / O.p);
      double val;
      if (shock < tol - O.avTol) val = 0.0;
      else if (shock > tol + O.avTol) val = full;
      else val = full * (1.0 + std::sin(kPi * (shock - tol) / (2.0 * O.avTol))) / 2.0;
      for (int k = 0; k < T.nbasic; k++) B.avElem[(size_t)e * T.nbasic + k] = val;
    }
  }
  std::fill(O.nodeAV.begin(), O.nodeAV.end(), 0.0);
  for (auto& bp : O.blk) if (bp) {   // maxElementArtificialViscosity :89-108
    ElemBlock& B = *bp; const int nb = B.tab.nbasic;
    for (int e = 0; e < B.n; e++) for (int k = 0; k < nb; k++) { double& a = O.nodeAV[B.nodeTag[(size_t)e * nb + k]]; a = std::max(a, B.avElem[(size_t)e * nb + k]); }
  }
  for (auto& bp : O.blk) if (bp) {   // storeElementArtificialViscosity :110-122
    ElemBlock& B = *bp; const int nb = B.tab.nbasic;
    for (int e = 0; e < B.n; e++) for (int k = 0; k < nb; k++) B.avElem[(size_t)e * nb + k] = O.nodeAV[B.nodeTag[(size_t)e * nb + k]];
  }
}

static void evalResidual(Oracle& O) {  // one full residual evaluation: G1-G4 (if needed) + R1-R3
  const bool grad = O.P.ns() || O.deadGradient || O.av;
  if (grad) {
    for (auto& b : O.blk) if (b) sweepG1(O, *b);
    sweepG2(O);
    for (auto& b : O.blk) if (b) sweepG3(O, *b);
    for (auto& b : O.blk) if (b) sweepG4(O, *b);
  }
  for (auto& b : O.blk) if (b) sweepR1(O, *b);
  sweepR2(O);
  for (auto& b : O.blk) if (b) sweepR3(O, *b);
}

// calculateRelativeError, TimeIntegration.cpp:279-324
static void relativeError(Oracle& O) {
  const int Nv = O.P.Nv;
  double tot[kMaxV] = {0, 0, 0, 0, 0};
  for (auto& bp : O.blk) if (bp) {
    ElemBlock& B = *bp; const Sizes s = sizes(O, B); const ElemTable& T = B.tab;
    double acc[kMaxV] = {0, 0, 0, 0, 0};
#pragma omp parallel
    {
      double loc[kMaxV] = {0, 0, 0, 0, 0};
      std::vector<double> rq((size_t)Nv * s.Nq);
#pragma omp for schedule(static)
      for (int e = 0; e < B.n; e++) {
        gemmNT(Nv, s.Nq, s.Nb, 1.0, &B.res[(size_t)e * Nv * s.Nb], Nv, T.Phi.data(), s.Nq, 0.0, rq.data(), Nv);
        for (int v = 0; v < Nv; v++) { double m = 0; for (int q = 0; q < s.Nq; q++) m += std::fabs(rq[(size_t)q * Nv + v]); loc[v] += m / s.Nq; }
      }
#pragma omp critical
      for (int v = 0; v < Nv; v++) acc[v] += loc[v];
    }
    // each calculateElementRelativeError ASSIGNS its sum to the shared vector (TimeIntegration.cpp:294-297): on a mixed mesh the last
    // element type (ascending ElementEnum) wins, and the result is still divided by the number of all elements (:323)
    for (int v = 0; v < Nv; v++) tot[v] = acc[v];
  }
  const int ne = O.totalElems();
  for (int v = 0; v < Nv; v++) O.relErr[v] = tot[v] / ne;
}

// Solver::stepSolver, TimeIntegration.cpp:326-350
static void step(Oracle& O, double dt) {
  for (auto& b : O.blk) if (b) b->coefLast = b->coef;  // copyBasisFunctionCoefficient :70-102
  if (O.av) artificialViscosity(O);                    // once per step, before the stages (TimeIntegration.cpp:336-338)
  int ns; double c[3][3]; rkTable(O.rk, ns, c);
  for (int i = 0; i < ns; i++) {
    evalResidual(O);
    for (auto& b : O.blk) if (b) sweepR4(O, *b, c[i], dt);
  }
  relativeError(O);
}

}  // namespace orc

// =====================================================================================================================
// C API (ctypes).  Every function returns 0 on success; on failure the message is available via orc_last_error().
// =====================================================================================================================
using namespace orc;

struct orc_config {
  int32_t dim, p, model, eos, transport, conv_flux, visc_flux, source, rk, dead_gradient, accurate;
  double cp, cv, mu, c0, rho0, beta, t_ref;
};

#define ORC_TRY try {
#define ORC_CATCH } catch (const std::exception& ex) { g_err = ex.what(); return 1; } return 0;

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

int orc_create(const orc_config* c, void** out) {
  ORC_TRY
  auto* O = new Oracle();
  Phys& P = O->P;
  P.D = c->dim; P.Nv = c->dim + 2; P.model = c->model; P.eos = c->eos; P.transport = c->transport; P.conv = c->conv_flux;
  P.visc = c->visc_flux; P.source = c->source; P.cp = c->cp; P.cv = c->cv; P.mu0 = c->mu;
  P.k0 = c->cp * c->mu / 0.71;  // calculateThermalConductivityFromDynamicViscosity, PhysicalModel.cpp:152-156 (Pr = 0.71)
  P.c0 = c->c0; P.rho0 = c->rho0; P.padd = 0.01 * c->rho0 * c->c0 * c->c0;  // PhysicalModel.cpp:63-66
  P.beta = c->beta; P.Tref = c->t_ref;
  O->p = c->p; O->rk = c->rk; O->deadGradient = c->dead_gradient != 0; O->accurate = c->accurate != 0;
  if (P.ns() && P.visc == kViscNone) throw std::runtime_error("oracle: NS model needs BR1 or BR2");
  if (!P.ns()) P.visc = kViscNone;
  if (c->p < 1 || c->p > 5 || c->dim < 1 || c->dim > 3) throw std::runtime_error("oracle: dim/p out of range");
  *out = O;
  ORC_CATCH
}
void orc_destroy(void* h) { delete (Oracle*)h; }

// node coordinates: n x nn x D in gmsh node order of Lagrange order geom_order
int orc_add_elements(void* h, int type, int n, int geom_order, const double* coords) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (elemDim(type) != O.P.D) throw std::runtime_error("oracle: element dimension mismatch");
  auto B = std::make_unique<ElemBlock>();
  B->type = type; B->n = n; B->tab.build(type, O.p, geom_order);
  B->X.assign(coords, coords + (size_t)n * B->tab.nn * O.P.D);
  O.blk[type] = std::move(B);
  ORC_CATCH
}

// Faces: interior first, then boundary (reference order).  For boundary faces the right_* entries are ignored.
int orc_set_faces(void* h, int n_int, int n_bnd, const int32_t* le, const int32_t* lt, const int32_t* lf, const int32_t* re,
                  const int32_t* rt, const int32_t* rf, const int32_t* rot, const int32_t* bc, const int32_t* phys) {
  ORC_TRY
  Oracle& O = *(Oracle*)h; FaceSet& F = O.F;
  const int nf = n_int + n_bnd;
  F.nInt = n_int; F.nBnd = n_bnd;
  F.elem[0].assign(le, le + nf); F.etype[0].assign(lt, lt + nf); F.lface[0].assign(lf, lf + nf);
  F.elem[1].assign(re, re + nf); F.etype[1].assign(rt, rt + nf); F.lface[1].assign(rf, rf + nf);
  F.rot.assign(rot, rot + nf); F.bc.assign(bc, bc + nf); F.phys.assign(phys, phys + nf);
  ORC_CATCH
}

int orc_finalize(void* h) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  int ft = -1;
  for (auto& b : O.blk) if (b) { int t = faceType(b->type); if (ft >= 0 && ft != t) throw std::runtime_error("oracle: mixed face types unsupported"); ft = t; }
  if (ft < 0) throw std::runtime_error("oracle: no elements");
  O.F.ftype = ft;
  O.F.Nqf = makeQuadrature(ft, 2 * O.p + 1).n;
  for (auto& b : O.blk) if (b) { elementGeometry(*b, O.accurate); allocSolver(O, *b); }
  const int nf = O.F.nInt + O.F.nBnd;
  for (int i = 0; i < nf; i++) {
    for (int s = 0; s < (i < O.F.nInt ? 2 : 1); s++) {
      const int t = O.F.etype[s][i];
      if (t < 0 || t > 6 || !O.blk[t] || O.F.elem[s][i] < 0 || O.F.elem[s][i] >= O.blk[t]->n || O.F.lface[s][i] < 0 || O.F.lface[s][i] >= O.blk[t]->tab.Nf)
        throw std::runtime_error("oracle: face record out of range");
    }
  }
  faceGeometry(O);
  O.F.dummy.assign((size_t)O.F.nBnd * O.F.Nqf, Var());
  O.finalized = true;
  ORC_CATCH
}

int orc_sizes(void* h, int type, int32_t* out /* n, Nb, Nq, Nf, Naq, nn, Nqf, Nv */) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  const ElemTable& T = O.blk[type]->tab;
  out[0] = O.blk[type]->n; out[1] = T.Nb; out[2] = T.Nq; out[3] = T.Nf; out[4] = T.Naq; out[5] = T.nn; out[6] = T.fquad.n; out[7] = O.P.Nv;
  ORC_CATCH
}

// ---- table / geometry access for tests -------------------------------------------------------------------------------
// which: 0 Phi(Nq x Nb) 1 dPhi(Nq*D x Nb) 2 PhiF(Naq x Nb) 3 LSinv(Nb x Nb) 4 quad pts(Nq x 3 row-major) 5 quad wts
//        6 face quad pts (Nqf x 3) 7 face quad wts 8 face points in parent coordinates (Naq x 3)
int orc_get_table(void* h, int type, int which, double* out) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  const ElemTable& T = O.blk[type]->tab;
  const std::vector<double>* v = nullptr;
  switch (which) {
    case 0: v = &T.Phi; break; case 1: v = &T.dPhi; break; case 2: v = &T.PhiF; break; case 3: v = &T.LSinv; break;
    case 4: v = &T.quad.pts; break; case 5: v = &T.quad.wts; break; case 6: v = &T.fquad.pts; break; case 7: v = &T.fquad.wts; break;
    case 8: v = &T.fxi; break;
    default: throw std::runtime_error("oracle: bad table id");
  }
  std::memcpy(out, v->data(), sizeof(double) * v->size());
  ORC_CATCH
}
// which: 0 xq (n x Nq x D) 1 jw (n x Nq) 2 mt (n x Nq x D*D) 3 Minv (n x Nb x Nb) 4 minEdge (n)
int orc_get_element_geometry(void* h, int type, int which, double* out) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  const ElemBlock& B = *O.blk[type];
  const std::vector<double>* v = which == 0 ? &B.xq : which == 1 ? &B.jw : which == 2 ? &B.mt : which == 3 ? &B.Minv : &B.minEdge;
  std::memcpy(out, v->data(), sizeof(double) * v->size());
  ORC_CATCH
}
// which: 0 xf (nf x Nqf x D) 1 normals (nf x Nqf x D) 2 jw (nf x Nqf)
int orc_get_face_geometry(void* h, int which, double* out) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  const std::vector<double>* v = which == 0 ? &O.F.xf : which == 1 ? &O.F.nrm : &O.F.jw;
  std::memcpy(out, v->data(), sizeof(double) * v->size());
  ORC_CATCH
}
// max over interior faces/points of | (xL_j - cL) - (xR_seq[j] - cR) |: checks rotation + permutation tables geometrically
int orc_check_face_match(void* h, double* out) {
  ORC_TRY
  Oracle& O = *(Oracle*)h; FaceSet& F = O.F;
  const int D = O.P.D, Nqf = F.Nqf;
  double worst = 0;
  for (int i = 0; i < F.nInt; i++) {
    const ElemBlock& B = *O.blk[F.etype[1][i]]; const ElemTable& T = B.tab;
    const int e = F.elem[1][i], f = F.lface[1][i];
    const double* X = &B.X[(size_t)e * T.nn * D];
    std::vector<double> xr((size_t)Nqf * D, 0.0);
    for (int j = 0; j < Nqf; j++) for (int m = 0; m < T.nn; m++) for (int l = 0; l < D; l++) xr[(size_t)j * D + l] += T.GNf[(size_t)m * T.Naq + T.off[f] + j] * X[m * D + l];
    double cl[3] = {0, 0, 0}, cr[3] = {0, 0, 0};
    for (int j = 0; j < Nqf; j++) for (int l = 0; l < D; l++) { cl[l] += F.xf[((size_t)i * Nqf + j) * D + l] / Nqf; cr[l] += xr[(size_t)j * D + l] / Nqf; }
    const std::vector<int> seq = faceQuadratureSequence(F.ftype, O.p, F.rot[i]);
    for (int j = 0; j < Nqf; j++) for (int l = 0; l < D; l++)
      worst = std::max(worst, std::fabs((F.xf[((size_t)i * Nqf + j) * D + l] - cl[l]) - (xr[(size_t)seq[j] * D + l] - cr[l])));
  }
  *out = worst;
  ORC_CATCH
}
int orc_reference_nodes(int type, int order, double* out /* nn x 3 */, int32_t* count) {
  ORC_TRY
  auto nodes = referenceNodes(type, order);
  *count = (int)nodes.size();
  if (out) for (size_t i = 0; i < nodes.size(); i++) for (int k = 0; k < 3; k++) out[i * 3 + k] = nodes[i][k];
  ORC_CATCH
}
int orc_face_sequence(int ftype, int p, int rotation, int32_t* out) {
  ORC_TRY
  auto s = faceQuadratureSequence(ftype, p, rotation);
  for (size_t i = 0; i < s.size(); i++) out[i] = s[i];
  ORC_CATCH
}

// ---- state ------------------------------------------------------------------------------------------------------------
// initializeElementSolver, InitialCondition.cpp:85-116: primitive values at the quadrature nodes (n x Nq x Nv) ->
// conserved -> U = Uq * Phi * (Phi^T Phi)^-1
int orc_set_state_from_primitive(void* h, int type, const double* prim) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  ElemBlock& B = *O.blk[type]; const Sizes s = sizes(O, B); const ElemTable& T = B.tab;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < B.n; e++) {
    std::vector<double> uq((size_t)s.Nv * s.Nq), t((size_t)s.Nv * s.Nb);
    for (int q = 0; q < s.Nq; q++) {
      Var v; for (int k = 0; k < s.Nv; k++) v.prim[k] = prim[((size_t)e * s.Nq + q) * s.Nv + k];
      consFromPrim(O.P, v);
      for (int k = 0; k < s.Nv; k++) uq[(size_t)q * s.Nv + k] = v.cons[k];
    }
    if (O.accurate) {
      // Uq Phi (Phi^T Phi)^-1 with every product accumulated in extended precision
      std::vector<long double> tl((size_t)s.Nv * s.Nb);
      for (int b = 0; b < s.Nb; b++) for (int v = 0; v < s.Nv; v++) { long double a = 0; for (int q = 0; q < s.Nq; q++) a += (long double)uq[(size_t)q * s.Nv + v] * (long double)T.Phi[(size_t)b * s.Nq + q]; tl[(size_t)b * s.Nv + v] = a; }
      for (int b = 0; b < s.Nb; b++) for (int v = 0; v < s.Nv; v++) { long double a = 0; for (int l = 0; l < s.Nb; l++) a += tl[(size_t)l * s.Nv + v] * T.LSinvL[(size_t)b * s.Nb + l]; B.coef[((size_t)e * s.Nb + b) * s.Nv + v] = (double)a; }
    } else {
      gemm(s.Nv, s.Nb, s.Nq, 1.0, uq.data(), s.Nv, T.Phi.data(), s.Nq, 0.0, t.data(), s.Nv);
      gemm(s.Nv, s.Nb, s.Nb, 1.0, t.data(), s.Nv, T.LSinv.data(), s.Nb, 0.0, &B.coef[(size_t)e * s.Nv * s.Nb], s.Nv);
    }
  }
  ORC_CATCH
}
// initializeAdjacencyElementSolver / updateAdjacencyElementBoundaryVariable (InitialCondition.cpp:118-149,
// BoundaryCondition.cpp:29-51): user primitive values at the boundary-face quadrature nodes (nBnd x Nqf x Nv)
int orc_set_boundary_primitive(void* h, const double* prim) {
  ORC_TRY
  Oracle& O = *(Oracle*)h; FaceSet& F = O.F;
  const int Nv = O.P.Nv;
  for (size_t k = 0; k < (size_t)F.nBnd * F.Nqf; k++) {
    Var v; for (int i = 0; i < Nv; i++) v.prim[i] = prim[k * Nv + i];
    consFromPrim(O.P, v); compFromPrim(O.P, v);
    F.dummy[k] = v;
  }
  ORC_CATCH
}
int orc_get_state(void* h, int type, double* U) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  std::memcpy(U, O.blk[type]->coef.data(), sizeof(double) * O.blk[type]->coef.size());
  ORC_CATCH
}
int orc_set_state(void* h, int type, const double* U) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  std::memcpy(O.blk[type]->coef.data(), U, sizeof(double) * O.blk[type]->coef.size());
  ORC_CATCH
}
// conserved variables at the volume quadrature nodes (n x Nq x Nv): basis-invariant view of the state
int orc_get_state_at_quadrature(void* h, int type, double* out) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  ElemBlock& B = *O.blk[type]; const Sizes s = sizes(O, B);
#pragma omp parallel for schedule(static)
  for (int e = 0; e < B.n; e++)
    gemmNT(s.Nv, s.Nq, s.Nb, 1.0, &B.coef[(size_t)e * s.Nv * s.Nb], s.Nv, B.tab.Phi.data(), s.Nq, 0.0, &out[(size_t)e * s.Nq * s.Nv], s.Nv);
  ORC_CATCH
}
// calculateElementDeltaTime, TimeIntegration.cpp:104-131
int orc_compute_dt(void* h, double cfl, double* dt) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  double best = 1.7976931348623157e308;
  for (auto& bp : O.blk) if (bp) {
    ElemBlock& B = *bp; const Sizes s = sizes(O, B);
#pragma omp parallel for schedule(static) reduction(min : best)
    for (int e = 0; e < B.n; e++) {
      std::vector<double> uq((size_t)s.Nv * s.Nq);
      gemmNT(s.Nv, s.Nq, s.Nb, 1.0, &B.coef[(size_t)e * s.Nv * s.Nb], s.Nv, B.tab.Phi.data(), s.Nq, 0.0, uq.data(), s.Nv);
      for (int q = 0; q < s.Nq; q++) {
        Var v; for (int k = 0; k < s.Nv; k++) v.cons[k] = uq[(size_t)q * s.Nv + k];
        compFromCons(O.P, v);
        const double sr = std::sqrt(vsq(O.P, v.comp)) + O.P.sound(v.comp[0], v.comp[s.D + 2]);
        best = std::min(best, cfl * B.minEdge[e] / (sr * (O.p + 1.0) * (O.p + 1.0)));
      }
    }
  }
  *dt = best;
  ORC_CATCH
}
int orc_step(void* h, double dt, int nsteps, double* rel_err) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.finalized) throw std::runtime_error("oracle: not finalized");
  for (int i = 0; i < nsteps; i++) step(O, dt);
  if (rel_err) for (int v = 0; v < O.P.Nv; v++) rel_err[v] = O.relErr[v];
  ORC_CATCH
}
// One residual evaluation (G1-G4, R1-R3) of the current state, all element types.
int orc_eval_residual(void* h) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.finalized) throw std::runtime_error("oracle: not finalized");
  if (O.av) artificialViscosity(O);   // parity hook: the viscosity of the CURRENT state (stepSolver evaluates it once per step)
  evalResidual(O);
  ORC_CATCH
}
// Fetch after orc_eval_residual.  Rmodal: variable_residual_ (n x Nb x Nv, basis dependent);
// rhsq: (R * M^-1) * Phi^T at the quadrature nodes (n x Nq x Nv, basis invariant = dU/dt).  Either may be null.
int orc_fetch_residual(void* h, int type, double* Rmodal, double* rhsq) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  ElemBlock& B = *O.blk[type]; const Sizes s = sizes(O, B);
  if (Rmodal) std::memcpy(Rmodal, B.res.data(), sizeof(double) * B.res.size());
  if (rhsq) {
#pragma omp parallel for schedule(static)
    for (int e = 0; e < B.n; e++) {
      std::vector<double> t((size_t)s.Nv * s.Nb);
      if (O.accurate) gemmLD(s.Nv, s.Nb, s.Nb, 1.0, &B.res[(size_t)e * s.Nv * s.Nb], s.Nv, &B.MinvL[(size_t)e * s.Nb * s.Nb], s.Nb, 0.0, t.data(), s.Nv);
      else gemm(s.Nv, s.Nb, s.Nb, 1.0, &B.res[(size_t)e * s.Nv * s.Nb], s.Nv, &B.Minv[(size_t)e * s.Nb * s.Nb], s.Nb, 0.0, t.data(), s.Nv);
      gemmNT(s.Nv, s.Nq, s.Nb, 1.0, t.data(), s.Nv, B.tab.Phi.data(), s.Nq, 0.0, &rhsq[(size_t)e * s.Nq * s.Nv], s.Nv);
    }
  }
  ORC_CATCH
}
// total gradient coefficients evaluated at the quadrature nodes (n x Nq x Nv*D): NS -> variable_gradient_..., Euler -> volume gradient
int orc_get_gradient_at_quadrature(void* h, int type, double* out) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  ElemBlock& B = *O.blk[type]; const Sizes s = sizes(O, B);
  const std::vector<double>& C = O.P.ns() ? B.gcoef : B.gcoefVol;
#pragma omp parallel for schedule(static)
  for (int e = 0; e < B.n; e++)
    gemmNT(s.G, s.Nq, s.Nb, 1.0, &C[(size_t)e * s.G * s.Nb], s.G, B.tab.Phi.data(), s.Nq, 0.0, &out[(size_t)e * s.Nq * s.G], s.G);
  ORC_CATCH
}
// RawBinary payload (RawBinary.cpp:75-88): variable_gradient_basis_function_coefficient_ of every element, [n][Nb][Nv*D]
int orc_get_gradient_state(void* h, int type, double* out) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  ElemBlock& B = *O.blk[type];
  const std::vector<double>& C = O.P.ns() ? B.gcoef : B.gcoefVol;
  std::copy(C.begin(), C.end(), out);
  ORC_CATCH
}
// RawBinary payload (RawBinary.cpp:89-154): per boundary face, in face order, the gradient coefficients written next to the parent's
// state: BR1 the total gradient, BR2 variable_volume_gradient_ + variable_interface_gradient_ of THAT face.  Rows are Nb(parent) x Nv*D.
int orc_get_boundary_gradient_state(void* h, double* out) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  const FaceSet& F = O.F;
  const int G = O.P.Nv * O.P.D;
  size_t at = 0;
  for (int b = 0; b < F.nBnd; b++) {
    const int i = F.nInt + b, t = F.etype[0][i], e = F.elem[0][i], f = F.lface[0][i];
    const ElemBlock& B = *O.blk[t]; const int Nb = B.tab.Nb;
    const size_t len = (size_t)G * Nb;
    if (O.P.visc == kBR2) {
      const double* a = &B.gcoefVol[(size_t)e * len];
      const double* c = &B.gicoef[((size_t)e * B.tab.Nf + f) * len];
      for (size_t k = 0; k < len; k++) out[at + k] = a[k] + c[k];
    } else {
      const double* a = &(O.P.ns() ? B.gcoef : B.gcoefVol)[(size_t)e * len];
      for (size_t k = 0; k < len; k++) out[at + k] = a[k];
    }
    at += len;
  }
  ORC_CATCH
}
// ViewVariable::get (VariableConvertor.cpp:754-872) at one point: computational variables, primitive gradient (zeros for Euler models), the
// interpolated artificial viscosity.  The switch of the reference falls through where a variable does not exist for the equation set.
static double viewValue(const Phys& P, int D, bool ns, const Var& v, const double* gp, double eps, int variable) {
      const double rho = v.comp[0], p = v.comp[D + 2];
      double v2 = 0; for (int d = 0; d < D; d++) v2 += v.comp[1 + d] * v.comp[1 + d];
      const double c = P.eos == kIdealGas ? std::sqrt(1.4 * p / rho) : P.c0;           // PhysicalModel.cpp:51-54,74-77
      auto dU = [&](int comp, int dir) { return gp[(1 + comp) * D + dir]; };
      double r = 0.0;
      int w = variable;
      for (;;) {
        if (w == 0) { r = rho; break; }
        if (w == 1) { r = std::sqrt(v2); break; }
        if (w == 2) { r = v.comp[D + 1] / P.cv; break; }
        if (w == 3) { r = p; break; }
        if (w == 4) { r = c; break; }
        if (w == 5) { r = std::sqrt(v2) / c; break; }
        if (w == 6) { if (P.comp()) { r = p / std::pow(rho, 1.4); break; } w = 7; continue; }
        if (w == 7) {
          if (ns && D == 2) { r = dU(1, 0) - dU(0, 1); break; }
          if (ns && D == 3) { const double a = dU(2, 1) - dU(1, 2), b = dU(0, 2) - dU(2, 0), cc = dU(1, 0) - dU(0, 1); r = std::sqrt(a * a + b * b + cc * cc); break; }
          w = 9; continue;
        }
        if (w == 9) { r = eps; break; }
        if (w >= 10 && w <= 12) { if (w - 10 >= D) throw std::runtime_error("oracle: view variable outside the dimension"); r = v.comp[1 + (w - 10)]; break; }
        if (w >= 13 && w <= 15) { if (w - 13 >= D) throw std::runtime_error("oracle: view variable outside the dimension"); r = v.comp[1 + (w - 13)] / c; break; }
        if (ns && w >= 16 && w <= 21) {
          if ((w == 16 || w == 17 || w == 21) && D < 3) throw std::runtime_error("oracle: view variable outside the dimension");
          if ((w == 18 || w == 20) && D < 2) throw std::runtime_error("oracle: view variable outside the dimension");
          if (w == 16) r = dU(2, 1) - dU(1, 2);
          else if (w == 17) r = dU(0, 2) - dU(2, 0);
          else if (w == 18) r = dU(1, 0) - dU(0, 1);
          else r = gp[(D + 1) * D + (w - 19)];
          break;
        }
        r = 0.0; break;
      }
      return r;
}
// ViewVariable::get (VariableConvertor.cpp:754-872) at the volume quadrature points, n x Nq.  variable = ViewVariableEnum value.  The
// switch of the reference falls through where a variable does not exist for the equation set; `goto`-free restatement of that chain.
int orc_view_variable(void* h, int type, int variable, double* out) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  ElemBlock& B = *O.blk[type]; const Sizes s = sizes(O, B); const ElemTable& T = B.tab; const Phys& P = O.P;
  const int D = s.D, Nv = s.Nv, G = s.G;
  const bool ns = P.ns();
  for (int e = 0; e < B.n; e++) {
    std::vector<double> uq((size_t)Nv * s.Nq), gq((size_t)G * s.Nq, 0.0);
    gemmNT(Nv, s.Nq, s.Nb, 1.0, &B.coef[(size_t)e * Nv * s.Nb], Nv, T.Phi.data(), s.Nq, 0.0, uq.data(), Nv);
    if (ns) gemmNT(G, s.Nq, s.Nb, 1.0, &B.gcoef[(size_t)e * G * s.Nb], G, T.Phi.data(), s.Nq, 0.0, gq.data(), G);
    for (int q = 0; q < s.Nq; q++) {
      Var v;
      for (int k = 0; k < Nv; k++) v.cons[k] = uq[(size_t)q * Nv + k];
      compFromCons(P, v);
      double gp[kMaxD * kMaxV] = {0};
      if (ns) primGradFromConsGrad(P, v, &gq[(size_t)q * G], gp);
      double eps = 0.0;
      if (O.av) for (int k = 0; k < T.nbasic; k++) eps += T.NodalQ[(size_t)k * s.Nq + q] * B.avElem[(size_t)e * T.nbasic + k];
      const double r = viewValue(P, D, ns, v, gp, eps, variable);
      out[(size_t)e * s.Nq + q] = r;
    }
  }
  ORC_CATCH
}
// System::setArtificialViscosity (SystemControl.cpp:105-108) with ShockCapturingEnum::ArtificialViscosity; n_nodes = Mesh::node_number_
int orc_set_artificial_viscosity(void* h, double empirical_tolerance, double factor, int n_nodes) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (O.P.ns()) throw std::runtime_error("oracle: artificial viscosity is restated for the Euler models only");
  O.av = true; O.avTol = empirical_tolerance; O.avFactor = factor; O.nodeAV.assign((size_t)n_nodes, 0.0);
  ORC_CATCH
}
// mesh data of one block: node_tag_ of the corner nodes (0-based, n x nbasic) and inner_radius_ (n), ReadControl.cpp:64,91
int orc_set_element_nodes(void* h, int type, const int32_t* tags, const double* inner_radius) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  ElemBlock& B = *O.blk[type];
  B.nodeTag.assign(tags, tags + (size_t)B.n * B.tab.nbasic);
  B.innerRadius.assign(inner_radius, inner_radius + B.n);
  B.avElem.assign((size_t)B.n * B.tab.nbasic, 0.0);
  for (int t : B.nodeTag) if (t < 0 || t >= (int)O.nodeAV.size()) throw std::runtime_error("oracle: node tag out of range (call orc_set_artificial_viscosity first)");
  ORC_CATCH
}
int orc_update_artificial_viscosity(void* h) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (O.av) artificialViscosity(O);
  ORC_CATCH
}
int orc_get_node_artificial_viscosity(void* h, double* out) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  std::copy(O.nodeAV.begin(), O.nodeAV.end(), out);
  ORC_CATCH
}
int orc_get_element_artificial_viscosity(void* h, int type, double* out) {
  ORC_TRY
  Oracle& O = *(Oracle*)h;
  if (!O.blk[type]) throw std::runtime_error("oracle: no such element block");
  std::copy(O.blk[type]->avElem.begin(), O.blk[type]->avElem.end(), out);
  ORC_CATCH
}
// Pointwise physics of the oracle behind the SAME entry point as oracle/ref_physics.cpp (the reference's own functions): what = 0 Riemann
// flux, 1 boundary face point, 2 viscous terms, 3 conversions / raw flux / source.  cfg = {dim, model, eos, transport, conv_flux, source},
// params = {cp, cv, mu, c0, rho0, beta, t_ref}; input / output layouts as documented there.  Pins the restatement against
// tests/golden/reference_physics.json.
int orc_physics(const int32_t* cfg, const double* params, int what, int bc, int n, const double* in, double* out) {
  ORC_TRY
  Phys P;
  P.D = cfg[0]; P.Nv = cfg[0] + 2; P.model = cfg[1]; P.eos = cfg[2]; P.transport = cfg[3]; P.conv = cfg[4]; P.source = cfg[5];
  P.visc = P.ns() ? kBR2 : kViscNone;
  P.cp = params[0]; P.cv = params[1]; P.mu0 = params[2]; P.k0 = params[0] * params[2] / 0.71;
  P.c0 = params[3]; P.rho0 = params[4]; P.padd = 0.01 * params[4] * params[3] * params[3]; P.beta = params[5]; P.Tref = params[6];
  const int D = P.D, Nv = P.Nv, NC = D + 3, G = Nv * D;
  for (int i = 0; i < n; i++) {
    if (what == 0) {
      const double* a = in + (size_t)i * (D + 2 * Nv);
      Var L, R;
      for (int k = 0; k < Nv; k++) { L.cons[k] = a[D + k]; R.cons[k] = a[D + Nv + k]; }
      compFromCons(P, L); compFromCons(P, R);
      convFlux(P, a, L, R, out + (size_t)i * Nv);
    } else if (what == 1) {
      const int ni = D + 2 * Nv + G, no = NC + 3 * Nv + (P.ns() ? NC + Nv : 0);
      const double* a = in + (size_t)i * ni; double* o = out + (size_t)i * no;
      Var L, dummy, b;
      for (int k = 0; k < Nv; k++) { L.cons[k] = a[D + k]; dummy.prim[k] = a[D + Nv + k]; }
      compFromCons(P, L);
      consFromPrim(P, dummy); compFromPrim(P, dummy);
      bcBoundaryVariable(P, bc, a, L, dummy, b.comp);
      bcBoundaryGradientVariable(P, bc, a, L, dummy, o + NC, o + NC + Nv);
      for (int k = 0; k < NC; k++) o[k] = b.comp[k];
      convNormalFlux(P, a, b.comp, o + NC + 2 * Nv);
      if (P.ns()) {
        double pL[kMaxD * kMaxV], gb[kMaxD * kMaxV], va[kMaxV], vb[kMaxV];
        primGradFromConsGrad(P, L, a + D + 2 * Nv, pL);
        if (bcIsWall(bc)) for (int k = 0; k < NC; k++) L.comp[k] = b.comp[k];
        for (int k = 0; k < G; k++) gb[k] = pL[k];
        if (bc == kAdiabaticSlipWall || bc == kAdiabaticNonSlipWall) for (int d = 0; d < D; d++) gb[(D + 1) * D + d] = 0.0;
        viscNormalFlux(P, a, L.comp, pL, va);
        viscNormalFlux(P, a, b.comp, gb, vb);
        for (int k = 0; k < NC; k++) o[NC + 3 * Nv + k] = L.comp[k];
        for (int k = 0; k < Nv; k++) o[2 * NC + 3 * Nv + k] = (va[k] + vb[k]) / 2.0;
      }
    } else if (what == 2) {
      if (!P.ns()) throw std::runtime_error("viscous terms need a Navier-Stokes model");
      const int ni = D + Nv + G, no = 2 * G + Nv;
      const double* a = in + (size_t)i * ni; double* o = out + (size_t)i * no;
      Var V;
      for (int k = 0; k < Nv; k++) V.cons[k] = a[D + k];
      compFromCons(P, V);
      primGradFromConsGrad(P, V, a + D + Nv, o);
      viscRawFlux(P, V.comp, o, o + G);
      viscNormalFlux(P, a, V.comp, o, o + 2 * G);
    } else if (what == 3) {
      const int no = NC + Nv + G + Nv;
      double* o = out + (size_t)i * no;
      Var V;
      for (int k = 0; k < Nv; k++) V.cons[k] = in[(size_t)i * Nv + k];
      compFromCons(P, V);
      for (int k = 0; k < NC; k++) o[k] = V.comp[k];
      o[NC] = V.comp[0]; for (int d = 0; d < D; d++) o[NC + 1 + d] = V.comp[1 + d]; o[NC + D + 1] = P.TFromE(V.comp[D + 1]);   // VariableConvertor.cpp:383-421
      convRawFlux(P, V.comp, o + NC + Nv);
      for (int k = 0; k < Nv; k++) o[NC + Nv + G + k] = 0.0;
      if (P.source != kSourceNone) sourceTerm(P, V.comp, o + NC + Nv + G);
    } else if (what == 4) {   // the 22 ViewVariableEnum values at one point: in = cons[Nv], conserved gradient[G], artificial viscosity
      const double* a = in + (size_t)i * (Nv + G + 1); double* o = out + (size_t)i * 22;
      Var V;
      for (int k = 0; k < Nv; k++) V.cons[k] = a[k];
      compFromCons(P, V);
      double gp[kMaxD * kMaxV] = {0};
      if (P.ns()) primGradFromConsGrad(P, V, a + Nv, gp);
      for (int w = 0; w < 22; w++) {
        const bool needs3 = w == 12 || w == 15 || w == 16 || w == 17 || w == 21, needs2 = w == 11 || w == 14 || w == 18 || w == 20;
        o[w] = ((needs3 && D < 3) || (needs2 && D < 2)) ? 0.0 : viewValue(P, D, P.ns(), V, gp, a[Nv + G], w);
      }
    } else {
      throw std::runtime_error("orc_physics: bad selector");
    }
  }
  ORC_CATCH
}

int orc_set_threads(int n) { omp_set_num_threads(n); return 0; }
int orc_max_threads() { return omp_get_max_threads(); }

}  // extern "C"
