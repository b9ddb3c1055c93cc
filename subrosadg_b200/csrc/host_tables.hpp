// host_tables.hpp — reference-element tables of the PRODUCT library (host side, built once at sdg_finalize).
//
// SubrosaDG obtains these from the Gmsh 4.13.1 library at run time (src/Mesh/Quadrature.cpp:27-34,
// src/Mesh/BasisFunction.cpp:31-74,136-230) and embeds the integer conventions in src/Solver/SimulationControl.cpp:26-523.
// The device path works in the collocation basis of the volume Gauss points for quadrangle/hexahedron (Nq == Nb,
// SimulationControl.cpp:268-273), so what is needed here is 1-D only: Gauss abscissae/weights, the differentiation matrix
// of the Lagrange polynomials through them, their end-point values, the 1-D Lobatto ("H1Legendre") shape functions at the
// Gauss points (seam transform to the reference's modal coefficients) and the gmsh node lattice of the geometry nodes.
// Independent of oracle/ by construction (nothing under oracle/ is included or linked).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace sdg {

enum ElemType { kPoint = 0, kLine = 1, kTriangle = 2, kQuadrangle = 3, kTetrahedron = 4, kPyramid = 5, kHexahedron = 6 };  // Enum.cpp:28-36

// ---- 1-D Legendre machinery -------------------------------------------------------------------------------------------
// P_n(x) and P_n'(x) by the three-term recurrence.
inline void legendrePair(int n, double x, double& p, double& dp) {
  double p0 = 1.0, p1 = x;
  if (n == 0) { p = 1.0; dp = 0.0; return; }
  for (int k = 2; k <= n; k++) { const double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k; p0 = p1; p1 = pk; }
  p = p1;
  dp = n * (x * p1 - p0) / (x * x - 1.0);
}
inline double legendreValue(int n, double x) {
  double p0 = 1.0, p1 = x;
  if (n == 0) return 1.0;
  for (int k = 2; k <= n; k++) { const double pk = ((2 * k - 1) * x * p1 - (k - 1) * p0) / k; p0 = p1; p1 = pk; }
  return p1;
}
// n-point Gauss-Legendre rule, ascending ("Gauss{o}" with n = o/2+1 points per direction; the counts are pinned by
// SimulationControl.cpp:268-273).
inline void gaussRule(int n, std::vector<double>& x, std::vector<double>& w) {
  x.assign(n, 0.0); w.assign(n, 0.0);
  const double pi = 3.14159265358979323846;
  for (int i = 0; i < (n + 1) / 2; i++) {
    double t = std::cos(pi * (i + 0.75) / (n + 0.5)), p, dp;  // i-th root from the right
    for (int it = 0; it < 60; it++) { legendrePair(n, t, p, dp); const double dt = p / dp; t -= dt; if (std::fabs(dt) < 1e-17) break; }
    legendrePair(n, t, p, dp);
    const double wt = 2.0 / ((1.0 - t * t) * dp * dp);
    x[n - 1 - i] = t; x[i] = -t; w[n - 1 - i] = wt; w[i] = wt;
  }
  if (n % 2) x[n / 2] = 0.0;
}
// Lobatto shape functions of Solin (docs/develop-note/develop-note.tex:284): l0=(1-x)/2, l1=(1+x)/2,
// l_k=(L_k-L_{k-2})/sqrt(2(2k-1)).
inline double lobattoShape(int k, double x) {
  if (k == 0) return 0.5 * (1.0 - x);
  if (k == 1) return 0.5 * (1.0 + x);
  return (legendreValue(k, x) - legendreValue(k - 2, x)) / std::sqrt(2.0 * (2.0 * k - 1.0));
}
// value / derivative at x of the Lagrange polynomials through arbitrary distinct nodes
inline void lagrangeAt(const std::vector<double>& nodes, double x, std::vector<double>& val, std::vector<double>& der) {
  const int n = (int)nodes.size();
  val.assign(n, 0.0); der.assign(n, 0.0);
  for (int j = 0; j < n; j++) {
    double v = 1.0;
    for (int m = 0; m < n; m++) if (m != j) v *= (x - nodes[m]) / (nodes[j] - nodes[m]);
    val[j] = v;
    double d = 0.0;
    for (int i = 0; i < n; i++) if (i != j) {
      double t = 1.0 / (nodes[j] - nodes[i]);
      for (int m = 0; m < n; m++) if (m != j && m != i) t *= (x - nodes[m]) / (nodes[j] - nodes[m]);
      d += t;
    }
    der[j] = d;
  }
}

// ---- gmsh element conventions ---------------------------------------------------------------------------------------------
inline int elemDim(int t) { return t == kLine ? 1 : (t == kTriangle || t == kQuadrangle) ? 2 : t == kPoint ? 0 : 3; }
inline int numFaces(int t) { static const int n[7] = {0, 2, 3, 4, 4, 5, 6}; return n[t]; }        // SimulationControl.cpp:99-122
inline int faceType(int t) { return t == kLine ? kPoint : (t == kTriangle || t == kQuadrangle) ? kLine : t == kHexahedron ? kQuadrangle : kTriangle; }
inline bool isTensor(int t) { return t == kLine || t == kQuadrangle || t == kHexahedron; }

// corner lattice bits (0 = low, 1 = high per axis) of quadrangle / hexahedron corners in gmsh order
static const int kQuadCorner[4][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}};
static const int kHexCorner[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
static const int kHexEdge[12][2] = {{0, 1}, {0, 3}, {0, 4}, {1, 2}, {1, 5}, {2, 3}, {2, 6}, {3, 7}, {4, 5}, {4, 7}, {5, 6}, {6, 7}};
// getElementPerAdjacencyNodeIndex, SimulationControl.cpp:177-216
static const int kQuadFace[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
static const int kHexFace[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {0, 4, 7, 3}, {1, 2, 6, 5}, {2, 3, 7, 6}, {4, 5, 6, 7}};

using Lat = std::array<int, 3>;

// Interior of a lattice square spanned from `o` along `eu`, `ev` (unit lattice steps), parameters in [lo,hi]^2, in gmsh
// order: corners, edge interiors (0-1, 1-2, 2-3, 3-0), then the nested square.
inline void latticeQuad(const Lat& o, const Lat& eu, const Lat& ev, int lo, int hi, std::vector<Lat>& out) {
  if (lo > hi) return;
  auto at = [&](int s, int t) { return Lat{o[0] + s * eu[0] + t * ev[0], o[1] + s * eu[1] + t * ev[1], o[2] + s * eu[2] + t * ev[2]}; };
  if (lo == hi) { out.push_back(at(lo, lo)); return; }
  const int cs[4] = {lo, hi, hi, lo}, ct[4] = {lo, lo, hi, hi};
  for (int c = 0; c < 4; c++) out.push_back(at(cs[c], ct[c]));
  const int len = hi - lo;
  for (int e = 0; e < 4; e++) {
    const int s0 = cs[e], t0 = ct[e], s1 = cs[(e + 1) % 4], t1 = ct[(e + 1) % 4];
    for (int i = 1; i < len; i++) out.push_back(at(s0 + (s1 - s0) / len * i, t0 + (t1 - t0) / len * i));
  }
  latticeQuad(o, eu, ev, lo + 1, hi - 1, out);
}
inline void latticeHex(int lo, int hi, std::vector<Lat>& out) {
  if (lo > hi) return;
  if (lo == hi) { out.push_back({lo, lo, lo}); return; }
  auto corner = [&](int c) { return Lat{kHexCorner[c][0] ? hi : lo, kHexCorner[c][1] ? hi : lo, kHexCorner[c][2] ? hi : lo}; };
  for (int c = 0; c < 8; c++) out.push_back(corner(c));
  const int len = hi - lo;
  for (auto& e : kHexEdge) {
    const Lat a = corner(e[0]), b = corner(e[1]);
    for (int i = 1; i < len; i++) out.push_back({a[0] + (b[0] - a[0]) / len * i, a[1] + (b[1] - a[1]) / len * i, a[2] + (b[2] - a[2]) / len * i});
  }
  for (auto& f : kHexFace) {
    const Lat a = corner(f[0]), b = corner(f[1]), d = corner(f[3]);
    const Lat eu = {(b[0] - a[0]) / len, (b[1] - a[1]) / len, (b[2] - a[2]) / len}, ev = {(d[0] - a[0]) / len, (d[1] - a[1]) / len, (d[2] - a[2]) / len};
    latticeQuad(a, eu, ev, 1, len - 1, out);
  }
  latticeHex(lo + 1, hi - 1, out);
}
// Lattice coordinates (0..g per axis) of the order-g Lagrange nodes of a tensor element in gmsh numbering
// (pinned by getAdjacencyElementViewNodeParentSequence, SimulationControl.cpp:525-887, in tests/).
inline std::vector<Lat> gmshNodeLattice(int type, int g) {
  std::vector<Lat> out;
  if (type == kLine) { out.push_back({0, 0, 0}); out.push_back({g, 0, 0}); for (int i = 1; i < g; i++) out.push_back({i, 0, 0}); }
  else if (type == kQuadrangle) latticeQuad({0, 0, 0}, {1, 0, 0}, {0, 1, 0}, 0, g, out);
  else if (type == kHexahedron) latticeHex(0, g, out);
  else throw std::runtime_error("gmshNodeLattice: tensor elements only");
  return out;
}

// Lobatto index triple of every modal ("H1Legendre{p}") function of a tensor element: vertex, edge, face, bubble
// functions in gmsh entity order, orientation block 0.
inline std::vector<Lat> modalFunctionIndex(int type, int p) {
  std::vector<Lat> out;
  if (type == kLine) { out.push_back({0, 0, 0}); out.push_back({1, 0, 0}); for (int k = 2; k <= p; k++) out.push_back({k, 0, 0}); return out; }
  if (type == kQuadrangle) {
    for (auto& c : kQuadCorner) out.push_back({c[0], c[1], 0});
    for (auto& e : kQuadFace) {
      const int* a = kQuadCorner[e[0]]; const int* b = kQuadCorner[e[1]];
      const int dir = a[0] != b[0] ? 0 : 1;
      for (int k = 2; k <= p; k++) { Lat f = {a[0], a[1], 0}; f[dir] = k; out.push_back(f); }
    }
    for (int i = 2; i <= p; i++) for (int j = 2; j <= p; j++) out.push_back({i, j, 0});
    return out;
  }
  if (type == kHexahedron) {
    for (auto& c : kHexCorner) out.push_back({c[0], c[1], c[2]});
    for (auto& e : kHexEdge) {
      const int* a = kHexCorner[e[0]]; const int* b = kHexCorner[e[1]];
      int dir = 0; for (int d = 0; d < 3; d++) if (a[d] != b[d]) dir = d;
      for (int k = 2; k <= p; k++) { Lat f = {a[0], a[1], a[2]}; f[dir] = k; out.push_back(f); }
    }
    for (auto& fc : kHexFace) {
      const int* a = kHexCorner[fc[0]]; const int* b = kHexCorner[fc[1]]; const int* d = kHexCorner[fc[3]];
      int ds = 0, dt = 0;
      for (int k = 0; k < 3; k++) { if (a[k] != b[k]) ds = k; if (a[k] != d[k]) dt = k; }
      for (int i = 2; i <= p; i++) for (int j = 2; j <= p; j++) { Lat f = {a[0], a[1], a[2]}; f[ds] = i; f[dt] = j; out.push_back(f); }
    }
    for (int i = 2; i <= p; i++) for (int j = 2; j <= p; j++) for (int k = 2; k <= p; k++) out.push_back({i, j, k});
    return out;
  }
  throw std::runtime_error("modalFunctionIndex: tensor elements only");
}

// getAdjacencyElementQuadratureSequence, SimulationControl.cpp:381-523: index of the RIGHT parent's face point that
// coincides with the left parent's face point j.  Line faces: reversal (:387-403).  Quadrangle faces (:456-520) with
// j = n*a + b: rotation 0 -> n*b+a, 1 -> n*(n-1-a)+b, 2 -> n*(n-1-b)+(n-1-a), 3 -> n*a+(n-1-b).
inline std::vector<int> faceSequence(int faceTypeId, int n, int rotation) {
  std::vector<int> s;
  if (faceTypeId == kPoint) return {0};
  if (faceTypeId == kLine) { for (int j = 0; j < n; j++) s.push_back(n - 1 - j); return s; }
  if (faceTypeId == kQuadrangle) {
    for (int a = 0; a < n; a++) for (int b = 0; b < n; b++) {
      if (rotation == 0) s.push_back(n * b + a);
      else if (rotation == 1) s.push_back(n * (n - 1 - a) + b);
      else if (rotation == 2) s.push_back(n * (n - 1 - b) + (n - 1 - a));
      else if (rotation == 3) s.push_back(n * a + (n - 1 - b));
      else throw std::runtime_error("faceSequence: bad rotation");
    }
    return s;
  }
  throw std::runtime_error("faceSequence: unsupported face type");
}

// Everything the tensor-product device path needs for one element type and order.
struct TensorTables {
  int type = 0, D = 0, p = 0, N = 0, NN = 0, NF = 0, NQF = 0;
  std::vector<double> x, w;          // Gauss points / weights (N)
  std::vector<double> Dm;            // Dm[a*N+b] = l_b'(x_a)
  std::vector<double> Lend;          // Lend[s*N+a] = l_a(-1) (s=0), l_a(+1) (s=1)
  std::vector<double> Phi1;          // Phi1[a*N+k] = lobatto_k(x_a)
  std::vector<double> K1;            // K1 = Phi1 Phi1^T (relative-error operator, TimeIntegration.cpp:279-298)
  std::vector<double> wq;            // volume weights per node (NN), node q = sum_d i_d N^(D-1-d)  (xi slowest)
  std::vector<double> wf;            // face weights per face point (NQF), first face coordinate slowest
  std::vector<double> Phi;           // Phi[q*NN+b]: modal function b at node q (modal_value_, BasisFunction.cpp:199-229)
  std::vector<double> PhiInv;        // inverse of Phi (nodal values -> modal coefficients)
  std::vector<int> faceDir, faceSide;  // per local face: normal axis, 0 = low side / 1 = high side
  std::vector<int> faceBase;         // [f*NQF+j]: node index of the face point's line with i_normal = 0
  std::vector<int> nodeFacePt;       // [f*NN+q]: face point j of face f whose normal line holds node q
  std::vector<double> faceTan;       // [f][a][k]: d(xi_k)/d(s_a) of the face's corner map (a < D-1)
  std::vector<Lat> modalIdx;
};

inline void invertDense(std::vector<double>& A, int n) {  // row-major Gauss-Jordan, partial pivoting
  std::vector<double> I((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) I[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++) if (std::fabs(A[(size_t)r * n + c]) > std::fabs(A[(size_t)piv * n + c])) piv = r;
    if (A[(size_t)piv * n + c] == 0.0) throw std::runtime_error("invertDense: singular matrix");
    if (piv != c) for (int k = 0; k < n; k++) { std::swap(A[(size_t)c * n + k], A[(size_t)piv * n + k]); std::swap(I[(size_t)c * n + k], I[(size_t)piv * n + k]); }
    const double d = 1.0 / A[(size_t)c * n + c];
    for (int k = 0; k < n; k++) { A[(size_t)c * n + k] *= d; I[(size_t)c * n + k] *= d; }
    for (int r = 0; r < n; r++) if (r != c) {
      const double f = A[(size_t)r * n + c];
      if (f != 0.0) for (int k = 0; k < n; k++) { A[(size_t)r * n + k] -= f * A[(size_t)c * n + k]; I[(size_t)r * n + k] -= f * I[(size_t)c * n + k]; }
    }
  }
  A.swap(I);
}

inline TensorTables buildTensorTables(int type, int p) {
  TensorTables T;
  T.type = type; T.D = elemDim(type); T.p = p; T.N = p + 1;
  const int D = T.D, N = T.N;
  T.NN = 1; for (int d = 0; d < D; d++) T.NN *= N;
  T.NQF = T.NN / N; T.NF = 2 * D;
  gaussRule(N, T.x, T.w);
  T.Dm.assign((size_t)N * N, 0.0); T.Lend.assign(2 * N, 0.0); T.Phi1.assign((size_t)N * N, 0.0); T.K1.assign((size_t)N * N, 0.0);
  std::vector<double> v, d;
  for (int a = 0; a < N; a++) { lagrangeAt(T.x, T.x[a], v, d); for (int b = 0; b < N; b++) T.Dm[a * N + b] = d[b]; }
  lagrangeAt(T.x, -1.0, v, d); for (int a = 0; a < N; a++) T.Lend[a] = v[a];
  lagrangeAt(T.x, 1.0, v, d); for (int a = 0; a < N; a++) T.Lend[N + a] = v[a];
  for (int a = 0; a < N; a++) for (int k = 0; k < N; k++) T.Phi1[a * N + k] = lobattoShape(k, T.x[a]);
  for (int a = 0; a < N; a++) for (int b = 0; b < N; b++) { double s = 0; for (int k = 0; k < N; k++) s += T.Phi1[a * N + k] * T.Phi1[b * N + k]; T.K1[a * N + b] = s; }
  auto idxOf = [&](int q, int dd) { int s = 1; for (int k = D - 1; k > dd; k--) s *= N; return (q / s) % N; };
  T.wq.assign(T.NN, 1.0);
  for (int q = 0; q < T.NN; q++) for (int dd = 0; dd < D; dd++) T.wq[q] *= T.w[idxOf(q, dd)];
  T.wf.assign(T.NQF, 1.0);
  for (int j = 0; j < T.NQF; j++) { int r = j; for (int a = D - 2; a >= 0; a--) { T.wf[j] *= T.w[r % N]; r /= N; } }
  T.modalIdx = modalFunctionIndex(type, p);
  T.Phi.assign((size_t)T.NN * T.NN, 0.0);
  for (int q = 0; q < T.NN; q++) for (int b = 0; b < T.NN; b++) {
    double s = 1.0; for (int dd = 0; dd < D; dd++) s *= T.Phi1[idxOf(q, dd) * N + T.modalIdx[b][dd]];
    T.Phi[(size_t)q * T.NN + b] = s;
  }
  T.PhiInv = T.Phi; invertDense(T.PhiInv, T.NN);
  // faces: corner map of the face (P1 Lagrange through the face corners, BasisFunction.cpp:76-111) evaluated at the face
  // Gauss points; the tangential abscissae coincide with volume Gauss abscissae because both rules have p+1 points per
  // direction (orders 2p and 2p+1, SimulationControl.cpp:275-283).
  T.faceDir.assign(T.NF, 0); T.faceSide.assign(T.NF, 0); T.faceBase.assign((size_t)T.NF * T.NQF, 0);
  T.nodeFacePt.assign((size_t)T.NF * T.NN, 0); T.faceTan.assign((size_t)T.NF * 2 * 3, 0.0);
  std::vector<int> stride(D); for (int dd = 0; dd < D; dd++) { int s = 1; for (int k = D - 1; k > dd; k--) s *= N; stride[dd] = s; }
  if (D == 1) {  // line element: the two faces are its end points (face f = node f, SimulationControl.cpp:177-216), one face "point" each
    for (int f = 0; f < 2; f++) {
      T.faceDir[f] = 0; T.faceSide[f] = f; T.faceBase[f] = 0;
      for (int a = 0; a < N; a++) T.nodeFacePt[(size_t)f * T.NN + a] = 0;
    }
    return T;
  }
  for (int f = 0; f < T.NF; f++) {
    double c[4][3] = {{0}}; int nc = D == 2 ? 2 : 4;
    for (int m = 0; m < nc; m++) {
      const int* bits = D == 2 ? kQuadCorner[kQuadFace[f][m]] : kHexCorner[kHexFace[f][m]];
      for (int k = 0; k < 3; k++) c[m][k] = k < D ? (bits[k] ? 1.0 : -1.0) : 0.0;
    }
    int dn = -1;
    for (int k = 0; k < D; k++) { bool same = true; for (int m = 1; m < nc; m++) if (c[m][k] != c[0][k]) same = false; if (same) dn = k; }
    T.faceDir[f] = dn; T.faceSide[f] = c[0][dn] > 0 ? 1 : 0;
    for (int k = 0; k < D; k++) {
      T.faceTan[((size_t)f * 2 + 0) * 3 + k] = 0.5 * (c[1][k] - c[0][k]);
      if (D == 3) T.faceTan[((size_t)f * 2 + 1) * 3 + k] = 0.5 * (c[3][k] - c[0][k]);
    }
    for (int j = 0; j < T.NQF; j++) {
      double xi[3] = {0, 0, 0};
      if (D == 2) { const double s = T.x[j]; for (int k = 0; k < 2; k++) xi[k] = 0.5 * (1 - s) * c[0][k] + 0.5 * (1 + s) * c[1][k]; }
      else {
        const double s = T.x[j / N], t = T.x[j % N];
        for (int k = 0; k < 3; k++) xi[k] = 0.25 * ((1 - s) * (1 - t) * c[0][k] + (1 + s) * (1 - t) * c[1][k] + (1 + s) * (1 + t) * c[2][k] + (1 - s) * (1 + t) * c[3][k]);
      }
      int base = 0;
      for (int k = 0; k < D; k++) if (k != dn) {
        int hit = -1; for (int a = 0; a < N; a++) if (std::fabs(T.x[a] - xi[k]) < 1e-13) hit = a;
        if (hit < 0) throw std::runtime_error("buildTensorTables: face point is not on the volume Gauss lattice");
        base += hit * stride[k];
      }
      T.faceBase[(size_t)f * T.NQF + j] = base;
      for (int a = 0; a < N; a++) T.nodeFacePt[(size_t)f * T.NN + base + a * stride[dn]] = j;
    }
  }
  return T;
}

}  // namespace sdg
