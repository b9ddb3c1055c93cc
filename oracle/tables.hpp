// oracle/tables.hpp — TEST INFRASTRUCTURE ONLY (CPU oracle). Never included, linked or called by the product path.
//
// Restates, without gmsh/Eigen, the reference-element tables that SubrosaDG obtains from the Gmsh 4.13.1 library
// at run time (src/Mesh/Quadrature.cpp:27-34, src/Mesh/BasisFunction.cpp:31-74,136-230) and the integer tables it
// embeds (src/Solver/SimulationControl.cpp:26-523).  PARITY UNPINNED with respect to the real gmsh binary: the
// reference ships no golden vectors and gmsh is not in this container; what IS pinned (tests/test_oracle_tables.py)
// are the integer tables of SimulationControl.cpp (quadrature counts, face->corner maps, face point permutations,
// high-order node numbering on faces) plus exactness/interpolation properties of every rule and basis below.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <stdexcept>
#include <vector>

namespace orc {

// ElementEnum values, src/Utils/Enum.cpp:28-36
enum ElemType { kPoint = 0, kLine = 1, kTriangle = 2, kQuadrangle = 3, kTetrahedron = 4, kPyramid = 5, kHexahedron = 6 };

inline int elemDim(int t) {
  switch (t) {
    case kPoint: return 0;
    case kLine: return 1;
    case kTriangle: case kQuadrangle: return 2;
    default: return 3;
  }
}

// ---- forward-mode dual number with 3 partials (exact gradients of every basis below) -------------------------------
struct Dual {
  double v;
  double d[3];
  Dual(double a = 0.0) : v(a), d{0, 0, 0} {}
  static Dual var(double a, int k) { Dual r(a); r.d[k] = 1.0; return r; }
};
inline Dual operator+(const Dual& a, const Dual& b) { Dual r(a.v + b.v); for (int k = 0; k < 3; k++) r.d[k] = a.d[k] + b.d[k]; return r; }
inline Dual operator-(const Dual& a, const Dual& b) { Dual r(a.v - b.v); for (int k = 0; k < 3; k++) r.d[k] = a.d[k] - b.d[k]; return r; }
inline Dual operator*(const Dual& a, const Dual& b) { Dual r(a.v * b.v); for (int k = 0; k < 3; k++) r.d[k] = a.d[k] * b.v + a.v * b.d[k]; return r; }
inline Dual operator*(double a, const Dual& b) { Dual r(a * b.v); for (int k = 0; k < 3; k++) r.d[k] = a * b.d[k]; return r; }
inline Dual operator/(const Dual& a, double b) { Dual r(a.v / b); for (int k = 0; k < 3; k++) r.d[k] = a.d[k] / b; return r; }

// ---- Legendre polynomials (value for any T supporting + - *) -------------------------------------------------------
template <typename T>
inline T legendre(int n, const T& x) {
  if (n == 0) return T(1.0);
  if (n == 1) return x;
  T pm = T(1.0), p = x;
  for (int k = 2; k <= n; k++) {
    T pn = ((2.0 * k - 1.0) * (x * p) - (k - 1.0) * pm) / double(k);
    pm = p; p = pn;
  }
  return p;
}
template <typename T>
inline T legendreDerivative(int n, const T& x) {  // L'_n via L'_n = sum_{k=n-1,n-3,..} (2k+1) L_k
  T s = T(0.0);
  for (int k = n - 1; k >= 0; k -= 2) s = s + (2.0 * k + 1.0) * legendre(k, x);
  return s;
}

// Gauss-Legendre rule with n points on [-1,1], ascending abscissae ("Gauss{o}" on a line, n = o/2+1;
// count pinned by kLineQuadratureNumber, SimulationControl.cpp:268).
inline void gaussLegendre(int n, std::vector<double>& x, std::vector<double>& w) {
  x.assign(n, 0.0); w.assign(n, 0.0);
  for (int i = 0; i < n; i++) {
    double t = -std::cos(M_PI * (i + 0.75) / (n + 0.5));
    for (int it = 0; it < 100; it++) {
      double p = legendre(n, t), dp = legendreDerivative(n, t);
      double dt = p / dp; t -= dt;
      if (std::fabs(dt) < 1e-16) break;
    }
    double dp = legendreDerivative(n, t);
    x[i] = t; w[i] = 2.0 / ((1.0 - t * t) * dp * dp);
  }
  for (int i = 0; i < n / 2; i++) {  // enforce exact symmetry (gmsh tables are symmetric)
    double a = 0.5 * (x[n - 1 - i] - x[i]); x[i] = -a; x[n - 1 - i] = a;
    double b = 0.5 * (w[i] + w[n - 1 - i]); w[i] = b; w[n - 1 - i] = b;
  }
  if (n % 2) x[n / 2] = 0.0;
}

// Quadrature point counts, SimulationControl.cpp:268-273 (index = quadrature order).
static const int kLineQuadratureNumber[12] = {1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6};
static const int kTriangleQuadratureNumber[12] = {1, 1, 3, 4, 6, 7, 12, 13, 16, 19, 25, 27};
static const int kQuadrangleQuadratureNumber[12] = {1, 3, 7, 4, 9, 9, 16, 16, 25, 25, 36, 36};
static const int kHexahedronQuadratureNumber[12] = {1, 6, 8, 8, 27, 27, 64, 64, 125, 125, 216, 216};

struct Quadrature {
  int n = 0;
  std::vector<double> pts;  // n x 3 (u,v,w), unused coordinates 0
  std::vector<double> wts;  // n
};

// Symmetric triangle rules on the unit triangle (0,0),(1,0),(0,1) (Dunavant 1985, degrees 2, 4, 6; the point
// counts 3/6/12 match kTriangleQuadratureNumber[2,4,6]).  Weights sum to the reference measure 1/2
// (getElementMeasure, SimulationControl.cpp:226-228).
inline Quadrature triangleRule(int order) {
  Quadrature q;
  auto add3 = [&](double w, double a, double b) {  // barycentric (a,b,b) and permutations
    const double P[3][3] = {{a, b, b}, {b, a, b}, {b, b, a}};
    for (auto& p : P) { q.pts.insert(q.pts.end(), {p[1], p[2], 0.0}); q.wts.push_back(0.5 * w); }
  };
  auto add6 = [&](double w, double a, double b, double c) {
    const double P[6][3] = {{a, b, c}, {a, c, b}, {b, a, c}, {b, c, a}, {c, a, b}, {c, b, a}};
    for (auto& p : P) { q.pts.insert(q.pts.end(), {p[1], p[2], 0.0}); q.wts.push_back(0.5 * w); }
  };
  if (order <= 2) {
    add3(1.0 / 3.0, 2.0 / 3.0, 1.0 / 6.0);
  } else if (order <= 4) {
    add3(0.223381589678011, 0.108103018168070, 0.445948490915965);
    add3(0.109951743655322, 0.816847572980459, 0.091576213509771);
  } else if (order <= 6) {
    add3(0.116786275726379, 0.501426509658179, 0.249286745170910);
    add3(0.050844906370207, 0.873821971016996, 0.063089014491502);
    add6(0.082851075618374, 0.053145049844817, 0.310352451033784, 0.636502499121399);
  } else {
    throw std::runtime_error("oracle: triangle quadrature order > 6 not tabulated (triangle P <= 3 only)");
  }
  q.n = (int)q.wts.size();
  // renormalise the 15-digit literature weights so that sum(w) == 1/2 to round-off
  double s = 0; for (double w : q.wts) s += w;
  for (double& w : q.wts) w *= 0.5 / s;
  return q;
}

// "Gauss{order}" for an element type.  Tensor rules: first coordinate is the SLOWEST index (pinned by the face
// permutation tables SimulationControl.cpp:456-520).
inline Quadrature makeQuadrature(int type, int order) {
  Quadrature q;
  if (type == kPoint) { q.n = 1; q.pts = {0, 0, 0}; q.wts = {1.0}; return q; }
  if (type == kTriangle) return triangleRule(order);
  if (type == kLine || type == kQuadrangle || type == kHexahedron) {
    const int n = order / 2 + 1;
    std::vector<double> x, w; gaussLegendre(n, x, w);
    const int D = elemDim(type);
    // Quadrangle order 2 (P1 quadrangles: thermalcavity_2d / naca0010_2d / shearlayer_2d of the reference's examples): Gmsh answers "Gauss2"
    // with a SEVEN-point rule (kQuadrangleQuadratureNumber[2] = 7, SimulationControl.cpp:270; the reference's fixed-size tables are sized by it) --
    // Radon's degree-5 formula for the square (Stroud C2 5-1): the centre with weight 8/7, (0, +-sqrt(14/15)) with 20/63 and
    // (+-sqrt(3/5), +-sqrt(1/3)) with 5/9.  Values from the formula; the point ORDER follows Gmsh's table as far as it is remembered
    // (it only permutes the quadrature arrays).  Orders 1 (3 / 6 points, quadrangle / hexahedron) are never requested (order = 2 P >= 2).
    if (type == kQuadrangle && order == 2) {
      const double r = std::sqrt(14.0 / 15.0), a = std::sqrt(3.0 / 5.0), b = std::sqrt(1.0 / 3.0);
      const double P7[7][2] = {{0, 0}, {0, r}, {0, -r}, {a, b}, {a, -b}, {-a, b}, {-a, -b}};
      const double W7[7] = {8.0 / 7.0, 20.0 / 63.0, 20.0 / 63.0, 5.0 / 9.0, 5.0 / 9.0, 5.0 / 9.0, 5.0 / 9.0};
      q.n = 7;
      for (int i = 0; i < 7; i++) { q.pts.insert(q.pts.end(), {P7[i][0], P7[i][1], 0}); q.wts.push_back(W7[i]); }
      return q;
    }
    if (D == 1) { q.n = n; for (int i = 0; i < n; i++) { q.pts.insert(q.pts.end(), {x[i], 0, 0}); q.wts.push_back(w[i]); } }
    if (D == 2) { q.n = n * n; for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { q.pts.insert(q.pts.end(), {x[i], x[j], 0}); q.wts.push_back(w[i] * w[j]); } }
    if (D == 3) { q.n = n * n * n; for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) for (int k = 0; k < n; k++) { q.pts.insert(q.pts.end(), {x[i], x[j], x[k]}); q.wts.push_back(w[i] * w[j] * w[k]); } }
    return q;
  }
  throw std::runtime_error("oracle: element type not supported (tetrahedron/pyramid)");
}

// ---- element topology tables ---------------------------------------------------------------------------------------
inline int numCorners(int t) { static const int n[7] = {1, 2, 3, 4, 4, 5, 8}; return n[t]; }
inline int numFaces(int t) { static const int n[7] = {0, 2, 3, 4, 4, 5, 6}; return n[t]; }  // getElementAdjacencyNumber :99-122
inline int faceType(int t) { return t == kLine ? kPoint : (t == kTriangle || t == kQuadrangle) ? kLine : t == kHexahedron ? kQuadrangle : kTriangle; }
inline int numNodes(int t, int p) {  // getElementNodeNumber :74-97
  switch (t) {
    case kPoint: return 1;
    case kLine: return p + 1;
    case kTriangle: return (p + 1) * (p + 2) / 2;
    case kQuadrangle: return (p + 1) * (p + 1);
    case kHexahedron: return (p + 1) * (p + 1) * (p + 1);
  }
  throw std::runtime_error("oracle: numNodes unsupported type");
}
// getElementPerAdjacencyNodeIndex, SimulationControl.cpp:177-216 (face -> corner ids, in face-local order)
inline std::vector<int> faceCorners(int t, int f) {
  static const int L[2][1] = {{0}, {1}};
  static const int T[3][2] = {{0, 1}, {1, 2}, {2, 0}};
  static const int Q[4][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}};
  static const int H[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {0, 4, 7, 3}, {1, 2, 6, 5}, {2, 3, 7, 6}, {4, 5, 6, 7}};
  switch (t) {
    case kLine: return {L[f][0]};
    case kTriangle: return {T[f][0], T[f][1]};
    case kQuadrangle: return {Q[f][0], Q[f][1]};
    case kHexahedron: return {H[f][0], H[f][1], H[f][2], H[f][3]};
  }
  throw std::runtime_error("oracle: faceCorners unsupported type");
}
// reference coordinates of the corner vertices (gmsh reference elements; measures pinned by :218-241)
inline std::array<double, 3> cornerCoord(int t, int c) {
  static const double L[2][3] = {{-1, 0, 0}, {1, 0, 0}};
  static const double T[3][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}};
  static const double Q[4][3] = {{-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}};
  static const double H[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
  const double* p = t == kPoint ? L[0] : t == kLine ? L[c] : t == kTriangle ? T[c] : t == kQuadrangle ? Q[c] : H[c];
  if (t == kPoint) return {0, 0, 0};
  return {p[0], p[1], p[2]};
}

// ---- gmsh high-order node numbering --------------------------------------------------------------------------------
// Reference coordinates of the order-g Lagrange nodes in gmsh order: corners, then edge-interior nodes edge by edge
// (directed first->second vertex), then face-interior nodes face by face (recursively a lower-order face element in
// the face's own corner frame), then volume-interior nodes (recursively).  Pinned against
// getAdjacencyElementViewNodeParentSequence (SimulationControl.cpp:525-887) in tests.
inline void lineNodes(const std::array<double, 3>& a, const std::array<double, 3>& b, int g, bool withEnds,
                      std::vector<std::array<double, 3>>& out) {
  if (withEnds) { out.push_back(a); out.push_back(b); }
  for (int i = 1; i < g; i++) {
    double s = double(i) / g;
    out.push_back({a[0] + s * (b[0] - a[0]), a[1] + s * (b[1] - a[1]), a[2] + s * (b[2] - a[2])});
  }
}
inline void quadNodes(const std::array<std::array<double, 3>, 4>& c, int g, std::vector<std::array<double, 3>>& out) {
  if (g == 0) { out.push_back({0.25 * (c[0][0] + c[1][0] + c[2][0] + c[3][0]), 0.25 * (c[0][1] + c[1][1] + c[2][1] + c[3][1]), 0.25 * (c[0][2] + c[1][2] + c[2][2] + c[3][2])}); return; }
  for (int i = 0; i < 4; i++) out.push_back(c[i]);
  if (g == 1) return;
  for (int e = 0; e < 4; e++) lineNodes(c[e], c[(e + 1) % 4], g, false, out);
  // interior: quad of order g-2 whose corners are the bilinear images of the (1/g) inset lattice corners
  auto bil = [&](double s, double t) {  // s,t in [0,1] along c0->c1 and c0->c3
    std::array<double, 3> r;
    for (int k = 0; k < 3; k++) r[k] = (1 - s) * (1 - t) * c[0][k] + s * (1 - t) * c[1][k] + s * t * c[2][k] + (1 - s) * t * c[3][k];
    return r;
  };
  double h = 1.0 / g;
  std::array<std::array<double, 3>, 4> ci = {bil(h, h), bil(1 - h, h), bil(1 - h, 1 - h), bil(h, 1 - h)};
  quadNodes(ci, g - 2, out);
}
inline void triNodes(const std::array<std::array<double, 3>, 3>& c, int g, std::vector<std::array<double, 3>>& out) {
  if (g == 0) { out.push_back({(c[0][0] + c[1][0] + c[2][0]) / 3, (c[0][1] + c[1][1] + c[2][1]) / 3, (c[0][2] + c[1][2] + c[2][2]) / 3}); return; }
  for (int i = 0; i < 3; i++) out.push_back(c[i]);
  if (g == 1) return;
  for (int e = 0; e < 3; e++) lineNodes(c[e], c[(e + 1) % 3], g, false, out);
  if (g < 3) return;
  auto bary = [&](double l0, double l1, double l2) {
    std::array<double, 3> r; for (int k = 0; k < 3; k++) r[k] = l0 * c[0][k] + l1 * c[1][k] + l2 * c[2][k]; return r;
  };
  double h = 1.0 / g;
  std::array<std::array<double, 3>, 3> ci = {bary(1 - 2 * h, h, h), bary(h, 1 - 2 * h, h), bary(h, h, 1 - 2 * h)};
  triNodes(ci, g - 3, out);
}
inline void hexNodes(const std::array<std::array<double, 3>, 8>& c, int g, std::vector<std::array<double, 3>>& out) {
  if (g == 0) { std::array<double, 3> m{0, 0, 0}; for (auto& p : c) for (int k = 0; k < 3; k++) m[k] += p[k] / 8; out.push_back(m); return; }
  for (int i = 0; i < 8; i++) out.push_back(c[i]);
  if (g == 1) return;
  static const int E[12][2] = {{0, 1}, {0, 3}, {0, 4}, {1, 2}, {1, 5}, {2, 3}, {2, 6}, {3, 7}, {4, 5}, {4, 7}, {5, 6}, {6, 7}};
  for (auto& e : E) lineNodes(c[e[0]], c[e[1]], g, false, out);
  for (int f = 0; f < 6; f++) {
    std::vector<int> fc = faceCorners(kHexahedron, f);
    std::array<std::array<double, 3>, 4> cf = {c[fc[0]], c[fc[1]], c[fc[2]], c[fc[3]]};
    std::vector<std::array<double, 3>> tmp; quadNodes(cf, g, tmp);
    out.insert(out.end(), tmp.begin() + 4 + 4 * (g - 1), tmp.end());  // face-interior part only
  }
  auto tri = [&](double s, double t, double u) {
    std::array<double, 3> r;
    for (int k = 0; k < 3; k++)
      r[k] = (1 - u) * ((1 - s) * (1 - t) * c[0][k] + s * (1 - t) * c[1][k] + s * t * c[2][k] + (1 - s) * t * c[3][k]) +
             u * ((1 - s) * (1 - t) * c[4][k] + s * (1 - t) * c[5][k] + s * t * c[6][k] + (1 - s) * t * c[7][k]);
    return r;
  };
  double h = 1.0 / g;
  std::array<std::array<double, 3>, 8> ci = {tri(h, h, h), tri(1 - h, h, h), tri(1 - h, 1 - h, h), tri(h, 1 - h, h),
                                             tri(h, h, 1 - h), tri(1 - h, h, 1 - h), tri(1 - h, 1 - h, 1 - h), tri(h, 1 - h, 1 - h)};
  hexNodes(ci, g - 2, out);
}
inline std::vector<std::array<double, 3>> referenceNodes(int t, int g) {
  std::vector<std::array<double, 3>> out;
  switch (t) {
    case kPoint: out.push_back({0, 0, 0}); break;
    case kLine: lineNodes(cornerCoord(t, 0), cornerCoord(t, 1), g, true, out); break;
    case kTriangle: triNodes({cornerCoord(t, 0), cornerCoord(t, 1), cornerCoord(t, 2)}, g, out); break;
    case kQuadrangle: quadNodes({cornerCoord(t, 0), cornerCoord(t, 1), cornerCoord(t, 2), cornerCoord(t, 3)}, g, out); break;
    case kHexahedron: {
      std::array<std::array<double, 3>, 8> c; for (int i = 0; i < 8; i++) c[i] = cornerCoord(t, i);
      hexNodes(c, g, out); break;
    }
    default: throw std::runtime_error("oracle: referenceNodes unsupported type");
  }
  return out;
}

// ---- Lagrange (geometry / "nodal") basis of order g on the gmsh-ordered nodes --------------------------------------
// getBasisFunctions("Lagrange{g}") restated.  Tensor elements: products of 1-D Lagrange polynomials on equispaced
// nodes (exact, well conditioned); triangle: inverse of a scaled-monomial Vandermonde matrix.
struct LagrangeBasis {
  int type = 0, g = 1, n = 0;
  std::vector<std::array<double, 3>> nodes;
  std::vector<std::array<int, 3>> idx;  // tensor elements: 1-D node index per direction
  std::vector<double> coef;             // triangle: n x n monomial coefficients
  std::vector<std::array<int, 2>> mono;

  LagrangeBasis() {}
  LagrangeBasis(int type_, int g_) : type(type_), g(g_) {
    nodes = referenceNodes(type, g); n = (int)nodes.size();
    const int D = elemDim(type);
    if (type == kLine || type == kQuadrangle || type == kHexahedron) {
      idx.resize(n);
      for (int i = 0; i < n; i++) for (int k = 0; k < 3; k++) idx[i][k] = k < D ? (int)std::lround((nodes[i][k] + 1.0) * 0.5 * g) : 0;
    } else if (type == kTriangle) {
      for (int a = 0; a <= g; a++) for (int b = 0; a + b <= g; b++) mono.push_back({a, b});
      std::vector<double> V(n * n), I(n * n, 0.0);
      for (int i = 0; i < n; i++) for (int m = 0; m < n; m++) V[i * n + m] = std::pow(nodes[i][0], mono[m][0]) * std::pow(nodes[i][1], mono[m][1]);
      for (int i = 0; i < n; i++) I[i * n + i] = 1.0;
      // solve V * C = I  (C[m][j] = coefficient of monomial m in basis j), Gauss-Jordan with partial pivoting
      for (int c = 0; c < n; c++) {
        int piv = c; for (int r = c + 1; r < n; r++) if (std::fabs(V[r * n + c]) > std::fabs(V[piv * n + c])) piv = r;
        for (int k = 0; k < n; k++) { std::swap(V[c * n + k], V[piv * n + k]); std::swap(I[c * n + k], I[piv * n + k]); }
        double d = V[c * n + c];
        for (int k = 0; k < n; k++) { V[c * n + k] /= d; I[c * n + k] /= d; }
        for (int r = 0; r < n; r++) if (r != c) { double f = V[r * n + c]; if (f != 0) for (int k = 0; k < n; k++) { V[r * n + k] -= f * V[c * n + k]; I[r * n + k] -= f * I[c * n + k]; } }
      }
      coef = I;
    }
  }
  static Dual lag1d(int g, int j, const Dual& x) {  // 1-D Lagrange polynomial j on g+1 equispaced nodes of [-1,1]
    Dual r(1.0);
    for (int m = 0; m <= g; m++) if (m != j) { double xm = -1.0 + 2.0 * m / g, xj = -1.0 + 2.0 * j / g; r = r * ((x - Dual(xm)) / (xj - xm)); }
    return r;
  }
  // values val[n], gradients grad[n][3] at reference point (u,v,w)
  void eval(double u, double v, double w, std::vector<double>& val, std::vector<std::array<double, 3>>& grad) const {
    val.assign(n, 0.0); grad.assign(n, {0, 0, 0});
    Dual U = Dual::var(u, 0), V = Dual::var(v, 1), W = Dual::var(w, 2);
    const int D = elemDim(type);
    if (type == kPoint) { val[0] = 1.0; return; }
    if (type == kTriangle) {
      std::vector<Dual> m(n);
      for (int k = 0; k < n; k++) { Dual r(1.0); for (int a = 0; a < mono[k][0]; a++) r = r * U; for (int b = 0; b < mono[k][1]; b++) r = r * V; m[k] = r; }
      for (int j = 0; j < n; j++) { Dual s(0.0); for (int k = 0; k < n; k++) s = s + coef[k * n + j] * m[k]; val[j] = s.v; grad[j] = {s.d[0], s.d[1], 0.0}; }
      return;
    }
    for (int j = 0; j < n; j++) {
      Dual r = lag1d(g, idx[j][0], U);
      if (D > 1) r = r * lag1d(g, idx[j][1], V);
      if (D > 2) r = r * lag1d(g, idx[j][2], W);
      val[j] = r.v; grad[j] = {r.d[0], r.d[1], r.d[2]};
    }
  }
};

// ---- modal "H1Legendre" (Lobatto, Solin) basis ----------------------------------------------------------------------
// getBasisFunctions("H1Legendre{p}") restated from docs/develop-note/develop-note.tex:272-322 and SURVEY.md App. B:
// function order = vertex, edge, face, bubble; orientation block 0 (no sign flips).  All physics is invariant under a
// change of basis of the same polynomial space; only raw modal coefficients depend on this ordering (unverifiable
// without gmsh).
inline Dual lobatto(int k, const Dual& x) {
  if (k == 0) return (Dual(1.0) - x) / 2.0;
  if (k == 1) return (Dual(1.0) + x) / 2.0;
  return (legendre(k, x) - legendre(k - 2, x)) / std::sqrt(2.0 * (2.0 * k - 1.0));
}
inline Dual lobattoKernel(int k, const Dual& x) {  // phi_k: l_{k+2}(x) = l_0 l_1 phi_k(x)
  const int kk = k + 2;
  return (-4.0 * (2.0 * kk - 1.0) / (kk * (kk - 1.0) * std::sqrt(2.0 * (2.0 * kk - 1.0)))) * legendreDerivative(kk - 1, x);
}

struct ModalBasis {
  int type = 0, p = 1, n = 0;
  std::vector<std::array<int, 3>> idx;  // tensor elements: Lobatto index per direction for every function
  ModalBasis() {}
  ModalBasis(int type_, int p_) : type(type_), p(p_) {
    if (type == kLine) {
      idx.push_back({0, 0, 0}); idx.push_back({1, 0, 0});
      for (int k = 2; k <= p; k++) idx.push_back({k, 0, 0});
    } else if (type == kQuadrangle) {
      const int V[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};
      for (auto& v : V) idx.push_back({v[0], v[1], 0});
      // edges (0,1),(1,2),(2,3),(3,0): varying direction gets k, fixed direction keeps the vertex index
      for (int e = 0; e < 4; e++) {
        const int* a = V[e]; const int* b = V[(e + 1) % 4];
        int dir = a[0] != b[0] ? 0 : 1;
        for (int k = 2; k <= p; k++) { std::array<int, 3> f = {a[0], a[1], 0}; f[dir] = k; idx.push_back(f); }
      }
      for (int i = 2; i <= p; i++) for (int j = 2; j <= p; j++) idx.push_back({i, j, 0});
    } else if (type == kHexahedron) {
      const int V[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
      for (auto& v : V) idx.push_back({v[0], v[1], v[2]});
      static const int E[12][2] = {{0, 1}, {0, 3}, {0, 4}, {1, 2}, {1, 5}, {2, 3}, {2, 6}, {3, 7}, {4, 5}, {4, 7}, {5, 6}, {6, 7}};
      for (auto& e : E) {
        int dir = 0; for (int d = 0; d < 3; d++) if (V[e[0]][d] != V[e[1]][d]) dir = d;
        for (int k = 2; k <= p; k++) { std::array<int, 3> f = {V[e[0]][0], V[e[0]][1], V[e[0]][2]}; f[dir] = k; idx.push_back(f); }
      }
      for (int fc = 0; fc < 6; fc++) {
        std::vector<int> c = faceCorners(kHexahedron, fc);
        int ds = 0, dt = 0;
        for (int d = 0; d < 3; d++) { if (V[c[0]][d] != V[c[1]][d]) ds = d; if (V[c[0]][d] != V[c[3]][d]) dt = d; }
        for (int i = 2; i <= p; i++) for (int j = 2; j <= p; j++) {
          std::array<int, 3> f = {V[c[0]][0], V[c[0]][1], V[c[0]][2]}; f[ds] = i; f[dt] = j; idx.push_back(f);
        }
      }
      for (int i = 2; i <= p; i++) for (int j = 2; j <= p; j++) for (int k = 2; k <= p; k++) idx.push_back({i, j, k});
    } else if (type == kTriangle) {
      // sizes only; evaluation below
    } else {
      throw std::runtime_error("oracle: ModalBasis unsupported type");
    }
    n = numNodes(type, p);
    if (type != kTriangle && (int)idx.size() != n) throw std::runtime_error("oracle: modal basis size mismatch");
  }
  void eval(double u, double v, double w, std::vector<double>& val, std::vector<std::array<double, 3>>& grad) const {
    val.assign(n, 0.0); grad.assign(n, {0, 0, 0});
    Dual U = Dual::var(u, 0), V = Dual::var(v, 1), W = Dual::var(w, 2);
    if (type == kTriangle) {
      Dual lam[3] = {Dual(1.0) - U - V, U, V};
      std::vector<Dual> f;
      for (int i = 0; i < 3; i++) f.push_back(lam[i]);
      for (int e = 0; e < 3; e++) {
        const Dual &a = lam[e], &b = lam[(e + 1) % 3];
        for (int k = 2; k <= p; k++) f.push_back(a * b * lobattoKernel(k - 2, b - a));
      }
      for (int n1 = 1; n1 <= p - 2; n1++) for (int n2 = 1; n1 + n2 <= p - 1; n2++)
        f.push_back(lam[0] * lam[1] * lam[2] * lobattoKernel(n1 - 1, lam[1] - lam[0]) * lobattoKernel(n2 - 1, lam[0] - lam[2]));
      if ((int)f.size() != n) throw std::runtime_error("oracle: triangle modal basis size mismatch");
      for (int j = 0; j < n; j++) { val[j] = f[j].v; grad[j] = {f[j].d[0], f[j].d[1], 0.0}; }
      return;
    }
    const int D = elemDim(type);
    for (int j = 0; j < n; j++) {
      Dual r = lobatto(idx[j][0], U);
      if (D > 1) r = r * lobatto(idx[j][1], V);
      if (D > 2) r = r * lobatto(idx[j][2], W);
      val[j] = r.v; grad[j] = {r.d[0], r.d[1], r.d[2]};
    }
  }
};

// getAdjacencyElementQuadratureSequence, SimulationControl.cpp:381-523: right-side face point index for left point j.
// Line: reversal (:387-403).  Quadrangle (:456-520), j = n*a + b: rot0 -> n*b+a, rot1 -> n*(n-1-a)+b,
// rot2 -> n*(n-1-b)+(n-1-a), rot3 -> n*a+(n-1-b).  Literal reference tables are checked in tests/golden.
inline std::vector<int> faceQuadratureSequence(int ftype, int p, int rotation) {
  std::vector<int> s;
  const int n = p + 1;
  if (ftype == kPoint) return {0};
  if (ftype == kLine) { for (int j = 0; j < n; j++) s.push_back(n - 1 - j); return s; }
  if (ftype == kQuadrangle) {
    for (int a = 0; a < n; a++) for (int b = 0; b < n; b++) {
      switch (rotation) {
        case 0: s.push_back(n * b + a); break;
        case 1: s.push_back(n * (n - 1 - a) + b); break;
        case 2: s.push_back(n * (n - 1 - b) + (n - 1 - a)); break;
        case 3: s.push_back(n * a + (n - 1 - b)); break;
        default: throw std::runtime_error("oracle: bad quadrangle rotation");
      }
    }
    return s;
  }
  throw std::runtime_error("oracle: faceQuadratureSequence unsupported face type");
}

// small dense helpers (column-major, A(r,c) = a[c*ld + r]) -----------------------------------------------------------
inline void invertInPlace(std::vector<double>& Ain, int n, std::vector<long double>* keep = nullptr) {  // Gauss-Jordan with partial pivoting in extended precision
  // (stand-in for Eigen's .inverse(); extended precision so that the inverse itself is correctly rounded for the
  // moderately ill-conditioned modal mass / least-squares matrices)
  std::vector<long double> A(Ain.begin(), Ain.end()), I((size_t)n * n, 0.0L);
  for (int i = 0; i < n; i++) I[i * n + i] = 1.0L;
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++) if (std::fabs((double)A[c * n + r]) > std::fabs((double)A[c * n + piv])) piv = r;
    if (A[c * n + piv] == 0.0L) throw std::runtime_error("oracle: singular matrix");
    if (piv != c) for (int k = 0; k < n; k++) { std::swap(A[k * n + c], A[k * n + piv]); std::swap(I[k * n + c], I[k * n + piv]); }
    long double d = 1.0L / A[c * n + c];
    for (int k = 0; k < n; k++) { A[k * n + c] *= d; I[k * n + c] *= d; }
    for (int r = 0; r < n; r++) if (r != c) {
      long double f = A[c * n + r];
      if (f != 0.0L) for (int k = 0; k < n; k++) { A[k * n + r] -= f * A[k * n + c]; I[k * n + r] -= f * I[k * n + c]; }
    }
  }
  for (size_t i = 0; i < Ain.size(); i++) Ain[i] = (double)I[i];
  if (keep) *keep = I;
}

}  // namespace orc
