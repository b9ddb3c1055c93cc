// mixed_tables.hpp — reference-element tables of the dense-operator ("mixed") device path: triangle and quadrangle blocks
// in one mesh (hybrid meshes of examples/karmanvortex_2d_cns.cpp, BASELINE configs[1]/[2] variants).
//
// SubrosaDG gets these from Gmsh 4.13.1 at run time (src/Mesh/Quadrature.cpp:27-34: "Gauss{2p}" volume / "Gauss{2p+1}" face
// rules; src/Mesh/BasisFunction.cpp:31-74,136-230: "H1Legendre{p}" modal basis and its gradient, "Lagrange{g}" geometry
// basis).  Triangles have Nq != Nb (SimulationControl.cpp:268-273: 12 points / 10 functions at p = 3), so the collocation
// trick of the tensor path does not apply: this path keeps the reference's modal coefficients and dense per-type operators
// Phi (Nq x Nb), grad Phi (Nq*D x Nb), Phi_f (Naq x Nb) exactly like ElementBasisFunction (BasisFunction.cpp:136-230).
// Host code only; nothing under oracle/ is included or linked.
#pragma once
#include <array>
#include <cmath>
#include <stdexcept>
#include <vector>

#include "host_tables.hpp"

namespace sdg {

// forward-mode dual number (value + 2 partials): exact gradients of the polynomial bases below
struct Dual2 {
  double v, d[2];
  Dual2(double a = 0.0) : v(a), d{0.0, 0.0} {}
  static Dual2 var(double a, int k) { Dual2 r(a); r.d[k] = 1.0; return r; }
};
inline Dual2 operator+(const Dual2& a, const Dual2& b) { Dual2 r(a.v + b.v); r.d[0] = a.d[0] + b.d[0]; r.d[1] = a.d[1] + b.d[1]; return r; }
inline Dual2 operator-(const Dual2& a, const Dual2& b) { Dual2 r(a.v - b.v); r.d[0] = a.d[0] - b.d[0]; r.d[1] = a.d[1] - b.d[1]; return r; }
inline Dual2 operator*(const Dual2& a, const Dual2& b) { Dual2 r(a.v * b.v); r.d[0] = a.d[0] * b.v + a.v * b.d[0]; r.d[1] = a.d[1] * b.v + a.v * b.d[1]; return r; }
inline Dual2 operator*(double a, const Dual2& b) { Dual2 r(a * b.v); r.d[0] = a * b.d[0]; r.d[1] = a * b.d[1]; return r; }

inline Dual2 legendreDual(int n, const Dual2& x) {
  if (n == 0) return Dual2(1.0);
  Dual2 pm(1.0), p = x;
  for (int k = 2; k <= n; k++) { Dual2 pn = (1.0 / k) * ((2.0 * k - 1.0) * (x * p) - (k - 1.0) * pm); pm = p; p = pn; }
  return p;
}
// L_n' = sum_{k = n-1, n-3, ...} (2k+1) L_k
inline Dual2 legendreDerivDual(int n, const Dual2& x) {
  Dual2 s(0.0);
  for (int k = n - 1; k >= 0; k -= 2) s = s + (2.0 * k + 1.0) * legendreDual(k, x);
  return s;
}
// Lobatto shape functions (Solin; docs/develop-note/develop-note.tex:272-322) and their kernels: l_{k+2} = l_0 l_1 phi_k
inline Dual2 lobattoDual(int k, const Dual2& x) {
  if (k == 0) return 0.5 * (Dual2(1.0) - x);
  if (k == 1) return 0.5 * (Dual2(1.0) + x);
  return (1.0 / std::sqrt(2.0 * (2.0 * k - 1.0))) * (legendreDual(k, x) - legendreDual(k - 2, x));
}
inline Dual2 lobattoKernelDual(int k, const Dual2& x) {
  const int kk = k + 2;
  return (-4.0 * (2.0 * kk - 1.0) / (kk * (kk - 1.0) * std::sqrt(2.0 * (2.0 * kk - 1.0)))) * legendreDerivDual(kk - 1, x);
}

inline int mixedNumBasis(int type, int p) { return type == kTriangle ? (p + 1) * (p + 2) / 2 : (p + 1) * (p + 1); }  // SimulationControl.cpp:243-266

// "H1Legendre{p}" functions at reference point (u, v): vertex, edge, interior functions, orientation block 0
inline void modalEval(int type, int p, double u, double v, std::vector<double>& val, std::vector<std::array<double, 2>>& grad) {
  const Dual2 U = Dual2::var(u, 0), V = Dual2::var(v, 1);
  std::vector<Dual2> f;
  if (type == kTriangle) {
    const Dual2 lam[3] = {Dual2(1.0) - U - V, U, V};
    for (int i = 0; i < 3; i++) f.push_back(lam[i]);
    for (int e = 0; e < 3; e++) {
      const Dual2 &a = lam[e], &b = lam[(e + 1) % 3];
      for (int k = 2; k <= p; k++) f.push_back(a * b * lobattoKernelDual(k - 2, b - a));
    }
    for (int n1 = 1; n1 <= p - 2; n1++) for (int n2 = 1; n1 + n2 <= p - 1; n2++)
      f.push_back(lam[0] * lam[1] * lam[2] * lobattoKernelDual(n1 - 1, lam[1] - lam[0]) * lobattoKernelDual(n2 - 1, lam[0] - lam[2]));
  } else if (type == kQuadrangle) {
    for (const Lat& m : modalFunctionIndex(kQuadrangle, p)) f.push_back(lobattoDual(m[0], U) * lobattoDual(m[1], V));
  } else {
    throw std::runtime_error("dense-operator path: triangle and quadrangle blocks only");
  }
  val.resize(f.size()); grad.resize(f.size());
  for (size_t j = 0; j < f.size(); j++) { val[j] = f[j].v; grad[j] = {f[j].d[0], f[j].d[1]}; }
}

// gmsh node order of the order-g Lagrange triangle: corners, edge interiors (0-1, 1-2, 2-0), then the nested triangle
inline void triangleNodes(const std::array<std::array<double, 2>, 3>& c, int g, std::vector<std::array<double, 2>>& out) {
  if (g == 0) { out.push_back({(c[0][0] + c[1][0] + c[2][0]) / 3.0, (c[0][1] + c[1][1] + c[2][1]) / 3.0}); return; }
  for (int i = 0; i < 3; i++) out.push_back(c[i]);
  if (g == 1) return;
  for (int e = 0; e < 3; e++) {
    const auto &a = c[e], &b = c[(e + 1) % 3];
    for (int i = 1; i < g; i++) { const double s = double(i) / g; out.push_back({a[0] + s * (b[0] - a[0]), a[1] + s * (b[1] - a[1])}); }
  }
  if (g < 3) return;
  auto bary = [&](double l0, double l1, double l2) { return std::array<double, 2>{l0 * c[0][0] + l1 * c[1][0] + l2 * c[2][0], l0 * c[0][1] + l1 * c[1][1] + l2 * c[2][1]}; };
  const double h = 1.0 / g;
  triangleNodes({bary(1 - 2 * h, h, h), bary(h, 1 - 2 * h, h), bary(h, h, 1 - 2 * h)}, g - 3, out);
}

// "Lagrange{g}" geometry basis on the gmsh-ordered nodes
struct GeomBasis2 {
  int type = 0, g = 1, nn = 0;
  std::vector<Lat> lat;                       // quadrangle: node lattice
  std::vector<std::array<int, 2>> mono;       // triangle: monomial exponents
  std::vector<double> coef;                   // triangle: coef[m*nn + j] = coefficient of monomial m in node function j
  GeomBasis2(int type_, int g_) : type(type_), g(g_) {
    if (type == kQuadrangle) { lat = gmshNodeLattice(kQuadrangle, g); nn = (int)lat.size(); return; }
    std::vector<std::array<double, 2>> nodes;
    triangleNodes({std::array<double, 2>{0.0, 0.0}, std::array<double, 2>{1.0, 0.0}, std::array<double, 2>{0.0, 1.0}}, g, nodes);
    nn = (int)nodes.size();
    for (int a = 0; a <= g; a++) for (int b = 0; a + b <= g; b++) mono.push_back({a, b});
    std::vector<double> V((size_t)nn * nn);
    for (int i = 0; i < nn; i++) for (int m = 0; m < nn; m++) V[(size_t)i * nn + m] = std::pow(nodes[i][0], mono[m][0]) * std::pow(nodes[i][1], mono[m][1]);
    invertDense(V, nn);   // V^-1[m][j]: coefficient of monomial m in the function that is 1 at node j
    coef = V;
  }
  void eval(double u, double v, std::vector<double>& val, std::vector<std::array<double, 2>>& grad) const {
    val.assign(nn, 0.0); grad.assign(nn, {0.0, 0.0});
    if (type == kQuadrangle) {
      std::vector<double> nodes1d, vu, du, vv, dv;
      for (int i = 0; i <= g; i++) nodes1d.push_back(-1.0 + 2.0 * i / g);
      lagrangeAt(nodes1d, u, vu, du); lagrangeAt(nodes1d, v, vv, dv);
      for (int m = 0; m < nn; m++) { val[m] = vu[lat[m][0]] * vv[lat[m][1]]; grad[m] = {du[lat[m][0]] * vv[lat[m][1]], vu[lat[m][0]] * dv[lat[m][1]]}; }
      return;
    }
    const Dual2 U = Dual2::var(u, 0), V = Dual2::var(v, 1);
    std::vector<Dual2> m(nn);
    for (int k = 0; k < nn; k++) { Dual2 r(1.0); for (int a = 0; a < mono[k][0]; a++) r = r * U; for (int b = 0; b < mono[k][1]; b++) r = r * V; m[k] = r; }
    for (int j = 0; j < nn; j++) { Dual2 s(0.0); for (int k = 0; k < nn; k++) s = s + coef[(size_t)k * nn + j] * m[k]; val[j] = s.v; grad[j] = {s.d[0], s.d[1]}; }
  }
};

// Symmetric triangle rules on the unit triangle (Dunavant 1985, degrees 2 / 4 / 6 = "Gauss{2p}" for p = 1, 2, 3; the point
// counts 3 / 6 / 12 are pinned by kTriangleQuadratureNumber, SimulationControl.cpp:269).  Weights sum to 1/2.
inline void triangleRule(int order, std::vector<std::array<double, 2>>& pts, std::vector<double>& wts) {
  pts.clear(); wts.clear();
  auto orbit3 = [&](double w, double a, double b) {
    const double P[3][3] = {{a, b, b}, {b, a, b}, {b, b, a}};
    for (auto& q : P) { pts.push_back({q[1], q[2]}); wts.push_back(0.5 * w); }
  };
  auto orbit6 = [&](double w, double a, double b, double c) {
    const double P[6][3] = {{a, b, c}, {a, c, b}, {b, a, c}, {b, c, a}, {c, a, b}, {c, b, a}};
    for (auto& q : P) { pts.push_back({q[1], q[2]}); wts.push_back(0.5 * w); }
  };
  if (order <= 2) orbit3(1.0 / 3.0, 2.0 / 3.0, 1.0 / 6.0);
  else if (order <= 4) { orbit3(0.223381589678011, 0.108103018168070, 0.445948490915965); orbit3(0.109951743655322, 0.816847572980459, 0.091576213509771); }
  else if (order <= 6) {
    orbit3(0.116786275726379, 0.501426509658179, 0.249286745170910);
    orbit3(0.050844906370207, 0.873821971016996, 0.063089014491502);
    orbit6(0.082851075618374, 0.053145049844817, 0.310352451033784, 0.636502499121399);
  } else throw std::runtime_error("triangle quadrature beyond degree 6 is not tabulated (triangle blocks: p <= 3)");
  double s = 0; for (double w : wts) s += w;
  for (double& w : wts) w *= 0.5 / s;   // 15-digit literature weights: renormalise to the reference measure (SimulationControl.cpp:226-228)
}

static const int kTriFace[3][2] = {{0, 1}, {1, 2}, {2, 0}};                       // SimulationControl.cpp:177-216
static const double kTriCornerXi[3][2] = {{0, 0}, {1, 0}, {0, 1}};
static const double kQuadCornerXi[4][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}};

// ElementBasisFunction + ElementQuadrature of one 2-D element type (row-major tables, variable count independent)
struct MixedTable {
  int type = 0, p = 0, g = 1, Nb = 0, Nq = 0, Nf = 0, Nqf = 0, Naq = 0, nn = 0;
  std::vector<double> xi, wq;            // volume points [Nq][2], weights
  std::vector<double> sf, wf;            // face rule (p+1 Gauss points on [-1,1])
  std::vector<double> Phi;               // [Nq][Nb]   modal_value_
  std::vector<double> dPhi;              // [Nq][2][Nb] modal_gradient_value_
  std::vector<double> PhiF;              // [Naq][Nb]  modal_adjacency_value_ (faces concatenated, SimulationControl.cpp:368-379)
  std::vector<double> Proj;              // [Nb][Nq]   (Phi^T Phi)^-1 Phi^T: unweighted least squares of InitialCondition.cpp:100-102
  std::vector<double> GN, dGN;           // geometry basis at volume points [Nq][nn], [Nq][2][nn]
  std::vector<double> GNf, dGNf;         // at face points [Naq][nn], [Naq][2][nn]
  std::vector<double> ftan;              // [Nf][2]: d(xi)/d(s) of the face's corner map
  int nbasic = 0;                        // kBasicNodeNumber (corner nodes)
  std::vector<double> NodalQ, NodalF;    // nodal_value_ [Nq][nbasic], nodal_adjacency_value_ [Naq][nbasic]: order-1 Lagrange basis (BasisFunction.cpp:149-208)

  void build(int type_, int p_, int g_) {
    type = type_; p = p_; g = g_;
    if (type != kTriangle && type != kQuadrangle) throw std::runtime_error("dense-operator path: triangle and quadrangle blocks only");
    if (type == kTriangle && p > 3) throw std::runtime_error("triangle blocks: p <= 3");
    Nb = mixedNumBasis(type, p); Nf = numFaces(type); Nqf = p + 1; Naq = Nf * Nqf;
    std::vector<double> x1, w1; gaussRule(p + 1, x1, w1);   // "Gauss{2p}" and "Gauss{2p+1}" on a line both have p+1 points
    sf = x1; wf = w1;
    if (type == kTriangle) {
      std::vector<std::array<double, 2>> pts; triangleRule(2 * p, pts, wq);
      for (auto& q : pts) { xi.push_back(q[0]); xi.push_back(q[1]); }
    } else if (p == 1) {
      // "Gauss2" on a quadrangle is Gmsh's SEVEN-point rule (kQuadrangleQuadratureNumber[2] = 7, SimulationControl.cpp:270): Radon's degree-5
      // formula for the square (Stroud C2 5-1) -- centre 8/7, (0, +-sqrt(14/15)) 20/63, (+-sqrt(3/5), +-sqrt(1/3)) 5/9.  P1 quadrangle blocks
      // (thermalcavity_2d / naca0010_2d / shearlayer_2d of the reference's examples) therefore run on this dense-operator path, not on the
      // collocation tensor kernels, whose nodes would be the 2 x 2 Gauss points.
      const double r = std::sqrt(14.0 / 15.0), a = std::sqrt(3.0 / 5.0), b = std::sqrt(1.0 / 3.0);
      const double P7[7][2] = {{0, 0}, {0, r}, {0, -r}, {a, b}, {a, -b}, {-a, b}, {-a, -b}};
      const double W7[7] = {8.0 / 7.0, 20.0 / 63.0, 20.0 / 63.0, 5.0 / 9.0, 5.0 / 9.0, 5.0 / 9.0, 5.0 / 9.0};
      for (int i = 0; i < 7; i++) { xi.push_back(P7[i][0]); xi.push_back(P7[i][1]); wq.push_back(W7[i]); }
    } else {
      for (int i = 0; i <= p; i++) for (int j = 0; j <= p; j++) { xi.push_back(x1[i]); xi.push_back(x1[j]); wq.push_back(w1[i] * w1[j]); }   // first coordinate slowest
    }
    Nq = (int)wq.size();
    GeomBasis2 gb(type, g); nn = gb.nn;
    GeomBasis2 g1(type, 1); nbasic = g1.nn;
    NodalQ.assign((size_t)Nq * nbasic, 0.0); NodalF.assign((size_t)Naq * nbasic, 0.0);
    std::vector<double> val; std::vector<std::array<double, 2>> grad;
    Phi.assign((size_t)Nq * Nb, 0.0); dPhi.assign((size_t)Nq * 2 * Nb, 0.0); GN.assign((size_t)Nq * nn, 0.0); dGN.assign((size_t)Nq * 2 * nn, 0.0);
    for (int q = 0; q < Nq; q++) {
      modalEval(type, p, xi[2 * q], xi[2 * q + 1], val, grad);
      if ((int)val.size() != Nb) throw std::runtime_error("internal: modal basis size mismatch");
      for (int b = 0; b < Nb; b++) { Phi[(size_t)q * Nb + b] = val[b]; for (int d = 0; d < 2; d++) dPhi[((size_t)q * 2 + d) * Nb + b] = grad[b][d]; }
      g1.eval(xi[2 * q], xi[2 * q + 1], val, grad);
      for (int m = 0; m < nbasic; m++) NodalQ[(size_t)q * nbasic + m] = val[m];
      gb.eval(xi[2 * q], xi[2 * q + 1], val, grad);
      for (int m = 0; m < nn; m++) { GN[(size_t)q * nn + m] = val[m]; for (int d = 0; d < 2; d++) dGN[((size_t)q * 2 + d) * nn + m] = grad[m][d]; }
    }
    {  // Proj = (Phi^T Phi)^-1 Phi^T in extended precision
      std::vector<long double> A((size_t)Nb * Nb), I((size_t)Nb * Nb, 0.0L);
      for (int a = 0; a < Nb; a++) for (int b = 0; b < Nb; b++) { long double s = 0; for (int q = 0; q < Nq; q++) s += (long double)Phi[(size_t)q * Nb + a] * Phi[(size_t)q * Nb + b]; A[(size_t)a * Nb + b] = s; }
      invertLong(A, I, Nb);
      Proj.assign((size_t)Nb * Nq, 0.0);
      for (int b = 0; b < Nb; b++) for (int q = 0; q < Nq; q++) { long double s = 0; for (int a = 0; a < Nb; a++) s += I[(size_t)b * Nb + a] * Phi[(size_t)q * Nb + a]; Proj[(size_t)b * Nq + q] = (double)s; }
    }
    PhiF.assign((size_t)Naq * Nb, 0.0); GNf.assign((size_t)Naq * nn, 0.0); dGNf.assign((size_t)Naq * 2 * nn, 0.0); ftan.assign((size_t)Nf * 2, 0.0);
    for (int f = 0; f < Nf; f++) {
      const double* c0 = type == kTriangle ? kTriCornerXi[kTriFace[f][0]] : kQuadCornerXi[kQuadFace[f][0]];
      const double* c1 = type == kTriangle ? kTriCornerXi[kTriFace[f][1]] : kQuadCornerXi[kQuadFace[f][1]];
      for (int k = 0; k < 2; k++) ftan[(size_t)f * 2 + k] = 0.5 * (c1[k] - c0[k]);   // P1 line map of the face corners, BasisFunction.cpp:76-111
      for (int j = 0; j < Nqf; j++) {
        const double s = sf[j], u = 0.5 * (1 - s) * c0[0] + 0.5 * (1 + s) * c1[0], v = 0.5 * (1 - s) * c0[1] + 0.5 * (1 + s) * c1[1];
        const int row = f * Nqf + j;
        modalEval(type, p, u, v, val, grad);
        for (int b = 0; b < Nb; b++) PhiF[(size_t)row * Nb + b] = val[b];
        g1.eval(u, v, val, grad);
        for (int m = 0; m < nbasic; m++) NodalF[(size_t)row * nbasic + m] = val[m];
        gb.eval(u, v, val, grad);
        for (int m = 0; m < nn; m++) { GNf[(size_t)row * nn + m] = val[m]; for (int d = 0; d < 2; d++) dGNf[((size_t)row * 2 + d) * nn + m] = grad[m][d]; }
      }
    }
  }

  // Gauss-Jordan with partial pivoting in extended precision (row-major): I <- A^-1
  static void invertLong(std::vector<long double>& A, std::vector<long double>& I, int n) {
    I.assign((size_t)n * n, 0.0L);
    for (int i = 0; i < n; i++) I[(size_t)i * n + i] = 1.0L;
    for (int c = 0; c < n; c++) {
      int piv = c;
      for (int r = c + 1; r < n; r++) if (fabsl(A[(size_t)r * n + c]) > fabsl(A[(size_t)piv * n + c])) piv = r;
      if (A[(size_t)piv * n + c] == 0.0L) throw std::runtime_error("singular matrix");
      if (piv != c) for (int k = 0; k < n; k++) { std::swap(A[(size_t)c * n + k], A[(size_t)piv * n + k]); std::swap(I[(size_t)c * n + k], I[(size_t)piv * n + k]); }
      const long double d = 1.0L / A[(size_t)c * n + c];
      for (int k = 0; k < n; k++) { A[(size_t)c * n + k] *= d; I[(size_t)c * n + k] *= d; }
      for (int r = 0; r < n; r++) if (r != c) {
        const long double f = A[(size_t)r * n + c];
        if (f != 0.0L) for (int k = 0; k < n; k++) { A[(size_t)r * n + k] -= f * A[(size_t)c * n + k]; I[(size_t)r * n + k] -= f * I[(size_t)c * n + k]; }
      }
    }
  }
};

}  // namespace sdg
