// host_plan.hpp — host-side flattening of a SubrosaDG mesh into the SoA arrays the device kernels consume.
//
// Restates the setup arithmetic of src/Mesh/Geometry.cpp (getElementJacobian :44-67, getAdjacencyElementJacobian :69-86,
// calculateNormalVector :102-169, getElementQuality "minEdge" :29-42) without gmsh, for quadrangle / hexahedron blocks
// with gmsh-ordered Lagrange geometry nodes of any order, and turns the reference's face records
// (src/Mesh/ReadControl.cpp:72-83) into per-thread-block face lists.  Pure host code: no CUDA, no oracle.
#pragma once
#include <algorithm>
#include <cstring>
#include <numeric>
#include <string>

#include "host_tables.hpp"

namespace sdg {

struct FaceInput {
  int nInt = 0, nBnd = 0;
  std::vector<int> le, lt, lf, re, rt, rf, rot, bc, phys;
};

struct BlockPlan {
  int type = 0, D = 0, p = 0, g = 1, n = 0, nGhost = 0, nOwned = 0, nn = 0;
  TensorTables T;
  std::vector<Lat> lattice;          // geometry node lattice (gmsh order)
  std::vector<double> X;             // [n][nn][D]
  bool affine = true;
  std::vector<int> perm, inv;        // caller index -> internal position, and back
  std::vector<double> geoE, invjw, minEdge;  // internal order
  int K = 1, nChunks = 0;
  std::vector<int> chunkFaceOff;     // [nChunks+1]
  std::vector<int> faceRec;          // 4 ints per entry: eL, eR (internal; -1 boundary), faceId, packed(lfL|lfR<<4|rot<<8|bc<<12)
  std::vector<int> chunkInterior, chunkBoundary;  // chunks without / with a ghost parent on one of their faces
};

// Host image of what the trace-based line kernels (nsl_kernels.cuh) read besides the states: per-(element, local face) link
// records and affine face geometry, boundary-face records, the face-point correspondence tables in natural point order.
struct LinePlan {
  std::vector<int> links;       // [nOwned][6] x {other parent | -1, face id, lfo | rot<<3 | bc<<5 | amRight<<8 | handles<<9 | inChunk<<10, 0}
  std::vector<double> lfGeo;    // affine: [nOwned][6][4] = normal of the face (outward from its LEFT parent), |J| scale
  std::vector<int> bndRec;      // [nBnd] x {left parent, its local face, boundary condition, face id}
  std::vector<unsigned char> partner, jLeft;
  double cLift = 0.0;
};

struct GeomEval {  // Lagrange geometry map of one tensor element type
  int D, g, nn;
  std::vector<Lat> lat;
  std::vector<double> nodes1d;
  GeomEval(int type, int g_) : D(elemDim(type)), g(g_) {
    lat = gmshNodeLattice(type, g); nn = (int)lat.size();
    for (int i = 0; i <= g; i++) nodes1d.push_back(-1.0 + 2.0 * i / g);
  }
  // node weights for value and for d/dxi_k at parent point xi: wv[nn], wd[D][nn]
  void weights(const double* xi, std::vector<double>& wv, std::vector<double>& wd) const {
    std::vector<double> v[3], d[3];
    for (int k = 0; k < D; k++) lagrangeAt(nodes1d, xi[k], v[k], d[k]);
    wv.assign(nn, 1.0); wd.assign((size_t)D * nn, 1.0);
    for (int m = 0; m < nn; m++) {
      for (int k = 0; k < D; k++) wv[m] *= v[k][lat[m][k]];
      for (int k = 0; k < D; k++) for (int kk = 0; kk < D; kk++) wd[(size_t)k * nn + m] *= (kk == k ? d[kk][lat[m][kk]] : v[kk][lat[m][kk]]);
    }
  }
};

inline double invertSmall(int D, const double* Jt, double* inv) {  // row-major
  if (D == 1) { inv[0] = 1.0 / Jt[0]; return Jt[0]; }
  if (D == 2) {
    const double det = Jt[0] * Jt[3] - Jt[1] * Jt[2];
    inv[0] = Jt[3] / det; inv[1] = -Jt[1] / det; inv[2] = -Jt[2] / det; inv[3] = Jt[0] / det;
    return det;
  }
  const double a = Jt[0], b = Jt[1], c = Jt[2], d = Jt[3], e = Jt[4], f = Jt[5], g = Jt[6], h = Jt[7], i = Jt[8];
  const double A = e * i - f * h, B = f * g - d * i, C = d * h - e * g;
  const double det = a * A + b * B + c * C;
  inv[0] = A / det; inv[1] = (c * h - b * i) / det; inv[2] = (b * f - c * e) / det;
  inv[3] = B / det; inv[4] = (a * i - c * g) / det; inv[5] = (c * d - a * f) / det;
  inv[6] = C / det; inv[7] = (b * g - a * h) / det; inv[8] = (a * e - b * d) / det;
  return det;
}

inline uint64_t spreadBits(uint64_t v, int D) {  // interleave the low 21 bits of v with D-1 zero bits
  uint64_t r = 0;
  for (int b = 0; b < 21; b++) r |= ((v >> b) & 1ull) << (D * b);
  return r;
}

// parent coordinates of face point j of local face f
inline void facePointParent(const TensorTables& T, int f, int j, double* xi) {
  const int D = T.D, N = T.N, dn = T.faceDir[f];
  const int base = T.faceBase[(size_t)f * T.NQF + j];
  for (int k = 0; k < D; k++) {
    int st = 1; for (int m = D - 1; m > k; m--) st *= N;
    xi[k] = k == dn ? (T.faceSide[f] ? 1.0 : -1.0) : T.x[(base / st) % N];
  }
}

struct MeshPlan {
  int D = 0, p = 0;
  BlockPlan blk;               // single tensor block (quadrangle or hexahedron)
  FaceInput F;
  std::vector<double> geoF;    // affine: [nf][D+1]; curved: [nf][D+1][NQF]
  int NQF = 0;

  // ---- element geometry ----------------------------------------------------------------------------------------------
  void buildBlock(int reorder, int chunk) {
    BlockPlan& B = blk;
    const int Dm = B.D, nn = B.nn, N = B.T.N, NN = B.T.NN, n = B.n;
    GeomEval G(B.type, B.g);
    B.lattice = G.lat;
    const int nc = Dm == 1 ? 2 : Dm == 2 ? 4 : 8;
    // affine test (order-1 geometry whose corners form a parallelogram / parallelepiped)
    B.affine = B.g == 1;
    if (B.affine && Dm > 1) {
      bool ok = true;
#pragma omp parallel for schedule(static) reduction(&& : ok)
      for (int e = 0; e < n; e++) {
        const double* X = &B.X[(size_t)e * nn * Dm];
        double scale = 0, dev = 0;
        for (int l = 0; l < Dm; l++) {
          const double x0 = X[l], x1 = X[1 * Dm + l], x3 = X[3 * Dm + l];
          scale = std::max(scale, std::max(std::fabs(x1 - x0), std::fabs(x3 - x0)));
          dev = std::max(dev, std::fabs(X[2 * Dm + l] - (x1 + x3 - x0)));
          if (Dm == 3) {
            const double x4 = X[4 * Dm + l];
            scale = std::max(scale, std::fabs(x4 - x0));
            dev = std::max(dev, std::fabs(X[5 * Dm + l] - (x1 + x4 - x0)));
            dev = std::max(dev, std::fabs(X[7 * Dm + l] - (x3 + x4 - x0)));
            dev = std::max(dev, std::fabs(X[6 * Dm + l] - (x1 + x3 + x4 - 2 * x0)));
          }
        }
        ok = ok && (dev <= 1e-13 * scale);
      }
      B.affine = ok;
    }
    // internal order: Morton curve through the centroids of the owned elements; ghosts keep their order at the end
    B.perm.resize(n); B.inv.resize(n);
    {
      std::vector<std::pair<uint64_t, int>> key(B.nOwned);
      double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
      std::vector<double> cen((size_t)B.nOwned * Dm);
      for (int e = 0; e < B.nOwned; e++) for (int l = 0; l < Dm; l++) {
        double s = 0; for (int c = 0; c < nc; c++) s += B.X[((size_t)e * nn + c) * Dm + l];
        s /= nc; cen[(size_t)e * Dm + l] = s; lo[l] = std::min(lo[l], s); hi[l] = std::max(hi[l], s);
      }
      // common cell size so that a uniform grid maps onto consecutive integers (bricks of 2^k cells)
      double ext = 0; for (int l = 0; l < Dm; l++) ext = std::max(ext, hi[l] - lo[l]);
      double h = 1e300;
      for (int e = 0; e < std::min(B.nOwned, 4096); e++) { const double* X = &B.X[(size_t)e * nn * Dm]; double s = 0; for (int l = 0; l < Dm; l++) s += (X[Dm + l] - X[l]) * (X[Dm + l] - X[l]); h = std::min(h, std::sqrt(s)); }
      if (!(h > 0) || ext / h > 1.9e6) h = ext / 1.9e6 + 1e-300;
      for (int e = 0; e < B.nOwned; e++) {
        uint64_t k = 0;
        if (reorder) for (int l = 0; l < Dm; l++) k |= spreadBits((uint64_t)std::floor((cen[(size_t)e * Dm + l] - lo[l]) / h + 0.5), Dm) << l;
        key[e] = {k, e};
      }
      if (reorder) std::stable_sort(key.begin(), key.end());
      for (int pos = 0; pos < B.nOwned; pos++) { B.inv[pos] = key[pos].second; B.perm[key[pos].second] = pos; }
      for (int e = B.nOwned; e < n; e++) { B.perm[e] = e; B.inv[e] = e; }
    }
    B.K = chunk;
    B.nChunks = (B.nOwned + B.K - 1) / B.K;
    // metric terms
    B.minEdge.assign(n, 0.0);
    if (B.affine) B.geoE.assign((size_t)n * (Dm * Dm + 1), 0.0);
    else { B.geoE.assign((size_t)n * Dm * Dm * NN, 0.0); B.invjw.assign((size_t)n * NN, 0.0); }
    std::vector<std::vector<double>> wd(NN);
    if (!B.affine) {
      std::vector<double> wv;
      for (int q = 0; q < NN; q++) {
        double xi[3]; for (int k = 0; k < Dm; k++) { int st = 1; for (int m = Dm - 1; m > k; m--) st *= N; xi[k] = B.T.x[(q / st) % N]; }
        G.weights(xi, wv, wd[q]);
      }
    }
    bool bad = false;
#pragma omp parallel for schedule(static) reduction(|| : bad)
    for (int pos = 0; pos < n; pos++) {
      const int e = B.inv[pos];
      const double* X = &B.X[(size_t)e * nn * Dm];
      if (B.affine) {
        double Jt[9], inv[9];
        const int cn[3] = {1, 3, 4};
        for (int k = 0; k < Dm; k++) for (int l = 0; l < Dm; l++) Jt[k * Dm + l] = 0.5 * (X[cn[k] * Dm + l] - X[l]);
        const double det = invertSmall(Dm, Jt, inv);
        if (!(det > 0)) bad = true;
        double* g = &B.geoE[(size_t)pos * (Dm * Dm + 1)];
        for (int c = 0; c < Dm; c++) for (int dd = 0; dd < Dm; dd++) g[dd * Dm + c] = inv[c * Dm + dd] * det;
        g[Dm * Dm] = det;
      } else {
        for (int q = 0; q < NN; q++) {
          double Jt[9] = {0}, inv[9];
          for (int k = 0; k < Dm; k++) for (int m = 0; m < nn; m++) { const double w = wd[q][(size_t)k * nn + m]; for (int l = 0; l < Dm; l++) Jt[k * Dm + l] += w * X[m * Dm + l]; }
          const double det = invertSmall(Dm, Jt, inv);
          if (!(det > 0)) bad = true;
          const double w = det * B.T.wq[q];
          for (int c = 0; c < Dm; c++) for (int dd = 0; dd < Dm; dd++) B.geoE[((size_t)pos * Dm * Dm + dd * Dm + c) * NN + q] = inv[c * Dm + dd] * w;
          B.invjw[(size_t)pos * NN + q] = 1.0 / w;
        }
      }
      double me = 1e300;
      auto dist = [&](int a, int b) { double s = 0; for (int l = 0; l < Dm; l++) { const double d = X[a * Dm + l] - X[b * Dm + l]; s += d * d; } return std::sqrt(s); };
      if (Dm == 1) me = dist(0, 1);
      else if (Dm == 2) for (int k = 0; k < 4; k++) me = std::min(me, dist(k, (k + 1) % 4));
      else for (auto& ed : kHexEdge) me = std::min(me, dist(ed[0], ed[1]));
      B.minEdge[pos] = me;
    }
    if (bad) throw std::runtime_error("non-positive Jacobian determinant");
  }

  // quadrature_node_coordinate_ in caller order [n][NN][D]
  void quadratureCoordinates(double* out) const {
    const BlockPlan& B = blk;
    const int Dm = B.D, nn = B.nn, N = B.T.N, NN = B.T.NN;
    GeomEval G(B.type, B.g);
    std::vector<std::vector<double>> wv(NN);
    std::vector<double> wd;
    for (int q = 0; q < NN; q++) {
      double xi[3]; for (int k = 0; k < Dm; k++) { int st = 1; for (int m = Dm - 1; m > k; m--) st *= N; xi[k] = B.T.x[(q / st) % N]; }
      G.weights(xi, wv[q], wd);
    }
#pragma omp parallel for schedule(static)
    for (int e = 0; e < B.n; e++) {
      const double* X = &B.X[(size_t)e * nn * Dm];
      for (int q = 0; q < NN; q++) for (int l = 0; l < Dm; l++) {
        double s = 0; for (int m = 0; m < nn; m++) s += wv[q][m] * X[m * Dm + l];
        out[((size_t)e * NN + q) * Dm + l] = s;
      }
    }
  }

  // ---- face geometry (from the LEFT parent's map) and boundary point coordinates -----------------------------------------------
  void buildFaces(double* xbOut /* may be null: [nBnd][NQF][D] */, bool onlyCoords) {
    const BlockPlan& B = blk;
    const int Dm = B.D, nn = B.nn;
    const int nf = F.nInt + F.nBnd;
    NQF = B.T.NQF;
    GeomEval G(B.type, B.g);
    const int NF = B.T.NF;
    std::vector<std::vector<double>> wv((size_t)NF * NQF), wd((size_t)NF * NQF);
    for (int f = 0; f < NF; f++) for (int j = 0; j < NQF; j++) { double xi[3]; facePointParent(B.T, f, j, xi); G.weights(xi, wv[(size_t)f * NQF + j], wd[(size_t)f * NQF + j]); }
    if (!onlyCoords) geoF.assign(B.affine ? (size_t)nf * (Dm + 1) : (size_t)nf * (Dm + 1) * NQF, 0.0);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < nf; i++) {
      const int e = F.le[i], f = F.lf[i];
      const double* X = &B.X[(size_t)e * nn * Dm];
      for (int j = 0; j < NQF; j++) {
        const std::vector<double>& v = wv[(size_t)f * NQF + j];
        const std::vector<double>& d = wd[(size_t)f * NQF + j];
        if (xbOut && i >= F.nInt) for (int l = 0; l < Dm; l++) { double s = 0; for (int m = 0; m < nn; m++) s += v[m] * X[m * Dm + l]; xbOut[((size_t)(i - F.nInt) * NQF + j) * Dm + l] = s; }
        if (onlyCoords) continue;
        if (B.affine && j > 0) continue;
        double Jt[9] = {0};
        for (int k = 0; k < Dm; k++) for (int m = 0; m < nn; m++) { const double w = d[(size_t)k * nn + m]; for (int l = 0; l < Dm; l++) Jt[k * Dm + l] += w * X[m * Dm + l]; }
        double nv[3] = {0, 0, 0}, scale;
        const double* ta = &B.T.faceTan[((size_t)f * 2 + 0) * 3];
        if (Dm == 1) {  // Geometry.cpp:102-112: end points of a line
          nv[0] = f == 0 ? -1.0 : 1.0; scale = 1.0;
        } else if (Dm == 2) {  // Geometry.cpp:114-129: normal = (t_y, -t_x)/|t|
          double t[2] = {0, 0};
          for (int k = 0; k < 2; k++) for (int l = 0; l < 2; l++) t[l] += ta[k] * Jt[k * 2 + l];
          scale = std::sqrt(t[0] * t[0] + t[1] * t[1]);
          nv[0] = t[1] / scale; nv[1] = -t[0] / scale;
        } else {  // :131-148: normal = d_s x  ×  d_t x, normalised
          const double* tb = &B.T.faceTan[((size_t)f * 2 + 1) * 3];
          double a[3] = {0, 0, 0}, b[3] = {0, 0, 0};
          for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++) { a[l] += ta[k] * Jt[k * 3 + l]; b[l] += tb[k] * Jt[k * 3 + l]; }
          const double c[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
          scale = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
          for (int l = 0; l < 3; l++) nv[l] = c[l] / scale;
        }
        if (B.affine) { double* g = &geoF[(size_t)i * (Dm + 1)]; for (int l = 0; l < Dm; l++) g[l] = nv[l]; g[Dm] = scale; }
        else { double* g = &geoF[(size_t)i * (Dm + 1) * NQF]; for (int l = 0; l < Dm; l++) g[(size_t)l * NQF + j] = nv[l]; g[(size_t)Dm * NQF + j] = scale * B.T.wf[j]; }
      }
    }
  }

  // ---- trace-based line kernels: every owned (element, local face) names its other parent -----------------------------------------
  LinePlan buildLinePlan() const {
    const BlockPlan& B = blk;
    if (B.D != 3 || B.T.N != 4) throw std::runtime_error("line plan: P3 hexahedra only");
    LinePlan P;
    const int nf = F.nInt + F.nBnd, NF = 6, NL = 16;
    P.links.assign((size_t)B.nOwned * NF * 4, -2);
    if (B.affine) P.lfGeo.assign((size_t)B.nOwned * NF * 4, 0.0);
    P.bndRec.assign((size_t)std::max(F.nBnd, 1) * 4, 0);
    auto sameChunk = [&](int a, int b) { return a < B.nOwned && b < B.nOwned && a / B.K == b / B.K; };
    for (int i = 0; i < nf; i++) {
      const bool interior = i < F.nInt;
      const int pL = B.perm[F.le[i]], pR = interior ? B.perm[F.re[i]] : -1;
      const int lfL = F.lf[i], lfR = interior ? F.rf[i] : 0, rot = interior ? F.rot[i] : 0, bc = interior ? 0 : (F.bc[i] & 7);
      const bool in = interior && sameChunk(pL, pR);
      auto put = [&](int pos, int lf, int other, int z) {
        int* r = &P.links[((size_t)pos * NF + lf) * 4];
        if (r[0] != -2) throw std::runtime_error("line plan: an element face appears in two face records");
        r[0] = other; r[1] = i; r[2] = z; r[3] = 0;
        if (B.affine) { double* g = &P.lfGeo[((size_t)pos * NF + lf) * 4]; for (int l = 0; l < 4; l++) g[l] = geoF[(size_t)i * 4 + l]; }
      };
      if (pL < B.nOwned) put(pL, lfL, pR, lfR | (rot << 3) | (bc << 5) | (1 << 9) | ((in ? 1 : 0) << 10));
      if (interior && pR < B.nOwned) put(pR, lfR, pL, lfL | (rot << 3) | (1 << 8) | ((in ? 0 : 1) << 9) | ((in ? 1 : 0) << 10));
      if (!interior) { int* r = &P.bndRec[(size_t)(i - F.nInt) * 4]; r[0] = pL; r[1] = lfL; r[2] = bc; r[3] = i; }
    }
    for (size_t k = 0; k < P.links.size(); k += 4) if (P.links[k] == -2) throw std::runtime_error("line plan: an element face without a face record");
    // natural point index of face point jf of face f: its two tangential lattice indices, lower axis first
    const TensorTables& T = B.T;
    std::vector<int> nat2jf(NF * NL), jf2nat(NF * NL);
    for (int f = 0; f < NF; f++) for (int jf = 0; jf < NL; jf++) {
      const int base = T.faceBase[(size_t)f * NL + jf], i = (base / 16) % 4, j = (base / 4) % 4, k = base % 4, dn = T.faceDir[f];
      const int nat = dn == 0 ? j * 4 + k : dn == 1 ? i * 4 + k : i * 4 + j;
      nat2jf[f * NL + nat] = jf; jf2nat[f * NL + jf] = nat;
    }
    P.partner.assign(6 * 6 * 4 * 2 * 16, 0); P.jLeft.assign(6 * 4 * 2 * 16, 0);
    for (int rot = 0; rot < 4; rot++) {
      const std::vector<int> seq = faceSequence(kQuadrangle, 4, rot);   // right parent's point of the left parent's point j
      std::vector<int> inv(NL); for (int jl = 0; jl < NL; jl++) inv[seq[jl]] = jl;
      for (int f = 0; f < NF; f++) for (int amR = 0; amR < 2; amR++) for (int t = 0; t < NL; t++) {
        const int jfMine = nat2jf[f * NL + t];
        const int jL = amR ? inv[jfMine] : jfMine, jO = amR ? jL : seq[jfMine];
        P.jLeft[((f * 4 + rot) * 2 + amR) * 16 + t] = (unsigned char)jL;
        for (int lfo = 0; lfo < NF; lfo++) P.partner[(((f * 6 + lfo) * 4 + rot) * 2 + amR) * 16 + t] = (unsigned char)jf2nat[lfo * NL + jO];
      }
    }
    for (int a = 0; a < 4; a++) P.cLift += T.Lend[a] * T.Lend[a] / T.w[a];
    return P;
  }

  // ---- per-chunk face lists ------------------------------------------------------------------------------------------------------
  void buildChunkFaces() {
    BlockPlan& B = blk;
    const int nf = F.nInt + F.nBnd;
    // Within a chunk's list the faces that need a parent from OUTSIDE the chunk (a long-latency gather) come first and the
    // cheap ones (both parents in the chunk, or boundary faces) last, so that the tail of the kernels' face loop is short.
    std::vector<int> count(B.nChunks + 1, 0), countCheap(B.nChunks + 1, 0);
    std::vector<char> touchesGhost(B.nChunks, 0);
    auto chunkOf = [&](int pos) { return pos < B.nOwned ? pos / B.K : -1; };
    for (int pass = 0; pass < 2; pass++) {
      std::vector<int> cursor, cursorCheap;
      if (pass == 1) {
        B.chunkFaceOff.assign(B.nChunks + 1, 0);
        for (int c = 0; c < B.nChunks; c++) B.chunkFaceOff[c + 1] = B.chunkFaceOff[c] + count[c];
        B.faceRec.assign((size_t)B.chunkFaceOff[B.nChunks] * 4, 0);
        cursor.assign(B.chunkFaceOff.begin(), B.chunkFaceOff.end() - 1);
        cursorCheap.resize(B.nChunks);
        for (int c = 0; c < B.nChunks; c++) cursorCheap[c] = B.chunkFaceOff[c + 1] - countCheap[c];
      }
      for (int i = 0; i < nf; i++) {
        const bool interior = i < F.nInt;
        const int pL = B.perm[F.le[i]], pR = interior ? B.perm[F.re[i]] : -1;
        const int cL = chunkOf(pL), cR = interior ? chunkOf(pR) : -1;
        const int packed = F.lf[i] | ((interior ? F.rf[i] : 0) << 4) | ((interior ? F.rot[i] : 0) << 8) | ((F.bc[i] & 15) << 12);
        const bool cheap = !interior || cL == cR;
        int targets[2] = {cL, (cR != cL) ? cR : -1};
        for (int t : targets) if (t >= 0) {
          if (pass == 0) { count[t]++; if (cheap) countCheap[t]++; if (pL >= B.nOwned || (interior && pR >= B.nOwned)) touchesGhost[t] = 1; }
          else { int* r = &B.faceRec[(size_t)(cheap ? cursorCheap[t]++ : cursor[t]++) * 4]; r[0] = pL; r[1] = pR; r[2] = i; r[3] = packed; }
        }
      }
    }
    B.chunkInterior.clear(); B.chunkBoundary.clear();
    for (int c = 0; c < B.nChunks; c++) (touchesGhost[c] ? B.chunkBoundary : B.chunkInterior).push_back(c);
  }
};

}  // namespace sdg
