// ns_kernels.cuh — Navier–Stokes (BR1 / BR2) stage kernels for quadrangle / hexahedron blocks on sm_100a (fp64).
//
// Two launches per RK stage instead of the reference's eight sweeps (src/Solver/TimeIntegration.cpp:339-348):
//
//   pass G (nsGradKernel)   G1 calculateElementGardientQuadrature      SpatialDiscrete.cpp:294-322
//                           G2 calculate…AdjacencyElementGardientQuadrature :844-968  (ViscousFlux.cpp:26-56)
//                           G3 calculateElementGardientResidual        :1034-1068
//                           G4 updateElementGardientBasisFunctionCoefficient  TimeIntegration.cpp:200-228
//     writes ONE gradient field to HBM: the volume gradient G_vol (BR2) or the total gradient G_vol + G_lift (BR1).
//
//   pass R (nsStageKernel)  R1-R4 + K as in tensor_kernels.cuh, with the viscous flux (ViscousFlux.cpp:59-153) and the
//                           wall / far-field boundary treatment of SpatialDiscrete.cpp:750-842.
//
// BR2 in the collocation basis: the lifting coefficient block of local face f (PerElementInterfaceGradientSolver<BR2>,
// SolveControl.cpp:92-105) is  G_f = M^-1 (A_int[:,f] Φ_f[f,:])  with M diagonal and Φ_f the end-point interpolation along the
// face-normal lines, i.e.  G_f(node) = invjw(node) · l_{a(node)}(±1) · [n ⊗ ½(U_R−U_L) |J|w](face point of that line).
// It is rank one per face point, so it is never stored: pass R rebuilds  trace_f(G_vol + G_f)  and  G_vol + Σ_f G_f  from
// the face jumps it needs anyway.  The 6 extra M^-1 applications and the Nf coefficient blocks of the reference disappear.
#pragma once
#include "tensor_kernels.cuh"

namespace sdg {

template <int D, int N, int K>
struct NsLayout {
  static constexpr int NV = D + 2, NG = NV * D, NN = Pow<N, D>::v, NQF = NN / N, NF = 2 * D, NAQ = NF * NQF;
  static constexpr int oU = 0;                          // [K][NV][NN]
  static constexpr int oG = oU + K * NV * NN;           // [K][NG][NN]   gradient field of the chunk's own elements (pass R)
  static constexpr int oF = oG + K * NG * NN;           // [K][NV][NN]   contravariant flux of one direction
  static constexpr int oFlux = oF + K * NV * NN;        // [K][NV][NAQ]  flux slots (pass G: {U} |J|w slots)
  static constexpr int oB = oFlux + K * NV * NAQ;       // [K][NV][NAQ]  jump slots  ½(U_R−U_L)|J|w  /  (U_b−U_L)|J|w
  static constexpr int oN = oB + K * NV * NAQ;          // [K][D][NAQ]   face normal at every slot
  static constexpr int oTab = oN + K * D * NAQ;
  static constexpr int nTabD = 2 * N * N + 2 * N + NN + NQF;
  static constexpr int nDoubles = oTab + nTabD;
  static constexpr int nBytesTab = NF * NQF + 4 * NQF + NF * NN;
  static constexpr size_t bytes = sizeof(double) * nDoubles + ((nBytesTab + 15) / 16) * 16;
  static constexpr int ITERS = (K * NN + kThreads - 1) / kThreads;
};

template <int D, int N, int K>
struct NsShared {
  using L = NsLayout<D, N, K>;
  double *sU, *sG, *sF, *sFlux, *sB, *sN, *sDm, *sLend, *sK1, *sWq, *sWf;
  unsigned char *sFaceBase, *sSeq, *sNodePt;
  __device__ __forceinline__ void carve(double* smem) {
    sU = smem + L::oU; sG = smem + L::oG; sF = smem + L::oF; sFlux = smem + L::oFlux; sB = smem + L::oB; sN = smem + L::oN;
    sDm = smem + L::oTab; sLend = sDm + N * N; sK1 = sLend + 2 * N; sWq = sK1 + N * N; sWf = sWq + L::NN;
    sFaceBase = reinterpret_cast<unsigned char*>(smem + L::nDoubles); sSeq = sFaceBase + L::NF * L::NQF; sNodePt = sSeq + 4 * L::NQF;
  }
  __device__ __forceinline__ void loadTables(const TensorDev& T, int tid) {
    for (int i = tid; i < N * N; i += kThreads) { sDm[i] = T.Dm[i]; sK1[i] = T.K1[i]; }
    for (int i = tid; i < 2 * N; i += kThreads) sLend[i] = T.Lend[i];
    for (int i = tid; i < L::NN; i += kThreads) sWq[i] = T.wq[i];
    for (int i = tid; i < L::NQF; i += kThreads) sWf[i] = T.wf[i];
    for (int i = tid; i < L::NF * L::NQF; i += kThreads) sFaceBase[i] = (unsigned char)T.faceBase[i];
    for (int i = tid; i < 4 * L::NQF; i += kThreads) sSeq[i] = (unsigned char)T.seq[i];
    for (int i = tid; i < L::NF * L::NN; i += kThreads) sNodePt[i] = T.nodeFacePt[i];
  }
};

// face geometry at one face point: unit normal (outward from the LEFT parent) and |J|·w
template <int D, int NQF, bool AFFINE>
__device__ __forceinline__ void faceGeometryAt(const StageArgs& A, int faceId, int j, const double* sWf, double* n, double& jw) {
  if constexpr (AFFINE) {
    const double* g = A.geoF + (size_t)faceId * (D + 1);
#pragma unroll
    for (int d = 0; d < D; d++) n[d] = __ldg(g + d);
    jw = __ldg(g + D) * sWf[j];
  } else {
    const double* g = A.geoF + (size_t)faceId * (D + 1) * NQF + j;
#pragma unroll
    for (int d = 0; d < D; d++) n[d] = __ldg(g + d * NQF);
    jw = __ldg(g + D * NQF);
  }
}
// 1 / (detJ w) at node q of element e (the diagonal inverse mass matrix, Geometry.cpp:88-100 in the collocation basis)
template <int D, int NN, bool AFFINE>
__device__ __forceinline__ double invJwAt(const StageArgs& A, int e, int q, const double* sWq) {
  if constexpr (AFFINE) return 1.0 / (__ldg(A.geoE + (size_t)e * ((D * D + 2) & ~1) + D * D) * sWq[q]);
  else return __ldg(A.invjw + (size_t)e * NN + q);
}
// entry [dd][c] of (J^T)^-1 detJ w at node q of element e
template <int D, int NN, bool AFFINE>
__device__ __forceinline__ double metricAt(const StageArgs& A, int e, int q, int dd, int c, const double* sWq) {
  if constexpr (AFFINE) return __ldg(A.geoE + (size_t)e * ((D * D + 2) & ~1) + dd * D + c) * sWq[q];
  else return __ldg(A.geoE + ((size_t)e * (D * D) + dd * D + c) * NN + q);
}

// =====================================================================================================================
// pass G
// =====================================================================================================================
template <int D, int N, int K, bool AFFINE>
__global__ void __launch_bounds__(kThreads, 1) nsGradKernel(const __grid_constant__ StageArgs A) {
  using L = NsLayout<D, N, K>;
  constexpr int NV = L::NV, NG = L::NG, NN = L::NN, NQF = L::NQF, NF = L::NF, NAQ = L::NAQ, ITERS = L::ITERS;
  extern __shared__ __align__(16) double smem[];
  NsShared<D, N, K> S; S.carve(smem);
  const int tid = threadIdx.x;
  const int chunk = A.chunkList ? A.chunkList[blockIdx.x] : blockIdx.x;
  const int e0 = chunk * K;
  const int ne = min(K, A.nOwned - e0);
  const int nNodes = ne * NN;
  const Phys<0> ph(A.phys);
  const bool br1 = A.phys.visc == kBR1;
  {
    const double* src = A.Uin + (size_t)e0 * NV * NN;
    for (int i = tid; i < ne * NV * NN; i += kThreads) S.sU[i] = src[i];
  }
  S.loadTables(*A.tab, tid);
  __syncthreads();

  // ---- G2: {U} n |J|w (volume-gradient flux) and ½(U_R−U_L) n |J|w (interface-gradient flux) at the face points -----------
  const int f0 = A.chunkFaceOff[chunk], nfc = A.chunkFaceOff[chunk + 1] - f0;
  for (int fp = tid; fp < nfc * NQF; fp += kThreads) {
    const int fi = fp / NQF, j = fp - fi * NQF;
    const int4 rec = __ldg(A.faceRec + f0 + fi);
    const int eL = rec.x, eR = rec.y, faceId = rec.z;
    const int lfL = rec.w & 15, lfR = (rec.w >> 4) & 15, rot = (rec.w >> 8) & 15, bc = (rec.w >> 12) & 15;
    double n[D], jw;
    faceGeometryAt<D, NQF, AFFINE>(A, faceId, j, S.sWf, n, jw);
    double consL[NV], avg[NV], jump[NV];
    const int locL = eL - e0, locR = eR - e0;
    const bool inL = locL >= 0 && locL < ne, inR = eR >= 0 && locR >= 0 && locR < ne;
    {
      const int dn = faceDirOf<D>(lfL), side = faceSideOf<D>(lfL);
      const int base = S.sFaceBase[lfL * NQF + j], stride = strideOf<N, D>(dn);
      if (inL) lineTrace<N, NV, NN>(S.sU + locL * NV * NN, base, stride, S.sLend + side * N, consL);
      else lineTrace<N, NV, NN>(A.Uin + (size_t)eL * NV * NN, base, stride, S.sLend + side * N, consL);
    }
    int jr = j;
    if (eR >= 0) {
      double consR[NV];
      jr = S.sSeq[rot * NQF + j];
      const int dn = faceDirOf<D>(lfR), side = faceSideOf<D>(lfR);
      const int base = S.sFaceBase[lfR * NQF + jr], stride = strideOf<N, D>(dn);
      if (inR) lineTrace<N, NV, NN>(S.sU + locR * NV * NN, base, stride, S.sLend + side * N, consR);
      else lineTrace<N, NV, NN>(A.Uin + (size_t)eR * NV * NN, base, stride, S.sLend + side * N, consR);
#pragma unroll
      for (int v = 0; v < NV; v++) { avg[v] = (consL[v] + consR[v]) / 2.0; jump[v] = (consR[v] - consL[v]) / 2.0; }  // ViscousFlux.cpp:33-56
    } else {
      double compL[D + 3], compR[D + 3];
      compFromCons<D>(ph, consL, compL);
      const double* dm = A.dummy + (size_t)(faceId - A.nInt) * (D + 3) * NQF + j;
#pragma unroll
      for (int k = 0; k < D + 3; k++) compR[k] = dm[k * NQF];
      bcBoundaryGradientVariable<D>(ph, bc, n, consL, compL, compR, avg, jump);  // BoundaryCondition.cpp:287-297,426-441,...
    }
    if (inL) {
      const int slot = lfL * NQF + j;
#pragma unroll
      for (int v = 0; v < NV; v++) { S.sFlux[(locL * NV + v) * NAQ + slot] = avg[v] * jw; S.sB[(locL * NV + v) * NAQ + slot] = jump[v] * jw; }
#pragma unroll
      for (int d = 0; d < D; d++) S.sN[(locL * D + d) * NAQ + slot] = n[d];
    }
    if (inR) {  // right parent: volume-gradient flux changes sign, interface-gradient flux does not (SpatialDiscrete.cpp:885-906)
      const int slot = lfR * NQF + jr;
#pragma unroll
      for (int v = 0; v < NV; v++) { S.sFlux[(locR * NV + v) * NAQ + slot] = -avg[v] * jw; S.sB[(locR * NV + v) * NAQ + slot] = jump[v] * jw; }
#pragma unroll
      for (int d = 0; d < D; d++) S.sN[(locR * D + d) * NAQ + slot] = n[d];
    }
  }
  __syncthreads();

  // ---- G1 + G3 + G4: G = M^-1 ( A Φ_f − (U ⊗ (J^T)^-1 detJ w) ∇Φ ) ------------------------------------------------------------
  for (int nd = tid; nd < nNodes; nd += kThreads) {
    const int el = nd / NN, q = nd - el * NN;
    const int e = e0 + el;
    double G[NV][D];
#pragma unroll
    for (int v = 0; v < NV; v++)
#pragma unroll
      for (int c = 0; c < D; c++) G[v][c] = 0.0;
#pragma unroll
    for (int dd = 0; dd < D; dd++) {
      const int st = strideOf<N, D>(dd);
      const int id = (q / st) % N, qb = q - id * st;
#pragma unroll
      for (int a = 0; a < N; a++) {
        const int qa = qb + a * st;
        const double dcoef = S.sDm[a * N + id];
        double u[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) u[v] = S.sU[(el * NV + v) * NN + qa];
#pragma unroll
        for (int c = 0; c < D; c++) {
          const double mc = metricAt<D, NN, AFFINE>(A, e, qa, dd, c, S.sWq) * dcoef;
#pragma unroll
          for (int v = 0; v < NV; v++) G[v][c] -= mc * u[v];
        }
      }
    }
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int dn = faceDirOf<D>(f), side = faceSideOf<D>(f);
      const int st = strideOf<N, D>(dn);
      const int id = (q / st) % N;
      const double cf = S.sLend[side * N + id];
      const int slot = f * NQF + S.sNodePt[f * NN + q];
#pragma unroll
      for (int v = 0; v < NV; v++) {
        double a = S.sFlux[(el * NV + v) * NAQ + slot];
        if (br1) a += S.sB[(el * NV + v) * NAQ + slot];   // BR1: single lifting block, G = G_vol + G_lift (TimeIntegration.cpp:208-215)
        a *= cf;
#pragma unroll
        for (int c = 0; c < D; c++) G[v][c] += a * S.sN[(el * D + c) * NAQ + slot];
      }
    }
    const double ijw = invJwAt<D, NN, AFFINE>(A, e, q, S.sWq);
    double* out = A.Gout + ((size_t)e * NG) * NN + q;
#pragma unroll
    for (int v = 0; v < NV; v++)
#pragma unroll
      for (int c = 0; c < D; c++) out[(size_t)(v * D + c) * NN] = G[v][c] * ijw;
  }
}

// lifting factor of a face point:  Σ_a l_a(±1)^2 / (detJ w)(node(a, j))   (trace of the rank-one BR2 lift at its own face)
template <int D, int N, int NN, bool AFFINE>
__device__ __forceinline__ double liftTraceFactor(const StageArgs& A, int e, int base, int stride, const double* lend, const double* sWq) {
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < N; a++) s += lend[a] * lend[a] * invJwAt<D, NN, AFFINE>(A, e, base + a * stride, sWq);
  return s;
}

// =====================================================================================================================
// pass R
// =====================================================================================================================
template <int D, int N, int K, bool AFFINE, int PH>
__global__ void __launch_bounds__(kThreads, 1) nsStageKernel(const __grid_constant__ StageArgs A) {
  using L = NsLayout<D, N, K>;
  constexpr int NV = L::NV, NG = L::NG, NN = L::NN, NQF = L::NQF, NF = L::NF, NAQ = L::NAQ, ITERS = L::ITERS;
  extern __shared__ __align__(16) double smem[];
  NsShared<D, N, K> S; S.carve(smem);
  const int tid = threadIdx.x;
  const int chunk = A.chunkList ? A.chunkList[blockIdx.x] : blockIdx.x;
  const int e0 = chunk * K;
  const int ne = min(K, A.nOwned - e0);
  const int nNodes = ne * NN;
  const Phys<PH> ph(A.phys);
  const bool br2 = A.phys.visc == kBR2;
  {
    const double* src = A.Uin + (size_t)e0 * NV * NN;
    for (int i = tid; i < ne * NV * NN; i += kThreads) S.sU[i] = src[i];
    const double* gsrc = A.Gvol + (size_t)e0 * NG * NN;
    for (int i = tid; i < ne * NG * NN; i += kThreads) S.sG[i] = gsrc[i];
  }
  S.loadTables(*A.tab, tid);
  __syncthreads();

  // ---- R2: Riemann flux minus averaged viscous normal flux at the face points ------------------------------------------------------
  const int f0 = A.chunkFaceOff[chunk], nfc = A.chunkFaceOff[chunk + 1] - f0;
  for (int fp = tid; fp < nfc * NQF; fp += kThreads) {
    const int fi = fp / NQF, j = fp - fi * NQF;
    const int4 rec = __ldg(A.faceRec + f0 + fi);
    const int eL = rec.x, eR = rec.y, faceId = rec.z;
    const int lfL = rec.w & 15, lfR = (rec.w >> 4) & 15, rot = (rec.w >> 8) & 15, bc = (rec.w >> 12) & 15;
    double n[D], jw;
    faceGeometryAt<D, NQF, AFFINE>(A, faceId, j, S.sWf, n, jw);
    double consL[NV], compL[D + 3], gL[NG], jump[NV], Fn[NV];
    const int locL = eL - e0, locR = eR - e0;
    const bool inL = locL >= 0 && locL < ne, inR = eR >= 0 && locR >= 0 && locR < ne;
    double lamL;
    {
      const int dn = faceDirOf<D>(lfL), side = faceSideOf<D>(lfL);
      const int base = S.sFaceBase[lfL * NQF + j], stride = strideOf<N, D>(dn);
      if (inL) { lineTrace<N, NV, NN>(S.sU + locL * NV * NN, base, stride, S.sLend + side * N, consL); lineTrace<N, NG, NN>(S.sG + locL * NG * NN, base, stride, S.sLend + side * N, gL); }
      else { lineTrace<N, NV, NN>(A.Uin + (size_t)eL * NV * NN, base, stride, S.sLend + side * N, consL); lineTrace<N, NG, NN>(A.Gvol + (size_t)eL * NG * NN, base, stride, S.sLend + side * N, gL); }
      lamL = br2 ? liftTraceFactor<D, N, NN, AFFINE>(A, eL, base, stride, S.sLend + side * N, S.sWq) : 0.0;
    }
    const double irL = compFromCons<D>(ph, consL, compL);
    int jr = j;
    if (eR >= 0) {
      double consR[NV], compR[D + 3], gR[NG];
      jr = S.sSeq[rot * NQF + j];
      const int dn = faceDirOf<D>(lfR), side = faceSideOf<D>(lfR);
      const int base = S.sFaceBase[lfR * NQF + jr], stride = strideOf<N, D>(dn);
      if (inR) { lineTrace<N, NV, NN>(S.sU + locR * NV * NN, base, stride, S.sLend + side * N, consR); lineTrace<N, NG, NN>(S.sG + locR * NG * NN, base, stride, S.sLend + side * N, gR); }
      else { lineTrace<N, NV, NN>(A.Uin + (size_t)eR * NV * NN, base, stride, S.sLend + side * N, consR); lineTrace<N, NG, NN>(A.Gvol + (size_t)eR * NG * NN, base, stride, S.sLend + side * N, gR); }
      const double irR = compFromCons<D>(ph, consR, compR);
#pragma unroll
      for (int v = 0; v < NV; v++) jump[v] = (consR[v] - consL[v]) / 2.0 * jw;
      if (br2) {  // trace of (G_vol + G_f) on both sides, VariableConvertor.cpp:674-688
        const double lamR = liftTraceFactor<D, N, NN, AFFINE>(A, eR, base, stride, S.sLend + side * N, S.sWq);
#pragma unroll
        for (int v = 0; v < NV; v++)
#pragma unroll
          for (int c = 0; c < D; c++) { gL[v * D + c] += lamL * n[c] * jump[v]; gR[v * D + c] += lamR * n[c] * jump[v]; }
      }
      double pL[NG], pR[NG], va[NV], vb[NV];
      primGradFromConsGrad<D>(ph, consL, compL, gL, pL);
      primGradFromConsGrad<D>(ph, consR, compR, gR, pR);
      convFlux<D>(ph, n, consL, compL, irL, consR, compR, irR, Fn);
      viscNormalFlux<D>(ph, n, compL, pL, va);   // calculateViscousFlux, ViscousFlux.cpp:139-153: average of both sides
      viscNormalFlux<D>(ph, n, compR, pR, vb);
#pragma unroll
      for (int v = 0; v < NV; v++) Fn[v] -= (va[v] + vb[v]) / 2.0;
    } else {
      double compR[D + 3], b[D + 3], volCons[NV], intCons[NV];
      const double* dm = A.dummy + (size_t)(faceId - A.nInt) * (D + 3) * NQF + j;
#pragma unroll
      for (int k = 0; k < D + 3; k++) compR[k] = dm[k * NQF];
      bcBoundaryGradientVariable<D>(ph, bc, n, consL, compL, compR, volCons, intCons);
#pragma unroll
      for (int v = 0; v < NV; v++) jump[v] = intCons[v] * jw;
      if (br2) {
#pragma unroll
        for (int v = 0; v < NV; v++)
#pragma unroll
          for (int c = 0; c < D; c++) gL[v * D + c] += lamL * n[c] * jump[v];
      }
      double pL[NG], gb[NG], va[NV], vb[NV];
      primGradFromConsGrad<D>(ph, consL, compL, gL, pL);        // from the UNMODIFIED interior trace (SpatialDiscrete.cpp:792-796)
      bcBoundaryVariable<D>(ph, bc, n, compL, compR, b);
      convNormalFlux<D>(ph, n, b, Fn);                          // :797-803
      // modifyBoundaryVariable (BoundaryCondition.cpp:299-307,443-452,490-501,535-546): walls overwrite the interior computational
      // state; boundary gradient = interior primitive gradient, adiabatic walls drop the temperature gradient
      if (bcIsWall(bc)) {
#pragma unroll
        for (int k = 0; k < D + 3; k++) compL[k] = b[k];
      }
#pragma unroll
      for (int k = 0; k < NG; k++) gb[k] = pL[k];
      if (bc == kAdiabaticSlipWall || bc == kAdiabaticNonSlipWall) {
#pragma unroll
        for (int d = 0; d < D; d++) gb[(D + 1) * D + d] = 0.0;
      }
      viscNormalFlux<D>(ph, n, compL, pL, va);
      viscNormalFlux<D>(ph, n, b, gb, vb);
#pragma unroll
      for (int v = 0; v < NV; v++) Fn[v] -= (va[v] + vb[v]) / 2.0;
    }
    if (inL) {
      const int slot = lfL * NQF + j;
#pragma unroll
      for (int v = 0; v < NV; v++) { S.sFlux[(locL * NV + v) * NAQ + slot] = Fn[v] * jw; S.sB[(locL * NV + v) * NAQ + slot] = jump[v]; }
#pragma unroll
      for (int d = 0; d < D; d++) S.sN[(locL * D + d) * NAQ + slot] = n[d];
    }
    if (inR) {
      const int slot = lfR * NQF + jr;
#pragma unroll
      for (int v = 0; v < NV; v++) { S.sFlux[(locR * NV + v) * NAQ + slot] = -Fn[v] * jw; S.sB[(locR * NV + v) * NAQ + slot] = jump[v]; }
#pragma unroll
      for (int d = 0; d < D; d++) S.sN[(locR * D + d) * NAQ + slot] = n[d];
    }
  }
  __syncthreads();

  // ---- R1: convective minus viscous flux at the nodes; total gradient = G_vol + Σ_f G_f (BR2) -----------------------------------------
  double R[ITERS][NV], Fv[ITERS][NG], velp[ITERS][D + 1];
#pragma unroll
  for (int it = 0; it < ITERS; it++) {
    const int nd = tid + it * kThreads;
#pragma unroll
    for (int v = 0; v < NV; v++) R[it][v] = 0.0;
    if (nd < nNodes) {
      const int el = nd / NN, q = nd - el * NN;
      double cons[NV], comp[D + 3], g[NG], gp[NG];
#pragma unroll
      for (int v = 0; v < NV; v++) cons[v] = S.sU[(el * NV + v) * NN + q];
#pragma unroll
      for (int r = 0; r < NG; r++) g[r] = S.sG[(el * NG + r) * NN + q];
      if (br2) {
        const double ijw = invJwAt<D, NN, AFFINE>(A, e0 + el, q, S.sWq);
#pragma unroll
        for (int f = 0; f < NF; f++) {
          const int dn = faceDirOf<D>(f), side = faceSideOf<D>(f);
          const int st = strideOf<N, D>(dn);
          const int id = (q / st) % N;
          const double cf = S.sLend[side * N + id] * ijw;
          const int slot = f * NQF + S.sNodePt[f * NN + q];
#pragma unroll
          for (int v = 0; v < NV; v++) {
            const double a = cf * S.sB[(el * NV + v) * NAQ + slot];
#pragma unroll
            for (int c = 0; c < D; c++) g[v * D + c] += a * S.sN[(el * D + c) * NAQ + slot];
          }
        }
      }
      if (A.mode == 3) {  // diagnostics: total gradient G_vol + Σ_f G_f at the nodes (variable_gradient_basis_function_coefficient_)
        double* out = A.Gout + ((size_t)(e0 + el) * NG) * NN + q;
#pragma unroll
        for (int r = 0; r < NG; r++) out[(size_t)r * NN] = g[r];
      }
      compFromCons<D>(ph, cons, comp);
      primGradFromConsGrad<D>(ph, cons, comp, g, gp);
      viscRawFlux<D>(ph, comp, gp, Fv[it]);
#pragma unroll
      for (int c = 0; c < D; c++) velp[it][c] = comp[1 + c];
      velp[it][D] = comp[D + 2];
    }
  }
  if (A.mode == 3) return;
#pragma unroll
  for (int dd = 0; dd < D; dd++) {
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
      const int nd = tid + it * kThreads;
      if (nd < nNodes) {
        const int el = nd / NN, q = nd - el * NN;
        double cons[NV], comp[D + 3], m[D], Ft[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) cons[v] = S.sU[(el * NV + v) * NN + q];
#pragma unroll
        for (int c = 0; c < D; c++) comp[1 + c] = velp[it][c];
        comp[D + 2] = velp[it][D];
#pragma unroll
        for (int c = 0; c < D; c++) m[c] = metricAt<D, NN, AFFINE>(A, e0 + el, q, dd, c, S.sWq);
        contravariantFlux<D>(ph, cons, comp, m, Ft);
#pragma unroll
        for (int v = 0; v < NV; v++) {
          double s = 0.0;
#pragma unroll
          for (int c = 0; c < D; c++) s += Fv[it][v * D + c] * m[c];
          S.sF[(el * NV + v) * NN + q] = Ft[v] - s;   // SpatialDiscrete.cpp:216-232: (F_c − F_v)ᵀ (J^T)^-1 detJ w
        }
      }
    }
    __syncthreads();
    const int st = strideOf<N, D>(dd);
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
      const int nd = tid + it * kThreads;
      if (nd < nNodes) {
        const int el = nd / NN, q = nd - el * NN;
        const int id = (q / st) % N;
        const double* f = S.sF + (el * NV) * NN + (q - id * st);
#pragma unroll
        for (int a = 0; a < N; a++) {
          const double dm = S.sDm[a * N + id];
#pragma unroll
          for (int v = 0; v < NV; v++) R[it][v] += f[v * NN + a * st] * dm;
        }
      }
    }
    if (dd + 1 < D) __syncthreads();
  }

  // ---- R3 (face part) + R4 -------------------------------------------------------------------------------------------------------
#pragma unroll
  for (int it = 0; it < ITERS; it++) {
    const int nd = tid + it * kThreads;
    if (nd < nNodes) {
      const int el = nd / NN, q = nd - el * NN;
#pragma unroll
      for (int f = 0; f < NF; f++) {
        const int dn = faceDirOf<D>(f), side = faceSideOf<D>(f);
        const int st = strideOf<N, D>(dn);
        const int id = (q / st) % N;
        const double cf = S.sLend[side * N + id];
        const int slot = f * NQF + S.sNodePt[f * NN + q];
#pragma unroll
        for (int v = 0; v < NV; v++) R[it][v] -= cf * S.sFlux[(el * NV + v) * NAQ + slot];
      }
      double cons[NV];
#pragma unroll
      for (int v = 0; v < NV; v++) cons[v] = S.sU[(el * NV + v) * NN + q];
      const double ijw = invJwAt<D, NN, AFFINE>(A, e0 + el, q, S.sWq);
      if (A.phys.source == kBoussinesq) {
        double comp[D + 3];
        compFromCons<D>(ph, cons, comp);
        R[it][D] += boussinesqSource<D>(ph, comp) / ijw;
      }
      const size_t g = ((size_t)(e0 + el) * NV) * NN + q;
      if (A.mode == 0) {
#pragma unroll
        for (int v = 0; v < NV; v++) {
          double u = A.aCur * cons[v] + A.bdt * (R[it][v] * ijw);
          if (A.aLast != 0.0) u += A.aLast * A.Ulast[g + (size_t)v * NN];
          A.Uout[g + (size_t)v * NN] = u;
        }
      } else {
#pragma unroll
        for (int v = 0; v < NV; v++) A.Uout[g + (size_t)v * NN] = A.mode == 1 ? R[it][v] * ijw : R[it][v];
      }
    }
  }

  // ---- K: relative error (same reduction as the Euler kernel) -------------------------------------------------------------------------
  if (A.normPartial != nullptr) {
    double* bufA = S.sF;
    double* bufB = S.sFlux;
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
      const int nd = tid + it * kThreads;
      if (nd < nNodes) {
        const int el = nd / NN, q = nd - el * NN;
#pragma unroll
        for (int v = 0; v < NV; v++) bufA[(el * NV + v) * NN + q] = R[it][v];
      }
    }
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) acc[v] = 0.0;
#pragma unroll
    for (int dd = 0; dd < D; dd++) {
      __syncthreads();
      const double* in = (dd & 1) ? bufB : bufA;
      double* out = (dd & 1) ? bufA : bufB;
      const int st = strideOf<N, D>(dd);
      for (int nd = tid; nd < nNodes; nd += kThreads) {
        const int el = nd / NN, q = nd - el * NN;
        const int id = (q / st) % N, qb = q - id * st;
#pragma unroll
        for (int v = 0; v < NV; v++) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < N; a++) s += S.sK1[a * N + id] * in[(el * NV + v) * NN + qb + a * st];
          if (dd == D - 1) acc[v] += fabs(s); else out[(el * NV + v) * NN + q] = s;
        }
      }
    }
    __syncthreads();
    double* red = S.sU;
#pragma unroll
    for (int v = 0; v < NV; v++) {
      double s = acc[v];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if ((tid & 31) == 0) red[(tid >> 5) * NV + v] = s;
    }
    __syncthreads();
    if (tid < NV) {
      double s = 0.0;
      for (int w = 0; w < kThreads / 32; w++) s += red[w * NV + tid];
      A.normPartial[(size_t)chunk * NV + tid] = s / NN;
    }
  }
}

}  // namespace sdg
