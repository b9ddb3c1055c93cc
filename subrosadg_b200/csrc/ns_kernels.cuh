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
//
// v2 (round 1): the chunk's states / gradients / metric and face records arrive by TMA bulk copies; the face-trace gathers
// walk the normal line in a per-lane rotated order (bank-conflict free for N = 4); the volume fluxes of all D directions
// are written once into the (dead) gradient tile, so the pass needs one barrier instead of D; affine meshes keep one
// normal per (element, face) and a precomputed lifting table, so no division is left in the face loop besides 1/rho.
#pragma once
#include "tensor_kernels.cuh"

namespace sdg {

template <int D, int N, int K, bool AFFINE, bool WITHG, int TH = kThreads>
struct NsLayout {
  static constexpr int NV = D + 2, NG = NV * D, NN = Pow<N, D>::v, NQF = NN / N, NF = 2 * D, NAQ = NF * NQF;
  static constexpr int REC = (D * D + 2) & ~1;
  static constexpr int MAXF = K * NF;
  static constexpr int oU = 0;                                      // [K][NV][NN]
  static constexpr int oG = oU + K * NV * NN;                       // [K][NG][NN]   gradient tile (pass R), later the volume fluxes
  static constexpr int oFlux = oG + (WITHG ? K * NG * NN : 0);      // [K][NV][NAQ]  flux slots (pass G: {U} |J|w slots)
  static constexpr int oB = oFlux + K * NV * NAQ;                   // [K][NV][NAQ]  jump slots  ½(U_R−U_L)|J|w  /  (U_b−U_L)|J|w
  static constexpr int oN = oB + K * NV * NAQ;                      // affine: [K][NF][D]; curved: [K][D][NAQ]
  static constexpr int nN = AFFINE ? ((K * NF * D + 1) & ~1) : K * D * NAQ;
  static constexpr int oTab = oN + nN;                              // Dm, K1, Lend, Wq, InvWq, Wf, LiftC
  static constexpr int nTab = (2 * N * N + 2 * N + 2 * NN + NQF + 1) & ~1;
  static constexpr int oGeoE = oTab + nTab;                         // affine: [K][REC]
  static constexpr int oInvDet = oGeoE + (AFFINE ? K * REC : 0);    // affine: [K]
  static constexpr int oCf = oInvDet + ((K + 1) & ~1);              // affine: [MAXF][kCF]
  static constexpr int oRec = oCf + (AFFINE ? MAXF * kCF : 0);      // [MAXF] int4
  static constexpr int nDoubles = oRec + MAXF * 2;
  static constexpr int nBytesTab = NF * NQF + 4 * NQF + NF * NN;
  static constexpr size_t bytes = sizeof(double) * nDoubles + ((nBytesTab + 15) / 16) * 16;
  static constexpr int ITERS = (K * NN + TH - 1) / TH;
};

template <int D, int N, int K, bool AFFINE, bool WITHG, int TH = kThreads>
struct NsShared {
  using L = NsLayout<D, N, K, AFFINE, WITHG, TH>;
  double *sU, *sG, *sFlux, *sB, *sN, *sDm, *sK1, *sLend, *sWq, *sInvWq, *sWf, *sGeoE, *sInvDet, *sCf;
  const int4* sRec;
  unsigned char *sFaceBase, *sSeq, *sNodePt;
  __device__ __forceinline__ void carve(double* smem) {
    sU = smem + L::oU; sG = smem + L::oG; sFlux = smem + L::oFlux; sB = smem + L::oB; sN = smem + L::oN;
    sDm = smem + L::oTab; sK1 = sDm + N * N; sLend = sK1 + N * N; sWq = sLend + 2 * N; sInvWq = sWq + L::NN; sWf = sInvWq + L::NN;
    sGeoE = smem + L::oGeoE; sInvDet = smem + L::oInvDet; sCf = smem + L::oCf;
    sRec = reinterpret_cast<const int4*>(smem + L::oRec);
    sFaceBase = reinterpret_cast<unsigned char*>(smem + L::nDoubles); sSeq = sFaceBase + L::NF * L::NQF; sNodePt = sSeq + 4 * L::NQF;
  }
  __device__ __forceinline__ void loadTables(const TensorDev& T, int tid) {
    for (int i = tid; i < N * N; i += TH) { sDm[i] = T.Dm[i]; sK1[i] = T.K1[i]; }
    for (int i = tid; i < 2 * N; i += TH) sLend[i] = T.Lend[i];
    for (int i = tid; i < L::NN; i += TH) { const double w = T.wq[i]; sWq[i] = w; sInvWq[i] = 1.0 / w; }
    for (int i = tid; i < L::NQF; i += TH) sWf[i] = T.wf[i];
    for (int i = tid; i < L::NF * L::NQF; i += TH) sFaceBase[i] = (unsigned char)T.faceBase[i];
    for (int i = tid; i < 4 * L::NQF; i += TH) sSeq[i] = (unsigned char)T.seq[i];
    for (int i = tid; i < L::NF * L::NN; i += TH) sNodePt[i] = T.nodeFacePt[i];
  }
};

// entry [dd][c] of (J^T)^-1 detJ w at node q of element e (curved meshes: read from HBM / L2)
template <int D, int NN>
__device__ __forceinline__ double metricCurved(const StageArgs& A, int e, int q, int dd, int c) {
  return __ldg(A.geoE + ((size_t)e * (D * D) + dd * D + c) * NN + q);
}
// trace of the rank-one BR2 lift at its own face point, without the 1/detJ (affine meshes):  Σ_a l_a(±1)^2 / w(node(a, j))
template <int N>
__device__ __forceinline__ double liftTraceFactorAffine(const double* sInvWq, int base, int stride, const double* lend) {
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < N; a++) s += lend[a] * lend[a] * sInvWq[base + a * stride];
  return s;
}
// lifting factor of a face point on curved meshes:  Σ_a l_a(±1)^2 / (detJ w)(node(a, j))
template <int D, int N, int NN>
__device__ __forceinline__ double liftTraceFactorCurved(const StageArgs& A, int e, int base, int stride, const double* lend) {
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < N; a++) s += lend[a] * lend[a] * __ldg(A.invjw + (size_t)e * NN + base + a * stride);
  return s;
}

// stages the chunk-owned contiguous ranges with TMA bulk copies; returns after the data has landed
template <class SH, int D, int N, int K, bool AFFINE, bool WITHG, int TH = kThreads>
__device__ __forceinline__ void nsStageIn(const StageArgs& A, SH& S, unsigned long long* mbar, int tid, int chunk, int e0, int ne, int f0, int nfc) {
  using L = NsLayout<D, N, K, AFFINE, WITHG>;
  constexpr int NV = L::NV, NG = L::NG, NN = L::NN;
  const unsigned bytesU = (unsigned)(ne * NV * NN * sizeof(double));
  const unsigned bytesG = WITHG ? (unsigned)(ne * NG * NN * sizeof(double)) : 0u;
  const bool bulkU = (bytesU & 15u) == 0, bulkG = (bytesG & 15u) == 0;
  if (tid == 0) mbarInit(mbar, 1);
  __syncthreads();
  if (tid == 0) {
    unsigned total = (unsigned)(nfc * sizeof(int4)) + (bulkU ? bytesU : 0u) + (bulkG ? bytesG : 0u);
    if constexpr (AFFINE) total += (unsigned)(ne * L::REC * sizeof(double)) + (unsigned)(nfc * kCF * sizeof(double));
    mbarExpectTx(mbar, total);
    if (bulkU) bulkLoad(S.sU, A.Uin + (size_t)e0 * NV * NN, bytesU, mbar);
    if constexpr (WITHG) { if (bulkG) bulkLoad(S.sG, A.Gvol + (size_t)e0 * NG * NN, bytesG, mbar); }
    bulkLoad(const_cast<int4*>(S.sRec), A.faceRec + f0, (unsigned)(nfc * sizeof(int4)), mbar);
    if constexpr (AFFINE) {
      bulkLoad(S.sGeoE, A.geoE + (size_t)e0 * L::REC, (unsigned)(ne * L::REC * sizeof(double)), mbar);
      bulkLoad(S.sCf, A.cfGeo + (size_t)f0 * kCF, (unsigned)(nfc * kCF * sizeof(double)), mbar);
    }
  }
  if (!bulkU) { const double* src = A.Uin + (size_t)e0 * NV * NN; for (int i = tid; i < ne * NV * NN; i += TH) S.sU[i] = src[i]; }
  if constexpr (WITHG) { if (!bulkG) { const double* src = A.Gvol + (size_t)e0 * NG * NN; for (int i = tid; i < ne * NG * NN; i += TH) S.sG[i] = src[i]; } }
  S.loadTables(*A.tab, tid);
  mbarWait(mbar, 0);
  if constexpr (AFFINE) { if (tid < ne) S.sInvDet[tid] = 1.0 / S.sGeoE[tid * L::REC + D * D]; }
  __syncthreads();
}

// =====================================================================================================================
// pass G
// =====================================================================================================================
#ifndef SDG_NSG_MINB
#define SDG_NSG_MINB 3
#endif
template <int D, int N, int K, bool AFFINE>
__global__ void __launch_bounds__(kThreads, SDG_NSG_MINB) nsGradKernel(const __grid_constant__ StageArgs A) {
  using L = NsLayout<D, N, K, AFFINE, false>;
  constexpr int NV = L::NV, NG = L::NG, NN = L::NN, NQF = L::NQF, NF = L::NF, NAQ = L::NAQ;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) unsigned long long mbar;
  NsShared<D, N, K, AFFINE, false> S; S.carve(smem);
  const int tid = threadIdx.x;
  const int chunk = A.chunkList ? A.chunkList[blockIdx.x] : blockIdx.x;
  const int e0 = chunk * K;
  const int ne = min(K, A.nOwned - e0);
  const int nNodes = ne * NN;
  const Phys<0> ph(A.phys);
  const bool br1 = A.phys.visc == kBR1;
  const int f0 = A.chunkFaceOff[chunk], nfc = A.chunkFaceOff[chunk + 1] - f0;
  nsStageIn<decltype(S), D, N, K, AFFINE, false>(A, S, &mbar, tid, chunk, e0, ne, f0, nfc);

  // ---- G2: {U} n |J|w (volume-gradient flux) and ½(U_R−U_L) n |J|w (interface-gradient flux) at the face points -----------
  for (int fp = tid; fp < nfc * NQF; fp += kThreads) {
    const int fi = fp / NQF, j = fp - fi * NQF;
    const int4 rec = S.sRec[fi];
    const int eL = rec.x, eR = rec.y, faceId = rec.z;
    const int lfL = rec.w & 15, lfR = (rec.w >> 4) & 15, rot = (rec.w >> 8) & 15, bc = (rec.w >> 12) & 15;
    const int rl = (j / N + fi) % N;
    double n[D], jw;
    if constexpr (AFFINE) {
      const double* g = S.sCf + fi * kCF;
#pragma unroll
      for (int d = 0; d < D; d++) n[d] = g[d];
      jw = g[D] * S.sWf[j];
    } else {
      const double* g = A.geoF + (size_t)faceId * (D + 1) * NQF + j;
#pragma unroll
      for (int d = 0; d < D; d++) n[d] = __ldg(g + d * NQF);
      jw = __ldg(g + D * NQF);
    }
    double consL[NV], avg[NV], jump[NV];
    const int locL = eL - e0, locR = eR - e0;
    const bool inL = locL >= 0 && locL < ne, inR = eR >= 0 && locR >= 0 && locR < ne;
    {
      const int dn = faceDirOf<D>(lfL), side = faceSideOf<D>(lfL);
      const int base = S.sFaceBase[lfL * NQF + j], stride = strideOf<N, D>(dn);
      if (inL) lineTraceRot<N, NV, NN>(S.sU + locL * NV * NN, base, stride, S.sLend + side * N, rl, consL);
      else lineTraceGlobal<N, NV, NN>(A.Uin + (size_t)eL * NV * NN, base, stride, S.sLend + side * N, consL);
    }
    int jr = j;
    if (eR >= 0) {
      double consR[NV];
      jr = S.sSeq[rot * NQF + j];
      const int dn = faceDirOf<D>(lfR), side = faceSideOf<D>(lfR);
      const int base = S.sFaceBase[lfR * NQF + jr], stride = strideOf<N, D>(dn);
      if (inR) lineTraceRot<N, NV, NN>(S.sU + locR * NV * NN, base, stride, S.sLend + side * N, rl, consR);
      else lineTraceGlobal<N, NV, NN>(A.Uin + (size_t)eR * NV * NN, base, stride, S.sLend + side * N, consR);
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
      if constexpr (AFFINE) {
        if (j == 0) {
#pragma unroll
          for (int d = 0; d < D; d++) S.sN[(locL * NF + lfL) * D + d] = n[d];
        }
      } else {
#pragma unroll
        for (int d = 0; d < D; d++) S.sN[(locL * D + d) * NAQ + slot] = n[d];
      }
    }
    if (inR) {  // right parent: volume-gradient flux changes sign, interface-gradient flux does not (SpatialDiscrete.cpp:885-906)
      const int slot = lfR * NQF + jr;
#pragma unroll
      for (int v = 0; v < NV; v++) { S.sFlux[(locR * NV + v) * NAQ + slot] = -avg[v] * jw; S.sB[(locR * NV + v) * NAQ + slot] = jump[v] * jw; }
      if constexpr (AFFINE) {
        if (j == 0) {
#pragma unroll
          for (int d = 0; d < D; d++) S.sN[(locR * NF + lfR) * D + d] = n[d];
        }
      } else {
#pragma unroll
        for (int d = 0; d < D; d++) S.sN[(locR * D + d) * NAQ + slot] = n[d];
      }
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
      if constexpr (AFFINE) {
        // metric = g[dd][c] * w(node): contract the line first, then the D x D metric
        double t[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) t[v] = 0.0;
#pragma unroll
        for (int a = 0; a < N; a++) {
          const int qa = qb + a * st;
          const double dw = S.sDm[a * N + id] * S.sWq[qa];
#pragma unroll
          for (int v = 0; v < NV; v++) t[v] += dw * S.sU[(el * NV + v) * NN + qa];
        }
#pragma unroll
        for (int c = 0; c < D; c++) {
          const double g = S.sGeoE[el * L::REC + dd * D + c];
#pragma unroll
          for (int v = 0; v < NV; v++) G[v][c] -= g * t[v];
        }
      } else {
#pragma unroll
        for (int a = 0; a < N; a++) {
          const int qa = qb + a * st;
          const double dcoef = S.sDm[a * N + id];
          double u[NV];
#pragma unroll
          for (int v = 0; v < NV; v++) u[v] = S.sU[(el * NV + v) * NN + qa];
#pragma unroll
          for (int c = 0; c < D; c++) {
            const double mc = metricCurved<D, NN>(A, e, qa, dd, c) * dcoef;
#pragma unroll
            for (int v = 0; v < NV; v++) G[v][c] -= mc * u[v];
          }
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
      double nn[D];
#pragma unroll
      for (int c = 0; c < D; c++) nn[c] = AFFINE ? S.sN[(el * NF + f) * D + c] : S.sN[(el * D + c) * NAQ + slot];
#pragma unroll
      for (int v = 0; v < NV; v++) {
        double a = S.sFlux[(el * NV + v) * NAQ + slot];
        if (br1) a += S.sB[(el * NV + v) * NAQ + slot];   // BR1: single lifting block, G = G_vol + G_lift (TimeIntegration.cpp:208-215)
        a *= cf;
#pragma unroll
        for (int c = 0; c < D; c++) G[v][c] += a * nn[c];
      }
    }
    double ijw;
    if constexpr (AFFINE) ijw = S.sInvDet[el] * S.sInvWq[q];
    else ijw = __ldg(A.invjw + (size_t)e * NN + q);
    double* out = A.Gout + ((size_t)e * NG) * NN + q;
#pragma unroll
    for (int v = 0; v < NV; v++)
#pragma unroll
      for (int c = 0; c < D; c++) out[(size_t)(v * D + c) * NN] = G[v][c] * ijw;
  }
}

// =====================================================================================================================
// pass R
// =====================================================================================================================
#ifndef SDG_NSR_MINB
#define SDG_NSR_MINB 2
#endif
template <int D, int N, int K, bool AFFINE, int PH, int TH = kThreads>
__global__ void __launch_bounds__(TH, SDG_NSR_MINB) nsStageKernel(const __grid_constant__ StageArgs A) {
  using L = NsLayout<D, N, K, AFFINE, true, TH>;
  constexpr int NV = L::NV, NG = L::NG, NN = L::NN, NQF = L::NQF, NF = L::NF, NAQ = L::NAQ, ITERS = L::ITERS;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) unsigned long long mbar;
  NsShared<D, N, K, AFFINE, true, TH> S; S.carve(smem);
  const int tid = threadIdx.x;
  const int chunk = A.chunkList ? A.chunkList[blockIdx.x] : blockIdx.x;
  const int e0 = chunk * K;
  const int ne = min(K, A.nOwned - e0);
  const int nNodes = ne * NN;
  const Phys<PH> ph(A.phys);
  const bool br2 = A.phys.visc == kBR2;
  const bool av = A.phys.av != 0;   // Euler + artificial viscosity: G is the volume gradient, the viscous terms are eps * grad(U)
  constexpr int NB = 1 << D;
  const int f0 = A.chunkFaceOff[chunk], nfc = A.chunkFaceOff[chunk + 1] - f0;
  if (A.mode == 0 && A.aLast != 0.0) {   // U_last is consumed at the very end: pull its lines into L2 now
    const char* p = reinterpret_cast<const char*>(A.Ulast + (size_t)e0 * NV * NN);
    const int bytes = ne * NV * NN * (int)sizeof(double);
    for (int o = tid * 128; o < bytes; o += TH * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
  }
  nsStageIn<decltype(S), D, N, K, AFFINE, true, TH>(A, S, &mbar, tid, chunk, e0, ne, f0, nfc);

  // ---- R2: Riemann flux minus averaged viscous normal flux at the face points ------------------------------------------------------
  for (int fp = tid; fp < nfc * NQF; fp += TH) {
    const int fi = fp / NQF, j = fp - fi * NQF;
    const int4 rec = S.sRec[fi];
    const int eL = rec.x, eR = rec.y, faceId = rec.z;
    const int lfL = rec.w & 15, lfR = (rec.w >> 4) & 15, rot = (rec.w >> 8) & 15, bc = (rec.w >> 12) & 15;
    const int rl = (j / N + fi) % N;
    double n[D], jw, invDetL = 0.0, invDetR = 0.0;
    if constexpr (AFFINE) {
      const double* g = S.sCf + fi * kCF;
#pragma unroll
      for (int d = 0; d < D; d++) n[d] = g[d];
      jw = g[D] * S.sWf[j];
      invDetL = g[D + 1]; invDetR = g[D + 2];
    } else {
      const double* g = A.geoF + (size_t)faceId * (D + 1) * NQF + j;
#pragma unroll
      for (int d = 0; d < D; d++) n[d] = __ldg(g + d * NQF);
      jw = __ldg(g + D * NQF);
    }
    const int locL = eL - e0, locR = eR - e0;
    const bool inL = locL >= 0 && locL < ne, inR = eR >= 0 && locR >= 0 && locR < ne;
    const int dnL = faceDirOf<D>(lfL), sideL = faceSideOf<D>(lfL);
    const int baseL = S.sFaceBase[lfL * NQF + j], strideL = strideOf<N, D>(dnL);
    double consL[NV], compL[D + 3], jump[NV], Fn[NV], va[NV];
    if (inL) lineTraceRot<N, NV, NN>(S.sU + locL * NV * NN, baseL, strideL, S.sLend + sideL * N, rl, consL);
    else lineTraceGlobal<N, NV, NN>(A.Uin + (size_t)eL * NV * NN, baseL, strideL, S.sLend + sideL * N, consL);
    const double irL = compFromCons<D>(ph, consL, compL);
    double lamL = 0.0;
    if (br2) {
      if constexpr (AFFINE) lamL = liftTraceFactorAffine<N>(S.sInvWq, baseL, strideL, S.sLend + sideL * N) * invDetL;
      else lamL = liftTraceFactorCurved<D, N, NN>(A, eL, baseL, strideL, S.sLend + sideL * N);
    }
    int jr = j;
    if (eR >= 0) {
      double consR[NV], compR[D + 3];
      jr = S.sSeq[rot * NQF + j];
      const int dnR = faceDirOf<D>(lfR), sideR = faceSideOf<D>(lfR);
      const int baseR = S.sFaceBase[lfR * NQF + jr], strideR = strideOf<N, D>(dnR);
      if (inR) lineTraceRot<N, NV, NN>(S.sU + locR * NV * NN, baseR, strideR, S.sLend + sideR * N, rl, consR);
      else lineTraceGlobal<N, NV, NN>(A.Uin + (size_t)eR * NV * NN, baseR, strideR, S.sLend + sideR * N, consR);
      const double irR = compFromCons<D>(ph, consR, compR);
#pragma unroll
      for (int v = 0; v < NV; v++) jump[v] = (consR[v] - consL[v]) / 2.0 * jw;
      convFlux<D>(ph, n, consL, compL, irL, consR, compR, irR, Fn);
      {  // left side: trace of (G_vol + G_f), VariableConvertor.cpp:674-688, then its normal viscous flux
        double g[NG], gp[NG];
        if (inL) lineTraceRot<N, NG, NN>(S.sG + locL * NG * NN, baseL, strideL, S.sLend + sideL * N, rl, g);
        else lineTraceGlobal<N, NG, NN>(A.Gvol + (size_t)eL * NG * NN, baseL, strideL, S.sLend + sideL * N, g);
        if (br2) {
#pragma unroll
          for (int v = 0; v < NV; v++)
#pragma unroll
            for (int c = 0; c < D; c++) g[v * D + c] += lamL * n[c] * jump[v];
        }
        if (av) {   // calculateArtificialViscousNormalFlux, ViscousFlux.cpp:126-136
          const double eps = avAt<NB>(A.avTabF + (size_t)(lfL * NQF + j) * NB, A.avElem + (size_t)eL * NB);
#pragma unroll
          for (int v = 0; v < NV; v++) {
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < D; c++) t += g[v * D + c] * n[c];
            va[v] = eps * t;
          }
        } else {
          primGradFromConsGrad<D>(ph, consL, compL, g, gp);
          viscNormalFlux<D>(ph, n, compL, gp, va);   // calculateViscousFlux, ViscousFlux.cpp:139-153: average of both sides
        }
      }
      {
        double g[NG], gp[NG], vb[NV];
        if (inR) lineTraceRot<N, NG, NN>(S.sG + locR * NG * NN, baseR, strideR, S.sLend + sideR * N, rl, g);
        else lineTraceGlobal<N, NG, NN>(A.Gvol + (size_t)eR * NG * NN, baseR, strideR, S.sLend + sideR * N, g);
        if (br2) {
          double lamR;
          if constexpr (AFFINE) lamR = liftTraceFactorAffine<N>(S.sInvWq, baseR, strideR, S.sLend + sideR * N) * invDetR;
          else lamR = liftTraceFactorCurved<D, N, NN>(A, eR, baseR, strideR, S.sLend + sideR * N);
#pragma unroll
          for (int v = 0; v < NV; v++)
#pragma unroll
            for (int c = 0; c < D; c++) g[v * D + c] += lamR * n[c] * jump[v];
        }
        if (av) {
          // the reference takes the right viscosity at the right face's own point j, not at the matching point sequence[j] that the
          // gradient column uses (right_quadrature_node_artificial_viscosity(j), SpatialDiscrete.cpp:714-719): reproduced as it is
          const double eps = avAt<NB>(A.avTabF + (size_t)(lfR * NQF + j) * NB, A.avElem + (size_t)eR * NB);
#pragma unroll
          for (int v = 0; v < NV; v++) {
            double t = 0.0;
#pragma unroll
            for (int c = 0; c < D; c++) t += g[v * D + c] * n[c];
            vb[v] = eps * t;
          }
        } else {
          primGradFromConsGrad<D>(ph, consR, compR, g, gp);
          viscNormalFlux<D>(ph, n, compR, gp, vb);
        }
#pragma unroll
        for (int v = 0; v < NV; v++) Fn[v] -= (va[v] + vb[v]) / 2.0;   // calculateArtificialViscousFlux averages as well (ViscousFlux.cpp:172-186)
      }
    } else {
      double compR[D + 3], b[D + 3], volCons[NV], intCons[NV], g[NG];
      const double* dm = A.dummy + (size_t)(faceId - A.nInt) * (D + 3) * NQF + j;
#pragma unroll
      for (int k = 0; k < D + 3; k++) compR[k] = dm[k * NQF];
      bcBoundaryGradientVariable<D>(ph, bc, n, consL, compL, compR, volCons, intCons);
#pragma unroll
      for (int v = 0; v < NV; v++) jump[v] = intCons[v] * jw;
      if (inL) lineTraceRot<N, NG, NN>(S.sG + locL * NG * NN, baseL, strideL, S.sLend + sideL * N, rl, g);
      else lineTraceGlobal<N, NG, NN>(A.Gvol + (size_t)eL * NG * NN, baseL, strideL, S.sLend + sideL * N, g);
      if (br2) {
#pragma unroll
        for (int v = 0; v < NV; v++)
#pragma unroll
          for (int c = 0; c < D; c++) g[v * D + c] += lamL * n[c] * jump[v];
      }
      double pL[NG], gb[NG], vb[NV];
      bcBoundaryVariable<D>(ph, bc, n, compL, compR, b);
      convNormalFlux<D>(ph, n, b, Fn);                          // :797-803
      if (av) {   // boundary faces: the interior side alone (SpatialDiscrete.cpp:813-819)
        const double eps = avAt<NB>(A.avTabF + (size_t)(lfL * NQF + j) * NB, A.avElem + (size_t)eL * NB);
#pragma unroll
        for (int v = 0; v < NV; v++) {
          double t = 0.0;
#pragma unroll
          for (int c = 0; c < D; c++) t += g[v * D + c] * n[c];
          Fn[v] -= eps * t;
        }
      } else {
      primGradFromConsGrad<D>(ph, consL, compL, g, pL);        // from the UNMODIFIED interior trace (SpatialDiscrete.cpp:792-796)
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
    }
    if (inL) {
      const int slot = lfL * NQF + j;
#pragma unroll
      for (int v = 0; v < NV; v++) { S.sFlux[(locL * NV + v) * NAQ + slot] = Fn[v] * jw; S.sB[(locL * NV + v) * NAQ + slot] = jump[v]; }
      if constexpr (AFFINE) {
        if (j == 0) {
#pragma unroll
          for (int d = 0; d < D; d++) S.sN[(locL * NF + lfL) * D + d] = n[d];
        }
      } else {
#pragma unroll
        for (int d = 0; d < D; d++) S.sN[(locL * D + d) * NAQ + slot] = n[d];
      }
    }
    if (inR) {
      const int slot = lfR * NQF + jr;
#pragma unroll
      for (int v = 0; v < NV; v++) { S.sFlux[(locR * NV + v) * NAQ + slot] = -Fn[v] * jw; S.sB[(locR * NV + v) * NAQ + slot] = jump[v]; }
      if constexpr (AFFINE) {
        if (j == 0) {
#pragma unroll
          for (int d = 0; d < D; d++) S.sN[(locR * NF + lfR) * D + d] = n[d];
        }
      } else {
#pragma unroll
        for (int d = 0; d < D; d++) S.sN[(locR * D + d) * NAQ + slot] = n[d];
      }
    }
  }
  __syncthreads();

  // ---- R1: convective minus viscous flux at the nodes; total gradient = G_vol + Σ_f G_f (BR2).  The contravariant fluxes of
  //      all D directions overwrite the node's own entries of the gradient tile (dead after this point). ---------------------
  double ijwv[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; it++) {
    const int nd = tid + it * TH;
    ijwv[it] = 0.0;
    if (nd < nNodes) {
      const int el = nd / NN, q = nd - el * NN;
      double cons[NV], comp[D + 3], g[NG], gp[NG], Fv[NG];
#pragma unroll
      for (int v = 0; v < NV; v++) cons[v] = S.sU[(el * NV + v) * NN + q];
#pragma unroll
      for (int r = 0; r < NG; r++) g[r] = S.sG[(el * NG + r) * NN + q];
      double ijw;
      if constexpr (AFFINE) ijw = S.sInvDet[el] * S.sInvWq[q];
      else ijw = __ldg(A.invjw + (size_t)(e0 + el) * NN + q);
      ijwv[it] = ijw;
      if (br2) {
#pragma unroll
        for (int f = 0; f < NF; f++) {
          if (A.mode == 3 && A.faceSel >= 0 && f != A.faceSel) continue;
          const int dn = faceDirOf<D>(f), side = faceSideOf<D>(f);
          const int st = strideOf<N, D>(dn);
          const int id = (q / st) % N;
          const double cf = S.sLend[side * N + id] * ijw;
          const int slot = f * NQF + S.sNodePt[f * NN + q];
          double nn[D];
#pragma unroll
          for (int c = 0; c < D; c++) nn[c] = (AFFINE ? S.sN[(el * NF + f) * D + c] : S.sN[(el * D + c) * NAQ + slot]) * cf;
#pragma unroll
          for (int v = 0; v < NV; v++) {
            const double a = S.sB[(el * NV + v) * NAQ + slot];
#pragma unroll
            for (int c = 0; c < D; c++) g[v * D + c] += a * nn[c];
          }
        }
      }
      if (A.mode == 3) {  // diagnostics: total gradient G_vol + Σ_f G_f at the nodes (variable_gradient_basis_function_coefficient_)
        double* out = A.Gout + ((size_t)(e0 + el) * NG) * NN + q;
#pragma unroll
        for (int r = 0; r < NG; r++) out[(size_t)r * NN] = g[r];
      } else {
        compFromCons<D>(ph, cons, comp);
        if (av) {   // calculateArtificialViscousRawFlux, ViscousFlux.cpp:105-113: eps(q) * volume gradient of the conserved variables
          const double eps = avAt<NB>(A.avTabQ + (size_t)q * NB, A.avElem + (size_t)(e0 + el) * NB);
#pragma unroll
          for (int r = 0; r < NG; r++) Fv[r] = eps * g[r];
        } else {
          primGradFromConsGrad<D>(ph, cons, comp, g, gp);
          viscRawFlux<D>(ph, comp, gp, Fv);
        }
#pragma unroll
        for (int dd = 0; dd < D; dd++) {
          double m[D], Ft[NV];
#pragma unroll
          for (int c = 0; c < D; c++) {
            if constexpr (AFFINE) m[c] = S.sGeoE[el * L::REC + dd * D + c] * S.sWq[q];
            else m[c] = metricCurved<D, NN>(A, e0 + el, q, dd, c);
          }
          contravariantFlux<D>(ph, cons, comp, m, Ft);
#pragma unroll
          for (int v = 0; v < NV; v++) {
            double s = 0.0;
#pragma unroll
            for (int c = 0; c < D; c++) s += Fv[v * D + c] * m[c];
            S.sG[(el * NG + dd * NV + v) * NN + q] = Ft[v] - s;   // SpatialDiscrete.cpp:216-232: (F_c − F_v)ᵀ (J^T)^-1 detJ w
          }
        }
      }
    }
  }
  if (A.mode == 3) return;
  __syncthreads();

  // ---- R3 + R4: R = Q·∇Φ − A·Φ_f, mass inverse, RK update ----------------------------------------------------------------------------
  double R[ITERS][NV];
#pragma unroll
  for (int it = 0; it < ITERS; it++) {
    const int nd = tid + it * TH;
#pragma unroll
    for (int v = 0; v < NV; v++) R[it][v] = 0.0;
    if (nd < nNodes) {
      const int el = nd / NN, q = nd - el * NN;
      const size_t g = ((size_t)(e0 + el) * NV) * NN + q;
      double ul[NV];   // issued now, consumed after the contraction below
      if (A.mode == 0 && A.aLast != 0.0) {
#pragma unroll
        for (int v = 0; v < NV; v++) ul[v] = __ldg(A.Ulast + g + (size_t)v * NN);
      }
#pragma unroll
      for (int dd = 0; dd < D; dd++) {
        const int st = strideOf<N, D>(dd);
        const int id = (q / st) % N;
        const double* f = S.sG + (el * NG + dd * NV) * NN + (q - id * st);
#pragma unroll
        for (int a = 0; a < N; a++) {
          const double dm = S.sDm[a * N + id];
#pragma unroll
          for (int v = 0; v < NV; v++) R[it][v] += f[v * NN + a * st] * dm;
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
        for (int v = 0; v < NV; v++) R[it][v] -= cf * S.sFlux[(el * NV + v) * NAQ + slot];
      }
      double cons[NV];
#pragma unroll
      for (int v = 0; v < NV; v++) cons[v] = S.sU[(el * NV + v) * NN + q];
      const double ijw = ijwv[it];
      if (A.phys.source == kBoussinesq) {
        double comp[D + 3];
        compFromCons<D>(ph, cons, comp);
        R[it][D] += boussinesqSource<D>(ph, comp) / ijw;
      }
      if (A.mode == 0) {
#pragma unroll
        for (int v = 0; v < NV; v++) {
          double u = A.aCur * cons[v] + A.bdt * (R[it][v] * ijw);
          if (A.aLast != 0.0) u += A.aLast * ul[v];
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
    double* bufA = S.sG;      // >= [K][NV][NN]
    double* bufB = (NAQ >= NN) ? S.sFlux : S.sU;   // [K][NV][NAQ] >= [K][NV][NN] in 3-D and for N <= 4 in 2-D; otherwise the (dead) state tile
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
      const int nd = tid + it * TH;
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
      for (int nd = tid; nd < nNodes; nd += TH) {
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
      for (int w = 0; w < TH / 32; w++) s += red[w * NV + tid];
      A.normPartial[(size_t)chunk * NV + tid] = s / NN;
    }
  }
}

}  // namespace sdg
