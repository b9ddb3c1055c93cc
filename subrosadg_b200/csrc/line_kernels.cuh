// line_kernels.cuh — register-blocked Euler stage kernel for hexahedron blocks (sm_100a, fp64).
//
// Same stage as eulerStageKernel (tensor_kernels.cuh): R1–R4 + K of Solver::stepSolver (src/Solver/TimeIntegration.cpp:326-350,
// SpatialDiscrete.cpp:194-266,633-842,1016-1032) in one launch, same collocation representation, same chunk / face-record
// plan.  What changes is the work decomposition, chosen for the latency-bound profile of the node-per-thread kernel
// (profiles/r01_euler_stage_v2_ncu.md):
//
//   * a thread owns one LINE of N nodes along the fastest axis (zeta) of one element: N^2 threads per element, K elements
//     per block.  Its N x Nv state values live in registers (loaded / stored with 16-byte accesses, fully coalesced), so the
//     zeta derivative, the zeta face traces and the zeta lifting never touch shared memory, and every thread carries N
//     independent dependency chains (instruction-level parallelism instead of occupancy);
//   * xi / eta derivatives exchange ONE flux tile per direction through shared memory (16-byte accesses, row pitch padded so
//     that the four i-rows fall into different banks), synchronised with __syncwarp (an element never spans two warps);
//   * every element computes its own 2·D face traces once ([face][var][point] tile); a face interior to the chunk reads both
//     traces from that tile, a face on the chunk boundary gathers the outside parent from global memory (L2);
//   * the Riemann flux overwrites the trace slots it was computed from (each (element, face, point) slot is read and written
//     by exactly one thread — the reference's race-free slot rule, SpatialDiscrete.cpp:406-439,738-744), no atomics;
//   * 1-D operator coefficients (differentiation matrix, end-point interpolation, K1) travel in the kernel parameter block,
//     i.e. in the constant bank: with compile-time indices they are instruction operands, not loads.
#pragma once
#include "tensor_kernels.cuh"

namespace sdg {

// local faces of an axis: hexahedron faces (zeta-,eta-,xi-,xi+,eta+,zeta+) -> axis 0 (xi): {2,3}, 1 (eta): {1,4}, 2 (zeta): {0,5}
__device__ __forceinline__ constexpr int hexFaceOfAxis(int d, int side) { return d == 0 ? (side ? 3 : 2) : d == 1 ? (side ? 4 : 1) : (side ? 5 : 0); }

template <int N, int K>
struct LineLayout {
  static constexpr int D = 3, NV = 5, NL = N * N, NN = N * N * N, NF = 6;
  static constexpr int PQ = N * N + N;          // pitch of an i-row (doubles): consecutive rows start N words (32 bytes) apart in the banks
  static constexpr int NNP = N * PQ;            // padded nodes per variable
  static constexpr int REC = (D * D + 2) & ~1;
  static constexpr int MAXF = K * NF;
  static constexpr int oU = 0;                                  // [2][K][NNP]  exchange tiles of TWO variables (state for the traces, then fluxes)
  static constexpr int oT = oU + 2 * K * NNP;                   // [K][NF][NV][NL] traces, later face fluxes
  static constexpr int oW = oT + K * NF * NV * NL;              // wq[NN], invWq[NN], wf[NL]
  static constexpr int oGeoE = oW + 2 * NN + NL;                // affine: [K][REC]
  static constexpr int oInvDet = oGeoE + K * REC;               // [K]
  static constexpr int oCf = oInvDet + ((K + 1) & ~1);          // affine: [MAXF][kCF]
  static constexpr int oRec = oCf + MAXF * kCF;                 // [MAXF] int4
  static constexpr int oRed = oRec + MAXF * 2;                  // [32][NV] block reduction scratch
  static constexpr int oL = oRed + 32 * NV;                     // [K][NV][NN]  U_last of the chunk, landed by TMA while the stage is computed
  static constexpr int nDoubles = oL + K * NV * NN;
  static constexpr int nBytes = 2 * NF * NL + 8 * NL + NF * NL + 2 * K * NF;   // nat2jf, jf2nat, seq, invseq, faceBase, sEF (int16)
  static constexpr size_t bytes = sizeof(double) * nDoubles + ((nBytes + 15) / 16) * 16;
  static constexpr int THREADS = K * NL;
};

__device__ __forceinline__ double2 ldg2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

#ifndef SDG_LINE_MINB
#define SDG_LINE_MINB 3
#endif
template <int N, int K, bool AFFINE, int PH>
__global__ void __launch_bounds__(K * N * N, (N == 4 && K == 8) ? SDG_LINE_MINB : 1) eulerLineKernel(const __grid_constant__ StageArgs A) {
  static_assert(N % 2 == 0, "16-byte accesses along the line need an even number of nodes");
  using L = LineLayout<N, K>;
  constexpr int D = 3, NV = 5, NL = L::NL, NN = L::NN, NF = 6, PQ = L::PQ, NNP = L::NNP, THREADS = L::THREADS;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) unsigned long long mbar, mbarLast;
  double* sU = smem + L::oU;
  double* sL = smem + L::oL;
  double* sT = smem + L::oT;
  double* sWq = smem + L::oW;
  double* sInvWq = sWq + NN;
  double* sWf = sInvWq + NN;
  double* sGeoE = smem + L::oGeoE;
  double* sInvDet = smem + L::oInvDet;
  double* sCf = smem + L::oCf;
  const int4* sRec = reinterpret_cast<const int4*>(smem + L::oRec);
  double* sRed = smem + L::oRed;
  unsigned char* sNat2Jf = reinterpret_cast<unsigned char*>(smem + L::nDoubles);
  unsigned char* sJf2Nat = sNat2Jf + NF * NL;
  unsigned char* sSeq = sJf2Nat + NF * NL;
  unsigned char* sInvSeq = sSeq + 4 * NL;
  unsigned char* sFaceBase = sInvSeq + 4 * NL;
  short* sEF = reinterpret_cast<short*>(sFaceBase + NF * NL);

  const int tid = threadIdx.x;
  const int chunk = A.chunkList ? A.chunkList[blockIdx.x] : blockIdx.x;
  const int e0 = chunk * K;
  const int ne = min(K, A.nOwned - e0);
  const int el = tid / NL, t = tid - el * NL;
  const int i = t / N, j = t - i * N;
  const bool active = el < ne;
  const Phys<PH> ph(A.phys);
  const TensorDev& T = *A.tab;
  const int f0 = A.chunkFaceOff[chunk], nfc = A.chunkFaceOff[chunk + 1] - f0;

  const bool needLast = A.mode == 0 && A.aLast != 0.0;
  if (tid == 0) { mbarInit(&mbar, 1); mbarInit(&mbarLast, 1); }
  __syncthreads();
  if (tid == 0) {
    if (needLast) {   // U_last is consumed at the very end: its TMA copy overlaps the whole stage
      const unsigned bytes = (unsigned)(ne * NV * NN * sizeof(double));
      mbarExpectTx(&mbarLast, bytes);
      bulkLoad(sL, A.Ulast + (size_t)e0 * NV * NN, bytes, &mbarLast);
    }
    unsigned total = (unsigned)(nfc * sizeof(int4));
    if constexpr (AFFINE) total += (unsigned)(ne * L::REC * sizeof(double)) + (unsigned)(nfc * kCF * sizeof(double));
    mbarExpectTx(&mbar, total);
    bulkLoad(smem + L::oRec, A.faceRec + f0, (unsigned)(nfc * sizeof(int4)), &mbar);
    if constexpr (AFFINE) {
      bulkLoad(sGeoE, A.geoE + (size_t)e0 * L::REC, (unsigned)(ne * L::REC * sizeof(double)), &mbar);
      bulkLoad(sCf, A.cfGeo + (size_t)f0 * kCF, (unsigned)(nfc * kCF * sizeof(double)), &mbar);
    }
  }

  // ---- own line -> registers (16-byte loads); the copy into the padded state tile follows the table set-up so that the
  //      two groups of global loads overlap ------------------------------------------------------------------------------------
  double u[NV][N];
  {
    const double* gU = A.Uin + ((size_t)(e0 + (active ? el : 0)) * NV) * NN + t * N;
#pragma unroll
    for (int v = 0; v < NV; v++) {
#pragma unroll
      for (int k = 0; k < N; k += 2) {
        const double2 x = ldg2(gU + v * NN + k);
        u[v][k] = x.x; u[v][k + 1] = x.y;
      }
    }
  }
  // tables
  for (int q = tid; q < NN; q += THREADS) { const double w = T.wq[q]; sWq[q] = w; sInvWq[q] = 1.0 / w; }
  for (int q = tid; q < NL; q += THREADS) sWf[q] = T.wf[q];
  for (int x = tid; x < NF * NL; x += THREADS) {
    const int f = x / NL, jf = x - f * NL, base = T.faceBase[x], dn = faceDirOf<3>(f);
    const int nat = dn == 2 ? base / N : dn == 0 ? base : (base / (N * N)) * N + base % N;
    sNat2Jf[f * NL + nat] = (unsigned char)jf;
    sJf2Nat[x] = (unsigned char)nat;
    sFaceBase[x] = (unsigned char)base;
  }
  for (int x = tid; x < 4 * NL; x += THREADS) { const int r = x / NL, jl = x - r * NL, jr = T.seq[x]; sSeq[x] = (unsigned char)jr; sInvSeq[r * NL + jr] = (unsigned char)jl; }

  // pressure and 1/rho of the own nodes; zeta-face traces straight from registers
  double ir[N], pr[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    double cons[NV], comp[D + 3];
#pragma unroll
    for (int v = 0; v < NV; v++) cons[v] = u[v][k];
    ir[k] = compFromCons<D>(ph, cons, comp);
    pr[k] = comp[D + 2];
  }
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int f = hexFaceOfAxis(2, s);
#pragma unroll
    for (int v = 0; v < NV; v++) {
      double x = 0.0;
#pragma unroll
      for (int k = 0; k < N; k++) x += A.lend[s * N + k] * u[v][k];
      sT[((el * NF + f) * NV + v) * NL + t] = x;
    }
  }
  mbarWait(&mbar, 0);
  // element-face -> chunk-face entry: bits 0-7 entry, bit 8 = this element is the RIGHT parent, bit 9 = this element handles the face
  for (int x = tid; x < nfc; x += THREADS) {
    const int4 rec = sRec[x];
    const int locL = rec.x - e0, locR = rec.y - e0;
    const bool inL = locL >= 0 && locL < ne, inR = rec.y >= 0 && locR >= 0 && locR < ne;
    if (inL) sEF[locL * NF + (rec.w & 15)] = (short)(x | 0x200);
    if (inR) sEF[locR * NF + ((rec.w >> 4) & 15)] = (short)(x | 0x100 | (inL ? 0 : 0x200));
  }
  if constexpr (AFFINE) { if (tid < ne) sInvDet[tid] = 1.0 / sGeoE[tid * L::REC + D * D]; }

  // ---- xi / eta face traces, one variable at a time through the exchange tile (both sides of an axis share the loads).  The
  //      tile of an element is written and read by the N^2 lines of that element only (half a warp): __syncwarp suffices. ------
  {
    double* sXe = sU + el * NNP;
#pragma unroll
    for (int v0 = 0; v0 < NV; v0 += 2) {   // two variables per round: twice the loads in flight, half the warp barriers
      if (v0 > 0) __syncwarp();
#pragma unroll
      for (int vv = v0; vv < NV && vv < v0 + 2; vv++)
#pragma unroll
        for (int k = 0; k < N; k += 2) *reinterpret_cast<double2*>(sXe + (vv - v0) * K * NNP + i * PQ + j * N + k) = make_double2(u[vv][k], u[vv][k + 1]);
      __syncwarp();
#pragma unroll
      for (int vv = v0; vv < NV && vv < v0 + 2; vv++) {
        const double* sX = sXe + (vv - v0) * K * NNP;
#pragma unroll
        for (int d = 0; d < 2; d++) {
          const int base = d == 0 ? t : i * PQ + j;          // nat = t: d = 0 -> (j',k') = t; d = 1 -> (i',k') = (i, j)
          const int stride = d == 0 ? PQ : N;
          double xm = 0.0, xp = 0.0;
#pragma unroll
          for (int a = 0; a < N; a++) {
            const double x = sX[base + a * stride];
            xm += A.lend[a] * x; xp += A.lend[N + a] * x;
          }
          sT[((el * NF + hexFaceOfAxis(d, 0)) * NV + vv) * NL + t] = xm;
          sT[((el * NF + hexFaceOfAxis(d, 1)) * NV + vv) * NL + t] = xp;
        }
      }
    }
  }
  __syncthreads();

  // ---- R2: one face point of every face of the own element; the flux replaces the traces in place -----------------------------
#pragma unroll 1
  for (int f = 0; f < NF; f++) {
    const int ef = active ? sEF[el * NF + f] : 0;
    if (!(ef & 0x200)) continue;
    const int entry = ef & 0xff;
    const bool amRight = (ef & 0x100) != 0;
    const int4 rec = sRec[entry];
    const int eL = rec.x, eR = rec.y, faceId = rec.z;
    const int lfL = rec.w & 15, lfR = (rec.w >> 4) & 15, rot = (rec.w >> 8) & 15, bc = (rec.w >> 12) & 15;
    const int jfMine = sNat2Jf[f * NL + t];
    const int jL = amRight ? sInvSeq[rot * NL + jfMine] : jfMine;
    const int jR = amRight ? jfMine : sSeq[rot * NL + jfMine];
    double n[D], jw;
    if constexpr (AFFINE) {
      const double* g = sCf + entry * kCF;
#pragma unroll
      for (int d = 0; d < D; d++) n[d] = g[d];
      jw = g[D] * sWf[jL];
    } else {
      const double* g = A.geoF + (size_t)faceId * (D + 1) * NL + jL;
#pragma unroll
      for (int d = 0; d < D; d++) n[d] = __ldg(g + d * NL);
      jw = __ldg(g + D * NL);
    }
    double* mine = sT + ((el * NF + f) * NV) * NL + t;
    double* other = nullptr;
    double cm[NV], co[NV], Fn[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) cm[v] = mine[v * NL];
    if (!amRight && eR < 0) {
      // boundary face: normal flux of the BC-constructed state, no Riemann solve (SpatialDiscrete.cpp:797-803)
      double compL[D + 3], compR[D + 3], b[D + 3];
      compFromCons<D>(ph, cm, compL);
      const double* dm = A.dummy + (size_t)(faceId - A.nInt) * (D + 3) * NL + jL;
#pragma unroll
      for (int k = 0; k < D + 3; k++) compR[k] = dm[k * NL];
      bcBoundaryVariable<D>(ph, bc, n, compL, compR, b);
      convNormalFlux<D>(ph, n, b, Fn);
    } else {
      const int eo = amRight ? eL : eR, lfo = amRight ? lfL : lfR, jo = amRight ? jL : jR;
      const int loco = eo - e0;
      if (!amRight && loco >= 0 && loco < ne) {
        other = sT + ((loco * NF + lfo) * NV) * NL + sJf2Nat[lfo * NL + jo];
#pragma unroll
        for (int v = 0; v < NV; v++) co[v] = other[v * NL];
      } else {
        const int dn = faceDirOf<3>(lfo), side = faceSideOf<3>(lfo);
        lineTraceGlobal<N, NV, NN>(A.Uin + (size_t)eo * NV * NN, sFaceBase[lfo * NL + jo], strideOf<N, 3>(dn), A.lend + side * N, co);
      }
      double consL[NV], consR[NV], compL[D + 3], compR[D + 3];
#pragma unroll
      for (int v = 0; v < NV; v++) { consL[v] = amRight ? co[v] : cm[v]; consR[v] = amRight ? cm[v] : co[v]; }
      const double irL = compFromCons<D>(ph, consL, compL), irR = compFromCons<D>(ph, consR, compR);
      convFlux<D>(ph, n, consL, compL, irL, consR, compR, irR, Fn);
    }
    const double sg = amRight ? -jw : jw;
#pragma unroll
    for (int v = 0; v < NV; v++) mine[v * NL] = Fn[v] * sg;
    if (other != nullptr) {
#pragma unroll
      for (int v = 0; v < NV; v++) other[v * NL] = -Fn[v] * jw;
    }
  }
  __syncthreads();

  // ---- R1 + R3: volume term and lifting, axis by axis --------------------------------------------------------------------------
  double R[NV][N];
#pragma unroll
  for (int v = 0; v < NV; v++)
#pragma unroll
    for (int k = 0; k < N; k++) R[v][k] = 0.0;
  // contravariant flux of reference direction dd at the own nodes
  auto fluxDir = [&](int dd, double (&F)[NV][N]) {
#pragma unroll
    for (int k = 0; k < N; k++) {
      double m[D];
      if constexpr (AFFINE) {
        const double w = sWq[t * N + k];
#pragma unroll
        for (int c = 0; c < D; c++) m[c] = sGeoE[el * L::REC + dd * D + c] * w;
      } else {
#pragma unroll
        for (int c = 0; c < D; c++) m[c] = active ? __ldg(A.geoE + ((size_t)(e0 + el) * (D * D) + dd * D + c) * NN + t * N + k) : 0.0;
      }
      double um = 0.0;
#pragma unroll
      for (int c = 0; c < D; c++) um += u[1 + c][k] * m[c];
      um *= ir[k];
      F[0][k] = u[0][k] * um;
#pragma unroll
      for (int c = 0; c < D; c++) F[1 + c][k] = u[1 + c][k] * um + pr[k] * m[c];
      F[D + 1][k] = ph.comp() ? (u[D + 1][k] + pr[k]) * um : u[D + 1][k] * um;
    }
  };
  {  // zeta: everything in registers
    double F[NV][N];
    fluxDir(2, F);
#pragma unroll
    for (int v = 0; v < NV; v++) {
      const double fm = sT[((el * NF + hexFaceOfAxis(2, 0)) * NV + v) * NL + t], fp = sT[((el * NF + hexFaceOfAxis(2, 1)) * NV + v) * NL + t];
#pragma unroll
      for (int k = 0; k < N; k++) {
        double r = -(A.lend[k] * fm + A.lend[N + k] * fp);
#pragma unroll
        for (int a = 0; a < N; a++) r += A.dm[a * N + k] * F[v][a];
        R[v][k] += r;
      }
    }
  }
#pragma unroll
  for (int d = 0; d < 2; d++) {   // xi (d = 0), eta (d = 1): the flux of one variable at a time through the exchange tile
    const int id = d == 0 ? i : j;                      // own index along the axis
    double dmi[N];
#pragma unroll
    for (int a = 0; a < N; a++) dmi[a] = A.dm[a * N + id];
    const double lm = A.lend[id], lp = A.lend[N + id];
    // metric row d of (J^T)^-1 detJ w at the own nodes: affine = (element constant) x (node weight); curved = read per use (L1 / L2)
    double gd[D], wk[N], um[N];
    if constexpr (AFFINE) {
#pragma unroll
      for (int c = 0; c < D; c++) gd[c] = sGeoE[el * L::REC + d * D + c];
#pragma unroll
      for (int k = 0; k < N; k++) wk[k] = sWq[t * N + k];
    }
    auto metric = [&](int c, int k) -> double {
      if constexpr (AFFINE) return gd[c] * wk[k];
      else return active ? __ldg(A.geoE + ((size_t)(e0 + el) * (D * D) + d * D + c) * NN + t * N + k) : 0.0;
    };
#pragma unroll
    for (int k = 0; k < N; k++) {
      double x = 0.0;
#pragma unroll
      for (int c = 0; c < D; c++) x += u[1 + c][k] * metric(c, k);
      um[k] = x * ir[k];
    }
    double* sXe = sU + el * NNP;
    const double* sFl = sXe + (d == 0 ? j * N : i * PQ);   // line start: (0, j, :) or (i, 0, :)
    const int stride = d == 0 ? PQ : N;
    const int natRow = d == 0 ? j * N : i * N;            // face point of the own nodes: (j, k) or (i, k)
    const double* fm = sT + ((el * NF + hexFaceOfAxis(d, 0)) * NV) * NL + natRow;
    const double* fp = sT + ((el * NF + hexFaceOfAxis(d, 1)) * NV) * NL + natRow;
#pragma unroll
    for (int v0 = 0; v0 < NV; v0 += 2) {   // two variables per round through the two exchange tiles
      __syncwarp();                                     // the previous tiles have been consumed by every line of the element
#pragma unroll
      for (int v = v0; v < NV && v < v0 + 2; v++) {
        double F[N];
#pragma unroll
        for (int k = 0; k < N; k++) {
          if (v == 0) F[k] = u[0][k] * um[k];
          else if (v <= D) F[k] = u[v][k] * um[k] + pr[k] * metric(v - 1, k);
          else F[k] = ph.comp() ? (u[D + 1][k] + pr[k]) * um[k] : u[D + 1][k] * um[k];
        }
#pragma unroll
        for (int k = 0; k < N; k += 2) *reinterpret_cast<double2*>(sXe + (v - v0) * K * NNP + i * PQ + j * N + k) = make_double2(F[k], F[k + 1]);
      }
      __syncwarp();
#pragma unroll
      for (int v = v0; v < NV && v < v0 + 2; v++) {
        const double* sFv = sFl + (v - v0) * K * NNP;
#pragma unroll
        for (int k = 0; k < N; k += 2) {
          const double2 a0 = *reinterpret_cast<const double2*>(fm + v * NL + k), a1 = *reinterpret_cast<const double2*>(fp + v * NL + k);
          double r0 = -(lm * a0.x + lp * a1.x), r1 = -(lm * a0.y + lp * a1.y);
#pragma unroll
          for (int a = 0; a < N; a++) {
            const double2 x = *reinterpret_cast<const double2*>(sFv + a * stride + k);
            r0 += dmi[a] * x.x; r1 += dmi[a] * x.y;
          }
          R[v][k] += r0; R[v][k + 1] += r1;
        }
      }
    }
  }

  // ---- R4: mass inverse, RK update (16-byte stores) -------------------------------------------------------------------------------
  double ijw[N];
#pragma unroll
  for (int k = 0; k < N; k++) {
    if constexpr (AFFINE) ijw[k] = sInvDet[el] * sInvWq[t * N + k];
    else ijw[k] = active ? __ldg(A.invjw + (size_t)(e0 + el) * NN + t * N + k) : 1.0;
  }
  if (A.phys.source == kBoussinesq) {  // SpatialDiscrete.cpp:254-262 + :1016-1032 (source·detJ w, times Φ)
#pragma unroll
    for (int k = 0; k < N; k++) {
      double cons[NV], comp[D + 3];
#pragma unroll
      for (int v = 0; v < NV; v++) cons[v] = u[v][k];
      compFromCons<D>(ph, cons, comp);
      R[D][k] += boussinesqSource<D>(ph, comp) / ijw[k];
    }
  }
  if (needLast) mbarWait(&mbarLast, 0);
  if (active) {
    const size_t g = ((size_t)(e0 + el) * NV) * NN + t * N;
    const double* sLe = sL + (el * NV) * NN + t * N;
#pragma unroll
    for (int v = 0; v < NV; v++) {
#pragma unroll
      for (int k = 0; k < N; k += 2) {
        double2 o;
        if (A.mode == 0) {
          o.x = A.aCur * u[v][k] + A.bdt * (R[v][k] * ijw[k]);
          o.y = A.aCur * u[v][k + 1] + A.bdt * (R[v][k + 1] * ijw[k + 1]);
          if (A.aLast != 0.0) { const double2 l = *reinterpret_cast<const double2*>(sLe + v * NN + k); o.x += A.aLast * l.x; o.y += A.aLast * l.y; }
        } else if (A.mode == 1) { o.x = R[v][k] * ijw[k]; o.y = R[v][k + 1] * ijw[k + 1]; }
        else { o.x = R[v][k]; o.y = R[v][k + 1]; }
        *reinterpret_cast<double2*>(A.Uout + g + (size_t)v * NN + k) = o;
      }
    }
  }

  // ---- K: relative error = mean_q |(K1⊗K1⊗K1) R|, summed over the chunk's elements (TimeIntegration.cpp:279-324) ------------------
  if (A.normPartial != nullptr) {
    double acc[NV];
    {  // zeta in registers
      double S[NV][N];
#pragma unroll
      for (int v = 0; v < NV; v++)
#pragma unroll
        for (int k = 0; k < N; k++) { double s = 0.0;
#pragma unroll
          for (int a = 0; a < N; a++) s += A.k1[a * N + k] * R[v][a];
          S[v][k] = s; }
#pragma unroll
      for (int v = 0; v < NV; v++)
#pragma unroll
        for (int k = 0; k < N; k++) R[v][k] = S[v][k];
    }
#pragma unroll
    for (int d = 0; d < 2; d++) {
      const int id = d == 0 ? i : j;
      double* sXe = sU + el * NNP;
      const double* sFl = sXe + (d == 0 ? j * N : i * PQ);
      const int stride = d == 0 ? PQ : N;
#pragma unroll
      for (int v = 0; v < NV; v++) {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < N; k += 2) *reinterpret_cast<double2*>(sXe + i * PQ + j * N + k) = make_double2(R[v][k], R[v][k + 1]);
        __syncwarp();
#pragma unroll
        for (int k = 0; k < N; k++) {
          double x = 0.0;
#pragma unroll
          for (int a = 0; a < N; a++) x += A.k1[a * N + id] * sFl[a * stride + k];
          R[v][k] = x;
        }
      }
    }
#pragma unroll
    for (int v = 0; v < NV; v++) { double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; k++) s += fabs(R[v][k]);
      acc[v] = active ? s : 0.0; }
    // deterministic block reduction: warp shuffle, then one thread sums the warp partials in order
#pragma unroll
    for (int v = 0; v < NV; v++) {
      double s = acc[v];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
      if ((tid & 31) == 0) sRed[(tid >> 5) * NV + v] = s;
    }
    __syncthreads();
    if (tid < NV) {
      double s = 0.0;
      for (int w = 0; w < THREADS / 32; w++) s += sRed[w * NV + tid];
      A.normPartial[(size_t)chunk * NV + tid] = s / NN;
    }
  }
}

}  // namespace sdg
