// tensor_kernels.cuh — fused DG stage kernels for quadrangle / hexahedron blocks on sm_100a (fp64).
//
// One launch = one RK stage of Solver::stepSolver (src/Solver/TimeIntegration.cpp:326-350) for a whole element block:
//   R1 calculateElementQuadrature            (SpatialDiscrete.cpp:194-266)   volume flux at the quadrature points
//   R2 calculateInterior/BoundaryAdjacency…  (SpatialDiscrete.cpp:633-842)   Riemann / boundary flux at the face points
//   R3 calculateElementResidual              (SpatialDiscrete.cpp:1016-1032) R = Q·∇Φ − A·Φ_f
//   R4 updateElementBasisFunctionCoefficient (TimeIntegration.cpp:181-198)   U ← a_cur U + a_last U_last + b Δt R M⁻¹
//   K  calculateElementRelativeError         (TimeIntegration.cpp:279-324)   fused block reduction on the last stage
// and, for Navier–Stokes, a preceding gradient pass (G1–G4, SpatialDiscrete.cpp:294-322,844-968,1034-1068 and
// TimeIntegration.cpp:200-228) that leaves only the volume gradient G_vol in HBM; the BR1/BR2 lifting terms are
// rank-one in the collocation basis and are rebuilt from the face jumps inside the residual pass.
//
// Representation: the block's unknowns are the conserved variables AT the volume Gauss points (Nq == Nb for these
// element types, SimulationControl.cpp:268-273), i.e. the reference's modal coefficients seen through the invertible
// map U_q = U·Φᵀ.  In that basis M = diag(detJ·w), ∇Φ is the 1-D differentiation matrix applied along lines and
// Φ_f is the 1-D end-point interpolation applied along the face-normal lines — the same discrete operator as the
// reference's dense Eigen products, evaluated by sum factorisation.
//
// Work decomposition: a thread block owns a chunk of K consecutive elements (consecutive along an internal
// space-filling curve, so a chunk is a compact brick).  The chunk's states are staged once in shared memory; every face
// that touches the chunk is evaluated once per block (faces interior to the brick once in total) and its flux is
// written to per-(element, local face, point) slots in shared memory — the reference's own race-free slot rule
// (SpatialDiscrete.cpp:406-439,738-744) — so there are no atomics.  Neighbour states outside the chunk are read
// straight from global memory (L2).
#pragma once
#include <cstdint>

#include "physics.cuh"

namespace sdg {

constexpr int kThreads = 256;
constexpr int kMaxN = 6;   // PolynomialOrderEnum P1..P5 (src/Utils/Enum.cpp)
constexpr int kCF = 6;   // chunk-face geometry record (affine meshes): {n[D], |J|, 1/detJ_left, 1/detJ_right}, padded to 6 doubles

// Device image of TensorTables (host_tables.hpp); lives in global memory, staged into shared memory by every block.
struct TensorDev {
  double Dm[kMaxN * kMaxN];     // Dm[a*N+b] = l_b'(x_a)
  double Lend[2 * kMaxN];       // [side*N+a]
  double K1[kMaxN * kMaxN];
  double wq[kMaxN * kMaxN * kMaxN];
  double wf[kMaxN * kMaxN];
  int faceDir[6], faceSide[6];
  int faceBase[6 * kMaxN * kMaxN];
  int seq[4 * kMaxN * kMaxN];   // [rot*NQF+j]
  unsigned char nodeFacePt[6 * kMaxN * kMaxN * kMaxN];
};

struct LineTabDev;   // nsl_kernels.cuh

struct StageArgs {
  const double* Uin;      // [nTotal][NV][NN] state read by this stage (owned + ghost elements)
  const double* Ulast;    // state at the beginning of the step (read when aLast != 0)
  double* Uout;
  const double* Gvol;     // NS residual pass: volume gradient [nTotal][NV*D][NN]
  double* Gout;           // NS gradient pass output
  const double* geoE;     // affine: [n][REC] (REC = D*D+1 rounded up to even: {(J^T)^-1 detJ rows, detJ}); curved: [n][D*D][NN]
  const double* invjw;    // curved: [n][NN]
  const double* geoF;     // affine: [nf][D+1]; curved: [nf][D+1][NQF]
  const double* cfGeo;    // affine: per chunk-face entry {n[D], |J| scale, 1/detJ_L, 1/detJ_R} padded to kCF doubles (same order as faceRec)
  const int4* faceRec;    // per chunk, the faces touching it
  const int* chunkFaceOff;
  const int* chunkList;   // chunk ids handled by this launch
  const double* dummy;    // boundary_dummy_variable_ (computational): [nBnd][D+3][NQF]
  const TensorDev* tab;
  double* normPartial;    // [nChunks][NV] (last stage) or nullptr
  double aLast, aCur, bdt;
  double dm[kMaxN * kMaxN], lend[2 * kMaxN], k1[kMaxN * kMaxN];   // 1-D operators in the parameter (constant) bank: Dm[a*N+b], Lend[side*N+a], K1
  int nOwned, nInt;
  int mode;               // 0 RK update, 1 write dU/dt, 2 write R (nodal, un-inverted)
  int faceSel;            // gradient diagnostics: -1 = lifts of all faces (total gradient), f = volume part + the lift of local face f only (RawBinary.cpp:110-135)
  PhysParams phys;
  // trace-based line kernels (nsl_kernels.cuh): published face traces [nTotal][6][5][16], per-(element, face) link / geometry records
  const double* TUin; double* TUout;   // traces of the state read / written by this stage
  const double* TVin; double* TVout;   // this side's viscous normal flux (pass G writes, pass R reads)
  double* TUb;                         // virtual neighbour traces of the boundary faces [nBnd][5][16]
  const int4* links;                   // [nOwned][6]
  const double* lfGeo;                 // affine: [nOwned][6][4] = face normal, |J| scale
  const LineTabDev* ltab;
  double w1[kMaxN];                    // 1-D Gauss weights
  int ahead;                           // thread blocks per wave: how far ahead a block prefetches the contiguous ranges of a later block into L2
  double cLift;                        // sum_a l_a(-1)^2 / w_a: BR2 lift trace factor of a face point, without 1 / (detJ w_face)
  // artificial viscosity (SpatialDiscrete.cpp:37-192): corner values per element [n][2^D] (variable_artificial_viscosity_), order-1 nodal
  // basis at the volume nodes [NN][2^D] (nodal_value_) and at the face points [NF*NQF][2^D] (nodal_adjacency_value_)
  const double* avElem; const double* avTabQ; const double* avTabF;
};

// eps at a point = nodal basis row * corner values (SpatialDiscrete.cpp:213-214, 539-546)
template <int NB>
__device__ __forceinline__ double avAt(const double* __restrict__ row, const double* __restrict__ corner) {
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < NB; k++) s += __ldg(row + k) * __ldg(corner + k);
  return s;
}

template <int N, int D> struct Pow { static constexpr int v = N * Pow<N, D - 1>::v; };
template <int N> struct Pow<N, 0> { static constexpr int v = 1; };

#ifndef SDG_MIN_BLOCKS
#define SDG_MIN_BLOCKS 3
#endif

template <int D, int N, int K>
struct Layout {
  static constexpr int NV = D + 2, NN = Pow<N, D>::v, NQF = NN / N, NF = 2 * D, NAQ = NF * NQF;
  static constexpr int oU = 0;                              // [K][NV][NN]   staged states
  static constexpr int oF = oU + K * NV * NN;               // [K][NV][NN]   contravariant flux of ONE reference direction
  static constexpr int oFlux = oF + K * NV * NN;            // [K][NV][NAQ]  face-flux slots
  static constexpr int oTab = oFlux + K * NV * NAQ;         // Dm, Lend, K1, wq, wf
  static constexpr int nTabD = 2 * N * N + 2 * N + NN + NQF + ((NN + NQF) & 1);   // kept even: what follows is 16-byte aligned
  static constexpr int REC = (D * D + 2) & ~1;                // per-element affine metric record
  static constexpr int MAXF = K * NF;                         // faces touching a chunk (upper bound)
  static constexpr int oGeoE = oTab + nTabD;                  // [K][REC]
  static constexpr int oCf = oGeoE + K * REC;                 // [MAXF][kCF]
  static constexpr int oRec = oCf + MAXF * kCF;               // [MAXF] int4
  static constexpr int nDoubles = oRec + MAXF * 2;
  static constexpr int nBytesTab = NF * NQF + 4 * NQF + NF * NN;   // faceBase, seq, nodeFacePt as bytes
  static constexpr size_t bytes = sizeof(double) * nDoubles + ((nBytesTab + 15) / 16) * 16;
  static constexpr int ITERS = (K * NN + kThreads - 1) / kThreads;
};

// ---- mbarrier + TMA bulk copy (global -> shared::cta), sm_90+ PTX ---------------------------------------------------------
__device__ __forceinline__ unsigned smemAddr(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // make the init visible to the async proxy
}
__device__ __forceinline__ void mbarExpectTx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulkLoad(void* dstShared, const void* srcGlobal, unsigned bytes, unsigned long long* bar) {
  if (bytes == 0) return;
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstShared)),
               "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
               : "memory");
}
__device__ __forceinline__ void mbarWait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smemAddr(bar)), "r"(parity)
                 : "memory");
  }
}

template <int N, int D>
__device__ __forceinline__ constexpr int strideOf(int d) { int s = 1; for (int k = D - 1; k > d; k--) s *= N; return s; }
// Normal axis / side of the gmsh local faces (getElementPerAdjacencyNodeIndex, SimulationControl.cpp:177-216, with the gmsh
// reference corners): quadrangle faces (eta-,xi+,eta+,xi-), hexahedron faces (zeta-,eta-,xi-,xi+,eta+,zeta+).
// sdg_finalize checks these against the numerically derived tables of host_tables.hpp.
template <int D>
__device__ __forceinline__ constexpr int faceDirOf(int f) {
  return D == 1 ? 0 : D == 2 ? ((0x1 | 0x0 << 2 | 0x1 << 4 | 0x0 << 6) >> (2 * f)) & 3 : ((0x2 | 0x1 << 2 | 0x0 << 4 | 0x0 << 6 | 0x1 << 8 | 0x2 << 10) >> (2 * f)) & 3;
}
template <int D>
__device__ __forceinline__ constexpr int faceSideOf(int f) { return D == 1 ? f : D == 2 ? (0x6 >> f) & 1 : (f >= 3 ? 1 : 0); }

// Trace of the conserved variables of one element at one face point: U·Φ_f[face]ᵀ (AdjacencyElementVariable::get,
// VariableConvertor.cpp:432-485) = end-point interpolation along the face-normal line.
template <int N, int NV, int NN>
__device__ __forceinline__ void lineTrace(const double* __restrict__ src, int base, int stride, const double* __restrict__ lend, double* out) {
  double val[NV][N];
#pragma unroll
  for (int v = 0; v < NV; v++)
#pragma unroll
    for (int a = 0; a < N; a++) val[v][a] = src[v * NN + base + a * stride];
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < N; a++) s += lend[a] * val[v][a];
    out[v] = s;
  }
}

// Same trace with the N nodes of the line visited starting at `rot`: the lanes of a warp (consecutive face points, several
// faces) then hit different shared-memory banks (N = 4: the plain order is 3-4 way conflicted on the faces whose normal is not
// the slowest axis).  Summation order differs per lane; parity is tolerance based (1e-12).
template <int N, int NFLD, int NN>
__device__ __forceinline__ void lineTraceRot(const double* __restrict__ src, int base, int stride, const double* __restrict__ lend, int rot, double* out) {
#pragma unroll
  for (int v = 0; v < NFLD; v++) out[v] = 0.0;
#pragma unroll
  for (int k = 0; k < N; k++) {
    int a = k + rot;
    a = a >= N ? a - N : a;
    const double l = lend[a];
    const double* s = src + base + a * stride;
#pragma unroll
    for (int v = 0; v < NFLD; v++) out[v] += l * s[v * NN];
  }
}

// Global-memory variant for a neighbour outside the chunk: no bank concern, so the line is read in natural order, and a
// stride-1 line (face normal = fastest axis) is read with 16-byte loads when N is even (each lane then touches every 32-byte
// sector once or twice instead of N times).
template <int N, int NFLD, int NN>
__device__ __forceinline__ void lineTraceGlobal(const double* __restrict__ src, int base, int stride, const double* __restrict__ lend, double* out) {
  if ((N % 2 == 0) && stride == 1) {
#pragma unroll
    for (int v = 0; v < NFLD; v++) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < N; a += 2) {
        const double2 x = __ldg(reinterpret_cast<const double2*>(src + v * NN + base + a));
        s += lend[a] * x.x + lend[a + 1] * x.y;
      }
      out[v] = s;
    }
  } else {
#pragma unroll
    for (int v = 0; v < NFLD; v++) out[v] = 0.0;
#pragma unroll
    for (int a = 0; a < N; a++) {
      const double l = lend[a];
      const double* s = src + base + a * stride;
#pragma unroll
      for (int v = 0; v < NFLD; v++) out[v] += l * __ldg(s + v * NN);
    }
  }
}

// Contravariant convective flux of reference direction dd at one node:  F(U)·m  with m = row dd of (J^T)^-1 detJ w
// (calculateConvectiveRawFlux, ConvectiveFlux.cpp:28-57, contracted like SpatialDiscrete.cpp:229-232).
template <int D, int PH>
__device__ __forceinline__ void contravariantFlux(const Phys<PH>& ph, const double* cons, const double* comp, const double* m, double* Ft) {
  double um = 0.0;
#pragma unroll
  for (int c = 0; c < D; c++) um += comp[1 + c] * m[c];
  const double p = comp[D + 2];
  Ft[0] = cons[0] * um;
#pragma unroll
  for (int k = 0; k < D; k++) Ft[1 + k] = cons[1 + k] * um + p * m[k];
  Ft[D + 1] = ph.comp() ? (cons[D + 1] + p) * um : cons[D + 1] * um;
}

template <int D, int N, int K, bool AFFINE, int PH>
__global__ void __launch_bounds__(kThreads, SDG_MIN_BLOCKS) eulerStageKernel(const __grid_constant__ StageArgs A) {
  using L = Layout<D, N, K>;
  constexpr int NV = L::NV, NN = L::NN, NQF = L::NQF, NF = L::NF, NAQ = L::NAQ, ITERS = L::ITERS;
  extern __shared__ __align__(16) double smem[];
  double* sU = smem + L::oU;
  double* sF = smem + L::oF;
  double* sFlux = smem + L::oFlux;
  double* sDm = smem + L::oTab;
  double* sLend = sDm + N * N;
  double* sK1 = sLend + 2 * N;
  double* sWq = sK1 + N * N;
  double* sWf = sWq + NN;
  unsigned char* sFaceBase = reinterpret_cast<unsigned char*>(smem + L::nDoubles);
  unsigned char* sSeq = sFaceBase + NF * NQF;
  unsigned char* sNodePt = sSeq + 4 * NQF;

  const int tid = threadIdx.x;
  const int chunk = A.chunkList ? A.chunkList[blockIdx.x] : blockIdx.x;
  const int e0 = chunk * K;
  const int ne = min(K, A.nOwned - e0);
  const int nNodes = ne * NN;
  // when the block size is a multiple of the nodes per element every thread keeps the same node q in all its iterations
  constexpr bool FIXQ = (kThreads % NN) == 0;
  constexpr int EL_STEP = kThreads / NN;
  const int q0 = tid % NN, el0 = tid / NN;
  const Phys<PH> ph(A.phys);
  const TensorDev& T = *A.tab;

  // Stage everything the chunk owns with TMA bulk copies (cp.async.bulk -> UBLKCP): its states, its affine metric
  // records and the records + geometry of the faces touching it are each ONE contiguous range in global memory.
  // One thread arms an mbarrier with the byte count and issues the copies; the block meets them at mbarWait below.
  double* sGeoE = smem + L::oGeoE;
  double* sCf = smem + L::oCf;
  const int4* sRec = reinterpret_cast<const int4*>(smem + L::oRec);
  __shared__ __align__(8) unsigned long long mbar;
  const int f0 = A.chunkFaceOff[chunk], nfc = A.chunkFaceOff[chunk + 1] - f0;
  const unsigned bytesU = (unsigned)(ne * NV * NN * sizeof(double));
  const bool bulkU = (bytesU & 15u) == 0;   // 16-byte granularity of the bulk copy engine
  if (tid == 0) { mbarInit(&mbar, 1); }
  __syncthreads();
  if (tid == 0) {
    unsigned total = (unsigned)(nfc * sizeof(int4)) + (bulkU ? bytesU : 0u);
    if constexpr (AFFINE) total += (unsigned)(ne * L::REC * sizeof(double)) + (unsigned)(nfc * kCF * sizeof(double));
    mbarExpectTx(&mbar, total);
    if (bulkU) bulkLoad(sU, A.Uin + (size_t)e0 * NV * NN, bytesU, &mbar);
    bulkLoad(smem + L::oRec, A.faceRec + f0, (unsigned)(nfc * sizeof(int4)), &mbar);
    if constexpr (AFFINE) {
      bulkLoad(sGeoE, A.geoE + (size_t)e0 * L::REC, (unsigned)(ne * L::REC * sizeof(double)), &mbar);
      bulkLoad(sCf, A.cfGeo + (size_t)f0 * kCF, (unsigned)(nfc * kCF * sizeof(double)), &mbar);
    }
  }
  if (!bulkU) {
    const double* src = A.Uin + (size_t)e0 * NV * NN;
    for (int i = tid; i < ne * NV * NN; i += kThreads) sU[i] = src[i];
  }
  for (int i = tid; i < N * N; i += kThreads) { sDm[i] = T.Dm[i]; sK1[i] = T.K1[i]; }
  for (int i = tid; i < 2 * N; i += kThreads) sLend[i] = T.Lend[i];
  for (int i = tid; i < NN; i += kThreads) sWq[i] = T.wq[i];
  for (int i = tid; i < NQF; i += kThreads) sWf[i] = T.wf[i];
  for (int i = tid; i < NF * NQF; i += kThreads) sFaceBase[i] = (unsigned char)T.faceBase[i];
  for (int i = tid; i < 4 * NQF; i += kThreads) sSeq[i] = (unsigned char)T.seq[i];
  for (int i = tid; i < NF * NN; i += kThreads) sNodePt[i] = T.nodeFacePt[i];

  // U_last is consumed at the very end (R4): pull its lines into L2 now
  if (A.mode == 0 && A.aLast != 0.0) {
    const char* p = reinterpret_cast<const char*>(A.Ulast + (size_t)e0 * NV * NN);
    const int bytes = ne * NV * NN * (int)sizeof(double);
    for (int o = tid * 128; o < bytes; o += kThreads * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + o));
  }
  __syncthreads();
  mbarWait(&mbar, 0);

  // ---- R2: face fluxes, once per face of the chunk ------------------------------------------------------------------------
  {
    for (int fp = tid; fp < nfc * NQF; fp += kThreads) {
      const int fi = fp / NQF, j = fp - fi * NQF;
      const int4 rec = sRec[fi];
      const int eL = rec.x, eR = rec.y, faceId = rec.z;
      const int lfL = rec.w & 15, lfR = (rec.w >> 4) & 15, rot = (rec.w >> 8) & 15, bc = (rec.w >> 12) & 15;
      double n[D], jw;
      if constexpr (AFFINE) {
        const double* g = sCf + fi * kCF;
#pragma unroll
        for (int d = 0; d < D; d++) n[d] = g[d];
        jw = g[D] * sWf[j];
      } else {
        const double* g = A.geoF + (size_t)faceId * (D + 1) * NQF + j;
#pragma unroll
        for (int d = 0; d < D; d++) n[d] = __ldg(g + d * NQF);
        jw = __ldg(g + D * NQF);
      }
      double consL[NV], compL[D + 3], consR[NV], compR[D + 3], Fn[NV];
      const int rl = (j / N + fi) % N;
      const int locL = eL - e0, locR = eR - e0;
      const bool inL = locL >= 0 && locL < ne, inR = eR >= 0 && locR >= 0 && locR < ne;
      {
        const int dn = faceDirOf<D>(lfL), side = faceSideOf<D>(lfL);
        const int base = sFaceBase[lfL * NQF + j], stride = strideOf<N, D>(dn);
        if (inL) lineTraceRot<N, NV, NN>(sU + locL * NV * NN, base, stride, sLend + side * N, rl, consL);
        else lineTraceGlobal<N, NV, NN>(A.Uin + (size_t)eL * NV * NN, base, stride, sLend + side * N, consL);
      }
      const double irL = compFromCons<D>(ph, consL, compL);
      int jr = j;
      if (eR >= 0) {
        jr = sSeq[rot * NQF + j];
        const int dn = faceDirOf<D>(lfR), side = faceSideOf<D>(lfR);
        const int base = sFaceBase[lfR * NQF + jr], stride = strideOf<N, D>(dn);
        if (inR) lineTraceRot<N, NV, NN>(sU + locR * NV * NN, base, stride, sLend + side * N, rl, consR);
        else lineTraceGlobal<N, NV, NN>(A.Uin + (size_t)eR * NV * NN, base, stride, sLend + side * N, consR);
        const double irR = compFromCons<D>(ph, consR, compR);
        convFlux<D>(ph, n, consL, compL, irL, consR, compR, irR, Fn);
      } else {
        // boundary face: normal flux of the BC-constructed state, no Riemann solve (SpatialDiscrete.cpp:797-803)
        double b[D + 3];
        const double* dm = A.dummy + (size_t)(faceId - A.nInt) * (D + 3) * NQF + j;
#pragma unroll
        for (int k = 0; k < D + 3; k++) compR[k] = dm[k * NQF];
        bcBoundaryVariable<D>(ph, bc, n, compL, compR, b);
        convNormalFlux<D>(ph, n, b, Fn);
      }
      if (inL) {
#pragma unroll
        for (int v = 0; v < NV; v++) sFlux[(locL * NV + v) * NAQ + lfL * NQF + j] = Fn[v] * jw;
      }
      if (inR) {
#pragma unroll
        for (int v = 0; v < NV; v++) sFlux[(locR * NV + v) * NAQ + lfR * NQF + jr] = -Fn[v] * jw;
      }
    }
  }

  // ---- R1 + R3: volume term, one reference direction at a time (sum factorisation) ------------------------------------------
  double R[ITERS][NV];
#pragma unroll
  for (int it = 0; it < ITERS; it++)
#pragma unroll
    for (int v = 0; v < NV; v++) R[it][v] = 0.0;
  double velp[ITERS][D + 1];  // velocity and pressure of the thread's nodes, reused by the D direction passes
#pragma unroll
  for (int it = 0; it < ITERS; it++) {
    const int nd = tid + it * kThreads;
    if (nd < nNodes) {
      const int el = FIXQ ? el0 + it * EL_STEP : nd / NN, q = FIXQ ? q0 : nd - el * NN;
      double cons[NV], comp[D + 3];
#pragma unroll
      for (int v = 0; v < NV; v++) cons[v] = sU[(el * NV + v) * NN + q];
      compFromCons<D>(ph, cons, comp);
#pragma unroll
      for (int c = 0; c < D; c++) velp[it][c] = comp[1 + c];
      velp[it][D] = comp[D + 2];
    }
  }
#pragma unroll
  for (int dd = 0; dd < D; dd++) {
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
      const int nd = tid + it * kThreads;
      if (nd < nNodes) {
        const int el = FIXQ ? el0 + it * EL_STEP : nd / NN, q = FIXQ ? q0 : nd - el * NN;
        double cons[NV], comp[D + 3], m[D], Ft[NV];
#pragma unroll
        for (int v = 0; v < NV; v++) cons[v] = sU[(el * NV + v) * NN + q];
#pragma unroll
        for (int c = 0; c < D; c++) comp[1 + c] = velp[it][c];
        comp[D + 2] = velp[it][D];
        if constexpr (AFFINE) {
          const double* g = sGeoE + el * L::REC + dd * D;
          const double w = sWq[q];
#pragma unroll
          for (int c = 0; c < D; c++) m[c] = g[c] * w;
        } else {
          const double* g = A.geoE + ((size_t)(e0 + el) * (D * D) + dd * D) * NN + q;
#pragma unroll
          for (int c = 0; c < D; c++) m[c] = __ldg(g + c * NN);
        }
        contravariantFlux<D>(ph, cons, comp, m, Ft);
#pragma unroll
        for (int v = 0; v < NV; v++) sF[(el * NV + v) * NN + q] = Ft[v];
      }
    }
    __syncthreads();   // (first pass: also orders the face-flux slot writes before their use below)
    const int st = strideOf<N, D>(dd);
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
      const int nd = tid + it * kThreads;
      if (nd < nNodes) {
        const int el = FIXQ ? el0 + it * EL_STEP : nd / NN, q = FIXQ ? q0 : nd - el * NN;
        const int id = (q / st) % N;
        const double* f = sF + (el * NV) * NN + (q - id * st);
#pragma unroll
        for (int a = 0; a < N; a++) {
          const double dm = sDm[a * N + id];
#pragma unroll
          for (int v = 0; v < NV; v++) R[it][v] += f[v * NN + a * st] * dm;
        }
      }
    }
    if (dd + 1 < D) __syncthreads();
  }

  // ---- R3 (face part) + R4: lifting of the face fluxes, mass inverse, RK update ---------------------------------------------
#pragma unroll
  for (int it = 0; it < ITERS; it++) {
    const int nd = tid + it * kThreads;
    if (nd < nNodes) {
      const int el = FIXQ ? el0 + it * EL_STEP : nd / NN, q = FIXQ ? q0 : nd - el * NN;
#pragma unroll
      for (int f = 0; f < NF; f++) {
        const int dn = faceDirOf<D>(f), side = faceSideOf<D>(f);
        const int st = strideOf<N, D>(dn);
        const int id = (q / st) % N;
        const double cf = sLend[side * N + id];
        const int j = sNodePt[f * NN + q];
        const double* fl = sFlux + (el * NV) * NAQ + f * NQF + j;
#pragma unroll
        for (int v = 0; v < NV; v++) R[it][v] -= cf * fl[v * NAQ];
      }
      double cons[NV];
#pragma unroll
      for (int v = 0; v < NV; v++) cons[v] = sU[(el * NV + v) * NN + q];
      double ijw;
      if constexpr (AFFINE) ijw = 1.0 / (sGeoE[el * L::REC + D * D] * sWq[q]);
      else ijw = __ldg(A.invjw + (size_t)(e0 + el) * NN + q);
      if (A.phys.source == kBoussinesq) {  // SpatialDiscrete.cpp:254-262 + :1016-1032 (source·detJ w, times Φ)
        double comp[D + 3];
        compFromCons<D>(ph, cons, comp);
        R[it][D] += boussinesqSource<D>(ph, comp) / ijw;
      }
      const size_t g = ((size_t)(e0 + el) * NV) * NN + q;
      if (A.mode == 0) {
#pragma unroll
        for (int v = 0; v < NV; v++) {
          double u = A.aCur * cons[v] + A.bdt * (R[it][v] * ijw);
          if (A.aLast != 0.0) u += A.aLast * __ldg(A.Ulast + g + (size_t)v * NN);
          A.Uout[g + (size_t)v * NN] = u;
        }
      } else {
#pragma unroll
        for (int v = 0; v < NV; v++) A.Uout[g + (size_t)v * NN] = A.mode == 1 ? R[it][v] * ijw : R[it][v];
      }
    }
  }

  // ---- K: relative error = mean_q |R_modal Φᵀ| = mean_q |(K1⊗…⊗K1) R_nodal|, summed over the chunk's elements --------------
  if (A.normPartial != nullptr) {
    double* bufA = sF;      // [K][NV][NN]
    double* bufB = (L::NAQ >= L::NN) ? sFlux : sU;   // [K][NV][NAQ] >= [K][NV][NN] in 3-D (6N^2 >= N^3) and for N <= 4 in 2-D; otherwise the (dead) state tile
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
      const int nd = tid + it * kThreads;
      if (nd < nNodes) {
        const int el = FIXQ ? el0 + it * EL_STEP : nd / NN, q = FIXQ ? q0 : nd - el * NN;
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
          for (int a = 0; a < N; a++) s += sK1[a * N + id] * in[(el * NV + v) * NN + qb + a * st];
          if (dd == D - 1) acc[v] += fabs(s); else out[(el * NV + v) * NN + q] = s;
        }
      }
    }
    __syncthreads();
    // deterministic block reduction: warp shuffle, then one thread sums the warp partials in order
    double* red = sU;
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

// ---- seam / utility kernels -------------------------------------------------------------------------------------------------
// Sum-factorised modal <-> nodal transform of the state setters / getters (sdg_set_state, sdg_get_state, sdg_residual).  The modal basis of
// a tensor element is Phi[q][b] = prod_d Phi1[i_d(q)][k_d(b)] (host_tables.hpp, modalIdx), so the dense NN x NN product per field
// factors into D passes with the N x N matrix M1 (row = output index, column = input index) — 3*N instead of N^3 multiply-adds per
// value, and the element's data goes through shared memory once.  lexOf[b] = sum_d k_d(b) N^(D-1-d).
//   dir 0: in [e][b][NV] (caller order, modal) -> out [perm[e]][NV][NN] (internal, nodal), M1 = Phi1
//   dir 1: in [perm[e]][NV][NN] -> out [e][b][NV], M1 = Phi1^-1 (coefficients) or Phi1^T (projection of a nodal residual)
// EPB elements per block iteration; dynamic shared memory 2 * EPB * NV * (NN + 1) doubles (rows padded by one: the modal side is
// accessed with the variable index fastest, which would otherwise hit one bank NV times).  <DT, NT> = compile-time D, N (0: run time).
template <int DT, int NT>
static __global__ void __launch_bounds__(256) tensorTransformKernel(const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ M1,
                                                                     const int* __restrict__ lexOf, const int* __restrict__ perm, int n, int Nrt, int Drt,
                                                                     int dir, int EPB) {
  extern __shared__ double sbuf[];
  __shared__ double sM[kMaxN * kMaxN];
  __shared__ int sLex[kMaxN * kMaxN * kMaxN];
  const int N = NT ? NT : Nrt, D = DT ? DT : Drt, NV = D + 2;
  int NN = 1;
  for (int d = 0; d < D; d++) NN *= N;
  const int per = NV * NN, RS = NN + 1, perP = NV * RS;
  double* A = sbuf;
  double* B = sbuf + (size_t)EPB * perP;
  for (int i = threadIdx.x; i < N * N; i += blockDim.x) sM[i] = M1[i];
  for (int i = threadIdx.x; i < NN; i += blockDim.x) sLex[i] = lexOf[i];
  __syncthreads();
  for (int e0 = blockIdx.x * EPB; e0 < n; e0 += gridDim.x * EPB) {
    const int ne = min(EPB, n - e0), total = ne * per;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int el = i / per, r = i - el * per;
      if (dir == 0) {
        const int b = r / NV, v = r - b * NV;
        A[el * perP + v * RS + sLex[b]] = in[(size_t)(e0 + el) * per + r];
      } else {
        const int pos = perm ? perm[e0 + el] : e0 + el;
        const int v = r / NN, q = r - v * NN;
        A[el * perP + v * RS + q] = in[(size_t)pos * per + r];
      }
    }
    __syncthreads();
    int stride = NN;
    for (int d = 0; d < D; d++) {
      stride /= N;
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int row = i / NN, q = i - row * NN, id = (q / stride) % N;   // row = el * NV + v
        const double* a = A + row * RS + (q - id * stride);
        const double* m = sM + id * N;
        double acc = 0.0;
        if constexpr (NT > 0) {
#pragma unroll
          for (int k = 0; k < NT; k++) acc += m[k] * a[k * stride];
        } else {
          for (int k = 0; k < N; k++) acc += m[k] * a[k * stride];
        }
        B[row * RS + q] = acc;
      }
      __syncthreads();
      double* t = A; A = B; B = t;
    }
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
      const int el = i / per, r = i - el * per;
      if (dir == 0) {
        const int pos = perm ? perm[e0 + el] : e0 + el;
        const int v = r / NN, q = r - v * NN;
        out[(size_t)pos * per + r] = A[el * perP + v * RS + q];
      } else {
        const int b = r / NV, v = r - b * NV;
        out[(size_t)(e0 + el) * per + r] = A[el * perP + v * RS + sLex[b]];
      }
    }
    __syncthreads();
  }
}

// RawBinary payload rows: nodal fields [pos][C][NN] (zslow: internal node index (q % 4) * 16 + q / 4) -> modal coefficients
// out[row][b][C] = Σ_q M[b][q] in[pos][c][q].  list == nullptr: every element e (pos = perm[e], row = e); else list[i] = {pos, row}.
static __global__ void modalRowsKernel(const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ M, const int* __restrict__ perm,
                                       const int2* __restrict__ list, int n, int C, int NN, int zslow) {
  extern __shared__ double sbuf[];  // one element: C*NN
  const int e = blockIdx.x;
  if (e >= n) return;
  const int pos = list ? list[e].x : (perm ? perm[e] : e);
  const int row = list ? list[e].y : e;
  const double* src = in + (size_t)pos * C * NN;
  for (int i = threadIdx.x; i < C * NN; i += blockDim.x) {
    const int c = i / NN, q = i - c * NN;
    sbuf[i] = src[c * NN + (zslow ? (q & 3) * 16 + (q >> 2) : q)];
  }
  __syncthreads();
  double* dst = out + (size_t)row * C * NN;
  for (int o = threadIdx.x; o < C * NN; o += blockDim.x) {
    const int b = o / C, c = o - b * C;
    double s = 0.0;
    for (int q = 0; q < NN; q++) s += M[(size_t)b * NN + q] * sbuf[c * NN + q];
    dst[o] = s;
  }
}

// [n][Nq][C] (caller order) <-> internal [pos][C][Nq]; dir 0: in -> internal, 1: internal -> out, 3: as 1 with internal node index (q % 4) * 16 + q / 4
static __global__ void seamTransposeKernel(const double* __restrict__ in, double* __restrict__ out, const int* __restrict__ perm, int n, int C, int NN, int dir) {
  const size_t total = (size_t)n * C * NN;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(i / ((size_t)C * NN));
    const int r = (int)(i - (size_t)e * C * NN);
    const int pos = perm ? perm[e] : e;
    if (dir == 0) { const int q = r / C, c = r - q * C; out[((size_t)pos * C + c) * NN + q] = in[i]; }
    else if (dir == 3) { const int q = r / C, c = r - q * C; out[i] = in[((size_t)pos * C + c) * NN + (q & 3) * 16 + (q >> 2)]; }   // P3 hexahedra, zeta-slowest internal node order
    else { const int q = r / C, c = r - q * C; out[i] = in[((size_t)pos * C + c) * NN + q]; }
  }
}

// Solver::initializeSolver for collocation blocks (InitialCondition.cpp:85-116): primitive (rho,u,T) at the quadrature
// points -> conserved; the unweighted least-squares projection is the identity in the collocation basis.
template <int D>
__global__ void primitiveToStateKernel(const double* __restrict__ prim, double* __restrict__ U, const int* __restrict__ perm, int n, int NN, PhysParams P) {
  constexpr int NV = D + 2;
  const size_t total = (size_t)n * NN;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(i / NN), q = (int)(i - (size_t)e * NN);
    const double* p = prim + i * NV;
    const double rho = p[0];
    double v2 = 0;
    for (int d = 0; d < D; d++) v2 += p[1 + d] * p[1 + d];
    const int pos = perm ? perm[e] : e;
    double* u = U + (size_t)pos * NV * NN + q;
    u[0] = rho;
    for (int d = 0; d < D; d++) u[(1 + d) * NN] = rho * p[1 + d];
    const double e_int = P.cv * p[D + 1];
    u[(D + 1) * NN] = P.compressible ? rho * (e_int + 0.5 * v2) : rho * e_int;  // VariableConvertor.cpp:341-366
  }
}

// boundary_dummy_variable_ (InitialCondition.cpp:118-149): primitive [nBnd][NQF][NV] -> computational [nBnd][D+3][NQF]
template <int D>
__global__ void boundaryPrimitiveKernel(const double* __restrict__ prim, double* __restrict__ dummy, int nBnd, int NQF, PhysParams P) {
  constexpr int NV = D + 2;
  const Phys<0> ph(P);
  const size_t total = (size_t)nBnd * NQF;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int f = (int)(i / NQF), j = (int)(i - (size_t)f * NQF);
    double pr[NV], comp[D + 3];
    for (int k = 0; k < NV; k++) pr[k] = prim[i * NV + k];
    compFromPrim<D>(ph, pr, comp);
    for (int k = 0; k < D + 3; k++) dummy[((size_t)f * (D + 3) + k) * NQF + j] = comp[k];
  }
}

// calculateElementDeltaTime, TimeIntegration.cpp:104-131: per-block minimum of CFL·minEdge / ((|u|+c)(p+1)^2)
template <int D>
__global__ void deltaTimeKernel(const double* __restrict__ U, const double* __restrict__ minEdge, int n, int NN, int p, double cfl, PhysParams P, double* __restrict__ partial) {
  constexpr int NV = D + 2;
  const Phys<0> ph(P);
  double best = 1.7976931348623157e308;
  const size_t total = (size_t)n * NN;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(i / NN), q = (int)(i - (size_t)e * NN);
    double cons[NV], comp[D + 3];
    for (int v = 0; v < NV; v++) cons[v] = U[((size_t)e * NV + v) * NN + q];
    compFromCons<D>(ph, cons, comp);
    const double sr = sqrt(vsq<D>(comp)) + ph.sound(comp[0], comp[D + 2]);
    const double dt = cfl * minEdge[e] / (sr * (p + 1.0) * (p + 1.0));
    best = dt < best ? dt : best;   // NaN never wins, like std::min
  }
  __shared__ double red[32];
  for (int o = 16; o > 0; o >>= 1) { const double t = __shfl_down_sync(0xffffffffu, best, o); best = t < best ? t : best; }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = best;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); w++) best = red[w] < best ? red[w] : best;
    partial[blockIdx.x] = best;
  }
}

// deterministic reduction of the per-chunk norm partials: block (v, y) sums its contiguous slice of the rows,
// out[y][v] = sum_{c in slice y} partial[c][v]; launched twice (slices, then the slice sums) so that a 2M-element mesh is not
// summed by NV thread blocks
static __global__ void normReduceKernel(const double* __restrict__ partial, int nRows, int NV, double* __restrict__ out) {
  __shared__ double red[32];
  const int v = blockIdx.x;
  const int per = (nRows + gridDim.y - 1) / gridDim.y;
  const int lo = blockIdx.y * per, hi = min(nRows, lo + per);
  double s = 0.0;
  for (int c = lo + threadIdx.x; c < hi; c += blockDim.x) s += partial[(size_t)c * NV + v];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w]; out[(size_t)blockIdx.y * NV + v] = t; }
}

// ---- peer-memory halo exchange (one process per GPU, CUDA IPC over NVLink / NVSwitch) -----------------------------------------------
// A peer's arrays as seen from this rank (cudaIpcOpenMemHandle) and where this rank's send elements land in them.
struct PeerDev {
  double* dst[4];        // the peer's U[0], U[1], U[2], G allocations
  long long* flags;      // the peer's arrival flags; this rank owns entry `slot`
  long long ghostFirst;  // element index, in the peer's arrays, of the first ghost element fed by this rank
  int sendFirst, sendCount, slot;
};
// Gather the send units (elements, or (element, face) trace rows) of field `which` and store them STRAIGHT into the peers' arrays (no
// staging buffer, no send/recv pair): one warp per unit, 16-byte loads and stores, the peer looked up once per unit.  Destination of
// unit k: dstUnit[k] if given (trace rows: the peer's row index), else the peer's contiguous ghost range.  The last thread block to
// finish publishes this exchange's epoch in every peer's flag entry.
static __global__ void haloPushKernel(const double* __restrict__ src, const int* __restrict__ units, int nSend, int stride, const PeerDev* __restrict__ peers,
                                      int nPeers, int which, unsigned int* counter, long long epoch, const long long* __restrict__ dstUnit) {
  const int lane = threadIdx.x & 31, warpsPerBlock = blockDim.x >> 5;
  const int half = stride >> 1;
  for (int k = blockIdx.x * warpsPerBlock + (threadIdx.x >> 5); k < nSend; k += gridDim.x * warpsPerBlock) {
    int p = 0;
    while (p + 1 < nPeers && k >= peers[p].sendFirst + peers[p].sendCount) p++;
    const long long d = dstUnit ? dstUnit[k] : peers[p].ghostFirst + (k - peers[p].sendFirst);
    const double* s1 = src + (size_t)units[k] * stride;
    double* d1 = peers[p].dst[which] + (size_t)d * stride;
    if (stride & 1) {   // odd unit size (e.g. 27-node hexahedra x 5 variables): units are not 16-byte aligned
      for (int r = lane; r < stride; r += 32) d1[r] = s1[r];
    } else {
      const double2* s2 = reinterpret_cast<const double2*>(s1);
      double2* d2 = reinterpret_cast<double2*>(d1);
      for (int r = lane; r < half; r += 32) d2[r] = s2[r];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(counter, 1u);
    if (prev == gridDim.x - 1) {   // every block's stores are fenced: the data is visible at the peers before the flag
      *counter = 0u;
      __threadfence_system();
      for (int p = 0; p < nPeers; p++) *reinterpret_cast<volatile long long*>(peers[p].flags + peers[p].slot) = epoch;
      __threadfence_system();
    }
  }
}
// Stream-side wait for the peers' pushes of exchange `epoch` (one lane per peer).  A wait that outlasts `timeoutNs` (ranks out
// of step for good) is FATAL: the flag is raised and the kernel traps, so every later call on this context fails loudly
// instead of computing a stage on stale ghost data.
static __global__ void haloWaitKernel(const long long* __restrict__ flags, int nPeers, long long epoch, int* errFlag, unsigned long long timeoutNs) {
  const int p = threadIdx.x;
  if (p < nPeers) {
    const volatile long long* f = flags + p;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (*f < epoch) {
      __nanosleep(200);
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeoutNs) { atomicExch(errFlag, 1); __threadfence_system(); __trap(); }
    }
  }
  __threadfence_system();
}

// scatter of received units: dst[units[i]][:] = in[i][:]  (trace rows land at their (ghost element, face) rows)
static __global__ void haloUnpackKernel(const double* __restrict__ in, const int* __restrict__ units, int n, int stride, double* __restrict__ dst) {
  const size_t total = (size_t)n * stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / stride);
    dst[(size_t)units[k] * stride + (i - (size_t)k * stride)] = in[i];
  }
}

// gather of the halo send list: out[i][:] = U[elems[i]][:]
static __global__ void haloPackKernel(const double* __restrict__ U, const int* __restrict__ elems, int n, int stride, double* __restrict__ out) {
  const size_t total = (size_t)n * stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / stride);
    out[i] = U[(size_t)elems[k] * stride + (i - (size_t)k * stride)];
  }
}

}  // namespace sdg
