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
constexpr int kMaxN = 4;

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

struct StageArgs {
  const double* Uin;      // [nTotal][NV][NN] state read by this stage (owned + ghost elements)
  const double* Ulast;    // state at the beginning of the step (read when aLast != 0)
  double* Uout;
  const double* Gvol;     // NS residual pass: volume gradient [nTotal][NV*D][NN]
  double* Gout;           // NS gradient pass output
  const double* geoE;     // affine: [n][D*D+1]; curved: [n][D*D][NN]
  const double* invjw;    // curved: [n][NN]
  const double* geoF;     // affine: [nf][D+1]; curved: [nf][D+1][NQF]
  const int4* faceRec;    // per chunk, the faces touching it
  const int* chunkFaceOff;
  const int* chunkList;   // chunk ids handled by this launch
  const double* dummy;    // boundary_dummy_variable_ (computational): [nBnd][D+3][NQF]
  const TensorDev* tab;
  double* normPartial;    // [nChunks][NV] (last stage) or nullptr
  double aLast, aCur, bdt;
  int nOwned, nInt;
  int mode;               // 0 RK update, 1 write dU/dt, 2 write R (nodal, un-inverted)
  PhysParams phys;
};

template <int N, int D> struct Pow { static constexpr int v = N * Pow<N, D - 1>::v; };
template <int N> struct Pow<N, 0> { static constexpr int v = 1; };

template <int D, int N, int K>
struct Layout {
  static constexpr int NV = D + 2, NN = Pow<N, D>::v, NQF = NN / N, NF = 2 * D, NAQ = NF * NQF;
  static constexpr int oU = 0;                              // [K][NV][NN]
  static constexpr int oF = oU + K * NV * NN;               // [K][D][NV][NN]
  static constexpr int oFlux = oF + K * D * NV * NN;        // [K][NV][NAQ]
  static constexpr int oTab = oFlux + K * NV * NAQ;         // Dm, Lend, K1, wq, wf
  static constexpr int nTabD = 2 * N * N + 2 * N + NN + NQF;
  static constexpr int nDoubles = oTab + nTabD;
  static constexpr int nInts = NF * NQF + 4 * NQF + NF * NN;
  static constexpr size_t bytes = sizeof(double) * nDoubles + sizeof(int) * nInts + 64;
};

template <int N, int D>
__device__ __forceinline__ int strideOf(int d) { int s = 1; for (int k = D - 1; k > d; k--) s *= N; return s; }
// Normal axis / side of the gmsh local faces (getElementPerAdjacencyNodeIndex, SimulationControl.cpp:177-216, with the gmsh
// reference corners): quadrangle faces (eta-,xi+,eta+,xi-), hexahedron faces (zeta-,eta-,xi-,xi+,eta+,zeta+).
// sdg_finalize checks these against the numerically derived tables of host_tables.hpp.
template <int D>
__device__ __forceinline__ int faceDirOf(int f) {
  return D == 2 ? ((0x1 | 0x0 << 2 | 0x1 << 4 | 0x0 << 6) >> (2 * f)) & 3 : ((0x2 | 0x1 << 2 | 0x0 << 4 | 0x0 << 6 | 0x1 << 8 | 0x2 << 10) >> (2 * f)) & 3;
}
template <int D>
__device__ __forceinline__ int faceSideOf(int f) { return D == 2 ? (0x6 >> f) & 1 : (f >= 3 ? 1 : 0); }

// Trace of the conserved variables of one element at one face point: U·Φ_f[face]ᵀ (AdjacencyElementVariable::get,
// VariableConvertor.cpp:432-485) = end-point interpolation along the face-normal line.
template <int N, int NV, int NN>
__device__ __forceinline__ void lineTrace(const double* __restrict__ src, int base, int stride, const double* __restrict__ lend, double* out) {
#pragma unroll
  for (int v = 0; v < NV; v++) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < N; a++) s += lend[a] * src[v * NN + base + a * stride];
    out[v] = s;
  }
}

template <int D, int N, int K, bool AFFINE, int PH>
__global__ void __launch_bounds__(kThreads) eulerStageKernel(const __grid_constant__ StageArgs A) {
  using L = Layout<D, N, K>;
  constexpr int NV = L::NV, NN = L::NN, NQF = L::NQF, NF = L::NF, NAQ = L::NAQ;
  extern __shared__ __align__(16) double smem[];
  double* sU = smem + L::oU;
  double* sF = smem + L::oF;
  double* sFlux = smem + L::oFlux;
  double* sDm = smem + L::oTab;
  double* sLend = sDm + N * N;
  double* sK1 = sLend + 2 * N;
  double* sWq = sK1 + N * N;
  double* sWf = sWq + NN;
  int* sFaceBase = reinterpret_cast<int*>(smem + L::nDoubles);
  int* sSeq = sFaceBase + NF * NQF;
  int* sNodePt = sSeq + 4 * NQF;

  const int tid = threadIdx.x;
  const int chunk = A.chunkList ? A.chunkList[blockIdx.x] : blockIdx.x;
  const int e0 = chunk * K;
  const int ne = min(K, A.nOwned - e0);
  const Phys<PH> ph(A.phys);
  const TensorDev& T = *A.tab;

  for (int i = tid; i < N * N; i += kThreads) { sDm[i] = T.Dm[i]; sK1[i] = T.K1[i]; }
  for (int i = tid; i < 2 * N; i += kThreads) sLend[i] = T.Lend[i];
  for (int i = tid; i < NN; i += kThreads) sWq[i] = T.wq[i];
  for (int i = tid; i < NQF; i += kThreads) sWf[i] = T.wf[i];
  for (int i = tid; i < NF * NQF; i += kThreads) sFaceBase[i] = T.faceBase[i];
  for (int i = tid; i < 4 * NQF; i += kThreads) sSeq[i] = T.seq[i];
  for (int i = tid; i < NF * NN; i += kThreads) sNodePt[i] = T.nodeFacePt[i];
  {  // stage the chunk's states (contiguous in global memory)
    const double* src = A.Uin + (size_t)e0 * NV * NN;
    for (int i = tid; i < ne * NV * NN; i += kThreads) sU[i] = src[i];
  }
  __syncthreads();

  // ---- R1: volume flux at the nodes, contracted with (J^T)^-1 detJ w ---------------------------------------------------
  for (int nd = tid; nd < ne * NN; nd += kThreads) {
    const int el = nd / NN, q = nd - el * NN;
    double cons[NV], comp[D + 3], F[NV * D];
#pragma unroll
    for (int v = 0; v < NV; v++) cons[v] = sU[(el * NV + v) * NN + q];
    compFromCons<D>(ph, cons, comp);
    convRawFlux<D>(ph, comp, F);
    double mt[D * D];
    if constexpr (AFFINE) {
      const double* g = A.geoE + (size_t)(e0 + el) * (D * D + 1);
      const double w = sWq[q];
#pragma unroll
      for (int k = 0; k < D * D; k++) mt[k] = g[k] * w;
    } else {
      const double* g = A.geoE + (size_t)(e0 + el) * (D * D) * NN + q;
#pragma unroll
      for (int k = 0; k < D * D; k++) mt[k] = g[k * NN];
    }
#pragma unroll
    for (int dd = 0; dd < D; dd++)
#pragma unroll
      for (int v = 0; v < NV; v++) {
        double t = 0.0;
#pragma unroll
        for (int c = 0; c < D; c++) t += F[v * D + c] * mt[dd * D + c];
        sF[((el * D + dd) * NV + v) * NN + q] = t;
      }
  }

  // ---- R2: face fluxes, once per face of the chunk ------------------------------------------------------------------------
  {
    const int f0 = A.chunkFaceOff[chunk], nfc = A.chunkFaceOff[chunk + 1] - f0;
    for (int fp = tid; fp < nfc * NQF; fp += kThreads) {
      const int fi = fp / NQF, j = fp - fi * NQF;
      const int4 rec = A.faceRec[f0 + fi];
      const int eL = rec.x, eR = rec.y, faceId = rec.z;
      const int lfL = rec.w & 15, lfR = (rec.w >> 4) & 15, rot = (rec.w >> 8) & 15, bc = (rec.w >> 12) & 15;
      double n[D], jw;
      if constexpr (AFFINE) {
        const double* g = A.geoF + (size_t)faceId * (D + 1);
#pragma unroll
        for (int d = 0; d < D; d++) n[d] = g[d];
        jw = g[D] * sWf[j];
      } else {
        const double* g = A.geoF + (size_t)faceId * (D + 1) * NQF + j;
#pragma unroll
        for (int d = 0; d < D; d++) n[d] = g[d * NQF];
        jw = g[D * NQF];
      }
      double consL[NV], compL[D + 3], Fn[NV];
      {
        const int dn = faceDirOf<D>(lfL), side = faceSideOf<D>(lfL);
        const int base = sFaceBase[lfL * NQF + j], stride = strideOf<N, D>(dn);
        const int loc = eL - e0;
        if (loc >= 0 && loc < ne) lineTrace<N, NV, NN>(sU + loc * NV * NN, base, stride, sLend + side * N, consL);
        else lineTrace<N, NV, NN>(A.Uin + (size_t)eL * NV * NN, base, stride, sLend + side * N, consL);
      }
      compFromCons<D>(ph, consL, compL);
      int jr = j;
      if (eR >= 0) {
        double consR[NV], compR[D + 3];
        jr = sSeq[rot * NQF + j];
        const int dn = faceDirOf<D>(lfR), side = faceSideOf<D>(lfR);
        const int base = sFaceBase[lfR * NQF + jr], stride = strideOf<N, D>(dn);
        const int loc = eR - e0;
        if (loc >= 0 && loc < ne) lineTrace<N, NV, NN>(sU + loc * NV * NN, base, stride, sLend + side * N, consR);
        else lineTrace<N, NV, NN>(A.Uin + (size_t)eR * NV * NN, base, stride, sLend + side * N, consR);
        compFromCons<D>(ph, consR, compR);
        convFlux<D>(ph, n, consL, compL, consR, compR, Fn);
      } else {
        // boundary face: normal flux of the BC-constructed state, no Riemann solve (SpatialDiscrete.cpp:797-803)
        double compR[D + 3], b[D + 3];
        const double* dm = A.dummy + (size_t)(faceId - A.nInt) * (D + 3) * NQF + j;
#pragma unroll
        for (int k = 0; k < D + 3; k++) compR[k] = dm[k * NQF];
        bcBoundaryVariable<D>(ph, bc, n, compL, compR, b);
        convNormalFlux<D>(ph, n, b, Fn);
      }
      {
        const int loc = eL - e0;
        if (loc >= 0 && loc < ne) {
#pragma unroll
          for (int v = 0; v < NV; v++) sFlux[(loc * NV + v) * NAQ + lfL * NQF + j] = Fn[v] * jw;
        }
      }
      if (eR >= 0) {
        const int loc = eR - e0;
        if (loc >= 0 && loc < ne) {
#pragma unroll
          for (int v = 0; v < NV; v++) sFlux[(loc * NV + v) * NAQ + lfR * NQF + jr] = -Fn[v] * jw;
        }
      }
    }
  }
  __syncthreads();

  // ---- R3 + R4: residual by sum factorisation, mass inverse, RK update ------------------------------------------------------
  const bool wantNorm = A.normPartial != nullptr;
  double* sR = sFlux;  // reused for the norm: [K][NV][NN] fits in [K][NV][NAQ] when NAQ >= NN (2D: 4N >= N^2 for N<=4; 3D: 6N^2 >= N^3 for N<=6)
  constexpr int ITERS = (K * NN + kThreads - 1) / kThreads;
  double Rkeep[ITERS][NV];
#pragma unroll
  for (int it = 0; it < ITERS; it++) {
    const int nd = tid + it * kThreads;
    if (nd >= ne * NN) break;
    const int el = nd / NN, q = nd - el * NN;
    double R[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) R[v] = 0.0;
#pragma unroll
    for (int dd = 0; dd < D; dd++) {
      const int st = strideOf<N, D>(dd);
      const int id = (q / st) % N;
      const int qb = q - id * st;
      const double* f = sF + ((el * D + dd) * NV) * NN + qb;
#pragma unroll
      for (int a = 0; a < N; a++) {
        const double dm = sDm[a * N + id];
#pragma unroll
        for (int v = 0; v < NV; v++) R[v] += f[v * NN + a * st] * dm;
      }
    }
#pragma unroll
    for (int f = 0; f < NF; f++) {
      const int dn = faceDirOf<D>(f), side = faceSideOf<D>(f);
      const int st = strideOf<N, D>(dn);
      const int id = (q / st) % N;
      const double cf = sLend[side * N + id];
      const int j = sNodePt[f * NN + q];
      const double* fl = sFlux + (el * NV) * NAQ + f * NQF + j;
#pragma unroll
      for (int v = 0; v < NV; v++) R[v] -= cf * fl[v * NAQ];
    }
    double cons[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) cons[v] = sU[(el * NV + v) * NN + q];
    double ijw;
    if constexpr (AFFINE) ijw = 1.0 / (A.geoE[(size_t)(e0 + el) * (D * D + 1) + D * D] * sWq[q]);
    else ijw = A.invjw[(size_t)(e0 + el) * NN + q];
    if (A.phys.source == kBoussinesq) {  // SpatialDiscrete.cpp:254-262 + :1016-1032 (source·detJ w, times Φ)
      double comp[D + 3];
      compFromCons<D>(ph, cons, comp);
      R[D] += boussinesqSource<D>(ph, comp) / ijw;
    }
    const size_t g = ((size_t)(e0 + el) * NV) * NN + q;
    if (A.mode == 0) {
#pragma unroll
      for (int v = 0; v < NV; v++) {
        double u = A.aCur * cons[v] + A.bdt * (R[v] * ijw);
        if (A.aLast != 0.0) u += A.aLast * A.Ulast[g + (size_t)v * NN];
        A.Uout[g + (size_t)v * NN] = u;
      }
    } else {
#pragma unroll
      for (int v = 0; v < NV; v++) A.Uout[g + (size_t)v * NN] = A.mode == 1 ? R[v] * ijw : R[v];
    }
#pragma unroll
    for (int v = 0; v < NV; v++) Rkeep[it][v] = R[v];
  }

  // ---- K: relative error = mean_q |R_modal Φᵀ| = mean_q |(K1⊗…⊗K1) R_nodal|, summed over the chunk's elements --------------
  if (wantNorm) {
    double* bufA = sF;                   // [K][NV][NN]
    double* bufB = sF + K * NV * NN;     // D >= 2 so sF holds at least two such buffers
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITERS; it++) {
      const int nd = tid + it * kThreads;
      if (nd < ne * NN) {
        const int el = nd / NN, q = nd - el * NN;
#pragma unroll
        for (int v = 0; v < NV; v++) bufA[(el * NV + v) * NN + q] = Rkeep[it][v];
      }
    }
    double acc[NV];
#pragma unroll
    for (int v = 0; v < NV; v++) acc[v] = 0.0;
    for (int dd = 0; dd < D; dd++) {
      __syncthreads();
      const double* in = (dd & 1) ? bufB : bufA;
      double* out = (dd & 1) ? bufA : bufB;
      const int st = strideOf<N, D>(dd);
      for (int nd = tid; nd < ne * NN; nd += kThreads) {
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
    double* red = sFlux;
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
  (void)sR;
}

// ---- seam / utility kernels -------------------------------------------------------------------------------------------------
// out[pos(e)][v][i] = sum_b M[i*nin+b] * in[e][b][v]      (modal -> nodal, M = Phi)        dir = 0
// out[e][i][v]      = sum_q M[i*nin+q] * in[pos(e)][v][q] (nodal -> modal, M = Phi^-1;
//                                                          or R_modal = Phi^T R_nodal)     dir = 1
__global__ void seamTransformKernel(const double* __restrict__ in, double* __restrict__ out, const double* __restrict__ M, const int* __restrict__ perm,
                                    int n, int NV, int NN, int dir) {
  extern __shared__ double sbuf[];  // one element: NV*NN
  const int e = blockIdx.x;
  if (e >= n) return;
  const int pos = perm ? perm[e] : e;
  const double* src = in + (size_t)(dir == 0 ? e : pos) * NV * NN;
  for (int i = threadIdx.x; i < NV * NN; i += blockDim.x) sbuf[i] = src[i];
  __syncthreads();
  double* dst = out + (size_t)(dir == 0 ? pos : e) * NV * NN;
  for (int o = threadIdx.x; o < NV * NN; o += blockDim.x) {
    double s = 0.0;
    if (dir == 0) {
      const int v = o / NN, i = o - v * NN;
      for (int b = 0; b < NN; b++) s += M[(size_t)i * NN + b] * sbuf[b * NV + v];
    } else {
      const int i = o / NV, v = o - i * NV;
      for (int q = 0; q < NN; q++) s += M[(size_t)i * NN + q] * sbuf[v * NN + q];
    }
    dst[o] = s;
  }
}

// [n][Nq][C] (caller order) <-> internal [pos][C][Nq]; dir 0: in -> internal, 1: internal -> out
__global__ void seamTransposeKernel(const double* __restrict__ in, double* __restrict__ out, const int* __restrict__ perm, int n, int C, int NN, int dir) {
  const size_t total = (size_t)n * C * NN;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int e = (int)(i / ((size_t)C * NN));
    const int r = (int)(i - (size_t)e * C * NN);
    const int pos = perm ? perm[e] : e;
    if (dir == 0) { const int q = r / C, c = r - q * C; out[((size_t)pos * C + c) * NN + q] = in[i]; }
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

// deterministic final reduction of the per-chunk norm partials: out[v] = sum_c partial[c][v]
__global__ void normReduceKernel(const double* __restrict__ partial, int nChunks, int NV, double* __restrict__ out) {
  __shared__ double red[8][32];
  const int v = blockIdx.x;
  double s = 0.0;
  for (int c = threadIdx.x; c < nChunks; c += blockDim.x) s += partial[(size_t)c * NV + v];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) { double t = 0.0; for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[0][w]; out[v] = t; }
}

// gather of the halo send list: out[i][:] = U[elems[i]][:]
__global__ void haloPackKernel(const double* __restrict__ U, const int* __restrict__ elems, int n, int stride, double* __restrict__ out) {
  const size_t total = (size_t)n * stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / stride);
    out[i] = U[(size_t)elems[k] * stride + (i - (size_t)k * stride)];
  }
}

}  // namespace sdg
