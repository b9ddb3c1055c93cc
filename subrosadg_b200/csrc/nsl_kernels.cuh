// nsl_kernels.cuh — trace-based, thread-per-line Navier–Stokes (BR1 / BR2) stage kernels for P3 hexahedra (sm_100a, fp64).
//
// Same two passes per RK stage as ns_kernels.cuh (pass G = G1–G4, pass R = R1–R4 + K of Solver::stepSolver,
// src/Solver/TimeIntegration.cpp:326-350), restructured after the round-1 profiles (profiles/r01_ns_s4_ncu.md: the face phase was
// 65 % of the residual pass — two 20-field line traces per face point, evaluated 5.0 times per element, neighbour lines gathered
// from L2 — at 16 warps per SM):
//
//   * every element PUBLISHES the traces of its own six faces once, like the reference's variable_adjacency_quadrature_ slots
//     (SolveControl.cpp:45-58) but per element side:  TU[e][f][v][pt]  = trace of the conserved variables (written by the residual pass
//     for the state it has just produced, read by both passes of the next stage) and  TV[e][f][v][pt] = this side's viscous normal
//     flux  F_v(U_own, trace_f(G_vol + G_f)) · n  (ViscousFlux.cpp:116-153; written by pass G, which knows the jump and therefore the
//     BR2 lift of that face, read by pass R).  A face point of the residual pass is then 20 coalesced loads, one Riemann solve and an
//     average — no line traces, no gathers of whole neighbour elements, no redundant own-side work;
//   * pass G stores the TOTAL gradient  G_vol + Σ_f G_f  at the nodes (what the volume term of pass R needs, TimeIntegration.cpp:208-223),
//     so pass R needs neither the jumps nor the lifting coefficients;
//   * a thread owns one zeta-line of one element (as eulerLineKernel): zeta operators in registers, xi / eta through a swizzled
//     shared-memory tile synchronised with __syncwarp (an element = half a warp);
//   * boundary faces look like interior faces to pass G: a small kernel writes the virtual neighbour trace  R = 2 U_vol − L  of every
//     boundary face point (BoundaryConditionImpl::calculateBoundaryGradientVariable, BoundaryCondition.cpp:287-297,426-441,...), for
//     which  ½(L+R) = U_vol  and  ½(R−L) = U_vol − L  are exactly the volume- and interface-gradient states of the reference.
//
// Point order inside a TU / TV row: the face point's position in the element's own node lattice ("natural" order: the two tangential
// lattice indices, lower axis first); LineTabDev maps a point to the partner's natural index for every (face, face, rotation, side).
#pragma once
#include "line_kernels.cuh"

namespace sdg {

// A/B switches of round 2 (64^3 NS / 96^3 Euler launch lists, profiles/r02_ns_line_ncu.md): the defaults are the measured winners
#ifndef SDG_NSLG_NBRSEL
#define SDG_NSLG_NBRSEL 1     // pass G, flux phase: partner rows from the offsets of phase one (select chain), no second table lookup
#endif
#ifndef SDG_NSLG_PRELOAD
#define SDG_NSLG_PRELOAD 2    // pass G, flux phase: partner (1) and own (2) trace values requested before the 15-field trace of G_vol: 1.445 -> 1.397 ms
#endif
#ifndef SDG_NSL_LAZYJL
#define SDG_NSL_LAZYJL 1
#endif
#ifndef SDG_NSLS_HOIST
#define SDG_NSLS_HOIST 1      // pass R, inviscid: partner rows of all six faces before the direction loop: 2.88 / 3.05 / 3.73 -> 2.83 / 2.98 / 3.58 ms at 96^3
#endif

// DIAGNOSTIC builds only (wrong numbers, tools/gpu_diag_euler.sh): what the face phase of the inviscid residual pass costs.
//   1 = the lower face of every direction is skipped (upper bound of evaluating every face once)
//   2 = the Riemann solve is replaced by the average of the two states (loads and slot traffic stay)
//   3 = the partner row is not loaded (own trace on both sides; the Riemann solve stays)
#ifndef SDG_NSL_DIAG
#define SDG_NSL_DIAG 0
#endif

#if SDG_NSLG_PRELOAD && !SDG_NSLG_NBRSEL
#error "SDG_NSLG_PRELOAD needs SDG_NSLG_NBRSEL"
#endif
constexpr int kLK = 8;                 // elements per thread block (a 2x2x2 brick of the internal Morton order)
constexpr int kRow = 5 * 16;           // doubles per (element, face) row of TU / TV
// (row * 16 + point) -> offset of the value of variable 0 inside a trace array
__device__ __forceinline__ size_t nbrOffset(int packed) { return (size_t)(packed >> 4) * kRow + (packed & 15); }
constexpr int kLG = 4;                 // affine per-(element, face) geometry record: face normal n[3] (left-outward), |J| scale

// link record of (element, local face): .x = other parent (internal position, -1 = boundary face), .y = face id,
// .z = lfo | rot << 3 | bc << 5 | amRight << 8 | handles << 9 | otherInChunk << 10, .w = unused
__device__ __forceinline__ int linkLfo(int z) { return z & 7; }
__device__ __forceinline__ int linkRot(int z) { return (z >> 3) & 3; }
__device__ __forceinline__ int linkBc(int z) { return (z >> 5) & 7; }
__device__ __forceinline__ bool linkAmRight(int z) { return (z >> 8) & 1; }
__device__ __forceinline__ bool linkHandles(int z) { return (z >> 9) & 1; }
__device__ __forceinline__ bool linkInChunk(int z) { return (z >> 10) & 1; }

struct LineTabDev {
  unsigned char partner[6 * 6 * 4 * 2 * 16];   // [(((f*6+lfo)*4+rot)*2+amRight)*16 + t]: natural point index at the other parent
  unsigned char jLeft[6 * 4 * 2 * 16];         // [((f*4+rot)*2+amRight)*16 + t]: point index (reference order) at the LEFT parent = column of geoF / dummy
};

// swizzled tile of one field of one element: node (i, j, k) at i*16 + ((j+i)&3)*4 + (k ^ 2(i&1)) — lines along xi and eta are read
// with 8-byte accesses, the own zeta line is written / read as two 16-byte pairs, all without bank conflicts and without padding
// (the pair swap by the parity of i separates the lines (0, j) and (1, j), which a quarter warp of 16-byte accesses covers together)
__device__ __forceinline__ int tIdx(int i, int j, int k) { return i * 16 + (((j + i) & 3) << 2) + (k ^ ((i & 1) << 1)); }
__device__ __forceinline__ int tPair(int i, int j, int p) { return i * 16 + (((j + i) & 3) << 2) + ((p ^ (i & 1)) << 1); }
// DRAM -> L2 ahead of use: the ranges a block owns are contiguous (bulk prefetch by one thread), the partners' rows are 128-byte lines
__device__ __forceinline__ void bulkPrefetchL2(const void* p, unsigned bytes) {
  if (bytes) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
#ifndef SDG_NSLS_PF_L1
#define SDG_NSLS_PF_L1 0      // A/B: the partners' rows of the residual pass requested into L1 instead of L2 at block start: measured 7.83 -> 7.88 ms (Euler 128^3), 10.16 -> 10.34 ms (NS 96^3): off
#endif
// The few KB a block waits for before it can do anything (link records, affine metric and face geometry) are fetched into L2 one wave of
// thread blocks ahead, by the block that currently occupies the slot: unlike the bulk data (measured: fetching THAT ahead only churns
// L2), they are small enough to stay, and the block-start wait on DRAM (6 % of the stall samples) becomes an L2 hit.
template <bool AFFINE>
__device__ __forceinline__ void prefetchBlockHeaderAhead(const StageArgs& A, int K) {
  const int b = (int)blockIdx.x + 148 * 3;
  if (b >= (int)gridDim.x) return;
  const int c2 = A.chunkList ? A.chunkList[b] : b;
  const int f0 = c2 * K, n2 = min(K, A.nOwned - f0);
  bulkPrefetchL2(A.links + (size_t)f0 * 6, (unsigned)(n2 * 6 * sizeof(int4)));
  if constexpr (AFFINE) {
    bulkPrefetchL2(A.geoE + (size_t)f0 * 10, (unsigned)(n2 * 10 * sizeof(double)));
    bulkPrefetchL2(A.lfGeo + (size_t)f0 * 6 * kLG, (unsigned)(n2 * 6 * kLG * sizeof(double)));
  }
}
__device__ __forceinline__ double2 lds2(const double* p) { return *reinterpret_cast<const double2*>(p); }
__device__ __forceinline__ void sts2(double* p, double a, double b) { *reinterpret_cast<double2*>(p) = make_double2(a, b); }

// Traces of the conserved variables of one element at its own 6 x 16 face points (AdjacencyElementVariable::get,
// VariableConvertor.cpp:432-485), from the line registers; sXe = the element's exchange tile [5][64].
__device__ __forceinline__ void lineTracesOut(const StageArgs& A, const double (&u)[5][4], double* sXe, int i, int j, int t, unsigned wm, double* gT) {
#pragma unroll
  for (int s = 0; s < 2; s++) {
    const int f = hexFaceOfAxis(2, s);
#pragma unroll
    for (int v = 0; v < 5; v++) {
      double x = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) x += A.lend[s * 4 + k] * u[v][k];
      gT[(f * 5 + v) * 16 + t] = x;
    }
  }
  const int p0 = tPair(i, j, 0), p1 = tPair(i, j, 1);
#pragma unroll
  for (int v = 0; v < 5; v++) { sts2(sXe + v * 64 + p0, u[v][0], u[v][1]); sts2(sXe + v * 64 + p1, u[v][2], u[v][3]); }
  __syncwarp(wm);
#pragma unroll
  for (int v = 0; v < 5; v++) {
    double xm = 0.0, xp = 0.0, ym = 0.0, yp = 0.0;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const double x = sXe[v * 64 + tIdx(a, i, j)];   // xi-normal faces: point (j', k') = (i, j), nodes (a, i, j)
      const double y = sXe[v * 64 + tIdx(i, a, j)];   // eta-normal faces: point (i', k') = (i, j), nodes (i, a, j)
      xm += A.lend[a] * x; xp += A.lend[4 + a] * x;
      ym += A.lend[a] * y; yp += A.lend[4 + a] * y;
    }
    gT[(2 * 5 + v) * 16 + t] = xm; gT[(3 * 5 + v) * 16 + t] = xp;
    gT[(1 * 5 + v) * 16 + t] = ym; gT[(4 * 5 + v) * 16 + t] = yp;
  }
}

// ---- U -> TU for a state that did not come out of the residual pass (initial condition, sdg_set_state) ------------------------------
static __global__ void __launch_bounds__(128) nslTraceKernel(const __grid_constant__ StageArgs A) {
  extern __shared__ __align__(16) double smem[];
  const int tid = threadIdx.x, chunk = A.chunkList ? A.chunkList[blockIdx.x] : blockIdx.x;
  const int e0 = chunk * kLK, ne = min(kLK, A.nOwned - e0);
  const int el = tid >> 4, t = tid & 15, i = t >> 2, j = t & 3;
  const bool active = el < ne;
  const unsigned wm = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const int e = e0 + el;
  double u[5][4];
#pragma unroll
  for (int v = 0; v < 5; v++)
#pragma unroll
    for (int k = 0; k < 4; k += 2) { const double2 x = ldg2(A.Uin + ((size_t)e * 5 + v) * 64 + t * 4 + k); u[v][k] = x.x; u[v][k + 1] = x.y; }
  lineTracesOut(A, u, smem + el * 5 * 64, i, j, t, wm, A.TUout + (size_t)e * 6 * kRow);
}

// ---- virtual neighbour traces of the boundary faces (see the header) ------------------------------------------------------------------
template <bool AFFINE>
__global__ void __launch_bounds__(128) nslBoundaryKernel(const __grid_constant__ StageArgs A, const int4* __restrict__ bndRec, int nBnd) {
  const int idx = blockIdx.x * 128 + threadIdx.x;
  const int t = idx & 15;
  if ((idx >> 4) >= nBnd) return;
  const int fb = A.chunkList ? A.chunkList[idx >> 4] : idx >> 4;   // chunkList: here a list of boundary faces (sdg_step_host launches them by level)
  const Phys<0> ph(A.phys);
  const int4 r = bndRec[fb];   // left parent (internal position), its local face, boundary condition, face id
  const int jL = A.ltab->jLeft[((r.y * 4 + 0) * 2 + 0) * 16 + t];
  double n[3];
  if constexpr (AFFINE) {
#pragma unroll
    for (int c = 0; c < 3; c++) n[c] = A.lfGeo[((size_t)r.x * 6 + r.y) * kLG + c];
  } else {
#pragma unroll
    for (int c = 0; c < 3; c++) n[c] = __ldg(A.geoF + ((size_t)r.w * 4 + c) * 16 + jL);
  }
  double consL[5], compL[6], compR[6], vol[5], itf[5];
#pragma unroll
  for (int v = 0; v < 5; v++) consL[v] = A.TUin[((size_t)r.x * 6 + r.y) * kRow + v * 16 + t];
  compFromCons<3>(ph, consL, compL);
#pragma unroll
  for (int k = 0; k < 6; k++) compR[k] = A.dummy[((size_t)fb * 6 + k) * 16 + jL];
  bcBoundaryGradientVariable<3>(ph, r.z, n, consL, compL, compR, vol, itf);
#pragma unroll
  for (int v = 0; v < 5; v++) A.TUb[(size_t)fb * kRow + v * 16 + t] = 2.0 * vol[v] - consL[v];
}

// ---- boundary faces, out of line: they are rare, and inlining the six boundary-condition bodies at every face of both passes pushes
//      the hot loops out of the 32 KB instruction cache (measured: stall_no_instruction 4.4 warps per issue) -------------------------------
struct Vals5 { double v[5]; };
struct Vals15 { double v[15]; };

// pass R (SpatialDiscrete.cpp:797-812): normal flux of the BC-constructed state minus the face's averaged viscous flux (published by pass G)
template <int PH, bool VISC>
__device__ __noinline__ Vals5 nslBoundaryFaceFlux(const PhysParams& P, int bc, double n0, double n1, double n2, Vals5 cm, Vals5 tm, const double* __restrict__ dummyCol) {
  const Phys<PH> ph(P);
  const double n[3] = {n0, n1, n2};
  double compL[6], compR[6], b[6], Fn[5];
  compFromCons<3>(ph, cm.v, compL);
#pragma unroll
  for (int k = 0; k < 6; k++) compR[k] = dummyCol[k * 16];
  bcBoundaryVariable<3>(ph, bc, n, compL, compR, b);
  convNormalFlux<3>(ph, n, b, Fn);
  Vals5 out;
#pragma unroll
  for (int v = 0; v < 5; v++) out.v[v] = VISC ? Fn[v] - tm.v[v] : Fn[v];
  return out;
}

// pass G (SpatialDiscrete.cpp:750-842): the complete averaged viscous flux of a boundary face from the (un-lifted) gradient trace g
static __device__ __noinline__ Vals5 nslBoundaryViscousFlux(const PhysParams& P, int bc, double n0, double n1, double n2, double jwLam, Vals5 cmv, Vals15 gv,
                                                            const double* __restrict__ dummyCol) {
  const Phys<0> ph(P);
  const double n[3] = {n0, n1, n2};
  double comp[6], compR[6], b[6], volCons[5], intCons[5], pL[15], gb[15], va[5], vb[5];
  double* g = gv.v;
  const double* cm = cmv.v;
  compFromCons<3>(ph, cm, comp);
#pragma unroll
  for (int k = 0; k < 6; k++) compR[k] = dummyCol[k * 16];
  bcBoundaryGradientVariable<3>(ph, bc, n, cm, comp, compR, volCons, intCons);
#pragma unroll
  for (int v = 0; v < 5; v++) {
    const double jl = intCons[v] * jwLam;
#pragma unroll
    for (int c = 0; c < 3; c++) g[v * 3 + c] += jl * n[c];
  }
  primGradFromConsGrad<3>(ph, cm, comp, g, pL);             // from the UNMODIFIED interior trace (:792-796)
  bcBoundaryVariable<3>(ph, bc, n, comp, compR, b);
  if (bcIsWall(bc)) {                                       // modifyBoundaryVariable, BoundaryCondition.cpp:443-452,490-501,535-546
#pragma unroll
    for (int k = 0; k < 6; k++) comp[k] = b[k];
  }
#pragma unroll
  for (int k = 0; k < 15; k++) gb[k] = pL[k];
  if (bc == kAdiabaticSlipWall || bc == kAdiabaticNonSlipWall) { gb[12] = 0.0; gb[13] = 0.0; gb[14] = 0.0; }
  viscNormalFlux<3>(ph, n, comp, pL, va);
  viscNormalFlux<3>(ph, n, b, gb, vb);
  Vals5 out;
#pragma unroll
  for (int v = 0; v < 5; v++) out.v[v] = 0.5 * (va[v] + vb[v]);
  return out;
}

// local faces of axis d: (xi-, xi+) = (2, 3), (eta-, eta+) = (1, 4), (zeta-, zeta+) = (0, 5)
__device__ __forceinline__ int hexFaceRt(int d, int side) { return side ? 3 + d : 2 - d; }

// =====================================================================================================================
// pass G: total gradient at the nodes + this side's viscous normal flux at the face points
// =====================================================================================================================
template <bool AFFINE>
struct NslGradLayout {
  static constexpr int SLC = AFFINE ? 1 : 3;                      // slot components: a (affine, the normal is a face constant) or a*n[c]
  static constexpr int oGv = 0;                                   // [K][15][64]  swizzled tiles of G_vol (BR2) / G (BR1)
  static constexpr int oSl = oGv + kLK * 15 * 64;                 // [K][4][2*SLC][16]  xi / eta face slots of the current variable
  static constexpr int oX = oSl + kLK * 4 * 2 * SLC * 16;         // curved: [K][6][64] exchange of metric x state products
  static constexpr int oGeoE = oX + (AFFINE ? 0 : kLK * 6 * 64);  // affine: [K][10]
  static constexpr int oLg = oGeoE + (AFFINE ? kLK * 10 : 0);     // affine: [K][6][kLG]
  static constexpr int oLink = oLg + (AFFINE ? kLK * 6 * kLG : 0);   // [K][6] int4
  static constexpr int nDoubles = oLink + kLK * 6 * 2;
  static constexpr size_t bytes = sizeof(double) * nDoubles;
};

// viscous normal flux of one side at one face point from the (lifted) trace of the conserved-variable gradient
template <int PH>
__device__ __forceinline__ void ownViscousNormalFlux(const Phys<PH>& ph, const double* n, const double* cons, const double* comp, const double* g, double* va) {
  double gp[15];
  primGradFromConsGrad<3>(ph, cons, comp, g, gp);
  viscNormalFlux<3>(ph, n, comp, gp, va);
}

template <bool AFFINE>
__global__ void __launch_bounds__(128, AFFINE ? 3 : 2) nslGradKernel(const __grid_constant__ StageArgs A) {
  using L = NslGradLayout<AFFINE>;
  constexpr int K = kLK, SLC = L::SLC;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) unsigned long long mbar;
  double* sGv = smem + L::oGv;
  double* sSl = smem + L::oSl;
  double* sGeoE = smem + L::oGeoE;
  double* sLg = smem + L::oLg;
  const int4* sLink = reinterpret_cast<const int4*>(smem + L::oLink);
  const int tid = threadIdx.x, chunk = A.chunkList ? A.chunkList[blockIdx.x] : blockIdx.x;
  const int e0 = chunk * K, ne = min(K, A.nOwned - e0);
  const int el = tid >> 4, t = tid & 15, i = t >> 2, j = t & 3;
  const bool active = el < ne;
  const Phys<0> ph(A.phys);
  const bool br1 = A.phys.visc == kBR1;
  if (tid == 0) mbarInit(&mbar, 1);
  __syncthreads();
  if (tid == 0) {
    unsigned total = (unsigned)(ne * 6 * sizeof(int4));
    if constexpr (AFFINE) total += (unsigned)(ne * 10 * sizeof(double)) + (unsigned)(ne * 6 * kLG * sizeof(double));
    mbarExpectTx(&mbar, total);
    bulkLoad(smem + L::oLink, A.links + (size_t)e0 * 6, (unsigned)(ne * 6 * sizeof(int4)), &mbar);
    if constexpr (AFFINE) {
      bulkLoad(sGeoE, A.geoE + (size_t)e0 * 10, (unsigned)(ne * 10 * sizeof(double)), &mbar);
      bulkLoad(sLg, A.lfGeo + (size_t)e0 * 6 * kLG, (unsigned)(ne * 6 * kLG * sizeof(double)), &mbar);
    }
  }
  if (tid == 64) prefetchBlockHeaderAhead<AFFINE>(A, K);
  if (tid == 32) {
    // DRAM -> L2 one wave of thread blocks AHEAD: the block that will run on this SM slot next finds its contiguous ranges in L2, and
    // the DRAM transfer overlaps this block's arithmetic (the first wave fetches its own)
    for (int b = (int)blockIdx.x < A.ahead ? (int)blockIdx.x : (int)blockIdx.x + A.ahead; b < (int)gridDim.x && b <= (int)blockIdx.x + A.ahead; b += A.ahead) {
      const int c2 = A.chunkList ? A.chunkList[b] : b;
      const int f0 = c2 * K, n2 = min(K, A.nOwned - f0);
      bulkPrefetchL2(A.TUin + (size_t)f0 * 6 * kRow, (unsigned)(n2 * 6 * kRow * sizeof(double)));
      bulkPrefetchL2(A.Uin + (size_t)f0 * 5 * 64, (unsigned)(n2 * 5 * 64 * sizeof(double)));
    }
  }
  const unsigned wm = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const int e = e0 + el;
  const int p0 = tPair(i, j, 0), p1 = tPair(i, j, 1);
  const double* gU = A.Uin + (size_t)e * 5 * 64 + t * 4;
  double un[4];   // the line of the NEXT variable (requested one iteration ahead)
  { const double2 a = ldg2(gU), b = ldg2(gU + 2); un[0] = a.x; un[1] = a.y; un[2] = b.x; un[3] = b.y; }
  mbarWait(&mbar, 0);
  for (int c = t; c < 30; c += 16) {   // the partners' rows of the six faces, one 128-byte line per (face, variable)
    const int f = c / 5, v = c - f * 5;
    const int4 lk = sLink[el * 6 + f];
    if (lk.x >= 0) prefetchL2(A.TUin + ((size_t)lk.x * 6 + linkLfo(lk.z)) * kRow + v * 16);
  }

  // ---- per-face set-up of this thread's face point (natural index t on each of the six faces) ---------------------------------------
  int nbr[6];          // >= 0: (partner's row in TU) * 16 + its point; < 0: -(1 + (boundary face row in TUb) * 16 + point).  Rows, not element offsets: 32 bits hold 2^27 rows = 22 M elements
  double jw[6];        // |J| w at the point
  unsigned amRightBits = 0;
  const double wij = A.w1[i] * A.w1[j];   // weight of the face point (t = two lattice indices) and of the line's nodes without w1[k]
#pragma unroll
  for (int f = 0; f < 6; f++) {
    const int4 lk = sLink[el * 6 + f];
    const int z = lk.z, lfo = linkLfo(z), rot = linkRot(z), amR = linkAmRight(z) ? 1 : 0;
    amRightBits |= (unsigned)amR << f;
    if (lk.x >= 0) nbr[f] = (lk.x * 6 + lfo) * 16 + A.ltab->partner[(((f * 6 + lfo) * 4 + rot) * 2 + amR) * 16 + t];
    else nbr[f] = -1 - ((lk.y - A.nInt) * 16 + t);
    if constexpr (AFFINE) jw[f] = sLg[(el * 6 + f) * kLG + 3] * wij;
    else jw[f] = __ldg(A.geoF + ((size_t)lk.y * 4 + 3) * 16 + A.ltab->jLeft[((f * 4 + rot) * 2 + amR) * 16 + t]);
  }
  double invDet = 0.0, ijw[4];
  if constexpr (AFFINE) {
    invDet = 1.0 / sGeoE[el * 10 + 9];
#pragma unroll
    for (int k = 0; k < 4; k++) ijw[k] = invDet / (wij * A.w1[k]);
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) ijw[k] = __ldg(A.invjw + (size_t)e * 64 + t * 4 + k);
  }
  auto normalAt = [&](int f, int jL, double* n) {   // face normal at this thread's point of face f
    if constexpr (AFFINE) {
#pragma unroll
      for (int c = 0; c < 3; c++) n[c] = sLg[(el * 6 + f) * kLG + c];
    } else {
      const int4 lk = sLink[el * 6 + f];
#pragma unroll
      for (int c = 0; c < 3; c++) n[c] = __ldg(A.geoF + ((size_t)lk.y * 4 + c) * 16 + jL);
    }
  };
  const double* gTU = A.TUin + (size_t)e * 6 * kRow + t;

  // ---- G1-G4, one conserved variable at a time (a rolled loop: the body is ~600 instructions); the line and the face values of the
  //      NEXT variable are requested before this one is worked on ---------------------------------------------------------------------
  const unsigned selBits = A.faceSel < 0 ? 63u : 1u << A.faceSel;   // diagnostics: the lift of one face only
  double fmine[6], fother[6];
  auto loadFaceValues = [&](int v) {
#pragma unroll
    for (int f = 0; f < 6; f++) {
      fmine[f] = __ldg(gTU + (f * 5 + v) * 16);
      fother[f] = nbr[f] >= 0 ? __ldg(A.TUin + nbrOffset(nbr[f]) + v * 16) : A.TUb[nbrOffset(-1 - nbr[f]) + v * 16];
    }
  };
  loadFaceValues(0);
#pragma unroll 1
  for (int v = 0; v < 5; v++) {
    double aZ[2][2][3];   // zeta faces (this thread's own line): [side][vol / tot][c] = a n[c]
    double cmine[6], cother[6], uv[4];
#pragma unroll
    for (int f = 0; f < 6; f++) { cmine[f] = fmine[f]; cother[f] = fother[f]; }
#pragma unroll
    for (int k = 0; k < 4; k++) uv[k] = un[k];
    if (v + 1 < 5) {
      loadFaceValues(v + 1);
      const double2 a = ldg2(gU + (v + 1) * 64), b2 = ldg2(gU + (v + 1) * 64 + 2);
      un[0] = a.x; un[1] = a.y; un[2] = b2.x; un[3] = b2.y;
    }
#pragma unroll
    for (int f = 0; f < 6; f++) {
      const bool amR = (amRightBits >> f) & 1;
      const double mine = cmine[f];
      const double other = cother[f];
      const double avg = 0.5 * (mine + other), jmp = 0.5 * (amR ? mine - other : other - mine);   // ViscousFlux.cpp:33-56
      const double aV = (amR ? -avg : avg) * jw[f], aT = aV + ((selBits >> f) & 1 ? jmp * jw[f] : 0.0);   // SpatialDiscrete.cpp:885-906
      constexpr int dnTab[6] = {2, 1, 0, 0, 1, 2};
      const int dn = dnTab[f], side = f >= 3 ? 1 : 0;
      if (dn == 2 || !AFFINE) {
        double n[3];
        int jL = 0;
        if constexpr (!AFFINE) { const int z = sLink[el * 6 + f].z; jL = A.ltab->jLeft[((f * 4 + linkRot(z)) * 2 + (amR ? 1 : 0)) * 16 + t]; }
        normalAt(f, jL, n);
        if (dn == 2) {
#pragma unroll
          for (int c = 0; c < 3; c++) { aZ[side][0][c] = aV * n[c]; aZ[side][1][c] = aT * n[c]; }
        } else {
          const int fs = dn == 0 ? side : 2 + side;
#pragma unroll
          for (int c = 0; c < 3; c++) { sSl[((el * 4 + fs) * 2 * SLC + c) * 16 + t] = aV * n[c]; sSl[((el * 4 + fs) * 2 * SLC + SLC + c) * 16 + t] = aT * n[c]; }
        }
      } else {
        const int fs = dn == 0 ? side : 2 + side;
        sSl[((el * 4 + fs) * 2 + 0) * 16 + t] = aV; sSl[((el * 4 + fs) * 2 + 1) * 16 + t] = aT;
      }
    }
    // volume term  − (U ⊗ (J^T)^-1 detJ w) ∇Φ  (SpatialDiscrete.cpp:294-322,1034-1068)
    double Gv[3][4], Gt[3][4];
    if constexpr (AFFINE) {
      double* sX = sGv + (el * 15 + 3 * v) * 64;   // tile of a field that has not been written yet
      double uw[4];
#pragma unroll
      for (int k = 0; k < 4; k++) uw[k] = uv[k] * (wij * A.w1[k]);
      sts2(sX + p0, uw[0], uw[1]); sts2(sX + p1, uw[2], uw[3]);
      __syncwarp(wm);
      double tx[4] = {0, 0, 0, 0}, ty[4] = {0, 0, 0, 0}, tz[4];
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const double dx = A.dm[a * 4 + i], dy = A.dm[a * 4 + j];
        const double2 x0 = lds2(sX + tPair(a, j, 0)), x1 = lds2(sX + tPair(a, j, 1));
        const double2 y0 = lds2(sX + tPair(i, a, 0)), y1 = lds2(sX + tPair(i, a, 1));
        tx[0] += dx * x0.x; tx[1] += dx * x0.y; tx[2] += dx * x1.x; tx[3] += dx * x1.y;
        ty[0] += dy * y0.x; ty[1] += dy * y0.y; ty[2] += dy * y1.x; ty[3] += dy * y1.y;
      }
#pragma unroll
      for (int k = 0; k < 4; k++) { double s = 0.0;
#pragma unroll
        for (int a = 0; a < 4; a++) s += A.dm[a * 4 + k] * uw[a];
        tz[k] = s; }
#pragma unroll
      for (int c = 0; c < 3; c++) {
        const double g0 = sGeoE[el * 10 + c], g1 = sGeoE[el * 10 + 3 + c], g2 = sGeoE[el * 10 + 6 + c];
#pragma unroll
        for (int k = 0; k < 4; k++) Gv[c][k] = -(g0 * tx[k] + g1 * ty[k] + g2 * tz[k]);
      }
    } else {
      // curved: the metric sits inside the line sums, so the products  metric[dd][c] * U  travel through the tile (3 fields per direction)
      double* sX = smem + L::oX + el * 6 * 64;
      double mz[3][4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const double* ge = A.geoE + (size_t)e * 9 * 64 + t * 4 + k;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          sX[(0 + c) * 64 + tIdx(i, j, k)] = __ldg(ge + (0 + c) * 64) * uv[k];
          sX[(3 + c) * 64 + tIdx(i, j, k)] = __ldg(ge + (3 + c) * 64) * uv[k];
          mz[c][k] = __ldg(ge + (6 + c) * 64) * uv[k];
        }
      }
      __syncwarp(wm);
#pragma unroll
      for (int c = 0; c < 3; c++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          double s = 0.0;
#pragma unroll
          for (int a = 0; a < 4; a++)
            s += A.dm[a * 4 + i] * sX[(0 + c) * 64 + tIdx(a, j, k)] + A.dm[a * 4 + j] * sX[(3 + c) * 64 + tIdx(i, a, k)] +
                 A.dm[a * 4 + k] * mz[c][a];
          Gv[c][k] = -s;
        }
      }
    }
    // lifting of the face terms  A Φ_f  (ViscousFlux.cpp:26-56): volume-gradient states into G_vol, volume + interface into the total
#pragma unroll
    for (int c = 0; c < 3; c++)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        Gt[c][k] = Gv[c][k] + A.lend[k] * aZ[0][1][c] + A.lend[4 + k] * aZ[1][1][c];
        Gv[c][k] += A.lend[k] * aZ[0][0][c] + A.lend[4 + k] * aZ[1][0][c];
      }
#pragma unroll
    for (int fs = 0; fs < 4; fs++) {   // 0, 1: xi- (face 2), xi+ (face 3): points (j, k);  2, 3: eta- (face 1), eta+ (face 4): points (i, k)
      const int side = fs & 1, f = fs == 0 ? 2 : fs == 1 ? 3 : fs == 2 ? 1 : 4;
      const double cf = A.lend[side * 4 + (fs < 2 ? i : j)];
      const int row = (fs < 2 ? j : i) * 4;
      if constexpr (AFFINE) {
        const double* s0 = sSl + ((el * 4 + fs) * 2 + 0) * 16 + row;
        const double2 v0 = lds2(s0), v1 = lds2(s0 + 2), t0 = lds2(s0 + 16), t1 = lds2(s0 + 18);
        const double av[4] = {v0.x, v0.y, v1.x, v1.y}, at[4] = {t0.x, t0.y, t1.x, t1.y};
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const double nc = sLg[(el * 6 + f) * kLG + c] * cf;
#pragma unroll
          for (int k = 0; k < 4; k++) { Gv[c][k] += nc * av[k]; Gt[c][k] += nc * at[k]; }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 3; c++) {
          const double* s0 = sSl + ((el * 4 + fs) * 6 + c) * 16 + row;
          const double2 v0 = lds2(s0), v1 = lds2(s0 + 2), t0 = lds2(s0 + 48), t1 = lds2(s0 + 50);
          Gv[c][0] += cf * v0.x; Gv[c][1] += cf * v0.y; Gv[c][2] += cf * v1.x; Gv[c][3] += cf * v1.y;
          Gt[c][0] += cf * t0.x; Gt[c][1] += cf * t0.y; Gt[c][2] += cf * t1.x; Gt[c][3] += cf * t1.y;
        }
      }
    }
    // G4: mass inverse (diagonal in the collocation basis, TimeIntegration.cpp:200-228).  The total goes to HBM with zeta as the SLOWEST
    // node index (thread t reads / writes entry k*16 + t: coalesced 8-byte accesses for a thread per line); G_vol stays in the tile.
    __syncwarp(wm);   // every lane of the element has finished reading the exchange tile and the slots of this variable
#pragma unroll
    for (int c = 0; c < 3; c++) {
      double* out = A.Gout + ((size_t)e * 15 + v * 3 + c) * 64 + t;
#pragma unroll
      for (int k = 0; k < 4; k++) out[k * 16] = Gt[c][k] * ijw[k];
      double* tile = sGv + (el * 15 + v * 3 + c) * 64;
      if (br1) { sts2(tile + p0, Gt[c][0] * ijw[0], Gt[c][1] * ijw[1]); sts2(tile + p1, Gt[c][2] * ijw[2], Gt[c][3] * ijw[3]); }
      else { sts2(tile + p0, Gv[c][0] * ijw[0], Gv[c][1] * ijw[1]); sts2(tile + p1, Gv[c][2] * ijw[2], Gv[c][3] * ijw[3]); }
    }
  }
  __syncwarp(wm);

  // ---- this side's viscous normal flux at the own face points: trace_f(G_vol) + BR2 lift of face f (VariableConvertor.cpp:674-688) ----
#pragma unroll 1
  for (int d = 0; d < 3; d++) {
    int off[4];   // tile offsets of the four nodes of the normal line through this thread's point (natural index t) of the direction's faces
#pragma unroll
    for (int a = 0; a < 4; a++) off[a] = d == 0 ? tIdx(a, i, j) : tIdx(i, a, j);   // (d == 2: the own line, read as two pairs below)
#if SDG_NSLG_NBRSEL
    // the partner rows of the direction's two faces: the offsets of phase one (a select chain instead of a second table lookup)
    const int nb0 = d == 0 ? nbr[2] : d == 1 ? nbr[1] : nbr[0], nb1 = d == 0 ? nbr[3] : d == 1 ? nbr[4] : nbr[5];
#endif
#if SDG_NSLG_PRELOAD
    double oth[2][5];   // requested before the 15-field trace below, consumed after it
#pragma unroll
    for (int v = 0; v < 5; v++) { oth[0][v] = nb0 >= 0 ? __ldg(A.TUin + nbrOffset(nb0) + v * 16) : 0.0; oth[1][v] = nb1 >= 0 ? __ldg(A.TUin + nbrOffset(nb1) + v * 16) : 0.0; }
#endif
#if SDG_NSLG_PRELOAD >= 2
    double own[2][5];
#pragma unroll
    for (int v = 0; v < 5; v++) { own[0][v] = __ldg(gTU + (hexFaceRt(d, 0) * 5 + v) * 16); own[1][v] = __ldg(gTU + (hexFaceRt(d, 1) * 5 + v) * 16); }
#endif
    double gm[15], gp[15];
#pragma unroll
    for (int fld = 0; fld < 15; fld++) {
      const double* tile = sGv + (el * 15 + fld) * 64;
      double x0, x1, x2, x3;
      if (d == 2) { const double2 a = lds2(tile + p0), b = lds2(tile + p1); x0 = a.x; x1 = a.y; x2 = b.x; x3 = b.y; }   // own line: 16-byte accesses (8-byte ones are 8-way bank conflicted)
      else { x0 = tile[off[0]]; x1 = tile[off[1]]; x2 = tile[off[2]]; x3 = tile[off[3]]; }
      gm[fld] = A.lend[0] * x0 + A.lend[1] * x1 + A.lend[2] * x2 + A.lend[3] * x3;
      gp[fld] = A.lend[4] * x0 + A.lend[5] * x1 + A.lend[6] * x2 + A.lend[7] * x3;
    }
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const int f = hexFaceRt(d, side);
      double* g = side ? gp : gm;
      const int4 lk = sLink[el * 6 + f];
      const int z = lk.z;
      const bool amR = linkAmRight(z);
      const int jL = (!SDG_NSL_LAZYJL || !AFFINE || lk.x < 0) ? A.ltab->jLeft[((f * 4 + linkRot(z)) * 2 + (amR ? 1 : 0)) * 16 + t] : 0;   // curved geometry / boundary values only
      double n[3], jwf;
      if constexpr (AFFINE) {
        const double* lg = sLg + (el * 6 + f) * kLG;
        n[0] = lg[0]; n[1] = lg[1]; n[2] = lg[2]; jwf = lg[3] * wij;
      } else {
        const double* gf = A.geoF + (size_t)lk.y * 4 * 16 + jL;
        n[0] = __ldg(gf); n[1] = __ldg(gf + 16); n[2] = __ldg(gf + 32); jwf = __ldg(gf + 48);
      }
      double lam = 0.0;   // trace at the own point of the rank-one lift of this face:  Σ_a l_a(±1)^2 / (detJ w)(node a of the normal line)
      if (!br1) {
        if constexpr (AFFINE) lam = A.cLift * invDet / wij;
        else {
#pragma unroll
          for (int a = 0; a < 4; a++) {
            const int node = d == 2 ? t * 4 + a : d == 0 ? a * 16 + i * 4 + j : i * 16 + a * 4 + j;
            lam += A.lend[side * 4 + a] * A.lend[side * 4 + a] * __ldg(A.invjw + (size_t)e * 64 + node);
          }
        }
      }
      double cm[5], va[5];
#if SDG_NSLG_PRELOAD >= 2
#pragma unroll
      for (int v = 0; v < 5; v++) cm[v] = own[side][v];
#else
#pragma unroll
      for (int v = 0; v < 5; v++) cm[v] = __ldg(gTU + (f * 5 + v) * 16);
#endif
      if (lk.x >= 0) {
#if SDG_NSLG_NBRSEL
        const size_t rowO = nbrOffset(side ? nb1 : nb0);
#else
        const size_t rowO = ((size_t)lk.x * 6 + linkLfo(z)) * kRow + A.ltab->partner[(((f * 6 + linkLfo(z)) * 4 + linkRot(z)) * 2 + (amR ? 1 : 0)) * 16 + t];
#endif
        double comp[6];
        compFromCons<3>(ph, cm, comp);
#pragma unroll
        for (int v = 0; v < 5; v++) {
#if SDG_NSLG_PRELOAD
          const double other = oth[side][v];
#else
          const double other = __ldg(A.TUin + rowO + v * 16);
#endif
          const double jl = 0.5 * (amR ? cm[v] - other : other - cm[v]) * jwf * lam;
#pragma unroll
          for (int c = 0; c < 3; c++) g[v * 3 + c] += jl * n[c];
        }
        ownViscousNormalFlux(ph, n, cm, comp, g, va);
      } else {
        Vals5 c5; Vals15 g15;
#pragma unroll
        for (int v = 0; v < 5; v++) c5.v[v] = cm[v];
#pragma unroll
        for (int k = 0; k < 15; k++) g15.v[k] = g[k];
        const Vals5 r = nslBoundaryViscousFlux(A.phys, linkBc(z), n[0], n[1], n[2], jwf * lam, c5, g15, A.dummy + (size_t)(lk.y - A.nInt) * 6 * 16 + jL);
#pragma unroll
        for (int v = 0; v < 5; v++) va[v] = r.v[v];
      }
      double* out = A.TVout + ((size_t)e * 6 + f) * kRow + t;
#pragma unroll
      for (int v = 0; v < 5; v++) out[v * 16] = va[v];
    }
  }
}

// =====================================================================================================================
// pass R: residual, RK update, traces of the new state, relative error
// =====================================================================================================================
template <bool GATHER, bool USM = false>
struct NslStageLayout {
  static constexpr int XE = GATHER ? 480 : 320;              // doubles per element of the exchange region
  static constexpr int oFl = 0;                              // [K][6][5][16] face-flux slots (natural point order)
  static constexpr int oX = oFl + kLK * 6 * 5 * 16;          // [K][320]: 10 half tiles [2 directions][5][32] of xi / eta fluxes, later [5][64] of the new state;
                                                             // GATHER: [K][6][5][16] own face traces first (dead after the face phase)
  static constexpr int oGeoE = oX + kLK * XE;                // affine: [K][10]
  static constexpr int oLg = oGeoE + kLK * 10;               // affine: [K][6][kLG]
  static constexpr int oLink = oLg + kLK * 6 * kLG;          // [K][6] int4
  static constexpr int oRed = oLink + kLK * 6 * 2;           // [4][5]
  static constexpr int oU = oRed + 4 * 5 + 4;                // USM: [K][5][64] the chunk's own states, staged by one TMA bulk copy
  static constexpr int nDoubles = oU + (USM ? kLK * 5 * 64 : 0);
  static constexpr size_t bytes = sizeof(double) * nDoubles;
};

// 1 (inviscid pass, published traces): the chunk's own states arrive in shared memory by a TMA bulk copy issued at block start, on their
// own mbarrier — the node phase and the update read them there instead of waiting for the L2 three times per line
#ifndef SDG_NSL_U_SMEM
#define SDG_NSL_U_SMEM 0
#endif
template <bool VISC, bool GATHER>
struct NslStageUsm { static constexpr bool value = SDG_NSL_U_SMEM != 0 && !VISC && !GATHER; };

#ifndef SDG_NSL_MINB
#define SDG_NSL_MINB 3
#endif
#ifndef SDG_NSL_MINB_EULER
#define SDG_NSL_MINB_EULER 3
#endif
// 1: the node phase reads the state two nodes at a time and the update reads it again (L1 / L2 hits) instead of holding all 20 values of
// the line in registers next to the 20 residual accumulators
#ifndef SDG_NSL_RELOAD_U
#define SDG_NSL_RELOAD_U 1
#endif
// 1: the xi / eta factors of the relative-error transform move the data by transposition (4 shared-memory reads per variable and direction
// instead of 16); 0: every thread keeps its zeta-line and reads the four lines it needs (round-2 first version)
// 1 (inviscid pass, published traces): a face between two elements of the same thread block in the eta / zeta direction of the 2 x 2 x 2
// brick is evaluated by ONE of its parents, which also writes the opposite value into the other parent's slot; the parents alternate so
// that every warp evaluates five of its six faces (the face between the two elements of a warp stays with both: no round would be saved).
// Measured at 128^3 on one box: 7.52-7.60 ms per stage against 7.57-7.63 without (+0.5 %): the block barrier the shared slots need costs
// what the sixth of the face evaluations saves (profiles/r02_ab_dedup.txt) — off
#ifndef SDG_NSL_DEDUP
#define SDG_NSL_DEDUP 0
#endif
// 1: L1 prefetch of the next direction's rows inside the face loop — measured 2.5 % SLOWER at 128^3 (profiles/r02_ab_pfnext.txt), off
#ifndef SDG_NSLS_PF_NEXT
#define SDG_NSLS_PF_NEXT 0
#endif
#ifndef SDG_NSL_NORM_TRANSPOSE
#define SDG_NSL_NORM_TRANSPOSE 1
#endif
// GATHER (inviscid only): no published traces — the own traces are computed into shared memory, partners inside the block are read from
// there, partners outside are interpolated from their nodal states in global memory (L2), as eulerLineKernel does; HBM traffic stays at
// the state itself (read U, read U_last, write U).
template <bool AFFINE, int PH, bool VISC, bool GATHER = false>
__global__ void __launch_bounds__(128, VISC ? SDG_NSL_MINB : SDG_NSL_MINB_EULER) nslStageKernel(const __grid_constant__ StageArgs A) {
  static_assert(!(GATHER && VISC), "the gathering variant is inviscid");
  constexpr bool USM = NslStageUsm<VISC, GATHER>::value;
  using L = NslStageLayout<GATHER, USM>;
  constexpr int K = kLK, XE = L::XE;
  extern __shared__ __align__(16) double smem[];
  __shared__ __align__(8) unsigned long long mbar;
  double* sFl = smem + L::oFl;
  double* sX = smem + L::oX;
  double* sGeoE = smem + L::oGeoE;
  double* sLg = smem + L::oLg;
  const int4* sLink = reinterpret_cast<const int4*>(smem + L::oLink);
  double* sRed = smem + L::oRed;
  const int tid = threadIdx.x, chunk = A.chunkList ? A.chunkList[blockIdx.x] : blockIdx.x;
  const int e0 = chunk * K, ne = min(K, A.nOwned - e0);
  const int el = tid >> 4, t = tid & 15, i = t >> 2, j = t & 3;
  const bool active = el < ne;
  const Phys<PH> ph(A.phys);
  const bool needLast = A.mode == 0 && A.aLast != 0.0;
  __shared__ __align__(8) unsigned long long mbarU;
  if (tid == 0) { mbarInit(&mbar, 1); if constexpr (USM) mbarInit(&mbarU, 1); }
  __syncthreads();
  if (tid == 0) {
    if constexpr (USM) {
      mbarExpectTx(&mbarU, (unsigned)(ne * 5 * 64 * sizeof(double)));
      bulkLoad(smem + L::oU, A.Uin + (size_t)e0 * 5 * 64, (unsigned)(ne * 5 * 64 * sizeof(double)), &mbarU);
    }
    unsigned total = (unsigned)(ne * 6 * sizeof(int4));
    if constexpr (AFFINE) total += (unsigned)(ne * 10 * sizeof(double)) + (unsigned)(ne * 6 * kLG * sizeof(double));
    mbarExpectTx(&mbar, total);
    bulkLoad(smem + L::oLink, A.links + (size_t)e0 * 6, (unsigned)(ne * 6 * sizeof(int4)), &mbar);
    if constexpr (AFFINE) {
      bulkLoad(sGeoE, A.geoE + (size_t)e0 * 10, (unsigned)(ne * 10 * sizeof(double)), &mbar);
      bulkLoad(sLg, A.lfGeo + (size_t)e0 * 6 * kLG, (unsigned)(ne * 6 * kLG * sizeof(double)), &mbar);
    }
  }
  if (tid == 64) prefetchBlockHeaderAhead<AFFINE>(A, K);
  if (tid == 32) {
    // DRAM -> L2 one wave of thread blocks AHEAD (see nslGradKernel): everything a block reads from its own contiguous ranges
    for (int b = (int)blockIdx.x < A.ahead ? (int)blockIdx.x : (int)blockIdx.x + A.ahead; b < (int)gridDim.x && b <= (int)blockIdx.x + A.ahead; b += A.ahead) {
      const int c2 = A.chunkList ? A.chunkList[b] : b;
      const int f0 = c2 * K, n2 = min(K, A.nOwned - f0);
      if constexpr (!GATHER) bulkPrefetchL2(A.TUin + (size_t)f0 * 6 * kRow, (unsigned)(n2 * 6 * kRow * sizeof(double)));
      if constexpr (VISC) bulkPrefetchL2(A.TVin + (size_t)f0 * 6 * kRow, (unsigned)(n2 * 6 * kRow * sizeof(double)));
      bulkPrefetchL2(A.Uin + (size_t)f0 * 5 * 64, (unsigned)(n2 * 5 * 64 * sizeof(double)));
      if constexpr (VISC) bulkPrefetchL2(A.Gvol + (size_t)f0 * 15 * 64, (unsigned)(n2 * 15 * 64 * sizeof(double)));
      if (needLast) bulkPrefetchL2(A.Ulast + (size_t)f0 * 5 * 64, (unsigned)(n2 * 5 * 64 * sizeof(double)));
    }
  }
  const unsigned wm = __ballot_sync(0xffffffffu, active);
  if (!active) return;
  const int e = e0 + el;
  const double wij = A.w1[i] * A.w1[j];
  mbarWait(&mbar, 0);
  for (int c = t; c < (GATHER ? 0 : VISC ? 60 : 30); c += 16) {   // the partners' rows: one 128-byte line per (face, field)
    const int f = c / (VISC ? 10 : 5), r = c - f * (VISC ? 10 : 5);
    const int4 lk = sLink[el * 6 + f];
    if (lk.x >= 0) {
      const size_t row = ((size_t)lk.x * 6 + linkLfo(lk.z)) * kRow;
#if SDG_NSLS_PF_L1
      prefetchL1(r < 5 ? A.TUin + row + r * 16 : A.TVin + row + (r - 5) * 16);
#else
      prefetchL2(r < 5 ? A.TUin + row + r * 16 : A.TVin + row + (r - 5) * 16);
#endif
    }
  }

  // ---- R2: Riemann flux minus the average of the two sides' viscous normal fluxes, at this thread's point of ALL six faces of its element.
  //      A face inside the block is evaluated by both of its parents (same inputs in the same left / right roles, hence bit-identical
  //      values): no slot of another element is ever written, so the phases of an element only need __syncwarp, and the loads of the two
  //      faces of a direction are requested together before either Riemann solve starts. ------------------------------------------------
  const double* gTU = A.TUin + (size_t)e * 6 * kRow + t;
  const double* gTV = A.TVin + (size_t)e * 6 * kRow + t;
  double u[5][4];
  double* sTr = sX;   // GATHER: [K][6][5][16] own face traces, in the exchange region (which the node phase reuses after a block barrier)
  if constexpr (GATHER) {
#pragma unroll
    for (int v = 0; v < 5; v++)
#pragma unroll
      for (int k = 0; k < 4; k += 2) { const double2 x = ldg2(A.Uin + ((size_t)e * 5 + v) * 64 + t * 4 + k); u[v][k] = x.x; u[v][k + 1] = x.y; }
    lineTracesOut(A, u, sFl + el * 480, i, j, t, wm, sTr + el * 480);   // tile = the (not yet used) flux slots of the element
    __syncthreads();                                                     // partners inside the block live in other warps
  }
  constexpr bool kDedup = SDG_NSL_DEDUP != 0 && !GATHER && !VISC;
#if SDG_NSLS_HOIST
  // partner rows of all six faces, (row * 16 + point) packed: the table lookups leave the direction loop (one latency instead of three).
  // Inviscid pass only: the viscous pass has no registers to spare (64 bytes of spills, 2.4 % slower when measured).
  constexpr bool kHoist = !GATHER && !VISC;
  int prow[6];
  if constexpr (kHoist) {
#pragma unroll
    for (int f = 0; f < 6; f++) {
      const int4 l = sLink[el * 6 + f];
      const int z = l.z, lfo = linkLfo(z), rot = linkRot(z), amR = linkAmRight(z) ? 1 : 0;
      prow[f] = l.x >= 0 ? (l.x * 6 + lfo) * 16 + A.ltab->partner[(((f * 6 + lfo) * 4 + rot) * 2 + amR) * 16 + t] : (e * 6 + f) * 16 + t;
    }
  }
#endif
#pragma unroll 1
  for (int d = 0; d < 3; d++) {
    int4 lk[2];
    int jL[2];
    size_t rowO[2];
    double cm[2][5], tm[2][5], co[2][5], to[2][5];
#pragma unroll
    for (int side = 0; side < 2; side++) {
      const int f = hexFaceRt(d, side);
      lk[side] = sLink[el * 6 + f];
      const int z = lk[side].z, lfo = linkLfo(z), rot = linkRot(z), amR = linkAmRight(z) ? 1 : 0;
      // the left parent's point index of this face point: curved geometry and boundary values only (affine interior faces never read it)
      jL[side] = (!SDG_NSL_LAZYJL || !AFFINE || lk[side].x < 0) ? A.ltab->jLeft[((f * 4 + rot) * 2 + amR) * 16 + t] : 0;
      // boundary face: the partner loads fall back on the own row (valid memory, values unused) so that no load sits behind a branch
#if SDG_NSLS_HOIST
      if constexpr (kHoist) rowO[side] = nbrOffset(side ? (d == 0 ? prow[3] : d == 1 ? prow[4] : prow[5]) : (d == 0 ? prow[2] : d == 1 ? prow[1] : prow[0]));
      else
#endif
      rowO[side] = lk[side].x >= 0 ? ((size_t)lk[side].x * 6 + lfo) * kRow + A.ltab->partner[(((f * 6 + lfo) * 4 + rot) * 2 + amR) * 16 + t]
                                   : ((size_t)e * 6 + f) * kRow + t;
    }
    // role of this element on the face: 0 = evaluates it for itself; 1 = ... and for the other parent, which sits in the same block
    // (its slot receives the opposite value); 2 = the other parent does.  Both parents derive the same answer from their block-local indices.
    int role[2] = {0, 0};
    if constexpr (kDedup) {
#pragma unroll
      for (int side = 0; side < 2; side++) {
        const int z = lk[side].z;
        if (lk[side].x >= 0 && linkInChunk(z)) {
          const int elP = lk[side].x - e0, lo = min(el, elP), hi = max(el, elP), diff = hi - lo;
          const int handler = diff == 2 ? ((lo & 4) ? hi : lo) : diff == 4 ? ((lo & 2) ? lo : hi) : -1;
          role[side] = handler < 0 ? 0 : handler == el ? 1 : 2;
        }
      }
    }
#pragma unroll
    for (int side = (SDG_NSL_DIAG == 1 && !VISC) ? 1 : 0; side < 2; side++) {
      const int f = hexFaceRt(d, side);
      if (kDedup && role[side] == 2) continue;
      if constexpr (GATHER) {
        const int z = lk[side].z, lfo = linkLfo(z);
        const int natO = (int)(rowO[side] % kRow);   // partner's natural point index (own point for a boundary face)
#pragma unroll
        for (int v = 0; v < 5; v++) cm[side][v] = sTr[((el * 6 + f) * 5 + v) * 16 + t];
        if (lk[side].x >= 0 && linkInChunk(z)) {
#pragma unroll
          for (int v = 0; v < 5; v++) co[side][v] = sTr[(((lk[side].x - e0) * 6 + lfo) * 5 + v) * 16 + natO];
        } else if (lk[side].x >= 0) {
          // partner outside the block: end-point interpolation along its face-normal line, nodal values from global memory (L2)
          const int p = natO >> 2, q = natO & 3, dn = lfo == 0 || lfo == 5 ? 2 : lfo == 1 || lfo == 4 ? 1 : 0;
          const int base = dn == 0 ? p * 4 + q : dn == 1 ? p * 16 + q : p * 16 + q * 4, stride = dn == 0 ? 16 : dn == 1 ? 4 : 1;
          const double* src = A.Uin + (size_t)lk[side].x * 5 * 64 + base;
          const double* le = A.lend + (lfo >= 3 ? 4 : 0);
          double x[5][4];
#pragma unroll
          for (int v = 0; v < 5; v++)
#pragma unroll
            for (int a = 0; a < 4; a++) x[v][a] = __ldg(src + v * 64 + a * stride);
#pragma unroll
          for (int v = 0; v < 5; v++) co[side][v] = le[0] * x[v][0] + le[1] * x[v][1] + le[2] * x[v][2] + le[3] * x[v][3];
        } else {
#pragma unroll
          for (int v = 0; v < 5; v++) co[side][v] = cm[side][v];
        }
      } else {
#pragma unroll
        for (int v = 0; v < 5; v++) {
          cm[side][v] = __ldg(gTU + (f * 5 + v) * 16);
#if SDG_NSL_DIAG == 3
          co[side][v] = cm[side][v];
#else
          co[side][v] = __ldg(A.TUin + rowO[side] + v * 16);
#endif
        }
        if constexpr (VISC) {
#pragma unroll
          for (int v = 0; v < 5; v++) { tm[side][v] = __ldg(gTV + (f * 5 + v) * 16); to[side][v] = __ldg(A.TVin + rowO[side] + v * 16); }
        }
      }
    }
#if SDG_NSLS_PF_NEXT
    // the rows of the NEXT direction's two faces into L1 while this direction's Riemann solves run (the direction loop is rolled: the
    // loads themselves cannot be hoisted across it without 40 more live registers)
    if constexpr (kHoist) {
      if (d < 2) {
#pragma unroll
        for (int side = 0; side < 2; side++) {
          const int fn = hexFaceRt(d + 1, side);
          const size_t ro = nbrOffset(side ? (d == 0 ? prow[4] : prow[5]) : (d == 0 ? prow[1] : prow[0]));
#pragma unroll
          for (int v = 0; v < 5; v++) { prefetchL1(gTU + (fn * 5 + v) * 16); prefetchL1(A.TUin + ro + v * 16); }
        }
      }
    }
#endif
#pragma unroll
    for (int side = (SDG_NSL_DIAG == 1 && !VISC) ? 1 : 0; side < 2; side++) {
      const int f = hexFaceRt(d, side);
      const int z = lk[side].z;
      const bool amR = linkAmRight(z);
      if (kDedup && role[side] == 2) continue;
      double n[3], jw, Fn[5];
      if constexpr (AFFINE) {
        const double* g = sLg + (el * 6 + f) * kLG;
        n[0] = g[0]; n[1] = g[1]; n[2] = g[2]; jw = g[3] * wij;
      } else {
        const double* g = A.geoF + (size_t)lk[side].y * 4 * 16 + jL[side];
        n[0] = __ldg(g); n[1] = __ldg(g + 16); n[2] = __ldg(g + 32); jw = __ldg(g + 48);
      }
      double* mine = sFl + ((el * 6 + f) * 5) * 16 + t;
      if (lk[side].x < 0) {
        Vals5 c5, t5;
#pragma unroll
        for (int v = 0; v < 5; v++) { c5.v[v] = cm[side][v]; t5.v[v] = VISC ? tm[side][v] : 0.0; }
        const Vals5 r = nslBoundaryFaceFlux<PH, VISC>(A.phys, linkBc(z), n[0], n[1], n[2], c5, t5, A.dummy + (size_t)(lk[side].y - A.nInt) * 6 * 16 + jL[side]);
#pragma unroll
        for (int v = 0; v < 5; v++) mine[v * 16] = r.v[v] * jw;
      } else {
        double consL[5], consR[5], compL[6], compR[6];
#pragma unroll
        for (int v = 0; v < 5; v++) { consL[v] = amR ? co[side][v] : cm[side][v]; consR[v] = amR ? cm[side][v] : co[side][v]; }
#if SDG_NSL_DIAG == 2
        if constexpr (!VISC) {
#pragma unroll
          for (int v = 0; v < 5; v++) Fn[v] = 0.5 * (consL[v] + consR[v]) * n[v % 3];
        } else
#endif
        {
        const double irL = compFromCons<3>(ph, consL, compL), irR = compFromCons<3>(ph, consR, compR);
        convFlux<3>(ph, n, consL, compL, irL, consR, compR, irR, Fn);
        }
        if constexpr (VISC) {
#pragma unroll
          for (int v = 0; v < 5; v++) Fn[v] -= 0.5 * (tm[side][v] + to[side][v]);   // calculateViscousFlux, ViscousFlux.cpp:139-153
        }
        const double sg = amR ? -jw : jw;   // left parent +, right parent - (SpatialDiscrete.cpp:738-744)
#pragma unroll
        for (int v = 0; v < 5; v++) mine[v * 16] = Fn[v] * sg;
        if constexpr (kDedup) {
          if (role[side] == 1) {   // the other parent's slot: its local face, its point, the opposite role
            double* theirs = sFl + (((lk[side].x - e0) * 6 + linkLfo(z)) * 5) * 16 + (int)(rowO[side] % kRow);
#pragma unroll
            for (int v = 0; v < 5; v++) theirs[v * 16] = -(Fn[v] * sg);
          }
        }
      }
    }
  }
#if SDG_NSL_DIAG == 1
  if constexpr (!VISC) {
    for (int f = 0; f < 3; f++)
      for (int v = 0; v < 5; v++) sFl[((el * 6 + f) * 5 + v) * 16 + t] = 0.0;
  }
#endif
  if constexpr (GATHER || kDedup) __syncthreads();   // GATHER: every warp has finished reading the trace tiles, the exchange region is free;
                                                    // kDedup: slots of this element may have been written by another warp
  else __syncwarp(wm);

  // ---- R1 + R3 volume part: fluxes at the own nodes, two nodes per round; zeta contraction in registers, xi / eta through half tiles ----
  constexpr bool RELOAD = SDG_NSL_RELOAD_U != 0 || USM;
  const double* gUl = A.Uin + (size_t)e * 5 * 64 + t * 4;
  const double* sUl = smem + L::oU + el * 5 * 64 + t * 4;   // USM only
  if constexpr (USM) mbarWait(&mbarU, 0);
  if constexpr (!GATHER && !RELOAD) {
#pragma unroll
    for (int v = 0; v < 5; v++)
#pragma unroll
      for (int k = 0; k < 4; k += 2) { const double2 x = ldg2(gUl + v * 64 + k); u[v][k] = x.x; u[v][k + 1] = x.y; }
  }
  double R[5][4];
#pragma unroll
  for (int v = 0; v < 5; v++)
#pragma unroll
    for (int k = 0; k < 4; k++) R[v][k] = 0.0;
  double* sXe = sX + el * XE;
  const int sw2 = i * 8 + (((j + i) & 3) << 1);
  double ge[9];   // affine: (J^T)^-1 detJ rows of the element, once (36 shared-memory reads per thread otherwise)
  if constexpr (AFFINE) {
#pragma unroll
    for (int q = 0; q < 9; q++) ge[q] = sGeoE[el * 10 + q];
  }
#pragma unroll
  for (int r = 0; r < 2; r++) {
    double Fx[5][2], Fy[5][2];
    double ur[5][2];
    if constexpr (RELOAD) {
#pragma unroll
      for (int v = 0; v < 5; v++) { const double2 x = USM ? lds2(sUl + v * 64 + 2 * r) : ldg2(gUl + v * 64 + 2 * r); ur[v][0] = x.x; ur[v][1] = x.y; }
    }
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
      const int k = 2 * r + kk;
      double cons[5], comp[6], Fv[15];
#pragma unroll
      for (int v = 0; v < 5; v++) cons[v] = RELOAD ? ur[v][kk] : u[v][k];
      compFromCons<3>(ph, cons, comp);
      if constexpr (VISC) {
        double g[15], gp[15];
        const double* gG = A.Gvol + (size_t)e * 15 * 64 + k * 16 + t;
#pragma unroll
        for (int fld = 0; fld < 15; fld++) g[fld] = __ldg(gG + fld * 64);
        primGradFromConsGrad<3>(ph, cons, comp, g, gp);
        viscRawFlux<3>(ph, comp, gp, Fv);
      }
#pragma unroll
      for (int dd = 0; dd < 3; dd++) {
        double m[3], Ft[5];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          if constexpr (AFFINE) m[c] = ge[dd * 3 + c] * (wij * A.w1[k]);
          else m[c] = __ldg(A.geoE + ((size_t)e * 9 + dd * 3 + c) * 64 + t * 4 + k);
        }
        contravariantFlux<3>(ph, cons, comp, m, Ft);
        if constexpr (VISC) {
#pragma unroll
          for (int v = 0; v < 5; v++) Ft[v] -= Fv[v * 3] * m[0] + Fv[v * 3 + 1] * m[1] + Fv[v * 3 + 2] * m[2];   // SpatialDiscrete.cpp:216-232
        }
        if (dd == 2) {
#pragma unroll
          for (int v = 0; v < 5; v++)
#pragma unroll
            for (int k2 = 0; k2 < 4; k2++) R[v][k2] += A.dm[k * 4 + k2] * Ft[v];
        } else {
#pragma unroll
          for (int v = 0; v < 5; v++) { if (dd == 0) Fx[v][kk] = Ft[v]; else Fy[v][kk] = Ft[v]; }
        }
      }
    }
    if (r > 0) __syncwarp(wm);   // the previous round's half tiles have been consumed by every line of the element
#pragma unroll
    for (int v = 0; v < 5; v++) { sts2(sXe + v * 32 + sw2, Fx[v][0], Fx[v][1]); sts2(sXe + (5 + v) * 32 + sw2, Fy[v][0], Fy[v][1]); }
    __syncwarp(wm);
#pragma unroll
    for (int v = 0; v < 5; v++) {
      double r0 = 0.0, r1 = 0.0;
#pragma unroll
      for (int a = 0; a < 4; a++) {
        const double2 x = lds2(sXe + v * 32 + a * 8 + (((j + a) & 3) << 1));
        const double2 y = lds2(sXe + (5 + v) * 32 + i * 8 + (((a + i) & 3) << 1));
        r0 += A.dm[a * 4 + i] * x.x + A.dm[a * 4 + j] * y.x;
        r1 += A.dm[a * 4 + i] * x.y + A.dm[a * 4 + j] * y.y;
      }
      R[v][2 * r] += r0; R[v][2 * r + 1] += r1;
    }
  }
  // ---- R3 face part: lifting of the face fluxes along the own line (zeta) and from the rows (j, :) / (i, :) of the xi / eta faces ----------
#pragma unroll
  for (int v = 0; v < 5; v++) {
    const double fm = sFl[((el * 6 + 0) * 5 + v) * 16 + t], fp = sFl[((el * 6 + 5) * 5 + v) * 16 + t];
    const double* x2 = sFl + ((el * 6 + 2) * 5 + v) * 16 + j * 4;
    const double* x3 = sFl + ((el * 6 + 3) * 5 + v) * 16 + j * 4;
    const double* y1 = sFl + ((el * 6 + 1) * 5 + v) * 16 + i * 4;
    const double* y4 = sFl + ((el * 6 + 4) * 5 + v) * 16 + i * 4;
    const double lxm = A.lend[i], lxp = A.lend[4 + i], lym = A.lend[j], lyp = A.lend[4 + j];
#pragma unroll
    for (int k = 0; k < 4; k += 2) {
      const double2 a2 = lds2(x2 + k), a3 = lds2(x3 + k), b1 = lds2(y1 + k), b4 = lds2(y4 + k);
      R[v][k] -= A.lend[k] * fm + A.lend[4 + k] * fp + lxm * a2.x + lxp * a3.x + lym * b1.x + lyp * b4.x;
      R[v][k + 1] -= A.lend[k + 1] * fm + A.lend[4 + k + 1] * fp + lxm * a2.y + lxp * a3.y + lym * b1.y + lyp * b4.y;
    }
  }

  // ---- R4: mass inverse, RK update -------------------------------------------------------------------------------------------------------
  double ijw[4];
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if constexpr (AFFINE) ijw[k] = 1.0 / (sGeoE[el * 10 + 9] * wij * A.w1[k]);
    else ijw[k] = __ldg(A.invjw + (size_t)e * 64 + t * 4 + k);
  }
  if (A.phys.source == kBoussinesq) {  // SpatialDiscrete.cpp:254-262 + :1016-1032 (source·detJ w, times Φ)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double cons[5], comp[6];
#pragma unroll
      for (int v = 0; v < 5; v++) cons[v] = USM ? sUl[v * 64 + k] : RELOAD ? __ldg(gUl + v * 64 + k) : u[v][k];
      compFromCons<3>(ph, cons, comp);
      R[3][k] += boussinesqSource<3>(ph, comp) / ijw[k];
    }
  }
  {
    const size_t g = ((size_t)e * 5) * 64 + t * 4;
    if (A.mode == 0) {
#pragma unroll
      for (int v = 0; v < 5; v++) {
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
          double uk = u[v][k], uk1 = u[v][k + 1];
          if constexpr (RELOAD) { const double2 x = USM ? lds2(sUl + v * 64 + k) : ldg2(gUl + v * 64 + k); uk = x.x; uk1 = x.y; }
          double ox = A.aCur * uk + A.bdt * (R[v][k] * ijw[k]), oy = A.aCur * uk1 + A.bdt * (R[v][k + 1] * ijw[k + 1]);
          if (needLast) { const double2 l = ldg2(A.Ulast + g + (size_t)v * 64 + k); ox += A.aLast * l.x; oy += A.aLast * l.y; }
          u[v][k] = ox; u[v][k + 1] = oy;
          *reinterpret_cast<double2*>(A.Uout + g + (size_t)v * 64 + k) = make_double2(ox, oy);
        }
      }
      // traces of the state just produced: what both passes of the next stage (and the neighbours) read
      if constexpr (!GATHER) {
        __syncwarp(wm);
        lineTracesOut(A, u, sXe, i, j, t, wm, A.TUout + (size_t)e * 6 * kRow);
      }
    } else {
#pragma unroll
      for (int v = 0; v < 5; v++)
#pragma unroll
        for (int k = 0; k < 4; k += 2) {
          const double ox = A.mode == 1 ? R[v][k] * ijw[k] : R[v][k], oy = A.mode == 1 ? R[v][k + 1] * ijw[k + 1] : R[v][k + 1];
          *reinterpret_cast<double2*>(A.Uout + g + (size_t)v * 64 + k) = make_double2(ox, oy);
        }
    }
  }

  // ---- K: relative error = mean_q |(K1⊗K1⊗K1) R|, summed over the chunk's elements (TimeIntegration.cpp:279-324) ------------------------
  if (A.normPartial != nullptr) {
    {  // zeta in registers
      double S[5][4];
#pragma unroll
      for (int v = 0; v < 5; v++)
#pragma unroll
        for (int k = 0; k < 4; k++) { double s = 0.0;
#pragma unroll
          for (int a = 0; a < 4; a++) s += A.k1[a * 4 + k] * R[v][a];
          S[v][k] = s; }
#pragma unroll
      for (int v = 0; v < 5; v++)
#pragma unroll
        for (int k = 0; k < 4; k++) R[v][k] = S[v][k];
    }
#if SDG_NSL_NORM_TRANSPOSE
    // xi, then eta, as TRANSPOSES through the element's [5][64] tile: the absolute sum does not care which thread holds which value, so
    // after the write of the zeta-transformed lines thread (i, j) takes the xi-line (., j, k = i), transforms it in registers, writes it
    // back and takes the eta-line (a = j, ., k = i): 4 reads per variable and direction instead of 16 (every value is read ONCE)
    const int p0 = tPair(i, j, 0), p1 = tPair(i, j, 1);
    __syncwarp(wm);
#pragma unroll
    for (int v = 0; v < 5; v++) { sts2(sXe + v * 64 + p0, R[v][0], R[v][1]); sts2(sXe + v * 64 + p1, R[v][2], R[v][3]); }
    __syncwarp(wm);
#pragma unroll
    for (int v = 0; v < 5; v++) {
      double x[4];
#pragma unroll
      for (int a = 0; a < 4; a++) x[a] = sXe[v * 64 + tIdx(a, j, i)];
#pragma unroll
      for (int q = 0; q < 4; q++) R[v][q] = A.k1[0 * 4 + q] * x[0] + A.k1[1 * 4 + q] * x[1] + A.k1[2 * 4 + q] * x[2] + A.k1[3 * 4 + q] * x[3];
    }
    __syncwarp(wm);
#pragma unroll
    for (int v = 0; v < 5; v++)
#pragma unroll
      for (int a = 0; a < 4; a++) sXe[v * 64 + tIdx(a, j, i)] = R[v][a];
    __syncwarp(wm);
#pragma unroll
    for (int v = 0; v < 5; v++) {
      double x[4];
#pragma unroll
      for (int b = 0; b < 4; b++) x[b] = sXe[v * 64 + tIdx(j, b, i)];
#pragma unroll
      for (int q = 0; q < 4; q++) R[v][q] = A.k1[0 * 4 + q] * x[0] + A.k1[1 * 4 + q] * x[1] + A.k1[2 * 4 + q] * x[2] + A.k1[3 * 4 + q] * x[3];
    }
#else
    const int p0 = tPair(i, j, 0), p1 = tPair(i, j, 1);
#pragma unroll
    for (int d = 0; d < 2; d++) {   // xi, then eta: all five variables through the element's [5][64] tile
      __syncwarp(wm);
#pragma unroll
      for (int v = 0; v < 5; v++) { sts2(sXe + v * 64 + p0, R[v][0], R[v][1]); sts2(sXe + v * 64 + p1, R[v][2], R[v][3]); }
      __syncwarp(wm);
#pragma unroll
      for (int v = 0; v < 5; v++) {
        double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
#pragma unroll
        for (int a = 0; a < 4; a++) {
          const double* p = sXe + v * 64;
          const double kk = A.k1[a * 4 + (d == 0 ? i : j)];
          const double2 lo = lds2(p + (d == 0 ? tPair(a, j, 0) : tPair(i, a, 0))), hi = lds2(p + (d == 0 ? tPair(a, j, 1) : tPair(i, a, 1)));
          x0 += kk * lo.x; x1 += kk * lo.y; x2 += kk * hi.x; x3 += kk * hi.y;
        }
        R[v][0] = x0; R[v][1] = x1; R[v][2] = x2; R[v][3] = x3;
      }
    }
#endif
    // deterministic block reduction: warp shuffle (inactive lanes contribute zero), then one thread sums the warp partials in order
#pragma unroll
    for (int v = 0; v < 5; v++) {
      double s = fabs(R[v][0]) + fabs(R[v][1]) + fabs(R[v][2]) + fabs(R[v][3]);
      for (int o = 16; o > 0; o >>= 1) { const double y = __shfl_down_sync(wm, s, o); if (((tid & 31) + o < 32) && ((wm >> ((tid & 31) + o)) & 1)) s += y; }
      if ((tid & 31) == 0) sRed[(tid >> 5) * 5 + v] = s;
    }
    __syncthreads();
    if (tid < 5) {
      double s = 0.0;
      for (int w = 0; w < (ne * 16 + 31) / 32; w++) s += sRed[w * 5 + tid];
      A.normPartial[(size_t)chunk * 5 + tid] = s / 64;
    }
  }
}

}  // namespace sdg
