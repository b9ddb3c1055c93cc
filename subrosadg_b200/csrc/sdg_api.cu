// sdg_api.cu — C ABI (include/subrosadg_b200.h) over the sm_100a kernels: context, uploads, stage sequencing.
//
// Mirrors what System<SC>::solve() asks of Solver<SC> (src/Utils/SystemControl.cpp:159-195): initializeSolver,
// calculateDeltaTime, stepSolver, relative_error_.  There is no CPU fallback anywhere in this file: without a CUDA
// device every compute entry point fails with an error message.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>

#include "../../include/subrosadg_b200.h"
#include "dev_util.cuh"
#include "host_plan.hpp"
#include "launchers.hpp"
#include "mixed_path.hpp"
#include "nsl_kernels.cuh"
#include "tensor_kernels.cuh"
#include <functional>
#include "av_kernels.cuh"
#include "view_variable.hpp"

using namespace sdg;

namespace {

thread_local std::string g_err;

}  // namespace

namespace sdg { void setLastError(const char* msg) { g_err = msg; } }   // for the other translation units behind sdg_last_error()

struct sdg_ctx {
  sdg_config cfg{};
  PhysParams phys{};
  int D = 0, NV = 0, nStages = 3;
  double rkc[3][3]{};
  bool hasDevice = false, finalized = false;
  MeshPlan plan;
  bool haveBlock = false, haveFaces = false;
  cudaStream_t stream = nullptr, copyStream = nullptr;
  // ShockCapturingEnum::ArtificialViscosity: Solver::empirical_tolerance_, artificial_viscosity_factor_, Mesh::node_number_; the mesh data of
  // the block in caller order (node_tag_ of the corner nodes, inner_radius_); device copies in internal order
  double avTol = 0.0, avFactor = 1.0; int avNodes = 0;
  std::vector<int> avTagsHost; std::vector<double> avRadiusHost;
  DevBuf<int> avTags; DevBuf<double> avRadius, avElem, avTabQ, avTabF, avH, avNode, avE;
  cudaEvent_t seamEvent[9] = {};
  int64_t launches = 0;

  // device state
  DevBuf<double> U[3], geoE, invjw, minEdge, geoF, dummy, Phi, PhiInv, PhiT, M1[3], normPartial, normSlices, normOut, dtPartial, scratch, sendBuf, cfGeo;
  DevBuf<int> perm, faceRec, chunkOff, chunkInterior, chunkBoundary, sendList, lexOf;
  DevBuf<TensorDev> tab;
  int cur = 0;          // index of the buffer holding the current state
  // buffer that held the INPUT of the last RK stage of the last completed step (-1: no step since the state was last set).  The reference
  // writes variable_gradient_basis_function_coefficient_ as the last stage left it (Solver::writeRawBinary after stepSolver,
  // SystemControl.cpp:166-183): the gradient of that stage's input, not of the new state; the gradient getters evaluate the same state.
  int gradSrc = -1;
  int latest = 0;       // buffer written last (halo source / target)
  StageFn eulerFn = nullptr, nsGradFn = nullptr, nsStageFn = nullptr;
  DevBuf<double> G, G2;   // NS: gradient field [n][NV*D][NN] (+ scratch for diagnostics)
  double stepDt = 0.0;
  int nSend = 0;
  std::vector<double> hostNorm;
  // one time step (nStages x passes launches) captured as a CUDA graph: launch-bound meshes (the reference's shipped configs have
  // 1e2-3e4 elements) replay it instead of issuing every launch from the host
  cudaGraphExec_t stepGraph = nullptr; double graphDt = 0.0; int graphCur = -1; bool graphWarm = false; int64_t graphLaunches = 0;
  // peer-memory halo exchange (CUDA IPC): the peers' arrays opened in this process, arrival flags, exchange counter
  std::vector<PeerDev> peerLinks; std::vector<void*> ipcOpened;
  bool rowHalo = false; int nRecvRows = 0;   // trace-row halo (sdg_set_halo_rows): units = (element, face) rows of TU / TV
  DevBuf<int> recvList; DevBuf<double> recvBuf; DevBuf<long long> dstUnit;
  DevBuf<PeerDev> peerDev; DevBuf<long long> ipcFlags; DevBuf<unsigned int> pushCounter; DevBuf<int> haloErr;
  long long pushEpoch = 0;
  std::unique_ptr<MixedSolver> mx;   // dense-operator path: meshes with triangle blocks / several element types (mixed_path.cu)
  // trace-based line kernels for P3 hexahedra (nsl_kernels.cuh): published face traces TU (one per stage buffer) and TV, virtual
  // neighbour traces of the boundary faces, link records; traceValid[b] = TU[b] holds the traces of U[b]
  bool lineTrace = false;   // link plan + nsl kernels
  bool traceTU = false;     // ... with published face traces (TU / TV arrays, trace rows in the halo)
  LineFns lineFns;
  DevBuf<double> TU[3], TV, TUb, lfGeo;
  DevBuf<int> links, bndRec;
  DevBuf<LineTabDev> ltab;
  double w1[kMaxN]{}, cLift = 0.0;
  LinePlan linePlan;   // host image (diagnostics: sdg_debug_plan 20-23)
  bool traceValid[3] = {false, false, false};
  int64_t stepCount = 0, bndKey = -1;
  // sdg_step_host: one time step streamed through host buffers (upload, stages and download overlapped per dependency level)
  const int* listOverride = nullptr; int listCount = 0;   // runStage: explicit chunk list of the launch (device pointer)
  struct HostPipe {
    int G = 0;                                   // upload groups = contiguous ranges of the caller's element order
    std::vector<int> first;                      // [G + 1] caller-order element ranges
    std::vector<int> off;                        // [(launches + 1) * G + 1]: chunk lists of (kind, level); kind 0 = traces, 1.. = the launches of a step in order
    std::vector<std::vector<int>> download;      // per level: the upload groups whose elements have all finished the last stage
    DevBuf<int> lists, bndLists; DevBuf<double> up, down;
    std::vector<int> bndOff;                     // [nStages * G + 1]: boundary faces by (stage, level of the parent chunk's gradient pass)
    std::vector<cudaEvent_t> upEv, downEv, doneEv;   // doneEv + t0: SDG_HOST_PIPE_TIMING=1 only (time line of the copies on stderr)
    cudaEvent_t t0 = nullptr; bool timing = false;
    cudaStream_t d2h = nullptr;
    double overlap = 0.0;                        // fraction of the upload groups downloaded before the last level (diagnostic)
  } pipe;

  size_t stateDoubles() const { return (size_t)plan.blk.n * NV * plan.blk.T.NN; }
  size_t elemDoubles() const { return (size_t)NV * plan.blk.T.NN; }
};

namespace {

bool twoPass(const sdg_ctx* c);
void needDevice(sdg_ctx* c) { if (!c->hasDevice) throw std::runtime_error("no CUDA device bound to this context (plan-only context); the product has no CPU path"); }
void needFinal(sdg_ctx* c) { if (!c->finalized) throw std::runtime_error("sdg_finalize has not been called"); }
// a stage buffer that holds neither the current state nor the input of the last stage (see gradSrc)
int scratchBuffer(const sdg_ctx* c) { const int a = (c->cur + 1) % 3; return a == c->gradSrc ? (c->cur + 2) % 3 : a; }
void needType(sdg_ctx* c, int type) { if (!c->haveBlock || c->plan.blk.type != type) throw std::runtime_error("no element block of this type"); }

void fillArgs(sdg_ctx* c, StageArgs& a) {
  const BlockPlan& B = c->plan.blk;
  a = StageArgs{};
  a.geoE = c->geoE.p; a.invjw = c->invjw.p; a.geoF = c->geoF.p; a.cfGeo = c->cfGeo.p;
  a.faceRec = reinterpret_cast<const int4*>(c->faceRec.p); a.chunkFaceOff = c->chunkOff.p; a.chunkList = nullptr;
  a.dummy = c->dummy.p; a.tab = c->tab.p; a.normPartial = nullptr;
  a.nOwned = B.nOwned; a.nInt = c->plan.F.nInt; a.mode = 0; a.faceSel = -1; a.phys = c->phys;
  a.avElem = c->avElem.p; a.avTabQ = c->avTabQ.p; a.avTabF = c->avTabF.p;
  const int N = B.T.N;
  for (int i = 0; i < N * N; i++) { a.dm[i] = B.T.Dm[i]; a.k1[i] = B.T.K1[i]; }
  for (int i = 0; i < 2 * N; i++) a.lend[i] = B.T.Lend[i];
  for (int i = 0; i < N; i++) a.w1[i] = B.T.w[i];
  if (c->lineTrace) {
    a.TVin = c->TV.p; a.TVout = c->TV.p; a.TUb = c->TUb.p;
    a.links = reinterpret_cast<const int4*>(c->links.p); a.lfGeo = c->lfGeo.p; a.ltab = c->ltab.p; a.cLift = c->cLift;
    static const int ahead = std::max(1, getenv("SDG_AHEAD") ? atoi(getenv("SDG_AHEAD")) : 1 << 30);   // default: every block fetches its OWN ranges; fetching one wave ahead (444, 148) measured slower: 1.90 / 1.66 vs 1.42 ms per residual pass at 64^3 (L2 churn)
    a.ahead = ahead;
  }
}

bool twoPass(const sdg_ctx* c) { return c->phys.ns != 0 || c->phys.av != 0; }

// Solver::calculateArtificialViscosity for the state in buffer `buf` (TimeIntegration.cpp:336-338: once per step, before the stages)
void avUpdate(sdg_ctx* c, int buf, cudaStream_t st) {
  static const double kTol[5] = {0.0, -1.20411998266, -1.90848501888, -2.40823996531, -2.79588001734};   // SimulationControl.cpp:892-893
  const BlockPlan& B = c->plan.blk;
  const int NB = 1 << c->D, n = B.nOwned;
  const int REC = (c->D * c->D + 2) & ~1;
  avIndicatorKernel<<<std::min((n + 7) / 8, 148 * 8), 256, 0, st>>>(c->U[buf].p, c->avH.p, c->geoE.p, c->invjw.p, c->tab.p->wq, B.affine ? 1 : 0, REC, c->D * c->D, n,
                                                                    c->NV, B.T.NN, c->cfg.p, kTol[c->cfg.p - 1], c->avTol, c->avFactor, c->avRadius.p, c->avE.p);
  CUDA_OK(cudaMemsetAsync(c->avNode.p, 0, c->avNode.n * sizeof(double), st));
  const int blocks = std::min((n * NB + 255) / 256, 148 * 8);
  avNodeMaxKernel<<<blocks, 256, 0, st>>>(c->avE.p, c->avTags.p, n, NB, reinterpret_cast<unsigned long long*>(c->avNode.p));
  // corner values of EVERY element of the context: between partitions the node array is first completed by a max-reduction over the
  // ranks (sdg_av_node_buffer / sdg_av_store), and the ghost elements' values enter the right-hand side of the cut faces
  avStoreKernel<<<std::min((B.n * NB + 255) / 256, 148 * 8), 256, 0, st>>>(c->avNode.p, c->avTags.p, B.n, NB, c->avElem.p);
  c->launches += 3;
  CUDA_OK(cudaGetLastError());
}

// traces of a state that did not come out of the residual pass (initial condition, state setters)
void ensureTraces(sdg_ctx* c, int buf, cudaStream_t s) {
  if (!c->traceTU || c->traceValid[buf]) return;
  StageArgs a; fillArgs(c, a);
  a.Uin = c->U[buf].p; a.TUout = c->TU[buf].p;
  c->lineFns.trace(a, c->plan.blk.nChunks, s);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  c->traceValid[buf] = true;
}

// one pass of one stage over a subset of the chunks
void runStage(sdg_ctx* c, const StageArgs& base, int part, cudaStream_t s, int pass = -1) {
  const BlockPlan& B = c->plan.blk;
  StageArgs a = base;
  int nBlocks = B.nChunks;
  if (part == 0) { a.chunkList = c->chunkInterior.p; nBlocks = (int)B.chunkInterior.size(); }
  else if (part == 1) { a.chunkList = c->chunkBoundary.p; nBlocks = (int)B.chunkBoundary.size(); }
  if (c->listOverride) { a.chunkList = c->listOverride; nBlocks = c->listCount; }
  if (nBlocks == 0) return;
  if (c->lineTrace) {
    if (pass == 0) c->lineFns.grad(a, nBlocks, s); else c->lineFns.stage(a, nBlocks, s);
  }
  else if (!twoPass(c)) c->eulerFn(a, nBlocks, s);
  else if (pass == 0) c->nsGradFn(a, nBlocks, s);
  else c->nsStageFn(a, nBlocks, s);
  c->launches++;
  CUDA_OK(cudaGetLastError());
}

// virtual neighbour traces of the boundary faces for the gradient pass reading TU[in]: once per (step, stage), before the first part
void lineBoundary(sdg_ctx* c, const StageArgs& a, int64_t key, cudaStream_t s) {
  if (c->plan.F.nBnd == 0 || (key >= 0 && c->bndKey == key)) return;
  c->lineFns.boundary(a, reinterpret_cast<const int4*>(c->bndRec.p), c->plan.F.nBnd, s);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  c->bndKey = key;
}

// buffers of stage s: in / out indices (SSP-RK tables TimeIntegration.cpp:45-65 realised with three rotating buffers)
void stageBuffers(sdg_ctx* c, int s, int& in, int& out) {
  const int cur = c->cur, a = (cur + 1) % 3, b = (cur + 2) % 3;
  if (c->nStages == 1) { in = cur; out = a; return; }
  if (s == 0) { in = cur; out = a; }
  else if (s == c->nStages - 1) { in = (s == 1) ? a : b; out = cur; }
  else { in = a; out = b; }
}

// pass: -1 = every pass of the stage (single GPU); 0 = gradient pass (NS); 1 = residual pass (the only pass for Euler)
void stageLaunch(sdg_ctx* c, int s, int part, cudaStream_t st, int pass = -1) {
  int in, out; stageBuffers(c, s, in, out);
  StageArgs a; fillArgs(c, a);
  a.Uin = c->U[in].p; a.Ulast = c->U[c->cur].p; a.Uout = c->U[out].p;
  a.Gvol = c->G.p; a.Gout = c->G.p;
  if (c->lineTrace) {
    ensureTraces(c, in, st);
    a.TUin = c->TU[in].p; a.TUout = c->TU[out].p;
    if (twoPass(c) && (pass == -1 || pass == 0)) lineBoundary(c, a, c->stepCount * 4 + s, st);
  }
  if (twoPass(c) && (pass == -1 || pass == 0)) { runStage(c, a, part, st, 0); if (pass == 0) return; }
  if (!twoPass(c) && pass == 0) return;
  a.aLast = s == 0 ? 0.0 : c->rkc[s][0];
  a.aCur = s == 0 ? 1.0 : c->rkc[s][1];
  a.bdt = c->rkc[s][2] * c->stepDt;
  a.normPartial = s == c->nStages - 1 ? c->normPartial.p : nullptr;
  runStage(c, a, part, st, 1);
  c->latest = out;
  if (c->traceTU) c->traceValid[out] = true;   // the residual pass publishes the traces of the state it writes
}

// where the input of the last stage of the step that has just finished lives (see sdg_ctx::gradSrc); `cur` is already rotated
void markLastStageInput(sdg_ctx* c) {
  if (c->nStages == 1) { c->gradSrc = (c->cur + 2) % 3; return; }
  int in, out; stageBuffers(c, c->nStages - 1, in, out);
  c->gradSrc = in;
}
void finishStep(sdg_ctx* c) { if (c->nStages == 1) c->cur = (c->cur + 1) % 3; c->latest = c->cur; c->stepCount++; markLastStageInput(c); }

void reduceNorm(sdg_ctx* c, double* sums) {
  constexpr int kSlices = 128;
  normReduceKernel<<<dim3(c->NV, kSlices), 256, 0, c->stream>>>(c->normPartial.p, c->plan.blk.nChunks, c->NV, c->normSlices.p);
  normReduceKernel<<<dim3(c->NV, 1), 128, 0, c->stream>>>(c->normSlices.p, kSlices, c->NV, c->normOut.p);
  c->launches += 2;
  CUDA_OK(cudaMemcpyAsync(c->hostNorm.data(), c->normOut.p, sizeof(double) * c->NV, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  for (int v = 0; v < c->NV; v++) sums[v] = c->hostNorm[v];
}

}  // namespace

#define SDG_TRY try {
#define SDG_CATCH                                                          \
  }                                                                        \
  catch (const std::exception& ex) { g_err = ex.what(); return 1; }        \
  catch (...) { g_err = "unknown error"; return 1; }                       \
  return 0;

extern "C" {

const char* sdg_last_error(void) { return g_err.c_str(); }
int sdg_version(void) { return 100; }

int sdg_create(const sdg_config* cfg, sdg_ctx** out) {
  SDG_TRY
  if (!cfg || !out) throw std::runtime_error("null argument");
  auto c = std::make_unique<sdg_ctx>();
  c->cfg = *cfg;
  if (cfg->dim < 1 || cfg->dim > 3) throw std::runtime_error("dim must be 1, 2 or 3");
  if (cfg->p < 1 || cfg->p > 5) throw std::runtime_error("polynomial order must be 1..5 (PolynomialOrderEnum P1..P5)");
  c->D = cfg->dim; c->NV = cfg->dim + 2;
  PhysParams& P = c->phys;
  P.model = cfg->model; P.eos = cfg->eos; P.transport = cfg->transport; P.conv = cfg->conv_flux; P.visc = cfg->visc_flux; P.source = cfg->source;
  P.compressible = (cfg->model == kCompresibleEuler || cfg->model == kCompresibleNS) ? 1 : 0;
  P.ns = (cfg->model == kCompresibleNS || cfg->model == kIncompresibleNS) ? 1 : 0;
  P.cp = cfg->cp; P.cv = cfg->cv; P.icv = 1.0 / cfg->cv; P.kg = 0.5 * (1.4 + 1.0) / 1.4; P.gamma = 1.4;  // EquationOfState<IdealGas>::kSpecificHeatRatio, PhysicalModel.cpp:45
  P.mu0 = cfg->mu; P.k0 = cfg->cp * cfg->mu / 0.71;  // Pr = 0.71, PhysicalModel.cpp:152-156
  P.c0 = cfg->c0; P.rho0 = cfg->rho0; P.padd = 0.01 * cfg->rho0 * cfg->c0 * cfg->c0;  // :63-66
  P.beta = cfg->beta; P.tref = cfg->t_ref;
  if (cfg->model < 0 || cfg->model > 3) throw std::runtime_error("equation model not supported");
  if (P.ns && (cfg->visc_flux != kBR1 && cfg->visc_flux != kBR2)) throw std::runtime_error("Navier-Stokes needs ViscousFluxEnum::BR1 or BR2");
  if (!P.ns) P.visc = kViscNone;
  if ((cfg->conv_flux == kHLLC || cfg->conv_flux == kRoe) && cfg->eos != kIdealGas) throw std::runtime_error("HLLC/Roe need the ideal-gas EOS (reference uses kSpecificHeatRatio)");
  if (cfg->conv_flux < 0 || cfg->conv_flux > 4) throw std::runtime_error("bad convective flux");
  // TimeIntegrationData, TimeIntegration.cpp:45-65: {a_last, a_cur, b}
  const double FE[1][3] = {{1.0, 0.0, 1.0}};
  const double H2[2][3] = {{1.0, 0.0, 1.0}, {0.5, 0.5, 0.5}};
  const double S3[3][3] = {{1.0, 0.0, 1.0}, {3.0 / 4.0, 1.0 / 4.0, 1.0 / 4.0}, {1.0 / 3.0, 2.0 / 3.0, 2.0 / 3.0}};
  if (cfg->rk == kForwardEuler) { c->nStages = 1; std::memcpy(c->rkc, FE, sizeof(FE)); }
  else if (cfg->rk == kHeunRK2) { c->nStages = 2; std::memcpy(c->rkc, H2, sizeof(H2)); }
  else if (cfg->rk == kSSPRK3) { c->nStages = 3; std::memcpy(c->rkc, S3, sizeof(S3)); }
  else throw std::runtime_error("bad time integration scheme");
  if (cfg->device >= 0) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= cfg->device) throw std::runtime_error(std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "ordinal out of range") + " — this library has no CPU path");
    CUDA_OK(cudaSetDevice(cfg->device));
    CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->hasDevice = true;
  }
  c->hostNorm.assign(8, 0.0);
  *out = c.release();
  SDG_CATCH
}

void sdg_destroy(sdg_ctx* c) {
  if (!c) return;
  if (c->hasDevice) { cudaSetDevice(c->cfg.device); cudaDeviceSynchronize(); }
  if (c->stepGraph) cudaGraphExecDestroy(c->stepGraph);
  for (void* q : c->ipcOpened) cudaIpcCloseMemHandle(q);
  if (c->copyStream) { cudaStreamDestroy(c->copyStream); for (cudaEvent_t e : c->seamEvent) if (e) cudaEventDestroy(e); }
  if (c->pipe.d2h) cudaStreamDestroy(c->pipe.d2h);
  for (cudaEvent_t e : c->pipe.upEv) cudaEventDestroy(e);
  for (cudaEvent_t e : c->pipe.downEv) cudaEventDestroy(e);
  for (cudaEvent_t e : c->pipe.doneEv) cudaEventDestroy(e);
  if (c->pipe.t0) cudaEventDestroy(c->pipe.t0);
  cudaStream_t s = c->stream; const bool dev = c->hasDevice;
  delete c;
  if (dev && s) cudaStreamDestroy(s);
}

int sdg_add_elements(sdg_ctx* c, int32_t type, int32_t n, int32_t n_ghost, int32_t geom_order, const double* coords) {
  SDG_TRY
  if (c->finalized) throw std::runtime_error("context already finalized");
  if (!(type == kLine || type == kTriangle || type == kQuadrangle || type == kHexahedron)) throw std::runtime_error("device path implements line, triangle, quadrangle and hexahedron blocks");
  if (elemDim(type) != c->D) throw std::runtime_error("element dimension mismatch");
  if (n <= 0 || n_ghost < 0 || n_ghost >= n || geom_order < 1 || geom_order > 5) throw std::runtime_error("bad element block arguments");
  // Single quadrangle / hexahedron block: collocation tensor path.  Triangles, several element types in one mesh, P1 quadrangles (Gmsh's
  // "Gauss2" on a quadrangle is a seven-point rule, not the 2 x 2 tensor rule: mixed_tables.hpp) or cfg.chunk == -1 (diagnostics):
  // dense-operator path in the reference's modal representation (mixed_path.cu).
  if (type == kTriangle || c->haveBlock || c->mx || c->cfg.chunk == -1 || (type == kQuadrangle && c->cfg.p == 1)) {
    if (c->D != 2) throw std::runtime_error("several element types in one mesh: 2-D (triangle / quadrangle) only");
    if (!c->mx) {
      c->mx = std::make_unique<MixedSolver>(c->cfg.p, c->phys, c->nStages, c->rkc, c->stream, c->hasDevice, c->cfg.device);
      if (c->phys.av) c->mx->setArtificialViscosity(c->avTol, c->avFactor, c->avNodes);
    }
    if (c->haveBlock) {  // a tensor block was registered first: hand it over
      BlockPlan& B = c->plan.blk;
      c->mx->addBlock(B.type, B.n, B.nGhost, B.g, B.X.data());
      B = BlockPlan{}; c->haveBlock = false;
    }
    c->mx->addBlock(type, n, n_ghost, geom_order, coords);
    return 0;
  }
  BlockPlan& B = c->plan.blk;
  B.type = type; B.D = c->D; B.p = c->cfg.p; B.g = geom_order; B.n = n; B.nGhost = n_ghost; B.nOwned = n - n_ghost;
  B.T = buildTensorTables(type, c->cfg.p);
  B.nn = (int)gmshNodeLattice(type, geom_order).size();
  B.X.assign(coords, coords + (size_t)n * B.nn * c->D);
  c->plan.D = c->D; c->plan.p = c->cfg.p;
  c->haveBlock = true;
  SDG_CATCH
}

int sdg_set_faces(sdg_ctx* c, int32_t n_int, int32_t n_bnd, const int32_t* le, const int32_t* lt, const int32_t* lf, const int32_t* re,
                  const int32_t* rt, const int32_t* rf, const int32_t* rot, const int32_t* bc, const int32_t* phys) {
  SDG_TRY
  if (c->finalized) throw std::runtime_error("context already finalized");
  FaceInput& F = c->plan.F;
  const int nf = n_int + n_bnd;
  F.nInt = n_int; F.nBnd = n_bnd;
  F.le.assign(le, le + nf); F.lt.assign(lt, lt + nf); F.lf.assign(lf, lf + nf);
  F.re.assign(re, re + nf); F.rt.assign(rt, rt + nf); F.rf.assign(rf, rf + nf);
  F.rot.assign(rot, rot + nf); F.bc.assign(bc, bc + nf); F.phys.assign(phys, phys + nf);
  c->haveFaces = true;
  SDG_CATCH
}

int sdg_finalize(sdg_ctx* c) {
  SDG_TRY
  if (c->finalized) throw std::runtime_error("context already finalized");
  if (c->mx) {
    if (!c->haveFaces) throw std::runtime_error("elements and faces must be set before sdg_finalize");
    c->mx->setFaces(c->plan.F); c->mx->finalize(); c->finalized = true;
    return 0;
  }
  if (!c->haveBlock || !c->haveFaces) throw std::runtime_error("elements and faces must be set before sdg_finalize");
  MeshPlan& M = c->plan; BlockPlan& B = M.blk; const FaceInput& F = M.F;
  const int nf = F.nInt + F.nBnd;
  for (int i = 0; i < nf; i++) {
    const bool interior = i < F.nInt;
    if (F.lt[i] != B.type || F.le[i] < 0 || F.le[i] >= B.n || F.lf[i] < 0 || F.lf[i] >= B.T.NF) throw std::runtime_error("face record out of range (left parent)");
    if (interior && (F.rt[i] != B.type || F.re[i] < 0 || F.re[i] >= B.n || F.rf[i] < 0 || F.rf[i] >= B.T.NF || F.rot[i] < 0 || F.rot[i] > 3)) throw std::runtime_error("face record out of range (right parent)");
    if (!interior && (F.bc[i] < 0 || F.bc[i] > 5)) throw std::runtime_error("boundary face without a boundary condition (Periodic is not a boundary type)");
  }
  const int N = B.T.N;
  int K = 0;
  const int ph = (c->phys.compressible && c->phys.eos == kIdealGas && c->phys.conv == kHLLC) ? 1 : 0;
  // decide the affine flag first (needs geometry), then the kernel
  M.buildBlock(c->cfg.reorder, 1);  // provisional chunk size; chunking is redone below once K is known
  // P3 hexahedra: Navier-Stokes runs on the trace-based line kernels (SDG_NS_NODE_KERNEL = A/B switch back to the node-per-thread
  // kernels of ns_kernels.cuh); SDG_EULER_TRACE routes Euler through the same residual pass for comparison with eulerLineKernel
  // Euler on P3 hexahedra: SDG_EULER_KERNEL = line (eulerLineKernel, chunk face lists) | link (the residual pass of the line NS kernels
  // without viscous terms, partners gathered from their nodal states) | trace (the same with published traces: twice the HBM traffic)
  const char* ek = getenv("SDG_EULER_KERNEL");
  const std::string eulerKernel = ek ? ek : "trace";   // measured at 128^3 (ms per stage): trace 8.12, line 8.80, link 8.83; 2 GPUs: 6.14 vs 6.85
  // artificial viscosity runs on the node-per-thread kernels of ns_kernels.cuh (gradient pass + residual pass with eps * grad(U))
  c->lineTrace = c->D == 3 && N == 4 && !c->phys.av && (c->phys.ns ? getenv("SDG_NS_NODE_KERNEL") == nullptr : eulerKernel != "line");
  c->traceTU = c->lineTrace && (c->phys.ns || eulerKernel == "trace");
  if (c->lineTrace) pickNslFns(B.affine, ph, c->phys.ns != 0, !c->traceTU, c->lineFns, K);
  else if (twoPass(c)) pickNsFn(c->D, N, B.affine, ph, c->nsGradFn, c->nsStageFn, K);
  else c->eulerFn = pickEulerFn(c->D, N, B.affine, ph, K);
  if (c->cfg.chunk > 0 && c->cfg.chunk != K) throw std::runtime_error("chunk override not available: kernels are compiled for K = " + std::to_string(K));
  B.K = K; B.nChunks = (B.nOwned + K - 1) / K;
  M.buildFaces(nullptr, false);
  M.buildChunkFaces();
  // the compile-time face direction / side tables of the kernels must agree with the numerically derived ones
  for (int f = 0; f < B.T.NF; f++) {
    const int dirRef = c->D == 1 ? 0 : c->D == 2 ? ((0x1 | 0x0 << 2 | 0x1 << 4 | 0x0 << 6) >> (2 * f)) & 3 : ((0x2 | 0x1 << 2 | 0x0 << 4 | 0x0 << 6 | 0x1 << 8 | 0x2 << 10) >> (2 * f)) & 3;
    const int sideRef = c->D == 1 ? f : c->D == 2 ? ((0x6 >> f) & 1) : (f >= 3);
    if (dirRef != B.T.faceDir[f] || sideRef != B.T.faceSide[f]) throw std::runtime_error("internal: face direction table mismatch");
  }
  if (c->lineTrace) c->linePlan = M.buildLinePlan();
  if (c->hasDevice) {
    CUDA_OK(cudaSetDevice(c->cfg.device));
    const size_t nd = c->stateDoubles();
    for (int i = 0; i < 3; i++) { c->U[i].alloc(nd); CUDA_OK(cudaMemsetAsync(c->U[i].p, 0, nd * sizeof(double), c->stream)); }
    if (B.affine) {
      // device images for the TMA staging of the stage kernel: element metric records padded to an even number of doubles,
      // face geometry duplicated per chunk-face entry (one contiguous 16-byte-granular range per thread block)
      const int DD = c->D * c->D, REC = (DD + 2) & ~1;
      std::vector<double> ge((size_t)B.n * REC, 0.0);
      for (int pos = 0; pos < B.n; pos++) for (int k = 0; k <= DD; k++) ge[(size_t)pos * REC + k] = B.geoE[(size_t)pos * (DD + 1) + k];
      c->geoE.upload(ge, c->stream);
      const size_t ncf = B.faceRec.size() / 4;
      std::vector<double> cf(std::max<size_t>(ncf, 1) * kCF, 0.0);
      for (size_t k = 0; k < ncf; k++) {
        for (int l = 0; l <= c->D; l++) cf[k * kCF + l] = M.geoF[(size_t)B.faceRec[k * 4 + 2] * (c->D + 1) + l];
        const int pL = B.faceRec[k * 4 + 0], pR = B.faceRec[k * 4 + 1];
        cf[k * kCF + c->D + 1] = 1.0 / B.geoE[(size_t)pL * (DD + 1) + DD];
        cf[k * kCF + c->D + 2] = pR >= 0 ? 1.0 / B.geoE[(size_t)pR * (DD + 1) + DD] : 0.0;
      }
      c->cfGeo.upload(cf, c->stream);
    } else {
      c->geoE.upload(B.geoE, c->stream);
    }
    c->invjw.upload(B.invjw, c->stream); c->minEdge.upload(B.minEdge, c->stream);
    c->geoF.upload(M.geoF, c->stream);
    c->dummy.alloc((size_t)std::max(F.nBnd, 1) * (c->D + 3) * B.T.NQF);
    CUDA_OK(cudaMemsetAsync(c->dummy.p, 0, c->dummy.n * sizeof(double), c->stream));
    c->perm.upload(B.perm, c->stream);
    c->faceRec.upload(B.faceRec, c->stream); c->chunkOff.upload(B.chunkFaceOff, c->stream);
    c->chunkInterior.upload(B.chunkInterior, c->stream); c->chunkBoundary.upload(B.chunkBoundary, c->stream);
    c->Phi.upload(B.T.Phi, c->stream); c->PhiInv.upload(B.T.PhiInv, c->stream);
    { std::vector<double> PT(B.T.Phi.size()); const int NN = B.T.NN; for (int q = 0; q < NN; q++) for (int b = 0; b < NN; b++) PT[(size_t)b * NN + q] = B.T.Phi[(size_t)q * NN + b]; c->PhiT.upload(PT, c->stream); }
    {   // 1-D factors of the modal transforms (tensorTransformKernel): Phi1, Phi1^-1, Phi1^T, and the lexicographic index of every mode
      const int N = B.T.N, D = B.T.D;
      std::vector<double> inv = B.T.Phi1, tr((size_t)N * N);
      invertDense(inv, N);
      for (int a = 0; a < N; a++) for (int k = 0; k < N; k++) tr[(size_t)k * N + a] = B.T.Phi1[(size_t)a * N + k];
      c->M1[0].upload(B.T.Phi1, c->stream); c->M1[1].upload(inv, c->stream); c->M1[2].upload(tr, c->stream);
      std::vector<int> lex(B.T.NN);
      for (int b = 0; b < B.T.NN; b++) { int q = 0; for (int d = 0; d < D; d++) q = q * N + B.T.modalIdx[b][d]; lex[b] = q; }
      c->lexOf.upload(lex, c->stream);
    }
    if (twoPass(c)) { c->G.alloc((size_t)B.n * c->NV * c->D * B.T.NN); CUDA_OK(cudaMemsetAsync(c->G.p, 0, c->G.n * sizeof(double), c->stream)); }
    c->normPartial.alloc((size_t)B.nChunks * c->NV); c->normSlices.alloc(128 * 8); c->normOut.alloc(8); c->dtPartial.alloc(1024);
    CUDA_OK(cudaMemsetAsync(c->normPartial.p, 0, c->normPartial.n * sizeof(double), c->stream));
    std::vector<TensorDev> td(1);
    TensorDev& t = td[0]; std::memset(&t, 0, sizeof(t));
    for (int i = 0; i < N * N; i++) { t.Dm[i] = B.T.Dm[i]; t.K1[i] = B.T.K1[i]; }
    for (int i = 0; i < 2 * N; i++) t.Lend[i] = B.T.Lend[i];
    for (int i = 0; i < B.T.NN; i++) t.wq[i] = B.T.wq[i];
    for (int i = 0; i < B.T.NQF; i++) t.wf[i] = B.T.wf[i];
    for (int f = 0; f < B.T.NF; f++) { t.faceDir[f] = B.T.faceDir[f]; t.faceSide[f] = B.T.faceSide[f]; }
    for (int i = 0; i < B.T.NF * B.T.NQF; i++) t.faceBase[i] = B.T.faceBase[i];
    for (int i = 0; i < B.T.NF * B.T.NN; i++) t.nodeFacePt[i] = (unsigned char)B.T.nodeFacePt[i];
    const int ft = faceType(B.type);
    for (int r = 0; r < 4; r++) {
      std::vector<int> s = faceSequence(ft, N, ft == kLine ? 0 : r);
      for (int j = 0; j < B.T.NQF; j++) t.seq[r * B.T.NQF + j] = s[j];
    }
    c->tab.upload(td, c->stream);
    if (c->phys.av) {
      // H = Phi[:, high] Phi^-1[high, :] (the part of a nodal field carried by the modes above order P-1; P1: every mode), the order-1 nodal
      // basis at the volume nodes and at the face points (nodal_value_ / nodal_adjacency_value_, BasisFunction.cpp:149-208), mesh data by position
      const int NN = B.T.NN, D = c->D, NB = 1 << D, p = c->cfg.p;
      int nbLow = 1; for (int d = 0; d < D; d++) nbLow *= p;       // getElementBasisFunctionNumber<type, P - 1> of line / quadrangle / hexahedron
      if (p == 1) nbLow = 0;
      std::vector<double> H((size_t)NN * NN, 0.0);
      for (int q = 0; q < NN; q++) for (int r = 0; r < NN; r++) { double a = 0.0; for (int b = nbLow; b < NN; b++) a += B.T.Phi[(size_t)q * NN + b] * B.T.PhiInv[(size_t)b * NN + r]; H[(size_t)q * NN + r] = a; }
      c->avH.upload(H, c->stream);
      GeomEval P1(B.type, 1);
      std::vector<double> wv, wd, tq((size_t)NN * NB), tf((size_t)B.T.NF * B.T.NQF * NB);
      auto coordOf = [&](int q, double* xi) { int s = NN; for (int d = 0; d < D; d++) { s /= N; xi[d] = B.T.x[(q / s) % N]; } };
      for (int q = 0; q < NN; q++) { double xi[3] = {0, 0, 0}; coordOf(q, xi); P1.weights(xi, wv, wd); for (int k = 0; k < NB; k++) tq[(size_t)q * NB + k] = wv[k]; }
      for (int f = 0; f < B.T.NF; f++) for (int j = 0; j < B.T.NQF; j++) {
        double xi[3] = {0, 0, 0}; coordOf(B.T.faceBase[f * B.T.NQF + j], xi);
        xi[B.T.faceDir[f]] = B.T.faceSide[f] ? 1.0 : -1.0;
        P1.weights(xi, wv, wd);
        for (int k = 0; k < NB; k++) tf[(size_t)(f * B.T.NQF + j) * NB + k] = wv[k];
      }
      c->avTabQ.upload(tq, c->stream); c->avTabF.upload(tf, c->stream);
      if ((int)c->avTagsHost.size() != B.n * NB || (int)c->avRadiusHost.size() != B.n) throw std::runtime_error("artificial viscosity needs sdg_set_element_nodes before sdg_finalize");
      std::vector<int> tg((size_t)B.n * NB); std::vector<double> rad(B.n);
      for (int e = 0; e < B.n; e++) { const int pos = B.perm[e]; rad[pos] = c->avRadiusHost[e]; for (int k = 0; k < NB; k++) tg[(size_t)pos * NB + k] = c->avTagsHost[(size_t)e * NB + k]; }
      c->avTags.upload(tg, c->stream); c->avRadius.upload(rad, c->stream);
      c->avElem.alloc((size_t)B.n * NB); c->avElem.zero(c->stream);
      c->avNode.alloc((size_t)std::max(c->avNodes, 1)); c->avNode.zero(c->stream);
      c->avE.alloc((size_t)B.n); c->avE.zero(c->stream);
    }
    if (c->lineTrace) {
      const LinePlan& LP = c->linePlan;
      c->links.upload(LP.links, c->stream); c->bndRec.upload(LP.bndRec, c->stream);
      if (B.affine) c->lfGeo.upload(LP.lfGeo, c->stream);
      std::vector<LineTabDev> lt(1);
      std::memcpy(lt[0].partner, LP.partner.data(), sizeof(lt[0].partner)); std::memcpy(lt[0].jLeft, LP.jLeft.data(), sizeof(lt[0].jLeft));
      c->ltab.upload(lt, c->stream);
      c->cLift = LP.cLift;
      if (c->traceTU) {
        const size_t nt = (size_t)B.n * 6 * kRow;
        for (int i = 0; i < 3; i++) { c->TU[i].alloc(nt); c->TU[i].zero(c->stream); }
        if (c->phys.ns) { c->TV.alloc(nt); c->TV.zero(c->stream); }
        c->TUb.alloc((size_t)std::max(F.nBnd, 1) * kRow); c->TUb.zero(c->stream);
      }
    }
    CUDA_OK(cudaStreamSynchronize(c->stream));
  }
  c->finalized = true;
  SDG_CATCH
}

int sdg_sizes(sdg_ctx* c, int32_t type, int32_t* out) {
  SDG_TRY
  if (c->mx) { c->mx->sizes(type, out); return 0; }
  needType(c, type);
  const BlockPlan& B = c->plan.blk;
  out[0] = B.n; out[1] = B.T.NN; out[2] = B.T.NN; out[3] = B.T.NF; out[4] = B.T.NF * B.T.NQF; out[5] = B.nn; out[6] = B.T.NQF; out[7] = c->NV;
  SDG_CATCH
}

int sdg_get_quadrature_coordinates(sdg_ctx* c, int32_t type, double* xq) {
  SDG_TRY
  if (c->mx) { c->mx->quadratureCoordinates(type, xq); return 0; }
  needType(c, type);
  c->plan.quadratureCoordinates(xq);
  SDG_CATCH
}

int sdg_get_boundary_quadrature_coordinates(sdg_ctx* c, double* xb) {
  SDG_TRY
  if (c->mx) { if (!c->haveFaces) throw std::runtime_error("elements and faces must be set first"); if (!c->finalized) c->mx->setFaces(c->plan.F); c->mx->boundaryQuadratureCoordinates(xb); return 0; }
  if (!c->haveBlock || !c->haveFaces) throw std::runtime_error("elements and faces must be set first");
  MeshPlan tmp = c->plan;  // coordinates only; leaves the finalized plan untouched
  tmp.buildFaces(xb, true);
  SDG_CATCH
}

int sdg_set_state_from_primitive(sdg_ctx* c, int32_t type, const double* prim) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->setStateFromPrimitive(type, prim); return 0; }
  needFinal(c); needDevice(c); needType(c, type);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  const size_t nd = c->stateDoubles();
  c->scratch.alloc(nd);
  CUDA_OK(cudaMemcpyAsync(c->scratch.p, prim, nd * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  const int nSet = B.nOwned;   // ghosts are fed by the halo exchange only (see transformModal)
  const int blocks = (int)std::min<size_t>(((size_t)nSet * B.T.NN + 255) / 256, 148 * 16);
  if (c->D == 1) primitiveToStateKernel<1><<<blocks, 256, 0, c->stream>>>(c->scratch.p, c->U[c->cur].p, c->perm.p, nSet, B.T.NN, c->phys);
  else if (c->D == 2) primitiveToStateKernel<2><<<blocks, 256, 0, c->stream>>>(c->scratch.p, c->U[c->cur].p, c->perm.p, nSet, B.T.NN, c->phys);
  else primitiveToStateKernel<3><<<blocks, 256, 0, c->stream>>>(c->scratch.p, c->U[c->cur].p, c->perm.p, nSet, B.T.NN, c->phys);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(c->stream));
  c->scratch.release();
  c->latest = c->cur; c->traceValid[c->cur] = false; c->gradSrc = -1;
  SDG_CATCH
}

int sdg_set_boundary_primitive(sdg_ctx* c, const double* prim) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->setBoundaryPrimitive(prim); return 0; }
  needFinal(c); needDevice(c);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk; const int nb = c->plan.F.nBnd;
  if (nb == 0) return 0;
  DevBuf<double> tmp; tmp.alloc((size_t)nb * B.T.NQF * c->NV);
  CUDA_OK(cudaMemcpyAsync(tmp.p, prim, tmp.n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  const int blocks = (nb * B.T.NQF + 255) / 256;
  if (c->D == 1) boundaryPrimitiveKernel<1><<<blocks, 256, 0, c->stream>>>(tmp.p, c->dummy.p, nb, B.T.NQF, c->phys);
  else if (c->D == 2) boundaryPrimitiveKernel<2><<<blocks, 256, 0, c->stream>>>(tmp.p, c->dummy.p, nb, B.T.NQF, c->phys);
  else boundaryPrimitiveKernel<3><<<blocks, 256, 0, c->stream>>>(tmp.p, c->dummy.p, nb, B.T.NQF, c->phys);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaStreamSynchronize(c->stream));
  SDG_CATCH
}

// nElems < B.n: the state setters of a partitioned block touch the OWNED elements only — the trailing ghost range of U[cur]
// belongs to the peers' halo pushes, which may land before or after this rank's setter (no receiver-ready handshake)
enum { kToNodal = 0, kToModal = 1, kProject = 2 };   // which: Phi1 (dir 0), Phi1^-1 (dir 1), Phi1^T (dir 1)
// elements [e0, e0 + ne) of the caller's order; the caller-order side of the transform is addressed by e, the internal side through perm
static void transformModalRange(sdg_ctx* c, const double* in, double* out, int which, int e0, int ne, cudaStream_t st) {
  const BlockPlan& B = c->plan.blk;
  if (ne <= 0) return;
  const int per = c->NV * B.T.NN, perP = c->NV * (B.T.NN + 1);
  const int epb = std::max(1, (40 << 10) / (16 * perP));
  const size_t smem = sizeof(double) * 2 * (size_t)epb * perP;
  using Fn = void (*)(const double*, double*, const double*, const int*, const int*, int, int, int, int, int);
  const Fn fn = (B.T.D == 3 && B.T.N == 4) ? tensorTransformKernel<3, 4> : (B.T.D == 2 && B.T.N == 4) ? tensorTransformKernel<2, 4> : tensorTransformKernel<0, 0>;
  const int blocks = std::min((ne + epb - 1) / epb, 148 * 8);
  const bool toNodal = which == kToNodal;
  fn<<<blocks, 256, smem, st>>>(toNodal ? in + (size_t)e0 * per : in, toNodal ? out : out + (size_t)e0 * per, c->M1[which].p, c->lexOf.p, c->perm.p + e0, ne,
                                B.T.N, B.T.D, toNodal ? 0 : 1, epb);
  c->launches++;
  CUDA_OK(cudaGetLastError());
}
static void transformModal(sdg_ctx* c, const double* in, double* out, int which, int nElems = -1) {
  transformModalRange(c, in, out, which, 0, nElems < 0 ? c->plan.blk.n : nElems, c->stream);
}
// Host <-> device seam of sdg_set_state / sdg_get_state: the copy runs on its own stream in kSeamChunks pieces so that the transform of
// one piece overlaps the PCIe transfer of the next (the transform is ~6-13 ms of a ~100 ms copy at 128^3 P3 hexahedra)
constexpr int kSeamChunks = 8;
static void seamStreams(sdg_ctx* c) {
  if (c->copyStream) return;
  CUDA_OK(cudaStreamCreateWithFlags(&c->copyStream, cudaStreamNonBlocking));
  for (auto& e : c->seamEvent) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
}

int sdg_set_state_device(sdg_ctx* c, int32_t type, const void* U_device) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->setStateDevice(type, U_device); return 0; }
  needFinal(c); needDevice(c); needType(c, type);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  transformModal(c, (const double*)U_device, c->U[c->cur].p, kToNodal, c->plan.blk.nOwned);
  c->latest = c->cur; c->traceValid[c->cur] = false; c->gradSrc = -1;
  SDG_CATCH
}
int sdg_get_state_device(sdg_ctx* c, int32_t type, void* U_device) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->getStateDevice(type, U_device); return 0; }
  needFinal(c); needDevice(c); needType(c, type);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  transformModal(c, c->U[c->cur].p, (double*)U_device, kToModal);
  SDG_CATCH
}

int sdg_set_state(sdg_ctx* c, int32_t type, const double* U) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->setState(type, U); return 0; }
  needFinal(c); needDevice(c); needType(c, type);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  const size_t per = (size_t)c->NV * B.T.NN;
  const int s = scratchBuffer(c);  // scratch: a stage buffer that holds no live data between steps
  c->traceValid[s] = false;
  const int n = B.nOwned;          // ghosts are fed by the halo exchange only (see transformModal)
  seamStreams(c);
  CUDA_OK(cudaEventRecord(c->seamEvent[kSeamChunks], c->stream));              // earlier work on the scratch buffer
  CUDA_OK(cudaStreamWaitEvent(c->copyStream, c->seamEvent[kSeamChunks], 0));
  const int nc = n >= 8192 ? kSeamChunks : 1;
  for (int k = 0; k < nc; k++) {
    const int e0 = (int)((long long)n * k / nc), e1 = (int)((long long)n * (k + 1) / nc);
    CUDA_OK(cudaMemcpyAsync(c->U[s].p + e0 * per, U + e0 * per, (size_t)(e1 - e0) * per * sizeof(double), cudaMemcpyHostToDevice, c->copyStream));
    CUDA_OK(cudaEventRecord(c->seamEvent[k], c->copyStream));
    CUDA_OK(cudaStreamWaitEvent(c->stream, c->seamEvent[k], 0));
    transformModalRange(c, c->U[s].p, c->U[c->cur].p, kToNodal, e0, e1 - e0, c->stream);
  }
  CUDA_OK(cudaStreamSynchronize(c->stream));
  c->latest = c->cur; c->traceValid[c->cur] = false; c->gradSrc = -1;
  SDG_CATCH
}
int sdg_get_state(sdg_ctx* c, int32_t type, double* U) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->getState(type, U); return 0; }
  needFinal(c); needDevice(c); needType(c, type);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  const size_t per = (size_t)c->NV * B.T.NN;
  const int s = scratchBuffer(c);
  c->traceValid[s] = false;
  const int n = B.n;
  seamStreams(c);
  const int nc = n >= 8192 ? kSeamChunks : 1;
  for (int k = 0; k < nc; k++) {
    const int e0 = (int)((long long)n * k / nc), e1 = (int)((long long)n * (k + 1) / nc);
    transformModalRange(c, c->U[c->cur].p, c->U[s].p, kToModal, e0, e1 - e0, c->stream);
    CUDA_OK(cudaEventRecord(c->seamEvent[k], c->stream));
    CUDA_OK(cudaStreamWaitEvent(c->copyStream, c->seamEvent[k], 0));
    CUDA_OK(cudaMemcpyAsync(U + e0 * per, c->U[s].p + e0 * per, (size_t)(e1 - e0) * per * sizeof(double), cudaMemcpyDeviceToHost, c->copyStream));
  }
  CUDA_OK(cudaStreamSynchronize(c->copyStream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  SDG_CATCH
}

int sdg_get_state_at_quadrature(sdg_ctx* c, int32_t type, double* Uq) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->stateAtQuadrature(type, Uq); return 0; }
  needFinal(c); needDevice(c); needType(c, type);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  const size_t nd = c->stateDoubles();
  const int s = scratchBuffer(c);
  c->traceValid[s] = false;
  seamTransposeKernel<<<148 * 8, 256, 0, c->stream>>>(c->U[c->cur].p, c->U[s].p, c->perm.p, B.n, c->NV, B.T.NN, 1);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(Uq, c->U[s].p, nd * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  SDG_CATCH
}

// Gradient at the nodes, on the device, of the state the reference's gradient coefficients belong to -- the input of the last RK stage
// after a step (gradSrc), the current state otherwise: all lifts (faceSel < 0: variable_gradient_basis_function_coefficient_)
// or the volume part plus the BR2 lift of local face faceSel only (variable_volume_gradient_ + variable_interface_gradient_(f),
// RawBinary.cpp:118-135).  Returns the device array [pos][NV*D][NN]; *zslow = node order of the line kernels.  c->G2 must be allocated.
static const double* nodalGradient(sdg_ctx* c, int faceSel, int* zslow) {
  StageArgs args; fillArgs(c, args);
  const int src = c->gradSrc >= 0 ? c->gradSrc : c->cur, spare = scratchBuffer(c);
  args.Uin = c->U[src].p; args.Ulast = c->U[src].p; args.Uout = c->U[spare].p;
  args.Gvol = c->G.p; args.Gout = c->G.p; args.faceSel = c->phys.visc == kBR2 ? faceSel : -1;
  if (c->lineTrace) {   // pass G of the line kernels leaves the TOTAL gradient in G (zeta-slowest node order)
    ensureTraces(c, src, c->stream);
    args.TUin = c->TU[src].p; args.TUout = c->TU[spare].p;
    lineBoundary(c, args, -1, c->stream);
  }
  runStage(c, args, -1, c->stream, 0);
  const bool second = !c->lineTrace && c->phys.visc == kBR2;
  if (second) { args.Gout = c->G2.p; args.mode = 3; runStage(c, args, -1, c->stream, 1); }
  *zslow = c->lineTrace ? 1 : 0;
  return second ? c->G2.p : c->G.p;
}

int sdg_get_gradient_at_quadrature(sdg_ctx* c, int32_t type, double* Gq) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->gradientAtQuadrature(type, Gq); return 0; }
  needFinal(c); needDevice(c); needType(c, type);
  if (!c->phys.ns) throw std::runtime_error("gradient state exists for Navier-Stokes models only");
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  const int NG = c->NV * c->D;
  c->G2.alloc(c->G.n);
  DevBuf<double> tmp; tmp.alloc(c->G.n);
  int zslow = 0;
  const double* g = nodalGradient(c, -1, &zslow);
  seamTransposeKernel<<<148 * 8, 256, 0, c->stream>>>(g, tmp.p, c->perm.p, B.n, NG, B.T.NN, zslow ? 3 : 1);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(Gq, tmp.p, c->G.n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  c->G2.release();
  SDG_CATCH
}

int sdg_get_gradient_state(sdg_ctx* c, int32_t type, double* G) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->gradientState(type, G); return 0; }
  needFinal(c); needDevice(c); needType(c, type);
  if (!c->phys.ns) throw std::runtime_error("gradient state exists for Navier-Stokes models only");
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  const int NG = c->NV * c->D;
  c->G2.alloc(c->G.n);
  DevBuf<double> tmp; tmp.alloc(c->G.n);
  int zslow = 0;
  const double* g = nodalGradient(c, -1, &zslow);
  modalRowsKernel<<<B.n, 128, sizeof(double) * NG * B.T.NN, c->stream>>>(g, tmp.p, c->PhiInv.p, c->perm.p, nullptr, B.n, NG, B.T.NN, zslow);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(cudaMemcpyAsync(G, tmp.p, c->G.n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  c->G2.release();
  SDG_CATCH
}

int sdg_get_boundary_gradient_state(sdg_ctx* c, double* Gb) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->boundaryGradientState(Gb); return 0; }
  needFinal(c); needDevice(c);
  if (!c->phys.ns) throw std::runtime_error("gradient state exists for Navier-Stokes models only");
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk; const FaceInput& F = c->plan.F;
  if (F.nBnd == 0) return 0;
  const int NG = c->NV * c->D, NF = 2 * c->D;
  const size_t row = (size_t)NG * B.T.NN;
  c->G2.alloc(c->G.n);
  DevBuf<double> out; out.alloc((size_t)F.nBnd * row);
  DevBuf<int> list; list.alloc((size_t)F.nBnd * 2);
  const bool perFace = c->phys.visc == kBR2;
  for (int f = perFace ? 0 : -1; f < (perFace ? NF : 0); f++) {
    std::vector<int> h;
    for (int b = 0; b < F.nBnd; b++) if (!perFace || F.lf[F.nInt + b] == f) { h.push_back(B.perm[F.le[F.nInt + b]]); h.push_back(b); }
    if (h.empty()) continue;
    int zslow = 0;
    const double* g = nodalGradient(c, f, &zslow);
    CUDA_OK(cudaMemcpyAsync(list.p, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    const int nl = (int)(h.size() / 2);
    modalRowsKernel<<<nl, 128, sizeof(double) * row, c->stream>>>(g, out.p, c->PhiInv.p, nullptr, reinterpret_cast<const int2*>(list.p), nl, NG, B.T.NN, zslow);
    c->launches++;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(c->stream));   // h is reused
  }
  CUDA_OK(cudaMemcpyAsync(Gb, out.p, out.n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  c->G2.release();
  SDG_CATCH
}

// System::setArtificialViscosity (SystemControl.cpp:105-108) for ShockCapturingEnum::ArtificialViscosity; before sdg_finalize
int sdg_set_artificial_viscosity(sdg_ctx* c, double empirical_tolerance, double artificial_viscosity_factor, int32_t node_number) {
  SDG_TRY
  if (c->finalized) throw std::runtime_error("sdg_set_artificial_viscosity must precede sdg_finalize");
  if (c->phys.ns) throw std::runtime_error("artificial viscosity is built for the Euler models (every example of the reference that uses it)");
  if (node_number < 1) throw std::runtime_error("node_number must be positive");
  c->phys.av = 1; c->avTol = empirical_tolerance; c->avFactor = artificial_viscosity_factor; c->avNodes = node_number;
  if (c->mx) c->mx->setArtificialViscosity(empirical_tolerance, artificial_viscosity_factor, node_number);
  SDG_CATCH
}
int sdg_set_element_nodes(sdg_ctx* c, int32_t type, const int32_t* node_tag, const double* inner_radius) {
  SDG_TRY
  if (c->finalized) throw std::runtime_error("sdg_set_element_nodes must precede sdg_finalize");
  if (!c->phys.av) throw std::runtime_error("sdg_set_artificial_viscosity first");
  if (c->mx) { c->mx->setElementNodes(type, node_tag, inner_radius); return 0; }
  needType(c, type);
  const BlockPlan& B = c->plan.blk;
  const int NB = 1 << c->D;
  c->avTagsHost.assign(node_tag, node_tag + (size_t)B.n * NB);
  c->avRadiusHost.assign(inner_radius, inner_radius + B.n);
  for (int t : c->avTagsHost) if (t < 0 || t >= c->avNodes) throw std::runtime_error("node tag out of range");
  SDG_CATCH
}
int sdg_get_node_artificial_viscosity(sdg_ctx* c, double* out) {
  SDG_TRY
  needFinal(c);
  if (!c->phys.av) throw std::runtime_error("artificial viscosity is not enabled");
  if (c->mx) { c->mx->nodeArtificialViscosity(out); return 0; }
  needDevice(c);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  CUDA_OK(cudaMemcpyAsync(out, c->avNode.p, (size_t)c->avNodes * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  SDG_CATCH
}
int sdg_get_element_artificial_viscosity(sdg_ctx* c, int32_t type, double* out) {
  SDG_TRY
  needFinal(c);
  if (!c->phys.av) throw std::runtime_error("artificial viscosity is not enabled");
  if (c->mx) { c->mx->elementArtificialViscosity(type, out); return 0; }
  needDevice(c); needType(c, type);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk; const int NB = 1 << c->D;
  std::vector<double> h((size_t)B.n * NB);
  CUDA_OK(cudaMemcpyAsync(h.data(), c->avElem.p, h.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  for (int e = 0; e < B.n; e++) for (int k = 0; k < NB; k++) out[(size_t)e * NB + k] = h[(size_t)B.perm[e] * NB + k];
  SDG_CATCH
}
int sdg_update_artificial_viscosity(sdg_ctx* c) {
  SDG_TRY
  needFinal(c);
  if (!c->phys.av) return 0;
  if (c->mx) { c->mx->updateArtificialViscosity(); return 0; }
  needDevice(c);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  avUpdate(c, c->cur, c->stream);
  CUDA_OK(cudaStreamSynchronize(c->stream));
  SDG_CATCH
}

// Partitioned runs: the node maximum of Solver::calculateArtificialViscosity (the cwiseMax combine of SpatialDiscrete.cpp:89-108) spans the
// ranks.  sdg_step_begin leaves each rank's maximum over its OWNED elements in the node array; the caller max-reduces that array over
// the ranks in place (it lives on the device: NCCL all-reduce) and sdg_av_store rewrites the corner values of owned and ghost elements.
int sdg_av_node_buffer(sdg_ctx* c, void** device_nodes, int64_t* count) {
  SDG_TRY
  if (c->mx) throw std::runtime_error("not available on the dense-operator (triangle / mixed-type) path: single GPU, sdg_step only");
  needFinal(c); needDevice(c);
  if (!c->phys.av) throw std::runtime_error("artificial viscosity is not enabled");
  *device_nodes = c->avNode.p; *count = (int64_t)c->avNode.n;
  SDG_CATCH
}
int sdg_av_store(sdg_ctx* c, void* stream) {
  SDG_TRY
  if (c->mx) throw std::runtime_error("not available on the dense-operator (triangle / mixed-type) path: single GPU, sdg_step only");
  needFinal(c); needDevice(c);
  if (!c->phys.av) return 0;
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  const int NB = 1 << c->D;
  avStoreKernel<<<std::min((B.n * NB + 255) / 256, 148 * 8), 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(c->avNode.p, c->avTags.p, B.n, NB, c->avElem.p);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  SDG_CATCH
}

// ViewVariable::get (VariableConvertor.cpp:754-872) at the volume quadrature points of the resident state, [n][Nq]
int sdg_get_view_variable(sdg_ctx* c, int32_t type, int32_t variable, double* out) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->viewVariable(type, variable, out); return 0; }
  needFinal(c); needDevice(c); needType(c, type);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  const size_t npts = (size_t)B.n * B.T.NN;
  const int NG = c->NV * c->D;
  DevBuf<double> cons, grad, eps, res;
  cons.alloc(npts * c->NV); res.alloc(npts);
  seamTransposeKernel<<<148 * 8, 256, 0, c->stream>>>(c->U[c->cur].p, cons.p, c->perm.p, B.n, c->NV, B.T.NN, 1);
  c->launches++;
  if (c->phys.ns) {
    c->G2.alloc(c->G.n); grad.alloc(npts * NG);
    int zslow = 0;
    const double* g = nodalGradient(c, -1, &zslow);
    seamTransposeKernel<<<148 * 8, 256, 0, c->stream>>>(g, grad.p, c->perm.p, B.n, NG, B.T.NN, zslow ? 3 : 1);
    c->launches++;
  }
  if (c->phys.av) {   // artificial_viscosity_ at the points: nodal basis * corner values (RawBinary.cpp:206-211)
    eps.alloc(npts);
    avAtNodesKernel<<<148 * 4, 256, 0, c->stream>>>(c->avElem.p, c->avTabQ.p, c->perm.p, B.n, B.T.NN, 1 << c->D, eps.p);
    c->launches++;
  }
  launchViewVariable(c->D, c->phys, variable, npts, cons.p, c->phys.ns ? grad.p : nullptr, c->phys.av ? eps.p : nullptr, res.p, c->stream);
  c->launches++;
  CUDA_OK(cudaMemcpyAsync(out, res.p, npts * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  c->G2.release();
  SDG_CATCH
}

int sdg_compute_dt(sdg_ctx* c, double cfl, double* dt) {
  SDG_TRY
  if (c->mx) { needFinal(c); *dt = c->mx->computeDt(cfl); return 0; }
  needFinal(c); needDevice(c);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  const int blocks = (int)std::min<size_t>(((size_t)B.nOwned * B.T.NN + 255) / 256, 1024);
  if (c->D == 1) deltaTimeKernel<1><<<blocks, 256, 0, c->stream>>>(c->U[c->cur].p, c->minEdge.p, B.nOwned, B.T.NN, c->cfg.p, cfl, c->phys, c->dtPartial.p);
  else if (c->D == 2) deltaTimeKernel<2><<<blocks, 256, 0, c->stream>>>(c->U[c->cur].p, c->minEdge.p, B.nOwned, B.T.NN, c->cfg.p, cfl, c->phys, c->dtPartial.p);
  else deltaTimeKernel<3><<<blocks, 256, 0, c->stream>>>(c->U[c->cur].p, c->minEdge.p, B.nOwned, B.T.NN, c->cfg.p, cfl, c->phys, c->dtPartial.p);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  std::vector<double> h(blocks);
  CUDA_OK(cudaMemcpyAsync(h.data(), c->dtPartial.p, sizeof(double) * blocks, cudaMemcpyDeviceToHost, c->stream));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  double best = 1.7976931348623157e308;
  for (double v : h) best = std::min(best, v);
  *dt = best;
  SDG_CATCH
}

int sdg_num_stages(sdg_ctx* c) { return c->nStages; }
int sdg_num_passes(sdg_ctx* c) { return twoPass(c) ? 2 : 1; }
void* sdg_stream(sdg_ctx* c) { return (void*)c->stream; }
int64_t sdg_launch_count(sdg_ctx* c) { return c->launches + (c->mx ? c->mx->launches : 0); }

int sdg_synchronize(sdg_ctx* c) {
  SDG_TRY
  needDevice(c);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  CUDA_OK(cudaStreamSynchronize(c->stream));
  SDG_CATCH
}

int sdg_step_begin(sdg_ctx* c, double dt) {
  SDG_TRY
  if (c->mx) throw std::runtime_error("not available on the dense-operator (triangle / mixed-type) path: single GPU, sdg_step only");
  needFinal(c); needDevice(c);
  c->stepDt = dt;
  if (c->phys.av) { CUDA_OK(cudaSetDevice(c->cfg.device)); avUpdate(c, c->cur, c->stream); }
  SDG_CATCH
}

int sdg_stage_pass(sdg_ctx* c, int32_t stage, int32_t pass, int32_t part, void* stream) {
  SDG_TRY
  if (c->mx) throw std::runtime_error("not available on the dense-operator (triangle / mixed-type) path: single GPU, sdg_step only");
  needFinal(c); needDevice(c);
  if (stage < 0 || stage >= c->nStages || pass < 0 || pass >= sdg_num_passes(c) || part < -1 || part > 1) throw std::runtime_error("bad stage/pass/part");
  CUDA_OK(cudaSetDevice(c->cfg.device));
  stageLaunch(c, stage, part, stream ? (cudaStream_t)stream : c->stream, twoPass(c) ? pass : 1);
  SDG_CATCH
}

int sdg_step_end(sdg_ctx* c, double* sums) {
  SDG_TRY
  if (c->mx) throw std::runtime_error("not available on the dense-operator (triangle / mixed-type) path: single GPU, sdg_step only");
  needFinal(c); needDevice(c);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  finishStep(c);
  if (sums) reduceNorm(c, sums);
  // a wait on the peers' pushes that timed out has trapped (haloWaitKernel): the sticky error surfaces here on every step,
  // with or without `sums`, without adding a synchronisation to the healthy path
  if (c->haloErr.p) {
    const cudaError_t q = cudaStreamQuery(c->stream);
    if (q != cudaSuccess && q != cudaErrorNotReady) throw std::runtime_error(std::string("peer-memory halo exchange: a rank waited in vain for a peer's push (") + cudaGetErrorString(q) + ")");
  }
  SDG_CATCH
}

namespace {
constexpr int kGraphMaxChunks = 1 << 15;   // above this a stage kernel runs for >= 100 us and launch overhead is irrelevant

void launchOneStep(sdg_ctx* c) {
  if (c->phys.av) avUpdate(c, c->cur, c->stream);
  for (int s = 0; s < c->nStages; s++) stageLaunch(c, s, -1, c->stream);
  finishStep(c);
}

// n_steps time steps on the context's stream
void runSteps(sdg_ctx* c, double dt, int n_steps) {
  c->stepDt = dt;
  ensureTraces(c, c->cur, c->stream);   // outside any graph capture
  const bool useGraph = c->plan.blk.nChunks <= kGraphMaxChunks && c->nStages > 1 && n_steps >= 4 && !getenv("SDG_NO_GRAPH");
  int it = 0;
  if (useGraph) {
    if (!c->graphWarm) { launchOneStep(c); it++; c->graphWarm = true; }   // first launches set the kernels' function attributes
    if (c->stepGraph == nullptr || c->graphDt != dt || c->graphCur != c->cur) {
      if (c->stepGraph) { cudaGraphExecDestroy(c->stepGraph); c->stepGraph = nullptr; }
      cudaGraph_t g = nullptr;
      const int64_t l0 = c->launches;
      CUDA_OK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      launchOneStep(c);                                   // nStages > 1: the buffer rotation returns to `cur`, every step is identical
      CUDA_OK(cudaStreamEndCapture(c->stream, &g));
      c->graphLaunches = c->launches - l0;                // kernels per replay
      c->launches = l0;                                   // captured, not launched
      CUDA_OK(cudaGraphInstantiate(&c->stepGraph, g, 0));
      cudaGraphDestroy(g);
      c->graphDt = dt; c->graphCur = c->cur;
    }
    for (; it < n_steps; it++) { CUDA_OK(cudaGraphLaunch(c->stepGraph, c->stream)); c->launches += c->graphLaunches; }
    if (n_steps > 0) markLastStageInput(c);   // a replay runs no host code: a setter may have cleared the mark since the capture
    return;
  }
  for (; it < n_steps; it++) launchOneStep(c);
}
}  // namespace

int sdg_step(sdg_ctx* c, double dt, int32_t n_steps, double* relative_error) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->step(dt, n_steps, relative_error, nullptr); return 0; }
  needFinal(c); needDevice(c);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  runSteps(c, dt, n_steps);
  if (relative_error) {
    double sums[8];
    reduceNorm(c, sums);
    for (int v = 0; v < c->NV; v++) relative_error[v] = sums[v] / c->plan.blk.nOwned;  // TimeIntegration.cpp:323
  } else {
    CUDA_OK(cudaStreamSynchronize(c->stream));
  }
  SDG_CATCH
}

int sdg_step_timed(sdg_ctx* c, double dt, int32_t n_steps, double* relative_error, float* milliseconds) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->step(dt, n_steps, relative_error, milliseconds); return 0; }
  needFinal(c); needDevice(c);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0)); CUDA_OK(cudaEventCreate(&e1));
  c->stepDt = dt;
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaEventRecord(e0, c->stream));
  runSteps(c, dt, n_steps);
  CUDA_OK(cudaEventRecord(e1, c->stream));
  CUDA_OK(cudaEventSynchronize(e1));
  if (milliseconds) CUDA_OK(cudaEventElapsedTime(milliseconds, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (relative_error) {
    double sums[8];
    reduceNorm(c, sums);
    for (int v = 0; v < c->NV; v++) relative_error[v] = sums[v] / c->plan.blk.nOwned;
  }
  SDG_CATCH
}

namespace {

// sdg_step_host can stream when a launch over a chunk reads, besides the chunk's own elements, only what the face neighbours' PREVIOUS
// launch wrote (their states or gradients, or the trace rows they published): every single-block context of the tensor kernels on one
// GPU without shock capturing (its viscosity pass runs over the whole mesh).  The virtual neighbour traces of the boundary faces, which
// the trace-based Navier-Stokes gradient pass reads, are launched per level as well (face lists sorted by the level of the parent's chunk).
bool hostPipeEligible(const sdg_ctx* c) {
  return c->haveBlock && !c->phys.av && c->plan.blk.nGhost == 0 && c->plan.blk.nOwned >= 8192 && !getenv("SDG_NO_HOST_PIPE");
}

// Dependency levels of the streamed step.  The caller's element order is cut into G contiguous upload groups; a chunk's traces can be
// taken once its own elements have arrived (level t0), launch k of a chunk (the stages in order; Navier-Stokes: gradient pass, then
// residual pass of each stage) can run once launch k - 1 (launch 0: the traces) of the chunk and of every face neighbour's chunk has
// run: r_k = max over {chunk, neighbours} of r_(k-1).  An upload group goes back to the host
// at the level at which the last stage of all of its elements is done.  Launching level after level on ONE stream (traces, stage 1,
// stage 2, ... of that level, in this order) satisfies every dependency without an event between kernels.
void buildHostPipe(sdg_ctx* c) {
  auto& P = c->pipe;
  const BlockPlan& B = c->plan.blk;
  const int n = B.nOwned, K = B.K, nCh = B.nChunks, S = c->nStages * (twoPass(c) ? 2 : 1);   // launches per step after the traces
  int G = std::max(2, std::min(64, n / 16384));
  if (const char* e = getenv("SDG_HOST_PIPE_GROUPS")) G = std::max(1, std::min(256, atoi(e)));
  P.G = G;
  P.first.resize(G + 1);
  for (int k = 0; k <= G; k++) P.first[k] = (int)((long long)n * k / G);
  std::vector<int> groupOf(n);   // by internal position
  for (int k = 0; k < G; k++) for (int ci = P.first[k]; ci < P.first[k + 1]; ci++) groupOf[B.perm[ci]] = k;
  std::vector<std::vector<int>> lvl(S + 1, std::vector<int>(nCh, 0));
  for (int e = 0; e < n; e++) lvl[0][e / K] = std::max(lvl[0][e / K], groupOf[e]);
  // pairs of different chunks that share a face: link records of the line plan, else the chunks' face lists
  std::vector<std::pair<int, int>> adj;
  if (c->lineTrace) {
    const std::vector<int>& L = c->linePlan.links;
    for (int e = 0; e < n; e++)
      for (int f = 0; f < 6; f++) {
        const int o = L[((size_t)e * 6 + f) * 4];
        if (o >= 0 && o < n && o / K != e / K) adj.emplace_back(e / K, o / K);
      }
  } else {
    for (size_t k = 0; k + 3 < B.faceRec.size(); k += 4) {
      const int a = B.faceRec[k], b = B.faceRec[k + 1];
      if (a >= 0 && b >= 0 && a < n && b < n && a / K != b / K) { adj.emplace_back(a / K, b / K); adj.emplace_back(b / K, a / K); }
    }
  }
  for (int s = 1; s <= S; s++) {
    lvl[s] = lvl[s - 1];
    for (const auto& ab : adj) lvl[s][ab.first] = std::max(lvl[s][ab.first], lvl[s - 1][ab.second]);
  }
  // chunk lists sorted by (kind, level); ascending chunk index inside a list
  P.off.assign((size_t)(S + 1) * G + 1, 0);
  std::vector<int> lists((size_t)(S + 1) * nCh);
  for (int k = 0; k <= S; k++) {
    std::vector<int> count(G, 0);
    for (int ch = 0; ch < nCh; ch++) count[lvl[k][ch]]++;
    int run = k * nCh;
    for (int g = 0; g < G; g++) { P.off[(size_t)k * G + g] = run; run += count[g]; }
    std::vector<int> fill(G);
    for (int g = 0; g < G; g++) fill[g] = P.off[(size_t)k * G + g];
    for (int ch = 0; ch < nCh; ch++) lists[fill[lvl[k][ch]]++] = ch;
  }
  P.off[(size_t)(S + 1) * G] = (S + 1) * nCh;
  // trace-based Navier-Stokes: the boundary faces whose virtual neighbour traces the gradient pass of (stage, level) reads
  const bool bndLists = c->lineTrace && c->traceTU && twoPass(c) && c->plan.F.nBnd > 0;
  std::vector<int> bnd;
  if (bndLists) {
    const int nB = c->plan.F.nBnd, nSt = c->nStages;
    P.bndOff.assign((size_t)nSt * G + 1, 0);
    bnd.resize((size_t)nSt * nB);
    for (int st = 0; st < nSt; st++) {
      const std::vector<int>& lv = lvl[1 + 2 * st];   // kind of the gradient pass of stage st
      std::vector<int> count(G, 0);
      for (int fb = 0; fb < nB; fb++) count[lv[c->linePlan.bndRec[(size_t)fb * 4] / K]]++;
      int run = st * nB;
      std::vector<int> fill(G);
      for (int g = 0; g < G; g++) { P.bndOff[(size_t)st * G + g] = run; fill[g] = run; run += count[g]; }
      for (int fb = 0; fb < nB; fb++) bnd[fill[lv[c->linePlan.bndRec[(size_t)fb * 4] / K]]++] = fb;
    }
    P.bndOff[(size_t)nSt * G] = nSt * nB;
  }
  P.download.assign(G, {});
  int early = 0;
  for (int g = 0; g < G; g++) {
    int d = 0;
    for (int ci = P.first[g]; ci < P.first[g + 1]; ci++) d = std::max(d, lvl[S][B.perm[ci] / K]);
    P.download[d].push_back(g);
    if (d < G - 1) early++;
  }
  P.overlap = (double)early / G;
  if (!c->hasDevice) return;   // plan-only context: the levels are all there is to inspect
  P.lists.upload(lists);
  if (bndLists) P.bndLists.upload(bnd);
  const size_t per = (size_t)c->NV * B.T.NN;
  P.up.alloc((size_t)n * per); P.down.alloc((size_t)n * per);
  CUDA_OK(cudaStreamCreateWithFlags(&P.d2h, cudaStreamNonBlocking));
  P.upEv.resize(G); P.downEv.resize(G);
  P.timing = getenv("SDG_HOST_PIPE_TIMING") != nullptr;
  const unsigned flags = P.timing ? cudaEventDefault : cudaEventDisableTiming;
  for (auto& e : P.upEv) CUDA_OK(cudaEventCreateWithFlags(&e, flags));
  for (auto& e : P.downEv) CUDA_OK(cudaEventCreateWithFlags(&e, flags));
  if (P.timing) {
    P.doneEv.resize(G);
    for (auto& e : P.doneEv) CUDA_OK(cudaEventCreate(&e));
    CUDA_OK(cudaEventCreate(&P.t0));
  }
}

void pipeLaunchList(sdg_ctx* c, int kind, int level, const std::function<void()>& launch) {
  const auto& P = c->pipe;
  const int i = kind * P.G + level, a = P.off[i], b = (level == P.G - 1) ? (kind + 1) * c->plan.blk.nChunks : P.off[i + 1];
  if (b <= a) return;
  struct Guard { sdg_ctx* c; ~Guard() { c->listOverride = nullptr; c->listCount = 0; } } guard{c};   // also when a launch throws
  c->listOverride = P.lists.p + a; c->listCount = b - a;
  launch();
}

}  // namespace

int sdg_step_host(sdg_ctx* c, int32_t type, double dt, const double* U_in, double* U_out, double* relative_error) {
  SDG_TRY
  needFinal(c);
  if (!U_in || !U_out) throw std::runtime_error("null argument");
  bool stream = !c->mx && c->hasDevice && hostPipeEligible(c) && c->pipe.G >= 0;
  if (stream && c->pipe.G == 0) {
    CUDA_OK(cudaSetDevice(c->cfg.device));
    try { buildHostPipe(c); }
    catch (const std::exception&) {   // no room for the two staging arrays: the phases reuse a stage buffer instead
      cudaGetLastError();
      c->pipe.lists.release(); c->pipe.up.release(); c->pipe.down.release();
      c->pipe.G = -1; stream = false;
    }
  }
  if (!stream) {   // same result, one phase after the other
    if (sdg_set_state(c, type, U_in)) throw std::runtime_error(g_err);
    if (sdg_step(c, dt, 1, relative_error)) throw std::runtime_error(g_err);
    if (sdg_get_state(c, type, U_out)) throw std::runtime_error(g_err);
    return 0;
  }
  needDevice(c); needType(c, type);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  auto& P = c->pipe;
  seamStreams(c);
  const BlockPlan& B = c->plan.blk;
  const size_t per = (size_t)c->NV * B.T.NN;
  const int G = P.G, S = c->nStages;
  c->stepDt = dt;
  int inLast, outLast; stageBuffers(c, S - 1, inLast, outLast);
  // whatever goes wrong below, no copy may still be reading or writing the caller's buffers when the call returns
  struct Drain { sdg_ctx* c; bool armed = true; ~Drain() { if (armed) { cudaStreamSynchronize(c->copyStream); cudaStreamSynchronize(c->stream); cudaStreamSynchronize(c->pipe.d2h); } } } drain{c};
  // uploads: nothing on the device waits for them except the transform of the same group
  CUDA_OK(cudaEventRecord(c->seamEvent[kSeamChunks], c->stream));   // an earlier call's transforms out of the staging array
  CUDA_OK(cudaStreamWaitEvent(c->copyStream, c->seamEvent[kSeamChunks], 0));
  if (P.timing) CUDA_OK(cudaEventRecord(P.t0, c->copyStream));
  for (int g = 0; g < G; g++) {
    const size_t e0 = P.first[g], ne = P.first[g + 1] - P.first[g];
    CUDA_OK(cudaMemcpyAsync(P.up.p + e0 * per, U_in + e0 * per, ne * per * sizeof(double), cudaMemcpyHostToDevice, c->copyStream));
    CUDA_OK(cudaEventRecord(P.upEv[g], c->copyStream));
  }
  c->latest = c->cur; c->gradSrc = -1;
  for (int b = 0; b < 3; b++) c->traceValid[b] = false;
  for (int g = 0; g < G; g++) {
    CUDA_OK(cudaStreamWaitEvent(c->stream, P.upEv[g], 0));
    transformModalRange(c, P.up.p, c->U[c->cur].p, kToNodal, P.first[g], P.first[g + 1] - P.first[g], c->stream);
    if (c->traceTU) pipeLaunchList(c, 0, g, [&] {
      StageArgs a; fillArgs(c, a);
      a.Uin = c->U[c->cur].p; a.TUout = c->TU[c->cur].p; a.chunkList = c->listOverride;
      c->lineFns.trace(a, c->listCount, c->stream);
      c->launches++;
      CUDA_OK(cudaGetLastError());
    });
    if (c->traceTU) c->traceValid[c->cur] = true;   // level by level: every row a launch below reads has been written by a launch above
    const int passes = twoPass(c) ? 2 : 1;
    for (int s = 0; s < S; s++)
      for (int q = 0; q < passes; q++) {
        if (q == 0 && !P.bndOff.empty()) {   // virtual neighbour traces of this level's boundary faces, then no launch over all of them
          const int b0 = P.bndOff[(size_t)s * G + g], b1 = P.bndOff[(size_t)s * G + g + 1];
          int in, out; stageBuffers(c, s, in, out);
          if (b1 > b0) {
            StageArgs a; fillArgs(c, a);
            a.Uin = c->U[in].p; a.TUin = c->TU[in].p; a.chunkList = P.bndLists.p + b0;
            c->lineFns.boundary(a, reinterpret_cast<const int4*>(c->bndRec.p), b1 - b0, c->stream);
            c->launches++;
            CUDA_OK(cudaGetLastError());
          }
          c->bndKey = c->stepCount * 4 + s;
        }
        pipeLaunchList(c, 1 + s * passes + q, g, [&] { stageLaunch(c, s, -1, c->stream, passes == 2 ? q : -1); });
      }
    for (int h : P.download[g]) {
      const size_t e0 = P.first[h], ne = P.first[h + 1] - P.first[h];
      transformModalRange(c, c->U[outLast].p, P.down.p, kToModal, (int)e0, (int)ne, c->stream);
      CUDA_OK(cudaEventRecord(P.downEv[h], c->stream));
      CUDA_OK(cudaStreamWaitEvent(P.d2h, P.downEv[h], 0));
      CUDA_OK(cudaMemcpyAsync(U_out + e0 * per, P.down.p + e0 * per, ne * per * sizeof(double), cudaMemcpyDeviceToHost, P.d2h));
      if (P.timing) CUDA_OK(cudaEventRecord(P.doneEv[h], P.d2h));
    }
  }
  finishStep(c);
  if (relative_error) {
    double sums[8];
    reduceNorm(c, sums);
    for (int v = 0; v < c->NV; v++) relative_error[v] = sums[v] / B.nOwned;
  }
  CUDA_OK(cudaStreamSynchronize(c->stream));
  CUDA_OK(cudaStreamSynchronize(P.d2h));
  drain.armed = false;
  if (P.timing) {   // time line of the copies, milliseconds after the first upload was queued
    std::fprintf(stderr, "sdg_step_host time line (ms): group, upload done, download queued (last stage + transform done), download done\n");
    for (int g = 0; g < G; g++) {
      float a = 0, b = 0, d = 0;
      cudaEventElapsedTime(&a, P.t0, P.upEv[g]); cudaEventElapsedTime(&b, P.t0, P.downEv[g]); cudaEventElapsedTime(&d, P.t0, P.doneEv[g]);
      std::fprintf(stderr, "%3d %8.2f %8.2f %8.2f\n", g, a, b, d);
    }
  }
  SDG_CATCH
}

/* diagnostics of the streamed step: number of upload groups and the fraction of them that go back to the host before the last level */
int sdg_step_host_info(sdg_ctx* c, int32_t* groups, double* early_fraction) {
  SDG_TRY
  needFinal(c);
  if (!c->mx && hostPipeEligible(c) && c->pipe.G == 0) { if (c->hasDevice) CUDA_OK(cudaSetDevice(c->cfg.device)); buildHostPipe(c); }
  if (groups) *groups = std::max(0, c->pipe.G);
  if (early_fraction) *early_fraction = c->pipe.overlap;
  SDG_CATCH
}

int sdg_residual(sdg_ctx* c, int32_t type, double* Rmodal, double* rhsq) {
  SDG_TRY
  if (c->mx) { needFinal(c); c->mx->residual(type, Rmodal, rhsq); return 0; }
  needFinal(c); needDevice(c); needType(c, type);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  const size_t nd = c->stateDoubles();
  const int a = (c->cur + 1) % 3, b = (c->cur + 2) % 3;
  c->traceValid[a] = c->traceValid[b] = false;   // both scratch buffers are overwritten below
  c->gradSrc = -1;
  if (c->phys.av) avUpdate(c, c->cur, c->stream);   // parity hook: the viscosity of the CURRENT state
  for (int mode = 1; mode <= 2; mode++) {
    double* host = mode == 1 ? rhsq : Rmodal;
    if (!host) continue;
    StageArgs args; fillArgs(c, args);
    args.Uin = c->U[c->cur].p; args.Ulast = c->U[c->cur].p; args.Uout = c->U[a].p; args.mode = mode;
    args.aLast = 0.0; args.aCur = 0.0; args.bdt = 1.0;
    CUDA_OK(cudaMemsetAsync(c->U[a].p, 0, nd * sizeof(double), c->stream));
    args.Gvol = c->G.p; args.Gout = c->G.p;
    if (c->lineTrace) {
      ensureTraces(c, c->cur, c->stream);
      args.TUin = c->TU[c->cur].p; args.TUout = c->TU[a].p;
      if (twoPass(c)) lineBoundary(c, args, -1, c->stream);
    }
    if (twoPass(c)) runStage(c, args, -1, c->stream, 0);
    runStage(c, args, -1, c->stream, 1);
    if (mode == 1) {
      seamTransposeKernel<<<148 * 8, 256, 0, c->stream>>>(c->U[a].p, c->U[b].p, c->perm.p, B.n, c->NV, B.T.NN, 1);
      c->launches++;
    } else {
      transformModal(c, c->U[a].p, c->U[b].p, kProject);
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(host, c->U[b].p, nd * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CUDA_OK(cudaStreamSynchronize(c->stream));
  }
  SDG_CATCH
}

int sdg_debug_plan(sdg_ctx* c, int32_t what, double* out_d, int32_t* out_i, int64_t* count) {
  SDG_TRY
  if (c->mx) {
    needFinal(c);
    const std::vector<double>& d = c->mx->debugArray(what);
    if (count) *count = (int64_t)d.size();
    if (out_d) std::memcpy(out_d, d.data(), d.size() * sizeof(double));
    return 0;
  }
  needFinal(c);
  const BlockPlan& B = c->plan.blk;
  const std::vector<double>* d = nullptr; const std::vector<int>* i = nullptr;
  std::vector<int> misc = {B.affine ? 1 : 0, B.K, B.nChunks, B.nOwned};
  std::vector<int> seqs, midx, lpart, ljl;
  for (unsigned char x : c->linePlan.partner) lpart.push_back(x);
  for (unsigned char x : c->linePlan.jLeft) ljl.push_back(x);
  {  // right-side face-point permutations for rotations 0..3 (the table the kernels stage), modal function index triples
    const int ft = faceType(B.type);
    for (int r = 0; r < 4; r++) { std::vector<int> q = faceSequence(ft, B.T.N, ft == kLine ? 0 : r); seqs.insert(seqs.end(), q.begin(), q.end()); }
    for (auto& t : B.T.modalIdx) for (int k = 0; k < 3; k++) midx.push_back(t[k]);
  }
  switch (what) {
    case 0: d = &B.geoE; break; case 1: d = &B.invjw; break; case 2: d = &B.minEdge; break; case 3: d = &c->plan.geoF; break;
    case 10: i = &B.perm; break; case 11: i = &B.chunkFaceOff; break; case 12: i = &B.faceRec; break;
    case 13: i = &B.chunkInterior; break; case 14: i = &B.chunkBoundary; break; case 15: i = &misc; break;
    case 16: i = &seqs; break; case 17: i = &B.T.faceBase; break; case 18: i = &B.T.nodeFacePt; break; case 19: i = &midx; break;
    case 20: i = &c->linePlan.links; break; case 21: i = &lpart; break; case 22: i = &ljl; break; case 23: i = &c->linePlan.bndRec; break;
    case 4: d = &B.T.Phi; break; case 5: d = &B.T.Dm; break; case 6: d = &B.T.Lend; break; case 7: d = &B.T.x; break; case 8: d = &B.T.w; break;
    default: throw std::runtime_error("bad diagnostics id");
  }
  if (d) { if (count) *count = (int64_t)d->size(); if (out_d) std::memcpy(out_d, d->data(), d->size() * sizeof(double)); }
  if (i) { if (count) *count = (int64_t)i->size(); if (out_i) std::memcpy(out_i, i->data(), i->size() * sizeof(int)); }
  SDG_CATCH
}

// what travels between ranks: whole elements of the state / volume gradient, or — trace-based line kernels — the elements' face-trace rows
// TU (what = 0) and TV (what = 1)
static int haloStride(sdg_ctx* c, int what) { return c->rowHalo ? kRow : c->traceTU ? 6 * kRow : (int)c->elemDoubles() * (what == 1 ? c->D : 1); }
static double* haloField(sdg_ctx* c, int what) { return c->traceTU ? (what == 1 ? c->TV.p : c->TU[c->latest].p) : (what == 1 ? c->G.p : c->U[c->latest].p); }

int sdg_halo_doubles_per_element(sdg_ctx* c, int32_t what) { return haloStride(c, what); }

int sdg_set_halo_send(sdg_ctx* c, int32_t type, int32_t n_send, const int32_t* elems) {
  SDG_TRY
  if (c->mx) throw std::runtime_error("not available on the dense-operator (triangle / mixed-type) path: single GPU, sdg_step only");
  needFinal(c); needDevice(c); needType(c, type);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  std::vector<int> pos(n_send);
  for (int i = 0; i < n_send; i++) { if (elems[i] < 0 || elems[i] >= B.nOwned) throw std::runtime_error("halo send element out of range"); pos[i] = B.perm[elems[i]]; }
  c->sendList.upload(pos, c->stream);
  c->nSend = n_send;
  c->sendBuf.alloc((size_t)std::max(n_send, 1) * std::max(haloStride(c, 0), twoPass(c) ? haloStride(c, 1) : 0));
  SDG_CATCH
}

int sdg_uses_trace_rows(sdg_ctx* c) { return c->traceTU ? 1 : 0; }

int sdg_set_halo_rows(sdg_ctx* c, int32_t type, int32_t n_send, const int32_t* send_elem, const int32_t* send_face, int32_t n_recv,
                      const int32_t* recv_elem, const int32_t* recv_face) {
  SDG_TRY
  needFinal(c); needDevice(c); needType(c, type);
  if (!c->traceTU) throw std::runtime_error("trace-row halo: this context does not publish face traces (sdg_uses_trace_rows)");
  CUDA_OK(cudaSetDevice(c->cfg.device));
  const BlockPlan& B = c->plan.blk;
  std::vector<int> srow(n_send), rrow(n_recv);
  for (int i = 0; i < n_send; i++) {
    if (send_elem[i] < 0 || send_elem[i] >= B.nOwned || send_face[i] < 0 || send_face[i] >= 6) throw std::runtime_error("halo send row out of range");
    srow[i] = B.perm[send_elem[i]] * 6 + send_face[i];
  }
  for (int i = 0; i < n_recv; i++) {
    if (recv_elem[i] < B.nOwned || recv_elem[i] >= B.n || recv_face[i] < 0 || recv_face[i] >= 6) throw std::runtime_error("halo receive row is not a ghost row");
    rrow[i] = B.perm[recv_elem[i]] * 6 + recv_face[i];
  }
  c->sendList.upload(srow, c->stream); c->recvList.upload(rrow, c->stream);
  c->nSend = n_send; c->nRecvRows = n_recv; c->rowHalo = true;
  c->sendBuf.alloc((size_t)std::max(n_send, 1) * kRow); c->recvBuf.alloc((size_t)std::max(n_recv, 1) * kRow);
  SDG_CATCH
}

int sdg_halo_unpack(sdg_ctx* c, int32_t type, int32_t what, void* stream) {
  SDG_TRY
  needFinal(c); needDevice(c); needType(c, type);
  if (!c->rowHalo) return 0;   // element halo: the ghost range received the data directly
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (c->nRecvRows == 0) return 0;
  const int blocks = (int)std::min<size_t>(((size_t)c->nRecvRows * kRow + 255) / 256, 148 * 8);
  haloUnpackKernel<<<blocks, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(c->recvBuf.p, c->recvList.p, c->nRecvRows, kRow, haloField(c, what));
  c->launches++;
  CUDA_OK(cudaGetLastError());
  SDG_CATCH
}

int sdg_ipc_set_destination_units(sdg_ctx* c, int32_t n, const int64_t* dst_units) {
  SDG_TRY
  needFinal(c); needDevice(c);
  if (n != c->nSend) throw std::runtime_error("destination units: one per send unit");
  CUDA_OK(cudaSetDevice(c->cfg.device));
  std::vector<long long> d(dst_units, dst_units + n);
  c->dstUnit.upload(d, c->stream);
  SDG_CATCH
}

int sdg_halo_pack(sdg_ctx* c, int32_t type, int32_t what, void* stream) {
  SDG_TRY
  if (c->mx) throw std::runtime_error("not available on the dense-operator (triangle / mixed-type) path: single GPU, sdg_step only");
  needFinal(c); needDevice(c); needType(c, type);
  if (what != 0 && !(what == 1 && twoPass(c))) throw std::runtime_error("halo field: 0 = state, 1 = volume gradient (Navier-Stokes models and shock-capturing runs)");
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (what == 0) ensureTraces(c, c->latest, stream ? (cudaStream_t)stream : c->stream);
  if (c->nSend == 0) return 0;
  const int stride = haloStride(c, what);
  const double* src = haloField(c, what);
  const int blocks = (int)std::min<size_t>(((size_t)c->nSend * stride + 255) / 256, 148 * 8);
  haloPackKernel<<<blocks, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(src, c->sendList.p, c->nSend, stride, c->sendBuf.p);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  SDG_CATCH
}

int sdg_halo_buffers_device(sdg_ctx* c, int32_t type, int32_t what, void** send, int64_t* send_doubles, void** recv, int64_t* recv_doubles) {
  SDG_TRY
  if (c->mx) throw std::runtime_error("not available on the dense-operator (triangle / mixed-type) path: single GPU, sdg_step only");
  needFinal(c); needDevice(c); needType(c, type);
  if (what != 0 && !(what == 1 && twoPass(c))) throw std::runtime_error("halo field: 0 = state, 1 = volume gradient (Navier-Stokes models and shock-capturing runs)");
  const BlockPlan& B = c->plan.blk;
  const int64_t per = haloStride(c, what);
  double* base = haloField(c, what);
  *send = c->sendBuf.p; *send_doubles = (int64_t)c->nSend * per;
  if (c->rowHalo) { *recv = c->recvBuf.p; *recv_doubles = (int64_t)c->nRecvRows * per; }   // staging: sdg_halo_unpack scatters the rows
  else { *recv = base + (size_t)B.nOwned * per; *recv_doubles = (int64_t)B.nGhost * per; }
  SDG_CATCH
}

// ---- peer-memory halo exchange ------------------------------------------------------------------------------------------------------
int sdg_ipc_export(sdg_ctx* c, unsigned char* handles) {
  SDG_TRY
  needFinal(c); needDevice(c);
  if (c->mx) throw std::runtime_error("not available on the dense-operator (triangle / mixed-type) path");
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (!c->ipcFlags.p) { c->ipcFlags.alloc(64); c->ipcFlags.zero(c->stream); c->pushCounter.alloc(1); c->pushCounter.zero(c->stream); c->haloErr.alloc(1); c->haloErr.zero(c->stream); CUDA_OK(cudaStreamSynchronize(c->stream)); }
  void* ptrs[5] = {c->U[0].p, c->U[1].p, c->U[2].p, c->G.p, c->ipcFlags.p};
  if (c->traceTU) { ptrs[0] = c->TU[0].p; ptrs[1] = c->TU[1].p; ptrs[2] = c->TU[2].p; ptrs[3] = c->TV.p; }
  std::memset(handles, 0, 5 * sizeof(cudaIpcMemHandle_t));
  for (int i = 0; i < 5; i++) if (ptrs[i]) { cudaIpcMemHandle_t h; CUDA_OK(cudaIpcGetMemHandle(&h, ptrs[i])); std::memcpy(handles + i * sizeof(h), &h, sizeof(h)); }
  SDG_CATCH
}

int sdg_ipc_connect(sdg_ctx* c, int32_t n_peers, const unsigned char* handles, const int64_t* ghost_first, const int32_t* send_first,
                    const int32_t* send_count, const int32_t* slot_at_peer) {
  SDG_TRY
  needFinal(c); needDevice(c);
  if (n_peers < 0 || n_peers > 32) throw std::runtime_error("bad peer count");
  CUDA_OK(cudaSetDevice(c->cfg.device));
  c->peerLinks.assign(n_peers, PeerDev{});
  for (int p = 0; p < n_peers; p++) {
    PeerDev& L = c->peerLinks[p];
    for (int i = 0; i < 5; i++) {
      cudaIpcMemHandle_t h; std::memcpy(&h, handles + ((size_t)p * 5 + i) * sizeof(h), sizeof(h));
      bool zero = true; for (size_t b = 0; b < sizeof(h); b++) zero = zero && reinterpret_cast<const unsigned char*>(&h)[b] == 0;
      void* q = nullptr;
      if (!zero) { CUDA_OK(cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess)); c->ipcOpened.push_back(q); }
      if (i < 4) L.dst[i] = static_cast<double*>(q); else L.flags = static_cast<long long*>(q);
    }
    L.ghostFirst = ghost_first[p]; L.sendFirst = send_first[p]; L.sendCount = send_count[p]; L.slot = slot_at_peer[p];
    if (L.slot < 0 || L.slot >= 64) throw std::runtime_error("bad flag slot");
  }
  c->peerDev.upload(c->peerLinks, c->stream);
  SDG_CATCH
}

int sdg_halo_push(sdg_ctx* c, int32_t type, int32_t what, void* stream) {
  SDG_TRY
  needFinal(c); needDevice(c); needType(c, type);
  if (what != 0 && !(what == 1 && twoPass(c))) throw std::runtime_error("halo field: 0 = state, 1 = volume gradient (Navier-Stokes models and shock-capturing runs)");
  if (c->peerLinks.empty() && c->nSend > 0) throw std::runtime_error("sdg_ipc_connect has not been called");
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (what == 0) ensureTraces(c, c->latest, stream ? (cudaStream_t)stream : c->stream);
  c->pushEpoch++;
  if (c->peerLinks.empty()) return 0;
  const int stride = haloStride(c, what);
  const double* src = haloField(c, what);
  const int which = what == 1 ? 3 : c->latest;
  static const int maxBlocks = getenv("SDG_PUSH_BLOCKS") ? std::max(1, atoi(getenv("SDG_PUSH_BLOCKS"))) : 148 * 2;   // 2 CTAs per SM measured best at 4 GPUs (96: 280, 296: 285, 592: 282 GDOF/s): more blocks steal SM slots from the interior launch
  if (c->rowHalo && !c->dstUnit.p && c->nSend > 0) throw std::runtime_error("trace-row halo: sdg_ipc_set_destination_units has not been called");
  const int blocks = (int)std::max<size_t>(1, std::min<size_t>(((size_t)c->nSend + 7) / 8, (size_t)maxBlocks));   // 8 warps per block, one unit per warp
  haloPushKernel<<<blocks, 256, 0, stream ? (cudaStream_t)stream : c->stream>>>(src, c->sendList.p, c->nSend, stride, c->peerDev.p, (int)c->peerLinks.size(), which,
                                                                                 c->pushCounter.p, c->pushEpoch, c->rowHalo ? c->dstUnit.p : nullptr);
  c->launches++;
  CUDA_OK(cudaGetLastError());
  SDG_CATCH
}

int sdg_halo_wait(sdg_ctx* c, void* stream) {
  SDG_TRY
  needFinal(c); needDevice(c);
  CUDA_OK(cudaSetDevice(c->cfg.device));
  if (c->peerLinks.empty()) return 0;
  // a rank may legitimately lag by a multi-GB host copy or a first-launch module load: generous default, SDG_HALO_TIMEOUT_S overrides
  static const double timeoutS = getenv("SDG_HALO_TIMEOUT_S") ? std::max(1.0, atof(getenv("SDG_HALO_TIMEOUT_S"))) : 120.0;
  haloWaitKernel<<<1, 32, 0, stream ? (cudaStream_t)stream : c->stream>>>(c->ipcFlags.p, (int)c->peerLinks.size(), c->pushEpoch, c->haloErr.p,
                                                                          (unsigned long long)(timeoutS * 1e9));
  c->launches++;
  CUDA_OK(cudaGetLastError());
  SDG_CATCH
}

}  // extern "C"
