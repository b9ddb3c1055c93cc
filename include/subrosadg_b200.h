/* subrosadg_b200.h — C ABI of the B200-native (sm_100a) DG residual + explicit SSP-RK path.
 *
 * The reference (SubrosaDG) has no FFI: its seam is the C++ class SubrosaDG::Solver<SimulationControl>
 * (src/Solver/SolveControl.cpp:327-436) driven by System<SC>::solve() (src/Utils/SystemControl.cpp:159-195).
 * Every entry point below replaces one member of that class (cited per function); the header-only shim
 * include/SubrosaDG_b200/SubrosaDG.hpp maps the reference's template configuration surface onto this ABI.
 *
 * Conventions: plain pointers and sizes, caller-owned HOST buffers unless the name says `_device`; every function
 * returns 0 on success and non-zero on failure with the message available from sdg_last_error().  NaNs in the state
 * propagate into `relative_error` (the reference's only run-time failure signal, SystemControl.cpp:185-191).
 * There is no CPU fallback: sdg_create fails when no CUDA device is usable.
 *
 * Data order at the seam (identical to the reference):
 *   element types      ElementEnum values (src/Utils/Enum.cpp:28-36): 1 line, 2 triangle, 3 quadrangle, 6 hexahedron.
 *                      One line, quadrangle or hexahedron block -> collocation tensor kernels; triangle blocks or several types in
 *                      one 2-D mesh -> dense-operator kernels in the reference's modal representation (single GPU).
 *   node coordinates   [n][nn][D], gmsh node order of Lagrange order `geom_order` (PerElementMesh::node_coordinate_,
 *                      src/Mesh/ReadControl.cpp:86-92)
 *   modal state        [n][Nb][Nv] = Eigen::Matrix<Real,Nv,Nb> column-major per element (SolveControl.cpp:45-58),
 *                      conserved order rho, rho u.., rho E (VariableConvertor.cpp:28-68)
 *   quadrature arrays  [n][Nq][..]  volume points in the reference's "Gauss{2p}" order (src/Mesh/Quadrature.cpp:27-34)
 *   face records       interior faces first, then boundary faces (AdjacencyElementMesh, ReadControl.cpp:72-83,143-155)
 */
#ifndef SUBROSADG_B200_H
#define SUBROSADG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdg_ctx sdg_ctx;

/* POD image of SimulationControl<...> (src/Solver/SimulationControl.cpp:1197-1279); integer fields carry the enum
 * values of src/Utils/Enum.cpp. */
typedef struct sdg_config {
  int32_t dim;        /* DimensionEnum 1..3 */
  int32_t p;          /* PolynomialOrderEnum 1..5 */
  int32_t model;      /* EquationModelEnum: 0 CompresibleEuler 1 CompresibleNS 2 IncompresibleEuler 3 IncompresibleNS */
  int32_t eos;        /* EquationOfStateEnum: 0 IdealGas 1 WeakCompressibleFluid */
  int32_t transport;  /* TransportModelEnum: 0 None 1 Constant 2 Sutherland */
  int32_t conv_flux;  /* ConvectiveFluxEnum: 0 Central 1 LaxFriedrichs 2 HLLC 3 Roe 4 Exact */
  int32_t visc_flux;  /* ViscousFluxEnum: 0 None 1 BR1 2 BR2 */
  int32_t source;     /* SourceTermEnum: 0 None 1 Boussinesq */
  int32_t rk;         /* TimeIntegrationEnum: 0 ForwardEuler 1 HeunRK2 2 SSPRK3 */
  int32_t device;     /* CUDA device ordinal */
  int32_t chunk;      /* elements per thread block (0 = automatic; -1 = force the dense-operator path, diagnostics) */
  int32_t reorder;    /* 1: internal space-filling-curve element order (invisible at the seam); 0: keep caller order */
  double cp, cv;      /* ThermodynamicModel<Constant>, PhysicalModel.cpp:26-38 */
  double mu;          /* TransportModel dynamic viscosity (reference value for Sutherland), PhysicalModel.cpp:83-123 */
  double c0, rho0;    /* EquationOfState<WeakCompressibleFluid>, PhysicalModel.cpp:57-78 */
  double beta, t_ref; /* SourceTerm<Boussinesq>, SourceTerm.cpp:29-58 */
} sdg_config;

const char* sdg_last_error(void);
int sdg_version(void);

/* System<SC>() + setters (SystemControl.cpp:55-140) collapsed into one POD. */
int sdg_create(const sdg_config* cfg, sdg_ctx** out);
void sdg_destroy(sdg_ctx* ctx);

/* Mesh<SC>::readMeshElement output for one element type (src/Mesh/Element.cpp:31-70): node coordinates only; Jacobians,
 * M^-1 and normals (src/Mesh/Geometry.cpp) are recomputed by the library and kept resident in HBM.
 * The last `n_ghost` elements are halo copies owned by another rank: they are read as neighbours, never advanced. */
int sdg_add_elements(sdg_ctx* ctx, int32_t type, int32_t n, int32_t n_ghost, int32_t geom_order, const double* coords);

/* AdjacencyElementMesh records (src/Mesh/Adjacency.cpp:330-430): parent_index_each_type_ (le/re),
 * parent_gmsh_type_number_ as ElementEnum (lt/rt), adjacency_sequence_in_parent_ (lf/rf), adjacency_right_rotation_,
 * boundary_condition_type_ (BoundaryConditionEnum), gmsh_physical_index_.  right_* are ignored for boundary faces. */
int sdg_set_faces(sdg_ctx* ctx, int32_t n_int, int32_t n_bnd, const int32_t* le, const int32_t* lt, const int32_t* lf,
                  const int32_t* re, const int32_t* rt, const int32_t* rf, const int32_t* rot, const int32_t* bc,
                  const int32_t* phys);

/* System::synchronize() (SystemControl.cpp:142-157): geometry, index flattening, upload. */
int sdg_finalize(sdg_ctx* ctx);

/* out[8] = n, Nb, Nq, Nf, Naq, nn, Nqf, Nv   (SimulationControl.cpp:74-97,243-379,1143-1151) */
int sdg_sizes(sdg_ctx* ctx, int32_t type, int32_t* out);

/* quadrature_node_coordinate_ of elements [n][Nq][D] / of boundary faces [n_bnd][Nqf][D] (Geometry.cpp:44-86): the
 * points at which the host evaluates the user's InitialCondition / BoundaryCondition callbacks. */
int sdg_get_quadrature_coordinates(sdg_ctx* ctx, int32_t type, double* xq);
int sdg_get_boundary_quadrature_coordinates(sdg_ctx* ctx, double* xb);

/* Solver::initializeSolver (InitialCondition.cpp:85-149): user primitive values (rho, u.., T) at the volume
 * quadrature points [n][Nq][Nv] -> least-squares projection; at the boundary-face points [n_bnd][Nqf][Nv] ->
 * boundary_dummy_variable_.  sdg_set_boundary_primitive is also Solver::updateBoundaryVariable
 * (BoundaryCondition.cpp:29-74) for BoundaryTimeEnum::TimeVarying. */
int sdg_set_state_from_primitive(sdg_ctx* ctx, int32_t type, const double* prim);
int sdg_set_boundary_primitive(sdg_ctx* ctx, const double* prim);

/* variable_basis_function_coefficient_ in the reference's modal (H1Legendre) basis, [n][Nb][Nv]
 * (what Solver::writeRawBinary serialises, src/View/RawBinary.cpp:75-88). */
int sdg_set_state(sdg_ctx* ctx, int32_t type, const double* U);
int sdg_get_state(sdg_ctx* ctx, int32_t type, double* U);
/* conserved variables / total gradient at the volume quadrature points, [n][Nq][Nv] / [n][Nq][Nv*D] (row var*D+dir) */
int sdg_get_state_at_quadrature(sdg_ctx* ctx, int32_t type, double* Uq);
int sdg_get_gradient_at_quadrature(sdg_ctx* ctx, int32_t type, double* Gq);
/* The gradient blocks Solver::writeRawBinary serialises for Navier-Stokes models (src/View/RawBinary.cpp:75-88, 89-154).  As in the
 * reference, whose arrays keep what the LAST RK STAGE computed, after sdg_step they are the gradient of the INPUT of the last stage of the
 * last step (not of the new state; pinned by the reference-written files tests/golden/reference_raw_*.zst); after a state setter, before any
 * step, they are evaluated for the current state (the reference's arrays are uninitialised at that point):
 *   sdg_get_gradient_state           variable_gradient_basis_function_coefficient_ of every element, [n][Nb][Nv*D] (row var*D+dir);
 *   sdg_get_boundary_gradient_state  per boundary face, in the order of sdg_set_faces, the block written next to the parent's state:
 *                                    BR1 the total gradient, BR2 variable_volume_gradient_ + variable_interface_gradient_ of THAT
 *                                    local face; rows of Nb(parent type) x Nv*D doubles, concatenated. */
int sdg_get_gradient_state(sdg_ctx* ctx, int32_t type, double* G);
int sdg_get_boundary_gradient_state(sdg_ctx* ctx, double* Gb);

/* ShockCapturingEnum::ArtificialViscosity (Euler models).  System::setArtificialViscosity (SystemControl.cpp:105-108) + the mesh data the
 * reference's Solver::calculateArtificialViscosity reads (SpatialDiscrete.cpp:37-192): node_number = Mesh::node_number_; per block the
 * 0-based tags of the corner nodes, [n][kBasicNodeNumber] in gmsh node order (PerElementMesh::node_tag_, ReadControl.cpp:64), and
 * inner_radius_ [n] (gmsh "innerRadius" element quality, Geometry.cpp:31-41).  All three calls precede sdg_finalize.  Every sdg_step then
 * evaluates the element indicator, the node maximum and the corner values once per step and adds eps * grad(U) to the volume and face
 * fluxes of every stage (SpatialDiscrete.cpp:210-253, 529-631, 694-746, 786-819; ViscousFlux.cpp:105-113,126-136,172-186). */
int sdg_set_artificial_viscosity(sdg_ctx* ctx, double empirical_tolerance, double artificial_viscosity_factor, int32_t node_number);
int sdg_set_element_nodes(sdg_ctx* ctx, int32_t type, const int32_t* node_tag, const double* inner_radius);
/* Solver::node_artificial_viscosity_ [node_number] (the tail of the RawBinary payload) and variable_artificial_viscosity_ [n][kBasicNodeNumber]
 * as of the last step; sdg_update_artificial_viscosity re-evaluates them for the current state (diagnostics / parity). */
int sdg_get_node_artificial_viscosity(sdg_ctx* ctx, double* out);
int sdg_get_element_artificial_viscosity(sdg_ctx* ctx, int32_t type, double* out);
int sdg_update_artificial_viscosity(sdg_ctx* ctx);
/* Partitioned runs (one context per GPU): the node maximum of Solver::calculateArtificialViscosity -- in the reference the cwiseMax
 * combine over the TBB threads' node arrays, SpatialDiscrete.cpp:89-108 -- spans the ranks, the one collective on this path.
 * sdg_step_begin leaves the maximum over the context's OWNED elements in the device array sdg_av_node_buffer returns ([node_number],
 * node tags global); the caller max-reduces it over the ranks in place (ncclAllReduce, ncclMax) and calls sdg_av_store, which rewrites
 * variable_artificial_viscosity_ of owned and ghost elements from it. */
int sdg_av_node_buffer(sdg_ctx* ctx, void** device_nodes, int64_t* count);
int sdg_av_store(sdg_ctx* ctx, void* stream);

/* ViewVariable::get (src/Solver/VariableConvertor.cpp:754-872) on the device: the scalar field `variable` (ViewVariableEnum value,
 * src/Utils/Enum.cpp: 0 Density, 1 Velocity, 2 Temperature, 3 Pressure, 4 SoundSpeed, 5 MachNumber, 6 Entropy, 7 Vorticity, 9
 * ArtificialViscosity, 10-12 VelocityX/Y/Z, 13-15 MachNumberX/Y/Z, 16-18 VorticityX/Y/Z, 19-21 HeatFluxX/Y/Z) of the resident state at
 * the volume quadrature points, [n][Nq].  Gradient-based entries use the total gradient of the current state (Navier-Stokes); entries
 * that do not exist for the equation set return what the reference's switch falls through to. */
int sdg_get_view_variable(sdg_ctx* ctx, int32_t type, int32_t variable, double* out);

/* Solver::calculateDeltaTime (TimeIntegration.cpp:104-179) */
int sdg_compute_dt(sdg_ctx* ctx, double cfl, double* dt);

/* Solver::stepSolver x n_steps (TimeIntegration.cpp:326-350); relative_error[Nv] = Solver::relative_error_ of the last
 * step (TimeIntegration.cpp:279-324), may be NULL. */
int sdg_step(sdg_ctx* ctx, double dt, int32_t n_steps, double* relative_error);

/* Same as sdg_step, additionally returning the device time of the n_steps (CUDA events on the context's stream). */
int sdg_step_timed(sdg_ctx* ctx, double dt, int32_t n_steps, double* relative_error, float* milliseconds);

/* One call of Solver::stepSolver (TimeIntegration.cpp:326-350) on a state that lives in HOST memory: the modal coefficients
 * U_in [n][Nb][Nv] (the layout of sdg_set_state) are the step's input, U_out receives the coefficients after the step, relative_error
 * [Nv] (may be NULL) Solver::relative_error_.  Same result, bit for bit, as sdg_set_state -> sdg_step(dt, 1) -> sdg_get_state; for
 * a single line / quadrangle / hexahedron block of at least 8192 elements on one GPU, without shock capturing, the three phases are
 * STREAMED: the caller's element order is cut into
 * upload groups, the stages of a thread-block chunk run as soon as the chunk's and its face neighbours' inputs have arrived, and a
 * group travels back while later groups are still arriving (PCIe in both directions at once).  U_in and U_out may be the same buffer;
 * pinned host memory is what makes the copies asynchronous.  SDG_NO_HOST_PIPE=1 forces the phase-after-phase composition,
 * SDG_HOST_PIPE_GROUPS=<G> sets the number of upload groups (default n / 16384, 2..64). */
int sdg_step_host(sdg_ctx* ctx, int32_t type, double dt, const double* U_in, double* U_out, double* relative_error);

/* Diagnostics of sdg_step_host: number of upload groups (0: the context does not stream) and the fraction of them whose download starts
 * before the last upload has arrived. */
int sdg_step_host_info(sdg_ctx* ctx, int32_t* groups, double* early_fraction);

/* Parity hook: one residual evaluation of the current state.  Rmodal [n][Nb][Nv] = variable_residual_
 * (SpatialDiscrete.cpp:1016-1032); rhsq [n][Nq][Nv] = (R M^-1) Phi^T, i.e. dU/dt at the quadrature points.
 * Either may be NULL. */
int sdg_residual(sdg_ctx* ctx, int32_t type, double* Rmodal, double* rhsq);

/* ---- element-block partitioning across the GPUs of one box (one context per rank) --------------------------------
 * Device-resident halo staging: sdg_halo_pack gathers the states of the `n_send` local elements listed by
 * sdg_set_halo_send into a contiguous device buffer; the ghost elements of sdg_add_elements form one contiguous
 * device range that NCCL receives into directly.  `what`: 0 conserved state, 1 volume-gradient state (NS).
 * The pointers are device addresses on cfg.device; `stream` is a cudaStream_t (0 = the context's stream). */
int sdg_set_halo_send(sdg_ctx* ctx, int32_t type, int32_t n_send, const int32_t* elems);
int sdg_halo_pack(sdg_ctx* ctx, int32_t type, int32_t what, void* stream);
int sdg_halo_buffers_device(sdg_ctx* ctx, int32_t type, int32_t what, void** send, int64_t* send_doubles, void** recv,
                            int64_t* recv_doubles);
/* Doubles exchanged per halo element for field `what`: the element's state / volume-gradient coefficients, or — P3 hexahedra with
 * Navier-Stokes, whose kernels read published face traces — its six face-trace rows (0: conserved variables, 1: viscous normal flux). */
int sdg_halo_doubles_per_element(sdg_ctx* ctx, int32_t what);
/* Trace-row halo (contexts whose kernels read published face traces, sdg_uses_trace_rows() == 1): a cut face needs ONE row of its
 * remote parent — 5 variables x 16 points = 640 bytes — instead of the whole element (640 B per face per direction).
 * sdg_set_halo_rows lists the (owned element, local face) rows to send and the (ghost element, local face) rows to receive, both in the
 * order the peers agree on; sdg_halo_pack / sdg_halo_buffers_device then work on rows (the receive buffer is a staging area) and
 * sdg_halo_unpack scatters received rows to their places.  sdg_halo_doubles_per_element returns the row size (80) in this mode. */
int sdg_uses_trace_rows(sdg_ctx* ctx);
int sdg_set_halo_rows(sdg_ctx* ctx, int32_t type, int32_t n_send, const int32_t* send_elem, const int32_t* send_face, int32_t n_recv,
                      const int32_t* recv_elem, const int32_t* recv_face);
int sdg_halo_unpack(sdg_ctx* ctx, int32_t type, int32_t what, void* stream);
/* peer-memory transport of the trace-row halo: row index, in the peer's arrays, of every send row (same order as the send list) */
int sdg_ipc_set_destination_units(sdg_ctx* ctx, int32_t n, const int64_t* dst_units);
/* Split stepping used by the multi-GPU driver: stage `s` of the current step, restricted to thread blocks that do not
 * (part 0) / do (part 1) touch ghost elements; part -1 = all.  sdg_step_begin / sdg_step_end bracket one step. */
int sdg_step_begin(sdg_ctx* ctx, double dt);
int sdg_stage_pass(sdg_ctx* ctx, int32_t stage, int32_t pass, int32_t part, void* stream);
int sdg_step_end(sdg_ctx* ctx, double* relative_error_sum /* Nv sums over OWNED elements, not yet divided */);
int sdg_num_passes(sdg_ctx* ctx);     /* passes per stage: 1 (Euler) or 2 (NS: gradient pass, residual pass) */
int sdg_num_stages(sdg_ctx* ctx);
void* sdg_stream(sdg_ctx* ctx);       /* the context's cudaStream_t */
int sdg_synchronize(sdg_ctx* ctx);

/* Peer-memory halo exchange (one process per GPU on one NVLink / NVSwitch box): instead of staging + send/recv, sdg_halo_push
 * gathers the send elements and stores them straight into the peers' ghost ranges, then raises this rank's arrival flag at every
 * peer; sdg_halo_wait makes `stream` wait for the matching pushes of all peers.  Set-up: every rank exports the CUDA IPC handles of
 * its U[0..2], G and flag allocations (sdg_ipc_export, 5 x 64 bytes), the driver exchanges them, and sdg_ipc_connect opens the
 * peers' handles: per peer the element index of the first ghost element this rank feeds in the PEER's arrays (its n_owned + the
 * start of its receive range for this rank), this rank's range in its own send list, and this rank's flag slot at the peer
 * (its position in the peer's ascending peer list).  Ranks must issue the same sequence of pushes. */
int sdg_ipc_export(sdg_ctx* ctx, unsigned char* handles /* 5 x 64 bytes */);
int sdg_ipc_connect(sdg_ctx* ctx, int32_t n_peers, const unsigned char* handles /* n_peers x 5 x 64 */, const int64_t* ghost_first,
                    const int32_t* send_first, const int32_t* send_count, const int32_t* slot_at_peer);
int sdg_halo_push(sdg_ctx* ctx, int32_t type, int32_t what, void* stream);
int sdg_halo_wait(sdg_ctx* ctx, void* stream);

/* Device-resident access for callers that already hold the state in HBM (bench.py's `value` leg):
 * modal state buffer [n][Nb][Nv] <-> internal representation, both on the device. */
int sdg_set_state_device(sdg_ctx* ctx, int32_t type, const void* U_device);
int sdg_get_state_device(sdg_ctx* ctx, int32_t type, void* U_device);

/* Diagnostics (host plan, no device needed): copies one of the flattened arrays and/or returns its length.
 * what: 0 geoE  1 invjw  2 minEdge  3 geoF  4 Phi[Nq][Nb]  5 1-D differentiation matrix  6 end-point interpolation
 *       7 Gauss abscissae  8 Gauss weights (doubles);  10 perm  11 chunkFaceOff  12 faceRec  13 chunkInterior
 *       14 chunkBoundary  15 {affine, K, nChunks, nOwned}  16 face-point permutations [4][Nqf]  17 faceBase  18 nodeFacePt
 *       19 modal function index triples (int32).  Element arrays are in INTERNAL order (position perm[e]).
 * Meshes with triangle blocks / several element types (dense-operator path): what = 100*type + {0 Phi[Nq][Nb], 1 grad Phi[Nq][D][Nb],
 *       2 Phi_f[Naq][Nb], 3 least-squares projection[Nb][Nq], 4 detJ w[n][Nq], 5 (J^T)^-1 detJ w[n][Nq][D*D], 6 M^-1[n][Nb][Nb],
 *       7 minEdge[n]}; 90 face normals[nf][Nqf][D]; 91 face |J| w[nf][Nqf] (doubles, caller element order). */
int sdg_debug_plan(sdg_ctx* ctx, int32_t what, double* out_d, int32_t* out_i, int64_t* count);

/* counters: number of kernels this library launched since creation (bench.py's gpu_launches) */
int64_t sdg_launch_count(sdg_ctx* ctx);

/* Parity hook: the pointwise device functions (conserved -> computational variables, the five Riemann fluxes of
 * src/Solver/ConvectiveFlux.cpp, the six BoundaryConditionImpl of BoundaryCondition.cpp:79-547 with their gradient states and
 * modifyBoundaryVariable, VariableGradient::calculatePrimitiveFromConserved, ViscousFlux.cpp:59-124, SourceTerm.cpp:29-58) evaluated on
 * caller-supplied points.  cfg = {dim, model, eos, transport, conv_flux, source} (Enum.cpp values), params = {cp, cv, mu, c0, rho0, beta,
 * t_ref}; `what` and the row layouts are those of tests/golden/reference_physics.json, which was generated from the reference's own
 * sources (oracle/ref_physics.cpp).  Needs a CUDA device. */
int sdg_debug_physics(const int32_t* cfg, const double* params, int32_t what, int32_t bc, int32_t n, const double* in, double* out);

#ifdef __cplusplus
}
#endif
#endif
