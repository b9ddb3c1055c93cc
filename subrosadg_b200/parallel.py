"""Element-block partitioning across the GPUs of one box and the per-stage face-neighbour (halo) exchange.

The reference is a single shared-memory process (tbb::parallel_for over elements / faces); its only
coupling between elements is through face traces (src/Solver/SpatialDiscrete.cpp:633-748,844-911) and two global
reductions (the time step, src/Solver/TimeIntegration.cpp:127-130, and relative_error_, :295-323).  That makes the path
shard naturally:

* rank r owns the contiguous element-index block [r*Ne/P, (r+1)*Ne/P) of every element type (z-slabs for the
  lexicographic structured meshes of the benchmark);
* every face that touches an owned element is kept, with the reference's left/right roles and rotation unchanged, so the
  per-element arithmetic is identical to the single-GPU run; the parents that live on another rank are appended to the
  element block as ghost elements (read as neighbours, never advanced);
* once per stage pass the states of the ghost elements are refreshed: a device pack kernel (sdg_halo_pack) gathers the
  owned elements other ranks need into one contiguous buffer per peer, NCCL send/recv (torch.distributed P2P over
  NVLink 5 / NVSwitch) delivers them straight into the ghost range of the state array, and the thread blocks that do
  not touch a ghost element (sdg_stage_pass part 0) run concurrently with the exchange; the blocks that do (part 1)
  wait for it.  Navier-Stokes adds the same exchange for the volume-gradient field between its two passes;
* relative_error_ is a sum over owned elements followed by an all-reduce(sum); the time step is an all-reduce(min).

`partition` is pure numpy (deterministic integer maps, covered by CPU tests); `HaloExchange` works on any
torch.distributed backend (gloo on CPU tensors in the tests, NCCL on device pointers in production).
"""
from __future__ import annotations

import ctypes
import json
import os
import time
from dataclasses import dataclass, field

import numpy as np

from . import mesh as M


# ---- partitioning ------------------------------------------------------------------------------------------------------------
@dataclass
class Partition:
    """The part of a mesh one rank works on (all indices int64 numpy arrays)."""
    rank: int
    world: int
    mesh: M.Mesh                       # owned elements followed by ghost elements; faces touching an owned element
    etype: int
    lo: int                            # owned global element range [lo, hi)
    hi: int
    n_owned: int
    n_ghost: int
    ghost_global: np.ndarray           # global ids of the ghost elements, ascending (=> grouped by owner rank)
    peers: list = field(default_factory=list)      # ascending peer ranks
    send_local: dict = field(default_factory=dict)  # peer -> local (owned) element indices, ascending global id
    recv_range: dict = field(default_factory=dict)  # peer -> (first ghost index relative to n_owned, count)
    face_global: np.ndarray | None = None           # global face ids of the local faces (interior first, then boundary)
    n_elements_global: int = 0
    # trace-row halo: per peer, the (local owned element, local face) rows to send and the (local ghost element, local face) rows to
    # receive, both ordered by global face id — the two ranks of a cut face enumerate the same faces in the same order
    send_rows: dict = field(default_factory=dict)
    recv_rows: dict = field(default_factory=dict)


def block_bounds(n: int, world: int) -> np.ndarray:
    """Contiguous element-index blocks: rank r owns [b[r], b[r+1])."""
    return np.array([(r * n) // world for r in range(world + 1)], dtype=np.int64)


def partition(mesh: M.Mesh, rank: int, world: int) -> Partition:
    """Rank `rank`'s share of a single-element-type mesh."""
    if len(mesh.blocks) != 1:
        raise ValueError("partitioning handles meshes with one element type")
    etype = next(iter(mesh.blocks))
    blk = mesh.blocks[etype]
    coords = blk["coords"]
    ne = coords.shape[0]
    b = block_bounds(ne, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    f = mesh.faces
    n_int, n_bnd = int(f["n_int"]), int(f["n_bnd"])
    le = np.asarray(f["le"], dtype=np.int64)
    re = np.asarray(f["re"], dtype=np.int64)
    interior = np.arange(n_int + n_bnd) < n_int
    own_l = (le >= lo) & (le < hi)
    own_r = interior & (re >= lo) & (re < hi)
    sel = np.flatnonzero(own_l | own_r)            # keeps the reference order: interior faces first, then boundary faces
    gl = le[sel][~own_l[sel]]
    gr = re[sel][interior[sel] & ~own_r[sel]]
    ghosts = np.unique(np.concatenate([gl, gr]))   # ascending global id == grouped by owner (owners are index blocks)
    n_owned, n_ghost = hi - lo, int(ghosts.size)

    def to_local(g):
        g = np.asarray(g, dtype=np.int64)
        out = g - lo
        outside = (g < lo) | (g >= hi)
        if outside.any():
            out = out.copy()
            out[outside] = n_owned + np.searchsorted(ghosts, g[outside])
        return out

    loc_faces = {k: np.asarray(f[k])[sel].copy() for k in ("le", "lt", "lf", "re", "rt", "rf", "rot", "bc", "phys")}
    loc_faces["le"] = to_local(le[sel]).astype(np.int32)
    r_loc = np.full(sel.size, -1, dtype=np.int64)
    isel = interior[sel]
    r_loc[isel] = to_local(re[sel][isel])
    loc_faces["re"] = r_loc.astype(np.int32)
    loc_faces["n_int"] = int(isel.sum())
    loc_faces["n_bnd"] = int(sel.size - isel.sum())
    loc_coords = np.concatenate([coords[lo:hi], coords[ghosts]]) if n_ghost else np.ascontiguousarray(coords[lo:hi])
    local = M.Mesh(dim=mesh.dim, blocks={etype: dict(coords=loc_coords, geom_order=blk["geom_order"], corners=None)}, faces=loc_faces,
                   phys_bc=dict(mesh.phys_bc), info=dict(mesh.info, partition=(rank, world)))
    part = Partition(rank=rank, world=world, mesh=local, etype=etype, lo=lo, hi=hi, n_owned=n_owned, n_ghost=n_ghost, ghost_global=ghosts,
                     face_global=sel, n_elements_global=ne)
    owner = np.searchsorted(b, ghosts, side="right") - 1
    for q in np.unique(owner):
        q = int(q)
        idx = np.flatnonzero(owner == q)
        part.recv_range[q] = (int(idx[0]), int(idx.size))
        part.peers.append(q)
    # what the peers need from this rank: owned elements that share a face with an element owned by the peer
    other_l = np.where(own_r[sel] & ~own_l[sel], le[sel], -1)       # remote left parent of a face whose right parent is owned
    other_r = np.where(own_l[sel] & isel & ~own_r[sel], re[sel], -1)  # remote right parent of a face whose left parent is owned
    mine_for_l = re[sel]    # the owned element seen by the remote left parent
    mine_for_r = le[sel]
    pairs_owner = np.concatenate([np.searchsorted(b, other_l[other_l >= 0], side="right") - 1,
                                  np.searchsorted(b, other_r[other_r >= 0], side="right") - 1])
    pairs_mine = np.concatenate([mine_for_l[other_l >= 0], mine_for_r[other_r >= 0]])
    for q in np.unique(pairs_owner):
        q = int(q)
        part.send_local[q] = (np.unique(pairs_mine[pairs_owner == q]) - lo).astype(np.int64)
        if q not in part.peers:
            part.peers.append(q)
    part.peers.sort()
    for q in part.peers:
        part.send_local.setdefault(q, np.zeros(0, dtype=np.int64))
        part.recv_range.setdefault(q, (0, 0))
    # cut faces in ascending global face id: the owned parent's row goes out, the remote parent's row comes in
    lf_g, rf_g = np.asarray(f["lf"], dtype=np.int64)[sel], np.asarray(f["rf"], dtype=np.int64)[sel]
    cut_l = own_l[sel] & isel & ~own_r[sel]          # I own the left parent, the right one is remote
    cut_r = own_r[sel] & ~own_l[sel]                 # I own the right parent
    mine_e = np.where(cut_l, le[sel], re[sel]); mine_f = np.where(cut_l, lf_g, rf_g)
    theirs_e = np.where(cut_l, re[sel], le[sel]); theirs_f = np.where(cut_l, rf_g, lf_g)
    cut = np.flatnonzero(cut_l | cut_r)
    cut_owner = np.searchsorted(b, theirs_e[cut], side="right") - 1
    for q in part.peers:
        k = cut[cut_owner == q]
        part.send_rows[q] = np.stack([mine_e[k] - lo, mine_f[k]], axis=1).astype(np.int64) if k.size else np.zeros((0, 2), dtype=np.int64)
        part.recv_rows[q] = np.stack([to_local(theirs_e[k]), theirs_f[k]], axis=1).astype(np.int64) if k.size else np.zeros((0, 2), dtype=np.int64)
    return part


def partition_node_data(mesh: M.Mesh, part: Partition):
    """Shock-capturing runs: (node tags, node number, inner radii) of a partition's owned + ghost elements with the tags of the GLOBAL mesh
    (Mesh::node_number_, PerElementMesh::node_tag_ / inner_radius_), so that the ranks' node arrays can be max-reduced entry by entry."""
    tags, n_nodes = M.node_tags(mesh)
    ids = np.concatenate([np.arange(part.lo, part.hi, dtype=np.int64), np.asarray(part.ghost_global, dtype=np.int64)])
    t = part.etype
    return {t: np.ascontiguousarray(tags[t][ids])}, n_nodes, {t: np.ascontiguousarray(M.inner_radius(mesh, t)[ids])}


# ---- halo exchange (any torch.distributed backend) ---------------------------------------------------------------------------
class HaloExchange:
    """One message per peer and direction.  `send` holds the packed states of send_local[peer] for all peers back to
    back (ascending peer rank), `recv` is the ghost range; both are 1-D float64 torch tensors (CPU or CUDA)."""

    def __init__(self, part: Partition, group=None, rows: bool = False):
        """rows = False: units are whole elements (received straight into the ghost range); rows = True: units are (element, face)
        trace rows (received into a staging buffer, scattered by sdg_halo_unpack)."""
        self.part = part
        self.group = group
        self.rows = rows
        off, roff = 0, 0
        self.send_off, self.send_count, self.recv_off, self.recv_count = {}, {}, {}, {}
        for q in part.peers:
            self.send_off[q] = off
            self.send_count[q] = int(part.send_rows[q].shape[0]) if rows else int(part.send_local[q].size)
            off += self.send_count[q]
            if rows:
                self.recv_off[q], self.recv_count[q] = roff, int(part.recv_rows[q].shape[0])
                roff += self.recv_count[q]
            else:
                self.recv_off[q], self.recv_count[q] = part.recv_range[q]
        self.n_send = off
        self.n_recv = roff if rows else int(part.n_ghost)
        self.send_elems = (np.concatenate([part.send_local[q] for q in part.peers]) if part.peers else np.zeros(0, dtype=np.int64)).astype(np.int32)
        cat = lambda d: (np.concatenate([d[q] for q in part.peers]) if part.peers else np.zeros((0, 2), dtype=np.int64)).astype(np.int32)
        self.send_rows, self.recv_rows = cat(part.send_rows), cat(part.recv_rows)

    def start(self, send, recv, elem_doubles: int):
        import torch.distributed as dist
        ops = []
        for q in self.part.peers:
            ns = self.send_count[q]
            r0, nr = self.recv_off[q], self.recv_count[q]
            if nr:
                ops.append(dist.P2POp(dist.irecv, recv[r0 * elem_doubles:(r0 + nr) * elem_doubles], q, group=self.group))
            if ns:
                s0 = self.send_off[q]
                ops.append(dist.P2POp(dist.isend, send[s0 * elem_doubles:(s0 + ns) * elem_doubles], q, group=self.group))
        return dist.batch_isend_irecv(ops) if ops else []

    @staticmethod
    def finish(reqs):
        for r in reqs:
            r.wait()


def _set_halo(lib, S, etype, halo):
    """Registers the halo send (and, for trace rows, receive) units of one context with the library."""
    ip = lambda a: np.ascontiguousarray(a, dtype=np.int32).ctypes.data_as(ctypes.POINTER(ctypes.c_int32))
    if halo.rows:
        sr, rr = halo.send_rows, halo.recv_rows
        rc = lib.sdg_set_halo_rows(S.h, etype, int(sr.shape[0]), ip(sr[:, 0]), ip(sr[:, 1]), int(rr.shape[0]), ip(rr[:, 0]), ip(rr[:, 1]))
    else:
        rc = lib.sdg_set_halo_send(S.h, etype, int(halo.n_send), ip(halo.send_elems))
    if rc != 0:
        raise RuntimeError(lib.sdg_last_error().decode())


# ---- the multi-GPU solver ----------------------------------------------------------------------------------------------------
class _DevArray:
    """A raw device range seen through __cuda_array_interface__ (torch.as_tensor aliases it without copying)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class DistributedSolver:
    """SubrosaDG::Solver<SC> (src/Solver/SolveControl.cpp:327-436) over one process per GPU.

    Each rank builds the library context for its element block (`Solver` with ghost elements) and drives the stages with
    the exchange overlapped: pack + NCCL on a communication stream, interior thread blocks on the library's stream.
    """

    def __init__(self, cfg: dict, mesh: M.Mesh, device: int | None = None, group=None, reorder: int = 1):
        import torch
        import torch.distributed as dist
        from .solver import Solver, load_library
        self.torch, self.dist = torch, dist
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.group = group
        self.device = int(os.environ.get("LOCAL_RANK", self.rank)) if device is None else device
        torch.cuda.set_device(self.device)
        self.part = partition(mesh, self.rank, self.world)
        self.etype = self.part.etype
        self.av = cfg.get("av_tolerance") is not None
        self.S = Solver(cfg, self.part.mesh, device=self.device, n_ghost={self.etype: self.part.n_ghost}, reorder=reorder,
                        node_data=partition_node_data(mesh, self.part) if self.av else None)
        self.lib = load_library()
        # kernels that read published face traces (P3 hexahedra: Navier-Stokes, Euler through traces) exchange 640-byte trace rows
        self.rows = bool(self.lib.sdg_uses_trace_rows(self.S.h)) and os.environ.get("SDG_HALO_ROWS", "1") != "0"
        self.halo = HaloExchange(self.part, group, rows=self.rows)
        self.Nv = self.S.Nv
        self.n_global = self.part.n_elements_global
        sz = self.S.sizes(self.etype)
        self.sizes = sz
        self.elem_doubles = sz.Nv * sz.Nb
        self.n_pass = int(self.lib.sdg_num_passes(self.S.h))
        self.n_stage = int(self.lib.sdg_num_stages(self.S.h))
        _set_halo(self.lib, self.S, self.etype, self.halo)
        self.main = torch.cuda.ExternalStream(int(self.lib.sdg_stream(self.S.h)), device=self.device)
        self.comm = torch.cuda.Stream(device=self.device, priority=-1)   # pack + NCCL get SM slots ahead of the queued interior blocks
        self.ev_ready = torch.cuda.Event()
        self.ev_halo = torch.cuda.Event()
        self._views = {}
        # halo transport: "ipc" = peer-memory stores over NVLink (CUDA IPC, no staging, no send/recv), "nccl" = pack + ncclSend/Recv
        self.transport = os.environ.get("SDG_HALO", "ipc" if (self.world > 1 and dist.get_backend(group) == "nccl") else "nccl")
        if self.transport == "ipc":
            # every rank must end up on the same transport: agree on the outcome of the IPC set-up, fall back to NCCL together
            try:
                self._connect_peers()
                ok = 1
            except Exception as exc:   # e.g. no peer access between two GPUs
                ok = 0
                print(f"[subrosadg_b200] rank {self.rank}: peer-memory halo set-up failed ({exc}); using NCCL send/recv", file=__import__("sys").stderr)
            flag = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{self.device}")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            if int(flag.item()) == 0:
                self.transport = "nccl"
        self.relative_error_ = np.zeros(self.Nv)
        self.delta_time_ = 0.0
        self.launches_extra = 0

    # -- plumbing ---------------------------------------------------------------------------------------------------------
    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.sdg_last_error().decode())

    def _view(self, ptr, n):
        key = (ptr, n)
        v = self._views.get(key)
        if v is None:
            v = self.torch.as_tensor(_DevArray(ptr, n), device=f"cuda:{self.device}") if n else self.torch.empty(0, dtype=self.torch.float64, device=f"cuda:{self.device}")
            self._views[key] = v
        return v

    def _buffers(self, what):
        sp, rp = ctypes.c_void_p(), ctypes.c_void_p()
        sn, rn = ctypes.c_int64(), ctypes.c_int64()
        self._chk(self.lib.sdg_halo_buffers_device(self.S.h, self.etype, what, ctypes.byref(sp), ctypes.byref(sn), ctypes.byref(rp), ctypes.byref(rn)))
        return self._view(sp.value or 0, sn.value), self._view(rp.value or 0, rn.value)

    def _connect_peers(self):
        """CUDA IPC set-up of the peer-memory exchange: every rank publishes the handles of its state / gradient / flag
        allocations and where it expects each peer's elements; the library opens the peers' handles (sdg_ipc_connect)."""
        lib, part = self.lib, self.part
        mine = np.zeros(5 * 64, dtype=np.uint8)
        self._chk(lib.sdg_ipc_export(self.S.h, mine.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte))))
        info = dict(handles=mine.tobytes(), n_owned=int(part.n_owned), peers=[int(q) for q in part.peers],
                    recv={int(q): (int(part.recv_range[q][0]), int(part.recv_range[q][1])) for q in part.peers},
                    recv_rows={int(q): part.recv_rows[q].tolist() for q in part.peers} if self.rows else {})
        everyone = [None] * self.world
        self.dist.all_gather_object(everyone, info, group=self.group)
        peers = [int(q) for q in part.peers]
        n = len(peers)
        handles = np.zeros(max(n, 1) * 5 * 64, dtype=np.uint8)
        ghost_first = np.zeros(max(n, 1), dtype=np.int64)
        send_first = np.zeros(max(n, 1), dtype=np.int32); send_count = np.zeros(max(n, 1), dtype=np.int32); slot = np.zeros(max(n, 1), dtype=np.int32)
        dst_units = np.zeros(max(int(self.halo.n_send), 1), dtype=np.int64)
        for k, q in enumerate(peers):
            other = everyone[q]
            handles[k * 320:(k + 1) * 320] = np.frombuffer(other["handles"], dtype=np.uint8)
            ns = self.halo.send_count[q]
            if self.rows:
                rows = np.asarray(other["recv_rows"].get(self.rank, np.zeros((0, 2), dtype=np.int64)), dtype=np.int64).reshape(-1, 2)
                if rows.shape[0] != ns:
                    raise RuntimeError(f"halo mismatch: rank {self.rank} sends {ns} rows to {q}, which expects {rows.shape[0]}")
                dst_units[self.halo.send_off[q]:self.halo.send_off[q] + ns] = rows[:, 0] * 6 + rows[:, 1]   # ghost rows keep their caller index
            else:
                r0, nr = other["recv"].get(self.rank, (0, 0))
                if nr != ns:
                    raise RuntimeError(f"halo mismatch: rank {self.rank} sends {ns} elements to {q}, which expects {nr}")
                ghost_first[k] = other["n_owned"] + r0
            send_first[k] = self.halo.send_off[q]; send_count[k] = ns
            slot[k] = other["peers"].index(self.rank)
        self._chk(lib.sdg_ipc_connect(self.S.h, n, handles.ctypes.data_as(ctypes.POINTER(ctypes.c_ubyte)),
                                      ghost_first.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), send_first.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                      send_count.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), slot.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))))
        if self.rows:
            self._chk(lib.sdg_ipc_set_destination_units(self.S.h, int(self.halo.n_send), dst_units.ctypes.data_as(ctypes.POINTER(ctypes.c_int64))))

    def _exchange(self, what):
        """Refresh the ghost copies of field `what` (0 state, 1 volume gradient); returns after ENQUEUEING the work on the
        communication stream and recording ev_halo there."""
        torch = self.torch
        self.ev_ready.record(self.main)
        self.comm.wait_event(self.ev_ready)
        if self.transport == "ipc":
            cs = ctypes.c_void_p(self.comm.cuda_stream)
            self._chk(self.lib.sdg_halo_push(self.S.h, self.etype, what, cs))   # stores into the peers' ghost ranges + arrival flags
            self._chk(self.lib.sdg_halo_wait(self.S.h, cs))                     # the peers' pushes into OUR ghost ranges
            self.ev_halo.record(self.comm)
            return
        with torch.cuda.stream(self.comm):
            self._chk(self.lib.sdg_halo_pack(self.S.h, self.etype, what, ctypes.c_void_p(self.comm.cuda_stream)))
            send, recv = self._buffers(what)
            per_elem = int(self.lib.sdg_halo_doubles_per_element(self.S.h, what))
            reqs = self.halo.start(send, recv, per_elem)
            self.halo.finish(reqs)      # stream-ordered for NCCL: the communication stream waits, the host does not
            self._chk(self.lib.sdg_halo_unpack(self.S.h, self.etype, what, ctypes.c_void_p(self.comm.cuda_stream)))   # trace rows: staging -> rows
            self.ev_halo.record(self.comm)

    def _reduce_node_viscosity(self):
        """The node maximum of Solver::calculateArtificialViscosity across the partitions (the cwiseMax combine of SpatialDiscrete.cpp:89-108):
        all-reduce(max) of the device node array on the library's stream, then the corner values of owned and ghost elements."""
        ptr, cnt = ctypes.c_void_p(), ctypes.c_int64()
        self._chk(self.lib.sdg_av_node_buffer(self.S.h, ctypes.byref(ptr), ctypes.byref(cnt)))
        nodes = self._view(ptr.value, cnt.value)
        with self.torch.cuda.stream(self.main):
            self.dist.all_reduce(nodes, op=self.dist.ReduceOp.MAX, group=self.group)
        self._chk(self.lib.sdg_av_store(self.S.h, ctypes.c_void_p(self.main.cuda_stream)))

    # -- Solver interface ---------------------------------------------------------------------------------------------------
    def initializeSolver(self, ic, bc=None):
        self.S.initializeSolver(ic, bc)

    def calculateDeltaTime(self, cfl) -> float:
        """Solver::calculateDeltaTime (TimeIntegration.cpp:104-179): global minimum over all ranks."""
        torch = self.torch
        dt = torch.tensor([self.S.calculateDeltaTime(cfl)], dtype=torch.float64, device=f"cuda:{self.device}")
        self.dist.all_reduce(dt, op=self.dist.ReduceOp.MIN, group=self.group)
        self.delta_time_ = float(dt.item())
        return self.delta_time_

    def stepSolver(self, dt=None, nsteps=1, want_error=True):
        """Solver::stepSolver x nsteps (TimeIntegration.cpp:326-350) with the halo exchange overlapped."""
        dt = self.delta_time_ if dt is None else dt
        lib, h = self.lib, self.S.h
        main = ctypes.c_void_p(self.main.cuda_stream)
        sums = np.zeros(8)
        # Every pass runs its ghost-reading thread blocks FIRST; as soon as they are done the exchange the NEXT pass needs is
        # started on the (high-priority) communication stream and runs under the long interior launch of the current pass.
        # Pass 0 needs the state U, pass 1 (NS) additionally the volume gradient.
        self._exchange(0)                                               # prime: ghosts of the current state
        for it in range(nsteps):
            self._chk(lib.sdg_step_begin(h, ctypes.c_double(dt)))
            if self.av:
                self._reduce_node_viscosity()
            for s in range(self.n_stage):
                for p in range(self.n_pass):
                    self.main.wait_event(self.ev_halo)
                    self._chk(lib.sdg_stage_pass(h, s, p, 1, main))     # thread blocks that read ghost elements
                    final = it == nsteps - 1 and s == self.n_stage - 1 and p == self.n_pass - 1
                    if not final:
                        self._exchange(p + 1 if p + 1 < self.n_pass else 0)
                    self._chk(lib.sdg_stage_pass(h, s, p, 0, main))     # thread blocks without ghost neighbours: overlap
            last = it == nsteps - 1
            self._chk(lib.sdg_step_end(h, sums.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if (last and want_error) else None))
        if want_error:
            t = self.torch.tensor(sums[:self.Nv], dtype=self.torch.float64, device=f"cuda:{self.device}")
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
            self.relative_error_ = t.cpu().numpy() / self.n_global   # TimeIntegration.cpp:323
        return self.relative_error_

    def synchronize(self):
        self.S.synchronize()
        self.comm.synchronize()

    # owned part of the fields, caller (global) element order within the block
    def get_state(self):
        return self.S.get_state(self.etype)[:self.part.n_owned]

    def state_at_quadrature(self):
        return self.S.state_at_quadrature(self.etype)[:self.part.n_owned]

    def gather_state_at_quadrature(self):
        """Global field on every rank (tests / output steps only)."""
        loc = self.torch.from_numpy(np.ascontiguousarray(self.state_at_quadrature())).to(f"cuda:{self.device}")
        counts = np.diff(block_bounds(self.n_global, self.world))
        outs = [self.torch.empty((int(c),) + tuple(loc.shape[1:]), dtype=loc.dtype, device=loc.device) for c in counts]
        self.dist.all_gather(outs, loc, group=self.group)
        return self.torch.cat(outs).cpu().numpy()

    @property
    def launch_count(self):
        return self.S.launch_count


class InProcessCluster:
    """`world` element-block contexts on ONE device, advanced in lock step with device-to-device copies in place of NCCL.
    Same partition, pack kernel, ghost ranges and part-0 / part-1 launches as DistributedSolver; used by the single-GPU
    parity tests of the multi-rank path (the result must equal the single-context run element for element)."""

    def __init__(self, cfg: dict, mesh: M.Mesh, world: int, device: int = 0, reorder: int = 1):
        import torch
        from .solver import Solver, load_library
        self.torch = torch
        self.lib = load_library()
        self.world = world
        self.device = device
        self.parts = [partition(mesh, r, world) for r in range(world)]
        self.etype = self.parts[0].etype
        self.av = cfg.get("av_tolerance") is not None
        self.S = [Solver(cfg, p.mesh, device=device, n_ghost={self.etype: p.n_ghost}, reorder=reorder, node_data=partition_node_data(mesh, p) if self.av else None)
                  for p in self.parts]
        self.rows = bool(self.lib.sdg_uses_trace_rows(self.S[0].h)) and os.environ.get("SDG_HALO_ROWS", "1") != "0"
        self.halos = [HaloExchange(p, rows=self.rows) for p in self.parts]
        for S, h in zip(self.S, self.halos):
            _set_halo(self.lib, S, self.etype, h)
        sz = self.S[0].sizes(self.etype)
        self.elem_doubles = sz.Nv * sz.Nb
        self.dim = self.S[0].dim
        self.n_pass = int(self.lib.sdg_num_passes(self.S[0].h))
        self.n_stage = int(self.lib.sdg_num_stages(self.S[0].h))
        self.n_global = self.parts[0].n_elements_global
        self.Nv = self.S[0].Nv

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.sdg_last_error().decode())

    def _buffers(self, r, what):
        sp, rp = ctypes.c_void_p(), ctypes.c_void_p()
        sn, rn = ctypes.c_int64(), ctypes.c_int64()
        self._chk(self.lib.sdg_halo_buffers_device(self.S[r].h, self.etype, what, ctypes.byref(sp), ctypes.byref(sn), ctypes.byref(rp), ctypes.byref(rn)))
        dev = f"cuda:{self.device}"
        mk = lambda p, n: self.torch.as_tensor(_DevArray(p, n), device=dev) if n else self.torch.empty(0, dtype=self.torch.float64, device=dev)
        return mk(sp.value or 0, sn.value), mk(rp.value or 0, rn.value)

    def _exchange(self, what):
        per = int(self.lib.sdg_halo_doubles_per_element(self.S[0].h, what))
        for S in self.S:
            self._chk(self.lib.sdg_halo_pack(S.h, self.etype, what, None))
            S.synchronize()
        bufs = [self._buffers(r, what) for r in range(self.world)]
        for r, p in enumerate(self.parts):
            for q in p.peers:
                r0, nr = self.halos[r].recv_off[q], self.halos[r].recv_count[q]
                if nr:
                    s0 = self.halos[q].send_off[r]
                    assert self.halos[q].send_count[r] == nr
                    bufs[r][1][r0 * per:(r0 + nr) * per].copy_(bufs[q][0][s0 * per:(s0 + nr) * per])
        self.torch.cuda.synchronize()
        for S in self.S:
            self._chk(self.lib.sdg_halo_unpack(S.h, self.etype, what, None))
            S.synchronize()

    def initializeSolver(self, ic, bc=None):
        for S in self.S:
            S.initializeSolver(ic, bc)

    def calculateDeltaTime(self, cfl):
        return min(S.calculateDeltaTime(cfl) for S in self.S)

    def stepSolver(self, dt, nsteps=1):
        sums = np.zeros((self.world, 8))
        for it in range(nsteps):
            for S in self.S:
                self._chk(self.lib.sdg_step_begin(S.h, ctypes.c_double(dt)))
            if self.av:      # the all-reduce(max) of DistributedSolver._reduce_node_viscosity, between contexts of one device
                arrays = []
                for S in self.S:
                    S.synchronize()
                    ptr, cnt = ctypes.c_void_p(), ctypes.c_int64()
                    self._chk(self.lib.sdg_av_node_buffer(S.h, ctypes.byref(ptr), ctypes.byref(cnt)))
                    arrays.append(self.torch.as_tensor(_DevArray(ptr.value, cnt.value), device=f"cuda:{self.device}"))
                top = self.torch.stack(arrays).max(dim=0).values
                for a in arrays:
                    a.copy_(top)
                self.torch.cuda.synchronize()
                for S in self.S:
                    self._chk(self.lib.sdg_av_store(S.h, None))
            for s in range(self.n_stage):
                for p in range(self.n_pass):
                    self._exchange(p)
                    for S in self.S:
                        self._chk(self.lib.sdg_stage_pass(S.h, s, p, 0, None))
                        self._chk(self.lib.sdg_stage_pass(S.h, s, p, 1, None))
            for r, S in enumerate(self.S):
                self._chk(self.lib.sdg_step_end(S.h, sums[r].ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if it == nsteps - 1 else None))
        return sums.sum(axis=0)[:self.Nv] / self.n_global

    def state_at_quadrature(self):
        return np.concatenate([S.state_at_quadrature(self.etype)[:p.n_owned] for S, p in zip(self.S, self.parts)])


# ---- bench.py entry for N > 1 ---------------------------------------------------------------------------------------------------
def bench_main(a, workload, metric, unit, bytes_per_dof, peaks, ClockSampler, ic, cfg_base, ns_cfg=None, ns_bytes_per_dof=144.0):
    """One rank of `torchrun ... bench.py --gpus N`: strong scaling of the same global mesh over N GPUs."""
    import torch
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    # stdout carries exactly one JSON line: NCCL prints its version banner (NCCL_DEBUG=VERSION/INFO) on file descriptor 1 when a
    # communicator is created, so fd 1 points at stderr until the communicators exist (end of the warm-up steps)
    import sys
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    os.environ.setdefault("TORCH_NCCL_HIGH_PRIORITY", "1")   # NCCL's internal stream: ahead of the interior thread blocks
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    mesh = M.periodic_box_fast(3, a.cells)
    cfg = dict(cfg_base); cfg["p"] = a.p
    D = DistributedSolver(cfg, mesh, device=local)
    D.initializeSolver(ic)
    sz = D.sizes
    dof_global = D.n_global * sz.Nb * sz.Nv
    nst = D.n_stage
    dt = D.calculateDeltaTime(1.0)
    warmup = max(a.warmup, 3)
    D.stepSolver(dt, warmup)
    D.synchronize(); sys.stdout.flush()
    os.dup2(saved_stdout, 1); os.close(saved_stdout)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    l0 = D.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.synchronize(); torch.cuda.synchronize(); dist.barrier()
    torch.cuda.synchronize()
    e0.record(D.main)
    D.stepSolver(dt, a.steps, want_error=False)
    e1.record(D.main)
    D.synchronize(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=f"cuda:{local}")
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    launches = D.launch_count - l0
    err = D.stepSolver(dt, 1)   # relative_error_ of one more step (outside the timed region)
    ck = clocks.stop() if rank == 0 else None
    # end to end through host buffers on every rank
    e2e = None
    if not a.no_e2e:
        t = D.etype
        U = torch.empty((sz.n, sz.Nb, sz.Nv), dtype=torch.float64).pin_memory()
        Un = U.numpy(); Un[...] = D.S.get_state(t)
        n_e2e = max(1, min(a.steps, 3))
        D.S.set_state(t, Un); D.stepSolver(dt, 1); D.S.get_state(t, out=Un)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            D.S.set_state(t, Un)
            D.stepSolver(dt, 1)
            D.S.get_state(t, out=Un)
        torch.cuda.synchronize(); dist.barrier()
        sec_e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(sec_e, op=dist.ReduceOp.MAX)
        own = D.part.n_owned * sz.Nb * sz.Nv
        e2e = {"value": dof_global * nst * n_e2e / float(sec_e.item()) / 1e9, "unit": unit, "h2d_bytes_per_step": int(sz.n * sz.Nb * sz.Nv * 8) * world,
               "d2h_bytes_per_step": int(sz.n * sz.Nb * sz.Nv * 8 + 8 * sz.Nv) * world, "steps": n_e2e,
               "note": f"every rank: host modal coefficients of its block ({own * 8 / 1e9:.2f} GB owned + ghosts) -> sdg_set_state -> stages with halo exchange -> sdg_get_state"}
    if rank == 0:
        hbm, how = peaks()
        sec = ms * 1e-3
        value = dof_global * nst * a.steps / sec / 1e9
        stage_ms = ms / (a.steps * nst)
        achieved = bytes_per_dof * dof_global / world / (stage_ms * 1e-3) / 1e9
        halo_bytes = int(D.halo.n_send) * int(D.lib.sdg_halo_doubles_per_element(D.S.h, 0)) * 8
        out = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": a.steps, "warmup": warmup, "ms_per_step": ms / a.steps,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
               "config": {"workload": workload, "elements": D.n_global, "scalar_dof": dof_global, "dt": dt,
                          "partition": f"contiguous element-index blocks (z-slabs), {D.part.n_owned} owned + {D.part.n_ghost} ghost elements on rank 0",
                          "halo": f"{halo_bytes / 1e6:.1f} MB per rank per stage pass, " + ("peer-memory stores into the neighbours' ghost ranges over NVLink (CUDA IPC)" if D.transport == "ipc" else "NCCL send/recv") + ", overlapped with the interior thread blocks",
                          "l2": "per-rank state much larger than the 126 MB L2 (no flush needed)", "relative_error": [float(x) for x in err]},
               "gpu_launches": int(launches) * world, "clocks": ck,
               "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm, "traffic": None,
                            "kernel": "stage kernels of one rank (interior + ghost-adjacent launches)", "kernel_ms": stage_ms,
                            "algorithmic_bytes_per_launch": bytes_per_dof * dof_global / world, "peak_source": how, "note": "per GPU"}}
        if e2e:
            out["e2e"] = e2e
    ns_line = None
    if ns_cfg is not None:
        # north_star's Navier-Stokes target cube (96^3 P3 hexahedra, BR2) on the same N GPUs: strong scaling of the NS stage, same timing rules
        D.S.close(); del D
        torch.cuda.empty_cache()
        cfg2 = dict(ns_cfg); cfg2["p"] = a.p
        D2 = DistributedSolver(cfg2, M.periodic_box_fast(3, 96), device=local)
        D2.initializeSolver(ic)
        dt2 = D2.calculateDeltaTime(1.0)
        D2.stepSolver(dt2, 3)
        steps2 = max(3, min(a.steps, 5))
        D2.synchronize(); torch.cuda.synchronize(); dist.barrier()
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(D2.main)
        D2.stepSolver(dt2, steps2, want_error=False)
        f1.record(D2.main)
        D2.synchronize(); torch.cuda.synchronize()
        ms2 = torch.tensor([f0.elapsed_time(f1)], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
        ms2 = float(ms2.item())
        dof2 = D2.n_global * D2.sizes.Nb * D2.sizes.Nv
        stage2 = ms2 / (steps2 * D2.n_stage)
        ach2 = ns_bytes_per_dof * dof2 / world / (stage2 * 1e-3) / 1e9
        ns_line = {"workload": f"periodic cube of configs[3] with CompresibleNS, HLLC, BR2, constant mu=1.4e-3, 96^3 hexes, p={a.p}, SSPRK3, {world} GPUs (strong scaling)",
                   "value": dof2 * D2.n_stage * steps2 / (ms2 * 1e-3) / 1e9, "unit": unit, "ms_per_stage": stage2,
                   "roofline": {"bound": "hbm", "achieved": ach2, "peak": peaks()[0], "unit": "GB/s", "frac": ach2 / peaks()[0], "note": "per GPU"},
                   "halo": f"{int(D2.halo.n_send) * int(D2.lib.sdg_halo_doubles_per_element(D2.S.h, 0)) * 8 / 1e6:.1f} MB per rank per stage pass"}
    if rank == 0:
        if ns_line:
            out["ns_target"] = ns_line
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
