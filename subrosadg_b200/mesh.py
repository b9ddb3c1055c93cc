"""In-code mesh producers (inputs of the hot path) — the stand-in for the Gmsh scripts in SubrosaDG's examples.

The reference obtains its meshes from the Gmsh library (`generateMesh` in every examples/*.cpp, then
src/Mesh/ReadControl.cpp:286-335 / src/Mesh/Adjacency.cpp:432-495).  Gmsh is not available here, so this module
produces the same *kind* of data deterministically in numpy:

* element blocks per element type with node coordinates in **gmsh node order** (corner, edge, face, volume nodes;
  `gmsh_reference_nodes`), the layout of `PerElementMesh::node_coordinate_` (src/Mesh/ReadControl.cpp:86-92);
* face records with the reference's index semantics (src/Mesh/ReadControl.cpp:72-83, src/Mesh/Adjacency.cpp:330-430):
  `parent_index_each_type_`, `adjacency_sequence_in_parent_`, parent element type, `adjacency_right_rotation_`,
  `boundary_condition_type_`, `gmsh_physical_index_`; interior faces first, boundary faces after.

Face order is canonical (order of first encounter while looping element types tri -> quad / hex, elements, local
faces) because the reference's own order is `std::unordered_map` iteration order (Adjacency.cpp:479-485) and therefore
not portable.  "Left" parent = first encounter; for periodic pairs this coincides with the reference's
"master face keeps its own parent as left" rule (Adjacency.cpp:242-253) when the masters are the low faces, which
is how every periodic example is set up (examples/periodic_2d_ceuler.cpp:85-86).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

POINT, LINE, TRIANGLE, QUADRANGLE, TETRAHEDRON, PYRAMID, HEXAHEDRON = range(7)  # ElementEnum, src/Utils/Enum.cpp:28-36

# BoundaryConditionEnum, src/Utils/Enum.cpp:92-100
RIEMANN_FARFIELD, VELOCITY_INFLOW, PRESSURE_OUTFLOW, ISOTHERMAL_NONSLIP_WALL, ADIABATIC_SLIP_WALL, ADIABATIC_NONSLIP_WALL, PERIODIC = range(7)

ELEM_DIM = {POINT: 0, LINE: 1, TRIANGLE: 2, QUADRANGLE: 2, HEXAHEDRON: 3}
FACE_TYPE = {LINE: POINT, TRIANGLE: LINE, QUADRANGLE: LINE, HEXAHEDRON: QUADRANGLE}

# getElementPerAdjacencyNodeIndex, src/Solver/SimulationControl.cpp:177-216
FACE_CORNERS = {
    LINE: np.array([[0], [1]]),
    TRIANGLE: np.array([[0, 1], [1, 2], [2, 0]]),
    QUADRANGLE: np.array([[0, 1], [1, 2], [2, 3], [3, 0]]),
    HEXAHEDRON: np.array([[0, 3, 2, 1], [0, 1, 5, 4], [0, 4, 7, 3], [1, 2, 6, 5], [2, 3, 7, 6], [4, 5, 6, 7]]),
}
CORNER_COORDS = {
    LINE: np.array([[-1.0, 0, 0], [1, 0, 0]]),
    TRIANGLE: np.array([[0.0, 0, 0], [1, 0, 0], [0, 1, 0]]),
    QUADRANGLE: np.array([[-1.0, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]]),
    HEXAHEDRON: np.array([[-1.0, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]]),
}
HEX_EDGES = [(0, 1), (0, 3), (0, 4), (1, 2), (1, 5), (2, 3), (2, 6), (3, 7), (4, 5), (4, 7), (5, 6), (6, 7)]


# ---- gmsh high-order node numbering (pinned by SimulationControl.cpp:525-887 in tests) ---------------------------------
def _line_interior(a, b, g):
    return [a + (b - a) * (i / g) for i in range(1, g)]


def _quad_nodes(c, g):
    if g == 0:
        return [sum(c) / 4.0]
    out = list(c)
    if g == 1:
        return out
    for e in range(4):
        out += _line_interior(c[e], c[(e + 1) % 4], g)

    def bil(s, t):
        return (1 - s) * (1 - t) * c[0] + s * (1 - t) * c[1] + s * t * c[2] + (1 - s) * t * c[3]

    h = 1.0 / g
    return out + _quad_nodes([bil(h, h), bil(1 - h, h), bil(1 - h, 1 - h), bil(h, 1 - h)], g - 2)


def _tri_nodes(c, g):
    if g == 0:
        return [sum(c) / 3.0]
    out = list(c)
    if g == 1:
        return out
    for e in range(3):
        out += _line_interior(c[e], c[(e + 1) % 3], g)
    if g < 3:
        return out
    h = 1.0 / g

    def bary(l0, l1, l2):
        return l0 * c[0] + l1 * c[1] + l2 * c[2]

    return out + _tri_nodes([bary(1 - 2 * h, h, h), bary(h, 1 - 2 * h, h), bary(h, h, 1 - 2 * h)], g - 3)


def _hex_nodes(c, g):
    if g == 0:
        return [sum(c) / 8.0]
    out = list(c)
    if g == 1:
        return out
    for a, b in HEX_EDGES:
        out += _line_interior(c[a], c[b], g)
    for f in range(6):
        fc = [c[i] for i in FACE_CORNERS[HEXAHEDRON][f]]
        out += _quad_nodes(fc, g)[4 + 4 * (g - 1):]

    def tri(s, t, u):
        lo = (1 - s) * (1 - t) * c[0] + s * (1 - t) * c[1] + s * t * c[2] + (1 - s) * t * c[3]
        hi = (1 - s) * (1 - t) * c[4] + s * (1 - t) * c[5] + s * t * c[6] + (1 - s) * t * c[7]
        return (1 - u) * lo + u * hi

    h = 1.0 / g
    ci = [tri(h, h, h), tri(1 - h, h, h), tri(1 - h, 1 - h, h), tri(h, 1 - h, h),
          tri(h, h, 1 - h), tri(1 - h, h, 1 - h), tri(1 - h, 1 - h, 1 - h), tri(h, 1 - h, 1 - h)]
    return out + _hex_nodes(ci, g - 2)


def gmsh_reference_nodes(etype: int, order: int) -> np.ndarray:
    """Reference coordinates (n x 3) of the Lagrange nodes of `order` in gmsh numbering."""
    c = [np.array(x, dtype=float) for x in CORNER_COORDS[etype]] if etype != POINT else None
    if etype == POINT:
        return np.zeros((1, 3))
    if etype == LINE:
        return np.array(c + _line_interior(c[0], c[1], order))
    if etype == TRIANGLE:
        return np.array(_tri_nodes(c, order))
    if etype == QUADRANGLE:
        return np.array(_quad_nodes(c, order))
    if etype == HEXAHEDRON:
        return np.array(_hex_nodes(c, order))
    raise ValueError("unsupported element type")


@dataclass
class Mesh:
    """Flat mesh description handed to the C ABI (and to the oracle)."""
    dim: int
    blocks: dict = field(default_factory=dict)   # type -> {"coords": [n, nn, D] gmsh order, "geom_order": g, "corners": [n, nc]}
    faces: dict = field(default_factory=dict)    # n_int, n_bnd, le, lt, lf, re, rt, rf, rot, bc, phys (int32 arrays)
    phys_bc: dict = field(default_factory=dict)  # gmsh physical index -> BoundaryConditionEnum value
    info: dict = field(default_factory=dict)

    @property
    def n_elements(self):
        return sum(b["coords"].shape[0] for b in self.blocks.values())


# ---- generic adjacency builder (restates the semantics of src/Mesh/Adjacency.cpp without gmsh) -------------------------
def build_faces(dim, corner_ids: dict, canon: np.ndarray, tagger, phys_bc: dict, face_centroid_fn):
    """corner_ids: type -> [n, nc] global corner node ids; canon: node id -> canonical id (periodic identification).

    Returns the face dict.  `tagger(centroids[nb, D]) -> phys[nb]`; `face_centroid_fn(type, elem, lface) -> [.., D]`.
    """
    recs_t, recs_e, recs_f, recs_c = [], [], [], []
    for t in sorted(corner_ids):  # tri (2) before quad (3): reference loops HasTriangle then HasQuadrangle (Adjacency.cpp:443-452)
        ids = corner_ids[t]
        fc = FACE_CORNERS[t]
        n, nf = ids.shape[0], fc.shape[0]
        c = canon[ids[:, fc]]  # [n, nf, nfc]
        recs_c.append(c.reshape(n * nf, -1))
        recs_t.append(np.full(n * nf, t, dtype=np.int64))
        recs_e.append(np.repeat(np.arange(n, dtype=np.int64), nf))
        recs_f.append(np.tile(np.arange(nf, dtype=np.int64), n))
    C = np.concatenate(recs_c)
    T = np.concatenate(recs_t); E = np.concatenate(recs_e); Fl = np.concatenate(recs_f)
    key = np.sort(C, axis=1)
    nrec = key.shape[0]
    order = np.lexsort(tuple(key[:, k] for k in range(key.shape[1] - 1, -1, -1)) if key.shape[1] > 1 else (key[:, 0],))
    ks = key[order]
    newgrp = np.ones(nrec, dtype=bool)
    newgrp[1:] = np.any(ks[1:] != ks[:-1], axis=1)
    gid_sorted = np.cumsum(newgrp) - 1
    ng = int(gid_sorted[-1]) + 1
    counts = np.bincount(gid_sorted, minlength=ng)
    if counts.max() > 2:
        raise ValueError("a face is shared by more than two elements (periodic direction with < 3 cells?)")
    # lexsort is stable => within a group the records are in encounter order
    first_pos = np.flatnonzero(newgrp)
    left_rec = order[first_pos]
    has_right = counts == 2
    right_rec = np.where(has_right, order[np.minimum(first_pos + 1, nrec - 1)], -1)
    # canonical order: by left record index (first encounter)
    g_order = np.argsort(left_rec, kind="stable")
    left_rec, right_rec, has_right = left_rec[g_order], right_rec[g_order], has_right[g_order]
    int_sel, bnd_sel = np.flatnonzero(has_right), np.flatnonzero(~has_right)
    L = np.concatenate([left_rec[int_sel], left_rec[bnd_sel]])
    R = np.concatenate([right_rec[int_sel], np.full(bnd_sel.size, -1)])
    n_int, n_bnd = int_sel.size, bnd_sel.size
    rot = np.zeros(n_int + n_bnd, dtype=np.int32)
    if dim == 3 and n_int:
        # adjacency_right_rotation_ = index, in the right parent's face-corner list, of the left's first corner
        # (Adjacency.cpp:166-171; periodic pairs: image of that corner)
        lc0 = C[L[:n_int], 0]
        rc = C[R[:n_int]]
        hit = rc == lc0[:, None]
        if not np.all(hit.sum(axis=1) == 1):
            raise ValueError("cannot determine face rotation")
        rot[:n_int] = np.argmax(hit, axis=1)
    faces = dict(n_int=n_int, n_bnd=n_bnd,
                 le=E[L].astype(np.int32), lt=T[L].astype(np.int32), lf=Fl[L].astype(np.int32),
                 re=np.where(R >= 0, E[R], -1).astype(np.int32), rt=np.where(R >= 0, T[R], -1).astype(np.int32),
                 rf=np.where(R >= 0, Fl[R], -1).astype(np.int32), rot=rot,
                 bc=np.full(n_int + n_bnd, PERIODIC, dtype=np.int32), phys=np.zeros(n_int + n_bnd, dtype=np.int32))
    if n_bnd:
        cent = np.zeros((n_bnd, dim))
        for t in sorted(corner_ids):
            sel = np.flatnonzero(faces["lt"][n_int:] == t)
            if sel.size:
                cent[sel] = face_centroid_fn(t, faces["le"][n_int:][sel], faces["lf"][n_int:][sel])
        phys = np.asarray(tagger(cent), dtype=np.int32)
        faces["phys"][n_int:] = phys
        faces["bc"][n_int:] = np.array([phys_bc[int(p)] for p in phys], dtype=np.int32)
    return faces


def _unique_ids(points: np.ndarray, tol=1e-9):
    """Global ids for coincident points ([m, D] -> [m] ids) by rounding to `tol`."""
    q = np.round(points / tol).astype(np.int64)
    _, inv = np.unique(q, axis=0, return_inverse=True)
    return inv.reshape(-1)


def mesh_from_blocks(dim, elems: dict, geom_order: int, phys_bc: dict, tagger=None, periodic=None, info=None, ident=None) -> Mesh:
    """elems: type -> [n, nn, D] node coordinates in gmsh order (order `geom_order`).

    periodic: list of (axis, lo, hi) — nodes on x[axis]==hi are identified with x[axis]==lo.
    """
    corner_xyz, counts = [], {}
    for t in sorted(elems):
        nc = CORNER_COORDS[t].shape[0]
        src = ident[t] if ident is not None else elems[t]  # `ident`: unwarped coordinates used for node identification
        corner_xyz.append(src[:, :nc, :].reshape(-1, dim))
        counts[t] = (elems[t].shape[0], nc)
    allc = np.concatenate(corner_xyz)
    ids = _unique_ids(allc)
    canon_xyz = allc.copy()
    if periodic:
        for axis, lo, hi in periodic:
            m = np.abs(canon_xyz[:, axis] - hi) < 1e-9 * max(1.0, abs(hi - lo))
            canon_xyz[m, axis] = lo
    cids = _unique_ids(canon_xyz)
    canon = np.zeros(ids.max() + 1, dtype=np.int64)
    canon[ids] = cids
    corner_ids, off = {}, 0
    for t in sorted(elems):
        n, nc = counts[t]
        corner_ids[t] = ids[off:off + n * nc].reshape(n, nc)
        off += n * nc

    def centroid(t, e, lf):
        nc = CORNER_COORDS[t].shape[0]
        src = ident[t] if ident is not None else elems[t]
        c = src[:, :nc, :][e[:, None], FACE_CORNERS[t][lf]]
        return c.mean(axis=1)

    faces = build_faces(dim, corner_ids, canon, tagger if tagger else (lambda c: np.ones(len(c), dtype=np.int32)), phys_bc, centroid)
    blocks = {t: dict(coords=np.ascontiguousarray(elems[t]), geom_order=geom_order, corners=corner_ids[t]) for t in elems}
    return Mesh(dim=dim, blocks=blocks, faces=faces, phys_bc=dict(phys_bc), info=info or {})


# ---- structured producers ------------------------------------------------------------------------------------------------
def _structured_elements(dim, shape, mapping, g):
    """Tensor block of prod(shape) quads/hexes (or lines); logical coords in [0,1]^dim; lexicographic numbering, x fastest.

    Returns [n, nn, dim] node coordinates in gmsh order.
    """
    etype = {1: LINE, 2: QUADRANGLE, 3: HEXAHEDRON}[dim]
    ref = gmsh_reference_nodes(etype, g)[:, :dim]  # [-1,1]^dim
    idx = np.indices(shape[::-1]).reshape(dim, -1)[::-1].T  # [n, dim] with axis 0 (x) fastest
    n = idx.shape[0]
    logical = (idx[:, None, :] + (ref[None, :, :] + 1.0) * 0.5) / np.asarray(shape, dtype=float)[None, None, :]
    return mapping(logical.reshape(-1, dim)).reshape(n, ref.shape[0], -1)


def box(dim, shape, lo, hi, periodic_axes=(), phys_bc=None, geom_order=1, warp=None, tagger=None) -> Mesh:
    """Structured [lo,hi]^dim box of lines/quads/hexes.  Non-periodic boundary faces get physical index
    1 + 2*axis + side (side 0 = low) unless `tagger` is given.  `warp(x)` optionally perturbs coordinates
    (use geom_order >= 2 for a genuinely curved mesh)."""
    lo = np.broadcast_to(np.asarray(lo, dtype=float), (dim,)); hi = np.broadcast_to(np.asarray(hi, dtype=float), (dim,))

    def mapping(s):
        x = lo + s * (hi - lo)
        return warp(x) if warp else x

    etype = {1: LINE, 2: QUADRANGLE, 3: HEXAHEDRON}[dim]
    elems = {etype: _structured_elements(dim, tuple(shape), mapping, geom_order)}
    ident = {etype: _structured_elements(dim, tuple(shape), lambda s: lo + s * (hi - lo), 1)} if warp else None
    periodic = [(a, lo[a], hi[a]) for a in periodic_axes]
    if phys_bc is None:
        phys_bc = {1 + 2 * a + s: RIEMANN_FARFIELD for a in range(dim) for s in range(2)}

    def default_tagger(c):
        out = np.zeros(len(c), dtype=np.int32)
        for a in range(dim):
            out[np.abs(c[:, a] - lo[a]) < 1e-9] = 1 + 2 * a
            out[np.abs(c[:, a] - hi[a]) < 1e-9] = 2 + 2 * a
        return out

    return mesh_from_blocks(dim, elems, geom_order, phys_bc, tagger or default_tagger, periodic, ident=ident,
                            info=dict(kind="box", shape=tuple(shape), lo=lo.tolist(), hi=hi.tolist(), periodic_axes=tuple(periodic_axes)))


def periodic_box(dim, n, lo=0.0, hi=2.0, geom_order=1, warp=None) -> Mesh:
    """Fully periodic [lo,hi]^dim box, n cells per direction (configs 1 and 4: examples/periodic_{2,3}d_ceuler.cpp)."""
    return box(dim, (n,) * dim, lo, hi, periodic_axes=tuple(range(dim)), phys_bc={1: PERIODIC}, geom_order=geom_order, warp=warp)


def periodic_box_fast(dim, n, lo=0.0, hi=2.0) -> Mesh:
    """Closed-form version of periodic_box (geom_order 1) for large n: identical arrays, no sorting."""
    shape = (n,) * dim
    etype = {1: LINE, 2: QUADRANGLE, 3: HEXAHEDRON}[dim]
    coords = _structured_elements(dim, shape, lambda s: lo + s * (hi - lo), 1)
    ne = n ** dim
    e = np.arange(ne, dtype=np.int64)
    ijk = [(e // n ** a) % n for a in range(dim)]
    # local face on the low / high side of each axis, SimulationControl.cpp:177-216 with the gmsh corner coordinates
    low_face = {1: [0], 2: [3, 0], 3: [2, 1, 0]}[dim]
    high_face = {1: [1], 2: [1, 2], 3: [3, 4, 5]}[dim]
    # rotation of the pair (left = low face of the element, right = high face of its low neighbour)
    rots = {1: [0], 2: [0, 0], 3: [0, 1, 0]}[dim]
    # first-encounter order: element e, local faces ascending; a face is first met by the lower-indexed parent,
    # except wrap-around faces which are first met by the element at index 0 along that axis (its LOW face).
    rec_key, le, lf, re, rf, rot = [], [], [], [], [], []
    for a in range(dim):
        nb_hi = e + np.where(ijk[a] == n - 1, -(n - 1), 1) * n ** a  # neighbour across the high face
        interior = ijk[a] < n - 1
        # interior faces: left = e (high face), right = nb_hi (low face)
        sel = np.flatnonzero(interior)
        le.append(sel); lf.append(np.full(sel.size, high_face[a])); re.append(nb_hi[sel]); rf.append(np.full(sel.size, low_face[a]))
        rec_key.append(sel * 8 + high_face[a])
        # rotation for left=high face / right=low face
        rot.append(np.full(sel.size, rots[a]))
        # wrap faces: left = element with index 0 along a (low face), right = element with index n-1 (high face)
        sel = np.flatnonzero(ijk[a] == 0)
        nb_lo = sel + (n - 1) * n ** a
        le.append(sel); lf.append(np.full(sel.size, low_face[a])); re.append(nb_lo); rf.append(np.full(sel.size, high_face[a]))
        rec_key.append(sel * 8 + low_face[a])
        rot.append(np.full(sel.size, rots[a]))
    key = np.concatenate(rec_key)
    order = np.argsort(key, kind="stable")
    cat = lambda xs: np.concatenate(xs)[order].astype(np.int32)
    nfa = key.size
    faces = dict(n_int=nfa, n_bnd=0, le=cat(le), lt=np.full(nfa, etype, dtype=np.int32), lf=cat(lf), re=cat(re),
                 rt=np.full(nfa, etype, dtype=np.int32), rf=cat(rf), rot=cat(rot),
                 bc=np.full(nfa, PERIODIC, dtype=np.int32), phys=np.zeros(nfa, dtype=np.int32))
    return Mesh(dim=dim, blocks={etype: dict(coords=coords, geom_order=1, corners=None)}, faces=faces, phys_bc={1: PERIODIC},
                info=dict(kind="box", shape=shape, lo=[lo] * dim, hi=[hi] * dim, periodic_axes=tuple(range(dim))))


def annulus(nr, nt, r0=0.5, r1=5.0, geom_order=3, stretch=1.0, tri_rings=0, phys_bc=None) -> Mesh:
    """O-mesh of curved quads between the circles r0 (wall, physical 2) and r1 (far field, physical 1); the outer
    `tri_rings` rings are split into triangles (hybrid tri/quad mesh as in examples/karmanvortex_2d_cns.cpp)."""
    phys_bc = phys_bc or {1: RIEMANN_FARFIELD, 2: ADIABATIC_SLIP_WALL}

    def mapping(s):
        t = s[:, 1] ** stretch if stretch != 1.0 else s[:, 1]
        r = r0 + (r1 - r0) * t
        th = -2.0 * np.pi * s[:, 0]  # clockwise so that (theta, r) is right-handed
        return np.stack([r * np.cos(th), r * np.sin(th)], axis=1)

    quads = _structured_elements(2, (nt, nr), mapping, geom_order)
    elems = {}
    nq_rings = nr - tri_rings
    ring = np.arange(nt * nr) // nt
    if tri_rings > 0:
        # triangles: straight-sided split of the (straight-sided image of the) outer quads; geometry order 1 there
        tq = _structured_elements(2, (nt, nr), mapping, 1)[ring >= nq_rings]
        t1 = tq[:, [0, 1, 2], :]; t2 = tq[:, [0, 2, 3], :]
        tris = np.empty((2 * tq.shape[0], 3, 2)); tris[0::2] = t1; tris[1::2] = t2
        if geom_order > 1:
            ref = gmsh_reference_nodes(TRIANGLE, geom_order)[:, :2]
            lam = np.stack([1 - ref[:, 0] - ref[:, 1], ref[:, 0], ref[:, 1]], axis=1)
            tris = np.einsum("nm,emd->end", lam, tris)
        elems[TRIANGLE] = tris
        qsel = quads[ring < nq_rings]
        if tri_rings < nr:
            # the quad ring adjacent to the triangles must share a STRAIGHT outer edge with them
            last = np.flatnonzero(ring[ring < nq_rings] == nq_rings - 1)
            qsel = qsel.copy()
            ref = gmsh_reference_nodes(QUADRANGLE, geom_order)[:, :2]
            top = np.flatnonzero(np.abs(ref[:, 1] - 1.0) < 1e-12)
            c2, c3 = qsel[last][:, 2, :], qsel[last][:, 3, :]
            s = (ref[top, 0] + 1.0) * 0.5  # 0 at corner 3 (xi=-1), 1 at corner 2
            qsel[np.ix_(last, top)] = c3[:, None, :] * (1 - s)[None, :, None] + c2[:, None, :] * s[None, :, None]
            elems[QUADRANGLE] = qsel
    else:
        elems[QUADRANGLE] = quads

    def tagger(c):
        r = np.hypot(c[:, 0], c[:, 1])
        return np.where(r < 0.5 * (r0 + r1), 2, 1).astype(np.int32)

    return mesh_from_blocks(2, elems, geom_order, phys_bc, tagger, None, info=dict(kind="annulus", r0=r0, r1=r1))


def naca0012(nr=19, nt=78, r_far=20.0, geom_order=3, phys_bc=None) -> Mesh:
    """O-mesh of curved P3 quads around a NACA0012 with closed trailing edge (config 2, examples/naca0012_2d_ceuler.cpp:
    far field = physical 1, slip wall = physical 2).  The reference uses a 4-block gmsh C-mesh (2204 cells); this
    in-code O-mesh has nr*nt cells of the same family (curved wall elements, far-field circle)."""
    phys_bc = phys_bc or {1: RIEMANN_FARFIELD, 2: ADIABATIC_SLIP_WALL}

    def wall(th):  # th in [0, 2pi): trailing edge -> upper -> leading edge -> lower -> trailing edge
        x = 0.5 * (1.0 + np.cos(th))
        yt = 0.6 * (0.2969 * np.sqrt(x) - 0.1260 * x - 0.3516 * x ** 2 + 0.2843 * x ** 3 - 0.1036 * x ** 4)
        return np.stack([x, np.where(np.sin(th) >= 0, yt, -yt)], axis=1)

    def mapping(s):
        th = -2.0 * np.pi * s[:, 0]  # clockwise so that (theta, r) is right-handed
        w = wall(th)
        far = np.stack([0.5 + r_far * np.cos(th), r_far * np.sin(th)], axis=1)
        t = (np.exp(4.0 * s[:, 1]) - 1.0) / (np.exp(4.0) - 1.0)
        return w + (far - w) * t[:, None]

    quads = _structured_elements(2, (nt, nr), mapping, geom_order)

    def tagger(c):
        r = np.hypot(c[:, 0] - 0.5, c[:, 1])
        return np.where(r < 0.5 * r_far, 2, 1).astype(np.int32)

    return mesh_from_blocks(2, {QUADRANGLE: quads}, geom_order, phys_bc, tagger, None, info=dict(kind="naca0012"))


def cubed_sphere_shell(n, nr, r0=0.5, r1=5.0, geom_order=3, stretch=2.0, phys_bc=None) -> Mesh:
    """Six-block curved hexahedral shell between the spheres r0 (wall, physical 2) and r1 (far field, physical 1):
    the in-code stand-in for the 6 sphere blocks + far-field blocks of examples/sphere_3d_cns.cpp:259-296."""
    phys_bc = phys_bc or {1: RIEMANN_FARFIELD, 2: ADIABATIC_NONSLIP_WALL}
    blocks = []
    # right-handed (a, b, outward) frames for the six cube faces
    frames = [((0, 1, 0), (0, 0, 1), (1, 0, 0)), ((0, 0, 1), (0, 1, 0), (-1, 0, 0)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)),
              ((1, 0, 0), (0, 0, 1), (0, -1, 0)), ((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((0, 1, 0), (1, 0, 0), (0, 0, -1))]
    for ea, eb, en in frames:
        ea, eb, en = np.array(ea, float), np.array(eb, float), np.array(en, float)

        def mapping(s, ea=ea, eb=eb, en=en):
            a = np.tan((2.0 * s[:, 0] - 1.0) * np.pi / 4.0)
            b = np.tan((2.0 * s[:, 1] - 1.0) * np.pi / 4.0)
            d = a[:, None] * ea + b[:, None] * eb + en
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            t = (np.exp(stretch * s[:, 2]) - 1.0) / (np.exp(stretch) - 1.0) if stretch else s[:, 2]
            return d * (r0 + (r1 - r0) * t)[:, None]

        blocks.append(_structured_elements(3, (n, n, nr), mapping, geom_order))
    hexes = np.concatenate(blocks)

    def tagger(c):
        r = np.linalg.norm(c, axis=1)
        return np.where(r < 0.5 * (r0 + r1), 2, 1).astype(np.int32)

    return mesh_from_blocks(3, {HEXAHEDRON: hexes}, geom_order, phys_bc, tagger, None, info=dict(kind="cubed_sphere_shell", r0=r0, r1=r1))


def _progression(n, ratio):
    """gmsh transfinite "Progression": n cells whose lengths grow by `ratio` from the start; returns the n+1 points in [0, 1]."""
    h = ratio ** np.arange(n, dtype=float)
    return np.concatenate([[0.0], np.cumsum(h)]) / h.sum()


def sphere_in_box(n_mid=11, n_out=9, n_rad=9, half=5.0, r_sphere=0.5, geom_order=3, phys_bc=None) -> Mesh:
    """Block topology of examples/sphere_3d_cns.cpp:72-296 (config 5): a 3x3x3 arrangement of transfinite blocks on [-half, half]^3
    whose centre block is replaced by SIX blocks between the sphere (wall, physical 2) and the faces of the centre cube, 26 + 6 blocks in
    total.  Cells per axis: n_out (progression 1.3 away from the centre, :259-274), n_mid, n_out; sphere blocks n_mid x n_mid x n_rad
    (radial progression 1.2, :286-292).  Defaults = the shipped mesh: 29^3 - 11^3 + 6 * 11 * 11 * 9 = 29,592 hexahedra.  The corners of the
    centre cube lie on the sphere of radius 1 (x = sqrt(3)/6, :73-78) and its edges are circle arcs (:108-113); gmsh's surface filling is
    not available here, so the curved faces are those of a transfinite blend of each block's edges (arcs on the centre cube, straight elsewhere).
    Far field = physical 1."""
    phys_bc = phys_bc or {1: RIEMANN_FARFIELD, 2: ADIABATIC_NONSLIP_WALL}
    a0 = 2.0 * np.sqrt(3.0) / 6.0          # half width of the centre cube; a0 * sqrt(3) = 1
    # 1-D point distributions of the three segments of an axis, as functions of the logical coordinate in [0, 1]
    seg_pts = [-half + (half - a0) * (1.0 - _progression(n_out, 1.3)[::-1]), np.linspace(-a0, a0, n_mid + 1), a0 + (half - a0) * _progression(n_out, 1.3)]
    seg_n = [n_out, n_mid, n_out]

    def axis_coord(seg, s):   # piecewise-linear in the cell index, so that element nodes of order g are equispaced inside a cell
        return np.interp(s * seg_n[seg], np.arange(seg_n[seg] + 1), seg_pts[seg])

    def morph(p):
        # transfinite blend per block: q = closest point of the centre cube; its displacement onto the unit sphere fades out linearly
        # along every axis that leaves the cube.  Face blocks: (1 - w) sphere patch + w flat far face; edge blocks: bilinear blend whose
        # inner edge is the circle arc; corner blocks: identity (the cube corners already lie on the sphere) — smooth inside each block.
        q = np.clip(p, -a0, a0)
        w = np.clip((np.abs(p) - a0) / (half - a0), 0.0, 1.0)
        fade = np.prod(1.0 - w, axis=1)
        sph = q / np.linalg.norm(q, axis=1, keepdims=True)
        return p + fade[:, None] * (sph - q)

    blocks = []
    for bk in range(3):
        for bj in range(3):
            for bi in range(3):
                if bi == 1 and bj == 1 and bk == 1:
                    continue

                def mapping(s, bi=bi, bj=bj, bk=bk):
                    p = np.stack([axis_coord(bi, s[:, 0]), axis_coord(bj, s[:, 1]), axis_coord(bk, s[:, 2])], axis=1)
                    return morph(p)

                blocks.append(_structured_elements(3, (seg_n[bi], seg_n[bj], seg_n[bk]), mapping, geom_order))
    frames = [((0, 1, 0), (0, 0, 1), (1, 0, 0)), ((0, 0, 1), (0, 1, 0), (-1, 0, 0)), ((0, 0, 1), (1, 0, 0), (0, 1, 0)),
              ((1, 0, 0), (0, 0, 1), (0, -1, 0)), ((1, 0, 0), (0, 1, 0), (0, 0, 1)), ((0, 1, 0), (1, 0, 0), (0, 0, -1))]
    rad_pts = r_sphere + (1.0 - r_sphere) * _progression(n_rad, 1.2)
    for ea, eb, en in frames:
        ea, eb, en = np.array(ea, float), np.array(eb, float), np.array(en, float)

        def mapping(s, ea=ea, eb=eb, en=en):
            a = a0 * (2.0 * s[:, 0] - 1.0); b = a0 * (2.0 * s[:, 1] - 1.0)
            q = a[:, None] * ea + b[:, None] * eb + a0 * en
            q /= np.linalg.norm(q, axis=1, keepdims=True)
            r = np.interp(s[:, 2] * n_rad, np.arange(n_rad + 1), rad_pts)
            return q * r[:, None]

        blocks.append(_structured_elements(3, (n_mid, n_mid, n_rad), mapping, geom_order))
    hexes = np.concatenate(blocks)

    def tagger(c):
        return np.where(np.linalg.norm(c, axis=1) < 0.5 * (r_sphere + half), 2, 1).astype(np.int32)

    return mesh_from_blocks(3, {HEXAHEDRON: hexes}, geom_order, phys_bc, tagger, None,
                            info=dict(kind="sphere_in_box", blocks=32, n_mid=n_mid, n_out=n_out, n_rad=n_rad))


# ---- mesh data of ShockCapturingEnum::ArtificialViscosity (PerElementMesh::node_tag_, inner_radius_; ReadControl.cpp:64,91) ------------
N_BASIC = {LINE: 2, TRIANGLE: 3, QUADRANGLE: 4, HEXAHEDRON: 8}   # kBasicNodeNumber: the corner nodes lead the gmsh node order


def node_tags(mesh: Mesh):
    """({type: int32 [n, kBasicNodeNumber] 0-based tags of the corner nodes}, node_number).  The flat mesh carries coordinates per
    element, not gmsh tags: nodes are the distinct coordinate tuples over all blocks (the two copies of a periodic pair stay distinct,
    as in a Gmsh mesh), numbered in lexicographic order of their coordinates — the same rule as MeshData::countNodes in the C++ shim."""
    pts = np.concatenate([np.asarray(b["coords"], dtype=np.float64).reshape(-1, mesh.dim) for b in mesh.blocks.values()])
    uniq, inv = np.unique(pts, axis=0, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    out, at = {}, 0
    for t, b in mesh.blocks.items():
        n, nn = np.asarray(b["coords"]).shape[:2]
        out[t] = np.ascontiguousarray(inv[at:at + n * nn].reshape(n, nn)[:, :N_BASIC[t]], dtype=np.int32)
        at += n * nn
    return out, int(uniq.shape[0])


def _quad_inner_radius(P):
    """P [n, 4, 2]: radius of the smallest circle tangent to three consecutive edge lines (centre = intersection of the interior angle
    bisectors at the two vertices of the middle edge) — Gmsh's MQuadrangle::getInnerRadius (third party, Gmsh 4.x; restated from its
    published algorithm, not checked against Gmsh here)."""
    unit = lambda v: v / np.linalg.norm(v, axis=-1, keepdims=True)
    cross = lambda a, b: a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]
    best = np.full(P.shape[0], np.inf)
    for i in range(4):
        A, B, prev, nxt = P[:, i], P[:, (i + 1) % 4], P[:, (i - 1) % 4], P[:, (i + 2) % 4]
        dA = unit(prev - A) + unit(B - A)
        dB = unit(A - B) + unit(nxt - B)
        den = cross(dA, dB)
        ok = np.abs(den) > 1e-300
        s_ = np.where(ok, cross(B - A, dB) / np.where(ok, den, 1.0), np.inf)
        C = A + s_[:, None] * dA
        r = np.abs(cross(B - A, C - A)) / np.linalg.norm(B - A, axis=-1)
        best = np.minimum(best, np.where(ok, r, np.inf))
    return best


def inner_radius(mesh: Mesh, t: int) -> np.ndarray:
    """inner_radius_ of every element of block t: gmsh::model::mesh::getElementQualities(..., "innerRadius") (Geometry.cpp:31-41).  Gmsh is
    a third-party dependency that is not in this image; its definitions are restated: line = half length, triangle = area / half
    perimeter (inscribed circle), quadrangle = see _quad_inner_radius, hexahedron = minimum of its six faces' quadrangle radii."""
    X = np.asarray(mesh.blocks[t]["coords"], dtype=np.float64)[:, :N_BASIC[t], :]
    if t == LINE:
        return 0.5 * np.linalg.norm(X[:, 1] - X[:, 0], axis=-1)
    if t == TRIANGLE:
        a = np.linalg.norm(X[:, 1] - X[:, 0], axis=-1); b = np.linalg.norm(X[:, 2] - X[:, 1], axis=-1); c = np.linalg.norm(X[:, 0] - X[:, 2], axis=-1)
        k = 0.5 * (a + b + c)
        return np.sqrt(k * (k - a) * (k - b) * (k - c)) / k
    if t == QUADRANGLE:
        return _quad_inner_radius(X[:, :, :2])
    if t == HEXAHEDRON:
        best = np.full(X.shape[0], np.inf)
        for fc in FACE_CORNERS[HEXAHEDRON]:
            V = X[:, fc, :]                                      # [n, 4, 3]
            nrm = np.cross(V[:, 2] - V[:, 0], V[:, 3] - V[:, 1])
            nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
            e1 = V[:, 1] - V[:, 0]; e1 /= np.linalg.norm(e1, axis=-1, keepdims=True)
            e2 = np.cross(nrm, e1)
            c0 = V.mean(axis=1, keepdims=True)
            P = np.stack([((V - c0) * e1[:, None, :]).sum(-1), ((V - c0) * e2[:, None, :]).sum(-1)], axis=-1)
            best = np.minimum(best, _quad_inner_radius(P))
        return best
    raise ValueError("inner_radius: unsupported element type")


def write_flat(mesh: Mesh, path) -> None:
    """Flat little-endian mesh file for the C++ host side (`SubrosaDG::MeshData::readFlat`,
    include/SubrosaDG_b200/SubrosaDG.hpp): magic "SDGM", dim, nblocks, per block {type, geom_order, n, nn, coords[n][nn][dim]},
    n_int, n_bnd, then le, lt, lf, re, rt, rf, rot, bc, phys as int32."""
    with open(path, "wb") as f:
        f.write(b"SDGM")
        f.write(np.array([mesh.dim, len(mesh.blocks)], dtype="<i4").tobytes())
        for t in sorted(mesh.blocks):
            c = np.ascontiguousarray(mesh.blocks[t]["coords"], dtype="<f8")
            f.write(np.array([t, mesh.blocks[t]["geom_order"], c.shape[0], c.shape[1]], dtype="<i4").tobytes())
            f.write(c.tobytes())
        fc = mesh.faces
        f.write(np.array([fc["n_int"], fc["n_bnd"]], dtype="<i4").tobytes())
        for k in ("le", "lt", "lf", "re", "rt", "rf", "rot", "bc", "phys"):
            f.write(np.ascontiguousarray(fc[k], dtype="<i4").tobytes())


# named meshes of the BASELINE configurations for the C++ example drivers (examples/*.cpp): `python -m subrosadg_b200.mesh NAME FILE [scale]`
EXAMPLE_MESHES = {
    # config 2: O-mesh of curved P3 quads around the NACA0012 (examples/naca0012_2d_ceuler.cpp:155-172 uses 4 transfinite blocks)
    "naca0012": lambda s: naca0012(nr=max(4, int(19 * s)), nt=max(12, 2 * int(39 * s))),
    # config 3: curved P3 quad rings at the cylinder + triangles outside (examples/karmanvortex_2d_cns.cpp: 4 ring blocks 15x11 + tris)
    "karmanvortex": lambda s: annulus(max(4, int(22 * s)), max(12, 4 * int(15 * s)), r0=0.5, r1=20.0, geom_order=3, stretch=2.0,
                                      tri_rings=max(2, int(11 * s)), phys_bc={1: RIEMANN_FARFIELD, 2: ADIABATIC_NONSLIP_WALL}),
    # config 5: cubed-sphere shell of curved P3 hexahedra (examples/sphere_3d_cns.cpp:259-296: 6 sphere blocks 11x11x9 + far blocks)
    # examples/sod_1d_ceuler.cpp:71-86: [0, 1] in 100 lines, physical 1 = left end, 2 = right end (both RiemannFarfield)
    "sod_1d": lambda s: box(1, (max(8, int(100 * s)),), 0.0, 1.0, phys_bc={1: RIEMANN_FARFIELD, 2: RIEMANN_FARFIELD}, geom_order=3),
    # examples/lidcavity_2d_incns.cpp:72-98: unit square, 20 x 20 quadrangles, physical 1 = bottom / right / left walls, 2 = the lid (y = 1)
    "lidcavity_2d": lambda s: box(2, (max(4, int(20 * s)),) * 2, 0.0, 1.0, phys_bc={1: ADIABATIC_NONSLIP_WALL, 2: ADIABATIC_NONSLIP_WALL}, geom_order=3,
                                  tagger=lambda c: np.where(np.abs(c[:, 1] - 1.0) < 1e-9, 2, 1).astype(np.int32)),
    # examples/thermalcavity_2d_incns.cpp:80-107: unit square, 80 x 80 P1 quadrangles, physical 1 = bottom / top (adiabatic), 2 = right wall (cold),
    # 3 = left wall (hot)
    "thermalcavity_2d": lambda s: box(2, (max(4, int(80 * s)),) * 2, 0.0, 1.0,
                                      phys_bc={1: ADIABATIC_NONSLIP_WALL, 2: ISOTHERMAL_NONSLIP_WALL, 3: ISOTHERMAL_NONSLIP_WALL},
                                      tagger=lambda c: np.where(np.abs(c[:, 0] - 1.0) < 1e-9, 2, np.where(np.abs(c[:, 0]) < 1e-9, 3, 1)).astype(np.int32)),
    "sphere": lambda s: cubed_sphere_shell(max(2, int(11 * s)), max(2, int(9 * s)), r0=0.5, r1=5.0, geom_order=3,
                                           phys_bc={1: RIEMANN_FARFIELD, 2: ADIABATIC_NONSLIP_WALL}),
}


if __name__ == "__main__":
    import sys
    if len(sys.argv) >= 4 and sys.argv[1] == "convert":
        # gmsh MSH 4.1 file -> flat mesh file; boundary types as physical:BoundaryConditionEnum pairs, e.g. 1:0,2:4
        from . import msh
        bc = {int(a): int(b) for a, b in (kv.split(":") for kv in (sys.argv[4] if len(sys.argv) > 4 else "1:0").split(","))}
        m = msh.read_msh(sys.argv[2], bc)
        write_flat(m, sys.argv[3])
        print({t: int(np.asarray(b["coords"]).shape[0]) for t, b in m.blocks.items()}, "faces", int(m.faces["n_int"]), "+", int(m.faces["n_bnd"]))
        raise SystemExit(0)
    if len(sys.argv) < 3 or sys.argv[1] not in EXAMPLE_MESHES:
        raise SystemExit(f"usage: python -m subrosadg_b200.mesh {{{'|'.join(EXAMPLE_MESHES)}}} out.sdgm [scale=1.0]\n"
                         "       python -m subrosadg_b200.mesh convert in.msh out.sdgm [phys:bc,...]")
    m = EXAMPLE_MESHES[sys.argv[1]](float(sys.argv[3]) if len(sys.argv) > 3 else 1.0)
    write_flat(m, sys.argv[2])
    print({t: int(np.asarray(b["coords"]).shape[0]) for t, b in m.blocks.items()}, "faces", int(m.faces["n_int"]), "+", int(m.faces["n_bnd"]))
