"""Gmsh MSH 4.1 (ASCII) reader / writer for the host side: the caller-side input format of the hot path.

SubrosaDG reads its meshes through the Gmsh API (src/Mesh/ReadControl.cpp:243-335: gmsh::open, getElementsByType, getNodes, ...)
and derives faces, boundary types and periodic pairs in src/Mesh/Adjacency.cpp.  Gmsh is not available here, so `read_msh` parses
the published MSH 4.1 text format directly and feeds the same adjacency builder as the in-code mesh producers
(`mesh.build_faces`, a restatement of Adjacency.cpp's semantics):

  * elements of the mesh dimension become the element blocks (coordinates in gmsh node order, sorted by element tag like
    getElementsByType), element order = geometry order like the reference (`gmsh::model::mesh::setOrder(P)`);
  * elements of dimension D-1 carry the boundary: their entity's first physical tag is `gmsh_physical_index_`
    (Adjacency.cpp:330-430), mapped to a BoundaryConditionEnum by `phys_bc` (System::addBoundaryCondition);
  * `$Periodic` node correspondences identify slave with master nodes, so periodic faces pair up as interior faces
    (Adjacency.cpp:226-320).

`write_msh` emits the same subset (one entity per physical group), which is what the round-trip tests use; a file written by
Gmsh itself has not been available to test against — the parser follows the format specification (MSH file format version 4.1).
"""
from __future__ import annotations

import numpy as np

from . import mesh as M

# gmsh element type number -> (ElementEnum, order), and back
GMSH_TYPES = {15: (M.POINT, 0),
              1: (M.LINE, 1), 8: (M.LINE, 2), 26: (M.LINE, 3), 27: (M.LINE, 4), 28: (M.LINE, 5),
              2: (M.TRIANGLE, 1), 9: (M.TRIANGLE, 2), 21: (M.TRIANGLE, 3), 23: (M.TRIANGLE, 4), 25: (M.TRIANGLE, 5),
              3: (M.QUADRANGLE, 1), 10: (M.QUADRANGLE, 2), 36: (M.QUADRANGLE, 3), 37: (M.QUADRANGLE, 4), 38: (M.QUADRANGLE, 5),
              5: (M.HEXAHEDRON, 1), 12: (M.HEXAHEDRON, 2), 92: (M.HEXAHEDRON, 3), 93: (M.HEXAHEDRON, 4), 94: (M.HEXAHEDRON, 5)}
GMSH_NUMBER = {v: k for k, v in GMSH_TYPES.items()}


def _sections(path):
    out, name, buf = {}, None, []
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s:
                continue
            if s.startswith("$End"):
                out[name] = buf
                name, buf = None, []
            elif s.startswith("$"):
                name, buf = s[1:], []
            elif name is not None:
                buf.append(s)
    return out


def read_msh(path, phys_bc: dict, dim: int | None = None) -> M.Mesh:
    """Parse an MSH 4.1 ASCII file into a `mesh.Mesh`.  phys_bc: gmsh physical index -> BoundaryConditionEnum value."""
    sec = _sections(path)
    ver = sec["MeshFormat"][0].split()
    if not ver[0].startswith("4.1") or int(ver[1]) != 0:
        raise ValueError("only MSH 4.1 ASCII files are supported")
    # entities: (dim, tag) -> first physical tag
    ent_phys = {}
    if "Entities" in sec:
        L = sec["Entities"]
        counts = [int(x) for x in L[0].split()]
        row = 1
        for d in range(4):
            for _ in range(counts[d]):
                tok = L[row].split(); row += 1
                tag = int(tok[0])
                k = 4 if d == 0 else 7   # point: tag x y z; others: tag + bounding box
                nphys = int(tok[k])
                if nphys:
                    ent_phys[(d, tag)] = abs(int(tok[k + 1]))
    # nodes
    L = sec["Nodes"]
    nblk, nnodes = int(L[0].split()[0]), int(L[0].split()[1])
    tags = np.zeros(nnodes, dtype=np.int64); xyz = np.zeros((nnodes, 3)); at, row = 0, 1
    for _ in range(nblk):
        _, _, parametric, n = (int(x) for x in L[row].split()); row += 1
        tags[at:at + n] = [int(L[row + i]) for i in range(n)]; row += n
        for i in range(n):
            xyz[at + i] = [float(x) for x in L[row + i].split()[:3]]
        row += n; at += n
    index_of = {int(t): i for i, t in enumerate(tags)}
    # elements
    L = sec["Elements"]
    nblk, row = int(L[0].split()[0]), 1
    by_type = {}    # gmsh type -> list of (element tag, node indices, physical)
    for _ in range(nblk):
        ed, etag, gtype, n = (int(x) for x in L[row].split()); row += 1
        phys = ent_phys.get((ed, etag), 0)
        for i in range(n):
            tok = [int(x) for x in L[row + i].split()]
            by_type.setdefault(gtype, []).append((tok[0], [index_of[t] for t in tok[1:]], phys))
        row += n
    for g in by_type:
        if g not in GMSH_TYPES:
            raise ValueError(f"gmsh element type {g} is not supported (line / triangle / quadrangle / hexahedron, order <= 5)")
    if dim is None:
        dim = max(M.ELEM_DIM[GMSH_TYPES[g][0]] for g in by_type)
    vol = {g: v for g, v in by_type.items() if M.ELEM_DIM[GMSH_TYPES[g][0]] == dim}
    bnd = {g: v for g, v in by_type.items() if M.ELEM_DIM[GMSH_TYPES[g][0]] == dim - 1}
    orders = {GMSH_TYPES[g][1] for g in vol}
    if len(orders) != 1:
        raise ValueError("all elements must have the same order")
    order = orders.pop()
    # periodic node identification: slave -> master (chains resolved)
    canon = np.arange(nnodes, dtype=np.int64)
    if "Periodic" in sec:
        L = sec["Periodic"]
        nlink, row = int(L[0]), 1
        for _ in range(nlink):
            row += 1                                  # entityDim entityTag entityTagMaster
            naff = int(L[row].split()[0]); row += 1   # affine transform on the same line
            ncorr = int(L[row]); row += 1
            for i in range(ncorr):
                s, m = (int(x) for x in L[row + i].split())
                canon[index_of[s]] = index_of[m]
            row += ncorr
            del naff
        for _ in range(8):                            # corners of a doubly / triply periodic box chain through several links
            canon = canon[canon]
    elems, corner_ids = {}, {}
    for g, lst in vol.items():
        t = GMSH_TYPES[g][0]
        lst = sorted(lst, key=lambda r: r[0])
        conn = np.array([r[1] for r in lst], dtype=np.int64)
        elems[t] = np.ascontiguousarray(xyz[conn][:, :, :dim])
        corner_ids[t] = conn[:, :M.CORNER_COORDS[t].shape[0]]
    # physical index of a boundary face: the (D-1)-element with the same corner nodes
    bnd_phys = {}
    for g, lst in bnd.items():
        ft = GMSH_TYPES[g][0]
        nc = {M.POINT: 1, M.LINE: 2, M.QUADRANGLE: 4, M.TRIANGLE: 3}[ft]
        for _, nodes, phys in lst:
            bnd_phys[tuple(sorted(nodes[:nc]))] = phys

    faces = M.build_faces(dim, corner_ids, canon, lambda c: np.zeros(len(c), dtype=np.int32), {0: M.RIEMANN_FARFIELD},
                          lambda t, e, lf: np.zeros((len(e), dim)))
    n_int, n_bnd = int(faces["n_int"]), int(faces["n_bnd"])
    for k in range(n_int, n_int + n_bnd):
        t, e, lf = int(faces["lt"][k]), int(faces["le"][k]), int(faces["lf"][k])
        key = tuple(sorted(int(x) for x in corner_ids[t][e][M.FACE_CORNERS[t][lf]]))
        if key not in bnd_phys:
            raise ValueError("boundary face without a physical group (every boundary needs a (D-1)-dimensional physical group or a $Periodic link)")
        p = bnd_phys[key]
        if p not in phys_bc:
            raise ValueError(f"no boundary condition for gmsh physical index {p} (System::addBoundaryCondition)")
        faces["phys"][k] = p
        faces["bc"][k] = phys_bc[p]
    blocks = {t: dict(coords=elems[t], geom_order=order, corners=corner_ids[t]) for t in elems}
    return M.Mesh(dim=dim, blocks=blocks, faces=faces, phys_bc=dict(phys_bc), info=dict(kind="msh", path=str(path)))


def write_msh(mesh: M.Mesh, path, periodic_pairs=None) -> None:
    """Write `mesh` as MSH 4.1 ASCII: one entity per physical boundary group and one for the elements of each type.
    periodic_pairs: optional list of (slave_xyz -> master_xyz) callables' result, i.e. arrays [(slave_node, master_node)] of 0-based
    node indices into the node list this function builds (use `node_table(mesh)` to get it)."""
    dim = mesh.dim
    xyz, conn = node_table(mesh)
    g = next(iter(mesh.blocks.values()))["geom_order"]
    f = mesh.faces
    n_int, n_bnd = int(f["n_int"]), int(f["n_bnd"])
    # boundary elements: the face nodes of the left parent in the face's own (gmsh) node order
    bnd = {}   # phys -> list of node index lists
    for k in range(n_int, n_int + n_bnd):
        t, e, lf = int(f["lt"][k]), int(f["le"][k]), int(f["lf"][k])
        bnd.setdefault(int(f["phys"][k]), []).append(_face_nodes(t, g, lf, conn[t][e]))
    ft = M.FACE_TYPE[next(iter(mesh.blocks))]
    with open(path, "w") as out:
        out.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
        nb, nv = len(bnd), len(mesh.blocks)
        cnt = [0, 0, 0, 0]; cnt[dim - 1] = nb; cnt[dim] = nv
        out.write("$Entities\n%d %d %d %d\n" % tuple(cnt))
        for d in (dim - 1, dim):
            tags = sorted(bnd) if d == dim - 1 else list(range(1, nv + 1))
            for tag in tags:
                if d == 0:
                    out.write(f"{tag} 0 0 0 1 {tag}\n")
                else:
                    out.write(f"{tag} 0 0 0 0 0 0 " + (f"1 {tag} 0\n" if d == dim - 1 else "0 0\n"))
        out.write("$EndEntities\n")
        out.write(f"$Nodes\n1 {len(xyz)} 1 {len(xyz)}\n{dim} 1 0 {len(xyz)}\n")
        out.write("\n".join(str(i + 1) for i in range(len(xyz))) + "\n")
        pad = np.zeros((len(xyz), 3)); pad[:, :dim] = xyz
        out.write("\n".join("%.17g %.17g %.17g" % tuple(p) for p in pad) + "\n$EndNodes\n")
        nel = sum(len(v) for v in bnd.values()) + sum(len(c) for c in conn.values())
        out.write(f"$Elements\n{nb + nv} {nel} 1 {nel}\n")
        tag = 1
        for p in sorted(bnd):
            out.write(f"{dim - 1} {p} {GMSH_NUMBER[(ft, g if ft != M.POINT else 0)]} {len(bnd[p])}\n")
            for nodes in bnd[p]:
                out.write(str(tag) + " " + " ".join(str(i + 1) for i in nodes) + "\n"); tag += 1
        for bi, t in enumerate(sorted(mesh.blocks)):
            out.write(f"{dim} {bi + 1} {GMSH_NUMBER[(t, g)]} {len(conn[t])}\n")
            for nodes in conn[t]:
                out.write(str(tag) + " " + " ".join(str(int(i) + 1) for i in nodes) + "\n"); tag += 1
        out.write("$EndElements\n")
        if periodic_pairs is not None and len(periodic_pairs):
            out.write(f"$Periodic\n1\n{dim - 1} 1 2\n0\n{len(periodic_pairs)}\n")
            out.write("\n".join(f"{int(s) + 1} {int(m) + 1}" for s, m in periodic_pairs) + "\n$EndPeriodic\n")


def node_table(mesh: M.Mesh, tol=1e-9):
    """Unique nodes of a mesh: (xyz [nnodes, D], {type: connectivity [n, nn]}) with coincident element nodes merged."""
    allc = np.concatenate([np.asarray(mesh.blocks[t]["coords"]).reshape(-1, mesh.dim) for t in sorted(mesh.blocks)])
    q = np.round(allc / tol).astype(np.int64)
    _, first, inv = np.unique(q, axis=0, return_index=True, return_inverse=True)
    inv = inv.reshape(-1)
    conn, off = {}, 0
    for t in sorted(mesh.blocks):
        n, nn = np.asarray(mesh.blocks[t]["coords"]).shape[:2]
        conn[t] = inv[off:off + n * nn].reshape(n, nn); off += n * nn
    return allc[first], conn


def periodic_node_pairs(xyz, axes_lo_hi, tol=1e-9):
    """(slave, master) node index pairs for an axis-aligned periodic box: nodes on x[a] = hi are slaves of x[a] = lo."""
    pairs = []
    key = {tuple(np.round(p / tol).astype(np.int64)): i for i, p in enumerate(xyz)}
    for a, lo, hi in axes_lo_hi:
        for i in np.flatnonzero(np.abs(xyz[:, a] - hi) < tol * max(1.0, abs(hi - lo))):
            p = xyz[i].copy(); p[a] = lo
            pairs.append((i, key[tuple(np.round(p / tol).astype(np.int64))]))
    return pairs


def _face_nodes(t, g, lf, elem_nodes):
    """Node indices of local face `lf` of an order-g element in the FACE element's own gmsh node order: the face's reference nodes
    (order g, gmsh order) mapped through the P1 map of the face corners into the parent and matched with the parent's nodes."""
    ft = M.FACE_TYPE[t]
    if ft == M.POINT:
        return [int(elem_nodes[M.FACE_CORNERS[t][lf][0]])]
    pref = M.gmsh_reference_nodes(t, g)
    fref = M.gmsh_reference_nodes(ft, g)
    corners = M.CORNER_COORDS[t][M.FACE_CORNERS[t][lf]]
    if ft == M.LINE:
        s = (fref[:, 0] + 1.0) * 0.5
        pts = corners[0][None, :] * (1 - s)[:, None] + corners[1][None, :] * s[:, None]
    else:  # quadrangle face
        s, r = (fref[:, 0] + 1.0) * 0.5, (fref[:, 1] + 1.0) * 0.5
        pts = (corners[0][None] * ((1 - s) * (1 - r))[:, None] + corners[1][None] * (s * (1 - r))[:, None] +
               corners[2][None] * (s * r)[:, None] + corners[3][None] * ((1 - s) * r)[:, None])
    out = []
    for p in pts:
        d = np.abs(pref - p[None, :]).sum(axis=1)
        k = int(np.argmin(d))
        if d[k] > 1e-9:
            raise ValueError("face node not found on the parent element")
        out.append(int(elem_nodes[k]))
    return out
