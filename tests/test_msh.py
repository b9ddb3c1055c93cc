"""MSH 4.1 reader / writer (subrosadg_b200/msh.py): a mesh written in gmsh's text format and read back must give the same element
blocks and, through the adjacency builder, the same face records (parents, local faces, rotations, boundary types, physical
indices) — bit for bit, the index-map bar of BASELINE.json."""
import numpy as np
import pytest

from subrosadg_b200 import mesh as M
from subrosadg_b200 import msh

FACE_KEYS = ("le", "lt", "lf", "re", "rt", "rf", "rot", "bc", "phys")

CASES = {
    "box2d_bc": (lambda: M.box(2, (5, 4), 0.0, 1.0, phys_bc={1: M.VELOCITY_INFLOW, 2: M.PRESSURE_OUTFLOW, 3: M.ADIABATIC_SLIP_WALL, 4: M.ADIABATIC_SLIP_WALL}), None),
    "box2d_curved": (lambda: M.box(2, (4, 4), 0.0, 1.0, geom_order=3, warp=lambda x: x + 0.04 * np.sin(np.pi * x[:, ::-1])), None),
    "periodic2d": (lambda: M.periodic_box(2, 5), [(0, 0.0, 2.0), (1, 0.0, 2.0)]),
    "periodic3d": (lambda: M.periodic_box(3, 3), [(0, 0.0, 2.0), (1, 0.0, 2.0), (2, 0.0, 2.0)]),
    "box3d": (lambda: M.box(3, (3, 2, 2), 0.0, 1.0), None),
    "hybrid": (lambda: M.annulus(4, 12, r0=0.5, r1=4.0, geom_order=3, tri_rings=2), None),
    "naca": (lambda: M.naca0012(nr=4, nt=12), None),
    "sphere": (lambda: M.cubed_sphere_shell(2, 2, geom_order=2), None),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_msh_roundtrip(tmp_path, name):
    make, periodic = CASES[name]
    m = make()
    path = tmp_path / f"{name}.msh"
    pairs = None
    if periodic:
        xyz, _ = msh.node_table(m)
        pairs = msh.periodic_node_pairs(xyz, periodic)
    msh.write_msh(m, path, periodic_pairs=pairs)
    head = open(path).read(64)
    assert head.startswith("$MeshFormat\n4.1 0 8\n$EndMeshFormat")
    r = msh.read_msh(path, m.phys_bc)
    assert r.dim == m.dim and sorted(r.blocks) == sorted(m.blocks)
    for t in m.blocks:
        assert r.blocks[t]["geom_order"] == m.blocks[t]["geom_order"]
        # %.17g round-trips doubles exactly; nodes shared by two elements are merged in the file (their copies may differ by an ulp)
        assert np.abs(r.blocks[t]["coords"] - np.asarray(m.blocks[t]["coords"])).max() < 1e-12
    assert int(r.faces["n_int"]) == int(m.faces["n_int"]) and int(r.faces["n_bnd"]) == int(m.faces["n_bnd"])
    for k in FACE_KEYS:
        assert np.array_equal(np.asarray(r.faces[k]), np.asarray(m.faces[k])), k


def test_msh_rejects_unknown_boundary(tmp_path):
    m = M.box(2, (3, 3), 0.0, 1.0)
    path = tmp_path / "b.msh"
    msh.write_msh(m, path)
    with pytest.raises(ValueError, match="no boundary condition"):
        msh.read_msh(path, {1: M.RIEMANN_FARFIELD})
    txt = open(path).read().replace("4.1 0 8", "2.2 0 8")
    open(path, "w").write(txt)
    with pytest.raises(ValueError, match="4.1"):
        msh.read_msh(path, m.phys_bc)


def test_gmsh_type_numbers():
    """element type numbers of the MSH format for the orders the reference instantiates (P1..P5)"""
    assert msh.GMSH_NUMBER[(M.QUADRANGLE, 3)] == 36 and msh.GMSH_NUMBER[(M.HEXAHEDRON, 3)] == 92 and msh.GMSH_NUMBER[(M.TRIANGLE, 3)] == 21
    assert msh.GMSH_NUMBER[(M.LINE, 3)] == 26 and msh.GMSH_NUMBER[(M.HEXAHEDRON, 1)] == 5 and msh.GMSH_NUMBER[(M.POINT, 0)] == 15
    for (t, g), num in msh.GMSH_NUMBER.items():
        if t != M.POINT:
            assert len(M.gmsh_reference_nodes(t, g)) == {M.LINE: g + 1, M.TRIANGLE: (g + 1) * (g + 2) // 2, M.QUADRANGLE: (g + 1) ** 2, M.HEXAHEDRON: (g + 1) ** 3}[t]


def test_file_in_gmsh_own_layout():
    """tests/data/lidcavity_2x2_gmsh_layout.msh is written BY HAND in the layout Gmsh 4.13 itself produces for the reference's lid-cavity script
    (examples/lidcavity_2d_incns.cpp:72-98 at 2 x 2 cells, order 1): $PhysicalNames, four point entities without physical tags, SEVERAL curves in
    one physical group, signed bounding-entity lists, one node block per entity (corner nodes on the point entities), elements only for
    entities in physical groups.  It is not a Gmsh-written file (no Gmsh here) — it checks that the reader does not lean on the simpler
    layout of this repository's own writer."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "lidcavity_2x2_gmsh_layout.msh")
    r = msh.read_msh(path, {1: M.ADIABATIC_NONSLIP_WALL, 2: M.ISOTHERMAL_NONSLIP_WALL})
    assert r.dim == 2 and sorted(r.blocks) == [M.QUADRANGLE] and r.blocks[M.QUADRANGLE]["geom_order"] == 1
    xy = np.asarray(r.blocks[M.QUADRANGLE]["coords"])
    assert xy.shape == (4, 4, 2)
    ref = M.box(2, (2, 2), 0.0, 1.0)
    cen = lambda c: sorted(map(tuple, np.round(np.asarray(c).mean(axis=1), 12)))
    assert cen(xy) == cen(ref.blocks[M.QUADRANGLE]["coords"])
    # counter-clockwise corners (positive area), as the reference's Jacobians need
    a = 0.5 * np.abs(np.sum(xy[:, :, 0] * np.roll(xy[:, :, 1], -1, axis=1) - np.roll(xy[:, :, 0], -1, axis=1) * xy[:, :, 1], axis=1))
    s = 0.5 * np.sum(xy[:, :, 0] * np.roll(xy[:, :, 1], -1, axis=1) - np.roll(xy[:, :, 0], -1, axis=1) * xy[:, :, 1], axis=1)
    assert np.allclose(a, 0.25) and np.all(s > 0)
    f = r.faces
    n_int, n_bnd = int(f["n_int"]), int(f["n_bnd"])
    assert (n_int, n_bnd) == (4, 8)
    phys = np.asarray(f["phys"])[n_int:]; bc = np.asarray(f["bc"])[n_int:]
    assert sorted(phys.tolist()) == [1] * 6 + [2] * 2 and np.array_equal(bc == M.ISOTHERMAL_NONSLIP_WALL, phys == 2)
    # the two lid faces (curve 3, physical 2) lie on y = 1
    for k in np.flatnonzero(phys == 2):
        e, lf = int(f["le"][n_int + k]), int(f["lf"][n_int + k])
        corners = xy[e][M.FACE_CORNERS[M.QUADRANGLE][lf]]
        assert np.allclose(corners[:, 1], 1.0)


def test_file_in_gmsh_own_layout_runs_through_the_oracle():
    """a uniform flow on that mesh is steady (all four walls replaced by far-field faces): the face records the reader produced are consistent"""
    import os
    import oracle
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "lidcavity_2x2_gmsh_layout.msh")
    r = msh.read_msh(path, {1: M.RIEMANN_FARFIELD, 2: M.RIEMANN_FARFIELD})
    O = oracle.Oracle(dict(p=2, conv_flux=2, rk=2), r)
    one = lambda x, *a: np.stack([1.4 * np.ones(x.shape[:-1]), 0.3 * np.ones(x.shape[:-1]), 0.1 * np.ones(x.shape[:-1]), np.ones(x.shape[:-1])], axis=-1)
    O.initialize(one, one)
    before = O.get_state(M.QUADRANGLE).copy()
    O.step(0.5 * O.compute_dt(1.0), 3)
    assert np.abs(O.get_state(M.QUADRANGLE) - before).max() < 1e-13
