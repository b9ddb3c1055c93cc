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
