"""CPU tests: the oracle's (and the product's host-side) tables against the integer tables embedded in the reference
(tests/golden/reference_tables.json, extracted from src/Solver/SimulationControl.cpp by tests/golden/make_golden.py) and
against exactness properties.  These literals are the only golden vectors the reference holds for the hot path."""
import json
import os

import numpy as np
import pytest

import oracle
from subrosadg_b200 import mesh as M

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_tables.json")))
NAMES = {1: "Line", 2: "Triangle", 3: "Quadrangle", 6: "Hexahedron"}


def test_golden_file_is_complete():
    assert set(G["quadrature_number"]) >= {"Line", "Triangle", "Quadrangle", "Hexahedron"}
    assert len(G["adjacency_quadrature_sequence"]["tables"]) == 41
    assert len(G["adjacency_view_node_parent_sequence"]["tables"]) >= 100


def test_face_corner_maps_match_reference(built):
    t = G["per_adjacency_node_index"]["tables"]
    for et, name in NAMES.items():
        assert M.FACE_CORNERS[et].reshape(-1).tolist() == t[name], name


@pytest.mark.parametrize("p", [1, 2, 3, 4, 5])
def test_face_point_permutations_match_reference(built, p):
    t = G["adjacency_quadrature_sequence"]["tables"]
    assert oracle.face_sequence(oracle.LINE, p, 0).tolist() == t[f"Line/P{p}"]
    for rot in range(4):
        assert oracle.face_sequence(oracle.QUADRANGLE, p, rot).tolist() == t[f"Quadrangle/P{p}/case{rot}"], (p, rot)


# P1 quadrangles run on the dense-operator path (Gmsh's seven-point "Gauss2" rule; the line faces there are reversed in place, mixed_path.cu)
@pytest.mark.parametrize("dim,p", [(2, 2), (2, 3), (3, 1), (3, 2), (3, 3)])
def test_product_face_point_permutations_match_reference(built, dim, p):
    from subrosadg_b200.solver import Solver
    S = Solver(dict(p=p), M.periodic_box(dim, 3), device=-1)
    s = S.sizes(S.types[0])
    seq = S.debug_plan(16).reshape(4, s.Nqf)
    t = G["adjacency_quadrature_sequence"]["tables"]
    for rot in range(4):
        want = t[f"Line/P{p}"] if dim == 2 else t[f"Quadrangle/P{p}/case{rot}"]
        assert seq[rot].tolist() == want


def test_quadrature_counts_match_reference(built):
    q = G["quadrature_number"]
    for p in (1, 2, 3):
        s = oracle.Oracle(dict(p=p), M.periodic_box(3, 3)).sizes(oracle.HEXAHEDRON)
        assert s.Nq == q["Hexahedron"][2 * p] and s.Nqf == q["Quadrangle"][2 * p + 1]
        s = oracle.Oracle(dict(p=p), M.periodic_box(2, 3)).sizes(oracle.QUADRANGLE)
        assert s.Nqf == q["Line"][2 * p + 1]
        if p >= 2:  # p = 1: gmsh's 7-point order-2 quadrangle rule is replaced by the 2x2 tensor rule (documented deviation)
            assert s.Nq == q["Quadrangle"][2 * p]
        s = oracle.Oracle(dict(p=p), M.annulus(3, 8, geom_order=1, tri_rings=3)).sizes(oracle.TRIANGLE)
        assert s.Nq == q["Triangle"][2 * p]


def _face_nodes_through_corner_map(parent, face, order):
    """reference nodes of the face element (gmsh order) pushed through the P1 map of the parent's face corners"""
    ft = M.FACE_TYPE[parent]
    fn = M.gmsh_reference_nodes(ft, order)
    c = M.CORNER_COORDS[parent][M.FACE_CORNERS[parent][face]]
    if ft == M.LINE:
        s = fn[:, 0]
        return 0.5 * (1 - s)[:, None] * c[0] + 0.5 * (1 + s)[:, None] * c[1]
    s, t = fn[:, 0], fn[:, 1]
    return 0.25 * ((1 - s) * (1 - t))[:, None] * c[0] + 0.25 * ((1 + s) * (1 - t))[:, None] * c[1] + \
        0.25 * ((1 + s) * (1 + t))[:, None] * c[2] + 0.25 * ((1 - s) * (1 + t))[:, None] * c[3]


@pytest.mark.parametrize("parent,order", [(3, 1), (3, 2), (3, 3), (3, 4), (6, 1), (6, 2), (6, 3)])
def test_high_order_node_numbering_matches_reference(built, parent, order):
    """getAdjacencyElementViewNodeParentSequence (SimulationControl.cpp:525-887) lists, per local face, the parent node ids
    of the face's nodes in the face's own gmsh order: pins the gmsh node numbering used for curved geometry."""
    t = G["adjacency_view_node_parent_sequence"]["tables"]
    ft = NAMES[M.FACE_TYPE[parent]]
    nodes_py = M.gmsh_reference_nodes(parent, order)
    nodes_or = oracle.reference_nodes(parent, order)
    assert np.allclose(nodes_py, nodes_or, atol=1e-14)
    for f in range(len(M.FACE_CORNERS[parent])):
        ids = t[f"{ft}/in{NAMES[parent]}/P{order}/case{f}"]
        want = _face_nodes_through_corner_map(parent, f, order)
        assert np.allclose(nodes_py[ids], want, atol=1e-13), (parent, order, f)


def test_product_node_lattice_matches_reference(built):
    """The product's own gmsh node lattice (host_tables.hpp) reproduces the same coordinates: checked through the
    quadrature coordinates of a curved mesh handed over in gmsh node order (see test_host_plan.py)."""
    from subrosadg_b200.solver import Solver
    warp = lambda x: x + 0.05 * np.sin(np.pi * x[:, ::-1])
    m = M.box(2, (3, 3), 0.0, 1.0, geom_order=4, warp=warp)
    S = Solver(dict(p=3), m, device=-1)
    O = oracle.Oracle(dict(p=3), m)
    assert np.abs(S.quadrature_coordinates(3) - O.quadrature_coordinates(3)).max() < 1e-14


@pytest.mark.parametrize("etype,mesh", [(3, lambda: M.periodic_box(2, 3)), (6, lambda: M.periodic_box(3, 3)),
                                        (2, lambda: M.annulus(3, 8, geom_order=1, tri_rings=3))])
@pytest.mark.parametrize("p", [1, 2, 3])
def test_quadrature_and_basis_properties(built, etype, mesh, p):
    O = oracle.Oracle(dict(p=p), mesh())
    pts, w = O.table(etype, 4), O.table(etype, 5)
    measure = {2: 0.5, 3: 4.0, 6: 8.0}[etype]   # getElementMeasure, SimulationControl.cpp:218-241
    assert abs(w.sum() - measure) < 1e-14
    D = 2 if etype in (2, 3) else 3
    # exactness for every monomial the scheme integrates (degree 2p per direction / total degree 2p on the triangle)
    deg = 2 * p
    for a in range(deg + 1):
        for b in range(deg + 1 - (a if etype == 2 else 0)):
            val = np.sum(w * pts[:, 0] ** a * pts[:, 1] ** b)
            if etype == 2:
                from math import factorial
                exact = factorial(a) * factorial(b) / factorial(a + b + 2)
            else:
                ex1 = lambda k: 0.0 if k % 2 else 2.0 / (k + 1)
                exact = ex1(a) * ex1(b) * (2.0 if D == 3 else 1.0)
            assert abs(val - exact) < 2e-14, (a, b, val, exact)
    Phi, dPhi = O.table(etype, 0), O.table(etype, 1)   # [Nq, Nb], [Nq*D, Nb]
    s = O.sizes(etype)
    assert np.linalg.matrix_rank(Phi) == s.Nb
    # the modal space contains the constants and the linear functions, and its gradient table differentiates them
    one = np.linalg.lstsq(Phi, np.ones(s.Nq), rcond=None)[0]
    assert np.abs(Phi @ one - 1).max() < 1e-12 and np.abs(dPhi @ one).max() < 1e-11
    cx = np.linalg.lstsq(Phi, pts[:, 0], rcond=None)[0]
    g = (dPhi @ cx).reshape(s.Nq, D)
    assert np.abs(g[:, 0] - 1).max() < 1e-11 and np.abs(g[:, 1:]).max() < 1e-11
    # (Phi^T Phi)^-1 (modal_least_squares_inverse_, BasisFunction.cpp:217)
    assert np.abs(O.table(etype, 3) @ (Phi.T @ Phi) - np.eye(s.Nb)).max() < 1e-9


def test_product_modal_convention_equals_oracle(built):
    """Phi (modal functions at the quadrature points) is the seam transform of the product; it must be the oracle's."""
    from subrosadg_b200.solver import Solver
    for dim, et in ((2, 3), (3, 6)):
        for p in (1, 2, 3):
            m = M.periodic_box(dim, 3)
            S = Solver(dict(p=p), m, device=-1)
            O = oracle.Oracle(dict(p=p), m)
            s = S.sizes(et)
            dense = dim == 2 and p == 1      # P1 quadrangles: seven-point rule, dense-operator tables (100 * type + 0 = Phi)
            assert s.Nq == (7 if dense else (p + 1) ** dim)
            assert np.abs(S.debug_plan(100 * et if dense else 4).reshape(s.Nq, s.Nb) - O.table(et, 0)).max() < 1e-15
