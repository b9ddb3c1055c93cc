"""CPU tests of the dense-operator ("mixed") host plan: triangle / quadrangle tables and geometry of the product library
(subrosadg_b200/csrc/mixed_tables.hpp, mixed_path.cu host part) against the oracle's independent restatement of
src/Mesh/BasisFunction.cpp:136-230, src/Mesh/Quadrature.cpp and src/Mesh/Geometry.cpp.  Plan-only contexts (device = -1)."""
import numpy as np
import pytest

import oracle
from subrosadg_b200 import mesh as M
from subrosadg_b200 import solver as sv

TRI, QUAD = M.TRIANGLE, M.QUADRANGLE


def hybrid(geom_order=3):
    return M.annulus(4, 12, r0=0.5, r1=4.0, geom_order=geom_order, tri_rings=2)


@pytest.mark.parametrize("p", [1, 2, 3])
def test_tables_match_oracle(built, p):
    m = hybrid()
    O = oracle.Oracle(dict(p=p), m)
    S = sv.Solver(dict(p=p), m, device=-1)
    assert S.types == [TRI, QUAD]
    for t in S.types:
        so, ss = O.sizes(t), S.sizes(t)
        assert vars(so) == vars(ss), (t, so, ss)
        Phi = S.debug_plan(100 * t + 0).reshape(ss.Nq, ss.Nb)
        dPhi = S.debug_plan(100 * t + 1).reshape(ss.Nq * 2, ss.Nb)
        PhiF = S.debug_plan(100 * t + 2).reshape(ss.Naq, ss.Nb)
        assert np.abs(Phi - O.table(t, 0)).max() < 1e-14
        assert np.abs(dPhi - O.table(t, 1)).max() < 1e-13
        assert np.abs(PhiF - O.table(t, 2)).max() < 1e-14
        # least-squares projection reproduces every modal function from its point values (InitialCondition.cpp:100-102)
        Proj = S.debug_plan(100 * t + 3).reshape(ss.Nb, ss.Nq)
        assert np.abs(Proj @ Phi - np.eye(ss.Nb)).max() < 1e-12


@pytest.mark.parametrize("g", [1, 3])
def test_geometry_matches_oracle(built, g):
    m = hybrid(g)
    O = oracle.Oracle(dict(p=3), m)
    S = sv.Solver(dict(p=3), m, device=-1)
    for t in S.types:
        s = S.sizes(t)
        assert np.abs(S.quadrature_coordinates(t) - O.quadrature_coordinates(t)).max() < 1e-13
        jw = S.debug_plan(100 * t + 4).reshape(s.n, s.Nq)
        mt = S.debug_plan(100 * t + 5).reshape(s.n, s.Nq, 4)
        Mi = S.debug_plan(100 * t + 6).reshape(s.n, s.Nb, s.Nb)
        me = S.debug_plan(100 * t + 7)
        assert np.allclose(jw, O.element_geometry(t, 1), rtol=1e-12, atol=0)
        assert np.allclose(mt, O.element_geometry(t, 2), rtol=1e-11, atol=1e-15)
        ref = O.element_geometry(t, 3)
        assert np.abs(Mi - ref).max() <= 1e-10 * np.abs(ref).max()
        assert np.allclose(me, O.element_geometry(t, 4), rtol=1e-14)
        # M^-1 really inverts Phi^T diag(detJ w) Phi (Geometry.cpp:88-100)
        Phi = S.debug_plan(100 * t + 0).reshape(s.Nq, s.Nb)
        Mass = np.einsum("qa,eq,qb->eab", Phi, jw, Phi)
        assert np.abs(np.einsum("eab,ebc->eac", Mass, Mi) - np.eye(s.Nb)).max() < 1e-9
    nrm = S.debug_plan(90).reshape(-1, 4, 2)
    fjw = S.debug_plan(91).reshape(-1, 4)
    assert np.abs(nrm - O.face_geometry(1)).max() < 1e-13
    assert np.allclose(fjw, O.face_geometry(2), rtol=1e-12)
    assert np.abs(S.boundary_quadrature_coordinates() - O.boundary_quadrature_coordinates()).max() < 1e-13


def test_triangle_rule_and_basis_properties(built):
    """Exactness of the Dunavant rules for total degree 2p and partition of unity of the vertex functions."""
    m = M.annulus(2, 6, tri_rings=2, geom_order=1)
    for p in (1, 2, 3):
        S = sv.Solver(dict(p=p), m, device=-1)
        s = S.sizes(TRI)
        assert (s.Nb, s.Nq) == ((p + 1) * (p + 2) // 2, {1: 3, 2: 6, 3: 12}[p])   # SimulationControl.cpp:243-273
        Phi = S.debug_plan(100 * TRI + 0).reshape(s.Nq, s.Nb)
        assert np.abs(Phi[:, :3].sum(axis=1) - 1.0).max() < 1e-14
        dPhi = S.debug_plan(100 * TRI + 1).reshape(s.Nq, 2, s.Nb)
        assert np.abs(dPhi[:, :, :3].sum(axis=2)).max() < 1e-13
        # sum_q detJ w = element area (affine triangles)
        jw = S.debug_plan(100 * TRI + 4).reshape(s.n, s.Nq)
        X = m.blocks[TRI]["coords"]
        area = 0.5 * np.abs((X[:, 1, 0] - X[:, 0, 0]) * (X[:, 2, 1] - X[:, 0, 1]) - (X[:, 2, 0] - X[:, 0, 0]) * (X[:, 1, 1] - X[:, 0, 1]))
        assert np.allclose(jw.sum(axis=1), area, rtol=1e-13)


def test_mixed_path_refuses_unsupported_requests(built):
    m = hybrid()
    S = sv.Solver(dict(p=3), m, device=-1)
    with pytest.raises(RuntimeError, match="no CUDA device|CPU path"):
        S.stepSolver(1e-3, 1)
    with pytest.raises(RuntimeError, match="one GPU|ghost"):
        sv.Solver(dict(p=3), m, device=-1, n_ghost={TRI: 1})
