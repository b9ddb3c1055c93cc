"""CPU checks of the artificial-viscosity restatement in the oracle (SpatialDiscrete.cpp:37-192 and the eps * grad(U) flux terms) and of the
mesh data it needs (node tags of the corner nodes, inner radius)."""
import numpy as np
import pytest

import cases
import oracle
from subrosadg_b200 import mesh as M


def test_inner_radius_known_shapes():
    assert np.allclose(M.inner_radius(M.box(1, (4,), 0.0, 2.0), M.LINE), 0.25)                       # half length
    assert np.allclose(M.inner_radius(M.box(2, (3, 2), 0.0, [3.0, 1.0]), M.QUADRANGLE), 0.25)      # rectangles 1 x 0.5: half the short side
    assert np.allclose(M.inner_radius(M.box(3, (2, 2, 2), 0.0, [2.0, 1.0, 4.0]), M.HEXAHEDRON), 0.25)
    tri = M.annulus(2, 8, r0=0.5, r1=2.0, geom_order=1, tri_rings=2)
    X = np.asarray(tri.blocks[M.TRIANGLE]["coords"])[:, :3, :]
    a, b, c = (np.linalg.norm(X[:, i] - X[:, (i + 1) % 3], axis=-1) for i in range(3))
    d1, d2 = X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]
    area = 0.5 * np.abs(d1[:, 0] * d2[:, 1] - d1[:, 1] * d2[:, 0])
    assert np.allclose(M.inner_radius(tri, M.TRIANGLE), 2.0 * area / (a + b + c))                    # inscribed circle
    # a rhombus with unit sides and a 60 degree angle: inscribed circle radius = height / 2
    rh = np.array([[[0.0, 0.0], [1.0, 0.0], [1.5, np.sqrt(3) / 2], [0.5, np.sqrt(3) / 2]]])
    assert np.allclose(M._quad_inner_radius(rh), np.sqrt(3) / 4)


def test_node_tags_count_shared_nodes_once():
    mesh = M.box(2, (3, 2), 0.0, 1.0)
    tags, n = M.node_tags(mesh)
    assert n == 4 * 3 and tags[M.QUADRANGLE].shape == (6, 4)
    X = np.asarray(mesh.blocks[M.QUADRANGLE]["coords"])
    uniq = {}
    for e in range(6):
        for k in range(4):
            uniq.setdefault(int(tags[M.QUADRANGLE][e, k]), []).append(tuple(X[e, k]))
    assert all(len(set(v)) == 1 for v in uniq.values())          # one coordinate per tag
    per = M.periodic_box(2, 4)                                     # the two copies of a periodic pair stay distinct nodes, as in Gmsh
    assert M.node_tags(per)[1] == 5 * 5


def _jump(x):
    s = np.tanh((x[..., 0] - 0.5) / 0.02)
    rho = 0.5625 - 0.4375 * s
    return np.stack([rho, np.zeros_like(rho), 1.4 * (0.55 - 0.45 * s) / rho], axis=-1)


def test_indicator_regimes_and_node_maximum():
    mesh = M.box(1, (24,), 0.0, 1.0)
    O = oracle.Oracle(dict(p=2, conv_flux=2, rk=2, av_tolerance=0.5, av_factor=1.5), mesh)
    O.initialize(_jump, lambda x, phys, time=None: _jump(x))
    O.update_artificial_viscosity()
    node, elem = O.node_artificial_viscosity(), O.element_artificial_viscosity(M.LINE)
    full = 1.5 * (1.0 / 24 / 2) / 2          # factor * inner_radius / p  (SpatialDiscrete.cpp:75-78)
    assert node.max() <= full * (1 + 1e-14) and node.max() > 0.0 and (node == 0.0).sum() > 15
    tags, _ = M.node_tags(mesh)
    assert np.array_equal(elem, node[tags[M.LINE]])                       # storeElementArtificialViscosity: corner values = node values
    flagged = np.nonzero(elem.max(axis=1) > 0)[0]
    assert flagged.min() >= 9 and flagged.max() <= 14                     # around the jump at x = 0.5, plus the neighbours that share its nodes
    # P1: every mode is "high": indicator = log10(1) = the P1 threshold 0 -> half of the full value everywhere
    O1 = oracle.Oracle(dict(p=1, conv_flux=2, rk=2, av_tolerance=0.5, av_factor=1.0), mesh)
    O1.initialize(_jump, lambda x, phys, time=None: _jump(x)); O1.update_artificial_viscosity()
    assert np.allclose(O1.node_artificial_viscosity(), 0.5 * (1.0 / 24 / 2) / 1, rtol=1e-12)


def test_viscosity_is_conservative_and_dissipative():
    """periodic box: the eps * grad(U) terms telescope over the faces (mass, momentum, energy stay constant to round-off) and the jump is
    smeared (the L2 norm of the density ends below the inviscid run's)"""
    mesh = M.periodic_box(2, 8)

    def ic(x):
        s = np.tanh((np.abs(x[..., 0] - 1.0) - 0.5) / 0.04)
        rho = 0.75 - 0.25 * s
        return np.stack([rho, 0.2 + 0 * rho, 0 * rho, 1.4 * (0.75 - 0.25 * s) / rho], axis=-1)
    runs = {}
    for av in (None, 1.0):
        cfg = dict(p=2, conv_flux=2, rk=2)
        if av:
            cfg.update(av_tolerance=av, av_factor=1.0)
        O = oracle.Oracle(cfg, mesh); O.initialize(ic)
        t = O.types[0]
        w = O.element_geometry(t, 1)                                     # detJ * w
        tot0 = (O.state_at_quadrature(t) * w[..., None]).sum(axis=(0, 1))
        O.step(0.1 * O.compute_dt(1.0), 10)
        u = O.state_at_quadrature(t)
        assert np.allclose((u * w[..., None]).sum(axis=(0, 1)), tot0, rtol=1e-12, atol=1e-13)
        runs[av] = float((u[..., 0] ** 2 * w).sum())
        if av:
            assert O.node_artificial_viscosity().max() > 0.0
    assert runs[1.0] < runs[None] * (1 - 1e-6)                            # diffusion of the density: its L2 norm drops below the inviscid run's
