"""CPU tests of the product's host logic (no GPU): mesh flattening, metric terms, face geometry, chunk face lists and the
C ABI surface.  Geometry is checked against the oracle's independent restatement of src/Mesh/Geometry.cpp."""
import ctypes
import os
import re

import numpy as np
import pytest

import oracle
from subrosadg_b200 import mesh as M
from subrosadg_b200 import solver as sv

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built):
    hdr = open(os.path.join(ROOT, "include", "subrosadg_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(sdg_[a-z_0-9]+)\s*\(", hdr)))
    lib = sv.load_library()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/subrosadg_b200.h but not exported"
    assert sorted(sv.EXPORTS) == declared
    assert lib.sdg_version() >= 100


def test_compute_fails_loudly_without_a_device(built):
    """No CPU fallback: a plan-only context refuses every compute entry point; a device context cannot be created when
    CUDA is unavailable."""
    m = M.periodic_box(2, 4)
    S = sv.Solver(dict(p=2), m, device=-1)
    with pytest.raises(RuntimeError, match="no CUDA device|CPU path"):
        S.stepSolver(1e-3, 1)
    with pytest.raises(RuntimeError, match="no CUDA device|CPU path"):
        S.initializeSolver(lambda x: np.ones(x.shape[:-1] + (4,)))
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            sv.Solver(dict(p=2), m, device=0)


def test_rejects_bad_input(built):
    m = M.periodic_box(2, 4)
    with pytest.raises(RuntimeError):
        sv.Solver(dict(p=7), m, device=-1)
    with pytest.raises(RuntimeError, match="BR1 or BR2"):
        sv.Solver(dict(p=2, model=1, visc_flux=0), m, device=-1)
    bad = M.periodic_box(2, 4)
    bad.faces["re"] = bad.faces["re"].copy(); bad.faces["re"][0] = 10 ** 6
    with pytest.raises(RuntimeError, match="out of range"):
        sv.Solver(dict(p=2), bad, device=-1)


MESHES = {
    "line1d": (lambda: M.box(1, (9,), 0.0, 2.0), 3),
    "line1d_periodic_curved": (lambda: M.box(1, (7,), 0.0, 2.0, periodic_axes=(0,), geom_order=3, warp=lambda x: x + 0.05 * np.sin(np.pi * x)), 4),
    "periodic2d": (lambda: M.periodic_box(2, 10), 3),
    "box2d_p5_curved": (lambda: M.box(2, (3, 3), 0, 1, geom_order=2, warp=lambda x: x + 0.03 * np.sin(np.pi * x[:, ::-1])), 5),
    "box3d_p4": (lambda: M.box(3, (2, 2, 2), 0, 1), 4),
    "periodic3d": (lambda: M.periodic_box(3, 4), 3),
    "periodic3d_fast": (lambda: M.periodic_box_fast(3, 5), 2),
    "box2d": (lambda: M.box(2, (5, 4), 0, 1), 3),
    "box2d_warped": (lambda: M.box(2, (5, 4), 0, 1, geom_order=3, warp=lambda x: x + 0.05 * np.sin(np.pi * x[:, ::-1])), 3),
    "box3d_warped": (lambda: M.box(3, (3, 3, 3), 0, 1, geom_order=2, warp=lambda x: x + 0.03 * np.sin(np.pi * np.roll(x, 1, axis=1))), 2),
    "naca": (lambda: M.naca0012(nr=6, nt=16), 3),
    "sphere": (lambda: M.cubed_sphere_shell(3, 3), 3),
    "annulus": (lambda: M.annulus(4, 12), 3),
}


@pytest.mark.parametrize("name", sorted(MESHES))
def test_flattened_geometry_matches_oracle(built, name):
    make, p = MESHES[name]
    m = make()
    O = oracle.Oracle(dict(p=p), m)
    S = sv.Solver(dict(p=p), m, device=-1)
    t = S.types[0]
    s, so = S.sizes(t), O.sizes(t)
    assert (s.n, s.Nb, s.Nq, s.Nf, s.Naq, s.nn, s.Nqf, s.Nv) == (so.n, so.Nb, so.Nq, so.Nf, so.Naq, so.nn, so.Nqf, so.Nv)
    D = m.dim
    tol = 2e-13
    assert np.abs(S.quadrature_coordinates(t) - O.quadrature_coordinates(t)).max() < tol
    affine, K, nch, nown = S.debug_plan(15)
    assert bool(affine) == (m.blocks[t]["geom_order"] == 1 and "warp" not in name)
    perm = S.debug_plan(10)
    assert sorted(perm.tolist()) == list(range(s.n))           # a permutation, invisible at the seam
    geoE, geoF = S.debug_plan(0), S.debug_plan(3)
    mt, jw, wq = O.element_geometry(t, 2), O.element_geometry(t, 1), O.table(t, 5)
    if affine:
        g = geoE.reshape(s.n, D * D + 1)[perm]
        mt2, jw2 = g[:, None, :D * D] * wq[None, :, None], g[:, D * D][:, None] * wq[None, :]
    else:
        mt2 = geoE.reshape(s.n, D * D, s.Nq)[perm].transpose(0, 2, 1)
        jw2 = 1.0 / S.debug_plan(1).reshape(s.n, s.Nq)[perm]
    scale = np.abs(mt).max()
    assert np.abs(mt2 - mt).max() < tol * max(scale, 1) and np.abs(jw2 / jw - 1).max() < 1e-12
    nrm, fjw, wf = O.face_geometry(1), O.face_geometry(2), O.table(t, 7)
    nf = nrm.shape[0]
    if affine:
        g = geoF.reshape(nf, D + 1)
        n2, j2 = np.broadcast_to(g[:, None, :D], nrm.shape), g[:, D][:, None] * wf[None, :]
    else:
        g = geoF.reshape(nf, D + 1, s.Nqf)
        n2, j2 = g[:, :D, :].transpose(0, 2, 1), g[:, D, :]
    assert np.abs(n2 - nrm).max() < 1e-12 and np.abs(j2 / fjw - 1).max() < 1e-12
    if m.faces["n_bnd"] > 0:
        assert np.abs(S.boundary_quadrature_coordinates() - O.boundary_quadrature_coordinates()).max() < tol
    assert np.abs(S.debug_plan(2)[perm] - O.element_geometry(t, 4)).max() == 0.0   # minEdge
    # per-chunk face lists: every face is listed once in the chunk of each of its owned parents
    off, rec = S.debug_plan(11), S.debug_plan(12).reshape(-1, 4)
    assert off[0] == 0 and off[-1] == len(rec) and len(off) == nch + 1
    f = m.faces
    n_int = int(f["n_int"])
    expect = {}
    for i in range(nf):
        cl = perm[f["le"][i]] // K
        expect.setdefault(cl, set()).add(i)
        if i < n_int:
            expect.setdefault(perm[f["re"][i]] // K, set()).add(i)
    for c in range(nch):
        got = rec[off[c]:off[c + 1]]
        assert sorted(got[:, 2].tolist()) == sorted(expect.get(c, set()))
        for eL, eR, i, packed in got:
            assert eL == perm[f["le"][i]] and (eR == (perm[f["re"][i]] if i < n_int else -1))
            assert (packed & 15) == f["lf"][i] and ((packed >> 12) & 15) == (f["bc"][i] & 15)
            if i < n_int:
                assert ((packed >> 4) & 15) == f["rf"][i] and ((packed >> 8) & 15) == f["rot"][i]
    assert sorted(S.debug_plan(13).tolist() + S.debug_plan(14).tolist()) == list(range(nch))


def test_internal_order_forms_bricks(built):
    """Morton order on a uniform box: consecutive groups of 8 hexes are 2x2x2 bricks (12 interior faces per chunk)."""
    m = M.periodic_box_fast(3, 8)
    S = sv.Solver(dict(p=3), m, device=-1)
    off = S.debug_plan(11)
    assert np.all(np.diff(off) == 36)
    S0 = sv.Solver(dict(p=3), m, device=-1, reorder=0)
    assert np.array_equal(S0.debug_plan(10), np.arange(8 ** 3))


def test_face_rotation_matches_geometry(built):
    """adjacency_right_rotation_ + the permutation table place the right parent's face points on the left parent's."""
    for m in (M.periodic_box(3, 4), M.cubed_sphere_shell(3, 2), M.box(3, (3, 2, 2), 0, 1)):
        assert oracle.Oracle(dict(p=3), m).check_face_match() < 1e-12
    for m in (M.periodic_box(2, 5), M.naca0012(nr=4, nt=12), M.annulus(3, 8, tri_rings=1)):
        assert oracle.Oracle(dict(p=3), m).check_face_match() < 1e-12
