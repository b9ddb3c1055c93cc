"""Host plan of the trace-based line kernels (nsl_kernels.cuh, P3 hexahedra): the link records and the face-point correspondence
tables must put every face point of every element on the SAME physical point as the partner entry they name — checked against the
mesh geometry for every (local face, local face, rotation) combination that the structured, cubed-sphere and 26+6-block sphere
meshes produce.  CPU only (plan-only context)."""
import numpy as np
import pytest

from subrosadg_b200 import mesh as M
from subrosadg_b200 import solver as sv

NS = dict(p=3, model=1, transport=1, mu=1e-3, visc_flux=2)
FACE_DIR = [2, 1, 0, 0, 1, 2]      # hexahedron faces (zeta-, eta-, xi-, xi+, eta+, zeta+)
FACE_SIDE = [0, 0, 0, 1, 1, 1]


def face_point_coordinates(S, t):
    """[n_internal][6][16][3]: coordinates of the face points in natural order (the two tangential lattice indices, lower axis first),
    by end-point interpolation of the volume quadrature coordinates along the normal lines (exact for geometry of order <= 3)."""
    xq = S.quadrature_coordinates(t)                       # caller order [n][64][3], node = i*16 + j*4 + k
    perm = S.debug_plan(10)
    x = np.empty_like(xq); x[perm] = xq                    # internal order
    lend = S.debug_plan(6).reshape(2, 4)
    X = x.reshape(-1, 4, 4, 4, 3)
    out = np.zeros((x.shape[0], 6, 16, 3))
    for f in range(6):
        tr = np.tensordot(lend[FACE_SIDE[f]], np.moveaxis(X, 1 + FACE_DIR[f], 0), axes=(0, 0))   # remaining lattice axes in order
        out[:, f] = tr.reshape(-1, 16, 3)
    return out


@pytest.mark.parametrize("name,make", [("periodic", lambda: M.periodic_box_fast(3, 4)), ("shell", lambda: M.cubed_sphere_shell(3, 2)),
                                       ("sphere_in_box", lambda: M.sphere_in_box(3, 3, 2)), ("warped", lambda: M.box(3, (3, 2, 2), 0.0, 1.0, geom_order=2, warp=lambda x: x + 0.03 * np.sin(np.pi * np.roll(x, 1, axis=1))))])
def test_links_and_partner_tables_match_geometry(built, name, make):
    m = make()
    S = sv.Solver(NS, m, device=-1)
    t = S.types[0]
    affine, K, nch, nown = S.debug_plan(15)
    assert K == 8
    links = S.debug_plan(20).reshape(nown, 6, 4)
    partner = S.debug_plan(21).reshape(6, 6, 4, 2, 16)
    xf = face_point_coordinates(S, t)
    span = np.ptp(S.quadrature_coordinates(t).reshape(-1, 3), axis=0) if name == "periodic" else None
    n_int = int(m.faces["n_int"])
    seen = set()
    for e in range(nown):
        for f in range(6):
            other, face_id, z, _ = links[e, f]
            lfo, rot, bc, am_r, handles, in_chunk = z & 7, (z >> 3) & 3, (z >> 5) & 7, (z >> 8) & 1, (z >> 9) & 1, (z >> 10) & 1
            if other < 0:
                assert face_id >= n_int and bc == m.faces["bc"][face_id]
                continue
            seen.add((f, lfo, rot))
            assert face_id < n_int and in_chunk == int(other // K == e // K)
            assert handles == (0 if (am_r and in_chunk) else 1)       # a face inside a block is evaluated by its left parent only
            back = links[other, lfo]
            assert back[0] == e and (back[2] & 7) == f and ((back[2] >> 8) & 1) == 1 - am_r and back[1] == face_id
            d = xf[e, f] - xf[other, lfo][partner[f, lfo, rot, am_r]]
            if span is not None:
                d -= np.round(d / 2.0) * 2.0                            # periodic images of [0, 2]^3
            assert np.abs(d).max() < 1e-12, (e, f, other, lfo, rot, am_r)
    assert len(seen) >= (3 if name == "periodic" else 6)
    # boundary records
    nb = int(m.faces["n_bnd"])
    if nb:
        rec = S.debug_plan(23).reshape(-1, 4)[:nb]
        perm = S.debug_plan(10)
        assert np.array_equal(rec[:, 0], perm[m.faces["le"][n_int:]]) and np.array_equal(rec[:, 1], m.faces["lf"][n_int:])
        assert np.array_equal(rec[:, 3], np.arange(n_int, n_int + nb))


def test_left_point_table_is_the_reference_face_point_order(built):
    """jLeft names the column of the face geometry / boundary_dummy_variable_ arrays (reference face-point order of the left parent)"""
    m = M.box(3, (2, 2, 2), 0.0, 1.0, geom_order=2, warp=lambda x: x + 0.03 * np.sin(np.pi * np.roll(x, 1, axis=1)))
    S = sv.Solver(NS, m, device=-1)
    t = S.types[0]
    jleft = S.debug_plan(22).reshape(6, 4, 2, 16)
    xf = face_point_coordinates(S, t)
    xb = S.boundary_quadrature_coordinates()                 # [nBnd][16][3] in the reference order of the left parent's face
    rec = S.debug_plan(23).reshape(-1, 4)
    for fb in range(int(m.faces["n_bnd"])):
        e, f = rec[fb, 0], rec[fb, 1]
        assert np.abs(xf[e, f] - xb[fb][jleft[f, 0, 0]]).max() < 1e-12
