"""Host logic of the experimental symmetry-orbit path (cmg_tqu_orbit): the classes of base-face pairs and the
store rules of tquOrbitKernel (cosmopp_b200/csrc/orbit.cuh), restated in numpy and run against the oracle's
matrix.  No GPU: this pins the index arithmetic -- every packed entry of the [T;Q;U] triangle is written exactly
once, with the value the oracle has there -- and the symmetry the path rests on."""
import numpy as np
import pytest

from cosmopp_b200 import capi
from cosmopp_b200.synthetic import synthetic_cl
from oracle import api

TI, TJ = 64, 32          # PQ_TI, PQ_TJ


def rotate(pix, face_pix, k=1):
    f, q = pix // face_pix, pix % face_pix
    return ((f & ~3) | ((f + k) & 3)) * face_pix + q


@pytest.fixture(scope="module")
def oracle_matrix():
    nside, lmax = 8, 14
    spectra = synthetic_cl(lmax, pol=True)
    n = 12 * nside * nside
    return nside, n, api.unpack_symmetric(api.tqu_matrix(*spectra, nside, 10.0), 3 * n)


def test_rotation_symmetry_of_the_oracle_matrix(oracle_matrix):
    """C[X Ra, Y Rb] = C[X a, Y b]: frames rotate with the pixels under the pi/2 rotation about the pole."""
    nside, n, M = oracle_matrix
    R = rotate(np.arange(n), nside * nside)
    v = api.unit_vectors(nside).reshape(n, 3)
    assert np.abs(v[R][:, 0] + v[:, 1]).max() < 1e-15 and np.abs(v[R][:, 1] - v[:, 0]).max() < 1e-15
    assert np.array_equal(v[R][:, 2], v[:, 2])
    idx = np.concatenate([R, n + R, 2 * n + R])
    for block, diag in ((slice(0, n), M[0, 0]), (slice(n, 3 * n), M[n, n])):
        assert np.abs(M[np.ix_(idx, idx)][block] - M[block]).max() < 1e-11 * diag


@pytest.mark.parametrize("mode", [0, 1])
def test_plan_covers_every_face_pair_once(mode):
    plan = capi.orbit_plan(16, mode)
    seen = {}
    units = 0.0
    for c in plan:
        assert c["row_face"] <= c["col_face"] and c["images"][0][:2] == (c["row_face"], c["col_face"])
        assert c["same_face"] == (c["row_face"] == c["col_face"])
        assert not c["same_face"] or c["tri"]
        units += 0.5 if c["tri"] else 1.0
        for k, (fr, fc, swap) in enumerate(c["images"]):
            assert (fr >> 2, fc >> 2) == (c["row_face"] >> 2, c["col_face"] >> 2)
            assert ((fr - c["row_face"]) & 3) == ((fc - c["col_face"]) & 3)       # one rotation moves both
            assert swap == (fr > fc)                                               # transposed: row image has the larger index
            assert mode == 0 or not swap
            key = (min(fr, fc), max(fr, fc))
            # a whole face pair is covered once; a q_row <= q_col class covers it with a straight and a transposed image
            seen[key] = seen.get(key, 0.0) + (0.5 if (c["tri"] and not c["same_face"]) else 1.0)
    assert seen == {(a, b): 1.0 for a in range(12) for b in range(a, 12)}
    assert units == (18.0 if mode == 0 else 22.5)                                  # of 72 face-pair units


def emulate_kernel_stores(plan, nside, n, M):
    """What tquOrbitKernel stores, entry by entry, taking the nine values of a source pair from M."""
    F = nside * nside
    dim = 3 * n
    out = np.full(capi.packed_size(dim), np.nan)
    count = np.zeros(capi.packed_size(dim), dtype=np.int32)

    def po(col):
        return col * (col + 1) // 2

    def put(pos, val, live):
        np.add.at(count, pos[live], 1)
        out[pos[live]] = val[live]

    il = np.arange(TI)[:, None]
    jl = np.arange(TJ)[None, :]
    for c in plan:
        for tr in range(F // TI):
            for tc in range(F // TJ):
                q_row0, q_col0 = tr * TI, tc * TJ
                if c["tri"] and q_row0 > q_col0 + TJ - 1:
                    continue
                a = c["row_face"] * F + q_row0 + il + 0 * jl
                b = c["col_face"] * F + q_col0 + jl + 0 * il
                dq = (q_col0 + jl) - (q_row0 + il)
                v = {(X, Y): M[X * n + a, Y * n + b] for X in range(3) for Y in range(3)}
                for fr, fc, swap in c["images"]:
                    ip = fr * F + q_row0 + il + 0 * jl
                    jp = fc * F + q_col0 + jl + 0 * il
                    col = [po(s * n + jp) for s in range(3)]                      # sColPtr
                    direct = (not c["tri"]) | (dq >= 0)
                    if not swap:
                        put(col[0] + ip, v[0, 0], direct)
                        put(col[1] + ip, v[0, 1], direct)
                        put(col[1] + n + ip, v[1, 1], direct)
                        put(col[2] + ip, v[0, 2], direct)
                        put(col[2] + n + ip, v[1, 2], direct)
                        put(col[2] + 2 * n + ip, v[2, 2], direct)
                    else:
                        strict = (not c["tri"]) | (dq > 0)
                        put(col[1] + ip, v[0, 1], strict)
                        put(col[2] + ip, v[0, 2], strict)
                        put(col[2] + n + ip, v[1, 2], strict)
                    min_gap = -(1 << 30) if not c["tri"] else (1 if (c["same_face"] or swap) else 0)
                    staged = [(1, 0), (2, 0), (2, 1)] + ([(0, 0), (1, 1), (2, 2)] if swap else [])
                    for X, Y in staged:
                        put(po(X * n + ip) + Y * n + jp, v[X, Y], dq >= min_gap)
    return out, count


@pytest.mark.parametrize("mode", [0, 1])
def test_store_rules_fill_the_packed_triangle_exactly_once(oracle_matrix, mode):
    nside, n, M = oracle_matrix
    out, count = emulate_kernel_stores(capi.orbit_plan(nside, mode), nside, n, M)
    assert count.min() == 1 and count.max() == 1
    dim = 3 * n
    iu = np.triu_indices(dim)
    want = np.empty(capi.packed_size(dim))
    want[iu[1] * (iu[1] + 1) // 2 + iu[0]] = M[iu]
    # images take the source pair's value: equal to the oracle's own entry up to the oracle's rounding
    scale = np.where(iu[1] < n, M[0, 0], M[n, n])
    err = np.empty(capi.packed_size(dim))
    err[iu[1] * (iu[1] + 1) // 2 + iu[0]] = scale
    assert (np.abs(out - want) / err).max() < 1e-11
