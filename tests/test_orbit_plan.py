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


def emulate_kernel_stores(plan, nside, n, M, bounds=None, rank=0):
    """What tquOrbitKernel stores for rank `rank` of the partition `bounds` (in-face column ranges), entry by entry, taking the
    nine values of a source pair from M.  Returns the values and store counts per packed position for the rank's own strips, the
    rank's outbox (values + store count per element, laid out as cmg_orbit_shard says) and, per outbox element, the packed
    position it belongs to (what orbitInboxScatterKernel computes on the receiving side)."""
    from cosmopp_b200 import partition
    F = nside * nside
    bounds = [0, F] if bounds is None else list(bounds)
    q0, q1 = bounds[rank], bounds[rank + 1]
    offsets = partition.orbit_outbox_offsets(plan, bounds, rank)
    ld = q1 - q0
    dim = 3 * n
    size = capi.packed_size(dim)
    out = np.full(size, np.nan)
    count = np.zeros(size, dtype=np.int32)
    box_out = np.full(offsets[-1], np.nan)
    box_count = np.zeros(offsets[-1], dtype=np.int32)
    box_pos = np.full(offsets[-1], -1, dtype=np.int64)

    def po(col):
        return col * (col + 1) // 2

    def owned(col):                      # packed column -> pixel -> in-face index inside [q0, q1)
        q = (col % n) % F
        return (q >= q0) & (q < q1)

    def put(pos, col, val, live):
        assert owned(col[live]).all()    # everything stored in the strips lies in this rank's own packed columns
        np.add.at(count, pos[live], 1)
        out[pos[live]] = val[live]

    il = np.arange(TI)[:, None]
    jl = np.arange(TJ)[None, :]
    for c in plan:
        kinds_before = 0
        for tr in range(F // TI):
            for tc in range(ld // TJ):
                q_row0, q_col0 = tr * TI, q0 + tc * TJ
                if c["tri"] and q_row0 > q_col0 + TJ - 1:
                    continue
                a = c["row_face"] * F + q_row0 + il + 0 * jl
                b = c["col_face"] * F + q_col0 + jl + 0 * il
                dq = (q_col0 + jl) - (q_row0 + il)
                v = {(X, Y): M[X * n + a, Y * n + b] for X in range(3) for Y in range(3)}
                for k, (fr, fc, swap) in enumerate(c["images"]):
                    ip = fr * F + q_row0 + il + 0 * jl
                    jp = fc * F + q_col0 + jl + 0 * il
                    cols = [s * n + jp for s in range(3)]
                    col = [po(x) for x in cols]                                   # sColPtr
                    direct = ((not c["tri"]) | (dq >= 0)) & (ip >= 0)
                    if not swap:
                        put(col[0] + ip, cols[0], v[0, 0], direct)
                        put(col[1] + ip, cols[1], v[0, 1], direct)
                        put(col[1] + n + ip, cols[1], v[1, 1], direct)
                        put(col[2] + ip, cols[2], v[0, 2], direct)
                        put(col[2] + n + ip, cols[2], v[1, 2], direct)
                        put(col[2] + 2 * n + ip, cols[2], v[2, 2], direct)
                    else:
                        strict = ((not c["tri"]) | (dq > 0)) & (ip >= 0)
                        put(col[1] + ip, cols[1], v[0, 1], strict)
                        put(col[2] + ip, cols[2], v[0, 2], strict)
                        put(col[2] + n + ip, cols[2], v[1, 2], strict)
                    # (mirror images: below, after the loop over the rotation images)
                    min_gap = -(1 << 30) if not c["tri"] else (1 if (c["same_face"] or swap) else 0)
                    staged = [(1, 0), (2, 0), (2, 1)] + ([(0, 0), (1, 1), (2, 2)] if swap else [])
                    qa = q_row0 + il + 0 * jl
                    local = (qa >= q0) & (qa < q1)
                    for t, (X, Y) in enumerate(staged):
                        live = dq >= min_gap
                        pos = po(X * n + ip) + Y * n + jp
                        put(pos, X * n + ip, v[X, Y], live & local)
                        far = live & ~local
                        if far.any():
                            e = partition.orbit_outbox_index(bounds, rank, offsets, c["combo_base"][k] + t, qa[far], (q_col0 + jl + 0 * il)[far])
                            np.add.at(box_count, e, 1)
                            box_out[e] = v[X, Y][far]
                            box_pos[e] = pos[far]
                # mode 3: the four rotations of the pair's image under the meridian mirror -- both in-face indices with their even
                # and odd bits swapped, entries with exactly one U index negated (single owner; whole face pairs of different rings)
                for fr, fc in c.get("mirror_images", []):
                    assert not c["tri"] and world_is_one(bounds)
                    ip = fr * F + swapbits(q_row0 + il + 0 * jl)
                    jp = fc * F + swapbits(q_col0 + jl + 0 * il)
                    assert (ip < jp).all()
                    every = np.ones((TI, TJ), dtype=bool)
                    for X in range(3):
                        for Y in range(3):
                            sign = -1.0 if (X == 2) != (Y == 2) else 1.0
                            if X <= Y:                       # direct stores: column (Y b'), row (X a')
                                put(po(Y * n + jp) + X * n + ip, Y * n + jp, sign * v[X, Y], every)
                            else:                            # staged: column (X a'), row (Y b')
                                put(po(X * n + ip) + Y * n + jp, X * n + ip, sign * v[X, Y], every)
    return out, count, box_out, box_count, box_pos


def swapbits(q):
    return ((q & 0x55555555) << 1) | ((q & 0xAAAAAAAA) >> 1)


def world_is_one(bounds):
    return len(bounds) == 2


def emulate_inbox_scatter(plan, nside, n, bounds, sender, receiver, block):
    """orbitInboxScatterKernel: packed position of every element of block(sender -> receiver), from the layout alone"""
    from cosmopp_b200 import partition
    F = nside * nside
    S = partition.ORB_SUB
    nct = (bounds[sender + 1] - bounds[sender]) // S
    h0, nh = bounds[receiver] // S, (bounds[receiver + 1] - bounds[receiver]) // S
    pos = np.full(block.size, -1, dtype=np.int64)
    X_of, Y_of = [1, 2, 2, 0, 1, 2], [0, 0, 1, 0, 1, 2]
    r = np.arange(S)[:, None] + 0 * np.arange(S)[None, :]
    l = np.arange(S)[None, :] + 0 * np.arange(S)[:, None]
    for c in plan:
        if c["tri"] and not receiver < sender:
            continue
        for k, (fr, fc, swap) in enumerate(c["images"]):
            for t in range(6 if swap else 3):
                combo = c["combo_base"][k] + t
                for ct in range(nct):
                    for hh in range(nh):
                        a = fr * F + (h0 + hh) * S + r
                        b = fc * F + bounds[sender] + ct * S + l
                        col = X_of[t] * n + a
                        e = ((combo * nct + ct) * nh + hh) * S * S + r * S + l
                        pos[e] = col * (col + 1) // 2 + Y_of[t] * n + b
    return pos


def packed_from_full(M):
    dim = M.shape[0]
    iu = np.triu_indices(dim)
    want = np.empty(capi.packed_size(dim))
    want[iu[1] * (iu[1] + 1) // 2 + iu[0]] = M[iu]
    return want, iu


def test_mirror_plan_evaluates_thirteen_units():
    plan = capi.orbit_plan(8, 3)
    units = sum(0.5 if c["tri"] else 1.0 for c in plan)
    assert units == 13.0
    with_mirror = [c for c in plan if c["mirror_images"]]
    assert len(with_mirror) == 5 and all(not c["tri"] and len(c["mirror_images"]) == 4 for c in with_mirror)
    assert all(not c["mirror_images"] for c in capi.orbit_plan(8, 0))


@pytest.mark.parametrize("mode", [0, 1, 3])
def test_store_rules_fill_the_packed_triangle_exactly_once(oracle_matrix, mode):
    nside, n, M = oracle_matrix
    out, count, _, box_count, _ = emulate_kernel_stores(capi.orbit_plan(nside, mode), nside, n, M)
    assert count.min() == 1 and count.max() == 1 and box_count.size == 0
    want, iu = packed_from_full(M)
    # images take the source pair's value: equal to the oracle's own entry up to the oracle's rounding
    scale = np.empty_like(want)
    scale[iu[1] * (iu[1] + 1) // 2 + iu[0]] = np.where(iu[1] < n, M[0, 0], M[n, n])
    assert (np.abs(out - want) / scale).max() < 1e-11


def check_sharded_store_rules(nside, n, M, mode, world, values=True):
    """Ranks own in-face column ranges of all twelve faces: together their strips and outboxes hold every entry once, strip
    stores stay inside the rank's own packed columns, every element of every (compact) outbox is written exactly once, and the
    receiver-side scatter puts each destination block where the sender's stores belong -- inside the receiver's columns."""
    from cosmopp_b200 import partition
    F = nside * nside
    plan = capi.orbit_plan(nside, mode)
    bounds = partition.orbit_partition(nside, world, mode)
    assert bounds[0] == 0 and bounds[-1] == F and all(b % 32 == 0 for b in bounds)
    size = capi.packed_size(3 * n)
    total = np.zeros(size, dtype=np.int32)
    merged = np.full(size, np.nan)
    pairs = 0
    for r in range(world):
        out, count, box_out, box_count, box_pos = emulate_kernel_stores(plan, nside, n, M, bounds, r)
        offsets = partition.orbit_outbox_offsets(plan, bounds, r)
        assert offsets == capi.orbit_outbox_layout(nside, mode, bounds, r)
        assert box_count.size == offsets[-1] and (box_count.size == 0 or (box_count.min() == 1 and box_count.max() == 1))
        for d in range(world):
            if d == r:
                assert offsets[d + 1] == offsets[d]
                continue
            pos = emulate_inbox_scatter(plan, nside, n, bounds, r, d, box_out[offsets[d]:offsets[d + 1]])
            assert (pos == box_pos[offsets[d]:offsets[d + 1]]).all()
            # the packed column of every element of the block is one of rank d's
            col = np.floor((np.sqrt(8.0 * pos + 1) - 1) / 2).astype(np.int64)
            col -= (col * (col + 1) // 2 > pos)
            q = (col % n) % F
            assert ((q >= bounds[d]) & (q < bounds[d + 1])).all()
        total += count
        np.add.at(total, box_pos, 1)
        merged = np.where(count > 0, out, merged)
        merged[box_pos] = box_out
        pairs += partition.orbit_pairs_in_range(bounds[r], bounds[r + 1], F, mode)
    assert total.min() == 1 and total.max() == 1
    if values:
        want, _ = packed_from_full(M)
        assert np.abs(merged - want).max() < 1e-11 * M[n, n]
    units = 18.0 if mode == 0 else 22.5
    assert abs(pairs - units * F * F) <= 6 * F          # the q_row <= q_col classes include their diagonal


@pytest.mark.parametrize("mode,world", [(0, 2), (1, 2), (0, 3)])
def test_sharded_store_rules_partition_the_triangle(oracle_matrix, mode, world):
    nside, n, M = oracle_matrix
    check_sharded_store_rules(nside, n, M, mode, world)


def emulate_tt_stores(plan, nside, M, q0, q1, out, count):
    """legendreSeriesOrbitKernel: 128 x 16 tiles over the columns [q0, q1) of the plan; a straight image is one direct store per
    column, a transposed image puts entry (a', b') into column a' (rows = the image of the tile's column pixels)"""
    F = nside * nside
    rows, cols = 128, 16                  # TT_ROWS, TT_COLS
    il = np.arange(rows)[:, None]
    jl = np.arange(cols)[None, :]
    for c in plan:
        for tr in range(F // rows):
            for tc in range((q1 - q0) // cols):
                q_row0, q_col0 = tr * rows, q0 + tc * cols
                if c["tri"] and q_row0 > q_col0 + cols - 1:
                    continue
                dq = (q_col0 + jl) - (q_row0 + il)
                val = M[c["row_face"] * F + q_row0 + il + 0 * jl, c["col_face"] * F + q_col0 + jl + 0 * il]
                for fr, fc, swap in c["images"]:
                    ip = fr * F + q_row0 + il + 0 * jl
                    jp = fc * F + q_col0 + jl + 0 * il
                    if not swap:
                        live = np.broadcast_to((not c["tri"]) | (dq >= 0), (rows, cols))
                        pos = jp * (jp + 1) // 2 + ip
                    else:
                        live = np.broadcast_to((not c["tri"]) | (dq > 0), (rows, cols))
                        assert (ip[live] > jp[live]).all()             # the row image has the larger pixel index
                        pos = ip * (ip + 1) // 2 + jp
                    np.add.at(count, pos[live], 1)
                    out[pos[live]] = val[live]


@pytest.fixture(scope="module")
def tt_matrix16():
    nside, lmax = 16, 20
    n = 12 * nside * nside
    return nside, n, api.unpack_symmetric(api.cl_to_cmatrix(synthetic_cl(lmax), nside, 10.0), n)


def test_tt_orbit_store_rules_fill_the_triangle_exactly_once(tt_matrix16):
    """one owner: the plan WITH transposed images (18 units), every entry of the triangle stored exactly once"""
    nside, n, M = tt_matrix16
    size = capi.packed_size(n)
    out = np.full(size, np.nan)
    count = np.zeros(size, dtype=np.int32)
    plan = capi.orbit_plan(nside, 0)
    assert any(swap for c in plan for _, _, swap in c["images"])
    emulate_tt_stores(plan, nside, M, 0, nside * nside, out, count)
    assert count.min() == 1 and count.max() == 1
    want, _ = packed_from_full(M)
    assert np.abs(out - want).max() < 1e-11 * M[0, 0]


@pytest.mark.parametrize("world", [2, 3])
def test_tt_orbit_sharded_store_rules(tt_matrix16, world):
    """several ranks: the plan WITHOUT transposed images; a rank's stores stay inside its own packed columns (no exchange) and
    the ranks together fill the triangle once"""
    from cosmopp_b200 import partition
    nside, n, M = tt_matrix16
    F = nside * nside
    size = capi.packed_size(n)
    plan = capi.orbit_plan(nside, 1)
    assert not any(swap for c in plan for _, _, swap in c["images"])
    bounds = partition.orbit_partition(nside, world, 1, align=16)
    total = np.zeros(size, dtype=np.int32)
    merged = np.full(size, np.nan)
    for r in range(world):
        out = np.full(size, np.nan)
        count = np.zeros(size, dtype=np.int32)
        emulate_tt_stores(plan, nside, M, bounds[r], bounds[r + 1], out, count)
        pos = np.nonzero(count)[0]
        col = np.floor((np.sqrt(8.0 * pos + 1) - 1) / 2).astype(np.int64)
        col -= (col * (col + 1) // 2 > pos)
        q = col % F
        assert ((q >= bounds[r]) & (q < bounds[r + 1])).all()
        total += count
        merged = np.where(count > 0, out, merged)
    assert total.min() == 1 and total.max() == 1
    want, _ = packed_from_full(M)
    assert np.abs(merged - want).max() < 1e-11 * M[0, 0]


def test_orbit_plan_rejects_bad_arguments():
    with pytest.raises(capi.CmgError):
        capi.orbit_plan(12)                 # nside must be a power of two
    with pytest.raises(capi.CmgError):
        capi.orbit_plan(16, mode=2)
    for mode in (0, 1):
        plan = capi.orbit_plan(8, mode)
        assert len(plan) <= 24 and all(1 <= len(c["images"]) <= 4 for c in plan)


def test_mirror_symmetries_of_the_oracle_matrix(oracle_matrix):
    """The rest of the grid's symmetry group (not used by the kernels yet; DESIGN.md, open items): the equatorial mirror
    (north face f <-> f + 8, in-face (ix, iy) -> (N-1-iy, N-1-ix)) and the meridian mirror phi -> pi/2 - phi (polar face position
    p -> -p, equatorial p -> 1 - p, ix <-> iy).  A reflection flips the sign of every entry with exactly one U index."""
    nside, n, M = oracle_matrix
    F = nside * nside

    def swapbits(q):
        return (((q & 0x55555555) << 1) | ((q & 0xAAAAAAAA) >> 1)) & (F - 1)

    pix = np.arange(n)
    face, q = pix // F, pix % F
    equatorial = np.where(face < 4, face + 8, np.where(face >= 8, face - 8, face)) * F + swapbits(~q & (F - 1))
    ring, pos = face // 4, face % 4
    meridian = (ring * 4 + np.where(ring == 1, (1 - pos) % 4, (-pos) % 4)) * F + swapbits(q)
    v = api.unit_vectors(nside).reshape(n, 3)
    assert np.abs(v[equatorial] - v * np.array([1.0, 1.0, -1.0])).max() < 1e-15
    assert np.abs(v[meridian] - v[:, [1, 0, 2]]).max() < 2e-15
    sign = np.concatenate([np.ones(n), np.ones(n), -np.ones(n)])
    for perm in (equatorial, meridian):
        assert np.array_equal(perm[perm], pix)                       # involutions
        idx = np.concatenate([perm, n + perm, 2 * n + perm])
        image = M[np.ix_(idx, idx)] * np.outer(sign, sign)
        assert np.abs(image - M)[:n].max() < 1e-11 * M[0, 0]
        assert np.abs(image - M)[n:].max() < 1e-11 * M[n, n]


@pytest.mark.parametrize("threads", [1, 3])
def test_host_expansion_rebuilds_the_matrix_from_the_last_face_columns(oracle_matrix, threads):
    """cmg_host_expand_rotations (pure CPU): keep only the columns of base faces 3, 7, 11 of every strip, the rest comes back
    as rotated images."""
    nside, n, M = oracle_matrix
    F = nside * nside
    want, iu = packed_from_full(M)
    col = np.empty(want.size, dtype=np.int64)
    col[iu[1] * (iu[1] + 1) // 2 + iu[0]] = iu[1]
    kept = ((col % n) // F) % 4 == 3
    assert 0.25 < kept.mean() < 0.29                               # 27 % of the entries cross PCIe
    got = np.where(kept, want, np.nan)
    capi.host_expand_rotations(got, nside, 3, threads)
    assert not np.isnan(got).any()
    assert np.array_equal(got[kept], want[kept])
    scale = np.where(col < n, M[0, 0], M[n, n])
    assert (np.abs(got - want) / scale).max() < 1e-11
    # TT alone (one strip)
    tt = want[:capi.packed_size(n)].copy()
    got = np.where(kept[:tt.size], tt, np.nan)
    capi.host_expand_rotations(got, nside, 1, threads)
    assert (np.abs(got - tt) / M[0, 0]).max() < 1e-11
    # a piece at a time, as the whole call does while later pieces are still in flight
    got = np.where(kept, want, np.nan)
    L = capi.library()
    for strip in range(3):
        for ring in range(3):
            assert L.cmg_host_expand_rotations(got.ctypes.data, nside, strip, strip + 1, 4 * ring, 4 * ring + 4, threads) == 0
    assert (np.abs(got - want) / scale).max() < 1e-11
    with pytest.raises(capi.CmgError):
        capi.host_expand_rotations(got, 12, 3)


@pytest.fixture(scope="module")
def oracle_matrix16():
    nside, lmax = 16, 12
    n = 12 * nside * nside
    return nside, n, api.unpack_symmetric(api.tqu_matrix(*synthetic_cl(lmax, pol=True), nside, 10.0), 3 * n)


@pytest.mark.parametrize("world,mode", [(2, 0), (3, 0), (3, 1)])
def test_sharded_store_rules_at_the_gpu_test_sizes(oracle_matrix16, world, mode):
    """the configurations tests/test_gpu_orbit.py runs on the device (Nside=16)"""
    nside, n, M = oracle_matrix16
    check_sharded_store_rules(nside, n, M, mode, world)
