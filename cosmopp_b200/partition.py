"""Host-side sharding of the packed upper triangle over GPUs.

Every entry depends only on replicated inputs (pixel frames, C_l), so the path shards with no
data-path collective: rank r computes all pixel pairs (i <= j) whose column j lies in its block of
pixel columns.  Column j holds j+1 pairs, so blocks of equal *area* have boundaries ~ N sqrt(k/G).
For TT the rank's output is one contiguous piece of the packed triangle; for [T;Q;U] it is the three
column strips j, N+j, 2N+j of its block (include/cmg.h, cmg_tqu_layout).
"""
import math


def packed_size(dim):
    return dim * (dim + 1) // 2


def column_partition(npix, n_parts, align=32):
    """Boundaries b[0..n_parts] (b[0]=0, b[-1]=npix) of equal-work pixel-column blocks, interior
    boundaries rounded to a multiple of `align` (the kernels' column tile)."""
    if n_parts < 1:
        raise ValueError("n_parts must be >= 1")
    b = [0]
    for k in range(1, n_parts):
        x = npix * math.sqrt(k / n_parts)
        v = int(round(x / align)) * align
        v = max(b[-1], min(npix, v))
        b.append(v)
    b.append(npix)
    return b


def pairs_in_block(begin, end):
    """pixel pairs (i <= j) with begin <= j < end"""
    return packed_size(end) - packed_size(begin)


def tt_shard_size(begin, end):
    """doubles in the contiguous packed piece holding columns [begin, end)"""
    return packed_size(end) - packed_size(begin)


def tqu_shard_sizes(npix, begin, end):
    """doubles in the T, Q and U strips (columns s*npix + [begin, end)) of a [T;Q;U] matrix"""
    return [packed_size(s * npix + end) - packed_size(s * npix + begin) for s in range(3)]


def tqu_strip_offsets(npix, begin):
    """offsets of entry (0, s*npix + begin) inside one whole packed 3N matrix"""
    return [packed_size(s * npix + begin) for s in range(3)]


def tqu_rank_plan(npix, bounds, rank):
    """What rank `rank` of a sharded [T;Q;U] generation holds (no GPU needed to compute this).

    Returns a dict:
      columns   (begin, end) pixel columns whose pairs (i <= j) this rank computes
      strips    sizes in doubles of its packed T, Q, U strips (cmg_tqu_layout kind 0)
      outbox    list of (owner, n_owner_columns, ld, row0): for every owner left of this rank, three dense column-major
                blocks of n_owner_columns x ld doubles holding <Q_i T_j>, <U_i T_j>, <U_i Q_j> for the owner's columns i and
                this rank's rows j (cmg_tqu_layout kind 1); these are the entries whose packed home is another rank's strip
      pairs     number of pixel pairs computed
    """
    a0, a1 = bounds[rank], bounds[rank + 1]
    outbox = []
    for k in range(rank):
        if bounds[k + 1] > bounds[k] and a1 > a0:
            outbox.append((k, bounds[k + 1] - bounds[k], a1 - a0, a0))
    return {"columns": (a0, a1), "strips": tqu_shard_sizes(npix, a0, a1), "outbox": outbox, "pairs": pairs_in_block(a0, a1)}


def tqu_entries_held(npix, bounds, rank):
    """Number of matrix entries rank `rank` writes: 6 per pair on the diagonal (i == j), 9 per pair off it."""
    a0, a1 = bounds[rank], bounds[rank + 1]
    pairs = pairs_in_block(a0, a1)
    return 9 * pairs - 3 * (a1 - a0)


def batch_partition(n_batch, n_parts, slab=16):
    """Batched mode shards along the batch axis (each GPU regenerates its share of the C_l proposals; no exchange):
    boundaries b[0..n_parts] in batch elements, interior boundaries on slab multiples so that no slab of
    cmg_tqu_batched_slab straddles two ranks."""
    if n_parts < 1 or n_batch < 0:
        raise ValueError("n_parts must be >= 1 and n_batch >= 0")
    slabs = (n_batch + slab - 1) // slab
    b = [0]
    for k in range(1, n_parts):
        b.append(min(n_batch, slab * ((slabs * k) // n_parts)))
    b.append(n_batch)
    return b


# ---------------------------------------------------------------- symmetry-orbit path (cmg_tqu_orbit_sharded)

# (whole-face-pair classes, q_row <= q_col classes) of cmg_tqu_orbit's plan: mode 0 / 2 with transposed images, 1 without,
# 3 with the meridian mirror on top of mode 0 (five whole-face-pair classes stored as mirror images; single owner only)
ORBIT_UNITS = {0: (15, 6), 1: (21, 3), 2: (15, 6), 3: (10, 6)}


def orbit_column_cost(q, face_pix, mode=0):
    """source pixel pairs evaluated for the in-face column index q: whole-face classes contribute face_pix rows each,
    q_row <= q_col classes q + 1 (cosmopp_b200/csrc/orbit.cuh: 15 + 6 classes with transposed images, 21 + 3 without)"""
    full, tri = ORBIT_UNITS[mode]
    return full * face_pix + tri * (q + 1)


def orbit_partition(nside, n_parts, mode=0, align=32):
    """Boundaries b[0..n_parts] of the in-face column index (0 .. nside^2) giving every rank the same number of source pixel
    pairs; a rank owns the columns q in [b[r], b[r+1]) of ALL twelve base faces (an orbit-closed set)."""
    face_pix = nside * nside
    if n_parts < 1 or face_pix % align:
        raise ValueError("n_parts must be >= 1 and nside^2 a multiple of the column tile")
    full, tri = ORBIT_UNITS[mode]
    total = full * face_pix * face_pix + tri * face_pix * (face_pix + 1) // 2
    b = [0]
    for k in range(1, n_parts):
        # cumulative cost up to q: full F q + tri q (q + 1) / 2 = target  ->  solve the quadratic
        target = total * k / n_parts
        a2, a1 = tri / 2.0, full * face_pix + tri / 2.0
        q = (-a1 + math.sqrt(a1 * a1 + 4 * a2 * target)) / (2 * a2)
        v = int(round(q / align)) * align
        b.append(max(b[-1], min(face_pix, v)))
    b.append(face_pix)
    return b


def orbit_partition_blocks(nside, n_parts, block=128):
    """Boundaries of the in-face column index in whole blocks of `block` columns, as equal as they come (the lower ranks, whose
    columns are the cheapest to generate, take the remainder): what a sharded Cholesky factorisation of the strips wants
    (multigpu.ShardedCholesky: every run boundary on its block grid, the same number of blocks per rank).  Generation is then
    balanced to ~ +-12 % at Nside = 64 over 8 ranks instead of exactly -- milliseconds against the seconds of the factorisation."""
    face_pix = nside * nside
    if n_parts < 1 or face_pix % block:
        raise ValueError("n_parts must be >= 1 and nside^2 a multiple of the block")
    nb = face_pix // block
    b = [0]
    for k in range(n_parts):
        b.append(b[-1] + block * (nb // n_parts + (1 if k < nb % n_parts else 0)))
    return b


def orbit_pairs_in_range(q0, q1, face_pix, mode=0):
    """source pixel pairs a rank owning [q0, q1) evaluates"""
    full, tri = ORBIT_UNITS[mode]
    return full * face_pix * (q1 - q0) + tri * (q1 * (q1 + 1) - q0 * (q0 + 1)) // 2


def orbit_strip_sizes(nside, q0, q1):
    """doubles in the 3 x 12 packed column runs s N + f nside^2 + [q0, q1) a rank owns"""
    face_pix = nside * nside
    n = 12 * face_pix
    return [[packed_size(s * n + f * face_pix + q1) - packed_size(s * n + f * face_pix + q0) for f in range(12)] for s in range(3)]


def orbit_column_runs(nside, q0, q1):
    """the 36 runs [col_begin, col_end) of packed columns s N + f nside^2 + [q0, q1), ascending"""
    face_pix = nside * nside
    n = 12 * face_pix
    return [(s * n + f * face_pix + q0, s * n + f * face_pix + q1) for s in range(3) for f in range(12)]


ORB_SUB = 32        # rows and columns of an outbox sub-tile (cosmopp_b200/csrc/orbit.cuh)


def orbit_combo_counts(plan):
    """(nComboA, nComboB): (class, image, staged kind) combinations of the whole-face-pair classes and of the q_row <= q_col
    classes; a transposed image stages six kinds, a straight one three"""
    n_a = sum(6 if swap else 3 for c in plan if not c["tri"] for _, _, swap in c["images"])
    n_b = sum(6 if swap else 3 for c in plan if c["tri"] for _, _, swap in c["images"])
    return n_a, n_b


def orbit_outbox_offsets(plan, bounds, rank):
    """Python statement of cmg_orbit_outbox_layout (include/cmg.h, cmg_orbit_shard): the outbox of `rank` is compact and ordered
    by destination; the block for rank d holds, for every combination that can address d (all for d < rank, only those of the
    whole-face-pair classes for d > rank), the 32 x 32 sub-tiles (row half-tile of d's range) x (column tile of rank's range).
    Returns offsets[d] in doubles, offsets[-1] = size of the outbox."""
    n_a, n_b = orbit_combo_counts(plan)
    nct = (bounds[rank + 1] - bounds[rank]) // ORB_SUB
    off, at = [], 0
    for d in range(len(bounds) - 1):
        off.append(at)
        if d != rank:
            nh = (bounds[d + 1] - bounds[d]) // ORB_SUB
            at += (n_a + (n_b if d < rank else 0)) * nct * nh * ORB_SUB * ORB_SUB
    off.append(at)
    return off


def orbit_outbox_index(bounds, rank, offsets, combo, q_row, q_col):
    """element of rank's outbox that takes (combination, row pixel index q_row outside the rank's range, column pixel index
    q_col inside it); numpy arrays broadcast"""
    import numpy as np
    b = np.asarray(bounds)
    d = np.searchsorted(b, q_row, side="right") - 1
    nh = (b[d + 1] - b[d]) // ORB_SUB
    nct = (bounds[rank + 1] - bounds[rank]) // ORB_SUB
    ct = (q_col - bounds[rank]) // ORB_SUB
    h = q_row // ORB_SUB - b[d] // ORB_SUB
    return (np.asarray(offsets)[d] + ((combo * nct + ct) * nh + h) * (ORB_SUB * ORB_SUB) + (q_row % ORB_SUB) * ORB_SUB + (q_col % ORB_SUB))
