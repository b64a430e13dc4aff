"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Tolerance (BASELINE.json north_star): every entry within 1e-11 relative to the diagonal of its block;
index layout bit-exact (checked by comparing whole packed arrays position by position).
"""
import numpy as np
import pytest

from conftest import synthetic_cl

pytestmark = pytest.mark.gpu

REL_TOL = 1e-11


def _torch():
    import torch
    return torch


def _tt_gpu(ctx, cl, nside, fwhm, good=None):
    torch = _torch()
    from cosmopp_b200 import capi
    ctx.set_pixels(nside, good)
    n = ctx.npix
    f = capi.window_beam(len(cl) - 1, fwhm)
    a = capi.tt_weights(cl, f)
    out = torch.full((capi.packed_size(n),), float("nan"), dtype=torch.float64, device="cuda")
    ctx.legendre_series(a, out)
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("nside,lmax,masked", [(1, 5, False), (2, 8, False), (4, 12, False), (8, 24, True),
                                                (16, 47, False), (16, 47, True), (16, 30, True)])
def test_tt_matches_oracle(gpu_ctx, oracle_api, nside, lmax, masked):
    cl = synthetic_cl(lmax)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside)) if masked else None
    got = _tt_gpu(gpu_ctx, cl, nside, 10.0, good)
    want = oracle_api.cl_to_cmatrix(cl, nside, 10.0, good=good)
    assert got.shape == want.shape
    assert not np.isnan(got).any()
    diag = want[0]
    assert np.abs(got - want).max() <= REL_TOL * diag


@pytest.mark.skipif("not __import__('oracle.api').api.have_ref()")
def test_tt_matches_reference_object_code(gpu_ctx, oracle_api):
    nside, lmax = 8, 20
    cl = synthetic_cl(lmax)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    got = _tt_gpu(gpu_ctx, cl, nside, 10.0, good)
    want = oracle_api.ref_cl_to_cmatrix(cl, nside, 10.0, good=good)
    assert np.abs(got - want).max() <= REL_TOL * want[0]


def test_tt_host_api_and_fiducial(gpu_ctx, oracle_api):
    torch = _torch()
    from cosmopp_b200 import capi
    nside, lmax = 8, 20
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    cl = synthetic_cl(4 * nside)
    gpu_ctx.set_pixels(nside, good)
    n = gpu_ctx.npix
    out = torch.empty(capi.packed_size(n), dtype=torch.float64, pin_memory=True)
    gpu_ctx.cl_to_cmatrix(cl[:lmax + 1], 10.0, out)
    want = oracle_api.cl_to_cmatrix(cl[:lmax + 1], nside, 10.0, good=good)
    assert np.abs(out.numpy() - want).max() <= REL_TOL * want[0]
    gpu_ctx.fiducial_matrix(cl, lmax, 10.0, out)
    want = oracle_api.fiducial_matrix(cl, nside, lmax, 10.0, good=good)
    assert np.abs(out.numpy() - want).max() <= REL_TOL * want[0]


def test_tt_column_shards_tile_the_triangle(gpu_ctx, oracle_api):
    torch = _torch()
    from cosmopp_b200 import capi, partition
    nside, lmax = 8, 16
    cl = synthetic_cl(lmax)
    gpu_ctx.set_pixels(nside)
    n = gpu_ctx.npix
    a = capi.tt_weights(cl, capi.window_beam(lmax, 10.0))
    whole = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    gpu_ctx.legendre_series(a, whole)
    for parts in (2, 3, 8):
        b = partition.column_partition(n, parts, align=16)
        pieces = []
        for k in range(parts):
            piece = torch.full((partition.tt_shard_size(b[k], b[k + 1]),), float("nan"), dtype=torch.float64, device="cuda")
            gpu_ctx.legendre_series(a, piece, b[k], b[k + 1])
            pieces.append(piece)
        assert torch.equal(torch.cat(pieces), whole)


def _tqu_gpu(ctx, spectra, nside, fwhm, good=None):
    torch = _torch()
    from cosmopp_b200 import capi
    ctx.set_pixels(nside, good)
    n = ctx.npix
    lmax = len(spectra[0]) - 1
    f = capi.window_beam(lmax, fwhm)
    a = capi.tqu_weights(*spectra, f, f)
    out = torch.full((capi.packed_size(3 * n),), float("nan"), dtype=torch.float64, device="cuda")
    ctx.tqu(*a, ctx.tqu_layout_single(out))
    torch.cuda.synchronize()
    return out.cpu().numpy(), n


def _block_diag_scales(packed, n):
    from oracle import api
    tt = packed[api.packed_size(1) - 1]
    qq = packed[(n) * (n + 1) // 2 + n]
    return tt, qq


def _assert_tqu_close(got, want, n):
    assert not np.isnan(got).any(), "entries never written: %d" % np.isnan(got).sum()
    from oracle import api
    M = api.unpack_symmetric(want, 3 * n)
    G = api.unpack_symmetric(got, 3 * n)
    dT = M[0, 0]
    dP = M[n, n]
    scale = np.empty((3 * n, 3 * n))
    s = np.array([dT] * n + [dP] * 2 * n)
    scale = np.sqrt(np.outer(s, s))       # TT block: dT, pol blocks: dP, cross: geometric mean
    assert (np.abs(G - M) / scale).max() <= REL_TOL


@pytest.mark.parametrize("nside,lmax,masked", [(1, 6, False), (2, 8, False), (4, 12, False), (4, 12, True), (8, 24, True)])
def test_tqu_matches_oracle(gpu_ctx, oracle_api, nside, lmax, masked):
    spectra = synthetic_cl(lmax, pol=True)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside)) if masked else None
    got, n = _tqu_gpu(gpu_ctx, spectra, nside, 10.0, good)
    want = oracle_api.tqu_matrix(*spectra, nside, 10.0, good=good)
    _assert_tqu_close(got, want, n)


def test_tqu_matches_bruteforce_definition(gpu_ctx, oracle_api):
    from oracle import pol_bruteforce as bf
    nside, lmax = 1, 6
    spectra = synthetic_cl(lmax, pol=True)
    got, n = _tqu_gpu(gpu_ctx, spectra, nside, 0.0)
    ang = [oracle_api.pix2ang_nest(nside, i) for i in range(n)]
    C = bf.covariance(ang, *spectra, lmax)
    G = oracle_api.unpack_symmetric(got, 3 * n)
    assert np.abs(G - C).max() <= 1e-12 * C[0, 0]


def test_tqu_nside16_masked_sampled(gpu_ctx, oracle_api):
    """BASELINE config 2 (polarized Nside=16 lmax=47 with the test_like_low mask): all diagonal blocks and a
    random sample of pairs against the oracle."""
    nside, lmax = 16, 47
    spectra = synthetic_cl(lmax, pol=True)
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    got, n = _tqu_gpu(gpu_ctx, spectra, nside, 10.0, good)
    assert not np.isnan(got).any()
    rs = np.random.RandomState(7)
    pj = rs.randint(0, n, 20000)
    pi = (rs.uniform(size=20000) * (pj + 1)).astype(np.int64)
    pi = np.concatenate([pi, np.arange(n)])
    pj = np.concatenate([pj, np.arange(n)])
    blocks = oracle_api.tqu_pairs(*spectra, nside, 10.0, pi, pj, good=good)
    dT = blocks[-1, 0, 0]
    dP = blocks[-1, 1, 1]
    s = np.array([dT, dP, dP])
    scale = np.sqrt(np.outer(s, s))
    idx = lambda r, c: np.where(r <= c, c * (c + 1) // 2 + r, r * (r + 1) // 2 + c)
    worst = 0.0
    for a in range(3):
        for b in range(3):
            r = a * n + pi
            c = b * n + pj
            worst = max(worst, (np.abs(got[idx(r, c)] - blocks[:, a, b]) / scale[a, b]).max())
    assert worst <= REL_TOL


def _whole_tqu(ctx, a, n):
    torch = _torch()
    from cosmopp_b200 import capi
    whole = torch.full((capi.packed_size(3 * n),), float("nan"), dtype=torch.float64, device="cuda")
    ctx.tqu(*a, ctx.tqu_layout_single(whole))
    torch.cuda.synchronize()
    return whole


@pytest.mark.parametrize("parts", [2, 3, 5])
def test_tqu_sharded_in_place_layout_is_bit_identical(gpu_ctx, oracle_api, parts):
    """All parts packed (kind 0), as with peer-mapped strips of other GPUs: every rank's launch writes the transposed
    partners straight into the owner's strip.  Run rank after rank on one GPU; the strips then tile the whole matrix."""
    torch = _torch()
    from cosmopp_b200 import capi, partition
    nside, lmax = 8, 18
    gpu_ctx.set_pixels(nside)
    n = gpu_ctx.npix
    f = capi.window_beam(lmax, 10.0)
    a = capi.tqu_weights(*synthetic_cl(lmax, pol=True), f, f)
    whole = _whole_tqu(gpu_ctx, a, n)
    b = partition.column_partition(n, parts, align=32)
    strips = [[torch.full((s,), float("nan"), dtype=torch.float64, device="cuda") for s in partition.tqu_shard_sizes(n, b[k], b[k + 1])]
              for k in range(parts)]
    for r in range(parts):
        lay = capi.TquLayout()
        lay.n_parts, lay.own = parts, r
        for k in range(parts + 1):
            lay.begin[k] = b[k]
        for k in range(parts):
            for s in range(3):
                lay.ptr[k][s] = strips[k][s].data_ptr()
        gpu_ctx.tqu(*a, lay)
    torch.cuda.synchronize()
    off = 0
    for s in range(3):
        for k in range(parts):
            piece = strips[k][s]
            assert torch.equal(piece, whole[partition.tqu_strip_offsets(n, b[k])[s]:][:piece.numel()])
            off += piece.numel()
    assert off == whole.numel()


def test_tqu_sharded_outbox_layout_assembles_to_whole(gpu_ctx, oracle_api):
    """The no-communication layout of the multi-GPU bench: transposed partners of cross pairs land in local dense
    blocks (kind 1).  Scatter them into place on the host and compare with the one-piece matrix, bit for bit."""
    torch = _torch()
    from cosmopp_b200 import capi, partition
    nside, lmax, parts = 8, 18, 3
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    gpu_ctx.set_pixels(nside, good)
    n = gpu_ctx.npix
    f = capi.window_beam(lmax, 10.0)
    a = capi.tqu_weights(*synthetic_cl(lmax, pol=True), f, f)
    whole = _whole_tqu(gpu_ctx, a, n).cpu().numpy()
    b = partition.column_partition(n, parts, align=32)
    assembled = np.full(capi.packed_size(3 * n), np.nan)
    pk = lambda r, c: c * (c + 1) // 2 + r
    for r in range(parts):
        plan = partition.tqu_rank_plan(n, b, r)
        strips = [torch.full((s,), float("nan"), dtype=torch.float64, device="cuda") for s in plan["strips"]]
        outbox = {o: [torch.full((nc * ld,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(3)] for o, nc, ld, _ in plan["outbox"]}
        lay = capi.make_tqu_layout(b, r, [t.data_ptr() for t in strips], {k: [t.data_ptr() for t in v] for k, v in outbox.items()})
        gpu_ctx.tqu(*a, lay)
        torch.cuda.synchronize()
        a0, a1 = plan["columns"]
        for s in range(3):
            got = strips[s].cpu().numpy()
            lo = partition.tqu_strip_offsets(n, a0)[s]
            seg = assembled[lo:lo + len(got)]
            seg[~np.isnan(got)] = got[~np.isnan(got)]
        for o, nc, ld, row0 in plan["outbox"]:
            for t in range(3):
                blk = outbox[o][t].cpu().numpy().reshape(nc, ld)          # [owner column i][row j]
                i = b[o] + np.arange(nc)[:, None]
                j = row0 + np.arange(ld)[None, :]
                if t == 0: dst = pk(j, n + i)                              # <Q_i T_j> -> column N+i, row j
                elif t == 1: dst = pk(j, 2 * n + i)                        # <U_i T_j> -> column 2N+i, row j
                else: dst = pk(n + j, 2 * n + i)                           # <U_i Q_j> -> column 2N+i, row N+j
                assert not np.isnan(blk).any()
                assembled[dst] = blk
    assert not np.isnan(assembled).any()
    assert np.array_equal(assembled, whole)


def test_mask_matrix_gather_kernel(gpu_ctx, oracle_api):
    torch = _torch()
    from cosmopp_b200 import capi
    n = 192
    rs = np.random.RandomState(5)
    packed = rs.standard_normal(capi.packed_size(n))
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(4))
    d_in = torch.from_numpy(packed).cuda()
    d_out = torch.empty(capi.packed_size(len(good)), dtype=torch.float64, device="cuda")
    gpu_ctx.mask_matrix(d_in, n, good, d_out)
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), oracle_api.mask_matrix(packed, good))


def test_batched_generation_matches_single_calls(gpu_ctx, oracle_api):
    torch = _torch()
    from cosmopp_b200 import capi
    nside, lmax, nb = 4, 12, 5
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    gpu_ctx.set_pixels(nside, good)
    n = gpu_ctx.npix
    f = capi.window_beam(lmax, 10.0)
    # TT
    a = np.stack([capi.tt_weights(synthetic_cl(lmax, seed=100 + b), f) for b in range(nb)])
    stride = capi.packed_size(n) + 7
    out = torch.full((nb * stride,), float("nan"), dtype=torch.float64, device="cuda")
    gpu_ctx.legendre_series_batched(a, out, stride)
    one = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    for b in range(nb):
        gpu_ctx.legendre_series(a[b], one)
        torch.cuda.synchronize()
        assert torch.equal(out[b * stride:b * stride + one.numel()], one)
        want = oracle_api.cl_to_cmatrix(synthetic_cl(lmax, seed=100 + b), nside, 10.0, good=good)
        assert np.abs(one.cpu().numpy() - want).max() <= REL_TOL * want[0]
    # TQU
    ab = np.stack([np.stack(capi.tqu_weights(*synthetic_cl(lmax, seed=200 + b, pol=True), f, f)) for b in range(nb)])
    stride = capi.packed_size(3 * n)
    wants = [oracle_api.tqu_matrix(*synthetic_cl(lmax, seed=200 + b, pol=True), nside, 10.0, good=good) for b in range(nb)]
    try:
        for variant in (0, 900, 901):     # per-element Clenshaw (default), shared-basis kernel, DMMA kernel
            gpu_ctx.set_kernel_variant(variant)
            outp = torch.full((nb * stride,), float("nan"), dtype=torch.float64, device="cuda")
            gpu_ctx.tqu_batched(ab, outp, stride)
            torch.cuda.synchronize()
            for b in range(nb):
                _assert_tqu_close(outp[b * stride:(b + 1) * stride].cpu().numpy(), wants[b], n)
    finally:
        gpu_ctx.set_kernel_variant(0)


def test_batched_slab_dmma_path(gpu_ctx, oracle_api):
    """DMMA batched generator with slab output (cmg_tqu_batched_slab): several 16-element slabs (3-stage weight ring), a
    batch size that is not a multiple of 16, a pixel count that is not a multiple of the 8 x 8 tile, the three
    fragment shapes (lmax <= 31, 47, 63); every element extracted with cmg_slab_unpack and compared with the oracle,
    the index layout position by position."""
    torch = _torch()
    from cosmopp_b200 import capi
    nside, nb = 4, 53
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    gpu_ctx.set_pixels(nside, good)
    n = gpu_ctx.npix
    assert n % 8 != 0
    packed = capi.packed_size(3 * n)
    sd = capi.slab_doubles(3 * n)
    assert sd == 16 * packed
    n_slabs = (nb + 15) // 16
    for lmax in (9, 40, 60):
        f = capi.window_beam(lmax, 10.0)
        ab = np.stack([np.stack(capi.tqu_weights(*synthetic_cl(lmax, seed=300 + b, pol=True), f, f)) for b in range(nb)])
        slabs = torch.full((n_slabs * sd,), float("nan"), dtype=torch.float64, device="cuda")
        gpu_ctx.tqu_batched_slab(ab, slabs)
        torch.cuda.synchronize()
        host = slabs.cpu().numpy().reshape(n_slabs, packed, 16)
        assert not np.isnan(host).any()
        # device-resident weights: the same slabs, bit for bit
        slabs_dev = torch.full_like(slabs, float("nan"))
        gpu_ctx.tqu_batched_slab_dev(torch.from_numpy(np.ascontiguousarray(ab)).cuda(), lmax, nb, slabs_dev)
        torch.cuda.synchronize()
        assert torch.equal(slabs_dev, slabs)
        assert (host[-1][:, nb % 16:] == 0).all()           # padding elements of the last slab are zeros
        outs = torch.full((16, packed + 5), float("nan"), dtype=torch.float64, device="cuda")
        one = torch.empty(packed, dtype=torch.float64, device="cuda")
        for k in range(n_slabs):
            live = min(16, nb - 16 * k)
            outs.fill_(float("nan"))
            gpu_ctx.slab_unpack(slabs[k * sd:(k + 1) * sd], 3 * n, outs, packed + 5, n_live=live)
            torch.cuda.synchronize()
            o = outs.cpu().numpy()
            assert np.isnan(o[:, packed:]).all() and np.isnan(o[live:]).all()
            assert np.array_equal(o[:live, :packed], host[k].T[:live])
            gpu_ctx.slab_unpack(slabs[k * sd:(k + 1) * sd], 3 * n, one, only_b=live - 1)
            torch.cuda.synchronize()
            assert np.array_equal(one.cpu().numpy(), host[k][:, live - 1])
        for b in (0, 15, 16, 31, 47, 52):
            want = oracle_api.tqu_matrix(*synthetic_cl(lmax, seed=300 + b, pol=True), nside, 10.0, good=good)
            _assert_tqu_close(np.ascontiguousarray(host[b // 16][:, b % 16]), want, n)
    # full sky, tile-aligned pixel count, against the per-element default path
    gpu_ctx.set_pixels(4)
    n = gpu_ctx.npix
    packed = capi.packed_size(3 * n)
    lmax, nb = 12, 16
    f = capi.window_beam(lmax, 10.0)
    ab = np.stack([np.stack(capi.tqu_weights(*synthetic_cl(lmax, seed=400 + b, pol=True), f, f)) for b in range(nb)])
    slabs = torch.full((capi.slab_doubles(3 * n),), float("nan"), dtype=torch.float64, device="cuda")
    gpu_ctx.tqu_batched_slab(ab, slabs)
    ref = torch.empty(nb * packed, dtype=torch.float64, device="cuda")
    gpu_ctx.tqu_batched(ab, ref, packed)
    torch.cuda.synchronize()
    got = slabs.cpu().numpy().reshape(packed, 16).T
    want = ref.cpu().numpy().reshape(nb, packed)
    dP = want[0][capi.packed_index(n, n)]
    assert np.abs(got - want).max() <= REL_TOL * dP
    with pytest.raises(Exception):
        gpu_ctx.tqu_batched_slab(np.zeros((2, 4, 65)), slabs)      # lmax = 64: refused, not silently rerouted


def test_batched_slab_at_config4_shape(gpu_ctx, oracle_api):
    """BASELINE config 4's shape (full-sky Nside=16, lmax=47), two slabs with a ragged second one: every element of the
    slab output against the single-matrix kernel, all 42.5 M entries, to 1e-11 of the block diagonal; one element (of the
    second, ragged slab) DIRECTLY against the CPU oracle; plus linearity in the weights."""
    torch = _torch()
    from cosmopp_b200 import capi
    nside, lmax, nb = 16, 47, 19
    gpu_ctx.set_pixels(nside)
    n = gpu_ctx.npix
    packed = capi.packed_size(3 * n)
    f = capi.window_beam(lmax, 10.0)
    ws = [capi.tqu_weights(*synthetic_cl(lmax, seed=12345 + b, pol=True), f, f) for b in range(nb)]
    ab = np.stack([np.stack(w) for w in ws])
    ab[nb - 1] = 0.5 * ab[0] - 2.0 * ab[1]                       # linear combination: its matrix must be the same combination
    slabs = torch.empty(2 * capi.slab_doubles(3 * n), dtype=torch.float64, device="cuda")
    gpu_ctx.tqu_batched_slab(ab, slabs)
    one = torch.empty(packed, dtype=torch.float64, device="cuda")
    el = torch.empty(packed, dtype=torch.float64, device="cuda")
    lay = gpu_ctx.tqu_layout_single(one)
    i_tt, i_qq = 0, capi.packed_index(n, n)
    keep = {}
    for b in (0, 1, 7, 16, 17):
        gpu_ctx.tqu(*[np.ascontiguousarray(x) for x in ab[b]], lay)
        gpu_ctx.slab_unpack(slabs[(b // 16) * capi.slab_doubles(3 * n):], 3 * n, el, only_b=b % 16)
        torch.cuda.synchronize()
        dT, dP = float(one[i_tt]), float(one[i_qq])
        assert dT > 0 and dP > 0
        # conservative scale: the smaller (polarization) diagonal for every entry
        assert float((el - one).abs().max()) <= REL_TOL * min(dT, dP)
        if b in (0, 1):
            keep[b] = el.clone()
    # element 17 straight against the oracle (not through this library's other kernels)
    gpu_ctx.slab_unpack(slabs[capi.slab_doubles(3 * n):], 3 * n, el, only_b=17 % 16)
    torch.cuda.synchronize()
    want = oracle_api.tqu_matrix(*synthetic_cl(lmax, seed=12345 + 17, pol=True), nside, 10.0)
    scale = np.full(want.shape, want[i_qq])
    scale[:capi.packed_size(n)] = want[i_tt]
    assert (np.abs(el.cpu().numpy() - want) / scale).max() <= REL_TOL
    gpu_ctx.slab_unpack(slabs[capi.slab_doubles(3 * n):], 3 * n, el, only_b=(nb - 1) % 16)
    torch.cuda.synchronize()
    combo = 0.5 * keep[0] - 2.0 * keep[1]
    assert float((el - combo).abs().max()) <= 1e-11 * float(keep[0][i_tt])


def test_cmatrix_file_streamed_from_device_shards(gpu_ctx, oracle_api, tmp_path):
    """The reference's binary CMatrix file written piece by piece from device-resident shards (three ranks' TT column
    blocks, out of order) and read back into device memory: byte-identical to the reference's own writer."""
    torch = _torch()
    import struct
    from cosmopp_b200 import capi, partition
    nside, lmax = 8, 20
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    gpu_ctx.set_pixels(nside, good)
    n = gpu_ctx.npix
    a = capi.tt_weights(synthetic_cl(lmax), capi.window_beam(lmax, 10.0))
    whole = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    gpu_ctx.legendre_series(a, whole)
    bounds = partition.column_partition(n, 3, align=16)
    pieces = []
    for r in (2, 0, 1):
        a0, a1 = bounds[r], bounds[r + 1]
        shard = torch.full((partition.tt_shard_size(a0, a1),), float("nan"), dtype=torch.float64, device="cuda")
        gpu_ctx.legendre_series(a, shard, a0, a1)
        pieces.append((partition.packed_size(a0), shard, shard.numel()))
    path = str(tmp_path / "c.dat")
    gpu_ctx.write_cmatrix_file(path, n, pieces, comment="streamed from shards")
    torch.cuda.synchronize()
    raw = open(path, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == n
    body = np.frombuffer(raw[4:4 + 8 * whole.numel()], dtype="<f8")
    assert np.array_equal(body, whole.cpu().numpy())
    assert raw[4 + 8 * whole.numel():] == struct.pack("<i", 20) + b"streamed from shards"
    if oracle_api.have_ref():
        oracle_api.ref_write_cmatrix(whole.cpu().numpy(), n, "streamed from shards", str(tmp_path / "ref.dat"), str(tmp_path / "ref.txt"))
        assert open(str(tmp_path / "ref.dat"), "rb").read() == raw
    back = torch.full((whole.numel(),), float("nan"), dtype=torch.float64, device="cuda")
    n_read, comment = gpu_ctx.read_cmatrix_file(path, back)
    assert n_read == n and comment == "streamed from shards" and torch.equal(back, whole)
    part = torch.empty(100, dtype=torch.float64, device="cuda")
    gpu_ctx.read_cmatrix_file(path, part, first=1234, count=100)
    assert torch.equal(part, whole[1234:1334])
    with pytest.raises(Exception):
        gpu_ctx.read_cmatrix_file(str(tmp_path / "missing.dat"), back)
    with pytest.raises(Exception):
        gpu_ctx.write_cmatrix_file(path, n, [(whole.numel() - 5, whole, 10)])      # piece beyond the triangle


def test_all_kernel_variants_agree(gpu_ctx, oracle_api):
    """Static-table and shared-memory-table kernels, all column counts: same matrix to rounding, on a full sky (whole
    tiles) and on a masked sky (ragged last tiles)."""
    torch = _torch()
    from cosmopp_b200 import capi
    nside, lmax = 8, 33
    f = capi.window_beam(lmax, 10.0)
    spectra = synthetic_cl(lmax, pol=True)
    a = capi.tqu_weights(*spectra, f, f)
    try:
        for good in (None, oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))):
            gpu_ctx.set_pixels(nside, good)
            n = gpu_ctx.npix
            want = oracle_api.tqu_matrix(*spectra, nside, 10.0, good=good)
            for v in (22, 42, 81, 114, 122, 123, 124, 142):
                gpu_ctx.set_kernel_variant(v)
                got = _whole_tqu(gpu_ctx, a, n).cpu().numpy()
                _assert_tqu_close(got, want, n)
    finally:
        gpu_ctx.set_kernel_variant(0)


def test_tt_kernel_variants_agree(gpu_ctx, oracle_api):
    """TT kernel: columns-per-thread / occupancy variants and the shared-memory-table kernel, against the oracle."""
    torch = _torch()
    from cosmopp_b200 import capi
    nside = 8
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    gpu_ctx.set_pixels(nside, good)
    n = gpu_ctx.npix
    out = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    try:
        for lmax in (20, 127, 128, 150):
            cl = synthetic_cl(lmax)
            a = capi.tt_weights(cl, capi.window_beam(lmax, 10.0))
            want = oracle_api.cl_to_cmatrix(cl, nside, 10.0, good=good)
            for v in (0, 1, 248, 2216):
                gpu_ctx.set_kernel_variant(v)
                out.fill_(float("nan"))
                gpu_ctx.legendre_series(a, out)
                torch.cuda.synchronize()
                assert np.abs(out.cpu().numpy() - want).max() <= REL_TOL * want[0], (lmax, v)
    finally:
        gpu_ctx.set_kernel_variant(0)


def test_argument_errors_are_reported_not_crashed(gpu_ctx):
    torch = _torch()
    from cosmopp_b200 import capi
    import cosmopp_b200 as cb
    gpu_ctx.set_pixels(2)
    out = torch.empty(capi.packed_size(48), dtype=torch.float64, device="cuda")
    with pytest.raises(cb.CmgError):
        gpu_ctx.legendre_series(np.ones(capi.LMAX_LIMIT + 2), out)            # lmax beyond the limit
    with pytest.raises(cb.CmgError):
        gpu_ctx.legendre_series(np.ones(5), out, 0, 49)                        # column range outside the matrix
    with pytest.raises(cb.CmgError):
        gpu_ctx.set_pixels(12)                                                 # not a power of two
    with pytest.raises(cb.CmgError):
        gpu_ctx.set_pixels(2, [0, 48])                                         # pixel index out of range
    fresh = cb.Context(0)
    with pytest.raises(cb.CmgError):
        fresh.legendre_series(np.ones(5), out)                                 # geometry not set
    fresh.close()


def test_tt_large_sample_nside32(gpu_ctx, oracle_api):
    """BASELINE config 3 size (TT Nside=32, lmax=96): every diagonal entry, every antipodal pair and a random sample
    of 10^6 entries against the oracle."""
    torch = _torch()
    from cosmopp_b200 import capi
    nside, lmax = 32, 96
    cl = synthetic_cl(lmax)
    got = _tt_gpu(gpu_ctx, cl, nside, 10.0)
    n = 12 * nside * nside
    assert not np.isnan(got).any()
    rs = np.random.RandomState(11)
    pj = rs.randint(0, n, 1000000)
    pi = (rs.uniform(size=len(pj)) * (pj + 1)).astype(np.int64)
    v = oracle_api.unit_vectors(nside)
    anti = np.argmin(v @ v[:2048].T, axis=0)                                   # antipodes of the first 2048 pixels
    ai, aj = np.minimum(anti, np.arange(2048)), np.maximum(anti, np.arange(2048))
    pi = np.concatenate([pi, np.arange(n), ai])
    pj = np.concatenate([pj, np.arange(n), aj])
    want = oracle_api.cl_to_cmatrix_pairs(cl, nside, 10.0, pi, pj, good=np.arange(n, dtype=np.int32))
    assert np.abs(got[pj * (pj + 1) // 2 + pi] - want).max() <= REL_TOL * want[len(want) - 2048 - n]


def test_scatter_block_places_outbox_entries(gpu_ctx, oracle_api):
    """cmg_tqu_scatter_block: outbox blocks of a 3-way sharded run land where the one-piece matrix has them."""
    torch = _torch()
    from cosmopp_b200 import capi, partition
    nside, lmax, parts = 8, 14, 3
    gpu_ctx.set_pixels(nside)
    n = gpu_ctx.npix
    f = capi.window_beam(lmax, 10.0)
    a = capi.tqu_weights(*synthetic_cl(lmax, pol=True), f, f)
    whole = _whole_tqu(gpu_ctx, a, n)
    b = partition.column_partition(n, parts, align=32)
    full = torch.full_like(whole, float("nan"))
    for r in range(parts):
        plan = partition.tqu_rank_plan(n, b, r)
        strips = [torch.full((s,), float("nan"), dtype=torch.float64, device="cuda") for s in plan["strips"]]
        outbox = {o: [torch.full((nc * ld,), float("nan"), dtype=torch.float64, device="cuda") for _ in range(3)] for o, nc, ld, _ in plan["outbox"]}
        gpu_ctx.tqu(*a, capi.make_tqu_layout(b, r, [t.data_ptr() for t in strips], {k: [t.data_ptr() for t in v] for k, v in outbox.items()}))
        torch.cuda.synchronize()
        for s in range(3):
            lo = partition.tqu_strip_offsets(n, b[r])[s]
            seg = full[lo:lo + strips[s].numel()]
            m = ~torch.isnan(strips[s])
            seg[m] = strips[s][m]
        for o, nc, ld, row0 in plan["outbox"]:
            for t in range(3):
                gpu_ctx.tqu_scatter_block(outbox[o][t], b[o], nc, ld, row0, t, full)
    torch.cuda.synchronize()
    assert torch.equal(full, whole)


@pytest.mark.parametrize("lmax", [0, 1, 2, 3, 441, 442, 700])
def test_tt_lmax_edges(gpu_ctx, oracle_api, lmax):
    """Series of length 1 and 2 (monopole/dipole only, as the fiducial term uses), the last lmax of the static-table
    kernel, the first of the shared-memory fallback, and a long series."""
    torch = _torch()
    from cosmopp_b200 import capi
    nside = 2
    gpu_ctx.set_pixels(nside)
    n = gpu_ctx.npix
    rs = np.random.RandomState(lmax)
    a = rs.uniform(0.5, 1.5, lmax + 1) / (1.0 + np.arange(lmax + 1)) ** 2
    out = torch.full((capi.packed_size(n),), float("nan"), dtype=torch.float64, device="cuda")
    gpu_ctx.legendre_series(a, out)
    torch.cuda.synchronize()
    v = oracle_api.unit_vectors(nside)
    z = np.clip(v @ v.T, -1, 1)
    from numpy.polynomial import legendre as L
    want = L.legval(z, a)
    got = oracle_api.unpack_symmetric(out.cpu().numpy(), n)
    assert np.abs(got - want).max() <= 1e-12 * np.abs(a).sum()


@pytest.mark.parametrize("lmax", [2, 3, 441, 442, 600])
def test_tqu_lmax_edges(gpu_ctx, oracle_api, lmax):
    nside = 2
    spectra = synthetic_cl(lmax, pol=True)
    got, n = _tqu_gpu(gpu_ctx, spectra, nside, 1.0)
    want = oracle_api.tqu_matrix(*spectra, nside, 1.0)
    _assert_tqu_close(got, want, n)


@pytest.mark.parametrize("pixels", [[7], [3, 100], list(range(33)), list(range(0, 192, 3)), [50, 2, 2, 191, 0, 77]])
def test_ragged_pixel_lists(gpu_ctx, oracle_api, pixels):
    """One pixel, sizes straddling the tile edges, and an unsorted list with a repeated pixel (the reference takes
    goodPixels in the caller's order and does not require uniqueness)."""
    nside, lmax = 4, 9
    good = np.array(pixels, dtype=np.int32)
    cl = synthetic_cl(lmax)
    got = _tt_gpu(gpu_ctx, cl, nside, 10.0, good)
    want = oracle_api.cl_to_cmatrix(cl, nside, 10.0, good=good)
    assert got.shape == want.shape and np.abs(got - want).max() <= REL_TOL * want[0]
    spectra = synthetic_cl(lmax, pol=True)
    gotp, n = _tqu_gpu(gpu_ctx, spectra, nside, 10.0, good)
    _assert_tqu_close(gotp, oracle_api.tqu_matrix(*spectra, nside, 10.0, good=good), n)


def test_linearity_in_the_spectra(gpu_ctx):
    """Size-independent property: the matrix is linear in C_l (checked at Nside=16 on whole matrices)."""
    torch = _torch()
    from cosmopp_b200 import capi
    nside, lmax = 16, 47
    gpu_ctx.set_pixels(nside)
    n = gpu_ctx.npix
    f = capi.window_beam(lmax, 10.0)
    s1, s2 = synthetic_cl(lmax, seed=1, pol=True), synthetic_cl(lmax, seed=2, pol=True)
    mats = []
    for sp in (s1, s2, tuple(2.0 * x + 0.5 * y for x, y in zip(s1, s2))):
        out = torch.empty(capi.packed_size(3 * n), dtype=torch.float64, device="cuda")
        gpu_ctx.tqu(*capi.tqu_weights(*sp, f, f), gpu_ctx.tqu_layout_single(out))
        mats.append(out)
    torch.cuda.synchronize()
    resid = (mats[2] - (2.0 * mats[0] + 0.5 * mats[1])).abs().max().item()
    assert resid <= 1e-12 * mats[2][0].item()


@pytest.mark.parametrize("orbit", [False, True])
def test_full_size_nside64_polarized_sampled(gpu_ctx, oracle_api, orbit):
    """BASELINE config 5 at full size (147456 x 147456, 87 GB on the device): all nine entries of 150k random pixel
    pairs, of every diagonal pair and of 2000 antipodal pairs against the oracle, plus the block structure on the diagonal.
    Once with every pair evaluated (cmg_tqu) and once over symmetry orbits (cmg_tqu_orbit, what bench.py's default runs)."""
    torch = _torch()
    from cosmopp_b200 import capi
    free, _total = torch.cuda.mem_get_info()
    nside, lmax = 64, 192
    n = 12 * nside * nside
    need = 8 * capi.packed_size(3 * n)
    if free < need + (4 << 30):
        pytest.skip("needs %.0f GB of free device memory" % (need / 1e9))
    spectra = synthetic_cl(lmax, pol=True)
    gpu_ctx.set_pixels(nside)
    f = capi.window_beam(lmax, 10.0)
    a = capi.tqu_weights(*spectra, f, f)
    out = torch.empty(capi.packed_size(3 * n), dtype=torch.float64, device="cuda")
    out.fill_(float("nan"))
    if orbit:
        gpu_ctx.tqu_orbit(*a, out, 0)
    else:
        gpu_ctx.tqu(*a, gpu_ctx.tqu_layout_single(out))
    torch.cuda.synchronize()
    assert not torch.isnan(out[::4099]).any()
    rs = np.random.RandomState(64)
    pj = rs.randint(0, n, 150000)
    pi = (rs.uniform(size=len(pj)) * (pj + 1)).astype(np.int64)
    v = oracle_api.unit_vectors(nside)
    anti = np.array([np.argmin(v @ v[k]) for k in range(2000)])
    ai, aj = np.minimum(anti, np.arange(2000)), np.maximum(anti, np.arange(2000))
    diag = np.arange(0, n, 7)
    pi = np.concatenate([pi, diag, ai])
    pj = np.concatenate([pj, diag, aj])
    blocks = oracle_api.tqu_pairs(*spectra, nside, 10.0, pi, pj, good=np.arange(n, dtype=np.int32))
    dT, dP = blocks[len(pj) - 2000 - 1, 0, 0], blocks[len(pj) - 2000 - 1, 1, 1]
    scale = np.sqrt(np.outer([dT, dP, dP], [dT, dP, dP]))
    worst = 0.0
    for ia in range(3):
        for ib in range(3):
            r, c = ia * n + pi, ib * n + pj
            lo, hi = np.minimum(r, c), np.maximum(r, c)
            idx = torch.from_numpy(hi * (hi + 1) // 2 + lo).cuda()
            got = out[idx].cpu().numpy()
            worst = max(worst, (np.abs(got - blocks[:, ia, ib]) / scale[ia, ib]).max())
    assert worst <= REL_TOL, worst
    del out
    torch.cuda.empty_cache()


def test_device_resident_likelihood_consumer(gpu_ctx, oracle_api):
    """C + F + N -> Cholesky -> chi2, logDet as reference source/likelihood.cpp:100-180, everything staying on the GPU,
    against a numpy restatement on the oracle's matrices."""
    torch = _torch()
    from cosmopp_b200 import capi
    from cosmopp_b200.likelihood import Likelihood, DET_OFFSET
    nside, lmax = 8, 20
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    cl = synthetic_cl(4 * nside)
    gpu_ctx.set_pixels(nside, good)
    n = gpu_ctx.npix
    f_l = capi.window_beam(4 * nside, 10.0)
    d_c = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    d_f = torch.empty_like(d_c)
    gpu_ctx.legendre_series(capi.tt_weights(cl[:lmax + 1], f_l[:lmax + 1]), d_c)
    gpu_ctx.legendre_series(capi.fiducial_weights(cl, f_l, nside, lmax), d_f)
    noise = oracle_api.mask_matrix(oracle_api.noise_matrix(nside, 0.5), good)
    d_n = torch.from_numpy(noise).cuda()
    rs = np.random.RandomState(4)
    fg = rs.standard_normal(n)
    maps = rs.standard_normal((3, n)) * 5.0
    like = Likelihood(gpu_ctx, d_c, d_f, d_n, n, foreground=fg)
    got_like, got_chi2, got_logdet = like.calculate(maps)

    C = oracle_api.unpack_symmetric(oracle_api.cl_to_cmatrix(cl[:lmax + 1], nside, 10.0, good=good)
                                    + oracle_api.fiducial_matrix(cl, nside, lmax, 10.0, good=good) + noise, n)
    Cinv = np.linalg.inv(C)
    sign, logdet = np.linalg.slogdet(C)
    assert sign > 0
    fCf = fg @ Cinv @ fg
    want_logdet = logdet - DET_OFFSET + np.log(fCf / n)
    want_chi2 = np.array([t @ Cinv @ t - (t @ Cinv @ fg) ** 2 / fCf for t in maps])
    assert abs(got_logdet - want_logdet) <= 1e-9 * abs(want_logdet)
    assert np.abs(got_chi2 - want_chi2).max() <= 1e-8 * np.abs(want_chi2).max()
    one = like.calculate(maps[1])
    assert abs(one[1] - want_chi2[1]) <= 1e-8 * abs(want_chi2[1]) and abs(one[0] - (want_chi2[1] + want_logdet)) <= 1e-8 * abs(one[0])
    # without the template
    plain = Likelihood(gpu_ctx, d_c, d_f, d_n, n)
    assert abs(plain.calculate(maps[0])[1] - maps[0] @ Cinv @ maps[0]) <= 1e-8 * abs(maps[0] @ Cinv @ maps[0])
    # not positive definite: the reference throws with this text (source/likelihood.cpp:119-124)
    with pytest.raises(ValueError, match="must be positive definite"):
        Likelihood(gpu_ctx, d_c, d_f, torch.from_numpy(-1e6 * noise).cuda(), n)

    # one element of a batched slab (element stride 16) as the covariance of a T,Q,U likelihood
    nside, lmax, nb = 4, 12, 3
    gpu_ctx.set_pixels(nside)
    n3 = 3 * gpu_ctx.npix
    f = capi.window_beam(lmax, 10.0)
    spectra = [synthetic_cl(lmax, seed=500 + b, pol=True) for b in range(nb)]
    ab = np.stack([np.stack(capi.tqu_weights(*sp, f, f)) for sp in spectra])
    slab = torch.empty(capi.slab_doubles(n3), dtype=torch.float64, device="cuda")
    gpu_ctx.tqu_batched_slab(ab, slab)
    sigma2 = 4.0
    noise3 = np.zeros(capi.packed_size(n3))
    noise3[[capi.packed_index(i, i) for i in range(n3)]] = sigma2
    d_n3 = torch.from_numpy(noise3).cuda()
    m3 = rs.standard_normal(n3) * 3.0
    for b in range(nb):
        lk = Likelihood(gpu_ctx, slab[b:], None, d_n3, n3, c_stride=capi.SLAB)
        C3 = oracle_api.unpack_symmetric(oracle_api.tqu_matrix(*spectra[b], nside, 10.0), n3) + sigma2 * np.eye(n3)
        total, chi2, logdet = lk.calculate(m3)
        want = m3 @ np.linalg.solve(C3, m3)
        assert abs(chi2 - want) <= 1e-8 * want
        assert abs(logdet - (np.linalg.slogdet(C3)[1] - DET_OFFSET)) <= 1e-9 * abs(logdet)
        lk.close()


def test_device_resident_weights_and_cuda_graph_replay(gpu_ctx, oracle_api):
    """MCMC-style use: weights live on the device, the generate call is captured once in a CUDA graph and replayed
    after the weights were overwritten in place; TT and T,Q,U."""
    torch = _torch()
    from cosmopp_b200 import capi
    nside, lmax = 4, 12
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    gpu_ctx.set_pixels(nside, good)
    n = gpu_ctx.npix
    f = capi.window_beam(lmax, 10.0)
    sets = [synthetic_cl(lmax, seed=s, pol=True) for s in (31, 32)]
    w = [np.concatenate(capi.tqu_weights(*sp, f, f)) for sp in sets]
    d_w = torch.from_numpy(w[0]).cuda()
    d_wtt = torch.from_numpy(capi.tt_weights(sets[0][0], f)).cuda()
    out = torch.empty(capi.packed_size(3 * n), dtype=torch.float64, device="cuda")
    out_tt = torch.empty(capi.packed_size(n), dtype=torch.float64, device="cuda")
    lay = gpu_ctx.tqu_layout_single(out)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    gpu_ctx.set_stream(side.cuda_stream)
    try:
        with torch.cuda.stream(side):
            gpu_ctx.tqu_dev(d_w, lmax, lay)                    # warm-up outside capture (function attributes, lazy init)
            gpu_ctx.legendre_series_dev(d_wtt, lmax, out_tt)
            side.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                gpu_ctx.tqu_dev(d_w, lmax, lay)
                gpu_ctx.legendre_series_dev(d_wtt, lmax, out_tt)
        for k in (1, 0, 1):
            d_w.copy_(torch.from_numpy(w[k]))
            d_wtt.copy_(torch.from_numpy(capi.tt_weights(sets[k][0], f)))
            torch.cuda.synchronize()
            out.fill_(float("nan"))
            out_tt.fill_(float("nan"))
            g.replay()
            torch.cuda.synchronize()
            _assert_tqu_close(out.cpu().numpy(), oracle_api.tqu_matrix(*sets[k], nside, 10.0, good=good), n)
            want = oracle_api.cl_to_cmatrix(sets[k][0], nside, 10.0, good=good)
            assert np.abs(out_tt.cpu().numpy() - want).max() <= REL_TOL * want[0]
    finally:
        gpu_ctx.set_stream(torch.cuda.current_stream().cuda_stream)
