"""The C-ABI shared library: it loads on a machine without a GPU, exports every symbol include/cmg.h declares, its
host-side entry points agree with the oracle, and compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, synthetic_cl
from cosmopp_b200 import capi

pytestmark = pytest.mark.skipif(not os.path.exists(capi.library_path()), reason="library not built (run __graft_entry__.build())")


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cmg.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmg_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    L = ctypes.CDLL(capi.library_path())
    names = declared_symbols()
    assert len(names) >= 40
    for name in names:
        assert hasattr(L, name), "include/cmg.h declares %s but the library does not export it" % name
    assert set(names) == set(capi.exported_names()), set(names) ^ set(capi.exported_names())


def test_layout_struct_matches_header():
    # int32 n_parts, own; int64 begin[17]; ptr[16][3]; int32 kind[16]; int64 ld[16]; int64 row0[16]
    assert ctypes.sizeof(capi.TquLayout) == 8 + 17 * 8 + 16 * 3 * 8 + 16 * 4 + 16 * 8 + 16 * 8
    assert capi.TquLayout.begin.offset == 8 and capi.TquLayout.kind.offset == 8 + 17 * 8 + 16 * 3 * 8


def test_host_entry_points_match_oracle(oracle_api):
    assert capi.library().cmg_nside2npix(64) == 49152
    assert capi.packed_size(147456) == 10871709696 and capi.packed_index(5, 3) == capi.packed_index(3, 5) == 18
    for i, j in [(0, 0), (7, 100000), (100000, 7), (46341, 46341), (147455, 147455)]:
        assert capi.packed_index(i, j) == oracle_api.lib().orc_packed_index(i, j)       # beyond the reference's int32 range
    for fwhm in (0.0, 10.0, 0.5):
        assert np.array_equal(capi.window_beam(64, fwhm), oracle_api.window_beam(64, fwhm))
    pw = np.linspace(1.0, 0.7, 65)
    assert np.array_equal(capi.window_beam(64, 10.0, pw), oracle_api.window_beam(64, 10.0, pw))
    mask = np.array([0.2, 0.51, 1.0, 0.5, 0.0, 3.0])
    assert list(capi.good_pixels_from_mask(mask)) == list(oracle_api.good_pixels_from_mask(mask)) == [1, 2, 5]
    assert np.array_equal(capi.noise_matrix(48, 0.01), oracle_api.noise_matrix(2, 0.01))


def test_weights_follow_reference_expressions():
    lmax = 12
    cl = synthetic_cl(lmax)
    f = capi.window_beam(lmax, 10.0)
    a = capi.tt_weights(cl, f)
    l = np.arange(lmax + 1)
    want = cl * (2 * l + 1) / (4 * 3.141592653589793) * f * f         # reference c_matrix_generator.cpp:192,222
    assert a[0] == a[1] == 0 and np.allclose(a[2:], want[2:], rtol=1e-15, atol=0)
    nside, lm = 4, 8
    clf = synthetic_cl(4 * nside)
    ff = capi.window_beam(4 * nside, 10.0)
    af = capi.fiducial_weights(clf, ff, nside, lm)
    md = 100 * clf[2] * ff[2] * ff[2]                                   # :762, (1+z) = P_0 + P_1
    assert af[0] == md and af[1] == md and not af[2:lm + 1].any() and (af[lm + 1:] > 0).all()
    tt, te, ee, bb = synthetic_cl(lmax, pol=True)
    att, ate, aee, abb = capi.tqu_weights(tt, te, ee, bb, f, 0.5 * f)
    w = (2 * l + 1) / (4 * 3.141592653589793)
    assert np.allclose(ate[2:], (te * w * f * 0.5 * f)[2:], rtol=1e-15) and np.allclose(abb[2:], (bb * w * 0.25 * f * f)[2:], rtol=1e-15)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this check is for machines without a GPU")
    assert capi.library().cmg_device_count() == 0
    with pytest.raises(capi.CmgError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under cosmopp_b200/ may reference it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "cosmopp_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or fn == "Makefile":
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "oracle" not in text.replace("oracle/ ", ""), "%s mentions the oracle" % fn
