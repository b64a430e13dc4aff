"""The mask recipe of the reference's only C-matrix test (source/test_like_low.cpp:99-118): the oracle's MT19937 /
uniform_real_distribution restatement against libstdc++ itself, and the good-pixel lists against committed goldens."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

SNIPPET = r"""
#include <cstdio>
#include <random>
int main() {
    const double pi = 3.141592653589793;
    std::mt19937 g1(1000000), g2(1000001), g3(1000002);
    std::uniform_real_distribution<double> d1(pi / 50, pi - pi / 50), d2(0, 2 * pi), d3(pi / 60, pi / 40);
    for (int i = 0; i < 25; ++i) std::printf("%.17g %.17g %.17g\n", d1(g1), d2(g2), d3(g3));
}
"""


def test_mt19937_uniform_real_matches_libstdcxx(oracle_api, tmp_path):
    src = tmp_path / "mt.cpp"
    src.write_text(SNIPPET)
    exe = str(tmp_path / "mt")
    r = subprocess.run(["g++", "-O1", "-o", exe, str(src)], capture_output=True, text=True)
    if r.returncode:
        pytest.skip("no working g++: " + r.stderr[-200:])
    want = np.array([[float(x) for x in ln.split()] for ln in subprocess.run([exe], capture_output=True, text=True).stdout.splitlines()])
    got = oracle_api.like_low_discs(25, 1000000)
    assert np.array_equal(got, want)          # bit-exact


@pytest.mark.parametrize("nside,count", [(4, 171), (8, 655), (16, 2548), (32, 10074)])
def test_like_low_good_pixels_match_golden(oracle_api, nside, count):
    good = oracle_api.good_pixels_from_mask(oracle_api.like_low_mask(nside))
    gold = np.load(os.path.join(ROOT, "tests", "golden", "like_low_good_pixels_nside%d.npy" % nside))
    assert len(good) == count and np.array_equal(good, gold)
    assert (np.diff(good) > 0).all()


def test_good_pixel_rule_threshold(oracle_api):
    # reference source/utils.cpp:45-51: strictly greater than 0.5, ascending
    mask = np.array([0.0, 0.5, 0.5000001, 1.0, -1.0, 2.0])
    assert list(oracle_api.good_pixels_from_mask(mask)) == [2, 3, 5]
    assert len(oracle_api.good_pixels_from_mask(np.zeros(12))) == 0


def test_reference_mask_fixture_golden():
    """slow_test_files/mask1.fits of the reference (Nside=32 NESTED): 9096 of 12288 pixels unmasked, first good
    indices 29,30,31,47,51, last 12260."""
    good = np.load(os.path.join(ROOT, "tests", "golden", "mask1_good_pixels.npy"))
    assert len(good) == 9096 and list(good[:5]) == [29, 30, 31, 47, 51] and good[-1] == 12260
