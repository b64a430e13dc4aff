"""Regenerates the committed golden vectors from the reference's own object code (oracle/_ref, built from
/root/reference by oracle/Makefile) -- run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

Outputs (small .npz files next to this script):
  ref_tt_nside4.npz        clToCMatrix full sky, Nside=4, lMax=10, fwhm=10 deg   (reference c_matrix_generator.cpp:164-232)
  ref_tt_nside8_masked.npz clToCMatrix with the test_like_low mask, Nside=8, lMax=16
  ref_fiducial_nside4.npz  getFiducialMatrix Nside=4, lMax=8 (terms 9..16 + monopole/dipole)  (:705-772)
  ref_noise_masked.npz     generateNoiseMatrix + maskMatrix, Nside=2                  (:774-787, c_matrix.cpp:182-201)
  ref_legendre.npz         Math::Legendre::calculate at the points of the reference's test_legendre.cpp plus a grid
  mask1_good_pixels.npy    good pixels of the reference fixture slow_test_files/mask1.fits (Nside=32 NESTED) via Utils::readMask rule
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from cosmopp_b200.synthetic import synthetic_cl  # noqa: E402
from oracle import api  # noqa: E402


def main():
    assert api.have_ref(), "build oracle/_ref first (make -C oracle)"
    cl = synthetic_cl(10)
    np.savez_compressed(os.path.join(HERE, "ref_tt_nside4.npz"), cl=cl, nside=4, fwhm=10.0, packed=api.ref_cl_to_cmatrix(cl, 4, 10.0))
    good = api.good_pixels_from_mask(api.like_low_mask(8))
    cl = synthetic_cl(16)
    np.savez_compressed(os.path.join(HERE, "ref_tt_nside8_masked.npz"), cl=cl, nside=8, fwhm=10.0, good=good,
                        packed=api.ref_cl_to_cmatrix(cl, 8, 10.0, good=good))
    cl = synthetic_cl(16)
    np.savez_compressed(os.path.join(HERE, "ref_fiducial_nside4.npz"), cl=cl, nside=4, lmax=8, fwhm=10.0,
                        packed=api.ref_fiducial_matrix(cl, 4, 8, 10.0))
    g2 = np.array([0, 3, 4, 17, 30, 47], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "ref_noise_masked.npz"), nside=2, noise=0.01, good=g2, packed=api.ref_noise_matrix_masked(2, 0.01, g2))
    ls = np.array([0, 1, 2, 3, 4, 10, 47, 64, 96, 192, 256, 1000], dtype=np.int64)
    xs = np.concatenate([np.linspace(-1, 1, 41), [2.0, 0.5, 0.1, -0.25, -0.5]])
    vals = np.array([[api.ref().ref_legendre(int(l), float(x)) for x in xs] for l in ls])
    np.savez_compressed(os.path.join(HERE, "ref_legendre.npz"), l=ls, x=xs, value=vals)
    mask1 = "/root/reference/slow_test_files/mask1.fits"
    if os.path.exists(mask1):
        # FITS binary table, one 1024D column (parsed with numpy here; the product's own reader is tested against this list)
        raw = open(mask1, "rb").read()
        pos = 0
        hdus = []
        while pos < len(raw):
            hdr = {}
            while True:
                block = raw[pos:pos + 2880].decode("ascii", "replace")
                pos += 2880
                done = False
                for c in range(36):
                    card = block[80 * c:80 * c + 80]
                    if card.startswith("END"):
                        done = True
                        break
                    if card[8:10] == "= ":
                        hdr[card[:8].strip()] = card[10:].split("/")[0].strip().strip("'").strip()
                if done:
                    break
            n = 0
            if int(hdr.get("NAXIS", 0)) > 0:
                n = abs(int(hdr["BITPIX"])) // 8
                for a in range(1, int(hdr["NAXIS"]) + 1):
                    n *= int(hdr["NAXIS%d" % a])
            hdus.append((hdr, pos, n))
            pos += (n + 2879) // 2880 * 2880
        hdr, start, n = hdus[1]
        assert hdr["TFORM1"].endswith("D") and hdr["ORDERING"].upper().startswith("NEST")
        mask = np.frombuffer(raw[start:start + n], dtype=">f8").astype(np.float64)
        np.save(os.path.join(HERE, "mask1_good_pixels.npy"), api.good_pixels_from_mask(mask).astype(np.int32))
        print("mask1:", len(mask), "pixels,", int((mask > 0.5).sum()), "good")
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
