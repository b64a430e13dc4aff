"""Python mirror of the CMatrix interface (host side only here; the generators are covered by the GPU tests)."""
import numpy as np
import pytest

from cosmopp_b200.generator import CMatrix, CMatrixGenerator, StandardException


def test_cmatrix_file_round_trips(tmp_path):
    m = CMatrix(6)
    k = 0
    for j in range(6):
        for i in range(j + 1):
            m.setElement(i, j, 0.5 + k)
            k += 1
    m.setComment("a comment")
    assert m.element(4, 2) == m.element(2, 4) == m.packed()[4 * 5 // 2 + 2]
    m.writeIntoFile(str(tmp_path / "m.dat"))
    r = CMatrix(str(tmp_path / "m.dat"))
    assert r.getNPix() == 6 and r.comment() == "a comment" and np.array_equal(r.packed(), m.packed())
    m.writeIntoTextFile(str(tmp_path / "m.txt"))
    t = CMatrix(2)
    t.readFromTextFile(str(tmp_path / "m.txt"))
    assert t.getNPix() == 6 and t.comment() == "a comment" and np.allclose(t.packed(), m.packed(), rtol=1e-5)
    with pytest.raises(StandardException):
        CMatrix(str(tmp_path / "missing.dat"))
    with pytest.raises(StandardException):
        CMatrix(0)


def test_mask_matrix_and_noise(oracle_api):
    noise = CMatrixGenerator.generateNoiseMatrix(2, 0.1)
    assert noise.getNPix() == 48 and noise.comment() == "noise matrix"
    good = [0, 5, 6, 40]
    full = noise.packed().copy()
    noise.maskMatrix(good)
    assert np.array_equal(noise.packed(), oracle_api.mask_matrix(full, good))
    rs = np.random.RandomState(3)
    m = CMatrix.fromPacked(9, rs.standard_normal(45))
    want = oracle_api.mask_matrix(m.packed(), [8, 1, 4])          # unsorted lists work too (gather semantics)
    m.maskMatrix([8, 1, 4])
    assert np.array_equal(m.packed(), want)
