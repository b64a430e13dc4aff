"""Python mirror of the reference's CMatrix / CMatrixGenerator interface over the C ABI.

Same names, argument meaning and error behaviour as reference include/c_matrix.hpp:10-83 and
include/c_matrix_generator.hpp:41-162 for the hot path, so the parity tests read like the reference's
own test (source/test_like_low.cpp:183-186).  The C++ drop-in headers under include/ are the primary
host interface; this mirror exists for the Python tests and bench.  All computing goes through
libcosmopp_b200.so -- there is no CPU path here.
"""
import struct

import numpy as np

from . import capi


class StandardException(Exception):
    """Mirror of the reference's exception type (include/exception_handler.hpp:10-31)."""


class CMatrix:
    """Symmetric nPix x nPix matrix, packed upper triangle, index j(j+1)/2+i for i<=j
    (reference source/c_matrix.cpp:27-39; 64-bit indices here)."""

    def __init__(self, arg):
        self._comment = ""
        if isinstance(arg, str):
            self.readFromFile(arg)
        else:
            n = int(arg)
            if n <= 0:
                raise StandardException("the number of pixels must be positive.")
            self._n = n
            self._m = np.zeros(capi.packed_size(n))

    @classmethod
    def fromPacked(cls, n, packed, comment=""):
        m = cls.__new__(cls)
        m._n = int(n)
        m._m = packed
        m._comment = comment
        return m

    def getNPix(self):
        return self._n

    def packed(self):
        return self._m

    def _index(self, i, j):
        if not (0 <= i < self._n) or not (0 <= j < self._n):
            raise StandardException("invalid index")
        if i > j:
            i, j = j, i
        return j * (j + 1) // 2 + i

    def element(self, i, j):
        return float(self._m[self._index(i, j)])

    def setElement(self, i, j, v):
        self._m[self._index(i, j)] = v

    def comment(self):
        return self._comment

    def setComment(self, s):
        self._comment = s

    # binary format of reference source/c_matrix.cpp:41-85: int32 nPix, packed doubles, int32 len, bytes
    def writeIntoFile(self, fileName):
        try:
            f = open(fileName, "wb")
        except OSError:
            raise StandardException("Cannot write into output file %s." % fileName)
        with f:
            f.write(struct.pack("<i", self._n))
            f.write(np.ascontiguousarray(self._m, dtype="<f8").tobytes())
            c = self._comment.encode()
            f.write(struct.pack("<i", len(c)))
            f.write(c)

    def readFromFile(self, fileName):
        try:
            f = open(fileName, "rb")
        except OSError:
            raise StandardException("Covariance matrix file %s cannot be read." % fileName)
        with f:
            (n,) = struct.unpack("<i", f.read(4))
            if n <= 0:
                raise StandardException("the number of pixels must be positive.")
            self._n = n
            self._m = np.frombuffer(f.read(8 * capi.packed_size(n)), dtype="<f8").copy()
            (cl,) = struct.unpack("<i", f.read(4))
            self._comment = f.read(cl).decode()

    # text format of reference source/c_matrix.cpp:87-158: nPix, comment line, "i\tj\tvalue" rows (j outer)
    def writeIntoTextFile(self, fileName):
        try:
            f = open(fileName, "w")
        except OSError:
            raise StandardException("Cannot write into output file %s." % fileName)
        with f:
            f.write("%d\n%s\n" % (self._n, self._comment))
            k = 0
            for j in range(self._n):
                for i in range(j + 1):
                    f.write("%d\t%d\t%s\n" % (i, j, _cxx_double(self._m[k])))
                    k += 1

    def readFromTextFile(self, fileName):
        """Reads what writeIntoTextFile writes: nPix, comment line, then "i j value" rows.  (The reference's reader,
        source/c_matrix.cpp:114-158, tokenises on white space and only picks up the first number of a row.)"""
        try:
            f = open(fileName, "r")
        except OSError:
            raise StandardException("Cannot read the input file %s." % fileName)
        with f:
            n = int(f.readline().split()[0])
            if n <= 0:
                raise StandardException("the number of pixels must be positive.")
            self._n = n
            self._m = np.zeros(capi.packed_size(n))
            self._comment = f.readline().rstrip("\n")
            for line in f:
                parts = line.split()
                if len(parts) < 3:
                    continue
                i, j, v = int(parts[0]), int(parts[1]), float(parts[2])
                if not (0 <= i < n):
                    raise StandardException("Invalid index i = %d." % i)
                if not (0 <= j < n):
                    raise StandardException("Invalid index j = %d." % j)
                self._m[self._index(i, j)] = v

    def maskMatrix(self, goodPixels):
        """reference source/c_matrix.cpp:182-201 (gather; host-side here because the object is host-resident)."""
        g = np.asarray(goodPixels, dtype=np.int64)
        n = len(g)
        out = np.empty(capi.packed_size(n))
        k = 0
        for b in range(n):
            gi = np.minimum(g[:b + 1], g[b])
            gj = np.maximum(g[:b + 1], g[b])
            out[k:k + b + 1] = self._m[gj * (gj + 1) // 2 + gi]
            k += b + 1
        self._n = n
        self._m = out


def _cxx_double(v):
    """operator<<(double) with default precision 6 (%g), as the reference's text writer."""
    return "%g" % v


class CMatrixGenerator:
    """Static generators of reference include/c_matrix_generator.hpp:41-162 (hot-path subset) plus
    the polarized / batched additions, each running the CUDA kernels through the C ABI."""

    _contexts = {}

    @classmethod
    def context(cls, device=0):
        ctx = cls._contexts.get(device)
        if ctx is None:
            ctx = capi.Context(device)
            cls._contexts[device] = ctx
        return ctx

    @staticmethod
    def _pinned(n):
        import torch
        return torch.empty(n, dtype=torch.float64, pin_memory=True)

    @classmethod
    def clToCMatrix(cls, cl, nSide, fwhm, goodPixels=None, lp=None, pixelWindow=None, device=0):
        """reference source/c_matrix_generator.cpp:164-232.  `lp` (a LegendrePolynomialContainer) is
        accepted and ignored: recomputing P_l on the GPU beats reading 1.8-58 GB of cached values."""
        if len(cl) == 0:
            raise StandardException("CHECK FAILED")
        ctx = cls.context(device)
        ctx.set_pixels(nSide, goodPixels)
        out = cls._pinned(capi.packed_size(ctx.npix))
        ctx.cl_to_cmatrix(cl, fwhm, out, pixwin=pixelWindow)
        return CMatrix.fromPacked(ctx.npix, out.numpy())

    @classmethod
    def getFiducialMatrix(cls, cl, nSide, lMax, fwhm, goodPixels=None, lp=None, pixelWindow=None, device=0):
        """reference source/c_matrix_generator.cpp:705-772"""
        if len(cl) < 4 * nSide + 1:
            raise StandardException("CHECK FAILED")
        ctx = cls.context(device)
        ctx.set_pixels(nSide, goodPixels)
        out = cls._pinned(capi.packed_size(ctx.npix))
        ctx.fiducial_matrix(cl, lMax, fwhm, out, pixwin=pixelWindow)
        return CMatrix.fromPacked(ctx.npix, out.numpy(), "fiducial matrix")

    @staticmethod
    def generateNoiseMatrix(nSide, noise=1e-3):
        """reference source/c_matrix_generator.cpp:774-787"""
        n = 12 * nSide * nSide
        return CMatrix.fromPacked(n, capi.noise_matrix(n, noise), "noise matrix")

    @classmethod
    def clToCMatrixPol(cls, clTT, clTE, clEE, clBB, nSide, fwhm, goodPixels=None, pixelWindowT=None, pixelWindowP=None,
                       device=0):
        """[T;Q;U] covariance (addition to the reference API; BASELINE.json north_star)."""
        ctx = cls.context(device)
        ctx.set_pixels(nSide, goodPixels)
        out = cls._pinned(capi.packed_size(3 * ctx.npix))
        ctx.cl_to_cmatrix_pol(clTT, clTE, clEE, clBB, fwhm, out, pixwinT=pixelWindowT, pixwinP=pixelWindowP)
        return CMatrix.fromPacked(3 * ctx.npix, out.numpy())
