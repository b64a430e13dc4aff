// Drop-in replacement of the reference's include/c_matrix_generator.hpp for the C_l -> pixel covariance
// path, running on a B200 through the C ABI of cmg.h.  Signatures, defaults, ownership ("the returned matrix
// must be deleted by the caller") and error behaviour (StandardException) follow the reference
// (include/c_matrix_generator.hpp:41-162); the polarized and batched generators are additions.
//
// There is no CPU fallback: without a usable sm_100 GPU every generator throws StandardException.
#ifndef COSMO_PP_C_MATRIX_GENERATOR_HPP
#define COSMO_PP_C_MATRIX_GENERATOR_HPP

#include <cstdint>
#include <vector>

#include <c_matrix.hpp>

class WholeMatrix;      // reference include/whole_matrix.hpp; only named by the SHT-based entry points below

// P_l(n_i . n_j) for all unmasked pixel pairs (reference include/c_matrix_generator.hpp:10-36).  The reference
// precomputes and stores (lMax+1) nPix (nPix+1)/2 doubles (1.8 GB at Nside=16, lMax=47) to speed up its CPU
// generator.  Here the container only records the pixel set: value() evaluates the recurrence on demand, and the
// generators ignore it -- recomputing P_l on the GPU is faster than reading the cache.  The file format of
// writeIntoFile / the file constructor is the reference's (source/c_matrix_generator.cpp:89-162).
class LegendrePolynomialContainer
{
public:
    LegendrePolynomialContainer(int lMax, long nSide, const std::vector<int>* goodPixels = NULL);
    LegendrePolynomialContainer(const char* fileName);

    // P_l of the cosine of the angle between pixels i and j (indices into goodPixels), i <= j
    double value(int l, int j, int i) const;
    void writeIntoFile(const char* fileName) const;

    int lMax() const { return lMax_; }
    int nPix() const { return nPix_; }

private:
    int lMax_;
    int nPix_;
    std::vector<double> xyz_;                 // unit vectors when built from a pixel set
    std::vector<std::vector<double> > file_;  // [l][packed(i,j)] when read from a file
};

class CMatrixGenerator
{
public:
    // S_ij = sum_{l=2}^{lMax} C_l (2l+1)/(4 pi) B_l^2 P_l(n_i.n_j), lMax = cl.size() - 1
    // (reference source/c_matrix_generator.cpp:164-232).  fwhm in degrees; goodPixels = NULL means all pixels.
    static CMatrix* clToCMatrix(const std::vector<double>& cl, long nSide, double fwhm, const std::vector<int>* goodPixels = NULL, const LegendrePolynomialContainer* lp = NULL);
    static CMatrix* clToCMatrix(const char* clFileName, long nSide, int lMax, double fwhm, const std::vector<int>* goodPixels = NULL, const LegendrePolynomialContainer* lp = NULL);

    // terms lMax < l <= 4 nSide plus the monopole/dipole marginalisation 100 C_2 (1 + cos) B_2^2
    // (reference source/c_matrix_generator.cpp:705-772); cl must reach l = 4 nSide
    static CMatrix* getFiducialMatrix(const std::vector<double>& cl, long nSide, int lMax, double fwhm, const std::vector<int>* goodPixels = NULL, const LegendrePolynomialContainer* lp = NULL);
    static CMatrix* getFiducialMatrix(const char* clFileName, long nSide, int lMax, double fwhm, const std::vector<int>* goodPixels = NULL, const LegendrePolynomialContainer* lp = NULL);

    // diagonal white-noise matrix over the full sky (reference source/c_matrix_generator.cpp:774-787)
    static CMatrix* generateNoiseMatrix(long nSide, double noise = 1e-3);

    // Harmonic-space (WholeMatrix) routes of the reference: they are spherical-harmonic-transform bound, need
    // HEALPix C++ and lie outside the C_l -> pixel hot path; they throw StandardException here.
    static void clToWholeMatrix(const std::vector<double>& cl, WholeMatrix& wm);
    static void clToWholeMatrix(const char* clFileName, WholeMatrix& tt, WholeMatrix& te, WholeMatrix& ee);
    static CMatrix* wholeMatrixToCMatrix(const WholeMatrix& wholeMatrix, long nSide, double fwhm, double phi = 0, double theta = 0, double psi = 0);
    static CMatrix* polarizationEEWholeMatrixToCMatrix(const WholeMatrix& ee, long nSide, double fwhm, double phi = 0, double theta = 0, double psi = 0);
    static CMatrix* calculateNoiseMatrix(const char* maskFileName, const char* noiseDataFileName, double sigma0, double fwhm, long nSideOriginal = 512, double fwhmOriginal = 1);

    // ---- additions ----
    // [T;Q;U] covariance of dimension 3 nPix from TT, TE, EE, BB spectra (each indexed by l, same length);
    // rows/columns [T_0.., Q_0.., U_0..] extending the reference's [Q;U] layout
    // (source/c_matrix_generator.cpp:678-681); Q,U in the local (e_theta, e_phi) frame, HEALPix convention.
    static CMatrix* clToCMatrixPol(const std::vector<double>& clTT, const std::vector<double>& clTE, const std::vector<double>& clEE,
                                   const std::vector<double>& clBB, long nSide, double fwhm, const std::vector<int>* goodPixels = NULL);

    // Per-l window the generators multiply the Gaussian beam by.  The reference reads it from
    // HEALPIX_DATA_DIR/pixel_window_nNNNN.fits on every call (source/utils.cpp:66-170); here the directory comes from
    // setHealpixDataDir() or the HEALPIX_DATA_DIR environment variable, or the table can be given directly.
    static void setHealpixDataDir(const char* dir);
    static void setPixelWindow(long nSide, const std::vector<double>& temperature, const std::vector<double>& polarization);
    static void clearPixelWindow(long nSide);

    // which of the two tables the Q, U part of clToCMatrixPol is smoothed with: HEALPix's polarization window (default), or the
    // temperature window as in the reference's own polarization routine (source/c_matrix_generator.cpp:534)
    static void setPolarizationUsesTemperatureWindow(bool on);

    // GPU the calling THREAD's generator calls run on (default 0): several host threads -- one per chain or rank -- can each
    // drive their own GPU from one process; calls on one GPU are serialised, calls on different GPUs run concurrently
    static void setDevice(int device);
    // true (default): the generators return matrices that live in GPU memory and reach the host lazily (CMatrix);
    // false: results are copied to host memory inside the call, as the reference's objects are
    static void setDeviceResident(bool on);
    // bytes the calling thread's GPU context has moved over PCIe so far (cmg_transfer_counters): what a test asserts on
    static void transferCounters(long long& hostToDevice, long long& deviceToHost);
};

#endif
