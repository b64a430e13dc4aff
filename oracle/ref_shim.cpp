// TEST INFRASTRUCTURE ONLY -- link shim for oracle/_ref/libcosmopp_ref.so.
//
// The reference's own c_matrix.cpp / c_matrix_generator.cpp / whole_matrix.cpp /
// macros.cpp are compiled unmodified from /root/reference (see Makefile).  This
// file supplies what they link against but cannot be built here:
//   * Utils::{beamFunction, readPixelWindowFunction, readMask, readClFromFile}
//     (reference source/utils.cpp needs cfitsio + HEALPix C++; restated from
//     source/utils.cpp:25-52,54-64,66-170,172-218 with the pixel window taken
//     from an injected vector instead of HEALPIX_DATA_DIR/pixel_window_nNNNN.fits)
//   * aborting definitions of the HEALPix SHT templates (never reached on the TT path)
//   * a C ABI (ref_*) so Python tests and bench.py can drive the reference objects.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <macros.hpp>
#include <exception_handler.hpp>
#include <math_constants.hpp>
#include <legendre.hpp>
#include <utils.hpp>
#include <c_matrix.hpp>
#include <c_matrix_generator.hpp>

#include "chealpix.h"

namespace
{
std::vector<double> g_pixwin_T, g_pixwin_P;   // empty => window == 1 for every l
std::string g_last_error;

[[noreturn]] void sht_unavailable(const char* what)
{
    std::fprintf(stderr, "oracle/_ref: %s needs HEALPix C++ (libsharp), which is not available\n", what);
    std::abort();
}
}

// ---- Utils (restated; reference source/utils.cpp) -------------------------------------------

double Utils::beamFunction(int l, double fwhm)
{
    if(fwhm == 0)
        return 1.0;
    const double sigma = std::sqrt(8 * std::log(2.0)) / (fwhm * Math::pi / 180);
    return std::exp(-l * (l + 1) / (2 * sigma * sigma));     // int product, as utils.cpp:63
}

void Utils::readPixelWindowFunction(std::vector<double>& f, long nSide, int lMax, double fwhm, bool polarization)
{
    (void)nSide;
    const std::vector<double>& w = polarization ? g_pixwin_P : g_pixwin_T;
    if(!w.empty() && (int)w.size() < lMax + 1)
    {
        StandardException exc;
        std::stringstream s;
        s << "The injected pixel window contains values only up to l = " << w.size() - 1 << ". Cannot read up to lMax = " << lMax << ".";
        exc.set(s.str());
        throw exc;
    }
    f.resize(lMax + 1);
    for(int l = 0; l <= lMax; ++l)
    {
        f[l] = (w.empty() ? 1.0 : w[l]);
        f[l] *= Utils::beamFunction(l, fwhm);
    }
}

void Utils::readMask(const char*, long&, std::vector<int>&)
{
    StandardException exc;
    exc.set("oracle/_ref: Utils::readMask needs cfitsio; pass goodPixels explicitly");
    throw exc;
}

void Utils::readClFromFile(const char* fileName, std::vector<double>& cl, bool hasL, bool isDl)
{
    std::ifstream in(fileName);
    if(!in)
    {
        StandardException exc;
        exc.set(std::string("Cannot open the input file ") + fileName + ".");
        throw exc;
    }
    cl.clear();
    std::string line;
    int l = 0;
    while(std::getline(in, line))
    {
        if(line.empty())
            break;
        std::stringstream str(line);
        double val;
        if(hasL)
        {
            int ll;
            str >> ll;
        }
        str >> val;
        if(isDl && l)
            val *= (2 * Math::pi / (l * (l + 1)));
        cl.push_back(val);
        ++l;
    }
}

// ---- HEALPix SHT entry points: declared in the stub headers, never reached on the TT path --------

template<> void Healpix_Map<double>::swap_scheme() { sht_unavailable("Healpix_Map::swap_scheme"); }
template<> void alm2map(const Alm<xcomplex<double> >&, Healpix_Map<double>&) { sht_unavailable("alm2map"); }
template<> void alm2map_pol(const Alm<xcomplex<double> >&, const Alm<xcomplex<double> >&, const Alm<xcomplex<double> >&,
                            Healpix_Map<double>&, Healpix_Map<double>&, Healpix_Map<double>&) { sht_unavailable("alm2map_pol"); }
template<> void map2alm_iter(const Healpix_Map<double>&, Alm<xcomplex<double> >&, int, const arr<double>&) { sht_unavailable("map2alm_iter"); }
template<> void rotate_alm(Alm<xcomplex<double> >&, const rotmatrix&) { sht_unavailable("rotate_alm"); }
template<> void rotate_alm(Alm<xcomplex<double> >&, Alm<xcomplex<double> >&, Alm<xcomplex<double> >&, const rotmatrix&) { sht_unavailable("rotate_alm"); }

// ---- C ABI over the reference objects ---------------------------------------------------------

namespace
{
void copyPacked(const CMatrix& m, double* out)
{
    const int n = m.getNPix();
    long k = 0;
    for(int j = 0; j < n; ++j)
        for(int i = 0; i <= j; ++i)
            out[k++] = m.element(i, j);
}
}

extern "C"
{

const char* ref_last_error() { return g_last_error.c_str(); }

void ref_set_pixel_window(const double* wT, int nT, const double* wP, int nP)
{
    g_pixwin_T.assign(wT, wT + (wT ? nT : 0));
    g_pixwin_P.assign(wP, wP + (wP ? nP : 0));
}

double ref_legendre(unsigned int l, double x)
{
    Math::Legendre leg;
    return leg.calculate(l, x);
}

double ref_beam_function(int l, double fwhm) { return Utils::beamFunction(l, fwhm); }

// CMatrixGenerator::clToCMatrix (reference source/c_matrix_generator.cpp:164-232).
// good == NULL => full sky.  out_packed holds n(n+1)/2 doubles, column j at j(j+1)/2.
int ref_cl_to_cmatrix(const double* cl, int lMax, long nSide, double fwhm, const int* good, int nGood, double* out_packed)
{
    try
    {
        std::vector<double> clv(cl, cl + lMax + 1);
        std::vector<int> gp;
        if(good) gp.assign(good, good + nGood);
        CMatrix* m = CMatrixGenerator::clToCMatrix(clv, nSide, fwhm, good ? &gp : NULL, NULL);
        copyPacked(*m, out_packed);
        delete m;
        return 0;
    }
    catch(std::exception& e) { g_last_error = e.what(); return 1; }
}

// same, through a LegendrePolynomialContainer (reference source/c_matrix_generator.cpp:30-87,221)
int ref_cl_to_cmatrix_lp(const double* cl, int lMax, long nSide, double fwhm, const int* good, int nGood, double* out_packed)
{
    try
    {
        std::vector<double> clv(cl, cl + lMax + 1);
        std::vector<int> gp;
        if(good) gp.assign(good, good + nGood);
        LegendrePolynomialContainer lp(lMax, nSide, good ? &gp : NULL);
        CMatrix* m = CMatrixGenerator::clToCMatrix(clv, nSide, fwhm, good ? &gp : NULL, &lp);
        copyPacked(*m, out_packed);
        delete m;
        return 0;
    }
    catch(std::exception& e) { g_last_error = e.what(); return 1; }
}

// CMatrixGenerator::getFiducialMatrix (reference source/c_matrix_generator.cpp:705-772); cl has 4*nSide+1 entries
int ref_fiducial_matrix(const double* cl, int nCl, long nSide, int lMax, double fwhm, const int* good, int nGood, double* out_packed)
{
    try
    {
        std::vector<double> clv(cl, cl + nCl);
        std::vector<int> gp;
        if(good) gp.assign(good, good + nGood);
        CMatrix* m = CMatrixGenerator::getFiducialMatrix(clv, nSide, lMax, fwhm, good ? &gp : NULL, NULL);
        copyPacked(*m, out_packed);
        delete m;
        return 0;
    }
    catch(std::exception& e) { g_last_error = e.what(); return 1; }
}

// generateNoiseMatrix + maskMatrix (reference source/c_matrix_generator.cpp:774-787, source/c_matrix.cpp:182-201)
int ref_noise_matrix_masked(long nSide, double noise, const int* good, int nGood, double* out_packed)
{
    try
    {
        CMatrix* m = CMatrixGenerator::generateNoiseMatrix(nSide, noise);
        if(good)
        {
            std::vector<int> gp(good, good + nGood);
            m->maskMatrix(gp);
        }
        copyPacked(*m, out_packed);
        delete m;
        return 0;
    }
    catch(std::exception& e) { g_last_error = e.what(); return 1; }
}

// CMatrix packed index as the reference computes it (source/c_matrix.cpp:27-39), via element() aliasing
long ref_packed_index(int nPix, int i, int j)
{
    CMatrix m(nPix);
    double* base = &m.element(0, 0);
    return (long)(&m.element(i, j) - base);
}

// CMatrix::maskMatrix on an arbitrary packed matrix (source/c_matrix.cpp:182-201)
int ref_mask_matrix(int nPix, const double* in_packed, const int* good, int nGood, double* out_packed)
{
    try
    {
        CMatrix m(nPix);
        long k = 0;
        for(int j = 0; j < nPix; ++j)
            for(int i = 0; i <= j; ++i)
                m.element(i, j) = in_packed[k++];
        std::vector<int> gp(good, good + nGood);
        m.maskMatrix(gp);
        copyPacked(m, out_packed);
        return 0;
    }
    catch(std::exception& e) { g_last_error = e.what(); return 1; }
}

// CMatrix binary / text writers (source/c_matrix.cpp:66-112), used to pin the drop-in file formats
int ref_write_cmatrix(int nPix, const double* in_packed, const char* comment, const char* binFile, const char* textFile)
{
    try
    {
        CMatrix m(nPix);
        long k = 0;
        for(int j = 0; j < nPix; ++j)
            for(int i = 0; i <= j; ++i)
                m.element(i, j) = in_packed[k++];
        m.comment() = comment;
        if(binFile) m.writeIntoFile(binFile);
        if(textFile) m.writeIntoTextFile(textFile);
        return 0;
    }
    catch(std::exception& e) { g_last_error = e.what(); return 1; }
}

int ref_read_cmatrix(const char* binFile, int* nPix, double* out_packed, long capacity, char* comment, int commentCap)
{
    try
    {
        CMatrix m(binFile);
        *nPix = m.getNPix();
        const long need = (long)m.getNPix() * (m.getNPix() + 1) / 2;
        if(need > capacity) { g_last_error = "capacity"; return 2; }
        copyPacked(m, out_packed);
        std::strncpy(comment, m.comment().c_str(), commentCap - 1);
        comment[commentCap - 1] = 0;
        return 0;
    }
    catch(std::exception& e) { g_last_error = e.what(); return 1; }
}

// LegendrePolynomialContainer file writer (source/c_matrix_generator.cpp:136-162)
int ref_write_legendre_container(int lMax, long nSide, const int* good, int nGood, const char* file)
{
    try
    {
        std::vector<int> gp;
        if(good) gp.assign(good, good + nGood);
        LegendrePolynomialContainer lp(lMax, nSide, good ? &gp : NULL);
        lp.writeIntoFile(file);
        return 0;
    }
    catch(std::exception& e) { g_last_error = e.what(); return 1; }
}

} // extern "C"
