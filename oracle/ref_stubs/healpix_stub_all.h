/* TEST INFRASTRUCTURE ONLY.
 * Declaration-only stand-ins for the HEALPix C++ / cfitsio headers that the
 * reference sources include.  They exist so that the reference's own
 * c_matrix.cpp / c_matrix_generator.cpp compile unmodified into oracle/_ref/.
 * None of the spherical-harmonic-transform entry points is implemented: the
 * TT path (clToCMatrix, getFiducialMatrix, generateNoiseMatrix, maskMatrix)
 * never calls them; the link shim (ref_shim.cpp) aborts if anything does.
 */
#ifndef ORACLE_HEALPIX_STUB_ALL_H
#define ORACLE_HEALPIX_STUB_ALL_H

#include <complex>
#include <vector>
#include <string>
#include <cstddef>

template <typename T> class xcomplex : public std::complex<T>
{
public:
    xcomplex() : std::complex<T>() {}
    xcomplex(const T& r, const T& i = T()) : std::complex<T>(r, i) {}
    xcomplex(const std::complex<T>& c) : std::complex<T>(c) {}
    xcomplex conj() const { return xcomplex(std::conj(static_cast<const std::complex<T>&>(*this))); }
    xcomplex& operator*=(const T& f) { std::complex<T>::operator*=(f); return *this; }
};
template <typename T> inline xcomplex<T> operator+(const xcomplex<T>& a, const xcomplex<T>& b)
{ return xcomplex<T>(static_cast<const std::complex<T>&>(a) + static_cast<const std::complex<T>&>(b)); }
template <typename T> inline xcomplex<T> operator-(const xcomplex<T>& a, const xcomplex<T>& b)
{ return xcomplex<T>(static_cast<const std::complex<T>&>(a) - static_cast<const std::complex<T>&>(b)); }
template <typename T> inline xcomplex<T> operator*(const xcomplex<T>& a, const xcomplex<T>& b)
{ return xcomplex<T>(static_cast<const std::complex<T>&>(a) * static_cast<const std::complex<T>&>(b)); }
template <typename T> inline xcomplex<T> operator*(const T& f, const xcomplex<T>& b)
{ return xcomplex<T>(f * static_cast<const std::complex<T>&>(b)); }
template <typename T> inline xcomplex<T> operator*(const xcomplex<T>& b, const T& f)
{ return xcomplex<T>(f * static_cast<const std::complex<T>&>(b)); }

template <typename T> class arr
{
public:
    arr() {}
    explicit arr(std::size_t n) : v_(n) {}
    T& operator[](std::size_t i) { return v_[i]; }
    const T& operator[](std::size_t i) const { return v_[i]; }
    std::size_t size() const { return v_.size(); }
private:
    std::vector<T> v_;
};

enum Healpix_Ordering_Scheme { RING, NEST };

class rotmatrix
{
public:
    rotmatrix() {}
    rotmatrix(double, double, double, double, double, double, double, double, double) {}
};

template <typename T> class Alm
{
public:
    Alm() : lmax_(0), mmax_(0) {}
    Alm(int lmax, int mmax) { Set(lmax, mmax); }
    void Set(int lmax, int mmax) { lmax_ = lmax; mmax_ = mmax; v_.assign(std::size_t(lmax + 1) * (mmax + 1), T()); }
    T& operator()(int l, int m) { return v_[std::size_t(l) * (mmax_ + 1) + m]; }
    const T& operator()(int l, int m) const { return v_[std::size_t(l) * (mmax_ + 1) + m]; }
    int Lmax() const { return lmax_; }
    int Mmax() const { return mmax_; }
private:
    int lmax_, mmax_;
    std::vector<T> v_;
};

template <typename T> class Healpix_Map
{
public:
    Healpix_Map() : nside_(0), scheme_(RING) {}
    void SetNside(long nside, Healpix_Ordering_Scheme s) { nside_ = nside; scheme_ = s; v_.assign(std::size_t(12) * nside * nside, T()); }
    T& operator[](std::size_t i) { return v_[i]; }
    const T& operator[](std::size_t i) const { return v_[i]; }
    long Nside() const { return nside_; }
    long Npix() const { return long(v_.size()); }
    Healpix_Ordering_Scheme Scheme() const { return scheme_; }
    void swap_scheme();                       /* not implemented: link shim aborts */
private:
    long nside_;
    Healpix_Ordering_Scheme scheme_;
    std::vector<T> v_;
};

template <typename T> void alm2map(const Alm<xcomplex<T> >&, Healpix_Map<T>&);
template <typename T> void alm2map_pol(const Alm<xcomplex<T> >&, const Alm<xcomplex<T> >&, const Alm<xcomplex<T> >&,
                                       Healpix_Map<T>&, Healpix_Map<T>&, Healpix_Map<T>&);
template <typename T> void map2alm_iter(const Healpix_Map<T>&, Alm<xcomplex<T> >&, int, const arr<double>&);
template <typename T> void rotate_alm(Alm<xcomplex<T> >&, const rotmatrix&);
template <typename T> void rotate_alm(Alm<xcomplex<T> >&, Alm<xcomplex<T> >&, Alm<xcomplex<T> >&, const rotmatrix&);

class fitshandle {};

#endif
