// CMatrixGenerator / LegendrePolynomialContainer of include/c_matrix_generator.hpp and the pixel-window part of
// Utils: thin host wrappers that marshal the reference's arguments into the C ABI (cmg.h) and turn non-zero
// statuses into StandardException, the reference's error convention.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <mutex>
#include <sstream>
#include <string>
#include <vector>

#include <c_matrix_generator.hpp>
#include <cmg.h>
#include <exception_handler.hpp>
#include <utils.hpp>

#include "dropin_internal.hpp"

namespace
{
[[noreturn]] void raise(const std::string& text) { throw StandardException(text); }

std::mutex g_mutex;                              // guards the registries below (not the GPU work: that is per device)
thread_local int t_device = 0;                   // CMatrixGenerator::setDevice is per host thread: one rank / chain per thread can
                                                 // each drive its own GPU from one process
bool g_deviceResident = true;
bool g_polarizationUsesTemperatureWindow = false;
std::string g_healpixDir;
struct Window { std::vector<double> t, p; };
std::map<long, Window> g_windows;
std::map<int, DropinDevice*> g_devices;

} // namespace

DropinDevice& cmgDropinDevice(int device)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    std::map<int, DropinDevice*>::iterator it = g_devices.find(device);
    if(it != g_devices.end())
        return *it->second;
    cmg_ctx* ctx = NULL;
    if(cmg_create(&ctx, device) != CMG_OK)
        raise(std::string("CMatrixGenerator: ") + cmg_last_error(NULL));
    DropinDevice* d = new DropinDevice;
    d->ctx = ctx;
    d->device = device;
    g_devices[device] = d;                       // lives until the process ends (objects on that GPU may outlive any scope)
    return *d;
}

int cmgDropinCurrentDevice() { return t_device; }

namespace
{
void check(cmg_ctx* ctx, cmg_status s)
{
    if(s != CMG_OK)
        raise(std::string("CMatrixGenerator: ") + cmg_last_error(ctx));
}

void setPixels(cmg_ctx* ctx, long nSide, const std::vector<int>* goodPixels)
{
    if(goodPixels)
    {
        if(goodPixels->empty())
            raise("CMatrixGenerator: the list of unmasked pixels is empty");
        check(ctx, cmg_set_pixels(ctx, nSide, &(*goodPixels)[0], static_cast<std::int64_t>(goodPixels->size())));
    }
    else
        check(ctx, cmg_set_pixels(ctx, nSide, NULL, 0));
}

// what cfitsio hands back when a numeric table cell is read as a string and parsed again, which is how the
// reference reads the window (source/utils.cpp:139-160): 7 significant digits for 'E', 16 for 'D' columns.
// cfitsio is not available here, so this detail is restated from its documented default display formats.
double throughText(double v, char type)
{
    char buf[64];
    std::snprintf(buf, sizeof(buf), type == 'E' ? "%#14.6G" : "%#23.15G", v);
    return std::strtod(buf, NULL);
}

// pixel window (no beam) up to lMax for this nSide: registered table, else HEALPix file, else error
void pixelWindow(long nSide, int lMax, bool polarization, std::vector<double>& w)
{
    std::map<long, Window>::const_iterator it = g_windows.find(nSide);
    if(it != g_windows.end())
    {
        const std::vector<double>& src = polarization ? it->second.p : it->second.t;
        if(static_cast<int>(src.size()) < lMax + 1)
        {
            std::stringstream s;
            s << "The pixel window registered for nSide = " << nSide << " contains values only up to l = " << static_cast<long>(src.size()) - 1
              << ". Cannot read up to lMax = " << lMax << ".";
            raise(s.str());
        }
        w.assign(src.begin(), src.begin() + lMax + 1);
        return;
    }
    std::string dir = g_healpixDir;
    if(dir.empty())
    {
        const char* env = std::getenv("HEALPIX_DATA_DIR");
        if(env) dir = env;
    }
    if(dir.empty())
        raise("No pixel window available: call CMatrixGenerator::setPixelWindow or setHealpixDataDir (or set HEALPIX_DATA_DIR)");
    char name[64];
    std::snprintf(name, sizeof(name), "/pixel_window_n%04ld.fits", nSide);
    const std::string file = dir + name;
    Utils::FitsTable t;
    Utils::readFitsTable(file.c_str(), t);
    if(t.columns.size() < 2)
    {
        std::stringstream s;
        s << "Invalid format of the pixel windows function file " << file << ". Number of columns is " << t.columns.size() << ", needs to be at least 2.";
        raise(s.str());
    }
    const size_t col = polarization ? 1 : 0;
    if(static_cast<int>(t.columns[col].size()) < lMax + 1)
    {
        std::stringstream s;
        s << "The pixel windows function file " << file << " contains values only up to l = " << static_cast<long>(t.columns[col].size()) - 1
          << ". Cannot read up to lMax = " << lMax << ".";
        raise(s.str());
    }
    w.resize(static_cast<size_t>(lMax + 1));
    for(int l = 0; l <= lMax; ++l)
        w[l] = throughText(t.columns[col][l], t.columnType[col]);
}
}

void Utils::readPixelWindowFunction(std::vector<double>& f, long nSide, int lMax, double fwhm, bool polarization)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    std::vector<double> w;
    pixelWindow(nSide, lMax, polarization, w);
    f.resize(static_cast<size_t>(lMax + 1));
    if(cmg_window_beam(&f[0], lMax, fwhm, &w[0]) != CMG_OK)
        raise("invalid fwhm");
}

// ---------------------------------------------------------------- CMatrixGenerator

void CMatrixGenerator::setHealpixDataDir(const char* dir)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    g_healpixDir = dir ? dir : "";
}

void CMatrixGenerator::setPixelWindow(long nSide, const std::vector<double>& temperature, const std::vector<double>& polarization)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    Window w;
    w.t = temperature;
    w.p = polarization;
    g_windows[nSide] = w;
}

void cmgDropinPixelWindow(long nSide, int lMax, bool polarization, std::vector<double>& w)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    pixelWindow(nSide, lMax, polarization, w);
}

void CMatrixGenerator::clearPixelWindow(long nSide)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    g_windows.erase(nSide);
}

void CMatrixGenerator::setDevice(int device) { t_device = device; }

void CMatrixGenerator::transferCounters(long long& hostToDevice, long long& deviceToHost)
{
    DropinLock lock(cmgDropinCurrentDevice());
    std::int64_t a = 0, b = 0;
    cmg_transfer_counters(lock.ctx(), &a, &b);
    hostToDevice = a;
    deviceToHost = b;
}

void CMatrixGenerator::setDeviceResident(bool on)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    g_deviceResident = on;
}

void CMatrixGenerator::setPolarizationUsesTemperatureWindow(bool on)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    g_polarizationUsesTemperatureWindow = on;
}

namespace
{
// the result of a generator: on the device (the default), or in host memory as the reference's objects are
CMatrix* newResult(int nPix, int device, double** dPacked, long fullSkyNSide, int strips)
{
    bool resident;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        resident = g_deviceResident;
    }
    if(resident)
        return CMatrix::newOnDevice(nPix, device, dPacked, fullSkyNSide, strips);
    *dPacked = NULL;
    return new CMatrix(nPix);
}
}

CMatrix* CMatrixGenerator::clToCMatrix(const std::vector<double>& cl, long nSide, double fwhm, const std::vector<int>* goodPixels,
                                       const LegendrePolynomialContainer* lp)
{
    if(cl.empty())
        raise("CHECK FAILED");                      // check(!cl.empty()) of reference source/c_matrix_generator.cpp:167
    const int lMax = static_cast<int>(cl.size()) - 1;
    // P_l is recomputed on the GPU whether or not a container is given (recomputing beats reading 1.8 - 58 GB of cached
    // values); a container that does not describe this call is still an error, as it would be in the reference (:198-204)
    if(lp && (lp->lMax() < lMax || lp->nPix() != static_cast<int>(goodPixels ? goodPixels->size() : cmg_nside2npix(nSide))))
        raise("CHECK FAILED");
    std::vector<double> w;
    cmgDropinPixelWindow(nSide, lMax, false, w);
    const int device = cmgDropinCurrentDevice();
    DropinLock lock(device);
    cmg_ctx* ctx = lock.ctx();
    setPixels(ctx, nSide, goodPixels);
    double* dOut = NULL;
    CMatrix* m = newResult(static_cast<int>(cmg_npix(ctx)), device, &dOut, goodPixels ? 0 : nSide, 1);
    const cmg_status s = dOut ? cmg_cl_to_cmatrix_dev(ctx, &cl[0], lMax, fwhm, &w[0], dOut) : cmg_cl_to_cmatrix(ctx, &cl[0], lMax, fwhm, &w[0], m->packed());
    if(s != CMG_OK || cmg_synchronize(ctx) != CMG_OK)
    {
        delete m;
        check(ctx, s == CMG_OK ? CMG_ECUDA : s);
    }
    return m;
}

CMatrix* CMatrixGenerator::clToCMatrix(const char* clFileName, long nSide, int, double fwhm, const std::vector<int>* goodPixels,
                                       const LegendrePolynomialContainer* lp)
{
    // the reference ignores its lMax argument here too (source/c_matrix_generator.cpp:234-240)
    std::vector<double> cl;
    Utils::readClFromFile(clFileName, cl);
    return clToCMatrix(cl, nSide, fwhm, goodPixels, lp);
}

CMatrix* CMatrixGenerator::getFiducialMatrix(const std::vector<double>& cl, long nSide, int lMax, double fwhm, const std::vector<int>* goodPixels,
                                             const LegendrePolynomialContainer* lp)
{
    const int lMaxMax = static_cast<int>(4 * nSide);
    if(static_cast<int>(cl.size()) < lMaxMax + 1)
        raise("CHECK FAILED");                      // check(cl.size() >= lMaxMax + 1), :709
    if(lp && (lp->lMax() < lMaxMax || lp->nPix() != static_cast<int>(goodPixels ? goodPixels->size() : cmg_nside2npix(nSide))))
        raise("CHECK FAILED");                      // :726-733
    std::vector<double> w;
    cmgDropinPixelWindow(nSide, lMaxMax, false, w);
    const int device = cmgDropinCurrentDevice();
    DropinLock lock(device);
    cmg_ctx* ctx = lock.ctx();
    setPixels(ctx, nSide, goodPixels);
    double* dOut = NULL;
    CMatrix* m = newResult(static_cast<int>(cmg_npix(ctx)), device, &dOut, goodPixels ? 0 : nSide, 1);
    m->comment() = "fiducial matrix";
    const cmg_status s = dOut ? cmg_fiducial_matrix_dev(ctx, &cl[0], lMax, fwhm, &w[0], dOut) : cmg_fiducial_matrix(ctx, &cl[0], lMax, fwhm, &w[0], m->packed());
    if(s != CMG_OK || cmg_synchronize(ctx) != CMG_OK)
    {
        delete m;
        check(ctx, s == CMG_OK ? CMG_ECUDA : s);
    }
    return m;
}

CMatrix* CMatrixGenerator::getFiducialMatrix(const char* clFileName, long nSide, int lMax, double fwhm, const std::vector<int>* goodPixels,
                                             const LegendrePolynomialContainer* lp)
{
    std::vector<double> cl;
    Utils::readClFromFile(clFileName, cl);
    return getFiducialMatrix(cl, nSide, lMax, fwhm, goodPixels, lp);
}

CMatrix* CMatrixGenerator::generateNoiseMatrix(long nSide, double noise)
{
    const std::int64_t nPix = cmg_nside2npix(nSide);
    CMatrix* m = new CMatrix(static_cast<int>(nPix));
    m->comment() = "noise matrix";
    for(int i = 0; i < nPix; ++i)
        m->element(i, i) = noise * noise;
    return m;
}

CMatrix* CMatrixGenerator::clToCMatrixPol(const std::vector<double>& clTT, const std::vector<double>& clTE, const std::vector<double>& clEE,
                                          const std::vector<double>& clBB, long nSide, double fwhm, const std::vector<int>* goodPixels)
{
    if(clTT.empty() || clTE.size() != clTT.size() || clEE.size() != clTT.size() || clBB.size() != clTT.size())
        raise("CMatrixGenerator::clToCMatrixPol: the four spectra must be non-empty and of equal length");
    const int lMax = static_cast<int>(clTT.size()) - 1;
    // HEALPix's temperature window for T, its polarization window for Q and U; the reference's own polarization routine takes
    // the temperature table for both (source/c_matrix_generator.cpp:534) -- setPolarizationUsesTemperatureWindow(true) does that
    bool sameWindow;
    {
        std::lock_guard<std::mutex> lock(g_mutex);
        sameWindow = g_polarizationUsesTemperatureWindow;
    }
    std::vector<double> wT, wP;
    cmgDropinPixelWindow(nSide, lMax, false, wT);
    cmgDropinPixelWindow(nSide, lMax, !sameWindow, wP);
    const int device = cmgDropinCurrentDevice();
    DropinLock lock(device);
    cmg_ctx* ctx = lock.ctx();
    setPixels(ctx, nSide, goodPixels);
    const std::int64_t dim = 3 * cmg_npix(ctx);
    if(dim > 2147483647)
        raise("CMatrixGenerator::clToCMatrixPol: dimension exceeds the int interface of CMatrix");
    double* dOut = NULL;
    CMatrix* m = newResult(static_cast<int>(dim), device, &dOut, goodPixels ? 0 : nSide, 3);
    m->comment() = "TQU covariance matrix";
    const cmg_status s = dOut ? cmg_cl_to_cmatrix_pol_dev(ctx, &clTT[0], &clTE[0], &clEE[0], &clBB[0], lMax, fwhm, &wT[0], &wP[0], dOut)
                              : cmg_cl_to_cmatrix_pol(ctx, &clTT[0], &clTE[0], &clEE[0], &clBB[0], lMax, fwhm, &wT[0], &wP[0], m->packed());
    if(s != CMG_OK || cmg_synchronize(ctx) != CMG_OK)
    {
        delete m;
        check(ctx, s == CMG_OK ? CMG_ECUDA : s);
    }
    return m;
}

namespace
{
[[noreturn]] void notOnThisPath(const char* what)
{
    raise(std::string("CMatrixGenerator::") + what + " is a spherical-harmonic-transform route of the reference (HEALPix C++); "
          "it is outside the C_l -> pixel covariance hot path this library implements");
}
}

void CMatrixGenerator::clToWholeMatrix(const std::vector<double>&, WholeMatrix&) { notOnThisPath("clToWholeMatrix"); }
void CMatrixGenerator::clToWholeMatrix(const char*, WholeMatrix&, WholeMatrix&, WholeMatrix&) { notOnThisPath("clToWholeMatrix"); }
CMatrix* CMatrixGenerator::wholeMatrixToCMatrix(const WholeMatrix&, long, double, double, double, double) { notOnThisPath("wholeMatrixToCMatrix"); }
CMatrix* CMatrixGenerator::polarizationEEWholeMatrixToCMatrix(const WholeMatrix&, long, double, double, double, double)
{
    notOnThisPath("polarizationEEWholeMatrixToCMatrix");
}
CMatrix* CMatrixGenerator::calculateNoiseMatrix(const char*, const char*, double, double, long, double) { notOnThisPath("calculateNoiseMatrix"); }

// ---------------------------------------------------------------- LegendrePolynomialContainer

LegendrePolynomialContainer::LegendrePolynomialContainer(int lMax, long nSide, const std::vector<int>* goodPixels) : lMax_(lMax), nPix_(0)
{
    if(lMax < 0)
        raise("CHECK FAILED");
    const std::int64_t full = cmg_nside2npix(nSide);
    nPix_ = static_cast<int>(goodPixels ? goodPixels->size() : full);
    xyz_.resize(static_cast<size_t>(3) * nPix_);
    for(int k = 0; k < nPix_; ++k)
    {
        double theta, phi;
        if(cmg_pix2ang_nest(nSide, goodPixels ? (*goodPixels)[k] : k, &theta, &phi) != CMG_OK)
            raise("LegendrePolynomialContainer: invalid nSide or pixel index");
        xyz_[3 * k + 0] = std::sin(theta) * std::cos(phi);
        xyz_[3 * k + 1] = std::sin(theta) * std::sin(phi);
        xyz_[3 * k + 2] = std::cos(theta);
    }
}

namespace
{
double legendreAt(int l, double x)
{
    if(l == 0) return 1.0;
    double pm2 = 1.0, pm1 = x;
    for(int k = 2; k <= l; ++k)
    {
        const double p = (2 - 1.0 / k) * x * pm1 - (1 - 1.0 / k) * pm2;     // recurrence of reference include/legendre.hpp:33-34
        pm2 = pm1;
        pm1 = p;
    }
    return pm1;
}
}

double LegendrePolynomialContainer::value(int l, int j, int i) const
{
    if(l < 0 || l > lMax_ || j < 0 || j >= nPix_ || i < 0 || i > j)
        raise("CHECK FAILED");
    if(!file_.empty())
        return file_[l][static_cast<size_t>(cmg_packed_index(i, j))];
    double dot = xyz_[3 * i] * xyz_[3 * j] + xyz_[3 * i + 1] * xyz_[3 * j + 1] + xyz_[3 * i + 2] * xyz_[3 * j + 2];
    if(dot > 1) dot = 1;
    if(dot < -1) dot = -1;
    return legendreAt(l, dot);
}

// int32 lMax | int32 nPix | for l, for j: (j+1) doubles   (reference source/c_matrix_generator.cpp:136-162)
void LegendrePolynomialContainer::writeIntoFile(const char* fileName) const
{
    std::ofstream out(fileName, std::ios::binary | std::ios::out);
    if(!out)
        raise(std::string("Cannot write into file ") + fileName + ".");
    const std::int32_t lMax = lMax_, nPix = nPix_;
    out.write(reinterpret_cast<const char*>(&lMax), sizeof(lMax));
    out.write(reinterpret_cast<const char*>(&nPix), sizeof(nPix));
    std::vector<double> row;
    for(int l = 0; l <= lMax_; ++l)
        for(int j = 0; j < nPix_; ++j)
        {
            row.resize(static_cast<size_t>(j + 1));
            for(int i = 0; i <= j; ++i)
                row[i] = value(l, j, i);
            out.write(reinterpret_cast<const char*>(&row[0]), static_cast<std::streamsize>(sizeof(double) * row.size()));
        }
}

LegendrePolynomialContainer::LegendrePolynomialContainer(const char* fileName) : lMax_(0), nPix_(0)
{
    std::ifstream in(fileName, std::ios::in | std::ios::binary);
    if(!in)
        raise(std::string("Cannot open input file ") + fileName + ".");
    std::int32_t lMax = -1, nPix = -1;
    in.read(reinterpret_cast<char*>(&lMax), sizeof(lMax));
    if(!in || lMax < 0)
    {
        std::stringstream s;
        s << "Invalid lMax = " << lMax << " read from file " << fileName << ".";
        raise(s.str());
    }
    in.read(reinterpret_cast<char*>(&nPix), sizeof(nPix));
    if(!in || nPix < 0)
    {
        std::stringstream s;
        s << "Invalid nPix = " << nPix << " read from file " << fileName << ".";
        raise(s.str());
    }
    lMax_ = lMax;
    nPix_ = nPix;
    const size_t tri = static_cast<size_t>(cmg_packed_size(nPix));
    file_.resize(static_cast<size_t>(lMax + 1));
    for(int l = 0; l <= lMax; ++l)
    {
        file_[l].resize(tri);
        in.read(reinterpret_cast<char*>(file_[l].data()), static_cast<std::streamsize>(tri * sizeof(double)));
        if(!in)
            raise(std::string("The file ") + fileName + " is truncated.");
    }
}
