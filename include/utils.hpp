// File-format side of the C-matrix path, same names as the reference's Utils (include/utils.hpp:11-62) for the
// members the path uses: mask -> good pixel list, Gaussian beam, pixel window x beam, C_l text files.
// (Utils::maskRegion(s) of the reference operate on HEALPix C++ maps and are not part of this library.)
#ifndef COSMO_PP_B200_UTILS_HPP
#define COSMO_PP_B200_UTILS_HPP

#include <string>
#include <vector>

class Utils
{
public:
    // HEALPix FITS map (binary table, NESTED) -> nSide and ascending indices of pixels with value > 0.5
    // (reference source/utils.cpp:25-52); throws StandardException if the ordering is not NESTED
    static void readMask(const char* maskFileName, long& nSide, std::vector<int>& goodPixels);

    // exp(-l(l+1) / (2 sigma^2)), sigma = sqrt(8 ln 2) / (fwhm pi / 180); 1 for fwhm == 0 (source/utils.cpp:54-64)
    static double beamFunction(int l, double fwhm);

    // f[l] = pixel window (temperature or polarization column) x beamFunction(l, fwhm), l = 0..lMax
    // (reference source/utils.cpp:66-170)
    static void readPixelWindowFunction(std::vector<double>& f, long nSide, int lMax, double fwhm = 0, bool polarization = false);

    // one C_l per line, optionally preceded by l, optionally D_l = l(l+1) C_l / 2 pi (source/utils.cpp:172-218)
    static void readClFromFile(const char* fileName, std::vector<double>& cl, bool hasL = false, bool isDl = false);

    // first column of the first binary-table extension of a FITS file, flattened, plus the header keywords the
    // path needs; enough of FITS for HEALPix masks and pixel-window tables (no cfitsio dependency)
    struct FitsTable
    {
        std::vector<std::vector<double> > columns;
        std::vector<char> columnType;      // 'E', 'D', ...
        std::string ordering;              // ORDERING keyword, upper case
        long nSide;                        // NSIDE keyword or 0
    };
    static void readFitsTable(const char* fileName, FitsTable& table);
};

#endif
