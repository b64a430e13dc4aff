// Drop-in replacement of the temperature part of the reference's include/likelihood.hpp (struct LikelihoodResult :19-31,
// class Likelihood :37-126): the low-l pixel likelihood that consumes the matrices of c_matrix_generator.hpp.
// Same public members, argument meaning and error behaviour (StandardException with the reference's messages).
// Underneath (cmg_like_* of cmg.h): C + F + N is summed and unpacked on the GPU, factorised once (Cholesky), and
// chi2 = |L^-1 t|^2 is evaluated for all maps of a calculateAll call together; the reference inverts the matrix on
// the host (dpptrf / dpptri) and runs an O(n^2) double loop per map (source/likelihood.cpp:100-160).
// LikelihoodPolarization (:129-): the pixel-space half of the reference's class -- inverse-noise weighted [Q;U] likelihood,
// cInv = N^-1 + N^-1 C N^-1 factorised on the GPU -- is here; its harmonic-space half (the E-mode map predicted from the
// temperature a_lm through the ET(TT)^-1 WholeMatrix, rotate_alm and alm2map_pol: source/likelihood.cpp:540-590) needs the
// HEALPix C++ SHT machinery and is left to the caller, who passes the predicted map in.
#ifndef COSMO_PP_LIKELIHOOD_HPP
#define COSMO_PP_LIKELIHOOD_HPP

#include <string>
#include <vector>

#include <c_matrix.hpp>

struct LikelihoodResult
{
    std::string mapName;
    double logDet;
    double chi2;
    double like;       // -2 log(likelihood) = logDet + chi2

    inline bool operator<(const LikelihoodResult& other) const { return like < other.like; }
    LikelihoodResult& operator=(const LikelihoodResult& other)
    {
        mapName = other.mapName; logDet = other.logDet; chi2 = other.chi2; like = other.like;
        return *this;
    }
};

struct cmg_like;

class Likelihood
{
public:
    // mask file -> good pixels; foregroundFileName NULL = no foreground marginalisation
    Likelihood(const CMatrix& cMatrix, const CMatrix& fiducialMatrix, const CMatrix& noiseMatrix, const char* maskFileName, const char* foregroundFileName = NULL);
    // foreground empty = no foreground marginalisation
    Likelihood(const CMatrix& cMatrix, const CMatrix& fiducialMatrix, const CMatrix& noiseMatrix, const std::vector<int>& goodPixels, const std::vector<double>& foreground);
    ~Likelihood();

    double calculate(const char* mapName, const char* noiseMapName, double& chi2, double& logDet) const;
    void calculateAll(const char* inputListName, std::vector<LikelihoodResult>& results) const;
    void calculateAll(const std::vector<std::vector<double> >& t, const std::vector<std::string>& mapNames, std::vector<LikelihoodResult>& results) const;
    double calculate(const std::vector<double>& t, double& chi2, double& logDet) const;

    static void readInput(const char* inputListName, const std::vector<int>& goodPixels, std::vector<std::vector<double> >& t, std::vector<std::string>& mapNames);
    static void readMapAndNoise(const char* mapName, const char* noiseMapName, const std::vector<int>& goodPixels, long& nSide, std::vector<double>& t);
    static void readForeground(const char* foregroundFileName, const std::vector<int>& goodPixels, long& nSide, std::vector<double>& f);

private:
    Likelihood(const Likelihood&);                 // owns a device factorisation: not copyable
    Likelihood& operator=(const Likelihood&);
    void construct(const CMatrix& cMatrix, const CMatrix& fiducialMatrix, const CMatrix& noiseMatrix, const std::vector<int>& goodPixels, const std::vector<double>& foreground);

    cmg_like* like_;
    int device_;                                   // the GPU the factorisation lives on (CMatrixGenerator::setDevice of the constructing thread)
    std::vector<int> goodPixels_;
};

// Pixel-space part of the reference's LikelihoodPolarization (include/likelihood.hpp:137-, source/likelihood.cpp:341-406 and
// 536-612).  cMatrix is a [Q;U] covariance over ALL pixels of the map (dimension 2 nPix: the layout of the reference's
// polarizationEEWholeMatrixToCMatrix, and of polarizationBlock() below applied to this library's [T;Q;U] matrix); the inverse
// noise matrix has the same layout (the reference reads it from the text file n_inv.txt: size x size numbers).  Both are
// restricted to the unmasked pixels [Q(good); U(good)], cInv = N^-1 + N^-1 C N^-1 is formed and factorised on the GPU,
// logDet = log det cInv - 16078.083180 (the reference's offset, :397).
class LikelihoodPolarization
{
public:
    LikelihoodPolarization(const CMatrix& cMatrix, long nSide, const std::vector<int>& goodPixels, const CMatrix& nInv);
    // inverse noise matrix from a text file of size x size numbers, row by row (the reference reads "n_inv.txt" from the
    // working directory)
    LikelihoodPolarization(const CMatrix& cMatrix, long nSide, const std::vector<int>& goodPixels, const char* nInvFileName = "n_inv.txt");
    ~LikelihoodPolarization();

    // v = N^-1 (Q, U) on the unmasked pixels (2 nGood values: Q then U), as in the reference (:560-564).  prediction: the
    // (Q, U) map predicted from temperature on the same pixels, subtracted as v - N^-1 prediction (:592-600); empty = none.
    // chi2 = v^T cInv^-1 v; returns chi2 + logDet.
    double calculate(const std::vector<double>& v, const std::vector<double>& prediction, double& chi2, double& logDet) const;

    // the [Q;U] block (dimension 2 nPix) of a [T;Q;U] matrix (dimension 3 nPix) of this library's clToCMatrixPol
    static CMatrix* polarizationBlock(const CMatrix& tqu);

private:
    LikelihoodPolarization(const LikelihoodPolarization&);
    LikelihoodPolarization& operator=(const LikelihoodPolarization&);
    void construct(const CMatrix& cMatrix, long nSide, const CMatrix& nInv);

    cmg_like* like_;
    int device_;
    std::vector<int> goodPixels_;
    std::vector<double> nInvGood_;                 // N^-1 over the unmasked pixels, dense 2g x 2g row-major (for v - N^-1 prediction)
};

#endif
