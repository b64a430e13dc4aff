// Drop-in replacement of the temperature part of the reference's include/likelihood.hpp (struct LikelihoodResult :19-31,
// class Likelihood :37-126): the low-l pixel likelihood that consumes the matrices of c_matrix_generator.hpp.
// Same public members, argument meaning and error behaviour (StandardException with the reference's messages).
// Underneath (cmg_like_* of cmg.h): C + F + N is summed and unpacked on the GPU, factorised once (Cholesky), and
// chi2 = |L^-1 t|^2 is evaluated for all maps of a calculateAll call together; the reference inverts the matrix on
// the host (dpptrf / dpptri) and runs an O(n^2) double loop per map (source/likelihood.cpp:100-160).
// LikelihoodPolarization (:129-) needs the reference's harmonic-space WholeMatrix / Alm machinery and is not part of
// this library.
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
    std::vector<int> goodPixels_;
};

#endif
