// Likelihood of include/likelihood.hpp: host wrapper over cmg_like_* (cmg.h).  Error texts follow the reference
// (source/likelihood.cpp:53-131, 186-277).
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <cmg.h>
#include <exception_handler.hpp>
#include <likelihood.hpp>
#include <utils.hpp>

#include "dropin_internal.hpp"

namespace
{
[[noreturn]] void raise(const std::string& text) { throw StandardException(text); }

// HEALPix map: first column of the first binary table, NESTED required; `what` names the map in the error text
void readNestedMap(const char* fileName, const char* what, long& nSide, std::vector<double>& map)
{
    Utils::FitsTable table;
    Utils::readFitsTable(fileName, table);
    if(table.ordering != "NESTED" && table.ordering != "NEST")
        raise(std::string("The ") + what + " must have nested ordering.");
    if(table.columns.empty())
        raise(std::string("No table column in ") + fileName + ".");
    map.swap(table.columns[0]);
    nSide = table.nSide;
    if(nSide <= 0)
    {
        long n = 1;
        while(12 * n * n < static_cast<long>(map.size())) n *= 2;
        nSide = n;
    }
    if(static_cast<long>(map.size()) != 12 * nSide * nSide)
        raise(std::string("The number of pixels in ") + fileName + " does not match its NSide.");
}

void pick(const std::vector<double>& map, const std::vector<int>& goodPixels, const char* fileName, std::vector<double>& out)
{
    out.resize(goodPixels.size());
    for(size_t i = 0; i < goodPixels.size(); ++i)
    {
        const int index = goodPixels[i];
        if(index < 0 || static_cast<size_t>(index) >= map.size())
            raise(std::string("Unmasked pixel index outside the map ") + fileName + ".");
        out[i] = map[index];
    }
}

}

Likelihood::Likelihood(const CMatrix& cMatrix, const CMatrix& fiducialMatrix, const CMatrix& noiseMatrix, const char* maskFileName, const char* foregroundFileName)
    : like_(NULL), device_(0)
{
    long nSideMask = 0, nSideFore = 0;
    std::vector<int> goodPixels;
    std::vector<double> f;
    Utils::readMask(maskFileName, nSideMask, goodPixels);
    if(foregroundFileName != NULL)
    {
        readForeground(foregroundFileName, goodPixels, nSideFore, f);
        if(nSideMask != nSideFore)
        {
            std::stringstream exceptionStr;
            exceptionStr << "Mask file " << maskFileName << " has nSide = " << nSideMask << " while the foreground file " << foregroundFileName
                         << " has nSide = " << nSideFore << ". They need to be the same.";
            raise(exceptionStr.str());
        }
    }
    construct(cMatrix, fiducialMatrix, noiseMatrix, goodPixels, f);
}

Likelihood::Likelihood(const CMatrix& cMatrix, const CMatrix& fiducialMatrix, const CMatrix& noiseMatrix, const std::vector<int>& goodPixels, const std::vector<double>& foreground)
    : like_(NULL), device_(0)
{
    construct(cMatrix, fiducialMatrix, noiseMatrix, goodPixels, foreground);
}

Likelihood::~Likelihood()
{
    if(like_)
    {
        DropinLock lock(device_);
        cmg_like_destroy(like_);
    }
}

void Likelihood::construct(const CMatrix& cMatrix, const CMatrix& fiducialMatrix, const CMatrix& noiseMatrix, const std::vector<int>& goodPixels, const std::vector<double>& foreground)
{
    goodPixels_ = goodPixels;
    const struct { const CMatrix* m; const char* what; const char* file; } parts[3] = {
        {&cMatrix, "covariance matrix", "c.dat"}, {&fiducialMatrix, "fiducial covariance matrix", "c_fiducial.dat"}, {&noiseMatrix, "noise covariance matrix", "c_noise.dat"}};
    for(int k = 0; k < 3; ++k)
        if(static_cast<size_t>(parts[k].m->getNPix()) != goodPixels_.size())
        {
            std::stringstream exceptionStr;
            exceptionStr << "There are " << goodPixels_.size() << " unmasked pixels, however the " << parts[k].what << " in " << parts[k].file
                         << " corresponds to " << parts[k].m->getNPix() << ". Please generate the " << parts[k].what << " with the same mask.";
            raise(exceptionStr.str());
        }
    if(!foreground.empty() && foreground.size() != goodPixels_.size())
        raise("The foreground map must have one value per unmasked pixel.");

    // the three matrices where they are: a matrix a generator produced is already on the GPU (CMatrix is a handle), a host-side
    // one (generateNoiseMatrix, a file) is uploaded once; nothing comes back to the host.  The factorisation owns its own
    // packed buffer (cmg_like_create), the inputs stay untouched.
    device_ = cmgDropinCurrentDevice();
    DropinLock lock(device_);
    cmg_ctx* ctx = lock.ctx();
    const double* c = cMatrix.devicePacked(device_);
    const double* f = fiducialMatrix.devicePacked(device_);
    const double* n = noiseMatrix.devicePacked(device_);
    const cmg_status s = cmg_like_create(ctx, c, 1, f, n, static_cast<std::int64_t>(goodPixels_.size()),
                                         foreground.empty() ? NULL : &foreground[0], &like_);
    if(s == CMG_ENUMERIC)
        raise(cmg_last_error(ctx));           // the reference's text: "The determinant of the covariance matrix is not positive. ..."
    if(s != CMG_OK)
        raise(std::string("Likelihood: ") + cmg_last_error(ctx));
}

double Likelihood::calculate(const std::vector<double>& t, double& chi2, double& logDet) const
{
    if(t.size() != goodPixels_.size())
        raise("CHECK FAILED");                // check() of the reference (source/likelihood.cpp:165)
    DropinLock lock(device_);
    const cmg_status s = cmg_like_calculate(like_, &t[0], 1, &chi2, &logDet);
    if(s != CMG_OK)
        raise(std::string("Likelihood: ") + cmg_last_error(lock.ctx()));
    return chi2 + logDet;
}

double Likelihood::calculate(const char* mapName, const char* noiseMapName, double& chi2, double& logDet) const
{
    std::vector<double> t;
    long nSide;
    readMapAndNoise(mapName, noiseMapName, goodPixels_, nSide, t);
    return calculate(t, chi2, logDet);
}

void Likelihood::calculateAll(const std::vector<std::vector<double> >& t, const std::vector<std::string>& mapNames, std::vector<LikelihoodResult>& results) const
{
    const size_t numOfMaps = t.size();
    if(mapNames.size() != numOfMaps)
        raise("CHECK FAILED");
    if(numOfMaps == 0)
        return;
    const size_t n = goodPixels_.size();
    std::vector<double> flat(numOfMaps * n), chi2(numOfMaps);
    for(size_t k = 0; k < numOfMaps; ++k)
    {
        if(t[k].size() != n)
            raise("CHECK FAILED");
        std::copy(t[k].begin(), t[k].end(), flat.begin() + k * n);
    }
    double logDet = 0;
    DropinLock lock(device_);
    const cmg_status s = cmg_like_calculate(like_, &flat[0], static_cast<std::int64_t>(numOfMaps), &chi2[0], &logDet);
    if(s != CMG_OK)
        raise(std::string("Likelihood: ") + cmg_last_error(lock.ctx()));
    LikelihoodResult res;
    for(size_t k = 0; k < numOfMaps; ++k)
    {
        res.mapName = mapNames[k];
        res.logDet = logDet;
        res.chi2 = chi2[k];
        res.like = res.logDet + res.chi2;
        results.push_back(res);
    }
}

void Likelihood::calculateAll(const char* inputListName, std::vector<LikelihoodResult>& results) const
{
    std::vector<std::string> mapNames;
    std::vector<std::vector<double> > t;
    readInput(inputListName, goodPixels_, t, mapNames);
    calculateAll(t, mapNames, results);
}

void Likelihood::readMapAndNoise(const char* mapName, const char* noiseMapName, const std::vector<int>& goodPixels, long& nSide, std::vector<double>& t)
{
    std::vector<double> map, noise;
    long nSideNoise = 0;
    readNestedMap(mapName, "map", nSide, map);
    readNestedMap(noiseMapName, "noise map", nSideNoise, noise);
    if(nSide != nSideNoise)
    {
        std::stringstream exceptionStr;
        exceptionStr << "Map and noise must have the same NSide, for the map it is " << nSide << " and for the noise it is " << nSideNoise << ".";
        raise(exceptionStr.str());
    }
    std::vector<double> tn;
    pick(map, goodPixels, mapName, t);
    pick(noise, goodPixels, noiseMapName, tn);
    for(size_t i = 0; i < t.size(); ++i)
        t[i] += tn[i];
}

void Likelihood::readForeground(const char* foregroundFileName, const std::vector<int>& goodPixels, long& nSide, std::vector<double>& f)
{
    std::vector<double> fore;
    readNestedMap(foregroundFileName, "foreground map", nSide, fore);
    pick(fore, goodPixels, foregroundFileName, f);
}

void Likelihood::readInput(const char* inputListName, const std::vector<int>& goodPixels, std::vector<std::vector<double> >& t, std::vector<std::string>& mapNames)
{
    std::ifstream inList(inputListName);
    if(!inList)
    {
        std::stringstream exceptionStr;
        exceptionStr << "Cannot read the map list file " << inputListName << ".";
        raise(exceptionStr.str());
    }
    int numOfMaps = 0;
    inList >> numOfMaps;
    if(numOfMaps < 0)
    {
        std::stringstream exceptionStr;
        exceptionStr << "The number of maps cannot be negative. It is " << numOfMaps << ".";
        raise(exceptionStr.str());
    }
    t.resize(numOfMaps);
    mapNames.resize(numOfMaps);
    long nSide;
    for(int i = 0; i < numOfMaps; ++i)
    {
        std::string mapName, noiseName;
        inList >> mapName >> noiseName;
        mapNames[i] = mapName;
        readMapAndNoise(mapName.c_str(), noiseName.c_str(), goodPixels, nSide, t[i]);
    }
}

// ---------------------------------------------------------------- LikelihoodPolarization (pixel-space part)

namespace
{
// [Q(good); U(good)] restriction of a matrix over [Q(all); U(all)], as reference source/likelihood.cpp:375-389
CMatrix* restrictQU(const CMatrix& full, const std::vector<int>& goodPixels)
{
    const int half = full.getNPix() / 2;
    std::vector<int> index(2 * goodPixels.size());
    for(size_t i = 0; i < goodPixels.size(); ++i)
    {
        if(goodPixels[i] < 0 || goodPixels[i] >= half)
            raise("LikelihoodPolarization: unmasked pixel index outside the map");
        index[i] = goodPixels[i];
        index[goodPixels.size() + i] = half + goodPixels[i];
    }
    CMatrix* reduced = new CMatrix(full);          // a device-resident matrix is copied and gathered on the device
    reduced->maskMatrix(index);
    return reduced;
}
}

LikelihoodPolarization::LikelihoodPolarization(const CMatrix& cMatrix, long nSide, const std::vector<int>& goodPixels, const CMatrix& nInv)
    : like_(NULL), device_(0), goodPixels_(goodPixels)
{
    construct(cMatrix, nSide, nInv);
}

LikelihoodPolarization::LikelihoodPolarization(const CMatrix& cMatrix, long nSide, const std::vector<int>& goodPixels, const char* nInvFileName)
    : like_(NULL), device_(0), goodPixels_(goodPixels)
{
    const int size = cMatrix.getNPix();
    std::ifstream in(nInvFileName);
    if(!in)
        raise(std::string("Cannot read the input file ") + nInvFileName + ".");      // the reference's text (source/likelihood.cpp:357)
    CMatrix nInv(size);
    for(int i = 0; i < size; ++i)
        for(int j = 0; j < size; ++j)
        {
            double v = 0;
            in >> v;
            if(!in)
                raise(std::string("The file ") + nInvFileName + " does not hold size x size numbers.");
            if(i <= j)
                nInv.element(i, j) = v;
        }
    construct(cMatrix, nSide, nInv);
}

void LikelihoodPolarization::construct(const CMatrix& cMatrix, long nSide, const CMatrix& nInv)
{
    const int size = cMatrix.getNPix(), goodSize = static_cast<int>(goodPixels_.size());
    if(size % 2 != 0)
        raise("polarization c matrix includes q and u parts, so size must be even");       // check() of the reference, :346
    if(size / 2 < goodSize || goodSize == 0)
        raise("CHECK FAILED");
    if(nSide > 0 && static_cast<std::int64_t>(size / 2) != cmg_nside2npix(nSide))
        raise("LikelihoodPolarization: the covariance matrix does not cover the pixels of this nSide");
    if(nInv.getNPix() != size)
        raise("LikelihoodPolarization: the inverse noise matrix must have the dimension of the covariance matrix");
    CMatrix* c = restrictQU(cMatrix, goodPixels_);
    CMatrix* ni = NULL;
    try
    {
        ni = restrictQU(nInv, goodPixels_);
        const int m = 2 * goodSize;
        nInvGood_.resize(static_cast<size_t>(m) * m);
        for(int i = 0; i < m; ++i)
            for(int j = 0; j < m; ++j)
                nInvGood_[static_cast<size_t>(i) * m + j] = ni->element(i, j);
        device_ = cmgDropinCurrentDevice();
        DropinLock lock(device_);
        const double detOffset = 16078.083180;       // reference source/likelihood.cpp:397
        const cmg_status s = cmg_like_create_ninv(lock.ctx(), c->devicePacked(device_), ni->devicePacked(device_), m, detOffset, &like_);
        if(s == CMG_ENUMERIC)
            raise(cmg_last_error(lock.ctx()));
        if(s != CMG_OK)
            raise(std::string("LikelihoodPolarization: ") + cmg_last_error(lock.ctx()));
    }
    catch(...)
    {
        delete c;
        delete ni;
        throw;
    }
    delete c;
    delete ni;
}

LikelihoodPolarization::~LikelihoodPolarization()
{
    if(like_)
    {
        DropinLock lock(device_);
        cmg_like_destroy(like_);
    }
}

double LikelihoodPolarization::calculate(const std::vector<double>& v, const std::vector<double>& prediction, double& chi2, double& logDet) const
{
    const size_t m = 2 * goodPixels_.size();
    if(v.size() != m || (!prediction.empty() && prediction.size() != m))
        raise("CHECK FAILED");                       // check(vCopy.size() == 2 * goodPixels_.size()), :548
    std::vector<double> vCopy(v);
    if(!prediction.empty())
        for(size_t i = 0; i < m; ++i)                // vCopy[i] -= nInv_(i, j) * subtractMap[j], :592-596
        {
            double s = 0;
            for(size_t j = 0; j < m; ++j)
                s += nInvGood_[i * m + j] * prediction[j];
            vCopy[i] -= s;
        }
    DropinLock lock(device_);
    const cmg_status s = cmg_like_calculate(like_, &vCopy[0], 1, &chi2, &logDet);
    if(s != CMG_OK)
        raise(std::string("LikelihoodPolarization: ") + cmg_last_error(lock.ctx()));
    return chi2 + logDet;
}

CMatrix* LikelihoodPolarization::polarizationBlock(const CMatrix& tqu)
{
    const int dim = tqu.getNPix();
    if(dim % 3 != 0)
        raise("LikelihoodPolarization::polarizationBlock: the matrix is not a [T;Q;U] matrix");
    const int n = dim / 3;
    std::vector<int> index(2 * n);
    for(int i = 0; i < 2 * n; ++i)
        index[i] = n + i;
    CMatrix* block = new CMatrix(tqu);
    block->maskMatrix(index);
    block->comment() = "QU covariance matrix";
    return block;
}
