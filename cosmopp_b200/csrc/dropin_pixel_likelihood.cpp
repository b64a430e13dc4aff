// PixelLikelihoodTT of include/pixel_likelihood.hpp: parameters -> C_l -> device C-matrix -> device likelihood.
#include <string>
#include <vector>

#include <cuda_runtime_api.h>

#include <cmg.h>
#include <exception_handler.hpp>
#include <pixel_likelihood.hpp>

#include "dropin_internal.hpp"

namespace
{
[[noreturn]] void raise(const std::string& text) { throw StandardException(text); }

void check(cmg_ctx* ctx, cmg_status s)
{
    if(s != CMG_OK)
        raise(std::string("PixelLikelihoodTT: ") + cmg_last_error(ctx));
}

double* upload(const CMatrix& m)
{
    double* p = NULL;
    const size_t bytes = sizeof(double) * static_cast<size_t>(m.packedSize());
    if(cudaMalloc(reinterpret_cast<void**>(&p), bytes) != cudaSuccess || cudaMemcpy(p, m.packed(), bytes, cudaMemcpyHostToDevice) != cudaSuccess)
        raise(std::string("PixelLikelihoodTT: cannot place the matrices on the GPU: ") + cudaGetErrorString(cudaGetLastError()));
    return p;
}
}

PixelLikelihoodTT::PixelLikelihoodTT(long nSide, int lMax, double fwhm, const std::vector<int>& goodPixels, const CMatrix& fiducialMatrix,
                                     const CMatrix& noiseMatrix, const std::vector<double>& map, const std::vector<double>& foreground, ClModel& model)
    : nSide_(nSide), lMax_(lMax), goodPixels_(goodPixels), map_(map), foreground_(foreground), model_(model), dFiducial_(NULL), dNoise_(NULL),
      dC_(NULL), cCapacity_(0), chi2_(0), logDet_(0)
{
    const size_t n = goodPixels_.size();
    if(n == 0 || lMax_ < 2)
        raise("PixelLikelihoodTT: empty pixel list or lMax < 2");
    if(static_cast<size_t>(fiducialMatrix.getNPix()) != n || static_cast<size_t>(noiseMatrix.getNPix()) != n || map_.size() != n ||
       (!foreground_.empty() && foreground_.size() != n))
        raise("PixelLikelihoodTT: fiducial matrix, noise matrix, map and foreground must all refer to the unmasked pixels");
    std::vector<double> w;
    cmgDropinPixelWindow(nSide_, lMax_, false, w);
    windowBeam_.resize(lMax_ + 1);
    cmg_window_beam(&windowBeam_[0], lMax_, fwhm, &w[0]);
    dFiducial_ = upload(fiducialMatrix);
    dNoise_ = upload(noiseMatrix);
}

PixelLikelihoodTT::~PixelLikelihoodTT()
{
    if(dFiducial_) cudaFree(dFiducial_);
    if(dNoise_) cudaFree(dNoise_);
    if(dC_) cudaFree(dC_);
}

double PixelLikelihoodTT::calculate(double* params, int nParams)
{
    double like = 0;
    calculateBatch(params, nParams, 1, &like);
    return like;
}

void PixelLikelihoodTT::calculateBatch(const double* params, int nParams, int nSets, double* like)
{
    if(nSets < 1 || !like)
        raise("PixelLikelihoodTT: nothing to calculate");
    cmg_ctx* ctx = cmgDropinContext();
    const std::int64_t n = static_cast<std::int64_t>(goodPixels_.size());
    const std::int64_t packed = cmg_packed_size(n);
    // the process-wide context may have been used for another mask meanwhile: (re)bind the pixel set (O(N) host work)
    check(ctx, cmg_set_pixels(ctx, nSide_, &goodPixels_[0], n));
    if(cCapacity_ < nSets)
    {
        if(dC_) cudaFree(dC_);
        dC_ = NULL;
        if(cudaMalloc(reinterpret_cast<void**>(&dC_), sizeof(double) * packed * nSets) != cudaSuccess)
            raise("PixelLikelihoodTT: out of device memory for the batch of matrices");
        cCapacity_ = nSets;
    }
    std::vector<double> cl(lMax_ + 1), a(static_cast<size_t>(nSets) * (lMax_ + 1));
    for(int s = 0; s < nSets; ++s)
    {
        cl.assign(lMax_ + 1, 0.0);
        model_.clTT(params + static_cast<size_t>(s) * nParams, nParams, cl);
        if(static_cast<int>(cl.size()) != lMax_ + 1)
            raise("PixelLikelihoodTT: the model must return lMax + 1 values");
        cmg_tt_weights(&cl[0], &windowBeam_[0], lMax_, &a[static_cast<size_t>(s) * (lMax_ + 1)]);
    }
    if(nSets == 1)
        check(ctx, cmg_legendre_series(ctx, &a[0], lMax_, 0, n, dC_));
    else
        check(ctx, cmg_legendre_series_batched(ctx, &a[0], lMax_, nSets, 0, n, dC_, packed));
    for(int s = 0; s < nSets; ++s)
    {
        cmg_like* lk = NULL;
        const cmg_status st = cmg_like_create(ctx, dC_ + s * packed, 1, dFiducial_, dNoise_, n, foreground_.empty() ? NULL : &foreground_[0], &lk);
        if(st == CMG_ENUMERIC)
            raise(cmg_last_error(ctx));
        check(ctx, st);
        const cmg_status sc = cmg_like_calculate(lk, &map_[0], 1, &chi2_, &logDet_);
        cmg_like_destroy(lk);
        check(ctx, sc);
        like[s] = chi2_ + logDet_;
    }
}
