// PixelLikelihoodTT of include/pixel_likelihood.hpp: parameters -> C_l -> device C-matrix -> device likelihood.
#include <string>
#include <vector>

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

}

PixelLikelihoodTT::PixelLikelihoodTT(long nSide, int lMax, double fwhm, const std::vector<int>& goodPixels, const CMatrix& fiducialMatrix,
                                     const CMatrix& noiseMatrix, const std::vector<double>& map, const std::vector<double>& foreground, ClModel& model)
    : nSide_(nSide), lMax_(lMax), goodPixels_(goodPixels), map_(map), foreground_(foreground), model_(model), fiducial_(fiducialMatrix),
      noise_(noiseMatrix), device_(cmgDropinCurrentDevice()), dC_(NULL), cCapacity_(0), chi2_(0), logDet_(0)
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
    if(cmg_window_beam(&windowBeam_[0], lMax_, fwhm, &w[0]) != CMG_OK)
        raise("PixelLikelihoodTT: fwhm must be >= 0");
    fiducial_.devicePacked(device_);
    noise_.devicePacked(device_);
}

PixelLikelihoodTT::~PixelLikelihoodTT()
{
    if(dC_)
    {
        DropinLock lock(device_);
        cmg_synchronize(lock.ctx());
        cmg_device_free(lock.ctx(), dC_);
    }
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
    // the GPU's context is shared with everything else this process runs on that GPU: the whole evaluation holds its lock
    DropinLock lock(device_);
    cmg_ctx* ctx = lock.ctx();
    const std::int64_t n = static_cast<std::int64_t>(goodPixels_.size());
    const std::int64_t packed = cmg_packed_size(n);
    // the context may have been used for another mask meanwhile: (re)bind the pixel set (O(N) host work)
    check(ctx, cmg_set_pixels(ctx, nSide_, &goodPixels_[0], n));
    const double* dFiducial = fiducial_.devicePacked(device_);
    const double* dNoise = noise_.devicePacked(device_);
    if(cCapacity_ < nSets)
    {
        if(dC_)
        {
            cmg_synchronize(ctx);
            cmg_device_free(ctx, dC_);
        }
        dC_ = NULL;
        cCapacity_ = 0;
        void* p = NULL;
        if(cmg_device_malloc(ctx, static_cast<std::int64_t>(sizeof(double)) * packed * nSets, &p) != CMG_OK)
            raise("PixelLikelihoodTT: out of device memory for the batch of matrices");
        dC_ = static_cast<double*>(p);
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
        const cmg_status st = cmg_like_create(ctx, dC_ + s * packed, 1, dFiducial, dNoise, n, foreground_.empty() ? NULL : &foreground_[0], &lk);
        if(st == CMG_ENUMERIC)
            raise(cmg_last_error(ctx));
        check(ctx, st);
        const cmg_status sc = cmg_like_calculate(lk, &map_[0], 1, &chi2_, &logDet_);
        cmg_like_destroy(lk);
        check(ctx, sc);
        like[s] = chi2_ + logDet_;
    }
}
