// C ABI of the generator (include/cmg.h): context, geometry, launches.  No CPU fallback anywhere:
// every compute entry point needs a live sm_100 context and fails with CMG_ECUDA otherwise.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include "../../include/cmg.h"
#include "healpix_nest.hpp"
#include "host_expand.hpp"
#include "cholesky.cuh"
#include "kernels.cuh"
#include "orbit.cuh"
#include "series.hpp"

namespace
{
const double kPi = 3.141592653589793;      // Math::pi of the reference (include/math_constants.hpp:8)
const int kTabLen = CMG_LMAX_LIMIT + 2;
thread_local std::string g_createError;
}

struct cmg_ctx
{
    int device = 0;
    cudaStream_t ownStream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;

    int64_t nside = 0;
    int64_t npix = 0;
    bool fullSky = false;            // all 12 nside^2 pixels in NESTED order (what the symmetry-orbit path needs)
    double* dGeo = nullptr;          // [8][npix]
    double* dTables = nullptr;       // 7 tables of kTabLen
    double* dWeights = nullptr;      // staging for host-supplied weights
    int64_t weightsCap = 0;          // doubles
    double* dScratch = nullptr;      // whole-call output buffer
    int64_t scratchCap = 0;          // bytes
    int32_t* dIndex = nullptr;       // maskMatrix indices
    int64_t indexCap = 0;

    cmg::SeriesTable hostT0, hostT20, hostT22;   // host copies of the recurrence tables
    int tquVariant = 0;                          // 0 = automatic choice (see launchTqu)
    int hostExpandDirectMask = 0;                // images that cross PCIe next to the last-face columns (bit 3 strip + k - 1)
    int hostExpandThreads = -1;                  // full-sky whole calls copy back 27 % and expand on the host: > 0 threads, 0 = plain copy,
                                                 // -1 = automatic (all host cores for matrices of 1 GiB and more; measured 1.27x at 87 GB)

    static const int kAux = 4;                   // side streams for many small independent launches (batched mode)
    cudaStream_t aux[kAux] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t auxDone[kAux] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t forkEv = nullptr;

    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timing = false;
    double lastMs = 0.0;
    int64_t launches = 0;
    int64_t bytesH2D = 0, bytesD2H = 0;          // what crossed PCIe through this context (cmg_transfer_counters)
    long long* dCholInfo = nullptr;              // status word of cmg_packed_cholesky
    double* dCholRed = nullptr;                  // log det / reductions of the packed solves
    double* dCholPanel = nullptr;                // dense panel planes of cmg_packed_cholesky (cholesky.cuh, cholSyrkKernel)
    size_t cholPanelDoubles = 0;
    int cholGroup = 0;                           // blocks of 128 rows per trailing update (cmg_set_cholesky_group); 0 = by size
    int cholLookAhead = 1;                       // the next group is factorised beside the trailing update (cmg_set_cholesky_lookahead)
    cudaStream_t cholSide = nullptr;             // high-priority stream of the look-ahead
    cudaEvent_t cholEvA = nullptr, cholEvF = nullptr;
    int likeMethod = 0;                          // cmg_like_create: 0 = this library's packed factorisation, 1 = cuSOLVER on the unpacked matrix
};

namespace
{

cmg_status fail(cmg_ctx* ctx, cmg_status s, const std::string& msg)
{
    if(ctx)
        ctx->err = msg;
    else
        g_createError = msg;
    return s;
}

cmg_status cudaFail(cmg_ctx* ctx, cudaError_t e, const char* what)
{
    return fail(ctx, e == cudaErrorMemoryAllocation ? CMG_ENOMEM : CMG_ECUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
}

#define CMG_CUDA(ctx, call)                                      \
    do                                                           \
    {                                                            \
        const cudaError_t e_ = (call);                           \
        if(e_ != cudaSuccess)                                    \
            return cudaFail((ctx), e_, #call);                   \
    } while(0)

cmg::Geometry geometryOf(const cmg_ctx* ctx)
{
    cmg::Geometry g;
    const double* b = ctx->dGeo;
    const int64_t n = ctx->npix;
    g.nx = b; g.ny = b + n; g.nz = b + 2 * n;
    g.tx = b + 3 * n; g.ty = b + 4 * n; g.tz = b + 5 * n;
    g.px = b + 6 * n; g.py = b + 7 * n;
    g.npix = n;
    return g;
}

cmg::DeviceTables tablesOf(const cmg_ctx* ctx)
{
    cmg::DeviceTables t;
    const double* b = ctx->dTables;
    t.N0 = b; t.g0 = b + kTabLen;
    t.N20 = b + 2 * kTabLen; t.g20 = b + 3 * kTabLen;
    t.N22 = b + 4 * kTabLen; t.g22 = b + 5 * kTabLen; t.c22 = b + 6 * kTabLen;
    return t;
}

cmg_status ensureWeights(cmg_ctx* ctx, int64_t doubles)
{
    if(doubles <= ctx->weightsCap)
        return CMG_OK;
    if(ctx->dWeights)
    {
        CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CMG_CUDA(ctx, cudaFree(ctx->dWeights));
        ctx->dWeights = nullptr;
        ctx->weightsCap = 0;
    }
    CMG_CUDA(ctx, cudaMalloc(&ctx->dWeights, sizeof(double) * doubles));
    ctx->weightsCap = doubles;
    return CMG_OK;
}

cmg_status ensureScratch(cmg_ctx* ctx, int64_t bytes)
{
    if(bytes <= ctx->scratchCap)
        return CMG_OK;
    if(ctx->dScratch)
    {
        CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        CMG_CUDA(ctx, cudaFree(ctx->dScratch));
        ctx->dScratch = nullptr;
        ctx->scratchCap = 0;
    }
    CMG_CUDA(ctx, cudaMalloc(&ctx->dScratch, bytes));
    ctx->scratchCap = bytes;
    return CMG_OK;
}

cmg_status checkReady(cmg_ctx* ctx, int lmax)
{
    if(!ctx)
        return CMG_EINVAL;
    if(ctx->npix <= 0)
        return fail(ctx, CMG_ESTATE, "cmg_set_pixels has not been called on this context");
    if(lmax < 0)
        return fail(ctx, CMG_EINVAL, "lmax must be >= 0");
    if(lmax > CMG_LMAX_LIMIT)
        return fail(ctx, CMG_EUNSUPPORTED, "lmax exceeds CMG_LMAX_LIMIT");
    return CMG_OK;
}

struct KernelTimer
{
    cmg_ctx* ctx;
    explicit KernelTimer(cmg_ctx* c) : ctx(c)
    {
        if(ctx->timing)
            cudaEventRecord(ctx->ev0, ctx->stream);
    }
    cmg_status finish()
    {
        if(!ctx->timing)
            return CMG_OK;
        CMG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CMG_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        CMG_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->lastMs = ms;
        return CMG_OK;
    }
};

// hostWeights != NULL (single matrix, weights known on the host) enables the static-table kernel
cmg_status launchLegendre(cmg_ctx* ctx, const double* dA, int64_t aStride, int lmax, int64_t nBatch,
                          int64_t colBegin, int64_t colEnd, double* dOut, int64_t outStride, const double* hostWeights,
                          cudaStream_t stream = nullptr)
{
    const bool sideStream = stream != nullptr;       // one of many launches of a batch: the caller times and counts
    if(!stream) stream = ctx->stream;
    if(colBegin < 0 || colEnd > ctx->npix || colBegin > colEnd)
        return fail(ctx, CMG_EINVAL, "column range outside [0, npix]");
    if(nBatch < 1 || nBatch > 65535)
        return fail(ctx, CMG_EINVAL, "n_batch must be in [1, 65535] per call");
    if(colBegin == colEnd)
        return CMG_OK;
    if(!dOut || !dA)
        return fail(ctx, CMG_EINVAL, "null device pointer");
    const int64_t rowBlocks = (colEnd + cmg::TT_ROWS - 1) / cmg::TT_ROWS;
    const int64_t colBlocks = (colEnd - colBegin + cmg::TT_COLS - 1) / cmg::TT_COLS;
    if(colBlocks > 65535)
        return fail(ctx, CMG_EUNSUPPORTED, "too many column blocks for one launch");
    const dim3 grid(static_cast<unsigned>(rowBlocks), static_cast<unsigned>(colBlocks), static_cast<unsigned>(nBatch));
    const cmg::DeviceTables t = tablesOf(ctx);
    const bool useStatic = hostWeights && nBatch == 1 && lmax + 1 <= cmg::TT_STATIC_STEPS && ctx->tquVariant != 1;
    static thread_local cmg::TtStaticTable T;
    int entrySlot = 0;
    if(useStatic)
    {
        for(int i = 0; i < cmg::TT_STATIC_STEPS; ++i)
        {
            const int k = cmg::TT_STATIC_STEPS - 1 - i;
            T.s[i] = make_double2(k <= lmax ? hostWeights[k] * ctx->hostT0.N[k] : 0.0, -ctx->hostT0.g[k + 1]);
        }
        entrySlot = cmg::TT_STATIC_STEPS - 1 - lmax;      // first slot with a non-zero weight (k = lmax)
    }
    cudaEvent_t timed0 = nullptr;
    if(ctx->timing && !sideStream) { timed0 = ctx->ev0; cudaEventRecord(ctx->ev0, ctx->stream); }
    // TT variants: 0 = R 8 columns per thread and pass, 4 CTAs/SM (with the rolled chunk loop the kernel is insensitive to
    // occupancy: 87.3-88.1 % of the FP64 peak for every (R, CTAs/SM) tried at Nside=64, profiles/r1_kernel_history.md);
    // 1 = shared-memory table; 248 and 2216 stay selectable for the variant test.
#define CMG_TT(R_, M_) cmg::legendreSeriesKernel<true, R_, M_><<<grid, cmg::TT_ROWS, 0, stream>>>(T, geometryOf(ctx), dA, aStride, t.N0, t.g0, lmax, entrySlot, colBegin, colEnd, dOut, outStride)
    if(useStatic)
    {
        switch(ctx->tquVariant)
        {
            case 2216: CMG_TT(2, 16); break;
            case 248: CMG_TT(4, 8); break;
            default: CMG_TT(8, 4); break;
        }
    }
#undef CMG_TT
    else
        cmg::legendreSeriesKernel<false, cmg::TT_R, 8><<<grid, cmg::TT_ROWS, sizeof(double2) * (lmax + 1), stream>>>(
            T, geometryOf(ctx), dA, aStride, t.N0, t.g0, lmax, entrySlot, colBegin, colEnd, dOut, outStride);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    if(timed0)
    {
        CMG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CMG_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        CMG_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        ctx->lastMs = ms;
    }
    return CMG_OK;
}

size_t tquSmemBytes(int lmax, bool isStatic)
{
    return (isStatic ? 0 : sizeof(double4) * 2 * (lmax + 1)) +
           sizeof(double) * cmg::PQ_SMEM_DOUBLES;
}

// coefficient table of the static-table kernel (kernels.cuh, TquStaticTable) from host weights
void fillStaticTable(const cmg_ctx* ctx, const double* att, const double* ate, const double* aee, const double* abb, int lmax,
                     cmg::TquStaticTable& T)
{
    const cmg::SeriesTable& t0 = ctx->hostT0;
    const cmg::SeriesTable& t20 = ctx->hostT20;
    const cmg::SeriesTable& t22 = ctx->hostT22;
    for(int i = 0; i < cmg::PQ_STATIC_STEPS; ++i)
    {
        const int k = cmg::PQ_STATIC_STEPS + 1 - i;
        double4 A = make_double4(0.0, 0.0, 0.0, 0.0);
        if(k <= lmax)
        {
            A.x = att[k] * t0.N[k];
            A.y = ate[k] * t20.N[k] * 0.61237243569579452455;
            A.z = (aee[k] + abb[k]) * t22.N[k] * 0.125;
            A.w = (aee[k] - abb[k]) * t22.N[k] * 0.125;
        }
        T.s[2 * i] = A;
        T.s[2 * i + 1] = make_double4(-t0.g[k + 1], -t20.g[k + 1], -t22.g[k + 1], t22.c[k]);
    }
    const double a1 = lmax >= 1 ? att[1] * t0.N[1] : 0.0;
    T.s[2 * cmg::PQ_STATIC_STEPS] = make_double4(a1, -t0.g[2], att[0] * t0.N[0], -t0.g[1]);
}

template <int R, bool STATIC, int MINB>
cmg_status launchTquVariant(cmg_ctx* ctx, const cmg::TquDynamicArgs& dyn, const cmg::TquStaticTable& T, int entrySlot,
                            const cmg::PartTable& P, dim3 grid, int64_t outStride, cudaStream_t stream = nullptr)
{
    if(!stream) stream = ctx->stream;
    const size_t smem = tquSmemBytes(dyn.lmax, STATIC);
    auto kernel = cmg::tquKernel<R, STATIC, MINB>;
    CMG_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    kernel<<<grid, cmg::PQ_THREADS, smem, stream>>>(T, geometryOf(ctx), dyn, entrySlot, P, outStride);
    CMG_CUDA(ctx, cudaGetLastError());
    return CMG_OK;
}

// hostWeights: the four host weight arrays when the caller has them (enables the static-table kernel)
cmg_status launchTqu(cmg_ctx* ctx, const double* dA, int64_t aStride, int lmax, int64_t nBatch,
                     const cmg_tqu_layout* layout, int64_t outStride, const double* const* hostWeights)
{
    if(!layout || layout->n_parts < 1 || layout->n_parts > CMG_MAX_PARTS || layout->own < 0 || layout->own >= layout->n_parts)
        return fail(ctx, CMG_EINVAL, "bad layout: n_parts / own");
    if(layout->begin[0] != 0 || layout->begin[layout->n_parts] != ctx->npix)
        return fail(ctx, CMG_EINVAL, "layout must cover pixel columns [0, npix]");
    cmg::PartTable P;
    std::memset(&P, 0, sizeof(P));
    P.n = layout->n_parts;
    P.own = layout->own;
    for(int k = 0; k <= P.n; ++k)
    {
        P.begin[k] = layout->begin[k];
        if(k > 0 && P.begin[k] < P.begin[k - 1])
            return fail(ctx, CMG_EINVAL, "layout boundaries must be non-decreasing");
    }
    for(int k = 0; k < P.n; ++k)
    {
        for(int s = 0; s < 3; ++s)
            P.ptr[k][s] = layout->ptr[k][s];
        P.kind[k] = layout->kind[k];
        P.ld[k] = layout->ld[k];
        P.row0[k] = layout->row0[k];
        if(P.kind[k] != 0 && P.kind[k] != 1)
            return fail(ctx, CMG_EINVAL, "layout kind must be 0 (packed) or 1 (dense blocks)");
        if(k < P.own && P.begin[k + 1] > P.begin[k] && (!P.ptr[k][1] || !P.ptr[k][2] || (P.kind[k] == 1 && !P.ptr[k][0])))
            return fail(ctx, CMG_EINVAL, "a part left of own has no storage for the transposed entries");
    }
    if(P.kind[P.own] != 0 || !P.ptr[P.own][0] || !P.ptr[P.own][1] || !P.ptr[P.own][2])
        return fail(ctx, CMG_EINVAL, "own part must be packed (kind 0) with three strips");
    const int64_t colBegin = P.begin[P.own], colEnd = P.begin[P.own + 1];
    if(colBegin == colEnd)
        return CMG_OK;
    if(nBatch < 1 || nBatch > 65535)
        return fail(ctx, CMG_EINVAL, "n_batch must be in [1, 65535] per call");
    const int64_t rowBlocks = (colEnd + cmg::PQ_TI - 1) / cmg::PQ_TI;
    const int64_t colBlocks = (colEnd - colBegin + cmg::PQ_TJ - 1) / cmg::PQ_TJ;
    if(colBlocks > 65535)
        return fail(ctx, CMG_EUNSUPPORTED, "too many column blocks for one launch");
    const dim3 grid(static_cast<unsigned>(rowBlocks), static_cast<unsigned>(colBlocks), static_cast<unsigned>(nBatch));

    cmg::TquDynamicArgs dyn;
    dyn.a = dA;
    dyn.aStride = aStride;
    dyn.tab = tablesOf(ctx);
    dyn.lmax = lmax;

    // variant code: 0 = automatic, else 100*static + 10*R + minBlocksPerSM, e.g. 123 = static table, R=2, 3 CTAs/SM;
    // 22 = shared-memory table, R=2, 2 CTAs/SM
    const bool canStatic = hostWeights && nBatch == 1 && lmax >= 2 && lmax <= cmg::PQ_STATIC_LMAX;
    int variant = ctx->tquVariant;
    if(variant == 0 || variant >= 900)
        variant = canStatic ? 142 : 42;
    if(variant >= 100 && !canStatic)
        return fail(ctx, CMG_EINVAL, "static-table kernel needs host weights, one batch element and 2 <= lmax <= PQ_STATIC_LMAX");

    static thread_local cmg::TquStaticTable T;      // 28 KB: keep it off the stack
    int entrySlot = 0;
    if(variant >= 100)
    {
        fillStaticTable(ctx, hostWeights[0], hostWeights[1], hostWeights[2], hostWeights[3], lmax, T);
        entrySlot = cmg::PQ_STATIC_STEPS + 1 - lmax;      // first slot with a non-zero weight (k = lmax)
    }
    KernelTimer timer(ctx);
    cmg_status s = CMG_OK;
    switch(variant)
    {
#define CMG_V(code, R, ST, MB) case code: s = launchTquVariant<R, ST, MB>(ctx, dyn, T, entrySlot, P, grid, outStride); break;
    CMG_V(22, 2, false, 2) CMG_V(42, 4, false, 2) CMG_V(81, 8, false, 1)
    CMG_V(114, 1, true, 4) CMG_V(122, 2, true, 2) CMG_V(123, 2, true, 3) CMG_V(124, 2, true, 4) CMG_V(142, 4, true, 2)
#undef CMG_V
    default: return fail(ctx, CMG_EINVAL, "unknown kernel variant");
    }
    if(s != CMG_OK) return s;
    ctx->launches += 1;
    return timer.finish();
}

} // namespace

extern "C"
{

int cmg_version(void) { return 100; }

int cmg_device_count(void)
{
    int n = 0;
    if(cudaGetDeviceCount(&n) != cudaSuccess)
    {
        cudaGetLastError();
        return 0;
    }
    return n;
}

cmg_status cmg_create(cmg_ctx** out, int device)
{
    if(!out)
        return fail(nullptr, CMG_EINVAL, "ctx output pointer is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if(e != cudaSuccess || n == 0)
    {
        cudaGetLastError();
        return fail(nullptr, CMG_ECUDA, std::string("no CUDA device available (this library has no CPU fallback): ") +
                                            (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    }
    if(device < 0 || device >= n)
        return fail(nullptr, CMG_EINVAL, "device index out of range");
    cudaDeviceProp prop;
    if((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return cudaFail(nullptr, e, "cudaGetDeviceProperties");
    if(prop.major != 10)
        return fail(nullptr, CMG_ECUDA, std::string("device ") + prop.name + " is not sm_100 (kernels are built for sm_100a only)");

    cmg_ctx* ctx = new(std::nothrow) cmg_ctx;
    if(!ctx)
        return fail(nullptr, CMG_ENOMEM, "out of host memory");
    ctx->device = device;
    auto bail = [&](cudaError_t err, const char* what) {
        const cmg_status s = cudaFail(nullptr, err, what);
        cmg_destroy(ctx);
        return s;
    };
    if((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if((e = cudaStreamCreateWithFlags(&ctx->ownStream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    ctx->stream = ctx->ownStream;
    if((e = cudaEventCreate(&ctx->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if((e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
    if((e = cudaEventCreateWithFlags(&ctx->forkEv, cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    for(int k = 0; k < cmg_ctx::kAux; ++k)
    {
        if((e = cudaStreamCreateWithFlags(&ctx->aux[k], cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
        if((e = cudaEventCreateWithFlags(&ctx->auxDone[k], cudaEventDisableTiming)) != cudaSuccess) return bail(e, "cudaEventCreate");
    }

    // recurrence tables, once per context
    std::vector<double> host(7 * kTabLen, 0.0);
    ctx->hostT0 = cmg::makeSeriesTable(CMG_LMAX_LIMIT, 0, 0);
    ctx->hostT20 = cmg::makeSeriesTable(CMG_LMAX_LIMIT, 2, 0);
    ctx->hostT22 = cmg::makeSeriesTable(CMG_LMAX_LIMIT, 2, 2);
    const cmg::SeriesTable& t0 = ctx->hostT0;
    const cmg::SeriesTable& t20 = ctx->hostT20;
    const cmg::SeriesTable& t22 = ctx->hostT22;
    for(int l = 0; l < kTabLen; ++l)
    {
        host[0 * kTabLen + l] = t0.N[l];
        host[1 * kTabLen + l] = t0.g[l];
        host[2 * kTabLen + l] = t20.N[l];
        host[3 * kTabLen + l] = t20.g[l];
        host[4 * kTabLen + l] = t22.N[l];
        host[5 * kTabLen + l] = t22.g[l];
        host[6 * kTabLen + l] = t22.c[l];
    }
    if((e = cudaMalloc(&ctx->dTables, sizeof(double) * host.size())) != cudaSuccess) return bail(e, "cudaMalloc(tables)");
    if((e = cudaMemcpy(ctx->dTables, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
        return bail(e, "cudaMemcpy(tables)");
    *out = ctx;
    return CMG_OK;
}

void cmg_destroy(cmg_ctx* ctx)
{
    if(!ctx)
        return;
    cudaSetDevice(ctx->device);
    if(ctx->ownStream) cudaStreamSynchronize(ctx->ownStream);
    if(ctx->dGeo) cudaFree(ctx->dGeo);
    if(ctx->dTables) cudaFree(ctx->dTables);
    if(ctx->dWeights) cudaFree(ctx->dWeights);
    if(ctx->dScratch) cudaFree(ctx->dScratch);
    if(ctx->dIndex) cudaFree(ctx->dIndex);
    if(ctx->dCholInfo) cudaFree(ctx->dCholInfo);
    if(ctx->dCholRed) cudaFree(ctx->dCholRed);
    if(ctx->dCholPanel) cudaFree(ctx->dCholPanel);
    if(ctx->cholSide) cudaStreamDestroy(ctx->cholSide);
    if(ctx->cholEvA) cudaEventDestroy(ctx->cholEvA);
    if(ctx->cholEvF) cudaEventDestroy(ctx->cholEvF);
    for(int k = 0; k < cmg_ctx::kAux; ++k)
    {
        if(ctx->aux[k]) { cudaStreamSynchronize(ctx->aux[k]); cudaStreamDestroy(ctx->aux[k]); }
        if(ctx->auxDone[k]) cudaEventDestroy(ctx->auxDone[k]);
    }
    if(ctx->forkEv) cudaEventDestroy(ctx->forkEv);
    if(ctx->ev0) cudaEventDestroy(ctx->ev0);
    if(ctx->ev1) cudaEventDestroy(ctx->ev1);
    if(ctx->ownStream) cudaStreamDestroy(ctx->ownStream);
    delete ctx;
}

const char* cmg_last_error(const cmg_ctx* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

cmg_status cmg_set_stream(cmg_ctx* ctx, void* s)
{
    if(!ctx) return CMG_EINVAL;
    ctx->stream = static_cast<cudaStream_t>(s);
    return CMG_OK;
}

cmg_status cmg_use_own_stream(cmg_ctx* ctx)
{
    if(!ctx) return CMG_EINVAL;
    ctx->stream = ctx->ownStream;
    return CMG_OK;
}

cmg_status cmg_synchronize(cmg_ctx* ctx)
{
    if(!ctx) return CMG_EINVAL;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CMG_OK;
}

int64_t cmg_launch_count(const cmg_ctx* ctx) { return ctx ? ctx->launches : 0; }

cmg_status cmg_device_malloc(cmg_ctx* ctx, int64_t bytes, void** p)
{
    if(!ctx || !p || bytes < 0) return CMG_EINVAL;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    CMG_CUDA(ctx, cudaMalloc(p, static_cast<size_t>(bytes)));
    return CMG_OK;
}

cmg_status cmg_device_free(cmg_ctx* ctx, void* p)
{
    if(!ctx) return CMG_EINVAL;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    CMG_CUDA(ctx, cudaFree(p));
    return CMG_OK;
}

cmg_status cmg_ipc_export(cmg_ctx* ctx, void* dPtr, void* handle)
{
    if(!ctx || !dPtr || !handle) return CMG_EINVAL;
    static_assert(sizeof(cudaIpcMemHandle_t) == CMG_IPC_HANDLE_BYTES, "IPC handle size");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CMG_CUDA(ctx, cudaIpcGetMemHandle(&h, dPtr));
    std::memcpy(handle, &h, sizeof(h));
    return CMG_OK;
}

cmg_status cmg_ipc_open(cmg_ctx* ctx, const void* handle, void** dPeer)
{
    if(!ctx || !handle || !dPeer) return CMG_EINVAL;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof(h));
    CMG_CUDA(ctx, cudaIpcOpenMemHandle(dPeer, h, cudaIpcMemLazyEnablePeerAccess));
    return CMG_OK;
}

cmg_status cmg_ipc_close(cmg_ctx* ctx, void* dPeer)
{
    if(!ctx || !dPeer) return CMG_EINVAL;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    CMG_CUDA(ctx, cudaIpcCloseMemHandle(dPeer));
    return CMG_OK;
}

cmg_status cmg_host_malloc_pinned(int64_t bytes, void** p)
{
    if(!p || bytes < 0) return CMG_EINVAL;
    const cudaError_t e = cudaHostAlloc(p, static_cast<size_t>(bytes), cudaHostAllocPortable);
    if(e != cudaSuccess)
        return cudaFail(nullptr, e, "cudaHostAlloc");
    return CMG_OK;
}

cmg_status cmg_host_free_pinned(void* p)
{
    const cudaError_t e = cudaFreeHost(p);
    if(e != cudaSuccess)
        return cudaFail(nullptr, e, "cudaFreeHost");
    return CMG_OK;
}

cmg_status cmg_copy_to_host(cmg_ctx* ctx, void* dst, const void* src, int64_t bytes)
{
    if(!ctx || !dst || !src || bytes < 0) return CMG_EINVAL;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    CMG_CUDA(ctx, cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDeviceToHost, ctx->stream));
    CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->bytesD2H += bytes;
    return CMG_OK;
}

cmg_status cmg_copy_to_device(cmg_ctx* ctx, void* dst, const void* src, int64_t bytes)
{
    if(!ctx || !dst || !src || bytes < 0) return CMG_EINVAL;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    CMG_CUDA(ctx, cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyHostToDevice, ctx->stream));
    ctx->bytesH2D += bytes;
    return CMG_OK;
}

cmg_status cmg_copy_on_device(cmg_ctx* ctx, void* dst, const void* src, int64_t bytes)
{
    if(!ctx || !dst || !src || bytes < 0) return CMG_EINVAL;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    CMG_CUDA(ctx, cudaMemcpyAsync(dst, src, static_cast<size_t>(bytes), cudaMemcpyDeviceToDevice, ctx->stream));
    return CMG_OK;
}

cmg_status cmg_transfer_counters(const cmg_ctx* ctx, int64_t* h2d, int64_t* d2h)
{
    if(!ctx) return CMG_EINVAL;
    if(h2d) *h2d = ctx->bytesH2D;
    if(d2h) *d2h = ctx->bytesD2H;
    return CMG_OK;
}

// ---------------------------------------------------------------- host pieces

int64_t cmg_nside2npix(int64_t nside) { return cmg::nside2npix(nside); }

cmg_status cmg_pix2ang_nest(int64_t nside, int64_t ipix, double* theta, double* phi)
{
    if(!theta || !phi || !cmg::validNside(nside) || ipix < 0 || ipix >= cmg::nside2npix(nside))
        return CMG_EINVAL;
    cmg::pix2angNest(nside, ipix, *theta, *phi);
    return CMG_OK;
}

int64_t cmg_packed_size(int64_t dim) { return dim * (dim + 1) / 2; }

int64_t cmg_packed_index(int64_t i, int64_t j)
{
    if(i > j)
        std::swap(i, j);
    return j * (j + 1) / 2 + i;
}

cmg_status cmg_good_pixels_from_mask(const double* mask, int64_t npix, int32_t* good, int64_t* nGood)
{
    if(!mask || !good || !nGood || npix < 0)
        return CMG_EINVAL;
    int64_t n = 0;
    for(int64_t i = 0; i < npix; ++i)
        if(mask[i] > 0.5)
            good[n++] = static_cast<int32_t>(i);
    *nGood = n;
    return CMG_OK;
}

double cmg_beam_function(int l, double fwhm)
{
    if(fwhm == 0)
        return 1.0;
    const double sigma = std::sqrt(8 * std::log(2.0)) / (fwhm * kPi / 180);
    const int ll1 = l * (l + 1);          // the reference forms l(l+1) in int and negates it (source/utils.cpp:63)
    return std::exp(-ll1 / (2 * sigma * sigma));
}

cmg_status cmg_window_beam(double* f, int lmax, double fwhm, const double* pixwin)
{
    if(!f || lmax < 0 || fwhm < 0)
        return CMG_EINVAL;
    for(int l = 0; l <= lmax; ++l)
    {
        double v = pixwin ? pixwin[l] : 1.0;
        v *= cmg_beam_function(l, fwhm);
        f[l] = v;
    }
    return CMG_OK;
}

cmg_status cmg_tt_weights(const double* cl, const double* f, int lmax, double* a)
{
    if(!cl || !f || !a || lmax < 0)
        return CMG_EINVAL;
    for(int l = 0; l <= lmax; ++l)
    {
        if(l < 2)
        {
            a[l] = 0.0;                                         // monopole and dipole excluded (:190,219)
            continue;
        }
        const double clCopy = cl[l] * (2 * l + 1) / (4 * kPi);  // reference c_matrix_generator.cpp:192
        a[l] = clCopy * f[l] * f[l];
    }
    return CMG_OK;
}

cmg_status cmg_fiducial_weights(const double* cl, const double* f, int64_t nside, int lmax, double* a)
{
    if(!cl || !f || !a || nside < 1 || lmax < 0)
        return CMG_EINVAL;
    const int lMaxMax = static_cast<int>(4 * nside);
    for(int l = 0; l <= lMaxMax; ++l)
        a[l] = (l > lmax) ? cl[l] * ((2 * l + 1) / (4 * kPi)) * f[l] * f[l] : 0.0;   // :758
    const double md = 100 * cl[2] * f[2] * f[2];                                       // :762, (1 + z) = P_0 + P_1
    a[0] += md;
    if(lMaxMax >= 1)
        a[1] += md;
    return CMG_OK;
}

cmg_status cmg_tqu_weights(const double* ctt, const double* cte, const double* cee, const double* cbb,
                           const double* fT, const double* fP, int lmax,
                           double* att, double* ate, double* aee, double* abb)
{
    if(!ctt || !cte || !cee || !cbb || !fT || !fP || !att || !ate || !aee || !abb || lmax < 0)
        return CMG_EINVAL;
    for(int l = 0; l <= lmax; ++l)
    {
        if(l < 2)
        {
            att[l] = ate[l] = aee[l] = abb[l] = 0.0;
            continue;
        }
        const double w = (2 * l + 1) / (4 * kPi);
        att[l] = ctt[l] * w * fT[l] * fT[l];
        ate[l] = cte[l] * w * fT[l] * fP[l];
        aee[l] = cee[l] * w * fP[l] * fP[l];
        abb[l] = cbb[l] * w * fP[l] * fP[l];
    }
    return CMG_OK;
}

cmg_status cmg_noise_matrix(int64_t npix, double noise, double* out)
{
    if(!out || npix < 1)
        return CMG_EINVAL;
    std::memset(out, 0, sizeof(double) * static_cast<size_t>(cmg_packed_size(npix)));
    for(int64_t i = 0; i < npix; ++i)
        out[cmg_packed_index(i, i)] = noise * noise;
    return CMG_OK;
}

// ---------------------------------------------------------------- geometry

cmg_status cmg_set_pixels(cmg_ctx* ctx, int64_t nside, const int32_t* good, int64_t nGood)
{
    if(!ctx)
        return CMG_EINVAL;
    if(!cmg::validNside(nside))
        return fail(ctx, CMG_EINVAL, "nside must be a power of two in [1, 8192]");
    const int64_t full = cmg::nside2npix(nside);
    const int64_t n = good ? nGood : full;
    if(n < 1)
        return fail(ctx, CMG_EINVAL, "empty pixel list");
    bool fullSky = n == full;
    std::vector<double> host(static_cast<size_t>(8 * n));
    for(int64_t k = 0; k < n; ++k)
    {
        const int64_t ipix = good ? good[k] : k;
        if(ipix < 0 || ipix >= full)
            return fail(ctx, CMG_EINVAL, "pixel index outside [0, 12 nside^2)");
        fullSky = fullSky && ipix == k;
        const cmg::PixelFrame f = cmg::pixelFrame(nside, ipix);
        host[0 * n + k] = f.n[0];
        host[1 * n + k] = f.n[1];
        host[2 * n + k] = f.n[2];
        host[3 * n + k] = f.eTheta[0];
        host[4 * n + k] = f.eTheta[1];
        host[5 * n + k] = f.eTheta[2];
        host[6 * n + k] = f.ePhi[0];
        host[7 * n + k] = f.ePhi[1];
    }
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if(ctx->dGeo)
    {
        CMG_CUDA(ctx, cudaFree(ctx->dGeo));
        ctx->dGeo = nullptr;
        ctx->npix = 0;
        ctx->fullSky = false;
    }
    CMG_CUDA(ctx, cudaMalloc(&ctx->dGeo, sizeof(double) * host.size()));
    CMG_CUDA(ctx, cudaMemcpy(ctx->dGeo, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice));
    ctx->nside = nside;
    ctx->npix = n;
    ctx->fullSky = fullSky;
    return CMG_OK;
}

int64_t cmg_npix(const cmg_ctx* ctx) { return ctx ? ctx->npix : 0; }

cmg_status cmg_get_geometry(cmg_ctx* ctx, double* out)
{
    if(!ctx || !out) return CMG_EINVAL;
    if(ctx->npix <= 0) return fail(ctx, CMG_ESTATE, "cmg_set_pixels has not been called on this context");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    CMG_CUDA(ctx, cudaMemcpy(out, ctx->dGeo, sizeof(double) * 8 * ctx->npix, cudaMemcpyDeviceToHost));
    return CMG_OK;
}

// ---------------------------------------------------------------- TT

cmg_status cmg_legendre_series_dev(cmg_ctx* ctx, const double* dA, int lmax, int64_t colBegin, int64_t colEnd, double* dOut)
{
    const cmg_status s = checkReady(ctx, lmax);
    if(s != CMG_OK) return s;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    return launchLegendre(ctx, dA, 0, lmax, 1, colBegin, colEnd, dOut, 0, nullptr);
}

cmg_status cmg_legendre_series(cmg_ctx* ctx, const double* a, int lmax, int64_t colBegin, int64_t colEnd, double* dOut)
{
    cmg_status s = checkReady(ctx, lmax);
    if(s != CMG_OK) return s;
    if(!a) return fail(ctx, CMG_EINVAL, "null weights");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    if((s = ensureWeights(ctx, lmax + 1)) != CMG_OK) return s;
    CMG_CUDA(ctx, cudaMemcpyAsync(ctx->dWeights, a, sizeof(double) * (lmax + 1), cudaMemcpyHostToDevice, ctx->stream));
    return launchLegendre(ctx, ctx->dWeights, 0, lmax, 1, colBegin, colEnd, dOut, 0, a);
}

cmg_status cmg_legendre_series_batched(cmg_ctx* ctx, const double* a, int lmax, int64_t nBatch,
                                       int64_t colBegin, int64_t colEnd, double* dOut, int64_t stride)
{
    cmg_status s = checkReady(ctx, lmax);
    if(s != CMG_OK) return s;
    if(!a || nBatch < 1) return fail(ctx, CMG_EINVAL, "bad batch arguments");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    if((s = ensureWeights(ctx, nBatch * (lmax + 1))) != CMG_OK) return s;
    CMG_CUDA(ctx, cudaMemcpyAsync(ctx->dWeights, a, sizeof(double) * nBatch * (lmax + 1), cudaMemcpyHostToDevice, ctx->stream));
    if(ctx->tquVariant == 0 && lmax + 1 <= cmg::TT_STATIC_STEPS)
    {
        // one static-table launch per element over the side streams (as cmg_tqu_batched does): the static kernel is 1.2-1.3x
        // faster than the shared-memory-table kernel and small grids overlap each other's heads and tails
        KernelTimer timer(ctx);
        CMG_CUDA(ctx, cudaEventRecord(ctx->forkEv, ctx->stream));
        for(int k = 0; k < cmg_ctx::kAux; ++k)
            CMG_CUDA(ctx, cudaStreamWaitEvent(ctx->aux[k], ctx->forkEv, 0));
        for(int64_t b = 0; b < nBatch; ++b)
            if((s = launchLegendre(ctx, ctx->dWeights, 0, lmax, 1, colBegin, colEnd, dOut + b * stride, 0, a + b * (lmax + 1),
                                   ctx->aux[b % cmg_ctx::kAux])) != CMG_OK) return s;
        for(int k = 0; k < cmg_ctx::kAux; ++k)
        {
            CMG_CUDA(ctx, cudaEventRecord(ctx->auxDone[k], ctx->aux[k]));
            CMG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->auxDone[k], 0));
        }
        return timer.finish();
    }
    return launchLegendre(ctx, ctx->dWeights, lmax + 1, lmax, nBatch, colBegin, colEnd, dOut, stride, nullptr);
}

// Full sky, host output, cmg_set_host_expand: only the columns of the last face of every ring cross PCIe (contiguous pieces of
// the packed triangle in ctx->dScratch, in chunks of ~64 MB, smallest columns first); worker threads fill in the rotated
// images of a chunk as soon as it has arrived, while the chunks behind it are in flight (host_expand.cpp).
static int hostExpandThreadsFor(const cmg_ctx* ctx, int64_t bytes)
{
    if(!ctx->fullSky || ctx->hostExpandThreads == 0) return 0;
    if(ctx->hostExpandThreads > 0) return ctx->hostExpandThreads;
    if(bytes < (int64_t(1) << 30)) return 0;             // a small matrix crosses PCIe faster than a thread pool starts
    return static_cast<int>(std::max(1u, std::min(32u, std::thread::hardware_concurrency())));
}

static cmg_status copyBackLastFacesAndExpand(cmg_ctx* ctx, const double* dPacked, int strips, double* outPacked, int threads)
{
    const int64_t n = ctx->npix, facePix = ctx->nside * ctx->nside;
    struct Chunk { int strip, ring, face; int64_t q0, q1; cudaEvent_t arrived; };
    std::vector<Chunk> chunks;
    // Images selected by cmg_set_host_expand_direct cross PCIe as well.  Off by default: on the measured host (16 cores, one
    // GPU) the copy engine idles for half of the call, yet handing it the third strip's first image made the call SLOWER
    // (991 ms against 840 ms) -- DMA writes and the host threads' streaming stores draw on one host memory write bandwidth
    // (~105 GB/s there), which is what bounds the call, not PCIe and not the cores.
    const int directMask = strips == 3 ? ctx->hostExpandDirectMask : 0;
    for(int pass = 0; pass <= 3; ++pass)
        for(int strip = 0; strip < strips; ++strip)
            for(int ring = 0; ring < 3; ++ring)
            {
                if(pass > 0 && !((directMask >> (3 * strip + pass - 1)) & 1))
                    continue;
                const int face = 4 * ring + 3 - pass;
                const int64_t col0 = strip * n + face * facePix;
                const int64_t step = std::max<int64_t>(8, std::min<int64_t>(facePix, (int64_t(8) << 20) / (col0 + facePix)));   // ~64 MB
                for(int64_t q = 0; q < facePix; q += step)
                    chunks.push_back({strip, ring, face, q, std::min(facePix, q + step), nullptr});
            }
    cmg_status st = CMG_OK;
    size_t issued = 0;
    for(; issued < chunks.size() && st == CMG_OK; ++issued)
    {
        Chunk& c = chunks[issued];
        const int64_t col0 = c.strip * n + c.face * facePix;
        const int64_t first = cmg_packed_size(col0 + c.q0), last = cmg_packed_size(col0 + c.q1);
        cudaError_t e = cudaMemcpyAsync(outPacked + first, dPacked + first, sizeof(double) * (last - first), cudaMemcpyDeviceToHost, ctx->stream);
        if(e == cudaSuccess) e = cudaEventCreateWithFlags(&c.arrived, cudaEventDisableTiming);
        if(e == cudaSuccess) e = cudaEventRecord(c.arrived, ctx->stream);
        if(e != cudaSuccess) st = cudaFail(ctx, e, "cudaMemcpyAsync (host expansion)");
        else ctx->bytesD2H += static_cast<int64_t>(sizeof(double)) * (last - first);
    }
    cmg::ExpandPipeline* pipe = cmg::expandBegin(outPacked, ctx->nside, threads, directMask);
    for(size_t k = 0; k < issued; ++k)
    {
        Chunk& c = chunks[k];
        if(!c.arrived) continue;
        const cudaError_t e = cudaEventSynchronize(c.arrived);
        cudaEventDestroy(c.arrived);
        if(e != cudaSuccess && st == CMG_OK) st = cudaFail(ctx, e, "cudaEventSynchronize");
        if(st == CMG_OK && (c.face & 3) == 3)
            cmg::expandPublish(pipe, c.strip, c.ring, c.q0, c.q1);
    }
    cmg::expandFinish(pipe);
    return st;
}

// a packed matrix of dimension strips x npix from device to host memory; on the full sky (the matrix of the context's current
// geometry) by the host expansion where that pays
static cmg_status matrixToHost(cmg_ctx* ctx, const double* dPacked, int strips, double* outPacked)
{
    const int64_t bytes = sizeof(double) * cmg_packed_size(strips * ctx->npix);
    if(const int threads = hostExpandThreadsFor(ctx, bytes))
        return copyBackLastFacesAndExpand(ctx, dPacked, strips, outPacked, threads);
    CMG_CUDA(ctx, cudaMemcpyAsync(outPacked, dPacked, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->bytesD2H += bytes;
    return CMG_OK;
}

static cmg_status wholeCallTT(cmg_ctx* ctx, const std::vector<double>& a, int lmax, double* dOut, double* outPacked)
{
    cmg_status s;
    if(!dOut)
    {
        if((s = ensureScratch(ctx, sizeof(double) * cmg_packed_size(ctx->npix))) != CMG_OK) return s;
        dOut = ctx->dScratch;
    }
    // full sky: one evaluation per orbit of pixel pairs under the pi/2 rotation of the grid (orbit.cuh), 3.2x less work
    const bool orbit = ctx->fullSky && ctx->nside >= 16 && lmax + 1 <= cmg::TT_STATIC_STEPS && ctx->tquVariant == 0;
    if((s = orbit ? cmg_legendre_series_orbit(ctx, a.data(), lmax, dOut)
                  : cmg_legendre_series(ctx, a.data(), lmax, 0, ctx->npix, dOut)) != CMG_OK) return s;
    return outPacked ? matrixToHost(ctx, dOut, 1, outPacked) : CMG_OK;
}

static cmg_status clToCMatrixImpl(cmg_ctx* ctx, const double* cl, int lmax, double fwhm, const double* pixwin, double* dOut, double* outPacked)
{
    cmg_status s = checkReady(ctx, lmax);
    if(s != CMG_OK) return s;
    if(!cl || (!outPacked && !dOut)) return fail(ctx, CMG_EINVAL, "null argument");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<double> f(lmax + 1), a(lmax + 1);
    // the reference has check(fwhm >= 0) (source/utils.cpp:56); a negative beam must not come back as an all-zero matrix
    if(cmg_window_beam(f.data(), lmax, fwhm, pixwin) != CMG_OK) return fail(ctx, CMG_EINVAL, "fwhm must be >= 0 (degrees)");
    if((s = cmg_tt_weights(cl, f.data(), lmax, a.data())) != CMG_OK) return fail(ctx, s, "bad C_l / window arguments");
    return wholeCallTT(ctx, a, lmax, dOut, outPacked);
}

static cmg_status fiducialImpl(cmg_ctx* ctx, const double* cl, int lmax, double fwhm, const double* pixwin, double* dOut, double* outPacked)
{
    if(!ctx) return CMG_EINVAL;
    const int lMaxMax = static_cast<int>(4 * ctx->nside);
    cmg_status s = checkReady(ctx, lMaxMax);
    if(s != CMG_OK) return s;
    if(!cl || (!outPacked && !dOut)) return fail(ctx, CMG_EINVAL, "null argument");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    std::vector<double> f(lMaxMax + 1), a(lMaxMax + 1);
    if(lmax < 2 || lmax > lMaxMax) return fail(ctx, CMG_EINVAL, "fiducial matrix: 2 <= lmax <= 4 nside");
    if(cmg_window_beam(f.data(), lMaxMax, fwhm, pixwin) != CMG_OK) return fail(ctx, CMG_EINVAL, "fwhm must be >= 0 (degrees)");
    if((s = cmg_fiducial_weights(cl, f.data(), ctx->nside, lmax, a.data())) != CMG_OK) return fail(ctx, s, "bad C_l / window arguments");
    return wholeCallTT(ctx, a, lMaxMax, dOut, outPacked);
}

cmg_status cmg_cl_to_cmatrix(cmg_ctx* ctx, const double* cl, int lmax, double fwhm, const double* pixwin, double* outPacked)
{
    if(!outPacked) return fail(ctx, CMG_EINVAL, "null argument");
    return clToCMatrixImpl(ctx, cl, lmax, fwhm, pixwin, nullptr, outPacked);
}

cmg_status cmg_cl_to_cmatrix_dev(cmg_ctx* ctx, const double* cl, int lmax, double fwhm, const double* pixwin, double* dOut)
{
    if(!dOut) return fail(ctx, CMG_EINVAL, "null argument");
    return clToCMatrixImpl(ctx, cl, lmax, fwhm, pixwin, dOut, nullptr);
}

cmg_status cmg_fiducial_matrix(cmg_ctx* ctx, const double* cl, int lmax, double fwhm, const double* pixwin, double* outPacked)
{
    if(!outPacked) return fail(ctx, CMG_EINVAL, "null argument");
    return fiducialImpl(ctx, cl, lmax, fwhm, pixwin, nullptr, outPacked);
}

cmg_status cmg_fiducial_matrix_dev(cmg_ctx* ctx, const double* cl, int lmax, double fwhm, const double* pixwin, double* dOut)
{
    if(!dOut) return fail(ctx, CMG_EINVAL, "null argument");
    return fiducialImpl(ctx, cl, lmax, fwhm, pixwin, dOut, nullptr);
}

cmg_status cmg_matrix_to_host(cmg_ctx* ctx, const double* dPacked, int64_t dim, int full_sky_strips, double* outPacked)
{
    if(!ctx || !dPacked || !outPacked || dim < 1) return CMG_EINVAL;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    if(full_sky_strips == 1 || full_sky_strips == 3)
    {
        // the caller vouches that this is the full-sky NESTED matrix of the context's current geometry ([T] or [T;Q;U])
        if(!ctx->fullSky || dim != full_sky_strips * ctx->npix)
            return fail(ctx, CMG_EINVAL, "cmg_matrix_to_host: not the full-sky matrix of the current geometry");
        return matrixToHost(ctx, dPacked, full_sky_strips, outPacked);
    }
    const int64_t bytes = sizeof(double) * cmg_packed_size(dim);
    CMG_CUDA(ctx, cudaMemcpyAsync(outPacked, dPacked, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->bytesD2H += bytes;
    return CMG_OK;
}

cmg_status cmg_mask_matrix(cmg_ctx* ctx, const double* dIn, int64_t npixIn, const int32_t* good, int64_t nGood, double* dOut)
{
    if(!ctx || !dIn || !good || !dOut || nGood < 1 || npixIn < 1) return CMG_EINVAL;
    for(int64_t k = 0; k < nGood; ++k)
        if(good[k] < 0 || good[k] >= npixIn)
            return fail(ctx, CMG_EINVAL, "good pixel index outside the input matrix");
    if(nGood > 65535)
        return fail(ctx, CMG_EUNSUPPORTED, "maskMatrix gather is limited to 65535 good pixels per call");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    if(nGood > ctx->indexCap)
    {
        if(ctx->dIndex)
        {
            CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            CMG_CUDA(ctx, cudaFree(ctx->dIndex));
            ctx->dIndex = nullptr;
            ctx->indexCap = 0;
        }
        CMG_CUDA(ctx, cudaMalloc(&ctx->dIndex, sizeof(int32_t) * nGood));
        ctx->indexCap = nGood;
    }
    CMG_CUDA(ctx, cudaMemcpyAsync(ctx->dIndex, good, sizeof(int32_t) * nGood, cudaMemcpyHostToDevice, ctx->stream));
    const dim3 grid(static_cast<unsigned>((nGood + 255) / 256), static_cast<unsigned>(nGood));
    cmg::maskGatherKernel<<<grid, 256, 0, ctx->stream>>>(dIn, ctx->dIndex, nGood, dOut);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return CMG_OK;
}

// ---------------------------------------------------------------- T,Q,U

cmg_status cmg_tqu_layout_single(cmg_ctx* ctx, double* dPacked, cmg_tqu_layout* layout)
{
    if(!ctx || !layout || !dPacked) return CMG_EINVAL;
    if(ctx->npix <= 0) return fail(ctx, CMG_ESTATE, "cmg_set_pixels has not been called on this context");
    std::memset(layout, 0, sizeof(*layout));
    const int64_t n = ctx->npix;
    layout->n_parts = 1;
    layout->own = 0;
    layout->begin[0] = 0;
    layout->begin[1] = n;
    layout->ptr[0][0] = dPacked;
    layout->ptr[0][1] = dPacked + cmg_packed_size(n);          // entry (0, N)
    layout->ptr[0][2] = dPacked + cmg_packed_size(2 * n);      // entry (0, 2N)
    return CMG_OK;
}

cmg_status cmg_tqu(cmg_ctx* ctx, const double* att, const double* ate, const double* aee, const double* abb, int lmax,
                   const cmg_tqu_layout* layout)
{
    cmg_status s = checkReady(ctx, lmax);
    if(s != CMG_OK) return s;
    if(!att || !ate || !aee || !abb) return fail(ctx, CMG_EINVAL, "null weights");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t n1 = lmax + 1;
    if((s = ensureWeights(ctx, 4 * n1)) != CMG_OK) return s;
    CMG_CUDA(ctx, cudaMemcpyAsync(ctx->dWeights, att, sizeof(double) * n1, cudaMemcpyHostToDevice, ctx->stream));
    CMG_CUDA(ctx, cudaMemcpyAsync(ctx->dWeights + n1, ate, sizeof(double) * n1, cudaMemcpyHostToDevice, ctx->stream));
    CMG_CUDA(ctx, cudaMemcpyAsync(ctx->dWeights + 2 * n1, aee, sizeof(double) * n1, cudaMemcpyHostToDevice, ctx->stream));
    CMG_CUDA(ctx, cudaMemcpyAsync(ctx->dWeights + 3 * n1, abb, sizeof(double) * n1, cudaMemcpyHostToDevice, ctx->stream));
    const double* hostW[4] = {att, ate, aee, abb};
    return launchTqu(ctx, ctx->dWeights, 0, lmax, 1, layout, 0, hostW);
}

cmg_status cmg_tqu_dev(cmg_ctx* ctx, const double* dA, int lmax, const cmg_tqu_layout* layout)
{
    const cmg_status s = checkReady(ctx, lmax);
    if(s != CMG_OK) return s;
    if(!dA) return fail(ctx, CMG_EINVAL, "null weights");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    return launchTqu(ctx, dA, 0, lmax, 1, layout, 0, nullptr);
}

// ---------------------------------------------------------------- T,Q,U and TT over symmetry orbits (full sky)

cmg_status cmg_tqu_orbit_plan(int64_t nside, int mode, int32_t* out, int32_t* nClasses)
{
    if(!out || !nClasses || !cmg::validNside(nside) || (mode != 0 && mode != 1 && mode != 3))
        return CMG_EINVAL;
    cmg::OrbitPlan plan;
    cmg::orbitBuildPlan(nside, mode, -1, plan);
    *nClasses = plan.n;
    for(int c = 0; c < plan.n; ++c)
    {
        const cmg::OrbitClass& oc = plan.c[c];
        int32_t* o = out + c * CMG_ORBIT_CLASS_INTS;
        o[0] = oc.rowFace; o[1] = oc.colFace; o[2] = oc.tri; o[3] = oc.sameFace; o[4] = oc.nImg;
        for(int k = 0; k < cmg::ORB_MAX_IMAGES; ++k)
        {
            o[5 + 3 * k] = oc.imgRowFace[k];
            o[6 + 3 * k] = oc.imgColFace[k];
            o[7 + 3 * k] = oc.imgSwap[k];
            o[17 + k] = oc.comboBase[k];
        }
    }
    return CMG_OK;
}

cmg_status cmg_tqu_orbit_plan_mirror(int64_t nside, int mode, int32_t* out, int32_t* nClasses)
{
    if(!out || !nClasses || !cmg::validNside(nside) || (mode != 0 && mode != 1 && mode != 3))
        return CMG_EINVAL;
    cmg::OrbitPlan plan;
    cmg::orbitBuildPlan(nside, mode, -1, plan);
    *nClasses = plan.n;
    for(int c = 0; c < plan.n; ++c)
    {
        out[c * 5] = plan.c[c].mirror;
        for(int k = 0; k < cmg::ORB_MAX_IMAGES; ++k)
            out[c * 5 + 1 + k] = plan.c[c].mirRowFace[k];
    }
    return CMG_OK;
}

namespace
{

// offsets of the destination blocks inside rank `rank`'s outbox (orbit.cuh, OrbitShardDev); off[nRanks] = total doubles
void orbitOutboxOffsets(const cmg::OrbitPlan& plan, int nRanks, const int64_t* bounds, int rank, int64_t* off)
{
    const int64_t nct = (bounds[rank + 1] - bounds[rank]) / cmg::PQ_TJ;
    int64_t at = 0;
    for(int d = 0; d < nRanks; ++d)
    {
        off[d] = at;
        if(d == rank)
            continue;
        const int64_t nh = (bounds[d + 1] - bounds[d]) / cmg::ORB_SUB;
        const int64_t combos = plan.nComboA + (d < rank ? plan.nComboB : 0);
        at += combos * nct * nh * (cmg::ORB_SUB * cmg::ORB_SUB);
    }
    off[nRanks] = at;
}

cmg_status orbitCheckBounds(cmg_ctx* ctx, int64_t nside, int nRanks, const int64_t* bounds, int rank)
{
    const int64_t facePix = nside * nside;
    if(nRanks < 1 || nRanks > cmg::ORB_MAX_RANKS || rank < 0 || rank >= nRanks || !bounds)
        return fail(ctx, CMG_EINVAL, "shard: 1 <= n_ranks <= 16 and 0 <= rank < n_ranks");
    if(bounds[0] != 0 || bounds[nRanks] != facePix)
        return fail(ctx, CMG_EINVAL, "shard: bounds must run from 0 to nside^2");
    for(int r = 0; r < nRanks; ++r)
        if(bounds[r] > bounds[r + 1] || bounds[r] % cmg::PQ_TJ)
            return fail(ctx, CMG_EINVAL, "shard: bounds must be ascending multiples of 32");
    return CMG_OK;
}

// checks a shard descriptor and turns it into the kernels' form (strip bases adjusted by the packed offset of their first column)
cmg_status orbitShardDev(cmg_ctx* ctx, const cmg_orbit_shard* shard, int mode, cmg::OrbitShardDev& sh)
{
    if(!shard) return fail(ctx, CMG_EINVAL, "null shard");
    if(mode < 0 || mode > 3)
        return fail(ctx, CMG_EINVAL, "orbit mode must be 0 (transposed images), 1 (none), 2 (0 without the row-pointer table) or 3 (0 + the meridian mirror)");
    if(mode == 3 && shard->n_ranks != 1)
        return fail(ctx, CMG_EUNSUPPORTED, "orbit mode 3 (meridian mirror) needs a single owner: the mirror image of a rank's columns are another rank's");
    if(mode == 2) mode = 0;                          // same classes, same storage
    if(!ctx->fullSky)
        return fail(ctx, CMG_EUNSUPPORTED, "the symmetry-orbit path needs the full sky in NESTED order (cmg_set_pixels with good_nest = NULL)");
    if(ctx->nside < 8)
        return fail(ctx, CMG_EUNSUPPORTED, "the symmetry-orbit path needs nside >= 8 (whole 64 x 32 tiles inside a base face)");
    const int64_t facePix = ctx->nside * ctx->nside, n = ctx->npix;
    cmg_status s = orbitCheckBounds(ctx, ctx->nside, shard->n_ranks, shard->bounds, shard->rank);
    if(s != CMG_OK) return s;
    std::memset(&sh, 0, sizeof(sh));
    sh.nRanks = shard->n_ranks;
    sh.rank = shard->rank;
    sh.q0 = static_cast<int>(shard->bounds[shard->rank]);
    sh.q1 = static_cast<int>(shard->bounds[shard->rank + 1]);
    for(int r = 0; r <= cmg::ORB_MAX_RANKS; ++r)       // padded behind the last rank: orbitOwnerOfHalfTile counts the boundaries <= h
        sh.boundH[r] = r < sh.nRanks ? static_cast<int>(shard->bounds[r] / cmg::ORB_SUB) : 0x7fffffff;
    sh.boundH[sh.nRanks] = static_cast<int>(shard->bounds[sh.nRanks] / cmg::ORB_SUB);
    cmg::OrbitPlan plan;
    cmg::orbitBuildPlan(ctx->nside, mode, -1, plan);
    int64_t off[cmg::ORB_MAX_RANKS + 1];
    orbitOutboxOffsets(plan, sh.nRanks, shard->bounds, sh.rank, off);
    for(int d = 0; d < sh.nRanks; ++d)
        sh.destOff[d] = off[d];
    if(off[sh.nRanks] > 0 && !shard->outbox && sh.q1 > sh.q0) return fail(ctx, CMG_EINVAL, "shard: null outbox");
    sh.outbox = shard->outbox;
    for(int f = 0; f < 12; ++f)
        for(int st = 0; st < 3; ++st)
        {
            if(!shard->strip[st][f] && sh.q1 > sh.q0) return fail(ctx, CMG_EINVAL, "shard: null strip");
            sh.strip[st][f] = shard->strip[st][f] - cmg::packedOffset(st * n + f * facePix + sh.q0);
        }
    return CMG_OK;
}

} // namespace

cmg_status cmg_tqu_orbit_sharded(cmg_ctx* ctx, const double* att, const double* ate, const double* aee, const double* abb, int lmax,
                                 const cmg_orbit_shard* shard, int mode)
{
    cmg_status s = checkReady(ctx, lmax);
    if(s != CMG_OK) return s;
    if(!att || !ate || !aee || !abb) return fail(ctx, CMG_EINVAL, "null weights");
    cmg::OrbitShardDev sh;
    if((s = orbitShardDev(ctx, shard, mode, sh)) != CMG_OK) return s;
    if(lmax < 2 || lmax > cmg::PQ_STATIC_LMAX)
        return fail(ctx, CMG_EUNSUPPORTED, "the symmetry-orbit path needs 2 <= lmax <= PQ_STATIC_LMAX");
    if(sh.q0 == sh.q1) return CMG_OK;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));

    static thread_local cmg::TquStaticTable T;      // 28 KB: keep it off the stack
    fillStaticTable(ctx, att, ate, aee, abb, lmax, T);
    const int entrySlot = cmg::PQ_STATIC_STEPS + 1 - lmax;
    const int64_t facePix = ctx->nside * ctx->nside;
    const int64_t tiles64 = (facePix / cmg::PQ_TI) * ((sh.q1 - sh.q0) / cmg::PQ_TJ);
    if(tiles64 > 0x7fffffffLL)
        return fail(ctx, CMG_EUNSUPPORTED, "too many tiles for one launch");
    const unsigned tiles = static_cast<unsigned>(tiles64);

    KernelTimer timer(ctx);
    const int masks[4] = {0, 8, 12, 16};             // classes by their transposed images: none, (3,0), (2,0) + (3,1); 16: with mirror images (mode 3)
    for(int m = 0; m < 4; ++m)
    {
        cmg::OrbitPlan plan;
        cmg::orbitBuildPlan(ctx->nside, mode == 2 ? 0 : mode, masks[m], plan);
        if(plan.n == 0)
            continue;
        const dim3 grid(tiles, static_cast<unsigned>(plan.n));
#define CMG_ORBIT_LAUNCH(MASK)                                                                                                      \
        {                                                                                                                           \
            auto kernel = cmg::tquOrbitKernel<4, 2, MASK>;                                                                          \
            const size_t smem = sizeof(double) * cmg::orbitSmemDoubles<(MASK) != 0>();                                              \
            CMG_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));       \
            kernel<<<grid, cmg::PQ_THREADS, smem, ctx->stream>>>(T, geometryOf(ctx), entrySlot, plan, sh);                          \
        }
        if(masks[m] == 16)
        {
            // whole face pairs of different rings + the four rotations of their mirror image: eight images per evaluated pair
            auto kernel = cmg::tquOrbitKernel<4, 2, 0, true, true>;
            const size_t smem = sizeof(double) * cmg::orbitSmemDoubles<false, true, true>();
            CMG_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            kernel<<<grid, cmg::PQ_THREADS, smem, ctx->stream>>>(T, geometryOf(ctx), entrySlot, plan, sh);
        }
        else if(masks[m] == 0 && (mode == 0 || mode == 3))
        {
            // classes without transposed images: store destinations from a per-tile table in shared memory (36.2 against 36.8 ms)
            auto kernel = cmg::tquOrbitKernel<4, 2, 0, true>;
            const size_t smem = sizeof(double) * cmg::orbitSmemDoubles<false, true>();
            CMG_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
            kernel<<<grid, cmg::PQ_THREADS, smem, ctx->stream>>>(T, geometryOf(ctx), entrySlot, plan, sh);
        }
        else if(masks[m] == 0) CMG_ORBIT_LAUNCH(0)
        else if(masks[m] == 8) CMG_ORBIT_LAUNCH(8)
        else CMG_ORBIT_LAUNCH(12)
#undef CMG_ORBIT_LAUNCH
        CMG_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    return timer.finish();
}

cmg_status cmg_tqu_orbit(cmg_ctx* ctx, const double* att, const double* ate, const double* aee, const double* abb, int lmax,
                         double* dPacked, int mode)
{
    if(!ctx) return CMG_EINVAL;
    if(!dPacked) return fail(ctx, CMG_EINVAL, "null output");
    if(ctx->npix <= 0) return fail(ctx, CMG_ESTATE, "cmg_set_pixels has not been called on this context");
    cmg_orbit_shard shard;
    std::memset(&shard, 0, sizeof(shard));
    const int64_t facePix = ctx->nside * ctx->nside, n = ctx->npix;
    shard.n_ranks = 1;
    shard.rank = 0;
    shard.bounds[0] = 0;
    shard.bounds[1] = facePix;
    for(int st = 0; st < 3; ++st)
        for(int f = 0; f < 12; ++f)
            shard.strip[st][f] = dPacked + cmg::packedOffset(st * n + f * facePix);
    return cmg_tqu_orbit_sharded(ctx, att, ate, aee, abb, lmax, &shard, mode);
}

namespace
{
// TT over orbits: columns [q0, q1) of all twelve faces into strip[f] (f F + [q0, q1) each); mode 0 needs the single owner
cmg_status launchTtOrbit(cmg_ctx* ctx, const double* a, int lmax, int mode, int64_t q0, int64_t q1, double* const* strip)
{
    const int64_t facePix = ctx->nside * ctx->nside;
    static thread_local cmg::TtStaticTable T;
    for(int i = 0; i < cmg::TT_STATIC_STEPS; ++i)
    {
        const int k = cmg::TT_STATIC_STEPS - 1 - i;
        T.s[i] = make_double2(k <= lmax ? a[k] * ctx->hostT0.N[k] : 0.0, -ctx->hostT0.g[k + 1]);
    }
    const int entrySlot = cmg::TT_STATIC_STEPS - 1 - lmax;
    cmg::OrbitTtShardDev sh;
    sh.q0 = static_cast<int>(q0);
    sh.q1 = static_cast<int>(q1);
    for(int f = 0; f < 12; ++f)
        sh.strip[f] = strip[f] - cmg::packedOffset(f * facePix + q0);
    const int64_t tiles64 = (facePix / cmg::TT_ROWS) * ((q1 - q0) / cmg::TT_COLS);
    if(tiles64 > 0x7fffffffLL)
        return fail(ctx, CMG_EUNSUPPORTED, "too many tiles for one launch");
    if(tiles64 == 0)
        return CMG_OK;
    KernelTimer timer(ctx);
    const int masks[3] = {0, 8, 12};
    for(int m = 0; m < (mode == 0 ? 3 : 1); ++m)
    {
        cmg::OrbitPlan plan;
        cmg::orbitBuildPlan(ctx->nside, mode, mode == 0 ? masks[m] : -1, plan);
        if(plan.n == 0)
            continue;
        const dim3 grid(static_cast<unsigned>(tiles64), static_cast<unsigned>(plan.n));
        if(mode != 0 || masks[m] == 0)
            cmg::legendreSeriesOrbitKernel<8, 4, 0><<<grid, cmg::TT_ROWS, 0, ctx->stream>>>(T, geometryOf(ctx), entrySlot, plan, sh);
        else if(masks[m] == 8)
            cmg::legendreSeriesOrbitKernel<8, 4, 8><<<grid, cmg::TT_ROWS, 0, ctx->stream>>>(T, geometryOf(ctx), entrySlot, plan, sh);
        else
            cmg::legendreSeriesOrbitKernel<8, 4, 12><<<grid, cmg::TT_ROWS, 0, ctx->stream>>>(T, geometryOf(ctx), entrySlot, plan, sh);
        CMG_CUDA(ctx, cudaGetLastError());
        ctx->launches += 1;
    }
    return timer.finish();
}

cmg_status ttOrbitReady(cmg_ctx* ctx, const double* a, int lmax)
{
    cmg_status s = checkReady(ctx, lmax);
    if(s != CMG_OK) return s;
    if(!a) return fail(ctx, CMG_EINVAL, "null argument");
    if(!ctx->fullSky)
        return fail(ctx, CMG_EUNSUPPORTED, "the symmetry-orbit path needs the full sky in NESTED order (cmg_set_pixels with good_nest = NULL)");
    if(ctx->nside < 16)
        return fail(ctx, CMG_EUNSUPPORTED, "the TT symmetry-orbit path needs nside >= 16 (whole 128 x 16 tiles inside a base face)");
    if(lmax + 1 > cmg::TT_STATIC_STEPS)
        return fail(ctx, CMG_EUNSUPPORTED, "lmax exceeds the static coefficient table");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    return CMG_OK;
}
} // namespace

cmg_status cmg_legendre_series_orbit(cmg_ctx* ctx, const double* a, int lmax, double* dOut)
{
    cmg_status s = ttOrbitReady(ctx, a, lmax);
    if(s != CMG_OK) return s;
    if(!dOut) return fail(ctx, CMG_EINVAL, "null argument");
    const int64_t facePix = ctx->nside * ctx->nside;
    double* strip[12];
    for(int f = 0; f < 12; ++f)
        strip[f] = dOut + cmg::packedOffset(f * facePix);
    // the single owner takes the plan with transposed images (a quarter of the pairs evaluated, one launch per transposed-image
    // mask) from Nside = 32 on; below that the three launches cost more than the 20 % of work they save (Nside = 16, lmax = 47:
    // 0.053 ms against 0.034 ms for the single launch of the plan without transposed images)
    return launchTtOrbit(ctx, a, lmax, ctx->nside >= 32 ? 0 : 1, 0, facePix, strip);
}

cmg_status cmg_legendre_series_orbit_sharded(cmg_ctx* ctx, const double* a, int lmax, int64_t qBegin, int64_t qEnd, double* const* dStrips)
{
    cmg_status s = ttOrbitReady(ctx, a, lmax);
    if(s != CMG_OK) return s;
    const int64_t facePix = ctx->nside * ctx->nside;
    if(!dStrips || qBegin < 0 || qEnd > facePix || qBegin > qEnd || qBegin % cmg::TT_COLS || qEnd % cmg::TT_COLS)
        return fail(ctx, CMG_EINVAL, "shard range must be multiples of 16 inside [0, nside^2]");
    for(int f = 0; f < 12 && qEnd > qBegin; ++f)
        if(!dStrips[f]) return fail(ctx, CMG_EINVAL, "shard: null strip");
    // without transposed images every entry lands in a column of the rank that evaluated it: no exchange.  A range that is the
    // whole face is the single owner again and may take them.
    const int mode = (qBegin == 0 && qEnd == facePix && ctx->nside >= 32) ? 0 : 1;
    return launchTtOrbit(ctx, a, lmax, mode, qBegin, qEnd, dStrips);
}

namespace
{
// block(sender -> receiver) -> the receiver's packed columns; `strips` = the receiver's ADJUSTED strip bases
cmg_status launchInboxScatter(cmg_ctx* ctx, int mode, int nRanks, const int64_t* bounds, int sender, int receiver, const double* dBlock,
                              double* const (*strips)[12])
{
    cmg::OrbitPlan plan;
    cmg::orbitBuildPlan(ctx->nside, mode, -1, plan);
    cmg::OrbitInboxArgs a;
    a.npix = ctx->npix;
    a.senderQ0 = static_cast<int>(bounds[sender]);
    a.nct = static_cast<int>((bounds[sender + 1] - bounds[sender]) / cmg::PQ_TJ);
    a.h0 = static_cast<int>(bounds[receiver] / cmg::ORB_SUB);
    a.nh = static_cast<int>((bounds[receiver + 1] - bounds[receiver]) / cmg::ORB_SUB);
    a.triLive = receiver < sender ? 1 : 0;
    a.block = dBlock;
    for(int st = 0; st < 3; ++st)
        for(int f = 0; f < 12; ++f)
            a.strip[st][f] = strips[st][f];
    (void) nRanks;
    if(a.nct == 0 || a.nh == 0) return CMG_OK;
    const dim3 grid(static_cast<unsigned>(a.nct) * static_cast<unsigned>(a.nh), static_cast<unsigned>(plan.n));
    cmg::orbitInboxScatterKernel<<<grid, cmg::PQ_THREADS, 0, ctx->stream>>>(plan, a);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return CMG_OK;
}
} // namespace

cmg_status cmg_orbit_outbox_layout(int64_t nside, int mode, int nRanks, const int64_t* bounds, int rank, int64_t* offsets)
{
    if(!offsets || !cmg::validNside(nside) || nside < 8 || mode < 0 || mode > 2) return CMG_EINVAL;
    if(orbitCheckBounds(nullptr, nside, nRanks, bounds, rank) != CMG_OK) return CMG_EINVAL;
    cmg::OrbitPlan plan;
    cmg::orbitBuildPlan(nside, mode == 2 ? 0 : mode, -1, plan);
    orbitOutboxOffsets(plan, nRanks, bounds, rank, offsets);
    return CMG_OK;
}

cmg_status cmg_tqu_orbit_scatter_inbox(cmg_ctx* ctx, const cmg_orbit_shard* shard, int mode, int sender, const double* dBlock)
{
    if(!ctx) return CMG_EINVAL;
    if(ctx->npix <= 0) return fail(ctx, CMG_ESTATE, "cmg_set_pixels has not been called on this context");
    cmg::OrbitShardDev sh;
    cmg_status s = orbitShardDev(ctx, shard, mode, sh);
    if(s != CMG_OK) return s;
    if(sender < 0 || sender >= sh.nRanks || sender == sh.rank) return fail(ctx, CMG_EINVAL, "sender must be another rank of the shard");
    if(!dBlock) return fail(ctx, CMG_EINVAL, "null block");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    return launchInboxScatter(ctx, mode == 2 ? 0 : mode, sh.nRanks, shard->bounds, sender, sh.rank, dBlock, sh.strip);
}

cmg_status cmg_tqu_orbit_assemble(cmg_ctx* ctx, const cmg_orbit_shard* shard, int mode, int parts, double* dFull)
{
    if(!ctx) return CMG_EINVAL;
    if(!dFull) return fail(ctx, CMG_EINVAL, "null output");
    if(ctx->npix <= 0) return fail(ctx, CMG_ESTATE, "cmg_set_pixels has not been called on this context");
    cmg::OrbitShardDev sh;
    cmg_status s = orbitShardDev(ctx, shard, mode, sh);
    if(s != CMG_OK) return s;
    if(mode == 2) mode = 0;                          // mode 2 stores exactly what mode 0 stores
    if(sh.q0 == sh.q1) return CMG_OK;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t facePix = ctx->nside * ctx->nside, n = ctx->npix;
    for(int st = 0; st < 3 && (parts & 1); ++st)
        for(int f = 0; f < 12; ++f)
        {
            const int64_t first = cmg::packedOffset(st * n + f * facePix + sh.q0), last = cmg::packedOffset(st * n + f * facePix + sh.q1);
            if(shard->strip[st][f] != dFull + first)
                CMG_CUDA(ctx, cudaMemcpyAsync(dFull + first, shard->strip[st][f], sizeof(double) * (last - first), cudaMemcpyDeviceToDevice, ctx->stream));
        }
    if(sh.nRanks == 1 || !(parts & 2))
        return CMG_OK;
    double* whole[3][12];
    for(int st = 0; st < 3; ++st)
        for(int f = 0; f < 12; ++f)
            whole[st][f] = dFull;                    // the adjusted base of every run of columns of a whole triangle is its first element
    for(int d = 0; d < sh.nRanks; ++d)
        if(d != sh.rank && sh.destOff[d] != (d + 1 < sh.nRanks ? sh.destOff[d + 1] : -1))
            if((s = launchInboxScatter(ctx, mode, sh.nRanks, shard->bounds, sh.rank, d, sh.outbox + sh.destOff[d], whole)) != CMG_OK) return s;
    return CMG_OK;
}

cmg_status cmg_host_register(void* ptr, int64_t bytes)
{
    if(!ptr || bytes <= 0) return CMG_EINVAL;
    const cudaError_t e = cudaHostRegister(ptr, static_cast<size_t>(bytes), cudaHostRegisterPortable);
    if(e != cudaSuccess) return cudaFail(nullptr, e, "cudaHostRegister");
    return CMG_OK;
}

cmg_status cmg_host_unregister(void* ptr)
{
    const cudaError_t e = cudaHostUnregister(ptr);
    if(e != cudaSuccess) return cudaFail(nullptr, e, "cudaHostUnregister");
    return CMG_OK;
}

cmg_status cmg_orbit_strips_to_host(cmg_ctx* ctx, const cmg_orbit_shard* shard, double* hostPacked, int threads, int directMask)
{
    if(!ctx) return CMG_EINVAL;
    if(!hostPacked || threads < 0 || directMask < 0 || directMask > 511) return fail(ctx, CMG_EINVAL, "null destination, negative thread count or direct mask outside 0..511");
    if(ctx->npix <= 0) return fail(ctx, CMG_ESTATE, "cmg_set_pixels has not been called on this context");
    cmg::OrbitShardDev sh;
    cmg_status s = orbitShardDev(ctx, shard, 0, sh);
    if(s != CMG_OK) return s;
    if(sh.q0 == sh.q1) return CMG_OK;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t facePix = ctx->nside * ctx->nside, n = ctx->npix;
    if(threads == 0)
    {
        for(int st = 0; st < 3; ++st)
            for(int f = 0; f < 12; ++f)
            {
                const int64_t first = cmg::packedOffset(st * n + f * facePix + sh.q0), last = cmg::packedOffset(st * n + f * facePix + sh.q1);
                CMG_CUDA(ctx, cudaMemcpyAsync(hostPacked + first, shard->strip[st][f], sizeof(double) * (last - first), cudaMemcpyDeviceToHost, ctx->stream));
                ctx->bytesD2H += static_cast<int64_t>(sizeof(double)) * (last - first);
            }
        CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return CMG_OK;
    }
    // the rank's columns of the last face of every ring, in chunks of ~64 MB; workers fill in the images of a chunk once it is there.
    // Images selected by directMask (bit 3 strip + k - 1: the face k below the last one) come over PCIe as well, behind the
    // last-face columns: the split between the copy engines and the host threads is the caller's to balance.
    struct Chunk { int strip, ring, face; int64_t q0, q1; cudaEvent_t arrived; };
    std::vector<Chunk> chunks;
    for(int pass = 0; pass <= 3; ++pass)
        for(int st = 0; st < 3; ++st)
            for(int ring = 0; ring < 3; ++ring)
            {
                if(pass > 0 && !((directMask >> (3 * st + pass - 1)) & 1))
                    continue;
                const int face = 4 * ring + 3 - pass;
                const int64_t col0 = st * n + face * facePix;
                const int64_t step = std::max<int64_t>(8, std::min<int64_t>(facePix, (int64_t(8) << 20) / (col0 + facePix)));
                for(int64_t q = sh.q0; q < sh.q1; q += step)
                    chunks.push_back({st, ring, face, q, std::min<int64_t>(sh.q1, q + step), nullptr});
            }
    cmg_status st = CMG_OK;
    size_t issued = 0;
    for(; issued < chunks.size() && st == CMG_OK; ++issued)
    {
        Chunk& c = chunks[issued];
        const int face = c.face;
        const int64_t col0 = c.strip * n + face * facePix;
        const int64_t first = cmg::packedOffset(col0 + c.q0), last = cmg::packedOffset(col0 + c.q1);
        const double* src = shard->strip[c.strip][face] + (first - cmg::packedOffset(col0 + sh.q0));
        cudaError_t e = cudaMemcpyAsync(hostPacked + first, src, sizeof(double) * (last - first), cudaMemcpyDeviceToHost, ctx->stream);
        if(e == cudaSuccess) e = cudaEventCreateWithFlags(&c.arrived, cudaEventDisableTiming);
        if(e == cudaSuccess) e = cudaEventRecord(c.arrived, ctx->stream);
        if(e != cudaSuccess) st = cudaFail(ctx, e, "cudaMemcpyAsync (strips to host)");
        else ctx->bytesD2H += static_cast<int64_t>(sizeof(double)) * (last - first);
    }
    cmg::ExpandPipeline* pipe = cmg::expandBegin(hostPacked, ctx->nside, threads, directMask);
    for(size_t k = 0; k < issued; ++k)
    {
        Chunk& c = chunks[k];
        if(!c.arrived) continue;
        const cudaError_t e = cudaEventSynchronize(c.arrived);
        cudaEventDestroy(c.arrived);
        if(e != cudaSuccess && st == CMG_OK) st = cudaFail(ctx, e, "cudaEventSynchronize");
        if(st == CMG_OK && (c.face & 3) == 3)
            cmg::expandPublish(pipe, c.strip, c.ring, c.q0, c.q1);
    }
    cmg::expandFinish(pipe);
    return st;
}

cmg_status cmg_tqu_batched(cmg_ctx* ctx, const double* a, int lmax, int64_t nBatch, double* dOut, int64_t stride)
{
    cmg_status s = checkReady(ctx, lmax);
    if(s != CMG_OK) return s;
    if(!a || !dOut || nBatch < 1) return fail(ctx, CMG_EINVAL, "bad batch arguments");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t per = 4 * (lmax + 1);
    if((s = ensureWeights(ctx, nBatch * per)) != CMG_OK) return s;
    CMG_CUDA(ctx, cudaMemcpyAsync(ctx->dWeights, a, sizeof(double) * nBatch * per, cudaMemcpyHostToDevice, ctx->stream));
    cmg_tqu_layout layout;
    if((s = cmg_tqu_layout_single(ctx, dOut, &layout)) != CMG_OK) return s;
    // Otherwise (long series, or an explicit variant): one Clenshaw pass per batch element in one launch (blockIdx.z),
    // 0.195 ms per Nside=16 lmax=47 matrix on a B200.
    // Variant 900 selects the shared-basis kernel (recurrences once per pixel pair and batch chunk, 4 accumulate-FMAs
    // per element and l).  It does 2.5x fewer FP64 operations but is bound by delivering one distinct weight per FMA
    // from shared memory (0.254 ms per matrix measured); the contraction belongs on the FP64 tensor path (DMMA) -- next round.
    if(ctx->tquVariant == 0 && lmax >= 2 && lmax <= cmg::PQ_STATIC_LMAX)
    {
        // one static-table launch per element (every DFMA with a uniform-register operand), spread over side streams so
        // that the small grids of low-Nside matrices overlap each other's heads and tails
        cmg::PartTable PS;
        std::memset(&PS, 0, sizeof(PS));
        PS.n = 1;
        PS.begin[1] = ctx->npix;
        const int64_t rowBlocks = (ctx->npix + cmg::PQ_TI - 1) / cmg::PQ_TI, colBlocks = (ctx->npix + cmg::PQ_TJ - 1) / cmg::PQ_TJ;
        if(colBlocks <= 65535)
        {
            const dim3 grid(static_cast<unsigned>(rowBlocks), static_cast<unsigned>(colBlocks), 1);
            cmg::TquDynamicArgs dyn;
            dyn.a = ctx->dWeights; dyn.aStride = 0; dyn.tab = tablesOf(ctx); dyn.lmax = lmax;
            static thread_local cmg::TquStaticTable TB;
            const int entrySlot = cmg::PQ_STATIC_STEPS + 1 - lmax;
            KernelTimer timerS(ctx);
            CMG_CUDA(ctx, cudaEventRecord(ctx->forkEv, ctx->stream));
            for(int k = 0; k < cmg_ctx::kAux; ++k)
                CMG_CUDA(ctx, cudaStreamWaitEvent(ctx->aux[k], ctx->forkEv, 0));
            for(int64_t b = 0; b < nBatch; ++b)
            {
                const double* wb = a + b * per;
                fillStaticTable(ctx, wb, wb + (lmax + 1), wb + 2 * (lmax + 1), wb + 3 * (lmax + 1), lmax, TB);
                for(int st = 0; st < 3; ++st) PS.ptr[0][st] = layout.ptr[0][st] + b * stride;
                if((s = launchTquVariant<4, true, 2>(ctx, dyn, TB, entrySlot, PS, grid, 0, ctx->aux[b % cmg_ctx::kAux])) != CMG_OK) return s;
            }
            for(int k = 0; k < cmg_ctx::kAux; ++k)
            {
                CMG_CUDA(ctx, cudaEventRecord(ctx->auxDone[k], ctx->aux[k]));
                CMG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->auxDone[k], 0));
            }
            ctx->launches += nBatch;
            return timerS.finish();
        }
    }
    if((ctx->tquVariant != 900 && ctx->tquVariant != 901) || lmax < 2)
        return launchTqu(ctx, ctx->dWeights, per, lmax, nBatch, &layout, stride, nullptr);
    if(ctx->tquVariant == 901)
    {
        // FP64 tensor path: the batch contraction as DMMA (kernels.cuh, tquBatchedMmaKernel)
        cmg::PartTable PM;
        std::memset(&PM, 0, sizeof(PM));
        PM.n = 1;
        PM.begin[1] = ctx->npix;
        for(int st = 0; st < 3; ++st) PM.ptr[0][st] = layout.ptr[0][st];
        const int64_t tilesM = (ctx->npix + cmg::MB_T - 1) / cmg::MB_T;
        const size_t smemM = cmg::tquMmaSmemBytes(lmax);
        if(smemM > 227 * 1024 || tilesM > 65535 || (cmg::mmaKPad(lmax) / 4) * 4 * 32 > cmg::MB_BN * 3 * cmg::MB_SP)
            return launchTqu(ctx, ctx->dWeights, per, lmax, nBatch, &layout, stride, nullptr);
        const int64_t nChunksM = (nBatch + cmg::MB_BN - 1) / cmg::MB_BN;
        const int64_t fragDoubles = nChunksM * (cmg::mmaKPad(lmax) / 4) * 4 * 32;
        if((s = ensureWeights(ctx, nBatch * per + fragDoubles)) != CMG_OK) return s;
        CMG_CUDA(ctx, cudaMemcpyAsync(ctx->dWeights, a, sizeof(double) * nBatch * per, cudaMemcpyHostToDevice, ctx->stream));
        double* dFrag = ctx->dWeights + nBatch * per;
        auto mmaKernel = cmg::mmaLd(lmax) == 52 ? cmg::tquBatchedMmaKernel<52> : cmg::tquBatchedMmaKernel<0>;   // lmax 44..47
        CMG_CUDA(ctx, cudaFuncSetAttribute(mmaKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smemM)));
        KernelTimer timerM(ctx);
        cmg::foldMmaWeightsKernel<<<static_cast<unsigned>(std::min<int64_t>(1024, (fragDoubles + 255) / 256)), 256, 0, ctx->stream>>>(
            ctx->dWeights, tablesOf(ctx), lmax, static_cast<int>(nBatch), dFrag);
        mmaKernel<<<dim3(static_cast<unsigned>(tilesM), static_cast<unsigned>(tilesM)), cmg::MB_THREADS, smemM, ctx->stream>>>(
            geometryOf(ctx), dFrag, tablesOf(ctx), lmax, static_cast<int>(nBatch), PM, stride);
        CMG_CUDA(ctx, cudaGetLastError());
        ctx->launches += 2;
        return timerM.finish();
    }
    cmg::PartTable P;
    std::memset(&P, 0, sizeof(P));
    P.n = 1;
    P.begin[1] = ctx->npix;
    for(int st = 0; st < 3; ++st) P.ptr[0][st] = layout.ptr[0][st];
    const int64_t tiles = (ctx->npix + cmg::PB_T - 1) / cmg::PB_T;
    const size_t smem = cmg::tquBatchedSmemBytes(lmax);
    if(smem > 227 * 1024 || tiles > 65535)
        return launchTqu(ctx, ctx->dWeights, per, lmax, nBatch, &layout, stride, nullptr);
    // folded weights live behind the raw ones in the staging buffer
    const int64_t nChunks = (nBatch + cmg::PB_BT - 1) / cmg::PB_BT;
    const int64_t foldedDoubles = nChunks * (lmax + 1) * 4 * cmg::PB_BT;
    if((s = ensureWeights(ctx, nBatch * per + foldedDoubles)) != CMG_OK) return s;
    CMG_CUDA(ctx, cudaMemcpyAsync(ctx->dWeights, a, sizeof(double) * nBatch * per, cudaMemcpyHostToDevice, ctx->stream));
    double* dFolded = ctx->dWeights + nBatch * per;
    CMG_CUDA(ctx, cudaFuncSetAttribute(cmg::tquBatchedKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    KernelTimer timer(ctx);
    cmg::foldBatchedWeightsKernel<<<static_cast<unsigned>(std::min<int64_t>(1024, (foldedDoubles + 255) / 256)), 256, 0, ctx->stream>>>(
        ctx->dWeights, tablesOf(ctx), lmax, static_cast<int>(nBatch), dFolded);
    cmg::tquBatchedKernel<<<dim3(static_cast<unsigned>(tiles), static_cast<unsigned>(tiles)), cmg::PB_T * cmg::PB_CW, smem, ctx->stream>>>(
        geometryOf(ctx), dFolded, tablesOf(ctx), lmax, static_cast<int>(nBatch), P, stride);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return timer.finish();
}

int64_t cmg_slab_doubles(int64_t dim) { return CMG_SLAB * cmg_packed_size(dim); }

static cmg_status slabGenerate(cmg_ctx* ctx, const double* a, bool weightsOnDevice, int lmax, int64_t nBatch, double* dSlabs);

cmg_status cmg_tqu_batched_slab(cmg_ctx* ctx, const double* a, int lmax, int64_t nBatch, double* dSlabs)
{
    return slabGenerate(ctx, a, false, lmax, nBatch, dSlabs);
}

cmg_status cmg_tqu_batched_slab_dev(cmg_ctx* ctx, const double* dA, int lmax, int64_t nBatch, double* dSlabs)
{
    return slabGenerate(ctx, dA, true, lmax, nBatch, dSlabs);
}

static cmg_status slabGenerate(cmg_ctx* ctx, const double* a, bool weightsOnDevice, int lmax, int64_t nBatch, double* dSlabs)
{
    cmg_status s = checkReady(ctx, lmax);
    if(s != CMG_OK) return s;
    if(!a || !dSlabs || nBatch < 1) return fail(ctx, CMG_EINVAL, "bad batch arguments");
    if(lmax < 2 || lmax > CMG_SLAB_LMAX)
        return fail(ctx, CMG_EUNSUPPORTED, "slab generation keeps the basis fragments in registers: 2 <= lmax <= 63");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t per = 4 * (lmax + 1);
    const int nkk = lmax <= 31 ? 8 : (lmax <= 47 ? 12 : 16);
    const int64_t tiles = (ctx->npix + cmg::M2_T - 1) / cmg::M2_T;
    const int64_t groups = (tiles + cmg::M2_GROUP - 1) / cmg::M2_GROUP;
    const int64_t nChunks = (nBatch + CMG_SLAB - 1) / CMG_SLAB;
    const int64_t fragDoubles = 2 * nChunks * nkk * 128;
    if(groups > 65535 || tiles * cmg::M2_GROUP > 2147483647LL)
        return fail(ctx, CMG_EUNSUPPORTED, "too many pixel tiles for one launch");
    if((s = ensureWeights(ctx, nBatch * per + fragDoubles)) != CMG_OK) return s;
    const double* dRaw = a;
    if(!weightsOnDevice)
    {
        CMG_CUDA(ctx, cudaMemcpyAsync(ctx->dWeights, a, sizeof(double) * nBatch * per, cudaMemcpyHostToDevice, ctx->stream));
        dRaw = ctx->dWeights;
    }
    double* dFrag = ctx->dWeights + nBatch * per;
    void (*kernel)(cmg::Geometry, const double*, cmg::DeviceTables, int, int, double*, long long);
    size_t smem;
    if(nkk == 8) { kernel = cmg::tquBatchedSlabKernel<8>; smem = cmg::M2Shape<8>::BYTES; }
    else if(nkk == 12) { kernel = cmg::tquBatchedSlabKernel<12>; smem = cmg::M2Shape<12>::BYTES; }
    else { kernel = cmg::tquBatchedSlabKernel<16>; smem = cmg::M2Shape<16>::BYTES; }
    CMG_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    KernelTimer timer(ctx);
    cmg::foldSlabWeightsKernel<<<static_cast<unsigned>(std::min<int64_t>(1024, (fragDoubles + 255) / 256)), 256, 0, ctx->stream>>>(
        dRaw, tablesOf(ctx), lmax, static_cast<int>(nBatch), nkk, dFrag);
    kernel<<<dim3(static_cast<unsigned>(tiles * cmg::M2_GROUP), static_cast<unsigned>(groups), 2), cmg::M2_THREADS, smem, ctx->stream>>>(
        geometryOf(ctx), dFrag, tablesOf(ctx), lmax, static_cast<int>(nBatch), dSlabs, cmg_slab_doubles(3 * ctx->npix));
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 2;
    return timer.finish();
}

cmg_status cmg_slab_unpack(cmg_ctx* ctx, const double* dSlab, int64_t dim, int nLive, int onlyB, double* dOut, int64_t outStride)
{
    if(!ctx) return CMG_EINVAL;
    if(!dSlab || !dOut || dim < 1 || nLive < 1 || nLive > CMG_SLAB || onlyB >= CMG_SLAB)
        return fail(ctx, CMG_EINVAL, "bad slab arguments");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t packed = cmg_packed_size(dim);
    const int64_t blocks = std::min<int64_t>((packed + 255) / 256, 148 * 16);
    cmg::slabUnpackKernel<<<static_cast<unsigned>(blocks), 256, 0, ctx->stream>>>(dSlab, packed, nLive, onlyB, dOut, outStride);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return CMG_OK;
}

static cmg_status clToCMatrixPolImpl(cmg_ctx* ctx, const double* ctt, const double* cte, const double* cee, const double* cbb,
                                     int lmax, double fwhm, const double* pixwinT, const double* pixwinP, double* dOut, double* outPacked)
{
    cmg_status s = checkReady(ctx, lmax);
    if(s != CMG_OK) return s;
    if(!ctt || !cte || !cee || !cbb || (!outPacked && !dOut)) return fail(ctx, CMG_EINVAL, "null argument");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n1 = lmax + 1;
    std::vector<double> fT(n1), fP(n1), a(4 * n1);
    if(cmg_window_beam(fT.data(), lmax, fwhm, pixwinT) != CMG_OK || cmg_window_beam(fP.data(), lmax, fwhm, pixwinP) != CMG_OK)
        return fail(ctx, CMG_EINVAL, "fwhm must be >= 0 (degrees)");
    if((s = cmg_tqu_weights(ctt, cte, cee, cbb, fT.data(), fP.data(), lmax, a.data(), a.data() + n1, a.data() + 2 * n1, a.data() + 3 * n1)) != CMG_OK)
        return fail(ctx, s, "bad C_l / window arguments");
    if(!dOut)
    {
        if((s = ensureScratch(ctx, sizeof(double) * cmg_packed_size(3 * ctx->npix))) != CMG_OK) return s;
        dOut = ctx->dScratch;
    }
    if(ctx->fullSky && ctx->nside >= 8 && lmax >= 2 && lmax <= cmg::PQ_STATIC_LMAX && ctx->tquVariant == 0)
    {
        // full sky: one evaluation per orbit of pixel pairs under the pi/2 rotation of the grid (orbit.cuh), a quarter of the work
        if((s = cmg_tqu_orbit(ctx, a.data(), a.data() + n1, a.data() + 2 * n1, a.data() + 3 * n1, lmax, dOut, 0)) != CMG_OK) return s;
    }
    else
    {
        cmg_tqu_layout layout;
        if((s = cmg_tqu_layout_single(ctx, dOut, &layout)) != CMG_OK) return s;
        if((s = cmg_tqu(ctx, a.data(), a.data() + n1, a.data() + 2 * n1, a.data() + 3 * n1, lmax, &layout)) != CMG_OK) return s;
    }
    return outPacked ? matrixToHost(ctx, dOut, 3, outPacked) : CMG_OK;
}

cmg_status cmg_cl_to_cmatrix_pol(cmg_ctx* ctx, const double* ctt, const double* cte, const double* cee, const double* cbb,
                                 int lmax, double fwhm, const double* pixwinT, const double* pixwinP, double* outPacked)
{
    if(!outPacked) return fail(ctx, CMG_EINVAL, "null argument");
    return clToCMatrixPolImpl(ctx, ctt, cte, cee, cbb, lmax, fwhm, pixwinT, pixwinP, nullptr, outPacked);
}

cmg_status cmg_cl_to_cmatrix_pol_dev(cmg_ctx* ctx, const double* ctt, const double* cte, const double* cee, const double* cbb,
                                     int lmax, double fwhm, const double* pixwinT, const double* pixwinP, double* dOut)
{
    if(!dOut) return fail(ctx, CMG_EINVAL, "null argument");
    return clToCMatrixPolImpl(ctx, ctt, cte, cee, cbb, lmax, fwhm, pixwinT, pixwinP, dOut, nullptr);
}

cmg_status cmg_tqu_scatter_block(cmg_ctx* ctx, const double* dBlock, int64_t col0, int64_t nCols, int64_t ld, int64_t row0, int kind,
                                 double* dFullPacked)
{
    if(!ctx) return CMG_EINVAL;
    if(ctx->npix <= 0) return fail(ctx, CMG_ESTATE, "cmg_set_pixels has not been called on this context");
    if(!dBlock || !dFullPacked || kind < 0 || kind > 2 || col0 < 0 || nCols < 0 || ld < 0 || row0 < 0 || col0 + nCols > ctx->npix ||
       row0 + ld > ctx->npix || (nCols > 0 && ld > 0 && row0 < col0 + nCols))
        return fail(ctx, CMG_EINVAL, "block outside the strictly-lower rectangle it can describe");
    if(nCols == 0 || ld == 0) return CMG_OK;
    if(nCols > 65535) return fail(ctx, CMG_EUNSUPPORTED, "more than 65535 owner columns in one block");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const dim3 grid(static_cast<unsigned>(std::min<int64_t>((ld + 255) / 256, 64)), static_cast<unsigned>(nCols));
    cmg::scatterBlockKernel<<<grid, 256, 0, ctx->stream>>>(dBlock, ctx->npix, col0, nCols, ld, row0, kind, dFullPacked);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return CMG_OK;
}

cmg_status cmg_sum_unpack_strided(cmg_ctx* ctx, const double* dC, int64_t cStride, const double* dF, const double* dN, int64_t n, double* dFull)
{
    if(!ctx || !dC || !dFull || n < 1 || cStride < 1) return CMG_EINVAL;
    if((n + 31) / 32 > 65535) return fail(ctx, CMG_EUNSUPPORTED, "matrix too large for one launch");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const unsigned blocks = static_cast<unsigned>((n + 31) / 32);
    cmg::sumUnpackKernel<<<dim3(blocks, blocks), 256, 0, ctx->stream>>>(dC, cStride, dF, dN, n, dFull);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return CMG_OK;
}

cmg_status cmg_sum_unpack(cmg_ctx* ctx, const double* dC, const double* dF, const double* dN, int64_t n, double* dFull)
{
    return cmg_sum_unpack_strided(ctx, dC, 1, dF, dN, n, dFull);
}

// ---------------------------------------------------------------- packed Cholesky (cholesky.cuh)

namespace
{
cmg_status cholBuffers(cmg_ctx* ctx)
{
    if(!ctx->dCholInfo)
        CMG_CUDA(ctx, cudaMalloc(&ctx->dCholInfo, sizeof(long long)));
    if(!ctx->dCholRed)
        CMG_CUDA(ctx, cudaMalloc(&ctx->dCholRed, sizeof(double) * (4 + cmg::CH_UKK_DOUBLES)));      // reductions, then packed U_kk + 1 / diagonal of a block
    return CMG_OK;
}
}

namespace
{
constexpr int CH_DIAG_SMEM = cmg::CH_DIAG_SMEM_DOUBLES * sizeof(double);
constexpr int CH_PANEL_SMEM = cmg::CH_PANEL_SMEM_DOUBLES * sizeof(double);
constexpr int CH_SYRK_SMEM = 2 * (cmg::CH_TILE + cmg::CH_TJ) * cmg::CH_SLD * sizeof(double);
constexpr int CH_MAX_GROUP = 4;                  // blocks per trailing update at most
constexpr int64_t CH_PLANE_SLACK = cmg::CH_TILE + cmg::CH_TJ;      // operand rows a tile may read behind the last column

cmg_status cholStepAttributes(cmg_ctx* ctx)
{
    CMG_CUDA(ctx, cudaFuncSetAttribute(cmg::cholDiagKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_DIAG_SMEM));
    CMG_CUDA(ctx, cudaFuncSetAttribute(cmg::cholPanelKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_PANEL_SMEM));
    CMG_CUDA(ctx, cudaFuncSetAttribute(cmg::cholSyrkKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SYRK_SMEM));
    return CMG_OK;
}

// one GPU, the whole triangle: a single run of the columns behind the block
cmg::CholRuns cholWholeRun(double* dA, int64_t colBegin, int64_t n)
{
    cmg::CholRuns r{};
    r.count = 1;
    r.colBegin[0] = colBegin;
    r.colEnd[0] = n;
    r.base[0] = dA;
    return r;
}

// A rank's runs clipped to the columns >= from (a multiple of 64 behind k1, or the start of a run), with the CTA prefix of a
// launch that gives `unit` columns to a CTA (panel: 128, solve update: 8); syrk: the tile counts of chSyrkTilesBefore relative
// to k1 instead (unit = 0), or one tile per 64-column block (unit < 0: the strip update).  Returns the total number of CTAs.
int64_t cholClipRuns(const cmg_chol_runs* in, int64_t from, int64_t k1, int unit, cmg::CholRuns* out)
{
    int64_t total = 0;
    out->count = 0;
    for(int r = 0; r < in->n_runs; ++r)
    {
        const int64_t c0 = std::max<int64_t>(in->col_begin[r], from), c1 = in->col_end[r];
        if(c1 <= c0)
            continue;
        const int k = out->count++;
        out->colBegin[k] = c0;
        out->colEnd[k] = c1;
        out->base[k] = in->d_run[r] - cmg::chOff(in->col_begin[r]);
        out->first[k] = total;
        if(unit > 0)
        {
            out->tile0[k] = 0;
            total += (c1 - c0 + unit - 1) / unit;
        }
        else
        {
            const int64_t b0 = (c0 - k1) / cmg::CH_TJ, b1 = (c1 - k1 + cmg::CH_TJ - 1) / cmg::CH_TJ;
            out->tile0[k] = unit == 0 ? cmg::chSyrkTilesBefore(b0) : b0;                              // unit < 0: first row tile only
            total += unit == 0 ? cmg::chSyrkTilesBefore(b1) - out->tile0[k] : b1 - b0;
        }
    }
    out->first[out->count] = total;
    return total;
}

bool cholRunsValid(const cmg_chol_runs* runs)
{
    if(!runs || runs->n_runs < 1 || runs->n_runs > CMG_CHOL_MAX_RUNS)
        return false;
    for(int r = 0; r < runs->n_runs; ++r)
    {
        if(!runs->d_run[r] || runs->col_begin[r] < 0 || runs->col_end[r] < runs->col_begin[r] || runs->col_begin[r] % cmg::CH_NB)
            return false;
        if(r > 0 && runs->col_begin[r] < runs->col_end[r - 1])
            return false;
    }
    return true;
}

// the run that holds column k0 (the owner of block k0), or -1
int cholRunOfColumn(const cmg_chol_runs* runs, int64_t k0)
{
    for(int r = 0; r < runs->n_runs; ++r)
        if(k0 >= runs->col_begin[r] && k0 < runs->col_end[r])
            return r;
    return -1;
}
}

cmg_status cmg_set_cholesky_group(cmg_ctx* ctx, int blocks)
{
    if(!ctx) return CMG_EINVAL;
    if(blocks < 0 || blocks > CH_MAX_GROUP) return fail(ctx, CMG_EINVAL, "cmg_set_cholesky_group: 1 .. 4 blocks of 128 rows, or 0 = chosen by size");
    ctx->cholGroup = blocks;
    return CMG_OK;
}

cmg_status cmg_set_cholesky_lookahead(cmg_ctx* ctx, int on)
{
    if(!ctx) return CMG_EINVAL;
    ctx->cholLookAhead = on ? 1 : 0;
    return CMG_OK;
}

// Right-looking in GROUPS of S blocks of 128 rows: inside a group every block is factorised, its 128 rows solved for all columns
// behind it (packed in place + dense plane s), and only the NEXT block's 128 rows are brought up to date (strip update, K = 128 s);
// the trailing matrix behind the group is then updated ONCE with all 128 S rows -- it is read and written n / (128 S) times
// instead of n / 128, and a tile's loads of C, its barriers and its stores are amortised over S times the DMMAs.
// Look-ahead: that update is issued in pieces -- first the S strips of rows the next group will factorise, then the rest -- and
// the next group's chain of small kernels (diagonal block, panel, strip, ...: latency-bound, a fraction of the SMs) runs on a
// high-priority side stream BESIDE the rest of the update; its rows of U go to the second set of planes.
cmg_status cmg_packed_cholesky(cmg_ctx* ctx, double* dA, int64_t n, int64_t* info)
{
    if(!ctx) return CMG_EINVAL;
    if(!dA || n < 1 || !info) return fail(ctx, CMG_EINVAL, "cmg_packed_cholesky: null argument or n < 1");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg_status s = cholBuffers(ctx);
    if(s != CMG_OK) return s;
    if((s = cholStepAttributes(ctx)) != CMG_OK) return s;
    // measured (profiles/r2_cholesky_bench_v9.log): n = 9216: 13.0 ms with groups of 2, 14.4 with 4 (the chain of small kernels
    // dominates and every block of a group adds a strip update to it); n = 36864: 536 / 510 ms
    const int S = ctx->cholGroup > 0 ? ctx->cholGroup : (n >= 16384 ? 4 : 2);
    const int64_t groupRows = static_cast<int64_t>(S) * cmg::CH_NB;
    const bool ahead = ctx->cholLookAhead && n > 2 * groupRows;
    const int64_t planeStride = (n + CH_PLANE_SLACK) * cmg::CH_NB;
    const int64_t bufStride = S * planeStride;                      // one set of planes; look-ahead alternates between two
    const size_t need = static_cast<size_t>(bufStride) * (ahead ? 2 : 1);
    if(ctx->cholPanelDoubles < need)
    {
        if(ctx->dCholPanel) CMG_CUDA(ctx, cudaFree(ctx->dCholPanel));
        ctx->dCholPanel = nullptr;
        ctx->cholPanelDoubles = 0;
        CMG_CUDA(ctx, cudaMalloc(&ctx->dCholPanel, sizeof(double) * need));
        ctx->cholPanelDoubles = need;
        CMG_CUDA(ctx, cudaMemsetAsync(ctx->dCholPanel, 0, sizeof(double) * need, ctx->stream));   // the slack rows are read
    }
    if(ahead && !ctx->cholSide)
    {
        int least = 0, greatest = 0;
        CMG_CUDA(ctx, cudaDeviceGetStreamPriorityRange(&least, &greatest));
        CMG_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->cholSide, cudaStreamNonBlocking, greatest));
        CMG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->cholEvA, cudaEventDisableTiming));
        CMG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->cholEvF, cudaEventDisableTiming));
    }
    CMG_CUDA(ctx, cudaMemsetAsync(ctx->dCholInfo, 0, sizeof(long long), ctx->stream));
    KernelTimer timer(ctx);

    auto syrk = [&](cudaStream_t st, int64_t kBase, int kb, int64_t shift, const double* planes, bool strip)
    {
        const int64_t k1 = kBase + kb + shift;
        if(k1 >= n) return;
        cmg::CholRuns runs = cholWholeRun(dA, k1, n);
        const int64_t colBlocks = (n - k1 + cmg::CH_TJ - 1) / cmg::CH_TJ;           // 64-column blocks; block b meets the row tiles 0 .. b / 2
        runs.first[1] = strip ? colBlocks : cmg::chSyrkTilesBefore(colBlocks);
        cmg::cholSyrkKernel<<<static_cast<unsigned>(runs.first[1]), cmg::CH_SYRK_THREADS, CH_SYRK_SMEM, st>>>(
            runs, kBase, kb, shift, ctx->dCholInfo, planes, kBase, planeStride, strip ? 1 : 0);
        ctx->launches += 1;
    };
    // the blocks of the group starting at kBase: true when the last block of the matrix has been factorised
    auto chain = [&](cudaStream_t st, int64_t kBase, double* planes) -> bool
    {
        for(int sub = 0; sub < S; ++sub)
        {
            const int64_t k0 = kBase + static_cast<int64_t>(sub) * cmg::CH_NB;
            if(k0 >= n) return true;
            const int kb = static_cast<int>(std::min<int64_t>(cmg::CH_NB, n - k0));
            if(sub > 0)                          // rows k0 .. k0 + 128 of every column from k0 on catch up with the sub blocks already solved
                syrk(st, kBase, sub * cmg::CH_NB, 0, planes, true);
            cmg::cholDiagKernel<<<1, 512, CH_DIAG_SMEM, st>>>(dA, k0, kb, ctx->dCholInfo, ctx->dCholRed + 4);
            ctx->launches += 1;
            const int64_t rem = n - k0 - kb;
            if(rem <= 0) return true;
            // (kb == CH_NB from here on: a short block can only be the last one)
            cmg::CholRuns runs = cholWholeRun(dA, k0 + kb, n);
            runs.first[1] = (rem + cmg::CH_PANEL_COLS - 1) / cmg::CH_PANEL_COLS;
            cmg::cholPanelKernel<<<static_cast<unsigned>(runs.first[1]), cmg::CH_PANEL_THREADS, CH_PANEL_SMEM, st>>>(
                runs, k0, kb, ctx->dCholInfo, ctx->dCholRed + 4, planes + sub * planeStride, kBase);
            ctx->launches += 1;
        }
        return kBase + groupRows >= n;
    };

    bool done = chain(ctx->stream, 0, ctx->dCholPanel);
    for(int64_t g = 0; !done; ++g)
    {
        const int64_t kBase = g * groupRows;
        double* planes = ctx->dCholPanel + (ahead ? (g & 1) * bufStride : 0);
        if(!ahead)
        {
            syrk(ctx->stream, kBase, S * cmg::CH_NB, 0, planes, false);
            done = chain(ctx->stream, kBase + groupRows, planes);
            continue;
        }
        for(int r = 0; r < S; ++r)               // the rows of the next group, strip by strip
            syrk(ctx->stream, kBase, S * cmg::CH_NB, static_cast<int64_t>(r) * cmg::CH_NB, planes, true);
        CMG_CUDA(ctx, cudaEventRecord(ctx->cholEvA, ctx->stream));
        CMG_CUDA(ctx, cudaStreamWaitEvent(ctx->cholSide, ctx->cholEvA, 0));
        done = chain(ctx->cholSide, kBase + groupRows, ctx->dCholPanel + ((g + 1) & 1) * bufStride);
        CMG_CUDA(ctx, cudaEventRecord(ctx->cholEvF, ctx->cholSide));
        syrk(ctx->stream, kBase, S * cmg::CH_NB, groupRows, planes, false);          // the rest, beside the next group's chain
        CMG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->cholEvF, 0));
    }
    CMG_CUDA(ctx, cudaGetLastError());
    if((s = timer.finish()) != CMG_OK) return s;
    long long hInfo = 0;
    CMG_CUDA(ctx, cudaMemcpyAsync(&hInfo, ctx->dCholInfo, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *info = hInfo;
    return CMG_OK;
}

// ---- the same factorisation over several GPUs, step by step (cosmopp_b200/multigpu.py: ShardedCholesky drives it) ----
// A rank holds runs of whole packed columns (the strips of cmg_orbit_shard after the exchange): cmg_chol_runs.  Block k lives in
// exactly one run of one rank (run boundaries are multiples of CH_NB).  Step k: the owner factorises the block (cmg_chol_diag)
// and everyone receives U_kk (66 KB broadcast); every rank solves the 128 rows of its own columns behind the block
// (cmg_chol_panel) into its packed columns and into a dense panel [column][128] that is zero elsewhere; one all-reduce later
// every rank holds the whole panel and updates its own columns with it (cmg_chol_syrk).

cmg_status cmg_chol_begin(cmg_ctx* ctx)
{
    if(!ctx) return CMG_EINVAL;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg_status s = cholBuffers(ctx);
    if(s != CMG_OK) return s;
    if((s = cholStepAttributes(ctx)) != CMG_OK) return s;
    CMG_CUDA(ctx, cudaMemsetAsync(ctx->dCholInfo, 0, sizeof(long long), ctx->stream));
    return CMG_OK;
}

cmg_status cmg_chol_end(cmg_ctx* ctx, int64_t* info)
{
    if(!ctx || !info) return CMG_EINVAL;
    if(!ctx->dCholInfo) return fail(ctx, CMG_EINVAL, "cmg_chol_end: no cmg_chol_begin before it");
    long long hInfo = 0;
    CMG_CUDA(ctx, cudaMemcpyAsync(&hInfo, ctx->dCholInfo, sizeof(long long), cudaMemcpyDeviceToHost, ctx->stream));
    CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *info = hInfo;
    return CMG_OK;
}

cmg_status cmg_chol_diag(cmg_ctx* ctx, const cmg_chol_runs* runs, int64_t k0, int kb, double* dUkk)
{
    if(!ctx) return CMG_EINVAL;
    if(!cholRunsValid(runs) || !dUkk || kb < 1 || kb > cmg::CH_NB || k0 % cmg::CH_NB || !ctx->dCholInfo)
        return fail(ctx, CMG_EINVAL, "cmg_chol_diag: bad arguments (after cmg_chol_begin; k0 a multiple of 128)");
    const int r = cholRunOfColumn(runs, k0);
    if(r < 0 || k0 + kb > runs->col_end[r]) return fail(ctx, CMG_EINVAL, "cmg_chol_diag: block k0 is not in this rank's columns");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    // d_ukk: kb (kb + 1) / 2 packed entries of U_kk, then the kb reciprocals of its diagonal
    cmg::cholDiagKernel<<<1, 512, CH_DIAG_SMEM, ctx->stream>>>(runs->d_run[r] - cmg::chOff(runs->col_begin[r]), k0, kb, ctx->dCholInfo, dUkk);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return CMG_OK;
}

cmg_status cmg_chol_panel(cmg_ctx* ctx, const cmg_chol_runs* runs, int64_t k0, int kb, const double* dUkk, double* dPanel, int64_t panelCol0)
{
    if(!ctx) return CMG_EINVAL;
    if(!cholRunsValid(runs) || !dUkk || !dPanel || kb != cmg::CH_NB || k0 % cmg::CH_NB || panelCol0 > k0 + kb || !ctx->dCholInfo ||
       reinterpret_cast<uintptr_t>(dPanel) % 16 || reinterpret_cast<uintptr_t>(dUkk) % 16)
        return fail(ctx, CMG_EINVAL, "cmg_chol_panel: bad arguments (a full block of 128 rows; panel_col0 <= k0 + 128; a 16-byte aligned plane)");
    cmg::CholRuns clipped;
    const int64_t ctas = cholClipRuns(runs, k0 + kb, k0 + kb, cmg::CH_PANEL_COLS, &clipped);
    if(ctas == 0) return CMG_OK;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg::cholPanelKernel<<<static_cast<unsigned>(ctas), cmg::CH_PANEL_THREADS, CH_PANEL_SMEM, ctx->stream>>>(clipped, k0, kb, ctx->dCholInfo, dUkk, dPanel,
                                                                                                           panelCol0);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return CMG_OK;
}

cmg_status cmg_chol_syrk(cmg_ctx* ctx, const cmg_chol_runs* runs, int64_t k0, int kb, int64_t shift, const double* dPanel, int64_t planeStride,
                         int64_t panelCol0, int stripOnly)
{
    if(!ctx) return CMG_EINVAL;
    if(!cholRunsValid(runs) || !dPanel || kb < cmg::CH_NB || kb > CH_MAX_GROUP * cmg::CH_NB || kb % cmg::CH_NB || k0 % cmg::CH_NB || panelCol0 > k0 + cmg::CH_NB ||
       panelCol0 % 2 || planeStride % 2 || shift < 0 || shift % cmg::CH_NB || !ctx->dCholInfo)
        return fail(ctx, CMG_EINVAL, "cmg_chol_syrk: bad arguments (kb = 128 .. 512 rows and shift in whole blocks; panel_col0 <= k0 + 128)");
    cmg::CholRuns clipped;
    const int64_t k1 = k0 + kb + shift;
    const int64_t tiles = cholClipRuns(runs, k1, k1, stripOnly ? -1 : 0, &clipped);
    if(tiles == 0) return CMG_OK;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg::cholSyrkKernel<<<static_cast<unsigned>(tiles), cmg::CH_SYRK_THREADS, CH_SYRK_SMEM, ctx->stream>>>(clipped, k0, kb, shift, ctx->dCholInfo, dPanel, panelCol0,
                                                                                                         planeStride, stripOnly ? 1 : 0);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return CMG_OK;
}

cmg_status cmg_chol_logdet_runs(cmg_ctx* ctx, const cmg_chol_runs* runs, double* logDetShare)
{
    if(!ctx) return CMG_EINVAL;
    if(!cholRunsValid(runs) || !logDetShare) return fail(ctx, CMG_EINVAL, "cmg_chol_logdet_runs: bad arguments");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg_status s = cholBuffers(ctx);
    if(s != CMG_OK) return s;
    cmg::CholRuns all;
    cholClipRuns(runs, 0, 0, cmg::CH_NB, &all);
    cmg::cholLogDetRunsKernel<<<1, 256, 0, ctx->stream>>>(all, ctx->dCholRed);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    CMG_CUDA(ctx, cudaMemcpyAsync(logDetShare, ctx->dCholRed, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CMG_OK;
}

namespace
{
void cholSolvePasses(int64_t nRhs, dim3* block, unsigned* passes)
{
    const int perPass = static_cast<int>(std::min<int64_t>(8, nRhs));
    *block = dim3(cmg::CH_NB, perPass);
    *passes = static_cast<unsigned>((nRhs + perPass - 1) / perPass);
}
}

cmg_status cmg_chol_solve_diag(cmg_ctx* ctx, const cmg_chol_runs* runs, int64_t k0, int kb, int64_t n, double* dT, int64_t nRhs)
{
    if(!ctx) return CMG_EINVAL;
    if(!cholRunsValid(runs) || !dT || kb < 1 || kb > cmg::CH_NB || k0 % cmg::CH_NB || nRhs < 1 || nRhs > 65535 || k0 + kb > n)
        return fail(ctx, CMG_EINVAL, "cmg_chol_solve_diag: bad arguments");
    const int r = cholRunOfColumn(runs, k0);
    if(r < 0 || k0 + kb > runs->col_end[r]) return fail(ctx, CMG_EINVAL, "cmg_chol_solve_diag: block k0 is not in this rank's columns");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    dim3 block;
    unsigned passes;
    cholSolvePasses(nRhs, &block, &passes);
    cmg::cholSolveDiagKernel<<<passes, block, 0, ctx->stream>>>(runs->d_run[r] - cmg::chOff(runs->col_begin[r]), k0, kb, n, dT, static_cast<int>(nRhs));
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return CMG_OK;
}

cmg_status cmg_chol_solve_update(cmg_ctx* ctx, const cmg_chol_runs* runs, int64_t k0, int kb, int64_t n, double* dT, int64_t nRhs)
{
    if(!ctx) return CMG_EINVAL;
    if(!cholRunsValid(runs) || !dT || kb < 1 || kb > cmg::CH_NB || k0 % cmg::CH_NB || nRhs < 1 || nRhs > 65535 || k0 + kb > n)
        return fail(ctx, CMG_EINVAL, "cmg_chol_solve_update: bad arguments");
    cmg::CholRuns clipped;
    const int64_t ctas = cholClipRuns(runs, k0 + kb, k0 + kb, 8, &clipped);
    if(ctas == 0) return CMG_OK;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg::cholSolveUpdateKernel<<<static_cast<unsigned>(ctas), 256, 0, ctx->stream>>>(clipped, k0, kb, n, dT, static_cast<int>(nRhs));
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return CMG_OK;
}

cmg_status cmg_packed_cholesky_logdet(cmg_ctx* ctx, const double* dU, int64_t n, double* logDet)
{
    if(!ctx) return CMG_EINVAL;
    if(!dU || n < 1 || !logDet) return fail(ctx, CMG_EINVAL, "cmg_packed_cholesky_logdet: null argument or n < 1");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg_status s = cholBuffers(ctx);
    if(s != CMG_OK) return s;
    cmg::cholLogDetKernel<<<1, 256, 0, ctx->stream>>>(dU, n, ctx->dCholRed);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    CMG_CUDA(ctx, cudaMemcpyAsync(logDet, ctx->dCholRed, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CMG_OK;
}

cmg_status cmg_packed_cholesky_solve(cmg_ctx* ctx, const double* dU, int64_t n, double* dT, int64_t nRhs)
{
    if(!ctx) return CMG_EINVAL;
    if(!dU || !dT || n < 1 || nRhs < 1 || nRhs > 65535) return fail(ctx, CMG_EINVAL, "cmg_packed_cholesky_solve: bad arguments");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int perPass = static_cast<int>(std::min<int64_t>(8, nRhs));
    const dim3 block(cmg::CH_NB, perPass);
    const unsigned passes = static_cast<unsigned>((nRhs + perPass - 1) / perPass);
    for(int64_t k0 = 0; k0 < n; k0 += cmg::CH_NB)
    {
        const int kb = static_cast<int>(std::min<int64_t>(cmg::CH_NB, n - k0));
        cmg::cholSolveDiagKernel<<<passes, block, 0, ctx->stream>>>(dU, k0, kb, n, dT, static_cast<int>(nRhs));
        ctx->launches += 1;
        const int64_t rem = n - k0 - kb;
        if(rem <= 0)
            break;
        cmg::CholRuns runs = cholWholeRun(const_cast<double*>(dU), k0 + kb, n);
        runs.first[1] = (rem + 7) / 8;
        cmg::cholSolveUpdateKernel<<<static_cast<unsigned>(runs.first[1]), 256, 0, ctx->stream>>>(runs, k0, kb, n, dT, static_cast<int>(nRhs));
        ctx->launches += 1;
    }
    CMG_CUDA(ctx, cudaGetLastError());
    return CMG_OK;
}

cmg_status cmg_packed_sum(cmg_ctx* ctx, const double* dC, int64_t cStride, const double* dF, const double* dN, int64_t n, double* dOut)
{
    if(!ctx) return CMG_EINVAL;
    if(!dC || !dOut || n < 1 || cStride < 1) return fail(ctx, CMG_EINVAL, "cmg_packed_sum: bad arguments");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t count = cmg_packed_size(n);
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((count + 255) / 256, 148 * 32));
    cmg::packedSumKernel<<<blocks, 256, 0, ctx->stream>>>(dC, cStride, dF, dN, count, dOut);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    return CMG_OK;
}

cmg_status cmg_set_like_method(cmg_ctx* ctx, int method)
{
    if(!ctx || method < 0 || method > 1) return CMG_EINVAL;
    ctx->likeMethod = method;
    return CMG_OK;
}

// ---------------------------------------------------------------- pixel likelihood on the device
// reference source/likelihood.cpp:68-134 (construct) and :136-180 (vmv, calculate).  The reference inverts C + F + N
// (LAPACK dpptrf / dpptri) and evaluates t^T C^-1 t as a double loop per map; here the Cholesky factor stays on the
// device and chi2 = |L^-1 t|^2 for all maps at once.  Default: this library's packed factorisation (cholesky.cuh), in place on
// the packed sum -- n (n + 1) / 2 doubles, no unpacked copy, any n that fits the GPU.  cmg_set_like_method(ctx, 1) selects the
// dense route instead (cusolverDnDpotrf + cublasDtrsm on the unpacked n x n matrix: plain library calls, kept as the bar to
// compare with; twice the memory, n <= 46340).

struct cmg_like
{
    cmg_ctx* ctx = nullptr;
    int64_t n = 0;
    double* dL = nullptr;            // dense method: n x n, lower triangle = Cholesky factor
    double* dU = nullptr;            // packed method: U of A = U^T U, packed upper triangle
    double* dYf = nullptr;           // L^-1 f
    double* dT = nullptr;            // maps / solutions, n x tCap
    int64_t tCap = 0;
    double* dRed = nullptr;          // 2 x tCap reductions
    bool hasF = false;
    double logDet = 0.0;             // log det (C + F + N) - offset
    double fCinvf = 0.0;
    cusolverDnHandle_t sol = nullptr;
    cublasHandle_t blas = nullptr;
};

namespace
{
const double kDetOffset = -29677.0566;          // reference source/likelihood.cpp:126

cmg_status likeReserve(cmg_like* L, int64_t nMaps)
{
    cmg_ctx* ctx = L->ctx;
    if(nMaps <= L->tCap)
        return CMG_OK;
    if(L->dT) { cudaStreamSynchronize(ctx->stream); cudaFree(L->dT); cudaFree(L->dRed); L->dT = L->dRed = nullptr; L->tCap = 0; }
    CMG_CUDA(ctx, cudaMalloc(&L->dT, sizeof(double) * L->n * nMaps));
    CMG_CUDA(ctx, cudaMalloc(&L->dRed, sizeof(double) * 2 * nMaps));
    L->tCap = nMaps;
    return CMG_OK;
}
}

void cmg_like_destroy(cmg_like* L)
{
    if(!L)
        return;
    if(L->ctx)
    {
        cudaSetDevice(L->ctx->device);
        cudaStreamSynchronize(L->ctx->stream);
    }
    if(L->dL) cudaFree(L->dL);
    if(L->dU) cudaFree(L->dU);
    if(L->dYf) cudaFree(L->dYf);
    if(L->dT) cudaFree(L->dT);
    if(L->dRed) cudaFree(L->dRed);
    if(L->sol) cusolverDnDestroy(L->sol);
    if(L->blas) cublasDestroy(L->blas);
    delete L;
}

cmg_status cmg_like_create(cmg_ctx* ctx, const double* dC, int64_t cStride, const double* dF, const double* dN, int64_t n,
                           const double* foreground, cmg_like** out)
{
    if(!ctx || !out) return CMG_EINVAL;
    *out = nullptr;
    if(!dC || n < 1 || cStride < 1) return fail(ctx, CMG_EINVAL, "bad likelihood arguments");
    if(ctx->likeMethod == 1 && n > 46340) return fail(ctx, CMG_EUNSUPPORTED, "dense factorisation limited to n <= 46340 (32-bit LAPACK-style interface)");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg_like* L = new(std::nothrow) cmg_like;
    if(!L) return fail(ctx, CMG_ENOMEM, "out of host memory");
    L->ctx = ctx;
    L->n = n;
    cmg_status s = CMG_OK;
    double* dWork = nullptr;
    int* dInfo = nullptr;
    auto bail = [&](cmg_status st) { if(dWork) cudaFree(dWork); if(dInfo) cudaFree(dInfo); cmg_like_destroy(L); return st; };
    cudaError_t e;
    if(ctx->likeMethod == 0)
    {
        // packed: C + F + N in one pass, factorised where it lies
        if((e = cudaMalloc(&L->dU, sizeof(double) * cmg_packed_size(n))) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMalloc (packed sum)"));
        if((s = cmg_packed_sum(ctx, dC, cStride, dF, dN, n, L->dU)) != CMG_OK) return bail(s);
        int64_t info = 0;
        if((s = cmg_packed_cholesky(ctx, L->dU, n, &info)) != CMG_OK) return bail(s);
        if(info != 0)
            return bail(fail(ctx, CMG_ENUMERIC, "The determinant of the covariance matrix is not positive. The covariance matrix must be positive definite."));
        double logDet = 0.0;
        if((s = cmg_packed_cholesky_logdet(ctx, L->dU, n, &logDet)) != CMG_OK) return bail(s);
        L->logDet = logDet - kDetOffset;
        if(foreground)
        {
            if((e = cudaMalloc(&L->dYf, sizeof(double) * n)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMalloc"));
            if((e = cudaMemcpyAsync(L->dYf, foreground, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMemcpy"));
            if((s = cmg_packed_cholesky_solve(ctx, L->dU, n, L->dYf, 1)) != CMG_OK) return bail(s);
            if((s = likeReserve(L, 1)) != CMG_OK) return bail(s);
            cmg::columnDotsKernel<<<1, 256, 0, ctx->stream>>>(L->dYf, nullptr, n, L->dRed, nullptr);
            if((e = cudaMemcpyAsync(&L->fCinvf, L->dRed, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMemcpy"));
            if((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaStreamSynchronize"));
            ctx->launches += 1;
            L->hasF = true;
        }
        *out = L;
        return CMG_OK;
    }
    if((e = cudaMalloc(&L->dL, sizeof(double) * n * n)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMalloc (full matrix)"));
    if((s = cmg_sum_unpack_strided(ctx, dC, cStride, dF, dN, n, L->dL)) != CMG_OK) return bail(s);
    if(cusolverDnCreate(&L->sol) != CUSOLVER_STATUS_SUCCESS || cublasCreate(&L->blas) != CUBLAS_STATUS_SUCCESS)
        return bail(fail(ctx, CMG_ECUDA, "cuSOLVER / cuBLAS handle creation failed"));
    cusolverDnSetStream(L->sol, ctx->stream);
    cublasSetStream(L->blas, ctx->stream);
    int lwork = 0;
    if(cusolverDnDpotrf_bufferSize(L->sol, CUBLAS_FILL_MODE_LOWER, static_cast<int>(n), L->dL, static_cast<int>(n), &lwork) != CUSOLVER_STATUS_SUCCESS)
        return bail(fail(ctx, CMG_ECUDA, "cusolverDnDpotrf_bufferSize failed"));
    if((e = cudaMalloc(&dWork, sizeof(double) * std::max(lwork, 1))) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMalloc (potrf workspace)"));
    if((e = cudaMalloc(&dInfo, sizeof(int))) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMalloc"));
    if(cusolverDnDpotrf(L->sol, CUBLAS_FILL_MODE_LOWER, static_cast<int>(n), L->dL, static_cast<int>(n), dWork, lwork, dInfo) != CUSOLVER_STATUS_SUCCESS)
        return bail(fail(ctx, CMG_ECUDA, "cusolverDnDpotrf failed"));
    int info = 0;
    std::vector<double> diag(n);
    if((e = cudaMemcpyAsync(&info, dInfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMemcpy"));
    if((e = cudaMemcpy2DAsync(diag.data(), sizeof(double), L->dL, sizeof(double) * (n + 1), sizeof(double), n, cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess)
        return bail(cudaFail(ctx, e, "cudaMemcpy2D (diagonal)"));
    if((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaStreamSynchronize"));
    if(info != 0)
        return bail(fail(ctx, CMG_ENUMERIC, "The determinant of the covariance matrix is not positive. The covariance matrix must be positive definite."));
    double logDet = 0.0;
    for(int64_t i = 0; i < n; ++i)
        logDet += std::log(diag[i]);
    L->logDet = 2.0 * logDet - kDetOffset;
    if(foreground)
    {
        if((e = cudaMalloc(&L->dYf, sizeof(double) * n)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMalloc"));
        if((e = cudaMemcpyAsync(L->dYf, foreground, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMemcpy"));
        const double one = 1.0;
        if(cublasDtrsm(L->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, static_cast<int>(n), 1, &one,
                       L->dL, static_cast<int>(n), L->dYf, static_cast<int>(n)) != CUBLAS_STATUS_SUCCESS)
            return bail(fail(ctx, CMG_ECUDA, "cublasDtrsm failed"));
        if((s = likeReserve(L, 1)) != CMG_OK) return bail(s);
        cmg::columnDotsKernel<<<1, 256, 0, ctx->stream>>>(L->dYf, nullptr, n, L->dRed, nullptr);
        if((e = cudaMemcpyAsync(&L->fCinvf, L->dRed, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMemcpy"));
        if((e = cudaStreamSynchronize(ctx->stream)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaStreamSynchronize"));
        ctx->launches += 1;
        L->hasF = true;
    }
    cudaFree(dWork);
    cudaFree(dInfo);
    *out = L;
    return CMG_OK;
}

// reference source/likelihood.cpp:341-406 (LikelihoodPolarization's constructor): cInv = N^-1 + N^-1 C N^-1, Cholesky, log det,
// inverse.  Here: the two products as plain library GEMMs on the unpacked matrices (cublasDgemm), the packed upper triangle of
// the result factorised by cmg_packed_cholesky; chi^2 = v^T cInv^-1 v = |U^-T v|^2 comes from cmg_like_calculate.
cmg_status cmg_like_create_ninv(cmg_ctx* ctx, const double* dC, const double* dNinv, int64_t m, double detOffset, cmg_like** out)
{
    if(!ctx || !out) return CMG_EINVAL;
    *out = nullptr;
    if(!dC || !dNinv || m < 1 || m > 46340) return fail(ctx, CMG_EINVAL, "cmg_like_create_ninv: null matrix or dimension outside 1 .. 46340");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg_like* L = new(std::nothrow) cmg_like;
    if(!L) return fail(ctx, CMG_ENOMEM, "out of host memory");
    L->ctx = ctx;
    L->n = m;
    double *fC = nullptr, *fN = nullptr, *fT = nullptr;
    cublasHandle_t blas = nullptr;
    auto bail = [&](cmg_status st)
    {
        if(fC) cudaFree(fC);
        if(fN) cudaFree(fN);
        if(fT) cudaFree(fT);
        if(blas) cublasDestroy(blas);
        cmg_like_destroy(L);
        return st;
    };
    cudaError_t e;
    const size_t dense = sizeof(double) * static_cast<size_t>(m) * static_cast<size_t>(m);
    if((e = cudaMalloc(&fC, dense)) != cudaSuccess || (e = cudaMalloc(&fN, dense)) != cudaSuccess || (e = cudaMalloc(&fT, dense)) != cudaSuccess)
        return bail(cudaFail(ctx, e, "cudaMalloc (dense work matrices)"));
    if((e = cudaMalloc(&L->dU, sizeof(double) * cmg_packed_size(m))) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMalloc (packed factor)"));
    cmg_status s;
    if((s = cmg_sum_unpack(ctx, dC, nullptr, nullptr, m, fC)) != CMG_OK) return bail(s);
    if((s = cmg_sum_unpack(ctx, dNinv, nullptr, nullptr, m, fN)) != CMG_OK) return bail(s);
    if(cublasCreate(&blas) != CUBLAS_STATUS_SUCCESS) return bail(fail(ctx, CMG_ECUDA, "cublasCreate failed"));
    cublasSetStream(blas, ctx->stream);
    const double one = 1.0, zero = 0.0;
    const int mi = static_cast<int>(m);
    // T = C N^-1;  K = N^-1 T + N^-1 (accumulated onto a copy of N^-1 held in fC afterwards)
    if(cublasDgemm(blas, CUBLAS_OP_N, CUBLAS_OP_N, mi, mi, mi, &one, fC, mi, fN, mi, &zero, fT, mi) != CUBLAS_STATUS_SUCCESS)
        return bail(fail(ctx, CMG_ECUDA, "cublasDgemm failed"));
    if((e = cudaMemcpyAsync(fC, fN, dense, cudaMemcpyDeviceToDevice, ctx->stream)) != cudaSuccess) return bail(cudaFail(ctx, e, "cudaMemcpy"));
    if(cublasDgemm(blas, CUBLAS_OP_N, CUBLAS_OP_N, mi, mi, mi, &one, fN, mi, fT, mi, &one, fC, mi) != CUBLAS_STATUS_SUCCESS)
        return bail(fail(ctx, CMG_ECUDA, "cublasDgemm failed"));
    cmg::packUpperKernel<<<dim3(static_cast<unsigned>(std::min<int64_t>((m + 255) / 256, 64)), static_cast<unsigned>(m)), 256, 0, ctx->stream>>>(fC, m, L->dU);
    if((e = cudaGetLastError()) != cudaSuccess) return bail(cudaFail(ctx, e, "packUpperKernel"));
    ctx->launches += 1;
    int64_t info = 0;
    if((s = cmg_packed_cholesky(ctx, L->dU, m, &info)) != CMG_OK) return bail(s);
    if(info != 0)
        return bail(fail(ctx, CMG_ENUMERIC, "The determinant of the covariance matrix is not positive. The covariance matrix must be positive definite."));
    double logDet = 0.0;
    if((s = cmg_packed_cholesky_logdet(ctx, L->dU, m, &logDet)) != CMG_OK) return bail(s);
    L->logDet = logDet - detOffset;
    cudaFree(fC);
    cudaFree(fN);
    cudaFree(fT);
    cublasDestroy(blas);
    *out = L;
    return CMG_OK;
}

cmg_status cmg_like_calculate(cmg_like* L, const double* t, int64_t nMaps, double* chi2, double* logDet)
{
    if(!L || !L->ctx) return CMG_EINVAL;
    cmg_ctx* ctx = L->ctx;
    if(!t || !chi2 || nMaps < 1 || nMaps > 2147483647LL) return fail(ctx, CMG_EINVAL, "bad likelihood arguments");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg_status s = likeReserve(L, nMaps);
    if(s != CMG_OK) return s;
    const int64_t n = L->n;
    CMG_CUDA(ctx, cudaMemcpyAsync(L->dT, t, sizeof(double) * n * nMaps, cudaMemcpyHostToDevice, ctx->stream));
    const double one = 1.0;
    if(L->dU)
    {
        if((s = cmg_packed_cholesky_solve(ctx, L->dU, n, L->dT, nMaps)) != CMG_OK) return s;
    }
    else if(cublasDtrsm(L->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, static_cast<int>(n), static_cast<int>(nMaps), &one,
                        L->dL, static_cast<int>(n), L->dT, static_cast<int>(n)) != CUBLAS_STATUS_SUCCESS)
        return fail(ctx, CMG_ECUDA, "cublasDtrsm failed");
    cmg::columnDotsKernel<<<static_cast<unsigned>(nMaps), 256, 0, ctx->stream>>>(L->dT, L->hasF ? L->dYf : nullptr, n, L->dRed, L->hasF ? L->dRed + nMaps : nullptr);
    CMG_CUDA(ctx, cudaGetLastError());
    ctx->launches += 1;
    std::vector<double> red(2 * nMaps, 0.0);
    CMG_CUDA(ctx, cudaMemcpyAsync(red.data(), L->dRed, sizeof(double) * (L->hasF ? 2 : 1) * nMaps, cudaMemcpyDeviceToHost, ctx->stream));
    CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    double ld = L->logDet;
    if(L->hasF)
        ld += std::log(L->fCinvf / static_cast<double>(n));                 // source/likelihood.cpp:174
    for(int64_t k = 0; k < nMaps; ++k)
    {
        chi2[k] = red[k];
        if(L->hasF)
            chi2[k] -= red[nMaps + k] * red[nMaps + k] / L->fCinvf;          // :175
    }
    if(logDet) *logDet = ld;
    return CMG_OK;
}

// ---------------------------------------------------------------- CMatrix files straight from / to device memory
// binary layout of reference source/c_matrix.cpp:41-104: int32 nPix, nPix (nPix + 1) / 2 doubles, int32 length, comment.
// Pieces (a rank's packed strip, a sub-range) are written at their element offset with pwrite, through two pinned
// bounce buffers so that the device-to-host copy of piece k+1 overlaps the file write of piece k.

struct cmg_file
{
    cmg_ctx* ctx = nullptr;
    int fd = -1;
    int64_t nPix = 0;
    int64_t total = 0;              // packed doubles
    bool writing = false;
    double* bounce[2] = {nullptr, nullptr};
    cudaEvent_t done[2] = {nullptr, nullptr};
    int64_t bounceDoubles = 0;
};

namespace
{
const int64_t kBounceDoubles = int64_t(1) << 22;      // 32 MiB per buffer

cmg_status fileBuffers(cmg_file* f)
{
    cmg_ctx* ctx = f->ctx;
    for(int k = 0; k < 2; ++k)
    {
        CMG_CUDA(ctx, cudaMallocHost(reinterpret_cast<void**>(&f->bounce[k]), sizeof(double) * kBounceDoubles));
        CMG_CUDA(ctx, cudaEventCreateWithFlags(&f->done[k], cudaEventDisableTiming));
    }
    f->bounceDoubles = kBounceDoubles;
    return CMG_OK;
}

void fileRelease(cmg_file* f)
{
    if(!f) return;
    for(int k = 0; k < 2; ++k)
    {
        if(f->bounce[k]) cudaFreeHost(f->bounce[k]);
        if(f->done[k]) cudaEventDestroy(f->done[k]);
    }
    if(f->fd >= 0) ::close(f->fd);
    delete f;
}

bool pwriteAll(int fd, const void* buf, size_t bytes, off_t off)
{
    const char* p = static_cast<const char*>(buf);
    while(bytes > 0)
    {
        const ssize_t w = ::pwrite(fd, p, bytes, off);
        if(w <= 0) return false;
        p += w; off += w; bytes -= static_cast<size_t>(w);
    }
    return true;
}

bool preadAll(int fd, void* buf, size_t bytes, off_t off)
{
    char* p = static_cast<char*>(buf);
    while(bytes > 0)
    {
        const ssize_t r = ::pread(fd, p, bytes, off);
        if(r <= 0) return false;
        p += r; off += r; bytes -= static_cast<size_t>(r);
    }
    return true;
}
}

cmg_status cmg_cmatrix_file_create(cmg_ctx* ctx, const char* path, int64_t nPix, cmg_file** out)
{
    if(!ctx || !out) return CMG_EINVAL;
    *out = nullptr;
    if(!path || nPix < 1 || nPix > 2147483647LL) return fail(ctx, CMG_EINVAL, "bad file arguments (the header stores nPix as int32)");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg_file* f = new(std::nothrow) cmg_file;
    if(!f) return fail(ctx, CMG_ENOMEM, "out of host memory");
    f->ctx = ctx; f->nPix = nPix; f->total = cmg_packed_size(nPix); f->writing = true;
    f->fd = ::open(path, O_CREAT | O_TRUNC | O_WRONLY, 0644);
    const int32_t n32 = static_cast<int32_t>(nPix);
    if(f->fd < 0 || !pwriteAll(f->fd, &n32, sizeof(n32), 0) || ::ftruncate(f->fd, static_cast<off_t>(4 + 8 * f->total)) != 0)
    {
        fileRelease(f);
        return fail(ctx, CMG_EINVAL, std::string("Cannot write into output file ") + path + ".");      // text of source/c_matrix.cpp:86
    }
    const cmg_status s = fileBuffers(f);
    if(s != CMG_OK) { fileRelease(f); return s; }
    *out = f;
    return CMG_OK;
}

cmg_status cmg_cmatrix_file_open(cmg_ctx* ctx, const char* path, int64_t* nPix, cmg_file** out)
{
    if(!ctx || !out) return CMG_EINVAL;
    *out = nullptr;
    if(!path) return fail(ctx, CMG_EINVAL, "null path");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cmg_file* f = new(std::nothrow) cmg_file;
    if(!f) return fail(ctx, CMG_ENOMEM, "out of host memory");
    f->ctx = ctx;
    f->fd = ::open(path, O_RDONLY);
    int32_t n32 = 0;
    struct stat st;
    if(f->fd < 0 || !preadAll(f->fd, &n32, sizeof(n32), 0) || n32 < 1 || ::fstat(f->fd, &st) != 0 ||
       st.st_size < static_cast<off_t>(4 + 8 * cmg_packed_size(n32) + 4))
    {
        fileRelease(f);
        return fail(ctx, CMG_EINVAL, std::string("Covariance matrix file ") + path + " cannot be read.");  // source/c_matrix.cpp:47-54
    }
    f->nPix = n32; f->total = cmg_packed_size(n32);
    const cmg_status s = fileBuffers(f);
    if(s != CMG_OK) { fileRelease(f); return s; }
    if(nPix) *nPix = n32;
    *out = f;
    return CMG_OK;
}

cmg_status cmg_cmatrix_file_write_device(cmg_file* f, int64_t firstElement, const double* dSrc, int64_t count)
{
    if(!f || !f->ctx) return CMG_EINVAL;
    cmg_ctx* ctx = f->ctx;
    if(!f->writing || !dSrc || firstElement < 0 || count < 0 || firstElement + count > f->total)
        return fail(ctx, CMG_EINVAL, "piece outside the packed triangle (or file opened for reading)");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nPieces = (count + f->bounceDoubles - 1) / f->bounceDoubles;
    auto pieceLen = [&](int64_t k) { return std::min(f->bounceDoubles, count - k * f->bounceDoubles); };
    if(nPieces > 0)
    {
        CMG_CUDA(ctx, cudaMemcpyAsync(f->bounce[0], dSrc, sizeof(double) * pieceLen(0), cudaMemcpyDeviceToHost, ctx->stream));
        CMG_CUDA(ctx, cudaEventRecord(f->done[0], ctx->stream));
    }
    for(int64_t k = 0; k < nPieces; ++k)
    {
        if(k + 1 < nPieces)
        {
            CMG_CUDA(ctx, cudaMemcpyAsync(f->bounce[(k + 1) & 1], dSrc + (k + 1) * f->bounceDoubles, sizeof(double) * pieceLen(k + 1),
                                          cudaMemcpyDeviceToHost, ctx->stream));
            CMG_CUDA(ctx, cudaEventRecord(f->done[(k + 1) & 1], ctx->stream));
        }
        CMG_CUDA(ctx, cudaEventSynchronize(f->done[k & 1]));
        if(!pwriteAll(f->fd, f->bounce[k & 1], sizeof(double) * pieceLen(k), static_cast<off_t>(4 + 8 * (firstElement + k * f->bounceDoubles))))
            return fail(ctx, CMG_EINVAL, "write to the covariance matrix file failed");
    }
    return CMG_OK;
}

cmg_status cmg_cmatrix_file_read_device(cmg_file* f, int64_t firstElement, double* dDst, int64_t count)
{
    if(!f || !f->ctx) return CMG_EINVAL;
    cmg_ctx* ctx = f->ctx;
    if(f->writing || !dDst || firstElement < 0 || count < 0 || firstElement + count > f->total)
        return fail(ctx, CMG_EINVAL, "piece outside the packed triangle (or file opened for writing)");
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t nPieces = (count + f->bounceDoubles - 1) / f->bounceDoubles;
    for(int64_t k = 0; k < nPieces; ++k)
    {
        const int64_t len = std::min(f->bounceDoubles, count - k * f->bounceDoubles);
        if(k >= 2)
            CMG_CUDA(ctx, cudaEventSynchronize(f->done[k & 1]));                 // the copy out of this buffer has finished
        if(!preadAll(f->fd, f->bounce[k & 1], sizeof(double) * len, static_cast<off_t>(4 + 8 * (firstElement + k * f->bounceDoubles))))
            return fail(ctx, CMG_EINVAL, "read from the covariance matrix file failed");
        CMG_CUDA(ctx, cudaMemcpyAsync(dDst + k * f->bounceDoubles, f->bounce[k & 1], sizeof(double) * len, cudaMemcpyHostToDevice, ctx->stream));
        CMG_CUDA(ctx, cudaEventRecord(f->done[k & 1], ctx->stream));
    }
    CMG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return CMG_OK;
}

cmg_status cmg_cmatrix_file_comment(cmg_file* f, char* buffer, int64_t capacity)
{
    if(!f || !f->ctx) return CMG_EINVAL;
    cmg_ctx* ctx = f->ctx;
    if(f->writing || !buffer || capacity < 1) return fail(ctx, CMG_EINVAL, "bad comment buffer");
    int32_t len = 0;
    const off_t at = static_cast<off_t>(4 + 8 * f->total);
    if(!preadAll(f->fd, &len, sizeof(len), at) || len < 0)
        return fail(ctx, CMG_EINVAL, "cannot read the comment");
    const int64_t take = std::min<int64_t>(len, capacity - 1);
    if(take > 0 && !preadAll(f->fd, buffer, static_cast<size_t>(take), at + 4))
        return fail(ctx, CMG_EINVAL, "cannot read the comment");
    buffer[take] = 0;
    return CMG_OK;
}

cmg_status cmg_cmatrix_file_close(cmg_file* f, const char* comment)
{
    if(!f) return CMG_EINVAL;
    cmg_ctx* ctx = f->ctx;
    cmg_status s = CMG_OK;
    if(f->writing)
    {
        const std::string c = comment ? comment : "";
        const int32_t len = static_cast<int32_t>(c.size());
        const off_t at = static_cast<off_t>(4 + 8 * f->total);
        if(!pwriteAll(f->fd, &len, sizeof(len), at) || (len > 0 && !pwriteAll(f->fd, c.data(), c.size(), at + 4)))
            s = fail(ctx, CMG_EINVAL, "write to the covariance matrix file failed");
    }
    fileRelease(f);
    return s;
}

// ---------------------------------------------------------------- measurement

cmg_status cmg_measure_fp64_peak(cmg_ctx* ctx, double* tflops)
{
    if(!ctx || !tflops) return CMG_EINVAL;
    CMG_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    CMG_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    double* sink = nullptr;
    CMG_CUDA(ctx, cudaMalloc(&sink, sizeof(double)));
    const int blocks = prop.multiProcessorCount * 16;
    double best = 0.0;
    for(int rep = 0; rep < 6; ++rep)
    {
        CMG_CUDA(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        for(int k = 0; k < 4; ++k)
            cmg::fp64PeakKernel<<<blocks, 256, 0, ctx->stream>>>(sink, 1.0000001, 0.9999999);
        CMG_CUDA(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        CMG_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        float ms = 0;
        CMG_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        const double flop = 4.0 * blocks * 256.0 * cmg::PEAK_CHAINS * cmg::PEAK_ITERS * 2.0;
        if(rep > 0)
            best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    CMG_CUDA(ctx, cudaFree(sink));
    *tflops = best;
    return CMG_OK;
}

cmg_status cmg_last_kernel_ms(cmg_ctx* ctx, double* ms)
{
    if(!ctx || !ms) return CMG_EINVAL;
    *ms = ctx->lastMs;
    return CMG_OK;
}

cmg_status cmg_set_kernel_variant(cmg_ctx* ctx, int variant)
{
    if(!ctx || variant < 0) return CMG_EINVAL;   // (1 also selects the shared-memory-table TT kernel)
    ctx->tquVariant = variant;
    return CMG_OK;
}

cmg_status cmg_set_host_expand(cmg_ctx* ctx, int threads)
{
    if(!ctx || threads < -1) return CMG_EINVAL;
    ctx->hostExpandThreads = threads;
    return CMG_OK;
}

cmg_status cmg_set_host_expand_direct(cmg_ctx* ctx, int directMask)
{
    if(!ctx || directMask < 0 || directMask > 511) return CMG_EINVAL;
    ctx->hostExpandDirectMask = directMask;
    return CMG_OK;
}

cmg_status cmg_set_timing(cmg_ctx* ctx, int enabled)
{
    if(!ctx) return CMG_EINVAL;
    ctx->timing = enabled != 0;
    return CMG_OK;
}

} // extern "C"
