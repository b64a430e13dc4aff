// Quick parity + timing probe of the experimental symmetry-orbit path (cmg_tqu_orbit) against cmg_tqu, through the C ABI
// only (no Python: starts in a second on a fresh box).
//   nvcc -O2 -std=c++17 -I include -o tools/bin/orbit_check tools/orbit_check.cu -L cosmopp_b200/lib -lcosmopp_b200 \
//        -Xlinker -rpath='$ORIGIN/../../cosmopp_b200/lib'
//   tools/bin/orbit_check [full | tt | ranks | prof]          (a call costs ~20 s of GPU box time: no Python start-up)
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "cmg.h"

#define OK(call)                                                                              \
    do                                                                                        \
    {                                                                                         \
        const cmg_status s_ = (call);                                                         \
        if(s_ != CMG_OK)                                                                      \
        {                                                                                     \
            std::printf("FAIL %s -> %d: %s\n", #call, (int) s_, cmg_last_error(ctx));          \
            return 1;                                                                         \
        }                                                                                     \
    } while(0)

static void weights(int lmax, std::vector<double>& tt, std::vector<double>& te, std::vector<double>& ee, std::vector<double>& bb)
{
    tt.assign(lmax + 1, 0.0); te = tt; ee = tt; bb = tt;
    unsigned s = 12345u;
    auto u = [&]() { s = s * 1664525u + 1013904223u; return 0.5 + (s >> 8) / 16777216.0; };
    for(int l = 2; l <= lmax; ++l)
    {
        const double w = (2 * l + 1) / (4 * 3.141592653589793) * std::exp(-l * (l + 1) * 0.0055);
        const double ctt = 1000.0 * u() / (l * (l + 1.0));
        const double cee = 0.03 * ctt * u(), cbb = 0.002 * ctt * u();
        tt[l] = w * ctt; ee[l] = w * cee; bb[l] = w * cbb; te[l] = w * (u() - 1.0) * std::sqrt(ctt * cee);
    }
}

// a rank's shard over freshly allocated strips + compact outbox (cmg_orbit_shard, include/cmg.h); mode fixes the outbox size
static int makeShard(cmg_ctx* ctx, int nside, int mode, const std::vector<int64_t>& b, int r, cmg_orbit_shard& sh, double*& dStrips, double*& dBox,
                     int64_t& stripDoubles, int64_t& boxDoubles, bool poison)
{
    const int64_t F = (int64_t) nside * nside, n = 12 * F;
    const int world = (int) b.size() - 1;
    std::memset(&sh, 0, sizeof(sh));
    sh.n_ranks = world;
    sh.rank = r;
    for(int k = 0; k <= world; ++k) sh.bounds[k] = b[k];
    stripDoubles = 0;
    for(int st = 0; st < 3; ++st)
        for(int f = 0; f < 12; ++f)
            stripDoubles += cmg_packed_size(st * n + f * F + b[r + 1]) - cmg_packed_size(st * n + f * F + b[r]);
    int64_t off[CMG_MAX_PARTS + 1];
    if(cmg_orbit_outbox_layout(nside, mode, world, b.data(), r, off) != CMG_OK) return 1;
    boxDoubles = off[world];
    int64_t other[CMG_MAX_PARTS + 1];             // callers may run several modes over one shard: size the buffer for the larger layout
    if(cmg_orbit_outbox_layout(nside, mode == 1 ? 0 : 1, world, b.data(), r, other) != CMG_OK) return 1;
    const int64_t boxAlloc = std::max(boxDoubles, other[world]);
    dStrips = dBox = nullptr;
    if(cmg_device_malloc(ctx, std::max<int64_t>(stripDoubles, 1) * 8, (void**) &dStrips) != CMG_OK) return 1;
    if(cmg_device_malloc(ctx, std::max<int64_t>(boxAlloc, 1) * 8, (void**) &dBox) != CMG_OK) return 1;
    if(poison)
    {
        cudaMemset(dStrips, 0xFF, stripDoubles * 8);
        cudaMemset(dBox, 0xFF, std::max<int64_t>(boxAlloc, 1) * 8);
        cudaDeviceSynchronize();
    }
    int64_t at = 0;
    for(int st = 0; st < 3; ++st)
        for(int f = 0; f < 12; ++f)
        {
            sh.strip[st][f] = dStrips + at;
            at += cmg_packed_size(st * n + f * F + b[r + 1]) - cmg_packed_size(st * n + f * F + b[r]);
        }
    sh.outbox = boxAlloc ? dBox : nullptr;
    return 0;
}

int main(int argc, char** argv)
{
    // modes: tt = parity + timing of the TT orbit kernel; full (default) = all parity checks + timings; ranks = per-rank timings of the balanced 2/4/8-way partitions at
    // Nside=64; prof = one cmg_tqu_orbit call at Nside=64 (what ncu profiles)
    const std::string what = argc > 1 ? argv[1] : "full";
    const int timingNside = 64;
    cmg_ctx* ctx = nullptr;
    if(cmg_create(&ctx, 0) != CMG_OK) { std::printf("no context: %s\n", cmg_last_error(nullptr)); return 1; }
    OK(cmg_set_timing(ctx, 1));
    int rc = 0;

    if(what == "tt")      // cmg_legendre_series_orbit (written without GPU time left: run this first in the next round)
    {
        const int cases[][2] = {{16, 47}, {32, 96}, {64, 192}};
        for(const auto& cs : cases)
        {
            const int nside = cs[0], lmax = cs[1];
            OK(cmg_set_pixels(ctx, nside, nullptr, 0));
            const int64_t n = cmg_npix(ctx), packed = cmg_packed_size(n);
            std::vector<double> tt, te, ee, bb;
            weights(lmax, tt, te, ee, bb);
            double *dA = nullptr, *dB = nullptr;
            OK(cmg_device_malloc(ctx, packed * 8, (void**) &dA));
            OK(cmg_device_malloc(ctx, packed * 8, (void**) &dB));
            double msBase = 1e30, ms = 1e30, t = 0;
            for(int rep = 0; rep < 3; ++rep)
            {
                OK(cmg_legendre_series(ctx, tt.data(), lmax, 0, n, dA));
                OK(cmg_last_kernel_ms(ctx, &t)); msBase = std::min(msBase, t);
            }
            cudaMemset(dB, 0xFF, packed * 8);
            cudaDeviceSynchronize();
            for(int rep = 0; rep < 3; ++rep)
            {
                OK(cmg_legendre_series_orbit(ctx, tt.data(), lmax, dB));
                OK(cmg_last_kernel_ms(ctx, &t)); ms = std::min(ms, t);
            }
            // whole matrix up to Nside=32; at Nside=64 the last 64 MB of the triangle (columns of the last face: every class writes there)
            const int64_t cmp = nside <= 32 ? packed : (int64_t) 8 << 20, first = packed - cmp;
            std::vector<double> hA(cmp), hB(cmp);
            double diag = 0;
            OK(cmg_copy_to_host(ctx, &diag, dA, 8));
            OK(cmg_copy_to_host(ctx, hA.data(), dA + first, cmp * 8));
            OK(cmg_copy_to_host(ctx, hB.data(), dB + first, cmp * 8));
            OK(cmg_synchronize(ctx));
            int64_t nan = 0; double worst = 0;
            for(int64_t e = 0; e < cmp; ++e)
            {
                if(std::isnan(hB[e])) { ++nan; continue; }
                worst = std::max(worst, std::fabs(hB[e] - hA[e]) / diag);
            }
            std::printf("TT nside %d lmax %d: unwritten %lld, max |orbit - every pair| / diag = %.3e; %.3f ms (every pair %.3f ms)\n",
                        nside, lmax, (long long) nan, worst, ms, msBase);
            if(nan || worst > 1e-11) rc = 1;
            OK(cmg_device_free(ctx, dA));
            OK(cmg_device_free(ctx, dB));
        }
        cmg_destroy(ctx);
        std::printf(rc ? "TT ORBIT CHECK FAILED\n" : "TT ORBIT CHECK OK\n");
        return rc;
    }

    if(what == "time")    // Nside 64: modes 3 and 0 on one GPU, best of 3
    {
        const int nside = 64, lmax = 192;
        OK(cmg_set_pixels(ctx, nside, nullptr, 0));
        const int64_t n = cmg_npix(ctx), packed = cmg_packed_size(3 * n);
        std::vector<double> tt, te, ee, bb;
        weights(lmax, tt, te, ee, bb);
        double* dA = nullptr;
        OK(cmg_device_malloc(ctx, packed * 8, (void**) &dA));
        for(int mode = 3; mode >= 0; mode -= 3)
        {
            double best = 1e30;
            for(int rep = 0; rep < 3; ++rep)
            {
                OK(cmg_tqu_orbit(ctx, tt.data(), te.data(), ee.data(), bb.data(), lmax, dA, mode));
                double ms = 0; OK(cmg_last_kernel_ms(ctx, &ms));
                best = std::min(best, ms);
            }
            std::printf("nside 64 lmax 192 mode %d: %.2f ms\n", mode, best);
        }
        cmg_destroy(ctx);
        return 0;
    }
    if(what == "prof" || what == "ranks")
    {
        const int nside = 64, lmax = 192;
        OK(cmg_set_pixels(ctx, nside, nullptr, 0));
        const int64_t n = cmg_npix(ctx), F = (int64_t) nside * nside, packed = cmg_packed_size(3 * n);
        std::vector<double> tt, te, ee, bb;
        weights(lmax, tt, te, ee, bb);
        if(what == "prof")
        {
            double* dA = nullptr;
            OK(cmg_device_malloc(ctx, packed * 8, (void**) &dA));
            const int profMode = argc > 2 ? std::atoi(argv[2]) : 0;
            OK(cmg_tqu_orbit(ctx, tt.data(), te.data(), ee.data(), bb.data(), lmax, dA, profMode));
            double ms = 0; OK(cmg_last_kernel_ms(ctx, &ms));
            std::printf("prof: nside 64 lmax 192 mode %d: %.2f ms\n", profMode, ms);
            cmg_destroy(ctx);
            return 0;
        }
        {
            double* dA = nullptr;
            OK(cmg_device_malloc(ctx, packed * 8, (void**) &dA));
            for(int mode = 3; mode >= 0; --mode)
            {
                double best = 1e30;
                for(int rep = 0; rep < 4; ++rep)
                {
                    OK(cmg_tqu_orbit(ctx, tt.data(), te.data(), ee.data(), bb.data(), lmax, dA, mode));
                    double ms = 0; OK(cmg_last_kernel_ms(ctx, &ms));
                    best = std::min(best, ms);
                }
                std::printf("nside 64 lmax 192, one rank, mode %d: %.2f ms\n", mode, best);
            }
            OK(cmg_device_free(ctx, dA));
        }
        const int worlds[3] = {2, 4, 8};
        for(int world : worlds)
        {
            // boundaries of equal numbers of evaluated pairs (cosmopp_b200/partition.py: orbit_partition, mode 0): cumulative
            // cost up to q = 15 F q + 6 q (q + 1) / 2
            std::vector<int64_t> b(world + 1, 0);
            const double total = 15.0 * F * F + 3.0 * F * (F + 1);
            for(int k = 1; k < world; ++k)
            {
                const double target = total * k / world, a2 = 3.0, a1 = 15.0 * F + 3.0;
                const double q = (-a1 + std::sqrt(a1 * a1 + 4 * a2 * target)) / (2 * a2);
                b[k] = std::min<int64_t>(F, std::max<int64_t>(b[k - 1], (int64_t) std::llround(q / 32) * 32));
            }
            b[world] = F;
            double worst = 0, sum = 0;
            for(int r = 0; r < world; ++r)
            {
                cmg_orbit_shard sh;
                double *dStrips = nullptr, *dBox = nullptr;
                int64_t stripDoubles = 0, boxDoubles = 0;
                if(makeShard(ctx, nside, 0, b, r, sh, dStrips, dBox, stripDoubles, boxDoubles, false)) return 1;
                double best = 1e30;
                for(int rep = 0; rep < 3; ++rep)
                {
                    OK(cmg_tqu_orbit_sharded(ctx, tt.data(), te.data(), ee.data(), bb.data(), lmax, &sh, 0));
                    double ms = 0; OK(cmg_last_kernel_ms(ctx, &ms));
                    best = std::min(best, ms);
                }
                worst = std::max(worst, best);
                sum += best;
                std::printf("  world %d rank %d: q [%lld, %lld): %.3f ms, strips %.2f GB, outbox %.2f GB\n", world, r, (long long) b[r], (long long) b[r + 1], best,
                            stripDoubles * 8e-9, boxDoubles * 8e-9);
                OK(cmg_device_free(ctx, dStrips));
                OK(cmg_device_free(ctx, dBox));
            }
            std::printf("nside 64 lmax 192, %d ranks (one after the other on this GPU), mode 0: slowest %.3f ms, mean %.3f ms\n", world, worst, sum / world);
        }
        cmg_destroy(ctx);
        return 0;
    }

    // full comparison at sizes that fit twice
    const int cases[][2] = {{8, 20}, {16, 47}, {32, 96}};
    for(const auto& cs : cases)
    {
        const int nside = cs[0], lmax = cs[1];
        OK(cmg_set_pixels(ctx, nside, nullptr, 0));
        const int64_t n = cmg_npix(ctx), packed = cmg_packed_size(3 * n);
        std::vector<double> tt, te, ee, bb;
        weights(lmax, tt, te, ee, bb);
        double *dA = nullptr, *dB = nullptr;
        OK(cmg_device_malloc(ctx, packed * 8, (void**) &dA));
        OK(cmg_device_malloc(ctx, packed * 8, (void**) &dB));
        cmg_tqu_layout lay;
        OK(cmg_tqu_layout_single(ctx, dA, &lay));
        OK(cmg_tqu(ctx, tt.data(), te.data(), ee.data(), bb.data(), lmax, &lay));
        double msBase = 0; OK(cmg_last_kernel_ms(ctx, &msBase));
        std::vector<double> hA(packed), hB(packed);
        OK(cmg_copy_to_host(ctx, hA.data(), dA, packed * 8));
        OK(cmg_synchronize(ctx));
        const double dT = hA[0], dQ = hA[cmg_packed_index(n, n)];
        for(int mode = 3; mode >= 0; --mode)
        {
            cudaMemset(dB, 0xFF, packed * 8);                      // NaN pattern: an entry nobody writes shows up
            cudaDeviceSynchronize();
            const cmg_status s = cmg_tqu_orbit(ctx, tt.data(), te.data(), ee.data(), bb.data(), lmax, dB, mode);
            if(s != CMG_OK) { std::printf("nside %d mode %d: launch failed: %s\n", nside, mode, cmg_last_error(ctx)); rc = 1; continue; }
            double ms = 0; OK(cmg_last_kernel_ms(ctx, &ms));
            const cmg_status s2 = cmg_synchronize(ctx);
            if(s2 != CMG_OK) { std::printf("nside %d mode %d: kernel fault: %s\n", nside, mode, cmg_last_error(ctx)); return 2; }
            OK(cmg_copy_to_host(ctx, hB.data(), dB, packed * 8));
            OK(cmg_synchronize(ctx));
            int64_t nan = 0, firstBad = -1;
            double worst = 0;
            for(int64_t col = 0, e = 0; col < 3 * n; ++col)
            {
                const double scale = col < n ? dT : dQ;
                for(int64_t row = 0; row <= col; ++row, ++e)
                {
                    if(std::isnan(hB[e])) { ++nan; if(firstBad < 0) firstBad = e; continue; }
                    const double d = std::fabs(hB[e] - hA[e]) / scale;
                    if(d > worst) { worst = d; if(d > 1e-11 && firstBad < 0) firstBad = e; }
                }
            }
            std::printf("nside %d lmax %d mode %d: unwritten %lld, max |orbit - cmg_tqu| / diag = %.3e, first bad entry %lld; %.3f ms (cmg_tqu %.3f ms)\n",
                        nside, lmax, mode, (long long) nan, worst, (long long) firstBad, ms, msBase);
            if(nan || worst > 1e-11) rc = 1;
        }
        OK(cmg_device_free(ctx, dA));
        OK(cmg_device_free(ctx, dB));
    }

    // sharded: every rank's pieces generated one after the other on this GPU, assembled, compared with cmg_tqu
    const int shardCases[][3] = {{16, 47, 2}, {16, 47, 3}, {32, 96, 8}};
    for(const auto& cs : shardCases)
    {
        const int nside = cs[0], lmax = cs[1], world = cs[2];
        OK(cmg_set_pixels(ctx, nside, nullptr, 0));
        const int64_t n = cmg_npix(ctx), packed = cmg_packed_size(3 * n), F = (int64_t) nside * nside;
        std::vector<double> tt, te, ee, bb;
        weights(lmax, tt, te, ee, bb);
        double *dA = nullptr, *dB = nullptr;
        OK(cmg_device_malloc(ctx, packed * 8, (void**) &dA));
        OK(cmg_device_malloc(ctx, packed * 8, (void**) &dB));
        cmg_tqu_layout lay;
        OK(cmg_tqu_layout_single(ctx, dA, &lay));
        OK(cmg_tqu(ctx, tt.data(), te.data(), ee.data(), bb.data(), lmax, &lay));
        std::vector<double> hA(packed), hB(packed);
        OK(cmg_copy_to_host(ctx, hA.data(), dA, packed * 8));
        OK(cmg_synchronize(ctx));
        const double dT = hA[0], dQ = hA[cmg_packed_index(n, n)];
        for(int mode = 2; mode >= 0; --mode)
        {
            cudaMemset(dB, 0xFF, packed * 8);
            double msMax = 0;
            std::vector<cmg_orbit_shard> shards(world);
            std::vector<double*> bufs;
            std::vector<int64_t> bnd(world + 1);
            for(int r = 0; r <= world; ++r)
                bnd[r] = r == world ? F : (F * r / world) / 32 * 32;
            for(int r = 0; r < world; ++r)
            {
                double *dStrips = nullptr, *dBox = nullptr;
                int64_t stripDoubles = 0, boxDoubles = 0;
                if(makeShard(ctx, nside, mode, bnd, r, shards[r], dStrips, dBox, stripDoubles, boxDoubles, true)) return 1;
                bufs.push_back(dStrips);
                bufs.push_back(dBox);
                OK(cmg_tqu_orbit_sharded(ctx, tt.data(), te.data(), ee.data(), bb.data(), lmax, &shards[r], mode));
                double ms = 0; OK(cmg_last_kernel_ms(ctx, &ms));
                msMax = std::max(msMax, ms);
            }
            for(int parts = 1; parts <= 2; ++parts)               // the strips of all ranks first (they have holes), then the outboxes
                for(int r = 0; r < world; ++r)
                    OK(cmg_tqu_orbit_assemble(ctx, &shards[r], mode, parts, dB));
            OK(cmg_synchronize(ctx));
            for(double* b : bufs)
                OK(cmg_device_free(ctx, b));
            OK(cmg_copy_to_host(ctx, hB.data(), dB, packed * 8));
            OK(cmg_synchronize(ctx));
            int64_t nan = 0; double worst = 0;
            for(int64_t col = 0, e = 0; col < 3 * n; ++col)
                for(int64_t row = 0; row <= col; ++row, ++e)
                {
                    if(std::isnan(hB[e])) { ++nan; continue; }
                    worst = std::max(worst, std::fabs(hB[e] - hA[e]) / (col < n ? dT : dQ));
                }
            std::printf("sharded nside %d lmax %d world %d mode %d: unwritten %lld, max |assembled - cmg_tqu| / diag = %.3e; slowest rank %.3f ms\n",
                        nside, lmax, world, mode, (long long) nan, worst, msMax);
            if(nan || worst > 1e-11) rc = 1;
        }
        OK(cmg_device_free(ctx, dA));
        OK(cmg_device_free(ctx, dB));
    }

    // flagship size, one rank of eight: time only (first and last range)
    if(timingNside >= 64)
    {
        const int nside = 64, lmax = 192, world = 8;
        OK(cmg_set_pixels(ctx, nside, nullptr, 0));
        const int64_t n = cmg_npix(ctx), F = (int64_t) nside * nside;
        std::vector<double> tt, te, ee, bb;
        weights(lmax, tt, te, ee, bb);
        for(int r = 0; r < world; r += world - 1)
        {
            cmg_orbit_shard sh;
            std::vector<int64_t> bnd(world + 1);
            for(int k = 0; k <= world; ++k)
                bnd[k] = F * k / world;
            double *dStrips = nullptr, *dBox = nullptr;
            int64_t stripDoubles = 0, boxDoubles = 0;
            if(makeShard(ctx, nside, 0, bnd, r, sh, dStrips, dBox, stripDoubles, boxDoubles, false)) return 1;
            for(int mode = 2; mode >= 0; --mode)
            {
                double best = 1e30;
                for(int rep = 0; rep < 3; ++rep)
                {
                    OK(cmg_tqu_orbit_sharded(ctx, tt.data(), te.data(), ee.data(), bb.data(), lmax, &sh, mode));
                    double ms = 0; OK(cmg_last_kernel_ms(ctx, &ms));
                    best = std::min(best, ms);
                }
                std::printf("nside 64 lmax 192, rank %d of 8 (equal q ranges), mode %d: %.2f ms, strips %.2f GB, outbox %.2f GB\n", r, mode, best, stripDoubles * 8e-9, boxDoubles * 8e-9);
            }
            OK(cmg_device_free(ctx, dStrips));
            OK(cmg_device_free(ctx, dBox));
        }
    }

    // flagship size: one buffer, three 32 MB windows compared, timings
    if(timingNside >= 64)
    {
        const int nside = 64, lmax = 192;
        OK(cmg_set_pixels(ctx, nside, nullptr, 0));
        const int64_t n = cmg_npix(ctx), packed = cmg_packed_size(3 * n);
        std::vector<double> tt, te, ee, bb;
        weights(lmax, tt, te, ee, bb);
        double* dA = nullptr;
        OK(cmg_device_malloc(ctx, packed * 8, (void**) &dA));
        const int64_t win = 4 << 20;
        const int64_t starts[3] = {0, packed / 2, packed - win};
        std::vector<double> ref(3 * win), got(3 * win);
        cmg_tqu_layout lay;
        OK(cmg_tqu_layout_single(ctx, dA, &lay));
        double best = 1e30;
        for(int rep = 0; rep < 3; ++rep)
        {
            OK(cmg_tqu(ctx, tt.data(), te.data(), ee.data(), bb.data(), lmax, &lay));
            double ms = 0; OK(cmg_last_kernel_ms(ctx, &ms));
            best = std::min(best, ms);
        }
        for(int w = 0; w < 3; ++w) OK(cmg_copy_to_host(ctx, ref.data() + w * win, dA + starts[w], win * 8));
        OK(cmg_synchronize(ctx));
        std::printf("nside 64 lmax 192: cmg_tqu %.2f ms\n", best);
        for(int mode = 3; mode >= 0; --mode)
        {
            cudaMemset(dA, 0xFF, packed * 8);
            cudaDeviceSynchronize();
            double bestO = 1e30;
            for(int rep = 0; rep < 3; ++rep)
            {
                OK(cmg_tqu_orbit(ctx, tt.data(), te.data(), ee.data(), bb.data(), lmax, dA, mode));
                double ms = 0; OK(cmg_last_kernel_ms(ctx, &ms));
                bestO = std::min(bestO, ms);
            }
            for(int w = 0; w < 3; ++w) OK(cmg_copy_to_host(ctx, got.data() + w * win, dA + starts[w], win * 8));
            OK(cmg_synchronize(ctx));
            double worst = 0; int64_t nan = 0;
            for(int64_t e = 0; e < 3 * win; ++e)
            {
                if(std::isnan(got[e])) { ++nan; continue; }
                worst = std::max(worst, std::fabs(got[e] - ref[e]));
            }
            std::printf("nside 64 lmax 192 mode %d: %.2f ms (%.2fx), windows: unwritten %lld, max abs diff %.3e (TT diag %.3e)\n",
                        mode, bestO, best / bestO, (long long) nan, worst, ref[0]);
            if(nan) rc = 1;
        }
        OK(cmg_device_free(ctx, dA));
    }
    cmg_destroy(ctx);
    std::printf(rc ? "ORBIT CHECK FAILED\n" : "ORBIT CHECK OK\n");
    return rc;
}
