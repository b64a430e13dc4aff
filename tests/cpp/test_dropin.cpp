// Exercises the C++ drop-in classes the way the reference's own test does
// (reference source/test_like_low.cpp:183-186: clToCMatrix, getFiducialMatrix, generateNoiseMatrix, maskMatrix),
// plus the file formats.  Mode "cpu" needs no GPU; mode "gpu" writes matrices that tests/test_dropin_cpp.py
// compares with the oracle.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include <c_matrix.hpp>
#include <c_matrix_generator.hpp>
#include <exception_handler.hpp>
#include <likelihood.hpp>
#include <pixel_likelihood.hpp>
#include <utils.hpp>

static int failures = 0;
#define EXPECT(cond)                                                                   \
    do                                                                                 \
    {                                                                                  \
        if(!(cond))                                                                    \
        {                                                                              \
            ++failures;                                                                \
            std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond);              \
        }                                                                              \
    } while(0)

template <typename F> static bool throwsStandard(F f)
{
    try { f(); }
    catch(StandardException&) { return true; }
    catch(...) { return false; }
    return false;
}

static std::vector<double> readDoubles(const std::string& path)
{
    std::ifstream in(path.c_str(), std::ios::binary);
    in.seekg(0, std::ios::end);
    const std::streamsize n = in.tellg();
    in.seekg(0);
    std::vector<double> v(static_cast<size_t>(n / 8));
    in.read(reinterpret_cast<char*>(v.data()), n);
    return v;
}

static std::vector<int> readInts(const std::string& path)
{
    std::ifstream in(path.c_str(), std::ios::binary);
    in.seekg(0, std::ios::end);
    const std::streamsize n = in.tellg();
    in.seekg(0);
    std::vector<int> v(static_cast<size_t>(n / 4));
    in.read(reinterpret_cast<char*>(v.data()), n);
    return v;
}

static void cpuTests(const std::string& dir)
{
    // packed layout: index(i,j) = j(j+1)/2 + i, symmetric access (reference source/c_matrix.cpp:27-39)
    CMatrix m(5);
    EXPECT(m.getNPix() == 5);
    for(int j = 0; j < 5; ++j)
        for(int i = 0; i <= j; ++i)
            m.element(i, j) = 10 * j + i + 0.25;
    EXPECT(m.element(3, 1) == m.element(1, 3));
    EXPECT(&m.element(2, 4) == m.packed() + (4 * 5 / 2 + 2));
    EXPECT(m.packedSize() == 15);
    m.comment() = "hello matrix";

    // binary round trip
    const std::string bin = dir + "/m.dat", txt = dir + "/m.txt";
    m.writeIntoFile(bin.c_str());
    CMatrix r(bin.c_str());
    EXPECT(r.getNPix() == 5 && r.comment() == "hello matrix");
    for(int k = 0; k < 15; ++k) EXPECT(r.packed()[k] == m.packed()[k]);
    // text round trip (default ostream precision: 6 significant digits)
    m.writeIntoTextFile(txt.c_str());
    CMatrix t(3);
    t.readFromTextFile(txt.c_str());
    EXPECT(t.getNPix() == 5 && t.comment() == "hello matrix");
    for(int k = 0; k < 15; ++k) EXPECT(std::fabs(t.packed()[k] - m.packed()[k]) <= 1e-5 * std::fabs(m.packed()[k]));

    // copy semantics are deep
    CMatrix c(m);
    c.element(0, 0) = -1;
    EXPECT(m.element(0, 0) == 0.25);
    CMatrix a(2);
    a = m;
    EXPECT(a.getNPix() == 5 && a.element(4, 4) == m.element(4, 4));

    // maskMatrix gather (reference source/c_matrix.cpp:182-201)
    std::vector<int> good;
    good.push_back(1); good.push_back(3); good.push_back(4);
    CMatrix g(m);
    g.maskMatrix(good);
    EXPECT(g.getNPix() == 3);
    for(int b = 0; b < 3; ++b)
        for(int aa = 0; aa <= b; ++aa)
            EXPECT(g.element(aa, b) == m.element(good[aa], good[b]));

    // error convention
    EXPECT(throwsStandard([] { CMatrix bad(0); }));
    EXPECT(throwsStandard([&] { CMatrix bad((dir + "/does_not_exist.dat").c_str()); }));
    EXPECT(throwsStandard([&] { m.writeIntoFile((dir + "/no/such/dir/x.dat").c_str()); }));

    // noise matrix (reference source/c_matrix_generator.cpp:774-787)
    CMatrix* noise = CMatrixGenerator::generateNoiseMatrix(2, 0.5);
    EXPECT(noise->getNPix() == 48 && noise->comment() == "noise matrix");
    EXPECT(noise->element(7, 7) == 0.25 && noise->element(7, 8) == 0);
    delete noise;

    // beam (reference source/utils.cpp:54-64)
    EXPECT(Utils::beamFunction(10, 0) == 1.0);
    const double sigma = std::sqrt(8 * std::log(2.0)) / (10.0 * 3.141592653589793 / 180);
    EXPECT(std::fabs(Utils::beamFunction(30, 10.0) - std::exp(-30 * 31 / (2 * sigma * sigma))) < 1e-16);

    // C_l text files
    {
        std::ofstream out((dir + "/cl.txt").c_str());
        out << "0\n0\n1.5\n2.5e-1\n";
    }
    std::vector<double> cl;
    Utils::readClFromFile((dir + "/cl.txt").c_str(), cl);
    EXPECT(cl.size() == 4 && cl[2] == 1.5 && cl[3] == 0.25);
    {
        std::ofstream out((dir + "/dl.txt").c_str());
        out << "0 0\n1 0\n2 6.0\n";
    }
    Utils::readClFromFile((dir + "/dl.txt").c_str(), cl, true, true);
    EXPECT(cl.size() == 3 && std::fabs(cl[2] - 6.0 * 2 * 3.141592653589793 / 6) < 1e-15);
    EXPECT(throwsStandard([&] { std::vector<double> x; Utils::readClFromFile((dir + "/nope.txt").c_str(), x); }));

    // FITS mask written by the Python side (NESTED, Nside=4) -> same good pixels as the list next to it
    {
        long nSide = 0;
        std::vector<int> gp;
        Utils::readMask((dir + "/mask_nest.fits").c_str(), nSide, gp);
        const std::vector<int> want = readInts(dir + "/mask_good.i32");
        EXPECT(nSide == 4 && gp == want);
        EXPECT(throwsStandard([&] { long n; std::vector<int> x; Utils::readMask((dir + "/mask_ring.fits").c_str(), n, x); }));
        // the same mask as a 64-bit integer column, and as a scaled integer column (TSCAL / TZERO)
        std::vector<int> gk, gs;
        long nk = 0, ns = 0;
        Utils::readMask((dir + "/mask_k.fits").c_str(), nk, gk);
        Utils::readMask((dir + "/mask_scaled.fits").c_str(), ns, gs);
        EXPECT(nk == 4 && ns == 4 && gk == want && gs == want);
    }

    // Legendre container: on-demand values and the reference's file layout
    {
        std::vector<int> px;
        for(int k = 0; k < 12; k += 2) px.push_back(k);
        LegendrePolynomialContainer lp(6, 1, &px);
        EXPECT(std::fabs(lp.value(0, 3, 1) - 1.0) < 1e-15);
        EXPECT(std::fabs(lp.value(4, 2, 2) - 1.0) < 1e-12);          // P_l(1) = 1 on the diagonal
        lp.writeIntoFile((dir + "/lp.dat").c_str());
        LegendrePolynomialContainer back((dir + "/lp.dat").c_str());
        EXPECT(back.lMax() == 6 && back.nPix() == 6);
        for(int l = 0; l <= 6; ++l)
            for(int j = 0; j < 6; ++j)
                for(int i = 0; i <= j; ++i)
                    EXPECT(back.value(l, j, i) == lp.value(l, j, i));
    }

    // the harmonic-space routes are declared but outside this library's path
    EXPECT(throwsStandard([] { CMatrixGenerator::calculateNoiseMatrix("a", "b", 1.0, 1.0); }));

    // HEALPix pixel window table (reference source/utils.cpp:66-170): HEALPIX_DATA_DIR/pixel_window_n0016.fits, HDU 2, column 1
    // = temperature, column 2 = polarization, times the beam.  tests/test_dropin_cpp.py wrote the files and checks the numbers.
    {
        std::ifstream probe((dir + "/pixel_window_n0016.fits").c_str());
        if(probe)
        {
            CMatrixGenerator::setHealpixDataDir(dir.c_str());
            std::vector<double> fT, fP, f0;
            Utils::readPixelWindowFunction(fT, 16, 47, 10.0, false);
            Utils::readPixelWindowFunction(fP, 16, 47, 10.0, true);
            Utils::readPixelWindowFunction(f0, 16, 20, 0.0, false);                       // fwhm = 0: the window alone
            EXPECT(fT.size() == 48 && fP.size() == 48 && f0.size() == 21);
            std::FILE* f = std::fopen((dir + "/window_read.txt").c_str(), "w");
            for(int l = 0; l <= 47; ++l)
                std::fprintf(f, "%.17g %.17g %.17g\n", fT[l], fP[l], l <= 20 ? f0[l] : 0.0);
            std::fclose(f);
            // the reference's error behaviour: a table that stops short of lMax, a missing file
            EXPECT(throwsStandard([&] { std::vector<double> g; Utils::readPixelWindowFunction(g, 16, 64, 10.0, false); }));
            EXPECT(throwsStandard([&] { std::vector<double> g; Utils::readPixelWindowFunction(g, 32, 10, 10.0, false); }));
            CMatrixGenerator::setHealpixDataDir("");
        }
    }
}

static void gpuTests(const std::string& dir)
{
    const long nSide = 8;
    const int lMax = 20, lMaxFid = 4 * nSide;
    const std::vector<double> cl = readDoubles(dir + "/cl_tt.f64");           // 4 nSide + 1 values
    const std::vector<int> good = readInts(dir + "/good.i32");
    const std::vector<double> ones(static_cast<size_t>(lMaxFid + 1), 1.0);
    CMatrixGenerator::setPixelWindow(nSide, ones, ones);
    CMatrixGenerator::setPixelWindow(4, ones, ones);

    // exactly the call sequence of reference source/test_like_low.cpp:181-186
    std::vector<double> clCopy(cl.begin(), cl.begin() + lMax + 1);
    CMatrix* cMatrix = CMatrixGenerator::clToCMatrix(clCopy, nSide, 10.0, &good);
    CMatrix* fiducialMatrix = CMatrixGenerator::getFiducialMatrix(cl, nSide, lMax, 10.0, &good);
    CMatrix* noiseMatrix = CMatrixGenerator::generateNoiseMatrix(nSide, 1e-2);
    noiseMatrix->maskMatrix(good);
    EXPECT(cMatrix->getNPix() == static_cast<int>(good.size()));
    EXPECT(fiducialMatrix->comment() == "fiducial matrix");
    EXPECT(noiseMatrix->getNPix() == static_cast<int>(good.size()));
    // the generated matrices live on the GPU; the consumer takes them there: generate -> mask -> factorise -> evaluate moves
    // no matrix data from the device to the host (the noise matrix is built on the host and goes up once)
    EXPECT(cMatrix->hasDeviceCopy() && !cMatrix->hasHostCopy() && fiducialMatrix->hasDeviceCopy() && !fiducialMatrix->hasHostCopy());
    {
        long long h2d0 = 0, d2h0 = 0, h2d1 = 0, d2h1 = 0;
        CMatrixGenerator::transferCounters(h2d0, d2h0);
        std::vector<double> noForeground;
        Likelihood residentLike(*cMatrix, *fiducialMatrix, *noiseMatrix, good, noForeground);
        const std::vector<double> maps = readDoubles(dir + "/maps.f64");
        std::vector<double> t0(maps.begin(), maps.begin() + good.size());
        double chi2 = 0, logDet = 0;
        residentLike.calculate(t0, chi2, logDet);
        CMatrixGenerator::transferCounters(h2d1, d2h1);
        EXPECT(d2h1 == d2h0);                                                  // zero bytes of matrix data came back
        EXPECT(h2d1 - h2d0 == 8LL * noiseMatrix->packedSize());                // and only the host-built noise matrix went up
        EXPECT(!cMatrix->hasHostCopy() && !fiducialMatrix->hasHostCopy());
        std::FILE* f = std::fopen((dir + "/like_resident.txt").c_str(), "w");
        std::fprintf(f, "%.17g %.17g\n", chi2, logDet);
        std::fclose(f);
        // a copy of a device-resident matrix is made on the device; writing to the copy leaves the original alone
        CMatrix copy(*cMatrix);
        EXPECT(copy.hasDeviceCopy() && !copy.hasHostCopy());
        const double c00 = copy.element(0, 0);                                 // lazily materialised on the host
        EXPECT(copy.hasHostCopy() && c00 > 0);
        copy.element(0, 0) = 2 * c00;
        EXPECT(!copy.hasDeviceCopy() && static_cast<const CMatrix&>(*cMatrix).element(0, 0) == c00);
        CMatrixGenerator::transferCounters(h2d1, d2h1);
        EXPECT(d2h1 - d2h0 == 2 * 8LL * cMatrix->packedSize());                // the copy and the original, once each
    }
    cMatrix->writeIntoFile((dir + "/c.dat").c_str());
    fiducialMatrix->writeIntoFile((dir + "/c_fiducial.dat").c_str());
    noiseMatrix->writeIntoFile((dir + "/c_noise.dat").c_str());

    // the consumer, as in reference source/test_like_low.cpp:187-191 (plus the foreground-template variant)
    {
        const std::vector<double> maps = readDoubles(dir + "/maps.f64"), fore = readDoubles(dir + "/fore.f64");
        const size_t ng = good.size(), nMaps = maps.size() / ng;
        std::vector<std::vector<double> > t(nMaps);
        for(size_t k = 0; k < nMaps; ++k)
            t[k].assign(maps.begin() + k * ng, maps.begin() + (k + 1) * ng);
        std::vector<double> foreground;
        Likelihood like(*cMatrix, *fiducialMatrix, *noiseMatrix, good, foreground);
        std::vector<std::string> mapNames(nMaps, "test_map");
        std::vector<LikelihoodResult> results;
        like.calculateAll(t, mapNames, results);
        EXPECT(results.size() == nMaps && results[0].mapName == "test_map");
        double chi2 = 0, logDet = 0;
        const double l0 = like.calculate(t[0], chi2, logDet);
        // (one map or six at a time: the library's triangular solve may block differently, so equal to rounding only)
        EXPECT(std::fabs(chi2 - results[0].chi2) <= 1e-11 * chi2 && logDet == results[0].logDet && std::fabs(l0 - results[0].like) <= 1e-11 * std::fabs(l0));
        Likelihood likeF(*cMatrix, *fiducialMatrix, *noiseMatrix, good, fore);
        std::vector<LikelihoodResult> resultsF;
        likeF.calculateAll(t, mapNames, resultsF);
        std::FILE* f = std::fopen((dir + "/like.txt").c_str(), "w");
        for(size_t k = 0; k < nMaps; ++k)
            std::fprintf(f, "%.17g %.17g %.17g %.17g\n", results[k].chi2, results[k].logDet, resultsF[k].chi2, resultsF[k].logDet);
        std::fclose(f);
        // error behaviour of the reference: size mismatches and a matrix that is not positive definite throw
        EXPECT(throwsStandard([&] { std::vector<int> fewer(good.begin(), good.end() - 1); Likelihood bad(*cMatrix, *fiducialMatrix, *noiseMatrix, fewer, foreground); }));
        EXPECT(throwsStandard([&] { std::vector<double> shortMap(ng - 1, 0.0); double a, b; like.calculate(shortMap, a, b); }));
        CMatrix negative(*noiseMatrix);
        for(int i = 0; i < negative.getNPix(); ++i) negative.element(i, i) = -1e6;
        EXPECT(throwsStandard([&] { Likelihood bad(*cMatrix, *fiducialMatrix, negative, good, foreground); }));
    }
    // sampler plug-in: parameters -> C_l -> device matrix -> device likelihood behind Math::LikelihoodFunction
    {
        struct Amplitude : public ClModel
        {
            std::vector<double> base;
            void clTT(const double* params, int nParams, std::vector<double>& out)
            {
                for(size_t l = 0; l < out.size(); ++l)
                    out[l] = params[0] * base[l] * std::pow((l + 1.0) / 10.0, nParams > 1 ? params[1] : 0.0);
            }
        } model;
        model.base = clCopy;
        const std::vector<double> maps = readDoubles(dir + "/maps.f64"), fore = readDoubles(dir + "/fore.f64");
        const std::vector<double> t0(maps.begin(), maps.begin() + good.size());
        std::vector<double> noForeground;
        PixelLikelihoodTT plug(nSide, lMax, 10.0, good, *fiducialMatrix, *noiseMatrix, t0, noForeground, model);
        Math::LikelihoodFunction* asSamplerSeesIt = &plug;
        double p1[2] = {1.0, 0.0};
        const double l1 = asSamplerSeesIt->calculate(p1, 2);
        // same point through the classes of the reference's own test
        Likelihood like(*cMatrix, *fiducialMatrix, *noiseMatrix, good, noForeground);
        double chi2 = 0, logDet = 0;
        const double lRef = like.calculate(t0, chi2, logDet);
        EXPECT(std::fabs(l1 - lRef) <= 1e-10 * std::fabs(lRef));
        EXPECT(std::fabs(plug.lastChi2() - chi2) <= 1e-10 * chi2);
        // a batch of proposals equals the points one by one
        double pts[6] = {1.0, 0.0, 1.3, 0.05, 0.7, -0.1}, batch[3];
        plug.calculateBatch(pts, 2, 3, batch);
        std::FILE* f = std::fopen((dir + "/plug.txt").c_str(), "w");
        for(int k = 0; k < 3; ++k)
        {
            const double one = plug.calculate(pts + 2 * k, 2);
            EXPECT(std::fabs(one - batch[k]) <= 1e-10 * std::fabs(one));
            std::fprintf(f, "%.17g %.17g %.17g\n", pts[2 * k], pts[2 * k + 1], batch[k]);
        }
        std::fclose(f);
        EXPECT(std::fabs(batch[0] - l1) <= 1e-10 * std::fabs(l1) && batch[1] != batch[0]);
        PixelLikelihoodTT plugF(nSide, lMax, 10.0, good, *fiducialMatrix, *noiseMatrix, t0, fore, model);
        Likelihood likeF(*cMatrix, *fiducialMatrix, *noiseMatrix, good, fore);
        EXPECT(std::fabs(plugF.calculate(p1, 2) - likeF.calculate(t0, chi2, logDet)) <= 1e-10 * std::fabs(lRef));
    }
    delete cMatrix;
    delete fiducialMatrix;
    delete noiseMatrix;

    // full sky through the file overload
    CMatrix* full = CMatrixGenerator::clToCMatrix((dir + "/cl_short.txt").c_str(), 4, 12, 10.0);
    EXPECT(full->getNPix() == 192);
    full->writeIntoFile((dir + "/c_full.dat").c_str());
    delete full;

    // polarized addition
    const std::vector<double> te = readDoubles(dir + "/cl_te.f64"), ee = readDoubles(dir + "/cl_ee.f64"), bb = readDoubles(dir + "/cl_bb.f64");
    std::vector<double> tt(cl.begin(), cl.begin() + te.size());
    CMatrix* pol = CMatrixGenerator::clToCMatrixPol(tt, te, ee, bb, nSide, 10.0, &good);
    EXPECT(pol->getNPix() == 3 * static_cast<int>(good.size()));
    pol->writeIntoFile((dir + "/c_pol.dat").c_str());
    delete pol;

    // which pixel window the polarized part is smoothed with: HEALPix's polarization table (default) or, as the reference's own
    // polarization routine does (source/c_matrix_generator.cpp:534), the temperature table
    {
        const std::vector<double> wT = readDoubles(dir + "/win_t.f64"), wP = readDoubles(dir + "/win_p.f64");
        CMatrixGenerator::setPixelWindow(nSide, wT, wP);
        CMatrix* a = CMatrixGenerator::clToCMatrixPol(tt, te, ee, bb, nSide, 10.0, &good);
        a->writeIntoFile((dir + "/c_pol_wtp.dat").c_str());
        delete a;
        CMatrixGenerator::setPolarizationUsesTemperatureWindow(true);
        CMatrix* b = CMatrixGenerator::clToCMatrixPol(tt, te, ee, bb, nSide, 10.0, &good);
        b->writeIntoFile((dir + "/c_pol_wtt.dat").c_str());
        delete b;
        CMatrixGenerator::setPolarizationUsesTemperatureWindow(false);
        CMatrixGenerator::setPixelWindow(nSide, ones, ones);
    }

    // LikelihoodPolarization, pixel-space part (reference source/likelihood.cpp:341-406, 536-612): full-sky [T;Q;U] matrix ->
    // its [Q;U] block -> restricted to the unmasked pixels with a diagonal N^-1 -> chi2 and log det
    {
        const long nSideP = 4;
        const std::vector<double> onesP(static_cast<size_t>(4 * nSideP + 1), 1.0);
        CMatrixGenerator::setPixelWindow(nSideP, onesP, onesP);
        std::vector<double> tt4(tt.begin(), tt.begin() + 13), te4(te.begin(), te.begin() + 13), ee4(ee.begin(), ee.begin() + 13), bb4(bb.begin(), bb.begin() + 13);
        CMatrix* tqu = CMatrixGenerator::clToCMatrixPol(tt4, te4, ee4, bb4, nSideP, 10.0);
        CMatrix* qu = LikelihoodPolarization::polarizationBlock(*tqu);
        EXPECT(qu->getNPix() == 2 * 192 && qu->hasDeviceCopy() && !qu->hasHostCopy());
        qu->writeIntoFile((dir + "/c_qu.dat").c_str());
        const std::vector<int> goodP = readInts(dir + "/good_p.i32");
        const std::vector<double> vP = readDoubles(dir + "/v_p.f64"), predP = readDoubles(dir + "/pred_p.f64"), nInvDiag = readDoubles(dir + "/ninv_diag.f64");
        CMatrix nInv(2 * 192);
        for(int i = 0; i < 2 * 192; ++i)
            nInv.element(i, i) = nInvDiag[i];
        nInv.element(3, 200) = 0.01;                                           // and one off-diagonal entry, so that N^-1 is not just a scaling
        LikelihoodPolarization likeP(*qu, nSideP, goodP, nInv);
        double chi2 = 0, logDet = 0, chi2b = 0, logDetb = 0;
        std::vector<double> none;
        const double l1 = likeP.calculate(vP, none, chi2, logDet);
        const double l2 = likeP.calculate(vP, predP, chi2b, logDetb);
        EXPECT(l1 == chi2 + logDet && l2 == chi2b + logDetb && logDet == logDetb && chi2 != chi2b);
        std::FILE* f = std::fopen((dir + "/like_pol.txt").c_str(), "w");
        std::fprintf(f, "%.17g %.17g %.17g\n", chi2, chi2b, logDet);
        std::fclose(f);
        // the text-file form of N^-1 (the reference's n_inv.txt) gives the same object
        {
            std::ofstream out((dir + "/n_inv.txt").c_str());
            out.precision(17);
            for(int i = 0; i < 2 * 192; ++i)
            {
                for(int j = 0; j < 2 * 192; ++j)
                    out << static_cast<const CMatrix&>(nInv).element(i, j) << ' ';
                out << '\n';
            }
        }
        LikelihoodPolarization likeFile(*qu, nSideP, goodP, (dir + "/n_inv.txt").c_str());
        double chi2f = 0, logDetf = 0;
        likeFile.calculate(vP, predP, chi2f, logDetf);
        EXPECT(std::fabs(chi2f - chi2b) <= 1e-12 * std::fabs(chi2b) && std::fabs(logDetf - logDetb) <= 1e-12 * std::fabs(logDetb));
        EXPECT(throwsStandard([&] { std::vector<double> shortV(3, 0.0); double a, b; likeP.calculate(shortV, none, a, b); }));
        EXPECT(throwsStandard([&] { CMatrix odd(5); LikelihoodPolarization bad(odd, 0, goodP, nInv); }));
        EXPECT(throwsStandard([&] { LikelihoodPolarization bad(*qu, nSideP, goodP, (dir + "/no_such_n_inv.txt").c_str()); }));
        delete tqu;
        delete qu;
    }

    // two host threads on one GPU: a sampler thread evaluating the plug-in while another thread generates matrices for a
    // different pixel set -- every call holds the device's lock for its duration, so neither sees the other's geometry
    {
        struct Flat : public ClModel
        {
            std::vector<double> base;
            void clTT(const double* params, int, std::vector<double>& out) { for(size_t l = 0; l < out.size(); ++l) out[l] = params[0] * base[l]; }
        } model;
        model.base = clCopy;
        const std::vector<double> maps = readDoubles(dir + "/maps.f64");
        const std::vector<double> t0(maps.begin(), maps.begin() + good.size());
        std::vector<double> noForeground;
        CMatrix* fid = CMatrixGenerator::getFiducialMatrix(cl, nSide, lMax, 10.0, &good);
        CMatrix* noise = CMatrixGenerator::generateNoiseMatrix(nSide, 1e-2);
        noise->maskMatrix(good);
        PixelLikelihoodTT plug(nSide, lMax, 10.0, good, *fid, *noise, t0, noForeground, model);
        double p[1] = {1.1};
        const double want = plug.calculate(p, 1);
        std::vector<double> short4(cl.begin(), cl.begin() + 13);
        CMatrix* ref4 = CMatrixGenerator::clToCMatrix(short4, 4, 10.0);
        const double ref00 = static_cast<const CMatrix&>(*ref4).element(0, 0), ref17 = static_cast<const CMatrix&>(*ref4).element(1, 7);
        int bad = 0;
        std::thread sampler([&] { for(int k = 0; k < 12; ++k) { double q[1] = {1.1}; if(plug.calculate(q, 1) != want) ++bad; } });
        std::thread generator([&]
        {
            for(int k = 0; k < 12; ++k)
            {
                CMatrix* m = CMatrixGenerator::clToCMatrix(short4, 4, 10.0);
                const CMatrix& cm = *m;
                if(cm.element(0, 0) != ref00 || cm.element(1, 7) != ref17) ++bad;
                delete m;
            }
        });
        sampler.join();
        generator.join();
        EXPECT(bad == 0);
        delete fid;
        delete noise;
        delete ref4;
    }

    // errors surface as StandardException
    EXPECT(throwsStandard([&] { std::vector<double> empty; CMatrixGenerator::clToCMatrix(empty, nSide, 10.0); }));
    EXPECT(throwsStandard([&] { CMatrixGenerator::getFiducialMatrix(clCopy, nSide, lMax, 10.0); }));      // cl too short
    EXPECT(throwsStandard([&] { CMatrixGenerator::clToCMatrix(clCopy, 12, 10.0); }));                    // nSide not a power of two
    CMatrixGenerator::clearPixelWindow(nSide);
    EXPECT(throwsStandard([&] { CMatrixGenerator::clToCMatrix(clCopy, nSide, 10.0, &good); }));          // no window available
}

int main(int argc, char** argv)
{
    if(argc < 3)
    {
        std::printf("usage: test_dropin cpu|gpu <dir>\n");
        return 2;
    }
    const std::string mode = argv[1], dir = argv[2];
    try
    {
        if(mode == "cpu") cpuTests(dir);
        else gpuTests(dir);
    }
    catch(std::exception& e)
    {
        std::printf("UNEXPECTED EXCEPTION: %s\n", e.what());
        return 1;
    }
    std::printf("%s: %d failure(s)\n", mode.c_str(), failures);
    return failures ? 1 : 0;
}
