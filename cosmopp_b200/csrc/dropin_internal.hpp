// shared by the drop-in classes: one generator context per GPU, created on first use, every use of it under its lock
#ifndef COSMO_PP_B200_DROPIN_INTERNAL_HPP
#define COSMO_PP_B200_DROPIN_INTERNAL_HPP

#include <cmg.h>

#include <mutex>
#include <vector>

// The C ABI's contexts are independent of each other; a context itself is one GPU + one stream + the resident geometry, so two
// host threads must not interleave on it (one would re-bind the pixel set under the other's launch).  Every drop-in operation
// holds the lock of its device for its whole duration; different devices run concurrently.
struct DropinDevice
{
    cmg_ctx* ctx;
    int device;
    std::recursive_mutex mutex;
};

DropinDevice& cmgDropinDevice(int device);       // throws StandardException when the GPU is not usable (no CPU fallback)
int cmgDropinCurrentDevice();                    // what CMatrixGenerator::setDevice chose on THIS thread (default 0)

struct DropinLock
{
    DropinDevice& dev;
    std::lock_guard<std::recursive_mutex> guard;
    explicit DropinLock(int device) : dev(cmgDropinDevice(device)), guard(dev.mutex) {}
    cmg_ctx* ctx() const { return dev.ctx; }
};

// pixel window of nSide up to lMax as CMatrixGenerator resolves it (setPixelWindow / HEALPix data directory)
void cmgDropinPixelWindow(long nSide, int lMax, bool polarization, std::vector<double>& w);

#endif
