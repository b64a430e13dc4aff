// CMatrix of include/c_matrix.hpp: host-side container with the reference's file formats
// (reference source/c_matrix.cpp:41-158) and 64-bit packed indexing.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <c_matrix.hpp>
#include <cmg.h>
#include <exception_handler.hpp>
#include <utils.hpp>

#include "dropin_internal.hpp"

namespace
{
[[noreturn]] void raise(const std::string& text)
{
    throw StandardException(text);
}
}

void CMatrix::allocate(int nPix)
{
    if(nPix <= 0)
        raise("the number of pixels must be positive.");
    nPix_ = nPix;
    const std::int64_t bytes = packedSize() * static_cast<std::int64_t>(sizeof(double));
    void* p = NULL;
    pinned_ = false;
    // page-locked when a GPU is there (copies to and from the device run at PCIe speed)
    if(cmg_device_count() > 0 && cmg_host_malloc_pinned(bytes, &p) == CMG_OK)
        pinned_ = true;
    else
    {
        p = std::malloc(static_cast<size_t>(bytes));
        if(!p)
            raise("out of memory allocating a covariance matrix");
    }
    data_ = static_cast<double*>(p);
    std::memset(data_, 0, static_cast<size_t>(bytes));
    state_.store(kHost, std::memory_order_release);
}

void CMatrix::dropDevice() const
{
    if(dev_)
    {
        DropinLock lock(device_);
        cmg_synchronize(lock.ctx());
        cmg_device_free(lock.ctx(), dev_);
        dev_ = NULL;
    }
    state_.store(state_.load() & ~kDevice, std::memory_order_release);
}

void CMatrix::release()
{
    dropDevice();
    if(data_)
    {
        if(pinned_)
            cmg_host_free_pinned(data_);
        else
            std::free(data_);
        data_ = NULL;
    }
    state_.store(0, std::memory_order_release);
}

CMatrix::CMatrix(int nPix) : nPix_(0), data_(NULL), pinned_(false), dev_(NULL), device_(0), state_(0), symNSide_(0), symStrips_(0) { allocate(nPix); }

CMatrix::CMatrix(const char* fileName) : nPix_(0), data_(NULL), pinned_(false), dev_(NULL), device_(0), state_(0), symNSide_(0), symStrips_(0)
{
    readFromFile(fileName);
}

CMatrix::CMatrix(DeviceOnly, int nPix) : nPix_(nPix), data_(NULL), pinned_(false), dev_(NULL), device_(0), state_(0), symNSide_(0), symStrips_(0)
{
    if(nPix <= 0)
        raise("the number of pixels must be positive.");
}

CMatrix* CMatrix::newOnDevice(int nPix, int device, double** dPacked, long fullSkyNSide, int strips)
{
    CMatrix* m = new CMatrix(DeviceOnly(), nPix);
    DropinLock lock(device);
    void* p = NULL;
    if(cmg_device_malloc(lock.ctx(), m->packedSize() * static_cast<std::int64_t>(sizeof(double)), &p) != CMG_OK)
    {
        const std::string text = std::string("CMatrix: ") + cmg_last_error(lock.ctx());
        delete m;
        raise(text);
    }
    m->dev_ = static_cast<double*>(p);
    m->device_ = device;
    m->symNSide_ = fullSkyNSide;
    m->symStrips_ = strips;
    m->state_.store(kDevice, std::memory_order_release);
    if(dPacked)
        *dPacked = m->dev_;
    return m;
}

// first host-side use of a matrix that lives on the device: one copy over PCIe (with the host expansion of the rotated images
// where the matrix is a full-sky one and large enough for that to pay, cmg_matrix_to_host)
void CMatrix::hostForRead() const
{
    DropinLock lock(device_);
    if(state_.load(std::memory_order_acquire) & kHost)
        return;                                  // another thread got here first
    if(!(state_.load() & kDevice) || !dev_)
        raise("CMatrix: the matrix holds no data");
    if(!data_)
    {
        const std::int64_t bytes = packedSize() * static_cast<std::int64_t>(sizeof(double));
        void* p = NULL;
        if(cmg_host_malloc_pinned(bytes, &p) == CMG_OK)
            pinned_ = true;
        else
        {
            p = std::malloc(static_cast<size_t>(bytes));
            pinned_ = false;
            if(!p)
                raise("out of memory allocating a covariance matrix");
        }
        data_ = static_cast<double*>(p);
    }
    int strips = 0;
    if(symNSide_ > 0 && (symStrips_ == 1 || symStrips_ == 3) && cmg_nside2npix(symNSide_) * symStrips_ == nPix_)
    {
        // the symmetric copy needs the full-sky geometry of that nSide bound to the context (someone may have re-bound it)
        if(cmg_set_pixels(lock.ctx(), symNSide_, NULL, 0) == CMG_OK)
            strips = symStrips_;
    }
    if(cmg_matrix_to_host(lock.ctx(), dev_, nPix_, strips, data_) != CMG_OK)
        raise(std::string("CMatrix: ") + cmg_last_error(lock.ctx()));
    state_.store(kHost | kDevice, std::memory_order_release);
}

void CMatrix::hostForWrite()
{
    if(!(state_.load(std::memory_order_acquire) & kHost))
        hostForRead();
    dropDevice();                                // the caller is about to change entries: the host copy is the matrix now
    symNSide_ = 0;
    symStrips_ = 0;
}

const double* CMatrix::devicePacked(int device) const
{
    DropinLock lock(device);
    if((state_.load(std::memory_order_acquire) & kDevice) && device_ == device)
        return dev_;
    if(state_.load() & kDevice)
    {
        hostForRead();                           // lives on another GPU: through the host
        dropDevice();
    }
    if(!(state_.load() & kHost))
        raise("CMatrix: the matrix holds no data");
    void* p = NULL;
    const std::int64_t bytes = packedSize() * static_cast<std::int64_t>(sizeof(double));
    if(cmg_device_malloc(lock.ctx(), bytes, &p) != CMG_OK || cmg_copy_to_device(lock.ctx(), p, data_, bytes) != CMG_OK ||
       cmg_synchronize(lock.ctx()) != CMG_OK)
    {
        const std::string text = std::string("CMatrix: cannot place the matrix on the GPU: ") + cmg_last_error(lock.ctx());
        if(p) cmg_device_free(lock.ctx(), p);
        raise(text);
    }
    dev_ = static_cast<double*>(p);
    device_ = device;
    state_.store(kHost | kDevice, std::memory_order_release);
    return dev_;
}

CMatrix::CMatrix(const CMatrix& other)
    : nPix_(other.nPix_), data_(NULL), pinned_(false), dev_(NULL), device_(other.device_), state_(0), symNSide_(other.symNSide_),
      symStrips_(other.symStrips_), comment_(other.comment_)
{
    const int st = other.state_.load(std::memory_order_acquire);
    if(st & kHost)
    {
        allocate(other.nPix_);
        std::memcpy(data_, other.data_, static_cast<size_t>(packedSize()) * sizeof(double));
    }
    else if(st & kDevice)
    {
        // a device-resident matrix is copied on the device
        DropinLock lock(other.device_);
        void* p = NULL;
        const std::int64_t bytes = packedSize() * static_cast<std::int64_t>(sizeof(double));
        if(cmg_device_malloc(lock.ctx(), bytes, &p) != CMG_OK || cmg_copy_on_device(lock.ctx(), p, other.dev_, bytes) != CMG_OK ||
           cmg_synchronize(lock.ctx()) != CMG_OK)
        {
            const std::string text = std::string("CMatrix: ") + cmg_last_error(lock.ctx());
            if(p) cmg_device_free(lock.ctx(), p);
            raise(text);
        }
        dev_ = static_cast<double*>(p);
        state_.store(kDevice, std::memory_order_release);
    }
}

void CMatrix::swap(CMatrix& other)
{
    std::swap(nPix_, other.nPix_);
    std::swap(data_, other.data_);
    std::swap(pinned_, other.pinned_);
    std::swap(dev_, other.dev_);
    std::swap(device_, other.device_);
    const int a = state_.load(), b = other.state_.load();
    state_.store(b);
    other.state_.store(a);
    std::swap(symNSide_, other.symNSide_);
    std::swap(symStrips_, other.symStrips_);
    comment_.swap(other.comment_);
}

CMatrix& CMatrix::operator=(const CMatrix& other)
{
    if(this == &other)
        return *this;
    CMatrix copy(other);                         // may throw: *this is untouched then
    swap(copy);
    return *this;
}

CMatrix::~CMatrix() { release(); }

std::int64_t CMatrix::index(int i, int j) const
{
#ifdef CHECKS_ON
    if(i < 0 || i >= nPix_ || j < 0 || j >= nPix_)
        raise("CHECK FAILED");
#endif
    return cmg_packed_index(i, j);
}

// int32 nPix | packed doubles | int32 comment length | comment bytes   (reference source/c_matrix.cpp:41-85)
void CMatrix::writeIntoFile(const char* fileName) const
{
    packed();                                    // the host copy (made now if the matrix lives on the device)
    std::FILE* f = std::fopen(fileName, "wb");
    if(!f)
        raise(std::string("Cannot write into output file ") + fileName + ".");
    const std::int32_t n = nPix_;
    const std::int32_t len = static_cast<std::int32_t>(comment_.size());
    bool ok = std::fwrite(&n, sizeof(n), 1, f) == 1;
    // large matrices: write in slices so a single fwrite never exceeds 1 GiB
    const std::int64_t total = packedSize();
    for(std::int64_t done = 0; ok && done < total;)
    {
        const std::int64_t chunk = std::min<std::int64_t>(total - done, std::int64_t(1) << 27);
        ok = std::fwrite(data_ + done, sizeof(double), static_cast<size_t>(chunk), f) == static_cast<size_t>(chunk);
        done += chunk;
    }
    ok = ok && std::fwrite(&len, sizeof(len), 1, f) == 1;
    if(ok && len > 0)
        ok = std::fwrite(comment_.data(), 1, static_cast<size_t>(len), f) == static_cast<size_t>(len);
    std::fclose(f);
    if(!ok)
        raise(std::string("Cannot write into output file ") + fileName + ".");
}

void CMatrix::readFromFile(const char* fileName)
{
    std::FILE* f = std::fopen(fileName, "rb");
    if(!f)
        raise(std::string("Covariance matrix file ") + fileName + " cannot be read.");
    std::int32_t n = 0;
    if(std::fread(&n, sizeof(n), 1, f) != 1 || n <= 0)
    {
        std::fclose(f);
        raise(std::string("Covariance matrix file ") + fileName + " cannot be read.");
    }
    release();
    symNSide_ = 0;
    symStrips_ = 0;
    allocate(n);
    const std::int64_t total = packedSize();
    bool ok = true;
    for(std::int64_t done = 0; ok && done < total;)
    {
        const std::int64_t chunk = std::min<std::int64_t>(total - done, std::int64_t(1) << 27);
        ok = std::fread(data_ + done, sizeof(double), static_cast<size_t>(chunk), f) == static_cast<size_t>(chunk);
        done += chunk;
    }
    std::int32_t len = 0;
    comment_.clear();
    if(ok && std::fread(&len, sizeof(len), 1, f) == 1 && len > 0)
    {
        comment_.resize(static_cast<size_t>(len));
        ok = std::fread(&comment_[0], 1, static_cast<size_t>(len), f) == static_cast<size_t>(len);
    }
    std::fclose(f);
    if(!ok)
        raise(std::string("Covariance matrix file ") + fileName + " is truncated.");
}

// nPix / comment line / "i<TAB>j<TAB>value" rows with j outer, operator<< default precision
// (reference source/c_matrix.cpp:87-112)
void CMatrix::writeIntoTextFile(const char* fileName) const
{
    packed();
    std::ofstream out(fileName);
    if(!out)
        raise(std::string("Cannot write into output file ") + fileName + ".");
    out << nPix_ << std::endl;
    out << comment_ << std::endl;
    std::int64_t k = 0;
    for(int j = 0; j < nPix_; ++j)
        for(int i = 0; i <= j; ++i)
            out << i << '\t' << j << '\t' << data_[k++] << '\n';
    out.close();
}

// Reads what writeIntoTextFile writes.  (The reference's reader, source/c_matrix.cpp:114-158, tokenises on white
// space and so only picks up the first number of each row; this one parses the whole "i j value" row.)
void CMatrix::readFromTextFile(const char* fileName)
{
    std::ifstream in(fileName);
    if(!in)
        raise(std::string("Cannot read the input file ") + fileName + ".");
    int n = 0;
    in >> n;
    if(!in || n <= 0)
        raise(std::string("Cannot read the input file ") + fileName + ".");
    release();
    symNSide_ = 0;
    symStrips_ = 0;
    allocate(n);
    std::string rest;
    std::getline(in, rest);          // remainder of the first line
    std::getline(in, comment_);
    std::string line;
    while(std::getline(in, line))
    {
        if(line.empty())
            continue;
        std::stringstream row(line);
        int i = -1, j = -1;
        double v = 0;
        row >> i >> j >> v;
        if(!row)
            continue;
        if(i < 0 || i >= nPix_)
        {
            std::stringstream s;
            s << "Invalid index i = " << i << ".";
            raise(s.str());
        }
        if(j < 0 || j >= nPix_)
        {
            std::stringstream s;
            s << "Invalid index j = " << j << ".";
            raise(s.str());
        }
        data_[cmg_packed_index(i, j)] = v;
    }
}

void CMatrix::maskMatrix(const char* maskFileName)
{
    long nSide = 0;
    std::vector<int> good;
    Utils::readMask(maskFileName, nSide, good);
    const std::int64_t nPix = cmg_nside2npix(nSide);
    if(nPix != nPix_)
    {
        std::stringstream s;
        s << "The covariance matrix has " << nPix_ << " pixels, while there are " << nPix << " pixels in the mask. They need to be the same.";
        raise(s.str());
    }
    maskMatrix(good);
}

// gather of reference source/c_matrix.cpp:182-201: on the device when the matrix lives there (maskGatherKernel through
// cmg_mask_matrix, nothing crosses PCIe), on the host otherwise
void CMatrix::maskMatrix(const std::vector<int>& goodPixels)
{
    const int n = static_cast<int>(goodPixels.size());
    if(n <= 0)
        raise("the number of pixels must be positive.");
    for(int a = 0; a < n; ++a)
        if(goodPixels[a] < 0 || goodPixels[a] >= nPix_)
            raise("invalid index in goodPixels");
    const int st = state_.load(std::memory_order_acquire);
    if((st & kDevice) && !(st & kHost) && n <= 65535)
    {
        double* dOut = NULL;
        CMatrix* reduced = newOnDevice(n, device_, &dOut);
        DropinLock lock(device_);
        const cmg_status s = cmg_mask_matrix(lock.ctx(), dev_, nPix_, &goodPixels[0], n, dOut);
        if(s != CMG_OK || cmg_synchronize(lock.ctx()) != CMG_OK)
        {
            const std::string text = std::string("CMatrix::maskMatrix: ") + cmg_last_error(lock.ctx());
            delete reduced;
            raise(text);
        }
        reduced->comment_ = comment_;
        swap(*reduced);
        delete reduced;
        return;
    }
    const double* src = packed();
    CMatrix reduced(n);
    std::int64_t k = 0;
    for(int b = 0; b < n; ++b)
        for(int a = 0; a <= b; ++a)
            reduced.data_[k++] = src[cmg_packed_index(goodPixels[a], goodPixels[b])];
    reduced.comment_ = comment_;
    swap(reduced);
}
