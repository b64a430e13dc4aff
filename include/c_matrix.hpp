// Drop-in replacement of the reference's include/c_matrix.hpp (class CMatrix, :10-83): a symmetric
// nPix x nPix covariance matrix in pixel space, stored as the packed upper triangle with entry (i <= j) at
// j(j+1)/2 + i (reference source/c_matrix.cpp:27-39 -- LAPACK 'U' packed order).  Same public members and
// meaning; what differs underneath:
//   * indices are 64-bit inside (the reference's int arithmetic overflows beyond nPix = 46340,
//     source/c_matrix.cpp:24,36), so the 147456-dimensional T,Q,U matrix of Nside=64 is representable;
//   * the object is a HANDLE over two copies of the packed triangle, either of which may be absent: one in (page-locked)
//     host memory and one in the memory of a GPU.  The generators of c_matrix_generator.hpp leave their result on the
//     device; the host copy is made the first time something asks for it (element(), packed(), the file writers), over
//     PCIe once.  The consumers of this library (Likelihood, PixelLikelihoodTT, maskMatrix) take the device copy, so the
//     sequence of reference source/test_like_low.cpp:181-191 -- generate, mask, factorise, evaluate -- moves no matrix
//     across PCIe at all.  Writing through element() / packed() makes the host copy the only valid one.
// File formats are byte-compatible with the reference (source/c_matrix.cpp:41-158).
#ifndef COSMO_PP_C_MATRIX_HPP
#define COSMO_PP_C_MATRIX_HPP

#include <atomic>
#include <cstdint>
#include <string>
#include <vector>

class CMatrix
{
public:
    // zero matrix of nPix pixels; throws StandardException unless nPix > 0
    CMatrix(int nPix);
    // read from a binary file written by writeIntoFile
    CMatrix(const char* fileName);
    CMatrix(const CMatrix& other);
    CMatrix& operator=(const CMatrix& other);
    ~CMatrix();

    // element (i, j) == element (j, i)
    double& element(int i, int j) { if(state_.load(std::memory_order_acquire) != kHost) hostForWrite(); return data_[index(i, j)]; }
    double element(int i, int j) const { if(!(state_.load(std::memory_order_acquire) & kHost)) hostForRead(); return data_[index(i, j)]; }

    void readFromFile(const char* fileName);
    void readFromTextFile(const char* fileName);
    void writeIntoFile(const char* fileName) const;
    void writeIntoTextFile(const char* fileName) const;

    int getNPix() const { return nPix_; }

    const std::string& comment() const { return comment_; }
    std::string& comment() { return comment_; }

    // keep only the unmasked pixels: new(a, b) = old(goodPixels[a], goodPixels[b])
    void maskMatrix(const char* maskFileName);
    void maskMatrix(const std::vector<int>& goodPixels);

    // ---- additions (not in the reference) ----
    // the packed triangle itself in host memory, nPix (nPix + 1) / 2 doubles (materialised on first use; the non-const
    // form invalidates the device copy)
    double* packed() { if(state_.load(std::memory_order_acquire) != kHost) hostForWrite(); return data_; }
    const double* packed() const { if(!(state_.load(std::memory_order_acquire) & kHost)) hostForRead(); return data_; }
    std::int64_t packedSize() const { return static_cast<std::int64_t>(nPix_) * (nPix_ + 1) / 2; }
    // the packed triangle in the memory of GPU `device` (uploaded on first use when only the host copy exists); valid until
    // the matrix is written to, masked, assigned or destroyed
    const double* devicePacked(int device) const;
    bool hasHostCopy() const { return (state_.load(std::memory_order_acquire) & kHost) != 0; }
    bool hasDeviceCopy() const { return (state_.load(std::memory_order_acquire) & kDevice) != 0; }
    // a matrix of nPix pixels that so far exists only in the memory of GPU `device` (what the generators return); *dPacked
    // receives the buffer to fill.  fullSkyNSide / strips (1 = [T], 3 = [T;Q;U]) tell the lazy host copy that the matrix
    // has the rotation symmetry of the full NESTED sky, 0 = no such promise.
    static CMatrix* newOnDevice(int nPix, int device, double** dPacked, long fullSkyNSide = 0, int strips = 0);

private:
    enum { kHost = 1, kDevice = 2 };
    std::int64_t index(int i, int j) const;
    void allocate(int nPix);
    void release();
    void hostForRead() const;
    void hostForWrite();
    void dropDevice() const;
    void swap(CMatrix& other);
    struct DeviceOnly {};
    CMatrix(DeviceOnly, int nPix);

    int nPix_;
    mutable double* data_;
    mutable bool pinned_;
    mutable double* dev_;
    mutable int device_;
    mutable std::atomic<int> state_;
    long symNSide_;
    int symStrips_;
    std::string comment_;
};

#endif
