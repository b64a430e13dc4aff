// Drop-in replacement of the reference's include/c_matrix.hpp (class CMatrix, :10-83): a symmetric
// nPix x nPix covariance matrix in pixel space, stored as the packed upper triangle with entry (i <= j) at
// j(j+1)/2 + i (reference source/c_matrix.cpp:27-39 -- LAPACK 'U' packed order).  Same public members and
// meaning; what differs underneath:
//   * indices are 64-bit inside (the reference's int arithmetic overflows beyond nPix = 46340,
//     source/c_matrix.cpp:24,36), so the 147456-dimensional T,Q,U matrix of Nside=64 is representable;
//   * the storage is page-locked host memory when a GPU is present, so the generators in
//     c_matrix_generator.hpp can copy their device result straight into it.
// File formats are byte-compatible with the reference (source/c_matrix.cpp:41-158).
#ifndef COSMO_PP_C_MATRIX_HPP
#define COSMO_PP_C_MATRIX_HPP

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
    double& element(int i, int j) { return data_[index(i, j)]; }
    double element(int i, int j) const { return data_[index(i, j)]; }

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
    // the packed triangle itself, nPix (nPix + 1) / 2 doubles
    double* packed() { return data_; }
    const double* packed() const { return data_; }
    std::int64_t packedSize() const { return static_cast<std::int64_t>(nPix_) * (nPix_ + 1) / 2; }

private:
    std::int64_t index(int i, int j) const;
    void allocate(int nPix);
    void release();

    int nPix_;
    double* data_;
    bool pinned_;
    std::string comment_;
};

#endif
