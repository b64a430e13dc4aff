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
    // page-locked when a GPU is there (the generators copy device results straight into it)
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
}

void CMatrix::release()
{
    if(!data_)
        return;
    if(pinned_)
        cmg_host_free_pinned(data_);
    else
        std::free(data_);
    data_ = NULL;
}

CMatrix::CMatrix(int nPix) : nPix_(0), data_(NULL), pinned_(false) { allocate(nPix); }

CMatrix::CMatrix(const char* fileName) : nPix_(0), data_(NULL), pinned_(false) { readFromFile(fileName); }

CMatrix::CMatrix(const CMatrix& other) : nPix_(0), data_(NULL), pinned_(false), comment_(other.comment_)
{
    allocate(other.nPix_);
    std::memcpy(data_, other.data_, static_cast<size_t>(packedSize()) * sizeof(double));
}

CMatrix& CMatrix::operator=(const CMatrix& other)
{
    if(this == &other)
        return *this;
    release();
    allocate(other.nPix_);
    std::memcpy(data_, other.data_, static_cast<size_t>(packedSize()) * sizeof(double));
    comment_ = other.comment_;
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

// gather of reference source/c_matrix.cpp:182-201, on the host: the object lives in host memory
void CMatrix::maskMatrix(const std::vector<int>& goodPixels)
{
    const int n = static_cast<int>(goodPixels.size());
    if(n <= 0)
        raise("the number of pixels must be positive.");
    for(int a = 0; a < n; ++a)
        if(goodPixels[a] < 0 || goodPixels[a] >= nPix_)
            raise("invalid index in goodPixels");
    CMatrix reduced(n);
    std::int64_t k = 0;
    for(int b = 0; b < n; ++b)
        for(int a = 0; a <= b; ++a)
            reduced.data_[k++] = data_[cmg_packed_index(goodPixels[a], goodPixels[b])];
    reduced.comment_ = comment_;
    std::swap(nPix_, reduced.nPix_);
    std::swap(data_, reduced.data_);
    std::swap(pinned_, reduced.pinned_);
}
