// Utils of include/utils.hpp: the file formats on the input side of the C-matrix path, without cfitsio / HEALPix.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include <cmg.h>
#include <exception_handler.hpp>
#include <utils.hpp>

namespace
{
[[noreturn]] void raise(const std::string& text) { throw StandardException(text); }

const int kBlock = 2880, kCard = 80;

struct Header
{
    std::vector<std::pair<std::string, std::string> > cards;
    std::string get(const std::string& key) const
    {
        for(size_t i = 0; i < cards.size(); ++i)
            if(cards[i].first == key)
                return cards[i].second;
        return "";
    }
    long getLong(const std::string& key, long dflt) const
    {
        const std::string v = get(key);
        return v.empty() ? dflt : std::atol(v.c_str());
    }
};

std::string trim(const std::string& s)
{
    size_t a = 0, b = s.size();
    while(a < b && std::isspace(static_cast<unsigned char>(s[a]))) ++a;
    while(b > a && std::isspace(static_cast<unsigned char>(s[b - 1]))) --b;
    return s.substr(a, b - a);
}

// reads header blocks until END; returns false at end of file
bool readHeader(std::FILE* f, Header& h)
{
    h.cards.clear();
    char block[kBlock];
    bool end = false;
    while(!end)
    {
        if(std::fread(block, 1, kBlock, f) != static_cast<size_t>(kBlock))
            return false;
        for(int c = 0; c < kBlock / kCard && !end; ++c)
        {
            const std::string card(block + c * kCard, kCard);
            const std::string key = trim(card.substr(0, 8));
            if(key == "END")
            {
                end = true;
                break;
            }
            if(card.size() < 10 || card[8] != '=')
                continue;
            std::string val = card.substr(10);
            if(!val.empty() && trim(val)[0] == '\'')
            {
                const size_t a = val.find('\'');
                const size_t b = val.find('\'', a + 1);
                val = trim(val.substr(a + 1, b == std::string::npos ? std::string::npos : b - a - 1));
            }
            else
            {
                const size_t slash = val.find('/');
                val = trim(val.substr(0, slash));
            }
            h.cards.push_back(std::make_pair(key, val));
        }
    }
    return true;
}

void skipData(std::FILE* f, const Header& h)
{
    const long bitpix = std::labs(h.getLong("BITPIX", 8));
    const long naxis = h.getLong("NAXIS", 0);
    std::int64_t n = naxis > 0 ? 1 : 0;
    for(long a = 1; a <= naxis; ++a)
    {
        std::stringstream k;
        k << "NAXIS" << a;
        n *= h.getLong(k.str(), 0);
    }
    std::int64_t bytes = (h.getLong("PCOUNT", 0) + n) * (bitpix / 8) * std::max<long>(1, h.getLong("GCOUNT", 1));
    bytes = (bytes + kBlock - 1) / kBlock * kBlock;
    std::fseek(f, static_cast<long>(bytes), SEEK_CUR);
}

double bigEndianDouble(const unsigned char* p)
{
    std::uint64_t v = 0;
    for(int b = 0; b < 8; ++b) v = (v << 8) | p[b];
    double d;
    std::memcpy(&d, &v, 8);
    return d;
}

float bigEndianFloat(const unsigned char* p)
{
    std::uint32_t v = 0;
    for(int b = 0; b < 4; ++b) v = (v << 8) | p[b];
    float d;
    std::memcpy(&d, &v, 4);
    return d;
}
}

void Utils::readFitsTable(const char* fileName, FitsTable& table)
{
    std::FILE* f = std::fopen(fileName, "rb");
    if(!f)
        raise(std::string("Cannot open the FITS file ") + fileName + ".");
    Header h;
    if(!readHeader(f, h) || h.get("SIMPLE").empty())
    {
        std::fclose(f);
        raise(std::string("The file ") + fileName + " is not a FITS file.");
    }
    skipData(f, h);
    bool found = false;
    while(readHeader(f, h))
    {
        if(h.get("XTENSION") == "BINTABLE")
        {
            found = true;
            break;
        }
        skipData(f, h);
    }
    if(!found)
    {
        std::fclose(f);
        raise(std::string("The FITS file ") + fileName + " has no binary table extension.");
    }
    const long rowBytes = h.getLong("NAXIS1", 0), rows = h.getLong("NAXIS2", 0), fields = h.getLong("TFIELDS", 0);
    table.columns.assign(static_cast<size_t>(fields), std::vector<double>());
    table.columnType.assign(static_cast<size_t>(fields), ' ');
    table.ordering = h.get("ORDERING");
    std::transform(table.ordering.begin(), table.ordering.end(), table.ordering.begin(), ::toupper);
    table.nSide = h.getLong("NSIDE", 0);
    std::vector<long> repeat(static_cast<size_t>(fields), 1), width(static_cast<size_t>(fields), 0), offset(static_cast<size_t>(fields), 0);
    std::vector<double> scale, zero;
    long off = 0;
    for(long c = 0; c < fields; ++c)
    {
        std::stringstream k;
        k << "TFORM" << (c + 1);
        const std::string form = h.get(k.str());
        size_t pos = 0;
        while(pos < form.size() && std::isdigit(static_cast<unsigned char>(form[pos]))) ++pos;
        repeat[c] = pos ? std::atol(form.substr(0, pos).c_str()) : 1;
        const char type = pos < form.size() ? form[pos] : ' ';
        table.columnType[c] = type;
        switch(type)
        {
        case 'D': case 'K': width[c] = 8; break;
        case 'E': case 'J': width[c] = 4; break;
        case 'I': width[c] = 2; break;
        case 'B': case 'L': case 'A': width[c] = 1; break;
        default:
            std::fclose(f);
            raise(std::string("Unsupported TFORM '") + form + "' in " + fileName + ".");
        }
        offset[c] = off;
        off += repeat[c] * width[c];
        // a scaled column (value = TZERO + TSCAL * stored) would be read wrongly without an error: apply the scaling
        std::stringstream ks, kz;
        ks << "TSCAL" << (c + 1);
        kz << "TZERO" << (c + 1);
        scale.push_back(h.get(ks.str()).empty() ? 1.0 : std::atof(h.get(ks.str()).c_str()));
        zero.push_back(h.get(kz.str()).empty() ? 0.0 : std::atof(h.get(kz.str()).c_str()));
    }
    if(off != rowBytes)
    {
        std::fclose(f);
        raise(std::string("Inconsistent binary table row length in ") + fileName + ".");
    }
    std::vector<unsigned char> row(static_cast<size_t>(rowBytes));
    for(long c = 0; c < fields; ++c)
        table.columns[c].reserve(static_cast<size_t>(rows * repeat[c]));
    for(long r = 0; r < rows; ++r)
    {
        if(std::fread(&row[0], 1, static_cast<size_t>(rowBytes), f) != static_cast<size_t>(rowBytes))
        {
            std::fclose(f);
            raise(std::string("The FITS file ") + fileName + " is truncated.");
        }
        for(long c = 0; c < fields; ++c)
            for(long e = 0; e < repeat[c]; ++e)
            {
                const unsigned char* p = &row[static_cast<size_t>(offset[c] + e * width[c])];
                double v = 0;
                switch(table.columnType[c])
                {
                case 'D': v = bigEndianDouble(p); break;
                case 'E': v = bigEndianFloat(p); break;
                case 'J': v = static_cast<std::int32_t>((std::uint32_t(p[0]) << 24) | (std::uint32_t(p[1]) << 16) | (std::uint32_t(p[2]) << 8) | p[3]); break;
                case 'I': v = static_cast<std::int16_t>((p[0] << 8) | p[1]); break;
                case 'K':
                {
                    std::uint64_t u = 0;
                    for(int b = 0; b < 8; ++b)
                        u = (u << 8) | p[b];
                    v = static_cast<double>(static_cast<std::int64_t>(u));
                    break;
                }
                case 'L': v = (p[0] == 'T') ? 1.0 : 0.0; break;             // logical: 'T' / 'F'
                default: v = p[0]; break;                                   // 'B' unsigned byte, 'A' character code
                }
                table.columns[c].push_back(zero[c] + scale[c] * v);
            }
    }
    std::fclose(f);
}

void Utils::readMask(const char* maskFileName, long& nSide, std::vector<int>& goodPixels)
{
    FitsTable t;
    readFitsTable(maskFileName, t);
    if(t.columns.empty())
        raise(std::string("The mask file ") + maskFileName + " has no columns.");
    if(t.ordering != "NESTED" && t.ordering != "NEST")
        raise("The mask must have nested ordering.");       // text of reference source/utils.cpp:34
    const std::vector<double>& m = t.columns[0];
    const std::int64_t nPix = static_cast<std::int64_t>(m.size());
    long ns = t.nSide;
    if(ns <= 0)
        ns = static_cast<long>(std::llround(std::sqrt(static_cast<double>(nPix) / 12.0)));
    if(cmg_nside2npix(ns) != nPix)
        raise(std::string("The mask file ") + maskFileName + " does not hold 12 nSide^2 pixels.");
    nSide = ns;
    goodPixels.resize(static_cast<size_t>(nPix));
    std::int64_t n = 0;
    if(cmg_good_pixels_from_mask(&m[0], nPix, &goodPixels[0], &n) != CMG_OK)
        raise("invalid mask");
    goodPixels.resize(static_cast<size_t>(n));
}

double Utils::beamFunction(int l, double fwhm) { return cmg_beam_function(l, fwhm); }

void Utils::readClFromFile(const char* fileName, std::vector<double>& cl, bool hasL, bool isDl)
{
    std::ifstream in(fileName);
    if(!in)
        raise(std::string("Cannot open the input file ") + fileName + ".");
    cl.clear();
    std::string line;
    int l = 0;
    const double pi = 3.141592653589793;
    while(std::getline(in, line))
    {
        if(line.empty())
            break;
        std::stringstream row(line);
        if(hasL)
        {
            int fileL = -1;
            row >> fileL;
            if(fileL != l)
            {
                std::stringstream s;
                s << "Invalid format of the input file " << fileName << ". Expected to read l = " << l << " but found l = " << fileL << ".";
                raise(s.str());
            }
        }
        double v = 0;
        row >> v;
        if(isDl && l)
            v *= (2 * pi / (l * (l + 1)));
        cl.push_back(v);
        ++l;
    }
}
