/* test infrastructure: see healpix_stub_all.h */
#ifndef ORACLE_STUB_HEALPIX_MAP_FITSIO_H
#define ORACLE_STUB_HEALPIX_MAP_FITSIO_H
#include "healpix_stub_all.h"
#endif
