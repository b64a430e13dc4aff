/* test infrastructure: see healpix_stub_all.h */
#ifndef ORACLE_STUB_ALM_HEALPIX_TOOLS_H
#define ORACLE_STUB_ALM_HEALPIX_TOOLS_H
#include "healpix_stub_all.h"
#endif
