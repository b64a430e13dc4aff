/* test infrastructure: see healpix_stub_all.h */
#ifndef ORACLE_STUB_XCOMPLEX_H
#define ORACLE_STUB_XCOMPLEX_H
#include "healpix_stub_all.h"
#endif
