/* test infrastructure: see healpix_stub_all.h */
#ifndef ORACLE_STUB_ARR_H
#define ORACLE_STUB_ARR_H
#include "healpix_stub_all.h"
#endif
