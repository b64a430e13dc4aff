// shared by the drop-in classes: the process-wide generator context (one per device, created on first use)
#ifndef COSMO_PP_B200_DROPIN_INTERNAL_HPP
#define COSMO_PP_B200_DROPIN_INTERNAL_HPP

#include <cmg.h>

#include <vector>

cmg_ctx* cmgDropinContext();          // throws StandardException when no sm_100 GPU is usable
// pixel window of nSide up to lMax as CMatrixGenerator resolves it (setPixelWindow / HEALPix data directory)
void cmgDropinPixelWindow(long nSide, int lMax, bool polarization, std::vector<double>& w);

#endif
