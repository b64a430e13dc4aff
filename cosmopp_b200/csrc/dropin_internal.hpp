// shared by the drop-in classes: the process-wide generator context (one per device, created on first use)
#ifndef COSMO_PP_B200_DROPIN_INTERNAL_HPP
#define COSMO_PP_B200_DROPIN_INTERNAL_HPP

#include <cmg.h>

cmg_ctx* cmgDropinContext();          // throws StandardException when no sm_100 GPU is usable

#endif
