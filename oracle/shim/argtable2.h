/* Stand-in for argtable2.h when compiling the reference's own sources into oracle/_ref (TEST
 * INFRASTRUCTURE): the option-table parser of the drop-in CLI has the same shape, so it is reused. */
#ifndef SHIM_ARGTABLE2_H
#define SHIM_ARGTABLE2_H
#include "../../msamtools_b200/csrc/host/margs.h"
#endif
