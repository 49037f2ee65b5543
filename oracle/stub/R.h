/* Stand-in for <R.h>; see Rinternals.h in this directory. Test infrastructure only. */
#ifndef EDB200_STUB_R_H
#define EDB200_STUB_R_H
#include "Rinternals.h"
#endif
