// Stand-in for src/fstext/fstext-lib.h, which pulls in all of OpenFst (not in this
// image).  src/ctc/ctc-transition-model.h:23 includes it but uses nothing from it.
#ifndef B200_SHIM_FSTEXT_LIB_H_
#define B200_SHIM_FSTEXT_LIB_H_
#include "fst/fst-decl.h"
#endif
