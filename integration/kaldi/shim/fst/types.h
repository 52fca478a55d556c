// Stand-in for OpenFst's <fst/types.h>, the only OpenFst header that
// src/base/kaldi-types.h:44 needs (OpenFst is a third-party dependency that is
// not in this image).  Used by the in-container builds of the reference's sources
// (integration/kaldi/check_compile.sh, oracle/ref/Makefile).
#ifndef B200_SHIM_FST_TYPES_H_
#define B200_SHIM_FST_TYPES_H_
#include <stdint.h>
typedef int8_t int8; typedef int16_t int16; typedef int32_t int32; typedef int64_t int64;
typedef uint8_t uint8; typedef uint16_t uint16; typedef uint32_t uint32; typedef uint64_t uint64;
// OpenFst's compat.h macro that src/hmm/transition-model.h:308 relies on
#ifndef DISALLOW_COPY_AND_ASSIGN
#define DISALLOW_COPY_AND_ASSIGN(type) type(const type&); void operator=(const type&)
#endif
#endif
