// Stand-in for OpenFst's <fst/types.h>, the only OpenFst header that
// src/base/kaldi-types.h:44 needs; used by the in-container compile check only.
#ifndef B200_SHIM_FST_TYPES_H_
#define B200_SHIM_FST_TYPES_H_
#include <stdint.h>
typedef int8_t int8; typedef int16_t int16; typedef int32_t int32; typedef int64_t int64;
typedef uint8_t uint8; typedef uint16_t uint16; typedef uint32_t uint32; typedef uint64_t uint64;
#endif
