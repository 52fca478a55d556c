// Stand-in for src/lat/lattice-functions.h (see lat/kaldi-lattice.h in this directory).
#ifndef B200_SHIM_LATTICE_FUNCTIONS_H_
#define B200_SHIM_LATTICE_FUNCTIONS_H_
#include "lat/kaldi-lattice.h"
namespace kaldi { int32 CompactLatticeStateTimes(const CompactLattice &clat, std::vector<int32> *times); }
#endif
