// integration/kaldi/harness/lattice_stub.cc -- definitions for the lattice I/O that
// src/ctc/ctc-nnet-example.cc:150,177,201 references for DiscriminativeNnetCtcExample.
// OpenFst is not in this image (integration/kaldi/shim/lat/kaldi-lattice.h); discriminative
// examples are not on the CTC training path, so these report an error if ever reached.
#include "lat/lattice-functions.h"

namespace kaldi {

bool WriteCompactLattice(std::ostream &, bool, const CompactLattice &) {
  KALDI_ERR << "lattice I/O is not available in this build (no OpenFst)";
  return false;
}

bool ReadCompactLattice(std::istream &, bool, CompactLattice **) {
  KALDI_ERR << "lattice I/O is not available in this build (no OpenFst)";
  return false;
}

int32 CompactLatticeStateTimes(const CompactLattice &, std::vector<int32> *) {
  KALDI_ERR << "lattice functions are not available in this build (no OpenFst)";
  return 0;
}

}  // namespace kaldi
