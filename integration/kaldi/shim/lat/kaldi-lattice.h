// Stand-in for src/lat/kaldi-lattice.h (OpenFst lattices; OpenFst is not in this
// image).  src/ctc/ctc-nnet-example.h:27 includes it for the denominator lattice
// of DiscriminativeNnetCtcExample, which is not on the CTC training path.  The
// type exists so that the reference's ctc-nnet-example.cc compiles unmodified;
// lattice I/O reports an error (integration/kaldi/harness/lattice_stub.cc).
#ifndef B200_SHIM_KALDI_LATTICE_H_
#define B200_SHIM_KALDI_LATTICE_H_
#include <iostream>
#include <vector>
#include "base/kaldi-common.h"
namespace kaldi {
struct CompactLattice { int NumStates() const { return 0; } };
struct Lattice { };
bool WriteCompactLattice(std::ostream &os, bool binary, const CompactLattice &clat);
bool ReadCompactLattice(std::istream &is, bool binary, CompactLattice **clat);
}
#endif
