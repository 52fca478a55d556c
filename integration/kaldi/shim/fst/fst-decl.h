// Stand-in for OpenFst's <fst/fst-decl.h> (forward declarations only), included by
// src/hmm/transition-model.h:27.  OpenFst is not in this image.
#ifndef B200_SHIM_FST_DECL_H_
#define B200_SHIM_FST_DECL_H_
#include "fst/types.h"
namespace fst {
template <class A> class VectorFst;
template <class W> class ArcTpl;
template <class T> class TropicalWeightTpl;
typedef ArcTpl<TropicalWeightTpl<float> > StdArc;
typedef VectorFst<StdArc> StdVectorFst;
}
#endif
