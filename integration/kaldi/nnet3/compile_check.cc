// Syntax/type check of the nnet3 adapter against the reference's own nnet3 headers
// (integration/kaldi/check_compile.sh): instantiates the class so every virtual is checked.
#include "nnet-b200-recurrent-component.h"

kaldi::nnet3::Component *MakeB200RecurrentComponent() { return new kaldi::nnet3::B200RecurrentComponent(); }
