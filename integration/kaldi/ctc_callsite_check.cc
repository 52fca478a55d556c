// Compile check of include/ctc.h against the way kaldi-ctc uses it: the same include
// form (extern "C" around "ctc.h", src/ctc/ctc-nnet-update.cc:27-29), the same
// option fields (:204-209) and the same three calls (:211-214, :224-231, :236-243).
// Written for this repo (ctc-nnet-update.cc itself needs OpenFst headers to compile).
#include <cstddef>
#include <vector>
extern "C" {
#include "ctc.h"
}
typedef struct CUstream_st *cudaStream_t;

int CallSiteShape(const float *output, float *deriv, std::vector<int> &flat_labels, std::vector<int> &label_lengths,
                  std::vector<int> &input_lengths, int alphabet_size, int mini_batch, float *costs,
                  cudaStream_t stream) {
  ctcOptions info;
  info.blank_label = 0;
  info.loc = CTC_GPU;
  info.stream = stream;
  size_t gpu_alloc_bytes;
  ctcStatus_t ret;
  if ((ret = get_workspace_size(label_lengths.data(), input_lengths.data(), alphabet_size, mini_batch, info,
                                &gpu_alloc_bytes)) != 0)
    return (int)ctcGetStatusString(ret)[0];
  char *ctc_gpu_workspace = 0;
  if ((ret = compute_ctc_loss(output, deriv, flat_labels.data(), label_lengths.data(), input_lengths.data(),
                              alphabet_size, mini_batch, costs, ctc_gpu_workspace, info)) != 0)
    return 1;
  if ((ret = compute_ctc_loss(output, NULL, flat_labels.data(), label_lengths.data(), input_lengths.data(),
                              alphabet_size, mini_batch, costs, ctc_gpu_workspace, info)) != 0)
    return 2;
  return 0;
}
