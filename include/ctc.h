/*
 * include/ctc.h -- drop-in for the warp-ctc header that kaldi-ctc includes as
 * `extern "C" { #include "ctc.h" }` at src/ctc/ctc-nnet-update.cc:27-29.
 *
 * Every name the reference uses at that call site is declared here with the
 * same meaning, so ctc-nnet-update.cc compiles unmodified against it:
 *   ctcStatus_t / ctcGetStatusString        ctc-nnet-update.cc:31-37
 *   ctcOptions {loc, stream, blank_label}   ctc-nnet-update.cc:204-209
 *   CTC_GPU                                 ctc-nnet-update.cc:207
 *   get_workspace_size                      ctc-nnet-update.cc:211-214
 *   compute_ctc_loss                        ctc-nnet-update.cc:224-231,236-243
 * Install as <root>/include/ctc.h next to <root>/build/libwarpctc.so (a copy or
 * symlink of libb200ctc.so), which is what `configure --warpctc-root=<root>`
 * looks for (src/configure:499-541).  See INTEGRATION.md.
 *
 * Implementation: kaldi_ctc_b200/csrc/ctc.cu (hand-written CUDA, sm_100a).
 * There is no CPU implementation behind this header: loc == CTC_CPU returns
 * CTC_STATUS_INVALID_VALUE.
 */
#ifndef B200_WARPCTC_COMPAT_CTC_H_
#define B200_WARPCTC_COMPAT_CTC_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* CUDA's stream handle, declared here so callers need no CUDA header. */
typedef struct CUstream_st *CUstream;

typedef enum {
  CTC_STATUS_SUCCESS = 0,
  CTC_STATUS_MEMOPS_FAILED = 1,
  CTC_STATUS_INVALID_VALUE = 2,
  CTC_STATUS_EXECUTION_FAILED = 3,
  CTC_STATUS_UNKNOWN_ERROR = 4
} ctcStatus_t;

typedef enum { CTC_CPU = 0, CTC_GPU = 1 } ctcComputeLocation;

struct ctcOptions {
  ctcComputeLocation loc; /* must be CTC_GPU */
  union {
    unsigned int num_threads; /* CTC_CPU only: unsupported here */
    CUstream stream;          /* CTC_GPU: stream the work is enqueued on */
  };
  int blank_label; /* the reference passes 0 (ctc-nnet-update.cc:205) */
};
#ifndef __cplusplus
typedef struct ctcOptions ctcOptions;
#endif

/* Library version (warp-ctc exports the same symbol). */
int get_warpctc_version(void);

/* Static, NUL-terminated description of a status code. */
const char *ctcGetStatusString(ctcStatus_t status);

/*
 * Bytes of DEVICE scratch compute_ctc_loss needs for this minibatch.
 * label_lengths / input_lengths: HOST arrays of `minibatch` ints.
 */
ctcStatus_t get_workspace_size(const int *const label_lengths,
                               const int *const input_lengths,
                               int alphabet_size, int minibatch,
                               struct ctcOptions options, size_t *size_bytes);

/*
 * CTC negative log-likelihood and its gradient w.r.t. the UN-normalised
 * activations (the softmax is applied inside).
 *   activations  DEVICE [T, minibatch, alphabet_size] fp32, index
 *                (t*minibatch + b)*alphabet_size + k, T = max(input_lengths)
 *   gradients    DEVICE, same shape, or NULL for loss only.  Every row is
 *                written: rows with t >= input_lengths[b] are set to 0.
 *   flat_labels  HOST, concatenated label sequences (no blanks)
 *   label_lengths, input_lengths   HOST [minibatch]
 *   costs        HOST [minibatch], -log p(labels | activations) per utterance;
 *                valid on return (the call synchronises options.stream)
 *   workspace    DEVICE, >= get_workspace_size() bytes, 256-byte aligned
 * An utterance whose labels cannot be aligned (L + #repeats > T) gets cost 0
 * and a zero gradient, as in warp-ctc.
 */
ctcStatus_t compute_ctc_loss(const float *const activations, float *gradients,
                             const int *const flat_labels,
                             const int *const label_lengths,
                             const int *const input_lengths, int alphabet_size,
                             int minibatch, float *costs, void *workspace,
                             struct ctcOptions options);

#ifdef __cplusplus
}
#endif
#endif /* B200_WARPCTC_COMPAT_CTC_H_ */
