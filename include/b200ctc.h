/*
 * include/b200ctc.h -- extended entry points of libb200ctc.so for callers that
 * want to drop the per-minibatch overheads around the warp-ctc call in
 * NnetCtcUpdater::ComputeObjfAndDeriv (src/ctc/ctc-nnet-update.cc:171-259):
 *   - cudaStreamCreate/Destroy + cudaMalloc/cudaFree per minibatch (:209-217,
 *     247-248)  -> caller-owned workspace, any stream, optional no-sync
 *   - deriv->Scale(-1) in NnetCtcUpdater::Backprop (:323) -> grad_scale
 *   - deriv->Sum() NaN check (:232-234) and the costs.Sum()==costs.Sum() assert (:254) -> a flag word the
 *     kernels set (bit 0: a non-finite cost, bit 1: a frame without a usable posterior), delivered through
 *     b200ctcOptions.nonfinite_dev so that the caller can skip the update of a bad minibatch without a sync
 *   - FindRowMaxId for the accuracy (:270-273), a third read of the slab -> argmax_dev
 * Same data layout and semantics as include/ctc.h.
 */
#ifndef B200CTC_H_
#define B200CTC_H_

#include "ctc.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int blank_label;    /* 0 in kaldi-ctc */
  float grad_scale;   /* gradients are multiplied by this (1 = warp-ctc, -1 =
                         what NnetCtcUpdater::Backprop feeds the network) */
  CUstream stream;    /* stream to enqueue on */
  int no_sync;        /* 1: do not synchronise; costs_host may then be NULL and
                         costs are only available in costs_dev */
  int *argmax_dev;    /* optional DEVICE [T*minibatch] ints: arg-max symbol of every valid row
                         (what FindRowMaxId gives NnetCtcUpdater::ComputeTotAccuracy,
                         ctc-nnet-update.cc:270-273), produced by the pass that already
                         streams the row; rows past input_lengths get -1.  NULL: skipped. */
  int *nonfinite_dev; /* optional DEVICE int: receives the call's flag word (0 = clean; bit 0 = a cost
                         is inf/nan, bit 1 = some frame had no usable posterior).  Replaces the
                         reference's blocking deriv->Sum() / costs.Sum() checks (:232-234, :254): a
                         caller gates its weight update on it (b200rnnClipAndUpdateGuarded). */
} b200ctcOptions;

/* Same as get_workspace_size. */
ctcStatus_t b200ctc_workspace_size(const int *label_lengths,
                                   const int *input_lengths, int alphabet_size,
                                   int minibatch, size_t *size_bytes);

/*
 * costs_host (HOST, may be NULL when no_sync) and costs_dev (DEVICE, may be
 * NULL) both receive the per-utterance NLL.  Everything else as
 * compute_ctc_loss.
 */
ctcStatus_t b200ctc_loss(const float *activations, float *gradients,
                         const int *flat_labels, const int *label_lengths,
                         const int *input_lengths, int alphabet_size,
                         int minibatch, float *costs_host, float *costs_dev,
                         void *workspace, size_t workspace_bytes,
                         b200ctcOptions options);

/* Algorithmic HBM bytes of one call:
 * 4*A*sum_b T_b + 4*A*T_max*B + 4*sum L_b + 4*B. */
size_t b200ctc_algorithmic_bytes(const int *label_lengths,
                                 const int *input_lengths, int alphabet_size,
                                 int minibatch);

/* Test / tuning hook, not needed in production: overrides one of the library's internal launch choices for
 * the rest of the process (the same switches the B200CTC_<KEY> environment variables set at load time:
 * "RING" -1/0/1, "GROUPS", "P", "NA", "NT", "PROFILE", "ONE_STREAM").  Every setting computes the
 * same function; the parity tests use it to force each kernel variant.  Returns 0, or -1 for an unknown key. */
int b200ctc_set_tuning(const char *key, int value);

/* Number of kernels one b200ctc_loss call launches (for bench.py's count). */
int b200ctc_launches_per_call(int with_gradients);

/*
 * Decoding side: everything CtcDecodableAmNnet's constructor does to the
 * network output after NnetComputation (src/ctc/ctc-decodable-am-nnet.cc:54-86), and
 * CtcDecodableAmNnetParallel::Compute's variant (:89-108), for `minibatch` utterances at once:
 *   keep frame t of utterance u  iff  p(blank | t,u) < blank_threshold     (:54-60; all frames when
 *     blank_threshold >= 1, and all frames of an utterance none of whose frames pass, :62-63)
 *   log_probs[kept row][k] = prob_scale * (log(max(p[k], floor_value)) - log(priors[k]))   (:72-83)
 * nnet_output: DEVICE [T_max*minibatch, alphabet_size], row t*minibatch+u (minibatch=1 is the
 *   reference's per-utterance matrix).  input_is_logits=1: the rows are the affine layer's output and
 *   the appended SoftmaxComponent (steps/ctc/train.sh:471-476) is folded into the same pass;
 *   0: the rows are already probabilities.
 * priors: DEVICE [alphabet_size] raw priors (AmNnet::Priors()) or NULL (Priors().Dim()==0, :76).
 * floor_value: 1e-10 (constructor, :72) or 1e-20 (Parallel::Compute, :93).
 * log_probs: DEVICE [sum_u input_lengths[u], alphabet_size]; utterance u's kept rows start at row
 *   sum_{v<u} input_lengths[v] and number kept[u] (<= input_lengths[u]); the rest is untouched.
 * kept_dev (DEVICE, may be NULL) / kept_host (HOST, may be NULL; non-NULL synchronises the stream).
 */
ctcStatus_t b200ctc_decodable_workspace_size(const int *input_lengths, int alphabet_size,
                                             int minibatch, size_t *size_bytes);
ctcStatus_t b200ctc_decodable(const float *nnet_output, int input_is_logits,
                              const int *input_lengths, int alphabet_size, int minibatch,
                              const float *priors, float prob_scale, float blank_threshold,
                              float floor_value, float *log_probs, int *kept_dev,
                              int *kept_host, void *workspace, size_t workspace_bytes,
                              CUstream stream);

/*
 * Training input path: kaldi::ctc::FormatNnetInput
 * (src/ctc/ctc-nnet-update.cc:351-424) with the CompressedMatrix decompression
 * (src/matrix/compressed-matrix.cc:493-529, bit-exact) moved onto the GPU, so that the compressed
 * bytes are what crosses PCIe.
 * examples_host[m]: HOST pointer to data[m].input_frames' in-memory image (CompressedMatrix::data_:
 *   GlobalHeader{int32 format; float min_value, range; int32 num_rows, num_cols}, then format 1:
 *   PerColHeader[num_cols] + bytes column-major, format 2: uint16 row-major; compressed-matrix.h:128-143).
 * spk_info_host[m]: HOST data[m].spk_info (spk_dim floats; the array may be NULL when spk_dim == 0).
 * left_context = data[0].left_context; nnet_left/right_context = nnet.LeftContext()/RightContext().
 * input_mat: DEVICE [max_num_frames * num_splice * minibatch, feat_dim + spk_dim], row
 *   (t*minibatch + m)*num_splice + s, padding rows zeroed -- exactly *input_mat of the reference.
 * staging_host (pinned recommended) / staging_dev: staging_bytes each, from b200ctc_format_input_size.
 *   The call packs into staging_host and enqueues ONE host-to-device copy + one kernel on `stream`;
 *   staging_host must stay untouched until the stream has passed the copy (alternate two buffers).
 */
ctcStatus_t b200ctc_format_input_size(const void *const *examples_host, int minibatch, int spk_dim,
                                      int left_context, int nnet_left_context, int nnet_right_context,
                                      int *max_num_frames, int *feat_dim, size_t *staging_bytes);
ctcStatus_t b200ctc_format_input(const void *const *examples_host, const float *const *spk_info_host,
                                 int spk_dim, int minibatch, int left_context, int nnet_left_context,
                                 int nnet_right_context, float *input_mat, size_t input_mat_floats,
                                 void *staging_host, void *staging_dev, size_t staging_bytes,
                                 CUstream stream);

#ifdef __cplusplus
}
#endif
#endif /* B200CTC_H_ */
