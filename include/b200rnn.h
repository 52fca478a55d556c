/*
 * include/b200rnn.h -- C ABI of libb200rnn.so: the recurrent forward/backward
 * that kaldi-ctc's nnet2 CuDNNRecurrentComponent obtains from cuDNN 5 through
 * the kaldi::cudnn::Recurrent* wrappers (src/cudamatrix/cudnn-recurrent.h).
 * Each entry point names the reference interface it replaces.
 *
 * Conventions (all from the reference's call sites):
 *   - tensors are fp32, row-major, stride == cols (asserted at
 *     src/nnet2/nnet-cudnn-component.cc:512-513,570-572); row index t*B + b
 *   - x [T*B x D], y / dy [T*B x H*dirs] with the bidirectional output
 *     [forward h_t | backward h_t]; every one of the B sequences is run for all
 *     T steps (zero-padded frames are real steps), hx = cx = 0 (:494-506)
 *   - w / dw: one flat fp32 blob in cuDNN-v5 packed order: for pseudo-layer
 *     p = layer*dirs + dir, nlin/2 input matrices [H x in] then nlin/2
 *     recurrent matrices [H x H] (row-major, y = W.x); after ALL matrices, per
 *     pseudo-layer nlin bias vectors of H.  nlin = 2 (RELU/TANH), 8 (LSTM:
 *     i,f,g,o), 6 (GRU: r,z,n).  This blob is also the <FilterParams> field of
 *     the model file (:673-721).
 *   - caller owns every buffer; calls are stream-ordered and never allocate;
 *     errors are status codes (the host wrapper maps them to KALDI_ERR)
 */
#ifndef B200RNN_H_
#define B200RNN_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct CUstream_st *b200rnnStream_t; /* == cudaStream_t */
typedef struct b200rnnPlan_st *b200rnnPlan_t;

typedef enum {
  B200RNN_STATUS_SUCCESS = 0,
  B200RNN_STATUS_INVALID_VALUE = 1,
  B200RNN_STATUS_ALLOC_FAILED = 2,
  B200RNN_STATUS_EXECUTION_FAILED = 3,
  B200RNN_STATUS_NOT_SUPPORTED = 4
} b200rnnStatus_t;

/* rnn-mode of the component's config line (nnet-cudnn-component.cc:252-258) */
typedef enum { B200RNN_RELU = 0, B200RNN_TANH = 1, B200RNN_LSTM = 2, B200RNN_GRU = 3 } b200rnnMode_t;

typedef enum {
  B200RNN_MATH_FP32 = 0,  /* CUDA-core FMA everywhere: matches the oracle to 1e-5 */
  B200RNN_MATH_TENSOR = 1 /* tcgen05: TF32 projections/weight-gradients, BF16 recurrent
                             operands, fp32 accumulation and fp32 cell state */
} b200rnnMath_t;

const char *b200rnnGetStatusString(b200rnnStatus_t status);

/*
 * Replaces the descriptor set-up of CuDNNRecurrentComponent::Init
 * (nnet-cudnn-component.cc:100-315: tensor/filter/dropout/RNN descriptors for
 * minibatch B and at most Tmax steps).  Host only; cheap enough to redo when B
 * changes (InitMiniBatch, :100-102).
 */
b200rnnStatus_t b200rnnCreatePlan(b200rnnPlan_t *plan, b200rnnMode_t mode, int bidirectional,
                                  int num_layers, int input_dim, int hidden_dim, int minibatch,
                                  int max_seq_length, b200rnnMath_t math);
b200rnnStatus_t b200rnnDestroyPlan(b200rnnPlan_t plan);

/* cudnnGetRNNParamsSize (cudnn-recurrent.h GetRecurrentParamsSize), in floats */
b200rnnStatus_t b200rnnGetParamCount(b200rnnPlan_t plan, size_t *count);

/*
 * cudnnGetRNNLinLayerMatrixParams / ...BiasParams (GetRecurrentLinLayer*Params,
 * used at nnet-cudnn-component.cc:336-408,417-483): float offset into the blob
 * and shape of linear layer `lin_id` of pseudo-layer `pseudo_layer`.
 */
b200rnnStatus_t b200rnnLocateParam(b200rnnPlan_t plan, int pseudo_layer, int lin_id, int is_bias,
                                   size_t *offset, int *rows, int *cols);

/* cudnnGetRNNWorkspaceSize / cudnnGetRNNTrainingReserveSize, in bytes */
b200rnnStatus_t b200rnnGetWorkspaceSize(b200rnnPlan_t plan, size_t *bytes);
b200rnnStatus_t b200rnnGetReserveSize(b200rnnPlan_t plan, size_t *bytes);

/*
 * cudnnRNNForwardTraining (reserve != NULL; RecurrentForwardTraining,
 * nnet-cudnn-component.cc:545-554) and cudnnRNNForwardInference (reserve == NULL;
 * :534-543).  seq_length <= max_seq_length.  The reserve needs no zeroing.
 */
b200rnnStatus_t b200rnnForward(b200rnnPlan_t plan, int seq_length, const float *x, const float *w,
                               float *y, void *workspace, void *reserve, b200rnnStream_t stream);

/*
 * cudnnRNNBackwardData (RecurrentBackwardData, :577-587) with dhy = dcy = 0:
 * dx [T*B x D] from dy, using the reserve of the matching forward call.
 * Leaves the gate gradients in the reserve for b200rnnBackwardWeights.
 */
b200rnnStatus_t b200rnnBackwardData(b200rnnPlan_t plan, int seq_length, const float *y,
                                    const float *dy, const float *w, float *dx, void *workspace,
                                    void *reserve, b200rnnStream_t stream);

/*
 * cudnnRNNBackwardWeights (RecurrentBackwardWeights, :595-599): ACCUMULATES the
 * weight gradient into dw (blob-shaped; the reference zeroes it first, :594).
 * Must follow b200rnnBackwardData on the same reserve.
 */
b200rnnStatus_t b200rnnBackwardWeights(b200rnnPlan_t plan, int seq_length, const float *x,
                                       const float *y, float *dw, void *workspace, void *reserve,
                                       b200rnnStream_t stream);

/*
 * The tail of CuDNNRecurrentComponent::Backprop fused into one pass (:602-603,
 * 612-614): w += learning_rate * clamp(dw, -clip, +clip).  clip <= 0 disables
 * the clamp.  n = number of floats.
 */
b200rnnStatus_t b200rnnClipAndUpdate(float *w, const float *dw, size_t n, float learning_rate,
                                     float clip, b200rnnStream_t stream);

/*
 * The same update as TrainNnetSimple applies it (src/ctc/ctc-nnet-train.cc:194-202, 220-245), in one pass:
 *   delta == NULL (momentum 0: components update the model directly):  w += learning_rate * clamp(dw)
 *   delta != NULL (the reference's gradient Nnet `delta_nnet`):        delta += learning_rate * clamp(dw);
 *                                                                      w += delta;  delta *= momentum
 * (`nnet->AddNnet(1.0, *delta_nnet); delta_nnet->Scale(momentum)`, :243-244).  0 <= momentum < 1 (:191).
 * skip_flag_dev: optional DEVICE int; when it is non-zero at execution time the call changes nothing --
 * the asynchronous form of the reference's "deriv sum is inf/nan" abort (ctc-nnet-update.cc:232-234);
 * pass b200ctcOptions.nonfinite_dev of the minibatch's CTC call.
 */
b200rnnStatus_t b200rnnUpdate(float *w, float *delta, const float *dw, size_t n, float learning_rate,
                              float clip, float momentum, const int *skip_flag_dev, b200rnnStream_t stream);

/* ClipGradientComponent::Backprop, norm-based (:936-957): each row of d
 * [rows x cols] is scaled to L2-norm <= threshold, in place. */
b200rnnStatus_t b200rnnClipRowNorm(float *d, int rows, int cols, float threshold,
                                   b200rnnStream_t stream);

/*
 * The whole of ClipGradientComponent::Backprop (nnet-cudnn-component.cc:912-1055), stream-ordered and without a
 * host round trip: norm-based row clipping of `deriv` [rows x cols] in place (:936-957), the component's counters,
 * and the stochastic self-repair term RepairGradients (:970-1055).
 *   in_value             the component's forward input [rows x cols] (what the self-repair pushes towards
 *                        self_repair_target); NULL disables self-repair
 *   attempt_repair       the caller's coin: RandUniform() <= repair_probability (0.5, :979-984), drawn on the host
 *   counters_dev         DEVICE int[4] = {num_clipped_, count_, num_self_repaired_, num_backpropped_} of the
 *                        component being updated (to_update), incremented as the reference does; may be NULL
 *   decide_counters_dev  DEVICE int[4] the repair decision reads (`this` in the reference: count_ > 0 and
 *                        num_clipped_/count_ > threshold, :983-995); pass counters_dev when the net updates itself
 *                        (with momentum the reference updates a copy, so `this` stays at its start-up values)
 *   workspace            DEVICE, b200rnnClipGradientWorkspaceSize(rows) bytes
 */
b200rnnStatus_t b200rnnClipGradientWorkspaceSize(int rows, size_t *bytes);
b200rnnStatus_t b200rnnClipGradientBackprop(float *deriv, const float *in_value, int rows, int cols,
                                            float clipping_threshold, float self_repair_clipped_proportion_threshold,
                                            float self_repair_target, float self_repair_scale, int attempt_repair,
                                            int *counters_dev, const int *decide_counters_dev, void *workspace,
                                            size_t workspace_bytes, b200rnnStream_t stream);

/* Plain row-major fp32 GEMM used for the adjacent AffineComponent
 * (src/nnet2/nnet-component.cc:1184-1226): C[M x N] = alpha*op(A)*op(B) + beta*C
 * (+ bias[n] broadcast over rows when bias != NULL).  trans: 0 = as stored
 * [rows x cols], 1 = transposed.  math selects FP32 or TF32-tcgen05.  An optional
 * DEVICE workspace enables a deterministic split-K for short-and-wide products
 * (weight gradients); NULL / 0 is always valid. */
b200rnnStatus_t b200rnnGemm(int transA, int transB, int M, int N, int K, float alpha,
                            const float *A, int lda, const float *B, int ldb, float beta, float *C,
                            int ldc, const float *bias, b200rnnMath_t math, void *workspace,
                            size_t workspace_bytes, b200rnnStream_t stream);

/* 1 if the most recent GEMM issued by this library ran on tcgen05 (MATH_TENSOR with
 * TMA-compatible operands), 0 if it took the fp32 FMA path. */
int b200rnnLastGemmUsedTensorCores(void);
/* 1 if that GEMM was the CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 tiles on the two SMs of a TPC). */
int b200rnnLastGemmUsedCtaPair(void);
/* Test / tuning hook, not needed in production: "GEMM_PAIR" = -1 (default: CTA pairs for M >= 2048, N > 128,
 * no split-K), 0 (never), 1 (whenever the tile shape allows); "GEMM_TMA_STORE" = 1 (default: the CTA-pair kernel
 * writes its output with TMA tile stores when beta = 0) or 0 (register stores).  Every setting computes the same function.
 * Returns 0, or -1 for an unknown key. */
int b200rnnSetTuning(const char *key, int value);

/* Column sums: out[c] = (accumulate ? out[c] : 0) + alpha * sum_r a[r, c]
 * (AffineComponent bias update: bias += lr * colsum(deriv)).  Deterministic. */
b200rnnStatus_t b200rnnColumnSums(const float *a, int rows, int cols, int lda, float alpha,
                                  float *out, int accumulate, void *workspace,
                                  size_t workspace_bytes, b200rnnStream_t stream);

/* FLOPs of one forward call at seq_length T (counting padded frames, as the
 * reference computes them): sum_layers dirs*2*ng*H*(D_l+H) per (frame, utt). */
double b200rnnForwardFlops(b200rnnPlan_t plan, int seq_length);

/* Optional CUDA-event timing of the kernels a plan launches (the reference times
 * every cuDNN call into CuDevice::AccuProfile, src/cudamatrix/cudnn-recurrent.cc:25-30).
 * Categories: 0 = recurrent forward kernel, 1 = recurrent backward kernel, 2 = projection / dx GEMMs (Forward,
 * BackwardData), 3 = weight-gradient GEMMs (BackwardWeights; callers run them on a side stream, where their
 * elapsed times overlap other kernels).
 * GetProfile synchronises the recorded events, returns the sums since the previous
 * GetProfile of that category and resets them. */
b200rnnStatus_t b200rnnSetProfiling(b200rnnPlan_t plan, int enable);
b200rnnStatus_t b200rnnGetProfile(b200rnnPlan_t plan, int category, float *total_ms, int *launches);

/* Kernels launched by the last Forward / BackwardData / BackwardWeights call on
 * this plan (bench.py's gpu_launches bookkeeping). */
int b200rnnLastLaunchCount(b200rnnPlan_t plan);

#ifdef __cplusplus
}
#endif
#endif /* B200RNN_H_ */
