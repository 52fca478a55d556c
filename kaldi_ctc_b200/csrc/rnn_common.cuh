// kaldi_ctc_b200/csrc/rnn_common.cuh -- shared declarations of libb200rnn.so.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/b200rnn.h"

namespace b200 {

inline __host__ __device__ int gates_of(int mode) { return mode == 2 ? 4 : (mode == 3 ? 3 : 1); }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// One pseudo-layer (layer, direction) of the packed blob; float offsets.
struct PseudoLayer {
  size_t w_in;   // [G*H x Din]  the G input matrices are contiguous
  size_t w_rec;  // [G*H x H]
  size_t b_in;   // [G*H]
  size_t b_rec;  // [G*H]
  int din;
};

// ------------------------------------------------------------------------
// fp32 GEMM (rnn_gemm.cu):  C[MxN] = alpha * A(MxK) * B(KxN) + beta * C + bias
// element (m,k) of A at A[m*sam + k*sak]; (k,n) of B at B[k*sbk + n*sbn].
// splits > 1: K is cut in `splits` ranges, partial products go through
// `partial` ([splits][M][N] floats) and are summed in a fixed order.
// bias (nullable) is bias_a[n] + (n < nb ? bias_b[n] : 0).
// ------------------------------------------------------------------------
struct GemmArgs {
  int M, N, K;
  float alpha, beta;
  const float *A;
  long long sam, sak;
  const float *B;
  long long sbk, sbn;
  float *C;
  int ldc;
  const float *bias_a, *bias_b;
  int nb;
  int splits;
  float *partial;
  // optional second operand pair of the same shape and strides: C = alpha * (A.B + A2.B2) + beta * C + bias
  // (the input gradient of a bidirectional layer: one product per direction, one pass over C)
  const float *A2, *B2;
};
cudaError_t gemm_fp32(const GemmArgs &g, cudaStream_t stream, int *launches);
// C = beta*C + bias + sum_z partial[z]  (fixed summation order)
cudaError_t splitk_reduce(const GemmArgs &g, cudaStream_t stream);
// tcgen05 kind::tf32 GEMM fed by TMA (rnn_gemm_tc.cu); cudaErrorNotSupported when the
// operands break TMA's alignment rules (16-byte base, row pitch % 4 floats)
cudaError_t gemm_tc(const GemmArgs &g, cudaStream_t stream, int *launches);
// math: 0 = fp32 FMA, 1 = tensor cores with fp32 fallback for unaligned operands
extern int g_last_gemm_tc;  // 1 if the last gemm_any ran on tcgen05
extern int g_gemm_pair;     // CTA-pair (cta_group::2) GEMM kernel: -1 = when the shape qualifies, 0 = never, 1 = whenever possible
extern int g_gemm_tma_store; // CTA-pair kernel: epilogue through TMA tile stores when beta = 0 (1) or always through registers (0)
extern int g_last_gemm_pair;  // 1 if the last gemm_tc launch was the CTA-pair kernel
cudaError_t gemm_any(int math, const GemmArgs &g, cudaStream_t stream, int *launches);
cudaError_t column_sums(const float *a, int rows, int cols, int lda, float alpha, float *out, int accumulate,
                        float *partial, size_t partial_floats, cudaStream_t stream, int *launches);
size_t column_sums_partial_floats(int rows, int cols);

// ------------------------------------------------------------------------
// ClipGradientComponent::Backprop with counters and self-repair (rnn_clip.cu)
// ------------------------------------------------------------------------
size_t clip_gradient_workspace_floats(int rows);
cudaError_t clip_gradient_backprop(float *deriv, const float *in_value, int rows, int cols, float thr,
                                   float prop_threshold, float target, float repair_scale, int attempt_repair,
                                   int *counters, const int *decide, float *ws, cudaStream_t stream);

// ------------------------------------------------------------------------
// recurrent kernels, fp32 math (rnn_rec_fp32.cu)
// ------------------------------------------------------------------------
struct RecArgs {
  int mode, T, B, H, dirs;
  int NC;        // CTAs per cluster (one cluster per direction and batch chunk)
  int U;         // hidden units per CTA = H / NC
  int BC;        // batch chunk (<= 16)
  const float *w_rec[2];  // [G*H x H]
  const float *b_rec[2];  // [G*H]   (GRU: the n-gate part is applied inside)
  float *gates[2];        // [T*B x G*H] fwd: pre-activations in, activations out
                          //             bwd: activations in, input-side gate gradients out
  float *cell[2];         // [T*B x H]   LSTM c_t / GRU q_t (bwd: GRU dq_t out)
  float *y;               // [T*B x H*dirs]
  const float *dy;        // bwd only
  int save;               // fwd: 1 = keep activations/cell for backward
  long long *dbg;         // per-phase cycle counters of cluster 0 / CTA 0; only read by kernels built with
                          // -DB200RNN_PHASE_COUNTERS (tuning builds), ignored by release kernels
  float *bias_partial;    // bwd (tensor kernels): [chunk][dir][side 0 = input, 1 = recurrent][G*H] sums of the
                          // gate gradients over time and the chunk's utterances (the bias gradients)
};
// smem bytes for a given geometry (0 = does not fit the fp32 persistent kernels)
size_t rec_fp32_smem_bytes(int mode, int H, int NC, bool backward);
cudaError_t rec_fp32_forward(const RecArgs &a, cudaStream_t stream);
cudaError_t rec_fp32_backward(const RecArgs &a, cudaStream_t stream);
// largest usable cluster size for this geometry (0 if none), probing the device
int rec_fp32_pick_cluster(int mode, int H);

// ------------------------------------------------------------------------
// recurrent kernels on tcgen05 (rnn_rec_tc.cu): BF16 operands, fp32 accumulate/state
// ------------------------------------------------------------------------
bool rec_tc_supported(int mode, int H);
int rec_tc_pick_chunk(int H, int B, int dirs);
cudaError_t rec_tc_forward(const RecArgs &a, cudaStream_t stream);  // a.NC = H/32, a.BC in {4,8,16}
cudaError_t rec_tc_backward(const RecArgs &a, cudaStream_t stream);
// dw[b_in[d] + n] += sum_chunk partial[chunk][d][0][n];  dw[b_rec[d] + n] += sum_chunk partial[chunk][d][1][n]
cudaError_t rec_tc_bias_finalize(const float *partial, int nchunks, int dirs, int GH, float *db_in0, float *db_rec0,
                                 float *db_in1, float *db_rec1, cudaStream_t stream);

// ------------------------------------------------------------------------
// general streaming path (rnn_rec_stream.cu): any shape, one launch per time step, fp32
// ------------------------------------------------------------------------
size_t rec_stream_scratch_floats(int dirs, int B, int H);
cudaError_t rec_stream_forward(const RecArgs &a, cudaStream_t stream);   // needs a.cell even when !a.save (LSTM)
cudaError_t rec_stream_backward(const RecArgs &a, float *scratch, cudaStream_t stream);

}  // namespace b200
