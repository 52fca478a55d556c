// kaldi_ctc_b200/csrc/rnn.cu -- C ABI of libb200rnn.so (include/b200rnn.h):
// plan, blob layout, workspace/reserve carving and the per-layer orchestration
// of forward / backward-data / backward-weights.  Replaces what
// CuDNNRecurrentComponent gets from cuDNN 5 (src/nnet2/nnet-cudnn-component.cc
// :100-315 descriptors, :534-555 forward, :576-599 backward).
#include <string.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <new>
#include <utility>
#include <vector>

#include "rnn_common.cuh"

using namespace b200;

struct b200rnnPlan_st {
  int mode, dirs, layers, D, H, B, Tmax, math;
  int G, GH, HO;
  std::vector<PseudoLayer> pl;
  size_t param_count;
  // recurrent-kernel geometry, probed on first use (needs the device)
  int NC, BC;        // fp32 kernels
  int tcNC, tcBC;    // tcgen05 kernels (0 = not usable for this shape)
  bool geometry_ready;
  int launches;
  // reserve layout (floats), per layer
  std::vector<size_t> r_gates[2], r_cell[2], r_y, r_bias;
  bool tc_bwd_used;  // the last BackwardData left fused bias gradients in the reserve
  // tuning switches, read from the environment ONCE when the plan is created
  bool force_stream, tc_no_bwd, phase_counters;
  int force_bc;
  size_t reserve_floats;
  // workspace layout (floats)
  size_t w_colsum, w_splitk, w_pp[2], w_gates[2], w_stream, w_cell[2], workspace_floats;
  size_t splitk_floats, colsum_floats;
  // optional event timing: [category] -> recorded (start, stop) pairs
  bool profiling;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[4];
  size_t ev_used[4];
};

namespace {

constexpr int kSplitK = 16;

int din_of(const b200rnnPlan_st *p, int layer) { return layer == 0 ? p->D : p->HO; }

b200rnnStatus_t ensure_geometry(b200rnnPlan_st *p) {
  if (p->geometry_ready) return B200RNN_STATUS_SUCCESS;
  p->NC = rec_fp32_pick_cluster(p->mode, p->H);  // 0: no on-chip configuration -> streaming kernels
  if (p->force_stream) p->NC = 0;
  p->BC = 16;
  p->tcNC = p->tcBC = 0;
  if (p->math == 1 && rec_tc_supported(p->mode, p->H)) {
    p->tcNC = p->H / 32;
    p->tcBC = rec_tc_pick_chunk(p->H, p->B, p->dirs);
    if (p->force_bc) p->tcBC = p->force_bc;  // tuning override: 4, 8 or 16
  }
  p->geometry_ready = true;
  return B200RNN_STATUS_SUCCESS;
}

__global__ void clip_update_kernel(float *w, const float *dw, size_t n, float lr, float clip) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float g = dw[i];
    if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);
    w[i] = fmaf(lr, g, w[i]);
  }
}

// TrainNnetSimple's update through a gradient Nnet (ctc-nnet-train.cc:194-202, 243-244):
//   delta += lr * clamp(dw);  w += delta;  delta *= momentum          (delta == NULL: w += lr * clamp(dw))
// skipped altogether when *skip_flag != 0 (the CTC call's non-finite flag, include/b200ctc.h)
__global__ void update_kernel(float *w, float *delta, const float *dw, size_t n, float lr, float clip, float momentum,
                              const int *skip_flag) {
  if (skip_flag && *skip_flag != 0) return;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    float g = dw[i];
    if (clip > 0.f) g = fminf(fmaxf(g, -clip), clip);
    if (delta) {
      const float d = fmaf(lr, g, delta[i]);
      w[i] += d;
      delta[i] = momentum * d;
    } else {
      w[i] = fmaf(lr, g, w[i]);
    }
  }
}

// one warp per row: scale rows whose L2 norm exceeds thr down to thr
__global__ void clip_row_norm_kernel(float *d, int rows, int cols, float thr) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float *p = d + (size_t)row * cols;
  float ss = 0.f;
  for (int c = lane; c < cols; c += 32) ss = fmaf(p[c], p[c], ss);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  const float r = ss / (thr * thr);
  if (r > 1.0f) {
    const float sc = rsqrtf(r);
    for (int c = lane; c < cols; c += 32) p[c] *= sc;
  }
}

b200rnnStatus_t to_status(cudaError_t e) {
  if (e == cudaSuccess) return B200RNN_STATUS_SUCCESS;
  fprintf(stderr, "b200rnn: CUDA error: %s\n", cudaGetErrorString(e));
  return B200RNN_STATUS_EXECUTION_FAILED;
}

struct Timed {  // records an event pair around a launch when profiling is on
  b200rnnPlan_st *p;
  int cat;
  cudaStream_t s;
  Timed(b200rnnPlan_st *p_, int cat_, cudaStream_t s_) : p(p_), cat(cat_), s(s_) {
    if (!p->profiling) return;
    if (p->ev_used[cat] == p->ev[cat].size()) {
      cudaEvent_t a, b;
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      p->ev[cat].push_back(std::make_pair(a, b));
    }
    cudaEventRecord(p->ev[cat][p->ev_used[cat]].first, s);
  }
  ~Timed() {
    if (!p->profiling) return;
    cudaEventRecord(p->ev[cat][p->ev_used[cat]].second, s);
    p->ev_used[cat]++;
  }
};

#define CK(expr)                                   \
  do {                                             \
    cudaError_t e__ = (expr);                      \
    if (e__ != cudaSuccess) return to_status(e__); \
  } while (0)

}  // namespace

extern "C" {

const char *b200rnnGetStatusString(b200rnnStatus_t s) {
  switch (s) {
    case B200RNN_STATUS_SUCCESS: return "success";
    case B200RNN_STATUS_INVALID_VALUE: return "invalid value";
    case B200RNN_STATUS_ALLOC_FAILED: return "allocation failed";
    case B200RNN_STATUS_EXECUTION_FAILED: return "execution failed";
    case B200RNN_STATUS_NOT_SUPPORTED: return "configuration not supported";
    default: return "unknown status";
  }
}

b200rnnStatus_t b200rnnCreatePlan(b200rnnPlan_t *plan, b200rnnMode_t mode, int bidirectional,
                                  int num_layers, int input_dim, int hidden_dim, int minibatch,
                                  int max_seq_length, b200rnnMath_t math) {
  if (!plan || (int)mode < 0 || (int)mode > 3 || num_layers < 1 || input_dim < 1 ||
      hidden_dim < 1 || minibatch < 1 || max_seq_length < 1 || (int)math < 0 || (int)math > 1)
    return B200RNN_STATUS_INVALID_VALUE;
  b200rnnPlan_st *p = new (std::nothrow) b200rnnPlan_st();
  if (!p) return B200RNN_STATUS_ALLOC_FAILED;
  p->mode = mode;
  p->dirs = bidirectional ? 2 : 1;
  p->layers = num_layers;
  p->D = input_dim;
  p->H = hidden_dim;
  p->B = minibatch;
  p->Tmax = max_seq_length;
  p->math = math;
  p->G = gates_of(mode);
  p->GH = p->G * p->H;
  p->HO = p->H * p->dirs;
  p->geometry_ready = false;
  p->NC = p->BC = 0;
  p->launches = 0;
  p->profiling = false;
  p->force_stream = getenv("B200RNN_FORCE_STREAM") != nullptr;
  p->tc_no_bwd = getenv("B200RNN_TC_NO_BWD") != nullptr;
#ifdef B200RNN_PHASE_COUNTERS
  p->phase_counters = getenv("B200RNN_TC_PROFILE") != nullptr;
#else
  p->phase_counters = false;
#endif
  p->force_bc = 0;
  if (const char *e = getenv("B200RNN_TC_BC")) {
    const int v = atoi(e);
    if (v == 4 || v == 8 || v == 16) p->force_bc = v;
  }
  p->ev_used[0] = p->ev_used[1] = p->ev_used[2] = p->ev_used[3] = 0;
  // blob: all matrices of all pseudo-layers, then all biases
  const int npl = p->layers * p->dirs;
  p->pl.resize(npl);
  size_t off = 0;
  for (int q = 0; q < npl; q++) {
    const int din = din_of(p, q / p->dirs);
    p->pl[q].din = din;
    p->pl[q].w_in = off;
    off += (size_t)p->GH * din;
    p->pl[q].w_rec = off;
    off += (size_t)p->GH * p->H;
  }
  for (int q = 0; q < npl; q++) {
    p->pl[q].b_in = off;
    off += p->GH;
    p->pl[q].b_rec = off;
    off += p->GH;
  }
  p->param_count = off;
  // reserve
  const size_t TB = (size_t)p->Tmax * p->B;
  size_t r = 0;
  for (int d = 0; d < 2; d++) {
    p->r_gates[d].assign(p->layers, 0);
    p->r_cell[d].assign(p->layers, 0);
  }
  p->r_y.assign(p->layers, 0);
  p->r_bias.assign(p->layers, 0);
  p->tc_bwd_used = false;
  for (int l = 0; l < p->layers; l++) {
    for (int d = 0; d < p->dirs; d++) {
      p->r_gates[d][l] = r;
      r = align_up(r + TB * p->GH, 64);
      p->r_cell[d][l] = r;
      r = align_up(r + TB * p->H, 64);
    }
    if (l + 1 < p->layers) {
      p->r_y[l] = r;
      r = align_up(r + TB * p->HO, 64);
    }
    p->r_bias[l] = r;  // [chunks of 4 utterances][dirs][2][GH]
    r = align_up(r + (size_t)((p->B + 3) / 4) * p->dirs * 2 * p->GH, 64);
  }
  p->reserve_floats = r;
  // workspace
  size_t w = 0;
  p->colsum_floats = column_sums_partial_floats((int)TB, p->GH);
  p->w_colsum = w;
  w = align_up(w + p->colsum_floats, 64);
  const int maxin = std::max(std::max(p->D, p->HO), p->H);
  p->splitk_floats = (size_t)kSplitK * p->GH * maxin;
  p->w_splitk = w;
  w = align_up(w + p->splitk_floats, 64);
  for (int i = 0; i < 2; i++) {
    p->w_pp[i] = w;
    if (p->layers > 1) w = align_up(w + TB * p->HO, 64);
  }
  for (int d = 0; d < p->dirs; d++) {
    p->w_gates[d] = w;
    w = align_up(w + TB * p->GH, 64);
    p->w_cell[d] = w;   // inference scratch of the streaming path (cell state through HBM)
    w = align_up(w + TB * p->H, 64);
  }
  p->w_stream = w;
  w = align_up(w + rec_stream_scratch_floats(p->dirs, p->B, p->H), 64);
  p->workspace_floats = w;
  *plan = p;
  return B200RNN_STATUS_SUCCESS;
}

b200rnnStatus_t b200rnnDestroyPlan(b200rnnPlan_t plan) {
  if (plan)
    for (int c = 0; c < 4; c++)
      for (auto &e : plan->ev[c]) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
      }
  delete plan;
  return B200RNN_STATUS_SUCCESS;
}

b200rnnStatus_t b200rnnGetParamCount(b200rnnPlan_t p, size_t *count) {
  if (!p || !count) return B200RNN_STATUS_INVALID_VALUE;
  *count = p->param_count;
  return B200RNN_STATUS_SUCCESS;
}

b200rnnStatus_t b200rnnLocateParam(b200rnnPlan_t p, int pseudo_layer, int lin_id, int is_bias,
                                   size_t *offset, int *rows, int *cols) {
  if (!p || !offset || !rows || !cols) return B200RNN_STATUS_INVALID_VALUE;
  if (pseudo_layer < 0 || pseudo_layer >= p->layers * p->dirs || lin_id < 0 || lin_id >= 2 * p->G)
    return B200RNN_STATUS_INVALID_VALUE;
  const PseudoLayer &q = p->pl[pseudo_layer];
  const bool rec = lin_id >= p->G;
  const int g = rec ? lin_id - p->G : lin_id;
  *rows = p->H;
  if (is_bias) {
    *offset = (rec ? q.b_rec : q.b_in) + (size_t)g * p->H;
    *cols = 1;
  } else {
    const int in = rec ? p->H : q.din;
    *offset = (rec ? q.w_rec : q.w_in) + (size_t)g * p->H * in;
    *cols = in;
  }
  return B200RNN_STATUS_SUCCESS;
}

b200rnnStatus_t b200rnnGetWorkspaceSize(b200rnnPlan_t p, size_t *bytes) {
  if (!p || !bytes) return B200RNN_STATUS_INVALID_VALUE;
  *bytes = p->workspace_floats * sizeof(float);
  return B200RNN_STATUS_SUCCESS;
}

b200rnnStatus_t b200rnnGetReserveSize(b200rnnPlan_t p, size_t *bytes) {
  if (!p || !bytes) return B200RNN_STATUS_INVALID_VALUE;
  *bytes = p->reserve_floats * sizeof(float);
  return B200RNN_STATUS_SUCCESS;
}

b200rnnStatus_t b200rnnForward(b200rnnPlan_t p, int T, const float *x, const float *w, float *y,
                               void *workspace, void *reserve, b200rnnStream_t stream_) {
  if (!p || !x || !w || !y || !workspace || T < 1 || T > p->Tmax) return B200RNN_STATUS_INVALID_VALUE;
  b200rnnStatus_t gs = ensure_geometry(p);
  if (gs != B200RNN_STATUS_SUCCESS) return gs;
  cudaStream_t stream = (cudaStream_t)stream_;
  float *ws = static_cast<float *>(workspace), *rs = static_cast<float *>(reserve);
  const int TB = T * p->B;
  p->launches = 0;
  for (int l = 0; l < p->layers; l++) {
    const int din = din_of(p, l);
    const float *in = l == 0 ? x : (rs ? rs + p->r_y[l - 1] : ws + p->w_pp[(l - 1) & 1]);
    float *out = l == p->layers - 1 ? y : (rs ? rs + p->r_y[l] : ws + p->w_pp[l & 1]);
    RecArgs a = {};
    a.mode = p->mode; a.T = T; a.B = p->B; a.H = p->H; a.dirs = p->dirs;
    a.NC = p->NC; a.U = p->NC ? p->H / p->NC : 0; a.BC = p->BC;
    a.y = out; a.dy = nullptr; a.save = rs ? 1 : 0;
    for (int d = 0; d < p->dirs; d++) {
      const PseudoLayer &q = p->pl[l * p->dirs + d];
      float *gates = rs ? rs + p->r_gates[d][l] : ws + p->w_gates[d];
      // hoisted input projection with both biases folded in (GRU: recurrent n-bias stays inside)
      GemmArgs g = {};
      g.M = TB; g.N = p->GH; g.K = din; g.alpha = 1.f; g.beta = 0.f;
      g.A = in; g.sam = din; g.sak = 1;
      g.B = w + q.w_in; g.sbk = 1; g.sbn = din;
      g.C = gates; g.ldc = p->GH;
      g.bias_a = w + q.b_in; g.bias_b = w + q.b_rec;
      g.nb = p->mode == 3 ? 2 * p->H : p->GH;
      g.splits = 1; g.partial = nullptr;
      {
        Timed tm(p, 2, stream);
        CK(gemm_any(p->math, g, stream, &p->launches));
      }
      a.w_rec[d] = w + q.w_rec;
      a.b_rec[d] = w + q.b_rec;
      a.gates[d] = gates;
      a.cell[d] = rs ? rs + p->r_cell[d][l] : ws + p->w_cell[d];
    }
    {
      Timed tm(p, 0, stream);
      if (p->tcNC) {
        a.NC = p->tcNC; a.U = 32; a.BC = p->tcBC;
        static long long *dbg = nullptr;
        const bool prof = p->phase_counters;
        if (prof && !dbg) cudaMalloc(&dbg, 64 * sizeof(long long));
        if (prof) cudaMemsetAsync(dbg, 0, 64 * sizeof(long long), stream);
        a.dbg = prof ? dbg : nullptr;
        CK(rec_tc_forward(a, stream));
        if (prof) {  // tuning aid: cycles per step of each phase (cluster 0, CTA 0)
          long long h[24];
          cudaStreamSynchronize(stream);
          cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
          const double n = T;
          fprintf(stderr, "[b200rnn fwd tc] cyc/step epilogue: wait_acc %.0f tmem_ld %.0f math %.0f pack+dsmem %.0f "
                  "bar %.0f arrive %.0f gstore+prefetch %.0f | mma warp: wait_h %.0f fence %.0f issue %.0f complete %.0f\n",
                  h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[8] / n, h[9] / n, h[10] / n,
                  h[11] / n);
          fprintf(stderr, "[b200rnn fwd tc] issuer-complete -> epilogue-sees-acc %.0f cyc; h sent -> next h seen by issuer %.0f cyc\n",
                  (h[13] - h[12]) / n, (h[14] - h[15]) / n + (double)(h[8] + h[9] + h[10] + h[11]) / n);
          fprintf(stderr, "[b200rnn fwd tc] tcgen05.fence::after_thread_sync in the epilogue: %.0f cyc\n", h[16] / n);
          long long hc[64];
          cudaMemcpy(hc, dbg, sizeof(hc), cudaMemcpyDeviceToHost);
          fprintf(stderr, "[b200rnn fwd tc] loop cycles/step per cluster (CTA 0):");
          for (int c = 0; c < 8; c++) fprintf(stderr, " %.0f", hc[32 + c] / n);
          fprintf(stderr, "\n");
        }
      } else if (p->NC) {
        CK(rec_fp32_forward(a, stream));
      } else {
        CK(rec_stream_forward(a, stream));
        p->launches += T - 1;
      }
    }
    p->launches++;
  }
  return B200RNN_STATUS_SUCCESS;
}

b200rnnStatus_t b200rnnBackwardData(b200rnnPlan_t p, int T, const float *y, const float *dy,
                                    const float *w, float *dx, void *workspace, void *reserve,
                                    b200rnnStream_t stream_) {
  if (!p || !y || !dy || !w || !workspace || !reserve || T < 1 || T > p->Tmax)
    return B200RNN_STATUS_INVALID_VALUE;
  b200rnnStatus_t gs = ensure_geometry(p);
  if (gs != B200RNN_STATUS_SUCCESS) return gs;
  cudaStream_t stream = (cudaStream_t)stream_;
  float *ws = static_cast<float *>(workspace), *rs = static_cast<float *>(reserve);
  const int TB = T * p->B;
  p->launches = 0;
  for (int l = p->layers - 1; l >= 0; l--) {
    const int din = din_of(p, l);
    const float *yl = l == p->layers - 1 ? y : rs + p->r_y[l];
    const float *dyl = l == p->layers - 1 ? dy : ws + p->w_pp[l & 1];
    float *dxl = l == 0 ? dx : ws + p->w_pp[(l - 1) & 1];
    RecArgs a = {};
    a.mode = p->mode; a.T = T; a.B = p->B; a.H = p->H; a.dirs = p->dirs;
    a.NC = p->NC; a.U = p->NC ? p->H / p->NC : 0; a.BC = p->BC;
    a.y = const_cast<float *>(yl); a.dy = dyl; a.save = 1;
    for (int d = 0; d < p->dirs; d++) {
      const PseudoLayer &q = p->pl[l * p->dirs + d];
      a.w_rec[d] = w + q.w_rec;
      a.b_rec[d] = w + q.b_rec;
      a.gates[d] = rs + p->r_gates[d][l];
      a.cell[d] = rs + p->r_cell[d][l];
    }
    {
      Timed tm(p, 1, stream);
      if (p->tcNC && !p->tc_no_bwd) {
        a.NC = p->tcNC; a.U = 32; a.BC = p->tcBC;
        a.bias_partial = rs + p->r_bias[l];
        static long long *dbgb = nullptr;
        const bool profb = p->phase_counters;
        if (profb && !dbgb) cudaMalloc(&dbgb, 64 * sizeof(long long));
        if (profb) cudaMemsetAsync(dbgb, 0, 64 * sizeof(long long), stream);
        a.dbg = profb ? dbgb : nullptr;
        CK(rec_tc_backward(a, stream));
        if (profb) {  // tuning aid: cycles per step of each phase (cluster 0, CTA 0)
          long long h[16];
          cudaStreamSynchronize(stream);
          cudaMemcpy(h, dbgb, sizeof(h), cudaMemcpyDeviceToHost);
          const double n = T;
          fprintf(stderr, "[b200rnn bwd tc] cyc/step epilogue: wait_recv %.0f sum %.0f math %.0f pack+fence+arrive %.0f "
                  "gstore+prefetch %.0f wait_acc %.0f tmem_ld+st.async %.0f | mma warp: wait_dg %.0f fence+issue+commit %.0f\n",
                  h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[8] / n, h[9] / n);
        }
        p->tc_bwd_used = true;
      } else if (p->NC) {
        CK(rec_fp32_backward(a, stream));
        p->tc_bwd_used = false;
      } else {
        CK(rec_stream_backward(a, ws + p->w_stream, stream));
        p->tc_bwd_used = false;
        p->launches += 2 * T - 2;
      }
    }
    p->launches++;
    if (dxl) {
      // dx = sum over directions of dG_d . Wi_d: both products in ONE pass over dx where the GEMM can
      // (gemm_any falls back to two launches, the second accumulating)
      const PseudoLayer &q0 = p->pl[l * p->dirs];
      GemmArgs g = {};
      g.M = TB; g.N = din; g.K = p->GH; g.alpha = 1.f; g.beta = 0.f;
      g.A = rs + p->r_gates[0][l]; g.sam = p->GH; g.sak = 1;
      g.B = w + q0.w_in; g.sbk = din; g.sbn = 1;
      g.C = dxl; g.ldc = din;
      g.splits = 1;
      if (p->dirs == 2) {
        g.A2 = rs + p->r_gates[1][l];
        g.B2 = w + p->pl[l * p->dirs + 1].w_in;
      }
      Timed tm(p, 2, stream);
      CK(gemm_any(p->math, g, stream, &p->launches));
    }
  }
  return B200RNN_STATUS_SUCCESS;
}

b200rnnStatus_t b200rnnBackwardWeights(b200rnnPlan_t p, int T, const float *x, const float *y,
                                       float *dw, void *workspace, void *reserve,
                                       b200rnnStream_t stream_) {
  if (!p || !x || !y || !dw || !workspace || !reserve || T < 1 || T > p->Tmax)
    return B200RNN_STATUS_INVALID_VALUE;
  cudaStream_t stream = (cudaStream_t)stream_;
  float *ws = static_cast<float *>(workspace), *rs = static_cast<float *>(reserve);
  const int TB = T * p->B, B = p->B, H = p->H, GH = p->GH, HO = p->HO;
  p->launches = 0;
  for (int l = 0; l < p->layers; l++) {
    const int din = din_of(p, l);
    const float *in = l == 0 ? x : rs + p->r_y[l - 1];
    const float *yl = l == p->layers - 1 ? y : rs + p->r_y[l];
    for (int d = 0; d < p->dirs; d++) {
      const PseudoLayer &q = p->pl[l * p->dirs + d];
      const float *dg = rs + p->r_gates[d][l];
      const float *dq = rs + p->r_cell[d][l];
      // dWi += dG^T . in          [GH x din], K = T*B
      GemmArgs g = {};
      g.M = GH; g.N = din; g.K = TB; g.alpha = 1.f; g.beta = 1.f;
      g.A = dg; g.sam = 1; g.sak = GH;
      g.B = in; g.sbk = din; g.sbn = 1;
      g.C = dw + q.w_in; g.ldc = din;
      g.splits = kSplitK; g.partial = ws + p->w_splitk;
      {
        Timed tm(p, 3, stream);
        CK(gemm_any(p->math, g, stream, &p->launches));
      }
      // dR += dGrec^T . h_prev    h_prev(t) = y(t -+ 1): a row shift of B
      if (T > 1) {
        const int K = (T - 1) * B;
        const size_t sh_g = d == 0 ? (size_t)B : 0, sh_y = d == 0 ? 0 : (size_t)B;
        GemmArgs r = {};
        r.N = H; r.K = K; r.alpha = 1.f; r.beta = 1.f;
        r.B = yl + sh_y * HO + (size_t)d * H; r.sbk = HO; r.sbn = 1;
        r.ldc = H; r.splits = kSplitK; r.partial = ws + p->w_splitk;
        r.M = p->mode == 3 ? 2 * H : GH;
        r.A = dg + sh_g * GH; r.sam = 1; r.sak = GH;
        r.C = dw + q.w_rec;
        {
          Timed tm(p, 3, stream);
          CK(gemm_any(p->math, r, stream, &p->launches));
        }
        if (p->mode == 3) {  // n-gate: recurrent-side gradient lives in the cell buffer
          r.M = H;
          r.A = dq + sh_g * H; r.sam = 1; r.sak = H;
          r.C = dw + q.w_rec + (size_t)2 * H * H;
          Timed tm(p, 3, stream);
          CK(gemm_any(p->math, r, stream, &p->launches));
        }
      }
      // biases: the tensor backward kernel already summed the gate gradients (per batch chunk)
      if (p->tc_bwd_used) {
        if (d == 0) {
          const PseudoLayer &q1 = p->pl[l * p->dirs + (p->dirs - 1)];
          CK(rec_tc_bias_finalize(rs + p->r_bias[l], (B + p->tcBC - 1) / p->tcBC, p->dirs, GH, dw + q.b_in,
                                  dw + q.b_rec, dw + q1.b_in, dw + q1.b_rec, stream));
          p->launches++;
        }
        continue;
      }
      float *part = ws + p->w_colsum;
      CK(column_sums(dg, TB, GH, GH, 1.f, dw + q.b_in, 1, part, p->colsum_floats, stream, &p->launches));
      if (p->mode == 3) {
        CK(column_sums(dg, TB, 2 * H, GH, 1.f, dw + q.b_rec, 1, part, p->colsum_floats, stream, &p->launches));
        CK(column_sums(dq, TB, H, H, 1.f, dw + q.b_rec + 2 * H, 1, part, p->colsum_floats, stream, &p->launches));
      } else {
        CK(column_sums(dg, TB, GH, GH, 1.f, dw + q.b_rec, 1, part, p->colsum_floats, stream, &p->launches));
      }
    }
  }
  return B200RNN_STATUS_SUCCESS;
}

b200rnnStatus_t b200rnnClipAndUpdate(float *w, const float *dw, size_t n, float lr, float clip,
                                     b200rnnStream_t stream) {
  if (!w || !dw) return B200RNN_STATUS_INVALID_VALUE;
  if (n == 0) return B200RNN_STATUS_SUCCESS;
  const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
  clip_update_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, dw, n, lr, clip);
  return to_status(cudaGetLastError());
}

b200rnnStatus_t b200rnnUpdate(float *w, float *delta, const float *dw, size_t n, float lr, float clip,
                              float momentum, const int *skip_flag_dev, b200rnnStream_t stream) {
  if (!w || !dw || momentum < 0.f || momentum >= 1.f) return B200RNN_STATUS_INVALID_VALUE;
  if (n == 0) return B200RNN_STATUS_SUCCESS;
  const unsigned grid = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
  update_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(w, delta, dw, n, lr, clip, momentum, skip_flag_dev);
  return to_status(cudaGetLastError());
}

b200rnnStatus_t b200rnnClipRowNorm(float *d, int rows, int cols, float threshold,
                                   b200rnnStream_t stream) {
  if (!d || rows < 0 || cols < 1 || threshold <= 0.f) return B200RNN_STATUS_INVALID_VALUE;
  if (rows == 0) return B200RNN_STATUS_SUCCESS;
  clip_row_norm_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>(d, rows, cols, threshold);
  return to_status(cudaGetLastError());
}

b200rnnStatus_t b200rnnClipGradientWorkspaceSize(int rows, size_t *bytes) {
  if (rows < 0 || !bytes) return B200RNN_STATUS_INVALID_VALUE;
  *bytes = clip_gradient_workspace_floats(rows) * sizeof(float);
  return B200RNN_STATUS_SUCCESS;
}

b200rnnStatus_t b200rnnClipGradientBackprop(float *deriv, const float *in_value, int rows, int cols,
                                            float clipping_threshold, float self_repair_clipped_proportion_threshold,
                                            float self_repair_target, float self_repair_scale, int attempt_repair,
                                            int *counters_dev, const int *decide_counters_dev, void *workspace,
                                            size_t workspace_bytes, b200rnnStream_t stream) {
  if (!deriv || rows < 0 || cols < 1 || clipping_threshold < 0.f || !workspace ||
      workspace_bytes < clip_gradient_workspace_floats(rows) * sizeof(float) ||
      (reinterpret_cast<uintptr_t>(workspace) & 3))
    return B200RNN_STATUS_INVALID_VALUE;
  if (rows == 0) return B200RNN_STATUS_SUCCESS;
  return to_status(clip_gradient_backprop(deriv, in_value, rows, cols, clipping_threshold,
                                          self_repair_clipped_proportion_threshold, self_repair_target,
                                          self_repair_scale, attempt_repair, counters_dev, decide_counters_dev,
                                          static_cast<float *>(workspace), (cudaStream_t)stream));
}

b200rnnStatus_t b200rnnGemm(int transA, int transB, int M, int N, int K, float alpha,
                            const float *A, int lda, const float *B, int ldb, float beta, float *C,
                            int ldc, const float *bias, b200rnnMath_t math, void *workspace,
                            size_t workspace_bytes, b200rnnStream_t stream) {
  if (!A || !B || !C || M < 0 || N < 0 || K < 0) return B200RNN_STATUS_INVALID_VALUE;
  GemmArgs g = {};
  g.M = M; g.N = N; g.K = K; g.alpha = alpha; g.beta = beta;
  g.A = A; g.sam = transA ? 1 : lda; g.sak = transA ? lda : 1;
  g.B = B; g.sbk = transB ? 1 : ldb; g.sbn = transB ? ldb : 1;
  g.C = C; g.ldc = ldc; g.bias_a = bias; g.bias_b = nullptr; g.nb = 0;
  g.splits = 1; g.partial = nullptr;
  if (workspace && (size_t)M * N > 0) {
    const size_t per = (size_t)M * N * sizeof(float);
    int s = (int)std::min<size_t>(workspace_bytes / per, 32);
    const int tiles = ((M + 127) / 128) * ((N + 127) / 128);
    if (s >= 2 && tiles < 148 && K >= 4096) {
      // tensor mode: the most the workspace allows, the kernel's cost model picks the count
      g.splits = math == B200RNN_MATH_TENSOR ? s : std::min(s, std::max(2, 296 / std::max(tiles, 1)));
      g.partial = static_cast<float *>(workspace);
    }
  }
  return to_status(gemm_any((int)math, g, (cudaStream_t)stream, nullptr));
}

b200rnnStatus_t b200rnnColumnSums(const float *a, int rows, int cols, int lda, float alpha,
                                  float *out, int accumulate, void *workspace,
                                  size_t workspace_bytes, b200rnnStream_t stream) {
  if (!a || !out || !workspace) return B200RNN_STATUS_INVALID_VALUE;
  cudaError_t e = column_sums(a, rows, cols, lda, alpha, out, accumulate, static_cast<float *>(workspace),
                              workspace_bytes / sizeof(float), (cudaStream_t)stream, nullptr);
  return e == cudaErrorInvalidValue ? B200RNN_STATUS_INVALID_VALUE : to_status(e);
}

double b200rnnForwardFlops(b200rnnPlan_t p, int T) {
  if (!p) return 0.0;
  double per = 0.0;
  for (int l = 0; l < p->layers; l++)
    per += (double)p->dirs * 2.0 * p->GH * ((double)din_of(p, l) + p->H);
  return per * T * p->B;
}

b200rnnStatus_t b200rnnSetProfiling(b200rnnPlan_t p, int enable) {
  if (!p) return B200RNN_STATUS_INVALID_VALUE;
  p->profiling = enable != 0;
  return B200RNN_STATUS_SUCCESS;
}

b200rnnStatus_t b200rnnGetProfile(b200rnnPlan_t p, int category, float *total_ms, int *launches) {
  if (!p || category < 0 || category > 3 || !total_ms || !launches) return B200RNN_STATUS_INVALID_VALUE;
  float tot = 0.f;
  for (size_t i = 0; i < p->ev_used[category]; i++) {
    float ms = 0.f;
    if (cudaEventSynchronize(p->ev[category][i].second) != cudaSuccess ||
        cudaEventElapsedTime(&ms, p->ev[category][i].first, p->ev[category][i].second) != cudaSuccess)
      return B200RNN_STATUS_EXECUTION_FAILED;
    tot += ms;
  }
  *total_ms = tot;
  *launches = (int)p->ev_used[category];
  p->ev_used[category] = 0;
  return B200RNN_STATUS_SUCCESS;
}

int b200rnnLastGemmUsedTensorCores(void) { return g_last_gemm_tc; }

int b200rnnLastGemmUsedCtaPair(void) { return g_last_gemm_pair; }

int b200rnnSetTuning(const char *key, int value) {
  if (key && strcmp(key, "GEMM_PAIR") == 0) {
    g_gemm_pair = value;
    return 0;
  }
  if (key && strcmp(key, "GEMM_TMA_STORE") == 0) {
    g_gemm_tma_store = value;
    return 0;
  }
  return -1;
}

int b200rnnLastLaunchCount(b200rnnPlan_t p) { return p ? p->launches : 0; }

}  // extern "C"
