/*
 * oracle/rnn_oracle.c -- CPU restatement of the recurrent forward/backward
 * that kaldi-ctc's nnet2 CuDNNRecurrentComponent obtains from cuDNN 5.
 * TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline /
 * --impl reference).  The product path never links it.
 *
 * PARITY PINNING.  cuDNN 5.x is a closed, un-vendored library; the v5 RNN API
 * no longer exists in the cuDNN 9 of this image, and the reference has no
 * test for the component (absent from src/nnet2/nnet-component-test.cc:872-902).
 * "parity unpinned by the reference"; the restatement follows
 *   - the call sites src/nnet2/nnet-cudnn-component.cc:508-556 (Propagate:
 *     hx = cx = 0, every one of the B sequences is run for all T steps, row
 *     index t*B + b, bidirectional output = [fwd h_t | bwd h_t]) and
 *     :558-610 (Backprop: dhy = dcy = 0, dW accumulated into a zeroed blob),
 *   - the weight blob as the reference locates it through
 *     cudnnGetRNNLinLayer{Matrix,Bias}Params (:336-408): per pseudo-layer
 *     p = layer*dirs + dir, nlin matrices (input ones first, then recurrent
 *     ones, each row-major [H x in]), all matrices of all pseudo-layers
 *     first, then per pseudo-layer nlin bias vectors of H,
 *   - cuDNN's documented cell equations (gate order LSTM i,f,g,o; GRU r,z,n;
 *     two bias vectors per gate; GRU's recurrent n-bias sits inside r*(...)),
 * and is pinned independently against torch.nn.LSTM/GRU/RNN in fp64
 * (tests/golden/make_golden.py), by finite differences (tests/test_rnn_oracle.py),
 * against the same layers written on the reference's own kaldi::Matrix operations
 * (oracle/ref/kaldi_cpu_path.cc, compiled from /root/reference by oracle/ref/Makefile)
 * and -- through the product's exact fp32 mode, which matches this file to 1e-6 --
 * against real cuDNN 9 on the GPU (tests/test_cudnn9_crosscheck_gpu.py).
 *
 * Built twice (oracle/Makefile): REAL=float rnn_oracle_f32 (the reference's
 * arithmetic type, also the timed CPU baseline) and REAL=double
 * rnn_oracle_f64.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL float
#endif
#ifndef FN
#define FN(x) x##_f32
#endif

static int nlin_of(int mode) { return mode == 2 ? 8 : mode == 3 ? 6 : 2; }

/* number of floats in the packed blob (cudnnGetRNNParamsSize / sizeof(float)) */
long FN(rnn_oracle_param_count)(int mode, int bidir, int layers, int D, int H) {
  const int dirs = bidir ? 2 : 1, ng = nlin_of(mode) / 2;
  long n = 0;
  for (int l = 0; l < layers; l++) {
    long in = l == 0 ? D : (long)H * dirs;
    n += (long)dirs * ((long)ng * H * in + (long)ng * H * H + 2L * ng * H);
  }
  return n;
}

/* offset of linear layer `lin` (0..nlin-1) of pseudo-layer p; is_bias selects
 * the bias vector.  rows/cols describe the matrix (cols = 1 for a bias). */
long FN(rnn_oracle_locate)(int mode, int bidir, int layers, int D, int H, int p,
                           int lin, int is_bias, int *rows, int *cols) {
  const int dirs = bidir ? 2 : 1, nlin = nlin_of(mode), ng = nlin / 2;
  long off = 0;
  if (!is_bias) {
    for (int q = 0; q < p; q++) {
      long in = (q / dirs) == 0 ? D : (long)H * dirs;
      off += (long)ng * H * in + (long)ng * H * H;
    }
    long in = (p / dirs) == 0 ? D : (long)H * dirs;
    if (lin < ng) {
      off += (long)lin * H * in;
      *rows = H;
      *cols = (int)in;
    } else {
      off += (long)ng * H * in + (long)(lin - ng) * H * H;
      *rows = H;
      *cols = H;
    }
    return off;
  }
  for (int q = 0; q < layers * dirs; q++) {
    long in = (q / dirs) == 0 ? D : (long)H * dirs;
    off += (long)ng * H * in + (long)ng * H * H;
  }
  off += (long)p * nlin * H + (long)lin * H;
  *rows = H;
  *cols = 1;
  return off;
}

static inline REAL sigm(REAL x) { return (REAL)(1.0 / (1.0 + exp(-(double)x))); }

/* C[M x N] (+)= A[M x K] * B^T, B is [N x K]; row-major, leading dims given */
static void gemm_nt(int M, int N, int K, const REAL *A, int lda, const REAL *Bm,
                    int ldb, REAL *C, int ldc, int accumulate) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; i++) {
    for (int j = 0; j < N; j++) {
      const REAL *a = A + (size_t)i * lda, *b = Bm + (size_t)j * ldb;
      REAL s = 0;
      for (int k = 0; k < K; k++) s += a[k] * b[k];
      if (accumulate)
        C[(size_t)i * ldc + j] += s;
      else
        C[(size_t)i * ldc + j] = s;
    }
  }
}
/* C[M x N] (+)= A[M x K] * B, B is [K x N] */
static void gemm_nn(int M, int N, int K, const REAL *A, int lda, const REAL *Bm,
                    int ldb, REAL *C, int ldc, int accumulate) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; i++) {
    REAL *c = C + (size_t)i * ldc;
    if (!accumulate)
      for (int j = 0; j < N; j++) c[j] = 0;
    for (int k = 0; k < K; k++) {
      REAL a = A[(size_t)i * lda + k];
      const REAL *b = Bm + (size_t)k * ldb;
      for (int j = 0; j < N; j++) c[j] += a * b[j];
    }
  }
}
/* C[M x N] += A^T * B, A is [K x M], B is [K x N] */
static void gemm_tn_acc(int M, int N, int K, const REAL *A, int lda,
                        const REAL *Bm, int ldb, REAL *C, int ldc) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; i++) {
    REAL *c = C + (size_t)i * ldc;
    for (int k = 0; k < K; k++) {
      REAL a = A[(size_t)k * lda + i];
      if (a == 0) continue;
      const REAL *b = Bm + (size_t)k * ldb;
      for (int j = 0; j < N; j++) c[j] += a * b[j];
    }
  }
}

typedef struct {
  REAL *act;  /* [T*B x ng*H] gate activations (post non-linearity)      */
  REAL *cell; /* LSTM: c_t [T*B x H]; GRU: q_t = R_n h + bR_n [T*B x H]  */
} reserve_t;

/*
 * x  [T*B x D]  float input (row t*B+b)            w  packed blob (float)
 * y  [T*B x H*dirs] output of the LAST layer
 * dy NULL => forward only.  Otherwise dx [T*B x D] and dw (blob-shaped,
 * ACCUMULATED into, caller zeroes) are produced (either may be NULL).
 * num_threads <= 0 keeps the OpenMP default.
 */
int FN(rnn_oracle)(int mode, int bidir, int layers, int D, int H, int B, int T,
                   const float *x, const float *w, REAL *y, const REAL *dy,
                   REAL *dx, REAL *dw, int num_threads) {
  if (mode < 0 || mode > 3 || layers < 1 || D < 1 || H < 1 || B < 1 || T < 1)
    return 2;
#ifdef _OPENMP
  if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
  const int dirs = bidir ? 2 : 1, nlin = nlin_of(mode), ng = nlin / 2;
  const int HO = H * dirs, GH = ng * H;
  const size_t TB = (size_t)T * B;
  const long np = FN(rnn_oracle_param_count)(mode, bidir, layers, D, H);
  REAL *W = (REAL *)malloc(sizeof(REAL) * np);
  for (long i = 0; i < np; i++) W[i] = (REAL)w[i];

  /* layer inputs: in[0] = x, in[l] = output of layer l-1 */
  REAL **in = (REAL **)calloc(layers + 1, sizeof(REAL *));
  in[0] = (REAL *)malloc(sizeof(REAL) * TB * D);
  for (size_t i = 0; i < TB * D; i++) in[0][i] = (REAL)x[i];
  reserve_t *res = (reserve_t *)calloc((size_t)layers * dirs, sizeof(reserve_t));
  REAL *pre = (REAL *)malloc(sizeof(REAL) * TB * GH);
  REAL *rec = (REAL *)malloc(sizeof(REAL) * (size_t)B * GH);
  REAL *hzero = (REAL *)calloc((size_t)B * H, sizeof(REAL));
  int r_, c_;

  for (int l = 0; l < layers; l++) {
    const int Din = l == 0 ? D : HO;
    in[l + 1] = (REAL *)calloc(TB * HO, sizeof(REAL));
    for (int d = 0; d < dirs; d++) {
      const int p = l * dirs + d;
      reserve_t *rs = &res[p];
      rs->act = (REAL *)malloc(sizeof(REAL) * TB * GH);
      rs->cell = (REAL *)calloc(TB * H, sizeof(REAL));
      const REAL *Wi = W + FN(rnn_oracle_locate)(mode, bidir, layers, D, H, p, 0, 0, &r_, &c_);
      const REAL *Rw = W + FN(rnn_oracle_locate)(mode, bidir, layers, D, H, p, ng, 0, &r_, &c_);
      const REAL *bW = W + FN(rnn_oracle_locate)(mode, bidir, layers, D, H, p, 0, 1, &r_, &c_);
      const REAL *bR = bW + (size_t)ng * H;
      /* hoisted input projection: pre = in * Wi^T  (Wi is [ng*H x Din], the
       * ng input matrices are contiguous in the blob) */
      gemm_nt((int)TB, GH, Din, in[l], Din, Wi, Din, pre, GH, 0);
      for (int step = 0; step < T; step++) {
        const int t = d == 0 ? step : T - 1 - step;
        const int tp = d == 0 ? t - 1 : t + 1;
        const REAL *hprev = step == 0 ? hzero : in[l + 1] + (size_t)tp * B * HO + d * H;
        const int ldh = step == 0 ? H : HO;
        gemm_nt(B, GH, H, hprev, ldh, Rw, H, rec, GH, 0);
        for (int b = 0; b < B; b++) {
          const size_t row = (size_t)t * B + b;
          const REAL *pr = pre + row * GH, *rc = rec + (size_t)b * GH;
          REAL *a = rs->act + row * GH;
          REAL *ho = in[l + 1] + row * HO + d * H;
          for (int j = 0; j < H; j++) {
            if (mode == 2) {
              REAL gi = sigm(pr[j] + rc[j] + bW[j] + bR[j]);
              REAL gf = sigm(pr[H + j] + rc[H + j] + bW[H + j] + bR[H + j]);
              REAL gg = (REAL)tanh((double)(pr[2 * H + j] + rc[2 * H + j] + bW[2 * H + j] + bR[2 * H + j]));
              REAL go = sigm(pr[3 * H + j] + rc[3 * H + j] + bW[3 * H + j] + bR[3 * H + j]);
              REAL cp = step == 0 ? 0 : rs->cell[((size_t)tp * B + b) * H + j];
              REAL c = gf * cp + gi * gg;
              rs->cell[row * H + j] = c;
              a[j] = gi; a[H + j] = gf; a[2 * H + j] = gg; a[3 * H + j] = go;
              ho[j] = go * (REAL)tanh((double)c);
            } else if (mode == 3) {
              REAL gr = sigm(pr[j] + rc[j] + bW[j] + bR[j]);
              REAL gz = sigm(pr[H + j] + rc[H + j] + bW[H + j] + bR[H + j]);
              REAL q = rc[2 * H + j] + bR[2 * H + j];
              REAL gn = (REAL)tanh((double)(pr[2 * H + j] + bW[2 * H + j] + gr * q));
              rs->cell[row * H + j] = q;
              a[j] = gr; a[H + j] = gz; a[2 * H + j] = gn;
              ho[j] = (1 - gz) * gn + gz * hprev[(size_t)b * ldh + j];
            } else {
              REAL v = pr[j] + rc[j] + bW[j] + bR[j];
              REAL h = mode == 0 ? (v > 0 ? v : 0) : (REAL)tanh((double)v);
              a[j] = h;
              ho[j] = h;
            }
          }
        }
      }
    }
  }
  if (y) memcpy(y, in[layers], sizeof(REAL) * TB * HO);

  if (dy != NULL) {
    REAL *dout = (REAL *)malloc(sizeof(REAL) * TB * HO); /* d(layer output) */
    memcpy(dout, dy, sizeof(REAL) * TB * HO);
    REAL *dgi = (REAL *)malloc(sizeof(REAL) * TB * GH); /* input-side dgates  */
    REAL *dgr = (REAL *)malloc(sizeof(REAL) * TB * GH); /* recurrent-side    */
    REAL *dh = (REAL *)malloc(sizeof(REAL) * (size_t)B * H);
    REAL *dc = (REAL *)malloc(sizeof(REAL) * (size_t)B * H);
    REAL *hprev_all = (REAL *)malloc(sizeof(REAL) * TB * H);
    for (int l = layers - 1; l >= 0; l--) {
      const int Din = l == 0 ? D : HO;
      REAL *din = (REAL *)calloc(TB * Din, sizeof(REAL));
      for (int d = 0; d < dirs; d++) {
        const int p = l * dirs + d;
        reserve_t *rs = &res[p];
        long oWi = FN(rnn_oracle_locate)(mode, bidir, layers, D, H, p, 0, 0, &r_, &c_);
        long oR = FN(rnn_oracle_locate)(mode, bidir, layers, D, H, p, ng, 0, &r_, &c_);
        long obW = FN(rnn_oracle_locate)(mode, bidir, layers, D, H, p, 0, 1, &r_, &c_);
        const REAL *Wi = W + oWi, *Rw = W + oR;
        memset(dh, 0, sizeof(REAL) * (size_t)B * H);
        memset(dc, 0, sizeof(REAL) * (size_t)B * H);
        for (int step = T - 1; step >= 0; step--) {
          const int t = d == 0 ? step : T - 1 - step;
          const int tp = d == 0 ? t - 1 : t + 1;
          for (int b = 0; b < B; b++) {
            const size_t row = (size_t)t * B + b;
            const REAL *a = rs->act + row * GH;
            REAL *gi_ = dgi + row * GH, *gr_ = dgr + row * GH;
            for (int j = 0; j < H; j++) {
              REAL dht = dout[row * HO + d * H + j] + dh[(size_t)b * H + j];
              REAL hp = step == 0 ? 0 : in[l + 1][((size_t)tp * B + b) * HO + d * H + j];
              hprev_all[row * H + j] = hp;
              if (mode == 2) {
                REAL i = a[j], f = a[H + j], g = a[2 * H + j], o = a[3 * H + j];
                REAL c = rs->cell[row * H + j];
                REAL cp = step == 0 ? 0 : rs->cell[((size_t)tp * B + b) * H + j];
                REAL tc = (REAL)tanh((double)c);
                REAL dct = dht * o * (1 - tc * tc) + dc[(size_t)b * H + j];
                gi_[j] = dct * g * i * (1 - i);
                gi_[H + j] = dct * cp * f * (1 - f);
                gi_[2 * H + j] = dct * i * (1 - g * g);
                gi_[3 * H + j] = dht * tc * o * (1 - o);
                dc[(size_t)b * H + j] = dct * f;
                dh[(size_t)b * H + j] = 0;
              } else if (mode == 3) {
                REAL r = a[j], z = a[H + j], n = a[2 * H + j];
                REAL q = rs->cell[row * H + j];
                REAL dn = dht * (1 - z) * (1 - n * n);
                gi_[j] = dn * q * r * (1 - r);
                gi_[H + j] = dht * (hp - n) * z * (1 - z);
                gi_[2 * H + j] = dn;
                gr_[j] = gi_[j];
                gr_[H + j] = gi_[H + j];
                gr_[2 * H + j] = dn * r;
                dh[(size_t)b * H + j] = dht * z;
              } else {
                REAL h = a[j];
                gi_[j] = dht * (mode == 0 ? (h > 0 ? 1 : 0) : (1 - h * h));
                dh[(size_t)b * H + j] = 0;
              }
            }
          }
          /* dh_{prev} += dgates_rec(t) * R   ([B x GH] * [GH x H]) */
          const REAL *gsrc = (mode == 3 ? dgr : dgi) + (size_t)t * B * GH;
          gemm_nn(B, H, GH, gsrc, GH, Rw, H, dh, H, 1);
        }
        const REAL *grec = mode == 3 ? dgr : dgi;
        /* d(layer input) += dgi * Wi */
        gemm_nn((int)TB, Din, GH, dgi, GH, Wi, Din, din, Din, 1);
        if (dw) {
          gemm_tn_acc(GH, Din, (int)TB, dgi, GH, in[l], Din, dw + oWi, Din);
          gemm_tn_acc(GH, H, (int)TB, grec, GH, hprev_all, H, dw + oR, H);
          for (size_t rI = 0; rI < TB; rI++)
            for (int j = 0; j < GH; j++) {
              dw[obW + j] += dgi[rI * GH + j];
              dw[obW + GH + j] += grec[rI * GH + j];
            }
        }
      }
      free(dout);
      dout = din; /* [TB x Din]; for l > 0, Din == HO */
    }
    if (dx) memcpy(dx, dout, sizeof(REAL) * TB * D);
    free(dout); free(dgi); free(dgr); free(dh); free(dc); free(hprev_all);
  }

  for (int p = 0; p < layers * dirs; p++) { free(res[p].act); free(res[p].cell); }
  for (int l = 0; l <= layers; l++) free(in[l]);
  free(in); free(res); free(pre); free(rec); free(hzero); free(W);
  return 0;
}
