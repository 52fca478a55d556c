// oracle/ref/kaldi_cpu_path.cc -- TEST / BASELINE INFRASTRUCTURE, never on the product path.
//
// "Kaldi's CPU matrix path for the recurrent layers" (BASELINE.json north_star): the LSTM / GRU forward,
// backward-data and backward-weights of one CuDNNRecurrentComponent layer written on the REFERENCE'S OWN
// kaldi::Matrix operations -- AddMatMat (-> cblas_sgemm of the BLAS the library is linked to), Sigmoid,
// Tanh, DiffSigmoid, DiffTanh, AddMatMatElements, MulElements, AddVecToRows, AddRowSumMat -- compiled
// from /root/reference/src/{base,matrix} by oracle/ref/Makefile into oracle/_ref/libkaldi_ref_cpu.so.
// The loop structure follows the reference's own CPU/GPU-agnostic LSTM, src/nnet/nnet-lstm-projected.h
// (:398-460 forward: one hoisted x -> gates GEMM, then per time step the recurrent GEMM and the gate
// nonlinearities on row ranges; :512-650 backward: per-step derivative chain, then the weight-gradient
// GEMMs over all frames), with the equations and the packed weight blob of the cuDNN-5 call the component
// makes (src/nnet2/nnet-cudnn-component.cc:252-265, 336-408, 534-599): no
// peepholes, no projection, two bias vectors, gate order i,f,g,o / r,z,n, zero initial state, every one of
// the B sequences run for all T steps, bidirectional output [fwd h | bwd h], rows t*B + b.
// Also exposes the affine layer (nnet-component.cc:1184-1226) on the same ops, the BLAS thread control and
// the real CompressedMatrix for round trips.
#include <stdint.h>
#include <string.h>

#include <vector>

#include "matrix/compressed-matrix.h"
#include "matrix/kaldi-matrix.h"
#include "matrix/kaldi-vector.h"

extern "C" {
void openblas_set_num_threads(int);
int openblas_get_num_threads(void);
char *openblas_get_config(void);
}

namespace {

using kaldi::kNoTrans;
using kaldi::kTrans;
using kaldi::Matrix;
using kaldi::MatrixBase;
using kaldi::SubMatrix;
using kaldi::SubVector;
using kaldi::Vector;
typedef float BaseFloat;

struct DirParams {  // one pseudo-layer of the cuDNN-v5 packed blob
  size_t w_in, w_rec, b_in, b_rec;
};

inline int GatesOf(int mode) { return mode == 2 ? 4 : (mode == 3 ? 3 : 1); }

void Layout(int mode, int dirs, int D, int H, std::vector<DirParams> *pl, size_t *total) {
  const size_t G = GatesOf(mode);
  size_t off = 0;
  pl->resize(dirs);
  for (int d = 0; d < dirs; d++) {
    (*pl)[d].w_in = off;
    off += G * H * D;
    (*pl)[d].w_rec = off;
    off += G * H * H;
  }
  for (int d = 0; d < dirs; d++) {
    (*pl)[d].b_in = off;
    off += G * H;
    (*pl)[d].b_rec = off;
    off += G * H;
  }
  *total = off;
}

// ---- LSTM, one direction ----------------------------------------------------------------------
// gates buffer columns: [i | f | g | o] (the blob's row order, so x.Wi^T lands there directly)
void LstmForward(int T, int B, int D, int H, bool reverse, const SubMatrix<BaseFloat> &x,
                 const SubMatrix<BaseFloat> &Wi, const SubMatrix<BaseFloat> &R, const SubVector<BaseFloat> &bW,
                 const SubVector<BaseFloat> &bR, Matrix<BaseFloat> *gates, Matrix<BaseFloat> *cell,
                 Matrix<BaseFloat> *tanh_c, SubMatrix<BaseFloat> y /* [T*B x H] view with the layer's stride */) {
  gates->Resize(T * B, 4 * H, kaldi::kUndefined);
  cell->Resize(T * B, H, kaldi::kSetZero);  // (AddMatMatElements with beta = 0 still reads its destination)
  tanh_c->Resize(T * B, H, kaldi::kSetZero);
  gates->AddMatMat(1.0, x, kNoTrans, Wi, kTrans, 0.0);  // hoisted input projection, all frames at once
  gates->AddVecToRows(1.0, bW);
  gates->AddVecToRows(1.0, bR);
  for (int s = 0; s < T; s++) {
    const int t = reverse ? T - 1 - s : s, tp = reverse ? t + 1 : t - 1;
    SubMatrix<BaseFloat> g_all(gates->RowRange(t * B, B));
    if (s > 0) g_all.AddMatMat(1.0, y.RowRange(tp * B, B), kNoTrans, R, kTrans, 1.0);  // h(t-1) -> i,f,g,o
    SubMatrix<BaseFloat> gi(g_all.ColRange(0, H)), gf(g_all.ColRange(H, H)), gg(g_all.ColRange(2 * H, H)),
        go(g_all.ColRange(3 * H, H));
    SubMatrix<BaseFloat> gif(g_all.ColRange(0, 2 * H));
    gif.Sigmoid(gif);
    gg.Tanh(gg);
    go.Sigmoid(go);
    SubMatrix<BaseFloat> c(cell->RowRange(t * B, B)), tc(tanh_c->RowRange(t * B, B)), h(y.RowRange(t * B, B));
    c.AddMatMatElements(1.0, gg, gi, 0.0);                                    // g * i -> c
    if (s > 0) c.AddMatMatElements(1.0, cell->RowRange(tp * B, B), gf, 1.0);  // c(t-1) * f -> c
    tc.Tanh(c);
    h.AddMatMatElements(1.0, tc, go, 0.0);  // h = o * tanh(c)
  }
}

// dgates overwrites `gates` in place (pre-activation derivatives); returns dx / dW contributions by GEMM
void LstmBackward(int T, int B, int D, int H, bool reverse, const SubMatrix<BaseFloat> &x,
                  const SubMatrix<BaseFloat> &Wi, const SubMatrix<BaseFloat> &R, Matrix<BaseFloat> *gates,
                  const Matrix<BaseFloat> &cell, const Matrix<BaseFloat> &tanh_c, const SubMatrix<BaseFloat> &y,
                  const SubMatrix<BaseFloat> &dy, SubMatrix<BaseFloat> dx, bool dx_accumulate,
                  SubMatrix<BaseFloat> dWi, SubMatrix<BaseFloat> dR, SubVector<BaseFloat> dbW,
                  SubVector<BaseFloat> dbR) {
  Matrix<BaseFloat> dh(B, H), dc(B, H), tmp(B, H), dc_next(B, H);
  for (int s = T - 1; s >= 0; s--) {
    const int t = reverse ? T - 1 - s : s, tp = reverse ? t + 1 : t - 1, tn = reverse ? t - 1 : t + 1;
    SubMatrix<BaseFloat> g_all(gates->RowRange(t * B, B));
    SubMatrix<BaseFloat> gi(g_all.ColRange(0, H)), gf(g_all.ColRange(H, H)), gg(g_all.ColRange(2 * H, H)),
        go(g_all.ColRange(3 * H, H));
    dh.CopyFromMat(dy.RowRange(t * B, B));
    if (s < T - 1)  // recurrent path: dh += dgates(t+1) . R
      dh.AddMatMat(1.0, gates->RowRange(tn * B, B), kNoTrans, R, kNoTrans, 1.0);
    const SubMatrix<BaseFloat> tc(tanh_c.RowRange(t * B, B));
    // dc = dh * o * (1 - tanh(c)^2) + dc(t+1) * f(t+1)
    tmp.CopyFromMat(dh);
    tmp.MulElements(go);
    dc.DiffTanh(tc, tmp);
    if (s < T - 1) dc.AddMat(1.0, dc_next);
    // o: do = dh * tanh(c) -> through the sigmoid
    tmp.CopyFromMat(dh);
    tmp.MulElements(tc);
    // (the gate values are still needed below: compute all pre-activation derivatives into temporaries)
    Matrix<BaseFloat> d_o(B, H, kaldi::kUndefined), d_i(B, H, kaldi::kUndefined), d_f(B, H, kaldi::kUndefined),
        d_g(B, H, kaldi::kUndefined);
    d_o.DiffSigmoid(go, tmp);
    tmp.CopyFromMat(dc);
    tmp.MulElements(gg);
    d_i.DiffSigmoid(gi, tmp);  // di = dc * g
    tmp.CopyFromMat(dc);
    tmp.MulElements(gi);
    d_g.DiffTanh(gg, tmp);     // dg = dc * i
    if (s > 0) {
      tmp.CopyFromMat(dc);
      tmp.MulElements(cell.RowRange(tp * B, B));
      d_f.DiffSigmoid(gf, tmp);  // df = dc * c(t-1)
    } else {
      d_f.SetZero();
    }
    // what the previous step receives through the cell: dc * f
    dc_next.CopyFromMat(dc);
    dc_next.MulElements(gf);
    gi.CopyFromMat(d_i);
    gf.CopyFromMat(d_f);
    gg.CopyFromMat(d_g);
    go.CopyFromMat(d_o);
  }
  // all frames at once: dx, dWi, dR, biases
  dx.AddMatMat(1.0, *gates, kNoTrans, Wi, kNoTrans, dx_accumulate ? 1.0 : 0.0);
  dWi.AddMatMat(1.0, *gates, kTrans, x, kNoTrans, 1.0);
  if (T > 1) {  // h_prev(t) = y(t -+ 1): a row shift of B
    const int n = (T - 1) * B;
    if (!reverse) dR.AddMatMat(1.0, gates->RowRange(B, n), kTrans, y.RowRange(0, n), kNoTrans, 1.0);
    else dR.AddMatMat(1.0, gates->RowRange(0, n), kTrans, y.RowRange(B, n), kNoTrans, 1.0);
  }
  dbW.AddRowSumMat(1.0, *gates, 1.0);
  dbR.AddRowSumMat(1.0, *gates, 1.0);
}

// ---- GRU, one direction -----------------------------------------------------------------------
// gates buffer columns: [r | z | n]; q = R_n h(t-1) + bR_n kept per frame for the backward pass
void GruForward(int T, int B, int D, int H, bool reverse, const SubMatrix<BaseFloat> &x,
                const SubMatrix<BaseFloat> &Wi, const SubMatrix<BaseFloat> &R, const SubVector<BaseFloat> &bW,
                const SubVector<BaseFloat> &bR, Matrix<BaseFloat> *gates, Matrix<BaseFloat> *q,
                SubMatrix<BaseFloat> y) {
  gates->Resize(T * B, 3 * H, kaldi::kUndefined);
  q->Resize(T * B, H, kaldi::kSetZero);
  gates->AddMatMat(1.0, x, kNoTrans, Wi, kTrans, 0.0);
  gates->AddVecToRows(1.0, bW);
  SubMatrix<BaseFloat> rz_all(gates->ColRange(0, 2 * H));
  rz_all.AddVecToRows(1.0, bR.Range(0, 2 * H));
  q->AddVecToRows(1.0, bR.Range(2 * H, H));
  const SubMatrix<BaseFloat> Rrz(R.RowRange(0, 2 * H)), Rn(R.RowRange(2 * H, H));
  Matrix<BaseFloat> tmp(B, H);
  for (int s = 0; s < T; s++) {
    const int t = reverse ? T - 1 - s : s, tp = reverse ? t + 1 : t - 1;
    SubMatrix<BaseFloat> g_all(gates->RowRange(t * B, B));
    SubMatrix<BaseFloat> rz(g_all.ColRange(0, 2 * H)), gr(g_all.ColRange(0, H)), gz(g_all.ColRange(H, H)),
        gn(g_all.ColRange(2 * H, H));
    SubMatrix<BaseFloat> qt(q->RowRange(t * B, B)), h(y.RowRange(t * B, B));
    if (s > 0) {
      rz.AddMatMat(1.0, y.RowRange(tp * B, B), kNoTrans, Rrz, kTrans, 1.0);
      qt.AddMatMat(1.0, y.RowRange(tp * B, B), kNoTrans, Rn, kTrans, 1.0);
    }
    rz.Sigmoid(rz);
    gn.AddMatMatElements(1.0, gr, qt, 1.0);  // n_pre = W_n x + bW_n + r * (R_n h + bR_n)
    gn.Tanh(gn);
    // h = (1 - z) * n + z * h(t-1) = n + z * (h(t-1) - n)
    h.CopyFromMat(gn);
    tmp.CopyFromMat(gn);
    tmp.Scale(-1.0);
    if (s > 0) tmp.AddMat(1.0, y.RowRange(tp * B, B));
    h.AddMatMatElements(1.0, gz, tmp, 1.0);
  }
}

void GruBackward(int T, int B, int D, int H, bool reverse, const SubMatrix<BaseFloat> &x,
                 const SubMatrix<BaseFloat> &Wi, const SubMatrix<BaseFloat> &R, Matrix<BaseFloat> *gates,
                 Matrix<BaseFloat> *q, const SubMatrix<BaseFloat> &y, const SubMatrix<BaseFloat> &dy,
                 SubMatrix<BaseFloat> dx, bool dx_accumulate, SubMatrix<BaseFloat> dWi, SubMatrix<BaseFloat> dR,
                 SubVector<BaseFloat> dbW, SubVector<BaseFloat> dbR) {
  const SubMatrix<BaseFloat> Rrz(R.RowRange(0, 2 * H)), Rn(R.RowRange(2 * H, H));
  Matrix<BaseFloat> dh(B, H), dh_carry(B, H), tmp(B, H), d_r(B, H), d_z(B, H), d_n(B, H), d_q(B, H);
  for (int s = T - 1; s >= 0; s--) {
    const int t = reverse ? T - 1 - s : s, tp = reverse ? t + 1 : t - 1, tn = reverse ? t - 1 : t + 1;
    SubMatrix<BaseFloat> g_all(gates->RowRange(t * B, B));
    SubMatrix<BaseFloat> gr(g_all.ColRange(0, H)), gz(g_all.ColRange(H, H)), gn(g_all.ColRange(2 * H, H));
    SubMatrix<BaseFloat> qt(q->RowRange(t * B, B));
    dh.CopyFromMat(dy.RowRange(t * B, B));
    if (s < T - 1) {  // from step t+1: through z (carry), through the r,z GEMM and through the n-gate GEMM
      dh.AddMat(1.0, dh_carry);
      dh.AddMatMat(1.0, gates->RowRange(tn * B, B).ColRange(0, 2 * H), kNoTrans, Rrz, kNoTrans, 1.0);
      dh.AddMatMat(1.0, q->RowRange(tn * B, B), kNoTrans, Rn, kNoTrans, 1.0);
    }
    // dz = dh * (h(t-1) - n)
    tmp.CopyFromMat(gn);
    tmp.Scale(-1.0);
    if (s > 0) tmp.AddMat(1.0, y.RowRange(tp * B, B));
    tmp.MulElements(dh);
    d_z.DiffSigmoid(gz, tmp);
    // dn = dh * (1 - z) -> through tanh
    tmp.CopyFromMat(dh);
    tmp.MulElements(gz);
    tmp.Scale(-1.0);
    tmp.AddMat(1.0, dh);
    d_n.DiffTanh(gn, tmp);
    // dq = dn_pre * r ;  dr = dn_pre * q -> through the sigmoid
    d_q.CopyFromMat(d_n);
    d_q.MulElements(gr);
    tmp.CopyFromMat(d_n);
    tmp.MulElements(qt);
    d_r.DiffSigmoid(gr, tmp);
    // carry to h(t-1) through z
    dh_carry.CopyFromMat(dh);
    dh_carry.MulElements(gz);
    gr.CopyFromMat(d_r);
    gz.CopyFromMat(d_z);
    gn.CopyFromMat(d_n);
    qt.CopyFromMat(d_q);
  }
  dx.AddMatMat(1.0, *gates, kNoTrans, Wi, kNoTrans, dx_accumulate ? 1.0 : 0.0);
  dWi.AddMatMat(1.0, *gates, kTrans, x, kNoTrans, 1.0);
  if (T > 1) {
    const int n = (T - 1) * B;
    const int go = reverse ? 0 : B, yo = reverse ? B : 0;
    SubMatrix<BaseFloat> dRrz(dR.RowRange(0, 2 * H)), dRn(dR.RowRange(2 * H, H));
    dRrz.AddMatMat(1.0, gates->RowRange(go, n).ColRange(0, 2 * H), kTrans, y.RowRange(yo, n), kNoTrans, 1.0);
    dRn.AddMatMat(1.0, q->RowRange(go, n), kTrans, y.RowRange(yo, n), kNoTrans, 1.0);
  }
  dbW.AddRowSumMat(1.0, *gates, 1.0);
  SubVector<BaseFloat> dbR_rz(dbR.Range(0, 2 * H)), dbR_n(dbR.Range(2 * H, H));
  dbR_rz.AddRowSumMat(1.0, gates->ColRange(0, 2 * H), 1.0);
  dbR_n.AddRowSumMat(1.0, *q, 1.0);
}

}  // namespace

extern "C" {

void kaldiref_set_num_threads(int n) { openblas_set_num_threads(n); }
int kaldiref_get_num_threads(void) { return openblas_get_num_threads(); }
const char *kaldiref_blas_config(void) { return openblas_get_config(); }

size_t kaldiref_rnn_param_count(int mode, int bidir, int D, int H) {
  std::vector<DirParams> pl;
  size_t total;
  Layout(mode, bidir ? 2 : 1, D, H, &pl, &total);
  return total;
}

// One CuDNNRecurrentComponent layer (num-layers=1), mode 2 (LSTM) or 3 (GRU).  x [T*B x D], w blob,
// y [T*B x H*dirs] out.  If dy != NULL: dx [T*B x D] out and dw (blob-shaped) ACCUMULATED into.
// Returns 0, or -1 for an unsupported mode.
int kaldiref_rnn_layer(int mode, int bidir, int T, int B, int D, int H, const float *x_, const float *w_,
                       float *y_, const float *dy_, float *dx_, float *dw_) {
  if (mode != 2 && mode != 3) return -1;
  const int dirs = bidir ? 2 : 1, G = GatesOf(mode), HO = H * dirs;
  std::vector<DirParams> pl;
  size_t total;
  Layout(mode, dirs, D, H, &pl, &total);
  float *w = const_cast<float *>(w_);
  SubMatrix<BaseFloat> x(const_cast<float *>(x_), T * B, D, D);
  for (int d = 0; d < dirs; d++) {
    SubMatrix<BaseFloat> Wi(w + pl[d].w_in, G * H, D, D), R(w + pl[d].w_rec, G * H, H, H);
    SubVector<BaseFloat> bW(w + pl[d].b_in, G * H), bR(w + pl[d].b_rec, G * H);
    SubMatrix<BaseFloat> y(y_ + d * H, T * B, H, HO);
    Matrix<BaseFloat> gates, cell, tanh_c;
    if (mode == 2) LstmForward(T, B, D, H, d == 1, x, Wi, R, bW, bR, &gates, &cell, &tanh_c, y);
    else GruForward(T, B, D, H, d == 1, x, Wi, R, bW, bR, &gates, &cell, y);
    if (!dy_) continue;
    SubMatrix<BaseFloat> dy(const_cast<float *>(dy_) + d * H, T * B, H, HO), dx(dx_, T * B, D, D);
    SubMatrix<BaseFloat> dWi(dw_ + pl[d].w_in, G * H, D, D), dR(dw_ + pl[d].w_rec, G * H, H, H);
    SubVector<BaseFloat> dbW(dw_ + pl[d].b_in, G * H), dbR(dw_ + pl[d].b_rec, G * H);
    if (mode == 2) LstmBackward(T, B, D, H, d == 1, x, Wi, R, &gates, cell, tanh_c, y, dy, dx, d > 0, dWi, dR, dbW, dbR);
    else GruBackward(T, B, D, H, d == 1, x, Wi, R, &gates, &cell, y, dy, dx, d > 0, dWi, dR, dbW, dbR);
  }
  return 0;
}

// AffineComponent on kaldi::Matrix (nnet-component.cc:1184-1226): out = in W^T + b;
// with deriv != NULL also in_deriv = deriv W and the UpdateSimple step W += lr deriv^T in, b += lr colsum(deriv).
void kaldiref_affine(int rows, int K, int N, const float *in_, float *W_, float *b_, float *out_,
                     const float *deriv_, float *in_deriv_, float lr) {
  SubMatrix<BaseFloat> in(const_cast<float *>(in_), rows, K, K), W(W_, N, K, K);
  SubVector<BaseFloat> b(b_, N);
  if (out_) {
    SubMatrix<BaseFloat> out(out_, rows, N, N);
    out.CopyRowsFromVec(b);
    out.AddMatMat(1.0, in, kNoTrans, W, kTrans, 1.0);
  }
  if (deriv_) {
    SubMatrix<BaseFloat> deriv(const_cast<float *>(deriv_), rows, N, N), in_deriv(in_deriv_, rows, K, K);
    in_deriv.AddMatMat(1.0, deriv, kNoTrans, W, kNoTrans, 0.0);
    b.AddRowSumMat(lr, deriv, 1.0);
    W.AddMatMat(lr, deriv, kTrans, in, kNoTrans, 1.0);
  }
}

// The real CompressedMatrix: compress a [rows x cols] float matrix and decompress it again
// (compressed-matrix.cc:41-121, 493-529).  back: [rows x cols] out.
void kaldiref_compress_roundtrip(const float *m_, int rows, int cols, float *back_) {
  SubMatrix<BaseFloat> m(const_cast<float *>(m_), rows, cols, cols), back(back_, rows, cols, cols);
  kaldi::CompressedMatrix cm(m);
  cm.CopyToMat(&back);
}

}  // extern "C"
