/*
 * oracle/ctc_oracle.c -- CPU restatement of the CTC loss + gradient that
 * kaldi-ctc obtains from warp-ctc.  TEST INFRASTRUCTURE ONLY: imported by
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs.  The product path (kaldi_ctc_b200/csrc) never links it.
 *
 * PARITY PINNING.  warp-ctc is an un-vendored, un-pinned third-party
 * dependency of the reference (tools/extras/install_warp_ctc.sh:9-13 clones
 * lifeiteng/warp-ctc master) and the reference holds no test or golden
 * vector for this call (src/ctc/Makefile:11 "TESTFILES =").  The oracle is
 * therefore anchored on
 *   - the call-site contract, src/ctc/ctc-nnet-update.cc:180-256
 *     (blank = 0 :205, time-major rows t*B+b :284,386-388, softmax inside,
 *     gradient w.r.t. the pre-softmax activations, NLL in costs[] :246-256,
 *     gradient slab pre-zeroed by the caller :222),
 *   - warp-ctc's published CPU algorithm (Graves 2006 eq. 16 in log space,
 *     one utterance per OpenMP iteration), restated from memory, and
 *   - independent pins made in this repo: torch.nn.functional.ctc_loss in
 *     fp64 through log_softmax autograd (tests/golden/make_golden.py) and a
 *     brute-force path enumerator (tests/test_ctc_oracle.py).
 * "parity unpinned by the reference; pinned by independent implementations".
 *
 * Two instantiations are built from this file (oracle/Makefile):
 *   REAL=float   ctc_oracle_f32  -- warp-ctc's arithmetic type; also the
 *                                    timed "reference CPU path"
 *   REAL=double  ctc_oracle_f64  -- same algorithm in double: the truth both
 *                                    fp32 implementations are judged against
 *
 * Layout (as warp-ctc): activations[(t*B + b)*A + k], T = max(input_lengths).
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
#define FN ctc_oracle_f32
#endif

#define NEG_INF (-INFINITY)

static inline REAL lse2(REAL a, REAL b) {
  if (a == NEG_INF) return b;
  if (b == NEG_INF) return a;
  REAL m = a > b ? a : b;
  return m + (REAL)log1p(exp(-fabs((double)(a - b))));
}

/* One utterance.  probs: [T_max*B*A] softmax outputs (linear domain, as
 * warp-ctc's CPU path keeps them).  Returns NLL; writes gradient rows
 * t < T of this utterance when grad != NULL. */
static REAL ctc_one(const REAL *probs, REAL *grad, const int *labels, int L,
                    int T, int A, int B, int b, int blank, int *feasible) {
  const int S = 2 * L + 1;
  int repeats = 0;
  for (int i = 1; i < L; i++) repeats += labels[i] == labels[i - 1];
  *feasible = (L + repeats <= T);
  if (!*feasible) return 0; /* warp-ctc: cost 0, gradient left untouched */

  int *lp = (int *)malloc(sizeof(int) * S);
  for (int s = 0; s < S; s++) lp[s] = (s & 1) ? labels[s / 2] : blank;

  REAL *alpha = (REAL *)malloc(sizeof(REAL) * (size_t)S * T);
  REAL *beta = (REAL *)malloc(sizeof(REAL) * S);
  REAL *beta_next = (REAL *)malloc(sizeof(REAL) * S);
  REAL *acc = (REAL *)malloc(sizeof(REAL) * A);

#define LOGP(t, k) ((REAL)log((double)probs[((size_t)(t)*B + b) * A + (k)]))

  /* alpha_0: only the first blank and the first label are reachable */
  for (int s = 0; s < S; s++) alpha[s] = NEG_INF;
  alpha[0] = LOGP(0, blank);
  if (S > 1) alpha[1] = LOGP(0, lp[1]);

  for (int t = 1; t < T; t++) {
    const REAL *ap = alpha + (size_t)(t - 1) * S;
    REAL *ac = alpha + (size_t)t * S;
    /* band: states that can still reach the end / be reached from the start */
    int lo = S - 2 * (T - t);
    if (lo < 0) lo = 0;
    int hi = 2 * (t + 1);
    if (hi > S) hi = S;
    for (int s = 0; s < S; s++) ac[s] = NEG_INF;
    for (int s = lo; s < hi; s++) {
      REAL v = ap[s];
      if (s >= 1) v = lse2(v, ap[s - 1]);
      if (s >= 2 && (s & 1) && lp[s] != lp[s - 2]) v = lse2(v, ap[s - 2]);
      ac[s] = v == NEG_INF ? NEG_INF : v + LOGP(t, lp[s]);
    }
  }
  REAL loglike = alpha[(size_t)(T - 1) * S + S - 1];
  if (S > 1) loglike = lse2(loglike, alpha[(size_t)(T - 1) * S + S - 2]);

  if (grad != NULL) {
    for (int t = T - 1; t >= 0; t--) {
      const REAL *ac = alpha + (size_t)t * S;
      if (t == T - 1) {
        for (int s = 0; s < S; s++) beta[s] = NEG_INF;
        beta[S - 1] = LOGP(t, blank);
        if (S > 1) beta[S - 2] = LOGP(t, lp[S - 2]);
      } else {
        for (int s = 0; s < S; s++) {
          REAL v = beta_next[s];
          if (s + 1 < S) v = lse2(v, beta_next[s + 1]);
          if (s + 2 < S && (s & 1) && lp[s] != lp[s + 2])
            v = lse2(v, beta_next[s + 2]);
          beta[s] = v == NEG_INF ? NEG_INF : v + LOGP(t, lp[s]);
        }
      }
      /* per-label log sum of alpha*beta (both carry y_t) */
      for (int k = 0; k < A; k++) acc[k] = NEG_INF;
      for (int s = 0; s < S; s++) acc[lp[s]] = lse2(acc[lp[s]], ac[s] + beta[s]);
      for (int k = 0; k < A; k++) {
        size_t idx = ((size_t)t * B + b) * A + k;
        REAL y = probs[idx];
        if (acc[k] == NEG_INF || y == 0)
          grad[idx] = y;
        else
          grad[idx] = y - (REAL)exp((double)(acc[k] - (REAL)log((double)y) - loglike));
      }
      REAL *tmp = beta;
      beta = beta_next;
      beta_next = tmp;
    }
  }
#undef LOGP
  free(lp);
  free(alpha);
  free(beta);
  free(beta_next);
  free(acc);
  return -loglike;
}

/* Same argument meaning as warp-ctc's compute_ctc_loss with loc == CTC_CPU
 * (call site: src/ctc/ctc-nnet-update.cc:224-231), all pointers host.
 * gradients may be NULL.  Returns 0 on success, 2 on invalid arguments. */
int FN(const float *activations, REAL *gradients, const int *flat_labels,
       const int *label_lengths, const int *input_lengths, int alphabet_size,
       int minibatch, REAL *costs, int blank_label, int num_threads) {
  if (!activations || !flat_labels || !label_lengths || !input_lengths ||
      !costs || alphabet_size <= 0 || minibatch <= 0)
    return 2;
  const int A = alphabet_size, B = minibatch;
  int T = 0;
  for (int b = 0; b < B; b++)
    if (input_lengths[b] > T) T = input_lengths[b];
  const size_t n = (size_t)T * B * A;
  REAL *probs = (REAL *)malloc(sizeof(REAL) * n);
  if (!probs) return 1;
#ifdef _OPENMP
  if (num_threads > 0) omp_set_num_threads(num_threads);
#endif
  /* softmax with max subtraction, one (t,b) row at a time */
#pragma omp parallel for schedule(static)
  for (long r = 0; r < (long)T * B; r++) {
    const float *a = activations + (size_t)r * A;
    REAL *p = probs + (size_t)r * A;
    REAL m = a[0];
    for (int k = 1; k < A; k++)
      if (a[k] > m) m = a[k];
    REAL den = 0;
    for (int k = 0; k < A; k++) {
      p[k] = (REAL)exp((double)((REAL)a[k] - m));
      den += p[k];
    }
    for (int k = 0; k < A; k++) p[k] /= den;
  }
  if (gradients) memset(gradients, 0, sizeof(REAL) * n);

  int *offs = (int *)malloc(sizeof(int) * (B + 1));
  offs[0] = 0;
  for (int b = 0; b < B; b++) offs[b + 1] = offs[b] + label_lengths[b];
  int bad = 0;
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; b++) {
    const int L = label_lengths[b], Tb = input_lengths[b];
    int ok = 1;
    for (int i = 0; i < L; i++) {
      int k = flat_labels[offs[b] + i];
      if (k < 0 || k >= A || k == blank_label) ok = 0;
    }
    if (!ok || Tb <= 0 || L < 0) {
#pragma omp atomic write
      bad = 1;
      costs[b] = 0;
      continue;
    }
    int feasible;
    costs[b] = ctc_one(probs, gradients, flat_labels + offs[b], L, Tb, A, B, b,
                       blank_label, &feasible);
  }
  free(offs);
  free(probs);
  return bad ? 2 : 0;
}
