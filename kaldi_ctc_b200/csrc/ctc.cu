// kaldi_ctc_b200/csrc/ctc.cu
//
// CTC loss + gradient for sm_100a behind the warp-ctc C ABI (include/ctc.h).
// Replaces the library call at src/ctc/ctc-nnet-update.cc:211-243 of the
// reference (get_workspace_size + compute_ctc_loss).
//
// Three kernels per call, all on the caller's stream:
//   K1 ctc_rowstats_gather   one warp per (t,b) row: online log-sum-exp of the
//        row (one HBM read of the activations), and a gather of the emission
//        log-probabilities of the utterance's own lattice symbols (blank + its
//        L labels) into a compact per-utterance table E[t][u], base-2 logs.
//   K2 ctc_alpha_beta<P>     one CTA per utterance.  Two warp groups run the
//        alpha (forward) and beta (backward) recursions CONCURRENTLY over the
//        blank-interleaved lattice, P (blank,label) state pairs per thread,
//        neighbour states by warp shuffle (+ one smem word per warp boundary),
//        emissions staged into shared memory by TMA bulk copies
//        (cp.async.bulk + mbarrier, 4-stage ring per direction).  Log-space,
//        base 2.  Every thread keeps its states relative to its OWN integer
//        offset, re-centred every kRenorm frames without any block-wide
//        reduction, so the stored values stay within ~+-100 and fp32 keeps
//        ~1e-5 resolution exactly where the posteriors live (un-normalised
//        fp32 alphas, as in warp-ctc, lose 5e-4 at |log p| ~ 8000).
//   K3 ctc_grad              one warp per row: y = softmax(row) (second HBM
//        read), state posteriors gamma_t(s) from alpha, beta, E (normalised per
//        frame so the common-mode rounding of the two recursions cancels),
//        per-label sums through a label->states CSR built on the host, and ONE
//        coalesced float4 write of the gradient row.  Padded rows are zeroed.
//
// HBM traffic: 2 reads + 1 write of the [T,B,A] slab, + the compact
// per-utterance tables (E, alpha, beta: 5*(L+1)*4 B per frame).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <mutex>
#include <vector>

#include "../../include/b200ctc.h"

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr double kLn2 = 0.6931471805599453;
constexpr float kNeg = -1.0e30f;  // "log 0": absorbs every finite addend
constexpr int kStages = 4;        // TMA ring depth per direction
constexpr int kStageFloats = 2048;  // 8 KB per stage
constexpr int kRenorm = 8;        // frames between per-thread re-centrings
constexpr int kK1Warps = 8;
constexpr int kK3Warps = 8;

struct UttMeta {
  int T;          // input length
  int L;          // label length
  int lab_off;    // offset into flat labels
  int pitch;      // floats per frame of E (>= L+1, multiple of 4)
  int feasible;   // L + repeats <= T
  int csr_off;    // offset into uniq_lab / uniq_start (L+1 entries reserved +1)
  int nuniq;      // number of distinct labels
  int pad_;
  long long e_off;   // float offset of E_b
  long long ab_off;  // float offset of alpha_b / beta_b (2*pitch per frame)
  long long off_off; // float offset of this utterance's offset tables
  long long vrow0;   // valid rows (frames of feasible utterances) before this utterance
};

struct CtcDev {
  const float *act;
  float *grad;
  int A, B, Tmax, blank;
  int b_lo, nb;           // utterance range [b_lo, b_lo+nb) this launch works on
  float grad_scale;
  const UttMeta *meta;
  const int *labels;      // flat labels
  const int *uniq_lab;    // per utt: distinct labels (ascending)
  const int *uniq_start;  // per utt: nuniq+1 offsets into pos
  const int *pos;         // per utt: label positions grouped by label
  const int *nuniq;       // [B] number of distinct labels (uploaded with the CSR, after K1/K2 are queued)
  float *lse2;            // [Tmax*B] base-2 log-sum-exp of each row
  float *E;
  float *alpha;
  float *beta;            // shifted by one state: beta[s+1]
  float *offA, *offB;     // [utt][frame block][thread] integer offsets of alpha / beta
  int P;                  // state pairs per thread in K2
  double *logp2;          // [2*B]: alpha-side and beta-side log2 p(l|x)
  float *costs;           // [B]
  int *flags;             // [0]: non-finite cost seen
  int *argmax;            // optional [Tmax*B]: arg-max symbol per row (-1 on padded rows)
  long long *pc;          // phase counters (tuning builds with -DB200CTC_PHASE_COUNTERS only; NULL otherwise)
};

// In-kernel phase counters of the streaming kernels are a BUILD-time option: release kernels carry none.
#ifdef B200CTC_PHASE_COUNTERS
#define PC_DECL long long pc_t0 = clock64(), pc_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define PC_MARK(i)                    \
  do {                                \
    const long long pc_t1 = clock64(); \
    pc_acc[i] += pc_t1 - pc_t0;       \
    pc_t0 = pc_t1;                    \
  } while (0)
#define PC_FLUSH(cond, base)                                             \
  do {                                                                   \
    if (d.pc && (cond))                                                  \
      for (int pc_i = 0; pc_i < 8; pc_i++) atomicAdd((unsigned long long *)d.pc + (base) + pc_i, (unsigned long long)pc_acc[pc_i]); \
  } while (0)
#else
#define PC_DECL
#define PC_MARK(i)
#define PC_FLUSH(cond, base)
#endif

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// log2(2^a + 2^b)
__device__ __forceinline__ float lse2_2(float a, float b) {
  float m = fmaxf(a, b), d = fminf(a, b) - m;
  return m + lg2_approx(1.0f + ex2_approx(d));
}
// log2(2^a + 2^b + 2^c): the largest term is exactly 1 after the shift
__device__ __forceinline__ float lse2_3(float a, float b, float c) {
  float hi = fmaxf(a, b), lo = fminf(a, b);
  float m = fmaxf(hi, c);
  float mid = fmaxf(lo, fminf(hi, c));
  float mn = fminf(lo, c);
  return m + lg2_approx(1.0f + ex2_approx(mid - m) + ex2_approx(mn - m));
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier / TMA bulk helpers (PTX) -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// 1-D TMA: global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes,
                                            uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ===========================================================================
// K1: per-row log-sum-exp + gather of the lattice emissions
// ===========================================================================
__global__ void __launch_bounds__(kK1Warps * 32)
ctc_rowstats_gather_kernel(CtcDev d) {
  const int lane = threadIdx.x & 31;
  const long long lrow = (long long)blockIdx.x * kK1Warps + (threadIdx.x >> 5);
  if (lrow >= (long long)d.Tmax * d.nb) return;
  const int t = (int)(lrow / d.nb), b = d.b_lo + (int)(lrow - (long long)t * d.nb);
  const long long row = (long long)t * d.B + b;
  const UttMeta um = d.meta[b];
  if (t >= um.T) {
    if (d.argmax && lane == 0) d.argmax[row] = -1;
    return;
  }
  if (!um.feasible && !d.argmax) return;
  const int A = d.A;
  const float *a = d.act + row * A;

  float m = -3.0e38f, s = 0.f;
  int am = 0;  // index of the running maximum (first occurrence)
  if ((A & 3) == 0) {
    const float4 *a4 = reinterpret_cast<const float4 *>(a);
    const int n4 = A >> 2;
    for (int k = lane; k < n4; k += 256) {
      float4 v[8];  // 8 x 16 B per lane in flight
#pragma unroll
      for (int u = 0; u < 8; u++)
        v[u] = (k + 32 * u < n4) ? __ldg(a4 + k + 32 * u)
                                 : make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f);
#pragma unroll
      for (int u = 0; u < 8; u++) {
        float mx = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
        if (mx > m) {
          s *= ex2_approx((m - mx) * kLog2e);
          m = mx;
          const int k0 = 4 * (k + 32 * u);
          am = v[u].x == mx ? k0 : (v[u].y == mx ? k0 + 1 : (v[u].z == mx ? k0 + 2 : k0 + 3));
        }
        s += ex2_approx((v[u].x - m) * kLog2e) + ex2_approx((v[u].y - m) * kLog2e) +
             ex2_approx((v[u].z - m) * kLog2e) + ex2_approx((v[u].w - m) * kLog2e);
      }
    }
  } else {
    for (int k = lane; k < A; k += 32) {
      float v = __ldg(a + k);
      if (v > m) {
        s *= ex2_approx((m - v) * kLog2e);
        m = v;
        am = k;
      }
      s += ex2_approx((v - m) * kLog2e);
    }
  }
  const float M = warp_max(m);
  if (d.argmax) {  // smallest index among the lanes that hold the row maximum (FindRowMaxId's tie rule)
    int cand = m == M ? am : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, o));
    if (lane == 0) d.argmax[row] = cand;
    if (!um.feasible) return;
  }
  s *= ex2_approx((m - M) * kLog2e);
  const float S = warp_sum(s);
  const float l2 = M * kLog2e + log2f(S);
  if (lane == 0) d.lse2[row] = l2;

  // gather: E[t][0] = blank, E[t][1+i] = label i  (base-2 log-probabilities).  The row was just
  // streamed, so these are L1/L2 hits; four independent gathers per lane in flight.
  float *e = d.E + um.e_off + (long long)t * um.pitch;
  const int *lab = d.labels + um.lab_off;
  for (int u0 = lane; u0 <= um.L; u0 += 128) {
    int kk[4];
    float vv[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int u = u0 + 32 * i;
      kk[i] = u <= um.L ? (u == 0 ? d.blank : __ldg(lab + u - 1)) : d.blank;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) vv[i] = __ldg(a + kk[i]);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int u = u0 + 32 * i;
      if (u <= um.L) e[u] = vv[i] * kLog2e - l2;
    }
  }
}

// ===========================================================================
// K2: concurrent alpha / beta recursions, one CTA per utterance
// ===========================================================================
// Pair i (0..L) = {Y_i: a blank state, X_i: the label state after it (alpha) /
// before it (beta)}.  With the label string reversed, beta obeys the SAME
// recurrence as alpha, so both warp groups run this code:
//   Y_i <- Eb   + lse(Y_i, X_{i-1})
//   X_i <- El_i + lse(X_i, Y_i, skip_i ? X_{i-1} : 0)
// alpha: Y_i = state 2i, X_i = state 2i+1, time ascending, label i.
// beta : Y_i = state 2(L-i), X_i = state 2(L-i)-1, time descending, label L-1-i.
// The per-frame loop is ISSUE-bound as much as latency-bound (a first version spent ~225 SASS
// instructions per frame on index arithmetic: step / F, step % 8, role selects, 64-bit address
// rebuilds), so the role is a template parameter and every index advances incrementally.
template <int P, int ROLE>
__device__ __forceinline__ void ctc_ab_run(const CtcDev &d, const UttMeta &um, int b, int r, int nthreads_needed,
                                           int nbar, uint64_t *my_bar, float *my_stage, float2 *bnd, float *fin,
                                           int F) {
  const int lane = r & 31, w = r >> 5;
  const int L = um.L, T = um.T, pitch = um.pitch;
  const float *Eg = d.E + um.e_off;
  const int nchunks = (T + F - 1) / F;

  // chunk k (in visiting order) -> first frame and frame count
  auto chunk_lo = [&](int k) { return ROLE ? max(0, T - (k + 1) * F) : k * F; };
  auto chunk_n = [&](int k) { return min(F, T - k * F); };
  auto issue = [&](int k) {
    const int st = k % kStages;
    const uint32_t bytes = (uint32_t)chunk_n(k) * pitch * 4u;
    mbar_expect_tx(my_bar + st, bytes);
    tma_load_1d(my_stage + st * kStageFloats, Eg + (long long)chunk_lo(k) * pitch, bytes, my_bar + st);
  };
  if (r == 0)
    for (int k = 0; k < min(kStages, nchunks); k++) issue(k);

  // per-thread lattice slice
  const int i0 = r * P;
  const int *lab = d.labels + um.lab_off;
  float X[P], Y[P];
  int eidx[P];      // index of El_i inside a frame of E
  int soff[P];      // float offset of this pair inside a stored frame
  bool skip[P], hasX[P], hasY[P];
#pragma unroll
  for (int p = 0; p < P; p++) {
    const int i = i0 + p;
    hasY[p] = i <= L;
    hasX[p] = i < L;
    int li = 0, lprev = -1;
    if (hasX[p]) {
      li = ROLE ? lab[L - 1 - i] : lab[i];
      if (i >= 1) lprev = ROLE ? lab[L - i] : lab[i - 1];
    }
    skip[p] = hasX[p] && i >= 1 && li != lprev;
    eidx[p] = hasX[p] ? (ROLE ? L - i : 1 + i) : 0;
    soff[p] = hasY[p] ? (ROLE ? 2 * (L - i) : 2 * i) : 0;
    X[p] = kNeg;
    Y[p] = (i == 0) ? 0.f : kNeg;  // virtual frame "-1": all mass on the first blank
  }
  // Values are kept relative to a per-thread integer offset c (a float holding an
  // integer): true log2 value = stored + c.  A thread whose states are all still
  // unreachable simply adopts its neighbour's offset.
  float xin = kNeg;  // X_{i0-1} of the previous frame, already relative to c
  float c = 0.f;
  bool live = (r == 0);

  const int pitch2 = 2 * pitch;
  float *o = (ROLE ? d.beta : d.alpha) + um.ab_off + (ROLE ? (long long)(T - 1) * pitch2 : 0);
  float *off_out = (ROLE ? d.offB : d.offA) + um.off_off + r;
  const int o_step = ROLE ? -pitch2 : pitch2, e_step = ROLE ? -pitch : pitch;
  const bool writes_off = r < nthreads_needed;
  const int off_stride = (nthreads_needed + 3) & ~3;
  float2 *bn_w = bnd + w;          // this warp's slot; the double buffer toggles by +-32
  int par = 0, rn = kRenorm, st = 0;
  uint32_t ph = 0;

  for (int k = 0; k < nchunks; k++) {
    const int n = min(F, T - k * F);
    mbar_wait(my_bar + st, ph);
    const float *e = my_stage + st * kStageFloats + (ROLE ? (n - 1) * pitch : 0);
    const bool last_chunk = k == nchunks - 1;
    for (int f = 0; f < n; f++) {
      const float Eb = e[0];
      float El[P];
#pragma unroll
      for (int p = 0; p < P; p++) El[p] = hasX[p] ? e[eidx[p]] : kNeg;  // kNeg keeps a missing X at "log 0"
      e += e_step;

      float nX[P], nY[P];
#pragma unroll
      for (int p = 0; p < P; p++) {
        const float xp = p == 0 ? xin : X[p - 1];
        nY[p] = Eb + lse2_2(Y[p], xp);
        nX[p] = El[p] + lse2_3(X[p], Y[p], skip[p] ? xp : kNeg);
      }
#pragma unroll
      for (int p = 0; p < P; p++) {
        Y[p] = nY[p];
        X[p] = nX[p];
      }
      // store this frame (relative to c)
#pragma unroll
      for (int p = 0; p < P; p++)
        if (hasY[p])
          *reinterpret_cast<float2 *>(o + soff[p]) = ROLE ? make_float2(X[p], Y[p]) : make_float2(Y[p], X[p]);
      o += o_step;
      // end of a frame block: publish the offset the block was stored with, re-centre.  c only changes
      // here: re-centred if the thread has reachable states, otherwise adopted from the left neighbour.
      const bool block_end = (--rn == 0) || (last_chunk && f == n - 1);
      if (block_end) {
        rn = kRenorm;
        if (writes_off) *off_out = c;
        off_out += off_stride;
        float mx = kNeg;
#pragma unroll
        for (int p = 0; p < P; p++) mx = fmaxf(mx, fmaxf(hasX[p] ? X[p] : kNeg, hasY[p] ? Y[p] : kNeg));
        live = mx > -1.0e29f;
        if (live) {
          const float sh = floorf(mx);
#pragma unroll
          for (int p = 0; p < P; p++) {
            X[p] = fmaxf(X[p] - sh, kNeg);
            Y[p] = fmaxf(Y[p] - sh, kNeg);
          }
          c += sh;
        }
      }
      // hand (X_last, c) to the next thread for the next frame
      const float xs = __shfl_up_sync(0xffffffffu, X[P - 1], 1);
      const float cs = __shfl_up_sync(0xffffffffu, c, 1);
      if (lane == 31) bn_w[par] = make_float2(X[P - 1], c);
      named_bar_sync(1 + ROLE, nbar);
      float xv = xs, cv = cs;
      if (lane == 0) {
        const float2 v = w == 0 ? make_float2(kNeg, c) : bn_w[par - 1];
        xv = v.x;
        cv = v.y;
      }
      par ^= 32;
      if (block_end && !live) c = cv;    // nothing reachable here yet: follow the neighbour
      xin = fmaxf(xv + (cv - c), kNeg);  // cv - c is an exact integer
    }
    // stage fully consumed (every thread is past the barrier of its last frame) -> refill it
    if (r == 0 && k + kStages < nchunks) issue(k + kStages);
    if (++st == kStages) {
      st = 0;
      ph ^= 1;
    }
  }

  // log2 p(l|x) = lse(Y_L, X_{L-1}) in absolute terms
#pragma unroll
  for (int p = 0; p < P; p++) {
    const int i = i0 + p;
    if (i == L) {
      fin[0] = Y[p];
      fin[1] = c;
    }
    if (i == L - 1) {
      fin[2] = X[p];
      fin[3] = c;
    }
  }
  if (r == 0 && L == 0) {
    fin[2] = kNeg;
    fin[3] = 0.f;
  }
  named_bar_sync(1 + ROLE, nbar);
  if (r == 0) {
    const double v0 = (double)fin[0] + (double)fin[1];
    const double v1 = (double)fin[2] + (double)fin[3];
    const double hi = fmax(v0, v1), lo = fmin(v0, v1);
    const double lp2 = hi + log2(1.0 + exp2(fmax(lo - hi, -1000.0)));
    d.logp2[ROLE * d.B + b] = lp2;
    if (ROLE == 0) {
      const float cost = (float)(-lp2 * kLn2);
      d.costs[b] = cost;
      if (!(fabsf(cost) < 3.0e38f)) atomicOr(d.flags, 1);
    }
  }
}

template <int P>
__global__ void __launch_bounds__(1024, 1) ctc_alpha_beta_kernel(CtcDev d, int frames_per_stage) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw);               // [2][kStages]
  float2 *bnd = reinterpret_cast<float2 *>(smem_raw + 64);               // [2][2][32] (value, offset)
  float *fin = reinterpret_cast<float *>(smem_raw + 64 + 1024);          // [2][4]
  float *stages = reinterpret_cast<float *>(smem_raw + 2048);            // [2][kStages][kStageFloats]

  const int b = d.b_lo + blockIdx.x;
  const UttMeta um = d.meta[b];
  if (!um.feasible) return;
  const int NT = blockDim.x >> 1;
  const int role = threadIdx.x >= NT ? 1 : 0;
  const int r = threadIdx.x - role * NT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2 * kStages; i++) mbar_init(mbar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int nthreads_needed = (um.L + 1 + P - 1) / P;
  const int nwarps_active = (nthreads_needed + 31) >> 5;
  if ((r >> 5) >= nwarps_active) return;  // idle warps leave; named barriers count the rest
  const int nbar = nwarps_active * 32;
  if (role == 0)
    ctc_ab_run<P, 0>(d, um, b, r, nthreads_needed, nbar, mbar, stages, bnd, fin, frames_per_stage);
  else
    ctc_ab_run<P, 1>(d, um, b, r, nthreads_needed, nbar, mbar + kStages, stages + kStages * kStageFloats,
                     bnd + 64, fin + 4, frames_per_stage);
}

// ===========================================================================
// K3: gradient rows
// ===========================================================================
__global__ void __launch_bounds__(kK3Warps * 32) ctc_grad_kernel(CtcDev d, int smem_pitch) {
  extern __shared__ __align__(16) float gsm[];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const long long lrow = (long long)blockIdx.x * kK3Warps + wi;
  if (lrow >= (long long)d.Tmax * d.nb) return;
  const int t = (int)(lrow / d.nb), b = d.b_lo + (int)(lrow - (long long)t * d.nb);
  const long long row = (long long)t * d.B + b;
  const UttMeta um = d.meta[b];
  const int A = d.A;
  float *g = d.grad + row * A;
  const bool vec = (A & 3) == 0;

  if (t >= um.T || !um.feasible) {  // padded frame / unalignable utterance: zero row
    if (vec) {
      float4 *g4 = reinterpret_cast<float4 *>(g);
      for (int k = lane; k < (A >> 2); k += 32) g4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int k = lane; k < A; k += 32) g[k] = 0.f;
    }
    return;
  }
  const float *a = d.act + row * A;
  const float l2 = d.lse2[row];
  const float gs = d.grad_scale;
  const int L = um.L, S = 2 * L + 1, pitch = um.pitch;

  // state posteriors gamma_t(s) ~ 2^(alpha + beta - E + offsets - log2 p)
  const float *al = d.alpha + um.ab_off + (long long)t * 2 * pitch;
  const float *be = d.beta + um.ab_off + (long long)t * 2 * pitch + 1;
  const float *e = d.E + um.e_off + (long long)t * pitch;
  // offsets: alpha pair i = s/2 lives in thread i/P; beta pair i' = L - ceil(s/2)
  const int psh = d.P == 1 ? 0 : (d.P == 2 ? 1 : 2);             // P is 1, 2 or 4: shifts, not divisions
  const int nthr = (((L + d.P) >> psh) + 3) & ~3;                // row stride of the offset tables
  const float *oa = d.offA + um.off_off + (long long)(t / kRenorm) * nthr;
  const float *ob = d.offB + um.off_off + (long long)((um.T - 1 - t) / kRenorm) * nthr;
  const double lp2 = d.logp2[b];
  float *sm = gsm + wi * smem_pitch;  // gamma of the LABEL states only: sm[i] = gamma(2i+1)
  float z = 0.f, zblank = 0.f;
  // four independent states per lane in flight (the loads dominate this loop)
  for (int s0 = lane; s0 < S; s0 += 128) {
    float av[4], bv[4], ev[4], cv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int s = s0 + 32 * u;
      const bool ok = s < S;
      av[u] = ok ? al[s] : kNeg;
      bv[u] = ok ? be[s] : 0.f;
      ev[u] = ok ? ((s & 1) ? e[1 + (s >> 1)] : e[0]) : 0.f;
      cv[u] = ok ? oa[(s >> 1) >> psh] + ob[(L - ((s + 1) >> 1)) >> psh] : 0.f;  // exact: integers
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int s = s0 + 32 * u;
      const float D = (float)((double)cv[u] - lp2);
      const float v = ex2_approx(fminf(fmaxf((av[u] + bv[u] - ev[u]) + D, -200.f), 100.f));
      if (s < S) {
        z += v;
        if (s & 1) sm[s >> 1] = v;
        else zblank += v;
      }
    }
  }
  const float Z = warp_sum(z);
  const float zb = warp_sum(zblank);  // even states are blanks
  const float invZ = Z > 0.f ? 1.0f / Z : 0.f;
  if (lane == 0 && !(Z > 0.f && Z < 3.0e38f)) atomicOr(d.flags, 2);   // no usable posterior for this frame

  // y = softmax(row), streamed
  if (vec) {
    const float4 *a4 = reinterpret_cast<const float4 *>(a);
    float4 *g4 = reinterpret_cast<float4 *>(g);
    const int n4 = A >> 2;
    for (int k = lane; k < n4; k += 256) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (k + 32 * u < n4) v[u] = __ldg(a4 + k + 32 * u);
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (k + 32 * u < n4) {
          float4 y;
          y.x = gs * ex2_approx(v[u].x * kLog2e - l2);
          y.y = gs * ex2_approx(v[u].y * kLog2e - l2);
          y.z = gs * ex2_approx(v[u].z * kLog2e - l2);
          y.w = gs * ex2_approx(v[u].w * kLog2e - l2);
          g4[k + 32 * u] = y;
        }
    }
  } else {
    for (int k = lane; k < A; k += 32) g[k] = gs * ex2_approx(__ldg(a + k) * kLog2e - l2);
  }
  __syncwarp();  // orders the row stores above (and sm[]) before the per-label overwrites below
  const int *ul = d.uniq_lab + um.csr_off;
  const int *us = d.uniq_start + um.csr_off + b;  // nuniq+1 entries per utterance
  const int *pos = d.pos + um.lab_off;

  if (lane == 0) g[d.blank] = gs * (ex2_approx(__ldg(a + d.blank) * kLog2e - l2) - zb * invZ);
  // overwrite the entries of the labels that occur.  Per-label posterior mass from shared memory in a
  // fixed order (deterministic); the activation gathers are L2 hits, four independent ones in flight.
  const int nuniq = __ldg(d.nuniq + b);
  for (int j0 = lane; j0 < nuniq; j0 += 128) {
    int kk[4];
    float vv[4], mass[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int j = j0 + 32 * i;
      kk[i] = d.blank;
      mass[i] = 0.f;
      if (j < nuniq) {
        kk[i] = __ldg(ul + j);
        const int q1 = __ldg(us + j + 1);
        for (int q = __ldg(us + j); q < q1; q++) mass[i] += sm[__ldg(pos + q)];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) vv[i] = __ldg(a + kk[i]);
#pragma unroll
    for (int i = 0; i < 4; i++)
      if (j0 + 32 * i < nuniq) g[kk[i]] = gs * (ex2_approx(vv[i] * kLog2e - l2) - mass[i] * invZ);
  }
}

// ===========================================================================
// K3 (ring variant): gradient rows for wide alphabets, persistent, TMA-fed
// ===========================================================================
// The row-per-warp kernel above keeps only what its registers hold in flight, and every row starts with
// a serial chain of table loads (alpha, beta, E, offsets) during which nothing streams: at A = 4000 it
// reaches ~60 % of the copy bandwidth.  Here ONE persistent CTA per SM walks a contiguous range of valid
// rows.  Thread 0 keeps two rings of TMA bulk copies ahead of the CTA: the activation rows (NA slots) and
// the rows of the per-utterance tables (NT slots), so the memory system always has several rows per SM
// outstanding while all 8 warps work on the current one out of shared memory:
//   gamma (tables slot) | y = grad_scale*softmax in place (activation slot)      -- sync --
//   label masses through the CSR (kept in shared memory per utterance), subtracted in place  -- sync --
//   one TMA bulk store of the finished row (16 KB, full lines, no registers).
// Padded rows and rows of infeasible utterances are zero-filled at the end.
struct RingCfg {
  int NA, NT;            // ring depths
  int act_bytes;         // bytes per activation slot (A*4 rounded up to 128)
  int tab_bytes;         // bytes per table slot
  int pitch_max;
  int pshift;            // log2(P)
  long long vrow_base;   // meta[b_lo].vrow0
  long long vrows;       // valid rows of this group
};
struct RowCur {
  int b, t, T, L, pitch, nstr;
  long long e_off, ab_off, off_off;
};
__device__ __forceinline__ void cur_load(const CtcDev &d, RowCur &c, int b, int pshift) {
  const UttMeta um = d.meta[b];
  c.b = b;
  c.T = um.T;
  c.L = um.L;
  c.pitch = um.pitch;
  c.nstr = ((((um.L + (1 << pshift)) >> pshift)) + 3) & ~3;
  c.e_off = um.e_off;
  c.ab_off = um.ab_off;
  c.off_off = um.off_off;
}
__device__ __forceinline__ void cur_next(const CtcDev &d, RowCur &c, int b_end, int pshift) {
  if (++c.t < c.T) return;
  int b = c.b + 1;
  while (b < b_end && !d.meta[b].feasible) b++;
  c.t = 0;
  if (b < b_end) cur_load(d, c, b, pshift);
  else c.b = b_end;
}
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr int kRingConsumers = 256;                  // 8 consumer warps
constexpr int kRingThreads = kRingConsumers + 32;    // + the producer warp

__global__ void __launch_bounds__(kRingThreads, 2) ctc_grad_ring_kernel(CtcDev d, RingCfg rc) {
  extern __shared__ __align__(128) unsigned char rsm[];
  // barriers: full_act[NA] full_tab[NT] (TMA bytes) | done_act[NA] free_tab[NT] (one arrive per consumer warp)
  uint64_t *full_act = reinterpret_cast<uint64_t *>(rsm);
  uint64_t *full_tab = full_act + rc.NA;
  uint64_t *done_act = full_tab + rc.NT;
  uint64_t *free_tab = done_act + rc.NA;                           // 2*(NA+NT) <= 32 barriers = 256 B
  float *partials = reinterpret_cast<float *>(rsm + 256);          // [2 rows][2][8]
  unsigned char *act_base = rsm + 512;
  unsigned char *tab_base = act_base + (size_t)rc.NA * rc.act_bytes;
  float *gam0 = reinterpret_cast<float *>(tab_base + (size_t)rc.NT * rc.tab_bytes);  // [2 rows][pitch_max]
  int *s_us = reinterpret_cast<int *>(gam0 + 2 * rc.pitch_max);    // [pitch_max + 4]
  int *s_pos = s_us + rc.pitch_max + 4;                            // [pitch_max]
  int *s_ul = s_pos + rc.pitch_max;                                // [pitch_max]

  const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
  const int A = d.A, b_end = d.b_lo + d.nb, pshift = rc.pshift;
  const long long v_lo = rc.vrows * blockIdx.x / gridDim.x, v_hi = rc.vrows * (blockIdx.x + 1) / gridDim.x;
  const int nrows = (int)(v_hi - v_lo);
  const float gs = d.grad_scale;

  if (tid == 0) {
    for (int i = 0; i < rc.NA + rc.NT; i++) mbar_init(full_act + i, 1);
    for (int i = 0; i < rc.NA + rc.NT; i++) mbar_init(done_act + i, kRingConsumers / 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // first row of this CTA: the utterance that holds valid row v_lo
  RowCur cur;
  cur.b = b_end;
  if (nrows > 0) {
    int b = d.b_lo;
    for (; b < b_end; b++) {
      const UttMeta um = d.meta[b];
      if (um.feasible && v_lo < um.vrow0 - rc.vrow_base + um.T) break;
    }
    cur_load(d, cur, b, pshift);
    cur.t = (int)(v_lo - (d.meta[b].vrow0 - rc.vrow_base));
  }

  if (wi == kRingConsumers / 32) {
    // ===================== producer: one thread keeps both rings full and drains finished rows =====
    if (lane == 0 && nrows > 0) {
      RowCur pa = cur, pt = cur, ps = cur;  // next activation row / next table row to request / next row to store
      auto issue_act = [&](int i) {
        uint64_t *bar = full_act + (i % rc.NA);
        mbar_expect_tx(bar, (uint32_t)A * 4u);
        tma_load_1d(act_base + (size_t)(i % rc.NA) * rc.act_bytes, d.act + ((long long)pa.t * d.B + pa.b) * A,
                    (uint32_t)A * 4u, bar);
        cur_next(d, pa, b_end, pshift);
      };
      auto issue_tab = [&](int i) {
        uint64_t *bar = full_tab + (i % rc.NT);
        float *dst = reinterpret_cast<float *>(tab_base + (size_t)(i % rc.NT) * rc.tab_bytes);
        const int p = pt.pitch, ns = pt.nstr;
        mbar_expect_tx(bar, (uint32_t)(5 * p + 2 * ns) * 4u);
        tma_load_1d(dst, d.alpha + pt.ab_off + (long long)pt.t * 2 * p, (uint32_t)p * 8u, bar);
        tma_load_1d(dst + 2 * p, d.beta + pt.ab_off + (long long)pt.t * 2 * p, (uint32_t)p * 8u, bar);
        tma_load_1d(dst + 4 * p, d.E + pt.e_off + (long long)pt.t * p, (uint32_t)p * 4u, bar);
        tma_load_1d(dst + 5 * p, d.offA + pt.off_off + (long long)(pt.t / kRenorm) * ns, (uint32_t)ns * 4u, bar);
        tma_load_1d(dst + 5 * p + ns, d.offB + pt.off_off + (long long)((pt.T - 1 - pt.t) / kRenorm) * ns,
                    (uint32_t)ns * 4u, bar);
        cur_next(d, pt, b_end, pshift);
      };
      for (int i = 0; i < min(rc.NT, nrows); i++) issue_tab(i);
      for (int i = 0; i < min(rc.NA - 1, nrows); i++) issue_act(i);
      for (int i = 0; i < nrows; i++) {
        if (i + rc.NT < nrows) {  // the consumers are done with row i's tables -> that slot takes row i+NT
          mbar_wait(free_tab + (i % rc.NT), (uint32_t)(i / rc.NT) & 1u);
          issue_tab(i + rc.NT);
        }
        mbar_wait(done_act + (i % rc.NA), (uint32_t)(i / rc.NA) & 1u);  // row i finished in its slot
        bulk_store(d.grad + ((long long)ps.t * d.B + ps.b) * A, act_base + (size_t)(i % rc.NA) * rc.act_bytes,
                   (uint32_t)A * 4u);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        cur_next(d, ps, b_end, pshift);
        // the slot of row i-1 (its store was committed one row ago) takes row i+NA-1
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        if (i + rc.NA - 1 < nrows) issue_act(i + rc.NA - 1);
      }
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  } else {
    // ===================== consumers: 8 warps on one row at a time ================================
    int loaded_b = -1, nuniq = 0;
    float lp_hi = 0.f, lp_lo = 0.f;
    for (int i = 0; i < nrows; i++) {
      if (cur.b != loaded_b) {  // new utterance: its label -> positions CSR into shared memory
        named_bar_sync(1, kRingConsumers);
        const UttMeta um = d.meta[cur.b];
        nuniq = d.nuniq[cur.b];
        const int *ul = d.uniq_lab + um.csr_off, *us = d.uniq_start + um.csr_off + cur.b, *pos = d.pos + um.lab_off;
        for (int k = tid; k <= nuniq; k += kRingConsumers) s_us[k] = us[k];
        for (int k = tid; k < nuniq; k += kRingConsumers) s_ul[k] = ul[k];
        for (int k = tid; k < um.L; k += kRingConsumers) s_pos[k] = pos[k];
        // log2 p(l|x) split into an integer and a fraction: (offset sum - integer) is exact in fp32
        const double lp2 = d.logp2[cur.b];
        const double fl = floor(lp2);
        lp_hi = (float)fl;
        lp_lo = (float)(lp2 - fl);
        loaded_b = cur.b;
        named_bar_sync(1, kRingConsumers);
      }
      const long long row = (long long)cur.t * d.B + cur.b;
      const float l2 = __ldg(d.lse2 + row);  // needed only after the gamma phase
      const int L = cur.L, S = 2 * L + 1, p = cur.pitch, ns = cur.nstr;
      float *gam = gam0 + (i & 1) * rc.pitch_max;
      float *part = partials + (i & 1) * 16;

      // ---- gamma_t(s) ~ 2^(alpha + beta - E + offsets - log2 p), label states to gam[], sums for Z
      const float *tab = reinterpret_cast<const float *>(tab_base + (size_t)(i % rc.NT) * rc.tab_bytes);
      const float *al = tab, *be = tab + 2 * p + 1, *e = tab + 4 * p, *oa = tab + 5 * p, *ob = oa + ns;
      mbar_wait(full_tab + (i % rc.NT), (uint32_t)(i / rc.NT) & 1u);
      float z = 0.f, zblank = 0.f;
      const float eb = e[0];
      {
        // one (blank, label) pair per thread and iteration: states 2i and 2i+1
        for (int i2 = tid; i2 <= L; i2 += kRingConsumers) {
          const float2 a2 = *reinterpret_cast<const float2 *>(al + 2 * i2);
          const float b0 = be[2 * i2], b1 = be[2 * i2 + 1];
          const float o_a = oa[i2 >> pshift];
          const float c0 = o_a + ob[(L - i2) >> pshift];          // state 2i   (exact: integers)
          const float v0 = ex2_approx(fminf(fmaxf((a2.x + b0 - eb) + ((c0 - lp_hi) - lp_lo), -200.f), 100.f));
          z += v0;
          zblank += v0;
          if (i2 < L) {
            const float c1 = o_a + ob[(L - i2 - 1) >> pshift];    // state 2i+1
            const float v1 =
                ex2_approx(fminf(fmaxf((a2.y + b1 - e[1 + i2]) + ((c1 - lp_hi) - lp_lo), -200.f), 100.f));
            z += v1;
            gam[i2] = v1;
          }
        }
      }
      z = warp_sum(z);
      zblank = warp_sum(zblank);
      if (lane == 0) {
        part[wi] = z;
        part[8 + wi] = zblank;
        mbar_arrive(free_tab + (i % rc.NT));  // (shuffles above: every lane of the warp is past its table reads)
      }

      // ---- y = grad_scale * softmax(row), in place in the activation slot
      float4 *a4 = reinterpret_cast<float4 *>(act_base + (size_t)(i % rc.NA) * rc.act_bytes);
      mbar_wait(full_act + (i % rc.NA), (uint32_t)(i / rc.NA) & 1u);
      const int n4 = A >> 2;
      for (int k0 = tid; k0 < n4; k0 += 4 * kRingConsumers) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (k0 + u * kRingConsumers < n4) v[u] = a4[k0 + u * kRingConsumers];
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (k0 + u * kRingConsumers < n4) {
            v[u].x = gs * ex2_approx(v[u].x * kLog2e - l2);
            v[u].y = gs * ex2_approx(v[u].y * kLog2e - l2);
            v[u].z = gs * ex2_approx(v[u].z * kLog2e - l2);
            v[u].w = gs * ex2_approx(v[u].w * kLog2e - l2);
            a4[k0 + u * kRingConsumers] = v[u];
          }
      }
      named_bar_sync(1, kRingConsumers);  // gam[], part[] and the scaled row are complete

      // ---- subtract the posterior mass of every distinct label (fixed summation order) and of the blank
      float Z = 0.f, zb = 0.f;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        Z += part[k];
        zb += part[8 + k];
      }
      const float invZ = Z > 0.f ? 1.0f / Z : 0.f;
      if (tid == 0 && !(Z > 0.f && Z < 3.0e38f)) atomicOr(d.flags, 2);   // no usable posterior for this frame
      float *arow = reinterpret_cast<float *>(a4);
      {
        for (int j = tid; j < nuniq; j += kRingConsumers) {
          const int q0 = s_us[j], q1 = s_us[j + 1];
          float acc = gam[s_pos[q0]];
          for (int q = q0 + 1; q < q1; q++) acc += gam[s_pos[q]];
          arow[s_ul[j]] -= gs * (acc * invZ);
        }
      }
      if (tid == 0) arow[d.blank] -= gs * (zb * invZ);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk store
      __syncwarp();
      if (lane == 0) mbar_arrive(done_act + (i % rc.NA));
      cur_next(d, cur, b_end, pshift);
    }
  }

  // ---- zero rows: padded frames and infeasible utterances of this group
  if (wi < kRingConsumers / 32) {
    const long long rows = (long long)d.Tmax * d.nb;
    for (long long lrow = (long long)blockIdx.x * 8 + wi; lrow < rows; lrow += (long long)gridDim.x * 8) {
      const int t = (int)(lrow / d.nb), b = d.b_lo + (int)(lrow - (long long)t * d.nb);
      const int Tb = d.meta[b].T, feas = d.meta[b].feasible;
      if (t < Tb && feas) continue;
      float4 *g4 = reinterpret_cast<float4 *>(d.grad + ((long long)t * d.B + b) * A);
      for (int k = lane; k < (A >> 2); k += 32) g4[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// ===========================================================================
// Streaming path (wide alphabets, many utterances): ONE persistent CTA per utterance, two kernels
// ===========================================================================
// The three-kernel path above reads the slab twice, writes it once and moves alpha, beta and E through HBM
// (38 GB for 21.6 GB algorithmic on the A = 4000, B = 256 stress case), with the alpha/beta recursion as a
// separate, latency-bound phase.  Here the recursions RIDE the two slab reads, warp-specialised so that the
// per-frame dependent chain is only the recursion itself:
//   SA ctc_stream_alpha   forward in time.  Producer thread: ring of TMA bulk copies of the utterance's rows.
//        4 "stats" warps take whole rows round-robin (log-sum-exp + arg-max out of shared memory, several frames
//        ahead).  4 "chain" warps gather their states' emissions from the row in shared memory and advance alpha
//        (same per-thread integer offsets and re-centring as K2), one named barrier per frame among themselves.
//        Written to HBM: alpha_t (8 B per state pair) with the row's log-sum-exp in the unused last slot.
//   SB ctc_stream_beta_grad   backward in time.  Row ring + a ring of alpha rows.  4 chain warps: gather, beta
//        step, gamma = alpha * beta / (y p) straight from registers into a small ring in shared memory.  4 "row"
//        warps, a few frames behind: grad_scale * softmax IN PLACE in the row slot, per-label masses through the
//        CSR subtracted in place, then ONE TMA bulk store of the row.  beta and E never exist in HBM.  Padded rows
//        are zero-filled by the CTA when its utterance is done, which keeps the memory system busy while the longest
//        utterances finish.
// HBM traffic: 2 reads + 1 write of the slab + alpha written and read once (8 B per pair and frame).
constexpr int kSChain = 128;         // chain threads (warps 0-3)
constexpr int kSC = 256;             // consumer threads (chain + stats / row warps)
constexpr int kSThreads = kSC + 32;  // + the producer warp
constexpr int kNG = 2;               // gamma ring depth (SB): how far the chain may run ahead of the row warps

// Row ring of SA: a multiple of the number of stats warps, so that a stats warp always works on the SAME slots and
// therefore sees every phase of their mbarriers in order (a parity wait is only sound for a waiter that has
// observed the previous phase).
constexpr int kNAs = 4;

struct StreamCfg {
  int NA, NT;        // ring depths of SB: rows / alpha-table rows
  int act_bytes;     // bytes per row slot (A*4 rounded up to 128)
  int tab_bytes;     // bytes per table slot
  int pitch_max;
  const int *order;  // CTA -> utterance, longest first
};

// state pairs per chain thread: 1, 2, 4 or 8 (L + 1 <= 128 * P)
__device__ __host__ __forceinline__ int stream_pshift(int L) {
  return L + 1 <= kSChain ? 0 : (L + 1 <= 2 * kSChain ? 1 : (L + 1 <= 4 * kSChain ? 2 : 3));
}

// The recursion state of one chain thread (P pairs).  ROLE 0: alpha (labels as given), ROLE 1: beta (the same
// recurrence on the reversed label string, pair i = {Y: state 2(L-i), X: state 2(L-i)-1}).
template <int P, int ROLE>
struct Chain {
  float X[P], Y[P];
  int lk[P];
  bool skip[P], hasX[P], hasY[P];
  float xin, c, xs_prev, cs_prev;
  bool live, blockend_prev;
  int rn;
  __device__ __forceinline__ void init(const int *lab, int L, int tid, int blank) {
#pragma unroll
    for (int p = 0; p < P; p++) {
      const int i = tid * P + p;
      hasY[p] = i <= L;
      hasX[p] = i < L;
      const int li = hasX[p] ? (ROLE ? lab[L - 1 - i] : lab[i]) : blank;
      const int lprev = (hasX[p] && i >= 1) ? (ROLE ? lab[L - i] : lab[i - 1]) : -1;
      skip[p] = hasX[p] && i >= 1 && li != lprev;
      lk[p] = li;
      X[p] = kNeg;
      Y[p] = (i == 0) ? 0.f : kNeg;  // virtual frame "-1": all mass on the first blank
    }
    xin = kNeg;
    c = 0.f;
    xs_prev = kNeg;
    cs_prev = 0.f;
    live = (tid == 0);
    blockend_prev = false;
    rn = kRenorm;
  }
  // what the left neighbour handed over at the end of the previous frame (lane 0 reads the warp-edge word)
  __device__ __forceinline__ void take_handover(const float2 *bnd_prev, int lane, int wi) {
    float xv = xs_prev, cv = cs_prev;
    if (lane == 0) {
      const float2 v = wi == 0 ? make_float2(kNeg, c) : bnd_prev[wi - 1];
      xv = v.x;
      cv = v.y;
    }
    if (blockend_prev && !live) c = cv;  // nothing reachable here yet: follow the neighbour's offset
    xin = fmaxf(xv + (cv - c), kNeg);    // cv - c is an exact integer
  }
  __device__ __forceinline__ void step(float Eb, const float (&El)[P]) {
    float nX[P], nY[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
      const float xp = p == 0 ? xin : X[p - 1];
      nY[p] = Eb + lse2_2(Y[p], xp);
      nX[p] = El[p] + lse2_3(X[p], Y[p], skip[p] ? xp : kNeg);
    }
#pragma unroll
    for (int p = 0; p < P; p++) {
      Y[p] = nY[p];
      X[p] = nX[p];
    }
  }
  // end of a frame block: re-centre the stored values around this thread's own integer offset
  __device__ __forceinline__ void recentre() {
    rn = kRenorm;
    float mx = kNeg;
#pragma unroll
    for (int p = 0; p < P; p++) mx = fmaxf(mx, fmaxf(hasX[p] ? X[p] : kNeg, hasY[p] ? Y[p] : kNeg));
    live = mx > -1.0e29f;
    if (live) {
      const float sh = floorf(mx);
#pragma unroll
      for (int p = 0; p < P; p++) {
        X[p] = fmaxf(X[p] - sh, kNeg);
        Y[p] = fmaxf(Y[p] - sh, kNeg);
      }
      c += sh;
    }
  }
  __device__ __forceinline__ void hand_over(bool block_end, float2 *bnd_cur, int lane, int wi) {
    xs_prev = __shfl_up_sync(0xffffffffu, X[P - 1], 1);
    cs_prev = __shfl_up_sync(0xffffffffu, c, 1);
    if (lane == 31) bnd_cur[wi] = make_float2(X[P - 1], c);
    blockend_prev = block_end;
  }
};

// ---- SA ----------------------------------------------------------------------------------------------
struct SaSmem {
  uint64_t *full, *freeb, *lready;   // [NA] each
  float *l2ring;                     // [NA]
  float2 *bnd;                       // [2][4]
  float *fin;                        // [4]
  unsigned char *ring;
};

template <int P>
__device__ __forceinline__ void sa_chain(const CtcDev &d, const UttMeta &um, int b, int tid, const StreamCfg &sc,
                                         const SaSmem &sm) {
  const int lane = tid & 31, wi = tid >> 5;
  const int L = um.L, T = um.T;
  const int nthreads_needed = (L + 1 + P - 1) / P;
  Chain<P, 0> ch;
  ch.init(d.labels + um.lab_off, L, tid, d.blank);
  const int pitch2 = 2 * um.pitch;
  float *o = d.alpha + um.ab_off + 2 * tid * P;
  const int off_stride = (nthreads_needed + 3) & ~3;
  float *off_out = d.offA + um.off_off + tid;
  const bool writes_off = tid < nthreads_needed;
  const bool store_alpha = d.grad != nullptr;
  PC_DECL;
  for (int t = 0; t < T; t++) {
    const int slot = t % kNAs, par = t & 1;
    const uint32_t ph = (uint32_t)(t / kNAs) & 1u;
    PC_MARK(7);
    mbar_wait(sm.lready + slot, ph);   // the row's statistics exist (hence the row has landed)
    mbar_wait(sm.full + slot, ph);
    PC_MARK(0);
    const float *row = reinterpret_cast<const float *>(sm.ring + (size_t)slot * sc.act_bytes);
    const float l2 = sm.l2ring[slot];
    if (t > 0) ch.take_handover(sm.bnd + (par ^ 1) * 4, lane, wi);
    const float Eb = row[d.blank] * kLog2e - l2;
    float El[P];
#pragma unroll
    for (int p = 0; p < P; p++) El[p] = ch.hasX[p] ? row[ch.lk[p]] * kLog2e - l2 : kNeg;
    __syncwarp();
    if (lane == 0) mbar_arrive(sm.freeb + slot);  // this warp is done with the row slot
    PC_MARK(1);
    ch.step(Eb, El);
    PC_MARK(2);
    if (store_alpha) {
#pragma unroll
      for (int p = 0; p < P; p++)   // the slot after the last blank is free: it carries the row's log-sum-exp
        if (ch.hasY[p]) *reinterpret_cast<float2 *>(o + 2 * p) = make_float2(ch.Y[p], ch.hasX[p] ? ch.X[p] : l2);
      o += pitch2;
    }
    const bool block_end = (--ch.rn == 0) || (t == T - 1);
    if (block_end) {
      if (writes_off && store_alpha) *off_out = ch.c;   // the offset this frame block was stored with
      off_out += off_stride;
      ch.recentre();
    }
    ch.hand_over(block_end, sm.bnd + par * 4, lane, wi);
    PC_MARK(3);
    named_bar_sync(1, kSChain);   // the warp-edge words of frame t are visible to the neighbours
    PC_MARK(4);
  }
  PC_FLUSH(tid == 0 && blockIdx.x == 0, 0);
#pragma unroll
  for (int p = 0; p < P; p++) {
    const int i = tid * P + p;
    if (i == L) {
      sm.fin[0] = ch.Y[p];
      sm.fin[1] = ch.c;
    }
    if (i == L - 1) {
      sm.fin[2] = ch.X[p];
      sm.fin[3] = ch.c;
    }
  }
  if (tid == 0 && L == 0) {
    sm.fin[2] = kNeg;
    sm.fin[3] = 0.f;
  }
  named_bar_sync(1, kSChain);
  if (tid == 0) {
    const double v0 = (double)sm.fin[0] + (double)sm.fin[1], v1 = (double)sm.fin[2] + (double)sm.fin[3];
    const double hi = fmax(v0, v1), lo = fmin(v0, v1);
    const double lp2 = hi + log2(1.0 + exp2(fmax(lo - hi, -1000.0)));
    d.logp2[b] = lp2;
    const float cost = (float)(-lp2 * kLn2);
    d.costs[b] = cost;
    if (!(fabsf(cost) < 3.0e38f)) atomicOr(d.flags, 1);
  }
}

__global__ void __launch_bounds__(kSThreads, 2) ctc_stream_alpha_kernel(CtcDev d, StreamCfg sc) {
  extern __shared__ __align__(128) unsigned char ssm[];
  SaSmem sm;
  sm.full = reinterpret_cast<uint64_t *>(ssm);
  sm.freeb = sm.full + 8;
  sm.lready = sm.freeb + 8;                                   // 3 x 8 barriers = 192 B
  sm.l2ring = reinterpret_cast<float *>(ssm + 192);            // [8]
  sm.bnd = reinterpret_cast<float2 *>(ssm + 256);              // [2][4]
  sm.fin = reinterpret_cast<float *>(ssm + 320);               // [4]
  sm.ring = ssm + 512;

  const int b = sc.order[blockIdx.x];
  const UttMeta um = d.meta[b];
  const bool feasible = um.feasible != 0;
  if (!feasible && !d.argmax) return;
  const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
  const int T = um.T, A = d.A;
  if (tid == 0) {
    for (int i = 0; i < kNAs; i++) {
      mbar_init(sm.full + i, 1);
      mbar_init(sm.freeb + i, 1 + (feasible ? kSChain / 32 : 0));   // the row's stats warp + the chain warps
      mbar_init(sm.lready + i, 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (wi == kSC / 32) {
    if (lane == 0) {  // producer: keeps NA rows of this utterance in flight
      const uint32_t bytes = (uint32_t)A * 4u;
      auto issue = [&](int t) {
        uint64_t *bar = sm.full + t % kNAs;
        mbar_expect_tx(bar, bytes);
        tma_load_1d(sm.ring + (size_t)(t % kNAs) * sc.act_bytes, d.act + ((long long)t * d.B + b) * A, bytes, bar);
      };
      for (int t = 0; t < min(kNAs, T); t++) issue(t);
      for (int t = 0; t + kNAs < T; t++) {
        mbar_wait(sm.freeb + t % kNAs, (uint32_t)(t / kNAs) & 1u);
        issue(t + kNAs);
      }
    }
    return;
  }
  if (wi >= kSChain / 32) {
    // ===================== stats warps: whole rows, round-robin, ahead of the chain =====================
    const int sw = wi - kSChain / 32, nsw = (kSC - kSChain) / 32;
    static_assert(kNAs % ((kSC - kSChain) / 32) == 0, "each stats warp must own a fixed set of ring slots");
    const int n4 = A >> 2;
    PC_DECL;
    for (int t = sw; t < T; t += nsw) {
      const int slot = t % kNAs;
      PC_MARK(7);
      mbar_wait(sm.full + slot, (uint32_t)(t / kNAs) & 1u);
      PC_MARK(0);
      const float4 *r4 = reinterpret_cast<const float4 *>(sm.ring + (size_t)slot * sc.act_bytes);
      // (one warp per row is latency-bound if it runs the online max/rescale recurrence: 3800 cycles per row
      // measured.  Two straight passes over shared memory -- max, then sum of 2^(x - max) with independent
      // accumulators and 8 loads in flight -- are throughput-shaped.)
      float m = -3.0e38f;
      int am = 0;
      for (int k0 = lane; k0 < n4; k0 += 256) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
          v[u] = (k0 + 32 * u < n4) ? r4[k0 + 32 * u] : make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f);
        if (d.argmax) {
#pragma unroll
          for (int u = 0; u < 8; u++) {   // ascending indices: the first occurrence of a new maximum wins
            const float mx = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
            if (mx > m) {
              m = mx;
              const int kk = 4 * (k0 + 32 * u);
              am = v[u].x == mx ? kk : (v[u].y == mx ? kk + 1 : (v[u].z == mx ? kk + 2 : kk + 3));
            }
          }
        } else {
          float mx[8];
#pragma unroll
          for (int u = 0; u < 8; u++) mx[u] = fmaxf(fmaxf(v[u].x, v[u].y), fmaxf(v[u].z, v[u].w));
          m = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])), m));
        }
      }
      const float M = warp_max(m);
      if (d.argmax) {  // smallest index among the lanes that hold the row maximum (FindRowMaxId's tie rule)
        int cand = m == M ? am : 0x7fffffff;
#pragma unroll
        for (int q = 16; q > 0; q >>= 1) cand = min(cand, __shfl_xor_sync(0xffffffffu, cand, q));
        if (lane == 0) d.argmax[(long long)t * d.B + b] = cand;
      }
      const float Ml = M * kLog2e;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int k0 = lane; k0 < n4; k0 += 256) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; u++)
          v[u] = (k0 + 32 * u < n4) ? r4[k0 + 32 * u] : make_float4(-3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f);
#pragma unroll
        for (int u = 0; u < 8; u++) {
          acc[0] += ex2_approx(fmaf(v[u].x, kLog2e, -Ml));
          acc[1] += ex2_approx(fmaf(v[u].y, kLog2e, -Ml));
          acc[2] += ex2_approx(fmaf(v[u].z, kLog2e, -Ml));
          acc[3] += ex2_approx(fmaf(v[u].w, kLog2e, -Ml));
        }
      }
      const float S = warp_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
      if (lane == 0) {
        sm.l2ring[slot] = Ml + log2f(S);
        mbar_arrive(sm.lready + slot);
        mbar_arrive(sm.freeb + slot);
      }
      PC_MARK(1);
    }
    PC_FLUSH(sw == 0 && lane == 0 && blockIdx.x == 0, 8);
    if (d.argmax)  // padded frames of this utterance
      for (int t = T + (tid - kSChain); t < d.Tmax; t += kSC - kSChain) d.argmax[(long long)t * d.B + b] = -1;
    return;
  }
  if (!feasible) return;   // (arg-max only: the stats warps do all the work)
  switch (stream_pshift(um.L)) {
    case 0: sa_chain<1>(d, um, b, tid, sc, sm); break;
    case 1: sa_chain<2>(d, um, b, tid, sc, sm); break;
    case 2: sa_chain<4>(d, um, b, tid, sc, sm); break;
    default: sa_chain<8>(d, um, b, tid, sc, sm); break;
  }
}

// ---- SB ----------------------------------------------------------------------------------------------
struct SbSmem {
  uint64_t *full_act, *full_tab, *done_act, *free_tab, *gready, *gfree;
  float *part;      // [kNG][12]: 4 x z, 4 x z_blank, l2
  float2 *bnd;      // [2][4]
  unsigned char *act_base, *tab_base;
  float *gam;       // [kNG][pitch_max]
  int *s_us, *s_pos, *s_ul;
};

template <int P>
__device__ __forceinline__ void sb_chain(const CtcDev &d, const UttMeta &um, int b, int tid, const StreamCfg &sc,
                                         const SbSmem &sm) {
  constexpr int psh = P == 1 ? 0 : (P == 2 ? 1 : (P == 4 ? 2 : 3));
  const int lane = tid & 31, wi = tid >> 5;
  const int L = um.L, T = um.T, pitch = um.pitch;
  Chain<P, 1> ch;
  ch.init(d.labels + um.lab_off, L, tid, d.blank);
  const double lp2 = d.logp2[b];
  const double fl = floor(lp2);
  const float lp_hi = (float)fl, lp_lo = (float)(lp2 - fl);  // (offset sum - integer part) is exact in fp32
  // The alpha-table ring is fed by chain thread 0 itself, right after the frame's named barrier (every chain warp is
  // then past its reads of the slot): a producer loop that also waits for the row warps would hand out table rows at
  // THEIR pace and starve the chain, which runs up to kNG frames ahead (measured: 830 cycles per frame waiting).
  const int nstr = ((((L + (1 << psh)) >> psh)) + 3) & ~3;
  auto issue_tab = [&](int k, int slot) {
    const int t = T - 1 - k;
    uint64_t *bar = sm.full_tab + slot;
    float *dst = reinterpret_cast<float *>(sm.tab_base + (size_t)slot * sc.tab_bytes);
    mbar_expect_tx(bar, (uint32_t)(2 * pitch + nstr) * 4u);
    tma_load_1d(dst, d.alpha + um.ab_off + (long long)t * 2 * pitch, (uint32_t)pitch * 8u, bar);
    tma_load_1d(dst + 2 * pitch, d.offA + um.off_off + (long long)(t / kRenorm) * nstr, (uint32_t)nstr * 4u, bar);
  };
  if (tid == 0)
    for (int k = 0; k < min(sc.NT, T); k++) issue_tab(k, k);
  // ring positions advance incrementally (NA, NT are run-time values: k % NA costs ~40 instructions a time)
  int slotA = 0, slotT = 0, slotG = 0;
  uint32_t phA = 0, phT = 0, phG = 0;
  PC_DECL;
  for (int k = 0; k < T; k++) {
    const int par = k & 1;
    const float *arow = reinterpret_cast<const float *>(sm.act_base + (size_t)slotA * sc.act_bytes);
    const float *tab = reinterpret_cast<const float *>(sm.tab_base + (size_t)slotT * sc.tab_bytes);
    const float *oa = tab + 2 * pitch;   // alpha's per-thread offsets of this frame block
    float *gam = sm.gam + slotG * sc.pitch_max;
    PC_MARK(7);
    if (k >= kNG) mbar_wait(sm.gfree + slotG, phG ^ 1u);   // the row warps released this gamma slot
    PC_MARK(0);
    mbar_wait(sm.full_tab + slotT, phT);
    PC_MARK(1);
    mbar_wait(sm.full_act + slotA, phA);
    PC_MARK(2);
    const float l2 = tab[2 * L + 1];     // the row's base-2 log-sum-exp, left there by SA
    if (k > 0) ch.take_handover(sm.bnd + (par ^ 1) * 4, lane, wi);
    const float Eb = arow[d.blank] * kLog2e - l2;
    float El[P];
#pragma unroll
    for (int p = 0; p < P; p++) El[p] = ch.hasX[p] ? arow[ch.lk[p]] * kLog2e - l2 : kNeg;
    // alpha of my states and the offsets they were stored with: loaded before the recursion step, used after it
    float aY[P], aX[P], oY[P], oX[P];
#pragma unroll
    for (int p = 0; p < P; p++) {
      const int i = tid * P + p;
      const int sY = ch.hasY[p] ? 2 * (L - i) : 0;
      aY[p] = tab[sY];
      oY[p] = oa[ch.hasY[p] ? (L - i) >> psh : 0];
      aX[p] = ch.hasX[p] ? tab[sY - 1] : 0.f;
      oX[p] = ch.hasX[p] ? oa[(L - i - 1) >> psh] : 0.f;
    }
    ch.step(Eb, El);
    PC_MARK(3);
    // state posteriors gamma_t(s) ~ 2^(alpha + beta - E + offsets - log2 p), straight from the registers
    float z = 0.f, zblank = 0.f;
    const float cb = ch.c - lp_hi;
#pragma unroll
    for (int p = 0; p < P; p++) {
      const int i = tid * P + p;
      if (ch.hasY[p]) {
        const float v0 = ex2_approx(fminf(fmaxf((aY[p] + ch.Y[p] - Eb) + ((oY[p] + cb) - lp_lo), -200.f), 100.f));
        zblank += v0;
        if (ch.hasX[p]) {
          const float v1 = ex2_approx(fminf(fmaxf((aX[p] + ch.X[p] - El[p]) + ((oX[p] + cb) - lp_lo), -200.f), 100.f));
          z += v1;
          gam[L - 1 - i] = v1;   // label position L-1-i
        }
      }
    }
    z += zblank;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {   // the two sums share the shuffle rounds
      z += __shfl_xor_sync(0xffffffffu, z, o);
      zblank += __shfl_xor_sync(0xffffffffu, zblank, o);
    }
    if (lane == 0) {
      sm.part[slotG * 12 + wi] = z;
      sm.part[slotG * 12 + 4 + wi] = zblank;
      if (wi == 0) sm.part[slotG * 12 + 8] = l2;
      mbar_arrive(sm.gready + slotG);     // release: this warp's gam[] / part[] writes and its gathers from the raw row
    }
    PC_MARK(4);
    const bool block_end = (--ch.rn == 0) || (k == T - 1);
    if (block_end) ch.recentre();
    ch.hand_over(block_end, sm.bnd + par * 4, lane, wi);
    named_bar_sync(1, kSChain);
    if (tid == 0 && k + sc.NT < T) issue_tab(k + sc.NT, slotT);   // every chain warp is past this table slot
    PC_MARK(5);
    if (++slotA == sc.NA) { slotA = 0; phA ^= 1u; }
    if (++slotT == sc.NT) { slotT = 0; phT ^= 1u; }
    if (++slotG == kNG) { slotG = 0; phG ^= 1u; }
  }
  PC_FLUSH(tid == 0 && blockIdx.x == 0, 16);
}

__device__ __forceinline__ void sb_rows(const CtcDev &d, const UttMeta &um, int rt, const StreamCfg &sc,
                                        const SbSmem &sm, int nuniq) {
  constexpr int NR = kSC - kSChain;   // row threads
  const int lane = rt & 31;
  const int T = um.T, n4 = d.A >> 2;
  const float gs = d.grad_scale;
  int slotA = 0, slotG = 0;
  uint32_t phA = 0, phG = 0;
  PC_DECL;
  for (int k = 0; k < T; k++) {
    float *arow = reinterpret_cast<float *>(sm.act_base + (size_t)slotA * sc.act_bytes);
    const float *gam = sm.gam + slotG * sc.pitch_max;
    PC_MARK(7);
    mbar_wait(sm.gready + slotG, phG);      // gamma of frame k is complete, the raw row is no longer needed
    mbar_wait(sm.full_act + slotA, phA);
    PC_MARK(0);
    const float *part = sm.part + slotG * 12;
    const float Z = (part[0] + part[1]) + (part[2] + part[3]);
    const float zb = (part[4] + part[5]) + (part[6] + part[7]);
    const float l2 = part[8];
    const float invZ = Z > 0.f ? 1.0f / Z : 0.f;
    if (rt == 0 && !(Z > 0.f && Z < 3.0e38f)) atomicOr(d.flags, 2);   // no usable posterior for this frame
    // ---- y = grad_scale * softmax(row), in place in the row slot: eight 16-byte loads in flight per thread
    float4 *a4 = reinterpret_cast<float4 *>(arow);
    for (int k0 = rt; k0 < n4; k0 += 8 * NR) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (k0 + u * NR < n4) v[u] = a4[k0 + u * NR];
#pragma unroll
      for (int u = 0; u < 8; u++)
        if (k0 + u * NR < n4) {
          v[u].x = gs * ex2_approx(fmaf(v[u].x, kLog2e, -l2));
          v[u].y = gs * ex2_approx(fmaf(v[u].y, kLog2e, -l2));
          v[u].z = gs * ex2_approx(fmaf(v[u].z, kLog2e, -l2));
          v[u].w = gs * ex2_approx(fmaf(v[u].w, kLog2e, -l2));
          a4[k0 + u * NR] = v[u];
        }
    }
    PC_MARK(1);
    named_bar_sync(2, NR);  // the scaled row is complete
    PC_MARK(2);
    // ---- subtract the posterior mass of every distinct label (fixed summation order) and of the blank;
    //      four labels per thread in flight (the chain of dependent shared-memory loads is what costs here)
    const float gz = gs * invZ;
    for (int j0 = rt; j0 < nuniq; j0 += 4 * NR) {
      int q0[4], q1[4], sym[4];
      float acc[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int j = j0 + u * NR;
        const bool ok = j < nuniq;
        q0[u] = ok ? sm.s_us[j] : 0;
        q1[u] = ok ? sm.s_us[j + 1] : 0;
        sym[u] = ok ? sm.s_ul[j] : 0;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) acc[u] = q1[u] > q0[u] ? gam[sm.s_pos[q0[u]]] : 0.f;
#pragma unroll
      for (int u = 0; u < 4; u++)
        for (int q = q0[u] + 1; q < q1[u]; q++) acc[u] += gam[sm.s_pos[q]];
#pragma unroll
      for (int u = 0; u < 4; u++)
        if (j0 + u * NR < nuniq) arow[sym[u]] -= gz * acc[u];
    }
    if (rt == 0) arow[d.blank] -= gz * zb;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the bulk store
    __syncwarp();
    if (lane == 0) {
      mbar_arrive(sm.done_act + slotA);
      mbar_arrive(sm.gfree + slotG);
    }
    PC_MARK(3);
    if (++slotA == sc.NA) { slotA = 0; phA ^= 1u; }
    if (++slotG == kNG) { slotG = 0; phG ^= 1u; }
  }
  PC_FLUSH(rt == 0 && blockIdx.x == 0, 24);
}

__global__ void __launch_bounds__(kSThreads, 2) ctc_stream_beta_grad_kernel(CtcDev d, StreamCfg sc) {
  extern __shared__ __align__(128) unsigned char ssm[];
  SbSmem sm;
  sm.full_act = reinterpret_cast<uint64_t *>(ssm);   // [NA <= 8]
  sm.done_act = sm.full_act + 8;                      // [NA]
  sm.full_tab = sm.done_act + 8;                      // [NT <= 4]
  sm.free_tab = sm.full_tab + 4;                      // [NT]
  sm.gready = sm.free_tab + 4;                        // [kNG <= 4]
  sm.gfree = sm.gready + 4;                           // [kNG]            32 barriers = 256 B
  sm.part = reinterpret_cast<float *>(ssm + 256);     // [kNG][12]
  sm.bnd = reinterpret_cast<float2 *>(ssm + 448);     // [2][4]
  sm.act_base = ssm + 512;
  sm.tab_base = sm.act_base + (size_t)sc.NA * sc.act_bytes;
  sm.gam = reinterpret_cast<float *>(sm.tab_base + (size_t)sc.NT * sc.tab_bytes);
  sm.s_us = reinterpret_cast<int *>(sm.gam + kNG * sc.pitch_max);   // [pitch_max + 4]
  sm.s_pos = sm.s_us + sc.pitch_max + 4;                            // [pitch_max]
  sm.s_ul = sm.s_pos + sc.pitch_max;                                // [pitch_max]

  const int b = sc.order[blockIdx.x];
  const UttMeta um = d.meta[b];
  const int tid = threadIdx.x, lane = tid & 31, wi = tid >> 5;
  const int A = d.A, T = um.feasible ? um.T : 0;
  if (tid == 0) {
    for (int i = 0; i < sc.NA; i++) {
      mbar_init(sm.full_act + i, 1);
      mbar_init(sm.done_act + i, (kSC - kSChain) / 32);
    }
    for (int i = 0; i < sc.NT; i++) {
      mbar_init(sm.full_tab + i, 1);
      mbar_init(sm.free_tab + i, kSChain / 32);   // (unused since the chain feeds the table ring itself)
    }
    for (int i = 0; i < kNG; i++) {
      mbar_init(sm.gready + i, kSChain / 32);
      mbar_init(sm.gfree + i, (kSC - kSChain) / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (wi == kSC / 32) {
    if (lane == 0 && T > 0) {
      // ============ producer: keeps the row ring full and drains finished rows (time runs backwards);
      //              the alpha-table ring is fed by the chain (see sb_chain) ============
      const uint32_t row_bytes = (uint32_t)A * 4u;
      auto issue_act = [&](int k, int slot) {
        const int t = T - 1 - k;
        uint64_t *bar = sm.full_act + slot;
        mbar_expect_tx(bar, row_bytes);
        tma_load_1d(sm.act_base + (size_t)slot * sc.act_bytes, d.act + ((long long)t * d.B + b) * A, row_bytes, bar);
      };
      for (int k = 0; k < min(sc.NA - 1, T); k++) issue_act(k, k);
      int slot = 0, nslot = sc.NA - 1;   // slot of row k / of row k + NA - 1
      uint32_t ph = 0;
      PC_DECL;
      for (int k = 0; k < T; k++) {
        PC_MARK(7);
        mbar_wait(sm.done_act + slot, ph);  // row k finished in its slot
        PC_MARK(1);
        bulk_store(d.grad + ((long long)(T - 1 - k) * d.B + b) * A, sm.act_base + (size_t)slot * sc.act_bytes, row_bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        // the slot of row k-1 (its store was committed one row ago) takes row k+NA-1
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        PC_MARK(2);
        if (k + sc.NA - 1 < T) issue_act(k + sc.NA - 1, nslot);
        if (++slot == sc.NA) { slot = 0; ph ^= 1u; }
        if (++nslot == sc.NA) nslot = 0;
      }
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      PC_FLUSH(blockIdx.x == 0, 32);
    }
    return;
  }
  if (T > 0) {
    if (wi >= kSChain / 32) {
      // row warps: the utterance's label -> positions CSR into shared memory, then the row loop
      const int rt = tid - kSChain, nrt = kSC - kSChain;
      const int nuniq = d.nuniq[b];
      const int *ul = d.uniq_lab + um.csr_off, *us = d.uniq_start + um.csr_off + b, *pos = d.pos + um.lab_off;
      for (int k = rt; k <= nuniq; k += nrt) sm.s_us[k] = us[k];
      for (int k = rt; k < nuniq; k += nrt) sm.s_ul[k] = ul[k];
      for (int k = rt; k < um.L; k += nrt) sm.s_pos[k] = pos[k];
      named_bar_sync(2, nrt);
      sb_rows(d, um, rt, sc, sm, nuniq);
    } else {
      switch (stream_pshift(um.L)) {
        case 0: sb_chain<1>(d, um, b, tid, sc, sm); break;
        case 1: sb_chain<2>(d, um, b, tid, sc, sm); break;
        case 2: sb_chain<4>(d, um, b, tid, sc, sm); break;
        default: sb_chain<8>(d, um, b, tid, sc, sm); break;
      }
    }
  }
  // ---- zero rows: padded frames (all frames of an unalignable utterance)
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int t = T; t < d.Tmax; t++) {
    float4 *g4 = reinterpret_cast<float4 *>(d.grad + ((long long)t * d.B + b) * A);
    for (int k = tid; k < (A >> 2); k += kSC) g4[k] = zero4;
  }
}

// ===========================================================================
// host side
// ===========================================================================
struct Plan {
  int A, B, Tmax, maxL, pitch_max;
  long long sumT, sumL;
  size_t off_meta, off_order, off_labels, off_uniq_lab, off_uniq_start, off_pos, off_nuniq;  // header block
  size_t header_bytes;
  size_t off_lse2, off_E, off_alpha, off_beta, off_offA, off_offB, off_logp2, off_costs, off_flags;
  size_t total;
  std::vector<UttMeta> meta;
};

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

ctcStatus_t make_plan(const int *label_lengths, const int *input_lengths, int A, int B, Plan *p) {
  if (!label_lengths || !input_lengths || A <= 0 || B <= 0) return CTC_STATUS_INVALID_VALUE;
  p->A = A;
  p->B = B;
  p->Tmax = 0;
  p->maxL = 0;
  p->sumT = p->sumL = 0;
  p->meta.assign(B, UttMeta());
  long long e_off = 0, ab_off = 0, off_off = 0;
  for (int b = 0; b < B; b++) {
    const int T = input_lengths[b], L = label_lengths[b];
    if (T <= 0 || L < 0) return CTC_STATUS_INVALID_VALUE;
    UttMeta &m = p->meta[b];
    m.T = T;
    m.L = L;
    m.lab_off = (int)p->sumL;
    m.pitch = (int)align_up((size_t)L + 1, 4);
    m.csr_off = (int)p->sumL;
    m.e_off = e_off;
    m.ab_off = ab_off;
    m.off_off = off_off;
    off_off += (long long)((T + kRenorm - 1) / kRenorm) * m.pitch;  // row stride <= align4(L+1) for any P
    e_off += (long long)T * m.pitch;
    ab_off += (long long)T * 2 * m.pitch;
    p->sumT += T;
    p->sumL += L;
    p->Tmax = std::max(p->Tmax, T);
    p->maxL = std::max(p->maxL, L);
  }
  p->pitch_max = (int)align_up((size_t)p->maxL + 1, 4);
  if (p->maxL + 1 > 512 * 4) return CTC_STATUS_INVALID_VALUE;  // P <= 4, 512 threads per direction
  size_t o = 0;
  p->off_meta = o;        o = align_up(o + sizeof(UttMeta) * B, 256);
  p->off_order = o;       o = align_up(o + sizeof(int) * B, 256);
  p->off_labels = o;      o = align_up(o + sizeof(int) * (p->sumL + 1), 256);
  p->off_uniq_lab = o;    o = align_up(o + sizeof(int) * (p->sumL + 1), 256);
  p->off_uniq_start = o;  o = align_up(o + sizeof(int) * (p->sumL + B + 1), 256);
  p->off_pos = o;         o = align_up(o + sizeof(int) * (p->sumL + 1), 256);
  p->off_nuniq = o;       o = align_up(o + sizeof(int) * B, 256);
  p->header_bytes = o;
  p->off_lse2 = o;   o = align_up(o + sizeof(float) * (size_t)p->Tmax * B, 256);
  p->off_E = o;      o = align_up(o + sizeof(float) * (size_t)e_off, 256);
  p->off_alpha = o;  o = align_up(o + sizeof(float) * (size_t)ab_off, 256);
  p->off_beta = o;   o = align_up(o + sizeof(float) * (size_t)ab_off, 256);
  p->off_offA = o;   o = align_up(o + sizeof(float) * (size_t)off_off, 256);
  p->off_offB = o;   o = align_up(o + sizeof(float) * (size_t)off_off, 256);
  p->off_logp2 = o;  o = align_up(o + sizeof(double) * 2 * B, 256);
  p->off_costs = o;  o = align_up(o + sizeof(float) * B, 256);
  p->off_flags = o;  o = align_up(o + 256, 256);
  p->total = o;
  return CTC_STATUS_SUCCESS;
}

// pinned staging for the header block (one H2D copy per call) and the costs
struct Staging {
  unsigned char *pinned = nullptr;
  size_t cap = 0;
  float *costs = nullptr;
  size_t costs_cap = 0;
  cudaEvent_t copied = nullptr;   // recorded after the last upload out of `pinned`
  bool pending = false;
};

// Tuning switches: read from the environment ONCE per process (none is needed in production).
struct Tuning {
  int force_p, ring, na, nt, profile, groups, one_stream, stream_path, stream_na;
};
const Tuning &tuning() {
  static const Tuning t = [] {
    auto geti = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
    Tuning v;
    v.force_p = geti("B200CTC_P", 0);
    v.ring = geti("B200CTC_RING", -1);
    v.na = geti("B200CTC_NA", 0);
    v.nt = geti("B200CTC_NT", 0);
    v.profile = geti("B200CTC_PROFILE", 0);
    v.groups = geti("B200CTC_GROUPS", 0);
    v.one_stream = geti("B200CTC_ONE_STREAM", 0);
    v.stream_path = geti("B200CTC_STREAM", -1);   // 0 / 1: never / whenever possible use the streaming path
    v.stream_na = geti("B200CTC_STREAM_NA", 0);
    return v;
  }();
  return t;
}

// Everything that belongs to ONE device: function attributes (cudaFuncSetAttribute is per device), the
// side streams and events of the grouped launch, the SM count and the staging event.  Keyed by
// cudaGetDevice() so that an in-process multi-GPU caller of the C ABI gets a consistent set per GPU.
constexpr int kMaxGroups = 8;
struct DeviceState {
  int num_sms = 0;
  bool k2_attr[3] = {false, false, false};   // P = 1, 2, 4
  size_t smem3_set = 0, ring_set = 0, sa_set = 0, sb_set = 0;
  cudaStream_t hp[kMaxGroups] = {};
  cudaEvent_t e1[kMaxGroups] = {}, e2[kMaxGroups] = {};
  Staging stage;
};
std::mutex g_mu;                       // one call at a time per process (the staging block is shared state)
DeviceState *device_state() {          // call with g_mu held
  static std::vector<DeviceState *> states;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) return nullptr;
  if ((size_t)dev >= states.size()) states.resize(dev + 1, nullptr);
  if (!states[dev]) {
    states[dev] = new DeviceState();
    cudaDeviceGetAttribute(&states[dev]->num_sms, cudaDevAttrMultiProcessorCount, dev);
  }
  return states[dev];
}

bool ensure_pinned(Staging &s, size_t bytes, size_t ncosts) {
  if (bytes > s.cap) {
    if (s.pinned) cudaFreeHost(s.pinned);
    s.pinned = nullptr;
    s.cap = 0;
    size_t want = align_up(bytes * 2, 4096);
    if (cudaMallocHost(&s.pinned, want) != cudaSuccess) return false;
    s.cap = want;
  }
  if (ncosts > s.costs_cap) {
    if (s.costs) cudaFreeHost(s.costs);
    s.costs = nullptr;
    s.costs_cap = 0;
    if (cudaMallocHost(&s.costs, sizeof(float) * ncosts * 2) != cudaSuccess) return false;
    s.costs_cap = ncosts * 2;
  }
  return true;
}

// costs / flag word out; synchronise unless the caller asked not to
ctcStatus_t finish_call(const CtcDev &dev, int B, float *costs_host, float *costs_dev, const b200ctcOptions &opt,
                        cudaStream_t stream, Staging &g_stage) {
  if (costs_dev &&
      cudaMemcpyAsync(costs_dev, dev.costs, sizeof(float) * B, cudaMemcpyDeviceToDevice, stream) !=
          cudaSuccess)
    return CTC_STATUS_MEMOPS_FAILED;
  if (opt.nonfinite_dev &&
      cudaMemcpyAsync(opt.nonfinite_dev, dev.flags, sizeof(int), cudaMemcpyDeviceToDevice, stream) != cudaSuccess)
    return CTC_STATUS_MEMOPS_FAILED;
  if (!opt.no_sync) {
    if (cudaMemcpyAsync(g_stage.costs, dev.costs, sizeof(float) * B, cudaMemcpyDeviceToHost,
                        stream) != cudaSuccess)
      return CTC_STATUS_MEMOPS_FAILED;
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) {
      fprintf(stderr, "b200ctc: %s\n", cudaGetErrorString(e));
      return CTC_STATUS_EXECUTION_FAILED;
    }
    if (costs_host) memcpy(costs_host, g_stage.costs, sizeof(float) * B);
  }
  return CTC_STATUS_SUCCESS;
}

template <int P>
cudaError_t launch_k2(const CtcDev &dev, int B, int NT, int F, cudaStream_t stream, DeviceState *ds) {
  const size_t smem = 2048 + sizeof(float) * 2 * kStages * kStageFloats;
  bool &attr_done = ds->k2_attr[P == 1 ? 0 : (P == 2 ? 1 : 2)];
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(ctc_alpha_beta_kernel<P>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  ctc_alpha_beta_kernel<P><<<dev.nb, 2 * NT, smem, stream>>>(dev, F);
  return cudaGetLastError();
}

ctcStatus_t run(const float *act, float *grad, const int *flat_labels, const int *label_lengths,
                const int *input_lengths, int A, int B, float *costs_host, float *costs_dev,
                void *workspace, size_t workspace_bytes, b200ctcOptions opt) {
  if (!act || !flat_labels || !workspace || (!costs_host && !costs_dev))
    return CTC_STATUS_INVALID_VALUE;
  if (opt.blank_label < 0 || opt.blank_label >= A) return CTC_STATUS_INVALID_VALUE;
  if ((uintptr_t)workspace % 256 != 0) return CTC_STATUS_INVALID_VALUE;
  const auto host_t0 = std::chrono::steady_clock::now();
  Plan p;
  ctcStatus_t st = make_plan(label_lengths, input_lengths, A, B, &p);
  if (st != CTC_STATUS_SUCCESS) return st;
  if (workspace_bytes < p.total) return CTC_STATUS_INVALID_VALUE;
  cudaStream_t stream = (cudaStream_t)opt.stream;

  std::lock_guard<std::mutex> lock(g_mu);
  DeviceState *ds = device_state();
  if (!ds) return CTC_STATUS_EXECUTION_FAILED;
  Staging &g_stage = ds->stage;
  const Tuning &tune = tuning();
  if (!ensure_pinned(g_stage, p.header_bytes, (size_t)B)) return CTC_STATUS_MEMOPS_FAILED;
  unsigned char *h = g_stage.pinned;
  UttMeta *hm = reinterpret_cast<UttMeta *>(h + p.off_meta);
  int *hl = reinterpret_cast<int *>(h + p.off_labels);
  int *hul = reinterpret_cast<int *>(h + p.off_uniq_lab);
  int *hus = reinterpret_cast<int *>(h + p.off_uniq_start);
  int *hpos = reinterpret_cast<int *>(h + p.off_pos);
  int *hnu = reinterpret_cast<int *>(h + p.off_nuniq);
  // the previous call's uploads out of the pinned block may still be queued (no_sync callers): wait for them
  if (!g_stage.copied && cudaEventCreateWithFlags(&g_stage.copied, cudaEventDisableTiming) != cudaSuccess)
    return CTC_STATUS_MEMOPS_FAILED;
  if (g_stage.pending) {
    cudaEventSynchronize(g_stage.copied);
    g_stage.pending = false;
  }
  // ---- stage 1: what K1 and K2 need (lengths, offsets, labels, feasibility) -> device, kernels queued;
  //      stage 2 (below, while those kernels run): the label -> positions CSR that only K3 reads
  long long vrows_total = 0;
  for (int b = 0; b < B; b++) {
    UttMeta &m = p.meta[b];
    const int *lab = flat_labels + m.lab_off;
    int repeats = 0;
    for (int i = 0; i < m.L; i++) {
      if (lab[i] < 0 || lab[i] >= A || lab[i] == opt.blank_label) return CTC_STATUS_INVALID_VALUE;
      if (i > 0 && lab[i] == lab[i - 1]) repeats++;
      hl[m.lab_off + i] = lab[i];
    }
    m.feasible = (m.L + repeats <= m.T) ? 1 : 0;
    m.nuniq = 0;  // (device code reads CtcDev::nuniq)
    m.vrow0 = vrows_total;
    if (m.feasible) vrows_total += m.T;
    hm[b] = m;
  }
  {  // CTA -> utterance map of the streaming kernels: longest utterances first (they bound the run time)
    int *ho = reinterpret_cast<int *>(h + p.off_order);
    for (int b = 0; b < B; b++) ho[b] = b;
    std::stable_sort(ho, ho + B, [&](int x, int y) { return p.meta[x].T > p.meta[y].T; });
  }
  unsigned char *w = static_cast<unsigned char *>(workspace);
  if (cudaMemcpyAsync(w, h, p.off_uniq_lab, cudaMemcpyHostToDevice, stream) != cudaSuccess)
    return CTC_STATUS_MEMOPS_FAILED;
  if (cudaMemsetAsync(w + p.off_costs, 0, (p.off_flags + 256) - p.off_costs, stream) != cudaSuccess)
    return CTC_STATUS_MEMOPS_FAILED;
  // label -> states CSR in O(L) per utterance: groups in order of first appearance, positions
  // ascending inside a group (a fixed summation order => deterministic gradients)
  auto build_and_upload_csr = [&]() -> bool {
    static thread_local std::vector<int> stamp, slot, cnt;
    if ((int)stamp.size() < A) {
      stamp.assign(A, -1);
      slot.assign(A, 0);
    }
    static thread_local int epoch = 0;
    for (int b = 0; b < B; b++) {
      const UttMeta &m = p.meta[b];
      const int *lab = flat_labels + m.lab_off;
      ++epoch;
      int nu = 0;
      int *us = hus + m.csr_off + b;
      cnt.assign(m.L + 1, 0);
      for (int i = 0; i < m.L; i++) {
        const int k = lab[i];
        if (stamp[k] != epoch) {
          stamp[k] = epoch;
          slot[k] = nu;
          hul[m.csr_off + nu] = k;
          nu++;
        }
        cnt[slot[k]]++;
      }
      int run_ = 0;
      for (int j = 0; j < nu; j++) {
        us[j] = run_;
        run_ += cnt[j];
        cnt[j] = us[j];  // becomes the write cursor of group j
      }
      us[nu] = m.L;
      for (int i = 0; i < m.L; i++) hpos[m.lab_off + cnt[slot[lab[i]]]++] = i;
      hnu[b] = nu;
    }
    if (cudaMemcpyAsync(w + p.off_uniq_lab, h + p.off_uniq_lab, p.header_bytes - p.off_uniq_lab,
                        cudaMemcpyHostToDevice, stream) != cudaSuccess)
      return false;
    cudaEventRecord(g_stage.copied, stream);
    g_stage.pending = true;
    return true;
  };

  CtcDev dev;
  dev.act = act;
  dev.grad = grad;
  dev.A = A;
  dev.B = B;
  dev.Tmax = p.Tmax;
  dev.blank = opt.blank_label;
  dev.grad_scale = opt.grad_scale;
  dev.meta = reinterpret_cast<const UttMeta *>(w + p.off_meta);
  dev.labels = reinterpret_cast<const int *>(w + p.off_labels);
  dev.uniq_lab = reinterpret_cast<const int *>(w + p.off_uniq_lab);
  dev.uniq_start = reinterpret_cast<const int *>(w + p.off_uniq_start);
  dev.pos = reinterpret_cast<const int *>(w + p.off_pos);
  dev.nuniq = reinterpret_cast<const int *>(w + p.off_nuniq);
  dev.lse2 = reinterpret_cast<float *>(w + p.off_lse2);
  dev.E = reinterpret_cast<float *>(w + p.off_E);
  dev.alpha = reinterpret_cast<float *>(w + p.off_alpha);
  dev.beta = reinterpret_cast<float *>(w + p.off_beta);
  dev.offA = reinterpret_cast<float *>(w + p.off_offA);
  dev.offB = reinterpret_cast<float *>(w + p.off_offB);
  dev.logp2 = reinterpret_cast<double *>(w + p.off_logp2);
  dev.costs = reinterpret_cast<float *>(w + p.off_costs);
  dev.flags = reinterpret_cast<int *>(w + p.off_flags);
  dev.argmax = opt.argmax_dev;
  dev.pc = nullptr;
#ifdef B200CTC_PHASE_COUNTERS
  static long long *pc_dev = nullptr;
  if (!pc_dev) cudaMalloc(&pc_dev, 64 * sizeof(long long));
  cudaMemsetAsync(pc_dev, 0, 64 * sizeof(long long), stream);
  dev.pc = pc_dev;
#endif

  // K2 geometry: P pairs per thread so that one direction fits 512 threads
  const int npairs = p.maxL + 1;
  // (measured: above ~8 warps per direction the frame loop is issue-bound and 2 pairs per thread win)
  int P = npairs <= 256 ? 1 : (npairs <= 1024 ? 2 : 4);
  const int force_p = tune.force_p;   // tuning aid
  if ((force_p == 2 || force_p == 4) && force_p > P) P = force_p;
  dev.P = P;
  // Launch K1 -> K2 -> K3 per utterance group.  With two groups on two streams the latency-bound
  // alpha/beta recursion of one group runs under the bandwidth-bound row kernels of the other.
  const int NT = (int)align_up((size_t)(npairs + P - 1) / P, 32);
  const int F = std::max(1, std::min(32, kStageFloats / p.pitch_max));
  const int smem_pitch = p.pitch_max;  // gammas of the label states of one row
  const size_t smem3 = sizeof(float) * (size_t)kK3Warps * smem_pitch;
  if (grad) {
    size_t &smem3_set = ds->smem3_set;
    if (smem3 > smem3_set) {
      if (cudaFuncSetAttribute(ctc_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem3) != cudaSuccess)
        return CTC_STATUS_EXECUTION_FAILED;
      smem3_set = smem3;
    }
  }
  // Wide alphabets: the persistent TMA-ring gradient kernel (see ctc_grad_ring_kernel)
  const int num_sms = ds->num_sms > 0 ? ds->num_sms : 148;
  const int env_ring = tune.ring, env_na = tune.na, env_nt = tune.nt;   // tuning aids
  RingCfg rc;
  rc.NA = env_na >= 2 ? env_na : 4;
  rc.NT = env_nt >= 1 ? env_nt : 2;
  rc.act_bytes = (int)align_up((size_t)A * 4, 128);
  rc.tab_bytes = (int)align_up((size_t)7 * p.pitch_max * 4, 128);
  rc.pitch_max = p.pitch_max;
  rc.pshift = P == 1 ? 0 : (P == 2 ? 1 : 2);
  rc.vrow_base = rc.vrows = 0;
  auto ring_bytes = [&]() {
    return (size_t)512 + (size_t)rc.NA * rc.act_bytes + (size_t)rc.NT * rc.tab_bytes +
           sizeof(float) * ((size_t)5 * p.pitch_max + 8);
  };
  // two CTAs per SM (16 consumer warps keep the issue slots busy): <= ~113 KB each
  bool use_ring = grad && (A & 3) == 0 && A >= 1024 && (size_t)p.Tmax * B * A >= ((size_t)64 << 20);
  if (env_ring == 1 && grad && (A & 3) == 0) use_ring = true;
  if (env_ring == 0) use_ring = false;
  while (use_ring && ring_bytes() > ((size_t)113 << 10) && rc.NA > 3) rc.NA--;
  if (ring_bytes() > ((size_t)113 << 10) || rc.NA + rc.NT > 16) use_ring = false;
  const size_t ring_smem = ring_bytes();
  if (use_ring) {
    size_t &ring_set = ds->ring_set;
    if (ring_smem > ring_set) {
      if (cudaFuncSetAttribute(ctc_grad_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)ring_smem) != cudaSuccess)
        return CTC_STATUS_EXECUTION_FAILED;
      ring_set = ring_smem;
    }
  }
  // ---- streaming path: one persistent CTA per utterance, the recursions ride the two slab reads
  {
    StreamCfg sc;
    sc.act_bytes = (int)align_up((size_t)A * 4, 128);
    sc.tab_bytes = (int)align_up((size_t)3 * p.pitch_max * 4, 128);
    sc.pitch_max = p.pitch_max;
    sc.order = reinterpret_cast<const int *>(w + p.off_order);
    sc.NT = 2;
    sc.NA = tune.stream_na >= 3 && tune.stream_na <= 8 ? tune.stream_na : 6;   // (SB; SA's ring is kNAs)
    auto sb_bytes = [&]() {
      return (size_t)512 + (size_t)sc.NA * sc.act_bytes + (size_t)sc.NT * sc.tab_bytes +
             sizeof(float) * ((size_t)(kNG + 3) * p.pitch_max + 8);
    };
    while (sb_bytes() > ((size_t)113 << 10) && sc.NA > 3) sc.NA--;   // two CTAs per SM when the rows allow it
    const size_t sb_smem = sb_bytes(), sa_smem = (size_t)512 + (size_t)kNAs * sc.act_bytes;
    const bool slab_big = (size_t)p.Tmax * B * A >= ((size_t)64 << 20);
    bool use_stream = (A & 3) == 0 && A >= 1024 && slab_big && B >= 96 && p.maxL + 1 <= 8 * kSChain &&
                      sb_smem <= ((size_t)220 << 10) && (uintptr_t)act % 16 == 0 && (uintptr_t)grad % 16 == 0;
    if (tune.stream_path != 1) use_stream = false;   // (not yet the default: slower than the three-kernel path, see profiles/)
    if (tune.stream_path == 1)
      use_stream = (A & 3) == 0 && p.maxL + 1 <= 8 * kSChain && sb_smem <= ((size_t)220 << 10) &&
                   (uintptr_t)act % 16 == 0 && (uintptr_t)grad % 16 == 0;
    if (use_stream) {
      if (sa_smem > ds->sa_set) {
        if (cudaFuncSetAttribute(ctc_stream_alpha_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sa_smem) !=
            cudaSuccess)
          return CTC_STATUS_EXECUTION_FAILED;
        ds->sa_set = sa_smem;
      }
      if (grad && sb_smem > ds->sb_set) {
        if (cudaFuncSetAttribute(ctc_stream_beta_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)sb_smem) != cudaSuccess)
          return CTC_STATUS_EXECUTION_FAILED;
        ds->sb_set = sb_smem;
      }
      dev.b_lo = 0;
      dev.nb = B;
      cudaEvent_t sev[3];
      const bool sprof = tune.profile == 1;
      if (sprof) {
        for (int i = 0; i < 3; i++) cudaEventCreate(&sev[i]);
        cudaEventRecord(sev[0], stream);
      }
      ctc_stream_alpha_kernel<<<B, kSThreads, sa_smem, stream>>>(dev, sc);
      if (cudaGetLastError() != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;
      if (sprof) cudaEventRecord(sev[1], stream);
      if (grad) {
        if (!build_and_upload_csr()) return CTC_STATUS_MEMOPS_FAILED;   // host work under the alpha kernel
        ctc_stream_beta_grad_kernel<<<B, kSThreads, sb_smem, stream>>>(dev, sc);
        if (cudaGetLastError() != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;
      } else {
        cudaEventRecord(g_stage.copied, stream);
        g_stage.pending = true;
      }
      if (sprof) {
        cudaEventRecord(sev[2], stream);
        cudaEventSynchronize(sev[2]);
        float t1 = 0, t2 = 0;
        cudaEventElapsedTime(&t1, sev[0], sev[1]);
        cudaEventElapsedTime(&t2, sev[1], sev[2]);
        fprintf(stderr, "[b200ctc] streaming path B=%d A=%d Tmax=%d maxL=%d NA=%d: alpha %.3f ms, beta+grad %.3f ms\n", B, A,
                p.Tmax, p.maxL, sc.NA, t1, t2);
        for (int i = 0; i < 3; i++) cudaEventDestroy(sev[i]);
#ifdef B200CTC_PHASE_COUNTERS
        long long h[64];
        cudaMemcpy(h, dev.pc, sizeof(h), cudaMemcpyDeviceToHost);
        const double n = p.meta[reinterpret_cast<const int *>(g_stage.pinned + p.off_order)[0]].T;
        fprintf(stderr, "[b200ctc pc] CTA 0 (T=%.0f), cycles per frame\n  SA chain: wait %.0f gather %.0f step %.0f store+handover %.0f barrier %.0f loop %.0f\n"
                "  SA stats(per own row): wait %.0f work %.0f loop %.0f\n"
                "  SB chain: wait_gfree %.0f wait_tab %.0f wait_act %.0f gather+step %.0f gamma+sums %.0f handover+barrier %.0f loop %.0f\n"
                "  SB rows: wait %.0f softmax %.0f barrier %.0f fixup+fence+arrive %.0f loop %.0f\n"
                "  SB producer: tab %.0f wait_done %.0f store+wait_read %.0f loop %.0f\n",
                n, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[7] / n, h[8] / n * 4, h[9] / n * 4, h[15] / n * 4,
                h[16] / n, h[17] / n, h[18] / n, h[19] / n, h[20] / n, h[21] / n, h[23] / n,
                h[24] / n, h[25] / n, h[26] / n, h[27] / n, h[31] / n, h[32] / n, h[33] / n, h[34] / n, h[39] / n);
#endif
      }
      return finish_call(dev, B, costs_host, costs_dev, opt, stream, g_stage);
    }
  }
  // tuning aid: B200CTC_PROFILE=1 serialises the groups and prints the duration of each kernel
  const int prof_mode = tune.profile;
  const bool prof = prof_mode == 1;
  const bool timeline = prof_mode == 2;   // keeps the groups; prints each kernel's start/end on its own stream
  // (only worth it when the row kernels are long: a slab of >= 256 MB)
  const bool big = (size_t)p.Tmax * B * A >= ((size_t)64 << 20);
  // Utterance groups on separate streams: the latency-bound alpha/beta recursion of one group runs
  // under the bandwidth-bound row kernels of the others (only the first K2 and the last K3 stay exposed).
  const int env_groups = tune.groups;
  int ngroups = B >= 16 ? 2 : 1;   // measured at B=256, A=4000: 1 -> 10.47 ms, 2 -> 9.95, 4 -> 9.91, 8 -> 9.99
  if (!big) ngroups = 1;
  if (env_groups >= 1 && env_groups <= kMaxGroups) ngroups = std::min(env_groups, B);
  if (prof || tune.one_stream) ngroups = 1;
  // Several groups: the row kernels (K1, K3) of all groups run back to back on the caller's stream; each
  // group's alpha/beta kernel runs on its own HIGH-PRIORITY stream between them, so its CTAs are placed as
  // soon as the group's K1 is done instead of queueing behind the row kernels of the other groups.
  cudaStream_t *hp = ds->hp;
  cudaEvent_t *e1 = ds->e1, *e2 = ds->e2;
  if (ngroups > 1) {
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    for (int gi = 0; gi < ngroups; gi++)
      if (!hp[gi] && (cudaStreamCreateWithPriority(&hp[gi], cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
                      cudaEventCreateWithFlags(&e1[gi], cudaEventDisableTiming) != cudaSuccess ||
                      cudaEventCreateWithFlags(&e2[gi], cudaEventDisableTiming) != cudaSuccess))
        return CTC_STATUS_EXECUTION_FAILED;
  }
  cudaEvent_t tl[kMaxGroups][4], tl0;
  if (timeline) {
    fprintf(stderr, "[b200ctc] host preparation %.3f ms\n",
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count());
    cudaEventCreate(&tl0);
    for (int gi = 0; gi < ngroups; gi++)
      for (int i = 0; i < 4; i++) cudaEventCreate(&tl[gi][i]);
    cudaEventRecord(tl0, stream);
  }
  cudaEvent_t ev[4];
  if (prof) {
    for (int i = 0; i < 4; i++) cudaEventCreate(&ev[i]);
    cudaEventRecord(ev[0], stream);
  }
  auto set_group = [&](int gi) {
    dev.b_lo = (int)((long long)B * gi / ngroups);
    dev.nb = (int)((long long)B * (gi + 1) / ngroups) - dev.b_lo;
    return (long long)p.Tmax * dev.nb;
  };
  for (int gi = 0; gi < ngroups; gi++) {  // K1
    const long long rows = set_group(gi);
    const unsigned g1 = (unsigned)((rows + kK1Warps - 1) / kK1Warps);
    if (timeline) cudaEventRecord(tl[gi][0], stream);
    ctc_rowstats_gather_kernel<<<g1, kK1Warps * 32, 0, stream>>>(dev);
    if (cudaGetLastError() != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;
    if (timeline) cudaEventRecord(tl[gi][1], stream);
    if (ngroups > 1) cudaEventRecord(e1[gi], stream);
  }
  if (prof) cudaEventRecord(ev[1], stream);
  for (int gi = 0; gi < ngroups; gi++) {  // K2
    set_group(gi);
    cudaStream_t st = ngroups > 1 ? hp[gi] : stream;
    if (ngroups > 1) cudaStreamWaitEvent(st, e1[gi], 0);
    cudaError_t ce = P == 1   ? launch_k2<1>(dev, B, NT, F, st, ds)
                     : P == 2 ? launch_k2<2>(dev, B, NT, F, st, ds)
                              : launch_k2<4>(dev, B, NT, F, st, ds);
    if (ce != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;
    if (timeline) cudaEventRecord(tl[gi][2], st);
    if (ngroups > 1) cudaEventRecord(e2[gi], st);
  }
  if (prof) cudaEventRecord(ev[2], stream);
  if (grad) {
    if (!build_and_upload_csr()) return CTC_STATUS_MEMOPS_FAILED;   // host work under the kernels queued above
  } else {
    cudaEventRecord(g_stage.copied, stream);
    g_stage.pending = true;
  }
  for (int gi = 0; gi < ngroups; gi++) {  // K3
    const long long rows = set_group(gi);
    if (ngroups > 1) cudaStreamWaitEvent(stream, e2[gi], 0);  // (also joins the side stream back)
    if (grad && use_ring) {
      rc.vrow_base = p.meta[dev.b_lo].vrow0;
      rc.vrows = (dev.b_lo + dev.nb < B ? p.meta[dev.b_lo + dev.nb].vrow0 : vrows_total) - rc.vrow_base;
      const unsigned g3 = (unsigned)std::max<long long>(1, std::min<long long>(2 * num_sms, (rows + 7) / 8));
      ctc_grad_ring_kernel<<<g3, kRingThreads, ring_smem, stream>>>(dev, rc);
      if (cudaGetLastError() != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;
    } else if (grad) {
      const unsigned g3 = (unsigned)((rows + kK3Warps - 1) / kK3Warps);
      ctc_grad_kernel<<<g3, kK3Warps * 32, smem3, stream>>>(dev, smem_pitch);
      if (cudaGetLastError() != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;
    }
    if (timeline) cudaEventRecord(tl[gi][3], stream);
  }
  if (prof) {
    cudaEventRecord(ev[3], stream);
    cudaEventSynchronize(ev[3]);
    float t1 = 0, t2 = 0, t3 = 0;
    cudaEventElapsedTime(&t1, ev[0], ev[1]);
    cudaEventElapsedTime(&t2, ev[1], ev[2]);
    cudaEventElapsedTime(&t3, ev[2], ev[3]);
    fprintf(stderr, "[b200ctc] B=%d A=%d Tmax=%d maxL=%d P=%d F=%d: rowstats %.3f ms, alpha_beta %.3f ms, grad %.3f ms\n",
            B, A, p.Tmax, p.maxL, P, F, t1, t2, t3);
    for (int i = 0; i < 4; i++) cudaEventDestroy(ev[i]);
  }
  if (timeline) {
    cudaStreamSynchronize(stream);
    for (int gi = 0; gi < ngroups; gi++) {
      float t[4];
      for (int i = 0; i < 4; i++) {
        cudaEventElapsedTime(&t[i], tl0, tl[gi][i]);
        cudaEventDestroy(tl[gi][i]);
      }
      fprintf(stderr, "[b200ctc] group %d/%d: K1 %.3f..%.3f  K2 ..%.3f  K3 ..%.3f ms\n", gi, ngroups, t[0], t[1], t[2], t[3]);
    }
    cudaEventDestroy(tl0);
  }
  return finish_call(dev, B, costs_host, costs_dev, opt, stream, g_stage);
}

}  // namespace

extern "C" {

int get_warpctc_version(void) { return 2; }

const char *ctcGetStatusString(ctcStatus_t status) {
  switch (status) {
    case CTC_STATUS_SUCCESS: return "no error";
    case CTC_STATUS_MEMOPS_FAILED: return "cuda memcpy or memset failed";
    case CTC_STATUS_INVALID_VALUE: return "invalid value";
    case CTC_STATUS_EXECUTION_FAILED: return "execution failed";
    default: return "unknown error";
  }
}

ctcStatus_t b200ctc_workspace_size(const int *label_lengths, const int *input_lengths,
                                   int alphabet_size, int minibatch, size_t *size_bytes) {
  if (!size_bytes) return CTC_STATUS_INVALID_VALUE;
  Plan p;
  ctcStatus_t st = make_plan(label_lengths, input_lengths, alphabet_size, minibatch, &p);
  if (st != CTC_STATUS_SUCCESS) return st;
  *size_bytes = p.total;
  return CTC_STATUS_SUCCESS;
}

ctcStatus_t get_workspace_size(const int *const label_lengths, const int *const input_lengths,
                               int alphabet_size, int minibatch, struct ctcOptions options,
                               size_t *size_bytes) {
  if (options.loc != CTC_GPU) return CTC_STATUS_INVALID_VALUE;  // no CPU path in this library
  return b200ctc_workspace_size(label_lengths, input_lengths, alphabet_size, minibatch, size_bytes);
}

ctcStatus_t b200ctc_loss(const float *activations, float *gradients, const int *flat_labels,
                         const int *label_lengths, const int *input_lengths, int alphabet_size,
                         int minibatch, float *costs_host, float *costs_dev, void *workspace,
                         size_t workspace_bytes, b200ctcOptions options) {
  try {
    return run(activations, gradients, flat_labels, label_lengths, input_lengths, alphabet_size,
               minibatch, costs_host, costs_dev, workspace, workspace_bytes, options);
  } catch (...) {
    return CTC_STATUS_UNKNOWN_ERROR;  // never throw across the C boundary
  }
}

ctcStatus_t compute_ctc_loss(const float *const activations, float *gradients,
                             const int *const flat_labels, const int *const label_lengths,
                             const int *const input_lengths, int alphabet_size, int minibatch,
                             float *costs, void *workspace, struct ctcOptions options) {
  if (options.loc != CTC_GPU || !costs) return CTC_STATUS_INVALID_VALUE;
  b200ctcOptions o;
  o.blank_label = options.blank_label;
  o.grad_scale = 1.0f;
  o.stream = options.stream;
  o.no_sync = 0;
  o.argmax_dev = nullptr;
  o.nonfinite_dev = nullptr;
  size_t need = 0;
  ctcStatus_t st = b200ctc_workspace_size(label_lengths, input_lengths, alphabet_size, minibatch, &need);
  if (st != CTC_STATUS_SUCCESS) return st;
  // the warp-ctc ABI carries no workspace size: the caller allocated get_workspace_size() bytes
  return b200ctc_loss(activations, gradients, flat_labels, label_lengths, input_lengths,
                      alphabet_size, minibatch, costs, nullptr, workspace, need, o);
}

size_t b200ctc_algorithmic_bytes(const int *label_lengths, const int *input_lengths,
                                 int alphabet_size, int minibatch) {
  size_t sumT = 0, sumL = 0;
  int Tmax = 0;
  for (int b = 0; b < minibatch; b++) {
    sumT += input_lengths[b];
    sumL += label_lengths[b];
    Tmax = std::max(Tmax, input_lengths[b]);
  }
  return 4 * (size_t)alphabet_size * sumT + 4 * (size_t)alphabet_size * Tmax * minibatch +
         4 * sumL + 4 * (size_t)minibatch;
}

int b200ctc_launches_per_call(int with_gradients) { return with_gradients ? 3 : 2; }

}  // extern "C"
