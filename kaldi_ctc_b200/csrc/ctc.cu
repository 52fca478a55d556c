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
//   K2 ctc_alpha_beta<P>     one CTA per (utterance, direction), the longest
//        utterances first: the alpha (forward) and beta (backward) recursions
//        over the blank-interleaved lattice run CONCURRENTLY in separate CTAs,
//        P (blank,label) state pairs per thread,
//        neighbour states by warp shuffle (+ one smem word per warp boundary),
//        emissions staged into shared memory by TMA bulk copies
//        (cp.async.bulk + mbarrier, 4-stage ring).  Log-space,
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
  const int *order;       // [B] K2: CTA i of a group works on utterance order[b_lo + i] (longest first)
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
};

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
// recurrence as alpha, so both directions run this code:
//   Y_i <- Eb   + lse(Y_i, X_{i-1})
//   X_i <- El_i + lse(X_i, Y_i, skip_i ? X_{i-1} : 0)
// alpha: Y_i = state 2i, X_i = state 2i+1, time ascending, label i.
// beta : Y_i = state 2(L-i), X_i = state 2(L-i)-1, time descending, label L-1-i.
// The per-frame loop is bound by instruction issue and by the special-function unit (16 ex2/lg2 per clock and
// SM), not by latency: one CTA's 20 warps already keep an SM busy.  So the loop is written for instruction
// count: the role and "every pair of this warp exists" (FULL) are template parameters, every address advances
// incrementally, the three-term sum of the label state is nested on the two-term sum of the blank state
// (4 instead of 5 special-function operations per pair), and the end-of-block work (publish the offset,
// re-centre) sits BETWEEN runs of frames instead of being tested in every frame: an iteration is
// "exchange the previous frame's boundary state, then compute this frame".
// log2(2^a + 2^b) with 2 special-function operations and no sorting
__device__ __forceinline__ float lse2_pair(float a, float b) {
  return fmaxf(a, b) + lg2_approx(1.0f + ex2_approx(-fabsf(a - b)));
}

template <int P, int ROLE, bool FULL>
__device__ __forceinline__ void ctc_ab_run(const CtcDev &d, const UttMeta &um, int b, int r, int nthreads_needed,
                                           int nbar, uint64_t *my_bar, float *my_stage, int stage_floats,
                                           float2 *bnd, float *fin, int F) {
  const int lane = r & 31, w = r >> 5;
  const int L = um.L, T = um.T, pitch = um.pitch;
  const float *Eg = d.E + um.e_off;
  const int nchunks = (T + F - 1) / F;
  const bool multi = nbar > 32;   // several warps per direction: the warp-boundary state goes through shared memory

  // chunk k (in visiting order) -> first frame and frame count
  auto chunk_lo = [&](int k) { return ROLE ? max(0, T - (k + 1) * F) : k * F; };
  auto chunk_n = [&](int k) { return min(F, T - k * F); };
  auto issue = [&](int k, int st) {
    const uint32_t bytes = (uint32_t)chunk_n(k) * pitch * 4u;
    mbar_expect_tx(my_bar + st, bytes);
    tma_load_1d(my_stage + st * stage_floats, Eg + (long long)chunk_lo(k) * pitch, bytes, my_bar + st);
  };
  if (r == 0)
    for (int k = 0; k < min(kStages, nchunks); k++) issue(k, k);

  // per-thread lattice slice: pairs i0 .. i0+P-1, adjacent in E (labels) and in a stored frame
  const int i0 = r * P;
  const int *lab = d.labels + um.lab_off;
  float X[P], Y[P];
  bool skip[P], hasX[P], hasY[P];
#pragma unroll
  for (int p = 0; p < P; p++) {
    const int i = i0 + p;
    hasY[p] = FULL || i <= L;
    hasX[p] = FULL || i < L;
    int li = 0, lprev = -1;
    if (hasX[p]) {
      li = ROLE ? lab[L - 1 - i] : lab[i];
      if (i >= 1) lprev = ROLE ? lab[L - i] : lab[i - 1];
    }
    skip[p] = hasX[p] && i >= 1 && li != lprev;
    X[p] = kNeg;
    Y[p] = (i == 0) ? 0.f : kNeg;  // virtual frame "-1": all mass on the first blank
  }
  constexpr int kDir = ROLE ? -1 : 1;                 // direction of the pair index in memory
  const int lab0 = ROLE ? L - i0 : 1 + i0;            // index of El_{i0} inside a frame of E (pair p: lab0 + kDir*p)
  // Values are kept relative to a per-thread integer offset c (a float holding an
  // integer): true log2 value = stored + c.  A thread whose states are all still
  // unreachable simply adopts its neighbour's offset.
  float xin = kNeg;  // X_{i0-1} of the previous frame, already relative to c
  float c = 0.f;
  bool adopt = false;

  const int pitch2 = 2 * pitch;
  float *o = (ROLE ? d.beta : d.alpha) + um.ab_off + (ROLE ? (long long)(T - 1) * pitch2 : 0) +
             (ROLE ? 2 * (L - i0) : 2 * i0);          // pair p of this frame: o[2*kDir*p], o[2*kDir*p + 1]
  float *off_out = (ROLE ? d.offB : d.offA) + um.off_off + r;
  const int o_step = ROLE ? -pitch2 : pitch2, e_step = ROLE ? -pitch : pitch;
  const bool writes_off = r < nthreads_needed;
  const int off_stride = (nthreads_needed + 3) & ~3;
  const uint32_t bn_mine = smem_u32(bnd + w);         // this warp's boundary slot; the double buffer toggles by 256 B
  const bool sends = multi && lane == 31, receives = multi && lane == 0 && w > 0, first = r == 0;
  uint32_t par = 0;
  int st = 0, done = 0;
  uint32_t ph = 0;
  const int R = min(F, kRenorm);                      // frames per run (F is a power of two: runs tile the blocks)

  for (int k = 0; k < nchunks; k++) {
    const int n = min(F, T - k * F);
    mbar_wait(my_bar + st, ph);
    const float *e = my_stage + st * stage_floats + (ROLE ? (n - 1) * pitch : 0);
    for (int f0 = 0; f0 < n; f0 += R) {
      const int m = min(R, n - f0);
      for (int j = 0; j < m; j++) {
        // ---- hand (X_last, c) of the previous frame to the next thread
        float xv = __shfl_up_sync(0xffffffffu, X[P - 1], 1);
        float cv = __shfl_up_sync(0xffffffffu, c, 1);
        if (multi) {
          if (sends)
            asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(bn_mine + par), "f"(X[P - 1]), "f"(c) : "memory");
          named_bar_sync(1 + ROLE, nbar);
          if (receives)
            asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(xv), "=f"(cv) : "r"(bn_mine + par - 8) : "memory");
          par ^= 256u;
        }
        if (first) {
          xv = kNeg;
          cv = c;
        }
        if (adopt) c = cv;                   // nothing reachable here at the last block end: follow the neighbour
        adopt = false;
        xin = fmaxf(xv + (cv - c), kNeg);    // cv - c is an exact integer

        // ---- this frame
        const float Eb = e[0];
        float El[P];
#pragma unroll
        for (int p = 0; p < P; p++) El[p] = hasX[p] ? e[lab0 + kDir * p] : kNeg;  // kNeg keeps a missing X at "log 0"
        e += e_step;
        float xp = xin;
#pragma unroll
        for (int p = 0; p < P; p++) {
          const float l1 = lse2_pair(Y[p], xp);            // blank state: Y_i, X_{i-1}
          const float xo = X[p];
          X[p] = El[p] + lse2_pair(xo, skip[p] ? l1 : Y[p]);   // label state: X_i, Y_i (, X_{i-1})
          Y[p] = Eb + l1;
          xp = xo;
          if (hasY[p])
            *reinterpret_cast<float2 *>(o + 2 * kDir * p) = ROLE ? make_float2(X[p], Y[p]) : make_float2(Y[p], X[p]);
        }
        o += o_step;
      }
      done += m;
      // ---- end of a frame block: publish the offset the block was stored with, re-centre.  c only changes
      //      here (re-centred if the thread has reachable states) or, for a thread with nothing reachable, at
      //      the next exchange (adopted from the left neighbour).
      if ((done & (kRenorm - 1)) == 0 || done == T) {
        if (writes_off) *off_out = c;
        off_out += off_stride;
        float mx = kNeg;
#pragma unroll
        for (int p = 0; p < P; p++) mx = fmaxf(mx, fmaxf(hasX[p] ? X[p] : kNeg, hasY[p] ? Y[p] : kNeg));
        if (mx > -1.0e29f) {
          const float sh = floorf(mx);
#pragma unroll
          for (int p = 0; p < P; p++) {
            X[p] = fmaxf(X[p] - sh, kNeg);
            Y[p] = fmaxf(Y[p] - sh, kNeg);
          }
          c += sh;
        } else {
          adopt = true;
        }
      }
    }
    // stage fully consumed -> refill it.  Thread 0 is past the exchange barrier of the chunk's LAST frame, which
    // every thread reaches only after reading the frame before it; the last frame's own reads are ordered by the
    // barrier below (one per chunk, several warps only).
    if (multi) named_bar_sync(1 + ROLE, nbar);
    else __syncwarp();
    if (r == 0 && k + kStages < nchunks) issue(k + kStages, st);
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

// One CTA per (utterance, direction): alpha and beta share nothing, and as separate CTAs two of them fit an SM
// (registers, ~79 KB of stages), so the kernel's end is not tied to the one SM that holds the longest utterance's
// 20 warps.  CTA i of a group works on utterance order[b_lo + i/2] (longest first), direction i % 2.
template <int P>
__global__ void __launch_bounds__(512, 2) ctc_alpha_beta_kernel(CtcDev d, int frames_per_stage, int stage_floats) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *mbar = reinterpret_cast<uint64_t *>(smem_raw);               // [kStages]
  float2 *bnd = reinterpret_cast<float2 *>(smem_raw + 64);               // [2][32] (value, offset)
  float *fin = reinterpret_cast<float *>(smem_raw + 64 + 1024);          // [4]
  float *stages = reinterpret_cast<float *>(smem_raw + 2048);            // [kStages][stage_floats]

  const int role = blockIdx.x & 1;
  const int b = d.order[d.b_lo + (blockIdx.x >> 1)];
  const UttMeta um = d.meta[b];
  if (!um.feasible) return;
  const int r = threadIdx.x;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; i++) mbar_init(mbar + i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int nthreads_needed = (um.L + 1 + P - 1) / P;
  const int nwarps_active = (nthreads_needed + 31) >> 5;
  if ((r >> 5) >= nwarps_active) return;  // idle warps leave; named barriers count the rest
  const int nbar = nwarps_active * 32;
  // FULL: every pair of every lane of this warp lies inside the lattice and has a label state (i < L)
  const bool full = ((r >> 5) + 1) * 32 * P <= um.L;
  if (role == 0) {
    if (full) ctc_ab_run<P, 0, true>(d, um, b, r, nthreads_needed, nbar, mbar, stages, stage_floats, bnd, fin, frames_per_stage);
    else ctc_ab_run<P, 0, false>(d, um, b, r, nthreads_needed, nbar, mbar, stages, stage_floats, bnd, fin, frames_per_stage);
  } else {
    if (full) ctc_ab_run<P, 1, true>(d, um, b, r, nthreads_needed, nbar, mbar, stages, stage_floats, bnd, fin, frames_per_stage);
    else ctc_ab_run<P, 1, false>(d, um, b, r, nthreads_needed, nbar, mbar, stages, stage_floats, bnd, fin, frames_per_stage);
  }
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
  long long zrows;       // rows of this group that only need zeros: padded frames, every frame of an infeasible utterance
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
  // one row of zeros: the producer stores it over the padded rows BETWEEN the gradient rows, so that these writes
  // ride along with the read-heavy main phase instead of forming a write-only tail (3 GB at configs[4])
  float *zrow = reinterpret_cast<float *>(rsm + (((size_t)(reinterpret_cast<unsigned char *>(s_ul + rc.pitch_max) - rsm) + 127) & ~(size_t)127));

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
  for (int k = tid; k < (A >> 2); k += kRingThreads) reinterpret_cast<float4 *>(zrow)[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy zeros -> visible to the bulk stores
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
    if (lane == 0) {
      // ---- this CTA's share of the zero rows: (zb, zt) walks the padded frames t in [T_b, Tmax) of feasible
      //      utterances and every frame of infeasible ones, in utterance order
      const long long z_lo = rc.zrows * blockIdx.x / gridDim.x, z_hi = rc.zrows * (blockIdx.x + 1) / gridDim.x;
      long long zleft = z_hi - z_lo;
      int zb = d.b_lo, zt = 0;
      if (zleft > 0) {
        long long skip = z_lo;
        for (; zb < b_end; zb++) {
          const UttMeta um = d.meta[zb];
          const int first = um.feasible ? um.T : 0;
          const long long cnt = d.Tmax - first;
          if (skip < cnt) {
            zt = first + (int)skip;
            break;
          }
          skip -= cnt;
        }
      }
      auto store_zero_row = [&]() {   // (joins the bulk group of the caller's next commit)
        bulk_store(d.grad + ((long long)zt * d.B + zb) * A, zrow, (uint32_t)A * 4u);
        zleft--;
        if (++zt >= d.Tmax && zleft > 0) {
          for (zb++; zb < b_end; zb++) {
            const UttMeta um = d.meta[zb];
            zt = um.feasible ? um.T : 0;
            if (zt < d.Tmax) break;
          }
        }
      };
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
        if (zleft > 0) store_zero_row();   // same bulk group: the slot accounting below is unchanged
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        cur_next(d, ps, b_end, pshift);
        // the slot of row i-1 (its store was committed one row ago) takes row i+NA-1
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        if (i + rc.NA - 1 < nrows) issue_act(i + rc.NA - 1);
      }
      // zero rows that outnumber this CTA's gradient rows (heavily padded batches)
      for (int n = 0; zleft > 0; n++) {
        store_zero_row();
        if ((n & 7) == 7) {
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
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

}

// ===========================================================================
// host side
// ===========================================================================
struct Plan {
  int A, B, Tmax, maxL, pitch_max;
  long long sumT, sumL;
  size_t off_meta, off_labels, off_order, off_uniq_lab, off_uniq_start, off_pos, off_nuniq;  // header block
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
  p->off_labels = o;      o = align_up(o + sizeof(int) * (p->sumL + 1), 256);
  p->off_order = o;       o = align_up(o + sizeof(int) * B, 256);   // (meta, labels, order: what K1 / K2 need)
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

// Tuning switches: read from the environment ONCE per process (none is needed in production); tests and the
// tools under tools/ change them in-process through b200ctc_set_tuning().
struct Tuning {
  int force_p, ring, na, nt, profile, groups, one_stream;
};
Tuning &tuning() {
  static Tuning t = [] {
    auto geti = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
    Tuning v;
    v.force_p = geti("B200CTC_P", 0);
    v.ring = geti("B200CTC_RING", -1);
    v.na = geti("B200CTC_NA", 0);
    v.nt = geti("B200CTC_NT", 0);
    v.profile = geti("B200CTC_PROFILE", 0);
    v.groups = geti("B200CTC_GROUPS", 0);
    v.one_stream = geti("B200CTC_ONE_STREAM", 0);
    return v;
  }();
  return t;
}
int *tuning_field(const char *key) {
  Tuning &t = tuning();
  struct { const char *k; int *v; } tab[] = {
      {"P", &t.force_p}, {"RING", &t.ring}, {"NA", &t.na}, {"NT", &t.nt}, {"PROFILE", &t.profile},
      {"GROUPS", &t.groups}, {"ONE_STREAM", &t.one_stream}};
  for (auto &e : tab)
    if (key && strcmp(key, e.k) == 0) return e.v;
  return nullptr;
}

// Everything that belongs to ONE device: function attributes (cudaFuncSetAttribute is per device), the
// side streams and events of the grouped launch, the SM count and the staging event.  Keyed by
// cudaGetDevice() so that an in-process multi-GPU caller of the C ABI gets a consistent set per GPU.
constexpr int kMaxGroups = 8;
struct DeviceState {
  int num_sms = 0;
  size_t k2_smem[3] = {0, 0, 0};   // P = 1, 2, 4: dynamic shared memory already granted to the kernel
  size_t smem3_set = 0, ring_set = 0;
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

template <int P>
cudaError_t launch_k2(const CtcDev &dev, int B, int NT, int F, int stage_floats, cudaStream_t stream, DeviceState *ds) {
  const size_t smem = 2048 + sizeof(float) * kStages * (size_t)stage_floats;
  size_t &granted = ds->k2_smem[P == 1 ? 0 : (P == 2 ? 1 : 2)];
  if (smem > granted) {
    cudaError_t e = cudaFuncSetAttribute(ctc_alpha_beta_kernel<P>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    granted = smem;
  }
  ctc_alpha_beta_kernel<P><<<2 * dev.nb, NT, smem, stream>>>(dev, F, stage_floats);
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
  const Tuning tune = tuning();
  if (!ensure_pinned(g_stage, p.header_bytes, (size_t)B)) return CTC_STATUS_MEMOPS_FAILED;
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
  // K2 runs one CTA per utterance and its CTAs differ 20-fold in work (frames x lattice states): inside every
  // utterance group they are handed out longest first, so that the long recursions start at once and the short
  // ones fill the SMs behind them (256 utterances on 148 SMs in input order: the kernel ended 1.6x later than its
  // longest CTA needs)
  {
    int *ho = reinterpret_cast<int *>(h + p.off_order);
    for (int gi = 0; gi < ngroups; gi++) {
      const int lo = (int)((long long)B * gi / ngroups), hi = (int)((long long)B * (gi + 1) / ngroups);
      for (int b = lo; b < hi; b++) ho[b] = b;
      std::stable_sort(ho + lo, ho + hi, [&](int x, int y) {
        return (long long)p.meta[x].T * (p.meta[x].L + 1) > (long long)p.meta[y].T * (p.meta[y].L + 1);
      });
    }
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
  dev.order = reinterpret_cast<const int *>(w + p.off_order);
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
  // emission stages of K2: 8 frames each when that fits (a stage then holds one re-centring block), a power of
  // two in any case; kStages stages per CTA (one direction of one utterance), at most 96 KB: two CTAs per SM
  const int stage_floats = std::min(std::max(kStageFloats, kRenorm * p.pitch_max), 6144);
  int F = 1;
  while (F < 32 && 2 * F * p.pitch_max <= stage_floats) F *= 2;
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
  rc.vrow_base = rc.vrows = rc.zrows = 0;
  auto ring_bytes = [&]() {
    return (size_t)512 + (size_t)rc.NA * rc.act_bytes + (size_t)rc.NT * rc.tab_bytes +
           sizeof(float) * ((size_t)5 * p.pitch_max + 8) + 128 + (size_t)rc.act_bytes;   // (+ the row of zeros)
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
    cudaError_t ce = P == 1   ? launch_k2<1>(dev, B, NT, F, stage_floats, st, ds)
                     : P == 2 ? launch_k2<2>(dev, B, NT, F, stage_floats, st, ds)
                              : launch_k2<4>(dev, B, NT, F, stage_floats, st, ds);
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
      rc.zrows = rows - rc.vrows;
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

int b200ctc_set_tuning(const char *key, int value) {
  std::lock_guard<std::mutex> lock(g_mu);
  int *f = tuning_field(key);
  if (!f) return -1;
  *f = value;
  return 0;
}

int b200ctc_launches_per_call(int with_gradients) { return with_gradients ? 3 : 2; }

}  // extern "C"
