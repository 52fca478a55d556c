// kaldi_ctc_b200/csrc/rnn_rec_tc.cu -- persistent recurrent kernels on tcgen05.
//
// One thread-block cluster per (direction, batch chunk).  CTA c of the cluster owns
// the hidden units [32c, 32c+32): its 128 gate rows (row = 4*unit + gate; unused
// gate slots of GRU / plain RNN are zero rows) of the recurrent matrix R stay ON
// CHIP as BF16 for all T steps -- in TENSOR MEMORY, as the A operand of
// tcgen05.mma (lane = gate row, H/2 columns), so a step does not re-read 80 KB of
// weights through the shared-memory port (measured: 1540 -> ~300 cycles of MMA issue
// per step).  Per time step:
//   MMA warp     waits until every CTA's slice of h_{t-1} has landed in the local
//                (double-buffered, swizzled) h tile, then one elected lane issues
//                H/16 tcgen05.mma (M=128, N=16, K=16, kind::f16/BF16) accumulating
//                the [128 gate rows x 16 utterances] pre-activations in TMEM;
//                tcgen05.commit signals the epilogue.
//   epilogue     4 warps: tcgen05.ld their 32 TMEM lanes, add the hoisted input
//                projection (prefetched from HBM one step ahead), apply the gate
//                non-linearity, 4x4 quad transposes by warp shuffle so that one
//                thread holds all gates of a (unit, utterance), fp32 cell update
//                (cell / hidden state live in registers), pack h_t to BF16 and
//                st.async the 16-byte chunks straight into every CTA's next
//                (swizzled) h tile through distributed shared memory -- each
//                store completes bytes on the destination's "h tile full"
//                mbarrier, so there is no fence, no barrier and no arrive on the
//                critical path; y / gates / cell go to HBM afterwards.
// No cluster-wide barrier inside the loop.
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "rnn_common.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace b200 {
bool make_map_nd(CUtensorMap *map, const float *base, int rank, const long long *dims, const long long *strides,
                 const int *box);   // rnn_gemm_tc.cu
namespace {

using namespace tc;

// TMA descriptors of the per-step operands (batch chunks of 8 / 16 only, see kBulk in the kernels)
struct RecMaps {
  CUtensorMap gates[2];   // [T*B rows][G gates][H]   box [BC][G][32]
  CUtensorMap cell[2];    // [T*B rows][H]            box [BC][32]
  CUtensorMap dy, y;      // [T*B rows][H*dirs]       box [BC][32]
};
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

constexpr int UT = 32;          // hidden units per CTA
constexpr int NPAD = 16;        // MMA N (utterances per chunk, zero padded)
constexpr int kIssuers = 4;     // MMA-issuing warps: the burst of small MMAs is bound by the ~13 instructions ptxas
                                // emits around every tcgen05.mma (tools/mma_bench.cu: 20 MMAs in 1135 / 855 / 480 cycles
                                // with 1 / 2 / 4 issuing warps), so four warps issue a quarter of the K slices each
constexpr int kThreads = 32 * kIssuers + 128;   // warps 0-3: MMA issuers (one accumulator each), warps 4-7: epilogue
// EH = 2 ("split epilogue", batch chunks of 8 / 16): warps 8-11 are a second set of epilogue warps.  A TMEM lane
// quarter can only be read by warps with the same (warp % 4), so warps q and q+4 share a quarter and each takes
// half of the chunk's utterance columns: the per-thread epilogue work -- which is on the per-step critical
// chain -- halves.
__host__ __device__ constexpr int threads_of(int EH) { return 32 * kIssuers + 128 * EH; }
constexpr int kTmemCols = 256;  // D: columns [0,64) (four accumulators); A (R slice): columns [64, 64 + H/2)
constexpr int kRingOffset = 48 * 1024;   // forward kernel: cp.async prefetch ring (32 KB) behind the h tiles + barriers
// In-kernel phase counters (cycles per step of cluster 0 / CTA 0) are a build-time option
// (-DB200RNN_PHASE_COUNTERS): release kernels carry no instrumentation in the per-step loops.
#ifdef B200RNN_PHASE_COUNTERS
constexpr bool kPhaseCounters = true;
#else
constexpr bool kPhaseCounters = false;
#endif
constexpr int kACol = 64;       // (several independent accumulators were measured: no gain, the
                                //  burst is issue-bound at ~24 cycles per MMA, tools/mma_bench.cu)

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_cluster_v4f(uint32_t raddr, float4 v) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// 4x4 transpose inside a lane quad: in x[j] = my gate's value for batch j of the group;
// out g[k] = gate k's value for batch (lane & 3).
__device__ __forceinline__ void quad_transpose(const float (&x)[4], float (&g)[4], int s) {
  const bool o1 = s & 1, o2 = s & 2;
  // stage 1 (xor 1): keep the batches with my parity, get the partner's gate for them
  const float k0 = o1 ? x[1] : x[0], k1 = o1 ? x[3] : x[2];          // my gate, batches (s&1), (s&1)+2
  const float r0 = __shfl_xor_sync(0xffffffffu, o1 ? x[0] : x[1], 1);  // partner gate, batch (s&1)
  const float r1 = __shfl_xor_sync(0xffffffffu, o1 ? x[2] : x[3], 1);  // partner gate, batch (s&1)+2
  // stage 2 (xor 2): keep batch s, get the other pair of gates for it
  const float w0 = o2 ? k1 : k0;                                       // gate s      , batch s
  const float w1 = o2 ? r1 : r0;                                       // gate s^1    , batch s
  const float w2 = __shfl_xor_sync(0xffffffffu, o2 ? k0 : k1, 2);      // gate s^2    , batch s
  const float w3 = __shfl_xor_sync(0xffffffffu, o2 ? r0 : r1, 2);      // gate s^3    , batch s
  // w[j] = gate (j ^ s); gate k = w[k ^ s]
  const float p0 = o1 ? w1 : w0, p1 = o1 ? w0 : w1, p2 = o1 ? w3 : w2, p3 = o1 ? w2 : w3;
  g[0] = o2 ? p2 : p0;
  g[1] = o2 ? p3 : p1;
  g[2] = o2 ? p0 : p2;
  g[3] = o2 ? p1 : p3;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t *>(&v);
}

// ===========================================================================
// forward
// ===========================================================================
// smem: [Rs: nkb x 128 rows x 128 B][hs: 2 x nkb x 16 rows x 128 B][barriers]
template <int MODE, int NJ, int NKB, int EH>
__global__ void __launch_bounds__(threads_of(EH), 1)
rec_tc_fwd_kernel(RecArgs a, const __grid_constant__ RecMaps tm) {
  constexpr int G = MODE == 2 ? 4 : (MODE == 3 ? 3 : 1);
  constexpr int BC = 4 * NJ;
  constexpr int NJL = NJ / EH, BCL = 4 * NJL;   // utterance groups / columns per epilogue thread
  constexpr int kThreads = threads_of(EH);
  static_assert(NJ % EH == 0, "split epilogue needs an even number of utterance groups");
  // 16 utterances: the h tile is laid out in 64-byte-swizzle k-blocks of 32 units, so that one CTA's slice
  // ([16 rows][64 B] = 1 KB) is contiguous in every peer's tile and travels as ONE bulk DSMEM copy per peer
  // instead of 1024 st.async per step and CTA (one mbarrier update each at the receiver: 788 cycles per step)
  constexpr bool kSW64 = BC >= 16;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int NC = a.NC;
  const int dir = blockIdx.y % a.dirs, chunk = blockIdx.y / a.dirs;
  const int b_lo = chunk * BC, nb = min(BC, a.B - b_lo);
  const int H = a.H, T = a.T, B = a.B, GH = G * H, HO = H * a.dirs;
  const int nkb = NKB > 0 ? NKB : H / 64;  // NKB > 0: compile-time K extent, MMA loop fully unrolled
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ uint8_t smem_dyn[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t *hs = smem;                                 // two buffers of nkb * 2048 B
  const int hs_bytes = nkb * 2048;
  uint64_t *hfull = reinterpret_cast<uint64_t *>(hs + 2 * hs_bytes);  // [2]
  uint64_t *acc_full = hfull + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

  // bytes every CTA's h tile receives per step: NC slices of (BC rows x 32 units x 2 B)
  const uint32_t h_bytes = (uint32_t)NC * BC * 64u;
  for (int idx = tid; idx < 2 * hs_bytes / 16; idx += kThreads)
    reinterpret_cast<uint4 *>(hs)[idx] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(hfull + 0, 1);
    mbar_init(hfull + 1, 1);
    mbar_init(acc_full, kIssuers);  // one commit from each issuing warp
    fence_barrier_init();
    // arm both h tiles for their first fill (st.async completes bytes on them)
    mbar_expect_tx(hfull + 0, h_bytes);
    mbar_expect_tx(hfull + 1, h_bytes);
  }
  const uint32_t tmem_cols = kACol + H / 2 <= kTmemCols ? (uint32_t)kTmemCols : 512u;
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  fence_proxy_async_all();  // zeroed h tiles -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // ---- one-time: my 128 gate rows of R -> BF16 -> tensor memory (lane = row, column = k/2)
  if (warp >= kIssuers && warp < kIssuers + 4) {
    const int q = warp & 3, row = q * 32 + lane;
    const int u = row >> 2, g = row & 3;
    const float *src = a.w_rec[dir] + ((size_t)(g < G ? g : 0) * H + crank * UT + u) * H;
    for (int c0 = 0; c0 < H / 2; c0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const float2 f = *reinterpret_cast<const float2 *>(src + 2 * (c0 + i));
        v[i] = g < G ? pack_bf16(f.x, f.y) : 0u;
      }
      tmem_st_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + kACol + c0, v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster.sync();  // every CTA's tiles and barriers exist before anyone writes remotely

  if (warp < kIssuers) {
    // ===================== MMA issuers =====================
    // The burst of H/16 small MMAs is issue-bound (~30 cycles each), so two warps issue half of
    // the K range each into their own accumulator (columns [16*warp, 16*warp+16)); the epilogue
    // adds the two.  Warp 1 also re-arms the h-tile barrier.
    constexpr uint32_t idesc = instr_desc(kFmtBF16, 0, 0, 128, NPAD);
    // warp-uniform copies (shuffle from lane 0) so the compiler keeps MMA operands in uniform registers
    const uint32_t hs0 = __shfl_sync(0xffffffffu, smem_u32(hs), 0);
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, tmem_base, 0);
    const bool prof = kPhaseCounters && a.dbg != nullptr && crank == 0 && blockIdx.y == 0 && warp == 1;
    long long pm[4] = {0, 0, 0, 0};
    for (int step = 0; step < T; step++) {
      const int p = step & 1;
      const long long m0 = prof ? clock64() : 0;
      if (step > 0) {
        const int use = p ? (step - 1) >> 1 : (step >> 1) - 1;
        mbar_wait(hfull + p, use & 1);
        // re-arm for the fill two steps ahead (both issuers are past the wait before any such
        // data can exist: it needs this step's h from every CTA first)
        if (warp == 1 && elect_one()) mbar_expect_tx(hfull + p, h_bytes);
      }
      const long long m1 = prof ? clock64() : 0;
      tc_fence_after();
      const long long m2 = prof ? clock64() : 0;
      {
        // warp-uniform issue: every lane computes the (uniform) descriptors, one elected lane
        // issues -- keeps the operands in uniform registers (no per-MMA R2UR/ELECT loop)
        // descriptor of the first K slice; later slices only add to the 14-bit address field
        const uint64_t bd0 = smem_desc(hs0 + p * hs_bytes, 0, kSW64 ? 512 : 1024, kSW64 ? kLayoutSw64 : kLayoutSw128);
        const uint32_t dacc = tmem_d + warp * NPAD;
        if (NKB > 0) {
#pragma unroll
          for (int i = 0; i < NKB; i++) {
            const int kk = kIssuers * i + warp;  // K slices interleaved over the issuing warps
            const uint64_t bd = bd0 + (uint64_t)((kSW64 ? (kk >> 1) * 1024 + (kk & 1) * 32
                                                        : (kk >> 2) * 2048 + (kk & 3) * 32) >> 4);
            if (elect_one()) mma_bf16_ts(dacc, tmem_d + kACol + kk * 8, bd, idesc, i ? 1u : 0u);
          }
        } else {
          for (int i = 0; i < nkb; i++) {
            const int kk = kIssuers * i + warp;
            const uint64_t bd = bd0 + (uint64_t)((kSW64 ? (kk >> 1) * 1024 + (kk & 1) * 32
                                                        : (kk >> 2) * 2048 + (kk & 3) * 32) >> 4);
            if (elect_one()) mma_bf16_ts(dacc, tmem_d + kACol + kk * 8, bd, idesc, i ? 1u : 0u);
          }
        }
        if (elect_one()) tc_commit(acc_full);
      }
      __syncwarp();
      const long long m3 = prof ? clock64() : 0;
      if (warp == 1) {
        // hand the accumulator to the epilogue through a named barrier: the issuing warp sees the
        // tcgen05.commit arrive ~60 cycles after its last MMA, warps sleeping on the mbarrier were
        // measured to resume ~300 cycles later
        mbar_wait(acc_full, step & 1);
        asm volatile("bar.arrive 2, %0;" ::"n"(32 + 128 * EH) : "memory");
      }
      if (prof) {
        const long long m4 = clock64();
        pm[0] += m1 - m0; pm[1] += m2 - m1; pm[2] += m3 - m2; pm[3] += m4 - m3;
        if (lane == 0) { a.dbg[12] += m4; a.dbg[14] += m1; }  // absolute SM clocks: completion / h-arrival seen
      }
    }
    if (prof && lane == 0) { a.dbg[8] = pm[0]; a.dbg[9] = pm[1]; a.dbg[10] = pm[2]; a.dbg[11] = pm[3]; }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;                  // TMEM lane quarter
    const int eh = (warp - kIssuers) >> 2;   // which half of the chunk's columns (split epilogue)
    const int jb = eh * NJL;                 // my first utterance group
    const int s = lane & 3;                  // gate slot of my row / batch slot after the transpose
    const int ul = q * 8 + (lane >> 2);      // local unit 0..31
    const int unit = crank * UT + ul;
    float *gates = a.gates[dir];
    float *cell = a.cell[dir];
    float brn = 0.f;
    if (MODE == 3) brn = a.b_rec[dir][2 * H + unit];
    float cst[NJL], hst[NJL];                // cell / hidden state of (unit, batch 4(jb+j)+s), fp32
#pragma unroll
    for (int j = 0; j < NJL; j++) cst[j] = hst[j] = 0.f;

    // which pre-activation column this ROW needs: LSTM gate s; GRU slots 0,1 -> r,z,
    // slot 3 carries the input part of n (slot 2 = recurrent part, nothing to load)
    const int pcol = MODE == 3 ? (s == 3 ? 2 : s) : s;
    const bool pload = MODE == 2 ? true : (MODE == 3 ? (s != 2) : (s == 0));
    // The hoisted projection rows are prefetched kPF steps ahead with cp.async into a per-thread ring in
    // shared memory (each thread only ever reads what it copied itself, so cp.async.wait_group is the only
    // synchronisation).  Register prefetch one step ahead cost 12 % of the step at T = 2000 (HBM latency
    // tails of the slowest thread of the slowest CTA gate every step); 2 / 4 steps ahead recovered 11 / 15 %.
    // Batch chunks of 8 / 16 (kBulk): 4-byte cp.async copies (2048 per step and CTA at 16 utterances) were
    // 0.7 us of a 2.1 us step (LSU issue), and one 128-byte bulk copy per (utterance, gate) was worse still
    // (64 small TMA operations per step).  There the whole [utterances x gates x my 32 units] box of a step
    // comes with ONE 3-D TMA tile load, completion counted on one mbarrier per ring slot.  At 4 utterances the
    // same tile load (2 KB boxes, 16 slots) replaces the cp.async ring as well: 0.772 -> 0.750 us per step
    // (kBulk = false keeps the cp.async variant for comparison).
    constexpr bool kBulk = true;
    constexpr int kPF = BC == 4 ? 16 : (BC == 8 ? 6 : 3);
    constexpr int kSlotFloats = BC * G * 32;                        // kBulk slot [utterance][gate][32]: 8 KB at 16 x 4
    float *bring = reinterpret_cast<float *>(smem + kRingOffset);
    uint64_t *pfbar = reinterpret_cast<uint64_t *>(smem + kRingOffset - 128);   // [kPF <= 16], kBulk only
    auto issue_bulk = [&](int step) {
      if (step < T && tid == 32 * kIssuers) {
        const int t = dir ? T - 1 - step : step;
        uint64_t *bar = pfbar + step % kPF;
        mbar_expect_tx(bar, (uint32_t)kSlotFloats * 4u);
        tma_load_3d(bring + (size_t)(step % kPF) * kSlotFloats, &tm.gates[dir], crank * UT, 0, t * B + b_lo, bar);
      }
    };
    // [slot][quarter thread 0..127][BC columns]; a thread owns columns [eh * BCL, eh * BCL + BCL)
    float *pring = reinterpret_cast<float *>(smem + kRingOffset) + (size_t)((tid - 32 * kIssuers) & 127) * BC + eh * BCL;
    auto issue_pre = [&](int step) {
      if (kBulk) {
        issue_bulk(step);
        return;
      }
      if (step < T && pload) {
        const int t = dir ? T - 1 - step : step;
        const float *pp = gates + ((size_t)t * B + b_lo) * GH + (size_t)pcol * H + unit;
        const uint32_t dst = smem_u32(pring + (size_t)(step % kPF) * 128 * BC);
#pragma unroll
        for (int b = 0; b < BCL; b++)
          if (eh * BCL + b < nb)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + 4 * b), "l"(pp + (size_t)(eh * BCL + b) * GH) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (kBulk) {
      if (tid == 32 * kIssuers) {
        for (int k = 0; k < kPF; k++) mbar_init(pfbar + k, 1);
        fence_barrier_init();
      }
      asm volatile("bar.sync 3, %0;" ::"n"(128 * EH) : "memory");
    } else {
#pragma unroll
      for (int k = 0; k < kPF; k++) {   // slots of (thread, b) pairs that are never copied stay zero
#pragma unroll
        for (int b = 0; b < BCL; b++) pring[(size_t)k * 128 * BC + b] = 0.f;
      }
    }
    for (int k = 0; k < kPF; k++) issue_pre(k);

    // remote addresses that never change
    const int kb_mine = crank >> 1, chunk_mine = (crank & 1) * 4 + q;
    const int peer_a = lane >> 2, peer_b = (lane >> 2) + 8;  // the two peers this lane serves
    // cluster addresses of the peers' h tiles / barriers (the offset inside a CTA is the same everywhere)
    const uint32_t rhs_a = mapa_u32(smem_u32(hs), peer_a < NC ? peer_a : 0);
    const uint32_t rhs_b = mapa_u32(smem_u32(hs), peer_b < NC ? peer_b : 0);
    const uint32_t rhf_a = mapa_u32(smem_u32(hfull), peer_a < NC ? peer_a : 0);
    const uint32_t rhf_b = mapa_u32(smem_u32(hfull), peer_b < NC ? peer_b : 0);
    const uint32_t rhs_l = mapa_u32(smem_u32(hs), lane < NC ? lane : 0);       // kSW64: lane i -> CTA i
    const uint32_t rhf_l = mapa_u32(smem_u32(hfull), lane < NC ? lane : 0);
    uint8_t *hstage = smem + kRingOffset + 32 * 1024;                          // kSW64: [2][16 rows][64 B]

    const bool prof = kPhaseCounters && a.dbg != nullptr && crank == 0 && blockIdx.y == 0 && warp == kIssuers;
    long long pe[7] = {0, 0, 0, 0, 0, 0, 0};
    const long long loop_t0 = (kPhaseCounters && a.dbg) ? clock64() : 0;
    for (int step = 0; step < T; step++) {
      const int t = dir ? T - 1 - step : step;
      const long long c0 = prof ? clock64() : 0;
      // this step's projection rows (copied kPF steps ago)
      float pre[BCL];
      if (kBulk) {
        mbar_wait(pfbar + step % kPF, (uint32_t)(step / kPF) & 1u);
        const float *ps = bring + (size_t)(step % kPF) * kSlotFloats + (eh * BCL * G + pcol) * 32 + ul;
#pragma unroll
        for (int b = 0; b < BCL; b++) pre[b] = pload ? ps[b * G * 32] : 0.f;
      } else {
        asm volatile("cp.async.wait_group %0;" ::"n"(kPF - 1) : "memory");
        const float *ps = pring + (size_t)(step % kPF) * 128 * BC;
#pragma unroll
        for (int b = 0; b < BCL; b++) pre[b] = ps[b];
      }
      asm volatile("bar.sync 2, %0;" ::"n"(32 + 128 * EH) : "memory");
      const long long c1a = prof ? clock64() : 0;
      tc_fence_after();
      const long long c1 = prof ? clock64() : 0;
      if (prof && lane == 0) a.dbg[16] += c1 - c1a;
      // only the BC columns of each issuer's accumulator that this chunk uses (TMEM reads are paced by bytes)
      uint32_t ra[kIssuers][BCL];
#pragma unroll
      for (int w = 0; w < kIssuers; w++)
        tmem_ld_32xN<BCL>(tmem_base + ((uint32_t)(q * 32) << 16) + w * NPAD + eh * BCL, ra[w]);
      tmem_ld_wait();
      tc_fence_before();
      float r[BCL];  // the issuers' partial sums, added in a fixed order
#pragma unroll
      for (int e = 0; e < BCL; e++)
        r[e] = (__uint_as_float(ra[0][e]) + __uint_as_float(ra[1][e])) + (__uint_as_float(ra[2][e]) + __uint_as_float(ra[3][e]));
      const long long c2 = prof ? clock64() : 0;

      // ---- gates -> (unit, batch) threads, cell update; results kept in registers
      float hnew[NJL], sv[NJL][4], sc[NJL];
#pragma unroll
      for (int j = 0; j < NJL; j++) {
        float x[4], g[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float v = r[4 * j + e] + pre[4 * j + e];
          if (MODE == 2) {
            const float scl = s == 2 ? 1.f : 0.5f;
            const float th = tanh_fast(scl * v);
            x[e] = s == 2 ? th : fmaf(0.5f, th, 0.5f);
          } else if (MODE == 3) {
            const float th = tanh_fast(0.5f * v);
            x[e] = s < 2 ? fmaf(0.5f, th, 0.5f) : (s == 2 ? v + brn : v);
          } else {
            x[e] = MODE == 0 ? fmaxf(v, 0.f) : tanh_fast(v);
          }
        }
        quad_transpose(x, g, s);
        const bool valid = 4 * (jb + j) + s < nb;
        float h;
        if (MODE == 2) {
          cst[j] = fmaf(g[1], cst[j], g[0] * g[2]);
          h = g[3] * tanh_fast(cst[j]);
          sv[j][0] = g[0]; sv[j][1] = g[1]; sv[j][2] = g[2]; sv[j][3] = g[3];
          sc[j] = cst[j];
        } else if (MODE == 3) {
          const float n = tanh_fast(fmaf(g[0], g[2], g[3]));
          h = fmaf(g[1], hst[j] - n, n);  // (1-z) n + z h_prev
          sv[j][0] = g[0]; sv[j][1] = g[1]; sv[j][2] = n; sv[j][3] = 0.f;
          sc[j] = g[2];
        } else {
          h = g[0];
          sv[j][0] = h; sv[j][1] = sv[j][2] = sv[j][3] = 0.f;
          sc[j] = 0.f;
        }
        if (!valid) h = 0.f;
        hst[j] = h;
        hnew[j] = h;
      }

      const long long c3 = prof ? clock64() : 0;
      long long c4 = c3, c5 = c3, c6 = c3;
      // ---- critical path first: ship h_t to every CTA, then signal
      if (step + 1 < T) {
        const int pn = (step + 1) & 1;
#pragma unroll
        for (int j = 0; j < NJL; j++) {
          // pack 8 consecutive units (one warp) of batch 4(jb+j)+s into one 16-byte chunk
          const float v = hnew[j];
          const float pv = __shfl_xor_sync(0xffffffffu, v, 4);
          const uint32_t pair = (lane & 4) ? pack_bf16(pv, v) : pack_bf16(v, pv);
          const uint32_t pq = __shfl_xor_sync(0xffffffffu, pair, 8);
          const uint32_t lo = (lane & 8) ? pq : pair, hi = (lane & 8) ? pair : pq;
          const uint32_t lo2 = __shfl_xor_sync(0xffffffffu, lo, 16);
          const uint32_t hi2 = __shfl_xor_sync(0xffffffffu, hi, 16);
          uint4 ch;
          if (lane & 16) ch = make_uint4(lo2, hi2, lo, hi);
          else ch = make_uint4(lo, hi, lo2, hi2);
          const int b = 4 * (jb + j) + s;
          if (kSW64) {   // my slice of the tile, in the peers' (64-byte swizzle) layout, into local staging
            if ((lane >> 2) == 0)
              *reinterpret_cast<uint4 *>(hstage + pn * 1024 + b * 64 + ((q ^ ((b >> 1) & 3)) << 4)) = ch;
          } else {
            const uint32_t off = pn * hs_bytes + kb_mine * 2048 + b * 128 + ((chunk_mine ^ (b & 7)) << 4);
            if (peer_a < NC) st_async_v4(rhs_a + off, ch.x, ch.y, ch.z, ch.w, rhf_a + pn * 8);
            if (peer_b < NC) st_async_v4(rhs_b + off, ch.x, ch.y, ch.z, ch.w, rhf_b + pn * 8);
          }
        }
        if (kSW64) {
          fence_proxy_async();
          const bool sender = warp == kIssuers && lane < NC;   // lane i ships the slice to CTA i
          // the copy of the previous step has read the other buffer; it (and all older ones) is done before
          // anyone passes the barrier, i.e. before the buffer it used is written again at the next step
          if (sender) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          asm volatile("bar.sync 3, %0;" ::"n"(128 * EH) : "memory");
          if (sender) {
            asm volatile(
                "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                    rhs_l + (uint32_t)(pn * hs_bytes + crank * 1024)),
                "r"(smem_u32(hstage + pn * 1024)), "n"(BC * 64), "r"(rhf_l + pn * 8)
                : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        if (prof) c4 = c5 = c6 = clock64();
      }

      // ---- off the critical path: results to HBM, next step's projection prefetch
#pragma unroll
      for (int j = 0; j < NJL; j++) {
        const int b = 4 * (jb + j) + s;
        if (b < nb) {
          const size_t row = (size_t)t * B + b_lo + b;
          a.y[row * HO + dir * H + unit] = hnew[j];
          if (a.save) {
            float *gp = gates + row * GH + unit;
            gp[0] = sv[j][0];
            if (G > 1) { gp[H] = sv[j][1]; gp[2 * H] = sv[j][2]; }
            if (G > 3) gp[3 * H] = sv[j][3];
            if (MODE >= 2) cell[row * H + unit] = sc[j];
          }
        }
      }
      issue_pre(step + kPF);   // refills the slot this step has just consumed
      if (prof) {
        const long long c7 = clock64();
        if (lane == 0) { a.dbg[13] += c1; a.dbg[15] += c4; }      // accumulator seen by the epilogue / h sent
        pe[0] += c1 - c0; pe[1] += c2 - c1; pe[2] += c3 - c2; pe[3] += c4 - c3;
        pe[4] += c5 - c4; pe[5] += c6 - c5; pe[6] += c7 - c6;
      }
    }
    if (prof && lane == 0)
      for (int i = 0; i < 7; i++) a.dbg[i] = pe[i];
    if (kPhaseCounters && a.dbg && warp == kIssuers && lane == 0 && blockIdx.y < 16) a.dbg[32 + blockIdx.y + 16 * (crank != 0)] = clock64() - loop_t0;
  }
  tc_fence_before();
  __syncthreads();
  cluster.sync();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

// ===========================================================================
// backward (data): gate gradients for all steps + the dh recurrence
// ===========================================================================
// CTA c keeps R_slice^T (its 128 gate rows, all H columns) in tensor memory as the A
// operand: M = hidden index k (H padded to MT tiles of 128 lanes), K = own gate row r.
// Per step:  epilogue threads (unit, utterance) form dh = dy + sum of the partial
// products received from every CTA, compute the gate gradients (fp32), write them
// to HBM (for the dW/dx GEMMs) and as BF16 into the swizzled [utterance][gate row]
// B tile; the MMA warp issues MT*8 tcgen05.mma giving the PARTIAL dh_{prev}[k][b]
// of this CTA's rows for ALL k; each epilogue warp then st.async's its 32 lanes
// (= the 32 units of exactly one peer CTA) into that peer's receive buffer: a
// reduce-scatter over distributed shared memory, completion counted on the
// peer's mbarrier.
template <int MODE, int NJ, int EH>
__global__ void __launch_bounds__(threads_of(EH), 1)
rec_tc_bwd_kernel(RecArgs a, const __grid_constant__ RecMaps tm) {
  constexpr int G = MODE == 2 ? 4 : (MODE == 3 ? 3 : 1);
  constexpr int BC = 4 * NJ;
  constexpr int NJL = NJ / EH, BCL = 4 * NJL;   // utterance groups / columns per epilogue thread
  constexpr int kThreads = threads_of(EH);
  static_assert(NJ % EH == 0, "split epilogue needs an even number of utterance groups");
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int NC = a.NC;
  const int dir = blockIdx.y % a.dirs, chunk = blockIdx.y / a.dirs;
  const int b_lo = chunk * BC, nb = min(BC, a.B - b_lo);
  const int H = a.H, T = a.T, B = a.B, GH = G * H, HO = H * a.dirs;
  const int MT = (H + 127) / 128;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t tmem_cols = 64 + MT * 64 <= 256 ? 256u : 512u;

  extern __shared__ uint8_t smem_dyn[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t *dgs = smem;                                              // 2 k-blocks x [16 rows x 128 B]
  float *recv = reinterpret_cast<float *>(smem + 4096);            // [2][NC][BC/4 groups][32 units][4 utterances]
  const int recv_floats = NC * 32 * BC;
  uint64_t *rfull = reinterpret_cast<uint64_t *>(smem + 4096 + 2 * recv_floats * 4);  // [2]
  uint64_t *acc_full = rfull + 2;          // [4]: one per M tile, so a tile can leave while the next computes
  uint64_t *dg_ready = acc_full + 4;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(dg_ready + 1);
  const uint32_t r_bytes = (uint32_t)recv_floats * 4u;  // bytes one step delivers into a receive buffer

  for (int idx = tid; idx < 4096 / 16; idx += kThreads) reinterpret_cast<uint4 *>(dgs)[idx] = make_uint4(0, 0, 0, 0);
  if (tid == 0) {
    mbar_init(rfull + 0, 1);
    mbar_init(rfull + 1, 1);
    for (int m = 0; m < 4; m++) mbar_init(acc_full + m, 1);  // committed by the warp that issued the tile
    mbar_init(dg_ready, 128 * EH);   // (kept for the layout; the hand-off itself uses named barrier 4)
    fence_barrier_init();
    mbar_expect_tx(rfull + 0, r_bytes);
    mbar_expect_tx(rfull + 1, r_bytes);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, tmem_cols);
    tmem_relinquish();
  }
  fence_proxy_async_all();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // ---- one-time: R_slice^T -> BF16 -> tensor memory.  lane = k within the tile, column c = rows 2c, 2c+1
  if (warp >= kIssuers && warp < kIssuers + 4) {
    const int q = warp & 3;
    const float *Rg = a.w_rec[dir];
    for (int m = 0; m < MT; m++) {
      const int k = m * 128 + q * 32 + lane;
      for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const int r0 = 2 * (c0 + i);  // rows r0 = 4u+g, r0+1 = 4u+g+1 (same unit)
          const int u = r0 >> 2, g0 = r0 & 3;
          float f0 = 0.f, f1 = 0.f;
          if (k < H) {
            if (g0 < G) f0 = Rg[((size_t)g0 * H + crank * UT + u) * H + k];
            if (g0 + 1 < G) f1 = Rg[((size_t)(g0 + 1) * H + crank * UT + u) * H + k];
          }
          v[i] = pack_bf16(f0, f1);
        }
        tmem_st_32x16(tmem_base + ((uint32_t)(q * 32) << 16) + 64 + m * 64 + c0, v);
      }
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  cluster.sync();

  if (warp < kIssuers) {
    // ===================== MMA issuers: warp w -> M tile w (MT <= 4) =====================
    constexpr uint32_t idesc = instr_desc(kFmtBF16, 0, 0, 128, NPAD);
    const uint32_t dg0 = __shfl_sync(0xffffffffu, smem_u32(dgs), 0);
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint64_t bd0 = smem_desc(dg0, 0, 1024, kLayoutSw128);
    const bool profm = kPhaseCounters && a.dbg != nullptr && crank == 0 && blockIdx.y == 0 && warp == 1;
    long long qm[3] = {0, 0, 0};
    for (int step = 0; step + 1 < T; step++) {
      const long long n0 = profm ? clock64() : 0;
      // (a named barrier, not the mbarrier: warps sleeping in mbarrier.try_wait resume later -- measured 0.735 -> 0.723 us per step)
      asm volatile("bar.sync 4, %0;" ::"n"(32 * kIssuers + 128 * EH) : "memory");
      const long long n1 = profm ? clock64() : 0;
      tc_fence_after();
      for (int m = warp; m < MT; m += kIssuers) {
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
          const uint64_t bd = bd0 + (uint64_t)(((kk >> 2) * 2048 + (kk & 3) * 32) >> 4);
          if (elect_one()) mma_bf16_ts(tmem_d + m * NPAD, tmem_d + 64 + m * 64 + kk * 8, bd, idesc, kk ? 1u : 0u);
        }
        if (elect_one()) tc_commit(acc_full + m);
      }
      __syncwarp();
      if (profm) {
        const long long n2 = clock64();
        qm[0] += n1 - n0; qm[1] += n2 - n1;
      }
    }
    if (profm && lane == 0) { a.dbg[8] = qm[0]; a.dbg[9] = qm[1]; }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int eh = (warp - kIssuers) >> 2;   // which half of the chunk's columns (split epilogue)
    const int jb = eh * NJL;                 // my first utterance group
    const int s = lane & 3;
    const int ul = q * 8 + (lane >> 2);
    const int unit = crank * UT + ul;
    float *gates = a.gates[dir];
    float *cell = a.cell[dir];
    const bool prof = kPhaseCounters && a.dbg != nullptr && crank == 0 && blockIdx.y == 0 && warp == kIssuers;
    long long pe[7] = {0, 0, 0, 0, 0, 0, 0};
    float carry[NJL];  // LSTM: dc carried to the previous step; GRU: dh * z
#pragma unroll
    for (int j = 0; j < NJL; j++) carry[j] = 0.f;
    float bsum[4] = {0.f, 0.f, 0.f, 0.f}, bsq = 0.f;  // bias gradients: sums over time and my utterances

    // operands of the step, prefetched kD steps ahead into rotating registers
    struct Ops {
      float dy[NJL], g[NJL][G], c[NJL], cp[NJL];
    };
    constexpr int kD = NJL == 1 ? 8 : 2;   // prefetch distance in steps (registers: 7 * NJL per step in flight)
    Ops opsR[kD];
    auto load_step = [&](Ops &o, int step) {
      float (&pdy)[NJL] = o.dy;
      float (&pg)[NJL][G] = o.g;
      float (&pc)[NJL] = o.c;
      float (&pcp)[NJL] = o.cp;
      const int fstep = T - 1 - step;
      const int t = dir ? T - 1 - fstep : fstep;
      const int tp = dir ? t + 1 : t - 1;
#pragma unroll
      for (int j = 0; j < NJL; j++) {
        const int b = 4 * (jb + j) + s;
        pdy[j] = 0.f;
        pc[j] = pcp[j] = 0.f;
#pragma unroll
        for (int g = 0; g < G; g++) pg[j][g] = 0.f;
        if (b < nb) {
          const size_t row = (size_t)t * B + b_lo + b;
          pdy[j] = a.dy[row * HO + dir * H + unit];
#pragma unroll
          for (int g = 0; g < G; g++) pg[j][g] = gates[row * GH + (size_t)g * H + unit];
          if (MODE >= 2) pc[j] = cell[row * H + unit];
          if (fstep > 0) {
            const size_t rowp = (size_t)tp * B + b_lo + b;
            if (MODE == 2) pcp[j] = cell[rowp * H + unit];
            if (MODE == 3) pcp[j] = a.y[rowp * HO + dir * H + unit];
          }
        }
      }
    };
    // Batch chunks of 8 / 16 (kBulk): the same operands through TMA tile loads into a 3-slot ring in shared
    // memory instead of 7 scalar loads per (thread, utterance) and step: dy [BC][32], gates [BC][G][32],
    // c [BC][32], c_prev / h_prev [BC][32] -- three or four TMA operations per step, issued by one thread,
    // counted on the slot's mbarrier.
    constexpr bool kBulk = BC >= 8;      // (at 4 utterances the rotating registers win: 0.78 vs 0.84 us per step)
    constexpr bool kBulkRS = BC >= 16;   // bulk-copy reduce-scatter (measured slower than st.async at 8 utterances)
    constexpr int kRB = 3;
    constexpr int kOffG = BC * 32, kOffC = kOffG + BC * G * 32, kOffP = kOffC + BC * 32;
    constexpr int kSlotFloats = kOffP + BC * 32;
    float *oring = reinterpret_cast<float *>(smem + 4096 + 2 * (size_t)recv_floats * 4 + 256);
    float *stage = oring + kRB * kSlotFloats;   // [epilogue warp][M tile][NJL x 32 units x 4 slots]: outgoing partial sums
    uint64_t *opbar = reinterpret_cast<uint64_t *>(smem + 4096 + 2 * (size_t)recv_floats * 4 + 128);  // [kRB]
    auto issue_ops = [&](int step) {
      if (tid != 32 * kIssuers) return;
      const int fstep = T - 1 - step;
      const int t = dir ? T - 1 - fstep : fstep;
      const int tp = dir ? t + 1 : t - 1;
      const bool prev = MODE >= 2 && fstep > 0;   // no previous frame at the first one
      uint64_t *bar = opbar + step % kRB;
      float *slot = oring + (size_t)(step % kRB) * kSlotFloats;
      const int row = t * B + b_lo;
      mbar_expect_tx(bar, (uint32_t)(BC * 32 * (1 + G + (MODE >= 2 ? 1 : 0) + (prev ? 1 : 0))) * 4u);
      tma_load_2d(slot, &tm.dy, dir * H + crank * UT, row, bar);
      tma_load_3d(slot + kOffG, &tm.gates[dir], crank * UT, 0, row, bar);
      if (MODE >= 2) tma_load_2d(slot + kOffC, &tm.cell[dir], crank * UT, row, bar);
      if (prev) {
        if (MODE == 2) tma_load_2d(slot + kOffP, &tm.cell[dir], crank * UT, tp * B + b_lo, bar);
        else tma_load_2d(slot + kOffP, &tm.y, dir * H + crank * UT, tp * B + b_lo, bar);
      }
    };
    auto read_ops = [&](Ops &o, int step) {
      mbar_wait(opbar + step % kRB, (uint32_t)(step / kRB) & 1u);
      const bool first_frame = step == T - 1;
      const float *slot = oring + (size_t)(step % kRB) * kSlotFloats;
#pragma unroll
      for (int j = 0; j < NJL; j++) {
        const int b = 4 * (jb + j) + s;
        o.dy[j] = slot[b * 32 + ul];
#pragma unroll
        for (int g = 0; g < G; g++) o.g[j][g] = slot[kOffG + (b * G + g) * 32 + ul];
        o.c[j] = MODE >= 2 ? slot[kOffC + b * 32 + ul] : 0.f;
        o.cp[j] = (MODE >= 2 && !first_frame) ? slot[kOffP + b * 32 + ul] : 0.f;
      }
    };
    if (kBulk) {
      if (tid == 32 * kIssuers) {
        for (int k = 0; k < kRB; k++) mbar_init(opbar + k, 1);
        fence_barrier_init();
      }
      asm volatile("bar.sync 3, %0;" ::"n"(128 * EH) : "memory");
      for (int i = 0; i < kRB; i++)
        if (i < T) issue_ops(i);
    } else {
#pragma unroll
      for (int i = 0; i < kD; i++)
        if (i < T) load_step(opsR[i], i);
    }

    // destination of my TMEM lanes' partial sums: for tile m, lanes of this warp are the 32
    // units of CTA 4m+q; inside its receive buffer: [parity][src = crank][lane][b]
    uint32_t rdst[4], rbar[4];
#pragma unroll
    for (int m = 0; m < 4; m++) {
      const int tgt = 4 * m + q;
      const int ok = m < MT && tgt < NC;
      // receive layout [src][utterance group][unit][slot]: a reading warp's 8 units x 4 slots are 32 consecutive
      // floats (with [src][unit][utterance] the reads were 4-way bank-conflicted at 16 utterances: 820 cycles per step)
      rdst[m] = mapa_u32(smem_u32(recv) + (uint32_t)(((crank * NJ + jb) * 32 + lane) * 4) * 4u, ok ? tgt : 0);
      rbar[m] = mapa_u32(smem_u32(rfull), ok ? tgt : 0);
    }

    auto do_step = [&](const int step, Ops &ops) {
      if (kBulk) read_ops(ops, step);
      float (&pdy)[NJL] = ops.dy;
      float (&pg)[NJL][G] = ops.g;
      float (&pc)[NJL] = ops.c;
      float (&pcp)[NJL] = ops.cp;
      const int fstep = T - 1 - step;
      const int t = dir ? T - 1 - fstep : fstep;
      const int p = step & 1;
      // ---- dh arriving from the step processed before (frame t +- 1)
      float dhr[NJL];
#pragma unroll
      for (int j = 0; j < NJL; j++) dhr[j] = 0.f;
      const long long c0 = prof ? clock64() : 0;
      long long c1 = c0;
      if (step > 0) {
        const int use = p ? (step - 1) >> 1 : (step >> 1) - 1;
        mbar_wait(rfull + p, use & 1);
        if (prof) c1 = clock64();
        // re-arm at once: the next fill of this buffer is two steps away
        if (warp == kIssuers && lane == 0) mbar_expect_tx(rfull + p, r_bytes);
        // all (<= 16) partial sums are loaded back to back, then added in a fixed pairwise order: a
        // rolled loop of dependent load->add pairs cost 412 cycles per step here (measured)
        const float *rc = recv + (size_t)p * recv_floats + (jb * 32 + ul) * 4 + s;
#pragma unroll
        for (int j = 0; j < NJL; j++) {
          float v[16];
#pragma unroll
          for (int src = 0; src < 16; src++) v[src] = src < NC ? rc[src * 32 * BC + j * 128] : 0.f;  // 16 loads in flight
#pragma unroll
          for (int w2 = 8; w2 >= 1; w2 >>= 1)   // fixed-order tree: deterministic
#pragma unroll
            for (int i = 0; i < w2; i++) v[i] += v[i + w2];
          dhr[j] = v[0];
        }
      }
      const long long c2 = prof ? clock64() : 0;
      // ---- gate gradients of (unit, batch 4j+s)
      float dgv[NJL][4], dq[NJL];
#pragma unroll
      for (int j = 0; j < NJL; j++) {
        float dh = pdy[j] + dhr[j];
        dgv[j][0] = dgv[j][1] = dgv[j][2] = dgv[j][3] = 0.f;
        dq[j] = 0.f;
        if (MODE == 2) {
          const float i = pg[j][0], f = pg[j][1 % G], g_ = pg[j][2 % G], o = pg[j][3 % G];
          const float tc_ = tanh_fast(pc[j]);
          const float dc = fmaf(dh * o, 1.f - tc_ * tc_, carry[j]);
          dgv[j][0] = dc * g_ * i * (1.f - i);
          dgv[j][1] = dc * pcp[j] * f * (1.f - f);
          dgv[j][2] = dc * i * (1.f - g_ * g_);
          dgv[j][3] = dh * tc_ * o * (1.f - o);
          carry[j] = dc * f;
        } else if (MODE == 3) {
          dh += carry[j];
          const float r = pg[j][0], z = pg[j][1 % G], n = pg[j][2 % G], qv = pc[j];
          const float dn = dh * (1.f - z) * (1.f - n * n);
          dgv[j][0] = dn * qv * r * (1.f - r);
          dgv[j][1] = dh * (pcp[j] - n) * z * (1.f - z);
          dgv[j][2] = dn;       // input side of the n gate
          dq[j] = dn * r;       // recurrent side
          carry[j] = dh * z;
        } else {
          const float h = pg[j][0];
          dgv[j][0] = dh * (MODE == 0 ? (h > 0.f ? 1.f : 0.f) : (1.f - h * h));
        }
        if (4 * (jb + j) + s >= nb) {
          dgv[j][0] = dgv[j][1] = dgv[j][2] = dgv[j][3] = 0.f;
          dq[j] = 0.f;
          carry[j] = 0.f;
        }
        bsum[0] += dgv[j][0]; bsum[1] += dgv[j][1]; bsum[2] += dgv[j][2]; bsum[3] += dgv[j][3];
        bsq += dq[j];
      }
      const long long c3 = prof ? clock64() : 0;
      if (step + 1 < T) {
        // ---- recurrent-side gradients -> BF16 B tile [utterance row][gate row r = 4*ul + g]
#pragma unroll
        for (int j = 0; j < NJL; j++) {
          const int b = 4 * (jb + j) + s;
          const float g2 = MODE == 3 ? dq[j] : dgv[j][2];
          const uint2 v = make_uint2(pack_bf16(dgv[j][0], dgv[j][1]), pack_bf16(g2, dgv[j][3]));
          const uint32_t off = (ul >> 4) * 2048 + b * 128 + ((((ul & 15) >> 1) ^ (b & 7)) << 4) + (ul & 1) * 8;
          *reinterpret_cast<uint2 *>(dgs + off) = v;
        }
        fence_proxy_async();  // my generic smem writes -> visible to the tensor core (async proxy)
        asm volatile("bar.arrive 4, %0;" ::"n"(32 * kIssuers + 128 * EH) : "memory");   // gate gradients of this step are in the B tile
      }
      const long long c4 = prof ? clock64() : 0;
      // ---- off the critical path: gradients to HBM, operands of the next step
#pragma unroll
      for (int j = 0; j < NJL; j++) {
        const int b = 4 * (jb + j) + s;
        if (b < nb) {
          const size_t row = (size_t)t * B + b_lo + b;
          float *gp = gates + row * GH + unit;
          gp[0] = dgv[j][0];
          if (G > 1) { gp[H] = dgv[j][1]; gp[2 * H] = dgv[j][2]; }
          if (G > 3) gp[3 * H] = dgv[j][3];
          if (MODE == 3) cell[row * H + unit] = dq[j];
        }
      }
      if (step + 1 < T) {
        // my (unit, batch) operands of later steps are not touched by anyone else: safe to prefetch now
        if (!kBulk && step + kD < T) load_step(ops, step + kD);
        const long long c5 = prof ? clock64() : 0;
        // ---- partial dh_{prev} of my rows, for all k: scatter to the owners
        const int pn = (step + 1) & 1;
        long long c6 = c5;
        // tile by tile (each issued and committed by its own warp):
        // the partial sums of an early tile are in flight while the later tiles still compute
        if (kBulkRS) {   // last step's bulk copies have long read their staging area; make it formal
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
        }
#pragma unroll
        for (int m = 0; m < 4; m++) {
          if (m < MT) {
            mbar_wait(acc_full + m, step & 1);
            if (prof && m == 0) {
              c6 = clock64();
              pe[0] += c1 - c0; pe[1] += c2 - c1; pe[2] += c3 - c2; pe[3] += c4 - c3; pe[4] += c5 - c4; pe[5] += c6 - c5;
            }
            tc_fence_after();
            uint32_t r[BCL];
            tmem_ld_32xN<BCL>(tmem_base + ((uint32_t)(q * 32) << 16) + m * NPAD + eh * BCL, r);
            tmem_ld_wait();
            if (4 * m + q < NC) {
              const uint32_t dst = rdst[m] + (uint32_t)pn * r_bytes, bar = rbar[m] + pn * 8;
              if (kBulkRS) {
                // 16 utterances: 1536 st.async per step and CTA (one mbarrier update each at the receiver) cost
                // ~1200 cycles per step.  My warp's part of peer 4m+q's block is contiguous there
                // ([group][unit][slot]): park it in shared memory and send it as ONE bulk copy.
                float4 *st4 = reinterpret_cast<float4 *>(stage) + ((warp - kIssuers) * 4 + m) * NJL * 32;
#pragma unroll
                for (int j = 0; j < NJL; j++)
                  st4[j * 32 + lane] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                   __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                  asm volatile(
                      "cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                          dst),
                      "r"(smem_u32(st4)), "n"(NJL * 512), "r"(bar)
                      : "memory");
                  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
              } else {
#pragma unroll
                for (int j = 0; j < NJL; j++)
                  st_async_v4(dst + j * 512, r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3], bar);
              }
            }
          }
        }
        tc_fence_before();
        if (prof) pe[6] += clock64() - c6;
        // every epilogue thread has arrived on the gate-gradient barrier (the MMAs above needed it), i.e. is past its reads of
        // this step's ring slot: refill it
        if (kBulk && step + kRB < T) issue_ops(step + kRB);
      }
    };
    for (int step = 0; step < T; step += kD) {
#pragma unroll
      for (int i = 0; i < kD; i++)
        if (step + i < T) do_step(step + i, opsR[i]);
    }
    if (prof && lane == 0)
      for (int i = 0; i < 7; i++) a.dbg[i] = pe[i];
    // ---- bias gradients of this chunk: add the four utterance slots, slot 0 writes
    if (a.bias_partial) {
#pragma unroll
      for (int g = 0; g < 4; g++) {
        bsum[g] += __shfl_xor_sync(0xffffffffu, bsum[g], 1);
        bsum[g] += __shfl_xor_sync(0xffffffffu, bsum[g], 2);
      }
      bsq += __shfl_xor_sync(0xffffffffu, bsq, 1);
      bsq += __shfl_xor_sync(0xffffffffu, bsq, 2);
      float *bp = a.bias_partial + (size_t)(chunk * a.dirs + dir) * 2 * GH;
      if (s == 0 && eh == 0) {
#pragma unroll
        for (int g = 0; g < G; g++) {
          bp[(size_t)g * H + unit] = bsum[g];                                        // input side
          bp[GH + (size_t)g * H + unit] = (MODE == 3 && g == 2) ? bsq : bsum[g];    // recurrent side
        }
      }
      if (EH > 1) {  // the second set of epilogue warps adds its utterances' sums (fixed order: deterministic)
        asm volatile("bar.sync 3, %0;" ::"n"(128 * EH) : "memory");
        if (s == 0 && eh == 1) {
#pragma unroll
          for (int g = 0; g < G; g++) {
            bp[(size_t)g * H + unit] += bsum[g];
            bp[GH + (size_t)g * H + unit] += (MODE == 3 && g == 2) ? bsq : bsum[g];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster.sync();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

__global__ void bias_finalize_kernel(const float *partial, int nchunks, int dirs, int GH, float *o00, float *o01,
                                     float *o10, float *o11) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= GH) return;
  for (int d = 0; d < dirs; d++)
    for (int side = 0; side < 2; side++) {
      float sum = 0.f;
      for (int c = 0; c < nchunks; c++) sum += partial[((size_t)(c * dirs + d) * 2 + side) * GH + n];
      float *o = d == 0 ? (side == 0 ? o00 : o01) : (side == 0 ? o10 : o11);
      o[n] += sum;
    }
}

// descriptors for the kBulk variants (boxes of BC utterance rows x this CTA's 32 units)
bool make_rec_maps(const RecArgs &a, bool backward, RecMaps *m) {
  memset(m, 0, sizeof(*m));
  if (a.BC < 8 && backward) return true;   // the backward kernel keeps its register prefetch at 4 utterances
  const int G = a.mode == 2 ? 4 : (a.mode == 3 ? 3 : 1);
  const long long rows = (long long)a.T * a.B, H = a.H, HO = (long long)a.H * a.dirs;
  for (int d = 0; d < a.dirs; d++) {
    const long long dims3[3] = {H, G, rows}, str3[2] = {H, G * H};
    const int box3[3] = {UT, G, a.BC};
    if (!make_map_nd(&m->gates[d], a.gates[d], 3, dims3, str3, box3)) return false;
    if (backward && a.mode >= 2) {
      const long long dims2[2] = {H, rows}, str2[1] = {H};
      const int box2[2] = {UT, a.BC};
      if (!make_map_nd(&m->cell[d], a.cell[d], 2, dims2, str2, box2)) return false;
    }
  }
  if (backward) {
    const long long dims2[2] = {HO, rows}, str2[1] = {HO};
    const int box2[2] = {UT, a.BC};
    if (!make_map_nd(&m->dy, a.dy, 2, dims2, str2, box2)) return false;
    if (a.mode == 3 && !make_map_nd(&m->y, a.y, 2, dims2, str2, box2)) return false;
  }
  return true;
}

template <typename K>
cudaError_t launch_cluster(K kernel, const RecArgs &a, const RecMaps &tm, size_t smem, cudaStream_t stream,
                           int threads = kThreads) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (a.NC > 8) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  const int nchunks = (a.B + a.BC - 1) / a.BC;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.NC, a.dirs * nchunks, 1);
  cfg.blockDim = dim3(threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.NC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, a, tm);
}

// tuning aid: B200RNN_SPLIT_EPILOGUE=0 keeps four epilogue warps at every batch chunk
bool split_epilogue() {
  static const bool on = !(getenv("B200RNN_SPLIT_EPILOGUE") && atoi(getenv("B200RNN_SPLIT_EPILOGUE")) == 0);
  return on;
}

// >= 116 KB so that two CTAs never share an SM (each owns 256+ TMEM columns and an issue slot)
constexpr size_t kSmemFloor = 116 * 1024;
// h tiles + barriers | operand ring (32 KB) | staging of the outgoing h slice (2 KB, 16 utterances)
size_t fwd_smem_bytes(int H) { return std::max(kSmemFloor, 1024 + (size_t)kRingOffset + 32 * 1024 + 2048); }

template <int MODE>
cudaError_t launch_fwd(const RecArgs &a, cudaStream_t stream) {
  const size_t smem = fwd_smem_bytes(a.H);
  RecMaps tm;
  if (!make_rec_maps(a, false, &tm)) return cudaErrorInvalidValue;
  const bool split = split_epilogue();
  if (a.H == 320) {  // the benchmark width: K extent known at compile time
    switch (a.BC) {
      case 4: return launch_cluster(rec_tc_fwd_kernel<MODE, 1, 5, 1>, a, tm, smem, stream);
      case 8:
        return split ? launch_cluster(rec_tc_fwd_kernel<MODE, 2, 5, 2>, a, tm, smem, stream, threads_of(2))
                     : launch_cluster(rec_tc_fwd_kernel<MODE, 2, 5, 1>, a, tm, smem, stream);
      default:
        return split ? launch_cluster(rec_tc_fwd_kernel<MODE, 4, 5, 2>, a, tm, smem, stream, threads_of(2))
                     : launch_cluster(rec_tc_fwd_kernel<MODE, 4, 5, 1>, a, tm, smem, stream);
    }
  }
  switch (a.BC) {
    case 4: return launch_cluster(rec_tc_fwd_kernel<MODE, 1, 0, 1>, a, tm, smem, stream);
    case 8:
      return split ? launch_cluster(rec_tc_fwd_kernel<MODE, 2, 0, 2>, a, tm, smem, stream, threads_of(2))
                   : launch_cluster(rec_tc_fwd_kernel<MODE, 2, 0, 1>, a, tm, smem, stream);
    default:
      return split ? launch_cluster(rec_tc_fwd_kernel<MODE, 4, 0, 2>, a, tm, smem, stream, threads_of(2))
                   : launch_cluster(rec_tc_fwd_kernel<MODE, 4, 0, 1>, a, tm, smem, stream);
  }
}

size_t bwd_smem_bytes(int H, int BC) {
  // operand ring of the TMA variant + staging of the outgoing partial sums (8 warps x 4 tiles x BC/8 x 512 B)
  const size_t ring = BC >= 8 ? 256 + (size_t)3 * BC * 7 * 32 * 4 + (size_t)32 * (BC / 8) * 512 : 128;
  return std::max(kSmemFloor, 1024 + 4096 + (size_t)2 * (H / UT) * 32 * BC * 4 + ring);
}

template <int MODE>
cudaError_t launch_bwd(const RecArgs &a, cudaStream_t stream) {
  const size_t smem = bwd_smem_bytes(a.H, a.BC);
  RecMaps tm;
  if (!make_rec_maps(a, true, &tm)) return cudaErrorInvalidValue;
  const bool split = split_epilogue();
  switch (a.BC) {
    case 4: return launch_cluster(rec_tc_bwd_kernel<MODE, 1, 1>, a, tm, smem, stream);
    case 8:
      return split ? launch_cluster(rec_tc_bwd_kernel<MODE, 2, 2>, a, tm, smem, stream, threads_of(2))
                   : launch_cluster(rec_tc_bwd_kernel<MODE, 2, 1>, a, tm, smem, stream);
    default:
      return split ? launch_cluster(rec_tc_bwd_kernel<MODE, 4, 2>, a, tm, smem, stream, threads_of(2))
                   : launch_cluster(rec_tc_bwd_kernel<MODE, 4, 1>, a, tm, smem, stream);
  }
}

}  // namespace

// The tcgen05 kernels need H a multiple of 64 (whole 128-byte swizzle rows of BF16)
// and at most 16 CTAs of 32 units per cluster.
bool rec_tc_supported(int mode, int H) {
  (void)mode;
  return H % 64 == 0 && H / UT <= 16 && kACol + H / 2 <= 512;
}

// batch chunk: the smallest of {4, 8, 16} that keeps all clusters resident at once
int rec_tc_pick_chunk(int H, int B, int dirs) {
  const int NC = H / UT;
  for (int bc : {4, 8, 16})
    if (dirs * ((B + bc - 1) / bc) * NC <= 144) return bc;
  return 16;
}

cudaError_t rec_tc_forward(const RecArgs &a, cudaStream_t stream) {
  if (!rec_tc_supported(a.mode, a.H) || a.NC != a.H / UT) return cudaErrorInvalidValue;
  switch (a.mode) {
    case 0: return launch_fwd<0>(a, stream);
    case 1: return launch_fwd<1>(a, stream);
    case 2: return launch_fwd<2>(a, stream);
    default: return launch_fwd<3>(a, stream);
  }
}

cudaError_t rec_tc_bias_finalize(const float *partial, int nchunks, int dirs, int GH, float *db_in0, float *db_rec0,
                                 float *db_in1, float *db_rec1, cudaStream_t stream) {
  bias_finalize_kernel<<<(GH + 255) / 256, 256, 0, stream>>>(partial, nchunks, dirs, GH, db_in0, db_rec0, db_in1, db_rec1);
  return cudaGetLastError();
}

cudaError_t rec_tc_backward(const RecArgs &a, cudaStream_t stream) {
  if (!rec_tc_supported(a.mode, a.H) || a.NC != a.H / UT) return cudaErrorInvalidValue;
  switch (a.mode) {
    case 0: return launch_bwd<0>(a, stream);
    case 1: return launch_bwd<1>(a, stream);
    case 2: return launch_bwd<2>(a, stream);
    default: return launch_bwd<3>(a, stream);
  }
}

}  // namespace b200
