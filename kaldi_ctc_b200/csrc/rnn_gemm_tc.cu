// kaldi_ctc_b200/csrc/rnn_gemm_tc.cu -- TMA-fed tcgen05 GEMM straight from fp32.
//
//   C[M x N] = alpha * A(M x K) * B(K x N) + beta * C + bias        (fp32 in HBM)
//
// The operands stay fp32 in global memory (they are Kaldi CuMatrix buffers); TMA
// (cp.async.bulk.tensor, 128-byte swizzle) stages 128x32 fp32 tiles into shared
// memory and tcgen05.mma kind::tf32 consumes them directly (the tensor core reads
// the top 19 bits), accumulating in fp32 in TMEM.  No conversion pass, no extra
// HBM traffic.  Both K-major and MN-major operands are supported through the UMMA
// descriptors, so the three products of a layer need no transposes:
//   pre = x . Wi^T   (A K-major, B K-major)     hoisted input projection
//   dx  = dG . Wi    (A K-major, B MN-major)
//   dW  = dG^T . x   (A MN-major, B MN-major), split-K with a fixed-order reduce
// Persistent kernel, one CTA per SM, dynamic tile scheduler:
//   warp 0   = tile scheduler (atomic ticket -> smem tile queue) + TMA producer
//   warp 1   = TMEM owner + MMA issuer (warp-uniform issue, one elected lane)
//   warps 2-9 = epilogue (TMEM -> registers -> global), overlapped with the next tile's main
//               loop through two TMEM accumulator buffers
// Tiles are 128 x TBN with TBN = 256 when N > 128 (fp32 operands make the main loop L2->SMEM
// bound: 48 KB per 2.1 MFLOP instead of 32 KB per 1.05 MFLOP), else 128.  ~192 KB smem ring.
// The ticket counter makes the kernel indifferent to how many of its CTAs are resident: the
// weight-gradient GEMMs run on a side stream next to the persistent recurrent kernels, which pin
// 80 SMs for milliseconds; CTAs that only become resident later find the queue drained and exit.
#include <stdio.h>
#include <string.h>

#include <algorithm>

#include "rnn_common.cuh"
#include "tc_common.cuh"

namespace b200 {
int g_gemm_pair = -1;
int g_gemm_tma_store = 1;
int g_last_gemm_pair = 0;
namespace {

using namespace tc;

constexpr int TBM = 128, TBK = 32;  // tf32: 32 elements = one 128-byte swizzle row
constexpr int kABytes = TBM * TBK * 4;  // 16 KB
constexpr int kEpiWarps = 8;            // two per TMEM lane quarter, each takes half of the tile's columns
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kQ = 4;                   // tile-queue depth

template <int TBN>
struct Cfg {
  static constexpr int kBBytes = TBN * TBK * 4;            // 16 / 32 KB
  static constexpr int kStageBytes = kABytes + kBBytes;    // 32 / 48 KB
  static constexpr int kStages = TBN == 256 ? 4 : 6;       // 192 KB either way
  static constexpr int kTmemCols = 2 * TBN;                // two accumulator buffers
  static constexpr size_t kSmem = 1024 + (size_t)kStages * kStageBytes + 512;
};

struct TcParams {
  int M, N, K;
  float alpha, beta;
  float *C;
  int ldc;
  const float *bias_a, *bias_b;
  int nb;
  int splits, kb_per_split;
  float *partial;
  int tiles_m, tiles_n, total;   // total = tiles_m * tiles_n * splits
  int *ticket;                   // [0] next tile, [1] CTAs finished (both self-resetting)
  int tma_store;                 // CTA-pair kernel: the epilogue leaves through TMA tile stores (beta = 0, no split-K)
  int nkb_total, nkb1;           // CTA-pair kernel: k-blocks in all / of the first operand pair (A, B); the rest come from (A2, B2)
};

template <bool A_KMAJOR, bool B_KMAJOR, int TBN>
__global__ void __launch_bounds__(kThreads, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
  using C = Cfg<TBN>;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem;                           // [kStages][16 KB]
  uint8_t *sB = smem + C::kStages * kABytes;    // [kStages][kBBytes]
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + C::kStages * C::kStageBytes);
  uint64_t *empty = full + C::kStages;
  uint64_t *acc_full = empty + C::kStages;   // [2]
  uint64_t *acc_empty = acc_full + 2;        // [2]
  uint64_t *q_full = acc_empty + 2;          // [kQ]
  uint64_t *q_empty = q_full + kQ;           // [kQ]
  int *tile_q = reinterpret_cast<int *>(q_empty + kQ);  // [kQ]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tile_q + kQ);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (p.K + TBK - 1) / TBK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < C::kStages; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(acc_full + b, 1);
      mbar_init(acc_empty + b, kEpiWarps);
    }
    for (int q = 0; q < kQ; q++) {
      mbar_init(q_full + q, 1);
      mbar_init(q_empty + q, 1 + kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile id -> (split z, row tile, column tile); n fastest so that concurrently processed tiles
  // share their A rows / B columns in L2
  auto decode = [&](int tile, int &m0, int &n0, int &z, int &kb0, int &kb1) {
    const int tn = tile % p.tiles_n;
    const int r = tile / p.tiles_n;
    const int tm = r % p.tiles_m;
    z = r / p.tiles_m;
    m0 = tm * TBM;
    n0 = tn * TBN;
    kb0 = z * p.kb_per_split;
    kb1 = min(nkb, kb0 + p.kb_per_split);
  };

  if (warp == 0) {
    // ===== tile scheduler + TMA producer =====
    if (lane == 0) {
      uint32_t it = 0;
      for (uint32_t qi = 0;; qi++) {
        int tile = atomicAdd(p.ticket, 1);
        if (tile >= p.total) tile = -1;
        mbar_wait(q_empty + (qi % kQ), ((qi / kQ) & 1) ^ 1);
        tile_q[qi % kQ] = tile;
        mbar_arrive(q_full + (qi % kQ));
        if (tile < 0) break;
        int m0, n0, z, kb0, kb1;
        decode(tile, m0, n0, z, kb0, kb1);
        for (int kb = kb0; kb < kb1; kb++, it++) {
          const uint32_t s = it % C::kStages;
          mbar_wait(empty + s, ((it / C::kStages) & 1) ^ 1);
          mbar_expect_tx(full + s, C::kStageBytes);
          uint8_t *a = sA + s * kABytes, *b = sB + s * C::kBBytes;
          if (A_KMAJOR) {
            tma_load_2d(a, &tmA, kb * TBK, m0, full + s);
          } else {
#pragma unroll
            for (int j = 0; j < TBM / 32; j++) tma_load_2d(a + j * (TBK * 128), &tmA, m0 + j * 32, kb * TBK, full + s);
          }
          if (B_KMAJOR) {
            tma_load_2d(b, &tmB, kb * TBK, n0, full + s);
          } else {
#pragma unroll
            for (int j = 0; j < TBN / 32; j++) tma_load_2d(b + j * (TBK * 128), &tmB, n0 + j * 32, kb * TBK, full + s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = instr_desc(kFmtTF32, A_KMAJOR ? 0 : 1, B_KMAJOR ? 0 : 1, TBM, TBN);
    uint32_t it = 0, ai = 0;
    for (uint32_t qi = 0;; qi++) {
      mbar_wait(q_full + (qi % kQ), (qi / kQ) & 1);
      const int tile = tile_q[qi % kQ];
      __syncwarp();
      if (lane == 0) mbar_arrive(q_empty + (qi % kQ));
      if (tile < 0) break;
      int m0, n0, z, kb0, kb1;
      decode(tile, m0, n0, z, kb0, kb1);
      if (kb1 <= kb0) continue;  // empty split: the epilogue writes zeros
      const uint32_t buf = ai & 1;
      mbar_wait(acc_empty + buf, ((ai >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t acc = tmem_base + buf * TBN;
      for (int kb = kb0; kb < kb1; kb++, it++) {
        const uint32_t s = it % C::kStages;
        mbar_wait(full + s, (it / C::kStages) & 1);
        tc_fence_after();
        // warp-uniform issue (descriptors stay in uniform registers), one elected lane issues
        const uint32_t a = smem_u32(sA + s * kABytes), b = smem_u32(sB + s * C::kBBytes);
#pragma unroll
        for (int k = 0; k < TBK / 8; k++) {
          // K-major (SW128, 16 B chunks): 8-row groups are 1024 B apart (SBO); a K step of 8 tf32
          //   is +32 B inside the 128 B swizzle row.
          // MN-major tf32 must use the 32 B-chunk flavour of the 128 B swizzle: atoms of 4 K-rows
          //   x 128 B (512 B apart = SBO), 32-wide MN blocks TBK*128 B apart (LBO); a K step of
          //   8 rows is +1024 B.
          const uint64_t ad = A_KMAJOR ? smem_desc(a + k * 32, 0, 1024, kLayoutSw128)
                                       : smem_desc(a + k * 1024, TBK * 128, 512, kLayoutSw128Base32);
          const uint64_t bd = B_KMAJOR ? smem_desc(b + k * 32, 0, 1024, kLayoutSw128)
                                       : smem_desc(b + k * 1024, TBK * 128, 512, kLayoutSw128Base32);
          if (elect_one()) mma_tf32(acc, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
        }
        if (elect_one()) {
          tc_commit(empty + s);                      // frees the smem stage when these MMAs retire
          if (kb == kb1 - 1) tc_commit(acc_full + buf);  // accumulator complete
        }
        __syncwarp();
      }
      ai++;
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;
    const bool direct = p.splits <= 1;
    const int ldo = direct ? p.ldc : p.N;
    uint32_t ai = 0;
    for (uint32_t qi = 0;; qi++) {
      mbar_wait(q_full + (qi % kQ), (qi / kQ) & 1);
      const int tile = tile_q[qi % kQ];
      __syncwarp();
      if (lane == 0) mbar_arrive(q_empty + (qi % kQ));
      if (tile < 0) break;
      int m0, n0, z, kb0, kb1;
      decode(tile, m0, n0, z, kb0, kb1);
      const bool have = kb1 > kb0;
      const uint32_t buf = ai & 1;
      if (have) {
        mbar_wait(acc_full + buf, (ai >> 1) & 1);
        tc_fence_after();
      }
      const int m = m0 + q * 32 + lane;
      float *out = direct ? p.C : p.partial + (size_t)z * p.M * p.N;
      const bool vec_ok = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
      // Each lane owns one output row and writes it 16 bytes at a time; the eight stores of a chunk
      // fill the lane's 128-byte line back to back and L2 merges them.  (A version that transposed
      // through shared memory to store whole lines per instruction measured 10-25 % slower: the
      // kernel is bound by the write path, not by store instructions.)
#pragma unroll 1
      for (int c = half * (TBN / 64); c < (half + 1) * (TBN / 64); c++) {
        const int nb0 = n0 + c * 32;
        if (nb0 >= p.N) break;
        uint32_t r[32];
        if (have) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * TBN + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) r[j] = 0u;
        }
        if (m < p.M) {
          float *row = out + (size_t)m * ldo;
#pragma unroll
          for (int j4 = 0; j4 < 8; j4++) {
            const int n = nb0 + j4 * 4;
            if (n >= p.N) break;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) v[e] = p.alpha * __uint_as_float(r[j4 * 4 + e]);
            if (vec_ok && n + 3 < p.N) {
              if (direct) {
                if (p.beta != 0.f) {
                  const float4 o = *reinterpret_cast<const float4 *>(row + n);
                  v[0] += p.beta * o.x; v[1] += p.beta * o.y; v[2] += p.beta * o.z; v[3] += p.beta * o.w;
                }
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  if (p.bias_a) v[e] += p.bias_a[n + e];
                  if (p.bias_b && n + e < p.nb) v[e] += p.bias_b[n + e];
                }
              }
              *reinterpret_cast<float4 *>(row + n) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; e++) {
                if (n + e >= p.N) break;
                float x = v[e];
                if (direct) {
                  if (p.beta != 0.f) x += p.beta * row[n + e];
                  if (p.bias_a) x += p.bias_a[n + e];
                  if (p.bias_b && n + e < p.nb) x += p.bias_b[n + e];
                }
                row[n + e] = x;
              }
            }
          }
        }
      }
      if (have) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + buf);  // this warp's quarter of the buffer is drained
        ai++;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
  if (threadIdx.x == 0) {  // last CTA out re-arms the ticket for the next launch that uses this slot
    __threadfence();
    if (atomicAdd(p.ticket + 1, 1) == (int)gridDim.x - 1) {
      p.ticket[0] = 0;
      p.ticket[1] = 0;
      __threadfence();
    }
  }
}

// ===========================================================================
// CTA-pair variant (tcgen05 cta_group::2): 256 x 256 tiles on two SMs of one TPC
// ===========================================================================
// The one-CTA kernel above stages 48 KB of fp32 operands per 2.1 MFLOP and is bound by L2 -> SM delivery
// (10.4 TB/s measured, 41 % of the TF32 rate).  Here a cluster of two CTAs owns a 256 x 256 tile: each CTA
// loads ITS 128 rows of A and ITS half (128 columns) of B -- 32 KB per 2.1 MFLOP and SM -- and the leader
// CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads both halves of B out of the two shared memories
// and accumulates each CTA's 128 rows in that CTA's tensor memory.  Same roles as above:
//   warp 0   leader: ticket scheduler (hands every tile to both CTAs) + TMA producer; peer: TMA producer
//   warp 1   leader: MMA issuer (commits multicast to both CTAs' barriers); both: tensor-memory owner
//   warps 2-9 epilogue of the CTA's own 128 rows
// Barriers that gate the LEADER on work of both CTAs live in the leader and take remote arrivals from the
// peer: full[] (TMA bytes of both CTAs, cp.async.bulk.tensor.cta_group::2), acc_empty[], q_empty[].
// Epilogue: with one output row per lane and 16-byte stores, a warp store touches 32 different lines, and the
// write path -- not the main loop -- bounds the kernel once the operands arrive faster (the K = 40 projection
// needs 87 us just to write its 164 MB).  So the pair kernel's epilogue warps park each [32 rows x 32 columns]
// block in shared memory (128-byte swizzle, conflict-free 16-byte stores) and hand it to a TMA tile store:
// whole 128-byte lines leave the SM without touching the load/store unit (beta = 0 and no split-K only;
// otherwise the register path of the one-CTA kernel).
#ifndef B200_GEMM2_STAGES
#define B200_GEMM2_STAGES 6
#endif
#ifndef B200_GEMM2_EPIBUFS
#define B200_GEMM2_EPIBUFS 1
#endif
constexpr int kStages2 = B200_GEMM2_STAGES;
constexpr int kEpiBufs = B200_GEMM2_EPIBUFS;            // boxes per epilogue warp for the TMA stores
constexpr int kB2Bytes = 128 * TBK * 4;                 // this CTA's half of the 256-column B tile
constexpr int kStage2Bytes = kABytes + kB2Bytes;        // 32 KB
constexpr int kEpiStageBytes = kEpiBufs * 32 * 128;     // per epilogue warp: [32 rows x 128 B] boxes for the TMA stores
constexpr size_t kSmem2 = 1024 + (size_t)kStages2 * kStage2Bytes + (size_t)kEpiWarps * kEpiStageBytes + 512;
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;          // shared::cluster address -> the even (leader) CTA of the pair

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint64_t *bar, uint32_t parity) {
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
// TMA load of a CTA pair: the bytes are counted on the LEADER CTA's barrier (same offset as `bar` here)
__device__ __forceinline__ void tma_load_2d_pair(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all prior tcgen05.mma of this thread arrive on `bar` of BOTH CTAs of the pair when they complete
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
tc_gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                    const __grid_constant__ CUtensorMap tmC, TcParams p) {
  constexpr int TBN = 256, TBM2 = 256;
  extern __shared__ uint8_t smem_dyn[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem;                           // [kStages2][16 KB]
  uint8_t *sB = smem + kStages2 * kABytes;      // [kStages2][16 KB]
  uint8_t *sE = smem + kStages2 * kStage2Bytes; // [kEpiWarps][2][32 rows x 128 B], 1024-byte aligned boxes
  uint64_t *full = reinterpret_cast<uint64_t *>(sE + kEpiWarps * kEpiStageBytes);
  uint64_t *empty = full + kStages2;
  uint64_t *acc_full = empty + kStages2;     // [2]
  uint64_t *acc_empty = acc_full + 2;        // [2]   (the leader's counts both CTAs' epilogue warps)
  uint64_t *q_full = acc_empty + 2;          // [kQ]
  uint64_t *q_empty = q_full + kQ;           // [kQ]  (the leader's counts both CTAs' consumers)
  int *tile_q = reinterpret_cast<int *>(q_empty + kQ);  // [kQ]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tile_q + kQ);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int nkb = p.nkb_total;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (p.nkb1 < p.nkb_total) {
      tma_prefetch_desc(&tmA2);
      tma_prefetch_desc(&tmB2);
    }
    for (int s = 0; s < kStages2; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(acc_full + b, 1);
      mbar_init(acc_empty + b, 2 * kEpiWarps);
    }
    for (int q = 0; q < kQ; q++) {
      mbar_init(q_full + q, 1);
      mbar_init(q_empty + q, 2 * (1 + kEpiWarps));   // leader: MMA + epilogue warps; peer: producer + epilogue warps
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers exist before anyone arrives remotely
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // the same objects in the leader CTA, as shared::cluster addresses
  const uint32_t ld_acc_empty = mapa_u32(smem_u32(acc_empty), 0), ld_q_empty = mapa_u32(smem_u32(q_empty), 0);

  auto decode = [&](int tile, int &m0, int &n0, int &z, int &kb0, int &kb1) {
    const int tn = tile % p.tiles_n;
    const int r = tile / p.tiles_n;
    const int tm = r % p.tiles_m;
    z = r / p.tiles_m;
    m0 = tm * TBM2 + (int)rank * TBM;   // this CTA's 128 rows
    n0 = tn * TBN;
    kb0 = z * p.kb_per_split;
    kb1 = min(nkb, kb0 + p.kb_per_split);
  };

  // Width of a column tile: MN-major B is staged in 32-column boxes, so the last tile of a row only loads and
  // multiplies the 64-column groups that exist (N = 640: tiles of 256, 256, 128 instead of three full ones).
  auto tile_width = [&](int n0) { return B_KMAJOR ? TBN : min(TBN, ((p.N - n0 + 63) >> 6) << 6); };

  if (warp == 0) {
    // ===== leader: tile scheduler; both: TMA producer =====
    if (lane == 0) {
      const uint32_t pr_q_full = mapa_u32(smem_u32(q_full), 1), pr_tile_q = mapa_u32(smem_u32(tile_q), 1);
      uint32_t it = 0;
      for (uint32_t qi = 0;; qi++) {
        const uint32_t slot = qi % kQ;
        int tile;
        if (leader) {
          tile = atomicAdd(p.ticket, 1);
          if (tile >= p.total) tile = -1;
          mbar_wait_acq_cluster(q_empty + slot, ((qi / kQ) & 1) ^ 1);
          tile_q[slot] = tile;
          asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(pr_tile_q + slot * 4), "r"(tile) : "memory");
          mbar_arrive(q_full + slot);
          mbar_arrive_cluster(pr_q_full + slot * 8);
        } else {
          mbar_wait_acq_cluster(q_full + slot, (qi / kQ) & 1);
          tile = tile_q[slot];
          mbar_arrive_cluster(ld_q_empty + slot * 8);
        }
        if (tile < 0) break;
        int m0, n0, z, kb0, kb1;
        decode(tile, m0, n0, z, kb0, kb1);
        const int nw = tile_width(n0), nwh = nw >> 1;   // tile width; this CTA stages half of it
        const int nh = n0 + (int)rank * nwh;
        for (int kbt = kb0; kbt < kb1; kbt++, it++) {
          const bool second = kbt >= p.nkb1;             // k-blocks of the second operand pair follow the first's
          const CUtensorMap *ma = second ? &tmA2 : &tmA, *mb = second ? &tmB2 : &tmB;
          const int kb = second ? kbt - p.nkb1 : kbt;
          const uint32_t s = it % kStages2;
          mbar_wait(empty + s, ((it / kStages2) & 1) ^ 1);
          if (leader) mbar_expect_tx(full + s, 2 * (kABytes + nwh * TBK * 4));   // both CTAs' bytes land on this barrier
          uint8_t *a = sA + s * kABytes, *b = sB + s * kB2Bytes;
          if (A_KMAJOR) {
            tma_load_2d_pair(a, ma, kb * TBK, m0, full + s);
          } else {
#pragma unroll
            for (int j = 0; j < TBM / 32; j++) tma_load_2d_pair(a + j * (TBK * 128), ma, m0 + j * 32, kb * TBK, full + s);
          }
          if (B_KMAJOR) {
            tma_load_2d_pair(b, mb, kb * TBK, nh, full + s);
          } else {
            for (int j = 0; j < nwh / 32; j++) tma_load_2d_pair(b + j * (TBK * 128), mb, nh + j * 32, kb * TBK, full + s);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== leader: MMA issuer =====
    if (leader) {
      constexpr uint32_t idesc0 = instr_desc(kFmtTF32, A_KMAJOR ? 0 : 1, B_KMAJOR ? 0 : 1, TBM2, 0);
      uint32_t it = 0, ai = 0;
      for (uint32_t qi = 0;; qi++) {
        mbar_wait(q_full + (qi % kQ), (qi / kQ) & 1);
        const int tile = tile_q[qi % kQ];
        __syncwarp();
        if (lane == 0) mbar_arrive(q_empty + (qi % kQ));
        if (tile < 0) break;
        int m0, n0, z, kb0, kb1;
        decode(tile, m0, n0, z, kb0, kb1);
        if (kb1 <= kb0) continue;  // empty split: the epilogue writes zeros
        const uint32_t idesc = idesc0 | ((uint32_t)(tile_width(n0) >> 3) << 17);   // N of this tile
        const uint32_t buf = ai & 1;
        mbar_wait_acq_cluster(acc_empty + buf, ((ai >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem_base + buf * TBN;
        for (int kb = kb0; kb < kb1; kb++, it++) {
          const uint32_t s = it % kStages2;
          mbar_wait(full + s, (it / kStages2) & 1);
          tc_fence_after();
          const uint32_t a = smem_u32(sA + s * kABytes), b = smem_u32(sB + s * kB2Bytes);
#pragma unroll
          for (int k = 0; k < TBK / 8; k++) {
            const uint64_t ad = A_KMAJOR ? smem_desc(a + k * 32, 0, 1024, kLayoutSw128)
                                         : smem_desc(a + k * 1024, TBK * 128, 512, kLayoutSw128Base32);
            const uint64_t bd = B_KMAJOR ? smem_desc(b + k * 32, 0, 1024, kLayoutSw128)
                                         : smem_desc(b + k * 1024, TBK * 128, 512, kLayoutSw128Base32);
            if (elect_one()) mma_tf32_pair(acc, ad, bd, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          if (elect_one()) {
            tc_commit_pair(empty + s);                          // frees the stage in both CTAs
            if (kb == kb1 - 1) tc_commit_pair(acc_full + buf);  // accumulator complete, both CTAs
          }
          __syncwarp();
        }
        ai++;
      }
    }
  } else {
    // ===== epilogue of this CTA's 128 rows: TMEM -> registers -> global =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const bool direct = p.splits <= 1;
    const int ldo = direct ? p.ldc : p.N;
    uint32_t ai = 0, nst = 0;   // nst: blocks this warp has handed to TMA stores
    for (uint32_t qi = 0;; qi++) {
      mbar_wait_acq_cluster(q_full + (qi % kQ), (qi / kQ) & 1);
      const int tile = tile_q[qi % kQ];
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(q_empty + (qi % kQ));
        else mbar_arrive_cluster(ld_q_empty + (qi % kQ) * 8);
      }
      if (tile < 0) break;
      int m0, n0, z, kb0, kb1;
      decode(tile, m0, n0, z, kb0, kb1);
      const bool have = kb1 > kb0;
      const uint32_t buf = ai & 1;
      if (have) {
        mbar_wait(acc_full + buf, (ai >> 1) & 1);
        tc_fence_after();
      }
      const int m = m0 + q * 32 + lane;
      float *out = direct ? p.C : p.partial + (size_t)z * p.M * p.N;
      const bool vec_ok = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll 1
      for (int c = half * (TBN / 64); c < (half + 1) * (TBN / 64); c++) {
        const int nb0 = n0 + c * 32;
        if (nb0 >= p.N) break;
        uint32_t r[32];
        if (have) {
          tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * TBN + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) r[j] = 0u;
        }
        if (p.tma_store) {
          // ---- [32 rows x 32 columns] -> shared memory (row = lane, 16-byte chunk j at j ^ (row & 7)) -> TMA store
          uint8_t *box = sE + (warp - 2) * kEpiStageBytes + (nst % kEpiBufs) * 4096;
          if (nst >= kEpiBufs) {   // the store that last read this box must be done with it
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kEpiBufs - 1) : "memory");
            __syncwarp();
          }
#pragma unroll
          for (int j4 = 0; j4 < 8; j4++) {
            const int n = nb0 + j4 * 4;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
              v[e] = p.alpha * __uint_as_float(r[j4 * 4 + e]);
              if (p.bias_a && n + e < p.N) v[e] += p.bias_a[n + e];
              if (p.bias_b && n + e < p.nb) v[e] += p.bias_b[n + e];
            }
            *reinterpret_cast<float4 *>(box + lane * 128 + ((j4 ^ (lane & 7)) << 4)) = make_float4(v[0], v[1], v[2], v[3]);
          }
          fence_proxy_async();   // generic-proxy writes -> visible to the TMA store
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, box, nb0, m0 + q * 32);   // rows >= M and columns >= N are clipped by the TMA unit
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          nst++;
        } else if (m < p.M) {
          float *row = out + (size_t)m * ldo;
#pragma unroll
          for (int j4 = 0; j4 < 8; j4++) {
            const int n = nb0 + j4 * 4;
            if (n >= p.N) break;
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; e++) v[e] = p.alpha * __uint_as_float(r[j4 * 4 + e]);
            if (vec_ok && n + 3 < p.N) {
              if (direct) {
                if (p.beta != 0.f) {
                  const float4 o = *reinterpret_cast<const float4 *>(row + n);
                  v[0] += p.beta * o.x; v[1] += p.beta * o.y; v[2] += p.beta * o.z; v[3] += p.beta * o.w;
                }
#pragma unroll
                for (int e = 0; e < 4; e++) {
                  if (p.bias_a) v[e] += p.bias_a[n + e];
                  if (p.bias_b && n + e < p.nb) v[e] += p.bias_b[n + e];
                }
              }
              *reinterpret_cast<float4 *>(row + n) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
              for (int e = 0; e < 4; e++) {
                if (n + e >= p.N) break;
                float x = v[e];
                if (direct) {
                  if (p.beta != 0.f) x += p.beta * row[n + e];
                  if (p.bias_a) x += p.bias_a[n + e];
                  if (p.bias_b && n + e < p.nb) x += p.bias_b[n + e];
                }
                row[n + e] = x;
              }
            }
          }
        }
      }
      if (have) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {   // this warp's quarter of the buffer is drained: tell the leader's MMA issuer
          if (leader) mbar_arrive(acc_empty + buf);
          else mbar_arrive_cluster(ld_acc_empty + buf * 8);
        }
        ai++;
      }
    }
    if (p.tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores have landed
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's tensor memory is part of the pair's allocation: both CTAs are done with it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (threadIdx.x == 0) {  // last CTA out re-arms the ticket for the next launch that uses this slot
    __threadfence();
    if (atomicAdd(p.ticket + 1, 1) == (int)gridDim.x - 1) {
      p.ticket[0] = 0;
      p.ticket[1] = 0;
      __threadfence();
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// fp32 matrix with `inner` contiguous elements per row, `outer` rows, row pitch ld
bool make_map(CUtensorMap *map, const float *base, long long inner, long long outer, long long ld,
              int box_inner, int box_outer, CUtensorMapSwizzle swz) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 4) || inner <= 0 || outer <= 0) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace

// fp32 tensor of rank 2 or 3 (dims / box innermost first; strides in elements for dims 1..rank-1), no swizzle:
// used by the recurrent kernels to fetch [utterances x gates x 32 units] boxes with one TMA operation
bool make_map_nd(CUtensorMap *map, const float *base, int rank, const long long *dims, const long long *strides,
                 const int *box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn || rank < 2 || rank > 3 || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
  cuuint64_t gdim[3], gstr[2];
  cuuint32_t bx[3], estr[3] = {1, 1, 1};
  for (int i = 0; i < rank; i++) {
    if (dims[i] <= 0 || box[i] <= 0 || box[i] > 256) return false;
    gdim[i] = (cuuint64_t)dims[i];
    bx[i] = (cuuint32_t)box[i];
    if (i > 0) {
      if (strides[i - 1] % 4) return false;
      gstr[i - 1] = (cuuint64_t)strides[i - 1] * 4;
    }
  }
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float *>(base), gdim, gstr, bx, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

namespace {

// Self-resetting ticket slots, handed out round-robin: a slot is reused kTicketSlots launches later,
// long after the launch that last used it has retired (the launch queue is far shallower).
constexpr int kTicketSlots = 2048;

int *ticket_slot() {
  static int *base[64] = {};
  static unsigned seq[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (!base[dev]) {
    if (cudaMalloc(&base[dev], sizeof(int) * 2 * kTicketSlots) != cudaSuccess) return nullptr;
    if (cudaMemset(base[dev], 0, sizeof(int) * 2 * kTicketSlots) != cudaSuccess) return nullptr;
  }
  return base[dev] + 2 * (seq[dev]++ % kTicketSlots);
}

int sm_count() {
  static int n[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  if (!n[dev]) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
  return n[dev] > 0 ? n[dev] : 148;
}

template <bool AK, bool BK, int TBN>
cudaError_t launch(const CUtensorMap &ta, const CUtensorMap &tb, const TcParams &p, cudaStream_t s) {
  const size_t smem = Cfg<TBN>::kSmem;
  static bool attr_done[64] = {};   // per device: cudaFuncSetAttribute does not carry over to other GPUs
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<AK, BK, TBN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const int grid = p.total < sm_count() ? p.total : sm_count();
  tc_gemm_kernel<AK, BK, TBN><<<grid, kThreads, smem, s>>>(ta, tb, p);
  return cudaGetLastError();
}

template <bool AK, bool BK>
cudaError_t launch_pair(const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &ta2, const CUtensorMap &tb2,
                        const CUtensorMap &tcm, const TcParams &p, cudaStream_t s) {
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !attr_done[dev]) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_pair_kernel<AK, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) attr_done[dev] = true;
  }
  const int pairs = std::min(p.total, sm_count() / 2);   // one cluster of two CTAs per TPC
  tc_gemm_pair_kernel<AK, BK><<<2 * pairs, kThreads, kSmem2, s>>>(ta, tb, ta2, tb2, tcm, p);
  return cudaGetLastError();
}
cudaError_t launch_pair_any(bool ak, bool bk, const CUtensorMap &ta, const CUtensorMap &tb, const CUtensorMap &ta2,
                            const CUtensorMap &tb2, const CUtensorMap &tcm, const TcParams &p, cudaStream_t s) {
  if (ak && bk) return launch_pair<true, true>(ta, tb, ta2, tb2, tcm, p, s);
  if (ak) return launch_pair<true, false>(ta, tb, ta2, tb2, tcm, p, s);
  if (bk) return launch_pair<false, true>(ta, tb, ta2, tb2, tcm, p, s);
  return launch_pair<false, false>(ta, tb, ta2, tb2, tcm, p, s);
}

template <int TBN>
cudaError_t launch_any(bool ak, bool bk, const CUtensorMap &ta, const CUtensorMap &tb, const TcParams &p,
                       cudaStream_t s) {
  if (ak && bk) return launch<true, true, TBN>(ta, tb, p, s);
  if (ak) return launch<true, false, TBN>(ta, tb, p, s);
  if (bk) return launch<false, true, TBN>(ta, tb, p, s);
  return launch<false, false, TBN>(ta, tb, p, s);
}

}  // namespace

// Returns cudaErrorNotSupported when the operands do not meet TMA's alignment rules
// (the caller then uses the fp32 path).
cudaError_t gemm_tc(const GemmArgs &g, cudaStream_t stream, int *launches) {
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  if (g.K <= 0) return cudaErrorNotSupported;
  const bool ak = g.sak == 1, bk = g.sbk == 1;
  if (!ak && g.sam != 1) return cudaErrorNotSupported;
  if (!bk && g.sbn != 1) return cudaErrorNotSupported;
  // Tile width and split count from a small cost model: equal-length work items are handed out
  // dynamically to SMs CTAs, so time ~ ceil(items / SMs) * (k-blocks per item * bytes staged per
  // k-block + the item's share of epilogue traffic).  g.splits is the most the workspace allows.
  static const int force_tbn = getenv("B200RNN_GEMM_TBN") ? atoi(getenv("B200RNN_GEMM_TBN")) : 0;
  const int nkb = (g.K + TBK - 1) / TBK;
  const int max_splits = (g.splits > 1 && g.partial) ? g.splits : 1;
  const int sms = sm_count();
  int tbn = 128, splits = 1;
  double best = 1e300;
  for (int w = 128; w <= 256; w += 128) {
    if ((force_tbn == 128 || force_tbn == 256) && w != force_tbn) continue;
    if (w == 256 && g.N <= 128) continue;
    const long tiles = (long)((g.M + TBM - 1) / TBM) * ((g.N + w - 1) / w);
    for (int sp = 1; sp <= max_splits; sp++) {
      const int per = (nkb + sp - 1) / sp;
      const int eff = (nkb + per - 1) / per;  // splits that actually get k-blocks
      if (eff != sp) continue;
      const long items = tiles * sp;
      const double waves = items <= 3L * sms ? (double)((items + sms - 1) / sms) : (double)items / sms + 0.5;
      const double cost = waves * ((double)per * (kABytes + w * TBK * 4) + 0.5 * TBM * w * 4) +
                          (sp > 1 ? 2e-3 * sp * (double)g.M * g.N : 0.0);   // + the reduce pass
      if (cost < best) {
        best = cost;
        tbn = w;
        splits = sp;
      }
    }
  }
  if (g_gemm_pair == 1 && g.N > 128) {   // test hook: the CTA-pair kernel wherever its tile shape exists
    tbn = 256;
    splits = 1;
  }
  // CTA pairs (256 x 256 tiles, each SM stages 32 instead of 48 KB per k-block) for the big products: the hoisted
  // projection and the input gradient of a layer.  Not for the short-and-wide split-K weight gradients (they run
  // on the side stream under a recurrent kernel and are hidden: split-K on CTA pairs was measured at 5-6 % faster per
  // GEMM and no change in the step).
  const bool pair = g_gemm_pair != 0 && tbn == 256 && splits == 1 && (g_gemm_pair == 1 || g.M >= 2048);
  CUtensorMap ta, tb;
  const CUtensorMapSwizzle kmaj = CU_TENSOR_MAP_SWIZZLE_128B, mnmaj = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  bool ok = ak ? make_map(&ta, g.A, g.K, g.M, g.sam, TBK, TBM, kmaj)
               : make_map(&ta, g.A, g.M, g.K, g.sak, 32, TBK, mnmaj);
  ok = ok && (bk ? make_map(&tb, g.B, g.K, g.N, g.sbn, TBK, pair ? 128 : tbn, kmaj)
                 : make_map(&tb, g.B, g.N, g.K, g.sbk, 32, TBK, mnmaj));
  if (!ok) return cudaErrorNotSupported;
  // second operand pair (same shape and strides): only the CTA-pair kernel walks both in one launch
  const bool dual = g.A2 != nullptr && g.B2 != nullptr;
  if (dual && !pair) return cudaErrorNotSupported;
  CUtensorMap ta2 = ta, tb2 = tb;
  if (dual) {
    ok = ak ? make_map(&ta2, g.A2, g.K, g.M, g.sam, TBK, TBM, kmaj) : make_map(&ta2, g.A2, g.M, g.K, g.sak, 32, TBK, mnmaj);
    ok = ok && (bk ? make_map(&tb2, g.B2, g.K, g.N, g.sbn, TBK, 128, kmaj) : make_map(&tb2, g.B2, g.N, g.K, g.sbk, 32, TBK, mnmaj));
    if (!ok) return cudaErrorNotSupported;
  }

  TcParams p;
  p.M = g.M; p.N = g.N; p.K = g.K; p.alpha = g.alpha; p.beta = g.beta;
  p.C = g.C; p.ldc = g.ldc; p.bias_a = g.bias_a; p.bias_b = g.bias_b; p.nb = g.nb;
  p.nkb1 = nkb;
  p.nkb_total = dual ? 2 * nkb : nkb;
  p.kb_per_split = pair ? p.nkb_total : (nkb + splits - 1) / splits;   // (the CTA-pair kernel never splits K)
  p.splits = splits;
  p.partial = g.partial;
  p.tiles_m = pair ? (g.M + 2 * TBM - 1) / (2 * TBM) : (g.M + TBM - 1) / TBM;
  p.tiles_n = (g.N + tbn - 1) / tbn;
  p.total = p.tiles_m * p.tiles_n * splits;
  p.ticket = ticket_slot();
  if (!p.ticket) return cudaErrorMemoryAllocation;
  g_last_gemm_pair = pair ? 1 : 0;
  CUtensorMap tcm;
  memset(&tcm, 0, sizeof(tcm));
  // output through TMA tile stores ([32 rows x 32 columns] boxes, 128-byte swizzle) when nothing has to be read back
  p.tma_store = (pair && g.beta == 0.f && g_gemm_tma_store != 0 &&
                 make_map(&tcm, g.C, g.N, g.M, g.ldc, 32, 32, kmaj)) ? 1 : 0;
  cudaError_t e = pair ? launch_pair_any(ak, bk, ta, tb, ta2, tb2, tcm, p, stream)
                       : (tbn == 256 ? launch_any<256>(ak, bk, ta, tb, p, stream) : launch_any<128>(ak, bk, ta, tb, p, stream));
  if (e != cudaSuccess) return e;
  if (launches) (*launches)++;
  if (splits > 1) {
    GemmArgs r = g;
    r.splits = splits;
    e = splitk_reduce(r, stream);
    if (launches) (*launches)++;
  }
  return e;
}

}  // namespace b200
