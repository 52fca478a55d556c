// kaldi_ctc_b200/csrc/rnn_rec_tc.cu -- persistent recurrent kernels on tcgen05.
//
// One thread-block cluster per (direction, batch chunk).  CTA c of the cluster owns
// the hidden units [32c, 32c+32): its 128 gate rows (row = 4*unit + gate; unused
// gate slots of GRU / plain RNN are zero rows) of the recurrent matrix R stay in
// shared memory as BF16 for all T steps, laid out as a K-major, 128-byte-swizzled
// UMMA operand.  Per time step:
//   MMA warp     waits until every CTA's slice of h_{t-1} has landed in the local
//                (double-buffered, swizzled) h tile, then one elected lane issues
//                H/16 tcgen05.mma (M=128, N=16, K=16, kind::f16/BF16) accumulating
//                the [128 gate rows x 16 utterances] pre-activations in TMEM;
//                tcgen05.commit signals the epilogue.
//   epilogue     4 warps: tcgen05.ld their 32 TMEM lanes, add the hoisted input
//                projection (prefetched from HBM one step ahead), apply the gate
//                non-linearity, 4x4 quad transposes by warp shuffle so that one
//                thread holds all gates of a (unit, utterance), fp32 cell update
//                (cell / hidden state live in registers), store y / gates / cell,
//                pack h_t to BF16 and write the 16-byte chunks straight into every
//                CTA's next h tile through distributed shared memory, then one
//                remote mbarrier arrive per peer.
// No cluster-wide barrier inside the loop: the only synchronisation is the
// per-CTA "h tile full" mbarrier fed by remote arrives.
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "rnn_common.cuh"
#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace b200 {
namespace {

using namespace tc;

constexpr int UT = 32;          // hidden units per CTA
constexpr int NPAD = 16;        // MMA N (utterances per chunk, zero padded)
constexpr int kThreads = 192;   // warp 0: spare/loader, warp 1: MMA, warps 2-5: epilogue

__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(cta));
  return r;
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
template <int MODE, int NJ>
__global__ void __launch_bounds__(kThreads, 1) rec_tc_fwd_kernel(RecArgs a) {
  constexpr int G = MODE == 2 ? 4 : (MODE == 3 ? 3 : 1);
  constexpr int BC = 4 * NJ;
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int NC = a.NC;
  const int dir = blockIdx.y % a.dirs, chunk = blockIdx.y / a.dirs;
  const int b_lo = chunk * BC, nb = min(BC, a.B - b_lo);
  const int H = a.H, T = a.T, B = a.B, GH = G * H, HO = H * a.dirs;
  const int nkb = H / 64;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  extern __shared__ uint8_t smem_dyn[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t *Rs = smem;
  uint8_t *hs = Rs + (size_t)nkb * 16384;             // two buffers of nkb * 2048 B
  const int hs_bytes = nkb * 2048;
  uint64_t *hfull = reinterpret_cast<uint64_t *>(hs + 2 * hs_bytes);  // [2]
  uint64_t *acc_full = hfull + 2;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);

  // ---- one-time set-up: R slice -> BF16, swizzled K-major; zero h tiles
  {
    const float *Rg = a.w_rec[dir];
    for (int idx = tid; idx < 128 * H; idx += kThreads) {
      const int k = idx % H, row = idx / H;
      const int u = row >> 2, g = row & 3;
      const float v = g < G ? Rg[((size_t)g * H + crank * UT + u) * H + k] : 0.f;
      const int kb = k >> 6, kk = k & 63;
      const uint32_t off = kb * 16384 + row * 128 + (((kk >> 3) ^ (row & 7)) << 4) + (kk & 7) * 2;
      *reinterpret_cast<__nv_bfloat16 *>(Rs + off) = __float2bfloat16_rn(v);
    }
    for (int idx = tid; idx < 2 * hs_bytes / 16; idx += kThreads)
      reinterpret_cast<uint4 *>(hs)[idx] = make_uint4(0, 0, 0, 0);
  }
  if (tid == 0) {
    mbar_init(hfull + 0, NC);
    mbar_init(hfull + 1, NC);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 32);
    tmem_relinquish();
  }
  fence_proxy_async_all();  // generic smem writes above -> visible to the tensor core (async proxy)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  cluster.sync();  // every CTA's tiles and barriers exist before anyone writes remotely

  if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = instr_desc(kFmtBF16, 0, 0, 128, NPAD);
    const uint32_t rs0 = smem_u32(Rs), hs0 = smem_u32(hs);
    for (int step = 0; step < T; step++) {
      const int p = step & 1;
      if (step > 0) {
        const int use = p ? (step - 1) >> 1 : (step >> 1) - 1;
        mbar_wait_cluster(hfull + p, use & 1);
      }
      fence_proxy_async_all();
      tc_fence_after();
      if (lane == 0) {
        const uint32_t hb = hs0 + p * hs_bytes;
        for (int kb = 0; kb < nkb; kb++) {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint64_t ad = smem_desc(rs0 + kb * 16384 + k * 32, 0, 1024, kLayoutSw128);
            const uint64_t bd = smem_desc(hb + kb * 2048 + k * 32, 0, 1024, kLayoutSw128);
            mma_bf16(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
          }
        }
        tc_commit(acc_full);
      }
      __syncwarp();
    }
  } else if (warp >= 2) {
    // ===================== epilogue =====================
    const int q = warp & 3;                  // TMEM lane quarter
    const int s = lane & 3;                  // gate slot of my row / batch slot after the transpose
    const int ul = q * 8 + (lane >> 2);      // local unit 0..31
    const int unit = crank * UT + ul;
    float *gates = a.gates[dir];
    float *cell = a.cell[dir];
    float brn = 0.f;
    if (MODE == 3) brn = a.b_rec[dir][2 * H + unit];
    float cst[NJ], hst[NJ];                  // cell / hidden state of (unit, batch 4j+s), fp32
#pragma unroll
    for (int j = 0; j < NJ; j++) cst[j] = hst[j] = 0.f;

    // which pre-activation column this ROW needs: LSTM gate s; GRU slots 0,1 -> r,z,
    // slot 3 carries the input part of n (slot 2 = recurrent part, nothing to load)
    const int pcol = MODE == 3 ? (s == 3 ? 2 : s) : s;
    const bool pload = MODE == 2 ? true : (MODE == 3 ? (s != 2) : (s == 0));
    float pre[BC];
    auto load_pre = [&](int step) {
      const int t = dir ? T - 1 - step : step;
      const float *pp = gates + ((size_t)t * B + b_lo) * GH + (size_t)pcol * H + unit;
#pragma unroll
      for (int b = 0; b < BC; b++) pre[b] = (pload && b < nb) ? pp[(size_t)b * GH] : 0.f;
    };
    load_pre(0);

    // remote addresses that never change
    const uint32_t hs0 = smem_u32(hs), hf0 = smem_u32(hfull);
    const int kb_mine = crank >> 1, chunk_mine = (crank & 1) * 4 + q;
    const int peer_a = lane >> 2, peer_b = (lane >> 2) + 8;  // the two peers this lane serves

    for (int step = 0; step < T; step++) {
      const int t = dir ? T - 1 - step : step;
      mbar_wait(acc_full, step & 1);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_32x16(tmem_base + ((uint32_t)(q * 32) << 16), r);
      tmem_ld_wait();
      tc_fence_before();

      float hnew[NJ];
#pragma unroll
      for (int j = 0; j < NJ; j++) {
        float x[4], g[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const float v = __uint_as_float(r[4 * j + e]) + pre[4 * j + e];
          if (MODE == 2) {
            const float sc = s == 2 ? 1.f : 0.5f;
            const float th = tanh_fast(sc * v);
            x[e] = s == 2 ? th : fmaf(0.5f, th, 0.5f);
          } else if (MODE == 3) {
            const float th = tanh_fast(0.5f * v);
            x[e] = s < 2 ? fmaf(0.5f, th, 0.5f) : (s == 2 ? v + brn : v);
          } else {
            x[e] = MODE == 0 ? fmaxf(v, 0.f) : tanh_fast(v);
          }
        }
        quad_transpose(x, g, s);
        const int b = 4 * j + s;
        const bool valid = b < nb;
        const size_t row = (size_t)t * B + b_lo + b;
        float h;
        if (MODE == 2) {
          cst[j] = fmaf(g[1], cst[j], g[0] * g[2]);
          h = g[3] * tanh_fast(cst[j]);
          if (valid && a.save) {
            float *gp = gates + row * GH + unit;
            gp[0] = g[0]; gp[H] = g[1]; gp[2 * H] = g[2]; gp[3 * H] = g[3];
            cell[row * H + unit] = cst[j];
          }
        } else if (MODE == 3) {
          const float n = tanh_fast(fmaf(g[0], g[2], g[3]));
          h = fmaf(g[1], hst[j] - n, n);  // (1-z) n + z h_prev
          if (valid && a.save) {
            float *gp = gates + row * GH + unit;
            gp[0] = g[0]; gp[H] = g[1]; gp[2 * H] = n;
            cell[row * H + unit] = g[2];
          }
        } else {
          h = g[0];
          if (valid && a.save) gates[row * GH + unit] = h;
        }
        if (!valid) h = 0.f;
        hst[j] = h;
        hnew[j] = h;
        if (valid) a.y[row * HO + dir * H + unit] = h;
      }

      if (step + 1 < T) {
        // ---- pack 8 consecutive units (one warp) of batch 4j+s into one 16-byte chunk
        const int pn = (step + 1) & 1;
#pragma unroll
        for (int j = 0; j < NJ; j++) {
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
          const int b = 4 * j + s;
          const uint32_t off = pn * hs_bytes + kb_mine * 2048 + b * 128 + ((chunk_mine ^ (b & 7)) << 4);
          if (peer_a < NC) st_cluster_v4(mapa_u32(hs0 + off, peer_a), ch);
          if (peer_b < NC) st_cluster_v4(mapa_u32(hs0 + off, peer_b), ch);
        }
        fence_proxy_async_all();
        load_pre(step + 1);
        epi_bar_sync();  // all 128 epilogue threads have issued their remote stores
        if (warp == 2 && lane < NC) mbar_arrive_remote(mapa_u32(hf0 + pn * 8, lane));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster.sync();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 32);
  }
}

template <typename K>
cudaError_t launch_cluster(K kernel, const RecArgs &a, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  if (a.NC > 8) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  const int nchunks = (a.B + a.BC - 1) / a.BC;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(a.NC, a.dirs * nchunks, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.NC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, a);
}

size_t fwd_smem_bytes(int H) { return 1024 + (size_t)(H / 64) * (16384 + 2 * 2048) + 64; }

template <int MODE>
cudaError_t launch_fwd(const RecArgs &a, cudaStream_t stream) {
  const size_t smem = fwd_smem_bytes(a.H);
  switch (a.BC) {
    case 4: return launch_cluster(rec_tc_fwd_kernel<MODE, 1>, a, smem, stream);
    case 8: return launch_cluster(rec_tc_fwd_kernel<MODE, 2>, a, smem, stream);
    default: return launch_cluster(rec_tc_fwd_kernel<MODE, 4>, a, smem, stream);
  }
}

}  // namespace

// The tcgen05 kernels need H a multiple of 64 (whole 128-byte swizzle rows of BF16)
// and at most 16 CTAs of 32 units per cluster.
bool rec_tc_supported(int mode, int H) {
  (void)mode;
  return H % 64 == 0 && H / UT <= 16 && fwd_smem_bytes(H) <= 227 * 1024;
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

}  // namespace b200
