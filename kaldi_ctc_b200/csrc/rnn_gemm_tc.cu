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
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA
// issuer (one elected lane), warps 2-5 = epilogue (TMEM -> registers -> global).
// 3-stage smem ring (96 KB) and 128 TMEM columns per CTA: two CTAs share an SM,
// so one tile's epilogue overlaps the other's main loop.
#include <stdio.h>

#include "rnn_common.cuh"
#include "tc_common.cuh"

namespace b200 {
namespace {

using namespace tc;

constexpr int TBM = 128, TBN = 128, TBK = 32;  // tf32: 32 elements = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kTileBytes = TBM * TBK * 4;       // 16 KB per operand per stage
constexpr int kThreads = 192;
constexpr int kTmemCols = 128;

struct TcParams {
  int M, N, K;
  float alpha, beta;
  float *C;
  int ldc;
  const float *bias_a, *bias_b;
  int nb;
  int splits, kb_per_split;
  float *partial;
};

template <bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(kThreads, 2)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, TcParams p) {
  extern __shared__ uint8_t smem_dyn[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
  uint8_t *sA = smem;                         // [kStages][16 KB]
  uint8_t *sB = smem + kStages * kTileBytes;  // [kStages][16 KB]
  uint64_t *full = reinterpret_cast<uint64_t *>(smem + 2 * kStages * kTileBytes);
  uint64_t *empty = full + kStages;
  uint64_t *tmem_full = empty + kStages;
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * TBM, n0 = blockIdx.x * TBN;
  const int nkb = (p.K + TBK - 1) / TBK;
  const int kb0 = blockIdx.z * p.kb_per_split;
  const int kb1 = min(nkb, kb0 + p.kb_per_split);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < kStages; s++) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(tmem_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = kb0; kb < kb1; kb++) {
        const int it = kb - kb0, s = it % kStages;
        mbar_wait(empty + s, ((it / kStages) & 1) ^ 1);
        mbar_expect_tx(full + s, 2 * kTileBytes);
        uint8_t *a = sA + s * kTileBytes, *b = sB + s * kTileBytes;
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
  } else if (warp == 1) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = instr_desc(kFmtTF32, A_KMAJOR ? 0 : 1, B_KMAJOR ? 0 : 1, TBM, TBN);
    for (int kb = kb0; kb < kb1; kb++) {
      const int it = kb - kb0, s = it % kStages;
      mbar_wait(full + s, (it / kStages) & 1);
      tc_fence_after();
      {
        // warp-uniform issue (descriptors stay in uniform registers), one elected lane issues
        const uint32_t a = smem_u32(sA + s * kTileBytes), b = smem_u32(sB + s * kTileBytes);
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
          if (elect_one()) mma_tf32(tmem_base, ad, bd, idesc, (it > 0 || k > 0) ? 1u : 0u);
        }
        if (elect_one()) {
          tc_commit(empty + s);                   // frees the smem stage when these MMAs retire
          if (kb == kb1 - 1) tc_commit(tmem_full);  // accumulator complete
        }
      }
      __syncwarp();
    }
  } else {
    // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int m = m0 + q * 32 + lane;
    const bool direct = p.splits <= 1;
    float *out = direct ? p.C : p.partial + (size_t)blockIdx.z * p.M * p.N;
    const int ldo = direct ? p.ldc : p.N;
    const bool have = kb1 > kb0;
    if (have) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    const bool vec_ok = (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
#pragma unroll 1
    for (int c = 0; c < TBN / 32; c++) {
      uint32_t r[32];
      if (have) {
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + c * 32, r);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int j = 0; j < 32; j++) r[j] = 0u;
      }
      if (m < p.M) {
        const int nb0 = n0 + c * 32;
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
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
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

template <bool AK, bool BK>
cudaError_t launch(const CUtensorMap &ta, const CUtensorMap &tb, const TcParams &p, dim3 grid, cudaStream_t s) {
  const size_t smem = 1024 + 2 * kStages * kTileBytes + 128;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<AK, BK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  tc_gemm_kernel<AK, BK><<<grid, kThreads, smem, s>>>(ta, tb, p);
  return cudaGetLastError();
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
  CUtensorMap ta, tb;
  const CUtensorMapSwizzle kmaj = CU_TENSOR_MAP_SWIZZLE_128B, mnmaj = CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
  bool ok = ak ? make_map(&ta, g.A, g.K, g.M, g.sam, TBK, TBM, kmaj)
               : make_map(&ta, g.A, g.M, g.K, g.sak, 32, TBK, mnmaj);
  ok = ok && (bk ? make_map(&tb, g.B, g.K, g.N, g.sbn, TBK, TBN, kmaj)
                 : make_map(&tb, g.B, g.N, g.K, g.sbk, 32, TBK, mnmaj));
  if (!ok) return cudaErrorNotSupported;

  TcParams p;
  p.M = g.M; p.N = g.N; p.K = g.K; p.alpha = g.alpha; p.beta = g.beta;
  p.C = g.C; p.ldc = g.ldc; p.bias_a = g.bias_a; p.bias_b = g.bias_b; p.nb = g.nb;
  const int nkb = (g.K + TBK - 1) / TBK;
  int splits = (g.splits > 1 && g.partial) ? g.splits : 1;
  p.kb_per_split = (nkb + splits - 1) / splits;
  splits = (nkb + p.kb_per_split - 1) / p.kb_per_split;
  p.splits = splits;
  p.partial = g.partial;
  dim3 grid((g.N + TBN - 1) / TBN, (g.M + TBM - 1) / TBM, splits);
  cudaError_t e;
  if (ak && bk) e = launch<true, true>(ta, tb, p, grid, stream);
  else if (ak) e = launch<true, false>(ta, tb, p, grid, stream);
  else if (bk) e = launch<false, true>(ta, tb, p, grid, stream);
  else e = launch<false, false>(ta, tb, p, grid, stream);
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
