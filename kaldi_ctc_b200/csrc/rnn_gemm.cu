// kaldi_ctc_b200/csrc/rnn_gemm.cu -- fp32 (CUDA-core FMA) GEMM with generic
// operand strides, deterministic split-K, fused bias; and deterministic column
// sums.  Used by the MATH_FP32 mode for the hoisted input projection
// (pre = x.Wi^T + bW + bR), the input gradient (dx = dG.Wi), the weight
// gradients (dWi = dG^T.x, dR = dG^T.h_prev) and the adjacent affine layer.
#include "rnn_common.cuh"

namespace b200 {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, LDS_ = 132;

// AKC: A is K-contiguous (sak == 1), else M-contiguous (sam == 1) or generic.
template <bool AKC, bool BKC>
__global__ void __launch_bounds__(256) sgemm_kernel(GemmArgs g, int k_per_split) {
  __shared__ __align__(16) float As[2][BK][LDS_];
  __shared__ __align__(16) float Bs[2][BK][LDS_];
  const int tid = threadIdx.x;
  const int bm = blockIdx.y * BM, bn = blockIdx.x * BN;
  const int k_lo = blockIdx.z * k_per_split;
  const int k_hi = min(g.K, k_lo + k_per_split);
  const int tx = tid & 15, ty = tid >> 4;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 8; j++) acc[i][j] = 0.f;

  float ra[8], rb[8];
  auto load_tile = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int e = tid + i * 256;
      int m, k;
      if (AKC) { k = e & 15; m = e >> 4; } else { m = e & 127; k = e >> 7; }
      const int gm = bm + m, gk = k0 + k;
      ra[i] = (gm < g.M && gk < k_hi) ? __ldg(g.A + (long long)gm * g.sam + (long long)gk * g.sak) : 0.f;
      int n, kb;
      if (BKC) { kb = e & 15; n = e >> 4; } else { n = e & 127; kb = e >> 7; }
      const int gn = bn + n, gkb = k0 + kb;
      rb[i] = (gn < g.N && gkb < k_hi) ? __ldg(g.B + (long long)gkb * g.sbk + (long long)gn * g.sbn) : 0.f;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      const int e = tid + i * 256;
      int m, k;
      if (AKC) { k = e & 15; m = e >> 4; } else { m = e & 127; k = e >> 7; }
      As[buf][k][m] = ra[i];
      int n, kb;
      if (BKC) { kb = e & 15; n = e >> 4; } else { n = e & 127; kb = e >> 7; }
      Bs[buf][kb][n] = rb[i];
    }
  };

  if (k_lo < k_hi) {
    load_tile(k_lo);
    store_tile(0);
    __syncthreads();
    int buf = 0;
    for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
      const bool more = k0 + BK < k_hi;
      if (more) load_tile(k0 + BK);
#pragma unroll
      for (int k = 0; k < BK; k++) {
        const float4 a0 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4 *>(&As[buf][k][ty * 4 + 64]);
        const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[buf][k][tx * 4 + 64]);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; i++)
#pragma unroll
          for (int j = 0; j < 8; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      if (more) {
        store_tile(buf ^ 1);
        __syncthreads();
        buf ^= 1;
      }
    }
  }

  const bool direct = g.splits <= 1;
  float *out = direct ? g.C : g.partial + (size_t)blockIdx.z * g.M * g.N;
  const int ldo = direct ? g.ldc : g.N;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const int m = bm + ty * 4 + (i & 3) + (i >> 2) * 64;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int n = bn + tx * 4 + (j & 3) + (j >> 2) * 64;
      if (n >= g.N) continue;
      float v = g.alpha * acc[i][j];
      if (direct) {
        if (g.beta != 0.f) v += g.beta * out[(size_t)m * ldo + n];
        if (g.bias_a) v += g.bias_a[n];
        if (g.bias_b && n < g.nb) v += g.bias_b[n];
      }
      out[(size_t)m * ldo + n] = v;
    }
  }
}

__global__ void splitk_reduce_kernel(GemmArgs g) {
  const size_t total = (size_t)g.M * g.N;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / g.N), n = (int)(i % g.N);
    float s = 0.f;
    for (int z = 0; z < g.splits; z++) s += g.partial[(size_t)z * total + i];
    float *c = g.C + (size_t)m * g.ldc + n;
    if (g.beta != 0.f) s += g.beta * *c;
    if (g.bias_a) s += g.bias_a[n];
    if (g.bias_b && n < g.nb) s += g.bias_b[n];
    *c = s;
  }
}

constexpr int kColSplit = 64;

// partial[z][c] = sum over the z-th row range
__global__ void colsum_partial_kernel(const float *a, int rows, int cols, int lda, float *partial) {
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  const int rg = threadIdx.x >> 5;  // 8 row groups per block
  const int per = (rows + kColSplit - 1) / kColSplit;
  const int r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  if (c < cols)
    for (int r = r0 + rg; r < r1; r += 8) s += a[(size_t)r * lda + c];
  __shared__ float sm[8][33];
  sm[rg][threadIdx.x & 31] = s;
  __syncthreads();
  if (rg == 0 && c < cols) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += sm[i][threadIdx.x & 31];
    partial[(size_t)blockIdx.y * cols + c] = t;
  }
}
__global__ void colsum_final_kernel(const float *partial, int cols, float alpha, float *out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float s = 0.f;
  for (int z = 0; z < kColSplit; z++) s += partial[(size_t)z * cols + c];
  out[c] = accumulate ? fmaf(alpha, s, out[c]) : alpha * s;
}

}  // namespace

cudaError_t gemm_fp32(const GemmArgs &g0, cudaStream_t stream, int *launches) {
  GemmArgs g = g0;
  if (g.M <= 0 || g.N <= 0) return cudaSuccess;
  if (g.splits < 1 || !g.partial) g.splits = 1;
  const int kps = (((g.K + g.splits - 1) / g.splits) + BK - 1) / BK * BK;
  g.splits = g.K > 0 ? (g.K + kps - 1) / kps : 1;
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.splits);
  const bool akc = g.sak == 1, bkc = g.sbk == 1;
  if (akc && bkc) sgemm_kernel<true, true><<<grid, 256, 0, stream>>>(g, kps);
  else if (akc) sgemm_kernel<true, false><<<grid, 256, 0, stream>>>(g, kps);
  else if (bkc) sgemm_kernel<false, true><<<grid, 256, 0, stream>>>(g, kps);
  else sgemm_kernel<false, false><<<grid, 256, 0, stream>>>(g, kps);
  if (launches) (*launches)++;
  if (g.splits > 1) {
    if (launches) (*launches)++;
    return splitk_reduce(g, stream);
  }
  return cudaGetLastError();
}

cudaError_t splitk_reduce(const GemmArgs &g, cudaStream_t stream) {
  const size_t total = (size_t)g.M * g.N;
  const size_t blocks = (total + 255) / 256;
  splitk_reduce_kernel<<<(unsigned)(blocks > 1184 ? 1184 : blocks), 256, 0, stream>>>(g);
  return cudaGetLastError();
}

int g_last_gemm_tc = 0;

cudaError_t gemm_any(int math, const GemmArgs &g, cudaStream_t stream, int *launches) {
  g_last_gemm_tc = 0;
  if (g.A2 && g.B2) {
    // two products into one C: in one launch where the CTA-pair kernel applies, else one after the other
    if (math == 1) {
      cudaError_t e = gemm_tc(g, stream, launches);
      if (e != cudaErrorNotSupported) {
        g_last_gemm_tc = 1;
        return e;
      }
      cudaGetLastError();
    }
    GemmArgs g1 = g;
    g1.A2 = g1.B2 = nullptr;
    cudaError_t e = gemm_any(math, g1, stream, launches);
    if (e != cudaSuccess) return e;
    g1.A = g.A2;
    g1.B = g.B2;
    g1.beta = 1.f;
    g1.bias_a = g1.bias_b = nullptr;
    return gemm_any(math, g1, stream, launches);
  }
  if (math == 1) {
    cudaError_t e = gemm_tc(g, stream, launches);
    if (e != cudaErrorNotSupported) {
      g_last_gemm_tc = 1;
      return e;
    }
    cudaGetLastError();
  }
  return gemm_fp32(g, stream, launches);
}

size_t column_sums_partial_floats(int rows, int cols) { return (size_t)kColSplit * cols; }

cudaError_t column_sums(const float *a, int rows, int cols, int lda, float alpha, float *out, int accumulate,
                        float *partial, size_t partial_floats, cudaStream_t stream, int *launches) {
  if (partial_floats < (size_t)kColSplit * cols) return cudaErrorInvalidValue;
  dim3 grid((cols + 31) / 32, kColSplit);
  colsum_partial_kernel<<<grid, 256, 0, stream>>>(a, rows, cols, lda, partial);
  colsum_final_kernel<<<(cols + 255) / 256, 256, 0, stream>>>(partial, cols, alpha, out, accumulate);
  if (launches) (*launches) += 2;
  return cudaGetLastError();
}

}  // namespace b200
