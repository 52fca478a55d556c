// kaldi_ctc_b200/csrc/rnn_rec_fp32.cu -- persistent recurrent kernels, fp32 math.
//
// One thread-block CLUSTER per (direction, batch chunk of <= 16 utterances).
// The recurrent weight matrix R [G*H x H] of that direction is split by hidden
// unit over the NC CTAs of the cluster and stays resident in shared memory for
// all T steps (fp32: 4*G*U*H bytes per CTA, U = H/NC).  Each step a CTA computes
// the gates of ITS U units for all utterances of the chunk, applies the cell
// update (cell state lives in registers), and broadcasts its slice of h_t into
// every CTA's double-buffered h tile through distributed shared memory; one
// cluster barrier per step.  The backward kernel keeps the same slice, forms the
// partial products R_slice^T . dgates (split-K over gate rows) and reduce-scatters
// them over DSMEM.
//
// This is the exact mode (B200RNN_MATH_FP32: matches the oracle to ~1e-6); the
// tensor-core mode lives in rnn_rec_tc.cu.
#include <cooperative_groups.h>

#include "rnn_common.cuh"

namespace cg = cooperative_groups;

namespace b200 {
namespace {

constexpr int BCP = 16;          // batch pitch of the smem tiles
constexpr int FWD_THREADS = 256; // (unit, batch pair) items
constexpr int BWD_THREADS = 320;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }
// precise enough for 1e-6: expf-based
__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// ===========================================================================
// forward
// ===========================================================================
template <int MODE>
__global__ void __launch_bounds__(FWD_THREADS, 1) rec_fwd_kernel(RecArgs a) {
  constexpr int G = MODE == 2 ? 4 : (MODE == 3 ? 3 : 1);
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int dir = blockIdx.y % a.dirs, chunk = blockIdx.y / a.dirs;
  const int b_lo = chunk * a.BC, nb = min(a.BC, a.B - b_lo);
  const int H = a.H, U = a.U, GU = G * U, T = a.T, B = a.B, GH = G * H, HO = H * a.dirs;
  const int tid = threadIdx.x;

  extern __shared__ __align__(16) float smem[];
  float *Rt = smem;                 // [H][GU]  Rt[k][u*G+g] = R_g[crank*U+u][k]
  float *ht = Rt + (size_t)H * GU;  // [2][H][BCP]

  const float *Rg = a.w_rec[dir];
  for (int idx = tid; idx < GU * H; idx += blockDim.x) {
    const int k = idx % H, row = idx / H;  // row = g*U + u (coalesced global reads over k)
    const int g = row / U, u = row % U;
    Rt[(size_t)k * GU + u * G + g] = Rg[((size_t)g * H + crank * U + u) * H + k];
  }
  for (int idx = tid; idx < 2 * H * BCP; idx += blockDim.x) ht[idx] = 0.f;
  cluster.sync();

  // item: unit u, batch pair b0,b0+1
  const int u = tid / (BCP / 2), b0 = (tid % (BCP / 2)) * 2;
  const bool act0 = u < U && b0 < nb, act1 = u < U && b0 + 1 < nb;
  const int unit = crank * U + u;
  float brn = 0.f;
  if (MODE == 3 && u < U) brn = a.b_rec[dir][2 * H + unit];
  float c0 = 0.f, c1 = 0.f;  // LSTM cell state
  float *gates = a.gates[dir];
  float *cell = a.cell[dir];

  float pre0[G], pre1[G];
  auto load_pre = [&](int step) {
    const int t = dir ? T - 1 - step : step;
    const size_t r0 = ((size_t)t * B + b_lo + b0) * GH + unit;
#pragma unroll
    for (int g = 0; g < G; g++) {
      pre0[g] = act0 ? gates[r0 + (size_t)g * H] : 0.f;
      pre1[g] = act1 ? gates[r0 + GH + (size_t)g * H] : 0.f;
    }
  };
  load_pre(0);

  for (int step = 0; step < T; step++) {
    const int t = dir ? T - 1 - step : step;
    const float *hc = ht + (size_t)(step & 1) * H * BCP;
    float acc0[G], acc1[G];
#pragma unroll
    for (int g = 0; g < G; g++) acc0[g] = acc1[g] = 0.f;
    if (u < U) {
      const float *rp = Rt + u * G;
      const float *hp = hc + b0;
#pragma unroll 8
      for (int k = 0; k < H; k++) {
        const float2 hv = *reinterpret_cast<const float2 *>(hp + (size_t)k * BCP);
        if (G == 4) {
          const float4 w = *reinterpret_cast<const float4 *>(rp + (size_t)k * GU);
          acc0[0] = fmaf(w.x, hv.x, acc0[0]); acc1[0] = fmaf(w.x, hv.y, acc1[0]);
          acc0[1 % G] = fmaf(w.y, hv.x, acc0[1 % G]); acc1[1 % G] = fmaf(w.y, hv.y, acc1[1 % G]);
          acc0[2 % G] = fmaf(w.z, hv.x, acc0[2 % G]); acc1[2 % G] = fmaf(w.z, hv.y, acc1[2 % G]);
          acc0[3 % G] = fmaf(w.w, hv.x, acc0[3 % G]); acc1[3 % G] = fmaf(w.w, hv.y, acc1[3 % G]);
        } else {
#pragma unroll
          for (int g = 0; g < G; g++) {
            const float w = rp[(size_t)k * GU + g];
            acc0[g] = fmaf(w, hv.x, acc0[g]);
            acc1[g] = fmaf(w, hv.y, acc1[g]);
          }
        }
      }
    }
    // previous h of this unit (GRU) before the tile is reused
    float hp0 = 0.f, hp1 = 0.f;
    if (MODE == 3 && u < U) {
      hp0 = hc[(size_t)unit * BCP + b0];
      hp1 = hc[(size_t)unit * BCP + b0 + 1];
    }
    float h0 = 0.f, h1 = 0.f;
    float out0[G], out1[G], cs0 = 0.f, cs1 = 0.f;
    if (MODE == 2) {
      const float i0 = sigm(pre0[0] + acc0[0]), i1 = sigm(pre1[0] + acc1[0]);
      const float f0 = sigm(pre0[1 % G] + acc0[1 % G]), f1 = sigm(pre1[1 % G] + acc1[1 % G]);
      const float g0 = tanhf(pre0[2 % G] + acc0[2 % G]), g1 = tanhf(pre1[2 % G] + acc1[2 % G]);
      const float o0 = sigm(pre0[3 % G] + acc0[3 % G]), o1 = sigm(pre1[3 % G] + acc1[3 % G]);
      c0 = f0 * c0 + i0 * g0;
      c1 = f1 * c1 + i1 * g1;
      h0 = o0 * tanhf(c0);
      h1 = o1 * tanhf(c1);
      out0[0] = i0; out0[1 % G] = f0; out0[2 % G] = g0; out0[3 % G] = o0;
      out1[0] = i1; out1[1 % G] = f1; out1[2 % G] = g1; out1[3 % G] = o1;
      cs0 = c0; cs1 = c1;
    } else if (MODE == 3) {
      const float r0 = sigm(pre0[0] + acc0[0]), r1 = sigm(pre1[0] + acc1[0]);
      const float z0 = sigm(pre0[1 % G] + acc0[1 % G]), z1 = sigm(pre1[1 % G] + acc1[1 % G]);
      const float q0 = acc0[2 % G] + brn, q1 = acc1[2 % G] + brn;
      const float n0 = tanhf(pre0[2 % G] + r0 * q0), n1 = tanhf(pre1[2 % G] + r1 * q1);
      h0 = (1.f - z0) * n0 + z0 * hp0;
      h1 = (1.f - z1) * n1 + z1 * hp1;
      out0[0] = r0; out0[1 % G] = z0; out0[2 % G] = n0;
      out1[0] = r1; out1[1 % G] = z1; out1[2 % G] = n1;
      cs0 = q0; cs1 = q1;
    } else {
      const float v0 = pre0[0] + acc0[0], v1 = pre1[0] + acc1[0];
      h0 = MODE == 0 ? fmaxf(v0, 0.f) : tanhf(v0);
      h1 = MODE == 0 ? fmaxf(v1, 0.f) : tanhf(v1);
      out0[0] = h0; out1[0] = h1;
    }
    if (!act0) h0 = 0.f;
    if (!act1) h1 = 0.f;
    // global stores
    {
      const size_t row = (size_t)t * B + b_lo + b0;
      if (act0) a.y[row * HO + dir * H + unit] = h0;
      if (act1) a.y[(row + 1) * HO + dir * H + unit] = h1;
      if (a.save) {
#pragma unroll
        for (int g = 0; g < G; g++) {
          if (act0) gates[row * GH + (size_t)g * H + unit] = out0[g];
          if (act1) gates[(row + 1) * GH + (size_t)g * H + unit] = out1[g];
        }
        if (MODE >= 2) {
          if (act0) cell[row * H + unit] = cs0;
          if (act1) cell[(row + 1) * H + unit] = cs1;
        }
      }
    }
    if (step + 1 < T) load_pre(step + 1);
    // broadcast h_t slice into every CTA's next tile
    if (u < U) {
      const size_t off = (size_t)((step + 1) & 1) * H * BCP + (size_t)unit * BCP + b0;
      const float2 hv = make_float2(h0, h1);
      for (int q = 0; q < a.NC; q++) {
        float *dst = cluster.map_shared_rank(ht, q);
        *reinterpret_cast<float2 *>(dst + off) = hv;
      }
    }
    cluster.sync();
  }
}

// ===========================================================================
// backward (data): dgates for all steps + dh recurrence
// ===========================================================================
template <int MODE>
__global__ void __launch_bounds__(BWD_THREADS, 1) rec_bwd_kernel(RecArgs a) {
  constexpr int G = MODE == 2 ? 4 : (MODE == 3 ? 3 : 1);
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int dir = blockIdx.y % a.dirs, chunk = blockIdx.y / a.dirs;
  const int b_lo = chunk * a.BC, nb = min(a.BC, a.B - b_lo);
  const int H = a.H, U = a.U, GU = G * U, T = a.T, B = a.B, GH = G * H, HO = H * a.dirs, NC = a.NC;
  const int tid = threadIdx.x;

  extern __shared__ __align__(16) float smem[];
  float *Rb = smem;                       // [GU][H]   row' = u*G+g
  float *dg = Rb + (size_t)GU * H;        // [GU][BCP] recurrent-side gate gradients of this step
  float *recv = dg + (size_t)GU * BCP;    // [2][NC][BCP][U] partial dh from every CTA
  float *st = recv + (size_t)2 * NC * BCP * U;  // [U][BCP] carried state: LSTM dc, GRU dh*z

  const float *Rg = a.w_rec[dir];
  for (int idx = tid; idx < GU * H; idx += blockDim.x) {
    const int k = idx % H, row = idx / H;
    const int g = row / U, u = row % U;
    Rb[(size_t)(u * G + g) * H + k] = Rg[((size_t)g * H + crank * U + u) * H + k];
  }
  for (int idx = tid; idx < 2 * NC * BCP * U; idx += blockDim.x) recv[idx] = 0.f;
  for (int idx = tid; idx < U * BCP; idx += blockDim.x) st[idx] = 0.f;
  for (int idx = tid; idx < GU * BCP; idx += blockDim.x) dg[idx] = 0.f;
  cluster.sync();

  float *gates = a.gates[dir];
  float *cell = a.cell[dir];

  for (int step = 0; step < T; step++) {
    // forward visited t in the order dir ? T-1..0 : 0..T-1; go back the other way
    const int fstep = T - 1 - step;             // forward step index of this frame
    const int t = dir ? T - 1 - fstep : fstep;  // frame
    const int tp = dir ? t + 1 : t - 1;         // frame of h_{prev}/c_{prev}
    const bool first = fstep == 0;
    const float *rc = recv + (size_t)(step & 1) * NC * BCP * U;

    // ---- phase A: gate gradients of my units
    for (int it = tid; it < U * BCP; it += blockDim.x) {
      const int u = it / BCP, b = it % BCP;
      if (b >= nb) continue;
      const int unit = crank * U + u;
      const size_t row = (size_t)t * B + b_lo + b;
      float dh = a.dy[row * HO + dir * H + unit];
      for (int q = 0; q < NC; q++) dh += rc[((size_t)q * BCP + b) * U + u];
      float *gp = gates + row * GH + unit;
      if (MODE == 2) {
        const float i = gp[0], f = gp[H], g_ = gp[2 * H], o = gp[3 * H];
        const float c = cell[row * H + unit];
        const float cp = first ? 0.f : cell[((size_t)tp * B + b_lo + b) * H + unit];
        const float tc = tanhf(c);
        const float dc = dh * o * (1.f - tc * tc) + st[u * BCP + b];
        const float di = dc * g_ * i * (1.f - i);
        const float df = dc * cp * f * (1.f - f);
        const float dgg = dc * i * (1.f - g_ * g_);
        const float dob = dh * tc * o * (1.f - o);
        st[u * BCP + b] = dc * f;
        gp[0] = di; gp[H] = df; gp[2 * H] = dgg; gp[3 * H] = dob;
        dg[(u * G + 0) * BCP + b] = di;
        dg[(u * G + 1 % G) * BCP + b] = df;
        dg[(u * G + 2 % G) * BCP + b] = dgg;
        dg[(u * G + 3 % G) * BCP + b] = dob;
      } else if (MODE == 3) {
        dh += st[u * BCP + b];
        const float r = gp[0], z = gp[H], n = gp[2 * H];
        const float q = cell[row * H + unit];
        const float hp = first ? 0.f : a.y[((size_t)tp * B + b_lo + b) * HO + dir * H + unit];
        const float dn = dh * (1.f - z) * (1.f - n * n);
        const float dr = dn * q * r * (1.f - r);
        const float dz = dh * (hp - n) * z * (1.f - z);
        st[u * BCP + b] = dh * z;
        gp[0] = dr; gp[H] = dz; gp[2 * H] = dn;   // input side
        cell[row * H + unit] = dn * r;            // recurrent side of the n gate (dq)
        dg[(u * G + 0) * BCP + b] = dr;
        dg[(u * G + 1 % G) * BCP + b] = dz;
        dg[(u * G + 2 % G) * BCP + b] = dn * r;
      } else {
        const float h = gp[0];
        const float d = dh * (MODE == 0 ? (h > 0.f ? 1.f : 0.f) : (1.f - h * h));
        gp[0] = d;
        dg[(u * G) * BCP + b] = d;
      }
    }
    __syncthreads();

    // ---- phase B: partial dh_prev[b][k] = sum_{row'} Rb[row'][k] * dg[row'][b], all k
    if (step + 1 < T) {
      float *rn = recv + (size_t)((step + 1) & 1) * NC * BCP * U;
      for (int it = tid; it < (H / 4) * (BCP / 4); it += blockDim.x) {
        const int k4 = it % (H / 4), bq = it / (H / 4);
        if (bq * 4 >= nb) continue;
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
#pragma unroll 4
        for (int r = 0; r < GU; r++) {
          const float4 w = *reinterpret_cast<const float4 *>(Rb + (size_t)r * H + k4 * 4);
          const float4 d = *reinterpret_cast<const float4 *>(dg + (size_t)r * BCP + bq * 4);
          const float wv[4] = {w.x, w.y, w.z, w.w}, dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < 4; j++) acc[i][j] = fmaf(wv[i], dv[j], acc[i][j]);
        }
        const int k = k4 * 4, q = k / U, ku = k - q * U;  // 4 | U: the four k share a target
        float *dst = cluster.map_shared_rank(rn, q) + ((size_t)crank * BCP + bq * 4) * U + ku;
#pragma unroll
        for (int j = 0; j < 4; j++)
          *reinterpret_cast<float4 *>(dst + (size_t)j * U) =
              make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
      }
    }
    cluster.sync();
  }
}

template <typename K>
cudaError_t launch_cluster(K kernel, const RecArgs &a, int threads, size_t smem, cudaStream_t stream) {
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
  return cudaLaunchKernelEx(&cfg, kernel, a);
}

}  // namespace

size_t rec_fp32_smem_bytes(int mode, int H, int NC, bool backward) {
  if (NC < 1 || H % NC) return 0;
  const int G = gates_of(mode), U = H / NC;
  if (U % 4) return 0;
  size_t fl;
  if (!backward) {
    if (U * (BCP / 2) > FWD_THREADS) return 0;
    fl = (size_t)H * G * U + 2 * (size_t)H * BCP;
  } else {
    fl = (size_t)G * U * H + (size_t)G * U * BCP + 2 * (size_t)NC * BCP * U + (size_t)U * BCP;
  }
  const size_t bytes = fl * sizeof(float);
  return bytes <= 227 * 1024 ? bytes : 0;
}

int rec_fp32_pick_cluster(int mode, int H) {
  static const int cand[] = {16, 10, 8, 12, 14, 6, 5, 4, 2, 1};
  for (int NC : cand) {
    const size_t sf = rec_fp32_smem_bytes(mode, H, NC, false), sb = rec_fp32_smem_bytes(mode, H, NC, true);
    if (!sf || !sb) continue;
    // can the device co-schedule such a cluster?
    int ok = 1;
    for (int pass = 0; pass < 2 && ok; pass++) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(NC, 1, 1);
      cfg.blockDim = dim3(pass ? BWD_THREADS : FWD_THREADS, 1, 1);
      cfg.dynamicSmemBytes = pass ? sb : sf;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = NC;
      attr[0].val.clusterDim.y = 1;
      attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      int n = 0;
      const void *fn = pass ? (const void *)rec_bwd_kernel<2> : (const void *)rec_fwd_kernel<2>;
      if (mode == 3) fn = pass ? (const void *)rec_bwd_kernel<3> : (const void *)rec_fwd_kernel<3>;
      if (mode == 0) fn = pass ? (const void *)rec_bwd_kernel<0> : (const void *)rec_fwd_kernel<0>;
      if (mode == 1) fn = pass ? (const void *)rec_bwd_kernel<1> : (const void *)rec_fwd_kernel<1>;
      cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.dynamicSmemBytes);
      if (NC > 8) cudaFuncSetAttribute(fn, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      if (cudaOccupancyMaxActiveClusters(&n, fn, &cfg) != cudaSuccess || n < 1) ok = 0;
    }
    cudaGetLastError();  // clear any probing error
    if (ok) return NC;
  }
  return 0;
}

cudaError_t rec_fp32_forward(const RecArgs &a, cudaStream_t stream) {
  const size_t smem = rec_fp32_smem_bytes(a.mode, a.H, a.NC, false);
  if (!smem) return cudaErrorInvalidValue;
  switch (a.mode) {
    case 0: return launch_cluster(rec_fwd_kernel<0>, a, FWD_THREADS, smem, stream);
    case 1: return launch_cluster(rec_fwd_kernel<1>, a, FWD_THREADS, smem, stream);
    case 2: return launch_cluster(rec_fwd_kernel<2>, a, FWD_THREADS, smem, stream);
    default: return launch_cluster(rec_fwd_kernel<3>, a, FWD_THREADS, smem, stream);
  }
}

cudaError_t rec_fp32_backward(const RecArgs &a, cudaStream_t stream) {
  const size_t smem = rec_fp32_smem_bytes(a.mode, a.H, a.NC, true);
  if (!smem) return cudaErrorInvalidValue;
  switch (a.mode) {
    case 0: return launch_cluster(rec_bwd_kernel<0>, a, BWD_THREADS, smem, stream);
    case 1: return launch_cluster(rec_bwd_kernel<1>, a, BWD_THREADS, smem, stream);
    case 2: return launch_cluster(rec_bwd_kernel<2>, a, BWD_THREADS, smem, stream);
    default: return launch_cluster(rec_bwd_kernel<3>, a, BWD_THREADS, smem, stream);
  }
}

}  // namespace b200
