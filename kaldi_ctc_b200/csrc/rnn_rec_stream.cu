// kaldi_ctc_b200/csrc/rnn_rec_stream.cu -- general recurrent path: one (or two) kernel launches
// per time step, recurrent weights streamed from L2 every step.  It has no shape restriction
// (any hidden size, any minibatch, all four modes) and exists for the configurations the
// persistent cluster kernels cannot hold on chip (H not a multiple of 4, or a weight slice larger
// than shared/tensor memory); exact fp32 arithmetic.  The benchmark shapes never take this path.
#include "rnn_common.cuh"

namespace b200 {
namespace {

constexpr int kWarps = 8;

__device__ __forceinline__ float sigm(float x) { return 1.0f / (1.0f + expf(-x)); }

// forward, one time step: one warp per (direction, hidden unit); lanes split K, all utterances
// of a chunk of <= 16 are accumulated per lane, then reduced by shuffles.
template <int MODE>
__global__ void __launch_bounds__(kWarps * 32) stream_fwd_step(RecArgs a, int step) {
  constexpr int G = MODE == 2 ? 4 : (MODE == 3 ? 3 : 1);
  constexpr int BCH = 8;
  const int lane = threadIdx.x & 31;
  const int unit = blockIdx.x * kWarps + (threadIdx.x >> 5);
  const int dir = blockIdx.y;
  const int H = a.H, B = a.B, T = a.T, GH = G * H, HO = H * a.dirs;
  if (unit >= H) return;
  const int t = dir ? T - 1 - step : step;
  const int tp = dir ? t + 1 : t - 1;
  const float *R = a.w_rec[dir];
  float *gates = a.gates[dir];
  float *cell = a.cell[dir];
  const float brn = MODE == 3 ? a.b_rec[dir][2 * H + unit] : 0.f;
  for (int b0 = 0; b0 < B; b0 += BCH) {
    float acc[G][BCH];
#pragma unroll
    for (int g = 0; g < G; g++)
#pragma unroll
      for (int j = 0; j < BCH; j++) acc[g][j] = 0.f;
    if (step > 0) {
      for (int k = lane; k < H; k += 32) {
        float w[G];
#pragma unroll
        for (int g = 0; g < G; g++) w[g] = R[((size_t)g * H + unit) * H + k];
#pragma unroll
        for (int j = 0; j < BCH; j++) {
          if (b0 + j < B) {
            const float h = a.y[((size_t)tp * B + b0 + j) * HO + dir * H + k];
#pragma unroll
            for (int g = 0; g < G; g++) acc[g][j] = fmaf(w[g], h, acc[g][j]);
          }
        }
      }
#pragma unroll
      for (int g = 0; g < G; g++)
#pragma unroll
        for (int j = 0; j < BCH; j++)
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc[g][j] += __shfl_xor_sync(0xffffffffu, acc[g][j], o);
    }
    // lane j finishes utterance b0 + j
#pragma unroll
    for (int j = 0; j < BCH; j++) {
      if (lane == j && b0 + j < B) {
        const size_t row = (size_t)t * B + b0 + j;
        float *gp = gates + row * GH + unit;
        float h;
        if (MODE == 2) {
          const float i = sigm(gp[0] + acc[0][j]), f = sigm(gp[H] + acc[1 % G][j]);
          const float g_ = tanhf(gp[2 * H] + acc[2 % G][j]), o = sigm(gp[3 * H] + acc[3 % G][j]);
          const float cp = step > 0 ? cell[((size_t)tp * B + b0 + j) * H + unit] : 0.f;
          const float c = f * cp + i * g_;
          h = o * tanhf(c);
          if (a.save) { gp[0] = i; gp[H] = f; gp[2 * H] = g_; gp[3 * H] = o; }
          cell[row * H + unit] = c;  // the cell state is carried through this buffer (needed even for inference)
        } else if (MODE == 3) {
          const float r = sigm(gp[0] + acc[0][j]), z = sigm(gp[H] + acc[1 % G][j]);
          const float q = acc[2 % G][j] + brn;
          const float n = tanhf(gp[2 * H] + r * q);
          const float hp = step > 0 ? a.y[((size_t)tp * B + b0 + j) * HO + dir * H + unit] : 0.f;
          h = (1.f - z) * n + z * hp;
          if (a.save) { gp[0] = r; gp[H] = z; gp[2 * H] = n; cell[row * H + unit] = q; }
        } else {
          const float v = gp[0] + acc[0][j];
          h = MODE == 0 ? fmaxf(v, 0.f) : tanhf(v);
          if (a.save) gp[0] = h;
        }
        a.y[row * HO + dir * H + unit] = h;
      }
    }
  }
}

// backward step, part A: gate gradients of frame t from dh = dy + dh_rec (+ carried terms)
template <int MODE>
__global__ void stream_bwd_gates(RecArgs a, int step, const float *dhrec, float *carry) {
  constexpr int G = MODE == 2 ? 4 : (MODE == 3 ? 3 : 1);
  const int H = a.H, B = a.B, T = a.T, GH = G * H, HO = H * a.dirs;
  const int dir = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, unit = idx % H;
  const int fstep = T - 1 - step;
  const int t = dir ? T - 1 - fstep : fstep;
  const int tp = dir ? t + 1 : t - 1;
  const bool first = fstep == 0;
  float *gates = a.gates[dir];
  float *cell = a.cell[dir];
  const size_t row = (size_t)t * B + b, sidx = ((size_t)dir * B + b) * H + unit;
  float dh = a.dy[row * HO + dir * H + unit] + (step > 0 ? dhrec[sidx] : 0.f);
  float *gp = gates + row * GH + unit;
  const float cr = step > 0 ? carry[sidx] : 0.f;
  if (MODE == 2) {
    const float i = gp[0], f = gp[H], g_ = gp[2 * H], o = gp[3 * H];
    const float c = cell[row * H + unit];
    const float cp = first ? 0.f : cell[((size_t)tp * B + b) * H + unit];
    const float tc = tanhf(c);
    const float dc = dh * o * (1.f - tc * tc) + cr;
    gp[0] = dc * g_ * i * (1.f - i);
    gp[H] = dc * cp * f * (1.f - f);
    gp[2 * H] = dc * i * (1.f - g_ * g_);
    gp[3 * H] = dh * tc * o * (1.f - o);
    carry[sidx] = dc * f;
  } else if (MODE == 3) {
    dh += cr;
    const float r = gp[0], z = gp[H], n = gp[2 * H];
    const float q = cell[row * H + unit];
    const float hp = first ? 0.f : a.y[((size_t)tp * B + b) * HO + dir * H + unit];
    const float dn = dh * (1.f - z) * (1.f - n * n);
    gp[0] = dn * q * r * (1.f - r);
    gp[H] = dh * (hp - n) * z * (1.f - z);
    gp[2 * H] = dn;
    cell[row * H + unit] = dn * r;
    carry[sidx] = dh * z;
  } else {
    const float h = gp[0];
    gp[0] = dh * (MODE == 0 ? (h > 0.f ? 1.f : 0.f) : (1.f - h * h));
  }
}

// backward step, part B: dh_rec[b][k] = sum_rows dgrec[b][row] * R[row][k]
template <int MODE>
__global__ void stream_bwd_dh(RecArgs a, int step, float *dhrec) {
  constexpr int G = MODE == 2 ? 4 : (MODE == 3 ? 3 : 1);
  const int H = a.H, B = a.B, T = a.T, GH = G * H;
  const int dir = blockIdx.y;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * H) return;
  const int b = idx / H, k = idx % H;
  const int fstep = T - 1 - step;
  const int t = dir ? T - 1 - fstep : fstep;
  const float *R = a.w_rec[dir];
  const float *dg = a.gates[dir] + ((size_t)t * B + b) * GH;
  const float *dq = a.cell[dir] + ((size_t)t * B + b) * H;
  float acc = 0.f;
  for (int row = 0; row < GH; row++) {
    const float d = (MODE == 3 && row >= 2 * H) ? dq[row - 2 * H] : dg[row];
    acc = fmaf(d, R[(size_t)row * H + k], acc);
  }
  dhrec[((size_t)dir * B + b) * H + k] = acc;
}

template <int MODE>
cudaError_t fwd(const RecArgs &a, cudaStream_t s) {
  dim3 grid((a.H + kWarps - 1) / kWarps, a.dirs);
  for (int step = 0; step < a.T; step++) stream_fwd_step<MODE><<<grid, kWarps * 32, 0, s>>>(a, step);
  return cudaGetLastError();
}
template <int MODE>
cudaError_t bwd(const RecArgs &a, float *scratch, cudaStream_t s) {
  float *dhrec = scratch, *carry = scratch + (size_t)a.dirs * a.B * a.H;
  dim3 grid((a.B * a.H + 255) / 256, a.dirs);
  for (int step = 0; step < a.T; step++) {
    stream_bwd_gates<MODE><<<grid, 256, 0, s>>>(a, step, dhrec, carry);
    if (step + 1 < a.T) stream_bwd_dh<MODE><<<grid, 256, 0, s>>>(a, step, dhrec);
  }
  return cudaGetLastError();
}

}  // namespace

size_t rec_stream_scratch_floats(int dirs, int B, int H) { return (size_t)2 * dirs * B * H; }

cudaError_t rec_stream_forward(const RecArgs &a, cudaStream_t s) {
  switch (a.mode) {
    case 0: return fwd<0>(a, s);
    case 1: return fwd<1>(a, s);
    case 2: return fwd<2>(a, s);
    default: return fwd<3>(a, s);
  }
}
cudaError_t rec_stream_backward(const RecArgs &a, float *scratch, cudaStream_t s) {
  switch (a.mode) {
    case 0: return bwd<0>(a, scratch, s);
    case 1: return bwd<1>(a, scratch, s);
    case 2: return bwd<2>(a, scratch, s);
    default: return bwd<3>(a, scratch, s);
  }
}

}  // namespace b200
