// kaldi_ctc_b200/csrc/rnn_clip.cu -- ClipGradientComponent::Backprop of the reference
// (src/nnet2/nnet-cudnn-component.cc:912-1055) as a stream-ordered device pipeline: norm-based row
// clipping with the component's counters (num_clipped_, count_, num_backpropped_, num_self_repaired_),
// and the stochastic self-repair term (RepairGradients, :970-1055).  The reference runs this as ~25
// CuMatrix/CuVector operations with host round trips for every scalar; here every scalar (the decision
// included) stays on the device, so the step never synchronises:
//   k1 clip rows, count the clipped ones, partial sums of the row norms after clipping
//   f1 (1 CTA) counters, sum of norms, DECISION: repair iff the caller's coin came up (RandUniform() <= 0.5 is
//      drawn on the host, :981), self_repair_scale != 0, threshold < 1, count > 0 and
//      num_clipped / count > threshold (:983-995)
//   k2 row norms of repair_mat = max(|in_value| - target, 0) .* sign(in_value)              (:1007-1017)
//   f2 coefficient  -scale * clipped_proportion * mean(|deriv rows|) / mean(|repair rows|) / repair_probability
//   k3 deriv += coefficient * repair_mat, partial sums of the new row norms                  (:1036-1040)
//   f3 + k4 rescale so that the summed row norm is what it was before the repair term         (:1041-1046)
// k2..k4 return at once when the decision is negative (and are not launched when the coin says no).
#include <algorithm>

#include "rnn_common.cuh"

namespace b200 {
namespace {

constexpr int kWarps = 8;

// scal layout (floats): [0] sum of row norms after clipping, [1] decision (0/1), [2] coefficient,
//                       [3] rescale factor, [4] clipped proportion
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void block_partial(float v, int lane, int wi, float *partial) {
  __shared__ float sh[kWarps];
  if (lane == 0) sh[wi] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; w++) s += sh[w];   // fixed order: deterministic
    partial[blockIdx.x] = s;
  }
}

__global__ void __launch_bounds__(kWarps * 32) clip_rows_kernel(float *d, int rows, int cols, float thr, float *partial,
                                                              int *clipped_now) {
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int row = blockIdx.x * kWarps + wi;
  float norm_after = 0.f;
  if (row < rows) {
    float *p = d + (size_t)row * cols;
    float ss = 0.f;
    for (int c = lane; c < cols; c += 32) ss = fmaf(p[c], p[c], ss);
    ss = warp_sum(ss);
    norm_after = sqrtf(ss);
    if (thr > 0.f) {
      const float r = ss / (thr * thr);
      if (r > 1.0f) {
        const float sc = rsqrtf(r);
        for (int c = lane; c < cols; c += 32) p[c] *= sc;
        norm_after *= sc;
        if (lane == 0) atomicAdd(clipped_now, 1);   // integer atomics: order-independent
      }
    }
  }
  block_partial(norm_after, lane, wi, partial);
}

// counters: [0] num_clipped, [1] count, [2] num_self_repaired, [3] num_backpropped (to_update's)
// decide:   the counters the decision reads (`this` of the reference: the same object when the net updates itself)
__global__ void clip_finalize_kernel(const float *partial, int nblocks, int rows, int *clipped_now, int *counters,
                                     const int *decide, int attempt, float prop_threshold, float repair_scale,
                                     float thr, float *scal) {
  __shared__ float sh[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += 256) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    scal[0] = sh[0];
    const int nc = *clipped_now;
    *clipped_now = 0;
    if (counters && thr > 0.f) {
      counters[0] += nc;
      counters[1] += rows;
      counters[3] += 1;
    }
    int go = 0;
    float prop = 0.f;
    if (attempt && decide && thr > 0.f && prop_threshold < 1.0f && repair_scale != 0.0f && decide[1] > 0) {
      prop = (float)decide[0] / (float)decide[1];
      go = prop > prop_threshold;
    }
    if (go && counters) counters[2] += 1;
    scal[1] = go ? 1.f : 0.f;
    scal[4] = prop;
  }
}

__device__ __forceinline__ float repair_elem(float v, float target) {
  const float m = fmaxf(fabsf(v) - target, 0.f);
  return v > 0.f ? m : -m;   // ApplyHeaviside: sign is +1 only for v > 0 (:1001-1005)
}

__global__ void __launch_bounds__(kWarps * 32) repair_norm_kernel(const float *v, int rows, int cols, float target,
                                                                const float *scal, float *partial) {
  if (scal[1] == 0.f) return;
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int row = blockIdx.x * kWarps + wi;
  float n = 0.f;
  if (row < rows) {
    const float *p = v + (size_t)row * cols;
    float ss = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float r = repair_elem(p[c], target);
      ss = fmaf(r, r, ss);
    }
    n = sqrtf(warp_sum(ss));
  }
  block_partial(n, lane, wi, partial);
}

// which: 0 -> coefficient from the repair norms, 1 -> rescale factor from the new derivative norms
__global__ void repair_finalize_kernel(const float *partial, int nblocks, int rows, float repair_scale, int which,
                                       float *scal) {
  if (scal[1] == 0.f) return;
  __shared__ float sh[256];
  float s = 0.f;
  for (int i = threadIdx.x; i < nblocks; i += 256) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    if (which == 0) {
      const float magnitude = repair_scale * scal[4] * (scal[0] / rows);
      const float mean_repair = sh[0] / rows;
      scal[2] = sh[0] != 0.f ? -(magnitude / mean_repair) / 0.5f : 0.f;   // repair_probability = 0.5 (:979)
    } else {
      scal[3] = sh[0] != 0.f ? scal[0] / sh[0] : 1.f;
    }
  }
}

__global__ void __launch_bounds__(kWarps * 32) repair_apply_kernel(float *d, const float *v, int rows, int cols,
                                                                 float target, const float *scal, float *partial) {
  if (scal[1] == 0.f) return;
  const float coef = scal[2];
  const int lane = threadIdx.x & 31, wi = threadIdx.x >> 5;
  const int row = blockIdx.x * kWarps + wi;
  float n = 0.f;
  if (row < rows) {
    float *p = d + (size_t)row * cols;
    const float *q = v + (size_t)row * cols;
    float ss = 0.f;
    for (int c = lane; c < cols; c += 32) {
      const float x = fmaf(coef, repair_elem(q[c], target), p[c]);
      p[c] = x;
      ss = fmaf(x, x, ss);
    }
    n = sqrtf(warp_sum(ss));
  }
  block_partial(n, lane, wi, partial);
}

__global__ void repair_rescale_kernel(float *d, size_t n, const float *scal) {
  if (scal[1] == 0.f) return;
  const float f = scal[3];
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) d[i] *= f;
}

}  // namespace

size_t clip_gradient_workspace_floats(int rows) { return (size_t)(rows + kWarps - 1) / kWarps + 16; }

cudaError_t clip_gradient_backprop(float *deriv, const float *in_value, int rows, int cols, float thr,
                                   float prop_threshold, float target, float repair_scale, int attempt_repair,
                                   int *counters, const int *decide, float *ws, cudaStream_t stream) {
  const int nblocks = (rows + kWarps - 1) / kWarps;
  float *scal = ws, *partial = ws + 16;
  int *clipped_now = reinterpret_cast<int *>(ws + 8);
  cudaError_t e = cudaMemsetAsync(clipped_now, 0, sizeof(int), stream);   // (the workspace arrives uninitialised)
  if (e != cudaSuccess) return e;
  clip_rows_kernel<<<nblocks, kWarps * 32, 0, stream>>>(deriv, rows, cols, thr, partial, clipped_now);
  const int attempt = attempt_repair && in_value != nullptr;
  clip_finalize_kernel<<<1, 256, 0, stream>>>(partial, nblocks, rows, clipped_now, counters, decide, attempt,
                                             prop_threshold, repair_scale, thr, scal);
  if (attempt && thr > 0.f && prop_threshold < 1.0f && repair_scale != 0.0f) {
    repair_norm_kernel<<<nblocks, kWarps * 32, 0, stream>>>(in_value, rows, cols, target, scal, partial);
    repair_finalize_kernel<<<1, 256, 0, stream>>>(partial, nblocks, rows, repair_scale, 0, scal);
    repair_apply_kernel<<<nblocks, kWarps * 32, 0, stream>>>(deriv, in_value, rows, cols, target, scal, partial);
    repair_finalize_kernel<<<1, 256, 0, stream>>>(partial, nblocks, rows, repair_scale, 1, scal);
    const size_t n = (size_t)rows * cols;
    repair_rescale_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, stream>>>(deriv, n, scal);
  }
  return cudaGetLastError();
}

}  // namespace b200
