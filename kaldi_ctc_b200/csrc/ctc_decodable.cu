// ctc_decodable.cu -- the post-network half of CtcDecodableAmNnet's constructor
// (src/ctc/ctc-decodable-am-nnet.cc:54-86) and of CtcDecodableAmNnetParallel::Compute
// (:89-108) as three small kernels behind b200ctc_decodable (include/b200ctc.h):
//
//   reference (one full-matrix pass each)            here
//   SoftmaxComponent::Propagate (appended for        rowstats: one read of the row -> log-sum-exp
//     decoding, steps/ctc/train.sh:471-476)            and the blank posterior
//   host loop over log_probs(i,0) (a D2H element     rowstats writes the keep flag; scan: one CTA per
//     read per frame!) + CopyRows        (:55-69)      utterance turns flags into destination rows
//   ApplyFloor, ApplyLog, AddVecToRows(-log prior),  write: one read of the kept rows, one write of the
//     Scale                               (:72-83)     compacted [kept, A] matrix, everything in registers
//
// HBM traffic: 2 reads of the input + 1 write of the kept rows (the reference: softmax r+w, CopyRows r+w,
// floor r+w, log r+w, prior r+w, scale r+w = 12 passes).  Rows are time-major t*B+u like every other matrix
// on the path; B=1 is the reference's per-utterance call.
#include <cuda_runtime.h>
#include <math_constants.h>

#include <cstdint>

#include "../../include/b200ctc.h"

namespace {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr int kWarpsPerCta = 8;

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

struct DecArgs {
  const float *in;      // [Tmax*B, A]
  float *out;           // [sum_u T_u, A], utterance u at row base[u]
  const int *len;       // [B] device
  const int *base;      // [B] device
  const float *priors;  // [A] or null
  float *log_priors;    // [A] workspace
  float *lse;           // [Tmax*B] natural-log sum-exp of the row (logits input)
  int *dest;            // [B*Tmax] utterance-major: flag, then destination row (or -1)
  int *kept;            // [B]
  int B, A, Tmax;
  int is_logits;
  float threshold, log_floor, floor_, scale;
};

// One warp per row: log-sum-exp (logits) and the keep decision  p(blank) < threshold  (:57).
__global__ void __launch_bounds__(kWarpsPerCta * 32) decodable_rowstats_kernel(DecArgs d) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= (long)d.Tmax * d.B) return;
  const int t = (int)(row / d.B), u = (int)(row % d.B);
  if (t >= d.len[u]) return;
  const float *a = d.in + row * d.A;
  float p0;
  if (d.is_logits) {
    float m = -CUDART_INF_F, s = 0.f;
    const int A = d.A;
    if ((A & 3) == 0 && (reinterpret_cast<uintptr_t>(a) & 15) == 0) {
      const float4 *a4 = reinterpret_cast<const float4 *>(a);
      for (int k = lane; k < A / 4; k += 32) {
        const float4 v = __ldg(a4 + k);
        const float mx = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
        if (mx > m) {
          s *= exp2f((m - mx) * kLog2e);
          m = mx;
        }
        s += exp2f((v.x - m) * kLog2e) + exp2f((v.y - m) * kLog2e) + exp2f((v.z - m) * kLog2e) +
             exp2f((v.w - m) * kLog2e);
      }
    } else {
      for (int k = lane; k < A; k += 32) {
        const float v = __ldg(a + k);
        if (v > m) {
          s *= exp2f((m - v) * kLog2e);
          m = v;
        }
        s += exp2f((v - m) * kLog2e);
      }
    }
    const float M = warp_max(m);
    s = warp_sum(m == -CUDART_INF_F ? 0.f : s * exp2f((m - M) * kLog2e));
    const float lse = M + log2f(s) * kLn2;
    if (lane == 0) d.lse[row] = lse;
    p0 = expf(__ldg(a) - lse);
  } else {
    p0 = __ldg(a);
  }
  if (lane == 0) d.dest[(long)u * d.Tmax + t] = (d.threshold >= 1.0f || p0 < d.threshold) ? 1 : 0;
}

// One CTA per utterance: flags -> destination rows (stable), the two corner cases of :61-68
// (nothing kept -> keep everything; everything kept -> identity), and log(prior) once.
__global__ void __launch_bounds__(1024) decodable_scan_kernel(DecArgs d) {
  __shared__ int warp_tot[32];
  __shared__ int carry_s;
  const int u = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (u == 0 && d.priors)
    for (int k = tid; k < d.A; k += blockDim.x) d.log_priors[k] = logf(d.priors[k]);
  const int T = d.len[u];
  int *f = d.dest + (long)u * d.Tmax;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int t0 = 0; t0 < T; t0 += blockDim.x) {
    const int t = t0 + tid;
    const int keep = t < T ? f[t] : 0;
    int inc = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += n;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
      int v = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
      }
      warp_tot[lane] = v;  // inclusive over warps
    }
    __syncthreads();
    const int before = carry_s + (w ? warp_tot[w - 1] : 0) + inc - keep;
    if (t < T) f[t] = keep ? before : -1;
    __syncthreads();
    if (tid == 0) carry_s += warp_tot[(blockDim.x >> 5) - 1];
    __syncthreads();
  }
  const int kept = carry_s;
  if (kept == 0) {  // "No Frame will be keeped ... don't skip blank" (:62-63)
    for (int t = tid; t < T; t += blockDim.x) f[t] = t;
  }
  if (tid == 0) d.kept[u] = kept == 0 ? T : kept;
}

// One warp per input row: floor, log, - log prior, * prob_scale, written to its compacted place.
__global__ void __launch_bounds__(kWarpsPerCta * 32) decodable_write_kernel(DecArgs d) {
  const int lane = threadIdx.x & 31;
  const long row = (long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= (long)d.Tmax * d.B) return;
  const int t = (int)(row / d.B), u = (int)(row % d.B);
  if (t >= d.len[u]) return;
  const int dst = d.dest[(long)u * d.Tmax + t];
  if (dst < 0) return;
  const int A = d.A;
  const float *a = d.in + row * A;
  float *o = d.out + ((long)d.base[u] + dst) * A;
  const float lse = d.is_logits ? d.lse[row] : 0.f;
  const float *lp = d.priors ? d.log_priors : nullptr;
  auto f = [&](float x, int k) {
    // log(max(softmax, floor)) == max(x - lse, log floor); probabilities: log(max(p, floor))
    const float l = d.is_logits ? fmaxf(x - lse, d.log_floor) : logf(fmaxf(x, d.floor_));
    return d.scale * (lp ? l - lp[k] : l);
  };
  if ((A & 3) == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(o)) & 15) == 0) {
    const float4 *a4 = reinterpret_cast<const float4 *>(a);
    float4 *o4 = reinterpret_cast<float4 *>(o);
    for (int k = lane; k < A / 4; k += 32) {
      const float4 v = __ldg(a4 + k);
      float4 r;
      r.x = f(v.x, 4 * k);
      r.y = f(v.y, 4 * k + 1);
      r.z = f(v.z, 4 * k + 2);
      r.w = f(v.w, 4 * k + 3);
      __stcs(o4 + k, r);
    }
  } else {
    for (int k = lane; k < A; k += 32) o[k] = f(__ldg(a + k), k);
  }
}

inline size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

struct DecLayout {
  size_t off_len, off_base, off_kept, off_lp, off_lse, off_dest, total;
};

DecLayout dec_layout(int B, int A, int Tmax) {
  DecLayout l;
  size_t o = 0;
  l.off_len = o;  o += align256(sizeof(int) * B);
  l.off_base = o; o += align256(sizeof(int) * B);
  l.off_kept = o; o += align256(sizeof(int) * B);
  l.off_lp = o;   o += align256(sizeof(float) * A);
  l.off_lse = o;  o += align256(sizeof(float) * (size_t)Tmax * B);
  l.off_dest = o; o += align256(sizeof(int) * (size_t)Tmax * B);
  l.total = o;
  return l;
}

}  // namespace

extern "C" {

ctcStatus_t b200ctc_decodable_workspace_size(const int *input_lengths, int alphabet_size, int minibatch,
                                             size_t *size_bytes) {
  if (!input_lengths || !size_bytes || alphabet_size <= 0 || minibatch <= 0) return CTC_STATUS_INVALID_VALUE;
  int Tmax = 0;
  for (int u = 0; u < minibatch; ++u) {
    if (input_lengths[u] < 0) return CTC_STATUS_INVALID_VALUE;
    Tmax = input_lengths[u] > Tmax ? input_lengths[u] : Tmax;
  }
  *size_bytes = dec_layout(minibatch, alphabet_size, Tmax).total;
  return CTC_STATUS_SUCCESS;
}

ctcStatus_t b200ctc_decodable(const float *nnet_output, int input_is_logits, const int *input_lengths,
                              int alphabet_size, int minibatch, const float *priors, float prob_scale,
                              float blank_threshold, float floor_value, float *log_probs, int *kept_dev,
                              int *kept_host, void *workspace, size_t workspace_bytes, CUstream stream) {
  if (!input_lengths || !workspace || alphabet_size <= 0 || minibatch <= 0 || !(floor_value > 0.f))
    return CTC_STATUS_INVALID_VALUE;
  int Tmax = 0;
  long total = 0;
  for (int u = 0; u < minibatch; ++u) {
    if (input_lengths[u] < 0) return CTC_STATUS_INVALID_VALUE;
    Tmax = input_lengths[u] > Tmax ? input_lengths[u] : Tmax;
    total += input_lengths[u];
  }
  if (total > 0x7fffffffL) return CTC_STATUS_INVALID_VALUE;
  const DecLayout l = dec_layout(minibatch, alphabet_size, Tmax);
  if (workspace_bytes < l.total || (reinterpret_cast<uintptr_t>(workspace) & 255)) return CTC_STATUS_INVALID_VALUE;
  if (Tmax == 0) {  // "Input with 0 rows will produce empty output" (:42-47)
    if (kept_host) for (int u = 0; u < minibatch; ++u) kept_host[u] = 0;
    if (kept_dev && cudaMemsetAsync(kept_dev, 0, sizeof(int) * minibatch, stream) != cudaSuccess)
      return CTC_STATUS_EXECUTION_FAILED;
    return CTC_STATUS_SUCCESS;
  }
  if (!nnet_output || !log_probs) return CTC_STATUS_INVALID_VALUE;
  char *w = static_cast<char *>(workspace);
  // lengths and output bases: pageable-source async copies are staged before the call returns
  int *base_h = new int[minibatch];
  int acc = 0;
  for (int u = 0; u < minibatch; ++u) {
    base_h[u] = acc;
    acc += input_lengths[u];
  }
  cudaError_t e = cudaMemcpyAsync(w + l.off_len, input_lengths, sizeof(int) * minibatch, cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(w + l.off_base, base_h, sizeof(int) * minibatch, cudaMemcpyHostToDevice, stream);
  delete[] base_h;
  if (e != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;

  DecArgs d;
  d.in = nnet_output;
  d.out = log_probs;
  d.len = reinterpret_cast<int *>(w + l.off_len);
  d.base = reinterpret_cast<int *>(w + l.off_base);
  d.priors = priors;
  d.log_priors = reinterpret_cast<float *>(w + l.off_lp);
  d.lse = reinterpret_cast<float *>(w + l.off_lse);
  d.dest = reinterpret_cast<int *>(w + l.off_dest);
  d.kept = kept_dev ? kept_dev : reinterpret_cast<int *>(w + l.off_kept);
  d.B = minibatch;
  d.A = alphabet_size;
  d.Tmax = Tmax;
  d.is_logits = input_is_logits ? 1 : 0;
  d.threshold = blank_threshold;
  d.floor_ = floor_value;
  d.log_floor = logf(floor_value);
  d.scale = prob_scale;

  const long rows = (long)Tmax * minibatch;
  const unsigned grid = (unsigned)((rows + kWarpsPerCta - 1) / kWarpsPerCta);
  decodable_rowstats_kernel<<<grid, kWarpsPerCta * 32, 0, stream>>>(d);
  decodable_scan_kernel<<<minibatch, 1024, 0, stream>>>(d);
  decodable_write_kernel<<<grid, kWarpsPerCta * 32, 0, stream>>>(d);
  if (cudaGetLastError() != cudaSuccess) return CTC_STATUS_EXECUTION_FAILED;
  if (kept_host) {
    if (cudaMemcpyAsync(kept_host, d.kept, sizeof(int) * minibatch, cudaMemcpyDeviceToHost, stream) != cudaSuccess ||
        cudaStreamSynchronize(stream) != cudaSuccess)
      return CTC_STATUS_EXECUTION_FAILED;
  }
  return CTC_STATUS_SUCCESS;
}

}  // extern "C"
