// ctc_format.cu -- kaldi::ctc::FormatNnetInput (src/ctc/ctc-nnet-update.cc:351-424) on the GPU.
//
// The reference decompresses every example's CompressedMatrix on the CPU
// (`Matrix<BaseFloat> full_src(data[chunk].input_frames)`, :393), copies the frames into the
// time-major slab on the CPU (:396-418, + one more full memcpy :421) and only then uploads 4 bytes per
// (padded) element.  Here the COMPRESSED bytes are uploaded (1 byte per element + 8 bytes per column,
// src/matrix/compressed-matrix.h:128-143) and one kernel decompresses, splices, appends the speaker
// vector, interleaves the utterances (row = (t*B + b)*num_splice + s) and zero-fills the padding.
//
// Decompression is BIT-EXACT with CompressedMatrix::CopyToMat (compressed-matrix.cc:493-529): the
// reference's expressions mix float and double (`p0 + (p25 - p0) * value * (1/64.0)`, :364-374), so the
// kernel does the same conversions with explicitly rounded, non-contracted operations.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>

#include "../../include/b200ctc.h"

namespace {

struct GlobalHeader {  // compressed-matrix.h:128-134
  int32_t format;
  float min_value;
  float range;
  int32_t num_rows;
  int32_t num_cols;
};
static_assert(sizeof(GlobalHeader) == 20, "CompressedMatrix::GlobalHeader is 20 bytes");

struct UttEntry {
  long long blob_off;   // byte offset of the blob inside the staging buffer (16-byte aligned)
  int num_rows;         // rows of the stored matrix
  int n_frames;         // output frames of this utterance (num_rows - num_splice - ignore_frames + 1)
  int format;           // 1: per-column headers + bytes (column-major); 2: uint16 row-major
  float min_value, range;
  long long spk_off;    // byte offset of the speaker vector (or -1)
};

constexpr int kTT = 64;       // output time steps per CTA
constexpr int kThreads = 256;

__device__ __forceinline__ float u16_to_float(float min_value, float range, unsigned v) {
  // min_value + range * 1.52590218966964e-05F * value      (compressed-matrix.cc:245-251), all float
  return __fadd_rn(min_value, __fmul_rn(__fmul_rn(range, 1.52590218966964e-05F), (float)v));
}

__device__ __forceinline__ float char_to_float(float p0, float p25, float p75, float p100, unsigned v) {
  // compressed-matrix.cc:364-374: (float difference * float(int)) -> double * (1/64.0) -> + double(p) -> float
  float base, diff;
  int k;
  double scale;
  if (v <= 64) {
    base = p0; diff = __fsub_rn(p25, p0); k = (int)v; scale = 1 / 64.0;
  } else if (v <= 192) {
    base = p25; diff = __fsub_rn(p75, p25); k = (int)v - 64; scale = 1 / 128.0;
  } else {
    base = p75; diff = __fsub_rn(p100, p75); k = (int)v - 192; scale = 1 / 63.0;
  }
  const float prod = __fmul_rn(diff, (float)k);
  return __double2float_rn(__dadd_rn((double)base, __dmul_rn((double)prod, scale)));
}

struct FmtArgs {
  const uint8_t *staging;   // blobs + speaker vectors
  const UttEntry *utts;     // [B]
  float *out;               // [max_frames * B * S, tot_dim]
  int B, S, feat_dim, spk_dim, ignore_frames, max_frames;
};

// grid (ceil(max_frames / kTT), B).  smem: pcol[feat_dim][4] floats, tile[(kTT+S-1)][feat_dim+1] floats.
__global__ void __launch_bounds__(kThreads) format_input_kernel(FmtArgs a) {
  extern __shared__ float sm[];
  const int b = blockIdx.y, t0 = blockIdx.x * kTT;
  const UttEntry u = a.utts[b];
  const int F = a.feat_dim, S = a.S, tot = F + a.spk_dim;
  const int R = kTT + S - 1;                       // source rows this CTA may need
  float *pcol = sm;                                // [F][4]
  float *tile = sm + 4 * F;                        // [R][F + 1]
  const int pitch = F + 1;
  const int nt = max(0, min(kTT, u.n_frames - t0));  // valid output steps in this tile
  const uint8_t *blob = a.staging + u.blob_off;
  if (nt > 0) {
    const int r0 = a.ignore_frames + t0;           // first source row
    const int nr = nt + S - 1;                     // source rows needed
    if (u.format == 1) {
      const uint16_t *hd = reinterpret_cast<const uint16_t *>(blob + sizeof(GlobalHeader));
      for (int i = threadIdx.x; i < 4 * F; i += kThreads) pcol[i] = u16_to_float(u.min_value, u.range, hd[i]);
      __syncthreads();
      const uint8_t *bytes = blob + sizeof(GlobalHeader) + 8 * (size_t)F;
      for (int i = threadIdx.x; i < F * nr; i += kThreads) {
        const int c = i / nr, r = i - c * nr;
        const unsigned v = bytes[(size_t)c * u.num_rows + r0 + r];
        tile[r * pitch + c] = char_to_float(pcol[4 * c], pcol[4 * c + 1], pcol[4 * c + 2], pcol[4 * c + 3], v);
      }
    } else {
      const uint16_t *d = reinterpret_cast<const uint16_t *>(blob + sizeof(GlobalHeader));
      for (int i = threadIdx.x; i < F * nr; i += kThreads) {
        const int r = i / F, c = i - r * F;
        tile[r * pitch + c] = u16_to_float(u.min_value, u.range, d[(size_t)(r0 + r) * F + c]);
      }
    }
  }
  __syncthreads();
  const float *spk = u.spk_off >= 0 ? reinterpret_cast<const float *>(a.staging + u.spk_off) : nullptr;
  const int tmax = min(kTT, a.max_frames - t0);
  const int per_t = S * tot;
  for (int i = threadIdx.x; i < tmax * per_t; i += kThreads) {
    const int tt = i / per_t, rem = i - tt * per_t;
    const int s = rem / tot, k = rem - s * tot;
    float v = 0.f;                                  // kSetZero padding (:385-387)
    if (tt < nt) v = k < F ? tile[(tt + s) * pitch + k] : spk[k - F];
    a.out[((size_t)(t0 + tt) * a.B + b) * per_t + rem] = v;
  }
}

inline size_t align16(size_t v) { return (v + 15) & ~size_t(15); }

struct Parsed {
  int feat_dim, num_splice, ignore_frames, max_frames;
  size_t blob_bytes_total, table_off, spk_off, total;
};

size_t blob_size(const GlobalHeader &g) {  // CompressedMatrix::DataSize, compressed-matrix.cc:28-38
  return g.format == 1 ? sizeof(GlobalHeader) + (size_t)g.num_cols * (8 + (size_t)g.num_rows)
                       : sizeof(GlobalHeader) + 2 * (size_t)g.num_rows * g.num_cols;
}

ctcStatus_t parse(const void *const *blobs, int minibatch, int spk_dim, int left_context, int nnet_left,
                  int nnet_right, Parsed *p) {
  if (!blobs || minibatch <= 0 || spk_dim < 0 || nnet_left < 0 || nnet_right < 0) return CTC_STATUS_INVALID_VALUE;
  if (left_context < nnet_left) return CTC_STATUS_INVALID_VALUE;  // KALDI_ASSERT(left_context >= nnet.LeftContext())
  p->num_splice = 1 + nnet_right + nnet_left;
  p->ignore_frames = left_context - nnet_left;
  p->max_frames = 0;
  size_t off = 0;
  for (int m = 0; m < minibatch; m++) {
    if (!blobs[m]) return CTC_STATUS_INVALID_VALUE;
    GlobalHeader g;
    memcpy(&g, blobs[m], sizeof(g));
    if ((g.format != 1 && g.format != 2) || g.num_rows <= 0 || g.num_cols <= 0) return CTC_STATUS_INVALID_VALUE;
    if (m == 0) {
      p->feat_dim = g.num_cols;
      if (g.num_rows < p->num_splice) return CTC_STATUS_INVALID_VALUE;  // KALDI_ASSERT, :357
    } else if (g.num_cols != p->feat_dim) {
      return CTC_STATUS_INVALID_VALUE;
    }
    const int n = g.num_rows - p->num_splice - p->ignore_frames + 1;
    if (n > p->max_frames) p->max_frames = n;
    off += align16(blob_size(g));
  }
  p->blob_bytes_total = off;
  p->spk_off = off;
  off += align16((size_t)minibatch * spk_dim * sizeof(float));
  p->table_off = off;
  off += align16((size_t)minibatch * sizeof(UttEntry));
  p->total = off;
  return p->max_frames > 0 ? CTC_STATUS_SUCCESS : CTC_STATUS_INVALID_VALUE;
}

}  // namespace

extern "C" {

ctcStatus_t b200ctc_format_input_size(const void *const *examples_host, int minibatch, int spk_dim,
                                      int left_context, int nnet_left_context, int nnet_right_context,
                                      int *max_num_frames, int *feat_dim, size_t *staging_bytes) {
  Parsed p;
  const ctcStatus_t st = parse(examples_host, minibatch, spk_dim, left_context, nnet_left_context,
                               nnet_right_context, &p);
  if (st != CTC_STATUS_SUCCESS) return st;
  if (max_num_frames) *max_num_frames = p.max_frames;
  if (feat_dim) *feat_dim = p.feat_dim;
  if (staging_bytes) *staging_bytes = p.total;
  return CTC_STATUS_SUCCESS;
}

ctcStatus_t b200ctc_format_input(const void *const *examples_host, const float *const *spk_info_host,
                                 int spk_dim, int minibatch, int left_context, int nnet_left_context,
                                 int nnet_right_context, float *input_mat, size_t input_mat_floats,
                                 void *staging_host, void *staging_dev, size_t staging_bytes,
                                 CUstream stream) {
  Parsed p;
  const ctcStatus_t st = parse(examples_host, minibatch, spk_dim, left_context, nnet_left_context,
                               nnet_right_context, &p);
  if (st != CTC_STATUS_SUCCESS) return st;
  if (!input_mat || !staging_host || !staging_dev || staging_bytes < p.total) return CTC_STATUS_INVALID_VALUE;
  if (spk_dim > 0 && !spk_info_host) return CTC_STATUS_INVALID_VALUE;
  const int tot = p.feat_dim + spk_dim;
  const size_t need = (size_t)p.max_frames * p.num_splice * minibatch * tot;
  if (input_mat_floats < need) return CTC_STATUS_INVALID_VALUE;
  const size_t smem = sizeof(float) * (4 * (size_t)p.feat_dim + (size_t)(kTT + p.num_splice - 1) * (p.feat_dim + 1));
  if (smem > 200 * 1024) return CTC_STATUS_INVALID_VALUE;

  // pack: blobs | speaker vectors | table  -> one upload
  uint8_t *h = static_cast<uint8_t *>(staging_host);
  UttEntry *tab = reinterpret_cast<UttEntry *>(h + p.table_off);
  size_t off = 0;
  for (int m = 0; m < minibatch; m++) {
    GlobalHeader g;
    memcpy(&g, examples_host[m], sizeof(g));
    const size_t n = blob_size(g);
    memcpy(h + off, examples_host[m], n);
    tab[m].blob_off = (long long)off;
    tab[m].num_rows = g.num_rows;
    tab[m].n_frames = g.num_rows - p.num_splice - p.ignore_frames + 1;
    if (tab[m].n_frames < 0) tab[m].n_frames = 0;
    tab[m].format = g.format;
    tab[m].min_value = g.min_value;
    tab[m].range = g.range;
    tab[m].spk_off = -1;
    if (spk_dim > 0) {
      if (!spk_info_host[m]) return CTC_STATUS_INVALID_VALUE;
      tab[m].spk_off = (long long)(p.spk_off + (size_t)m * spk_dim * sizeof(float));
      memcpy(h + tab[m].spk_off, spk_info_host[m], sizeof(float) * spk_dim);
    }
    off += align16(n);
  }
  if (cudaMemcpyAsync(staging_dev, h, p.total, cudaMemcpyHostToDevice, stream) != cudaSuccess)
    return CTC_STATUS_EXECUTION_FAILED;

  FmtArgs a;
  a.staging = static_cast<const uint8_t *>(staging_dev);
  a.utts = reinterpret_cast<const UttEntry *>(a.staging + p.table_off);
  a.out = input_mat;
  a.B = minibatch;
  a.S = p.num_splice;
  a.feat_dim = p.feat_dim;
  a.spk_dim = spk_dim;
  a.ignore_frames = p.ignore_frames;
  a.max_frames = p.max_frames;
  // (the attribute is per device: set it on every call that needs it -- a host-side table update, no sync)
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(format_input_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return CTC_STATUS_EXECUTION_FAILED;
  dim3 grid((p.max_frames + kTT - 1) / kTT, minibatch);
  format_input_kernel<<<grid, kThreads, smem, stream>>>(a);
  return cudaGetLastError() == cudaSuccess ? CTC_STATUS_SUCCESS : CTC_STATUS_EXECUTION_FAILED;
}

}  // extern "C"
