/*
 * oracle/feat_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the training input path of kaldi-ctc:
 *   kaldi::CompressedMatrix            src/matrix/compressed-matrix.{h,cc}
 *     GlobalHeader / PerColHeader         .h:128-143
 *     CopyFromMat (compress)              .cc:41-121
 *     FloatToUint16 / Uint16ToFloat       .cc:234-251
 *     ComputeColHeader                    .cc:253-331
 *     FloatToChar / CharToFloat           .cc:334-374
 *     CopyToMat (decompress)              .cc:493-529
 *   kaldi::ctc::FormatNnetInput        src/ctc/ctc-nnet-update.cc:351-424
 *   FrameSubsamplingShiftFeatureTimes  src/ctc/ctc-nnet-example.cc:78-93 (row selection on a plain Matrix)
 *
 * Arithmetic follows the C++ expression by expression, including its float/double mixing
 * (e.g. `p0 + (p25 - p0) * value * (1/64.0)` is float*float -> double multiply -> double add ->
 * float).  Built with -ffp-contract=off and without FMA, like Kaldi's own build (-msse -msse2).
 *
 * Pinning: PINNED AGAINST THE REFERENCE'S OWN CODE.  oracle/ref/Makefile compiles the reference's
 * src/matrix/compressed-matrix.cc and src/ctc/ctc-nnet-update.cc where they lie (oracle/_ref/), and
 * tests/test_reference_linked_cpu.py checks this file against them: decompression (CopyToMat) bit for bit,
 * compression byte for byte on random and pathological matrices of both storage formats, FormatNnetInput bit
 * for bit with splicing contexts, left context and speaker vectors.  (The reference holds no golden vectors of
 * its own for this path: matrix-lib-test.cc:4126-4230 only checks properties on random matrices;
 * tests/test_feat_oracle.py keeps those properties and hand-computed known answers as a second pin.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  int32_t format;
  float min_value;
  float range;
  int32_t num_rows;
  int32_t num_cols;
} GlobalHeader; /* 20 bytes (.h:128-134) */

typedef struct {
  uint16_t percentile_0, percentile_25, percentile_75, percentile_100;
} PerColHeader; /* .h:138-143 */

long feat_oracle_data_size(int format, int rows, int cols) { /* .cc:28-38 */
  if (format == 1) return (long)sizeof(GlobalHeader) + (long)cols * ((long)sizeof(PerColHeader) + rows);
  return (long)sizeof(GlobalHeader) + 2L * rows * cols;
}

static uint16_t float_to_uint16(const GlobalHeader *h, float value) { /* .cc:234-243 */
  float f = (value - h->min_value) / h->range;
  if (f > 1.0) f = 1.0;
  if (f < 0.0) f = 0.0;
  return (uint16_t)(int)(f * 65535 + 0.499);
}

static float uint16_to_float(const GlobalHeader *h, uint16_t value) { /* .cc:245-251 */
  return h->min_value + h->range * 1.52590218966964e-05F * value;
}

static unsigned char float_to_char(float p0, float p25, float p75, float p100, float value) { /* .cc:334-361 */
  int ans;
  if (value < p25) {
    float f = (value - p0) / (p25 - p0);
    ans = (int)(f * 64 + 0.5);
    if (ans < 0) ans = 0;
    if (ans > 64) ans = 64;
  } else if (value < p75) {
    float f = (value - p25) / (p75 - p25);
    ans = 64 + (int)(f * 128 + 0.5);
    if (ans < 64) ans = 64;
    if (ans > 192) ans = 192;
  } else {
    float f = (value - p75) / (p100 - p75);
    ans = 192 + (int)(f * 63 + 0.5);
    if (ans < 192) ans = 192;
    if (ans > 255) ans = 255;
  }
  return (unsigned char)ans;
}

static float char_to_float(float p0, float p25, float p75, float p100, unsigned char value) { /* .cc:364-374 */
  if (value <= 64) {
    return p0 + (p25 - p0) * value * (1 / 64.0);
  } else if (value <= 192) {
    return p25 + (p75 - p25) * (value - 64) * (1 / 128.0);
  } else {
    return p75 + (p100 - p75) * (value - 192) * (1 / 63.0);
  }
}

static int cmp_float(const void *a, const void *b) {
  const float x = *(const float *)a, y = *(const float *)b;
  return (x > y) - (x < y);
}
static uint16_t min16(uint16_t a, uint16_t b) { return a < b ? a : b; }
static uint16_t max16(uint16_t a, uint16_t b) { return a > b ? a : b; }

/* .cc:253-331.  nth_element leaves the order statistics 0, n/4, 3(n/4), n-1 in place; a full sort
 * puts the same values there. */
static void compute_col_header(const GlobalHeader *g, const float *data, int stride, int num_rows,
                               PerColHeader *h, float *sdata) {
  for (int i = 0; i < num_rows; i++) sdata[i] = data[(size_t)i * stride];
  qsort(sdata, num_rows, sizeof(float), cmp_float);
  if (num_rows >= 5) {
    const int q = num_rows / 4;
    h->percentile_0 = min16(float_to_uint16(g, sdata[0]), 65532);
    h->percentile_25 = min16(max16(float_to_uint16(g, sdata[q]), (uint16_t)(h->percentile_0 + 1)), 65533);
    h->percentile_75 = min16(max16(float_to_uint16(g, sdata[3 * q]), (uint16_t)(h->percentile_25 + 1)), 65534);
    h->percentile_100 = max16(float_to_uint16(g, sdata[num_rows - 1]), (uint16_t)(h->percentile_75 + 1));
  } else {
    h->percentile_0 = min16(float_to_uint16(g, sdata[0]), 65532);
    if (num_rows > 1)
      h->percentile_25 = min16(max16(float_to_uint16(g, sdata[1]), (uint16_t)(h->percentile_0 + 1)), 65533);
    else
      h->percentile_25 = h->percentile_0 + 1;
    if (num_rows > 2)
      h->percentile_75 = min16(max16(float_to_uint16(g, sdata[2]), (uint16_t)(h->percentile_25 + 1)), 65534);
    else
      h->percentile_75 = h->percentile_25 + 1;
    if (num_rows > 3)
      h->percentile_100 = max16(float_to_uint16(g, sdata[3]), (uint16_t)(h->percentile_75 + 1));
    else
      h->percentile_100 = h->percentile_75 + 1;
  }
}

/* CopyFromMat, .cc:41-121.  mat row-major [rows x cols] with `stride`; out holds
 * feat_oracle_data_size(format, rows, cols) bytes.  force_format 0 = the reference's rule
 * (format 1 iff rows > 8).  Returns the number of bytes written (0 for an empty matrix). */
long feat_oracle_compress(const float *mat, int rows, int cols, int stride, int force_format, void *out) {
  if (rows == 0 || cols == 0) return 0;
  GlobalHeader g;
  float mn = mat[0], mx = mat[0];
  for (int r = 0; r < rows; r++)
    for (int c = 0; c < cols; c++) {
      const float v = mat[(size_t)r * stride + c];
      if (v < mn) mn = v;
      if (v > mx) mx = v;
    }
  if (mx == mn) mx = mn + (1.0 + fabs(mn));
  g.min_value = mn;
  g.range = mx - mn;
  if (g.range <= 0.0) g.range = 1.0e-05;
  g.num_rows = rows;
  g.num_cols = cols;
  g.format = force_format ? force_format : (rows > 8 ? 1 : 2);
  memcpy(out, &g, sizeof(g));
  if (g.format == 1) {
    PerColHeader *hd = (PerColHeader *)((char *)out + sizeof(GlobalHeader));
    unsigned char *bytes = (unsigned char *)(hd + cols);
    float *sdata = (float *)malloc(sizeof(float) * rows);
    for (int c = 0; c < cols; c++) {
      PerColHeader h;
      compute_col_header(&g, mat + c, stride, rows, &h, sdata);
      memcpy(hd + c, &h, sizeof(h));
      const float p0 = uint16_to_float(&g, h.percentile_0), p25 = uint16_to_float(&g, h.percentile_25),
                  p75 = uint16_to_float(&g, h.percentile_75), p100 = uint16_to_float(&g, h.percentile_100);
      for (int r = 0; r < rows; r++)
        bytes[(size_t)c * rows + r] = float_to_char(p0, p25, p75, p100, mat[(size_t)r * stride + c]);
    }
    free(sdata);
  } else {
    uint16_t *d = (uint16_t *)((char *)out + sizeof(GlobalHeader));
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) d[(size_t)r * cols + c] = float_to_uint16(&g, mat[(size_t)r * stride + c]);
  }
  return feat_oracle_data_size(g.format, rows, cols);
}

/* CopyToMat (kNoTrans), .cc:493-529: blob -> row-major [rows x cols]. */
int feat_oracle_decompress(const void *blob, float *out) {
  GlobalHeader g;
  memcpy(&g, blob, sizeof(g));
  const int rows = g.num_rows, cols = g.num_cols;
  if (g.format == 1) {
    const PerColHeader *hd = (const PerColHeader *)((const char *)blob + sizeof(GlobalHeader));
    const unsigned char *bytes = (const unsigned char *)(hd + cols);
    for (int c = 0; c < cols; c++) {
      PerColHeader h;
      memcpy(&h, hd + c, sizeof(h));
      const float p0 = uint16_to_float(&g, h.percentile_0), p25 = uint16_to_float(&g, h.percentile_25),
                  p75 = uint16_to_float(&g, h.percentile_75), p100 = uint16_to_float(&g, h.percentile_100);
      for (int r = 0; r < rows; r++)
        out[(size_t)r * cols + c] = char_to_float(p0, p25, p75, p100, bytes[(size_t)c * rows + r]);
    }
  } else if (g.format == 2) {
    const uint16_t *d = (const uint16_t *)((const char *)blob + sizeof(GlobalHeader));
    for (int r = 0; r < rows; r++)
      for (int c = 0; c < cols; c++) {
        uint16_t v;
        memcpy(&v, d + (size_t)r * cols + c, 2);
        out[(size_t)r * cols + c] = uint16_to_float(&g, v);
      }
  } else {
    return 1;
  }
  return 0;
}

/*
 * FormatNnetInput, ctc-nnet-update.cc:351-424.  blobs[m] = data[m].input_frames (in-memory
 * CompressedMatrix), spk[m] = data[m].spk_info (spk_dim floats, may be NULL when spk_dim == 0),
 * left_context = data[0].left_context, nnet_left/right = nnet.LeftContext()/RightContext().
 * out: [max_num_frames * num_splice * minibatch, feat_dim + spk_dim], zero filled here.
 * The tmp_feat matrix of the reference, [max_num_frames x (minibatch * tot_dim * num_splice)], IS
 * that buffer (the final memcpy just re-labels rows).  Returns max_num_frames, or -1 on bad input.
 */
int feat_oracle_format_nnet_input(const void *const *blobs, const float *const *spk, int spk_dim, int minibatch,
                                  int left_context, int nnet_left, int nnet_right, float *out,
                                  long out_floats) {
  if (minibatch <= 0) return -1;
  const int num_splice = 1 + nnet_right + nnet_left;
  GlobalHeader g0;
  memcpy(&g0, blobs[0], sizeof(g0));
  if (g0.num_rows < num_splice || left_context < nnet_left) return -1;
  const int feat_dim = g0.num_cols, tot_dim = feat_dim + spk_dim;
  const int ignore_frames = left_context - nnet_left;
  int max_num_frames = 0;
  for (int m = 0; m < minibatch; m++) {
    GlobalHeader g;
    memcpy(&g, blobs[m], sizeof(g));
    const int n = g.num_rows - num_splice - ignore_frames + 1;
    if (n > max_num_frames) max_num_frames = n;
  }
  const long need = (long)max_num_frames * num_splice * minibatch * tot_dim;
  if (need > out_floats) return -1;
  memset(out, 0, sizeof(float) * need);
  const long row_floats = (long)minibatch * tot_dim * num_splice; /* one row of tmp_feat */
  long off = 0;                                                   /* feat_dim_offset */
  for (int m = 0; m < minibatch; m++) {
    GlobalHeader g;
    memcpy(&g, blobs[m], sizeof(g));
    float *full = (float *)malloc(sizeof(float) * (size_t)g.num_rows * g.num_cols);
    feat_oracle_decompress(blobs[m], full);
    const int n = g.num_rows - num_splice - ignore_frames + 1;
    for (int s = 0; s < num_splice; s++) {
      for (int t = 0; t < n; t++)
        memcpy(out + t * row_floats + off, full + (size_t)(ignore_frames + s + t) * feat_dim,
               sizeof(float) * feat_dim);
      off += feat_dim;
      if (spk_dim != 0) {
        for (int t = 0; t < n; t++) memcpy(out + t * row_floats + off, spk[m], sizeof(float) * spk_dim);
        off += spk_dim;
      }
    }
    free(full);
  }
  return max_num_frames;
}
