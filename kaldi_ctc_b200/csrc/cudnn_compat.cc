// kaldi_ctc_b200/csrc/cudnn_compat.cc -- the cuDNN-5 entry points kaldi-ctc binds to
// (include/cudnn_v5_compat/cudnn.h), implemented on the B200-native recurrent
// kernels through the C ABI of include/b200rnn.h.  Built into libb200cudnn.so, which is
// installed as <root>/lib64/libcudnn.so for `configure --cudnn-root=<root>`.
// Host code only: descriptors are plain structs, the RNN descriptor owns the
// b200rnn plans (one per (minibatch, input width) it has been queried for).
#include <stdlib.h>
#include <string.h>

#include <map>
#include <new>
#include <utility>

#include "../../include/b200rnn.h"
#include "../../include/cudnn_v5_compat/cudnn.h"

struct cudnnContext {
  int unused;
};
struct cudnnTensorStruct {
  cudnnDataType_t dtype;
  int nb;
  int dim[CUDNN_DIM_MAX], stride[CUDNN_DIM_MAX];
};
struct cudnnFilterStruct {
  cudnnDataType_t dtype;
  cudnnTensorFormat_t format;
  int nb;
  int dim[CUDNN_DIM_MAX];
};
struct cudnnConvolutionStruct {
  int n;
  int pad[CUDNN_DIM_MAX], stride[CUDNN_DIM_MAX], upscale[CUDNN_DIM_MAX];
  cudnnConvolutionMode_t mode;
  cudnnDataType_t dtype;
};
struct cudnnDropoutStruct {
  float p;
};
struct cudnnRNNStruct {
  int hidden, layers, dirs, mode, math;
  bool set;
  // (minibatch, input width) -> plan sized for the longest sequence asked about so far
  std::map<std::pair<int, int>, std::pair<b200rnnPlan_t, int> > plans;
};

namespace {

int g_math = -1;

// cudnnSetRNNDescriptor only accepts CUDNN_DATA_FLOAT and the reference is fp32 throughout
// (nnet-cudnn-component.cc:153), so an unmodified binary gets the EXACT fp32 kernels (oracle parity 1e-5).
// The tensor-core mode (BF16 recurrent operands, TF32/BF16 projections; tolerance in DESIGN.md section 5)
// is an explicit opt-in: B200_CUDNN_MATH=tensor in the environment (read once) or b200cudnnSetMath(1).
int default_math() {
  if (g_math >= 0) return g_math;
  static const int env_math = [] {
    const char *e = getenv("B200_CUDNN_MATH");
    return (e && !strcmp(e, "tensor")) ? 1 : 0;
  }();
  return env_math;
}

cudnnStatus_t to_cudnn(b200rnnStatus_t s) {
  switch (s) {
    case B200RNN_STATUS_SUCCESS: return CUDNN_STATUS_SUCCESS;
    case B200RNN_STATUS_INVALID_VALUE: return CUDNN_STATUS_BAD_PARAM;
    case B200RNN_STATUS_ALLOC_FAILED: return CUDNN_STATUS_ALLOC_FAILED;
    case B200RNN_STATUS_NOT_SUPPORTED: return CUDNN_STATUS_NOT_SUPPORTED;
    default: return CUDNN_STATUS_EXECUTION_FAILED;
  }
}

// plan for this (B, D) able to run `T` steps; grow = size queries may enlarge it
b200rnnPlan_t plan_for(cudnnRNNStruct *r, const cudnnTensorStruct *x, int T, bool grow, cudnnStatus_t *st) {
  *st = CUDNN_STATUS_BAD_PARAM;
  if (!r || !r->set || !x || x->nb < 2 || x->dtype != CUDNN_DATA_FLOAT || T < 1) return NULL;
  const std::pair<int, int> key(x->dim[0], x->dim[1]);
  std::map<std::pair<int, int>, std::pair<b200rnnPlan_t, int> >::iterator it = r->plans.find(key);
  if (it != r->plans.end() && it->second.second >= T) {
    *st = CUDNN_STATUS_SUCCESS;
    return it->second.first;
  }
  if (it != r->plans.end() && !grow) return NULL;  // buffers were sized for a shorter sequence
  b200rnnPlan_t p = NULL;
  b200rnnStatus_t s = b200rnnCreatePlan(&p, (b200rnnMode_t)r->mode, r->dirs == 2, r->layers, key.second,
                                        r->hidden, key.first, T, (b200rnnMath_t)r->math);
  if (s != B200RNN_STATUS_SUCCESS) {
    *st = to_cudnn(s);
    return NULL;
  }
  if (it != r->plans.end()) b200rnnDestroyPlan(it->second.first);
  r->plans[key] = std::make_pair(p, T);
  *st = CUDNN_STATUS_SUCCESS;
  return p;
}

}  // namespace

extern "C" {

void b200cudnnSetMath(int math) { g_math = math ? 1 : 0; }

cudnnStatus_t cudnnCreate(cudnnHandle_t *h) {
  if (!h) return CUDNN_STATUS_BAD_PARAM;
  *h = new (std::nothrow) cudnnContext();
  return *h ? CUDNN_STATUS_SUCCESS : CUDNN_STATUS_ALLOC_FAILED;
}
cudnnStatus_t cudnnDestroy(cudnnHandle_t h) {
  delete h;
  return CUDNN_STATUS_SUCCESS;
}
const char *cudnnGetErrorString(cudnnStatus_t s) {
  switch (s) {
    case CUDNN_STATUS_SUCCESS: return "CUDNN_STATUS_SUCCESS";
    case CUDNN_STATUS_NOT_INITIALIZED: return "CUDNN_STATUS_NOT_INITIALIZED";
    case CUDNN_STATUS_ALLOC_FAILED: return "CUDNN_STATUS_ALLOC_FAILED";
    case CUDNN_STATUS_BAD_PARAM: return "CUDNN_STATUS_BAD_PARAM";
    case CUDNN_STATUS_INTERNAL_ERROR: return "CUDNN_STATUS_INTERNAL_ERROR";
    case CUDNN_STATUS_INVALID_VALUE: return "CUDNN_STATUS_INVALID_VALUE";
    case CUDNN_STATUS_ARCH_MISMATCH: return "CUDNN_STATUS_ARCH_MISMATCH";
    case CUDNN_STATUS_MAPPING_ERROR: return "CUDNN_STATUS_MAPPING_ERROR";
    case CUDNN_STATUS_EXECUTION_FAILED: return "CUDNN_STATUS_EXECUTION_FAILED";
    case CUDNN_STATUS_NOT_SUPPORTED: return "CUDNN_STATUS_NOT_SUPPORTED";
    case CUDNN_STATUS_LICENSE_ERROR: return "CUDNN_STATUS_LICENSE_ERROR";
  }
  return "CUDNN_UNKNOWN_STATUS";
}

// ---- tensor descriptors
cudnnStatus_t cudnnCreateTensorDescriptor(cudnnTensorDescriptor_t *d) {
  if (!d) return CUDNN_STATUS_BAD_PARAM;
  *d = new (std::nothrow) cudnnTensorStruct();
  return *d ? CUDNN_STATUS_SUCCESS : CUDNN_STATUS_ALLOC_FAILED;
}
cudnnStatus_t cudnnSetTensorNdDescriptor(cudnnTensorDescriptor_t d, cudnnDataType_t t, int nb, const int dimA[],
                                         const int strideA[]) {
  if (!d || !dimA || !strideA || nb < 1 || nb > CUDNN_DIM_MAX) return CUDNN_STATUS_BAD_PARAM;
  d->dtype = t;
  d->nb = nb;
  for (int i = 0; i < nb; i++) {
    d->dim[i] = dimA[i];
    d->stride[i] = strideA[i];
  }
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnGetTensorNdDescriptor(const cudnnTensorDescriptor_t d, int req, cudnnDataType_t *t, int *nb,
                                         int dimA[], int strideA[]) {
  if (!d) return CUDNN_STATUS_BAD_PARAM;
  if (t) *t = d->dtype;
  if (nb) *nb = d->nb;
  for (int i = 0; i < d->nb && i < req; i++) {
    if (dimA) dimA[i] = d->dim[i];
    if (strideA) strideA[i] = d->stride[i];
  }
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnDestroyTensorDescriptor(cudnnTensorDescriptor_t d) {
  delete d;
  return CUDNN_STATUS_SUCCESS;
}

// ---- filter descriptors
cudnnStatus_t cudnnCreateFilterDescriptor(cudnnFilterDescriptor_t *d) {
  if (!d) return CUDNN_STATUS_BAD_PARAM;
  *d = new (std::nothrow) cudnnFilterStruct();
  return *d ? CUDNN_STATUS_SUCCESS : CUDNN_STATUS_ALLOC_FAILED;
}
cudnnStatus_t cudnnSetFilterNdDescriptor(cudnnFilterDescriptor_t d, cudnnDataType_t t, cudnnTensorFormat_t f, int nb,
                                         const int dimA[]) {
  if (!d || !dimA || nb < 1 || nb > CUDNN_DIM_MAX) return CUDNN_STATUS_BAD_PARAM;
  d->dtype = t;
  d->format = f;
  d->nb = nb;
  for (int i = 0; i < nb; i++) d->dim[i] = dimA[i];
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnGetFilterNdDescriptor(const cudnnFilterDescriptor_t d, int req, cudnnDataType_t *t,
                                         cudnnTensorFormat_t *f, int *nb, int dimA[]) {
  if (!d) return CUDNN_STATUS_BAD_PARAM;
  if (t) *t = d->dtype;
  if (f) *f = d->format;
  if (nb) *nb = d->nb;
  for (int i = 0; i < d->nb && i < req; i++)
    if (dimA) dimA[i] = d->dim[i];
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnSetFilterNdDescriptor_v3(cudnnFilterDescriptor_t d, cudnnDataType_t t, int nb, const int dimA[]) {
  return cudnnSetFilterNdDescriptor(d, t, CUDNN_TENSOR_NCHW, nb, dimA);
}
cudnnStatus_t cudnnGetFilterNdDescriptor_v3(const cudnnFilterDescriptor_t d, int req, cudnnDataType_t *t, int *nb,
                                            int dimA[]) {
  return cudnnGetFilterNdDescriptor(d, req, t, NULL, nb, dimA);
}
cudnnStatus_t cudnnDestroyFilterDescriptor(cudnnFilterDescriptor_t d) {
  delete d;
  return CUDNN_STATUS_SUCCESS;
}

// ---- convolution descriptors (bookkeeping only)
cudnnStatus_t cudnnCreateConvolutionDescriptor(cudnnConvolutionDescriptor_t *d) {
  if (!d) return CUDNN_STATUS_BAD_PARAM;
  *d = new (std::nothrow) cudnnConvolutionStruct();
  return *d ? CUDNN_STATUS_SUCCESS : CUDNN_STATUS_ALLOC_FAILED;
}
cudnnStatus_t cudnnSetConvolutionNdDescriptor(cudnnConvolutionDescriptor_t d, int n, const int padA[],
                                              const int strideA[], const int upscaleA[], cudnnConvolutionMode_t mode,
                                              cudnnDataType_t t) {
  if (!d || n < 0 || n > CUDNN_DIM_MAX) return CUDNN_STATUS_BAD_PARAM;
  d->n = n;
  for (int i = 0; i < n; i++) {
    d->pad[i] = padA ? padA[i] : 0;
    d->stride[i] = strideA ? strideA[i] : 1;
    d->upscale[i] = upscaleA ? upscaleA[i] : 1;
  }
  d->mode = mode;
  d->dtype = t;
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnGetConvolutionNdDescriptor(const cudnnConvolutionDescriptor_t d, int req, int *n, int padA[],
                                              int strideA[], int upscaleA[], cudnnConvolutionMode_t *mode,
                                              cudnnDataType_t *t) {
  if (!d) return CUDNN_STATUS_BAD_PARAM;
  if (n) *n = d->n;
  for (int i = 0; i < d->n && i < req; i++) {
    if (padA) padA[i] = d->pad[i];
    if (strideA) strideA[i] = d->stride[i];
    if (upscaleA) upscaleA[i] = d->upscale[i];
  }
  if (mode) *mode = d->mode;
  if (t) *t = d->dtype;
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnDestroyConvolutionDescriptor(cudnnConvolutionDescriptor_t d) {
  delete d;
  return CUDNN_STATUS_SUCCESS;
}

// ---- dropout (the reference only ever sets p = 0)
cudnnStatus_t cudnnCreateDropoutDescriptor(cudnnDropoutDescriptor_t *d) {
  if (!d) return CUDNN_STATUS_BAD_PARAM;
  *d = new (std::nothrow) cudnnDropoutStruct();
  return *d ? CUDNN_STATUS_SUCCESS : CUDNN_STATUS_ALLOC_FAILED;
}
cudnnStatus_t cudnnDropoutGetStatesSize(cudnnHandle_t, size_t *bytes) {
  if (!bytes) return CUDNN_STATUS_BAD_PARAM;
  *bytes = 1024;  // nothing is stored; non-zero and a multiple of sizeof(float) as the caller asserts
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnSetDropoutDescriptor(cudnnDropoutDescriptor_t d, cudnnHandle_t, float dropout, void *, size_t,
                                        unsigned long long) {
  if (!d) return CUDNN_STATUS_BAD_PARAM;
  if (dropout != 0.f) return CUDNN_STATUS_NOT_SUPPORTED;
  d->p = dropout;
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnDestroyDropoutDescriptor(cudnnDropoutDescriptor_t d) {
  delete d;
  return CUDNN_STATUS_SUCCESS;
}

// ---- RNN descriptor
cudnnStatus_t cudnnCreateRNNDescriptor(cudnnRNNDescriptor_t *d) {
  if (!d) return CUDNN_STATUS_BAD_PARAM;
  *d = new (std::nothrow) cudnnRNNStruct();
  if (!*d) return CUDNN_STATUS_ALLOC_FAILED;
  (*d)->set = false;
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnSetRNNDescriptor(cudnnRNNDescriptor_t d, int hiddenSize, int numLayers, cudnnDropoutDescriptor_t,
                                    cudnnRNNInputMode_t inputMode, cudnnDirectionMode_t direction, cudnnRNNMode_t mode,
                                    cudnnDataType_t dataType) {
  if (!d || hiddenSize < 1 || numLayers < 1) return CUDNN_STATUS_BAD_PARAM;
  if (inputMode != CUDNN_LINEAR_INPUT || dataType != CUDNN_DATA_FLOAT) return CUDNN_STATUS_NOT_SUPPORTED;
  d->hidden = hiddenSize;
  d->layers = numLayers;
  d->dirs = direction == CUDNN_BIDIRECTIONAL ? 2 : 1;
  d->mode = (int)mode;
  d->math = default_math();
  d->set = true;
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnDestroyRNNDescriptor(cudnnRNNDescriptor_t d) {
  if (d)
    for (std::map<std::pair<int, int>, std::pair<b200rnnPlan_t, int> >::iterator it = d->plans.begin();
         it != d->plans.end(); ++it)
      b200rnnDestroyPlan(it->second.first);
  delete d;
  return CUDNN_STATUS_SUCCESS;
}

cudnnStatus_t cudnnGetRNNParamsSize(cudnnHandle_t, const cudnnRNNDescriptor_t r, const cudnnTensorDescriptor_t x,
                                    size_t *bytes, cudnnDataType_t t) {
  if (!bytes || t != CUDNN_DATA_FLOAT) return CUDNN_STATUS_BAD_PARAM;
  cudnnStatus_t st;
  b200rnnPlan_t p = plan_for(r, x, 1, true, &st);
  if (!p) return st;
  size_t n = 0;
  b200rnnGetParamCount(p, &n);
  *bytes = n * sizeof(float);
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnGetRNNWorkspaceSize(cudnnHandle_t, const cudnnRNNDescriptor_t r, const int T,
                                       const cudnnTensorDescriptor_t *x, size_t *bytes) {
  if (!bytes || !x) return CUDNN_STATUS_BAD_PARAM;
  cudnnStatus_t st;
  b200rnnPlan_t p = plan_for(r, x[0], T, true, &st);
  if (!p) return st;
  return to_cudnn(b200rnnGetWorkspaceSize(p, bytes));
}
cudnnStatus_t cudnnGetRNNTrainingReserveSize(cudnnHandle_t, const cudnnRNNDescriptor_t r, const int T,
                                             const cudnnTensorDescriptor_t *x, size_t *bytes) {
  if (!bytes || !x) return CUDNN_STATUS_BAD_PARAM;
  cudnnStatus_t st;
  b200rnnPlan_t p = plan_for(r, x[0], T, true, &st);
  if (!p) return st;
  return to_cudnn(b200rnnGetReserveSize(p, bytes));
}

static cudnnStatus_t lin_layer(const cudnnRNNDescriptor_t r, int layer, const cudnnTensorDescriptor_t x, const void *w,
                               int lin, int is_bias, cudnnFilterDescriptor_t out, void **ptr) {
  if (!out || !ptr || !w) return CUDNN_STATUS_BAD_PARAM;
  cudnnStatus_t st;
  b200rnnPlan_t p = plan_for(r, x, 1, true, &st);
  if (!p) return st;
  size_t off = 0;
  int rows = 0, cols = 0;
  b200rnnStatus_t s = b200rnnLocateParam(p, layer, lin, is_bias, &off, &rows, &cols);
  if (s != B200RNN_STATUS_SUCCESS) return to_cudnn(s);
  const int dims[3] = {1, rows, cols};
  cudnnSetFilterNdDescriptor(out, CUDNN_DATA_FLOAT, CUDNN_TENSOR_NCHW, 3, dims);
  *ptr = const_cast<float *>(static_cast<const float *>(w)) + off;
  return CUDNN_STATUS_SUCCESS;
}
cudnnStatus_t cudnnGetRNNLinLayerMatrixParams(cudnnHandle_t, const cudnnRNNDescriptor_t r, const int layer,
                                              const cudnnTensorDescriptor_t x, const cudnnFilterDescriptor_t,
                                              const void *w, const int lin, cudnnFilterDescriptor_t out, void **ptr) {
  return lin_layer(r, layer, x, w, lin, 0, out, ptr);
}
cudnnStatus_t cudnnGetRNNLinLayerBiasParams(cudnnHandle_t, const cudnnRNNDescriptor_t r, const int layer,
                                            const cudnnTensorDescriptor_t x, const cudnnFilterDescriptor_t,
                                            const void *w, const int lin, cudnnFilterDescriptor_t out, void **ptr) {
  return lin_layer(r, layer, x, w, lin, 1, out, ptr);
}

// ---- compute
cudnnStatus_t cudnnRNNForwardInference(cudnnHandle_t, const cudnnRNNDescriptor_t r, const int T,
                                       const cudnnTensorDescriptor_t *xDesc, const void *x,
                                       const cudnnTensorDescriptor_t, const void *, const cudnnTensorDescriptor_t,
                                       const void *, const cudnnFilterDescriptor_t, const void *w,
                                       const cudnnTensorDescriptor_t *, void *y, const cudnnTensorDescriptor_t, void *,
                                       const cudnnTensorDescriptor_t, void *, void *workspace, size_t) {
  if (!xDesc) return CUDNN_STATUS_BAD_PARAM;
  cudnnStatus_t st;
  b200rnnPlan_t p = plan_for(r, xDesc[0], T, false, &st);
  if (!p) return st;
  return to_cudnn(b200rnnForward(p, T, static_cast<const float *>(x), static_cast<const float *>(w),
                                 static_cast<float *>(y), workspace, NULL, NULL));
}
cudnnStatus_t cudnnRNNForwardTraining(cudnnHandle_t, const cudnnRNNDescriptor_t r, const int T,
                                      const cudnnTensorDescriptor_t *xDesc, const void *x,
                                      const cudnnTensorDescriptor_t, const void *, const cudnnTensorDescriptor_t,
                                      const void *, const cudnnFilterDescriptor_t, const void *w,
                                      const cudnnTensorDescriptor_t *, void *y, const cudnnTensorDescriptor_t, void *,
                                      const cudnnTensorDescriptor_t, void *, void *workspace, size_t, void *reserve,
                                      size_t) {
  if (!xDesc || !reserve) return CUDNN_STATUS_BAD_PARAM;
  cudnnStatus_t st;
  b200rnnPlan_t p = plan_for(r, xDesc[0], T, false, &st);
  if (!p) return st;
  return to_cudnn(b200rnnForward(p, T, static_cast<const float *>(x), static_cast<const float *>(w),
                                 static_cast<float *>(y), workspace, reserve, NULL));
}
cudnnStatus_t cudnnRNNBackwardData(cudnnHandle_t, const cudnnRNNDescriptor_t r, const int T,
                                   const cudnnTensorDescriptor_t *, const void *y, const cudnnTensorDescriptor_t *,
                                   const void *dy, const cudnnTensorDescriptor_t, const void *,
                                   const cudnnTensorDescriptor_t, const void *, const cudnnFilterDescriptor_t,
                                   const void *w, const cudnnTensorDescriptor_t, const void *,
                                   const cudnnTensorDescriptor_t, const void *, const cudnnTensorDescriptor_t *dxDesc,
                                   void *dx, const cudnnTensorDescriptor_t, void *, const cudnnTensorDescriptor_t,
                                   void *, void *workspace, size_t, const void *reserve, size_t) {
  if (!dxDesc || !reserve) return CUDNN_STATUS_BAD_PARAM;
  cudnnStatus_t st;
  b200rnnPlan_t p = plan_for(r, dxDesc[0], T, false, &st);
  if (!p) return st;
  return to_cudnn(b200rnnBackwardData(p, T, static_cast<const float *>(y), static_cast<const float *>(dy),
                                      static_cast<const float *>(w), static_cast<float *>(dx), workspace,
                                      const_cast<void *>(reserve), NULL));
}
cudnnStatus_t cudnnRNNBackwardWeights(cudnnHandle_t, const cudnnRNNDescriptor_t r, const int T,
                                      const cudnnTensorDescriptor_t *xDesc, const void *x,
                                      const cudnnTensorDescriptor_t, const void *, const cudnnTensorDescriptor_t *,
                                      const void *y, const void *workspace, size_t, const cudnnFilterDescriptor_t,
                                      void *dw, const void *reserve, size_t) {
  if (!xDesc || !reserve) return CUDNN_STATUS_BAD_PARAM;
  cudnnStatus_t st;
  b200rnnPlan_t p = plan_for(r, xDesc[0], T, false, &st);
  if (!p) return st;
  return to_cudnn(b200rnnBackwardWeights(p, T, static_cast<const float *>(x), static_cast<const float *>(y),
                                         static_cast<float *>(dw), const_cast<void *>(workspace),
                                         const_cast<void *>(reserve), NULL));
}

}  // extern "C"
