/*
 * include/cudnn_v5_compat/cudnn.h -- the subset of the cuDNN 5.x C API that
 * kaldi-ctc compiles against, implemented by libb200cudnn.so on top of the
 * B200-native recurrent kernels (include/b200rnn.h).  With this header at
 * <root>/include/cudnn.h and the library at <root>/lib64/libcudnn.so,
 * `configure --cudnn-root=<root>` (src/configure:188-189, makefiles/cudnn_64bit.mk)
 * builds the reference UNMODIFIED: src/cudamatrix/{cu-device,cudnn-utils,
 * cudnn-recurrent}.cc and src/nnet2/nnet-cudnn-component.cc bind to these symbols.
 *
 * Every entry point below is one the reference calls (file:line given); nothing
 * else of cuDNN is provided.  cuDNN 5 itself cannot run on sm_100 and its v5 RNN
 * API no longer exists in cuDNN 9.
 */
#ifndef B200_CUDNN_V5_COMPAT_H_
#define B200_CUDNN_V5_COMPAT_H_

#include <stddef.h>

#define CUDNN_MAJOR 5
#define CUDNN_MINOR 1
#define CUDNN_PATCHLEVEL 0
#define CUDNN_VERSION (CUDNN_MAJOR * 1000 + CUDNN_MINOR * 100 + CUDNN_PATCHLEVEL)
#define CUDNN_DIM_MAX 8 /* src/cudamatrix/cudnn-utils.cc:31-32 */

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cudnnContext *cudnnHandle_t;
typedef struct cudnnTensorStruct *cudnnTensorDescriptor_t;
typedef struct cudnnFilterStruct *cudnnFilterDescriptor_t;
typedef struct cudnnConvolutionStruct *cudnnConvolutionDescriptor_t;
typedef struct cudnnDropoutStruct *cudnnDropoutDescriptor_t;
typedef struct cudnnRNNStruct *cudnnRNNDescriptor_t;

typedef enum {
  CUDNN_STATUS_SUCCESS = 0,
  CUDNN_STATUS_NOT_INITIALIZED = 1,
  CUDNN_STATUS_ALLOC_FAILED = 2,
  CUDNN_STATUS_BAD_PARAM = 3,
  CUDNN_STATUS_INTERNAL_ERROR = 4,
  CUDNN_STATUS_INVALID_VALUE = 5,
  CUDNN_STATUS_ARCH_MISMATCH = 6,
  CUDNN_STATUS_MAPPING_ERROR = 7,
  CUDNN_STATUS_EXECUTION_FAILED = 8,
  CUDNN_STATUS_NOT_SUPPORTED = 9,
  CUDNN_STATUS_LICENSE_ERROR = 10
} cudnnStatus_t;

typedef enum { CUDNN_DATA_FLOAT = 0, CUDNN_DATA_DOUBLE = 1, CUDNN_DATA_HALF = 2 } cudnnDataType_t;
typedef enum { CUDNN_TENSOR_NCHW = 0, CUDNN_TENSOR_NHWC = 1 } cudnnTensorFormat_t;
typedef enum { CUDNN_CONVOLUTION = 0, CUDNN_CROSS_CORRELATION = 1 } cudnnConvolutionMode_t;
typedef enum { CUDNN_RNN_RELU = 0, CUDNN_RNN_TANH = 1, CUDNN_LSTM = 2, CUDNN_GRU = 3 } cudnnRNNMode_t;
typedef enum { CUDNN_UNIDIRECTIONAL = 0, CUDNN_BIDIRECTIONAL = 1 } cudnnDirectionMode_t;
typedef enum { CUDNN_LINEAR_INPUT = 0, CUDNN_SKIP_INPUT = 1 } cudnnRNNInputMode_t;

/* src/cudamatrix/cu-device.cc:220,578; src/cudamatrix/cu-common.h:55-63 */
cudnnStatus_t cudnnCreate(cudnnHandle_t *handle);
cudnnStatus_t cudnnDestroy(cudnnHandle_t handle);
const char *cudnnGetErrorString(cudnnStatus_t status);

/* src/nnet2/nnet-cudnn-component.cc:156-225; src/cudamatrix/cudnn-utils.cc:25-50 */
cudnnStatus_t cudnnCreateTensorDescriptor(cudnnTensorDescriptor_t *desc);
cudnnStatus_t cudnnSetTensorNdDescriptor(cudnnTensorDescriptor_t desc, cudnnDataType_t dataType,
                                         int nbDims, const int dimA[], const int strideA[]);
cudnnStatus_t cudnnGetTensorNdDescriptor(const cudnnTensorDescriptor_t desc, int nbDimsRequested,
                                         cudnnDataType_t *dataType, int *nbDims, int dimA[],
                                         int strideA[]);
cudnnStatus_t cudnnDestroyTensorDescriptor(cudnnTensorDescriptor_t desc);

/* src/nnet2/nnet-cudnn-component.cc:267-285,352-361; src/cudamatrix/cudnn-utils.cc:53-75 */
cudnnStatus_t cudnnCreateFilterDescriptor(cudnnFilterDescriptor_t *desc);
cudnnStatus_t cudnnSetFilterNdDescriptor(cudnnFilterDescriptor_t desc, cudnnDataType_t dataType,
                                         cudnnTensorFormat_t format, int nbDims,
                                         const int filterDimA[]);
cudnnStatus_t cudnnGetFilterNdDescriptor(const cudnnFilterDescriptor_t desc, int nbDimsRequested,
                                         cudnnDataType_t *dataType, cudnnTensorFormat_t *format,
                                         int *nbDims, int filterDimA[]);
cudnnStatus_t cudnnSetFilterNdDescriptor_v3(cudnnFilterDescriptor_t desc, cudnnDataType_t dataType,
                                            int nbDims, const int filterDimA[]);
cudnnStatus_t cudnnGetFilterNdDescriptor_v3(const cudnnFilterDescriptor_t desc, int nbDimsRequested,
                                            cudnnDataType_t *dataType, int *nbDims,
                                            int filterDimA[]);
cudnnStatus_t cudnnDestroyFilterDescriptor(cudnnFilterDescriptor_t desc);

/* src/cudamatrix/cudnn-utils.cc:78-115 (descriptor copy helper only; no convolution is computed) */
cudnnStatus_t cudnnCreateConvolutionDescriptor(cudnnConvolutionDescriptor_t *desc);
cudnnStatus_t cudnnSetConvolutionNdDescriptor(cudnnConvolutionDescriptor_t desc, int arrayLength,
                                              const int padA[], const int filterStrideA[],
                                              const int upscaleA[], cudnnConvolutionMode_t mode,
                                              cudnnDataType_t dataType);
cudnnStatus_t cudnnGetConvolutionNdDescriptor(const cudnnConvolutionDescriptor_t desc,
                                              int arrayLengthRequested, int *arrayLength, int padA[],
                                              int strideA[], int upscaleA[],
                                              cudnnConvolutionMode_t *mode, cudnnDataType_t *dataType);
cudnnStatus_t cudnnDestroyConvolutionDescriptor(cudnnConvolutionDescriptor_t desc);

/* src/nnet2/nnet-cudnn-component.cc:230-246 (dropout probability is always 0 in the reference) */
cudnnStatus_t cudnnCreateDropoutDescriptor(cudnnDropoutDescriptor_t *desc);
cudnnStatus_t cudnnDropoutGetStatesSize(cudnnHandle_t handle, size_t *sizeInBytes);
cudnnStatus_t cudnnSetDropoutDescriptor(cudnnDropoutDescriptor_t desc, cudnnHandle_t handle,
                                        float dropout, void *states, size_t stateSizeInBytes,
                                        unsigned long long seed);
cudnnStatus_t cudnnDestroyDropoutDescriptor(cudnnDropoutDescriptor_t desc);

/* src/nnet2/nnet-cudnn-component.cc:252-265 */
cudnnStatus_t cudnnCreateRNNDescriptor(cudnnRNNDescriptor_t *desc);
cudnnStatus_t cudnnSetRNNDescriptor(cudnnRNNDescriptor_t desc, int hiddenSize, int numLayers,
                                    cudnnDropoutDescriptor_t dropoutDesc,
                                    cudnnRNNInputMode_t inputMode, cudnnDirectionMode_t direction,
                                    cudnnRNNMode_t mode, cudnnDataType_t dataType);
cudnnStatus_t cudnnDestroyRNNDescriptor(cudnnRNNDescriptor_t desc);

/* src/nnet2/nnet-cudnn-component.cc:270-314 */
cudnnStatus_t cudnnGetRNNParamsSize(cudnnHandle_t handle, const cudnnRNNDescriptor_t rnnDesc,
                                    const cudnnTensorDescriptor_t xDesc, size_t *sizeInBytes,
                                    cudnnDataType_t dataType);
cudnnStatus_t cudnnGetRNNWorkspaceSize(cudnnHandle_t handle, const cudnnRNNDescriptor_t rnnDesc,
                                       const int seqLength, const cudnnTensorDescriptor_t *xDesc,
                                       size_t *sizeInBytes);
cudnnStatus_t cudnnGetRNNTrainingReserveSize(cudnnHandle_t handle,
                                             const cudnnRNNDescriptor_t rnnDesc,
                                             const int seqLength,
                                             const cudnnTensorDescriptor_t *xDesc,
                                             size_t *sizeInBytes);

/* src/nnet2/nnet-cudnn-component.cc:342-350,380-388,425-433,455-463 */
cudnnStatus_t cudnnGetRNNLinLayerMatrixParams(cudnnHandle_t handle,
                                              const cudnnRNNDescriptor_t rnnDesc, const int layer,
                                              const cudnnTensorDescriptor_t xDesc,
                                              const cudnnFilterDescriptor_t wDesc, const void *w,
                                              const int linLayerID,
                                              cudnnFilterDescriptor_t linLayerMatDesc,
                                              void **linLayerMat);
cudnnStatus_t cudnnGetRNNLinLayerBiasParams(cudnnHandle_t handle,
                                            const cudnnRNNDescriptor_t rnnDesc, const int layer,
                                            const cudnnTensorDescriptor_t xDesc,
                                            const cudnnFilterDescriptor_t wDesc, const void *w,
                                            const int linLayerID,
                                            cudnnFilterDescriptor_t linLayerBiasDesc,
                                            void **linLayerBias);

/*
 * src/cudamatrix/cudnn-recurrent.cc:26,49,75,95.  Restrictions that hold at every call the
 * reference makes: fp32; every xDesc[t] has the same batch; hx, cx, dhy, dcy are zero
 * (nnet-cudnn-component.cc:494-506) and hy, cy, dhx, dcx are never read back (they are
 * left untouched); work is enqueued on the legacy default stream.
 */
cudnnStatus_t cudnnRNNForwardInference(
    cudnnHandle_t handle, const cudnnRNNDescriptor_t rnnDesc, const int seqLength,
    const cudnnTensorDescriptor_t *xDesc, const void *x, const cudnnTensorDescriptor_t hxDesc,
    const void *hx, const cudnnTensorDescriptor_t cxDesc, const void *cx,
    const cudnnFilterDescriptor_t wDesc, const void *w, const cudnnTensorDescriptor_t *yDesc, void *y,
    const cudnnTensorDescriptor_t hyDesc, void *hy, const cudnnTensorDescriptor_t cyDesc, void *cy,
    void *workspace, size_t workSpaceSizeInBytes);
cudnnStatus_t cudnnRNNForwardTraining(
    cudnnHandle_t handle, const cudnnRNNDescriptor_t rnnDesc, const int seqLength,
    const cudnnTensorDescriptor_t *xDesc, const void *x, const cudnnTensorDescriptor_t hxDesc,
    const void *hx, const cudnnTensorDescriptor_t cxDesc, const void *cx,
    const cudnnFilterDescriptor_t wDesc, const void *w, const cudnnTensorDescriptor_t *yDesc, void *y,
    const cudnnTensorDescriptor_t hyDesc, void *hy, const cudnnTensorDescriptor_t cyDesc, void *cy,
    void *workspace, size_t workSpaceSizeInBytes, void *reserveSpace,
    size_t reserveSpaceSizeInBytes);
cudnnStatus_t cudnnRNNBackwardData(
    cudnnHandle_t handle, const cudnnRNNDescriptor_t rnnDesc, const int seqLength,
    const cudnnTensorDescriptor_t *yDesc, const void *y, const cudnnTensorDescriptor_t *dyDesc,
    const void *dy, const cudnnTensorDescriptor_t dhyDesc, const void *dhy,
    const cudnnTensorDescriptor_t dcyDesc, const void *dcy, const cudnnFilterDescriptor_t wDesc,
    const void *w, const cudnnTensorDescriptor_t hxDesc, const void *hx,
    const cudnnTensorDescriptor_t cxDesc, const void *cx, const cudnnTensorDescriptor_t *dxDesc,
    void *dx, const cudnnTensorDescriptor_t dhxDesc, void *dhx, const cudnnTensorDescriptor_t dcxDesc,
    void *dcx, void *workspace, size_t workSpaceSizeInBytes, const void *reserveSpace,
    size_t reserveSpaceSizeInBytes);
cudnnStatus_t cudnnRNNBackwardWeights(
    cudnnHandle_t handle, const cudnnRNNDescriptor_t rnnDesc, const int seqLength,
    const cudnnTensorDescriptor_t *xDesc, const void *x, const cudnnTensorDescriptor_t hxDesc,
    const void *hx, const cudnnTensorDescriptor_t *yDesc, const void *y, const void *workspace,
    size_t workSpaceSizeInBytes, const cudnnFilterDescriptor_t dwDesc, void *dw,
    const void *reserveSpace, size_t reserveSpaceSizeInBytes);

/* Selects the arithmetic of the kernels behind this API for descriptors created afterwards:
 * 0 = fp32 FMA (DEFAULT: what CUDNN_DATA_FLOAT asks for; matches an fp32 reference to ~1e-6),
 * 1 = tensor cores (BF16 recurrent operands, TF32/BF16 projections, fp32 accumulate/state): ~10x
 *     faster, outputs within the tolerance stated in DESIGN.md section 5.  Opt-in only.
 * Also settable through the environment (read once): B200_CUDNN_MATH=tensor.  Not part of cuDNN. */
void b200cudnnSetMath(int math);

#ifdef __cplusplus
}
#endif
#endif /* B200_CUDNN_V5_COMPAT_H_ */
