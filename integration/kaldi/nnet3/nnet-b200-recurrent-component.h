// integration/kaldi/nnet3/nnet-b200-recurrent-component.h
//
// An nnet3::Component in front of the recurrent C ABI (include/b200rnn.h): what a kaldi-ctc maintainer
// adds to src/nnet3/ to run the LSTM/GRU/RNN layers of an nnet3 network on libb200rnn.so, with the
// Propagate/Backprop surface of src/nnet3/nnet-component-itf.h:116-165.  It mirrors the
// nnet2 CuDNNRecurrentComponent (src/nnet2/nnet-cudnn-component.{h,cc}): same config keys, same packed
// <FilterParams> blob, same clip-then-update rule (:640-657), so models move between the two.
//
// Row layout: nnet3 sorts Index by (t, n, x) (src/nnet3/nnet-common.h:55-61), i.e. row = t*B + n --
// exactly the time-major layout of the kernels.  The component only asks (ReorderIndexes /
// PrecomputeIndexes) that the rows it is given form a full T x B rectangle in that order.
#ifndef KALDI_NNET3_NNET_B200_RECURRENT_COMPONENT_H_
#define KALDI_NNET3_NNET_B200_RECURRENT_COMPONENT_H_

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "b200rnn.h"
#include "nnet3/nnet-component-itf.h"
#include "nnet3/nnet-parse.h"

namespace kaldi {
namespace nnet3 {

class B200RecurrentPrecomputedIndexes: public ComponentPrecomputedIndexes {
 public:
  int32 num_frames, minibatch;
  B200RecurrentPrecomputedIndexes(): num_frames(0), minibatch(0) { }
  virtual ComponentPrecomputedIndexes *Copy() const { return new B200RecurrentPrecomputedIndexes(*this); }
  virtual void Write(std::ostream &os, bool binary) const {
    WriteToken(os, binary, "<B200RecurrentPrecomputedIndexes>");
    WriteBasicType(os, binary, num_frames);
    WriteBasicType(os, binary, minibatch);
    WriteToken(os, binary, "</B200RecurrentPrecomputedIndexes>");
  }
  virtual void Read(std::istream &is, bool binary) {
    ExpectOneOrTwoTokens(is, binary, "<B200RecurrentPrecomputedIndexes>", "<NumFrames>");
    ReadBasicType(is, binary, &num_frames);
    ReadBasicType(is, binary, &minibatch);
    ExpectToken(is, binary, "</B200RecurrentPrecomputedIndexes>");
  }
  virtual std::string Type() const { return "B200RecurrentPrecomputedIndexes"; }
};

class B200RecurrentComponent: public UpdatableComponent {
 public:
  B200RecurrentComponent(): input_dim_(0), hidden_dim_(0), num_layers_(1), rnn_mode_(B200RNN_LSTM),
      bidirectional_(true), max_seq_length_(2000), clip_gradient_(5.0), math_(B200RNN_MATH_FP32) { }
  B200RecurrentComponent(const B200RecurrentComponent &o):
      UpdatableComponent(o), input_dim_(o.input_dim_), hidden_dim_(o.hidden_dim_), num_layers_(o.num_layers_),
      rnn_mode_(o.rnn_mode_), bidirectional_(o.bidirectional_), max_seq_length_(o.max_seq_length_),
      clip_gradient_(o.clip_gradient_), math_(o.math_), filter_params_(o.filter_params_) { }
  virtual ~B200RecurrentComponent() { DestroyPlans(); }

  virtual int32 InputDim() const { return input_dim_; }
  virtual int32 OutputDim() const { return hidden_dim_ * (bidirectional_ ? 2 : 1); }
  virtual std::string Type() const { return "B200RecurrentComponent"; }
  virtual int32 Properties() const {
    return kUpdatableComponent | kBackpropNeedsInput | kBackpropNeedsOutput | kReordersIndexes;
  }
  virtual Component *Copy() const { return new B200RecurrentComponent(*this); }

  // config keys of the nnet2 component (nnet-cudnn-component.cc:411-470)
  virtual void InitFromConfig(ConfigLine *cfl) {
    bool ok = cfl->GetValue("input-dim", &input_dim_) && cfl->GetValue("output-dim", &hidden_dim_);
    BaseFloat param_stddev = 0.02, bias_stddev = 0.2;
    cfl->GetValue("learning-rate", &learning_rate_);
    cfl->GetValue("num-layers", &num_layers_);
    cfl->GetValue("rnn-mode", &rnn_mode_);
    cfl->GetValue("bidirectional", &bidirectional_);
    cfl->GetValue("max-seq-length", &max_seq_length_);
    cfl->GetValue("clip-gradient", &clip_gradient_);
    cfl->GetValue("param-stddev", &param_stddev);
    cfl->GetValue("bias-stddev", &bias_stddev);
    // exact fp32 kernels by default (what an fp32 nnet3 model expects); exact-fp32=0 opts into the tensor-core mode
    // (BF16 recurrent operands, TF32 projections; tolerance in DESIGN.md section 5)
    int32 exact = 1;
    cfl->GetValue("exact-fp32", &exact);
    math_ = exact ? B200RNN_MATH_FP32 : B200RNN_MATH_TENSOR;
    if (!ok || cfl->HasUnusedValues() || rnn_mode_ < 0 || rnn_mode_ > 3)
      KALDI_ERR << "Bad initializer " << cfl->WholeLine();
    b200rnnPlan_t plan = GetPlan(1);
    size_t n = 0;
    Check(b200rnnGetParamCount(plan, &n), "b200rnnGetParamCount");
    filter_params_.Resize(n);
    filter_params_.SetRandn();           // matrices ~ N(0, param_stddev), biases = bias_stddev (:336-408)
    filter_params_.Scale(param_stddev);
    const int32 gates = rnn_mode_ == B200RNN_LSTM ? 4 : (rnn_mode_ == B200RNN_GRU ? 3 : 1);
    const int32 nbias = num_layers_ * (bidirectional_ ? 2 : 1) * 2 * gates * hidden_dim_;
    filter_params_.Range(n - nbias, nbias).Set(bias_stddev);
  }

  // rows must be a full T x B rectangle sorted by (t, n): sort, the framework permutes the data
  virtual void ReorderIndexes(std::vector<Index> *input_indexes, std::vector<Index> *output_indexes) const {
    std::sort(input_indexes->begin(), input_indexes->end());
    std::sort(output_indexes->begin(), output_indexes->end());
  }
  virtual ComponentPrecomputedIndexes *PrecomputeIndexes(const MiscComputationInfo &misc_info,
                                                         const std::vector<Index> &input_indexes,
                                                         const std::vector<Index> &output_indexes,
                                                         bool need_backprop) const {
    KALDI_ASSERT(input_indexes == output_indexes && !input_indexes.empty());
    int32 B = 0;
    while (B < static_cast<int32>(input_indexes.size()) && input_indexes[B].t == input_indexes[0].t) B++;
    KALDI_ASSERT(input_indexes.size() % B == 0);
    const int32 T = input_indexes.size() / B;
    for (int32 t = 0; t < T; t++)
      for (int32 b = 0; b < B; b++)
        KALDI_ASSERT(input_indexes[t * B + b].t == input_indexes[t * B].t &&
                     input_indexes[t * B + b].n == input_indexes[b].n);  // rectangle, row = t*B + n
    B200RecurrentPrecomputedIndexes *ans = new B200RecurrentPrecomputedIndexes();
    ans->num_frames = T;
    ans->minibatch = B;
    return ans;
  }

  virtual void Propagate(const ComponentPrecomputedIndexes *indexes_in, const CuMatrixBase<BaseFloat> &in,
                         CuMatrixBase<BaseFloat> *out) const {
    const B200RecurrentPrecomputedIndexes *ix = dynamic_cast<const B200RecurrentPrecomputedIndexes*>(indexes_in);
    KALDI_ASSERT(ix != NULL && in.NumRows() == ix->num_frames * ix->minibatch);
    KALDI_ASSERT(in.Stride() == in.NumCols() && out->Stride() == out->NumCols());
    b200rnnPlan_t plan = GetPlan(ix->minibatch);
    EnsureBuffers(plan);
    // minibatch 1 = decoding: no reserve space (nnet-cudnn-component.cc:534-543)
    Check(b200rnnForward(plan, ix->num_frames, in.Data(), filter_params_.Data(), out->Data(), work_space_.Data(),
                         ix->minibatch == 1 ? NULL : reserve_space_.Data(), 0), "b200rnnForward");
  }

  virtual void Backprop(const std::string &debug_info, const ComponentPrecomputedIndexes *indexes_in,
                        const CuMatrixBase<BaseFloat> &in_value, const CuMatrixBase<BaseFloat> &out_value,
                        const CuMatrixBase<BaseFloat> &out_deriv, Component *to_update_in,
                        CuMatrixBase<BaseFloat> *in_deriv) const {
    const B200RecurrentPrecomputedIndexes *ix = dynamic_cast<const B200RecurrentPrecomputedIndexes*>(indexes_in);
    KALDI_ASSERT(ix != NULL && ix->minibatch > 1);
    b200rnnPlan_t plan = GetPlan(ix->minibatch);
    Check(b200rnnBackwardData(plan, ix->num_frames, out_value.Data(), out_deriv.Data(), filter_params_.Data(),
                              in_deriv ? in_deriv->Data() : NULL, work_space_.Data(), reserve_space_.Data(), 0),
          "b200rnnBackwardData");
    B200RecurrentComponent *to_update = dynamic_cast<B200RecurrentComponent*>(to_update_in);
    if (to_update != NULL) {
      CuVector<BaseFloat> grad(filter_params_.Dim());   // kSetZero: BackwardWeights accumulates
      Check(b200rnnBackwardWeights(plan, ix->num_frames, in_value.Data(), out_value.Data(), grad.Data(),
                                   work_space_.Data(), reserve_space_.Data(), 0), "b200rnnBackwardWeights");
      // clip to +-clip_gradient_, then w += lr * g (nnet-cudnn-component.cc:640-657), one fused pass
      Check(b200rnnClipAndUpdate(to_update->filter_params_.Data(), grad.Data(), grad.Dim(),
                                 to_update->learning_rate_, clip_gradient_, 0), "b200rnnClipAndUpdate");
    }
  }

  virtual void Read(std::istream &is, bool binary) {
    ExpectOneOrTwoTokens(is, binary, "<B200RecurrentComponent>", "<LearningRate>");
    ReadBasicType(is, binary, &learning_rate_);
    ExpectToken(is, binary, "<IsGradient>");    ReadBasicType(is, binary, &is_gradient_);
    ExpectToken(is, binary, "<ClipGradient>");  ReadBasicType(is, binary, &clip_gradient_);
    ExpectToken(is, binary, "<InputDim>");      ReadBasicType(is, binary, &input_dim_);
    ExpectToken(is, binary, "<HiddenDim>");     ReadBasicType(is, binary, &hidden_dim_);
    ExpectToken(is, binary, "<NumLayers>");     ReadBasicType(is, binary, &num_layers_);
    ExpectToken(is, binary, "<Bidirectional>"); ReadBasicType(is, binary, &bidirectional_);
    ExpectToken(is, binary, "<RNNMode>");       ReadBasicType(is, binary, &rnn_mode_);
    ExpectToken(is, binary, "<MaxSeqLength>");  ReadBasicType(is, binary, &max_seq_length_);
    ExpectToken(is, binary, "<FilterParams>");  filter_params_.Read(is, binary);
    ExpectToken(is, binary, "</B200RecurrentComponent>");
    DestroyPlans();
  }
  virtual void Write(std::ostream &os, bool binary) const {  // field for field the nnet2 component (:698-721)
    WriteToken(os, binary, "<B200RecurrentComponent>");
    WriteToken(os, binary, "<LearningRate>");  WriteBasicType(os, binary, learning_rate_);
    WriteToken(os, binary, "<IsGradient>");    WriteBasicType(os, binary, is_gradient_);
    WriteToken(os, binary, "<ClipGradient>");  WriteBasicType(os, binary, clip_gradient_);
    WriteToken(os, binary, "<InputDim>");      WriteBasicType(os, binary, input_dim_);
    WriteToken(os, binary, "<HiddenDim>");     WriteBasicType(os, binary, hidden_dim_);
    WriteToken(os, binary, "<NumLayers>");     WriteBasicType(os, binary, num_layers_);
    WriteToken(os, binary, "<Bidirectional>"); WriteBasicType(os, binary, bidirectional_);
    WriteToken(os, binary, "<RNNMode>");       WriteBasicType(os, binary, rnn_mode_);
    WriteToken(os, binary, "<MaxSeqLength>");  WriteBasicType(os, binary, max_seq_length_);
    WriteToken(os, binary, "<FilterParams>");  filter_params_.Write(os, binary);
    WriteToken(os, binary, "</B200RecurrentComponent>");
  }

  // UpdatableComponent surface on the flat blob
  virtual void SetZero(bool treat_as_gradient) {
    if (treat_as_gradient) { learning_rate_ = 1.0; is_gradient_ = true; }
    filter_params_.SetZero();
  }
  virtual void Scale(BaseFloat scale) { filter_params_.Scale(scale); }
  virtual void Add(BaseFloat alpha, const Component &other_in) {
    const B200RecurrentComponent *other = dynamic_cast<const B200RecurrentComponent*>(&other_in);
    KALDI_ASSERT(other != NULL);
    filter_params_.AddVec(alpha, other->filter_params_);
  }
  virtual BaseFloat DotProduct(const UpdatableComponent &other_in) const {
    const B200RecurrentComponent *other = dynamic_cast<const B200RecurrentComponent*>(&other_in);
    KALDI_ASSERT(other != NULL);
    return VecVec(filter_params_, other->filter_params_);
  }
  virtual void PerturbParams(BaseFloat stddev) {
    CuVector<BaseFloat> tmp(filter_params_.Dim());
    tmp.SetRandn();
    filter_params_.AddVec(stddev, tmp);
  }
  virtual int32 NumParameters() const { return filter_params_.Dim(); }
  virtual void Vectorize(VectorBase<BaseFloat> *params) const { filter_params_.CopyToVec(params); }
  virtual void UnVectorize(const VectorBase<BaseFloat> &params) { filter_params_.CopyFromVec(params); }

 private:
  static void Check(b200rnnStatus_t st, const char *what) {
    if (st != B200RNN_STATUS_SUCCESS) KALDI_ERR << what << " failed with status " << static_cast<int>(st);
  }
  // one plan per minibatch size, like the descriptors of the nnet2 component (:283-334)
  b200rnnPlan_t GetPlan(int32 minibatch) const {
    std::map<int32, b200rnnPlan_t>::iterator it = plans_.find(minibatch);
    if (it != plans_.end()) return it->second;
    b200rnnPlan_t plan = NULL;
    Check(b200rnnCreatePlan(&plan, static_cast<b200rnnMode_t>(rnn_mode_), bidirectional_ ? 1 : 0, num_layers_,
                            input_dim_, hidden_dim_, minibatch, max_seq_length_, math_), "b200rnnCreatePlan");
    plans_[minibatch] = plan;
    return plan;
  }
  void EnsureBuffers(b200rnnPlan_t plan) const {
    size_t ws = 0, rs = 0;
    Check(b200rnnGetWorkspaceSize(plan, &ws), "b200rnnGetWorkspaceSize");
    Check(b200rnnGetReserveSize(plan, &rs), "b200rnnGetReserveSize");
    const int32 wf = (ws + 3) / 4, rf = (rs + 3) / 4;
    if (work_space_.Dim() < wf) work_space_.Resize(wf, kUndefined);
    if (reserve_space_.Dim() < rf) reserve_space_.Resize(rf, kUndefined);
  }
  void DestroyPlans() {
    for (std::map<int32, b200rnnPlan_t>::iterator it = plans_.begin(); it != plans_.end(); ++it)
      b200rnnDestroyPlan(it->second);
    plans_.clear();
  }

  int32 input_dim_, hidden_dim_, num_layers_, rnn_mode_;
  bool bidirectional_;
  int32 max_seq_length_;
  BaseFloat clip_gradient_;
  b200rnnMath_t math_;
  CuVector<BaseFloat> filter_params_;
  mutable std::map<int32, b200rnnPlan_t> plans_;
  mutable CuVector<BaseFloat> work_space_, reserve_space_;   // Propagate is const in nnet3
};

}  // namespace nnet3
}  // namespace kaldi

#endif  // KALDI_NNET3_NNET_B200_RECURRENT_COMPONENT_H_
