// integration/kaldi/harness/component_harness.cc
//
// Test driver that is COMPILED AND LINKED against the reference's own, unmodified objects
// (src/cudamatrix, src/nnet2, src/ctc, src/hmm, src/tree; recipe: oracle/ref/Makefile) with
// kaldi_ctc_b200/libb200cudnn.so standing in for -lcudnn and libb200ctc.so for -lwarpctc.  It drives the
// reference's real host code on the GPU and dumps raw little-endian floats for tests/ to compare with
// the oracle:
//
//   component <config-line> <B> <w.f32> <x.f32> <dy.f32> <out-prefix>
//       CuDNNRecurrentComponent::InitFromString (nnet-cudnn-component.cc:72-98) -> InitMiniBatch(B) ->
//       Propagate (:508-556) -> Backprop (:558-610) with to_update = the component itself.
//       Writes <out-prefix>.{y,dx,w}.f32 (w = filter_params_ after the clipped update).
//   make-model <nnet-in> <num-phones> <model-out>
//       wraps an Nnet (as written by kaldi_ctc_b200/model_io.py) into what nnet2-ctc-train-simple reads:
//       a (monophone) TransitionModel followed by an AmNnet (ctcbin/nnet2-ctc-train-simple.cc:86-93).
//   extract-nnet <model-in> <nnet-out>
//       the inverse, so that model_io.read_nnet can read what the reference's trainer wrote.
//   dump-egs <egs-rspecifier> <out-prefix>            (CPU only)
//       reads an archive with the reference's SequentialNnetCtcExampleReader (ctc-nnet-example.cc:46-60) and
//       writes <out-prefix>.meta.txt ("key rows cols left_context nlabels labels...") plus the frames as the
//       reference's CompressedMatrix::CopyToMat decompresses them (<out-prefix>.<i>.frames.f32).
//   format-input <nnet-in> <egs-rspecifier> <B> <out.f32>   (CPU only)
//       kaldi::ctc::FormatNnetInput (ctc-nnet-update.cc:351-424) on the first B examples.
//   compress <in.f32> <rows> <cols> <out.cm>          (CPU only)
//       CompressedMatrix(Matrix) (compressed-matrix.cc:41-121) written in binary mode.
//   step <nnet-in> <egs-rspecifier> <B> <out-prefix> [update]
//       NnetCtcUpdater::ComputeForMinibatch (ctc-nnet-update.cc:94-112) on the first B examples:
//       FormatNnetInput, Propagate, ComputeObjfAndDeriv (the warp-ctc call), Backprop.  Writes
//       <out-prefix>.objf.txt ("tot_objf tot_accuracy"), <out-prefix>.output.f32 (the logits) and, with
//       `update`, <out-prefix>.nnet (the updated network).
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "base/kaldi-common.h"
#include "ctc/ctc-nnet-update.h"
#include "ctc/ctc-transition-model.h"
#include "cudamatrix/cu-device.h"
#include "hmm/hmm-topology.h"
#include "nnet2/am-nnet.h"
#include "nnet2/nnet-cudnn-component.h"
#include "tree/context-dep.h"
#include "util/common-utils.h"

namespace {

using namespace kaldi;

std::vector<float> ReadFloats(const std::string &path) {
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f) KALDI_ERR << "cannot open " << path;
  f.seekg(0, std::ios::end);
  const size_t n = static_cast<size_t>(f.tellg()) / sizeof(float);
  f.seekg(0);
  std::vector<float> v(n);
  f.read(reinterpret_cast<char *>(v.data()), n * sizeof(float));
  return v;
}

void WriteFloats(const std::string &path, const float *p, size_t n) {
  std::ofstream f(path.c_str(), std::ios::binary);
  f.write(reinterpret_cast<const char *>(p), n * sizeof(float));
  if (!f) KALDI_ERR << "cannot write " << path;
}

void DumpCuMatrix(const std::string &path, const CuMatrixBase<BaseFloat> &m) {
  Matrix<BaseFloat> h(m.NumRows(), m.NumCols(), kUndefined, kStrideEqualNumCols);
  m.CopyToMat(&h);
  WriteFloats(path, h.Data(), static_cast<size_t>(h.NumRows()) * h.NumCols());
}

void FillCuMatrix(const std::vector<float> &v, int32 rows, int32 cols, CuMatrix<BaseFloat> *m) {
  KALDI_ASSERT(v.size() == static_cast<size_t>(rows) * cols);
  Matrix<BaseFloat> h(rows, cols, kUndefined, kStrideEqualNumCols);
  std::copy(v.begin(), v.end(), h.Data());
  m->Resize(rows, cols, kUndefined, kStrideEqualNumCols);
  m->CopyFromMat(h);
}

int RunComponent(int argc, char **argv) {
  if (argc != 8) KALDI_ERR << "usage: component <config> <B> <w> <x> <dy> <out-prefix>";
  const std::string config = argv[2], prefix = argv[7];
  const int32 B = atoi(argv[3]);
  nnet2::CuDNNRecurrentComponent comp;
  comp.InitFromString(config);
  comp.InitMiniBatch(B);
  const std::vector<float> w = ReadFloats(argv[4]), x = ReadFloats(argv[5]), dy = ReadFloats(argv[6]);
  {  // the weights come from the test, through the component's own Vectorize/UnVectorize pair
    KALDI_ASSERT(static_cast<int32>(w.size()) == comp.GetParameterDim());
    Vector<BaseFloat> params(w.size());
    std::copy(w.begin(), w.end(), params.Data());
    comp.UnVectorize(params);
  }
  const int32 D = comp.InputDim(), O = comp.OutputDim();
  const int32 rows = static_cast<int32>(x.size()) / D;
  KALDI_ASSERT(rows % B == 0);
  CuMatrix<BaseFloat> in, out(rows, O, kSetZero, kStrideEqualNumCols), out_deriv, in_deriv(rows, D, kSetZero, kStrideEqualNumCols);
  FillCuMatrix(x, rows, D, &in);
  FillCuMatrix(dy, rows, O, &out_deriv);
  nnet2::ChunkInfo in_info(D, 1, 0, rows - 1), out_info(O, 1, 0, rows - 1);
  comp.Propagate(in_info, out_info, in, &out);
  DumpCuMatrix(prefix + ".y.f32", out);
  comp.Backprop(in_info, out_info, in, out, out_deriv, &comp, &in_deriv);
  DumpCuMatrix(prefix + ".dx.f32", in_deriv);
  Vector<BaseFloat> params(comp.GetParameterDim());
  comp.Vectorize(&params);
  WriteFloats(prefix + ".w.f32", params.Data(), params.Dim());
  std::cout << "component: rows " << rows << " B " << B << " " << comp.Info() << std::endl;
  return 0;
}

int MakeModel(int argc, char **argv) {
  if (argc != 5) KALDI_ERR << "usage: make-model <nnet-in> <num-phones> <model-out>";
  nnet2::Nnet nnet;
  {
    bool binary;
    Input ki(argv[2], &binary);
    nnet.Read(ki.Stream(), binary);
  }
  const int32 num_phones = atoi(argv[3]);
  std::ostringstream topo;  // one emitting state per phone, as the CTC recipes use (steps/ctc: 1-state topology)
  topo << "<Topology>\n<TopologyEntry>\n<ForPhones>\n";
  std::vector<int32> phones, num_pdf_classes(num_phones + 1, 1);
  for (int32 p = 1; p <= num_phones; p++) {
    phones.push_back(p);
    topo << p << " ";
  }
  topo << "\n</ForPhones>\n<State> 0 <PdfClass> 0 <Transition> 0 0.5 <Transition> 1 0.5 </State>\n<State> 1 </State>\n"
       << "</TopologyEntry>\n</Topology>\n";
  HmmTopology hmm_topo;
  std::istringstream is(topo.str());
  hmm_topo.Read(is, false);
  ContextDependency *ctx_dep = MonophoneContextDependency(phones, num_pdf_classes);
  ctc::CtcTransitionModel trans_model(*ctx_dep, hmm_topo);
  KALDI_ASSERT(trans_model.NumPdfs() == nnet.OutputDim());
  nnet2::AmNnet am_nnet(nnet);
  Output ko(argv[4], true);
  trans_model.Write(ko.Stream(), true);
  am_nnet.Write(ko.Stream(), true);
  delete ctx_dep;
  std::cout << "make-model: " << nnet.NumComponents() << " components, " << trans_model.NumPdfs() << " pdfs" << std::endl;
  return 0;
}

int ExtractNnet(int argc, char **argv) {
  if (argc != 4) KALDI_ERR << "usage: extract-nnet <model-in> <nnet-out>";
  ctc::CtcTransitionModel trans_model;
  nnet2::AmNnet am_nnet;
  {
    bool binary;
    Input ki(argv[2], &binary);
    trans_model.Read(ki.Stream(), binary);
    am_nnet.Read(ki.Stream(), binary);
  }
  Output ko(argv[3], true);
  am_nnet.GetNnet().Write(ko.Stream(), true);
  return 0;
}

int DumpEgs(int argc, char **argv) {
  if (argc != 4) KALDI_ERR << "usage: dump-egs <egs-rspecifier> <out-prefix>";
  const std::string prefix = argv[3];
  ctc::SequentialNnetCtcExampleReader reader(argv[2]);
  std::ofstream meta((prefix + ".meta.txt").c_str());
  int32 i = 0;
  for (; !reader.Done(); reader.Next(), i++) {
    const ctc::NnetCtcExample &eg = reader.Value();
    Matrix<BaseFloat> frames(eg.input_frames.NumRows(), eg.input_frames.NumCols(), kUndefined, kStrideEqualNumCols);
    eg.input_frames.CopyToMat(&frames);
    std::ostringstream name;
    name << prefix << "." << i << ".frames.f32";
    WriteFloats(name.str(), frames.Data(), static_cast<size_t>(frames.NumRows()) * frames.NumCols());
    meta << reader.Key() << " " << frames.NumRows() << " " << frames.NumCols() << " " << eg.left_context << " "
         << eg.labels.size();
    for (size_t k = 0; k < eg.labels.size(); k++) meta << " " << eg.labels[k];
    meta << " spk " << eg.spk_info.Dim();
    for (int32 k = 0; k < eg.spk_info.Dim(); k++) meta << " " << eg.spk_info(k);
    meta << "\n";
  }
  std::cout << "dump-egs: " << i << " examples" << std::endl;
  return 0;
}

int FormatInput(int argc, char **argv) {
  if (argc != 6) KALDI_ERR << "usage: format-input <nnet-in> <egs-rspecifier> <B> <out.f32>";
  nnet2::Nnet nnet;
  {
    bool binary;
    Input ki(argv[2], &binary);
    nnet.Read(ki.Stream(), binary);
  }
  const int32 B = atoi(argv[4]);
  std::vector<ctc::NnetCtcExample> egs;
  ctc::SequentialNnetCtcExampleReader reader(argv[3]);
  for (; !reader.Done() && static_cast<int32>(egs.size()) < B; reader.Next()) egs.push_back(reader.Value());
  Matrix<BaseFloat> input;
  ctc::FormatNnetInput(nnet, egs, &input);
  Matrix<BaseFloat> packed(input.NumRows(), input.NumCols(), kUndefined, kStrideEqualNumCols);
  packed.CopyFromMat(input);
  WriteFloats(argv[5], packed.Data(), static_cast<size_t>(packed.NumRows()) * packed.NumCols());
  std::cout << "format-input: " << packed.NumRows() << " x " << packed.NumCols() << std::endl;
  return 0;
}

int Compress(int argc, char **argv) {
  if (argc != 6) KALDI_ERR << "usage: compress <in.f32> <rows> <cols> <out.cm>";
  const int32 rows = atoi(argv[3]), cols = atoi(argv[4]);
  const std::vector<float> v = ReadFloats(argv[2]);
  KALDI_ASSERT(v.size() == static_cast<size_t>(rows) * cols);
  Matrix<BaseFloat> m(rows, cols, kUndefined, kStrideEqualNumCols);
  std::copy(v.begin(), v.end(), m.Data());
  CompressedMatrix cm(m);
  Output ko(argv[5], true, false);  // binary, no "\0B" header: exactly CompressedMatrix::Write's bytes
  cm.Write(ko.Stream(), true);
  return 0;
}

int RunStep(int argc, char **argv) {
  if (argc != 6 && argc != 7) KALDI_ERR << "usage: step <nnet-in> <egs-rspecifier> <B> <out-prefix> [update]";
  nnet2::Nnet nnet;
  {
    bool binary;
    Input ki(argv[2], &binary);
    nnet.Read(ki.Stream(), binary);
  }
  const int32 B = atoi(argv[4]);
  const std::string prefix = argv[5];
  const bool update = argc == 7;
  std::vector<ctc::NnetCtcExample> egs;
  ctc::SequentialNnetCtcExampleReader reader(argv[3]);
  for (; !reader.Done() && static_cast<int32>(egs.size()) < B; reader.Next()) egs.push_back(reader.Value());
  KALDI_ASSERT(static_cast<int32>(egs.size()) == B);
  ctc::NnetCtcUpdater updater(nnet, update ? &nnet : NULL);
  double tot_accuracy = 0.0;
  const double objf = updater.ComputeForMinibatch(egs, &tot_accuracy);
  CuMatrix<BaseFloat> output;
  updater.GetOutput(&output);
  DumpCuMatrix(prefix + ".output.f32", output);
  {
    std::ofstream f((prefix + ".objf.txt").c_str());
    f.precision(12);
    f << objf << " " << tot_accuracy << "\n";
  }
  if (update) {
    Output ko(prefix + ".nnet", true);
    nnet.Write(ko.Stream(), true);
  }
  std::cout << "step: objf " << objf << " accuracy " << tot_accuracy << " rows " << output.NumRows() << std::endl;
  return 0;
}

}  // namespace

int main(int argc, char **argv) {
  try {
    if (argc < 2) KALDI_ERR << "usage: ref_component_harness component|make-model|extract-nnet|step ...";
    const std::string cmd = argv[1];
    if (cmd == "make-model") return MakeModel(argc, argv);
    if (cmd == "extract-nnet") return ExtractNnet(argc, argv);
    if (cmd == "dump-egs") return DumpEgs(argc, argv);
    if (cmd == "format-input") return FormatInput(argc, argv);
    if (cmd == "compress") return Compress(argc, argv);
#if HAVE_CUDA == 1
    kaldi::CuDevice::Instantiate().SelectGpuId("yes");
#endif
    if (cmd == "component") return RunComponent(argc, argv);
    if (cmd == "step") return RunStep(argc, argv);
    KALDI_ERR << "unknown command " << cmd;
    return 2;
  } catch (const std::exception &e) {
    std::cerr << e.what() << '\n';
    return 1;
  }
}
