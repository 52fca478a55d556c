// integration/kaldi/harness/nnet3_harness.cc
//
// Runs the nnet3 adapter (integration/kaldi/nnet3/nnet-b200-recurrent-component.h) LINKED against the
// reference's own src/nnet3 objects (nnet-component-itf, nnet-parse, nnet-common, ...; recipe:
// oracle/ref/Makefile) through the surface of src/nnet3/nnet-component-itf.h:116-165:
//   InitFromConfig -> ReorderIndexes / PrecomputeIndexes -> Propagate -> Backprop (to_update = itself)
//   -> Write / Read round trip -> Vectorize.
// usage: ref_nnet3_harness <config-line> <T> <B> <w.f32> <x.f32> <dy.f32> <out-prefix>
// Writes <out-prefix>.{y,dx,w}.f32 for tests/test_reference_linked_gpu.py.
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

#include "cudamatrix/cu-device.h"
#include "nnet-b200-recurrent-component.h"
#include "nnet3/nnet-computation-graph.h"   // MiscComputationInfo

namespace {
using namespace kaldi;
using namespace kaldi::nnet3;

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
}
void Dump(const std::string &path, const CuMatrixBase<BaseFloat> &m) {
  Matrix<BaseFloat> h(m.NumRows(), m.NumCols(), kUndefined, kStrideEqualNumCols);
  m.CopyToMat(&h);
  WriteFloats(path, h.Data(), static_cast<size_t>(h.NumRows()) * h.NumCols());
}
void Fill(const std::vector<float> &v, int32 rows, int32 cols, CuMatrix<BaseFloat> *m) {
  KALDI_ASSERT(v.size() == static_cast<size_t>(rows) * cols);
  Matrix<BaseFloat> h(rows, cols, kUndefined, kStrideEqualNumCols);
  std::copy(v.begin(), v.end(), h.Data());
  m->Resize(rows, cols, kUndefined, kStrideEqualNumCols);
  m->CopyFromMat(h);
}
}  // namespace

int main(int argc, char **argv) {
  try {
    if (argc != 8) KALDI_ERR << "usage: ref_nnet3_harness <config-line> <T> <B> <w> <x> <dy> <out-prefix>";
    CuDevice::Instantiate().SelectGpuId("yes");
    const int32 T = atoi(argv[2]), B = atoi(argv[3]);
    const std::string prefix = argv[7];
    B200RecurrentComponent comp;
    ConfigLine cfl;
    if (!cfl.ParseLine(std::string("B200RecurrentComponent ") + argv[1])) KALDI_ERR << "bad config line";
    comp.InitFromConfig(&cfl);
    const std::vector<float> w = ReadFloats(argv[4]), x = ReadFloats(argv[5]), dy = ReadFloats(argv[6]);
    {
      KALDI_ASSERT(static_cast<int32>(w.size()) == comp.NumParameters());
      Vector<BaseFloat> params(w.size());
      std::copy(w.begin(), w.end(), params.Data());
      comp.UnVectorize(params);
    }
    // the indexes as nnet3 would hand them over, deliberately NOT sorted: n-major
    std::vector<Index> in_idx, out_idx;
    for (int32 n = 0; n < B; n++)
      for (int32 t = 0; t < T; t++) in_idx.push_back(Index(n, t, 0));
    out_idx = in_idx;
    comp.ReorderIndexes(&in_idx, &out_idx);
    KALDI_ASSERT(in_idx[1].n == 1 && in_idx[1].t == 0 && in_idx[B].t == 1);   // now (t, n): row = t*B + n
    MiscComputationInfo misc;
    ComponentPrecomputedIndexes *pre = comp.PrecomputeIndexes(misc, in_idx, out_idx, true);
    const int32 D = comp.InputDim(), O = comp.OutputDim(), rows = T * B;
    CuMatrix<BaseFloat> in, out(rows, O, kSetZero, kStrideEqualNumCols), out_deriv, in_deriv(rows, D, kSetZero, kStrideEqualNumCols);
    Fill(x, rows, D, &in);
    Fill(dy, rows, O, &out_deriv);
    comp.Propagate(pre, in, &out);
    Dump(prefix + ".y.f32", out);
    comp.Backprop("harness", pre, in, out, out_deriv, &comp, &in_deriv);
    Dump(prefix + ".dx.f32", in_deriv);
    // Write -> Read round trip through the reference's token I/O, then the parameters
    std::ostringstream os;
    comp.Write(os, true);
    B200RecurrentComponent back;
    std::istringstream is(os.str());
    back.Read(is, true);
    Vector<BaseFloat> params(back.NumParameters());
    back.Vectorize(&params);
    WriteFloats(prefix + ".w.f32", params.Data(), params.Dim());
    std::cout << "nnet3 harness: " << comp.Type() << " rows " << rows << " params " << comp.NumParameters()
              << " dot " << comp.DotProduct(back) << std::endl;
    delete pre;
    return 0;
  } catch (const std::exception &e) {
    std::cerr << e.what() << '\n';
    return 1;
  }
}
