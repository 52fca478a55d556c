#!/bin/bash
# Compiles the reference's UNMODIFIED cuDNN-facing sources against include/cudnn_v5_compat/cudnn.h
# (syntax + type check, no link), and the warp-ctc call-site shape against include/ctc.h.
# Needs the reference tree (default /root/reference); used by tests/test_integration_cpu.py.
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
CXX=${CXX_CHECK:-/usr/bin/g++}
FLAGS="-std=c++11 -fsyntax-only -w -DHAVE_CUDA=1 -DHAVE_CUDNN=1 -DHAVE_CLAPACK -DKALDI_DOUBLEPRECISION=0
       -I$ROOT/include/cudnn_v5_compat -I$HERE/shim -I$REF/src -I$REF/tools/CLAPACK -I/usr/local/cuda/include"
for f in cudamatrix/cudnn-utils.cc cudamatrix/cudnn-recurrent.cc cudamatrix/cu-device.cc nnet2/nnet-cudnn-component.cc nnet2/nnet-nnet.cc; do
  echo "checking $f"
  $CXX $FLAGS "$REF/src/$f"
done
echo "checking ctc.h call-site shape"
$CXX -std=c++11 -fsyntax-only -Wall -I"$ROOT/include" "$HERE/ctc_callsite_check.cc"
echo "checking the nnet3 adapter (integration/kaldi/nnet3) against the reference's nnet3 headers"
$CXX -std=c++11 -fsyntax-only -w -DHAVE_CUDA=1 -DHAVE_CLAPACK -DKALDI_DOUBLEPRECISION=0 -I"$ROOT/include" -I"$HERE/shim" \
     -I"$HERE/nnet3" -I"$REF/src" -I"$REF/tools/CLAPACK" -I/usr/local/cuda/include "$HERE/nnet3/compile_check.cc"
echo "OK: reference sources compile against the drop-in headers"
