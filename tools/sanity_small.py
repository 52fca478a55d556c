"""Small end-to-end exercise of every kernel family (CTC, tensor/fp32/streaming recurrent, both GEMMs,
the training-step mirror) for compute-sanitizer runs:
   compute-sanitizer --tool memcheck python tools/sanity_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kaldi_ctc_b200 import ctc, nnet, rnn, synth  # noqa: E402

bt = synth.ctc_batch(5, 48, 60, 90, 5, 20, seed=9)
op = ctc.CtcLoss("cuda:0")
costs, grad = op.compute(torch.from_numpy(bt.activations).cuda(), bt.flat_labels, bt.label_lengths, bt.input_lengths)
print("ctc costs", costs[:3], float(grad.abs().sum()))
bt = synth.ctc_batch(2, 9, 1300, 1400, 600, 640, seed=10)      # P = 2 path
costs, grad = op.compute(torch.from_numpy(bt.activations).cuda(), bt.flat_labels, bt.label_lengths, bt.input_lengths)
print("ctc costs (L~620)", costs)

for math, H, B, mode in [(rnn.MATH_TENSOR, 64, 5, 2), (rnn.MATH_TENSOR, 128, 9, 3), (rnn.MATH_FP32, 32, 3, 2),
                         (rnn.MATH_FP32, 33, 3, 3)]:
    c = rnn.CuDNNRecurrentComponent("cuda:0", math=math)
    c.InitFromString("learning-rate=0.01 num-layers=1 input-dim=24 output-dim=%d rnn-mode=%d bidirectional=true "
                     "max-seq-length=20 mini-batch=%d" % (H, mode, B))
    x = torch.randn(12 * B, 24, device="cuda")
    y = c.Propagate(x)
    dx = c.Backprop(x, y, torch.randn_like(y), to_update=c)
    torch.cuda.synchronize()
    print("rnn math", math, "H", H, float(y.abs().sum()), float(dx.abs().sum()))

spec = synth.ModelSpec(mode=2, layers=2, D=10, H=64, A=12, learning_rate=0.01, param_stddev=0.2)
blobs, aw, ab = synth.model_weights(spec, 3)
x, fl, L, T = synth.features(4, spec.D, 20, 24, 2, 5, spec.A, seed=5)
up = nnet.NnetCtcUpdater(spec, blobs, aw, ab, 4, int(T.max()), math=rnn.MATH_TENSOR)
print("step objf", up.ComputeForMinibatch(torch.from_numpy(x).pin_memory(), int(T.max()), fl, L, T))
print("SANITY DONE")
