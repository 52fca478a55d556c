"""Times one recurrent layer (forward / backward-data / backward-weights) with CUDA events.
Usage: python tools/rnn_time.py [math=tensor|fp32] [T] [B] [D] [H] [mode]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kaldi_ctc_b200 import rnn, _lib  # noqa: E402

if os.environ.get("B200RNN_LIB"):   # tuning builds
    import ctypes
    _lib._cache["libb200rnn.so"] = ctypes.CDLL(os.path.abspath(os.environ["B200RNN_LIB"]))

math = sys.argv[1] if len(sys.argv) > 1 else "tensor"
T, B, D, H, mode = [int(v) for v in (sys.argv[2:7] + ["2000", "16", "640", "320", "2"][len(sys.argv) - 2:])]
c = rnn.CuDNNRecurrentComponent("cuda:0", math=rnn.MATH_TENSOR if math == "tensor" else rnn.MATH_FP32)
c.InitFromString("learning-rate=0.001 num-layers=1 input-dim=%d output-dim=%d rnn-mode=%d bidirectional=true "
                 "max-seq-length=%d mini-batch=%d" % (D, H, mode, T, B))
x = torch.randn(T * B, D, device="cuda")
dy = torch.randn(T * B, 2 * H, device="cuda")
y = torch.empty(T * B, 2 * H, device="cuda")
c.plan.set_profiling(True)


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    for k in range(3):
        c.plan.get_profile(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    prof = [c.plan.get_profile(k) for k in range(3)]
    return e0.elapsed_time(e1) / n, [p[0] / n for p in prof]


class Grab:
    def Update(self, g, clip):
        pass


fwd, pf = timed(lambda: c.Propagate(x, y))
bwd, pb = timed(lambda: c.Backprop(x, y, dy, to_update=Grab()))
print(json.dumps({"math": math, "T": T, "B": B, "D": D, "H": H, "mode": mode, "fwd_ms": fwd, "bwd_ms": bwd,
                  "fwd_rec_ms": pf[0], "fwd_gemm_ms": pf[2], "bwd_rec_ms": pb[1], "bwd_gemm_ms": pb[2],
                  "us_per_step_fwd_rec": pf[0] * 1e3 / T, "us_per_step_bwd_rec": pb[1] * 1e3 / T,
                  "tc_bc": os.environ.get("B200RNN_TC_BC", "auto")}))
