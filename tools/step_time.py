"""Times one full training step (NnetCtcUpdater mirror) for an arbitrary topology.
Usage: python tools/step_time.py mode layers D H A B Tlo Thi Llo Lhi [math]"""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kaldi_ctc_b200 import nnet, rnn, synth  # noqa: E402

mode, layers, D, H, A, B, Tlo, Thi, Llo, Lhi = [int(v) for v in sys.argv[1:11]]
math = rnn.MATH_FP32 if (len(sys.argv) > 11 and sys.argv[11] == "fp32") else rnn.MATH_TENSOR
spec = synth.ModelSpec(mode=mode, layers=layers, D=D, H=H, A=A)
blobs, aw, ab = synth.model_weights(spec, 7)
x, fl, L, T = synth.features(B, D, Tlo, Thi, Llo, Lhi, A, seed=1002)
Tmax = int(T.max())
up = nnet.NnetCtcUpdater(spec, blobs, aw, ab, B, Tmax, math=math)
up.FormatInput(torch.from_numpy(x).pin_memory(), Tmax)
for _ in range(3):
    up.ComputeForMinibatch(None, Tmax, fl, L, T, host_sync=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
e0.record()
for _ in range(n):
    up.ComputeForMinibatch(None, Tmax, fl, L, T, host_sync=False)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(json.dumps({"mode": mode, "layers": layers, "D": D, "H": H, "A": A, "B": B, "Tmax": Tmax,
                  "valid_frames": int(T.sum()), "ms_per_step": ms, "frames_per_s": float(T.sum()) / ms * 1e3,
                  "objf": up.last_objf()}))
