"""Host and device time of the GPU FormatNnetInput (b200ctc_format_input) on the benchmark minibatch."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from kaldi_ctc_b200 import egs, synth  # noqa: E402

B, D = 16, 40
ex = synth.examples(B, D, 1200, 2000, 120, 180, 48, seed=1002)
st = egs.InputStager()
out = torch.empty(2000 * B, D, device="cuda")
for _ in range(3):
    egs.FormatNnetInput(0, 0, ex, input_mat=out, stager=st)
torch.cuda.synchronize()
n = 20
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(n):
    egs.FormatNnetInput(0, 0, ex, input_mat=out, stager=st)
e1.record()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
frames = sum(e.NumFrames() for e in ex)
print(json.dumps({"host_enqueue_ms_per_call": (t1 - t0) * 1e3 / n, "device_ms_per_call": e0.elapsed_time(e1) / n,
                  "wall_ms_per_call": (t2 - t0) * 1e3 / n, "h2d_bytes": st.h2d_bytes,
                  "slab_bytes": 2000 * B * D * 4, "valid_frames": frames}))
