"""Times one b200ctc_loss call on BASELINE configs[4] (A=4000, T_b~U{1500..3000}, L_b~U{50..600}) at a given
batch size, activations drawn on the device; prints ms and algorithmic GB/s.  B200CTC_LIB=<path> times another build
of the library (A/B runs of kernel variants).  Usage: python tools/ctc_stress_time.py B [iters]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kaldi_ctc_b200 import _lib, ctc, synth  # noqa: E402

if os.environ.get("B200CTC_LIB"):   # tuning builds (e.g. tools/build/libb200ctc_prof.so with phase counters)
    import ctypes
    _lib._cache["libb200ctc.so"] = ctypes.CDLL(os.path.abspath(os.environ["B200CTC_LIB"]))

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
A = 4000
rng = np.random.Generator(np.random.PCG64(1005))
T, L = synth._lengths(rng, B, 1500, 3000, 50, 600)
labels = np.concatenate([rng.integers(1, A, size=int(l)) for l in L]).astype(np.int32)
Tmax = int(T.max())
g0 = torch.Generator(device="cuda")
g0.manual_seed(1005)
a = torch.randn(Tmax, B, A, device="cuda", generator=g0) * 2.0
a *= (torch.arange(Tmax, device="cuda")[:, None] < torch.from_numpy(T.astype(np.int64)).cuda()[None, :])[:, :, None]
op = ctc.CtcLoss("cuda:0")
g = torch.empty_like(a)
cd = torch.zeros(B, device="cuda")
flag = torch.zeros(1, dtype=torch.int32, device="cuda")
run = lambda: op.compute_extended(a, labels, L, T, gradients=g, costs_dev=cd, no_sync=True, nonfinite_dev=flag)
for _ in range(2):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
nbytes = ctc.algorithmic_bytes(L, T, A)
print(json.dumps({"B": B, "ms": ms, "min_ms": float(min(ts)), "alg_GBps": nbytes / ms / 1e6,
                  "frac_of_6528": nbytes / ms / 1e6 / 6528.4, "costs_sum": float(cd.sum()), "flag": int(flag.item()),
                  "env": {k: v for k, v in os.environ.items() if k.startswith("B200CTC")}}))
