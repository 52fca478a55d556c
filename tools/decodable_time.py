"""Time b200ctc_decodable on device-resident network output and report it against
its algorithmic HBM bytes: 2 reads of the valid rows + 1 write of the kept rows."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from kaldi_ctc_b200 import decodable  # noqa: E402

peaks = json.load(open("MEASURED_PEAKS.json")) if __import__("os").path.exists("MEASURED_PEAKS.json") else {}
for (T, B, A, thr) in [(2000, 1, 46, 0.98), (2000, 16, 46, 0.98), (2000, 16, 8000, 1.0), (2000, 16, 8000, 0.98)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(T * B, A, device="cuda", generator=g) * 2
    x[:, 0] += torch.where(torch.rand(T * B, device="cuda", generator=g) < 0.6, 12.0, 0.0)
    il = np.full(B, T, dtype=np.int32)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        out, kept = decodable.decodable_log_probs(torch, x, il, B, None, 1.0, thr, 1e-10, workspace=ws)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    n = 20
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(n):
        out, kept = decodable.decodable_log_probs(torch, x, il, B, None, 1.0, thr, 1e-10, workspace=ws)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / n
    nbytes = 4.0 * A * (2 * T * B + int(kept.sum()))
    print(json.dumps({"T": T, "B": B, "A": A, "blank_threshold": thr, "kept_frac": float(kept.sum()) / (T * B),
                      "ms_per_call_incl_host_sync": round(ms, 4), "GB_per_s": round(nbytes / ms / 1e6, 1)}))
