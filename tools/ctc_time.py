"""Times the CTC call (b200ctc_loss, no host sync) on the BASELINE configs with CUDA
events; prints algorithmic GB/s.  Usage: python tools/ctc_time.py [cfg...]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kaldi_ctc_b200 import ctc, synth  # noqa: E402


def time_cfg(cfg, scale=1.0, iters=10):
    bt = synth.config_ctc(cfg, scale=scale)
    op = ctc.CtcLoss("cuda:0")
    a = torch.from_numpy(bt.activations).cuda()
    g = torch.empty_like(a)
    cd = torch.zeros(a.shape[1], device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    run = lambda: op.compute_extended(a, bt.flat_labels, bt.label_lengths, bt.input_lengths,
                                      gradients=g, costs_dev=cd, no_sync=True)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    nbytes = ctc.algorithmic_bytes(bt.label_lengths, bt.input_lengths, a.shape[2])
    return {"cfg": cfg, "shape": list(a.shape), "ms": ms, "min_ms": float(min(ts)),
            "alg_GB": nbytes / 1e9, "alg_GBps": nbytes / ms / 1e6,
            "frames_per_s": float(bt.input_lengths.sum()) / ms * 1e3}


if __name__ == "__main__":
    cfgs = [c for c in sys.argv[1:]] or ["1", "4", "5:0.125"]
    for c in cfgs:
        cfg, _, sc = c.partition(":")
        print(json.dumps(time_cfg(int(cfg), float(sc) if sc else 1.0)))
