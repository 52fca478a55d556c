"""Debug aid: runs one CTC call and compares every utterance with the fp64 oracle: cost, max gradient error and where
it is (frame, symbol).  Usage: python tools/ctc_utt_debug.py <BASELINE config index | ring>"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from kaldi_ctc_b200 import ctc, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "1"
if cfg == "ring":
    rng = np.random.default_rng(4001)
    il = np.array([90, 400, 37, 333, 20, 256, 1], np.int32)
    ll = np.array([0, 300, 12, 150, 30, 100, 1], np.int32)
    T, B, A = int(il.max()), len(il), 4000
    act = (rng.standard_normal((T, B, A)) * 2).astype(np.float32)
    for b in range(B):
        act[il[b]:, b, :] = 0
    fl = np.concatenate([rng.integers(1, A, size=int(l)) for l in ll]).astype(np.int32)
else:
    bt = synth.config_ctc(int(cfg))
    act, fl, ll, il = bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths
c_ref, g_ref = pyoracle.ctc(act, fl, ll, il, dtype=np.float64)
op = ctc.CtcLoss("cuda:0")
costs, grad = op.compute(torch.from_numpy(act).cuda(), fl, ll, il)
g = grad.cpu().numpy()
offs = np.concatenate([[0], np.cumsum(ll)])
for b in range(act.shape[1]):
    Tb = int(il[b])
    e = np.abs(g[:, b] - g_ref[:, b])
    t, k = np.unravel_index(np.argmax(e), e.shape)
    lab = fl[offs[b]:offs[b + 1]]
    print("utt %2d T=%4d L=%3d cost %.4f ref %.4f rel %.1e | grad err %.2e at t=%d k=%d (label? %s blank? %s) got %.5f want %.5f pad_nonzero=%d rowsum_max=%.1e"
          % (b, Tb, ll[b], costs[b], c_ref[b], abs(costs[b] - c_ref[b]) / max(1e-30, abs(c_ref[b])), e.max(), t, k, k in set(lab.tolist()),
             k == 0, g[t, b, k], g_ref[t, b, k], int((g[Tb:, b] != 0).sum()), np.abs(g[:Tb, b].sum(-1)).max()))
