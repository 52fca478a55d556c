"""Pins oracle/rnn_oracle.c (CPU restatement of the cuDNN-5 recurrent
forward/backward behind CuDNNRecurrentComponent, call sites
src/nnet2/nnet-cudnn-component.cc:534-555,576-599) against committed
torch.nn.{LSTM,GRU,RNN} fp64 goldens and finite differences."""
import os

import numpy as np
import pytest

from oracle import pyoracle

CASES = ["rnn_lstm_bi", "rnn_lstm_uni2", "rnn_gru_bi", "rnn_gru_bi2", "rnn_relu_bi", "rnn_tanh_uni", "rnn_gru_odd", "rnn_lstm_odd"]


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    return {k: z[k] for k in z.files}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_torch_fp64_golden(golden_dir, name):
    g = _load(golden_dir, name)
    mode, bidir, layers, H, B = int(g["mode"]), bool(g["bidir"]), int(g["layers"]), int(g["H"]), int(g["B"])
    y, dx, dw = pyoracle.rnn(mode, bidir, layers, H, g["x"], g["w"], B, dy=g["dy"], dtype=np.float64)
    assert np.abs(y - g["y"]).max() < 1e-12
    assert np.abs(dx - g["dx"]).max() < 1e-11
    assert np.abs(dw - g["dw"]).max() < 1e-10
    y32, dx32, dw32 = pyoracle.rnn(mode, bidir, layers, H, g["x"], g["w"], B, dy=g["dy"], dtype=np.float32)
    assert np.abs(y32 - g["y"]).max() < 1e-5       # north_star fp32-mode tolerance
    assert np.abs(dx32 - g["dx"]).max() < 1e-4
    assert np.abs(dw32 - g["dw"]).max() < 1e-4


def test_blob_layout_matches_reference_sizes():
    # layer-1 BLSTM 926 720 floats, layers 2-5 2 462 720 each
    assert pyoracle.rnn_param_count(2, True, 1, 40, 320) == 926720
    assert pyoracle.rnn_param_count(2, True, 1, 640, 320) == 2462720
    # matrices of every pseudo-layer first, then biases (cudnnGetRNNLinLayer*Params order)
    off, r, c = pyoracle.rnn_locate(2, True, 1, 40, 320, 1, 0, False)
    assert (off, r, c) == (4 * 320 * 40 + 4 * 320 * 320, 320, 40)
    off, r, c = pyoracle.rnn_locate(2, True, 1, 40, 320, 0, 4, False)
    assert (off, r, c) == (4 * 320 * 40, 320, 320)
    off, r, c = pyoracle.rnn_locate(2, True, 1, 40, 320, 0, 0, True)
    assert off == 2 * (4 * 320 * 40 + 4 * 320 * 320) and (r, c) == (320, 1)
    off, _, _ = pyoracle.rnn_locate(2, True, 1, 40, 320, 1, 7, True)
    assert off == 926720 - 320


@pytest.mark.parametrize("mode", [2, 3])
def test_finite_difference_like_nnet_component_test(mode):
    """The reference's generic component check (src/nnet2/nnet-component-test.cc:78-208):
    random linear objective, perturb input / parameters, compare predicted and
    observed objective change."""
    rng = np.random.default_rng(7)
    D, H, B, T = 5, 6, 2, 4
    n = pyoracle.rnn_param_count(mode, True, 1, D, H)
    w = (rng.standard_normal(n) * 0.4).astype(np.float32)
    x = rng.standard_normal((T * B, D)).astype(np.float32)
    dy = rng.standard_normal((T * B, 2 * H))
    y, dx, dw = pyoracle.rnn(mode, True, 1, H, x, w, B, dy=dy, dtype=np.float64)
    f0 = (y * dy).sum()
    for _ in range(5):
        px = (rng.standard_normal(x.shape) * 1e-4).astype(np.float32)
        pw = (rng.standard_normal(w.shape) * 1e-4).astype(np.float32)
        f1 = (pyoracle.rnn(mode, True, 1, H, x + px, w, B, dtype=np.float64) * dy).sum()
        pred = (dx * ((x + px).astype(np.float64) - x)).sum()
        assert abs((f1 - f0) - pred) < 5e-2 * abs(pred) + 1e-6
        f2 = (pyoracle.rnn(mode, True, 1, H, x, w + pw, B, dtype=np.float64) * dy).sum()
        pred = (dw * ((w + pw).astype(np.float64) - w)).sum()
        assert abs((f2 - f0) - pred) < 5e-2 * abs(pred) + 1e-6
