"""nnet2 binary model I/O and <FilterParams> <-> PyTorch conversion (host code): byte layout known
answers, round trips, and the layout cross-checked against the oracle's own blob locator and against
torch.nn modules running the converted weights."""
import io
import struct

import numpy as np
import pytest

from kaldi_ctc_b200 import model_io, synth
from oracle import pyoracle


def test_component_byte_layout_known_answer():
    c = {"type": "AffineComponent", "learning_rate": 0.5, "linear_params": np.array([[1.0, 2.0]], dtype=np.float32),
         "bias_params": np.array([3.0], dtype=np.float32)}
    buf = io.BytesIO()
    model_io.write_nnet(buf, [c])
    want = (b"\x00B<Nnet> <NumComponents> \x04\x01\x00\x00\x00<Components> <AffineComponent> <LearningRate> \x04" +
            struct.pack("<f", 0.5) + b"<LinearParams> FM \x04\x01\x00\x00\x00\x04\x02\x00\x00\x00" +
            struct.pack("<ff", 1.0, 2.0) + b"<BiasParams> FV \x04\x01\x00\x00\x00" + struct.pack("<f", 3.0) +
            b"<IsGradient> F</AffineComponent> </Components> </Nnet> ")
    assert buf.getvalue() == want


def test_nnet_round_trip():
    spec = synth.ModelSpec(mode=2, layers=2, D=10, H=16, A=12)
    blobs, aw, ab = synth.model_weights(spec, 1)
    comps = model_io.components_of(spec, blobs, aw, ab, softmax=True)
    buf = io.BytesIO()
    model_io.write_nnet(buf, comps)
    back = model_io.read_nnet(io.BytesIO(buf.getvalue()))
    assert [c["type"] for c in back] == ["CuDNNRecurrentComponent", "ClipGradientComponent"] * 2 + \
        ["AffineComponent", "SoftmaxComponent"]
    assert np.array_equal(back[0]["filter_params"], blobs[0]) and back[0]["bidirectional"] is True
    assert back[0]["rnn_mode"] == 2 and back[2]["input_dim"] == 32 and back[1]["clipping_threshold"] == spec.clipping_threshold
    assert np.array_equal(back[4]["linear_params"], aw) and np.array_equal(back[4]["bias_params"], ab)
    buf2 = io.BytesIO()
    model_io.write_nnet(buf2, back)
    assert buf2.getvalue() == buf.getvalue()


@pytest.mark.parametrize("mode,bidir,layers", [(2, True, 1), (3, True, 2), (1, False, 2), (2, False, 1)])
def test_blob_layout_matches_oracle_locator(mode, bidir, layers):
    D, H = 6, 5
    n = pyoracle.rnn_param_count(mode, bidir, layers, D, H)
    blob = np.arange(n, dtype=np.float32)
    sd = model_io.filter_params_to_torch(blob, mode, bidir, layers, D, H)
    G, dirs = {0: 1, 1: 1, 2: 4, 3: 3}[mode], 2 if bidir else 1
    for pl in range(layers * dirs):
        sfx = "_l%d%s" % (pl // dirs, "_reverse" if pl % dirs else "")
        for g in range(G):
            loc = lambda lin, is_bias: pyoracle.rnn_locate(mode, bidir, layers, D, H, pl, lin, is_bias)[0]
            assert sd["weight_ih" + sfx][g * H, 0] == loc(g, 0) and sd["weight_hh" + sfx][g * H, 0] == loc(G + g, 0)
            assert sd["bias_ih" + sfx][g * H] == loc(g, 1) and sd["bias_hh" + sfx][g * H] == loc(G + g, 1)
    assert np.array_equal(model_io.torch_to_filter_params(sd, mode, bidir, layers, D, H), blob)


@pytest.mark.parametrize("mode", [2, 3, 1])
def test_converted_weights_drive_torch_module_like_the_oracle(mode):
    import torch
    D, H, T, B, layers = 7, 6, 9, 3, 2
    rng = np.random.default_rng(mode)
    n = pyoracle.rnn_param_count(mode, True, layers, D, H)
    blob = (rng.standard_normal(n) * 0.3).astype(np.float32)
    x = rng.standard_normal((T * B, D)).astype(np.float32)
    y_ref = pyoracle.rnn(mode, True, layers, H, x, blob, B, dtype=np.float64)
    cls = {1: lambda: torch.nn.RNN(D, H, layers, nonlinearity="tanh", bidirectional=True),
           2: lambda: torch.nn.LSTM(D, H, layers, bidirectional=True),
           3: lambda: torch.nn.GRU(D, H, layers, bidirectional=True)}[mode]
    mod = cls().double()
    sd = model_io.filter_params_to_torch(blob, mode, True, layers, D, H)
    mod.load_state_dict({k: torch.from_numpy(v).double() for k, v in sd.items()})
    y, _ = mod(torch.from_numpy(x).double().reshape(T, B, D))
    np.testing.assert_allclose(y.detach().numpy().reshape(T * B, 2 * H), y_ref, atol=1e-6)
