"""Parity of the CUDA recurrent path (libb200rnn.so, through the C ABI of
include/b200rnn.h and the CuDNNRecurrentComponent mirror) with the oracle.
fp32 mode tolerance is north_star's 1e-5 on outputs (gradients: 1e-4 max-abs,
scaled by their magnitude where sums over T*B rows make them large)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLDEN_OK = ["rnn_lstm_bi", "rnn_lstm_uni2", "rnn_gru_bi", "rnn_gru_bi2", "rnn_relu_bi", "rnn_tanh_uni",
             "rnn_gru_odd", "rnn_lstm_odd"]   # the odd hidden sizes take the general streaming kernels


@pytest.fixture(scope="module")
def T():
    import torch
    assert torch.cuda.is_available()
    return torch


def _component(rnn, mode, bidir, layers, D, H, B, Tmax, w, math=0):
    c = rnn.CuDNNRecurrentComponent("cuda:0", math=math)
    c.InitFromString("learning-rate=0.01 num-layers=%d input-dim=%d output-dim=%d rnn-mode=%d "
                     "bidirectional=%s max-seq-length=%d mini-batch=%d" %
                     (layers, D, H, mode, "true" if bidir else "false", Tmax, B))
    c.SetParams(w)
    return c


def _run(torch, rnn, mode, bidir, layers, D, H, B, x, w, dy, math=0):
    Tn = x.shape[0] // B
    c = _component(rnn, mode, bidir, layers, D, H, B, Tn, w, math)
    xt, dyt = torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda()
    y = c.Propagate(xt)

    class Grab:
        def Update(self, g, clip):
            self.g = g.clone()
    grab = Grab()
    dx = c.Backprop(xt, y, dyt, to_update=grab)
    torch.cuda.synchronize()
    return y.cpu().numpy(), dx.cpu().numpy(), grab.g.cpu().numpy()


def _assert_close(got, ref, atol, name):
    scale = max(1.0, float(np.abs(ref).max()))
    err = float(np.abs(got - ref).max())
    assert err < atol * scale, "%s: max-abs err %g (scale %g)" % (name, err, scale)


@pytest.mark.parametrize("name", GOLDEN_OK)
def test_matches_committed_torch_fp64_golden(T, golden_dir, name):
    from kaldi_ctc_b200 import rnn
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    mode, bidir, layers, D, H, B = (int(z[k]) for k in ("mode", "bidir", "layers", "D", "H", "B"))
    y, dx, dw = _run(T, rnn, mode, bool(bidir), layers, D, H, B, z["x"], z["w"], z["dy"])
    _assert_close(y, z["y"], 1e-5, "y")
    _assert_close(dx, z["dx"], 1e-4, "dx")
    _assert_close(dw, z["dw"], 1e-4, "dw")


@pytest.mark.parametrize("mode,D,H,B,Tn", [(2, 40, 320, 16, 12), (2, 640, 320, 16, 6), (3, 40, 320, 64, 7),
                                          (2, 24, 64, 5, 9), (3, 24, 64, 17, 5), (1, 16, 128, 3, 8)])
def test_benchmark_shapes_against_fp64_oracle(T, mode, D, H, B, Tn):
    """The benchmark layer shapes (BLSTM-320 layer 1 / layers 2-5, BiGRU-320 at B=64:
    several batch chunks) and ragged minibatch sizes."""
    from kaldi_ctc_b200 import rnn
    from oracle import pyoracle
    rng = np.random.default_rng(mode * 100 + B)
    n = pyoracle.rnn_param_count(mode, True, 1, D, H)
    w = (rng.standard_normal(n) * 0.05).astype(np.float32)
    x = rng.standard_normal((Tn * B, D)).astype(np.float32)
    dy = rng.standard_normal((Tn * B, 2 * H)).astype(np.float32)
    yr, dxr, dwr = pyoracle.rnn(mode, True, 1, H, x, w, B, dy=dy, dtype=np.float64)
    y, dx, dw = _run(T, rnn, mode, True, 1, D, H, B, x, w, dy)
    _assert_close(y, yr, 1e-5, "y")
    _assert_close(dx, dxr, 1e-4, "dx")
    _assert_close(dw, dwr, 1e-4, "dw")


def test_blob_layout_and_sizes(T):
    from kaldi_ctc_b200 import rnn
    from oracle import pyoracle
    p = rnn.Plan(2, True, 1, 40, 320, 16, 100)
    assert p.param_count == 926720
    for pl in range(2):
        for lin in range(8):
            for bias in (False, True):
                assert p.locate(pl, lin, bias) == pyoracle.rnn_locate(2, True, 1, 40, 320, pl, lin, bias)
    p2 = rnn.Plan(3, True, 2, 30, 16, 4, 10)
    assert p2.param_count == pyoracle.rnn_param_count(3, True, 2, 30, 16)
    assert p2.locate(3, 5, True) == pyoracle.rnn_locate(3, True, 2, 30, 16, 3, 5, True)
    with pytest.raises(rnn.RnnError):
        rnn.Plan(7, True, 1, 4, 4, 1, 1)


def test_inference_path_minibatch_one(T):
    """mini_batch == 1 takes the ForwardInference branch (nnet-cudnn-component.cc:534-543): no reserve."""
    from kaldi_ctc_b200 import rnn
    from oracle import pyoracle
    rng = np.random.default_rng(1)
    D, H, Tn = 20, 32, 30
    w = (rng.standard_normal(pyoracle.rnn_param_count(2, True, 1, D, H)) * 0.1).astype(np.float32)
    x = rng.standard_normal((Tn, D)).astype(np.float32)
    c = _component(rnn, 2, True, 1, D, H, 1, Tn, w)
    y = c.Propagate(T.from_numpy(x).cuda()).cpu().numpy()
    _assert_close(y, pyoracle.rnn(2, True, 1, H, x, w, 1, dtype=np.float64), 1e-5, "y")


def test_update_clips_then_applies(T):
    """Backprop's tail (:602-614): w += lr * clamp(dW, +-clip)."""
    from kaldi_ctc_b200 import rnn
    from oracle import pyoracle
    rng = np.random.default_rng(3)
    D, H, B, Tn = 8, 16, 4, 20
    w = (rng.standard_normal(pyoracle.rnn_param_count(2, True, 1, D, H)) * 0.3).astype(np.float32)
    x = rng.standard_normal((Tn * B, D)).astype(np.float32) * 3
    dy = rng.standard_normal((Tn * B, 2 * H)).astype(np.float32) * 3
    _, _, dwr = pyoracle.rnn(2, True, 1, H, x, w, B, dy=dy, dtype=np.float64)
    c = _component(rnn, 2, True, 1, D, H, B, Tn, w)
    c.clip_gradient_ = 0.5
    assert np.abs(dwr).max() > 0.5    # the clamp is exercised
    xt = T.from_numpy(x).cuda()
    y = c.Propagate(xt)
    c.Backprop(xt, y, T.from_numpy(dy).cuda(), to_update=c)
    want = w + 0.01 * np.clip(dwr, -0.5, 0.5)
    _assert_close(c.Vectorize(), want, 1e-5, "updated blob")


@pytest.mark.parametrize("tA,tB,M,N,K", [(0, 1, 300, 130, 77), (0, 0, 129, 65, 200), (1, 0, 70, 33, 1000),
                                        (1, 1, 17, 19, 23), (1, 0, 48, 640, 9000)])
def test_gemm_against_numpy(T, tA, tB, M, N, K):
    from kaldi_ctc_b200 import rnn
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
    Bm = rng.standard_normal((N, K) if tB else (K, N)).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    want = 0.5 * ((A.T if tA else A).astype(np.float64) @ (Bm.T if tB else Bm)) + 2.0 * C0 + bias
    At, Bt, Ct, bt = (T.from_numpy(v).cuda() for v in (A, Bm, C0.copy(), bias))
    ws = T.empty(64 << 20, dtype=T.uint8, device="cuda")
    rnn.gemm(T, tA, tB, M, N, K, 0.5, At, A.shape[1], Bt, Bm.shape[1], 2.0, Ct, N, bias=bt, workspace=ws)
    T.cuda.synchronize()
    _assert_close(Ct.cpu().numpy(), want, 1e-5, "gemm")


def test_clip_row_norm_and_column_sums(T):
    from kaldi_ctc_b200 import rnn
    rng = np.random.default_rng(0)
    d = (rng.standard_normal((1000, 640)) * np.linspace(0.1, 3, 1000)[:, None]).astype(np.float32)
    dt = T.from_numpy(d.copy()).cuda()
    rnn.clip_row_norm(T, dt, 30.0)
    nrm = np.linalg.norm(d, axis=1, keepdims=True)
    want = d * np.minimum(1.0, 30.0 / nrm)
    assert (nrm > 30).any() and (nrm < 30).any()
    _assert_close(dt.cpu().numpy(), want, 1e-5, "clip")
    out = T.ones(640, device="cuda")
    ws = T.empty(1 << 20, dtype=T.uint8, device="cuda")
    rnn.column_sums(T, dt, out, True, ws)
    _assert_close(out.cpu().numpy(), 1.0 + want.astype(np.float64).sum(0), 1e-5, "colsum")


@pytest.mark.parametrize("mode,D,H,B,Tn", [(2, 20, 64, 5, 9), (3, 12, 48, 19, 6), (0, 8, 33, 2, 5), (2, 16, 1100, 3, 4)])
def test_general_streaming_path(T, monkeypatch, mode, D, H, B, Tn):
    """Shapes the persistent kernels cannot hold on chip (any H, e.g. 33 or 1100) and, forced through
    B200RNN_FORCE_STREAM, ordinary ones: per-time-step launches, weights streamed from L2."""
    from kaldi_ctc_b200 import rnn
    from oracle import pyoracle
    monkeypatch.setenv("B200RNN_FORCE_STREAM", "1")
    rng = np.random.default_rng(H + B)
    n = pyoracle.rnn_param_count(mode, True, 1, D, H)
    w = (rng.standard_normal(n) * (0.3 / np.sqrt(H))).astype(np.float32)
    x = rng.standard_normal((Tn * B, D)).astype(np.float32)
    dy = rng.standard_normal((Tn * B, 2 * H)).astype(np.float32)
    yr, dxr, dwr = pyoracle.rnn(mode, True, 1, H, x, w, B, dy=dy, dtype=np.float64)
    y, dx, dw = _run(T, rnn, mode, True, 1, D, H, B, x, w, dy)
    _assert_close(y, yr, 1e-5, "y")
    _assert_close(dx, dxr, 1e-4, "dx")
    _assert_close(dw, dwr, 1e-4, "dw")
    # inference entry point on the same path
    c = _component(rnn, mode, True, 1, D, H, 1, Tn, w)
    y1 = c.Propagate(T.from_numpy(x[:Tn]).cuda()).cpu().numpy()
    _assert_close(y1, pyoracle.rnn(mode, True, 1, H, x[:Tn], w, 1, dtype=np.float64), 1e-5, "y (B=1)")
