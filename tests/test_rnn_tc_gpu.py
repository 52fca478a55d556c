"""Tensor-core mode (B200RNN_MATH_TENSOR) of the recurrent path.

Stated tolerance of this mode (north_star: "within a stated bf16/TF32 tolerance"):
recurrent operands (R, h_{t-1}) are rounded to BF16 (8 mantissa bits), projections
and weight gradients run in TF32 (10 bits), accumulation / cell state / outputs are
fp32, gate non-linearities use tanh.approx (abs err 5e-4).  Against the fp64 oracle:
    outputs y            max-abs <= 5e-3   (values in [-1, 1]; measured 3e-4 .. 2e-3)
    dx, dw               max-abs <= 1e-2 * max|ref|   (measured 6e-4 .. 1.6e-3)
and against a numpy emulation that applies the SAME operand roundings (so only
accumulation order and tanh.approx differ): y max-abs <= 1e-3 (measured <= 3.6e-4).
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bf16(x):
    """round-to-nearest-even to bfloat16, returned as float32/64 values"""
    x = np.asarray(x, np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).astype(np.float64)


def _tf32(x):
    """truncate to 10 mantissa bits (what kind::tf32 reads of an fp32 word)"""
    u = np.asarray(x, np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return u.view(np.float32).astype(np.float64)


def _emulate_forward(mode, D, H, B, x, w, T):
    """bidirectional single layer, same operand roundings as the tensor-core kernels"""
    from oracle import pyoracle
    G = {2: 4, 3: 3}[mode]
    y = np.zeros((T * B, 2 * H))
    sig = lambda v: 1.0 / (1.0 + np.exp(-v))
    for d in range(2):
        o, _, _ = pyoracle.rnn_locate(mode, True, 1, D, H, d, 0, False)
        Wi = w[o:o + G * H * D].reshape(G * H, D)
        o, _, _ = pyoracle.rnn_locate(mode, True, 1, D, H, d, G, False)
        R = w[o:o + G * H * H].reshape(G * H, H)
        o, _, _ = pyoracle.rnn_locate(mode, True, 1, D, H, d, 0, True)
        bW, bR = w[o:o + G * H].astype(np.float64), w[o + G * H:o + 2 * G * H].astype(np.float64)
        pre = _tf32(x) @ _tf32(Wi).T + bW
        if mode == 2:
            pre += bR
        else:
            pre[:, :2 * H] += bR[:2 * H]
        Rb = _bf16(R)
        h = np.zeros((B, H))
        c = np.zeros((B, H))
        for step in range(T):
            t = step if d == 0 else T - 1 - step
            rec = _bf16(h) @ Rb.T
            p = pre[t * B:(t + 1) * B]
            if mode == 2:
                i, f = sig(p[:, :H] + rec[:, :H]), sig(p[:, H:2 * H] + rec[:, H:2 * H])
                g, o_ = np.tanh(p[:, 2 * H:3 * H] + rec[:, 2 * H:3 * H]), sig(p[:, 3 * H:] + rec[:, 3 * H:])
                c = f * c + i * g
                h = o_ * np.tanh(c)
            else:
                r, z = sig(p[:, :H] + rec[:, :H]), sig(p[:, H:2 * H] + rec[:, H:2 * H])
                n = np.tanh(p[:, 2 * H:] + r * (rec[:, 2 * H:] + bR[2 * H:]))
                h = (1 - z) * n + z * h
            y[t * B:(t + 1) * B, d * H:(d + 1) * H] = h
    return y


def _run_tc(mode, D, H, B, Tn, x, w, dy):
    import torch
    from kaldi_ctc_b200 import rnn
    c = rnn.CuDNNRecurrentComponent("cuda:0", math=rnn.MATH_TENSOR)
    c.InitFromString("learning-rate=0.01 num-layers=1 input-dim=%d output-dim=%d rnn-mode=%d "
                     "bidirectional=true max-seq-length=%d mini-batch=%d" % (D, H, mode, Tn, B))
    c.SetParams(w)
    xt, dyt = torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda()
    y = c.Propagate(xt)

    class Grab:
        def Update(self, g, clip):
            self.g = g.clone()
    grab = Grab()
    dx = c.Backprop(xt, y, dyt, to_update=grab)
    torch.cuda.synchronize()
    return y.cpu().numpy(), dx.cpu().numpy(), grab.g.cpu().numpy()


@pytest.mark.parametrize("mode,D,H,B,Tn", [(2, 40, 320, 16, 40), (2, 640, 320, 16, 12), (3, 40, 320, 64, 25),
                                          (2, 24, 64, 5, 30), (3, 24, 128, 17, 20), (2, 40, 320, 3, 200),
                                          (2, 40, 320, 32, 20),   # batch chunk 8 (two utterance slots per thread)
                                          (3, 24, 64, 4, 3), (2, 24, 64, 4, 1)])   # T below the prefetch depth
def test_tensor_mode_within_stated_tolerance(mode, D, H, B, Tn):
    from oracle import pyoracle
    rng = np.random.default_rng(mode * 100 + B + H)
    n = pyoracle.rnn_param_count(mode, True, 1, D, H)
    w = (rng.standard_normal(n) * 0.05).astype(np.float32)
    x = rng.standard_normal((Tn * B, D)).astype(np.float32)
    dy = rng.standard_normal((Tn * B, 2 * H)).astype(np.float32)
    yr, dxr, dwr = pyoracle.rnn(mode, True, 1, H, x, w, B, dy=dy, dtype=np.float64)
    y, dx, dw = _run_tc(mode, D, H, B, Tn, x, w, dy)
    ye = _emulate_forward(mode, D, H, B, x, w, Tn)
    e_emul, e_y = np.abs(y - ye).max(), np.abs(y - yr).max()
    e_dx, e_dw = np.abs(dx - dxr).max() / np.abs(dxr).max(), np.abs(dw - dwr).max() / np.abs(dwr).max()
    print("tensor-mode errors: y vs emulation %.2e, y vs fp64 %.2e, dx rel %.2e, dw rel %.2e" % (e_emul, e_y, e_dx, e_dw))
    assert e_emul < 1e-3
    assert e_y < 5e-3
    assert e_dx < 1e-2 and e_dw < 1e-2


@pytest.mark.parametrize("mode,D,H,B,Tn,force", [(2, 640, 320, 16, 12, 1), (3, 640, 320, 16, 9, 1), (2, 136, 64, 7, 10, 1),
                                                (2, 640, 320, 16, 130, -1)])
def test_input_gradient_through_the_cta_pair_gemm(mode, D, H, B, Tn, force):
    """dx = dG_fwd . Wi_fwd + dG_bwd . Wi_bwd in ONE launch of the CTA-pair GEMM (second operand pair, tiles of 256 /
    256 / 128 columns at D = 640): forced at small sizes (GEMM_PAIR = 1), and chosen by the library itself at
    T * B >= 2048 rows (the last case)."""
    from kaldi_ctc_b200 import rnn
    from oracle import pyoracle
    rng = np.random.default_rng(mode * 10 + D + Tn)
    n = pyoracle.rnn_param_count(mode, True, 1, D, H)
    w = (rng.standard_normal(n) * 0.05).astype(np.float32)
    x = rng.standard_normal((Tn * B, D)).astype(np.float32)
    dy = rng.standard_normal((Tn * B, 2 * H)).astype(np.float32)
    yr, dxr, dwr = pyoracle.rnn(mode, True, 1, H, x, w, B, dy=dy, dtype=np.float64)
    rnn.set_tuning("GEMM_PAIR", force)
    try:
        y, dx, dw = _run_tc(mode, D, H, B, Tn, x, w, dy)
    finally:
        rnn.set_tuning("GEMM_PAIR", -1)
    assert np.abs(y - yr).max() < 5e-3
    assert np.abs(dx - dxr).max() / np.abs(dxr).max() < 1e-2
    assert np.abs(dw - dwr).max() / np.abs(dwr).max() < 1e-2


@pytest.mark.parametrize("mode,H,B,bc", [(2, 64, 23, 16), (3, 128, 23, 16), (2, 320, 50, 16), (3, 64, 13, 8),
                                         (1, 64, 21, 16)])
def test_split_epilogue_chunks(monkeypatch, mode, H, B, bc):
    """Batch chunks of 8 / 16 run a second set of epilogue warps (each takes half of the chunk's utterance
    columns); ragged last chunks (23 = 16 + 7, 50 = 3 x 16 + 2, 13 = 8 + 5), all recurrent modes, and the
    same results as four epilogue warps (B200RNN_SPLIT_EPILOGUE=0)."""
    from oracle import pyoracle
    D, Tn = 24, 11
    monkeypatch.setenv("B200RNN_TC_BC", str(bc))
    rng = np.random.default_rng(mode * 1000 + B + H)
    n = pyoracle.rnn_param_count(mode, True, 1, D, H)
    w = (rng.standard_normal(n) * 0.05).astype(np.float32)
    x = rng.standard_normal((Tn * B, D)).astype(np.float32)
    dy = rng.standard_normal((Tn * B, 2 * H)).astype(np.float32)
    yr, dxr, dwr = pyoracle.rnn(mode, True, 1, H, x, w, B, dy=dy, dtype=np.float64)
    y, dx, dw = _run_tc(mode, D, H, B, Tn, x, w, dy)
    assert np.abs(y - yr).max() < 5e-3
    assert np.abs(dx - dxr).max() / np.abs(dxr).max() < 1e-2
    assert np.abs(dw - dwr).max() / np.abs(dwr).max() < 1e-2
