"""b200ctc_decodable (softmax + blank-frame skipping + floor/log/prior/scale in one pass) against the
oracle's restatement of CtcDecodableAmNnet (ctc-decodable-am-nnet.cc:54-86), and the decodable objects
on top of the inference-mode recurrent stack."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ATOL = 2e-5  # fp32 log-probabilities down to log(1e-10) = -23; x |prob_scale|


def _logits(rng, Tmax, B, A, blank_frac):
    x = (rng.standard_normal((Tmax, B, A)) * 2.0).astype(np.float32)
    boost = rng.random((Tmax, B)) < blank_frac
    x[..., 0] += np.where(boost, 9.0 + 3.0 * rng.random((Tmax, B)), 0.0).astype(np.float32)
    return x


def _p_blank(x):
    e = np.exp(x.astype(np.float64) - x.max(-1, keepdims=True))
    return e[..., 0] / e.sum(-1)


@pytest.mark.parametrize("A", [7, 46, 131, 8000])
@pytest.mark.parametrize("B", [1, 5])
@pytest.mark.parametrize("thr,use_priors,scale", [(1.0, False, 1.0), (0.98, True, 0.5), (0.5, True, 1.0),
                                                   (1e-9, False, 2.0)])
def test_kernel_matches_oracle(A, B, thr, use_priors, scale):
    import torch
    from kaldi_ctc_b200 import decodable
    from oracle import pyoracle
    rng = np.random.default_rng(A * 31 + B)
    T = rng.integers(30, 90, size=B).astype(np.int32)
    Tmax = int(T.max())
    x = _logits(rng, Tmax, B, A, 0.6)
    if B > 1:
        x[:, 1, 0] += 30.0     # utterance 1: every frame is blank-dominated -> "keep everything" corner (:62-63)
    if thr < 1.0:   # keep every frame's decision clear of fp32 rounding: push borderline frames over
        x[..., 0] += np.where(np.abs(_p_blank(x) / thr - 1.0) < 1e-2, 1.0, 0.0).astype(np.float32)
        assert np.abs(_p_blank(x) / thr - 1.0).min() > 1e-4
    pri = (rng.random(A) + 0.05).astype(np.float32) if use_priors else None
    xd = torch.from_numpy(x.reshape(Tmax * B, A)).cuda()
    out, kept = decodable.decodable_log_probs(torch, xd, T, B, None if pri is None else torch.from_numpy(pri).cuda(),
                                              scale, thr, 1e-10)
    out = out.cpu().numpy()
    base = 0
    for u in range(B):
        want = pyoracle.decodable(x[:T[u], u], scale, thr, pri, 1e-10)
        assert kept[u] == want.shape[0]
        np.testing.assert_allclose(out[base:base + kept[u]], want, atol=ATOL * abs(scale), rtol=0)
        base += T[u]
    if B > 1 and thr < 1.0:
        assert kept[1] == T[1]
    if thr == 1e-9:
        assert np.array_equal(kept, T)   # nothing passes -> nothing is skipped


def test_probability_input_and_floor():
    """input_is_logits=0: the reference's exact boundary (post-softmax matrix), zeros hit the floor."""
    import torch
    from kaldi_ctc_b200 import decodable
    from oracle import pyoracle
    rng = np.random.default_rng(3)
    T, A = 50, 46
    p = rng.random((T, A)).astype(np.float32)
    p[rng.random((T, A)) < 0.1] = 0.0
    p[:, 0] = np.where(rng.random(T) < 0.5, 0.99, 0.2)
    for floor in (1e-10, 1e-20):
        out, kept = decodable.decodable_log_probs(torch, torch.from_numpy(p).cuda(), [T], 1, None, 1.0, 0.98, floor,
                                                  is_logits=False)
        want = pyoracle.decodable(p, 1.0, 0.98, None, floor, is_logits=False)
        assert kept[0] == want.shape[0] and 0 < kept[0] < T
        np.testing.assert_allclose(out[:kept[0]].cpu().numpy(), want, atol=ATOL, rtol=0)
        assert np.isclose(out[:kept[0]].min().item(), np.log(floor), atol=1e-4)


def test_empty_and_invalid():
    import torch
    from kaldi_ctc_b200 import ctc, decodable
    x = torch.zeros(4, 5, device="cuda")
    out, kept = decodable.decodable_log_probs(torch, x, [0], 1)
    assert out.shape[0] == 0 and kept[0] == 0
    with pytest.raises(ctc.CtcError):
        decodable.decodable_log_probs(torch, x, [4], 1, floor=0.0)


def _model(mode=2, H=32, layers=2, stddev=0.3):
    from kaldi_ctc_b200 import synth
    spec = synth.ModelSpec(mode=mode, layers=layers, D=10, H=H, A=12, learning_rate=0.0, param_stddev=stddev)
    return spec, synth.model_weights(spec, 7)


@pytest.mark.parametrize("mode", [2, 3])
def test_decodable_object_matches_oracle(mode):
    """Inference-mode forward (no reserve) + decodable vs the oracle's forward + decodable."""
    from kaldi_ctc_b200 import decodable, synth
    from oracle import pymodel, pyoracle
    spec, (blobs, aw, ab) = _model(mode)
    T = 37
    x, fl, L, Tl = synth.features(1, spec.D, T, T, 2, 4, spec.A, seed=2)
    ref = pymodel.train_step(spec, blobs, aw, ab, x, fl, L, Tl, 1, dtype=np.float64)
    pri = np.linspace(0.02, 0.2, spec.A).astype(np.float32)
    am = decodable.AmNnet(spec, blobs, aw, ab, priors=pri, max_frames=64)
    tm = decodable.CtcTransitionModel(np.arange(-1, spec.A - 1))   # graph label k -> pdf k-1
    p_blank = _p_blank(ref["logits"])
    thr = float(np.sort(p_blank)[T // 2] + 1e-4)  # skips about half the frames
    dec = decodable.CtcDecodableAmNnet(tm, am, x, True, 0.7, thr)
    want = pyoracle.decodable(ref["logits"], 0.7, thr, pri)
    assert dec.NumFramesReady() == want.shape[0] and 0 < want.shape[0] < T
    assert dec.NumIndices() == spec.A and dec.IsLastFrame(dec.NumFramesReady() - 1)
    np.testing.assert_allclose(dec.log_probs_, want, atol=1e-4, rtol=0)
    assert dec.LogLikelihood(3, 1) == pytest.approx(want[3, 0], abs=1e-4)       # tid 1 = blank
    assert dec.LogLikelihood(3, 5) == pytest.approx(want[3, 4], abs=1e-4)
    par = decodable.CtcDecodableAmNnetParallel(tm, am, x, True, 0.7)
    assert par.NumFramesReady() == T
    wantp = pyoracle.decodable(ref["logits"], 0.7, 1.0, pri, floor=1e-20)
    assert par.LogLikelihood(0, 2) == pytest.approx(wantp[0, 1], abs=1e-4)
    np.testing.assert_allclose(par.log_probs_, wantp, atol=1e-4, rtol=0)


def test_batched_decode_equals_per_utterance():
    """Equal-length utterances through one launch == one utterance at a time (fp32 mode)."""
    from kaldi_ctc_b200 import decodable
    spec, (blobs, aw, ab) = _model(2)
    rng = np.random.default_rng(5)
    feats = [rng.standard_normal((41, spec.D)).astype(np.float32) for _ in range(4)]
    am = decodable.AmNnet(spec, blobs, aw, ab, max_frames=64)
    tm = decodable.CtcTransitionModel(np.arange(-1, spec.A - 1))
    batched = decodable.decode_batch(am, feats, 1.0, 0.3)
    for u, f in enumerate(feats):
        one = decodable.CtcDecodableAmNnet(tm, am, f, True, 1.0, 0.3)
        assert batched[u].shape == one.log_probs_.shape
        np.testing.assert_allclose(batched[u], one.log_probs_, atol=1e-5, rtol=0)


def test_tensor_mode_decode_close_to_fp32():
    from kaldi_ctc_b200 import decodable, rnn
    spec, (blobs, aw, ab) = _model(2, H=64, layers=2, stddev=0.1)
    rng = np.random.default_rng(6)
    feats = [rng.standard_normal((50, spec.D)).astype(np.float32) for _ in range(3)]
    a32 = decodable.AmNnet(spec, blobs, aw, ab, max_frames=64)
    atc = decodable.AmNnet(spec, blobs, aw, ab, max_frames=64, math=rnn.MATH_TENSOR)
    r32 = decodable.decode_batch(a32, feats)
    rtc = decodable.decode_batch(atc, feats)
    for a, b in zip(r32, rtc):
        assert a.shape == b.shape and np.abs(a - b).max() < 5e-2   # bf16 recurrent operands, TF32 projections
