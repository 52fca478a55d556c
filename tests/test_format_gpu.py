"""b200ctc_format_input (GPU FormatNnetInput + CompressedMatrix decompression) against the oracle's
restatement of ctc-nnet-update.cc:351-424 / compressed-matrix.cc:493-529 -- BIT-EXACT."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _egs(rng, lengths, D, left_context, spk_dim, scale=3.0):
    from kaldi_ctc_b200 import egs
    out = []
    for T in lengths:
        m = (rng.standard_normal((T, D)) * scale + rng.standard_normal()).astype(np.float32)
        out.append(egs.NnetCtcExample([1], egs.CompressedMatrix.from_matrix(m), left_context,
                                      rng.standard_normal(spk_dim).astype(np.float32) if spk_dim else []))
    return out


def _check(examples, nl, nr):
    import torch
    from kaldi_ctc_b200 import egs
    from oracle import pyoracle
    spk = [e.spk_info for e in examples] if examples[0].spk_info.size else None
    want, mf = pyoracle.format_nnet_input([e.input_frames.blob for e in examples], spk, examples[0].left_context, nl, nr)
    got, mf2 = egs.FormatNnetInput(nl, nr, examples)
    torch.cuda.synchronize()
    assert mf2 == mf and tuple(got.shape) == want.shape
    g = got.cpu().numpy()
    assert np.array_equal(g.view(np.uint32), want.view(np.uint32)), "max diff %g" % np.abs(g - want).max()
    return g


@pytest.mark.parametrize("lengths,D,left,nl,nr,spk", [
    ([50, 37, 64, 12], 40, 0, 0, 0, 0),          # the RNN recipe: no splicing
    ([9, 8, 3, 20], 13, 0, 0, 0, 0),             # formats 1 and 2 mixed (<= 8 rows -> uint16 storage)
    ([30, 45, 33], 7, 3, 1, 2, 4),               # splice 4, extra left context ignored, speaker vector
    ([200], 40, 0, 0, 0, 0),                     # minibatch 1
    ([65, 64, 63, 129, 1], 5, 0, 0, 0, 1),       # tile edges of the 64-step CTAs, a 1-frame utterance
])
def test_matches_oracle_bit_exact(lengths, D, left, nl, nr, spk):
    rng = np.random.default_rng(sum(lengths) + D)
    _check(_egs(rng, lengths, D, left, spk), nl, nr)


def test_benchmark_shape_and_padding():
    rng = np.random.default_rng(7)
    lengths = list(rng.integers(1200, 2001, size=16))
    ex = _egs(rng, lengths, 40, 0, 0)
    g = _check(ex, 0, 0).reshape(max(lengths), 16, 40)
    for b, T in enumerate(lengths):
        assert not g[T:, b].any()


def test_staging_is_double_buffered():
    """Two different minibatches formatted back to back through one stager, no sync in between."""
    import torch
    from kaldi_ctc_b200 import egs
    from oracle import pyoracle
    rng = np.random.default_rng(8)
    st = egs.InputStager()
    batches = [_egs(rng, [40, 31], 12, 0, 0), _egs(rng, [25, 44], 12, 0, 0), _egs(rng, [33, 10], 12, 0, 0)]
    outs = [egs.FormatNnetInput(0, 0, b, stager=st)[0] for b in batches]
    torch.cuda.synchronize()
    for b, o in zip(batches, outs):
        want, _ = pyoracle.format_nnet_input([e.input_frames.blob for e in b], None, 0, 0, 0)
        assert np.array_equal(o.cpu().numpy(), want)


def test_invalid_inputs_are_rejected():
    from kaldi_ctc_b200 import ctc, egs
    rng = np.random.default_rng(9)
    ex = _egs(rng, [20, 20], 6, 0, 0)
    ex[1] = _egs(rng, [20], 7, 0, 0)[0]            # feature dimension differs
    with pytest.raises(ctc.CtcError):
        egs.FormatNnetInput(0, 0, ex)
    with pytest.raises(ctc.CtcError):
        egs.FormatNnetInput(2, 0, _egs(rng, [20], 6, 1, 0))   # left_context < nnet.LeftContext() (:366)


def test_training_step_from_examples_equals_step_from_slab():
    import torch
    from kaldi_ctc_b200 import nnet, synth
    from oracle import pyoracle
    spec = synth.ModelSpec(mode=2, layers=2, D=10, H=32, A=12, learning_rate=0.01, param_stddev=0.2)
    blobs, aw, ab = synth.model_weights(spec, 3)
    ex = synth.examples(4, spec.D, 20, 30, 2, 5, spec.A, seed=4)
    slab, T = pyoracle.format_nnet_input([e.input_frames.blob for e in ex], None, 0, 0, 0)
    fl = np.concatenate([np.asarray(e.labels, dtype=np.int32) for e in ex])
    ll = np.array([e.NumLabels() for e in ex])
    il = np.array([e.NumFrames() for e in ex])
    a = nnet.NnetCtcUpdater(spec, blobs, aw, ab, 4, T)
    b = nnet.NnetCtcUpdater(spec, blobs, aw, ab, 4, T)
    oa = a.ComputeForMinibatchFromExamples(ex)
    ob = b.ComputeForMinibatch(torch.from_numpy(slab).pin_memory(), T, fl, ll, il)
    assert oa == ob
    for ca, cb in zip(a.rnns, b.rnns):
        assert torch.equal(ca.filter_params_, cb.filter_params_)
