"""One full training step (NnetCtcUpdater::ComputeForMinibatch mirror) on the GPU
against the CPU restatement oracle/pymodel.py: objective, updated weights."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,H,B,Tn", [(2, 32, 4, 24), (3, 16, 3, 20)])
def test_training_step_matches_oracle(mode, H, B, Tn):
    import torch
    from kaldi_ctc_b200 import nnet, synth
    from oracle import pymodel
    spec = synth.ModelSpec(mode=mode, layers=2, D=10, H=H, A=12, learning_rate=0.01, param_stddev=0.2)
    blobs, aw, ab = synth.model_weights(spec, 3)
    x, fl, L, T = synth.features(B, spec.D, Tn - 5, Tn, 2, 5, spec.A, seed=5)
    Tmax = int(T.max())
    ref = pymodel.train_step(spec, blobs, aw, ab, x, fl, L, T, B, dtype=np.float64)
    up = nnet.NnetCtcUpdater(spec, blobs, aw, ab, B, Tmax)
    objf = up.ComputeForMinibatch(torch.from_numpy(x).pin_memory(), Tmax, fl, L, T)
    assert abs(objf - ref["objf"]) < 1e-5 * abs(ref["objf"])
    np.testing.assert_allclose(up.logits[:Tmax * B].cpu().numpy(), ref["logits"], atol=2e-5)
    for l in range(spec.layers):
        got = up.rnns[l].Vectorize()
        # the update itself is lr * clip(dW): compare the applied delta
        d_got, d_ref = got - blobs[l], ref["new_blobs"][l] - blobs[l]
        assert np.abs(d_got - d_ref).max() < 1e-4 * max(1.0, np.abs(d_ref).max() / spec.learning_rate) * spec.learning_rate + 1e-7
    np.testing.assert_allclose(up.affine.linear_params_.cpu().numpy(), ref["new_aff_w"], atol=1e-5)
    np.testing.assert_allclose(up.affine.bias_params_.cpu().numpy(), ref["new_aff_b"], atol=1e-5)
    # a second step runs on the updated model and lowers nothing silently: finite objective
    objf2 = up.ComputeForMinibatch(torch.from_numpy(x).pin_memory(), Tmax, fl, L, T)
    assert np.isfinite(objf2)


def test_tot_accuracy_matches_oracle():
    """ComputeTotAccuracy through the fused arg-max vs the oracle on the product's own logits."""
    import torch
    from kaldi_ctc_b200 import nnet, synth
    from oracle import pyoracle
    spec = synth.ModelSpec(mode=2, layers=1, D=10, H=32, A=12, learning_rate=0.0, param_stddev=0.5)
    blobs, aw, ab = synth.model_weights(spec, 4)
    B, Tn = 6, 40
    x, fl, L, T = synth.features(B, spec.D, Tn - 15, Tn, 3, 8, spec.A, seed=9)
    Tmax = int(T.max())
    up = nnet.NnetCtcUpdater(spec, blobs, aw, ab, B, Tmax)
    up.ComputeForMinibatch(torch.from_numpy(x).pin_memory(), Tmax, fl, L, T, update=False, want_best_pdf=True)
    acc, w = up.ComputeTotAccuracy(Tmax, fl, L, T)
    want = pyoracle.tot_accuracy(up.logits[:Tmax * B].cpu().numpy(), fl, L, T, B)
    assert (acc, w) == want and w == float(np.sum(L))
