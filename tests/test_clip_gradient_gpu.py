"""ClipGradientComponent::Backprop on the device (b200rnnClipGradientBackprop): row-norm clipping, the component's
counters, and the stochastic self-repair term of RepairGradients (src/nnet2/nnet-cudnn-component.cc:936-1055)
against the numpy restatement (the coin of :981 is forced, so the comparison is deterministic)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols,thr,scale,target", [(37, 64, 2.0, 1.0, 0.0), (1000, 640, 20.0, 0.5, 0.3), (5, 7, 1.0, 1.0, 0.0)])
def test_clip_counters_and_self_repair(rows, cols, thr, scale, target):
    import torch
    from kaldi_ctc_b200 import nnet
    from oracle import pyoracle
    rng = np.random.default_rng(rows)
    comp = nnet.ClipGradientComponent(cols, clipping_threshold=thr, self_repair_scale=scale, self_repair_target=target)
    ref_counters = [0, 0, 0, 0]
    for it in range(4):
        d = (rng.standard_normal((rows, cols)) * rng.uniform(0.05, 1.5, size=(rows, 1))).astype(np.float32)
        v = np.tanh(rng.standard_normal((rows, cols))).astype(np.float32)
        attempt = it % 2 == 1
        want = pyoracle.clip_gradient_backprop(d, v, thr, ref_counters, target=target, scale=scale, attempt_repair=attempt)
        dt = torch.from_numpy(d).cuda()
        comp.Backprop(dt, in_value=torch.from_numpy(v).cuda(), to_update=comp, force_attempt=attempt)
        got = dt.cpu().numpy()
        assert [int(x) for x in comp.counters.cpu()] == ref_counters
        assert np.abs(got - want).max() < 2e-5 * max(1.0, np.abs(want).max())
    assert ref_counters[0] > 0 and ref_counters[2] > 0       # rows were clipped and a repair really happened
    # no statistics, no repair when there is nothing to update (to_update == NULL in the reference)
    before = comp.counters.clone()
    d = (rng.standard_normal((rows, cols)) * 3).astype(np.float32)
    dt = torch.from_numpy(d).cuda()
    comp.Backprop(dt, in_value=None, to_update=None)
    assert torch.equal(before, comp.counters)
    n = np.sqrt((dt.cpu().numpy().astype(np.float64) ** 2).sum(1))
    assert n.max() <= thr * (1 + 1e-5)
