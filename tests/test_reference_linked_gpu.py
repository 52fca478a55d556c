"""The drop-in, LINKED AND RUN (VERDICT r01 "what's missing" 2, "do this" 4).

oracle/ref/Makefile compiles the reference's own, unmodified sources -- src/cudamatrix (cu-kernels.cu for
sm_100a), src/nnet2/nnet-cudnn-component.cc, nnet-component.cc, nnet-nnet.cc, src/ctc/ctc-nnet-update.cc,
ctc-nnet-train.cc, ctc-nnet-example.cc and src/ctcbin/nnet2-ctc-train-simple.cc -- and links them with
kaldi_ctc_b200/libb200cudnn.so as -lcudnn and libb200ctc.so as -lwarpctc:

  oracle/_ref/ref_component_harness   integration/kaldi/harness/component_harness.cc on those objects
  oracle/_ref/ref_ctc_train           the reference's training binary itself
  oracle/_ref/ref_nnet3_harness       the nnet3 adapter (integration/kaldi/nnet3) on the reference's src/nnet3 objects

These tests run the binaries on the GPU and compare what the REFERENCE'S host code computed on top of this
repo's kernels with the CPU oracle: CuDNNRecurrentComponent::InitFromString -> Propagate -> Backprop
(nnet-cudnn-component.cc:72-98, 508-610), NnetCtcUpdater::ComputeForMinibatch (ctc-nnet-update.cc:94-127, the
warp-ctc call at :211-243), and whole runs of nnet2-ctc-train-simple (TrainNnetSimple, ctc-nnet-train.cc:181-284,
with and without momentum) whose written model is read back with model_io."""
import os
import re

import numpy as np
import pytest

import refbin
from kaldi_ctc_b200 import egs, model_io, synth
from oracle import pymodel, pyoracle

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not refbin.ensure_built("gpu"), reason="oracle/_ref binaries are not built")]


def _spec(mode=2, layers=2, D=10, H=32, A=12, lr=0.01):
    return synth.ModelSpec(mode=mode, layers=layers, D=D, H=H, A=A, learning_rate=lr, param_stddev=0.2)


@pytest.mark.parametrize("mode,bidir,math", [(2, True, "fp32"), (3, True, "fp32"), (2, False, "fp32"), (1, True, "fp32"),
                                             (2, True, "tensor")])
def test_reference_component_on_the_dropin(tmp_path, mode, bidir, math):
    D, H, B, T = 24, 64, 4, 12
    dirs = 2 if bidir else 1
    rng = np.random.default_rng(5)
    n = pyoracle.rnn_param_count(mode, bidir, 1, D, H)
    w = (rng.standard_normal(n) * 0.2).astype(np.float32)
    x = rng.standard_normal((T * B, D)).astype(np.float32)
    dy = rng.standard_normal((T * B, H * dirs)).astype(np.float32)
    for name, arr in (("w", w), ("x", x), ("dy", dy)):
        arr.tofile(tmp_path / (name + ".f32"))
    lr, clip = 0.05, 0.7
    cfg = ("learning-rate=%g num-layers=1 input-dim=%d output-dim=%d rnn-mode=%d bidirectional=%s max-seq-length=16 "
           "clip-gradient=%g" % (lr, D, H, mode, "true" if bidir else "false", clip))
    env = {"B200_CUDNN_MATH": "tensor"} if math == "tensor" else {"B200_CUDNN_MATH": ""}
    out = refbin.run(refbin.HARNESS, "component", cfg, B, tmp_path / "w.f32", tmp_path / "x.f32", tmp_path / "dy.f32",
                     tmp_path / "o", env=env).stdout
    assert "CuDNNRecurrentComponent" in out
    y = np.fromfile(tmp_path / "o.y.f32", np.float32).reshape(T * B, H * dirs)
    dx = np.fromfile(tmp_path / "o.dx.f32", np.float32).reshape(T * B, D)
    w2 = np.fromfile(tmp_path / "o.w.f32", np.float32)
    yr, dxr, dwr = pyoracle.rnn(mode, bidir, 1, H, x, w, B, dy=dy, dtype=np.float64)
    w2r = w + lr * np.clip(dwr, -clip, clip)          # ApplyFloor/ApplyCeiling + Update (:602-614)
    ty, tg = (1e-5, 1e-4) if math == "fp32" else (5e-3, 1e-2)
    assert np.abs(y - yr).max() < ty
    assert np.abs(dx - dxr).max() < tg * max(1.0, np.abs(dxr).max())
    assert np.abs(w2 - w2r).max() < tg * lr * max(1.0, np.abs(dwr).max()) + 1e-7
    assert np.abs(w2 - w).max() > 1e-4                # the update really happened


@pytest.mark.parametrize("mode,bidir,exact", [(2, True, 1), (3, True, 1), (2, False, 1), (2, True, 0)])
def test_nnet3_adapter_linked_against_reference_nnet3(tmp_path, mode, bidir, exact):
    """integration/kaldi/nnet3/nnet-b200-recurrent-component.h compiled and LINKED with the reference's src/nnet3
    objects (nnet-component-itf.cc, nnet-parse.cc, nnet-common.cc, ...), driven through the nnet3 Component surface
    (src/nnet3/nnet-component-itf.h:116-165): InitFromConfig, ReorderIndexes (the harness hands over n-major
    indexes), PrecomputeIndexes, Propagate, Backprop with to_update, Write/Read, Vectorize."""
    D, H, B, T = 24, 64, 4, 12
    dirs = 2 if bidir else 1
    rng = np.random.default_rng(9)
    n = pyoracle.rnn_param_count(mode, bidir, 1, D, H)
    w = (rng.standard_normal(n) * 0.2).astype(np.float32)
    x = rng.standard_normal((T * B, D)).astype(np.float32)
    dy = rng.standard_normal((T * B, H * dirs)).astype(np.float32)
    for name, arr in (("w", w), ("x", x), ("dy", dy)):
        arr.tofile(tmp_path / (name + ".f32"))
    lr, clip = 0.05, 0.7
    cfg = ("input-dim=%d output-dim=%d learning-rate=%g num-layers=1 rnn-mode=%d bidirectional=%s max-seq-length=16 "
           "clip-gradient=%g exact-fp32=%d" % (D, H, lr, mode, "true" if bidir else "false", clip, exact))
    out = refbin.run(refbin.NNET3, cfg, T, B, tmp_path / "w.f32", tmp_path / "x.f32", tmp_path / "dy.f32", tmp_path / "o").stdout
    assert "B200RecurrentComponent" in out
    y = np.fromfile(tmp_path / "o.y.f32", np.float32).reshape(T * B, H * dirs)
    dx = np.fromfile(tmp_path / "o.dx.f32", np.float32).reshape(T * B, D)
    w2 = np.fromfile(tmp_path / "o.w.f32", np.float32)
    yr, dxr, dwr = pyoracle.rnn(mode, bidir, 1, H, x, w, B, dy=dy, dtype=np.float64)
    w2r = w + lr * np.clip(dwr, -clip, clip)
    ty, tg = (1e-5, 1e-4) if exact else (5e-3, 1e-2)
    assert np.abs(y - yr).max() < ty
    assert np.abs(dx - dxr).max() < tg * max(1.0, np.abs(dxr).max())
    assert np.abs(w2 - w2r).max() < tg * lr * max(1.0, np.abs(dwr).max()) + 1e-7
    assert np.abs(w2 - w).max() > 1e-4


def _write_model_and_egs(tmp_path, spec, n_utts, t_lo, t_hi, l_lo, l_hi, seed):
    blobs, aw, ab = synth.model_weights(spec, 3)
    comps = model_io.components_of(spec, blobs, aw, ab, max_seq_length=t_hi + 8)
    with open(tmp_path / "in.nnet", "wb") as f:
        model_io.write_nnet(f, comps)
    ex = synth.examples(n_utts, spec.D, t_lo, t_hi, l_lo, l_hi, spec.A, seed=seed)
    egs.write_egs_ark(str(tmp_path / "egs.ark"), [("utt%03d" % i, e) for i, e in enumerate(ex)])
    return blobs, aw, ab, ex


def _minibatch_arrays(ex):
    """What FormatNnetInput + ComputeObjfAndDeriv hand to the network and to warp-ctc for these examples."""
    x, T = pyoracle.format_nnet_input([e.input_frames.blob for e in ex], None, 0, 0, 0)
    il = np.array([e.NumFrames() for e in ex], np.int32)
    ll = np.array([e.NumLabels() for e in ex], np.int32)
    fl = np.concatenate([np.asarray(e.labels, np.int32) for e in ex])
    return x, fl, ll, il


@pytest.mark.parametrize("mode", [2, 3])
def test_reference_updater_step_on_the_dropin(tmp_path, mode):
    """NnetCtcUpdater::ComputeForMinibatch of the reference, its own FormatNnetInput / Propagate /
    ComputeObjfAndDeriv / ComputeTotAccuracy / Backprop, on libb200cudnn.so + libb200ctc.so."""
    spec, B = _spec(mode=mode), 4
    blobs, aw, ab, ex = _write_model_and_egs(tmp_path, spec, B, 20, 28, 2, 5, seed=11)
    refbin.run(refbin.HARNESS, "step", tmp_path / "in.nnet", "ark:%s" % (tmp_path / "egs.ark"), B, tmp_path / "s", "update")
    objf, acc = [float(v) for v in open(tmp_path / "s.objf.txt").read().split()]
    x, fl, ll, il = _minibatch_arrays(ex)
    ref = pymodel.train_step(spec, blobs, aw, ab, x, fl, ll, il, B, dtype=np.float64)
    assert abs(objf - ref["objf"]) < 1e-5 * abs(ref["objf"])
    logits = np.fromfile(tmp_path / "s.output.f32", np.float32).reshape(ref["logits"].shape)
    assert np.abs(logits - ref["logits"]).max() < 2e-5
    want_acc, _ = pyoracle.tot_accuracy(logits, fl, ll, il, B)
    assert acc == want_acc
    with open(tmp_path / "s.nnet", "rb") as f:
        trained = model_io.read_nnet(f)
    rnns = [c for c in trained if c["type"] == "CuDNNRecurrentComponent"]
    lr = spec.learning_rate
    for l, c in enumerate(rnns):
        d_got, d_ref = c["filter_params"] - blobs[l], ref["new_blobs"][l] - blobs[l]
        assert np.abs(d_got - d_ref).max() < 1e-4 * lr * max(1.0, np.abs(d_ref).max() / lr) + 1e-7
    aff = [c for c in trained if c["type"] == "AffineComponent"][0]
    np.testing.assert_allclose(aff["linear_params"], ref["new_aff_w"], atol=1e-5)
    np.testing.assert_allclose(aff["bias_params"], ref["new_aff_b"], atol=1e-5)


@pytest.mark.parametrize("momentum", [0.0, 0.9])
def test_reference_training_binary_on_the_dropin(tmp_path, momentum):
    """src/ctcbin/nnet2-ctc-train-simple.cc, unmodified: reads the model (TransitionModel + AmNnet) and the egs
    archive, trains 3 minibatches of 4 utterances (background reader, DoBackprop, momentum through delta_nnet),
    writes the model.  The oracle replays the same minibatches on the CPU in fp64."""
    spec, B, n_mb = _spec(), 4, 3
    blobs, aw, ab, ex = _write_model_and_egs(tmp_path, spec, B * n_mb, 18, 26, 2, 5, seed=21)
    refbin.run(refbin.HARNESS, "make-model", tmp_path / "in.nnet", spec.A - 1, tmp_path / "in.mdl")
    r = refbin.run(refbin.TRAIN, "--minibatch-size=%d" % B, "--momentum=%g" % momentum, "--max-allow-frames=100",
                   tmp_path / "in.mdl", "ark:%s" % (tmp_path / "egs.ark"), tmp_path / "out.mdl")
    log = r.stderr + r.stdout
    m = re.search(r"Did backprop on (\S+) examples, average log-prob per frame is (\S+)", log)
    assert m, log[-2000:]
    tot_weight, avg = float(m.group(1)), float(m.group(2))
    refbin.run(refbin.HARNESS, "extract-nnet", tmp_path / "out.mdl", tmp_path / "out.nnet")
    with open(tmp_path / "out.nnet", "rb") as f:
        trained = model_io.read_nnet(f)
    # oracle replay: delta += lr*clip(g); w += delta; delta *= momentum  (ctc-nnet-train.cc:220-245)
    cur_b, cur_w, cur_ab = [b.astype(np.float64) for b in blobs], aw.astype(np.float64), ab.astype(np.float64)
    d_b, d_w, d_ab = [np.zeros_like(b) for b in cur_b], np.zeros_like(cur_w), np.zeros_like(cur_ab)
    tot_objf, tot_labels = 0.0, 0
    for k in range(n_mb):
        mb = ex[k * B:(k + 1) * B]
        x, fl, ll, il = _minibatch_arrays(mb)
        ref = pymodel.train_step(spec, [b.astype(np.float32) for b in cur_b], cur_w.astype(np.float32),
                                 cur_ab.astype(np.float32), x, fl, ll, il, B, dtype=np.float64)
        tot_objf += ref["objf"]
        tot_labels += int(ll.sum())
        for l in range(spec.layers):
            d_b[l] += ref["new_blobs"][l] - cur_b[l].astype(np.float32)
            cur_b[l] = cur_b[l] + d_b[l]
            d_b[l] *= momentum
        d_w += ref["new_aff_w"] - cur_w.astype(np.float32)
        cur_w = cur_w + d_w
        d_w *= momentum
        d_ab += ref["new_aff_b"] - cur_ab.astype(np.float32)
        cur_ab = cur_ab + d_ab
        d_ab *= momentum
    # TrainNnetSimple's "tot_weight" is TotalNnetTrainingWeight = number of labels; objective = sum of costs
    assert tot_weight == tot_labels
    assert abs(avg - tot_objf / tot_labels) < 2e-4 * abs(tot_objf / tot_labels)
    rnns = [c for c in trained if c["type"] == "CuDNNRecurrentComponent"]
    for l, c in enumerate(rnns):
        scale = max(1e-3, np.abs(cur_b[l] - blobs[l]).max())
        assert np.abs(c["filter_params"] - cur_b[l]).max() < 2e-3 * scale, "layer %d" % l
    aff = [c for c in trained if c["type"] == "AffineComponent"][0]
    assert np.abs(aff["linear_params"] - cur_w).max() < 2e-3 * np.abs(cur_w - aw).max()
    assert np.abs(aff["bias_params"] - cur_ab).max() < 2e-3 * max(1e-3, np.abs(cur_ab - ab).max())


def test_python_mirror_momentum_equals_reference_semantics():
    """kaldi_ctc_b200.nnet.NnetCtcUpdater(momentum=0.9) (b200rnnUpdate through delta buffers) against the same
    oracle replay as the reference binary above."""
    import torch
    from kaldi_ctc_b200 import nnet
    spec, B, momentum = _spec(), 4, 0.9
    blobs, aw, ab = synth.model_weights(spec, 3)
    up = nnet.NnetCtcUpdater(spec, blobs, aw, ab, B, 32, momentum=momentum)
    cur_b, cur_w = [b.astype(np.float64) for b in blobs], aw.astype(np.float64)
    d_b, d_w = [np.zeros_like(b) for b in cur_b], np.zeros_like(cur_w)
    for k in range(3):
        x, fl, L, T = synth.features(B, spec.D, 20, 28, 2, 5, spec.A, seed=40 + k)
        Tmax = int(T.max())
        ref = pymodel.train_step(spec, [b.astype(np.float32) for b in cur_b], cur_w.astype(np.float32), up.affine.bias_params_.cpu().numpy(),
                                 x, fl, L, T, B, dtype=np.float64)
        up.ComputeForMinibatch(torch.from_numpy(x).pin_memory(), Tmax, fl, L, T)
        for l in range(spec.layers):
            d_b[l] += ref["new_blobs"][l] - cur_b[l].astype(np.float32)
            cur_b[l] = cur_b[l] + d_b[l]
            d_b[l] *= momentum
        d_w += ref["new_aff_w"] - cur_w.astype(np.float32)
        cur_w = cur_w + d_w
        d_w *= momentum
    for l in range(spec.layers):
        assert np.abs(up.rnns[l].Vectorize() - cur_b[l]).max() < 1e-3 * max(1e-3, np.abs(cur_b[l] - blobs[l]).max())
    assert np.abs(up.affine.linear_params_.cpu().numpy() - cur_w).max() < 1e-3 * np.abs(cur_w - aw).max()
