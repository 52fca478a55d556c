"""Parity of the CUDA CTC path (libb200ctc.so, through the warp-ctc C ABI) with
the oracle.  Tolerances are north_star's: loss 1e-5 relative, gradient 1e-4
max-abs -- held against the fp64 instantiation of the oracle (and the committed
torch-fp64 goldens), which is stricter than matching the fp32 one."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

LOSS_RTOL = 1e-5
GRAD_ATOL = 1e-4


@pytest.fixture(scope="module")
def ctx():
    import torch
    from kaldi_ctc_b200 import ctc
    assert torch.cuda.is_available()
    return torch, ctc, ctc.CtcLoss("cuda:0")


def _run(ctx, act, fl, ll, il, **kw):
    torch, ctc, op = ctx
    a = torch.from_numpy(np.ascontiguousarray(act)).cuda()
    costs, grad = op.compute(a, fl, ll, il, **kw)
    torch.cuda.synchronize()
    return costs, (grad.cpu().numpy() if grad is not None else None)


def _check(costs, grad, c_ref, g_ref):
    np.testing.assert_allclose(costs, c_ref, rtol=LOSS_RTOL)
    assert np.abs(grad - g_ref).max() < GRAD_ATOL


@pytest.mark.parametrize("name", ["ctc_small", "ctc_repeats", "ctc_ragged", "ctc_peaky", "ctc_wide"])
def test_matches_committed_golden(ctx, golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"))
    costs, grad = _run(ctx, z["activations"], z["flat_labels"], z["label_lengths"], z["input_lengths"])
    _check(costs, grad, z["costs"], z["grads"])


@pytest.mark.parametrize("cfg,peaky", [(1, False), (1, True), (4, False)])
def test_baseline_configs_against_fp64_oracle(ctx, cfg, peaky):
    from kaldi_ctc_b200 import synth
    from oracle import pyoracle
    bt = synth.config_ctc(cfg, peaky=peaky)
    c_ref, g_ref = pyoracle.ctc(bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths,
                                dtype=np.float64)
    costs, grad = _run(ctx, bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths)
    _check(costs, grad, c_ref, g_ref)
    for b, Tb in enumerate(bt.input_lengths):
        assert np.all(grad[Tb:, b, :] == 0)   # padded rows exactly zero
    # the fp32 restatement of warp-ctc's arithmetic agrees to ITS accuracy
    c32, g32 = pyoracle.ctc(bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths,
                            dtype=np.float32)
    np.testing.assert_allclose(costs, c32, rtol=LOSS_RTOL)
    assert np.abs(grad - g32).max() < np.abs(g32 - g_ref).max() + GRAD_ATOL


def test_wide_alphabet_slice_of_config5(ctx):
    """Config 5 (A=4000) at B=4: parity with the oracle + row-sum property."""
    from kaldi_ctc_b200 import synth
    from oracle import pyoracle
    bt = synth.ctc_batch(4, 4000, 300, 500, 50, 200, seed=1005, sigma=2.0)
    c_ref, g_ref = pyoracle.ctc(bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths,
                                dtype=np.float64)
    costs, grad = _run(ctx, bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths)
    _check(costs, grad, c_ref, g_ref)
    assert np.abs(grad.sum(-1)).max() < 1e-4


@pytest.mark.parametrize("A,groups", [(4000, 1), (4000, 3), (1028, 2)])
def test_wide_alphabet_ring_kernel(ctx, A, groups):
    """The persistent TMA-ring gradient kernel (normally chosen for slabs >= 64 MB), forced on a small
    ragged batch: an empty label string, P=1 and P=2 lattices in one call is not possible (P is per call),
    so L up to 300 gives P=2; one infeasible utterance (zero row block), padded frames, several utterance
    groups on side streams."""
    from oracle import pyoracle
    ctc = ctx[1]
    # (b200ctc_set_tuning, not the environment: the library reads B200CTC_* once per process)
    ctc.set_tuning("RING", 1)
    ctc.set_tuning("GROUPS", groups)
    rng = np.random.default_rng(A + groups)
    il = np.array([90, 400, 37, 333, 20, 256, 1], np.int32)
    ll = np.array([0, 300, 12, 150, 30, 100, 1], np.int32)   # utterance 4: L=30 > T=20 -> infeasible
    T, B = int(il.max()), len(il)
    act = (rng.standard_normal((T, B, A)) * 2).astype(np.float32)
    for b in range(B):
        act[il[b]:, b, :] = 0
    fl = np.concatenate([rng.integers(1, A, size=int(l)) for l in ll]).astype(np.int32)
    c_ref, g_ref = pyoracle.ctc(act, fl, ll, il, dtype=np.float64)
    try:
        costs, grad = _run(ctx, act, fl, ll, il)
    except Exception:
        ctc.set_tuning("RING", -1)
        ctc.set_tuning("GROUPS", 0)
        raise
    feas = np.array([0, 1, 2, 3, 5, 6])
    np.testing.assert_allclose(costs[feas], c_ref[feas], rtol=LOSS_RTOL)
    assert np.abs(grad[:, feas] - g_ref[:, feas]).max() < GRAD_ATOL
    assert np.abs(grad[:, 4]).max() == 0
    for b in range(B):
        assert np.abs(grad[il[b]:, b]).max(initial=0) == 0
    # and it is the same function as the row-per-warp kernel
    try:
        ctc.set_tuning("RING", 0)
        costs2, grad2 = _run(ctx, act, fl, ll, il)
    finally:
        ctc.set_tuning("RING", -1)
        ctc.set_tuning("GROUPS", 0)
    np.testing.assert_array_equal(costs, costs2)
    assert np.abs(grad - grad2).max() < 1e-6


@pytest.mark.parametrize("L,T,A", [(0, 7, 5), (1, 3, 4), (511, 1100, 40), (512, 1100, 40),
                                   (700, 1500, 31), (1500, 3100, 9), (2047, 4100, 6)])
def test_label_length_edges(ctx, L, T, A):
    """L=0 (blank only), and every pairs-per-thread variant (P=1: L<=511, P=2: L<=1023,
    P=4: L<=2047).  The reference's warp-ctc stopped at L=639 (ctc-nnet-train.cc:25-26)."""
    from oracle import pyoracle
    rng = np.random.default_rng(L + T)
    act = (rng.standard_normal((T, 2, A)) * 2).astype(np.float32)
    labs = [rng.integers(1, A, size=L).astype(np.int32), rng.integers(1, A, size=max(L - 1, 0)).astype(np.int32)]
    il = np.array([T, T - 1], np.int32)
    act[T - 1:, 1, :] = 0
    fl = np.concatenate(labs) if L else np.zeros(0, np.int32)
    ll = np.array([len(l) for l in labs], np.int32)
    c_ref, g_ref = pyoracle.ctc(act, fl, ll, il, dtype=np.float64)
    costs, grad = _run(ctx, act, fl, ll, il)
    _check(costs, grad, c_ref, g_ref)


@pytest.mark.parametrize("seed", range(12))
def test_random_ragged_batches_against_oracle(ctx, seed):
    """Seeded random batches that stress the alpha/beta frame loop: tiny alphabets (many repeated labels), label
    strings as long as the frames allow (tight alignments: every state on the edge of the reachable cone carries the
    path), long T with short L (the forward-tilted lattice that broke the rejected shared-exponent variant), peaky
    logits, every pairs-per-thread variant forced in turn (B200CTC_P hook), T not a multiple of the 8-frame block."""
    from oracle import pyoracle
    ctc = ctx[1]
    rng = np.random.default_rng(9000 + seed)
    B = int(rng.integers(1, 7))
    A = int(rng.choice([2, 3, 5, 12, 48, 130]))
    kind = seed % 4
    il, labs = [], []
    for b in range(B):
        if kind == 0:      # tight: L + repeats == T or T - 1
            L = int(rng.integers(1, 60))
            lab = rng.integers(1, A, size=L)
            rep = int((lab[1:] == lab[:-1]).sum())
            T = L + rep + int(rng.integers(0, 2))
        elif kind == 1:    # long T, short L
            L = int(rng.integers(0, 6))
            lab = rng.integers(1, A, size=L)
            T = int(rng.integers(300, 900))
        elif kind == 2:    # many repeats, medium
            L = int(rng.integers(20, 200))
            lab = rng.integers(1, min(A, 3), size=L) if A > 2 else np.ones(L, np.int64)
            T = L + int((lab[1:] == lab[:-1]).sum()) + int(rng.integers(0, 300))
        else:              # general
            L = int(rng.integers(1, 300))
            lab = rng.integers(1, A, size=L)
            T = L + int((lab[1:] == lab[:-1]).sum()) + int(rng.integers(1, 500))
        il.append(T)
        labs.append(lab.astype(np.int32))
    il = np.array(il, np.int32)
    ll = np.array([len(l) for l in labs], np.int32)
    fl = np.concatenate(labs) if ll.sum() else np.zeros(0, np.int32)
    Tm = int(il.max())
    scale = 6.0 if seed % 3 == 0 else 2.0   # peaky / ordinary logits
    act = (rng.standard_normal((Tm, B, A)) * scale).astype(np.float32)
    for b in range(B):
        act[il[b]:, b, :] = 0
    c_ref, g_ref = pyoracle.ctc(act, fl, ll, il, dtype=np.float64)
    try:
        for P in (0, 2, 4):
            ctc.set_tuning("P", P)
            costs, grad = _run(ctx, act, fl, ll, il)
            np.testing.assert_allclose(costs, c_ref, rtol=LOSS_RTOL, atol=1e-5)
            assert np.abs(grad - g_ref).max() < GRAD_ATOL, "P=%d kind=%d" % (P, kind)
            for b in range(B):
                assert np.abs(grad[il[b]:, b]).max(initial=0) == 0
    finally:
        ctc.set_tuning("P", 0)


def test_alphabet_not_multiple_of_four_and_nonzero_blank(ctx):
    from oracle import pyoracle
    rng = np.random.default_rng(2)
    T, B, A, blank = 50, 3, 7, 3
    act = rng.standard_normal((T, B, A)).astype(np.float32)
    labs = [np.array([1, 1, 2, 6], np.int32), np.array([5], np.int32), np.array([0, 4, 0], np.int32)]
    fl, ll, il = np.concatenate(labs), np.array([4, 1, 3], np.int32), np.array([50, 20, 33], np.int32)
    for b in range(B):
        act[il[b]:, b, :] = 0
    c_ref, g_ref = pyoracle.ctc(act, fl, ll, il, blank=blank, dtype=np.float64)
    costs, grad = _run(ctx, act, fl, ll, il, blank=blank)
    _check(costs, grad, c_ref, g_ref)


def test_loss_only_and_infeasible_and_errors(ctx):
    torch, ctc, op = ctx
    from oracle import pyoracle
    rng = np.random.default_rng(3)
    act = rng.standard_normal((6, 2, 5)).astype(np.float32)
    fl, ll, il = np.array([1, 1, 1, 2], np.int32), np.array([3, 1], np.int32), np.array([4, 6], np.int32)
    act[4:, 0, :] = 0
    # utterance 0: L + repeats = 5 > T = 4 -> cost 0, zero gradient (warp-ctc behaviour)
    costs, grad = _run(ctx, act, fl, ll, il)
    c_ref, g_ref = pyoracle.ctc(act, fl, ll, il, dtype=np.float64)
    assert costs[0] == 0 and np.all(grad[:, 0, :] == 0)
    _check(costs, grad, c_ref, g_ref)
    costs2, none = _run(ctx, act, fl, ll, il, want_grad=False)
    assert none is None
    np.testing.assert_array_equal(costs, costs2)
    a = torch.from_numpy(act).cuda()
    with pytest.raises(ctc.CtcError):
        op.compute(a, np.array([0, 1, 1, 2], np.int32), ll, il)   # blank inside the labels
    with pytest.raises(ctc.CtcError):
        op.compute(a, np.array([5, 1, 1, 2], np.int32), ll, il)   # label >= alphabet


def test_extended_entry_point(ctx):
    """b200ctc_loss: grad_scale=-1 (fuses NnetCtcUpdater::Backprop's Scale(-1),
    ctc-nnet-update.cc:323), device costs, no host sync."""
    torch, ctc, op = ctx
    from kaldi_ctc_b200 import synth
    bt = synth.ctc_batch(5, 33, 60, 90, 5, 20, seed=9)
    a = torch.from_numpy(bt.activations).cuda()
    costs, grad = op.compute(a, bt.flat_labels, bt.label_lengths, bt.input_lengths)
    g2 = torch.full_like(a, 7.0)
    cd = torch.zeros(5, device="cuda")
    op.compute_extended(a, bt.flat_labels, bt.label_lengths, bt.input_lengths, gradients=g2,
                        grad_scale=-1.0, costs_dev=cd, no_sync=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(cd.cpu().numpy(), costs)
    np.testing.assert_array_equal(g2.cpu().numpy(), -grad.cpu().numpy())


def test_deterministic(ctx):
    from kaldi_ctc_b200 import synth
    bt = synth.ctc_batch(8, 48, 300, 400, 60, 90, seed=4)
    r1 = _run(ctx, bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths)
    r2 = _run(ctx, bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths)
    np.testing.assert_array_equal(r1[0], r2[0])
    np.testing.assert_array_equal(r1[1], r2[1])


def test_fused_row_argmax_matches_find_row_max_id(ctx):
    """argmax_dev of b200ctc_loss == CuMatrix::FindRowMaxId on the valid rows (first maximum wins),
    -1 on padded rows; costs/gradients unchanged by asking for it."""
    torch, ctc, op = ctx
    from kaldi_ctc_b200 import synth
    for A, seed in ((46, 1), (131, 2), (8000, 3), (7, 4)):
        bt = synth.ctc_batch(5, A, 20, 60, 2, 9, seed)
        act = bt.activations.copy()
        T, B, _ = act.shape
        act[3, 1, :] = 0.25            # full tie -> index 0
        act[4, 2, A - 1] = act[4, 2, 2] = 50.0   # two-way tie -> the lower index
        a = torch.from_numpy(act).cuda()
        g0 = torch.empty_like(a)
        c0 = op.compute_extended(a, bt.flat_labels, bt.label_lengths, bt.input_lengths, gradients=g0)
        g1 = torch.empty_like(a)
        am = torch.full((T * B,), -7, dtype=torch.int32, device="cuda")
        c1 = op.compute_extended(a, bt.flat_labels, bt.label_lengths, bt.input_lengths, gradients=g1,
                                 argmax_dev=am)
        torch.cuda.synchronize()
        assert np.array_equal(c0, c1) and torch.equal(g0, g1)
        got = am.cpu().numpy().reshape(T, B)
        want = act.argmax(axis=2)
        for b, Tb in enumerate(bt.input_lengths):
            assert np.array_equal(got[:Tb, b], want[:Tb, b])
            assert np.all(got[Tb:, b] == -1)
        assert got[3, 1] == 0 and got[4, 2] == 2


def test_argmax_written_for_infeasible_utterances_too(ctx):
    torch, ctc, op = ctx
    rng = np.random.default_rng(0)
    act = rng.standard_normal((4, 2, 6)).astype(np.float32)
    fl, ll, il = np.array([1, 1, 1, 2], dtype=np.int32), np.array([3, 1]), np.array([4, 4])  # utt 0 needs 5 frames
    a = torch.from_numpy(act).cuda()
    am = torch.full((8,), -7, dtype=torch.int32, device="cuda")
    op.compute_extended(a, fl, ll, il, gradients=torch.empty_like(a), argmax_dev=am)
    torch.cuda.synchronize()
    assert np.array_equal(am.cpu().numpy().reshape(4, 2), act.argmax(2))


def test_wide_alphabet_default_path_properties(ctx):
    """A slab large enough (154 MB) that the library picks the persistent TMA-ring gradient kernel and two
    utterance groups by itself: parity with the fp64 oracle, gradient rows sum to zero, padded rows are exactly
    zero, and two calls give bit-identical results (fixed summation orders everywhere)."""
    from kaldi_ctc_b200 import synth
    from oracle import pyoracle
    bt = synth.ctc_batch(16, 4000, 300, 600, 50, 280, seed=77, sigma=2.0)
    costs, grad = _run(ctx, bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths)
    costs2, grad2 = _run(ctx, bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths)
    np.testing.assert_array_equal(costs, costs2)
    np.testing.assert_array_equal(grad, grad2)
    assert np.all(np.isfinite(costs))
    assert np.abs(grad.sum(-1)).max() < 1e-4
    for b, Tb in enumerate(bt.input_lengths):
        assert np.all(grad[Tb:, b, :] == 0)
    c_ref, g_ref = pyoracle.ctc(bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths,
                                dtype=np.float64)
    _check(costs, grad, c_ref, g_ref)
