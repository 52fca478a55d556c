"""Generates the committed golden fixtures tests/golden/*.npz.

The reference holds no golden vectors for this path, and
neither warp-ctc nor cuDNN 5 can run anywhere here, so the pins are made with
INDEPENDENT fp64 implementations available in this container:
  * CTC  : torch.nn.functional.ctc_loss(log_softmax(act)) with autograd through
           log_softmax  ==  d(NLL)/d(activations), what warp-ctc returns
           (call site src/ctc/ctc-nnet-update.cc:224-231).
  * RNN  : torch.nn.LSTM / GRU / RNN (bidirectional, zero initial state, NO
           packing: padded frames are real steps, nnet-cudnn-component.cc:508-556)
           with weights scattered from the cuDNN-v5-ordered blob.
Run:  python tests/golden/make_golden.py     (CPU only, a few seconds)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from kaldi_ctc_b200 import synth  # noqa: E402


def ctc_case(name, B, A, t_lo, t_hi, l_lo, l_hi, seed, sigma=3.0, peaky=False,
             force_repeats=False):
    bt = synth.ctc_batch(B, A, t_lo, t_hi, l_lo, l_hi, seed, sigma, peaky)
    if force_repeats:  # exercise the blank-between-repeats rule
        off = 0
        for L in bt.label_lengths:
            if L >= 2:
                bt.flat_labels[off + 1] = bt.flat_labels[off]
            off += int(L)
        # keep feasible: need T >= L + repeats; generator keeps T >= 2L+1
    act = torch.tensor(bt.activations, dtype=torch.float64, requires_grad=True)
    lp = torch.log_softmax(act, dim=-1)
    loss = torch.nn.functional.ctc_loss(
        lp, torch.tensor(bt.flat_labels, dtype=torch.long),
        torch.tensor(bt.input_lengths, dtype=torch.long),
        torch.tensor(bt.label_lengths, dtype=torch.long),
        blank=0, reduction="none", zero_infinity=False)
    loss.sum().backward()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        activations=bt.activations, flat_labels=bt.flat_labels,
        label_lengths=bt.label_lengths, input_lengths=bt.input_lengths,
        costs=loss.detach().numpy(), grads=act.grad.numpy().astype(np.float64))
    print(name, bt.activations.shape, "costs", loss.detach().numpy()[:3])


def scatter_blob(mod, w, mode, bidir, layers, D, H):
    """Copy a cuDNN-v5-ordered blob into a torch.nn.{RNN,LSTM,GRU}."""
    dirs = 2 if bidir else 1
    ng = {0: 1, 1: 1, 2: 4, 3: 3}[mode]
    off = 0
    with torch.no_grad():
        for p in range(layers * dirs):
            l, d = divmod(p, dirs)
            Din = D if l == 0 else H * dirs
            sfx = "_l%d%s" % (l, "_reverse" if d else "")
            n = ng * H * Din
            getattr(mod, "weight_ih" + sfx).copy_(torch.tensor(w[off:off + n]).view(ng * H, Din))
            off += n
            n = ng * H * H
            getattr(mod, "weight_hh" + sfx).copy_(torch.tensor(w[off:off + n]).view(ng * H, H))
            off += n
        for p in range(layers * dirs):
            l, d = divmod(p, dirs)
            sfx = "_l%d%s" % (l, "_reverse" if d else "")
            n = ng * H
            getattr(mod, "bias_ih" + sfx).copy_(torch.tensor(w[off:off + n]))
            off += n
            getattr(mod, "bias_hh" + sfx).copy_(torch.tensor(w[off:off + n]))
            off += n
    assert off == w.size


def gather_grad_blob(mod, mode, bidir, layers, D, H):
    dirs = 2 if bidir else 1
    mats, biases = [], []
    for p in range(layers * dirs):
        l, d = divmod(p, dirs)
        sfx = "_l%d%s" % (l, "_reverse" if d else "")
        mats += [getattr(mod, "weight_ih" + sfx).grad.reshape(-1),
                 getattr(mod, "weight_hh" + sfx).grad.reshape(-1)]
        biases += [getattr(mod, "bias_ih" + sfx).grad, getattr(mod, "bias_hh" + sfx).grad]
    return torch.cat(mats + biases).numpy()


def rnn_case(name, mode, bidir, layers, D, H, B, T, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    dirs = 2 if bidir else 1
    n = sum(synth.blob_size(mode, bidir, D if l == 0 else H * dirs, H) for l in range(layers))
    # blob for a multi-layer component: all matrices first, then all biases
    w = (rng.standard_normal(n) * 0.3).astype(np.float32)
    x = rng.standard_normal((T * B, D)).astype(np.float32)
    x.reshape(T, B, D)[T - 2:, B - 1, :] = 0.0  # a zero-padded tail: still real steps
    dy = rng.standard_normal((T * B, H * dirs)).astype(np.float32)
    cls = {0: torch.nn.RNN, 1: torch.nn.RNN, 2: torch.nn.LSTM, 3: torch.nn.GRU}[mode]
    kw = dict(nonlinearity="relu" if mode == 0 else "tanh") if mode < 2 else {}
    mod = cls(D, H, num_layers=layers, bidirectional=bidir, **kw).double()
    scatter_blob(mod, w.astype(np.float64), mode, bidir, layers, D, H)
    xt = torch.tensor(x, dtype=torch.float64).view(T, B, D).requires_grad_(True)
    y, _ = mod(xt)
    (y * torch.tensor(dy, dtype=torch.float64).view(T, B, -1)).sum().backward()
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"), mode=mode, bidir=int(bidir), layers=layers,
        D=D, H=H, B=B, T=T, x=x, w=w, dy=dy,
        y=y.detach().numpy().reshape(T * B, -1), dx=xt.grad.numpy().reshape(T * B, D),
        dw=gather_grad_blob(mod, mode, bidir, layers, D, H))
    print(name, "y", y.shape)


if __name__ == "__main__":
    torch.manual_seed(0)
    ctc_case("ctc_small", B=4, A=6, t_lo=9, t_hi=15, l_lo=1, l_hi=4, seed=11)
    ctc_case("ctc_repeats", B=5, A=5, t_lo=12, t_hi=20, l_lo=2, l_hi=5, seed=12, force_repeats=True)
    ctc_case("ctc_ragged", B=8, A=48, t_lo=40, t_hi=120, l_lo=5, l_hi=19, seed=13)
    ctc_case("ctc_peaky", B=6, A=30, t_lo=80, t_hi=160, l_lo=20, l_hi=39, seed=14, peaky=True)
    ctc_case("ctc_wide", B=3, A=500, t_lo=30, t_hi=60, l_lo=3, l_hi=14, seed=15, sigma=2.0)
    rnn_case("rnn_lstm_bi", 2, True, 1, 7, 12, 3, 6, 21)
    rnn_case("rnn_lstm_uni2", 2, False, 2, 5, 8, 2, 5, 22)
    rnn_case("rnn_gru_bi", 3, True, 1, 6, 12, 3, 7, 23)
    rnn_case("rnn_gru_bi2", 3, True, 2, 4, 8, 2, 5, 24)
    rnn_case("rnn_relu_bi", 0, True, 1, 5, 12, 2, 6, 25)
    rnn_case("rnn_tanh_uni", 1, False, 1, 5, 8, 2, 6, 26)
    # odd hidden sizes (no divisibility by 4): exercised once the generic path exists
    rnn_case("rnn_gru_odd", 3, True, 1, 6, 10, 3, 7, 27)
    rnn_case("rnn_lstm_odd", 2, True, 1, 5, 9, 2, 6, 28)
