"""Deterministic synthetic inputs for the five BASELINE.json configs
(SURVEY.md section 8(d)).  numpy only; generator = PCG64(1000 + config id).

Shapes follow the reference's conventions: time-major slabs with row index
t*B + b (src/ctc/ctc-nnet-update.cc:386-388), labels in 1..A-1 with blank 0
(:205,281-282), T >= 2L+1 (src/ctc/ctc-nnet-train.cc:84-94), zero padding past
each utterance's length.
"""
from dataclasses import dataclass

import numpy as np


@dataclass
class CtcBatch:
    activations: np.ndarray    # [T, B, A] float32, zero on padded rows
    flat_labels: np.ndarray    # [sum L] int32
    label_lengths: np.ndarray  # [B] int32
    input_lengths: np.ndarray  # [B] int32

    @property
    def shape(self):
        return self.activations.shape


def _lengths(rng, B, t_lo, t_hi, l_lo, l_hi):
    T = rng.integers(t_lo, t_hi + 1, size=B).astype(np.int32)
    T[rng.integers(0, B)] = t_hi  # force the maximum, SURVEY 8(d) table
    L = rng.integers(l_lo, l_hi + 1, size=B).astype(np.int32)
    L = np.minimum(L, (T - 1) // 2).astype(np.int32)  # keep T >= 2L+1
    return T, L


def ctc_batch(B, A, t_lo, t_hi, l_lo, l_hi, seed, sigma=3.0, peaky=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    T, L = _lengths(rng, B, t_lo, t_hi, l_lo, l_hi)
    Tmax = int(T.max())
    labels = [rng.integers(1, A, size=int(l)).astype(np.int32) for l in L]
    act = (rng.standard_normal((Tmax, B, A), dtype=np.float32) * np.float32(sigma))
    for b in range(B):
        act[T[b]:, b, :] = 0.0
        if peaky:  # late-training regime: +8 on a random monotone alignment
            Lb, Tb = int(L[b]), int(T[b])
            cuts = np.sort(rng.choice(np.arange(1, Tb), size=2 * Lb, replace=False))
            seq = np.zeros(Tb, dtype=np.int64)
            for i in range(Lb):
                seq[cuts[2 * i]:cuts[2 * i + 1]] = labels[b][i]
            act[np.arange(Tb), b, seq] += 8.0
    flat = np.concatenate(labels) if B else np.zeros(0, np.int32)
    return CtcBatch(act, flat.astype(np.int32), L, T)


def config_ctc(cfg, scale=1.0, peaky=False):
    """CTC inputs of BASELINE.json config `cfg` (1, 4 or 5); `scale` < 1 shrinks
    B (config 5) for bounded samples."""
    if cfg == 1:
        return ctc_batch(16, 48, 1200, 2000, 120, 180, 1001, 3.0, peaky)
    if cfg == 4:
        return ctc_batch(64, 30, 1200, 2000, 350, 450, 1004, 3.0, peaky)
    if cfg == 5:
        B = max(1, int(round(256 * scale)))
        return ctc_batch(B, 4000, 1500, 3000, 50, 600, 1005, 2.0, peaky)
    raise ValueError(cfg)


@dataclass
class ModelSpec:
    mode: int = 2           # 0 relu, 1 tanh, 2 lstm, 3 gru (rnn-mode)
    layers: int = 5         # stacked CuDNNRecurrentComponent's (num-layers=1 each)
    D: int = 40
    H: int = 320
    A: int = 48
    bidir: bool = True
    param_stddev: float = 0.02   # nnet-cudnn-component.cc:490
    bias_const: float = 0.2      # :402  (bias_stddev_ used as a constant)
    clip_gradient: float = 5.0   # :490
    clipping_threshold: float = 30.0  # steps/ctc/train.sh:63
    learning_rate: float = 5e-4  # egs/librispeech/ctc/run_ctc_phone.sh:32


def blob_size(mode, bidir, D, H):
    ng = {0: 1, 1: 1, 2: 4, 3: 3}[mode]
    dirs = 2 if bidir else 1
    return dirs * (ng * H * D + ng * H * H + 2 * ng * H)


def model_weights(spec: ModelSpec, seed):
    """Per-layer packed blobs (cuDNN-v5 order, see include/b200rnn.h) with the
    reference's initialisation, plus the final affine [A x 2H] and bias [A]."""
    rng = np.random.Generator(np.random.PCG64(seed))
    dirs = 2 if spec.bidir else 1
    ng = {0: 1, 1: 1, 2: 4, 3: 3}[spec.mode]
    blobs = []
    for l in range(spec.layers):
        Din = spec.D if l == 0 else spec.H * dirs
        n = blob_size(spec.mode, spec.bidir, Din, spec.H)
        nb = dirs * 2 * ng * spec.H
        w = np.empty(n, dtype=np.float32)
        w[: n - nb] = rng.standard_normal(n - nb, dtype=np.float32) * np.float32(spec.param_stddev)
        w[n - nb:] = np.float32(spec.bias_const)
        blobs.append(w)
    HO = spec.H * dirs
    aff_w = (rng.standard_normal((spec.A, HO), dtype=np.float32) / np.float32(np.sqrt(HO)))
    aff_b = np.zeros(spec.A, dtype=np.float32)
    return blobs, aff_w, aff_b


def features(B, D, t_lo, t_hi, l_lo, l_hi, A, seed):
    """Zero-padded time-major features [T*B, D] ~ N(0,1) plus labels/lengths."""
    rng = np.random.Generator(np.random.PCG64(seed))
    T, L = _lengths(rng, B, t_lo, t_hi, l_lo, l_hi)
    Tmax = int(T.max())
    x = rng.standard_normal((Tmax, B, D), dtype=np.float32)
    for b in range(B):
        x[T[b]:, b, :] = 0.0
    labels = [rng.integers(1, A, size=int(l)).astype(np.int32) for l in L]
    return x.reshape(Tmax * B, D), np.concatenate(labels).astype(np.int32), L, T


def examples(B, D, t_lo, t_hi, l_lo, l_hi, A, seed, left_context=0, spk_dim=0):
    """The same kind of minibatch as features(), as the reference stores it on disk: one
    NnetCtcExample per utterance with CompressedMatrix frames (egs.py)."""
    from . import egs
    rng = np.random.Generator(np.random.PCG64(seed))
    T, L = _lengths(rng, B, t_lo, t_hi, l_lo, l_hi)
    out = []
    for b in range(B):
        frames = rng.standard_normal((int(T[b]) + left_context, D), dtype=np.float32)
        labels = rng.integers(1, A, size=int(L[b])).astype(np.int32)
        spk = rng.standard_normal(spk_dim).astype(np.float32) if spk_dim else []
        out.append(egs.NnetCtcExample(labels, egs.CompressedMatrix.from_matrix(frames), left_context, spk))
    return out
