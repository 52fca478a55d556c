"""ctypes front-end of oracle/liboracle.so (the CPU restatement of the path).

TEST INFRASTRUCTURE ONLY: may be imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs, never by the product
package kaldi_ctc_b200.  See the headers of ctc_oracle.c / rnn_oracle.c for
what is restated and how it is pinned.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        for name in ("rnn_oracle_param_count_f32", "rnn_oracle_param_count_f64",
                     "rnn_oracle_locate_f32", "rnn_oracle_locate_f64"):
            getattr(_LIB, name).restype = ctypes.c_long
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def ctc(activations, flat_labels, label_lengths, input_lengths, blank=0,
        want_grad=True, dtype=np.float32, num_threads=0):
    """activations [T, B, A] float32 (time-major, as warp-ctc).  Returns
    (costs[B], grads[T,B,A] or None) in `dtype` (float32 | float64)."""
    act = np.ascontiguousarray(activations, dtype=np.float32)
    T, B, A = act.shape
    ll = np.ascontiguousarray(label_lengths, dtype=np.int32)
    il = np.ascontiguousarray(input_lengths, dtype=np.int32)
    fl = np.ascontiguousarray(flat_labels, dtype=np.int32)
    assert T == int(il.max()), "slab length must equal max(input_lengths)"
    costs = np.zeros(B, dtype=dtype)
    grads = np.zeros((T, B, A), dtype=dtype) if want_grad else None
    fn = lib().ctc_oracle_f32 if dtype == np.float32 else lib().ctc_oracle_f64
    st = fn(_p(act), _p(grads), _p(fl), _p(ll), _p(il), ctypes.c_int(A),
            ctypes.c_int(B), _p(costs), ctypes.c_int(blank),
            ctypes.c_int(num_threads))
    if st != 0:
        raise ValueError("ctc oracle status %d" % st)
    return costs, grads


def rnn_param_count(mode, bidir, layers, D, H):
    return int(lib().rnn_oracle_param_count_f32(mode, int(bidir), layers, D, H))


def rnn_locate(mode, bidir, layers, D, H, p, lin, is_bias):
    r, c = ctypes.c_int(), ctypes.c_int()
    off = lib().rnn_oracle_locate_f32(mode, int(bidir), layers, D, H, p, lin,
                                      int(is_bias), ctypes.byref(r),
                                      ctypes.byref(c))
    return int(off), r.value, c.value


def rnn(mode, bidir, layers, H, x, w, B, dy=None, dtype=np.float32,
        num_threads=0):
    """x [T*B, D] float32 rows t*B+b; w packed blob float32.
    Returns y [T*B, H*dirs]; with dy also (dx, dw)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    w = np.ascontiguousarray(w, dtype=np.float32)
    TB, D = x.shape
    T = TB // B
    dirs = 2 if bidir else 1
    assert w.size == rnn_param_count(mode, bidir, layers, D, H)
    y = np.zeros((TB, H * dirs), dtype=dtype)
    fn = lib().rnn_oracle_f32 if dtype == np.float32 else lib().rnn_oracle_f64
    if dy is None:
        st = fn(mode, int(bidir), layers, D, H, B, T, _p(x), _p(w), _p(y), None,
                None, None, num_threads)
        assert st == 0
        return y
    dy = np.ascontiguousarray(dy, dtype=dtype)
    dx = np.zeros((TB, D), dtype=dtype)
    dw = np.zeros(w.size, dtype=dtype)
    st = fn(mode, int(bidir), layers, D, H, B, T, _p(x), _p(w), _p(y), _p(dy),
            _p(dx), _p(dw), num_threads)
    assert st == 0
    return y, dx, dw


def tot_accuracy(output, flat_labels, label_lengths, input_lengths, minibatch, blank_id=0):
    """NnetCtcUpdater::ComputeTotAccuracy, src/ctc/ctc-nnet-update.cc:261-314, restated.

    output: [T*minibatch, A] network output, row t*minibatch+m (:283).  Returns
    (tot_accuracy, tot_weight).  FindRowMaxId (:272) = first index of the row maximum
    (np.argmax has the same tie rule); the collapse loop (:290-302) starts at i=j=1, so
    frame 0's symbol always survives, blank or not; the edit distance is
    util/edit-distance-inl.h's unit-cost Levenshtein, here as the full DP table."""
    output = np.asarray(output)
    best = output.argmax(axis=1)
    tot_num = err_num = 0
    off = 0
    for m in range(minibatch):
        L, Tm = int(label_lengths[m]), int(input_lengths[m])
        ref = [int(v) for v in flat_labels[off:off + L]]
        off += L
        tot_num += L
        hyp = [int(best[i * minibatch + m]) for i in range(Tm)]
        i = 1
        for j in range(1, Tm):
            if hyp[j] != hyp[j - 1] and hyp[j] != blank_id:
                hyp[i] = hyp[j]
                i += 1
        hyp = hyp[:i]
        d = np.zeros((len(ref) + 1, len(hyp) + 1), dtype=np.int64)
        d[:, 0] = np.arange(len(ref) + 1)
        d[0, :] = np.arange(len(hyp) + 1)
        for a in range(1, len(ref) + 1):
            for b in range(1, len(hyp) + 1):
                d[a, b] = min(d[a - 1, b - 1] + (ref[a - 1] != hyp[b - 1]), d[a - 1, b] + 1, d[a, b - 1] + 1)
        err_num += int(d[len(ref), len(hyp)])
    return float(tot_num - err_num), float(tot_num)


def decodable(nnet_output, prob_scale=1.0, blank_threshold=1.0, priors=None, floor=1e-10, is_logits=True,
              dtype=np.float64):
    """CtcDecodableAmNnet's constructor after NnetComputation, src/ctc/ctc-decodable-am-nnet.cc:54-86
    (floor=1e-20, blank_threshold=1.0 gives CtcDecodableAmNnetParallel::Compute, :89-108), for ONE
    utterance.  nnet_output: [T, A]; is_logits -> the appended SoftmaxComponent
    (nnet-component.cc SoftmaxComponent::Propagate = ApplySoftMaxPerRow) is applied first."""
    p = np.asarray(nnet_output, dtype=dtype)
    if is_logits:
        e = np.exp(p - p.max(axis=1, keepdims=True))
        p = e / e.sum(axis=1, keepdims=True)
    if blank_threshold < 1.0:                       # :54
        keep = np.nonzero(p[:, 0] < dtype(blank_threshold))[0]   # :56-59
        if len(keep) != p.shape[0] and len(keep) != 0:          # :60-68
            p = p[keep]
    lp = np.log(np.maximum(p, dtype(floor)))        # :72-73
    if priors is not None:
        lp = lp - np.log(np.asarray(priors, dtype=dtype))[None, :]   # :76-81
    return lp * dtype(prob_scale)                   # :83


def clip_gradient_backprop(deriv, in_value, threshold, counters, prop_threshold=0.01, target=0.0, scale=1.0,
                           attempt_repair=False):
    """ClipGradientComponent::Backprop + RepairGradients (src/nnet2/nnet-cudnn-component.cc:936-1055), restated in
    numpy (fp64).  counters = [num_clipped, count, num_self_repaired, num_backpropped] of the component (updated in
    place, the net updating itself); attempt_repair stands for RandUniform() <= repair_probability.
    Returns the new in_deriv."""
    d = np.array(deriv, np.float64)
    nrm2 = (d * d).sum(1) / threshold ** 2
    clipped = nrm2 > 1.0
    d[clipped] *= (nrm2[clipped] ** -0.5)[:, None]
    counters[0] += int(clipped.sum())
    counters[1] += d.shape[0]
    counters[3] += 1
    if not attempt_repair or prop_threshold >= 1.0 or scale == 0.0 or counters[1] == 0:
        return d
    prop = counters[0] / counters[1]
    if prop <= prop_threshold:
        return d
    counters[2] += 1
    v = np.asarray(in_value, np.float64)
    sign = np.where(v > 0, 1.0, -1.0)
    repair = np.maximum(np.abs(v) - target, 0.0) * sign
    dn = np.sqrt((d * d).sum(1))
    magnitude = scale * prop * dn.mean()
    rn = np.sqrt((repair * repair).sum(1))
    s2 = magnitude / rn.mean() if rn.sum() != 0 else 0.0
    d = d - (s2 / 0.5) * repair
    dn2 = np.sqrt((d * d).sum(1))
    if dn2.sum() != 0:
        d *= dn.sum() / dn2.sum()
    return d


# ---- training input path (oracle/feat_oracle.c) ---------------------------------------------------
def cm_compress(mat, force_format=0):
    """CompressedMatrix::CopyFromMat -> the in-memory image as bytes (b"" for an empty matrix)."""
    mat = np.ascontiguousarray(mat, dtype=np.float32)
    rows, cols = mat.shape
    if rows == 0 or cols == 0:
        return b""
    L = lib()
    L.feat_oracle_data_size.restype = ctypes.c_long
    L.feat_oracle_compress.restype = ctypes.c_long
    fmt = force_format or (1 if rows > 8 else 2)
    out = np.zeros(L.feat_oracle_data_size(fmt, rows, cols), dtype=np.uint8)
    n = L.feat_oracle_compress(_p(mat), rows, cols, cols, force_format, _p(out))
    assert n == out.size
    return out.tobytes()


def cm_decompress(blob):
    """CompressedMatrix::CopyToMat -> float32 [rows, cols]."""
    fmt, _, _, rows, cols = np.frombuffer(blob[:20], dtype=[("f", "<i4"), ("a", "<f4"), ("b", "<f4"),
                                                            ("r", "<i4"), ("c", "<i4")])[0]
    out = np.zeros((int(rows), int(cols)), dtype=np.float32)
    buf = np.frombuffer(blob, dtype=np.uint8)
    assert lib().feat_oracle_decompress(_p(buf), _p(out)) == 0
    return out


def format_nnet_input(blobs, spk_infos, left_context, nnet_left, nnet_right):
    """kaldi::ctc::FormatNnetInput -> (input_mat [max_frames*num_splice*B, tot_dim] float32, max_frames)."""
    B = len(blobs)
    bufs = [np.frombuffer(b, dtype=np.uint8) for b in blobs]
    ptrs = (ctypes.c_void_p * B)(*[b.ctypes.data for b in bufs])
    spk_dim = 0 if spk_infos is None else int(np.asarray(spk_infos[0]).size)
    spks = [np.ascontiguousarray(s, dtype=np.float32) for s in (spk_infos or [])]
    sptr = (ctypes.c_void_p * B)(*[s.ctypes.data for s in spks]) if spk_dim else None
    rows = max(int(np.frombuffer(b[12:16], dtype="<i4")[0]) for b in blobs)
    cols = int(np.frombuffer(blobs[0][16:20], dtype="<i4")[0]) + spk_dim
    S = 1 + nnet_left + nnet_right
    out = np.empty(rows * S * B * cols, dtype=np.float32)
    mf = lib().feat_oracle_format_nnet_input(ptrs, sptr, spk_dim, B, left_context, nnet_left, nnet_right,
                                             _p(out), ctypes.c_long(out.size))
    if mf < 0:
        raise ValueError("FormatNnetInput: bad input")
    return out[:mf * S * B * cols].reshape(mf * S * B, cols).copy(), mf
