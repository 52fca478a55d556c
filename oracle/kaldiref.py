"""ctypes access to oracle/_ref/libkaldi_ref_cpu.so: the reference's OWN src/base + src/matrix compiled from
/root/reference (oracle/ref/Makefile) plus the recurrent layers written on kaldi::Matrix ops
(oracle/ref/kaldi_cpu_path.cc).  TEST / BASELINE INFRASTRUCTURE ONLY: this is "Kaldi's CPU matrix path for
the recurrent layers" that BASELINE.json asks to be timed beside the GPU path; nothing under kaldi_ctc_b200/
may import it."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
PATH = os.path.join(_HERE, "_ref", "libkaldi_ref_cpu.so")
_lib = None


def available():
    return os.path.exists(PATH)


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(PATH)
        L.kaldiref_rnn_param_count.restype = ctypes.c_size_t
        L.kaldiref_blas_config.restype = ctypes.c_char_p
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def set_num_threads(n):
    lib().kaldiref_set_num_threads(int(n))


def blas_info():
    return {"config": lib().kaldiref_blas_config().decode(), "threads": int(lib().kaldiref_get_num_threads())}


def rnn_layer(mode, bidir, H, x, w, B, dy=None):
    """One CuDNNRecurrentComponent layer on kaldi::Matrix.  x [T*B, D] fp32, w the packed blob.
    Returns y, or (y, dx, dw) when dy is given."""
    x = np.ascontiguousarray(x, np.float32)
    w = np.ascontiguousarray(w, np.float32)
    TB, D = x.shape
    T, dirs = TB // B, 2 if bidir else 1
    assert lib().kaldiref_rnn_param_count(mode, int(bidir), D, H) == w.size
    y = np.zeros((TB, H * dirs), np.float32)
    if dy is None:
        rc = lib().kaldiref_rnn_layer(mode, int(bidir), T, B, D, H, _p(x), _p(w), _p(y), None, None, None)
        assert rc == 0
        return y
    dy = np.ascontiguousarray(dy, np.float32)
    dx, dw = np.zeros_like(x), np.zeros_like(w)
    rc = lib().kaldiref_rnn_layer(mode, int(bidir), T, B, D, H, _p(x), _p(w), _p(y), _p(dy), _p(dx), _p(dw))
    assert rc == 0
    return y, dx, dw


def affine(x, W, b, deriv=None, lr=0.0):
    """AffineComponent on kaldi::Matrix: forward only (deriv None) -> out; else updates W, b IN PLACE
    (UpdateSimple) and returns in_deriv."""
    x = np.ascontiguousarray(x, np.float32)
    rows, K = x.shape
    N = W.shape[0]
    assert W.dtype == np.float32 and W.flags.c_contiguous and b.dtype == np.float32
    if deriv is None:
        out = np.empty((rows, N), np.float32)
        lib().kaldiref_affine(rows, K, N, _p(x), _p(W), _p(b), _p(out), None, None, ctypes.c_float(0.0))
        return out
    deriv = np.ascontiguousarray(deriv, np.float32)
    ind = np.empty((rows, K), np.float32)
    lib().kaldiref_affine(rows, K, N, _p(x), _p(W), _p(b), None, _p(deriv), _p(ind), ctypes.c_float(lr))
    return ind


def compress_roundtrip(m):
    m = np.ascontiguousarray(m, np.float32)
    back = np.empty_like(m)
    lib().kaldiref_compress_roundtrip(_p(m), m.shape[0], m.shape[1], _p(back))
    return back


def train_step(spec, blobs, aff_w, aff_b, x, flat_labels, label_lengths, input_lengths, B, num_threads=0):
    """One ComputeForMinibatch step (same flow as oracle/pymodel.train_step) with the recurrent and affine
    layers on kaldi::Matrix/OpenBLAS and the CTC on the OpenMP restatement of warp-ctc's CPU path
    (oracle/ctc_oracle.c, fp32).  Returns dict(objf, new_blobs, new_aff_w, new_aff_b)."""
    from . import pyoracle
    TB = x.shape[0]
    T = TB // B
    acts = [np.ascontiguousarray(x, np.float32)]
    for blob in blobs:
        acts.append(rnn_layer(spec.mode, spec.bidir, spec.H, acts[-1], blob, B))
    W, b = np.array(aff_w, np.float32, order="C"), np.array(aff_b, np.float32)
    logits = affine(acts[-1], W, b)
    costs, grad = pyoracle.ctc(logits.reshape(T, B, -1), flat_labels, label_lengths, input_lengths,
                               dtype=np.float32, num_threads=num_threads)
    deriv = np.ascontiguousarray(-grad.reshape(TB, -1), np.float32)
    lr = spec.learning_rate
    d = affine(acts[-1], W, b, deriv=deriv, lr=lr)
    new_blobs = [None] * len(blobs)
    for l in range(len(blobs) - 1, -1, -1):
        nrm = np.sqrt((d * d).sum(1, keepdims=True))
        d = (d * np.where(nrm > spec.clipping_threshold, spec.clipping_threshold / np.maximum(nrm, 1e-30), 1.0)).astype(np.float32)
        _, dx, dw = rnn_layer(spec.mode, spec.bidir, spec.H, acts[l], blobs[l], B, dy=d)
        new_blobs[l] = blobs[l] + np.float32(lr) * np.clip(dw, -spec.clip_gradient, spec.clip_gradient)
        d = dx
    return dict(objf=float(costs.sum()), costs=costs, logits=logits, new_blobs=new_blobs, new_aff_w=W, new_aff_b=b)
