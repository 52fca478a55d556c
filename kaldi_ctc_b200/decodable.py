"""Host-side mirror of the decoding front end: nnet2::AmNnet (the CTC topology) and
kaldi::ctc::CtcDecodableAmNnet / CtcDecodableAmNnetParallel
(src/ctc/ctc-decodable-am-nnet.{h,cc}).  The recurrent stack runs in inference mode
(no reserve space, nnet-cudnn-component.cc:534-543); everything after the affine layer
(softmax, blank-frame skipping, floor, log, prior division, scale; :54-86) is the one
b200ctc_decodable call of include/b200ctc.h.  Several utterances may be pushed through
one launch (rows t*B+u), which the reference cannot do.
"""
import ctypes

import numpy as np

from . import _lib, ctc, rnn


class CtcTransitionModel:
    """Only what the decodable needs (ctc-transition-model.h:56-62): graph label 1 is blank
    (pdf 0); label tid>1 maps to inner_tid_to_pdf[tid-1] + 1."""

    def __init__(self, inner_tid_to_pdf):
        self.map = np.asarray(inner_tid_to_pdf, dtype=np.int64)  # index 0 unused (tids are 1-based)

    def NumGraphLabels(self):
        return len(self.map)  # NumTransitionIds() + 1

    def TransitionIdToPdf(self, tid):
        assert 1 <= tid <= self.NumGraphLabels()
        return 0 if tid == 1 else int(self.map[tid - 1]) + 1


class AmNnet:
    """nnet2::AmNnet for [CuDNNRecurrent + ClipGradient]*n + Affine (+ Softmax, folded into
    b200ctc_decodable) and its Priors()."""

    def __init__(self, spec, blobs, affine_w, affine_b, priors=None, device="cuda:0", math=rnn.MATH_FP32,
                 max_frames=2000):
        from .nnet import AffineComponent
        self.torch = t = _lib.require_cuda()
        self.device = t.device(device)
        self.spec, self.math, self.max_frames = spec, math, max_frames
        dirs = 2 if spec.bidir else 1
        self.rnns = []
        for l, blob in enumerate(blobs):
            c = rnn.CuDNNRecurrentComponent(device, math=math)
            c.InitFromString(
                "learning-rate=0 num-layers=1 input-dim=%d output-dim=%d rnn-mode=%d bidirectional=%s "
                "max-seq-length=%d mini-batch=1" %
                (spec.D if l == 0 else spec.H * dirs, spec.H, spec.mode, "true" if spec.bidir else "false", max_frames))
            c.SetParams(blob)
            self.rnns.append(c)
        self.affine = AffineComponent(affine_w, affine_b, 0.0, device, math)
        self.priors_ = None if priors is None else t.as_tensor(np.asarray(priors, dtype=np.float32)).to(self.device)

    def NumPdfs(self):
        return self.spec.A

    def Priors(self):
        return self.priors_

    def Compute(self, feats_dev, minibatch):
        """NnetComputation up to (not including) the softmax: [T*B, D] -> [T*B, A] logits."""
        t = self.torch
        h = feats_dev
        for c in self.rnns:
            c.InitMiniBatch(minibatch)
            h = c.Propagate(h, inference=True)   # ClipGradientComponent is the identity forward
        out = t.empty(h.shape[0], self.spec.A, device=self.device)
        return self.affine.Propagate(h, out)


def _configure(L):
    if getattr(L, "_dec_configured", False):
        return L
    L.b200ctc_decodable_workspace_size.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                   ctypes.POINTER(ctypes.c_size_t)]
    L.b200ctc_decodable.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                    ctypes.c_void_p, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                                    ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                    ctypes.c_size_t, ctypes.c_void_p]
    L._dec_configured = True
    return L


def decodable_log_probs(torch, nnet_output, input_lengths, minibatch, priors=None, prob_scale=1.0,
                        blank_threshold=1.0, floor=1e-10, is_logits=True, workspace=None):
    """b200ctc_decodable on a device matrix; returns (log_probs [sum T, A] device, kept [B] numpy)."""
    L = _configure(ctc.lib())
    il = np.ascontiguousarray(input_lengths, dtype=np.int32)
    assert len(il) == minibatch and nnet_output.is_contiguous() and nnet_output.dtype == torch.float32
    A = nnet_output.shape[1]
    assert nnet_output.shape[0] >= int(il.max(initial=0)) * minibatch
    n = ctypes.c_size_t(0)
    ctc._check(L.b200ctc_decodable_workspace_size(il.ctypes.data, A, minibatch, ctypes.byref(n)),
               "b200ctc_decodable_workspace_size")
    if workspace is None or workspace.numel() < n.value:
        workspace = torch.empty(n.value + 256, dtype=torch.uint8, device=nnet_output.device)
    out = torch.empty(int(il.sum()), A, device=nnet_output.device)
    kept = np.zeros(minibatch, dtype=np.int32)
    with torch.cuda.device(nnet_output.device):
        st = L.b200ctc_decodable(nnet_output.data_ptr(), 1 if is_logits else 0, il.ctypes.data, A, minibatch,
                                 priors.data_ptr() if priors is not None else None, prob_scale, blank_threshold,
                                 floor, out.data_ptr(), None, kept.ctypes.data, workspace.data_ptr(),
                                 workspace.numel(), torch.cuda.current_stream(nnet_output.device).cuda_stream)
    ctc._check(st, "b200ctc_decodable")
    return out, kept


class CtcDecodableAmNnet:
    """Same constructor arguments and DecodableInterface methods as the reference's class.
    feats: host [T, D] float32 (pad_input is moot: the recurrent topology has no context)."""

    floor_ = 1e-10  # ctc-decodable-am-nnet.cc:72

    def __init__(self, trans_model, am_nnet, feats, pad_input=True, prob_scale=1.0, blank_threshold=1.0):
        self.trans_model_, self.am_nnet_ = trans_model, am_nnet
        t = am_nnet.torch
        feats = np.ascontiguousarray(feats, dtype=np.float32)
        if feats.shape[0] <= 0:
            self.log_probs_ = np.zeros((0, am_nnet.NumPdfs()), dtype=np.float32)  # KALDI_WARN, empty output (:42-47)
            return
        x = t.from_numpy(feats).to(am_nnet.device)
        logits = am_nnet.Compute(x, 1)
        lp, kept = decodable_log_probs(t, logits, [feats.shape[0]], 1, am_nnet.Priors(), prob_scale,
                                       blank_threshold, self.floor_)
        # "Transfer the log-probs to the CPU for faster access by the decoding process" (:84-86)
        self.log_probs_ = lp[:int(kept[0])].cpu().numpy()

    def LogLikelihood(self, frame, tid):
        return float(self.log_probs_[frame, self.trans_model_.TransitionIdToPdf(tid)])

    def NumFramesReady(self):
        return self.log_probs_.shape[0]

    def NumIndices(self):
        return self.am_nnet_.NumPdfs()

    def IsLastFrame(self, frame):
        assert frame < self.NumFramesReady()
        return frame == self.NumFramesReady() - 1


class CtcDecodableAmNnetParallel(CtcDecodableAmNnet):
    """Lazy variant (:89-108): no blank skipping, floor 1e-20, computed on first LogLikelihood."""

    floor_ = 1e-20

    def __init__(self, trans_model, am_nnet, feats, pad_input=True, prob_scale=1.0):
        assert feats is not None
        self.trans_model_, self.am_nnet_ = trans_model, am_nnet
        self.feats_, self.prob_scale_ = np.ascontiguousarray(feats, dtype=np.float32), prob_scale
        self.log_probs_ = None

    def Compute(self):
        CtcDecodableAmNnet.__init__(self, self.trans_model_, self.am_nnet_, self.feats_, True, self.prob_scale_, 1.0)
        self.feats_ = None

    def LogLikelihood(self, frame, tid):
        if self.feats_ is not None:
            self.Compute()
        return CtcDecodableAmNnet.LogLikelihood(self, frame, tid)

    def NumFramesReady(self):
        return self.feats_.shape[0] if self.feats_ is not None else self.log_probs_.shape[0]


def decode_batch(am_nnet, feats_list, prob_scale=1.0, blank_threshold=1.0, floor=1e-10):
    """Several utterances through ONE pass of the network and ONE b200ctc_decodable call (what SURVEY
    8(f).2 asks for).  feats_list: host [T_u, D] arrays.  Returns per-utterance host log-prob matrices.
    Shorter utterances are zero-padded to T_max like the training minibatch (FormatNnetInput), so the
    backward direction of a BLSTM sees the same padding the model was trained with."""
    t = am_nnet.torch
    B = len(feats_list)
    T = np.array([f.shape[0] for f in feats_list], dtype=np.int32)
    Tmax, D = int(T.max()), feats_list[0].shape[1]
    slab = np.zeros((Tmax, B, D), dtype=np.float32)
    for u, f in enumerate(feats_list):
        slab[:T[u], u] = f
    x = t.from_numpy(slab.reshape(Tmax * B, D)).to(am_nnet.device)
    logits = am_nnet.Compute(x, B)
    lp, kept = decodable_log_probs(t, logits, T, B, am_nnet.Priors(), prob_scale, blank_threshold, floor)
    lp = lp.cpu().numpy()
    base = np.concatenate([[0], np.cumsum(T)[:-1]])
    return [lp[base[u]:base[u] + kept[u]] for u in range(B)]
