"""Host-side mirror of kaldi::nnet2::CuDNNRecurrentComponent
(src/nnet2/nnet-cudnn-component.{h,cc}) over libb200rnn.so (include/b200rnn.h).

Same config keys (InitFromString, :72-98), same packed weight blob
(<FilterParams>, :673-721), same Propagate / Backprop semantics (:508-610):
hx = cx = 0, every sequence runs all T steps, dW is accumulated into a zeroed
blob, clipped element-wise to +-clip_gradient and applied as w += lr * dW.
All compute is in the CUDA library; there is no CPU path here.
"""
import ctypes
import re

import numpy as np

from . import _lib

RELU, TANH, LSTM, GRU = 0, 1, 2, 3
MATH_FP32, MATH_TENSOR = 0, 1

_configured = False


class RnnError(RuntimeError):
    """Raised where the reference would KALDI_ERR (CUDNN_SAFE_CALL, cu-common.h:55-63)."""


def lib():
    global _configured
    L = _lib.load("libb200rnn.so")
    if not _configured:
        vp, sz, i, f = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_float
        L.b200rnnGetStatusString.restype = ctypes.c_char_p
        L.b200rnnCreatePlan.argtypes = [ctypes.POINTER(vp), i, i, i, i, i, i, i, i]
        L.b200rnnDestroyPlan.argtypes = [vp]
        L.b200rnnGetParamCount.argtypes = [vp, ctypes.POINTER(sz)]
        L.b200rnnLocateParam.argtypes = [vp, i, i, i, ctypes.POINTER(sz), ctypes.POINTER(i), ctypes.POINTER(i)]
        L.b200rnnGetWorkspaceSize.argtypes = [vp, ctypes.POINTER(sz)]
        L.b200rnnGetReserveSize.argtypes = [vp, ctypes.POINTER(sz)]
        L.b200rnnForward.argtypes = [vp, i, vp, vp, vp, vp, vp, vp]
        L.b200rnnBackwardData.argtypes = [vp, i, vp, vp, vp, vp, vp, vp, vp]
        L.b200rnnBackwardWeights.argtypes = [vp, i, vp, vp, vp, vp, vp, vp]
        L.b200rnnClipAndUpdate.argtypes = [vp, vp, sz, f, f, vp]
        L.b200rnnUpdate.argtypes = [vp, vp, vp, sz, f, f, f, vp, vp]
        L.b200rnnClipRowNorm.argtypes = [vp, i, i, f, vp]
        L.b200rnnClipGradientWorkspaceSize.argtypes = [i, ctypes.POINTER(sz)]
        L.b200rnnClipGradientBackprop.argtypes = [vp, vp, i, i, f, f, f, f, i, vp, vp, vp, sz, vp]
        L.b200rnnGemm.argtypes = [i, i, i, i, i, f, vp, i, vp, i, f, vp, i, vp, i, vp, sz, vp]
        L.b200rnnColumnSums.argtypes = [vp, i, i, i, f, vp, i, vp, sz, vp]
        L.b200rnnSetProfiling.argtypes = [vp, i]
        L.b200rnnGetProfile.argtypes = [vp, i, ctypes.POINTER(f), ctypes.POINTER(i)]
        L.b200rnnForwardFlops.restype = ctypes.c_double
        L.b200rnnForwardFlops.argtypes = [vp, i]
        L.b200rnnLastLaunchCount.argtypes = [vp]
        _configured = True
    return L


def _check(status, what):
    if status != 0:
        raise RnnError("b200rnnStatus_t %d : \"%s\" returned from '%s'" %
                       (status, lib().b200rnnGetStatusString(status).decode(), what))


class Plan:
    """b200rnnPlan_t: what CuDNNRecurrentComponent::Init builds for one minibatch size."""

    def __init__(self, mode, bidirectional, num_layers, input_dim, hidden_dim, minibatch,
                 max_seq_length, math=MATH_FP32):
        self.h = ctypes.c_void_p()
        _check(lib().b200rnnCreatePlan(ctypes.byref(self.h), mode, int(bool(bidirectional)), num_layers,
                                       input_dim, hidden_dim, minibatch, max_seq_length, math),
               "b200rnnCreatePlan")
        self.mode, self.dirs, self.layers = mode, 2 if bidirectional else 1, num_layers
        self.D, self.H, self.B, self.Tmax, self.math = input_dim, hidden_dim, minibatch, max_seq_length, math

    def __del__(self):
        try:
            if self.h:
                lib().b200rnnDestroyPlan(self.h)
        except Exception:
            pass

    def _size(self, fn, name):
        n = ctypes.c_size_t()
        _check(fn(self.h, ctypes.byref(n)), name)
        return n.value

    @property
    def param_count(self):
        return self._size(lib().b200rnnGetParamCount, "b200rnnGetParamCount")

    @property
    def workspace_bytes(self):
        return self._size(lib().b200rnnGetWorkspaceSize, "b200rnnGetWorkspaceSize")

    @property
    def reserve_bytes(self):
        return self._size(lib().b200rnnGetReserveSize, "b200rnnGetReserveSize")

    def locate(self, pseudo_layer, lin_id, is_bias):
        off, r, c = ctypes.c_size_t(), ctypes.c_int(), ctypes.c_int()
        _check(lib().b200rnnLocateParam(self.h, pseudo_layer, lin_id, int(is_bias), ctypes.byref(off),
                                        ctypes.byref(r), ctypes.byref(c)), "b200rnnLocateParam")
        return off.value, r.value, c.value

    def forward_flops(self, T):
        return lib().b200rnnForwardFlops(self.h, T)

    def last_launches(self):
        return lib().b200rnnLastLaunchCount(self.h)

    def set_profiling(self, enable):
        _check(lib().b200rnnSetProfiling(self.h, int(enable)), "b200rnnSetProfiling")

    def get_profile(self, category):
        ms, n = ctypes.c_float(), ctypes.c_int()
        _check(lib().b200rnnGetProfile(self.h, category, ctypes.byref(ms), ctypes.byref(n)), "b200rnnGetProfile")
        return ms.value, n.value


def _stream(torch, device):
    return torch.cuda.current_stream(device).cuda_stream


class CuDNNRecurrentComponent:
    """Python mirror of the nnet2 component (same names; tensors are torch CUDA fp32)."""

    def __init__(self, device="cuda:0", math=MATH_FP32):
        self.torch = _lib.require_cuda()
        self.device = self.torch.device(device)
        self.math = math
        # defaults of the reference's constructor (nnet-cudnn-component.cc:486-491)
        self.learning_rate_ = 0.001
        self.input_dim_ = self.hidden_dim_ = self.num_layers_ = 0
        self.max_seq_length_ = 2000
        self.bidirectional_ = True
        self.rnn_mode_ = LSTM
        self.mini_batch_ = 0
        self.param_stddev_, self.bias_stddev_, self.clip_gradient_ = 0.02, 0.2, 5.0
        self.filter_params_ = None
        self.plan = None
        self.launch_counts = {}
        # optional torch.cuda.Stream: BackwardWeights + Update run there, overlapping whatever the
        # caller enqueues next (the next component's BackwardData); the caller joins it
        self.side_stream = None

    def Type(self):
        return "CuDNNRecurrentComponent"

    def InputDim(self):
        return self.input_dim_

    def OutputDim(self):
        return self.hidden_dim_ * (2 if self.bidirectional_ else 1)

    def InitFromString(self, args):
        kv = dict(re.findall(r"([\w-]+)=(\S+)", args))
        try:
            self.learning_rate_ = float(kv["learning-rate"])
            self.num_layers_ = int(kv["num-layers"])
            self.input_dim_ = int(kv["input-dim"])
            self.hidden_dim_ = int(kv["output-dim"])
            self.rnn_mode_ = int(kv["rnn-mode"])
            self.bidirectional_ = kv["bidirectional"].lower() in ("true", "t", "1")
            self.max_seq_length_ = int(kv["max-seq-length"])
        except KeyError:
            raise RnnError("Bad initializer " + args)
        self.param_stddev_ = float(kv.get("param-stddev", self.param_stddev_))
        self.bias_stddev_ = float(kv.get("bias-stddev", self.bias_stddev_))
        self.clip_gradient_ = float(kv.get("clip-gradient", self.clip_gradient_))
        if self.rnn_mode_ not in (0, 1, 2, 3):
            raise RnnError("rnn_mode_ = %d, should in [0, 1, 2, 3]." % self.rnn_mode_)
        self.InitMiniBatch(int(kv["mini-batch"]) if "mini-batch" in kv else 1)

    def InitMiniBatch(self, mini_batch, seq_length=0):
        """(Re)build the plan for a minibatch size; weights are initialised once
        (matrices ~ N(0, param_stddev), every bias = bias_stddev, :336-408)."""
        if self.mini_batch_ == mini_batch and seq_length == 0 and self.plan is not None:
            return
        if seq_length:
            self.max_seq_length_ = seq_length
        self.mini_batch_ = mini_batch
        self.plan = Plan(self.rnn_mode_, self.bidirectional_, self.num_layers_, self.input_dim_,
                         self.hidden_dim_, mini_batch, self.max_seq_length_, self.math)
        torch = self.torch
        n = self.plan.param_count
        if self.filter_params_ is None:
            w = torch.randn(n, device=self.device) * self.param_stddev_
            nb = self.num_layers_ * self.plan.dirs * 2 * (n_gates(self.rnn_mode_)) * self.hidden_dim_
            w[n - nb:] = self.bias_stddev_
            self.filter_params_ = w
        assert self.filter_params_.numel() == n
        self.work_space_ = torch.empty(self.plan.workspace_bytes, dtype=torch.uint8, device=self.device)
        self.reserve_space_ = torch.empty(self.plan.reserve_bytes, dtype=torch.uint8, device=self.device)
        self.filter_params_grad_ = torch.zeros(n, device=self.device)

    def SetParams(self, blob):
        self.filter_params_ = self.torch.as_tensor(np.asarray(blob, dtype=np.float32)).to(self.device).clone()

    def Propagate(self, inp, out=None, inference=False):
        torch = self.torch
        if self.mini_batch_ == 0:
            self.InitMiniBatch(1)
        assert inp.is_contiguous() and inp.dtype == torch.float32 and inp.shape[1] == self.input_dim_
        assert inp.shape[0] % self.mini_batch_ == 0
        T = inp.shape[0] // self.mini_batch_
        if T > self.max_seq_length_:
            self.InitMiniBatch(self.mini_batch_, T)
        if out is None:
            out = torch.empty(inp.shape[0], self.OutputDim(), device=self.device)
        # :534: B==1 -> cudnnRNNForwardInference (no reserve space)
        reserve = None if (self.mini_batch_ == 1 or inference) else self.reserve_space_.data_ptr()
        with torch.cuda.device(self.device):
            _check(lib().b200rnnForward(self.plan.h, T, inp.data_ptr(), self.filter_params_.data_ptr(),
                                        out.data_ptr(), self.work_space_.data_ptr(), reserve,
                                        _stream(torch, self.device)), "b200rnnForward")
        self.launch_counts["fwd"] = self.plan.last_launches()
        return out

    def Backprop(self, in_value, out_value, out_deriv, to_update=None, want_in_deriv=True, defer_weights=False):
        """defer_weights (side stream only): BackwardWeights + Update are not enqueued here but by
        LaunchDeferredWeights(), which the caller invokes once the main stream holds everything up to
        the NEXT component's recurrent kernel: the persistent GEMM CTAs must not grab the SMs before
        that kernel's clusters are placed (a 10-CTA cluster needs most of a GPC)."""
        torch = self.torch
        T = in_value.shape[0] // self.mini_batch_
        assert 0 < T <= self.max_seq_length_ and out_deriv.is_contiguous()
        in_deriv = torch.empty_like(in_value) if want_in_deriv else None
        s = _stream(torch, self.device)
        with torch.cuda.device(self.device):
            _check(lib().b200rnnBackwardData(self.plan.h, T, out_value.data_ptr(), out_deriv.data_ptr(),
                                             self.filter_params_.data_ptr(),
                                             in_deriv.data_ptr() if want_in_deriv else None,
                                             self.work_space_.data_ptr(), self.reserve_space_.data_ptr(), s),
                   "b200rnnBackwardData")
            self.launch_counts["bwd_data"] = self.plan.last_launches()
            if to_update is not None:
                if self.side_stream is not None and defer_weights:
                    self._deferred = (T, in_value, out_value, to_update)
                elif self.side_stream is not None:
                    self.side_stream.wait_stream(torch.cuda.current_stream(self.device))
                    with torch.cuda.stream(self.side_stream):
                        self._backward_weights(T, in_value, out_value, to_update)
                else:
                    self._backward_weights(T, in_value, out_value, to_update)
        return in_deriv

    def LaunchDeferredWeights(self, then=None):
        """then: optional callable run on the side stream right after (the data-parallel all-reduce)."""
        if getattr(self, "_deferred", None) is None:
            return
        torch = self.torch
        args, self._deferred = self._deferred, None
        with torch.cuda.device(self.device):
            self.side_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.side_stream):
                self._backward_weights(*args)
                if then is not None:
                    then()

    def _backward_weights(self, T, in_value, out_value, to_update):
        torch = self.torch
        self.filter_params_grad_.zero_()
        _check(lib().b200rnnBackwardWeights(self.plan.h, T, in_value.data_ptr(), out_value.data_ptr(),
                                            self.filter_params_grad_.data_ptr(), self.work_space_.data_ptr(),
                                            self.reserve_space_.data_ptr(), _stream(torch, self.device)),
               "b200rnnBackwardWeights")
        self.launch_counts["bwd_weights"] = self.plan.last_launches()
        to_update.Update(self.filter_params_grad_, self.clip_gradient_)

    def Update(self, filter_params_grad, clip):
        """filter_params_ += learning_rate_ * clamp(grad) (:602-603, 612-614).  With `delta_` set (the
        reference's gradient Nnet of TrainNnetSimple, ctc-nnet-train.cc:194-202) the step goes through it with
        `momentum_`; `skip_flag_` (device int, the CTC call's non-finite flag) turns the update into a no-op."""
        torch = self.torch
        with torch.cuda.device(self.device):
            update(torch, self.filter_params_, filter_params_grad, self.learning_rate_, clip,
                   delta=getattr(self, "delta_", None), momentum=getattr(self, "momentum_", 0.0),
                   skip_flag=getattr(self, "skip_flag_", None))

    # the remaining UpdatableComponent surface acts on the flat blob (:723-772)
    def NumParameters(self):
        return int(self.filter_params_.numel())

    def Vectorize(self):
        return self.filter_params_.detach().cpu().numpy().copy()

    def UnVectorize(self, params):
        self.SetParams(params)

    def Scale(self, s):
        self.filter_params_ *= s

    def Add(self, alpha, other):
        self.filter_params_ += alpha * other.filter_params_

    def DotProduct(self, other):
        return float((self.filter_params_ * other.filter_params_).sum())


def n_gates(mode):
    return {0: 1, 1: 1, 2: 4, 3: 3}[mode]


def gemm(torch, transA, transB, M, N, K, alpha, A, lda, B, ldb, beta, C, ldc, bias=None, math=MATH_FP32,
         workspace=None):
    _check(lib().b200rnnGemm(int(transA), int(transB), M, N, K, alpha, A.data_ptr(), lda, B.data_ptr(), ldb,
                             beta, C.data_ptr(), ldc, bias.data_ptr() if bias is not None else None, math,
                             workspace.data_ptr() if workspace is not None else None,
                             workspace.numel() * workspace.element_size() if workspace is not None else 0,
                             _stream(torch, C.device)), "b200rnnGemm")


def set_tuning(key, value):
    """Test / tuning hook (include/b200rnn.h b200rnnSetTuning), e.g. ("GEMM_PAIR", -1 | 0 | 1)."""
    L = lib()
    L.b200rnnSetTuning.argtypes = [ctypes.c_char_p, ctypes.c_int]
    if L.b200rnnSetTuning(key.encode(), int(value)) != 0:
        raise KeyError(key)


def column_sums(torch, a, out, accumulate, workspace, alpha=1.0):
    rows, cols = a.shape
    _check(lib().b200rnnColumnSums(a.data_ptr(), rows, cols, a.stride(0), alpha, out.data_ptr(), int(accumulate),
                                   workspace.data_ptr(), workspace.numel() * workspace.element_size(),
                                   _stream(torch, a.device)), "b200rnnColumnSums")


def column_sums_scaled(torch, a, out, alpha, workspace):
    """out += alpha * colsum(a)"""
    column_sums(torch, a, out, True, workspace, alpha=alpha)


def clip_and_update(torch, w, dw, lr, clip):
    _check(lib().b200rnnClipAndUpdate(w.data_ptr(), dw.data_ptr(), w.numel(), lr, clip,
                                      _stream(torch, w.device)), "b200rnnClipAndUpdate")


def update(torch, w, dw, lr, clip, delta=None, momentum=0.0, skip_flag=None):
    """b200rnnUpdate: delta None -> w += lr*clamp(dw); else delta += lr*clamp(dw); w += delta; delta *= momentum."""
    _check(lib().b200rnnUpdate(w.data_ptr(), delta.data_ptr() if delta is not None else None, dw.data_ptr(),
                               w.numel(), lr, clip, momentum,
                               skip_flag.data_ptr() if skip_flag is not None else None,
                               _stream(torch, w.device)), "b200rnnUpdate")


def clip_gradient_workspace_bytes(rows):
    n = ctypes.c_size_t()
    _check(lib().b200rnnClipGradientWorkspaceSize(int(rows), ctypes.byref(n)), "b200rnnClipGradientWorkspaceSize")
    return n.value


def clip_gradient_backprop(torch, deriv, in_value, threshold, prop_threshold, target, scale, attempt_repair,
                           counters, decide_counters, workspace):
    """b200rnnClipGradientBackprop: ClipGradientComponent::Backprop incl. counters and self-repair, in place."""
    rows, cols = deriv.shape
    _check(lib().b200rnnClipGradientBackprop(
        deriv.data_ptr(), in_value.data_ptr() if in_value is not None else None, rows, cols, threshold,
        prop_threshold, target, scale, int(bool(attempt_repair)),
        counters.data_ptr() if counters is not None else None,
        decide_counters.data_ptr() if decide_counters is not None else None,
        workspace.data_ptr(), workspace.numel() * workspace.element_size(), _stream(torch, deriv.device)),
        "b200rnnClipGradientBackprop")


def clip_row_norm(torch, d, threshold):
    _check(lib().b200rnnClipRowNorm(d.data_ptr(), d.shape[0], d.shape[1], threshold,
                                    _stream(torch, d.device)), "b200rnnClipRowNorm")
