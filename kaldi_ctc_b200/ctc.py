"""Host-side mirror of the warp-ctc call made by NnetCtcUpdater::ComputeObjfAndDeriv
(src/ctc/ctc-nnet-update.cc:171-259): same argument meaning, same layout
(time-major [T, B, A], row t*B+b), same status-code -> exception behaviour as
WARPCTC_SAFE_CALL (:31-37).  All compute happens in libb200ctc.so.
"""
import ctypes

import numpy as np

from . import _lib

CTC_GPU = 1


class CtcOptionsUnion(ctypes.Union):
    _fields_ = [("num_threads", ctypes.c_uint), ("stream", ctypes.c_void_p)]


class CtcOptions(ctypes.Structure):  # struct ctcOptions of include/ctc.h
    _anonymous_ = ("u",)
    _fields_ = [("loc", ctypes.c_int), ("u", CtcOptionsUnion), ("blank_label", ctypes.c_int)]


class B200CtcOptions(ctypes.Structure):  # b200ctcOptions of include/b200ctc.h
    _fields_ = [("blank_label", ctypes.c_int), ("grad_scale", ctypes.c_float),
                ("stream", ctypes.c_void_p), ("no_sync", ctypes.c_int), ("argmax_dev", ctypes.c_void_p),
                ("nonfinite_dev", ctypes.c_void_p)]


class CtcError(RuntimeError):
    """Raised where the reference would KALDI_ERR on a non-zero ctcStatus_t."""


_configured = False


def lib():
    global _configured
    L = _lib.load("libb200ctc.so")
    if not _configured:
        L.ctcGetStatusString.restype = ctypes.c_char_p
        L.ctcGetStatusString.argtypes = [ctypes.c_int]
        L.get_workspace_size.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                         CtcOptions, ctypes.POINTER(ctypes.c_size_t)]
        L.compute_ctc_loss.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, CtcOptions]
        L.b200ctc_workspace_size.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                             ctypes.POINTER(ctypes.c_size_t)]
        L.b200ctc_loss.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                   ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, B200CtcOptions]
        L.b200ctc_algorithmic_bytes.restype = ctypes.c_size_t
        L.b200ctc_algorithmic_bytes.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        _configured = True
    return L


def _check(status, what):
    if status != 0:
        raise CtcError('ctcStatus_t %d : "%s" returned from \'%s\'' %
                       (status, lib().ctcGetStatusString(status).decode(), what))


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def workspace_size(label_lengths, input_lengths, alphabet_size):
    ll, il = _i32(label_lengths), _i32(input_lengths)
    n = ctypes.c_size_t(0)
    opt = CtcOptions()
    opt.loc = CTC_GPU
    opt.stream = None
    opt.blank_label = 0
    _check(lib().get_workspace_size(_ptr(ll), _ptr(il), int(alphabet_size), len(ll), opt,
                                    ctypes.byref(n)), "get_workspace_size")
    return n.value


def set_tuning(key, value):
    """Test / tuning hook (include/b200ctc.h b200ctc_set_tuning): forces one of the library's launch choices
    ("RING", "GROUPS", "P", ...) for the rest of the process; -1 / 0 restore the default of most keys."""
    L = lib()
    L.b200ctc_set_tuning.argtypes = [ctypes.c_char_p, ctypes.c_int]
    if L.b200ctc_set_tuning(key.encode(), int(value)) != 0:
        raise KeyError(key)


def algorithmic_bytes(label_lengths, input_lengths, alphabet_size):
    ll, il = _i32(label_lengths), _i32(input_lengths)
    return lib().b200ctc_algorithmic_bytes(_ptr(ll), _ptr(il), int(alphabet_size), len(ll))


class CtcLoss:
    """Re-usable caller-owned workspace + the two entry points.

    compute()          -> the warp-ctc ABI exactly as the reference calls it
    compute_extended() -> b200ctc_loss (grad_scale, device costs, no_sync)
    """

    def __init__(self, device="cuda:0"):
        self.torch = _lib.require_cuda()
        self.device = self.torch.device(device)
        self.ws = None

    def _workspace(self, nbytes):
        if self.ws is None or self.ws.numel() < nbytes:
            self.ws = self.torch.empty(int(nbytes * 1.25) + 256, dtype=self.torch.uint8, device=self.device)
        assert self.ws.data_ptr() % 256 == 0
        return self.ws

    def compute(self, activations, flat_labels, label_lengths, input_lengths, blank=0,
                want_grad=True, gradients=None):
        torch = self.torch
        assert activations.is_cuda and activations.dtype == torch.float32 and activations.is_contiguous()
        T, B, A = activations.shape
        fl, ll, il = _i32(flat_labels), _i32(label_lengths), _i32(input_lengths)
        assert len(ll) == B and len(il) == B and int(il.max()) == T
        ws = self._workspace(workspace_size(ll, il, A))
        if want_grad and gradients is None:
            gradients = torch.empty_like(activations)
        costs = np.zeros(B, dtype=np.float32)
        opt = CtcOptions()
        opt.loc = CTC_GPU
        opt.stream = torch.cuda.current_stream(self.device).cuda_stream
        opt.blank_label = blank
        with torch.cuda.device(self.device):
            st = lib().compute_ctc_loss(activations.data_ptr(), gradients.data_ptr() if want_grad else None,
                                        _ptr(fl), _ptr(ll), _ptr(il), A, B, _ptr(costs), ws.data_ptr(), opt)
        _check(st, "compute_ctc_loss")
        return costs, (gradients if want_grad else None)

    def compute_extended(self, activations, flat_labels, label_lengths, input_lengths, blank=0,
                         gradients=None, grad_scale=1.0, costs_dev=None, no_sync=False, argmax_dev=None,
                         nonfinite_dev=None):
        torch = self.torch
        assert activations.is_cuda and activations.dtype == torch.float32 and activations.is_contiguous()
        T, B, A = activations.shape
        fl, ll, il = _i32(flat_labels), _i32(label_lengths), _i32(input_lengths)
        # the library derives T from input_lengths and writes gradient rows for t < max(input_lengths) only:
        # a longer slab would leave stale rows (the reference zero-fills deriv first, ctc-nnet-update.cc:222)
        assert len(ll) == B and len(il) == B and int(il.max()) == T
        assert gradients is None or (gradients.dtype == torch.float32 and gradients.is_contiguous()
                                     and tuple(gradients.shape) == (T, B, A))
        ws = self._workspace(workspace_size(ll, il, A))
        costs = None if no_sync else np.zeros(B, dtype=np.float32)
        opt = B200CtcOptions(blank, grad_scale, torch.cuda.current_stream(self.device).cuda_stream,
                             1 if no_sync else 0, argmax_dev.data_ptr() if argmax_dev is not None else None,
                             nonfinite_dev.data_ptr() if nonfinite_dev is not None else None)
        with torch.cuda.device(self.device):
            st = lib().b200ctc_loss(activations.data_ptr(),
                                    gradients.data_ptr() if gradients is not None else None,
                                    _ptr(fl), _ptr(ll), _ptr(il), A, B,
                                    _ptr(costs) if costs is not None else None,
                                    costs_dev.data_ptr() if costs_dev is not None else None,
                                    ws.data_ptr(), ws.numel(), opt)
        _check(st, "b200ctc_loss")
        return costs
