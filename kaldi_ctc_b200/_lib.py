"""Loads the in-tree product libraries (C ABI of include/*.h) with ctypes.

No fallback: a missing library raises, naming the build command.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_cache = {}


class LibraryMissing(RuntimeError):
    pass


def path(name):
    return os.path.join(_HERE, name)


def load(name):
    """name: 'libb200ctc.so' | 'libb200rnn.so'."""
    if name not in _cache:
        p = path(name)
        if not os.path.exists(p):
            raise LibraryMissing(
                "%s not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or make -C kaldi_ctc_b200/csrc). There is no CPU fallback." % p)
        _cache[name] = ctypes.CDLL(p)  # RTLD_LOCAL: libb200cudnn.so exports cuDNN names, keep them private
    return _cache[name]


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("kaldi_ctc_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch
