"""libb200cudnn.so driven through the cuDNN-5 call sequence that
CuDNNRecurrentComponent::Init / Propagate / Backprop make
(src/nnet2/nnet-cudnn-component.cc:146-314, 336-408, 534-599), via ctypes."""
import ctypes
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t


def _ok(st):
    assert st == 0, "cudnnStatus_t %d" % st


@pytest.mark.parametrize("mode,math", [(2, 0), (3, 0), (2, 1)])
def test_reference_call_sequence(mode, math):
    import torch
    from oracle import pyoracle
    L = ctypes.CDLL(os.path.join(ROOT, "kaldi_ctc_b200", "libb200cudnn.so"))   # RTLD_LOCAL
    L.cudnnGetErrorString.restype = ctypes.c_char_p
    L.b200cudnnSetMath(math)
    D, H, B, T, Tmax = 24, 64, 6, 11, 16
    h = vp()
    _ok(L.cudnnCreate(ctypes.byref(h)))

    def tensor_desc(dims):
        d = vp()
        _ok(L.cudnnCreateTensorDescriptor(ctypes.byref(d)))
        strides = [int(np.prod(dims[i + 1:])) for i in range(len(dims))]
        _ok(L.cudnnSetTensorNdDescriptor(d, 0, len(dims), (ci * 3)(*dims), (ci * 3)(*strides)))
        return d
    xd = (vp * Tmax)(*[tensor_desc([B, D, 1]) for _ in range(Tmax)])
    yd = (vp * Tmax)(*[tensor_desc([B, 2 * H, 1]) for _ in range(Tmax)])
    hd = tensor_desc([2, B, H])
    drop, rnn, wd = vp(), vp(), vp()
    _ok(L.cudnnCreateDropoutDescriptor(ctypes.byref(drop)))
    n = sz()
    _ok(L.cudnnDropoutGetStatesSize(h, ctypes.byref(n)))
    assert n.value > 0 and n.value % 4 == 0
    _ok(L.cudnnSetDropoutDescriptor(drop, h, ctypes.c_float(0.0), None, n, ctypes.c_ulonglong(1337)))
    _ok(L.cudnnCreateRNNDescriptor(ctypes.byref(rnn)))
    _ok(L.cudnnSetRNNDescriptor(rnn, H, 1, drop, 0, 1, mode, 0))
    wbytes, wsbytes, rsbytes = sz(), sz(), sz()
    _ok(L.cudnnGetRNNParamsSize(h, rnn, xd[0], ctypes.byref(wbytes), 0))
    nparam = pyoracle.rnn_param_count(mode, True, 1, D, H)
    assert wbytes.value == 4 * nparam
    _ok(L.cudnnCreateFilterDescriptor(ctypes.byref(wd)))
    _ok(L.cudnnSetFilterNdDescriptor(wd, 0, 0, 3, (ci * 3)(nparam, 1, 1)))
    _ok(L.cudnnGetRNNWorkspaceSize(h, rnn, Tmax, xd, ctypes.byref(wsbytes)))
    _ok(L.cudnnGetRNNTrainingReserveSize(h, rnn, Tmax, xd, ctypes.byref(rsbytes)))
    assert wsbytes.value > 0 and rsbytes.value > 0 and wsbytes.value % 4 == 0 and rsbytes.value % 4 == 0

    rng = np.random.default_rng(mode)
    w = torch.from_numpy((rng.standard_normal(nparam) * 0.2).astype(np.float32)).cuda()
    # cudnnGetRNNLinLayer{Matrix,Bias}Params: same offsets and sizes as the oracle's blob
    nlin = {2: 8, 3: 6}[mode]
    for pl in range(2):
        for lin in range(nlin):
            for fn, is_bias in ((L.cudnnGetRNNLinLayerMatrixParams, False), (L.cudnnGetRNNLinLayerBiasParams, True)):
                fd, ptr = vp(), vp()
                _ok(L.cudnnCreateFilterDescriptor(ctypes.byref(fd)))
                _ok(fn(h, rnn, pl, xd[0], wd, vp(w.data_ptr()), lin, fd, ctypes.byref(ptr)))
                dt, fmt, nb, dims = ci(), ci(), ci(), (ci * 3)()
                _ok(L.cudnnGetFilterNdDescriptor(fd, 3, ctypes.byref(dt), ctypes.byref(fmt), ctypes.byref(nb), dims))
                off, r, c = pyoracle.rnn_locate(mode, True, 1, D, H, pl, lin, is_bias)
                assert (ptr.value - w.data_ptr()) // 4 == off and dims[0] * dims[1] * dims[2] == r * c
                _ok(L.cudnnDestroyFilterDescriptor(fd))

    x = torch.from_numpy(rng.standard_normal((T * B, D)).astype(np.float32)).cuda()
    dy = torch.from_numpy(rng.standard_normal((T * B, 2 * H)).astype(np.float32)).cuda()
    y, dx = torch.empty(T * B, 2 * H, device="cuda"), torch.empty(T * B, D, device="cuda")
    dw = torch.zeros(nparam, device="cuda")
    ws = torch.zeros(wsbytes.value // 4, device="cuda")
    rs = torch.zeros(rsbytes.value // 4, device="cuda")
    st = torch.zeros(2 * B * H, device="cuda")   # hx = cx = ... = 0, as SetBufferZero() leaves them
    P = lambda t: vp(t.data_ptr())
    torch.cuda.synchronize()
    _ok(L.cudnnRNNForwardTraining(h, rnn, T, xd, P(x), hd, P(st), hd, P(st), wd, P(w), yd, P(y), hd, P(st), hd, P(st),
                                  P(ws), wsbytes, P(rs), rsbytes))
    _ok(L.cudnnRNNBackwardData(h, rnn, T, yd, P(y), yd, P(dy), hd, P(st), hd, P(st), wd, P(w), hd, P(st), hd, P(st),
                               xd, P(dx), hd, P(st), hd, P(st), P(ws), wsbytes, P(rs), rsbytes))
    _ok(L.cudnnRNNBackwardWeights(h, rnn, T, xd, P(x), hd, P(st), yd, P(y), P(ws), wsbytes, wd, P(dw), P(rs), rsbytes))
    torch.cuda.synchronize()
    yr, dxr, dwr = pyoracle.rnn(mode, True, 1, H, x.cpu().numpy(), w.cpu().numpy(), B, dy=dy.cpu().numpy(),
                                dtype=np.float64)
    tol = (1e-5, 1e-4) if math == 0 else (5e-3, 1e-2)
    assert np.abs(y.cpu().numpy() - yr).max() < tol[0]
    assert np.abs(dx.cpu().numpy() - dxr).max() < tol[1] * max(1, np.abs(dxr).max())
    assert np.abs(dw.cpu().numpy() - dwr).max() < tol[1] * max(1, np.abs(dwr).max())
    # a longer sequence than the buffers were sized for must be refused, not overrun
    assert L.cudnnRNNForwardTraining(h, rnn, Tmax + 1, xd, P(x), hd, P(st), hd, P(st), wd, P(w), yd, P(y), hd, P(st),
                                     hd, P(st), P(ws), wsbytes, P(rs), rsbytes) != 0
    # inference entry point (mini_batch == 1 in the reference): no reserve
    y2 = torch.empty_like(y)
    _ok(L.cudnnRNNForwardInference(h, rnn, T, xd, P(x), hd, P(st), hd, P(st), wd, P(w), yd, P(y2), hd, P(st), hd, P(st),
                                   P(ws), wsbytes))
    torch.cuda.synchronize()
    assert torch.equal(y, y2)
    _ok(L.cudnnDestroyRNNDescriptor(rnn))
    _ok(L.cudnnDestroy(h))
