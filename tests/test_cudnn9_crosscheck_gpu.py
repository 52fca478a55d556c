"""Out-of-product-path pin of SURVEY 8 row R4/R5 (gate order, packed-blob layout) against REAL cuDNN: the
cuDNN 9 of this image through torch.nn.LSTM / GRU / RNN on the GPU.  The blob is scattered
into torch's parameters with model_io.filter_params_to_torch (i,f,g,o / r,z,n, two biases = cuDNN's order);
outputs, input gradients and weight gradients of the library's exact fp32 mode must agree with cuDNN's."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,bidir,layers", [(2, True, 1), (3, True, 1), (2, False, 2), (3, True, 2), (1, True, 1), (0, False, 1)])
def test_against_cudnn9(mode, bidir, layers):
    import torch
    from kaldi_ctc_b200 import model_io, rnn
    assert torch.backends.cudnn.is_available() and torch.backends.cudnn.enabled
    D, H, B, T = 24, 64, 5, 17
    dirs = 2 if bidir else 1
    comp = rnn.CuDNNRecurrentComponent()
    comp.InitFromString("learning-rate=0.0 num-layers=%d input-dim=%d output-dim=%d rnn-mode=%d bidirectional=%s "
                        "max-seq-length=32 clip-gradient=0 mini-batch=%d" % (layers, D, H, mode, "true" if bidir else "false", B))
    rng = np.random.default_rng(3)
    # (plain tanh / relu recurrences amplify round-off when the recurrent matrix is expansive: keep them contractive)
    w = (rng.standard_normal(comp.NumParameters()) * (0.2 if mode >= 2 else 0.08)).astype(np.float32)
    comp.SetParams(w)
    x = rng.standard_normal((T * B, D)).astype(np.float32)
    dy = rng.standard_normal((T * B, H * dirs)).astype(np.float32)
    xt, dyt = torch.from_numpy(x).cuda(), torch.from_numpy(dy).cuda()
    y = comp.Propagate(xt)

    class Grab:
        def Update(self, g, clip):
            self.g = g.clone()
    grab = Grab()
    dx = comp.Backprop(xt, y, dyt, to_update=grab)

    cls = {0: torch.nn.RNN, 1: torch.nn.RNN, 2: torch.nn.LSTM, 3: torch.nn.GRU}[mode]
    kw = dict(nonlinearity="relu" if mode == 0 else "tanh") if mode < 2 else {}
    net = cls(D, H, num_layers=layers, bidirectional=bidir, **kw).cuda()
    state = model_io.filter_params_to_torch(w, mode, bidir, layers, D, H)
    with torch.no_grad():
        for k, v in state.items():
            getattr(net, k).copy_(torch.from_numpy(v))
    net.flatten_parameters()
    xin = xt.view(T, B, D).clone().requires_grad_(True)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        out, _ = net(xin)                       # cuDNN 9 cudnnRNNForward
        out.backward(dyt.view(T, B, H * dirs))  # cudnnRNNBackwardData_v8 / Weights_v8
    y9 = out.detach().reshape(T * B, H * dirs).cpu().numpy()
    dx9 = xin.grad.reshape(T * B, D).cpu().numpy()
    dw9 = model_io.torch_to_filter_params({k: getattr(net, k).grad.cpu().numpy() for k in state}, mode, bidir, layers, D, H)
    assert np.abs(y.cpu().numpy() - y9).max() < 2e-5
    assert np.abs(dx.cpu().numpy() - dx9).max() < 1e-4 * max(1.0, np.abs(dx9).max())
    assert np.abs(grab.g.cpu().numpy() - dw9).max() < 1e-4 * max(1.0, np.abs(dw9).max())
