"""CPU restatement of ONE training step of nnet2-ctc-train-simple on the
'cudnn_google' topology, assembled from the C oracles (oracle/pyoracle.py) and
numpy for the stock-Kaldi neighbours.  TEST INFRASTRUCTURE ONLY.

Follows NnetCtcUpdater::ComputeForMinibatch (src/ctc/ctc-nnet-update.cc:94-127):
  Propagate   [CuDNNRecurrentComponent -> ClipGradientComponent(identity)] x N -> AffineComponent
  objective   warp-ctc NLL + d(NLL)/d(activations)                          (:171-259)
  Backprop    deriv *= -1 (:323); Affine (nnet-component.cc:1198-1226);
              ClipGradient row-norm clip (nnet-cudnn-component.cc:936-957);
              recurrent BackwardData/BackwardWeights, dW clamp +-clip, w += lr*dW (:576-614)
"""
import numpy as np

from . import pyoracle


def train_step(spec, blobs, aff_w, aff_b, x, flat_labels, label_lengths, input_lengths, B,
               dtype=np.float64, num_threads=0):
    """x [T*B, D] float32.  Returns dict(objf, costs, logits, new_blobs, new_aff_w, new_aff_b, dx0)."""
    TB = x.shape[0]
    T = TB // B
    acts = [np.asarray(x, np.float32)]
    for blob in blobs:
        y = pyoracle.rnn(spec.mode, spec.bidir, 1, spec.H, acts[-1], blob, B, dtype=dtype,
                         num_threads=num_threads)
        acts.append(y.astype(np.float32))   # the reference's buffers are fp32 CuMatrix
    top = acts[-1].astype(dtype)
    logits = top @ aff_w.astype(dtype).T + aff_b.astype(dtype)
    costs, grad = pyoracle.ctc(logits.astype(np.float32).reshape(T, B, -1), flat_labels, label_lengths,
                               input_lengths, dtype=dtype, num_threads=num_threads)
    deriv = -grad.reshape(TB, -1)
    lr = spec.learning_rate
    d = deriv @ aff_w.astype(dtype)
    new_aff_w = aff_w.astype(dtype) + lr * (deriv.T @ top)
    new_aff_b = aff_b.astype(dtype) + lr * deriv.sum(0)
    new_blobs = [None] * len(blobs)
    for l in range(len(blobs) - 1, -1, -1):
        nrm = np.sqrt((d * d).sum(1, keepdims=True))
        d = d * np.where(nrm > spec.clipping_threshold, spec.clipping_threshold / np.maximum(nrm, 1e-30), 1.0)
        _, dx, dw = pyoracle.rnn(spec.mode, spec.bidir, 1, spec.H, acts[l], blobs[l], B, dy=d, dtype=dtype,
                                 num_threads=num_threads)
        new_blobs[l] = blobs[l].astype(dtype) + lr * np.clip(dw, -spec.clip_gradient, spec.clip_gradient)
        d = dx
    return dict(objf=float(costs.sum()), costs=costs, logits=logits, new_blobs=new_blobs,
                new_aff_w=new_aff_w, new_aff_b=new_aff_b, dx0=d)
