"""Host-side mirror of the training step of nnet2-ctc-train-simple:
kaldi::ctc::NnetCtcUpdater::ComputeForMinibatch (src/ctc/ctc-nnet-update.cc:94-127)
over the 'cudnn_google' topology that steps/ctc/nnet2/make_configs.py emits:

    [CuDNNRecurrentComponent -> ClipGradientComponent] x N -> AffineComponent -> CTC

Same order of operations as the reference: FormatInput (H2D of the time-major
slab, :107-111), SetMiniBatch (:117), Propagate (:136-169), ComputeObjfAndDeriv
(:171-259, the warp-ctc call), Backprop with the derivative negated (:320-348),
each updatable component applying w += lr * grad as it goes.  Every FLOP runs in
libb200ctc.so / libb200rnn.so; torch only owns device memory and streams.
"""
import numpy as np

from . import _lib, ctc, rnn


class AffineComponent:
    """nnet2 AffineComponent (src/nnet2/nnet-component.cc:1184-1226): y = x W^T + b."""

    def __init__(self, linear_params, bias_params, learning_rate, device="cuda:0", math=rnn.MATH_FP32):
        self.torch = _lib.require_cuda()
        t = self.torch
        self.device = t.device(device)
        self.linear_params_ = t.as_tensor(np.asarray(linear_params, np.float32)).to(self.device).clone()
        self.bias_params_ = t.as_tensor(np.asarray(bias_params, np.float32)).to(self.device).clone()
        self.learning_rate_ = learning_rate
        self.math = math
        self.ws = t.empty(32 << 20, dtype=t.uint8, device=self.device)

    def InputDim(self):
        return self.linear_params_.shape[1]

    def OutputDim(self):
        return self.linear_params_.shape[0]

    def Propagate(self, inp, out=None):
        t = self.torch
        rows, K, N = inp.shape[0], self.InputDim(), self.OutputDim()
        if out is None:
            out = t.empty(rows, N, device=self.device)
        rnn.gemm(t, 0, 1, rows, N, K, 1.0, inp, K, self.linear_params_, K, 0.0, out, N,
                 bias=self.bias_params_, math=self.math)
        return out

    def Backprop(self, in_value, out_deriv, to_update=None, in_deriv=None, grad_out=None):
        """UpdateSimple (bias += lr * colsum(dY), W += lr * dY^T X) goes through raw gradient buffers and
        Update(), so that it can be gated on the CTC call's non-finite flag, routed through the momentum
        buffers, or summed over ranks first (grad_out given: only the raw gradients are produced)."""
        t = self.torch
        rows, K, N = in_value.shape[0], self.InputDim(), self.OutputDim()
        if in_deriv is None:
            in_deriv = t.empty(rows, K, device=self.device)
        rnn.gemm(t, 0, 0, rows, K, N, 1.0, out_deriv, N, self.linear_params_, K, 0.0, in_deriv, K, math=self.math)
        if to_update is not None or grad_out is not None:
            if grad_out is None:
                if getattr(self, "gW_", None) is None:
                    self.gW_, self.gb_ = t.empty_like(self.linear_params_), t.empty_like(self.bias_params_)
                grad_out = (self.gW_, self.gb_)
            gW, gb = grad_out
            rnn.column_sums(t, out_deriv, gb, False, self.ws)
            rnn.gemm(t, 1, 0, N, K, rows, 1.0, out_deriv, N, in_value, K, 0.0, gW, K, math=self.math,
                     workspace=self.ws)
            if to_update is not None:
                to_update.Update(gW, gb)
        return in_deriv

    def Update(self, gW, gb):
        t = self.torch
        dW, db = getattr(self, "delta_", None) or (None, None)
        mom, flag = getattr(self, "momentum_", 0.0), getattr(self, "skip_flag_", None)
        rnn.update(t, self.linear_params_, gW, self.learning_rate_, 0.0, delta=dW, momentum=mom, skip_flag=flag)
        rnn.update(t, self.bias_params_, gb, self.learning_rate_, 0.0, delta=db, momentum=mom, skip_flag=flag)


class ClipGradientComponent:
    """nnet2 ClipGradientComponent, norm-based (nnet-cudnn-component.cc:912-1055): identity forward; backward scales
    each derivative row to L2 norm <= threshold, keeps the component's counters (num_clipped_, count_,
    num_self_repaired_, num_backpropped_) and adds the stochastic self-repair term (RepairGradients, :970-1055) --
    all on the device, stream-ordered (b200rnnClipGradientBackprop).  Defaults are InitFromString's (:883-910) except
    self_repair_scale, which is 0 here (the reference's default of 1.0 is what steps/ctc/nnet2/components.py:60-63
    leaves in place; pass it explicitly to get the recipe's behaviour)."""

    def __init__(self, dim, clipping_threshold=30.0, self_repair_clipped_proportion_threshold=0.01,
                 self_repair_target=0.0, self_repair_scale=0.0, device="cuda:0", seed=0):
        self.dim_, self.clipping_threshold_ = dim, clipping_threshold
        self.self_repair_clipped_proportion_threshold_ = self_repair_clipped_proportion_threshold
        self.self_repair_target_, self.self_repair_scale_ = self_repair_target, self_repair_scale
        self.torch = t = _lib.require_cuda()
        self.device = t.device(device)
        # [num_clipped_, count_, num_self_repaired_, num_backpropped_] of THIS component (what the decision reads)
        self.counters = t.zeros(4, dtype=t.int32, device=self.device)
        self.ws = None
        self.rng = np.random.default_rng(seed)   # the reference draws RandUniform() from the C library's rand()
        self.repair_probability = 0.5            # hard-coded in the reference (:979)

    def Propagate(self, inp):
        return inp  # out->CopyFromMat(in): the copy is elided, the values are identical

    def Backprop(self, out_deriv, in_value=None, to_update=None, force_attempt=None):
        """in place on out_deriv.  to_update: the ClipGradientComponent whose counters are incremented (the same
        object when the net updates itself; a copy when training through a gradient Nnet; None: no statistics and
        no self-repair, as in the reference when to_update is NULL, :960-963)."""
        if self.clipping_threshold_ <= 0:
            return out_deriv
        t = self.torch
        rows = out_deriv.shape[0]
        need = rnn.clip_gradient_workspace_bytes(rows)
        if self.ws is None or self.ws.numel() < need:
            self.ws = t.empty(int(need * 1.5) + 64, dtype=t.uint8, device=self.device)
        attempt = False
        if to_update is not None and self.self_repair_scale_ != 0.0 and in_value is not None:
            attempt = (self.rng.random() <= self.repair_probability) if force_attempt is None else bool(force_attempt)
        rnn.clip_gradient_backprop(t, out_deriv, in_value, self.clipping_threshold_,
                                   self.self_repair_clipped_proportion_threshold_, self.self_repair_target_,
                                   self.self_repair_scale_, attempt,
                                   to_update.counters if to_update is not None else None, self.counters, self.ws)
        return out_deriv

    def Info(self):
        nc, cnt, nsr, nb = [int(v) for v in self.counters.cpu()]
        return ("ClipGradientComponent, dim=%d, norm-based-clipping=true, clipping-threshold=%g, clipped-proportion=%g, "
                "num-self-repaired=%d, num-backpropped=%d" % (self.dim_, self.clipping_threshold_, nc / cnt if cnt else 0.0,
                                                              nsr, nb))

    def ZeroStats(self):
        self.counters.zero_()


class _Grab:
    """to_update stand-in that leaves the raw gradient in the component's buffer."""

    def Update(self, grad, clip):
        pass


def collapse_best_path(best_pdf, blank_id=0):
    """The 'remove blank, and uniq labels' loop of ComputeTotAccuracy (ctc-nnet-update.cc:290-302),
    quirk included: frame 0's symbol is always kept, blank or not ("at least one label")."""
    hyp = [int(best_pdf[0])]
    for j in range(1, len(best_pdf)):
        if best_pdf[j] != best_pdf[j - 1] and best_pdf[j] != blank_id:
            hyp.append(int(best_pdf[j]))
    return hyp


def levenshtein(a, b):
    """util/edit-distance-inl.h LevenshteinEditDistance: unit costs, two rolling rows."""
    prev = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        cur = [i] + [0] * len(b)
        ai = a[i - 1]
        for j in range(1, len(b) + 1):
            cur[j] = min(prev[j - 1] + (ai != b[j - 1]), prev[j] + 1, cur[j - 1] + 1)
        prev = cur
    return prev[len(b)]


def tot_accuracy(best_pdf_tb, flat_labels, label_lengths, input_lengths, blank_id=0):
    """best_pdf_tb: [T, B] ints (row t*B+b of the reference's best_pdf_cpu)."""
    tot_num, err_num, off = 0, 0, 0
    for m, (L, Tm) in enumerate(zip(label_lengths, input_lengths)):
        labels = [int(v) for v in flat_labels[off:off + L]]
        off += L
        assert blank_id not in labels
        tot_num += L
        err_num += levenshtein(labels, collapse_best_path(best_pdf_tb[:Tm, m], blank_id))
    return float(tot_num - err_num), float(tot_num)


class NnetCtcUpdater:
    """Mirror of kaldi::ctc::NnetCtcUpdater for the BLSTM/BiGRU + CTC topology."""

    def __init__(self, spec, blobs, affine_w, affine_b, minibatch, max_frames, device="cuda:0",
                 math=rnn.MATH_FP32, world=1, overlap_weights=True, momentum=0.0, self_repair_scale=0.0):
        self.torch = t = _lib.require_cuda()
        self.device = t.device(device)
        self.spec, self.B, self.math, self.world = spec, minibatch, math, world
        dirs = 2 if spec.bidir else 1
        self.rnns, self.clips = [], []
        for l, blob in enumerate(blobs):
            c = rnn.CuDNNRecurrentComponent(device, math=math)
            c.InitFromString(
                "learning-rate=%g num-layers=1 input-dim=%d output-dim=%d rnn-mode=%d bidirectional=%s "
                "max-seq-length=%d clip-gradient=%g mini-batch=%d" %
                (spec.learning_rate, spec.D if l == 0 else spec.H * dirs, spec.H, spec.mode,
                 "true" if spec.bidir else "false", max_frames, spec.clip_gradient, minibatch))
            c.SetParams(blob)
            self.rnns.append(c)
            self.clips.append(ClipGradientComponent(spec.H * dirs, spec.clipping_threshold,
                                                    self_repair_scale=self_repair_scale, device=device, seed=l))
        self.affine = AffineComponent(affine_w, affine_b, spec.learning_rate, device, math)
        self.ctc = ctc.CtcLoss(device)
        self.max_frames = max_frames
        rows = max_frames * minibatch
        # forward_data_ of the reference: one buffer per component boundary, reused every minibatch
        self.x_dev = t.empty(rows, spec.D, device=self.device)
        self.acts = [t.empty(rows, spec.H * dirs, device=self.device) for _ in blobs]
        self.logits = t.empty(rows, spec.A, device=self.device)
        self.deriv = t.empty(rows, spec.A, device=self.device)
        self.dact = [t.empty(rows, spec.H * dirs, device=self.device) for _ in range(2)]
        # the CTC call's flag word (non-finite cost / no usable posterior): every weight update of the step is
        # gated on it on the device, the asynchronous form of the reference's aborts (ctc-nnet-update.cc:232-234,254)
        self.nonfinite_dev = t.zeros(1, dtype=t.int32, device=self.device)
        self.nonfinite_host = t.zeros(1, dtype=t.int32).pin_memory()
        # momentum: the reference trains through a zeroed copy of the model, `delta_nnet`
        # (ctc-nnet-train.cc:194-202): delta += lr*grad; nnet += delta; delta *= momentum (:243-244)
        assert 0.0 <= momentum < 1.0
        self.momentum = momentum
        for c in self.rnns:
            c.skip_flag_, c.momentum_ = self.nonfinite_dev, momentum
            c.delta_ = t.zeros_like(c.filter_params_) if momentum != 0.0 else None
        self.clip_twins = [ClipGradientComponent(c.dim_, c.clipping_threshold_, device=device) for c in self.clips] \
            if momentum != 0.0 else None
        a = self.affine
        a.skip_flag_, a.momentum_ = self.nonfinite_dev, momentum
        a.delta_ = (t.zeros_like(a.linear_params_), t.zeros_like(a.bias_params_)) if momentum != 0.0 else None
        self.costs_dev = t.zeros(minibatch, device=self.device)
        self.costs_host = t.zeros(minibatch, dtype=t.float32).pin_memory()
        self.best_pdf = None       # [rows] int32, filled by the CTC pass when accuracy is wanted
        self.stager = None         # pinned/device staging of the compressed minibatch (egs.InputStager)
        self.best_pdf_host = None
        # The weight-gradient GEMMs of layer l (and its clip+update / all-reduce) do not feed layer l-1's
        # backward: they run on a side stream under the (latency-bound, 80-SM) recurrent kernel of the
        # next layer.  The step itself runs on a high-priority stream so that the recurrent kernel's
        # clusters are placed ahead of queued GEMM tiles.
        lo, hi = t.cuda.Stream.priority_range()
        self.main_stream = t.cuda.Stream(device=self.device, priority=hi)
        self.side_stream = t.cuda.Stream(device=self.device, priority=lo) if overlap_weights else None
        for c in self.rnns:
            c.side_stream = self.side_stream
        if world > 1:  # data-parallel gradient buffers of the affine layer
            self.gW = t.zeros_like(self.affine.linear_params_)
            self.gb = t.zeros_like(self.affine.bias_params_)

    def SetMiniBatch(self, B):
        for c in self.rnns:
            c.InitMiniBatch(B)

    def FormatInput(self, feats_host, T):
        """feats_host: pinned [T*B, D] float32 (FormatNnetInput's output) -> device."""
        rows = T * self.B
        self.x_dev[:rows].copy_(feats_host[:rows], non_blocking=True)

    def FormatInputFromExamples(self, examples):
        """FormatNnetInput on the GPU (b200ctc_format_input): the compressed frames are what crosses PCIe.
        Returns (T, flat_labels, label_lengths, input_lengths) of the minibatch (:351-424, :171-206)."""
        from . import egs
        if self.stager is None:
            self.stager = egs.InputStager(self.device)
        assert len(examples) == self.B
        _, T = egs.FormatNnetInput(0, 0, examples, input_mat=self.x_dev, stager=self.stager)
        assert T <= self.max_frames
        ignore = examples[0].left_context
        il = np.array([eg.NumFrames() - ignore for eg in examples], dtype=np.int32)
        ll = np.array([eg.NumLabels() for eg in examples], dtype=np.int32)
        fl = np.concatenate([np.asarray(eg.labels, dtype=np.int32) for eg in examples])
        return T, fl, ll, il

    def ComputeForMinibatchFromExamples(self, examples, update=True, host_sync=True, want_best_pdf=False):
        """NnetCtcUpdater::ComputeForMinibatch(const std::vector<NnetCtcExample>&, ...) (:94-117)."""
        t = self.torch
        cur = t.cuda.current_stream(self.device)
        self.main_stream.wait_stream(cur)
        with t.cuda.stream(self.main_stream):
            T, fl, ll, il = self.FormatInputFromExamples(examples)
        cur.wait_stream(self.main_stream)
        return self.ComputeForMinibatch(None, T, fl, ll, il, update=update, host_sync=host_sync,
                                        want_best_pdf=want_best_pdf)

    def Propagate(self, T):
        rows = T * self.B
        h = self.x_dev[:rows]
        for c, clip, out in zip(self.rnns, self.clips, self.acts):
            h = clip.Propagate(c.Propagate(h, out[:rows]))
        return self.affine.Propagate(h, self.logits[:rows])

    def ComputeObjfAndDeriv(self, T, flat_labels, label_lengths, input_lengths, sync=True, want_best_pdf=False):
        t = self.torch
        rows = T * self.B
        act = self.logits[:rows].view(T, self.B, self.spec.A)
        grad = self.deriv[:rows].view(T, self.B, self.spec.A)
        if want_best_pdf and self.best_pdf is None:
            self.best_pdf = t.empty(self.max_frames * self.B, dtype=t.int32, device=self.device)
            self.best_pdf_host = t.empty(self.max_frames * self.B, dtype=t.int32).pin_memory()
        # grad_scale=-1 fuses deriv->Scale(-1) of NnetCtcUpdater::Backprop (:323); argmax_dev fuses
        # output.FindRowMaxId of ComputeTotAccuracy (:270-273) into the pass that reads the rows anyway
        self.ctc.compute_extended(act, flat_labels, label_lengths, input_lengths, blank=0, gradients=grad,
                                  grad_scale=-1.0, costs_dev=self.costs_dev, no_sync=True,
                                  argmax_dev=self.best_pdf if want_best_pdf else None,
                                  nonfinite_dev=self.nonfinite_dev)
        self.costs_host.copy_(self.costs_dev, non_blocking=True)
        self.nonfinite_host.copy_(self.nonfinite_dev, non_blocking=True)
        if want_best_pdf:
            self.best_pdf_host[:rows].copy_(self.best_pdf[:rows], non_blocking=True)
        if sync:
            return self.last_objf()
        return None

    def Backprop(self, T, update=True):
        """world == 1: the reference's flow, every component updates itself as the
        derivative passes through it.  world > 1 (utterance-sharded data parallel,
        SURVEY 8(e)): each component's raw gradient is summed over ranks with an
        asynchronous NCCL all-reduce issued as soon as it exists (top layer first, so
        it overlaps the lower layers' backward), then clipped and applied on every rank."""
        rows = T * self.B
        n = len(self.rnns)
        t = self.torch
        top_in = self.acts[-1][:rows]
        dp = update and self.world > 1
        side = self.side_stream
        for c in self.rnns:
            c.side_stream = side
        if dp:
            from .parallel import GradientReducer
            red = GradientReducer()
            red.submit_max(self.nonfinite_dev)   # a bad minibatch on ANY rank skips the update on EVERY rank
            d = self.affine.Backprop(top_in, self.deriv[:rows], None, in_deriv=self.dact[0][:rows],
                                     grad_out=(self.gW, self.gb))
            # (queued only: this runs on the main stream, which must not wait for a transfer; the first recurrent
            #  component's submission, on the side stream, applies it)
            red.submit([self.gW, self.gb], lambda: self.affine.Update(self.gW, self.gb), drain=False)
        else:
            d = self.affine.Backprop(top_in, self.deriv[:rows], self.affine if update else None,
                                     in_deriv=self.dact[0][:rows])
        pending = None   # (component, what to run after its weight GEMMs) held back until the main stream
        #                  has queued everything up to the next recurrent kernel

        def release():
            comp_, then_ = pending
            comp_.LaunchDeferredWeights(then_)

        for l in range(n - 1, -1, -1):
            # (in_value of the clip component = the recurrent layer's output; with momentum the reference updates a
            #  copy of the net, so the component that decides never sees its own counters move: to_update is a twin)
            d = self.clips[l].Backprop(d, in_value=self.acts[l][:rows],
                                       to_update=(self.clips[l] if self.momentum == 0.0 else self.clip_twins[l]) if update else None)
            if pending is not None:
                release()
                pending = None
            inp = self.x_dev[:rows] if l == 0 else self.acts[l - 1][:rows]
            comp = self.rnns[l]
            defer = update and side is not None and l > 0
            if dp:
                then = (lambda c=comp: red.submit([c.filter_params_grad_],
                                                  lambda c=c: c.Update(c.filter_params_grad_, c.clip_gradient_)))
                d = comp.Backprop(inp, self.acts[l][:rows], d, to_update=_Grab(), want_in_deriv=(l > 0),
                                  defer_weights=defer)
                if defer:
                    pending = (comp, then)
                elif side is not None:   # bottom layer: nothing left to hide under, but keep the stream order
                    with t.cuda.stream(side):
                        then()
                else:
                    then()
            else:
                d = comp.Backprop(inp, self.acts[l][:rows], d, to_update=comp if update else None,
                                  want_in_deriv=(l > 0), defer_weights=defer)
                pending = (comp, None) if defer else None
        if pending is not None:
            release()
        if dp:
            if side is not None:
                side.wait_stream(t.cuda.current_stream(self.device))
                with t.cuda.stream(side):
                    red.finish()
            else:
                red.finish()
        if side is not None:
            cur = t.cuda.current_stream(self.device)
            ev = getattr(self, "tail_events", None)
            if ev is not None:   # measurement aid: how long the step waits for the side stream (weight GEMMs of the
                #                  bottom layer, all-reduce, updates) after the main stream has run dry
                e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
                e0.record(cur)
            cur.wait_stream(side)
            if ev is not None:
                e1.record(cur)
                ev.append((e0, e1))
        return d

    def ComputeTotAccuracy(self, T, flat_labels, label_lengths, input_lengths):
        """NnetCtcUpdater::ComputeTotAccuracy (ctc-nnet-update.cc:261-314) on the best_pdf of the
        last ComputeObjfAndDeriv(want_best_pdf=True).  Returns (tot_accuracy, tot_weight) =
        (sum |labels| - sum edit distance, sum |labels|).  The device part (row arg-max) came out of
        the CTC kernel; the per-utterance collapse + edit distance is host work, as in the reference."""
        self.torch.cuda.current_stream(self.device).synchronize()
        best = self.best_pdf_host[:T * self.B].numpy().reshape(T, self.B)
        return tot_accuracy(best, flat_labels, label_lengths, input_lengths)

    def ComputeForMinibatch(self, feats_host, T, flat_labels, label_lengths, input_lengths, update=True,
                            host_sync=True, want_best_pdf=False):
        """One training step; returns tot_objf = sum of the per-utterance NLLs (:256).
        feats_host=None keeps the slab already resident in x_dev; host_sync=False leaves the
        step asynchronous (read the objective later with last_objf())."""
        t = self.torch
        cur = t.cuda.current_stream(self.device)
        self.main_stream.wait_stream(cur)
        with t.cuda.stream(self.main_stream):
            if feats_host is not None:
                self.FormatInput(feats_host, T)
            self.Propagate(T)
            self.ComputeObjfAndDeriv(T, flat_labels, label_lengths, input_lengths, sync=False,
                                     want_best_pdf=want_best_pdf)
            self.Backprop(T, update)
        cur.wait_stream(self.main_stream)
        if not host_sync:
            return None
        return self.last_objf()

    def last_objf(self):
        """tot_objf of the last minibatch.  Raises where the reference aborts ("Error in this batch, deriv sum
        is inf/nan", ctc-nnet-update.cc:232-234; costs.Sum() != costs.Sum(), :254); the weight updates of such a
        minibatch were skipped on the device."""
        self.torch.cuda.current_stream(self.device).synchronize()
        flag = int(self.nonfinite_host[0])
        if flag:
            raise ctc.CtcError("Error in this batch: non-finite CTC cost or posterior (flag %d); "
                               "the weight update of this minibatch was skipped" % flag)
        return float(self.costs_host.sum())

    def launches_per_step(self):
        """Kernels of OUR libraries launched by one ComputeForMinibatch (fp32 mode
        bookkeeping comes from the plans; + affine 3 GEMMs (+split-K reduce) + colsum 2 + CTC 3
        + clip per layer + update per layer)."""
        n = 0
        for c in self.rnns:
            n += c.launch_counts.get("fwd", 0) + c.launch_counts.get("bwd_data", 0) + \
                c.launch_counts.get("bwd_weights", 0) + 1 + 2  # + update + ClipGradient (clip rows, finalize)
        return n + 3 + 4 + 2 + 2   # CTC 3, affine GEMMs 3 (+ split-K reduce), column sums 2, affine updates 2
