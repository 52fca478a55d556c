#!/usr/bin/env python
"""bench.py -- BLSTM+CTC training frames/s (BASELINE.json metric) on N B200s.

A "step" is one NnetCtcUpdater::ComputeForMinibatch (src/ctc/ctc-nnet-update.cc:94-127)
on one synthetic minibatch of the workload BASELINE.json quotes the metric on
(configs[1]: 5 x BLSTM-320, 40-dim input, 48 outputs, minibatch 16, T_b ~ U{1200..2000}):
forward, CTC loss + gradient, backward, weight update.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--math fp32|tensor]
  python bench.py --impl reference ...      # CPU restatement of the reference's path
Under torchrun (N > 1) every rank takes its own 16 utterances (weak scaling,
utterance-sharded as SURVEY.md 8(e)); weight gradients are summed with NCCL
all-reduce before the clip-and-update.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "blstm_ctc_train_frames_per_sec"
UNIT = "frames/s"
WORKLOAD = "configs[1]: cudnn_google 5xBLSTM-320 (D=40) + affine 640->48 + CTC, minibatch 16/GPU, T_b~U{1200..2000}, L_b~U{120..180}"


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons with nvidia-smi during the timed region."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples = index, threading.Event(), []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        mhz = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": float(self.samples[0][1]),
                "reasons": reasons, "samples": len(self.samples)}


def make_workload(spec, rank, B=16, t_lo=1200, t_hi=2000):
    from kaldi_ctc_b200 import synth
    x, fl, L, T = synth.features(B, spec.D, t_lo, t_hi, 120, 180, spec.A, seed=1002 + 97 * rank)
    return x, fl, L, T


def _best_blas_core():
    """OpenBLAS 0.3.15 (the BLAS bundled with this image) does not recognise recent CPUs and falls back to its
    SSE3 'Prescott' kernels; give the reference's CPU path the best kernels the host supports."""
    if os.environ.get("OPENBLAS_CORETYPE"):
        return os.environ["OPENBLAS_CORETYPE"]
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return None
    core = "SkylakeX" if " avx512f" in flags else ("Haswell" if " avx2" in flags else None)
    if core:
        os.environ["OPENBLAS_CORETYPE"] = core      # must be set before the library is loaded
    return core


def cpu_step_timer(spec, blobs, aw, ab, frames, steps, warmup):
    """Times the reference's CPU path on this box's host cores: the recurrent + affine layers on the reference's
    own kaldi::Matrix code (oracle/_ref/libkaldi_ref_cpu.so = /root/reference/src/{base,matrix} compiled by
    oracle/ref/Makefile, linked to OpenBLAS) and warp-ctc's OpenMP CPU CTC as restated in oracle/ctc_oracle.c
    (warp-ctc itself is not vendored in the reference).  Falls back to the plain-C port when the reference
    library was not built.  Returns (frames_per_s, ms_per_step, description dict)."""
    from kaldi_ctc_b200 import synth
    from oracle import kaldiref, pymodel, pyoracle
    pyoracle.build()
    cores = os.cpu_count() or 1
    B = 16
    x, fl, L, T = synth.features(B, spec.D, frames, frames, max(1, frames // 12), max(2, frames // 8), spec.A, seed=1002)
    core = _best_blas_core()
    if kaldiref.available():
        kaldiref.set_num_threads(cores)
        kind, blas = "reference", kaldiref.blas_info()
        step = lambda: kaldiref.train_step(spec, blobs, aw, ab, x, fl, L, T, B, num_threads=cores)
        what = ("recurrent + affine layers on the reference's kaldi::Matrix (AddMatMat -> OpenBLAS sgemm, %d BLAS threads, "
                "core type %s), CTC = OpenMP restatement of warp-ctc's CPU path (%d threads)" % (blas["threads"], core, cores))
    else:
        kind, blas = "port", None
        step = lambda: pymodel.train_step(spec, blobs, aw, ab, x, fl, L, T, B, dtype=np.float32, num_threads=cores)
        what = "plain-C OpenMP restatement (oracle/_ref not built)"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    ms = 1e3 * float(np.mean(times))
    sample = ("%d step(s) of %d utts x %d frames of the same 5xBLSTM-320+CTC training step (configs[1] has T_b<=2000: "
              "every operation of the CPU path is per frame or per time step, so frames/s does not depend on T); %s"
              % (steps, B, frames, what))
    return float(T.sum()) / (ms / 1e3), ms, {"cores": cores, "kind": kind, "sample": sample, "blas": blas}


def run_reference(args, rank, world):
    """The reference's own CPU path (see cpu_step_timer), timed on this box's host cores."""
    if rank != 0:
        return
    from kaldi_ctc_b200 import synth
    spec = synth.ModelSpec()
    blobs, aw, ab = synth.model_weights(spec, 7)
    value, ms, d = cpu_step_timer(spec, blobs, aw, ab, args.ref_frames, args.steps, args.warmup)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": d["sample"]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": d["cores"], "kind": d["kind"], "sample": d["sample"],
                         "blas": d["blas"]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def cpu_baseline(spec, blobs, aw, ab, frames=500):
    value, ms, d = cpu_step_timer(spec, blobs, aw, ab, frames, 1, 0)
    return {"value": value, "unit": UNIT, "cores": d["cores"], "kind": d["kind"], "sample": d["sample"], "blas": d["blas"],
            "ms_per_step": ms}


def cpu_ctc_figures():
    """CTC-only CPU figure: the OpenMP restatement of warp-ctc's CPU path on configs 1, 4
    and a reduced configs[5]... (B = 32), all host threads, fp32, algorithmic GB/s."""
    from kaldi_ctc_b200 import ctc as _ctc, synth
    from oracle import pyoracle
    cores = os.cpu_count() or 1
    out = {}
    for name, cfg, scale in (("configs[0/1] A=48 B=16 T<=2000", 1, 1.0), ("configs[3] A=30 B=64 T<=2000 L~400", 4, 1.0),
                             ("configs[4] reduced: A=4000 B=32 T<=3000", 5, 0.125)):
        bt = synth.config_ctc(cfg, scale)
        ts = []
        for _ in range(1 if cfg == 5 else 2):
            t0 = time.perf_counter()
            pyoracle.ctc(bt.activations, bt.flat_labels, bt.label_lengths, bt.input_lengths, dtype=np.float32, num_threads=cores)
            ts.append(time.perf_counter() - t0)
        A, B = bt.activations.shape[2], bt.activations.shape[1]
        nbytes = 4 * A * int(np.sum(bt.input_lengths)) + 4 * A * int(np.max(bt.input_lengths)) * B
        out[name] = {"ms_per_call": 1e3 * min(ts), "GBps_algorithmic": nbytes / min(ts) / 1e9, "threads": cores}
    return out


def ctc_roofline(dev, pk, pk_src, B=256):
    """Second half of BASELINE.json's metric: CTC loss+grad HBM GB/s on configs[4] (the stress sweep:
    A=4000, batch 256, T_b~U{1500..3000}, L_b~U{50..600}) on one GPU.  Algorithmic bytes of BASELINE.md
    section 3; CUDA events around one b200ctc_loss call (its three kernels); the 24.6 GB slab pair is far
    larger than L2.  Lengths/labels from the deterministic generator, activations N(0, 2^2) drawn on the
    device (12 GB of host RNG would dominate the run)."""
    import torch
    from kaldi_ctc_b200 import ctc, synth
    rng = np.random.Generator(np.random.PCG64(1005))
    T, L = synth._lengths(rng, B, 1500, 3000, 50, 600)
    labels = np.concatenate([rng.integers(1, 4000, size=int(l)) for l in L]).astype(np.int32)
    Tmax, A = int(T.max()), 4000
    g0 = torch.Generator(device=dev)
    g0.manual_seed(1005)
    a = torch.randn(Tmax, B, A, device=dev, generator=g0) * 2.0
    tmask = (torch.arange(Tmax, device=dev)[:, None] < torch.from_numpy(T.astype(np.int64)).to(dev)[None, :])
    a *= tmask[:, :, None]
    op = ctc.CtcLoss(dev)
    g = torch.empty_like(a)
    cd = torch.zeros(B, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    run = lambda: op.compute_extended(a, labels, L, T, gradients=g, costs_dev=cd, no_sync=True, nonfinite_dev=flag)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    costs = cd.cpu().numpy()
    # parity at FULL size: the fp64 oracle on 4 utterances of this very batch (CTC is independent per utterance)
    from oracle import pyoracle
    offs = np.concatenate([[0], np.cumsum(L)])
    spot = {"utterances": [], "loss_rel_max": 0.0, "grad_maxabs_max": 0.0}
    for b in [0, int(np.argmax(T)), int(np.argmax(L)), int(np.argmin(T))]:
        Tb = int(T[b])
        c_ref, g_ref = pyoracle.ctc(a[:Tb, b, :].contiguous().cpu().numpy().reshape(Tb, 1, A), labels[offs[b]:offs[b + 1]],
                                    [int(L[b])], [Tb], dtype=np.float64)
        spot["utterances"].append(b)
        spot["loss_rel_max"] = max(spot["loss_rel_max"], float(abs(costs[b] - c_ref[0]) / abs(c_ref[0])))
        spot["grad_maxabs_max"] = max(spot["grad_maxabs_max"],
                                      float(np.abs(g[:Tb, b, :].cpu().numpy() - g_ref.reshape(Tb, A)).max()))
    spot["within_tolerance"] = bool(spot["loss_rel_max"] < 1e-5 and spot["grad_maxabs_max"] < 1e-4)
    nbytes = ctc.algorithmic_bytes(L, T, A)
    ach = nbytes / ms / 1e6
    # what the three kernels must physically move: the slab is read twice (row statistics, then softmax
    # for the gradient) and written once, plus the per-utterance alpha/beta/E tables written and read once
    pitch = (L.astype(np.int64) + 1 + 3) // 4 * 4
    tables = float((T.astype(np.int64) * pitch * 5 * 2 * 4).sum())
    physical = 2.0 * 4 * A * float(T.sum()) + 4.0 * A * Tmax * B + tables
    traffic, traffic_src = ncu_traffic("ctc_grad_ring")
    return {"workload": "configs[4]: A=4000, B=%d, T_b~U{1500..3000}, L_b~U{50..600}, 1 GPU" % B, "bound": "hbm",
            "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
            "ms_per_call": ms, "algorithmic_bytes": int(nbytes), "peak_source": pk_src,
            "physical_bytes_3_passes_plus_tables": int(physical), "physical_GBps": physical / ms / 1e6,
            "physical_frac": physical / ms / 1e6 / pk["hbm_gbs"],
            "traffic": traffic, "traffic_source": (traffic_src + " (ctc_grad_ring_kernel, B=32 slice)") if traffic_src else None,
            "costs_finite": bool(np.isfinite(costs).all() and (costs > 0).all()), "nonfinite_flag": int(flag.item()),
            "parity_fp64_oracle_full_size": spot,
            "l2": "inputs (12.3 GB) + outputs (12.3 GB) >> 126 MB L2"}


def gemm_roofline(dev, pk, pk_src):
    """The hoisted input projection of layers 2-5 (32000 x 1280 x 640) timed alone with CUDA events, L2 flushed
    between launches, against dense tensor-core rates MEASURED in this run with cuBLAS (measured_gemm_peaks) and the
    driver-measured bf16 burst of MEASURED_PEAKS.json."""
    import torch
    from kaldi_ctc_b200 import rnn
    M, N, K = 32000, 1280, 640
    A = torch.randn(M, K, device=dev)
    Bm = torch.randn(N, K, device=dev)
    C = torch.empty(M, N, device=dev)
    ws = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    run = lambda: rnn.gemm(torch, 0, 1, M, N, K, 1.0, A, K, Bm, K, 0.0, C, N, math=rnn.MATH_TENSOR, workspace=ws)
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    ach = 2.0 * M * N * K / ms / 1e9
    del A, Bm, C
    gp = measured_gemm_peaks(dev)
    bf16_burst = pk.get("bf16_tflops", pk.get("bf16_tflops_sustained"))
    peak = gp["tf32"]
    traffic, src = ncu_traffic("tc_gemm_pair")
    if traffic is None:
        traffic, src = ncu_traffic("tc_gemm")
    pair = int(rnn.lib().b200rnnLastGemmUsedCtaPair())
    return {"kernel": "%s (x.Wi^T, %dx%dx%d, fp32 operands in HBM)" % ("tc_gemm_pair_kernel" if pair else "tc_gemm_kernel", M, N, K),
            "bound": "tensor", "achieved": ach,
            "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "us_per_launch": ms * 1e3,
            "peak_source": "measured in this run: cuBLAS TF32 8192^3 best of 10",
            "measured_peaks_TFLOPs": {"cublas_tf32_8192": gp["tf32"], "cublas_bf16_8192": gp["bf16"],
                                      "bf16_burst_" + pk_src: bf16_burst},
            "frac_of_bf16_burst": ach / bf16_burst, "traffic": traffic, "traffic_source": src,
            "l2": "256 MB flush write between launches"}


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the kernel, from the committed
    `ncu --set full` raw pages under profiles/ (captured once per change, never during a timed run)."""
    import csv
    import glob
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    for path in sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "*_raw.csv")), reverse=True):   # latest round first
        try:
            rows = list(csv.reader(open(path)))
            hdr, units = rows[0], rows[1]
            ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            for r in rows[2:]:
                if kernel_substr in r[ki]:
                    return (float(r[ri]) * scale[units[ri]] + float(r[wi]) * scale[units[wi]],
                            "profiles/" + os.path.basename(path))
        except (ValueError, IndexError, KeyError, OSError):
            continue
    return None, None


def measured_gemm_peaks(dev):
    """Dense tensor-core GEMM rates measured live with cuBLAS (torch.matmul, 8192^3, best of 10): the TF32 rate with
    fp32 operands and the BF16 rate -- the denominators of gemm_roofline (library calls used only as yardsticks)."""
    import torch
    n = 8192
    out = {}
    old = torch.backends.cuda.matmul.allow_tf32
    for name, dt, tf32 in (("tf32", torch.float32, True), ("bf16", torch.bfloat16, False)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        a, b = torch.randn(n, n, device=dev, dtype=dt), torch.randn(n, n, device=dev, dtype=dt)
        for _ in range(3):
            torch.matmul(a, b)
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name] = 2.0 * n ** 3 / best / 1e9
        del a, b
    torch.backends.cuda.matmul.allow_tf32 = old
    return out


def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    from kaldi_ctc_b200 import nnet, rnn, synth
    assert torch.cuda.is_available(), "bench.py needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = "cuda:%d" % local_rank
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(dev))
    math = rnn.MATH_TENSOR if args.math == "tensor" else rnn.MATH_FP32
    spec = synth.ModelSpec()
    blobs, aw, ab = synth.model_weights(spec, 7)        # same initial model on every rank
    x, fl, L, T = make_workload(spec, rank)
    B, Tmax = 16, int(T.max())
    feats = torch.from_numpy(x).pin_memory()
    valid_frames = int(T.sum())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn, steps):
        """EXACTLY `steps` calls of fn between barrier + synchronize on both sides; device time, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        barrier()
        return float(ms.item())

    # ---- objf_check: the SAME minibatch and initial model through the benchmarked arithmetic (tensor mode) and
    #      through the exact fp32 kernels, no update: a parity number in every driver run
    objf_check = None
    if rank == 0 and not args.no_objf_check:
        res = {}
        for name, m in (("tensor", rnn.MATH_TENSOR), ("fp32", rnn.MATH_FP32)):
            u = nnet.NnetCtcUpdater(spec, blobs, aw, ab, B, Tmax, device=dev, math=m, world=1)
            res[name] = u.ComputeForMinibatch(feats, Tmax, fl, L, T, update=False)
            res[name + "_logits"] = u.logits[:Tmax * B].clone()
            del u
        objf_check = {"what": "step-0 objective (sum of CTC costs) of the benchmark minibatch at T=%d: tensor mode vs exact fp32 mode" % Tmax,
                      "objf_tensor": res["tensor"], "objf_fp32": res["fp32"],
                      "rel_diff": abs(res["tensor"] - res["fp32"]) / abs(res["fp32"]),
                      "logits_maxabs_diff": float((res["tensor_logits"] - res["fp32_logits"]).abs().max())}
        del res
        torch.cuda.empty_cache()

    # self_repair_scale = 1.0 is what the recipe's ClipGradientComponent lines leave in place (InitFromString default)
    up = nnet.NnetCtcUpdater(spec, blobs, aw, ab, B, Tmax, device=dev, math=math, world=world, self_repair_scale=1.0)
    # the same minibatch as the reference holds it on the host: NnetCtcExamples with compressed frames
    egs_batch = synth.examples(B, spec.D, 1200, 2000, 120, 180, spec.A, seed=1002 + 97 * rank)

    def step(resident):
        if resident == "egs":     # ComputeForMinibatch(const std::vector<NnetCtcExample>&): GPU FormatNnetInput
            return up.ComputeForMinibatchFromExamples(egs_batch, host_sync=True)
        if resident:
            return up.ComputeForMinibatch(None, Tmax, fl, L, T, host_sync=False)
        return up.ComputeForMinibatch(feats, Tmax, fl, L, T, host_sync=True)

    def set_profiling(on):
        for c in up.rnns:
            c.plan.set_profiling(on)

    up.FormatInput(feats, Tmax)
    for _ in range(args.warmup):
        step(True)
    sampler = ClockSampler(local_rank)
    sampler.start()
    set_profiling(False)
    ms_res = timed_loop(lambda: step(True), args.steps)          # `value`: no per-launch events in the way
    sampler.stop_flag.set()
    sampler.join()
    # the same timed region once more with CUDA events around every kernel the recurrent plans launch: per-kernel
    # split for the roofline (events recorded on the launching streams, inside the timed region)
    set_profiling(True)
    up.tail_events = []
    ms_prof = timed_loop(lambda: step(True), args.steps)
    prof = [[c.plan.get_profile(k) for k in range(4)] for c in up.rnns]
    tail_ms = float(np.mean([a.elapsed_time(b) for a, b in up.tail_events])) if up.tail_events else None
    up.tail_events = None
    set_profiling(False)
    ms_e2e = timed_loop(lambda: step(False), args.steps)
    objf = up.last_objf()
    for _ in range(args.warmup):
        step("egs")
    ms_egs = timed_loop(lambda: step("egs"), args.steps)
    egs_frames = int(sum(e.NumFrames() for e in egs_batch))

    # ---- roofline of the dominant kernel (CUDA events recorded inside the timed region)
    pk, pk_src = peaks()
    cat_ms = [sum(p[k][0] for p in prof) for k in range(4)]
    cat_n = [sum(p[k][1] for p in prof) for k in range(4)]
    names = ["recurrent_forward (rec_tc_fwd_kernel)", "recurrent_backward (rec_tc_bwd_kernel)",
             "projection + dx GEMMs (tc_gemm_pair_kernel, main stream)",
             "weight-gradient GEMMs (tc_gemm_kernel, side stream: elapsed times overlap the recurrent kernels)"]
    # dominant kernel = the largest of the kernels on the step's critical path (the side-stream GEMMs hide under the
    # recurrent kernels; the ncu launch list under profiles/ gives the same ranking from serialised durations)
    k = int(np.argmax(cat_ms[:3]))
    G, H, dirs = 4, spec.H, 2
    rec_flops_per_launch = dirs * 2.0 * G * H * H * Tmax * B            # per layer, padded frames are real work
    gemm_flops_per_step = 0.0          # main-stream GEMMs: x.Wi^T for every layer, dG.Wi for layers 2..5
    for l in range(spec.layers):
        din = spec.D if l == 0 else H * dirs
        gemm_flops_per_step += dirs * 2.0 * G * H * Tmax * B * din * (2 if l > 0 else 1)
    if k < 2:
        flops_per_launch = rec_flops_per_launch
    else:
        flops_per_launch = gemm_flops_per_step * args.steps / max(cat_n[2], 1)
    avg_s = cat_ms[k] / max(cat_n[k], 1) / 1e3
    achieved = flops_per_launch / avg_s / 1e12 if avg_s > 0 else 0.0
    peak = pk.get("bf16_tflops_sustained", pk.get("bf16_tflops"))
    traffic, traffic_src = ncu_traffic(["rec_tc_fwd", "rec_tc_bwd", "tc_gemm"][k])
    kernel_ms_per_launch = cat_ms[k] / max(cat_n[k], 1)
    roofline = {"bound": "tensor", "kernel": names[k], "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes/launch (dram read+write)",
                "traffic_source": traffic_src, "peak_source": pk_src + " (bf16 sustained)",
                "share_of_step": cat_ms[k] / ms_prof, "ms_per_step_profiled_run": ms_prof / args.steps,
                "ms_per_launch": kernel_ms_per_launch, "launches_per_step": cat_n[k] / args.steps,
                "ms_per_step_by_kernel": {names[i]: cat_ms[i] / args.steps for i in range(4)},
                "note": "latency-bound per-time-step chain at minibatch 16 (DESIGN.md 4.3); the GEMM-shaped part of the "
                        "step is reported against its own peak in gemm_roofline"}

    # ---- BASELINE configs[2]: frame_subsampling_factor = 3 (T_b ~ U{400..667}), 16 utterances per GPU (weak), and the
    #      strong-scaling split of ONE 16-utterance minibatch of the headline workload over the ranks
    def side_config(updater, xs, fls, Ls, Ts, steps):
        f = torch.from_numpy(xs).pin_memory()
        Tm = int(Ts.max())
        updater.FormatInput(f, Tm)
        run = lambda: updater.ComputeForMinibatch(None, Tm, fls, Ls, Ts, host_sync=False)
        for _ in range(args.warmup):
            run()
        updater.tail_events = []
        ms = timed_loop(run, steps)
        tail = float(np.mean([a.elapsed_time(b) for a, b in updater.tail_events])) if updater.tail_events else None
        updater.tail_events = None
        fr = torch.tensor([float(Ts.sum())], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(fr)
        return {"ms_per_step": ms / steps, "frames_per_s": float(fr.item()) * steps / (ms / 1e3),
                "side_stream_wait_ms_per_step": tail}

    x3, fl3, L3, T3 = synth.features(B, spec.D, 400, 667, 120, 180, spec.A, seed=1003 + 97 * rank)
    fs3 = side_config(up, x3, fl3, L3, T3, args.steps)
    fs3.update({"workload": "configs[2]: same model after frame_subsampling_factor=3, T_b~U{400..667}, L_b~U{120..180}, "
                            "16 utts/GPU (weak scaling), frames = subsampled frames", "scaling": "weak"})
    strong = None
    if world > 1 and B % world == 0:
        xg, flg, Lg, Tg = make_workload(spec, 0)            # ONE global minibatch, utterance-sharded
        from kaldi_ctc_b200.parallel import shard_utterances
        mine = shard_utterances(B, world, rank)
        Bl = len(mine)
        offs = np.concatenate([[0], np.cumsum(Lg)])
        fll = np.concatenate([flg[offs[b]:offs[b + 1]] for b in mine])
        Ll, Tl = Lg[mine], Tg[mine]
        Tml = int(Tl.max())                                  # the local slab keeps the t*B_local + b layout
        xl = np.ascontiguousarray(xg.reshape(int(Tg.max()), B, spec.D)[:Tml, mine, :]).reshape(Tml * Bl, spec.D)
        up_s = nnet.NnetCtcUpdater(spec, blobs, aw, ab, Bl, Tml, device=dev, math=math, world=world)
        strong = side_config(up_s, xl, fll, Ll, Tl, args.steps)
        strong.update({"workload": "configs[1]'s ONE minibatch of 16 utterances sharded %d per GPU" % Bl, "scaling": "strong"})
        del up_s

    h2d = int(Tmax * B * spec.D * 4)
    d2h = int(B * 4)
    line = {
        "metric": METRIC, "value": valid_frames * world * args.steps / (ms_res / 1e3), "unit": UNIT,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if math == rnn.MATH_FP32 else "tf32/bf16 tensor-core operands, f32 accumulate+state; CTC f32",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "math": args.math, "global_batch": B * world,
                   "valid_frames_per_step": valid_frames * world, "padded_frames_per_step": Tmax * B * world,
                   "parallelism": "dp%d (utterance-sharded, NCCL all-reduce of weight gradients)" % world,
                   "l2": "working set per step (activations+reserve ~2.5 GB) >> 126 MB L2; no explicit flush"},
        "e2e": {"value": valid_frames * world * args.steps / (ms_e2e / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
        "e2e_from_examples": {"value": egs_frames * world * args.steps / (ms_egs / 1e3), "unit": UNIT,
                              "h2d_bytes_per_step": int(up.stager.h2d_bytes), "d2h_bytes_per_step": d2h,
                              "ms_per_step": ms_egs / args.steps, "valid_frames_per_step": egs_frames * world,
                              "what": "host NnetCtcExamples (CompressedMatrix frames) -> GPU decompress+format -> step"},
        "gpu_launches": up.launches_per_step() * args.steps,
        "clocks": sampler.summary(),
        "roofline": roofline,
        "objf_last_step": objf,
        "objf_check": objf_check,
        "side_stream_wait_ms_per_step": tail_ms,
        "fs3": fs3,
        "strong": strong,
    }
    if rank == 0 and world == 1 and not args.no_ctc_roofline:
        del up
        torch.cuda.empty_cache()
        # BASELINE configs[3]: CTC-character GRU, 30 outputs, L ~ 400, minibatch 64, T <= 2000
        spec_g = synth.ModelSpec(mode=3, A=30)
        bg, awg, abg = synth.model_weights(spec_g, 7)
        xg, flg, Lg, Tg = synth.features(64, spec_g.D, 1200, 2000, 350, 450, spec_g.A, seed=1004)
        up_g = nnet.NnetCtcUpdater(spec_g, bg, awg, abg, 64, int(Tg.max()), device=dev, math=math, world=1)
        c3 = side_config(up_g, xg, flg, Lg, Tg, max(2, args.steps // 2))
        c3.update({"workload": "configs[3]: 5xBiGRU-320 + CTC, 30 outputs, L_b~U{350..450}, minibatch 64, T_b~U{1200..2000}"})
        line["configs3_gru"] = c3
        del up_g
        torch.cuda.empty_cache()
        line["ctc_roofline"] = ctc_roofline(dev, pk, pk_src)
        line["gemm_roofline"] = gemm_roofline(dev, pk, pk_src)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(spec, blobs, aw, ab)
            if not args.no_ctc_roofline:
                line["ctc_roofline"]["cpu_ctc"] = cpu_ctc_figures()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--math", default=os.environ.get("B200_MATH", "tensor"), choices=["fp32", "tensor"])
    ap.add_argument("--ref-frames", type=int, default=500, help="frames per utterance of the bounded CPU sample")
    ap.add_argument("--no-objf-check", action="store_true", help="skip the tensor-vs-fp32 step-0 objective check")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ctc-roofline", action="store_true", help="skip the secondary CTC HBM-roofline measurement")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
