"""Tuning aid: times the projection / input-gradient GEMMs with the one-CTA and the CTA-pair kernel.
B200RNN_LIB=<path> loads another build of libb200rnn.so.  Usage: python tools/gemm_pair_time.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from kaldi_ctc_b200 import rnn, _lib
if os.environ.get("B200RNN_LIB"):
    _lib._cache["libb200rnn.so"] = ctypes.CDLL(os.path.abspath(os.environ["B200RNN_LIB"]))
for pr in (0, 1):
    rnn.set_tuning("GEMM_PAIR", pr)
    for (tA, tB, M, N, K) in [(0, 1, 32000, 1280, 640), (0, 0, 32000, 640, 1280), (0, 1, 32000, 1280, 40), (0, 0, 32000, 640, 2560)]:
        A = torch.randn((K, M) if tA else (M, K), device="cuda")
        B = torch.randn((N, K) if tB else (K, N), device="cuda")
        C = torch.zeros(M, N, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ts = []
        for i in range(9):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rnn.gemm(torch, tA, tB, M, N, K, 1.0, A, A.shape[1], B, B.shape[1], 0.0, C, N, math=rnn.MATH_TENSOR)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print("pair=%d %s M=%d N=%d K=%d: %.1f us  %.0f TFLOP/s (pair kernel used: %d)"
              % (pr, ("T" if tA else "N") + ("T" if tB else "N"), M, N, K, ms * 1e3, 2.0 * M * N * K / ms / 1e9,
                 rnn.lib().b200rnnLastGemmUsedCtaPair()))
