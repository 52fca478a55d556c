"""Time b200rnnGemm (tensor mode) on the shapes of the benchmark model's step and print TFLOP/s.
L2 is flushed between timed launches (a 256 MB write) so every launch streams its operands."""
import json
import sys

import torch

sys.path.insert(0, ".")
from kaldi_ctc_b200 import rnn  # noqa: E402

SHAPES = [
    # name, tA, tB, M, N, K
    ("proj  x.Wi^T", 0, 1, 32000, 1280, 640),
    ("proj1 x.Wi^T (D=40)", 0, 1, 32000, 1280, 40),
    ("dx    dG.Wi", 0, 0, 32000, 640, 1280),
    ("dWi   dG^T.x", 1, 0, 1280, 640, 32000),
    ("dR    dG^T.h", 1, 0, 1280, 320, 31984),
    ("affine h.W^T", 0, 1, 32000, 48, 640),
    ("affine dh", 0, 0, 32000, 640, 48),
    ("affine dW", 1, 0, 48, 640, 32000),
]
ws = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
only = sys.argv[1] if len(sys.argv) > 1 else None
for name, tA, tB, M, N, K in SHAPES:
    if only and not name.startswith(only):
        continue
    A = torch.randn((K, M) if tA else (M, K), device="cuda")
    B = torch.randn((N, K) if tB else (K, N), device="cuda")
    C = torch.zeros(M, N, device="cuda")
    for _ in range(3):
        rnn.gemm(torch, tA, tB, M, N, K, 1.0, A, A.shape[1], B, B.shape[1], 0.0, C, N, math=rnn.MATH_TENSOR, workspace=ws)
    tc = rnn.lib().b200rnnLastGemmUsedTensorCores()
    n, tot = 10, 0.0
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rnn.gemm(torch, tA, tB, M, N, K, 1.0, A, A.shape[1], B, B.shape[1], 0.0, C, N, math=rnn.MATH_TENSOR, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / n
    print(json.dumps({"gemm": name, "M": M, "N": N, "K": K, "tensor": tc, "us": round(ms * 1e3, 1),
                      "TFLOPs": round(2.0 * M * N * K / ms / 1e9, 1)}))
