"""The TMA-fed tcgen05 (kind::tf32) GEMM of libb200rnn.so against numpy fp64, for
every operand-major combination the recurrent layers use.  TF32 keeps 10 mantissa
bits of each operand (fp32 accumulation): tolerance 2e-3 of sum|a||b|."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES = [
    # tA tB  M     N    K      (what it is in the model)
    (0, 1, 512, 1280, 40),     # layer-1 input projection   x . Wi^T
    (0, 1, 777, 1280, 640),    # layers 2-5 projection, ragged M
    (0, 0, 640, 640, 1280),    # dx = dG . Wi
    (1, 0, 1280, 640, 4096),   # dWi = dG^T . x   (split-K)
    (1, 0, 1280, 320, 3000),   # dR  = dG^T . h_prev
    (0, 1, 1000, 48, 640),     # affine forward, N = 48
    (1, 0, 48, 640, 5000),     # affine weight gradient
    (1, 1, 200, 136, 96),      # remaining combination
    (0, 1, 130, 260, 33 * 4),  # K not a multiple of the 32-wide k-block
    (0, 1, 2100, 1280, 640),   # several 256-row tiles, ragged last pair (M % 256 = 52: the peer CTA's rows are all padding)
    (0, 0, 2500, 640, 1280),   # dx at a size where the CTA-pair kernel is the default
]


@pytest.mark.parametrize("pair", [0, 1])
@pytest.mark.parametrize("tA,tB,M,N,K", CASES)
def test_tc_gemm(tA, tB, M, N, K, pair):
    """pair = 1 forces the CTA-pair kernel (tcgen05 cta_group::2) wherever its tile shape applies (N > 128, no
    split-K); pair = 0 the one-CTA kernel.  Same tolerance: both are TF32 with fp32 accumulation."""
    import torch
    from kaldi_ctc_b200 import rnn
    rnn.set_tuning("GEMM_PAIR", pair)
    try:
        _run_case(torch, rnn, tA, tB, M, N, K, pair)
    finally:
        rnn.set_tuning("GEMM_PAIR", -1)


def _run_case(torch, rnn, tA, tB, M, N, K, pair):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
    Bm = rng.standard_normal((N, K) if tB else (K, N)).astype(np.float32)
    C0 = rng.standard_normal((M, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    opA, opB = (A.T if tA else A).astype(np.float64), (Bm.T if tB else Bm).astype(np.float64)
    want = 0.5 * (opA @ opB) + 2.0 * C0 + bias
    bound = 0.5 * (np.abs(opA) @ np.abs(opB))
    At, Bt, Ct, bt = (torch.from_numpy(v).cuda() for v in (A, Bm, C0.copy(), bias))
    ws = torch.empty(128 << 20, dtype=torch.uint8, device="cuda")
    rnn.gemm(torch, tA, tB, M, N, K, 0.5, At, A.shape[1], Bt, Bm.shape[1], 2.0, Ct, N, bias=bt,
             math=rnn.MATH_TENSOR, workspace=ws)
    torch.cuda.synchronize()
    assert rnn.lib().b200rnnLastGemmUsedTensorCores() == 1, "fell back to the fp32 path"
    used_pair = rnn.lib().b200rnnLastGemmUsedCtaPair()
    if pair == 0:
        assert used_pair == 0
    elif (M, N, K) == (2100, 1280, 640):   # 256-wide tiles, no split-K
        assert used_pair == 1, "CTA-pair kernel was not used"
    err = np.abs(Ct.cpu().numpy() - want)
    assert (err <= 2e-3 * bound + 1e-4).all(), "max err %g (bound %g)" % (err.max(), (2e-3 * bound).max())
    # and it is genuinely TF32 (not fp32): some rounding must be visible at K >= 640
    if K >= 640:
        assert err.max() > 1e-6


@pytest.mark.parametrize("tma_store", [1, 0])
@pytest.mark.parametrize("tA,tB", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(2100, 640, 300), (256, 1280, 64), (3000, 260, 1000)])
def test_cta_pair_kernel_every_operand_major(tA, tB, M, N, K, tma_store):
    """The CTA-pair kernel (tcgen05.mma.cta_group::2, TMA loads counted on the leader's barrier, multicast
    commits) forced for all four operand-major combinations: no split-K workspace, so every case runs on it.
    Ragged M (the peer CTA's rows partly or wholly padding), ragged N, K not a multiple of 32; output through
    TMA tile stores (the default when beta = 0) and through the register path."""
    import torch
    from kaldi_ctc_b200 import rnn
    rng = np.random.default_rng(M + 3 * N + 7 * K + tA + 2 * tB)
    A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
    Bm = rng.standard_normal((N, K) if tB else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32)
    opA, opB = (A.T if tA else A).astype(np.float64), (Bm.T if tB else Bm).astype(np.float64)
    want = opA @ opB + bias
    bound = np.abs(opA) @ np.abs(opB)
    At, Bt, bt = (torch.from_numpy(v).cuda() for v in (A, Bm, bias))
    Ct = torch.full((M, N), 7.0, device="cuda")
    rnn.set_tuning("GEMM_PAIR", 1)
    rnn.set_tuning("GEMM_TMA_STORE", tma_store)
    try:
        rnn.gemm(torch, tA, tB, M, N, K, 1.0, At, A.shape[1], Bt, Bm.shape[1], 0.0, Ct, N, bias=bt, math=rnn.MATH_TENSOR)
        torch.cuda.synchronize()
        assert rnn.lib().b200rnnLastGemmUsedTensorCores() == 1
        assert rnn.lib().b200rnnLastGemmUsedCtaPair() == 1
    finally:
        rnn.set_tuning("GEMM_PAIR", -1)
        rnn.set_tuning("GEMM_TMA_STORE", 1)
    err = np.abs(Ct.cpu().numpy() - want)
    assert (err <= 2e-3 * bound + 1e-4).all(), "max err %g" % err.max()


@pytest.mark.parametrize("seed", range(10))
def test_cta_pair_kernel_random_shapes(seed):
    """Seeded random shapes through the forced CTA-pair kernel: M from below one CTA's 128 rows to several 256-row
    tiles, N from 132 to 900 (column tiles of every width the MN-major path produces: 64, 128, 192, 256), K from one
    partial k-block to many, alpha / bias present or not."""
    import torch
    from kaldi_ctc_b200 import rnn
    rng = np.random.default_rng(500 + seed)
    tA, tB = int(rng.integers(0, 2)), int(rng.integers(0, 2))
    M = int(rng.integers(1, 1200)) * (4 if tA else 1)     # MN-major A needs a row pitch of whole 16-byte chunks
    N = int(rng.integers(33, 226)) * 4
    K = int(rng.integers(1, 300)) * 4
    alpha = float(rng.choice([1.0, -0.75]))
    A = rng.standard_normal((K, M) if tA else (M, K)).astype(np.float32)
    Bm = rng.standard_normal((N, K) if tB else (K, N)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32) if seed % 2 else None
    opA, opB = (A.T if tA else A).astype(np.float64), (Bm.T if tB else Bm).astype(np.float64)
    want = alpha * (opA @ opB) + (bias if bias is not None else 0.0)
    bound = abs(alpha) * (np.abs(opA) @ np.abs(opB))
    At, Bt = torch.from_numpy(A).cuda(), torch.from_numpy(Bm).cuda()
    bt = torch.from_numpy(bias).cuda() if bias is not None else None
    Ct = torch.full((M, N), -3.0, device="cuda")
    rnn.set_tuning("GEMM_PAIR", 1)
    try:
        rnn.gemm(torch, tA, tB, M, N, K, alpha, At, A.shape[1], Bt, Bm.shape[1], 0.0, Ct, N, bias=bt, math=rnn.MATH_TENSOR)
        torch.cuda.synchronize()
        assert rnn.lib().b200rnnLastGemmUsedCtaPair() == 1
    finally:
        rnn.set_tuning("GEMM_PAIR", -1)
    err = np.abs(Ct.cpu().numpy() - want)
    assert (err <= 2e-3 * bound + 1e-4).all(), "tA=%d tB=%d M=%d N=%d K=%d max err %g" % (tA, tB, M, N, K, err.max())


def test_unaligned_operands_fall_back_to_fp32():
    import torch
    from kaldi_ctc_b200 import rnn
    A = torch.randn(50, 37, device="cuda")      # row pitch 37 floats: not 16-byte aligned rows
    Bm = torch.randn(20, 37, device="cuda")
    C = torch.zeros(50, 20, device="cuda")
    rnn.gemm(torch, 0, 1, 50, 20, 37, 1.0, A, 37, Bm, 37, 0.0, C, 20, math=rnn.MATH_TENSOR)
    torch.cuda.synchronize()
    assert rnn.lib().b200rnnLastGemmUsedTensorCores() == 0
    np.testing.assert_allclose(C.cpu().numpy(), (A @ Bm.T).cpu().numpy(), atol=1e-4)
