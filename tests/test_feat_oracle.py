"""Pins oracle/feat_oracle.c (CompressedMatrix + FormatNnetInput restatement) with hand-computed known
answers and the properties the reference's own test checks (matrix-lib-test.cc:4126-4230), and checks the
product's host-side egs code (compressor, binary I/O) against it byte for byte."""
import io
import struct

import numpy as np
import pytest

from kaldi_ctc_b200 import egs
from oracle import pyoracle


def _pathological(rng, rows, cols):
    """The matrix generator of UnitTestCompressedMatrix (:4145-4162)."""
    M = rng.standard_normal((rows, cols)).astype(np.float32) if rng.integers(3) else \
        np.full((rows, cols), rng.standard_normal(), dtype=np.float32)
    if rng.integers(2) and rows:
        M[rng.integers(rows)] = rng.standard_normal() * 4.0
    val = np.float32(rng.standard_normal() * 4.0)
    modulus = 1 + rng.integers(5)
    M[rng.integers(modulus, size=M.shape) != 0] = val
    return M


def test_format2_known_answer():
    M = np.array([[0, 1], [2, 3]], dtype=np.float32)
    blob = pyoracle.cm_compress(M)
    fmt, mn, rng, rows, cols = struct.unpack_from("<iffii", blob)
    assert (fmt, mn, rng, rows, cols) == (2, 0.0, 3.0, 2, 2) and len(blob) == 20 + 8
    assert list(np.frombuffer(blob[20:], dtype="<u2")) == [0, 21845, 43690, 65535]
    np.testing.assert_allclose(pyoracle.cm_decompress(blob), M, atol=3.0 / 65535)


def test_format1_known_answer():
    M = np.stack([np.arange(9, dtype=np.float32), np.full(9, 4.0, dtype=np.float32)], axis=1)
    blob = pyoracle.cm_compress(M)
    fmt, mn, rng, rows, cols = struct.unpack_from("<iffii", blob)
    assert (fmt, mn, rng, rows, cols) == (1, 0.0, 8.0, 9, 2) and len(blob) == 20 + 2 * (8 + 9)
    hdr = np.frombuffer(blob[20:36], dtype="<u2").reshape(2, 4)
    assert list(hdr[0]) == [0, 16384, 49151, 65535]          # order statistics 0, 2, 6, 8 of 0..8
    assert list(hdr[1]) == [32767, 32768, 32769, 32770]      # constant column: forced strictly increasing
    col0 = np.frombuffer(blob[36:45], dtype=np.uint8)
    assert col0[0] == 0 and col0[8] == 255 and col0[2] in (63, 64) and col0[6] in (191, 192)
    assert (np.diff(col0.astype(int)) > 0).all()
    back = pyoracle.cm_decompress(blob)
    np.testing.assert_allclose(back[:, 0], M[:, 0], atol=8.0 / 128)
    np.testing.assert_allclose(back[:, 1], 4.0, atol=1e-3)


def test_reference_test_properties():
    """Sizes, near-lossless reconstruction and re-compression stability (the reference tolerates rare failures)."""
    rng = np.random.default_rng(0)
    unstable = 0
    for n in range(300):
        rows, cols = int(rng.integers(1, 20)), int(rng.integers(1, 15))
        M = _pathological(rng, rows, cols)
        blob = pyoracle.cm_compress(M)
        assert len(blob) == egs.CompressedMatrix.data_size(1 if rows > 8 else 2, rows, cols)
        M2 = pyoracle.cm_decompress(blob)
        span = max(float(M.max() - M.min()), 1e-5)
        assert np.abs(M2 - M).max() <= span / 60 + 1e-6      # worst bucket: quarter of the range over 63 steps
        M3 = pyoracle.cm_decompress(pyoracle.cm_compress(M2))
        if not np.allclose(M2, M3, rtol=0, atol=1e-4 * max(np.abs(M2).max(), 1e-3) * np.sqrt(M2.size)):
            unstable += 1
    assert unstable <= 3


def test_product_compressor_is_byte_identical_to_oracle():
    rng = np.random.default_rng(1)
    for n in range(200):
        rows, cols = int(rng.integers(1, 40)), int(rng.integers(1, 15))
        M = _pathological(rng, rows, cols) if n % 2 else (rng.standard_normal((rows, cols)) * 3).astype(np.float32)
        assert egs.CompressedMatrix.from_matrix(M).blob == pyoracle.cm_compress(M), (n, rows, cols)
    big = (rng.standard_normal((700, 40)) * 5 + 1).astype(np.float32)
    assert egs.CompressedMatrix.from_matrix(big).blob == pyoracle.cm_compress(big)


def test_format_nnet_input_oracle_against_direct_construction():
    rng = np.random.default_rng(2)
    left_context, nl, nr, spk_dim, D = 3, 1, 2, 2, 5
    S, ign = 1 + nl + nr, left_context - nl
    mats = [(rng.standard_normal((T, D)) * 2).astype(np.float32) for T in (20, 9, 14)]
    blobs = [pyoracle.cm_compress(m) for m in mats]
    spk = [rng.standard_normal(spk_dim).astype(np.float32) for _ in mats]
    out, mf = pyoracle.format_nnet_input(blobs, spk, left_context, nl, nr)
    n = [m.shape[0] - S - ign + 1 for m in mats]
    assert mf == max(n) and out.shape == (mf * S * 3, D + spk_dim)
    dec = [pyoracle.cm_decompress(b) for b in blobs]
    for b in range(3):
        for t in range(mf):
            for s in range(S):
                row = out[(t * 3 + b) * S + s]
                if t < n[b]:
                    assert np.array_equal(row[:D], dec[b][ign + s + t]) and np.array_equal(row[D:], spk[b])
                else:
                    assert not row.any()


def test_example_binary_io_round_trip_and_layout():
    rng = np.random.default_rng(3)
    eg = egs.NnetCtcExample([3, 1, 2], egs.CompressedMatrix.from_matrix(rng.standard_normal((12, 4))), 2, [0.5, -1.0])
    buf = io.BytesIO()
    eg.Write(buf)
    raw = buf.getvalue()
    assert raw.startswith(b"<NnetCtcExample> <Labels> \x04\x03\x00\x00\x00\x03\x00\x00\x00\x01\x00\x00\x00\x02\x00\x00\x00<InputFrames> CM ")
    assert raw.endswith(b"<LeftContext> \x04\x02\x00\x00\x00<SpkInfo> FV \x04\x02\x00\x00\x00" +
                        struct.pack("<ff", 0.5, -1.0) + b"</NnetCtcExample> ")
    back = egs.NnetCtcExample.Read(io.BytesIO(raw))
    assert back.labels == [3, 1, 2] and back.left_context == 2 and back.input_frames.blob == eg.input_frames.blob
    assert np.array_equal(back.spk_info, eg.spk_info) and back.NumFrames() == 12
    small = egs.NnetCtcExample([1], egs.CompressedMatrix.from_matrix(rng.standard_normal((3, 2))), 0, [])
    ark = io.BytesIO()
    egs.write_egs_ark(ark, [("utt1", eg), ("utt2", small)])
    items = egs.read_egs_ark(io.BytesIO(ark.getvalue()))
    assert [k for k, _ in items] == ["utt1", "utt2"] and items[1][1].input_frames.blob == small.input_frames.blob
    assert b"CM2 " in ark.getvalue()   # <= 8 rows -> uint16 format (compressed-matrix.cc:85-89)


def test_frame_subsampling_row_selection():
    f = np.arange(20, dtype=np.float32).reshape(10, 2)
    assert np.array_equal(egs.FrameSubsamplingShiftFeatureTimes(3, 1, f), f[[1, 4, 7]])
    short = f[:2]
    assert egs.FrameSubsamplingShiftFeatureTimes(3, 2, short) is short   # no row qualifies: unchanged (:87-88)
