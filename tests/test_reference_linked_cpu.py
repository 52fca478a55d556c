"""Host-side formats checked against the REFERENCE'S OWN CODE, compiled from /root/reference by
oracle/ref/Makefile into oracle/_ref/ref_component_harness (integration/kaldi/harness/component_harness.cc):

  * kaldi_ctc_b200/model_io.py  vs  Nnet::Read / Nnet::Write (src/nnet2/nnet-nnet.cc:170-205) and the
    TransitionModel + AmNnet container that nnet2-ctc-train-simple reads
  * kaldi_ctc_b200/egs.py       vs  NnetCtcExample::Read (src/ctc/ctc-nnet-example.cc:46-60) through the
    reference's SequentialTableReader, and CompressedMatrix compress / decompress
    (src/matrix/compressed-matrix.cc:41-121, 493-529) -- byte / bit exact
  * oracle/feat_oracle.c        vs  the same (this is what pins the restatement that the GPU
    FormatNnetInput is compared with), and kaldi::ctc::FormatNnetInput (ctc-nnet-update.cc:351-424) itself
None of this needs a GPU."""
import io
import os

import numpy as np
import pytest

import refbin
from kaldi_ctc_b200 import egs, model_io, synth
from oracle import pyoracle

pytestmark = pytest.mark.skipif(not refbin.ensure_built("gpu"),
                                reason="oracle/_ref not built and /root/reference absent")


def _pathological(rng, rows, cols):
    M = rng.standard_normal((rows, cols)).astype(np.float32) if rng.integers(3) else \
        np.full((rows, cols), rng.standard_normal(), dtype=np.float32)
    if rng.integers(2) and rows:
        M[rng.integers(rows)] = rng.standard_normal() * 4.0
    val = np.float32(rng.standard_normal() * 4.0)
    M[rng.integers(1 + rng.integers(5), size=M.shape) != 0] = val
    return M


def test_model_file_roundtrip_through_reference_reader_and_writer(tmp_path):
    spec = synth.ModelSpec(D=8, H=12, layers=2, A=10)
    blobs, aw, ab = synth.model_weights(spec, 3)
    comps = model_io.components_of(spec, blobs, aw, ab, max_seq_length=50, softmax=True)
    with open(tmp_path / "in.nnet", "wb") as f:
        model_io.write_nnet(f, comps)
    out = refbin.run(refbin.HARNESS, "make-model", tmp_path / "in.nnet", 9, tmp_path / "model.mdl").stdout
    assert "6 components, 10 pdfs" in out
    refbin.run(refbin.HARNESS, "extract-nnet", tmp_path / "model.mdl", tmp_path / "out.nnet")
    with open(tmp_path / "out.nnet", "rb") as f:
        back = model_io.read_nnet(f)
    assert [c["type"] for c in back] == [c["type"] for c in comps]
    for a, b in zip(comps, back):
        for k, v in a.items():
            if isinstance(v, np.ndarray):
                assert np.array_equal(v, b[k]), (a["type"], k)
            elif k != "type":
                assert b[k] == pytest.approx(v), (a["type"], k)


def test_egs_archive_read_by_the_reference_reader(tmp_path):
    rng = np.random.default_rng(11)
    items = []
    for i in range(6):
        rows, cols = int(rng.integers(1, 40)), 7          # both storage formats (rows <= 8: uint16)
        frames = _pathological(rng, rows, cols)
        labels = rng.integers(1, 30, size=int(rng.integers(0, 9))).astype(np.int32)
        spk = rng.standard_normal(3).astype(np.float32) if i % 2 else []
        items.append(("utt%d" % i, egs.NnetCtcExample(labels, egs.CompressedMatrix.from_matrix(frames), i % 3, spk)))
    egs.write_egs_ark(str(tmp_path / "egs.ark"), items)
    refbin.run(refbin.HARNESS, "dump-egs", "ark:%s" % (tmp_path / "egs.ark"), tmp_path / "d")
    meta = [l.split() for l in open(tmp_path / "d.meta.txt")]
    assert len(meta) == len(items)
    for i, ((key, eg), m) in enumerate(zip(items, meta)):
        rows, cols, left, nl = int(m[1]), int(m[2]), int(m[3]), int(m[4])
        assert m[0] == key and rows == eg.NumFrames() and left == eg.left_context and nl == eg.NumLabels()
        assert [int(v) for v in m[5:5 + nl]] == [int(v) for v in eg.labels]
        ns = int(m[6 + nl])
        assert ns == len(eg.spk_info)
        np.testing.assert_allclose([float(v) for v in m[7 + nl:7 + nl + ns]], eg.spk_info, rtol=1e-5)
        ref = np.fromfile(tmp_path / ("d.%d.frames.f32" % i), dtype=np.float32).reshape(rows, cols)
        # the reference's CompressedMatrix::CopyToMat == the oracle's restatement, bit for bit
        assert np.array_equal(ref, pyoracle.cm_decompress(eg.input_frames.blob)), "example %d" % i


def test_compress_matches_the_reference_bytes(tmp_path):
    rng = np.random.default_rng(5)
    for n in range(40):
        rows, cols = int(rng.integers(1, 30)), int(rng.integers(1, 12))
        M = _pathological(rng, rows, cols)
        M.tofile(tmp_path / "m.f32")
        refbin.run(refbin.HARNESS, "compress", tmp_path / "m.f32", rows, cols, tmp_path / "m.cm")
        ref_bytes = open(tmp_path / "m.cm", "rb").read()
        mine = io.BytesIO()
        egs.CompressedMatrix.from_matrix(M).Write(mine)
        assert mine.getvalue() == ref_bytes, "case %d (%dx%d): egs.py differs from CompressedMatrix::Write" % (n, rows, cols)
        # and the C restatement produces the same in-memory image
        image = pyoracle.cm_compress(M)
        assert image == egs.CompressedMatrix.from_matrix(M).blob


@pytest.mark.parametrize("context,left_context,spk_dim", [((0,), 0, 0), ((-1, 0, 1), 1, 0), ((-2, 0), 3, 2), ((0, 1), 0, 1)])
def test_format_nnet_input_matches_the_reference(tmp_path, context, left_context, spk_dim):
    rng = np.random.default_rng(7)
    D, B = 5, 4
    nl, nr = -min(context), max(context)
    splice = 1 + nl + nr
    in_dim = D + spk_dim
    out_dim = (in_dim - spk_dim) * len(context) + spk_dim
    comps = [{"type": "SpliceComponent", "input_dim": in_dim, "context": list(context), "const_component_dim": spk_dim},
             {"type": "AffineComponent", "learning_rate": 0.1, "linear_params": np.zeros((3, out_dim), np.float32),
              "bias_params": np.zeros(3, np.float32)}]
    with open(tmp_path / "n.nnet", "wb") as f:
        model_io.write_nnet(f, comps)
    items = []
    for i in range(B):
        rows = int(rng.integers(splice + left_context - nl, 25))
        frames = rng.standard_normal((rows, D)).astype(np.float32)
        spk = rng.standard_normal(spk_dim).astype(np.float32) if spk_dim else []
        items.append(("u%d" % i, egs.NnetCtcExample([1, 2], egs.CompressedMatrix.from_matrix(frames), left_context, spk)))
    egs.write_egs_ark(str(tmp_path / "e.ark"), items)
    out = refbin.run(refbin.HARNESS, "format-input", tmp_path / "n.nnet", "ark:%s" % (tmp_path / "e.ark"), B,
                     tmp_path / "x.f32").stdout
    ref = np.fromfile(tmp_path / "x.f32", dtype=np.float32)
    want, mf = pyoracle.format_nnet_input([e.input_frames.blob for _, e in items],
                                          [e.spk_info for _, e in items] if spk_dim else None, left_context, nl, nr)
    assert "%d x %d" % want.shape in out
    assert np.array_equal(ref.reshape(want.shape), want)
