"""Host-side mirror of the training input path:

  kaldi::CompressedMatrix    src/matrix/compressed-matrix.{h,cc}  (in-memory image, compress, binary I/O)
  kaldi::ctc::NnetCtcExample src/ctc/ctc-nnet-example.{h,cc}      (struct, binary Read/Write, archives)
  kaldi::ctc::FormatNnetInput src/ctc/ctc-nnet-update.cc:351-424  -> b200ctc_format_input (GPU)
  FrameSubsamplingShiftFeatureTimes  ctc-nnet-example.cc:78-93

Decompression happens ONLY on the GPU (libb200ctc.so); this module never expands a CompressedMatrix on
the host.  Compression (needed to write egs) follows the reference expression by expression, float/double
mixing included, so that files written here are byte-identical to what the reference would write.
"""
import ctypes
import io
import struct

import numpy as np

from . import _lib, ctc

_F32 = np.float32


class CompressedMatrix:
    """In-memory image = what CompressedMatrix::data_ points to (GlobalHeader incl. `format`)."""

    def __init__(self, blob=b""):
        self.blob = bytes(blob)

    # -- header access
    def _hdr(self):
        return struct.unpack_from("<iffii", self.blob, 0) if self.blob else (1, 0.0, 0.0, 0, 0)

    def NumRows(self):
        return self._hdr()[3]

    def NumCols(self):
        return self._hdr()[4]

    @staticmethod
    def data_size(fmt, rows, cols):  # CompressedMatrix::DataSize, .cc:28-38
        return 20 + cols * (8 + rows) if fmt == 1 else 20 + 2 * rows * cols

    # -- compress (CopyFromMat, .cc:41-121)
    @staticmethod
    def _float_to_uint16(v, mn, rng):  # .cc:234-243: float division, float*65535 then + 0.499 in double
        f = (v.astype(_F32) - _F32(mn)) / _F32(rng)
        f = np.minimum(np.maximum(f, _F32(0.0)), _F32(1.0))
        return ((f * _F32(65535)).astype(np.float64) + 0.499).astype(np.int64).astype(np.uint16)

    @staticmethod
    def _uint16_to_float(v, mn, rng):  # .cc:245-251, all float
        return _F32(mn) + (_F32(rng) * _F32(1.52590218966964e-05)) * v.astype(_F32)

    @classmethod
    def from_matrix(cls, mat):
        mat = np.ascontiguousarray(mat, dtype=_F32)
        rows, cols = mat.shape
        if rows == 0 or cols == 0:
            return cls()
        assert np.isfinite(mat).all(), "cannot compress inf/nan"
        mn, mx = _F32(mat.min()), _F32(mat.max())
        if mx == mn:
            mx = _F32(np.float64(mn) + (1.0 + abs(np.float64(mn))))
        rng = _F32(mx - mn)
        if rng <= 0.0:
            rng = _F32(1.0e-05)
        fmt = 1 if rows > 8 else 2
        out = bytearray(struct.pack("<iffii", fmt, float(mn), float(rng), rows, cols))
        if fmt == 2:
            out += cls._float_to_uint16(mat, mn, rng).astype("<u2").tobytes()
            return cls(out)
        srt = np.sort(mat, axis=0)  # nth_element leaves exactly these order statistics in place (.cc:263-286)
        q = rows // 4
        u = lambda a: cls._float_to_uint16(a, mn, rng).astype(np.int64)
        p0 = np.minimum(u(srt[0]), 65532)
        p25 = np.minimum(np.maximum(u(srt[q]), p0 + 1), 65533)
        p75 = np.minimum(np.maximum(u(srt[3 * q]), p25 + 1), 65534)
        p100 = np.maximum(u(srt[rows - 1]), p75 + 1)
        hdr = np.stack([p0, p25, p75, p100], axis=1).astype("<u2")
        out += hdr.tobytes()
        f0, f25, f75, f100 = (cls._uint16_to_float(h.astype(np.uint16), mn, rng) for h in (p0, p25, p75, p100))
        # FloatToChar (.cc:334-361), column-wise
        v = mat
        lo = v < f25[None, :]
        mid = (~lo) & (v < f75[None, :])
        r1 = ((v - f0) / (f25 - f0) * _F32(64)).astype(np.float64) + 0.5
        r2 = ((v - f25) / (f75 - f25) * _F32(128)).astype(np.float64) + 0.5
        r3 = ((v - f75) / (f100 - f75) * _F32(63)).astype(np.float64) + 0.5
        a1 = np.clip(np.trunc(r1), 0, 64)
        a2 = np.clip(64 + np.trunc(r2), 64, 192)
        a3 = np.clip(192 + np.trunc(r3), 192, 255)
        ans = np.where(lo, a1, np.where(mid, a2, a3)).astype(np.uint8)
        out += np.ascontiguousarray(ans.T).tobytes()  # column-major bytes
        return cls(out)

    # -- binary I/O (.cc:405-470): token CM / CM2, then the image minus its first 4 bytes
    def Write(self, os):
        if not self.blob:
            os.write(b"CM " + struct.pack("<iffii", 0, 0.0, 0.0, 0, 0))  # the reference writes sizeof(h) bytes here
            return
        os.write(b"CM " if self._hdr()[0] == 1 else b"CM2 ")
        os.write(self.blob[4:])

    @classmethod
    def Read(cls, is_):
        tok = _read_token(is_)
        if tok not in ("CM", "CM2"):
            raise ValueError("Unexpected token %s, expecting CM or CM2." % tok)
        fmt = 1 if tok == "CM" else 2
        rest = is_.read(16)
        mn, rng, rows, cols = struct.unpack("<ffii", rest)
        if cols == 0:
            # empty matrix; the reference's writer emitted a full 20-byte header after "CM " (format included)
            is_.read(4)
            return cls()
        n = cls.data_size(fmt, rows, cols) - 20
        body = is_.read(n)
        if len(body) != n:
            raise ValueError("Failed to read compressed matrix data")
        return cls(struct.pack("<i", fmt) + rest + body)


def _read_token(is_):
    out = bytearray()
    while True:
        c = is_.read(1)
        if not c:
            raise ValueError("ReadToken, failed to read token")
        if c in b" \t\n":
            if out:
                return out.decode()
            continue
        out += c


def _expect(is_, tok):
    got = _read_token(is_)
    if got != tok:
        raise ValueError("Expected token \"%s\", got instead \"%s\"." % (tok, got))


class NnetCtcExample:
    """struct NnetCtcExample (ctc-nnet-example.h): labels, input_frames (CompressedMatrix), left_context, spk_info."""

    def __init__(self, labels=(), input_frames=None, left_context=0, spk_info=()):
        self.labels = [int(v) for v in labels]
        self.input_frames = input_frames if input_frames is not None else CompressedMatrix()
        self.left_context = int(left_context)
        self.spk_info = np.asarray(spk_info, dtype=_F32).reshape(-1)

    def NumFrames(self):
        return self.input_frames.NumRows()

    def NumLabels(self):
        return len(self.labels)

    def SetLabels(self, alignment):
        assert len(alignment) <= self.input_frames.NumRows()
        self.labels = [int(v) for v in alignment]

    def Write(self, os):  # binary mode, ctc-nnet-example.cc:28-44
        os.write(b"<NnetCtcExample> <Labels> ")
        os.write(b"\x04" + struct.pack("<i", len(self.labels)) + np.asarray(self.labels, dtype="<i4").tobytes())
        os.write(b"<InputFrames> ")
        self.input_frames.Write(os)
        os.write(b"<LeftContext> \x04" + struct.pack("<i", self.left_context))
        os.write(b"<SpkInfo> FV \x04" + struct.pack("<i", self.spk_info.size) + self.spk_info.astype("<f4").tobytes())
        os.write(b"</NnetCtcExample> ")

    @classmethod
    def Read(cls, is_):  # :46-60
        _expect(is_, "<NnetCtcExample>")
        _expect(is_, "<Labels>")
        if is_.read(1) != b"\x04":
            raise ValueError("ReadIntegerVector: expected size-of-int 4")
        n = struct.unpack("<i", is_.read(4))[0]
        labels = np.frombuffer(is_.read(4 * n), dtype="<i4")
        _expect(is_, "<InputFrames>")
        frames = CompressedMatrix.Read(is_)
        _expect(is_, "<LeftContext>")
        if is_.read(1) != b"\x04":
            raise ValueError("ReadBasicType: expected size 4")
        left = struct.unpack("<i", is_.read(4))[0]
        _expect(is_, "<SpkInfo>")
        _expect(is_, "FV")
        is_.read(1)
        d = struct.unpack("<i", is_.read(4))[0]
        spk = np.frombuffer(is_.read(4 * d), dtype="<f4")
        _expect(is_, "</NnetCtcExample>")
        return cls(labels, frames, left, spk)


def write_egs_ark(path_or_file, items):
    """Binary Kaldi archive: `key ` + "\\0B" + object, per entry (util/kaldi-holder, kaldi-table)."""
    f = open(path_or_file, "wb") if isinstance(path_or_file, str) else path_or_file
    for key, eg in items:
        f.write(key.encode() + b" \x00B")
        eg.Write(f)
    if isinstance(path_or_file, str):
        f.close()


def read_egs_ark(path_or_file):
    f = open(path_or_file, "rb") if isinstance(path_or_file, str) else path_or_file
    out = []
    while True:
        c = f.read(1)
        if not c:
            break
        key = bytearray(c)
        while True:
            c = f.read(1)
            if c == b" ":
                break
            key += c
        if f.read(2) != b"\x00B":
            raise ValueError("only binary archives are supported")
        out.append((key.decode(), NnetCtcExample.Read(f)))
    if isinstance(path_or_file, str):
        f.close()
    return out


def FrameSubsamplingShiftFeatureTimes(frame_subsampling_factor, frame_shift, feature):
    """ctc-nnet-example.cc:78-93: rows frame_shift, frame_shift + f, ... (unchanged when none qualifies)."""
    idx = np.arange(frame_shift, feature.shape[0], frame_subsampling_factor)
    return feature if idx.size == 0 else np.ascontiguousarray(feature[idx])


class InputStager:
    """Pinned + device staging for b200ctc_format_input, double-buffered so that packing minibatch
    k+1 never touches the buffer minibatch k's upload is still reading."""

    def __init__(self, device="cuda:0"):
        self.torch = _lib.require_cuda()
        self.device = self.torch.device(device)
        self.host = [None, None]
        self.dev = [None, None]
        self.done = [None, None]
        self.k = 0
        L = ctc.lib()
        if not getattr(L, "_fmt_configured", False):
            L.b200ctc_format_input_size.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                                    ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                                    ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_size_t)]
            L.b200ctc_format_input.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                               ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                               ctypes.c_void_p]
            L._fmt_configured = True
        self.L = L

    def _buffers(self, nbytes):
        t, k = self.torch, self.k
        if self.host[k] is None or self.host[k].numel() < nbytes:
            cap = int(nbytes * 1.25) + 4096
            for j in (0, 1):   # both halves at once: pinning is slow, keep it out of the steady state
                if self.done[j] is not None:
                    self.done[j].synchronize()
                self.host[j] = t.empty(cap, dtype=t.uint8).pin_memory()
                self.dev[j] = t.empty(cap, dtype=t.uint8, device=self.device)
                self.done[j] = None
        if self.done[k] is not None:
            self.done[k].synchronize()  # the upload that last used this pair has completed
        return self.host[k], self.dev[k]


def FormatNnetInput(nnet_left_context, nnet_right_context, data, input_mat=None, stager=None, device="cuda:0"):
    """kaldi::ctc::FormatNnetInput: `data` = list of NnetCtcExample -> device matrix
    [max_num_frames * num_splice * len(data), feat_dim + spk_dim] (returned, with max_num_frames).
    input_mat: optional preallocated device buffer (first rows are used)."""
    assert len(data) > 0
    stager = stager or InputStager(device)
    t, L = stager.torch, stager.L
    B = len(data)
    spk_dim = int(data[0].spk_info.size)
    keep = [np.frombuffer(eg.input_frames.blob, dtype=np.uint8) for eg in data]   # keeps the images alive
    ptrs = (ctypes.c_void_p * B)(*[k.ctypes.data for k in keep])
    spk_keep = [np.ascontiguousarray(eg.spk_info, dtype=_F32) for eg in data]
    spk = (ctypes.c_void_p * B)(*[s.ctypes.data for s in spk_keep]) if spk_dim else None
    mf, fd, nb = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_size_t(0)
    left = data[0].left_context
    ctc._check(L.b200ctc_format_input_size(ptrs, B, spk_dim, left, nnet_left_context, nnet_right_context,
                                           ctypes.byref(mf), ctypes.byref(fd), ctypes.byref(nb)),
               "b200ctc_format_input_size")
    num_splice = 1 + nnet_left_context + nnet_right_context
    rows, cols = mf.value * num_splice * B, fd.value + spk_dim
    if input_mat is None:
        input_mat = t.empty(rows, cols, device=stager.device)
    assert input_mat.is_contiguous() and input_mat.shape[1] == cols and input_mat.shape[0] >= rows
    host, dev = stager._buffers(nb.value)
    stream = t.cuda.current_stream(stager.device)
    with t.cuda.device(stager.device):
        ctc._check(L.b200ctc_format_input(ptrs, spk, spk_dim, B, left, nnet_left_context, nnet_right_context,
                                          input_mat.data_ptr(), input_mat.numel(), host.data_ptr(), dev.data_ptr(),
                                          host.numel(), stream.cuda_stream), "b200ctc_format_input")
    ev = t.cuda.Event()
    ev.record(stream)
    stager.done[stager.k] = ev
    stager.k ^= 1
    stager.h2d_bytes = nb.value
    return input_mat[:rows], mf.value
