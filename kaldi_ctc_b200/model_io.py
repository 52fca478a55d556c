"""Model-file side of the drop-in: the binary nnet2 `Nnet` format for the components of
the CTC topology, and the <FilterParams> blob <-> PyTorch parameter order.

  Nnet::Write / Read                     src/nnet2/nnet-nnet.cc:170-205
  Component::ReadNew                     src/nnet2/nnet-component.cc:38-48
  CuDNNRecurrentComponent::Write / Read  src/nnet2/nnet-cudnn-component.cc:673-721
  ClipGradientComponent::Write           src/nnet2/nnet-cudnn-component.cc:814-837
  AffineComponent::Write                 src/nnet2/nnet-component.cc:1260-1274
  NonlinearComponent::Write (Softmax)    src/nnet2/nnet-component.cc:398-412
  SpliceComponent::Write                 src/nnet2/nnet-component.cc:2822-2831
  basic types / vectors / matrices       src/base/io-funcs.cc:26-75, matrix/kaldi-vector.cc:1202-1222,
                                         matrix/kaldi-matrix.cc:1213-1245

A model trained by the reference (its <FilterParams> is cuDNN 5's packed blob) loads straight into
CuDNNRecurrentComponent.SetParams, and the other way round.  Pure host code (numpy).
"""
import struct

import numpy as np


# ---- primitives -----------------------------------------------------------------------------------
def _tok(os, t):
    os.write(t.encode() + b" ")


def _read_tok(is_):
    out = bytearray()
    while True:
        c = is_.read(1)
        if not c:
            raise ValueError("ReadToken: unexpected end of file")
        if c in b" \t\n\r":
            if out:
                return out.decode()
            continue
        out += c


def _expect(is_, t):
    got = _read_tok(is_)
    if got != t:
        raise ValueError("Expected token \"%s\", got instead \"%s\"." % (t, got))


def _w_i32(os, v):
    os.write(b"\x04" + struct.pack("<i", int(v)))


def _w_f32(os, v):
    os.write(b"\x04" + struct.pack("<f", float(v)))


def _w_f64(os, v):
    os.write(b"\x08" + struct.pack("<d", float(v)))


def _w_bool(os, v):
    os.write(b"T" if v else b"F")


def _r_sized(is_, size, fmt):
    n = is_.read(1)
    if n != bytes([size]):
        raise ValueError("ReadBasicType: expected a %d-byte value, saw size byte %r" % (size, n))
    return struct.unpack(fmt, is_.read(size))[0]


def _r_i32(is_):
    return _r_sized(is_, 4, "<i")


def _r_f32(is_):
    return _r_sized(is_, 4, "<f")


def _r_f64(is_):
    return _r_sized(is_, 8, "<d")


def _r_bool(is_):
    c = is_.read(1)
    if c not in (b"T", b"F"):
        raise ValueError("ReadBasicType<bool>: expected T or F, got %r" % c)
    return c == b"T"


def _w_ivec(os, v):  # WriteIntegerVector<int32>, base/io-funcs-inl.h:198-211
    v = np.ascontiguousarray(v, dtype="<i4").reshape(-1)
    os.write(b"\x04" + struct.pack("<i", v.size) + v.tobytes())


def _r_ivec(is_):
    if is_.read(1) != b"\x04":
        raise ValueError("ReadIntegerVector: expected 4-byte elements")
    n = struct.unpack("<i", is_.read(4))[0]
    return np.frombuffer(is_.read(4 * n), dtype="<i4").copy()


def _w_vec(os, v, dtype="<f4"):
    v = np.ascontiguousarray(v, dtype=dtype).reshape(-1)
    _tok(os, "FV" if dtype == "<f4" else "DV")
    _w_i32(os, v.size)
    os.write(v.tobytes())


def _r_vec(is_):
    t = _read_tok(is_)
    if t not in ("FV", "DV"):
        raise ValueError("expected FV or DV, got " + t)
    n = _r_i32(is_)
    dt = "<f4" if t == "FV" else "<f8"
    return np.frombuffer(is_.read(n * int(dt[2])), dtype=dt).copy()


def _w_mat(os, m):
    m = np.ascontiguousarray(m, dtype="<f4")
    _tok(os, "FM")
    _w_i32(os, m.shape[0])
    _w_i32(os, m.shape[1])
    os.write(m.tobytes())


def _r_mat(is_):
    t = _read_tok(is_)
    if t not in ("FM", "DM"):
        raise ValueError("expected FM or DM, got " + t)
    r, c = _r_i32(is_), _r_i32(is_)
    dt = "<f4" if t == "FM" else "<f8"
    return np.frombuffer(is_.read(r * c * int(dt[2])), dtype=dt).reshape(r, c).copy()


# ---- components (dicts with a "type" key) -------------------------------------------------------------
def _write_component(os, c):
    t = c["type"]
    _tok(os, "<%s>" % t)
    if t == "CuDNNRecurrentComponent":
        _tok(os, "<LearningRate>"); _w_f32(os, c["learning_rate"])
        _tok(os, "<IsGradient>"); _w_bool(os, c.get("is_gradient", False))
        _tok(os, "<ClipGradient>"); _w_f32(os, c["clip_gradient"])
        _tok(os, "<InputDim>"); _w_i32(os, c["input_dim"])
        _tok(os, "<HiddenDim>"); _w_i32(os, c["hidden_dim"])
        _tok(os, "<NumLayers>"); _w_i32(os, c["num_layers"])
        _tok(os, "<Bidirectional>"); _w_bool(os, c["bidirectional"])
        _tok(os, "<RNNMode>"); _w_i32(os, c["rnn_mode"])
        _tok(os, "<MaxSeqLength>"); _w_i32(os, c["max_seq_length"])
        _tok(os, "<FilterParams>"); _w_vec(os, c["filter_params"])
    elif t == "ClipGradientComponent":
        _tok(os, "<Dim>"); _w_i32(os, c["dim"])
        _tok(os, "<ClippingThreshold>"); _w_f32(os, c["clipping_threshold"])
        _tok(os, "<NormBasedClipping>"); _w_bool(os, c.get("norm_based_clipping", True))
        _tok(os, "<SelfRepairClippedProportionThreshold>"); _w_f32(os, c.get("self_repair_clipped_proportion_threshold", 0.01))
        _tok(os, "<SelfRepairTarget>"); _w_f32(os, c.get("self_repair_target", 0.0))
        _tok(os, "<SelfRepairScale>"); _w_f32(os, c.get("self_repair_scale", 0.0))
        _tok(os, "<NumElementsClipped>"); _w_i32(os, c.get("num_clipped", 0))
        _tok(os, "<NumElementsProcessed>"); _w_i32(os, c.get("count", 0))
        _tok(os, "<NumSelfRepaired>"); _w_i32(os, c.get("num_self_repaired", 0))
        _tok(os, "<NumBackpropped>"); _w_i32(os, c.get("num_backpropped", 0))
    elif t == "AffineComponent":
        _tok(os, "<LearningRate>"); _w_f32(os, c["learning_rate"])
        _tok(os, "<LinearParams>"); _w_mat(os, c["linear_params"])
        _tok(os, "<BiasParams>"); _w_vec(os, c["bias_params"])
        _tok(os, "<IsGradient>"); _w_bool(os, c.get("is_gradient", False))
    elif t == "SpliceComponent":
        _tok(os, "<InputDim>"); _w_i32(os, c["input_dim"])
        _tok(os, "<Context>"); _w_ivec(os, c["context"])
        _tok(os, "<ConstComponentDim>"); _w_i32(os, c.get("const_component_dim", 0))
    elif t == "SoftmaxComponent":
        _tok(os, "<Dim>"); _w_i32(os, c["dim"])
        _tok(os, "<ValueSum>"); _w_vec(os, c.get("value_sum", np.zeros(0)), "<f8")
        _tok(os, "<DerivSum>"); _w_vec(os, c.get("deriv_sum", np.zeros(0)), "<f8")
        _tok(os, "<Count>"); _w_f64(os, c.get("count", 0.0))
    else:
        raise ValueError("Unknown component type " + t)
    _tok(os, "</%s>" % t)


def _read_component(is_):
    tok = _read_tok(is_)
    t = tok[1:-1]
    c = {"type": t}
    if t == "CuDNNRecurrentComponent":
        _expect(is_, "<LearningRate>"); c["learning_rate"] = _r_f32(is_)
        _expect(is_, "<IsGradient>"); c["is_gradient"] = _r_bool(is_)
        _expect(is_, "<ClipGradient>"); c["clip_gradient"] = _r_f32(is_)
        _expect(is_, "<InputDim>"); c["input_dim"] = _r_i32(is_)
        _expect(is_, "<HiddenDim>"); c["hidden_dim"] = _r_i32(is_)
        _expect(is_, "<NumLayers>"); c["num_layers"] = _r_i32(is_)
        _expect(is_, "<Bidirectional>"); c["bidirectional"] = _r_bool(is_)
        _expect(is_, "<RNNMode>"); c["rnn_mode"] = _r_i32(is_)
        _expect(is_, "<MaxSeqLength>"); c["max_seq_length"] = _r_i32(is_)
        _expect(is_, "<FilterParams>"); c["filter_params"] = _r_vec(is_)
    elif t == "ClipGradientComponent":
        _expect(is_, "<Dim>"); c["dim"] = _r_i32(is_)
        _expect(is_, "<ClippingThreshold>"); c["clipping_threshold"] = _r_f32(is_)
        _expect(is_, "<NormBasedClipping>"); c["norm_based_clipping"] = _r_bool(is_)
        _expect(is_, "<SelfRepairClippedProportionThreshold>"); c["self_repair_clipped_proportion_threshold"] = _r_f32(is_)
        _expect(is_, "<SelfRepairTarget>"); c["self_repair_target"] = _r_f32(is_)
        _expect(is_, "<SelfRepairScale>"); c["self_repair_scale"] = _r_f32(is_)
        _expect(is_, "<NumElementsClipped>"); c["num_clipped"] = _r_i32(is_)
        _expect(is_, "<NumElementsProcessed>"); c["count"] = _r_i32(is_)
        _expect(is_, "<NumSelfRepaired>"); c["num_self_repaired"] = _r_i32(is_)
        _expect(is_, "<NumBackpropped>"); c["num_backpropped"] = _r_i32(is_)
    elif t == "AffineComponent":
        _expect(is_, "<LearningRate>"); c["learning_rate"] = _r_f32(is_)
        _expect(is_, "<LinearParams>"); c["linear_params"] = _r_mat(is_)
        _expect(is_, "<BiasParams>"); c["bias_params"] = _r_vec(is_)
        _expect(is_, "<IsGradient>"); c["is_gradient"] = _r_bool(is_)
    elif t == "SpliceComponent":
        _expect(is_, "<InputDim>"); c["input_dim"] = _r_i32(is_)
        _expect(is_, "<Context>"); c["context"] = _r_ivec(is_)
        _expect(is_, "<ConstComponentDim>"); c["const_component_dim"] = _r_i32(is_)
    elif t == "SoftmaxComponent":
        _expect(is_, "<Dim>"); c["dim"] = _r_i32(is_)
        _expect(is_, "<ValueSum>"); c["value_sum"] = _r_vec(is_)
        _expect(is_, "<DerivSum>"); c["deriv_sum"] = _r_vec(is_)
        _expect(is_, "<Count>"); c["count"] = _r_f64(is_)
    else:
        raise ValueError("Unknown component type " + t)
    _expect(is_, "</%s>" % t)
    return c


def write_nnet(os, components, binary_header=True):
    """Nnet::Write in binary mode (a file written by the reference starts with "\\0B")."""
    if binary_header:
        os.write(b"\x00B")
    _tok(os, "<Nnet>")
    _tok(os, "<NumComponents>"); _w_i32(os, len(components))
    _tok(os, "<Components>")
    for c in components:
        _write_component(os, c)
    _tok(os, "</Components>")
    _tok(os, "</Nnet>")


def read_nnet(is_):
    head = is_.read(2)
    if head != b"\x00B":
        raise ValueError("only binary nnet2 files are supported")
    _expect(is_, "<Nnet>")
    _expect(is_, "<NumComponents>")
    n = _r_i32(is_)
    _expect(is_, "<Components>")
    comps = [_read_component(is_) for _ in range(n)]
    _expect(is_, "</Components>")
    _expect(is_, "</Nnet>")
    return comps


def components_of(spec, blobs, affine_w, affine_b, max_seq_length=2000, softmax=False):
    """The `cudnn_google` topology as component dicts (steps/ctc/nnet2/make_configs.py)."""
    dirs = 2 if spec.bidir else 1
    out = []
    for l, blob in enumerate(blobs):
        out.append({"type": "CuDNNRecurrentComponent", "learning_rate": spec.learning_rate,
                    "clip_gradient": spec.clip_gradient, "input_dim": spec.D if l == 0 else spec.H * dirs,
                    "hidden_dim": spec.H, "num_layers": 1, "bidirectional": bool(spec.bidir),
                    "rnn_mode": spec.mode, "max_seq_length": max_seq_length,
                    "filter_params": np.asarray(blob, dtype=np.float32)})
        out.append({"type": "ClipGradientComponent", "dim": spec.H * dirs,
                    "clipping_threshold": spec.clipping_threshold, "norm_based_clipping": True})
    out.append({"type": "AffineComponent", "learning_rate": spec.learning_rate,
                "linear_params": np.asarray(affine_w, dtype=np.float32),
                "bias_params": np.asarray(affine_b, dtype=np.float32)})
    if softmax:
        out.append({"type": "SoftmaxComponent", "dim": int(np.asarray(affine_b).size)})
    return out


# ---- <FilterParams> blob <-> PyTorch ---------------------------------------------------------------
_GATES = {0: 1, 1: 1, 2: 4, 3: 3}


def _layout(mode, bidir, layers, D, H):
    """(pseudo-layer, kind, offset, shape) of every block of the cuDNN-v5 packed blob: all matrices of all
    pseudo-layers first (per pseudo-layer: G input matrices [H x in], then G recurrent [H x H]), then all
    biases (per pseudo-layer: G input biases, G recurrent biases).  Gate order i,f,g,o / r,z,n = PyTorch's."""
    G, dirs = _GATES[mode], 2 if bidir else 1
    off, mats, biases = 0, [], []
    for pl in range(layers * dirs):
        din = D if pl // dirs == 0 else H * dirs
        mats.append((pl, "w_ih", off, (G * H, din))); off += G * H * din
        mats.append((pl, "w_hh", off, (G * H, H))); off += G * H * H
    for pl in range(layers * dirs):
        biases.append((pl, "b_ih", off, (G * H,))); off += G * H
        biases.append((pl, "b_hh", off, (G * H,))); off += G * H
    return mats + biases, off


def filter_params_to_torch(blob, mode, bidir, layers, D, H):
    """-> {"weight_ih_l0": ..., "weight_hh_l0_reverse": ..., "bias_ih_l0": ...} (numpy arrays), the
    state-dict names of torch.nn.LSTM / GRU / RNN(bidirectional=bidir, num_layers=layers)."""
    blob = np.asarray(blob, dtype=np.float32)
    blocks, total = _layout(mode, bidir, layers, D, H)
    assert blob.size == total, "blob has %d floats, layout needs %d" % (blob.size, total)
    dirs = 2 if bidir else 1
    names = {"w_ih": "weight_ih", "w_hh": "weight_hh", "b_ih": "bias_ih", "b_hh": "bias_hh"}
    out = {}
    for pl, kind, off, shape in blocks:
        key = "%s_l%d%s" % (names[kind], pl // dirs, "_reverse" if (pl % dirs) == 1 else "")
        out[key] = blob[off:off + int(np.prod(shape))].reshape(shape).copy()
    return out


def torch_to_filter_params(state, mode, bidir, layers, D, H):
    blocks, total = _layout(mode, bidir, layers, D, H)
    dirs = 2 if bidir else 1
    names = {"w_ih": "weight_ih", "w_hh": "weight_hh", "b_ih": "bias_ih", "b_hh": "bias_hh"}
    blob = np.zeros(total, dtype=np.float32)
    for pl, kind, off, shape in blocks:
        key = "%s_l%d%s" % (names[kind], pl // dirs, "_reverse" if (pl % dirs) == 1 else "")
        v = np.asarray(state[key], dtype=np.float32)
        assert v.shape == tuple(shape), (key, v.shape, shape)
        blob[off:off + v.size] = v.reshape(-1)
    return blob
