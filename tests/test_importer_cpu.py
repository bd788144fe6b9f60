"""Checkpoint importer, host half (reference: initialize(), obj_detect_tracking.py:392-448): the
variable manifest in the reference's checkpoint naming, ':0' tensor names, ignored extras, loud failure on missing /
mis-shaped variables (the reference would silently keep random initial values)."""
import numpy as np
import pytest

from object_detection_tracking_b200.backend import check_weights
from object_detection_tracking_b200.config import make_config
from object_detection_tracking_b200.synth import frcnn_weight_shapes, synth_weights


def test_manifest_equals_synthetic_checkpoint_and_reference_param_counts():
    cfg = make_config(resnet_num_block=(1, 1, 2, 1))
    W = synth_weights(cfg, 3)
    sh = frcnn_weight_shapes(cfg)
    assert set(sh) == set(W)
    for k in sh:
        assert tuple(W[k].shape) == tuple(sh[k]), k
    # full-size graph: ResNet-101 convs 42.39 M + FPN 3.34 M + RPN 0.59 M = 46.33 M kernel weights (SURVEY 8's 48.7 M also
    # figure is an estimate that also counts BatchNorm vectors and biases), 14.0 M fc for 15 classes
    r101 = frcnn_weight_shapes(make_config())
    conv = sum(int(np.prod(s)) for n, s in r101.items() if n.endswith("/W") and len(s) == 4)
    fc = sum(int(np.prod(s)) for n, s in r101.items() if n.endswith("/W") and len(s) == 2)
    assert abs(conv - 46.33e6) < 0.02e6 and abs(fc - 14.0e6) < 0.1e6
    assert len([n for n in r101 if n.endswith("/conv2/W")]) == 33 and "group2/block22/conv3/bn/variance/EMA" in r101
    r50 = frcnn_weight_shapes(make_config(resnet50=True, num_class=81))
    assert "group2/block5/conv1/W" in r50 and "group2/block6/conv1/W" not in r50
    assert r50["fastrcnn/outputs/box/W"] == (1024, 324)
    agn = frcnn_weight_shapes(make_config(use_frcnn_class_agnostic=True))
    assert agn["fastrcnn/outputs/box/W"] == (1024, 4)


def test_importer_accepts_tensor_names_and_ignores_extras():
    cfg = make_config(resnet_num_block=(1, 1, 1, 1))
    W = synth_weights(cfg, 4)
    ck = {k + ":0": v.astype(np.float64) for k, v in W.items()}          # tensor names, other dtype
    ck["global_step:0"] = np.int64(90000)
    ck["learning_rate"] = np.float32(0.001)
    ck["conv0/W/Momentum:0"] = np.zeros_like(W["conv0/W"])
    ck["maskrcnn/fcn0/W:0"] = np.zeros((3, 3, 256, 256), np.float32)      # head this graph does not build
    got = check_weights(cfg, ck)
    assert set(got) == set(W)
    for k in W:
        assert got[k].dtype == np.float32 and got[k].flags["C_CONTIGUOUS"]
        np.testing.assert_array_equal(got[k], W[k])


def test_importer_fails_loudly_on_missing_or_misshaped():
    cfg = make_config(resnet_num_block=(1, 1, 1, 1))
    W = synth_weights(cfg, 5)
    bad = dict(W)
    del bad["group3/block0/conv2/bn/mean/EMA"]
    bad["fastrcnn/outputs/class/W"] = np.zeros((1024, 81), np.float32)    # COCO head into a 15-class graph
    with pytest.raises(ValueError) as e:
        check_weights(cfg, bad)
    assert "group3/block0/conv2/bn/mean/EMA" in str(e.value) and "fastrcnn/outputs/class/W" in str(e.value)
    with pytest.raises(ValueError):
        check_weights(make_config(), W)                                    # R101 graph, 4-block checkpoint


def test_model_load_npz_roundtrip(tmp_path):
    from object_detection_tracking_b200.backend import get_model
    cfg = make_config(resnet_num_block=(1, 1, 1, 1))
    W = synth_weights(cfg, 6)
    path = str(tmp_path / "ckpt.npz")
    np.savez(path, **{k + ":0": v for k, v in W.items()})
    model = get_model(cfg, gpuid=0)
    model.load_npz(path)                                                   # host side only: no device needed yet
    assert set(model._weights) == set(W)
    np.testing.assert_array_equal(model._weights["rpn/box/b"], W["rpn/box/b"])


def test_mask_head_variables_join_the_manifest_with_add_mask():
    cfg = make_config(resnet_num_block=(1, 1, 1, 1), add_mask=True)
    sh = frcnn_weight_shapes(cfg)
    assert sh["maskrcnn/fcn3/W"] == (3, 3, 256, 256) and sh["maskrcnn/deconv/W"] == (2, 2, 256, 256)
    assert sh["maskrcnn/conv/W"] == (1, 1, 256, 14) and sh["maskrcnn/conv/b"] == (14,)
    W = synth_weights(cfg, 9)
    assert set(W) == set(sh) and all(tuple(W[k].shape) == tuple(sh[k]) for k in sh)
    base = synth_weights(make_config(resnet_num_block=(1, 1, 1, 1)), 9)
    assert "maskrcnn/conv/W" not in base
    np.testing.assert_array_equal(base["fastrcnn/fc7/W"], W["fastrcnn/fc7/W"])      # detector weights independent of add_mask
    with pytest.raises(ValueError):
        check_weights(cfg, base)                                                      # mask graph, checkpoint without the head


# ---- frozen GraphDef (.pb) reader: a minimal protobuf WRITER for the test, following the same public schema -------------
def _vi(x):
    out = bytearray()
    while True:
        b = x & 0x7F
        x >>= 7
        out.append(b | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _ld(fnum, payload):
    return _vi((fnum << 3) | 2) + _vi(len(payload)) + payload


def _tensor_proto(arr, mode):
    dt = {np.dtype(np.float32): 1, np.dtype(np.int32): 3}[arr.dtype]
    shape = b"".join(_ld(2, _vi((1 << 3) | 0) + _vi(int(d))) for d in arr.shape)
    msg = _vi((1 << 3) | 0) + _vi(dt) + _ld(2, shape)
    if mode == "content":
        msg += _ld(4, arr.tobytes())
    elif mode == "packed":
        msg += _ld(5 if dt == 1 else 7, arr.astype("<f4").tobytes() if dt == 1 else b"".join(_vi(int(v)) for v in arr.ravel()))
    elif mode == "splat":                       # one value for a constant-filled tensor
        msg += _ld(5, np.float32(arr.ravel()[0]).tobytes())
    return msg


def _node(name, op, tensor=None):
    msg = _ld(1, name.encode()) + _ld(2, op.encode())
    if tensor is not None:
        msg += _ld(5, _ld(1, b"value") + _ld(2, _ld(8, tensor)))
        msg += _ld(5, _ld(1, b"dtype") + _ld(2, _vi((6 << 3) | 0) + _vi(1)))
    return _ld(1, msg)


def test_frozen_graph_reader_roundtrip_and_model_load_pb(tmp_path):
    from object_detection_tracking_b200.backend import get_model
    from object_detection_tracking_b200.pbreader import read_frozen_graph
    cfg = make_config(resnet_num_block=(1, 1, 1, 1))
    W = synth_weights(cfg, 8)
    W["group0/block0/conv1/bn/beta"][:] = 0.25                           # stored as a splat
    graph = _node("image", "Placeholder")
    for i, (k, v) in enumerate(sorted(W.items())):
        mode = "splat" if k == "group0/block0/conv1/bn/beta" else ("packed" if v.size <= 64 else "content")
        graph += _node(k, "Const", _tensor_proto(v, mode))
        graph += _node(k + "/read", "Identity")
    graph += _node("anchors_p2", "Const", _tensor_proto(np.arange(12, dtype=np.int32).reshape(3, 4), "packed"))
    graph += _vi((3 << 3) | 0) + _vi(27)                                  # GraphDef.version (varint field, skipped)
    path = str(tmp_path / "frozen.pb")
    with open(path, "wb") as f:
        f.write(graph)
    consts = read_frozen_graph(path)
    assert set(consts) == set(W)                                          # the int32 constant is not a weight
    for k in W:
        np.testing.assert_array_equal(consts[k], W[k])
    assert read_frozen_graph(path, float_only=False)["anchors_p2"].tolist() == np.arange(12).reshape(3, 4).tolist()
    cfg.is_load_from_pb, cfg.load_from = True, path
    model = get_model(cfg, gpuid=0)                                       # models.py:102-109 path
    assert set(model._weights) == set(W)
    np.testing.assert_array_equal(model._weights["fastrcnn/fc6/W"], W["fastrcnn/fc6/W"])


def test_frozen_graph_reader_against_the_official_tf_protos(tmp_path):
    """The same check with the GraphDef built by the generated classes of TensorFlow's own .proto files (shipped with
    tensorboard: tensorboard.compat.proto) -- pins the field numbers the reader assumes, tensor_content and *_val forms."""
    graph_pb2 = pytest.importorskip("tensorboard.compat.proto.graph_pb2")
    from tensorboard.compat.proto import tensor_pb2, tensor_shape_pb2, types_pb2
    from object_detection_tracking_b200.pbreader import read_frozen_graph
    rng = np.random.default_rng(2)
    g = graph_pb2.GraphDef()
    want = {}

    def const(name, arr, how):
        n = g.node.add()
        n.name, n.op = name, "Const"
        t = tensor_pb2.TensorProto()
        t.dtype = {np.dtype(np.float32): types_pb2.DT_FLOAT, np.dtype(np.float64): types_pb2.DT_DOUBLE,
                   np.dtype(np.float16): types_pb2.DT_HALF, np.dtype(np.int32): types_pb2.DT_INT32}[arr.dtype]
        t.tensor_shape.CopyFrom(tensor_shape_pb2.TensorShapeProto(
            dim=[tensor_shape_pb2.TensorShapeProto.Dim(size=int(d)) for d in arr.shape]))
        if how == "content":
            t.tensor_content = arr.tobytes()
        elif how == "val":
            if arr.dtype == np.float32:
                t.float_val.extend(arr.ravel().tolist())
            elif arr.dtype == np.float64:
                t.double_val.extend(arr.ravel().tolist())
            elif arr.dtype == np.float16:
                t.half_val.extend(arr.ravel().view(np.uint16).tolist())
            else:
                t.int_val.extend(arr.ravel().tolist())
        elif how == "splat":
            t.float_val.append(float(arr.ravel()[0]))
        n.attr["value"].tensor.CopyFrom(t)
        n.attr["dtype"].type = t.dtype
        want[name] = arr

    ph = g.node.add()
    ph.name, ph.op = "image", "Placeholder"
    const("conv0/W", rng.standard_normal((7, 7, 3, 64)).astype(np.float32), "content")
    const("group0/block0/conv1/bn/gamma", rng.standard_normal(64).astype(np.float32), "val")
    const("group0/block0/conv1/bn/beta", np.full(64, 0.5, np.float32), "splat")
    const("fp64/const", rng.standard_normal((3, 2)), "val")
    const("fp16/const", rng.standard_normal((4,)).astype(np.float16), "val")
    const("scalar", np.float32(3.25).reshape(()), "val")
    const("shape/const", np.array([1, -1, 7, 7], np.int32), "val")
    rd = g.node.add()
    rd.name, rd.op = "conv0/W/read", "Identity"
    rd.input.append("conv0/W")
    g.versions.producer = 27
    path = str(tmp_path / "official.pb")
    with open(path, "wb") as f:
        f.write(g.SerializeToString())
    got = read_frozen_graph(path, float_only=False)
    assert set(got) == set(want)
    for k, v in want.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape, k
        np.testing.assert_array_equal(got[k], v)
    assert "shape/const" not in read_frozen_graph(path)
