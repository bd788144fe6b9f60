"""Reads the weights out of a frozen TensorFlow GraphDef (`.pb`) without TensorFlow.

The reference packs a trained model with tf.graph_util.convert_variables_to_constants (models.py:134-191 `pack`) and loads
it back with tf.import_graph_def (`Mask_RCNN_FPN_frozen`, models.py:196-238).  Every variable becomes a `Const` node that
keeps the variable's name ("conv0/W", "group1/block0/conv2/bn/mean/EMA", ...), its value in the `value` attribute.  This
module walks the protobuf wire format directly -- GraphDef.node (1) -> NodeDef {name 1, op 2, attr 5 map<string, AttrValue>}
-> AttrValue.tensor (8) -> TensorProto {dtype 1, tensor_shape 2, tensor_content 4, float_val 5, double_val 6, int_val 7,
int64_val 10, half_val 13} (tensorflow/core/framework/{graph,node_def,attr_value,tensor,tensor_shape,types}.proto) -- and
returns {node name: ndarray} for the float Const nodes.  Host-side importer code: no device involved."""
from __future__ import annotations

import struct

import numpy as np

_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 19: np.float16, 10: np.bool_}


def _varint(buf, pos):
    x = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        x |= (b & 0x7F) << shift
        if not b & 0x80:
            return x, pos
        shift += 7


def _fields(buf):
    """Yields (field number, wire type, value) of one message; length-delimited values are memoryview slices."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fnum, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)
        yield fnum, wt, v


def _shape(buf):
    dims = []
    for f, wt, v in _fields(buf):
        if f == 2 and wt == 2:                       # Dim
            size = 0
            for f2, wt2, v2 in _fields(v):
                if f2 == 1 and wt2 == 0:
                    size = v2 if v2 < (1 << 63) else v2 - (1 << 64)
            dims.append(size)
    return dims


def _packed(v, wt, fmt, size):
    if wt == 2:                                       # packed repeated
        return list(struct.unpack("<%d%s" % (len(v) // size, fmt), bytes(v)))
    return [struct.unpack("<" + fmt, bytes(v))[0]]


def _tensor(buf):
    dtype, shape, content, vals = 0, [], None, []
    for f, wt, v in _fields(buf):
        if f == 1 and wt == 0:
            dtype = v
        elif f == 2 and wt == 2:
            shape = _shape(v)
        elif f == 4 and wt == 2:
            content = bytes(v)
        elif f == 5:
            vals += _packed(v, wt, "f", 4)
        elif f == 6:
            vals += _packed(v, wt, "d", 8)
        elif f in (7, 10, 13):                        # int_val / int64_val / half_val: varints, packed or not
            raw = []
            if wt == 2:
                p = 0
                while p < len(v):
                    x, p = _varint(v, p)
                    raw.append(x)
            else:
                raw.append(v)
            vals += [x - (1 << 64) if (f != 13 and x >= (1 << 63)) else x for x in raw]   # negatives are 64-bit two's complement
    np_dt = _DTYPES.get(dtype)
    if np_dt is None:
        return None
    n = int(np.prod(shape)) if shape else 1
    if content is not None and len(content):
        arr = np.frombuffer(content, dtype=np_dt).copy()
    elif vals:
        if dtype == 19:                               # half_val carries the raw 16-bit patterns
            arr = np.asarray(vals, dtype=np.uint16).view(np.float16)
        else:
            arr = np.asarray(vals, dtype=np_dt)
        if arr.size == 1 and n > 1:                   # a single value stands for a constant-filled tensor
            arr = np.full(n, arr[0], dtype=np_dt)
    else:
        arr = np.zeros(n, dtype=np_dt)
    if arr.size != n:
        raise ValueError("tensor with %d values for shape %s" % (arr.size, shape))
    return arr.reshape(shape)


def read_frozen_graph(path: str, float_only: bool = True) -> dict:
    """{Const node name: ndarray} of a frozen GraphDef file."""
    with open(path, "rb") as f:
        buf = memoryview(f.read())
    out = {}
    for f, wt, node in _fields(buf):
        if f != 1 or wt != 2:
            continue
        name, op, value = None, None, None
        for f2, wt2, v2 in _fields(node):
            if f2 == 1 and wt2 == 2:
                name = bytes(v2).decode()
            elif f2 == 2 and wt2 == 2:
                op = bytes(v2).decode()
            elif f2 == 5 and wt2 == 2:                # map entry {key 1, value 2}
                key, val = None, None
                for f3, wt3, v3 in _fields(v2):
                    if f3 == 1 and wt3 == 2:
                        key = bytes(v3).decode()
                    elif f3 == 2 and wt3 == 2:
                        val = v3
                if key == "value" and val is not None:
                    for f4, wt4, v4 in _fields(val):
                        if f4 == 8 and wt4 == 2:      # AttrValue.tensor
                            value = v4
        if op == "Const" and name and value is not None:
            arr = _tensor(value)
            if arr is not None and (not float_only or arr.dtype in (np.float32, np.float64, np.float16)):
                out[name] = arr
    return out
