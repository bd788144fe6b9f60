"""Frame-resize ingest (reference: resizeImage / get_new_hw, nn.py:1540-1560): the oracle against cv2
fixtures (tests/golden/resize_cv2.npz, generated with the cv2 of the authoring container) and the CUDA kernel's per-pixel
arithmetic (csrc/resize_math.h, compiled here into a host harness) against the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import resize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_oracle_matches_cv2_fixtures(golden_dir):
    g = np.load(os.path.join(golden_dir, "resize_cv2.npz"))
    for k in range(6):
        nw, nh = (int(v) for v in g["size%d" % k])
        got = resize.resize_linear(g["src%d" % k], nw, nh)
        np.testing.assert_array_equal(got, g["dst%d" % k])        # bit-exact with OpenCV's generic bilinear path
        assert float(g["opt_dev%d" % k]) < 1e-2                   # what OpenCV's own SIMD path deviates by (0..255 scale)


def test_get_new_hw_known_sizes():
    from object_detection_tracking_b200.engine import get_new_hw
    for h, w, exp in ((1080, 1920, (1280, 720)), (720, 1280, (1280, 720)), (480, 640, (960, 720)), (1920, 1080, (720, 1280)),
                      (1000, 3000, (1280, 427)), (333, 500, (1081, 720))):
        assert get_new_hw(h, w, 720, 1280) == exp == resize.get_new_hw(h, w, 720, 1280)


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("resize") / "libresize_harness.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", out, os.path.join(ROOT, "tests", "resize_harness.cpp")], check=True)
    lib = ctypes.CDLL(out)
    lib.resize_harness.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    lib.resize_harness.restype = None
    return lib


def test_kernel_arithmetic_matches_oracle(harness, golden_dir):
    g = np.load(os.path.join(golden_dir, "resize_cv2.npz"))
    rng = np.random.default_rng(8)
    cases = [(g["src%d" % k], int(g["size%d" % k][0]), int(g["size%d" % k][1])) for k in range(6)]
    cases += [(rng.integers(0, 256, (h, w, 3)).astype(np.uint8), nw, nh)
              for h, w, nw, nh in ((108, 192, 128, 72), (33, 77, 231, 99), (5, 4, 40, 50), (50, 40, 4, 5), (1, 9, 3, 7))]
    for src, nw, nh in cases:
        src = np.ascontiguousarray(src)
        dst = np.empty((nh, nw, 3), np.float32)
        harness.resize_harness(src.ctypes.data, src.shape[0], src.shape[1], dst.ctypes.data, nh, nw)
        np.testing.assert_array_equal(dst, resize.resize_linear(src, nw, nh))
