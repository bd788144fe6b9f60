"""TEST INFRASTRUCTURE (CPU oracle; never imported by the product path).

cv2.resize(im, (neww, newh), interpolation=cv2.INTER_LINEAR) on a float32 HWC frame, restated in numpy from OpenCV's
published bilinear algorithm, plus the reference's own size rule.  Reference call sites: nn.py:1540-1560 (resizeImage /
get_new_hw), obj_detect_tracking.py:597-608, enqueuer_thread.py:259-266.  OpenCV itself is third-party (absent from
/root/reference); this restatement is pinned against the cv2 installed in the authoring container through
tests/golden/resize_cv2.npz (tests/golden/make_golden_tmot.py: resize_cv2)."""
import numpy as np


def get_new_hw(h, w, size, max_size):          # nn.py:1548-1560, verbatim arithmetic
    scale = size * 1.0 / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    return int(neww + 0.5), int(newh + 0.5)


def _taps(dst, src, reset_weight):
    scale = 1.0 / (float(dst) / float(src))
    f = ((np.arange(dst) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if reset_weight:                           # horizontal: border taps collapse to a single sample
        f = np.where((s < 0) | (s >= src - 1), np.float32(0), f)
        s = np.clip(s, 0, src - 1)
    i0 = np.clip(s, 0, src - 1)
    i1 = np.clip(s + 1, 0, src - 1)
    return i0, i1, (np.float32(1) - f).astype(np.float32), f.astype(np.float32)


def resize_linear(im, new_w, new_h):
    """im: [h, w, c] (any dtype, converted to float32 like `frame.astype("float32")`) -> [new_h, new_w, c] float32."""
    im = np.asarray(im, dtype=np.float32)
    h, w = im.shape[:2]
    x0, x1, a0, a1 = _taps(new_w, w, True)
    y0, y1, b0, b1 = _taps(new_h, h, False)
    rows = im[:, x0] * a0[None, :, None] + im[:, x1] * a1[None, :, None]          # horizontal pass (float32)
    return (rows[y0] * b0[:, None, None] + rows[y1] * b1[:, None, None]).astype(np.float32)


def resize_image(im, short_size, max_size):    # nn.py:1540-1545
    h, w = im.shape[:2]
    neww, newh = get_new_hw(h, w, short_size, max_size)
    if h == newh and w == neww:
        return np.asarray(im, dtype=np.float32)
    return resize_linear(im, neww, newh)
