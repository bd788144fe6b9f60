"""ORACLE (test infrastructure only).  CPU fp32 restatement of the EfficientNet backbone the reference's EfficientDet
wrapper builds (efficientdet_arch.build_backbone :396-437) and of the wrapper's input pre-processing.

Reference: efficientdet/backbone/efficientnet_builder.py (`efficientnet_params` :37-52, `_DEFAULT_BLOCKS_ARGS`
:172-178, BlockDecoder :55-120), efficientnet_model.py (`round_filters` :137-151, `round_repeats` :154-159,
`MBConvBlock` :162-392 -- expand 1x1 + BN + swish, depthwise kxk + BN + swish, squeeze-excite (reduce_mean -> 1x1 +
swish -> 1x1 -> sigmoid gate), project 1x1 + BN, identity skip; `Model._build/call` :504-704 -- stem 3x3/2 + BN +
swish, endpoints reduction_1..5), efficientdet_wrapper.py `build_preprocess` (:40-61) with dataloader.py
`InputProcessor` (:56-123: convert_image_dtype, mean/std, bilinear resize_images, pad_to_bounding_box).
TensorFlow is absent here => **parity unpinned** for the TF op arithmetic (Conv2D/DepthwiseConv2D SAME padding,
legacy ResizeBilinear follow TF's documented semantics); the filter/repeat rounding is pinned against the reference's
own `round_filters`/`round_repeats` (tests/golden/effdet_numpy.npz).

Weight naming (keras variable names under the model scope, e.g. "efficientnet-b6/"):
  stem/conv2d/kernel, stem/tpu_batch_normalization/{gamma,beta,moving_mean,moving_variance}
  blocks_i/conv2d/kernel              expand (expand_ratio != 1) -- or the projection when expand_ratio == 1
  blocks_i/tpu_batch_normalization    bn0 (after expand)
  blocks_i/depthwise_conv2d/depthwise_kernel, blocks_i/tpu_batch_normalization_1
  blocks_i/se/conv2d/{kernel,bias}, blocks_i/se/conv2d_1/{kernel,bias}
  blocks_i/conv2d_1/kernel            projection (expand_ratio != 1), blocks_i/tpu_batch_normalization_2
"""
from __future__ import annotations

import math
from collections import namedtuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-3
PARAMS = {   # (width, depth) efficientnet_builder.py:37-52
    "efficientnet-b0": (1.0, 1.0), "efficientnet-b1": (1.0, 1.1), "efficientnet-b2": (1.1, 1.2),
    "efficientnet-b3": (1.2, 1.4), "efficientnet-b4": (1.4, 1.8), "efficientnet-b5": (1.6, 2.2),
    "efficientnet-b6": (1.8, 2.6), "efficientnet-b7": (2.0, 3.1),
}
# (repeats, kernel, stride, expand, in, out, se_ratio) efficientnet_builder.py:172-178
DEFAULT_BLOCKS = ((1, 3, 1, 1, 32, 16, 0.25), (2, 3, 2, 6, 16, 24, 0.25), (2, 5, 2, 6, 24, 40, 0.25),
                  (3, 3, 2, 6, 40, 80, 0.25), (3, 5, 1, 6, 80, 112, 0.25), (4, 5, 2, 6, 112, 192, 0.25),
                  (1, 3, 1, 6, 192, 320, 0.25))
Block = namedtuple("Block", "kernel stride expand cin cout se")


def round_filters(filters, width, divisor=8):
    """efficientnet_model.py:137-151."""
    filters *= width
    new = max(divisor, int(filters + divisor / 2) // divisor * divisor)
    if new < 0.9 * filters:
        new += divisor
    return int(new)


def round_repeats(repeats, depth):
    """efficientnet_model.py:154-159."""
    return int(math.ceil(depth * repeats))


def block_specs(name, blocks=DEFAULT_BLOCKS):
    """Model._build (:535-584): per-stage filter / repeat rounding, first block of a stage carries the stride."""
    width, depth = PARAMS[name]
    out = []
    for rep, k, s, e, ci, co, se in blocks:
        ci, co, rep = round_filters(ci, width), round_filters(co, width), round_repeats(rep, depth)
        out.append(Block(k, s, e, ci, co, se))
        for _ in range(rep - 1):
            out.append(Block(k, 1, e, co, co, se))
    return round_filters(32, width), out


def endpoint_blocks(blocks):
    """Model.call (:644-668): reduction_k = output of the last block before the next stride-2 block (and the last)."""
    red = []
    for i in range(len(blocks)):
        if i == len(blocks) - 1 or blocks[i + 1].stride > 1:
            red.append(i)
    return red          # red[k-1] = block index of reduction_k


def _t(a):
    """float32 by default; float64 inside `with oracle.frcnn.exact():` (the rounding-noise-free evaluation, see frcnn.py)."""
    from .frcnn import _DT
    t = torch.from_numpy(np.ascontiguousarray(a))
    return t.to(_DT[-1]) if t.is_floating_point() else t


def _np(t):
    return t.to(torch.float32).numpy()


def swish(x):
    return x * torch.sigmoid(x)


def bn(x, W, name):
    g, b = _t(W[name + "/gamma"]), _t(W[name + "/beta"])
    m, v = _t(W[name + "/moving_mean"]), _t(W[name + "/moving_variance"])
    inv = torch.rsqrt(v + BN_EPS) * g
    return x * inv.view(1, -1, 1, 1) + (b - m * inv).view(1, -1, 1, 1)


def same_pad(x, k, s):
    """TF 'SAME': total = max((ceil(n/s)-1)*s + k - n, 0), before = total // 2 (the extra pixel goes after)."""
    h, w = x.shape[2:]
    ph = max((-(-h // s) - 1) * s + k - h, 0)
    pw = max((-(-w // s) - 1) * s + k - w, 0)
    return F.pad(x, (pw // 2, pw - pw // 2, ph // 2, ph - ph // 2))


def conv(x, W, name, stride=1, bias=False):
    k = _t(W[name + "/kernel"])                       # HWIO
    x = same_pad(x, k.shape[0], stride)
    return F.conv2d(x, k.permute(3, 2, 0, 1).contiguous(), _t(W[name + "/bias"]) if bias else None, stride=stride)


def dwconv(x, W, name, stride):
    k = _t(W[name + "/depthwise_kernel"])             # [k,k,C,1]
    x = same_pad(x, k.shape[0], stride)
    return F.conv2d(x, k.permute(2, 3, 0, 1).contiguous(), None, stride=stride, groups=k.shape[2])


def mbconv(x, W, pre, b: Block):
    """MBConvBlock.call (:331-392), inference (no drop-connect)."""
    inp = x
    proj = "conv2d"
    if b.expand != 1:
        x = swish(bn(conv(x, W, pre + "/conv2d"), W, pre + "/tpu_batch_normalization"))
        proj = "conv2d_1"
    x = swish(bn(dwconv(x, W, pre + "/depthwise_conv2d", b.stride), W, pre + "/tpu_batch_normalization_1"))
    if b.se and 0 < b.se <= 1:
        s = x.mean(dim=(2, 3), keepdim=True)
        s = conv(swish(conv(s, W, pre + "/se/conv2d", bias=True)), W, pre + "/se/conv2d_1", bias=True)
        x = torch.sigmoid(s) * x
    x = bn(conv(x, W, pre + "/" + proj), W, pre + "/tpu_batch_normalization_2")
    if b.stride == 1 and b.cin == b.cout:
        x = x + inp
    return x


def forward(image_nhwc, W, name, stages=False):
    """image [H,W,3] fp32 (pre-processed) -> {3,4,5: [C,h,w]} numpy (reduction_3/4/5)."""
    stem_c, blocks = block_specs(name)
    red = endpoint_blocks(blocks)
    out, every = {}, {}
    with torch.no_grad():
        x = _t(np.asarray(image_nhwc, np.float32)).permute(2, 0, 1)[None]
        x = swish(bn(conv(x, W, name + "/stem/conv2d", stride=2), W, name + "/stem/tpu_batch_normalization"))
        every["stem"] = _np(x[0])
        for i, b in enumerate(blocks):
            x = mbconv(x, W, "%s/blocks_%d" % (name, i), b)
            if stages:
                every["block_%d" % i] = _np(x[0])
            if i in red:
                out[red.index(i) + 1] = _np(x[0])
    res = {l: out[l] for l in (3, 4, 5)}
    if stages:
        res["stages"] = every
    return res


def resize_bilinear_legacy(img, oh, ow):
    """tf.image.resize_images(BILINEAR) of TF1: align_corners=False, half_pixel_centers=False.  img [H,W,C] fp32."""
    f32 = np.float32
    h, w = img.shape[:2]
    ys = (np.arange(oh, dtype=f32) * (f32(h) / f32(oh))).astype(f32)
    xs = (np.arange(ow, dtype=f32) * (f32(w) / f32(ow))).astype(f32)
    y0 = np.floor(ys).astype(np.int64); y1 = np.minimum(np.ceil(ys).astype(np.int64), h - 1); ly = (ys - np.floor(ys)).astype(f32)
    x0 = np.floor(xs).astype(np.int64); x1 = np.minimum(np.ceil(xs).astype(np.int64), w - 1); lx = (xs - np.floor(xs)).astype(f32)
    tl, tr = img[y0][:, x0], img[y0][:, x1]
    bl, br = img[y1][:, x0], img[y1][:, x1]
    lx_, ly_ = lx[None, :, None], ly[:, None, None]
    top = (tl + ((tr - tl).astype(f32) * lx_).astype(f32)).astype(f32)
    bot = (bl + ((br - bl).astype(f32) * lx_).astype(f32)).astype(f32)
    return (top + ((bot - top).astype(f32) * ly_).astype(f32)).astype(f32)


def preprocess(frame_bgr_u8, out_h, out_w):
    """EfficientDet.build_preprocess (wrapper :40-61): BGR->RGB, /255, mean/std, resize to fit, zero-pad bottom/right.
    Returns (image [out_h,out_w,3] fp32, image_scale_to_original)."""
    f32 = np.float32
    rgb = frame_bgr_u8[:, :, ::-1].astype(f32) * f32(1.0 / 255.0)            # convert_image_dtype
    rgb = ((rgb - np.array([0.485, 0.456, 0.406], f32)) / np.array([0.229, 0.224, 0.225], f32)).astype(f32)
    h, w = frame_bgr_u8.shape[:2]
    scale = min(f32(out_w) / f32(w), f32(out_h) / f32(h))                   # set_scale_factors_to_output_size
    sh, sw = int(f32(h) * scale), int(f32(w) * scale)
    img = resize_bilinear_legacy(rgb, sh, sw)[:out_h, :out_w]
    out = np.zeros((out_h, out_w, 3), f32)
    out[:img.shape[0], :img.shape[1]] = img
    return out, f32(1.0) / scale
