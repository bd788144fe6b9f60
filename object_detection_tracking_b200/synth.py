"""Seeded synthetic weights / frames for the detector, in the reference's variable naming.

No model weights ship with the reference (download links only, README.md:83,105) and
this environment has no network, so parity and the bench run on seeded random
weights.  Names follow the Tensorpack-npz convention the reference loads
(`obj_detect_tracking.py:417-443`; scopes from `nn.py:886-1005`,
`models.py:984-1101`) so a real `.npz` drops in unchanged:

  conv0/W, conv0/bn/{gamma,beta,mean/EMA,variance/EMA}
  group{0..3}/block{i}/{conv1,conv2,conv3,convshortcut}/W (+ bn/..)
  fpn/lateral_1x1_c{2..5}/{W,b}, fpn/posthoc_3x3_p{2..5}/{W,b}
  rpn/conv0/{W,b}, rpn/class/{W,b}, rpn/box/{W,b}
  fastrcnn/fc6/{W,b}, fastrcnn/fc7/{W,b}, fastrcnn/outputs/{class,box}/{W,b}

Conv kernels are HWIO (`nn.py:350,367`), dense kernels [in,out] (`nn.py:753`).
The distributions are chosen so activations stay O(1) through 33 residual blocks
and the RPN / class logits have spread (otherwise every proposal collapses and
parity would be vacuous).
"""
from __future__ import annotations

import numpy as np


def _conv_w(rng, k, cin, cout, gain=2.0):
    std = np.sqrt(gain / (k * k * cin))
    return (rng.standard_normal((k, k, cin, cout)) * std).astype(np.float32)


def _bn(rng, w, name, c, gamma_lo=0.8, gamma_hi=1.2):
    w[name + "/bn/gamma"] = rng.uniform(gamma_lo, gamma_hi, c).astype(np.float32)
    w[name + "/bn/beta"] = (rng.standard_normal(c) * 0.05).astype(np.float32)
    w[name + "/bn/mean/EMA"] = (rng.standard_normal(c) * 0.05).astype(np.float32)
    w[name + "/bn/variance/EMA"] = rng.uniform(0.8, 1.2, c).astype(np.float32)


def synth_weights(cfg, seed: int = 1234) -> dict:
    rng = np.random.default_rng(seed)
    w = {}
    w["conv0/W"] = _conv_w(rng, 7, 3, 64)
    _bn(rng, w, "conv0", 64)
    cin = 64
    for g, (feat, count) in enumerate(zip((64, 128, 256, 512), cfg.resnet_num_block)):
        for i in range(count):
            p = "group%d/block%d" % (g, i)
            w[p + "/conv1/W"] = _conv_w(rng, 1, cin, feat)
            _bn(rng, w, p + "/conv1", feat)
            w[p + "/conv2/W"] = _conv_w(rng, 3, feat, feat)
            _bn(rng, w, p + "/conv2", feat)
            w[p + "/conv3/W"] = _conv_w(rng, 1, feat, feat * 4)
            # residual-branch last BN: small gamma keeps the running sum bounded
            _bn(rng, w, p + "/conv3", feat * 4, 0.1, 0.3)
            if cin != feat * 4:
                w[p + "/convshortcut/W"] = _conv_w(rng, 1, cin, feat * 4, gain=1.0)
                _bn(rng, w, p + "/convshortcut", feat * 4)
            cin = feat * 4
    nc = cfg.fpn_num_channel
    for i, c in enumerate((256, 512, 1024, 2048)):
        w["fpn/lateral_1x1_c%d/W" % (i + 2)] = _conv_w(rng, 1, c, nc, gain=0.25)
        w["fpn/lateral_1x1_c%d/b" % (i + 2)] = (rng.standard_normal(nc) * 0.02).astype(np.float32)
        w["fpn/posthoc_3x3_p%d/W" % (i + 2)] = _conv_w(rng, 3, nc, nc, gain=(0.12, 0.3, 0.8, 2.5)[i])
        w["fpn/posthoc_3x3_p%d/b" % (i + 2)] = (rng.standard_normal(nc) * 0.02).astype(np.float32)
    na = len(cfg.anchor_ratios)
    w["rpn/conv0/W"] = _conv_w(rng, 3, nc, nc, gain=1.0)
    w["rpn/conv0/b"] = (rng.standard_normal(nc) * 0.02).astype(np.float32)
    w["rpn/class/W"] = (rng.standard_normal((1, 1, nc, na)) * 0.12).astype(np.float32)
    w["rpn/class/b"] = (rng.standard_normal(na) * 0.1).astype(np.float32)
    w["rpn/box/W"] = (rng.standard_normal((1, 1, nc, 4 * na)) * 0.03).astype(np.float32)
    w["rpn/box/b"] = (rng.standard_normal(4 * na) * 0.02).astype(np.float32)
    dim = cfg.fpn_frcnn_fc_head_dim
    fin = nc * 7 * 7
    w["fastrcnn/fc6/W"] = (rng.standard_normal((fin, dim)) * np.sqrt(2.0 / fin)).astype(np.float32)
    w["fastrcnn/fc6/b"] = (rng.standard_normal(dim) * 0.02).astype(np.float32)
    w["fastrcnn/fc7/W"] = (rng.standard_normal((dim, dim)) * np.sqrt(2.0 / dim)).astype(np.float32)
    w["fastrcnn/fc7/b"] = (rng.standard_normal(dim) * 0.02).astype(np.float32)
    ncls = cfg.num_class
    w["fastrcnn/outputs/class/W"] = (rng.standard_normal((dim, ncls)) * 0.15).astype(np.float32)
    w["fastrcnn/outputs/class/b"] = (rng.standard_normal(ncls) * 0.1).astype(np.float32)
    nbox = ncls if not cfg.use_frcnn_class_agnostic else 1
    w["fastrcnn/outputs/box/W"] = (rng.standard_normal((dim, nbox * 4)) * 0.03).astype(np.float32)
    w["fastrcnn/outputs/box/b"] = (rng.standard_normal(nbox * 4) * 0.02).astype(np.float32)
    if getattr(cfg, "add_mask", False):              # maskrcnn_up4conv_head (models.py:1173-1199); drawn last so that the
        md = getattr(cfg, "mrcnn_head_dim", 256)     # detector weights of a seed do not depend on add_mask
        for k in range(4):
            w["maskrcnn/fcn%d/W" % k] = _conv_w(rng, 3, nc if k == 0 else md, md)
            w["maskrcnn/fcn%d/b" % k] = (rng.standard_normal(md) * 0.02).astype(np.float32)
        w["maskrcnn/deconv/W"] = (rng.standard_normal((2, 2, md, md)) * np.sqrt(2.0 / md)).astype(np.float32)   # [kh,kw,out,in]
        w["maskrcnn/deconv/b"] = (rng.standard_normal(md) * 0.02).astype(np.float32)
        w["maskrcnn/conv/W"] = (rng.standard_normal((1, 1, md, ncls - 1)) * np.sqrt(4.0 / md)).astype(np.float32)
        w["maskrcnn/conv/b"] = (rng.standard_normal(ncls - 1) * 0.5).astype(np.float32)
    return w


def frcnn_weight_shapes(cfg) -> dict:
    """Name -> shape of every variable the inference graph reads, in the reference's checkpoint naming (Tensorpack-style:
    conv2d `W`/`b` nn.py:367,376, BatchNorm `gamma`/`beta`/`mean/EMA`/`variance/EMA` nn.py:1739-1826, dense nn.py:762;
    scopes: nn.py:843-944 resnet, :947-1014 fpn, models.py:979-1009 rpn, :1030-1108 fastrcnn).  Used by the importer
    (backend.check_weights) to refuse incomplete or mis-shaped checkpoints before anything reaches the device."""
    sh = {}

    def conv(name, k, cin, cout, bias=False):
        sh[name + "/W"] = (k, k, cin, cout)
        if bias:
            sh[name + "/b"] = (cout,)

    def bn(name, c):
        for k in ("gamma", "beta", "mean/EMA", "variance/EMA"):
            sh["%s/bn/%s" % (name, k)] = (c,)

    conv("conv0", 7, 3, 64)
    bn("conv0", 64)
    cin = 64
    for g, (feat, count) in enumerate(zip((64, 128, 256, 512), cfg.resnet_num_block)):
        for i in range(count):
            p = "group%d/block%d" % (g, i)
            for name, k, ci, co in (("conv1", 1, cin, feat), ("conv2", 3, feat, feat), ("conv3", 1, feat, feat * 4)):
                conv("%s/%s" % (p, name), k, ci, co)
                bn("%s/%s" % (p, name), co)
            if cin != feat * 4:
                conv(p + "/convshortcut", 1, cin, feat * 4)
                bn(p + "/convshortcut", feat * 4)
            cin = feat * 4
    nc = cfg.fpn_num_channel
    for i, c in enumerate((256, 512, 1024, 2048)):
        conv("fpn/lateral_1x1_c%d" % (i + 2), 1, c, nc, bias=True)
        conv("fpn/posthoc_3x3_p%d" % (i + 2), 3, nc, nc, bias=True)
    na = len(cfg.anchor_ratios)
    conv("rpn/conv0", 3, nc, nc, bias=True)
    conv("rpn/class", 1, nc, na, bias=True)
    conv("rpn/box", 1, nc, 4 * na, bias=True)
    dim = cfg.fpn_frcnn_fc_head_dim
    nbox = 1 if cfg.use_frcnn_class_agnostic else cfg.num_class
    for name, fin, fout in (("fastrcnn/fc6", nc * 49, dim), ("fastrcnn/fc7", dim, dim),
                            ("fastrcnn/outputs/class", dim, cfg.num_class), ("fastrcnn/outputs/box", dim, nbox * 4)):
        sh[name + "/W"] = (fin, fout)
        sh[name + "/b"] = (fout,)
    if getattr(cfg, "add_mask", False):              # models.py:1173-1199; Conv2DTranspose kernel is [kh, kw, out, in]
        md = getattr(cfg, "mrcnn_head_dim", 256)
        for k in range(4):
            conv("maskrcnn/fcn%d" % k, 3, nc if k == 0 else md, md, bias=True)
        sh["maskrcnn/deconv/W"] = (2, 2, md, md)
        sh["maskrcnn/deconv/b"] = (md,)
        conv("maskrcnn/conv", 1, md, cfg.num_class - 1, bias=True)
    return sh


def synth_frame(h: int, w: int, seed: int = 0, n_rects: int = 4) -> np.ndarray:
    """uint8 [h, w, 3] BGR frame: low-pass noise + a few flat rectangles + white noise
    (SURVEY.md section 8d), so RPN logits have spatial structure."""
    rng = np.random.default_rng(seed)
    gh, gw = max(2, h // 16), max(2, w // 16)
    coarse = rng.uniform(0, 255, (gh, gw, 3)).astype(np.float32)
    # separable bilinear up-sampling of the coarse grid (no cv2 dependency)
    ys = np.linspace(0, gh - 1, h, dtype=np.float32)
    xs = np.linspace(0, gw - 1, w, dtype=np.float32)
    y0 = np.floor(ys).astype(np.int64); y1 = np.minimum(y0 + 1, gh - 1); fy = (ys - y0)[:, None, None]
    x0 = np.floor(xs).astype(np.int64); x1 = np.minimum(x0 + 1, gw - 1); fx = (xs - x0)[None, :, None]
    rows = coarse[y0] * (1 - fy) + coarse[y1] * fy
    img = rows[:, x0] * (1 - fx) + rows[:, x1] * fx
    for _ in range(n_rects):
        rh, rw = int(rng.integers(h // 8, h // 2)), int(rng.integers(w // 10, w // 3))
        ry, rx = int(rng.integers(0, h - rh)), int(rng.integers(0, w - rw))
        img[ry:ry + rh, rx:rx + rw] = rng.uniform(0, 255, 3).astype(np.float32)
    img += rng.uniform(-10, 10, img.shape).astype(np.float32)
    return np.clip(img, 0, 255).astype(np.uint8)


def osnet_param_shapes() -> dict:
    """Parameter / buffer names and shapes of torchreid's osnet_x1_0 state_dict (osnet.py:282-438, :522-534),
    minus the unused classifier."""
    shapes = {}

    def bn(prefix, c):
        for k in ("weight", "bias", "running_mean", "running_var"):
            shapes[prefix + "." + k] = (c,)

    def light(prefix, c):
        shapes[prefix + ".conv1.weight"] = (c, c, 1, 1)
        shapes[prefix + ".conv2.weight"] = (c, 1, 3, 3)
        bn(prefix + ".bn", c)

    def osblock(prefix, cin, cout):
        mid = cout // 4
        shapes[prefix + ".conv1.conv.weight"] = (mid, cin, 1, 1)
        bn(prefix + ".conv1.bn", mid)
        light(prefix + ".conv2a", mid)
        for name, n in (("conv2b", 2), ("conv2c", 3), ("conv2d", 4)):
            for j in range(n):
                light("%s.%s.%d" % (prefix, name, j), mid)
        shapes[prefix + ".gate.fc1.weight"] = (mid // 16, mid, 1, 1)
        shapes[prefix + ".gate.fc1.bias"] = (mid // 16,)
        shapes[prefix + ".gate.fc2.weight"] = (mid, mid // 16, 1, 1)
        shapes[prefix + ".gate.fc2.bias"] = (mid,)
        shapes[prefix + ".conv3.conv.weight"] = (cout, mid, 1, 1)
        bn(prefix + ".conv3.bn", cout)
        if cin != cout:
            shapes[prefix + ".downsample.conv.weight"] = (cout, cin, 1, 1)
            bn(prefix + ".downsample.bn", cout)

    shapes["conv1.conv.weight"] = (64, 3, 7, 7)
    bn("conv1.bn", 64)
    chans = (64, 256, 384, 512)
    for s in range(3):
        name = "conv%d" % (s + 2)
        osblock(name + ".0", chans[s], chans[s + 1])
        osblock(name + ".1", chans[s + 1], chans[s + 1])
        if s < 2:
            shapes[name + ".2.0.conv.weight"] = (chans[s + 1], chans[s + 1], 1, 1)
            bn(name + ".2.0.bn", chans[s + 1])
    shapes["conv5.conv.weight"] = (512, 512, 1, 1)
    bn("conv5.bn", 512)
    shapes["fc.0.weight"] = (512, 512)
    shapes["fc.0.bias"] = (512,)
    bn("fc.1", 512)
    return shapes


def synth_osnet_state(seed: int = 4321) -> dict:
    """Seeded float32 state_dict for osnet_x1_0 with non-trivial BatchNorm statistics (a fresh torch model has
    mean 0 / var 1 / gamma 1, which would not exercise the BN folding)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shp in osnet_param_shapes().items():
        if name.endswith("running_var"):
            v = rng.uniform(0.6, 1.4, shp)
        elif name.endswith("running_mean"):
            v = rng.standard_normal(shp) * 0.1
        elif name.endswith("conv3.bn.weight"):              # residual-branch BN: small gamma keeps the sum bounded
            v = rng.uniform(0.1, 0.3, shp)
        elif name.endswith("bn.weight") or name == "fc.1.weight":
            v = rng.uniform(0.7, 1.3, shp)
        elif name.endswith("bias"):
            v = rng.standard_normal(shp) * 0.05
        elif len(shp) == 4 and shp[1] == 1:                 # depthwise 3x3: fan_in 9
            v = rng.standard_normal(shp) * np.sqrt(1.0 / 9)
        elif len(shp) >= 2:
            fan_in = int(np.prod(shp[1:]))
            relu_follows = name.endswith("conv1.conv.weight") or name.endswith(".2.0.conv.weight") \
                or name in ("conv5.conv.weight", "fc.0.weight")
            v = rng.standard_normal(shp) * np.sqrt((2.0 if relu_follows else 1.0) / fan_in)
        else:
            v = rng.standard_normal(shp) * 0.05
        out[name] = v.astype(np.float32)
    return out


def resnet101_reid_param_shapes() -> dict:
    """Parameter / buffer names and shapes of torchreid's resnet101 state_dict (torchreid/models/resnet.py:157-290,
    441-455: Bottleneck [3, 4, 23, 3], last_stride 2), minus the unused classifier."""
    shapes = {"conv1.weight": (64, 3, 7, 7)}

    def bn(prefix, c):
        for k in ("weight", "bias", "running_mean", "running_var"):
            shapes[prefix + "." + k] = (c,)

    bn("bn1", 64)
    cin = 64
    for li, (n, planes) in enumerate(zip((3, 4, 23, 3), (64, 128, 256, 512))):
        for bi in range(n):
            p = "layer%d.%d" % (li + 1, bi)
            shapes[p + ".conv1.weight"] = (planes, cin, 1, 1)
            bn(p + ".bn1", planes)
            shapes[p + ".conv2.weight"] = (planes, planes, 3, 3)
            bn(p + ".bn2", planes)
            shapes[p + ".conv3.weight"] = (planes * 4, planes, 1, 1)
            bn(p + ".bn3", planes * 4)
            if bi == 0:
                shapes[p + ".downsample.0.weight"] = (planes * 4, cin, 1, 1)
                bn(p + ".downsample.1", planes * 4)
            cin = planes * 4
    return shapes


def synth_resnet101_reid_state(seed: int = 2468) -> dict:
    """Seeded float32 state_dict for torchreid's resnet101 with non-trivial BatchNorm statistics; the last BN of every
    residual branch gets a small gamma so that 33 stacked blocks keep activations O(1) (fp16-plane range)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shp in resnet101_reid_param_shapes().items():
        if name.endswith("running_var"):
            v = rng.uniform(0.6, 1.4, shp)
        elif name.endswith("running_mean"):
            v = rng.standard_normal(shp) * 0.1
        elif name.endswith("bn3.weight"):
            v = rng.uniform(0.1, 0.3, shp)
        elif name.endswith(".weight") and len(shp) == 1:
            v = rng.uniform(0.7, 1.3, shp)
        elif name.endswith("bias"):
            v = rng.standard_normal(shp) * 0.05
        else:
            fan_in = int(np.prod(shp[1:]))
            gain = 1.0 if (name.endswith("conv3.weight") or "downsample" in name) else 2.0
            v = rng.standard_normal(shp) * np.sqrt(gain / fan_in)
        out[name] = v.astype(np.float32)
    return out


def synth_effdet_weights(cfg, seed: int = 99) -> dict:
    """Seeded weights of the EfficientDet feature network + class/box nets in the TF variable naming of
    efficientdet_arch.py (kernels HWIO, depthwise [3,3,C,1]); BN statistics non-trivial; predict biases spread
    so that the top-k / NMS path sees a non-degenerate score distribution."""
    from .effdet_config import BIFPN_NODES
    rng = np.random.default_rng(seed)
    W = {}
    F_ = cfg.fpn_num_filters

    def bn(name, c):
        W[name + "/gamma"] = rng.uniform(0.7, 1.3, c).astype(np.float32)
        W[name + "/beta"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        W[name + "/moving_mean"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        W[name + "/moving_variance"] = rng.uniform(0.6, 1.4, c).astype(np.float32)

    def conv(name, cin, cout):
        W[name + "/kernel"] = (rng.standard_normal((1, 1, cin, cout)) * np.sqrt(1.0 / cin)).astype(np.float32)
        W[name + "/bias"] = (rng.standard_normal(cout) * 0.05).astype(np.float32)

    def sep(name, c, cout, gain=1.6, bias_std=0.05):
        W[name + "/depthwise_kernel"] = (rng.standard_normal((3, 3, c, 1)) * np.sqrt(1.0 / 9)).astype(np.float32)
        W[name + "/pointwise_kernel"] = (rng.standard_normal((1, 1, c, cout)) * np.sqrt(gain / c)).astype(np.float32)
        W[name + "/bias"] = (rng.standard_normal(cout) * bias_std).astype(np.float32)

    node_gain_sum = 0.55     # un-normalised "sum" nodes add 2-3 maps: smaller pointwise gain keeps 8 cells O(1)
    c5 = cfg.backbone_channels[2]
    if c5 != F_:
        conv("resample_p6/conv2d", c5, F_)
        bn("resample_p6/bn", F_)
    chans = list(cfg.backbone_channels) + [F_, F_]
    for rep in range(cfg.fpn_cell_repeats):
        cur = list(chans) if rep == 0 else [F_] * 5
        for i, (lvl, offsets) in enumerate(BIFPN_NODES):
            pre = "fpn_cells/cell_%d/fnode%d" % (rep, i)
            for idx, o in enumerate(offsets):
                if cur[o] != F_:
                    nm = "%s/resample_%d_%d_%d" % (pre, idx, o, len(cur))
                    conv(nm + "/conv2d", cur[o], F_)
                    bn(nm + "/bn", F_)
                if cfg.fpn_weight_method == "fastattn":
                    W["%s/WSM%s" % (pre, "" if idx == 0 else "_%d" % idx)] = np.float32(rng.uniform(0.5, 1.5))
            op = "%s/op_after_combine%d" % (pre, len(cur))
            sep(op + "/conv", F_, F_, gain=3.0 if cfg.fpn_weight_method == "fastattn" else node_gain_sum)
            bn(op + "/bn", F_)
            cur.append(F_)
    na = cfg.num_scales * len(cfg.aspect_ratios)
    for kind, nout in (("class", cfg.num_classes * na), ("box", 4 * na)):
        for i in range(cfg.box_class_repeats):
            sep("%s_net/%s-%d" % (kind, kind, i), F_, F_, gain=3.0)
            for level in range(cfg.min_level, cfg.max_level + 1):
                bn("%s_net/%s-%d-bn-%d" % (kind, kind, i, level), F_)
        sep("%s_net/%s-predict" % (kind, kind), F_, nout, gain=6.0 if kind == "class" else 0.05,
            bias_std=0.5 if kind == "class" else 0.02)
    if "class_net/class-predict/bias" in W:
        W["class_net/class-predict/bias"] -= np.float32(3.0)     # mostly-negative logits, like a trained detector
    return W



def synth_efficientnet_weights(name: str, seed: int = 77) -> dict:
    """Seeded EfficientNet backbone weights in the keras variable naming of the reference graph
    (efficientnet_model.py:162-392, 504-704); gains chosen so that activations stay O(1) through all blocks."""
    from .effdet_config import efficientnet_blocks
    rng = np.random.default_rng(seed)
    stem_c, blocks = efficientnet_blocks(name)
    W = {}

    def bn(pre, c, g_lo=0.8, g_hi=1.2):
        W[pre + "/gamma"] = rng.uniform(g_lo, g_hi, c).astype(np.float32)
        W[pre + "/beta"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        W[pre + "/moving_mean"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        W[pre + "/moving_variance"] = rng.uniform(0.6, 1.4, c).astype(np.float32)

    def conv(pre, k, cin, cout, gain, bias=None):
        W[pre + "/kernel"] = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(gain / (k * k * cin))).astype(np.float32)
        if bias is not None:
            W[pre + "/bias"] = (rng.standard_normal(cout) * 0.1 + bias).astype(np.float32)

    conv(name + "/stem/conv2d", 3, 3, stem_c, 1.0)
    bn(name + "/stem/tpu_batch_normalization", stem_c)
    for i, b in enumerate(blocks):
        pre = "%s/blocks_%d" % (name, i)
        mid = b.cin * b.expand
        proj = "conv2d"
        if b.expand != 1:
            conv(pre + "/conv2d", 1, b.cin, mid, 2.5)
            bn(pre + "/tpu_batch_normalization", mid)
            proj = "conv2d_1"
        W[pre + "/depthwise_conv2d/depthwise_kernel"] = (
            rng.standard_normal((b.kernel, b.kernel, mid, 1)) * np.sqrt(2.5 / (b.kernel ** 2))).astype(np.float32)
        bn(pre + "/tpu_batch_normalization_1", mid)
        nr = max(1, int(b.cin * b.se))
        conv(pre + "/se/conv2d", 1, mid, nr, 1.0, bias=0.0)
        conv(pre + "/se/conv2d_1", 1, nr, mid, 1.0, bias=0.5)
        conv(pre + "/" + proj, 1, mid, b.cout, 2.0)
        bn(pre + "/tpu_batch_normalization_2", b.cout, 0.45, 0.75)
    return W
