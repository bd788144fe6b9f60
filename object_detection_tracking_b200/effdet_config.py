"""EfficientDet configuration mirror (efficientdet_wrapper.py:160-252 `get_efficientdet_config`, parameter
table :511-587) for the BiFPN / heads / post-processing path."""
from __future__ import annotations

from types import SimpleNamespace

# name -> (fpn_num_filters, fpn_cell_repeats, box_class_repeats, anchor_scale, weight_method, backbone C3/C4/C5 channels)
_TABLE = {
    "efficientdet-d0": (64, 3, 3, 4.0, "fastattn", (40, 112, 320)),
    "efficientdet-d1": (88, 4, 3, 4.0, "fastattn", (40, 112, 320)),
    "efficientdet-d2": (112, 5, 3, 4.0, "fastattn", (48, 120, 352)),
    "efficientdet-d3": (160, 6, 4, 4.0, "fastattn", (48, 136, 384)),
    "efficientdet-d4": (224, 7, 4, 4.0, "fastattn", (56, 160, 448)),
    "efficientdet-d5": (288, 7, 4, 4.0, "fastattn", (64, 176, 512)),
    "efficientdet-d6": (384, 8, 5, 4.0, "sum", (72, 200, 576)),
    "efficientdet-d7": (384, 8, 5, 5.0, "sum", (72, 200, 576)),
}

# efficientdet_arch.py:508-522 bifpn_sum_config / bifpn_fa_config (levels 3..7)
BIFPN_NODES = (
    (6, (3, 4)), (5, (2, 5)), (4, (1, 6)), (3, (0, 7)),
    (4, (1, 7, 8)), (5, (2, 6, 9)), (6, (3, 5, 10)), (7, (4, 11)),
)


def make_effdet_config(name="efficientdet-d7", height=1536, width=1536, **overrides) -> SimpleNamespace:
    w, cells, reps, ascale, method, bc = _TABLE[name]
    cfg = SimpleNamespace(
        name=name, image_size=(int(height), int(width)), min_level=3, max_level=7,
        num_classes=90, num_scales=3, aspect_ratios=((1.0, 1.0), (1.4, 0.7), (0.7, 1.4)), anchor_scale=ascale,
        fpn_num_filters=w, fpn_cell_repeats=cells, box_class_repeats=reps, fpn_weight_method=method,
        backbone_channels=bc, apply_bn_for_resampling=True, conv_after_downsample=False,
        max_detection_topk=5000, result_score_thres=1e-4, result_per_im=100, nms_iou_threshold=0.5,
    )
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def feat_sizes(cfg):
    """efficientdet/utils.py:467-484 get_feat_sizes: (h-1)//2+1 per level, index = level."""
    h, w = cfg.image_size
    out = [(h, w)]
    for _ in range(1, cfg.max_level + 1):
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        out.append((h, w))
    return out


# EfficientNet backbone geometry (efficientnet_builder.py:37-52,172-178; efficientnet_model.py:137-159,535-584)
_EFFNET = {"efficientnet-b0": (1.0, 1.0), "efficientnet-b1": (1.0, 1.1), "efficientnet-b2": (1.1, 1.2),
           "efficientnet-b3": (1.2, 1.4), "efficientnet-b4": (1.4, 1.8), "efficientnet-b5": (1.6, 2.2),
           "efficientnet-b6": (1.8, 2.6), "efficientnet-b7": (2.0, 3.1)}
# (repeats, kernel, stride, expand, in, out, se_ratio)
_EFFNET_STAGES = ((1, 3, 1, 1, 32, 16, 0.25), (2, 3, 2, 6, 16, 24, 0.25), (2, 5, 2, 6, 24, 40, 0.25),
                  (3, 3, 2, 6, 40, 80, 0.25), (3, 5, 1, 6, 80, 112, 0.25), (4, 5, 2, 6, 112, 192, 0.25),
                  (1, 3, 1, 6, 192, 320, 0.25))
# efficientdet_wrapper.py:511-587: detector -> backbone
BACKBONE_OF = {"efficientdet-d0": "efficientnet-b0", "efficientdet-d1": "efficientnet-b1", "efficientdet-d2": "efficientnet-b2",
               "efficientdet-d3": "efficientnet-b3", "efficientdet-d4": "efficientnet-b4", "efficientdet-d5": "efficientnet-b5",
               "efficientdet-d6": "efficientnet-b6", "efficientdet-d7": "efficientnet-b6"}


def _round_filters(filters, width, divisor=8):
    filters *= width
    new = max(divisor, int(filters + divisor / 2) // divisor * divisor)
    if new < 0.9 * filters:
        new += divisor
    return int(new)


def efficientnet_blocks(name):
    """-> (stem channels, [block(kernel, stride, expand, cin, cout, se)]) after width / depth scaling."""
    import math
    width, depth = _EFFNET[name]
    out = []
    for rep, k, s, e, ci, co, se in _EFFNET_STAGES:
        ci, co, rep = _round_filters(ci, width), _round_filters(co, width), int(math.ceil(depth * rep))
        out.append(SimpleNamespace(kernel=k, stride=s, expand=e, cin=ci, cout=co, se=se))
        for _ in range(rep - 1):
            out.append(SimpleNamespace(kernel=k, stride=1, expand=e, cin=co, cout=co, se=se))
    return _round_filters(32, width), out
