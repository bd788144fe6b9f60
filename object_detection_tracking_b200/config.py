"""Model configuration namespace for the detect hot path.

Mirrors the derived-constant block of the reference CLI
(`obj_detect_tracking.py:236-389`): the reference passes its argparse namespace
straight into `models.get_model(config, ...)`, so the drop-in backend accepts an
object with the same attribute names.  `make_config()` builds such a namespace
without argparse; `normalize_config()` fills anything a caller left out.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

# obj_detect_tracking.py:327-387 -- fixed hyper-parameters of the inference graph
_DEFAULTS = dict(
    num_class=15,                      # obj_detect_tracking.py:122 (ActEV 15-class incl. BG)
    version=3,                         # --version 3  => use_dilations (:270-271)
    use_dilations=True,
    use_frcnn_class_agnostic=False,
    resnet_num_block=(3, 4, 23, 3),    # :348 ResNet-101
    short_edge_size=720,
    max_size=1280,
    rpn_test_post_nms_topk=300,
    rpn_min_size=0,
    rpn_proposal_nms_thres=0.7,
    anchor_strides=(4, 8, 16, 32, 64),
    anchor_sizes=(32, 64, 128, 256, 512),
    anchor_ratios=(0.5, 1.0, 2.0),
    fpn_num_channel=256,
    fpn_frcnn_fc_head_dim=1024,
    fastrcnn_bbox_reg_weights=(10.0, 10.0, 5.0, 5.0),
    fastrcnn_nms_iou_thres=0.5,
    threshold_conf=1e-4,               # -> result_score_thres (:385)
    result_per_im=100,
    im_batch_size=1,
    is_train=False,
    add_mask=False,
    mrcnn_head_dim=256,                # obj_detect_tracking.py:324
    use_partial_classes=False,
    is_efficientdet=False,
)


def make_config(**overrides) -> SimpleNamespace:
    cfg = dict(_DEFAULTS)
    cfg.update(overrides)
    return normalize_config(SimpleNamespace(**cfg))


def normalize_config(cfg):
    """Fill derived fields exactly as obj_detect_tracking.py:320-387 does."""
    for k, v in _DEFAULTS.items():
        if not hasattr(cfg, k):
            setattr(cfg, k, v)
    if getattr(cfg, "resnet50", False):
        cfg.resnet_num_block = (3, 4, 6, 3)          # :351-352
    cfg.resnet_num_block = tuple(int(x) for x in cfg.resnet_num_block)
    cfg.fpn_resolution_requirement = float(cfg.anchor_strides[3])   # :323
    # :330-331  max_size rounded up to a multiple of 32
    cfg.max_size = float(np.ceil(cfg.max_size / cfg.fpn_resolution_requirement)
                         * cfg.fpn_resolution_requirement)
    cfg.bbox_decode_clip = float(np.log(cfg.max_size / 16.0))        # :370
    cfg.result_score_thres = float(cfg.threshold_conf)               # :385
    cfg.num_anchors_per_loc = len(cfg.anchor_ratios)
    return cfg


def backbone_geometry(h: int, w: int, cfg):
    """Feature-map sizes for an (h, w) input, from the padding rules of
    `nn.py:843-944` (tf_pad_reverse=True) and the crop of `models.py:372-400`.

    Returns dict with c1, pool, c2..c5 (H, W), p2..p6 full sizes and the cropped
    p2..p6 sizes the RPN / ROIAlign see.
    """
    mult = int(cfg.fpn_resolution_requirement)
    ph = int(math.ceil(h / mult) * mult)
    pw = int(math.ceil(w / mult) * mult)
    # pad [3, 2 + pad_to_32], conv 7x7 s2 VALID
    ih, iw = ph + 5, pw + 5
    c1 = ((ih - 7) // 2 + 1, (iw - 7) // 2 + 1)
    # pad [1, 0], maxpool 3x3 s2 VALID
    pool = ((c1[0] + 1 - 3) // 2 + 1, (c1[1] + 1 - 3) // 2 + 1)
    c2 = pool
    def down(s):  # stride-2 group: pad [1,0] + 3x3 s2 VALID (dilated variant pads again)
        return ((s[0] + 1 - 3) // 2 + 1, (s[1] + 1 - 3) // 2 + 1)
    c3 = down(c2)
    c4 = down(c3)
    c5 = down(c4)
    full = [c2, c3, c4, c5, ((c5[0] - 1) // 2 + 1, (c5[1] - 1) // 2 + 1)]
    crop = []
    for i, s in enumerate(cfg.anchor_strides):
        if i < 3:
            crop.append((min(full[i][0], int(math.ceil(h / float(s)))),
                         min(full[i][1], int(math.ceil(w / float(s))))))
        else:
            crop.append(full[i])
    return dict(padded=(ph, pw), c1=c1, pool=pool, c=[c2, c3, c4, c5], p_full=full, p=crop)
