"""GPU probe: run the detector on a small frame and compare every stage with the CPU oracle.
Usage: python tools/gpu_pipeline_probe.py H W impl precision [blocks] [batch]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from object_detection_tracking_b200.config import make_config  # noqa: E402
from object_detection_tracking_b200.engine import Detector  # noqa: E402
from object_detection_tracking_b200.synth import synth_frame, synth_weights  # noqa: E402
from oracle import frcnn  # noqa: E402  (probe tool = test infrastructure)


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-12))


def main():
    H, W = int(sys.argv[1]), int(sys.argv[2])
    impl, prec = sys.argv[3], sys.argv[4]
    blocks = tuple(int(v) for v in sys.argv[5].split(",")) if len(sys.argv) > 5 else (3, 4, 23, 3)
    B = int(sys.argv[6]) if len(sys.argv) > 6 else 1
    cfg = make_config(resnet_num_block=blocks, max_size=max(H, W), short_edge_size=min(H, W))
    Wt = synth_weights(cfg, 1234)
    frames = np.stack([synth_frame(H, W, seed=s) for s in range(B)]).astype(np.float32)
    det = Detector(cfg, B, H, W, precision=prec, conv_impl=impl, use_cuda_graph=False)
    det.load_weights(Wt)
    det.set_stage("image", frames)
    t = time.time(); det.run_phases(255); print("gpu pass %.1f ms" % ((time.time() - t) * 1e3), det.phase_times())
    t = time.time(); det.run_phases(255); print("gpu pass %.1f ms" % ((time.time() - t) * 1e3), det.phase_times())
    for b in range(B):
        o = frcnn.forward(cfg, Wt, frames[b])
        with frcnn.exact():
            x = frcnn.forward(cfg, Wt, frames[b])
        # summary against the float64 evaluation (the centre both float32 realisations scatter around)
        def sd(a, bb):
            d = np.abs(a[:, None, :].astype(np.float64) - bb[None, :, :].astype(np.float64)).max(-1)
            return max(d.min(1).max(), d.min(0).max())
        for i in range(4):
            g = det.get_stage("c%d" % (i + 2))[b].transpose(2, 0, 1).astype(np.float64)
            r = x["c2345"][i].astype(np.float64)
            print("EXACT b%d c%d rel %.3e  magnitude bias (sum|gpu|-sum|ref|)/sum|ref| %+.3e   [oracle32: rel %.3e bias %+.3e]" % (
                b, i + 2, np.abs(g - r).max() / np.abs(r).max(), (np.abs(g).sum() - np.abs(r).sum()) / np.abs(r).sum(),
                np.abs(o["c2345"][i] - r).max() / np.abs(r).max(),
                (np.abs(o["c2345"][i].astype(np.float64)).sum() - np.abs(r).sum()) / np.abs(r).sum()))
        fcx = int(det.get_stage("final_count")[b].reshape(-1)[0])
        fbx = det.get_stage("final_boxes")[b].reshape(-1, 4)[:fcx]
        pcx = int(det.get_stage("proposal_count")[b].reshape(-1)[0])
        pbx = det.get_stage("proposal_boxes")[b].reshape(-1, 4)[:pcx]
        print("EXACT b%d proposals set-dist gpu %.3e oracle32 %.3e | final boxes set-dist gpu %.3e oracle32 %.3e (counts gpu %d exact %d)" % (
            b, sd(pbx, x["proposal_boxes"]), sd(o["proposal_boxes"], x["proposal_boxes"]),
            sd(fbx, x["final_boxes"]), sd(o["final_boxes"], x["final_boxes"]), fcx, len(x["final_boxes"])))
        for i in range(4):
            g = det.get_stage("c%d" % (i + 2))[b].transpose(2, 0, 1)
            e = (g.astype(np.float64) - o["c2345"][i]) / np.abs(o["c2345"][i]).max()
            print("b%d c%d rel %.3e  signed-mean %.2e rms %.2e  (|gpu|-|ref|)/|ref| mean %.2e" % (
                b, i + 2, rel(g, o["c2345"][i]), e.mean(), np.sqrt((e ** 2).mean()),
                ((np.abs(g) - np.abs(o["c2345"][i])).sum() / np.abs(o["c2345"][i]).sum())))
        for i in range(5):
            g = det.get_stage("p%d" % (i + 2))[b].transpose(2, 0, 1)
            ref = o["p23456"][i]
            g = g[:, :ref.shape[1], :ref.shape[2]]
            print("b%d p%d rel %.3e" % (b, i + 2, rel(g, ref)))
        for i in range(5):
            g = det.get_stage("rpn_l%d" % i)[b]
            cls, box = o["rpn"][i]
            print("b%d rpn%d cls rel %.3e box rel %.3e" % (b, i, rel(g[..., :3], cls), rel(g[..., 3:15].reshape(box.shape), box)))
        cnt = det.get_stage("lvl_count")[b].reshape(-1)
        lb = det.get_stage("lvl_boxes")[b].reshape(5, -1, 4); ls = det.get_stage("lvl_scores")[b].reshape(5, -1)
        for i in range(5):
            rb, rs = o["level_proposals"][i]
            n = min(int(cnt[i]), len(rs))
            print("b%d lvl%d count gpu %d ref %d  box maxabs %.3e score maxabs %.3e" % (
                b, i, cnt[i], len(rs), np.abs(lb[i, :n] - rb[:n]).max() if n else 0, np.abs(ls[i, :n] - rs[:n]).max() if n else 0))
        pc = int(det.get_stage("proposal_count")[b].reshape(-1)[0])
        pb = det.get_stage("proposal_boxes")[b].reshape(-1, 4)
        n = min(pc, len(o["proposal_boxes"]))
        print("b%d proposals gpu %d ref %d box maxabs %.3e" % (b, pc, len(o["proposal_boxes"]), np.abs(pb[:n] - o["proposal_boxes"][:n]).max()))
        K = cfg.rpn_test_post_nms_topk
        rf = det.get_stage("roi_feat").reshape(-1, 7, 7, 256)[b * K:b * K + n].transpose(0, 3, 1, 2)
        print("b%d roi_feat rel %.3e" % (b, rel(rf, o["roi_feat"][:n])))
        hl = det.get_stage("head_logits")[b, :n]
        nc = cfg.num_class
        print("b%d cls_logits rel %.3e box_logits rel %.3e" % (b, rel(hl[:, :nc, 0], o["cls_logits"][:n]),
              rel(hl[:, nc + 4:nc + 4 * nc, 0].reshape(n, nc - 1, 4), o["box_logits"][:n])))
        fc = int(det.get_stage("final_count")[b].reshape(-1)[0])
        fb = det.get_stage("final_boxes")[b].reshape(-1, 4); fp = det.get_stage("final_probs")[b].reshape(-1); fl = det.get_stage("final_labels")[b].reshape(-1)
        n = min(fc, len(o["final_probs"]))
        print("b%d final gpu %d ref %d labels_equal %s box maxabs %.3e prob maxabs %.3e" % (
            b, fc, len(o["final_probs"]), bool(np.array_equal(fl[:n], o["final_labels"][:n])),
            np.abs(fb[:n] - o["final_boxes"][:n]).max() if n else 0, np.abs(fp[:n] - o["final_probs"][:n]).max() if n else 0))
        bf = det.get_stage("fpn_box_feat")[b * cfg.result_per_im: b * cfg.result_per_im + n]
        print("b%d fpn_box_feat rel %.3e" % (b, rel(bf, o["fpn_box_feat"][:n])))


if __name__ == "__main__":
    main()
