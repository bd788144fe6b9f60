"""BASELINE config 5 on N GPUs of one node (one camera stream per GPU): OSNet-x1.0 embedding of every track crop ->
device-resident gallery -> NCCL all-gather over NVLink -> every rank scores its share of the camera pairs
(multi_video_reid.py:448-476 pair loop, :308-324 feature distance, :512 lap.lapjv) -> track-id matches.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/c5_bench.py
  [--tracks 50 --crops 100 --batch 64 --precision split --p2p]

Rank 0 prints one JSON line: crops/s per GPU (host uint8 crops in, features left in HBM), all-gather ms / bytes against
the NVLink figure, pair-matching ms, and whether the matches are the planted identities.  Timing: CUDA events for the
embedding and the all-gather, host clock (max over ranks) for the pair phase, which contains host code (trajectory
distance, lapjv)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth_camera(cam, n_tracks, crops_per_track, seed=5):
    """Planted identities: identity k has a base crop shared by every camera (plus per-camera / per-crop noise) and a
    top-down trajectory shared up to jitter; identities come in spatial clusters of 5 so that the trajectory gate alone
    cannot decide.  Track ids are permuted per camera."""
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (n_tracks, 32, 16, 3)).astype(np.float32)                 # identity appearance, low-res
    centers = rng.uniform(0, 4000, (n_tracks // 5 + 1, 2))
    vel = rng.uniform(-3, 3, (n_tracks, 2))
    crng = np.random.default_rng(1000 + cam)
    perm = crng.permutation(n_tracks)
    variants = 8
    crops = np.empty((n_tracks, variants, 256, 128, 3), np.uint8)
    tracks_meta, ident_of = {}, {}
    for k in range(n_tracks):
        up = np.kron(base[k], np.ones((8, 8, 1), np.float32))
        for v in range(variants):
            crops[k, v] = np.clip(up + crng.standard_normal(up.shape) * 20, 0, 255).astype(np.uint8)
        frames = np.arange(0, crops_per_track, dtype=np.float64)
        pts = centers[k // 5] + np.array([(k % 5) * 8.0, 0.0]) + frames[:, None] * vel[k] + crng.standard_normal((len(frames), 2))
        tid = int(perm[k]) + 1
        tracks_meta[tid] = (np.concatenate([frames[:, None], pts], 1), crops_per_track)
        ident_of[tid] = k
    order = sorted(tracks_meta)                                                         # gallery rows in track-id order
    idx = np.array([ident_of[t] for t in order])
    return tracks_meta, ident_of, crops, idx, variants


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=50)
    ap.add_argument("--crops", type=int, default=100)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--precision", default="split")
    ap.add_argument("--p2p", action="store_true", help="also time the peer-memory form (reid.match_cameras_p2p)")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from object_detection_tracking_b200 import reid
    from object_detection_tracking_b200.synth import synth_osnet_state

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    gloo = dist.new_group(backend="gloo")                     # host-side objects (ids, trajectories) travel over gloo

    eng = reid.ReidEngine(args.batch, device=local, precision=args.precision)
    eng.load_state(synth_osnet_state(4321))
    meta, ident_of, crops, idx, variants = synth_camera(rank, args.tracks, args.crops)
    rows = args.tracks * args.crops
    D = eng.feat_dim
    gallery = torch.zeros((rows, D), dtype=torch.float32, device="cuda")

    # ---- embed: host uint8 crops -> features in HBM (row r = crop r % crops of track idx[r // crops]) ----
    def batch_crops(r0, r1):
        r = np.arange(r0, r1)
        return crops[idx[r // args.crops], (r % args.crops) % variants]

    pinned = [torch.from_numpy(batch_crops(r0, min(rows, r0 + args.batch))).pin_memory() for r0 in range(0, rows, args.batch)]
    eng.embed_dev(pinned[0].numpy(), gallery[:pinned[0].shape[0]])                       # warm-up + graph capture
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r0 = 0
    for pb in pinned:
        n = pb.shape[0]
        eng.embed_dev(pb.numpy(), gallery[r0:r0 + n])
        r0 += n
    e1.record()
    torch.cuda.synchronize()
    embed_ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    dist.all_reduce(embed_ms, op=dist.ReduceOp.MAX)

    # ---- exchange + pair matching ----
    timing = {}
    reid.match_cameras_allgather(meta, gallery, rows, device=local, object_group=gloo, tol=50, precision=args.precision)   # warm-up
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = reid.match_cameras_allgather(meta, gallery, rows, device=local, object_group=gloo, tol=50,
                                       precision=args.precision, timing=timing)
    torch.cuda.synchronize()
    pair_ms = torch.tensor([(time.perf_counter() - t0) * 1e3, timing["allgather_ms"]], device="cuda")
    dist.all_reduce(pair_ms, op=dist.ReduceOp.MAX)

    # ---- correctness: every match must pair the same planted identity; every identity must be found ----
    idents = [None] * world
    dist.all_gather_object(idents, ident_of, group=gloo)
    ok, total, found = 0, 0, 0
    for (i, j), matches in res.items():
        for a, b in matches:
            total += 1
            ok += int(idents[i][a] == idents[j][b])
        found += len(matches)
    stats = torch.tensor([ok, total, found, len(res) * args.tracks], device="cuda", dtype=torch.float64)
    dist.all_reduce(stats)

    p2p_ms = None
    if args.p2p:
        feats_host = gallery.cpu().numpy()
        order = sorted(meta)
        cam = {t: (meta[t][0], feats_host[k * args.crops:(k + 1) * args.crops]) for k, t in enumerate(order)}
        reid.match_cameras_p2p(cam, device=local, group=gloo, tol=50, precision=args.precision)      # warm-up
        dist.barrier()
        t0 = time.perf_counter()
        res2 = reid.match_cameras_p2p(cam, device=local, group=gloo, tol=50, precision=args.precision)
        t = torch.tensor([(time.perf_counter() - t0) * 1e3], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        p2p_ms = float(t[0])
        same = torch.tensor([int(res2 == res)], device="cuda")
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        p2p_same = bool(same[0])
    if rank == 0:
        n_pairs = world * (world - 1) // 2
        ag_ms = float(pair_ms[1])
        recv = timing["allgather_bytes_recv"]
        out = {"what": "config 5: multi-camera ReID, one camera per GPU", "n_gpus": world, "tracks_per_camera": args.tracks,
               "crops_per_track": args.crops, "feat_dim": D, "precision": args.precision,
               "embed_crops_per_s_per_gpu": rows / (float(embed_ms[0]) * 1e-3), "embed_ms": float(embed_ms[0]),
               "embed_batch": args.batch,
               "allgather_ms": ag_ms, "allgather_bytes_recv_per_gpu": recv,
               "allgather_gbs_per_gpu": (recv / (ag_ms * 1e-3) / 1e9) if ag_ms > 0 else None,
               "nvlink_peak_gbs_per_direction": 900.0,
               "camera_pairs": n_pairs, "exchange_plus_matching_ms": float(pair_ms[0]),
               "matching_ms_per_pair_per_rank": (float(pair_ms[0]) - ag_ms) / max(1, -(-n_pairs // world)),
               "matches_correct": int(stats[0]), "matches_total": int(stats[1]), "identities_expected": int(stats[3]),
               "timing": "CUDA events (embed, all-gather), host clock max over ranks (exchange + matching: contains host trajectory distance + lapjv)"}
        if p2p_ms is not None:
            out["p2p_exchange_plus_matching_ms"] = p2p_ms
            out["p2p_same_matches"] = p2p_same
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
