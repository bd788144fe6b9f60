"""Multi-camera track-pair matching (BASELINE config 5, multi_video_reid.py:260-324, 486-534): the native trajectory
distance and the assignment step against fixtures produced by the REFERENCE's own compute_spatial_dist /
compute_feature_dist (tests/golden/make_golden_tmot.py: reid_pairs).  The feature distance itself is GPU work
(b2_track_pair_cost; tests/test_widen_gpu.py) -- here a float64 numpy checker stands in for it."""
import os

import numpy as np


def load_cameras(g):
    cams = []
    for name in ("c1", "c2"):
        cams.append({int(t): (g["%s_t%d_rows" % (name, t)], g["%s_t%d_feat" % (name, t)]) for t in g[name + "_ids"]})
    return cams


def feature_dist_checker(tracks1, tracks2, spatial):          # compute_feature_dist restated in float64 numpy
    out = np.full(spatial.shape, 999.0)
    for i, t1 in enumerate(sorted(tracks1)):
        for j, t2 in enumerate(sorted(tracks2)):
            if spatial[i, j] < 9999.:
                a, b = tracks1[t1][1].astype(np.float64), tracks2[t2][1].astype(np.float64)
                d = (a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2 * a @ b.T
                out[i, j] = np.maximum(d, 0).min()
    return out


def test_spatial_distance_matches_reference(golden_dir):
    from object_detection_tracking_b200 import reid
    g = np.load(os.path.join(golden_dir, "reid_pairs.npz"))
    c1, c2 = load_cameras(g)
    sp = reid.compute_spatial_dist(c1, c2, frame_offset=4, tol=50, ignore_pairs=[list(g["ignore0"]), list(g["ignore1"])])
    np.testing.assert_array_equal(sp < 9999, g["spatial"] < 9999)
    assert (g["spatial"] < 9999).sum() >= 10
    np.testing.assert_allclose(sp, g["spatial"], rtol=0, atol=1e-10)
    # without the frame offset the common frames change
    sp0 = reid.compute_spatial_dist(c1, c2, frame_offset=0, tol=50)
    assert not np.array_equal(sp0 < 9999, sp < 9999)


def test_checker_and_matching_match_reference(golden_dir):
    from object_detection_tracking_b200 import reid
    g = np.load(os.path.join(golden_dir, "reid_pairs.npz"))
    c1, c2 = load_cameras(g)
    fd = feature_dist_checker(c1, c2, g["spatial"])
    np.testing.assert_allclose(fd, g["feature"], rtol=1e-6, atol=5e-5)     # sklearn keeps float32 inputs in float32 (~1e-7 of the squared norms)
    got = reid.match_tracks(c1, c2, frame_offset=4, tol=50, ignore_pairs=[list(g["ignore0"]), list(g["ignore1"])],
                            feature_dist_fn=feature_dist_checker)
    ids1, ids2 = sorted(c1), sorted(c2)
    ref = [(ids1[i], ids2[int(j)]) for i, j in enumerate(g["x"]) if j >= 0]
    assert got == ref and len(ref) >= 5
    assert reid.match_tracks({}, c2) == []


def test_camera_pairs_cover_the_bubble_compare():
    from object_detection_tracking_b200.reid import camera_pairs
    allp = camera_pairs(8)
    assert len(allp) == 28 and allp[0] == (0, 1) and allp[-1] == (6, 7)
    dealt = [camera_pairs(8, r, 8) for r in range(8)]
    assert sorted(p for d in dealt for p in d) == sorted(allp)
    assert max(len(d) for d in dealt) - min(len(d) for d in dealt) <= 1


def _c5_worker(rank, world, port, q, golden_path):
    """One rank = one camera: contribute the camera's gallery, all-gather, then score this rank's share of the camera pairs."""
    import os as _os
    import torch
    import torch.distributed as dist
    from object_detection_tracking_b200 import reid
    _os.environ["MASTER_ADDR"] = "127.0.0.1"
    _os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(golden_path)
    cams = load_cameras(g)
    mine = cams[rank]
    ids = sorted(mine)
    local = torch.from_numpy(np.concatenate([mine[t][1] for t in ids], 0))
    allf, counts = reid.allgather_gallery(local)
    # the gallery of the OTHER camera as received over the wire replaces the local copy of it
    other = 1 - rank
    o_ids = sorted(cams[other])
    off = sum(counts[:other])
    recv = allf[off:off + counts[other]].numpy()
    pos = 0
    rebuilt = {}
    for t in o_ids:
        k = len(cams[other][t][1])
        rebuilt[t] = (cams[other][t][0], recv[pos:pos + k])
        pos += k
    pair_cams = [mine, rebuilt] if rank == 0 else [rebuilt, mine]
    res = []
    for (i, j) in reid.camera_pairs(world, rank, world):
        res.append(((i, j), reid.match_tracks(pair_cams[i], pair_cams[j], frame_offset=4, tol=50,
                                              ignore_pairs=[list(g["ignore0"]), list(g["ignore1"])],
                                              feature_dist_fn=feature_dist_checker)))
    q.put((rank, counts, res))
    dist.barrier()
    dist.destroy_process_group()


def test_c5_exchange_then_pair_matching_world2_gloo(golden_dir):
    """Config 5 on two ranks (gloo on the CPU; NCCL on the GPUs): all-gather of the per-camera galleries, the camera pairs
    dealt to the ranks, each rank's matches equal to the reference's assignment."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    path = os.path.join(golden_dir, "reid_pairs.npz")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_c5_worker, args=(r, 2, port, q, path)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=180) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.load(path)
    c1, c2 = load_cameras(g)
    ids1, ids2 = sorted(c1), sorted(c2)
    ref = [(ids1[i], ids2[int(j)]) for i, j in enumerate(g["x"]) if j >= 0]
    n1 = sum(len(c1[t][1]) for t in c1)
    n2 = sum(len(c2[t][1]) for t in c2)
    assert out[0][1] == out[1][1] == [n1, n2]
    pairs = [r for rank_out in out for r in rank_out[2]]
    assert [p[0] for p in pairs] == [(0, 1)]                 # one camera pair on two cameras, owned by rank 0
    assert pairs[0][1] == ref
