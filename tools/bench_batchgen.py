#!/usr/bin/env python
"""Batch generation on the device vs the reference's host path, per training step of config 3:
4096 RGB rays (LLFFDataset.__getitem__) + 2048 event pairs (sample_events: successor walk, pose interpolation, rays).
    python tools/bench_batchgen.py"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    from scipy.spatial.transform import Rotation
    import evdeblur_oracle as oc
    from evdeblurnerf_b200.batchgen import EventBatchSampler, PoseInterpolator, RayBatchSampler
    rng = np.random.default_rng(0)
    g = torch.Generator().manual_seed(0)
    n_img, H, W, Kn = 30, 400, 400, 300
    K = [[400.0, 0, 200.0], [0, 400.0, 200.0], [0, 0, 1.0]]
    times = np.cumsum(rng.uniform(0.8e4, 1.2e4, Kn)) + 1.0e9
    rots = Rotation.from_rotvec(np.cumsum(rng.normal(0, 0.01, (Kn, 3)), 0)).as_matrix()
    trans = np.cumsum(rng.normal(0, 0.01, (Kn, 3)), 0)
    poses = torch.tensor(np.concatenate([rots[:n_img], trans[:n_img, :, None]], -1), dtype=torch.float32)
    images = torch.rand(n_img, H, W, 3, generator=g)
    n_ev = 2_000_000
    ts = np.sort(rng.uniform(times[0], times[-1], n_ev))
    succ = torch.arange(n_ev) + torch.randint(1, 2000, (n_ev,), generator=g)
    succ[succ >= n_ev] = 0
    pol = torch.randint(0, 2, (n_ev,), generator=g) * 2 - 1
    coord_id = torch.randint(0, H * W, (n_ev,), generator=g)
    events = torch.stack([coord_id.double(), torch.zeros(n_ev).double(), torch.as_tensor(ts), pol.double(), succ.double()], -1)
    id_to_coords = torch.stack([torch.arange(H * W) % W, torch.arange(H * W) // W], -1).float()
    rgb = RayBatchSampler(images, poses, K)
    interp = PoseInterpolator(times, rots, trans, bd_scale=0.5)
    evs = EventBatchSampler(events, id_to_coords, K, interp, num_successors=torch.full((n_ev,), 10), accum_steps=lambda s: (1, 4))
    ray_ids = torch.randint(0, len(rgb), (4096,), generator=g).cuda()
    ev_ids = torch.randint(0, n_ev, (2048,), generator=g).cuda()
    hops = torch.randint(0, 4, (2048,), generator=g).cuda()
    for _ in range(3):
        rgb[ray_ids]; evs.sample_events(ev_ids, 0, hops)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 50
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        rgb[ray_ids]; evs.sample_events(ev_ids, 0, hops)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / n
    dev_ms = e0.elapsed_time(e1) / n
    # the reference's host path for the same products (torch CPU indexing + scipy Slerp / interp1d every step)
    o_interp = oc.pose_interpolator(times, rots, trans)
    ids_c, hops_c, rid_c = ev_ids.cpu(), hops.cpu(), ray_ids.cpu()
    t0 = time.perf_counter()
    for _ in range(5):
        oc.make_rgb_batch(rid_c, images, poses, K)
        end, neg, pos = oc.gather_successor(ids_c, hops_c, succ, pol.int())
        for tq in (ts[ids_c.numpy()], ts[end.numpy()]):
            p = torch.as_tensor(oc.interpolate_event_poses(o_interp, tq, 0.5))
            oc.rays_from_pixels(id_to_coords[coord_id[ids_c]], K, p)
    cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
    print(json.dumps({"metric": "batch generation per training step (4096 rgb rays + 2048 event pairs)", "device_ms": dev_ms,
                      "host_wall_ms_including_launch_overhead": wall * 1e3, "cpu_oracle_port_ms": cpu_ms}))


if __name__ == "__main__":
    main()
