#!/usr/bin/env python
"""Training-branch forward WITH the AWP branch (kernel_use_awp, every shipped config) on the headline batch:
NeRFAll.forward(force_naive=False) = RBK warp -> c2f render emitting depth_feature [R,128,128] -> AWP -> blends.
    python tools/bench_awp_forward.py [--steps K] [--precision bf16|fp32]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


awp_params = bench.awp_params      # moved to bench.py (the bench line reports the AWP-on figures too)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--precision", default="bf16")
    a = ap.parse_args()
    from evdeblurnerf_b200 import NeRFAll
    dev = torch.device("cuda")
    P = bench.make_params(dev)
    P.update(awp_params(dev, bench.N_EXPOSURE))
    rays, idx = bench.make_rays(bench.N_RAYS, 1)
    rays, idx = rays.to(dev), idx.to(dev)
    out = {}
    for use_awp in (False, True):
        nerf = NeRFAll(P, *bench.AABB, kernel_ptnum=bench.N_EXPOSURE, precision=a.precision, use_awp=use_awp).train()
        kw = dict(rays=rays, rays_info={"images_idx": idx}, force_naive=False, retraw=True, N_samples=bench.NC, N_importance=bench.NI,
                  perturb=1., raw_noise_std=1.)
        with torch.no_grad():
            for _ in range(3):
                nerf(bench.H, bench.W, bench.KMAT, **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(a.steps):
                nerf(bench.H, bench.W, bench.KMAT, **kw)
            e1.record()
            torch.cuda.synchronize()
        out["awp" if use_awp else "no_awp"] = e0.elapsed_time(e1) / a.steps
    print(json.dumps({"metric": "training-branch forward ms (4096 x 5 rays, 64+64)", "precision": a.precision, "ms": out}))


if __name__ == "__main__":
    main()
