"""Profiling driver: runs the headline step a few times (used under ncu; numbers printed here are never bench values)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from evdeblurnerf_b200 import NeRFAll

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
P = bench.make_params(dev)
nerf = NeRFAll(P, *bench.AABB, kernel_ptnum=bench.N_EXPOSURE, precision=prec).eval()
rays, idx = bench.make_rays(bench.N_RAYS, seed=1000)
rays, idx = rays.to(dev), idx.to(dev)
for _ in range(n):
    nerf.render_blurred(bench.H, bench.W, bench.KMAT, rays, idx, N_samples=bench.NC, N_importance=bench.NI, perturb=0., raw_noise_std=0.)
torch.cuda.synchronize()
print("done")
