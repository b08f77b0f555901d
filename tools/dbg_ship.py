"""dev repro: shipped-config forward (AWP, perturb, noise) with another rank's rays on ONE GPU."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from evdeblurnerf_b200 import NeRFAll
dev = torch.device("cuda")
P = bench.make_params(dev)
P_all = dict(P); P_all.update(bench.awp_params(dev))
nerf = NeRFAll(P_all, *bench.AABB, kernel_ptnum=bench.N_EXPOSURE, precision="bf16", use_awp=True).train()
kw = dict(force_naive=False, retraw=True, N_samples=bench.NC, N_importance=bench.NI, perturb=1., raw_noise_std=1.)
for seed in (1000, 1001, 1002, 1003):
    rays, idx = bench.make_rays(bench.N_RAYS, seed=seed)
    rays, idx = rays.to(dev), idx.to(dev)
    for i in range(12):
        with torch.no_grad():
            nerf(bench.H, bench.W, bench.KMAT, rays=rays, rays_info={"images_idx": idx}, **kw)
        torch.cuda.synchronize()
    print("ok", seed, flush=True)
