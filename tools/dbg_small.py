"""dev: the bench legs at a reduced ray count on ONE GPU (repro helper for the strong-scaling leg)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from evdeblurnerf_b200 import NeRFAll
dev = torch.device("cuda")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
P = bench.make_params(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for prec in ("bf16", "tc32"):
    nerf = NeRFAll(P, *bench.AABB, kernel_ptnum=bench.N_EXPOSURE, precision=prec).eval()
    for seed in (4242, 1):
        rays, idx = bench.make_rays(bench.N_RAYS, seed=seed)
        for lo in (0, n):
            r, i = rays[lo:lo + n].to(dev), idx[lo:lo + n].to(dev)
            for _ in range(3):
                nerf.render_blurred(bench.H, bench.W, bench.KMAT, r, i, N_samples=bench.NC, N_importance=bench.NI, perturb=0., raw_noise_std=0.)
            torch.cuda.synchronize()
            print("ok forward", prec, seed, lo, flush=True)
P_all = dict(P); P_all.update(bench.awp_params(dev))
out = bench.train_leg(P_all, dev, "bf16", n, 1, 0, 4, 3, flush)
torch.cuda.synchronize()
print("ok train", {k: out[k] for k in ("ms_per_step", "loss_finite")}, flush=True)
