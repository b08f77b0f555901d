"""Soak test of the tensor-core kernels' synchronisation (run under gpurun): many launches with fresh random draws / rays in every
mode -- bf16 and tc32 forward with perturb + noise, shipped-configuration forward (AWP), 120 full training steps (AWP on) -- and odd
batch sizes.  Any protocol bug shows as a trap (bounded waits) or a non-finite result.   python tools/stress.py [iters]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from evdeblurnerf_b200 import NeRFAll
dev = torch.device("cuda")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 40
P = bench.make_params(dev)
t0 = time.time()
for prec in ("bf16", "tc32"):
    nerf = NeRFAll(P, *bench.AABB, kernel_ptnum=bench.N_EXPOSURE, precision=prec).eval()
    for i in range(iters):
        n = (4096, 1000, 37, 2048, 4095)[i % 5]
        rays, idx = bench.make_rays(n, seed=9000 + i)
        rgb, rgb0 = nerf.render_blurred(bench.H, bench.W, bench.KMAT, rays.to(dev), idx.to(dev), N_samples=bench.NC, N_importance=bench.NI,
                                        perturb=float(i % 2), raw_noise_std=float(i % 2))
        assert bool(torch.isfinite(rgb).all()) and bool(torch.isfinite(rgb0).all()), (prec, i)
    torch.cuda.synchronize()
    print("ok forward", prec, iters, "launch sets,", round(time.time() - t0, 1), "s", flush=True)
    del nerf
P_all = dict(P); P_all.update(bench.awp_params(dev))
nerf = NeRFAll(P_all, *bench.AABB, kernel_ptnum=bench.N_EXPOSURE, precision="bf16", use_awp=True).train()
kw = dict(force_naive=False, retraw=True, N_samples=bench.NC, N_importance=bench.NI, perturb=1., raw_noise_std=1.)
for i in range(iters):
    rays, idx = bench.make_rays(bench.N_RAYS, seed=7000 + i)
    with torch.no_grad():
        rgb, rgb1, _, other = nerf(bench.H, bench.W, bench.KMAT, rays=rays.to(dev), rays_info={"images_idx": idx.to(dev)}, **kw)
    assert bool(torch.isfinite(rgb).all()) and bool(torch.isfinite(other["rgb_awp"]).all()), i
torch.cuda.synchronize()
print("ok shipped forward", iters, round(time.time() - t0, 1), "s", flush=True)
del nerf
flush = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
out = bench.train_leg(P_all, dev, "bf16", bench.N_RAYS, 1, 0, 3 * iters, 3, flush)
assert out["loss_finite"], out
print("ok train", 3 * iters, "steps", {k: out[k] for k in ("ms_per_step", "loss", "loss_finite")}, round(time.time() - t0, 1), "s", flush=True)
