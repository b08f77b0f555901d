"""dev: host enqueue time vs device time of Trainer.step (is the training step launch bound?)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from evdeblurnerf_b200.trainer import Trainer
dev = torch.device("cuda")
P = bench.make_params(dev)
P_all = dict(P); P_all.update(bench.awp_params(dev))
for awp in (False, True):
    tr = Trainer(P_all if awp else P, None, *bench.AABB, kernel_ptnum=bench.N_EXPOSURE, precision="bf16", tv_loss_weight=1e-2, device=dev, use_awp=awp,
                 render_kwargs=dict(N_samples=bench.NC, N_importance=bench.NI, perturb=1., raw_noise_std=1.), check_numerics_every=0)
    tr.nerf.backward_chunk_rays = 20480
    rays, idx = bench.make_rays(bench.N_RAYS, seed=1)
    batch = {"rays": rays.to(dev), "images_idx": idx.to(dev), "rgbsf": torch.rand(bench.N_RAYS, 3).to(dev)}
    for _ in range(5):
        tr.step(batch, bench.H, bench.W, bench.KMAT)
    torch.cuda.synchronize()
    cpu, gpu = [], []
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter(); s.record()
        tr.step(batch, bench.H, bench.W, bench.KMAT)
        e.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        cpu.append((t1 - t0) * 1e3); gpu.append(s.elapsed_time(e))
    print("awp", awp, "host enqueue ms", round(sum(cpu) / 10, 2), "device span ms", round(sum(gpu) / 10, 2), flush=True)
    del tr
