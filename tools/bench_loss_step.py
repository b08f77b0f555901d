"""BASELINE config[2] (blurbatteries-style): forward + loss evaluation of one training iteration on one GPU, no backward.
   4096 frame rays x 5 exposures (RBK) + 2 x 2048 event rays (E = 1) -> CRF (gamma rgb / learned event) -> MSE (fine + coarse)
   + colour EGM loss (stage 0 + stage 1) + TV, everything through the public host mirror.  Prints one JSON line.
       python tools/bench_loss_step.py [fp32|bf16]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import evdeblurnerf_b200 as edn

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
dev = torch.device("cuda", 0)
g = torch.Generator().manual_seed(0)
P = bench.make_params(dev)
nerf = edn.NeRFAll(P, *bench.AABB, kernel_ptnum=bench.N_EXPOSURE, precision=prec).train()
Pc = {}
for i, (o, k) in zip((0, 2, 4, 6), ((16, 3), (16, 16), (16, 16), (1, 16))):
    Pc[f"tonemapping_event.linear.{i}.weight"] = ((torch.rand(o, k, generator=g) * 2 - 1) / k ** 0.5).to(dev)
    Pc[f"tonemapping_event.linear.{i}.bias"] = torch.zeros(o, device=dev)
crf = edn.TonemappingTransform(Pc, map_type_rgb="gamma", map_type_event="learn", extra_features_event=2)
N, M = bench.N_RAYS, 2048
rays, idx = bench.make_rays(N, seed=1)
ev0, _ = bench.make_rays(M, seed=2)
ev1 = ev0 + 0.002 * torch.randn(ev0.shape, generator=g)
target = torch.rand(N, 3, generator=g)
pol = torch.stack([-(torch.rand(M, generator=g) < 0.5).float(), (torch.rand(M, generator=g) < 0.5).float()], -1)
cmask = torch.nn.functional.one_hot(torch.randint(0, 3, (M,), generator=g), 3).bool()
cpol = torch.zeros(M, 3, 2)
cpol[cmask] = pol
bii = (torch.tensor([0.25, 0.25]) * pol).sum(-1)
rays, idx, ev0, ev1, target, cpol, cmask, bii = [t.to(dev) for t in (rays, idx, ev0, ev1, target, cpol, cmask, bii)]
kw = dict(N_samples=bench.NC, N_importance=bench.NI, perturb=0., raw_noise_std=0.)


def step():
    rgb, rgb0, other_loss, _ = nerf(bench.H, bench.W, bench.KMAT, rays=rays, rays_info={"images_idx": idx}, force_naive=False, **kw)
    loss = edn.img2mse(crf(rgb, mode="encode_rgb"), target) + edn.img2mse(crf(rgb0, mode="encode_rgb"), target)
    s1, s0, _, _ = nerf(bench.H, bench.W, bench.KMAT, rays=ev0, rays_info=None, force_naive=True, **kw)
    e1, e0, _, _ = nerf(bench.H, bench.W, bench.KMAT, rays=ev1, rays_info=None, force_naive=True, **kw)
    egm = 0
    for a_, b_ in ((s0, e0), (s1, e1)):
        la = crf(a_, mode="encode_luma", ev_extra_feat=cpol, tonemap_only=True)
        lb = crf(b_, mode="encode_luma", ev_extra_feat=cpol, tonemap_only=True)
        egm = egm + edn.egm_loss(la, lb, bii, color_mask=cmask, color_weight=[0.4, 0.2, 0.4])
    return loss + other_loss["TV"] * 1e-1 + egm * 1e-1


for _ in range(3):
    step()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(10):
    flush.fill_(1)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); val = step(); e.record()
    torch.cuda.synchronize()
    ts.append(s.elapsed_time(e))
ms = sum(ts) / len(ts)
print(json.dumps({"workload": "config[2]: forward + loss (4096x5 frame rays, 2x2048 event rays, CRF, MSE, colour EGM, TV x3), no backward",
                  "precision": prec, "ms_per_iteration": ms, "frame_rays_per_s": N / (ms / 1e3), "loss": float(val)}))
