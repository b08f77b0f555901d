#!/usr/bin/env python
"""Config 4 (BASELINE.json): full training step -- forward, losses, backward, gradient all-reduce, fused Adam -- on the
headline batch (4096 primary rays x 5 exposures x 64+64 samples, full-size grids).  Prints one JSON line per precision.

    python tools/bench_train_step.py [--steps K] [--warmup W] [--precision bf16 fp32]
    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train_step.py   (weak scaling, NCCL all-reduce)
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", nargs="+", default=["bf16", "fp32"])
    ap.add_argument("--rays", type=int, default=bench.N_RAYS)
    ap.add_argument("--chunk-rays", type=int, default=20480)
    ap.add_argument("--events", type=int, default=0, help="config 3: add M start/end event-ray pairs + EGM loss + learnable CRF")
    ap.add_argument("--awp", action="store_true", help="train with the AWP branch (kernel_use_awp, as in the shipped configs)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        torch.distributed.init_process_group("nccl")
    from evdeblurnerf_b200.parallel import max_over_ranks
    from evdeblurnerf_b200.trainer import Trainer
    dev = torch.device("cuda", local)
    for precision in args.precision:
        P = bench.make_params(dev, seed=0)
        if args.awp:
            from bench_awp_forward import awp_params
            P.update(awp_params(dev, bench.N_EXPOSURE))
        crf_state = None
        if args.events:
            g = torch.Generator().manual_seed(5)
            crf_state = {}
            for idx, (o, i) in ((0, (16, 3)), (2, (16, 16)), (4, (16, 16)), (6, (1, 16))):
                crf_state[f"tonemapping_event.linear.{idx}.weight"] = (torch.rand(o, i, generator=g) * 2 - 1) / i ** 0.5
                crf_state[f"tonemapping_event.linear.{idx}.bias"] = (torch.rand(o, generator=g) * 2 - 1) / i ** 0.5
        tr = Trainer(P, crf_state, *bench.AABB, event_loss_weight=1.0 if args.events else 0.0, kernel_ptnum=bench.N_EXPOSURE, precision=precision, tv_loss_weight=1e-2, device=dev,
                     use_awp=args.awp,
                     render_kwargs=dict(N_samples=bench.NC, N_importance=bench.NI, perturb=1., raw_noise_std=1.))
        del P
        tr.nerf.backward_chunk_rays = args.chunk_rays
        batches = []
        for s in range(args.steps + args.warmup):
            rays, idx = bench.make_rays(args.rays, seed=1000 * rank + s)
            g = torch.Generator().manual_seed(s)
            b = {"rays": rays.to(dev), "images_idx": idx.to(dev), "rgbsf": torch.rand(args.rays, 3, generator=g).to(dev)}
            if args.events:
                ev0, _ = bench.make_rays(args.events, seed=77000 + s)
                ev1 = ev0.clone()
                ev1[:, :, 0] += 0.005 * torch.randn(args.events, 3, generator=g)
                b.update(ev_rays_start=ev0.to(dev), ev_rays_end=ev1.to(dev), ev_extra_feat=torch.rand(args.events, 2, generator=g).to(dev),
                         bii=(0.25 * torch.randint(-2, 3, (args.events,), generator=g).float()).to(dev))
            batches.append(b)
        for s in range(args.warmup):
            tr.step(batches[s], bench.H, bench.W, bench.KMAT)
        tr.nerf.engine.profile = {}
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(args.steps):
            out = tr.step(batches[args.warmup + s], bench.H, bench.W, bench.KMAT)
        e1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(e0.elapsed_time(e1) / args.steps, dev)
        kern = {k: sum(a.elapsed_time(b) for a, b in v) / args.steps for k, v in tr.nerf.engine.profile.items()}
        if rank == 0:
            print(json.dumps({"metric": "training rays/s (fwd + loss + bwd + all-reduce + Adam)", "value": world * args.rays / (ms * 1e-3),
                              "unit": "rays/s", "n_gpus": world, "ms_per_step": ms, "precision": precision, "awp": bool(args.awp), "event_rays": args.events,
                              "kernels_ms": kern, "loss": float(out["loss"]),
                              "config": {"workload": f"{args.rays} rays x {bench.N_EXPOSURE} exposures x {bench.NC}+{bench.NI} samples, "
                                                     "full-size VM grids, TV + MSE losses, Adam over 36.8M parameters",
                                         "backward_chunk_rays": args.chunk_rays}}), flush=True)
        del tr
        torch.cuda.empty_cache()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
