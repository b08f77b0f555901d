#!/usr/bin/env python
"""Headline benchmark of the fused render path (BASELINE.json: primary rays/s, 4096-ray / 64+64-sample / 5-exposure).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

One "step" = one forward pass of the hot path over one synthetic batch: 4096 primary rays -> 5 exposures (20480
sub-rays) -> coarse pass (64 samples) -> hierarchical sampling -> fine pass (128 samples) -> composited colours.
N > 1 (torchrun): every rank renders its own 4096-ray batch (weak scaling, no data-path collective); value = all
rays / max-over-ranks time.  Prints ONE JSON line on rank 0.

Beside the headline (kept as in round 1 for continuity) the line carries, at EVERY N:
  "shipped_forward": the training-branch forward of the shipped configuration (AWP on, perturb = 1, raw_noise_std = 1);
  "train_step":      BASELINE config 4 -- forward + losses + backward + the NCCL gradient all-reduce (+ the synchronised-BatchNorm
                     exchange of the AWP branch) + Adam, weak scaling, with `loss_finite` and the all-reduce's share;
  "strong":          SURVEY 8(e)'s own partition -- the SAME 4096-ray batch split across the N ranks -- forward and training.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS, N_EXPOSURE, NC, NI = 4096, 5, 64, 64
AABB = ((-1.5, -1.5, -1.0), (1.5, 1.5, 1.0))
COARSE_VOX, FINE_VOX = 16777248, 134217984          # configs/.../tx_blurfactory_*.txt:66,74
H = W = 400
FOCAL = 400.0
N_IMGS = 30
KMAT = [[FOCAL, 0.0, 200.0], [0.0, FOCAL, 200.0], [0.0, 0.0, 1.0]]

# algorithmic work, MACs
MAC_COARSE_SAMPLE = 17152                            # basis 3072 + sigma_net 7104 + color_net 6976
MAC_FINE_SAMPLE = 171520                             # sigma_net 65536 + color_net 105984
MAC_BASIS = 3072


def flops_per_subray(nc=NC, ni=NI):
    return 2 * (nc * MAC_COARSE_SAMPLE + 3 * MAC_BASIS * ni + (nc + ni) * MAC_FINE_SAMPLE)


def flops_fine_kernel_per_subray(nc=NC, ni=NI):
    return 2 * (3 * MAC_BASIS * ni + (nc + ni) * MAC_FINE_SAMPLE)


def grid_size(n_voxels):
    """voxnerf.py:87-92."""
    import torch
    amin, amax = torch.tensor(AABB[0]), torch.tensor(AABB[1])
    voxel = ((amax - amin).prod() / n_voxels).pow(1 / 3)
    return ((amax - amin) / voxel).long().tolist()


def make_params(device, seed=0):
    """Random-init parameters with the reference's names, shapes and init scales (voxnerf.py:104-118, nn.Linear)."""
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    P = {}

    def lin(name, out_c, in_c):
        bound = 1.0 / (in_c ** 0.5)
        P[name] = ((torch.rand(out_c, in_c, generator=g) * 2 - 1) * bound).to(device)

    for pre, nvox, hid, geo in (("mlp_coarse.", COARSE_VOX, 64, 15), ("mlp_fine.", FINE_VOX, 256, 128)):
        gs = grid_size(nvox)
        gd = torch.Generator(device=device).manual_seed(seed + (1 if hid == 64 else 2))
        for i, (m, v) in enumerate((((0, 1), 2), ((0, 2), 1), ((1, 2), 0))):
            c = (64, 16, 16)[i]
            P[pre + f"app_plane.{i}"] = 0.1 * torch.randn((1, c, gs[m[1]], gs[m[0]]), generator=gd, device=device)
            P[pre + f"app_line.{i}"] = 0.1 * torch.randn((1, c, gs[v], 1), generator=gd, device=device)
        lin(pre + "basis_mat.weight", 32, 96)
        lin(pre + "sigma_net.0.weight", hid, (32 if hid == 64 else 64) + 63)
        lin(pre + "sigma_net.1.weight", 1 + geo, hid)
        lin(pre + "color_net.0.weight", hid, geo + 27)
        lin(pre + "color_net.1.weight", hid, hid)
        lin(pre + "color_net.2.weight", 3, hid)
    # DP-NeRF rigid blur kernel (dpnerf/blurmodel.py:38-45); latents / r,v heads randomised so that the warp is non-trivial
    pre = "kernelsnet."
    P[pre + "view_embed_module.img_embed"] = (0.5 * torch.randn(N_IMGS, 32, generator=g)).to(device)
    for h in ("r", "v", "w"):
        lin(pre + f"{h}_branch.0.weight", 32, 32)
        P[pre + f"{h}_branch.0.bias"] = torch.zeros(32, device=device)
    for h, n_out in (("r", 3 * (N_EXPOSURE - 1)), ("v", 3 * (N_EXPOSURE - 1)), ("w", N_EXPOSURE)):
        P[pre + f"{h}_linear.weight"] = (0.05 * torch.randn(n_out, 32, generator=g)).to(device)
        P[pre + f"{h}_linear.bias"] = torch.zeros(n_out, device=device)
    return P


def awp_params(device, E=N_EXPOSURE, seed=7):
    """AdaptiveWeightProposal parameters (dpnerf/awp.py:9-47, mam.py) with nn.Linear / Conv1d default-init scales."""
    import torch
    g = torch.Generator().manual_seed(seed)
    P = {}

    def lin(name, o, i, bias=True, extra=()):
        b = 1.0 / (i ** 0.5)
        P[name + ".weight"] = ((torch.rand(o, i, *extra, generator=g) * 2 - 1) * b).to(device)
        if bias:
            P[name + ".bias"] = ((torch.rand(o, generator=g) * 2 - 1) * b).to(device)

    pre = "awpnet."
    lin(pre + "sample_feature_embed_layer.0", 64, 128)
    for l in (1, 2, 3):
        lin(pre + f"sample_feature_embed_layer.{l}", 64, 64)
    lin(pre + "motion_feature_embed_layer.0", 32, 111)
    lin(pre + "motion_feature_embed_layer.1", 32, 32)
    lin(pre + "MAM.linear", 32, 64)
    P[pre + "MAM.Corr.line_conv_att.weight"] = (torch.randn(1, 32, 1, 1, generator=g) * 0.2).to(device)
    for n, (o, i) in (("conva", (16, 32)), ("convb", (16, 32)), ("convc", (16, 32)), ("convn", (16, 16)), ("convl", (16, 16))):
        lin(pre + "MAM.Corr." + n, o, i, bias=False, extra=(1,))
    lin(pre + "MAM.Corr.convd.0", 32, 32, bias=False, extra=(1,))
    P[pre + "MAM.Corr.convd.1.weight"] = torch.ones(32, device=device)
    P[pre + "MAM.Corr.convd.1.bias"] = torch.zeros(32, device=device)
    lin(pre + "w_linear", E, 32)
    return P


def make_rays(n, seed):
    """SURVEY 8(d) synthetic primary rays [n,3,2] (camera looking down -z so that NDC is well posed) + image ids [n,1]."""
    import torch
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(n, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0])
    d = torch.cat([torch.randn(n, 2, generator=g) * 0.3, -torch.ones(n, 1)], -1)
    idx = torch.randint(0, N_IMGS, (n, 1), generator=g)
    return torch.stack([o, d], -1).contiguous(), idx


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML in a thread every ~2 ms (the region is tens of
    milliseconds; `nvidia-smi -lms` starts too slowly for that), `nvidia-smi` polling as the fallback."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
    REASONS = (("hw_slowdown", 0x8), ("sw_thermal_slowdown", 0x20), ("hw_thermal_slowdown", 0x40), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.idx, self.sm, self.mx, self.reasons = gpu_index, [], None, set()
        self.stop_flag, self.thread, self.proc, self.rows = threading.Event(), None, None, []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
        try:
            mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
        except Exception:
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
        for name, bit in self.REASONS:
            if mask & bit:
                self.reasons.add(name)

    def _loop(self):
        while not self.stop_flag.is_set():
            try:
                self._sample_nvml()
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            try:
                self._sample_nvml()          # at least one sample taken while the last timed kernels are still in flight
            except Exception:
                pass
            self.stop_flag.set()
            if self.thread is not None:
                self.thread.join(timeout=1.0)
            sm = sorted(self.sm)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml"}
        if self.proc is not None:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def _reference_kwargs(perturb=0., raw_noise_std=0.):
    """render_kwargs_train of run_nerf.py:306-330 for the blurfactory configs (ndc, near 0, far 1, use_viewdirs)."""
    return dict(retraw=True, force_naive=False, perturb=perturb, N_importance=NI, N_samples=NC, use_viewdirs=True,
                white_bkgd=False, raw_noise_std=raw_noise_std, inference=False, near=0., far=1.)


def _harness():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_harness as rh
    return rh


def cpu_reference_arm(sample_rays, steps, warmup, threads):
    """The reference's own CPU path on the host cores: the UNMODIFIED reference (oracle/_ref staged copy, or /root/reference in the
    build container; kind "reference") through its public NeRFAll.forward in training mode -- its chunk loop, NaN checks and
    per-exposure Python included -- else the oracle port (kind "port").  Returns (kind, ms per step)."""
    import torch
    torch.set_num_threads(threads)
    rh = _harness()
    P = make_params("cpu")
    rays, idx = make_rays(sample_rays, seed=100)
    if rh.reference_available():
        nerf = rh.build_bench_reference(P, N_EXPOSURE, False, "cpu")
        kw = _reference_kwargs()

        def one():
            with torch.no_grad():
                nerf(H, W, KMAT, chunk=1024 * 32, rays=rays, rays_info={"images_idx": idx}, **kw)
        kind = "reference"
    else:
        import evdeblur_oracle as oc
        cfg = {"aabb_min": AABB[0], "aabb_max": AABB[1], "rmnearplane": 0}

        def one():
            with torch.no_grad():
                oc.forward_train(P, cfg, H, W, FOCAL, rays, idx, N_EXPOSURE, NC, NI, use_awp=False)
        kind = "port"
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        one()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return kind, 1e3 * sum(times) / max(len(times), 1)


def gpu_eager_reference(P_dev, dev, n_rays, steps=3):
    """Baseline leg, second figure: the reference's own single-GPU PyTorch path ON the B200 (north star: >= 10x this) -- the
    UNMODIFIED reference from oracle/_ref with CUDA as the default device (run_nerf.py:779), eager fp32, its own chunk loop
    (chunk = 32768) and host-synchronising NaN checks; the oracle port as eager PyTorch when nothing is staged.  Forward
    (no_grad) of the bench workload, then forward + autograd backward of the image loss; CUDA-event timed, best of `steps`."""
    import torch
    rh = _harness()
    rays, idx = make_rays(n_rays, seed=100)
    rays, idx = rays.to(dev), idx.to(dev)
    real = rh.reference_available()
    out = {"kind": "reference" if real else "port", "rays": n_rays, "unit": "rays/s",
           "what": ("UNMODIFIED reference (oracle/_ref) NeRFAll.forward, training branch, eager fp32, default device = this GPU, chunk 32768"
                    if real else "oracle restatement of the reference path as eager PyTorch fp32 on this GPU, whole batch in one chunk")}

    def timed(fn):
        best = None
        for i in range(steps + 1):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); fn(); e.record()
            torch.cuda.synchronize()
            if i > 0:
                best = s.elapsed_time(e) if best is None else min(best, s.elapsed_time(e))
        return best

    with torch.device(dev):
        target = torch.rand(n_rays, 3, device=dev)
        if real:
            nerf = rh.build_bench_reference(P_dev, N_EXPOSURE, False, dev)
            kw = _reference_kwargs()
            call = lambda: nerf(H, W, KMAT, chunk=1024 * 32, rays=rays, rays_info={"images_idx": idx}, **kw)

            def fwd():
                with torch.no_grad():
                    call()

            def fwd_bwd():
                rgb, rgb0, _, _ = call()
                loss = torch.mean((rgb - target) ** 2) + torch.mean((rgb0 - target) ** 2)
                loss.backward()
                nerf.zero_grad(set_to_none=True)
        else:
            import evdeblur_oracle as oc
            cfg = {"aabb_min": AABB[0], "aabb_max": AABB[1], "rmnearplane": 0}
            leaves = {k: v.detach().clone().requires_grad_(True) for k, v in P_dev.items()}

            def fwd():
                with torch.no_grad():
                    oc.forward_train(P_dev, cfg, H, W, FOCAL, rays, idx, N_EXPOSURE, NC, NI, use_awp=False)

            def fwd_bwd():
                o = oc.forward_train(leaves, cfg, H, W, FOCAL, rays, idx, N_EXPOSURE, NC, NI, use_awp=False)
                loss = oc.img2mse(o["rgb"], target) + oc.img2mse(o["rgb1"], target)
                loss.backward()
                for v in leaves.values():
                    v.grad = None
        ms = timed(fwd)
        out.update(fwd_ms=ms, value=n_rays / (ms / 1e3))
        ms = timed(fwd_bwd)
        out.update(fwd_bwd_ms=ms, fwd_bwd_value=n_rays / (ms / 1e3))
        if real:
            # the configuration every shipped config runs (kernel_use_awp, perturb = 1, raw_noise_std = 1): what `shipped_forward`
            # and `train_step` of this line are to be compared with
            try:
                del nerf
                torch.cuda.empty_cache()
                nerf = rh.build_bench_reference(P_dev, N_EXPOSURE, True, dev)
                kw = _reference_kwargs(perturb=1., raw_noise_std=1.)
                call = lambda: nerf(H, W, KMAT, chunk=1024 * 32, rays=rays, rays_info={"images_idx": idx}, **kw)

                def fwd_awp():
                    with torch.no_grad():
                        call()

                def fwd_bwd_awp():
                    rgb, rgb0, _, other = call()
                    loss = (torch.mean((rgb - target) ** 2) + torch.mean((rgb0 - target) ** 2) + torch.mean((other["rgb_awp"] - target) ** 2))
                    loss.backward()
                    nerf.zero_grad(set_to_none=True)
                ms_f, ms_fb = timed(fwd_awp), timed(fwd_bwd_awp)
                out["shipped"] = {"what": "same module with kernel_use_awp, perturb = 1, raw_noise_std = 1; loss = MSE x 3", "fwd_ms": ms_f,
                                  "fwd_value": n_rays / (ms_f / 1e3), "fwd_bwd_ms": ms_fb, "fwd_bwd_value": n_rays / (ms_fb / 1e3)}
            except Exception as e:      # a reported comparison, never a reason to lose the bench line
                out["shipped"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_rays = args.sample_rays
    kind, ms = cpu_reference_arm(sample_rays, args.steps, args.warmup, threads)
    val = sample_rays / (ms / 1e3)
    line = {"impl": "reference", "metric": "primary_rays_per_sec_fwd", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config("fp32", sample=f"{sample_rays} primary rays x {N_EXPOSURE} exposures per step"),
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": threads, "kind": kind,
                             "sample": f"{sample_rays} primary rays x {N_EXPOSURE} exposures, {NC}+{NI} samples, full-size VM grids"
                                       + ("; UNMODIFIED reference NeRFAll.forward (training branch, AWP off) from oracle/_ref" if kind == "reference" else "")},
            "e2e": {"value": val, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def workload_config(precision, **extra):
    cfg = {"workload": f"blurfactory c2f render fwd (RBK warp + NDC -> coarse -> sample_pdf -> fine -> exposure blend): "
                       f"{N_RAYS} primary rays x {N_EXPOSURE} exposures, {NC}+{NI} samples, "
                       f"PDRF coarse {grid_size(COARSE_VOX)} / fine {grid_size(FINE_VOX)} VM grids",
           "rays": N_RAYS, "exposures": N_EXPOSURE, "samples": [NC, NI], "precision": precision,
           "perturb": 0, "l2": "flushed between timed steps (256 MiB write)"}
    cfg.update(extra)
    return cfg


def timed_steps(fn, steps, warmup, world, dev, before_each=None):
    """W untimed + K timed calls of fn(i); barrier + synchronize on both sides; CUDA events; MAX over ranks -> ms per step."""
    import torch
    import torch.distributed as dist
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = []
    for i in range(steps):
        if before_each is not None:
            before_each()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn(warmup + i)
        e.record()
        ev.append((s, e))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([sum(s.elapsed_time(e) for s, e in ev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


def train_leg(P_all, dev, precision, n_rays_rank, world, rank, steps, warmup, flush):
    """BASELINE config 4: Trainer.step on this rank's rays -- forward (AWP on, perturb = 1, raw_noise_std = 1) + MSE (fine, coarse,
    AWP) + TV losses + backward + ONE flat gradient all-reduce over NCCL (+ the synchronised-BatchNorm exchange) + fused Adam."""
    import torch
    from evdeblurnerf_b200.trainer import Trainer
    tr = Trainer(P_all, None, *AABB, kernel_ptnum=N_EXPOSURE, precision=precision, tv_loss_weight=1e-2, device=dev, use_awp=True,
                 render_kwargs=dict(N_samples=NC, N_importance=NI, perturb=1., raw_noise_std=1.), check_numerics_every=0)
    tr.nerf.backward_chunk_rays = 20480
    batches = []
    for s_ in range(steps + warmup):
        rays, idx = make_rays(n_rays_rank, seed=5000 + 1000 * rank + s_)
        g = torch.Generator().manual_seed(s_)
        batches.append({"rays": rays.to(dev), "images_idx": idx.to(dev), "rgbsf": torch.rand(n_rays_rank, 3, generator=g).to(dev)})
    last = {}
    tr.flat.profile = []

    def one(i):
        last["out"] = tr.step(batches[i], H, W, KMAT)
    ms = timed_steps(one, steps, warmup, world, dev, before_each=lambda: flush.fill_(1))
    ar = tr.flat.profile[-steps:] if world > 1 else []
    ar_ms = sum(a.elapsed_time(b) for a, b in ar) / max(len(ar), 1) if ar else 0.0
    loss = float(last["out"]["loss"])
    errs = tr.nerf.engine.numerical_errors()
    finite = torch.tensor([1.0 if (loss == loss and abs(loss) != float("inf") and not errs and bool(torch.isfinite(tr.flat.param).all())) else 0.0],
                          device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(finite, op=dist.ReduceOp.MIN)
    res = {"ms_per_step": ms, "value": world * n_rays_rank / (ms / 1e3), "unit": "rays/s", "rays_per_rank": n_rays_rank, "loss": loss,
           "loss_finite": bool(finite.item() > 0.5), "numerical_errors": errs, "all_reduce_ms": ar_ms,
           "all_reduce_share": ar_ms / ms if ms else None, "all_reduce_bytes": tr.flat.numel * 4 if world > 1 else 0}
    del tr, batches
    torch.cuda.empty_cache()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("EDN_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16", "tc32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the shipped-config / training / strong-scaling legs")
    ap.add_argument("--sample-rays", type=int, default=512, help="--impl reference: primary rays per CPU step (bounded sample)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from evdeblurnerf_b200 import NeRFAll

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    P = make_params(dev)
    nerf = NeRFAll(P, *AABB, kernel_ptnum=N_EXPOSURE, precision=args.precision).eval()
    eng = nerf.engine
    rays_host, idx_host = make_rays(N_RAYS, seed=1000 + rank)
    rays_host, idx_host = rays_host.pin_memory(), idx_host.pin_memory()
    rays_dev, idx_dev = rays_host.to(dev), idx_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out_host = torch.empty((N_RAYS, 3), dtype=torch.float32).pin_memory()

    def step(rays, idx):
        # public API call of one render: NeRFAll.render_blurred == the render part of NeRFAll.forward (training branch)
        return nerf.render_blurred(H, W, KMAT, rays, idx, N_samples=NC, N_importance=NI, perturb=0., raw_noise_std=0.)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(rays_dev, idx_dev)
    torch.cuda.synchronize()

    # ---- device-resident timing: K steps, L2 flushed before each, CUDA events around each step -----------------------
    sampler = ClockSampler(local)
    eng.profile = {}
    ev = []
    barrier()
    sampler.start()
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step(rays_dev, idx_dev)
        e.record()
        ev.append((s, e))
    barrier()
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    total_ms = sum(s.elapsed_time(e) for s, e in ev)
    kern_ms = {k: sum(s.elapsed_time(e) for s, e in v) / len(v) for k, v in eng.profile.items()}
    eng.profile = None

    # ---- end to end through the public API: pinned host rays -> H2D -> render -> D2H colours, every step ------------
    ev2 = []
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r_d = rays_host.to(dev, non_blocking=True)
        i_d = idx_host.to(dev, non_blocking=True)
        rgb, _ = step(r_d, i_d)
        out_host.copy_(rgb, non_blocking=True)
        e.record()
        ev2.append((s, e))
    barrier()
    e2e_ms = sum(s.elapsed_time(e) for s, e in ev2)
    headline_errs = eng.numerical_errors()

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = t.tolist()

    # ---- the legs beside the headline, at every N (all ranks take part) ------------------------------------------------------
    extra = {}
    if not args.no_train:
        k2, w2 = max(4, min(args.steps, 8)), 3
        flush_fn = lambda: flush.fill_(1)
        # (0) the same headline render in the tensor-core PARITY mode (bf16 x 3 split operands: fp32-grade, tests/test_render_gpu.py)
        if args.precision != "tc32":
            nerf_p = NeRFAll(P, *AABB, kernel_ptnum=N_EXPOSURE, precision="tc32").eval()
            ms = timed_steps(lambda i: nerf_p.render_blurred(H, W, KMAT, rays_dev, idx_dev, N_samples=NC, N_importance=NI, perturb=0., raw_noise_std=0.),
                             k2, w2, world, dev, before_each=flush_fn)
            extra["parity_mode"] = {"precision": "tc32", "ms_per_step": ms, "value": world * N_RAYS / (ms / 1e3), "unit": "rays/s",
                                    "what": "same workload, coarse and fine passes on tcgen05 with bf16 x 3 split operands + fp32 TMEM accumulation "
                                            "(meets the 1e-4 parity bar: tests/test_render_gpu.py runs in this precision)",
                                    "numerical_errors": nerf_p.engine.numerical_errors()}
            del nerf_p
            torch.cuda.empty_cache()
        # (a) shipped configuration, forward: training branch with the AWP branch on, perturb = 1, raw_noise_std = 1
        P_all = dict(P)
        P_all.update(awp_params(dev))
        nerf_awp = NeRFAll(P_all, *AABB, kernel_ptnum=N_EXPOSURE, precision=args.precision, use_awp=True).train()
        kw_ship = dict(force_naive=False, retraw=True, N_samples=NC, N_importance=NI, perturb=1., raw_noise_std=1.)

        def ship(i):
            with torch.no_grad():
                nerf_awp(H, W, KMAT, rays=rays_dev, rays_info={"images_idx": idx_dev}, **kw_ship)
        ms = timed_steps(ship, k2, w2, world, dev, before_each=flush_fn)
        extra["shipped_forward"] = {"ms_per_step": ms, "value": world * N_RAYS / (ms / 1e3), "unit": "rays/s",
                                    "what": "NeRFAll.forward training branch: RBK warp -> c2f render emitting depth_feature -> AWP -> "
                                            "exposure blends + TV; kernel_use_awp, perturb = 1, raw_noise_std = 1 (every shipped config)",
                                    "numerical_errors": nerf_awp.engine.numerical_errors()}
        del nerf_awp
        torch.cuda.empty_cache()
        # (b) config 4: full training step, weak scaling (4096 rays per rank), the gradient all-reduce inside
        extra["train_step"] = dict(train_leg(P_all, dev, args.precision, N_RAYS, world, rank, k2, w2, flush), scaling="weak", n_gpus=world,
                                   what="Trainer.step: fwd (AWP on, perturb / noise on) + MSE x3 + TV + bwd + flat gradient all-reduce "
                                        "(NCCL, avg) + sync-BN exchange + Adam over 36.9M parameters")
        # (c) SURVEY 8(e)'s partition: the SAME 4096-ray batch split across the ranks (strong scaling), forward and training
        n_loc = N_RAYS // world
        rays_s, idx_s = make_rays(N_RAYS, seed=4242)
        rays_s, idx_s = rays_s[rank * n_loc:(rank + 1) * n_loc].to(dev), idx_s[rank * n_loc:(rank + 1) * n_loc].to(dev)
        ms = timed_steps(lambda i: step(rays_s, idx_s), k2, w2, world, dev, before_each=flush_fn)
        strong = {"rays_total": n_loc * world, "rays_per_rank": n_loc, "fwd_ms_per_step": ms, "fwd_value": n_loc * world / (ms / 1e3), "unit": "rays/s"}
        tl = train_leg(P_all, dev, args.precision, n_loc, world, rank, k2, w2, flush)
        strong.update(train_ms_per_step=tl["ms_per_step"], train_value=tl["value"], train_loss_finite=tl["loss_finite"],
                      train_all_reduce_ms=tl["all_reduce_ms"])
        extra["strong"] = strong
        del P_all

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    ms_per_step = total_ms / args.steps
    value = world * N_RAYS / (ms_per_step / 1e3)
    e2e_value = world * N_RAYS / (e2e_ms / args.steps / 1e3)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops (burst)" if "bf16_tflops" in peaks else "fallback 1.59 PFLOP/s"
    R = N_RAYS * N_EXPOSURE
    fine_ms = kern_ms.get("fine", float("nan"))
    traffic, traffic_src = None, None   # dram__bytes_read.sum + dram__bytes_write.sum of the fine kernel: committed ncu --set full capture
    for name in ("r2_tc_kernels_ncu_summary.txt", "r1_tc_kernels_ncu_summary.txt"):
        try:
            txt = open(os.path.join(ROOT, "profiles", name)).read().split("=" * 100)
            want = {"bf16": ("fine_fwd_tc2_kernel", "fine_fwd_tc_kernel"), "tc32": ("fine_fwd_tc3_kernel",)}.get(args.precision, ())
            blk = next(b for b in txt if any(w in b for w in want))
            unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            tot = 0.0
            for key in ("dram__bytes_read.sum ", "dram__bytes_write.sum "):
                ln = next(l for l in blk.splitlines() if l.startswith(key)).split()
                tot += float(ln[1]) * unit[ln[2]]
            traffic, traffic_src = tot, "profiles/" + name
            break
        except Exception:
            continue
    achieved = flops_fine_kernel_per_subray() * R / (fine_ms / 1e3) / 1e12
    line = {
        "metric": "primary_rays_per_sec_fwd", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"bf16": "bf16", "fp32": "f32", "tc32": "f32 via bf16x3 split operands on tcgen05"}[args.precision], "data": "synthetic",
        "config": workload_config(args.precision),
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": rays_host.numel() * 4 + idx_host.numel() * 8,
                "d2h_bytes_per_step": out_host.numel() * 4},
        # rbk_warp_ndc, coarse, sample_pdf_merge, fine, NaN/Inf guard, 2 x weighted_sum (+ heads_stage in front of the bf16 fine kernel);
        # checked against profiles/r2_launches_bench_bf16.csv
        "gpu_launches": (8 if args.precision == "bf16" else 7) * args.steps,
        "clocks": clocks,
        "kernels_ms": kern_ms,
        "numerical_errors": headline_errs,
        "roofline": {"bound": "tensor", "kernel": "edn_render_fine_fwd", "achieved": achieved, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                     "traffic_unit": "bytes/launch (ncu dram read+write, parsed from the committed capture, not measured in this run)",
                     "traffic_source": traffic_src, "peak_source": peak_src,
                     "peak_note": "a burst figure against the burst peak: the timed region is tens of ms, power never reaches the cap",
                     "flops_per_launch": flops_fine_kernel_per_subray() * R, "ms_per_launch": fine_ms,
                     "whole_step_tflops": flops_per_subray() * R / (ms_per_step / 1e3) / 1e12},
        "wall_s_timed_region": t_wall,
    }
    line.update(extra)
    del nerf
    torch.cuda.empty_cache()
    if not args.no_cpu_baseline and world == 1:      # reported baselines: rank 0 at N = 1 only
        # the reference's CPU path in a FRESH process (its harness turns Tensor.cuda into a no-op: never in this process)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1"],
                               capture_output=True, text=True, timeout=600)
            ref_line = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
            line["cpu_baseline"] = dict(ref_line["cpu_baseline"], ms_per_step=ref_line["ms_per_step"])
            line["cpu_baseline"]["sample"] += "; mean of 3 steps after 1 warm-up, all host threads"
        except Exception as exc:
            line["cpu_baseline"] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
    if not args.no_cpu_baseline and not args.no_gpu_eager_baseline and world == 1:
        try:
            line["gpu_eager_reference"] = gpu_eager_reference(P, dev, N_RAYS)
        except Exception as exc:     # a reported extra, never the measured arm: an OOM here must not lose the bench line
            line["gpu_eager_reference"] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
