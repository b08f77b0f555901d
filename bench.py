#!/usr/bin/env python
"""Headline benchmark of the fused render path (BASELINE.json: primary rays/s, 4096-ray / 64+64-sample / 5-exposure).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision bf16|fp32]

One "step" = one forward pass of the hot path over one synthetic batch: 4096 primary rays -> 5 exposures (20480
sub-rays) -> coarse pass (64 samples) -> hierarchical sampling -> fine pass (128 samples) -> composited colours.
N > 1 (torchrun): every rank renders its own 4096-ray batch (weak scaling, no data-path collective); value = all
rays / max-over-ranks time.  Prints ONE JSON line on rank 0.
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


def cpu_reference_step(P_cpu, rays_cpu, idx_cpu, threads):
    """One bounded CPU step of the reference algorithm (oracle port; the reference is Python and cannot travel):
    blur-kernel warp -> NDC ray batch -> c2f render of the N*E sub-rays -> exposure-weighted sum."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import evdeblur_oracle as oc
    torch.set_num_threads(threads)
    cfg = {"aabb_min": AABB[0], "aabb_max": AABB[1], "rmnearplane": 0}
    with torch.no_grad():
        t0 = time.perf_counter()
        out = oc.forward_train(P_cpu, cfg, H, W, FOCAL, rays_cpu, idx_cpu, N_EXPOSURE, NC, NI, use_awp=False)
        dt = time.perf_counter() - t0
    return dt, out


def gpu_eager_port(P_dev, dev, n_rays, steps=3):
    """Baseline leg, second figure: the same oracle port run as EAGER PyTorch fp32 on the B200 itself -- the stand-in for
    the reference's own single-GPU path (the reference is Python under /root/reference and cannot travel to the
    GPU box).  Like the reference (run_nerf.py:779) it runs with CUDA as the default device.  Forward (no_grad) of the bench
    workload, then forward + autograd backward of the image loss; CUDA-event timed, best of `steps` after one warm-up."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import evdeblur_oracle as oc
    cfg = {"aabb_min": AABB[0], "aabb_max": AABB[1], "rmnearplane": 0}
    rays, idx = make_rays(n_rays, seed=100)
    rays, idx = rays.to(dev), idx.to(dev)
    out = {"kind": "port", "what": "oracle restatement of the reference path as eager PyTorch fp32 on this GPU, whole batch in one chunk",
           "rays": n_rays, "unit": "rays/s"}

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
        def fwd():
            with torch.no_grad():
                oc.forward_train(P_dev, cfg, H, W, FOCAL, rays, idx, N_EXPOSURE, NC, NI, use_awp=False)
        ms = timed(fwd)
        out.update(fwd_ms=ms, value=n_rays / (ms / 1e3))
        leaves = {k: v.detach().clone().requires_grad_(True) for k, v in P_dev.items()}
        target = torch.rand(n_rays, 3, device=dev)

        def fwd_bwd():
            o = oc.forward_train(leaves, cfg, H, W, FOCAL, rays, idx, N_EXPOSURE, NC, NI, use_awp=False)
            loss = oc.img2mse(o["rgb"], target) + oc.img2mse(o["rgb1"], target)
            loss.backward()
            for v in leaves.values():
                v.grad = None
        ms = timed(fwd_bwd)
        out.update(fwd_bwd_ms=ms, fwd_bwd_value=n_rays / (ms / 1e3))
    del leaves
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port) on the host cores, bounded sample per step."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample_rays = 512
    P = make_params("cpu")
    rays, idx = make_rays(sample_rays, seed=100)
    times = []
    for i in range(args.warmup + args.steps):
        dt, _ = cpu_reference_step(P, rays, idx, threads)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / max(len(times), 1)
    val = sample_rays / (ms / 1e3)
    line = {"impl": "reference", "metric": "primary_rays_per_sec_fwd", "value": val, "unit": "rays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config("fp32", sample=f"{sample_rays} primary rays x {N_EXPOSURE} exposures per step"),
            "cpu_baseline": {"value": val, "unit": "rays/s", "cores": threads, "kind": "port",
                             "sample": f"{sample_rays} primary rays x {N_EXPOSURE} exposures, {NC}+{NI} samples, full-size VM grids"},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("EDN_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
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

    t = torch.tensor([total_ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms = t.tolist()
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
    traffic = None        # dram__bytes_read.sum + dram__bytes_write.sum of the fine kernel from the committed ncu --set full capture
    try:
        txt = open(os.path.join(ROOT, "profiles", "r1_tc_kernels_ncu_summary.txt")).read().split("=" * 100)
        blk = next(b for b in txt if "fine_fwd_tc_kernel" in b)
        unit = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        tot = 0.0
        for key in ("dram__bytes_read.sum ", "dram__bytes_write.sum "):
            ln = next(l for l in blk.splitlines() if l.startswith(key)).split()
            tot += float(ln[1]) * unit[ln[2]]
        traffic = tot if args.precision == "bf16" else None
    except Exception:
        traffic = None
    achieved = flops_fine_kernel_per_subray() * R / (fine_ms / 1e3) / 1e12
    line = {
        "metric": "primary_rays_per_sec_fwd", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
        "config": workload_config(args.precision),
        "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": rays_host.numel() * 4 + idx_host.numel() * 8,
                "d2h_bytes_per_step": out_host.numel() * 4},
        "gpu_launches": 6 * args.steps,   # rbk_warp_ndc, coarse, sample_pdf_merge, fine, 2 x weighted_sum
        "clocks": clocks,
        "kernels_ms": kern_ms,
        "roofline": {"bound": "tensor", "kernel": "edn_render_fine_fwd", "achieved": achieved, "peak": peak_tf,
                     "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
                     "peak_source": peak_src,
                     "flops_per_launch": flops_fine_kernel_per_subray() * R, "ms_per_launch": fine_ms,
                     "whole_step_tflops": flops_per_subray() * R / (ms_per_step / 1e3) / 1e12},
        "wall_s_timed_region": t_wall,
    }
    if not args.no_cpu_baseline and world == 1:      # reported baselines: rank 0 at N = 1 only
        threads = os.cpu_count() or 1
        sample = 512
        Pc = {k: v.cpu() for k, v in P.items()}
        rays_c, idx_c = make_rays(sample, seed=100)
        best, t_spent = None, 0.0
        for i in range(4):
            dt, _ = cpu_reference_step(Pc, rays_c, idx_c, threads)
            t_spent += dt
            if i > 0:
                best = dt if best is None else min(best, dt)
            if t_spent > 25.0 and best is not None:
                break
        line["cpu_baseline"] = {"value": sample / best, "unit": "rays/s", "cores": threads, "kind": "port",
                                "sample": f"{sample} primary rays x {N_EXPOSURE} exposures, {NC}+{NI} samples, same VM grids; "
                                          f"best of {i} after 1 warm-up"}
    if not args.no_cpu_baseline and not args.no_gpu_eager_baseline and world == 1:
        try:
            line["gpu_eager_port"] = gpu_eager_port(P, dev, N_RAYS)
        except Exception as exc:     # a reported extra, never the measured arm: an OOM here must not lose the bench line
            line["gpu_eager_port"] = {"error": f"{type(exc).__name__}: {str(exc)[:200]}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
