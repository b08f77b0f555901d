"""BASELINE config[5] sweep: primary rays {1k,4k,16k,65k} x samples {32+32, 64+64, 96+96} x exposures {1,5,9}, on 1 GPU or, under
torchrun, on N ranks (every rank renders its own batch of that size: weak scaling, no data-path collective; time = max over ranks).
Rank 0 prints one JSON line per point (whole-job rays/s, ms, fraction of the measured bf16 peak over the step's algorithmic FLOPs).
    python tools/sweep.py [--quick] > profiles/rN_sweep.jsonl
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/sweep.py --quick
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from evdeblurnerf_b200 import NeRFAll


def main():
    quick = "--quick" in sys.argv
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for E in (1, 5, 9):
        bench.N_EXPOSURE = E
        P = bench.make_params(dev)
        nerf = NeRFAll(P, *bench.AABB, kernel_ptnum=E, precision="bf16").eval()
        for nc, ni in ((32, 32), (64, 64), (96, 96)):
            for n in ((1024, 4096) if quick else (1024, 4096, 16384, 65536)):
                if (nc + ni) > 128 and n * E > 100000:
                    continue            # 192-sample rays run the fp32 fine kernel: keep the sweep short
                rays, idx = bench.make_rays(n, seed=7 + rank)
                rays, idx = rays.to(dev), idx.to(dev)
                step = lambda: nerf.render_blurred(bench.H, bench.W, bench.KMAT, rays, idx, N_samples=nc, N_importance=ni, perturb=0., raw_noise_std=0.)
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    flush.fill_(1)
                    if world > 1:
                        dist.barrier()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(); step(); e.record()
                    torch.cuda.synchronize()
                    ts.append(s.elapsed_time(e))
                t = torch.tensor([sum(ts) / len(ts)], device=dev)
                if world > 1:
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
                flops = bench.flops_per_subray(nc, ni) * n * E
                if rank == 0:
                    print(json.dumps({"n_gpus": world, "rays_per_gpu": n, "rays": n * world, "exposures": E, "samples": [nc, ni], "ms": ms,
                                      "rays_per_s": world * n / (ms / 1e3), "subrays_per_s": world * n * E / (ms / 1e3),
                                      "tflops_algorithmic_per_gpu": flops / (ms / 1e3) / 1e12,
                                      "frac_of_measured_bf16_peak": flops / (ms / 1e3) / 1e12 / peaks["bf16_tflops"],
                                      "fine_path": f"tcgen05 bf16, {(nc + ni + 127) // 128} x 128-row tile(s) per ray"}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
