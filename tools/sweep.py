"""BASELINE config[5] sweep on one GPU: primary rays {1k,4k,16k,65k} x samples {32+32, 64+64, 96+96} x exposures {1,5,9}.
Prints one JSON line per point (rays/s, ms, fraction of the measured bf16 peak over the whole step's algorithmic FLOPs).
    python tools/sweep.py [--quick] > profiles/rN_sweep.jsonl
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
    dev = torch.device("cuda", 0)
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
                rays, idx = bench.make_rays(n, seed=7)
                rays, idx = rays.to(dev), idx.to(dev)
                step = lambda: nerf.render_blurred(bench.H, bench.W, bench.KMAT, rays, idx, N_samples=nc, N_importance=ni, perturb=0., raw_noise_std=0.)
                for _ in range(3):
                    step()
                torch.cuda.synchronize()
                ts = []
                for _ in range(5):
                    flush.fill_(1)
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record(); step(); e.record()
                    torch.cuda.synchronize()
                    ts.append(s.elapsed_time(e))
                ms = sum(ts) / len(ts)
                flops = bench.flops_per_subray(nc, ni) * n * E
                print(json.dumps({"rays": n, "exposures": E, "samples": [nc, ni], "ms": ms, "rays_per_s": n / (ms / 1e3),
                                  "subrays_per_s": n * E / (ms / 1e3), "tflops_algorithmic": flops / (ms / 1e3) / 1e12,
                                  "frac_of_measured_bf16_peak": flops / (ms / 1e3) / 1e12 / peaks["bf16_tflops"],
                                  "fine_path": f"tcgen05 bf16, {(nc + ni + 127) // 128} x 128-row tile(s) per ray"}), flush=True)


if __name__ == "__main__":
    main()
