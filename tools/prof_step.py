"""Profiling driver: runs the bf16 step a few times (used under ncu; numbers printed here are never bench values)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from evdeblurnerf_b200 import RenderEngine
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
P = bench.make_params(dev)
eng = RenderEngine(P, *bench.AABB, precision=prec)
rb = bench.build_ray_batch(bench.make_rays(bench.N_RAYS, seed=1000)).to(dev)
for _ in range(n):
    eng.render_rays(rb, bench.NC, N_importance=bench.NI, is_train=False)
torch.cuda.synchronize()
print("done")
