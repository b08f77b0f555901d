"""dev tooling: one bench-sized fine pass with EDN_TC_TRACE=1 (prints the in-kernel clock64 time line of CTA 0 to stderr)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from evdeblurnerf_b200 import NeRFAll
dev = torch.device("cuda")
P = bench.make_params(dev)
nerf = NeRFAll(P, *bench.AABB, kernel_ptnum=bench.N_EXPOSURE, precision="bf16").eval()
rays, idx = bench.make_rays(bench.N_RAYS, 1)
rays, idx = rays.to(dev), idx.to(dev)
for _ in range(2):
    nerf.render_blurred(bench.H, bench.W, bench.KMAT, rays, idx, N_samples=bench.NC, N_importance=bench.NI)
torch.cuda.synchronize()
os.environ["EDN_TC_TRACE"] = "1"
nerf.render_blurred(bench.H, bench.W, bench.KMAT, rays, idx, N_samples=bench.NC, N_importance=bench.NI)
torch.cuda.synchronize()
