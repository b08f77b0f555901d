import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
import evdeblur_oracle as oc
from util import *
from evdeblurnerf_b200 import RenderEngine
P, _ = small_params()
eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision="bf16")
rays, _ = synthetic_rays(int(sys.argv[2]) if len(sys.argv) > 2 else 150, seed=31)
rb = oc.build_ray_batch(H, W, FOCAL, rays)
awp = sys.argv[1] == "full"
out = eng.render_rays(rb.cuda(), 64, retraw=True, N_importance=64, use_awp=awp)
torch.cuda.synchronize()
z = out["z_vals"].cpu()
emu = emulated_bf16_fine(P, rb, z, lean=not awp)
for k in ("weights", "rgb_map"):
    print(k, (out[k].cpu() - emu[k]).abs().max().item())
