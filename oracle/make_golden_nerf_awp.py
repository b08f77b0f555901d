"""Golden vectors for the AWP branch on mode = nerf features (256-channel depth_feature, run_nerf.py:203-212, networks/nerf.py:140-150),
produced by the UNMODIFIED reference NeRFAll (mode = nerf, RBK + AWP) imported from /root/reference in the build container.  The two
NeRF MLPs share case5's mlp_fine weights (already committed in tests/golden/case5_nerf24.npz), so this fixture only adds the kernel
net, the AWP net and the outputs.   Output: tests/golden/case10_nerf_awp.npz.   Run:  python oracle/make_golden_nerf_awp.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_harness as rh  # noqa: E402
import evdeblur_oracle as oc  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
H = W = 400
FOCAL = 400.0
KMAT = torch.tensor([[FOCAL, 0, 200.0], [0, FOCAL, 200.0], [0, 0, 1.0]])
E, N = 5, 12


def main():
    args = rh.blurfactory_args(E=E, mode="nerf", use_awp=True, rgb_add_bias=True)
    nerf, _ = rh.build_reference(args, seed=5)
    c5 = np.load(os.path.join(OUT, "case5_nerf24.npz"))
    sd = nerf.state_dict()
    for k in sd:
        if k.startswith("mlp_fine."):
            assert np.array_equal(sd[k].numpy(), c5["P." + k]), k          # same construction as case 5
    nerf.mlp_coarse.load_state_dict(nerf.mlp_fine.state_dict())            # one committed weight set serves both passes
    with torch.no_grad():                                                   # make the AWP net's small heads count
        g = torch.Generator().manual_seed(77)
        for n_, p in nerf.awpnet.named_parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
    nerf.train()
    rays, idx = rh.synthetic_rays(N, seed=31)
    kw = dict(force_naive=False, return_pts0_rgb=True, retraw=True, N_samples=64, N_importance=64, perturb=0., raw_noise_std=0.,
              use_viewdirs=True, white_bkgd=False, inference=False, near=0., far=1.)
    with torch.no_grad():
        rgb, rgb1, other_loss, other = nerf(H, W, KMAT, chunk=32768, rays=rays, rays_info={"images_idx": idx}, **kw)
        new_rays, weight1, _, kex = nerf.kernelsnet(H, W, KMAT, rays, {"images_idx": idx}, return_img_embed=True)
        _, _, _, ex = nerf.render(H, W, KMAT, 32768, new_rays.reshape(-1, 3, 2), retraw=True, N_samples=64, N_importance=64, perturb=0.,
                                  raw_noise_std=0., use_viewdirs=True, white_bkgd=False, inference=False, near=0., far=1.)
    P = {k: v.detach().clone() for k, v in nerf.state_dict().items()}
    # oracle: AWP restatement on the reference's own depth_feature / z_vals (the nerf-mode render itself is pinned by case 5)
    rb = oc.build_ray_batch(H, W, FOCAL, new_rays.reshape(-1, 3, 2))
    with torch.no_grad():
        ccw = oc.awp_forward(P, ex["depth_feature"], ex["z_vals"], rb[:, 3:6], kex["img_embed"], E)
        ccw = ccw + ccw * 0.05
        ccw = ccw / ccw.sum(-1, keepdim=True)
    assert ex["depth_feature"].shape[-1] == 256
    print("  depth_feature", tuple(ex["depth_feature"].shape), "rgb_awp", tuple(other["rgb_awp"].shape))
    out = {"rays": rays, "images_idx": idx, "rgb": rgb, "rgb1": rgb1, "rgb_awp": other["rgb_awp"], "ccw_fine": ccw,
           "z_vals": ex["z_vals"], "depth_feature_sum": ex["depth_feature"].double().sum(-1).float()}
    # check the oracle's ccw against the reference's rgb_awp through the reference's own per-exposure colours
    with torch.no_grad():
        rgb_s, _, _, _ = nerf.render(H, W, KMAT, 32768, new_rays.reshape(-1, 3, 2), retraw=True, N_samples=64, N_importance=64, perturb=0.,
                                     raw_noise_std=0., use_viewdirs=True, white_bkgd=False, inference=False, near=0., far=1.)
    mine = oc.rbk_weighted_sum(rgb_s, ccw)
    err = (mine - other["rgb_awp"]).abs().max().item()
    print(f"  oracle AWP (256-channel features) vs reference rgb_awp: max abs diff {err:.3e}")
    assert err < 1e-5
    for k, v in P.items():
        if k.startswith(("kernelsnet.", "awpnet.")):
            out["P." + k] = v
    np.savez_compressed(os.path.join(OUT, "case10_nerf_awp.npz"), **{k: v.numpy() for k, v in out.items()})
    print("wrote case10_nerf_awp.npz", os.path.getsize(os.path.join(OUT, "case10_nerf_awp.npz")), "bytes")


if __name__ == "__main__":
    main()
