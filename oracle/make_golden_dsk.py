"""Golden vectors for the deformable sparse kernel (DSK) blur model, produced by the UNMODIFIED reference
(networks/pdrf/blurmodel.py BlurModel, networks/renderer.py NeRFAll with kernel_type = DSK) imported from /root/reference
in the build container.  Forward outputs AND the parameter gradients of a fixed scalar functional (torch autograd on the
reference module), for three option sets; one end-to-end NeRFAll.forward on the small grids of tests/golden/params_small.npz.
Output: tests/golden/case9_dsk.npz.   Run:  python oracle/make_golden_dsk.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import reference_harness as rh  # noqa: E402
import evdeblur_oracle as oc  # noqa: E402

rh._install_shims(True)
from networks.embedding import ViewEmbedding  # noqa: E402
from networks.pdrf.blurmodel import BlurModel  # noqa: E402
from networks.renderer import NeRFAll  # noqa: E402

OUT = os.path.join(HERE, "..", "tests", "golden")
H = W = 400
FOCAL = 400.0
KMAT = torch.tensor([[FOCAL, 0, W / 2], [0, FOCAL, H / 2], [0, 0, 1]], dtype=torch.float32)
N_IMG, NPT = 30, 5

CONFIGS = {
    # name: BlurModel keyword arguments (run_nerf.py:184-203 option names in comments)
    "a": dict(random_hwindow=0.25, in_embed=3, spatial_embed=0, num_hidden=3, num_wide=64, short_cut=False, isglobal=False,
              optim_trans=False, optim_spatialvariant_trans=False),                               # defaults of options.py
    "b": dict(random_hwindow=0.0, in_embed=2, spatial_embed=2, num_hidden=2, num_wide=64, short_cut=True, isglobal=False,
              optim_trans=True, optim_spatialvariant_trans=False),                                # kernel_shortcut, kernel_global_trans
    "c": dict(random_hwindow=0.0, in_embed=3, spatial_embed=0, num_hidden=3, num_wide=32, short_cut=False, isglobal=True,
              optim_trans=False, optim_spatialvariant_trans=True),                                # kernel_isglobal, kernel_spatialvariant_trans
}


def make_inputs(N, seed):
    g = torch.Generator().manual_seed(seed)
    rays_x = torch.rand(N, 1, generator=g) * (W - 1)
    rays_y = torch.rand(N, 1, generator=g) * (H - 1)
    idx = torch.randint(0, N_IMG, (N, 1), generator=g)
    ang = torch.randn(N, 3, generator=g) * 0.05
    poses = torch.zeros(N, 3, 4)
    for n in range(N):        # small rotations about identity + translations: camera looks down -z (NDC well posed)
        ax, ay, az = ang[n].tolist()
        Rx = torch.tensor([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]], dtype=torch.float32)
        Ry = torch.tensor([[np.cos(ay), 0, np.sin(ay)], [0, 1, 0], [-np.sin(ay), 0, np.cos(ay)]], dtype=torch.float32)
        Rz = torch.tensor([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]], dtype=torch.float32)
        poses[n, :, :3] = Rz @ Ry @ Rx
    poses[:, :, 3] = torch.randn(N, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0])
    return rays_x, rays_y, idx, poses


def build(cfg, seed):
    torch.manual_seed(seed)
    ve = ViewEmbedding(num_embed=N_IMG, embed_dim=32, init_params="normal")
    m = BlurModel(N_IMG, NPT, 10, "DSK", ve, img_wh=[W, H], view_embed_cnl=32, random_mode="input", depth_embed=0,
                  pattern_init_radius=0.1, **cfg)
    with torch.no_grad():      # the 2- / 3-row output layers start at xavier gain 0.1 and zero bias: make every path count
        for p in m.parameters():
            p.add_(torch.randn_like(p) * 0.05)
    return m


def oracle_cfg(cfg):
    return dict(num_pt=NPT, kernel_hwindow=10, in_embed=cfg["in_embed"], spatial_embed=cfg["spatial_embed"], num_hidden=cfg["num_hidden"],
                short_cut=cfg["short_cut"], isglobal=cfg["isglobal"], optim_trans=cfg["optim_trans"], optim_sv_trans=cfg["optim_spatialvariant_trans"])


def main():
    out = {}
    for ci, (name, cfg) in enumerate(CONFIGS.items()):
        N = 37
        m = build(cfg, 100 + ci).train()
        rays_x, rays_y, idx, poses = make_inputs(N, 200 + ci)
        info = {"rays_x": rays_x, "rays_y": rays_y, "images_idx": idx, "poses": poses}
        torch.manual_seed(300 + ci)
        new_rays, weight, align, _ = m(H, W, KMAT, None, info)
        noise = None
        if cfg["random_hwindow"] > 0:
            torch.manual_seed(300 + ci)          # replay blurmodel.py:126 (the only draw in forward)
            noise = torch.randn(N, NPT, 2) * cfg["random_hwindow"]
        g = torch.Generator().manual_seed(400 + ci)
        G_rays, G_w = torch.randn(new_rays.shape, generator=g), torch.randn(weight.shape, generator=g)
        g_align = 0.7
        loss = (new_rays * G_rays).sum() + (weight * G_w).sum() + align * g_align
        loss.backward()
        P = {"kernelsnet." + k: v.detach().clone() for k, v in m.state_dict().items()}
        mine = oc.dsk_forward(P, oracle_cfg(cfg), H, W, KMAT, rays_x, rays_y, idx, poses, noise)
        for a, b, nm in ((mine[0], new_rays, "new_rays"), (mine[1], weight, "weight"), (mine[2], align, "align")):
            err = (a - b.detach()).abs().max().item()
            print(f"  {name}: oracle vs reference {nm}: max abs diff {err:.3e}")
            assert err <= 1e-6 * max(1.0, b.detach().abs().max().item()), nm
        pre = f"{name}."
        out.update({pre + "rays_x": rays_x, pre + "rays_y": rays_y, pre + "images_idx": idx, pre + "poses": poses,
                    pre + "new_rays": new_rays.detach(), pre + "weight": weight.detach(), pre + "align": align.detach().reshape(1),
                    pre + "G_rays": G_rays, pre + "G_w": G_w, pre + "g_align": torch.tensor([g_align])})
        if noise is not None:
            out[pre + "noise"] = noise
        for k, v in m.state_dict().items():
            out[pre + "param." + k] = v.detach()
        for k, p in m.named_parameters():
            out[pre + "grad." + k] = p.grad.detach() if p.grad is not None else torch.zeros_like(p)

    # ---- end to end: NeRFAll(kernel_type = DSK).forward on the small grids, config "a" without the input noise -----------------------
    small = np.load(os.path.join(OUT, "params_small.npz"))
    args = rh.blurfactory_args(E=NPT, coarse_n_voxels=18 * 18 * 12, fine_n_voxels=36 * 36 * 24, use_awp=False)
    AABB = ((-1.5, -1.5, -1.0), (1.5, 1.5, 1.0))
    args.kernel_type = "DSK"
    cfg = dict(CONFIGS["a"], random_hwindow=0.0)
    kn = build(cfg, 777)
    torch.manual_seed(0)
    nerf = NeRFAll(args, kn, None)
    sd = nerf.state_dict()
    loaded = 0
    for k in sd:
        if k in small.files and tuple(small[k].shape) == tuple(sd[k].shape):
            sd[k] = torch.from_numpy(small[k]); loaded += 1
    nerf.load_state_dict(sd)
    print("  e2e: field tensors taken from params_small.npz:", loaded)
    assert loaded >= 20
    nerf.train()
    N = 24
    rays_x, rays_y, idx, poses = make_inputs(N, 900)
    info = {"rays_x": rays_x, "rays_y": rays_y, "images_idx": idx, "poses": poses}
    with torch.no_grad():
        rgb, rgb1, other_loss, other = nerf(H, W, KMAT, chunk=32768, rays=torch.zeros(N, 3, 2), rays_info=info, force_naive=False, return_pts0_rgb=True,
                                            retraw=True, N_samples=64, N_importance=64, perturb=0., raw_noise_std=0., ndc=True,
                                            near=0., far=1., use_viewdirs=True, lindisp=False, white_bkgd=False, inference=False)
    P = {k: v.detach().clone() for k, v in nerf.state_dict().items()}
    CFG = {"aabb_min": AABB[0], "aabb_max": AABB[1], "rmnearplane": 0, "pdf_norm": "torch_sum"}
    with torch.no_grad():
        mine = oc.forward_train_dsk(P, CFG, oracle_cfg(cfg), H, W, KMAT, rays_x, rays_y, idx, poses, 64, 64)
    for a, b, nm in ((mine["rgb"], rgb, "rgb"), (mine["rgb1"], rgb1, "rgb1"), (mine["align"], other_loss["align"].reshape(()), "align")):
        err = (a - b).abs().max().item()
        print(f"  e2e: oracle vs reference {nm}: max abs diff {err:.3e}")
        assert err <= 2e-5, nm
    out.update({"e2e.rays_x": rays_x, "e2e.rays_y": rays_y, "e2e.images_idx": idx, "e2e.poses": poses, "e2e.rgb": rgb, "e2e.rgb1": rgb1,
                "e2e.align": other_loss["align"].reshape(1), "e2e.TV": other_loss["TV"].reshape(1),
                "e2e.stage1_rgb_pts0": other["stage1_rgb_pts0"]})
    for k, v in kn.state_dict().items():
        out["e2e.param." + k] = v.detach()
    # ---- the same with the AWP branch on (renderer.py:310-343 with a non-RBK kernel): rgb_awp from AWP's weights over the DSK points --
    from networks.dpnerf.awp import AdaptiveWeightProposal
    args2 = rh.blurfactory_args(E=NPT, coarse_n_voxels=18 * 18 * 12, fine_n_voxels=36 * 36 * 24, use_awp=True)
    args2.kernel_type = "DSK"
    torch.manual_seed(1)
    awpnet = AdaptiveWeightProposal(input_ch=args2.fine_geo_feat_dim, num_motion=NPT - 1, use_origin=True, D_sam=4, W_sam=64, D_mot=1, W_mot=32,
                                    dir_freq=2, rgb_freq=2, depth_freq=3, ray_dir_freq=2, view_feature_ch=32)
    nerf2 = NeRFAll(args2, kn, awpnet)
    sd2 = nerf2.state_dict()
    loaded = 0
    for k in sd2:
        if k in small.files and tuple(small[k].shape) == tuple(sd2[k].shape) and not k.startswith("kernelsnet."):
            sd2[k] = torch.from_numpy(small[k]); loaded += 1
    nerf2.load_state_dict(sd2)
    print("  e2e + AWP: tensors taken from params_small.npz:", loaded)
    assert loaded >= 50
    nerf2.train()
    with torch.no_grad():
        rgb_a, rgb1_a, _, other_a = nerf2(H, W, KMAT, chunk=32768, rays=torch.zeros(N, 3, 2), rays_info=info, force_naive=False,
                                          return_pts0_rgb=True, retraw=True, N_samples=64, N_importance=64, perturb=0., raw_noise_std=0.,
                                          ndc=True, near=0., far=1., use_viewdirs=True, lindisp=False, white_bkgd=False, inference=False)
    P2 = {k: v.detach().clone() for k, v in nerf2.state_dict().items()}
    with torch.no_grad():
        mine2 = oc.forward_train_dsk(P2, CFG, oracle_cfg(cfg), H, W, KMAT, rays_x, rays_y, idx, poses, 64, 64, use_awp=True)
    for a, b, nm in ((mine2["rgb"], rgb_a, "rgb"), (mine2["rgb1"], rgb1_a, "rgb1"), (mine2["rgb_awp"], other_a["rgb_awp"], "rgb_awp")):
        err = (a - b).abs().max().item()
        print(f"  e2e + AWP: oracle vs reference {nm}: max abs diff {err:.3e}")
        assert err <= 5e-5, nm
    out.update({"e2e.awp.rgb": rgb_a, "e2e.awp.rgb1": rgb1_a, "e2e.awp.rgb_awp": other_a["rgb_awp"], "e2e.awp.ccw_fine": mine2["ccw_fine"]})
    np.savez_compressed(os.path.join(OUT, "case9_dsk.npz"), **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in out.items()})
    print("wrote case9_dsk.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
