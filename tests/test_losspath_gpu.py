"""GPU parity tests of the blur-kernel / loss path against the oracle and the
golden vectors written from the unmodified reference.  fp32; tolerance 1e-4 relative (north_star)."""
import numpy as np
import pytest
import torch

import evdeblur_oracle as oc
from util import AABB, CFG, FOCAL, H, W, assert_close, golden, small_params, synthetic_rays

pytestmark = pytest.mark.gpu
KMAT = [[FOCAL, 0, 200.0], [0, FOCAL, 200.0], [0, 0, 1.0]]


@pytest.fixture(scope="module")
def cuda_params():
    P, Pc = small_params()
    return P, Pc, {k: v.cuda() for k, v in P.items()}, {k: v.cuda() for k, v in Pc.items()}


def test_rbk_warp_and_ray_batch(cuda_params):
    from evdeblurnerf_b200 import RigidBlurringModel, build_ray_batch
    P, _, Pg, _ = cuda_params
    g = golden("case1_train48x5")
    rbk = RigidBlurringModel(Pg, 4)
    o = rbk.warp(H, W, FOCAL, g["rays"].cuda(), g["images_idx"].cuda())
    assert_close(o["new_rays"], g["new_rays"], "new_rays", atol=1e-6)
    assert_close(o["weight"], g["weight1"], "weight", atol=1e-7)
    assert torch.equal(o["img_embed"].cpu(), g["img_embed"])
    rb = oc.build_ray_batch(H, W, FOCAL, g["new_rays"].reshape(-1, 3, 2))
    assert_close(o["ray_batch"], rb, "ray_batch (fused)", atol=2e-6)
    assert_close(build_ray_batch(H, W, FOCAL, g["new_rays"].cuda().reshape(-1, 3, 2)), rb, "ray_batch", atol=2e-6)
    nr, wgt, align, ex = rbk(H, W, KMAT, g["rays"].cuda(), {"images_idx": g["images_idx"].cuda()}, return_img_embed=True)
    assert align is None and "img_embed" in ex and nr.shape == (48, 5, 3, 2)


def test_rbk_identity_and_single_exposure(cuda_params):
    from evdeblurnerf_b200 import RigidBlurringModel
    _, _, Pg, _ = cuda_params
    Pz = dict(Pg)
    for k in list(Pz):
        if "r_linear" in k or "v_linear" in k:
            Pz[k] = torch.zeros_like(Pz[k])
    rays, idx = synthetic_rays(17, seed=4)
    o = RigidBlurringModel(Pz, 4).warp(H, W, FOCAL, rays.cuda(), idx.cuda())
    # zero twist -> theta = 1e-10 -> identity warp up to rounding (rigid_warping.py:24)
    assert_close(o["new_rays"], rays[:, None].expand(17, 5, 3, 2), "identity warp", rtol=1e-5, atol=1e-6)
    assert_close(o["weight"].sum(-1), torch.ones(17), "weights sum to 1", rtol=1e-5)


def test_forward_train_blend_and_tv(cuda_params):
    from evdeblurnerf_b200 import NeRFAll
    P, _, Pg, _ = cuda_params
    g = golden("case1_train48x5")
    nerf = NeRFAll(Pg, *AABB, kernel_ptnum=5, precision="fp32").train()
    rgb, rgb1, other_loss, other = nerf(H, W, KMAT, chunk=32768, rays=g["rays"].cuda(), rays_info={"images_idx": g["images_idx"].cuda()},
                                        force_naive=False, return_pts0_rgb=True, retraw=True, N_samples=64, N_importance=64,
                                        perturb=0., raw_noise_std=0., use_viewdirs=True, white_bkgd=False, inference=False,
                                        near=0., far=1.)
    assert_close(rgb, g["rgb"], "blended rgb", rtol=1e-4, atol=2e-4)
    assert_close(rgb1, g["rgb1"], "blended rgb1", rtol=1e-4, atol=2e-5)
    assert_close(other["stage1_rgb_pts0"], g["stage1_rgb_pts0"], "stage1_rgb_pts0", rtol=1e-4, atol=2e-4)
    assert_close(other["stage1_rgb1_pts0"], g["stage1_rgb1_pts0"], "stage1_rgb1_pts0", rtol=1e-4, atol=2e-5)
    assert_close(other_loss["TV"], g["TV"], "TV", rtol=1e-5)
    # force_naive branch (the event-ray calls, run_nerf.py:534,547)
    rgb_n, rgb0_n, ol, ot = nerf(H, W, KMAT, chunk=32768, rays=g["new_rays"].cuda().reshape(-1, 3, 2)[:50], rays_info=None,
                                 force_naive=True, retraw=True, N_samples=64, N_importance=64, perturb=0., raw_noise_std=0.)
    assert_close(rgb_n, g["rgb_map"][:50], "naive rgb", rtol=1e-4, atol=2e-4)
    assert_close(rgb0_n, g["rgb0"][:50], "naive rgb0", rtol=1e-4, atol=2e-5)


def test_weighted_sum_shapes():
    from evdeblurnerf_b200 import weighted_sum
    g = torch.Generator().manual_seed(0)
    x, w = torch.randn(35, 7, 3, generator=g), torch.rand(7, 5, generator=g)
    assert_close(weighted_sum(x.cuda(), w.cuda()), oc.rbk_weighted_sum(x, w), "ws 3d", rtol=1e-6, atol=1e-6)
    assert_close(weighted_sum(x[:, 0, 0].contiguous().cuda(), w.cuda()), oc.rbk_weighted_sum(x[:, 0, 0], w), "ws 1d", rtol=1e-6, atol=1e-6)


def test_crf_egm_mse_golden(cuda_params):
    from evdeblurnerf_b200 import TonemappingTransform, egm_loss, img2mse
    _, _, _, Pcg = cuda_params
    g = {k: v.cuda() for k, v in golden("case3_loss").items()}
    crf = TonemappingTransform(Pcg, map_type_rgb="gamma", map_type_event="learn", extra_features_event=2, gamma=2.2)
    assert_close(crf(g["x0"], mode="encode_rgb", skip_learn_crf=False), g["enc"], "encode_rgb", rtol=1e-5)
    l0 = crf(g["x0"], mode="encode_luma", skip_learn_crf=False, ev_extra_feat=g["pol"])
    l1 = crf(g["x1"], mode="encode_luma", skip_learn_crf=False, ev_extra_feat=g["pol"])
    assert_close(l0, g["l0"], "l0", rtol=1e-5); assert_close(l1, g["l1"], "l1", rtol=1e-5)
    assert_close(crf(g["x0"], mode="encode_luma", skip_learn_crf=True, ev_extra_feat=g["pol"]), g["l0s"], "l0s", rtol=1e-5)
    c0 = crf(g["x0"], mode="encode_luma", skip_learn_crf=False, ev_extra_feat=g["cpol"], tonemap_only=True)
    c1 = crf(g["x1"], mode="encode_luma", skip_learn_crf=False, ev_extra_feat=g["cpol"], tonemap_only=True)
    assert_close(c0, g["c0"], "c0", rtol=1e-5); assert_close(c1, g["c1"], "c1", rtol=1e-5)
    assert_close(crf(g["x0"], mode="encode_luma", skip_learn_crf=False), g["l0n"], "l0n (zero padded feats)", rtol=1e-5)
    assert_close(egm_loss(l0, l1, g["bii"]), g["e_gray"], "egm gray", rtol=1e-5)
    assert_close(egm_loss(c0, c1, g["bii"], color_mask=g["cmask"], color_weight=[0.4, 0.2, 0.4]), g["e_col"], "egm colour", rtol=1e-5)
    assert_close(img2mse(crf(g["x0"], mode="encode_rgb"), g["target"]), g["mse"], "img2mse", rtol=1e-5)


def test_crf_none_type(cuda_params):
    from evdeblurnerf_b200 import TonemappingTransform
    _, _, _, Pcg = cuda_params
    crf = TonemappingTransform(Pcg, map_type_rgb="none", map_type_event="gamma")
    x = torch.rand(33, 3).cuda()
    assert torch.equal(crf(x, mode="encode_rgb"), x)
    xl = x.cpu() ** (1 / 2.2)
    assert_close(crf(x, mode="encode_luma"), 0.299 * xl[:, :1] + 0.587 * xl[:, 1:2] + 0.114 * xl[:, 2:], "gamma luma", rtol=1e-5)


def test_edi_prior_golden():
    from evdeblurnerf_b200 import edi_prior_image
    g = {k: v.numpy() for k, v in golden("case4_edi").items()}
    sharp, bii = edi_prior_image(g["ev_x"], g["ev_y"], g["ev_t"], g["ev_p"], g["blurry"], float(g["t0"]), float(g["t1"]),
                                 float(g["cpos"]), float(g["cneg"]), int(g["steps"]))
    assert_close(bii, g["bii"], "bii", rtol=1e-5, atol=1e-5)          # fp32 atomics: summation order differs
    assert_close(sharp, g["sharp"], "sharp", rtol=1e-4, atol=1e-6)


def test_edi_no_events_is_identity():
    from evdeblurnerf_b200 import edi_prior_image
    blurry = np.random.default_rng(0).uniform(0.1, 1, (8, 9, 3)).astype(np.float32)
    ev = np.zeros((0,), np.float32)
    sharp, bii = edi_prior_image(ev, ev, np.zeros((0,)), ev, blurry, 0.0, 1.0, 0.2, 0.2, 9)
    assert float(bii.abs().max()) == 0.0
    assert_close(sharp, blurry, "sharp == blurry", rtol=1e-6)


def test_tv_loss(cuda_params):
    from evdeblurnerf_b200 import tv_loss_app
    P, _, Pg, _ = cuda_params
    for pre in ("mlp_coarse.", "mlp_fine."):
        assert_close(tv_loss_app(Pg, pre), oc.tv_loss_app(P, pre), "TV " + pre, rtol=1e-5)


def test_render_path_eval(cuda_params):
    from evdeblurnerf_b200 import NeRFAll
    P, _, Pg, _ = cuda_params
    nerf = NeRFAll(Pg, *AABB, precision="fp32").eval()
    Hs, Ws = 6, 5
    K = [[7.0, 0, 2.5], [0, 7.0, 3.0], [0, 0, 1.0]]
    c2w = torch.tensor([[1.0, 0.02, 0.0, 0.05], [-0.02, 1.0, 0.01, -0.03], [0.0, -0.01, 1.0, 1.1]])
    rgbs, depths = nerf(Hs, Ws, K, chunk=4096, poses=[c2w], render_kwargs=dict(N_samples=64, N_importance=64, perturb=0., raw_noise_std=0.))
    assert rgbs.shape == (1, Hs, Ws, 3) and depths.shape == (1, Hs, Ws)
    i, j = torch.meshgrid(torch.linspace(0, Ws - 1, Ws), torch.linspace(0, Hs - 1, Hs), indexing="xy")
    dirs = torch.stack([(i + (0.5 - K[0][2])) / K[0][0], -(j + (0.5 - K[1][2])) / K[1][1], -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays = torch.stack([c2w[:3, -1].expand(rays_d.shape), rays_d], -1).reshape(-1, 3, 2)
    rb = oc.build_ray_batch(Hs, Ws, K[0][0], rays)
    ref = oc.render_rays(P, CFG, rb, 64, 64, is_train=False)
    assert_close(rgbs[0].reshape(-1, 3), ref["rgb_map"], "eval rgb", rtol=1e-4, atol=5e-4)


def test_awp_forward_against_oracle(cuda_params):
    """Row a11: AdaptiveWeightProposal.forward (train-mode BatchNorm) on the oracle's own depth_feature."""
    from evdeblurnerf_b200.renderer import AdaptiveWeightProposal
    P, _, Pg, _ = cuda_params
    g = golden("case1_train48x5")
    rb = oc.build_ray_batch(H, W, FOCAL, g["new_rays"].reshape(-1, 3, 2))
    ref = oc.render_rays(P, CFG, rb, 64, 64, want_feature=True)
    ccw_ref = oc.awp_forward(P, ref["depth_feature"], ref["z_vals"], rb[:, 3:6], g["img_embed"], 5)
    awp = AdaptiveWeightProposal(Pg, 4)
    ccw = awp(ref["depth_feature"].cuda(), ref["z_vals"].cuda(), rb.cuda()[:, 3:6], g["img_embed"].cuda())
    assert_close(ccw, ccw_ref, "ccw", rtol=1e-4, atol=1e-6)
    assert_close(ccw.sum(-1), torch.ones(48), "rows sum to 1", rtol=1e-5)
    # bf16-mode path: the per-sample MLP as TF32 tensor-core GEMMs (10-bit mantissa operands): 2e-3 relative
    from evdeblurnerf_b200 import _lib
    awp_tc = AdaptiveWeightProposal(Pg, 4, precision=_lib.EDN_BF16)
    ccw_tc = awp_tc(ref["depth_feature"].cuda(), ref["z_vals"].cuda(), rb.cuda()[:, 3:6], g["img_embed"].cuda())
    assert_close(ccw_tc, ccw_ref, "ccw (TF32 GEMM path)", rtol=2e-3, atol=1e-5)


def test_forward_train_with_awp_golden(cuda_params):
    from evdeblurnerf_b200 import NeRFAll
    _, _, Pg, _ = cuda_params
    g = golden("case1_train48x5")
    nerf = NeRFAll(Pg, *AABB, kernel_ptnum=5, precision="fp32", use_awp=True).train()
    rgb, rgb1, other_loss, other = nerf(H, W, KMAT, chunk=32768, rays=g["rays"].cuda(), rays_info={"images_idx": g["images_idx"].cuda()},
                                        force_naive=False, return_pts0_rgb=True, retraw=True, N_samples=64, N_importance=64,
                                        perturb=0., raw_noise_std=0.)
    assert_close(other["ccw_fine"], g["ccw_fine"], "ccw_fine", rtol=2e-4, atol=2e-5)
    assert_close(other["rgb_awp"], g["rgb_awp"], "rgb_awp", rtol=1e-4, atol=2e-4)
    assert_close(rgb, g["rgb"], "rgb", rtol=1e-4, atol=2e-4)
