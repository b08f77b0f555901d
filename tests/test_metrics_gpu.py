"""Eval-path helpers added in round 2: image metrics (utils/metrics.py), the reference-style NeRFAll constructor / nn.Module surface
(networks/renderer.py:15-127), render(c2w_staticcam=...) (renderer.py:428-431)."""
import math

import pytest
import torch

import evdeblur_oracle as oc
from util import AABB, FOCAL, H, W, assert_close, small_params, synthetic_rays

pytestmark = pytest.mark.gpu
KMAT = [[FOCAL, 0, 200.0], [0, FOCAL, 200.0], [0, 0, 1.0]]


def test_mse_psnr_ssim_known_answers():
    from evdeblurnerf_b200.metrics import compute_img_metric
    g = torch.Generator().manual_seed(1)
    a = torch.rand(2, 40, 48, 3, generator=g).cuda()
    b = (a + 0.05 * torch.randn(2, 40, 48, 3, generator=g).cuda()).clamp(0, 1)
    d = ((a * 2 - 1) - (b * 2 - 1)).double()
    mse = float((d ** 2).reshape(2, -1).mean(1).mean())
    assert compute_img_metric(a, b, "mse") == pytest.approx(mse, rel=1e-5)
    psnr = sum(10 * math.log10(4.0 / float((d[i] ** 2).mean())) for i in range(2)) / 2
    assert compute_img_metric(a, b, "psnr") == pytest.approx(psnr, rel=1e-5)
    assert compute_img_metric(a, a, "ssim") == pytest.approx(1.0, abs=1e-6)
    s = compute_img_metric(a, b, "ssim")
    assert 0.5 < s < 1.0
    # direct evaluation of the SSIM definition on one 7 x 7 window (sample covariance, K1 = 0.01, K2 = 0.03, data range 2)
    x, y = (a[0, :7, :7, 0] * 2 - 1).double().flatten(), (b[0, :7, :7, 0] * 2 - 1).double().flatten()
    ux, uy = x.mean(), y.mean()
    vx, vy, vxy = x.var(unbiased=True), y.var(unbiased=True), ((x - ux) * (y - uy)).sum() / 48
    c1, c2 = (0.01 * 2) ** 2, (0.03 * 2) ** 2
    want = float((2 * ux * uy + c1) * (2 * vxy + c2) / ((ux ** 2 + uy ** 2 + c1) * (vx + vy + c2)))
    from evdeblurnerf_b200.metrics import _prep, _ssim_map
    pa, pb = _prep(a[:1], b[:1], None)
    assert float(_ssim_map(pa, pb)[0, 0, 0, 0]) == pytest.approx(want, rel=1e-4)
    # masked psnr: the reference's correction term
    mask = torch.zeros(2, 40, 48).cuda()
    mask[:, 5:30, 4:40] = 1
    pm = compute_img_metric(a, b, "psnr", mask=mask)
    dm = d * mask[..., None].double()
    want_m = sum(10 * math.log10(4.0 / float((dm[i] ** 2).mean())) - 10 * math.log10(40 * 48 / float(mask[i].sum())) for i in range(2)) / 2
    assert pm == pytest.approx(want_m, rel=1e-5)


def test_reference_style_constructor_and_module_surface():
    """NeRFAll(args, kernelsnet, awpnet) (renderer.py:15): reference state_dict names / shapes, parameters() usable by torch.optim,
    state_dict round trip, get_parameters() regular-expression groups (run_nerf.py:246-250)."""
    import reference_harness as rh
    from evdeblurnerf_b200 import NeRFAll
    P, _ = small_params()
    args = rh.blurfactory_args(E=5, coarse_n_voxels=20 * 20 * 14, fine_n_voxels=40 * 40 * 28, use_awp=True)
    kn = {k[len("kernelsnet."):]: v for k, v in P.items() if k.startswith("kernelsnet.")}
    awp = {k[len("awpnet."):]: v for k, v in P.items() if k.startswith("awpnet.")}
    nerf = NeRFAll(args, kn, awp, precision="fp32", seed=3)
    names = [k for k, _ in nerf.named_parameters()]
    assert "mlp_coarse.app_plane.0" in names and "mlp_fine.color_net.2.weight" in names and "kernelsnet.r_linear.weight" in names
    assert not any(k.endswith("running_mean") for k in names) and "awpnet.MAM.Corr.convd.1.running_mean" in nerf.state_dict()
    sd = nerf.state_dict()
    assert tuple(sd["mlp_fine.sigma_net.0.weight"].shape) == (256, 127) and tuple(sd["mlp_coarse.color_net.0.weight"].shape) == (64, 42)
    assert sd["mlp_fine.app_plane.0"].shape[1] == 64 and abs(float(sd["mlp_fine.app_plane.0"].std()) - 0.1) < 0.01
    wd = nerf.get_parameters("net", match_re=r"color_net\.[0-9]+\.weight")
    assert len(wd) == 6 and len(nerf.get_parameters("vol")) == 12
    # trains through torch.optim like the reference module
    opt = torch.optim.Adam(nerf.parameters(), lr=1e-3)
    rays, idx = synthetic_rays(16, seed=5)
    rgb, rgb0, other, tensors = nerf(H, W, KMAT, rays=rays.cuda(), rays_info={"images_idx": idx.cuda()}, retraw=True, force_naive=False,
                                     N_samples=32, N_importance=32, perturb=0., raw_noise_std=0.)
    loss = (rgb ** 2).mean() + (rgb0 ** 2).mean() + (tensors["rgb_awp"] ** 2).mean() + 1e-2 * other["TV"]
    before = sd["mlp_fine.color_net.2.weight"].clone()
    loss.backward()
    opt.step()
    assert float((nerf.state_dict()["mlp_fine.color_net.2.weight"] - before).abs().max()) > 0
    nerf.zero_grad()
    assert all(p.grad is None for p in nerf.parameters())
    # state_dict round trip into a second instance
    nerf2 = NeRFAll(args, kn, awp, precision="fp32", seed=4)
    nerf2.load_state_dict(nerf.state_dict())
    with torch.no_grad():
        a = nerf.eval().render_rays(oc.build_ray_batch(H, W, FOCAL, rays).cuda(), 32, N_importance=32)["rgb_map"]
        b = nerf2.eval().render_rays(oc.build_ray_batch(H, W, FOCAL, rays).cuda(), 32, N_importance=32)["rgb_map"]
    assert torch.equal(a, b)


def test_render_with_static_camera_view_directions():
    """renderer.py:428-431: c2w_staticcam -> camera rays of the static pose, view directions of `rays`."""
    from evdeblurnerf_b200 import NeRFAll
    from evdeblurnerf_b200.renderer import get_rays
    P, _ = small_params()
    nerf = NeRFAll({k: v.cuda() for k, v in P.items()}, *AABB, precision="fp32").eval()
    Hs = Ws = 12
    K = [[30.0, 0, 6.0], [0, 30.0, 6.0], [0, 0, 1.0]]
    pose = torch.tensor([[1.0, 0, 0, 0.02], [0, 1, 0, -0.01], [0, 0, 1, 1.0]])
    pose2 = torch.tensor([[0.995, 0, 0.0998, 0.1], [0, 1, 0, 0.0], [-0.0998, 0, 0.995, 1.0]])
    rays = get_rays(Hs, Ws, K, pose2.cuda())
    kw = dict(N_samples=32, N_importance=32, perturb=0., raw_noise_std=0., inference=True)
    rgb, depth, acc, _ = nerf.render(Hs, Ws, K, 1024, rays=rays, c2w_staticcam=pose, **kw)
    assert rgb.shape == (Hs, Ws, 3)
    # oracle: ray batch of the static pose with the view-direction columns of `rays`
    rb_s = oc.build_ray_batch(Hs, Ws, 30.0, get_rays(Hs, Ws, K, pose.cuda()).cpu().reshape(-1, 3, 2))
    rb_v = oc.build_ray_batch(Hs, Ws, 30.0, rays.cpu().reshape(-1, 3, 2))
    rb_s[:, 8:11] = rb_v[:, 8:11]
    ref = nerf.render_rays(rb_s.cuda(), **kw)
    assert_close(rgb.reshape(-1, 3), ref["rgb_map"], "static-camera render", rtol=1e-5, atol=1e-6)
    plain, _, _, _ = nerf.render(Hs, Ws, K, 1024, rays=rays, **kw)
    assert float((plain - rgb).abs().max()) > 1e-4
