"""GPU parity of mode = nerf ("run_network") against the golden vectors of the unmodified reference
and the oracle.  fp32, 1e-4 relative."""
import pytest
import torch

import evdeblur_oracle as oc
from util import AABB, assert_close, golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case5():
    g = golden("case5_nerf24")
    Pn = {k[2:]: v for k, v in g.items() if k.startswith("P.")}
    return g, Pn


def test_nerf_mlpforward_and_raw2outputs_golden(case5):
    from evdeblurnerf_b200.nerf_mode import NeRF
    g, Pn = case5
    net = NeRF({k: v.cuda() for k, v in Pn.items()}, "mlp_fine.")
    rb, z = g["ray_batch"].cuda(), g["z_vals"].cuda()
    raw, feat = net.mlpforward_at(rb, z, want_feature=True)
    assert_close(raw, g["raw"], "raw", rtol=1e-4, atol=2e-5)
    assert_close(feat.double().sum(-1).float(), g["feature_sum"], "feature (before_linear) row sums", rtol=1e-4, atol=1e-3)
    rgb, depth, acc, w = net.raw2outputs(raw, z, rb)
    assert_close(rgb, g["rgb_map"], "rgb_map", rtol=1e-4, atol=2e-5)
    assert_close(depth, g["depth_map"], "depth_map", rtol=1e-4, atol=2e-5)
    assert_close(acc, g["acc_map"], "acc_map", rtol=1e-4, atol=2e-5)
    assert_close(w, g["weights"], "weights", rtol=1e-4, atol=2e-5)
    o = oc.nerf_raw2outputs(g["raw"], g["z_vals"], g["ray_batch"][:, 3:6], white_bkgd=True)
    rgb_w, _, _, _ = net.raw2outputs(g["raw"].cuda(), z, rb, white_bkgd=True)
    assert_close(rgb_w, o[0], "white background", rtol=1e-4, atol=2e-5)


def test_nerf_render_rays_c2f_sampling(case5):
    from evdeblurnerf_b200 import RenderEngine
    from evdeblurnerf_b200.nerf_mode import NeRF, render_rays_nerf
    from util import small_params
    g, Pn = case5
    Pg = {k: v.cuda() for k, v in Pn.items()}
    net = NeRF(Pg, "mlp_fine.")
    P, _ = small_params()
    eng = RenderEngine({k: v.cuda() for k, v in P.items()}, *AABB, precision="fp32")     # sampler host of the nerf-mode render
    rb = g["ray_batch"]
    out = render_rays_nerf(eng, net, net, rb.cuda(), 64, retraw=True, N_importance=64)
    # oracle: same field for the coarse and the fine pass
    z0 = oc.place_samples(rb[:, 6:7], rb[:, 7:8], 64)
    assert torch.equal(out["z_vals0"].cpu(), z0)
    o, d, vd = rb[:, :3], rb[:, 3:6], rb[:, -3:]
    raw0, _ = oc.nerf_mlpforward(Pn, "mlp_fine.", o[:, None] + d[:, None] * z0[..., None], vd)
    r0 = oc.nerf_raw2outputs(raw0, z0, d)
    assert_close(out["rgb0"], r0[0], "rgb0", rtol=1e-4, atol=2e-5)
    assert_close(out["weights0"], r0[3], "weights0", rtol=1e-4, atol=2e-5)
    zf = out["z_vals"].cpu()
    raw1, _ = oc.nerf_mlpforward(Pn, "mlp_fine.", o[:, None] + d[:, None] * zf[..., None], vd)
    r1 = oc.nerf_raw2outputs(raw1, zf, d)
    assert_close(out["rgb_map"], r1[0], "rgb_map at the CUDA depths", rtol=1e-4, atol=2e-5)
    assert_close(out["weights"], r1[3], "weights", rtol=1e-4, atol=2e-5)
    assert_close(out["rgb_map"], g["rgb_map"], "golden rgb_map (end to end)", rtol=1e-4, atol=5e-4)


def test_nerf_ragged(case5):
    from evdeblurnerf_b200.nerf_mode import NeRF
    g, Pn = case5
    net = NeRF({k: v.cuda() for k, v in Pn.items()}, "mlp_fine.", extract_feature="after_linear")
    rb = g["ray_batch"][:5]
    z = torch.sort(torch.rand(5, 37, generator=torch.Generator().manual_seed(1)), -1)[0]
    raw, feat = net.mlpforward_at(rb.cuda(), z.cuda(), want_feature=True)
    o, d, vd = rb[:, :3], rb[:, 3:6], rb[:, -3:]
    raw_o, feat_o = oc.nerf_mlpforward(Pn, "mlp_fine.", o[:, None] + d[:, None] * z[..., None], vd, before_linear=False)
    assert_close(raw, raw_o, "raw", rtol=1e-4, atol=2e-5)
    assert_close(feat, feat_o, "feature (after_linear)", rtol=1e-4, atol=2e-5)


def test_nerfall_facade_nerf_mode(case5):
    from evdeblurnerf_b200 import NeRFAll
    g, Pn = case5
    Pg = {k: v.cuda() for k, v in Pn.items()}
    Pg.update({k.replace("mlp_fine.", "mlp_coarse."): v for k, v in Pg.items()})
    nerf = NeRFAll(Pg, *AABB).train()
    assert nerf.mode == "nerf"
    out = nerf.render_rays(g["ray_batch"].cuda(), 64, retraw=True, N_importance=64)
    assert_close(out["rgb_map"], g["rgb_map"], "facade rgb_map", rtol=1e-4, atol=5e-4)
