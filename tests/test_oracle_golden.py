"""CPU: the oracle restatement replays the committed golden vectors (written by oracle/make_golden.py from the
UNMODIFIED reference).  This is the pin of the oracle; it needs neither a GPU nor /root/reference."""
import numpy as np
import torch

import evdeblur_oracle as oc
from util import CFG, FOCAL, H, W, AABB, assert_close, golden, small_params

CFG_REFNORM = dict(CFG, pdf_norm="torch_sum")


def test_case0_coarse_render():
    P, _ = small_params()
    g = golden("case0_coarse256")
    rb = oc.build_ray_batch(H, W, FOCAL, g["rays"])
    assert_close(rb, g["ray_batch"], "ray_batch", rtol=1e-6)
    out = oc.render_rays(P, CFG, rb, 64, 0)
    for k in ("rgb_map", "depth_map", "acc_map", "weights", "z_vals"):
        assert_close(out[k], g[k], k, rtol=1e-5, atol=1e-6)


def test_case1_train_forward():
    P, _ = small_params()
    g = golden("case1_train48x5")
    out = oc.forward_train(P, CFG_REFNORM, H, W, FOCAL, g["rays"], g["images_idx"], 5, 64, 64)
    r = out["render"]
    assert_close(out["new_rays"], g["new_rays"], "new_rays", rtol=1e-5)
    assert_close(out["weight1"], g["weight1"], "weight1", rtol=1e-5)
    for k in ("rgb_map", "depth_map", "acc_map", "rgb0", "weights0", "z_vals", "weights", "z_std"):
        assert_close(r[k], g[k], k, rtol=2e-5, atol=2e-6)
    assert torch.equal(r["inds"], g["inds"])
    assert_close(r["depth_feature"][:8], g["depth_feature_head"], "depth_feature", rtol=2e-5, atol=2e-6)
    assert_close(out["rgb"], g["rgb"], "rgb"); assert_close(out["rgb1"], g["rgb1"], "rgb1")
    assert_close(out["rgb_awp"], g["rgb_awp"], "rgb_awp")
    tv = (oc.tv_loss_app(P, "mlp_coarse.") + oc.tv_loss_app(P, "mlp_fine.")) * 5
    assert_close(tv, g["TV"], "TV", rtol=1e-5)


def test_case2_injected_randomness():
    P, _ = small_params()
    g = golden("case2_perturb32")
    rand = {k: g[k] for k in ("t_rand", "noise0", "u", "noise1")}
    out = oc.render_rays(P, CFG_REFNORM, g["ray_batch"], 64, 64, perturb=1., rand=rand)
    for k in ("rgb_map", "depth_map", "acc_map", "rgb0", "weights0", "z_vals", "weights"):
        assert_close(out[k], g[k], k, rtol=2e-5, atol=2e-6)
    assert torch.equal(out["inds"], g["inds"])


def test_case3_loss_heads():
    _, Pc = small_params()
    g = golden("case3_loss")
    x0, x1, pol, cpol, cmask = g["x0"], g["x1"], g["pol"], g["cpol"], g["cmask"]
    assert_close(oc.encode_rgb(Pc, x0), g["enc"], "encode_rgb", rtol=1e-5)
    l0 = oc.encode_luma(Pc, x0, ev_extra_feat=pol)
    l1 = oc.encode_luma(Pc, x1, ev_extra_feat=pol)
    assert_close(l0, g["l0"], "l0", rtol=1e-5); assert_close(l1, g["l1"], "l1", rtol=1e-5)
    assert_close(oc.encode_luma(Pc, x0, ev_extra_feat=pol, skip_learn=True), g["l0s"], "l0s", rtol=1e-5)
    c0 = oc.encode_luma(Pc, x0, ev_extra_feat=cpol, tonemap_only=True)
    c1 = oc.encode_luma(Pc, x1, ev_extra_feat=cpol, tonemap_only=True)
    assert_close(c0, g["c0"], "c0", rtol=1e-5)
    assert_close(oc.encode_luma(Pc, x0), g["l0n"], "l0n", rtol=1e-5)
    assert_close(oc.egm_loss(l0, l1, g["bii"]), g["e_gray"], "egm gray", rtol=1e-5)
    assert_close(oc.egm_loss(c0, c1, g["bii"], cmask, [0.4, 0.2, 0.4]), g["e_col"], "egm col", rtol=1e-5)
    assert_close(oc.img2mse(oc.encode_rgb(Pc, x0), g["target"]), g["mse"], "mse", rtol=1e-5)


def test_case4_edi():
    g = {k: v.numpy() for k, v in golden("case4_edi").items()}
    sharp = oc.edi_prior_image(g["ev_x"], g["ev_y"], g["ev_t"], g["ev_p"], g["blurry"], float(g["t0"]), float(g["t1"]),
                               40, 32, float(g["cpos"]), float(g["cneg"]), int(g["steps"]))
    assert_close(sharp, g["sharp"], "edi sharp", rtol=1e-5)


def test_case5_nerf_mode():
    g = golden("case5_nerf24")
    Pn = {k[2:]: v for k, v in g.items() if k.startswith("P.")}
    rb = g["ray_batch"]
    o, d, vd = rb[:, :3], rb[:, 3:6], rb[:, -3:]
    pts = o[:, None] + d[:, None] * g["z_vals"][..., None]
    raw, feat = oc.nerf_mlpforward(Pn, "mlp_fine.", pts, vd)
    assert_close(raw, g["raw"], "raw", rtol=2e-5, atol=2e-6)
    out = oc.nerf_raw2outputs(raw, g["z_vals"], d)
    assert_close(out[0], g["rgb_map"], "rgb_map", rtol=2e-5); assert_close(out[3], g["weights"], "weights", rtol=2e-5, atol=2e-6)


def test_case6_vm_known_answers():
    P, _ = small_params()
    g = golden("case6_vm")
    assert_close(oc.vm_sample(P, "mlp_coarse.", g["pts"], *AABB), g["ft_coarse"], "grid_sample", rtol=1e-5, atol=1e-6)
    assert_close(oc.vm_sample_taps(P, "mlp_fine.", g["pts"], *AABB), g["ft_fine"], "taps", rtol=1e-4, atol=1e-6)


def test_searchsorted_right_known_answer():
    # independent known-answer check of the index rule the CUDA kernel implements: first idx with cdf[idx] > u
    bins = torch.linspace(0, 1, 5)[None]
    w = torch.tensor([[1.0, 0.0, 3.0, 0.0]])
    z, inds = oc.sample_pdf(bins, w, 5)
    cdf = torch.cat([torch.zeros(1), ((w[0] + 1e-5) / (w[0] + 1e-5).sum()).cumsum(0)])
    u = torch.linspace(0, 1, 5)
    exp = torch.tensor([int((cdf <= x).sum()) for x in u])
    assert torch.equal(inds[0], exp)


def test_batchgen_oracle_replays_reference_vectors():
    """case7: the batch-generation restatements against the vectors written from the reference's own get_rays_pix,
    gather_successor and SLERP / cubic-spline pose interpolator + recenter_poses."""
    import numpy as np
    import evdeblur_oracle as oc
    from util import golden
    g = golden("case7_batchgen")
    K = g["K"].tolist()
    b = oc.make_rgb_batch(g["ray_ids"], g["images"], g["poses"], K)
    assert torch.equal(b["rays"], g["rays"]) and torch.equal(b["rgbsf"], g["rgbsf"]) and torch.equal(b["images_idx"], g["images_idx"])
    idx, neg, pos = oc.gather_successor(g["q_idx"], g["q_hops"], g["succ"], g["pol"])
    assert torch.equal(idx, g["succ_idx"]) and torch.equal(neg, g["neg"]) and torch.equal(pos, g["pos"])
    poses = oc.interpolate_event_poses(oc.pose_interpolator(g["times"].numpy(), g["rots"].numpy(), g["trans"].numpy()), g["tq"].numpy(),
                                       float(g["bd_scale"]), g["recenter_c2w"].numpy())
    assert np.allclose(poses, g["event_poses"].numpy(), rtol=0, atol=1e-6)
    assert torch.equal(oc.rays_from_pixels(g["ev_xy"], K, g["event_poses"]), g["ev_rays"])
