"""Oracle restatement of the deformable sparse kernel (DSK) against vectors written by the UNMODIFIED reference
(oracle/make_golden_dsk.py -> tests/golden/case9_dsk.npz): forward outputs and, through torch autograd on the oracle,
every parameter gradient of a fixed scalar functional.  Runs on any CPU box."""
import pytest
import torch

import evdeblur_oracle as oc
from util import golden

H = W = 400
KMAT = torch.tensor([[400.0, 0, 200.0], [0, 400.0, 200.0], [0, 0, 1.0]])
DSK_CFG = {   # the option sets of oracle/make_golden_dsk.py
    "a": dict(num_pt=5, kernel_hwindow=10, in_embed=3, spatial_embed=0, num_hidden=3, short_cut=False, isglobal=False, optim_trans=False, optim_sv_trans=False),
    "b": dict(num_pt=5, kernel_hwindow=10, in_embed=2, spatial_embed=2, num_hidden=2, short_cut=True, isglobal=False, optim_trans=True, optim_sv_trans=False),
    "c": dict(num_pt=5, kernel_hwindow=10, in_embed=3, spatial_embed=0, num_hidden=3, short_cut=False, isglobal=True, optim_trans=False, optim_sv_trans=True),
}


def case(g, name):
    pre = name + "."
    P = {"kernelsnet." + k[len(pre) + 6:]: v for k, v in g.items() if k.startswith(pre + "param.")}
    G = {k[len(pre) + 5:]: v for k, v in g.items() if k.startswith(pre + "grad.")}
    io = {k[len(pre):]: v for k, v in g.items() if k.startswith(pre) and ".param." not in k and ".grad." not in k}
    return P, G, io


@pytest.mark.parametrize("name", ["a", "b", "c"])
def test_oracle_dsk_matches_the_reference(name):
    P, G, io = case(golden("case9_dsk"), name)
    Pg = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in P.items()}
    new_rays, weight, align = oc.dsk_forward(Pg, DSK_CFG[name], H, W, KMAT, io["rays_x"], io["rays_y"], io["images_idx"], io["poses"], io.get("noise"))
    assert torch.allclose(new_rays, io["new_rays"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(weight, io["weight"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(align.reshape(1), io["align"], rtol=1e-6, atol=1e-7)
    loss = (new_rays * io["G_rays"]).sum() + (weight * io["G_w"]).sum() + align * io["g_align"][0]
    names = [k for k in G]
    grads = torch.autograd.grad(loss, [Pg["kernelsnet." + k] for k in names], allow_unused=True)
    for k, gr in zip(names, grads):
        ref = G[k]
        gr = torch.zeros_like(ref) if gr is None else gr
        assert torch.allclose(gr, ref, rtol=1e-4, atol=1e-5 * float(ref.abs().max() + 1e-12)), k


def test_oracle_awp_on_nerf_mode_features_matches_the_reference():
    """case 10 (oracle/make_golden_nerf_awp.py): the oracle's AWP restatement on the 256-channel trunk features of mode = nerf reproduces
    the reference NeRFAll's ccw_fine / rgb_awp (MLP weights of case 5 for both passes, depths from the fixture)."""
    g5, g = golden("case5_nerf24"), golden("case10_nerf_awp")
    P = {k[2:]: v for k, v in g5.items() if k.startswith("P.mlp_fine.")}
    P.update({k[2:]: v for k, v in g.items() if k.startswith("P.")})
    new_rays, weight1, emb = oc.rbk_forward(P, g["rays"], g["images_idx"], 4)
    rb = oc.build_ray_batch(400, 400, 400.0, new_rays.reshape(-1, 3, 2))
    o, d, vd, z = rb[:, :3], rb[:, 3:6], rb[:, -3:], g["z_vals"]
    with torch.no_grad():
        raw, feat = oc.nerf_mlpforward(P, "mlp_fine.", o[:, None] + d[:, None] * z[..., None], vd)
        assert torch.allclose(feat.double().sum(-1).float(), g["depth_feature_sum"], rtol=1e-4, atol=1e-3)
        rgb_s = oc.nerf_raw2outputs(raw, z, d)[0]
        ccw = oc.awp_forward(P, feat, z, d, emb, 5)
        ccw = ccw + ccw * 0.05
        ccw = ccw / ccw.sum(-1, keepdim=True)
    assert torch.allclose(ccw, g["ccw_fine"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(oc.rbk_weighted_sum(rgb_s, ccw), g["rgb_awp"], rtol=1e-4, atol=1e-5)
    assert torch.allclose(oc.rbk_weighted_sum(rgb_s, weight1), g["rgb"], rtol=1e-4, atol=1e-5)
