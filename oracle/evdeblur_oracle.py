"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32 on CPU + numpy) of the EvDeblurNeRF render / blur-loss path.

This file is the *oracle* the CUDA path is checked against.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import it; the product package `evdeblurnerf_b200`
never does (it fails loudly when its CUDA library is missing).

Every function cites the reference file:line (relative to uzh-rpg/EvDeblurNeRF @ 4111020) it restates.  The reference
is plain PyTorch, so the restatement is functional torch on CPU: parameters come in as a flat dict keyed by the
reference's own `state_dict()` names, tensors keep the reference's shapes and layouts.

Pinning: `oracle/make_golden.py` runs the UNMODIFIED reference (imported from /root/reference, this container only)
and this oracle on identical seeded inputs, asserts agreement and writes `tests/golden/*.npz`.  The reference has no
tests or golden vectors of its own, so those fixtures are the pin.

Specified arithmetic where the reference leaves it to the backend (see DESIGN.md "Numerics contract"):
  * sample_pdf: normaliser = fp32(sum of the 62 fp32 terms accumulated sequentially in fp64); cdf_k =
    fp32(prefix sums accumulated sequentially in fp64) -- this is what torch's CPU cumsum does (it accumulates
    fp32 in double) and makes `inds` platform independent; the CUDA kernel implements exactly this.
  * sort of the merged z values is stable (ties keep coarse-before-fine / lower index first).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

MATMODE = ((0, 1), (0, 2), (1, 2))  # voxnerf.py:99
VECMODE = (2, 1, 0)                 # voxnerf.py:100


# --------------------------------------------------------------------------------------------------------------
# a5  positional encoding                                                  networks/embedding.py:88-98, 101-115
# --------------------------------------------------------------------------------------------------------------
def posenc(x, L):
    """[x, sin(2^0 x), cos(2^0 x), ..., sin(2^(L-1) x), cos(2^(L-1) x)]; freq = 2**linspace(0, L-1, L)."""
    freqs = 2.0 ** torch.linspace(0.0, L - 1, steps=L)
    out = [x]
    for f in freqs:
        out.append(torch.sin(x * f))
        out.append(torch.cos(x * f))
    return torch.cat(out, -1)


# --------------------------------------------------------------------------------------------------------------
# a1  DP-NeRF rigid blur kernel              networks/dpnerf/blurmodel.py:129-173, 51-82; utils/rigid_warping.py
# --------------------------------------------------------------------------------------------------------------
def _skew(w):  # rigid_warping.py:116-132
    z = torch.zeros_like(w[..., 0])
    return torch.stack([torch.stack([z, -w[..., 2], w[..., 1]], -1),
                        torch.stack([w[..., 2], z, -w[..., 0]], -1),
                        torch.stack([-w[..., 1], w[..., 0], z], -1)], -2)


def se3_transform(rot, trans):
    """rigid_warping.py:18-30 (get_transform) + :75-95 (exp_se3) + :97-113 (exp_so3) -> R [n,3,3], p [n,3]."""
    theta = torch.linalg.norm(rot, dim=-1) + 1.0e-10
    w = rot / theta[..., None]
    v = trans / theta[..., None]
    Wm = _skew(w)
    th = theta[..., None, None]
    eye = torch.eye(3)[None]
    WW = torch.matmul(Wm, Wm)
    R = eye + torch.sin(th) * Wm + (1.0 - torch.cos(th)) * WW
    p = torch.matmul(th * eye + (1.0 - torch.cos(th)) * Wm + (th - torch.sin(th)) * WW, v[..., None])[..., 0]
    return R, p


def rbk_forward(P, rays, images_idx, num_motion, rv_window=0.1, prefix="kernelsnet."):
    """RigidBlurringModel.forward, feat_ch = 0, depth-1 branches, use_origin=True (all shipped configs).
    rays [N,3,2], images_idx [N] int64 -> new_rays [N,E,3,2], weight [N,E], img_embed [N,32]."""
    emb = P[prefix + "view_embed_module.img_embed"][images_idx.reshape(-1)]  # embedding.py:31-32

    def branch(name):
        return F.relu(F.linear(emb, P[prefix + f"{name}_branch.0.weight"], P[prefix + f"{name}_branch.0.bias"]))

    r = F.linear(branch("r"), P[prefix + "r_linear.weight"], P[prefix + "r_linear.bias"]) * rv_window
    v = F.linear(branch("v"), P[prefix + "v_linear.weight"], P[prefix + "v_linear.bias"]) * rv_window
    wgt = torch.sigmoid(F.linear(branch("w"), P[prefix + "w_linear.weight"], P[prefix + "w_linear.bias"]))
    wgt = wgt / (wgt.sum(-1, keepdim=True) + 1e-10)
    N = rays.shape[0]
    r = r.reshape(N, 3, num_motion)  # blurmodel.py:52-53 : component-major, motion-minor
    v = v.reshape(N, 3, num_motion)
    o, d = rays[..., 0], rays[..., 1]
    end = o + d
    out = [torch.stack([o, d], -1)]
    for i in range(num_motion):
        R, p = se3_transform(r[:, :, i], v[:, :, i])
        wo = torch.matmul(R, o[..., None])[..., 0] + p   # homogeneous w stays 1 (rigid_warping.py:44-47, 153-154)
        we = torch.matmul(R, end[..., None])[..., 0] + p
        out.append(torch.stack([wo, we - wo], -1))
    return torch.stack(out, 1), wgt, emb


def rbk_weighted_sum(x, ccw):
    """blurmodel.py:112-127 for one tensor: [N*E, ...] -> [N, ...] with weights ccw [N,E]."""
    N, E = ccw.shape
    xs = x.reshape(N, E, *x.shape[1:])
    w = ccw.reshape(N, E, *([1] * (xs.dim() - 2)))
    return (xs * w).sum(1)


# --------------------------------------------------------------------------------------------------------------
# a2  render() prologue + NDC                        networks/renderer.py:423-446; utils/rays.py:104-145
# --------------------------------------------------------------------------------------------------------------
def ndc_rays(H, W, focal, near, rays_o, rays_d):
    t = -(near + rays_o[..., 2]) / rays_d[..., 2]
    rays_o = rays_o + t[..., None] * rays_d
    ox_oz = rays_o[..., 0] / rays_o[..., 2]
    oy_oz = rays_o[..., 1] / rays_o[..., 2]
    o0 = -1. / (W / (2. * focal)) * ox_oz
    o1 = -1. / (H / (2. * focal)) * oy_oz
    o2 = 1. + 2. * near / rays_o[..., 2]
    d0 = -1. / (W / (2. * focal)) * (rays_d[..., 0] / rays_d[..., 2] - ox_oz)
    d1 = -1. / (H / (2. * focal)) * (rays_d[..., 1] / rays_d[..., 2] - oy_oz)
    d2 = 1 - o2
    return torch.stack([o0, o1, o2], -1), torch.stack([d0, d1, d2], -1)


def build_ray_batch(H, W, focal, rays, near=0.0, far=1.0, ndc=True, use_viewdirs=True):
    """rays [R,3,2] -> ray_batch [R, 8|11] = [o, d, near, far, viewdirs]; viewdirs from the pre-NDC direction."""
    o, d = rays[..., 0], rays[..., 1]
    viewdirs = d / torch.norm(d, dim=-1, keepdim=True)
    if ndc:
        o, d = ndc_rays(H, W, focal, 1.0, o, d)
    o, d = o.reshape(-1, 3).float(), d.reshape(-1, 3).float()
    cols = [o, d, near * torch.ones_like(d[..., :1]), far * torch.ones_like(d[..., :1])]
    if use_viewdirs:
        cols.append(viewdirs.reshape(-1, 3).float())
    return torch.cat(cols, -1)


# --------------------------------------------------------------------------------------------------------------
# a3  coarse sample placement                                             networks/renderer.py:157-180
# --------------------------------------------------------------------------------------------------------------
def place_samples(near, far, n_samples, lindisp=False, t_rand=None):
    """near/far [R,1]; t_rand [R,Nc] in [0,1) or None (perturb == 0)."""
    t = torch.linspace(0., 1., steps=n_samples)
    if not lindisp:
        z = near * (1. - t) + far * t
    else:
        z = 1. / (1. / near * (1. - t) + 1. / far * t)
    z = z.expand(near.shape[0], n_samples)
    if t_rand is not None:
        mids = .5 * (z[..., 1:] + z[..., :-1])
        upper = torch.cat([mids, z[..., -1:]], -1)
        lower = torch.cat([z[..., :1], mids], -1)
        z = lower + (upper - lower) * t_rand
    return z


# --------------------------------------------------------------------------------------------------------------
# a4  VM-decomposed feature lookup                       networks/pdrf/voxnerf.py:203-208 (sample), 132-151
# --------------------------------------------------------------------------------------------------------------
def vm_grid_size(aabb_min, aabb_max, n_voxels):
    """voxnerf.py:87-92 -> [Gx, Gy, Gz]."""
    amin, amax = torch.as_tensor(aabb_min, dtype=torch.float32), torch.as_tensor(aabb_max, dtype=torch.float32)
    voxel = ((amax - amin).prod() / n_voxels).pow(1 / 3)
    return ((amax - amin) / voxel).long().tolist()


def vm_products(P, prefix, pts, aabb_min, aabb_max):
    """compute_appfeature up to (plane (.) line), voxnerf.py:132-149, with F.grid_sample exactly as the reference
    issues it.  pts [..., 3] -> [n, 96]."""
    amin, amax = torch.as_tensor(aabb_min, dtype=torch.float32), torch.as_tensor(aabb_max, dtype=torch.float32)
    inv = 2.0 / (amax - amin)
    xyz = (pts.reshape(-1, 3) - amin) * inv - 1
    feats_p, feats_l = [], []
    for i in range(3):
        cp = xyz[..., list(MATMODE[i])].view(1, -1, 1, 2)
        cl = torch.stack((torch.zeros_like(xyz[..., 0]), xyz[..., VECMODE[i]]), -1).view(1, -1, 1, 2)
        feats_p.append(F.grid_sample(P[prefix + f"app_plane.{i}"], cp, align_corners=True).view(-1, xyz.shape[0]))
        feats_l.append(F.grid_sample(P[prefix + f"app_line.{i}"], cl, align_corners=True).view(-1, xyz.shape[0]))
    return (torch.cat(feats_p) * torch.cat(feats_l)).T


def vm_sample(P, prefix, pts, aabb_min, aabb_max):
    """`VoxelNeRFBase.sample` (voxnerf.py:203-208). pts [R,S,3] -> [R,S,app_dim]."""
    out = F.linear(vm_products(P, prefix, pts, aabb_min, aabb_max), P[prefix + "basis_mat.weight"])
    return out.reshape(pts.shape[0], pts.shape[1], -1)


def _bilinear_taps(img, x, y):
    """Independent restatement of grid_sample(bilinear, zeros padding, align_corners=True) on one [C,H,W] image:
    ix = ((x+1)/2)*(W-1); taps nw,ne,sw,se with weights (ix_se-ix)(iy_se-iy) ...; out-of-range taps contribute 0."""
    C, Hh, Ww = img.shape
    ix = ((x + 1) / 2) * (Ww - 1)
    iy = ((y + 1) / 2) * (Hh - 1)
    x0, y0 = torch.floor(ix), torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    out = torch.zeros(x.shape[0], C)
    for xi, yi, wgt in ((x0, y0, (x1 - ix) * (y1 - iy)), (x1, y0, (ix - x0) * (y1 - iy)),
                        (x0, y1, (x1 - ix) * (iy - y0)), (x1, y1, (ix - x0) * (iy - y0))):
        ok = (xi >= 0) & (xi <= Ww - 1) & (yi >= 0) & (yi <= Hh - 1)
        xc, yc = xi.clamp(0, Ww - 1).long(), yi.clamp(0, Hh - 1).long()
        out = out + img[:, yc, xc].T * (wgt * ok)[:, None]
    return out


def vm_sample_taps(P, prefix, pts, aabb_min, aabb_max):
    """Same as vm_sample but with the tap arithmetic written out (what the CUDA gather implements)."""
    amin, amax = torch.as_tensor(aabb_min, dtype=torch.float32), torch.as_tensor(aabb_max, dtype=torch.float32)
    xyz = (pts.reshape(-1, 3) - amin) * (2.0 / (amax - amin)) - 1
    cols = []
    for i in range(3):
        pl = _bilinear_taps(P[prefix + f"app_plane.{i}"][0], xyz[:, MATMODE[i][0]], xyz[:, MATMODE[i][1]])
        ln = _bilinear_taps(P[prefix + f"app_line.{i}"][0], torch.zeros_like(xyz[:, 0]), xyz[:, VECMODE[i]])
        cols.append(pl * ln)
    out = F.linear(torch.cat(cols, -1), P[prefix + "basis_mat.weight"])
    return out.reshape(pts.shape[0], pts.shape[1], -1)


# --------------------------------------------------------------------------------------------------------------
# a7  sigma -> alpha compositing                                      networks/pdrf/voxnerf.py:153-201
# --------------------------------------------------------------------------------------------------------------
def raw2outputs(raw, z_vals, rays_d, noise=None, is_train=True, rmnearplane=0, rgb_act="none"):
    """raw [R,S,4] = [sigma_raw, rgb]; noise [R,S-1] = N(0,1)*raw_noise_std already scaled, or None.
    Returns rgb_map, density, acc_map, weights, depth_map (the reference's order)."""
    dists = (z_vals[..., 1:] - z_vals[..., :-1]) * torch.norm(rays_d[..., None, :], dim=-1)
    rgb = torch.relu(raw[..., 1:]) if rgb_act == "relu" else raw[..., 1:]
    sig = raw[..., :-1, 0]
    if noise is not None:
        sig = sig + noise
    density = torch.relu(sig)
    if (not is_train) and rmnearplane > 0:
        density = (z_vals[:, 1:] > rmnearplane / 128).type_as(density) * density
    alpha = -torch.exp(-density * dists) + 1.
    alpha = torch.cat([alpha, torch.ones_like(alpha[:, :1])], -1)
    T = torch.cumprod(torch.cat([torch.ones(alpha.shape[0], 1), -alpha + (1. + 1e-10)], -1), -1)[:, :-1]
    weights = alpha * T
    rgb_map = torch.sum(weights[..., None] * rgb, -2)
    depth_map = torch.sum(weights * z_vals, -1)
    acc_map = torch.sum(weights, -1)
    return rgb_map, density, acc_map, weights, depth_map


# --------------------------------------------------------------------------------------------------------------
# a6  PDRF field forward (CRR coarse / FVR fine, non-composite branch)   networks/pdrf/voxnerf.py:210-259
# --------------------------------------------------------------------------------------------------------------
def _mlp(P, prefix, h, n_layers):
    for l in range(n_layers):
        h = F.linear(h, P[prefix + f"{l}.weight"], P.get(prefix + f"{l}.bias"))
        if l != n_layers - 1:
            h = F.relu(h)
    return h


def _count_layers(P, prefix):
    n = 0
    while prefix + f"{n}.weight" in P:
        n += 1
    return n


def field_forward(P, prefix, pts, viewdirs, fts, z_vals, rays_d, noise=None, is_train=True, rmnearplane=0,
                  rgb_act="none", L_pts=10, L_dir=4):
    """-> color [R,3], depth [R], acc [R], weights [R,S], feature_map [R,S,geo]."""
    R, S = pts.shape[:2]
    h = torch.cat([fts.reshape(R * S, -1), posenc(pts.reshape(-1, 3), L_pts)], -1)
    h = _mlp(P, prefix + "sigma_net.", h, _count_layers(P, prefix + "sigma_net."))
    feature_map = h[..., 1:].reshape(R, S, -1)
    dirs = posenc(viewdirs[:, None].expand(R, S, 3).reshape(-1, 3), L_dir)
    sigma = h[..., :1].reshape(R, S, 1)
    c = _mlp(P, prefix + "color_net.", torch.cat([h[..., 1:], dirs], -1), _count_layers(P, prefix + "color_net."))
    color = torch.sigmoid(c).reshape(R, S, 3)
    rgb_map, _, acc_map, weights, depth_map = raw2outputs(
        torch.cat([sigma, color], -1), z_vals, rays_d, noise, is_train, rmnearplane, rgb_act)
    return rgb_map, depth_map, acc_map, weights, feature_map


# --------------------------------------------------------------------------------------------------------------
# a8  hierarchical sampling                                                         utils/rays.py:149-193
# --------------------------------------------------------------------------------------------------------------
def sample_pdf(bins, weights, n_samples, u=None, norm="fp64"):
    """bins [R,B] (z mids), weights [R,B-1]; u [R,n] or None (det: linspace(0,1,n)).
    Returns (z_samples [R,n] fp32, inds [R,n] int64).  Specified arithmetic, see module docstring.
    norm="torch_sum" uses the reference's own expression torch.sum (rays.py:152), whose fp32 summation order depends
    on the CPU's vector width; it exists only so make_golden.py can show that the normaliser is the sole difference."""
    w = (weights + 1e-5).to(torch.float32)
    if norm == "torch_sum":
        norm = torch.sum(w, -1, keepdim=True)
    else:
        norm = w.double().cumsum(-1)[..., -1:].float()      # sequential fp64 accumulation, rounded once
    pdf = w / norm
    cdf = pdf.double().cumsum(-1).float()                    # == torch CPU cumsum on fp32 (accumulates in double)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    if u is None:
        u = torch.linspace(0., 1., steps=n_samples).expand(list(cdf.shape[:-1]) + [n_samples])
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = torch.clamp(inds - 1, min=0)
    above = torch.clamp(inds, max=cdf.shape[-1] - 1)
    cdf_b, cdf_a = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_b, bin_a = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_a - cdf_b
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    t = (u - cdf_b) / denom
    return bin_b + t * (bin_a - bin_b), inds


def merge_samples(z_coarse, z_samples):
    """renderer.py:205 -> z_vals [R,Nc+Ni], order [R,Nc+Ni] int64 (stable)."""
    return torch.sort(torch.cat([z_coarse, z_samples], -1), dim=-1, stable=True)


# --------------------------------------------------------------------------------------------------------------
# render_rays (mode = c2f)                                                networks/renderer.py:129-264
# --------------------------------------------------------------------------------------------------------------
def render_rays(P, cfg, ray_batch, n_samples, n_importance=0, perturb=0., lindisp=False, is_train=True,
                rand=None, want_feature=False):
    """cfg: dict(aabb_min, aabb_max, rmnearplane).  rand: optional dict(t_rand [R,Nc], u [R,Ni], noise0 [R,Nc-1],
    noise1 [R,Nc+Ni-1]) of injected random tensors (noise already multiplied by raw_noise_std)."""
    rand = rand or {}
    o, d = ray_batch[:, 0:3], ray_batch[:, 3:6]
    viewdirs = ray_batch[:, -3:]
    near, far = ray_batch[:, 6:7], ray_batch[:, 7:8]
    z = place_samples(near, far, n_samples, lindisp, rand.get("t_rand") if perturb > 0 else None)
    pts = o[:, None, :] + d[:, None, :] * z[..., None]
    amin, amax, rm = cfg["aabb_min"], cfg["aabb_max"], cfg.get("rmnearplane", 0)
    ftc = vm_sample(P, "mlp_coarse.", pts, amin, amax)
    rgb0, depth0, acc0, w0, feat = field_forward(P, "mlp_coarse.", pts, viewdirs, ftc, z, d, rand.get("noise0"),
                                                 is_train, rm, rgb_act="relu")
    ret = {"rgb_map": rgb0, "depth_map": depth0, "acc_map": acc0, "z_vals": z, "weights": w0}
    if n_importance > 0:
        ftf = vm_sample(P, "mlp_fine.", pts, amin, amax)
        ft0 = torch.cat([ftc, ftf], -1)
        z_mid = .5 * (z[..., 1:] + z[..., :-1])
        z_s, inds = sample_pdf(z_mid, w0[..., 1:-1], n_importance, rand.get("u") if perturb > 0 else None,
                               norm=cfg.get("pdf_norm", "fp64"))
        z_s = z_s.detach()                                                           # renderer.py:203
        z_all, order = merge_samples(z, z_s)
        pts1 = o[:, None, :] + d[:, None, :] * z_s[..., None]
        ft1 = torch.cat([vm_sample(P, "mlp_coarse.", pts1, amin, amax), vm_sample(P, "mlp_fine.", pts1, amin, amax)], -1)
        rows = torch.arange(pts1.shape[0])[:, None]
        pts_all = torch.cat([pts, pts1], 1)[rows, order]
        ft_all = torch.cat([ft0, ft1], 1)[rows, order]
        rgb, depth, acc, w, feat = field_forward(P, "mlp_fine.", pts_all, viewdirs, ft_all, z_all, d,
                                                 rand.get("noise1"), is_train, rm, rgb_act="none")
        ret = {"rgb_map": rgb, "depth_map": depth, "acc_map": acc, "z_vals": z_all, "weights": w,
               "rgb0": rgb0, "depth0": depth0, "acc0": acc0, "z_vals0": z, "weights0": w0,
               "z_std": torch.std(z_s, dim=-1, unbiased=False), "inds": inds, "order": order, "z_samples": z_s}
    if want_feature:
        ret["depth_feature"] = feat
    return ret


# --------------------------------------------------------------------------------------------------------------
# a10  vanilla NeRF field (mode = nerf), run_network                      networks/nerf.py:46-72, 131-162, 74-129
# --------------------------------------------------------------------------------------------------------------
def nerf_mlpforward(P, prefix, pts, viewdirs, L_pts=10, L_dir=4, skips=(4,), before_linear=True):
    """-> raw [R,S,4] = [rgb, sigma], feature [R,S,W]."""
    R, S = pts.shape[:2]
    x = posenc(pts.reshape(-1, 3), L_pts)
    dirs = posenc(viewdirs[:, None].expand(R, S, 3).reshape(-1, 3), L_dir)
    h = x
    D = _count_layers(P, prefix + "pts_linears.")
    for i in range(D):
        h = F.relu(F.linear(h, P[prefix + f"pts_linears.{i}.weight"], P[prefix + f"pts_linears.{i}.bias"]))
        if i in skips:
            h = torch.cat([x, h], -1)
    feat_before = h
    alpha = F.linear(h, P[prefix + "alpha_linear.weight"], P[prefix + "alpha_linear.bias"])
    feature = F.linear(h, P[prefix + "feature_linear.weight"], P[prefix + "feature_linear.bias"])
    h = F.relu(F.linear(torch.cat([feature, dirs], -1), P[prefix + "views_linears.0.weight"],
                        P[prefix + "views_linears.0.bias"]))
    rgb = F.linear(h, P[prefix + "rgb_linear.weight"], P.get(prefix + "rgb_linear.bias"))
    raw = torch.cat([rgb, alpha], -1).reshape(R, S, 4)
    f = feat_before if before_linear else feature
    return raw, f.reshape(R, S, -1)


def nerf_raw2outputs(raw, z_vals, rays_d, noise=None, white_bkgd=False, is_train=True, rmnearplane=0):
    """nerf.py:74-129: sigma is channel 3, rgb = sigmoid(channels 0..2)."""
    raw_v = torch.cat([raw[..., 3:4], torch.sigmoid(raw[..., :3])], -1)
    rgb_map, density, acc_map, weights, depth_map = raw2outputs(raw_v, z_vals, rays_d, noise, is_train, rmnearplane)
    if white_bkgd:
        rgb_map = rgb_map + (1. - acc_map[..., None])
    return rgb_map, density, acc_map, weights, depth_map


# --------------------------------------------------------------------------------------------------------------
# a11  adaptive weight proposal                    networks/dpnerf/awp.py:79-117, 49-77; networks/dpnerf/mam.py
# --------------------------------------------------------------------------------------------------------------
def awp_feature_integration(feat, z_vals, rays_d):
    """awp.py:49-77 literally: the cumprod runs over the LAST (channel) axis of a tensor concatenated along -2."""
    dists = (z_vals[..., 1:] - z_vals[..., :-1]) * torch.norm(rays_d[..., None, :], dim=-1)
    alpha = -torch.exp(-feat[..., :-1, :] * dists[..., None]) + 1
    alpha = torch.cat([alpha, torch.zeros_like(alpha[:, 0:1])], dim=-2)
    T = torch.cumprod(torch.cat([torch.ones((alpha.shape[0], 1, alpha.shape[-1])), -alpha + (1. + 1e-10)], -2), -1)
    weights = alpha * T[:, :-1, :]
    return torch.sum(weights * feat, dim=-2)


def awp_forward(P, depth_feature, z_vals, rays_d, view_feature, E, prefix="awpnet.", bn_eps=1e-5):
    """depth_feature [N*E,S,F], z_vals [N*E,S], rays_d [N*E,3], view_feature [N,32] -> ccw [N,E]
    (train mode: BatchNorm1d uses batch statistics, mam.py:24-27)."""
    NE, S, _ = depth_feature.shape
    N = NE // E
    vd = rays_d.reshape(N, E, -1)[:, 0, :]
    vd = vd / torch.norm(vd, dim=-1, keepdim=True)
    view = torch.cat([view_feature, posenc(vd, 2)], -1)
    h = depth_feature
    for i in range(_count_layers(P, prefix + "sample_feature_embed_layer.")):
        h = F.relu(F.linear(h, P[prefix + f"sample_feature_embed_layer.{i}.weight"],
                            P[prefix + f"sample_feature_embed_layer.{i}.bias"]))
    h_local = h
    g = awp_feature_integration(h, z_vals, rays_d).reshape(N, E, -1)
    g = torch.cat([g, view[:, None].repeat(1, E, 1)], -1)
    for i in range(_count_layers(P, prefix + "motion_feature_embed_layer.")):
        g = F.relu(F.linear(g, P[prefix + f"motion_feature_embed_layer.{i}.weight"],
                            P[prefix + f"motion_feature_embed_layer.{i}.bias"]))
    # MotionAggregationModule.forward (mam.py:67-84) + CorrelationModule.forward (mam.py:31-53)
    xl = F.linear(h_local.reshape(N, E, S, -1), P[prefix + "MAM.linear.weight"], P[prefix + "MAM.linear.bias"])
    curves = xl.permute(0, 3, 1, 2)                      # B C N L
    x = g.permute(0, 2, 1)                               # B C N
    c = prefix + "MAM.Corr."
    att = F.conv2d(curves, P[c + "line_conv_att.weight"])
    inter = torch.sum(curves * F.softmax(att, dim=-1), dim=-1)
    intra = torch.sum(curves * F.softmax(att, dim=-2), dim=-2)
    inter = F.conv1d(inter, P[c + "conva.weight"])
    intra = F.conv1d(intra, P[c + "convb.weight"])
    xlog = F.conv1d(x, P[c + "convc.weight"]).transpose(1, 2)
    x_inter = F.softmax(torch.bmm(xlog, inter), dim=-1)
    x_intra = F.softmax(torch.bmm(xlog, intra), dim=-1)
    inter = F.conv1d(inter, P[c + "convn.weight"]).transpose(1, 2)
    intra = F.conv1d(intra, P[c + "convl.weight"]).transpose(1, 2)
    cf = torch.cat((torch.bmm(x_inter, inter), torch.bmm(x_intra, intra)), dim=-1).transpose(1, 2)
    y = F.conv1d(cf, P[c + "convd.0.weight"])
    mean = y.mean(dim=(0, 2), keepdim=True)
    var = y.var(dim=(0, 2), unbiased=False, keepdim=True)
    y = (y - mean) / torch.sqrt(var + bn_eps) * P[c + "convd.1.weight"][None, :, None] + P[c + "convd.1.bias"][None, :, None]
    res = F.leaky_relu(x + y, negative_slope=0.2).permute(0, 2, 1)  # B N C
    pooled = res.mean(1)                                  # adaptive_avg_pool1d over exposures, awp.py:112
    w = torch.sigmoid(F.linear(pooled, P[prefix + "w_linear.weight"], P[prefix + "w_linear.bias"]))
    return w / w.sum(-1, keepdim=True)


# --------------------------------------------------------------------------------------------------------------
# NeRFAll.forward, training branch, RBK (+AWP)                            networks/renderer.py:277-378
# --------------------------------------------------------------------------------------------------------------
def forward_train(P, cfg, H, W, focal, rays, images_idx, E, n_samples, n_importance, perturb=0., rand=None,
                  use_awp=True, near=0., far=1.):
    new_rays, weight1, emb = rbk_forward(P, rays, images_idx, E - 1)
    rb = build_ray_batch(H, W, focal, new_rays.reshape(-1, 3, 2), near, far)
    ret = render_rays(P, cfg, rb, n_samples, n_importance, perturb, rand=rand, want_feature=use_awp)
    out = {"weight1": weight1, "new_rays": new_rays, "ray_batch": rb, "render": ret}
    out["rgb"] = rbk_weighted_sum(ret["rgb_map"], weight1)
    out["depth"] = rbk_weighted_sum(ret["depth_map"], weight1)
    out["acc"] = rbk_weighted_sum(ret["acc_map"], weight1)
    if n_importance > 0:
        out["rgb1"] = rbk_weighted_sum(ret["rgb0"], weight1)
    N = rays.shape[0]
    out["stage1_rgb_pts0"] = ret["rgb_map"].reshape(N, E, 3)[:, 0]
    if use_awp:
        ccw = awp_forward(P, ret["depth_feature"], ret["z_vals"], rb[:, 3:6], emb, E)
        ccw = ccw + ccw * 0.05                                         # renderer.py:316-317
        ccw = ccw / torch.sum(ccw, -1, keepdim=True)
        out["ccw_fine"] = ccw
        out["rgb_awp"] = rbk_weighted_sum(ret["rgb_map"], ccw)
    return out


# --------------------------------------------------------------------------------------------------------------
# f3  Deformable sparse kernel (DSK)                                          networks/pdrf/blurmodel.py:9-224
# --------------------------------------------------------------------------------------------------------------
def dsk_forward(P, cfg, H, W, K, rays_x, rays_y, images_idx, poses, noise=None, prefix="kernelsnet."):
    """BlurModel.forward with kernel_type = DSK (blurmodel.py:109-224).  cfg: num_pt, kernel_hwindow, in_embed, spatial_embed,
    num_hidden, short_cut, isglobal, optim_trans, optim_sv_trans.  rays_x / rays_y [N,1] pixel coordinates, poses [N,3,4],
    K = 3x3 intrinsics; noise [N,P,2] = randn_like(pt_pos) * random_hwindow (blurmodel.py:125-127) or None.
    Returns new_rays [N,P,3,2], weight [N,P], align (scalar)."""
    npt, hw = int(cfg["num_pt"]), float(cfg["kernel_hwindow"])
    idx = images_idx.reshape(-1).long()
    N = idx.shape[0]
    emb = P[prefix + "img_embed.img_embed"][idx]                                       # blurmodel.py:119
    pt_pos = P[prefix + "pattern_pos"]
    pt_pos = pt_pos.expand(N, -1, -1) if cfg.get("isglobal") else pt_pos[idx]        # blurmodel.py:122-124
    pt_pos = torch.tanh(pt_pos) * hw
    if noise is not None:
        pt_pos = pt_pos + noise
    input_pos = pt_pos
    if int(cfg.get("in_embed", 0)) > 0:
        pt_pos = posenc(pt_pos * (math.pi / hw), int(cfg["in_embed"]))              # blurmodel.py:130-132
    x = torch.cat([pt_pos, emb[:, None].expand(N, npt, emb.shape[-1])], -1)
    if int(cfg.get("spatial_embed", 0)) > 0:                                           # blurmodel.py:149-155
        sp = torch.cat([rays_x / (W / 2 / math.pi) - math.pi, rays_y / (H / 2 / math.pi) - math.pi], -1)
        sp = posenc(sp, int(cfg["spatial_embed"]))
        x = torch.cat([x, sp[:, None].expand(N, npt, sp.shape[-1])], -1)
    h = x
    for l in range(int(cfg["num_hidden"])):                                          # linears: Linear + ReLU, num_hidden times
        h = torch.relu(h @ P[prefix + f"linears.{2 * l}.weight"].t() + P[prefix + f"linears.{2 * l}.bias"])
    if cfg.get("short_cut"):
        h = torch.cat([x, h], -1)
    h = torch.relu(h @ P[prefix + "linears1.0.weight"].t() + P[prefix + "linears1.0.bias"])
    out = h @ P[prefix + "linears1.2.weight"].t() + P[prefix + "linears1.2.bias"]
    if cfg.get("optim_sv_trans"):
        delta_trans, delta_pos, wl = out[..., 0:2], out[..., 2:4], out[..., 4]
    else:
        delta_pos, wl = out[..., 0:2], out[..., 2]
        delta_trans = None
    if cfg.get("optim_trans"):                                                        # blurmodel.py:174-176
        pt = P[prefix + "pattern_trans"]
        delta_trans = pt.expand(N, -1, -1) if cfg.get("isglobal") else pt[idx]
    if delta_trans is None:
        delta_trans = torch.zeros_like(delta_pos)
    delta_trans = delta_trans * 0.01
    new_xy = delta_pos + input_pos
    align = new_xy[:, 0, :].abs().mean() + delta_trans[:, 0, :].abs().mean() * 10    # blurmodel.py:190-191
    weight = torch.softmax(wl, -1)
    rx = (rays_x - K[0, 2] + new_xy[..., 0]) / K[0, 0]                                 # blurmodel.py:199-200
    ry = -(rays_y - K[1, 2] + new_xy[..., 1]) / K[1, 1]
    dirs = torch.stack([rx - delta_trans[..., 0], ry - delta_trans[..., 1], -torch.ones_like(rx)], -1)
    rays_d = torch.sum(dirs[..., None, :] * poses[:, None, :3, :3], -1)
    tr = torch.stack([delta_trans[..., 0], delta_trans[..., 1], torch.zeros_like(rx), torch.ones_like(rx)], -1)
    rays_o = torch.sum(tr[..., None, :] * poses[:, None], -1)
    return torch.stack([rays_o, rays_d], -1), weight, align


def forward_train_dsk(P, cfg, dsk_cfg, H, W, K, rays_x, rays_y, images_idx, poses, n_samples, n_importance, noise=None, use_awp=False):
    """NeRFAll.forward, kernel_type = DSK (renderer.py:301-378): DSK rays -> render -> per-point weighted sums (+ the AWP branch)."""
    new_rays, weight, align = dsk_forward(P, dsk_cfg, H, W, K, rays_x, rays_y, images_idx, poses, noise)
    N, npt = weight.shape
    rb = build_ray_batch(H, W, float(K[0, 0]), new_rays.reshape(-1, 3, 2))
    ret = render_rays(P, cfg, rb, n_samples, n_importance, want_feature=use_awp)
    out = {"new_rays": new_rays, "weight": weight, "align": align, "render": ret,
           "rgb": torch.sum(ret["rgb_map"].reshape(N, npt, 3) * weight[..., None], 1)}
    if n_importance > 0:
        out["rgb1"] = torch.sum(ret["rgb0"].reshape(N, npt, 3) * weight[..., None], 1)
    if use_awp:                                                            # renderer.py:310-336 with a non-RBK kernel
        emb = P["kernelsnet.img_embed.img_embed"][images_idx.reshape(-1).long()]
        ccw = awp_forward(P, ret["depth_feature"], ret["z_vals"], rb[:, 3:6], emb, npt)
        ccw = ccw + ccw * 0.05
        ccw = ccw / torch.sum(ccw, -1, keepdim=True)
        out["ccw_fine"] = ccw
        out["rgb_awp"] = torch.sum(ret["rgb_map"].reshape(N, npt, 3) * ccw[..., None], 1)
    return out


# --------------------------------------------------------------------------------------------------------------
# a17  TV regulariser                                networks/pdrf/voxnerf.py:126-130, 306-324; renderer.py:361-365
# --------------------------------------------------------------------------------------------------------------
def tv_reg(x):
    count_h = x.shape[1] * (x.shape[2] - 1) * x.shape[3]
    count_w = max(x.shape[1] * x.shape[2] * (x.shape[3] - 1), 1)
    h_tv = torch.pow(x[:, :, 1:, :] - x[:, :, :-1, :], 2).sum()
    w_tv = torch.pow(x[:, :, :, 1:] - x[:, :, :, :-1], 2).sum()
    return 2 * (h_tv / count_h + w_tv / count_w) / x.shape[0]


def tv_loss_app(P, prefix):
    total = 0
    for i in range(3):
        total = total + tv_reg(P[prefix + f"app_plane.{i}"]) * 1e-2 + tv_reg(P[prefix + f"app_line.{i}"]) * 1e-3
    return total


# --------------------------------------------------------------------------------------------------------------
# a13  camera response / tone mapping                                  networks/tonemapping.py:59-93, 111-139
# --------------------------------------------------------------------------------------------------------------
def crf_apply(P, prefix, x, map_type, gamma=2.2, x_feat=None, extra_features=0, skip_learn=False):
    if map_type == "none":
        return x
    if "gamma" in map_type:
        x = x ** (1. / gamma)
    if skip_learn or map_type != "learn":
        return x
    shape = x.shape
    x_in = x.reshape(-1, 1)
    if x_feat is not None and extra_features > 0:
        f = x_feat.to(x_in.dtype)
        if f.ndim != 3:
            f = f[:, None].repeat(1, 3, 1)
        inp = torch.cat([x_in, f.reshape(-1, extra_features)], -1)
    else:
        inp = x_in
        if extra_features > 0:
            inp = torch.cat([x_in, torch.zeros(x_in.shape[0], extra_features)], -1)
    h = inp
    for i in (0, 2, 4, 6):
        h = F.linear(h, P[prefix + f"linear.{i}.weight"], P[prefix + f"linear.{i}.bias"])
        if i != 6:
            h = F.relu(h)
    return torch.sigmoid(h * 0.1 + x_in).reshape(shape)


def encode_rgb(P, x, map_type_rgb="gamma", gamma=2.2):
    return crf_apply(P, "tonemapping_rgb.", x, map_type_rgb, gamma)


def encode_luma(P, x, map_type_event="learn", gamma=2.2, ev_extra_feat=None, extra_features=2, skip_learn=False,
                tonemap_only=False, luma_standard="rec601"):
    x = crf_apply(P, "tonemapping_event.", x, map_type_event, gamma, ev_extra_feat, extra_features, skip_learn)
    if not tonemap_only:
        if luma_standard == "rec601":
            x = 0.299 * x[..., [0]] + 0.587 * x[..., [1]] + 0.114 * x[..., [2]]  # tonemapping.py:128-129
        elif luma_standard == "rec709":
            x = 0.2126 * x[..., [0]] + 0.7152 * x[..., [1]] + 0.0722 * x[..., [2]]  # tonemapping.py:130-131
        elif luma_standard == "avg":
            x = x.mean(axis=-1, keepdims=True)  # tonemapping.py:132-133
        else:
            raise ValueError(f"Unknown luma_standard {luma_standard}")
    return x


# --------------------------------------------------------------------------------------------------------------
# a14 / a15  event generation-model loss and photometric loss       utils/events.py:260-284; utils/metrics.py:7-8
# --------------------------------------------------------------------------------------------------------------
def egm_loss(luma_start, luma_end, bii, color_mask=None, color_weight=None, log_eps=1e-5):
    pred = (torch.log(luma_end + log_eps) - torch.log(luma_start + log_eps)).squeeze(-1)
    if color_mask is not None:
        pred = pred[color_mask]
        cw = None
        if color_weight is not None:
            cw = torch.as_tensor(color_weight, dtype=torch.float32)[torch.where(color_mask)[1]]
    else:
        cw = None
    if cw is None:
        cw = torch.ones(pred.shape[0])
    return (((pred - bii) ** 2) * cw).sum() / cw.sum()


def img2mse(x, y):
    return torch.mean((x - y) ** 2)


def mse2psnr(x):
    return -10. * torch.log(x) / math.log(10.)


# --------------------------------------------------------------------------------------------------------------
# a16  event double integral prior (numpy, like the reference)                         utils/edi.py:7-95
# --------------------------------------------------------------------------------------------------------------
def edi_splat(x, y, w, h):
    """interpolate_subpixel with unit values: bilinear splat on the floor/ceil taps (edi.py:7-41).  A tap is kept
    when ref < w,h and it is not the duplicate ceil tap of an integer coordinate; negative refs wrap like np.add.at."""
    img = np.zeros((h, w), dtype=np.float32)
    if x.size == 0:
        return img
    for fx in (np.floor, np.ceil):
        for fy in (np.floor, np.ceil):
            xr, yr = fx(x), fy(y)
            ok = ((xr != x) | (fx is np.floor)) & ((yr != y) | (fy is np.floor)) & (xr < w) & (yr < h)
            xr, yr = xr[ok], yr[ok]
            if xr.shape[0] > 0:
                val = (np.ones_like(xr, dtype=np.float32) * np.maximum(0, 1 - np.abs(xr - x[ok]))
                       * np.maximum(0, 1 - np.abs(yr - y[ok]))).astype(np.float32)
                np.add.at(img, (yr.astype(np.int64), xr.astype(np.int64)), val)
    return img


def edi_bii(x, y, p, w, h, c_pos, c_neg):
    """brightness_increment_image, interpolate=True, color_events=False (edi.py:44-70)."""
    pos = p > 0
    return edi_splat(x[pos], y[pos], w, h) * c_pos - edi_splat(x[~pos], y[~pos], w, h) * c_neg


def edi_inner(bii):
    """inner_double_integral (edi.py:73-88): bii [2N,...] -> [2N+1,...] signed partial sums around mid-exposure."""
    N = bii.shape[0] // 2
    imgs = [-bii[i:N].sum(axis=0) for i in range(N)]
    imgs.append(np.zeros_like(imgs[0]))
    imgs += [bii[N:N + 1 + i].sum(axis=0) for i in range(N)]
    return np.stack(imgs, 0)


def edi_deblur(blurry, bii):
    """deblur_double_integral (edi.py:91-95)."""
    N = bii.shape[0] // 2
    return (2 * N + 1) * blurry / np.exp(edi_inner(bii)).sum(axis=0)


def edi_prior_image(ev_x, ev_y, ev_t, ev_p, blurry, t_start, t_end, w, h, c_pos, c_neg, steps=9):
    """compute_edi_prior for one image (data/loader_events.py:99-131): `steps` time stamps -> steps-1 sub-intervals,
    each taking events with searchsorted(left) of t_j .. searchsorted(right) of t_{j+1} (inclusive both ends)."""
    ts = np.linspace(t_start, t_end, steps)
    i0 = np.searchsorted(ev_t, ts[:-1], side="left")
    i1 = np.searchsorted(ev_t, ts[1:], side="right")
    bii = np.stack([edi_bii(ev_x[a:b], ev_y[a:b], ev_p[a:b], w, h, c_pos, c_neg) for a, b in zip(i0, i1)], 0)
    bii = np.repeat(bii[..., None], blurry.shape[-1], axis=-1) if blurry.ndim == 3 else bii
    return edi_deblur(blurry, bii)


# --------------------------------------------------------------------------------------------------------------
# (f).2  ray / event batch generation     data/loader.py:325-356, utils/rays.py:25-36, data/loader_events.py:133-148,
#        259-304, utils/events.py:221-257, utils/data.py:34-61, 167-183
# --------------------------------------------------------------------------------------------------------------
def rays_from_pixels(coords, K, c2ws, add_halfpix=True):
    """get_rays_pix (utils/rays.py:25-36): coords [n,2] = (x, y), c2ws [n,3,4] (or [3,4]) -> rays [n,3,2]."""
    half = 0.5 if add_halfpix else 0.0
    x, y = coords[:, 0], coords[:, 1]
    dirs = torch.stack([(x + (half - K[0][2])) / K[0][0], -(y + (half - K[1][2])) / K[1][1], -torch.ones_like(x)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2ws[..., :3, :3], -1)
    rays_o = c2ws[..., :3, -1].expand(rays_d.shape)
    return torch.stack([rays_o, rays_d], -1)


def make_rgb_batch(ray_ids, images, poses, K):
    """LLFFDataset.__getitem__ (data/loader.py:325-356): ray ids over [n_img, H, W] -> the training batch dict."""
    n_img, Hh, Ww, _ = images.shape
    img_id = ray_ids // (Hh * Ww)
    ray_y = (ray_ids % (Hh * Ww)) // Ww
    ray_x = ray_ids % Ww
    c2w = poses[img_id]
    rays = rays_from_pixels(torch.stack([ray_x, ray_y], -1).to(poses.dtype), K, c2w)
    return {"rays": rays, "rays_x": (ray_x + 0.5).reshape(-1, 1), "rays_y": (ray_y + 0.5).reshape(-1, 1),
            "images_idx": img_id.reshape(-1, 1), "rgbsf": images[img_id, ray_y, ray_x].reshape(-1, 3), "poses": c2w.reshape(-1, 3, 4)}


def gather_successor(query_idx, query_hops, successor_map, polarities):
    """utils/events.py:221-257: follow the per-pixel successor map `hops + 1` times, accumulating the polarities of the
    visited events; an out-of-range successor invalidates the query (-1, 0, 0)."""
    n_ev = successor_map.shape[0]
    out_idx = query_idx.clone()
    pos = torch.zeros_like(query_idx, dtype=polarities.dtype)
    neg = torch.zeros_like(query_idx, dtype=polarities.dtype)
    invalid = torch.zeros(query_idx.shape[0], dtype=torch.bool)
    for i in range(query_idx.shape[0]):
        cur = int(query_idx[i])
        for _ in range(int(query_hops[i]) + 1):
            nxt = int(successor_map[cur])
            if nxt < 0 or nxt >= n_ev:
                invalid[i] = True
                break
            p = int(polarities[nxt])
            if p > 0:
                pos[i] += p
            elif p < 0:
                neg[i] += p
            cur = nxt
        out_idx[i] = cur
    out_idx[invalid] = -1
    pos[invalid] = 0
    neg[invalid] = 0
    return out_idx, neg, pos


def pose_interpolator(times, rots, trans):
    """utils/data.py:34-61: SLERP of the rotations (scipy Slerp) + cubic-spline interpolation of the translations
    (scipy interp1d(kind='cubic')), queries clipped to the known range.  Returns f(t [n]) -> (R [n,3,3], T [n,3])."""
    from scipy.interpolate import interp1d
    from scipy.spatial.transform import Rotation, Slerp
    slerp = Slerp(times, Rotation.from_matrix(rots))
    cubic = interp1d(x=times, y=trans, axis=0, kind="cubic", bounds_error=True)

    def f(t):
        t = np.clip(t, times[0], times[-1])
        return slerp(t).as_matrix(), cubic(t)
    return f


def interpolate_event_poses(interp, t, bd_scale=1.0, recenter_c2w=None):
    """LLFFEventsDataset.interpolate_poses (data/loader_events.py:133-148) without spherify: interpolated [R | T] -> column
    reorder [c1, -c0, c2, T] -> T *= bd_scale -> inv(recenter_c2w) @ pose.  Returns [n,3,4] float32."""
    Rm, T = interp(np.asarray(t, dtype=np.float64))
    pose = np.concatenate([Rm, T[..., None]], -1)                                        # [n,3,4]
    pose = np.concatenate([pose[..., 1:2], -pose[..., 0:1], pose[..., 2:]], -1).astype(np.float32)
    pose[..., :3, 3] *= bd_scale
    if recenter_c2w is not None:
        bottom = np.tile(np.array([0, 0, 0, 1.0], dtype=np.float64).reshape(1, 1, 4), [pose.shape[0], 1, 1])
        full = np.concatenate([pose[:, :3, :4], bottom], -2)
        pose = (np.linalg.inv(recenter_c2w) @ full)[:, :3, :4]
    return pose.astype(np.float32)
