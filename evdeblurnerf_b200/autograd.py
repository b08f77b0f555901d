"""torch.autograd bridges: the reference trains by `loss.backward()` through its PyTorch graph (run_nerf.py:594); here each
fused forward entry point is paired with its hand-written backward (field_bwd.cu, ray_bwd.cu, loss_bwd.cu) in a
`torch.autograd.Function`, so outputs of the host mirrors participate in autograd exactly like the reference's and
parameter `.grad`s come out under the reference's own names / shapes (SURVEY.md 8(b) "Ownership / memory").

Forward arithmetic is identical to the no-grad path (same kernels); nothing per-sample is saved -- the render backward
recomputes activations chunk by chunk (see field_bwd.cu).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import AwpGrads, CrfGrads, RbkGrads, check, ptr, stream_ptr

_RBK_NAMES = ("view_embed_module.img_embed", "r_branch.0.weight", "r_branch.0.bias", "v_branch.0.weight", "v_branch.0.bias",
              "w_branch.0.weight", "w_branch.0.bias", "r_linear.weight", "r_linear.bias", "v_linear.weight", "v_linear.bias",
              "w_linear.weight", "w_linear.bias")
_RBK_FIELDS = ("img_embed", "r_branch_w", "r_branch_b", "v_branch_w", "v_branch_b", "w_branch_w", "w_branch_b", "r_linear_w",
               "r_linear_b", "v_linear_w", "v_linear_b", "w_linear_w", "w_linear_b")


def _c(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


class WeightedSumFn(torch.autograd.Function):
    """rbk_weighted_sum for one tensor (blurmodel.py:112-127): x [N*E, ...], ccw [N,E] -> [N, ...]."""

    @staticmethod
    def forward(ctx, x, ccw):
        N, E = ccw.shape
        xf, wf = _c(x), _c(ccw)
        Cn = max(1, xf.numel() // max(1, N * E))
        out = torch.empty((N,) + tuple(xf.shape[1:]), dtype=torch.float32, device=xf.device)
        check(_lib.load().edn_weighted_sum(ptr(xf), ptr(wf), ptr(out), N, E, Cn, stream_ptr()), "edn_weighted_sum")
        ctx.save_for_backward(xf, wf)
        ctx.dims = (N, E, Cn)
        return out

    @staticmethod
    def backward(ctx, d_out):
        xf, wf = ctx.saved_tensors
        N, E, Cn = ctx.dims
        d_x = torch.empty_like(xf) if ctx.needs_input_grad[0] else None
        d_w = torch.empty_like(wf) if ctx.needs_input_grad[1] else None
        if N > 0:
            check(_lib.load().edn_weighted_sum_bwd(ptr(xf), ptr(wf), ptr(_c(d_out)), N, E, Cn, ptr(d_x), ptr(d_w), stream_ptr()),
                  "edn_weighted_sum_bwd")
        return d_x, d_w


class Img2MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y):
        xf, yf = _c(x), _c(y)
        out = torch.empty((1,), dtype=torch.float32, device=xf.device)
        check(_lib.load().edn_img2mse(ptr(xf), ptr(yf), xf.numel(), ptr(out), stream_ptr()), "edn_img2mse")
        ctx.save_for_backward(xf, yf)
        return out[0]

    @staticmethod
    def backward(ctx, d_loss):
        xf, yf = ctx.saved_tensors
        d_x = torch.empty_like(xf)
        check(_lib.load().edn_img2mse_bwd(ptr(xf), ptr(yf), xf.numel(), ptr(_c(d_loss).reshape(1)), ptr(d_x), stream_ptr()),
              "edn_img2mse_bwd")
        return d_x, None


class EgmLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ls, le, b, mask, cw, log_eps):
        M, Cn = ls.shape
        out = torch.empty((1,), dtype=torch.float32, device=ls.device)
        check(_lib.load().edn_egm_loss_fwd(ptr(ls), ptr(le), ptr(b), ptr(mask), ptr(cw), Cn, M, float(log_eps), ptr(out), stream_ptr()),
              "edn_egm_loss_fwd")
        ctx.save_for_backward(ls, le, b)
        ctx.aux = (mask, cw, float(log_eps))
        return out[0]

    @staticmethod
    def backward(ctx, d_loss):
        ls, le, b = ctx.saved_tensors
        mask, cw, eps = ctx.aux
        M, Cn = ls.shape
        d_ls, d_le = torch.empty_like(ls), torch.empty_like(le)
        check(_lib.load().edn_egm_loss_bwd(ptr(ls), ptr(le), ptr(b), ptr(mask), ptr(cw), Cn, M, eps, ptr(_c(d_loss).reshape(1)),
                                           ptr(d_ls), ptr(d_le), stream_ptr()), "edn_egm_loss_bwd")
        return d_ls, d_le, None, None, None, None


class CrfFn(torch.autograd.Function):
    """CRF.forward / encode_rgb / encode_luma (tonemapping.py:59-139).  `weights`: the 8 CRF tensors w0, b0, ..., w3, b3 (may
    be empty when the mapping has no learnt part)."""

    @staticmethod
    def forward(ctx, x, feat, per_channel, flags, cp, *weights):
        M = x.shape[0]
        luma = bool(flags & _lib.CRF_LUMA)
        out = torch.empty((M, 1 if luma else 3), dtype=torch.float32, device=x.device)
        check(_lib.load().edn_crf_fwd(C.byref(cp), ptr(x), ptr(feat), per_channel, flags, M, ptr(out), stream_ptr()), "edn_crf_fwd")
        ctx.save_for_backward(x, *([feat] if feat is not None else []))
        ctx.aux = (feat is not None, per_channel, flags, cp, [tuple(w.shape) for w in weights])
        return out

    @staticmethod
    def backward(ctx, d_out):
        has_feat, per_channel, flags, cp, shapes = ctx.aux
        x = ctx.saved_tensors[0]
        feat = ctx.saved_tensors[1] if has_feat else None
        M = x.shape[0]
        d_x = torch.empty_like(x)
        gw = [torch.zeros(s, dtype=torch.float32, device=x.device) for s in shapes]
        learn = bool(flags & _lib.CRF_LEARN) and not (flags & _lib.CRF_SKIP_LEARN)
        g = None
        if learn and len(gw) == 8:
            g = CrfGrads()
            for slot, t in zip(("w0", "b0", "w1", "b1", "w2", "b2", "w3", "b3"), gw):
                setattr(g, slot, t.data_ptr())
        check(_lib.load().edn_crf_bwd(C.byref(cp), ptr(x), ptr(feat), per_channel, flags, M, ptr(_c(d_out)), ptr(d_x),
                                      C.byref(g) if g is not None else None, stream_ptr()), "edn_crf_bwd")
        return (d_x, None, None, None, None) + tuple(gw)


class TvLossFn(torch.autograd.Function):
    """VoxelNeRFBase.TV_loss_app (voxnerf.py:126-130) over 3 planes + 3 lines in the reference layout."""

    @staticmethod
    def _arrays(planes, lines):
        pp = (C.c_void_p * 3)(*[t.data_ptr() for t in planes])
        lp = (C.c_void_p * 3)(*[t.data_ptr() for t in lines])
        ph = (C.c_int32 * 3)(*[t.shape[2] for t in planes])
        pw = (C.c_int32 * 3)(*[t.shape[3] for t in planes])
        ll = (C.c_int32 * 3)(*[t.shape[2] for t in lines])
        nc = (C.c_int32 * 3)(*[t.shape[1] for t in planes])
        return pp, lp, ph, pw, ll, nc

    @staticmethod
    def forward(ctx, *tensors):
        ts = [_c(t) for t in tensors]
        planes, lines = ts[:3], ts[3:]
        pp, lp, ph, pw, ll, nc = TvLossFn._arrays(planes, lines)
        dev = planes[0].device
        ws = torch.empty((12,), dtype=torch.float64, device=dev)
        out = torch.empty((1,), dtype=torch.float32, device=dev)
        check(_lib.load().edn_tv_loss_app(C.byref(pp), C.byref(lp), C.byref(ph), C.byref(pw), C.byref(ll), C.byref(nc), ptr(ws),
                                          ptr(out), stream_ptr()), "edn_tv_loss_app")
        ctx.save_for_backward(*ts)
        return out[0]

    @staticmethod
    def backward(ctx, d_loss):
        ts = list(ctx.saved_tensors)
        planes, lines = ts[:3], ts[3:]
        pp, lp, ph, pw, ll, nc = TvLossFn._arrays(planes, lines)
        grads = [torch.zeros_like(t) for t in ts]
        gp = (C.c_void_p * 3)(*[t.data_ptr() for t in grads[:3]])
        gl = (C.c_void_p * 3)(*[t.data_ptr() for t in grads[3:]])
        check(_lib.load().edn_tv_loss_app_bwd(C.byref(pp), C.byref(lp), C.byref(ph), C.byref(pw), C.byref(ll), C.byref(nc),
                                              ptr(_c(d_loss).reshape(1)), C.byref(gp), C.byref(gl), stream_ptr()),
              "edn_tv_loss_app_bwd")
        return tuple(grads)


class RenderSubRaysFn(torch.autograd.Function):
    """Blur-kernel warp + render() prologue + c2f render_rays of the N*E sub-rays (renderer.py:303-309 -> 129-264) as ONE
    autograd node.  Inputs: the parameter tensors (for graph connectivity; the kernels read the engine's packed copies).
    Outputs: rgb_map, depth_map, acc_map, rgb0, depth0, acc0 [N*E, ...] and the blur weights [N,E]; everything else the
    forward produced (z_vals, weights, img_embed ...) is returned through `owner.last_render` without gradient."""

    @staticmethod
    def forward(ctx, owner, kn, H, W, focal, rays, images_idx, near, far, ndc, kwargs, names, *params):
        eng = owner.engine
        kw = dict(kwargs)
        kw["retraw"] = True
        if kn is not None:
            k = kn.warp(H, W, focal, rays, images_idx, near, far, ndc, want_new_rays=False)
            rb, weight = k["ray_batch"], k["weight"]
        else:
            from .renderer import build_ray_batch
            rb, weight, k = build_ray_batch(H, W, focal, rays, near, far, ndc), None, {}
        extra_rays = kw.pop("extra_rays", None)
        if extra_rays is not None and extra_rays.shape[0] > 0:
            # rays rendered WITHOUT the blur kernel in the same launch (the event start / end rays of run_nerf.py:534-557): their
            # rows follow the warped sub-rays; no gradient flows to their ray batch
            from .renderer import build_ray_batch
            rb = torch.cat([rb, build_ray_batch(H, W, focal, extra_rays, near, far, ndc)], 0)
        R, dev = rb.shape[0], rb.device
        Nc, Ni = int(kw["N_samples"]), int(kw.get("N_importance", 0))
        # draw the density noise here so that backward sees the very same tensors (voxnerf.py:175)
        rand = dict(kw.pop("rand", None) or {})
        std = float(kw.get("raw_noise_std", 0.))
        if std > 0.:
            eng._calls += 1
            if "noise0" not in rand:
                rand["noise0"] = eng._random((R, Nc - 1), 1, normal=True, scale=std)
            if Ni > 0 and "noise1" not in rand:
                rand["noise1"] = eng._random((R, Nc + Ni - 1), 3, normal=True, scale=std)
        if Ni > 0 and owner.mode == "c2f":
            kw["want_indices"] = True            # the merged order lets the backward scatter every coarse position once
        out = owner.render_rays(rb, rand=rand, **kw)
        owner.last_render = dict(out, ray_batch=rb, img_embed=k.get("img_embed"), weight=weight)
        ctx.owner, ctx.kn, ctx.names = owner, kn, names
        ctx.geom = (H, W, focal, ndc)
        ctx.two_stage = Ni > 0
        ctx.white_bkgd = bool(kw.get("white_bkgd", False))
        ctx.saved = {"ray_batch": rb, "z_vals0": out["z_vals0"] if Ni > 0 else out["z_vals"], "z_vals": out["z_vals"] if Ni > 0 else None,
                     "noise0": rand.get("noise0"), "noise1": rand.get("noise1"), "order": out.get("order") if Ni > 0 else None}
        ctx.rays = _c(rays)
        ctx.rays_grad = kn is None and isinstance(rays, torch.Tensor) and rays.requires_grad      # rays from a learned kernel (DSK)
        ctx.idx = images_idx.reshape(-1).to(torch.int64).contiguous() if images_idx is not None else None
        zero3, zero1 = torch.zeros((R, 3), device=dev), torch.zeros((R,), device=dev)
        empty = torch.zeros((0,), device=dev)
        ctx.has_feat = "depth_feature" in out
        res = (out["rgb_map"], out["depth_map"], out["acc_map"], out.get("rgb0", zero3), out.get("depth0", zero1), out.get("acc0", zero1),
               weight if weight is not None else empty, out.get("depth_feature", empty), rb.clone(),
               k["img_embed"] if k.get("img_embed") is not None else empty)
        return res

    @staticmethod
    def backward(ctx, d_rgb, d_depth, d_acc, d_rgb0, d_depth0, d_acc0, d_weight, d_feat, d_rb_out, d_img_embed):
        owner, kn = ctx.owner, ctx.kn
        eng = owner.engine
        lib = _lib.load()
        d_out = {"rgb_map": d_rgb, "depth_map": d_depth, "acc_map": d_acc}
        if ctx.two_stage:
            d_out.update(rgb0=d_rgb0, depth0=d_depth0, acc0=d_acc0)
        if ctx.has_feat and d_feat is not None and d_feat.numel():
            d_out["depth_feature"] = d_feat
        named, d_rb = eng.backward(ctx.saved, d_out, chunk_rays=owner.backward_chunk_rays, white_bkgd=ctx.white_bkgd)
        if d_rb_out is not None and d_rb_out.numel():
            d_rb += d_rb_out                                   # direct uses of the ray batch (AWP reads rays_d)
        if d_img_embed is not None and not d_img_embed.numel():
            d_img_embed = None
        if kn is not None:
            H, W, focal, ndc = ctx.geom
            N = ctx.rays.shape[0]
            g = RbkGrads()
            for nm, field in zip(_RBK_NAMES, _RBK_FIELDS):
                t = torch.zeros_like(kn.tensors[nm])
                named[kn.prefix + nm] = t
                setattr(g, field, t.data_ptr())
            ws = torch.empty((int(lib.edn_rbk_bwd_workspace_floats(N, kn.num_motion)),), dtype=torch.float32, device=d_rb.device)
            check(lib.edn_rbk_warp_ndc_bwd(C.byref(kn.p), ptr(ctx.rays), ptr(ctx.idx), N, int(H), int(W), float(focal), 1 if ndc else 0,
                                           ptr(d_rb), ptr(_c(d_weight)) if d_weight is not None and d_weight.numel() else None,
                                           ptr(_c(d_img_embed)) if d_img_embed is not None else None,
                                           C.byref(g), ptr(ws), stream_ptr()), "edn_rbk_warp_ndc_bwd")
        d_rays = None
        if ctx.rays_grad:          # explicit rays that carry a graph: d ray_batch -> d rays (their rows come first)
            H, W, focal, ndc = ctx.geom
            R = ctx.rays.reshape(-1, 3, 2).shape[0]
            d_rays = torch.empty((R, 3, 2), dtype=torch.float32, device=d_rb.device)
            check(lib.edn_build_ray_batch_bwd(ptr(ctx.rays), R, int(H), int(W), float(focal), 1 if ndc else 0, ptr(_c(d_rb[:R])),
                                              ptr(d_rays), stream_ptr()), "edn_build_ray_batch_bwd")
            d_rays = d_rays.reshape(ctx.rays.shape)
        out = []
        for nm in ctx.names:
            gr = named.get(nm)
            out.append(gr)
        return (None,) * 5 + (d_rays,) + (None,) * 6 + tuple(out)


_AWP_FIELDS = (   # (struct field, index | None, state_dict name, transposed)
    [("sample_t", l, f"sample_feature_embed_layer.{l}.weight", True) for l in range(4)] +
    [("sample_b", l, f"sample_feature_embed_layer.{l}.bias", False) for l in range(4)] +
    [("motion_w", l, f"motion_feature_embed_layer.{l}.weight", False) for l in range(2)] +
    [("motion_b", l, f"motion_feature_embed_layer.{l}.bias", False) for l in range(2)] +
    [("mam_linear_t", None, "MAM.linear.weight", True), ("mam_linear_b", None, "MAM.linear.bias", False),
     ("line_conv_att", None, "MAM.Corr.line_conv_att.weight", False), ("conva", None, "MAM.Corr.conva.weight", False),
     ("convb", None, "MAM.Corr.convb.weight", False), ("convc", None, "MAM.Corr.convc.weight", False),
     ("convn", None, "MAM.Corr.convn.weight", False), ("convl", None, "MAM.Corr.convl.weight", False),
     ("convd_w", None, "MAM.Corr.convd.0.weight", False), ("bn_weight", None, "MAM.Corr.convd.1.weight", False),
     ("bn_bias", None, "MAM.Corr.convd.1.bias", False), ("w_linear_w", None, "w_linear.weight", False),
     ("w_linear_b", None, "w_linear.bias", False)])
AWP_PARAM_NAMES = tuple(f[2] for f in _AWP_FIELDS)


def awp_grad_buffers(shapes, device):
    """Zeroed gradient buffers in the layout edn_awp_bwd writes (transposed where the packed weights are) + their struct."""
    g, bufs = AwpGrads(), []
    for (field, idx, _, transposed), shp in zip(_AWP_FIELDS, shapes):
        t = torch.zeros(tuple(shp)[::-1] if transposed else tuple(shp), dtype=torch.float32, device=device)
        bufs.append(t)
        if idx is None:
            setattr(g, field, t.data_ptr())
        else:
            getattr(g, field)[idx] = t.data_ptr()
    return g, bufs


def awp_grads_to_reference(bufs):
    return [(b.t().contiguous() if f[3] else b) for f, b in zip(_AWP_FIELDS, bufs)]


class AwpFn(torch.autograd.Function):
    """AdaptiveWeightProposal.forward (awp.py:79-117) with its hand-written backward (awp_bwd.cu).
    Inputs: depth_feature [N*E,S,128], z_vals [N*E,S] (no gradient), rays_d [N*E,3], view_feature [N,32], then the AWP
    parameter tensors in AWP_PARAM_NAMES order (graph connectivity; the kernels read the module's packed copies)."""

    @staticmethod
    def forward(ctx, awp, depth_feature, z_vals, rays_d, view_feature, *params):
        df, z, rd, vf = _c(depth_feature), _c(z_vals), _c(rays_d), _c(view_feature)
        ctx.awp = awp
        ctx.save_for_backward(df, z, rd, vf)
        ctx.shapes = [tuple(t.shape) for t in params]
        # run the forward straight into a backward-sized workspace: in bf16 mode it keeps the layer activations, so the
        # backward does not recompute them
        NE, S, _ = df.shape
        ctx.ws = torch.empty((int(_lib.load().edn_awp_bwd_workspace_floats(NE // awp.E, awp.E, S)),), dtype=torch.float32, device=df.device)
        return awp.run(df, z, rd, vf, workspace=ctx.ws, keep_activations=True)

    @staticmethod
    def backward(ctx, d_ccw):
        awp = ctx.awp
        df, z, rd, vf = ctx.saved_tensors
        lib = _lib.load()
        NE, S, _ = df.shape
        E = awp.E
        N = NE // E
        dev = df.device
        g, bufs = awp_grad_buffers(ctx.shapes, dev)
        d_df = torch.empty_like(df)
        d_rd = torch.zeros_like(rd)
        d_vf = torch.empty_like(vf)
        ws, ctx.ws = ctx.ws, None
        world = awp._world()
        # phases 1 / 2: BatchNorm row count = the all-reduced one the forward left in the workspace (bn_rows_total = 0)
        phases = [awp.options(True)] if world == 1 else [awp.options(True, 1), awp.options(True, 2)]
        dcc = _c(d_ccw)
        for o in phases:      # the forward's workspace still holds the activations and the (all-reduced) batch sums
            check(lib.edn_awp_bwd(C.byref(awp.p), ptr(df), ptr(z), ptr(rd), 3, ptr(vf), N, E, S, awp.bn_eps, C.byref(o), 1, ptr(dcc),
                                  C.byref(g), ptr(d_df), ptr(d_rd), 3, ptr(d_vf), ptr(ws), stream_ptr()), "edn_awp_bwd")
            if o.phase == 1:
                awp._all_reduce_block(ws, int(lib.edn_awp_bwd_sums_offset_floats(N, E, S)), 64)
        return (None, d_df, None, d_rd, d_vf) + tuple(awp_grads_to_reference(bufs))
