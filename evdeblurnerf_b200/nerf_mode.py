"""mode = nerf: host mirror of the vanilla NeRF field (reference networks/nerf.py) and of the mode = nerf branch of
NeRFAll.render_rays (networks/renderer.py:219-240).  fp32 parity path (SIMT kernels of csrc/nerf_f32.cu)."""
import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib
from ._lib import FLAG_LINDISP, FLAG_TRAIN, FLAG_WHITE_BKGD, NerfMlp, NerfWeights, check, ptr, stream_ptr

_NERF_TENSORS = ([(f"pts_linears.{i}.weight", "pts_w", i) for i in range(8)] + [(f"pts_linears.{i}.bias", "pts_b", i) for i in range(8)] +
                 [("alpha_linear.weight", "alpha_w", None), ("alpha_linear.bias", "alpha_b", None),
                  ("feature_linear.weight", "feature_w", None), ("feature_linear.bias", "feature_b", None),
                  ("views_linears.0.weight", "views_w", None), ("views_linears.0.bias", "views_b", None),
                  ("rgb_linear.weight", "rgb_w", None), ("rgb_linear.bias", "rgb_b", None)])


def _nerf_struct(T, prefix):
    w = NerfWeights()
    for name, field, idx in _NERF_TENSORS:
        t = T.get(prefix + name)
        if idx is None:
            setattr(w, field, ptr(t))
        else:
            getattr(w, field)[idx] = ptr(t)
    return w


class NeRF:
    """networks/nerf.py:7-177 with D = 8, W = 256, skips = [4], use_viewdirs = True (the reference defaults)."""

    def __init__(self, params, prefix, extract_feature="before_linear", render_rmnearplane=0):
        self.extract_feature, self.rmnearplane = extract_feature, float(render_rmnearplane)
        self.keep = []
        self.prefix = prefix
        # reference-layout fp32 views: what the backward pass reads (edn_nerf_field_bwd)
        self.params = {k: v.detach().float().contiguous() for k, v in params.items() if k.startswith(prefix)}
        m = NerfMlp()

        def g(name):
            t = params[prefix + name].detach().to(torch.float32)
            if not t.is_cuda:
                raise RuntimeError("NeRF parameters must be CUDA tensors (no CPU path)")
            return t

        def keep(t):
            t = t.contiguous()
            self.keep.append(t)
            return t.data_ptr()

        if (prefix + "pts_linears.8.weight") in params or tuple(g("pts_linears.5.weight").shape) != (256, 319):
            raise RuntimeError("unsupported NeRF MLP (expected D=8, W=256, skips=[4], 63-d input)")
        for i in range(8):
            w = g(f"pts_linears.{i}.weight").t()                      # [in][256]
            if i == 0:
                w = F.pad(w, (0, 0, 0, 1))                                # 63 -> 64 rows
            elif i == 5:
                w = torch.cat([F.pad(w[:63], (0, 0, 0, 1)), w[63:]], 0)   # [PE 63 | 0 | h 256] -> 320 rows
            m.pts_t[i] = keep(w)
            m.pts_b[i] = keep(g(f"pts_linears.{i}.bias"))
        m.alpha_w, m.alpha_b = keep(g("alpha_linear.weight").reshape(-1)), keep(g("alpha_linear.bias"))
        m.feature_t, m.feature_b = keep(g("feature_linear.weight").t()), keep(g("feature_linear.bias"))
        if tuple(g("views_linears.0.weight").shape) != (128, 283):
            raise RuntimeError("unsupported views_linears shape")
        m.views_t, m.views_b = keep(g("views_linears.0.weight").t()), keep(g("views_linears.0.bias"))
        m.rgb_t = keep(F.pad(g("rgb_linear.weight").t(), (0, 1)))
        m.rgb_b = keep(g("rgb_linear.bias")) if (prefix + "rgb_linear.bias") in params else None
        self.m = m

    def mlpforward_at(self, ray_batch, z_vals, want_feature=False):
        """NeRF.mlpforward (nerf.py:46) at pts = o + d * z_vals -> (raw [R,S,4], feature [R,S,256] | None)."""
        rb, z = ray_batch.float().contiguous(), z_vals.float().contiguous()
        R, S = z.shape
        raw = torch.empty((R, S, 4), dtype=torch.float32, device=z.device)
        feat = torch.empty((R, S, 256), dtype=torch.float32, device=z.device) if want_feature else None
        check(_lib.load().edn_nerf_mlp_fwd(C.byref(self.m), ptr(rb), ptr(z), R, S, 1 if self.extract_feature == "after_linear" else 0,
                                           ptr(raw), ptr(feat), stream_ptr()), "edn_nerf_mlp_fwd")
        return raw, feat

    def backward(self, engine, ray_batch, z_vals, noise, d_rgb, d_depth, d_acc, d_ray_batch, white_bkgd=False, d_feat=None,
                 chunk_rays=4096, precision=_lib.EDN_F32):
        """Backward of mlpforward_at + raw2outputs -> {state_dict name: gradient}; d_ray_batch is accumulated in place."""
        lib = _lib.load()
        R, S = z_vals.shape
        grads = {k: torch.zeros_like(v) for k, v in self.params.items()}
        if R == 0:
            return grads
        if white_bkgd and d_rgb is not None:      # rgb_map += 1 - acc_map (nerf.py:126-127)
            d_acc = (d_acc if d_acc is not None else 0) - d_rgb.sum(-1)
        f = lambda t: None if t is None else t.detach().float().contiguous()
        keep = [f(t) for t in (ray_batch, z_vals, noise, d_rgb, d_depth, d_acc, d_feat)]
        w, g = _nerf_struct(self.params, self.prefix), _nerf_struct(grads, self.prefix)
        nbytes = int(lib.edn_nerf_bwd_workspace_bytes(min(R, int(chunk_rays)), S))
        ws = engine.workspace(nbytes)
        check(lib.edn_nerf_field_bwd(C.byref(w), ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), R, S, 0, int(precision), ptr(keep[3]), ptr(keep[4]),
                                     ptr(keep[5]), None, ptr(keep[6]), 1 if self.extract_feature == "after_linear" else 0, C.byref(g),
                                     ptr(d_ray_batch), ptr(ws), nbytes, stream_ptr()), "edn_nerf_field_bwd")
        return grads

    def raw2outputs(self, raw, z_vals, ray_batch, noise=None, white_bkgd=False, is_train=True):
        """NeRF.raw2outputs (nerf.py:74) -> (rgb_map, depth_map, acc_map, weights)."""
        R, S = z_vals.shape
        f32 = dict(dtype=torch.float32, device=z_vals.device)
        w, rgb, depth, acc = torch.empty((R, S), **f32), torch.empty((R, 3), **f32), torch.empty((R,), **f32), torch.empty((R,), **f32)
        flags = (FLAG_TRAIN if is_train else 0) | (FLAG_WHITE_BKGD if white_bkgd else 0)
        nz = None if noise is None else noise.float().contiguous()
        check(_lib.load().edn_nerf_raw2outputs(ptr(raw.contiguous()), ptr(z_vals.float().contiguous()), ptr(ray_batch.float().contiguous()),
                                               ptr(nz), R, S, flags, self.rmnearplane, ptr(w), ptr(rgb), ptr(depth), ptr(acc), stream_ptr()),
              "edn_nerf_raw2outputs")
        return rgb, depth, acc, w


def render_rays_nerf(engine, mlp_coarse, mlp_fine, ray_batch, N_samples, retraw=False, lindisp=False, perturb=0., N_importance=0,
                     white_bkgd=False, raw_noise_std=0., is_train=True, use_awp=False, inference=False, force_naive=False, rand=None):
    """mode = nerf branch of NeRFAll.render_rays (renderer.py:219-264).  `engine` supplies the sampler / linspace helpers."""
    rand = dict(rand or {})
    rb = ray_batch.float().contiguous()
    R, dev = rb.shape[0], rb.device
    Nc, Ni = int(N_samples), int(N_importance)
    f32 = dict(dtype=torch.float32, device=dev)
    t_rand = None
    engine._calls += 1
    if perturb > 0.:
        t_rand = (rand["t_rand"] if "t_rand" in rand else engine._random((R, Nc), 0)).float().contiguous()
    z0 = torch.empty((R, Nc), **f32)
    check(_lib.load().edn_place_samples(ptr(rb), ptr(engine._linspace(Nc)), ptr(t_rand), R, Nc, FLAG_LINDISP if lindisp else 0, ptr(z0),
                                        stream_ptr()), "edn_place_samples")
    want_feat = use_awp and not force_naive and not inference

    def noise(n, key):
        if key in rand:
            return rand[key]
        return engine._random((R, n), 1 if key == "noise0" else 3, normal=True, scale=raw_noise_std) if raw_noise_std > 0. else None

    raw0, feat0 = mlp_coarse.mlpforward_at(rb, z0, want_feat and Ni == 0)
    rgb0, depth0, acc0, w0 = mlp_coarse.raw2outputs(raw0, z0, rb, noise(Nc - 1, "noise0"), white_bkgd, is_train)
    ret = {"rgb_map": rgb0, "depth_map": depth0, "acc_map": acc0}
    if retraw:
        ret["z_vals"], ret["weights"] = z0, w0
    if Ni > 0:
        u = (rand["u"] if "u" in rand else engine._random((R, Ni), 2)) if perturb > 0. else None
        m = engine.sample_pdf_merge(z0, w0, Ni, u=u, want_indices=False)
        raw1, feat1 = mlp_fine.mlpforward_at(rb, m["z_vals"], want_feat)
        rgb1, depth1, acc1, w1 = mlp_fine.raw2outputs(raw1, m["z_vals"], rb, noise(Nc + Ni - 1, "noise1"), white_bkgd, is_train)
        ret = {"rgb_map": rgb1, "depth_map": depth1, "acc_map": acc1, "rgb0": rgb0, "depth0": depth0, "acc0": acc0, "z_std": m["z_std"]}
        if retraw:
            ret.update(z_vals=m["z_vals"], weights=w1, z_vals0=z0, weights0=w0)
        if want_feat:
            ret["depth_feature"], ret["z_vals"] = feat1, m["z_vals"]
    elif want_feat:
        ret["depth_feature"], ret["z_vals"] = feat0, z0
    return ret
