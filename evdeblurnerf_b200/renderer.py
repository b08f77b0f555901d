"""Host mirror of the reference renderer façade `NeRFAll` (networks/renderer.py:14-626), mode = c2f | nerf, kernel_type = RBK | DSK | none.

Same method names, argument order and return structure as the reference so that `run_nerf.py`'s call sites
(`nerf(H, W, K, chunk, rays=..., rays_info=..., **render_kwargs)`) keep working; every arithmetic
step is a kernel of libevdeblur_b200.so.  Round-1 scope: forward passes (training-branch forward and eval render);
parameters are read from a reference `state_dict()` (SURVEY Appendix A names).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import AwpParams, RbkParams, check, ptr, stream_ptr
from .engine import NerfRenderEngine, RenderEngine
from .losses import tv_loss_app


class RigidBlurringModel:
    """networks/dpnerf/blurmodel.py:8-173 (feat_ch = 0, depth-1 branches, use_origin = True: every shipped config)."""

    def __init__(self, params, num_motion, rv_window=0.1, prefix="kernelsnet."):
        self.num_motion, self.rv_window, self.prefix = int(num_motion), float(rv_window), prefix
        self.keep = []
        self.tensors = {}       # name (without prefix) -> fp32 view of the parameter (shares storage when already fp32)
        p = RbkParams()

        def g(name):
            t = params[prefix + name].detach().to(torch.float32).contiguous()
            if not t.is_cuda:
                raise RuntimeError("RigidBlurringModel parameters must be CUDA tensors")
            self.keep.append(t)
            self.tensors[name] = t
            return t

        emb = g("view_embed_module.img_embed")
        if emb.shape[1] != 32:
            raise RuntimeError("unsupported view-latent width %d (expected 32)" % emb.shape[1])
        p.img_embed, p.n_img = emb.data_ptr(), emb.shape[0]
        for h in ("r", "v", "w"):
            w, b = g(f"{h}_branch.0.weight"), g(f"{h}_branch.0.bias")
            if tuple(w.shape) != (32, 32) or (prefix + f"{h}_branch.1.weight") in params:
                raise RuntimeError("unsupported RBK branch shape (expected one 32x32 layer per head)")
            setattr(p, f"{h}_branch_w", w.data_ptr())
            setattr(p, f"{h}_branch_b", b.data_ptr())
            w, b = g(f"{h}_linear.weight"), g(f"{h}_linear.bias")
            exp_out = (self.num_motion + 1) if h == "w" else 3 * self.num_motion
            if tuple(w.shape) != (exp_out, 32):
                raise RuntimeError(f"unsupported {h}_linear shape {tuple(w.shape)} for num_motion={self.num_motion}")
            setattr(p, f"{h}_linear_w", w.data_ptr())
            setattr(p, f"{h}_linear_b", b.data_ptr())
        p.num_motion, p.rv_window = self.num_motion, self.rv_window
        self.p = p

    def warp(self, H, W, focal, rays, images_idx, near=0., far=1., ndc=True, want_new_rays=True, want_ray_batch=True):
        """Fused forward + render() prologue -> dict(new_rays [N,E,3,2], weight [N,E], img_embed [N,32], ray_batch [N*E,11])."""
        r = rays.detach().to(torch.float32).contiguous()
        idx = images_idx.reshape(-1).to(torch.int64).contiguous()
        N, E, dev = r.shape[0], self.num_motion + 1, r.device
        f32 = dict(dtype=torch.float32, device=dev)
        out = {"weight": torch.empty((N, E), **f32), "img_embed": torch.empty((N, 32), **f32),
               "new_rays": torch.empty((N, E, 3, 2), **f32) if want_new_rays else None,
               "ray_batch": torch.empty((N * E, 11), **f32) if want_ray_batch else None}
        check(_lib.load().edn_rbk_warp_ndc_fwd(C.byref(self.p), ptr(r), ptr(idx), N, int(H), int(W), float(focal), float(near),
                                               float(far), 1 if ndc else 0, ptr(out["new_rays"]), ptr(out["weight"]),
                                               ptr(out["img_embed"]), ptr(out["ray_batch"]), stream_ptr()), "edn_rbk_warp_ndc_fwd")
        return out

    def __call__(self, H, W, K, rays, rays_info, feats=None, return_img_embed=False):
        """Reference signature (blurmodel.py:129): -> (new_rays, weight, align_loss=None, extras)."""
        o = self.warp(H, W, float(K[0][0]), rays, rays_info["images_idx"], want_ray_batch=False)
        extras = {"img_embed": o["img_embed"]} if return_img_embed else {}
        return o["new_rays"], o["weight"], None, extras

    @staticmethod
    def rbk_weighted_sum(rgb, depth, acc, extras, ccw):
        """blurmodel.py:112-127: reduce [N*E, ...] -> [N, ...] with weights ccw [N,E] for rgb, depth, acc and every extra."""
        N, E = ccw.shape
        ws = lambda x: weighted_sum(x, ccw)
        out_extras = {k: (ws(v) if (isinstance(v, torch.Tensor) and v.shape[:1] == (N * E,)) else v) for k, v in extras.items()}
        return ws(rgb), ws(depth), ws(acc), out_extras


class AdaptiveWeightProposal:
    """networks/dpnerf/awp.py:9-117 + mam.py (D_sam = 4, W_sam = 64, D_mot = 1, W_mot = 32, ray_dir_freq = 2, view latent 32:
    the values run_nerf.py:203-212 passes for every shipped config).  Forward only, train-mode BatchNorm statistics."""

    ccw_fine_scale = 0.05     # awp.py:22

    def __init__(self, params, num_motion, prefix="awpnet.", bn_eps=1e-5, precision=_lib.EDN_F32):
        self.E, self.bn_eps, self.keep = int(num_motion) + 1, float(bn_eps), []
        self.params, self.prefix = params, prefix
        self.precision = int(precision)     # EDN_F32: fused fp32 kernels (parity); EDN_BF16: TF32 GEMM chain for the sample MLP
        p = AwpParams()

        def g(name, transpose=False, shape=None):
            t = params[prefix + name].detach().to(torch.float32)
            if not t.is_cuda:
                raise RuntimeError("AdaptiveWeightProposal parameters must be CUDA tensors")
            if shape is not None and tuple(t.shape) != shape:
                raise RuntimeError(f"unsupported shape for {prefix + name}: {tuple(t.shape)} (expected {shape})")
            t = t.t() if transpose else t
            t = t.contiguous()
            self.keep.append(t)
            return t.data_ptr()

        # depth_feature width = in-features of the first sample layer: 128 (c2f geo features) or 256 (mode = nerf, run_nerf.py:203-212)
        self.input_ch = int(params[prefix + "sample_feature_embed_layer.0.weight"].shape[1])
        for l in range(4):
            p.sample_t[l] = g(f"sample_feature_embed_layer.{l}.weight", True, (64, self.input_ch) if l == 0 else (64, 64))
            p.sample_b[l] = g(f"sample_feature_embed_layer.{l}.bias")
        p.input_ch = self.input_ch
        if (prefix + "sample_feature_embed_layer.4.weight") in params or (prefix + "motion_feature_embed_layer.2.weight") in params:
            raise RuntimeError("unsupported AWP depth")
        p.motion_w[0], p.motion_b[0] = g("motion_feature_embed_layer.0.weight", shape=(32, 111)), g("motion_feature_embed_layer.0.bias")
        p.motion_w[1], p.motion_b[1] = g("motion_feature_embed_layer.1.weight", shape=(32, 32)), g("motion_feature_embed_layer.1.bias")
        p.mam_linear_t, p.mam_linear_b = g("MAM.linear.weight", True, (32, 64)), g("MAM.linear.bias")
        p.line_conv_att = g("MAM.Corr.line_conv_att.weight")
        for n_, shp in (("conva", (16, 32, 1)), ("convb", (16, 32, 1)), ("convc", (16, 32, 1)), ("convn", (16, 16, 1)), ("convl", (16, 16, 1))):
            setattr(p, n_, g(f"MAM.Corr.{n_}.weight", shape=shp))
        p.convd_w = g("MAM.Corr.convd.0.weight", shape=(32, 32, 1))
        p.bn_weight, p.bn_bias = g("MAM.Corr.convd.1.weight"), g("MAM.Corr.convd.1.bias")
        p.w_linear_w, p.w_linear_b = g("w_linear.weight", shape=(self.E, 32)), g("w_linear.bias")
        self.p = p

    def __call__(self, depth_feature, z_vals, rays_d, view_feature):
        """awp.py:79: depth_feature [N*E,S,128], z_vals [N*E,S], rays_d [N*E,3] (may be a strided view), view_feature [N,32]."""
        from .autograd import AWP_PARAM_NAMES, AwpFn
        ps = [self.params[self.prefix + n] for n in AWP_PARAM_NAMES]
        if torch.is_grad_enabled() and (depth_feature.requires_grad or rays_d.requires_grad or view_feature.requires_grad
                                        or any(t.requires_grad for t in ps)):
            return AwpFn.apply(self, depth_feature, z_vals, rays_d, view_feature, *ps)
        return self.run(depth_feature, z_vals, rays_d, view_feature)

    # ---- synchronised BatchNorm: when `sync_bn` is set (the Trainer does, for the process group it trains on)
    #      the batch sums of CorrelationModule's BatchNorm1d -- and the row count behind them, so ragged shards are fine -- are
    #      all-reduced between two phases of the pass: N ranks x N/world rays reproduce the single-GPU statistics.  Off by
    #      default: a stand-alone (e.g. rank-0-only validation) call must not enter a collective -------------------------------
    sync_bn = False
    group = None            # torch.distributed process group of the exchange (None = default group)

    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size(self.group) if (self.sync_bn and dist.is_available() and dist.is_initialized()) else 1

    def options(self, keep_activations=False, phase=0, bn_rows_total=0):
        # widths other than 128 have no fused per-sample kernel: they always run the materialised (GEMM) path
        keep = keep_activations or self.input_ch != 128
        return _lib.AwpOptions(self.precision, 1 if keep else 0, int(phase), int(bn_rows_total))

    def _all_reduce_block(self, ws, offset_floats, n_doubles):
        import torch.distributed as dist
        block = ws[offset_floats: offset_floats + 2 * n_doubles].view(torch.float64)
        dist.all_reduce(block, group=self.group)

    def run(self, depth_feature, z_vals, rays_d, view_feature, workspace=None, keep_activations=False):
        df, z = depth_feature.detach().float().contiguous(), z_vals.detach().float().contiguous()
        NE, S, Fd = df.shape
        if Fd != self.input_ch:
            raise RuntimeError(f"AWP: depth_feature has {Fd} channels, sample_feature_embed_layer.0 expects {self.input_ch}")
        E = self.E
        N = NE // E
        rd = rays_d.detach().float()
        if rd.stride(-1) != 1:
            rd = rd.contiguous()
        vf = view_feature.detach().float().contiguous()
        lib = _lib.load()
        opt = self.options(keep_activations)
        ws = workspace if workspace is not None else torch.empty((int(lib.edn_awp_workspace_floats(N, E, S, C.byref(opt))),),
                                                                 dtype=torch.float32, device=df.device)
        ccw = torch.empty((N, E), dtype=torch.float32, device=df.device)
        world = self._world()
        # phases 1 / 2 with bn_rows_total = 0: the kernels take the row count from the all-reduced block (sums + count)
        phases = [self.options(keep_activations)] if world == 1 else [self.options(keep_activations, 1), self.options(keep_activations, 2)]
        for o in phases:
            check(lib.edn_awp_fwd(C.byref(self.p), ptr(df), ptr(z), rd.data_ptr(), int(rd.stride(0)), ptr(vf), N, E, S, self.bn_eps,
                                  C.byref(o), ptr(ws), ptr(ccw), stream_ptr()), "edn_awp_fwd")
            if o.phase == 1:
                self._all_reduce_block(ws, int(lib.edn_awp_stats_offset_floats(N, E, S)), 66)
        self.last_stats = ws[int(lib.edn_awp_stats_offset_floats(N, E, S)):][:132].view(torch.float64)   # sums, sums of squares, rows
        return ccw


def weighted_sum(x, ccw):
    """x [N*E, ...] , ccw [N,E] -> [N, ...]."""
    if torch.is_grad_enabled() and (x.requires_grad or ccw.requires_grad):
        from .autograd import WeightedSumFn
        return WeightedSumFn.apply(x, ccw)
    N, E = ccw.shape
    xf = x.detach().to(torch.float32).contiguous()
    Cn = max(1, xf.numel() // max(1, N * E))
    out = torch.empty((N,) + tuple(xf.shape[1:]), dtype=torch.float32, device=xf.device)
    check(_lib.load().edn_weighted_sum(ptr(xf), ptr(ccw.detach().float().contiguous()), ptr(out), N, E, Cn, stream_ptr()),
          "edn_weighted_sum")
    return out


def normalize_ccw(ccw, scale):
    """renderer.py:316-317: ccw + ccw * scale, renormalised over the exposures (row-wise, [N,E] is tiny: host tensor ops)."""
    c = ccw + ccw * scale
    return c / torch.sum(c, -1, keepdim=True)


def get_rays(H, W, K, c2w):
    """utils/rays.py:8-22 (HALF_PIX = 0.5): full-image camera rays [H, W, 3, 2] of pose c2w [3|4, 4] (eval-path ray generation)."""
    dev = c2w.device
    i, j = torch.meshgrid(torch.linspace(0, W - 1, W, device=dev), torch.linspace(0, H - 1, H, device=dev), indexing="xy")
    dirs = torch.stack([(i + (0.5 - float(K[0][2]))) / float(K[0][0]), -(j + (0.5 - float(K[1][2]))) / float(K[1][1]),
                        -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return torch.stack([rays_o, rays_d], -1)


def build_ray_batch(H, W, focal, rays, near=0., far=1., ndc=True):
    """render() prologue (renderer.py:423-446): rays [R,3,2] -> ray_batch [R,11]."""
    r = rays.detach().to(torch.float32).contiguous().reshape(-1, 3, 2)
    out = torch.empty((r.shape[0], 11), dtype=torch.float32, device=r.device)
    check(_lib.load().edn_build_ray_batch(ptr(r), r.shape[0], int(H), int(W), float(focal), float(near), float(far),
                                          1 if ndc else 0, ptr(out), stream_ptr()), "edn_build_ray_batch")
    return out


def init_reference_parameters(args, device="cuda", seed=None):
    """The parameters `NeRFAll(args, ...)` creates in the reference for mode = c2f (renderer.py:49-82; voxnerf.py:40-118): same
    state_dict names, shapes and initialisation -- VM planes / lines 0.1 * randn (init_one_svd), nn.Linear default init
    (U(-1/sqrt(in), 1/sqrt(in))), biases only with --rgb_add_bias."""
    g = torch.Generator().manual_seed(int(seed)) if seed is not None else None
    P = {}

    def lin(name, out_c, in_c, bias=False):
        b = 1.0 / (in_c ** 0.5)
        P[name + ".weight"] = ((torch.rand(out_c, in_c, generator=g) * 2 - 1) * b).to(device)
        if bias:
            P[name + ".bias"] = ((torch.rand(out_c, generator=g) * 2 - 1) * b).to(device)

    amin, amax = (torch.as_tensor(t, dtype=torch.float32).cpu() for t in args.bounding_box)
    pe_pts = 3 + 6 * args.multires
    pe_dir = (3 + 6 * args.multires_views) if args.use_viewdirs else 0
    fields = [("mlp_coarse.", args.coarse_n_voxels, args.coarse_app_n_comp, args.coarse_app_dim, args.coarse_app_dim + pe_pts,
               args.coarse_num_layers, args.coarse_hidden_dim, args.kernel_feat_cnl, args.coarse_num_layers_color)]
    if args.N_importance > 0:
        fields.append(("mlp_fine.", args.fine_n_voxels, args.fine_app_n_comp, args.fine_app_dim,
                       args.coarse_app_dim + args.fine_app_dim + pe_pts, args.fine_num_layers, args.fine_hidden_dim,
                       args.fine_geo_feat_dim, args.fine_num_layers_color))
    for pre, n_vox, n_comp, app_dim, in_ch, n_layers, hidden, geo, n_layers_color in fields:
        voxel = ((amax - amin).prod() / n_vox).pow(1 / 3)                   # voxnerf.py:87-89
        gs = ((amax - amin) / voxel).long().tolist()
        for i, ((m0, m1), v) in enumerate((((0, 1), 2), ((0, 2), 1), ((1, 2), 0))):
            P[pre + f"app_plane.{i}"] = (0.1 * torch.randn((1, n_comp[i], gs[m1], gs[m0]), generator=g)).to(device)
            P[pre + f"app_line.{i}"] = (0.1 * torch.randn((1, n_comp[i], gs[v], 1), generator=g)).to(device)
        lin(pre + "basis_mat", app_dim, sum(n_comp))
        for l in range(n_layers):                                              # voxnerf.py:49-62
            lin(pre + f"sigma_net.{l}", (1 + geo) if l == n_layers - 1 else hidden, in_ch if l == 0 else hidden)
        for l in range(n_layers_color):                                        # voxnerf.py:68-80
            lin(pre + f"color_net.{l}", 3 if l == n_layers_color - 1 else hidden, (pe_dir + geo) if l == 0 else hidden, bias=args.rgb_add_bias)
    return P


class NeRFAll:
    """Drop-in for the reference façade (mode = c2f or nerf, kernel_type = RBK, DSK or none -- all detected from the parameter
    names; `kernel_cfg` carries the DSK options the shapes cannot tell: in_embed, spatial_embed, kernel_hwindow, random_hwindow).
    Two constructor shapes:
      NeRFAll(params, aabb_min, aabb_max, kernel_ptnum=5, precision=..., ...)   `params`: a reference state_dict
      NeRFAll(args, kernelsnet=None, awpnet=None, precision=..., device=...)    the reference's own signature (renderer.py:15):
          `args` is the option namespace (bounding_box, *_n_voxels, kernel_ptnum, kernel_use_awp, render_rmnearplane ...); the PDRF
          fields are created with the reference's initialisation, `kernelsnet` / `awpnet` are objects with a `state_dict()` (the
          reference's nn.Modules) or plain dicts.  Either way the object then offers the nn.Module surface run_nerf.py uses:
          parameters() / named_parameters() / state_dict() / load_state_dict() / get_parameters() / train() / eval() / zero_grad()."""

    def __init__(self, params, aabb_min=None, aabb_max=None, kernel_ptnum=5, precision="fp32", render_rmnearplane=0, use_awp=False,
                 awpnet=None, device="cuda", seed=None, kernel_cfg=None):
        if not isinstance(params, dict):                     # reference signature: (args, kernelsnet, awpnet)
            args, kernelsnet = params, aabb_min
            awpnet = aabb_max if aabb_max is not None else awpnet
            if getattr(args, "mode", "c2f") != "c2f":
                raise NotImplementedError("NeRFAll(args, ...): mode = nerf models are built from a state_dict (NeRFAll(params, ...))")
            if getattr(args, "kernel_type", "RBK") not in ("RBK", "DSK", "none"):
                raise NotImplementedError(f"kernel_type {args.kernel_type!r}: the PBE two-stage blur kernel is not built (DESIGN.md section 8)")
            if getattr(args, "kernel_type", "RBK") == "DSK":      # run_nerf.py:184-203 option names
                kernel_cfg = dict(in_embed=getattr(args, "kernel_rand_embed", 3), spatial_embed=getattr(args, "kernel_spatial_embed", 0),
                                  kernel_hwindow=getattr(args, "kernel_hwindow", 10), random_hwindow=getattr(args, "kernel_random_hwindow", 0.25))
            P = init_reference_parameters(args, device=device, seed=seed)
            for pre, mod in (("kernelsnet.", kernelsnet), ("awpnet.", awpnet)):
                if mod is not None:
                    sd = mod if isinstance(mod, dict) else mod.state_dict()
                    P.update({pre + k: v.detach().to(device) for k, v in sd.items()})
            for v in P.values():
                if v.is_floating_point():
                    v.requires_grad_(True)
            self.args = args
            params = P
            aabb_min, aabb_max = ([float(x) for x in t] for t in args.bounding_box)
            kernel_ptnum, render_rmnearplane = args.kernel_ptnum, args.render_rmnearplane
            use_awp = bool(args.kernel_use_awp) and awpnet is not None
        self.params = {k: v for k, v in params.items() if isinstance(v, torch.Tensor)}
        self.mode = "nerf" if "mlp_coarse.pts_linears.0.weight" in self.params else "c2f"
        if self.mode == "nerf":
            self.engine = NerfRenderEngine(self.params, rmnearplane=render_rmnearplane, use_awp=bool(use_awp))
        else:
            self.engine = RenderEngine(self.params, aabb_min, aabb_max, precision=precision, rmnearplane=render_rmnearplane)
        self.kernelsnet = None
        if "kernelsnet.r_linear.weight" in self.params:
            self.kernelsnet = RigidBlurringModel(self.params, kernel_ptnum - 1)
        self.kernel_type, self.use_awp = "RBK", bool(use_awp)
        if "kernelsnet.pattern_pos" in self.params:           # deformable sparse kernel (pdrf/blurmodel.py)
            from .dsk import BlurModel
            self.kernelsnet, self.kernel_type = BlurModel(self.params, kernel_ptnum, **(kernel_cfg or {})), "DSK"
        self.awpnet = (AdaptiveWeightProposal(self.params, kernel_ptnum - 1,
                                              precision=_lib.EDN_BF16 if precision == "bf16" else _lib.EDN_F32) if self.use_awp else None)
        self.training = True
        self.backward_chunk_rays = 8192     # rays per recompute chunk of the backward pass (workspace ~ 8.6 KB x samples)
        self.last_render = None
        self._grad_names = sorted(k for k in self.params if k.startswith(("mlp_coarse.", "mlp_fine.", "kernelsnet."))
                                  and self.params[k].is_floating_point())
        self._packed_version = self._param_version()

    # ---- nn.Module surface used by run_nerf.py (optimizer construction :243-274, checkpoints :282-295, 628-634) --------------
    _BUFFERS = (".running_mean", ".running_var", ".num_batches_tracked")

    def named_parameters(self):
        return [(k, v) for k, v in self.params.items() if v.is_floating_point() and not k.endswith(self._BUFFERS)]

    def parameters(self):
        return [v for _, v in self.named_parameters()]

    def get_parameters(self, type, match_re=None, not_match_re=None):
        """renderer.py:108-127: "net" / "vol" parameter lists filtered by regular expressions on the names."""
        import re
        hit = lambda text, rx: len(re.findall(rx, text)) > 0
        vol = lambda k: "app_plane" in k or "app_line" in k
        return [v for k, v in self.named_parameters()
                if (match_re is None or hit(k, match_re)) and (not_match_re is None or not hit(k, not_match_re))
                and ((type == "net" and not vol(k)) or (type == "vol" and vol(k)))]

    def state_dict(self):
        return {k: v.detach() for k, v in self.params.items()}

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self.params if k not in sd]
        unexpected = [k for k in sd if k not in self.params]
        if strict and (missing or unexpected):
            raise RuntimeError(f"load_state_dict: missing {missing[:4]}, unexpected {unexpected[:4]}")
        with torch.no_grad():
            for k, v in sd.items():
                if k in self.params:
                    self.params[k].copy_(v)
        self.repack()
        return missing, unexpected

    def zero_grad(self, set_to_none=True):
        for v in self.parameters():
            if set_to_none:
                v.grad = None
            elif v.grad is not None:
                v.grad.zero_()

    def cuda(self, *a, **k):
        return self

    def to(self, *a, **k):
        return self

    # ---- autograd plumbing (the reference trains through torch autograd, run_nerf.py:594) --------------------------------
    def _param_version(self):
        return sum(int(v._version) for v in self.params.values())

    def _wants_grad(self):
        return torch.is_grad_enabled() and any(v.requires_grad for v in self.params.values())

    def repack(self):
        """Refresh the render-layout copies after the parameters changed (optimizer.step(), load_state_dict)."""
        self.engine.repack(self.params)
        if self.use_awp:
            old = self.awpnet
            self.awpnet = AdaptiveWeightProposal(self.params, old.E - 1, precision=old.precision)
            self.awpnet.sync_bn, self.awpnet.group = old.sync_bn, old.group
        self._packed_version = self._param_version()

    def _maybe_repack(self):
        # in-place updates (optimizer steps) bump the tensors' version counters
        if self._param_version() != self._packed_version:
            self.repack()

    def _render_sub_rays(self, H, W, K, rays, images_idx, near, far, ndc, kwargs, blur=True):
        """Differentiable warp + render of the sub-rays -> (rgb, depth, acc, rgb0, depth0, acc0, weight1), see autograd.py."""
        from .autograd import RenderSubRaysFn
        names = self._grad_names
        return RenderSubRaysFn.apply(self, self.kernelsnet if blur else None, H, W, float(K[0][0]), rays, images_idx, near, far, ndc,
                                     kwargs, names,
                                     *[self.params[n] for n in names])

    def train(self, mode=True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    # ---- renderer.py:129 ----------------------------------------------------------------------------------------------
    def render_rays(self, ray_batch, N_samples, retraw=False, lindisp=False, perturb=0., N_importance=0, white_bkgd=False,
                    raw_noise_std=0., pytest=False, force_naive=False, inference=False, **extra):
        self._maybe_repack()      # parameters updated in place (optimizer.step()) since the last render
        return self.engine.render_rays(ray_batch, N_samples, retraw=retraw, lindisp=lindisp, perturb=perturb,
                                       N_importance=N_importance, white_bkgd=white_bkgd, raw_noise_std=raw_noise_std,
                                       pytest=pytest, force_naive=force_naive, inference=inference, is_train=self.training,
                                       use_awp=self.use_awp, **extra)

    def _render_batch(self, ray_batch, rays_shape, **kwargs):
        all_ret = self.render_rays(ray_batch, **kwargs)     # no chunk loop: the fused kernels keep no per-sample activations
        for k in all_ret:
            all_ret[k] = all_ret[k].reshape(list(rays_shape) + list(all_ret[k].shape[1:]))
        k_extract = ["rgb_map", "depth_map", "acc_map"]
        return [all_ret[k] for k in k_extract] + [{k: all_ret[k] for k in all_ret if k not in k_extract}]

    # ---- renderer.py:399 ----------------------------------------------------------------------------------------------
    def render(self, H, W, K, chunk, rays=None, c2w=None, ndc=True, near=0., far=1., use_viewdirs=False, c2w_staticcam=None,
               **kwargs):
        # (`c2w` is accepted and, as in the reference, not used: renderer.py:399-446 only ever reads `rays`)
        rb = build_ray_batch(H, W, float(K[0][0]), rays, near, far, ndc)
        if c2w_staticcam is not None:       # renderer.py:428-431: camera rays from the static pose, view directions from `rays`
            static = get_rays(H, W, K, torch.as_tensor(c2w_staticcam, dtype=torch.float32, device=rb.device)).reshape(-1, 3, 2)
            if static.shape[0] != rb.shape[0]:
                raise RuntimeError("c2w_staticcam: `rays` must cover the full H x W image")
            rb_static = build_ray_batch(H, W, float(K[0][0]), static, near, far, ndc)
            rb_static[:, 8:11] = rb[:, 8:11]
            rb = rb_static
        return self._render_batch(rb, rays.shape[:-2], **kwargs)

    # ---- renderer.py:266 ----------------------------------------------------------------------------------------------
    def forward(self, H, W, K, chunk=1024 * 32, rays=None, rays_info=None, poses=None, **kwargs):
        if not self.training:
            assert poses is not None, "Please specify poses when in the eval model"
            return self.render_path(H, W, K, chunk, poses, **kwargs)
        assert rays is not None, "Please specify rays when in the training mode"
        force_baseline = kwargs.pop("force_naive", True)
        return_pts0_rgb = kwargs.pop("return_pts0_rgb", False)
        want_tv = kwargs.pop("want_tv", True)     # not in the reference: lets a caller skip the TV term it would discard
        N_importance = kwargs.get("N_importance", 0)
        ndc, near, far = kwargs.pop("ndc", True), kwargs.pop("near", 0.), kwargs.pop("far", 1.)
        kwargs.pop("use_viewdirs", None)
        other_loss, other_tensors = {}, {}
        self._maybe_repack()
        if self.kernel_type == "DSK" and not force_baseline:
            return self._forward_dsk(H, W, K, rays, rays_info, return_pts0_rgb, N_importance, ndc, near, far, kwargs, want_tv)
        if self._wants_grad():
            return self._forward_with_grad(H, W, K, rays, rays_info, force_baseline, return_pts0_rgb, N_importance, ndc, near, far,
                                           kwargs, want_tv)
        if self.kernelsnet is not None and self.kernel_type == "RBK" and not force_baseline:
            k = self.kernelsnet.warp(H, W, float(K[0][0]), rays, rays_info["images_idx"], near, far, ndc, want_new_rays=False)
            weight1 = k["weight"]
            N, E = weight1.shape
            rgb, depth, acc, extras = self._render_batch(k["ray_batch"], (N * E,), **kwargs)
            rgb_pts = rgb.reshape(N, E, 3)
            if self.use_awp:     # renderer.py:310-330
                ccw = self.awpnet(extras["depth_feature"], extras["z_vals"], k["ray_batch"][:, 3:6], k["img_embed"])
                ccw = normalize_ccw(ccw, self.awpnet.ccw_fine_scale)
                other_tensors["rgb_awp"] = weighted_sum(rgb, ccw)
                other_tensors["ccw_fine"] = ccw
                other_tensors["stage1_img_embed"] = k["img_embed"]
            rgb_b = weighted_sum(rgb, weight1)
            rgb1 = None
            if N_importance > 0:
                rgb1_pts = extras["rgb0"].reshape(N, E, 3)
                rgb1 = weighted_sum(extras["rgb0"], weight1)
            if self.mode == "c2f" and want_tv:
                other_loss["TV"] = self.tv_loss(N_importance > 0)
            if return_pts0_rgb:
                other_tensors["stage1_rgb_pts0"] = rgb_pts[:, 0]
                if N_importance > 0:
                    other_tensors["stage1_rgb1_pts0"] = rgb1_pts[:, 0]
            return rgb_b, rgb1, other_loss, other_tensors
        rgb, depth, acc, extras = self.render(H, W, K, chunk, rays, ndc=ndc, near=near, far=far, **kwargs)
        other_tensors["stage1_rgb_pts0"] = rgb
        if N_importance > 0:
            other_tensors["stage1_rgb1_pts0"] = extras["rgb0"]
        if self.mode == "c2f" and want_tv:
            other_loss["TV"] = self.tv_loss(N_importance > 0)
        return rgb, extras.get("rgb0"), other_loss, other_tensors

    __call__ = forward

    def forward_fused(self, H, W, K, rays, rays_info, naive_rays, **kwargs):
        """Training-branch forward of the blurred rays (renderer.py:277-378) TOGETHER with the force_naive renders of the
        ray sets in `naive_rays` (the event start / end rays, run_nerf.py:534-557): one fused render launch and one backward
        instead of three -> (rgb, rgb0, other_loss, other_tensors, [(rgb_i, rgb0_i), ...]).  Autograd mode only."""
        self._maybe_repack()
        return_pts0_rgb = kwargs.pop("return_pts0_rgb", False)
        want_tv = kwargs.pop("want_tv", True)
        ndc, near, far = kwargs.pop("ndc", True), kwargs.pop("near", 0.), kwargs.pop("far", 1.)
        kwargs.pop("use_viewdirs", None)
        kwargs.pop("force_naive", None)
        return self._forward_with_grad(H, W, K, rays, rays_info, False, return_pts0_rgb, kwargs.get("N_importance", 0), ndc, near, far,
                                       kwargs, want_tv, naive_rays=list(naive_rays))

    def _forward_with_grad(self, H, W, K, rays, rays_info, force_baseline, return_pts0_rgb, N_importance, ndc, near, far, kwargs,
                           want_tv=True, naive_rays=None):
        """Training branch of forward() (renderer.py:277-378) with outputs attached to the autograd graph."""
        other_loss, other_tensors = {}, {}
        blur = self.kernelsnet is not None and self.kernel_type == "RBK" and not force_baseline
        if naive_rays:
            kwargs = dict(kwargs, extra_rays=torch.cat([r.reshape(-1, 3, 2) for r in naive_rays], 0))
        rgb, depth, acc, rgb0, depth0, acc0, weight1, feat, rb, img_embed = self._render_sub_rays(
            H, W, K, rays, rays_info["images_idx"] if blur else None, near, far, ndc, kwargs, blur=blur)
        naive_out = []
        if naive_rays:     # rows of the extra rays follow the main sub-rays
            n_main = rb.shape[0] - sum(r.reshape(-1, 3, 2).shape[0] for r in naive_rays)
            lo = n_main
            for r in naive_rays:
                hi = lo + r.reshape(-1, 3, 2).shape[0]
                naive_out.append((rgb[lo:hi], rgb0[lo:hi] if N_importance > 0 else None))
                lo = hi
            rgb, depth, acc, rgb0, depth0, acc0 = (t[:n_main] for t in (rgb, depth, acc, rgb0, depth0, acc0))
            rb = rb[:n_main]
            if feat.numel():
                feat = feat[:n_main]
        if blur:
            N, E = weight1.shape
            if self.use_awp:     # renderer.py:310-330
                ccw = self.awpnet(feat, self.last_render["z_vals"][:rb.shape[0]], rb[:, 3:6], img_embed)
                ccw = normalize_ccw(ccw, self.awpnet.ccw_fine_scale)
                other_tensors["rgb_awp"] = weighted_sum(rgb, ccw)
                other_tensors["ccw_fine"] = ccw
                other_tensors["stage1_img_embed"] = img_embed
            rgb_b = weighted_sum(rgb, weight1)
            rgb1 = weighted_sum(rgb0, weight1) if N_importance > 0 else None
            if return_pts0_rgb:
                other_tensors["stage1_rgb_pts0"] = rgb.reshape(N, E, 3)[:, 0]
                if N_importance > 0:
                    other_tensors["stage1_rgb1_pts0"] = rgb0.reshape(N, E, 3)[:, 0]
        else:
            rgb_b, rgb1 = rgb, (rgb0 if N_importance > 0 else None)
            other_tensors["stage1_rgb_pts0"] = rgb
            if N_importance > 0:
                other_tensors["stage1_rgb1_pts0"] = rgb0
        if want_tv:
            other_loss["TV"] = self.tv_loss(N_importance > 0)
        if naive_rays:
            return rgb_b, rgb1, other_loss, other_tensors, naive_out
        return rgb_b, rgb1, other_loss, other_tensors

    def _forward_dsk(self, H, W, K, rays, rays_info, return_pts0_rgb, N_importance, ndc, near, far, kwargs, want_tv=True):
        """renderer.py:301-378 with kernel_type = DSK: kernel rays -> render of the N * num_pt rays -> per-point weighted sums (and,
        with kernel_use_awp, AWP's weights over the points: renderer.py:310-343); `align` joins the extra losses.  With parameters
        that require grad the stages are autograd nodes (DskRaysFn -> RenderSubRaysFn, which returns d rays -> AwpFn / WeightedSumFn)."""
        other_loss, other_tensors = {}, {}
        new_rays, weight1, align, kex = self.kernelsnet(H, W, K, rays, rays_info, return_img_embed=self.use_awp,
                                                        noise=kwargs.pop("dsk_noise", None))
        N, E = weight1.shape
        sub = new_rays.reshape(-1, 3, 2)
        if self._wants_grad():
            rgb, _, _, rgb0, _, _, _, feat, rb, _ = self._render_sub_rays(H, W, K, sub, None, near, far, ndc, kwargs, blur=False)
            z_vals = self.last_render["z_vals"]
        else:
            rb = build_ray_batch(H, W, float(K[0][0]), sub, near, far, ndc)
            rgb, _, _, extras = self._render_batch(rb, (N * E,), **kwargs)
            rgb0, feat, z_vals = extras.get("rgb0"), extras.get("depth_feature"), extras.get("z_vals")
        if self.use_awp:
            ccw = self.awpnet(feat, z_vals, rb[:, 3:6], kex["img_embed"])
            ccw = normalize_ccw(ccw, self.awpnet.ccw_fine_scale)
            other_tensors["rgb_awp"] = weighted_sum(rgb, ccw)
            other_tensors["ccw_fine"] = ccw
            other_tensors["stage1_img_embed"] = kex["img_embed"]
        rgb_b = weighted_sum(rgb, weight1)
        rgb1 = weighted_sum(rgb0, weight1) if N_importance > 0 else None
        if self.mode == "c2f" and want_tv:
            other_loss["TV"] = self.tv_loss(N_importance > 0)
        other_loss["align"] = align.reshape(1, 1)
        if return_pts0_rgb:
            other_tensors["stage1_rgb_pts0"] = rgb.reshape(N, E, 3)[:, 0]
            if N_importance > 0:
                other_tensors["stage1_rgb1_pts0"] = rgb0.reshape(N, E, 3)[:, 0]
        return rgb_b, rgb1, other_loss, other_tensors

    def render_blurred(self, H, W, K, rays, images_idx, near=0., far=1., ndc=True, **kwargs):
        """The render part of the training forward (renderer.py:303-343 without the loss terms): blur-kernel warp ->
        NDC ray batch -> c2f render of the N*E sub-rays -> exposure-weighted sum.  Returns (rgb [N,3], rgb0 [N,3] | None)."""
        self._maybe_repack()
        if self.kernel_type == "DSK":
            raise NotImplementedError("render_blurred takes RBK rays; with kernel_type = DSK call forward(rays_info=...) (pixel coordinates + poses)")
        if self._wants_grad():
            rgb, _, _, rgb0, _, _, weight1 = self._render_sub_rays(H, W, K, rays, images_idx, near, far, ndc, kwargs)[:7]
            if self.kernelsnet is None:
                return rgb, (rgb0 if kwargs.get("N_importance", 0) > 0 else None)
            return weighted_sum(rgb, weight1), (weighted_sum(rgb0, weight1) if kwargs.get("N_importance", 0) > 0 else None)
        if self.kernelsnet is None:           # kernel_type = none: one exposure per ray, no blending
            out = self.render_rays(build_ray_batch(H, W, float(K[0][0]), rays, near, far, ndc), **kwargs)
            return out["rgb_map"], out.get("rgb0")
        k = self.kernelsnet.warp(H, W, float(K[0][0]), rays, images_idx, near, far, ndc, want_new_rays=False)
        out = self.render_rays(k["ray_batch"], **kwargs)
        rgb = weighted_sum(out["rgb_map"], k["weight"])
        rgb0 = weighted_sum(out["rgb0"], k["weight"]) if "rgb0" in out else None
        return rgb, rgb0

    def tv_loss(self, with_fine=True):
        """renderer.py:361-365: (TV_loss_app(coarse) [+ TV_loss_app(fine)]) * 5 (mode = c2f only)."""
        if self.mode != "c2f":
            return None
        self._maybe_repack()
        tv = tv_loss_app(self.params, "mlp_coarse.")
        if with_fine:
            tv = tv + tv_loss_app(self.params, "mlp_fine.")
        return tv * 5

    # ---- renderer.py:594 + utils/rays.py:8-22 ---------------------------------------------------------------------------
    def render_path(self, H, W, K, chunk, render_poses, render_kwargs=None, render_factor=0, **kwargs):
        if render_factor != 0:
            H, W = H // render_factor, W // render_factor
        kw = dict(render_kwargs or {})
        kw.update(kwargs)
        ndc, near, far = kw.pop("ndc", True), kw.pop("near", 0.), kw.pop("far", 1.)
        kw.pop("use_viewdirs", None)
        kw.setdefault("inference", True)
        rgbs, depths = [], []
        dev = self.engine.device
        for c2w in render_poses:
            c2w = torch.as_tensor(c2w, dtype=torch.float32, device=dev)
            rays = get_rays(H, W, K, c2w).reshape(-1, 3, 2)
            rb = build_ray_batch(H, W, float(K[0][0]), rays, near, far, ndc)
            rgb, depth, acc, _ = self._render_batch(rb, (H, W), **kw)
            rgbs.append(rgb)
            depths.append(depth)
        return torch.stack(rgbs, 0), torch.stack(depths, 0)
