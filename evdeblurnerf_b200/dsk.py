"""Host mirror of the reference's deformable sparse kernel `BlurModel` with kernel_type = DSK (networks/pdrf/blurmodel.py:9-224):
same call signature and return tuple, parameters under the reference's state_dict names, compute through the C ABI
(`edn_dsk_rays_fwd` / `edn_dsk_rays_bwd`, csrc/dsk.cu).  The PBE variant (two-stage render with composited coarse features,
renderer.py:289-299, voxnerf.py:223-239) is not built.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import DskGrads, DskParams, check, ptr, stream_ptr


def _f32(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


class BlurModel:
    """`BlurModel(params, num_pt, kernel_hwindow=10, in_embed=3, spatial_embed=0, random_hwindow=0.25)`: `params` is a dict of CUDA
    tensors under `prefix` with the reference's names -- img_embed.img_embed [n_img, C], pattern_pos [n_img | 1, num_pt, 2],
    optional pattern_trans (optim_trans), linears.{0,2,..}.{weight,bias}, linears1.{0,2}.{weight,bias}.  Layer count, width,
    short_cut, isglobal and optim_spatialvariant_trans are read from the shapes; the two embedding depths cannot be and are
    checked against linears.0's input width (blurmodel.py:94-97)."""

    kernel_type = "DSK"

    def __init__(self, params, num_pt, kernel_hwindow=10, in_embed=3, spatial_embed=0, random_hwindow=0.25, prefix="kernelsnet."):
        self.prefix, self.num_pt = prefix, int(num_pt)
        self.kernel_hwindow, self.random_hwindow = float(kernel_hwindow), float(random_hwindow)
        self.in_embed, self.spatial_embed = int(in_embed), int(spatial_embed)
        self.params = params
        has = lambda n: (prefix + n) in params
        if has("view_embed_linears.0.weight") or has("img_embed.view_embed_linears.0.weight"):
            raise NotImplementedError("BlurModel: kernel_img_embed_type = param_mlp is not built (plain per-image embedding only)")
        names = ["img_embed.img_embed", "pattern_pos"] + (["pattern_trans"] if has("pattern_trans") else [])
        self.num_hidden = 0
        while has(f"linears.{2 * self.num_hidden}.weight"):
            names += [f"linears.{2 * self.num_hidden}.weight", f"linears.{2 * self.num_hidden}.bias"]
            self.num_hidden += 1
        names += ["linears1.0.weight", "linears1.0.bias", "linears1.2.weight", "linears1.2.bias"]
        self.names = names
        if not 1 <= self.num_hidden <= _lib.DSK_MAX_HIDDEN:
            raise RuntimeError(f"BlurModel: unsupported num_hidden {self.num_hidden}")
        self._bind()

    def _bind(self):
        prm, pre = self.params, self.prefix
        self.tensors = {}
        for n in self.names:
            t = _f32(prm[pre + n])
            if not t.is_cuda:
                raise RuntimeError("BlurModel parameters must be CUDA tensors")
            self.tensors[n] = t
        T = self.tensors
        p = DskParams()
        p.img_embed, p.pattern_pos = T["img_embed.img_embed"].data_ptr(), T["pattern_pos"].data_ptr()
        p.pattern_trans = T["pattern_trans"].data_ptr() if "pattern_trans" in T else None
        for l in range(self.num_hidden):
            p.lin_w[l], p.lin_b[l] = T[f"linears.{2 * l}.weight"].data_ptr(), T[f"linears.{2 * l}.bias"].data_ptr()
        p.out0_w, p.out0_b = T["linears1.0.weight"].data_ptr(), T["linears1.0.bias"].data_ptr()
        p.out1_w, p.out1_b = T["linears1.2.weight"].data_ptr(), T["linears1.2.bias"].data_ptr()
        emb, pat = T["img_embed.img_embed"], T["pattern_pos"]
        wide, in_cnl = T["linears.0.weight"].shape
        p.n_img, p.n_pt, p.embed = emb.shape[0], self.num_pt, emb.shape[1]
        p.in_embed, p.spatial_embed, p.num_hidden, p.wide = self.in_embed, self.spatial_embed, self.num_hidden, wide
        expect = (2 + 4 * self.in_embed) + emb.shape[1] + ((2 + 4 * self.spatial_embed) if self.spatial_embed > 0 else 0)
        if in_cnl != expect:
            raise RuntimeError(f"BlurModel: linears.0 takes {in_cnl} inputs but in_embed = {self.in_embed}, spatial_embed = "
                               f"{self.spatial_embed}, view embedding {emb.shape[1]} give {expect} (depth_embed / PBE features are not built)")
        if tuple(pat.shape[1:]) != (self.num_pt, 2) or pat.shape[0] not in (1, emb.shape[0]):
            raise RuntimeError(f"BlurModel: pattern_pos shape {tuple(pat.shape)} does not match num_pt = {self.num_pt}, n_img = {emb.shape[0]}")
        o0_in = T["linears1.0.weight"].shape[1]
        if o0_in not in (wide, wide + in_cnl):
            raise RuntimeError("BlurModel: unsupported linears1.0 input width")
        p.short_cut = int(o0_in != wide)
        p.isglobal = int(pat.shape[0] == 1 and emb.shape[0] != 1)
        oc = T["linears1.2.weight"].shape[0]
        if oc not in (3, 5):
            raise RuntimeError("BlurModel: linears1.2 must have 3 or 5 outputs")
        p.optim_sv_trans = int(oc == 5)
        p.kernel_hwindow = self.kernel_hwindow
        self.p = p
        self._versions = self._version()

    def _version(self):
        return tuple(int(self.params[self.prefix + n]._version) for n in self.names)

    def _refresh(self):
        # fp32 contiguous parameters are bound by pointer (in-place optimizer steps are seen as they happen); converted copies go
        # stale when the source changes: rebinding is only pointer bookkeeping, so do it on any version change
        if self._version() != self._versions:
            self._bind()

    # ---- raw entry points ------------------------------------------------------------------------------------------------------
    def _inputs(self, K, rays_info):
        rx, ry = _f32(rays_info["rays_x"].reshape(-1)), _f32(rays_info["rays_y"].reshape(-1))
        idx = rays_info["images_idx"].reshape(-1).to(torch.int64).contiguous()
        poses = _f32(rays_info["poses"][:, :3, :4])
        if not (rx.is_cuda and poses.is_cuda):
            raise RuntimeError("BlurModel inputs must be CUDA tensors")
        k4 = (float(K[0][0]), float(K[1][1]), float(K[0][2]), float(K[1][2]))
        return rx, ry, idx, poses, k4

    def run(self, H, W, k4, rx, ry, idx, poses, noise):
        N, dev = rx.shape[0], rx.device
        f32 = dict(dtype=torch.float32, device=dev)
        new_rays, weight, align = torch.empty((N, self.num_pt, 3, 2), **f32), torch.empty((N, self.num_pt), **f32), torch.zeros((1,), **f32)
        lib = _lib.load()
        ws = torch.empty((max(1, int(lib.edn_dsk_workspace_floats(C.byref(self.p), N))),), **f32)
        check(lib.edn_dsk_rays_fwd(C.byref(self.p), ptr(rx), ptr(ry), ptr(idx), ptr(poses), ptr(noise), N, int(H), int(W), *k4,
                                   ptr(new_rays), ptr(weight), ptr(align), ptr(ws), stream_ptr()), "edn_dsk_rays_fwd")
        return new_rays, weight, align

    def run_backward(self, H, W, k4, rx, ry, idx, poses, noise, d_new_rays, d_weight, d_align):
        """-> dict name -> gradient (fp32, the parameter's shape)."""
        N, dev = rx.shape[0], rx.device
        g, out = DskGrads(), {}
        for n in self.names:
            out[n] = torch.zeros_like(self.tensors[n])
        g.img_embed, g.pattern_pos = out["img_embed.img_embed"].data_ptr(), out["pattern_pos"].data_ptr()
        g.pattern_trans = out["pattern_trans"].data_ptr() if "pattern_trans" in out else None
        for l in range(self.num_hidden):
            g.lin_w[l], g.lin_b[l] = out[f"linears.{2 * l}.weight"].data_ptr(), out[f"linears.{2 * l}.bias"].data_ptr()
        g.out0_w, g.out0_b = out["linears1.0.weight"].data_ptr(), out["linears1.0.bias"].data_ptr()
        g.out1_w, g.out1_b = out["linears1.2.weight"].data_ptr(), out["linears1.2.bias"].data_ptr()
        lib = _lib.load()
        ws = torch.empty((max(1, int(lib.edn_dsk_workspace_floats(C.byref(self.p), N))),), dtype=torch.float32, device=dev)
        check(lib.edn_dsk_rays_bwd(C.byref(self.p), ptr(rx), ptr(ry), ptr(idx), ptr(poses), ptr(noise), N, int(H), int(W), *k4,
                                   ptr(_f32(d_new_rays)), ptr(_f32(d_weight)), ptr(_f32(d_align)), C.byref(g), ptr(ws), stream_ptr()),
              "edn_dsk_rays_bwd")
        return out

    # ---- reference signature (blurmodel.py:109) -----------------------------------------------------------------------------------
    def __call__(self, H, W, K, rays, rays_info, feats=None, return_img_embed=False, noise=None):
        """-> (new_rays [N, num_pt, 3, 2], weight [N, num_pt], align, extras).  `rays` is unused, as in the reference (the rays are
        rebuilt from rays_info's pixel coordinates and poses).  `noise` [N, num_pt, 2] replaces the reference's internal
        randn_like(pt_pos) * random_hwindow draw (blurmodel.py:125-127) when given."""
        if feats is not None:
            raise NotImplementedError("BlurModel: PBE features are not built")
        self._refresh()
        rx, ry, idx, poses, k4 = self._inputs(K, rays_info)
        if noise is None and self.random_hwindow > 0:
            noise = torch.randn((rx.shape[0], self.num_pt, 2), dtype=torch.float32, device=rx.device) * self.random_hwindow
        noise = _f32(noise)
        ps = [self.params[self.prefix + n] for n in self.names]
        if torch.is_grad_enabled() and any(t.requires_grad for t in ps):
            new_rays, weight, align = DskRaysFn.apply(self, int(H), int(W), k4, rx, ry, idx, poses, noise, *ps)
        else:
            new_rays, weight, align = self.run(H, W, k4, rx, ry, idx, poses, noise)
        # the view latents handed to AWP: a plain row gather of the parameter itself, so that AWP's gradient reaches it through autograd
        extras = {"img_embed": self.params[self.prefix + "img_embed.img_embed"].to(torch.float32)[idx]} if return_img_embed else {}
        return new_rays, weight, align.reshape(()), extras


class DskRaysFn(torch.autograd.Function):
    """edn_dsk_rays_fwd / edn_dsk_rays_bwd as one autograd node; the parameter tensors are inputs for graph connectivity."""

    @staticmethod
    def forward(ctx, model, H, W, k4, rx, ry, idx, poses, noise, *params):
        ctx.model, ctx.args = model, (H, W, k4, rx, ry, idx, poses, noise)
        return model.run(H, W, k4, rx, ry, idx, poses, noise)

    @staticmethod
    def backward(ctx, d_new_rays, d_weight, d_align):
        m = ctx.model
        grads = m.run_backward(*ctx.args, d_new_rays, d_weight, d_align)
        out = []
        for n in m.names:
            ref = m.params[m.prefix + n]
            out.append(grads[n].to(ref.dtype).reshape(ref.shape) if ref.requires_grad else None)
        return (None,) * 9 + tuple(out)
