"""Host side of the render path's backward pass (what `loss.backward()` does in the reference, run_nerf.py:594, through
NeRFAll.render_rays): gradient buffers in the reference's parameter layout and the two field-backward launches.

`RenderGradients` owns one gradient tensor per reference parameter of the two PDRF fields (state_dict names, SURVEY
Appendix A).  MLP / basis_mat gradients are accumulated directly in the nn.Linear layout; VM plane / line gradients are
scatter-added in the render layout (channel-last fp32) and moved into the [1,C,H,W] layout by `finish()`.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import FieldWeights, VmGridGrad, check, ptr, stream_ptr

_LINEAR = ("sigma_net.0.weight", "sigma_net.1.weight", "color_net.0.weight", "color_net.1.weight", "color_net.2.weight")
_BIAS = ("color_net.0.bias", "color_net.1.bias", "color_net.2.bias")


def _field_struct(T, pre, basis_pres):
    fw = FieldWeights()
    for g, bp in enumerate(basis_pres):
        fw.basis[g] = T[bp + "basis_mat.weight"].data_ptr()
    fw.sigma0, fw.sigma1 = T[pre + _LINEAR[0]].data_ptr(), T[pre + _LINEAR[1]].data_ptr()
    fw.color0, fw.color1, fw.color2 = (T[pre + n].data_ptr() for n in _LINEAR[2:])
    fw.color0_b, fw.color1_b, fw.color2_b = (ptr(T.get(pre + n)) for n in _BIAS)
    fw.hidden = T[pre + _LINEAR[0]].shape[0]
    fw.geo_feat = T[pre + _LINEAR[1]].shape[0] - 1
    fw.n_grids = len(basis_pres)
    return fw


class RenderGradients:
    """Gradient accumulators for a c2f RenderEngine (see module docstring)."""

    def __init__(self, engine):
        self.engine = engine
        P = engine.params
        self.prefixes = ["mlp_coarse."] + (["mlp_fine."] if engine.fine is not None else [])
        self.w = {}
        self.hwc = {}
        for pre in self.prefixes:
            for n in _LINEAR + _BIAS + ("basis_mat.weight",):
                if pre + n in P:
                    self.w[pre + n] = torch.zeros_like(P[pre + n], dtype=torch.float32)
            for i in range(3):
                for kind in ("app_plane", "app_line"):
                    _, Cc, Hh, Ww = P[pre + f"{kind}.{i}"].shape
                    self.hwc[pre + f"{kind}.{i}"] = torch.zeros((Hh, Ww, Cc), dtype=torch.float32, device=engine.device)
        self.grid_struct = {}
        for pre in self.prefixes:
            g = VmGridGrad()
            for i in range(3):
                g.plane[i] = self.hwc[pre + f"app_plane.{i}"].data_ptr()
                g.line[i] = self.hwc[pre + f"app_line.{i}"].data_ptr()
            self.grid_struct[pre] = g

    def zero_(self):
        for t in list(self.w.values()) + list(self.hwc.values()):
            t.zero_()

    def finish(self):
        """-> {state_dict name: gradient in the reference layout}."""
        lib = _lib.load()
        out = dict(self.w)
        for name, g in self.hwc.items():
            Hh, Ww, Cc = g.shape
            dst = torch.empty((1, Cc, Hh, Ww), dtype=torch.float32, device=g.device)
            check(lib.edn_unpack_vm_plane_grad(ptr(g), ptr(dst), Cc, Hh, Ww, 0, stream_ptr()), "edn_unpack_vm_plane_grad")
            out[name] = dst
        return out


def field_backward(engine, grads, fine, ray_batch, z_vals, noise, d_rgb, d_depth, d_acc, d_ray_batch, d_weights=None,
                   d_feat=None, chunk_rays=2048, merge=None):
    """Accumulates the gradients of one field's render pass (edn_render_field_bwd) into `grads` and `d_ray_batch`."""
    lib = _lib.load()
    P = engine.params
    pre = "mlp_fine." if fine else "mlp_coarse."
    basis_pres = ["mlp_coarse.", "mlp_fine."] if fine else ["mlp_coarse."]
    w = _field_struct(P, pre, basis_pres)
    gw = _field_struct(grads.w, pre, basis_pres)
    R, S = z_vals.shape
    if R == 0:
        return
    chunk = min(R, int(chunk_rays))
    nbytes = int(lib.edn_field_bwd_workspace_bytes(w.n_grids, w.hidden, w.geo_feat, chunk, S))
    ws = engine.workspace(nbytes)

    def f(t):
        return None if t is None else t.float().contiguous()
    keep = [f(t) for t in (ray_batch, z_vals, noise, d_rgb, d_depth, d_acc, d_weights, d_feat)]
    g0 = engine.coarse.grid
    g1 = engine.fine.grid if fine else None
    mg = None
    if merge is not None:          # (order [R][S] int64 | None, n_coarse, moved buffer): see edn_field_bwd_merge
        mg = _lib.FieldBwdMerge(ptr(merge[0]) if fine else None, int(merge[1]), merge[2].data_ptr())
    check(engine._launch("field_bwd_fine" if fine else "field_bwd_coarse", lambda: lib.edn_render_field_bwd(
        C.byref(g0), C.byref(g1) if g1 is not None else None, C.byref(w), ptr(keep[0]), ptr(keep[1]), ptr(keep[2]), R, S,
        engine.prec_code, ptr(keep[3]), ptr(keep[4]), ptr(keep[5]), ptr(keep[6]), ptr(keep[7]), C.byref(gw),
        C.byref(grads.grid_struct["mlp_coarse."]), C.byref(grads.grid_struct["mlp_fine."]) if fine else None,
        ptr(d_ray_batch), ptr(ws), nbytes, C.byref(mg) if mg is not None else None, stream_ptr())), "edn_render_field_bwd")


def render_rays_backward(engine, saved, d_out, grads=None, chunk_rays=2048):
    """Backward of RenderEngine.render_rays (mode = c2f).

    saved: dict(ray_batch, z_vals0, noise0 [, z_vals, noise1]) from the forward call (`retraw=True` gives the depths);
    d_out: upstream gradients keyed like the forward result: rgb_map / depth_map / acc_map and, for N_importance > 0,
    rgb0 / depth0 / acc0 (missing or None = zero).
    -> (RenderGradients, d_ray_batch [R, 11])."""
    if grads is None:
        grads = RenderGradients(engine)
    rb = saved["ray_batch"].float().contiguous()
    d_rb = torch.zeros_like(rb)
    two_stage = saved.get("z_vals") is not None and engine.fine is not None and saved["z_vals"].shape[1] != saved["z_vals0"].shape[1]
    if two_stage:
        # the coarse positions inside the merged depths are scattered into the coarse grid ONCE: the fine call hands their d P rows to
        # the coarse call (needs the merged order of the forward; without it both calls scatter their own rows)
        merge = None
        order = saved.get("order")
        if order is not None:
            R, Nc = saved["z_vals0"].shape
            from ._lib import EDN_F32
            moved = torch.empty((R * Nc * 96,), dtype=torch.float32 if engine.prec_code == EDN_F32 else torch.bfloat16, device=rb.device)
            merge = (order.contiguous(), Nc, moved)
        field_backward(engine, grads, True, rb, saved["z_vals"], saved.get("noise1"), d_out.get("rgb_map"), d_out.get("depth_map"),
                       d_out.get("acc_map"), d_rb, d_feat=d_out.get("depth_feature"), chunk_rays=chunk_rays, merge=merge)
        field_backward(engine, grads, False, rb, saved["z_vals0"], saved.get("noise0"), d_out.get("rgb0"), d_out.get("depth0"),
                       d_out.get("acc0"), d_rb, chunk_rays=chunk_rays, merge=merge)
    else:
        field_backward(engine, grads, False, rb, saved["z_vals0"], saved.get("noise0"), d_out.get("rgb_map"), d_out.get("depth_map"),
                       d_out.get("acc_map"), d_rb, d_feat=d_out.get("depth_feature"), chunk_rays=chunk_rays)
    return grads, d_rb
