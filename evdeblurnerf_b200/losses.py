"""Host mirror of the reference's loss path (SURVEY.md 8(b) "loss path"): same names / argument meaning as
networks/tonemapping.py::TonemappingTransform, utils/events.py::egm_loss, utils/metrics.py::img2mse,
VoxelNeRFBase.TV_loss_app and LLFFEventsDataset.compute_edi_prior; arithmetic runs in libevdeblur_b200.so."""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import CRF_GAMMA, CRF_LEARN, CRF_LUMA, CRF_LUMA_AVG, CRF_LUMA_REC709, CRF_SKIP_LEARN, CrfParams, check, ptr, stream_ptr


def _f32c(t):
    return t.detach().to(torch.float32).contiguous()


class TonemappingTransform:
    """networks/tonemapping.py:96-154.  `params`: the reference state_dict of the module (keys
    `tonemapping_event.linear.{0,2,4,6}.{weight,bias}` / `tonemapping_rgb...`) on the CUDA device."""

    def __init__(self, params, map_type_rgb="gamma", map_type_event="learn", extra_features_event=2,
                 extra_features_rgb=0, gamma=2.2, luma_standard="rec601"):
        if luma_standard not in ("rec601", "rec709", "avg"):
            raise ValueError(f"Unknown luma_standard {luma_standard}")          # tonemapping.py:134-135
        self.luma_standard = luma_standard
        self.map_type = {"rgb": map_type_rgb, "event": map_type_event}
        self.extra = {"rgb": extra_features_rgb, "event": extra_features_event}
        self.gamma = float(gamma)
        self.keep = []
        self.cp = {}
        for which in ("rgb", "event"):
            cp = CrfParams()
            cp.extra_features, cp.gamma = self.extra[which], self.gamma
            if self.map_type[which] == "learn":
                pre = f"tonemapping_{which}.linear."
                for slot, idx in (("0", 0), ("1", 2), ("2", 4), ("3", 6)):
                    w, b = _f32c(params[pre + f"{idx}.weight"]), _f32c(params[pre + f"{idx}.bias"])
                    if not w.is_cuda:
                        raise RuntimeError("TonemappingTransform parameters must be CUDA tensors")
                    self.keep += [w, b]
                    setattr(cp, "w" + slot, w.data_ptr())
                    setattr(cp, "b" + slot, b.data_ptr())
            self.cp[which] = cp
        self.params = params

    def _weights(self, which):
        pre = f"tonemapping_{which}.linear."
        return [self.params[pre + f"{idx}.{kind}"] for idx in (0, 2, 4, 6) for kind in ("weight", "bias")]

    def _run(self, which, x, feat, skip_learn, luma):
        mt = self.map_type[which]
        if mt not in ("none", "gamma", "learn"):
            raise NotImplementedError(f"map_type {mt!r}")
        flags = (CRF_GAMMA if "gamma" in mt else 0) | (CRF_LEARN if mt == "learn" else 0)
        luma_bits = (CRF_LUMA | {"rec601": 0, "rec709": CRF_LUMA_REC709, "avg": CRF_LUMA_AVG}[self.luma_standard]) if luma else 0
        flags |= (CRF_SKIP_LEARN if skip_learn else 0) | luma_bits
        if mt == "none":
            flags = luma_bits
        shape = x.shape
        xf = _f32c(x).reshape(-1, 3)
        M = xf.shape[0]
        per_channel, f = 0, None
        if feat is not None and self.extra[which] > 0 and mt == "learn":
            f = _f32c(feat)
            per_channel = 1 if f.ndim == 3 else 0
        ws = self._weights(which) if mt == "learn" else []
        if torch.is_grad_enabled() and (x.requires_grad or any(w.requires_grad for w in ws)):
            from .autograd import CrfFn
            xin = x.to(torch.float32).reshape(-1, 3).contiguous()       # keeps the graph (no detach)
            out = CrfFn.apply(xin, f, per_channel, flags, self.cp[which], *ws)
            return out.reshape(*shape[:-1], 1 if luma else 3)
        out = torch.empty((M, 1 if luma else 3), dtype=torch.float32, device=xf.device)
        check(_lib.load().edn_crf_fwd(C.byref(self.cp[which]), ptr(xf), ptr(f), per_channel, flags, M, ptr(out), stream_ptr()),
              "edn_crf_fwd")
        return out.reshape(*shape[:-1], 1 if luma else 3)

    def encode_rgb(self, x, skip_learn_crf=False, rgb_extra_feat=None, **kwargs):
        assert x.shape[-1] == 3
        return self._run("rgb", x, rgb_extra_feat, skip_learn_crf, False)

    def encode_luma(self, x, keep_rgb=False, tonemap_only=False, skip_learn_crf=False, ev_extra_feat=None, **kwargs):
        y = self._run("event", x, ev_extra_feat, skip_learn_crf, not tonemap_only)
        if keep_rgb and not tonemap_only:
            y = y.expand(*y.shape[:-1], 3).contiguous()
        return y

    def forward(self, x, mode="encode", chunk=None, **kwargs):
        # the reference chunks to bound memory (tonemapping.py:141-154); the fused kernel has no intermediates
        if mode == "encode_rgb":
            return self.encode_rgb(x, **kwargs)
        if mode == "encode_luma":
            return self.encode_luma(x, **kwargs)
        raise RuntimeError(f"mode '{mode}' not recognized")

    __call__ = forward


def egm_loss(luma_start, luma_end, bii, color_mask=None, color_weight=None, log_eps=1e-5):
    """utils/events.py:260-284."""
    ls, le, b = _f32c(luma_start), _f32c(luma_end), _f32c(bii).reshape(-1)
    M = b.shape[0]
    ls, le = ls.reshape(M, -1), le.reshape(M, -1)
    Cn = ls.shape[1]
    mask = cw = None
    if color_mask is not None:
        assert color_mask.shape == (M, 3)
        mask = color_mask.to(torch.uint8).contiguous()
        if color_weight is not None:
            cw = torch.as_tensor(color_weight, dtype=torch.float32, device=ls.device).contiguous()
    elif Cn != 1:
        raise RuntimeError("egm_loss: 3-channel luma needs a color_mask")
    if torch.is_grad_enabled() and (luma_start.requires_grad or luma_end.requires_grad):
        from .autograd import EgmLossFn
        return EgmLossFn.apply(luma_start.to(torch.float32).reshape(M, -1).contiguous(),
                               luma_end.to(torch.float32).reshape(M, -1).contiguous(), b, mask, cw, float(log_eps))
    out = torch.empty((1,), dtype=torch.float32, device=ls.device)
    check(_lib.load().edn_egm_loss_fwd(ptr(ls), ptr(le), ptr(b), ptr(mask), ptr(cw), Cn, M, float(log_eps), ptr(out), stream_ptr()),
          "edn_egm_loss_fwd")
    return out[0]


def img2mse(x, y):
    """utils/metrics.py:7."""
    xf, yf = _f32c(x), _f32c(y)
    assert xf.shape == yf.shape
    if torch.is_grad_enabled() and x.requires_grad:
        from .autograd import Img2MseFn
        return Img2MseFn.apply(x.to(torch.float32).contiguous(), yf)
    out = torch.empty((1,), dtype=torch.float32, device=xf.device)
    check(_lib.load().edn_img2mse(ptr(xf), ptr(yf), xf.numel(), ptr(out), stream_ptr()), "edn_img2mse")
    return out[0]


def mse2psnr(x):
    """utils/metrics.py:8: -10 log10(mse) (a scalar; host-side tensor op)."""
    return -10.0 * torch.log(x) / 2.302585092994046


def tv_loss_app(params, prefix):
    """VoxelNeRFBase.TV_loss_app (voxnerf.py:126-130) on the reference-layout parameters `prefix + app_plane.i / app_line.i`."""
    raw = [params[prefix + f"app_plane.{i}"] for i in range(3)] + [params[prefix + f"app_line.{i}"] for i in range(3)]
    if torch.is_grad_enabled() and any(t.requires_grad for t in raw):
        from .autograd import TvLossFn
        return TvLossFn.apply(*raw)
    planes = [_f32c(params[prefix + f"app_plane.{i}"]) for i in range(3)]
    lines = [_f32c(params[prefix + f"app_line.{i}"]) for i in range(3)]
    dev = planes[0].device
    pp = (C.c_void_p * 3)(*[t.data_ptr() for t in planes])
    lp = (C.c_void_p * 3)(*[t.data_ptr() for t in lines])
    ph = (C.c_int32 * 3)(*[t.shape[2] for t in planes])
    pw = (C.c_int32 * 3)(*[t.shape[3] for t in planes])
    ll = (C.c_int32 * 3)(*[t.shape[2] for t in lines])
    nc = (C.c_int32 * 3)(*[t.shape[1] for t in planes])
    ws = torch.empty((12,), dtype=torch.float64, device=dev)
    out = torch.empty((1,), dtype=torch.float32, device=dev)
    check(_lib.load().edn_tv_loss_app(C.byref(pp), C.byref(lp), C.byref(ph), C.byref(pw), C.byref(ll), C.byref(nc), ptr(ws),
                                      ptr(out), stream_ptr()), "edn_tv_loss_app")
    return out[0]


def edi_prior_image(ev_x, ev_y, ev_t, ev_p, blurry, t_start, t_end, c_pos, c_neg, steps=9, device="cuda"):
    """compute_edi_prior for one frame (data/loader_events.py:99-131, utils/edi.py): `steps` time stamps ->
    steps - 1 sub-intervals [searchsorted(t_j, left), searchsorted(t_{j+1}, right)) (inclusive both ends).
    ev_* are host numpy arrays (the event stream lives on the host, as in the reference); blurry [H,W,C] numpy or tensor.
    Returns (sharp [H,W,C], bii [steps-1,H,W]) CUDA tensors."""
    ts = np.linspace(t_start, t_end, steps)
    i0 = np.searchsorted(ev_t, ts[:-1], side="left")
    i1 = np.searchsorted(ev_t, ts[1:], side="right")
    lo, hi = int(i0.min()), int(i1.max())
    x = torch.as_tensor(np.ascontiguousarray(ev_x[lo:hi]), dtype=torch.float32).to(device)
    y = torch.as_tensor(np.ascontiguousarray(ev_y[lo:hi]), dtype=torch.float32).to(device)
    p = torch.as_tensor(np.ascontiguousarray(ev_p[lo:hi]), dtype=torch.float32).to(device)
    s0 = torch.as_tensor(i0 - lo, dtype=torch.int64).to(device)
    s1 = torch.as_tensor(i1 - lo, dtype=torch.int64).to(device)
    bl = torch.as_tensor(blurry, dtype=torch.float32).to(device).contiguous()
    if bl.ndim == 2:
        bl = bl[..., None]
    H, W, Cn = bl.shape
    n_seg = steps - 1
    bii = torch.empty((n_seg, H, W), dtype=torch.float32, device=device)
    sharp = torch.empty((H, W, Cn), dtype=torch.float32, device=device)
    if x.numel() == 0:        # no events in the exposure: keep valid pointers for the (empty) splat
        x = y = p = torch.zeros((1,), dtype=torch.float32, device=device)
    check(_lib.load().edn_edi_prior(ptr(x), ptr(y), ptr(p), ptr(s0), ptr(s1), n_seg, int((i1 - i0).max()), ptr(bl), H, W, Cn,
                                    float(c_pos), float(c_neg), ptr(bii), ptr(sharp), stream_ptr()), "edn_edi_prior")
    return sharp, bii
