"""Host side of the fused render path: parameter packing into the render layout and the `render_rays` surface.

`RenderEngine.render_rays` mirrors `NeRFAll.render_rays` (reference networks/renderer.py:129-264, mode = c2f): same
argument names and meaning, same result dict.  PyTorch is used for device memory and streams only; all arithmetic
runs in libevdeblur_b200.so through the C ABI (include/evdeblur_b200.h).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import EDN_BF16, EDN_F32, EDN_TC32, FLAG_LINDISP, FLAG_RELU_RGB, FLAG_TRAIN, FieldMlp, VmGrid, check, ptr, stream_ptr

MATMODE = ((0, 1), (0, 2), (1, 2))   # voxnerf.py:99
VECMODE = (2, 1, 0)                  # voxnerf.py:100
APP_N_COMP = (64, 16, 16)


def _require_cuda(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError(f"evdeblurnerf_b200: `{name}` must be a CUDA tensor (there is no CPU path)")


class PackedField:
    """Render-layout copy of one PDRF field (VoxelNeRFBase, voxnerf.py:6-118): channel-last VM planes/lines (fp32 or
    bf16), transposed / padded MLP weights, and the ctypes structs that point at them."""

    def __init__(self, P, prefix, aabb_min, aabb_max, coarse, grid_dtype):
        lib = _lib.load()
        self.coarse = coarse
        self.keep = []          # owns the device buffers the structs point at
        g = VmGrid()
        tdt = torch.float32 if grid_dtype == EDN_F32 else torch.bfloat16
        for i in range(3):
            pl = P[prefix + f"app_plane.{i}"]
            ln = P[prefix + f"app_line.{i}"]
            _require_cuda(pl, prefix + f"app_plane.{i}")
            _, Cc, H, W = pl.shape
            if Cc != APP_N_COMP[i] or ln.shape[1] != Cc or ln.shape[3] != 1:
                raise RuntimeError(f"unsupported VM component shape {tuple(pl.shape)} / {tuple(ln.shape)}")
            src = pl.detach().to(torch.float32).contiguous()
            dst = torch.empty((H, W, Cc), dtype=tdt, device=pl.device)
            check(lib.edn_pack_vm_plane(ptr(src), ptr(dst), Cc, H, W, grid_dtype, stream_ptr()), "edn_pack_vm_plane")
            Ln = ln.shape[2]
            lsrc = ln.detach().to(torch.float32).contiguous()
            ldst = torch.empty((Ln, Cc), dtype=tdt, device=pl.device)
            check(lib.edn_pack_vm_plane(ptr(lsrc), ptr(ldst), Cc, Ln, 1, grid_dtype, stream_ptr()), "edn_pack_vm_plane")
            self.keep += [src, dst, lsrc, ldst]
            g.plane[i], g.line[i] = dst.data_ptr(), ldst.data_ptr()
            g.plane_h[i], g.plane_w[i], g.line_len[i], g.n_comp[i] = H, W, Ln, Cc
        g.dtype = grid_dtype
        basis_t = P[prefix + "basis_mat.weight"].detach().float().t().contiguous()
        if tuple(basis_t.shape) != (96, 32):
            raise RuntimeError(f"unsupported basis_mat shape {tuple(basis_t.shape)}")
        self.keep.append(basis_t)
        g.basis_t = basis_t.data_ptr()
        for i in range(3):
            g.aabb_min[i], g.aabb_max[i] = float(aabb_min[i]), float(aabb_max[i])
        self.grid = g

        def wt(name, pad_rows=0, pad_cols=0):
            w = P[prefix + name].detach().float().t()
            if pad_rows or pad_cols:
                w = torch.nn.functional.pad(w, (0, pad_cols, 0, pad_rows))
            w = w.contiguous()
            self.keep.append(w)
            return w

        def bias(name, pad=0):
            b = P.get(prefix + name)
            if b is None:
                return None
            b = torch.nn.functional.pad(b.detach().float(), (0, pad)).contiguous()
            self.keep.append(b)
            return b

        m = FieldMlp()
        if coarse:
            exp = {"sigma_net.0.weight": (64, 95), "sigma_net.1.weight": (16, 64), "color_net.0.weight": (64, 42),
                   "color_net.1.weight": (64, 64), "color_net.2.weight": (3, 64)}
        else:
            exp = {"sigma_net.0.weight": (256, 127), "sigma_net.1.weight": (129, 256), "color_net.0.weight": (256, 155),
                   "color_net.1.weight": (256, 256), "color_net.2.weight": (3, 256)}
        for k, shp in exp.items():
            if tuple(P[prefix + k].shape) != shp:
                raise RuntimeError(f"unsupported shape for {prefix + k}: {tuple(P[prefix + k].shape)} (expected {shp})")
        if prefix + "sigma_net.2.weight" in P or prefix + "color_net.3.weight" in P:
            raise RuntimeError("unsupported MLP depth (expected sigma_net: 2 layers, color_net: 3 layers)")
        m.sigma0_t = wt("sigma_net.0.weight", pad_rows=1).data_ptr()
        if coarse:
            m.sigma1_t = wt("sigma_net.1.weight").data_ptr()
            m.sigma1_v = None
            m.hidden, m.geo_feat = 64, 15
        else:
            w1 = P[prefix + "sigma_net.1.weight"].detach().float()
            v = w1[0].contiguous()
            t = w1[1:].t().contiguous()
            self.keep += [v, t]
            m.sigma1_v, m.sigma1_t = v.data_ptr(), t.data_ptr()
            m.hidden, m.geo_feat = 256, 128
        m.color0_t = wt("color_net.0.weight").data_ptr()
        m.color1_t = wt("color_net.1.weight").data_ptr()
        m.color2_t = wt("color_net.2.weight", pad_cols=1).data_ptr()
        b0, b1, b2 = bias("color_net.0.bias"), bias("color_net.1.bias"), bias("color_net.2.bias", pad=1)
        m.color0_b = ptr(b0)
        m.color1_b = ptr(b1)
        m.color2_b = ptr(b2)
        m.tc_blob = None
        self.mlp = m
        self.basis_t = basis_t

    def pack_tensor_core_operands(self, coarse_field=None):
        """bf16 UMMA operand blob (edn_pack_fine_tc: fine field + both basis_mat's; edn_pack_coarse_tc: coarse field)."""
        lib = _lib.load()
        if self.coarse:
            n = int(lib.edn_coarse_tc_blob_bytes())
            blob = torch.empty((n,), dtype=torch.uint8, device=self.basis_t.device)
            ws = torch.empty((int(lib.edn_coarse_tc_pack_workspace_floats()),), dtype=torch.float32, device=self.basis_t.device)
            check(lib.edn_pack_coarse_tc(C.byref(self.mlp), ptr(self.basis_t), ptr(ws), ptr(blob), stream_ptr()), "edn_pack_coarse_tc")
        else:
            n = int(lib.edn_fine_tc_blob_bytes())
            blob = torch.empty((n,), dtype=torch.uint8, device=self.basis_t.device)
            ws = torch.empty((int(lib.edn_fine_tc_pack_workspace_floats()),), dtype=torch.float32, device=self.basis_t.device)
            check(lib.edn_pack_fine_tc(C.byref(self.mlp), ptr(coarse_field.basis_t), ptr(self.basis_t), ptr(ws), ptr(blob),
                                       stream_ptr()), "edn_pack_fine_tc")
        self.keep.append(blob)
        self.mlp.tc_blob = blob.data_ptr()


class SamplerHost:
    """Shared host pieces of the c2f and nerf-mode renderers: cached linspace vectors, profiling hooks, the sampler."""

    def seed(self, seed):
        """Seed of the in-library Philox generator used for perturb / raw_noise_std draws; every render call advances a
        call counter, so draws differ between calls and are reproducible for a given (seed, call order)."""
        self._seed, self._calls = int(seed) & ((1 << 64) - 1), 0

    def _random(self, shape, stream_id, normal=False, scale=1.0):
        out = torch.empty(shape, dtype=torch.float32, device=self.device)
        sid = (self._calls * 8 + stream_id) & 0xFFFFFFFF
        check(_lib.load().edn_fill_random(ptr(out), out.numel(), self._seed, sid, 1 if normal else 0, float(scale), stream_ptr()),
              "edn_fill_random")
        return out

    def _init_host(self, device):
        self.seed(0)
        self.device = torch.device(device if device is not None else "cuda")
        self._lin = {}
        self.profile = None     # set to {} to collect (start, end) CUDA events per kernel (bench.py roofline leg)
        # NaN / Inf guard (renderer.py:259-263): True = every returned tensor except depth_feature (a non-finite feature reaches
        # rgb_map through color_net anyway), "full" = depth_feature too (1.3 GB scan on the headline batch), False = off
        self.check_numerics = True
        self._err_flags = torch.zeros((2,), dtype=torch.int32, device=self.device)

    GUARD_KEYS = ("rgb_map", "depth_map", "acc_map", "z_vals", "weights", "rgb0", "depth0", "acc0", "z_std", "z_vals0", "weights0",
                  "depth_feature")

    def _guard(self, ret):
        """One launch ORs a bit per result tensor into the device flag words; nothing is read back here."""
        if not self.check_numerics:
            return ret
        keys = [k for k in self.GUARD_KEYS if k in ret and ret[k] is not None and ret[k].numel() > 0
                and (k != "depth_feature" or self.check_numerics == "full")]
        ts = [ret[k] if ret[k].is_contiguous() else ret[k].contiguous() for k in keys]
        ptrs = (C.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        sizes = (C.c_int64 * len(ts))(*[t.numel() for t in ts])
        # bit t of the flag words = position of the key in GUARD_KEYS: pad the call with empty slots to keep positions fixed
        idx = [self.GUARD_KEYS.index(k) for k in keys]
        n = (max(idx) + 1) if idx else 0
        full_p, full_n = (C.c_void_p * max(n, 1))(), (C.c_int64 * max(n, 1))()
        for j, i in enumerate(idx):
            full_p[i], full_n[i] = ptrs[j], sizes[j]
        check(_lib.load().edn_check_finite(full_p, full_n, n, ptr(self._err_flags), stream_ptr()), "edn_check_finite")
        return ret

    def numerical_errors(self, reset=True):
        """Reads the guard's flag words (synchronises) -> the reference's messages, e.g. ['rgb_map contains nan.']."""
        nan_bits, inf_bits = (int(v) & 0xFFFFFFFF for v in self._err_flags.tolist())
        if reset and (nan_bits or inf_bits):
            self._err_flags.zero_()
        msgs = [f"{k} contains nan." for i, k in enumerate(self.GUARD_KEYS) if nan_bits >> i & 1]
        return msgs + [f"{k} contains inf." for i, k in enumerate(self.GUARD_KEYS) if inf_bits >> i & 1]


class RenderEngine(SamplerHost):
    """Fused c2f renderer over a reference `NeRFAll.state_dict()`-style parameter dict.

    precision: "fp32" -> fp32 SIMT kernels everywhere (parity mode, 1e-4 rel against the reference);
               "tc32" -> tensor-core parity mode: the GEMMs of both render passes on tcgen05 with bf16 x 3 split operands and
                         fp32 TMEM accumulation (fp32-grade, same tolerances as "fp32"); fp32 VM planes, fp32 backward;
               "bf16" -> tcgen05 tensor-core fine pass with bf16 operands / fp32 accumulation and bf16 VM planes.
    """

    def __init__(self, params, aabb_min, aabb_max, precision="fp32", rmnearplane=0, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError("evdeblurnerf_b200.RenderEngine needs a CUDA device (no CPU fallback)")
        _lib.load()
        if precision not in ("fp32", "bf16", "tc32"):
            raise ValueError(f"precision must be 'fp32', 'tc32' or 'bf16', got {precision!r}")
        self.precision = precision
        self.prec_code = EDN_BF16 if precision == "bf16" else EDN_F32      # storage / backward precision
        self._init_host(device)
        self.rmnearplane = float(rmnearplane)
        self.aabb_min, self.aabb_max = [float(x) for x in aabb_min], [float(x) for x in aabb_max]
        self.repack(params)

    def workspace(self, nbytes):
        """Cached scratch buffer (grown on demand) for the backward pass's recomputed activations."""
        ws = getattr(self, "_workspace", None)
        if ws is None or ws.numel() < nbytes:
            self._workspace = ws = torch.empty((int(nbytes),), dtype=torch.uint8, device=self.device)
        return ws

    def backward(self, saved, d_out, chunk_rays=8192, white_bkgd=False):
        """Backward of render_rays -> ({state_dict name: gradient in the reference layout}, d_ray_batch [R, 11])."""
        from .backward import RenderGradients, render_rays_backward
        grads, d_rb = render_rays_backward(self, saved, d_out, grads=RenderGradients(self), chunk_rays=chunk_rays)
        return grads.finish(), d_rb

    def _launch(self, name, fn):
        if self.profile is None:
            return fn()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn()
        e.record()
        self.profile.setdefault(name, []).append((s, e))
        return rc

    def repack(self, params):
        """(Re)build the render-layout copies; call after the parameters changed (optimizer step / checkpoint load)."""
        P = {k: (v if v.is_cuda else v.to(self.device)) for k, v in params.items() if isinstance(v, torch.Tensor)}
        # reference-layout fp32 views of the two fields' parameters: what the backward pass reads (backward.py)
        self.params = {k: v.detach().float().contiguous() for k, v in P.items() if k.startswith(("mlp_coarse.", "mlp_fine."))}
        grid_dtype = self.prec_code
        self.coarse = PackedField(P, "mlp_coarse.", self.aabb_min, self.aabb_max, True, grid_dtype)
        if self.prec_code == EDN_BF16 or self.precision == "tc32":
            self.coarse.pack_tensor_core_operands()
        self.fine = None
        if "mlp_fine.sigma_net.0.weight" in P:
            self.fine = PackedField(P, "mlp_fine.", self.aabb_min, self.aabb_max, False, grid_dtype)
            if self.prec_code == EDN_BF16 or self.precision == "tc32":
                self.fine.pack_tensor_core_operands(self.coarse)

    @staticmethod
    def _empty_result(dev, Nc, Ni, retraw, want_feat):
        f32 = dict(dtype=torch.float32, device=dev)
        S = Nc + Ni
        ret = {"rgb_map": torch.empty((0, 3), **f32), "depth_map": torch.empty((0,), **f32), "acc_map": torch.empty((0,), **f32)}
        if retraw:
            ret["z_vals"], ret["weights"] = torch.empty((0, S), **f32), torch.empty((0, S), **f32)
        if Ni > 0:
            ret.update(rgb0=torch.empty((0, 3), **f32), depth0=torch.empty((0,), **f32), acc0=torch.empty((0,), **f32),
                       z_std=torch.empty((0,), **f32))
            if retraw:
                ret["z_vals0"], ret["weights0"] = torch.empty((0, Nc), **f32), torch.empty((0, Nc), **f32)
        if want_feat:
            ret["depth_feature"] = torch.empty((0, S, 128 if Ni > 0 else 15), **f32)
            ret["z_vals"] = torch.empty((0, S), **f32)
        return ret

    def _fine_precision(self, n_total, want_feat=False):
        # the bf16 tensor-core fine kernel maps a ray to ceil(S / 128) UMMA M tiles (e.g. 96 + 96 samples = 128 + 64 rows); the
        # bf16 x 3 parity kernel handles rays of <= 128 samples without depth_feature, everything else stays on the fp32 SIMT kernel
        if self.precision == "tc32":
            return EDN_TC32 if (n_total <= 128 and not want_feat) else EDN_F32
        return self.prec_code

    def _coarse_precision(self, n_samples, want_feat=False):
        # the tensor-core coarse kernels tile 128 rows = floor(128 / n_samples) rays; other sample counts use the SIMT kernel, and so
        # does the parity mode when the coarse feature_map is requested (the bf16 x 3 kernel runs the folded schedule)
        if self.precision == "tc32":
            return EDN_TC32 if (32 <= n_samples <= 128 and not want_feat) else EDN_F32
        return EDN_BF16 if (self.prec_code == EDN_BF16 and 32 <= n_samples <= 128) else EDN_F32

    def _linspace(self, n):
        # computed by torch on the CPU (bit-identical to the reference's torch.linspace), cached on the device
        if n not in self._lin:
            self._lin[n] = torch.linspace(0., 1., steps=n).to(self.device)
        return self._lin[n]

    def vm_sample(self, pts, field="coarse"):
        """VoxelNeRFBase.sample (voxnerf.py:203): pts [..., 3] -> [..., 32]."""
        _require_cuda(pts, "pts")
        f = self.coarse if field == "coarse" else self.fine
        p = pts.reshape(-1, 3).float().contiguous()
        out = torch.empty((p.shape[0], 32), dtype=torch.float32, device=p.device)
        check(_lib.load().edn_vm_sample(C.byref(f.grid), ptr(p), ptr(out), p.shape[0], stream_ptr()), "edn_vm_sample")
        return out.reshape(*pts.shape[:-1], 32)

    def sample_pdf_merge(self, z_vals0, weights0, n_importance, u=None, want_indices=True):
        """sample_pdf + sort (utils/rays.py:149, renderer.py:199-205) -> dict(z_samples, inds, z_vals, order, z_std)."""
        _require_cuda(z_vals0, "z_vals0")
        R, Nc = z_vals0.shape
        dev = z_vals0.device
        z0, w0 = z_vals0.float().contiguous(), weights0.float().contiguous()
        zs = torch.empty((R, n_importance), dtype=torch.float32, device=dev)
        zv = torch.empty((R, Nc + n_importance), dtype=torch.float32, device=dev)
        zstd = torch.empty((R,), dtype=torch.float32, device=dev)
        inds = torch.empty((R, n_importance), dtype=torch.int64, device=dev) if want_indices else None
        order = torch.empty((R, Nc + n_importance), dtype=torch.int64, device=dev) if want_indices else None
        u_det = self._linspace(n_importance) if u is None else None
        u_rand = None if u is None else u.float().contiguous()
        check(self._launch("sample_pdf", lambda: _lib.load().edn_sample_pdf_merge(
            ptr(z0), ptr(w0), ptr(u_det), ptr(u_rand), R, Nc, n_importance, ptr(zs), ptr(inds), ptr(zv), ptr(order),
            ptr(zstd), stream_ptr())), "edn_sample_pdf_merge")
        return {"z_samples": zs, "inds": inds, "z_vals": zv, "order": order, "z_std": zstd}

    def render_rays(self, ray_batch, N_samples, retraw=False, lindisp=False, perturb=0., N_importance=0,
                    white_bkgd=False, raw_noise_std=0., pytest=False, force_naive=False, inference=False,
                    is_train=True, use_awp=False, rand=None, want_indices=False):
        """Drop-in for NeRFAll.render_rays (renderer.py:129-264), mode = c2f.

        Extra keyword arguments (not in the reference): `is_train` (= module.training), `use_awp` (= self.use_awp),
        `rand` = dict of injected random tensors {t_rand [R,Nc], u [R,Ni], noise0 [R,Nc-1], noise1 [R,Nc+Ni-1]} (noise
        already scaled by raw_noise_std) -- when absent and perturb / raw_noise_std are non-zero they are drawn by the
        library's counter-based Philox generator (edn_fill_random; seed with `engine.seed(s)`), one stream per draw site
        (renderer.py:176, voxnerf.py:175, rays.py:162, voxnerf.py:175).
        """
        _require_cuda(ray_batch, "ray_batch")
        if ray_batch.shape[-1] != 11:
            raise RuntimeError("render_rays: ray_batch must be [R, 11] (use_viewdirs=True), got %r" % (tuple(ray_batch.shape),))
        lib = _lib.load()
        rb = ray_batch.float().contiguous()
        R, dev = rb.shape[0], rb.device
        if R == 0:
            return self._empty_result(dev, int(N_samples), int(N_importance), retraw, use_awp and not force_naive and not inference)
        rand = dict(rand or {})
        Nc, Ni = int(N_samples), int(N_importance)
        flags = (FLAG_LINDISP if lindisp else 0) | (FLAG_TRAIN if is_train else 0)
        f32 = dict(dtype=torch.float32, device=dev)

        t_rand = noise0 = None
        self._calls += 1
        if perturb > 0.:
            t_rand = rand["t_rand"] if "t_rand" in rand else self._random((R, Nc), 0)
            t_rand = t_rand.float().contiguous()
        if "noise0" in rand:
            noise0 = rand["noise0"].float().contiguous()
        elif raw_noise_std > 0.:
            noise0 = self._random((R, Nc - 1), 1, normal=True, scale=raw_noise_std)
        z0 = torch.empty((R, Nc), **f32)
        w0 = torch.empty((R, Nc), **f32)
        rgb0 = torch.empty((R, 3), **f32)
        depth0 = torch.empty((R,), **f32)
        acc0 = torch.empty((R,), **f32)
        want_feat = use_awp and not force_naive and not inference
        feat0 = torch.empty((R, Nc, 15), **f32) if (want_feat and Ni == 0) else None
        tv = self._linspace(Nc)
        check(self._launch("coarse", lambda: lib.edn_render_coarse_fwd(
            C.byref(self.coarse.grid), C.byref(self.coarse.mlp), ptr(rb), ptr(tv), ptr(t_rand), ptr(noise0), R, Nc,
            flags | FLAG_RELU_RGB, self.rmnearplane, self._coarse_precision(Nc, feat0 is not None), ptr(z0), ptr(w0), ptr(rgb0), ptr(depth0),
            ptr(acc0), ptr(feat0),
            stream_ptr())), "edn_render_coarse_fwd")
        if Ni <= 0:
            ret = {"rgb_map": rgb0, "depth_map": depth0, "acc_map": acc0}
            if retraw:
                ret["z_vals"], ret["weights"] = z0, w0
            if want_feat:
                ret["depth_feature"], ret["z_vals"] = feat0, z0
            return self._guard(ret)

        if self.fine is None:
            raise RuntimeError("render_rays: N_importance > 0 needs mlp_fine.* parameters")
        u = None
        if perturb > 0.:
            u = rand["u"] if "u" in rand else self._random((R, Ni), 2)
        m = self.sample_pdf_merge(z0, w0, Ni, u=u, want_indices=want_indices)
        S = Nc + Ni
        noise1 = None
        if "noise1" in rand:
            noise1 = rand["noise1"].float().contiguous()
        elif raw_noise_std > 0.:
            noise1 = self._random((R, S - 1), 3, normal=True, scale=raw_noise_std)
        w1 = torch.empty((R, S), **f32)
        rgb1 = torch.empty((R, 3), **f32)
        depth1 = torch.empty((R,), **f32)
        acc1 = torch.empty((R,), **f32)
        feat1 = torch.empty((R, S, 128), **f32) if want_feat else None
        check(self._launch("fine", lambda: lib.edn_render_fine_fwd(
            C.byref(self.coarse.grid), C.byref(self.fine.grid), C.byref(self.fine.mlp), ptr(rb), ptr(m["z_vals"]),
            ptr(noise1), R, S, flags, self.rmnearplane, self._fine_precision(S, want_feat), ptr(w1), ptr(rgb1), ptr(depth1), ptr(acc1),
            ptr(feat1), stream_ptr())), "edn_render_fine_fwd")
        ret = {"rgb_map": rgb1, "depth_map": depth1, "acc_map": acc1}
        if retraw:
            ret["z_vals"], ret["weights"] = m["z_vals"], w1
        ret["rgb0"], ret["depth0"], ret["acc0"], ret["z_std"] = rgb0, depth0, acc0, m["z_std"]
        if retraw:
            ret["z_vals0"], ret["weights0"] = z0, w0
        if want_feat:
            ret["depth_feature"], ret["z_vals"] = feat1, m["z_vals"]
        if want_indices:
            ret["inds"], ret["order"], ret["z_samples"] = m["inds"], m["order"], m["z_samples"]
        return self._guard(ret)


# the sampler / launch helpers do not depend on the VM fields: share them with the nerf-mode renderer
SamplerHost._launch = RenderEngine._launch
SamplerHost._linspace = RenderEngine._linspace
SamplerHost.sample_pdf_merge = RenderEngine.sample_pdf_merge
SamplerHost.workspace = RenderEngine.workspace


class NerfRenderEngine(SamplerHost):
    """mode = nerf renderer over a reference state_dict (mlp_coarse.* / mlp_fine.* = networks/nerf.py::NeRF)."""

    def __init__(self, params, rmnearplane=0, use_awp=False, device=None):
        from .nerf_mode import NeRF
        if not torch.cuda.is_available():
            raise RuntimeError("evdeblurnerf_b200.NerfRenderEngine needs a CUDA device (no CPU fallback)")
        _lib.load()
        self._init_host(device)
        self.rmnearplane, self.use_awp = rmnearplane, use_awp
        self.prec_code = EDN_F32
        self.repack(params)

    def repack(self, params):
        from .nerf_mode import NeRF
        P = {k: (v if v.is_cuda else v.to(self.device)) for k, v in params.items() if isinstance(v, torch.Tensor)}
        ef = "before_linear" if self.use_awp else "after_linear"       # renderer.py:88-99
        self.mlp_coarse = NeRF(P, "mlp_coarse.", ef, self.rmnearplane)
        self.mlp_fine = NeRF(P, "mlp_fine.", ef, self.rmnearplane) if "mlp_fine.pts_linears.0.weight" in P else None

    def backward(self, saved, d_out, chunk_rays=4096, white_bkgd=False):
        """Backward of render_rays (mode = nerf) -> ({state_dict name: gradient}, d_ray_batch [R, 11])."""
        rb = saved["ray_batch"].float().contiguous()
        d_rb = torch.zeros_like(rb)
        named = {}
        two_stage = saved.get("z_vals") is not None and self.mlp_fine is not None
        if two_stage:
            named.update(self.mlp_fine.backward(self, rb, saved["z_vals"], saved.get("noise1"), d_out.get("rgb_map"), d_out.get("depth_map"),
                                                d_out.get("acc_map"), d_rb, white_bkgd, d_out.get("depth_feature"), chunk_rays))
            named.update(self.mlp_coarse.backward(self, rb, saved["z_vals0"], saved.get("noise0"), d_out.get("rgb0"), d_out.get("depth0"),
                                                  d_out.get("acc0"), d_rb, white_bkgd, None, chunk_rays))
        else:
            named.update(self.mlp_coarse.backward(self, rb, saved["z_vals0"], saved.get("noise0"), d_out.get("rgb_map"), d_out.get("depth_map"),
                                                  d_out.get("acc_map"), d_rb, white_bkgd, d_out.get("depth_feature"), chunk_rays))
        return named, d_rb

    def render_rays(self, ray_batch, N_samples, **kw):
        from .nerf_mode import render_rays_nerf
        kw.pop("pytest", None)
        kw.pop("want_indices", None)
        return self._guard(render_rays_nerf(self, self.mlp_coarse, self.mlp_fine, ray_batch, N_samples, **kw))
