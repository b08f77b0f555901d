"""The reference's training iteration (run_nerf.py:423-613), mode = c2f + RBK, re-laid-out for one process per GPU.

* Every trainable tensor (the two PDRF fields, the DP-NeRF kernel net, the CRF) is a VIEW into one flat fp32 buffer, under
  its reference state_dict name and shape (checkpoints round-trip); `.grad`s are views into a second flat buffer.  The
  gradient exchange across ranks is therefore ONE all-reduce (NCCL over NVLink) and the optimizer ONE fused
  Adam launch per parameter group (edn_adam_step) -- torch.optim.Adam semantics, betas (0.9, 0.999), optional L2 weight
  decay on `color_net.N.weight` (run_nerf.py:243-274), learning-rate warm-up + exponential decay (run_nerf.py:604-613).
* Rays shard across ranks (each rank renders its own slice of the batch); losses are means over rays, so gradients are
  averaged (all-reduce sum / world size).
* Loss mixing follows run_nerf.py:448-504, 539-594: photometric MSE on CRF-encoded fine + coarse colours, TV regulariser,
  optional event generation-model loss on start/end event rays rendered through the force_naive branch.
"""
import re

import torch

from . import _lib
from ._lib import check, stream_ptr
from .schedules import annealing_interpolator, checkpoint_dict, exponential_scale_fine_loss_weight, load_checkpoint_dict, optimizer_groups
from .losses import TonemappingTransform, egm_loss, img2mse
from .renderer import NeRFAll

_WD_RE = re.compile(r"\.color_net\.[0-9]+\.weight$")     # run_nerf.py:246
_BUFFER_SUFFIXES = (".running_mean", ".running_var", ".num_batches_tracked")


def lr_at(step, lrate, decay_k, warmup_iters=0, warmup_factor=1.0):
    """run_nerf.py:604-613: the learning rate in force AFTER `step` updates (global_step)."""
    if warmup_iters > 0 and step < warmup_iters:
        return lrate * ((1 - warmup_factor) * step / warmup_iters + warmup_factor)
    return lrate * (0.1 ** (step / (decay_k * 1000)))


class FlatParams:
    """Flat fp32 parameter / gradient / Adam-moment buffers with named views; group 0 = weight-decayed tensors."""

    def __init__(self, tensors, device, weight_decay_re=_WD_RE):
        names = sorted(tensors)
        wd = [n for n in names if weight_decay_re.search(n)]
        rest = [n for n in names if n not in wd]
        self.order = wd + rest
        self.offset, off = {}, 0
        for n in self.order:
            self.offset[n] = off
            off += (tensors[n].numel() + 3) // 4 * 4          # 16-byte aligned views
            if n == (wd[-1] if wd else None):
                self.split = off
        if not wd:
            self.split = 0
        self.numel = off
        self.param = torch.zeros(off, dtype=torch.float32, device=device)
        self.grad = torch.zeros_like(self.param)
        self.exp_avg = torch.zeros_like(self.param)
        self.exp_avg_sq = torch.zeros_like(self.param)
        self.views = {}
        for n in self.order:
            t = tensors[n]
            v = self.param[self.offset[n]: self.offset[n] + t.numel()].view(t.shape)
            v.copy_(t.detach().to(device=device, dtype=torch.float32))
            v.requires_grad_(True)
            v.grad = self.grad[self.offset[n]: self.offset[n] + t.numel()].view(t.shape)
            self.views[n] = v

    def named(self, flat):
        """name -> view of `flat` (a buffer laid out like `param`, e.g. exp_avg) with the parameter's shape."""
        return {n: flat[self.offset[n]: self.offset[n] + v.numel()].view(v.shape) for n, v in self.views.items()}

    profile = None      # set to [] to collect (start, end) CUDA events around the gradient exchange (bench.py)

    def all_reduce_mean(self, group=None):
        """The one gradient exchange of a step: mean over ranks of the flat gradient buffer, in place.  NCCL averages inside the
        collective (ReduceOp.AVG: no separate division sweep over the 147 MB buffer); gloo (CPU tests) sums, then divides."""
        dist = torch.distributed
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            ev = None
            if self.profile is not None and self.grad.is_cuda:
                ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                ev[0].record()
            if dist.get_backend(group) == "nccl":
                dist.all_reduce(self.grad, op=dist.ReduceOp.AVG, group=group)
            else:
                dist.all_reduce(self.grad, group=group)
                self.grad.div_(dist.get_world_size(group))
            if ev is not None:
                ev[1].record()
                self.profile.append(ev)

    def adam(self, lr, step, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8):
        lib = _lib.load()
        for lo, hi, wd in ((0, self.split, weight_decay), (self.split, self.numel, 0.0)):
            if hi > lo:
                check(lib.edn_adam_step(self.param[lo:].data_ptr(), self.grad[lo:].data_ptr(), self.exp_avg[lo:].data_ptr(),
                                        self.exp_avg_sq[lo:].data_ptr(), hi - lo, float(lr), betas[0], betas[1], eps, float(wd),
                                        int(step), stream_ptr()), "edn_adam_step")


class Trainer:
    def __init__(self, state, crf_state, aabb_min, aabb_max, kernel_ptnum=5, precision="bf16", lrate=5e-4, lrate_decay=250,
                 lrate_warmup_iters=0, lrate_warmup_factor=1.0, colornet_weightdecay=0.0, tv_loss_weight=1e-2,
                 event_loss_weight=0.0, crf_kwargs=None, render_kwargs=None, device=None, process_group=None, seed=0,
                 use_awp=False, awp_fine_loss_weight=None, schedule=None, check_numerics_every=100, kernel_cfg=None):
        if not torch.cuda.is_available():
            raise RuntimeError("evdeblurnerf_b200.Trainer needs a CUDA device (no CPU fallback)")
        dev = torch.device(device if device is not None else "cuda")
        # nn.Module buffers (the BatchNorm running statistics of awpnet.MAM) are state, not parameters: they stay out of the flat
        # parameter buffer, the optimizer groups and the Adam state (run_nerf.py:243-261 iterates named_parameters())
        trainable = {k: v for k, v in state.items() if isinstance(v, torch.Tensor) and v.is_floating_point()
                     and k.startswith(("mlp_coarse.", "mlp_fine.", "kernelsnet.") + (("awpnet.",) if use_awp else ()))
                     and not k.endswith(_BUFFER_SUFFIXES)}
        crf_train = {"crf." + k: v for k, v in (crf_state or {}).items() if isinstance(v, torch.Tensor) and v.is_floating_point()}
        self.flat = FlatParams({**trainable, **crf_train}, dev)
        P = {k: self.flat.views[k] for k in trainable}
        Pc = {k[4:]: self.flat.views[k] for k in crf_train}
        self.nerf = NeRFAll(P, aabb_min, aabb_max, kernel_ptnum=kernel_ptnum, precision=precision, use_awp=use_awp,
                            kernel_cfg=kernel_cfg).train()
        if use_awp:     # BatchNorm statistics over the WHOLE batch, as on one GPU (SURVEY 8(e) caveat 1)
            self.nerf.awpnet.sync_bn, self.nerf.awpnet.group = True, process_group
        self.awp_fine_loss_weight = awp_fine_loss_weight
        # one render + one backward for the blurred rays and both event ray sets (RBK; the DSK rays come out of their own autograd node)
        self.fuse_event_renders = self.nerf.kernel_type != "DSK"
        self.crf = TonemappingTransform(Pc, **(crf_kwargs or dict(map_type_rgb="gamma", map_type_event="learn" if Pc else "gamma",
                                                                  extra_features_event=2)))
        self.hp = dict(lrate=lrate, decay=lrate_decay, warm_it=lrate_warmup_iters, warm_f=lrate_warmup_factor,
                       wd=colornet_weightdecay, tv_w=tv_loss_weight, ev_w=event_loss_weight)
        self.render_kwargs = dict(render_kwargs or dict(N_samples=64, N_importance=64, perturb=1., raw_noise_std=1.))
        self.group = process_group
        self.world = torch.distributed.get_world_size(process_group) if torch.distributed.is_initialized() else 1
        self.rank = torch.distributed.get_rank(process_group) if torch.distributed.is_initialized() else 0
        self.nerf.engine.seed(seed * 1000003 + self.rank)      # per-rank draw streams (SURVEY 8(e) caveat 2)
        self.global_step = 0
        # NaN / Inf guard (renderer.py:259-263): the render kernels' flag words are read back -- one host synchronisation -- every
        # `check_numerics_every` steps (0 = never); findings are printed like the reference's and kept in `numerical_errors`
        self.check_numerics_every, self.numerical_errors = int(check_numerics_every), []
        # run_nerf.py:121-142, 437-499 loss schedule.  Keys (reference option names, options.py:39,173-232): N_iters,
        # kernel_start_iter, kernel_start_warmup_mode ("step" | "linear" | "cosine"), kernel_start_warmup_iters,
        # kernel_awp_use_coarse_to_fine_opt, use_pts0_prior, pts0_target_{weight,weight_end,weight_steps,weight_scheduler,
        # start_iter,end_iter}, blur_loss_after, event_egm_{weight,weight_end,weight_steps,weight_scheduler}, clip_grads_norm,
        # tone_mapping_start_learn_iter, add_event_egm_startiter, add_event_egm_stages, event_egm_use_color_weights,
        # event_egm_color_weights_start_iter, kernel_awp_fine_loss_start_ratio, kernel_align_weight, align_start_iter, align_end_iter
        sc = dict(kernel_start_iter=0, kernel_start_warmup_mode="step", kernel_start_warmup_iters=1, N_iters=200000,
                  kernel_awp_use_coarse_to_fine_opt=False, use_pts0_prior=None, pts0_target_weight=0.1, pts0_target_weight_end=1.0,
                  pts0_target_weight_steps=None, pts0_target_weight_scheduler="constant", pts0_target_start_iter=-1,
                  pts0_target_end_iter=9999999, blur_loss_after=-1, event_egm_weight=event_loss_weight,
                  event_egm_weight_end=event_loss_weight, event_egm_weight_steps=None, event_egm_weight_scheduler="constant",
                  clip_grads_norm=None, tone_mapping_start_learn_iter=0, add_event_egm_startiter=None,
                  add_event_egm_stages=("stage0", "stage1"), event_egm_use_color_weights=None, event_egm_color_weights_start_iter=-1,
                  kernel_awp_fine_loss_start_ratio=0.1, kernel_align_weight=0.0, align_start_iter=0, align_end_iter=1e10)
        unknown = set(schedule or {}) - set(sc)
        if unknown:
            raise ValueError(f"unknown schedule keys {sorted(unknown)}")
        sc.update(schedule or {})
        self.schedule = sc
        self.w_events_egm = annealing_interpolator(sc["event_egm_weight"], sc["event_egm_weight_end"], sc["event_egm_weight_steps"],
                                                   sc["event_egm_weight_scheduler"])
        self.w_pts0_target = annealing_interpolator(sc["pts0_target_weight"], sc["pts0_target_weight_end"],
                                                    sc["pts0_target_weight_steps"], sc["pts0_target_weight_scheduler"])
        self.w_kernel, self.kernel_end_warmup_iter = (lambda step: 1.0), -1
        if sc["kernel_start_warmup_mode"] != "step":
            self.kernel_end_warmup_iter = sc["kernel_start_iter"] + sc["kernel_start_warmup_iters"]
            self.w_kernel = annealing_interpolator(0.0, 1.0, self.kernel_end_warmup_iter, sc["kernel_start_warmup_mode"],
                                                   start_step=sc["kernel_start_iter"])
        self._fine_loss_weight = None
        self._names = {"net": list(trainable), "crf": list(crf_state or {})}
        self._buffers = {k: v.detach().clone() for k, v in state.items() if isinstance(v, torch.Tensor) and k not in trainable}

    # run_nerf.py:438-504 (+ 506-592 when event rays are given)
    def loss(self, batch, H, W, K):
        out = {}
        sc, g = self.schedule, self.global_step
        i = g + 1      # the reference's loop index: `start = global_step; for i in range(start + 1, ...)` (run_nerf.py:421-423), so
        #                comparisons written against `i` see global_step + 1 and the schedule functions see global_step
        force_naive = i < sc["kernel_start_iter"]                       # run_nerf.py:440: the blur kernel is off at first
        skip_crf = i < sc["tone_mapping_start_learn_iter"]              # run_nerf.py:443: the learnt CRF is bypassed at first
        w_ev = self.w_events_egm(g)
        events = ("ev_rays_start" in batch and (w_ev or 0.0) > 0
                  and (sc["add_event_egm_startiter"] is None or i >= sc["add_event_egm_startiter"]))
        use_pts0 = sc["use_pts0_prior"] is not None and sc["pts0_target_start_iter"] <= i < sc["pts0_target_end_iter"]
        warm = sc["kernel_start_warmup_mode"] != "step" and sc["kernel_start_iter"] <= g < self.kernel_end_warmup_iter
        want_pts0 = g < self.kernel_end_warmup_iter or use_pts0        # run_nerf.py:441
        render_kwargs = dict(self.render_kwargs, return_pts0_rgb=True) if want_pts0 else self.render_kwargs
        crf = lambda x, **kw: self.crf(x, skip_learn_crf=skip_crf, **kw)
        naive = None
        if events and self.fuse_event_renders and not force_naive:
            # the reference calls nerf() three times (blurred rays, event start rays, event end rays; run_nerf.py:438, 534, 547);
            # the rays are independent, so one fused render + one backward gives the same result with a third of the launches
            rgb, rgb0, extra_loss, extra_tensor, naive = self.nerf.forward_fused(
                H, W, K, batch["rays"], batch, [batch["ev_rays_start"], batch["ev_rays_end"]], retraw=True, **render_kwargs)
        else:
            rgb, rgb0, extra_loss, extra_tensor = self.nerf(H, W, K, rays=batch["rays"], rays_info=batch, retraw=True,
                                                            force_naive=force_naive, **render_kwargs)
        target = batch["rgbsf"].reshape(-1, 3)
        loss = 0.0
        if i > sc["blur_loss_after"]:                          # run_nerf.py:452-462: no photometric term before blur_loss_after
            img_loss = img2mse(crf(rgb, mode="encode_rgb"), target)
            out["img_loss"] = img_loss.detach()
            if rgb0 is not None:
                img_loss = img_loss + img2mse(crf(rgb0, mode="encode_rgb"), target)
            loss = img_loss
        if extra_tensor.get("rgb_awp") is not None:          # run_nerf.py:464-475
            fine = img2mse(crf(extra_tensor["rgb_awp"], mode="encode_rgb"), target)
            out["img_fine_loss"] = fine.detach()
            flw = self.awp_fine_loss_weight                   # kernel_awp_use_coarse_to_fine_opt: annealed mix, else plain sum
            if flw is None and sc["kernel_awp_use_coarse_to_fine_opt"]:
                if self._fine_loss_weight is None:            # run_nerf.py:416: starts at kernel_awp_fine_loss_start_ratio
                    self._fine_loss_weight = sc["kernel_awp_fine_loss_start_ratio"]
                if i % 10000 == 0:                            # refreshed every 10 000 iterations (run_nerf.py:467-470; N_iters + 1 there)
                    self._fine_loss_weight = exponential_scale_fine_loss_weight(sc["N_iters"] + 1, sc["kernel_start_iter"], 0.1, 0.9, i)
                flw = self._fine_loss_weight
            loss = loss + fine if flw is None else loss * (1 - flw) + fine * flw
        if warm or use_pts0:                                  # run_nerf.py:475-499
            tgt0 = batch["rgbsf_pts0"].reshape(-1, 3) if use_pts0 else target
            pts0_loss = 0.0
            for name in ("stage0_rgb_pts0", "stage1_rgb_pts0", "stage1_rgb1_pts0"):
                if name in extra_tensor:
                    pts0_loss = pts0_loss + img2mse(crf(extra_tensor[name], mode="encode_rgb"), tgt0)
            out["pts0_loss"] = pts0_loss.detach()
            if use_pts0:
                w_pts0 = 1.0 if i <= sc["blur_loss_after"] else self.w_pts0_target(g)
                loss = loss + pts0_loss * w_pts0
            else:
                loss = self.w_kernel(g) * loss + (1 - self.w_kernel(g)) * pts0_loss
        if self.hp["tv_w"] > 0 and extra_loss.get("TV") is not None:
            loss = loss + extra_loss["TV"] * self.hp["tv_w"]
        if "align" in extra_loss and sc["align_start_iter"] <= i <= sc["align_end_iter"]:      # run_nerf.py:502-504 (DSK kernels)
            loss = loss + extra_loss["align"].reshape(()) * sc["kernel_align_weight"]
        if events:
            feat = batch.get("ev_extra_feat")
            cmask = batch.get("ev_color_map")
            cweight = sc["event_egm_use_color_weights"] if i > sc["event_egm_color_weights_start_iter"] else None
            ev_kw = {"tonemap_only": True} if cmask is not None else {}      # event_egm_use_colorevents (run_nerf.py:522)
            lum = []
            for j, key in enumerate(("ev_rays_start", "ev_rays_end")):
                if naive is not None:
                    c, c0 = naive[j]
                else:
                    c, c0, _, _ = self.nerf(H, W, K, rays=batch[key], rays_info=None, retraw=True, force_naive=True, want_tv=False,
                                            **self.render_kwargs)   # TV of these calls is never used (run_nerf.py:500-501)
                lum.append((crf(c, mode="encode_luma", ev_extra_feat=feat, **ev_kw),
                            crf(c0, mode="encode_luma", ev_extra_feat=feat, **ev_kw) if c0 is not None else None))
            ev = 0.0
            if "stage0" in sc["add_event_egm_stages"] and lum[0][1] is not None:          # run_nerf.py:562-567
                ev = ev + egm_loss(lum[0][1], lum[1][1], batch["bii"], color_mask=cmask, color_weight=cweight)
            if "stage1" in sc["add_event_egm_stages"]:                                     # run_nerf.py:568-571
                ev = ev + egm_loss(lum[0][0], lum[1][0], batch["bii"], color_mask=cmask, color_weight=cweight)
            out["event_loss"] = ev.detach() if isinstance(ev, torch.Tensor) else ev
            loss = loss + ev * w_ev
        out["loss"] = loss
        return out

    def step(self, batch, H, W, K):
        """One optimisation step on this rank's shard of the batch -> dict of detached loss terms."""
        hp = self.hp
        self.flat.grad.zero_()
        out = self.loss(batch, H, W, K)
        out["loss"].backward()
        self.flat.all_reduce_mean(self.group)
        if self.schedule["clip_grads_norm"] is not None:       # run_nerf.py:596-599: over nerf.parameters() only (not the CRF)
            grads = [self.flat.views[k].grad for k in self._names["net"]]
            total = torch.linalg.vector_norm(torch.stack(torch._foreach_norm(grads)))
            coef = torch.clamp(self.schedule["clip_grads_norm"] / (total + 1e-6), max=1.0)
            torch._foreach_mul_(grads, coef)
            out["grad_norm"] = total.detach()
        self.global_step += 1
        # iteration g of the reference runs with the rate it set at the end of iteration g - 1 from global_step = g - 1
        lr = lr_at(max(self.global_step - 2, 0), hp["lrate"], hp["decay"], hp["warm_it"], hp["warm_f"])
        self.flat.adam(lr, self.global_step, weight_decay=hp["wd"])
        self.nerf.repack()
        out["loss"] = out["loss"].detach()
        out["lr"] = lr
        if self.check_numerics_every > 0 and self.global_step % self.check_numerics_every == 0:
            out["numerical_errors"] = self.poll_numerics()
        return out

    def poll_numerics(self):
        """Lazy read of the device NaN / Inf flags (synchronises) -> messages, printed in the reference's format."""
        msgs = self.nerf.engine.numerical_errors()
        for m_ in msgs:
            print(f"! [Numerical Error] {m_}")            # renderer.py:260-263
        self.numerical_errors += [(self.global_step, m_) for m_ in msgs]
        return msgs

    def state_dict(self):
        """Reference-format tensors (run_nerf.py:628-634 saves network / optimizer state dicts)."""
        return {k: v.detach().clone() for k, v in self.flat.views.items()}

    def _groups(self):
        return optimizer_groups(self._names["net"], self._names["crf"], mode="c2f", colornet_weightdecay=self.hp["wd"])

    def checkpoint(self):
        """The dict run_nerf.py:628-634 passes to torch.save: network / crf state dicts (buffers included), a genuine
        torch.optim.Adam state_dict in the reference's group order, global_step."""
        views = self.flat.views
        net = {**{k: views[k] for k in self._names["net"]}, **self._buffers}
        crf = {k: views["crf." + k] for k in self._names["crf"]}
        m, v = self.flat.named(self.flat.exp_avg), self.flat.named(self.flat.exp_avg_sq)
        lr = lr_at(max(self.global_step - 1, 0), self.hp["lrate"], self.hp["decay"], self.hp["warm_it"], self.hp["warm_f"])
        return checkpoint_dict(self.global_step, net, crf, m, v, self._groups(), lr=lr, initial_lr=self.hp["lrate"],
                               colornet_weightdecay=self.hp["wd"])

    def load_checkpoint(self, ckpt):
        """Reload what `checkpoint()` -- or the reference itself -- saved (run_nerf.py:282-295)."""
        step, net, crf, m, v, _ = load_checkpoint_dict(ckpt, self._groups())
        views = self.flat.views
        ma, va = self.flat.named(self.flat.exp_avg), self.flat.named(self.flat.exp_avg_sq)
        with torch.no_grad():
            for k in self._names["net"]:
                views[k].copy_(net[k])
            for k in self._names["crf"]:
                views["crf." + k].copy_(crf[k])
            for k, t in m.items():
                ma[k].copy_(t)
                va[k].copy_(v[k])
        self._buffers.update({k: t.detach().clone() for k, t in net.items() if k in self._buffers})
        self.global_step = step
        self.nerf.repack()
