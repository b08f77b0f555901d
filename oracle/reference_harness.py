"""TEST / BENCH INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference: from /root/reference in the build container, else from
the byte-for-byte staged copy under oracle/_ref/EvDeblurNeRF (oracle/stage_reference.py; git-ignored, travels to the GPU box).

Builds the reference's own modules (NeRFAll + RigidBlurringModel + AdaptiveWeightProposal +
TonemappingTransform) so that `oracle/make_golden.py` can (a) pin the oracle restatement in
`oracle/evdeblur_oracle.py` against the real reference and (b) write the golden fixtures under `tests/golden/`, and so that
`bench.py --impl reference` and bench.py's eager-GPU baseline leg can TIME the reference's own code (never the product path).

Shims; no reference file is modified:
  * kornia / imageio / h5py / configargparse / tensorboardX / skimage / numba -> empty stub modules
    (imported at module top in utils/rays.py:3, networks/dpnerf/mam.py:6, utils/events.py:2 ... never used on the path)
  * Tensor.cuda / Module.cuda -> no-ops (voxnerf.py:86, renderer.py:609, tonemapping.py:147 hard-code .cuda()).
"""
import sys
import types
from types import SimpleNamespace

import torch

import os

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "EvDeblurNeRF")
REFERENCE_ROOT = "/root/reference" if os.path.isdir("/root/reference/networks") else _STAGED


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "networks"))


def _install_shims(cpu=True):
    """cpu = True additionally turns Tensor.cuda / Module.cuda into no-ops (CPU-only process!); with cpu = False the reference
    runs on the GPU as it does in production: the caller wraps construction and calls in `torch.device("cuda")`, the modern
    spelling of run_nerf.py:779's torch.set_default_tensor_type('torch.cuda.FloatTensor')."""
    for name in ("kornia", "imageio", "h5py", "configargparse", "tensorboardX", "wandb", "skimage",
                 "skimage.metrics", "numba"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                m.__dict__["create_meshgrid"] = None
                if name == "numba":
                    m.njit = lambda *a, **k: (a[0] if a and callable(a[0]) else (lambda f: f))
                    m.jit = m.njit
                sys.modules[name] = m
    if cpu:
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def blurfactory_args(E=5, mode="c2f", n_imgs=30, aabb=((-1.5, -1.5, -1.0), (1.5, 1.5, 1.0)),
                     coarse_n_voxels=16777248, fine_n_voxels=134217984, use_awp=True,
                     N_samples=64, N_importance=64, render_rmnearplane=0, rgb_add_bias=False):
    """Namespace with the fields NeRFAll/run_nerf.py read, values from
    configs/evdeblurnerf_blender/tx_blurfactory_evdeblurnerf_ediprior_evcrf.txt:50-110 (kernel_ptnum=E)."""
    a = SimpleNamespace()
    a.mode = mode
    a.multires, a.multires_views, a.use_viewdirs = 10, 4, True
    a.kernel_type, a.kernel_use_awp, a.kernel_ptnum = "RBK", use_awp, E
    a.N_samples, a.N_importance = N_samples, N_importance
    a.bounding_box = (torch.tensor(aabb[0], dtype=torch.float32), torch.tensor(aabb[1], dtype=torch.float32))
    a.coarse_num_layers, a.coarse_hidden_dim = 2, 64
    a.coarse_num_layers_color, a.coarse_hidden_dim_color = 3, 64
    a.coarse_app_dim, a.coarse_app_n_comp, a.coarse_n_voxels = 32, [64, 16, 16], coarse_n_voxels
    a.coarse_app_actfn = "none"
    a.fine_num_layers, a.fine_hidden_dim = 2, 256
    a.fine_num_layers_color, a.fine_hidden_dim_color = 3, 256
    a.fine_geo_feat_dim, a.fine_app_dim, a.fine_app_n_comp, a.fine_n_voxels = 128, 32, [64, 16, 16], fine_n_voxels
    a.fine_app_actfn = "none"
    a.kernel_feat_cnl = 15
    a.rgb_add_bias, a.rgb_activate, a.sigma_activate = rgb_add_bias, "sigmoid", "relu"
    a.render_rmnearplane = render_rmnearplane
    a.netdepth = a.netdepth_fine = 8
    a.netwidth = a.netwidth_fine = 256
    a.n_imgs = n_imgs
    a.kernel_img_embed = 32
    return a


def build_reference(args, seed=0, nontrivial_rbk=True, cpu=True):
    """Construct the reference modules exactly as run_nerf.py:166-244 does (RBK + AWP + CRF + NeRFAll)."""
    _install_shims(cpu)
    from networks.embedding import ViewEmbedding
    from networks.dpnerf.blurmodel import RigidBlurringModel
    from networks.dpnerf.awp import AdaptiveWeightProposal
    from networks.renderer import NeRFAll
    from networks.tonemapping import TonemappingTransform

    torch.manual_seed(seed)
    view_embed = ViewEmbedding(num_embed=args.n_imgs, embed_dim=args.kernel_img_embed, init_params="zero")
    kernelnet = RigidBlurringModel(
        feat_ch=0, num_motion=args.kernel_ptnum - 1, D_r=1, W_r=32, D_v=1, W_v=32, D_w=1, W_w=32,
        output_ch_r=3, output_ch_v=3, rv_window=0.1, use_origin=True, view_embed=view_embed, W=32)
    awpnet = None
    if args.kernel_use_awp:
        awpnet = AdaptiveWeightProposal(
            input_ch=args.fine_geo_feat_dim if args.mode == "c2f" else args.netwidth,
            num_motion=args.kernel_ptnum - 1, use_origin=True, D_sam=4, W_sam=64, D_mot=1, W_mot=32,
            dir_freq=2, rgb_freq=2, depth_freq=3, ray_dir_freq=2, view_feature_ch=32)
    crf = TonemappingTransform(map_type_rgb="gamma", map_type_event="learn", extra_features_event=2,
                               gamma=2.2, init_learn_identity=False)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        nerf = NeRFAll(args, kernelnet, awpnet)
    if nontrivial_rbk:
        # zero latents + 1e-5 heads give identity warps; make the warp observable
        g = torch.Generator().manual_seed(seed + 1)
        with torch.no_grad():
            view_embed.img_embed.copy_(torch.randn(view_embed.img_embed.shape, generator=g) * 0.5)
            kernelnet.r_linear.weight.copy_(torch.randn(kernelnet.r_linear.weight.shape, generator=g) * 0.05)
            kernelnet.v_linear.weight.copy_(torch.randn(kernelnet.v_linear.weight.shape, generator=g) * 0.05)
    return nerf, crf


def synthetic_rays(N, seed=0, n_imgs=30, H=400, W=400):
    """SURVEY 8(d) synthetic batch: camera looking down -z so that NDC is well posed."""
    g = torch.Generator().manual_seed(seed)
    o = torch.randn(N, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0])
    d = torch.cat([torch.randn(N, 2, generator=g) * 0.3, -torch.ones(N, 1)], -1)
    rays = torch.stack([o, d], -1)  # [N,3,2]
    images_idx = torch.randint(0, n_imgs, (N, 1), generator=g)
    return rays, images_idx


def build_bench_reference(P, E=5, use_awp=False, device="cpu"):
    """The reference's NeRFAll for bench.py's synthetic blurfactory workload: constructed by the reference's own code, then
    loaded with the bench parameters `P` (reference state_dict names; tensors missing from P -- e.g. awpnet.* -- keep the
    reference's own initialisation).  device = "cpu": Tensor.cuda shimmed away; "cuda[:i]": CUDA is the default device."""
    import contextlib
    cpu = str(device) == "cpu"
    args = blurfactory_args(E=E, use_awp=use_awp)
    ctx = contextlib.nullcontext() if cpu else torch.device(device)
    with ctx:
        args.bounding_box = (args.bounding_box[0].to(device), args.bounding_box[1].to(device))
        nerf, _ = build_reference(args, nontrivial_rbk=False, cpu=cpu)
        sd = nerf.state_dict()
        bad = [k for k, v in P.items() if k in sd and tuple(sd[k].shape) != tuple(v.shape)]
        if bad:
            raise RuntimeError(f"bench parameters do not fit the reference model: {bad[:4]}")
        missing = [k for k in sd if k not in P and not k.startswith("awpnet.")]
        if missing:
            raise RuntimeError(f"bench parameters miss reference tensors: {missing[:4]}")
        nerf.load_state_dict({k: v.to(device) for k, v in P.items() if k in sd}, strict=False)
        nerf = nerf.to(device)
    return nerf.train()
