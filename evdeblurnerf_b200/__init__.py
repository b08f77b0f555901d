"""evdeblurnerf_b200 -- B200-native (sm_100a) render / blur-loss hot path of uzh-rpg/EvDeblurNeRF.

Python here is the host-side mirror of the reference's operator surface; all arithmetic runs in
csrc/libevdeblur_b200.so through the C ABI declared in include/evdeblur_b200.h.  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .engine import NerfRenderEngine, RenderEngine, PackedField  # noqa: F401
from .renderer import NeRFAll, RigidBlurringModel, build_ray_batch, weighted_sum  # noqa: F401
from .dsk import BlurModel  # noqa: F401
from .losses import TonemappingTransform, egm_loss, img2mse, mse2psnr, tv_loss_app, edi_prior_image  # noqa: F401

__all__ = ["RenderEngine", "PackedField", "NeRFAll", "RigidBlurringModel", "BlurModel", "build_ray_batch", "weighted_sum",
           "TonemappingTransform", "egm_loss", "img2mse", "mse2psnr", "tv_loss_app", "edi_prior_image"]
