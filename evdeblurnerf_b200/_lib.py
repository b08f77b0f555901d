"""ctypes binding of the C ABI in include/evdeblur_b200.h (libevdeblur_b200.so, built in-tree by csrc/build.py).

There is no CPU fallback: importing the package works without the library (so that host-side logic can be tested on a
CPU box), but every compute entry point raises if the library is missing or CUDA is unavailable.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libevdeblur_b200.so")

EDN_F32, EDN_BF16, EDN_TC32 = 0, 1, 2
FLAG_LINDISP, FLAG_TRAIN, FLAG_RELU_RGB = 1, 2, 4

c_float_p = C.c_void_p   # device pointers travel as plain integers


class VmGrid(C.Structure):
    _fields_ = [("plane", C.c_void_p * 3), ("line", C.c_void_p * 3), ("plane_h", C.c_int32 * 3),
                ("plane_w", C.c_int32 * 3), ("line_len", C.c_int32 * 3), ("n_comp", C.c_int32 * 3),
                ("dtype", C.c_int32), ("basis_t", C.c_void_p), ("aabb_min", C.c_float * 3),
                ("aabb_max", C.c_float * 3)]


class FieldMlp(C.Structure):
    _fields_ = [("sigma0_t", C.c_void_p), ("sigma1_t", C.c_void_p), ("sigma1_v", C.c_void_p),
                ("color0_t", C.c_void_p), ("color1_t", C.c_void_p), ("color2_t", C.c_void_p),
                ("color0_b", C.c_void_p), ("color1_b", C.c_void_p), ("color2_b", C.c_void_p),
                ("hidden", C.c_int32), ("geo_feat", C.c_int32), ("tc_blob", C.c_void_p)]


class FieldWeights(C.Structure):
    _fields_ = [("basis", C.c_void_p * 2)] + [(n, C.c_void_p) for n in (
        "sigma0", "sigma1", "color0", "color1", "color2", "color0_b", "color1_b", "color2_b")] + [
        ("hidden", C.c_int32), ("geo_feat", C.c_int32), ("n_grids", C.c_int32)]


class VmGridGrad(C.Structure):
    _fields_ = [("plane", C.c_void_p * 3), ("line", C.c_void_p * 3)]


class FieldBwdMerge(C.Structure):
    _fields_ = [("order", C.c_void_p), ("n_coarse", C.c_int32), ("moved", C.c_void_p)]


class RbkParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("img_embed", "r_branch_w", "r_branch_b", "v_branch_w", "v_branch_b", "w_branch_w",
                                           "w_branch_b", "r_linear_w", "r_linear_b", "v_linear_w", "v_linear_b", "w_linear_w",
                                           "w_linear_b")] + [("num_motion", C.c_int32), ("n_img", C.c_int32),
                                                             ("rv_window", C.c_float)]


class RbkGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("img_embed", "r_branch_w", "r_branch_b", "v_branch_w", "v_branch_b", "w_branch_w",
                                           "w_branch_b", "r_linear_w", "r_linear_b", "v_linear_w", "v_linear_b", "w_linear_w",
                                           "w_linear_b")]


DSK_MAX_HIDDEN = 4
_DSK_PTRS = [("img_embed", C.c_void_p), ("pattern_pos", C.c_void_p), ("pattern_trans", C.c_void_p),
             ("lin_w", C.c_void_p * DSK_MAX_HIDDEN), ("lin_b", C.c_void_p * DSK_MAX_HIDDEN),
             ("out0_w", C.c_void_p), ("out0_b", C.c_void_p), ("out1_w", C.c_void_p), ("out1_b", C.c_void_p)]


class DskParams(C.Structure):
    _fields_ = _DSK_PTRS + [(n, C.c_int32) for n in ("n_img", "n_pt", "embed", "in_embed", "spatial_embed", "num_hidden", "wide",
                                                      "short_cut", "isglobal", "optim_sv_trans")] + [("kernel_hwindow", C.c_float)]


class DskGrads(C.Structure):
    _fields_ = list(_DSK_PTRS)


class CrfGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w0", "b0", "w1", "b1", "w2", "b2", "w3", "b3")]


class CrfParams(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w0", "b0", "w1", "b1", "w2", "b2", "w3", "b3")] + [
        ("extra_features", C.c_int32), ("gamma", C.c_float)]


class NerfMlp(C.Structure):
    _fields_ = [("pts_t", C.c_void_p * 8), ("pts_b", C.c_void_p * 8)] + [(n, C.c_void_p) for n in (
        "alpha_w", "alpha_b", "feature_t", "feature_b", "views_t", "views_b", "rgb_t", "rgb_b")]


class NerfWeights(C.Structure):
    _fields_ = [("pts_w", C.c_void_p * 8), ("pts_b", C.c_void_p * 8)] + [(n, C.c_void_p) for n in (
        "alpha_w", "alpha_b", "feature_w", "feature_b", "views_w", "views_b", "rgb_w", "rgb_b")]


class AwpParams(C.Structure):
    _fields_ = [("sample_t", C.c_void_p * 4), ("sample_b", C.c_void_p * 4), ("motion_w", C.c_void_p * 2),
                ("motion_b", C.c_void_p * 2)] + [(n, C.c_void_p) for n in (
                    "mam_linear_t", "mam_linear_b", "line_conv_att", "conva", "convb", "convc", "convn", "convl", "convd_w",
                    "bn_weight", "bn_bias", "w_linear_w", "w_linear_b")] + [("input_ch", C.c_int32)]


class AwpOptions(C.Structure):
    _fields_ = [("precision", C.c_int32), ("keep_activations", C.c_int32), ("phase", C.c_int32), ("bn_rows_total", C.c_int64)]


class AwpGrads(AwpParams):
    pass


CRF_GAMMA, CRF_LEARN, CRF_SKIP_LEARN, CRF_LUMA = 1, 2, 4, 8
CRF_LUMA_REC709, CRF_LUMA_AVG = 16, 32
FLAG_WHITE_BKGD = 8

# name -> (restype, argtypes); must list every symbol include/evdeblur_b200.h declares (tests/test_abi.py checks)
_P, _I32, _I64, _F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
SIGNATURES = {
    "edn_last_error": (C.c_char_p, []),
    "edn_abi_version": (C.c_int, []),
    "edn_pack_vm_plane": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P]),
    "edn_fill_random": (C.c_int, [_P, _I64, C.c_uint64, C.c_uint32, _I32, _F, _P]),
    "edn_vm_sample": (C.c_int, [C.POINTER(VmGrid), _P, _P, _I64, _P]),
    "edn_check_finite": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int64), _I32, _P, _P]),
    "edn_render_coarse_fwd": (C.c_int, [C.POINTER(VmGrid), C.POINTER(FieldMlp), _P, _P, _P, _P, _I64, _I32, _I32, _F,
                                        _I32, _P, _P, _P, _P, _P, _P, _P]),
    "edn_coarse_tc_blob_bytes": (C.c_int64, []),
    "edn_coarse_tc_pack_workspace_floats": (C.c_int64, []),
    "edn_pack_coarse_tc": (C.c_int, [C.POINTER(FieldMlp), _P, _P, _P, _P]),
    "edn_sample_pdf_merge": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _P, _P, _P, _P, _P, _P]),
    "edn_fine_tc_blob_bytes": (C.c_int64, []),
    "edn_fine_tc_pack_workspace_floats": (C.c_int64, []),
    "edn_pack_fine_tc": (C.c_int, [C.POINTER(FieldMlp), _P, _P, _P, _P, _P]),
    "edn_render_fine_fwd": (C.c_int, [C.POINTER(VmGrid), C.POINTER(VmGrid), C.POINTER(FieldMlp), _P, _P, _P, _I64,
                                      _I32, _I32, _F, _I32, _P, _P, _P, _P, _P, _P]),
    "edn_field_bwd_workspace_bytes": (C.c_int64, [_I32, _I32, _I32, _I64, _I32]),
    "edn_render_field_bwd": (C.c_int, [C.POINTER(VmGrid), C.POINTER(VmGrid), C.POINTER(FieldWeights), _P, _P, _P, _I64, _I32,
                                       _I32, _P, _P, _P, _P, _P, C.POINTER(FieldWeights), C.POINTER(VmGridGrad),
                                       C.POINTER(VmGridGrad), _P, _P, _I64, C.POINTER(FieldBwdMerge), _P]),
    "edn_unpack_vm_plane_grad": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _P]),
    "edn_rbk_bwd_workspace_floats": (C.c_int64, [_I64, _I32]),
    "edn_rbk_warp_ndc_bwd": (C.c_int, [C.POINTER(RbkParams), _P, _P, _I64, _I32, _I32, _F, _I32, _P, _P, _P, C.POINTER(RbkGrads), _P, _P]),
    "edn_awp_bwd_workspace_floats": (C.c_int64, [_I64, _I32, _I32]),
    "edn_awp_bwd": (C.c_int, [C.POINTER(AwpParams), _P, _P, _P, _I32, _P, _I64, _I32, _I32, _F, C.POINTER(AwpOptions), _I32, _P, C.POINTER(AwpGrads), _P,
                              _P, _I32, _P, _P, _P]),
    "edn_weighted_sum_bwd": (C.c_int, [_P, _P, _P, _I64, _I32, _I64, _P, _P, _P]),
    "edn_crf_bwd": (C.c_int, [C.POINTER(CrfParams), _P, _P, _I32, _I32, _I64, _P, _P, C.POINTER(CrfGrads), _P]),
    "edn_egm_loss_bwd": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I64, _F, _P, _P, _P, _P]),
    "edn_img2mse_bwd": (C.c_int, [_P, _P, _I64, _P, _P, _P]),
    "edn_tv_loss_app_bwd": (C.c_int, [C.POINTER(C.c_void_p * 3), C.POINTER(C.c_void_p * 3), C.POINTER(C.c_int32 * 3),
                                      C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3), _P,
                                      C.POINTER(C.c_void_p * 3), C.POINTER(C.c_void_p * 3), _P]),
    "edn_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F, _F, _F, _F, _F, _I64, _P]),
    "edn_nerf_bwd_workspace_bytes": (C.c_int64, [_I64, _I32]),
    "edn_nerf_field_bwd": (C.c_int, [C.POINTER(NerfWeights), _P, _P, _P, _I64, _I32, _I32, _I32, _P, _P, _P, _P, _P, _I32,
                                     C.POINTER(NerfWeights), _P, _P, _I64, _P]),
    "edn_nerf_mlp_fwd": (C.c_int, [C.POINTER(NerfMlp), _P, _P, _I64, _I32, _I32, _P, _P, _P]),
    "edn_nerf_raw2outputs": (C.c_int, [_P, _P, _P, _P, _I64, _I32, _I32, _F, _P, _P, _P, _P, _P]),
    "edn_place_samples": (C.c_int, [_P, _P, _P, _I64, _I32, _I32, _P, _P]),
    "edn_awp_workspace_floats": (C.c_int64, [_I64, _I32, _I32, C.POINTER(AwpOptions)]),
    "edn_awp_stats_offset_floats": (C.c_int64, [_I64, _I32, _I32]),
    "edn_awp_bwd_sums_offset_floats": (C.c_int64, [_I64, _I32, _I32]),
    "edn_awp_fwd": (C.c_int, [C.POINTER(AwpParams), _P, _P, _P, _I32, _P, _I64, _I32, _I32, _F, C.POINTER(AwpOptions), _P, _P, _P]),
    "edn_rbk_warp_ndc_fwd": (C.c_int, [C.POINTER(RbkParams), _P, _P, _I64, _I32, _I32, _F, _F, _F, _I32, _P, _P, _P, _P, _P]),
    "edn_build_ray_batch": (C.c_int, [_P, _I64, _I32, _I32, _F, _F, _F, _I32, _P, _P]),
    "edn_build_ray_batch_bwd": (C.c_int, [_P, _I64, _I32, _I32, _F, _I32, _P, _P, _P]),
    "edn_dsk_workspace_floats": (C.c_int64, [C.POINTER(DskParams), _I64]),
    "edn_dsk_rays_fwd": (C.c_int, [C.POINTER(DskParams), _P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _F, _F, _F, _P, _P, _P, _P, _P]),
    "edn_dsk_rays_bwd": (C.c_int, [C.POINTER(DskParams), _P, _P, _P, _P, _P, _I64, _I32, _I32, _F, _F, _F, _F, _P, _P, _P,
                                   C.POINTER(DskGrads), _P, _P]),
    "edn_weighted_sum": (C.c_int, [_P, _P, _P, _I64, _I32, _I64, _P]),
    "edn_crf_fwd": (C.c_int, [C.POINTER(CrfParams), _P, _P, _I32, _I32, _I64, _P, _P]),
    "edn_egm_loss_fwd": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I64, _F, _P, _P]),
    "edn_img2mse": (C.c_int, [_P, _P, _I64, _P, _P]),
    "edn_tv_loss_app": (C.c_int, [C.POINTER(C.c_void_p * 3), C.POINTER(C.c_void_p * 3), C.POINTER(C.c_int32 * 3),
                                  C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3), C.POINTER(C.c_int32 * 3), _P, _P, _P]),
    "edn_rays_from_pixels": (C.c_int, [_P, _P, _I32, _I64, C.c_double, C.c_double, C.c_double, C.c_double, _I32, _P, _P]),
    "edn_make_rgb_batch": (C.c_int, [_P, _I64, _P, _P, _I32, _I32, _I32, C.c_double, C.c_double, C.c_double, C.c_double, _P, _P, _P, _P, _P,
                                     _P, _P]),
    "edn_gather_successor": (C.c_int, [_P, _P, _I64, _P, _P, _I64, _P, _P, _P, _P]),
    "edn_interpolate_poses": (C.c_int, [_P, _I64, _P, _P, _I32, _P, _P, _I32, C.c_double, _P, _P, _P]),
    "edn_edi_prior": (C.c_int, [_P, _P, _P, _P, _P, _I32, _I64, _P, _I32, _I32, _I32, _F, _F, _P, _P, _P]),
}

_lib = None


class NativeLibraryError(RuntimeError):
    pass


def load():
    """Returns the loaded CDLL; raises NativeLibraryError when the CUDA library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: build it with `python evdeblurnerf_b200/csrc/build.py` "
            f"(or __graft_entry__.build()). evdeblurnerf_b200 has no CPU / PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(rc, what):
    if rc != 0:
        msg = load().edn_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a (contiguous) tensor or None."""
    if t is None:
        return None
    assert t.is_contiguous(), "edn: tensor must be contiguous"
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
