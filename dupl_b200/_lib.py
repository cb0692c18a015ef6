"""ctypes binding of libdupl.so (include/dupl.h).

The product path has NO fallback: if the CUDA library is missing or a call fails, a
RuntimeError is raised.  PyTorch is used for storage only (device buffers, streams).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# DUPL_LIB: an alternative build of the same library (tools/precision_table.py loads the TF32-emulation experiment builds)
LIB_PATH = os.environ.get("DUPL_LIB") or os.path.join(_HERE, "libdupl.so")

MAX_SEGMENTS = 8
MAX_GROUPS = 2
MAX_KSPLIT = 8
PAR_MAX_DIL = 8

EPI_F32, EPI_SPLIT, EPI_GELU_SPLIT, EPI_RESID, EPI_PATCH, EPI_RELU_SPLIT = 0, 1, 2, 3, 4, 5

c_f32p = C.POINTER(C.c_float)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)


class Segment(C.Structure):
    _fields_ = [("batch", C.c_int32), ("gh", C.c_int32), ("gw", C.c_int32), ("tokens", C.c_int32),
                ("row_offset", C.c_int32), ("patch_row_offset", C.c_int32)]


class GemmGroup(C.Structure):
    _fields_ = [("a_hi", C.c_void_p), ("a_lo", C.c_void_p), ("w_hi", C.c_void_p), ("w_lo", C.c_void_p),
                ("bias", C.c_void_p), ("resid", C.c_void_p), ("out_f32", C.c_void_p),
                ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("pos", C.c_void_p * MAX_SEGMENTS),
                ("splitk_ws", C.c_void_p)]


class GemmArgs(C.Structure):
    _fields_ = [("groups", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
                ("lda", C.c_int32), ("ldo", C.c_int32), ("epilogue", C.c_int32), ("nseg", C.c_int32),
                ("ldw", C.c_int32), ("max_ksplit", C.c_int32), ("f32_rows", C.c_int32), ("passes", C.c_int32),
                ("a_mn_major", C.c_int32), ("b_mn_major", C.c_int32),
                ("seg", Segment * MAX_SEGMENTS), ("g", GemmGroup * MAX_GROUPS)]


class AttentionArgs(C.Structure):
    _fields_ = [("nseg", C.c_int32), ("seg", Segment * MAX_SEGMENTS), ("M", C.c_int32), ("heads", C.c_int32),
                ("scale", C.c_float), ("qkv_hi", C.c_void_p), ("qkv_lo", C.c_void_p),
                ("out_hi", C.c_void_p), ("out_lo", C.c_void_p), ("lse", C.c_void_p)]


class AttentionBwdArgs(C.Structure):
    _fields_ = [("qkv_hi", C.c_void_p), ("qkv_lo", C.c_void_p), ("o_hi", C.c_void_p), ("o_lo", C.c_void_p),
                ("do_hi", C.c_void_p), ("do_lo", C.c_void_p), ("lse", C.c_void_p), ("Dvec", C.c_void_p), ("dqkv", C.c_void_p),
                ("M", C.c_int32), ("batch", C.c_int32), ("tokens", C.c_int32), ("row_offset", C.c_int32), ("heads", C.c_int32),
                ("scale", C.c_float)]


class MscamArgs(C.Structure):
    _fields_ = [("nscale", C.c_int32), ("lowres", C.c_void_p * MAX_SEGMENTS),
                ("gh", C.c_int32 * MAX_SEGMENTS), ("gw", C.c_int32 * MAX_SEGMENTS),
                ("b", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("out", C.c_void_p), ("minmax", C.c_void_p)]


class CamToLabelArgs(C.Structure):
    _fields_ = [("cam", C.c_void_p), ("cls_label", C.c_void_p), ("img_box", C.c_void_p), ("high_thre", C.c_void_p),
                ("high_thre_scalar", C.c_float), ("low_thre", C.c_float), ("bkg_thre", C.c_float),
                ("ignore_mid", C.c_int32), ("ignore_index", C.c_int64),
                ("b", C.c_int32), ("K", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
                ("valid_cam", C.c_void_p), ("label", C.c_void_p)]


class RefinePrologueArgs(C.Structure):
    _fields_ = [("images", C.c_void_p), ("cams", C.c_void_p), ("cls_label", C.c_void_p), ("bkg_h", C.c_void_p),
                ("bkg_h_scalar", C.c_float), ("bkg_l_scalar", C.c_float),
                ("b", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("images_ds", C.c_void_p), ("masks", C.c_void_p), ("nactive", C.c_void_p)]


class RefineEpilogueArgs(C.Structure):
    _fields_ = [("masks", C.c_void_p), ("cls_label", C.c_void_p), ("img_box", C.c_void_p),
                ("b", C.c_int32), ("K", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("ignore_index", C.c_float), ("label", C.c_void_p), ("label_h", C.c_void_p), ("label_l", C.c_void_p)]


class CrfArgs(C.Structure):
    _fields_ = [("W", C.c_int32), ("H", C.c_int32), ("C", C.c_int32),
                ("pos_w", C.c_float), ("pos_xy_std", C.c_float), ("bi_w", C.c_float), ("bi_xy_std", C.c_float),
                ("bi_rgb_std", C.c_float), ("iters", C.c_int32), ("input_is_energy", C.c_int32),
                ("image", C.c_void_p), ("unary_or_probs", C.c_void_p), ("out", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("values", C.c_void_p), ("values_bytes", C.c_size_t), ("meta", C.c_void_p)]


MAX_TRANSPOSE_ITEMS = 16


class TransposeItem(C.Structure):
    _fields_ = [("in_hi", C.c_void_p), ("in_lo", C.c_void_p), ("out_hi", C.c_void_p), ("out_lo", C.c_void_p),
                ("R", C.c_int32), ("Cc", C.c_int32), ("ld", C.c_int32), ("Rpad", C.c_int32)]


class AdamwParam(C.Structure):
    _fields_ = [("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
                ("plane_hi", C.c_void_p), ("plane_lo", C.c_void_p), ("numel", C.c_int64), ("lr", C.c_float), ("reserved", C.c_int32)]


class AdamwArgs(C.Structure):
    _fields_ = [("params", C.c_void_p), ("items", C.c_void_p), ("active", C.c_void_p), ("steps", C.c_void_p), ("coef", C.c_void_p),
                ("lr_scale", C.c_void_p), ("n_params", C.c_int32), ("n_items", C.c_int64),
                ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float)]


# name -> (restype, argtypes); every symbol declared in include/dupl.h
_PROTOTYPES = {
    "dupl_version": (C.c_int, []),
    "dupl_last_error": (C.c_char_p, []),
    "dupl_launch_count": (C.c_int64, []),
    "dupl_split_bf16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "dupl_gemm_bf16x3": (C.c_int, [C.POINTER(GemmArgs), C.c_void_p]),
    "dupl_layernorm_split": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    "dupl_im2col3x3": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int32] * 8 + [C.c_void_p]),
    "dupl_rows_to_nchw": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 7 + [C.c_void_p]),
    "dupl_gmp_classify": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] * 7 + [C.c_void_p]),
    "dupl_attention_fwd": (C.c_int, [C.POINTER(AttentionArgs), C.c_void_p]),
    "dupl_patchify": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(Segment), C.c_int32, C.c_int32,
                                C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dupl_pos_embed_resize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "dupl_cls_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(Segment), C.c_int32,
                                C.c_int32, C.c_void_p]),
    "dupl_cam_contract": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int32, C.c_int32,
                                    C.POINTER(Segment), C.c_int32, C.c_void_p, c_i64p, C.c_void_p]),
    "dupl_mscam_post": (C.c_int, [C.POINTER(MscamArgs), C.c_void_p]),
    "dupl_cam_to_label": (C.c_int, [C.POINTER(CamToLabelArgs), C.c_void_p]),
    "dupl_label_to_aff_mask": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p]),
    "dupl_par_affinity": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, c_i32p,
                                    C.c_int32, C.c_float, C.c_float, C.c_void_p]),
    "dupl_par_propagate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_int32, c_i32p, C.c_int32, C.c_int32, c_i32p, C.c_void_p]),
    "dupl_refine_prologue": (C.c_int, [C.POINTER(RefinePrologueArgs), C.c_void_p]),
    "dupl_refine_epilogue": (C.c_int, [C.POINTER(RefineEpilogueArgs), C.c_void_p]),
    "dupl_split_transpose": (C.c_int, [C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 3),
    "dupl_split_transpose_gelu": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 4 + [C.c_int32] + [C.c_void_p] * 3),
    "dupl_transpose_planes": (C.c_int, [C.c_void_p] * 2 + [C.c_int32] * 6 + [C.c_void_p] * 2 + [C.c_int32, C.c_void_p]),
    "dupl_transpose_planes_multi": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "dupl_colsum": (C.c_int, [C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p, C.c_void_p]),
    "dupl_layernorm_bwd": (C.c_int, [C.c_void_p] * 7 + [C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    "dupl_gelu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "dupl_relu_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "dupl_col2im3x3": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 9 + [C.c_void_p]),
    "dupl_nchw_to_rows_add": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 6 + [C.c_void_p]),
    "dupl_gmp_classify_bwd": (C.c_int, [C.c_void_p] * 7 + [C.c_int32] * 7 + [C.c_void_p]),
    "dupl_attention_bwd": (C.c_int, [C.POINTER(AttentionBwdArgs), C.c_void_p]),
    "dupl_seg_loss_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int64,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dupl_seg_loss_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] * 4 + [C.c_int64, C.c_void_p, C.c_void_p]),
    "dupl_seg_loss_up_fwd": (C.c_int, [C.c_void_p, C.c_void_p] + [C.c_int32] * 6 + [C.c_int64] + [C.c_void_p] * 4),
    "dupl_seg_loss_up_bwd": (C.c_int, [C.c_void_p] * 5 + [C.c_int32] * 6 + [C.c_int64, C.c_void_p, C.c_void_p]),
    "dupl_ptc_loss_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 5),
    "dupl_ptc_loss_bwd": (C.c_int, [C.c_void_p] * 6 + [C.c_int32] * 3 + [C.c_void_p] * 3),
    "dupl_gemm_plan": (C.c_int, [C.c_int32] * 5 + [C.POINTER(C.c_int32)] * 3),
    "dupl_set_gemm_sm_limit": (C.c_int, [C.c_int32]),
    "dupl_ptc_prepare": (C.c_int, [C.c_void_p] + [C.c_int32] * 4 + [C.c_void_p] * 6),
    "dupl_ptc_mask_reduce": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "dupl_ptc_dg": (C.c_int, [C.c_void_p] * 4 + [C.c_int32] * 3 + [C.c_void_p] * 3),
    "dupl_ptc_norm_bwd_rows": (C.c_int, [C.c_void_p] * 3 + [C.c_int32] * 3 + [C.c_void_p] * 2),
    "dupl_cls_loss_fwd": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "dupl_cls_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "dupl_sim_loss_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "dupl_sim_loss_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p]),
    "dupl_randaug_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "dupl_randaug": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, c_f32p, C.c_void_p,
                               C.c_size_t, C.c_void_p]),
    "dupl_adamw_items": (C.c_int, [c_i64p, C.c_int32, c_i32p, C.c_int64, c_i64p]),
    "dupl_adamw_step": (C.c_int, [C.POINTER(AdamwArgs), C.c_void_p]),
    "dupl_gmm_filter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_float,
                                  C.c_float, C.c_float, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    "dupl_crf_workspace_bytes": (C.c_int, [C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "dupl_crf_values_bytes": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_size_t)]),
    "dupl_crf_build": (C.c_int, [C.POINTER(CrfArgs), C.c_void_p]),
    "dupl_crf_infer": (C.c_int, [C.POINTER(CrfArgs), C.c_int32, C.c_int32, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOTYPES)

_lib = None


def lib():
    """Loads libdupl.so once; raises if it has not been built (python -c 'import __graft_entry__ as g; g.build()')."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: the CUDA library has not been built. Run "
                f"`python -c 'import __graft_entry__ as g; g.build()'` at the repo root. "
                f"dupl_b200 has no CPU or PyTorch fallback.")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().dupl_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else '?'}")


def stream_ptr(device=None):
    """Raw cudaStream_t of torch's current stream on `device` (fast path: no Stream object is built)."""
    idx = device.index if isinstance(device, torch.device) and device.index is not None else torch.cuda.current_device()
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(idx))


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("dupl_b200 runs on CUDA tensors only (no CPU fallback); got a tensor on " + str(t.device))


def f32c(t):
    """contiguous fp32 view/copy (storage plumbing only)."""
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()
