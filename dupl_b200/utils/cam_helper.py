"""CAM / pseudo-label helpers with the reference's signatures (utils/cam_helper.py).

Only the functions on the hot path are provided: cam_to_label,
cam_to_label_dynamic_cls, get_valid_cam, ignore_img_box, multi_scale_cam2, multi_scale_cam2_siamese,
label_to_aff_mask, refine_cams_with_bkg_v2, refine_cams_with_dynamic_thres, _refine_cams.
"""
import torch

from .. import _lib as L
from .. import ops


# ----------------------------------------------------------------------------- cam_to_label
def cam_to_label(cam, cls_label, img_box=None, bkg_thre=None, high_thre=None, low_thre=None, ignore_mid=False,
                 ignore_index=None):
    """utils/cam_helper.py:8-30.  Returns int64 labels [b,h,w] when img_box is None, else
    (valid_cam, pseudo_label)."""
    valid, label = ops.cam_to_label(cam, cls_label, img_box, bkg_thre, high_thre, low_thre, ignore_mid, ignore_index,
                                    want_valid=img_box is not None)
    if img_box is None:
        return label
    return valid, label


def cam_to_label_dynamic_cls(cam, cls_label, img_box=None, bkg_thre=None, high_thre=None, low_thre=None,
                             ignore_mid=False, ignore_index=None):
    """utils/cam_helper.py:33-55: `high_thre` is a tensor with one threshold per image."""
    return cam_to_label(cam, cls_label, img_box, bkg_thre, high_thre, low_thre, ignore_mid, ignore_index)


def get_valid_cam(cam, cls_label):
    """utils/cam_helper.py:72-78"""
    valid, _ = ops.cam_to_label(cam, cls_label, None, 0.0, None, None, False, 0, want_valid=True)
    return valid


def label_to_aff_mask(cam_label, ignore_index=255):
    """utils/cam_helper.py:323-335"""
    return ops.label_to_aff_mask(cam_label, ignore_index)


# ----------------------------------------------------------------------------- multi-scale CAM
def _unwrap(model):
    return model.module if hasattr(model, "module") else model


def _scale_order(scales):
    return [1.0] + [s for s in scales if s != 1.0]  # scale 1.0 always first (cam_helper.py:168-181)


def _fused_mscam(nets, inputs, scales):
    """All scales (and flipped twins) of all given students through the encoder in one pass."""
    from ..model.model_dupl import cam_only_forward
    inputs = L.f32c(inputs)
    b, _, h, w = inputs.shape
    sizes = [(int(s * h), int(s * w)) for s in _scale_order(scales)]  # cam_helper.py:183
    # The un-flipped scale-1.0 half of this pass IS the training forward of `inputs` (no dropout / batch statistics in
    # the model): a fused step driver may ask to keep its activations (network._keep_next) and skip that forward.
    keep = b if all(getattr(n, "_keep_next", False) for n in nets) else 0
    res = cam_only_forward(nets, inputs, seg_images=[inputs] * len(sizes), seg_sizes=sizes, flip_twin=True, keep_batch=keep)
    return [(ops.mscam_post(cams, b, h, w), ops.mscam_post(aux, b, h, w)) for (aux, cams) in res]


def multi_scale_cam2_siamese(model, inputs, scales, branch=1):
    """utils/cam_helper.py:164-204 -> (cam, cam_aux), each [b,K,h,w] in [0,1)."""
    from ..model.model_dupl import siamese_network
    net = _unwrap(model)
    if not isinstance(net, siamese_network):
        raise TypeError("multi_scale_cam2_siamese expects a dupl_b200 siamese_network (optionally DDP-wrapped)")
    return _fused_mscam([net.branch1 if branch == 1 else net.branch2], inputs, scales)[0]


def multi_scale_cam2_pair(model, inputs, scales):
    """Both students in one pass (one grouped GEMM per layer): ((cam1, aux1), (cam2, aux2))."""
    net = _unwrap(model)
    return _fused_mscam([net.branch1, net.branch2], inputs, scales)


def multi_scale_cam2(model, inputs, scales):
    """utils/cam_helper.py:121-161 for a single `network`."""
    return _fused_mscam([_unwrap(model)], inputs, scales)[0]


# ----------------------------------------------------------------------------- refine
def _refine(ref_mod, images, cams, cls_labels, bkg_h, low_thre, ignore_index, img_box, down_scale):
    if down_scale != 2:
        raise NotImplementedError("refine_cams_*: only down_scale=2 (the reference's only call) is built")
    L.require_cuda(images, cams)
    b, _, h, w = images.shape
    images_ds, masks, nactive, cls = ops.refine_prologue(images, cams, cls_labels, bkg_h, low_thre)
    aff = ops.par_affinity(images_ds, ref_mod.dilations, ref_mod.w1, ref_mod.w2)
    masks = ops.par_propagate(aff, masks, ref_mod.dilations, ref_mod.num_iter, nactive=nactive)
    return ops.refine_epilogue(masks, cls, img_box, h, w, ignore_index)


def refine_cams_shared_affinity(ref_mod, images, cams_list, cls_labels, bkg_h, low_thre, ignore_index, img_box):
    """refine_cams_with_dynamic_thres / _bkg_v2 for several CAM tensors over the SAME images (the two students of
    train_final_voc.py:330-343): PAR's affinity depends on the images only (PAR.py:68-87), so it is computed once
    instead of once per call (4x per image in the reference).  Returns one label map per entry."""
    L.require_cuda(images, *cams_list)
    b, _, h, w = images.shape
    aff = None
    out = []
    for cams in cams_list:
        images_ds, masks, nactive, cls = ops.refine_prologue(images, cams, cls_labels, bkg_h, low_thre)
        if aff is None:
            aff = ops.par_affinity(images_ds, ref_mod.dilations, ref_mod.w1, ref_mod.w2)
        masks = ops.par_propagate(aff, masks, ref_mod.dilations, ref_mod.num_iter, nactive=nactive)
        out.append(ops.refine_epilogue(masks, cls, img_box, h, w, ignore_index))
    return out


def refine_cams_with_bkg_v2(ref_mod=None, images=None, cams=None, cls_labels=None, high_thre=None, low_thre=None,
                            ignore_index=False, img_box=None, down_scale=2):
    """utils/cam_helper.py:338-383 (scalar high threshold) -> float32 [b,h,w] in {0..K, ignore}."""
    return _refine(ref_mod, images, cams, cls_labels, float(high_thre), low_thre, ignore_index, img_box, down_scale)


def refine_cams_with_dynamic_thres(ref_mod=None, images=None, cams=None, cls_labels=None, high_thre_map=None,
                                   low_thre=None, ignore_index=False, img_box=None, down_scale=2):
    """utils/cam_helper.py:386-431 (per-pixel high-threshold map [b,1,h,w])."""
    return _refine(ref_mod, images, cams, cls_labels, high_thre_map, low_thre, ignore_index, img_box, down_scale)


def _refine_cams(ref_mod, images, cams, valid_key, orig_size):
    """utils/cam_helper.py:434-440 for already gathered+soft-maxed `cams` [1,n,h,w]."""
    L.require_cuda(images, cams)
    n = cams.shape[1]
    refined = ref_mod(images, cams)
    # arg-max of the up-sampled maps via the refine epilogue: every channel present, identity box
    b = refined.shape[0]
    H, W = orig_size
    if (H, W) != (2 * refined.shape[2], 2 * refined.shape[3]):
        raise NotImplementedError("_refine_cams: orig_size must be twice the refined size")
    masks = torch.zeros(b, 2 * n, refined.shape[2], refined.shape[3], dtype=torch.float32, device=refined.device)
    masks[:, :n] = refined
    masks[:, n:] = refined
    cls = torch.ones(b, n - 1, dtype=torch.float32, device=refined.device)
    box = torch.tensor([[0, H, 0, W]] * b, dtype=torch.int32)
    _, lh, _ = ops.refine_epilogue(masks, cls, box, H, W, -1.0, want_parts=True)
    return valid_key[lh.long()]
