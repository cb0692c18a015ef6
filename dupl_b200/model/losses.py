"""Losses with the reference's interface (model/losses.py:6-39), fused CUDA forward + backward.

Each loss is one torch.autograd.Function around libdupl.so kernels, so it plugs into the reference
script's autograd graph (`loss.backward()`) without a PyTorch re-implementation of the arithmetic.
"""
import torch

from .. import _lib as L


class _SegLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, label, ignore_index):
        L.require_cuda(pred, label)
        pred = L.f32c(pred.detach())
        label = label.to(torch.int64).contiguous()
        b, Cn, H, W = pred.shape
        dev = pred.device
        lse = torch.empty(b, H, W, dtype=torch.float32, device=dev)
        partials = torch.empty(4 * ((b * H * W + 255) // 256), dtype=torch.float32, device=dev)
        stats = torch.empty(5, dtype=torch.float32, device=dev)
        L.check(L.lib().dupl_seg_loss_fwd(L.ptr(pred), L.ptr(label), b, Cn, H, W, int(ignore_index), L.ptr(lse),
                                          L.ptr(partials), L.ptr(stats), L.stream_ptr(dev)), "dupl_seg_loss_fwd")
        ctx.save_for_backward(pred, label, lse, stats)
        ctx.ignore_index = int(ignore_index)
        return stats[4].clone()

    @staticmethod
    def backward(ctx, grad_out):
        pred, label, lse, stats = ctx.saved_tensors
        b, Cn, H, W = pred.shape
        dpred = torch.empty_like(pred)
        g = L.f32c(grad_out).reshape(1)
        L.check(L.lib().dupl_seg_loss_bwd(L.ptr(pred), L.ptr(label), L.ptr(lse), L.ptr(stats), L.ptr(g), b, Cn, H, W,
                                          ctx.ignore_index, L.ptr(dpred), L.stream_ptr(pred.device)), "dupl_seg_loss_bwd")
        return dpred, None, None


class _SegLossUp(torch.autograd.Function):
    """get_seg_loss(F.interpolate(pred, size=label.shape[1:], mode='bilinear', align_corners=False), label) in one
    pass over the label pixels; the gradient comes back at the low resolution of `pred`."""

    @staticmethod
    def forward(ctx, pred, label, ignore_index):
        L.require_cuda(pred, label)
        pred = L.f32c(pred.detach())
        label = label.to(torch.int64).contiguous()
        b, Cn, h, w = pred.shape
        H, W = label.shape[1:]
        dev = pred.device
        lse = torch.empty(b, H, W, dtype=torch.float32, device=dev)
        partials = torch.empty(4 * b * ((H + 31) // 32) * ((W + 31) // 32), dtype=torch.float32, device=dev)
        stats = torch.empty(5, dtype=torch.float32, device=dev)
        L.check(L.lib().dupl_seg_loss_up_fwd(L.ptr(pred), L.ptr(label), b, Cn, h, w, H, W, int(ignore_index), L.ptr(lse),
                                             L.ptr(partials), L.ptr(stats), L.stream_ptr(dev)), "dupl_seg_loss_up_fwd")
        ctx.save_for_backward(pred, label, lse, stats)
        ctx.ignore_index = int(ignore_index)
        return stats[4].clone()

    @staticmethod
    def backward(ctx, grad_out):
        pred, label, lse, stats = ctx.saved_tensors
        b, Cn, h, w = pred.shape
        H, W = label.shape[1:]
        dpred = torch.empty_like(pred)
        g = L.f32c(grad_out).reshape(1)
        L.check(L.lib().dupl_seg_loss_up_bwd(L.ptr(pred), L.ptr(label), L.ptr(lse), L.ptr(stats), L.ptr(g), b, Cn, h, w, H, W,
                                             ctx.ignore_index, L.ptr(dpred), L.stream_ptr(pred.device)), "dupl_seg_loss_up_bwd")
        return dpred, None, None


class _PtcLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, mask):
        L.require_cuda(inputs, mask)
        x = L.f32c(inputs.detach())
        b, c, h, w = x.shape
        n = h * w
        mask = mask.to(torch.int64).contiguous()
        if tuple(mask.shape) != (b, n, n):
            raise ValueError(f"mask must be [b, h*w, h*w] = {(b, n, n)}, got {tuple(mask.shape)}")
        dev = x.device
        inv = torch.empty(b, n, dtype=torch.float32, device=dev)
        Gs = torch.empty(b, n, n, dtype=torch.float32, device=dev)
        tiles = (n + 63) // 64
        partials = torch.empty(4 * b * tiles * tiles, dtype=torch.float32, device=dev)
        stats = torch.empty(5, dtype=torch.float32, device=dev)
        L.check(L.lib().dupl_ptc_loss_fwd(L.ptr(x), L.ptr(mask), b, c, n, L.ptr(inv), L.ptr(Gs), L.ptr(partials), L.ptr(stats),
                                          L.stream_ptr(dev)), "dupl_ptc_loss_fwd")
        ctx.save_for_backward(x, mask, inv, Gs, stats)
        return stats[4].clone()

    @staticmethod
    def backward(ctx, grad_out):
        x, mask, inv, Gs, stats = ctx.saved_tensors
        b, c, h, w = x.shape
        dxh = torch.empty_like(x)
        dx = torch.empty_like(x)
        g = L.f32c(grad_out).reshape(1)
        L.check(L.lib().dupl_ptc_loss_bwd(L.ptr(x), L.ptr(mask), L.ptr(inv), L.ptr(Gs), L.ptr(stats), L.ptr(g), b, c, h * w,
                                          L.ptr(dxh), L.ptr(dx), L.stream_ptr(x.device)), "dupl_ptc_loss_bwd")
        return dx, None


def get_masked_ptc_loss(inputs, mask):
    """model/losses.py:6-21"""
    return _PtcLoss.apply(inputs, mask)


def get_seg_loss(pred, label, ignore_index=255):
    """model/losses.py:24-39"""
    return _SegLoss.apply(pred, label, ignore_index)


def get_seg_loss_upsampled(pred_lowres, label, ignore_index=255):
    """get_seg_loss(F.interpolate(pred_lowres, size=label.shape[1:], mode='bilinear', align_corners=False), label)
    — the pair of calls at train_final_voc.py:345-352 — fused: the [b, C, H, W] logits are never written."""
    return _SegLossUp.apply(pred_lowres, label, ignore_index)


def get_seg_loss_conflict_v2(*args, **kwargs):
    """train_final_coco.py:21 imports this name but the reference never defines or calls it."""
    raise NotImplementedError("get_seg_loss_conflict_v2 does not exist in the reference's model/losses.py either")
