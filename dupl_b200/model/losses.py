"""Losses with the reference's interface (model/losses.py:6-39), fused CUDA forward + backward.

Each loss is one torch.autograd.Function around libdupl.so kernels, so it plugs into the reference
script's autograd graph (`loss.backward()`) without a PyTorch re-implementation of the arithmetic.
"""
import torch

from .. import _lib as L
from .. import ops


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


class _PtcLossSimt(torch.autograd.Function):
    """fp32 CUDA-core variant (any shape; also the A/B reference of the tensor-core path in tests)."""

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


class _CeSumUp(torch.autograd.Function):
    """sum over non-ignored pixels of CE(F.interpolate(pred, size=target.shape[1:], bilinear), target) / max(#pixels, 1)
    — the consistency term of train_final_voc.py:407-436 (`ce_criterion(segs_aug, pseudo_seg).sum() / uncertain_mask.sum()`,
    0 when the mask is empty) — on the fused up-sample + cross-entropy kernels: the [b, C, H, W] logits are never written and
    the gradient comes back at the low resolution (torch's upsample_bilinear2d_backward alone cost 4.8 ms per student)."""

    @staticmethod
    def forward(ctx, pred, target, ignore_index):
        L.require_cuda(pred, target)
        pred = L.f32c(pred.detach())
        target = target.to(torch.int64).contiguous()
        b, Cn, h, w = pred.shape
        H, W = target.shape[1:]
        dev = pred.device
        lse = torch.empty(b, H, W, dtype=torch.float32, device=dev)
        partials = torch.empty(4 * b * ((H + 31) // 32) * ((W + 31) // 32), dtype=torch.float32, device=dev)
        stats = torch.empty(5, dtype=torch.float32, device=dev)
        L.check(L.lib().dupl_seg_loss_up_fwd(L.ptr(pred), L.ptr(target), b, Cn, h, w, H, W, int(ignore_index), L.ptr(lse),
                                             L.ptr(partials), L.ptr(stats), L.stream_ptr(dev)), "dupl_seg_loss_up_fwd")
        # stats = [sum over label-0 pixels, sum over the other valid pixels, their counts, get_seg_loss value]
        count = (stats[2] + stats[3]).clamp_min(1.0)
        # the backward kernel weights a pixel by 0.5 / (class-group count + 1e-6): counts that give every valid pixel 1 / count
        bstats = stats.clone()
        bstats[2:4] = count / 2 - 1e-6
        ctx.save_for_backward(pred, target, lse, bstats)
        ctx.ignore_index = int(ignore_index)
        return (stats[0] + stats[1]) / count

    @staticmethod
    def backward(ctx, grad_out):
        pred, target, lse, bstats = ctx.saved_tensors
        b, Cn, h, w = pred.shape
        H, W = target.shape[1:]
        dpred = torch.empty_like(pred)
        g = L.f32c(grad_out).reshape(1)
        L.check(L.lib().dupl_seg_loss_up_bwd(L.ptr(pred), L.ptr(target), L.ptr(lse), L.ptr(bstats), L.ptr(g), b, Cn, h, w, H, W,
                                             ctx.ignore_index, L.ptr(dpred), L.stream_ptr(pred.device)), "dupl_seg_loss_up_bwd")
        return dpred, None, None


def ce_sum_upsampled(pred_lowres, target, ignore_index=255):
    """CE(F.interpolate(pred_lowres -> target size), target).sum() / max(number of non-ignored pixels, 1)."""
    return _CeSumUp.apply(pred_lowres, target, ignore_index)


def _pad64(n):
    return (n + 63) // 64 * 64


class _PtcLoss(torch.autograd.Function):
    """get_masked_ptc_loss with both contractions on the tensor cores (split-bf16 GEMM): forward G_i = x_hat_i x_hat_i^T per
    image, backward dX_hat_i = (S_i + S_i^T) x_hat_i.  Shapes the GEMM cannot take (h*w not a multiple of 16, channels not a
    multiple of 64) use the fp32 CUDA-core kernels (_PtcLossSimt)."""

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
        npad = _pad64(n)
        bf = dict(dtype=torch.bfloat16, device=dev)
        inv = torch.empty(b, n, dtype=torch.float32, device=dev)
        rows = (torch.empty(b * n, c, **bf), torch.empty(b * n, c, **bf))
        cm = (torch.empty(b, c, npad, **bf), torch.empty(b, c, npad, **bf))
        st = L.stream_ptr(dev)
        L.check(L.lib().dupl_ptc_prepare(L.ptr(x), b, c, n, npad, L.ptr(inv), L.ptr(rows[0]), L.ptr(rows[1]), L.ptr(cm[0]),
                                         L.ptr(cm[1]), st), "dupl_ptc_prepare")
        G = torch.empty(b, n, n, dtype=torch.float32, device=dev)
        for i in range(0, b, L.MAX_GROUPS):
            ops.gemm_bf16x3([dict(a=(rows[0][j * n:(j + 1) * n], rows[1][j * n:(j + 1) * n]),
                                  w=(rows[0][j * n:(j + 1) * n], rows[1][j * n:(j + 1) * n]), out_f32=G[j])
                             for j in range(i, min(i + L.MAX_GROUPS, b))], n, n, c, L.EPI_F32)
        nblocks = 592
        partials = torch.empty(4 * nblocks, dtype=torch.float32, device=dev)
        stats = torch.empty(5, dtype=torch.float32, device=dev)
        L.check(L.lib().dupl_ptc_mask_reduce(L.ptr(G), L.ptr(mask), L.ptr(x), L.ptr(inv), b, c, n, L.ptr(partials), nblocks, L.ptr(stats), st),
                "dupl_ptc_mask_reduce")
        ctx.save_for_backward(x, mask, inv, G, stats, cm[0], cm[1])
        return stats[4].clone()

    @staticmethod
    def backward(ctx, grad_out):
        x, mask, inv, G, stats, cm_hi, cm_lo = ctx.saved_tensors
        b, c, h, w = x.shape
        n = h * w
        npad = cm_hi.shape[2]
        dev = x.device
        st = L.stream_ptr(dev)
        g = L.f32c(grad_out).reshape(1)
        bf = dict(dtype=torch.bfloat16, device=dev)
        t = (torch.empty(b * n, npad, **bf), torch.empty(b * n, npad, **bf))
        L.check(L.lib().dupl_ptc_dg(L.ptr(G), L.ptr(mask), L.ptr(stats), L.ptr(g), b, n, npad, L.ptr(t[0]), L.ptr(t[1]), st),
                "dupl_ptc_dg")
        dxh = torch.empty(b * n, c, dtype=torch.float32, device=dev)
        for i in range(0, b, L.MAX_GROUPS):
            ops.gemm_bf16x3([dict(a=(t[0][j * n:(j + 1) * n], t[1][j * n:(j + 1) * n]), w=(cm_hi[j], cm_lo[j]),
                                  out_f32=dxh[j * n:(j + 1) * n]) for j in range(i, min(i + L.MAX_GROUPS, b))],
                            n, c, npad, L.EPI_F32)
        dx = torch.empty_like(x)
        L.check(L.lib().dupl_ptc_norm_bwd_rows(L.ptr(x), L.ptr(inv), L.ptr(dxh), b, c, n, L.ptr(dx), st), "dupl_ptc_norm_bwd_rows")
        return dx, None


def get_masked_ptc_loss(inputs, mask):
    """model/losses.py:6-21"""
    n = inputs.shape[2] * inputs.shape[3]
    if n % 16 == 0 and inputs.shape[1] % 64 == 0:
        return _PtcLoss.apply(inputs, mask)
    return _PtcLossSimt.apply(inputs, mask)


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


# ---------------------------------------------------------------------------------------------------------------------
# The two losses the scripts compute inline with torch ops (train_final_voc.py:299-305 and :440-447), fused
# ---------------------------------------------------------------------------------------------------------------------
def _ptr_table(tensors):
    """HOST array of the tensors' device pointers (passed by value into the kernel parameters: capturable)."""
    import ctypes as C
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


class _ClsLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_label, *logits):
        L.require_cuda(cls_label, *logits)
        xs = [L.f32c(x.detach()) for x in logits]
        y = L.f32c(cls_label.detach())
        if any(x.shape != y.shape for x in xs):
            raise ValueError("multilabel_soft_margin_sum: every logits tensor must have the shape of cls_label")
        dev = y.device
        loss = torch.empty((), dtype=torch.float32, device=dev)
        tab = _ptr_table(xs)
        L.check(L.lib().dupl_cls_loss_fwd(tab, len(xs), L.ptr(y), y.numel(), L.ptr(loss), L.stream_ptr(dev)), "dupl_cls_loss_fwd")
        ctx.xs, ctx.y, ctx.tab = xs, y, tab
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        xs, y = ctx.xs, ctx.y
        grads = [torch.empty_like(x) for x in xs]
        gtab = _ptr_table(grads)
        g = L.f32c(grad_out).reshape(1)
        L.check(L.lib().dupl_cls_loss_bwd(ctx.tab, gtab, len(xs), L.ptr(y), y.numel(), L.ptr(g), L.stream_ptr(y.device)),
                "dupl_cls_loss_bwd")
        return (None, *grads)


def multilabel_soft_margin_sum(logits, cls_label):
    """sum_t F.multilabel_soft_margin_loss(logits[t], cls_label) for the 4 classification heads of the two students
    (train_final_voc.py:299-305) in one kernel (+ one for the gradients)."""
    return _ClsLoss.apply(cls_label, *logits)


class _SimLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1, f2, eps):
        L.require_cuda(f1, f2)
        a, b = L.f32c(f1.detach()), L.f32c(f2.detach())
        if a.shape != b.shape or a.dim() != 4:
            raise ValueError("discrepancy_loss expects two [b, C, h, w] feature maps of the same shape")
        rows, n = a.shape[0] * a.shape[1], a.shape[2] * a.shape[3]
        dev = a.device
        cos_rows = torch.empty(rows, dtype=torch.float32, device=dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        L.check(L.lib().dupl_sim_loss_fwd(L.ptr(a), L.ptr(b), rows, n, float(eps), L.ptr(cos_rows), L.ptr(loss), L.stream_ptr(dev)),
                "dupl_sim_loss_fwd")
        ctx.a, ctx.b, ctx.eps = a, b, float(eps)
        return loss

    @staticmethod
    def backward(ctx, grad_out):
        a, b = ctx.a, ctx.b
        d1, d2 = torch.empty_like(a), torch.empty_like(b)
        g = L.f32c(grad_out).reshape(1)
        L.check(L.lib().dupl_sim_loss_bwd(L.ptr(a), L.ptr(b), a.shape[0] * a.shape[1], a.shape[2] * a.shape[3], ctx.eps, L.ptr(g),
                                          L.ptr(d1), L.ptr(d2), L.stream_ptr(a.device)), "dupl_sim_loss_bwd")
        return d1, d2, None


def discrepancy_loss(fmap_1, fmap_2, eps=1e-6):
    """(1 + cos(fmap_1.detach(), fmap_2).mean()) + (1 + cos(fmap_2.detach(), fmap_1).mean()) with
    nn.CosineSimilarity(dim=-1, eps) over the flattened spatial axis (train_final_voc.py:440-447), fused."""
    return _SimLoss.apply(fmap_1, fmap_2, eps)
