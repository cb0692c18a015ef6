"""GPU: fused loss kernels (dupl_b200.model.losses) vs the oracle's autograd on CPU — values and gradients."""
import os

import numpy as np
import pytest
import torch

from helpers import rel_err

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _labels(b, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    lab = torch.randint(0, 6, (b, H, W), generator=g)
    lab[lab == 5] = 255
    lab[lab == 4] = 0
    return lab


@pytest.mark.parametrize("b,C,H,W", [(2, 21, 32, 48), (1, 81, 17, 23), (3, 5, 64, 64)])
def test_seg_loss_value_and_gradient(b, C, H, W):
    from dupl_b200.model.losses import get_seg_loss
    from oracle import dupl_oracle as O
    g = torch.Generator().manual_seed(b * C)
    pred = torch.randn(b, C, H, W, generator=g) * 2
    lab = _labels(b, H, W, seed=H)
    p_ref = pred.clone().requires_grad_(True)
    want = O.seg_loss(p_ref, lab)
    (want * 1.7).backward()
    p_gpu = pred.cuda().requires_grad_(True)
    got = get_seg_loss(p_gpu, lab.cuda(), ignore_index=255)
    (got * 1.7).backward()
    assert abs(got.item() - want.item()) < 1e-5 * max(1.0, abs(want.item()))
    assert rel_err(p_gpu.grad, p_ref.grad) < 1e-4


@pytest.mark.parametrize("b,C,h,w,H,W", [(2, 21, 28, 28, 448, 448), (1, 81, 5, 7, 37, 50), (2, 21, 21, 21, 448, 448), (1, 4, 6, 6, 6, 6),
                                         (1, 3, 9, 8, 5, 4)])
def test_seg_loss_upsampled_equals_interpolate_then_seg_loss(b, C, h, w, H, W):
    """get_seg_loss_upsampled == get_seg_loss(F.interpolate(..., bilinear, align_corners=False)) of the oracle (CPU autograd),
    value and the gradient at the low resolution; ratios 16x, ragged, 21.3x (the 0.75 view), 1x and a down-sampling."""
    import torch.nn.functional as F
    from dupl_b200.model.losses import get_seg_loss_upsampled
    from oracle import dupl_oracle as O
    g = torch.Generator().manual_seed(b * C + h)
    pred = torch.randn(b, C, h, w, generator=g) * 2
    lab = _labels(b, H, W, seed=H + w)
    p_ref = pred.clone().requires_grad_(True)
    want = O.seg_loss(F.interpolate(p_ref, size=(H, W), mode="bilinear", align_corners=False), lab)
    (want * 0.2).backward()
    p_gpu = pred.cuda().requires_grad_(True)
    got = get_seg_loss_upsampled(p_gpu, lab.cuda(), ignore_index=255)
    (got * 0.2).backward()
    assert abs(got.item() - want.item()) < 1e-5 * max(1.0, abs(want.item()))
    assert rel_err(p_gpu.grad, p_ref.grad) < 1e-4


def test_seg_loss_without_foreground_or_background_is_finite():
    from dupl_b200.model.losses import get_seg_loss
    pred = torch.randn(1, 4, 8, 8).cuda().requires_grad_(True)
    lab = torch.full((1, 8, 8), 255, dtype=torch.long).cuda()
    loss = get_seg_loss(pred, lab)
    loss.backward()
    assert loss.item() == 0.0 and torch.count_nonzero(pred.grad) == 0


@pytest.mark.parametrize("b,C,h,w", [(2, 768, 7, 9), (1, 96, 12, 11), (3, 768, 28, 28), (2, 768, 4, 4), (4, 768, 28, 28), (1, 64, 8, 6)])
def test_ptc_loss_value_and_gradient(b, C, h, w):
    from dupl_b200.model.losses import get_masked_ptc_loss
    from dupl_b200.utils import cam_helper
    from oracle import dupl_oracle as O
    g = torch.Generator().manual_seed(C + h)
    fmap = torch.randn(b, C, h, w, generator=g)
    lab = torch.randint(0, 4, (b, h, w), generator=g)
    lab[lab == 3] = 255
    aff = O.label_to_aff_mask(lab)
    f_ref = fmap.clone().requires_grad_(True)
    want = O.masked_ptc_loss(f_ref, aff)
    (want * 0.5).backward()
    f_gpu = fmap.cuda().requires_grad_(True)
    got = get_masked_ptc_loss(f_gpu, cam_helper.label_to_aff_mask(lab.cuda()))
    (got * 0.5).backward()
    assert abs(got.item() - want.item()) < 1e-5
    assert rel_err(f_gpu.grad, f_ref.grad) < 1e-3


def test_losses_match_reference_golden_values():
    from dupl_b200.model.losses import get_masked_ptc_loss, get_seg_loss
    d = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "losses.npz")).items()}
    ptc = get_masked_ptc_loss(d["fmap"].cuda(), d["aff"].cuda())
    assert abs(ptc.item() - float(d["ptc"])) < 1e-5
    seg = get_seg_loss(d["pred"].float().cuda(), d["label"].long().cuda(), ignore_index=255)
    assert abs(seg.item() - float(d["seg"])) < 1e-4


def test_ptc_tensor_core_path_agrees_with_the_fp32_simt_path():
    """Same loss through the split-bf16 GEMM (Gram and dX_hat on tcgen05) and through the fp32 CUDA-core kernels."""
    from dupl_b200.model.losses import _PtcLoss, _PtcLossSimt
    from dupl_b200.utils import cam_helper
    g = torch.Generator().manual_seed(3)
    fmap = torch.randn(4, 768, 28, 28, generator=g).cuda()
    lab = torch.randint(0, 5, (4, 28, 28), generator=g)
    lab[lab == 4] = 255
    aff = cam_helper.label_to_aff_mask(lab.cuda())
    outs = []
    for fn in (_PtcLoss, _PtcLossSimt):
        f = fmap.clone().requires_grad_(True)
        loss = fn.apply(f, aff)
        (loss * 0.2).backward()
        outs.append((loss.item(), f.grad.clone()))
    assert abs(outs[0][0] - outs[1][0]) < 2e-6
    assert rel_err(outs[0][1], outs[1][1]) < 1e-4


@pytest.mark.parametrize("b,C,h,w,H,W,frac", [(2, 21, 21, 21, 448, 448, 0.3), (1, 81, 14, 14, 224, 224, 0.05), (2, 21, 7, 9, 64, 80, 0.0)])
def test_consistency_ce_sum_upsampled_matches_torch(b, C, h, w, H, W, frac):
    """ce_criterion(F.interpolate(aug), pseudo_seg).sum() / mask.sum() of train_final_voc.py:407-436 (0 for an empty mask)."""
    import torch.nn.functional as F
    from dupl_b200.model.losses import ce_sum_upsampled
    g = torch.Generator().manual_seed(C + h)
    pred = torch.randn(b, C, h, w, generator=g) * 3
    target = torch.randint(0, C, (b, H, W), generator=g)
    target[torch.rand(b, H, W, generator=g) >= frac] = 255
    p_ref = pred.clone().requires_grad_(True)
    up = F.interpolate(p_ref, size=(H, W), mode="bilinear", align_corners=False)
    n = (target != 255).sum()
    want = F.cross_entropy(up, target, ignore_index=255, reduction="none").sum() / n if n > 0 else up.sum() * 0.0
    (want * 0.05).backward()
    p_gpu = pred.cuda().requires_grad_(True)
    got = ce_sum_upsampled(p_gpu, target.cuda(), 255)
    (got * 0.05).backward()
    assert abs(got.item() - want.item()) < 1e-5 * max(1.0, abs(want.item()))
    if n > 0:
        assert rel_err(p_gpu.grad, p_ref.grad) < 1e-4
    else:
        assert torch.count_nonzero(p_gpu.grad) == 0


def test_fused_classification_and_discrepancy_losses_match_torch():
    """train_final_voc.py:299-305 (4 x F.multilabel_soft_margin_loss) and :440-447 (cosine discrepancy along the spatial axis):
    value and gradients of the fused kernels against the torch expressions of the script."""
    import torch.nn.functional as F
    from dupl_b200.model.losses import discrepancy_loss, multilabel_soft_margin_sum
    g = torch.Generator().manual_seed(3)
    for K, dt in ((20, torch.float32), (80, torch.uint8)):
        y = (torch.rand(4, K, generator=g) < 0.15).to(dt).cuda()
        xs = [(torch.randn(4, K, generator=g) * 3).cuda().requires_grad_() for _ in range(4)]
        ref = sum(F.multilabel_soft_margin_loss(x, y) for x in xs)
        gr = torch.autograd.grad(ref * 1.7, xs)
        xs2 = [x.detach().clone().requires_grad_() for x in xs]
        got = multilabel_soft_margin_sum(xs2, y.float())
        gg = torch.autograd.grad(got * 1.7, xs2)
        assert abs(got.item() - ref.item()) < 1e-6 * max(1.0, abs(ref.item()))
        for a, b_ in zip(gg, gr):
            assert rel_err(a, b_) < 1e-5
    f1 = torch.randn(4, 768, 28, 28, generator=g).cuda().requires_grad_()
    f2 = (0.3 * f1.detach() + torch.randn(4, 768, 28, 28, generator=g).cuda()).requires_grad_()
    cos = torch.nn.CosineSimilarity(dim=-1, eps=1e-6)
    a, b_ = f1.view(4, 768, -1), f2.view(4, 768, -1)
    ref = (1 + cos(a.detach(), b_).mean()) + (1 + cos(b_.detach(), a).mean())
    g1, g2 = torch.autograd.grad(ref * 0.1, (f1, f2))
    x1, x2 = f1.detach().clone().requires_grad_(), f2.detach().clone().requires_grad_()
    got = discrepancy_loss(x1, x2)
    h1, h2 = torch.autograd.grad(got * 0.1, (x1, x2))
    assert abs(got.item() - ref.item()) < 1e-6
    assert rel_err(h1, g1) < 1e-4 and rel_err(h2, g2) < 1e-4
    z = torch.zeros(1, 2, 3, 3, device="cuda")                      # zero vectors: norms clamp at eps, loss = 2, no NaN
    assert discrepancy_loss(z, z).item() == 2.0
