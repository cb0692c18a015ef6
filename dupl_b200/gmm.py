"""GPU replacement of the sklearn GaussianMixture noise filter that train_final_voc.py:358-394 runs inline
(one GPU->CPU->GPU round trip per image and student in the reference)."""
import torch

from . import _lib as L


def gmm_noise_filter(seg_loss, refined_label, ignore_index=255, loss_min=0.1, min_count=1000, gmm_valid_thre=1.0, gamma=0.95,
                     reg_covar=5e-4, max_iter=10, tol=1e-2):
    """seg_loss: per-pixel CE [b,H,W] (detached); refined_label: float32 [b,H,W] in {0..K, ignore}, modified IN PLACE
    like the script does (`refined_pseudo_label[i][noise_mask] = 255`).  Returns int32 [b,4] diagnostics on the device
    (samples used, filter applied?, EM iterations, pixels flipped) — nothing is read back on the host."""
    L.require_cuda(seg_loss, refined_label)
    if refined_label.dtype != torch.float32 or not refined_label.is_contiguous():
        raise ValueError("refined_label must be a contiguous float32 tensor (it is updated in place)")
    loss = L.f32c(seg_loss.detach())
    b = loss.shape[0]
    n = loss[0].numel()
    info = torch.empty(b, 4, dtype=torch.int32, device=loss.device)
    L.check(L.lib().dupl_gmm_filter(L.ptr(loss), L.ptr(refined_label), b, n, float(ignore_index), loss_min, min_count, gmm_valid_thre,
                                    gamma, reg_covar, max_iter, tol, L.ptr(info), L.stream_ptr(loss.device)), "dupl_gmm_filter")
    return info
