"""Segmentation scores with the reference's interface (utils/evaluate.py:9-67), accumulated on the device.

The reference gathers every prediction and ground truth of the validation set as numpy arrays on rank 0 and builds
the confusion matrix there (utils/train_helper.py:141-175 -> evaluate.scores).  Here the C x C histogram is an int64
tensor that lives where the predictions are produced (one `bincount` per batch, exact integer arithmetic); with
several ranks each one accumulates its shard of the images (eval_sweep.shard_indices) and ONE C x C all-reduce merges
them — the only collective of the evaluation path.
"""
import numpy as np
import torch


def _as_tensor(a, device=None):
    t = a if torch.is_tensor(a) else torch.as_tensor(np.asarray(a))
    return t.to(device) if device is not None else t


def fast_hist(label_true, label_pred, num_classes):
    """utils/evaluate.py:9-15 (_fast_hist): int64 [C, C], rows = ground truth, columns = prediction."""
    lt = _as_tensor(label_true).reshape(-1).long()
    lp = _as_tensor(label_pred, lt.device).reshape(-1).long()
    mask = (lt >= 0) & (lt < num_classes)
    idx = num_classes * lt[mask] + lp[mask]
    return torch.bincount(idx, minlength=num_classes ** 2).reshape(num_classes, num_classes)


def scores_from_hist(hist):
    """utils/evaluate.py:21-35 on an accumulated histogram -> {"pAcc", "mAcc", "miou", "iou"} (numpy floats like the reference)."""
    hist = hist.detach().cpu().numpy().astype(np.float64)
    num_classes = hist.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.diag(hist).sum() / hist.sum()
        acc_cls = np.nanmean(np.diag(hist) / hist.sum(axis=1))
        iu = np.diag(hist) / (hist.sum(axis=1) + hist.sum(axis=0) - np.diag(hist))
        valid = hist.sum(axis=1) > 0
        mean_iu = np.nanmean(iu[valid])
    return {"pAcc": acc, "mAcc": acc_cls, "miou": mean_iu, "iou": dict(zip(range(num_classes), iu))}


class ConfusionMatrix:
    """Running C x C histogram on `device`; update() per batch, all_reduce() once at the end of a sharded sweep."""

    def __init__(self, num_classes=21, device=None):
        self.num_classes = num_classes
        self.hist = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=device)

    def update(self, label_trues, label_preds, pseudo=False):
        """label_trues / label_preds: tensors or lists of per-image label maps.  pseudo=True applies pseudo_scores'
        convention (utils/evaluate.py:44-49): pixels predicted 255 are dropped from the ground truth."""
        if torch.is_tensor(label_trues) or isinstance(label_trues, np.ndarray):
            label_trues, label_preds = [label_trues], [label_preds]
        for lt, lp in zip(label_trues, label_preds):
            lt = _as_tensor(lt, self.hist.device).reshape(-1).long()
            lp = _as_tensor(lp, self.hist.device).reshape(-1).long()
            if pseudo:
                lt = torch.where(lp == 255, torch.full_like(lt, 255), lt)
                lp = torch.where(lp == 255, torch.zeros_like(lp), lp)
            self.hist += fast_hist(lt, lp, self.num_classes)
        return self

    def all_reduce(self):
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.hist, op=dist.ReduceOp.SUM)
        return self

    def scores(self):
        return scores_from_hist(self.hist)


def scores(label_trues, label_preds, num_classes=21):
    """Drop-in for utils/evaluate.py:17-35."""
    return ConfusionMatrix(num_classes).update(list(label_trues), list(label_preds)).scores()


def pseudo_scores(label_trues, label_preds, num_classes=21):
    """Drop-in for utils/evaluate.py:37-60."""
    return ConfusionMatrix(num_classes).update(list(label_trues), list(label_preds), pseudo=True).scores()
