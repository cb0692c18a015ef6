"""Host logic: device-side confusion matrix / scores (dupl_b200/utils/evaluate.py) against a numpy restatement of
utils/evaluate.py:9-60, against the reference module itself when it is mounted, and sharded over 2 ranks (gloo)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _np_scores(gts, preds, C, pseudo=False):
    hist = np.zeros((C, C))
    for lt, lp in zip(gts, preds):
        lt, lp = lt.flatten().copy(), lp.flatten().copy()
        if pseudo:
            lt[lp == 255] = 255
            lp[lp == 255] = 0
        m = (lt >= 0) & (lt < C)
        hist += np.bincount(C * lt[m].astype(int) + lp[m], minlength=C ** 2).reshape(C, C)
    with np.errstate(divide="ignore", invalid="ignore"):
        iu = np.diag(hist) / (hist.sum(1) + hist.sum(0) - np.diag(hist))
        return dict(pAcc=np.diag(hist).sum() / hist.sum(), mAcc=np.nanmean(np.diag(hist) / hist.sum(1)),
                    miou=np.nanmean(iu[hist.sum(1) > 0]), iou=iu), hist


def _data(seed, n=5, C=21, with_ignore_pred=False):
    rng = np.random.RandomState(seed)
    gts, preds = [], []
    for _ in range(n):
        h, w = rng.randint(20, 40), rng.randint(20, 40)
        gt = rng.randint(0, C - 6, (h, w)).astype(np.int16)     # some classes never occur (valid-class handling)
        gt[rng.rand(h, w) < 0.1] = 255
        pr = rng.randint(0, C, (h, w)).astype(np.int16)
        if with_ignore_pred:
            pr[rng.rand(h, w) < 0.2] = 255
        gts.append(gt)
        preds.append(pr)
    return gts, preds


@pytest.mark.parametrize("pseudo", [False, True])
def test_scores_match_numpy_restatement(pseudo):
    from dupl_b200.utils import evaluate
    gts, preds = _data(0, with_ignore_pred=pseudo)
    want, hist = _np_scores(gts, preds, 21, pseudo)
    got = (evaluate.pseudo_scores if pseudo else evaluate.scores)(gts, preds, num_classes=21)
    cm = evaluate.ConfusionMatrix(21).update([torch.from_numpy(g) for g in gts], [torch.from_numpy(p) for p in preds], pseudo=pseudo)
    assert np.array_equal(cm.hist.numpy(), hist.astype(np.int64))
    for k in ("pAcc", "mAcc", "miou"):
        assert got[k] == want[k]
    assert all((np.isnan(got["iou"][c]) and np.isnan(want["iou"][c])) or got["iou"][c] == want["iou"][c] for c in range(21))


def test_scores_match_reference_module_when_mounted():
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference tree not mounted")
    sys.path.insert(0, "/root/reference")
    try:
        import importlib.util
        spec = importlib.util.spec_from_file_location("ref_evaluate", "/root/reference/utils/evaluate.py")
        ref_eval = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_eval)
    finally:
        sys.path.remove("/root/reference")
    from dupl_b200.utils import evaluate
    gts, preds = _data(3)
    want = ref_eval.scores(gts, preds, num_classes=21)
    got = evaluate.scores(gts, preds, num_classes=21)
    for k in ("pAcc", "mAcc", "miou"):
        assert got[k] == want[k]
    gts, preds = _data(4, with_ignore_pred=True)
    want = ref_eval.pseudo_scores([g.copy() for g in gts], [p.copy() for p in preds], num_classes=21)
    got = evaluate.pseudo_scores(gts, preds, num_classes=21)
    for k in ("pAcc", "mAcc", "miou"):
        assert got[k] == want[k]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dupl_b200.eval_sweep import shard_indices
    from dupl_b200.utils.evaluate import ConfusionMatrix
    gts, preds = _data(7, n=9)
    cm = ConfusionMatrix(21)
    for i in shard_indices(len(gts), rank, world):
        cm.update(gts[i], preds[i])
    cm.all_reduce()
    out[rank] = cm.hist.tolist()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_confusion_matrix_all_reduce_world_size_2():
    port = 29700 + os.getpid() % 200
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        gts, preds = _data(7, n=9)
        _, hist = _np_scores(gts, preds, 21)
        assert np.array_equal(np.array(out[0]), hist.astype(np.int64))
        assert out[0] == out[1]
