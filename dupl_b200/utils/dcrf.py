"""DenseCRF post-processing with the reference's interface (utils/dcrf.py:7-69), on the GPU.

The reference delegates to pydensecrf (CPU, third-party).  Here the same mean-field inference runs in
libdupl.so (crf.cu); inputs and outputs stay numpy arrays like the reference's, torch CUDA tensors are
accepted too (and then returned as CUDA tensors, skipping the host round trip).
"""
import numpy as np
import torch

from .. import ops

_WS = ops.CrfWorkspace()
_PINNED = {}     # (tag, shape, dtype) -> page-locked staging tensor, reused across calls of the same geometry


def _pinned(tag, shape, dtype):
    key = (tag, tuple(shape), dtype)
    buf = _PINNED.get(key)
    if buf is None:
        if len(_PINNED) >= 8:                       # a handful of image sizes at most stay page-locked
            _PINNED.pop(next(iter(_PINNED)))
        buf = torch.empty(tuple(shape), dtype=dtype).pin_memory()
        _PINNED[key] = buf
    return buf


def _host_to_device(tag, arr, dev):
    """numpy (pageable) -> device through a cached page-locked staging buffer: torch's multi-threaded host copy fills the
    staging buffer and the DMA runs at PCIe speed, instead of the driver's single-threaded pageable path (the probability
    map of a 640x480x81 image is 100 MB each way: the interface of the reference costs more than the inference)."""
    t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.numel() < (1 << 18):
        return t.to(dev)
    stage = _pinned(tag, t.shape, t.dtype)
    stage.copy_(t)
    return stage.to(dev, non_blocking=True)


def _device_to_numpy(q):
    if q.numel() < (1 << 18):
        return q.cpu().numpy()
    stage = _pinned("out", q.shape, q.dtype)
    stage.copy_(q, non_blocking=True)
    torch.cuda.current_stream(q.device).synchronize()
    out = torch.empty(q.shape, dtype=q.dtype)       # a fresh array per call, like the reference returns
    out.copy_(stage)
    return out.numpy()


def _to_cuda(image, arr):
    was_numpy = not torch.is_tensor(arr)
    dev = arr.device if torch.is_tensor(arr) and arr.is_cuda else torch.device("cuda", torch.cuda.current_device())
    img = (image if torch.is_tensor(image) else torch.from_numpy(np.ascontiguousarray(image))).to(dev)
    x = _host_to_device("probs", arr, dev) if was_numpy else arr.to(dev)
    return img, x, was_numpy


def _infer(image, x, was_numpy, **kw):
    q, _ = ops.crf_inference(image, x, ws=_WS, **kw)
    return _device_to_numpy(q) if was_numpy else q


class DenseCRF(object):
    def __init__(self, iter_max, pos_w, pos_xy_std, bi_w, bi_xy_std, bi_rgb_std):
        self.iter_max = iter_max
        self.pos_w = pos_w
        self.pos_xy_std = pos_xy_std
        self.bi_w = bi_w
        self.bi_xy_std = bi_xy_std
        self.bi_rgb_std = bi_rgb_std

    def __call__(self, image, probmap):
        """image uint8 [H,W,3], probmap float32 [C,H,W] -> Q float32 [C,H,W] (dcrf.py:51-69)."""
        img, x, was_numpy = _to_cuda(image, probmap)
        return _infer(img, x, was_numpy, iters=self.iter_max, pos_w=self.pos_w, pos_xy_std=self.pos_xy_std, bi_w=self.bi_w,
                      bi_xy_std=self.bi_xy_std, bi_rgb_std=self.bi_rgb_std)


def crf_inference(img, probs, t=10, scale_factor=1, labels=21):
    """dcrf.py:7-24: Gaussian sxy=3/scale compat 3; bilateral sxy=80/scale, srgb=13, compat 10."""
    image, x, was_numpy = _to_cuda(img, probs)
    return _infer(image, x, was_numpy, iters=t, pos_w=3.0, pos_xy_std=3.0 / scale_factor, bi_w=10.0,
                  bi_xy_std=80.0 / scale_factor, bi_rgb_std=13.0)


def crf_inference_label(img, labels, t=10, n_labels=21, gt_prob=0.7):
    """dcrf.py:26-40: unary_from_labels(gt_prob, zero_unsure=False), Gaussian 3/3, bilateral 50/5/10 -> arg-max labels."""
    lab = np.asarray(labels.cpu() if torch.is_tensor(labels) else labels).reshape(-1).astype(np.int64)
    h, w = np.asarray(img.cpu() if torch.is_tensor(img) else img).shape[:2]
    n_energy = -np.log((1.0 - gt_prob) / (n_labels - 1))
    p_energy = -np.log(gt_prob)
    U = np.full((n_labels, lab.size), n_energy, dtype=np.float32)
    U[lab, np.arange(lab.size)] = p_energy
    image, x, _ = _to_cuda(img, U.reshape(n_labels, h, w))
    q, _ = ops.crf_inference(image, x, iters=t, pos_w=3.0, pos_xy_std=3.0, bi_w=10.0, bi_xy_std=50.0, bi_rgb_std=5.0,
                             input_is_energy=True, ws=_WS)
    return q.argmax(0).cpu().numpy()
