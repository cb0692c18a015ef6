"""DenseCRF post-processing with the reference's interface (utils/dcrf.py:7-69), on the GPU.

The reference delegates to pydensecrf (CPU, third-party).  Here the same mean-field inference runs in
libdupl.so (crf.cu); inputs and outputs stay numpy arrays like the reference's, torch CUDA tensors are
accepted too (and then returned as CUDA tensors, skipping the host round trip).
"""
import numpy as np
import torch

from .. import ops

_WS = ops.CrfWorkspace()


def _to_cuda(image, arr):
    was_numpy = not torch.is_tensor(arr)
    dev = arr.device if torch.is_tensor(arr) and arr.is_cuda else torch.device("cuda", torch.cuda.current_device())
    img = torch.as_tensor(np.ascontiguousarray(image) if not torch.is_tensor(image) else image).to(dev)
    x = torch.as_tensor(np.ascontiguousarray(arr) if was_numpy else arr).to(dev)
    return img, x, was_numpy


def _infer(image, x, was_numpy, **kw):
    q, _ = ops.crf_inference(image, x, ws=_WS, **kw)
    return q.cpu().numpy() if was_numpy else q


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
