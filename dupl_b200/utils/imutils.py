"""The two functions of the reference's utils/imutils.py that sit on the training loop (train_final_voc.py:186-191), with the
reference's names and signatures, on the GPU:

    denormalize_img2(imgs)                    utils/imutils.py:17-31  (x*std + mean -> uint8 truncation -> /255)
    augment_data_strong(images, n=4, m=20)    utils/imutils.py:305-317 (ToPILImage -> RandAugment(n, m) -> ToTensor -> Normalize -> flip)

The reference runs the second one through Pillow on the host in every iteration; here the batch stays on the device
(libdupl.so, csrc/augment.cu: every operation bit-exact with Pillow) and only the choice of operations — `random.choices` over
`augment_list()`, utils/randomaug.py:260 — happens on the host, consuming Python's `random` stream exactly like the reference.
"""
import ctypes as C
import random

import torch

from .. import _lib as L
from ..pipeline import denormalize_img2  # noqa: F401  (same arithmetic as the reference's script-side helper)

# utils/randomaug.py:185-200 — augment_list(): (operation, min, max), in this order
AUGMENT_LIST = (("AutoContrast", 0, 1), ("Equalize", 0, 1), ("Posterize", 0, 6), ("Color", 0.1, 1.9), ("Contrast", 0.1, 1.9),
                ("Brightness", 0.1, 1.9), ("Sharpness", 0.1, 1.9))

_WS = {}


def _magnitudes(m):
    return [(float(m) / 30) * float(hi - lo) + lo for _, lo, hi in AUGMENT_LIST]       # utils/randomaug.py:262


def draw_ops(batch, n):
    """The reference builds RandAugment(n, m) per image and calls `random.choices(self.augment_list, k=n)` once per image, in
    batch order (imutils.py:310-313, randomaug.py:260): the same calls here -> [n][batch] operation indices."""
    idx = list(range(len(AUGMENT_LIST)))
    per_image = [random.choices(idx, k=n) for _ in range(batch)]
    return [[per_image[b][s] for b in range(batch)] for s in range(n)]


def augment_data_strong(images, n=4, m=20, ops=None, out=None):
    """images: float [B,3,H,W] in [0,1] on the GPU (`inputs_denorm`) -> normalised, flipped, augmented batch (same shape).
    ops: optional [n][B] operation indices (default: drawn like the reference does); a device int32 tensor is used as is
    (CUDA-graph replay: refill it before each step)."""
    L.require_cuda(images)
    x = L.f32c(images)
    B, ch, H, W = x.shape
    if ch != 3:
        raise ValueError("augment_data_strong expects RGB images [B,3,H,W]")
    dev = x.device
    if torch.is_tensor(ops) and ops.is_cuda:
        ops_dev = ops.to(torch.int32).contiguous()
    else:
        host = torch.as_tensor(ops if ops is not None else draw_ops(B, n), dtype=torch.int32).reshape(-1, B)
        ops_dev = host.pin_memory().to(dev, non_blocking=True) if host.numel() else torch.zeros(1, B, dtype=torch.int32, device=dev)[:0]
    if ops_dev.dim() != 2 or ops_dev.shape[1] != B:
        raise ValueError("ops must be [n][batch]")
    n_ops = ops_dev.shape[0]
    key = (B, H, W, dev)
    ws = _WS.get(key)
    if ws is None:
        nbytes = C.c_size_t(0)
        L.check(L.lib().dupl_randaug_workspace_bytes(B, H, W, C.byref(nbytes)), "dupl_randaug_workspace_bytes")
        ws = _WS[key] = torch.empty(nbytes.value, dtype=torch.uint8, device=dev)
    out = torch.empty_like(x) if out is None else out
    mags = (C.c_float * 7)(*_magnitudes(m))
    ops_ptr = L.ptr(ops_dev) if n_ops else L.ptr(ws)        # never dereferenced when n_ops == 0
    L.check(L.lib().dupl_randaug(L.ptr(x), L.ptr(out), B, H, W, ops_ptr, n_ops, mags, L.ptr(ws), ws.numel(),
                                 L.stream_ptr(dev)), "dupl_randaug")
    return out
