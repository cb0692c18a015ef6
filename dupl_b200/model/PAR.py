"""Pixel-adaptive refinement with the reference's interface (model/PAR.py:26-91)."""
import numpy as np
import torch
import torch.nn as nn

from .. import _lib as L
from .. import ops


def get_kernel():
    """The 8 one-hot 3x3 taps of the reference (PAR.py:10-24); kept as a buffer for state-dict
    compatibility — the CUDA kernels gather neighbours directly."""
    weight = torch.zeros(8, 1, 3, 3)
    for i, (r, c) in enumerate(((0, 0), (0, 1), (0, 2), (1, 0), (1, 2), (2, 0), (2, 1), (2, 2))):
        weight[i, 0, r, c] = 1
    return weight


class PAR(nn.Module):
    def __init__(self, dilations, num_iter):
        super().__init__()
        self.dilations = list(dilations)
        self.num_iter = num_iter
        self.register_buffer("kernel", get_kernel())
        self.pos = self.get_pos()
        self.dim = 2
        self.w1 = 0.3
        self.w2 = 0.01

    def get_pos(self):
        ker = torch.ones(1, 1, 8, 1, 1)
        for i in (0, 2, 5, 7):
            ker[0, 0, i, 0, 0] = np.sqrt(2)
        return torch.cat([ker * d for d in self.dilations], dim=2)

    def affinity(self, imgs):
        L.require_cuda(imgs)
        return ops.par_affinity(imgs, self.dilations, self.w1, self.w2)

    def forward(self, imgs, masks):
        L.require_cuda(imgs, masks)
        if masks.shape[-2:] != imgs.shape[-2:]:
            # PAR.py:66: masks are first brought to the image size (bilinear, align_corners=True).  Every call on the hot path
            # passes equal sizes (identity); the general case is the reference's own one-line torch resize in front of the kernels.
            masks = torch.nn.functional.interpolate(masks, size=imgs.shape[-2:], mode="bilinear", align_corners=True)
        if masks.shape[0] != imgs.shape[0]:
            raise ValueError("PAR: imgs and masks must have the same batch size")
        aff = self.affinity(imgs)
        m = L.f32c(masks).clone()
        return ops.par_propagate(aff, m, self.dilations, self.num_iter)
