"""LargeFOV parameter container (reference: model/decoder/conv_head.py:11-41):
conv6 3x3 d5 (in->512), conv7 3x3 d5 (512->512), conv8 1x1 (512->classes), no bias."""
import torch.nn as nn


class LargeFOV(nn.Module):
    def __init__(self, in_planes, out_planes, dilation=5):
        super().__init__()
        self.embed_dim = 512
        self.dilation = dilation
        self.conv6 = nn.Conv2d(in_planes, 512, 3, padding=dilation, dilation=dilation, bias=False)
        self.conv7 = nn.Conv2d(512, 512, 3, padding=dilation, dilation=dilation, bias=False)
        self.conv8 = nn.Conv2d(512, out_planes, 1, bias=False)

    def forward(self, *a, **k):
        raise RuntimeError("the decoder is driven by dupl_b200 CUDA kernels, not by nn.Module.forward")
