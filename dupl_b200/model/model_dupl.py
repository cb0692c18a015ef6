"""Dual-student model with the reference's interface (model/model_dupl.py:9-213).

`network` = ViT-B/16 encoder + LargeFOV decoder + two bias-free 1x1 classifiers;
`siamese_network` = two independent `network`s with the mode switches val / cam_only /
cam_with_grad / branch / need_sp.  The modules own nn.Parameters under the reference's state-dict
names; every forward computation is issued to libdupl.so (no torch.nn math, no CPU fallback).
"""
import torch
import torch.nn as nn

from .. import _lib as L
from .. import encoder as E
from .. import ops
from . import backbone as encoder
from . import decoder


class network(nn.Module):
    def __init__(self, backbone, num_classes=None, pretrained=None, aux_layer=None, add_mlp=False):
        super().__init__()
        if add_mlp:
            raise NotImplementedError("add_mlp=True is never used by the reference scripts (model_dupl.py:112-117)")
        self.num_classes = num_classes
        self.add_mlp = add_mlp
        self.encoder = getattr(encoder, backbone)(pretrained=pretrained, aux_layer=aux_layer)
        self.in_channels = [self.encoder.embed_dim] * 4
        self.decoder = decoder.LargeFOV(in_planes=self.in_channels[-1], out_planes=self.num_classes)
        self.classifier = nn.Conv2d(self.in_channels[-1], self.num_classes - 1, kernel_size=1, bias=False)
        self.aux_classifier = nn.Conv2d(self.in_channels[-1], self.num_classes - 1, kernel_size=1, bias=False)
        self._planes = None  # split-bf16 weight planes live outside the module state (no buffers: SURVEY §8(b))
        self._kept = None    # encoder.KeptActivations left by an MS-CAM pass run with keep_activations (train_step.py)
        self._keep_next = False

    # -- reference API -----------------------------------------------------------------------
    def get_param_groups(self):
        groups = [[], [], [], []]  # backbone; backbone norms; cls heads; seg head  (model_dupl.py:41-62)
        for name, p in self.encoder.named_parameters():
            groups[1 if "norm" in name else 0].append(p)
        groups[2] += [self.classifier.weight, self.aux_classifier.weight]
        groups[3] += list(self.decoder.parameters())
        return groups

    def to_2D(self, x, h, w):
        n, hw, c = x.shape
        return x.transpose(1, 2).reshape(n, c, h, w)

    # -- kernels -----------------------------------------------------------------------------
    def planes(self):
        if self._planes is None:
            self._planes = E.StudentPlanes(self.encoder)
        return self._planes

    def forward(self, x, cam_only=False, val=False, cam_with_grad=False):
        if cam_only:
            cam_aux, cam = cam_only_forward([self], x)[0]
            return cam_aux, cam
        from .. import dense  # decoder / heads / autograd path
        return dense.network_forward(self, x, val=val, cam_with_grad=cam_with_grad)


def cam_only_forward(nets, x, seg_images=None, seg_sizes=None, flip_twin=False, keep_batch=0):
    """cam_only path of `network.forward` (model_dupl.py:69-84) for one or two students at once.

    Default: one segment holding the batch `x` as given.  multi_scale_cam2_siamese passes several
    segments (one per scale, with flipped twins) so that all scales share each GEMM launch.
    keep_batch > 0: the encoder keeps the activations of the first keep_batch images of segment 0 on each
    student (`net._kept`, encoder.KeptActivations) for a training forward of the same images in the same step.
    Returns, per student, (cam_aux, cam) as lists over segments when seg_sizes is given, else tensors.
    """
    L.require_cuda(x)
    single = seg_sizes is None
    if single:
        seg_images, seg_sizes = [x], [tuple(x.shape[-2:])]
    if any(hs < 16 or ws < 16 for hs, ws in seg_sizes):
        raise ValueError("inputs must be at least one 16x16 patch large")
    with torch.no_grad():
        planes = [n.planes() for n in nets]
        aux_idx = nets[0].encoder.aux_block_index()
        aux_out = [None] * len(nets)

        def on_aux(g, tok_g, segs):
            w = nets[g].aux_classifier.weight.detach().reshape(nets[g].num_classes - 1, -1)
            aux_out[g] = ops.cam_contract(tok_g, None, None, L.f32c(w), segs)

        if keep_batch > 0:
            segs, tok, kept = E.run_encoder(planes, seg_images, seg_sizes, flip_twin, aux_idx, on_aux, keep_batch=keep_batch)
            for n, k in zip(nets, kept):
                n._kept = k
        else:
            segs, tok = E.run_encoder(planes, seg_images, seg_sizes, flip_twin, aux_idx, on_aux)
        res = []
        for g, n in enumerate(nets):
            gam, bet = n.encoder.norm.weight.detach(), n.encoder.norm.bias.detach()
            w = L.f32c(n.classifier.weight.detach().reshape(n.num_classes - 1, -1))
            cams = ops.cam_contract(tok[g], gam, bet, w, segs, eps=E.LN_EPS)
            if aux_out[g] is None:  # aux_layer points at the last entry of `embeds` == final-normed tokens
                wa = L.f32c(n.aux_classifier.weight.detach().reshape(n.num_classes - 1, -1))
                aux_out[g] = ops.cam_contract(tok[g], gam, bet, wa, segs, eps=E.LN_EPS)
            res.append((aux_out[g][0], cams[0]) if single else (aux_out[g], cams))
    return res


class siamese_network(nn.Module):
    def __init__(self, backbone, num_classes=None, pretrained=None, aux_layer=None):
        super().__init__()
        self.branch1 = network(backbone, num_classes=num_classes, pretrained=pretrained, aux_layer=aux_layer)
        self.branch2 = network(backbone, num_classes=num_classes, pretrained=pretrained, aux_layer=aux_layer)

    def get_param_groups(self):
        g1, g2 = self.branch1.get_param_groups(), self.branch2.get_param_groups()
        # order of model_dupl.py:119-154: encoders of both students, then heads, then decoders
        return [g1[0] + g2[0], g1[1] + g2[1], g1[2] + g2[2], g1[3] + g2[3]]

    def _pick(self, branch):
        return self.branch1 if branch == 1 else self.branch2

    def forward(self, x, val=False, cam_only=False, cam_with_grad=False, branch=None, need_sp=False):
        if cam_only:
            if branch is None:
                (a1, c1), (a2, c2) = cam_only_forward([self.branch1, self.branch2], x)
                return a1, c1, a2, c2
            return self._pick(branch)(x, cam_only=True)
        from .. import dense
        if val or cam_with_grad:
            if branch is None:
                r1, r2 = dense.pair_forward(self.branch1, self.branch2, x, val=val, cam_with_grad=cam_with_grad)
                return {"branch1": r1, "branch2": r2}
            return self._pick(branch)(x, val=val, cam_with_grad=cam_with_grad)
        if branch is None:
            if need_sp:
                # model_dupl.py:190-205 first runs both students on the 2b cat batch and discards the
                # result; only the two passes below reach the caller, so only they are computed.
                x, x_aug = x.chunk(2)
                r1, r2 = dense.pair_forward(self.branch1, self.branch2, x)
                s1, s2 = dense.pair_forward_aug(self.branch1, self.branch2, x_aug, scale=0.75)
                return {"branch1": r1, "branch2": r2, "branch1_aug": s1, "branch2_aug": s2}
            r1, r2 = dense.pair_forward(self.branch1, self.branch2, x)
            return {"branch1": r1, "branch2": r2}
        return self._pick(branch)(x)
