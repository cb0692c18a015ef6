"""Decoder + classification heads on top of the encoder (reference: model/model_dupl.py:86-106,
model/decoder/conv_head.py:33-41): the non-`cam_only` outputs `(cls_x4, seg, _x4, cls_aux)`.

No-grad calls run the inference kernels below; with gradients enabled the call is routed to
dupl_b200.train (one autograd.Function per student with CUDA forward and backward).  There is no
PyTorch fallback on either path.
"""
import os

import torch

from . import _lib as L
from . import encoder as E
from . import ops

DECODER_DIL = 5


class DecoderPlanes:
    """split-bf16 planes of the LargeFOV weights in implicit-GEMM layout [Cout][tap][Cin] (conv8 padded to a
    multiple of 16 output rows), refreshed when a parameter changes."""

    def __init__(self, net):
        self.net = net
        self._cache = {}

    def get(self, name):
        conv = getattr(self.net.decoder, name)
        p = conv.weight
        key = (p.data_ptr(), p._version)
        hit = self._cache.get(name)
        if hit is None or hit[0] != key:
            w = p.detach()
            cout = w.shape[0]
            w2 = w.permute(0, 2, 3, 1).reshape(cout, -1)  # storage re-layout only: [Cout, ky, kx, Cin]
            pad = (-cout) % 16
            if pad:
                w2 = torch.cat([w2, w2.new_zeros(pad, w2.shape[1])], 0)
            hit = (key, ops.split_bf16(w2.contiguous()))
            self._cache[name] = hit
        return hit[1]


def _invalidate(self):
    for name, (key, planes) in list(self._cache.items()):
        self._cache[name] = ((0, -1), planes)


DecoderPlanes.invalidate = _invalidate


def _get_t(self, name):
    """Transposed implicit-GEMM weight planes [9*Cin, Cout] (B operand of the conv dgrad GEMM)."""
    conv = getattr(self.net.decoder, name)
    p = conv.weight
    key = (p.data_ptr(), p._version)
    hit = self._cache.get(name + "^T")
    if hit is None or hit[0] != key:
        w = self.get(name)
        hit = (key, E.transpose_planes(w, w[0].shape[0], w[0].shape[1]))
        self._cache[name + "^T"] = hit
    return hit[1]


DecoderPlanes.get_t = _get_t


def _decoder_planes(net):
    if getattr(net, "_dec_planes", None) is None:
        net._dec_planes = DecoderPlanes(net)
    return net._dec_planes


def _wants_grad(nets):
    return torch.is_grad_enabled() and any(p.requires_grad for n in nets for p in n.parameters())


@torch.no_grad()
def _heads(nets, x, aux_seg_only=False):
    """Runs the encoder for all `nets` (students) on batch x and then decoder + heads per student."""
    L.require_cuda(x)
    x = L.f32c(x)
    B, _, H, W = x.shape
    if H < 16 or W < 16:
        raise ValueError("inputs must be at least one 16x16 patch large")
    planes = [n.planes() for n in nets]
    aux_idx = nets[0].encoder.aux_block_index()
    aux_tok = [None] * len(nets)

    def on_aux(g, tok_g, segs):
        aux_tok[g] = tok_g.clone()  # block-9 residual stream (un-normed), vit.py:319-326

    segs, tok = E.run_encoder(planes, [x], [(H, W)], False, aux_idx, on_aux)
    sg = segs[0]
    gh, gw, np_, M = sg.gh, sg.gw, sg.gh * sg.gw, tok[0].shape[0]
    dev = x.device
    out = []
    for g, n in enumerate(nets):
        K = n.num_classes - 1
        enc = n.encoder
        xn_hi = torch.empty(M, E.EMBED, dtype=torch.bfloat16, device=dev)
        xn_lo = torch.empty_like(xn_hi)
        xn = torch.empty(M, E.EMBED, dtype=torch.float32, device=dev)
        ops.layernorm_split(tok[g], enc.norm.weight.detach(), enc.norm.bias.detach(), xn_hi, xn_lo, eps=E.LN_EPS, out_f32=xn)
        # --- LargeFOV: conv6 (3x3 d5) + ReLU, conv7 (3x3 d5) + ReLU, conv8 (1x1)   conv_head.py:33-41
        dp = _decoder_planes(n)
        c6 = ops.im2col3x3(xn_hi, xn_lo, B, gh, gw, DECODER_DIL, 0, sg.tokens, 1)
        h6 = (torch.empty(B * np_, 512, dtype=torch.bfloat16, device=dev), torch.empty(B * np_, 512, dtype=torch.bfloat16, device=dev))
        ops.gemm_bf16x3([dict(a=c6, w=dp.get("conv6"), out=h6)], B * np_, 512, 9 * E.EMBED, L.EPI_RELU_SPLIT)
        c7 = ops.im2col3x3(h6[0], h6[1], B, gh, gw, DECODER_DIL, 0, np_, 0)
        h7 = (torch.empty_like(h6[0]), torch.empty_like(h6[0]))
        ops.gemm_bf16x3([dict(a=c7, w=dp.get("conv7"), out=h7)], B * np_, 512, 9 * 512, L.EPI_RELU_SPLIT)
        w8 = dp.get("conv8")
        n8 = w8[0].shape[0]
        seg_rows = torch.empty(B * np_, n8, dtype=torch.float32, device=dev)
        ops.gemm_bf16x3([dict(a=h7, w=w8, out_f32=seg_rows)], B * np_, n8, 512, L.EPI_F32)
        seg = ops.rows_to_nchw(seg_rows, B, gh, gw, n.num_classes, 0, np_, 0)
        if aux_seg_only:
            out.append(seg)
            continue
        # --- _x4 = to_2D(final-normed patch tokens)   model_dupl.py:73
        x4 = ops.rows_to_nchw(xn, B, gh, gw, E.EMBED, 0, sg.tokens, 1)
        # --- GMP + 1x1 classifiers   model_dupl.py:88-98
        wc = L.f32c(n.classifier.weight.detach().reshape(K, -1))
        wa = L.f32c(n.aux_classifier.weight.detach().reshape(K, -1))
        cls_x4 = ops.gmp_classify(xn, wc, B, np_, 0, sg.tokens, 1)
        aux_src = xn if aux_tok[g] is None else aux_tok[g]  # aux_layer == last entry of `embeds` -> final-normed tokens
        cls_aux = ops.gmp_classify(aux_src, wa, B, np_, 0, sg.tokens, 1)
        out.append((cls_x4, seg, x4, cls_aux))
    return out


def _with_cam_grad(net, outs):
    """model_dupl.py:100-104 (cam_with_grad=True; no reference script uses it): a CAM that carries gradient w.r.t. the feature
    map, built from `_x4` exactly as the reference does — on top of the kernels' outputs, so autograd reaches the encoder
    through the feature-map gradient of the student's backward."""
    import torch.nn.functional as F
    cls_x4, seg, x4, cls_aux = outs
    cam_grad = F.conv2d(x4, net.classifier.weight.detach())
    cam_grad = cam_grad + F.adaptive_max_pool2d(-cam_grad, (1, 1))
    cam_grad = cam_grad / F.adaptive_max_pool2d(cam_grad, (1, 1)) + 1e-5
    return cls_x4, seg, x4, cls_aux, cam_grad


def network_forward(net, x, val=False, cam_with_grad=False):
    if _wants_grad([net]):
        from . import train
        outs = train.student_forward(net, x)
    else:
        outs = _heads([net], x)[0]
    return _with_cam_grad(net, outs) if cam_with_grad else outs


def pair_forward(net1, net2, x, val=False, cam_with_grad=False):
    if cam_with_grad:
        return network_forward(net1, x, cam_with_grad=True), network_forward(net2, x, cam_with_grad=True)
    if _wants_grad([net1, net2]):
        from . import train
        if train._arena_mode(net1) and train._arena_mode(net2) and os.environ.get("DUPL_PAIR_BACKWARD", "1") != "0":
            return train.pair_student_forward(net1, net2, x)      # one autograd node: grouped dgrad / wgrad launches
        return train.student_forward(net1, x), train.student_forward(net2, x)
    r1, r2 = _heads([net1, net2], x)
    return r1, r2


def pair_forward_aug(net1, net2, x_aug, scale=0.75):
    """seg logits of the strongly-augmented view resized by `scale` (model_dupl.py:195-205: F.interpolate with
    scale_factor=0.75, bilinear, then both students; only the seg output reaches the caller)."""
    from . import train
    H, W = x_aug.shape[-2:]
    size = (int(H * scale), int(W * scale))
    # F.interpolate(scale_factor=s) maps coordinates with 1/s, a size-based resize with H/out: identical only when H*s is an
    # integer; the patch grid further needs multiples of 16 (448 -> 336 for the scripts' crop size)
    if size[0] != H * scale or size[1] != W * scale or size[0] % 16 or size[1] % 16:
        raise ValueError(f"need_sp: {H}x{W} * {scale} must be integer multiples of 16 (got {H * scale}x{W * scale})")
    if _wants_grad([net1, net2]):
        return train.student_forward(net1, x_aug, size)[1], train.student_forward(net2, x_aug, size)[1]
    with torch.no_grad():
        return train._forward(net1, x_aug, size)[0][1], train._forward(net2, x_aug, size)[0][1]
