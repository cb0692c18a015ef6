"""Host-side driver of the ViT-B/16 encoder kernels (reference: model/backbone/vit.py:289-326).

All images of all scales are laid out as segments of ONE token matrix per student, so each linear
layer of each block is a single grouped tcgen05 GEMM launch covering every scale and both students
(dupl.h: dupl_segment).  This module only sequences launches and owns buffers; arithmetic is in
libdupl.so.
"""
import torch

from . import _lib as L
from . import ops

EMBED = 768
HEADS = 12
DEPTH = 12
LN_EPS = 1e-6  # deit.py:100


def transpose_planes(planes, R, Cc, tokens=0, np_=0, first=0):
    """(hi, lo) [R (mapped rows), Cc] bf16 -> (hi^T, lo^T) [Cc, pad64(R)] (zero padded): pure data movement, one launch."""
    Rpad = (R + 63) // 64 * 64
    hi, lo = planes
    ohi = torch.empty(Cc, Rpad, dtype=torch.bfloat16, device=hi.device)
    olo = torch.empty(Cc, Rpad, dtype=torch.bfloat16, device=hi.device)
    L.check(L.lib().dupl_transpose_planes(L.ptr(hi), L.ptr(lo), R, Cc, hi.shape[1], tokens, np_, first, L.ptr(ohi), L.ptr(olo),
                                          Rpad, L.stream_ptr(hi.device)), "dupl_transpose_planes")
    return ohi, olo


class StudentPlanes:
    """split-bf16 planes of one student's encoder weights, refreshed lazily when a parameter changes
    (torch bumps `_version` on every in-place update, e.g. by the optimizer)."""

    GEMM_WEIGHTS = ("patch_embed.proj.weight",) + tuple(
        f"blocks.{i}.{n}.weight" for i in range(DEPTH) for n in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2"))

    def __init__(self, encoder):
        self.encoder = encoder
        self._planes = {}
        self._pos = {}
        self._pdict = None

    def _params(self):
        # nn.Module keeps the same Parameter objects across .to()/.cuda()/optimizer steps (only .data moves),
        # so the name -> Parameter map is built once; walking named_parameters() per lookup cost ~10 ms per pass.
        if self._pdict is None:
            self._pdict = dict(self.encoder.named_parameters())
        return self._pdict

    def plane(self, name):
        p = self._params()[name]
        L.require_cuda(p)
        key = (p.data_ptr(), p._version)
        hit = self._planes.get(name)
        if hit is None or hit[0] != key:
            w = p.detach().reshape(p.shape[0], -1)
            if hit is not None and hit[1][0].device == p.device:
                ops.split_bf16(w, out=hit[1])   # refresh in place: plane addresses stay valid (CUDA graphs hold them)
                hit = (key, hit[1])
            else:
                hit = (key, ops.split_bf16(w))
            self._planes[name] = hit
        return hit[1]

    def invalidate(self):
        """Marks every cached plane stale (keeps the buffers): used after a CUDA-graph replay changed the parameters
        without bumping their `_version`."""
        for name, (key, planes) in list(self._planes.items()):
            self._planes[name] = ((0, -1), planes)

    def refresh_all(self):
        """Re-splits every GEMM weight into its (persistent) planes unconditionally.  Called inside a CUDA-graph
        capture so that each replay picks up the weights the optimizer wrote since the previous one."""
        for name in self.GEMM_WEIGHTS:
            p = self._params()[name]
            hit = self._planes.get(name)
            w = p.detach().reshape(p.shape[0], -1)
            if hit is None:
                self._planes[name] = ((p.data_ptr(), p._version), ops.split_bf16(w))
            else:
                ops.split_bf16(w, out=hit[1])

    def sync_keys(self):
        """Declares every cached plane current for the parameters as they are now (after refresh_all(), or when an optimizer
        that writes the planes itself — optim.FusedPolyWarmupAdamW — owns them)."""
        for name, (key, planes) in list(self._planes.items()):
            if name.endswith("^T"):
                continue
            p = self._params()[name]
            self._planes[name] = ((p.data_ptr(), p._version), planes)

    def plane_t(self, name):
        """Transposed planes [K, N] of a Linear weight [N, K] (B operand of the dgrad GEMM), cached like plane()."""
        p = self._params()[name]
        key = (p.data_ptr(), p._version)
        hit = self._planes.get(name + "^T")
        if hit is None or hit[0] != key:
            hit = (key, transpose_planes(self.plane(name), p.shape[0], p[0].numel()))
            self._planes[name + "^T"] = hit
        return hit[1]

    def vec(self, name):
        return self._params()[name].detach()

    def pos(self, gh, gw):
        p = self.encoder.pos_embed
        key = (p.data_ptr(), p._version)
        hit = self._pos.get((gh, gw))
        if hit is None or hit[0] != key:
            hit = (key, ops.pos_embed_resize(p.detach(), gh, gw))
            self._pos[(gh, gw)] = hit
        return hit[1]


class KeptBlock:
    """Activations of one encoder block that the training backward reads (same fields as train._forward saves)."""
    __slots__ = ("x_in", "xn1", "qkv", "att", "lse", "x_mid", "xn2", "hid", "h_pre")


class KeptActivations:
    """What a no-grad encoder pass leaves behind for the leading `rows` token rows (the un-flipped scale-1.0 images of
    multi_scale_cam2_siamese): with no dropout and no batch statistics (vit.py:87-184, drop rates 0) those rows are
    bit-identical to the training forward `model(inputs)` of the same step (SURVEY §7 step 4 "dedupe"), so the
    training pass can start from them instead of recomputing 12 blocks."""
    __slots__ = ("rows", "patch_rows", "batch", "gh", "gw", "patch", "blocks", "tok_final")


def run_encoder(planes, images_per_seg, seg_sizes, flip_twin, aux_index, on_aux=None, keep_batch=0):
    """Runs the 12 blocks for every student in `planes` over the given segments.

    images_per_seg: list of source image tensors [b,3,H,W] (fp32, cuda); segment i is patchified from
    images_per_seg[i] resized to seg_sizes[i] = (hs, ws), with a flipped twin batch when flip_twin.
    aux_index: block index whose output feeds the aux head (embeds[aux_layer], vit.py:319-326).
    on_aux(g, tok_g, segs): called right after block `aux_index` (0-based) for every student when that
    index is not the last block (the last entry of `embeds` is the final-normed tensor, vit.py:323-324).
    keep_batch > 0: every block writes into buffers of its own (no reuse across blocks), also emits the attention
    log-sum-exp and, for the first keep_batch images of segment 0, the fc1 pre-activation; a KeptActivations per
    student is returned as third value.  Costs memory (0.8 GB per block and student at 21 976 rows), no extra pass.
    Returns (segs, [tok_g]) with tok_g the fp32 residual stream BEFORE the final LayerNorm.
    """
    G = len(planes)
    mult = 2 if flip_twin else 1
    segs, M, Mp = ops.make_segments([(mult * img.shape[0], hs // 16, ws // 16) for img, (hs, ws) in zip(images_per_seg, seg_sizes)])
    dev = images_per_seg[0].device
    bf = dict(dtype=torch.bfloat16, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    keep = keep_batch > 0
    kr = keep_batch * segs[0].tokens            # kept token rows
    kpr = keep_batch * segs[0].gh * segs[0].gw  # kept patch rows

    patch_hi = torch.empty(Mp, EMBED, **bf)
    patch_lo = torch.empty(Mp, EMBED, **bf)
    for s, img, size in zip(segs, images_per_seg, seg_sizes):
        ops.patchify(L.f32c(img), s, size, flip_twin, patch_hi, patch_lo)

    def planes_buf(cols):
        return [(torch.empty(M, cols, **bf), torch.empty(M, cols, **bf)) for _ in range(G)]

    tok = [torch.empty(M, EMBED, **f32) for _ in range(G)]
    if not keep:
        xn, qkv, att, hid = planes_buf(EMBED), planes_buf(3 * EMBED), planes_buf(EMBED), planes_buf(4 * EMBED)
    kept = None
    if keep:
        kept = [KeptActivations() for _ in range(G)]
        for k in kept:
            k.rows, k.patch_rows, k.batch, k.gh, k.gw = kr, kpr, keep_batch, segs[0].gh, segs[0].gw
            k.patch = (patch_hi[:kpr], patch_lo[:kpr])
            k.blocks = []

    pos = [[pl.pos(s.gh, s.gw) for s in segs] for pl in planes]
    ops.gemm_bf16x3(
        [dict(a=(patch_hi, patch_lo), w=pl.plane("patch_embed.proj.weight"), bias=pl.vec("patch_embed.proj.bias"),
              out_f32=tok[g], pos=pos[g]) for g, pl in enumerate(planes)],
        Mp, EMBED, EMBED, L.EPI_PATCH, segs=segs)
    for g, pl in enumerate(planes):
        ops.cls_rows(tok[g], pl.vec("cls_token").reshape(-1), pos[g], segs)

    scale = (EMBED // HEADS) ** -0.5
    for i in range(DEPTH):
        bp = f"blocks.{i}."
        if keep:
            xn1, qkv, att, xn2, hid = planes_buf(EMBED), planes_buf(3 * EMBED), planes_buf(EMBED), planes_buf(EMBED), planes_buf(4 * EMBED)
            lse = [torch.empty(M, HEADS, **f32) for _ in range(G)]
            h_pre = [torch.empty(kr, 4 * EMBED, **f32) for _ in range(G)]
            tok_mid = [torch.empty(M, EMBED, **f32) for _ in range(G)]
            tok_out = [torch.empty(M, EMBED, **f32) for _ in range(G)]
        else:
            xn1 = xn2 = xn
            lse = h_pre = [None] * G
            tok_mid = tok_out = tok
        for g, pl in enumerate(planes):
            ops.layernorm_split(tok[g], pl.vec(bp + "norm1.weight"), pl.vec(bp + "norm1.bias"), *xn1[g], eps=LN_EPS)
        ops.gemm_bf16x3([dict(a=xn1[g], w=pl.plane(bp + "attn.qkv.weight"), bias=pl.vec(bp + "attn.qkv.bias"), out=qkv[g])
                         for g, pl in enumerate(planes)], M, 3 * EMBED, EMBED, L.EPI_SPLIT)
        for g in range(G):
            ops.attention_fwd(qkv[g][0], qkv[g][1], att[g][0], att[g][1], segs, HEADS, scale, lse=lse[g])
        ops.gemm_bf16x3([dict(a=att[g], w=pl.plane(bp + "attn.proj.weight"), bias=pl.vec(bp + "attn.proj.bias"),
                              resid=tok[g], out_f32=tok_mid[g]) for g, pl in enumerate(planes)], M, EMBED, EMBED, L.EPI_RESID)
        for g, pl in enumerate(planes):
            ops.layernorm_split(tok_mid[g], pl.vec(bp + "norm2.weight"), pl.vec(bp + "norm2.bias"), *xn2[g], eps=LN_EPS)
        ops.gemm_bf16x3([dict(a=xn2[g], w=pl.plane(bp + "mlp.fc1.weight"), bias=pl.vec(bp + "mlp.fc1.bias"), out=hid[g],
                              out_f32=h_pre[g]) for g, pl in enumerate(planes)], M, 4 * EMBED, EMBED, L.EPI_GELU_SPLIT,
                        f32_rows=kr if keep else 0)
        ops.gemm_bf16x3([dict(a=hid[g], w=pl.plane(bp + "mlp.fc2.weight"), bias=pl.vec(bp + "mlp.fc2.bias"),
                              resid=tok_mid[g], out_f32=tok_out[g]) for g, pl in enumerate(planes)], M, EMBED, 4 * EMBED, L.EPI_RESID)
        if keep:
            for g in range(G):
                kb = KeptBlock()
                kb.x_in, kb.x_mid, kb.lse, kb.h_pre = tok[g][:kr], tok_mid[g][:kr], lse[g][:kr], h_pre[g]
                kb.xn1, kb.qkv, kb.att = (xn1[g][0][:kr], xn1[g][1][:kr]), (qkv[g][0][:kr], qkv[g][1][:kr]), (att[g][0][:kr], att[g][1][:kr])
                kb.xn2, kb.hid = (xn2[g][0][:kr], xn2[g][1][:kr]), (hid[g][0][:kr], hid[g][1][:kr])
                kept[g].blocks.append(kb)
        tok = tok_out
        if on_aux is not None and i == aux_index and i != DEPTH - 1:
            for g in range(G):
                on_aux(g, tok[g], segs)
    if keep:
        for g in range(G):
            kept[g].tok_final = tok[g][:kr]
        return segs, tok, kept
    return segs, tok
