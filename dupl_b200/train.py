"""Training-mode forward + backward of one student (`network.forward`, model/model_dupl.py:69-106 and
everything below it: vit.py:289-326, conv_head.py:33-41) as ONE torch.autograd.Function.

The Function takes the student's parameters as inputs and returns (cls_x4, seg, _x4, cls_aux); its
backward returns one gradient per parameter, so the reference script's `loss.backward()` and
DistributedDataParallel's reducer hooks work unchanged (unused `head.*` parameters stay unused, as in the
reference).  Every arithmetic step is a libdupl.so kernel: tcgen05 split-bf16 GEMMs for all forward,
dgrad and wgrad contractions, plus the HBM-bound kernels of train_kernels.cu.
"""
import ctypes as C
import os

import torch

from . import _lib as L
from . import encoder as E
from . import ops
from .dense import DECODER_DIL, _decoder_planes

D = E.EMBED


def _pad64(n):
    return (n + 63) // 64 * 64


def _mn_major():
    """wgrad / dgrad GEMMs read their transposed operands in place (MN-major tcgen05 operands: dupl_gemm_args.a_mn_major /
    b_mn_major) instead of from transposed copies.  DUPL_MN_MAJOR=0 restores the copies (same products in the same order:
    bit-identical gradients, tests/test_gpu_train.py)."""
    return os.environ.get("DUPL_MN_MAJOR", "1") != "0"


# ------------------------------------------------------------------ thin wrappers of the backward kernels
def _st(dev):
    return L.stream_ptr(dev)


def split_transpose(src, R, Cc, want_planes=True, want_t=True, tokens=0, np_=0, first=0, want_colsum=False, gelu_pre=None,
                    colsum_out=None):
    """fp32 rows -> (hi, lo) planes, transposed planes [Cc, pad64(R)] and (optionally, same pass) the column sums.
    gelu_pre: src is the gradient w.r.t. GELU's output; it is multiplied by GELU'(gelu_pre) on the way (no separate pass)."""
    dev = src.device
    bf = dict(dtype=torch.bfloat16, device=dev)
    Rpad = _pad64(R)
    hi = lo = thi = tlo = ws = cs = None
    if want_planes:
        hi, lo = torch.empty(R, Cc, **bf), torch.empty(R, Cc, **bf)
    if want_t:
        thi, tlo = torch.empty(Cc, Rpad, **bf), torch.empty(Cc, Rpad, **bf)
    if want_colsum:
        ws = torch.empty(((Rpad if want_t else R) + 63) // 64 * Cc, dtype=torch.float32, device=dev)
        cs = torch.empty(Cc, dtype=torch.float32, device=dev) if colsum_out is None else colsum_out.view(Cc)
    if gelu_pre is not None:
        if tokens or src.shape[1] != Cc or tuple(gelu_pre.shape) != (R, Cc):
            raise ValueError("split_transpose(gelu_pre=...) takes dense [R, Cc] operands")
        L.check(L.lib().dupl_split_transpose_gelu(L.ptr(src), L.ptr(gelu_pre), R, Cc, L.ptr(hi), L.ptr(lo), L.ptr(thi), L.ptr(tlo), Rpad,
                                                  L.ptr(ws), L.ptr(cs), _st(dev)), "dupl_split_transpose_gelu")
    else:
        L.check(L.lib().dupl_split_transpose(L.ptr(src), R, Cc, src.shape[1], tokens, np_, first, L.ptr(hi), L.ptr(lo),
                                             L.ptr(thi), L.ptr(tlo), Rpad, L.ptr(ws), L.ptr(cs), _st(dev)), "dupl_split_transpose")
    if want_colsum:
        return (hi, lo), (thi, tlo), cs
    return (hi, lo), (thi, tlo)


def transpose_planes(planes, R, Cc, tokens=0, np_=0, first=0):
    """(hi, lo) [*, Cc] -> (hi^T, lo^T) [Cc, pad64(R)]"""
    return E.transpose_planes(planes, R, Cc, tokens, np_, first)


def transpose_planes_multi(items):
    """items: [(planes (hi, lo) dense [R, Cc] bf16, R, Cc)] -> [(hi^T, lo^T) [Cc, pad64(R)]], ONE launch for up to 16 pairs."""
    outs, tab, dev = [], (L.TransposeItem * len(items))(), items[0][0][0].device
    for i, ((hi, lo), R, Cc) in enumerate(items):
        Rpad = _pad64(R)
        ohi = torch.empty(Cc, Rpad, dtype=torch.bfloat16, device=dev)
        olo = torch.empty(Cc, Rpad, dtype=torch.bfloat16, device=dev)
        e = tab[i]
        e.in_hi, e.in_lo, e.out_hi, e.out_lo = hi.data_ptr(), lo.data_ptr(), ohi.data_ptr(), olo.data_ptr()
        e.R, e.Cc, e.ld, e.Rpad = R, Cc, hi.shape[1], Rpad
        outs.append((ohi, olo))
    L.check(L.lib().dupl_transpose_planes_multi(tab, len(items), _st(dev)), "dupl_transpose_planes_multi")
    return outs


def colsum(x, R, Cc, tokens=0, np_=0, first=0):
    out = torch.empty(Cc, dtype=torch.float32, device=x.device)
    L.check(L.lib().dupl_colsum(L.ptr(x), R, Cc, x.shape[1], tokens, np_, first, L.ptr(out), _st(x.device)), "dupl_colsum")
    return out


def layernorm_bwd(dy, x, gamma, dres, dg_out=None, db_out=None):
    rows, cols = x.shape
    dev = x.device
    partial = torch.empty(2 * cols * ((rows + 7) // 8), dtype=torch.float32, device=dev)
    dg = torch.empty(cols, dtype=torch.float32, device=dev) if dg_out is None else dg_out.view(cols)
    db = torch.empty(cols, dtype=torch.float32, device=dev) if db_out is None else db_out.view(cols)
    L.check(L.lib().dupl_layernorm_bwd(L.ptr(dy), L.ptr(x), L.ptr(gamma), L.ptr(dres), L.ptr(partial), L.ptr(dg), L.ptr(db),
                                       rows, cols, E.LN_EPS, _st(dev)), "dupl_layernorm_bwd")
    return dg, db


def dgrad(dy_planes, wt_planes, M, n_out, k_contr):
    """dX[M, n_out] = dY[M, k_contr] @ W, with W^T given as planes [n_out, k_contr]."""
    out = torch.empty(M, n_out, dtype=torch.float32, device=dy_planes[0].device)
    ops.gemm_bf16x3([dict(a=dy_planes, w=wt_planes, out_f32=out)], M, n_out, k_contr, L.EPI_F32)
    return out


def wgrad(dyt_planes, xt_planes, n_rows, n_cols, k_contr, out=None):
    """dW[n_rows, n_cols] = dY^T @ X, operands as transposed planes [n_rows, k_contr], [n_cols, k_contr].
    out: fp32 buffer of n_rows*n_cols elements to write into (a gradient-arena view), else a fresh tensor."""
    out = torch.empty(n_rows, n_cols, dtype=torch.float32, device=dyt_planes[0].device) if out is None else out.view(n_rows, n_cols)
    ops.gemm_bf16x3([dict(a=dyt_planes, w=xt_planes, out_f32=out)], n_rows, n_cols, k_contr, L.EPI_F32, ksplit=L.MAX_KSPLIT)
    return out


def dgrad_multi(pairs, M, n_out, k_contr, w_in_place=False):
    """dgrad for several students in ONE grouped launch: pairs = [(dy_planes, wt_planes)] -> [dX].
    w_in_place: the second plane pair is the weight as stored, [k_contr, n_out] (consumed as an MN-major operand)."""
    outs = [torch.empty(M, n_out, dtype=torch.float32, device=dy[0].device) for dy, _ in pairs]
    ops.gemm_bf16x3([dict(a=dy, w=wt, out_f32=o) for (dy, wt), o in zip(pairs, outs)], M, n_out, k_contr, L.EPI_F32,
                    b_mn=w_in_place)
    return outs


def wgrad_multi(pairs, n_rows, n_cols, k_contr, outs=None, in_place=False):
    """wgrad for several students in ONE grouped launch: pairs = [(dyt_planes, xt_planes)], outs = per-student arena views or
    None -> [dW [n_rows, n_cols]].
    in_place: the pairs are (dy_planes [k_contr, n_rows], x_planes [k_contr, n_cols]) as the forward / backward left them
    (row stride = their own width); both are consumed as MN-major operands, k_contr need not be a multiple of 64."""
    outs = list(outs) if outs is not None else [None] * len(pairs)
    for i, (dyt, _) in enumerate(pairs):
        outs[i] = (torch.empty(n_rows, n_cols, dtype=torch.float32, device=dyt[0].device) if outs[i] is None
                   else outs[i].view(n_rows, n_cols))
    kw = {}
    if in_place:
        kw = dict(a_mn=True, b_mn=True, lda=pairs[0][0][0].shape[1], ldw=pairs[0][1][0].shape[1])
    ops.gemm_bf16x3([dict(a=dyt, w=xt, out_f32=o) for (dyt, xt), o in zip(pairs, outs)], n_rows, n_cols, k_contr, L.EPI_F32,
                    ksplit=L.MAX_KSPLIT, **kw)
    return outs


class _Saved:
    pass


# ------------------------------------------------------------------ forward with saved activations
def _forward(net, x, size=None, kept=None):
    """size: (hs, ws) to resize the input to first (bilinear, align_corners=False) — the 0.75x view of need_sp.
    kept: encoder.KeptActivations of the SAME images left by this step's MS-CAM pass: the 12 blocks are not run again."""
    L.require_cuda(x)
    x = L.f32c(x)
    B, _, H, W = x.shape
    hs, ws = size if size is not None else (H, W)
    pl = net.planes()
    dp = _decoder_planes(net)
    segs, M, Mp = ops.make_segments([(B, hs // 16, ws // 16)])
    sg = segs[0]
    gh, gw, np_, N = sg.gh, sg.gw, sg.gh * sg.gw, sg.tokens
    dev = x.device
    bf = dict(dtype=torch.bfloat16, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    S = _Saved()
    S.B, S.gh, S.gw, S.np, S.N, S.M, S.Mp, S.sg = B, gh, gw, np_, N, M, Mp, sg

    if kept is not None:
        if (kept.batch, kept.gh, kept.gw, kept.rows, kept.patch_rows) != (B, gh, gw, M, Mp):
            raise RuntimeError("kept activations do not match this batch")
        S.patch, S.blocks, tok = kept.patch, kept.blocks, kept.tok_final
        return _forward_heads(net, S, tok, pl, dp)
    S.patch = (torch.empty(Mp, D, **bf), torch.empty(Mp, D, **bf))
    ops.patchify(x, sg, (hs, ws), False, *S.patch)
    pos = [pl.pos(gh, gw)]
    tok = torch.empty(M, D, **f32)
    ops.gemm_bf16x3([dict(a=S.patch, w=pl.plane("patch_embed.proj.weight"), bias=pl.vec("patch_embed.proj.bias"),
                          out_f32=tok, pos=pos)], Mp, D, D, L.EPI_PATCH, segs=segs)
    ops.cls_rows(tok, pl.vec("cls_token").reshape(-1), pos, segs)

    scale = (D // E.HEADS) ** -0.5
    S.blocks = []
    aux_idx = net.encoder.aux_block_index()
    for i in range(E.DEPTH):
        bp = f"blocks.{i}."
        b = _Saved()
        b.x_in = tok
        b.xn1 = (torch.empty(M, D, **bf), torch.empty(M, D, **bf))
        ops.layernorm_split(tok, pl.vec(bp + "norm1.weight"), pl.vec(bp + "norm1.bias"), *b.xn1, eps=E.LN_EPS)
        b.qkv = (torch.empty(M, 3 * D, **bf), torch.empty(M, 3 * D, **bf))
        ops.gemm_bf16x3([dict(a=b.xn1, w=pl.plane(bp + "attn.qkv.weight"), bias=pl.vec(bp + "attn.qkv.bias"), out=b.qkv)],
                        M, 3 * D, D, L.EPI_SPLIT)
        b.att = (torch.empty(M, D, **bf), torch.empty(M, D, **bf))
        b.lse = torch.empty(M, E.HEADS, **f32)
        ops.attention_fwd(b.qkv[0], b.qkv[1], b.att[0], b.att[1], segs, E.HEADS, scale, lse=b.lse)
        b.x_mid = torch.empty(M, D, **f32)
        ops.gemm_bf16x3([dict(a=b.att, w=pl.plane(bp + "attn.proj.weight"), bias=pl.vec(bp + "attn.proj.bias"),
                              resid=tok, out_f32=b.x_mid)], M, D, D, L.EPI_RESID)
        b.xn2 = (torch.empty(M, D, **bf), torch.empty(M, D, **bf))
        ops.layernorm_split(b.x_mid, pl.vec(bp + "norm2.weight"), pl.vec(bp + "norm2.bias"), *b.xn2, eps=E.LN_EPS)
        b.hid = (torch.empty(M, 4 * D, **bf), torch.empty(M, 4 * D, **bf))
        b.h_pre = torch.empty(M, 4 * D, **f32)
        ops.gemm_bf16x3([dict(a=b.xn2, w=pl.plane(bp + "mlp.fc1.weight"), bias=pl.vec(bp + "mlp.fc1.bias"), out=b.hid,
                              out_f32=b.h_pre)], M, 4 * D, D, L.EPI_GELU_SPLIT, ldo=4 * D)
        tok = torch.empty(M, D, **f32)
        ops.gemm_bf16x3([dict(a=b.hid, w=pl.plane(bp + "mlp.fc2.weight"), bias=pl.vec(bp + "mlp.fc2.bias"),
                              resid=b.x_mid, out_f32=tok)], M, D, 4 * D, L.EPI_RESID)
        S.blocks.append(b)
    return _forward_heads(net, S, tok, pl, dp)


def _forward_heads(net, S, tok, pl, dp):
    """Final LayerNorm, LargeFOV decoder and the two GMP classifiers on top of the residual stream `tok`."""
    B, gh, gw, np_, N, M, Mp = S.B, S.gh, S.gw, S.np, S.N, S.M, S.Mp
    dev = tok.device
    bf = dict(dtype=torch.bfloat16, device=dev)
    f32 = dict(dtype=torch.float32, device=dev)
    aux_idx = net.encoder.aux_block_index()
    S.tok_final = tok
    S.aux_is_final = aux_idx == E.DEPTH - 1
    S.aux_idx = aux_idx

    # final norm, heads
    S.xn_planes = (torch.empty(M, D, **bf), torch.empty(M, D, **bf))
    S.xn = torch.empty(M, D, **f32)
    ops.layernorm_split(tok, pl.vec("norm.weight"), pl.vec("norm.bias"), *S.xn_planes, eps=E.LN_EPS, out_f32=S.xn)
    S.col6 = ops.im2col3x3(S.xn_planes[0], S.xn_planes[1], B, gh, gw, DECODER_DIL, 0, N, 1)
    S.h6 = (torch.empty(Mp, 512, **bf), torch.empty(Mp, 512, **bf))
    ops.gemm_bf16x3([dict(a=S.col6, w=dp.get("conv6"), out=S.h6)], Mp, 512, 9 * D, L.EPI_RELU_SPLIT)
    S.col7 = ops.im2col3x3(S.h6[0], S.h6[1], B, gh, gw, DECODER_DIL, 0, np_, 0)
    S.h7 = (torch.empty(Mp, 512, **bf), torch.empty(Mp, 512, **bf))
    ops.gemm_bf16x3([dict(a=S.col7, w=dp.get("conv7"), out=S.h7)], Mp, 512, 9 * 512, L.EPI_RELU_SPLIT)
    w8 = dp.get("conv8")
    S.n8 = w8[0].shape[0]
    seg_rows = torch.empty(Mp, S.n8, **f32)
    ops.gemm_bf16x3([dict(a=S.h7, w=w8, out_f32=seg_rows)], Mp, S.n8, 512, L.EPI_F32)
    seg = ops.rows_to_nchw(seg_rows, B, gh, gw, net.num_classes, 0, np_, 0)
    x4 = ops.rows_to_nchw(S.xn, B, gh, gw, D, 0, N, 1)
    K = net.num_classes - 1
    S.wc = L.f32c(net.classifier.weight.detach().reshape(K, -1))
    S.wa = L.f32c(net.aux_classifier.weight.detach().reshape(K, -1))
    cls_x4, S.arg_c = ops.gmp_classify(S.xn, S.wc, B, np_, 0, N, 1, want_argmax=True)
    S.aux_src = S.xn if S.aux_is_final else S.blocks[aux_idx + 1].x_in  # output of block aux_idx (un-normed)
    cls_aux, S.arg_a = ops.gmp_classify(S.aux_src, S.wa, B, np_, 0, N, 1, want_argmax=True)
    return (cls_x4, seg, x4, cls_aux), S


# ------------------------------------------------------------------ where the parameter gradients go
class _DictSink(dict):
    """Default: gradients are fresh tensors handed back to autograd (AccumulateGrad / DistributedDataParallel hooks)."""

    def out(self, name):
        return None

    def put(self, name, g):
        if g is not None:
            self[name] = g


class _ArenaSink:
    """Gradients land in the student's flat arena (grad_arena.GradArena): big wgrad GEMMs write their view in place, the rest
    is copied; autograd sees no parameter gradients.  `put` order == backward_order(net)."""

    def __init__(self, arena):
        self.arena = arena

    def out(self, name):
        return self.arena.out(name)

    def put(self, name, g):
        self.arena.put(name, g)


def backward_order(net):
    """Names of the trainable parameters of one student in the order `_backward` finishes their gradients."""
    aux_idx = net.encoder.aux_block_index()
    aux_final = aux_idx == E.DEPTH - 1
    order = ["classifier.weight"]
    if aux_final:
        order.append("aux_classifier.weight")
    order += ["decoder.conv8.weight", "decoder.conv7.weight", "decoder.conv6.weight", "encoder.norm.weight", "encoder.norm.bias"]
    for i in reversed(range(E.DEPTH)):
        bp = f"encoder.blocks.{i}."
        if not aux_final and i == aux_idx:
            order.append("aux_classifier.weight")
        order += [bp + "mlp.fc2.bias", bp + "mlp.fc2.weight", bp + "mlp.fc1.bias", bp + "mlp.fc1.weight", bp + "norm2.weight",
                  bp + "norm2.bias", bp + "attn.proj.bias", bp + "attn.proj.weight", bp + "attn.qkv.bias", bp + "attn.qkv.weight",
                  bp + "norm1.weight", bp + "norm1.bias"]
    order += ["encoder.patch_embed.proj.bias", "encoder.patch_embed.proj.weight", "encoder.cls_token"]
    return order


def make_grad_arena(net, chunk_elems=6 << 20, group=None):
    from .grad_arena import GradArena
    net._grad_arena = GradArena(trainable_parameters(net), backward_order(net), chunk_elems=chunk_elems, group=group)
    return net._grad_arena


# ------------------------------------------------------------------ backward
def _gmp_bwd(x_rows, w, dlogits, argmax, dx, S):
    B, K = dlogits.shape
    dev = dx.device
    dwp = torch.empty(B, K, D, dtype=torch.float32, device=dev)
    dw = torch.empty(K, D, dtype=torch.float32, device=dev)
    L.check(L.lib().dupl_gmp_classify_bwd(L.ptr(x_rows), L.ptr(w), L.ptr(L.f32c(dlogits)), L.ptr(argmax), L.ptr(dx), L.ptr(dwp),
                                          L.ptr(dw), B, S.np, D, K, D, S.N, 1, _st(dev)), "dupl_gmp_classify_bwd")
    return dw


def _conv_bwd_multi(d_outs, acts, cols, dec_planes, name, S, cin, d_ins, in_tokens, in_first, accumulate):
    """Backward of relu(conv3x3_d5(in)) for several students: d_outs fp32 [Mp, 512] each (grad wrt the relu output) ->
    d_ins (+)=, returns [dWmat [512, 9*cin]]; the dgrad and wgrad GEMMs are one grouped launch each."""
    dev = d_outs[0].device
    Mp = S.Mp
    mn = _mn_major()
    dpls, dts = [], []
    for d_out, act in zip(d_outs, acts):
        L.check(L.lib().dupl_relu_bwd(L.ptr(d_out), L.ptr(act[0]), L.ptr(act[1]), d_out.numel(), _st(dev)), "dupl_relu_bwd")
        dpl, dt = split_transpose(d_out, Mp, 512, want_t=not mn)
        dpls.append(dpl)
        dts.append(dt)
    if mn:
        dcols = dgrad_multi([(dpl, dp.get(name)) for dpl, dp in zip(dpls, dec_planes)], Mp, 9 * cin, 512, w_in_place=True)
    else:
        dcols = dgrad_multi([(dpl, dp.get_t(name)) for dpl, dp in zip(dpls, dec_planes)], Mp, 9 * cin, 512)
    for dcol, d_in in zip(dcols, d_ins):
        L.check(L.lib().dupl_col2im3x3(L.ptr(dcol), L.ptr(d_in), S.B, S.gh, S.gw, cin, DECODER_DIL, d_in.shape[1], in_tokens, in_first,
                                       1 if accumulate else 0, _st(dev)), "dupl_col2im3x3")
    if mn:
        return wgrad_multi(list(zip(dpls, cols)), 512, 9 * cin, Mp, in_place=True)
    col_ts = [transpose_planes(c, Mp, 9 * cin) for c in cols]
    return wgrad_multi(list(zip(dts, col_ts)), 512, 9 * cin, _pad64(Mp))


def _backward(net, S, g_cls, g_seg, g_x4, g_aux, sink=None):
    """One student.  -> sink (a dict name -> gradient by default)."""
    return _backward_multi([net], [S], [g_cls], [g_seg], [g_x4], [g_aux], [sink])[0]


def _backward_multi(nets, Ss, g_cls, g_seg, g_x4, g_aux, sinks):
    """Backward of `_forward` for one or both students in lock step: every dgrad / wgrad GEMM is ONE grouped launch over the
    students (twice the tiles per launch fill the 74 CTA pairs better: 495 vs 631 us per encoder block, tools/gemm_shapes.py
    train), everything else runs per student.  All arguments are lists over the students; every trainable parameter of every
    student is `put` exactly once, in backward_order().  -> list of sinks."""
    G = len(nets)
    R = range(G)
    pls = [n.planes() for n in nets]
    dps = [_decoder_planes(n) for n in nets]
    S0 = Ss[0]
    dev = S0.xn.device
    f32 = dict(dtype=torch.float32, device=dev)
    M, Mp, N, B, np_ = S0.M, S0.Mp, S0.N, S0.B, S0.np
    if any((S.M, S.Mp, S.N, S.B, S.aux_idx) != (M, Mp, N, B, S0.aux_idx) for S in Ss):
        raise RuntimeError("grouped backward needs students with identical shapes")
    Mpad = _pad64(M)
    mn = _mn_major()
    grads = [s if s is not None else _DictSink() for s in sinks]
    K = nets[0].num_classes - 1

    d_xn = [torch.zeros(M, D, **f32) for _ in R]          # grad wrt the final-normed tokens
    for g in R:
        S = Ss[g]
        if g_x4[g] is not None:
            L.check(L.lib().dupl_nchw_to_rows_add(L.ptr(L.f32c(g_x4[g])), L.ptr(d_xn[g]), B, np_, D, D, N, 1, _st(dev)), "dupl_nchw_to_rows_add")
        grads[g].put("classifier.weight", None if g_cls[g] is None else _gmp_bwd(S.xn, S.wc, g_cls[g], S.arg_c, d_xn[g], S).reshape(K, D, 1, 1))
        if S.aux_is_final:
            grads[g].put("aux_classifier.weight",
                         None if g_aux[g] is None else _gmp_bwd(S.aux_src, S.wa, g_aux[g], S.arg_a, d_xn[g], S).reshape(K, D, 1, 1))
    with_seg = [g for g in R if g_seg[g] is not None]
    if with_seg:
        # conv8 (1x1): seg_rows = h7 @ W8^T
        Cn = nets[0].num_classes
        Cp = _pad64(Cn)                             # contraction dim of the dgrad padded to a multiple of 64 (21 -> 64, 81 -> 128)
        dsps, dsts, w8ts, h7ts = [], [], [], []
        for g in with_seg:
            S = Ss[g]
            d_seg_rows = torch.zeros(Mp, Cp, **f32)
            L.check(L.lib().dupl_nchw_to_rows_add(L.ptr(L.f32c(g_seg[g])), L.ptr(d_seg_rows), B, np_, Cn, Cp, 0, 0, _st(dev)), "dupl_nchw_to_rows_add")
            dsp, dst = split_transpose(d_seg_rows, Mp, Cp)
            dsps.append(dsp)
            dsts.append(dst)
            w8ts.append(transpose_planes(dps[g].get("conv8"), S.n8, 512))                      # [512, pad64(n8)]
            h7ts.append(transpose_planes(S.h7, Mp, 512))
        d_h7 = dgrad_multi(list(zip(dsps, w8ts)), Mp, 512, Cp)
        dw8 = wgrad_multi(list(zip(dsts, h7ts)), Cp, 512, _pad64(Mp))
        for i, g in enumerate(with_seg):
            grads[g].put("decoder.conv8.weight", dw8[i][:Cn].reshape(Cn, 512, 1, 1))
        # conv7 + relu, conv6 + relu
        d_h6 = [torch.empty(Mp, 512, **f32) for _ in with_seg]
        dw7 = _conv_bwd_multi(d_h7, [Ss[g].h7 for g in with_seg], [Ss[g].col7 for g in with_seg], [dps[g] for g in with_seg], "conv7",
                              S0, 512, d_h6, 0, 0, False)
        for i, g in enumerate(with_seg):
            grads[g].put("decoder.conv7.weight", dw7[i].reshape(512, 3, 3, 512).permute(0, 3, 1, 2))
        dw6 = _conv_bwd_multi(d_h6, [Ss[g].h6 for g in with_seg], [Ss[g].col6 for g in with_seg], [dps[g] for g in with_seg], "conv6",
                              S0, D, [d_xn[g] for g in with_seg], N, 1, True)
        for i, g in enumerate(with_seg):
            grads[g].put("decoder.conv6.weight", dw6[i].reshape(512, 3, 3, D).permute(0, 3, 1, 2))
    for g in R:
        if g_seg[g] is None:
            for n in ("decoder.conv8.weight", "decoder.conv7.weight", "decoder.conv6.weight"):
                grads[g].put(n, None)

    # final LayerNorm
    d_tok = [torch.zeros(M, D, **f32) for _ in R]
    for g in R:
        dg, db = layernorm_bwd(d_xn[g], Ss[g].tok_final, pls[g].vec("norm.weight"), d_tok[g],
                               grads[g].out("encoder.norm.weight"), grads[g].out("encoder.norm.bias"))
        grads[g].put("encoder.norm.weight", dg)
        grads[g].put("encoder.norm.bias", db)

    scale = (D // E.HEADS) ** -0.5
    bfk = dict(dtype=torch.bfloat16, device=dev)
    for i in reversed(range(E.DEPTH)):
        bp = f"blocks.{i}."
        ep = "encoder." + bp
        bl = [S.blocks[i] for S in Ss]
        for g in R:
            S = Ss[g]
            if not S.aux_is_final and i == S.aux_idx:
                # cls_aux reads the output of this block: its gradient joins the residual-stream gradient here
                grads[g].put("aux_classifier.weight",
                             None if g_aux[g] is None else _gmp_bwd(S.aux_src, S.wa, g_aux[g], S.arg_a, d_tok[g], S).reshape(K, D, 1, 1))
        w_names = ("mlp.fc2", "mlp.fc1", "attn.proj", "attn.qkv")
        if mn:
            # MN-major operands: the wgrad GEMMs read dY and the saved activation planes, the dgrad GEMMs the weight planes,
            # exactly as they are stored
            act_t = [dict(hid=bl[g].hid, xn2=bl[g].xn2, att=bl[g].att, xn1=bl[g].xn1) for g in R]
            w_t = [{wn: pls[g].plane(bp + wn + ".weight") for wn in w_names} for g in R]
        else:
            # every transposed operand of this block's wgrad (activations) and dgrad (weights) GEMMs, all students: ONE launch
            items = []
            for g in R:
                items += [(bl[g].hid, M, 4 * D), (bl[g].xn2, M, D), (bl[g].att, M, D), (bl[g].xn1, M, D)]
            for g in R:
                for wn in w_names:
                    wp = pls[g].plane(bp + wn + ".weight")
                    items.append((wp, wp[0].shape[0], wp[0].shape[1]))
            tr = transpose_planes_multi(items)
            act_t = [dict(zip(("hid", "xn2", "att", "xn1"), tr[4 * g:4 * g + 4])) for g in R]
            w_t = [dict(zip(w_names, tr[4 * G + 4 * g:4 * G + 4 * g + 4])) for g in R]

        def wgrad_of(act, n_rows, n_cols, pname):
            """dW of one layer for all students: dY = dpl (in place) or dt (transposed copy), X = the saved activation."""
            dw = wgrad_multi([((dpl if mn else dt)[g], act_t[g][act]) for g in R], n_rows, n_cols, M if mn else Mpad,
                             outs=[grads[g].out(ep + pname) for g in R], in_place=mn)
            for g in R:
                grads[g].put(ep + pname, dw[g])

        # ---- MLP: x_out = x_mid + fc2(gelu(fc1(LN2(x_mid))))
        dpl, dt = [None] * G, [None] * G
        for g in R:
            dpl[g], dt[g], cs = split_transpose(d_tok[g], M, D, want_t=not mn, want_colsum=True,
                                                colsum_out=grads[g].out(ep + "mlp.fc2.bias"))
            grads[g].put(ep + "mlp.fc2.bias", cs)
        wgrad_of("hid", D, 4 * D, "mlp.fc2.weight")
        d_hid = dgrad_multi([(dpl[g], w_t[g]["mlp.fc2"]) for g in R], M, 4 * D, D, w_in_place=mn)
        for g in R:
            # GELU'(fc1 pre-activation) is applied inside the split / transpose pass (dupl_split_transpose_gelu)
            dpl[g], dt[g], cs = split_transpose(d_hid[g], M, 4 * D, want_t=not mn, want_colsum=True, gelu_pre=bl[g].h_pre,
                                                colsum_out=grads[g].out(ep + "mlp.fc1.bias"))
            grads[g].put(ep + "mlp.fc1.bias", cs)
        wgrad_of("xn2", 4 * D, D, "mlp.fc1.weight")
        d_xn2 = dgrad_multi([(dpl[g], w_t[g]["mlp.fc1"]) for g in R], M, D, 4 * D, w_in_place=mn)
        for g in R:
            dg, db = layernorm_bwd(d_xn2[g], bl[g].x_mid, pls[g].vec(bp + "norm2.weight"), d_tok[g],
                                   grads[g].out(ep + "norm2.weight"), grads[g].out(ep + "norm2.bias"))
            grads[g].put(ep + "norm2.weight", dg)
            grads[g].put(ep + "norm2.bias", db)
        # ---- attention: x_mid = x_in + proj(attn(qkv(LN1(x_in))))
        for g in R:
            dpl[g], dt[g], cs = split_transpose(d_tok[g], M, D, want_t=not mn, want_colsum=True,
                                                colsum_out=grads[g].out(ep + "attn.proj.bias"))
            grads[g].put(ep + "attn.proj.bias", cs)
        wgrad_of("att", D, D, "attn.proj.weight")
        d_att = [(torch.empty(M, D, **bfk), torch.empty(M, D, **bfk)) for _ in R]   # dO as split planes: operand of the attention backward
        ops.gemm_bf16x3([dict(a=dpl[g], w=w_t[g]["attn.proj"], out=d_att[g]) for g in R], M, D, D, L.EPI_SPLIT, b_mn=mn)
        for g in R:
            d_qkv = ops.attention_bwd(bl[g].qkv, bl[g].att, d_att[g], bl[g].lse, B, N, E.HEADS, scale)
            dpl[g], dt[g], cs = split_transpose(d_qkv, M, 3 * D, want_t=not mn, want_colsum=True,
                                                colsum_out=grads[g].out(ep + "attn.qkv.bias"))
            grads[g].put(ep + "attn.qkv.bias", cs)
        wgrad_of("xn1", 3 * D, D, "attn.qkv.weight")
        d_xn1 = dgrad_multi([(dpl[g], w_t[g]["attn.qkv"]) for g in R], M, D, 3 * D, w_in_place=mn)
        for g in R:
            dg, db = layernorm_bwd(d_xn1[g], bl[g].x_in, pls[g].vec(bp + "norm1.weight"), d_tok[g],
                                   grads[g].out(ep + "norm1.weight"), grads[g].out(ep + "norm1.bias"))
            grads[g].put(ep + "norm1.weight", dg)
            grads[g].put(ep + "norm1.bias", db)

    # ---- patch embedding (pos_embed is frozen, vit.py:243)
    dts = []
    for g in R:
        _, dt_g, cs = split_transpose(d_tok[g], Mp, D, want_planes=False, tokens=N, np_=np_, first=1, want_colsum=True)
        grads[g].put("encoder.patch_embed.proj.bias", cs)
        dts.append(dt_g)
    dwpe = wgrad_multi([(dts[g], transpose_planes(Ss[g].patch, Mp, D)) for g in R], D, D, _pad64(Mp),
                       outs=[grads[g].out("encoder.patch_embed.proj.weight") for g in R])
    for g in R:
        grads[g].put("encoder.patch_embed.proj.weight", dwpe[g].reshape(D, 3, 16, 16))
        grads[g].put("encoder.cls_token", colsum(d_tok[g], B, D, tokens=N, np_=1, first=0).reshape(1, 1, D))
    return grads


def trainable_parameters(net):
    """(name, parameter) pairs that take part in the forward pass, in a fixed order."""
    out = []
    for name, p in net.named_parameters():
        if name.startswith("encoder.head.") or not p.requires_grad:
            continue
        out.append((name, p))
    return out


class StudentFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, net, x, size, *params):
        ctx.set_materialize_grads(False)
        kept = None
        if size is None and getattr(net, "_use_kept", False) and net._kept is not None:
            kept, net._use_kept = net._kept, False
        outs, S = _forward(net, x, size, kept)
        ctx.net, ctx.S = net, S
        ctx.names = [(n, tuple(p.shape)) for n, p in trainable_parameters(net)]
        ctx.n_inputs = len(params)
        return outs

    @staticmethod
    def backward(ctx, g_cls, g_seg, g_x4, g_aux):
        arena = getattr(ctx.net, "_grad_arena", None) if getattr(ctx.net, "_use_arena", False) else None
        with torch.no_grad():
            if arena is not None:
                # gradients go straight into the student's flat arena (and from there, chunk by chunk, into the
                # overlapped all-reduce); the parameters' .grad are views of it, autograd gets nothing to accumulate
                _backward(ctx.net, ctx.S, g_cls, g_seg, g_x4, g_aux, sink=_ArenaSink(arena))
                arena.end_call()
                ctx.S = None
                return (None, None, None, *([None] * ctx.n_inputs))
            grads = _backward(ctx.net, ctx.S, g_cls, g_seg, g_x4, g_aux)
        ctx.S = None
        out = []
        for n, shape in ctx.names:
            g = grads.get(n)
            out.append(None if g is None else g.reshape(shape).contiguous())
        return (None, None, None, *out)


class PairFunction(torch.autograd.Function):
    """Both students of `siamese_network.forward(x)` (model_dupl.py:207-209) as ONE autograd node — arena mode only.  Its
    backward runs the two students in lock step (_backward_multi), so every dgrad / wgrad GEMM of a layer is one grouped launch
    over both students instead of two launches that each leave part of the 74 CTA pairs idle."""

    @staticmethod
    def forward(ctx, net1, net2, x, anchor):
        ctx.set_materialize_grads(False)
        outs, Ss = [], []
        for net in (net1, net2):
            kept = None
            if getattr(net, "_use_kept", False) and net._kept is not None:
                kept, net._use_kept = net._kept, False
            o, S = _forward(net, x, None, kept)
            outs.append(o)
            Ss.append(S)
        ctx.nets, ctx.Ss = (net1, net2), Ss
        return (*outs[0], *outs[1])

    @staticmethod
    def backward(ctx, *g):
        arenas = [n._grad_arena for n in ctx.nets]
        with torch.no_grad():
            _backward_multi(list(ctx.nets), ctx.Ss, [g[0], g[4]], [g[1], g[5]], [g[2], g[6]], [g[3], g[7]],
                            [_ArenaSink(a) for a in arenas])
            for a in arenas:
                a.end_call()
        ctx.Ss = None
        return (None, None, None, None)


def pair_student_forward(net1, net2, x):
    """-> ((cls_x4, seg, _x4, cls_aux) of student 1, same of student 2); both nets must be in arena mode."""
    anchor = torch.empty(0, device=x.device).requires_grad_()
    o = PairFunction.apply(net1, net2, x, anchor)
    return tuple(o[:4]), tuple(o[4:])


def _arena_mode(net):
    return getattr(net, "_use_arena", False) and getattr(net, "_grad_arena", None) is not None


def student_forward(net, x, size=None):
    if _arena_mode(net):
        # arena mode: the backward writes the parameter gradients into the student's flat arena itself, so the parameters
        # are not autograd inputs at all.  A fresh leaf per call keeps the outputs differentiable; being new, its
        # AccumulateGrad node belongs to the stream of THIS forward (a node that survives from an earlier step would make a
        # capturing stream wait for the stream it was created on).
        anchor = torch.empty(0, device=x.device).requires_grad_()
        return StudentFunction.apply(net, x, size, anchor)
    params = [p for _, p in trainable_parameters(net)]
    return StudentFunction.apply(net, x, size, *params)
