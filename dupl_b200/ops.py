"""Tensor-level wrappers over the C ABI: allocate outputs with torch, pass raw pointers, launch on the
current stream.  No arithmetic happens in Python/PyTorch here."""
import ctypes as C

import torch

from . import _lib as L


def _segs(seg_list):
    arr = (L.Segment * L.MAX_SEGMENTS)()
    for i, s in enumerate(seg_list):
        arr[i] = s
    return arr


def make_segments(shapes):
    """shapes: list of (batch, gh, gw) -> (list[Segment], total token rows, total patch rows)."""
    if not 1 <= len(shapes) <= L.MAX_SEGMENTS:
        raise ValueError(f"1..{L.MAX_SEGMENTS} segments supported, got {len(shapes)}")
    segs, row, prow = [], 0, 0
    for (b, gh, gw) in shapes:
        tokens = 1 + gh * gw
        segs.append(L.Segment(b, gh, gw, tokens, row, prow))
        row += b * tokens
        prow += b * gh * gw
    return segs, row, prow


# ------------------------------------------------------------------ dense path
def split_bf16(x, out=None):
    L.require_cuda(x)
    x = L.f32c(x)
    if out is not None:
        hi, lo = out
    else:
        hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        lo = torch.empty_like(hi)
    L.check(L.lib().dupl_split_bf16(L.ptr(x), L.ptr(hi), L.ptr(lo), x.numel(), L.stream_ptr(x.device)), "dupl_split_bf16")
    return hi, lo


def gemm_bf16x3(groups, M, N, K, epilogue, lda=None, ldo=None, segs=None, ldw=None, ksplit=0, f32_rows=0, passes=3,
                a_mn=False, b_mn=False):
    """groups: list of dicts with keys a=(hi,lo), w=(hi,lo), bias, resid, out_f32, out=(hi,lo), pos=[...].
    ksplit > 1 lets the library split the contraction up to that many ways (plain F32 epilogue, no bias): the
    partial sums go through a workspace allocated here and are added in a fixed order.
    a_mn / b_mn: the operand planes are stored transposed ([K, M] resp. [K, N]) and are consumed in place."""
    a = L.GemmArgs()
    a.groups, a.M, a.N, a.K = len(groups), M, N, K
    a.a_mn_major, a.b_mn_major = int(a_mn), int(b_mn)
    a.lda = (M if a_mn else K) if lda is None else lda
    a.ldo = N if ldo is None else ldo
    a.ldw = 0 if ldw is None else ldw
    a.max_ksplit = ksplit
    a.f32_rows = f32_rows
    a.passes = passes
    a.epilogue = epilogue
    ws = []
    a.nseg = 0
    if segs is not None:
        a.nseg = len(segs)
        for i, s in enumerate(segs):
            a.seg[i] = s
    dev = None
    for gi, g in enumerate(groups):
        G = a.g[gi]
        dev = g["a"][0].device
        G.a_hi, G.a_lo = g["a"][0].data_ptr(), g["a"][1].data_ptr()
        G.w_hi, G.w_lo = g["w"][0].data_ptr(), g["w"][1].data_ptr()
        G.bias = g["bias"].data_ptr() if g.get("bias") is not None else None
        G.resid = g["resid"].data_ptr() if g.get("resid") is not None else None
        G.out_f32 = g["out_f32"].data_ptr() if g.get("out_f32") is not None else None
        if g.get("out") is not None:
            G.out_hi, G.out_lo = g["out"][0].data_ptr(), g["out"][1].data_ptr()
        for i, p in enumerate(g.get("pos") or []):
            G.pos[i] = p.data_ptr()
        if ksplit > 1:
            ws.append(torch.empty(ksplit * M * a.ldo, dtype=torch.float32, device=dev))
            G.splitk_ws = ws[-1].data_ptr()
    L.check(L.lib().dupl_gemm_bf16x3(C.byref(a), L.stream_ptr(dev)), "dupl_gemm_bf16x3")


def layernorm_split(x, gamma, beta, out_hi, out_lo, eps=1e-6, out_f32=None):
    rows, cols = x.shape
    L.check(L.lib().dupl_layernorm_split(L.ptr(x), L.ptr(gamma), L.ptr(beta), L.ptr(out_hi), L.ptr(out_lo), L.ptr(out_f32),
                                         rows, cols, eps, L.stream_ptr(x.device)), "dupl_layernorm_split")


def im2col3x3(in_hi, in_lo, B, gh, gw, dilation, row_offset, row_stride, first):
    cin = in_hi.shape[1]
    out_hi = torch.empty(B * gh * gw, 9 * cin, dtype=torch.bfloat16, device=in_hi.device)
    out_lo = torch.empty_like(out_hi)
    L.check(L.lib().dupl_im2col3x3(L.ptr(in_hi), L.ptr(in_lo), L.ptr(out_hi), L.ptr(out_lo), B, gh, gw, cin, dilation,
                                   row_offset, row_stride, first, L.stream_ptr(in_hi.device)), "dupl_im2col3x3")
    return out_hi, out_lo


def rows_to_nchw(src, B, gh, gw, Cc, row_offset, row_stride, first):
    out = torch.empty(B, Cc, gh, gw, dtype=torch.float32, device=src.device)
    L.check(L.lib().dupl_rows_to_nchw(L.ptr(src), L.ptr(out), B, gh * gw, Cc, src.shape[1], row_offset, row_stride, first,
                                      L.stream_ptr(src.device)), "dupl_rows_to_nchw")
    return out


def gmp_classify(x, w, B, np_, row_offset, row_stride, first, want_argmax=False):
    K, D = w.shape
    logits = torch.empty(B, K, dtype=torch.float32, device=x.device)
    arg = torch.empty(B, D, dtype=torch.int32, device=x.device) if want_argmax else None
    pooled = torch.empty(B, D, dtype=torch.float32, device=x.device)
    L.check(L.lib().dupl_gmp_classify(L.ptr(x), L.ptr(w), L.ptr(logits), L.ptr(pooled), L.ptr(arg), B, np_, D, K, row_offset, row_stride,
                                      first, L.stream_ptr(x.device)), "dupl_gmp_classify")
    return (logits, arg) if want_argmax else logits


def attention_fwd(qkv_hi, qkv_lo, out_hi, out_lo, segs, heads, scale, lse=None):
    a = L.AttentionArgs()
    a.nseg = len(segs)
    for i, s in enumerate(segs):
        a.seg[i] = s
    a.M, a.heads, a.scale = qkv_hi.shape[0], heads, scale
    a.qkv_hi, a.qkv_lo = qkv_hi.data_ptr(), qkv_lo.data_ptr()
    a.out_hi, a.out_lo = out_hi.data_ptr(), out_lo.data_ptr()
    a.lse = lse.data_ptr() if lse is not None else None
    L.check(L.lib().dupl_attention_fwd(C.byref(a), L.stream_ptr(qkv_hi.device)), "dupl_attention_fwd")


def attention_bwd(qkv, att, d_att, lse, batch, tokens, heads, scale, row_offset=0):
    """Backward of softmax(Q K^T * scale) V per (image, head) (autograd of vit.py:120-135) for `batch` images of `tokens`
    rows starting at `row_offset`.  qkv / att / d_att are (hi, lo) split-bf16 plane pairs [M, 3D] / [M, D] / [M, D], lse the
    forward's log-sum-exp [M, heads].  -> d_qkv fp32 [M, 3D]."""
    M, D = att[0].shape
    dev = att[0].device
    d_qkv = torch.empty(M, 3 * D, dtype=torch.float32, device=dev)
    dvec = torch.empty(M, heads, dtype=torch.float32, device=dev)
    a = L.AttentionBwdArgs()
    a.qkv_hi, a.qkv_lo, a.o_hi, a.o_lo = qkv[0].data_ptr(), qkv[1].data_ptr(), att[0].data_ptr(), att[1].data_ptr()
    a.do_hi, a.do_lo = d_att[0].data_ptr(), d_att[1].data_ptr()
    a.lse, a.Dvec, a.dqkv = lse.data_ptr(), dvec.data_ptr(), d_qkv.data_ptr()
    a.M, a.batch, a.tokens, a.row_offset, a.heads, a.scale = M, batch, tokens, row_offset, heads, scale
    L.check(L.lib().dupl_attention_bwd(C.byref(a), L.stream_ptr(dev)), "dupl_attention_bwd")
    return d_qkv


def patchify(images, seg, size, flip_twin, out_hi, out_lo):
    """size = (hs, ws): the resolution the images are resized to before the 16x16 patch grid is cut."""
    b, _, H, W = images.shape
    L.check(L.lib().dupl_patchify(L.ptr(images), b, H, W, C.byref(seg), size[0], size[1], 1 if flip_twin else 0, L.ptr(out_hi),
                                  L.ptr(out_lo), L.stream_ptr(images.device)), "dupl_patchify")


def pos_embed_resize(pos_embed, gh, gw):
    """pos_embed [1, 1+S*S, D] -> [1+gh*gw, D]"""
    pe = L.f32c(pos_embed).reshape(pos_embed.shape[-2], pos_embed.shape[-1])
    S = int(round((pe.shape[0] - 1) ** 0.5))
    out = torch.empty(1 + gh * gw, pe.shape[1], dtype=torch.float32, device=pe.device)
    L.check(L.lib().dupl_pos_embed_resize(L.ptr(pe), L.ptr(out), S, gh, gw, pe.shape[1], L.stream_ptr(pe.device)),
            "dupl_pos_embed_resize")
    return out


def cls_rows(tok, cls_token, pos_list, segs):
    arr = (C.c_void_p * len(segs))(*[p.data_ptr() for p in pos_list])
    L.check(L.lib().dupl_cls_rows(L.ptr(tok), L.ptr(cls_token), arr, _segs(segs), len(segs), tok.shape[1],
                                  L.stream_ptr(tok.device)), "dupl_cls_rows")


def cam_contract(tok, gamma, beta, w, segs, eps=1e-6):
    """-> list of [batch, K, gh, gw] tensors (views of one buffer), one per segment."""
    K, D = w.shape[0], tok.shape[1]
    sizes = [s.batch * K * s.gh * s.gw for s in segs]
    out = torch.empty(sum(sizes), dtype=torch.float32, device=tok.device)
    offs, o = [], 0
    for n in sizes:
        offs.append(o)
        o += n
    off_arr = (C.c_int64 * len(segs))(*offs)
    L.check(L.lib().dupl_cam_contract(L.ptr(tok), L.ptr(gamma), L.ptr(beta), eps, L.ptr(w), K, D, _segs(segs), len(segs),
                                      L.ptr(out), off_arr, L.stream_ptr(tok.device)), "dupl_cam_contract")
    return [out[o:o + n].view(s.batch, K, s.gh, s.gw) for o, n, s in zip(offs, sizes, segs)]


# ------------------------------------------------------------------ CAM post / labels
def mscam_post(lowres, b, H, W):
    """lowres: list of [2b, K, gh, gw] (scale 1.0 first) -> [b, K, H, W]"""
    K = lowres[0].shape[1]
    dev = lowres[0].device
    a = L.MscamArgs()
    a.nscale = len(lowres)
    keep = []
    for i, t in enumerate(lowres):
        t = L.f32c(t)
        keep.append(t)
        if t.shape[0] != 2 * b or t.shape[1] != K:
            raise ValueError("mscam_post: every scale must be [2b, K, gh, gw]")
        a.lowres[i] = t.data_ptr()
        a.gh[i], a.gw[i] = t.shape[2], t.shape[3]
    out = torch.empty(b, K, H, W, dtype=torch.float32, device=dev)
    mm = torch.empty(b * K * 2, dtype=torch.float32, device=dev)
    a.b, a.K, a.H, a.W = b, K, H, W
    a.out, a.minmax = out.data_ptr(), mm.data_ptr()
    L.check(L.lib().dupl_mscam_post(C.byref(a), L.stream_ptr(dev)), "dupl_mscam_post")
    return out


def box_to_device(img_box, device):
    """img_box [b,4] (int16 CPU tensor in the reference's loader) -> int32 device tensor."""
    t = torch.as_tensor(img_box)
    return t.to(device=device, dtype=torch.int32, non_blocking=True).contiguous()


def cam_to_label(cam, cls_label, img_box, bkg_thre, high_thre, low_thre, ignore_mid, ignore_index, want_valid):
    L.require_cuda(cam)
    cam = L.f32c(cam)
    b, K, h, w = cam.shape
    dev = cam.device
    cls = L.f32c(cls_label.to(dev))
    a = L.CamToLabelArgs()
    a.cam, a.cls_label = cam.data_ptr(), cls.data_ptr()
    box = None
    if img_box is not None:
        box = box_to_device(img_box, dev)
        a.img_box = box.data_ptr()
    ht = None
    if torch.is_tensor(high_thre):
        ht = L.f32c(high_thre.to(dev)).reshape(-1)
        if ht.numel() != b:
            raise ValueError("high_thre tensor must have one entry per image")
        a.high_thre = ht.data_ptr()
    else:
        a.high_thre_scalar = float(high_thre) if high_thre is not None else 0.0
    a.low_thre = float(low_thre) if low_thre is not None else 0.0
    a.bkg_thre = float(bkg_thre)
    a.ignore_mid = 1 if ignore_mid else 0
    a.ignore_index = int(ignore_index) if ignore_index is not None else 0
    a.b, a.K, a.h, a.w = b, K, h, w
    valid = torch.empty_like(cam) if want_valid else None
    label = torch.empty(b, h, w, dtype=torch.int64, device=dev)
    a.valid_cam = valid.data_ptr() if valid is not None else None
    a.label = label.data_ptr()
    L.check(L.lib().dupl_cam_to_label(C.byref(a), L.stream_ptr(dev)), "dupl_cam_to_label")
    return valid, label


def label_to_aff_mask(label, ignore_index):
    L.require_cuda(label)
    lab = label.to(torch.int64).contiguous()
    b = lab.shape[0]
    n = lab[0].numel()
    aff = torch.empty(b, n, n, dtype=torch.int64, device=lab.device)
    L.check(L.lib().dupl_label_to_aff_mask(L.ptr(lab), L.ptr(aff), b, n, int(ignore_index), L.stream_ptr(lab.device)),
            "dupl_label_to_aff_mask")
    return aff


# ------------------------------------------------------------------ PAR / refine
def _dil(dilations):
    return (C.c_int32 * len(dilations))(*[int(d) for d in dilations]), len(dilations)


def par_affinity(imgs, dilations, w1=0.3, w2=0.01):
    imgs = L.f32c(imgs)
    B, Cc, h, w = imgs.shape
    arr, nd = _dil(dilations)
    aff = torch.empty(B, 8 * nd, h, w, dtype=torch.float32, device=imgs.device)
    L.check(L.lib().dupl_par_affinity(L.ptr(imgs), L.ptr(aff), B, Cc, h, w, arr, nd, w1, w2, L.stream_ptr(imgs.device)),
            "dupl_par_affinity")
    return aff


def par_propagate(aff, masks, dilations, num_iter, nactive=None):
    """masks [B,P,h,w] fp32 contiguous (consumed as scratch) -> propagated masks [B,P,h,w]."""
    B, P, h, w = masks.shape
    arr, nd = _dil(dilations)
    scratch = torch.empty_like(masks) if nactive is None else torch.zeros_like(masks)
    where = C.c_int32(0)
    L.check(L.lib().dupl_par_propagate(L.ptr(aff), L.ptr(masks), L.ptr(scratch), L.ptr(nactive), B, P, h, w, arr, nd,
                                       int(num_iter), C.byref(where), L.stream_ptr(masks.device)), "dupl_par_propagate")
    return scratch if where.value else masks


def refine_prologue(images, cams, cls_label, bkg_h, bkg_l):
    images, cams = L.f32c(images), L.f32c(cams)
    b, K, H, W = cams.shape
    dev = cams.device
    cls = L.f32c(cls_label.to(dev))
    a = L.RefinePrologueArgs()
    a.images, a.cams, a.cls_label = images.data_ptr(), cams.data_ptr(), cls.data_ptr()
    keep = None
    if torch.is_tensor(bkg_h):
        keep = L.f32c(bkg_h.to(dev))
        if keep.numel() != b * H * W:
            raise ValueError("high_thre_map must be [b,1,H,W]")
        a.bkg_h = keep.data_ptr()
    else:
        a.bkg_h_scalar = float(bkg_h)
    a.bkg_l_scalar = float(bkg_l)
    a.b, a.K, a.H, a.W = b, K, H, W
    images_ds = torch.empty(b, 3, H // 2, W // 2, dtype=torch.float32, device=dev)
    masks = torch.empty(b, 2 * (K + 1), H // 2, W // 2, dtype=torch.float32, device=dev)
    nactive = torch.empty(b, dtype=torch.int32, device=dev)
    a.images_ds, a.masks, a.nactive = images_ds.data_ptr(), masks.data_ptr(), nactive.data_ptr()
    L.check(L.lib().dupl_refine_prologue(C.byref(a), L.stream_ptr(dev)), "dupl_refine_prologue")
    return images_ds, masks, nactive, cls


def refine_epilogue(masks, cls, img_box, H, W, ignore_index, want_parts=False):
    b, P = masks.shape[0], masks.shape[1]
    K = P // 2 - 1
    dev = masks.device
    box = box_to_device(img_box, dev)
    a = L.RefineEpilogueArgs()
    a.masks, a.cls_label, a.img_box = masks.data_ptr(), cls.data_ptr(), box.data_ptr()
    a.b, a.K, a.H, a.W = b, K, H, W
    a.ignore_index = float(ignore_index)
    label = torch.empty(b, H, W, dtype=torch.float32, device=dev)
    a.label = label.data_ptr()
    lh = ll = None
    if want_parts:
        lh, ll = torch.empty_like(label), torch.empty_like(label)
        a.label_h, a.label_l = lh.data_ptr(), ll.data_ptr()
    L.check(L.lib().dupl_refine_epilogue(C.byref(a), L.stream_ptr(dev)), "dupl_refine_epilogue")
    return (label, lh, ll) if want_parts else label


# ------------------------------------------------------------------ DenseCRF
class CrfWorkspace:
    """Lattice workspace re-used across images of the same size (one allocation per (H, W))."""

    def __init__(self):
        self.key, self.buf, self.values, self.meta = None, None, None, None

    def get(self, H, W, device):
        if self.key != (H, W, device):
            n = C.c_size_t(0)
            L.check(L.lib().dupl_crf_workspace_bytes(W, H, C.byref(n)), "dupl_crf_workspace_bytes")
            self.buf = torch.empty(n.value, dtype=torch.uint8, device=device)
            self.meta = torch.zeros(4, dtype=torch.int32, device=device)
            self.key = (H, W, device)
        return self.buf, self.meta

    def get_values(self, nbytes, device):
        if self.values is None or self.values.numel() < nbytes or self.values.device != device:
            self.values = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self.values


def crf_inference(image_u8, unary_or_probs, iters, pos_w, pos_xy_std, bi_w, bi_xy_std, bi_rgb_std, input_is_energy=False,
                  ws=None):
    """image_u8: uint8 [H,W,3] cuda; unary_or_probs: fp32 [C,H,W] cuda -> (Q fp32 [C,H,W] cuda, (M_gauss, M_bilateral))."""
    L.require_cuda(image_u8, unary_or_probs)
    ws = ws or CrfWorkspace()
    x = L.f32c(unary_or_probs)
    img = image_u8.contiguous()
    Cn, H, W = x.shape
    if img.dtype != torch.uint8 or tuple(img.shape) != (H, W, 3):
        raise ValueError("image must be uint8 [H, W, 3] matching the probability map")
    dev = x.device
    buf, meta = ws.get(H, W, dev)
    out = torch.empty_like(x)
    a = L.CrfArgs()
    a.W, a.H, a.C = W, H, Cn
    a.pos_w, a.pos_xy_std, a.bi_w, a.bi_xy_std, a.bi_rgb_std = pos_w, pos_xy_std, bi_w, bi_xy_std, bi_rgb_std
    a.iters, a.input_is_energy = iters, 1 if input_is_energy else 0
    a.image, a.unary_or_probs, a.out = img.data_ptr(), x.data_ptr(), out.data_ptr()
    a.workspace, a.workspace_bytes, a.meta = buf.data_ptr(), buf.numel(), meta.data_ptr()
    st = L.stream_ptr(dev)
    L.check(L.lib().dupl_crf_build(C.byref(a), st), "dupl_crf_build")
    m = meta.tolist()  # the one host sync of the CRF: vertex counts size the value arrays
    if m[2]:
        raise RuntimeError("DenseCRF: a lattice coordinate left the packed key range (features too large for this build)")
    nv = C.c_size_t(0)
    L.check(L.lib().dupl_crf_values_bytes(W, H, Cn, m[0], m[1], C.byref(nv)), "dupl_crf_values_bytes")
    vals = ws.get_values(nv.value, dev)
    a.values, a.values_bytes = vals.data_ptr(), vals.numel()
    L.check(L.lib().dupl_crf_infer(C.byref(a), m[0], m[1], st), "dupl_crf_infer")
    _LAST_CRF_SIZES[:] = [m[0], m[1]]
    return out, (m[0], m[1])


_LAST_CRF_SIZES = [0, 0]


def last_crf_lattice_sizes():
    """(vertices of the d=2 Gaussian lattice, vertices of the d=5 bilateral lattice) of the most recent crf_inference."""
    return tuple(_LAST_CRF_SIZES)
