"""CUDA-vs-oracle measurements at BASELINE.json's sizes (448x448; N = 197 / 442 / 785 / 1765 tokens; K = 20 / 80).

Each function RETURNS the error figures (so that tests/test_gpu_baseline_sizes.py can assert on them and
tools/precision_table.py can tabulate them for alternative tensor-core operand schemes); nothing here asserts.
The CPU oracle results are cached under $DUPL_ORACLE_CACHE (default /tmp/dupl_oracle_cache): the phase-B loop of one
448x448 image costs ~25 s on 8 cores, and the precision table re-uses it for every scheme.
"""
import hashlib
import os

import torch

from helpers import init_state_dict, mscam_err, rel_err, synth_boxes, synth_cls_labels, synth_images

CACHE = os.environ.get("DUPL_ORACLE_CACHE", "/tmp/dupl_oracle_cache")
SCALES = (1.0, 0.5, 1.5)


def _cached(key, fn):
    os.makedirs(CACHE, exist_ok=True)
    path = os.path.join(CACHE, hashlib.sha1(key.encode()).hexdigest()[:16] + ".pt")
    if os.path.exists(path):
        return torch.load(path)
    val = fn()
    torch.save(val, path)
    return val


def nrel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build_model(num_classes=21, aux_layer=-3, train=False):
    from dupl_b200.model.model_dupl import siamese_network
    P = init_state_dict(num_classes)
    m = siamese_network("deit_base_patch16_224", num_classes=num_classes, pretrained=False, aux_layer=aux_layer)
    m.load_state_dict(P, strict=True)
    m = m.cuda()
    return (m.train() if train else m.eval()), P


# ------------------------------------------------------------------------------------------------ attention kernels
def _attention_ref_fp64(qkv64, dO64, B, N, heads=12, scale=0.125):
    """softmax(Q K^T * scale) V per (image, head) in fp64 on the GPU with torch autograd (vit.py:120-135)."""
    D = heads * 64
    x = qkv64.clone().requires_grad_(True)
    blk = x.reshape(B, N, 3, heads, 64).permute(2, 0, 3, 1, 4)
    att = torch.softmax(blk[0] @ blk[1].transpose(-1, -2) * scale, -1)
    out = (att @ blk[2]).permute(0, 2, 1, 3).reshape(B * N, D)
    if dO64 is None:
        return out.detach(), None
    out.backward(dO64)
    return out.detach(), x.grad


def attention_errors(B, gh, gw, seed=0, backward=True):
    """dupl_attention_fwd / dupl_attention_bwd against fp64 autograd on the SAME (split-rounded) operands.
    -> dict(fwd=max-rel error of O, dq / dk / dv = max-rel error of each gradient third)."""
    from dupl_b200 import ops
    segs, M, _ = ops.make_segments([(B, gh, gw)])
    N = segs[0].tokens
    g = torch.Generator().manual_seed(100 + seed)
    qkv = torch.randn(M, 2304, generator=g).cuda()
    dO = torch.randn(M, 768, generator=g).cuda()
    qh, ql = ops.split_bf16(qkv)
    oh = torch.zeros(M, 768, dtype=torch.bfloat16, device="cuda")
    ol = torch.zeros_like(oh)
    lse = torch.empty(M, 12, dtype=torch.float32, device="cuda")
    ops.attention_fwd(qh, ql, oh, ol, segs, 12, 0.125, lse=lse)
    out = oh.float() + ol.float()
    q64 = qh.double() + ql.double()                       # what the kernels actually see
    res = {"tokens": N}
    if not backward:
        ref, _ = _attention_ref_fp64(q64, None, B, N)
        res["fwd"] = rel_err(out, ref)
        return res
    dh, dl = ops.split_bf16(dO)
    d64 = dh.double() + dl.double()
    ref, gref = _attention_ref_fp64(q64, d64, B, N)
    res["fwd"] = rel_err(out, ref)
    dqkv = ops.attention_bwd((qh, ql), (oh, ol), (dh, dl), lse, B, N, 12, 0.125)
    got = dqkv.reshape(M, 3, 768)
    want = gref.reshape(M, 3, 768)
    for i, n in enumerate(("dq", "dk", "dv")):
        res[n] = rel_err(got[:, i], want[:, i])
    return res


# ------------------------------------------------------------------------------------------------ MS-CAM + refine at 448^2
def _oracle_mscam(num_classes, aux_layer, b, S, seed, branch):
    from oracle import dupl_oracle as O

    def run():
        P = init_state_dict(num_classes)
        x = synth_images(b, S, S, seed=seed)
        with torch.no_grad():
            return O.multi_scale_cam(P, branch, x, SCALES, aux_layer=aux_layer, return_sums=True)
    return _cached(f"mscam-{num_classes}-{aux_layer}-{b}-{S}-{seed}-{branch}", run)


def mscam_errors(model, num_classes=21, aux_layer=-3, b=1, S=448, seed=3, pair=True):
    """multi_scale_cam2_siamese / multi_scale_cam2_pair at S x S (N = 785 / 197 / 1765 in one grouped pass) vs the oracle.
    -> dict with, per student, the error in un-normalised CAM units (helpers.mscam_err) and the raw max error on the
    well-conditioned planes; plus the refine-label mismatch fraction end to end and stage-isolated."""
    from dupl_b200.model.PAR import PAR
    from dupl_b200.utils import cam_helper
    from oracle import dupl_oracle as O
    K = num_classes - 1
    x = synth_images(b, S, S, seed=seed)
    cls = synth_cls_labels(b, K, seed=seed)
    box = synth_boxes(b, S, S, seed=seed)
    with torch.no_grad():
        if pair:
            got = cam_helper.multi_scale_cam2_pair(model, x.cuda(), SCALES)
        else:
            got = [cam_helper.multi_scale_cam2_siamese(model, x.cuda(), SCALES, branch=br) for br in (1, 2)]
    res = {}
    par = PAR(num_iter=10, dilations=[1, 2, 4, 8, 12, 24]).cuda()
    den = O.denormalize_img2(x)
    for br in (1, 2):
        cam, aux = got[br - 1]
        ocam, oaux, osum, oaux_sum = _oracle_mscam(num_classes, aux_layer, b, S, seed, br)
        well = O.mscam_condition(osum) < 20
        res[f"cam{br}"] = mscam_err(cam, ocam, osum)
        res[f"aux{br}"] = mscam_err(aux, oaux, oaux_sum)
        res[f"cam{br}_raw_well"] = ((cam.cpu() - ocam).abs() * well).max().item()
        # refine labels (cam_helper.py:338-440): end to end (GPU CAM -> GPU labels vs oracle CAM -> oracle labels) and
        # stage-isolated (the SAME GPU CAM through both refine implementations)
        clsb = cls[:, :, None, None]
        lab = cam_helper.refine_cams_with_bkg_v2(par, den.cuda(), cam * clsb.cuda(), cls.cuda(), high_thre=0.65, low_thre=0.25,
                                                 ignore_index=255, img_box=box).cpu()
        olab = _cached(f"refine-{num_classes}-{aux_layer}-{b}-{S}-{seed}-{br}",
                       lambda: O.refine_cams(den, ocam * clsb, cls, 0.65, 0.25, 255, box))
        res[f"label{br}_mismatch"] = (lab != olab).float().mean().item()
        if br == 1:
            iso = O.refine_cams(den, cam.cpu() * clsb, cls, 0.65, 0.25, 255, box)
            res["label1_mismatch_isolated"] = (lab != iso).float().mean().item()
    return res


# ------------------------------------------------------------------------------------------------ training step at 448^2
def _oracle_train(num_classes, b, S, seed, n_iter, coco):
    from oracle import dupl_oracle as O

    def run():
        P = init_state_dict(num_classes)
        x = synth_images(b, S, S, seed=seed)
        cls = synth_cls_labels(b, num_classes - 1, seed=seed + 1)
        if coco:
            cls = cls.to(torch.uint8)
        box = synth_boxes(b, S, S, seed=seed + 2)
        Pg = {k: v.clone().requires_grad_("pos_embed" not in k and ".head." not in k) for k, v in P.items()}
        cfg = O.COCO_CFG if coco else O.VOC_CFG
        from dupl_b200.train_step import Args, CocoArgs
        target = list((CocoArgs if coco else Args).high_thres_target)
        loss, parts, labels = O.train_losses(Pg, x, cls, box, n_iter, cfg, thres_target=target)
        loss.backward()
        grads = {k: v.grad.clone() for k, v in Pg.items() if v.grad is not None}
        with torch.no_grad():
            fmaps = [O.network_forward(P, br, x, cfg["aux_layer"])[2] for br in (1, 2)]
        return dict(loss=loss.detach(), parts={k: torch.as_tensor(v).detach() for k, v in parts.items()},
                    labels=labels, grads=grads, fmaps=fmaps)
    return _cached(f"train-{num_classes}-{b}-{S}-{seed}-{n_iter}-{coco}", run)


def train_step_errors(num_classes=21, b=1, S=448, seed=40, n_iter=3000, coco=False):
    """TrainStep.losses + backward (phase B of train_final_voc.py:260-456 / train_final_coco.py) at S x S against the
    oracle's CPU autograd: every loss part (relative), label mismatch, and the norm-relative error of ALL parameter
    gradients of both students (2 x 154 tensors)."""
    from dupl_b200.train_step import Args, CocoArgs, TrainStep
    want = _oracle_train(num_classes, b, S, seed, n_iter, coco)
    m, _ = build_model(num_classes, aux_layer=9 if coco else -3, train=True)
    x = synth_images(b, S, S, seed=seed)
    cls = synth_cls_labels(b, num_classes - 1, seed=seed + 1)
    if coco:
        cls = cls.to(torch.uint8)
    box = synth_boxes(b, S, S, seed=seed + 2)
    step = TrainStep(m, None, args=CocoArgs if coco else Args)
    loss, parts, labels = step.losses(x.cuda(), cls.cuda(), box, n_iter)
    loss.backward()
    res = {"loss": abs(loss.item() - want["loss"].item()) / max(1.0, abs(want["loss"].item()))}
    for k, v in want["parts"].items():
        res[k] = abs(float(parts[k]) - float(v)) / max(1.0, abs(float(v)))
    res["label_mismatch"] = max((a.cpu() != w).float().mean().item() for a, w in zip(labels, want["labels"]))
    worst, checked, errs = (0.0, ""), 0, {}
    for name, p in m.named_parameters():
        if ".head." in name or "pos_embed" in name:
            continue
        ref = want["grads"][name]
        e = nrel(p.grad, ref)
        errs[name] = e
        worst = max(worst, (e, name))
        checked += 1
    res["grads_checked"] = checked
    res["grad_worst"], res["grad_worst_name"] = worst
    vals = sorted(errs.values())
    res["grad_median"] = vals[len(vals) // 2]
    res["grad_over_1e-3"] = sum(1 for v in vals if v >= 1e-3)
    # Global max pooling (model_dupl.py:88-95) routes the classification gradient to the arg-max token of every channel: where
    # two tokens tie within the forward tolerance the route is ill-defined and one flipped channel of 768 moves the gradient
    # norm by ~sqrt(2/768) = 5 %.  Report the flips and how close the tie was in the oracle's own feature map.
    with torch.no_grad():
        out = m(x.cuda())
    flips, worst_margin = 0, 0.0
    for br in (1, 2):
        fo = want["fmaps"][br - 1].flatten(2)                       # [b, 768, hw]
        fg = out[f"branch{br}"][2].detach().cpu().flatten(2)
        ao, ag = fo.argmax(2), fg.argmax(2)
        diff = ao != ag
        flips += int(diff.sum())
        if diff.any():
            top = fo.gather(2, ao[..., None])[..., 0]
            alt = fo.gather(2, ag[..., None])[..., 0]
            worst_margin = max(worst_margin, float(((top - alt)[diff] / fo.abs().max()).max()))
        res[f"grad_worst_branch{br}"] = max(e for n, e in errs.items() if n.startswith(f"branch{br}."))
    res["gmp_argmax_flips"], res["gmp_flip_margin_rel"] = flips, worst_margin
    return res
