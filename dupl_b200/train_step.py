"""The dual-student training step as one callable (driver "M2" of SURVEY §8(b)): a restatement of the
loop body of train_final_voc.py:186-472 for the phase cam_iters <= n_iter < gmm_iters ("phase B": CAM +
PAR pseudo-labels + decoder + all losses, no GMM filter / consistency term), built from the drop-in
modules of this package.  The script-side glue the reference executes with stock torch ops between the
calls into the model / helpers (classification loss, F.interpolate of logits and CAMs, cosine
discrepancy loss, loss weighting, AdamW) stays stock torch here as well; everything the reference
reaches through model(...), cam_helper, PAR and model.losses runs in libdupl.so.

Differences from the script, none of which changes a number:
  * the per-image high threshold is computed on the device (masked max) instead of through
    torch.nonzero + a Python loop (train_final_voc.py:268-275) — no host sync;
  * the up-sampling of the seg logits and the seg loss are one kernel pair (model.losses.get_seg_loss_upsampled);
  * no logging / validation / checkpointing.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .model.losses import get_masked_ptc_loss, get_seg_loss_upsampled
from .model.PAR import PAR
from .pipeline import CamParStep, denormalize_img2
from .utils import cam_helper

VOC_HIGH_THRES_TARGET = (0.70, 0.70, 0.70, 0.70, 0.55, 0.55, 0.55, 0.55, 0.70, 0.55,
                         0.55, 0.55, 0.55, 0.55, 0.55, 0.55, 0.55, 0.55, 0.70, 0.55)  # train_final_voc.py:163-166


class Args:
    """Defaults of train_final_voc.py:33-90 that the step uses."""
    cam_scales = (1.0, 0.5, 1.5)
    bkg_thre = 0.5
    high_thre = 0.7
    low_thre = 0.25
    ignore_index = 255
    w_ptc = 0.2
    w_seg = 0.2
    cam_iters = 2000
    gmm_iters = 8000
    max_iters = 20000
    lr = 6e-5
    wt_decay = 1e-2
    betas = (0.9, 0.999)
    warmup_iters = 1500
    warmup_lr = 1e-6
    power = 0.9


def cosine_descent(max_thres, min_thres, step, num_steps):
    """utils/train_helper.py:340-349"""
    if step < 0:
        return max_thres
    if step >= num_steps:
        return min_thres
    f = step / (num_steps - 1)
    return max_thres + (min_thres - max_thres) * (1 - np.cos(np.pi * f)) / 2


class PolyWarmupAdamW(torch.optim.AdamW):
    """utils/optimizer.py:38-68: AdamW whose step() first sets the warm-up / polynomial learning rate."""

    def __init__(self, params, lr, weight_decay, betas, warmup_iter=None, max_iter=None, warmup_ratio=None, power=None,
                 fused=None):
        super().__init__(params, lr=lr, betas=betas, weight_decay=weight_decay, eps=1e-8, fused=fused)
        self.global_step = 0
        self.warmup_iter, self.warmup_ratio, self.max_iter, self.power = warmup_iter, warmup_ratio, max_iter, power
        self._init_lr = [g["lr"] for g in self.param_groups]

    def step(self, closure=None):
        if self.global_step < self.warmup_iter:
            m = 1 - (1 - self.global_step / self.warmup_iter) * (1 - self.warmup_ratio)
        elif self.global_step < self.max_iter:
            m = (1 - self.global_step / self.max_iter) ** self.power
        else:
            m = None
        if m is not None:
            for g, lr0 in zip(self.param_groups, self._init_lr):
                g["lr"] = lr0 * m
        super().step(closure)
        self.global_step += 1


def make_optimizer(model, args=Args):
    """utils/train_helper.py:21-87 (get_optimizer): 4 groups, heads and decoders at 10x learning rate."""
    g = model.get_param_groups()
    on_gpu = all(p.is_cuda for grp in g for p in grp)  # torch's single-kernel-per-chunk AdamW (same update rule)
    return PolyWarmupAdamW(
        params=[{"params": g[0], "lr": args.lr, "weight_decay": args.wt_decay},
                {"params": g[1], "lr": args.lr, "weight_decay": args.wt_decay},
                {"params": g[2], "lr": args.lr * 10, "weight_decay": args.wt_decay},
                {"params": g[3], "lr": args.lr * 10, "weight_decay": args.wt_decay}],
        lr=args.lr, weight_decay=args.wt_decay, betas=args.betas, warmup_iter=args.warmup_iters, max_iter=args.max_iters,
        warmup_ratio=args.warmup_lr, power=args.power, fused=True if on_gpu else None)


class PhaseBStep:
    def __init__(self, model, optim=None, args=Args, device=None, graph=True, reuse_forward=True):
        self.model = model          # siamese_network or DistributedDataParallel(siamese_network)
        self.optim = optim
        self.args = args
        dev = device or next(model.parameters()).device
        self.par = PAR(num_iter=10, dilations=[1, 2, 4, 8, 12, 24]).to(dev)
        # the no-grad half of the step (MS-CAM of both students + PAR refinement) replayed as one CUDA graph
        # reuse_forward: `model(inputs)` of train_final_voc.py:287 recomputes what the MS-CAM pass has just computed for
        # the un-flipped scale-1.0 images (identical arithmetic: no dropout, no batch statistics); keep those activations
        # and start the training pass at the final LayerNorm.
        self.reuse_forward = reuse_forward and 1.0 in tuple(args.cam_scales)
        self.pseudo = CamParStep(model, args.cam_scales, low_thre=args.low_thre, ignore_index=args.ignore_index, graph=graph,
                                 keep_activations=self.reuse_forward)
        self.pseudo.par = self.par
        self.thres_start = torch.ones(20, device=dev) * args.high_thre
        self.thres_target = torch.tensor(VOC_HIGH_THRES_TARGET, device=dev)

    def losses(self, inputs, cls_label, img_box, n_iter):
        a = self.args
        model = self.model
        b, _, h, w = inputs.shape
        # per-image high threshold = max over the present classes of the cosine-annealed class thresholds
        thres = cosine_descent(self.thres_start, self.thres_target, n_iter - a.cam_iters, a.max_iters - a.cam_iters)
        high_thres = torch.where(cls_label > 0, thres[None, :], thres.new_full((), -math.inf)).amax(1)

        # multi_scale_cam2_siamese x2 and refine_cams_with_dynamic_thres x2 (train_final_voc.py:279-284, 330-343).
        # cams * cls_label_rep of the script is the identity on the channels the refine kernels read (one-hot labels).
        label_1, label_2, (cams_1, cams_aux_1), (cams_2, cams_aux_2) = self.pseudo(inputs, cls_label, img_box, high_thres)
        net = model.module if hasattr(model, "module") else model
        for n in (net.branch1, net.branch2):
            n._use_kept = self.reuse_forward
        res = model(inputs)
        for n in (net.branch1, net.branch2):
            n._use_kept = False
        cls_1, segs_1, fmap_1, cls_aux_1 = res["branch1"]
        cls_2, segs_2, fmap_2, cls_aux_2 = res["branch2"]

        cls_loss = (F.multilabel_soft_margin_loss(cls_1, cls_label) + F.multilabel_soft_margin_loss(cls_aux_1, cls_label) +
                    F.multilabel_soft_margin_loss(cls_2, cls_label) + F.multilabel_soft_margin_loss(cls_aux_2, cls_label))

        ptc_loss = 0
        for cams_aux, fmap in ((cams_aux_1, fmap_1), (cams_aux_2, fmap_2)):
            resized = F.interpolate(cams_aux, size=fmap.shape[2:], mode="bilinear", align_corners=False)
            _, pseudo = cam_helper.cam_to_label_dynamic_cls(resized.detach(), cls_label=cls_label, img_box=img_box, ignore_mid=True,
                                                            bkg_thre=a.bkg_thre, high_thre=high_thres, low_thre=a.low_thre,
                                                            ignore_index=a.ignore_index)
            ptc_loss = ptc_loss + get_masked_ptc_loss(fmap, cam_helper.label_to_aff_mask(pseudo))

        # F.interpolate(segs, size=label.shape[1:]) + get_seg_loss (train_final_voc.py:345-352), fused
        seg_loss = get_seg_loss_upsampled(segs_1, label_2) + get_seg_loss_upsampled(segs_2, label_1)

        f1 = fmap_1.view(fmap_1.shape[0], fmap_1.shape[1], -1)
        f2 = fmap_2.view(fmap_2.shape[0], fmap_2.shape[1], -1)
        cos = torch.nn.CosineSimilarity(dim=-1, eps=1e-6)
        sim_loss = (1 + cos(f1.detach(), f2).mean()) + (1 + cos(f2.detach(), f1).mean())

        loss = 1.0 * cls_loss + a.w_ptc * ptc_loss + a.w_seg * seg_loss + 0.1 * sim_loss
        return loss, dict(cls_loss=cls_loss, ptc_loss=ptc_loss, seg_loss=seg_loss, sim_loss=sim_loss), (label_1, label_2)

    def __call__(self, inputs, cls_label, img_box, n_iter):
        loss, parts, _ = self.losses(inputs, cls_label, img_box, n_iter)
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        return loss.detach(), {k: v.detach() for k, v in parts.items()}
