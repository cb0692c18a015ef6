"""The dual-student training step as one callable (driver "M2" of SURVEY §8(b)): a restatement of the
loop body of train_final_voc.py:186-472 / train_final_coco.py:182-464 for all three phases (A: n_iter < cam_iters,
CAM + cls/PTC losses; B: + PAR pseudo-labels and seg loss; C: n_iter >= gmm_iters, + the strongly augmented view, the GMM
noise filter on the GPU and the consistency term), built from the drop-in modules of this package.  The script-side glue the reference executes with stock torch ops between the
calls into the model / helpers is fused where it costs launches (classification loss x4, cosine discrepancy loss,
up-sampling + CE, AdamW) and stock torch for the rest (F.interpolate of the CAMs, loss weighting); everything the reference
reaches through model(...), cam_helper, PAR and model.losses runs in libdupl.so.

Differences from the script:
  * the per-image high threshold is computed on the device (masked max) instead of through
    torch.nonzero + a Python loop (train_final_voc.py:268-275) — no host sync, same numbers;
  * the up-sampling of the seg logits and the seg loss are one kernel pair (model.losses.get_seg_loss_upsampled), same numbers;
  * phase C: the GMM noise filter (train_final_voc.py:358-394, scikit-learn on the host in the reference) runs on the device
    with deterministic Lloyd iterations in place of scikit-learn's k-means++ seeding.  Pinned against scikit-learn 1.9 on
    unimodal, weakly bimodal and well-separated CE histograms (oracle/gmm_ref.py, tests/test_gmm_oracle.py): the `|d mean| > 1`
    gate decides identically and the masks differ by <= 2 pixels of 36 864 — the band scikit-learn shows against itself with
    another random_state.  It is NOT bit-pinned to the reference's scikit-learn 1.0.2;
  * phase C without `inputs_aug`: the strong augmentation (utils/imutils.augment_data_strong, PIL on the host in the
    reference) runs on the device, bit-exact with Pillow for the same operation draw (utils/imutils.py of this package);
  * the refined pseudo-labels handed back by `losses()` are the step's own buffers (graph-static under capture=True and
    overwritten — also by the GMM filter, in place — on the next call): clone them to keep them;
  * no logging / validation / checkpointing.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .gmm import gmm_noise_filter
from .model.losses import (ce_sum_upsampled, discrepancy_loss, get_masked_ptc_loss, get_seg_loss_upsampled,
                           multilabel_soft_margin_sum)
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
    num_classes = 21
    gmm_valid_thre = 1.0
    gamma = 0.95
    high_thres_target = VOC_HIGH_THRES_TARGET
    thres_anneal_from = 2000       # cosine_descent(.., n_iter - cam_iters, max_iters - cam_iters), train_final_voc.py:263-265
    aux_refine_until = None        # COCO only: refine_cams_with_bkg_v2 on cams_aux while n_iter <= 12000
    ptc_in_phase_a = True          # COCO sets ptc_loss = 1 (weight 0) before cam_iters, train_final_coco.py:216

    @staticmethod
    def loss_weights(n_iter):
        """train_final_voc.py:451-456"""
        if n_iter <= Args.cam_iters:
            return dict(cls=1.0, ptc=Args.w_ptc, seg=0.0, sim=0.1, reg=0.0)
        if n_iter <= Args.gmm_iters:
            return dict(cls=1.0, ptc=Args.w_ptc, seg=Args.w_seg, sim=0.1, reg=0.0)
        return dict(cls=1.0, ptc=Args.w_ptc, seg=Args.w_seg, sim=0.1, reg=0.05)

    @staticmethod
    def coco():
        return CocoArgs


class CocoArgs(Args):
    """Defaults of train_final_coco.py:33-90 and the constants inlined in its loop body."""
    num_classes = 81
    bkg_thre = 0.45
    high_thre = 0.65
    cam_iters = 8000
    gmm_iters = 32000
    max_iters = 80000
    high_thres_target = (0.55,) * 80   # train_final_coco.py:161-162
    thres_anneal_from = 12000          # train_final_coco.py:240-242
    aux_refine_until = 12000           # train_final_coco.py:312-333
    ptc_in_phase_a = False

    @staticmethod
    def loss_weights(n_iter):
        """train_final_coco.py:441-448"""
        if n_iter <= 8000:
            return dict(cls=1.0, ptc=0.0, seg=0.0, sim=0.0, reg=0.0)
        if n_iter <= 12000:
            return dict(cls=1.0, ptc=0.0, seg=0.2, sim=0.05, reg=0.0)
        return dict(cls=1.0, ptc=0.2, seg=0.2, sim=0.05, reg=0.05)


def cosine_descent(max_thres, min_thres, step, num_steps):
    """utils/train_helper.py:340-349"""
    if step < 0:
        return max_thres
    if step >= num_steps:
        return min_thres
    f = step / (num_steps - 1)
    return max_thres + (min_thres - max_thres) * (1 - np.cos(np.pi * f)) / 2


class PolyWarmupAdamW(torch.optim.AdamW):
    """utils/optimizer.py:38-68: AdamW whose step() first sets the warm-up / polynomial learning rate.
    capturable=True keeps the learning rates and step counters in device tensors so that step() can be part of a CUDA
    graph (TrainStep(capture=True)); the schedule is then advanced by advance_schedule() before each replay."""

    def __init__(self, params, lr, weight_decay, betas, warmup_iter=None, max_iter=None, warmup_ratio=None, power=None,
                 fused=None, capturable=False):
        params = list(params)
        self._init_lr = [g["lr"] for g in params]
        if capturable:
            dev = params[0]["params"][0].device
            for g in params:
                g["lr"] = torch.tensor(g["lr"], dtype=torch.float32, device=dev)
            lr = torch.tensor(lr, dtype=torch.float32, device=dev)
        super().__init__(params, lr=lr, betas=betas, weight_decay=weight_decay, eps=1e-8, fused=fused, capturable=capturable)
        self.global_step = 0
        self.warmup_iter, self.warmup_ratio, self.max_iter, self.power = warmup_iter, warmup_ratio, max_iter, power

    def advance_schedule(self):
        """Sets this step's learning rates (utils/optimizer.py:52-66) and counts the step."""
        if self.global_step < self.warmup_iter:
            m = 1 - (1 - self.global_step / self.warmup_iter) * (1 - self.warmup_ratio)
        elif self.global_step < self.max_iter:
            m = (1 - self.global_step / self.max_iter) ** self.power
        else:
            m = None
        if m is not None:
            for g, lr0 in zip(self.param_groups, self._init_lr):
                if torch.is_tensor(g["lr"]):
                    g["lr"].fill_(lr0 * m)
                else:
                    g["lr"] = lr0 * m
        self.global_step += 1

    def step(self, closure=None):
        self.advance_schedule()
        super().step(closure)

    def step_captured(self, closure=None):
        """The parameter update alone (what a CUDA graph captures); the caller advances the schedule."""
        super().step(closure)


def make_optimizer(model, args=Args, capturable=False, fused=None):
    """utils/train_helper.py:21-87 (get_optimizer): 4 groups, heads and decoders at 10x learning rate.
    fused (default: on for capturable on CUDA; DUPL_FUSED_ADAMW=0 switches it off): optim.FusedPolyWarmupAdamW, one launch
    for all tensors of both students that also rewrites the split-bf16 planes of the GEMM weights."""
    import os
    g = model.get_param_groups()
    if fused is None:
        fused = capturable and os.environ.get("DUPL_FUSED_ADAMW", "1") != "0"
    if fused and all(p.is_cuda for grp in g for p in grp):
        from .optim import FusedPolyWarmupAdamW
        lookup = {}
        for net in (model.branch1, model.branch2):
            pl = net.planes()
            pdict = dict(net.encoder.named_parameters())
            for name in pl.GEMM_WEIGHTS:
                lookup[id(pdict[name])] = pl.plane(name)
        return FusedPolyWarmupAdamW(
            params=[{"params": g[0], "lr": args.lr, "weight_decay": args.wt_decay},
                    {"params": g[1], "lr": args.lr, "weight_decay": args.wt_decay},
                    {"params": g[2], "lr": args.lr * 10, "weight_decay": args.wt_decay},
                    {"params": g[3], "lr": args.lr * 10, "weight_decay": args.wt_decay}],
            lr=args.lr, weight_decay=args.wt_decay, betas=args.betas, warmup_iter=args.warmup_iters, max_iter=args.max_iters,
            warmup_ratio=args.warmup_lr, power=args.power, plane_lookup=lambda p: lookup.get(id(p)))
    on_gpu = all(p.is_cuda for grp in g for p in grp)  # torch's single-kernel-per-chunk AdamW (same update rule)
    return PolyWarmupAdamW(
        params=[{"params": g[0], "lr": args.lr, "weight_decay": args.wt_decay},
                {"params": g[1], "lr": args.lr, "weight_decay": args.wt_decay},
                {"params": g[2], "lr": args.lr * 10, "weight_decay": args.wt_decay},
                {"params": g[3], "lr": args.lr * 10, "weight_decay": args.wt_decay}],
        lr=args.lr, weight_decay=args.wt_decay, betas=args.betas, warmup_iter=args.warmup_iters, max_iter=args.max_iters,
        warmup_ratio=args.warmup_lr, power=args.power, fused=True if on_gpu else None, capturable=capturable and on_gpu)


class TrainStep:
    """One iteration of the training loop for any n_iter (phases A / B / C of train_final_voc.py:186-472 and
    train_final_coco.py:182-464).  `PhaseBStep` is the historical name of the same class."""

    def __init__(self, model, optim=None, args=Args, device=None, graph=True, reuse_forward=True, capture=False):
        """capture=True: the WHOLE iteration (MS-CAM, PAR, forward heads, losses, backward, AdamW) is captured once per phase
        as one CUDA graph and replayed; ~1500 kernel launches per step otherwise keep the host as busy as the GPU.  Needs a
        capturable optimizer (make_optimizer(..., capturable=True)) and the bare model: with several ranks the gradients are
        averaged by one NCCL all-reduce inside the graph instead of by DistributedDataParallel's hooks."""
        self.model = model          # siamese_network or DistributedDataParallel(siamese_network)
        self.optim = optim
        self.args = args
        self.capture = capture
        self._graphs = {}
        self._aug_ops = None        # captured phase C: device tensor [5][b] of RandAugment operation indices, refilled per step
        self._last_graph_key = None
        self._sched = None          # static device scalars of the threshold schedule while capturing / replaying
        dev = device or next(model.parameters()).device
        self.par = PAR(num_iter=10, dilations=[1, 2, 4, 8, 12, 24]).to(dev)
        # reuse_forward: `model(inputs)` of train_final_voc.py:287 recomputes what the MS-CAM pass has just computed for
        # the un-flipped scale-1.0 images (identical arithmetic: no dropout, no batch statistics); keep those activations
        # and start the training pass at the final LayerNorm.
        self.reuse_forward = reuse_forward and 1.0 in tuple(args.cam_scales)
        # the no-grad half of the step (MS-CAM of both students + PAR refinement) replayed as one CUDA graph
        self.pseudo = CamParStep(model, args.cam_scales, low_thre=args.low_thre, ignore_index=args.ignore_index, graph=graph,
                                 keep_activations=self.reuse_forward)
        self.pseudo.par = self.par
        # phase A and the COCO 8000 < n_iter <= 12000 window need other label sources: separate (graph-replayed) variants
        self._cam_only = CamParStep(model, args.cam_scales, low_thre=args.low_thre, ignore_index=args.ignore_index, graph=graph,
                                    keep_activations=self.reuse_forward, refine=None)
        self._cam_aux = CamParStep(model, args.cam_scales, low_thre=args.low_thre, ignore_index=args.ignore_index, graph=graph,
                                   keep_activations=self.reuse_forward, refine="aux_scalar", scalar_high_thre=args.high_thre)
        for st in (self._cam_only, self._cam_aux):
            st.par = self.par
        if capture:                 # one graph for everything: the CAM half is captured inline, not replayed as a sub-graph
            for st in (self.pseudo, self._cam_only, self._cam_aux):
                st.graph = False
        # capture=True: the students' gradients live in flat arenas (grad_arena.GradArena) and, with several ranks, are averaged
        # chunk by chunk WHILE the backward pass runs (what DistributedDataParallel's reducer does for the reference,
        # train_final_voc.py:155,470-471) by a process group of its own whose NCCL kernels are kept to a few CTAs
        self._arenas = None
        self._comm_sms = 0
        if capture:
            import os
            from . import train
            net = model.module if hasattr(model, "module") else model
            group = None
            world = self._world()
            # Measured on B200 (profiles/r02_summary.md §4): the gradients are averaged chunk by chunk while the backward runs,
            # NCCL kept to 16 CTAs and the persistent GEMM grid leaving 16 SMs alone — without that reservation the
            # collective's CTAs delay statically assigned GEMM tiles (55.8 vs 50.6 ms at 8 ranks).  At the end of round 2:
            # 2 ranks 47.85 ms chunked + overlapped vs 48.2-48.9 with ONE all-reduce after the backward (8 / 8 and 24 / 24:
            # 48.7 / 48.4); 8 ranks 50.6 vs 51.3 (mid-round build).  DUPL_GRAD_OVERLAP=0 selects the single all-reduce.
            overlap = os.environ.get("DUPL_GRAD_OVERLAP", "1" if world >= 2 else "0") != "0"
            chunk = int(os.environ.get("DUPL_GRAD_CHUNK_ELEMS", str(6 << 20 if overlap else 1 << 30)))
            if world > 1 and overlap:
                import torch.distributed as dist
                ctas = int(os.environ.get("DUPL_NCCL_MAX_CTAS", "16"))
                if ctas > 0 and dist.get_backend() == "nccl":
                    opts = dist.ProcessGroupNCCL.Options()
                    opts.config.max_ctas = ctas
                    opts.config.min_ctas = min(ctas, 4)
                    group = dist.new_group(backend="nccl", pg_options=opts)
                    # SMs the persistent GEMM grid leaves alone while gradient chunks are in flight
                    self._comm_sms = int(os.environ.get("DUPL_COMM_SMS", "16"))
            self._arenas = [train.make_grad_arena(n, chunk_elems=chunk, group=group) for n in (net.branch1, net.branch2)]
        K = args.num_classes - 1
        self.thres_start = torch.ones(K, device=dev) * args.high_thre
        self.thres_target = torch.tensor(args.high_thres_target, dtype=torch.float32, device=dev)

    def close(self):
        """Drops the captured graphs (they hold the NCCL work of the gradient group: release them before
        torch.distributed.destroy_process_group(), which otherwise waits forever for the communicator)."""
        self._graphs.clear()
        for st in (self.pseudo, self._cam_only, self._cam_aux):
            st._g = None

    # ------------------------------------------------------------------ pieces
    def _forward(self, inputs, inputs_aug=None):
        model = self.model
        net = model.module if hasattr(model, "module") else model
        for n in (net.branch1, net.branch2):
            n._use_kept = self.reuse_forward
            n._use_arena = self._arenas is not None
        if self._arenas is not None:
            for a in self._arenas:
                a.begin_step(1 if inputs_aug is None else 2)   # phase C: the plain and the augmented view both reach every parameter
        if inputs_aug is None:
            res = model(inputs)
        else:
            res = model(torch.cat([inputs, inputs_aug], dim=0), need_sp=True)
        for n in (net.branch1, net.branch2):
            n._use_kept = False
        return res

    def _ptc(self, cams_aux, fmap, cls_label, img_box, high):
        a = self.args
        resized = F.interpolate(cams_aux, size=fmap.shape[2:], mode="bilinear", align_corners=False)
        fn = cam_helper.cam_to_label if not torch.is_tensor(high) else cam_helper.cam_to_label_dynamic_cls
        _, pseudo = fn(resized.detach(), cls_label=cls_label, img_box=img_box, ignore_mid=True, bkg_thre=a.bkg_thre,
                       high_thre=high, low_thre=a.low_thre, ignore_index=a.ignore_index)
        return get_masked_ptc_loss(fmap, cam_helper.label_to_aff_mask(pseudo))

    def losses(self, inputs, cls_label, img_box, n_iter, inputs_aug=None):
        """-> (loss, dict of parts, (refined_label_1, refined_label_2) or None).  inputs_aug: the strongly augmented view
        (imutils.augment_data_strong; CPU PIL in the reference); once n_iter >= gmm_iters it is drawn on the device when None."""
        a = self.args
        cls_f = cls_label.float()
        one = inputs.new_ones(())
        labels = None
        if n_iter < a.cam_iters:
            # ---------------- phase A (train_final_voc.py:194-258, train_final_coco.py:190-231)
            (cams_1, cams_aux_1), (cams_2, cams_aux_2) = self._cam_only(inputs, cls_f, img_box, None)[2:]
            res = self._forward(inputs)
            high_thres = a.high_thre
        else:
            # per-image high threshold = max over the present classes of the cosine-annealed class thresholds
            if self._sched is not None:   # captured: same expression with the two host scalars read from device memory
                omc, done = self._sched
                thres = torch.where(done, self.thres_target, self.thres_start + (self.thres_target - self.thres_start) * omc / 2)
            else:
                thres = cosine_descent(self.thres_start, self.thres_target, n_iter - a.thres_anneal_from, a.max_iters - a.thres_anneal_from)
            high_thres = torch.where(cls_f > 0, thres[None, :], thres.new_full((), -math.inf)).amax(1)
            # multi_scale_cam2_siamese x2 and refine_cams_with_* x2 (train_final_voc.py:279-284, 330-343; coco :312-333).
            # cams * cls_label_rep of the script is the identity on the channels the refine kernels read (one-hot labels).
            pseudo = self._cam_aux if (a.aux_refine_until is not None and n_iter <= a.aux_refine_until) else self.pseudo
            label_1, label_2, (cams_1, cams_aux_1), (cams_2, cams_aux_2) = pseudo(inputs, cls_f, img_box, high_thres)
            labels = (label_1, label_2)
            phase_c = n_iter >= a.gmm_iters
            if phase_c and inputs_aug is None:
                # train_final_voc.py:186-191: inputs_aug = augment_data_strong(denormalize_img2(inputs), n=5, m=10) — Pillow on
                # the host in the reference, seven bit-exact kernels on the device here (utils/imutils.py of this package)
                from .utils import imutils
                inputs_aug = imutils.augment_data_strong(denormalize_img2(inputs.clone()), n=5, m=10, ops=self._aug_ops)
            res = self._forward(inputs, inputs_aug if phase_c else None)
        cls_1, segs_1, fmap_1, cls_aux_1 = res["branch1"]
        cls_2, segs_2, fmap_2, cls_aux_2 = res["branch2"]

        # 4 x F.multilabel_soft_margin_loss (train_final_voc.py:299-305) in one kernel
        cls_loss = multilabel_soft_margin_sum((cls_1, cls_aux_1, cls_2, cls_aux_2), cls_f)

        if n_iter < a.cam_iters and not a.ptc_in_phase_a:
            ptc_loss = one                                     # train_final_coco.py:216
        else:
            ptc_loss = (self._ptc(cams_aux_1, fmap_1, cls_label, img_box, high_thres) +
                        self._ptc(cams_aux_2, fmap_2, cls_label, img_box, high_thres))

        reg_loss = one * 0
        if labels is None:
            seg_loss = one                                     # train_final_voc.py:244
        else:
            label_1, label_2 = labels
            if n_iter >= a.gmm_iters:
                # ---------------- phase C: GMM noise filter on each student's OWN per-pixel CE (train_final_voc.py:358-394)
                size = label_1.shape[1:]
                with torch.no_grad():
                    up_1 = F.interpolate(segs_1, size=size, mode="bilinear", align_corners=False)
                    up_2 = F.interpolate(segs_2, size=size, mode="bilinear", align_corners=False)
                    ce_1 = F.cross_entropy(up_1, label_1.long(), ignore_index=a.ignore_index, reduction="none")
                    ce_2 = F.cross_entropy(up_2, label_2.long(), ignore_index=a.ignore_index, reduction="none")
                    gmm_noise_filter(ce_1, label_1, ignore_index=a.ignore_index, gmm_valid_thre=a.gmm_valid_thre, gamma=a.gamma)
                    gmm_noise_filter(ce_2, label_2, ignore_index=a.ignore_index, gmm_valid_thre=a.gmm_valid_thre, gamma=a.gamma)
            # F.interpolate(segs, size=label.shape[1:]) + get_seg_loss (train_final_voc.py:345-352, 396-401), fused
            seg_loss = get_seg_loss_upsampled(segs_1, label_2, a.ignore_index) + get_seg_loss_upsampled(segs_2, label_1, a.ignore_index)
            if n_iter >= a.gmm_iters:
                # ---------------- consistency regularisation on the augmented view (train_final_voc.py:407-436)
                regs = []
                for up, other_label, aug in ((up_1, label_2, res["branch1_aug"]), (up_2, label_1, res["branch2_aug"])):
                    with torch.no_grad():
                        conf, pseudo_seg = torch.softmax(up, dim=1).max(1)
                        uncertain = (other_label == a.ignore_index) & (conf > 0.9)
                        target = torch.where(uncertain, pseudo_seg, torch.full_like(pseudo_seg, a.ignore_index))
                    # F.interpolate + ce_criterion(...).sum() / uncertain.sum() fused; `if uncertain.sum() > 0` of the script
                    # without the host sync: an empty mask gives 0 / 1 = 0
                    regs.append(ce_sum_upsampled(torch.flip(aug, dims=[3]), target, a.ignore_index))
                reg_loss = regs[0] + regs[1]

        sim_loss = discrepancy_loss(fmap_1, fmap_2, eps=1e-6)     # train_final_voc.py:440-447, fused

        w = a.loss_weights(n_iter)
        loss = w["cls"] * cls_loss + w["ptc"] * ptc_loss + w["seg"] * seg_loss + w["sim"] * sim_loss + w["reg"] * reg_loss
        return loss, dict(cls_loss=cls_loss, ptc_loss=ptc_loss, seg_loss=seg_loss, sim_loss=sim_loss, reg_loss=reg_loss), labels

    def __call__(self, inputs, cls_label, img_box, n_iter, inputs_aug=None):
        if self.capture:
            return self._replay(inputs, cls_label, img_box, n_iter, inputs_aug)
        loss, parts, _ = self.losses(inputs, cls_label, img_box, n_iter, inputs_aug)
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        return loss.detach(), {k: v.detach() for k, v in parts.items()}

    # ------------------------------------------------------------------ data parallel without the DDP reducer
    @staticmethod
    def _world():
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1

    def _backward(self, loss):
        """loss.backward() with the gradients landing in the students' arenas; finished chunks are averaged over the ranks
        while the rest of the backward pass runs (grad_arena.GradArena), the persistent GEMM grid leaving `_comm_sms` SMs to
        the NCCL kernels meanwhile.  On return every .grad is final on the current stream."""
        from . import _lib as L
        limit = self._comm_sms if self._world() > 1 else 0
        prev = 0
        if limit > 0:
            prev = L.lib().dupl_set_gemm_sm_limit(torch.cuda.get_device_properties(loss.device).multi_processor_count - limit)
        try:
            loss.backward()
        finally:
            if limit > 0:
                L.lib().dupl_set_gemm_sm_limit(prev)
        for a in self._arenas:
            a.finish()
            a.bind_grads()

    # ------------------------------------------------------------------ whole-iteration CUDA graph
    def _phase_key(self, n_iter, inputs, cls_label, inputs_aug):
        a = self.args
        w = a.loss_weights(n_iter)
        return (n_iter < a.cam_iters, a.aux_refine_until is not None and n_iter <= a.aux_refine_until, n_iter >= a.gmm_iters,
                tuple(sorted(w.items())), tuple(inputs.shape), tuple(cls_label.shape), cls_label.dtype, inputs_aug is not None)

    def _set_schedule(self, st, n_iter):
        a = self.args
        step, num = n_iter - a.thres_anneal_from, a.max_iters - a.thres_anneal_from
        if step < 0:
            omc, done = 0.0, False
        elif step >= num:
            omc, done = 0.0, True
        else:
            omc, done = float(1 - np.cos(np.pi * (step / (num - 1)))), False
        st["omc"].fill_(omc)
        st["done"].fill_(done)

    def _replay(self, inputs, cls_label, img_box, n_iter, inputs_aug):
        if hasattr(self.model, "module"):
            raise RuntimeError("TrainStep(capture=True) takes the bare siamese_network: with several ranks it all-reduces the "
                               "gradients itself inside the graph (DistributedDataParallel's hooks need the eager autograd pass)")
        dev = inputs.device
        key = self._phase_key(n_iter, inputs, cls_label, inputs_aug)
        st = self._graphs.get(key)
        first = st is None
        if first:
            st = dict(x=torch.empty_like(inputs), cls=torch.empty_like(cls_label),
                      box=torch.empty(inputs.shape[0], 4, dtype=torch.int32, device=dev),
                      aug=None if inputs_aug is None else torch.empty_like(inputs_aug),
                      omc=torch.zeros((), dtype=torch.float32, device=dev), done=torch.zeros((), dtype=torch.bool, device=dev))
            self._graphs[key] = st
        phase_c_own_aug = n_iter >= self.args.gmm_iters and inputs_aug is None
        if phase_c_own_aug:
            from .utils import imutils
            if st.get("aug_ops") is None:
                st["aug_ops"] = torch.zeros(5, inputs.shape[0], dtype=torch.int32, device=dev)
                st["aug_ops_host"] = torch.zeros(5, inputs.shape[0], dtype=torch.int32).pin_memory()
            st["aug_ops_host"].copy_(torch.tensor(imutils.draw_ops(inputs.shape[0], 5), dtype=torch.int32))
            st["aug_ops"].copy_(st["aug_ops_host"], non_blocking=True)
        self._aug_ops = st.get("aug_ops") if phase_c_own_aug else None
        st["x"].copy_(inputs, non_blocking=True)
        st["cls"].copy_(cls_label, non_blocking=True)
        st["box"].copy_(torch.as_tensor(img_box).to(torch.int32), non_blocking=True)
        if inputs_aug is not None:
            st["aug"].copy_(inputs_aug, non_blocking=True)
        self._set_schedule(st, n_iter)
        self._sched = (st["omc"], st["done"])
        try:
            if first:
                # the warm-up runs on a side stream on purpose: torch's stream-mismatch warning for AccumulateGrad nodes is noise here
                if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
                    torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
                # warm-up outside capture (lazy caches, optimizer state tensors, allocator pools); parameters, optimizer
                # state and schedule are put back afterwards, so the first captured call is ONE training step like any other
                params = [p for g in self.optim.param_groups for p in g["params"]]
                saved_p = [p.detach().clone() for p in params]
                fused_opt = hasattr(self.optim, "snapshot")
                saved_s = self.optim.snapshot() if fused_opt else \
                    {p: {k: v.clone() for k, v in s.items() if torch.is_tensor(v)} for p, s in self.optim.state.items()}
                saved_step = self.optim.global_step
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    for _ in range(2):
                        loss, parts, _ = self.losses(st["x"], st["cls"], st["box"], n_iter, st["aug"])
                        self._backward(loss)
                        self.optim.step_captured()
                    with torch.no_grad():
                        for p, q in zip(params, saved_p):
                            p.copy_(q)
                        if fused_opt:
                            self.optim.restore(saved_s)
                        else:
                            for p, s in self.optim.state.items():
                                for k, v in s.items():
                                    if torch.is_tensor(v):
                                        if p in saved_s and k in saved_s[p]:
                                            v.copy_(saved_s[p][k])
                                        else:
                                            v.zero_()
                        if fused_opt:
                            # the weight planes follow the restored parameters; with the fused optimizer they are never
                            # re-split inside the graph (its update kernel rewrites them), so they must be current when capture
                            # starts.  (With torch's AdamW the stale keys are what makes the capture record the re-split.)
                            for b in (self.model.branch1, self.model.branch2):
                                b.planes().refresh_all()
                                b.planes().sync_keys()
                torch.cuda.current_stream(dev).wait_stream(side)
                self.optim.global_step = saved_step
                del saved_p, saved_s
                # the warm-up's autograd graph must be gone before capture: its AccumulateGrad nodes (kept alive through
                # `loss`) belong to the warm-up, and autograd would make the capturing stream wait for their stream at the end
                # of the captured backward ("dependency created on uncaptured work"); capturing on the warm-up's stream makes
                # the two coincide in any case
                del loss, parts
                graph = torch.cuda.CUDAGraph()
                multi = self._world() > 1
                with torch.cuda.graph(graph, stream=side, **(dict(capture_error_mode="thread_local") if multi else {})):
                    loss, parts, _ = self.losses(st["x"], st["cls"], st["box"], n_iter, st["aug"])
                    self._backward(loss)
                    self.optim.step_captured()
                st["graph"], st["loss"], st["parts"] = graph, loss.detach(), {k: v.detach() for k, v in parts.items()}
                if fused_opt:
                    st["opt_tables"] = self.optim.export_tables()
            if st.get("opt_tables") is not None and self._last_graph_key not in (None, key):
                self.optim.import_tables(st["opt_tables"])   # another phase's graph ran in between: its flags / pointers differ
            self._last_graph_key = key
            self.optim.advance_schedule()
            st["graph"].replay()
            # the replay updated the parameters without bumping `_version`: eager users of the weight planes must re-split
            net = self.model
            for b in (net.branch1, net.branch2):
                if hasattr(self.optim, "snapshot"):
                    b.planes().sync_keys()      # the fused optimizer rewrote the planes itself; transposed copies stay graph-internal
                    for name in [n for n in b.planes()._planes if n.endswith("^T")]:
                        del b.planes()._planes[name]
                else:
                    b.planes().invalidate()
                if getattr(b, "_dec_planes", None) is not None:
                    b._dec_planes.invalidate()
        finally:
            self._sched = None
            self._aug_ops = None
        return st["loss"], st["parts"]


PhaseBStep = TrainStep
