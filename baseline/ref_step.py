"""The loop body of the reference's training script (train_final_voc.py:186-472) executed with the reference's OWN modules
(model.model_dupl, utils.cam_helper, model.PAR, model.losses, utils.imutils, utils.optimizer from baseline/_ref) on stock
PyTorch — the comparator of the north_star's ">= 5x the reference's stock PyTorch-GPU images/sec" target and, on the host
cores, of `bench.py --impl reference`.  None of dupl_b200's modules, kernels or library is imported here.

The body is inline in the reference's `train()` (not importable), so it is restated below statement by statement with the
script's line numbers; skipped: progress printing, tensorboard, validation / checkpointing (`:474-538`).  Everything that
costs time in the script is kept, including its host synchronisations (`torch.nonzero` per image `:268-275`, the sklearn
F1 and the six `.item()` reads `:458-468`) and, optionally, the PIL RandAugment round trip of every iteration (`:191`).
"""
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import compat

VOC_ARGS = dict(  # defaults of train_final_voc.py:33-90
    backbone="deit_base_patch16_224", num_classes=21, aux_layer=-3, samples_per_gpu=4, optimizer="PolyWarmupAdamW",
    warmup_iters=1500, lr=6e-5, warmup_lr=1e-6, wt_decay=1e-2, betas=(0.9, 0.999), power=0.9, ignore_index=255,
    max_iters=20000, cam_iters=2000, high_thre=0.7, low_thre=0.25, bkg_thre=0.5, cam_scales=(1.0, 0.5, 1.5), w_ptc=0.2,
    w_seg=0.2, gmm_iters=8000, gmm_valid_thre=1.0, gamma=0.95)


def cosine_descent(max_thres, min_thres, step, num_steps):
    """utils/train_helper.py:340-349 (train_helper itself imports the dataset modules; this is its 8-line schedule)."""
    if step < 0:
        return max_thres
    if step >= num_steps:
        return min_thres
    f = step / (num_steps - 1)
    return max_thres + (min_thres - max_thres) * (1 - np.cos(np.pi * f)) / 2


class ReferenceStep:
    def __init__(self, device, state_dict=None, samples_per_gpu=4, ddp=False, local_rank=0, strong_aug=False, seed=0):
        self.ref = ref = compat.load_reference()
        self.args = args = types.SimpleNamespace(**VOC_ARGS)
        args.samples_per_gpu = samples_per_gpu
        self.device = device
        self.strong_aug = strong_aug
        # setup_seed, train_final_voc.py:95-102
        torch.manual_seed(seed)
        if device.type == "cuda":
            torch.cuda.manual_seed_all(seed)
        np.random.seed(seed)
        torch.backends.cudnn.benchmark = False
        torch.backends.cudnn.deterministic = True
        model = ref.model_dupl.siamese_network(backbone=args.backbone, num_classes=args.num_classes, pretrained=False,
                                               aux_layer=args.aux_layer)                        # :145-150
        if state_dict is not None:
            model.load_state_dict(state_dict, strict=True)
        param_groups = model.get_param_groups()                                                    # :152
        model.to(device)                                                                           # :154
        if ddp:
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], find_unused_parameters=True)  # :155
        model.train()
        self.model = model
        self.optim = getattr(ref.optimizer, args.optimizer)(                                       # train_helper.py:21-52
            params=[{"params": param_groups[0], "lr": args.lr, "weight_decay": args.wt_decay},
                    {"params": param_groups[1], "lr": args.lr, "weight_decay": args.wt_decay},
                    {"params": param_groups[2], "lr": args.lr * 10, "weight_decay": args.wt_decay},
                    {"params": param_groups[3], "lr": args.lr * 10, "weight_decay": args.wt_decay}],
            lr=args.lr, weight_decay=args.wt_decay, betas=args.betas, warmup_iter=args.warmup_iters, max_iter=args.max_iters,
            warmup_ratio=args.warmup_lr, power=args.power)
        self.par = ref.PAR.PAR(num_iter=10, dilations=[1, 2, 4, 8, 12, 24]).to(device)             # :160
        self.high_thres_start = torch.ones(20, device=device) * args.high_thre                     # :162
        self.high_thres_target = torch.tensor([0.70, 0.70, 0.70, 0.70, 0.55, 0.55, 0.55, 0.55, 0.70, 0.55,
                                               0.55, 0.55, 0.55, 0.55, 0.55, 0.55, 0.55, 0.55, 0.70, 0.55]).to(device)
        self.ce_criterion = nn.CrossEntropyLoss(ignore_index=args.ignore_index, reduction="none").to(device)

    def __call__(self, inputs, cls_label, img_box, n_iter):
        """inputs / cls_label: CPU (or device) tensors as the DataLoader yields them; img_box int16 CPU."""
        ref, args, model, par, device = self.ref, self.args, self.model, self.par, self.device
        cam_helper, imutils = ref.cam_helper, ref.imutils
        get_masked_ptc_loss, get_seg_loss = ref.losses.get_masked_ptc_loss, ref.losses.get_seg_loss
        ce_criterion = self.ce_criterion
        one = lambda v: torch.full((1,), float(v), device=device)                                   # noqa: E731  (torch.ones(1).cuda())

        inputs = inputs.to(device, non_blocking=True)                                              # :185
        inputs_denorm = imutils.denormalize_img2(inputs.clone())                                   # :186
        cls_label = cls_label.to(device, non_blocking=True)                                        # :187
        if self.strong_aug:                                                                        # :190-191 (PIL, host)
            inputs_aug = imutils.augment_data_strong(inputs_denorm.clone(), n=5, m=10)
        else:
            inputs_aug = None

        if n_iter < args.cam_iters:                                                                # :193-258
            cams_1, cams_aux_1 = cam_helper.multi_scale_cam2_siamese(model, inputs=inputs, scales=args.cam_scales, branch=1)
            cams_2, cams_aux_2 = cam_helper.multi_scale_cam2_siamese(model, inputs=inputs, scales=args.cam_scales, branch=2)
            res = model(inputs)
            cls_1, segs_1, fmap_1, cls_aux_1 = res["branch1"]
            cls_2, segs_2, fmap_2, cls_aux_2 = res["branch2"]
            cls_loss = (F.multilabel_soft_margin_loss(cls_1, cls_label) + F.multilabel_soft_margin_loss(cls_aux_1, cls_label) +
                        F.multilabel_soft_margin_loss(cls_2, cls_label) + F.multilabel_soft_margin_loss(cls_aux_2, cls_label))
            ptc_loss = 0
            for cams_aux, fmap in ((cams_aux_1, fmap_1), (cams_aux_2, fmap_2)):
                resized = F.interpolate(cams_aux, size=fmap.shape[2:], mode="bilinear", align_corners=False)
                _, pseudo = cam_helper.cam_to_label(resized.detach(), cls_label=cls_label, img_box=img_box, ignore_mid=True,
                                                    bkg_thre=args.bkg_thre, high_thre=args.high_thre, low_thre=args.low_thre,
                                                    ignore_index=args.ignore_index)
                ptc_loss = ptc_loss + get_masked_ptc_loss(fmap, cam_helper.label_to_aff_mask(pseudo))
            seg_loss = one(1)
            reg_loss = one(0)
        else:
            high_thres = cosine_descent(self.high_thres_start, self.high_thres_target, n_iter - args.cam_iters,
                                        args.max_iters - args.cam_iters)                           # :263-265
            b, _, h, w = inputs.shape
            high_thres_mask_list, high_thres_list = [], []
            for i in range(args.samples_per_gpu):                                                  # :268-272 (host sync per image)
                high_thres_i = torch.max(high_thres[torch.nonzero(cls_label[i]).squeeze(-1)])
                high_thres_list.append(high_thres_i)
                high_thres_mask_list.append(torch.ones((h, w), device=device) * high_thres_i)
            high_thres = torch.stack(high_thres_list, dim=0)
            high_thres_mask = torch.stack(high_thres_mask_list, dim=0).unsqueeze(1)
            cams_1, cams_aux_1 = cam_helper.multi_scale_cam2_siamese(model, inputs=inputs, scales=args.cam_scales, branch=1)
            cams_2, cams_aux_2 = cam_helper.multi_scale_cam2_siamese(model, inputs=inputs, scales=args.cam_scales, branch=2)
            if n_iter < args.gmm_iters:                                                            # :286-296
                res = model(inputs)
            else:
                if inputs_aug is None:
                    raise ValueError("n_iter >= gmm_iters needs strong_aug=True (the script always computes inputs_aug)")
                res = model(torch.cat([inputs, inputs_aug], dim=0), need_sp=True)
                segs_1_aug, segs_2_aug = res["branch1_aug"], res["branch2_aug"]
            cls_1, segs_1, fmap_1, cls_aux_1 = res["branch1"]
            cls_2, segs_2, fmap_2, cls_aux_2 = res["branch2"]
            cls_loss = (F.multilabel_soft_margin_loss(cls_1, cls_label) + F.multilabel_soft_margin_loss(cls_aux_1, cls_label) +
                        F.multilabel_soft_margin_loss(cls_2, cls_label) + F.multilabel_soft_margin_loss(cls_aux_2, cls_label))
            ptc_loss = 0
            for cams_aux, fmap in ((cams_aux_1, fmap_1), (cams_aux_2, fmap_2)):                    # :307-325
                resized = F.interpolate(cams_aux, size=fmap.shape[2:], mode="bilinear", align_corners=False)
                _, pseudo = cam_helper.cam_to_label_dynamic_cls(resized.detach(), cls_label=cls_label, img_box=img_box,
                                                                ignore_mid=True, bkg_thre=args.bkg_thre, high_thre=high_thres,
                                                                low_thre=args.low_thre, ignore_index=args.ignore_index)
                ptc_loss = ptc_loss + get_masked_ptc_loss(fmap, cam_helper.label_to_aff_mask(pseudo))
            b, c, h, w = cams_1.shape                                                              # :329-343
            cls_label_rep = cls_label.unsqueeze(-1).unsqueeze(-1).repeat([1, 1, h, w])
            refined_pseudo_label_1 = cam_helper.refine_cams_with_dynamic_thres(
                par, inputs_denorm, cams=cams_1.detach() * cls_label_rep, cls_labels=cls_label, high_thre_map=high_thres_mask,
                low_thre=args.low_thre, ignore_index=args.ignore_index, img_box=img_box)
            refined_pseudo_label_2 = cam_helper.refine_cams_with_dynamic_thres(
                par, inputs_denorm, cams=cams_2.detach() * cls_label_rep, cls_labels=cls_label, high_thre_map=high_thres_mask,
                low_thre=args.low_thre, ignore_index=args.ignore_index, img_box=img_box)
            segs_1 = F.interpolate(segs_1, size=refined_pseudo_label_1.shape[1:], mode="bilinear", align_corners=False)
            segs_2 = F.interpolate(segs_2, size=refined_pseudo_label_2.shape[1:], mode="bilinear", align_corners=False)
            if n_iter < args.gmm_iters:                                                            # :348-356
                seg_loss_1 = get_seg_loss(segs_1, refined_pseudo_label_2.type(torch.long))
                seg_loss_2 = get_seg_loss(segs_2, refined_pseudo_label_1.type(torch.long))
                seg_loss = seg_loss_1 + seg_loss_2
                reg_loss = seg_loss_1 * 0 + seg_loss_2 * 0
            else:                                                                                  # :358-436
                from sklearn.mixture import GaussianMixture
                seg_loss_1 = ce_criterion(segs_1, refined_pseudo_label_1.type(torch.long)).detach()
                seg_loss_2 = ce_criterion(segs_2, refined_pseudo_label_2.type(torch.long)).detach()
                roi_mask_1 = (refined_pseudo_label_1 != 0).bool() & (refined_pseudo_label_1 != 255).bool()
                roi_mask_2 = (refined_pseudo_label_2 != 0).bool() & (refined_pseudo_label_2 != 255).bool()
                for i in range(args.samples_per_gpu):
                    for seg_loss_s, roi, lab in ((seg_loss_1, roi_mask_1, refined_pseudo_label_1),
                                                 (seg_loss_2, roi_mask_2, refined_pseudo_label_2)):
                        m = seg_loss_s[i][roi[i]]
                        if (m > 0.1).sum().item() > 1000:
                            gmm = GaussianMixture(n_components=2, max_iter=10, tol=1e-2, reg_covar=5e-4, random_state=0)
                            gmm.fit(m[m > 0.1].unsqueeze(-1).cpu().detach().numpy())
                            means = gmm.means_
                            if abs(means[0, 0] - means[1, 0]) > args.gmm_valid_thre:
                                noise_idx = gmm.means_.argmax()
                                prob = gmm.predict_proba(seg_loss_s[i].view(-1).unsqueeze(-1).cpu().detach().numpy())
                                noise_mask = torch.tensor(prob[:, noise_idx] > args.gamma, device=device).reshape(h, w)
                                noise_mask = noise_mask & (lab[i] != 0).bool()
                                lab[i][noise_mask] = 255
                seg_loss_1 = get_seg_loss(segs_1, refined_pseudo_label_2.type(torch.long), ignore_index=args.ignore_index)
                seg_loss_2 = get_seg_loss(segs_2, refined_pseudo_label_1.type(torch.long), ignore_index=args.ignore_index)
                seg_loss = seg_loss_1 + seg_loss_2
                segs_1_aug = F.interpolate(torch.flip(segs_1_aug, dims=[3]), size=inputs_denorm.shape[2:], mode="bilinear", align_corners=False)
                segs_2_aug = F.interpolate(torch.flip(segs_2_aug, dims=[3]), size=inputs_denorm.shape[2:], mode="bilinear", align_corners=False)
                pseudo_seg_1 = segs_1.detach().data.max(1)[1]
                pseudo_seg_2 = segs_2.detach().data.max(1)[1]
                confidence_map_1 = torch.softmax(segs_1.detach(), dim=1).max(1)[0]
                confidence_map_2 = torch.softmax(segs_2.detach(), dim=1).max(1)[0]
                uncertain_mask_1 = (refined_pseudo_label_2 == args.ignore_index).bool() & (confidence_map_1 > 0.9)
                uncertain_mask_2 = (refined_pseudo_label_1 == args.ignore_index).bool() & (confidence_map_2 > 0.9)
                pseudo_seg_1[~uncertain_mask_1] = args.ignore_index
                pseudo_seg_2[~uncertain_mask_2] = args.ignore_index
                reg_loss_1, reg_loss_2 = seg_loss_1 * 0.0, seg_loss_2 * 0.0
                if uncertain_mask_1.sum() > 0:
                    reg_loss_1 = (ce_criterion(segs_1_aug, pseudo_seg_1)).sum() / uncertain_mask_1.sum()
                if uncertain_mask_2.sum() > 0:
                    reg_loss_2 = (ce_criterion(segs_2_aug, pseudo_seg_2)).sum() / uncertain_mask_2.sum()
                reg_loss = reg_loss_1 + reg_loss_2

        fmap_1_flat = fmap_1.view(fmap_1.shape[0], fmap_1.shape[1], -1)                            # :438-447 (and :246-253)
        fmap_2_flat = fmap_2.view(fmap_2.shape[0], fmap_2.shape[1], -1)
        cos_simi = nn.CosineSimilarity(dim=-1, eps=1e-6)
        sim_loss = (1 + cos_simi(fmap_1_flat.detach(), fmap_2_flat).mean()) + (1 + cos_simi(fmap_2_flat.detach(), fmap_1_flat).mean())

        if n_iter <= args.cam_iters:                                                               # :451-456
            loss = 1.0 * cls_loss + args.w_ptc * ptc_loss + 0.0 * seg_loss + 0.1 * sim_loss
        elif n_iter <= args.gmm_iters:
            loss = 1.0 * cls_loss + args.w_ptc * ptc_loss + args.w_seg * seg_loss + 0.1 * sim_loss + 0.00 * reg_loss
        else:
            loss = 1.0 * cls_loss + args.w_ptc * ptc_loss + args.w_seg * seg_loss + 0.1 * sim_loss + 0.05 * reg_loss

        cls_pred = (cls_1 > 0).type(torch.int16)                                                   # :458-468
        cls_score = ref.evaluate.multilabel_score(cls_label.cpu().numpy()[0], cls_pred.cpu().numpy()[0])
        parts = {"cls_score": cls_score.item(), "cls_loss": cls_loss.item(), "ptc_loss": ptc_loss.item(),
                 "seg_loss": seg_loss.item(), "sim_loss": sim_loss.item(), "reg_loss": reg_loss.item()}

        self.optim.zero_grad()                                                                     # :470-472
        loss.backward()
        self.optim.step()
        return loss.detach(), parts
