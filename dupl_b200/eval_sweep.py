"""Test-time sweep of the reference's eval tools as ONE device-resident pipeline (BASELINE.json configs[4]):

    tools/eval_seg_coco_ddp.py:75-132   multi-scale + flip segmentation logits of both students ("msc_seg")
    tools/eval_seg_voc.py:50-84         same, VOC flavour (max over scales at label resolution)
    utils/train_helper.py:120-150       MS-CAM + cam_to_label of the validation loop
    tools/eval_seg_*.py crf_proc        F.interpolate -> softmax -> DenseCRF(10, 1, 1, 4, 121, 5) -> argmax

The reference writes the logits of every image to .npy files and runs pydensecrf in a joblib CPU pool afterwards
(rank 0 only).  Here the logits never leave the device and the mean-field runs on the GPU that produced them; images
are strided over ranks exactly like tools/eval_seg_coco_ddp.py:241 (`range(rank, N, world)`), no collective.

The interpolate / flip / stack glue below is the scripts' own inline torch code (it is not part of any reference
module); all dense math and the CRF are libdupl.so kernels behind the drop-in modules.
"""
import torch
import torch.nn.functional as F

from .utils import cam_helper
from .utils.dcrf import DenseCRF
from .utils.evaluate import ConfusionMatrix


def shard_indices(n_items, rank, world):
    """tools/eval_seg_coco_ddp.py:241: image i belongs to rank i % world."""
    return list(range(rank, n_items, world))


class SegCrfSweep:
    def __init__(self, model, flavour="coco", scales=None, cam_scales=(1.0, 0.5, 1.5), crop_size=448,
                 crf=None, bkg_thre=0.5, high_thre=0.7, low_thre=0.25, ignore_index=255):
        if flavour not in ("coco", "voc"):
            raise ValueError("flavour must be 'coco' or 'voc'")
        self.model = model
        self.flavour = flavour
        # argparse defaults of the two tools (eval_seg_coco_ddp.py:51, eval_seg_voc.py:35)
        self.scales = tuple(scales) if scales is not None else ((1.0, 1.25, 1.5) if flavour == "coco" else (1.0, 1.5, 1.25))
        self.cam_scales = tuple(cam_scales)
        self.crop_size = crop_size
        self.crf = crf or DenseCRF(iter_max=10, pos_w=1, pos_xy_std=1, bi_w=4, bi_xy_std=121, bi_rgb_std=5)
        self.bkg_thre, self.high_thre, self.low_thre, self.ignore_index = bkg_thre, high_thre, low_thre, ignore_index

    # ------------------------------------------------------------------ multi-scale + flip seg logits
    @torch.no_grad()
    def _pair_segs(self, x):
        """both students on [x, flip(x)]; returns the seg logits of the cat batch (eval_seg_*: `model(inputs_cat)`)."""
        res = self.model(torch.cat([x, x.flip(-1)], dim=0))
        return res["branch1"][1], res["branch2"][1]

    @staticmethod
    def _unflip_sum(segs, b):
        return segs[:b] + segs[b:].flip(-1)

    @torch.no_grad()
    def msc_seg(self, inputs, label_size=None):
        """-> (seg_1, seg_2): what the tools store as {"msc_seg": ...} per image (here for a batch of b images).

        coco (eval_seg_coco_ddp.py:77-122): inputs resized to crop_size^2; the un-scaled pass defines (h_s, w_s); other
        scales are resized to it; flip twin added; SUM over scales.
        voc (eval_seg_voc.py:54-78): every scale is up-sampled to label_size first; flip twin added; MAX over scales."""
        b = inputs.shape[0]
        if self.flavour == "coco":
            inputs = F.interpolate(inputs, size=[self.crop_size, self.crop_size], mode="bilinear", align_corners=False)
            _, _, h, w = inputs.shape
            s1, s2 = self._pair_segs(inputs)
            acc = [self._unflip_sum(s1, b), self._unflip_sum(s2, b)]
            hs, ws = acc[0].shape[-2:]
            for sc in self.scales:
                if sc == 1.0:
                    continue
                x = F.interpolate(inputs, size=[int(h * sc), int(w * sc)], mode="bilinear", align_corners=False)
                for k, s in enumerate(self._pair_segs(x)):
                    s = F.interpolate(s, size=(hs, ws), mode="bilinear", align_corners=False)
                    acc[k] = acc[k] + self._unflip_sum(s, b)
            return acc[0], acc[1]
        if label_size is None:
            label_size = tuple(inputs.shape[-2:])
        _, _, h, w = inputs.shape
        best = [None, None]
        for sc in self.scales:
            x = F.interpolate(inputs, size=[int(h * sc), int(w * sc)], mode="bilinear", align_corners=False)
            for k, s in enumerate(self._pair_segs(x)):
                s = F.interpolate(s, size=label_size, mode="bilinear", align_corners=False)
                s = self._unflip_sum(s, b)
                best[k] = s if best[k] is None else torch.maximum(best[k], s)
        return best[0], best[1]

    # ------------------------------------------------------------------ CRF stage (crf_proc._job)
    @torch.no_grad()
    def crf_prob(self, image_u8, logits):
        """image_u8 uint8 [H,W,3] (cuda), logits [1,C,h,w] -> Q float32 [C,H,W] on the device."""
        H, W = image_u8.shape[:2]
        logit = F.interpolate(logits, size=(H, W), mode="bilinear", align_corners=False)
        prob = F.softmax(logit, dim=1)[0]
        return self.crf(image_u8, prob)

    # ------------------------------------------------------------------ CAM pseudo labels (validate_siamase*)
    @torch.no_grad()
    def cam_label(self, inputs, cls_label, label_size, branch=1):
        x = F.interpolate(inputs, size=[self.crop_size, self.crop_size], mode="bilinear", align_corners=False)
        cams, _aux = cam_helper.multi_scale_cam2_siamese(self.model, inputs=x, scales=self.cam_scales, branch=branch)
        resized = F.interpolate(cams, size=label_size, mode="bilinear", align_corners=False)
        return cam_helper.cam_to_label(resized, cls_label, bkg_thre=self.bkg_thre, high_thre=self.high_thre,
                                       low_thre=self.low_thre, ignore_index=self.ignore_index)

    @torch.no_grad()
    def __call__(self, images_u8, inputs, cls_label, branch=1):
        """One batch of the sweep.  images_u8: list of uint8 [H_i,W_i,3] cuda tensors (original images, CRF appearance
        term); inputs [b,3,H,W] normalised (already at crop size for coco); cls_label [b,K].
        Returns dict(seg_pred=[...], crf_pred=[...], cam_label=Tensor) with per-image int64 label maps on the device."""
        b = inputs.shape[0]
        label_size = tuple(images_u8[0].shape[:2])
        seg_1, seg_2 = self.msc_seg(inputs, label_size)
        seg = seg_1 if branch == 1 else seg_2
        out = dict(seg_pred=[], crf_pred=[])
        for i in range(b):
            H, W = images_u8[i].shape[:2]
            up = F.interpolate(seg[i:i + 1], size=(H, W), mode="bilinear", align_corners=False)
            out["seg_pred"].append(up.argmax(1)[0])
            out["crf_pred"].append(self.crf_prob(images_u8[i], seg[i:i + 1]).argmax(0))
        out["cam_label"] = self.cam_label(inputs, cls_label, label_size, branch)
        return out

    @torch.no_grad()
    def validate(self, samples, num_classes, rank=0, world=1, branch=1):
        """validate_siamase(_coco) + crf_proc over a list of samples (utils/train_helper.py:90-283, tools/eval_seg_*.py): every
        rank takes images rank, rank + world, ... (tools/eval_seg_coco_ddp.py:241), accumulates three confusion matrices on
        its device (seg arg-max, CRF arg-max, CAM pseudo-labels) and ONE C x C all-reduce per matrix merges them.
        samples[i] = (image_u8 [H,W,3] uint8, inputs [1,3,h,w] normalised, cls_label [1,K], label [H,W] int).
        Returns {"seg": scores, "crf": scores, "cam": scores} (utils/evaluate.py:17-60 semantics)."""
        dev = next(self.model.parameters()).device
        cms = {k: ConfusionMatrix(num_classes, dev) for k in ("seg", "crf", "cam")}
        for i in shard_indices(len(samples), rank, world):
            img, x, cls, gt = (t.to(dev, non_blocking=True) for t in samples[i])
            out = self([img], x, cls, branch=branch)
            cms["seg"].update(gt, out["seg_pred"][0])
            cms["crf"].update(gt, out["crf_pred"][0])
            cms["cam"].update(gt, out["cam_label"][0], pseudo=True)
        return {k: cm.all_reduce().scores() for k, cm in cms.items()}


# ---------------------------------------------------------------------------------------------------------------------
# Interop with the files the reference's tools write and read
# ---------------------------------------------------------------------------------------------------------------------
def load_checkpoint(model, path_or_state_dict, map_location="cpu"):
    """tools/eval_seg_voc.py:172-177 / eval_seg_coco_ddp.py: the training script saves `model.state_dict()` of the
    DistributedDataParallel wrapper (train_final_voc.py:508), so every key carries a `module.` prefix; the tools strip it and
    load with strict=True.  Same here (into a dupl_b200 siamese_network, whose state-dict schema is the reference's)."""
    from collections import OrderedDict
    sd = torch.load(path_or_state_dict, map_location=map_location) if isinstance(path_or_state_dict, (str, bytes)) or hasattr(
        path_or_state_dict, "__fspath__") else path_or_state_dict
    new = OrderedDict((k.replace("module.", ""), v) for k, v in sd.items())
    model.load_state_dict(state_dict=new, strict=True)
    return model


def save_msc_seg(path, seg):
    """tools/eval_seg_voc.py:83-84: the multi-scale logits of one image as a pickled dict `{"msc_seg": float32[1,C,h,w]}`."""
    import numpy as np
    np.save(path, {"msc_seg": seg.detach().float().cpu().numpy()})


def load_msc_seg(path, device=None):
    """tools/eval_seg_voc.py:116-117 (`np.load(..., allow_pickle=True).item()["msc_seg"]`) -> float32 tensor [1,C,h,w]."""
    import numpy as np
    arr = np.load(path, allow_pickle=True).item()["msc_seg"]
    t = torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32))
    return t.to(device) if device is not None else t


def crf_from_logits_file(logit_path, image_u8, crf=None, device="cuda"):
    """crf_proc._job (tools/eval_seg_voc.py:113-137) for one image on the GPU: logits file + uint8 image [H,W,3] (numpy or
    tensor) -> (pred uint8 [H,W] numpy, Q float32 [C,H,W] cuda)."""
    import numpy as np
    crf = crf or DenseCRF(iter_max=10, pos_w=1, pos_xy_std=1, bi_w=4, bi_xy_std=121, bi_rgb_std=5)
    img = torch.as_tensor(np.ascontiguousarray(image_u8) if not torch.is_tensor(image_u8) else image_u8).to(device=device, dtype=torch.uint8)
    H, W = img.shape[:2]
    logit = F.interpolate(load_msc_seg(logit_path, device), size=(H, W), mode="bilinear", align_corners=False)
    q = crf(img, F.softmax(logit, dim=1)[0])
    return q.argmax(0).to(torch.uint8).cpu().numpy(), q


def save_label_png(path, pred):
    """imageio.imsave(segs_dir/name.png, pred.astype(uint8)) of the tools (tools/eval_seg_voc.py:140): 8-bit single-channel PNG."""
    import numpy as np
    from PIL import Image
    arr = pred.detach().cpu().numpy() if torch.is_tensor(pred) else np.asarray(pred)
    Image.fromarray(np.squeeze(arr).astype(np.uint8)).save(path)
