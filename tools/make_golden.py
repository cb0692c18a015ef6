"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference, imported on CPU
through oracle/ref_import.py) on seeded inputs.  Run in the authoring container:

    python tools/make_golden.py

The fixtures pin oracle/dupl_oracle.py (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_gpu_golden.py, GPU) to the reference's own outputs.  Model weights are not stored: they are
regenerated bit-identically from tests/helpers.init_state_dict (CPU generator, fixed seed).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import init_state_dict, synth_boxes, synth_cls_labels, synth_images  # noqa: E402
from oracle import ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DIL = [1, 2, 4, 8, 12, 24]


def smooth_img(b, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    u8 = torch.randint(0, 256, (b, 3, h, w), generator=g).float()
    u8 = torch.nn.functional.avg_pool2d(u8, 5, stride=1, padding=2, count_include_pad=False).round()
    return u8 / 255.0


def main():
    torch.set_num_threads(8)
    ref = ref_import.load()
    os.makedirs(OUT, exist_ok=True)
    npz = lambda name, **kw: np.savez_compressed(os.path.join(OUT, name), **{k: (v.numpy() if torch.is_tensor(v) else v) for k, v in kw.items()})

    # ---- A6 PAR.forward (model/PAR.py:64-91)
    par = ref.PAR.PAR(num_iter=10, dilations=DIL)
    img = smooth_img(1, 40, 56, seed=11)
    g = torch.Generator().manual_seed(12)
    masks = torch.rand(1, 3, 40, 56, generator=g).softmax(1)
    with torch.no_grad():
        out = par(img, masks)
    npz("par.npz", imgs=img, masks=masks, out=out, dilations=np.array(DIL), num_iter=10)

    # ---- A5a/b/c cam_to_label(_dynamic_cls), label_to_aff_mask (utils/cam_helper.py:8-55,323-335)
    g = torch.Generator().manual_seed(13)
    cam = torch.rand(4, 20, 28, 28, generator=g)
    cam[0, :, 0, 0] = 0.5
    cls = synth_cls_labels(4, 20, seed=14)
    box = synth_boxes(4, 448, 448, seed=15)
    box[1] = torch.tensor([30, 448, 0, 10])
    box[2] = torch.tensor([3, 20, 5, 26])
    kw = dict(bkg_thre=0.45, low_thre=0.25, ignore_mid=True, ignore_index=255)
    v1, l1 = ref.cam_helper.cam_to_label(cam.clone(), cls, box, high_thre=0.65, **kw)
    ht = torch.tensor([0.6, 0.7, 0.55, 0.65])
    v2, l2 = ref.cam_helper.cam_to_label_dynamic_cls(cam.clone(), cls, box, high_thre=ht, **kw)
    l3 = ref.cam_helper.cam_to_label(cam.clone(), cls, bkg_thre=0.45)
    small = l1[:, :7, :9].contiguous()
    aff = ref.cam_helper.label_to_aff_mask(small)
    npz("cam_to_label.npz", cam=cam, cls=cls, box=box, valid=v1, label=l1, high_thre_dyn=ht, label_dyn=l2, label_nobox=l3,
        aff_in=small, aff=aff)

    # ---- A7 refine_cams_with_dynamic_thres / refine_cams_with_bkg_v2 (utils/cam_helper.py:338-431)
    b, K, H, W = 3, 20, 64, 96
    imgs = smooth_img(b, H, W, seed=16)
    cls = synth_cls_labels(b, K, seed=17)
    g = torch.Generator().manual_seed(18)
    cams = torch.rand(b, K, H, W, generator=g).half().float() * cls[:, :, None, None]  # fp16-representable: compresses
    box = synth_boxes(b, H, W, seed=19)
    htm = torch.stack([torch.full((1, H, W), v) for v in (0.6, 0.7, 0.55)])
    with torch.no_grad():
        r_dyn = ref.cam_helper.refine_cams_with_dynamic_thres(par, imgs, cams, cls, high_thre_map=htm, low_thre=0.25,
                                                              ignore_index=255, img_box=box)
        r_v2 = ref.cam_helper.refine_cams_with_bkg_v2(par, imgs, cams, cls, high_thre=0.65, low_thre=0.25,
                                                      ignore_index=255, img_box=box)
    npz("refine.npz", images=imgs, cams=cams, cls=cls, box=box, high_thre_map=htm[:, :, :1, :1],
        label_dyn=r_dyn.to(torch.uint8), label_v2=r_v2.to(torch.uint8))

    # ---- A1/A2/A4 model: cam_only, train-mode forward, multi_scale_cam2_siamese (random-init weights, seed 0)
    P = init_state_dict(21)
    model = ref.model_dupl.siamese_network("deit_base_patch16_224", num_classes=21, pretrained=False, aux_layer=-3)
    model.load_state_dict(P, strict=True)
    model.eval()
    x = synth_images(2, 32, 48, seed=20)
    with torch.no_grad():
        ca1, c1, ca2, c2 = model(x, cam_only=True)
        res = model(x)
        cam, cam_aux = ref.cam_helper.multi_scale_cam2_siamese(model, x, (1.0, 0.5, 1.5), branch=2)
    npz("model.npz", x=x, cam_aux_1=ca1, cam_1=c1, cam_aux_2=ca2, cam_2=c2,
        cls_1=res["branch1"][0], seg_1=res["branch1"][1], fmap_1=res["branch1"][2], cls_aux_1=res["branch1"][3],
        cls_2=res["branch2"][0], seg_2=res["branch2"][1], cls_aux_2=res["branch2"][3],
        mscam_2=cam, mscam_aux_2=cam_aux)

    # ---- A5d / A8 losses (model/losses.py:6-39)
    g = torch.Generator().manual_seed(21)
    fmap = torch.randn(2, 768, 7, 9, generator=g)
    ptc = ref.losses.get_masked_ptc_loss(fmap, aff[:2])
    pred = torch.randn(3, 21, 64, 96, generator=g).half().float()
    seg = ref.losses.get_seg_loss(pred, r_dyn.long(), ignore_index=255)
    npz("losses.npz", fmap=fmap, aff=aff[:2], ptc=ptc, pred=pred.half(), label=r_dyn.to(torch.uint8), seg=seg)

    # ---- A0 denormalize_img2 (utils/imutils.py:17-31)
    xin = synth_images(1, 16, 16, seed=22)
    npz("denorm.npz", x=xin, out=ref.imutils.denormalize_img2(xin.clone()))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
