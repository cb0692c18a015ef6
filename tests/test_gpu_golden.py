"""GPU: the CUDA path against the reference's own outputs (tests/golden/*.npz, made by tools/make_golden.py
from the unmodified reference).  Labels bit-exact where the inputs are identical; floats to tolerance."""
import os

import numpy as np
import pytest
import torch

from helpers import init_state_dict, mscam_err, rel_err

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return {k: torch.from_numpy(v) if v.ndim else v for k, v in np.load(os.path.join(G, name)).items()}


def test_par_forward_golden():
    from dupl_b200.model.PAR import PAR
    d = load("par.npz")
    par = PAR(num_iter=int(d["num_iter"]), dilations=[int(x) for x in d["dilations"]]).cuda()
    out = par(d["imgs"].cuda(), d["masks"].cuda())
    assert (out.cpu() - d["out"]).abs().max().item() < 2e-6


def test_cam_to_label_golden_bit_exact():
    from dupl_b200.utils import cam_helper, camutils
    d = load("cam_to_label.npz")
    kw = dict(bkg_thre=0.45, low_thre=0.25, ignore_mid=True, ignore_index=255)
    v, lab = cam_helper.cam_to_label(d["cam"].cuda(), d["cls"].cuda(), d["box"], high_thre=0.65, **kw)
    assert torch.equal(lab.cpu(), d["label"]) and torch.equal(v.cpu(), d["valid"])
    _, lab = cam_helper.cam_to_label_dynamic_cls(d["cam"].cuda(), d["cls"].cuda(), d["box"], high_thre=d["high_thre_dyn"].cuda(), **kw)
    assert torch.equal(lab.cpu(), d["label_dyn"])
    assert torch.equal(camutils.cam_to_label(d["cam"].cuda(), d["cls"].cuda(), bkg_thre=0.45).cpu(), d["label_nobox"])
    assert torch.equal(cam_helper.label_to_aff_mask(d["aff_in"].cuda()).cpu(), d["aff"])
    assert torch.equal(cam_helper.get_valid_cam(d["cam"].cuda(), d["cls"].cuda()).cpu(), d["valid"])


def test_refine_golden():
    """End-to-end labels vs the reference: PAR runs in a different summation order (1e-7 differences), so a
    label may flip only at an arg-max near-tie; allow < 1e-4 of the pixels."""
    from dupl_b200.model.PAR import PAR
    from dupl_b200.utils import cam_helper
    d = load("refine.npz")
    b, _, H, W = d["images"].shape
    par = PAR(num_iter=10, dilations=[1, 2, 4, 8, 12, 24]).cuda()
    htm = d["high_thre_map"].expand(b, 1, H, W).contiguous()
    got = cam_helper.refine_cams_with_dynamic_thres(par, d["images"].cuda(), d["cams"].cuda(), d["cls"].cuda(),
                                                    high_thre_map=htm.cuda(), low_thre=0.25, ignore_index=255, img_box=d["box"])
    assert (got.cpu() != d["label_dyn"].float()).float().mean().item() < 1e-4
    got = cam_helper.refine_cams_with_bkg_v2(par, d["images"].cuda(), d["cams"].cuda(), d["cls"].cuda(), high_thre=0.65,
                                             low_thre=0.25, ignore_index=255, img_box=d["box"])
    assert (got.cpu() != d["label_v2"].float()).float().mean().item() < 1e-4


def test_model_golden():
    from dupl_b200.model.model_dupl import siamese_network
    from dupl_b200.utils import cam_helper
    from oracle import dupl_oracle as O
    d = load("model.npz")
    P = init_state_dict(21)
    m = siamese_network("deit_base_patch16_224", num_classes=21, pretrained=False, aux_layer=-3)
    m.load_state_dict(P, strict=True)
    m = m.cuda().eval()
    ca1, c1, ca2, c2 = m(d["x"].cuda(), cam_only=True)
    for got, want in ((ca1, d["cam_aux_1"]), (c1, d["cam_1"]), (ca2, d["cam_aux_2"]), (c2, d["cam_2"])):
        assert rel_err(got, want) < 1e-3
    cam, aux = cam_helper.multi_scale_cam2_siamese(m, d["x"].cuda(), (1.0, 0.5, 1.5), branch=2)
    _, _, osum, oaux_sum = O.multi_scale_cam(P, 2, d["x"], (1.0, 0.5, 1.5), return_sums=True)
    assert mscam_err(cam, d["mscam_2"], osum) < 1e-3
    assert mscam_err(aux, d["mscam_aux_2"], oaux_sum) < 1e-3


def test_val_forward_golden():
    """model(x) outputs of the reference (cls, seg, fmap, cls_aux) — model_dupl.py:86-106."""
    from dupl_b200.model.model_dupl import siamese_network
    d = load("model.npz")
    P = init_state_dict(21)
    m = siamese_network("deit_base_patch16_224", num_classes=21, pretrained=False, aux_layer=-3)
    m.load_state_dict(P, strict=True)
    m = m.cuda().eval()
    with torch.no_grad():
        res = m(d["x"].cuda(), val=True)
    cls1, seg1, fmap1, aux1 = res["branch1"]
    for got, want in ((cls1, d["cls_1"]), (seg1, d["seg_1"]), (fmap1, d["fmap_1"]), (aux1, d["cls_aux_1"]),
                      (res["branch2"][0], d["cls_2"]), (res["branch2"][1], d["seg_2"]), (res["branch2"][3], d["cls_aux_2"])):
        assert got.shape == want.shape
        assert rel_err(got, want) < 1e-3
