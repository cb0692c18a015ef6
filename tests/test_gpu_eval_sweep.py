"""GPU: the eval sweep (multi-scale + flip seg logits -> softmax -> DenseCRF, MS-CAM labels) of
tools/eval_seg_coco_ddp.py / tools/eval_seg_voc.py / utils/train_helper.py:90-283 through dupl_b200.eval_sweep,
against the CPU oracle (seg logits) and the C restatement of the mean-field (CRF stage)."""
import numpy as np
import pytest
import torch

from helpers import init_state_dict, rel_err, synth_cls_labels, synth_images

pytestmark = pytest.mark.gpu


def _model(num_classes):
    from dupl_b200.model.model_dupl import siamese_network
    P = init_state_dict(num_classes)
    m = siamese_network("deit_base_patch16_224", num_classes=num_classes, pretrained=False, aux_layer=-3 if num_classes == 21 else 9)
    m.load_state_dict(P, strict=True)
    return m.cuda().eval(), P


def _u8_image(x):
    from oracle import dupl_oracle as O
    return (O.denormalize_img2(x)[0] * 255.0).round().permute(1, 2, 0).to(torch.uint8).contiguous()


@pytest.mark.parametrize("flavour,num_classes", [("coco", 81), ("voc", 21)])
def test_msc_seg_logits_match_oracle(flavour, num_classes):
    from dupl_b200.eval_sweep import SegCrfSweep
    from oracle import dupl_oracle as O
    m, P = _model(num_classes)
    inputs = synth_images(1, 48, 80, seed=31)
    sweep = SegCrfSweep(m, flavour=flavour, crop_size=64)
    s1, s2 = sweep.msc_seg(inputs.cuda(), label_size=(48, 80))
    aux = -3 if num_classes == 21 else 9
    with torch.no_grad():
        w1 = O.msc_seg(P, 1, inputs, sweep.scales, flavour, crop_size=64, label_size=(48, 80), aux_layer=aux)
        w2 = O.msc_seg(P, 2, inputs, sweep.scales, flavour, crop_size=64, label_size=(48, 80), aux_layer=aux)
    assert s1.shape == w1.shape and s2.shape == w2.shape
    assert rel_err(s1, w1) < 1e-3 and rel_err(s2, w2) < 1e-3


def test_sweep_crf_stage_matches_c_restatement_and_labels_agree():
    """logits (device) -> up-sample -> softmax -> GPU mean-field vs the same logits through the CPU restatement."""
    import torch.nn.functional as F
    from dupl_b200.eval_sweep import SegCrfSweep
    from oracle.densecrf_ref import DenseCRF as RefCRF
    m, P = _model(21)
    inputs = synth_images(1, 64, 96, seed=33)
    img = _u8_image(inputs)
    sweep = SegCrfSweep(m, flavour="voc")
    s1, _ = sweep.msc_seg(inputs.cuda())
    q = sweep.crf_prob(img.cuda(), s1 * 4.0)  # sharpen the random-init logits so that the CRF has something to move
    prob = F.softmax(F.interpolate((s1 * 4.0).cpu(), size=(64, 96), mode="bilinear", align_corners=False), dim=1)[0].numpy()
    want = RefCRF(10, 1, 1, 4, 121, 5)(img.numpy(), prob)
    assert np.abs(q.cpu().numpy() - want).max() < 1e-4
    assert (q.argmax(0).cpu().numpy() == want.argmax(0)).mean() > 0.999


def test_sweep_call_returns_device_label_maps():
    from dupl_b200.eval_sweep import SegCrfSweep
    m, _ = _model(21)
    inputs = synth_images(2, 64, 64, seed=35)
    cls = synth_cls_labels(2, 20, seed=35)
    imgs = [_u8_image(inputs[i:i + 1]).cuda() for i in range(2)]
    out = SegCrfSweep(m, flavour="coco", crop_size=64)(imgs, inputs.cuda(), cls.cuda(), branch=2)
    assert len(out["seg_pred"]) == len(out["crf_pred"]) == 2
    for a, b in zip(out["seg_pred"], out["crf_pred"]):
        assert a.shape == b.shape == (64, 64) and a.is_cuda and b.dtype == torch.int64
        assert int(a.max()) <= 20 and int(b.max()) <= 20
    assert out["cam_label"].shape == (2, 64, 64) and out["cam_label"].dtype == torch.int64


def test_validate_accumulates_device_side_confusion_matrices():
    """validate_siamase + crf_proc as one sharded sweep: scores equal utils/evaluate.py applied to the per-image outputs."""
    from dupl_b200.eval_sweep import SegCrfSweep
    from dupl_b200.utils import evaluate
    m, _ = _model(21)
    sweep = SegCrfSweep(m, flavour="coco", crop_size=64)
    samples, segs, crfs, gts = [], [], [], []
    g = torch.Generator().manual_seed(9)
    for i in range(3):
        x = synth_images(1, 64, 64, seed=40 + i)
        gt = torch.randint(0, 21, (64, 64), generator=g)
        gt[torch.rand(64, 64, generator=g) < 0.1] = 255
        samples.append((_u8_image(x), x, synth_cls_labels(1, 20, seed=40 + i), gt))
    got = sweep.validate(samples, 21)
    for img, x, cls, gt in samples:
        out = sweep([img.cuda()], x.cuda(), cls.cuda(), branch=1)
        segs.append(out["seg_pred"][0].cpu().numpy())
        crfs.append(out["crf_pred"][0].cpu().numpy())
        gts.append(gt.numpy())
    assert got["seg"]["miou"] == evaluate.scores(gts, segs, 21)["miou"]
    assert got["crf"]["pAcc"] == evaluate.scores(gts, crfs, 21)["pAcc"]
    assert set(got) == {"seg", "crf", "cam"}
