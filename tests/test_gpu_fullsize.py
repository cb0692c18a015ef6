"""GPU: size-independent properties at BASELINE.json's FULL sizes (b=4, 448x448, K=20; CRF at 640x480x81), where the CPU
oracle would take minutes: flip equivariance of the multi-scale CAM, mass conservation of the PAR propagation, label
alphabet / box semantics of the refine step, normalisation of the mean-field marginals."""
import numpy as np
import pytest
import torch

from helpers import init_state_dict, synth_boxes, synth_cls_labels, synth_images

pytestmark = pytest.mark.gpu
B, S, K = 4, 448, 20


@pytest.fixture(scope="module")
def model():
    from dupl_b200.model.model_dupl import siamese_network
    m = siamese_network("deit_base_patch16_224", num_classes=K + 1, pretrained=False, aux_layer=-3)
    m.load_state_dict(init_state_dict(K + 1), strict=True)
    return m.cuda().eval()


def test_mscam_is_flip_equivariant_and_normalised(model):
    """multi_scale_cam2_siamese fuses f(x) and flip(f(flip x)) by a max (cam_helper.py:187-194): cam(flip x) == flip(cam x)."""
    from dupl_b200.utils import cam_helper
    x = synth_images(B, S, S, seed=3).cuda()
    cam, aux = cam_helper.multi_scale_cam2_siamese(model, x, (1.0, 0.5, 1.5), branch=1)
    cam_f, aux_f = cam_helper.multi_scale_cam2_siamese(model, x.flip(-1), (1.0, 0.5, 1.5), branch=1)
    assert cam.shape == (B, K, S, S)
    for a, b_ in ((cam, cam_f), (aux, aux_f)):
        assert torch.isfinite(a).all()
        assert float(a.min()) >= 0.0 and float(a.max()) < 1.0
        assert float(a.amin(dim=(2, 3)).max()) == 0.0                      # cam += max(-cam): every plane touches 0
        # well-conditioned planes only (a plane that ReLU leaves almost constant is divided by ~1e-5)
        diff = (a - b_.flip(-1)).abs().amax(dim=(2, 3))
        assert float(diff.median()) < 1e-3


def test_par_propagation_conserves_constants_at_full_size():
    """Every affinity row sums to 1 + 0.01 (PAR.py:84-87), so a constant mask c becomes c * 1.01^10 after 10 iterations."""
    from dupl_b200 import ops
    from oracle import dupl_oracle as O
    x = synth_images(B, S, S, seed=5)
    imgs = torch.nn.functional.avg_pool2d(O.denormalize_img2(x), 2).cuda()   # the 224x224 images the refine step hands to PAR
    dil = [1, 2, 4, 8, 12, 24]
    aff = ops.par_affinity(imgs, dil)
    assert aff.shape == (B, 48, 224, 224)
    assert float((aff.sum(1) - 1.01).abs().max()) < 1e-5
    masks = torch.full((B, 6, 224, 224), 0.25, device="cuda")
    masks[:, 1] = 0.5
    out = ops.par_propagate(aff, masks.clone(), dil, 10)
    assert float((out[:, 0] - 0.25 * 1.01 ** 10).abs().max()) < 1e-5
    assert float((out[:, 1] - 0.5 * 1.01 ** 10).abs().max()) < 1e-5


def test_refine_labels_alphabet_and_box_at_full_size(model):
    from dupl_b200.model.PAR import PAR
    from dupl_b200.utils import cam_helper
    from oracle import dupl_oracle as O
    x = synth_images(B, S, S, seed=7)
    cls = synth_cls_labels(B, K, seed=7)
    box = synth_boxes(B, S, S, seed=7)
    par = PAR(num_iter=10, dilations=[1, 2, 4, 8, 12, 24]).cuda()
    cam, _ = cam_helper.multi_scale_cam2_siamese(model, x.cuda(), (1.0, 0.5, 1.5), branch=2)
    lab = cam_helper.refine_cams_with_bkg_v2(par, O.denormalize_img2(x).cuda(), cam * cls.cuda()[:, :, None, None], cls.cuda(),
                                             high_thre=0.65, low_thre=0.25, ignore_index=255, img_box=box).cpu()
    assert lab.shape == (B, S, S) and lab.dtype == torch.float32
    for i in range(B):
        allowed = {0.0, 255.0} | {float(k + 1) for k in torch.nonzero(cls[i]).flatten().tolist()}
        assert set(lab[i].unique().tolist()) <= allowed
        y0, y1, x0, x1 = box[i].tolist()
        outside = torch.ones(S, S, dtype=torch.bool)
        outside[y0:y1, x0:x1] = False
        assert (lab[i][outside] == 255).all()


def test_crf_marginals_are_normalised_at_coco_size():
    from dupl_b200.utils.dcrf import DenseCRF
    rng = np.random.RandomState(0)
    H, W, C = 480, 640, 81
    img = torch.from_numpy(rng.randint(0, 256, (H, W, 3)).astype(np.uint8)).cuda()
    logits = torch.from_numpy(rng.randn(1, C, H // 16, W // 16).astype(np.float32)).cuda() * 3
    prob = torch.softmax(torch.nn.functional.interpolate(logits, size=(H, W), mode="bilinear", align_corners=False), 1)[0]
    q = DenseCRF(10, 1, 1, 4, 121, 5)(img, prob)
    assert q.shape == (C, H, W) and torch.isfinite(q).all()
    assert float((q.sum(0) - 1.0).abs().max()) < 1e-4
    q2 = DenseCRF(10, 1, 1, 4, 121, 5)(img, prob)
    assert torch.equal(q, q2)                                              # fixed-point splat: bit-reproducible
    assert float((q.argmax(0) == prob.argmax(0)).float().mean()) > 0.5     # smoothing, not scrambling
