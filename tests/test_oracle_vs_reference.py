"""Authoring-container only: the oracle against the live reference on fresh seeded inputs (skipped when
/root/reference is absent, e.g. on the GPU box)."""
import pytest
import torch

from oracle import dupl_oracle as O
from oracle import ref_import

pytestmark = [pytest.mark.reference, pytest.mark.skipif(not ref_import.available(), reason="reference tree not mounted")]


def test_bilinear_is_bit_exact_for_the_upsampling_ratios_on_the_path():
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(0)
    for hi, ho in [(8, 128), (14, 448), (28, 448), (42, 448), (224, 448), (448, 224), (21, 448)]:
        t = torch.randn(1, 2, hi, hi, generator=g)
        assert torch.equal(F.interpolate(t, size=(ho, ho), mode="bilinear", align_corners=False), O.bilinear(t, ho, ho))


def test_par_matches_reference_module():
    ref = ref_import.load()
    par = ref.PAR.PAR(num_iter=10, dilations=[1, 2, 4, 8, 12, 24])
    g = torch.Generator().manual_seed(1)
    img = torch.randint(0, 256, (2, 3, 48, 40), generator=g).float() / 255
    mk = torch.rand(2, 4, 48, 40, generator=g).softmax(1)
    assert (par(img, mk) - O.par_forward(img, mk)).abs().max().item() < 1e-6


def test_refine_matches_reference_bit_exact():
    ref = ref_import.load()
    par = ref.PAR.PAR(num_iter=10, dilations=[1, 2, 4, 8, 12, 24])
    g = torch.Generator().manual_seed(2)
    imgs = torch.randint(0, 256, (2, 3, 64, 64), generator=g).float() / 255
    cls = torch.zeros(2, 20)
    cls[0, [2, 5]] = 1
    cls[1, [7]] = 1
    cams = torch.rand(2, 20, 64, 64, generator=g) * cls[:, :, None, None]
    box = torch.tensor([[0, 64, 0, 64], [5, 60, 3, 50]], dtype=torch.int16)
    want = ref.cam_helper.refine_cams_with_bkg_v2(par, imgs, cams, cls, high_thre=0.65, low_thre=0.25, ignore_index=255, img_box=box)
    assert torch.equal(O.refine_cams(imgs, cams, cls, 0.65, 0.25, 255, box), want)
