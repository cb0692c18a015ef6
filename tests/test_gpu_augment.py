"""GPU: augment_data_strong on the device (csrc/augment.cu) against Pillow itself — the exact pipeline of
utils/imutils.py:305-317 (ToPILImage -> the chosen operations -> ToTensor -> Normalize -> flip) — bit for bit on the uint8
stage and to the last float bit on the normalised output; and the operation draw consumes Python's `random` stream like the
reference's RandAugment."""
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pil_pipeline(denorm, ops_per_image, m):
    """imutils.augment_data_strong with the random draw replaced by the given operations (torchvision + Pillow on the host)."""
    import PIL.ImageEnhance
    import PIL.ImageOps
    from torchvision import transforms
    from oracle import randaug_ref as R
    to_pil, to_tensor = transforms.ToPILImage(), transforms.ToTensor()
    norm = transforms.Normalize((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))
    out = torch.empty_like(denorm)
    for i in range(denorm.shape[0]):
        img = to_pil(denorm[i])
        for k in ops_per_image[i]:
            v = R.magnitude(k, m)
            if k == 0:
                img = PIL.ImageOps.autocontrast(img)
            elif k == 1:
                img = PIL.ImageOps.equalize(img)
            elif k == 2:
                img = PIL.ImageOps.posterize(img, max(1, int(v)))
            else:
                img = {3: PIL.ImageEnhance.Color, 4: PIL.ImageEnhance.Contrast, 5: PIL.ImageEnhance.Brightness,
                       6: PIL.ImageEnhance.Sharpness}[k](img).enhance(v)
        out[i] = torch.flip(norm(to_tensor(img)), dims=[2])
    return out


@pytest.mark.parametrize("H,W,m", [(448, 448, 10), (61, 97, 10), (64, 48, 25)])
def test_device_randaugment_is_bit_exact_with_pillow(H, W, m):
    from helpers import synth_images
    from dupl_b200.utils import imutils
    from oracle import dupl_oracle as O
    B = 4
    denorm = O.denormalize_img2(synth_images(B, H, W, seed=H))            # what the script hands over (k/255 floats)
    rnd = random.Random(H + m)
    per_image = [rnd.choices(range(7), k=5) for _ in range(B)]
    per_image[0] = [6, 1, 0, 4, 3]                                          # every neighbourhood / histogram operation at least once
    per_image[1] = [2, 5, 6, 6, 1]
    ops = [[per_image[b][s] for b in range(B)] for s in range(5)]
    got = imutils.augment_data_strong(denorm.cuda(), n=5, m=m, ops=ops).cpu()
    want = _pil_pipeline(denorm, per_image, m)
    assert torch.equal(got, want), int((got != want).sum())
    # no operations: ToPILImage -> ToTensor -> Normalize -> flip alone
    none = imutils.augment_data_strong(denorm.cuda(), n=0, m=m, ops=torch.zeros(0, B, dtype=torch.int32)).cpu()
    assert torch.equal(none, _pil_pipeline(denorm, [[]] * B, m))


def test_operation_draw_follows_the_reference_random_stream():
    from dupl_b200.utils import imutils
    random.seed(7)
    ours = imutils.draw_ops(3, 5)
    random.seed(7)
    lst = list(range(7))
    ref = [random.choices(lst, k=5) for _ in range(3)]                      # RandAugment.__call__ per image (randomaug.py:260)
    assert ours == [[ref[b][s] for b in range(3)] for s in range(5)]
