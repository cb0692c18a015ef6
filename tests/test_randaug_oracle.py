"""CPU: oracle/randaug_ref.py (numpy restatement of the seven Pillow operations of utils/randomaug.py:161-262) pinned against
Pillow itself, bit for bit — per operation and through random 5-operation chains at the reference's magnitude (n=5, m=10)."""
import random

import numpy as np
import PIL.ImageEnhance
import PIL.ImageOps
import pytest
from PIL import Image

from oracle import randaug_ref as R


def _pil_op(img, k, v):
    if k == 0:
        return PIL.ImageOps.autocontrast(img)
    if k == 1:
        return PIL.ImageOps.equalize(img)
    if k == 2:
        return PIL.ImageOps.posterize(img, max(1, int(v)))
    cls = {3: PIL.ImageEnhance.Color, 4: PIL.ImageEnhance.Contrast, 5: PIL.ImageEnhance.Brightness, 6: PIL.ImageEnhance.Sharpness}[k]
    return cls(img).enhance(v)


def _images():
    rng = np.random.RandomState(0)
    smooth = np.asarray(Image.fromarray(rng.randint(0, 256, (12, 16, 3)).astype(np.uint8)).resize((97, 61), Image.BILINEAR))
    yield smooth
    yield rng.randint(0, 256, (40, 33, 3)).astype(np.uint8)
    yield (rng.randint(60, 120, (31, 45, 3))).astype(np.uint8)                 # narrow histogram
    flat = np.full((20, 20, 3), 77, np.uint8)
    flat[3:9, 4:11, 1] = 200
    yield flat                                                                # nearly constant bands (degenerate LUT branches)


@pytest.mark.parametrize("k", range(7))
def test_each_operation_is_bit_exact_against_pillow(k):
    for m in (10, 3, 27):
        v = R.magnitude(k, m)
        for a in _images():
            want = np.asarray(_pil_op(Image.fromarray(a), k, v))
            got = R.apply_op(a, k, v)
            assert np.array_equal(got, want), (R.OPS[k], m, int((got != want).sum()))


def test_random_chains_at_the_reference_magnitude():
    rnd = random.Random(0)
    for a in _images():
        for _ in range(6):
            ops = rnd.choices(range(7), k=5)
            img = Image.fromarray(a)
            for k in ops:
                img = _pil_op(img, k, R.magnitude(k, 10))
            assert np.array_equal(R.augment_u8(a, ops, 10), np.asarray(img)), ops
