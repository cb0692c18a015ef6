"""CPU: the DenseCRF C restatement (oracle/densecrf_ref.c).  pydensecrf is absent, so the only checks
available are internal: the permutohedral-lattice mean-field must stay close to the exact O(N^2)
dense-Gaussian mean-field with the same kernels, normalisation and update rule on small images."""
import numpy as np
import pytest

from oracle.densecrf_ref import DenseCRF


def _tiny(seed=0, H=24, W=20, C=4):
    rng = np.random.RandomState(seed)
    img = np.zeros((H, W, 3), np.uint8)
    img[:, :W // 2] = [200, 50, 50]
    img[:, W // 2:] = [40, 180, 90]
    img[H // 2:, :, 2] += 60
    img = (img.astype(int) + rng.randint(-8, 8, img.shape)).clip(0, 255).astype(np.uint8)
    logits = rng.randn(C, H, W).astype(np.float32)
    logits[0, :, :W // 2] += 1.5
    logits[1, :, W // 2:] += 1.5
    p = np.exp(logits)
    p /= p.sum(0, keepdims=True)
    return img, p.astype(np.float32)


@pytest.mark.parametrize("params", [(10, 3, 3, 4, 20, 13), (10, 1, 1, 4, 121, 5), (5, 3, 3, 0, 20, 13), (5, 0, 3, 5, 8, 10)])
def test_lattice_mean_field_tracks_exact_dense_mean_field(params):
    img, p = _tiny()
    crf = DenseCRF(*params)
    q, qb = crf(img, p), crf.bruteforce(img, p)
    assert np.abs(q.sum(0) - 1).max() < 1e-5
    assert np.abs(q - qb).max() < 0.05                      # lattice approximation error
    assert (q.argmax(0) == qb.argmax(0)).mean() > 0.98
    assert (q.argmax(0) != p.argmax(0)).mean() > 0.05       # the CRF actually changes labels


def test_zero_iterations_returns_the_softmax_of_the_unary():
    img, p = _tiny(seed=1)
    q = DenseCRF(0, 1, 1, 4, 121, 5)(img, p)
    assert np.abs(q - p).max() < 1e-5
