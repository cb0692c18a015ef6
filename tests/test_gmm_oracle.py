"""CPU: the numpy restatement of the GMM noise filter (oracle/gmm_ref.py) pinned against scikit-learn's GaussianMixture driven
exactly like train_final_voc.py:358-394.  Bar: mask mismatch <= 1e-4, on well-separated and on overlapping
mixtures; scikit-learn's own seed-to-seed spread on the same fixtures is measured beside it (0 .. 1 pixel): the k-means++
seeding that the restatement replaces by deterministic Lloyd iterations is not what decides a pixel."""
import numpy as np
import pytest

from oracle import gmm_ref

sklearn_mixture = pytest.importorskip("sklearn.mixture")


def sklearn_filter(loss, label, ignore=255, gamma=0.95, valid=1.0, random_state=0):
    label = label.copy()
    for i in range(loss.shape[0]):
        roi = (label[i] != 0) & (label[i] != ignore)
        m = loss[i][roi]
        if (m > 0.1).sum() > 1000:
            gmm = sklearn_mixture.GaussianMixture(n_components=2, max_iter=10, tol=1e-2, reg_covar=5e-4, random_state=random_state)
            gmm.fit(m[m > 0.1].reshape(-1, 1))
            if abs(gmm.means_[0, 0] - gmm.means_[1, 0]) > valid:
                k = gmm.means_.argmax()
                prob = gmm.predict_proba(loss[i].reshape(-1, 1))
                noise = (prob[:, k] > gamma).reshape(label[i].shape) & (label[i] != 0)
                label[i][noise] = ignore
    return label


def mixture_case(seed, mu_noisy=3.5, sd_noisy=0.6, frac=0.25, H=96, W=128):
    rng = np.random.RandomState(seed)
    b = 3
    label = rng.choice([0, 3, 7, 255], size=(b, H, W), p=[0.4, 0.3, 0.2, 0.1]).astype(np.float32)
    clean = rng.gamma(2.0, 0.15, size=(b, H, W))
    noisy = rng.normal(mu_noisy, sd_noisy, size=(b, H, W)).clip(0.2)
    loss = np.where(rng.rand(b, H, W) < frac, noisy, clean).astype(np.float32)
    loss[label == 255] = 0.0
    return loss, label


CASES = [(3.5, 0.6, 0.25), (2.0, 0.6, 0.25), (1.5, 0.5, 0.3), (2.5, 1.0, 0.1), (3.5, 0.6, 0.02), (1.2, 0.4, 0.5)]


@pytest.mark.parametrize("mu,sd,frac", CASES)
def test_restatement_matches_sklearn_masks(mu, sd, frac):
    worst = 0.0
    for seed in range(3):
        loss, label = mixture_case(seed, mu, sd, frac)
        want = sklearn_filter(loss, label)
        got, info = gmm_ref.gmm_noise_filter(loss, label)
        worst = max(worst, float((got != want).mean()))
        assert (info[:, 1] == 1).all() == bool((want != label).any())
    assert worst <= 1e-4, worst


def test_sklearn_own_seed_spread_is_the_same_band():
    """scikit-learn against itself with another random_state: 0 .. 1 pixel of 36 864 (<= 3e-5) on these fixtures — the band
    the restatement sits in (<= 2 pixels); the k-means++ seeding is not what decides a pixel."""
    worst = 0.0
    for mu, sd, frac in CASES[:4]:
        loss, label = mixture_case(1, mu, sd, frac)
        ref = sklearn_filter(loss, label, random_state=0)
        for rs in (1, 2, 3):
            worst = max(worst, float((sklearn_filter(loss, label, random_state=rs) != ref).mean()))
    assert worst <= 1e-4, worst
