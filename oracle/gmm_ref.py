"""CPU restatement (numpy, fp64) of the GMM noise filter of the training loop (TEST INFRASTRUCTURE — never the product).

train_final_voc.py:358-394 (same block in train_final_coco.py): per image and student, the per-pixel CE losses of the
foreground pseudo-labels that exceed 0.1 are fitted with `sklearn.mixture.GaussianMixture(n_components=2, max_iter=10, tol=1e-2,
reg_covar=5e-4)`; if the two means are further apart than 1.0, every labelled pixel whose posterior under the high-loss
component exceeds 0.95 is set to the ignore label.

The algorithm lives in scikit-learn (third-party; the reference pins 1.0.2, this image has 1.9): k-means initialisation ->
one-hot responsibilities -> M-step, then E/M steps until |d lower_bound| < tol or max_iter, posterior = softmax of the weighted
log-densities.  The only part that is NOT restated literally is the k-means++ seeding (NumPy RandomState): 1-D two-means is
run here as deterministic Lloyd iterations from mean -/+ std — for two clusters on a line it converges to the same partition.
PINNED against scikit-learn itself by tests/test_gmm_oracle.py on well-separated AND overlapping mixtures (mask mismatch
<= 1e-4, i.e. a pixel or two whose posterior sits within fp32 rounding of 0.95: sklearn computes in float32); the CUDA kernel
(dupl_b200/csrc/gmm.cu) is compared with this restatement and with scikit-learn on the GPU box.
"""
import numpy as np

EPS10 = 10 * np.finfo(np.float32).eps       # sklearn: nk = resp.sum(0) + 10 * eps of the DATA dtype (float32 losses)
LOG2PI = np.log(2 * np.pi)


def _log_prob(x, w, mu, var):
    """weighted log-densities [n, 2] and their log-sum-exp [n]  (sklearn _estimate_weighted_log_prob / _estimate_log_prob_resp)."""
    lw = np.log(w) - 0.5 * (LOG2PI + np.log(var))
    l = lw[None, :] - 0.5 * (x[:, None] - mu[None, :]) ** 2 / var[None, :]
    m = l.max(1)
    return l, m + np.log(np.exp(l - m[:, None]).sum(1))


def fit_two_gaussians(x, reg_covar=5e-4, max_iter=10, tol=1e-2):
    """x: 1-D float64 samples -> (weights [2], means [2], variances [2], EM iterations)."""
    n = x.size
    mean = x.mean()
    sd = np.sqrt(max((x * x).mean() - mean * mean, 0.0))
    c0, c1 = mean - sd, mean + sd
    for _ in range(100):                                   # Lloyd: the k-means initialisation of GaussianMixture
        lo = x < 0.5 * (c0 + c1)
        m0 = x[lo].mean() if lo.any() else c0
        m1 = x[~lo].mean() if (~lo).any() else c1
        done = abs(m0 - c0) + abs(m1 - c1) < 1e-7 * (abs(mean) + 1e-12)
        c0, c1 = m0, m1
        if done:
            break
    lo = x < 0.5 * (c0 + c1)                                # one-hot responsibilities -> first M-step
    nk = np.array([lo.sum() + EPS10, (~lo).sum() + EPS10])
    mu = np.array([x[lo].sum() / nk[0], x[~lo].sum() / nk[1]])
    var = np.array([((x[lo] - mu[0]) ** 2).sum() / nk[0], ((x[~lo] - mu[1]) ** 2).sum() / nk[1]]) + reg_covar
    w = nk / n
    lower, iters = -np.inf, 0
    for it in range(1, max_iter + 1):
        l, lse = _log_prob(x, w, mu, var)
        r0 = np.exp(l[:, 0] - lse)
        nk = np.array([r0.sum() + EPS10, (n - r0.sum()) + EPS10])
        nm = np.array([(r0 * x).sum() / nk[0], ((1 - r0) * x).sum() / nk[1]])
        var = np.array([(r0 * (x - nm[0]) ** 2).sum() / nk[0], ((1 - r0) * (x - nm[1]) ** 2).sum() / nk[1]]) + reg_covar
        mu = nm
        w = nk / n
        w = w / w.sum()
        new_lower = lse.sum() / n                           # lower bound of the parameters the E-step used
        change, lower, iters = new_lower - lower, new_lower, it
        if abs(change) < tol:
            break
    return w, mu, var, iters


def gmm_noise_filter(loss, label, ignore_index=255.0, loss_min=0.1, min_count=1000, gmm_valid_thre=1.0, gamma=0.95,
                     reg_covar=5e-4, max_iter=10, tol=1e-2):
    """loss, label: float32 [b, H, W] -> (filtered label copy, info [b, 4] = samples, applied?, EM iterations, pixels flipped)."""
    label = np.array(label, dtype=np.float32, copy=True)
    info = np.zeros((loss.shape[0], 4), dtype=np.int64)
    for i in range(loss.shape[0]):
        L = loss[i].astype(np.float64).ravel()
        lab = label[i].reshape(-1)
        sel = (lab != 0) & (lab != ignore_index) & (loss[i].ravel() > loss_min)
        x = L[sel]
        info[i, 0] = x.size
        if x.size <= min_count:                             # train_final_voc.py:366
            continue
        w, mu, var, iters = fit_two_gaussians(x, reg_covar, max_iter, tol)
        info[i, 2] = iters
        if not abs(mu[0] - mu[1]) > gmm_valid_thre:         # :372
            continue
        hi = 1 if mu[1] > mu[0] else 0
        l, lse = _log_prob(L, w, mu, var)                   # predict_proba on EVERY pixel, :378
        noise = (np.exp(l[:, hi] - lse) > gamma) & (lab != 0)
        info[i, 1] = 1
        info[i, 3] = int((noise & (lab != ignore_index)).sum())
        lab[noise] = ignore_index
    return label, info
