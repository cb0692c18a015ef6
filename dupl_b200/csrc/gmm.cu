// GMM noise filter of the training loop (train_final_voc.py:358-394): per image, a 2-component
// 1-D Gaussian mixture is fitted to the per-pixel CE losses of the foreground pseudo-labels
// (sklearn.mixture.GaussianMixture(2, max_iter=10, tol=1e-2, reg_covar=5e-4, init k-means) in the
// reference, a GPU -> CPU -> GPU round trip per image and student); pixels whose posterior under the
// high-loss component exceeds gamma are set to the ignore label.
//
// PARITY UNPINNED against scikit-learn (third-party, k-means++ seeded by NumPy's RandomState): the
// k-means initialisation is replaced by deterministic Lloyd iterations from (mean -/+ std); 1-D two-means
// has few fixed points, and the result is checked against sklearn on bimodal data in tests/ (mask
// agreement), not bit for bit.  Everything else follows sklearn's EM: one-hot responsibilities ->
// weights / means / variances (+ reg_covar), E-step / M-step until |d lower_bound| < tol or max_iter,
// posterior = softmax of the weighted log-densities.
// One thread block per image; no host synchronisation, the "valid" decisions stay on the device.
#include "common.cuh"

namespace dupl {

constexpr int GMM_THREADS = 1024;

struct GmmArgsDev {
  const float* loss;   // [b, n]
  float* label;        // [b, n] in {0..K, ignore}, modified in place
  int n;
  float ignore, loss_min, valid_gap, gamma, reg_covar, tol;
  int min_count, max_iter;
  int* info;           // [b, 4]: samples, filtered?, EM iterations, pixels ignored
};

__device__ double gmm_block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < GMM_THREADS / 32; ++w) r += sh[w];  // every thread sums in the same order
  return r;
}

__global__ void __launch_bounds__(GMM_THREADS) gmm_filter_kernel(GmmArgsDev a) {
  __shared__ double sh[GMM_THREADS / 32];
  const int img = blockIdx.x;
  const float* L = a.loss + static_cast<long>(img) * a.n;
  float* lab = a.label + static_cast<long>(img) * a.n;
  auto selected = [&](int i) -> bool {
    const float l = lab[i];
    return l != 0.0f && l != a.ignore && L[i] > a.loss_min;
  };
  // ---- sample count, mean, std
  double c = 0, s = 0, q = 0;
  for (int i = threadIdx.x; i < a.n; i += GMM_THREADS)
    if (selected(i)) {
      const double x = L[i];
      c += 1; s += x; q += x * x;
    }
  const double n = gmm_block_sum(c, sh);
  const double sum = gmm_block_sum(s, sh), sq = gmm_block_sum(q, sh);
  if (threadIdx.x == 0) {
    a.info[4 * img + 0] = static_cast<int>(n);
    a.info[4 * img + 1] = 0; a.info[4 * img + 2] = 0; a.info[4 * img + 3] = 0;
  }
  if (n <= a.min_count) return;                       // (seg_loss_m > 0.1).sum() > 1000
  const double mean = sum / n;
  const double sd = sqrt(fmax(sq / n - mean * mean, 0.0));
  // ---- deterministic 1-D 2-means (Lloyd) from mean -/+ std
  double c0 = mean - sd, c1 = mean + sd;
  for (int it = 0; it < 100; ++it) {
    const double thr = 0.5 * (c0 + c1);
    double n0 = 0, s0 = 0, s1 = 0;
    for (int i = threadIdx.x; i < a.n; i += GMM_THREADS)
      if (selected(i)) {
        const double x = L[i];
        if (x < thr) { n0 += 1; s0 += x; } else { s1 += x; }
      }
    n0 = gmm_block_sum(n0, sh); s0 = gmm_block_sum(s0, sh); s1 = gmm_block_sum(s1, sh);
    const double n1 = n - n0;
    const double m0 = n0 > 0 ? s0 / n0 : c0, m1 = n1 > 0 ? s1 / n1 : c1;
    const bool done = fabs(m0 - c0) + fabs(m1 - c1) < 1e-7 * (fabs(mean) + 1e-12);
    c0 = m0; c1 = m1;
    if (done) break;
  }
  // ---- initial parameters from the hard assignment (sklearn _initialize_parameters)
  const double eps10 = 10.0 * 1.1920928955078125e-07;  // 10 * np.finfo(float32).eps
  double w[2], mu[2], var[2];
  {
    const double thr = 0.5 * (c0 + c1);
    double n0 = 0, s0 = 0, s1 = 0;
    for (int i = threadIdx.x; i < a.n; i += GMM_THREADS)
      if (selected(i)) {
        const double x = L[i];
        if (x < thr) { n0 += 1; s0 += x; } else { s1 += x; }
      }
    n0 = gmm_block_sum(n0, sh); s0 = gmm_block_sum(s0, sh); s1 = gmm_block_sum(s1, sh);
    const double nk0 = n0 + eps10, nk1 = (n - n0) + eps10;
    mu[0] = s0 / nk0; mu[1] = s1 / nk1;
    double v0 = 0, v1 = 0;
    for (int i = threadIdx.x; i < a.n; i += GMM_THREADS)
      if (selected(i)) {
        const double x = L[i];
        if (x < thr) v0 += (x - mu[0]) * (x - mu[0]); else v1 += (x - mu[1]) * (x - mu[1]);
      }
    v0 = gmm_block_sum(v0, sh); v1 = gmm_block_sum(v1, sh);
    var[0] = v0 / nk0 + a.reg_covar; var[1] = v1 / nk1 + a.reg_covar;
    w[0] = nk0 / n; w[1] = nk1 / n;
  }
  // ---- EM
  const double LOG2PI = 1.8378770664093453;
  double lower = -INFINITY;
  int iters = 0;
  for (int it = 1; it <= a.max_iter; ++it) {
    const double lw0 = log(w[0]) - 0.5 * (LOG2PI + log(var[0])), lw1 = log(w[1]) - 0.5 * (LOG2PI + log(var[1]));
    double r0s = 0, r0x = 0, r1x = 0, ll = 0;
    for (int i = threadIdx.x; i < a.n; i += GMM_THREADS)
      if (selected(i)) {
        const double x = L[i];
        const double l0 = lw0 - 0.5 * (x - mu[0]) * (x - mu[0]) / var[0], l1 = lw1 - 0.5 * (x - mu[1]) * (x - mu[1]) / var[1];
        const double m = fmax(l0, l1), lse = m + log(exp(l0 - m) + exp(l1 - m));
        const double r0 = exp(l0 - lse);
        ll += lse; r0s += r0; r0x += r0 * x; r1x += (1.0 - r0) * x;
      }
    r0s = gmm_block_sum(r0s, sh); r0x = gmm_block_sum(r0x, sh); r1x = gmm_block_sum(r1x, sh); ll = gmm_block_sum(ll, sh);
    const double nk0 = r0s + eps10, nk1 = (n - r0s) + eps10;
    const double nm0 = r0x / nk0, nm1 = r1x / nk1;
    double v0 = 0, v1 = 0;
    for (int i = threadIdx.x; i < a.n; i += GMM_THREADS)
      if (selected(i)) {
        const double x = L[i];
        const double l0 = lw0 - 0.5 * (x - mu[0]) * (x - mu[0]) / var[0], l1 = lw1 - 0.5 * (x - mu[1]) * (x - mu[1]) / var[1];
        const double m = fmax(l0, l1), lse = m + log(exp(l0 - m) + exp(l1 - m));
        const double r0 = exp(l0 - lse);
        v0 += r0 * (x - nm0) * (x - nm0); v1 += (1.0 - r0) * (x - nm1) * (x - nm1);
      }
    v0 = gmm_block_sum(v0, sh); v1 = gmm_block_sum(v1, sh);
    mu[0] = nm0; mu[1] = nm1;
    var[0] = v0 / nk0 + a.reg_covar; var[1] = v1 / nk1 + a.reg_covar;
    w[0] = nk0 / n; w[1] = nk1 / n;
    const double ws = w[0] + w[1];
    w[0] /= ws; w[1] /= ws;
    const double new_lower = ll / n;
    const double change = new_lower - lower;
    lower = new_lower;
    iters = it;
    if (fabs(change) < a.tol) break;
  }
  if (threadIdx.x == 0) a.info[4 * img + 2] = iters;
  if (!(fabs(mu[0] - mu[1]) > a.valid_gap)) return;   // "two normal distributions"
  // ---- posterior of the high-mean component on EVERY pixel; noisy & labelled -> ignore
  const int hi = mu[1] > mu[0] ? 1 : 0;
  const double lw0 = log(w[0]) - 0.5 * (LOG2PI + log(var[0])), lw1 = log(w[1]) - 0.5 * (LOG2PI + log(var[1]));
  double flipped = 0;
  for (int i = threadIdx.x; i < a.n; i += GMM_THREADS) {
    const double x = L[i];
    const double l0 = lw0 - 0.5 * (x - mu[0]) * (x - mu[0]) / var[0], l1 = lw1 - 0.5 * (x - mu[1]) * (x - mu[1]) / var[1];
    const double m = fmax(l0, l1), lse = m + log(exp(l0 - m) + exp(l1 - m));
    const double p = exp((hi ? l1 : l0) - lse);
    if (p > a.gamma && lab[i] != 0.0f) {
      if (lab[i] != a.ignore) flipped += 1;
      lab[i] = a.ignore;
    }
  }
  flipped = gmm_block_sum(flipped, sh);
  if (threadIdx.x == 0) {
    a.info[4 * img + 1] = 1;
    a.info[4 * img + 3] = static_cast<int>(flipped);
  }
}

}  // namespace dupl

extern "C" int dupl_gmm_filter(const float* loss, float* label, int32_t b, int32_t n, float ignore_index, float loss_min,
                               int32_t min_count, float valid_gap, float gamma, float reg_covar, int32_t max_iter, float tol,
                               int32_t* info, void* stream) {
  using namespace dupl;
  DUPL_CHECK_ARG(loss && label && info && b > 0 && n > 0 && max_iter > 0, "dupl_gmm_filter: bad arguments");
  GmmArgsDev a;
  a.loss = loss; a.label = label; a.n = n; a.ignore = ignore_index; a.loss_min = loss_min; a.valid_gap = valid_gap;
  a.gamma = gamma; a.reg_covar = reg_covar; a.tol = tol; a.min_count = min_count; a.max_iter = max_iter; a.info = info;
  gmm_filter_kernel<<<b, GMM_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(a);
  DUPL_LAUNCH_OK();
  return DUPL_OK;
}
