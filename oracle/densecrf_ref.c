/* CPU oracle for DenseCRF mean-field inference (TEST INFRASTRUCTURE — never the product).
 *
 * PARITY UNPINNED: the reference calls pydensecrf (utils/dcrf.py:1-3, 42-69), an un-vendored
 * third-party package (git master of lucasb-eyer/pydensecrf wrapping Krähenbühl's densecrf C++),
 * absent from /root/reference and not installable here.  This file restates the PUBLISHED
 * algorithm (Krähenbühl & Koltun, NIPS 2011; Adams et al., permutohedral lattice, 2010) as
 * exposed through the reference's call sites:
 *   DenseCRF.__call__                utils/dcrf.py:51-69
 *     U = -log(clip(p, 1e-5, 1))     (unary_from_softmax)
 *     DenseCRF2D(W,H,C); setUnaryEnergy(U)
 *     addPairwiseGaussian(sxy, compat=pos_w)              features (x/sxy, y/sxy)
 *     addPairwiseBilateral(sxy, srgb, rgbim, compat=bi_w) features (x/sxy, y/sxy, r/srgb, g/srgb, b/srgb)
 *     inference(iter_max)
 *   kernel DIAG, normalisation SYMMETRIC: norm = 1/sqrt(K 1 + 1e-20), K~Q = norm .* lattice(norm .* Q)
 *   Potts compatibility: message = -w K~Q;  Q <- softmax(-U + sum_k w_k K~_k Q)
 * Lattice filter = splat (barycentric weights) -> blur along each of the d+1 axes
 * (v + 0.5 (v[n-] + v[n+])) -> slice, scaled by alpha = 1/(1 + 2^-d).
 *
 * Build: gcc -O2 -fPIC -shared (oracle/Makefile) -> oracle/libdensecrf_ref.so, used by tests/ and by
 * bench_crf's cpu_baseline leg only.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXD 5

typedef struct {
  int d, N, M;
  int* offset;      /* [N*(d+1)] lattice vertex of each (pixel, remainder) */
  float* bary;      /* [N*(d+1)] */
  int* n1;          /* [(d+1)*M] */
  int* n2;
} Lattice;

/* ---- open-addressing hash table over d-short keys, ids in insertion order ---- */
typedef struct {
  int d, cap, filled;
  short* keys; /* [cap_keys * d] in id order */
  int* table;  /* [cap] -> id or -1 */
  int key_cap;
} Hash;

static unsigned hash_key(const short* k, int d) {
  unsigned r = 0;
  for (int i = 0; i < d; ++i) {
    r += (unsigned)(int)k[i];
    r *= 1664525u;
  }
  return r;
}

static void hash_init(Hash* h, int d, int n_elements) {
  h->d = d;
  h->cap = 1;
  while (h->cap < 2 * n_elements) h->cap <<= 1;
  h->filled = 0;
  h->key_cap = n_elements;
  h->keys = (short*)malloc(sizeof(short) * (size_t)n_elements * d);
  h->table = (int*)malloc(sizeof(int) * (size_t)h->cap);
  for (int i = 0; i < h->cap; ++i) h->table[i] = -1;
}

static int hash_find(Hash* h, const short* k, int create) {
  unsigned p = hash_key(k, h->d) & (unsigned)(h->cap - 1);
  for (;;) {
    int e = h->table[p];
    if (e == -1) {
      if (!create) return -1;
      memcpy(h->keys + (size_t)h->filled * h->d, k, sizeof(short) * h->d);
      h->table[p] = h->filled;
      return h->filled++;
    }
    if (memcmp(h->keys + (size_t)e * h->d, k, sizeof(short) * h->d) == 0) return e;
    p = (p + 1) & (unsigned)(h->cap - 1);
  }
}

static void hash_free(Hash* h) {
  free(h->keys);
  free(h->table);
}

static void lattice_init(Lattice* L, const float* feature /* [N][d] */, int d, int N) {
  L->d = d;
  L->N = N;
  L->offset = (int*)malloc(sizeof(int) * (size_t)N * (d + 1));
  L->bary = (float*)malloc(sizeof(float) * (size_t)N * (d + 1));
  Hash H;
  hash_init(&H, d, N * (d + 1));

  int canonical[(MAXD + 1) * (MAXD + 1)];
  for (int i = 0; i <= d; ++i) {
    for (int j = 0; j <= d - i; ++j) canonical[i * (d + 1) + j] = i;
    for (int j = d - i + 1; j <= d; ++j) canonical[i * (d + 1) + j] = i - (d + 1);
  }
  float scale_factor[MAXD];
  const float inv_std_dev = sqrtf(2.0f / 3.0f) * (float)(d + 1);
  for (int i = 0; i < d; ++i) scale_factor[i] = 1.0f / sqrtf((float)((i + 2) * (i + 1))) * inv_std_dev;

  for (int k = 0; k < N; ++k) {
    const float* f = feature + (size_t)k * d;
    float elevated[MAXD + 1], rem0[MAXD + 1], barycentric[MAXD + 2];
    int rank[MAXD + 1];
    short key[MAXD + 1];
    float sm = 0.0f;
    for (int j = d; j > 0; --j) {
      const float cf = f[j - 1] * scale_factor[j - 1];
      elevated[j] = sm - (float)j * cf;
      sm += cf;
    }
    elevated[0] = sm;

    const float down_factor = 1.0f / (float)(d + 1), up_factor = (float)(d + 1);
    int sum = 0;
    for (int i = 0; i <= d; ++i) {
      const float v = down_factor * elevated[i];
      const float up = ceilf(v) * up_factor, down = floorf(v) * up_factor;
      const int rd2 = (up - elevated[i] < elevated[i] - down) ? (int)up : (int)down;
      rem0[i] = (float)rd2;
      sum += (int)((float)rd2 * down_factor);
    }
    for (int i = 0; i <= d; ++i) rank[i] = 0;
    for (int i = 0; i < d; ++i) {
      const float di = elevated[i] - rem0[i];
      for (int j = i + 1; j <= d; ++j) {
        if (di < elevated[j] - rem0[j]) rank[i]++;
        else rank[j]++;
      }
    }
    for (int i = 0; i <= d; ++i) {
      rank[i] += sum;
      if (rank[i] < 0) {
        rank[i] += d + 1;
        rem0[i] += (float)(d + 1);
      } else if (rank[i] > d) {
        rank[i] -= d + 1;
        rem0[i] -= (float)(d + 1);
      }
    }
    for (int i = 0; i <= d + 1; ++i) barycentric[i] = 0.0f;
    for (int i = 0; i <= d; ++i) {
      const float v = (elevated[i] - rem0[i]) * down_factor;
      barycentric[d - rank[i]] += v;
      barycentric[d - rank[i] + 1] -= v;
    }
    barycentric[0] += 1.0f + barycentric[d + 1];

    for (int r = 0; r <= d; ++r) {
      for (int i = 0; i < d; ++i) key[i] = (short)((int)rem0[i] + canonical[r * (d + 1) + rank[i]]);
      L->offset[(size_t)k * (d + 1) + r] = hash_find(&H, key, 1);
      L->bary[(size_t)k * (d + 1) + r] = barycentric[r];
    }
  }
  const int M = H.filled;
  L->M = M;
  L->n1 = (int*)malloc(sizeof(int) * (size_t)(d + 1) * M);
  L->n2 = (int*)malloc(sizeof(int) * (size_t)(d + 1) * M);
  for (int j = 0; j <= d; ++j) {
    for (int i = 0; i < M; ++i) {
      const short* key = H.keys + (size_t)i * d;
      short a[MAXD], b[MAXD];
      for (int k = 0; k < d; ++k) {
        a[k] = (short)(key[k] - 1);
        b[k] = (short)(key[k] + 1);
      }
      if (j < d) {  /* the (d+1)-th coordinate is implicit (keys store the first d) */
        a[j] = (short)(key[j] + d);
        b[j] = (short)(key[j] - d);
      }
      L->n1[(size_t)j * M + i] = hash_find(&H, a, 0);
      L->n2[(size_t)j * M + i] = hash_find(&H, b, 0);
    }
  }
  hash_free(&H);
}

static void lattice_free(Lattice* L) {
  free(L->offset);
  free(L->bary);
  free(L->n1);
  free(L->n2);
}

/* out[N][vs] = lattice filter of in[N][vs] */
static void lattice_compute(const Lattice* L, float* out, const float* in, int vs) {
  const int d = L->d, N = L->N, M = L->M;
  float* values = (float*)calloc((size_t)(M + 2) * vs, sizeof(float));
  float* new_values = (float*)calloc((size_t)(M + 2) * vs, sizeof(float));
  for (int i = 0; i < N; ++i)
    for (int j = 0; j <= d; ++j) {
      const int o = L->offset[(size_t)i * (d + 1) + j] + 1;
      const float w = L->bary[(size_t)i * (d + 1) + j];
      for (int k = 0; k < vs; ++k) values[(size_t)o * vs + k] += w * in[(size_t)i * vs + k];
    }
  for (int j = 0; j <= d; ++j) {
    for (int i = 0; i < M; ++i) {
      const float* old_val = values + (size_t)(i + 1) * vs;
      float* new_val = new_values + (size_t)(i + 1) * vs;
      const float* n1 = values + (size_t)(L->n1[(size_t)j * M + i] + 1) * vs;
      const float* n2 = values + (size_t)(L->n2[(size_t)j * M + i] + 1) * vs;
      for (int k = 0; k < vs; ++k) new_val[k] = old_val[k] + 0.5f * (n1[k] + n2[k]);
    }
    float* t = values;
    values = new_values;
    new_values = t;
  }
  const float alpha = 1.0f / (1.0f + powf(2.0f, -(float)d));
  for (int i = 0; i < N; ++i) {
    for (int k = 0; k < vs; ++k) out[(size_t)i * vs + k] = 0.0f;
    for (int j = 0; j <= d; ++j) {
      const int o = L->offset[(size_t)i * (d + 1) + j] + 1;
      const float w = L->bary[(size_t)i * (d + 1) + j];
      for (int k = 0; k < vs; ++k) out[(size_t)i * vs + k] += w * values[(size_t)o * vs + k] * alpha;
    }
  }
  free(values);
  free(new_values);
}

static void exp_and_normalize(float* Q, const float* in, int N, int C) {
  for (int i = 0; i < N; ++i) {
    const float* b = in + (size_t)i * C;
    float mx = b[0];
    for (int k = 1; k < C; ++k) mx = b[k] > mx ? b[k] : mx;
    float s = 0.0f;
    for (int k = 0; k < C; ++k) {
      Q[(size_t)i * C + k] = expf(b[k] - mx);
      s += Q[(size_t)i * C + k];
    }
    for (int k = 0; k < C; ++k) Q[(size_t)i * C + k] /= s;
  }
}

/* img: uint8 [H][W][3]; unary: float [C][N] energies; Q_out: float [C][N]; lattice_sizes: int[2] (may be NULL).
 * Either kernel is skipped when its weight is 0. */
int densecrf_ref_inference(const unsigned char* img, const float* unary, int W, int H, int C, float pos_w,
                           float pos_sxy, float bi_w, float bi_sxy, float bi_srgb, int iters, float* Q_out,
                           int* lattice_sizes) {
  const int N = W * H;
  Lattice L[2];
  float* norm[2] = {NULL, NULL};
  float weight[2] = {pos_w, bi_w};
  int use[2] = {pos_w != 0.0f, bi_w != 0.0f};
  float* feat = (float*)malloc(sizeof(float) * (size_t)N * 5);
  float* ones = (float*)malloc(sizeof(float) * (size_t)N);
  for (int i = 0; i < N; ++i) ones[i] = 1.0f;
  for (int k = 0; k < 2; ++k) {
    if (!use[k]) continue;
    const int d = k == 0 ? 2 : 5;
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        float* f = feat + (size_t)(y * W + x) * d;
        if (k == 0) {
          f[0] = (float)x / pos_sxy;
          f[1] = (float)y / pos_sxy;
        } else {
          f[0] = (float)x / bi_sxy;
          f[1] = (float)y / bi_sxy;
          for (int c = 0; c < 3; ++c) f[2 + c] = (float)img[(size_t)(y * W + x) * 3 + c] / bi_srgb;
        }
      }
    lattice_init(&L[k], feat, d, N);
    norm[k] = (float*)malloc(sizeof(float) * (size_t)N);
    lattice_compute(&L[k], norm[k], ones, 1);
    for (int i = 0; i < N; ++i) norm[k][i] = 1.0f / sqrtf(norm[k][i] + 1e-20f);
    if (lattice_sizes) lattice_sizes[k] = L[k].M;
  }
  float* U = (float*)malloc(sizeof(float) * (size_t)N * C);  /* pixel-major [N][C] like Eigen's column-major M x N */
  float* Q = (float*)malloc(sizeof(float) * (size_t)N * C);
  float* t1 = (float*)malloc(sizeof(float) * (size_t)N * C);
  float* t2 = (float*)malloc(sizeof(float) * (size_t)N * C);
  float* t3 = (float*)malloc(sizeof(float) * (size_t)N * C);
  for (int i = 0; i < N; ++i)
    for (int k = 0; k < C; ++k) U[(size_t)i * C + k] = unary[(size_t)k * N + i];
  for (size_t e = 0; e < (size_t)N * C; ++e) t1[e] = -U[e];
  exp_and_normalize(Q, t1, N, C);
  for (int it = 0; it < iters; ++it) {
    for (size_t e = 0; e < (size_t)N * C; ++e) t1[e] = -U[e];
    for (int k = 0; k < 2; ++k) {
      if (!use[k]) continue;
      for (int i = 0; i < N; ++i)
        for (int c = 0; c < C; ++c) t3[(size_t)i * C + c] = Q[(size_t)i * C + c] * norm[k][i];
      lattice_compute(&L[k], t2, t3, C);
      for (int i = 0; i < N; ++i)
        for (int c = 0; c < C; ++c) {
          const float msg = -weight[k] * (t2[(size_t)i * C + c] * norm[k][i]); /* Potts: -w K~Q */
          t1[(size_t)i * C + c] -= msg;
        }
    }
    exp_and_normalize(Q, t1, N, C);
  }
  for (int i = 0; i < N; ++i)
    for (int k = 0; k < C; ++k) Q_out[(size_t)k * N + i] = Q[(size_t)i * C + k];
  for (int k = 0; k < 2; ++k)
    if (use[k]) {
      lattice_free(&L[k]);
      free(norm[k]);
    }
  free(feat); free(ones); free(U); free(Q); free(t1); free(t2); free(t3);
  return 0;
}

/* Exact (brute-force) dense mean-field with the same kernels, O(N^2): used on tiny images to bound the
 * permutohedral approximation.  Same normalisation (symmetric) and update rule. */
int densecrf_bruteforce_inference(const unsigned char* img, const float* unary, int W, int H, int C, float pos_w,
                                  float pos_sxy, float bi_w, float bi_sxy, float bi_srgb, int iters, float* Q_out) {
  const int N = W * H;
  float* Kmat[2] = {NULL, NULL};
  float* norm[2] = {NULL, NULL};
  float weight[2] = {pos_w, bi_w};
  for (int k = 0; k < 2; ++k) {
    if (weight[k] == 0.0f) continue;
    Kmat[k] = (float*)malloc(sizeof(float) * (size_t)N * N);
    norm[k] = (float*)malloc(sizeof(float) * (size_t)N);
    for (int i = 0; i < N; ++i) {
      double s = 0.0;
      for (int j = 0; j < N; ++j) {
        const float sxy = k == 0 ? pos_sxy : bi_sxy;
        float dx = (float)(i % W - j % W) / sxy, dy = (float)(i / W - j / W) / sxy;
        float d2 = dx * dx + dy * dy;
        if (k == 1)
          for (int c = 0; c < 3; ++c) {
            const float dc = ((float)img[(size_t)i * 3 + c] - (float)img[(size_t)j * 3 + c]) / bi_srgb;
            d2 += dc * dc;
          }
        Kmat[k][(size_t)i * N + j] = expf(-0.5f * d2);
        s += Kmat[k][(size_t)i * N + j];
      }
      norm[k][i] = 1.0f / sqrtf((float)s + 1e-20f);
    }
  }
  float* U = (float*)malloc(sizeof(float) * (size_t)N * C);
  float* Q = (float*)malloc(sizeof(float) * (size_t)N * C);
  float* t1 = (float*)malloc(sizeof(float) * (size_t)N * C);
  for (int i = 0; i < N; ++i)
    for (int k = 0; k < C; ++k) U[(size_t)i * C + k] = unary[(size_t)k * N + i];
  for (size_t e = 0; e < (size_t)N * C; ++e) t1[e] = -U[e];
  exp_and_normalize(Q, t1, N, C);
  for (int it = 0; it < iters; ++it) {
    for (size_t e = 0; e < (size_t)N * C; ++e) t1[e] = -U[e];
    for (int k = 0; k < 2; ++k) {
      if (weight[k] == 0.0f) continue;
      for (int i = 0; i < N; ++i)
        for (int c = 0; c < C; ++c) {
          double s = 0.0;
          for (int j = 0; j < N; ++j) s += (double)Kmat[k][(size_t)i * N + j] * norm[k][j] * Q[(size_t)j * C + c];
          t1[(size_t)i * C + c] += weight[k] * (float)s * norm[k][i];
        }
    }
    exp_and_normalize(Q, t1, N, C);
  }
  for (int i = 0; i < N; ++i)
    for (int k = 0; k < C; ++k) Q_out[(size_t)k * N + i] = Q[(size_t)i * C + k];
  for (int k = 0; k < 2; ++k) {
    free(Kmat[k]);
    free(norm[k]);
  }
  free(U); free(Q); free(t1);
  return 0;
}
