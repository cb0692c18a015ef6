/* libdupl.so — C ABI of the B200-native DuPL hot path.
 *
 * The reference (Wu0409/DuPL) is pure Python: its "operator API" for this path is the module
 * surface model/model_dupl.py, utils/cam_helper.py (= utils/camutils.py), model/PAR.py,
 * model/losses.py, utils/dcrf.py.  There is no FFI in the reference; each entry point below
 * names the reference function (file:line, relative to the reference root) whose arithmetic it
 * replaces.  dupl_b200/ binds these with ctypes and re-exports the reference's Python
 * signatures.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; the caller (PyTorch's
 *    caching allocator) owns every buffer, including workspaces; the library never allocates
 *    or frees device memory and keeps no pointer across calls;
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises
 *    and no call reads device data on the host;
 *  - tensors are dense row-major ("contiguous") in the stated shape, fp32 unless stated;
 *    "split bf16" means two bf16 planes hi/lo with x ~= hi + lo (|err| <= 2^-17 |x|), the
 *    operand format of the 3-pass tcgen05 GEMMs (hi*hi + hi*lo + lo*hi, fp32 accumulate);
 *  - return value 0 = success, negative = DUPL_ERR_*; dupl_last_error() returns a
 *    thread-local message for the last failure on the calling thread.
 */
#ifndef DUPL_H_
#define DUPL_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DUPL_ABI_VERSION 1

#define DUPL_OK 0
#define DUPL_ERR_INVALID_ARGUMENT (-1)
#define DUPL_ERR_CUDA (-2)
#define DUPL_ERR_UNSUPPORTED (-3)

int dupl_version(void);
const char* dupl_last_error(void);
/* Number of CUDA kernels launched by this library in the calling process so far. */
int64_t dupl_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Dense path: ViT-B/16 encoder (model/backbone/vit.py:87-184,223-334) and the CAM contraction
 * (model/model_dupl.py:69-84).
 * ---------------------------------------------------------------------------------------- */

#define DUPL_MAX_SEGMENTS 8
#define DUPL_MAX_GROUPS 2
#define DUPL_MAX_KSPLIT 8

/* A "segment" is a batch of equally-sized images flowing through the encoder: `batch` images of
 * `tokens` = 1 + gh*gw tokens each, stored at rows [row_offset, row_offset + batch*tokens) of the
 * token matrix (cls token first within each image, as vit.py:299-301).  The three scales of
 * multi_scale_cam2_siamese (utils/cam_helper.py:164-204) are three segments of one matrix, so
 * every linear layer is ONE GEMM over all scales. */
typedef struct {
  int32_t batch;
  int32_t gh, gw;      /* patch grid (H/16, W/16) */
  int32_t tokens;      /* 1 + gh*gw */
  int32_t row_offset;  /* first token row of this segment */
  int32_t patch_row_offset; /* first row of this segment in the patch matrix (no cls rows) */
} dupl_segment;

/* x (fp32) -> split bf16 planes.  n elements. */
int dupl_split_bf16(const float* x, void* hi, void* lo, int64_t n, void* stream);

/* Epilogues of dupl_gemm_bf16x3 */
#define DUPL_EPI_F32 0        /* out_f32 = acc + bias                                         */
#define DUPL_EPI_SPLIT 1      /* out_hi/lo = split(acc + bias)            (qkv: vit.py:122)     */
#define DUPL_EPI_GELU_SPLIT 2 /* out_hi/lo = split(gelu_erf(acc + bias))  (fc1+act: vit.py:98-99); out_f32 (optional) = acc + bias */
#define DUPL_EPI_RESID 3      /* out_f32 = resid + acc + bias  (proj / fc2 + residual: vit.py:158-159) */
#define DUPL_EPI_PATCH 4      /* out_f32[token row] = acc + bias + pos_embed (vit.py:292-304)  */
#define DUPL_EPI_RELU_SPLIT 5 /* out_hi/lo = split(relu(acc + bias))   (conv6/conv7 + ReLU: conv_head.py:34-38) */

typedef struct {
  const void* a_hi;
  const void* a_lo; /* [M, K] split bf16, row stride lda */
  const void* w_hi;
  const void* w_lo;    /* [N, K] split bf16 (nn.Linear weight layout), row stride K */
  const float* bias;   /* [N] or NULL */
  const float* resid;  /* DUPL_EPI_RESID: [M, ldo] (may alias out_f32) */
  float* out_f32;      /* F32 / RESID / PATCH */
  void* out_hi;
  void* out_lo;        /* SPLIT / GELU_SPLIT: [M, ldo] */
  const float* pos[DUPL_MAX_SEGMENTS]; /* PATCH: per segment [tokens, N] resized pos_embed */
  float* splitk_ws;    /* max_ksplit > 1: [max_ksplit, M, ldo] fp32 workspace for the split-K partial sums */
} dupl_gemm_group;

/* C[M,N] = A[M,K] * W[N,K]^T for `groups` independent problems of identical shape (the two
 * students).  K % 64 == 0, N % 16 == 0, 16-byte aligned planes and row strides.
 * a_mn_major / b_mn_major: the operand is read in place from its transposed storage (MN-major tcgen05 operand: the
 * contraction runs along the rows of the stored matrix) — what autograd's wgrad (dW = dY^T X: both operands) and dgrad
 * (dX = dY W: the weight) need, without materialising transposed copies.  With BOTH operands MN-major K may be any positive
 * number (rows past K read as zeros).
 * Tile width (256/192/128 columns per CTA pair) and, when max_ksplit > 1, a split-K factor are chosen per
 * shape so that the work items fill the 74 CTA pairs of a B200; split-K partial sums go through splitk_ws
 * and are added in a fixed order (bit-reproducible). */
typedef struct {
  int32_t groups;
  int32_t M, N, K;
  int32_t lda, ldo;
  int32_t epilogue;
  int32_t nseg;                         /* PATCH only */
  int32_t ldw;                          /* row stride of the W planes in elements (0 = K) */
  int32_t max_ksplit;                   /* 0/1 = no split-K; else the workspace holds this many partials (F32 epilogue, no bias) */
  int32_t f32_rows;                     /* GELU_SPLIT: only rows < f32_rows get the out_f32 side output (0 = all rows) */
  int32_t passes;                       /* 0/3: hi*hi + hi*lo + lo*hi (error ~2^-16 of sum|a_i w_i|); 4: also lo*lo (fp32-level) */
  int32_t a_mn_major;                   /* != 0: the A planes are stored [K, M] (row stride lda >= M): A is consumed TRANSPOSED in place */
  int32_t b_mn_major;                   /* != 0: the W planes are stored [K, N] (row stride ldw >= N, 0 = N); needs N >= 128 */
  dupl_segment seg[DUPL_MAX_SEGMENTS];  /* PATCH only: patch row -> token row mapping */
  dupl_gemm_group g[DUPL_MAX_GROUPS];
} dupl_gemm_args;

int dupl_gemm_bf16x3(const dupl_gemm_args* args, void* stream);
/* Host-only: the tile width (columns per CTA pair), split-K factor and number of work items dupl_gemm_bf16x3 would use for
 * this shape on the current device (148 SMs assumed without a device).  No launch. */
int dupl_gemm_plan(int32_t M, int32_t N, int32_t K, int32_t groups, int32_t max_ksplit, int32_t* tile_n, int32_t* ksplit,
                   int32_t* work_items);
/* Host-only: the persistent GEMM grid uses at most `sms` SMs from now on (0 = all).  The data-parallel training step
 * (train_final_voc.py:155,470: DistributedDataParallel overlaps the gradient all-reduce with the backward pass) leaves a few
 * SMs to the NCCL kernels that average finished gradient chunks while the remaining wgrad GEMMs run: a persistent grid
 * that owns every SM would make the collective wait for a kernel boundary and then delay the next GEMM's statically
 * assigned tiles.  Returns the previous limit. */
int dupl_set_gemm_sm_limit(int32_t sms);

/* LayerNorm over the last dim (biased variance, eps inside the sqrt; vit.py:146,152,256 with
 * eps=1e-6 from deit.py:100) of x[rows, cols] -> split bf16 planes and/or fp32 (either output may be
 * NULL; cols % 128 == 0, <= 1024). */
int dupl_layernorm_split(const float* x, const float* gamma, const float* beta, void* out_hi, void* out_lo,
                         float* out_f32, int32_t rows, int32_t cols, float eps, void* stream);

/* softmax(Q K^T * scale) V per (image, head)  (vit.py:120-135), fused flash-style.
 * qkv planes: [M, 3*heads*64] split bf16 as produced by the qkv GEMM (q | k | v, head-major);
 * out planes: [M, heads*64] split bf16.  Attention never crosses an image boundary. */
typedef struct {
  int32_t nseg;
  dupl_segment seg[DUPL_MAX_SEGMENTS];
  int32_t M;
  int32_t heads; /* head_dim is fixed at 64 */
  float scale;
  const void* qkv_hi;
  const void* qkv_lo;
  void* out_hi;
  void* out_lo;
  float* lse; /* optional [M, heads]: ln sum_j exp(scale * s_ij) per row and head (for dupl_attention_bwd) */
} dupl_attention_args;

int dupl_attention_fwd(const dupl_attention_args* args, void* stream);

/* Image -> patch matrix rows (PatchEmbed's im2col, vit.py:176-183), fused with the bilinear
 * (align_corners=False) resize of the input and the horizontal flip that
 * multi_scale_cam2_siamese applies (cam_helper.py:168,183-184).
 * images [b,3,H,W] fp32 are resized to hs x ws (hs == H and ws == W: no resize); the patch grid is
 * gh = hs/16, gw = ws/16 (a ragged right/bottom border is ignored, like the stride-16 conv does);
 * the segment gets 2b images when flip_twin != 0 (second half = resized images flipped along x).
 * Output rows [patch_row_offset ...) of split-bf16 planes [*, 768], column = c*256 + py*16 + px. */
int dupl_patchify(const float* images, int32_t b, int32_t H, int32_t W, const dupl_segment* seg, int32_t hs,
                  int32_t ws, int32_t flip_twin, void* out_hi, void* out_lo, void* stream);

/* Bicubic (A=-0.75, align_corners=False) resize of the 14x14 grid of pos_embed[197, D] to gh x gw,
 * cls row copied (vit.py:294-298). out [1 + gh*gw, D]. */
int dupl_pos_embed_resize(const float* pos_embed, float* out, int32_t src, int32_t gh, int32_t gw, int32_t D,
                          void* stream);

/* Writes the cls rows of the token matrix: tok[row_offset + i*tokens] = cls_token + pos[0]. */
int dupl_cls_rows(float* tok, const float* cls_token, const float* const* pos, const dupl_segment* seg,
                  int32_t nseg, int32_t D, void* stream);

/* The feature x class-weight contraction of cam_only (model_dupl.py:82-83): for every patch token
 * (cls rows skipped) cam[img, k, p] = sum_d f(tok[row, d]) * w[k, d], with f = final LayerNorm
 * (vit.py:322) when gamma != NULL, identity otherwise (aux CAM from the un-normed block-9 output).
 * out: per segment a dense [batch, K, gh, gw] block at out + out_offset[s] floats. */
int dupl_cam_contract(const float* tok, const float* gamma, const float* beta, float eps, const float* w,
                      int32_t K, int32_t D, const dupl_segment* seg, int32_t nseg, float* out,
                      const int64_t* out_offset_host, void* stream);

/* ------------------------------------------------------------------------------------------
 * Decoder and classification heads (model/decoder/conv_head.py:33-41, model_dupl.py:86-95).
 * Rows of a token-major matrix are addressed as row(b, p) = row_offset + b*row_stride + first + p
 * (first = 1 skips the cls token of each image).
 * ---------------------------------------------------------------------------------------- */
/* im2col of a 3x3 / dilation d / zero-padding d convolution on token-major split-bf16 planes [*, Cin]:
 * out[(b*gh*gw + y*gw + x)][tap*Cin + c], tap = ky*3 + kx; the conv is then dupl_gemm_bf16x3 against
 * the weight re-laid out as [Cout][tap][Cin]. */
int dupl_im2col3x3(const void* in_hi, const void* in_lo, void* out_hi, void* out_lo, int32_t B, int32_t gh, int32_t gw,
                   int32_t Cin, int32_t dilation, int32_t row_offset, int32_t row_stride, int32_t first, void* stream);
/* out[b][c][p] = src[row(b,p)][c], c < C (src row stride ld): token-major -> NCHW (to_2D, model_dupl.py:64-67). */
int dupl_rows_to_nchw(const float* src, float* out, int32_t B, int32_t np, int32_t C, int32_t ld, int32_t row_offset,
                      int32_t row_stride, int32_t first, void* stream);
/* logits[b][k] = sum_d (max_p x[row(b,p)][d]) w[k][d]  (GMP + bias-free 1x1 classifier, model_dupl.py:88-98);
 * pooled: [B][D] output of the max pool (scratch for the caller); argmax (optional, int32 [B][D]) keeps the pooled rows. */
int dupl_gmp_classify(const float* x, const float* w, float* logits, float* pooled, int32_t* argmax, int32_t B, int32_t np,
                      int32_t D, int32_t K, int32_t row_offset, int32_t row_stride, int32_t first, void* stream);

/* ------------------------------------------------------------------------------------------
 * multi_scale_cam2_siamese post-processing (utils/cam_helper.py:173-202): per scale bilinear
 * up-sample to (H,W), max with the flipped twin, ReLU; sum over scales; per-(b,k) min-shift and
 * max-normalise (+1e-5).  lowres[s]: [2b, K, gh_s, gw_s]; out [b, K, H, W].
 * minmax: workspace of 2*b*K floats.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t nscale;
  const float* lowres[DUPL_MAX_SEGMENTS];
  int32_t gh[DUPL_MAX_SEGMENTS], gw[DUPL_MAX_SEGMENTS];
  int32_t b, K, H, W;
  float* out;
  float* minmax;
} dupl_mscam_args;

int dupl_mscam_post(const dupl_mscam_args* args, void* stream);

/* cam_to_label / cam_to_label_dynamic_cls (utils/cam_helper.py:8-55).
 * cam [b,K,h,w]; cls_label [b,K] fp32; img_box [b,4] int32 (y0,y1,x0,x1) or NULL;
 * high_thre: per-image [b] (device) or NULL to use high_thre_scalar.
 * valid_cam (optional) [b,K,h,w]; label int64 [b,h,w]. */
typedef struct {
  const float* cam;
  const float* cls_label;
  const int32_t* img_box;
  const float* high_thre;
  float high_thre_scalar, low_thre, bkg_thre;
  int32_t ignore_mid;
  int64_t ignore_index;
  int32_t b, K, h, w;
  float* valid_cam;
  int64_t* label;
} dupl_cam_to_label_args;

int dupl_cam_to_label(const dupl_cam_to_label_args* args, void* stream);

/* label_to_aff_mask (utils/cam_helper.py:323-335): label int64 [b,n] -> aff int64 [b,n,n]. */
int dupl_label_to_aff_mask(const int64_t* label, int64_t* aff, int32_t b, int32_t n, int64_t ignore_index,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * PAR (model/PAR.py:64-91) and refine_cams_with_* (utils/cam_helper.py:338-440)
 * ---------------------------------------------------------------------------------------- */
#define DUPL_PAR_MAX_DIL 8

/* aff[B, 8*ndil, h, w] = softmax_n(-mean_c((|I_c(n)-I_c(p)|/(std_c(p)+1e-8)/w1)^2)) + w2*softmax_n(pos prior)
 * (PAR.py:70-89), replicate border. imgs [B,C,h,w]. */
int dupl_par_affinity(const float* imgs, float* aff, int32_t B, int32_t C, int32_t h, int32_t w,
                      const int32_t* dilations_host, int32_t ndil, float w1, float w2, void* stream);

/* num_iter x: m_c(p) <- sum_n aff(p,n) m_c(nbr_n(p)) (PAR.py:88-90).  masks / scratch: [B, P, h, w]
 * plane stacks used as ping-pong buffers; nactive (device int32 [B], may be NULL = all P) gives the
 * number of leading live planes per image, the rest is neither read nor written.
 * *result_in_scratch_host (host, optional) reports where the result lives (num_iter odd => scratch). */
int dupl_par_propagate(const float* aff, float* masks, float* scratch, const int32_t* nactive, int32_t B, int32_t P,
                       int32_t h, int32_t w, const int32_t* dilations_host, int32_t ndil, int32_t num_iter,
                       int32_t* result_in_scratch_host, void* stream);

/* Prologue of refine_cams_with_dynamic_thres / _bkg_v2 (cam_helper.py:338-417): 2x2-mean
 * down-sample of the image and of [bkg | cams], softmax over the PRESENT channels (bkg + classes
 * with cls_label != 0; absent channels excluded exactly like the reference's nonzero+gather),
 * for the high and the low background variant — without the reference's torch.nonzero host sync.
 * images [b,3,H,W]; cams [b,K,H,W]; cls_label [b,K]; bkg_h: map [b,1,H,W] or NULL (use bkg_h_scalar).
 * out: images_ds [b,3,H/2,W/2]; masks [b, P = 2*(K+1), H/2, W/2]: live planes of image i are
 * a = v*nch_i + slot with v in {0: high, 1: low}, slot 0 = background, slot s = s-th present class in
 * ascending order (the reference's valid_key), nch_i = 1 + #present; nactive[i] = 2*nch_i (device). */
typedef struct {
  const float* images;
  const float* cams;
  const float* cls_label;
  const float* bkg_h;
  float bkg_h_scalar, bkg_l_scalar;
  int32_t b, K, H, W;
  float* images_ds;
  float* masks;
  int32_t* nactive;
} dupl_refine_prologue_args;

int dupl_refine_prologue(const dupl_refine_prologue_args* args, void* stream);

/* Epilogue (cam_helper.py:419-431,434-440): bilinear x2 up-sample of the propagated masks, arg-max
 * over the live channels (first index on ties), valid_key lookup, img_box paste on an ignore canvas,
 * high/low merge.  masks as above -> label fp32 [b,H,W] in {0..K, ignore}. */
typedef struct {
  const float* masks;
  const float* cls_label;
  const int32_t* img_box;
  int32_t b, K, H, W;
  float ignore_index;
  float* label;
  float* label_h; /* optional per-variant outputs (may be NULL) */
  float* label_l;
} dupl_refine_epilogue_args;

int dupl_refine_epilogue(const dupl_refine_epilogue_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Backward pass of the training step (train_final_voc.py:470-471 `loss.backward()` through
 * model/model_dupl.py, vit.py, conv_head.py).  dgrad and wgrad of every Linear / conv are
 * dupl_gemm_bf16x3 calls on transposed operand planes; the kernels below provide the operands and the
 * non-GEMM gradients.  Row addressing: row(r) = r when tokens == 0, else (r / np)*tokens + first + r % np.
 * ---------------------------------------------------------------------------------------- */
/* src fp32 [R rows (mapped), Cc] (row stride ld) -> planes hi/lo [R, Cc] (optional) and transposed planes
 * t_hi/t_lo [Cc, Rpad] (optional; columns R..Rpad-1 zero so that Rpad % 64 == 0 can be a GEMM contraction). */
int dupl_split_transpose(const float* src, int32_t R, int32_t Cc, int32_t ld, int32_t tokens, int32_t np, int32_t first,
                         void* hi, void* lo, void* t_hi, void* t_lo, int32_t Rpad, float* colsum_ws, float* colsum,
                         void* stream);
/* Same for the gradient w.r.t. GELU's output (autograd of vit.py:97-103): every element of src [R, Cc] is multiplied by
 * GELU'(gelu_pre) (the saved fc1 pre-activation, [R, Cc] fp32) on the way — the hidden gradient never makes a separate
 * element-wise round trip through HBM. */
int dupl_split_transpose_gelu(const float* src, const float* gelu_pre, int32_t R, int32_t Cc, void* hi, void* lo, void* t_hi, void* t_lo,
                              int32_t Rpad, float* colsum_ws, float* colsum, void* stream);
/* (with colsum != NULL the same pass also produces colsum[c] = sum_r src[row(r)][c], the bias gradient of the
 * layer; colsum_ws: scratch of ceil(rows/64)*Cc floats, rows = Rpad when t_hi is given, else R.) */
/* one or two bf16 planes [R (mapped), Cc] (row stride ld) -> their transposes [Cc, Rpad], zero padded;
 * in_lo/out_lo may be NULL.  Cc, ld, Rpad even. */
int dupl_transpose_planes(const void* in_hi, const void* in_lo, int32_t R, int32_t Cc, int32_t ld, int32_t tokens,
                          int32_t np, int32_t first, void* out_hi, void* out_lo, int32_t Rpad, void* stream);
/* The same for up to DUPL_MAX_TRANSPOSE_ITEMS dense plane pairs in ONE launch (all activations / weights one encoder block's
 * wgrad and dgrad GEMMs consume, both students): in [R, ld >= Cc] -> out [Cc, Rpad] zero padded. */
#define DUPL_MAX_TRANSPOSE_ITEMS 16
typedef struct dupl_transpose_item {
  const void* in_hi; const void* in_lo;
  void* out_hi; void* out_lo;
  int32_t R, Cc, ld, Rpad;
} dupl_transpose_item;
int dupl_transpose_planes_multi(const dupl_transpose_item* items, int32_t n_items, void* stream);
/* out[c] = sum_r x[row(r)][c]  (bias gradients) */
int dupl_colsum(const float* x, int32_t R, int32_t Cc, int32_t ld, int32_t tokens, int32_t np, int32_t first, float* out,
                void* stream);
/* LayerNorm backward (vit.py:146,152,256): dres[rows, cols] += dL/dx; dgamma, dbeta [cols];
 * partial: scratch of 2*cols*ceil(rows/8) floats.  cols == 768. */
int dupl_layernorm_bwd(const float* dy, const float* x, const float* gamma, float* dres, float* partial, float* dgamma,
                       float* dbeta, int32_t rows, int32_t cols, float eps, void* stream);
/* d[i] *= gelu'(pre[i]) (exact erf form) */
int dupl_gelu_bwd(float* d, const float* pre, int64_t n, void* stream);
/* d[i] = 0 where the saved relu output (split planes) is 0 */
int dupl_relu_bwd(float* d, const void* act_hi, const void* act_lo, int64_t n, void* stream);
/* gradient of dupl_im2col3x3: din[row(b,y,x)][c] (+)= sum_tap dcol[(b, y-dy, x-dx)][tap*Cin + c] */
int dupl_col2im3x3(const float* dcol, float* din, int32_t B, int32_t gh, int32_t gw, int32_t Cin, int32_t dilation,
                   int32_t ld_in, int32_t tokens, int32_t first, int32_t accumulate, void* stream);
/* dst[row(b*np + p)][c] += src[b][c][p]  (gradient of dupl_rows_to_nchw) */
int dupl_nchw_to_rows_add(const float* src, float* dst, int32_t B, int32_t np, int32_t Cc, int32_t ld, int32_t tokens,
                          int32_t first, void* stream);
/* gradient of dupl_gmp_classify: dx[row(b, argmax[b][d])][d] += sum_k dlogits[b][k] w[k][d];
 * dw[k][d] = sum_b dlogits[b][k] * pooled[b][d]; dw_partial: scratch [B, K, D]. */
int dupl_gmp_classify_bwd(const float* x, const float* w, const float* dlogits, const int32_t* argmax, float* dx,
                          float* dw_partial, float* dw, int32_t B, int32_t np, int32_t D, int32_t K, int32_t ld,
                          int32_t tokens, int32_t first, void* stream);
/* Attention backward for one segment (vit.py:120-135) on tcgen05: from the forward's qkv / output planes, lse and
 * the gradient dO (split-bf16 planes [M, heads*64]) -> dqkv fp32 [M, 3*heads*64].  Dvec: scratch [M, heads].
 * M = number of rows of the planes (bounds the TMA tensor maps). */
typedef struct {
  const void* qkv_hi;
  const void* qkv_lo;
  const void* o_hi;
  const void* o_lo;
  const void* do_hi;
  const void* do_lo;
  const float* lse;
  float* Dvec;
  float* dqkv;
  int32_t M, batch, tokens, row_offset, heads;
  float scale;
} dupl_attention_bwd_args;

int dupl_attention_bwd(const dupl_attention_bwd_args* args, void* stream);

/* ------------------------------------------------------------------------------------------
 * Losses (model/losses.py) with fused forward + backward.
 * ---------------------------------------------------------------------------------------- */
/* get_seg_loss (losses.py:24-39): 0.5 * [ sum CE over label==0 / (n_bg + 1e-6) + sum CE over label not in
 * {0, ignore} / (n_fg + 1e-6) ].  pred fp32 [b,C,H,W]; label int64 [b,H,W].
 * lse: [b,H,W] scratch kept for backward; partials: 4*ceil(b*H*W/256) floats; stats: 5 floats
 * (sum_bg, sum_fg, n_bg, n_fg, loss). */
int dupl_seg_loss_fwd(const float* pred, const int64_t* label, int32_t b, int32_t C, int32_t H, int32_t W,
                      int64_t ignore_index, float* lse, float* partials, float* stats, void* stream);
/* dpred = grad_out[0] * d loss / d pred. */
int dupl_seg_loss_bwd(const float* pred, const int64_t* label, const float* lse, const float* stats,
                      const float* grad_out, int32_t b, int32_t C, int32_t H, int32_t W, int64_t ignore_index,
                      float* dpred, void* stream);
/* The same loss on logits that the caller would first up-sample (train_final_voc.py:345-352: F.interpolate(segs,
 * size=label.shape[1:], mode='bilinear', align_corners=False) then get_seg_loss), without materialising them:
 * pred fp32 [b,C,h,w] low resolution, label int64 [b,H,W]; lse [b,H,W]; partials: 4*b*ceil(H/32)*ceil(W/32) floats;
 * dpred [b,C,h,w] is the gradient wrt the LOW-resolution logits (the transpose of the interpolation applied). */
int dupl_seg_loss_up_fwd(const float* pred, const int64_t* label, int32_t b, int32_t C, int32_t h, int32_t w, int32_t H,
                         int32_t W, int64_t ignore_index, float* lse, float* partials, float* stats, void* stream);
int dupl_seg_loss_up_bwd(const float* pred, const int64_t* label, const float* lse, const float* stats,
                         const float* grad_out, int32_t b, int32_t C, int32_t h, int32_t w, int32_t H, int32_t W,
                         int64_t ignore_index, float* dpred, void* stream);
/* get_masked_ptc_loss (losses.py:6-21): x fp32 [b,C,n] (n = h*w), mask int64 [b,n,n] with values {0,1,other};
 * inv: [b,n]; Gs: [b,n,n] signed cosine matrix kept for backward; partials: 4*b*ceil(n/64)^2 floats;
 * stats: 5 floats (sum_pos, sum_neg, n_pos, n_neg, loss). */
int dupl_ptc_loss_fwd(const float* x, const int64_t* mask, int32_t b, int32_t C, int32_t n, float* inv, float* Gs,
                      float* partials, float* stats, void* stream);
int dupl_ptc_loss_bwd(const float* x, const int64_t* mask, const float* inv, const float* Gs, const float* stats,
                      const float* grad_out, int32_t b, int32_t C, int32_t n, float* dxh_scratch, float* dx, void* stream);

/* The same loss with its two contractions on the tensor cores (model/losses.py:6-21): the host sequences
 *   dupl_ptc_prepare  ->  dupl_gemm_bf16x3 (G_i = x_hat_i x_hat_i^T per image)  ->  dupl_ptc_mask_reduce          (forward)
 *   dupl_ptc_dg  ->  dupl_gemm_bf16x3 (dX_hat_i = T_i x_hat_i)  ->  dupl_ptc_norm_bwd_rows                          (backward)
 * rows_*: split-bf16 x_hat [b*n, C] (token-major); cm_*: split-bf16 x_hat [b, C, npad] (channel-major, zero padded to
 * npad = multiple of 64); G fp32 [b, n, n]; t_*: split-bf16 T = S + S^T [b*n, npad]; dxh_rows fp32 [b*n, C]. */
int dupl_ptc_prepare(const float* x, int32_t b, int32_t C, int32_t n, int32_t npad, float* inv, void* rows_hi, void* rows_lo,
                     void* cm_hi, void* cm_lo, void* stream);
/* also re-derives in fp32 (from x, inv) the entries of G closer to zero than the split-bf16 rounding, in place */
int dupl_ptc_mask_reduce(float* G, const int64_t* mask, const float* x, const float* inv, int32_t b, int32_t C, int32_t n,
                         float* partials, int32_t nblocks, float* stats, void* stream);
int dupl_ptc_dg(const float* G, const int64_t* mask, const float* stats, const float* grad_out, int32_t b, int32_t n,
                int32_t npad, void* t_hi, void* t_lo, void* stream);
int dupl_ptc_norm_bwd_rows(const float* x, const float* inv, const float* dxh_rows, int32_t b, int32_t C, int32_t n, float* dx,
                           void* stream);

/* Classification loss of the loop (train_final_voc.py:299-305): sum over the T (= 4) logit tensors [b, K] of
 * F.multilabel_soft_margin_loss(logits_t, cls_label) = mean over b*K of -[y logsigmoid(x) + (1 - y) logsigmoid(-x)].
 * logits / grads: HOST arrays of T (<= 8) device pointers; label fp32 [n = b*K]; loss / grad_out: device scalars. */
int dupl_cls_loss_fwd(const float* const* logits, int32_t T, const float* label, int32_t n, float* loss, void* stream);
int dupl_cls_loss_bwd(const float* const* logits, float* const* grads, int32_t T, const float* label, int32_t n,
                      const float* grad_out, void* stream);
/* Discrepancy loss (train_final_voc.py:440-447): (1 + mean cos(f1.detach(), f2)) + (1 + mean cos(f2.detach(), f1)) with
 * nn.CosineSimilarity(dim=-1, eps) over rows = b*768 vectors of n = 784 spatial positions.  bwd: d1 = d loss / d f1 (from the
 * second term), d2 = d loss / d f2 (from the first), both scaled by *grad_out. */
int dupl_sim_loss_fwd(const float* f1, const float* f2, int32_t rows, int32_t n, float eps, float* cos_rows, float* loss, void* stream);
int dupl_sim_loss_bwd(const float* f1, const float* f2, int32_t rows, int32_t n, float eps, const float* grad_out, float* d1, float* d2,
                      void* stream);

/* ------------------------------------------------------------------------------------------
 * Strong augmentation on the device (utils/imutils.py:305-317 augment_data_strong, utils/randomaug.py:161-262).
 * images fp32 [B,3,H,W] in [0,1] (the denormalised batch) -> out fp32 [B,3,H,W] = normalised, horizontally flipped result of
 * n_ops Pillow operations per image.  ops_dev: DEVICE int32 [n_ops][B], operation index per step and image in the order of
 * augment_list() (0 AutoContrast, 1 Equalize, 2 Posterize, 3 Color, 4 Contrast, 5 Brightness, 6 Sharpness), drawn by the caller
 * (random.choices on the host, as the reference does); magnitudes7: HOST array of the 7 operations' magnitudes
 * (m/30 * (max - min) + min).  Bit-exact with Pillow.  The caller owns the workspace.
 * ---------------------------------------------------------------------------------------- */
int dupl_randaug_workspace_bytes(int32_t B, int32_t H, int32_t W, size_t* bytes);
int dupl_randaug(const float* images, float* out, int32_t B, int32_t H, int32_t W, const int32_t* ops_dev, int32_t n_ops,
                 const float* magnitudes7, void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * Optimizer: fused multi-tensor PolyWarmupAdamW (utils/optimizer.py:38-68, utils/train_helper.py:21-52).
 * ---------------------------------------------------------------------------------------- */
typedef struct dupl_adamw_param {
  void* param;        /* fp32, updated in place */
  const void* grad;   /* fp32 (a view of the student's gradient arena) */
  void* exp_avg;      /* fp32 */
  void* exp_avg_sq;   /* fp32 */
  void* plane_hi;     /* bf16 split planes of the updated parameter in the same element order, or NULL */
  void* plane_lo;
  int64_t numel;      /* multiple of 4; every pointer 16-byte aligned */
  float lr;           /* initial learning rate of the parameter's group (utils/train_helper.py: lr, or 10 lr for heads/decoders) */
  int32_t reserved;
} dupl_adamw_param;

typedef struct dupl_adamw_args {
  const dupl_adamw_param* params; /* device table [n_params] */
  const int32_t* items;           /* device work list [n_items][2] = (parameter index, chunk of 2048 elements): dupl_adamw_items */
  const int32_t* active;          /* device [n_params]: 1 = this parameter has a gradient this step (torch skips p.grad is None) */
  int32_t* steps;                 /* device [n_params]: per-parameter step count, incremented for active parameters */
  float* coef;                    /* device scratch [n_params][2]: 1/bias_correction1, sqrt(bias_correction2) */
  const float* lr_scale;          /* device scalar: the schedule multiplier of this step (utils/optimizer.py:52-66) */
  int32_t n_params;
  int64_t n_items;
  float beta1, beta2, eps, weight_decay;
} dupl_adamw_args;

/* Host-only: fills the work list for tensors of the given sizes (items_xy may be NULL to query the count). */
int dupl_adamw_items(const int64_t* numel, int32_t n_params, int32_t* items_xy, int64_t capacity, int64_t* n_items);
/* One AdamW step of every active parameter (decoupled weight decay, bias correction from the per-parameter step count,
 * torch.optim.AdamW's operation order) + refresh of the split planes.  Two launches, no host synchronisation. */
int dupl_adamw_step(const dupl_adamw_args* args, void* stream);

/* GMM noise filter of the training loop (train_final_voc.py:358-394, sklearn GaussianMixture in the
 * reference): per image, 2-component 1-D mixture on loss[label not in {0, ignore} and loss > loss_min]; when more
 * than min_count samples exist and the means differ by more than valid_gap, every pixel whose posterior under
 * the high-mean component exceeds gamma and whose label != 0 gets label = ignore_index.
 * loss fp32 [b, n]; label fp32 [b, n] (in place); info int32 [b, 4] = samples, filtered?, EM iterations, #flipped. */
int dupl_gmm_filter(const float* loss, float* label, int32_t b, int32_t n, float ignore_index, float loss_min,
                    int32_t min_count, float valid_gap, float gamma, float reg_covar, int32_t max_iter, float tol,
                    int32_t* info, void* stream);

/* ------------------------------------------------------------------------------------------
 * DenseCRF mean-field inference (utils/dcrf.py:42-69 -> pydensecrf DenseCRF2D: setUnaryEnergy,
 * addPairwiseGaussian(sxy=pos_xy_std, compat=pos_w), addPairwiseBilateral(sxy=bi_xy_std,
 * srgb=bi_rgb_std, compat=bi_w), inference(iters)); also serves crf_inference / crf_inference_label
 * (dcrf.py:7-40) through the same parameters.  Permutohedral-lattice filtering, symmetric
 * normalisation, Potts compatibility.  Two phases so that the vertex arrays can be sized exactly:
 *   dupl_crf_build  : builds both lattices for `image` into `workspace`, writes meta (device int32[4]:
 *                     #vertices of the Gaussian lattice, #vertices of the bilateral lattice, key-range
 *                     overflow flag, 0).  The caller reads meta (its only host sync) and then calls
 *   dupl_crf_infer  : iters mean-field updates; `values` is a scratch of dupl_crf_values_bytes().
 * image: uint8 [H][W][3]; unary_or_probs: fp32 [C][H][W] (probabilities -> -log(clip(p,1e-5,1)), or
 * energies when input_is_energy != 0); out: fp32 Q [C][H][W].  C <= 96.  A pairwise term whose weight is
 * 0 is skipped.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
  int32_t W, H, C;
  float pos_w, pos_xy_std, bi_w, bi_xy_std, bi_rgb_std;
  int32_t iters;
  int32_t input_is_energy;
  const uint8_t* image;
  const float* unary_or_probs;
  float* out;
  void* workspace;
  size_t workspace_bytes;
  void* values;
  size_t values_bytes;
  int32_t* meta;
} dupl_crf_args;

int dupl_crf_workspace_bytes(int32_t W, int32_t H, size_t* bytes);
int dupl_crf_values_bytes(int32_t W, int32_t H, int32_t C, int32_t M_gauss, int32_t M_bilateral, size_t* bytes);
int dupl_crf_build(const dupl_crf_args* args, void* stream);
int dupl_crf_infer(const dupl_crf_args* args, int32_t M_gauss, int32_t M_bilateral, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DUPL_H_ */
