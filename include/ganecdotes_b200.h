/* ganecdotes_b200 - C ABI of the B200 (sm_100a) per-pixel hidden-feature clustering path.
 *
 * This is the drop-in boundary: every entry point replaces one op (or one fused
 * stage) of the reference's path.  Plain pointers and sizes only; all pointers are
 * DEVICE pointers unless stated; outputs are caller-allocated; calls are
 * asynchronous on `stream` (a cudaStream_t passed as void*); no hidden allocation
 * or synchronisation (the tcgen05 entry points encode TMA descriptors on the host,
 * which is synchronous CPU work only).  Return value: GX_OK or a negative error;
 * nothing throws across the ABI.  Re-entrant, no global state besides the
 * last-error slot.
 *
 * Citations `ref:` are relative to the reference repository root.
 */
#ifndef GANECDOTES_B200_H
#define GANECDOTES_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GX_OK 0
#define GX_ERR_ARG (-1)         /* invalid argument / unsupported shape          */
#define GX_ERR_CUDA (-2)        /* CUDA runtime/driver error: gx_last_cuda_error */
#define GX_ERR_UNSUPPORTED (-3) /* device is not sm_100                           */

#define GX_ABI_VERSION 203
int gx_version(void);              /* == GX_ABI_VERSION of the header the library was built from */
int gx_abi_sizeof(int which);      /* sizeof of gx_conv_desc (0), gx_gemm_desc (1), gx_gather_desc (2), gx_ll_desc (3); -1 otherwise */
int gx_last_cuda_error(void);             /* cudaError_t of the last GX_ERR_CUDA  */
const char* gx_error_string(int gx_code); /* static string                        */
int gx_device_ok(void);                   /* 1 if the current device is sm_100    */
/* CTA (= SM) budgets of the persistent kernels: tcgen05 contractions and the streaming
 * Sinkhorn/loss kernels; 0 = all SMs.  Disjoint budgets let both families run concurrently. */
int gx_set_sm_budget(int umma_ctas, int stream_ctas);

/* ------------------------------------------------------------------------------------
 * L0 native ops of the reference
 * ---------------------------------------------------------------------------------- */

/* upfirdn2d forward.  ref: lib/gan/optim/upfirdn2d.cpp:18-39 (pybind `upfirdn2d`),
 * lib/gan/optim/upfirdn2d_kernel.cu:217-379.  input [major,in_h,in_w,minor] fp32,
 * kernel [kh,kw] fp32 (NOT flipped by the caller; the op correlates with the
 * flipped kernel like the reference), out [major,out_h,out_w,minor] with
 * out = (in*up + pad0 + pad1 - k + down) / down.  Negative pads crop. */
int gx_upfirdn2d(const float* input, const float* kernel, float* out, int major, int in_h, int in_w, int minor,
                 int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                 int pad_y1, void* stream);

/* fused_bias_act.  ref: lib/gan/optim/fused_bias_act.cpp:18-36,
 * fused_bias_act_kernel.cu:18-85.  x += b[(i/step_b)%size_b]; act*10+grad in
 * {10,11,12,30,31,32}; out = y*scale.  bias/refer may be NULL ("empty tensor"). */
int gx_fused_bias_act(const float* input, const float* bias, const float* refer, float* out, long long n,
                      int step_b, int size_b, int act, int grad, float alpha, float scale, void* stream);

/* The same two ops on float16 / float64 tensors (dtype: 0 float32, 1 float16, 2 float64; every tensor of the call in
 * that type, FIR taps and bias included) - the reference instantiates its kernels with
 * AT_DISPATCH_FLOATING_TYPES_AND_HALF (upfirdn2d_kernel.cu:321, fused_bias_act_kernel.cu:127).  Arithmetic in
 * float for float16, in double for float64. */
int gx_upfirdn2d_t(int dtype, const void* input, const void* kernel, void* out, int major, int in_h, int in_w,
                   int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1,
                   int pad_y0, int pad_y1, void* stream);
int gx_fused_bias_act_t(int dtype, const void* input, const void* bias, const void* refer, void* out, long long n,
                        int step_b, int size_b, int act, int grad, float alpha, float scale, void* stream);

/* ------------------------------------------------------------------------------------
 * StyleGAN2 synthesis (ref: models/stylegan2/model.py)
 * ---------------------------------------------------------------------------------- */

/* PixelNorm, ref: model.py:105-110.  x,y [n,dim]. */
int gx_pixel_norm(const float* x, float* y, int n, int dim, void* stream);

/* EqualLinear, ref: model.py:223-252.  y[n,o] = act(sum_i x[n,i]*w[o,i]*w_scale + b[o]*b_scale);
 * act: 0 none, 1 fused leaky relu (0.2, *sqrt2).  b may be NULL.  ldx: row pitch of x in floats (>= in_dim;
 * a row of W+ [B, n_latent, D] is addressed in place with ldx = n_latent*D). */
int gx_equal_linear(const float* x, long long ldx, const float* w, const float* b, float* y, int n, int in_dim,
                    int out_dim, float w_scale, float b_scale, int act, void* stream);

/* out[i,:] = mean + psi*(w[i,:]-mean): the truncation trick, ref: model.py:594-602. */
int gx_truncate(const float* w, const float* mean, float* out, long long rows, int dim, float psi, void* stream);

/* One-time weight preparation of a modulated conv (ref: model.py:316-320,330):
 * w [cout,cin,k,k] fp32 -> split-bf16 planes w_hi/w_lo [cout, k*k*cin_ld] holding
 * scale*w with k index = (ky*k+kx)*cin_ld + ci (zero for ci >= cin; cin_ld = cin rounded up to 64),
 * and wsq [cout,cin] = sum_taps (scale*w)^2. */
int gx_modconv_prepare(const float* w, float scale, void* w_hi, void* w_lo, float* wsq, int cout, int cin,
                       int cin_ld, int k, void* stream);

/* demod[b,co] = rsqrt(sum_ci wsq[co,ci]*s[b,ci]^2 + 1e-8), ref: model.py:332-334. */
int gx_modconv_demod(const float* wsq, const float* s, float* demod, int batch, int cin, int cout, void* stream);

/* x_mod = x * s[b,c] split into bf16 hi/lo planes (NHWC).  x has `x_batch_stride`
 * elements between samples (0 = broadcast, e.g. the constant input, ref: model.py:385-395).
 * Planes have c_ld >= c channels per pixel (the caller zero-fills channels [c, c_ld)). */
int gx_modulate_split(const float* x, long long x_batch_stride, const float* s, void* hi, void* lo, int batch,
                      long long hw, int c, int c_ld, void* stream);

typedef struct gx_conv_desc {
  /* operands (bf16 planes; *_lo may be NULL when passes == 1) */
  const void* x_hi; /* [B,H,W,Cin] NHWC, already modulated by the style     */
  const void* x_lo;
  const void* w_hi; /* [Cout, taps*Cin] from gx_modconv_prepare              */
  const void* w_lo;
  int batch, h, w, cin, cout;
  int cin_ld;   /* channels per pixel of the x planes / per tap of the w planes (multiple of 64)   */
  int upsample; /* 0: 3x3 pad 1 -> [B,H,W,Cout]; 1: transposed stride 2 -> [B,2H+1,2W+1,Cout] */
  int passes;   /* 1: bf16, 3: split-bf16 (fp32-equivalent)                  */
  /* epilogue: v = acc*demod[b,co] + strength*noise[b,y,x] + bias[co]; act; */
  const float* demod;          /* [B,Cout] or NULL                            */
  const float* noise;          /* [*,Ho,Wo] or NULL                           */
  long long noise_batch_stride; /* 0: shared across the batch                 */
  const float* noise_strength; /* device scalar (ref: NoiseInjection.weight) */
  const float* bias;           /* [Cout] or NULL                              */
  int act;                     /* 0 none, 1 lrelu(0.2)*sqrt2, 2 lrelu(0.2) (nn.LeakyReLU, segmentor head) */
  float* out;                  /* fp32 NHWC [B,Ho,Wo,Cout]                    */
  const float* next_style;     /* [B,Cout] or NULL: also emit next conv input */
  void* next_hi;               /* bf16 NHWC [B,Ho,Wo,next_ld]                 */
  void* next_lo;
  int next_ld;                 /* channels per pixel of the next planes (>= cout; caller zero-fills the rest) */
  int block_n; /* 0 = auto (128 or 256) */
  int stages;  /* 0 = auto               */
  int cluster_pair; /* != 0: CTA pairs (tcgen05 cta_group::2) on adjacent pixel tiles, half of the weight tile
                       per SM (used when the layer has >= 16 pixel tiles) */
  int dilation; /* plain conv only: tap spacing d with padding d (0 = 1); ref: OneShotSegmentor's dilated
                   Conv2d stack, hfc_with_swav/swav_clustering.py:716-742 */
} gx_conv_desc;

/* Modulated 3x3 conv as implicit GEMM on tcgen05 (TMA im2col boxes, TMEM accumulators),
 * ref: ModulatedConv2d.forward model.py:327-368 in the algebraic form
 * y = demod * conv(scale*W, s*x).  Requires cin_ld % 64 == 0 (channels zero-padded), cout % 4 == 0.
 * With demod = noise = NULL and unmodulated planes it is a plain (optionally dilated) 3x3 conv + bias +
 * activation: the layers of the one-shot segmentor head. */
int gx_modconv(const gx_conv_desc* d, void* stream);

/* The same modulated 3x3 conv (stride 1, padding 1) for layers with FEW channels - the 128^2 / 256^2 layers of the
 * BagGAN generator have 32 / 16 (ref models/baggan/models.py:383-390) - as a direct fp32 conv: the style-modulated
 * halo tile of the fp32 NHWC input and the whole weight tensor in shared memory, register accumulators, nothing padded
 * to 64 channels.  x [B,H,W,cin] fp32 (UNmodulated), style [B,cin], w [9,cin,cout] fp32 with the equalised-lr scale
 * folded in (tap-major, cross-correlation order), demod [B,cout] or NULL; noise / bias / act / next_* as in gx_modconv.
 * Supported (cin, cout): see gx_modconv_small_supported; anything else returns GX_ERR_ARG (use gx_modconv). */
int gx_modconv_small(const float* x, const float* style, const float* w, const float* demod, const float* noise,
                     long long noise_batch_stride, const float* noise_strength, const float* bias, int act,
                     float* out, const float* next_style, void* next_hi, void* next_lo, int next_ld, int batch,
                     int h, int w_, int cin, int cout, void* stream);
int gx_modconv_small_supported(int cin, int cout);

/* Blur (upfirdn2d up=1, down=1, pad=(p0,p1)) of the transposed-conv output fused with
 * noise + bias + leaky-relu*sqrt2 and with the next conv's modulate+split, NHWC.
 * ref: Blur model.py:166-182 + NoiseInjection :371-382 + FusedLeakyReLU :15-43.
 * in [B,Hi,Wi,C]; fir [kh,kw] device fp32 (unflipped); out [B,Ho,Wo,C], Ho = Hi+p0+p1-kh+1. */
int gx_blur_noise_bias_act(const float* in, const float* fir, int kh, int kw, int pad0, int pad1,
                           const float* noise, long long noise_batch_stride, const float* noise_strength,
                           const float* bias, int act, float* out, const float* next_style, void* next_hi,
                           void* next_lo, int next_ld, int batch, int hi, int wi, int c, void* stream);

/* The same stage for a separable filter fir[ky][kx] = fir_y[ky] * fir_x[kx] (the shipped blur, make_kernel of a
 * 1-D kernel, ref model.py:113-121): horizontal then vertical 4-tap pass, 16 instead of 64 multiply-adds per 4
 * channels (the 2-D form is bound by instruction issue).  fir_x, fir_y: device fp32 [ntaps], unflipped; ntaps = 4.
 * Results differ from the 2-D form by rounding only. */
int gx_blur_sep_noise_bias_act(const float* in, const float* fir_x, const float* fir_y, int ntaps, int pad0, int pad1,
                               const float* noise, long long noise_batch_stride, const float* noise_strength,
                               const float* bias, int act, float* out, const float* next_style, void* next_hi,
                               void* next_lo, int next_ld, int batch, int hi, int wi, int c, void* stream);

/* ToRGB: 1x1 modulated conv without demodulation + bias (+ skip), ref: model.py:435-454.
 * x fp32 NHWC [B,H,W,C]; w [3,C]; s [B,C]; bias [3]; skip (already upsampled) [B,3,H,W] or NULL;
 * out [B,3,H,W] NCHW fp32. */
int gx_torgb(const float* x, const float* w, float w_scale, const float* s, const float* bias, const float* skip,
             float* out, int batch, int hw, int c, void* stream);

/* ------------------------------------------------------------------------------------
 * Dense contractions (tcgen05)
 * ---------------------------------------------------------------------------------- */
typedef struct gx_gemm_desc {
  const void* a_hi; /* bf16 plane(s) of A */
  const void* a_lo;
  const void* b_hi; /* bf16 plane(s) of B */
  const void* b_lo;
  long long lda, ldb; /* leading dimensions in elements                         */
  int a_mn_major;     /* 0: A stored [M,K] (K contiguous); 1: stored [K,M]      */
  int b_mn_major;     /* 0: B stored [N,K] (K contiguous); 1: stored [K,N]      */
  int m, n, k;
  int passes;        /* 1 or 3                                                   */
  float* c;          /* fp32 [M,N] row-major                                     */
  long long ldc;
  const float* bias; /* [N] or NULL, added once                                  */
  int split_k;       /* >= 1; > 1 accumulates atomically into c (caller zeroes c) */
  int accumulate;    /* != 0: c += A*B^T (atomic adds) even when split_k == 1       */
  int force_m128;    /* != 0: never use 256-row CTA tiles (tuning / tests)          */
  float* colexp_sum; /* optional [N], caller-zeroed: += sum_m exp2(colexp_scale*C[m,n]) - the first
                        Sinkhorn pass (u_k = sum_n exp(S_nk/eps)) fused into the score GEMM's epilogue */
  float colexp_scale;
  int block_n;       /* 0 = auto                                                 */
  int stages;        /* 0 = auto                                                 */
  int cluster_pair;  /* != 0: CTA pairs (tcgen05 cta_group::2) - one M=256 MMA spans two SMs,
                        each CTA loads its 128 A rows and half of the B tile: a third less operand traffic
                        per SM than two independent 128-row tiles */
  int ab_f16;        /* != 0: a_hi / b_hi are IEEE fp16 planes (passes must be 1): 11-bit significands
                        for operands bounded by 1 (normalised rows), error ~8x below bf16  */
} gx_gemm_desc;

/* C = A * B^T (+ bias): projection / prototype / gradient GEMMs
 * (ref: hfc_with_swav/swav_clustering.py:171,175 and their autograd backward). */
int gx_gemm(const gx_gemm_desc* d, void* stream);

/* Plain fp32 SIMT GEMM on the same operands (A = hi+lo, B = hi+lo): the on-device
 * cross-check used by the tests, not used on the product path. */
int gx_gemm_check(const gx_gemm_desc* d, void* stream);

/* fp32 [rows,cols] (row stride ld) -> bf16 hi (+lo) planes with row pitch ld_out (0: dense),
 * optionally transposed ([cols,rows], dense).  lo may be NULL.  With a pitch the planes of several
 * maps are written side by side (the K-concatenated operand of one resolution). */
int gx_split_planes(const float* x, long long ld, void* hi, void* lo, long long rows, long long cols,
                    int transpose, long long ld_out, void* stream);

/* fp32 [rows,cols] (row stride ld) -> one IEEE fp16 plane (dense), round to nearest even. */
int gx_round_f16(const float* x, long long ld, void* out, long long rows, long long cols, void* stream);

/* ------------------------------------------------------------------------------------
 * Per-pixel feature vectors (ref: hfc_with_swav/swav_clustering.py:108-182)
 * ---------------------------------------------------------------------------------- */
#define GX_MAX_LEVELS 16
typedef struct gx_gather_desc {
  int nlevels;
  const float* feat[GX_MAX_LEVELS]; /* fp32 NHWC [nimg,h_l,w_l,c_l]             */
  int h[GX_MAX_LEVELS], w[GX_MAX_LEVELS], c[GX_MAX_LEVELS];
  int out_h, out_w; /* resolution every map is (nearest) upsampled to           */
  int hlen;         /* channels kept after concatenation ([:, :hlen])           */
  /* rows: row r reads image row_img[r], source pixel row_src[r] (flat index in
   * out_h*out_w AFTER undoing rotate/flip; -1 = rotation fill -> zeros).
   * If row_src == NULL the rows are all pixels of all images in order. */
  const int* row_img;
  const int* row_src;
  long long nrows;
  void* a_hi; /* bf16 [nrows, ld]; may be NULL when only a_f32 / row_norm are wanted */
  void* a_lo; /* may be NULL                                                   */
  float* a_f32; /* optional fp32 copy [nrows, ld] (tests), may be NULL          */
  long long ld;
  float* row_norm; /* optional [nrows]: L2 norm of every gathered row (ref :361-362)     */
} gx_gather_desc;

/* nearest-upsample + concat + [:hlen] + rotate/flip + random-pixel sampling in one
 * gather, emitting the projection GEMM's A operand (ref: :108-130,:158-167,:358-359). */
int gx_gather_rows(const gx_gather_desc* d, void* stream);

/* zn[r,:] = z[src,:] / max(||z[src,:]||,1e-12) as split planes, src = row_idx ? row_idx[r] : r
 * (row_idx[r] = -1: an all-zero input row -> zero output); inv_norm[n] kept for backward
 * (ref: F.normalize at swav_clustering.py:174).  z [*,c]; outputs have n rows.  zn_f16 (optional):
 * the same rows rounded to one IEEE fp16 plane (operand of the single-pass fp16 score GEMM). */
int gx_l2norm_split(const float* z, const int* row_idx, void* zn_hi, void* zn_lo, void* zn_f16, float* inv_norm,
                    long long n, int c, void* stream);

/* backward of the normalisation: dz = (dzn - zn*(zn.dzn)) * inv_norm, emitted as split planes
 * and / or fp32 rows (either output may be NULL). */
int gx_l2norm_bwd_split(const float* dzn, const void* zn_hi, const void* zn_lo, const float* inv_norm, void* dz_hi,
                        void* dz_lo, float* dz_f32, long long n, int c, void* stream);

/* out[seg,:] = sum_{r in [seg_off[seg], seg_off[seg+1])} rows[order[r],:] as bf16 planes: folds the
 * dZ rows of pixels that were sampled by several patches into one row per pixel (deterministic,
 * no atomics), so the projection-weight gradient GEMM runs over pixels, not samples.  rows: fp32 [*, c], or
 * - rows_bf16 != 0 - one bf16 plane (the bf16-backward mode keeps the dZ rows at half the bytes). */
int gx_segment_sum_rows(const void* rows, int rows_bf16, const int* order, const int* seg_off, void* hi, void* lo,
                        float* out_f32, long long nseg, int c, void* stream);

/* Z[b,y,x,:] = sum_l P_l[b, y*h_l/out_h, x*w_l/out_w, :] (fp32 NHWC, c channels).  By linearity the
 * projection of the nearest-upsampled + concatenated per-pixel vector (ref swav_clustering.py:108-130,
 * :171) is the sum of per-level projections computed at each level's native resolution.
 * hi / lo (optional, out may then be NULL): the same values as bf16 split planes [batch*out_h*out_w, c].
 * labels (optional, int64 [batch*out_h*out_w]): first arg-max over the c channels of every pixel, taken from the
 * sums in registers - the label map of predict_swav_codes (ref :691) without a second pass over Z.
 * bilinear != 0: the levels are upsampled like F.interpolate(mode='bilinear', align_corners=False) instead of
 * nearest (swav_args['hf_interp'], ref :112-126) - still linear, so the per-level projection identity holds. */
int gx_upsample_sum(int nlevels, const float* const* p, const int* h, const int* w, int batch, int out_h,
                    int out_w, int c, float* out, void* hi, void* lo, long long* labels, int bilinear, void* stream);

/* Adjoint of 1-D bilinear upsampling (align_corners = False) along the middle axis:
 * in [outer, n_in, inner] -> out [outer, n_out, inner], n_out <= n_in, inner % 4 == 0.  Applied along x and
 * then y it folds dZ onto a coarser level when swav_args['hf_interp'] == 'bilinear'. */
int gx_pool1d_bilinear(const float* in, long long outer, int n_in, int n_out, long long inner, float* out,
                       void* stream);

/* Second half of a 3x3 (dilated, padding = dilation) conv with few output channels computed as one GEMM
 * over all nine taps: g [batch*h*w, 9*cout] with column tap*cout + co (tap = ky*3 + kx);
 * out[b,y,x,co] = act(bias[co] + sum_tap g[(y + (ky-1)*dilation, x + (kx-1)*dilation), tap*cout + co]),
 * zeros outside the image; act 0 none, 2 lrelu(0.2).  out fp32 NHWC and / or bf16 split planes with
 * next_ld channels per pixel for the next layer (caller zero-fills [cout, next_ld)).  cout % 4 == 0.
 * ref: the Conv2d + LeakyReLU layers of OneShotSegmentor, hfc_with_swav/swav_clustering.py:733-742. */
int gx_tap_sum(const float* g, int batch, int h, int w, int cout, int dilation, const float* bias, int act,
               float* out, void* next_hi, void* next_lo, int next_ld, void* stream);

/* Adjoint of gx_tap_sum for the fine-tune backward of those layers: dg[pix, tap*cout + co] =
 * dout[pix - offset(tap), co] (zero outside the image) as bf16 split planes [batch*h*w, 9*cout].
 * Then dW_all [9*cout, cin] = dg^T x (one MN-major split-K GEMM) and dx = dg W_all. */
int gx_tap_spread(const float* dout, int batch, int h, int w, int cout, int dilation, void* dg_hi, void* dg_lo,
                  void* stream);

/* out[b,y,x,:] = sum of the (in_h/out_h x in_w/out_w) block of `in` - the adjoint of nearest upsampling,
 * used to fold dZ onto a level's native resolution; fp32 out and/or bf16 planes (any may be NULL). */
int gx_pool_sum(const float* in, int batch, int in_h, int in_w, int out_h, int out_w, int c, float* out, void* hi,
                void* lo, void* stream);

/* prototype row normalisation in place + split planes (+ transposed planes for dZ),
 * ref: swav_clustering.py:328-331. w [k,c]. */
int gx_normalize_rows(float* w, long long rows, int cols, void* stream);

/* W+ latents of BOTH perturbed views in one launch (ref: swav_clustering.py:593-640 with
 * lib/oneshot/image_augmentor.py:42-53,75-79): out[row, r, :] for row < rows (= 2*b: view s then view t),
 * r < n_latent: v = trunc(w[row % b]); rows 2l, 2l+1 (l = layer_no[row]) are blended with the mapped
 * perturbation draws, v = (1 - sigma[row]) v + sigma[row] noise_w[2 row + (r - 2l)]; then the second
 * truncation (trunc(x) = mean + psi (x - mean), skipped for psi >= 1). */
int gx_view_wplus(const float* w, const float* noise_w, const int* layer_no, const float* sigma, const float* mean,
                  float psi, int b, int rows, int n_latent, int dim, float* out, void* stream);

/* Sampled-pixel bookkeeping of one view (ref: swav_clustering.py:158-167 on the rotated / flipped tensor,
 * :358-359): row_src [patches, bn] = source pixel of every sample inside its image (-1: rotation fill),
 * row_img [bn] = image of the sample.  Outputs: ridx [patches, bn] = row of Z (image*hw + pixel, or -1),
 * and the CSR list of the samples of every pixel: seg_off [npix + 1], order [patches*bn] (first seg_off[npix]
 * entries valid; ascending sample index inside a segment - deterministic).  counts [npix] and tile_scratch
 * [gx_pixel_segments_scratch(npix)] are int scratch.  Several views in one call: patch p belongs to view
 * p / patches_per_group, whose images start at (p / patches_per_group) * img_group_stride (0, 0: one view). */
int gx_pixel_segments_scratch(long long npix);
int gx_pixel_segments(const int* row_src, const int* row_img, int patches, long long bn, int hw, long long npix,
                      int patches_per_group, int img_group_stride, int* ridx, int* counts, int* tile_scratch,
                      int* seg_off, int* order, void* stream);

/* out[k] = (accumulate ? out[k] : 0) + scale * sum_p parts[p,k]  (deterministic order): loss and bias-gradient
 * reductions of the per-CTA partials of gx_swav_loss. */
int gx_colsum(const float* parts, int nparts, int k, float scale, int accumulate, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Sinkhorn-Knopp (ref: hfc_with_swav/swav_clustering.py:509-544)
 *
 * Scaling-vector form: Q = diag(b) * exp(S/eps) * diag(a) (S is [N,K]).  One pass
 * over S per iteration: for every row n: t = sum_k a_k e_nk ; b_n = c_n / t ;
 * u'_k += e_nk * b_n.  Between passes a_k = r_k / u_k.  The final per-pixel
 * normalisation removes b, so the result is q[n,:] = softmax_k(S[n,k]/eps + log a_k).
 * ---------------------------------------------------------------------------------- */

/* Low-latency exchange of the K-vector of column marginals between the GPUs of one box - the only exchange
 * step of the distributed Sinkhorn (ref: swav_clustering.py:519-544 with SwAV's all-reduce of the marginals).
 * Every rank owns one exchange buffer of 8-byte words {fp32 value, u32 sequence number}; all buffers are
 * mapped into every process (gx_peer_*: CUDA IPC).  A block of the buffer is [world][K] words: the sender of
 * rank r stores its K values, tagged with `seq`, into slot r of the block in EVERY rank's buffer (plain 8-byte
 * stores over NVLink, single-copy atomic, no fence); a consumer spins on the tags of its OWN buffer and adds the
 * `world` slots in rank order, so every rank gets the bit-identical sum.  Blocks alternate (parity of the
 * sequence number) so that a fast rank cannot overwrite a slot a slow rank is still reading.
 * world == 0 means "no exchange" wherever a gx_ll_desc is accepted. */
#define GX_MAX_PEERS 16
typedef struct gx_ll_desc {
  void* peers[GX_MAX_PEERS]; /* base of rank r's exchange buffer as mapped in THIS process (own buffer at [rank]) */
  int world, rank;
  long long block_words;     /* offset of the block, in 8-byte words, from the buffer base */
  unsigned int seq;          /* tag of this exchange (non-zero, increases by one per exchange of a channel) */
  int* err;                  /* device int: set to 1 if a consumer gave up waiting (~30 s); may be NULL */
} gx_ll_desc;

/* exchange buffers: cudaMalloc'ed + zeroed / exported as a 64-byte CUDA IPC handle / mapped from a peer's handle */
int gx_peer_alloc(long long bytes, void** ptr);
int gx_peer_free(void* ptr);
int gx_peer_export(void* ptr, void* handle64);
int gx_peer_open(const void* handle64, void** ptr);
int gx_peer_close(void* ptr);

/* One pass.  first != 0: u'_k = sum_n e_nk (a and b are 1).  Otherwise a_k = r_k/u_k with u = u_in, or - when
 * u_ll (may be NULL) has world > 0 - the sum over ranks of the tagged slots of u_ll's block, received in the
 * kernel prologue while the first rows are already in flight.  r == NULL: 1/K; c == NULL: c_n = 1/n_total.
 * reverse != 0 streams the rows last-to-first (alternating the direction from pass to pass turns the tail of S
 * the 126 MB L2 still holds into hits).  Writes per-CTA partials [nparts,K]; returns the number of partials
 * through *nparts_out (host int).  partials must hold at least gx_sinkhorn_max_parts()*K floats. */
int gx_sinkhorn_max_parts(void);
int gx_sinkhorn_pass(const float* s, long long n, int k, long long lds, float inv_eps, int first, const float* u_in,
                     const gx_ll_desc* u_ll, const float* r, const float* c, long long n_total, int reverse,
                     float* partials, int* nparts_out, void* stream);
/* The same pass through a 16-bit cache of the scaled kernel matrix (never a first pass).  write_cache != 0: streams S
 * like gx_sinkhorn_pass and also stores e16[n,lde] = half(2^15 a_k e_nk / sum_k a_k e_nk) and la1[k] = log2 a_k.
 * write_cache == 0: streams e16 instead of S (s may be NULL) - half the bytes and no exponentials: with
 * rho_k = a_k / a1_k the row totals are t'_n = sum_k e16_nk rho_k and u_k = (1/a1_k) sum_n e16_nk c_n / t'_n (the row
 * factor stored in e16 cancels).  The marginals differ from the fp32 pass by the rounding of e16 (2^-11 per term,
 * averaged over a column: <~ 3e-4 in the final codes); the codes themselves are always computed from
 * the fp32 scores (gx_swav_loss / gx_sinkhorn_q).  lde % 8 == 0 (16-byte rows), e16 and la1 16-byte aligned.
 * Meant for uniform (or mildly non-uniform) prototype marginals r: a column whose target mass is below ~1e-8 of the
 * typical one falls under the fp16 range of the row-normalised plane - the host layer uses the fp32 pass for
 * source_pdf == 'image', where an empty histogram bin has a mass of 1e-9 counts. */
int gx_sinkhorn_pass_cached(const float* s, long long n, int k, long long lds, float inv_eps, const float* u_in,
                            const gx_ll_desc* u_ll, const float* r, const float* c, long long n_total, int reverse,
                            float* partials, int* nparts_out, void* e16, long long lde, float* la1, int write_cache,
                            void* stream);
/* u[k] = sum_p partials[p,k] (deterministic order). */
int gx_sinkhorn_reduce(const float* partials, int nparts, int k, float* u, void* stream);
/* The same column sums, pushed as tagged words into slot `rank` of ll's block on every rank (fused reduce +
 * exchange over NVLink peer memory; replaces reduce + NCCL all-reduce).  u_local (optional): the local sums. */
int gx_sinkhorn_reduce_send(const float* partials, int nparts, int k, const gx_ll_desc* ll, float* u_local,
                            void* stream);
/* u[k] = sum over ranks of ll's block (waits for the tags). */
int gx_ll_recv_sum(const gx_ll_desc* ll, int k, float* u, void* stream);
/* log_a[k] = log(r_k / u[k]); u from u (u_ll NULL / world 0) or received from u_ll's block. */
int gx_sinkhorn_log_a(const float* u, const gx_ll_desc* u_ll, const float* r, int k, float* log_a, void* stream);
/* Materialise Q [N,K] = softmax_k(S/eps + log_a) (API parity with sinkhorn_knopp's return). */
int gx_sinkhorn_q(const float* s, long long n, int k, long long lds, float inv_eps, const float* log_a, float* q,
                  void* stream);

/* ------------------------------------------------------------------------------------
 * Swapped-prediction loss, forward + d/dscores fused
 * (ref: hfc_with_swav/swav_clustering.py:547-570, p = S/temperature, q from Sinkhorn)
 * loss_sum += sum_n -0.5*(sum_k q_s*logsoftmax(p_t) + sum_k q_t*logsoftmax(p_s))   (not yet / N)
 * dS_t = grad_scale * (softmax(p_t) - q_s) / (2*T), dS_s likewise; written as bf16 planes.
 * grad_scale = 1 / (N_total * num_patches).  loss_parts: [gx_loss_max_parts()] per-CTA sums;
 * db_parts (optional): [gx_loss_max_parts(), K] per-CTA column sums of dS_s + dS_t (the
 * prototype-bias gradient), reduced with gx_sinkhorn_reduce.
 * ---------------------------------------------------------------------------------- */
int gx_loss_max_parts(void);
int gx_swav_loss(const float* s_s, const float* s_t, long long n, int k, long long lds, float inv_eps,
                 float inv_temp, const float* log_a_s, const float* log_a_t, float grad_scale, float* loss_parts,
                 float* db_parts, int* nparts_out, void* ds_s_hi, void* ds_s_lo, void* ds_t_hi, void* ds_t_lo,
                 long long ldd, float* ds_s_f32, float* ds_t_f32, void* stream);

/* ------------------------------------------------------------------------------------
 * Generic convolution pieces of the conv2d_gradfix drop-in (ref: lib/gan/optim/conv2d_gradfix.py:129-270):
 * conv2d = im2col + gx_gemm, conv_transpose2d = gx_gemm + col2im; the two maps are an adjoint pair, so gradients
 * of any order are made of the same three calls.  x [b,c,h,w] fp32 NCHW; cols [b*ho*wo, ld] fp32, column
 * (ci*kh + ky)*kw + kx, ld >= c*kh*kw (extra columns zero-filled by im2col, ignored by col2im).
 * ---------------------------------------------------------------------------------- */
int gx_im2col(const float* x, int b, int c, int h, int w, int kh, int kw, int sy, int sx, int py, int px, int dy,
              int dx, int ho, int wo, long long ld, float* cols, void* stream);
int gx_col2im(const float* cols, int b, int c, int h, int w, int kh, int kw, int sy, int sx, int py, int px, int dy,
              int dx, int ho, int wo, long long ld, float* x, void* stream);

/* ------------------------------------------------------------------------------------
 * SimCLR baseline head (ref: baseline/hfc_with_simclr/simclr_clustering.py:133-281, 362-401):
 * Linear(no bias) -> BatchNorm1d -> LeakyReLU(0.01) -> Linear(no bias) on channel-normalised per-pixel vectors.
 * The two Linear layers are gx_gemm calls; these are the pieces in between.  hraw [n,c] (row pitch ldh) is the
 * first GEMM's output on the UN-normalised rows; rscale[n] = 1 / max(|row|, 1e-12) is F.normalize folded in
 * (h = hraw * rscale; NULL = 1).
 * ---------------------------------------------------------------------------------- */
/* out[i] = 1 / max(x[i], eps) (mode 0: the 1/|f| of F.normalize from the row norms of gx_gather_rows) or
 * rsqrt(x[i] + eps) (mode 1: the eval-mode BatchNorm scale from the running variance). */
int gx_recip(const float* x, long long n, float eps, int mode, float* out, void* stream);
/* training-mode batch statistics over the n rows: mean[c], invstd[c] = rsqrt(biased var + eps); optionally the
 * running statistics update run = (1 - momentum) run + momentum (mean, unbiased var)  (nn.BatchNorm1d). */
int gx_bn_stats(const float* hraw, long long ldh, const float* rscale, int n, int c, float eps, float* mean,
                float* invstd, float* run_mean, float* run_var, float momentum, void* stream);
/* a = lrelu_slope((h - mean) * invstd * gamma + beta): fp32 out [n,c] and/or split-bf16 planes (any may be NULL). */
int gx_bn_act_apply(const float* hraw, long long ldh, const float* rscale, long long n, int c, const float* mean,
                    const float* invstd, const float* gamma, const float* beta, float slope, float* out, void* hi,
                    void* lo, void* stream);
/* backward through lrelu(bn(.)) with batch statistics: dhs[n,c] = dL/dhraw (= dL/dh * rscale), dgamma[c], dbeta[c]. */
int gx_bn_act_bwd(const float* da, const float* hraw, long long ldh, const float* rscale, int n, int c,
                  const float* mean, const float* invstd, const float* gamma, const float* beta, float slope,
                  float* dhs, float* dgamma, float* dbeta, void* stream);
/* The reference's contrastive loss on the projection output z [n2, c] (n2 = 2 * batch_size <= 64 rows interleaved
 * s_0, t_0, s_1, ...; ref :235-265 with both of its quirks, see oracle simclr_loss) and its gradient dz [n2, c]. */
int gx_simclr_loss(const float* z, int n2, int c, float inv_temperature, float* loss, float* dz, void* stream);

/* ------------------------------------------------------------------------------------
 * Optimiser: apex LARC(clip=False) around SGD(momentum)
 * (ref: swav_clustering.py:286-292,458-460).  norms: scratch of gx_larc_scratch_floats() floats (per-block partial
 * sums of |p|^2, |g|^2, folded in a fixed order: the trust ratio is deterministic, so data-parallel replicas
 * that apply the same all-reduced gradient stay bit-identical).  first_step != 0: momentum buffer := g.
 * ---------------------------------------------------------------------------------- */
int gx_larc_scratch_floats(void);
int gx_larc_sgd(float* p, const float* g, float* buf, long long n, float lr, float momentum, float trust,
                float weight_decay, float eps, int first_step, float* norms, void* stream);

/* ------------------------------------------------------------------------------------
 * Inference
 * ---------------------------------------------------------------------------------- */
/* labels[n] = first argmax_c x[n,c]  (ref: out_preds.max(1)[1], swav_clustering.py:691). int64 out. */
int gx_argmax_rows(const float* x, long long n, int c, long long ldx, long long* labels, void* stream);

/* k-means assignment: labels[n] = first argmin_k ||x_n - c_k||^2 with x_n = concat(x1[n,:c1], x2[n,:c2])
 * (the two same-resolution maps the reference concatenates, image_augmentor.py:80-90; x2 may be NULL)
 * (ref: clusterer.predict, baseline/hfc_kmeans/hfc_kmeans_clustering.py:184).  centers [k,c1+c2]; int32 out.
 * dist (optional, labels may then be NULL): squared distance to the assigned centre - inertia and the
 * k-means++ potentials of the Lloyd fit (ref: KMeans.fit, hfc_kmeans_clustering.py:146-166). */
int gx_kmeans_assign(const float* x1, int c1, const float* x2, int c2, long long n, const float* centers, int k,
                     int* labels, float* dist, void* stream);

/* labels[n] = first argmin_k (bias[k] + scale * s[n,k]).  With s = X C^T from gx_gemm (3-pass split-bf16), bias =
 * ||c_k||^2 and scale = -2 this is the same nearest-centre assignment on the tensor cores (the GEMM form of the
 * distance that scikit-learn's predict uses, ref: hfc_kmeans_clustering.py:184). int32 out. */
int gx_argmin_affine(const float* s, long long n, int k, long long lds, const float* bias, float scale, int* labels,
                     void* stream);

/* The same assignment for few centres (k <= 64) and long fp32 rows (c1, c2 multiples of 16), fused: the feature rows
 * are read once (4 B per element, no operand planes, no score matrix in HBM), split into bf16 hi / lo in registers and
 * contracted against centre fragments resident in shared memory on the tensor cores (three bf16 MMAs per product:
 * fp32-grade), arg-min of ||c_k||^2 - 2 x.c_k in the epilogue.
 *   gx_kmeans_frag_bytes:   size of the centre-fragment buffer of a [k, c] centre matrix; 0 = shape not supported
 *                           (k > 64, c % 16, or fragments larger than shared memory) - use gx_gemm + gx_argmin_affine.
 *   gx_kmeans_center_frags: builds the fragments from fp32 centres [k, c] (once per fitted model).
 *   gx_kmeans_assign_mma:   cn_pad = ||c_k||^2 padded with +inf to 64 entries; int32 labels out.
 * (ref: clusterer.predict, baseline/hfc_kmeans/hfc_kmeans_clustering.py:184) */
long long gx_kmeans_frag_bytes(int k, int c);
int gx_kmeans_center_frags(const float* centers, int k, int c, void* frags, void* stream);
int gx_kmeans_assign_mma(const float* x1, int c1, const float* x2, int c2, long long n, const void* frags,
                         const float* cn_pad, int k, int* labels, void* stream);

/* one-hot cluster maps [b,k,out_h,out_w] from labels [b,h,w], nearest-neighbour resize
 * (ref: hfc_kmeans_clustering.py:190-206).  out_batch_stride (floats; 0 = k*out_h*out_w): a layer can write its K
 * channels straight into the concatenated [B, sum K, out_h, out_w] maps (16-byte aligned when out_w % 4 == 0).
 * on_value / off_value: (1, 0) = the maps of clusterer.predict; (1, -1) = the `hier_preds * 2 - 1` encoding that
 * predict_hfc_vectors returns (ref: baseline/hfc_kmeans/segmentor.py:222-226) without a second pass. */
int gx_onehot_nearest(const int* labels, int b, int h, int w, int k, int out_h, int out_w, float* out,
                      long long out_batch_stride, float on_value, float off_value, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GANECDOTES_B200_H */
