// SimCLR baseline head (ref: baseline/hfc_with_simclr/simclr_clustering.py:133-281, 362-401): the pieces between the
// tensor-core GEMMs - BatchNorm1d (+ the per-pixel 1/|f| of F.normalize folded in as a row scale) + LeakyReLU,
// its backward, and the contrastive loss with its gradient.  The training batch is 2 x batch_size = 40 rows, so
// these are latency-sized SIMT kernels; the inference path (every pixel) uses the element-wise apply kernel.
#include "gx_common.cuh"

namespace {

// h[n,c] = hraw[n,c] * rscale[n]; per-channel batch statistics over the n rows (thread per channel, rows strided by
// ldh: coalesced across the warp).  Also folds the running statistics: run = (1 - mom) run + mom stat (unbiased var).
__global__ void bn_stats_kernel(const float* __restrict__ hraw, long long ldh, const float* __restrict__ rscale, int n,
                                int c, float eps, float* __restrict__ mean_out, float* __restrict__ invstd_out,
                                float* __restrict__ run_mean, float* __restrict__ run_var, float momentum) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float s = 0.f;
  for (int r = 0; r < n; ++r) s += hraw[r * ldh + ch] * (rscale ? rscale[r] : 1.f);
  const float mean = s / (float)n;
  float v = 0.f;
  for (int r = 0; r < n; ++r) {
    const float d = hraw[r * ldh + ch] * (rscale ? rscale[r] : 1.f) - mean;
    v = fmaf(d, d, v);
  }
  mean_out[ch] = mean;
  invstd_out[ch] = rsqrtf(v / (float)n + eps);                    // biased variance normalises (training mode)
  if (run_mean) {
    run_mean[ch] = (1.f - momentum) * run_mean[ch] + momentum * mean;
    run_var[ch] = (1.f - momentum) * run_var[ch] + momentum * (n > 1 ? v / (float)(n - 1) : v);
  }
}

// a = lrelu((h - mean) * invstd * gamma + beta), h = hraw * rscale[row]; fp32 out and/or split-bf16 planes
__global__ void bn_apply_kernel(const float* __restrict__ hraw, long long ldh, const float* __restrict__ rscale,
                                long long n, int c, const float* __restrict__ mean, const float* __restrict__ invstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta, float slope,
                                float* __restrict__ out, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int cq = c >> 2;
  const long long total = n * cq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cq;
    const int q = (int)(i - r * cq) * 4;
    const float rs = rscale ? rscale[r] : 1.f;
    const float4 x = *reinterpret_cast<const float4*>(hraw + r * ldh + q);
    const float4 m = __ldg(reinterpret_cast<const float4*>(mean + q));
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd + q));
    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + q));
    const float4 b = __ldg(reinterpret_cast<const float4*>(beta + q));
    float4 y;
    y.x = fmaf((x.x * rs - m.x) * is.x, g.x, b.x);
    y.y = fmaf((x.y * rs - m.y) * is.y, g.y, b.y);
    y.z = fmaf((x.z * rs - m.z) * is.z, g.z, b.z);
    y.w = fmaf((x.w * rs - m.w) * is.w, g.w, b.w);
    y.x = y.x > 0.f ? y.x : y.x * slope;
    y.y = y.y > 0.f ? y.y : y.y * slope;
    y.z = y.z > 0.f ? y.z : y.z * slope;
    y.w = y.w > 0.f ? y.w : y.w * slope;
    if (out) *reinterpret_cast<float4*>(out + r * c + q) = y;
    if (hi) {
      uint2 h2, l2;
      gx_split4(y, h2, l2);
      *reinterpret_cast<uint2*>(hi + r * c + q) = h2;
      if (lo) *reinterpret_cast<uint2*>(lo + r * c + q) = l2;
    }
  }
}

// backward of lrelu(bn(h)) for the training batch: thread per channel.
// dhs[n,c] = dL/dh * rscale[n] (the gradient w.r.t. hraw), dgamma, dbeta.
__global__ void bn_bwd_kernel(const float* __restrict__ da, const float* __restrict__ hraw, long long ldh,
                              const float* __restrict__ rscale, int n, int c, const float* __restrict__ mean,
                              const float* __restrict__ invstd, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float slope, float* __restrict__ dhs,
                              float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const float m = mean[ch], is = invstd[ch], g = gamma[ch], b = beta[ch];
  float s1 = 0.f, s2 = 0.f;                       // sum dy, sum dy * hn   (dy = da * lrelu'(pre))
  for (int r = 0; r < n; ++r) {
    const float hn = (hraw[r * ldh + ch] * (rscale ? rscale[r] : 1.f) - m) * is;
    const float pre = fmaf(hn, g, b);
    const float dy = da[(long long)r * c + ch] * (pre > 0.f ? 1.f : slope);
    s1 += dy;
    s2 = fmaf(dy, hn, s2);
  }
  dgamma[ch] = s2;
  dbeta[ch] = s1;
  const float inv_n = 1.f / (float)n;
  for (int r = 0; r < n; ++r) {
    const float rs = rscale ? rscale[r] : 1.f;
    const float hn = (hraw[r * ldh + ch] * rs - m) * is;
    const float pre = fmaf(hn, g, b);
    const float dy = da[(long long)r * c + ch] * (pre > 0.f ? 1.f : slope);
    const float dh = g * is * (dy - inv_n * s1 - hn * inv_n * s2);
    dhs[(long long)r * c + ch] = dh * rs;
  }
}

// The reference's contrastive loss (its two O(n^2) loops, :235-265, both quirks kept - see the oracle's simclr_loss):
// vectors v_i = z[:, i] for the first n2 CHANNELS (each a vector over the n2 samples), sim = cos(v_i, v_j) / T,
// l[i,j] = -sim_ij + log sum_{m != i} exp(sim_im), loss = sum_k (l[a_k, 2k] + l[2k, a_k]) / n2, a_k = (2k - 1) mod n2.
// One block; n2 <= 64.  dz[n2, c] gets the gradient (zero outside the first n2 channels).
constexpr int SC_MAX = 64;

// dL/dsim[p][q] * n2: every index is the row of exactly one pair term (a_k runs over the odd, 2k over the even indices)
__device__ __forceinline__ float simclr_gsim(int p, int q, int n2, float sim_pq, float den_p) {
  if (p == q) return 0.f;
  float g = expf(sim_pq) / den_p;
  if ((q & 1) == 0 && p == (q - 1 + n2) % n2) g -= 1.f;       // l[a_k, 2k]
  if ((p & 1) == 0 && q == (p - 1 + n2) % n2) g -= 1.f;       // l[2k, a_k]
  return g;
}

__global__ void __launch_bounds__(256)
simclr_loss_kernel(const float* __restrict__ z, int n2, int c, float inv_t, float* __restrict__ loss,
                   float* __restrict__ dz) {
  __shared__ float v[SC_MAX][SC_MAX + 1];        // v[i][s] = z[s, i]
  __shared__ float sim[SC_MAX][SC_MAX + 1];
  __shared__ float nrm[SC_MAX], den[SC_MAX];
  const int tid = threadIdx.x;
  for (int i = tid; i < n2 * n2; i += blockDim.x) {
    const int s = i / n2, ch = i - s * n2;
    v[ch][s] = z[(long long)s * c + ch];
  }
  for (long long i = tid; i < (long long)n2 * c; i += blockDim.x) dz[i] = 0.f;
  __syncthreads();
  if (tid < n2) {
    float s = 0.f;
    for (int k = 0; k < n2; ++k) s = fmaf(v[tid][k], v[tid][k], s);
    nrm[tid] = sqrtf(s);
  }
  __syncthreads();
  for (int i = tid; i < n2 * n2; i += blockDim.x) {
    const int p = i / n2, q = i - p * n2;
    float d = 0.f;
    for (int k = 0; k < n2; ++k) d = fmaf(v[p][k], v[q][k], d);
    sim[p][q] = d / fmaxf(nrm[p] * nrm[q], 1e-8f) * inv_t;
  }
  __syncthreads();
  if (tid < n2) {
    float s = 0.f;
    for (int m = 0; m < n2; ++m)
      if (m != tid) s += expf(sim[tid][m]);
    den[tid] = s;
  }
  __syncthreads();
  const float inv_n = 1.f / (float)n2;
  if (tid == 0) {
    float part = 0.f;
    for (int k = 0; k < n2 / 2; ++k) {
      const int a = (2 * k - 1 + n2) % n2, b = 2 * k;
      part += (-sim[a][b] + logf(den[a])) + (-sim[b][a] + logf(den[b]));
    }
    loss[0] = part * inv_n;
  }
  // dv_p[k] = sum_q (G_pq + G_qp) * d sim_pq / d v_p[k];  thread per (p, k)
  for (int i = tid; i < n2 * n2; i += blockDim.x) {
    const int p = i / n2, k = i - p * n2;
    float acc = 0.f;
    const float np_ = nrm[p];
    for (int q = 0; q < n2; ++q) {
      if (q == p) continue;
      const float nn = np_ * nrm[q];
      if (nn < 1e-8f) continue;                   // clamped denominator: the cosine no longer depends on the norms
      const float w = inv_n * (simclr_gsim(p, q, n2, sim[p][q], den[p]) + simclr_gsim(q, p, n2, sim[q][p], den[q]));
      const float dot = sim[p][q] * nn / inv_t;   // v_p . v_q
      acc += w * inv_t * (v[q][k] / nn - dot * v[p][k] / (np_ * np_ * nn));
    }
    dz[(long long)k * c + p] = acc;               // v[p][k] = z[k, p]
  }
}

// out = 1 / max(x, eps)  (mode 0: the 1/|f| of F.normalize)   or   rsqrt(x + eps)  (mode 1: BatchNorm eval)
__global__ void recip_kernel(const float* __restrict__ x, long long n, float eps, int mode, float* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = mode == 0 ? 1.f / fmaxf(x[i], eps) : rsqrtf(x[i] + eps);
}

}  // namespace

extern "C" int gx_recip(const float* x, long long n, float eps, int mode, float* out, void* stream) {
  GX_CHECK_ARG(x && out && n > 0 && (mode == 0 || mode == 1));
  recip_kernel<<<(int)min((long long)gx_sm_count() * 8, (n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, n, eps, mode,
                                                                                                       out);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_bn_stats(const float* hraw, long long ldh, const float* rscale, int n, int c, float eps, float* mean,
                           float* invstd, float* run_mean, float* run_var, float momentum, void* stream) {
  GX_CHECK_ARG(hraw && mean && invstd && n > 0 && c > 0 && ldh >= c && ((run_mean == nullptr) == (run_var == nullptr)));
  bn_stats_kernel<<<gx_cdiv(c, 128), 128, 0, (cudaStream_t)stream>>>(hraw, ldh, rscale, n, c, eps, mean, invstd,
                                                                    run_mean, run_var, momentum);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_bn_act_apply(const float* hraw, long long ldh, const float* rscale, long long n, int c,
                               const float* mean, const float* invstd, const float* gamma, const float* beta,
                               float slope, float* out, void* hi, void* lo, void* stream) {
  GX_CHECK_ARG(hraw && mean && invstd && gamma && beta && (out || hi) && n > 0 && c > 0 && c % 4 == 0 && ldh % 4 == 0);
  const long long total = n * (c >> 2);
  const int grid = (int)min((long long)gx_sm_count() * 8, (total + 255) / 256);
  bn_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(hraw, ldh, rscale, n, c, mean, invstd, gamma, beta, slope, out,
                                                          reinterpret_cast<__nv_bfloat16*>(hi),
                                                          reinterpret_cast<__nv_bfloat16*>(lo));
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_bn_act_bwd(const float* da, const float* hraw, long long ldh, const float* rscale, int n, int c,
                             const float* mean, const float* invstd, const float* gamma, const float* beta, float slope,
                             float* dhs, float* dgamma, float* dbeta, void* stream) {
  GX_CHECK_ARG(da && hraw && mean && invstd && gamma && beta && dhs && dgamma && dbeta && n > 0 && c > 0 && ldh >= c);
  bn_bwd_kernel<<<gx_cdiv(c, 128), 128, 0, (cudaStream_t)stream>>>(da, hraw, ldh, rscale, n, c, mean, invstd, gamma, beta,
                                                                  slope, dhs, dgamma, dbeta);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_simclr_loss(const float* z, int n2, int c, float inv_temperature, float* loss, float* dz,
                              void* stream) {
  GX_CHECK_ARG(z && loss && dz && n2 >= 2 && n2 % 2 == 0 && n2 <= SC_MAX && c >= n2);
  simclr_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(z, n2, c, inv_temperature, loss, dz);
  GX_LAUNCH_CHECK();
  return GX_OK;
}
