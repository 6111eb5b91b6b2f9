// Small StyleGAN2 synthesis helpers: mapping network pieces, modulation / demodulation
// coefficients, weight preparation, activation modulate+split.
#include "gx_common.cuh"

namespace {

__global__ void pixel_norm_kernel(const float* __restrict__ x, float* __restrict__ y, int n, int dim) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* xr = x + (long long)row * dim;
  float ss = 0.f;
  for (int i = lane; i < dim; i += 32) ss = fmaf(xr[i], xr[i], ss);
  ss = gx_warp_sum(ss);
  const float r = rsqrtf(ss / (float)dim + 1e-8f);
  for (int i = lane; i < dim; i += 32) y[(long long)row * dim + i] = xr[i] * r;
}

// one warp per output element (row, o); 8 rows share one weight row per block
__global__ void equal_linear_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ w,
                                    const float* __restrict__ b, float* __restrict__ y, int n, int in_dim, int out_dim,
                                    float w_scale, float b_scale, int act) {
  const int lane = threadIdx.x & 31;
  const int wi = threadIdx.x >> 5;               // warp in block -> row offset
  const int o = blockIdx.x;
  const int row = blockIdx.y * (blockDim.x >> 5) + wi;
  if (row >= n) return;
  const float* xr = x + (long long)row * ldx;
  const float* wr = w + (long long)o * in_dim;
  float acc = 0.f;
  for (int i = lane; i < in_dim; i += 32) acc = fmaf(xr[i], wr[i] * w_scale, acc);
  acc = gx_warp_sum(acc);
  if (lane == 0) {
    if (b) acc += b[o] * b_scale;
    if (act) acc = (acc > 0.f ? acc : acc * 0.2f) * 1.41421356237309515f;
    y[(long long)row * out_dim + o] = acc;
  }
}

__global__ void truncate_kernel(const float* __restrict__ w, const float* __restrict__ mean, float* __restrict__ out,
                                long long total, int dim, float psi) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float m = mean[i % dim];
  out[i] = m + psi * (w[i] - m);
}

__global__ void modconv_prepare_kernel(const float* __restrict__ w, float scale, __nv_bfloat16* __restrict__ w_hi,
                                       __nv_bfloat16* __restrict__ w_lo, float* __restrict__ wsq, int cout, int cin,
                                       int cin_ld, int kk) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)cout * cin_ld) return;
  const int ci = (int)(idx % cin_ld);
  const int co = (int)(idx / cin_ld);
  float ss = 0.f;
  for (int t = 0; t < kk; ++t) {
    const float v = ci < cin ? w[((long long)co * cin + ci) * kk + t] * scale : 0.f;
    ss = fmaf(v, v, ss);
    __nv_bfloat16 h, l;
    gx_split_bf16(v, h, l);
    const long long o = ((long long)co * kk + t) * cin_ld + ci;
    w_hi[o] = h;
    if (w_lo) w_lo[o] = l;
  }
  if (wsq && ci < cin) wsq[(long long)co * cin + ci] = ss;
}

__global__ void modconv_demod_kernel(const float* __restrict__ wsq, const float* __restrict__ s,
                                     float* __restrict__ demod, int batch, int cin, int cout) {
  const int lane = threadIdx.x & 31;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (long long)batch * cout) return;
  const int co = (int)(wid % cout);
  const int b = (int)(wid / cout);
  const float* wr = wsq + (long long)co * cin;
  const float* sr = s + (long long)b * cin;
  float acc = 0.f;
  for (int i = lane; i < cin; i += 32) acc = fmaf(wr[i], sr[i] * sr[i], acc);
  acc = gx_warp_sum(acc);
  if (lane == 0) demod[wid] = rsqrtf(acc + 1e-8f);
}

__global__ void modulate_split_kernel(const float* __restrict__ x, long long xbs, const float* __restrict__ s,
                                      __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int batch,
                                      long long hw, int c, int c_ld) {
  const int cq = c >> 2;
  const long long total = (long long)batch * hw * cq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cq);
    const long long pix = i / cq;
    const int b = (int)(pix / hw);
    const long long p = pix - (long long)b * hw;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + (long long)b * xbs + p * c) + q);
    const float4 sv = __ldg(reinterpret_cast<const float4*>(s + (long long)b * c) + q);
    uint2 h, l;
    gx_split4(make_float4(v.x * sv.x, v.y * sv.y, v.z * sv.z, v.w * sv.w), h, l);
    const long long o = pix * (c_ld >> 2) + q;
    reinterpret_cast<uint2*>(hi)[o] = h;
    if (lo) reinterpret_cast<uint2*>(lo)[o] = l;
  }
}

}  // namespace

extern "C" int gx_pixel_norm(const float* x, float* y, int n, int dim, void* stream) {
  GX_CHECK_ARG(x && y && n > 0 && dim > 0);
  pixel_norm_kernel<<<gx_cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(x, y, n, dim);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_equal_linear(const float* x, long long ldx, const float* w, const float* b, float* y, int n,
                               int in_dim, int out_dim, float w_scale, float b_scale, int act, void* stream) {
  GX_CHECK_ARG(x && w && y && n > 0 && in_dim > 0 && out_dim > 0 && ldx >= in_dim);
  GX_CHECK_ARG(gx_cdiv(n, 8) <= 65535);
  dim3 grid(out_dim, gx_cdiv(n, 8));
  equal_linear_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ldx, w, b, y, n, in_dim, out_dim, w_scale, b_scale,
                                                              act);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_truncate(const float* w, const float* mean, float* out, long long rows, int dim, float psi,
                           void* stream) {
  GX_CHECK_ARG(w && mean && out && rows > 0 && dim > 0);
  const long long total = rows * dim;
  truncate_kernel<<<gx_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(w, mean, out, total, dim, psi);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_modconv_prepare(const float* w, float scale, void* w_hi, void* w_lo, float* wsq, int cout, int cin,
                                  int cin_ld, int k, void* stream) {
  GX_CHECK_ARG(w && w_hi && cout > 0 && cin > 0 && k > 0);
  if (cin_ld <= 0) cin_ld = cin;
  GX_CHECK_ARG(cin_ld >= cin);
  const long long total = (long long)cout * cin_ld;
  modconv_prepare_kernel<<<gx_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(
      w, scale, reinterpret_cast<__nv_bfloat16*>(w_hi), reinterpret_cast<__nv_bfloat16*>(w_lo), wsq, cout, cin, cin_ld,
      k * k);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_modconv_demod(const float* wsq, const float* s, float* demod, int batch, int cin, int cout,
                                void* stream) {
  GX_CHECK_ARG(wsq && s && demod && batch > 0 && cin > 0 && cout > 0);
  const long long warps = (long long)batch * cout;
  modconv_demod_kernel<<<gx_cdiv(warps, 8), 256, 0, (cudaStream_t)stream>>>(wsq, s, demod, batch, cin, cout);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_modulate_split(const float* x, long long x_batch_stride, const float* s, void* hi, void* lo,
                                 int batch, long long hw, int c, int c_ld, void* stream) {
  GX_CHECK_ARG(x && s && hi && batch > 0 && hw > 0 && c % 4 == 0);
  if (c_ld <= 0) c_ld = c;
  GX_CHECK_ARG(c_ld >= c && c_ld % 4 == 0);
  const long long total = (long long)batch * hw * (c / 4);
  int grid = gx_cdiv(total, 256);
  const int cap = gx_sm_count() * 16;
  if (grid > cap) grid = cap;
  modulate_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, x_batch_stride, s,
                                                               reinterpret_cast<__nv_bfloat16*>(hi),
                                                               reinterpret_cast<__nv_bfloat16*>(lo), batch, hw, c, c_ld);
  GX_LAUNCH_CHECK();
  return GX_OK;
}
