// Memory-bound stages of the SwAV head: per-pixel feature gather, L2 normalisation,
// Sinkhorn-Knopp passes, fused swapped-prediction loss fwd+bwd, LARC+SGD, arg-max label
// maps and k-means assignment.  All kernels use 128-bit coalesced accesses and
// warp-shuffle + shared-memory reductions; none of them is reshaped into a GEMM.
#include <stdlib.h>
#include <string.h>

#include "gx_common.cuh"
#include "gx_ll.cuh"
#include "gx_ptx.cuh"

namespace {

constexpr float LOG2E = 1.4426950408889634f;

// ---------------------------------------------------------------------------
// gather: rotate/flip + random-pixel sampling + nearest upsample + concat + [:hlen]
// one 128-thread block per output row (= one sampled pixel), 16 B per thread per access
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gather_rows_kernel(const gx_gather_desc d) {
  const long long r = blockIdx.x;
  if (r >= d.nrows) return;
  int img, src;
  if (d.row_src) {
    img = d.row_img ? d.row_img[r] : 0;
    src = d.row_src[r];
  } else {
    const long long per = (long long)d.out_h * d.out_w;
    img = (int)(r / per);
    src = (int)(r - (long long)img * per);
  }
  const int sy = src >= 0 ? src / d.out_w : 0;
  const int sx = src >= 0 ? src - sy * d.out_w : 0;
  uint2* ah = d.a_hi ? reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(d.a_hi) + r * d.ld) : nullptr;
  uint2* al = d.a_lo ? reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(d.a_lo) + r * d.ld) : nullptr;
  float4* af = d.a_f32 ? reinterpret_cast<float4*>(d.a_f32 + r * d.ld) : nullptr;
  float ss = 0.f;
  int off = 0;
  for (int l = 0; l < d.nlevels && off < d.hlen; ++l) {
    const int cl = d.c[l];
    const int keep = min(cl, d.hlen - off);
    // F.interpolate(mode='nearest'): src = floor(dst * in / out)
    const int ly = (int)(((long long)sy * d.h[l]) / d.out_h);
    const int lx = (int)(((long long)sx * d.w[l]) / d.out_w);
    const float4* f = reinterpret_cast<const float4*>(
        d.feat[l] + (((long long)img * d.h[l] + ly) * d.w[l] + lx) * cl);
    for (int q = threadIdx.x; q < (keep >> 2); q += blockDim.x) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (src >= 0) v = __ldg(f + q);
      const int o = (off >> 2) + q;
      if (ah) {
        uint2 h, lo;
        gx_split4(v, h, lo);
        ah[o] = h;
        if (al) al[o] = lo;
      }
      if (af) af[o] = v;
      ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    off += keep;
  }
  // zero the padding columns [hlen, ld)
  for (int q = (d.hlen >> 2) + threadIdx.x; q < (int)(d.ld >> 2); q += blockDim.x) {
    if (ah) ah[q] = make_uint2(0u, 0u);
    if (al) al[q] = make_uint2(0u, 0u);
    if (af) af[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (d.row_norm) {   // ||row||_2 (ref: torch.norm(hfeat, p=2, dim=1), swav_clustering.py:361-362)
    __shared__ float red[4];
    ss = gx_warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) d.row_norm[r] = sqrtf(red[0] + red[1] + red[2] + red[3]);
  }
}

// ---------------------------------------------------------------------------
// L2 normalisation of projected rows (warp per row)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint2 f16x4(const float4 v) {
  const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
}

__global__ void l2norm_split_kernel(const float* __restrict__ z, const int* __restrict__ row_idx,
                                    __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                                    __half* __restrict__ f16, float* __restrict__ inv_norm, long long n, int c) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int cq = c >> 2;
  uint2* oh = reinterpret_cast<uint2*>(hi + row * c);
  uint2* ol = lo ? reinterpret_cast<uint2*>(lo + row * c) : nullptr;
  uint2* of = f16 ? reinterpret_cast<uint2*>(f16 + row * c) : nullptr;
  // optional gather: output row `row` normalises input row row_idx[row]; -1 = an all-zero row
  // (rotation fill), whose normalisation is 0 (F.normalize clamps the norm at 1e-12)
  const long long src = row_idx ? (long long)row_idx[row] : row;
  if (src < 0) {
    if (lane == 0 && inv_norm) inv_norm[row] = 1e12f;
    for (int i = lane; i < cq; i += 32) {
      oh[i] = make_uint2(0u, 0u);
      if (ol) ol[i] = make_uint2(0u, 0u);
      if (of) of[i] = make_uint2(0u, 0u);
    }
    return;
  }
  const float4* zr = reinterpret_cast<const float4*>(z + src * c);
  float ss = 0.f;
  for (int i = lane; i < cq; i += 32) {
    const float4 v = __ldg(zr + i);
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = gx_warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  if (lane == 0 && inv_norm) inv_norm[row] = inv;
  for (int i = lane; i < cq; i += 32) {
    const float4 v = __ldg(zr + i);
    uint2 h, l;
    const float4 zn = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    gx_split4(zn, h, l);
    oh[i] = h;
    if (ol) ol[i] = l;
    if (of) of[i] = f16x4(zn);
  }
}

__global__ void round_f16_kernel(const float* __restrict__ x, long long ld, __half* __restrict__ out, long long rows,
                                 long long cols) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    out[i] = __float2half_rn(x[r * ld + c]);
  }
}

__device__ __forceinline__ float4 planes_to_float4(uint2 h, const uint2* lo_ptr) {
  float4 v;
  v.x = __uint_as_float(h.x << 16);
  v.y = __uint_as_float(h.x & 0xffff0000u);
  v.z = __uint_as_float(h.y << 16);
  v.w = __uint_as_float(h.y & 0xffff0000u);
  if (lo_ptr) {
    const uint2 l = *lo_ptr;
    v.x += __uint_as_float(l.x << 16);
    v.y += __uint_as_float(l.x & 0xffff0000u);
    v.z += __uint_as_float(l.y << 16);
    v.w += __uint_as_float(l.y & 0xffff0000u);
  }
  return v;
}

__global__ void l2norm_bwd_split_kernel(const float* __restrict__ dzn, const __nv_bfloat16* __restrict__ zh,
                                        const __nv_bfloat16* __restrict__ zl, const float* __restrict__ inv_norm,
                                        __nv_bfloat16* __restrict__ dh, __nv_bfloat16* __restrict__ dl,
                                        float* __restrict__ dz_f32, long long n, int c) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int cq = c >> 2;
  const float4* gr = reinterpret_cast<const float4*>(dzn + row * c);
  const uint2* hr = reinterpret_cast<const uint2*>(zh + row * c);
  const uint2* lr = zl ? reinterpret_cast<const uint2*>(zl + row * c) : nullptr;
  float dot = 0.f;
  for (int i = lane; i < cq; i += 32) {
    const float4 g = __ldg(gr + i);
    const float4 zv = planes_to_float4(hr[i], lr ? lr + i : nullptr);
    dot += g.x * zv.x + g.y * zv.y + g.z * zv.z + g.w * zv.w;
  }
  dot = gx_warp_sum(dot);
  const float inv = inv_norm[row];
  uint2* oh = dh ? reinterpret_cast<uint2*>(dh + row * c) : nullptr;
  uint2* ol = dl ? reinterpret_cast<uint2*>(dl + row * c) : nullptr;
  float4* of = dz_f32 ? reinterpret_cast<float4*>(dz_f32 + row * c) : nullptr;
  for (int i = lane; i < cq; i += 32) {
    const float4 g = __ldg(gr + i);
    const float4 zv = planes_to_float4(hr[i], lr ? lr + i : nullptr);
    const float4 d = make_float4((g.x - zv.x * dot) * inv, (g.y - zv.y * dot) * inv, (g.z - zv.z * dot) * inv,
                                 (g.w - zv.w * dot) * inv);
    if (oh) {
      uint2 h, l;
      gx_split4(d, h, l);
      oh[i] = h;
      if (ol) ol[i] = l;
    }
    if (of) of[i] = d;
  }
}

// out[seg,:] = sum over r in [seg_off[seg], seg_off[seg+1]) of rows[order[r], :]  (deterministic
// gather-reduce; used to fold the per-patch dZ rows of pixels sampled several times into one
// row per pixel), written as bf16 planes.  Warp per segment.
// (BF16_ROWS: the rows are single bf16 planes - the bf16-backward mode stores dZ rows at half the bytes)
template <bool BF16_ROWS>
__global__ void segment_sum_rows_kernel(const void* __restrict__ rows_v, const int* __restrict__ order,
                                        const int* __restrict__ seg_off, __nv_bfloat16* __restrict__ hi,
                                        __nv_bfloat16* __restrict__ lo, float* __restrict__ out_f32, long long nseg,
                                        int c) {
  const float* rows = reinterpret_cast<const float*>(rows_v);
  const __nv_bfloat16* rows_h = reinterpret_cast<const __nv_bfloat16*>(rows_v);
  const int lane = threadIdx.x & 31;
  const long long seg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (seg >= nseg) return;
  const int r0 = seg_off[seg], r1 = seg_off[seg + 1];
  const int cq = c >> 2;
  uint2* oh = hi ? reinterpret_cast<uint2*>(hi + seg * c) : nullptr;
  uint2* ol = lo ? reinterpret_cast<uint2*>(lo + seg * c) : nullptr;
  float4* of = out_f32 ? reinterpret_cast<float4*>(out_f32 + seg * c) : nullptr;
  // 512 channels per sweep: the four 128-bit loads of a row are independent and in flight together
  for (int i0 = 0; i0 < cq; i0 += 128) {
    float4 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = r0; r < r1; ++r) {
      const float4* src = reinterpret_cast<const float4*>(rows + (long long)order[r] * c) + i0;
      const uint2* src_h = reinterpret_cast<const uint2*>(rows_h + (long long)order[r] * c) + i0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = lane + 32 * j;
        if (i0 + i < cq) {
          float4 v;
          if constexpr (BF16_ROWS) v = planes_to_float4(__ldg(src_h + i), nullptr);
          else v = gx_ldg_stream(src + i);
          acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = i0 + lane + 32 * j;
      if (i < cq) {
        if (oh) {
          uint2 h, l;
          gx_split4(acc[j], h, l);
          oh[i] = h;
          if (ol) ol[i] = l;
        }
        if (of) of[i] = acc[j];
      }
    }
  }
}

// Z[b,y,x,:] = sum_l P_l[b, y*h_l/H, x*w_l/W, :]  - the projection of the nearest-upsampled,
// channel-concatenated feature vector is the sum of the per-level projections evaluated at
// each level's native resolution (linearity), so the 11x redundant upsampled GEMM is never run.
struct UpsumDesc {
  int nlevels;
  const float* p[GX_MAX_LEVELS];
  int h[GX_MAX_LEVELS], w[GX_MAX_LEVELS];
  int out_h, out_w, c;
  int bilinear;     // 0: nearest (floor(dst * in / out)), 1: F.interpolate(mode='bilinear', align_corners=False)
  long long npix;   // B*out_h*out_w
};

// source coordinate of torch's upsample_bilinear2d with align_corners = False:
// src = max(0, scale * (dst + 0.5) - 0.5), i0 = floor(src), i1 = i0 + (i0 < n - 1), lambda = src - i0
__device__ __forceinline__ void bilinear_src(int dst, int n_in, int n_out, int& i0, int& i1, float& lam) {
  const float scale = (float)n_in / (float)n_out;
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > n_in - 1) i0 = n_in - 1;
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
  lam = src - (float)i0;
}
// first arg-max over the channels of one pixel from per-lane candidates (value, channel): larger value wins, the
// smaller channel on ties (ref: out_preds.max(1)[1], swav_clustering.py:691)
__device__ __forceinline__ void upsum_argmax_take(float& best, int& bi, const float4 a, int ch0) {
  if (a.x > best) { best = a.x; bi = ch0; }
  if (a.y > best) { best = a.y; bi = ch0 + 1; }
  if (a.z > best) { best = a.z; bi = ch0 + 2; }
  if (a.w > best) { best = a.w; bi = ch0 + 3; }
}
__device__ __forceinline__ int upsum_argmax_warp(float best, int bi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  return bi == 0x7fffffff ? 0 : bi;
}

__global__ void upsample_sum_kernel(const UpsumDesc d, float* __restrict__ out, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, long long* __restrict__ labels) {
  const int lane = threadIdx.x & 31;
  const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= d.npix) return;
  float am_best = -INFINITY;
  int am_i = 0x7fffffff;
  const int per = d.out_h * d.out_w;
  const int b = (int)(pix / per);
  const int r = (int)(pix - (long long)b * per);
  const int y = r / d.out_w, x = r - y * d.out_w;
  const int cq = d.c >> 2;
  for (int c0 = 0; c0 < cq; c0 += 128) {        // 512 channels per sweep, 4 float4 per lane
    float4 acc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int l = 0; l < d.nlevels; ++l) {
      if (d.bilinear && (d.h[l] != d.out_h || d.w[l] != d.out_w)) {
        int y0, y1, x0, x1;
        float ly, lx;
        bilinear_src(y, d.h[l], d.out_h, y0, y1, ly);
        bilinear_src(x, d.w[l], d.out_w, x0, x1, lx);
        const float* base = d.p[l] + (long long)b * d.h[l] * d.w[l] * d.c;
        const float4* p00 = reinterpret_cast<const float4*>(base + ((long long)y0 * d.w[l] + x0) * d.c) + c0;
        const float4* p01 = reinterpret_cast<const float4*>(base + ((long long)y0 * d.w[l] + x1) * d.c) + c0;
        const float4* p10 = reinterpret_cast<const float4*>(base + ((long long)y1 * d.w[l] + x0) * d.c) + c0;
        const float4* p11 = reinterpret_cast<const float4*>(base + ((long long)y1 * d.w[l] + x1) * d.c) + c0;
        const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int i = lane + 32 * j;
          if (c0 + i < cq) {
            const float4 a = __ldg(p00 + i), bq = __ldg(p01 + i), cc = __ldg(p10 + i), dd = __ldg(p11 + i);
            acc[j].x += w00 * a.x + w01 * bq.x + w10 * cc.x + w11 * dd.x;
            acc[j].y += w00 * a.y + w01 * bq.y + w10 * cc.y + w11 * dd.y;
            acc[j].z += w00 * a.z + w01 * bq.z + w10 * cc.z + w11 * dd.z;
            acc[j].w += w00 * a.w + w01 * bq.w + w10 * cc.w + w11 * dd.w;
          }
        }
        continue;
      }
      const int ly = (y * d.h[l]) / d.out_h, lx = (x * d.w[l]) / d.out_w;
      const float4* src =
          reinterpret_cast<const float4*>(d.p[l] + (((long long)b * d.h[l] + ly) * d.w[l] + lx) * d.c) + c0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = lane + 32 * j;
        if (c0 + i < cq) {
          const float4 v = __ldg(src + i);
          acc[j].x += v.x; acc[j].y += v.y; acc[j].z += v.z; acc[j].w += v.w;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = lane + 32 * j;
      if (c0 + i < cq) {
        if (out) gx_stg_stream(reinterpret_cast<float4*>(out + pix * d.c) + c0 + i, acc[j]);
        if (labels) upsum_argmax_take(am_best, am_i, acc[j], 4 * (c0 + i));
        if (hi) {      // operand planes of a consumer conv / GEMM, emitted in the same pass
          uint2 h, l;
          gx_split4(acc[j], h, l);
          reinterpret_cast<uint2*>(hi + pix * d.c)[c0 + i] = h;
          if (lo) reinterpret_cast<uint2*>(lo + pix * d.c)[c0 + i] = l;
        }
      }
    }
  }
  if (labels) {
    const int a = upsum_argmax_warp(am_best, am_i);
    if (lane == 0) labels[pix] = a;
  }
}

// Nearest upsampling on a pyramid: a level whose upsampling factor is an even integer in both directions gives the
// same source vector to the four pixels of an aligned 2x2 output quad.  One warp per quad: the `n_shared` coarse
// levels (they come first) are loaded and summed once, in level order, the remaining (at most two) levels are loaded
// per pixel and added on top - the same additions in the same order as the per-pixel kernel above, so the result is
// bit-identical, with 2.3x fewer 128-bit loads through L1 (the per-pixel kernel is bound by L1 wavefronts, not by
// HBM: 7 loads + 1 store of 2 KB per pixel).  All loads of a 128-channel sweep are in flight together.
constexpr int UPQ_SHARED = 8;
constexpr int UPQ_FINE = 2;
// LABELS: the arg-max bookkeeping is compiled in only for the label-map path (as a run-time test it cost the
// training step's instantiation 20 %: 0.97 -> 1.19 ms per 16 images)
template <int UPQ_S, int UPQ_F, bool LABELS>
__global__ void __launch_bounds__(256, 2)
upsample_sum_quad_kernel(const UpsumDesc d, int n_shared, float* __restrict__ out, __nv_bfloat16* __restrict__ hi,
                         __nv_bfloat16* __restrict__ lo, long long* __restrict__ labels) {
  const int lane = threadIdx.x & 31;
  const int qh = d.out_h >> 1, qw = d.out_w >> 1;
  const long long quad = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (quad >= (d.npix >> 2)) return;
  const int per = qh * qw;
  const int b = (int)(quad / per);
  const int r = (int)(quad - (long long)b * per);
  const int y = (r / qw) * 2, x = (r - (r / qw) * qw) * 2;
  const int cq = d.c >> 2;
  const int n_fine = d.nlevels - n_shared;
  const float4* src[UPQ_S];
#pragma unroll
  for (int l = 0; l < UPQ_S; ++l) {
    src[l] = nullptr;
    if (l < n_shared) {
      const int ly = (y * d.h[l]) / d.out_h, lx = (x * d.w[l]) / d.out_w;
      src[l] = reinterpret_cast<const float4*>(d.p[l] + (((long long)b * d.h[l] + ly) * d.w[l] + lx) * d.c);
    }
  }
  const float4* fsrc[UPQ_F];
  int foff[UPQ_F][4];        // float4 offsets of the four pixels' sources from fsrc
#pragma unroll
  for (int m = 0; m < UPQ_F; ++m) {
    fsrc[m] = nullptr;
    if (m < n_fine) {
      const int l = n_shared + m;
      fsrc[m] = reinterpret_cast<const float4*>(d.p[l] + (long long)b * d.h[l] * d.w[l] * d.c);
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int ly = ((y + (p >> 1)) * d.h[l]) / d.out_h, lx = ((x + (p & 1)) * d.w[l]) / d.out_w;
        foff[m][p] = (ly * d.w[l] + lx) * cq;
      }
    }
  }
  const long long pix0 = ((long long)b * d.out_h + y) * d.out_w + x;
  float am_best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
  int am_i[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
  for (int i = lane; i < cq; i += 32) {
    float4 v[UPQ_S], f[UPQ_F][4];
#pragma unroll
    for (int l = 0; l < UPQ_S; ++l)
      if (l < n_shared) v[l] = __ldg(src[l] + i);
#pragma unroll
    for (int m = 0; m < UPQ_F; ++m)
#pragma unroll
      for (int p = 0; p < 4; ++p)
        if (m < n_fine) f[m][p] = gx_ldg_stream(fsrc[m] + foff[m][p] + i);
    float4 acc_s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int l = 0; l < UPQ_S; ++l)
      if (l < n_shared) { acc_s.x += v[l].x; acc_s.y += v[l].y; acc_s.z += v[l].z; acc_s.w += v[l].w; }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      float4 acc = acc_s;
#pragma unroll
      for (int m = 0; m < UPQ_F; ++m)
        if (m < n_fine) { acc.x += f[m][p].x; acc.y += f[m][p].y; acc.z += f[m][p].z; acc.w += f[m][p].w; }
      const long long pix = pix0 + (p >> 1) * d.out_w + (p & 1);
      if (out) gx_stg_stream(reinterpret_cast<float4*>(out + pix * d.c) + i, acc);
      if constexpr (LABELS) upsum_argmax_take(am_best[p], am_i[p], acc, 4 * i);
      if (hi) {
        uint2 h2, l2;
        gx_split4(acc, h2, l2);
        reinterpret_cast<uint2*>(hi + pix * d.c)[i] = h2;
        if (lo) reinterpret_cast<uint2*>(lo + pix * d.c)[i] = l2;
      }
    }
  }
  if constexpr (LABELS) {      // the label map of predict_swav_codes from the sums in registers: Z is not read again
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int a = upsum_argmax_warp(am_best[p], am_i[p]);
      if (lane == 0) labels[pix0 + (p >> 1) * d.out_w + (p & 1)] = a;
    }
  }
}

// Adjoint of 1-D bilinear upsampling along one axis: in [outer, n_in(fine), inner] -> out [outer, n_out(coarse),
// inner], out[o, J, :] = sum_j w(j -> J) in[o, j, :] with the forward weights of `bilinear_src`.  Two calls
// (x then y) give the adjoint of the separable 2-D bilinear upsampling - the weight-gradient fold of dZ onto a
// coarser level when hf_interp = 'bilinear'.  One warp per (o, J, 512-float slice of inner).
__global__ void pool1d_bilinear_kernel(const float* __restrict__ in, long long outer, int n_in, int n_out,
                                       long long inner, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const long long iq = inner >> 2;                      // float4 per (o, j)
  const long long slices = (iq + 127) / 128;
  const long long wid = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= outer * n_out * slices) return;
  const long long sl = wid % slices;
  const int J = (int)((wid / slices) % n_out);
  const long long o = wid / (slices * n_out);
  // fine indices whose source coordinate falls in (J - 1, J + 1): a superset is enough, weights decide
  const float inv = (float)n_in / (float)n_out;         // fine per coarse
  int jlo = (int)floorf(((float)J - 0.5f) * inv - 0.5f) - 1;
  int jhi = (int)ceilf(((float)J + 1.5f) * inv - 0.5f) + 1;
  if (jlo < 0) jlo = 0;
  if (jhi > n_in - 1) jhi = n_in - 1;
  float4 acc[4];
#pragma unroll
  for (int t = 0; t < 4; ++t) acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j = jlo; j <= jhi; ++j) {
    int i0, i1;
    float lam;
    bilinear_src(j, n_out, n_in, i0, i1, lam);
    const float wj = (i0 == J ? 1.f - lam : 0.f) + (i1 == J ? lam : 0.f);
    if (wj == 0.f) continue;                            // warp-uniform
    const float4* src = reinterpret_cast<const float4*>(in + (o * n_in + j) * inner) + sl * 128;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const long long i = sl * 128 + lane + 32 * t;
      if (i < iq) {
        const float4 v = __ldg(src + lane + 32 * t);
        acc[t].x = fmaf(wj, v.x, acc[t].x); acc[t].y = fmaf(wj, v.y, acc[t].y);
        acc[t].z = fmaf(wj, v.z, acc[t].z); acc[t].w = fmaf(wj, v.w, acc[t].w);
      }
    }
  }
  float4* dst = reinterpret_cast<float4*>(out + (o * n_out + J) * inner) + sl * 128;
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const long long i = sl * 128 + lane + 32 * t;
    if (i < iq) dst[lane + 32 * t] = acc[t];
  }
}

// 3x3 (dilated) conv with few output channels as ONE GEMM + a stencil sum: G[pix, tap*cout + co] =
// x[pix,:] . W[co,:,tap] for all nine taps at once (the input is read once instead of nine times), then
// out[y,x,co] = act(bias[co] + sum_tap G[(y,x) + offset(tap), tap*cout + co]) with zeros outside the image.
__global__ void tap_sum_kernel(const float* __restrict__ g, int batch, int h, int w, int cout, int dil,
                               const float* __restrict__ bias, int act, float* __restrict__ out,
                               __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo, int next_ld) {
  const int cq = cout >> 2;
  const long long total = (long long)batch * h * w * cq;
  const long long ldg = 9LL * cout;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cq);
    const long long pix = i / cq;
    const int x = (int)(pix % w);
    const int y = (int)((pix / w) % h);
    float4 acc = bias ? __ldg(reinterpret_cast<const float4*>(bias) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + (ky - 1) * dil;
      if (yy < 0 || yy >= h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + (kx - 1) * dil;
        if (xx < 0 || xx >= w) continue;
        const long long np = pix + (long long)(yy - y) * w + (xx - x);
        const float4 v = __ldg(reinterpret_cast<const float4*>(g + np * ldg + (ky * 3 + kx) * cout) + q);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (act == 2) {
      acc.x = acc.x > 0.f ? acc.x : acc.x * 0.2f; acc.y = acc.y > 0.f ? acc.y : acc.y * 0.2f;
      acc.z = acc.z > 0.f ? acc.z : acc.z * 0.2f; acc.w = acc.w > 0.f ? acc.w : acc.w * 0.2f;
    }
    if (out) reinterpret_cast<float4*>(out + pix * cout)[q] = acc;
    if (hi) {
      uint2 hh, ll;
      gx_split4(acc, hh, ll);
      reinterpret_cast<uint2*>(hi + pix * next_ld)[q] = hh;
      if (lo) reinterpret_cast<uint2*>(lo + pix * next_ld)[q] = ll;
    }
  }
}

// Adjoint of tap_sum: dG[q, tap*cout + co] = dOut[q - offset(tap), co] (zero outside the image), emitted
// as bf16 split planes - the operand of the weight-gradient GEMM dW_all = dG^T X and of dX = dG W_all.
__global__ void tap_spread_kernel(const float* __restrict__ dout, int batch, int h, int w, int cout, int dil,
                                  __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  const int cq = cout >> 2;
  const long long total = (long long)batch * h * w * 9 * cq;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int q = (int)(i % cq);
    const int tap = (int)((i / cq) % 9);
    const long long pix = i / (9 * cq);
    const int x = (int)(pix % w);
    const int y = (int)((pix / w) % h);
    const int ky = tap / 3, kx = tap - 3 * ky;
    const int yy = y - (ky - 1) * dil, xx = x - (kx - 1) * dil;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
      const long long sp = pix + (long long)(yy - y) * w + (xx - x);
      v = __ldg(reinterpret_cast<const float4*>(dout + sp * cout) + q);
    }
    uint2 hh, ll;
    gx_split4(v, hh, ll);
    const long long o = (pix * 9 + tap) * cq + q;
    reinterpret_cast<uint2*>(hi)[o] = hh;
    if (lo) reinterpret_cast<uint2*>(lo)[o] = ll;
  }
}

// out[b,y,x,:] = sum of the f x f block of in (f = H/h): the adjoint of nearest upsampling; also
// emitted as bf16 planes (operand of the per-level weight-gradient GEMM).  Warp per output pixel.
__global__ void pool_sum_kernel(const float* __restrict__ in, int H, int W, int h, int w, int c, long long nout,
                                float* __restrict__ out, __nv_bfloat16* __restrict__ hi,
                                __nv_bfloat16* __restrict__ lo) {
  const int lane = threadIdx.x & 31;
  const long long o = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (o >= nout) return;
  const long long per = (long long)h * w;
  const int b = (int)(o / per);
  const int r = (int)(o - (long long)b * per);
  const int oy = r / w, ox = r - oy * w;
  const int fy = H / h, fx = W / w;
  const int cq = c >> 2;
  if (fy == 2 && fx == 2) {       // the 2x2 steps of the resolution chain: all loads of a sweep in flight at once
    const float4* p00 = reinterpret_cast<const float4*>(in + (((long long)b * H + 2 * oy) * W + 2 * ox) * c);
    const float4* p10 = p00 + (long long)W * cq;
    for (int i0 = 0; i0 < cq; i0 += 128) {
      float4 v[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + lane + 32 * j;
        if (i < cq) {
          v[j][0] = gx_ldg_stream(p00 + i); v[j][1] = gx_ldg_stream(p00 + cq + i);
          v[j][2] = gx_ldg_stream(p10 + i); v[j][3] = gx_ldg_stream(p10 + cq + i);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int i = i0 + lane + 32 * j;
        if (i < cq) {
          float4 acc;
          acc.x = (v[j][0].x + v[j][1].x) + (v[j][2].x + v[j][3].x);
          acc.y = (v[j][0].y + v[j][1].y) + (v[j][2].y + v[j][3].y);
          acc.z = (v[j][0].z + v[j][1].z) + (v[j][2].z + v[j][3].z);
          acc.w = (v[j][0].w + v[j][1].w) + (v[j][2].w + v[j][3].w);
          if (out) reinterpret_cast<float4*>(out + o * c)[i] = acc;
          if (hi) {
            uint2 hh, ll;
            gx_split4(acc, hh, ll);
            reinterpret_cast<uint2*>(hi + o * c)[i] = hh;
            if (lo) reinterpret_cast<uint2*>(lo + o * c)[i] = ll;
          }
        }
      }
    }
    return;
  }
  for (int i = lane; i < cq; i += 32) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int dy = 0; dy < fy; ++dy)
      for (int dx = 0; dx < fx; ++dx) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(
                                   in + (((long long)b * H + oy * fy + dy) * W + ox * fx + dx) * c) + i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    if (out) reinterpret_cast<float4*>(out + o * c)[i] = acc;
    if (hi) {
      uint2 hh, ll;
      gx_split4(acc, hh, ll);
      reinterpret_cast<uint2*>(hi + o * c)[i] = hh;
      if (lo) reinterpret_cast<uint2*>(lo + o * c)[i] = ll;
    }
  }
}

__global__ void normalize_rows_kernel(float* __restrict__ w, long long rows, int cols) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* wr = w + row * cols;
  float ss = 0.f;
  for (int i = lane; i < cols; i += 32) ss = fmaf(wr[i], wr[i], ss);
  ss = gx_warp_sum(ss);
  const float inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);
  for (int i = lane; i < cols; i += 32) wr[i] *= inv;
}

// ---------------------------------------------------------------------------
// fp32 -> split planes (optionally transposed through a padded smem tile)
// ---------------------------------------------------------------------------
__global__ void split_planes_kernel(const float* __restrict__ x, long long ld, __nv_bfloat16* __restrict__ hi,
                                    __nv_bfloat16* __restrict__ lo, long long rows, long long cols, long long ldo) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / cols, c = i - r * cols;
    __nv_bfloat16 h, l;
    gx_split_bf16(x[r * ld + c], h, l);
    hi[r * ldo + c] = h;
    if (lo) lo[r * ldo + c] = l;
  }
}

// 4 columns per thread (all pitches and base addresses multiples of 4 elements)
__global__ void split_planes_v4_kernel(const float* __restrict__ x, long long ld, __nv_bfloat16* __restrict__ hi,
                                       __nv_bfloat16* __restrict__ lo, long long rows, int cq, long long ldo) {
  const long long total = rows * cq;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // four independent 128-bit loads in flight per thread
  for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    float4 v[4];
    long long r[4];
    int q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = i0 + u * stride;
      r[u] = i / cq;
      q[u] = (int)(i - r[u] * cq);
      if (i < total) v[u] = gx_ldg_stream(reinterpret_cast<const float4*>(x + r[u] * ld) + q[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (i0 + u * stride >= total) break;
      uint2 h, l;
      gx_split4(v[u], h, l);
      reinterpret_cast<uint2*>(hi + r[u] * ldo)[q[u]] = h;
      if (lo) reinterpret_cast<uint2*>(lo + r[u] * ldo)[q[u]] = l;
    }
  }
}

__global__ void split_planes_t_kernel(const float* __restrict__ x, long long ld, __nv_bfloat16* __restrict__ hi,
                                      __nv_bfloat16* __restrict__ lo, long long rows, long long cols) {
  __shared__ float tile[32][33];
  const long long r0 = (long long)blockIdx.y * 32, c0 = (long long)blockIdx.x * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? x[r * ld + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long c = c0 + j, r = r0 + threadIdx.x;  // output [cols, rows]
    if (c < cols && r < rows) {
      __nv_bfloat16 h, l;
      gx_split_bf16(tile[threadIdx.x][j], h, l);
      hi[c * rows + r] = h;
      if (lo) lo[c * rows + r] = l;
    }
  }
}

// ---------------------------------------------------------------------------
// Sinkhorn-Knopp pass (scaling-vector form).  512 threads own the K columns
// (4 consecutive columns per thread per 2048-column sweep, J sweeps), stream rows of S
// once, keep the column accumulators u'_k in registers.
// ---------------------------------------------------------------------------
constexpr int SK_THREADS = 512;
constexpr int SK_ROWS = 2;

// ---------------------------------------------------------------------------
// Streaming kernels over the rows of S.  Rows are brought into a shared-memory ring by the
// TMA engine (cp.async.bulk, one contiguous K*4-byte copy per row) several iterations
// ahead of the compute.  The CTA is split into NG independent groups of GT threads (named
// barriers), each group owning all K columns (4 consecutive columns per thread per
// GT*4-column sweep, J sweeps) and consuming every NG-th ring stage, so the exp / shuffle /
// barrier phases of one group overlap those of the other.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 ex2_2(float2 a) { return make_float2(ex2_fast(a.x), ex2_fast(a.y)); }

__device__ __forceinline__ void group_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// MODE 0: rows of S (fp32) in.  MODE 1: the same pass, which also stores its row-normalised terms
//   e16[n,k] = half(2^15 * a_k e_nk / sum_k a_k e_nk)   and   la1[k] = log2 a_k
// - the 16-bit cache of the scaled kernel matrix that sinkhorn_pass16_kernel streams in the later passes.
template <int J, int R, int GT, int MODE>
__global__ void __launch_bounds__(SK_THREADS, 1)
sinkhorn_pass_kernel(const float* __restrict__ s, long long n, int k, long long lds, float scale_log2, int first,
                     const float* __restrict__ u_in, const gx_ll_desc u_ll, const float* __restrict__ r,
                     const float* __restrict__ cvec, float c_uniform, float* __restrict__ partials, int stages,
                     int reverse, __half* __restrict__ e16, long long lde, float* __restrict__ la1) {
  constexpr int NG = SK_THREADS / GT;   // groups per CTA
  constexpr int NW = GT / 32;           // warps per group
  extern __shared__ __align__(128) uint8_t sk_smem[];
  __shared__ __align__(16) float red[NG][2][R][NW];
  __shared__ uint64_t full_bar[8];
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid / GT, gt = tid % GT, gwarp = gt >> 5;
  const uint32_t row_bytes = (uint32_t)k * 4u;
  const uint32_t stage_bytes = row_bytes * R;
  if (tid == 0) {
    for (int i = 0; i < stages; ++i) gxptx::mbar_init(&full_bar[i], 1);
    gxptx::fence_mbar_init();
  }
  __syncthreads();
  // iteration `it` of this CTA covers the row group g = blockIdx.x + it*gridDim.x (rows [g*R, +R)), or - in a
  // reverse pass - the mirrored group, so that consecutive passes sweep S in opposite directions
  const long long groups_total = (n + R - 1) / R;
  const int n_iters = (int)((groups_total - blockIdx.x + gridDim.x - 1) / gridDim.x);
  auto first_row = [&](int it) -> long long {
    const long long g = (long long)blockIdx.x + (long long)it * gridDim.x;
    return (reverse ? groups_total - 1 - g : g) * R;
  };
  auto issue = [&](int it, int st) {   // st == it % stages
    const long long row0 = first_row(it);
    const int valid = (int)min((long long)R, n - row0);
    gxptx::mbar_arrive_expect_tx(&full_bar[st], row_bytes * valid);
    for (int rr = 0; rr < valid; ++rr)
      gxptx::bulk_load_1d(sk_smem + (size_t)st * stage_bytes + (size_t)rr * row_bytes, s + (row0 + rr) * lds,
                          row_bytes, &full_bar[st]);
  };
  if (tid == 0)
    for (int it = 0; it < stages && it < n_iters; ++it) issue(it, it);

  // per-thread columns; out-of-range columns read a clamped (valid) address and are
  // neutralised by log2(a) = -inf  ->  exp2(-inf) = 0
  // (packed fp32x2 arithmetic: the pass is close enough to the issue limit that its HBM rate follows the
  // SM clock under the power cap; FFMA2/FADD2 cut the FMA-pipe instructions per score from 3 to 1.5)
  float2 la2[J][2], acc[J][2];
  int coff[J];
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int col = j * (GT * 4) + gt * 4;
    coff[j] = min(col, k - 4);
    float l4[4], u4[4] = {1.f, 1.f, 1.f, 1.f};
    if (!first && col < k) {
      // marginals of the previous pass: local vector, or the tagged slots the peers pushed into this rank's
      // exchange buffer (they arrived while the pass in between was running; the rows above are already in flight)
      if (u_ll.world > 0) gxll::recv4(u_ll, k, col, u4);
      else {
        const float4 uu = *reinterpret_cast<const float4*>(u_in + col);
        u4[0] = uu.x; u4[1] = uu.y; u4[2] = uu.z; u4[3] = uu.w;
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (col < k) {
        l4[e] = 0.f;
        if (!first) {
          const float rk = r ? r[col + e] : 1.f / (float)k;
          l4[e] = log2f(rk / u4[e]);
        }
      } else {
        l4[e] = -INFINITY;
      }
    }
    if (MODE == 1 && col < k && blockIdx.x == 0 && grp == 0)
      *reinterpret_cast<float4*>(la1 + col) = make_float4(l4[0], l4[1], l4[2], l4[3]);
    la2[j][0] = make_float2(l4[0], l4[1]);
    la2[j][1] = make_float2(l4[2], l4[3]);
    acc[j][0] = acc[j][1] = make_float2(0.f, 0.f);
  }
  const float2 sc2 = make_float2(scale_log2, scale_log2);
  int buf = 0, st = grp;   // st == it % stages, phase == (it / stages) & 1, kept incrementally (no integer division)
  uint32_t phase = 0;
  for (int it = grp; it < n_iters; it += NG) {
    const long long row0 = first_row(it);
    gxptx::mbar_wait(&full_bar[st], phase);
    const float* srow = reinterpret_cast<const float*>(sk_smem + (size_t)st * stage_bytes);
    float2 p[R][J][2];
    float t[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      t[rr] = 0.f;
      if (row0 + rr < n) {   // warp-uniform
        float2 t2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < J; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(srow + (size_t)rr * k + coff[j]);
          p[rr][j][0] = ex2_2(fma2(make_float2(v.x, v.y), sc2, la2[j][0]));
          p[rr][j][1] = ex2_2(fma2(make_float2(v.z, v.w), sc2, la2[j][1]));
          t2 = add2(t2, add2(p[rr][j][0], p[rr][j][1]));
        }
        t[rr] = t2.x + t2.y;
      } else {
#pragma unroll
        for (int j = 0; j < J; ++j) p[rr][j][0] = p[rr][j][1] = make_float2(0.f, 0.f);
      }
    }
    if (!first) {
#pragma unroll
      for (int rr = 0; rr < R; ++rr) {
        const float w = gx_warp_sum(t[rr]);
        if (lane == 0) red[grp][buf][rr][gwarp] = w;
      }
    }
    group_bar(1 + grp, GT);   // the group has consumed the stage (and published its partial sums)
    if (gt == 0 && it + stages < n_iters) issue(it + stages, st);
    float bn[R], e16n[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      bn[rr] = 1.f;
      e16n[rr] = 0.f;
      if (!first) {
        float tot = 0.f;
#pragma unroll
        for (int w = 0; w < NW; w += 4) {
          const float4 q = *reinterpret_cast<const float4*>(&red[grp][buf][rr][w]);
          tot += (q.x + q.y) + (q.z + q.w);
        }
        const long long row = row0 + rr;
        const float cn = cvec ? ((row < n) ? cvec[row] : 0.f) : c_uniform;
        bn[rr] = __fdividef(cn, tot);
        if (MODE == 1) e16n[rr] = __fdividef(32768.f, tot);
      }
    }
    buf ^= 1;
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      if (row0 + rr < n) {
        const float2 b2 = make_float2(bn[rr], bn[rr]);
#pragma unroll
        for (int j = 0; j < J; ++j) {
          acc[j][0] = fma2(p[rr][j][0], b2, acc[j][0]);
          acc[j][1] = fma2(p[rr][j][1], b2, acc[j][1]);
        }
        if (MODE == 1) {
          const float2 n2 = make_float2(e16n[rr], e16n[rr]);
          __half* erow = e16 + (row0 + rr) * lde;
#pragma unroll
          for (int j = 0; j < J; ++j) {
            const int col = j * (GT * 4) + gt * 4;
            if (col < k) {
              const float2 a = mul2(p[rr][j][0], n2), b = mul2(p[rr][j][1], n2);
              const __half2 h0 = __floats2half2_rn(a.x, a.y), h1 = __floats2half2_rn(b.x, b.y);
              *reinterpret_cast<uint2*>(erow + col) =
                  make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            }
          }
        }
      }
    }
    st += NG;
    if (st >= stages) { st -= stages; phase ^= 1u; }
  }
  // acc holds sum_n a_k e_nk b_n; the marginal of the unscaled matrix is acc / a_k
  float* prow = partials + ((long long)blockIdx.x * NG + grp) * k;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int col = j * (GT * 4) + gt * 4;
    if (col < k)
      *reinterpret_cast<float4*>(prow + col) =
          make_float4(acc[j][0].x * ex2_fast(-la2[j][0].x), acc[j][0].y * ex2_fast(-la2[j][0].y),
                      acc[j][1].x * ex2_fast(-la2[j][1].x), acc[j][1].y * ex2_fast(-la2[j][1].y));
  }
}

// ---------------------------------------------------------------------------
// A later Sinkhorn pass on the 16-bit cache e16 (written by sinkhorn_pass_kernel<.., 1>) instead of S: half the bytes
// and no exponentials.  With rho_k = a_k / a1_k (a1 = the column scaling the cache was written with) the row totals
// are t'_n = sum_k e16_nk rho_k and the marginals u_k = (1 / a1_k) sum_n e16_nk c_n / t'_n; the row factor
// 2^15 / t_n stored in e16 cancels.  Same decomposition as the fp32 pass (NG groups of GT threads, each thread owning
// 4 consecutive columns per sweep), but R rows per ring stage loaded by ONE bulk copy (the plane is contiguous) and
// incremental ring bookkeeping: the pass is bound by instruction issue and latency at 16 warps per SM, not by HBM.
// ---------------------------------------------------------------------------
template <int J, int R, int GT>
__global__ void __launch_bounds__(SK_THREADS, 1)
sinkhorn_pass16_kernel(const __half* __restrict__ e16, long long n, int k, long long lde,
                       const float* __restrict__ u_in, const gx_ll_desc u_ll, const float* __restrict__ r,
                       const float* __restrict__ cvec, float c_uniform, const float* __restrict__ la1,
                       float* __restrict__ partials, int stages, int reverse) {
  constexpr int NG = SK_THREADS / GT;   // groups per CTA
  constexpr int NW = GT / 32;           // warps per group
  extern __shared__ __align__(128) uint8_t sk_smem[];
  __shared__ __align__(16) float red[NG][2][NW][R];
  __shared__ uint64_t full_bar[8];
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid / GT, gt = tid % GT, gwarp = gt >> 5;
  const uint32_t row_bytes = (uint32_t)lde * 2u;
  const uint32_t stage_bytes = row_bytes * R;
  if (tid == 0) {
    for (int i = 0; i < stages; ++i) gxptx::mbar_init(&full_bar[i], 1);
    gxptx::fence_mbar_init();
  }
  __syncthreads();
  const int groups_total = (int)((n + R - 1) / R);
  const int n_iters = (groups_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int last_valid = (int)(n - (long long)(groups_total - 1) * R);   // rows of the last (possibly ragged) group
  auto group_of = [&](int it) -> int {
    const int g = (int)blockIdx.x + it * (int)gridDim.x;
    return reverse ? groups_total - 1 - g : g;
  };
  auto issue = [&](int it, int st) {   // st == it % stages
    const int g = group_of(it);
    const uint32_t bytes = row_bytes * (uint32_t)(g == groups_total - 1 ? last_valid : R);
    gxptx::mbar_arrive_expect_tx(&full_bar[st], bytes);
    gxptx::bulk_load_1d(sk_smem + (size_t)st * stage_bytes, e16 + (long long)g * R * lde, bytes, &full_bar[st]);
  };
  if (tid == 0)
    for (int it = 0; it < stages && it < n_iters; ++it) issue(it, it);

  float2 rho[J][2], acc[J][2];
  // byte offset of this thread's 4 halves in a row: gt * 8 + j * GT * 8; only the last sweep can run past column k
  // (J = ceil(k / (4 GT))): its slot is clamped to re-read valid columns, with rho = 0
  const uint32_t coff0 = (uint32_t)gt * 8u;
  const uint32_t coff_last = (uint32_t)min((J - 1) * (GT * 4) + gt * 4, k - 4) * 2u;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int col = j * (GT * 4) + gt * 4;
    float l4[4] = {0.f, 0.f, 0.f, 0.f};
    if (col < k) {
      float u4[4];
      if (u_ll.world > 0) gxll::recv4(u_ll, k, col, u4);
      else {
        const float4 uu = *reinterpret_cast<const float4*>(u_in + col);
        u4[0] = uu.x; u4[1] = uu.y; u4[2] = uu.z; u4[3] = uu.w;
      }
      const float4 l1 = *reinterpret_cast<const float4*>(la1 + col);
      const float l1v[4] = {l1.x, l1.y, l1.z, l1.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float rk = r ? r[col + e] : 1.f / (float)k;
        l4[e] = ex2_fast(log2f(rk / u4[e]) - l1v[e]);
      }
    }
    rho[j][0] = make_float2(l4[0], l4[1]);
    rho[j][1] = make_float2(l4[2], l4[3]);
    acc[j][0] = acc[j][1] = make_float2(0.f, 0.f);
  }
  int buf = 0, sg = grp;
  uint32_t phase = 0;
  for (int it = grp; it < n_iters; it += NG) {
    const int g = group_of(it);
    const int valid = g == groups_total - 1 ? last_valid : R;
    gxptx::mbar_wait(&full_bar[sg], phase);
    const uint8_t* srow = sk_smem + (size_t)sg * stage_bytes;
    float2 p[R][J][2];
    float t[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      // unconditional: the rows of a ragged last group beyond `valid` hold whatever the stage held before - their
      // totals are never used and their terms are skipped below (no branch here, so the rows' loads and conversions
      // interleave)
      float2 ta = make_float2(0.f, 0.f), tb = make_float2(0.f, 0.f);
      const uint8_t* sr = srow + (size_t)rr * row_bytes;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const uint2 w = *reinterpret_cast<const uint2*>(j < J - 1 ? sr + coff0 + j * (GT * 8) : sr + coff_last);
        p[rr][j][0] = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
        p[rr][j][1] = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
        ta = fma2(p[rr][j][0], rho[j][0], ta);
        tb = fma2(p[rr][j][1], rho[j][1], tb);
      }
      t[rr] = (ta.x + tb.x) + (ta.y + tb.y);
    }
    if (R == 2) {
      // both rows in one butterfly: after the first exchange lanes 0-15 carry row 0 and lanes 16-31 row 1
      const bool up = (lane & 16) != 0;
      float v = (up ? t[1] : t[0]) + __shfl_xor_sync(0xffffffffu, up ? t[0] : t[1], 16);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((lane & 15) == 0) red[grp][buf][gwarp][lane >> 4] = v;
    } else {
#pragma unroll
      for (int rr = 0; rr < R; ++rr) t[rr] = gx_warp_sum(t[rr]);
      if (lane == 0) {
#pragma unroll
        for (int rr = 0; rr < R; ++rr) red[grp][buf][gwarp][rr] = t[rr];
      }
    }
    group_bar(1 + grp, GT);   // the group has consumed the stage (and published its partial sums)
    if (gt == 0 && it + stages < n_iters) issue(it + stages, sg);
    float tot[R];
#pragma unroll
    for (int rr = 0; rr < R; ++rr) tot[rr] = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w)
#pragma unroll
      for (int rr = 0; rr < R; ++rr) tot[rr] += red[grp][buf][w][rr];
    buf ^= 1;
#pragma unroll
    for (int rr = 0; rr < R; ++rr) {
      if (rr < valid) {
        const float cn = cvec ? cvec[(long long)g * R + rr] : c_uniform;
        float rt;   // t' is a sum of positive terms of order 2^15: no range handling needed
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(tot[rr]));
        const float bn = cn * rt;
        const float2 b2 = make_float2(bn, bn);
#pragma unroll
        for (int j = 0; j < J; ++j) {
          acc[j][0] = fma2(p[rr][j][0], b2, acc[j][0]);
          acc[j][1] = fma2(p[rr][j][1], b2, acc[j][1]);
        }
      }
    }
    sg += NG;
    if (sg >= stages) { sg -= stages; phase ^= 1u; }
  }
  // acc holds sum_n e16_nk c_n / t'_n = a1_k * (the marginal of the unscaled matrix)
  float* prow = partials + ((long long)blockIdx.x * NG + grp) * k;
#pragma unroll
  for (int j = 0; j < J; ++j) {
    const int col = j * (GT * 4) + gt * 4;
    if (col < k) {
      const float4 l = *reinterpret_cast<const float4*>(la1 + col);
      *reinterpret_cast<float4*>(prow + col) = make_float4(acc[j][0].x * ex2_fast(-l.x), acc[j][0].y * ex2_fast(-l.y),
                                                           acc[j][1].x * ex2_fast(-l.z), acc[j][1].y * ex2_fast(-l.w));
    }
  }
}

// u[col] = sum_p parts[p, col]; 32 columns x 8 part-slices per block, fixed summation order
__global__ void __launch_bounds__(256)
colsum_parts_kernel(const float* __restrict__ parts, int nparts, int k, float* __restrict__ u) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (col < k)
    for (int p = ty; p < nparts; p += 8) acc += parts[(long long)p * k + col];
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < k) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    u[col] = t;
  }
}

// the same sums, pushed as tagged words into this rank's slot of the exchange block on every rank
__global__ void __launch_bounds__(256)
colsum_send_kernel(const float* __restrict__ parts, int nparts, int k, const gx_ll_desc ll,
                   float* __restrict__ u_local) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (col < k)
    for (int p = ty; p < nparts; p += 8) acc += parts[(long long)p * k + col];
  red[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && col < k) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][tx];
    if (u_local) u_local[col] = t;
    gxll::send1(ll, k, col, t);
  }
}

__global__ void ll_recv_sum_kernel(const gx_ll_desc ll, int k, float* __restrict__ u) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col < k) u[col] = gxll::recv1(ll, k, col);
}

__global__ void log_a_kernel(const float* __restrict__ u, const gx_ll_desc u_ll, const float* __restrict__ r, int k,
                             float* __restrict__ log_a) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= k) return;
  const float rk = r ? r[col] : 1.f / (float)k;
  const float uk = u_ll.world > 0 ? gxll::recv1(u_ll, k, col) : u[col];
  log_a[col] = logf(rk / uk);
}

// block-wide reductions of NV values at once (512 threads)
template <int NV, bool IS_MAX>
__device__ __forceinline__ void block_reduce(float (&v)[NV], float (*red)[SK_THREADS / 32]) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float w = IS_MAX ? gx_warp_max(v[i]) : gx_warp_sum(v[i]);
    if (lane == 0) red[i][warp] = w;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float tot = red[i][0];
#pragma unroll
    for (int w = 1; w < SK_THREADS / 32; ++w) tot = IS_MAX ? fmaxf(tot, red[i][w]) : tot + red[i][w];
    v[i] = tot;
  }
  __syncthreads();
}

// Q = softmax_k(S/eps + log_a)  (materialised only for API parity / tests)
template <int J>
__global__ void __launch_bounds__(SK_THREADS, 1)
sinkhorn_q_kernel(const float* __restrict__ s, long long n, int k, long long lds, float inv_eps,
                  const float* __restrict__ log_a, float* __restrict__ q) {
  __shared__ float red[1][SK_THREADS / 32];
  const int tid = threadIdx.x;
  for (long long row = blockIdx.x; row < n; row += gridDim.x) {
    float x[J][4];
    float mx[1] = {-INFINITY};
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int col = j * (SK_THREADS * 4) + tid * 4;
#pragma unroll
      for (int e = 0; e < 4; ++e) x[j][e] = -INFINITY;
      if (col < k) {
        const float4 v = *reinterpret_cast<const float4*>(s + row * lds + col);
        const float4 la = __ldg(reinterpret_cast<const float4*>(log_a + col));
        x[j][0] = fmaf(v.x, inv_eps, la.x); x[j][1] = fmaf(v.y, inv_eps, la.y);
        x[j][2] = fmaf(v.z, inv_eps, la.z); x[j][3] = fmaf(v.w, inv_eps, la.w);
#pragma unroll
        for (int e = 0; e < 4; ++e) mx[0] = fmaxf(mx[0], x[j][e]);
      }
    }
    block_reduce<1, true>(mx, red);
    float sm[1] = {0.f};
#pragma unroll
    for (int j = 0; j < J; ++j)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        x[j][e] = exp2f((x[j][e] - mx[0]) * LOG2E);
        sm[0] += x[j][e];
      }
    block_reduce<1, false>(sm, red);
    const float inv = 1.f / sm[0];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int col = j * (SK_THREADS * 4) + tid * 4;
      if (col < k)
        *reinterpret_cast<float4*>(q + row * (long long)k + col) =
            make_float4(x[j][0] * inv, x[j][1] * inv, x[j][2] * inv, x[j][3] * inv);
    }
  }
}

// ---------------------------------------------------------------------------
// Swapped-prediction loss, forward + d/dS fused.  Same TMA row ring / group structure as
// the Sinkhorn pass: one (S_s row, S_t row) pair per ring stage, one pair per group
// iteration.  All exponentials are ex2 of a single FFMA (log2-domain constants folded).
// ---------------------------------------------------------------------------
template <int NV, bool IS_MAX, int NW>
__device__ __forceinline__ void group_reduce(float (&v)[NV], float* red /* [NV][NW] */, int gwarp, int lane,
                                             int bar_id, int nthreads) {
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float w = IS_MAX ? gx_warp_max(v[i]) : gx_warp_sum(v[i]);
    if (lane == 0) red[i * NW + gwarp] = w;
  }
  group_bar(bar_id, nthreads);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float tot = IS_MAX ? -INFINITY : 0.f;
    if constexpr (NW % 4 == 0) {
#pragma unroll
      for (int w = 0; w < NW; w += 4) {
        const float4 q = *reinterpret_cast<const float4*>(red + i * NW + w);
        tot = IS_MAX ? fmaxf(tot, fmaxf(fmaxf(q.x, q.y), fmaxf(q.z, q.w))) : tot + ((q.x + q.y) + (q.z + q.w));
      }
    } else {
#pragma unroll
      for (int w = 0; w < NW; ++w) tot = IS_MAX ? fmaxf(tot, red[i * NW + w]) : tot + red[i * NW + w];
    }
    v[i] = tot;
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2_rn(float a, float b) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

template <int J, int GT, int NT>
__global__ void __launch_bounds__(NT, 1)
swav_loss_kernel(const float* __restrict__ ss, const float* __restrict__ st, long long n, int k, long long lds,
                 float inv_eps, float inv_temp, const float* __restrict__ la_s, const float* __restrict__ la_t,
                 float grad_scale, float* __restrict__ loss_parts, float* __restrict__ db_parts,
                 __nv_bfloat16* __restrict__ ds_s_hi, __nv_bfloat16* __restrict__ ds_s_lo,
                 __nv_bfloat16* __restrict__ ds_t_hi, __nv_bfloat16* __restrict__ ds_t_lo, long long ldd,
                 float* __restrict__ ds_s_f32, float* __restrict__ ds_t_f32, int stages) {
  constexpr int NG = NT / GT;
  constexpr int NW = GT / 32;
  extern __shared__ __align__(128) uint8_t sk_smem[];
  __shared__ __align__(16) float red_max[NG][4 * NW];
  __shared__ __align__(16) float red_sum[NG][6 * NW];
  __shared__ uint64_t full_bar[8];
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid / GT, gt = tid % GT, gwarp = gt >> 5;
  const uint32_t row_bytes = (uint32_t)k * 4u;
  const uint32_t stage_bytes = row_bytes * 2u;   // one row of S_s and one of S_t
  if (tid == 0) {
    for (int i = 0; i < stages; ++i) gxptx::mbar_init(&full_bar[i], 1);
    gxptx::fence_mbar_init();
  }
  __syncthreads();
  const int n_iters = (int)((n - blockIdx.x + gridDim.x - 1) / gridDim.x);
  auto issue = [&](int it, int sg) {   // sg == it % stages
    const long long row = (long long)blockIdx.x + (long long)it * gridDim.x;
    gxptx::mbar_arrive_expect_tx(&full_bar[sg], stage_bytes);
    gxptx::bulk_load_1d(sk_smem + (size_t)sg * stage_bytes, ss + row * lds, row_bytes, &full_bar[sg]);
    gxptx::bulk_load_1d(sk_smem + (size_t)sg * stage_bytes + row_bytes, st + row * lds, row_bytes, &full_bar[sg]);
  };
  if (tid == 0)
    for (int it = 0; it < stages && it < n_iters; ++it) issue(it, it);

  // log2-domain constants.  log2(a) of both views lives in shared memory behind the ring
  // (K floats each).  Out-of-range columns read a clamped address and are removed by a -inf
  // additive mask.
  const float ce = inv_eps * LOG2E, ct = inv_temp * LOG2E;
  float* sla_s = reinterpret_cast<float*>(sk_smem + (size_t)stages * stage_bytes);
  float* sla_t = sla_s + k;
  for (int i = tid; i < k; i += NT) {
    sla_s[i] = la_s[i] * LOG2E;
    sla_t[i] = la_t[i] * LOG2E;
  }
  __syncthreads();
  // Out-of-range column slots (col >= k) read a clamped, valid address; their contributions to
  // the sums are removed with a -inf additive mask and they store nothing.
  float2 db[J][2];
#pragma unroll
  for (int j = 0; j < J; ++j) db[j][0] = db[j][1] = make_float2(0.f, 0.f);
  const int col0 = gt * 4;
  const float2 ce2 = make_float2(ce, ce), ct2 = make_float2(ct, ct);
  float loss_acc = 0.f;
  const float gs = grad_scale * 0.5f * inv_temp;
  int sg = grp;   // == it % stages, with phase == (it / stages) & 1, kept incrementally (no integer division)
  uint32_t phase = 0;
  for (int it = grp; it < n_iters; it += NG, sg = sg + NG >= stages ? (phase ^= 1u, sg + NG - stages) : sg + NG) {
    const long long row = (long long)blockIdx.x + (long long)it * gridDim.x;
    gxptx::mbar_wait(&full_bar[sg], phase);
    const float* srow_s = reinterpret_cast<const float*>(sk_smem + (size_t)sg * stage_bytes);
    const float* srow_t = srow_s + k;
    float2 vs[J][2], vt[J][2];     // raw scores, later softmax(p) numerators
    float2 e1s[J][2], e1t[J][2];   // log2-domain S/eps + log a, later q numerators
    // pass 1: maxima.  Clamped slots re-read valid columns, which leaves the maxima unchanged.
    float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};  // max x1_s, max s_s, max x1_t, max s_t
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int cofs = min(col0 + j * (GT * 4), k - 4);
      const float4 a = *reinterpret_cast<const float4*>(srow_s + cofs);
      const float4 b = *reinterpret_cast<const float4*>(srow_t + cofs);
      const float4 ls = *reinterpret_cast<const float4*>(sla_s + cofs);
      const float4 lt = *reinterpret_cast<const float4*>(sla_t + cofs);
      vs[j][0] = make_float2(a.x, a.y); vs[j][1] = make_float2(a.z, a.w);
      vt[j][0] = make_float2(b.x, b.y); vt[j][1] = make_float2(b.z, b.w);
      e1s[j][0] = fma2(vs[j][0], ce2, make_float2(ls.x, ls.y));
      e1s[j][1] = fma2(vs[j][1], ce2, make_float2(ls.z, ls.w));
      e1t[j][0] = fma2(vt[j][0], ce2, make_float2(lt.x, lt.y));
      e1t[j][1] = fma2(vt[j][1], ce2, make_float2(lt.z, lt.w));
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        mx[0] = fmaxf(mx[0], fmaxf(e1s[j][h].x, e1s[j][h].y));
        mx[1] = fmaxf(mx[1], fmaxf(vs[j][h].x, vs[j][h].y));
        mx[2] = fmaxf(mx[2], fmaxf(e1t[j][h].x, e1t[j][h].y));
        mx[3] = fmaxf(mx[3], fmaxf(vt[j][h].x, vt[j][h].y));
      }
    }
    group_reduce<4, true, NW>(mx, red_max[grp], gwarp, lane, 1 + grp, GT);   // stage consumed after this barrier
    if (gt == 0 && it + stages < n_iters) issue(it + stages, sg);
    // pass 2: sums  Z1_s, Z2_s, D_st = sum e1s*s_t, Z1_t, Z2_t, D_ts = sum e1t*s_s   (packed partials)
    float2 s2[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) s2[i] = make_float2(0.f, 0.f);
    const float m2s = mx[1] * ct, m2t = mx[3] * ct;   // log2 domain
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const float mskj = (col0 + j * (GT * 4) < k) ? 0.f : -INFINITY;
      const float2 c1s = make_float2(mskj - mx[0], mskj - mx[0]), c1t = make_float2(mskj - mx[2], mskj - mx[2]);
      const float2 c2s = make_float2(mskj - m2s, mskj - m2s), c2t = make_float2(mskj - m2t, mskj - m2t);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 rs = vs[j][h], rt = vt[j][h];
        const float2 q1s = ex2_2(add2(e1s[j][h], c1s));
        const float2 q1t = ex2_2(add2(e1t[j][h], c1t));
        const float2 p2s = ex2_2(fma2(rs, ct2, c2s));
        const float2 p2t = ex2_2(fma2(rt, ct2, c2t));
        s2[0] = add2(s2[0], q1s);
        s2[3] = add2(s2[3], q1t);
        s2[2] = fma2(q1s, rt, s2[2]);
        s2[5] = fma2(q1t, rs, s2[5]);
        s2[1] = add2(s2[1], p2s);
        s2[4] = add2(s2[4], p2t);
        e1s[j][h] = q1s; e1t[j][h] = q1t;
        vs[j][h] = p2s; vt[j][h] = p2t;
      }
    }
    float sm[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) sm[i] = s2[i].x + s2[i].y;
    group_reduce<6, false, NW>(sm, red_sum[grp], gwarp, lane, 1 + grp, GT);
    const float iz1s = __fdividef(1.f, sm[0]), iz2s = __fdividef(1.f, sm[1]);
    const float iz1t = __fdividef(1.f, sm[3]), iz2t = __fdividef(1.f, sm[4]);
    if (gt == 0) {
      const float ln2 = 0.69314718055994531f;
      const float lse_s = (m2s + log2f(sm[1])) * ln2, lse_t = (m2t + log2f(sm[4])) * ln2;
      const float qs_pt = sm[2] * iz1s * inv_temp - lse_t;  // sum_k q_s * log_softmax(p_t)
      const float qt_ps = sm[5] * iz1t * inv_temp - lse_s;
      loss_acc += -0.5f * (qs_pt + qt_ps);
    }
    // pass 3: gradients.  dL/dS_s uses q_t ; dL/dS_t uses q_s
    const float2 a2s = make_float2(gs * iz2s, gs * iz2s), a1t = make_float2(-gs * iz1t, -gs * iz1t);
    const float2 a2t = make_float2(gs * iz2t, gs * iz2t), a1s = make_float2(-gs * iz1s, -gs * iz1s);
    __nv_bfloat16* ps_hi = ds_s_hi + row * ldd;
    __nv_bfloat16* pt_hi = ds_t_hi + row * ldd;
    const bool want_lo = (ds_s_lo != nullptr) || (ds_t_lo != nullptr);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int col = col0 + j * (GT * 4);
      if (col < k) {
        const float2 gs0 = fma2(vs[j][0], a2s, mul2(e1t[j][0], a1t)), gs1 = fma2(vs[j][1], a2s, mul2(e1t[j][1], a1t));
        const float2 gt0 = fma2(vt[j][0], a2t, mul2(e1s[j][0], a1s)), gt1 = fma2(vt[j][1], a2t, mul2(e1s[j][1], a1s));
        db[j][0] = add2(db[j][0], add2(gs0, gt0));
        db[j][1] = add2(db[j][1], add2(gs1, gt1));
        const float4 g_s = make_float4(gs0.x, gs0.y, gs1.x, gs1.y);
        const float4 g_t = make_float4(gt0.x, gt0.y, gt1.x, gt1.y);
        if (want_lo) {
          uint2 h, l;
          gx_split4(g_s, h, l);
          *reinterpret_cast<uint2*>(ps_hi + col) = h;
          if (ds_s_lo) *reinterpret_cast<uint2*>(ds_s_lo + row * ldd + col) = l;
          gx_split4(g_t, h, l);
          *reinterpret_cast<uint2*>(pt_hi + col) = h;
          if (ds_t_lo) *reinterpret_cast<uint2*>(ds_t_lo + row * ldd + col) = l;
        } else {
          *reinterpret_cast<uint2*>(ps_hi + col) =
              make_uint2(pack_bf16x2_rn(g_s.x, g_s.y), pack_bf16x2_rn(g_s.z, g_s.w));
          *reinterpret_cast<uint2*>(pt_hi + col) =
              make_uint2(pack_bf16x2_rn(g_t.x, g_t.y), pack_bf16x2_rn(g_t.z, g_t.w));
        }
        if (ds_s_f32) *reinterpret_cast<float4*>(ds_s_f32 + row * (long long)k + col) = g_s;
        if (ds_t_f32) *reinterpret_cast<float4*>(ds_t_f32 + row * (long long)k + col) = g_t;
      }
    }
  }
  if (gt == 0) loss_parts[blockIdx.x * NG + grp] = loss_acc;
  if (db_parts) {
    float* prow = db_parts + ((long long)blockIdx.x * NG + grp) * k;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int col = j * (GT * 4) + gt * 4;
      if (col < k)
        *reinterpret_cast<float4*>(prow + col) = make_float4(db[j][0].x, db[j][0].y, db[j][1].x, db[j][1].y);
    }
  }
}

// ---------------------------------------------------------------------------
// Same stage when T / eps is 1 or 2 (every shipped config: eps 0.005 / 0.01, T 0.01).  Then
//   exp(S/eps + log a_k) = exp(S/T)^RHO * a_k,
// so the Sinkhorn code numerators are products of the softmax(S/T) numerators that are needed anyway:
// 2 instead of 4 `ex2` per score pair, no maxima for the q branch (the common factors cancel in the row
// normalisation), and only two register arrays per thread.  a_k is stored in shared memory divided by
// its maximum, so every numerator is <= 1.
// ---------------------------------------------------------------------------
template <int J, int GT, int NT, int RHO>
__global__ void __launch_bounds__(NT, 1)
swav_loss_pow_kernel(const float* __restrict__ ss, const float* __restrict__ st, long long n, int k, long long lds,
                     float inv_temp, const float* __restrict__ la_s, const float* __restrict__ la_t,
                     float grad_scale, float* __restrict__ loss_parts, float* __restrict__ db_parts,
                     __nv_bfloat16* __restrict__ ds_s_hi, __nv_bfloat16* __restrict__ ds_s_lo,
                     __nv_bfloat16* __restrict__ ds_t_hi, __nv_bfloat16* __restrict__ ds_t_lo, long long ldd,
                     float* __restrict__ ds_s_f32, float* __restrict__ ds_t_f32, int stages) {
  constexpr int NG = NT / GT;
  constexpr int NW = GT / 32;
  extern __shared__ __align__(128) uint8_t sk_smem[];
  __shared__ __align__(16) float red_max[NG][2 * NW];
  __shared__ __align__(16) float red_sum[NG][6 * NW];
  __shared__ float red_la[2][NT / 32];
  __shared__ uint64_t full_bar[8];
  const int tid = threadIdx.x, lane = tid & 31;
  const int grp = tid / GT, gt = tid % GT, gwarp = gt >> 5;
  const uint32_t row_bytes = (uint32_t)k * 4u;
  const uint32_t stage_bytes = row_bytes * 2u;   // one row of S_s and one of S_t
  if (tid == 0) {
    for (int i = 0; i < stages; ++i) gxptx::mbar_init(&full_bar[i], 1);
    gxptx::fence_mbar_init();
  }
  __syncthreads();
  const int n_iters = (int)((n - blockIdx.x + gridDim.x - 1) / gridDim.x);
  auto issue = [&](int it, int sg) {   // sg == it % stages
    const long long row = (long long)blockIdx.x + (long long)it * gridDim.x;
    gxptx::mbar_arrive_expect_tx(&full_bar[sg], stage_bytes);
    gxptx::bulk_load_1d(sk_smem + (size_t)sg * stage_bytes, ss + row * lds, row_bytes, &full_bar[sg]);
    gxptx::bulk_load_1d(sk_smem + (size_t)sg * stage_bytes + row_bytes, st + row * lds, row_bytes, &full_bar[sg]);
  };
  if (tid == 0)
    for (int it = 0; it < stages && it < n_iters; ++it) issue(it, it);

  // a_k / max_k a_k of both views behind the ring (K floats each)
  float* sa_s = reinterpret_cast<float*>(sk_smem + (size_t)stages * stage_bytes);
  float* sa_t = sa_s + k;
  {
    float ms = -INFINITY, mt = -INFINITY;
    for (int i = tid; i < k; i += NT) {
      ms = fmaxf(ms, la_s[i]);
      mt = fmaxf(mt, la_t[i]);
    }
    ms = gx_warp_max(ms);
    mt = gx_warp_max(mt);
    if (lane == 0) { red_la[0][tid >> 5] = ms; red_la[1][tid >> 5] = mt; }
    __syncthreads();
    ms = mt = -INFINITY;
    for (int w = 0; w < NT / 32; ++w) { ms = fmaxf(ms, red_la[0][w]); mt = fmaxf(mt, red_la[1][w]); }
    for (int i = tid; i < k; i += NT) {
      sa_s[i] = ex2_fast((la_s[i] - ms) * LOG2E);
      sa_t[i] = ex2_fast((la_t[i] - mt) * LOG2E);
    }
  }
  __syncthreads();
  const float ct = inv_temp * LOG2E;
  float2 db[J][2];
#pragma unroll
  for (int j = 0; j < J; ++j) db[j][0] = db[j][1] = make_float2(0.f, 0.f);
  const int col0 = gt * 4;
  const float2 ct2 = make_float2(ct, ct);
  float loss_acc = 0.f;
  const float gs = grad_scale * 0.5f * inv_temp;
  int sg = grp;   // == it % stages, with phase == (it / stages) & 1, kept incrementally (no integer division)
  uint32_t phase = 0;
  for (int it = grp; it < n_iters; it += NG, sg = sg + NG >= stages ? (phase ^= 1u, sg + NG - stages) : sg + NG) {
    const long long row = (long long)blockIdx.x + (long long)it * gridDim.x;
    gxptx::mbar_wait(&full_bar[sg], phase);
    const float* srow_s = reinterpret_cast<const float*>(sk_smem + (size_t)sg * stage_bytes);
    const float* srow_t = srow_s + k;
    float2 vs[J][2], vt[J][2];     // raw scores, later softmax(S/T) numerators
    // pass 1: row maxima of the raw scores (clamped slots re-read valid columns)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int cofs = min(col0 + j * (GT * 4), k - 4);
      const float4 a = *reinterpret_cast<const float4*>(srow_s + cofs);
      const float4 b = *reinterpret_cast<const float4*>(srow_t + cofs);
      vs[j][0] = make_float2(a.x, a.y); vs[j][1] = make_float2(a.z, a.w);
      vt[j][0] = make_float2(b.x, b.y); vt[j][1] = make_float2(b.z, b.w);
      mx[0] = fmaxf(mx[0], fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)));
      mx[1] = fmaxf(mx[1], fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
    }
    group_reduce<2, true, NW>(mx, red_max[grp], gwarp, lane, 1 + grp, GT);   // stage consumed after this barrier
    if (gt == 0 && it + stages < n_iters) issue(it + stages, sg);
    // pass 2: sums  Z1_s, Z2_s, D_st = sum q_s*s_t, Z1_t, Z2_t, D_ts = sum q_t*s_s   (packed partials)
    float2 s2[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) s2[i] = make_float2(0.f, 0.f);
    const float m2s = mx[0] * ct, m2t = mx[1] * ct;   // log2 domain
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int cofs = min(col0 + j * (GT * 4), k - 4);
      const float mskj = (col0 + j * (GT * 4) < k) ? 0.f : -INFINITY;
      const float2 c2s = make_float2(mskj - m2s, mskj - m2s), c2t = make_float2(mskj - m2t, mskj - m2t);
      const float4 as4 = *reinterpret_cast<const float4*>(sa_s + cofs);
      const float4 at4 = *reinterpret_cast<const float4*>(sa_t + cofs);
      const float2 as2[2] = {make_float2(as4.x, as4.y), make_float2(as4.z, as4.w)};
      const float2 at2[2] = {make_float2(at4.x, at4.y), make_float2(at4.z, at4.w)};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float2 rs = vs[j][h], rt = vt[j][h];
        const float2 p2s = ex2_2(fma2(rs, ct2, c2s));
        const float2 p2t = ex2_2(fma2(rt, ct2, c2t));
        const float2 q1s = mul2(RHO == 2 ? mul2(p2s, p2s) : p2s, as2[h]);
        const float2 q1t = mul2(RHO == 2 ? mul2(p2t, p2t) : p2t, at2[h]);
        s2[0] = add2(s2[0], q1s);
        s2[3] = add2(s2[3], q1t);
        s2[2] = fma2(q1s, rt, s2[2]);
        s2[5] = fma2(q1t, rs, s2[5]);
        s2[1] = add2(s2[1], p2s);
        s2[4] = add2(s2[4], p2t);
        vs[j][h] = p2s; vt[j][h] = p2t;
      }
    }
    float sm[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) sm[i] = s2[i].x + s2[i].y;
    group_reduce<6, false, NW>(sm, red_sum[grp], gwarp, lane, 1 + grp, GT);
    const float iz1s = __fdividef(1.f, sm[0]), iz2s = __fdividef(1.f, sm[1]);
    const float iz1t = __fdividef(1.f, sm[3]), iz2t = __fdividef(1.f, sm[4]);
    if (gt == 0) {
      const float ln2 = 0.69314718055994531f;
      const float lse_s = (m2s + log2f(sm[1])) * ln2, lse_t = (m2t + log2f(sm[4])) * ln2;
      const float qs_pt = sm[2] * iz1s * inv_temp - lse_t;  // sum_k q_s * log_softmax(p_t)
      const float qt_ps = sm[5] * iz1t * inv_temp - lse_s;
      loss_acc += -0.5f * (qs_pt + qt_ps);
    }
    // pass 3: gradients.  dL/dS_s uses q_t ; dL/dS_t uses q_s
    const float2 a2s = make_float2(gs * iz2s, gs * iz2s), a1t = make_float2(-gs * iz1t, -gs * iz1t);
    const float2 a2t = make_float2(gs * iz2t, gs * iz2t), a1s = make_float2(-gs * iz1s, -gs * iz1s);
    __nv_bfloat16* ps_hi = ds_s_hi + row * ldd;
    __nv_bfloat16* pt_hi = ds_t_hi + row * ldd;
    const bool want_lo = (ds_s_lo != nullptr) || (ds_t_lo != nullptr);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int col = col0 + j * (GT * 4);
      if (col < k) {
        const float4 as4 = *reinterpret_cast<const float4*>(sa_s + col);
        const float4 at4 = *reinterpret_cast<const float4*>(sa_t + col);
        // q numerators again from the p numerators in registers (two multiplies)
        const float2 qs0 = mul2(RHO == 2 ? mul2(vs[j][0], vs[j][0]) : vs[j][0], mul2(make_float2(as4.x, as4.y), a1s));
        const float2 qs1 = mul2(RHO == 2 ? mul2(vs[j][1], vs[j][1]) : vs[j][1], mul2(make_float2(as4.z, as4.w), a1s));
        const float2 qt0 = mul2(RHO == 2 ? mul2(vt[j][0], vt[j][0]) : vt[j][0], mul2(make_float2(at4.x, at4.y), a1t));
        const float2 qt1 = mul2(RHO == 2 ? mul2(vt[j][1], vt[j][1]) : vt[j][1], mul2(make_float2(at4.z, at4.w), a1t));
        const float2 gs0 = fma2(vs[j][0], a2s, qt0), gs1 = fma2(vs[j][1], a2s, qt1);
        const float2 gt0 = fma2(vt[j][0], a2t, qs0), gt1 = fma2(vt[j][1], a2t, qs1);
        db[j][0] = add2(db[j][0], add2(gs0, gt0));
        db[j][1] = add2(db[j][1], add2(gs1, gt1));
        const float4 g_s = make_float4(gs0.x, gs0.y, gs1.x, gs1.y);
        const float4 g_t = make_float4(gt0.x, gt0.y, gt1.x, gt1.y);
        if (want_lo) {
          uint2 h, l;
          gx_split4(g_s, h, l);
          *reinterpret_cast<uint2*>(ps_hi + col) = h;
          if (ds_s_lo) *reinterpret_cast<uint2*>(ds_s_lo + row * ldd + col) = l;
          gx_split4(g_t, h, l);
          *reinterpret_cast<uint2*>(pt_hi + col) = h;
          if (ds_t_lo) *reinterpret_cast<uint2*>(ds_t_lo + row * ldd + col) = l;
        } else {
          *reinterpret_cast<uint2*>(ps_hi + col) =
              make_uint2(pack_bf16x2_rn(g_s.x, g_s.y), pack_bf16x2_rn(g_s.z, g_s.w));
          *reinterpret_cast<uint2*>(pt_hi + col) =
              make_uint2(pack_bf16x2_rn(g_t.x, g_t.y), pack_bf16x2_rn(g_t.z, g_t.w));
        }
        if (ds_s_f32) *reinterpret_cast<float4*>(ds_s_f32 + row * (long long)k + col) = g_s;
        if (ds_t_f32) *reinterpret_cast<float4*>(ds_t_f32 + row * (long long)k + col) = g_t;
      }
    }
  }
  if (gt == 0) loss_parts[blockIdx.x * NG + grp] = loss_acc;
  if (db_parts) {
    float* prow = db_parts + ((long long)blockIdx.x * NG + grp) * k;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int col = j * (GT * 4) + gt * 4;
      if (col < k)
        *reinterpret_cast<float4*>(prow + col) = make_float4(db[j][0].x, db[j][0].y, db[j][1].x, db[j][1].y);
    }
  }
}

// ---------------------------------------------------------------------------
// LARC + SGD(momentum)
// ---------------------------------------------------------------------------
// Deterministic two-stage norms: per-block partial sums (no atomics), folded in a fixed order by every block of the
// update kernel - the LARC trust ratio is then bit-identical from run to run and from rank to rank, which keeps
// data-parallel replicas in lock-step (float atomics would let them drift apart by an ulp per step).
__global__ void sq_norms_kernel(const float* __restrict__ p, const float* __restrict__ g, long long n,
                                float* __restrict__ parts) {
  float sp = 0.f, sg = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    sp = fmaf(p[i], p[i], sp);
    sg = fmaf(g[i], g[i], sg);
  }
  sp = gx_warp_sum(sp);
  sg = gx_warp_sum(sg);
  __shared__ float rp[8], rg[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { rp[warp] = sp; rg[warp] = sg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { a += rp[w]; b += rg[w]; }
    parts[2 * blockIdx.x] = a;
    parts[2 * blockIdx.x + 1] = b;
  }
}

__global__ void larc_sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ buf,
                                long long n, float lr, float momentum, float trust, float wd, float eps, int first,
                                const float* __restrict__ parts, int nparts) {
  __shared__ float tot[2];
  if (threadIdx.x < 32) {
    float a = 0.f, b = 0.f;
    for (int i = threadIdx.x; i < nparts; i += 32) { a += parts[2 * i]; b += parts[2 * i + 1]; }
    a = gx_warp_sum(a);
    b = gx_warp_sum(b);
    if (threadIdx.x == 0) { tot[0] = a; tot[1] = b; }
  }
  __syncthreads();
  const float pn = sqrtf(tot[0]), gn = sqrtf(tot[1]);
  const bool adapt = (pn != 0.f) && (gn != 0.f);
  const float alr = adapt ? trust * pn / (gn + pn * wd + eps) : 1.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float gi = g[i];
    if (adapt) gi = (gi + wd * p[i]) * alr;
    const float b = first ? gi : momentum * buf[i] + gi;
    buf[i] = b;
    p[i] -= lr * b;
  }
}

// ---------------------------------------------------------------------------
// label maps
// ---------------------------------------------------------------------------
__global__ void argmax_rows_kernel(const float* __restrict__ x, long long n, int c, long long ldx,
                                   long long* __restrict__ labels) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* xr = x + row * ldx;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = lane; i < c; i += 32) {
    const float v = xr[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) labels[row] = (long long)bi;
}

// x = concat(x1 [n,c1], x2 [n,c2]) along channels (x2 may be null): the two maps of one
// resolution are never concatenated in memory
__global__ void kmeans_assign_kernel(const float* __restrict__ x1, int c1, const float* __restrict__ x2, int c2,
                                     long long n, const float* __restrict__ centers, int k,
                                     int* __restrict__ labels, float* __restrict__ dist) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const float* xr1 = x1 + row * c1;
  const float* xr2 = x2 ? x2 + row * c2 : nullptr;
  const int c = c1 + c2;
  float best = INFINITY;
  int bi = 0;
  for (int kk = 0; kk < k; ++kk) {
    const float* cr = centers + (long long)kk * c;
    float d = 0.f;
    for (int i = lane; i < c1; i += 32) {
      const float t = xr1[i] - __ldg(cr + i);
      d = fmaf(t, t, d);
    }
    for (int i = lane; i < c2; i += 32) {
      const float t = xr2[i] - __ldg(cr + c1 + i);
      d = fmaf(t, t, d);
    }
    d = gx_warp_sum(d);
    if (d < best) { best = d; bi = kk; }
  }
  if (lane == 0) {
    if (labels) labels[row] = bi;
    if (dist) dist[row] = best;      // squared distance to the assigned centre (k-means fit: inertia, k-means++)
  }
}

// labels[n] = first argmin_k (bias[k] + scale * s[n,k]): nearest centre from the x.c scores of the tensor-core path,
// ||x - c||^2 = ||x||^2 - 2 x.c + ||c||^2 (the ||x||^2 term does not change the arg-min).  One warp per row.
__global__ void argmin_affine_kernel(const float* __restrict__ s, long long n, int k, long long lds,
                                     const float* __restrict__ bias, float scale, int* __restrict__ labels) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  float best = INFINITY;
  int bi = 0x7fffffff;
  for (int kk = lane; kk < k; kk += 32) {
    const float v = fmaf(scale, s[row * lds + kk], __ldg(bias + kk));
    if (v < best) { best = v; bi = kk; }           // kk ascends: the first minimum of this lane is kept
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov < best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if (lane == 0) labels[row] = bi == 0x7fffffff ? 0 : bi;
}

// one-hot cluster maps resized with nearest neighbour (ref hfc_kmeans_clustering.py:190-206)
__global__ void onehot_nearest_kernel(const int* __restrict__ labels, int b, int h, int w, int k, int oh, int ow,
                                      long long out_bstride, int vw, float on, float off, float* __restrict__ out) {
  // vw = 4: four consecutive output columns per thread (one 128-bit store; the host checks ow % 4 and alignment);
  // out[bb] starts at bb * out_bstride so that a layer writes its K channels straight into the concatenated
  // [B, sum K, oh, ow] maps
  const int owv = ow / vw;
  const long long total = (long long)b * k * oh * owv;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int oxv = (int)(r % owv); r /= owv;
    const int oy = (int)(r % oh); r /= oh;
    const int kk = (int)(r % k);
    const int bb = (int)(r / k);
    const int sy = (int)(((long long)oy * h) / oh);
    const int* lrow = labels + ((long long)bb * h + sy) * w;
    float* dst = out + (long long)bb * out_bstride + ((long long)kk * oh + oy) * ow + (long long)oxv * vw;
    if (vw == 4) {
      float4 v;
      v.x = lrow[(int)(((long long)(oxv * 4 + 0) * w) / ow)] == kk ? on : off;
      v.y = lrow[(int)(((long long)(oxv * 4 + 1) * w) / ow)] == kk ? on : off;
      v.z = lrow[(int)(((long long)(oxv * 4 + 2) * w) / ow)] == kk ? on : off;
      v.w = lrow[(int)(((long long)(oxv * 4 + 3) * w) / ow)] == kk ? on : off;
      *reinterpret_cast<float4*>(dst) = v;
    } else {
      dst[0] = lrow[(int)(((long long)oxv * w) / ow)] == kk ? on : off;
    }
  }
}

inline int sk_grid(long long n, int rows_per_iter) {
  long long g = (n + rows_per_iter - 1) / rows_per_iter;
  const int cap = gx_stream_cta_budget();
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

extern "C" int gx_gather_rows(const gx_gather_desc* d, void* stream) {
  GX_CHECK_ARG(d && (d->a_hi || d->a_f32 || d->row_norm) && d->nlevels > 0 && d->nlevels <= GX_MAX_LEVELS &&
               d->nrows > 0);
  GX_CHECK_ARG(d->a_lo == nullptr || d->a_hi != nullptr);
  GX_CHECK_ARG(d->hlen % 4 == 0 && d->ld % 4 == 0 && d->ld >= d->hlen);
  GX_CHECK_ARG(d->nrows < (1ll << 31));
  for (int l = 0; l < d->nlevels; ++l) GX_CHECK_ARG(d->feat[l] && d->c[l] % 4 == 0 && d->h[l] > 0 && d->w[l] > 0);
  gather_rows_kernel<<<(unsigned)d->nrows, 128, 0, (cudaStream_t)stream>>>(*d);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_l2norm_split(const float* z, const int* row_idx, void* zn_hi, void* zn_lo, void* zn_f16,
                               float* inv_norm, long long n, int c, void* stream) {
  GX_CHECK_ARG(z && zn_hi && n > 0 && c % 4 == 0);
  l2norm_split_kernel<<<gx_cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(
      z, row_idx, reinterpret_cast<__nv_bfloat16*>(zn_hi), reinterpret_cast<__nv_bfloat16*>(zn_lo),
      reinterpret_cast<__half*>(zn_f16), inv_norm, n, c);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_l2norm_bwd_split(const float* dzn, const void* zn_hi, const void* zn_lo, const float* inv_norm,
                                   void* dz_hi, void* dz_lo, float* dz_f32, long long n, int c, void* stream) {
  GX_CHECK_ARG(dzn && zn_hi && inv_norm && (dz_hi || dz_f32) && n > 0 && c % 4 == 0);
  GX_CHECK_ARG(dz_lo == nullptr || dz_hi != nullptr);
  l2norm_bwd_split_kernel<<<gx_cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(
      dzn, reinterpret_cast<const __nv_bfloat16*>(zn_hi), reinterpret_cast<const __nv_bfloat16*>(zn_lo), inv_norm,
      reinterpret_cast<__nv_bfloat16*>(dz_hi), reinterpret_cast<__nv_bfloat16*>(dz_lo), dz_f32, n, c);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_segment_sum_rows(const void* rows, int rows_bf16, const int* order, const int* seg_off, void* hi,
                                   void* lo, float* out_f32, long long nseg, int c, void* stream) {
  GX_CHECK_ARG(rows && order && seg_off && (hi || out_f32) && nseg > 0 && c % 4 == 0);
  GX_CHECK_ARG(lo == nullptr || hi != nullptr);
  if (rows_bf16)
    segment_sum_rows_kernel<true><<<gx_cdiv(nseg, 8), 256, 0, (cudaStream_t)stream>>>(
        rows, order, seg_off, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), out_f32, nseg,
        c);
  else
    segment_sum_rows_kernel<false><<<gx_cdiv(nseg, 8), 256, 0, (cudaStream_t)stream>>>(
      rows, order, seg_off, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), out_f32, nseg,
      c);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_pool1d_bilinear(const float* in, long long outer, int n_in, int n_out, long long inner, float* out,
                                  void* stream) {
  GX_CHECK_ARG(in && out && outer > 0 && n_in > 0 && n_out > 0 && n_out <= n_in && inner > 0 && inner % 4 == 0);
  const long long warps = outer * n_out * (((inner >> 2) + 127) / 128);
  GX_CHECK_ARG(gx_cdiv(warps, 8) < 2147483647LL);
  pool1d_bilinear_kernel<<<(unsigned)gx_cdiv(warps, 8), 256, 0, (cudaStream_t)stream>>>(in, outer, n_in, n_out, inner,
                                                                                       out);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_upsample_sum(int nlevels, const float* const* p, const int* h, const int* w, int batch, int out_h,
                               int out_w, int c, float* out, void* hi, void* lo, long long* labels, int bilinear,
                               void* stream) {
  GX_CHECK_ARG(nlevels > 0 && nlevels <= GX_MAX_LEVELS && p && h && w && (out || hi || labels) && batch > 0 &&
               c % 4 == 0);
  GX_CHECK_ARG(lo == nullptr || hi != nullptr);
  UpsumDesc d;
  d.nlevels = nlevels;
  for (int l = 0; l < nlevels; ++l) {
    GX_CHECK_ARG(p[l] && h[l] > 0 && w[l] > 0);
    d.p[l] = p[l]; d.h[l] = h[l]; d.w[l] = w[l];
  }
  d.out_h = out_h; d.out_w = out_w; d.c = c;
  d.bilinear = bilinear ? 1 : 0;
  d.npix = (long long)batch * out_h * out_w;
  // pyramid fast path: leading levels shared by aligned 2x2 output quads, at most UPQ_FINE per-pixel levels after them
  int n_shared = 0;
  bool quad_ok = !bilinear && out_h % 2 == 0 && out_w % 2 == 0 && (long long)out_h * out_w * (c / 4) < (1LL << 31);
  if (quad_ok) {
    auto shared = [&](int l) {
      return out_h % h[l] == 0 && (out_h / h[l]) % 2 == 0 && out_w % w[l] == 0 && (out_w / w[l]) % 2 == 0;
    };
    while (n_shared < nlevels && shared(n_shared)) ++n_shared;
    quad_ok = n_shared <= UPQ_SHARED && nlevels - n_shared <= UPQ_FINE;
    for (int l = n_shared; l < nlevels && quad_ok; ++l)
      quad_ok = (long long)h[l] * w[l] * (c / 4) < (1LL << 31);
  }
  const char* quad_env = getenv("GX_UPSUM_QUAD");   // GX_UPSUM_QUAD=0: per-pixel kernel (A/B timing)
  if (quad_env && atoi(quad_env) == 0) quad_ok = false;
  if (quad_ok) {
    const dim3 qgrid(gx_cdiv(d.npix / 4, 8));
    __nv_bfloat16* bhi = reinterpret_cast<__nv_bfloat16*>(hi);
    __nv_bfloat16* blo = reinterpret_cast<__nv_bfloat16*>(lo);
    cudaStream_t cs = (cudaStream_t)stream;
    const bool small = n_shared <= 6 && nlevels - n_shared <= 1;
    if (small && labels)
      upsample_sum_quad_kernel<6, 1, true><<<qgrid, 256, 0, cs>>>(d, n_shared, out, bhi, blo, labels);
    else if (small)
      upsample_sum_quad_kernel<6, 1, false><<<qgrid, 256, 0, cs>>>(d, n_shared, out, bhi, blo, nullptr);
    else if (labels)
      upsample_sum_quad_kernel<UPQ_SHARED, UPQ_FINE, true><<<qgrid, 256, 0, cs>>>(d, n_shared, out, bhi, blo, labels);
    else
      upsample_sum_quad_kernel<UPQ_SHARED, UPQ_FINE, false><<<qgrid, 256, 0, cs>>>(d, n_shared, out, bhi, blo, nullptr);
  } else {
    upsample_sum_kernel<<<gx_cdiv(d.npix, 8), 256, 0, (cudaStream_t)stream>>>(
        d, out, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), labels);
  }
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_tap_sum(const float* g, int batch, int h, int w, int cout, int dilation, const float* bias, int act,
                          float* out, void* next_hi, void* next_lo, int next_ld, void* stream) {
  GX_CHECK_ARG(g && (out || next_hi) && batch > 0 && h > 0 && w > 0 && cout > 0 && cout % 4 == 0 && dilation >= 1);
  GX_CHECK_ARG(act == 0 || act == 2);
  GX_CHECK_ARG(next_lo == nullptr || next_hi != nullptr);
  GX_CHECK_ARG(next_hi == nullptr || (next_ld >= cout && next_ld % 4 == 0));
  const long long total = (long long)batch * h * w * (cout / 4);
  int grid = gx_cdiv(total, 256);
  const int cap = gx_sm_count() * 32;
  if (grid > cap) grid = cap;
  tap_sum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g, batch, h, w, cout, dilation, bias, act, out,
                                                         reinterpret_cast<__nv_bfloat16*>(next_hi),
                                                         reinterpret_cast<__nv_bfloat16*>(next_lo), next_ld);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_tap_spread(const float* dout, int batch, int h, int w, int cout, int dilation, void* dg_hi,
                             void* dg_lo, void* stream) {
  GX_CHECK_ARG(dout && dg_hi && batch > 0 && h > 0 && w > 0 && cout > 0 && cout % 4 == 0 && dilation >= 1);
  const long long total = (long long)batch * h * w * 9 * (cout / 4);
  int grid = gx_cdiv(total, 256);
  const int cap = gx_sm_count() * 32;
  if (grid > cap) grid = cap;
  tap_spread_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dout, batch, h, w, cout, dilation,
                                                            reinterpret_cast<__nv_bfloat16*>(dg_hi),
                                                            reinterpret_cast<__nv_bfloat16*>(dg_lo));
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_pool_sum(const float* in, int batch, int in_h, int in_w, int out_h, int out_w, int c, float* out,
                           void* hi, void* lo, void* stream) {
  GX_CHECK_ARG(in && (out || hi) && batch > 0 && c % 4 == 0 && out_h > 0 && out_w > 0);
  GX_CHECK_ARG(in_h % out_h == 0 && in_w % out_w == 0);
  GX_CHECK_ARG(lo == nullptr || hi != nullptr);
  const long long nout = (long long)batch * out_h * out_w;
  pool_sum_kernel<<<gx_cdiv(nout, 8), 256, 0, (cudaStream_t)stream>>>(
      in, in_h, in_w, out_h, out_w, c, nout, out, reinterpret_cast<__nv_bfloat16*>(hi),
      reinterpret_cast<__nv_bfloat16*>(lo));
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_normalize_rows(float* w, long long rows, int cols, void* stream) {
  GX_CHECK_ARG(w && rows > 0 && cols > 0);
  normalize_rows_kernel<<<gx_cdiv(rows, 8), 256, 0, (cudaStream_t)stream>>>(w, rows, cols);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_split_planes(const float* x, long long ld, void* hi, void* lo, long long rows, long long cols,
                               int transpose, long long ld_out, void* stream) {
  GX_CHECK_ARG(x && hi && rows > 0 && cols > 0 && ld >= cols);
  if (ld_out <= 0) ld_out = transpose ? rows : cols;
  if (!transpose) {
    GX_CHECK_ARG(ld_out >= cols);
    const bool v4 = cols % 4 == 0 && ld % 4 == 0 && ld_out % 4 == 0 && ((uintptr_t)x & 15) == 0 &&
                    ((uintptr_t)hi & 7) == 0 && ((uintptr_t)lo & 7) == 0;
    const long long work = v4 ? rows * (cols / 4) : rows * cols;
    int grid = gx_cdiv(work, 256);
    const int cap = gx_sm_count() * 16;
    if (grid > cap) grid = cap;
    if (v4)
      split_planes_v4_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
          x, ld, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), rows, (int)(cols / 4),
          ld_out);
    else
      split_planes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ld, reinterpret_cast<__nv_bfloat16*>(hi),
                                                                 reinterpret_cast<__nv_bfloat16*>(lo), rows, cols,
                                                                 ld_out);
  } else {
    GX_CHECK_ARG(ld_out == rows);
    dim3 grid(gx_cdiv(cols, 32), gx_cdiv(rows, 32));
    GX_CHECK_ARG(grid.y <= 65535);
    split_planes_t_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(
        x, ld, reinterpret_cast<__nv_bfloat16*>(hi), reinterpret_cast<__nv_bfloat16*>(lo), rows, cols);
  }
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_round_f16(const float* x, long long ld, void* out, long long rows, long long cols, void* stream) {
  GX_CHECK_ARG(x && out && rows > 0 && cols > 0 && ld >= cols);
  int grid = gx_cdiv(rows * cols, 256);
  const int cap = gx_sm_count() * 16;
  if (grid > cap) grid = cap;
  round_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, ld, reinterpret_cast<__half*>(out), rows, cols);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_sinkhorn_max_parts(void) { return 2 * gx_sm_count(); }
extern "C" int gx_loss_max_parts(void) { return 2 * gx_sm_count(); }

#define GX_DISPATCH_J(J_, ...)                           \
  switch (J_) {                                          \
    case 1: { constexpr int J = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int J = 2; __VA_ARGS__; } break; \
    case 3: { constexpr int J = 3; __VA_ARGS__; } break; \
    case 4: { constexpr int J = 4; __VA_ARGS__; } break; \
    default: return GX_ERR_ARG;                          \
  }

template <int J, int R, int GT, int MODE>
static int launch_sinkhorn_pass(const float* s, long long n, int k, long long lds, float scale_log2, int first,
                                const float* u_in, const gx_ll_desc& u_ll, const float* r, const float* c, float cu,
                                float* partials, int grid, int reverse, __half* e16, long long lde, float* la1,
                                cudaStream_t st) {
  const int stage_bytes = k * 4 * R;
  // Each of the NG groups must own a fixed subset of the ring (stage index parity == iteration
  // parity): with a stage shared between groups, a group running ahead would observe the
  // mbarrier of a fill it does not own one phase early (parity aliasing).
  constexpr int NG = SK_THREADS / GT;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  stages -= stages % NG;
  GX_CHECK_ARG(stages >= 2 * NG || (NG == 1 && stages >= 2));
  static bool attr = false;
  if (!attr) {
    GX_CHECK_CUDA(cudaFuncSetAttribute(sinkhorn_pass_kernel<J, R, GT, MODE>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  sinkhorn_pass_kernel<J, R, GT, MODE><<<grid, SK_THREADS, stages * stage_bytes, st>>>(
      s, n, k, lds, scale_log2, first, u_in, u_ll, r, c, cu, partials, stages, reverse, e16, lde, la1);
  return GX_OK;
}

template <int J, int R, int GT>
static int launch_sinkhorn_pass16(const __half* e16, long long n, int k, long long lde, const float* u_in,
                                  const gx_ll_desc& u_ll, const float* r, const float* c, float cu, const float* la1,
                                  float* partials, int reverse, int* nparts_out, cudaStream_t st) {
  constexpr int NG = SK_THREADS / GT;
  const int stage_bytes = (int)lde * 2 * R;
  int stages = (200 * 1024) / stage_bytes;
  if (stages > 8) stages = 8;
  stages -= stages % NG;
  GX_CHECK_ARG(stages >= 2 * NG || (NG == 1 && stages >= 2));
  GX_CHECK_ARG((n + R - 1) / R < (1ll << 30));
  const int grid = sk_grid(n, R);
  if (nparts_out) *nparts_out = grid * NG;
  static bool attr = false;
  if (!attr) {
    GX_CHECK_CUDA(cudaFuncSetAttribute(sinkhorn_pass16_kernel<J, R, GT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       200 * 1024));
    attr = true;
  }
  sinkhorn_pass16_kernel<J, R, GT><<<grid, SK_THREADS, stages * stage_bytes, st>>>(
      e16, n, k, lde, u_in, u_ll, r, c, cu, la1, partials, stages, reverse);
  return GX_OK;
}

static int ll_desc_ok(const gx_ll_desc* d) {
  if (!d || d->world == 0) return 1;
  if (d->world < 0 || d->world > GX_MAX_PEERS || d->rank < 0 || d->rank >= d->world || d->seq == 0 ||
      d->block_words < 0 || (d->block_words & 3))
    return 0;
  for (int r = 0; r < d->world; ++r)
    if (!d->peers[r] || (reinterpret_cast<uintptr_t>(d->peers[r]) & 31)) return 0;
  return 1;
}
static gx_ll_desc ll_or_none(const gx_ll_desc* d) {
  gx_ll_desc z;
  memset(&z, 0, sizeof(z));
  return (d && d->world > 0) ? *d : z;
}

template <int MODE>
static int sinkhorn_pass_mode(const float* s, long long n, int k, long long lds, float inv_eps, int first,
                              const float* u_in, const gx_ll_desc& u_ll, const float* r, const float* c,
                              long long n_total, int reverse, float* partials, int* nparts_out, __half* e16,
                              long long lde, float* la1, cudaStream_t st) {
  const int grid = sk_grid(n, SK_ROWS);
  const float cu = 1.f / (float)(n_total > 0 ? n_total : n);
  const float sl = inv_eps * LOG2E;
  int rc;
  const int j256 = gx_cdiv(k, 1024);
#define GX_SK(J_, GT_)                                                                                               \
  rc = launch_sinkhorn_pass<J_, SK_ROWS, GT_, MODE>(s, n, k, lds, sl, first, u_in, u_ll, r, c, cu, partials, grid, \
                                                    reverse != 0, e16, lde, la1, st)
  if (j256 <= 5) {
    if (nparts_out) *nparts_out = grid * 2;
    switch (j256) {
      case 1: GX_SK(1, 256); break;
      case 2: GX_SK(2, 256); break;
      case 3: GX_SK(3, 256); break;
      case 4: GX_SK(4, 256); break;
      default: GX_SK(5, 256); break;
    }
  } else {
    if (nparts_out) *nparts_out = grid;
    if (k <= 6144) GX_SK(3, 512);
    else GX_SK(4, 512);
  }
#undef GX_SK
  return rc;
}

static int sinkhorn_pass_any(const float* s, long long n, int k, long long lds, float inv_eps, int first,
                             const float* u_in, const gx_ll_desc* u_ll_p, const float* r, const float* c,
                             long long n_total, int reverse, float* partials, int* nparts_out, void* e16,
                             long long lde, float* la1, int mode, void* stream) {
  GX_CHECK_ARG(partials && n > 0 && k >= 4 && k % 4 == 0 && k <= 8192);
  GX_CHECK_ARG(mode >= 0 && mode <= 2);
  if (mode != 2) {
    GX_CHECK_ARG(s && lds % 4 == 0 && lds >= k && (reinterpret_cast<uintptr_t>(s) & 15) == 0);
  }
  if (mode != 0) {   // the 16-bit plane: 16-byte aligned rows (bulk copies), written by a normalising (non-first) pass
    GX_CHECK_ARG(e16 && la1 && !first && lde % 8 == 0 && lde >= k);
    GX_CHECK_ARG((reinterpret_cast<uintptr_t>(e16) & 15) == 0 && (reinterpret_cast<uintptr_t>(la1) & 15) == 0);
  }
  GX_CHECK_ARG(ll_desc_ok(u_ll_p));
  const gx_ll_desc u_ll = ll_or_none(u_ll_p);
  GX_CHECK_ARG(first || u_in || u_ll.world > 0);
  GX_CHECK_ARG(!u_in || (reinterpret_cast<uintptr_t>(u_in) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  __half* e = reinterpret_cast<__half*>(e16);
  int rc;
  if (mode == 0)
    rc = sinkhorn_pass_mode<0>(s, n, k, lds, inv_eps, first, u_in, u_ll, r, c, n_total, reverse, partials, nparts_out,
                               e, lde, la1, st);
  else if (mode == 1)
    rc = sinkhorn_pass_mode<1>(s, n, k, lds, inv_eps, first, u_in, u_ll, r, c, n_total, reverse, partials, nparts_out,
                               e, lde, la1, st);
  else {
    const float cu = 1.f / (float)(n_total > 0 ? n_total : n);
    const int j256 = gx_cdiv(k, 1024);
    // 2 rows per ring stage (8 stages of 20 KB at K = 5000): 0.259 ms per pass at N = 160000 (6.18 TB/s); 3 / 4 rows per
    // stage measured 0.262 / 0.307 ms (register spills at 4) - profiles/r2g_sinkhorn_cache16.txt
#define GX_SK16R(J_, GT_)                                                                                          \
  rc = launch_sinkhorn_pass16<J_, SK_ROWS, GT_>(e, n, k, lde, u_in, u_ll, r, c, cu, la1, partials, reverse != 0, \
                                                nparts_out, st)
    switch (j256 <= 5 ? j256 : (k <= 6144 ? 6 : 7)) {
      case 1: GX_SK16R(1, 256); break;
      case 2: GX_SK16R(2, 256); break;
      case 3: GX_SK16R(3, 256); break;
      case 4: GX_SK16R(4, 256); break;
      case 5: GX_SK16R(5, 256); break;
      case 6: GX_SK16R(3, 512); break;
      default: GX_SK16R(4, 512); break;
    }
#undef GX_SK16R
  }
  if (rc != GX_OK) return rc;
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_sinkhorn_pass(const float* s, long long n, int k, long long lds, float inv_eps, int first,
                                const float* u_in, const gx_ll_desc* u_ll_p, const float* r, const float* c,
                                long long n_total, int reverse, float* partials, int* nparts_out, void* stream) {
  return sinkhorn_pass_any(s, n, k, lds, inv_eps, first, u_in, u_ll_p, r, c, n_total, reverse, partials, nparts_out,
                           nullptr, 0, nullptr, 0, stream);
}

extern "C" int gx_sinkhorn_pass_cached(const float* s, long long n, int k, long long lds, float inv_eps,
                                       const float* u_in, const gx_ll_desc* u_ll_p, const float* r, const float* c,
                                       long long n_total, int reverse, float* partials, int* nparts_out, void* e16,
                                       long long lde, float* la1, int write_cache, void* stream) {
  return sinkhorn_pass_any(s, n, k, lds, inv_eps, 0, u_in, u_ll_p, r, c, n_total, reverse, partials, nparts_out, e16,
                           lde, la1, write_cache ? 1 : 2, stream);
}

extern "C" int gx_sinkhorn_reduce(const float* partials, int nparts, int k, float* u, void* stream) {
  GX_CHECK_ARG(partials && u && nparts > 0 && k > 0);
  colsum_parts_kernel<<<gx_cdiv(k, 32), 256, 0, (cudaStream_t)stream>>>(partials, nparts, k, u);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_sinkhorn_reduce_send(const float* partials, int nparts, int k, const gx_ll_desc* ll,
                                       float* u_local, void* stream) {
  GX_CHECK_ARG(partials && nparts > 0 && k > 0 && k % 4 == 0 && ll && ll->world > 0 && ll_desc_ok(ll));
  colsum_send_kernel<<<gx_cdiv(k, 32), 256, 0, (cudaStream_t)stream>>>(partials, nparts, k, *ll, u_local);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_ll_recv_sum(const gx_ll_desc* ll, int k, float* u, void* stream) {
  GX_CHECK_ARG(u && k > 0 && k % 4 == 0 && ll && ll->world > 0 && ll_desc_ok(ll));
  ll_recv_sum_kernel<<<gx_cdiv(k, 128), 128, 0, (cudaStream_t)stream>>>(*ll, k, u);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_sinkhorn_log_a(const float* u, const gx_ll_desc* u_ll_p, const float* r, int k, float* log_a,
                                 void* stream) {
  GX_CHECK_ARG(log_a && k > 0 && ll_desc_ok(u_ll_p));
  const gx_ll_desc u_ll = ll_or_none(u_ll_p);
  GX_CHECK_ARG(u || u_ll.world > 0);
  GX_CHECK_ARG(u_ll.world == 0 || k % 4 == 0);
  log_a_kernel<<<gx_cdiv(k, 256), 256, 0, (cudaStream_t)stream>>>(u, u_ll, r, k, log_a);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_sinkhorn_q(const float* s, long long n, int k, long long lds, float inv_eps, const float* log_a,
                             float* q, void* stream) {
  GX_CHECK_ARG(s && log_a && q && n > 0 && k > 0 && k % 4 == 0 && lds % 4 == 0);
  const int jn = gx_cdiv(k, SK_THREADS * 4);
  const int grid = sk_grid(n, 1);
  GX_DISPATCH_J(jn, (sinkhorn_q_kernel<J><<<grid, SK_THREADS, 0, (cudaStream_t)stream>>>(s, n, k, lds, inv_eps,
                                                                                      log_a, q)));
  GX_LAUNCH_CHECK();
  return GX_OK;
}

template <int J, int GT, int NT>
static int launch_swav_loss(const float* s_s, const float* s_t, long long n, int k, long long lds, float inv_eps,
                            float inv_temp, const float* la_s, const float* la_t, float grad_scale, float* loss_parts,
                            float* db_parts, void* ds_s_hi, void* ds_s_lo, void* ds_t_hi, void* ds_t_lo,
                            long long ldd, float* fs, float* ft, int grid, cudaStream_t st) {
  const int stage_bytes = k * 4 * 2;
  constexpr int NG = NT / GT;
  int stages = (200 * 1024 - stage_bytes) / stage_bytes;   // one stage worth of smem holds log2(a) of both views
  if (stages > 8) stages = 8;
  stages -= stages % NG;                                   // fixed stage ownership per group (see sinkhorn pass)
  GX_CHECK_ARG(stages >= NG);   // a stage is refilled once its group holds the row in registers
  // exp(S/eps + log a) = exp(S/T)^rho * a when rho = T/eps is 1 or 2: half the exponentials
  const float rho = inv_eps / inv_temp;
  const int irho = (fabsf(rho - 2.f) < 1e-5f) ? 2 : ((fabsf(rho - 1.f) < 1e-5f) ? 1 : 0);
  static const bool no_pow = getenv("GX_LOSS_GENERIC") != nullptr;   // profiling / tests: force the general kernel
#define GX_LOSS_ARGS                                                                                           \
  loss_parts, db_parts, reinterpret_cast<__nv_bfloat16*>(ds_s_hi), reinterpret_cast<__nv_bfloat16*>(ds_s_lo),    \
      reinterpret_cast<__nv_bfloat16*>(ds_t_hi), reinterpret_cast<__nv_bfloat16*>(ds_t_lo), ldd, fs, ft, stages
  const size_t smem = (size_t)(stages + 1) * stage_bytes;
  if (irho == 2 && !no_pow) {
    static bool attr2 = false;
    if (!attr2) {
      GX_CHECK_CUDA(cudaFuncSetAttribute(swav_loss_pow_kernel<J, GT, NT, 2>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr2 = true;
    }
    swav_loss_pow_kernel<J, GT, NT, 2><<<grid, NT, smem, st>>>(s_s, s_t, n, k, lds, inv_temp, la_s, la_t, grad_scale,
                                                              GX_LOSS_ARGS);
  } else if (irho == 1 && !no_pow) {
    static bool attr1 = false;
    if (!attr1) {
      GX_CHECK_CUDA(cudaFuncSetAttribute(swav_loss_pow_kernel<J, GT, NT, 1>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr1 = true;
    }
    swav_loss_pow_kernel<J, GT, NT, 1><<<grid, NT, smem, st>>>(s_s, s_t, n, k, lds, inv_temp, la_s, la_t, grad_scale,
                                                              GX_LOSS_ARGS);
  } else {
    static bool attr = false;
    if (!attr) {
      GX_CHECK_CUDA(cudaFuncSetAttribute(swav_loss_kernel<J, GT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         200 * 1024));
      attr = true;
    }
    swav_loss_kernel<J, GT, NT><<<grid, NT, smem, st>>>(s_s, s_t, n, k, lds, inv_eps, inv_temp, la_s, la_t,
                                                       grad_scale, GX_LOSS_ARGS);
  }
#undef GX_LOSS_ARGS
  return GX_OK;
}

extern "C" int gx_swav_loss(const float* s_s, const float* s_t, long long n, int k, long long lds, float inv_eps,
                            float inv_temp, const float* log_a_s, const float* log_a_t, float grad_scale,
                            float* loss_parts, float* db_parts, int* nparts_out, void* ds_s_hi, void* ds_s_lo,
                            void* ds_t_hi, void* ds_t_lo, long long ldd, float* ds_s_f32, float* ds_t_f32,
                            void* stream) {
  GX_CHECK_ARG(s_s && s_t && log_a_s && log_a_t && loss_parts && ds_s_hi && ds_t_hi);
  GX_CHECK_ARG(n > 0 && k >= 4 && k % 4 == 0 && k <= 8192 && lds % 4 == 0 && ldd % 4 == 0 && ldd >= k);
  GX_CHECK_ARG((reinterpret_cast<uintptr_t>(s_s) & 15) == 0 && (reinterpret_cast<uintptr_t>(s_t) & 15) == 0);
  const int grid = sk_grid(n, 1);
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  // Two independent groups of NWG warps with J <= 8 sweeps of 4 columns per thread, so that the
  // four exponential arrays + bias-gradient accumulators of a row pair stay in registers
  // (K = 5000: 5 warps per group, 8 sweeps).
  const int nwg = gx_cdiv(k, 1024);
#define GX_LS(J_, GT_, NT_)                                                                                     \
  rc = launch_swav_loss<J_, GT_, NT_>(s_s, s_t, n, k, lds, inv_eps, inv_temp, log_a_s, log_a_t, grad_scale,     \
                                      loss_parts, db_parts, ds_s_hi, ds_s_lo, ds_t_hi, ds_t_lo, ldd, ds_s_f32, \
                                      ds_t_f32, grid, st)
  if (nwg == 1) {
    if (nparts_out) *nparts_out = grid * 2;
    switch (gx_cdiv(k, 128)) {
      case 1: GX_LS(1, 32, 64); break;
      case 2: GX_LS(2, 32, 64); break;
      case 3: GX_LS(3, 32, 64); break;
      case 4: GX_LS(4, 32, 64); break;
      case 5: GX_LS(5, 32, 64); break;
      case 6: GX_LS(6, 32, 64); break;
      case 7: GX_LS(7, 32, 64); break;
      default: GX_LS(8, 32, 64); break;
    }
  } else if (k > 5120) {
    // two groups of 256 threads (register-limited: spills, still correct)
    if (nparts_out) *nparts_out = grid * 2;
    switch (nwg) {
      case 6: GX_LS(6, 256, 512); break;
      case 7: GX_LS(7, 256, 512); break;
      default: GX_LS(8, 256, 512); break;
    }
  } else {
    // two independent groups of 128 threads
    if (nparts_out) *nparts_out = grid * 2;
    switch (gx_cdiv(k, 512)) {
      case 3: GX_LS(3, 128, 256); break;
      case 4: GX_LS(4, 128, 256); break;
      case 5: GX_LS(5, 128, 256); break;
      case 6: GX_LS(6, 128, 256); break;
      case 7: GX_LS(7, 128, 256); break;
      case 8: GX_LS(8, 128, 256); break;
      case 9: GX_LS(9, 128, 256); break;
      default: GX_LS(10, 128, 256); break;
    }
  }
#undef GX_LS
  if (rc != GX_OK) return rc;
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_larc_scratch_floats(void) { return 2 * 4 * gx_sm_count(); }

extern "C" int gx_larc_sgd(float* p, const float* g, float* buf, long long n, float lr, float momentum, float trust,
                           float weight_decay, float eps, int first_step, float* norms, void* stream) {
  GX_CHECK_ARG(p && g && buf && norms && n > 0);
  cudaStream_t st = (cudaStream_t)stream;
  int grid = gx_cdiv(n, 256 * 8);
  const int cap = gx_larc_scratch_floats() / 2;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  sq_norms_kernel<<<grid, 256, 0, st>>>(p, g, n, norms);
  GX_LAUNCH_CHECK();
  larc_sgd_kernel<<<grid, 256, 0, st>>>(p, g, buf, n, lr, momentum, trust, weight_decay, eps, first_step, norms, grid);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_argmax_rows(const float* x, long long n, int c, long long ldx, long long* labels, void* stream) {
  GX_CHECK_ARG(x && labels && n > 0 && c > 0);
  argmax_rows_kernel<<<gx_cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(x, n, c, ldx, labels);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_kmeans_assign(const float* x1, int c1, const float* x2, int c2, long long n, const float* centers,
                                int k, int* labels, float* dist, void* stream) {
  GX_CHECK_ARG(x1 && centers && (labels || dist) && n > 0 && c1 > 0 && k > 0 && c2 >= 0);
  GX_CHECK_ARG((x2 != nullptr) == (c2 > 0));
  kmeans_assign_kernel<<<gx_cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(x1, c1, x2, c2, n, centers, k, labels,
                                                                        dist);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_argmin_affine(const float* s, long long n, int k, long long lds, const float* bias, float scale,
                                int* labels, void* stream) {
  GX_CHECK_ARG(s && bias && labels && n > 0 && k > 0 && lds >= k);
  argmin_affine_kernel<<<gx_cdiv(n, 8), 256, 0, (cudaStream_t)stream>>>(s, n, k, lds, bias, scale, labels);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_onehot_nearest(const int* labels, int b, int h, int w, int k, int out_h, int out_w, float* out,
                                 long long out_batch_stride, float on_value, float off_value, void* stream) {
  GX_CHECK_ARG(labels && out && b > 0 && h > 0 && w > 0 && k > 0 && out_h > 0 && out_w > 0);
  if (out_batch_stride <= 0) out_batch_stride = (long long)k * out_h * out_w;
  GX_CHECK_ARG(out_batch_stride >= (long long)k * out_h * out_w);
  const bool vec = (out_w % 4 == 0) && (out_batch_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
  const long long total = (long long)b * k * out_h * (vec ? out_w / 4 : out_w);
  int grid = gx_cdiv(total, 256);
  const int cap = gx_sm_count() * 16;
  if (grid > cap) grid = cap;
  onehot_nearest_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(labels, b, h, w, k, out_h, out_w, out_batch_stride,
                                                                vec ? 4 : 1, on_value, off_value, out);
  GX_LAUNCH_CHECK();
  return GX_OK;
}
