// FIR resampling (upfirdn2d), fused bias+activation, and the hot-path fusion of both:
// blur + noise + bias + leaky-relu*sqrt2 (+ next conv's modulate/split) on NHWC maps.
#include "gx_common.cuh"

namespace {

// --------------------------------------------------------------------------
// Generic upfirdn2d, layout [major, h, w, minor] (ref: upfirdn2d_kernel.cu:52-215).
// out[oy,ox] = sum_{ky,kx} k[kh-1-ky][kw-1-kx] * u[oy*dy+ky][ox*dx+kx], where u is the
// zero-inserted, padded/cropped input.  Only taps that land on real samples are visited.
// --------------------------------------------------------------------------
struct UpfirParams {
  int major, in_h, in_w, minor, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_y0, out_h, out_w;
};

__device__ __forceinline__ int floor_div(int a, int b) {
  int q = a / b;
  return (q * b > a) ? q - 1 : q;
}

__global__ void upfirdn2d_kernel(const float* __restrict__ in, const float* __restrict__ kern,
                                 float* __restrict__ out, const UpfirParams p, long long total) {
  extern __shared__ float sk[];  // flipped taps
  for (int i = threadIdx.x; i < p.kh * p.kw; i += blockDim.x) {
    const int ky = i / p.kw, kx = i - ky * p.kw;
    sk[i] = kern[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)];
  }
  __syncthreads();
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx;
    const int mi = (int)(r % p.minor); r /= p.minor;
    const int ox = (int)(r % p.out_w); r /= p.out_w;
    const int oy = (int)(r % p.out_h); r /= p.out_h;
    const int mj = (int)r;
    // u index of tap ky is oy*dy+ky; it maps to input row (oy*dy+ky-pad_y0)/up_y when divisible
    const int base_y = oy * p.down_y - p.pad_y0;
    const int base_x = ox * p.down_x - p.pad_x0;
    int ky0 = ((-base_y) % p.up_y + p.up_y) % p.up_y;  // smallest ky >= 0 with (base_y+ky) % up_y == 0
    int kx0 = ((-base_x) % p.up_x + p.up_x) % p.up_x;
    float acc = 0.f;
    const float* src = in + (long long)mj * p.in_h * p.in_w * p.minor + mi;
    for (int ky = ky0; ky < p.kh; ky += p.up_y) {
      const int iy = floor_div(base_y + ky, p.up_y);
      if (iy < 0 || iy >= p.in_h) continue;
      for (int kx = kx0; kx < p.kw; kx += p.up_x) {
        const int ix = floor_div(base_x + kx, p.up_x);
        if (ix < 0 || ix >= p.in_w) continue;
        acc = fmaf(__ldg(src + ((long long)iy * p.in_w + ix) * p.minor), sk[ky * p.kw + kx], acc);
      }
    }
    out[idx] = acc;
  }
}

// --------------------------------------------------------------------------
// Planar fast path (minor == 1: the NCHW tensors the Python-level upfirdn2d / Blur / Upsample / Downsample of
// the reference feed the op, ref upfirdn2d.py:146-162): filters up to 4x4, up/down in {1, 2} with up*down <= 2.
// A CTA owns a 32 x 64 output tile of one plane: the input rows it needs are staged once in shared memory with
// row-contiguous (coalesced) loads, zero-filled outside the image, and every thread produces 8 consecutive rows of
// one output column from the tile, so the taps of neighbouring rows are common sub-expressions and the inner loop
// has neither bounds checks nor index divisions.  (The per-element kernel above issues ~16 predicated scalar
// global loads per output: 0.2 TB/s on the 256^2 blur against 1.7 TB/s for the reference's tiled kernel.)
// --------------------------------------------------------------------------
constexpr int UT_OW = 64, UT_OH = 32, UT_K = 4, UT_ROWS = 8;

__device__ __forceinline__ void cp_async_f32(float* smem_dst, const float* gsrc, bool valid) {
  // 4-byte asynchronous copy; src-size 0 zero-fills the destination (the zero border of the tile)
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}

// NBUF = 2: the CTA walks over several planes and the tile of the next plane is in flight (cp.async) while the
// current one is filtered; NBUF = 1 where two tiles do not fit the static shared-memory budget (DOWN = 2).
template <int UP, int DOWN, int NBUF>
__global__ void __launch_bounds__(256)
upfirdn2d_planar_kernel(const float* __restrict__ in, const float* __restrict__ kern, float* __restrict__ out,
                        const UpfirParams p) {
  constexpr int IN_H = ((UT_OH - 1) * DOWN + UT_K - 1) / UP + 2;
  constexpr int IN_W = ((UT_OW - 1) * DOWN + UT_K - 1) / UP + 2;
  constexpr int TAPS = (UT_K + UP - 1) / UP;            // taps per axis that land on real samples
  __shared__ float tile[NBUF][IN_H][IN_W + 1];
  __shared__ float fk[UT_K + UP][UT_K + UP];            // flipped taps, zero beyond kh x kw
  const int tid = threadIdx.x;
  for (int i = tid; i < (UT_K + UP) * (UT_K + UP); i += 256) {
    const int ky = i / (UT_K + UP), kx = i % (UT_K + UP);
    fk[ky][kx] = (ky < p.kh && kx < p.kw) ? kern[(p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx)] : 0.f;
  }
  const int ox0 = blockIdx.x * UT_OW, oy0 = blockIdx.y * UT_OH;
  const int uy0 = oy0 * DOWN - p.pad_y0, ux0 = ox0 * DOWN - p.pad_x0;   // zero-inserted coordinates of tap 0
  const int iy0 = floor_div(uy0, UP), ix0 = floor_div(ux0, UP);
  const int tx = tid % UT_OW, ty = tid / UT_OW;
  const int ox = ox0 + tx;
  const int ux = ux0 + tx * DOWN;
  const int kx0 = ((-ux) % UP + UP) % UP;
  const int lx0 = (ux + kx0) / UP - ix0;                // exact division: ux + kx0 is a multiple of UP
  auto fetch = [&](int mj, int buf) {
    const float* src = in + (long long)mj * p.in_h * p.in_w;
    for (int i = tid; i < IN_H * IN_W; i += 256) {
      const int r = i / IN_W, c = i - r * IN_W;
      const int iy = iy0 + r, ix = ix0 + c;
      const bool ok = iy >= 0 && iy < p.in_h && ix >= 0 && ix < p.in_w;
      cp_async_f32(&tile[buf][r][c], ok ? src + (long long)iy * p.in_w + ix : src, ok);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  int it = 0;
  if (blockIdx.z < (unsigned)p.major) fetch(blockIdx.z, 0);
  for (int mj = blockIdx.z; mj < p.major; mj += gridDim.z, ++it) {
    const int buf = NBUF == 2 ? (it & 1) : 0;
    if (NBUF == 2 && mj + (int)gridDim.z < p.major) {
      fetch(mj + gridDim.z, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    if (ox < p.out_w) {
      float* dst = out + (long long)mj * p.out_h * p.out_w + ox;
      if constexpr (UP == 1 && DOWN == 1) {
        // sliding window: each of the 8 + 3 tile rows of this column is read once (4 LDS) and feeds the up to four
        // output rows it overlaps - 5.5 instead of 16 shared-memory loads per output
        float kr[UT_K][UT_K], acc[UT_ROWS];
#pragma unroll
        for (int a = 0; a < UT_K; ++a)
#pragma unroll
          for (int b = 0; b < UT_K; ++b) kr[a][b] = fk[a][b];
#pragma unroll
        for (int j = 0; j < UT_ROWS; ++j) acc[j] = 0.f;
        const int ly = ty * UT_ROWS + (uy0 - iy0);            // uy0 - iy0 == 0 for UP == 1
#pragma unroll
        for (int r = 0; r < UT_ROWS + UT_K - 1; ++r) {
          const float v0 = tile[buf][ly + r][lx0], v1 = tile[buf][ly + r][lx0 + 1], v2 = tile[buf][ly + r][lx0 + 2],
                      v3 = tile[buf][ly + r][lx0 + 3];
#pragma unroll
          for (int j = 0; j < UT_ROWS; ++j) {
            const int ky = r - j;
            if (ky >= 0 && ky < UT_K)
              acc[j] = fmaf(v0, kr[ky][0], fmaf(v1, kr[ky][1], fmaf(v2, kr[ky][2], fmaf(v3, kr[ky][3], acc[j]))));
          }
        }
#pragma unroll
        for (int j = 0; j < UT_ROWS; ++j) {
          const int oy = oy0 + ty * UT_ROWS + j;
          if (oy < p.out_h) dst[(long long)oy * p.out_w] = acc[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < UT_ROWS; ++j) {
          const int oyl = ty * UT_ROWS + j;
          const int uy = uy0 + oyl * DOWN;
          const int ky0 = ((-uy) % UP + UP) % UP;
          const int ly0 = (uy + ky0) / UP - iy0;
          float acc = 0.f;
#pragma unroll
          for (int a = 0; a < TAPS; ++a)
#pragma unroll
            for (int b = 0; b < TAPS; ++b)
              acc = fmaf(tile[buf][ly0 + a][lx0 + b], fk[ky0 + a * UP][kx0 + b * UP], acc);
          if (oy0 + oyl < p.out_h) dst[(long long)(oy0 + oyl) * p.out_w] = acc;
        }
      }
    }
    __syncthreads();                       // the tile may be overwritten by the next fetch
    if (NBUF == 1 && mj + (int)gridDim.z < p.major) fetch(mj + gridDim.z, 0);
  }
}

// --------------------------------------------------------------------------
// fused_bias_act (ref: fused_bias_act_kernel.cu:18-85)
// --------------------------------------------------------------------------
__device__ __forceinline__ float bias_act_one(float x, float b, float ref, int mode, float alpha, float scale) {
  x += b;
  float y;
  switch (mode) {
    default:
    case 10: case 11: y = x; break;
    case 12: y = 0.f; break;
    case 30: y = (x > 0.f) ? x : x * alpha; break;
    case 31: y = (ref > 0.f) ? x : x * alpha; break;
    case 32: y = 0.f; break;
  }
  return y * scale;
}

__global__ void fused_bias_act_kernel(const float* __restrict__ x, const float* __restrict__ b,
                                      const float* __restrict__ ref, float* __restrict__ out, long long n, int step_b,
                                      int size_b, int mode, float alpha, float scale, int vec) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  if (vec) {
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      const float4 v = gx_ldg_stream(reinterpret_cast<const float4*>(x) + i);
      float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ref) rv = gx_ldg_stream(reinterpret_cast<const float4*>(ref) + i);
      float bb = 0.f;
      if (b) bb = __ldg(b + ((i * 4) / step_b) % size_b);  // step_b % 4 == 0: one bias per vector
      float4 o;
      o.x = bias_act_one(v.x, bb, rv.x, mode, alpha, scale);
      o.y = bias_act_one(v.y, bb, rv.y, mode, alpha, scale);
      o.z = bias_act_one(v.z, bb, rv.z, mode, alpha, scale);
      o.w = bias_act_one(v.w, bb, rv.w, mode, alpha, scale);
      gx_stg_stream(reinterpret_cast<float4*>(out) + i, o);
    }
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float bb = b ? __ldg(b + (i / step_b) % size_b) : 0.f;
      const float rv = ref ? ref[i] : 0.f;
      out[i] = bias_act_one(x[i], bb, rv, mode, alpha, scale);
    }
  }
}

// --------------------------------------------------------------------------
// Hot path: blur (4x4 FIR, up=down=1) of the (2H+1)^2 transposed-conv output, fused with
// noise + bias + lrelu*sqrt2 and the next layer's modulate+split.  NHWC, 4 channels per
// thread (128-bit accesses), a sliding 4-row window of input vectors held in registers so
// each input vector is fetched once per output column.
// --------------------------------------------------------------------------
constexpr int BLUR_STRIP = 16;
constexpr int BLUR_PF = 3;      // L2 prefetch distance of the separable blur kernel, in rows

template <int KH, int KW>
__global__ void __launch_bounds__(256)
blur_fused_kernel(const float* __restrict__ in, const float* __restrict__ fir, int pad0, const float* __restrict__ noise,
                  long long noise_bstride, const float* __restrict__ noise_strength, const float* __restrict__ bias,
                  int act, float* __restrict__ out, const float* __restrict__ next_style,
                  __nv_bfloat16* __restrict__ next_hi, __nv_bfloat16* __restrict__ next_lo, int next_ld, int batch,
                  int hi, int wi, int ho, int wo, int c) {
  __shared__ float sk[KH * KW];
  if (threadIdx.x < KH * KW) {
    const int ky = threadIdx.x / KW, kx = threadIdx.x % KW;
    sk[threadIdx.x] = fir[(KH - 1 - ky) * KW + (KW - 1 - kx)];
  }
  __syncthreads();
  const int cq = c >> 2;                       // channel quads
  const int cq_blk = cq < 256 ? cq : 256;      // quads handled per block row
  const int xs_per_blk = 256 / cq_blk;
  const int qi = threadIdx.x % cq_blk;
  const int xi = threadIdx.x / cq_blk;
  const int cq_groups = (cq + cq_blk - 1) / cq_blk;
  const int cq0 = (blockIdx.z % cq_groups) * cq_blk;
  const int b = blockIdx.z / cq_groups;
  const int ox = blockIdx.x * xs_per_blk + xi;
  const int q = cq0 + qi;
  if (ox >= wo || q >= cq || b >= batch) return;
  const int oy0 = blockIdx.y * BLUR_STRIP;
  const int oy1 = min(ho, oy0 + BLUR_STRIP);
  float kreg[KH * KW];
#pragma unroll
  for (int i = 0; i < KH * KW; ++i) kreg[i] = sk[i];
  const float4* src = reinterpret_cast<const float4*>(in) + (long long)b * hi * wi * cq + q;
  float4 win[KH][KW];
  auto load_row = [&](int iy, float4 (&dst)[KW]) {
#pragma unroll
    for (int kx = 0; kx < KW; ++kx) {
      const int ix = ox + kx - pad0;
      if (iy >= 0 && iy < hi && ix >= 0 && ix < wi)
        dst[kx] = __ldg(src + ((long long)iy * wi + ix) * cq);
      else
        dst[kx] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
#pragma unroll
  for (int ky = 0; ky < KH - 1; ++ky) load_row(oy0 + ky - pad0, win[ky + 1]);
  const float nstr = noise ? __ldg(noise_strength) : 0.f;
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias) bs = __ldg(reinterpret_cast<const float4*>(bias) + q);
  float4 st = make_float4(1.f, 1.f, 1.f, 1.f);
  if (next_style) st = __ldg(reinterpret_cast<const float4*>(next_style) + (long long)b * cq + q);
  for (int oy = oy0; oy < oy1; ++oy) {
#pragma unroll
    for (int ky = 0; ky < KH - 1; ++ky)
#pragma unroll
      for (int kx = 0; kx < KW; ++kx) win[ky][kx] = win[ky + 1][kx];
    load_row(oy + KH - 1 - pad0, win[KH - 1]);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int ky = 0; ky < KH; ++ky)
#pragma unroll
      for (int kx = 0; kx < KW; ++kx) {
        const float kv = kreg[ky * KW + kx];
        acc.x = fmaf(win[ky][kx].x, kv, acc.x);
        acc.y = fmaf(win[ky][kx].y, kv, acc.y);
        acc.z = fmaf(win[ky][kx].z, kv, acc.z);
        acc.w = fmaf(win[ky][kx].w, kv, acc.w);
      }
    if (noise) {
      const float nz = nstr * __ldg(noise + (long long)b * noise_bstride + (long long)oy * wo + ox);
      acc.x += nz; acc.y += nz; acc.z += nz; acc.w += nz;
    }
    acc.x += bs.x; acc.y += bs.y; acc.z += bs.z; acc.w += bs.w;
    if (act) {
      const float s2 = 1.41421356237309515f;
      acc.x = (acc.x > 0.f ? acc.x : acc.x * 0.2f) * s2;
      acc.y = (acc.y > 0.f ? acc.y : acc.y * 0.2f) * s2;
      acc.z = (acc.z > 0.f ? acc.z : acc.z * 0.2f) * s2;
      acc.w = (acc.w > 0.f ? acc.w : acc.w * 0.2f) * s2;
    }
    const long long o = (((long long)b * ho + oy) * wo + ox) * cq + q;
    reinterpret_cast<float4*>(out)[o] = acc;
    if (next_hi) {
      uint2 h, l;
      gx_split4(make_float4(acc.x * st.x, acc.y * st.y, acc.z * st.z, acc.w * st.w), h, l);
      const long long on = (((long long)b * ho + oy) * wo + ox) * (next_ld >> 2) + q;
      reinterpret_cast<uint2*>(next_hi)[on] = h;
      if (next_lo) reinterpret_cast<uint2*>(next_lo)[on] = l;
    }
  }
}

// Separable variant (the shipped blur is outer([1,3,3,1]) * const): out = sum_ky gy[ky] * (sum_kx gx[kx] * in).  The kernel
// above is bound by instruction issue (64 FFMA per 4 channels + a 127-register window that limits occupancy), not
// by HBM; here a row costs 8 + 8 packed FFMA2 per 4 channels, the register window holds 4 filtered vectors instead
// of 16 raw ones, and the next row's raw vectors are requested before the current row's epilogue.
// Threads: channel quad fastest (32 quads or all of them), then x, so the 4-tap horizontal re-reads hit L1.
template <int KT>
__global__ void __launch_bounds__(256, 3)
blur_sep_kernel(const float* __restrict__ in, const float* __restrict__ fir_x, const float* __restrict__ fir_y, int pad0,
                const float* __restrict__ noise, long long noise_bstride, const float* __restrict__ noise_strength,
                const float* __restrict__ bias, int act, float* __restrict__ out, const float* __restrict__ next_style,
                __nv_bfloat16* __restrict__ next_hi, __nv_bfloat16* __restrict__ next_lo, int next_ld, int hi, int wi,
                int ho, int wo, int c, int cqb) {
  static_assert(BLUR_STRIP % KT == 0, "the row loop is unrolled by the tap count (static window rotation)");
  const int cq = c >> 2;
  const int cq_groups = cq / cqb;
  const int g = blockIdx.z % cq_groups, b = blockIdx.z / cq_groups;
  const int qi = threadIdx.x % cqb, xi = threadIdx.x / cqb;
  const int q = g * cqb + qi;
  const int ox = blockIdx.x * (256 / cqb) + xi;
  if (ox >= wo) return;
  const int oy0 = blockIdx.y * BLUR_STRIP;
  const int oy1 = min(ho, oy0 + BLUR_STRIP);
  // flipped taps (correlation with the flipped filter, ref upfirdn2d_native).  A column outside the input gets a
  // zero tap and a clamped address, so the row loads need no per-column predicate; addresses are four pointers
  // advanced by one input row per output row (the previous version recomputed 64-bit addresses: ~80 of its ~210
  // instructions per row, which bound the kernel by instruction issue at half of the HBM rate).
  float2 gx2[KT], gy2[KT];
  const float4* cp[KT];
  const long long in_row = (long long)wi * cq;
  const float4* src = reinterpret_cast<const float4*>(in) + (long long)b * hi * in_row + q;
  int iy = oy0 - pad0;                          // input row of the next load
#pragma unroll
  for (int i = 0; i < KT; ++i) {
    const int ix = ox + i - pad0;
    const bool ok = ix >= 0 && ix < wi;
    const float a = ok ? __ldg(fir_x + KT - 1 - i) : 0.f, bq = __ldg(fir_y + KT - 1 - i);
    gx2[i] = make_float2(a, a);
    gy2[i] = make_float2(bq, bq);
    cp[i] = src + (long long)iy * in_row + (long long)min(max(ix, 0), wi - 1) * cq;
  }
  float4 raw[KT];
  bool raw_ok = false;                          // block-uniform: the row in `raw` lies inside the input
  auto load_raw = [&]() {                       // row iy, then advance
    raw_ok = iy >= 0 && iy < hi;
    if (raw_ok) {
#pragma unroll
      for (int kx = 0; kx < KT; ++kx) raw[kx] = __ldg(cp[kx]);
    }
    // the loop holds one row of loads in flight per thread (one DRAM latency per output row): pull the row
    // BLUR_PF rows ahead into L2 meanwhile, no registers needed
    if (iy + BLUR_PF < hi) asm volatile("prefetch.global.L2 [%0];" ::"l"(cp[KT / 2] + BLUR_PF * in_row));
#pragma unroll
    for (int kx = 0; kx < KT; ++kx) cp[kx] += in_row;
    ++iy;
  };
  auto hfilter = [&]() {
    float2 a = make_float2(0.f, 0.f), bq = make_float2(0.f, 0.f);
    if (raw_ok) {
#pragma unroll
      for (int kx = 0; kx < KT; ++kx) {
        a = fma2(make_float2(raw[kx].x, raw[kx].y), gx2[kx], a);
        bq = fma2(make_float2(raw[kx].z, raw[kx].w), gx2[kx], bq);
      }
    }
    return make_float4(a.x, a.y, bq.x, bq.y);
  };
  float4 hw[KT];                                // hw[(r + ky) % KT]: filtered input row of tap ky of output row r
#pragma unroll
  for (int ky = 0; ky < KT - 1; ++ky) {
    load_raw();
    hw[ky] = hfilter();
  }
  load_raw();
  const float nstr = noise ? __ldg(noise_strength) : 0.f;
  const float2 ns2 = make_float2(nstr, nstr);
  float2 b01 = make_float2(0.f, 0.f), b23 = make_float2(0.f, 0.f);
  if (bias) {
    const float4 bs = __ldg(reinterpret_cast<const float4*>(bias) + q);
    b01 = make_float2(bs.x, bs.y);
    b23 = make_float2(bs.z, bs.w);
  }
  float2 s01 = make_float2(1.f, 1.f), s23 = make_float2(1.f, 1.f);
  if (next_style) {
    const float4 st = __ldg(reinterpret_cast<const float4*>(next_style) + (long long)b * cq + q);
    s01 = make_float2(st.x, st.y);
    s23 = make_float2(st.z, st.w);
  }
  const long long opix = ((long long)b * ho + oy0) * wo + ox;
  float4* op = reinterpret_cast<float4*>(out) + opix * cq + q;
  const long long out_row = (long long)wo * cq;
  const int nq = next_ld >> 2;
  const bool has_noise = noise != nullptr, has_hi = next_hi != nullptr, has_lo = next_lo != nullptr;   // uniform
  uint2* hp = reinterpret_cast<uint2*>(next_hi) + opix * nq + q;
  uint2* lp = reinterpret_cast<uint2*>(next_lo) + opix * nq + q;
  const long long pl_row = (long long)wo * nq;
  const float* np = noise + (long long)b * noise_bstride + (long long)oy0 * wo + ox;
  float nz_next = has_noise ? __ldg(np) : 0.f;
  const float2 k02 = make_float2(0.2f, 0.2f), ks2 = make_float2(1.41421356237309515f, 1.41421356237309515f);
  for (int oyb = oy0; oyb < oy1; oyb += KT) {
#pragma unroll
    for (int r = 0; r < KT; ++r) {
      if (oyb + r >= oy1) break;
      hw[(r + KT - 1) % KT] = hfilter();
      const float nz = nz_next;
      if (oyb + r + 1 < oy1) {                  // next row (and its noise value): in flight during this row's epilogue
        load_raw();
        if (has_noise) {
          np += wo;
          nz_next = __ldg(np);
        }
      }
      float2 a01 = make_float2(0.f, 0.f), a23 = make_float2(0.f, 0.f);
#pragma unroll
      for (int ky = 0; ky < KT; ++ky) {
        const float4 h = hw[(r + ky) % KT];
        a01 = fma2(make_float2(h.x, h.y), gy2[ky], a01);
        a23 = fma2(make_float2(h.z, h.w), gy2[ky], a23);
      }
      // reference order: + strength * noise (NoiseInjection), + bias, lrelu(0.2), * sqrt 2 (FusedLeakyReLU)
      const float2 nz2 = make_float2(nz, nz);
      a01 = add2(fma2(nz2, ns2, a01), b01);
      a23 = add2(fma2(nz2, ns2, a23), b23);
      if (act) {
        const float2 t01 = mul2(a01, k02), t23 = mul2(a23, k02);     // slope < 1: lrelu(x) = max(x, 0.2 x)
        a01 = mul2(make_float2(fmaxf(a01.x, t01.x), fmaxf(a01.y, t01.y)), ks2);
        a23 = mul2(make_float2(fmaxf(a23.x, t23.x), fmaxf(a23.y, t23.y)), ks2);
      }
      gx_stg_stream(op, make_float4(a01.x, a01.y, a23.x, a23.y));
      op += out_row;
      if (has_hi) {
        const float2 m01 = mul2(a01, s01), m23 = mul2(a23, s23);
        uint2 h, l;
        gx_split2(m01.x, m01.y, h.x, l.x);
        gx_split2(m23.x, m23.y, h.y, l.y);
        *hp = h;
        hp += pl_row;
        if (has_lo) {
          *lp = l;
          lp += pl_row;
        }
      }
    }
  }
}

// ToRGB: warp per pixel, 3 dot products over C with the per-sample modulated 1x1 weights.
__global__ void torgb_kernel(const float* __restrict__ x, const float* __restrict__ w, float w_scale,
                             const float* __restrict__ s, const float* __restrict__ bias,
                             const float* __restrict__ skip, float* __restrict__ out, int batch, int hw, int c) {
  const int lane = threadIdx.x & 31;
  const long long pix = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= (long long)batch * hw) return;
  const int b = (int)(pix / hw);
  const int p = (int)(pix - (long long)b * hw);
  const float4* xr = reinterpret_cast<const float4*>(x + pix * c);
  const float4* sr = reinterpret_cast<const float4*>(s + (long long)b * c);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f;
  for (int i = lane; i < (c >> 2); i += 32) {
    const float4 xv = __ldg(xr + i);
    const float4 sv = __ldg(sr + i);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w) + i);
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + c) + i);
    const float4 w2 = __ldg(reinterpret_cast<const float4*>(w + 2 * c) + i);
    const float m0 = xv.x * sv.x, m1 = xv.y * sv.y, m2 = xv.z * sv.z, m3 = xv.w * sv.w;
    a0 += m0 * (w0.x * w_scale) + m1 * (w0.y * w_scale) + m2 * (w0.z * w_scale) + m3 * (w0.w * w_scale);
    a1 += m0 * (w1.x * w_scale) + m1 * (w1.y * w_scale) + m2 * (w1.z * w_scale) + m3 * (w1.w * w_scale);
    a2 += m0 * (w2.x * w_scale) + m1 * (w2.y * w_scale) + m2 * (w2.z * w_scale) + m3 * (w2.w * w_scale);
  }
  a0 = gx_warp_sum(a0); a1 = gx_warp_sum(a1); a2 = gx_warp_sum(a2);
  if (lane < 3) {
    float v = lane == 0 ? a0 : (lane == 1 ? a1 : a2);
    v += bias ? bias[lane] : 0.f;
    const long long o = ((long long)b * 3 + lane) * hw + p;
    if (skip) v += skip[o];
    out[o] = v;
  }
}

// --------------------------------------------------------------------------
// half / double instantiations of the two native ops (the reference dispatches them with
// AT_DISPATCH_FLOATING_TYPES_AND_HALF, upfirdn2d_kernel.cu:321, fused_bias_act_kernel.cu:127): same per-element
// kernels on the storage type T, arithmetic in A (float for half, double for double).
// --------------------------------------------------------------------------
template <typename T> struct Acc { using type = float; };
template <> struct Acc<double> { using type = double; };
template <typename T> __device__ __forceinline__ typename Acc<T>::type ldv(const T* p) { return (typename Acc<T>::type)(*p); }
template <> __device__ __forceinline__ float ldv<__half>(const __half* p) { return __half2float(*p); }
template <typename T, typename A> __device__ __forceinline__ T stv(A v) { return (T)v; }
template <> __device__ __forceinline__ __half stv<__half, float>(float v) { return __float2half_rn(v); }

template <typename T>
__global__ void upfirdn2d_typed_kernel(const T* __restrict__ in, const T* __restrict__ kern, T* __restrict__ out,
                                       const UpfirParams p, long long total) {
  using A = typename Acc<T>::type;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    long long r = idx;
    const int mi = (int)(r % p.minor); r /= p.minor;
    const int ox = (int)(r % p.out_w); r /= p.out_w;
    const int oy = (int)(r % p.out_h); r /= p.out_h;
    const int mj = (int)r;
    const int base_y = oy * p.down_y - p.pad_y0;
    const int base_x = ox * p.down_x - p.pad_x0;
    const int ky0 = ((-base_y) % p.up_y + p.up_y) % p.up_y;
    const int kx0 = ((-base_x) % p.up_x + p.up_x) % p.up_x;
    A acc = (A)0;
    const T* src = in + (long long)mj * p.in_h * p.in_w * p.minor + mi;
    for (int ky = ky0; ky < p.kh; ky += p.up_y) {
      const int iy = floor_div(base_y + ky, p.up_y);
      if (iy < 0 || iy >= p.in_h) continue;
      for (int kx = kx0; kx < p.kw; kx += p.up_x) {
        const int ix = floor_div(base_x + kx, p.up_x);
        if (ix < 0 || ix >= p.in_w) continue;
        acc += ldv<T>(src + ((long long)iy * p.in_w + ix) * p.minor) *
               ldv<T>(kern + (p.kh - 1 - ky) * p.kw + (p.kw - 1 - kx));
      }
    }
    out[idx] = stv<T, A>(acc);
  }
}

template <typename T>
__global__ void fused_bias_act_typed_kernel(const T* __restrict__ x, const T* __restrict__ b, const T* __restrict__ ref,
                                            T* __restrict__ out, long long n, int step_b, int size_b, int mode,
                                            float alpha, float scale) {
  using A = typename Acc<T>::type;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    A v = ldv<T>(x + i);
    if (b) v += ldv<T>(b + (i / step_b) % size_b);
    const A rv = ref ? ldv<T>(ref + i) : (A)0;
    A y;
    switch (mode) {
      default:
      case 10: case 11: y = v; break;
      case 12: case 32: y = (A)0; break;
      case 30: y = (v > (A)0) ? v : v * (A)alpha; break;
      case 31: y = (rv > (A)0) ? v : v * (A)alpha; break;
    }
    out[i] = stv<T, A>(y * (A)scale);
  }
}

}  // namespace

// dtype: 0 float32, 1 float16, 2 float64 (all tensors of the call, FIR taps included, in that type)
extern "C" int gx_upfirdn2d_t(int dtype, const void* input, const void* kernel, void* out, int major, int in_h,
                              int in_w, int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y,
                              int pad_x0, int pad_x1, int pad_y0, int pad_y1, void* stream) {
  if (dtype == 0)
    return gx_upfirdn2d((const float*)input, (const float*)kernel, (float*)out, major, in_h, in_w, minor, kh, kw, up_x,
                        up_y, down_x, down_y, pad_x0, pad_x1, pad_y0, pad_y1, stream);
  GX_CHECK_ARG((dtype == 1 || dtype == 2) && input && kernel && out);
  GX_CHECK_ARG(major > 0 && in_h > 0 && in_w > 0 && minor > 0 && kh > 0 && kw > 0);
  GX_CHECK_ARG(up_x > 0 && up_y > 0 && down_x > 0 && down_y > 0);
  UpfirParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) / down_y;
  p.out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) / down_x;
  GX_CHECK_ARG(p.out_h > 0 && p.out_w > 0);
  const long long total = (long long)major * p.out_h * p.out_w * minor;
  int grid = (int)((total + 255) / 256);
  const int cap = gx_sm_count() * 16;
  if (grid > cap) grid = cap;
  if (dtype == 1)
    upfirdn2d_typed_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>((const __half*)input, (const __half*)kernel,
                                                                           (__half*)out, p, total);
  else
    upfirdn2d_typed_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)input, (const double*)kernel,
                                                                           (double*)out, p, total);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_fused_bias_act_t(int dtype, const void* input, const void* bias, const void* refer, void* out,
                                   long long n, int step_b, int size_b, int act, int grad, float alpha, float scale,
                                   void* stream) {
  if (dtype == 0)
    return gx_fused_bias_act((const float*)input, (const float*)bias, (const float*)refer, (float*)out, n, step_b, size_b,
                             act, grad, alpha, scale, stream);
  GX_CHECK_ARG((dtype == 1 || dtype == 2) && input && out && n >= 0);
  if (n == 0) return GX_OK;
  const int mode = act * 10 + grad;
  GX_CHECK_ARG(mode == 10 || mode == 11 || mode == 12 || mode == 30 || mode == 31 || mode == 32);
  GX_CHECK_ARG(mode != 31 || refer != nullptr);
  if (bias) GX_CHECK_ARG(step_b > 0 && size_b > 0);
  else { step_b = 1; size_b = 1; }
  int grid = (int)((n + 255) / 256);
  const int cap = gx_sm_count() * 16;
  if (grid > cap) grid = cap;
  if (dtype == 1)
    fused_bias_act_typed_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const __half*)input, (const __half*)bias, (const __half*)refer, (__half*)out, n, step_b, size_b, mode, alpha, scale);
  else
    fused_bias_act_typed_kernel<double><<<grid, 256, 0, (cudaStream_t)stream>>>(
        (const double*)input, (const double*)bias, (const double*)refer, (double*)out, n, step_b, size_b, mode, alpha, scale);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

namespace {
}  // namespace

extern "C" int gx_upfirdn2d(const float* input, const float* kernel, float* out, int major, int in_h, int in_w,
                            int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0,
                            int pad_x1, int pad_y0, int pad_y1, void* stream) {
  GX_CHECK_ARG(input && kernel && out);
  GX_CHECK_ARG(major > 0 && in_h > 0 && in_w > 0 && minor > 0 && kh > 0 && kw > 0);
  GX_CHECK_ARG(up_x > 0 && up_y > 0 && down_x > 0 && down_y > 0);
  UpfirParams p;
  p.major = major; p.in_h = in_h; p.in_w = in_w; p.minor = minor; p.kh = kh; p.kw = kw;
  p.up_x = up_x; p.up_y = up_y; p.down_x = down_x; p.down_y = down_y; p.pad_x0 = pad_x0; p.pad_y0 = pad_y0;
  p.out_h = (in_h * up_y + pad_y0 + pad_y1 - kh + down_y) / down_y;
  p.out_w = (in_w * up_x + pad_x0 + pad_x1 - kw + down_x) / down_x;
  GX_CHECK_ARG(p.out_h > 0 && p.out_w > 0);
  GX_CHECK_ARG(kh * kw * 4 <= 48 * 1024);
  const long long total = (long long)major * p.out_h * p.out_w * minor;
  cudaStream_t st = (cudaStream_t)stream;
  if (minor == 1 && kh <= UT_K && kw <= UT_K && up_x == up_y && down_x == down_y && up_x <= 2 && down_x <= 2 &&
      up_x * down_x <= 2 && p.out_w >= 32 && p.out_h >= 8) {
    // enough CTAs for ~3 waves of 4 CTAs per SM; each CTA then walks over major / grid.z planes with the next
    // plane's tile in flight
    const int tiles = gx_cdiv(p.out_w, UT_OW) * gx_cdiv(p.out_h, UT_OH);
    int gz = gx_cdiv((long long)gx_sm_count() * 12, tiles);
    if (gz > major) gz = major;
    if (gz > 32768) gz = 32768;
    if (gz < 1) gz = 1;
    dim3 grid(gx_cdiv(p.out_w, UT_OW), gx_cdiv(p.out_h, UT_OH), gz);
    if (up_x == 2) upfirdn2d_planar_kernel<2, 1, 2><<<grid, 256, 0, st>>>(input, kernel, out, p);
    else if (down_x == 2) upfirdn2d_planar_kernel<1, 2, 1><<<grid, 256, 0, st>>>(input, kernel, out, p);
    else upfirdn2d_planar_kernel<1, 1, 2><<<grid, 256, 0, st>>>(input, kernel, out, p);
    GX_LAUNCH_CHECK();
    return GX_OK;
  }
  int grid = (int)((total + 255) / 256);
  const int cap = gx_sm_count() * 16;
  if (grid > cap) grid = cap;
  upfirdn2d_kernel<<<grid, 256, kh * kw * sizeof(float), st>>>(input, kernel, out, p, total);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_fused_bias_act(const float* input, const float* bias, const float* refer, float* out, long long n,
                                 int step_b, int size_b, int act, int grad, float alpha, float scale, void* stream) {
  GX_CHECK_ARG(input && out && n >= 0);
  if (n == 0) return GX_OK;
  const int mode = act * 10 + grad;
  GX_CHECK_ARG(mode == 10 || mode == 11 || mode == 12 || mode == 30 || mode == 31 || mode == 32);
  GX_CHECK_ARG(mode != 31 || refer != nullptr);
  if (bias) GX_CHECK_ARG(step_b > 0 && size_b > 0);
  else { step_b = 1; size_b = 1; }
  const bool aligned = ((reinterpret_cast<uintptr_t>(input) | reinterpret_cast<uintptr_t>(out) |
                         reinterpret_cast<uintptr_t>(refer)) & 15) == 0;
  const int vec = (aligned && (n % 4 == 0) && (step_b % 4 == 0)) ? 1 : 0;
  const long long work = vec ? n / 4 : n;
  int grid = (int)((work + 255) / 256);
  const int cap = gx_sm_count() * 16;
  if (grid > cap) grid = cap;
  fused_bias_act_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(input, bias, refer, out, n, step_b, size_b, mode, alpha,
                                                                scale, vec);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_blur_noise_bias_act(const float* in, const float* fir, int kh, int kw, int pad0, int pad1,
                                      const float* noise, long long noise_batch_stride, const float* noise_strength,
                                      const float* bias, int act, float* out, const float* next_style, void* next_hi,
                                      void* next_lo, int next_ld, int batch, int hi, int wi, int c, void* stream) {
  if (next_ld <= 0) next_ld = c;
  GX_CHECK_ARG(next_ld >= c && next_ld % 4 == 0);
  GX_CHECK_ARG(in && fir && out && batch > 0 && hi > 0 && wi > 0);
  GX_CHECK_ARG(c % 4 == 0);
  GX_CHECK_ARG(kh == 4 && kw == 4);  // StyleGAN2 blur_kernel=[1,3,3,1]; other sizes go through gx_upfirdn2d
  GX_CHECK_ARG(noise == nullptr || noise_strength != nullptr);
  GX_CHECK_ARG(next_style == nullptr || next_hi != nullptr);
  const int ho = hi + pad0 + pad1 - kh + 1, wo = wi + pad0 + pad1 - kw + 1;
  GX_CHECK_ARG(ho > 0 && wo > 0);
  const int cq = c / 4;
  const int cq_blk = cq < 256 ? cq : 256;
  const int xs = 256 / cq_blk;
  dim3 grid(gx_cdiv(wo, xs), gx_cdiv(ho, BLUR_STRIP), batch * gx_cdiv(cq, cq_blk));
  GX_CHECK_ARG(256 % cq_blk == 0);
  blur_fused_kernel<4, 4><<<grid, 256, 0, (cudaStream_t)stream>>>(
      in, fir, pad0, noise, noise_batch_stride, noise_strength, bias, act, out, next_style,
      reinterpret_cast<__nv_bfloat16*>(next_hi), reinterpret_cast<__nv_bfloat16*>(next_lo), next_ld, batch, hi, wi, ho,
      wo, c);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_blur_sep_noise_bias_act(const float* in, const float* fir_x, const float* fir_y, int ntaps, int pad0,
                                          int pad1, const float* noise, long long noise_batch_stride,
                                          const float* noise_strength, const float* bias, int act, float* out,
                                          const float* next_style, void* next_hi, void* next_lo, int next_ld,
                                          int batch, int hi, int wi, int c, void* stream) {
  if (next_ld <= 0) next_ld = c;
  GX_CHECK_ARG(next_ld >= c && next_ld % 4 == 0);
  GX_CHECK_ARG(in && fir_x && fir_y && out && batch > 0 && hi > 0 && wi > 0);
  GX_CHECK_ARG(c % 4 == 0 && ntaps == 4);
  GX_CHECK_ARG(noise == nullptr || noise_strength != nullptr);
  GX_CHECK_ARG(next_style == nullptr || next_hi != nullptr);
  const int ho = hi + pad0 + pad1 - ntaps + 1, wo = wi + pad0 + pad1 - ntaps + 1;
  GX_CHECK_ARG(ho > 0 && wo > 0);
  const int cq = c / 4;
  int cqb = cq < 32 ? cq : 32;
  while (cqb > 1 && (cq % cqb != 0 || 256 % cqb != 0)) --cqb;
  const int tx = 256 / cqb;
  dim3 grid(gx_cdiv(wo, tx), gx_cdiv(ho, BLUR_STRIP), batch * (cq / cqb));
  GX_CHECK_ARG(grid.y <= 65535 && grid.z <= 65535);
  blur_sep_kernel<4><<<grid, 256, 0, (cudaStream_t)stream>>>(
      in, fir_x, fir_y, pad0, noise, noise_batch_stride, noise_strength, bias, act, out, next_style,
      reinterpret_cast<__nv_bfloat16*>(next_hi), reinterpret_cast<__nv_bfloat16*>(next_lo), next_ld, hi, wi, ho, wo, c,
      cqb);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_torgb(const float* x, const float* w, float w_scale, const float* s, const float* bias,
                        const float* skip, float* out, int batch, int hw, int c, void* stream) {
  GX_CHECK_ARG(x && w && s && out && batch > 0 && hw > 0 && c % 4 == 0);
  const long long npix = (long long)batch * hw;
  torgb_kernel<<<gx_cdiv(npix, 8), 256, 0, (cudaStream_t)stream>>>(x, w, w_scale, s, bias, skip, out, batch, hw, c);
  GX_LAUNCH_CHECK();
  return GX_OK;
}
