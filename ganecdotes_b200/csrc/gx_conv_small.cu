// Direct fp32 modulated 3x3 conv for layers with FEW channels (cin, cout <= 32): the 256^2 / 128^2 layers of the BagGAN
// (pidray) generator have 16 / 32 channels (ref models/baggan/models.py:383-390).  On the tcgen05 path such a layer is
// padded to 64-channel operand planes and re-reads a 128-pixel x 64-channel TMA box per tap: 0.60 ms per 16 images for
// 4.8 GFLOP, bound by the SM<->L2 port at 3/4 zero padding.  Here a CTA stages one (16+2) x (32+2) pixel halo tile of
// the style-modulated input in shared memory (channel-major: the per-tap reads of a warp are consecutive words) next to
// the whole weight tensor, and every thread accumulates two output pixels x all output channels in registers with
// packed fp32x2 FMAs: the input is read once, nothing is padded, and the arithmetic is the reference's fp32.
//   out[b,y,x,co] = act(demod[b,co] * sum_{tap,ci} w[tap,ci,co] * (s[b,ci] * x[b,y+dy,x+dx,ci]) + strength*noise + bias[co])
// (ModulatedConv2d.forward models/stylegan2/model.py:327-368 in the algebraic form y = demod * conv(scale*W, s*x),
//  NoiseInjection :371-382, FusedLeakyReLU :15-43); optionally also the next conv's modulated split-bf16 planes.
#include "gx_common.cuh"

namespace {

constexpr int CS_TH = 16, CS_TW = 32;             // output tile
constexpr int CS_HW = CS_TW + 2;                  // halo tile width
constexpr int CS_POS = (CS_TH + 2) * CS_HW;       // 612 halo positions
constexpr int CS_PITCH = CS_POS + 1;              // channel pitch in words (odd: the transposing tile fill spreads over banks)

template <int CIN, int COUT>
__global__ void __launch_bounds__(256)
modconv_small_kernel(const float* __restrict__ x, const float* __restrict__ s, const float* __restrict__ w,
                     const float* __restrict__ demod, const float* __restrict__ noise, long long noise_bstride,
                     const float* __restrict__ noise_strength, const float* __restrict__ bias, int act,
                     float* __restrict__ out, const float* __restrict__ next_style, __nv_bfloat16* __restrict__ next_hi,
                     __nv_bfloat16* __restrict__ next_lo, int next_ld, int h, int wd) {
  extern __shared__ __align__(16) float cs_smem[];
  float* ws = cs_smem;                            // [9][CIN][COUT]
  float* xs = ws + 9 * CIN * COUT;                // [CIN][CS_PITCH]
  __shared__ float sh_style[CIN];
  __shared__ __align__(16) float sh_demod[COUT], sh_bias[COUT], sh_next[COUT];
  const int tid = threadIdx.x;
  const int b = blockIdx.z;
  const int y0 = blockIdx.y * CS_TH, x0 = blockIdx.x * CS_TW;
  if (tid < CIN) sh_style[tid] = __ldg(s + (long long)b * CIN + tid);
  if (tid < COUT) {
    sh_demod[tid] = demod ? __ldg(demod + (long long)b * COUT + tid) : 1.f;
    sh_bias[tid] = bias ? __ldg(bias + tid) : 0.f;
    sh_next[tid] = next_style ? __ldg(next_style + (long long)b * COUT + tid) : 1.f;
  }
  for (int i = tid; i < 9 * CIN * COUT / 4; i += 256)
    reinterpret_cast<float4*>(ws)[i] = __ldg(reinterpret_cast<const float4*>(w) + i);
  __syncthreads();
  // halo tile, modulated by the style, zero outside the image (the conv's padding)
  constexpr int CQ = CIN / 4;
  const float4* xin = reinterpret_cast<const float4*>(x) + (long long)b * h * wd * CQ;
  for (int i = tid; i < CS_POS * CQ; i += 256) {
    const int pos = i / CQ, cq = i - pos * CQ;
    const int py = pos / CS_HW, px = pos - py * CS_HW;
    const int gy = y0 + py - 1, gx = x0 + px - 1;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gy >= 0 && gy < h && gx >= 0 && gx < wd) v = __ldg(xin + ((long long)gy * wd + gx) * CQ + cq);
    float* dst = xs + (4 * cq) * CS_PITCH + pos;
    dst[0] = v.x * sh_style[4 * cq];
    dst[CS_PITCH] = v.y * sh_style[4 * cq + 1];
    dst[2 * CS_PITCH] = v.z * sh_style[4 * cq + 2];
    dst[3 * CS_PITCH] = v.w * sh_style[4 * cq + 3];
  }
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;          // pixels (ty, tx) and (ty + 8, tx) of the tile
  float2 acc0[COUT / 2], acc1[COUT / 2];
#pragma unroll
  for (int i = 0; i < COUT / 2; ++i) acc0[i] = acc1[i] = make_float2(0.f, 0.f);
  const float* xp = xs + ty * CS_HW + tx;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int off = (tap / 3) * CS_HW + (tap % 3);
    const float4* wt = reinterpret_cast<const float4*>(ws + tap * CIN * COUT);
#pragma unroll 4
    for (int ci = 0; ci < CIN; ++ci) {
      const float a0 = xp[ci * CS_PITCH + off], a1 = xp[ci * CS_PITCH + off + 8 * CS_HW];
      const float2 a02 = make_float2(a0, a0), a12 = make_float2(a1, a1);
#pragma unroll
      for (int q = 0; q < COUT / 4; ++q) {
        const float4 w4 = wt[ci * (COUT / 4) + q];        // broadcast read
        acc0[2 * q] = fma2(make_float2(w4.x, w4.y), a02, acc0[2 * q]);
        acc0[2 * q + 1] = fma2(make_float2(w4.z, w4.w), a02, acc0[2 * q + 1]);
        acc1[2 * q] = fma2(make_float2(w4.x, w4.y), a12, acc1[2 * q]);
        acc1[2 * q + 1] = fma2(make_float2(w4.z, w4.w), a12, acc1[2 * q + 1]);
      }
    }
  }
  const float nstr = noise ? __ldg(noise_strength) : 0.f;
  auto emit = [&](const float2 (&acc)[COUT / 2], int half) {
    const int oy = y0 + ty + 8 * half, ox = x0 + tx;
    if (oy >= h || ox >= wd) return;
    const long long pix = ((long long)b * h + oy) * wd + ox;
    float nz = 0.f;
    if (noise) nz = nstr * __ldg(noise + (long long)b * noise_bstride + (long long)oy * wd + ox);
    float4* dst = reinterpret_cast<float4*>(out + pix * COUT);
    uint2* dh = next_hi ? reinterpret_cast<uint2*>(next_hi + pix * next_ld) : nullptr;
    uint2* dl = next_lo ? reinterpret_cast<uint2*>(next_lo + pix * next_ld) : nullptr;
#pragma unroll
    for (int q = 0; q < COUT / 4; ++q) {
      const float4 dm = *reinterpret_cast<const float4*>(sh_demod + 4 * q);
      const float4 bs = *reinterpret_cast<const float4*>(sh_bias + 4 * q);
      // reference order: demodulate, + strength * noise, + bias, leaky-relu(0.2) * sqrt 2
      float4 v = make_float4(acc[2 * q].x * dm.x + nz + bs.x, acc[2 * q].y * dm.y + nz + bs.y,
                             acc[2 * q + 1].x * dm.z + nz + bs.z, acc[2 * q + 1].y * dm.w + nz + bs.w);
      if (act == 1) {
        const float s2 = 1.41421356237309515f;
        v.x = (v.x > 0.f ? v.x : v.x * 0.2f) * s2; v.y = (v.y > 0.f ? v.y : v.y * 0.2f) * s2;
        v.z = (v.z > 0.f ? v.z : v.z * 0.2f) * s2; v.w = (v.w > 0.f ? v.w : v.w * 0.2f) * s2;
      } else if (act == 2) {
        v.x = v.x > 0.f ? v.x : v.x * 0.2f; v.y = v.y > 0.f ? v.y : v.y * 0.2f;
        v.z = v.z > 0.f ? v.z : v.z * 0.2f; v.w = v.w > 0.f ? v.w : v.w * 0.2f;
      }
      dst[q] = v;
      if (dh) {
        const float4 st = *reinterpret_cast<const float4*>(sh_next + 4 * q);
        uint2 hh, ll;
        gx_split4(make_float4(v.x * st.x, v.y * st.y, v.z * st.z, v.w * st.w), hh, ll);
        dh[q] = hh;
        if (dl) dl[q] = ll;
      }
    }
  };
  emit(acc0, 0);
  emit(acc1, 1);
}

template <int CIN, int COUT>
int launch_small(const float* x, const float* s, const float* w, const float* demod, const float* noise,
                 long long nbs, const float* nstr, const float* bias, int act, float* out, const float* next_style,
                 void* next_hi, void* next_lo, int next_ld, int batch, int h, int wd, cudaStream_t st) {
  const size_t smem = (size_t)(9 * CIN * COUT + CIN * CS_PITCH) * sizeof(float);
  static bool attr = false;
  if (!attr) {
    GX_CHECK_CUDA(cudaFuncSetAttribute(modconv_small_kernel<CIN, COUT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    attr = true;
  }
  const dim3 grid(gx_cdiv(wd, CS_TW), gx_cdiv(h, CS_TH), batch);
  modconv_small_kernel<CIN, COUT><<<grid, 256, smem, st>>>(
      x, s, w, demod, noise, nbs, nstr, bias, act, out, next_style, reinterpret_cast<__nv_bfloat16*>(next_hi),
      reinterpret_cast<__nv_bfloat16*>(next_lo), next_ld, h, wd);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

}  // namespace

extern "C" int gx_modconv_small(const float* x, const float* style, const float* w, const float* demod,
                                const float* noise, long long noise_batch_stride, const float* noise_strength,
                                const float* bias, int act, float* out, const float* next_style, void* next_hi,
                                void* next_lo, int next_ld, int batch, int h, int wd, int cin, int cout,
                                void* stream) {
  GX_CHECK_ARG(x && style && w && out && batch > 0 && h > 0 && wd > 0);
  GX_CHECK_ARG(batch <= 65535 && gx_cdiv(h, CS_TH) <= 65535);
  GX_CHECK_ARG(noise == nullptr || noise_strength != nullptr);
  GX_CHECK_ARG(next_style == nullptr || next_hi != nullptr);
  if (next_ld <= 0) next_ld = cout;
  GX_CHECK_ARG(next_hi == nullptr || (next_ld >= cout && next_ld % 4 == 0));
  cudaStream_t st = (cudaStream_t)stream;
#define GX_CS(CI, CO)                                                                                              \
  if (cin == CI && cout == CO)                                                                                     \
    return launch_small<CI, CO>(x, style, w, demod, noise, noise_batch_stride, noise_strength, bias, act, out,      \
                                next_style, next_hi, next_lo, next_ld, batch, h, wd, st)
  GX_CS(16, 16);
  GX_CS(32, 32);
  GX_CS(32, 16);
  GX_CS(16, 32);
  GX_CS(8, 8);
#undef GX_CS
  return GX_ERR_ARG;      // other channel counts run on the tcgen05 path (gx_modconv)
}

extern "C" int gx_modconv_small_supported(int cin, int cout) {
  return ((cin == 16 || cin == 32) && (cout == 16 || cout == 32)) || (cin == 8 && cout == 8);
}
