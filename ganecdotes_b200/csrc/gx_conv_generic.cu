// Generic 2-D convolution support for the conv2d_gradfix drop-in (ref: lib/gan/optim/conv2d_gradfix.py:129-270 -
// GAN training with R1 / path-length regularisation needs conv gradients of arbitrary order).  A convolution with
// any stride / padding / dilation is  cols = im2col(x)  followed by a tensor-core GEMM (gx_gemm); its adjoint is a
// GEMM followed by col2im.  The two kernels below are an adjoint pair of LINEAR maps, so every derivative of the
// convolution is again built from these two kernels and gx_gemm.  NCHW fp32 tensors; cols [B*Ho*Wo, ld] with the
// column index (c*kh + ky)*kw + kx and ld >= C*kh*kw (padding columns are zero-filled: 16-byte TMA pitch).
#include "gx_common.cuh"

namespace {

struct ConvGeom {
  int b, c, h, w, kh, kw, sy, sx, py, px, dy, dx, ho, wo;
  long long ld;
};

__global__ void im2col_kernel(const float* __restrict__ x, const ConvGeom g, float* __restrict__ cols) {
  const long long total = (long long)g.b * g.ho * g.wo * g.ld;
  const int kk = g.c * g.kh * g.kw;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int col = (int)(i % g.ld);
    long long r = i / g.ld;
    float v = 0.f;
    if (col < kk) {
      const int ox = (int)(r % g.wo); r /= g.wo;
      const int oy = (int)(r % g.ho);
      const int bi = (int)(r / g.ho);
      const int kx = col % g.kw;
      const int ky = (col / g.kw) % g.kh;
      const int ci = col / (g.kw * g.kh);
      const int iy = oy * g.sy - g.py + ky * g.dy, ix = ox * g.sx - g.px + kx * g.dx;
      if (iy >= 0 && iy < g.h && ix >= 0 && ix < g.w) v = x[(((long long)bi * g.c + ci) * g.h + iy) * g.w + ix];
    }
    cols[i] = v;
  }
}

// x[b,c,iy,ix] = sum over the (ky,kx,oy,ox) with oy*sy - py + ky*dy == iy (same in x) of cols[(b,oy,ox),(c,ky,kx)]:
// gather form of the adjoint (no atomics, fixed summation order)
__global__ void col2im_kernel(const float* __restrict__ cols, const ConvGeom g, float* __restrict__ x) {
  const long long total = (long long)g.b * g.c * g.h * g.w;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i;
    const int ix = (int)(r % g.w); r /= g.w;
    const int iy = (int)(r % g.h); r /= g.h;
    const int ci = (int)(r % g.c);
    const int bi = (int)(r / g.c);
    float acc = 0.f;
    for (int ky = 0; ky < g.kh; ++ky) {
      const int ty = iy + g.py - ky * g.dy;
      if (ty < 0 || ty % g.sy) continue;
      const int oy = ty / g.sy;
      if (oy >= g.ho) continue;
      for (int kx = 0; kx < g.kw; ++kx) {
        const int tx = ix + g.px - kx * g.dx;
        if (tx < 0 || tx % g.sx) continue;
        const int ox = tx / g.sx;
        if (ox >= g.wo) continue;
        acc += cols[(((long long)bi * g.ho + oy) * g.wo + ox) * g.ld + (ci * g.kh + ky) * g.kw + kx];
      }
    }
    x[i] = acc;
  }
}

int geom_ok(const ConvGeom& g) {
  return g.b > 0 && g.c > 0 && g.h > 0 && g.w > 0 && g.kh > 0 && g.kw > 0 && g.sy > 0 && g.sx > 0 && g.dy > 0 &&
         g.dx > 0 && g.py >= 0 && g.px >= 0 && g.ho > 0 && g.wo > 0 && g.ld >= (long long)g.c * g.kh * g.kw;
}

}  // namespace

extern "C" int gx_im2col(const float* x, int b, int c, int h, int w, int kh, int kw, int sy, int sx, int py, int px,
                         int dy, int dx, int ho, int wo, long long ld, float* cols, void* stream) {
  const ConvGeom g{b, c, h, w, kh, kw, sy, sx, py, px, dy, dx, ho, wo, ld};
  GX_CHECK_ARG(x && cols && geom_ok(g));
  const long long total = (long long)b * ho * wo * ld;
  im2col_kernel<<<(int)min((long long)gx_sm_count() * 16, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, g, cols);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_col2im(const float* cols, int b, int c, int h, int w, int kh, int kw, int sy, int sx, int py, int px,
                         int dy, int dx, int ho, int wo, long long ld, float* x, void* stream) {
  const ConvGeom g{b, c, h, w, kh, kw, sy, sx, py, px, dy, dx, ho, wo, ld};
  GX_CHECK_ARG(x && cols && geom_ok(g));
  const long long total = (long long)b * c * h * w;
  col2im_kernel<<<(int)min((long long)gx_sm_count() * 16, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(cols, g, x);
  GX_LAUNCH_CHECK();
  return GX_OK;
}
