// Fused nearest-centre assignment of the k-means baseline (ref: clusterer.predict,
// baseline/hfc_kmeans/hfc_kmeans_clustering.py:184) for the shipped shapes: few centres (K <= 64), long rows
// (C = 512 ... 1024 fp32 channels, two same-resolution maps read in place).
//
// HBM-bound: the only large operand is the feature rows, 4 B per element, read ONCE.  The scores x.c_k of a row
// against all K centres are accumulated on the tensor cores straight from registers: a warp owns 16 rows, every
// lane loads 16 B of two rows per 16-channel step, splits them into bf16 hi / lo in registers (fp32-grade
// products, same split as the GEMM route) and issues hi*hi + hi*lo + lo*hi as mma.sync m16n8k16 against centre
// fragments resident in shared memory.  No operand planes, no score matrix in HBM (the GEMM route writes and
// re-reads 2 B + 2 B per element of planes and 4 B per score).  arg-min of ||c_k||^2 - 2 x.c_k in the epilogue.
//
// Measured (B200, 1 GB of rows): 0.83-0.87 of the HBM copy peak for K <= 32; the K = 64 layer is bound by the legacy
// tensor path (ncu: HMMA pipe 75 % busy at 8.7 % of the tcgen05 bf16 peak - mma.sync runs at 1/8 of the tcgen05 rate
// on sm_100) at 0.57.  A tcgen05 version (converter warps writing 128B-swizzled hi / lo planes of 128-row x 64-channel
// blocks to shared memory, accumulator in TMEM, arg-min from tcgen05.ld) was built and verified bit-equal, but its
// converter -> mbarrier -> MMA -> commit ring costs 0.9 us per 32 KB block with the loads switched off (the 128 KB of
// resident centre planes leave room for three stages only) against 0.74 us at the HBM peak: 0.41-0.52 end to end,
// slower than this kernel on every layer, so it is not shipped.
//
// The channel order inside a 16-channel step is permuted consistently for both operands (a dot product does not
// care): lane t of a quad holds channels 4t..4t+3, which the MMA sees as k-slots (2t, 2t+1, 2t+8, 2t+9); the centre
// fragments are stored in exactly the order the lanes read them (one conflict-free LDS.128 per n-tile and step).
#include "gx_common.cuh"
#include <stdlib.h>

namespace {

constexpr int KM_THREADS = 512;
constexpr int KM_WARPS = KM_THREADS / 32;
constexpr int KM_MAX_SMEM = 200 * 1024;

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// hi = bf16x2(x, y), lo = bf16x2(x - hi.x, y - hi.y)
__device__ __forceinline__ void split2(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x, y);
  hi = reinterpret_cast<const uint32_t&>(h);
  const float hx = __uint_as_float(hi << 16), hy = __uint_as_float(hi & 0xffff0000u);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x - hx, y - hy);
  lo = reinterpret_cast<const uint32_t&>(l);
}

// frags[ks][nt][lane] = {hi(ch0,ch1), hi(ch2,ch3), lo(ch0,ch1), lo(ch2,ch3)} of centre nt*8 + lane/4,
// channels 16 ks + 4 (lane % 4) + 0..3; centres >= k are zero
__global__ void kmeans_center_frags_kernel(const float* __restrict__ centers, int k, int c, int nt_count,
                                           uint4* __restrict__ frags) {
  const int total = (c / 16) * nt_count * 32;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int lane = i & 31;
    const int nt = (i >> 5) % nt_count;
    const int ks = (i >> 5) / nt_count;
    const int kc = nt * 8 + (lane >> 2);
    uint4 f = make_uint4(0u, 0u, 0u, 0u);
    if (kc < k) {
      const float4 v = *reinterpret_cast<const float4*>(centers + (long long)kc * c + ks * 16 + 4 * (lane & 3));
      split2(v.x, v.y, f.x, f.z);
      split2(v.z, v.w, f.y, f.w);
    }
    frags[i] = f;
  }
}

template <int NT>
__device__ __forceinline__ void km_step(float (&acc)[NT][4], const float4 va, const float4 vb,
                                        const uint4* __restrict__ fr) {
  uint32_t ah0, al0, ah1, al1, ah2, al2, ah3, al3;
  split2(va.x, va.y, ah0, al0);
  split2(vb.x, vb.y, ah1, al1);
  split2(va.z, va.w, ah2, al2);
  split2(vb.z, vb.w, ah3, al3);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const uint4 b = fr[nt * 32];
    mma_bf16_16816(acc[nt], al0, al1, al2, al3, b.x, b.y);      // lo * hi
    mma_bf16_16816(acc[nt], ah0, ah1, ah2, ah3, b.z, b.w);      // hi * lo
    mma_bf16_16816(acc[nt], ah0, ah1, ah2, ah3, b.x, b.y);      // hi * hi
  }
}

template <int NT, int U>
__global__ void __launch_bounds__(KM_THREADS, 1)
kmeans_assign_mma_kernel(const float* __restrict__ x1, int c1, const float* __restrict__ x2, int c2, long long n,
                         const uint4* __restrict__ frags, const float* __restrict__ cn_pad,
                         int* __restrict__ labels) {
  extern __shared__ uint4 km_smem[];
  const int ksteps = (c1 + c2) / 16;
  const int nfrag = ksteps * NT * 32;
  for (int i = threadIdx.x; i < nfrag; i += KM_THREADS) km_smem[i] = frags[i];
  float* cn_s = reinterpret_cast<float*>(km_smem + nfrag);
  if (threadIdx.x < NT * 8) cn_s[threadIdx.x] = cn_pad[threadIdx.x];
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const long long ntiles = (n + 15) / 16;
  for (long long tile = (long long)blockIdx.x * KM_WARPS + warp; tile < ntiles; tile += (long long)gridDim.x * KM_WARPS) {
    const long long ra = tile * 16 + g, rb = ra + 8;
    const long long rac = ra < n ? ra : n - 1, rbc = rb < n ? rb : n - 1;
    float acc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    const uint4* fr = km_smem + lane;
#pragma unroll 1
    for (int seg = 0; seg < 2; ++seg) {
      const int cs = seg == 0 ? c1 : c2;
      if (cs == 0) continue;
      const float* base = seg == 0 ? x1 : x2;
      const float4* pa = reinterpret_cast<const float4*>(base + rac * cs) + t;
      const float4* pb = reinterpret_cast<const float4*>(base + rbc * cs) + t;
      const int steps = cs / 16;
      int s = 0;
#pragma unroll 1
      for (; s + U <= steps; s += U) {
        float4 va[U], vb[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          va[u] = gx_ldg_stream(pa + (s + u) * 4);
          vb[u] = gx_ldg_stream(pb + (s + u) * 4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) km_step<NT>(acc, va[u], vb[u], fr + (s + u) * NT * 32);
      }
      for (; s < steps; ++s) km_step<NT>(acc, gx_ldg_stream(pa + s * 4), gx_ldg_stream(pb + s * 4), fr + s * NT * 32);
      fr += steps * NT * 32;
    }
    // arg-min of ||c_k||^2 - 2 x.c_k: this lane holds columns nt*8 + 2t + {0,1} of rows g (acc[..][0..1]) and g+8
    float best_a = INFINITY, best_b = INFINITY;
    int ia = 0x7fffffff, ib = 0x7fffffff;
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = nt * 8 + 2 * t + j;
        const float cn = cn_s[idx];
        const float v0 = fmaf(-2.f, acc[nt][j], cn), v1 = fmaf(-2.f, acc[nt][2 + j], cn);
        if (v0 < best_a) { best_a = v0; ia = idx; }          // idx ascends within the lane: first minimum kept
        if (v1 < best_b) { best_b = v1; ib = idx; }
      }
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      const float oa = __shfl_xor_sync(0xffffffffu, best_a, o), ob = __shfl_xor_sync(0xffffffffu, best_b, o);
      const int oia = __shfl_xor_sync(0xffffffffu, ia, o), oib = __shfl_xor_sync(0xffffffffu, ib, o);
      if (oa < best_a || (oa == best_a && oia < ia)) { best_a = oa; ia = oia; }
      if (ob < best_b || (ob == best_b && oib < ib)) { best_b = ob; ib = oib; }
    }
    if (t == 0) {
      if (ra < n) labels[ra] = ia == 0x7fffffff ? 0 : ia;
      if (rb < n) labels[rb] = ib == 0x7fffffff ? 0 : ib;
    }
  }
}

inline int km_nt(int k) { return k <= 8 ? 1 : k <= 16 ? 2 : k <= 32 ? 4 : k <= 64 ? 8 : 0; }

template <int NT, int U>
int km_launch(const float* x1, int c1, const float* x2, int c2, long long n, const void* frags, const float* cn_pad,
              int* labels, cudaStream_t st) {
  const int smem = ((c1 + c2) / 16) * NT * 32 * 16 + NT * 8 * 4;
  static bool attr_done = false;
  if (!attr_done) {
    GX_CHECK_CUDA(cudaFuncSetAttribute(kmeans_assign_mma_kernel<NT, U>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       KM_MAX_SMEM));
    attr_done = true;
  }
  const long long ntiles = (n + 15) / 16;
  int grid = gx_cdiv(ntiles, KM_WARPS);
  const int cap = gx_stream_cta_budget();
  if (grid > cap) grid = cap;
  kmeans_assign_mma_kernel<NT, U><<<grid, KM_THREADS, smem, st>>>(x1, c1, x2, c2, n, (const uint4*)frags, cn_pad,
                                                                 labels);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

}  // namespace

extern "C" long long gx_kmeans_frag_bytes(int k, int c) {
  const int nt = km_nt(k);
  if (nt == 0 || c <= 0 || c % 16) return 0;
  const long long bytes = (long long)(c / 16) * nt * 32 * 16;
  return bytes + nt * 8 * 4 <= KM_MAX_SMEM ? bytes : 0;
}

extern "C" int gx_kmeans_center_frags(const float* centers, int k, int c, void* frags, void* stream) {
  GX_CHECK_ARG(centers && frags && gx_kmeans_frag_bytes(k, c) > 0);
  GX_CHECK_ARG((reinterpret_cast<uintptr_t>(centers) & 15) == 0 && (reinterpret_cast<uintptr_t>(frags) & 15) == 0);
  const int nt = km_nt(k);
  const int total = (c / 16) * nt * 32;
  kmeans_center_frags_kernel<<<gx_cdiv(total, 256), 256, 0, (cudaStream_t)stream>>>(centers, k, c, nt, (uint4*)frags);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_kmeans_assign_mma(const float* x1, int c1, const float* x2, int c2, long long n, const void* frags,
                                    const float* cn_pad, int k, int* labels, void* stream) {
  GX_CHECK_ARG(x1 && frags && cn_pad && labels && n > 0 && c1 > 0 && c1 % 16 == 0 && c2 >= 0 && c2 % 16 == 0);
  GX_CHECK_ARG(c2 == 0 || x2);
  GX_CHECK_ARG(gx_kmeans_frag_bytes(k, c1 + c2) > 0);
  GX_CHECK_ARG((reinterpret_cast<uintptr_t>(x1) & 15) == 0 && (reinterpret_cast<uintptr_t>(x2) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(frags) & 15) == 0);
  cudaStream_t st = (cudaStream_t)stream;
  static const int u_env = getenv("GX_KM_U") ? atoi(getenv("GX_KM_U")) : 0;      // A/B timing of the unroll depth
  switch (km_nt(k)) {
    case 1: return km_launch<1, 4>(x1, c1, x2, c2, n, frags, cn_pad, labels, st);
    case 2: return km_launch<2, 4>(x1, c1, x2, c2, n, frags, cn_pad, labels, st);
    case 4: return u_env == 2 ? km_launch<4, 2>(x1, c1, x2, c2, n, frags, cn_pad, labels, st)
                              : km_launch<4, 4>(x1, c1, x2, c2, n, frags, cn_pad, labels, st);
    case 8: return u_env == 2 ? km_launch<8, 2>(x1, c1, x2, c2, n, frags, cn_pad, labels, st)
                              : km_launch<8, 4>(x1, c1, x2, c2, n, frags, cn_pad, labels, st);
  }
  return GX_ERR_ARG;
}
