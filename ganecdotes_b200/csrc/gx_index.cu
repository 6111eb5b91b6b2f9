// Index bookkeeping of the all-pixel ("dedup") training path: which samples of the 2 x P patches fall on which
// pixel, as a CSR list (ref: the random-pixel sampling of hfc_with_swav/swav_clustering.py:158-167 applied to the
// rotated / flipped tensor, :358-359 - here a gather of rows of Z by pixel index, and for the backward a
// per-pixel segment sum of the dZ rows).  A deterministic counting sort: histogram -> exclusive scan -> scatter
// with atomic cursors -> every (tiny) segment sorted by sample index, so the summation order of
// gx_segment_sum_rows does not depend on the order the atomics were served in.
#include "gx_common.cuh"

namespace {

// ridx[p*bn + j] = row_img[j]*hw + row_src[p*bn + j]  (or -1 on rotation fill); counts[pixel] += 1
__global__ void pixel_keys_kernel(const int* __restrict__ row_src, const int* __restrict__ row_img, long long total,
                                  long long bn, int hw, int patches_per_group, int img_group_stride,
                                  int* __restrict__ ridx, int* __restrict__ counts) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int src = row_src[i];
    const int img = row_img[i % bn] + (int)((i / bn) / patches_per_group) * img_group_stride;
    const int key = src >= 0 ? img * hw + src : -1;
    ridx[i] = key;
    if (key >= 0) atomicAdd(counts + key, 1);
  }
}

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;                        // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int* smem /* [33] */, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < (SCAN_THREADS >> 5) ? smem[lane] : 0;
    int wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    smem[lane] = wi - w;                              // exclusive warp offsets
    if (lane == 31) smem[32] = wi;
  }
  __syncthreads();
  total = smem[32];
  const int r = smem[warp] + incl - v;
  __syncthreads();
  return r;
}

// phase 1: tile sums
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_sums_kernel(const int* __restrict__ counts, long long n,
                                                                       int* __restrict__ tile_sums) {
  __shared__ int sm[33];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int v = 0;
#pragma unroll
  for (int e = 0; e < SCAN_ITEMS; ++e)
    if (base + e < n) v += counts[base + e];
  int total;
  block_exclusive_scan(v, sm, total);
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// phase 2: exclusive scan of the tile sums in place (one block, any number of tiles), grand total -> seg_off[n]
__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_offsets_kernel(int* __restrict__ tile_sums, int ntiles,
                                                                          int* __restrict__ grand_total) {
  __shared__ int sm[33];
  int carry = 0;
  for (int t0 = 0; t0 < ntiles; t0 += SCAN_THREADS) {
    const int i = t0 + threadIdx.x;
    const int v = i < ntiles ? tile_sums[i] : 0;
    int total;
    const int ex = block_exclusive_scan(v, sm, total);
    if (i < ntiles) tile_sums[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0) *grand_total = carry;
}

// phase 3: seg_off[i] = tile offset + exclusive scan inside the tile; cursors := 0
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(int* __restrict__ counts, long long n,
                                                                   const int* __restrict__ tile_offs,
                                                                   int* __restrict__ seg_off) {
  __shared__ int sm[33];
  const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_ITEMS;
  int c[SCAN_ITEMS], v = 0;
#pragma unroll
  for (int e = 0; e < SCAN_ITEMS; ++e) {
    c[e] = base + e < n ? counts[base + e] : 0;
    v += c[e];
  }
  int total;
  int off = tile_offs[blockIdx.x] + block_exclusive_scan(v, sm, total);
#pragma unroll
  for (int e = 0; e < SCAN_ITEMS; ++e)
    if (base + e < n) {
      seg_off[base + e] = off;
      off += c[e];
      counts[base + e] = 0;                           // re-used as the scatter cursor
    }
}

__global__ void scatter_samples_kernel(const int* __restrict__ ridx, long long total, const int* __restrict__ seg_off,
                                       int* __restrict__ cursor, int* __restrict__ order) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int key = ridx[i];
    if (key >= 0) order[seg_off[key] + atomicAdd(cursor + key, 1)] = (int)i;
  }
}

// ascending sample index inside every segment (insertion sort: a pixel is hit by a handful of samples)
__global__ void sort_segments_kernel(const int* __restrict__ seg_off, long long nseg, int* __restrict__ order) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < nseg;
       s += (long long)gridDim.x * blockDim.x) {
    const int b = seg_off[s], e = seg_off[s + 1];
    for (int i = b + 1; i < e; ++i) {
      const int v = order[i];
      int j = i - 1;
      while (j >= b && order[j] > v) {
        order[j + 1] = order[j];
        --j;
      }
      order[j + 1] = v;
    }
  }
}

// W+ of both perturbed views in one launch (ref: swav_clustering.py:593-640, image_augmentor.py:42-53,75-79)
__global__ void view_wplus_kernel(const float* __restrict__ w, const float* __restrict__ noise_w,
                                  const int* __restrict__ layer_no, const float* __restrict__ sigma,
                                  const float* __restrict__ mean, float psi, int b, int rows, int n_latent, int dim,
                                  float* __restrict__ out) {
  const long long total = (long long)rows * n_latent * dim;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % dim);
    const int r = (int)((i / dim) % n_latent);
    const int row = (int)(i / ((long long)dim * n_latent));
    const float m = mean[d];
    float v = w[(long long)(row % b) * dim + d];
    if (psi < 1.f) v = m + psi * (v - m);                                   // first truncation (:603-607)
    const int l2 = 2 * layer_no[row];
    if (r == l2 || r == l2 + 1) {
      const float sg = sigma[row];
      v = (1.f - sg) * v + sg * noise_w[((long long)2 * row + (r - l2)) * dim + d];
    }
    if (psi < 1.f) v = m + psi * (v - m);                                   // second truncation (aug:75-79)
    out[i] = v;
  }
}

// out[col] = (accumulate ? out[col] : 0) + scale * sum_p parts[p, col]   (fixed summation order)
__global__ void __launch_bounds__(256)
colsum_scale_kernel(const float* __restrict__ parts, int nparts, int k, float scale, int accumulate,
                    float* __restrict__ out) {
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
    out[col] = (accumulate ? out[col] : 0.f) + scale * t;
  }
}

}  // namespace

extern "C" int gx_pixel_segments(const int* row_src, const int* row_img, int patches, long long bn, int hw,
                                 long long npix, int patches_per_group, int img_group_stride, int* ridx, int* counts,
                                 int* tile_scratch, int* seg_off, int* order, void* stream) {
  GX_CHECK_ARG(row_src && row_img && ridx && counts && tile_scratch && seg_off && order);
  if (patches_per_group <= 0) { patches_per_group = patches; img_group_stride = 0; }
  GX_CHECK_ARG(patches > 0 && bn > 0 && hw > 0 && npix > 0 && npix < (1LL << 31) && patches * bn < (1LL << 31));
  cudaStream_t st = (cudaStream_t)stream;
  const long long total = (long long)patches * bn;
  const int ntiles = gx_cdiv(npix, SCAN_TILE);
  const int grid = (int)min((long long)gx_sm_count() * 8, (total + 255) / 256);
  GX_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)npix * sizeof(int), st));
  pixel_keys_kernel<<<grid, 256, 0, st>>>(row_src, row_img, total, bn, hw, patches_per_group, img_group_stride, ridx,
                                          counts);
  scan_tile_sums_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(counts, npix, tile_scratch);
  scan_tile_offsets_kernel<<<1, SCAN_THREADS, 0, st>>>(tile_scratch, ntiles, seg_off + npix);
  scan_apply_kernel<<<ntiles, SCAN_THREADS, 0, st>>>(counts, npix, tile_scratch, seg_off);
  scatter_samples_kernel<<<grid, 256, 0, st>>>(ridx, total, seg_off, counts, order);
  sort_segments_kernel<<<(int)min((long long)gx_sm_count() * 8, (npix + 255) / 256), 256, 0, st>>>(seg_off, npix, order);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_pixel_segments_scratch(long long npix) { return gx_cdiv(npix, SCAN_TILE); }

extern "C" int gx_view_wplus(const float* w, const float* noise_w, const int* layer_no, const float* sigma,
                             const float* mean, float psi, int b, int rows, int n_latent, int dim, float* out,
                             void* stream) {
  GX_CHECK_ARG(w && noise_w && layer_no && sigma && mean && out && b > 0 && rows > 0 && n_latent > 0 && dim > 0);
  const long long total = (long long)rows * n_latent * dim;
  view_wplus_kernel<<<(int)min((long long)gx_sm_count() * 8, (total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      w, noise_w, layer_no, sigma, mean, psi, b, rows, n_latent, dim, out);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

extern "C" int gx_colsum(const float* parts, int nparts, int k, float scale, int accumulate, float* out,
                         void* stream) {
  GX_CHECK_ARG(parts && out && nparts > 0 && k > 0);
  colsum_scale_kernel<<<gx_cdiv(k, 32), 256, 0, (cudaStream_t)stream>>>(parts, nparts, k, scale, accumulate, out);
  GX_LAUNCH_CHECK();
  return GX_OK;
}
