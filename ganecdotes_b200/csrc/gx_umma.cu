// tcgen05 / TMEM / TMA contraction engine for sm_100a.
//
// One persistent, warp-specialised kernel serves every dense contraction on the path:
//   * gx_gemm    : C[M,N] = A * B^T (+bias), A/B as bf16 planes (K- or MN-major), split-K
//                  -> projection / prototype scores / their gradient GEMMs
//   * gx_modconv : modulated 3x3 conv (plain, or transposed stride 2 as 4 sub-pixel
//                  phases) as implicit GEMM: the A tile of a tap is ONE 4-D TMA box of
//                  the NHWC activation shifted by the tap offset; TMA's out-of-bounds
//                  zero fill is the conv padding.
// Roles (384 threads, 1 CTA / SM): warp 0 = TMA producer, warp 1 = MMA issuer,
// warp 2 = TMEM allocator, warps 4-11 = epilogue (TMEM -> registers -> smem transpose -> global).
// Pipelines: smem ring (full/empty mbarriers) and a 2-deep TMEM accumulator ring
// (tmem_full/tmem_empty) so the epilogue of tile i overlaps the MMAs of tile i+1.
// Precision: passes == 3 issues A_lo*B_hi + A_hi*B_lo + A_hi*B_hi into the same fp32
// accumulator (split-bf16, ~2^-16 relative per product); passes == 1 is plain bf16.
#include <stdlib.h>

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "gx_common.cuh"
#include "gx_ptx.cuh"

using namespace gxptx;

namespace {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle atom
constexpr int A_PLANE_BYTES = BM * BK * 2;
constexpr int MAX_STAGES = 8;
constexpr int NTHREADS = 384;      // warps 0-3: producer / MMA / TMEM alloc / spare, warps 4-11: epilogue
constexpr int EPI_WARPS = 8;       // two warps per TMEM lane quarter, each taking every other 32-column chunk
constexpr int TMEM_COLS = 512;
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int STAGING_BYTES = EPI_WARPS * 32 * 33 * 4;  // per-warp 32x33 fp32 transpose tiles of the epilogue

struct UmmaParams {
  int mode;  // 0 = gemm, 1 = conv
  int passes, block_n, stages;
  int debug;       // bit 0: skip epilogue global stores (profiling aid, GX_UMMA_DEBUG)
  int mtiles;      // 128-row sub-tiles per CTA tile (2 = 256-row tiles, halves B traffic per FLOP)
  int acc_stages;  // TMEM accumulator ring depth (2 when mtiles*block_n <= 256)
  int a_mn, b_mn;
  int f16;         // operand planes are fp16 (single-pass only) instead of bf16
  int nsub;        // pair gemm: 256-column accumulator sub-tiles per CTA tile (2 = 256 x 512 pair tiles: one A tile
                   // feeds two N=256 MMAs, a quarter less operand ingest per FLOP; the whole TMEM is one accumulator)
  int pair;        // gemm: CTA pairs (clusters of 2, tcgen05 cta_group::2) on adjacent 128-row tiles of one
                   // n-tile: one M=256 MMA spans both SMs, each CTA loads its A rows and HALF of the B tile
                   // (a third less operand ingest per SM and FLOP than two independent 128-row tiles)
  // ---- gemm
  int M, N, K;
  int tiles_m, tiles_n, split_k, kiters_total;
  float* c;
  long long ldc;
  const float* bias;
  int atomic;
  float* colexp_sum;   // optional [N]: += sum_m exp2(colexp_scale * C[m,n])  (first Sinkhorn pass fused)
  float colexp_scale;
  // ---- conv
  int B, H, W, Cin, Cout, upsample;
  int th, tw, nb;
  int Ho, Wo;
  int nphases;
  int phase_tile_start[5];
  int phase_ty[4], phase_tx[4], phase_eh[4], phase_ew[4], phase_a[4], phase_b[4];
  int ntaps[4];
  signed char tap_dy[4][9], tap_dx[4][9], tap_w[4][9];
  const float* demod;
  const float* noise;
  long long noise_bstride;
  const float* noise_strength;
  int act;
  float* out;
  const float* next_style;
  __nv_bfloat16* next_hi;
  __nv_bfloat16* next_lo;
  int next_ld;
  float* ws;            // split-K conv: per-split partial results [split_k][B*Ho*Wo*Cout]
  long long ws_stride;
};

struct TileInfo {
  // gemm
  int m0, n0, kbeg, kend;
  // conv
  int phase, b0, y0, x0;
  int split;
};

__device__ __forceinline__ TileInfo decode_tile(const UmmaParams& p, int w, int rank = 0) {
  TileInfo t;
  const int tiles_m_units = p.pair ? (p.tiles_m + 1) >> 1 : p.tiles_m;
  const int tiles_mn = tiles_m_units * p.tiles_n;
  const int split = w / tiles_mn;
  const int tmn = w - split * tiles_mn;
  const int tile_n = tmn % p.tiles_n;
  const int tile_m = p.pair ? 2 * (tmn / p.tiles_n) + rank : tmn / p.tiles_n;
  t.n0 = tile_n * p.block_n * p.nsub;
  t.m0 = tile_m * BM * p.mtiles;
  t.phase = 0; t.b0 = 0; t.y0 = 0; t.x0 = 0;
  t.split = split;
  if (p.mode == 0) {
    const int per = (p.kiters_total + p.split_k - 1) / p.split_k;
    t.kbeg = split * per;
    t.kend = min(p.kiters_total, t.kbeg + per);
  } else {
    int ph = 0;
    while (ph + 1 < p.nphases && tile_m >= p.phase_tile_start[ph + 1]) ++ph;
    int r = tile_m - p.phase_tile_start[ph];
    const int tx = r % p.phase_tx[ph];
    r /= p.phase_tx[ph];
    const int ty = r % p.phase_ty[ph];
    const int g = r / p.phase_ty[ph];
    t.phase = ph;
    t.b0 = g * p.nb;
    t.y0 = ty * p.th;
    t.x0 = tx * p.tw;
    // split-K (few pixel tiles): every split owns a non-empty slice of this phase's tap x channel-block loop
    const int ktot = p.ntaps[ph] * (p.Cin / BK);
    const int per = (ktot + p.split_k - 1) / p.split_k;
    t.kbeg = split * per;
    t.kend = min(ktot, t.kbeg + per);
  }
  return t;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float lrelu_sqrt2(float v) {
  return (v > 0.f ? v : v * 0.2f) * 1.41421356237309515f;
}

// Epilogue of one 32x32 accumulator chunk of a warp (thread = row on entry).  Staging tile of 32 rows x
// 8 16-byte pieces, piece j of row r at physical piece j ^ (r & 7): conflict-free both for the row owner's
// 128-bit writes and for the reads, after which 8 lanes hold one row (4 columns per lane) so that every
// 128-bit global store of the warp covers four contiguous 128 B row segments.
template <bool COLEXP, bool ATOMIC>
__device__ __forceinline__ void epi_chunk_vec(const uint32_t (&r)[32], uint32_t stg, int lane, float* cptr,
                                              long long rstep, const float* bias4, int rows_valid, float colexp_scale,
                                              float* colsum4, bool do_store) {
  // issued first: its latency hides behind the smem transpose
  float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
  if (bias4 != nullptr) bv = __ldg(reinterpret_cast<const float4*>(bias4));
  {
    const uint32_t wa = stg + (uint32_t)(lane * 128);
    const uint32_t sw = (uint32_t)(lane & 7);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wa + ((j ^ sw) << 4)), "r"(r[4 * j]),
                   "r"(r[4 * j + 1]), "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                   : "memory");
  }
  __syncwarp();
  const int lr = lane >> 3;
  float2 v[8][2];
  {
    // row 4g + lr: (row & 7) = (4g + lr) & 7 alternates between lr and lr + 4
    const uint32_t ra = stg + (uint32_t)(lr * 128);
    const uint32_t p0 = (uint32_t)(((lane & 7) ^ lr) << 4), p1 = (uint32_t)(((lane & 7) ^ (lr + 4)) << 4);
#pragma unroll
    for (int g = 0; g < 8; ++g)
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(v[g][0].x), "=f"(v[g][0].y), "=f"(v[g][1].x), "=f"(v[g][1].y)
                   : "r"(ra + (uint32_t)(g * 512) + ((g & 1) ? p1 : p0))
                   : "memory");
  }
  __syncwarp();   // the tile may be overwritten from here on; the rest works from registers
  if (bias4 != nullptr) {
    const float2 b0 = make_float2(bv.x, bv.y), b1 = make_float2(bv.z, bv.w);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      v[g][0] = add2(v[g][0], b0);
      v[g][1] = add2(v[g][1], b1);
    }
  }
  const bool full = rows_valid == 32;   // warp-uniform
  if (do_store) {
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (full || 4 * g + lr < rows_valid) {
        if (ATOMIC)
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cptr), "f"(v[g][0].x), "f"(v[g][0].y),
                       "f"(v[g][1].x), "f"(v[g][1].y)
                       : "memory");
        else
          *reinterpret_cast<float4*>(cptr) = make_float4(v[g][0].x, v[g][0].y, v[g][1].x, v[g][1].y);
      }
      cptr += rstep;
    }
  }
  if (COLEXP) {
    const float2 sc = make_float2(colexp_scale, colexp_scale);
    float2 e0 = make_float2(0.f, 0.f), e1 = make_float2(0.f, 0.f);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      if (full || 4 * g + lr < rows_valid) {
        const float2 t0 = mul2(v[g][0], sc), t1 = mul2(v[g][1], sc);
        e0 = add2(e0, make_float2(ex2_approx(t0.x), ex2_approx(t0.y)));
        e1 = add2(e1, make_float2(ex2_approx(t1.x), ex2_approx(t1.y)));
      }
    }
    // fold the 4 row groups, then one vector reduction per column quad
    e0.x += __shfl_xor_sync(0xffffffffu, e0.x, 8);  e0.y += __shfl_xor_sync(0xffffffffu, e0.y, 8);
    e1.x += __shfl_xor_sync(0xffffffffu, e1.x, 8);  e1.y += __shfl_xor_sync(0xffffffffu, e1.y, 8);
    e0.x += __shfl_xor_sync(0xffffffffu, e0.x, 16); e0.y += __shfl_xor_sync(0xffffffffu, e0.y, 16);
    e1.x += __shfl_xor_sync(0xffffffffu, e1.x, 16); e1.y += __shfl_xor_sync(0xffffffffu, e1.y, 16);
    if (lr == 0 && colsum4 != nullptr)
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(colsum4), "f"(e0.x), "f"(e0.y), "f"(e1.x),
                   "f"(e1.y)
                   : "memory");
  }
}

// PAIR = true: the cta_group::2 instantiation, only ever launched as clusters of two CTAs
template <bool PAIR>
__global__ void __launch_bounds__(NTHREADS, 1)
gx_umma_kernel(const UmmaParams p, const __grid_constant__ CUtensorMap tm_a_hi,
               const __grid_constant__ CUtensorMap tm_a_lo, const __grid_constant__ CUtensorMap tm_b_hi,
               const __grid_constant__ CUtensorMap tm_b_lo, int total_work) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve (stage buffers must be 1024-aligned for the 128B swizzle)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_sub_bytes = (PAIR ? p.block_n / 2 : p.block_n) * BK * 2;
  const int b_plane_bytes = b_sub_bytes * p.nsub;
  const int nplanes = (p.passes == 3) ? 2 : 1;
  const int a_plane_bytes = A_PLANE_BYTES * p.mtiles;
  const int stage_bytes = nplanes * (a_plane_bytes + b_plane_bytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + MAX_STAGES;
  uint64_t* tfull_bar = bars + 2 * MAX_STAGES;
  uint64_t* tempty_bar = bars + 2 * MAX_STAGES + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);
  float* staging = reinterpret_cast<float*>(bars + 2 * MAX_STAGES + 6);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = PAIR ? (int)cluster_ctarank() : 0;
  const int w_first = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int w_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a_hi);
    prefetch_tmap(&tm_b_hi);
    if (p.passes == 3) {
      prefetch_tmap(&tm_a_lo);
      prefetch_tmap(&tm_b_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      // pair: the leader's MMA thread waits for the epilogue warps of both CTAs
      mbar_init(&tempty_bar[i], PAIR ? 2 * EPI_WARPS : EPI_WARPS);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const int cblocks = (p.mode == 1) ? p.Cin / BK : 1;
      for (int w = w_first; w < total_work; w += w_step) {
        const TileInfo t = decode_tile(p, w, crank);
        for (int kit = t.kbeg; kit < t.kend; ++kit) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * stage_bytes;
          uint8_t* sb = sa + nplanes * a_plane_bytes;
          if constexpr (PAIR) {
            // both CTAs' loads are counted on the LEADER's barrier (its MMA thread consumes both halves)
            if (crank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (uint32_t)stage_bytes);
            const uint32_t lead_bar = cluster_addr_of(&full_bar[stage], 0);
            const int k0 = kit * BK;
            const int half_rows = p.block_n >> 1;
            for (int pl = 0; pl < nplanes; ++pl) {
              const CUtensorMap* ma = pl ? &tm_a_lo : &tm_a_hi;
              const CUtensorMap* mb = pl ? &tm_b_lo : &tm_b_hi;
              uint8_t* da = sa + pl * a_plane_bytes;
              uint8_t* db = sb + pl * b_plane_bytes;
              if (p.mode == 1) {
                // conv: this CTA's own pixel tile shifted by the tap, and its half of the weight tile
                const int tap = kit / cblocks;
                const int c0 = (kit - tap * cblocks) * BK;
                const int dy = p.tap_dy[t.phase][tap], dx = p.tap_dx[t.phase][tap];
                const int wk = p.tap_w[t.phase][tap];
                tma_load_4d_pair(da, ma, lead_bar, c0, t.x0 + dx, t.y0 + dy, t.b0);
                tma_load_2d_pair(db, mb, lead_bar, wk * p.Cin + c0, t.n0 + crank * half_rows);
                continue;
              }
              if (!p.a_mn) {
                tma_load_2d_pair(da, ma, lead_bar, k0, t.m0);
              } else {
                for (int i = 0; i < 2; ++i) tma_load_2d_pair(da + i * 8192, ma, lead_bar, t.m0 + i * 64, k0);
              }
              for (int j = 0; j < p.nsub; ++j) {     // this CTA's half of every 256-column sub-tile of B
                const int nrow = t.n0 + j * p.block_n + crank * half_rows;
                if (!p.b_mn) {
                  tma_load_2d_pair(db + j * b_sub_bytes, mb, lead_bar, k0, nrow);
                } else {
                  for (int i = 0; i < half_rows / 64; ++i)
                    tma_load_2d_pair(db + j * b_sub_bytes + i * 8192, mb, lead_bar, nrow + i * 64, k0);
                }
              }
            }
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
            continue;
          }
          mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)stage_bytes);
          for (int pl = 0; pl < nplanes; ++pl) {
            const CUtensorMap* ma = pl ? &tm_a_lo : &tm_a_hi;
            const CUtensorMap* mb = pl ? &tm_b_lo : &tm_b_hi;
            uint8_t* da = sa + pl * a_plane_bytes;
            uint8_t* db = sb + pl * b_plane_bytes;
            if (p.mode == 0) {
              const int k0 = kit * BK;
              if (!p.a_mn) {
                for (int i = 0; i < p.mtiles; ++i)
                  tma_load_2d(da + i * A_PLANE_BYTES, ma, &full_bar[stage], k0, t.m0 + i * BM);
              } else {
                for (int i = 0; i < 2 * p.mtiles; ++i)
                  tma_load_2d(da + i * 8192, ma, &full_bar[stage], t.m0 + i * 64, k0);
              }
              if (!p.b_mn) {
                tma_load_2d(db, mb, &full_bar[stage], k0, t.n0);
              } else {
                for (int i = 0; i < p.block_n / 64; ++i)
                  tma_load_2d(db + i * 8192, mb, &full_bar[stage], t.n0 + i * 64, k0);
              }
            } else {
              const int tap = kit / cblocks;
              const int c0 = (kit - tap * cblocks) * BK;
              const int dy = p.tap_dy[t.phase][tap], dx = p.tap_dx[t.phase][tap];
              const int wk = p.tap_w[t.phase][tap];
              tma_load_4d(da, ma, &full_bar[stage], c0, t.x0 + dx, t.y0 + dy, t.b0);
              tma_load_2d(db, mb, &full_bar[stage], wk * p.Cin + c0, t.n0);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ================================
    if (lane == 0 && crank == 0) {     // pair: the leader CTA issues for both SMs
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * BM : BM, p.block_n, p.a_mn, p.b_mn, p.f16);
      const uint32_t a_lbo = p.a_mn ? 8192u : 0u, b_lbo = p.b_mn ? 8192u : 0u;
      const uint32_t a_kstep = p.a_mn ? 2048u : 32u, b_kstep = p.b_mn ? 2048u : 32u;
      int it = 0;
      for (int w = w_first; w < total_work; w += w_step, ++it) {
        const TileInfo t = decode_tile(p, w, crank);
        const int acc = it % p.acc_stages;
        const uint32_t acc_phase = (it / p.acc_stages) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * p.mtiles * p.block_n);
        uint32_t accumulate = 0;
        for (int kit = t.kbeg; kit < t.kend; ++kit) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * stage_bytes);
          const uint32_t sb = sa + nplanes * a_plane_bytes;
          for (int ps = 0; ps < p.passes; ++ps) {
            // passes==3: (A_lo,B_hi) (A_hi,B_lo) (A_hi,B_hi); passes==1: (A_hi,B_hi)
            const int apl = (p.passes == 3 && ps == 0) ? 1 : 0;
            const int bpl = (p.passes == 3 && ps == 1) ? 1 : 0;
            const uint32_t abase = sa + apl * a_plane_bytes;
            const uint32_t bbase = sb + bpl * b_plane_bytes;
#pragma unroll
            for (int k4 = 0; k4 < BK / 16; ++k4) {
              const uint64_t bdesc = make_smem_desc(bbase + k4 * b_kstep, b_lbo, 1024u);
              if constexpr (PAIR) {
                const uint64_t adesc = make_smem_desc(abase + k4 * a_kstep, a_lbo, 1024u);
                umma_bf16_pair(d_tmem, adesc, bdesc, idesc, accumulate);
                if (p.nsub == 2)
                  umma_bf16_pair(d_tmem + (uint32_t)p.block_n, adesc,
                                 make_smem_desc(bbase + b_sub_bytes + k4 * b_kstep, b_lbo, 1024u), idesc, accumulate);
              } else {
                for (int sub = 0; sub < p.mtiles; ++sub) {
                  const uint64_t adesc = make_smem_desc(abase + sub * A_PLANE_BYTES + k4 * a_kstep, a_lbo, 1024u);
                  umma_bf16(d_tmem + (uint32_t)(sub * p.block_n), adesc, bdesc, idesc, accumulate);
                }
              }
              accumulate = 1;
            }
          }
          // frees the smem slot when these MMAs retire (pair: in both CTAs)
          if constexpr (PAIR) umma_commit_pair(&empty_bar[stage], (uint16_t)3);
          else umma_commit(&empty_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        // accumulator complete -> epilogue (pair: of both CTAs)
        if constexpr (PAIR) umma_commit_pair(&tfull_bar[acc], (uint16_t)3);
        else umma_commit(&tfull_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    // ============================== epilogue ==================================
    const int q = warp & 3;             // TMEM lane quarter owned by this warp
    const int ehalf = (warp - 4) >> 2;  // which half of the chunks this warp handles
    const int row = q * 32 + lane;      // tile row == TMEM lane
    int it = 0;
    for (int w = w_first; w < total_work; w += w_step, ++it) {
      const TileInfo t = decode_tile(p, w, crank);
      const int acc = it % p.acc_stages;
      const uint32_t acc_phase = (it / p.acc_stages) & 1;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * p.mtiles * p.block_n);
      const int nchunks = p.block_n * p.nsub / 32;

      if (p.mode == 0) {
        // TMEM -> registers (thread = row) -> per-warp 32x33 smem tile -> registers (8 lanes per
        // row, 4 columns per lane) so that every 128-bit global store of a warp covers four
        // contiguous 128 B row segments.
        const uint32_t stg = smem_u32(staging) + (uint32_t)((warp - 4) * (32 * 33 * 4));
        const bool add_bias = p.bias != nullptr && (!p.atomic || t.kbeg == 0);
        const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.c) & 15) == 0);
        const int lr = lane >> 3, lc = (lane & 7) * 4;   // vector path: row-in-group, first column
        for (int sub = 0; sub < p.mtiles; ++sub) {
          const long long mrow0 = (long long)t.m0 + sub * BM + q * 32;   // first row of this warp
          if (mrow0 >= p.M) break;                                        // warp-uniform
          const int rows_valid = (int)min((long long)32, (long long)p.M - mrow0);
          const uint32_t tsub = taddr0 + (uint32_t)(sub * p.block_n);
          auto process = [&](const uint32_t (&r)[32], int ch) {
            const int n_base = t.n0 + ch * 32;
            if (p.debug & 2) return;
            if (vec_ok && n_base + 32 <= p.N) {
              float* cptr = p.c + (mrow0 + lr) * p.ldc + n_base + lc;
              const float* bptr = add_bias ? p.bias + n_base + lc : nullptr;
              float* uptr = (p.colexp_sum != nullptr && !(p.debug & 4)) ? p.colexp_sum + n_base + lc : nullptr;
              if (p.colexp_sum != nullptr)
                epi_chunk_vec<true, false>(r, stg, lane, cptr, 4 * p.ldc, bptr, rows_valid, p.colexp_scale, uptr,
                                           !(p.debug & 1));
              else if (p.atomic)
                epi_chunk_vec<false, true>(r, stg, lane, cptr, 4 * p.ldc, bptr, rows_valid, 0.f, nullptr, true);
              else
                epi_chunk_vec<false, false>(r, stg, lane, cptr, 4 * p.ldc, bptr, rows_valid, 0.f, nullptr,
                                            !(p.debug & 1));
            } else {
              {
                const uint32_t wa = stg + (uint32_t)(lane * 33 * 4);
#pragma unroll
                for (int i = 0; i < 32; ++i)
                  asm volatile("st.shared.b32 [%0], %1;" ::"r"(wa + i * 4), "r"(r[i]) : "memory");
              }
              __syncwarp();
              const int n = n_base + lane;
              const bool nvalid = n < p.N;
              const float bv = (add_bias && nvalid) ? __ldg(p.bias + n) : 0.f;
              float* cptr = p.c + mrow0 * p.ldc + n;
              const uint32_t ra = stg + (uint32_t)(lane * 4);
              float es = 0.f;
              for (int rr = 0; rr < rows_valid; ++rr) {
                float v;
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(ra + (uint32_t)(rr * 33 * 4)));
                if (nvalid) {
                  if (p.atomic) atomicAdd(cptr, v + bv);
                  else cptr[0] = v + bv;
                  if (p.colexp_sum != nullptr) es += ex2_approx((v + bv) * p.colexp_scale);
                }
                cptr += p.ldc;
              }
              if (p.colexp_sum != nullptr && nvalid && !p.atomic) atomicAdd(p.colexp_sum + n, es);
            }
            __syncwarp();
          };
          // two register sets: the TMEM load of the next chunk is in flight while this one is staged / stored
          uint32_t ra[32], rb[32];
          int ch = ehalf;
          bool have = ch < nchunks && t.n0 + ch * 32 < p.N;   // warp-uniform
          if (have) tmem_ld_32x32(tsub + ch * 32, ra);
          while (have) {
            tmem_ld_wait();
            const int ch2 = ch + 2;
            const bool have2 = ch2 < nchunks && t.n0 + ch2 * 32 < p.N;
            if (have2) tmem_ld_32x32(tsub + ch2 * 32, rb);
            process(ra, ch);
            if (!have2) break;
            tmem_ld_wait();
            ch = ch2 + 2;
            have = ch < nchunks && t.n0 + ch * 32 < p.N;
            if (have) tmem_ld_32x32(tsub + ch * 32, ra);
            process(rb, ch2);
          }
        }
      } else {
        const int ph = t.phase;
        const int per_img = p.th * p.tw;
        const int bi = row / per_img;
        const int rem = row - bi * per_img;
        const int iy = rem / p.tw, ix = rem - iy * p.tw;
        const int b = t.b0 + bi, i = t.y0 + iy, j = t.x0 + ix;
        const bool valid = (b < p.B) && (i < p.phase_eh[ph]) && (j < p.phase_ew[ph]);
        const int oy = p.upsample ? 2 * i + p.phase_a[ph] : i;
        const int ox = p.upsample ? 2 * j + p.phase_b[ph] : j;
        const long long pix = valid ? ((long long)b * p.Ho + oy) * p.Wo + ox : 0;
        float nz = 0.f;
        if (valid && p.noise != nullptr)
          nz = __ldg(p.noise_strength) * __ldg(p.noise + (long long)b * p.noise_bstride + (long long)oy * p.Wo + ox);
        for (int ch = ehalf; ch < nchunks; ch += 2) {
          const int co = t.n0 + ch * 32;
          if (co >= p.Cout) break;
          uint32_t r[32];
          tmem_ld_32x32(taddr0 + ch * 32, r);
          tmem_ld_wait();
          if (!valid) continue;
          float v[32];
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __uint_as_float(r[e]);
          const int nq = min(8, (p.Cout - co) >> 2);   // valid float4 groups of this chunk (warp-uniform)
          if (p.split_k > 1) {
            // this split's partial sums into its own slice of the workspace (plain stores: every split covers every
            // pixel); conv_finish_kernel adds the slices in split order and applies the epilogue - deterministic
            float4* dstp = reinterpret_cast<float4*>(p.ws + (long long)t.split * p.ws_stride + pix * p.Cout + co);
#pragma unroll
            for (int e = 0; e < 8; ++e)
              if (e < nq) dstp[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
            continue;
          }
          if (p.demod != nullptr) {
            const float4* dm = reinterpret_cast<const float4*>(p.demod + (long long)b * p.Cout + co);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (e < nq) {
                const float4 d4 = __ldg(dm + e);
                v[4 * e] *= d4.x; v[4 * e + 1] *= d4.y; v[4 * e + 2] *= d4.z; v[4 * e + 3] *= d4.w;
              }
            }
          }
          if (p.noise != nullptr) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] += nz;
          }
          if (p.bias != nullptr) {
            const float4* bs = reinterpret_cast<const float4*>(p.bias + co);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (e < nq) {
                const float4 b4 = __ldg(bs + e);
                v[4 * e] += b4.x; v[4 * e + 1] += b4.y; v[4 * e + 2] += b4.z; v[4 * e + 3] += b4.w;
              }
            }
          }
          if (p.act == 1) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = lrelu_sqrt2(v[e]);
          } else if (p.act == 2) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = v[e] > 0.f ? v[e] : v[e] * 0.2f;
          }
          float4* dst = reinterpret_cast<float4*>(p.out + pix * p.Cout + co);
#pragma unroll
          for (int e = 0; e < 8; ++e)
            if (e < nq) dst[e] = make_float4(v[4 * e], v[4 * e + 1], v[4 * e + 2], v[4 * e + 3]);
          if (p.next_style != nullptr) {
            const float4* st = reinterpret_cast<const float4*>(p.next_style + (long long)b * p.Cout + co);
            uint2* dh = reinterpret_cast<uint2*>(p.next_hi + pix * p.next_ld + co);
            uint2* dl = (p.next_lo != nullptr) ? reinterpret_cast<uint2*>(p.next_lo + pix * p.next_ld + co) : nullptr;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              if (e < nq) {
                const float4 s0 = __ldg(st + e);
                uint2 h0, l0;
                gx_split4(make_float4(v[4 * e] * s0.x, v[4 * e + 1] * s0.y, v[4 * e + 2] * s0.z, v[4 * e + 3] * s0.w),
                          h0, l0);
                dh[e] = h0;
                if (dl != nullptr) dl[e] = l0;
              }
            }
          }
        }
      }
      // release the accumulator stage
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (crank == 0) mbar_arrive(&tempty_bar[acc]);
        else mbar_arrive_cluster(cluster_addr_of(&tempty_bar[acc], 0));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) cluster_sync_all();   // no CTA leaves while its peer can still signal its barriers
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// Grow-only device workspace of the split-K convs (a few MB: only layers of at most ~64 pixel tiles split).  Allocated
// outside stream capture only; under capture a too small workspace means "no split" (captured graphs are built after
// an eager pass of the same shapes, which sizes it).
float* conv_workspace(size_t bytes, cudaStream_t st) {
  static float* ws = nullptr;
  static size_t ws_bytes = 0;
  if (bytes <= ws_bytes) return ws;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return nullptr;
  }
  // a smaller buffer is NOT freed: a kernel in flight or a captured graph may still refer to it (a few MB, and the
  // sizes grow geometrically).  All split-K convs of a process share the workspace: they must be stream-ordered, as the
  // engine's are (GX_CONV_SPLITK=0 switches the split off).
  float* fresh = nullptr;
  const size_t want = 2 * bytes;
  if (cudaMalloc(&fresh, want) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  ws = fresh;
  ws_bytes = want;
  return ws;
}

// Second stage of a split-K conv: the epilogue of the conv branch above (same order of operations) applied in place to
// the summed partial results, 4 channels per thread.
__global__ void conv_finish_kernel(const float* __restrict__ ws, int splits, float* __restrict__ out,
                                   const float* __restrict__ demod,
                                   const float* __restrict__ noise, long long noise_bstride,
                                   const float* __restrict__ noise_strength, const float* __restrict__ bias, int act,
                                   const float* __restrict__ next_style, __nv_bfloat16* __restrict__ next_hi,
                                   __nv_bfloat16* __restrict__ next_lo, int next_ld, int hw, int cout, long long n4) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int cq = cout >> 2;
  const long long pix = i / cq;
  const int co = (int)(i - pix * cq) * 4;
  const int b = (int)(pix / hw);
  float4 v = __ldg(reinterpret_cast<const float4*>(ws) + i);
  for (int sp = 1; sp < splits; ++sp) {             // fixed order: run-to-run identical results
    const float4 u = __ldg(reinterpret_cast<const float4*>(ws) + (long long)sp * n4 + i);
    v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
  }
  if (demod != nullptr) {
    const float4 d4 = __ldg(reinterpret_cast<const float4*>(demod + (long long)b * cout + co));
    v.x *= d4.x; v.y *= d4.y; v.z *= d4.z; v.w *= d4.w;
  }
  if (noise != nullptr) {
    const float nz = __ldg(noise_strength) * __ldg(noise + (long long)b * noise_bstride + (pix - (long long)b * hw));
    v.x += nz; v.y += nz; v.z += nz; v.w += nz;
  }
  if (bias != nullptr) {
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + co));
    v.x += b4.x; v.y += b4.y; v.z += b4.z; v.w += b4.w;
  }
  if (act == 1) {
    v.x = lrelu_sqrt2(v.x); v.y = lrelu_sqrt2(v.y); v.z = lrelu_sqrt2(v.z); v.w = lrelu_sqrt2(v.w);
  } else if (act == 2) {
    v.x = v.x > 0.f ? v.x : v.x * 0.2f; v.y = v.y > 0.f ? v.y : v.y * 0.2f;
    v.z = v.z > 0.f ? v.z : v.z * 0.2f; v.w = v.w > 0.f ? v.w : v.w * 0.2f;
  }
  reinterpret_cast<float4*>(out)[i] = v;
  if (next_style != nullptr) {
    const float4 s0 = __ldg(reinterpret_cast<const float4*>(next_style + (long long)b * cout + co));
    uint2 h0, l0;
    gx_split4(make_float4(v.x * s0.x, v.y * s0.y, v.z * s0.z, v.w * s0.w), h0, l0);
    *reinterpret_cast<uint2*>(next_hi + pix * next_ld + co) = h0;
    if (next_lo != nullptr) *reinterpret_cast<uint2*>(next_lo + pix * next_ld + co) = l0;
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(ptr);
  return fn;
}

// bf16 tensor, dims listed innermost first; strides in elements for dims 1..rank-1
int make_tmap(CUtensorMap* m, const void* base, int rank, const unsigned long long* dims,
              const unsigned long long* strides_elems, const unsigned* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return GX_ERR_CUDA;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_elems[i] * 2ull;
  if (reinterpret_cast<uintptr_t>(base) & 15) return GX_ERR_ARG;
  for (int i = 0; i + 1 < rank; ++i)
    if (gstr[i] & 15) return GX_ERR_ARG;
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    gx_set_last_cuda_error(1000 + (int)r);
    return GX_ERR_CUDA;
  }
  return GX_OK;
}

int pick_stages(int passes, int block_n, int mtiles, int want, int nsub = 1) {
  const int nplanes = passes == 3 ? 2 : 1;
  const int stage_bytes = nplanes * (A_PLANE_BYTES * mtiles + nsub * block_n * BK * 2);
  const int avail = SMEM_LIMIT - 1024 /*align slack*/ - 256 /*barriers*/ - STAGING_BYTES;
  int s = avail / stage_bytes;
  if (s > MAX_STAGES) s = MAX_STAGES;
  if (want > 0 && want < s) s = want;
  return s;
}

int launch(const UmmaParams& p_in, const CUtensorMap* maps, int total_work, cudaStream_t st) {
  UmmaParams p = p_in;
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("GX_UMMA_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    p.debug = dbg;
  }
  const int nplanes = p.passes == 3 ? 2 : 1;
  const int stage_bytes = nplanes * (A_PLANE_BYTES * p.mtiles + p.nsub * (p.pair ? p.block_n / 2 : p.block_n) * BK * 2);
  const int smem_bytes = p.stages * stage_bytes + 256 + STAGING_BYTES + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    GX_CHECK_CUDA(cudaFuncSetAttribute(gx_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    GX_CHECK_CUDA(cudaFuncSetAttribute(gx_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_set = true;
  }
  int grid = gx_umma_cta_budget();
  if (p.pair) {
    // total_work counts PAIRS of tiles; one 2-CTA cluster per unit of work in flight
    grid &= ~1;
    if (grid > 2 * total_work) grid = 2 * total_work;
    if (grid < 2) return GX_OK;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NTHREADS);
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    GX_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gx_umma_kernel<true>, p, maps[0], maps[1], maps[2], maps[3], total_work));
    GX_LAUNCH_CHECK();
    return GX_OK;
  }
  if (grid > total_work) grid = total_work;
  if (grid < 1) return GX_OK;
  gx_umma_kernel<false><<<grid, NTHREADS, smem_bytes, st>>>(p, maps[0], maps[1], maps[2], maps[3], total_work);
  GX_LAUNCH_CHECK();
  return GX_OK;
}

}  // namespace

extern "C" int gx_gemm(const gx_gemm_desc* d, void* stream) {
  GX_CHECK_ARG(d != nullptr && d->a_hi && d->b_hi && d->c);
  GX_CHECK_ARG(d->m > 0 && d->n > 0 && d->k > 0);
  GX_CHECK_ARG(d->passes == 1 || d->passes == 3);
  GX_CHECK_ARG(d->passes == 1 || (d->a_lo && d->b_lo));
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.mode = 0;
  p.passes = d->passes;
  p.f16 = d->ab_f16 ? 1 : 0;
  GX_CHECK_ARG(!p.f16 || d->passes == 1);
  p.a_mn = d->a_mn_major ? 1 : 0;
  p.b_mn = d->b_mn_major ? 1 : 0;
  p.M = d->m; p.N = d->n; p.K = d->k;
  int bn = d->block_n;
  if (bn == 0) bn = (d->n > 128) ? 256 : 128;
  GX_CHECK_ARG(bn == 64 || bn == 128 || bn == 256);
  p.block_n = bn;
  // 256-row tiles for single-pass GEMMs: one B tile in smem feeds two M=128 MMAs, which keeps the
  // smem fill rate (L2 -> SM) at the level of the 3-pass mode instead of 1.5x above it
  // (MN-major B is loaded in 64-wide boxes: each CTA's half must hold at least one)
  p.pair = (d->cluster_pair && d->m > BM && bn >= (p.b_mn ? 128 : 64)) ? 1 : 0;
  p.mtiles = (p.passes == 1 && bn == 256 && d->m > BM && !d->force_m128 && !p.pair) ? 2 : 1;
  // 256 x 512 pair tiles for the single-pass GEMMs with a wide N and a long K loop (dZn = dS Wk, gWk = dS^T Zn):
  // these are bound by operand ingest through the SM<->L2 port, and the A tile then feeds two N = 256 MMAs.
  // Not for the score GEMM (short K: the store-heavy epilogue needs the second accumulator stage to overlap).
  static const int nsub_env = getenv("GX_UMMA_NSUB") ? atoi(getenv("GX_UMMA_NSUB")) : 2;
  p.nsub = (p.pair && p.passes == 1 && bn == 256 && d->n % 512 == 0 && d->colexp_sum == nullptr && d->k >= 2048 &&
            nsub_env == 2) ? 2 : 1;
  p.acc_stages = (p.mtiles * bn * p.nsub <= 256) ? 2 : 1;
  p.stages = pick_stages(p.passes, p.pair ? bn / 2 : bn, p.mtiles, d->stages, p.nsub);
  GX_CHECK_ARG(p.stages >= 2);
  p.tiles_m = gx_cdiv(d->m, BM * p.mtiles);
  p.tiles_n = gx_cdiv(d->n, bn * p.nsub);
  p.kiters_total = gx_cdiv(d->k, BK);
  int sk = d->split_k < 1 ? 1 : d->split_k;
  if (sk > p.kiters_total) sk = p.kiters_total;
  // every split must own at least one k-iteration
  while (sk > 1 && (long long)(sk - 1) * gx_cdiv(p.kiters_total, sk) >= p.kiters_total) --sk;
  p.split_k = sk;
  p.atomic = (sk > 1 || d->accumulate) ? 1 : 0;
  p.c = d->c; p.ldc = d->ldc; p.bias = d->bias;
  p.colexp_sum = d->colexp_sum;
  p.colexp_scale = d->colexp_scale;
  GX_CHECK_ARG(d->colexp_sum == nullptr || !p.atomic);

  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  const void* aptr[2] = {d->a_hi, d->a_lo};
  const void* bptr[2] = {d->b_hi, d->b_lo};
  const int nplanes = p.passes == 3 ? 2 : 1;
  for (int pl = 0; pl < nplanes; ++pl) {
    int rc;
    if (!p.a_mn) {
      unsigned long long dims[2] = {(unsigned long long)d->k, (unsigned long long)d->m};
      unsigned long long str[1] = {(unsigned long long)d->lda};
      unsigned box[2] = {BK, BM};
      rc = make_tmap(&maps[pl], aptr[pl], 2, dims, str, box);
    } else {
      unsigned long long dims[2] = {(unsigned long long)d->m, (unsigned long long)d->k};
      unsigned long long str[1] = {(unsigned long long)d->lda};
      unsigned box[2] = {64, BK};
      rc = make_tmap(&maps[pl], aptr[pl], 2, dims, str, box);
    }
    if (rc != GX_OK) return rc;
    if (!p.b_mn) {
      unsigned long long dims[2] = {(unsigned long long)d->k, (unsigned long long)d->n};
      unsigned long long str[1] = {(unsigned long long)d->ldb};
      unsigned box[2] = {BK, (unsigned)(p.pair ? bn / 2 : bn)};
      rc = make_tmap(&maps[2 + pl], bptr[pl], 2, dims, str, box);
    } else {
      unsigned long long dims[2] = {(unsigned long long)d->n, (unsigned long long)d->k};
      unsigned long long str[1] = {(unsigned long long)d->ldb};
      unsigned box[2] = {64, BK};
      rc = make_tmap(&maps[2 + pl], bptr[pl], 2, dims, str, box);
    }
    if (rc != GX_OK) return rc;
  }
  if (nplanes == 1) {
    maps[1] = maps[0];
    maps[3] = maps[2];
  }
  const int total = (p.pair ? (p.tiles_m + 1) / 2 : p.tiles_m) * p.tiles_n * p.split_k;
  return launch(p, maps, total, (cudaStream_t)stream);
}

extern "C" int gx_modconv(const gx_conv_desc* d, void* stream) {
  GX_CHECK_ARG(d != nullptr && d->x_hi && d->w_hi && d->out);
  GX_CHECK_ARG(d->passes == 1 || d->passes == 3);
  GX_CHECK_ARG(d->passes == 1 || (d->x_lo && d->w_lo));
  GX_CHECK_ARG(d->batch > 0 && d->h > 0 && d->w > 0);
  const int cin_ld = d->cin_ld > 0 ? d->cin_ld : d->cin;
  GX_CHECK_ARG(cin_ld % BK == 0 && cin_ld >= d->cin && d->cout % 4 == 0);
  GX_CHECK_ARG(d->next_style == nullptr || d->next_hi != nullptr);
  UmmaParams p;
  memset(&p, 0, sizeof(p));
  p.mode = 1;
  p.nsub = 1;
  p.passes = d->passes;
  p.B = d->batch; p.H = d->h; p.W = d->w; p.Cin = cin_ld; p.Cout = d->cout;  // the K loop runs over padded channels
  p.upsample = d->upsample ? 1 : 0;
  const int dil = d->dilation > 0 ? d->dilation : 1;
  GX_CHECK_ARG(dil <= 64 && (dil == 1 || !p.upsample));
  int bn = d->block_n;
  if (bn == 0) bn = (d->cout >= 256) ? 256 : 128;
  if (bn > d->cout) bn = (d->cout >= 128) ? 128 : 64;
  GX_CHECK_ARG(bn == 64 || bn == 128 || bn == 256);
  p.block_n = bn;
  p.mtiles = 1;
  p.acc_stages = 2;
  p.stages = pick_stages(p.passes, bn, 1, d->stages);
  GX_CHECK_ARG(p.stages >= 2);
  // pixel tile: 128 rows = nb images x th x tw
  const int ext_w = p.upsample ? d->w + 1 : d->w;
  const int ext_h = p.upsample ? d->h + 1 : d->h;
  int tw = 16, th = 8, nb = 1;
  if (ext_w <= 4 && ext_h <= 4) { tw = 4; th = 4; nb = 8; }
  else if (ext_w <= 8 && ext_h <= 8) { tw = 8; th = 8; nb = 2; }
  else if (ext_w <= 8) { tw = 8; th = 16; nb = 1; }
  p.tw = tw; p.th = th; p.nb = nb;
  p.Ho = p.upsample ? 2 * d->h + 1 : d->h;
  p.Wo = p.upsample ? 2 * d->w + 1 : d->w;
  const int groups = gx_cdiv(d->batch, nb);
  if (!p.upsample) {
    p.nphases = 1;
    p.phase_eh[0] = d->h; p.phase_ew[0] = d->w;
    p.phase_ty[0] = gx_cdiv(d->h, th); p.phase_tx[0] = gx_cdiv(d->w, tw);
    p.ntaps[0] = 9;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int tpi = ky * 3 + kx;
        p.tap_dy[0][tpi] = (signed char)((ky - 1) * dil);
        p.tap_dx[0][tpi] = (signed char)((kx - 1) * dil);
        p.tap_w[0][tpi] = (signed char)tpi;
      }
    p.phase_tile_start[0] = 0;
    p.phase_tile_start[1] = groups * p.phase_ty[0] * p.phase_tx[0];
  } else {
    // out[2i+a, 2j+b] = sum over taps (ky,kx) with ky%2==a, kx%2==b of x[i-ky/2, j-kx/2] * w[ky,kx]
    p.nphases = 4;
    int start = 0;
    for (int ph = 0; ph < 4; ++ph) {
      const int a = ph >> 1, b = ph & 1;
      p.phase_a[ph] = a; p.phase_b[ph] = b;
      p.phase_eh[ph] = a ? d->h : d->h + 1;
      p.phase_ew[ph] = b ? d->w : d->w + 1;
      p.phase_ty[ph] = gx_cdiv(p.phase_eh[ph], th);
      p.phase_tx[ph] = gx_cdiv(p.phase_ew[ph], tw);
      int nt = 0;
      for (int ky = a; ky < 3; ky += 2)
        for (int kx = b; kx < 3; kx += 2) {
          p.tap_dy[ph][nt] = (signed char)(-(ky / 2));
          p.tap_dx[ph][nt] = (signed char)(-(kx / 2));
          p.tap_w[ph][nt] = (signed char)(ky * 3 + kx);
          ++nt;
        }
      p.ntaps[ph] = nt;
      p.phase_tile_start[ph] = start;
      start += groups * p.phase_ty[ph] * p.phase_tx[ph];
    }
    p.phase_tile_start[4] = start;
  }
  p.tiles_m = p.phase_tile_start[p.nphases];
  // CTA pairs (two adjacent pixel tiles of one phase on two SMs, half of the weight tile each): every
  // phase gets an even number of tiles; the padding tile decodes to an image group past the batch
  // (TMA zero fill, no stores)
  p.pair = (d->cluster_pair && p.tiles_m >= 16) ? 1 : 0;
  if (p.pair) {
    int orig[5];
    for (int ph = 0; ph <= p.nphases; ++ph) orig[ph] = p.phase_tile_start[ph];
    int start = 0;
    for (int ph = 0; ph < p.nphases; ++ph) {
      p.phase_tile_start[ph] = start;
      start += (orig[ph + 1] - orig[ph] + 1) & ~1;
    }
    p.phase_tile_start[p.nphases] = start;
    p.tiles_m = start;
  }
  // few pixel tiles (4x4 ... 16x16 layers): narrower channel tiles put more SMs on the layer, which is
  // bound by the serial K loop of a handful of CTAs otherwise
  p.split_k = 1;
  {
    // 4x4 / 8x8 layers: a handful of CTAs each walking the whole tap x channel loop (0.1 ms per layer whatever its
    // size).  Split that loop over otherwise idle SMs (full-width channel tiles, >= 4 splits): every split stores its
    // partial sums into its own slice of a workspace and a small second kernel adds the slices in a fixed order,
    // applies demodulation / noise / bias / activation and emits the next layer's planes.
    static const bool no_split = getenv("GX_CONV_SPLITK") != nullptr && atoi(getenv("GX_CONV_SPLITK")) == 0;
    const int units = (p.pair ? p.tiles_m / 2 : p.tiles_m) * gx_cdiv(d->cout, bn);
    const int cblocks = cin_ld / BK;
    int sk = (no_split || d->block_n != 0) ? 1 : gx_sm_count() / (units * (p.pair ? 2 : 1));
    if (sk > 16) sk = 16;
    auto all_nonempty = [&](int s_) {
      for (int ph = 0; ph < p.nphases; ++ph) {
        const int ktot = p.ntaps[ph] * cblocks;
        if ((s_ - 1) * gx_cdiv(ktot, s_) >= ktot) return false;
      }
      return true;
    };
    while (sk > 1 && !all_nonempty(sk)) --sk;
    // (two splits pay for the second kernel on a plain conv - 72 k-blocks per tile - but not on the sub-pixel phases
    //  of a transposed conv, whose longest phase has 32)
    if (sk >= (p.upsample ? 4 : 2)) {
      p.ws_stride = (long long)d->batch * p.Ho * p.Wo * d->cout;
      p.ws = conv_workspace((size_t)sk * (size_t)p.ws_stride * sizeof(float), (cudaStream_t)stream);
      if (p.ws != nullptr) p.split_k = sk;      // no workspace (allocation refused, e.g. under stream capture): no split
    }
  }
  if (d->block_n == 0 && p.split_k == 1) {
    while (bn > 64 && p.tiles_m * gx_cdiv(d->cout, bn) * 2 <= gx_sm_count()) bn >>= 1;
    p.block_n = bn;
  }
  p.stages = pick_stages(p.passes, p.pair ? bn / 2 : bn, 1, d->stages);
  GX_CHECK_ARG(p.stages >= 2);
  p.tiles_n = gx_cdiv(d->cout, bn);
  p.demod = d->demod; p.noise = d->noise; p.noise_bstride = d->noise_batch_stride;
  p.noise_strength = d->noise_strength; p.bias = d->bias; p.act = d->act;
  GX_CHECK_ARG(d->noise == nullptr || d->noise_strength != nullptr);
  p.out = d->out; p.next_style = d->next_style;
  p.next_hi = reinterpret_cast<__nv_bfloat16*>(d->next_hi);
  p.next_lo = reinterpret_cast<__nv_bfloat16*>(d->next_lo);
  p.next_ld = d->next_ld > 0 ? d->next_ld : d->cout;
  GX_CHECK_ARG(p.next_ld >= d->cout && p.next_ld % 4 == 0);

  CUtensorMap maps[4];
  memset(maps, 0, sizeof(maps));
  const void* xptr[2] = {d->x_hi, d->x_lo};
  const void* wptr[2] = {d->w_hi, d->w_lo};
  const int nplanes = p.passes == 3 ? 2 : 1;
  for (int pl = 0; pl < nplanes; ++pl) {
    unsigned long long dims[4] = {(unsigned long long)cin_ld, (unsigned long long)d->w, (unsigned long long)d->h,
                                  (unsigned long long)d->batch};
    unsigned long long str[3] = {(unsigned long long)cin_ld, (unsigned long long)cin_ld * d->w,
                                 (unsigned long long)cin_ld * d->w * d->h};
    unsigned box[4] = {BK, (unsigned)tw, (unsigned)th, (unsigned)nb};
    // a box may not exceed the tensor extent in TMA encoding only through its
    // dimension limit of 256; partial boxes are zero-filled
    int rc = make_tmap(&maps[pl], xptr[pl], 4, dims, str, box);
    if (rc != GX_OK) return rc;
    unsigned long long wd[2] = {(unsigned long long)9 * cin_ld, (unsigned long long)d->cout};
    unsigned long long ws[1] = {(unsigned long long)9 * cin_ld};
    unsigned wb[2] = {BK, (unsigned)(p.pair ? bn / 2 : bn)};
    rc = make_tmap(&maps[2 + pl], wptr[pl], 2, wd, ws, wb);
    if (rc != GX_OK) return rc;
  }
  if (nplanes == 1) {
    maps[1] = maps[0];
    maps[3] = maps[2];
  }
  const int total = (p.pair ? p.tiles_m / 2 : p.tiles_m) * p.tiles_n * p.split_k;
  if (p.split_k > 1) {
    cudaStream_t st = (cudaStream_t)stream;
    const int rc = launch(p, maps, total, st);
    if (rc != GX_OK) return rc;
    const long long n4 = p.ws_stride / 4;
    conv_finish_kernel<<<gx_cdiv(n4, 256), 256, 0, st>>>(p.ws, p.split_k, p.out, p.demod, p.noise, p.noise_bstride, p.noise_strength,
                                                         p.bias, p.act, p.next_style, p.next_hi, p.next_lo, p.next_ld,
                                                         p.Ho * p.Wo, d->cout, n4);
    GX_LAUNCH_CHECK();
    return GX_OK;
  }
  return launch(p, maps, total, (cudaStream_t)stream);
}

// --------------------------------------------------------------------------
// fp32 SIMT cross-check GEMM on the same planes (tests only)
// --------------------------------------------------------------------------
namespace {
__device__ __forceinline__ float plane_val(const __nv_bfloat16* hi, const __nv_bfloat16* lo, long long idx,
                                           int f16 = 0) {
  if (f16) return __half2float(reinterpret_cast<const __half*>(hi)[idx]);
  float v = __bfloat162float(hi[idx]);
  if (lo) v += __bfloat162float(lo[idx]);
  return v;
}
__global__ void gemm_check_kernel(const __nv_bfloat16* ah, const __nv_bfloat16* al, const __nv_bfloat16* bh,
                                  const __nv_bfloat16* bl, long long lda, long long ldb, int a_mn, int b_mn, int M,
                                  int N, int K, float* c, long long ldc, const float* bias, int f16) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) {
    const long long ia = a_mn ? (long long)k * lda + m : (long long)m * lda + k;
    const long long ib = b_mn ? (long long)k * ldb + n : (long long)n * ldb + k;
    acc = fmaf(plane_val(ah, al, ia, f16), plane_val(bh, bl, ib, f16), acc);
  }
  if (bias) acc += bias[n];
  c[(long long)m * ldc + n] = acc;
}
}  // namespace

extern "C" int gx_gemm_check(const gx_gemm_desc* d, void* stream) {
  GX_CHECK_ARG(d != nullptr && d->a_hi && d->b_hi && d->c);
  dim3 grid(gx_cdiv(d->n, 128), d->m);
  gemm_check_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)d->a_hi, d->passes == 3 ? (const __nv_bfloat16*)d->a_lo : nullptr,
      (const __nv_bfloat16*)d->b_hi, d->passes == 3 ? (const __nv_bfloat16*)d->b_lo : nullptr, d->lda, d->ldb,
      d->a_mn_major, d->b_mn_major, d->m, d->n, d->k, d->c, d->ldc, d->bias, d->ab_f16);
  GX_LAUNCH_CHECK();
  return GX_OK;
}
