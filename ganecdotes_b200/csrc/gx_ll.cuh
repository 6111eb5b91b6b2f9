// Low-latency K-vector exchange over NVLink peer memory (see gx_ll_desc in include/ganecdotes_b200.h).
// A word is {fp32 value, u32 tag}; 8-byte aligned 8-byte stores / loads are single-copy atomic, so a consumer
// that sees the expected tag also sees the value written with it: no fence, no separate flag, one NVLink trip.
#pragma once
#include "gx_common.cuh"

#ifdef __CUDACC__
namespace gxll {

constexpr long long kTimeoutCycles = 60000000000LL;   // ~30 s at 1.9 GHz: a peer that never arrives is an error

__device__ __forceinline__ void st_word(void* p, float v, unsigned seq) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(seq) : "memory");
}
__device__ __forceinline__ uint2 ld_word(const void* p) {
  uint2 v;
  asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint4 ld_word2(const void* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p)
               : "memory");
  return v;
}

// slot of `from_rank` in the block, as seen in the buffer of `in_rank`
__device__ __forceinline__ uint2* slot(const gx_ll_desc& d, int in_rank, int from_rank, int k) {
  return reinterpret_cast<uint2*>(d.peers[in_rank]) + d.block_words + (long long)from_rank * k;
}

// push v (column col) of this rank to every rank's buffer
__device__ __forceinline__ void send1(const gx_ll_desc& d, int k, int col, float v) {
  for (int r = 0; r < d.world; ++r) st_word(slot(d, r, d.rank, k) + col, v, d.seq);
}

// sum over ranks (rank order) of column col; spins until every tag matches
__device__ __forceinline__ float recv1(const gx_ll_desc& d, int k, int col) {
  float t = 0.f;
  for (int r = 0; r < d.world; ++r) {
    const uint2* p = slot(d, d.rank, r, k) + col;
    uint2 w = ld_word(p);
    if (w.y != d.seq) {
      const long long t0 = clock64();
      do {
        w = ld_word(p);
        if (clock64() - t0 > kTimeoutCycles) {
          if (d.err) atomicExch(d.err, 1);
          break;
        }
      } while (w.y != d.seq);
    }
    t += __uint_as_float(w.x);
  }
  return t;
}

// four consecutive columns (col % 4 == 0; the block is 32-byte aligned for k % 4 == 0)
__device__ __forceinline__ void recv4(const gx_ll_desc& d, int k, int col, float (&out)[4]) {
  out[0] = out[1] = out[2] = out[3] = 0.f;
  for (int r = 0; r < d.world; ++r) {
    const uint2* p = slot(d, d.rank, r, k) + col;
    uint4 a = ld_word2(p), b = ld_word2(p + 2);
    if (a.y != d.seq || a.w != d.seq || b.y != d.seq || b.w != d.seq) {
      const long long t0 = clock64();
      do {
        a = ld_word2(p);
        b = ld_word2(p + 2);
        if (clock64() - t0 > kTimeoutCycles) {
          if (d.err) atomicExch(d.err, 1);
          break;
        }
      } while (a.y != d.seq || a.w != d.seq || b.y != d.seq || b.w != d.seq);
    }
    out[0] += __uint_as_float(a.x);
    out[1] += __uint_as_float(a.z);
    out[2] += __uint_as_float(b.x);
    out[3] += __uint_as_float(b.z);
  }
}

}  // namespace gxll
#endif
