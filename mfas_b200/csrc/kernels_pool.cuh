// Feature-cache builder: global average pooling of one backbone tap into its column slice of the cache
// (GlobalPooling2D.forward, /root/reference/models/auxiliary/aux_models.py:58-64; applied to every selected tap at
// /root/reference/models/search/ntu_searchable.py:224-225 -- here once per sample at cache-build time).
//
// in  : [B][C][S] fp32 contiguous (S = product of the trailing dims; S = 1 for a tap that is already a vector)
// out : out[b * out_ld + c] = (sum_s in[b][c][s]) / S      (a column slice of the [N, sum(D)] cache matrix)
//
// HBM-bound streaming reduction: one warp per (b, c) row, lanes walk the row in 16-byte words (4 independent loads in
// flight per lane, evict-first: every byte is read once), warp-shuffle tree, one 4-byte store per row.  Algorithmic bytes
// per row: 4 (S + 1).  Deterministic (fixed summation order).
#pragma once
#include "common.cuh"

namespace mfas {

constexpr int kPoolThreads = 256;

// G lanes share one row (G = 32: a warp per row; G = 8: four rows per warp, for short rows -- a 392-float row is only
// three 16-byte words per lane of a full warp, too few loads in flight to cover the HBM latency: r01H 0.52 of the roofline).
template <int G>
__global__ void __launch_bounds__(kPoolThreads)
k_global_pool(const float* __restrict__ in, long long rows, long long C, long long S, float* __restrict__ out, long long out_ld,
              int vec4) {
  constexpr int RPW = 32 / G;                            // rows per warp
  const int lane = threadIdx.x & 31, gl = lane % G, grp = lane / G;
  const long long nw = (long long)gridDim.x * (kPoolThreads / 32);
  // the row loop is warp-uniform (base), so every lane reaches the shuffles
  for (long long base = ((long long)blockIdx.x * (kPoolThreads / 32) + (threadIdx.x >> 5)) * RPW; base < rows; base += nw * RPW) {
    const long long r = base + grp;
    const bool valid = r < rows;
    const float* p = in + (valid ? r : 0) * S;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if (valid) {
      if (vec4) {                                        // S % 4 == 0 and a 16-byte aligned base: every row is aligned
        const float4* p4 = reinterpret_cast<const float4*>(p);
        const long long n4 = S >> 2;
        long long i = gl;
        for (; i + 3 * G < n4; i += 4 * G) {
          const float4 v0 = __ldcs(p4 + i), v1 = __ldcs(p4 + i + G), v2 = __ldcs(p4 + i + 2 * G), v3 = __ldcs(p4 + i + 3 * G);
          a0 += (v0.x + v0.y) + (v0.z + v0.w); a1 += (v1.x + v1.y) + (v1.z + v1.w);
          a2 += (v2.x + v2.y) + (v2.z + v2.w); a3 += (v3.x + v3.y) + (v3.z + v3.w);
        }
        for (; i < n4; i += G) { const float4 v = __ldcs(p4 + i); a0 += (v.x + v.y) + (v.z + v.w); }
      } else {
        long long i = gl;
        for (; i + 3 * G < S; i += 4 * G) { a0 += __ldcs(p + i); a1 += __ldcs(p + i + G); a2 += __ldcs(p + i + 2 * G); a3 += __ldcs(p + i + 3 * G); }
        for (; i < S; i += G) a0 += __ldcs(p + i);
      }
    }
    float s = (a0 + a1) + (a2 + a3);
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);      // o < G: stays inside the row's lanes
    if (valid && gl == 0) out[(r / C) * out_ld + (r % C)] = s / (float)S;
  }
}

}  // namespace mfas
