// Micro-benchmark: how fast can B200 stream a row-major [R][K] fp32 matrix when every CTA touches a
// "slab" of 128 rows x SEG floats (the access shape of a GEMM operand tile / a fused-Adam weight tile)?
// Variants: read-only (forward W tiles) and read-modify-write of three arrays (Adam p/m/v).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tests/cuda/slab_bw.bin tests/cuda/slab_bw.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int SEG, bool RMW, int UNROLL>
__global__ void __launch_bounds__(256) k_slab(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, int K,
                                              long long ntiles, int segs_per_row, float* sink) {
  constexpr int F4_PER_ROW = SEG / 4;
  constexpr int F4_PER_TILE = 128 * F4_PER_ROW;
  constexpr int PER_THREAD = F4_PER_TILE / 256;
  static_assert(PER_THREAD % UNROLL == 0 || PER_THREAD < UNROLL, "");
  float acc = 0.f;
  for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const long long rb = t / segs_per_row;
    const int sg = (int)(t % segs_per_row);
    const long long base = rb * 128 * (long long)K + (long long)sg * SEG;
    constexpr int U = PER_THREAD < UNROLL ? PER_THREAD : UNROLL;
    for (int j0 = 0; j0 < PER_THREAD; j0 += U) {
      float4 a[U], b[U], c[U];
      long long off[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int i = threadIdx.x + 256 * (j0 + u);
        off[u] = base + (long long)(i / F4_PER_ROW) * K + (i % F4_PER_ROW) * 4;
        a[u] = *reinterpret_cast<const float4*>(p + off[u]);
        if (RMW) { b[u] = *reinterpret_cast<const float4*>(m + off[u]); c[u] = *reinterpret_cast<const float4*>(v + off[u]); }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (RMW) {
          a[u].x += 1e-3f * b[u].x; b[u].x = 0.9f * b[u].x + c[u].x; c[u].x *= 0.999f;
          a[u].y += 1e-3f * b[u].y; b[u].y = 0.9f * b[u].y + c[u].y; c[u].y *= 0.999f;
          a[u].z += 1e-3f * b[u].z; b[u].z = 0.9f * b[u].z + c[u].z; c[u].z *= 0.999f;
          a[u].w += 1e-3f * b[u].w; b[u].w = 0.9f * b[u].w + c[u].w; c[u].w *= 0.999f;
          *reinterpret_cast<float4*>(p + off[u]) = a[u];
          *reinterpret_cast<float4*>(m + off[u]) = b[u];
          *reinterpret_cast<float4*>(v + off[u]) = c[u];
        } else {
          acc += a[u].x + a[u].y + a[u].z + a[u].w;
        }
      }
    }
  }
  if (acc == 123.456f) *sink = acc;
}

template <int SEG, bool RMW, int UNROLL>
static void run(float* p, float* m, float* v, long long R, int K, float* sink, int ctas_per_sm) {
  const int segs = K / SEG;
  const long long ntiles = (R / 128) * segs;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = 148 * ctas_per_sm;
  k_slab<SEG, RMW, UNROLL><<<grid, 256>>>(p, m, v, K, ntiles, segs, sink);
  cudaEventRecord(e0);
  for (int i = 0; i < 3; ++i) k_slab<SEG, RMW, UNROLL><<<grid, 256>>>(p, m, v, K, ntiles, segs, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  ms /= 3;
  const double bytes = (double)R * segs * SEG * 4 * (RMW ? 6 : 1);
  printf("seg=%4d floats (%4d B/row)  %s unroll=%d ctas/sm=%d : %8.3f ms  %7.1f GB/s  (%s)\n", SEG, SEG * 4,
         RMW ? "rmw p/m/v" : "read     ", UNROLL, ctas_per_sm, ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int K = 2688;
  const long long R = 128LL * 400;            // 51200 rows x 2688 x 4 B = 550 MB per array
  float *p, *m, *v, *sink;
  cudaMalloc(&p, R * K * 4); cudaMalloc(&m, R * K * 4); cudaMalloc(&v, R * K * 4); cudaMalloc(&sink, 4);
  cudaMemset(p, 0, R * K * 4); cudaMemset(m, 0, R * K * 4); cudaMemset(v, 0, R * K * 4);
  for (int c : {2, 4, 8}) {
    run<32, false, 4>(p, m, v, R, K, sink, c);
    run<64, false, 4>(p, m, v, R, K, sink, c);
    run<128, false, 4>(p, m, v, R, K, sink, c);
    run<384, false, 4>(p, m, v, R, K, sink, c);
    run<2688, false, 4>(p, m, v, R, K, sink, c);
  }
  for (int c : {2, 4, 8}) {
    run<32, true, 4>(p, m, v, R, K, sink, c);
    run<64, true, 4>(p, m, v, R, K, sink, c);
    run<128, true, 4>(p, m, v, R, K, sink, c);
    run<384, true, 4>(p, m, v, R, K, sink, c);
    run<2688, true, 4>(p, m, v, R, K, sink, c);
  }
  run<128, true, 2>(p, m, v, R, K, sink, 4);
  run<128, true, 8>(p, m, v, R, K, sink, 4);
  run<128, false, 8>(p, m, v, R, K, sink, 4);
  return 0;
}
