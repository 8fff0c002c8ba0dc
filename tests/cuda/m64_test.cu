// Where does tcgen05.mma put a 64-row accumulator (M = 64, cta_group::1, kind::tf32)?  (run on the B200 box)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/m64_test.bin tests/cuda/m64_test.cu && /tmp/m64_test.bin
// D[i][n] = (i + 1) * (n + 1) for i < 64, n < N from one k-step; every TMEM lane is dumped and decoded back to (row, col).
#include <cstdio>
#include <vector>
#include "../../mfas_b200/csrc/umma.cuh"
using namespace umma;

template <int N>
__global__ void __launch_bounds__(128) k_m64(float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  uint8_t* a = smem;            // 64 rows x 128 B (K-major, SW128): A[i][0] = i + 1
  uint8_t* b = smem + 16384;    // N rows x 128 B: B[n][0] = n + 1
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, 32);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  for (int i = tid; i < 16384 / 4 + 4096 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.f;
  __syncthreads();
  if (tid < 64) *reinterpret_cast<float*>(a + sw128(tid, 0)) = (float)(tid + 1);
  if (tid < N) *reinterpret_cast<float*>(b + sw128(tid, 0)) = (float)(tid + 1);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  // clear all 128 lanes x 32 columns first (an M = 128 product of zeros), then the M = 64 product
  if (tid == 0) {
    const uint64_t dz = smem_desc(smem_u32(smem) + 8192 + 64, 16, 1024);      // a region of zeros
    mma_tf32(tm, dz, dz, idesc_tf32(128, 32, false, false), 0u);
    mma_tf32(tm, smem_desc(smem_u32(a), 16, 1024), smem_desc(smem_u32(b), 16, 1024), idesc_tf32(64, N, false, false), 0u);
    mma_commit(&bar);
  }
  cta_wait(&bar, 0, (int*)&tmem_base + 0 == nullptr ? nullptr : (int*)smem);   // (ok flag into scratch)
  tc_fence_after();
  float v[32];
  tmem_ld32(tm + ((uint32_t)(warp * 32) << 16), v);
  for (int c = 0; c < 32; ++c) out[tid * 32 + c] = v[c];
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tm, 32);
}

template <int N> int run() {
  float* d;
  cudaMalloc(&d, 128 * 32 * 4);
  cudaFuncSetAttribute(k_m64<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  k_m64<N><<<1, 128, 32768>>>(d);
  std::vector<float> h(128 * 32);
  cudaError_t e = cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) { printf("N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
  printf("M=64 N=%d: TMEM lane -> accumulator row (col 0 value - 1), '.' = untouched lane\n", N);
  int identity = 1, split = 1;
  for (int lane = 0; lane < 128; ++lane) {
    const float v0 = h[lane * 32];
    const int row = v0 > 0 ? (int)(v0 + 0.5f) - 1 : -1;
    if (lane % 32 == 0) printf("  lanes %3d..: ", lane);
    if (row < 0) printf("  ."); else printf("%3d", row);
    if (lane % 32 == 31) printf("\n");
    if (row != (lane < 64 ? lane : -1)) identity = 0;
    if (row != ((lane % 32) < 16 ? (lane / 32) * 16 + lane % 32 : -1)) split = 0;
    if (row >= 0) for (int c = 0; c < N; ++c) if (h[lane * 32 + c] != (float)((row + 1) * (c + 1))) { printf("  bad value lane %d col %d: %g\n", lane, c, h[lane * 32 + c]); return 1; }
  }
  printf("  layout: %s\n", identity ? "rows 0..63 in lanes 0..63" : split ? "16 rows in the first 16 lanes of every 32-lane quarter" : "other (see map)");
  return 0;
}
int main() { return run<16>() | run<32>(); }
