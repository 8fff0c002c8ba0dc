// Standalone check of the TMA primitives in mfas_b200/csrc/umma.cuh (run on the B200 box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o /tmp/tma_test.bin tests/cuda/tma_test.cu && /tmp/tma_test.bin
// (1) a 2-D tile {32 floats, 128 rows} of a row-major [H][K] matrix through a tensor map held in GLOBAL memory, SWIZZLE_128B:
//     must land exactly where umma::sw128(row, byte) puts it (the K-major operand layout of the tcgen05 descriptors), rows
//     beyond H as zeros;  (2) tile::gather4: four arbitrary rows of a [N][ld] matrix per instruction, 64 rows in 16 instructions,
//     into the same layout.
#include <cuda.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../mfas_b200/csrc/umma.cuh"

using namespace umma;

__global__ void __launch_bounds__(128) k_tile(const CUtensorMap* maps, int which, int c0, int c1, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    tmap_acquire(maps + which);
    mbar_arrive_expect_tx(&bar, 16384);
    tma_load_2d(smem_u32(smem), maps + which, c0, c1, &bar, pol);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

__global__ void __launch_bounds__(128) k_gather(const __grid_constant__ CUtensorMap map, const int* rows, int c0, float* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = align1024(smem_raw);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x < 16) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
    if (threadIdx.x == 0) mbar_arrive_expect_tx(&bar, 64 * 128);
    __syncwarp(0xFFFF);
    const int j = threadIdx.x;
    tma_gather4(smem_u32(smem) + j * 512, &map, c0, rows[4 * j], rows[4 * j + 1], rows[4 * j + 2], rows[4 * j + 3], &bar, pol);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < 2048; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q) != cudaSuccess || !encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
  const int H = 80, K = 416, N = 300, LD = 1920;
  std::vector<float> W((size_t)H * K), X((size_t)N * LD);
  for (size_t i = 0; i < W.size(); ++i) W[i] = (float)(i % 9973) * 0.25f + 1.f;
  for (size_t i = 0; i < X.size(); ++i) X[i] = (float)(i % 7919) * 0.5f - 3.f;
  float *dW, *dX, *dout;
  cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dout, 4096 * 4);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  int bad = 0;
  {   // (1) tile through a map in global memory
    CUtensorMap hm[2];
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)H}, strides[1] = {(cuuint64_t)K * 4};
    cuuint32_t box[2] = {32, 128}, es[2] = {1, 1};
    memset(hm, 0, sizeof(hm));
    CUresult r = encode(&hm[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dW, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    CUtensorMap* dmaps;
    cudaMalloc(&dmaps, sizeof(hm));
    cudaMemcpy(dmaps, hm, sizeof(hm), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 20480);
    const int c0 = 96;
    k_tile<<<1, 128, 20480>>>(dmaps, 1, c0, 0, dout);
    std::vector<float> out(4096);
    cudaError_t e = cudaMemcpy(out.data(), dout, 4096 * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("k_tile: %s\n", cudaGetErrorString(e)); return 1; }
    for (int row = 0; row < 128; ++row)
      for (int c = 0; c < 32; ++c) {
        const float want = row < H ? W[(size_t)row * K + c0 + c] : 0.f;
        const float got = out[sw128(row, c * 4) / 4];
        if (want != got && ++bad < 5) printf("tile mismatch row %d col %d: got %g want %g\n", row, c, got, want);
      }
    printf("tile {32 x 128} SW128 via a global-memory tensor map, OOB rows zero-filled: %s\n", bad ? "MISMATCH" : "ok");
  }
  {   // (2) gather4
    CUtensorMap gm;
    cuuint64_t dims[2] = {1024, (cuuint64_t)N}, strides[1] = {(cuuint64_t)LD * 4};       // a 1024-column tap inside the 1920-wide matrix
    cuuint32_t box[2] = {32, 1}, es[2] = {1, 1};
    CUresult r = encode(&gm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, dX + 384, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode (gather) failed %d\n", (int)r); return 1; }
    std::vector<int> rows(64);
    for (int i = 0; i < 64; ++i) rows[i] = (i * 37 + 11) % N;
    int* drows;
    cudaMalloc(&drows, 64 * 4);
    cudaMemcpy(drows, rows.data(), 64 * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, 12288);
    const int c0 = 160;
    k_gather<<<1, 128, 12288>>>(gm, drows, c0, dout);
    std::vector<float> out(2048);
    cudaError_t e = cudaMemcpy(out.data(), dout, 2048 * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("k_gather: %s\n", cudaGetErrorString(e)); return 1; }
    int bad2 = 0;
    for (int row = 0; row < 64; ++row)
      for (int c = 0; c < 32; ++c) {
        const float want = X[(size_t)rows[row] * LD + 384 + c0 + c];
        const float got = out[sw128(row, c * 4) / 4];
        if (want != got && ++bad2 < 5) printf("gather4 mismatch row %d col %d: got %g want %g\n", row, c, got, want);
      }
    printf("tile::gather4, 64 rows in 16 instructions, SW128: %s\n", bad2 ? "MISMATCH" : "ok");
    bad += bad2;
  }
  return bad ? 1 : 0;
}
