// Standalone check of the tcgen05 primitives in mfas_b200/csrc/umma.cuh (run on the B200 box):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tests/cuda/umma_test.bin tests/cuda/umma_test.cu
// D[128 x N] = A * B^T with A, B given K-major ([rows][K]) or MN-major ([K][rows]), single tf32 pass
// (checked against a CPU product of tf32-truncated inputs) and 3xTF32 (checked against fp64).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../mfas_b200/csrc/umma.cuh"

using namespace umma;

template <int N, bool MN_MAJOR, int PASSES>
__global__ void __launch_bounds__(128) k_test(const float* __restrict__ A, const float* __restrict__ B, float* D, int K,
                                              int* err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  // tiles of one k-block (32 K values): A hi/lo 16 KB each, B hi/lo N*128 B each
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + 16384;
  uint8_t* b_hi = a_lo + 16384;
  uint8_t* b_lo = b_hi + N * 128;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tmem_alloc(&tmem_base, N < 32 ? 32 : N);
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base;
  constexpr uint32_t idesc = idesc_tf32(128, N, MN_MAJOR, MN_MAJOR);
  uint32_t phase = 0;
  for (int kb = 0; kb < K / 32; ++kb) {
    // stage the k-block: element (row r, k) -> K-major: tile row r, byte 4*(k%32)
    //                                          MN-major: block r/32, tile row k%32, byte 4*(r%32)
    for (int i = tid; i < 128 * 32; i += 128) {
      int r, k;
      float x;
      if (!MN_MAJOR) { r = i / 32; k = i % 32; x = A[(size_t)r * K + kb * 32 + k]; }
      else { k = i / 128; r = i % 128; x = A[(size_t)(kb * 32 + k) * 128 + r]; }
      float hi, lo;
      if (PASSES == 3) split_tf32(x, hi, lo); else { hi = x; lo = 0.f; }
      const uint32_t off = !MN_MAJOR ? sw128(r, 4 * k) : (uint32_t)(r / 32) * 4096u + sw128_b32(k, 4 * (r % 32));
      *(float*)(a_hi + off) = hi;
      *(float*)(a_lo + off) = lo;
    }
    for (int i = tid; i < N * 32; i += 128) {
      int r, k;
      float x;
      if (!MN_MAJOR) { r = i / 32; k = i % 32; x = B[(size_t)r * K + kb * 32 + k]; }
      else { k = i / N; r = i % N; x = B[(size_t)(kb * 32 + k) * N + r]; }
      float hi, lo;
      if (PASSES == 3) split_tf32(x, hi, lo); else { hi = x; lo = 0.f; }
      const uint32_t off = !MN_MAJOR ? sw128(r, 4 * k) : (uint32_t)(r / 32) * 4096u + sw128_b32(k, 4 * (r % 32));
      *(float*)(b_hi + off) = hi;
      *(float*)(b_lo + off) = lo;
    }
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      for (int ks = 0; ks < 4; ++ks) {        // 4 MMA k-steps of 8 inside the 32-wide k-block
        // K-major: advance 32 bytes inside the 128-byte row; MN-major: advance 8 rows = one 1024-byte atom
        const uint32_t adv = MN_MAJOR ? ks * 1024u : ks * 32u;
        // MN-major: LBO = stride between 32-wide M/N groups (32 rows x 128 B), SBO = 1024 (8 K rows)
        // (tf32 MN-major: SWIZZLE_128B_BASE32B, atoms of 4 K rows -> SBO = 512)
        const uint32_t lbo = MN_MAJOR ? 4096u : 16u, sbo = MN_MAJOR ? 512u : 1024u;
        const uint64_t lt = MN_MAJOR ? kLayoutSw128Base32 : kLayoutSw128;
        const uint64_t dah = smem_desc(smem_u32(a_hi) + adv, lbo, sbo, lt), dal = smem_desc(smem_u32(a_lo) + adv, lbo, sbo, lt);
        const uint64_t dbh = smem_desc(smem_u32(b_hi) + adv, lbo, sbo, lt), dbl = smem_desc(smem_u32(b_lo) + adv, lbo, sbo, lt);
        const uint32_t first = (kb == 0 && ks == 0) ? 0u : 1u;
        if (PASSES == 3) {
          mma_tf32(tm, dal, dbh, idesc, first);
          mma_tf32(tm, dah, dbl, idesc, 1u);
          mma_tf32(tm, dah, dbh, idesc, 1u);
        } else {
          mma_tf32(tm, dah, dbh, idesc, first);
        }
      }
      mma_commit(&bar);
    }
    if (!mbar_wait(&bar, phase)) { if (tid == 0) atomicExch(err, 1 + kb); break; }
    phase ^= 1;
  }
  tc_fence_after();
  // epilogue: warp w owns TMEM lanes [32w, 32w+32); thread = one row of D
  for (int c0 = 0; c0 < N; c0 += 32) {
    float v[32];
    tmem_ld32(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int j = 0; j < 32; ++j)
      if (c0 + j < N) D[(size_t)tid * N + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_free(tm, N < 32 ? 32 : N);
}

static float tf32_trunc(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  u &= 0xFFFFE000u;
  memcpy(&x, &u, 4);
  return x;
}

template <int N, bool MN, int PASSES>
static int run(int K, const char* name) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N, -1.f);
  srand(7);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& x : B) x = (float)rand() / RAND_MAX * 2.f - 1.f;
  float *dA, *dB, *dD;
  int* derr;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&derr, 4);
  cudaMemset(derr, 0, 4);
  // logical A(m,k), B(n,k); memory layout depends on MN
  std::vector<float> Am(A.size()), Bm(B.size());
  for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) Am[MN ? (size_t)k * 128 + m : (size_t)m * K + k] = A[(size_t)m * K + k];
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) Bm[MN ? (size_t)k * N + n : (size_t)n * K + k] = B[(size_t)n * K + k];
  cudaMemcpy(dA, Am.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bm.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const size_t smem = 1024 + 2 * 16384 + 2 * N * 128;
  cudaFuncSetAttribute(k_test<N, MN, PASSES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_test<N, MN, PASSES><<<1, 128, smem>>>(dA, dB, dD, K, derr);
  cudaError_t e = cudaDeviceSynchronize();
  int herr = 0;
  cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) {
        const float a = A[(size_t)m * K + k], b = B[(size_t)n * K + k];
        s += PASSES == 3 ? (double)a * b : (double)tf32_trunc(a) * tf32_trunc(b);
      }
      maxerr = fmax(maxerr, fabs(s - D[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(s));
    }
  const double rel = maxerr / maxref;
  const double tol = 5e-6;
  const bool ok = e == cudaSuccess && herr == 0 && rel < tol;
  printf("%-28s N=%3d K=%4d passes=%d : cuda=%s barrier_err=%d max_rel_err=%.3e %s\n", name, N, K, PASSES,
         cudaGetErrorString(e), herr, rel, ok ? "OK" : "FAIL");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(derr);
  return ok ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<64, false, 1>(64, "K-major tf32");
  bad += run<64, false, 3>(64, "K-major 3xTF32");
  bad += run<128, false, 3>(256, "K-major 3xTF32");
  bad += run<64, true, 1>(64, "MN-major tf32");
  bad += run<64, true, 3>(64, "MN-major 3xTF32");
  bad += run<32, true, 3>(128, "MN-major 3xTF32");
  bad += run<128, true, 3>(64, "MN-major 3xTF32");
  printf(bad ? "UMMA TEST FAILED (%d)\n" : "UMMA TEST PASSED\n", bad);
  return bad;
}
