// Thin inline-PTX layer over the Blackwell tensor-core path used by the "tc" engine:
// tcgen05.mma (kind::tf32, cta_group::1) with operands in shared memory and the accumulator in
// tensor memory, mbarrier completion, tcgen05.ld for the epilogue.
//
// Operand tiles use the canonical SWIZZLE_128B layouts (rows of 128 bytes, 16-byte chunks XOR-ed with
// row%8 inside 1024-byte atoms); the same bytes serve as a K-major tile (row = an M/N index, 32 tf32
// of K per row) or as an MN-major tile (row = a K index, 32 consecutive M/N per row).
// Every wait is bounded: a stuck barrier sets a flag and returns instead of hanging the GPU.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// first 1024-byte aligned address at or after p; written as pointer + offset (not integer arithmetic on the
// pointer) so the compiler keeps the shared address space and emits LDS/STS instead of generic LD/ST
__device__ __forceinline__ uint8_t* align1024(uint8_t* p) { return p + ((1024u - (smem_u32(p) & 1023u)) & 1023u); }

// byte offset of (row, byte_in_row) inside a SWIZZLE_128B tile whose base is 1024-byte aligned
__host__ __device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t byte_in_row) {
  return row * 128u + ((((byte_in_row >> 4) ^ (row & 7u)) << 4) | (byte_in_row & 15u));
}

// MN-major tf32 operands must use SWIZZLE_128B_BASE32B (CUTLASS sm100_common.inl: "for mn-major tf32
// operands, SW128_32B is the only available smem layout"): rows of 128 bytes (= 32 consecutive M/N for
// one K index), 32-byte chunks XOR-ed with row%4 inside 512-byte atoms of 4 K rows.
__device__ __forceinline__ uint32_t sw128_b32(uint32_t row, uint32_t byte_in_row) {
  return row * 128u + ((((byte_in_row >> 5) ^ (row & 3u)) << 5) | (byte_in_row & 31u));
}

constexpr uint64_t kLayoutSw128 = 2, kLayoutSw128Base32 = 1;

// 64-bit shared-memory matrix descriptor (sm_100 format: version=1 at bit 46, layout type in [61,64))
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint64_t layout = kLayoutSw128) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;      // descriptor version (Blackwell)
  d |= layout << 61;
  return d;
}

// 32-bit instruction descriptor, kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4)                         // c_format = F32
         | (2u << 7) | (2u << 10)          // a_format = b_format = TF32
         | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
         | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; one thread issues on behalf of the CTA
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// One lane of a converged warp (CUTLASS's elect_one_sync).  Branching on this instead of `lane == 0` tells the compiler that
// exactly one thread issues the tcgen05 instructions that follow: it then keeps the descriptors in uniform registers and
// emits straight UTCHMMAs, where a `lane == 0` branch gets an ELECT / R2UR.BROADCAST / BRA.U.ANY loop around every MMA
// (12-20 instructions, ~90 cycles per MMA against the tensor pipe's 32: r01 SASS of k_tc_fwd_ws).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0, laneid = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 %%rx;\n\t"
      ".reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, %2;\n\t"
      "@%%px mov.s32 %1, 1;\n\t"
      "mov.s32 %0, %%rx;\n\t"
      "}\n"
      : "+r"(laneid), "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}

// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- TMA (cp.async.bulk.tensor): operand tiles straight from a CUtensorMap into SWIZZLE_128B shared memory -------------
// The tensor maps live in global memory (one per work item, written by the host): the issuing thread acquires a map through
// the tensormap proxy once before its first use.
__device__ __forceinline__ void tmap_acquire(const void* tmap) {
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(tmap) : "memory");
}
// this thread arrives on the mbarrier and announces `bytes` of asynchronous-copy traffic that will complete on it
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 2-D box {c0.., c1..} (c0 = innermost coordinate, elements) -> shared memory; out-of-bounds elements arrive as zeros
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void* tmap, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
               ::"r"(dst_smem), "l"(tmap), "r"(c0), "r"(c1), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
// four rows r0..r3 of a 2-D tensor, box-width columns from column c0 -> four consecutive 128-byte rows of shared memory
__device__ __forceinline__ void tma_gather4(uint32_t dst_smem, const void* tmap, int c0, int r0, int r1, int r2, int r3, uint64_t* bar,
                                            uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4, %5, %6}], [%7], %8;"
               ::"r"(dst_smem), "l"(tmap), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// bounded wait (~2 s of SM clock at most); returns false if the phase never completed.  Under compute-sanitizer every
// instrumented access is orders of magnitude slower: profiles/run_gpu_sanitize.sh builds the library with both bounds raised.
#ifndef MFAS_WAIT_SPINS
#define MFAS_WAIT_SPINS (1u << 24)
#endif
#ifndef MFAS_WAIT_CYCLES
#define MFAS_WAIT_CYCLES 4000000000LL
#endif
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, uint32_t max_spins = MFAS_WAIT_SPINS) {
  const uint32_t a = smem_u32(bar);
  const long long t0 = clock64();
  for (uint32_t i = 0; i < max_spins; ++i) {
    if ((i & 1023u) == 1023u && clock64() - t0 > MFAS_WAIT_CYCLES) return false;
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(a), "r"(parity)
        : "memory");
    if (done) return true;
  }
  return false;
}

// generic-proxy smem writes -> visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one full warp allocates / frees `cols` TMEM columns (power of two >= 32)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// warp-collective: 32 lanes x 32 consecutive columns; thread i gets lane (taddr.lane + i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// warp-collective: 32 lanes x 16 consecutive columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// warp-collective: 32 lanes x 4 consecutive columns
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float* v) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

// warp-collective: 32 lanes x 8 consecutive columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 3xTF32 split: x = hi + lo (+ O(2^-22 |x|)), hi and lo both rounded to nearest tf32 (ties away from
// zero, i.e. cvt.rna.tf32.f32) so the error of the three-product scheme is unbiased -- the tensor core
// itself would truncate the low 13 bits.  Done with two integer ops per rounding (add half an ulp of
// tf32 to the magnitude, clear the low 13 bits): the conversion pipe (cvt) runs at quarter rate and
// would otherwise cost as many issue slots as the loads it decorates.
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = round_tf32(x);
  lo = round_tf32(x - hi);
}

// warp 0 polls the mbarrier, everybody else sleeps on the CTA barrier (no issue slots burnt spinning)
__device__ __forceinline__ bool cta_wait(uint64_t* bar, uint32_t parity, int* ok_flag) {
  if (threadIdx.x < 32) {
    const bool ok = mbar_wait(bar, parity);
    if (threadIdx.x == 0) *ok_flag = ok ? 1 : 0;
  }
  __syncthreads();
  return *ok_flag != 0;
}

}  // namespace umma
