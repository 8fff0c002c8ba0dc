// Engine "tc": the GEMM-shaped stages of a training step on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM), fp32-accurate through a TF32 split
// (x = hi + lo, both rounded to tf32; D = [A_lo*B_lo +] A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, fp32
// accumulate; the small chain GEMMs take all four products, the two streaming GEMMs three).
//
// The step is HBM-bound (10 FLOP/B): 95 % of its bytes are the feature columns of the L weight
// matrices (forward: read once; backward: p/m/v read + written by Adam).  Those columns do not
// depend on the chain h_0 -> h_1 -> ... so they are processed for ALL layers in one launch each way:
//
//   k_tc_fwd_all   zT_l[h,b] partials over the feature columns of every layer (split-K work items)
//   k_fwd_layer    per layer: + h_{l-1} W_l[:,hid]^T (small), + bias, activation, BatchNorm, dropout
//   k_head         classifier + softmax-CE (+ backward + Adam of the classifier)      [kernels_ffma.cuh]
//   k_dzx          per layer, reverse: dh_l = dz_{l+1} W_{l+1}[:,hid] (small), BN/act backward -> dz_l
//   k_tc_bwd_all   dW_l^T[k,h] = x^T dz for every layer, every 128-column chunk, Adam(L2) fused in the
//                  epilogue straight out of TMEM: the gradient never reaches HBM
//
// Operands are staged by coalesced float4 loads through registers -- the gather of the batch rows, the
// concat bookkeeping and the hi/lo split all happen on that pass -- into the canonical 128B-swizzled
// shared-memory tiles the tensor core reads (umma.cuh).
// Shapes: H multiple of 64 (<= 256), tap widths multiples of 128, batch <= 128.
#pragma once
#include <cuda.h>      // CUtensorMap (type only: the encoder is fetched through cudaGetDriverEntryPoint, nothing links libcuda)

#include "common.cuh"
#include "umma.cuh"

namespace mfas {

constexpr int TC_THREADS = 256;
constexpr int TC_KB_PER_ITEM = 32;    // forward: at most this many k-blocks (32 columns each) per work item
constexpr int TC_BWD_KT = 128;        // backward: weight columns (TMEM lanes) per work item
constexpr int TC_BWD_HT = 64;         // backward: output rows h (TMEM columns) per work item

struct TcErr {
  int* flag;                          // set when a bounded barrier wait expires (never hangs the GPU)
  long long* timeline;                // optional [n_cand][16] clock64 stamps of k_chain_all's phases (MFAS_CHAIN_TIMELINE=1)
  int l2_hints;                       // bit 0: k_tc_fwd_ws loads its once-per-launch streams with L2 evict-first priority; bit 1: k_tc_bwd_ws too
                                      // (MFAS_L2_HINTS; default 1 -- on the backward's read-modify-write stream the hint costs 55 %, r01n)
};

// `per`: k-blocks per work item at most.  TC_KB_PER_ITEM unless the group chose a smaller one (DCand::kb_item: small groups of the
// inner_repr 16 / 32 path, so that every SM gets about the same number of k-blocks); every producer and consumer of a group's
// partial sums uses the group's value.
__host__ __device__ __forceinline__ int tc_fwd_items(int d_ske, int d_rgb, int per = TC_KB_PER_ITEM) {
  return (((d_ske + d_rgb) >> 5) + per - 1) / per;
}
// k-block range of split `split` of a layer with nkb feature k-blocks (even split over tc_fwd_items pieces)
__host__ __device__ __forceinline__ void tc_fwd_range(int nkb, int split, int& kb0, int& kb1, int per = TC_KB_PER_ITEM) {
  const int n = (nkb + per - 1) / per, each = (nkb + n - 1) / n;
  kb0 = split * each;
  kb1 = kb0 + each < nkb ? kb0 + each : nkb;
}
// Alpha-gated candidates (MFAS_FLAG_ALPHAS): forward work items never straddle the boundary between the two modalities -- the
// first tc_fwd_items_of(d_ske) items of a layer cover the ske columns, the rest the rgb columns -- so the consumer of the partial
// sums applies the gate of a modality as ONE factor per partial: z = s * (W_ske x_ske) + (1 - s) * (W_rgb x_rgb) + ...
__host__ __device__ __forceinline__ int tc_fwd_items_of(int width, int per = TC_KB_PER_ITEM) { return ((width >> 5) + per - 1) / per; }
__host__ __device__ __forceinline__ int tc_fwd_items_g(int d_ske, int d_rgb, bool gated, int per = TC_KB_PER_ITEM) {
  return gated ? tc_fwd_items_of(d_ske, per) + tc_fwd_items_of(d_rgb, per) : tc_fwd_items(d_ske, d_rgb, per);
}
// k-block range of item `split` of a layer (gated: an even split of the item's own modality)
__host__ __device__ __forceinline__ void tc_fwd_range_g(int d_ske, int d_rgb, int split, bool gated, int& kb0, int& kb1, int per = TC_KB_PER_ITEM) {
  if (!gated) { tc_fwd_range((d_ske + d_rgb) >> 5, split, kb0, kb1, per); return; }
  const int ns = tc_fwd_items_of(d_ske, per);
  if (split < ns) tc_fwd_range(d_ske >> 5, split, kb0, kb1, per);
  else { tc_fwd_range(d_rgb >> 5, split - ns, kb0, kb1, per); kb0 += d_ske >> 5; kb1 += d_ske >> 5; }
}
__host__ __device__ __forceinline__ int tc_bwd_items(int K) { return (K + TC_BWD_KT - 1) / TC_BWD_KT; }

__device__ __forceinline__ void store_split(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, float4 x) {
  float4 h, l;
  umma::split_tf32(x.x, h.x, l.x); umma::split_tf32(x.y, h.y, l.y);
  umma::split_tf32(x.z, h.z, l.z); umma::split_tf32(x.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// ---------------------------------------------------------------------------------------------
// forward, all layers: partial zT over a slice of FEATURE columns of one layer.
// grid = (max items per candidate, ceil(H/128), candidates);  NPAD = batch padded to the MMA N.
// part[cand][item][NPAD/4][Hp][4]: four consecutive batch rows of one output column are one float4, and for a fixed
// group of four rows consecutive columns are consecutive float4s -- the producer (TMEM lane = column) and the consumers
// (lane = column) both move 512 contiguous bytes per warp instruction
// dynamic smem (1024-aligned): A_hi 16K | A_lo 16K | B_hi NPAD*128 | B_lo NPAD*128
// (Measured alternative, kept out: double-buffering the smem stage so the stores of block k+1 overlap
//  the MMAs of block k costs a third of the resident CTAs -- 2 instead of 3 per SM -- and is 30 % slower.)
// ---------------------------------------------------------------------------------------------
template <int NPAD>
__global__ void __launch_bounds__(TC_THREADS, NPAD == 64 ? 3 : 2)
k_tc_fwd_all(const DCand* __restrict__ cands, DCache cache, BatchRef batch, float* part_base,
             long long part_stride_cand, TcErr err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);
  const int cand = blockIdx.z;
  const DCand& cd = cands[cand];
  const int H = cd.H;
  const int m0 = blockIdx.y * 128;
  if (m0 >= H) return;
  // decode the work item -> (layer, split)
  int layer = 0, split = blockIdx.x;
  for (; layer < cd.L; ++layer) {
    const int n = tc_fwd_items(cd.layer[layer].d_ske, cd.layer[layer].d_rgb, cd.kb_item);
    if (split < n) break;
    split -= n;
  }
  if (layer >= cd.L) return;
  const DLayer& ly = cd.layer[layer];
  const int K = ly.K, fs = ly.d_ske, fr = ly.d_rgb;
  const int nkb = (fs + fr) >> 5;
  int kb0, kb1;
  tc_fwd_range(nkb, split, kb0, kb1, cd.kb_item);
  const int nrows = batch.n_rows, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + 16384;
  uint8_t* b_hi = a_lo + 16384;
  uint8_t* b_lo = b_hi + NPAD * 128;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int ok_flag;
  __shared__ int rowid[MFAS_MAX_BATCH];

  if (warp == 0) umma::tmem_alloc(&tmem_slot, NPAD);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
  for (int r = tid; r < MFAS_MAX_BATCH; r += TC_THREADS) rowid[r] = r < nrows ? batch_row(batch, cand, r) : 0;
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tm = tmem_slot;
  const float* Wbase = cd.p + ly.oW;

  constexpr int A_IT = 128 * 8 / TC_THREADS;     // 4 float4 of W per thread per k-block
  constexpr int B_IT = NPAD * 8 / TC_THREADS;    // 2 or 4 float4 of x
  struct Regs { float4 a[A_IT], b[B_IT]; };
  auto load_kb = [&](Regs& rg, int kb) {
    const int kg = kb << 5;                      // concat column; tap widths are multiples of 32
    const float* src; long long ld; int kl;
    if (kg < fs) { src = cache.ske[ly.ske_tap]; ld = cache.ske_ld[ly.ske_tap]; kl = kg; }
    else { src = cache.rgb[ly.rgb_tap]; ld = cache.rgb_ld[ly.rgb_tap]; kl = kg - fs; }
#pragma unroll
    for (int i = 0; i < A_IT; ++i) {
      const int idx = tid + TC_THREADS * i, r = idx >> 3, c4 = idx & 7;
      rg.a[i] = (m0 + r < H) ? *reinterpret_cast<const float4*>(Wbase + (long long)(m0 + r) * K + kg + c4 * 4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_IT; ++i) {
      const int idx = tid + TC_THREADS * i, r = idx >> 3, c4 = idx & 7;
      rg.b[i] = (r < nrows) ? __ldg(reinterpret_cast<const float4*>(src + (long long)rowid[r] * ld + kl + c4 * 4))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_kb = [&](const Regs& rg) {
#pragma unroll
    for (int i = 0; i < A_IT; ++i) {
      const int idx = tid + TC_THREADS * i;
      store_split(a_hi, a_lo, umma::sw128(idx >> 3, (idx & 7) * 16), rg.a[i]);
    }
#pragma unroll
    for (int i = 0; i < B_IT; ++i) {
      const int idx = tid + TC_THREADS * i;
      store_split(b_hi, b_lo, umma::sw128(idx >> 3, (idx & 7) * 16), rg.b[i]);
    }
  };

  constexpr uint32_t idesc = umma::idesc_tf32(128, NPAD, false, false);
  uint32_t phase = 0;
  bool ok = true;
  // one k-block: (wait for the tensor core to release the tiles) -> store this block's registers ->
  // refill the same registers with the block two steps ahead -> issue this block's MMAs.
  // Two register sets alternate, so every thread always has two k-blocks of global loads in flight.
  auto stage = [&](Regs& rg, int kb) {
    if (kb > kb0) {
      ok = umma::cta_wait(&bar, phase, &ok_flag);
      phase ^= 1;
      if (!ok) return;
    }
    store_kb(rg);
    umma::fence_async_smem();
    __syncthreads();
    if (kb + 2 < kb1) load_kb(rg, kb + 2);
    if (tid == 0) {
      umma::tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint32_t adv = ks * 32u;               // 8 tf32 along K inside the 128-byte row
        const uint64_t dah = umma::smem_desc(umma::smem_u32(a_hi) + adv, 16, 1024);
        const uint64_t dal = umma::smem_desc(umma::smem_u32(a_lo) + adv, 16, 1024);
        const uint64_t dbh = umma::smem_desc(umma::smem_u32(b_hi) + adv, 16, 1024);
        const uint64_t dbl = umma::smem_desc(umma::smem_u32(b_lo) + adv, 16, 1024);
        umma::mma_tf32(tm, dal, dbh, idesc, (kb > kb0 || ks > 0) ? 1u : 0u);
        umma::mma_tf32(tm, dah, dbl, idesc, 1u);
        umma::mma_tf32(tm, dah, dbh, idesc, 1u);
      }
      umma::mma_commit(&bar);
    }
  };
  Regs r0, r1;
  load_kb(r0, kb0);
  if (kb0 + 1 < kb1) load_kb(r1, kb0 + 1);
  for (int kb = kb0; kb < kb1 && ok; kb += 2) {
    stage(r0, kb);
    if (kb + 1 < kb1 && ok) stage(r1, kb + 1);
  }
  if (ok) ok = umma::cta_wait(&bar, phase, &ok_flag);
  if (!ok && tid == 0) atomicExch(err.flag, 1);
  umma::tc_fence_after();

  // epilogue: thread = one output column h (TMEM lane), 32 consecutive batch rows per tcgen05.ld
  const int Hp = ((H + 127) >> 7) << 7;
  float* part = part_base + (long long)cand * part_stride_cand + (long long)blockIdx.x * Hp * NPAD;
  const int h_loc = (warp & 3) * 32 + lane;
#pragma unroll
  for (int cc = 0; cc < NPAD / 64; ++cc) {
    const int c0 = (warp >> 2) * (NPAD / 2) + cc * 32;
    float v[32];
    umma::tmem_ld32(tm + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
    float4* dst = reinterpret_cast<float4*>(part) + (long long)(c0 >> 2) * Hp + m0 + h_loc;
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[(long long)j * Hp] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tm, NPAD);
}

// ---------------------------------------------------------------------------------------------
// forward, all layers, persistent + warp-specialised.  Same work items and partial-sum layout as
// k_tc_fwd_all, but one CTA per SM walks a static list of items, the bytes in flight are bounded by
// shared memory instead of registers, and no role ever waits for another inside a k-block:
//   warp  17    W producer : one thread issues the W_hi tiles by TMA into their own ring (7 deep; cp.async by the loaders when
//                            no tensor-map encoder is available)                       (wfree[s] <- MMA; wland[s] -> converters)
//   warps 0-3   loaders    : cp.async (16 B per lane, swizzle applied to the destination) of the gathered x rows straight into
//                            canonical SW128 K-major tiles, a ring of [x_hi | x_lo] stages (5 deep).  The raw tiles ARE the "hi"
//                            operands: kind::tf32 reads the top 19 bits of each 32-bit container, i.e. hi = trunc_tf32(x) for free.
//                            (xfree[s] <- MMA; xland[s] -> converters, signalled by cp.async itself)
//   warps 4-11  converters : two groups of four warps, every other k-block: lo = rna_tf32(x - trunc_tf32(x)) (the subtraction is
//                            exact in fp32): x_lo next to x_hi, W_lo into a ring of two            (lofree[s] <- MMA; lofull[s] -> MMA)
//   warp  12    MMA        : per k-block 4 x {W_hi [x_hi; x_lo]^T -> [main | correction] accumulator columns (ONE N = 2 NPAD MMA);
//                            W_lo x_hi^T -> correction}.  The tensor core truncates when it adds into the fp32
//                            accumulator, so the long chain of adds is the dominant error (measured 5e-6..1e-5
//                            of max|logit| against 1e-6 for the fp32 reference); keeping the small cross terms
//                            out of the main chain cuts its length by three            (tfull[t] -> epilogue)
//   warps 13-16 epilogue   : main + correction -> partial sums in global memory                      (tempty[t] -> MMA)
// The stream of k-blocks runs across item boundaries, so the pipeline never drains inside a launch.
// (r01 ncu of k_tc_fwd_all: 3.4 TB/s, stalled on the CTA barrier + MMA round trip of every 24 KB k-block.)
// ---------------------------------------------------------------------------------------------
struct __align__(16) FwdItem {
  const float* W;                 // &params[oW + m0 * K]
  long long part_off;             // float offset of this item's tile (+ 4 * m0: first row of this CTA's 128) in the partial-sum buffer
  int K, kb0, kb1, fs_kb;         // row stride of W; k-block range [kb0, kb1); first k-block of the rgb tap
  int ske_tap, rgb_tap, cand, rows_valid;   // rows_valid = min(128, H - m0)
  int Hp, pad0, pad1, pad2;       // H rounded up to 128: rows of the partial-sum tile
};

// Stage layout (r02): raw stage = [W_hi 16 KB | x_hi | x_lo] -- the x halves ADJACENT, so that [x_hi; x_lo] is one N = 2 NPAD
// operand -- and a ring of W_lo tiles.  Per 8 columns: W_hi [x_hi; x_lo]^T -> [main | correction] accumulator columns in ONE MMA
// (W_hi is read from shared memory once instead of twice), W_lo x_hi^T -> correction: 8 MMAs per k-block instead of 12 (what a
// k-block cost on the issuing thread: ~70 cycles per MMA, profiles/r02dk_roles.txt).  The converter warps work in two groups of
// four, each with a W_lo stage of its own and every other k-block (a pass is two barrier waits, a shared-memory round trip, a
// proxy fence and an arrive: its latency, not its bytes, was the converters' rate).
template <int NPAD, int XR = 0> struct FwdWs {     // XR = 1: W / x rings 8 / 4 deep instead of 7 / 5 (128 batch rows: 6 / 3 instead of 4 / 4); A/B switch MFAS_FWD_XR
  // three rings: W_hi tiles (the stream's HBM bytes: as deep as shared memory allows -- what bounds the stream is loaded HBM
  // latency x tiles in flight), [x_hi | x_lo] tiles (gathered rows, mostly L2 hits), W_lo tiles
  static constexpr int WR = NPAD == 64 ? 7 + XR : 4 + 2 * XR, XS = NPAD == 64 ? 5 - XR : 4 - XR, LO = 2;
  static constexpr uint32_t A_BYTES = 16384, B_BYTES = NPAD * 128, X_TILE = 2 * B_BYTES, LO_TILE = A_BYTES;
  static constexpr size_t SMEM = 1024 + (size_t)WR * A_BYTES + (size_t)XS * X_TILE + (size_t)LO * LO_TILE;
  static constexpr int LOADERS = 128, CONVERTERS = 256, CGROUPS = 2, THREADS = 18 * 32;     // + one W-producer warp (TMA)
};

// L2 eviction priority for the streams that are read exactly once per launch (weights, Adam moments, gathered feature
// rows: >> L2 in total).  Marked evict-first they stop flushing what IS re-read out of the 126 MB L2 -- the split-K
// partial sums between k_tc_fwd_ws and the chain, the activations / dz between the chain and k_tc_bwd_ws.
__device__ __forceinline__ uint64_t l2_stream_policy(bool evict_first) {
  uint64_t p;
  if (evict_first) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_keep_policy() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async16_zfill(uint32_t dst_smem, const void* src, bool valid, uint64_t policy) {
  const int n = valid ? 16 : 0;    // src-size 0: the 16 destination bytes are zero-filled, src is not read
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(dst_smem), "l"(src), "r"(n), "l"(policy) : "memory");
}
// the mbarrier receives one arrival from this thread once all of its earlier cp.async have landed
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}

// TMA operand staging of the forward stream (default; MFAS_FWD_TMA=0 keeps the cp.async loaders):
//   W tile  : one cp.async.bulk.tensor.2d per k-block -- box {32 columns, 128 rows} of the work item's own tensor map (base =
//             first row of the item's 128-row tile, extent = its valid rows, so rows beyond inner_repr arrive as zeros),
//             SWIZZLE_128B: exactly the K-major operand tile the MMA descriptors describe;
//   x tile  : tile::gather4 -- four batch rows per instruction from the tap's tensor map (box {32 columns, 1 row}), one
//             instruction per lane of the loader warp, row indices = the candidate's batch order; rows beyond the batch name
//             a row outside the tensor and arrive as zeros.
// All of a stage's copies complete on landed[stage] through complete_tx; the 16 x (NPAD / 4 + 1) address computations and
// issue slots per k-block that 4 loader warps spent on cp.async become NPAD / 4 + 1 instructions of one warp.
// role clock stamps of the persistent streams (diagnostic builds only: tests/cuda/fwd_small_timeline.py): event e of step n of CTA 0
#ifdef MFAS_KSTAMPS
#define MFAS_KSTAMP(first, n, e) do { if (err.timeline && blockIdx.x == 0 && (n) >= (first) && (n) < (first) + 16) err.timeline[(first == 32 ? 0 : 128) + ((n) - (first)) * 8 + (e)] = clock64(); } while (0)
#else
#define MFAS_KSTAMP(first, n, e) do { } while (0)
#endif
struct TapMaps { CUtensorMap ske[MFAS_NUM_TAPS], rgb[MFAS_NUM_TAPS]; };      // feature taps of one cache: dims {width, n_rows}, box {32, 1}

template <int NPAD, int XR>
__global__ void __launch_bounds__((FwdWs<NPAD, XR>::THREADS), 1)
k_tc_fwd_ws(const FwdItem* __restrict__ items, int n_items, DCache cache, BatchRef batch, float* part_base, TcErr err,
            const CUtensorMap* __restrict__ wmaps, const __grid_constant__ TapMaps taps, int use_tma) {
  using Cfg = FwdWs<NPAD, XR>;
  constexpr int WR = Cfg::WR, XS = Cfg::XS, LQ = Cfg::LO;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);                       // W ring
  uint8_t* x_base = smem + WR * Cfg::A_BYTES;                      // [x_hi | x_lo] ring
  uint8_t* lo_base = x_base + XS * Cfg::X_TILE;                    // W_lo ring
  __shared__ uint64_t wland[WR], wfree[WR], xland[XS], xfree[XS], lofull[LQ], lofree[LQ], tfull[2], tempty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nrows = batch.n_rows;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 4 * NPAD);
  if (tid == 32) {
    // use_tma: 0 = cp.async loaders, 1 = W tiles through TMA + x through cp.async, 2 = W through TMA + x through tile::gather4
    for (int i = 0; i < WR; ++i) { umma::mbar_init(&wland[i], use_tma ? 1 : Cfg::LOADERS); umma::mbar_init(&wfree[i], 1); }
    for (int i = 0; i < XS; ++i) { umma::mbar_init(&xland[i], use_tma == 2 ? 1 : Cfg::LOADERS); umma::mbar_init(&xfree[i], 1); }
    for (int i = 0; i < LQ; ++i) { umma::mbar_init(&lofull[i], Cfg::CONVERTERS / 32 / Cfg::CGROUPS); umma::mbar_init(&lofree[i], 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&tfull[i], 1); umma::mbar_init(&tempty[i], 4); }
    umma::fence_mbar_init();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  griddep_launch();
  griddep_wait();                                                  // the weights come from the previous step's backward stream
  const uint32_t tm = tmem_slot;
  const int n_my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  bool ok = true;

  if (warp == 17) {
    // ================================ W producer (TMA): one thread ================================
    // runs ahead of the x loaders by the difference of the ring depths: W_hi tiles are the HBM bytes of the stream
    if (use_tma && lane == 0) {
      const uint32_t s0 = umma::smem_u32(smem);
      const uint64_t stream_policy = l2_stream_policy((err.l2_hints & 1) != 0);
      int n = 0;
      for (int i = 0; i < n_my && ok; ++i) {
        const int idx = blockIdx.x + i * gridDim.x;
        const int kb0 = items[idx].kb0, kb1 = items[idx].kb1;
        const CUtensorMap* wm = wmaps + idx;
        umma::tmap_acquire(wm);
        for (int kb = kb0; kb < kb1; ++kb, ++n) {
          const int sw = n % WR;
          if (n >= WR && !umma::mbar_wait(&wfree[sw], ((n / WR) & 1) ^ 1)) { ok = false; break; }
          umma::mbar_arrive_expect_tx(&wland[sw], Cfg::A_BYTES);
          umma::tma_load_2d(s0 + sw * Cfg::A_BYTES, wm, 32 * kb, 0, &wland[sw], stream_policy);
          MFAS_KSTAMP(32, n, 1);
        }
      }
    }
  } else if (warp < 4 && use_tma == 2) {
    // ================================ x loader (TMA gather4): warp 0 ===============================
    if (warp == 0) {
      const uint32_t x0 = umma::smem_u32(x_base);
      const uint64_t stream_policy = l2_stream_policy((err.l2_hints & 1) != 0);
      const uint64_t x_policy = (err.l2_hints & 8) ? l2_keep_policy() : (err.l2_hints & 4) ? l2_stream_policy(false) : stream_policy;
      const int oob_row = (int)cache.n_rows;                       // a row index outside every tap tensor: arrives as zeros
      auto load_rows = [&](const FwdItem& it, int* rows) {         // lane j gathers batch rows 4j .. 4j+3
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int row = 4 * lane + q;
          rows[q] = (row < nrows && row < NPAD) ? batch_row(batch, it.cand, row) : oob_row;
        }
      };
      FwdItem cur{}, nxt{};
      int rows_cur[4] = {0, 0, 0, 0}, rows_nxt[4] = {0, 0, 0, 0};
      if (n_my > 0) { cur = items[blockIdx.x]; load_rows(cur, rows_cur); }
      int n = 0;
      for (int i = 0; i < n_my && ok; ++i) {
        const int idx = blockIdx.x + i * gridDim.x;
        if (i + 1 < n_my) { nxt = items[idx + gridDim.x]; load_rows(nxt, rows_nxt); }      // in flight while this item's k-blocks are issued
        const CUtensorMap* sm = &taps.ske[cur.ske_tap];
        const CUtensorMap* rm = &taps.rgb[cur.rgb_tap];
#pragma unroll 1
        for (int kb = cur.kb0; kb < cur.kb1; ++kb, ++n) {
          const int sx = n % XS;
          if (n >= XS && !umma::mbar_wait(&xfree[sx], ((n / XS) & 1) ^ 1)) { ok = false; break; }
          const uint32_t b = x0 + sx * Cfg::X_TILE;
          if (lane == 0) umma::mbar_arrive_expect_tx(&xland[sx], Cfg::B_BYTES);
          __syncwarp();
          if (lane < NPAD / 4) {
            const bool ske = kb < cur.fs_kb;
            umma::tma_gather4(b + lane * 512, ske ? sm : rm, 32 * (ske ? kb : kb - cur.fs_kb), rows_cur[0], rows_cur[1], rows_cur[2],
                              rows_cur[3], &xland[sx], x_policy);
          }
        }
        cur = nxt;
#pragma unroll
        for (int q = 0; q < 4; ++q) rows_cur[q] = rows_nxt[q];
      }
    }
  } else if (warp < 4) {
    // ================================ loaders (cp.async) ==========================================
    constexpr int WJ = 8, XJ = NPAD / 16;                          // rows r + 16 j of the W / x tile
    const int r = tid >> 3, c = tid & 7;                           // 16-byte chunk c of the 128-byte row
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);   // sw128(r + 16 j, 16 c) = off + 2048 j
    const uint32_t s0 = umma::smem_u32(smem), x0 = umma::smem_u32(x_base);
    const uint64_t stream_policy = l2_stream_policy((err.l2_hints & 1) != 0);
    // the gathered x rows are read again by the backward stream of the same step: bit 2 = default priority, bit 3 = evict-last
    const uint64_t x_policy = (err.l2_hints & 8) ? l2_keep_policy() : (err.l2_hints & 4) ? l2_stream_policy(false) : stream_policy;
    // Two-deep software pipeline over the item list, so that no item starts on a chain of dependent loads (descriptor ->
    // gather indices -> first cp.async): the descriptor of item i+2 and the row indices of item i+1 are requested while
    // the k-blocks of item i are issued (r01t ncu: long-scoreboard was this kernel's top stall; an item is only ~24 k-blocks).
    auto load_rows = [&](const FwdItem& it, int* rows) {
#pragma unroll
      for (int j = 0; j < XJ; ++j) {
        const int row = r + 16 * j;
        rows[j] = batch_row(batch, it.cand, row < nrows ? row : 0);
      }
    };
    FwdItem cur{}, nxt{};
    int rows_cur[XJ], rows_nxt[XJ];
#pragma unroll
    for (int j = 0; j < XJ; ++j) rows_cur[j] = rows_nxt[j] = 0;
    if (n_my > 0) { cur = items[blockIdx.x]; load_rows(cur, rows_cur); }
    if (n_my > 1) nxt = items[blockIdx.x + gridDim.x];
    int n = 0;
    for (int i = 0; i < n_my && ok; ++i) {
      if (i + 1 < n_my) load_rows(nxt, rows_nxt);
      FwdItem nx2{};
      if (i + 2 < n_my) nx2 = items[blockIdx.x + (i + 2) * gridDim.x];
      const FwdItem& it = cur;
      const long long wstride = 16LL * it.K;
      const float* wp = it.W + (long long)r * it.K + c * 4;        // &W[m0 + r][4 c]
      const float* xs[XJ]; const float* xr[XJ];                    // gathered rows of the two taps, indexed by concat column
#pragma unroll
      for (int j = 0; j < XJ; ++j) {
        const long long gr = rows_cur[j];
        xs[j] = cache.ske[it.ske_tap] + gr * cache.ske_ld[it.ske_tap] + c * 4;
        xr[j] = cache.rgb[it.rgb_tap] + gr * cache.rgb_ld[it.rgb_tap] + c * 4 - 32LL * it.fs_kb;
      }
#pragma unroll 1
      for (int kb = it.kb0; kb < it.kb1; ++kb, ++n) {
        const int sx = n % XS;
        if (!use_tma) {                                            // the W tile by cp.async too (no TMA descriptor encoder)
          const int sw = n % WR;
          if (n >= WR && !umma::mbar_wait(&wfree[sw], ((n / WR) & 1) ^ 1)) { ok = false; break; }
          const uint32_t a = s0 + sw * Cfg::A_BYTES + off;
          const float* w = wp + 32LL * kb;
#pragma unroll
          for (int j = 0; j < WJ; ++j) cp_async16_zfill(a + j * 2048, w + j * wstride, r + 16 * j < it.rows_valid, stream_policy);
          cp_async_arrive_noinc(&wland[sw]);
        }
        if (n >= XS && !umma::mbar_wait(&xfree[sx], ((n / XS) & 1) ^ 1)) { ok = false; break; }
        if (tid == 0) MFAS_KSTAMP(32, n, 7);
        const uint32_t b = x0 + sx * Cfg::X_TILE + off;
        const bool ske = kb < it.fs_kb;
#pragma unroll
        for (int j = 0; j < XJ; ++j) cp_async16_zfill(b + j * 2048, (ske ? xs[j] : xr[j]) + 32LL * kb, r + 16 * j < nrows, x_policy);
        cp_async_arrive_noinc(&xland[sx]);
        if (tid == 0) MFAS_KSTAMP(32, n, 0);
      }
      cur = nxt; nxt = nx2;
#pragma unroll
      for (int j = 0; j < XJ; ++j) rows_cur[j] = rows_nxt[j];
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp < 12) {
    // ================================ converters ==================================================
    // group g: k-blocks g, g + CG, ...; x_lo goes next to x_hi in the x stage, W_lo into the group's turn of the W_lo ring
    const int ct = tid - 128;
    constexpr int CG = Cfg::CGROUPS, TG = Cfg::CONVERTERS / CG;
    constexpr int NA = (int)(Cfg::A_BYTES / 16), NCH = (int)((Cfg::A_BYTES + Cfg::B_BYTES) / 16), CJ = NCH / TG;
    const int grp = ct / TG, gt = ct % TG;
    int total = 0;
    for (int i = 0; i < n_my; ++i) { const FwdItem& it = items[blockIdx.x + i * gridDim.x]; total += it.kb1 - it.kb0; }
#pragma unroll 1
    for (int n = grp; n < total; n += CG) {
      const int sw = n % WR, sx = n % XS, sl = n % LQ;
      if (!umma::mbar_wait(&wland[sw], (n / WR) & 1)) { ok = false; break; }
      if (!umma::mbar_wait(&xland[sx], (n / XS) & 1)) { ok = false; break; }
      if (gt == 0) MFAS_KSTAMP(32, n, 2);
      if (n >= LQ && !umma::mbar_wait(&lofree[sl], ((n / LQ) & 1) ^ 1)) { ok = false; break; }
      if (gt == 0) MFAS_KSTAMP(32, n, 6);
      const uint8_t* wh = smem + sw * Cfg::A_BYTES;
      uint8_t* xh = x_base + sx * Cfg::X_TILE;
      uint8_t* wl = lo_base + sl * Cfg::LO_TILE;
      float4 x[CJ];
#pragma unroll
      for (int j = 0; j < CJ; ++j) {
        const int ch = gt + j * TG;
        x[j] = ch < NA ? *reinterpret_cast<const float4*>(wh + (size_t)ch * 16) : *reinterpret_cast<const float4*>(xh + (size_t)(ch - NA) * 16);
      }
#pragma unroll
      for (int j = 0; j < CJ; ++j) {
        const int ch = gt + j * TG;
        float4 l;
        l.x = umma::round_tf32(x[j].x - __uint_as_float(__float_as_uint(x[j].x) & 0xFFFFE000u));
        l.y = umma::round_tf32(x[j].y - __uint_as_float(__float_as_uint(x[j].y) & 0xFFFFE000u));
        l.z = umma::round_tf32(x[j].z - __uint_as_float(__float_as_uint(x[j].z) & 0xFFFFE000u));
        l.w = umma::round_tf32(x[j].w - __uint_as_float(__float_as_uint(x[j].w) & 0xFFFFE000u));
        if (ch < NA) *reinterpret_cast<float4*>(wl + (size_t)ch * 16) = l;
        else *reinterpret_cast<float4*>(xh + Cfg::B_BYTES + (size_t)(ch - NA) * 16) = l;
      }
      umma::fence_async_smem();                                    // lo (generic proxy) and the landed raw tiles -> tensor core
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&lofull[sl]);
      if (gt == 0) MFAS_KSTAMP(32, n, 3);
    }
  } else if (warp == 12) {
    // ================================ MMA issuer ==================================================
    constexpr uint32_t idesc = umma::idesc_tf32(128, NPAD, false, false), idesc_cat = umma::idesc_tf32(128, 2 * NPAD, false, false);
    int n = 0;
    auto nkb_of = [&](int i) { const FwdItem& it = items[blockIdx.x + i * gridDim.x]; return it.kb1 - it.kb0; };
    int nkb_next = n_my > 0 ? nkb_of(0) : 0;                       // one item ahead: the issue loop never waits for a descriptor
    for (int i = 0; i < n_my && ok; ++i) {
      const int nkb = nkb_next, tb = i & 1;
      if (i + 1 < n_my) nkb_next = nkb_of(i + 1);
      if (!umma::mbar_wait(&tempty[tb], ((i >> 1) & 1) ^ 1)) { ok = false; break; }
      for (int k = 0; k < nkb; ++k, ++n) {
        const int sw = n % WR, sx = n % XS, sl = n % LQ;
        if (!umma::mbar_wait(&lofull[sl], (n / LQ) & 1)) { ok = false; break; }
        if (lane == 0) MFAS_KSTAMP(32, n, 4);
        umma::tc_fence_after();
        if (umma::elect_one()) {
          const uint32_t a_hi = umma::smem_u32(smem) + sw * Cfg::A_BYTES, b_hi = umma::smem_u32(x_base) + sx * Cfg::X_TILE;
          const uint32_t a_lo = umma::smem_u32(lo_base) + sl * Cfg::LO_TILE;
          const uint64_t dah0 = umma::smem_desc(a_hi, 16, 1024), dal0 = umma::smem_desc(a_lo, 16, 1024), dbh0 = umma::smem_desc(b_hi, 16, 1024);
          const uint32_t dt = tm + tb * 2 * NPAD;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t acc = (k > 0 || ks > 0) ? 1u : 0u;
            umma::mma_tf32(dt, dah0 + 2 * ks, dbh0 + 2 * ks, idesc_cat, acc);     // W_hi [x_hi; x_lo]^T -> [main | correction]
            umma::mma_tf32(dt + NPAD, dal0 + 2 * ks, dbh0 + 2 * ks, idesc, 1u);   // W_lo x_hi^T -> correction
          }
          umma::mma_commit(&wfree[sw]);
          umma::mma_commit(&xfree[sx]);
          umma::mma_commit(&lofree[sl]);
          if (k == nkb - 1) umma::mma_commit(&tfull[tb]);
        }
        __syncwarp();
        if (lane == 0) MFAS_KSTAMP(32, n, 5);
      }
    }
  } else if (warp < 17) {
    // ================================ epilogue ====================================================
    const int q = warp & 3;                                        // TMEM lane quarter this warp may read
    struct Ep { long long part_off; int rows_valid, Hp; };
    auto ep_of = [&](int i) { const FwdItem& f = items[blockIdx.x + i * gridDim.x]; return Ep{f.part_off, f.rows_valid, f.Hp}; };
    Ep ep_next = n_my > 0 ? ep_of(0) : Ep{0, 0, 0};
    for (int i = 0; i < n_my; ++i) {
      const Ep it = ep_next;
      if (i + 1 < n_my) ep_next = ep_of(i + 1);
      const int tb = i & 1;
      float4* dst = reinterpret_cast<float4*>(part_base + it.part_off) + (q * 32 + lane);   // part_off includes the row tile
      if (!umma::mbar_wait(&tfull[tb], (i >> 1) & 1)) { ok = false; break; }
      umma::tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < NPAD / 32; ++cc) {
        if (q * 32 >= it.rows_valid) break;                        // lane quarter beyond H (inner_repr 16 / 32 / 64): rows of zeros nobody reads
        float v[32], w[32];
        umma::tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(tb * 2 * NPAD + cc * 32), v);
        umma::tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(tb * 2 * NPAD + NPAD + cc * 32), w);
        float4* d4 = dst + (long long)(cc * 8) * it.Hp;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          d4[(long long)j * it.Hp] = make_float4(v[4 * j] + w[4 * j], v[4 * j + 1] + w[4 * j + 1], v[4 * j + 2] + w[4 * j + 2], v[4 * j + 3] + w[4 * j + 3]);
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&tempty[tb]);
    }
  }
  if (!ok) atomicExch(err.flag, 6);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tm, 4 * NPAD);
}

// ---------------------------------------------------------------------------------------------
// forward, all layers, inner_repr 16 / 32 (the search default): the same persistent, warp-specialised stream as
// k_tc_fwd_ws with the product TRANSPOSED -- z[b, h] = sum_k x[b, k] W[h, k]: A = the gathered x tile (batch rows),
// B = the W tile (N = HN = 16 or 32 rows).  With W as the M operand (k_tc_fwd_ws) 112 or 96 of the 128 MMA rows are zeros,
// yet every k-block moves the full 24 KB tile six times through shared memory: search256 on one B200 spent 174 us per
// step in a forward stream whose bytes take 38 us.
// What a k-block costs here is not bytes but the fixed price of its steps -- ~40 cycles per tcgen05.mma on the issuing
// thread however small the MMA, ~500 cycles for a converter pass (two barrier waits, a shared-memory round trip, a proxy
// fence, an arrive) -- measured per role with clock stamps (tests/cuda/fwd_small_timeline.py,
// profiles/r02dd_fwd_small_roles.txt: 12 MMAs = 518 cycles, a converter pass by all eight warps 1150).  So:
//   * hi and lo tiles of a stage are ADJACENT, [x_hi | x_lo | W_hi | W_lo], and the split products are ONE MMA per 8 columns:
//     64 batch rows:  A' = [x_hi; x_lo] (M = 128), B' = [W_hi; W_lo] (N = 2 HN): D = [hi hi | hi lo ; lo hi | lo lo] --
//                     4 MMAs per k-block instead of 12 (lo lo comes for free); the epilogue adds the two column halves,
//                     and the lo rows (TMEM lanes 64..127) to the hi rows through a 4 KB shared-memory exchange;
//     128 batch rows: x_hi B' (N = 2 HN), then x_lo W_hi into the second half: 8 MMAs per k-block;
//   * the converter warps work in four pairs, each pair a k-block of its own (pair g: k-blocks g, g + 4, ...);
//   * a loader warp signals "landed" once per stage (cp.async groups complete in order: the stage issued LAG k-blocks ago),
//     not once per thread.
// Work items, partial-sum layout (part[item][NPAD/4][Hp][4]: element (b, h) at ((b >> 2) Hp + h) 4 + (b & 3)) and the
// consumers (chain kernels) are unchanged.
// ---------------------------------------------------------------------------------------------
template <int NPAD, int HN> struct FwdSmall {
  static constexpr uint32_t A_BYTES = NPAD * 128, B_BYTES = HN * 128, TILE = 2 * (A_BYTES + B_BYTES);
  static constexpr uint32_t EX_BYTES = NPAD == 64 ? 2 * 64 * (HN + 1) * 4 : 0;       // lo-row sums of the two accumulator buffers
  static constexpr int STAGES = (int)((230000 - EX_BYTES) / TILE);
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * TILE + EX_BYTES;
  static constexpr int LOADERS = 128, CONVERTERS = 256, CGROUPS = 4, THREADS = 18 * 32;     // warps 12 and 17: MMA issuers
};


template <int NPAD, int HN>
__global__ void __launch_bounds__((FwdSmall<NPAD, HN>::THREADS), 1)
k_tc_fwd_small(const FwdItem* __restrict__ items, int n_items, DCache cache, BatchRef batch, float* part_base, TcErr err) {
  using Cfg = FwdSmall<NPAD, HN>;
  constexpr int S = Cfg::STAGES, CG = Cfg::CGROUPS;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);
  __shared__ uint64_t landed[S], split[S], sfree[S], tfull[2], tempty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nrows = batch.n_rows;
  constexpr uint32_t TM_COLS = 8 * HN;                             // two accumulator buffers x two MMA warps x 2 HN columns
  if (warp == 0) umma::tmem_alloc(&tmem_slot, TM_COLS);
  if (tid == 32) {
    for (int i = 0; i < S; ++i) { umma::mbar_init(&landed[i], Cfg::LOADERS); umma::mbar_init(&split[i], Cfg::CONVERTERS / 32 / CG); umma::mbar_init(&sfree[i], 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&tfull[i], 2); umma::mbar_init(&tempty[i], 4); }
    umma::fence_mbar_init();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  griddep_launch();
  const int n_my = (n_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  griddep_wait();                                                  // the weights come from the previous step's backward stream
  const uint32_t tm = tmem_slot;
  bool ok = true;

  if (warp < 4) {
    // ================================ loaders (cp.async) ==========================================
    constexpr int XJ = NPAD / 16;                                  // x rows r + 16 j; W rows r + 16 j for j < HN / 16
    const int r = tid >> 3, c = tid & 7;
    const uint32_t off = (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);   // sw128(r + 16 j, 16 c) = off + 2048 j
    const uint32_t s0 = umma::smem_u32(smem);
    const uint64_t stream_policy = l2_stream_policy((err.l2_hints & 1) != 0);
    const uint64_t x_policy = (err.l2_hints & 8) ? l2_keep_policy() : (err.l2_hints & 4) ? l2_stream_policy(false) : stream_policy;
    auto load_rows = [&](const FwdItem& it, int* rows) {
#pragma unroll
      for (int j = 0; j < XJ; ++j) {
        const int row = r + 16 * j;
        rows[j] = batch_row(batch, it.cand, row < nrows ? row : 0);
      }
    };
    FwdItem cur{}, nxt{};
    int rows_cur[XJ], rows_nxt[XJ];
#pragma unroll
    for (int j = 0; j < XJ; ++j) rows_cur[j] = rows_nxt[j] = 0;
    if (n_my > 0) { cur = items[blockIdx.x]; load_rows(cur, rows_cur); }
    if (n_my > 1) nxt = items[blockIdx.x + gridDim.x];
    int n = 0;
    for (int i = 0; i < n_my && ok; ++i) {
      if (i + 1 < n_my) load_rows(nxt, rows_nxt);
      FwdItem nx2{};
      if (i + 2 < n_my) nx2 = items[blockIdx.x + (i + 2) * gridDim.x];
      const FwdItem& it = cur;
      const long long wstride = 16LL * it.K;
      const float* wp = it.W + (long long)r * it.K + c * 4;        // &W[r][4 c]
      const float* xs[XJ]; const float* xr[XJ];
#pragma unroll
      for (int j = 0; j < XJ; ++j) {
        const long long gr = rows_cur[j];
        xs[j] = cache.ske[it.ske_tap] + gr * cache.ske_ld[it.ske_tap] + c * 4;
        xr[j] = cache.rgb[it.rgb_tap] + gr * cache.rgb_ld[it.rgb_tap] + c * 4 - 32LL * it.fs_kb;
      }
#pragma unroll 1
      for (int kb = it.kb0; kb < it.kb1; ++kb, ++n) {
        const int sg = n % S;
        if (n >= S && !umma::mbar_wait(&sfree[sg], ((n / S) & 1) ^ 1)) { ok = false; break; }
        if (tid == 0) MFAS_KSTAMP(32, n, 7);
        const uint32_t a = s0 + sg * Cfg::TILE + off, b = a + 2 * Cfg::A_BYTES;
        const bool ske = kb < it.fs_kb;
#pragma unroll
        for (int j = 0; j < XJ; ++j) cp_async16_zfill(a + j * 2048, (ske ? xs[j] : xr[j]) + 32LL * kb, r + 16 * j < nrows, x_policy);
        const float* w = wp + 32LL * kb;
#pragma unroll
        for (int j = 0; j < HN / 16; ++j) cp_async16_zfill(b + j * 2048, w + j * wstride, r + 16 * j < it.rows_valid, stream_policy);
        // completion: every thread's cp.async.mbarrier.arrive fires when ITS copies have landed -- the loaders never wait for
        // data, so all S stages can be in flight (a gathered row takes ~3500 cycles under load; a per-warp signal behind
        // cp.async.wait_group held the stream to wait-depth / latency, profiles/r02dk_roles.txt)
        cp_async_arrive_noinc(&landed[sg]);
        if (tid == 0) MFAS_KSTAMP(32, n, 0);
      }
      cur = nxt; nxt = nx2;
#pragma unroll
      for (int j = 0; j < XJ; ++j) rows_cur[j] = rows_nxt[j];
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp < 12) {
    // ================================ converters ==================================================
    // lo = rna_tf32(x - trunc_tf32(x)) of the x tile and of the W tile, into the same stage
    const int ct = tid - 128;
    constexpr int TG = Cfg::CONVERTERS / CG, NA = (int)(Cfg::A_BYTES / 16), NCH = (int)((Cfg::A_BYTES + Cfg::B_BYTES) / 16);
    const int grp = ct / TG, gt = ct % TG;
    int total = 0;
    for (int i = 0; i < n_my; ++i) { const FwdItem& it = items[blockIdx.x + i * gridDim.x]; total += it.kb1 - it.kb0; }
#pragma unroll 1
    for (int n = grp; n < total; n += CG) {
      const int sg = n % S;
      if (!umma::mbar_wait(&landed[sg], (n / S) & 1)) { ok = false; break; }
      if (gt == 0) MFAS_KSTAMP(32, n, 2);
      uint8_t* st = smem + sg * Cfg::TILE;
#pragma unroll
      for (int j = 0; j < (NCH + TG - 1) / TG; ++j) {
        const int ch = gt + j * TG;
        if (ch < NCH) {
          const uint32_t so = ch < NA ? (uint32_t)ch * 16u : 2 * Cfg::A_BYTES + (uint32_t)(ch - NA) * 16u;
          const float4 x = *reinterpret_cast<const float4*>(st + so);
          float4 l;
          l.x = umma::round_tf32(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
          l.y = umma::round_tf32(x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
          l.z = umma::round_tf32(x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
          l.w = umma::round_tf32(x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
          *reinterpret_cast<float4*>(st + so + (ch < NA ? Cfg::A_BYTES : Cfg::B_BYTES)) = l;
        }
      }
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&split[sg]);
      if (gt == 0) MFAS_KSTAMP(32, n, 3);
    }
  } else if (warp == 12 || warp == 17) {
    // ================================ MMA issuers =================================================
    // Two warps, every other k-block of an item each, an accumulator of its own each (the epilogue adds them in a fixed order):
    // the issue loop of one warp -- barrier poll, fence, 4 MMAs, commits: ~600 cycles per k-block -- was the stream's rate.
    constexpr uint32_t idesc_cat = umma::idesc_tf32(128, 2 * HN, false, false), idesc_lo = umma::idesc_tf32(128, HN, false, false);
    const int mw = warp == 12 ? 0 : 1;
    int n0 = 0;
    auto nkb_of = [&](int i) { const FwdItem& it = items[blockIdx.x + i * gridDim.x]; return it.kb1 - it.kb0; };
    int nkb_next = n_my > 0 ? nkb_of(0) : 0;
    for (int i = 0; i < n_my && ok; ++i) {
      const int nkb = nkb_next, tb = i & 1;
      if (i + 1 < n_my) nkb_next = nkb_of(i + 1);
      if (!umma::mbar_wait(&tempty[tb], ((i >> 1) & 1) ^ 1)) { ok = false; break; }
      if (mw >= nkb) { if (lane == 0) umma::mbar_arrive(&tfull[tb]); }      // a one-k-block item: nothing for the second warp
      for (int k = mw; k < nkb; k += 2) {
        const int n = n0 + k, sg = n % S;
        if (!umma::mbar_wait(&split[sg], (n / S) & 1)) { ok = false; break; }
        if (lane == 0) MFAS_KSTAMP(32, n, 4);
        umma::tc_fence_after();
        if (umma::elect_one()) {
          const uint32_t a_hi = umma::smem_u32(smem) + sg * Cfg::TILE, b_hi = a_hi + 2 * Cfg::A_BYTES;
          const uint64_t da0 = umma::smem_desc(a_hi, 16, 1024), db0 = umma::smem_desc(b_hi, 16, 1024);
          const uint32_t dt = tm + tb * 4 * HN + mw * 2 * HN;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t acc = (k > mw || ks > 0) ? 1u : 0u;
            if (NPAD == 64) {
              umma::mma_tf32(dt, da0 + 2 * ks, db0 + 2 * ks, idesc_cat, acc);               // [x_hi; x_lo] [W_hi; W_lo]^T
            } else {
              umma::mma_tf32(dt, da0 + 2 * ks, db0 + 2 * ks, idesc_cat, acc);               // x_hi [W_hi; W_lo]^T
              umma::mma_tf32(dt + HN, da0 + (Cfg::A_BYTES >> 4) + 2 * ks, db0 + 2 * ks, idesc_lo, 1u);   // x_lo W_hi^T -> the second (correction) half
            }
          }
          umma::mma_commit(&sfree[sg]);
          if (k + 2 >= nkb) umma::mma_commit(&tfull[tb]);
        }
        __syncwarp();
        if (lane == 0) MFAS_KSTAMP(32, n, 5);
      }
      n0 += nkb;
    }
  } else {
    // ================================ epilogue ====================================================
    const int q = warp & 3;                                        // TMEM lane quarter this warp may read
    // 64 batch rows: lanes 0..63 hold the x_hi rows, 64..127 the x_lo rows of the same batch rows; 128: lane = batch row
    const int brow = NPAD == 128 ? q * 32 + lane : (q & 1) * 32 + lane;
    float* ex = reinterpret_cast<float*>(smem + (size_t)S * Cfg::TILE);          // [2][64][HN + 1]
    struct Ep { long long part_off; int rows_valid, Hp, nkb; };
    auto ep_of = [&](int i) { const FwdItem& f = items[blockIdx.x + i * gridDim.x]; return Ep{f.part_off, f.rows_valid, f.Hp, f.kb1 - f.kb0}; };
    Ep ep_next = n_my > 0 ? ep_of(0) : Ep{0, 0, 0, 0};
    for (int i = 0; i < n_my; ++i) {
      const Ep it = ep_next;
      if (i + 1 < n_my) ep_next = ep_of(i + 1);
      const int tb = i & 1;
      if (!umma::mbar_wait(&tfull[tb], (i >> 1) & 1)) ok = false;   // (no early exit: the four warps meet at the exchange barrier)
      umma::tc_fence_after();
      float v[HN], w[HN];
      auto ld_acc = [&](uint32_t col, float* o) {
        if (HN == 16) umma::tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + col, o);
        else umma::tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + col, o);
      };
      ld_acc((uint32_t)(tb * 4 * HN), v);                           // first MMA warp: k-blocks 0, 2, ...
      ld_acc((uint32_t)(tb * 4 * HN + HN), w);
#pragma unroll
      for (int h = 0; h < HN; ++h) v[h] += w[h];
      if (it.nkb > 1) {                                             // second MMA warp: k-blocks 1, 3, ...
        float v2[HN];
        ld_acc((uint32_t)(tb * 4 * HN + 2 * HN), v2);
        ld_acc((uint32_t)(tb * 4 * HN + 3 * HN), w);
#pragma unroll
        for (int h = 0; h < HN; ++h) v[h] += v2[h] + w[h];
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&tempty[tb]);
      if (NPAD == 64) {
        float* e = ex + ((size_t)tb * 64 + brow) * (HN + 1);
        if (q >= 2) {
#pragma unroll
          for (int h = 0; h < HN; ++h) e[h] = v[h];
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");              // the four epilogue warps (ex[tb] is next written two items on)
        if (q < 2) {
#pragma unroll
          for (int h = 0; h < HN; ++h) v[h] += e[h];
        }
      }
      if ((NPAD == 128 || q < 2) && brow < nrows) {
        float* dst = part_base + it.part_off + ((long long)(brow >> 2) * it.Hp) * 4 + (brow & 3);
#pragma unroll
        for (int h = 0; h < HN; ++h)
          if (h < it.rows_valid) dst[h * 4] = v[h];
      }
      if (!ok) break;
    }
  }
  if (!ok) atomicExch(err.flag, 6);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tm, TM_COLS);
}

// ---------------------------------------------------------------------------------------------
// forward, one layer: z = (sum of this layer's feature partials, fixed order) + h_{l-1} W[:,hid]^T + b,
// activation, BatchNorm over the batch (train: batch statistics + running-stat update; eval: running
// statistics), dropout.  These kernels sit on the serial chain h_0 -> h_1 -> ..., so they are cut fine
// for latency: grid = (H/8, candidates), one warp per output column, lanes run over batch rows.
// dynamic smem: hp[bmax][H+1] | wh[8][H]
// ---------------------------------------------------------------------------------------------
constexpr int TC_CB = 16;             // output columns per CTA in the per-layer chain kernels (one warp each)
constexpr int TC_CHAIN_THREADS = TC_CB * 32;

template <bool TRAIN, int NPAD>
__global__ void __launch_bounds__(TC_CHAIN_THREADS)
k_fwd_layer(const DCand* __restrict__ cands, int layer, int nrows, int bmax, const float* part_base,
            long long part_stride_cand, uint32_t drop_seed, float drop_p, uint32_t step) {
  extern __shared__ __align__(16) float fsm[];
  const DCand& cd = cands[blockIdx.y];
  if (layer >= cd.L) return;
  const int H = cd.H;
  const int col0 = blockIdx.x * TC_CB;
  if (col0 >= H) return;
  const DLayer& ly = cd.layer[layer];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  int item0 = 0;
  for (int l = 0; l < layer; ++l) item0 += tc_fwd_items(cd.layer[l].d_ske, cd.layer[l].d_rgb, cd.kb_item);
  const int nsplit = tc_fwd_items(ly.d_ske, ly.d_rgb, cd.kb_item);
  const int Hp = ((H + 127) >> 7) << 7;
  const float* part = part_base + (long long)blockIdx.y * part_stride_cand + (long long)item0 * Hp * NPAD;
  const bool bn = (cd.flags & MFAS_FLAG_BN) != 0;
  const bool drop = TRAIN && (cd.flags & MFAS_FLAG_DROPOUT);
  const uint32_t dkey = drop ? dropout_key(drop_seed, (uint32_t)cd.cand_id, step, (uint32_t)layer) : 0u;
  const float dscale = drop ? 1.f / (1.f - drop_p) : 1.f;
  constexpr int NJ = NPAD / 32;
  __shared__ float sa[MFAS_MAX_BATCH][TC_CB + 1], sh[MFAS_MAX_BATCH][TC_CB + 1];
  float* hp = fsm;                               // [nrows][H+1]
  float* wh = fsm + (size_t)bmax * (H + 1);      // [TC_CB][H]
  const bool has_hid = layer > 0;
  const int c = col0 + warp;                     // this warp's output column (H % 8 == 0)

  // split-K partials first: their loads overlap the staging of h_{l-1} below
  float z[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) z[j] = 0.f;
  for (int s = 0; s < nsplit; ++s) {
    const float* p = part + (long long)s * Hp * NPAD + (long long)c * 4;
#pragma unroll
    for (int j = 0; j < NJ; ++j) { const int b = lane + 32 * j; z[j] += p[(long long)(b >> 2) * Hp * 4 + (b & 3)]; }
  }
  if (has_hid) {
    const float* hprev = cd.hid + (long long)(layer - 1) * bmax * H;
    {   // h_{l-1} -> smem, 8 independent loads in flight per thread
      const int total = nrows * H;
#pragma unroll 1
      for (int i0 = tid; i0 < total; i0 += TC_CHAIN_THREADS * 8) {
        float t[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * TC_CHAIN_THREADS; t[u] = i < total ? hprev[i] : 0.f; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + u * TC_CHAIN_THREADS; if (i < total) hp[(i / H) * (H + 1) + (i % H)] = t[u]; }
      }
    }
    const float* Wh = cd.p + ly.oW + ly.d_ske + ly.d_rgb + (long long)c * ly.K;
    for (int j = lane; j < H; j += 32) wh[warp * H + j] = Wh[j];
    __syncthreads();
    const float* w = wh + warp * H;
    float zh[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) zh[j] = 0.f;
    for (int jj = 0; jj < H; ++jj) {
      const float wv = w[jj];
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        if (lane + 32 * j < nrows) zh[j] = fmaf(hp[(lane + 32 * j) * (H + 1) + jj], wv, zh[j]);
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) z[j] += zh[j];
  }
  const float bias = cd.p[ly.ob + c];
  float a[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) a[j] = (lane + 32 * j < nrows) ? act_fwd(z[j] + bias, ly.act) : 0.f;
  float mean = 0.f, var = 1.f, istd = 1.f, gamma = 1.f, beta = 0.f;
  if (bn) {
    if (TRAIN) {
      float s1 = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) s1 += a[j];
      mean = warp_sum(s1) / (float)nrows;
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) if (lane + 32 * j < nrows) { const float d = a[j] - mean; q = fmaf(d, d, q); }
      var = warp_sum(q) / (float)nrows;
    } else {
      mean = cd.bufs[ly.orm + c];
      var = cd.bufs[ly.orv + c];
    }
    istd = 1.f / sqrtf(var + kBnEps);
    gamma = cd.p[ly.og + c]; beta = cd.p[ly.obe + c];
    if (TRAIN && lane == 0) {
      cd.mu[layer * H + c] = mean;
      cd.invstd[layer * H + c] = istd;
      const float n = (float)nrows;
      float& rm = cd.bufs[ly.orm + c];
      float& rv = cd.bufs[ly.orv + c];
      rm = (1.f - kBnMomentum) * rm + kBnMomentum * mean;
      rv = (1.f - kBnMomentum) * rv + kBnMomentum * (var * (n / (n - 1.f)));
      if (c == 0) cd.nbt[layer] += 1;
    }
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int r = lane + 32 * j;
    if (r < nrows) {
      float h = bn ? (a[j] - mean) * istd * gamma + beta : a[j];
      if (drop) h = dropout_keep(dkey, (uint32_t)(r * H + c), drop_p) ? h * dscale : 0.f;
      sa[r][warp] = a[j];
      sh[r][warp] = h;
    }
  }
  __syncthreads();
  float* actp = cd.act + (long long)layer * bmax * H;       // write-out: 16 columns = two 32-byte sectors per row
  float* hidp = cd.hid + (long long)layer * bmax * H;
  for (int r = tid / TC_CB; r < nrows; r += TC_CHAIN_THREADS / TC_CB) {
    const int cc = tid % TC_CB;
    if (TRAIN) actp[r * H + col0 + cc] = sa[r][cc];
    hidp[r * H + col0 + cc] = sh[r][cc];
  }
}

// ---------------------------------------------------------------------------------------------
// backward chain, one layer: dh_l = dz_{l+1} W_{l+1}[:, hidden columns] (pre-update weights; for the
// last layer dh comes from k_head), then dropout mask, BatchNorm backward, activation backward ->
// dz_l (kept per layer for k_tc_bwd_all), plus gradients + Adam of the per-column vectors b, gamma, beta.
// grid = (H/8, candidates); CTA = 8 columns x 32 row groups.  dynamic smem: dzn[bmax][H] | wt[H][9]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_dzx(const DCand* __restrict__ cands, int layer, int nrows, int bmax, AdamH adam, float step_size, float bc2_sqrt,
      uint32_t drop_seed, float drop_p, uint32_t step) {
  extern __shared__ __align__(16) float dsm[];
  const DCand& cd = cands[blockIdx.y];
  if (layer >= cd.L) return;
  const int H = cd.H;
  if (blockIdx.x * TC_CB >= H) return;
  const int tid = threadIdx.x, col = tid % TC_CB, rg = tid / TC_CB;   // NRG row groups: rows rg, rg+NRG, ...
  constexpr int NRG = kThreads / TC_CB, NI = MFAS_MAX_BATCH / NRG;
  const int c = blockIdx.x * TC_CB + col;
  const DLayer& ly = cd.layer[layer];
  const bool bn = (cd.flags & MFAS_FLAG_BN) != 0;
  const bool drop = (cd.flags & MFAS_FLAG_DROPOUT) != 0;
  const uint32_t dkey = drop ? dropout_key(drop_seed, (uint32_t)cd.cand_id, step, (uint32_t)layer) : 0u;
  const float dscale = drop ? 1.f / (1.f - drop_p) : 1.f;
  const float* actp = cd.act + (long long)layer * bmax * H;
  float* dzl = cd.dzs + (long long)layer * bmax * H;
  __shared__ float red1[NRG][TC_CB + 1], red2[NRG][TC_CB + 1];
  float dhv[NI];                                                  // dh of rows rg + NRG*i
#pragma unroll
  for (int i = 0; i < NI; ++i) dhv[i] = 0.f;

  if (layer + 1 < cd.L) {
    // dh_l[b][c] = sum_h dz_{l+1}[b][h] * W_{l+1}[h][fs+fr+c]
    const DLayer& up = cd.layer[layer + 1];
    float* dzn = dsm;                          // [nrows][H]
    float* wt = dsm + (size_t)bmax * H;        // [H][9]
    const float* dzg = cd.dzs + (long long)(layer + 1) * bmax * H;
    {   // dz_{l+1} and the hidden columns of W_{l+1} -> smem, loads batched (latency is paid once)
      const int total = nrows * H;
#pragma unroll 1
      for (int i0 = tid * 4; i0 < total; i0 += kThreads * 4 * 4) {
        float4 t[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + u * kThreads * 4; t[u] = i < total ? *reinterpret_cast<const float4*>(dzg + i) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + u * kThreads * 4; if (i < total) *reinterpret_cast<float4*>(dzn + i) = t[u]; }
      }
      const float* Wg = cd.p + up.oW + up.d_ske + up.d_rgb + blockIdx.x * TC_CB;
      float wv[MFAS_MAX_HIDDEN / NRG];
#pragma unroll
      for (int u = 0; u < MFAS_MAX_HIDDEN / NRG; ++u) { const int h = rg + NRG * u; wv[u] = h < H ? Wg[(long long)h * up.K + col] : 0.f; }
#pragma unroll
      for (int u = 0; u < MFAS_MAX_HIDDEN / NRG; ++u) { const int h = rg + NRG * u; if (h < H) wt[h * (TC_CB + 1) + col] = wv[u]; }
    }
    __syncthreads();
    for (int h = 0; h < H; h += 4) {
      const float w0 = wt[h * (TC_CB + 1) + col], w1 = wt[(h + 1) * (TC_CB + 1) + col], w2 = wt[(h + 2) * (TC_CB + 1) + col], w3 = wt[(h + 3) * (TC_CB + 1) + col];
#pragma unroll
      for (int i = 0; i < NI; ++i) {
        const int r = rg + NRG * i;
        if (r < nrows) {
          const float4 d = *reinterpret_cast<const float4*>(dzn + (size_t)r * H + h);
          dhv[i] = fmaf(d.x, w0, dhv[i]); dhv[i] = fmaf(d.y, w1, dhv[i]);
          dhv[i] = fmaf(d.z, w2, dhv[i]); dhv[i] = fmaf(d.w, w3, dhv[i]);
        }
      }
    }
  } else {
    const float* dhp = cd.dh + (long long)layer * bmax * H;
#pragma unroll
    for (int i = 0; i < NI; ++i) if (rg + NRG * i < nrows) dhv[i] = dhp[(rg + NRG * i) * H + c];
  }
  if (drop) {
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      const int r = rg + NRG * i;
      if (r < nrows) dhv[i] = dropout_keep(dkey, (uint32_t)(r * H + c), drop_p) ? dhv[i] * dscale : 0.f;
    }
  }
  float av[NI];
#pragma unroll
  for (int i = 0; i < NI; ++i) av[i] = (rg + NRG * i < nrows) ? actp[(rg + NRG * i) * H + c] : 0.f;

  float mu = 0.f, istd = 1.f, gam = 1.f, m1 = 0.f, m2 = 0.f, S1 = 0.f, S2 = 0.f;
  if (bn) {
    mu = cd.mu[layer * H + c]; istd = cd.invstd[layer * H + c]; gam = cd.p[ly.og + c];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i)
      if (rg + NRG * i < nrows) { s1 += dhv[i]; s2 = fmaf(dhv[i], (av[i] - mu) * istd, s2); }
    red1[rg][col] = s1; red2[rg][col] = s2;
    __syncthreads();
#pragma unroll
    for (int g = 0; g < NRG; ++g) { S1 += red1[g][col]; S2 += red2[g][col]; }
    m1 = S1 / (float)nrows; m2 = S2 / (float)nrows;
    __syncthreads();
  }
  float sdz = 0.f;
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int r = rg + NRG * i;
    if (r < nrows) {
      float da = dhv[i];
      if (bn) { const float ah = (av[i] - mu) * istd; da = gam * istd * (dhv[i] - m1 - ah * m2); }
      const float dz = da * act_bwd(av[i], ly.act);
      dzl[r * H + c] = dz;
      sdz += dz;
    }
  }
  red1[rg][col] = sdz;
  __syncthreads();
  if (rg == 0) {
    float db = 0.f;
#pragma unroll
    for (int g = 0; g < NRG; ++g) db += red1[g][col];
    auto upd = [&](long long o, float g) {
      if (cd.grad) cd.grad[o] = g;
      float p = cd.p[o], m = cd.m[o], v = cd.v[o];
      adam_update(g, p, m, v, adam, step_size, bc2_sqrt);
      cd.p[o] = p; cd.m[o] = m; cd.v[o] = v;
    };
    upd(ly.ob + c, db);
    if (bn) { upd(ly.og + c, S2); upd(ly.obe + c, S1); }
  }
}

// 16-byte cp.async without a cache hint; src-size 0 zero-fills the destination and reads nothing
__device__ __forceinline__ void cp_async16_zf(uint32_t dst_smem, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// lo = rna_tf32(x - trunc_tf32(x)) of one 16-byte chunk of a raw operand tile (the raw tile itself is the hi operand:
// kind::tf32 reads the top 19 bits of each container), written to the same offset of the lo tile
__device__ __forceinline__ void lo_of_chunk(const uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off) {
  const float4 x = *reinterpret_cast<const float4*>(hi_tile + off);
  float4 l;
  l.x = umma::round_tf32(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
  l.y = umma::round_tf32(x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
  l.z = umma::round_tf32(x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
  l.w = umma::round_tf32(x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// ---------------------------------------------------------------------------------------------
// Chain kernels on the tensor core.  The per-layer steps of the serial chain (hidden-state GEMM +
// BatchNorm forward, and its mirror image backward) are tiny, so what matters is latency: one CTA per
// (candidate, 128 output columns) does the whole thing.  The MMA is oriented so that a TMEM lane is an
// output column: each epilogue thread then owns ALL batch rows of its column in registers, and the
// BatchNorm reductions over the batch are plain sequential sums inside a thread (deterministic, no
// shuffles, no shared memory), while global reads/writes stay coalesced across the 32 lanes of a warp.
//
// k_chain_fwd:  zT[c,b] = sum_j W_l[c, hid j] h_{l-1}[b, j]  (+ feature partials + bias) -> act -> BN
//               A = W_l hidden columns (K-major), B = h_{l-1} (K-major); K = H in passes of 128
// dynamic smem (1024-aligned), per pass of <= 4 k-blocks: A_hi | A_lo (4 x 16 KB) | B_hi | B_lo (4 x NPAD*128 B)
// ---------------------------------------------------------------------------------------------
template <int NPAD> struct ChainCfg {
  static constexpr int THREADS = 512;                     // 16 warps: 4 per TMEM lane quarter
  static constexpr int NB = NPAD / 4;                     // batch rows per epilogue thread
  static constexpr int PASS = NPAD == 64 ? 128 : 64;      // K columns staged per pass (smem budget)
  static constexpr int NKB = PASS / 32;
  static constexpr size_t SMEM = 1024 + 2 * (size_t)NKB * 16384 + 2 * (size_t)NKB * NPAD * 128;
};

// Shared state of the chain kernels: operand tiles, the MMA-completion mbarrier and its running phase, the TMEM
// accumulator.  One layer step is a device function so that the per-layer kernels and the fused k_chain_all
// (every layer forward, the head, every layer backward in ONE launch per step) run the same code.
struct ChainCtx {
  uint8_t* smem;          // 1024-aligned dynamic shared memory
  uint64_t* bar;
  int* ok_flag;
  uint32_t tm;            // TMEM accumulator (NPAD columns)
  uint32_t phase;
  bool ok;
  long long* tl;          // debug timeline slots 10.. of this candidate (or null)
  int tli, tl_layer;      // next slot; the forward layer whose inner phases are stamped
};
__device__ __forceinline__ void chain_stamp(ChainCtx& cx, int layer) {
  if (cx.tl && layer == cx.tl_layer && threadIdx.x == 0 && cx.tli < 16) cx.tl[cx.tli++] = clock64();
}

template <int NPAD>
__device__ __forceinline__ void chain_ctx_open(ChainCtx& cx, uint8_t* smem_raw, uint64_t* bar, uint32_t* tmem_slot, int* ok_flag) {
  const int tid = threadIdx.x, warp = tid >> 5;
  cx.smem = umma::align1024(smem_raw);
  cx.bar = bar; cx.ok_flag = ok_flag; cx.phase = 0; cx.ok = true; cx.tl = nullptr; cx.tli = 10; cx.tl_layer = 0;
  if (warp == 0) umma::tmem_alloc(tmem_slot, NPAD);
  if (tid == 0) { umma::mbar_init(bar, 1); umma::fence_mbar_init(); }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  cx.tm = *tmem_slot;
}
template <int NPAD>
__device__ __forceinline__ void chain_ctx_close(ChainCtx& cx) {
  umma::tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == 0) umma::tmem_free(cx.tm, NPAD);
}
__device__ __forceinline__ void chain_wait(ChainCtx& cx) {      // wait for the MMAs committed last; flips the phase
  if (cx.ok) cx.ok = umma::cta_wait(cx.bar, cx.phase, cx.ok_flag);
  cx.phase ^= 1;
}

// accT[m, b] = sum_{j < Kd} A[m, j] X[b, j]: stage (hi/lo split) and issue, in passes of PASS columns of j.
//   A: rows m < rows_valid of a row-major matrix with row stride ldA (rows beyond are zero), X: batch rows b < nrows, stride ldX
// Both operands K-major SW128; the caller waits for the last commit with chain_wait().  Kd in {16, 32} or Kd % 64 == 0.
template <int NPAD>
__device__ __forceinline__ void chain_mma_kmajor(ChainCtx& cx, const float* A, long long ldA, int rows_valid,
                                                 const float* X, int ldX, int Kd, int nrows, int stamp_layer = -1) {
  constexpr int THREADS = ChainCfg<NPAD>::THREADS;
  constexpr uint32_t A_KB = 16384, B_KB = NPAD * 128;
  constexpr int PASS = ChainCfg<NPAD>::PASS, NKB = ChainCfg<NPAD>::NKB;
  const int tid = threadIdx.x;
  uint8_t* a_hi = cx.smem;
  uint8_t* a_lo = a_hi + NKB * A_KB;
  uint8_t* b_hi = a_lo + NKB * A_KB;
  uint8_t* b_lo = b_hi + NKB * B_KB;
  const uint32_t tm = cx.tm;
  constexpr uint32_t idesc = umma::idesc_tf32(128, NPAD, false, false);
  for (int j0 = 0; j0 < Kd; j0 += PASS) {                             // passes of <= NKB k-blocks
    const int jw = min(PASS, Kd - j0), f4 = jw >> 2;                  // float4 per row in this pass
    const int fsh = 31 - __clz(f4);                                   // jw in {16, 32, 64, 128}: f4 is a power of two -- no integer divisions in the staging loops
    if (j0 > 0) { chain_wait(cx); if (!cx.ok) break; }
    // Staging: every 16-byte chunk of both operands goes global -> shared with cp.async (all of a thread's chunks in
    // flight at once: ONE memory round trip per pass instead of three dependent LDG -> split -> STS rounds, r01o
    // timeline: 8 k cycles), landing as the hi tiles; the thread then derives the lo chunks of what it fetched.
    const uint32_t a_hi_u = umma::smem_u32(a_hi), b_hi_u = umma::smem_u32(b_hi);
    for (int i = tid; i < 128 * f4; i += THREADS) {                   // A: 128 rows
      const int r = i >> fsh, c4 = i & (f4 - 1);
      const bool v = r < rows_valid;
      cp_async16_zf(a_hi_u + (uint32_t)(c4 >> 3) * A_KB + umma::sw128(r, (c4 & 7) * 16), v ? A + (long long)r * ldA + j0 + c4 * 4 : A, v);
    }
    for (int i = tid; i < NPAD * f4; i += THREADS) {                  // B: batch rows
      const int r = i >> fsh, c4 = i & (f4 - 1);
      const bool v = r < nrows;
      cp_async16_zf(b_hi_u + (uint32_t)(c4 >> 3) * B_KB + umma::sw128(r, (c4 & 7) * 16), v ? X + (long long)r * ldX + j0 + c4 * 4 : X, v);
    }
    cp_async_wait_all();
    chain_stamp(cx, stamp_layer);    // 11: raw tiles landed
    for (int i = tid; i < 128 * f4; i += THREADS) {
      const int r = i >> fsh, c4 = i & (f4 - 1);
      lo_of_chunk(a_hi, a_lo, (uint32_t)(c4 >> 3) * A_KB + umma::sw128(r, (c4 & 7) * 16));
    }
    for (int i = tid; i < NPAD * f4; i += THREADS) {
      const int r = i >> fsh, c4 = i & (f4 - 1);
      lo_of_chunk(b_hi, b_lo, (uint32_t)(c4 >> 3) * B_KB + umma::sw128(r, (c4 & 7) * 16));
    }
    chain_stamp(cx, stamp_layer);    // 12: B staged
    umma::fence_async_smem();
    __syncthreads();
    if (tid < 32 && umma::elect_one()) {
      // One thread issues on behalf of the CTA, so its instruction stream IS the critical path here (r01 timeline:
      // 4.9 k cycles for 64 MMAs behind `tid == 0`, i.e. an ELECT / R2UR / BRA.U.ANY loop around every MMA): elect.sync
      // leader, compile-time offsets in a fully unrolled loop, and three products (lo x lo is 2^-22 of the result).
      umma::tc_fence_after();
      const uint64_t dah0 = umma::smem_desc(umma::smem_u32(a_hi), 16, 1024), dal0 = umma::smem_desc(umma::smem_u32(a_lo), 16, 1024);
      const uint64_t dbh0 = umma::smem_desc(umma::smem_u32(b_hi), 16, 1024), dbl0 = umma::smem_desc(umma::smem_u32(b_lo), 16, 1024);
      const int nks = jw >> 3;
#pragma unroll
      for (int ks = 0; ks < PASS / 8; ++ks) {
        if (ks < nks) {
          const uint64_t oa = ((uint32_t)(ks >> 2) * A_KB + (ks & 3) * 32u) >> 4, ob = ((uint32_t)(ks >> 2) * B_KB + (ks & 3) * 32u) >> 4;
          umma::mma_tf32(tm, dal0 + oa, dbh0 + ob, idesc, (j0 > 0 || ks > 0) ? 1u : 0u);
          umma::mma_tf32(tm, dah0 + oa, dbl0 + ob, idesc, 1u);
          umma::mma_tf32(tm, dah0 + oa, dbh0 + ob, idesc, 1u);
        }
      }
      umma::mma_commit(cx.bar);
    }
    chain_stamp(cx, stamp_layer);    // 13: MMAs issued
  }
}

template <bool TRAIN, int NPAD>
__device__ __forceinline__ void chain_fwd_layer(ChainCtx& cx, const DCand& cd, int cand, int layer, int m0, int nrows, int bmax,
                                                const float* part_base, long long part_stride_cand, uint32_t drop_seed,
                                                float drop_p, uint32_t step) {
  const int H = cd.H;
  const DLayer& ly = cd.layer[layer];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool has_hid = layer > 0;
  const uint32_t tm = cx.tm;
  chain_stamp(cx, layer);          // 10: layer entered
  // epilogue mapping: thread = (output column c = m0 + 32q + lane, batch rows [cg*NB, cg*NB+NB)), q = warp%4 (the TMEM
  // lane quarter this warp may read), cg = warp/4.  The per-column vectors are requested first: the chain is a string of
  // dependent HBM/L2 round trips (r01 timeline: 2-3 k cycles each), so every load that does not depend on the chain is
  // issued at the top of its phase and lands while the operands are staged and the MMAs run.
  constexpr int NB = ChainCfg<NPAD>::NB;
  const int q = warp & 3, cg = warp >> 2, cl = q * 32 + lane, b0 = cg * NB;
  const int c = m0 + cl;
  const bool mine = c < H;
  const bool bn = (cd.flags & MFAS_FLAG_BN) != 0;
  const bool drop = TRAIN && (cd.flags & MFAS_FLAG_DROPOUT);
  float bias = 0.f, gamma = 1.f, beta = 0.f, rm0 = 0.f, rv0 = 1.f;
  long long nbt0 = 0;
  if (TRAIN && bn && c == 0 && cg == 0) nbt0 = cd.nbt[layer];
  if (mine) {
    bias = cd.p[ly.ob + c];
    if (bn) {
      gamma = cd.p[ly.og + c]; beta = cd.p[ly.obe + c];
      if (!TRAIN || cg == 0) { rm0 = cd.bufs[ly.orm + c]; rv0 = cd.bufs[ly.orv + c]; }
    }
  }

  if (has_hid) {
    const float* Wh = cd.p + ly.oW + ly.d_ske + ly.d_rgb;              // hidden columns of W_l
    const float* hprev = cd.hid + (long long)(layer - 1) * bmax * H;
    chain_mma_kmajor<NPAD>(cx, Wh + (long long)m0 * ly.K, ly.K, H - m0, hprev, H, H, nrows, layer);
  }

  // ---- epilogue.  BatchNorm sums: per-thread partial, then a fixed-order combine of the 4 row groups through
  //      shared memory (deterministic).
  __shared__ float red[4][128];
  float z[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) z[b] = 0.f;
  if (mine) {      // feature partials of this layer: all requested at once (they fly while the MMAs run), summed in split order
    // alpha gates (aux_models.py:103-111: ske * sigmoid(alpha), rgb * (1 - sigmoid(alpha))): the items of a gated candidate are
    // cut at the modality boundary, so the gate is one factor per partial sum
    const bool gated = (cd.flags & MFAS_FLAG_ALPHAS) != 0;
    const float sg = gated ? gate_of(cd.p[ly.oalpha]) : 1.f;
    const int n_ske = tc_fwd_items_of(ly.d_ske, cd.kb_item);
    int item0 = 0;
    for (int l = 0; l < layer; ++l) item0 += tc_fwd_items_g(cd.layer[l].d_ske, cd.layer[l].d_rgb, gated, cd.kb_item);
    const int nsplit = tc_fwd_items_g(ly.d_ske, ly.d_rgb, gated, cd.kb_item);
    const int Hp = ((H + 127) >> 7) << 7;
    const float4* part = reinterpret_cast<const float4*>(part_base + (long long)cand * part_stride_cand + (long long)item0 * Hp * NPAD) +
                         (long long)(b0 >> 2) * Hp + c;
    constexpr int SU = NB == 16 ? 3 : 1;           // splits in flight per round (K_feat <= 3072 -> nsplit <= 3)
    for (int s0 = 0; s0 < nsplit; s0 += SU) {
      float4 v[SU][NB / 4];
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const float4* p4 = part + (long long)min(s0 + u, nsplit - 1) * Hp * (NPAD / 4);
#pragma unroll
        for (int k = 0; k < NB / 4; ++k) v[u][k] = p4[(long long)k * Hp];
      }
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        if (s0 + u < nsplit) {
          if (gated) {
            const float gt = s0 + u < n_ske ? sg : 1.0f - sg;
#pragma unroll
            for (int k = 0; k < NB / 4; ++k) {
              z[4 * k] = fmaf(gt, v[u][k].x, z[4 * k]); z[4 * k + 1] = fmaf(gt, v[u][k].y, z[4 * k + 1]);
              z[4 * k + 2] = fmaf(gt, v[u][k].z, z[4 * k + 2]); z[4 * k + 3] = fmaf(gt, v[u][k].w, z[4 * k + 3]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < NB / 4; ++k) { z[4 * k] += v[u][k].x; z[4 * k + 1] += v[u][k].y; z[4 * k + 2] += v[u][k].z; z[4 * k + 3] += v[u][k].w; }
          }
        }
      }
    }
  }
  if (has_hid) {
    chain_wait(cx);
    umma::tc_fence_after();
    float v[NB];
    if (NB == 16) umma::tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + b0, v);
    else umma::tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + b0, v);
#pragma unroll
    for (int b = 0; b < NB; ++b) z[b] += v[b];
  }
#pragma unroll
  for (int b = 0; b < NB; ++b) z[b] = (mine && b0 + b < nrows) ? act_fwd(z[b] + bias, ly.act) : 0.f;   // z <- a = phi(z)
  chain_stamp(cx, layer);          // 14: partials + MMA result + bias + activation
  float mean = 0.f, var = 1.f, istd = 1.f;
  if (bn) {
    if (TRAIN) {
      float s1 = 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b) s1 += z[b];
      red[cg][cl] = s1;
      __syncthreads();
      mean = (red[0][cl] + red[1][cl] + red[2][cl] + red[3][cl]) / (float)nrows;
      __syncthreads();
      float qq = 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b) if (b0 + b < nrows) { const float d = z[b] - mean; qq = fmaf(d, d, qq); }
      red[cg][cl] = qq;
      __syncthreads();
      var = (red[0][cl] + red[1][cl] + red[2][cl] + red[3][cl]) / (float)nrows;
    } else if (mine) {
      mean = rm0;
      var = rv0;
    }
    istd = 1.f / sqrtf(var + kBnEps);
    if (TRAIN && mine && cg == 0) {
      cd.mu[layer * H + c] = mean;
      cd.invstd[layer * H + c] = istd;
      const float n = (float)nrows;
      cd.bufs[ly.orm + c] = (1.f - kBnMomentum) * rm0 + kBnMomentum * mean;
      cd.bufs[ly.orv + c] = (1.f - kBnMomentum) * rv0 + kBnMomentum * (var * (n / (n - 1.f)));
      if (c == 0) cd.nbt[layer] = nbt0 + 1;
    }
  }
  chain_stamp(cx, layer);          // 15: BatchNorm statistics
  if (mine) {
    const uint32_t dkey = drop ? dropout_key(drop_seed, (uint32_t)cd.cand_id, step, (uint32_t)layer) : 0u;
    const float dscale = drop ? 1.f / (1.f - drop_p) : 1.f;
    float* actp = cd.act + ((long long)layer * bmax + b0) * H + c;     // lanes = consecutive columns: coalesced rows
    float* hidp = cd.hid + ((long long)layer * bmax + b0) * H + c;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (b0 + b < nrows) {
        float h = bn ? (z[b] - mean) * istd * gamma + beta : z[b];
        if (drop) h = dropout_keep(dkey, (uint32_t)((b0 + b) * H + c), drop_p) ? h * dscale : 0.f;
        if (TRAIN) *actp = z[b];
        *hidp = h;
      }
      actp += H; hidp += H;
    }
  }
}

template <bool TRAIN, int NPAD>
__global__ void __launch_bounds__(ChainCfg<NPAD>::THREADS)
k_chain_fwd(const DCand* __restrict__ cands, int layer, int nrows, int bmax, const float* part_base,
            long long part_stride_cand, uint32_t drop_seed, float drop_p, uint32_t step, TcErr err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int ok_flag;
  const DCand& cd = cands[blockIdx.y];
  if (layer >= cd.L || (int)blockIdx.x * 128 >= cd.H) return;
  ChainCtx cx;
  chain_ctx_open<NPAD>(cx, smem_raw, &bar, &tmem_slot, &ok_flag);
  chain_fwd_layer<TRAIN, NPAD>(cx, cd, blockIdx.y, layer, blockIdx.x * 128, nrows, bmax, part_base, part_stride_cand, drop_seed, drop_p, step);
  if (!cx.ok && threadIdx.x == 0) atomicExch(err.flag, 3);
  chain_ctx_close<NPAD>(cx);
}


// ---------------------------------------------------------------------------------------------
// k_chain_bwd: dhT_l[c,b] = sum_h W_{l+1}[h, hid c] dz_{l+1}[b,h]  (pre-update weights; the last layer
//              takes dh from k_head) -> dropout mask -> BatchNorm backward -> activation backward -> dz_l,
//              plus gradients and Adam of the per-column vectors b_l, gamma_l, beta_l.
//              A = hidden columns of W_{l+1} (MN-major: one smem row per h), B = dz_{l+1} (K-major)
// dynamic smem (1024-aligned), per pass of PASS h: A_hi | A_lo (4 blocks x PASS rows x 128 B) | B_hi | B_lo (NKB x NPAD*128 B)
// ---------------------------------------------------------------------------------------------
// accT[c, b] = sum_{h < Kd} U[h, c] G[b, h]: stage (hi/lo split) and issue, in passes of PASS rows h.
//   U: rows h < k_valid (zero beyond) of a row-major matrix with row stride ldU, columns c < mw (zero beyond);
//   G: batch rows b < nrows with row stride ldG, readable (and zero) up to column Kd.   Kd in {16, 32} or Kd % 64 == 0.
// A = U as an MN-major tile (one smem row per h), B = G K-major; the caller waits for the last commit with chain_wait().
template <int NPAD>
__device__ __forceinline__ void chain_mma_mnmajor(ChainCtx& cx, const float* U, long long ldU, int k_valid, int Kd, int mw,
                                                  const float* G, int ldG, int nrows) {
  constexpr int THREADS = ChainCfg<NPAD>::THREADS;
  constexpr int PASS = ChainCfg<NPAD>::PASS, NKB = ChainCfg<NPAD>::NKB;
  constexpr uint32_t A_BLK = PASS * 128, B_KB = NPAD * 128;      // A block = 32 columns x PASS h rows
  const int tid = threadIdx.x;
  uint8_t* a_hi = cx.smem;
  uint8_t* a_lo = a_hi + 4 * A_BLK;
  uint8_t* b_hi = a_lo + 4 * A_BLK;
  uint8_t* b_lo = b_hi + NKB * B_KB;
  const uint32_t tm = cx.tm;
  constexpr uint32_t idesc = umma::idesc_tf32(128, NPAD, true, false);
  for (int h0 = 0; h0 < Kd; h0 += PASS) {                            // passes of PASS rows of U
    const int hw = min(PASS, Kd - h0);
    if (h0 > 0) { chain_wait(cx); if (!cx.ok) break; }
    const uint32_t a_hi_u = umma::smem_u32(a_hi), b_hi_u = umma::smem_u32(b_hi);      // (staging: see chain_mma_kmajor)
    const int f4 = hw >> 2, fsh = 31 - __clz(f4);
    for (int i = tid; i < hw * 32; i += THREADS) {                   // A: hw rows (h) x 32 float4 (128 columns)
      const int r = i >> 5, c4 = i & 31;
      const bool v = c4 * 4 < mw && h0 + r < k_valid;
      cp_async16_zf(a_hi_u + (uint32_t)(c4 >> 3) * A_BLK + umma::sw128_b32(r, (c4 & 7) * 16), v ? U + (long long)(h0 + r) * ldU + c4 * 4 : U, v);
    }
    for (int i = tid; i < NPAD * f4; i += THREADS) {                 // B: batch rows of G[:, h0:h0+hw]
      const int r = i >> fsh, c4 = i & (f4 - 1);
      const bool v = r < nrows;
      cp_async16_zf(b_hi_u + (uint32_t)(c4 >> 3) * B_KB + umma::sw128(r, (c4 & 7) * 16), v ? G + (long long)r * ldG + h0 + c4 * 4 : G, v);
    }
    cp_async_wait_all();
    for (int i = tid; i < hw * 32; i += THREADS) {
      const int r = i >> 5, c4 = i & 31;
      lo_of_chunk(a_hi, a_lo, (uint32_t)(c4 >> 3) * A_BLK + umma::sw128_b32(r, (c4 & 7) * 16));
    }
    for (int i = tid; i < NPAD * f4; i += THREADS) {
      const int r = i >> fsh, c4 = i & (f4 - 1);
      lo_of_chunk(b_hi, b_lo, (uint32_t)(c4 >> 3) * B_KB + umma::sw128(r, (c4 & 7) * 16));
    }
    umma::fence_async_smem();
    __syncthreads();
    if (tid < 32 && umma::elect_one()) {                              // (see chain_mma_kmajor)
      umma::tc_fence_after();
      const uint64_t dah0 = umma::smem_desc(umma::smem_u32(a_hi), A_BLK, 512, umma::kLayoutSw128Base32);
      const uint64_t dal0 = umma::smem_desc(umma::smem_u32(a_lo), A_BLK, 512, umma::kLayoutSw128Base32);
      const uint64_t dbh0 = umma::smem_desc(umma::smem_u32(b_hi), 16, 1024), dbl0 = umma::smem_desc(umma::smem_u32(b_lo), 16, 1024);
      const int nks = hw >> 3;
#pragma unroll
      for (int ks = 0; ks < PASS / 8; ++ks) {
        if (ks < nks) {
          const uint64_t oa = (ks * 1024u) >> 4, ob = ((uint32_t)(ks >> 2) * B_KB + (ks & 3) * 32u) >> 4;
          umma::mma_tf32(tm, dal0 + oa, dbh0 + ob, idesc, (h0 > 0 || ks > 0) ? 1u : 0u);
          umma::mma_tf32(tm, dah0 + oa, dbl0 + ob, idesc, 1u);
          umma::mma_tf32(tm, dah0 + oa, dbh0 + ob, idesc, 1u);
        }
      }
      umma::mma_commit(cx.bar);
    }
  }
}

constexpr int TC_DLOG_LD = 64;        // row stride of DCand::dlog (dL/dlogits, zero-padded): the K extent of the classifier "layer"

// head_up: the classifier is the "upper layer" of the last fusion step (dh_L = dlogits W_c on the tensor core, with
// dlogits in cd.dlog as k_chain_all's head leaves it); otherwise the last layer takes dh from k_head.
template <int NPAD>
__device__ __forceinline__ void chain_bwd_layer(ChainCtx& cx, const DCand& cd, int layer, int m0, int nrows, int bmax, AdamH adam,
                                                float step_size, float bc2_sqrt, uint32_t drop_seed, float drop_p, uint32_t step,
                                                bool head_up = false) {
  const int H = cd.H;
  const DLayer& ly = cd.layer[layer];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool inner = layer + 1 < cd.L;
  const bool has_up = inner || head_up;
  const uint32_t tm = cx.tm;
  // epilogue mapping and early requests (see chain_fwd_layer): thread = (column c = m0 + 32q + lane of layer l, batch
  // rows [cg*NB, cg*NB+NB)); the statistics of the forward pass and, for the row group that applies them, the Adam
  // state of b / gamma / beta are on their way while the operands are staged
  constexpr int NB = ChainCfg<NPAD>::NB;
  const int q = warp & 3, cg = warp >> 2, cl = q * 32 + lane, b0 = cg * NB;
  const int c = m0 + cl;
  const bool mine = c < H;
  const bool bn = (cd.flags & MFAS_FLAG_BN) != 0;
  const bool drop = (cd.flags & MFAS_FLAG_DROPOUT) != 0;
  float mu = 0.f, istd = 1.f, gam = 1.f;
  float pb = 0.f, mb = 0.f, vb = 0.f, mg = 0.f, vg = 0.f, pe = 0.f, me = 0.f, ve = 0.f;
  if (mine) {
    if (bn) { mu = cd.mu[layer * H + c]; istd = cd.invstd[layer * H + c]; gam = cd.p[ly.og + c]; }
    if (cg == 0) {
      pb = cd.p[ly.ob + c]; mb = cd.m[ly.ob + c]; vb = cd.v[ly.ob + c];
      if (bn) { mg = cd.m[ly.og + c]; vg = cd.v[ly.og + c]; pe = cd.p[ly.obe + c]; me = cd.m[ly.obe + c]; ve = cd.v[ly.obe + c]; }
    }
  }

  if (has_up) {
    const int mw = min(128, H - m0);                                   // valid output columns of this tile
    if (inner) {
      const DLayer& up = cd.layer[layer + 1];
      const float* Wu = cd.p + up.oW + up.d_ske + up.d_rgb + m0;       // W_{l+1}[h][hid m0 + .]
      const float* dzu = cd.dzs + (long long)(layer + 1) * bmax * H;
      chain_mma_mnmajor<NPAD>(cx, Wu, up.K, H, H, mw, dzu, H, nrows);
    } else {
      chain_mma_mnmajor<NPAD>(cx, cd.p + cd.oWc + m0, H, cd.C, TC_DLOG_LD, mw, cd.dlog, TC_DLOG_LD, nrows);
    }
  }

  // ---- epilogue: thread = (column c = m0 + 32q + lane of layer l, batch rows [cg*NB, cg*NB+NB)) ----------
  __shared__ float red1[4][128], red2[4][128];
  float av[NB], dh[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) { av[b] = 0.f; dh[b] = 0.f; }
  if (mine) {                                   // activations of this column (loads fly during the MMAs)
    const float* actp = cd.act + ((long long)layer * bmax + b0) * H + c;
#pragma unroll
    for (int b = 0; b < NB; ++b) if (b0 + b < nrows) av[b] = actp[(long long)b * H];
    if (!has_up) {
      const float* dhp = cd.dh + ((long long)layer * bmax + b0) * H + c;
#pragma unroll
      for (int b = 0; b < NB; ++b) if (b0 + b < nrows) dh[b] = dhp[(long long)b * H];
    }
  }
  if (has_up) {
    chain_wait(cx);
    umma::tc_fence_after();
    if (NB == 16) umma::tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + b0, dh);
    else umma::tmem_ld32(tm + ((uint32_t)(q * 32) << 16) + b0, dh);
  }
  if (drop && mine) {
    const uint32_t dkey = dropout_key(drop_seed, (uint32_t)cd.cand_id, step, (uint32_t)layer);
    const float dscale = 1.f / (1.f - drop_p);
#pragma unroll
    for (int b = 0; b < NB; ++b)
      if (b0 + b < nrows) dh[b] = dropout_keep(dkey, (uint32_t)((b0 + b) * H + c), drop_p) ? dh[b] * dscale : 0.f;
  }
  float S1 = 0.f, S2 = 0.f;
  if (bn) {
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int b = 0; b < NB; ++b) if (b0 + b < nrows) { s1 += dh[b]; s2 = fmaf(dh[b], (av[b] - mu) * istd, s2); }
    red1[cg][cl] = s1; red2[cg][cl] = s2;
    __syncthreads();
    S1 = red1[0][cl] + red1[1][cl] + red1[2][cl] + red1[3][cl];
    S2 = red2[0][cl] + red2[1][cl] + red2[2][cl] + red2[3][cl];
    __syncthreads();
  }
  const float m1 = S1 / (float)nrows, m2 = S2 / (float)nrows;
  float db = 0.f;
  if (mine) {
    float* dzl = cd.dzs + ((long long)layer * bmax + b0) * H + c;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (b0 + b < nrows) {
        float da = dh[b];
        if (bn) { const float ah = (av[b] - mu) * istd; da = gam * istd * (dh[b] - m1 - ah * m2); }
        const float dz = da * act_bwd(av[b], ly.act);
        *dzl = dz;
        db += dz;
      }
      dzl += H;
    }
  }
  red1[cg][cl] = db;
  __syncthreads();
  if (mine && cg == 0) {
    db = red1[0][cl] + red1[1][cl] + red1[2][cl] + red1[3][cl];
    auto upd = [&](long long o, float g, float p, float m, float v) {
      if (cd.grad) cd.grad[o] = g;
      adam_update(g, p, m, v, adam, step_size, bc2_sqrt);
      cd.p[o] = p; cd.m[o] = m; cd.v[o] = v;
    };
    upd(ly.ob + c, db, pb, mb, vb);
    if (bn) { upd(ly.og + c, S2, gam, mg, vg); upd(ly.obe + c, S1, pe, me, ve); }
  }
}

template <int NPAD>
__global__ void __launch_bounds__(ChainCfg<NPAD>::THREADS)
k_chain_bwd(const DCand* __restrict__ cands, int layer, int nrows, int bmax, AdamH adam, float step_size,
            float bc2_sqrt, uint32_t drop_seed, float drop_p, uint32_t step, TcErr err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int ok_flag;
  const DCand& cd = cands[blockIdx.y];
  if (layer >= cd.L || (int)blockIdx.x * 128 >= cd.H) return;
  ChainCtx cx;
  chain_ctx_open<NPAD>(cx, smem_raw, &bar, &tmem_slot, &ok_flag);
  chain_bwd_layer<NPAD>(cx, cd, layer, blockIdx.x * 128, nrows, bmax, adam, step_size, bc2_sqrt, drop_seed, drop_p, step);
  if (!cx.ok && threadIdx.x == 0) atomicExch(err.flag, 4);
  chain_ctx_close<NPAD>(cx);
}


// ---------------------------------------------------------------------------------------------
// k_chain_all: the whole serial chain of one step in ONE launch -- for every candidate (one CTA each, H <= 128)
// layer 0..L-1 forward, the classifier head (loss, dlogits, classifier Adam), layer L-1..0 backward.
// The chain is latency-bound (a dozen dependent phases of a few microseconds); fusing it removes eight launch
// boundaries per step and lets the weights the later phases need (hidden columns of W_1.., the classifier)
// be pulled into L2 while the first phases run.  Intermediate activations go through global memory (L2) exactly
// as between the separate kernels.
// dynamic smem: max(chain operand tiles, head tiles)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Row-wise state of the head shared between its phases (static shared memory of k_chain_all)
struct HeadRows {
  float rowloss[MFAS_MAX_BATCH];
  int rowok[MFAS_MAX_BATCH], lab[MFAS_MAX_BATCH], grow[MFAS_MAX_BATCH];
};

// The classifier head on the tensor core (C <= 64, H <= 128, one CTA per candidate):
//   logitsT[c, b] = sum_h W_c[c, h] h_L[b, h]   -- the forward-chain MMA with W_c as a 128-row tile (rows >= C zero)
//   -> + b_c -> shared lg[b][c] -> head_rows (softmax-CE, argmax, dlogits) -> dlogits to cd.dlog (zero-padded)
//   -> loss / accuracy statistics, gradient + Adam of b_c.
// What the CUDA-core head did after that is not on the chain any more: dh_L = dlogits W_c is the "upper layer" MMA of
// the last fusion step's backward (chain_bwd_layer, head_up), and dW_c = dlogits^T h_L with its Adam step is one more
// tile of the weight-streaming kernel (k_tc_bwd_ws, layer index L).  (r01 timeline: the CUDA-core head was 86 k of the
// chain's 280 k cycles -- shared-memory bandwidth on dh/dW_c and the p/m/v round trip of W_c.)
// ML: the multi-label head of the MM-IMDB network (head_rows_ml: weighted BCE-with-logits, per-sample F1 statistic) in place of
// softmax-CE; hr.lab then carries tp | den << 8 per row.
template <bool TRAIN, int NPAD, bool ML>
__device__ __forceinline__ void chain_head_tc(ChainCtx& cx, const DCand& cd, int cand, const DCache& cache, const BatchRef& batch,
                                              int bmax, const AdamH& adam, float step_size, float bc2_sqrt, const HeadOut& out,
                                              HeadRows& hr) {
  constexpr int NB = ChainCfg<NPAD>::NB, LG_LD = TC_DLOG_LD;
  const int H = cd.H, C = cd.C, nrows = batch.n_rows, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, cg = warp >> 2, c = q * 32 + lane, b0 = cg * NB;
  const bool mine = c < C;
  const float bias = mine ? cd.p[cd.obc + c] : 0.f;
  // classifier-bias Adam state for the threads that will apply it (warps 2-3), requested now
  const int cb = tid - 64;
  const bool bias_thread = TRAIN && cb >= 0 && cb < C;
  float pbc = 0.f, mbc = 0.f, vbc = 0.f;
  if (bias_thread) { pbc = cd.p[cd.obc + cb]; mbc = cd.m[cd.obc + cb]; vbc = cd.v[cd.obc + cb]; }

  const float* hl = cd.hid + (long long)(cd.L - 1) * bmax * H;
  chain_mma_kmajor<NPAD>(cx, cd.p + cd.oWc, H, C, hl, H, H, nrows);
  chain_wait(cx);
  umma::tc_fence_after();
  float v[NB];
  if (NB == 16) umma::tmem_ld16(cx.tm + ((uint32_t)(q * 32) << 16) + b0, v);
  else umma::tmem_ld32(cx.tm + ((uint32_t)(q * 32) << 16) + b0, v);
  float* lg = reinterpret_cast<float*>(cx.smem);          // [nrows][LG_LD]; the operand tiles are free (MMAs complete)
  if (mine) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (b0 + b < nrows) {
        const float s = v[b] + bias;
        lg[(b0 + b) * LG_LD + c] = s;
        cd.logits[(b0 + b) * C + c] = s;
        if (out.logits) out.logits[((long long)cand * bmax + b0 + b) * C + c] = s;
      }
    }
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  if (ML) head_rows_ml<TRAIN>(cd, cache, nrows, lg, LG_LD, hr.rowloss, hr.rowok, hr.lab, hr.grow, TRAIN ? cd.dlog : nullptr);
  else head_rows<TRAIN>(cd, cache, nrows, lg, LG_LD, hr.rowloss, hr.rowok, hr.lab, hr.grow, TRAIN ? cd.dlog : nullptr);
  __syncthreads();
  if (warp == 0) {                                         // batch statistics: fixed-order tree (deterministic)
    float ls = 0.f;
    int ok = 0;
    double f1 = 0.0;                                       // ML: sum over rows of 2 tp / (|pred| + |true|), 0 when both are empty
    for (int r = lane; r < nrows; r += 32) {
      ls += hr.rowloss[r]; ok += hr.rowok[r];
      if (ML) { const int tp = hr.lab[r] & 255, den = hr.lab[r] >> 8; f1 += den > 0 ? 2.0 * (double)tp / (double)den : 0.0; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ls += __shfl_xor_sync(0xffffffffu, ls, o); ok += __shfl_xor_sync(0xffffffffu, ok, o);
      if (ML) f1 += __shfl_xor_sync(0xffffffffu, f1, o);
    }
    if (lane == 0) {
      // CrossEntropyLoss(reduction='mean'); ML: torch.mean over all B*C elements (aux_models.py:146)
      const float mean_loss = ML ? ls / ((float)nrows * (float)cd.C) : ls / (float)nrows;
      if (out.loss) out.loss[cand] = mean_loss;
      if (out.correct) out.correct[cand] = ok;
      if (out.stats) {                                              // running_loss += loss.item()*B (ntu.py:72-73, mmimdb.py:88)
        double* st = out.stats + (long long)cand * out.stat_stride + out.stat_off;
        st[0] += (double)mean_loss * (double)nrows;
        st[1] += ML ? f1 : (double)ok;
      }
    }
  }
  if (bias_thread) {                                       // db_c = sum_b dlogits, in row order
    float g = 0.f;
    for (int b = 0; b < nrows; ++b) g += lg[b * LG_LD + cb];
    const long long o = cd.obc + cb;
    if (cd.grad) cd.grad[o] = g;
    adam_update(g, pbc, mbc, vbc, adam, step_size, bc2_sqrt);
    cd.p[o] = pbc; cd.m[o] = mbc; cd.v[o] = vbc;
  }
}

// all threads of all CTAs of the cluster: global writes made before it are visible to the whole cluster after it
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// cl: CTAs per candidate.  1: one CTA walks every 128-column tile of a layer in turn.  2 (inner_repr 256 and at most half as
// many candidates as SMs, e.g. MM-IMDB configs[3]): a 2-CTA thread-block cluster per candidate, one tile each, a cluster
// barrier (release / acquire: the tiles exchange h_l and dz_l through global memory) where the single CTA has __syncthreads;
// the head runs on rank 0.  The chain is latency-bound, so two tiles side by side take the time of one.
template <bool TRAIN, int NPAD, bool TCHEAD, bool ML = false>
__global__ void __launch_bounds__(ChainCfg<NPAD>::THREADS)
k_chain_all(const DCand* __restrict__ cands, DCache cache, BatchRef batch, int bmax, const float* part_base,
            long long part_stride_cand, int hs_ld, int lg_ld, AdamH adam, float step_size, float bc2_sqrt,
            uint32_t drop_seed, float drop_p, uint32_t step, HeadOut ho, TcErr err, int cl) {
  static_assert(ChainCfg<NPAD>::THREADS == kHeadThreads, "the head body is written for the chain CTA size");
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int ok_flag;
  __shared__ DCand scd;            // this candidate's descriptor: every field access below is an LDS, not a dependent global load
  __shared__ HeadRows hr;
  const int cand = (int)blockIdx.x / cl, rank = (int)blockIdx.x % cl, tid = threadIdx.x;
  {
    static_assert(sizeof(DCand) % 4 == 0, "DCand is copied as words");
    const int* src = reinterpret_cast<const int*>(cands + cand);
    int* dst = reinterpret_cast<int*>(&scd);
    for (int i = tid; i < (int)(sizeof(DCand) / 4); i += ChainCfg<NPAD>::THREADS) dst[i] = src[i];
  }
  const int nrows = batch.n_rows;
  if (TCHEAD) {                    // labels of this batch: two dependent loads that nothing on the chain has to wait for
    for (int r = tid; r < nrows; r += ChainCfg<NPAD>::THREADS) {
      const int gr = batch_row(batch, cand, r);
      hr.grow[r] = gr;
      if (!ML) hr.lab[r] = (int)cache.labels[gr];
    }
  }
  ChainCtx cx;
  int stamp_i = 0;
#ifdef MFAS_KSTAMPS
  err.timeline = nullptr;          // the buffer belongs to the streams' role stamps in such a build
#endif
  auto stamp = [&]() { if (err.timeline && tid == 0 && rank == 0 && stamp_i < 16) err.timeline[cand * 16 + stamp_i] = clock64(); ++stamp_i; };
  stamp();
  chain_ctx_open<NPAD>(cx, smem_raw, &bar, &tmem_slot, &ok_flag);      // (__syncthreads inside: scd is visible)
  griddep_launch();
  griddep_wait();                                                      // the partial sums come from the forward stream
  const DCand& cd = scd;
  const int L = cd.L, H = cd.H;
  if (err.timeline && rank == 0) { cx.tl = err.timeline + cand * 16; cx.tl_layer = L > 1 ? 1 : 0; }
  // L2 prefetch, one phase ahead, of what the next phase reads from HBM: the hidden columns of W_{l+1} (H rows x H
  // floats), then W_c.  (Issued all at once at kernel start the 29 MB burst of 128 CTAs delayed the first layer by 3 us.)
  auto prefetch_next = [&](int l) {
    if (l < L) {
      const DLayer& ly = cd.layer[l];
      const float* Wh = cd.p + ly.oW + ly.d_ske + ly.d_rgb;
      const int lsh = 31 - __clz(H >> 5) , lines = H >> 5;       // 128-byte lines per row (H a power of two >= 64: 2, 4 or 8; below: no prefetch)
      for (int i = tid; i < H * lines; i += ChainCfg<NPAD>::THREADS) prefetch_l2(Wh + (long long)(i >> lsh) * ly.K + (i & (lines - 1)) * 32);
    } else {
      for (int i = tid; i < (cd.C * H) >> 5; i += ChainCfg<NPAD>::THREADS) prefetch_l2(cd.p + cd.oWc + i * 32);
    }
  };

  // inner_repr 256: two 128-column tiles per layer -- walked in turn by one CTA (cl = 1) or one per CTA of the cluster (cl = 2);
  // both read h_{l-1} / dz_{l+1} in full and write disjoint columns, so the only ordering is layer by layer
  const int m_first = rank * 128, m_step = 128 * cl;
  for (int l = 0; l < L; ++l) {
    prefetch_next(l + 1);
    for (int m0 = m_first; m0 < H; m0 += m_step) {
      chain_fwd_layer<TRAIN, NPAD>(cx, cd, cand, l, m0, nrows, bmax, part_base, part_stride_cand, drop_seed, drop_p, step);
      umma::tc_fence_before();
      __syncthreads();                                           // h_l (global) and the TMEM reads are done
      umma::tc_fence_after();
    }
    if (cl > 1) cluster_sync_all();                              // ... in every tile of the layer
    stamp();
  }
  if (rank == 0) {
    if (TCHEAD) chain_head_tc<TRAIN, NPAD, ML>(cx, cd, cand, cache, batch, bmax, adam, step_size, bc2_sqrt, ho, hr);
    else head_body<TRAIN, ML>(cd, cand, cache, batch, bmax, hs_ld, lg_ld, adam, step_size, bc2_sqrt, ho, reinterpret_cast<float*>(cx.smem));
  }
  stamp();
  if (TRAIN) {
    for (int l = L - 1; l >= 0; --l) {
      if (cl > 1) { __syncthreads(); cluster_sync_all(); }       // dlogits / dh_L / dz_{l+1} of every tile (global) visible
      for (int m0 = m_first; m0 < H; m0 += m_step) {
        __syncthreads();                                         // dlogits / dh_L / dz_{l+1} (global) visible, smem tiles free
        chain_bwd_layer<NPAD>(cx, cd, l, m0, nrows, bmax, adam, step_size, bc2_sqrt, drop_seed, drop_p, step, TCHEAD);
        umma::tc_fence_before();
        __syncthreads();
        umma::tc_fence_after();
      }
      stamp();
    }
  }
  if (!cx.ok && tid == 0) atomicExch(err.flag, 7);
  chain_ctx_close<NPAD>(cx);
}

// ---------------------------------------------------------------------------------------------
// k_chain_small: the serial chain of one step for inner_repr 16 / 32, on the CUDA cores, ENTIRELY ON CHIP.
// A 64 x 16 hidden state does not need a tensor core: k_chain_all spends 11-14 k cycles per phase on it (operand
// staging -> fence -> MMA -> commit -> tcgen05.ld, and a global-memory round trip of h_l / a_l / dz_l between phases;
// profiles/r02z_chain_timeline_search.txt), 16 warps with 16 useful lanes in 4 of them, one CTA per SM.  Here ONE round
// of cp.async brings everything the chain reads -- the forward stream's partial sums, the hidden columns of W_1.., the
// classifier, every per-column vector with its Adam state -- into shared memory, where h_l, a_l and dz_l stay from the
// first phase to the last; a phase is H (or C) FMAs per output element, a fixed-order column reduction and one or two
// __syncthreads; the Adam steps of all b / gamma / beta / b_c run side by side after the last phase instead of on each
// phase's critical path; two CTAs fit an SM, so 256 candidates are one wave.  Same inputs and the same outputs for the
// backward stream (h_l, dz_l, dlogits in global memory) as k_chain_all.
//   thread = (column c = tid % HN, row-group slot tid / HN): G groups of 4 consecutive batch rows (the float4 of the
//   partial-sum layout); HN == cd.H.
// dynamic smem (floats): h[L][NPAD][HN] | a[L][NPAD][HN] (TRAIN) | dz[2][NPAD][HN] (TRAIN) | W_hid[L][HN][HN+1] |
//                        W_c[64][HN+1] | lg[NPAD][64] | vec[L][NV][HN] | red[5][16][32] | partial sums [items][NPAD/4][HN] float4
// ---------------------------------------------------------------------------------------------
template <int NPAD, int HN> struct ChainSmall {
  static constexpr int THREADS = kHeadThreads;
  static constexpr int SLOTS = THREADS / HN;               // row-group slots
  static constexpr int NG = NPAD / 4;                      // groups of 4 batch rows
  static constexpr int G = NG > SLOTS ? NG / SLOTS : 1;    // groups per thread
  static constexpr int WLD = HN + 1;                       // row stride of the weight tiles: conflict-free by row AND by column
  static constexpr int LG_LD = TC_DLOG_LD;
  static constexpr int H_L = NPAD * HN;
  // b g be | m_b v_b m_g v_g m_be v_be | running mean, var | batch mean, invstd | d b, d gamma, d beta
  static constexpr int NV = 16;
  static constexpr size_t smem(int L, bool train, int items) {
    // the partial sums share their space with what is written after they are consumed: a_l (layer l's items end at or after
    // (l + 1) H_L) and dz (backward only); evaluation: with h_l
    const size_t state = train ? (size_t)L * H_L + (size_t)(items > L + 2 ? items : L + 2) * H_L : (size_t)(items > L ? items : L) * H_L;
    return sizeof(float) * (state + (size_t)L * HN * WLD + 64 * WLD + NPAD * LG_LD + (size_t)L * NV * HN + 5 * 16 * 32);
  }
};

__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async16_ca(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}

// sum over the batch rows of one column: lanes of a warp hold different columns (HN = 32) or two row-group slots of 16
// columns (HN = 16); fixed order: slot pair, then warp 0..15.  Every thread of the CTA calls it (idle slots pass 0).
template <int HN>
__device__ __forceinline__ float col_sum(float v, float* red, int c) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (HN == 16) v += __shfl_xor_sync(0xffffffffu, v, 16);
  if (lane < HN) red[warp * 32 + lane] = v;
  __syncthreads();
  float t[4] = {0.f, 0.f, 0.f, 0.f};           // fixed order, four chains of four (not one chain of sixteen dependent adds)
#pragma unroll
  for (int w = 0; w < 16; ++w) t[w & 3] += red[w * 32 + c];
  return (t[0] + t[1]) + (t[2] + t[3]);
}
template <int HN>
__device__ __forceinline__ void col_sum2(float& v1, float& v2, float* red1, float* red2, int c) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (HN == 16) { v1 += __shfl_xor_sync(0xffffffffu, v1, 16); v2 += __shfl_xor_sync(0xffffffffu, v2, 16); }
  if (lane < HN) { red1[warp * 32 + lane] = v1; red2[warp * 32 + lane] = v2; }
  __syncthreads();
  float t1[4] = {0.f, 0.f, 0.f, 0.f}, t2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int w = 0; w < 16; ++w) { t1[w & 3] += red1[w * 32 + c]; t2[w & 3] += red2[w * 32 + c]; }
  v1 = (t1[0] + t1[1]) + (t1[2] + t1[3]); v2 = (t2[0] + t2[1]) + (t2[2] + t2[3]);
}

// head_rows (kernels_ffma.cuh) for four rows of a warp at once -- rows warp + 16 i of the round: the same arithmetic per
// row, but the four dependent shuffle trees are interleaved instead of run one after the other (the chain's longest phase
// at inner_repr 16).  Single-task softmax-CE only; multitask and the multi-label head take the row-at-a-time functions.
template <bool TRAIN, int NPAD>
__device__ __forceinline__ void head_rows_x4(const DCand& cd, int nrows, float* lg, int lg_ld, float* rowloss, int* rowok,
                                             const int* lab, float* dlog) {
  const int C = cd.C, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_n = 1.f / (float)nrows;
  for (int rb = 0; rb < NPAD; rb += 64) {
    if (rb >= nrows) break;
    float v0[4], v1[4], mx[4];
    int am[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = rb + warp + 16 * i;
      const bool ok = r < nrows;
      const float* row = lg + r * lg_ld;
      v0[i] = (ok && lane < C) ? row[lane] : -INFINITY; v1[i] = (ok && lane + 32 < C) ? row[lane + 32] : -INFINITY;
      mx[i] = fmaxf(v0[i], v1[i]);
      am[i] = (v1[i] > v0[i]) ? lane + 32 : lane;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float om = __shfl_xor_sync(0xffffffffu, mx[i], o);
        const int oa = __shfl_xor_sync(0xffffffffu, am[i], o);
        if (om > mx[i] || (om == mx[i] && oa < am[i])) { mx[i] = om; am[i] = oa; }
      }
    }
    float se[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) se[i] = (lane < C ? expf(v0[i] - mx[i]) : 0.f) + (lane + 32 < C ? expf(v1[i] - mx[i]) : 0.f);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) se[i] += __shfl_xor_sync(0xffffffffu, se[i], o);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = rb + warp + 16 * i;
      if (r < nrows) {                                                  // (warp-uniform)
        float* row = lg + r * lg_ld;
        const float lse = logf(se[i]);
        const int y = lab[r];
        const float vy = __shfl_sync(0xffffffffu, y < 32 ? v0[i] : v1[i], y & 31);
        if (lane == 0) { rowloss[r] = -((vy - mx[i]) - lse); rowok[r] = (am[i] == y) ? 1 : 0; }
        if (TRAIN) {
          const float d0 = lane < C ? (expf((v0[i] - mx[i]) - lse) - (lane == y ? 1.f : 0.f)) * inv_n : 0.f;
          const float d1 = lane + 32 < C ? (expf((v1[i] - mx[i]) - lse) - (lane + 32 == y ? 1.f : 0.f)) * inv_n : 0.f;
          if (lane < C) row[lane] = d0;
          if (lane + 32 < C) row[lane + 32] = d1;
          if (dlog) { dlog[r * 64 + lane] = d0; dlog[r * 64 + 32 + lane] = d1; }
        }
      }
    }
  }
}

template <bool TRAIN, int NPAD, int HN, bool ML>
__global__ void __launch_bounds__((ChainSmall<NPAD, HN>::THREADS), 2)
k_chain_small(const DCand* __restrict__ cands, DCache cache, BatchRef batch, int bmax, const float* __restrict__ part_base,
              long long part_stride_cand, AdamH adam, float step_size, float bc2_sqrt, uint32_t drop_seed, float drop_p,
              uint32_t step, HeadOut ho, TcErr err, int raw_items) {
  using Cfg = ChainSmall<NPAD, HN>;
  constexpr int THREADS = Cfg::THREADS, G = Cfg::G, WLD = Cfg::WLD, LG_LD = Cfg::LG_LD, H_L = Cfg::H_L, NV = Cfg::NV, NG = Cfg::NG;
  extern __shared__ __align__(16) float csm[];
  __shared__ DCand scd;
  __shared__ HeadRows hr;
  __shared__ long long nbt_s[MFAS_MAX_LAYERS];
  __shared__ float bc_s[4][64];                      // classifier bias: p, m, v, gradient
  const int cand = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  {
    const int* src = reinterpret_cast<const int*>(cands + cand);
    int* dst = reinterpret_cast<int*>(&scd);
    for (int i = tid; i < (int)(sizeof(DCand) / 4); i += THREADS) dst[i] = src[i];
  }
  const int nrows = batch.n_rows;
  for (int r = tid; r < nrows; r += THREADS) {
    const int gr = batch_row(batch, cand, r);
    hr.grow[r] = gr;
    if (!ML) hr.lab[r] = (int)cache.labels[gr];
  }
  int stamp_i = 0;
#ifdef MFAS_KSTAMPS
  const bool tl_on = false;                          // the buffer belongs to the streams' role stamps in such a build
#else
  const bool tl_on = err.timeline != nullptr;
#endif
  auto stamp = [&]() { if (tl_on && tid == 0 && stamp_i < 10) err.timeline[cand * 16 + stamp_i] = clock64(); ++stamp_i; };
  auto stamp_at = [&](int i) { if (tl_on && tid == 0) err.timeline[cand * 16 + i] = clock64(); };      // slots 10..: inside the phases
  stamp();
  __syncthreads();
  const DCand& cd = scd;
  const int L = cd.L, C = cd.C;
  const bool bn = (cd.flags & MFAS_FLAG_BN) != 0;
  const bool drop = TRAIN && (cd.flags & MFAS_FLAG_DROPOUT);
  const bool gated = (cd.flags & MFAS_FLAG_ALPHAS) != 0;
  const float dscale = drop ? 1.f / (1.f - drop_p) : 1.f;
  const float inv_n = 1.f / (float)nrows;            // the batch means below are sum * (1 / n): one division per launch on the chain
  int n_items = 0;
  for (int l = 0; l < L; ++l) n_items += tc_fwd_items_g(cd.layer[l].d_ske, cd.layer[l].d_rgb, gated, cd.kb_item);
  float* h_s = csm;
  float* a_s = h_s + (TRAIN ? (size_t)L * H_L : 0);
  float* raw = a_s;                                  // the forward stream's partial sums, [item][NPAD/4][HN] float4 (see ChainSmall::smem)
  float* dz_s = a_s + (size_t)L * H_L;
  // raw_items: the partial sums the launch has room for (the group's largest candidate, or 0 when that does not fit: the
  // layers then read them from global memory -- the same sums in the same order)
  const bool staged = n_items <= raw_items;
  float* wh_s = TRAIN ? a_s + (size_t)max(L + 2, raw_items) * H_L : h_s + (size_t)max(L, raw_items) * H_L;
  float* wc_s = wh_s + (size_t)L * HN * WLD;
  float* lg = wc_s + 64 * WLD;
  float* vec = lg + NPAD * LG_LD;
  float* red = vec + (size_t)L * NV * HN;
  const int c = tid % HN, slot = tid / HN;
  const bool active = slot < NG;                     // inner_repr 16 with 64 rows: the upper 8 warps only help with staging and the head
  if (TRAIN) for (int i = tid; i < NPAD * LG_LD; i += THREADS) lg[i] = 0.f;         // columns >= C stay zero (dh_L reads whole float4s)
  for (int i = C * HN + tid; i < 64 * HN; i += THREADS) wc_s[(i / HN) * WLD + (i % HN)] = 0.f;
  griddep_launch();
  griddep_wait();                                    // the partial sums come from the forward stream; the weights from the last step
  stamp_at(10);

  // ---- stage everything the chain reads from global memory: one round of cp.async ----------------------------------
  {
    const float4* pc = reinterpret_cast<const float4*>(part_base + (long long)cand * part_stride_cand);
    for (int i = tid; i < (staged ? n_items * NG * HN : 0); i += THREADS) {
      const int cc = i % HN, grp = (i / HN) % NG, item = i / (HN * NG);
      cp_async16_ca(umma::smem_u32(raw + 4 * (size_t)i), pc + ((long long)item * NG + grp) * 128 + cc);
    }
    for (int i = tid; i < L * HN; i += THREADS) {      // per-column vectors and their Adam state
      const int l = i / HN, cc = i % HN;
      const DLayer& ly = cd.layer[l];
      const uint32_t vv = umma::smem_u32(vec + (size_t)l * NV * HN + cc);
      cp_async4(vv, cd.p + ly.ob + cc);
      if (TRAIN) { cp_async4(vv + 4 * 3 * HN, cd.m + ly.ob + cc); cp_async4(vv + 4 * 4 * HN, cd.v + ly.ob + cc); }
      if (bn) {
        cp_async4(vv + 4 * 1 * HN, cd.p + ly.og + cc); cp_async4(vv + 4 * 2 * HN, cd.p + ly.obe + cc);
        cp_async4(vv + 4 * 9 * HN, cd.bufs + ly.orm + cc); cp_async4(vv + 4 * 10 * HN, cd.bufs + ly.orv + cc);
        if (TRAIN) {
          cp_async4(vv + 4 * 5 * HN, cd.m + ly.og + cc); cp_async4(vv + 4 * 6 * HN, cd.v + ly.og + cc);
          cp_async4(vv + 4 * 7 * HN, cd.m + ly.obe + cc); cp_async4(vv + 4 * 8 * HN, cd.v + ly.obe + cc);
        }
      }
    }
    for (int i = tid; i < (L - 1) * HN * HN; i += THREADS) {      // hidden columns of W_1..
      const int l = 1 + i / (HN * HN), r = (i / HN) % HN, j = i % HN;
      const DLayer& ly = cd.layer[l];
      cp_async4(umma::smem_u32(wh_s + (size_t)l * HN * WLD + r * WLD + j), cd.p + ly.oW + (long long)r * ly.K + ly.d_ske + ly.d_rgb + j);
    }
    for (int i = tid; i < C * HN; i += THREADS) cp_async4(umma::smem_u32(wc_s + (i / HN) * WLD + (i % HN)), cd.p + cd.oWc + i);
    if (tid < C) {
      cp_async4(umma::smem_u32(&bc_s[0][tid]), cd.p + cd.obc + tid);
      if (TRAIN) { cp_async4(umma::smem_u32(&bc_s[1][tid]), cd.m + cd.obc + tid); cp_async4(umma::smem_u32(&bc_s[2][tid]), cd.v + cd.obc + tid); }
    }
    if (TRAIN && bn && tid < L) nbt_s[tid] = cd.nbt[tid];
    cp_async_wait_all();
  }
  __syncthreads();
  stamp_at(11);

  // ---- forward ------------------------------------------------------------------------------------------------------
  int item0 = 0;
  for (int l = 0; l < L; ++l) {
    const DLayer& ly = cd.layer[l];
    float* vv = vec + (size_t)l * NV * HN + c;
    const int nsplit = tc_fwd_items_g(ly.d_ske, ly.d_rgb, gated, cd.kb_item);
    float a[G][4];
#pragma unroll
    for (int k = 0; k < G; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) a[k][i] = 0.f;
    if (active) {
      // feature partial sums of this layer, in split order (alpha gates: one factor per partial, see chain_fwd_layer)
      const float sg = gated ? gate_of(cd.p[ly.oalpha]) : 1.f;
      const int n_ske = tc_fwd_items_of(ly.d_ske, cd.kb_item);
      const float4* ps = staged ? reinterpret_cast<const float4*>(raw) + c
                                : reinterpret_cast<const float4*>(part_base + (long long)cand * part_stride_cand) + c;
      const int pstr = staged ? HN : 128;            // float4s per row group: packed in shared memory, Hp = 128 in the forward stream's layout
      for (int s = 0; s < nsplit; ++s) {
        const float gt = !gated ? 1.f : (s < n_ske ? sg : 1.0f - sg);
#pragma unroll
        for (int k = 0; k < G; ++k) {
          const float4 v = ps[((long long)(item0 + s) * NG + slot + k * Cfg::SLOTS) * pstr];
          if (gated) { a[k][0] = fmaf(gt, v.x, a[k][0]); a[k][1] = fmaf(gt, v.y, a[k][1]); a[k][2] = fmaf(gt, v.z, a[k][2]); a[k][3] = fmaf(gt, v.w, a[k][3]); }
          else { a[k][0] += v.x; a[k][1] += v.y; a[k][2] += v.z; a[k][3] += v.w; }
        }
      }
      if (l > 0) {                                   // + W_hid h_{l-1}
        const float* w = wh_s + (size_t)l * HN * WLD + c * WLD;
        const float* hp = h_s + (size_t)(l - 1) * H_L;
        float acc[G][4];
#pragma unroll
        for (int k = 0; k < G; ++k)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[k][i] = 0.f;
#pragma unroll 4
        for (int j = 0; j < HN; j += 4) {
          const float w0 = w[j], w1 = w[j + 1], w2 = w[j + 2], w3 = w[j + 3];
#pragma unroll
          for (int k = 0; k < G; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 hv = *reinterpret_cast<const float4*>(hp + (4 * (slot + k * Cfg::SLOTS) + i) * HN + j);
              acc[k][i] = fmaf(hv.x, w0, acc[k][i]); acc[k][i] = fmaf(hv.y, w1, acc[k][i]);
              acc[k][i] = fmaf(hv.z, w2, acc[k][i]); acc[k][i] = fmaf(hv.w, w3, acc[k][i]);
            }
        }
#pragma unroll
        for (int k = 0; k < G; ++k)
#pragma unroll
          for (int i = 0; i < 4; ++i) a[k][i] += acc[k][i];
      }
      const float bias = vv[0];
#pragma unroll
      for (int k = 0; k < G; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) a[k][i] = (4 * (slot + k * Cfg::SLOTS) + i < nrows) ? act_fwd(a[k][i] + bias, ly.act) : 0.f;
    }
    item0 += nsplit;
    if (l == 1) stamp_at(14);
    float mean = 0.f, istd = 1.f;
    if (bn) {
      if (TRAIN) {
        float s1 = 0.f;
#pragma unroll
        for (int k = 0; k < G; ++k)
#pragma unroll
          for (int i = 0; i < 4; ++i) s1 += a[k][i];
        mean = col_sum<HN>(s1, red, c) * inv_n;
        if (l == 1) stamp_at(6);
        float qq = 0.f;
#pragma unroll
        for (int k = 0; k < G; ++k)
#pragma unroll
          for (int i = 0; i < 4; ++i) if (4 * (slot + k * Cfg::SLOTS) + i < nrows) { const float d = a[k][i] - mean; qq = fmaf(d, d, qq); }
        if (l == 1) stamp_at(7);
        const float var = col_sum<HN>(qq, red + 512, c) * inv_n;
        if (l == 1) stamp_at(8);
        istd = rsqrtf(var + kBnEps);
        if (tid >= THREADS - HN) {                   // one thread per column, in the last warp (idle at inner_repr 16, and never the slowest)
          vv[11 * HN] = mean; vv[12 * HN] = istd;
          const float n = (float)nrows;
          cd.bufs[ly.orm + c] = (1.f - kBnMomentum) * vv[9 * HN] + kBnMomentum * mean;
          cd.bufs[ly.orv + c] = (1.f - kBnMomentum) * vv[10 * HN] + kBnMomentum * (var * (n / (n - 1.f)));
          if (c == 0) cd.nbt[l] = nbt_s[l] + 1;
        }
      } else {
        mean = vv[9 * HN];
        istd = rsqrtf(vv[10 * HN] + kBnEps);
      }
    }
    if (l == 1) stamp_at(15);
    if (!(TRAIN && bn)) __syncthreads();             // a_l / h_l overwrite partial sums of this layer that another thread may still be reading
    if (active) {
      const float gamma = bn ? vv[1 * HN] : 1.f, beta = bn ? vv[2 * HN] : 0.f;
      const uint32_t dkey = drop ? dropout_key(drop_seed, (uint32_t)cd.cand_id, step, (uint32_t)l) : 0u;
      float* hs = h_s + (size_t)l * H_L + c;
      float* as = a_s + (size_t)l * H_L + c;
      float* hg = cd.hid + (long long)l * bmax * HN + c;
#pragma unroll
      for (int k = 0; k < G; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int b = 4 * (slot + k * Cfg::SLOTS) + i;
          float h = bn ? (a[k][i] - mean) * istd * gamma + beta : a[k][i];
          if (drop) h = dropout_keep(dkey, (uint32_t)(b * HN + c), drop_p) ? h * dscale : 0.f;
          if (b >= nrows) h = 0.f;
          hs[b * HN] = h;
          if (TRAIN) as[b * HN] = a[k][i];
          if (b < nrows) hg[(long long)b * HN] = h;                    // the backward stream's x operand (hidden columns)
        }
    }
    __syncthreads();
    stamp();
  }

  // ---- head: logits = h_L W_c^T + b_c; thread = (class k = tid % 64, rows tid / 64 + 8 i) ---------------------------
  {
    const int k = tid & 63, r0 = tid >> 6;
    constexpr int NR = NPAD / 8;
    const float* hp = h_s + (size_t)(L - 1) * H_L;
    const float* w = wc_s + k * WLD;
    float acc[NR];
#pragma unroll
    for (int i = 0; i < NR; ++i) acc[i] = 0.f;
#pragma unroll 1
    for (int j = 0; j < HN; j += 4) {
      const float w0 = w[j], w1 = w[j + 1], w2 = w[j + 2], w3 = w[j + 3];
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const float4 hv = *reinterpret_cast<const float4*>(hp + (r0 + 8 * i) * HN + j);
        acc[i] = fmaf(hv.x, w0, acc[i]); acc[i] = fmaf(hv.y, w1, acc[i]); acc[i] = fmaf(hv.z, w2, acc[i]); acc[i] = fmaf(hv.w, w3, acc[i]);
      }
    }
    if (k < C) {
      const float bias = bc_s[0][k];
#pragma unroll
      for (int i = 0; i < NR; ++i) {
        const int b = r0 + 8 * i;
        if (b < nrows) {
          const float s = acc[i] + bias;
          lg[b * LG_LD + k] = s;
          cd.logits[b * C + k] = s;
          if (ho.logits) ho.logits[((long long)cand * bmax + b) * C + k] = s;
        }
      }
    }
  }
  __syncthreads();
  stamp_at(12);
  if (ML) head_rows_ml<TRAIN>(cd, cache, nrows, lg, LG_LD, hr.rowloss, hr.rowok, hr.lab, hr.grow, TRAIN ? cd.dlog : nullptr);
  else if ((cd.flags & MFAS_FLAG_MULTITASK) && cache.logit_rgb && cache.logit_ske)
    head_rows<TRAIN>(cd, cache, nrows, lg, LG_LD, hr.rowloss, hr.rowok, hr.lab, hr.grow, TRAIN ? cd.dlog : nullptr);
  else head_rows_x4<TRAIN, NPAD>(cd, nrows, lg, LG_LD, hr.rowloss, hr.rowok, hr.lab, TRAIN ? cd.dlog : nullptr);
  __syncthreads();
  stamp_at(13);
  if (warp == 15) {                                        // batch statistics: fixed-order tree (see chain_head_tc)
    float ls = 0.f;
    int ok = 0;
    double f1 = 0.0;
    for (int r = lane; r < nrows; r += 32) {
      ls += hr.rowloss[r]; ok += hr.rowok[r];
      if (ML) { const int tp = hr.lab[r] & 255, den = hr.lab[r] >> 8; f1 += den > 0 ? 2.0 * (double)tp / (double)den : 0.0; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ls += __shfl_xor_sync(0xffffffffu, ls, o); ok += __shfl_xor_sync(0xffffffffu, ok, o);
      if (ML) f1 += __shfl_xor_sync(0xffffffffu, f1, o);
    }
    if (lane == 0) {
      const float mean_loss = ML ? ls / ((float)nrows * (float)cd.C) : ls / (float)nrows;
      if (ho.loss) ho.loss[cand] = mean_loss;
      if (ho.correct) ho.correct[cand] = ok;
      if (ho.stats) {
        double* st = ho.stats + (long long)cand * ho.stat_stride + ho.stat_off;
        st[0] += (double)mean_loss * (double)nrows;
        st[1] += ML ? f1 : (double)ok;
      }
    }
  }
  stamp();
  if (!TRAIN) return;

  // ---- backward (pre-update weights: W_hid and W_c are stepped by the backward stream after this kernel) -------------
  {                                                        // db_c = sum_b dlogits: 8 partial sums per class, combined after the last phase
    const int k = tid & 63, r0 = tid >> 6;
    float g = 0.f;
#pragma unroll
    for (int i = 0; i < NPAD / 8; ++i) g += lg[(r0 + 8 * i) * LG_LD + k];      // (rows >= nrows and columns >= C are zero)
    red[4 * 512 + r0 * 64 + k] = g;
  }
  for (int l = L - 1; l >= 0; --l) {
    const DLayer& ly = cd.layer[l];
    float* vv = vec + (size_t)l * NV * HN + c;
    float* dz_cur = dz_s + ((l + 1) & 1) * H_L;            // dz_{l+1}
    float* dz_nxt = dz_s + (l & 1) * H_L;
    float dh[G][4], av[G][4];
#pragma unroll
    for (int k = 0; k < G; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) { dh[k][i] = 0.f; av[k][i] = 0.f; }
    if (active) {
      if (l == L - 1) {                                    // dh_L = dlogits W_c
        const float* w = wc_s + c;
        const int C4 = (C + 3) & ~3;
#pragma unroll 4
        for (int k4 = 0; k4 < C4; k4 += 4) {
          const float w0 = w[k4 * WLD], w1 = w[(k4 + 1) * WLD], w2 = w[(k4 + 2) * WLD], w3 = w[(k4 + 3) * WLD];
#pragma unroll
          for (int k = 0; k < G; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 d = *reinterpret_cast<const float4*>(lg + (4 * (slot + k * Cfg::SLOTS) + i) * LG_LD + k4);
              dh[k][i] = fmaf(d.x, w0, dh[k][i]); dh[k][i] = fmaf(d.y, w1, dh[k][i]); dh[k][i] = fmaf(d.z, w2, dh[k][i]); dh[k][i] = fmaf(d.w, w3, dh[k][i]);
            }
        }
      } else {                                             // dh_l = dz_{l+1} W_hid,l+1
        const float* w = wh_s + (size_t)(l + 1) * HN * WLD + c;
#pragma unroll 4
        for (int h4 = 0; h4 < HN; h4 += 4) {
          const float w0 = w[h4 * WLD], w1 = w[(h4 + 1) * WLD], w2 = w[(h4 + 2) * WLD], w3 = w[(h4 + 3) * WLD];
#pragma unroll
          for (int k = 0; k < G; ++k)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 d = *reinterpret_cast<const float4*>(dz_cur + (4 * (slot + k * Cfg::SLOTS) + i) * HN + h4);
              dh[k][i] = fmaf(d.x, w0, dh[k][i]); dh[k][i] = fmaf(d.y, w1, dh[k][i]); dh[k][i] = fmaf(d.z, w2, dh[k][i]); dh[k][i] = fmaf(d.w, w3, dh[k][i]);
            }
        }
      }
      const float* as = a_s + (size_t)l * H_L + c;
      const uint32_t dkey = drop ? dropout_key(drop_seed, (uint32_t)cd.cand_id, step, (uint32_t)l) : 0u;
#pragma unroll
      for (int k = 0; k < G; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int b = 4 * (slot + k * Cfg::SLOTS) + i;
          av[k][i] = as[b * HN];
          if (b >= nrows) dh[k][i] = 0.f;
          else if (drop) dh[k][i] = dropout_keep(dkey, (uint32_t)(b * HN + c), drop_p) ? dh[k][i] * dscale : 0.f;
        }
    }
    const float mu = bn ? vv[11 * HN] : 0.f, istd = bn ? vv[12 * HN] : 1.f, gam = bn ? vv[1 * HN] : 1.f;
    float S1 = 0.f, S2 = 0.f;
    if (bn) {
#pragma unroll
      for (int k = 0; k < G; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) if (4 * (slot + k * Cfg::SLOTS) + i < nrows) { S1 += dh[k][i]; S2 = fmaf(dh[k][i], (av[k][i] - mu) * istd, S2); }
      col_sum2<HN>(S1, S2, red, red + 512, c);
    }
    const float m1 = S1 * inv_n, m2 = S2 * inv_n;
    float db = 0.f;
    if (active) {
      float* dzg = cd.dzs + (long long)l * bmax * HN + c;
#pragma unroll
      for (int k = 0; k < G; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int b = 4 * (slot + k * Cfg::SLOTS) + i;
          float dz = 0.f;
          if (b < nrows) {
            float da = dh[k][i];
            if (bn) { const float ah = (av[k][i] - mu) * istd; da = gam * istd * (dh[k][i] - m1 - ah * m2); }
            dz = da * act_bwd(av[k][i], ly.act);
            dzg[(long long)b * HN] = dz;
            db += dz;
          }
          dz_nxt[b * HN + c] = dz;
        }
    }
    db = col_sum<HN>(db, red + 1024 + (l & 1) * 512, c);   // (its barrier also publishes dz_l for the layer below; two buffers in turn:
                                                           //  without BatchNorm it is the only barrier of a layer)
    if (tid >= THREADS - HN) { vv[13 * HN] = db; vv[14 * HN] = S2; vv[15 * HN] = S1; }
    stamp();
  }
  __syncthreads();
  // ---- Adam of every per-column vector, side by side: b_l, gamma_l, beta_l (gradients: sum dz, sum dh a_hat, sum dh) and b_c
  for (int t = tid; t < L * HN * 3; t += THREADS) {
    const int l = t / (3 * HN), which = (t / HN) % 3, cc = t % HN;
    if (which > 0 && !bn) continue;
    const DLayer& ly = cd.layer[l];
    const float* vv = vec + (size_t)l * NV * HN + cc;
    const long long o = (which == 0 ? ly.ob : which == 1 ? ly.og : ly.obe) + cc;
    const float g = vv[(13 + which) * HN];
    float p = vv[which * HN], m = vv[(3 + 2 * which) * HN], v = vv[(4 + 2 * which) * HN];
    if (cd.grad) cd.grad[o] = g;
    adam_update(g, p, m, v, adam, step_size, bc2_sqrt);
    cd.p[o] = p; cd.m[o] = m; cd.v[o] = v;
  }
  if (tid >= THREADS - 64 && tid - (THREADS - 64) < C) {
    const int k = tid - (THREADS - 64);
    float g = 0.f;
#pragma unroll
    for (int r0 = 0; r0 < 8; ++r0) g += red[4 * 512 + r0 * 64 + k];
    float p = bc_s[0][k], m = bc_s[1][k], v = bc_s[2][k];
    const long long o = cd.obc + k;
    if (cd.grad) cd.grad[o] = g;
    adam_update(g, p, m, v, adam, step_size, bc2_sqrt);
    cd.p[o] = p; cd.m[o] = m; cd.v[o] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// backward, all layers: one work item = 128 weight columns (k) x 64 output rows (h) of one layer.
//   dW^T[k,h] = sum_b x[b,k] dz[b,h]      A = x chunk (MN-major, M = k), B = dz slice (MN-major, N = h)
// TMEM lane = weight column k, so for a fixed h the 32 lanes of a warp touch 32 consecutive floats of
// W[h,:], m[h,:], v[h,:]: every Adam access is a coalesced 128-byte line with no transpose.
// 512 threads: 4 warps per TMEM lane quarter, each owning 16 of the 64 rows; every thread keeps
// 2 x 12 independent loads in flight (two register sets, software-pipelined; the first two sets are
// requested before the MMAs are waited for).  2 CTAs/SM -> 32 warps/SM hide the HBM latency.
// (Measured alternatives, kept out: draining TMEM to a shared-memory tile and running Adam on float4s
//  halves the instruction count but is 11 % slower; looping several chunks per CTA is 25 % slower.)
// grid = (max items per candidate, H/64, candidates);  BP = batch rows padded (64 or 128) = MMA K extent
// dynamic smem (1024-aligned): A_hi | A_lo (4 blocks x BP x 128 B) | B_hi | B_lo (2 blocks x BP x 128 B)
// ---------------------------------------------------------------------------------------------
constexpr int TC_BWD_THREADS = 512;

template <int BP, bool KEEP_GRAD>
__global__ void __launch_bounds__(TC_BWD_THREADS, BP == 64 ? 2 : 1)
k_tc_bwd_all(const DCand* __restrict__ cands, DCache cache, BatchRef batch, int bmax, AdamH adam, float step_size,
             float bc2_sqrt, TcErr err, int dbg) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);
  const int cand = blockIdx.z;
  const DCand& cd = cands[cand];
  const int H = cd.H;
  const int h0 = blockIdx.y * TC_BWD_HT;
  if (h0 >= H) return;
  int layer = 0, chunk = blockIdx.x;
  for (; layer < cd.L; ++layer) {
    const int n = tc_bwd_items(cd.layer[layer].K);
    if (chunk < n) break;
    chunk -= n;
  }
  if (layer >= cd.L) return;
  const DLayer& ly = cd.layer[layer];
  const int K = ly.K;
  const int kc0 = chunk * TC_BWD_KT;
  const int kw = min(TC_BWD_KT, K - kc0);
  const int nrows = batch.n_rows, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  constexpr uint32_t BLK = BP * 128, A_TILE = 4 * BLK, B_TILE = 2 * BLK;
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + A_TILE;
  uint8_t* b_hi = a_lo + A_TILE;
  uint8_t* b_lo = b_hi + B_TILE;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int ok_flag;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 64);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }

  // ---- A tile: x[:, kc0:kc0+128] (one smem row per batch row, 4 blocks of 32 columns) and
  //      B tile: dz_l[:, h0:h0+64] (2 blocks); all loads of a thread are issued before the first store
  const int fs = ly.d_ske, fr = ly.d_rgb;
  const float* src; long long ld; int kl; bool gather = true;
  if (kc0 < fs) { src = cache.ske[ly.ske_tap]; ld = cache.ske_ld[ly.ske_tap]; kl = kc0; }
  else if (kc0 < fs + fr) { src = cache.rgb[ly.rgb_tap]; ld = cache.rgb_ld[ly.rgb_tap]; kl = kc0 - fs; }
  else { src = cd.hid + (long long)(layer - 1) * bmax * H; ld = H; kl = kc0 - fs - fr; gather = false; }
  {
    constexpr int XIT = BP * 32 / TC_BWD_THREADS, DIT = BP * 16 / TC_BWD_THREADS;
    const float* dzl = cd.dzs + (long long)layer * bmax * H + h0;
    float4 xv[XIT], dv[DIT];
#pragma unroll
    for (int i = 0; i < XIT; ++i) {
      const int idx = tid + TC_BWD_THREADS * i, r = idx >> 5, c4 = idx & 31;
      xv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows && c4 * 4 < kw) {
        const long long row = gather ? (long long)batch_row(batch, cand, r) : (long long)r;
        xv[i] = __ldg(reinterpret_cast<const float4*>(src + row * ld + kl + c4 * 4));
      }
    }
#pragma unroll
    for (int i = 0; i < DIT; ++i) {
      const int idx = tid + TC_BWD_THREADS * i, r = idx >> 4, c4 = idx & 15;
      dv[i] = (r < nrows) ? *reinterpret_cast<const float4*>(dzl + (long long)r * H + c4 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < XIT; ++i) {
      const int idx = tid + TC_BWD_THREADS * i, r = idx >> 5, c4 = idx & 31;
      store_split(a_hi, a_lo, (uint32_t)(c4 >> 3) * BLK + umma::sw128_b32(r, (c4 & 7) * 16), xv[i]);
    }
#pragma unroll
    for (int i = 0; i < DIT; ++i) {
      const int idx = tid + TC_BWD_THREADS * i, r = idx >> 4, c4 = idx & 15;
      store_split(b_hi, b_lo, (uint32_t)(c4 >> 3) * BLK + umma::sw128_b32(r, (c4 & 7) * 16), dv[i]);
    }
  }
  umma::fence_async_smem();
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tm = tmem_slot;

  if (tid == 0) {
    constexpr uint32_t idesc = umma::idesc_tf32(128, TC_BWD_HT, true, true);
    const int ksteps = (nrows + 7) >> 3;
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint32_t adv = ks * 1024u;               // 8 batch rows = two 512-byte swizzle atoms
      const uint64_t dah = umma::smem_desc(umma::smem_u32(a_hi) + adv, BLK, 512, umma::kLayoutSw128Base32);
      const uint64_t dal = umma::smem_desc(umma::smem_u32(a_lo) + adv, BLK, 512, umma::kLayoutSw128Base32);
      const uint64_t dbh = umma::smem_desc(umma::smem_u32(b_hi) + adv, BLK, 512, umma::kLayoutSw128Base32);
      const uint64_t dbl = umma::smem_desc(umma::smem_u32(b_lo) + adv, BLK, 512, umma::kLayoutSw128Base32);
      umma::mma_tf32(tm, dal, dbh, idesc, ks > 0 ? 1u : 0u);
      umma::mma_tf32(tm, dah, dbl, idesc, 1u);
      umma::mma_tf32(tm, dah, dbh, idesc, 1u);
    }
    umma::mma_commit(&bar);
  }

  // ---- epilogue: thread = weight column kc0 + 32q + lane (q = warp%4), rows h0 + 16*(warp/4) + [0,16) --
  const int q = warp & 3, cg = warp >> 2;
  const int kcol = q * 32 + lane;
  const bool kvalid = q * 32 < kw && !(dbg & 1);      // kw is a multiple of 32: uniform per warp
  const int hb = h0 + cg * 16;
  float* Wg = cd.p + ly.oW + (long long)hb * K + kc0 + kcol;
  const long long moff = (long long)(cd.m - cd.p), voff = (long long)(cd.v - cd.p), goff = KEEP_GRAD ? (long long)(cd.grad - cd.p) : 0;
  const float inv_bc2 = 1.f / bc2_sqrt;
  struct Set { float p[4], m[4], v[4]; };
  Set sa, sb;
  auto prefetch = [&](Set& st, int j0) {
    const float* w = Wg + (long long)j0 * K;
#pragma unroll
    for (int j = 0; j < 4; ++j) { st.p[j] = w[0]; st.m[j] = w[moff]; st.v[j] = w[voff]; w += K; }
  };
  auto update = [&](Set& st, int j0) {
    float g[4];
    umma::tmem_ld4(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(cg * 16 + j0), g);
    if (kvalid) {
      float* w = Wg + (long long)j0 * K;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (KEEP_GRAD) w[goff] = g[j];
        adam_update_fast(g[j], st.p[j], st.m[j], st.v[j], adam, step_size, inv_bc2);
        w[0] = st.p[j]; w[moff] = st.m[j]; w[voff] = st.v[j];
        w += K;
      }
    }
  };
  if (kvalid) { prefetch(sa, 0); prefetch(sb, 4); }   // p/m/v loads fly while the tensor core works
  const bool ok = umma::cta_wait(&bar, 0, &ok_flag);
  if (!ok && tid == 0) atomicExch(err.flag, 2);
  umma::tc_fence_after();
  update(sa, 0);
  if (kvalid) prefetch(sa, 8);
  update(sb, 4);
  if (kvalid) prefetch(sb, 12);
  update(sa, 8);
  update(sb, 12);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tm, 64);
}

// ---------------------------------------------------------------------------------------------
// backward, all layers, persistent + warp-specialised (operand stage of BP = 64 batch rows; batches up to 128 rows in two
// passes over the stage into the same accumulator): one CTA per SM walks a static list of
// tiles (the same 128 weight columns x 64 rows tile as k_tc_bwd_all) with three decoupled roles, so
// the x/dz operand traffic, the tensor-core work and the Adam p/m/v stream of different tiles overlap
// all the time instead of taking turns inside a CTA:
//   warps 0-7   stagers : gather x rows + dz slice -> hi/lo split -> smem operand stage   (full  -> MMA)
//   warp  8     MMA     : 3 x tcgen05.mma per 8 batch rows into TMEM buffer t             (empty -> stagers,
//                                                                                          tfull[t] -> Adam)
//   warps 9-16  Adam    : p/m/v arrive through a per-warp cp.async ring in shared memory (WS_RING batches of
//                         8 rows x 3 arrays x 128 B, requested one whole tile ahead, so the bytes in flight
//                         are bounded by shared memory instead of registers); gradient straight out of
//                         TMEM, update, coalesced 128-byte stores                         (tempty[t] -> MMA)
// (r01 ncu of the register-prefetch version: 3.8 TB/s, long-scoreboard bound with 25 KB of p/m/v in flight
//  per SM; the ring keeps 72-96 KB in flight.)
// smem: 1 operand stage of 96 KB + 8 rings x WS_RING x 3 KB; TMEM: 2 x 64 columns.
// Tile list: int4 {candidate, layer, first column, first row}.
// ---------------------------------------------------------------------------------------------
// Self-contained tile record (built on the host whenever arenas are (re)bound): the Adam warps need no
// dependent descriptor loads, one 64-byte record fetched two tiles ahead is all they read.
// end of the concat source [ske | rgb | hidden] that weight column kc0 of a layer belongs to
__host__ __device__ __forceinline__ int tc_bwd_seg_end(int d_ske, int d_rgb, int K, int kc0) {
  return kc0 < d_ske ? d_ske : (kc0 < d_ske + d_rgb ? d_ske + d_rgb : K);
}
struct __align__(16) BwdTile {
  float* W;                       // &params[oW + h0 * K + kc0]
  long long moff, voff, goff;     // adam_m - params, adam_v - params, grad - params (floats; goff 0 when no grad arena)
  int K, kw, rows, pad1;          // rows: valid rows of the 64 (64 for a fusion layer, C for the classifier tile)
  int cand, layer, kc0, h0;       // 16-byte aligned: the stagers read these four as one int4 (layer == L: the classifier)
  const float* alpha;             // alpha gates: &params[oalpha] of the tile's layer (pre-update value: the step's last launch updates it)
  int gate, slot;                 // gate: 0 none, 1 ske columns (x sigmoid(alpha)), 2 rgb columns (x (1 - sigmoid(alpha))); slot: tile index
  // operand sources, resolved on the host (k_tc_bwd_small's loaders: no walk tile -> candidate -> layer per fill)
  const float* xsrc;              // xkind 0: &hid[layer - 1][0][kc0 - d_ske - d_rgb] (classifier: &hid[L - 1][0][kc0]); else unused
  const float* dz;                // &dzs[layer][0][h0] (classifier: &dlog[0][h0])
  int xkind, xtap, xcol, xld;     // xkind 1 / 2: columns xcol.. of ske / rgb tap xtap, rows gathered by the batch; xld: row stride of xsrc
  int dzld, hw, pad2, pad3;       // row stride of dz; valid dz columns of the tile
};
static_assert(sizeof(BwdTile) == 128, "one 128-byte line per tile record");
constexpr int TC_DSP_PER_TILE = 8;                      // one d(loss)/d(sigmoid(alpha)) partial per Adam warp and tile
constexpr int TC_WS_THREADS = 17 * 32;
constexpr int WS_RING = 4;                              // ring slots per Adam warp (= batches per tile)
constexpr int WS_SLOT = 3 * 8 * 128;                    // bytes: 3 arrays x 8 rows x 32 floats
constexpr size_t TC_WS_SMEM = 1024 + 98304 + 8 * (size_t)WS_RING * WS_SLOT;

__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst_smem), "l"(src), "l"(policy) : "memory");
}

// ALPHA (groups with the modality gates on): the staged x is unscaled, so the accumulator holds G = dz^T x; the weight gradient
// of a gated tile is gate * G, and d(loss)/d(sigmoid(alpha)) = +- sum(W o G) over the tile (pre-update W) -- summed per Adam warp
// in a fixed order and left in dsp[tile][warp] for k_alpha_step_tc, which finishes the sum and updates alpha after this launch.
template <bool KEEP_GRAD, bool ALPHA = false>
__global__ void __launch_bounds__(TC_WS_THREADS, 1)
k_tc_bwd_ws(const DCand* __restrict__ cands, DCache cache, BatchRef batch, int bmax, AdamH adam, float step_size,
            float bc2_sqrt, const BwdTile* __restrict__ tiles, int n_tiles, TcErr err, float* __restrict__ dsp = nullptr) {
  constexpr int BP = 64;
  constexpr uint32_t BLK = BP * 128, A_TILE = 4 * BLK, B_TILE = 2 * BLK, STAGE = 2 * A_TILE + 2 * B_TILE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);
  __shared__ uint64_t full, empty, tfull[2], tempty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nrows = batch.n_rows;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (tid == 32) {
    umma::mbar_init(&full, 8); umma::mbar_init(&empty, 1);
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&tfull[i], 1); umma::mbar_init(&tempty[i], 8); }
    umma::fence_mbar_init();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  griddep_launch();
  griddep_wait();                                                  // dz / activations / dlogits come from the chain kernel
  const uint32_t tm = tmem_slot;
  const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  bool ok = true;

  if (warp < 8) {
    // ================================ stagers =====================================================
    // Two-deep software pipeline so that no wait sits on a chain of dependent loads: the tile descriptor
    // and the 8 gather indices of tile i+2 are fetched while the x / dz loads of tile i+1 are in flight
    // (r01 ncu: with index -> x load pairs issued one after the other the stagers, not HBM, set the pace).
    // (r02fj: host-resolved tile records as in k_tc_bwd_small -- the descriptor walk gone, the x loads issued ~6000 cycles
    //  earlier -- made this kernel 12 % SLOWER: the gathers then queue behind the Adam warps' ring requests for ~10 k cycles.)
    // Batches above 64 rows: the batch is the reduction dimension of dW = x^T dz, so a tile is staged and multiplied in
    // `npass` passes of <= 64 rows into the SAME accumulator (the 96 KB operand stage stays as it is); the unit the stagers and
    // the MMA warp hand over is a "fill" f = tile * npass + pass.
    struct Desc { const float* src; const float* dz; long long ld; int H, kw, row0; int row[8]; };   // kw = valid x columns | valid dz columns (H < 64) << 16; row0 = first batch row of this pass
    float4 xv[8], dv[4];
    const int npass = (nrows + BP - 1) / BP, n_fill = n_my * npass;
    auto fetch_desc = [&](int f, Desc& d) {
      const int i = npass == 1 ? f : (f >> 1), row0 = npass == 1 ? 0 : (f & 1) * BP;
      const int4 t = *reinterpret_cast<const int4*>(&tiles[blockIdx.x + i * gridDim.x].cand);   // {cand, layer, kc0, h0}
      const DCand& cd = cands[t.x];
      const int H = cd.H, kc0 = t.z;
      bool gather = true;
      d.row0 = row0;
      if (t.y >= cd.L) {                                // the classifier as one more layer: x = h_L, dz = dlogits (zero-padded to 64 classes)
        d.src = cd.hid + (long long)(cd.L - 1) * bmax * H + kc0; d.ld = H; gather = false;
        d.H = TC_DLOG_LD; d.kw = min(TC_BWD_KT, H - kc0) | (TC_BWD_HT << 16);
        d.dz = cd.dlog;
      } else {
        const DLayer& ly = cd.layer[t.y];
        const int fs = ly.d_ske, fr = ly.d_rgb;
        // a tile never straddles two concat sources: the host cuts the tile list at the source boundaries (tc_bwd_seg_end),
        // so taps narrower than / not a multiple of the 128-column tile (the 64-wide MM-IMDB text tap) are short tiles
        int seg_end = ly.K;
        if (kc0 < fs) { d.src = cache.ske[ly.ske_tap] + kc0; d.ld = cache.ske_ld[ly.ske_tap]; seg_end = fs; }
        else if (kc0 < fs + fr) { d.src = cache.rgb[ly.rgb_tap] + (kc0 - fs); d.ld = cache.rgb_ld[ly.rgb_tap]; seg_end = fs + fr; }
        else { d.src = cd.hid + (long long)(t.y - 1) * bmax * H + (kc0 - fs - fr); d.ld = H; gather = false; }
        d.H = H; d.kw = min(TC_BWD_KT, seg_end - kc0) | (min(TC_BWD_HT, H - t.w) << 16);
        d.dz = cd.dzs + (long long)t.y * bmax * H + t.w;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {                     // rows row0 + warp, row0 + warp + 8, ...: one index per warp and j
        const int r = min(row0 + warp + 8 * j, nrows - 1);
        d.row[j] = gather ? batch_row(batch, t.x, r) : r;
      }
    };
    auto issue_loads = [&](const Desc& d) {             // unconditional loads from clamped addresses, then select
      const int kw = d.kw & 0xffff, hw = d.kw >> 16;
      const int c4 = tid & 31, cc = min(c4 * 4, kw - 4);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(d.src + (long long)d.row[j] * d.ld + cc));
        xv[j] = (d.row0 + warp + 8 * j < nrows && c4 * 4 < kw) ? x : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = tid + 256 * j, r = d.row0 + (idx >> 4), cd4 = idx & 15;
        const float4 x = *reinterpret_cast<const float4*>(d.dz + (long long)min(r, nrows - 1) * d.H + min(cd4 * 4, hw - 4));
        dv[j] = (r < nrows && cd4 * 4 < hw) ? x : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    };
    Desc d1{};
    if (n_fill > 0) { fetch_desc(0, d1); issue_loads(d1); }
    if (n_fill > 1) fetch_desc(1, d1);
    for (int i = 0; i < n_fill; ++i) {
      if (!umma::mbar_wait(&empty, (i & 1) ^ 1)) { ok = false; break; }   // MMAs of fill i-1 have read the stage
      if (tid == 0) MFAS_KSTAMP(16, i, 7);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int idx = tid + 256 * j, r = idx >> 5, c4 = idx & 31;
        store_split(smem, smem + A_TILE, (uint32_t)(c4 >> 3) * BLK + umma::sw128_b32(r, (c4 & 7) * 16), xv[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = tid + 256 * j, r = idx >> 4, c4 = idx & 15;
        store_split(smem + 2 * A_TILE, smem + 2 * A_TILE + B_TILE, (uint32_t)(c4 >> 3) * BLK + umma::sw128_b32(r, (c4 & 7) * 16), dv[j]);
      }
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&full);
      if (tid == 0) MFAS_KSTAMP(16, i, 0);
      if (i + 1 < n_fill) issue_loads(d1);             // in flight while this fill's MMAs and the Adam pass run
      if (i + 2 < n_fill) fetch_desc(i + 2, d1);
    }
  } else if (warp == 8) {
    // ================================ MMA issuer ==================================================
    // Two products per 8 batch rows instead of three (r02): dz_hi and dz_lo are adjacent tiles with the same block stride, so
    // [dz_hi | dz_lo] is ONE N = 128 operand: x_hi [dz_hi | dz_lo] -> accumulator columns [0, 64) and [64, 128) (x_hi read from
    // shared memory once), x_lo dz_hi -> [0, 64).  The Adam warps add the two column ranges.  These MN-major tf32 MMAs cost
    // ~170-250 cycles each, and with ONE operand stage their issue alternates with the stagers' stores: under the power cap
    // (1.6 GHz) that serial chain, not HBM, was the tile time.
    constexpr uint32_t idesc = umma::idesc_tf32(128, TC_BWD_HT, true, true), idesc_cat = umma::idesc_tf32(128, 2 * TC_BWD_HT, true, true);
    const int npass = (nrows + BP - 1) / BP, n_fill = n_my * npass;
    for (int f = 0; f < n_fill; ++f) {
      const int i = npass == 1 ? f : (f >> 1), pass = npass == 1 ? 0 : (f & 1), tb = i & 1;
      if (!umma::mbar_wait(&full, f & 1)) { ok = false; break; }
      if (lane == 0) MFAS_KSTAMP(16, f, 3);
      if (pass == 0 && !umma::mbar_wait(&tempty[tb], ((i >> 1) & 1) ^ 1)) { ok = false; break; }
      if (lane == 0) MFAS_KSTAMP(16, f, 1);
      umma::tc_fence_after();
      if (umma::elect_one()) {
        const uint32_t a_hi = umma::smem_u32(smem), a_lo = a_hi + A_TILE, b_hi = a_lo + A_TILE;      // b_lo = b_hi + B_TILE: the next two blocks
        static_assert(B_TILE == 2 * BLK, "[dz_hi | dz_lo] must be four blocks of one stride");
        const uint32_t d = tm + tb * 128;
        const int ksteps = (min(BP, nrows - pass * BP) + 7) >> 3;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t adv = ks * 1024u;
          const uint64_t dah = umma::smem_desc(a_hi + adv, BLK, 512, umma::kLayoutSw128Base32);
          const uint64_t dal = umma::smem_desc(a_lo + adv, BLK, 512, umma::kLayoutSw128Base32);
          const uint64_t dbh = umma::smem_desc(b_hi + adv, BLK, 512, umma::kLayoutSw128Base32);
          umma::mma_tf32(d, dah, dbh, idesc_cat, (pass > 0 || ks > 0) ? 1u : 0u);
          umma::mma_tf32(d, dal, dbh, idesc, 1u);
        }
        umma::mma_commit(&empty);                      // operand stage free once these MMAs have read it
        if (pass == npass - 1) umma::mma_commit(&tfull[tb]);   // accumulator ready for the Adam warps
      }
      __syncwarp();
      if (lane == 0) MFAS_KSTAMP(16, f, 4);
    }
  } else {
    // ================================ Adam warps ==================================================
    const int aw = warp - 9;                           // 0..7
    const int q = warp & 3, cg = aw >> 2;              // TMEM lane quarter this warp may read; row half
    const float inv_bc2 = 1.f / bc2_sqrt;
    uint8_t* ring = smem + STAGE + (size_t)aw * WS_RING * WS_SLOT;
    const uint32_t ring_u32 = umma::smem_u32(ring);
    const int srow = lane >> 3, schunk = lane & 7;     // cp.async: a lane moves 16 B of row (4*u + srow)
    const uint64_t stream_policy = l2_stream_policy((err.l2_hints & 2) != 0);
    struct Tile { float* W; long long K, moff, voff, goff; int rc; float gsc, gsign; int slot; };   // rc = rows | cols << 8: valid rows / columns of this warp's 32 x 32 (0 columns: nothing to do)
    struct Raw { int4 a, b, c, d; };                   // first 48 bytes of a BwdTile (+ the gate record when ALPHA)
    auto fetch_raw = [&](int i) {
      const int4* r = reinterpret_cast<const int4*>(&tiles[blockIdx.x + i * gridDim.x]);
      Raw o; o.a = __ldg(r); o.b = __ldg(r + 1); o.c = __ldg(r + 2);
      o.d = ALPHA ? __ldg(r + 4) : make_int4(0, 0, 0, 0);
      return o;
    };
    auto open_tile = [&](const Raw& r) {
      Tile o;
      const long long Wbits = ((long long)(uint32_t)r.a.y << 32) | (uint32_t)r.a.x;
      o.moff = ((long long)(uint32_t)r.a.w << 32) | (uint32_t)r.a.z;
      o.voff = ((long long)(uint32_t)r.b.y << 32) | (uint32_t)r.b.x;
      o.goff = KEEP_GRAD ? (((long long)(uint32_t)r.b.w << 32) | (uint32_t)r.b.z) : 0;
      o.K = r.c.x;
      // 64-row x 128-column tiles of a fusion layer: 32 | 32 << 8; fewer rows in the classifier tile and for inner_repr < 64,
      // fewer columns in the last chunk of a layer (the 16 / 32 / 64 hidden columns)
      o.rc = min(32, max(0, r.c.z - cg * 32)) | (min(32, max(0, r.c.y - q * 32)) << 8);
      o.W = reinterpret_cast<float*>(Wbits) + (long long)(cg * 32) * o.K + q * 32;   // first row / column of this warp
      o.gsc = 1.f; o.gsign = 0.f; o.slot = 0;
      if (ALPHA && r.d.z) {
        const float* ap = reinterpret_cast<const float*>(((long long)(uint32_t)r.d.y << 32) | (uint32_t)r.d.x);
        const float sg = gate_of(__ldg(ap));
        o.gsc = r.d.z == 1 ? sg : 1.0f - sg;
        o.gsign = r.d.z == 1 ? 1.f : -1.f;
        o.slot = r.d.w;
      }
      return o;
    };
    // request batch j (rows 8j .. 8j+7 of this warp's 32) of tile t into ring slot j; always commits a group
    auto request = [&](const Tile& t, int j) {
      if (t.rc >> 8) {
        const int rows = t.rc & 255, cols = t.rc >> 8;
        const uint32_t dst = ring_u32 + j * WS_SLOT + srow * 128 + schunk * 16;
        const float* w = t.W + (long long)(8 * j + srow) * t.K + schunk * 4;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (8 * j + 4 * u + srow < rows && schunk * 4 < cols) {
            const float* wu = w + (long long)(4 * u) * t.K;
            cp_async16(dst + u * 512, wu, stream_policy);
            cp_async16(dst + 1024 + u * 512, wu + t.moff, stream_policy);
            cp_async16(dst + 2048 + u * 512, wu + t.voff, stream_policy);
          }
        }
      }
      cp_async_commit();
    };
    Tile cur{}, nxt{};
    Raw ahead{};                                       // record of tile i+2, in flight during iteration i
    if (n_my > 0) {
      cur = open_tile(fetch_raw(0));
#pragma unroll
      for (int j = 0; j < WS_RING; ++j) request(cur, j);
    }
    if (n_my > 1) ahead = fetch_raw(1);
    for (int i = 0; i < n_my; ++i) {
      const int tb = i & 1;
      const bool more = i + 1 < n_my;
      if (more) nxt = open_tile(ahead);
      if (i + 2 < n_my) ahead = fetch_raw(i + 2);
      if (!umma::mbar_wait(&tfull[tb], (i >> 1) & 1)) { ok = false; break; }
      if (aw == 0 && lane == 0) MFAS_KSTAMP(16, i, 5);
      umma::tc_fence_after();
      float dsum = 0.f;                                // ALPHA: sum over this warp's part of the tile of W (pre-update) o G
#pragma unroll
      for (int j = 0; j < WS_RING; ++j) {
        cp_async_wait<WS_RING - 1>();                  // the oldest outstanding batch (this one) has landed
        __syncwarp();
        float g[8], g2[8];
        umma::tmem_ld8(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(tb * 128 + cg * 32 + 8 * j), g);
        umma::tmem_ld8(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(tb * 128 + 64 + cg * 32 + 8 * j), g2);
#pragma unroll
        for (int r = 0; r < 8; ++r) g[r] += g2[r];
        if (cur.rc >> 8) {
          const float* sp = reinterpret_cast<const float*>(ring + j * WS_SLOT) + lane;
          float* w = cur.W + (long long)(8 * j) * cur.K + lane;
          float p[8], m[8], v[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) { p[r] = sp[r * 32]; m[r] = sp[256 + r * 32]; v[r] = sp[512 + r * 32]; }
          if (ALPHA) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              if (8 * j + r < (cur.rc & 255) && lane < (cur.rc >> 8)) dsum = fmaf(p[r], g[r], dsum);   // (ring slots of masked elements hold stale bytes)
              g[r] *= cur.gsc;
            }
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) adam_update_fast(g[r], p[r], m[r], v[r], adam, step_size, inv_bc2);
          if (cur.rc == (32 | (32 << 8))) {             // full 32 x 32 (always, except in the classifier tile and for inner_repr < 64)
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              if (KEEP_GRAD) w[cur.goff] = g[r];
              w[0] = p[r]; w[cur.moff] = m[r]; w[cur.voff] = v[r];
              w += cur.K;
            }
          } else {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              if (8 * j + r < (cur.rc & 255) && lane < (cur.rc >> 8)) {
                if (KEEP_GRAD) w[cur.goff] = g[r];
                w[0] = p[r]; w[cur.moff] = m[r]; w[cur.voff] = v[r];
              }
              w += cur.K;
            }
          }
        }
        __syncwarp();                                  // every lane has read slot j before it is refilled
        if (more) request(nxt, j); else cp_async_commit();
      }
      if (ALPHA && cur.gsign != 0.f) {                 // (warp-uniform) fixed-order tree; k_alpha_step_tc adds the 8 x tiles partials in order
        dsum = warp_sum(dsum);
        if (lane == 0) dsp[(long long)cur.slot * TC_DSP_PER_TILE + aw] = cur.gsign * dsum;
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&tempty[tb]);   // this warp has drained its part of the accumulator
      if (aw == 0 && lane == 0) MFAS_KSTAMP(16, i, 6);
      cur = nxt;
    }
    cp_async_wait<0>();
  }
  if (!ok) atomicExch(err.flag, 5);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tm, 256);
}

// ---------------------------------------------------------------------------------------------
// backward, all layers, inner_repr 16 / 32 (the search default): the persistent weight-gradient + Adam stream with the
// operand staging of the forward stream.  A tile is 128 weight columns x HN rows: its p / m / v bytes are 1/4 or 1/8 of a
// k_tc_bwd_ws tile, so what is left is the x tile (64 gathered rows x 512 bytes) -- and k_tc_bwd_ws keeps exactly ONE such
// tile in flight per SM (registers of its stager warps): search256 on one B200 took 5 us per tile, 203 us per step, for bytes
// that take 75 us.  Here the gathered x rows and the dz slice go global -> shared with cp.async straight into the MN-major
// operand layout (the raw tile IS the hi operand), three raw stages deep; converters derive the lo tiles (two lo stages); per 8 batch
// rows the MMA warp issues x_hi [dz_hi | dz_lo] (N = 32 + HN, one operand across the raw and the lo stage) and x_lo dz_hi:
//   warps 0-3   loaders    : cp.async into raw stage f % RAW, sources from the host-resolved tile record   (rawfree <- MMA; landed -> converters)
//   warps 4-11  converters : two groups of four warps, every other fill: lo = rna_tf32(x - trunc_tf32(x)) into the group's lo stage
//                                                                                         (lofree[g] <- MMA; lofull[g] -> MMA)
//   warp  12    MMA        : 2 x tcgen05.mma per 8 batch rows, M = 128 columns             (tfull[t] -> Adam)
//   warps 13-16 Adam       : one TMEM lane quarter each (32 columns x HN rows); p / m / v through a per-warp cp.async ring
//                            requested one tile ahead; gradient straight out of TMEM (the two column ranges added)   (tempty[t] -> MMA)
// Tile list: as k_tc_bwd_ws with row tiles of HN (the classifier's C rows are HN-row tiles h0 = 0, HN, ...).
// Batches above 64 rows: two passes over the 64-row stage into the same accumulator.
// (Measured and kept out, r02fv: converter warps that TRANSPOSE while they split, so that both operands are K-major tiles -- correct,
//  MMA issue 2760 -> ~1600 cycles per tile, but the transposed hi tile can no longer be the raw stage, the operand stage is then
//  72 KB and fits only once next to three raw stages: split and MMA issue alternate again, 157 us against 140.)
// ---------------------------------------------------------------------------------------------
template <int HN> struct BwdSmall {
  static constexpr int BP = 64, RAW = HN <= 16 ? 3 : 2, LO = 2, NB = HN / 8;   // NB: 8-row p / m / v batches per tile and warp = ring slots
  static constexpr uint32_t A_BLK = BP * 128, A_BYTES = 4 * A_BLK, B_BYTES = BP * 128, TILE = A_BYTES + B_BYTES;
  static constexpr size_t SMEM = 1024 + (size_t)(RAW + LO) * TILE + 4 * (size_t)NB * WS_SLOT;
  static constexpr int THREADS = 17 * 32;
};

template <int HN, bool KEEP_GRAD, bool ALPHA>
__global__ void __launch_bounds__((BwdSmall<HN>::THREADS), 1)
k_tc_bwd_small(const DCand* __restrict__ cands, DCache cache, BatchRef batch, int bmax, AdamH adam, float step_size,
               float bc2_sqrt, const BwdTile* __restrict__ tiles, int n_tiles, TcErr err, float* __restrict__ dsp) {
  using Cfg = BwdSmall<HN>;
  constexpr int R = Cfg::RAW, BP = Cfg::BP, NB = Cfg::NB;
  constexpr uint32_t A_BLK = Cfg::A_BLK, A_BYTES = Cfg::A_BYTES, TILE = Cfg::TILE;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = umma::align1024(smem_raw);
  uint8_t* lo_base = smem + R * TILE;
  __shared__ uint64_t landed[R], rawfree[R], lofull[2], lofree[2], tfull[2], tempty[2];
  __shared__ uint32_t tmem_slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nrows = batch.n_rows;
  constexpr uint32_t TM_COLS = 128;                                // two accumulator buffers x 64 columns (see the MMA issuer)
  if (warp == 0) umma::tmem_alloc(&tmem_slot, TM_COLS);
  if (tid == 32) {
    for (int i = 0; i < R; ++i) { umma::mbar_init(&landed[i], 128); umma::mbar_init(&rawfree[i], 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&lofull[i], 4); umma::mbar_init(&lofree[i], 1); }
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&tfull[i], 1); umma::mbar_init(&tempty[i], 4); }
    umma::fence_mbar_init();
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  griddep_launch();
  griddep_wait();                                                  // dz / activations / dlogits come from the chain kernel
  const uint32_t tm = tmem_slot;
  const int n_my = (n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int npass = (nrows + BP - 1) / BP, n_fill = n_my * npass;
  bool ok = true;

  if (warp < 4) {
    // ================================ loaders =====================================================
    // warp w moves batch rows w, w + 4, ... of the pass: a row of the x tile is 512 contiguous bytes = one 16-byte chunk per lane
    struct Desc { const float* src; const float* dz; long long ld; int H, kw, hw, row0; int row[16]; };
    struct Rec { int4 shape, id, ptrs, x, z; };          // the tile record's loader fields (one 128-byte line, fetched two fills ahead)
    const uint32_t s0 = umma::smem_u32(smem);
    const uint64_t x_policy = l2_stream_policy(false);
    auto fetch_rec = [&](int f) {
      const int i = npass == 1 ? f : (f >> 1);
      const int4* r = reinterpret_cast<const int4*>(tiles + blockIdx.x + (long long)i * gridDim.x);
      Rec o; o.shape = __ldg(r + 2); o.id = __ldg(r + 3); o.ptrs = __ldg(r + 5); o.x = __ldg(r + 6); o.z = __ldg(r + 7);
      return o;
    };
    auto open_desc = [&](int f, const Rec& rc, Desc& d) {
      d.row0 = npass == 1 ? 0 : (f & 1) * BP;
      const bool gather = rc.x.x != 0;
      if (rc.x.x == 1) { d.src = cache.ske[rc.x.y] + rc.x.z; d.ld = cache.ske_ld[rc.x.y]; }
      else if (rc.x.x == 2) { d.src = cache.rgb[rc.x.y] + rc.x.z; d.ld = cache.rgb_ld[rc.x.y]; }
      else { d.src = reinterpret_cast<const float*>(((long long)(uint32_t)rc.ptrs.y << 32) | (uint32_t)rc.ptrs.x); d.ld = rc.x.w; }
      d.dz = reinterpret_cast<const float*>(((long long)(uint32_t)rc.ptrs.w << 32) | (uint32_t)rc.ptrs.z);
      d.H = rc.z.x; d.hw = rc.z.y; d.kw = rc.shape.y;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = min(d.row0 + warp + 4 * j, nrows - 1);
        d.row[j] = gather ? batch_row(batch, rc.id.x, r) : r;
      }
    };
    Desc d{};
    Rec rnext{};
    if (n_fill > 0) open_desc(0, fetch_rec(0), d);
    if (n_fill > 1) rnext = fetch_rec(1);
    for (int f = 0; f < n_fill; ++f) {
      const int sg = f % R;
      if (f >= R && !umma::mbar_wait(&rawfree[sg], ((f / R) & 1) ^ 1)) { ok = false; break; }
      if (tid == 0) MFAS_KSTAMP(16, f, 7);
      const uint32_t a = s0 + sg * TILE, b = a + A_BYTES;
      const int c4 = lane;                              // x: 16-byte chunk c4 of the row's 512 bytes
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = warp + 4 * j;                     // row of the stage
        // sw128_b32(r, 16 (c4 & 7)) = r 128 + (((c4 & 7) >> 1) ^ (r & 3)) 32 + (c4 & 1) 16
        const uint32_t dst = a + (uint32_t)(c4 >> 3) * A_BLK + (uint32_t)r * 128u + (uint32_t)(((((c4 & 7) >> 1) ^ (r & 3))) << 5) + (uint32_t)((c4 & 1) << 4);
        cp_async16_zfill(dst, d.src + (long long)d.row[j] * d.ld + min(c4 * 4, d.kw - 4), d.row0 + r < nrows && c4 * 4 < d.kw, x_policy);
      }
      {   // dz slice: 64 rows x HN floats: HN / 4 chunks per row, 64 * HN / 4 chunks over 128 threads
        constexpr int CPR = HN / 4, PER = 64 * CPR / 128;
#pragma unroll
        for (int j = 0; j < PER; ++j) {
          const int idx = tid + 128 * j, r = idx / CPR, cd4 = idx % CPR;
          const uint32_t dst = b + (uint32_t)r * 128u + (uint32_t)((((cd4 >> 1) ^ (r & 3))) << 5) + (uint32_t)((cd4 & 1) << 4);
          cp_async16_zf(dst, d.dz + (long long)min(d.row0 + r, nrows - 1) * d.H + min(cd4 * 4, d.hw - 4), d.row0 + r < nrows && cd4 * 4 < d.hw);
        }
      }
      cp_async_arrive_noinc(&landed[sg]);
      if (tid == 0) MFAS_KSTAMP(16, f, 0);
      if (f + 1 < n_fill) open_desc(f + 1, rnext, d);   // gather indices of the next fill (one load deep) while this one flies
      if (f + 2 < n_fill) rnext = fetch_rec(f + 2);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  } else if (warp < 12) {
    // ================================ converters ==================================================
    // two groups of four warps, each with a lo stage of its own and every other fill: the split of fill f + 1 runs while the
    // MMAs of fill f are issued (one lo stage and all eight warps on the same fill made split and MMA issue strictly alternate:
    // 3900 cycles per tile)
    const int ct = tid - 128, grp = ct >> 7, gt = ct & 127;
    constexpr int NCH = (int)(TILE / 16);
#pragma unroll 1
    for (int f = grp; f < n_fill; f += 2) {
      const int sg = f % R, k = f >> 1;
      if (!umma::mbar_wait(&landed[sg], (f / R) & 1)) { ok = false; break; }
      if (gt == 0) MFAS_KSTAMP(16, f, 1);
      if (k >= 1 && !umma::mbar_wait(&lofree[grp], (k & 1) ^ 1)) { ok = false; break; }
      const float4* src = reinterpret_cast<const float4*>(smem + sg * TILE);
      float4* dst = reinterpret_cast<float4*>(lo_base + grp * TILE);
#pragma unroll 5
      for (int j = 0; j < NCH / 128; ++j) {
        const float4 x = src[gt + j * 128];
        float4 l;
        l.x = umma::round_tf32(x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u));
        l.y = umma::round_tf32(x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u));
        l.z = umma::round_tf32(x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u));
        l.w = umma::round_tf32(x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u));
        dst[gt + j * 128] = l;
      }
      umma::fence_async_smem();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&lofull[grp]);
      if (gt == 0) MFAS_KSTAMP(16, f, 2);
    }
  } else if (warp == 12) {
    // ================================ MMA issuer ==================================================
    // Two products per 8 batch rows instead of three: B' = [dz_hi | dz_lo] as ONE N = 32 + HN operand -- the MN-major
    // descriptor's leading-dimension offset reaches from the raw stage's dz tile (columns 0..31) to the lo stage's (columns
    // 32..) -- so x_hi is read from shared memory once for x_hi dz_hi (accumulator columns [0, HN)) and x_hi dz_lo (columns
    // [32, 32 + HN)); x_lo dz_hi adds into [0, HN).  The Adam warps add the two column ranges.  (Issue of these small MMAs, ~40
    // cycles each, is what the tile costs on this warp; columns [HN, 32) of the raw dz tile are never written nor read back.)
    constexpr uint32_t idesc = umma::idesc_tf32(128, HN, true, true), idesc_cat = umma::idesc_tf32(128, 32 + HN, true, true);
    for (int f = 0; f < n_fill; ++f) {
      const int i = npass == 1 ? f : (f >> 1), pass = npass == 1 ? 0 : (f & 1), tb = i & 1, sg = f % R, sl = f & 1;
      if (!umma::mbar_wait(&lofull[sl], (f >> 1) & 1)) { ok = false; break; }
      if (lane == 0) MFAS_KSTAMP(16, f, 3);
      if (pass == 0 && !umma::mbar_wait(&tempty[tb], ((i >> 1) & 1) ^ 1)) { ok = false; break; }
      umma::tc_fence_after();
      if (umma::elect_one()) {
        const uint32_t a_hi = umma::smem_u32(smem) + sg * TILE, b_hi = a_hi + A_BYTES;
        const uint32_t a_lo = umma::smem_u32(lo_base) + sl * TILE, b_lo = a_lo + A_BYTES;
        const uint32_t dt = tm + tb * 64;
        const int ksteps = (min(BP, nrows - pass * BP) + 7) >> 3;
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint32_t adv = ks * 1024u;               // 8 batch rows = two 512-byte swizzle atoms
          const uint64_t dah = umma::smem_desc(a_hi + adv, A_BLK, 512, umma::kLayoutSw128Base32);
          const uint64_t dal = umma::smem_desc(a_lo + adv, A_BLK, 512, umma::kLayoutSw128Base32);
          const uint64_t dbh = umma::smem_desc(b_hi + adv, A_BLK, 512, umma::kLayoutSw128Base32);
          const uint64_t dbc = umma::smem_desc(b_hi + adv, b_lo - b_hi, 512, umma::kLayoutSw128Base32);
          umma::mma_tf32(dt, dah, dbc, idesc_cat, (pass > 0 || ks > 0) ? 1u : 0u);
          umma::mma_tf32(dt, dal, dbh, idesc, 1u);
        }
        umma::mma_commit(&rawfree[sg]);
        umma::mma_commit(&lofree[sl]);
        if (pass == npass - 1) umma::mma_commit(&tfull[tb]);
      }
      __syncwarp();
      if (lane == 0) MFAS_KSTAMP(16, f, 4);
    }
  } else {
    // ================================ Adam warps ==================================================
    const int q = warp & 3;                            // TMEM lane quarter this warp may read = 32 weight columns of the tile
    const int aw = warp - 13;
    const float inv_bc2 = 1.f / bc2_sqrt;
    uint8_t* ring = smem + (R + Cfg::LO) * TILE + (size_t)aw * NB * WS_SLOT;
    const uint32_t ring_u32 = umma::smem_u32(ring);
    const int srow = lane >> 3, schunk = lane & 7;
    const uint64_t stream_policy = l2_stream_policy((err.l2_hints & 2) != 0);
    struct Tile { float* W; long long K, moff, voff, goff; int rc; float gsc, gsign; int slot; };
    struct Raw { int4 a, b, c, d; };
    auto fetch_raw = [&](int i) {
      const int4* r = reinterpret_cast<const int4*>(&tiles[blockIdx.x + i * gridDim.x]);
      Raw o; o.a = __ldg(r); o.b = __ldg(r + 1); o.c = __ldg(r + 2);
      o.d = ALPHA ? __ldg(r + 4) : make_int4(0, 0, 0, 0);
      return o;
    };
    auto open_tile = [&](const Raw& r) {
      Tile o;
      const long long Wbits = ((long long)(uint32_t)r.a.y << 32) | (uint32_t)r.a.x;
      o.moff = ((long long)(uint32_t)r.a.w << 32) | (uint32_t)r.a.z;
      o.voff = ((long long)(uint32_t)r.b.y << 32) | (uint32_t)r.b.x;
      o.goff = KEEP_GRAD ? (((long long)(uint32_t)r.b.w << 32) | (uint32_t)r.b.z) : 0;
      o.K = r.c.x;
      o.rc = min(HN, max(0, r.c.z)) | (min(32, max(0, r.c.y - q * 32)) << 8);       // valid rows | valid columns of this warp's HN x 32
      o.W = reinterpret_cast<float*>(Wbits) + q * 32;
      o.gsc = 1.f; o.gsign = 0.f; o.slot = 0;
      if (ALPHA && r.d.z) {
        const float* ap = reinterpret_cast<const float*>(((long long)(uint32_t)r.d.y << 32) | (uint32_t)r.d.x);
        const float sg = gate_of(__ldg(ap));
        o.gsc = r.d.z == 1 ? sg : 1.0f - sg;
        o.gsign = r.d.z == 1 ? 1.f : -1.f;
        o.slot = r.d.w;
      }
      return o;
    };
    auto request = [&](const Tile& t, int j) {           // batch j (rows 8j .. 8j+7) of tile t into ring slot j; always commits a group
      if (t.rc >> 8) {
        const int rows = t.rc & 255, cols = t.rc >> 8;
        const uint32_t dst = ring_u32 + j * WS_SLOT + srow * 128 + schunk * 16;
        const float* w = t.W + (long long)(8 * j + srow) * t.K + schunk * 4;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          if (8 * j + 4 * u + srow < rows && schunk * 4 < cols) {
            const float* wu = w + (long long)(4 * u) * t.K;
            cp_async16(dst + u * 512, wu, stream_policy);
            cp_async16(dst + 1024 + u * 512, wu + t.moff, stream_policy);
            cp_async16(dst + 2048 + u * 512, wu + t.voff, stream_policy);
          }
        }
      }
      cp_async_commit();
    };
    Tile cur{}, nxt{};
    Raw ahead{};
    if (n_my > 0) {
      cur = open_tile(fetch_raw(0));
#pragma unroll
      for (int j = 0; j < NB; ++j) request(cur, j);
    }
    if (n_my > 1) ahead = fetch_raw(1);
    for (int i = 0; i < n_my; ++i) {
      const int tb = i & 1;
      const bool more = i + 1 < n_my;
      if (more) nxt = open_tile(ahead);
      if (i + 2 < n_my) ahead = fetch_raw(i + 2);
      if (!umma::mbar_wait(&tfull[tb], (i >> 1) & 1)) { ok = false; break; }
      if (aw == 0 && lane == 0) MFAS_KSTAMP(16, i, 5);
      umma::tc_fence_after();
      float dsum = 0.f;
#pragma unroll
      for (int j = 0; j < NB; ++j) {
        cp_async_wait<NB - 1>();
        __syncwarp();
        float g[8], g2[8];
        umma::tmem_ld8(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(tb * 64 + 8 * j), g);
        umma::tmem_ld8(tm + ((uint32_t)(q * 32) << 16) + (uint32_t)(tb * 64 + 32 + 8 * j), g2);
#pragma unroll
        for (int r = 0; r < 8; ++r) g[r] += g2[r];
        if (cur.rc >> 8) {
          const float* sp = reinterpret_cast<const float*>(ring + j * WS_SLOT) + lane;
          float* w = cur.W + (long long)(8 * j) * cur.K + lane;
          float p[8], m[8], v[8];
#pragma unroll
          for (int r = 0; r < 8; ++r) { p[r] = sp[r * 32]; m[r] = sp[256 + r * 32]; v[r] = sp[512 + r * 32]; }
          if (ALPHA) {
#pragma unroll
            for (int r = 0; r < 8; ++r) {
              if (8 * j + r < (cur.rc & 255) && lane < (cur.rc >> 8)) dsum = fmaf(p[r], g[r], dsum);
              g[r] *= cur.gsc;
            }
          }
#pragma unroll
          for (int r = 0; r < 8; ++r) adam_update_fast(g[r], p[r], m[r], v[r], adam, step_size, inv_bc2);
#pragma unroll
          for (int r = 0; r < 8; ++r) {
            if (8 * j + r < (cur.rc & 255) && lane < (cur.rc >> 8)) {
              if (KEEP_GRAD) w[cur.goff] = g[r];
              w[0] = p[r]; w[cur.moff] = m[r]; w[cur.voff] = v[r];
            }
            w += cur.K;
          }
        }
        __syncwarp();
        if (more) request(nxt, j); else cp_async_commit();
      }
      if (ALPHA && cur.gsign != 0.f) {
        dsum = warp_sum(dsum);
        if (lane == 0) dsp[(long long)cur.slot * TC_DSP_PER_TILE + aw] = cur.gsign * dsum;
      }
      umma::tc_fence_before();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(&tempty[tb]);
      if (aw == 0 && lane == 0) MFAS_KSTAMP(16, i, 6);
      cur = nxt;
    }
    cp_async_wait<0>();
  }
  if (!ok) atomicExch(err.flag, 5);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tm, TM_COLS);
}

// ---------------------------------------------------------------------------------------------
// alpha gates, tensor-core engine: d(alpha_l) = (sum of the per-tile, per-warp partials of k_tc_bwd_ws<., true>, fixed order)
// * s (1 - s), then Adam (the arithmetic of k_alpha_step).  One thread per (candidate, layer); rng[cand][layer] = {first
// tile, number of feature-column tiles} of the layer in the tile list.  Launched after the backward stream of every train step.
// ---------------------------------------------------------------------------------------------
__global__ void k_alpha_step_tc(const DCand* __restrict__ cands, int n_cand, const int2* __restrict__ rng, const float* __restrict__ dsp,
                                AdamH adam, float step_size, float bc2_sqrt) {
  const int cand = blockIdx.x * (blockDim.x / MFAS_MAX_LAYERS) + threadIdx.x / MFAS_MAX_LAYERS, l = threadIdx.x % MFAS_MAX_LAYERS;
  if (cand >= n_cand) return;
  const DCand& cd = cands[cand];
  if (l >= cd.L || !(cd.flags & MFAS_FLAG_ALPHAS)) return;
  const int2 r = rng[cand * MFAS_MAX_LAYERS + l];
  float ds = 0.f;
  for (int i = 0; i < r.y * TC_DSP_PER_TILE; ++i) ds += dsp[(long long)r.x * TC_DSP_PER_TILE + i];
  const DLayer& ly = cd.layer[l];
  float p = cd.p[ly.oalpha], m = cd.m[ly.oalpha], v = cd.v[ly.oalpha];
  const float sg = gate_of(p);
  const float g = ds * sg * (1.0f - sg);
  if (cd.grad) cd.grad[ly.oalpha] = g;
  adam_update(g, p, m, v, adam, step_size, bc2_sqrt);
  cd.p[ly.oalpha] = p; cd.m[ly.oalpha] = m; cd.v[ly.oalpha] = v;
}

}  // namespace mfas
