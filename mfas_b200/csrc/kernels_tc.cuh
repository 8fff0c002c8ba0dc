// Engine "tc": the two GEMM-shaped stages of a fusion step on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM), fp32-accurate through a 3xTF32 split
// (x = hi + lo; D = A_lo*B_hi + A_hi*B_lo + A_hi*B_hi, fp32 accumulate).  The problem is HBM-bound
// (10 FLOP/B), so operands are staged by coalesced vectorised loads through registers -- the gather of
// the batch rows, the concat bookkeeping and the hi/lo split all happen on that pass -- and stored into
// the canonical 128B-swizzled shared-memory tiles the tensor core reads.
//
//   forward :  zT[h, b] = sum_k W[h, k] * x[b, k]            A = W tile   (K-major), B = x tile (K-major)
//              split over K across CTAs; partial sums reduced, then bias/act/BN by k_tc_fwd_epi
//   backward:  dW[h, k] = sum_b dz[b, h] * x[b, k]           A = dz tile (MN-major), B = x tile (MN-major)
//              fused Adam(L2) epilogue: the gradient never reaches HBM; hidden columns also give dh_{l-1}
//
// Shapes: H in {64,128,256} (M tile 128, rows >= H are zero), batch <= 128.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace mfas {

constexpr int TC_THREADS = 256;
constexpr int TC_KB_PER_CTA = 16;     // forward: k-blocks (of 32 columns) per CTA = 512 columns of K
constexpr int TC_BWD_KT = 64;         // backward: weight columns per CTA
constexpr int TC_G_LD = 65;           // epilogue staging tile leading dimension (conflict-free transpose)

struct TcErr { int* flag; };          // set when a bounded barrier wait expires (never hangs the GPU)

__device__ __forceinline__ void store_split(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, float4 x) {
  float4 h, l;
  umma::split_tf32(x.x, h.x, l.x); umma::split_tf32(x.y, h.y, l.y);
  umma::split_tf32(x.z, h.z, l.z); umma::split_tf32(x.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi_tile + off) = h;
  *reinterpret_cast<float4*>(lo_tile + off) = l;
}

// ---------------------------------------------------------------------------------------------
// forward GEMM: partial zT over a K slice.  grid = (splits, H/128 m-tiles, candidates)
// NPAD = batch rows padded to the MMA N (64 or 128).
// dynamic smem (1024-aligned): A_hi 16K | A_lo 16K | B_hi NPAD*128 | B_lo NPAD*128
// ---------------------------------------------------------------------------------------------
template <int NPAD>
__global__ void __launch_bounds__(TC_THREADS)
k_tc_fwd(const DCand* __restrict__ cands, DCache cache, BatchRef batch, int layer, int bmax, float* part_base,
         long long part_stride_cand, TcErr err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int cand = blockIdx.z;
  const DCand& cd = cands[cand];
  if (layer >= cd.L) return;
  const int H = cd.H;
  const int m0 = blockIdx.y * 128;
  if (m0 >= H) return;
  const DLayer& ly = cd.layer[layer];
  const int K = ly.K, nkb = K >> 5;
  const int kb0 = blockIdx.x * TC_KB_PER_CTA;
  if (kb0 >= nkb) return;
  const int kb1 = min(nkb, kb0 + TC_KB_PER_CTA);
  const int nrows = batch.n_rows, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + 16384;
  uint8_t* b_hi = a_lo + 16384;
  uint8_t* b_lo = b_hi + NPAD * 128;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int rowid[MFAS_MAX_BATCH];

  if (warp == 0) umma::tmem_alloc(&tmem_slot, NPAD);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
  for (int r = tid; r < MFAS_MAX_BATCH; r += TC_THREADS) rowid[r] = r < nrows ? batch_row(batch, cand, r) : 0;
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tm = tmem_slot;

  const int fs = ly.d_ske, fr = ly.d_rgb;
  const float* Wbase = cd.p + ly.oW;
  const float* hid_prev = layer > 0 ? cd.hid + (long long)(layer - 1) * bmax * H : nullptr;

  // register staging of one k-block: A = 128 rows x 8 float4, B = NPAD rows x 8 float4
  constexpr int A_IT = 128 * 8 / TC_THREADS;     // 4
  constexpr int B_IT = NPAD * 8 / TC_THREADS;    // 2 or 4
  float4 ar[A_IT], br[B_IT];
  auto load_kb = [&](int kb) {
    const int kg = kb << 5;                      // first concat column of this k-block
    // which concat source? (segment widths are multiples of 32, so a k-block never straddles)
    const float* src; long long ld; int kl; bool gather = true;
    if (kg < fs) { src = cache.ske[ly.ske_tap]; ld = cache.ske_ld[ly.ske_tap]; kl = kg; }
    else if (kg < fs + fr) { src = cache.rgb[ly.rgb_tap]; ld = cache.rgb_ld[ly.rgb_tap]; kl = kg - fs; }
    else { src = hid_prev; ld = H; kl = kg - fs - fr; gather = false; }
#pragma unroll
    for (int i = 0; i < A_IT; ++i) {
      const int idx = tid + TC_THREADS * i, r = idx >> 3, c4 = idx & 7;
      ar[i] = (m0 + r < H) ? *reinterpret_cast<const float4*>(Wbase + (long long)(m0 + r) * K + kg + c4 * 4)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < B_IT; ++i) {
      const int idx = tid + TC_THREADS * i, r = idx >> 3, c4 = idx & 7;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows) {
        const long long row = gather ? (long long)rowid[r] : (long long)r;
        v = __ldg(reinterpret_cast<const float4*>(src + row * ld + kl + c4 * 4));
      }
      br[i] = v;
    }
  };
  auto store_kb = [&]() {
#pragma unroll
    for (int i = 0; i < A_IT; ++i) {
      const int idx = tid + TC_THREADS * i;
      store_split(a_hi, a_lo, umma::sw128(idx >> 3, (idx & 7) * 16), ar[i]);
    }
#pragma unroll
    for (int i = 0; i < B_IT; ++i) {
      const int idx = tid + TC_THREADS * i;
      store_split(b_hi, b_lo, umma::sw128(idx >> 3, (idx & 7) * 16), br[i]);
    }
  };

  constexpr uint32_t idesc = umma::idesc_tf32(128, NPAD, false, false);
  uint32_t phase = 0;
  bool ok = true;
  load_kb(kb0);
  for (int kb = kb0; kb < kb1; ++kb) {
    if (kb > kb0) {                                  // the tensor core must be done reading the tiles
      ok = umma::mbar_wait(&bar, phase);
      phase ^= 1;
      if (!ok) break;
    }
    store_kb();
    umma::fence_async_smem();
    __syncthreads();
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
    if (kb + 1 < kb1) load_kb(kb + 1);               // global loads fly while the MMAs run
  }
  if (ok) ok = umma::mbar_wait(&bar, phase);
  if (!ok && tid == 0) atomicExch(err.flag, 1);
  umma::tc_fence_after();

  // epilogue: thread = one output column h (TMEM lane), 32 consecutive batch rows per tcgen05.ld
  float* part = part_base + (long long)cand * part_stride_cand +
                ((long long)blockIdx.x * (((H + 127) >> 7) << 7) + m0) * NPAD;
  const int h_loc = (warp & 3) * 32 + lane;
#pragma unroll
  for (int c0 = (warp >> 2) * (NPAD / 2); c0 < (warp >> 2) * (NPAD / 2) + NPAD / 2; c0 += 32) {
    float v[32];
    umma::tmem_ld32(tm + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
    float4* dst = reinterpret_cast<float4*>(part + (long long)h_loc * NPAD + c0);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tm, NPAD);
}

// ---------------------------------------------------------------------------------------------
// forward epilogue: z = sum of the split-K partials (fixed order), + bias, activation, BatchNorm over
// the batch (train: batch statistics + running-stat update; eval: running statistics), dropout.
// grid = (H/32, candidates); a warp owns one output column at a time, lanes run over batch rows.
// ---------------------------------------------------------------------------------------------
template <bool TRAIN, int NPAD>
__global__ void __launch_bounds__(TC_THREADS)
k_tc_fwd_epi(const DCand* __restrict__ cands, int layer, int nrows, int bmax, const float* part_base,
             long long part_stride_cand, uint32_t drop_seed, float drop_p, uint32_t step) {
  const DCand& cd = cands[blockIdx.y];
  if (layer >= cd.L) return;
  const int H = cd.H;
  const int col0 = blockIdx.x * 32;
  if (col0 >= H) return;
  const DLayer& ly = cd.layer[layer];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsplit = ((ly.K >> 5) + TC_KB_PER_CTA - 1) / TC_KB_PER_CTA;
  const int Hp = ((H + 127) >> 7) << 7;
  const float* part = part_base + (long long)blockIdx.y * part_stride_cand;
  const bool bn = (cd.flags & MFAS_FLAG_BN) != 0;
  const bool drop = TRAIN && (cd.flags & MFAS_FLAG_DROPOUT);
  const uint32_t dkey = drop ? dropout_key(drop_seed, (uint32_t)cd.cand_id, step, (uint32_t)layer) : 0u;
  const float dscale = drop ? 1.f / (1.f - drop_p) : 1.f;
  constexpr int NJ = NPAD / 32;
  __shared__ float sa[MFAS_MAX_BATCH][33], sh[MFAS_MAX_BATCH][33];

  for (int cc = warp; cc < 32; cc += 8) {
    const int c = col0 + cc;
    if (c >= H) break;
    float z[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) z[j] = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float* p = part + ((long long)s * Hp + c) * NPAD;
#pragma unroll
      for (int j = 0; j < NJ; ++j) z[j] += p[lane + 32 * j];
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
        sa[r][cc] = a[j];
        sh[r][cc] = h;
      }
    }
  }
  __syncthreads();
  // coalesced write-out: a row of 32 columns = 128 bytes
  float* actp = cd.act + (long long)layer * bmax * H;
  float* hidp = cd.hid + (long long)layer * bmax * H;
  for (int r = warp; r < nrows; r += 8) {
    if (col0 + lane < H) {
      if (TRAIN) actp[r * H + col0 + lane] = sa[r][lane];
      hidp[r * H + col0 + lane] = sh[r][lane];
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward: dW[:, kc0:kc0+64] = dz^T x[:, kc0:kc0+64] on the tensor core, then, straight from TMEM:
// Adam(L2) on p/m/v (coalesced through a transposed staging tile), and for hidden columns
// dh_{l-1} = dz W[:, cols] from the pre-update weights.   grid = (ceil(Kmax/64), candidates)
// BP = batch rows padded to the tile (64 or 128); MMA K runs over the batch.
// dynamic smem (1024-aligned): A_hi | A_lo (4 blocks x BP x 128 B each) | 1 KB | B_hi | B_lo (2 blocks x BP x 128 B)
//                              after the MMAs the A region (+1 KB) is reused: G[128][65] staging, Wsm[128][64]
// ---------------------------------------------------------------------------------------------
template <int BP>
__global__ void __launch_bounds__(TC_THREADS)
k_tc_bwd(const DCand* __restrict__ cands, DCache cache, BatchRef batch, int layer, int bmax, AdamH adam,
         float step_size, float bc2_sqrt, TcErr err) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int cand = blockIdx.y;
  const DCand& cd = cands[cand];
  if (layer >= cd.L) return;
  const DLayer& ly = cd.layer[layer];
  const int K = ly.K, H = cd.H;
  const int kc0 = blockIdx.x * TC_BWD_KT;
  if (kc0 >= K) return;
  const int kw = min(TC_BWD_KT, K - kc0);
  const int nrows = batch.n_rows, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  constexpr uint32_t A_BLK = BP * 128, A_TILE = 4 * A_BLK, B_TILE = 2 * A_BLK;
  uint8_t* a_hi = smem;
  uint8_t* a_lo = a_hi + A_TILE;
  uint8_t* b_hi = a_lo + A_TILE + 1024;
  uint8_t* b_lo = b_hi + B_TILE;
  float* G = reinterpret_cast<float*>(smem);                        // [128][65]  aliases the A tiles (dead after the MMAs)
  float* Wsm = reinterpret_cast<float*>(smem + 128 * TC_G_LD * 4);  // [128][64]  aliases A_lo + the 1 KB gap; never the B tiles
  static_assert(128 * TC_G_LD * 4 + 128 * 64 * 4 <= 2 * A_TILE + 1024, "staging tiles must fit in the A region");
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;

  if (warp == 0) umma::tmem_alloc(&tmem_slot, 64);
  if (tid == 0) { umma::mbar_init(&bar, 1); umma::fence_mbar_init(); }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tm = tmem_slot;

  // ---- B tile: x[:, kc0:kc0+64] (MN-major: one row per batch row, 2 blocks of 32 columns) ----------
  const int fs = ly.d_ske, fr = ly.d_rgb;
  const float* src; long long ld; int kl; bool gather = true;
  if (kc0 < fs) { src = cache.ske[ly.ske_tap]; ld = cache.ske_ld[ly.ske_tap]; kl = kc0; }
  else if (kc0 < fs + fr) { src = cache.rgb[ly.rgb_tap]; ld = cache.rgb_ld[ly.rgb_tap]; kl = kc0 - fs; }
  else { src = cd.hid + (long long)(layer - 1) * bmax * H; ld = H; kl = kc0 - fs - fr; gather = false; }
  const bool hidden = !gather;
#pragma unroll
  for (int i = 0; i < BP * 16 / TC_THREADS; ++i) {
    const int idx = tid + TC_THREADS * i, r = idx >> 4, c4 = idx & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < nrows && c4 * 4 < kw) {
      const long long row = gather ? (long long)batch_row(batch, cand, r) : (long long)r;
      v = __ldg(reinterpret_cast<const float4*>(src + row * ld + kl + c4 * 4));
    }
    store_split(b_hi, b_lo, (uint32_t)(c4 >> 3) * A_BLK + umma::sw128_b32(r, (c4 & 7) * 16), v);
  }

  constexpr uint32_t idesc = umma::idesc_tf32(128, 64, true, true);
  const int ksteps = (nrows + 7) >> 3;
  uint32_t phase = 0;
  bool ok = true;
  float* Wg = cd.p + ly.oW + kc0;
  float* Mg = cd.m + ly.oW + kc0;
  float* Vg = cd.v + ly.oW + kc0;
  float* Gg = cd.grad ? cd.grad + ly.oW + kc0 : nullptr;

  for (int m0 = 0; m0 < H; m0 += 128) {
    // ---- A tile: dz[:, m0:m0+128] (MN-major: one row per batch row, 4 blocks of 32 columns) -------
#pragma unroll
    for (int i = 0; i < BP * 32 / TC_THREADS; ++i) {
      const int idx = tid + TC_THREADS * i, r = idx >> 5, c4 = idx & 31;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows && m0 + c4 * 4 < H) v = *reinterpret_cast<const float4*>(cd.dz + (long long)r * H + m0 + c4 * 4);
      store_split(a_hi, a_lo, (uint32_t)(c4 >> 3) * A_BLK + umma::sw128_b32(r, (c4 & 7) * 16), v);
    }
    umma::fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      umma::tc_fence_after();
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint32_t adv = ks * 1024u;             // 8 batch rows = two 512-byte atoms
        const uint64_t dah = umma::smem_desc(umma::smem_u32(a_hi) + adv, A_BLK, 512, umma::kLayoutSw128Base32);
        const uint64_t dal = umma::smem_desc(umma::smem_u32(a_lo) + adv, A_BLK, 512, umma::kLayoutSw128Base32);
        const uint64_t dbh = umma::smem_desc(umma::smem_u32(b_hi) + adv, A_BLK, 512, umma::kLayoutSw128Base32);
        const uint64_t dbl = umma::smem_desc(umma::smem_u32(b_lo) + adv, A_BLK, 512, umma::kLayoutSw128Base32);
        umma::mma_tf32(tm, dal, dbh, idesc, ks > 0 ? 1u : 0u);
        umma::mma_tf32(tm, dah, dbl, idesc, 1u);
        umma::mma_tf32(tm, dah, dbh, idesc, 1u);
      }
      umma::mma_commit(&bar);
    }
    ok = umma::mbar_wait(&bar, phase);
    phase ^= 1;
    if (!ok) break;
    umma::tc_fence_after();

    // ---- TMEM -> registers -> transposed staging tile G[h][col] ------------------------------------
    {
      float v[32];
      const int c0 = (warp >> 2) * 32, hl = (warp & 3) * 32 + lane;
      umma::tmem_ld32(tm + ((uint32_t)((warp & 3) * 32) << 16) + c0, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) G[hl * TC_G_LD + c0 + j] = v[j];
    }
    umma::tc_fence_before();
    // hidden columns: dh_{l-1}[b][kl+j] += sum_h dz[b][m0+h] * W[m0+h][kc0+j]   (pre-update weights)
    const int hrows = min(128, H - m0);
    if (hidden) {
      for (int i = tid; i < hrows * 16; i += TC_THREADS) {
        const int h = i >> 4, c4 = i & 15;
        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
        if (c4 * 4 < kw) w = *reinterpret_cast<const float4*>(Wg + (long long)(m0 + h) * K + c4 * 4);
        *reinterpret_cast<float4*>(Wsm + h * 64 + c4 * 4) = w;
      }
    }
    __syncthreads();
    if (hidden) {
      float* dprev = cd.dh + (long long)(layer - 1) * bmax * H;
      const int j = tid & 63, bg = tid >> 6;
      for (int b = bg; b < nrows; b += 4) {
        const float* dzr = cd.dz + (long long)b * H + m0;
        float s = 0.f;
        for (int h = 0; h < hrows; ++h) s = fmaf(dzr[h], Wsm[h * 64 + j], s);
        if (j < kw) {
          float* o = dprev + b * H + kl + j;
          *o = (m0 == 0) ? s : *o + s;
        }
      }
      __syncthreads();      // Wsm reads done before the weights below are overwritten in place
    }
    // ---- Adam(L2) straight from the staging tile; every p/m/v access is a coalesced 128-byte row ----
    for (int it = 0; it < hrows * 64 / TC_THREADS; it += 4) {
      float pp[4], mm[4], vv[4];
      long long off[4];
      bool act4[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = (it + u) * TC_THREADS + tid, r = idx >> 6, c = idx & 63;
        act4[u] = c < kw;
        off[u] = (long long)(m0 + r) * K + c;
        if (act4[u]) { pp[u] = Wg[off[u]]; mm[u] = Mg[off[u]]; vv[u] = Vg[off[u]]; }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (!act4[u]) continue;
        const int idx = (it + u) * TC_THREADS + tid, r = idx >> 6, c = idx & 63;
        const float g = G[r * TC_G_LD + c];
        if (Gg) Gg[off[u]] = g;
        adam_update(g, pp[u], mm[u], vv[u], adam, step_size, bc2_sqrt);
        Wg[off[u]] = pp[u]; Mg[off[u]] = mm[u]; Vg[off[u]] = vv[u];
      }
    }
    __syncthreads();        // G / Wsm (aliasing the A tiles) free before the next m-tile is staged
  }
  if (!ok && tid == 0) atomicExch(err.flag, 2);
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tm, 64);
}

}  // namespace mfas
