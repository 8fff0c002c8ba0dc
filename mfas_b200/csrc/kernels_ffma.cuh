// Engine "ffma": plain fp32 CUDA-core kernels, one launch per stage, batched over candidates
// (blockIdx.y = candidate).  Bit-for-bit deterministic (fixed-order reductions, no atomics).
// This is the first correct CUDA path and the bisecting reference for the tcgen05 engine.
//
// Per train step:  L x k_fusion_fwd  ->  k_head  ->  L x (k_dz -> k_fusion_bwd)
// Reference arithmetic: /root/reference/models/search/ntu_searchable.py:228-242 (forward),
// models/search/train_searchable/ntu.py:53-69 (loss, backward via autograd, Adam step).
#pragma once
#include "common.cuh"

namespace mfas {

constexpr int kThreads = 256;
constexpr int FWD_HT = 16;            // output columns per CTA
constexpr int FWD_KC = 32;            // K chunk (floats)
constexpr int FWD_LD = FWD_KC + 4;    // padded smem row, keeps 16-byte alignment
constexpr int BWD_KT = 32;            // weight columns per CTA in the backward

// ---------------------------------------------------------------------------------------------
// K1: fused fusion step forward.
//   x = concat(ske_tap[rows], rgb_tap[rows], h_{l-1})   (gathered on the fly, never materialised)
//   z = x W^T + b ; a = phi(z) ; h = BN(a) [; dropout]
// CTA = all batch rows x 16 output columns, so BatchNorm statistics stay inside the CTA.
// ---------------------------------------------------------------------------------------------
template <bool TRAIN>
__global__ void __launch_bounds__(kThreads)
k_fusion_fwd(const DCand* __restrict__ cands, DCache cache, BatchRef batch, int layer, int bmax,
             uint32_t drop_seed, float drop_p, uint32_t step) {
  const int cand = blockIdx.y;
  const DCand& cd = cands[cand];
  if (layer >= cd.L) return;
  const int H = cd.H;
  const int col0 = blockIdx.x * FWD_HT;
  if (col0 >= H) return;
  const DLayer& ly = cd.layer[layer];
  const int nrows = batch.n_rows;
  const int tid = threadIdx.x;

  __shared__ __align__(16) float Xs[MFAS_MAX_BATCH][FWD_LD];
  __shared__ __align__(16) float Ws[FWD_HT][FWD_LD];
  __shared__ float red[16][FWD_HT + 1];
  __shared__ int rowid[MFAS_MAX_BATCH];

  for (int r = tid; r < MFAS_MAX_BATCH; r += kThreads) rowid[r] = r < nrows ? batch_row(batch, cand, r) : 0;
  __syncthreads();

  // the three concat sources, in the reference's column order [ske | rgb | hidden]
  const float* seg_ptr[3];
  long long seg_ld[3];
  int seg_w[3];
  bool seg_gather[3];
  seg_ptr[0] = cache.ske[ly.ske_tap]; seg_ld[0] = cache.ske_ld[ly.ske_tap]; seg_w[0] = ly.d_ske; seg_gather[0] = true;
  seg_ptr[1] = cache.rgb[ly.rgb_tap]; seg_ld[1] = cache.rgb_ld[ly.rgb_tap]; seg_w[1] = ly.d_rgb; seg_gather[1] = true;
  seg_ptr[2] = layer > 0 ? cd.hid + (long long)(layer - 1) * bmax * H : nullptr;
  seg_ld[2] = H; seg_w[2] = ly.d_hid; seg_gather[2] = false;
  int nch[3];
  for (int s = 0; s < 3; ++s) nch[s] = (seg_w[s] + FWD_KC - 1) / FWD_KC;
  const int nchunks = nch[0] + nch[1] + nch[2];
  const float* Wbase = cd.p + ly.oW;
  const int K = ly.K;
  // AlphaScalarMultiplication (aux_models.py:103-111): ske * sigmoid(alpha), rgb * (1 - sigmoid(alpha))
  float seg_gate[3] = {1.f, 1.f, 1.f};
  if (cd.flags & MFAS_FLAG_ALPHAS) {
    const float sg = gate_of(cd.p[ly.oalpha]);
    seg_gate[0] = sg; seg_gate[1] = 1.0f - sg;
  }

  // register staging of the next chunk (global loads in flight while the current chunk computes)
  float4 xr[4];
  float4 wr;
  auto load_chunk = [&](int ci) {
    int s = 0, c = ci;
    if (c >= nch[0]) { c -= nch[0]; s = 1; if (c >= nch[1]) { c -= nch[1]; s = 2; } }
    const int k0 = c * FWD_KC;
    const int kw = min(FWD_KC, seg_w[s] - k0);
    int kglob = k0;
    if (s >= 1) kglob += seg_w[0];
    if (s >= 2) kglob += seg_w[1];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + kThreads * i;
      const int r = idx >> 3, c4 = (idx & 7) * 4;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r < nrows && c4 < kw) {
        const long long row = seg_gather[s] ? (long long)rowid[r] : (long long)r;
        val = __ldg(reinterpret_cast<const float4*>(seg_ptr[s] + row * seg_ld[s] + k0 + c4));
        const float gt = seg_gate[s];
        val.x *= gt; val.y *= gt; val.z *= gt; val.w *= gt;      // exact no-op when the gate is off (1.0f)
      }
      xr[i] = val;
    }
    wr = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < FWD_HT * 8) {
      const int hr = tid >> 3, c4 = (tid & 7) * 4;
      if (col0 + hr < H && c4 < kw)
        wr = *reinterpret_cast<const float4*>(Wbase + (long long)(col0 + hr) * K + kglob + c4);
    }
  };
  auto store_chunk = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + kThreads * i;
      *reinterpret_cast<float4*>(&Xs[idx >> 3][(idx & 7) * 4]) = xr[i];
    }
    if (tid < FWD_HT * 8) *reinterpret_cast<float4*>(&Ws[tid >> 3][(tid & 7) * 4]) = wr;
  };

  const int col = tid & (FWD_HT - 1);
  const int rg = tid >> 4;                       // 16 row groups; thread owns rows rg + 16*i
  const int nr = (nrows + 15) >> 4;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;

  load_chunk(0);
  for (int ci = 0; ci < nchunks; ++ci) {
    store_chunk();
    __syncthreads();
    if (ci + 1 < nchunks) load_chunk(ci + 1);
#pragma unroll
    for (int kk = 0; kk < FWD_KC; kk += 4) {
      const float4 w = *reinterpret_cast<const float4*>(&Ws[col][kk]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (i < nr) {
          const float4 x = *reinterpret_cast<const float4*>(&Xs[rg + 16 * i][kk]);
          acc[i] = fmaf(x.x, w.x, acc[i]);
          acc[i] = fmaf(x.y, w.y, acc[i]);
          acc[i] = fmaf(x.z, w.z, acc[i]);
          acc[i] = fmaf(x.w, w.w, acc[i]);
        }
      }
    }
    __syncthreads();
  }

  // ---- epilogue: bias, activation, BatchNorm over the batch, dropout -------------------------
  const int c = col0 + col;
  const bool cvalid = c < H;
  const float bias = cvalid ? cd.p[ly.ob + c] : 0.f;
  const bool bn = (cd.flags & MFAS_FLAG_BN) != 0;
  float a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = rg + 16 * i;
    a[i] = (r < nrows && cvalid) ? act_fwd(acc[i] + bias, ly.act) : 0.f;
  }
  float mean = 0.f, var = 1.f, istd = 1.f, gamma = 1.f, beta = 0.f;
  if (bn) {
    if (TRAIN) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) if (rg + 16 * i < nrows) s += a[i];
      red[rg][col] = s;
      __syncthreads();
      float tot = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) tot += red[g][col];
      mean = tot / (float)nrows;
      __syncthreads();
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) if (rg + 16 * i < nrows) { const float d = a[i] - mean; q = fmaf(d, d, q); }
      red[rg][col] = q;
      __syncthreads();
      tot = 0.f;
#pragma unroll
      for (int g = 0; g < 16; ++g) tot += red[g][col];
      var = tot / (float)nrows;
    } else if (cvalid) {
      mean = cd.bufs[ly.orm + c];
      var = cd.bufs[ly.orv + c];
    }
    istd = 1.f / sqrtf(var + kBnEps);
    if (cvalid) { gamma = cd.p[ly.og + c]; beta = cd.p[ly.obe + c]; }
    if (TRAIN && rg == 0 && cvalid) {
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
  const bool drop = TRAIN && (cd.flags & MFAS_FLAG_DROPOUT);
  const uint32_t dkey = drop ? dropout_key(drop_seed, (uint32_t)cd.cand_id, step, (uint32_t)layer) : 0u;
  const float dscale = drop ? 1.f / (1.f - drop_p) : 1.f;
  float* actp = cd.act + (long long)layer * bmax * H;
  float* hidp = cd.hid + (long long)layer * bmax * H;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = rg + 16 * i;
    if (r < nrows && cvalid) {
      float h = bn ? (a[i] - mean) * istd * gamma + beta : a[i];
      if (drop) h = dropout_keep(dkey, (uint32_t)(r * H + c), drop_p) ? h * dscale : 0.f;
      if (TRAIN) actp[r * H + c] = a[i];
      hidp[r * H + c] = h;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K3: classifier + softmax cross-entropy (+ its backward and the classifier's Adam step).
// One CTA of 512 threads per candidate; every inner product runs on float4 shared-memory reads laid
// out so that a warp either broadcasts an address or walks consecutive 16-byte words.
// dynamic smem: hs[bmax][hs_ld] | wcs[C][hs_ld] | lg[bmax][lg_ld]      (hs_ld = H+4 when it fits)
// ---------------------------------------------------------------------------------------------
constexpr int kHeadThreads = 512;

struct HeadOut {
  float* logits;          // [n_cand][bmax][C] or null
  float* loss;            // [n_cand] or null
  int* correct;           // [n_cand] or null
  double* stats;          // accumulators or null: stats[cand*stride + off] += loss*n, [+1] += correct
  long long stat_stride, stat_off;
};

// per-row log-softmax, NLL, argmax (first maximum), dlogits = (softmax - onehot)/n : one warp per row.
// lg[nrows][lg_ld] (shared) holds the logits and, when TRAIN, receives dlogits in place; dlog (global, [.][64], or null)
// receives a zero-padded copy for the tensor-core consumers.  blockDim.x == kHeadThreads, C <= 64.
template <bool TRAIN>
__device__ __forceinline__ void head_rows(const DCand& cd, const DCache& cache, int nrows, float* lg, int lg_ld,
                                          float* rowloss, int* rowok, const int* lab, const int* grow, float* dlog) {
  const int C = cd.C, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool multitask = (cd.flags & MFAS_FLAG_MULTITASK) && cache.logit_rgb && cache.logit_ske;
  for (int r = warp; r < nrows; r += kHeadThreads / 32) {
    float* row = lg + r * lg_ld;
    float v0 = lane < C ? row[lane] : -INFINITY, v1 = lane + 32 < C ? row[lane + 32] : -INFINITY;
    float mx = fmaxf(v0, v1);
    int am = (v1 > v0) ? lane + 32 : lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, am, o);
      if (om > mx || (om == mx && oa < am)) { mx = om; am = oa; }
    }
    const float e0 = lane < C ? expf(v0 - mx) : 0.f, e1 = lane + 32 < C ? expf(v1 - mx) : 0.f;
    const float lse = logf(warp_sum(e0 + e1));
    const int y = lab[r];
    float extra = 0.f;
    if (multitask) {
      const float* lr_ = cache.logit_rgb + (long long)grow[r] * C;
      const float* ls_ = cache.logit_ske + (long long)grow[r] * C;
      const float r0 = lane < C ? lr_[lane] : -INFINITY, r1 = lane + 32 < C ? lr_[lane + 32] : -INFINITY;
      const float s0 = lane < C ? ls_[lane] : -INFINITY, s1 = lane + 32 < C ? ls_[lane + 32] : -INFINITY;
      auto ce_of = [&](float a0, float a1) {                       // -log_softmax(a)[y], warp-wide
        float m = fmaxf(a0, a1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float e = (lane < C ? expf(a0 - m) : 0.f) + (lane + 32 < C ? expf(a1 - m) : 0.f);
        const float l2 = logf(warp_sum(e));
        const float ay = __shfl_sync(0xffffffffu, y < 32 ? a0 : a1, y & 31);
        return -((ay - m) - l2);
      };
      extra = ce_of(r0, r1) + ce_of(s0, s1);
      // preds = argmax(out + visual + skeleton), first maximum (torch.max(sum(output), 1))
      const float t0 = lane < C ? (v0 + r0) + s0 : -INFINITY, t1 = lane + 32 < C ? (v1 + r1) + s1 : -INFINITY;
      float tm = fmaxf(t0, t1);
      am = (t1 > t0) ? lane + 32 : lane;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, tm, o);
        const int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > tm || (om == tm && oa < am)) { tm = om; am = oa; }
      }
    }
    if (lane == 0) {
      rowloss[r] = -((row[y] - mx) - lse) + extra;
      rowok[r] = (am == y) ? 1 : 0;
    }
    __syncwarp();
    if (TRAIN) {
      const float inv_n = 1.f / (float)nrows;
      const float d0 = lane < C ? (expf((v0 - mx) - lse) - (lane == y ? 1.f : 0.f)) * inv_n : 0.f;
      const float d1 = lane + 32 < C ? (expf((v1 - mx) - lse) - (lane + 32 == y ? 1.f : 0.f)) * inv_n : 0.f;
      if (lane < C) row[lane] = d0;
      if (lane + 32 < C) row[lane + 32] = d1;
      if (dlog) { dlog[r * 64 + lane] = d0; dlog[r * 64 + 32 + lane] = d1; }     // zero-padded to 64 columns
    }
  }
}

// Multi-label head of the MM-IMDB fusion network, one warp per row, C <= 64:
//   loss   L[b][c] = q_c z (-log s) + (1 - z)(-log(1 - s)), s = sigmoid(x), through the explicit sigmoid and logs of
//          WeightedCrossEntropyWithLogits.forward (/root/reference/models/auxiliary/aux_models.py:136-146) -- NOT the
//          stable softplus form, so a saturated logit gives the same inf / nan the reference gives;
//   dlogits = (-q z (1 - s) + (1 - z) s) / (B C)   (mean over all B*C elements), written in place when TRAIN;
//   metric  sigmoid(x) > 0.3 against z: true positives and |pred| + |true| of the row, from which the caller forms the
//          per-sample F1 in fp64 (f1_score(average='samples'), train_searchable/mmimdb.py:84,101).
// rowloss[r] = sum_c L[r][c]; rowok[r] = 1 when the thresholded set equals the label set; tpden[r] = tp | (den << 8).
// dlog (global, [.][64], or null) receives a zero-padded copy of dlogits for the tensor-core consumers (see head_rows).
template <bool TRAIN>
__device__ __forceinline__ void head_rows_ml(const DCand& cd, const DCache& cache, int nrows, float* lg, int lg_ld,
                                             float* rowloss, int* rowok, int* tpden, const int* grow, float* dlog = nullptr) {
  const int C = cd.C, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float inv_bc = 1.f / ((float)nrows * (float)C);
  for (int r = warp; r < nrows; r += kHeadThreads / 32) {
    float* row = lg + r * lg_ld;
    const float* zt = cache.targets + (long long)grow[r] * C;
    float ls = 0.f;
    int tp = 0, den = 0, wrong = 0;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int c = lane + 32 * half;
      const bool ok = c < C;
      const float x = ok ? row[c] : 0.f, z = ok ? zt[c] : 0.f, q = ok ? cache.pos_weight[c] : 0.f;
      const float s = 1.f / (1.f + expf(-x));
      const float L = q * z * -logf(s) + (1.f - z) * -logf(1.f - s);
      ls += ok ? L : 0.f;
      const bool pred = ok && s > 0.3f, tru = ok && z > 0.5f;
      const unsigned pb = __ballot_sync(0xffffffffu, pred), tb = __ballot_sync(0xffffffffu, tru);
      tp += __popc(pb & tb);
      den += __popc(pb) + __popc(tb);
      wrong += __popc(pb ^ tb);
      const float dl = ok ? (-q * z * (1.f - s) + (1.f - z) * s) * inv_bc : 0.f;
      if (TRAIN && ok) row[c] = dl;
      if (TRAIN && dlog) dlog[r * 64 + c] = dl;                    // zero-padded to 64 columns
    }
    ls = warp_sum(ls);
    if (lane == 0) {
      rowloss[r] = ls;
      rowok[r] = wrong == 0 ? 1 : 0;
      tpden[r] = tp | (den << 8);
    }
    __syncwarp();
  }
}

// body shared by k_head and the fused chain kernel (kernels_tc.cuh: k_chain_all); blockDim.x == kHeadThreads
template <bool TRAIN, bool ML = false>
__device__ __forceinline__ void head_body(const DCand& cd, int cand, const DCache& cache, const BatchRef& batch, int bmax,
                                          int hs_ld, int lg_ld, const AdamH& adam, float step_size, float bc2_sqrt,
                                          const HeadOut& out, float* smem) {
  const int H = cd.H, C = cd.C, nrows = batch.n_rows, tid = threadIdx.x;
  const int H4 = H >> 2;
  float* hs = smem;                         // [bmax][hs_ld]
  float* wcs = hs + (size_t)bmax * hs_ld;   // [C][hs_ld]
  float* lg = wcs + (size_t)C * hs_ld;      // [bmax][lg_ld]
  __shared__ float rowloss[MFAS_MAX_BATCH];
  __shared__ int rowok[MFAS_MAX_BATCH], lab[MFAS_MAX_BATCH], grow[MFAS_MAX_BATCH];
  // multitask (train_searchable/ntu.py:59-61): loss = CE(fusion) + CE(rgb backbone) + CE(ske backbone), preds from
  // the sum of the three logit vectors.  The backbone logits are cached constants, so gradients are unchanged (head_rows).

  const float* hl = cd.hid + (long long)(cd.L - 1) * bmax * H;
  const float* Wc = cd.p + cd.oWc;
  for (int i = tid; i < nrows * H4; i += kHeadThreads) {
    const int r = i / H4, c4 = i % H4;
    *reinterpret_cast<float4*>(hs + r * hs_ld + c4 * 4) = *reinterpret_cast<const float4*>(hl + r * H + c4 * 4);
  }
  for (int i = tid; i < C * H4; i += kHeadThreads) {
    const int r = i / H4, c4 = i % H4;
    *reinterpret_cast<float4*>(wcs + r * hs_ld + c4 * 4) = *reinterpret_cast<const float4*>(Wc + r * H + c4 * 4);
  }
  for (int r = tid; r < nrows; r += kHeadThreads) {
    const int gr = batch_row(batch, cand, r);
    grow[r] = gr;
    lab[r] = ML ? 0 : (int)cache.labels[gr];
  }
  __syncthreads();

  // logits = h W_c^T + b_c (ntu_searchable.py:242): item = (class quad, batch row), rows run over lanes
  const int CQ = (C + 3) >> 2, bpad = (nrows + 31) & ~31;
  for (int it = tid; it < CQ * bpad; it += kHeadThreads) {
    const int cq = it / bpad, b = it % bpad;
    if (b >= nrows) continue;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    const float4* hp = reinterpret_cast<const float4*>(hs + b * hs_ld);
    const float4* wp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wp[i] = reinterpret_cast<const float4*>(wcs + min(cq * 4 + i, C - 1) * hs_ld);
    for (int h = 0; h < H4; ++h) {
      const float4 x = hp[h];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 w = wp[i][h];
        acc[i] = fmaf(x.x, w.x, acc[i]); acc[i] = fmaf(x.y, w.y, acc[i]);
        acc[i] = fmaf(x.z, w.z, acc[i]); acc[i] = fmaf(x.w, w.w, acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c = cq * 4 + i;
      if (c < C) {
        const float s = acc[i] + cd.p[cd.obc + c];
        lg[b * lg_ld + c] = s;
        cd.logits[b * C + c] = s;
        if (out.logits) out.logits[((long long)cand * bmax + b) * C + c] = s;
      }
    }
  }
  __syncthreads();

  if (ML) head_rows_ml<TRAIN>(cd, cache, nrows, lg, lg_ld, rowloss, rowok, lab, grow);
  else head_rows<TRAIN>(cd, cache, nrows, lg, lg_ld, rowloss, rowok, lab, grow, nullptr);
  __syncthreads();
  if (tid == 0) {
    float ls = 0.f;
    int ok = 0;
    double f1 = 0.0;                                              // ML: sum over rows of 2 tp / (|pred| + |true|), 0 when both are empty
    for (int r = 0; r < nrows; ++r) {
      ls += rowloss[r]; ok += rowok[r];
      if (ML) { const int tp = lab[r] & 255, den = lab[r] >> 8; f1 += den > 0 ? 2.0 * (double)tp / (double)den : 0.0; }
    }
    // CrossEntropyLoss(reduction='mean'); ML: torch.mean over all B*C elements (aux_models.py:146)
    const float mean_loss = ML ? ls / ((float)nrows * (float)C) : ls / (float)nrows;
    if (out.loss) out.loss[cand] = mean_loss;
    if (out.correct) out.correct[cand] = ok;
    if (out.stats) {                                              // running_loss += loss.item()*B (ntu.py:72-73, mmimdb.py:88)
      double* st = out.stats + (long long)cand * out.stat_stride + out.stat_off;
      st[0] += (double)mean_loss * (double)nrows;
      st[1] += ML ? f1 : (double)ok;
    }
  }
  if (!TRAIN) return;

  // dh_L = dlogits W_c (pre-update classifier in smem): item = (row, 4 columns), columns run over lanes
  float* dhp = cd.dh + (long long)(cd.L - 1) * bmax * H;
  for (int it = tid; it < nrows * H4; it += kHeadThreads) {
    const int b = it / H4, h4 = it % H4;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* dl = lg + b * lg_ld;
    for (int c = 0; c < C; ++c) {
      const float d = dl[c];
      const float4 w = *reinterpret_cast<const float4*>(wcs + c * hs_ld + h4 * 4);
      s.x = fmaf(d, w.x, s.x); s.y = fmaf(d, w.y, s.y); s.z = fmaf(d, w.z, s.z); s.w = fmaf(d, w.w, s.w);
    }
    *reinterpret_cast<float4*>(dhp + b * H + h4 * 4) = s;
  }
  // dW_c = dlogits^T h_L -> Adam : item = (class, 4 columns)
  for (int it = tid; it < C * H4; it += kHeadThreads) {
    const int c = it / H4, h4 = it % H4;
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < nrows; ++b) {
      const float d = lg[b * lg_ld + c];
      const float4 x = *reinterpret_cast<const float4*>(hs + b * hs_ld + h4 * 4);
      g.x = fmaf(d, x.x, g.x); g.y = fmaf(d, x.y, g.y); g.z = fmaf(d, x.z, g.z); g.w = fmaf(d, x.w, g.w);
    }
    const long long o = cd.oWc + (long long)c * H + h4 * 4;
    if (cd.grad) *reinterpret_cast<float4*>(cd.grad + o) = g;
    float4 p = *reinterpret_cast<float4*>(cd.p + o), m = *reinterpret_cast<float4*>(cd.m + o), v = *reinterpret_cast<float4*>(cd.v + o);
    adam_update(g.x, p.x, m.x, v.x, adam, step_size, bc2_sqrt);
    adam_update(g.y, p.y, m.y, v.y, adam, step_size, bc2_sqrt);
    adam_update(g.z, p.z, m.z, v.z, adam, step_size, bc2_sqrt);
    adam_update(g.w, p.w, m.w, v.w, adam, step_size, bc2_sqrt);
    *reinterpret_cast<float4*>(cd.p + o) = p; *reinterpret_cast<float4*>(cd.m + o) = m; *reinterpret_cast<float4*>(cd.v + o) = v;
  }
  for (int c = tid; c < C; c += kHeadThreads) {                   // db_c = sum_b dlogits
    float g = 0.f;
    for (int b = 0; b < nrows; ++b) g += lg[b * lg_ld + c];
    const long long o = cd.obc + c;
    if (cd.grad) cd.grad[o] = g;
    float p = cd.p[o], m = cd.m[o], v = cd.v[o];
    adam_update(g, p, m, v, adam, step_size, bc2_sqrt);
    cd.p[o] = p; cd.m[o] = m; cd.v[o] = v;
  }
}

template <bool TRAIN, bool ML = false>
__global__ void __launch_bounds__(kHeadThreads)
k_head(const DCand* __restrict__ cands, DCache cache, BatchRef batch, int bmax, int hs_ld, int lg_ld, AdamH adam,
       float step_size, float bc2_sqrt, HeadOut out) {
  extern __shared__ __align__(16) float smem[];
  head_body<TRAIN, ML>(cands[blockIdx.x], blockIdx.x, cache, batch, bmax, hs_ld, lg_ld, adam, step_size, bc2_sqrt, out, smem);
}

// ---------------------------------------------------------------------------------------------
// K2a: dL/dh_l -> dL/dz_l  (dropout mask, BatchNorm backward, activation backward), plus the
// gradients and Adam steps of the per-column vectors b_l, gamma_l, beta_l.
// CTA = 32 columns x 8 row groups.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_dz(const DCand* __restrict__ cands, int layer, int nrows, int bmax, AdamH adam, float step_size, float bc2_sqrt,
     uint32_t drop_seed, float drop_p, uint32_t step) {
  const DCand& cd = cands[blockIdx.y];
  if (layer >= cd.L) return;
  const int H = cd.H;
  const int c = blockIdx.x * 32 + (threadIdx.x & 31);
  if (blockIdx.x * 32 >= H) return;
  const bool cvalid = c < H;
  const int col = threadIdx.x & 31, rg = threadIdx.x >> 5;
  const DLayer& ly = cd.layer[layer];
  const bool bn = (cd.flags & MFAS_FLAG_BN) != 0;
  const bool drop = (cd.flags & MFAS_FLAG_DROPOUT) != 0;
  const uint32_t dkey = drop ? dropout_key(drop_seed, (uint32_t)cd.cand_id, step, (uint32_t)layer) : 0u;
  const float dscale = drop ? 1.f / (1.f - drop_p) : 1.f;
  const float* dhp = cd.dh + (long long)layer * bmax * H;
  const float* actp = cd.act + (long long)layer * bmax * H;
  __shared__ float red1[8][33], red2[8][33];

  auto dh_at = [&](int r) {
    float d = dhp[r * H + c];
    if (drop) d = dropout_keep(dkey, (uint32_t)(r * H + c), drop_p) ? d * dscale : 0.f;
    return d;
  };
  float mu = 0.f, istd = 1.f, gam = 1.f, m1 = 0.f, m2 = 0.f, S1 = 0.f, S2 = 0.f;
  if (bn) {
    if (cvalid) { mu = cd.mu[layer * H + c]; istd = cd.invstd[layer * H + c]; gam = cd.p[ly.og + c]; }
    float s1 = 0.f, s2 = 0.f;
    if (cvalid)
      for (int r = rg; r < nrows; r += 8) {
        const float d = dh_at(r);
        const float ah = (actp[r * H + c] - mu) * istd;
        s1 += d;
        s2 = fmaf(d, ah, s2);
      }
    red1[rg][col] = s1; red2[rg][col] = s2;
    __syncthreads();
#pragma unroll
    for (int g = 0; g < 8; ++g) { S1 += red1[g][col]; S2 += red2[g][col]; }
    m1 = S1 / (float)nrows; m2 = S2 / (float)nrows;
    __syncthreads();
  }
  float sdz = 0.f;
  if (cvalid)
    for (int r = rg; r < nrows; r += 8) {
      const float d = dh_at(r);
      const float a = actp[r * H + c];
      float da = d;
      if (bn) { const float ah = (a - mu) * istd; da = gam * istd * (d - m1 - ah * m2); }
      const float dz = da * act_bwd(a, ly.act);
      cd.dz[r * H + c] = dz;
      sdz += dz;
    }
  red1[rg][col] = sdz;
  __syncthreads();
  if (rg == 0 && cvalid) {
    float db = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) db += red1[g][col];
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

// ---------------------------------------------------------------------------------------------
// K2b + K4: dW_l[:, cols] = dz^T x[:, cols] with the Adam(L2) update fused into the epilogue
// (the gradient never reaches HBM), and for the hidden columns dh_{l-1} = dz W_l[:, cols]
// computed from the pre-update weights.  Feature columns need no dX.
// CTA = 32 weight columns x all H rows.  dynamic smem: dzs[bmax][H] | Xs[bmax][32] | Wsm[H][32]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads)
k_fusion_bwd(const DCand* __restrict__ cands, DCache cache, BatchRef batch, int layer, int bmax, AdamH adam,
             float step_size, float bc2_sqrt) {
  extern __shared__ __align__(16) float smem[];
  const int cand = blockIdx.y;
  const DCand& cd = cands[cand];
  if (layer >= cd.L) return;
  const DLayer& ly = cd.layer[layer];
  const int K = ly.K, H = cd.H;
  const int kc0 = blockIdx.x * BWD_KT;
  if (kc0 >= K) return;
  const int kw = min(BWD_KT, K - kc0);
  const int nrows = batch.n_rows, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* dzs = smem;                   // [bmax][H]
  float* Xs = dzs + bmax * H;          // [bmax][32]
  float* Wsm = Xs + bmax * BWD_KT;     // [H][32]

  for (int i = tid * 4; i < nrows * H; i += kThreads * 4)
    *reinterpret_cast<float4*>(dzs + i) = *reinterpret_cast<const float4*>(cd.dz + i);

  // which concat source do these columns come from?
  const float* src; long long ld; int k_local; bool gather = true;
  const int fs = ly.d_ske, fr = ly.d_rgb;
  if (kc0 < fs) { src = cache.ske[ly.ske_tap]; ld = cache.ske_ld[ly.ske_tap]; k_local = kc0; }
  else if (kc0 < fs + fr) { src = cache.rgb[ly.rgb_tap]; ld = cache.rgb_ld[ly.rgb_tap]; k_local = kc0 - fs; }
  else { src = cd.hid + (long long)(layer - 1) * bmax * H; ld = H; k_local = kc0 - fs - fr; gather = false; }
  const bool hidden = !gather;
  // alpha gate: the columns of this CTA belong to one segment; x is kept unscaled in smem, so acc below is the
  // unscaled G = dz^T x: dW = gate * G, and d(loss)/d(sigmoid) = +-sum(W o G) over the gated columns
  const bool gated = (cd.flags & MFAS_FLAG_ALPHAS) && !hidden;
  float gsc = 1.f, gsign = 0.f;
  if (gated) {
    const float sg = gate_of(cd.p[ly.oalpha]);
    if (kc0 < fs) { gsc = sg; gsign = 1.f; } else { gsc = 1.0f - sg; gsign = -1.f; }
  }
  float dsum = 0.f;
  for (int r = warp; r < nrows; r += 8) {
    const long long row = gather ? (long long)batch_row(batch, cand, r) : (long long)r;
    Xs[r * BWD_KT + lane] = lane < kw ? __ldg(src + row * ld + k_local + lane) : 0.f;
  }
  float* Wg = cd.p + ly.oW + kc0;
  if (hidden)
    for (int h = warp; h < H; h += 8) Wsm[h * BWD_KT + lane] = lane < kw ? Wg[(long long)h * K + lane] : 0.f;
  __syncthreads();

  if (hidden) {   // dh_{l-1}[b][k_local+lane] = sum_h dz[b][h] * W[h][kc0+lane]   (pre-update W)
    float* dprev = cd.dh + (long long)(layer - 1) * bmax * H;
    for (int b = warp; b < nrows; b += 8) {
      float s = 0.f;
      const float* dzr = dzs + b * H;
      for (int h = 0; h < H; ++h) s = fmaf(dzr[h], Wsm[h * BWD_KT + lane], s);
      if (lane < kw) dprev[b * H + k_local + lane] = s;
    }
  }

  float* Mg = cd.m + ly.oW + kc0;
  float* Vg = cd.v + ly.oW + kc0;
  float* Gg = cd.grad ? cd.grad + ly.oW + kc0 : nullptr;
  for (int h0 = warp * 16; h0 < H; h0 += 128) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int b = 0; b < nrows; ++b) {
      const float x = Xs[b * BWD_KT + lane];
      const float4* dz4 = reinterpret_cast<const float4*>(dzs + b * H + h0);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 d = dz4[q];
        acc[4 * q + 0] = fmaf(d.x, x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(d.y, x, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(d.z, x, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(d.w, x, acc[4 * q + 3]);
      }
    }
    if (lane < kw) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const long long o = (long long)(h0 + i) * K + lane;
        float p = Wg[o], m = Mg[o], v = Vg[o];
        dsum = fmaf(p, acc[i], dsum);
        const float gw = gated ? acc[i] * gsc : acc[i];
        if (Gg) Gg[o] = gw;
        adam_update(gw, p, m, v, adam, step_size, bc2_sqrt);
        Wg[o] = p; Mg[o] = m; Vg[o] = v;
      }
    }
  }
  if (gated) {     // (CTA-uniform) fixed-order reduction -> one partial per feature-column CTA; k_alpha_step finishes the sum.
    // Only the feature-column CTAs own a slot: blockIdx.x < (d_ske + d_rgb) / BWD_KT <= MFAS_DSP_SLOTS (mfas_group_create
    // rejects alpha groups with wider taps); the hidden-column CTAs write nothing.
    __shared__ float dsw[8];
    dsum = warp_sum(dsum);
    if (lane == 0) dsw[warp] = dsum;
    __syncthreads();
    if (tid == 0 && (int)blockIdx.x < MFAS_DSP_SLOTS) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += dsw[w];
      cd.dsp[layer * MFAS_DSP_SLOTS + blockIdx.x] = gsign * t;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// alpha gates: d(alpha_l) = (sum of the per-CTA partials of k_fusion_bwd, fixed order) * s (1 - s), then Adam.
// One thread per (candidate, layer); launched once per train step after every k_fusion_bwd.
// ---------------------------------------------------------------------------------------------
__global__ void k_alpha_step(const DCand* __restrict__ cands, int n_cand, AdamH adam, float step_size, float bc2_sqrt) {
  const int cand = blockIdx.x * (blockDim.x / MFAS_MAX_LAYERS) + threadIdx.x / MFAS_MAX_LAYERS, l = threadIdx.x % MFAS_MAX_LAYERS;
  if (cand >= n_cand) return;
  const DCand& cd = cands[cand];
  if (l >= cd.L || !(cd.flags & MFAS_FLAG_ALPHAS)) return;
  const DLayer& ly = cd.layer[l];
  const int nt = (ly.d_ske + ly.d_rgb + BWD_KT - 1) / BWD_KT;      // feature-column CTAs of k_fusion_bwd
  float ds = 0.f;
  for (int i = 0; i < nt; ++i) ds += cd.dsp[l * MFAS_DSP_SLOTS + i];
  float p = cd.p[ly.oalpha], m = cd.m[ly.oalpha], v = cd.v[ly.oalpha];
  const float sg = gate_of(p);
  const float g = ds * sg * (1.0f - sg);
  if (cd.grad) cd.grad[ly.oalpha] = g;
  adam_update(g, p, m, v, adam, step_size, bc2_sqrt);
  cd.p[ly.oalpha] = p; cd.m[ly.oalpha] = m; cd.v[ly.oalpha] = v;
}

// ---------------------------------------------------------------------------------------------
// K6: best-dev bookkeeping (train_searchable/ntu.py:82-86): strict '>' on fp64 accuracy, then a
// device-side snapshot of parameters + BN buffers; final rollback to the snapshot.
// ---------------------------------------------------------------------------------------------
__global__ void k_best_update(int n_cand, const double* stats, long long stat_stride, long long stat_off,
                              long long n_dev, int epoch, double* best_acc, int* best_epoch, int* improved) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cand) return;
  const double acc = stats[(long long)c * stat_stride + stat_off + 1] / (double)n_dev;
  if (acc > best_acc[c]) { best_acc[c] = acc; best_epoch[c] = epoch; improved[c] = 1; }
  else improved[c] = 0;
}

// ---------------------------------------------------------------------------------------------
// Initial weights of a whole group in one launch (mfas_group_init_params): counter-based draws keyed by (seed, candidate id,
// tensor, element) -- placement independent.  grid = (tiles, candidates); every thread walks the candidate's tensors.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float unit_uniform(uint32_t key, uint32_t i) {          // [0, 1): 24 bits
  return (float)(mix32(key + i * 0x9E3779B9u) >> 8) * (1.0f / 16777216.0f);
}
__global__ void __launch_bounds__(kThreads)
k_init_params(const DCand* __restrict__ cands, uint32_t seed_lo, uint32_t seed_hi) {
  const DCand& cd = cands[blockIdx.y];
  const uint32_t base = mix32(mix32(seed_lo ^ 0xA511E9B3u) + seed_hi * 0x85EBCA6Bu + (uint32_t)cd.cand_id * 0xC2B2AE35u);
  const long long t0 = (long long)blockIdx.x * kThreads + threadIdx.x, stride = (long long)gridDim.x * kThreads;
  const bool bn = (cd.flags & MFAS_FLAG_BN) != 0;
  auto uniform = [&](long long off, long long n, float bound, uint32_t tensor) {
    const uint32_t key = mix32(base + tensor * 0x27D4EB2Fu);
    for (long long i = t0; i < n; i += stride) cd.p[off + i] = bound * (2.0f * unit_uniform(key, (uint32_t)i) - 1.0f);
  };
  auto fill = [&](float* dst, long long off, long long n, float v) { for (long long i = t0; i < n; i += stride) dst[off + i] = v; };
  for (int l = 0; l < cd.L; ++l) {
    const DLayer& ly = cd.layer[l];
    const float bound = rsqrtf((float)ly.K);                      // kaiming_uniform_(a = sqrt(5)) and the bias bound: both 1 / sqrt(fan_in)
    uniform(ly.oW, (long long)cd.H * ly.K, bound, 4u * l);
    uniform(ly.ob, cd.H, bound, 4u * l + 1u);
    if (bn) {
      fill(cd.p, ly.og, cd.H, 1.f); fill(cd.p, ly.obe, cd.H, 0.f);
      fill(cd.bufs, ly.orm, cd.H, 0.f); fill(cd.bufs, ly.orv, cd.H, 1.f);
      if (t0 == 0) cd.nbt[l] = 0;
    }
    if (t0 == 0) {                                                // alpha ~ N(0, 0.1): Box-Muller on two draws
      const uint32_t key = mix32(base + (4u * l + 2u) * 0x27D4EB2Fu);
      const float u1 = 1.0f - unit_uniform(key, 0u), u2 = unit_uniform(key, 1u);
      cd.p[ly.oalpha] = 0.1f * sqrtf(-2.0f * logf(u1)) * cosf(6.2831853071795864f * u2);
    }
  }
  const float cb = rsqrtf((float)cd.H);
  uniform(cd.oWc, (long long)cd.C * cd.H, cb, 4u * MFAS_MAX_LAYERS);
  uniform(cd.obc, cd.C, cb, 4u * MFAS_MAX_LAYERS + 1u);
  for (long long i = t0; i < cd.n_params; i += stride) { cd.m[i] = 0.f; cd.v[i] = 0.f; }
}

// dir = 0: params -> best (when force or improved[c]);  dir = 1: best -> params
__global__ void __launch_bounds__(kThreads)
k_snapshot(const DCand* __restrict__ cands, const int* improved, int force, int dir) {
  const DCand& cd = cands[blockIdx.y];
  if (!force && !improved[blockIdx.y]) return;
  float* src_p = dir ? cd.best_p : cd.p;
  float* dst_p = dir ? cd.p : cd.best_p;
  float* src_b = dir ? cd.best_bufs : cd.bufs;
  float* dst_b = dir ? cd.bufs : cd.best_bufs;
  const long long n4 = cd.n_params >> 2;
  for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads)
    reinterpret_cast<float4*>(dst_p)[i] = reinterpret_cast<const float4*>(src_p)[i];
  if (blockIdx.x == 0) {
    for (long long i = threadIdx.x; i < cd.n_bufs; i += kThreads) dst_b[i] = src_b[i];
    if (threadIdx.x < cd.L) {
      if (dir) cd.nbt[threadIdx.x] = cd.best_nbt[threadIdx.x];
      else cd.best_nbt[threadIdx.x] = cd.nbt[threadIdx.x];
    }
  }
}

}  // namespace mfas
