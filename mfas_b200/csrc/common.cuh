// Device-side descriptors and small helpers shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mfas_b200.h"

namespace mfas {

constexpr float kBnEps = 1e-5f;        // torch.nn.BatchNorm1d default
constexpr float kBnMomentum = 0.1f;
constexpr float kLeakySlope = 0.01f;   // torch.nn.LeakyReLU default

// One fusion step of one candidate, as the kernels see it.
struct DLayer {
  int ske_tap, rgb_tap, act;
  int d_ske, d_rgb, d_hid;             // widths of the three concat sources (d_hid = 0 for step 0)
  int K;                               // d_ske + d_rgb + d_hid
  long long oW, ob, og, obe, oalpha;   // offsets into the parameter arena (og/obe < 0: no BN)
  long long orm, orv;                  // offsets into the buffer arena
};

struct DCand {
  int L, H, C, flags;
  int cand_id;                         // global candidate index (dropout key)
  int kb_item;                         // k-blocks per forward work item of this candidate's group (kernels_tc.cuh: tc_fwd_items)
  DLayer layer[MFAS_MAX_LAYERS];
  long long oWc, obc;
  long long n_params, n_bufs;
  // caller-owned arenas
  float *p, *m, *v, *grad, *bufs;
  long long* nbt;
  // library-owned: best-dev snapshot
  float *best_p, *best_bufs;
  long long* best_nbt;
  // library-owned workspace, all [L][Bmax][H] unless noted
  float* act;      // a_l  = phi(z_l)
  float* hid;      // h_l  = layer output (after BN / dropout)
  float* dh;       // dL/dh_l
  float* dz;       // [Bmax][H] dL/dz of the layer being back-propagated (ffma engine)
  float* dzs;      // [L][Bmax][H] dL/dz of every layer (tc engine: all layers stream in one launch)
  float* mu;       // [L][H] batch (or running) mean used by the last forward
  float* invstd;   // [L][H]
  float* logits;   // [Bmax][C]
  float* dsp;      // [L][MFAS_DSP_SLOTS] per-CTA partials of d(loss)/d(sigmoid(alpha_l)) (alpha gates, ffma engine)
  float* dlog;     // [Bmax][64] dL/dlogits, zero-padded (tc engine, tensor-core head: read by the last layer's backward and by the classifier tile of k_tc_bwd_ws)
};

constexpr int MFAS_DSP_SLOTS = 128;    // >= (widest ske tap + widest rgb tap) / 32

struct DCache {
  long long n_rows;
  const float* ske[MFAS_NUM_TAPS];
  const float* rgb[MFAS_NUM_TAPS];
  long long ske_ld[MFAS_NUM_TAPS];
  long long rgb_ld[MFAS_NUM_TAPS];
  const long long* labels;
  const float* logit_rgb;
  const float* logit_ske;
  const float* targets;        // [n_rows][C] multi-hot (MFAS_FLAG_MULTILABEL)
  const float* pos_weight;     // [C]
};

struct AdamH {
  float beta1, beta2, eps, wd;
  float one_minus_beta1, one_minus_beta2;
};

// Where a batch comes from: candidate c reads rows[c*stride + r], r < n_rows.
struct BatchRef {
  const int* rows;       // may be null => identity (row = base + r)
  long long stride;
  long long base;        // offset added when rows == null, or start inside the perm when not
  int n_rows;
};

__device__ __forceinline__ int batch_row(const BatchRef& b, int cand, int r) {
  if (b.rows == nullptr) return (int)(b.base + r);
  return b.rows[(long long)cand * b.stride + b.base + r];
}

// sigmoid(alpha) of the modality gate, the arithmetic of torch.sigmoid on fp32
__device__ __forceinline__ float gate_of(float alpha) { return 1.f / (1.f + expf(-alpha)); }

__device__ __forceinline__ float act_fwd(float z, int kind) {
  if (kind == MFAS_ACT_RELU) return fmaxf(z, 0.f);
  if (kind == MFAS_ACT_SIGMOID) return 1.f / (1.f + expf(-z));
  return z > 0.f ? z : kLeakySlope * z;
}
// phi'(z) through a = phi(z)
__device__ __forceinline__ float act_bwd(float a, int kind) {
  if (kind == MFAS_ACT_RELU) return a > 0.f ? 1.f : 0.f;
  if (kind == MFAS_ACT_SIGMOID) return a * (1.f - a);
  return a > 0.f ? 1.f : kLeakySlope;
}

// ---- counter-based dropout mask (mirrored bit-for-bit by oracle/mfas_oracle.py) -------------
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t dropout_key(uint32_t seed, uint32_t cand, uint32_t step,
                                                         uint32_t layer) {
  uint32_t k = mix32(seed ^ 0x9E3779B9u);
  k = mix32(k + cand * 0x85EBCA6Bu);
  k = mix32(k + step);
  k = mix32(k + layer * 0xC2B2AE35u);
  return k;
}
__device__ __forceinline__ bool dropout_keep(uint32_t key, uint32_t idx, float p) {
  uint32_t r = mix32(key + idx);
  float u = (float)(r >> 8) * (1.0f / 16777216.0f);
  return u >= p;
}

// Adam with coupled L2, the arithmetic of torch 2.11 _single_tensor_adam
// (grad.add(param, alpha=wd); exp_avg.lerp_; exp_avg_sq.mul_().addcmul_(); addcdiv_).
__device__ __forceinline__ void adam_update(float g, float& p, float& m, float& v, const AdamH& a,
                                            float step_size, float bc2_sqrt) {
  g = g + a.wd * p;
  m = m + a.one_minus_beta1 * (g - m);
  v = v * a.beta2 + a.one_minus_beta2 * g * g;
  float denom = sqrtf(v) / bc2_sqrt + a.eps;
  p = p - step_size * (m / denom);
}

// Same update with the two IEEE divisions and the IEEE sqrt replaced by sqrt.approx / rcp.approx and a
// host-side reciprocal of sqrt(1-beta2^t): ~12 instructions per element instead of ~50, which is what
// keeps the fused weight-streaming kernel HBM-bound rather than issue-bound.  Differs from
// adam_update by <= 3 ulp of the update term (tests hold it to 1e-6 of |p|).
__device__ __forceinline__ void adam_update_fast(float g, float& p, float& m, float& v, const AdamH& a,
                                                 float step_size, float inv_bc2_sqrt) {
  g = fmaf(a.wd, p, g);
  m = fmaf(a.one_minus_beta1, g - m, m);
  v = fmaf(a.one_minus_beta2 * g, g, v * a.beta2);
  float s, r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(s) : "f"(v));
  const float denom = fmaf(s, inv_bc2_sqrt, a.eps);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(denom));
  p = fmaf(-step_size * m, r, p);
}

// Programmatic dependent launch (the hot kernels of a step are launched with cudaLaunchAttributeProgrammaticStreamSerialization):
// griddep_launch() lets the NEXT kernel of the stream be scheduled on SMs as this grid's CTAs leave them -- its prologue
// (barrier init, TMEM allocation, descriptor fetch) then overlaps this grid's tail instead of waiting for a full drain plus a
// launch; griddep_wait() blocks until the PREVIOUS kernel of the stream has completed and its writes are visible.  Both are
// no-ops in a kernel launched without the attribute.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

}  // namespace mfas
