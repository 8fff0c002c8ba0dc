// C ABI of the B200-native MFAS candidate-training hot path (see include/mfas_b200.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <map>
#include <mutex>
#include <vector>

#include "kernels_ffma.cuh"
#include "kernels_tc.cuh"
#include "kernels_pool.cuh"

using namespace mfas;

// ---------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CUDA_TRY(expr)                                                                            \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return fail(MFAS_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// ---------------------------------------------------------------------------------------------
// TMA tensor maps (cuTensorMapEncodeTiled is a driver entry point: fetched at run time, nothing links libcuda)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TmapEncodeFn tmap_encoder() {
  static TmapEncodeFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    cudaGetLastError();
    return (TmapEncodeFn)p;
  }();
  return fn;
}
// row-major fp32 matrix [rows][cols] with row stride ld (floats) -> SWIZZLE_128B tiles of {32 columns, box_rows rows}
static bool encode_tile_map(CUtensorMap* out, const float* base, long long cols, long long rows, long long ld, int box_rows) {
  TmapEncodeFn enc = tmap_encoder();
  if (!enc) return false;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_rows}, es[2] = {1, 1};
  return enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" int mfas_abi_version(void) { return MFAS_ABI_VERSION; }
extern "C" const char* mfas_last_error(void) { return g_err; }

// ---------------------------------------------------------------------------------------------
// layout (host only)
// ---------------------------------------------------------------------------------------------
static inline int64_t align4(int64_t x) { return (x + 3) & ~int64_t(3); }

extern "C" int mfas_plan_layout(int32_t L, const int32_t* conf, int32_t H, int32_t C, int32_t flags,
                                const int32_t d_ske[MFAS_NUM_TAPS], const int32_t d_rgb[MFAS_NUM_TAPS],
                                mfas_layout* out) {
  if (!conf || !d_ske || !d_rgb || !out) return fail(MFAS_ERR_INVALID, "null argument");
  if (L < 1 || L > MFAS_MAX_LAYERS) return fail(MFAS_ERR_INVALID, "L=%d outside [1,%d]", L, MFAS_MAX_LAYERS);
  if (H < 16 || H > MFAS_MAX_HIDDEN || (H % 16) != 0)
    return fail(MFAS_ERR_INVALID, "inner_representation_size=%d must be a multiple of 16 in [16,%d]", H, MFAS_MAX_HIDDEN);
  if (C < 2 || C > MFAS_MAX_CLASSES) return fail(MFAS_ERR_INVALID, "num_outputs=%d outside [2,%d]", C, MFAS_MAX_CLASSES);
  const bool bn = flags & MFAS_FLAG_BN, drop = flags & MFAS_FLAG_DROPOUT;
  if ((flags & MFAS_FLAG_PLAIN) && bn)   // avmnist_searchable.py:276-285: the BatchNorm branches are commented out there
    return fail(MFAS_ERR_INVALID, "MFAS_FLAG_PLAIN (the AV-MNIST recipe Linear -> act [-> Dropout]) excludes MFAS_FLAG_BN");
  if (!bn && !drop && !(flags & MFAS_FLAG_PLAIN))   // ntu_searchable.py:274-284: no branch assigns `op` -> UnboundLocalError
    return fail(MFAS_ERR_UNSUPPORTED, "no layer recipe for drpt<1e-10 and batchnorm=False (reference raises UnboundLocalError)");
  for (int t = 0; t < MFAS_NUM_TAPS; ++t)
    if (d_ske[t] <= 0 || d_rgb[t] <= 0 || d_ske[t] % 32 || d_rgb[t] % 32)
      return fail(MFAS_ERR_INVALID, "tap widths must be positive multiples of 32 (ske[%d]=%d rgb[%d]=%d)", t, d_ske[t], t, d_rgb[t]);
  memset(out, 0, sizeof(*out));
  out->L = L; out->H = H; out->C = C; out->flags = flags;
  int64_t o = 0;
  for (int l = 0; l < L; ++l) {
    const int i = conf[3 * l + 0], j = conf[3 * l + 1], a = conf[3 * l + 2];
    if (i < 0 || i >= MFAS_NUM_TAPS || j < 0 || j >= MFAS_NUM_TAPS)
      return fail(MFAS_ERR_INVALID, "conf[%d] tap index out of range (%d,%d)", l, i, j);
    if (a < 0 || a > 2) return fail(MFAS_ERR_INVALID, "conf[%d] activation %d not in {0,1,2}", l, a);
    out->conf[l][0] = i; out->conf[l][1] = j; out->conf[l][2] = a;
    out->d_ske[l] = d_ske[i]; out->d_rgb[l] = d_rgb[j];
    out->K[l] = d_ske[i] + d_rgb[j] + (l > 0 ? H : 0);
    out->off_W[l] = o; o += align4((int64_t)H * out->K[l]);
    out->off_b[l] = o; o += align4(H);
    if (bn) { out->off_gamma[l] = o; o += align4(H); out->off_beta[l] = o; o += align4(H); }
    else { out->off_gamma[l] = -1; out->off_beta[l] = -1; }
  }
  out->off_Wc = o; o += align4((int64_t)C * H);
  out->off_bc = o; o += align4(C);
  for (int l = 0; l < L; ++l) { out->off_alpha[l] = o; o += 4; }
  out->n_params = o;
  int64_t b = 0;
  for (int l = 0; l < L; ++l) {
    if (bn) { out->off_rm[l] = b; b += align4(H); out->off_rv[l] = b; b += align4(H); }
    else { out->off_rm[l] = -1; out->off_rv[l] = -1; }
  }
  out->n_bufs = b > 0 ? b : 4;
  return MFAS_OK;
}

extern "C" int mfas_algorithmic_counts(const mfas_layout* lay, int32_t batch, double out[4]) {
  if (!lay || !out) return fail(MFAS_ERR_INVALID, "null argument");
  const double B = batch, H = lay->H, C = lay->C, L = lay->L;
  double F_sel = 0, sumK = 0, P = 0;
  for (int l = 0; l < lay->L; ++l) {
    F_sel += lay->d_ske[l] + lay->d_rgb[l];
    sumK += lay->K[l];
    P += (double)lay->K[l] * H + H;
  }
  if (lay->flags & MFAS_FLAG_BN) P += 2 * H * L;
  P += C * H + C;
  // labels: B int64 class ids, or (multi-label head) B x C fp32 targets + C positive-class weights
  const double lab = (lay->flags & MFAS_FLAG_MULTILABEL) ? 4 * (B * C + C) : 8 * B;
  out[0] = 4 * (B * F_sel + 6 * P) + lab;        // train step bytes
  out[1] = 4 * (B * F_sel + P) + lab;            // eval step bytes
  out[2] = 2 * B * (sumK * H + H * C);           // forward flops
  out[3] = out[2] + 2 * B * H * H * (L - 1) + 2 * B * H * C;
  return MFAS_OK;
}

// ---------------------------------------------------------------------------------------------
// Device-memory cache for the per-group allocations: workspace + partial sums (~0.8 GB for a 148-candidate cfg2 group)
// and the small descriptor / tile-list blocks.  A search driver creates and destroys a group per train_sampled_models
// call; cudaMalloc / cudaFree of these blocks cost 10-400 ms on the B200 box and single cudaFree calls up to 1.3 s
// (r01 e2e timings) -- more than all the per-call host work that remains.  Freed blocks are parked here and handed to
// the next group that fits; mfas_release_cached_memory() returns them to the driver.  MFAS_POOL_MB caps what is
// parked (default 4096, 0 disables).
// ---------------------------------------------------------------------------------------------
namespace {
struct PoolBlock { int device; void* p; size_t bytes; };
std::mutex g_pool_mu;
std::vector<PoolBlock> g_pool;
size_t pool_cap_bytes() {
  static const size_t cap = [] { const char* e = getenv("MFAS_POOL_MB"); return (size_t)(e ? atoll(e) : 4096) << 20; }();
  return cap;
}
cudaError_t pool_alloc(int device, size_t bytes, void** out, size_t* got) {
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    int best = -1;
    for (int i = 0; i < (int)g_pool.size(); ++i)
      if (g_pool[i].device == device && g_pool[i].bytes >= bytes && g_pool[i].bytes <= bytes + bytes / 4 + 65536 &&
          (best < 0 || g_pool[i].bytes < g_pool[best].bytes)) best = i;
    if (best >= 0) {
      *out = g_pool[best].p; *got = g_pool[best].bytes;
      g_pool.erase(g_pool.begin() + best);
      return cudaSuccess;
    }
  }
  *got = bytes;
  cudaError_t e = cudaMalloc(out, bytes);
  if (e == cudaErrorMemoryAllocation) {               // make room: drop everything parked on this device and retry
    cudaGetLastError();
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (int i = (int)g_pool.size() - 1; i >= 0; --i)
      if (g_pool[i].device == device) { cudaFree(g_pool[i].p); g_pool.erase(g_pool.begin() + i); }
    e = cudaMalloc(out, bytes);
  }
  return e;
}
// caller guarantees no work that touches the block is in flight (mfas_group_destroy synchronises the device first)
void pool_free(int device, void* p, size_t bytes) {
  if (!p) return;
  {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    size_t parked = 0;
    for (const PoolBlock& b : g_pool) parked += b.bytes;
    if (parked + bytes <= pool_cap_bytes() && g_pool.size() < 256) { g_pool.push_back({device, p, bytes}); return; }
  }
  cudaFree(p);
}
template <class T> cudaError_t pool_alloc_t(int device, size_t bytes, T** out, size_t* got) {
  void* p = nullptr;
  const cudaError_t e = pool_alloc(device, bytes, &p, got);
  *out = (T*)p;
  return e;
}
}  // namespace

extern "C" int mfas_release_cached_memory(void) {
  std::lock_guard<std::mutex> lk(g_pool_mu);
  int cur = 0;
  cudaGetDevice(&cur);
  for (const PoolBlock& b : g_pool) { cudaSetDevice(b.device); cudaFree(b.p); }
  g_pool.clear();
  cudaSetDevice(cur);
  return MFAS_OK;
}

// Tile list of the persistent backward (k_tc_bwd_ws): {candidate, layer, first weight column, first row} of every
// 128-column x 64-row tile of every fusion layer -- cut at the boundaries of the concat sources [ske | rgb | hidden], so that
// a tile reads its x columns from ONE source (a tap narrower than / not a multiple of 128 columns is a short tile) --
// followed, when the classifier head runs on the tensor core, by the classifier tiles (layer == L).  Returns the number
// of fusion-layer tiles.  Host only.
// ht: rows per tile -- TC_BWD_HT (64) for k_tc_bwd_ws, 16 / 32 for k_tc_bwd_small (whose classifier tiles are cut in ht-row pieces too)
static int build_bwd_tiles(const mfas_layout* lay, int n_cand, bool head_tiles, std::vector<int4>& tl, int ht = TC_BWD_HT) {
  for (int c = 0; c < n_cand; ++c)
    for (int l = 0; l < lay[c].L; ++l)
      for (int s0 = 0; s0 < lay[c].K[l]; s0 = tc_bwd_seg_end(lay[c].d_ske[l], lay[c].d_rgb[l], lay[c].K[l], s0))
        for (int kc0 = s0; kc0 < tc_bwd_seg_end(lay[c].d_ske[l], lay[c].d_rgb[l], lay[c].K[l], s0); kc0 += TC_BWD_KT)
          for (int h0 = 0; h0 < lay[c].H; h0 += ht) tl.push_back(make_int4(c, l, kc0, h0));
  const int n_layer_tiles = (int)tl.size();
  if (head_tiles)                                       // classifier tiles last: a step that ran k_head instead simply stops short of them
    for (int c = 0; c < n_cand; ++c)
      for (int kc0 = 0; kc0 < lay[c].H; kc0 += TC_BWD_KT)
        for (int c0 = 0; c0 < (ht < TC_BWD_HT ? lay[c].C : 1); c0 += ht) tl.push_back(make_int4(c, lay[c].L, kc0, c0));
  return n_layer_tiles;
}

extern "C" int mfas_plan_bwd_tiles(const mfas_layout* layouts, int32_t n_cand, int32_t head_tiles, int32_t* out, int64_t max_tiles,
                                   int64_t* n_tiles, int64_t* n_layer_tiles) {
  if (!layouts || n_cand < 1 || !n_tiles) return fail(MFAS_ERR_INVALID, "null argument / n_cand=%d", n_cand);
  std::vector<int4> tl;
  const int nl = build_bwd_tiles(layouts, n_cand, head_tiles != 0, tl);
  *n_tiles = (int64_t)tl.size();
  if (n_layer_tiles) *n_layer_tiles = nl;
  if (out) {
    if ((int64_t)tl.size() > max_tiles) return fail(MFAS_ERR_INVALID, "%zu tiles do not fit max_tiles=%lld", tl.size(), (long long)max_tiles);
    for (size_t i = 0; i < tl.size(); ++i) { out[4 * i] = tl[i].x; out[4 * i + 1] = tl[i].y; out[4 * i + 2] = tl[i].z; out[4 * i + 3] = tl[i].w; }
  }
  return MFAS_OK;
}

// ---------------------------------------------------------------------------------------------
// group
// ---------------------------------------------------------------------------------------------
struct mfas_group {
  int device = 0, n_cand = 0, bmax = 0;
  int Lmax = 0, Hmax = 0, Cmax = 0;
  int Kmax[MFAS_MAX_LAYERS] = {0};
  float drop_p = 0.f;
  uint32_t drop_seed = 0;
  AdamH adam;
  std::vector<mfas_layout> lay;
  std::vector<DCand> hc;          // host mirror of the device descriptors
  std::vector<char> bound;
  DCand* dc = nullptr;            // device descriptors
  char* ws = nullptr;             // one workspace allocation
  size_t ws_bytes = 0, part_bytes = 0, dc_bytes = 0, improved_bytes = 0, tiles_bytes = 0, items_bytes = 0, err_bytes = 0, tl_bytes = 0;   // block sizes as the pool handed them out
  int* improved = nullptr;        // [n_cand]
  bool dirty = true;
  size_t smem_head = 0, smem_bwd = 0;
  int hs_ld = 0, lg_ld = 0;       // leading dimensions of the head kernel's smem tiles
  int64_t launches = 0;
  // engine "tc" (tcgen05 tensor cores); engine 0 = "ffma"
  int engine = 0;
  bool multilabel = false;        // every candidate carries MFAS_FLAG_MULTILABEL (the MM-IMDB head)
  int npad = 64;                  // batch rows padded to the MMA tile (64 or 128)
  int items_fwd = 0, items_bwd = 0;   // work items per candidate (max over the group)
  float* part = nullptr;          // forward partial sums [n_cand][items_fwd][Hp][npad]
  long long part_stride = 0;
  int* tc_err = nullptr;          // device flag set by a timed-out barrier wait
  long long* timeline = nullptr;  // [n_cand][16] phase stamps of k_chain_all (debug: MFAS_CHAIN_TIMELINE=1)
  size_t smem_tc_fwd = 0, smem_tc_bwd = 0, smem_fl = 0, smem_dzx = 0, smem_chain = 0;
  BwdTile* bwd_tiles = nullptr;   // tile list of the persistent backward kernel (device), rebuilt when arenas are rebound
  std::vector<int4> bwd_tl;       // host: {cand, layer, first column, first row}
  int n_bwd_tiles = 0, n_bwd_layer_tiles = 0, n_sms = 148, bwd_ws = 1;
  FwdItem* fwd_items = nullptr;   // item list of the persistent forward kernel (device), rebuilt when arenas are rebound
  int n_fwd_items = 0, fwd_ws = 1, fwd_xr = 0;
  CUtensorMap* fwd_wmaps = nullptr; // one tensor map per forward item (its W tile rows), same order as fwd_items (device)
  size_t wmaps_bytes = 0;
  int bwd_small = 0;               // inner_repr <= 32: the persistent backward k_tc_bwd_small, rows per tile (16 / 32; 0 = k_tc_bwd_ws; MFAS_BWD_SMALL=0)
  int fwd_small = 1;               // inner_repr <= 32: the transposed forward stream k_tc_fwd_small (MFAS_FWD_SMALL=0: k_tc_fwd_ws with masked rows)
  int fwd_tma = 1;                 // forward stream operands: 1 = W tiles through TMA (cp.async.bulk.tensor.2d) + gathered x rows through
                                   // cp.async (default); 2 = x through tile::gather4 as well; 0 = cp.async loaders only (MFAS_FWD_TMA).
                                   // r02l on B200, 148 cfg2 candidates: forward stream 196 us (0) / 196 us (1) / 259 us (2) -- sixteen
                                   // 512-byte gather4 requests per k-block are slower than 512 cp.async chunks from 4 warps; the kernel
                                   // is bound by shared-memory bandwidth (144 KB of traffic per 24 KB k-block), not by the loaders' issue
  struct TapKey { const float* ske[MFAS_NUM_TAPS]; const float* rgb[MFAS_NUM_TAPS]; long long n_rows, ske_ld[MFAS_NUM_TAPS], rgb_ld[MFAS_NUM_TAPS]; };
  TapKey tap_key[4];               // small cache of feature-tap tensor maps: train / dev / test caches alternate
  TapMaps tap_maps[4];
  int tap_used[4] = {0, 0, 0, 0}, tap_clock = 0;
  FwdItem* fwd_items_ev = nullptr; // the same items with the partial-sum offsets of the 128-row eval layout
  size_t items_ev_bytes = 0;
  bool ev128 = false;             // dev / test passes run 128 rows per step (tc engine, batch <= 64): W is read half as often
  long long part_stride_ev = 0;
  bool any_alphas = false;
  float* dsp_tc = nullptr;        // alpha gates on the tc engine: [n_bwd_tiles][TC_DSP_PER_TILE] partials of d(loss)/d(sigmoid(alpha))
  int2* alpha_rng = nullptr;      // [n_cand][MFAS_MAX_LAYERS] {first tile, feature-column tiles} of every layer in the tile list
  size_t dsp_bytes = 0, rng_bytes = 0;
  int l2_hints = 9;               // bit 0: the forward's weight stream evict-first; bit 3: its gathered x rows evict-last (the backward stream of the same step re-reads them: r02gc)
  bool tchead = false;            // classifier head on the tensor core inside k_chain_all (+ its dW as a k_tc_bwd_ws tile)
  cudaEvent_t prof_ev[4] = {nullptr, nullptr, nullptr, nullptr};   // mfas_group_set_profiling: around fwd / chain / bwd of a step
  bool prof = false, prof_valid = false;
  int dbg = 0;                    // MFAS_TC_DEBUG bit 0: skip the Adam epilogue, bit 1: skip operand staging (timing experiments only)
  int kb_item = TC_KB_PER_ITEM;   // k-blocks per forward work item (smaller for small groups of the inner_repr 16 / 32 path)
  int chain_small_items = 0;      // partial sums per candidate staged by k_chain_small (largest of the group)
  int chain_small = 0;            // 16 / 32: the chain of a step runs in k_chain_small (inner_repr 16 / 32, on-chip, CUDA cores)
  int chain_cl = 1;               // CTAs (thread-block cluster size) per candidate in the fused chain: 2 for inner_repr 256 with <= n_sms / 2 candidates
  int chain = 2;                  // 2: fused chain kernel (H <= 128), 1: per-layer tensor-core chain kernels (MFAS_CHAIN=layers),
                                  // 0: per-layer CUDA-core chain kernels (MFAS_CHAIN=ffma)
  size_t smem_chain_all = 0;
};

static void set_adam(AdamH& a, double b1, double b2, double eps, double wd) {
  a.beta1 = (float)b1; a.beta2 = (float)b2; a.eps = (float)eps; a.wd = (float)wd;
  a.one_minus_beta1 = (float)(1.0 - b1);
  a.one_minus_beta2 = (float)(1.0 - b2);
}

extern "C" int mfas_group_set_adam(mfas_group_t g, const mfas_adam_hparams* hp) {
  if (!g || !hp) return fail(MFAS_ERR_INVALID, "null argument");
  // torch passes python doubles: 1-beta is formed in fp64 and then rounded to fp32
  set_adam(g->adam, hp->beta1, hp->beta2, hp->eps, hp->weight_decay);
  return MFAS_OK;
}

extern "C" int mfas_group_destroy(mfas_group_t g) {
  if (!g) return MFAS_OK;
  DeviceGuard dg(g->device);
  cudaDeviceSynchronize();                            // what cudaFree did implicitly: nothing of this group is in flight any more
  pool_free(g->device, g->dc, g->dc_bytes);
  pool_free(g->device, g->ws, g->ws_bytes);
  pool_free(g->device, g->improved, g->improved_bytes);
  pool_free(g->device, g->part, g->part_bytes);
  pool_free(g->device, g->bwd_tiles, g->tiles_bytes);
  pool_free(g->device, g->fwd_items, g->items_bytes);
  pool_free(g->device, g->fwd_items_ev, g->items_ev_bytes);
  pool_free(g->device, g->fwd_wmaps, g->wmaps_bytes);
  for (auto& ev : g->prof_ev) if (ev) cudaEventDestroy(ev);
  pool_free(g->device, g->tc_err, g->err_bytes);
  pool_free(g->device, g->timeline, g->tl_bytes);
  pool_free(g->device, g->dsp_tc, g->dsp_bytes);
  pool_free(g->device, g->alpha_rng, g->rng_bytes);
  delete g;
  return MFAS_OK;
}

// Forward work items of a small inner_repr-16 / 32 group: with items of up to 32 k-blocks, 32 two-step candidates are ~230 items on
// 148 persistent CTAs -- 82 CTAs get two, the launch takes two rounds.  The item size that minimises the busiest CTA's k-blocks
// (items dealt biggest first, round robin, as the launch does; 2 k-blocks of overhead per item) is chosen per group -- if the
// partial sums of the deepest candidate still fit k_chain_small's shared-memory stage.
static int choose_kb_item(const std::vector<mfas_layout>& lay, int n_cand, int n_sms, int Hmax, int Lmax, int npad) {
  auto sizes_of = [&](int per, std::vector<int>& sz, int& per_cand_max) {
    sz.clear(); per_cand_max = 0;
    for (int c = 0; c < n_cand; ++c) {
      int n_c = 0;
      const bool gated = (lay[c].flags & MFAS_FLAG_ALPHAS) != 0;
      for (int l = 0; l < lay[c].L; ++l) {
        const int ns = tc_fwd_items_g(lay[c].d_ske[l], lay[c].d_rgb[l], gated, per);
        for (int sp = 0; sp < ns; ++sp) {
          int kb0, kb1;
          tc_fwd_range_g(lay[c].d_ske[l], lay[c].d_rgb[l], sp, gated, kb0, kb1, per);
          sz.push_back(kb1 - kb0);
        }
        n_c += ns;
      }
      per_cand_max = std::max(per_cand_max, n_c);
    }
    std::sort(sz.begin(), sz.end(), std::greater<int>());
  };
  auto busiest = [&](const std::vector<int>& sz) {
    std::vector<int> load(n_sms, 0);
    for (size_t i = 0; i < sz.size(); ++i) load[i % n_sms] += sz[i] + 2;
    return *std::max_element(load.begin(), load.end());
  };
  auto fits = [&](int items) {
    const size_t tr = Hmax == 16 ? (npad == 64 ? ChainSmall<64, 16>::smem(Lmax, true, items) : ChainSmall<128, 16>::smem(Lmax, true, items))
                                 : (npad == 64 ? ChainSmall<64, 32>::smem(Lmax, true, items) : ChainSmall<128, 32>::smem(Lmax, true, items));
    const size_t ev = Hmax == 16 ? ChainSmall<128, 16>::smem(Lmax, false, items) : ChainSmall<128, 32>::smem(Lmax, false, items);
    return std::max(tr, ev) <= 216 * 1024;
  };
  std::vector<int> sz;
  int pc = 0;
  sizes_of(TC_KB_PER_ITEM, sz, pc);
  if ((int)sz.size() >= 6 * n_sms) return TC_KB_PER_ITEM;          // many items per CTA: the imbalance is a fraction of one item
  int best = TC_KB_PER_ITEM, best_load = busiest(sz);
  for (int per : {28, 24, 20, 16, 14, 12, 10, 8}) {
    sizes_of(per, sz, pc);
    if (!fits(pc)) continue;
    const int b = busiest(sz);
    if (b * 100 < best_load * 95) { best = per; best_load = b; }
  }
  return best;
}

extern "C" int mfas_group_create(int32_t device, int32_t n_cand, const mfas_layout* layouts, int32_t batch_max,
                                 float dropout_p, uint32_t dropout_seed, const int32_t* cand_ids,
                                 mfas_group_t* out) {
  if (!layouts || !out) return fail(MFAS_ERR_INVALID, "null argument");
  if (n_cand < 1) return fail(MFAS_ERR_INVALID, "n_cand=%d", n_cand);
  if (batch_max < 1 || batch_max > MFAS_MAX_BATCH)
    return fail(MFAS_ERR_INVALID, "batch_max=%d outside [1,%d]", batch_max, MFAS_MAX_BATCH);
  if (dropout_p < 0.f || dropout_p >= 1.f) return fail(MFAS_ERR_INVALID, "dropout p=%f outside [0,1)", dropout_p);
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(MFAS_ERR_INVALID, "device %d of %d", device, ndev);
  DeviceGuard dg(device);
  if (!dg.ok) return fail(MFAS_ERR_CUDA, "cudaSetDevice(%d) failed", device);

  mfas_group* g = new (std::nothrow) mfas_group();
  if (!g) return fail(MFAS_ERR_NOMEM, "host allocation failed");
  g->device = device; g->n_cand = n_cand; g->bmax = batch_max;
  g->drop_p = dropout_p; g->drop_seed = dropout_seed;
  set_adam(g->adam, 0.9, 0.999, 1e-8, 1e-4);
  g->lay.assign(layouts, layouts + n_cand);
  g->hc.resize(n_cand);
  g->bound.assign(n_cand, 0);

  // workspace carve-up
  auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
  size_t total = 0;
  std::vector<size_t> base(n_cand);
  struct Off { size_t act, hid, dh, dz, dzs, mu, invstd, logits, dsp, dlog, best_p, best_bufs, best_nbt; };
  std::vector<Off> off(n_cand);
  for (int c = 0; c < n_cand; ++c) {
    const mfas_layout& l = g->lay[c];
    if (l.L < 1 || l.L > MFAS_MAX_LAYERS || l.H < 16 || l.H > MFAS_MAX_HIDDEN || l.H % 16 || l.C < 2 ||
        l.C > MFAS_MAX_CLASSES || l.n_params <= 0) {
      delete g;
      return fail(MFAS_ERR_INVALID, "layout %d was not produced by mfas_plan_layout", c);
    }
    g->Lmax = l.L > g->Lmax ? l.L : g->Lmax;
    g->Hmax = l.H > g->Hmax ? l.H : g->Hmax;
    g->Cmax = l.C > g->Cmax ? l.C : g->Cmax;
    for (int k = 0; k < l.L; ++k) g->Kmax[k] = l.K[k] > g->Kmax[k] ? l.K[k] : g->Kmax[k];
    const size_t lbh = (size_t)l.L * batch_max * l.H * sizeof(float);
    const int rows_ev = batch_max <= 64 ? 128 : batch_max;       // dev passes may run 128 rows per step (eval rows are independent)
    Off& o = off[c];
    o.act = total; total += up(lbh);
    o.hid = total; total += up((size_t)l.L * rows_ev * l.H * sizeof(float));
    o.dh = total; total += up(lbh);
    o.dz = total; total += up((size_t)batch_max * l.H * sizeof(float));
    o.dzs = total; total += up(lbh);
    o.mu = total; total += up((size_t)l.L * l.H * sizeof(float));
    o.invstd = total; total += up((size_t)l.L * l.H * sizeof(float));
    o.logits = total; total += up((size_t)rows_ev * l.C * sizeof(float));
    o.dsp = total; total += up((size_t)MFAS_MAX_LAYERS * MFAS_DSP_SLOTS * sizeof(float));
    o.dlog = total; total += up((size_t)batch_max * TC_DLOG_LD * sizeof(float));
    o.best_p = total; total += up((size_t)l.n_params * sizeof(float));
    o.best_bufs = total; total += up((size_t)l.n_bufs * sizeof(float));
    o.best_nbt = total; total += up((size_t)MFAS_MAX_LAYERS * sizeof(long long));
  }
  cudaError_t e = pool_alloc(device, total, (void**)&g->ws, &g->ws_bytes);
  if (e == cudaSuccess) e = cudaMemset(g->ws, 0, total);
  if (e == cudaSuccess) e = pool_alloc_t(device, sizeof(DCand) * n_cand, &g->dc, &g->dc_bytes);
  if (e == cudaSuccess) e = pool_alloc_t(device, sizeof(int) * n_cand, &g->improved, &g->improved_bytes);
  if (e == cudaSuccess) e = cudaMemset(g->improved, 0, sizeof(int) * n_cand);
  if (e != cudaSuccess) {
    int code = fail(e == cudaErrorMemoryAllocation ? MFAS_ERR_NOMEM : MFAS_ERR_CUDA, "workspace allocation (%zu bytes): %s",
                    total, cudaGetErrorString(e));
    mfas_group_destroy(g);
    return code;
  }
  for (int c = 0; c < n_cand; ++c) {
    const mfas_layout& l = g->lay[c];
    DCand& d = g->hc[c];
    memset(&d, 0, sizeof(d));
    d.L = l.L; d.H = l.H; d.C = l.C; d.flags = l.flags;
    d.cand_id = cand_ids ? cand_ids[c] : c;
    d.kb_item = TC_KB_PER_ITEM;
    for (int k = 0; k < l.L; ++k) {
      DLayer& y = d.layer[k];
      y.ske_tap = l.conf[k][0]; y.rgb_tap = l.conf[k][1]; y.act = l.conf[k][2];
      y.d_ske = l.d_ske[k]; y.d_rgb = l.d_rgb[k]; y.d_hid = k > 0 ? l.H : 0; y.K = l.K[k];
      y.oW = l.off_W[k]; y.ob = l.off_b[k]; y.og = l.off_gamma[k]; y.obe = l.off_beta[k];
      y.oalpha = l.off_alpha[k]; y.orm = l.off_rm[k]; y.orv = l.off_rv[k];
    }
    d.oWc = l.off_Wc; d.obc = l.off_bc; d.n_params = l.n_params; d.n_bufs = l.n_bufs;
    const Off& o = off[c];
    d.act = (float*)(g->ws + o.act); d.hid = (float*)(g->ws + o.hid); d.dh = (float*)(g->ws + o.dh);
    d.dz = (float*)(g->ws + o.dz); d.dzs = (float*)(g->ws + o.dzs); d.mu = (float*)(g->ws + o.mu); d.invstd = (float*)(g->ws + o.invstd);
    d.logits = (float*)(g->ws + o.logits); d.dsp = (float*)(g->ws + o.dsp); d.dlog = (float*)(g->ws + o.dlog);
    d.best_p = (float*)(g->ws + o.best_p);
    d.best_bufs = (float*)(g->ws + o.best_bufs); d.best_nbt = (long long*)(g->ws + o.best_nbt);
  }
  // dynamic shared memory of the two big-smem kernels
  // head kernel tiles: rows padded by 4 floats (conflict-free float4 walks) when that still fits
  auto head_bytes = [&](int hs_ld, int lg_ld) {
    return sizeof(float) * ((size_t)batch_max * hs_ld + (size_t)g->Cmax * hs_ld + (size_t)batch_max * lg_ld);
  };
  g->hs_ld = g->Hmax + 4; g->lg_ld = g->Cmax + 1;
  if (head_bytes(g->hs_ld, g->lg_ld) > 220 * 1024) { g->hs_ld = g->Hmax; g->lg_ld = g->Cmax; }
  g->smem_head = head_bytes(g->hs_ld, g->lg_ld);
  g->smem_bwd = sizeof(float) * ((size_t)batch_max * g->Hmax + (size_t)batch_max * BWD_KT + (size_t)g->Hmax * BWD_KT);
  const size_t lim = 227 * 1024;
  if (g->smem_head > lim || g->smem_bwd > lim) {
    int code = fail(MFAS_ERR_UNSUPPORTED, "shared memory need (%zu / %zu B) exceeds 227 KB", g->smem_head, g->smem_bwd);
    mfas_group_destroy(g);
    return code;
  }
  e = cudaFuncSetAttribute(k_head<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_head);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_head<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_head);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_head<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_head);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_head<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_head);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_fusion_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g->smem_bwd);
  if (e != cudaSuccess) {
    int code = fail(MFAS_ERR_CUDA, "cudaFuncSetAttribute: %s (is this an sm_100 device?)", cudaGetErrorString(e));
    mfas_group_destroy(g);
    return code;
  }
  // ---- engine selection: tensor cores whenever the shapes fit the MMA tiles -----------------------
  auto l_small_ok = [](const mfas_layout& l) { return l.C <= TC_DLOG_LD; };     // small inner_repr needs the tensor-core head
  bool tc_ok = true, ragged = false;
  const bool stage_override = getenv("MFAS_BWD") || getenv("MFAS_FWD") || getenv("MFAS_CHAIN") || getenv("MFAS_HEAD");
  for (int c = 0; c < n_cand; ++c) {
    // inner_repr 16 / 32 (the search default) ride the same tiles with most rows masked; that needs the row / column
    // masks of the persistent kernels and the fused chain with the tensor-core head, i.e. none of the stage-by-stage overrides
    const bool small_ok = !stage_override && l_small_ok(g->lay[c]);
    tc_ok = tc_ok && (g->lay[c].H % 64 == 0 || ((g->lay[c].H == 16 || g->lay[c].H == 32) && small_ok));
    g->multilabel = g->multilabel || (g->lay[c].flags & MFAS_FLAG_MULTILABEL);
    if (((g->lay[c].flags ^ g->lay[0].flags) & MFAS_FLAG_MULTILABEL) ||
        ((g->lay[c].flags & MFAS_FLAG_MULTILABEL) && (g->lay[c].flags & MFAS_FLAG_MULTITASK))) {
      int code = fail(MFAS_ERR_INVALID, "MFAS_FLAG_MULTILABEL must be set for all candidates of a group or none, and excludes MFAS_FLAG_MULTITASK");
      mfas_group_destroy(g);
      return code;
    }
    g->any_alphas = g->any_alphas || (g->lay[c].flags & MFAS_FLAG_ALPHAS);
    if (g->lay[c].flags & MFAS_FLAG_ALPHAS)               // one d(sigmoid(alpha)) partial per 32 feature columns (DCand::dsp)
      for (int l = 0; l < g->lay[c].L; ++l)
        if ((g->lay[c].d_ske[l] + g->lay[c].d_rgb[l] + BWD_KT - 1) / BWD_KT > MFAS_DSP_SLOTS) {
          int code = fail(MFAS_ERR_UNSUPPORTED, "alpha gates: the taps of candidate %d step %d are %d + %d columns wide, more than the %d the gate-gradient partials cover",
                          c, l, g->lay[c].d_ske[l], g->lay[c].d_rgb[l], MFAS_DSP_SLOTS * BWD_KT);
          mfas_group_destroy(g);
          return code;
        }
    for (int l = 0; l < g->lay[c].L; ++l) ragged = ragged || g->lay[c].d_ske[l] % 128 || g->lay[c].d_rgb[l] % 128;
  }
  // Tap widths that are not multiples of the 128-column backward tile (the 64-wide MM-IMDB text tap): only the persistent
  // backward walks a host-built tile list, which is cut at the concat-source boundaries; the grid-indexed backward is not.
  if (ragged && (getenv("MFAS_BWD") || getenv("MFAS_FWD"))) tc_ok = false;
  // the modality gates ride the persistent kernels (item / tile lists cut at the source boundaries) and the fused chain
  if (g->any_alphas && stage_override) tc_ok = false;
  const char* env = getenv("MFAS_ENGINE");
  if (env && !strcmp(env, "ffma")) tc_ok = false;
  if (env && !strcmp(env, "tc") && !tc_ok) {
    int code = fail(MFAS_ERR_UNSUPPORTED, "MFAS_ENGINE=tc needs inner_representation_size 16, 32 or a multiple of 64 (16 / 32 and the alpha gates: without the stage-by-stage overrides MFAS_FWD / MFAS_BWD / MFAS_CHAIN / MFAS_HEAD)");
    mfas_group_destroy(g);
    return code;
  }
  if (tc_ok) {
    g->engine = 1;
    g->npad = batch_max <= 64 ? 64 : 128;
    { int nsm = 0; if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && nsm > 0) g->n_sms = nsm; }
    {
      bool same = true;
      for (int c = 0; c < n_cand; ++c) same = same && g->lay[c].H == g->Hmax;
      // MFAS_KB_ITEM: a fixed item size, or "auto" = choose_kb_item.  NOT the default: the split of a layer's columns decides the
      // order in which its partial products are rounded, so a size chosen from the group would make a candidate's numbers depend
      // on how many candidates share the GPU -- and 1-GPU, N-rank and fan-out runs are bit-identical by contract (measured with
      // "auto": forward stream 20.3 -> 18.6 us at 32 candidates, 21.6 -> 15.9 us at 32 one-step candidates, r02fo).
      const char* ke = getenv("MFAS_KB_ITEM");
      if (ke && !strcmp(ke, "auto")) { if (same && (g->Hmax == 16 || g->Hmax == 32)) g->kb_item = choose_kb_item(g->lay, n_cand, g->n_sms, g->Hmax, g->Lmax, g->npad); }
      else if (ke && atoi(ke) >= 4 && atoi(ke) <= TC_KB_PER_ITEM) g->kb_item = atoi(ke);
      for (int c = 0; c < n_cand; ++c) g->hc[c].kb_item = g->kb_item;
    }
    for (int c = 0; c < n_cand; ++c) {
      int nf = 0, nb = 0;
      for (int l = 0; l < g->lay[c].L; ++l) {
        nf += tc_fwd_items_g(g->lay[c].d_ske[l], g->lay[c].d_rgb[l], (g->lay[c].flags & MFAS_FLAG_ALPHAS) != 0, g->kb_item);
        nb += tc_bwd_items(g->lay[c].K[l]);
      }
      g->items_fwd = nf > g->items_fwd ? nf : g->items_fwd;
      g->items_bwd = nb > g->items_bwd ? nb : g->items_bwd;
    }
    const int Hp = ((g->Hmax + 127) / 128) * 128;
    g->part_stride = (long long)g->items_fwd * Hp * g->npad;
    g->smem_tc_fwd = 1024 + 32768 + 2 * (size_t)g->npad * 128;
    g->smem_tc_bwd = 1024 + 2 * (size_t)(4 * g->npad * 128) + 2 * (size_t)(2 * g->npad * 128);
    g->smem_fl = sizeof(float) * ((size_t)batch_max * (g->Hmax + 1) + TC_CB * (size_t)g->Hmax);
    g->smem_dzx = sizeof(float) * ((size_t)batch_max * g->Hmax + (TC_CB + 1) * (size_t)g->Hmax);
    e = pool_alloc_t(device, sizeof(int), &g->tc_err, &g->err_bytes);
    if (e == cudaSuccess) e = cudaMemset(g->tc_err, 0, sizeof(int));
    // The limit is a property of the FUNCTION (per device), not of the group: groups whose kernels need different amounts
    // (k_chain_small's depends on the deepest candidate, k_chain_all's on the head) may be alive together, on any host thread
    // (fan-out) -- so the limit only ever grows: the largest request seen for (device, function) stays in force.
    auto attr = [&](const void* f, size_t bytes) {
      static std::mutex mu;
      static std::map<std::pair<int, const void*>, size_t> seen;
      std::lock_guard<std::mutex> lk(mu);
      size_t& cur = seen[std::make_pair(device, f)];
      if (bytes > cur && e == cudaSuccess) {
        e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess) cur = bytes;
      }
    };
    attr((const void*)k_tc_fwd_all<64>, 1024 + 32768 + 2 * 64 * 128);
    attr((const void*)k_tc_fwd_all<128>, 1024 + 32768 + 2 * 128 * 128);
    attr((const void*)k_tc_bwd_all<64, false>, 1024 + 6 * 64 * 128 * 2);
    attr((const void*)k_tc_bwd_all<64, true>, 1024 + 6 * 64 * 128 * 2);
    attr((const void*)k_tc_bwd_all<128, false>, 1024 + 6 * 128 * 128 * 2);
    attr((const void*)k_tc_bwd_all<128, true>, 1024 + 6 * 128 * 128 * 2);
    attr((const void*)k_fwd_layer<true, 64>, g->smem_fl);
    attr((const void*)k_fwd_layer<false, 64>, g->smem_fl);
    attr((const void*)k_fwd_layer<true, 128>, g->smem_fl);
    attr((const void*)k_fwd_layer<false, 128>, g->smem_fl);
    attr((const void*)k_dzx, g->smem_dzx);
    g->smem_chain = g->npad == 64 ? ChainCfg<64>::SMEM : ChainCfg<128>::SMEM;
    attr((const void*)k_chain_fwd<true, 64>, ChainCfg<64>::SMEM);
    attr((const void*)k_chain_fwd<false, 64>, ChainCfg<64>::SMEM);
    attr((const void*)k_chain_fwd<true, 128>, ChainCfg<128>::SMEM);
    attr((const void*)k_chain_fwd<false, 128>, ChainCfg<128>::SMEM);
    attr((const void*)k_chain_bwd<64>, ChainCfg<64>::SMEM);
    attr((const void*)k_chain_bwd<128>, ChainCfg<128>::SMEM);
    { int nsm = 0; if (e == cudaSuccess) e = cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device); if (nsm > 0) g->n_sms = nsm; }
    { const char* ce = getenv("MFAS_CHAIN"); if (ce && !strcmp(ce, "ffma")) g->chain = 0; if (ce && !strcmp(ce, "layers")) g->chain = 1; }
    // the fused kernel walks the 128-column tiles of a layer one after the other (inner_repr a power of two up to 256:
    // its L2 prefetch indexes lines by shifts); inner_repr 192 takes the per-layer chain kernels
    if ((g->Hmax > 256 || (g->Hmax & (g->Hmax - 1))) && g->chain == 2) g->chain = 1;
    { const char* be = getenv("MFAS_BWD"); if (be && !strcmp(be, "cta")) g->bwd_ws = 0; }
    // tensor-core head: needs the fused chain (its tiles and TMEM), the persistent backward (the classifier's dW + Adam
    // become one of its tiles) and C <= 64 (one 64-row tile; head_rows' two values per lane)
    g->tchead = g->chain == 2 && g->bwd_ws && g->Cmax <= TC_DLOG_LD;
    { const char* he = getenv("MFAS_HEAD"); if (he && !strcmp(he, "ffma")) g->tchead = false; }
    // dynamic shared memory of the fused chain: its operand tiles, or -- only with the CUDA-core head inside it -- that head's
    // tiles if they are larger; ~12 KB of the 227 KB are the kernel's static arrays (descriptor, row state, reductions)
    g->smem_chain_all = (g->tchead || g->smem_chain > 1024 + g->smem_head) ? g->smem_chain : 1024 + g->smem_head;
    if (g->smem_chain_all > 214 * 1024 && g->chain == 2) { g->chain = 1; g->tchead = false; }
    if (g->chain == 2 && g->Hmax > 128 && 2 * n_cand <= g->n_sms) {
      const char* ce = getenv("MFAS_CHAIN_CLUSTER");
      g->chain_cl = (ce && !atoi(ce)) ? 1 : 2;
    }
    // inner_repr 16 / 32 with the tensor-core head's hand-off (dlogits to the backward stream): the on-chip chain
    if (g->chain == 2 && g->tchead && (g->Hmax == 16 || g->Hmax == 32)) {
      bool same = true;
      for (int c = 0; c < n_cand; ++c) same = same && g->lay[c].H == g->Hmax;
      const char* se = getenv("MFAS_CHAIN_SMALL");
      int items = 0;                                      // most partial sums of a candidate (all staged in shared memory)
      for (int c = 0; c < n_cand; ++c) {
        int n = 0;
        for (int l = 0; l < g->lay[c].L; ++l) n += tc_fwd_items_g(g->lay[c].d_ske[l], g->lay[c].d_rgb[l], (g->lay[c].flags & MFAS_FLAG_ALPHAS) != 0, g->kb_item);
        items = std::max(items, n);
      }
      auto need_of = [&](int npad, bool tr, int it) {
        return g->Hmax == 16 ? (npad == 64 ? ChainSmall<64, 16>::smem(g->Lmax, tr, it) : ChainSmall<128, 16>::smem(g->Lmax, tr, it))
                             : (npad == 64 ? ChainSmall<64, 32>::smem(g->Lmax, tr, it) : ChainSmall<128, 32>::smem(g->Lmax, tr, it));
      };
      // training steps at the group's row padding; a 128-row dev / test step.  Partial sums that do not fit are read from global memory.
      auto need_all = [&](int it) { return std::max(need_of(g->npad, true, it), need_of(128, false, it)); };
      { const char* st = getenv("MFAS_CHAIN_SMALL_STAGE"); if (st && !atoi(st)) items = 0; }      // (test switch: the global-memory path)
      if (need_all(items) > 216 * 1024) items = 0;
      g->chain_small_items = items;
      const size_t need = need_all(items);
      if (same && !(se && !atoi(se)) && need <= 216 * 1024) {
        g->chain_small = g->Hmax;
#define CS_ATTR(T, N, HN_) attr((const void*)k_chain_small<T, N, HN_, false>, ChainSmall<N, HN_>::smem(g->Lmax, T, g->chain_small_items)); attr((const void*)k_chain_small<T, N, HN_, true>, ChainSmall<N, HN_>::smem(g->Lmax, T, g->chain_small_items));
        // the variants this group launches: training and dev steps at its own row padding, 128-row dev / test steps
        if (g->Hmax == 16) { if (g->npad == 64) { CS_ATTR(true, 64, 16) CS_ATTR(false, 64, 16) } else { CS_ATTR(true, 128, 16) } CS_ATTR(false, 128, 16) }
        else { if (g->npad == 64) { CS_ATTR(true, 64, 32) CS_ATTR(false, 64, 32) } else { CS_ATTR(true, 128, 32) } CS_ATTR(false, 128, 32) }
#undef CS_ATTR
      }
    }
    if (g->chain == 2) {
#define CHAIN_ATTR(ML) \
      attr((const void*)k_chain_all<true, 64, false, ML>, g->smem_chain_all); attr((const void*)k_chain_all<false, 64, false, ML>, g->smem_chain_all); \
      attr((const void*)k_chain_all<true, 128, false, ML>, g->smem_chain_all); attr((const void*)k_chain_all<false, 128, false, ML>, g->smem_chain_all); \
      attr((const void*)k_chain_all<true, 64, true, ML>, g->smem_chain_all); attr((const void*)k_chain_all<false, 64, true, ML>, g->smem_chain_all); \
      attr((const void*)k_chain_all<true, 128, true, ML>, g->smem_chain_all); attr((const void*)k_chain_all<false, 128, true, ML>, g->smem_chain_all);
      if (g->multilabel) { CHAIN_ATTR(true) } else { CHAIN_ATTR(false) }
#undef CHAIN_ATTR
    }
    if (g->bwd_ws) {
      std::vector<int4> tl;
      { const char* se = getenv("MFAS_BWD_SMALL"); if (g->Hmax <= 32 && g->tchead && !(se && !atoi(se))) g->bwd_small = g->Hmax <= 16 ? 16 : 32; }
      g->n_bwd_layer_tiles = build_bwd_tiles(g->lay.data(), n_cand, g->tchead, tl, g->bwd_small ? g->bwd_small : TC_BWD_HT);
      if (g->bwd_small) {
#define BS_ATTR(HN_) attr((const void*)k_tc_bwd_small<HN_, false, false>, BwdSmall<HN_>::SMEM); attr((const void*)k_tc_bwd_small<HN_, true, false>, BwdSmall<HN_>::SMEM); \
                     attr((const void*)k_tc_bwd_small<HN_, false, true>, BwdSmall<HN_>::SMEM); attr((const void*)k_tc_bwd_small<HN_, true, true>, BwdSmall<HN_>::SMEM);
        BS_ATTR(16) BS_ATTR(32)
#undef BS_ATTR
      }
      g->n_bwd_tiles = (int)tl.size();
      if (e == cudaSuccess) e = pool_alloc_t(device, sizeof(BwdTile) * tl.size(), &g->bwd_tiles, &g->tiles_bytes);
      g->bwd_tl = std::move(tl);
      attr((const void*)k_tc_bwd_ws<false, false>, TC_WS_SMEM);
      attr((const void*)k_tc_bwd_ws<true, false>, TC_WS_SMEM);
      if (g->any_alphas) {
        attr((const void*)k_tc_bwd_ws<false, true>, TC_WS_SMEM);
        attr((const void*)k_tc_bwd_ws<true, true>, TC_WS_SMEM);
        std::vector<int2> rng((size_t)n_cand * MFAS_MAX_LAYERS, make_int2(0, 0));
        for (int i = 0; i < g->n_bwd_layer_tiles; ++i) {         // tiles are listed candidate by candidate, layer by layer, feature columns first
          const int4 t = g->bwd_tl[i];
          int2& r = rng[(size_t)t.x * MFAS_MAX_LAYERS + t.y];
          if (t.z < g->lay[t.x].d_ske[t.y] + g->lay[t.x].d_rgb[t.y]) { if (r.y == 0) r.x = i; ++r.y; }
        }
        if (e == cudaSuccess) e = pool_alloc_t(device, sizeof(float) * TC_DSP_PER_TILE * g->bwd_tl.size(), &g->dsp_tc, &g->dsp_bytes);
        if (e == cudaSuccess) e = cudaMemset(g->dsp_tc, 0, sizeof(float) * TC_DSP_PER_TILE * g->bwd_tl.size());
        if (e == cudaSuccess) e = pool_alloc_t(device, sizeof(int2) * rng.size(), &g->alpha_rng, &g->rng_bytes);
        if (e == cudaSuccess) e = cudaMemcpy(g->alpha_rng, rng.data(), sizeof(int2) * rng.size(), cudaMemcpyHostToDevice);
      }
    }
    if (g->any_alphas && (!g->bwd_ws || !g->chain)) {
      int code = fail(MFAS_ERR_UNSUPPORTED, "alpha gates on the tensor-core engine need the persistent backward and the tensor-core chain kernels");
      mfas_group_destroy(g);
      return code;
    }
    if (getenv("MFAS_CHAIN_TIMELINE") && e == cudaSuccess) {
      e = pool_alloc_t(device, sizeof(long long) * 16 * n_cand, &g->timeline, &g->tl_bytes);
      if (e == cudaSuccess) e = cudaMemset(g->timeline, 0, sizeof(long long) * 16 * n_cand);
    }
    { const char* fe = getenv("MFAS_FWD"); if (fe && !strcmp(fe, "cta")) g->fwd_ws = 0; }
    if (g->fwd_ws) {
      int n = 0;
      for (int c = 0; c < n_cand; ++c)
        for (int l = 0; l < g->lay[c].L; ++l)
          n += tc_fwd_items_g(g->lay[c].d_ske[l], g->lay[c].d_rgb[l], (g->lay[c].flags & MFAS_FLAG_ALPHAS) != 0, g->kb_item) * ((g->lay[c].H + 127) / 128);
      g->n_fwd_items = n;
      if (e == cudaSuccess) e = pool_alloc_t(device, sizeof(FwdItem) * n, &g->fwd_items, &g->items_bytes);
      attr((const void*)k_tc_fwd_ws<64, 0>, FwdWs<64, 0>::SMEM);
      attr((const void*)k_tc_fwd_ws<128, 0>, FwdWs<128, 0>::SMEM);
      attr((const void*)k_tc_fwd_ws<64, 1>, FwdWs<64, 1>::SMEM);
      attr((const void*)k_tc_fwd_ws<128, 1>, FwdWs<128, 1>::SMEM);
      { const char* xe = getenv("MFAS_FWD_XR"); if (xe) g->fwd_xr = atoi(xe) ? 1 : 0; }
      { const char* se = getenv("MFAS_FWD_SMALL"); if (se) g->fwd_small = atoi(se) ? 1 : 0; }
      if (g->Hmax > 32) g->fwd_small = 0;
      if (g->fwd_small) {
        attr((const void*)k_tc_fwd_small<64, 16>, FwdSmall<64, 16>::SMEM); attr((const void*)k_tc_fwd_small<64, 32>, FwdSmall<64, 32>::SMEM);
        attr((const void*)k_tc_fwd_small<128, 16>, FwdSmall<128, 16>::SMEM); attr((const void*)k_tc_fwd_small<128, 32>, FwdSmall<128, 32>::SMEM);
      }
      { const char* te = getenv("MFAS_FWD_TMA"); if (te) g->fwd_tma = atoi(te) < 0 ? 0 : (atoi(te) > 2 ? 2 : atoi(te)); }
      if (!tmap_encoder()) g->fwd_tma = 0;
      if (g->fwd_tma && e == cudaSuccess) e = pool_alloc_t(device, sizeof(CUtensorMap) * n, &g->fwd_wmaps, &g->wmaps_bytes);
    }
    // wide eval: needs the persistent forward (its item list carries the partial-sum offsets) and the tensor-core head
    g->ev128 = g->fwd_ws && g->tchead && g->npad == 64;
    { const char* ee = getenv("MFAS_EVAL128"); if (ee && !atoi(ee)) g->ev128 = false; }
    g->part_stride_ev = (long long)g->items_fwd * Hp * 128;
    if (e == cudaSuccess)
      e = pool_alloc(device, sizeof(float) * (g->ev128 ? g->part_stride_ev : g->part_stride) * n_cand, (void**)&g->part, &g->part_bytes);
    if (g->ev128) {
      if (e == cudaSuccess) e = pool_alloc_t(device, sizeof(FwdItem) * g->n_fwd_items, &g->fwd_items_ev, &g->items_ev_bytes);
    }
    { const char* de = getenv("MFAS_TC_DEBUG"); if (de) g->dbg = atoi(de); }
    { const char* le = getenv("MFAS_L2_HINTS"); if (le) g->l2_hints = atoi(le) & 15; }
    if (e != cudaSuccess) {
      int code = fail(MFAS_ERR_CUDA, "tc engine setup: %s", cudaGetErrorString(e));
      mfas_group_destroy(g);
      return code;
    }
  }
  *out = g;
  return MFAS_OK;
}

extern "C" int mfas_group_engine(mfas_group_t g, int32_t* out) {
  if (!g || !out) return fail(MFAS_ERR_INVALID, "null argument");
  *out = g->engine;
  return MFAS_OK;
}

// Synchronises the device and reports sticky kernel-side failures (a tcgen05 barrier that timed out).
extern "C" int mfas_group_status(mfas_group_t g) {
  if (!g) return fail(MFAS_ERR_INVALID, "null group");
  DeviceGuard dg(g->device);
  CUDA_TRY(cudaDeviceSynchronize());
  if (g->tc_err) {
    int flag = 0;
    CUDA_TRY(cudaMemcpy(&flag, g->tc_err, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) return fail(MFAS_ERR_CUDA, "tensor-core pipeline barrier timed out in kernel %s", flag == 1 ? "k_tc_fwd_all" : flag == 2 ? "k_tc_bwd_all" : flag == 3 ? "k_chain_fwd" : flag == 4 ? "k_chain_bwd" : flag == 5 ? "k_tc_bwd_ws" : flag == 6 ? "k_tc_fwd_ws" : "k_chain_all");
  }
  return MFAS_OK;
}

extern "C" int mfas_group_bind(mfas_group_t g, int32_t cand, const mfas_arenas* a) {
  if (!g || !a) return fail(MFAS_ERR_INVALID, "null argument");
  if (cand < 0 || cand >= g->n_cand) return fail(MFAS_ERR_INVALID, "candidate %d of %d", cand, g->n_cand);
  if (!a->params || !a->adam_m || !a->adam_v || !a->bufs || !a->nbt)
    return fail(MFAS_ERR_INVALID, "params/adam_m/adam_v/bufs/nbt must be non-null");
  const uintptr_t bits = (uintptr_t)a->params | (uintptr_t)a->adam_m | (uintptr_t)a->adam_v | (uintptr_t)a->bufs |
                         (uintptr_t)a->grad;
  if (bits & 15) return fail(MFAS_ERR_INVALID, "arenas must be 16-byte aligned");
  DCand& d = g->hc[cand];
  d.p = a->params; d.m = a->adam_m; d.v = a->adam_v; d.grad = a->grad; d.bufs = a->bufs;
  d.nbt = (long long*)a->nbt;
  g->bound[cand] = 1;
  g->dirty = true;
  return MFAS_OK;
}

static int sync_descriptors(mfas_group* g, cudaStream_t st);

extern "C" int mfas_group_init_params(mfas_group_t g, uint64_t seed, void* stream) {
  if (!g) return fail(MFAS_ERR_INVALID, "null group");
  DeviceGuard dg(g->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = sync_descriptors(g, st);
  if (rc) return rc;
  long long nmax = 0;
  for (int c = 0; c < g->n_cand; ++c) nmax = g->lay[c].n_params > nmax ? g->lay[c].n_params : nmax;
  int tiles = (int)((nmax + kThreads * 8 - 1) / (kThreads * 8));
  tiles = tiles < 1 ? 1 : (tiles > 64 ? 64 : tiles);
  k_init_params<<<dim3(tiles, g->n_cand), kThreads, 0, st>>>(g->dc, (uint32_t)seed, (uint32_t)(seed >> 32));
  ++g->launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MFAS_ERR_CUDA, "k_init_params launch failed: %s", cudaGetErrorString(e));
  return MFAS_OK;
}

extern "C" int mfas_group_num_launches(mfas_group_t g, int64_t* out) {
  if (!g || !out) return fail(MFAS_ERR_INVALID, "null argument");
  *out = g->launches;
  return MFAS_OK;
}

extern "C" int mfas_group_chain_timeline(mfas_group_t g, int64_t* out, int32_t n_cand) {
  if (!g || !out) return fail(MFAS_ERR_INVALID, "null argument");
  if (!g->timeline) return fail(MFAS_ERR_INVALID, "create the group with MFAS_CHAIN_TIMELINE=1");
  DeviceGuard dg(g->device);
  CUDA_TRY(cudaDeviceSynchronize());
  CUDA_TRY(cudaMemcpy(out, g->timeline, sizeof(long long) * 16 * (n_cand < g->n_cand ? n_cand : g->n_cand), cudaMemcpyDeviceToHost));
  return MFAS_OK;
}

extern "C" int mfas_group_set_profiling(mfas_group_t g, int32_t on) {
  if (!g) return fail(MFAS_ERR_INVALID, "null group");
  DeviceGuard dg(g->device);
  if (on)
    for (auto& ev : g->prof_ev)
      if (!ev) CUDA_TRY(cudaEventCreate(&ev));
  g->prof = on != 0;
  g->prof_valid = false;
  return MFAS_OK;
}

extern "C" int mfas_group_last_step_ms(mfas_group_t g, float* ms3) {
  if (!g || !ms3) return fail(MFAS_ERR_INVALID, "null argument");
  if (!g->prof || !g->prof_valid) return fail(MFAS_ERR_INVALID, "no profiled train step (tensor-core engine with the fused chain, after mfas_group_set_profiling(g, 1))");
  DeviceGuard dg(g->device);
  CUDA_TRY(cudaEventSynchronize(g->prof_ev[3]));
  for (int i = 0; i < 3; ++i) CUDA_TRY(cudaEventElapsedTime(&ms3[i], g->prof_ev[i], g->prof_ev[i + 1]));
  return MFAS_OK;
}

static int sync_descriptors(mfas_group* g, cudaStream_t st) {
  for (int c = 0; c < g->n_cand; ++c)
    if (!g->bound[c]) return fail(MFAS_ERR_UNBOUND, "candidate %d has no arenas bound", c);
  if (g->dirty) {
    // pageable source: the copy is staged before the call returns, so hc may change afterwards
    CUDA_TRY(cudaMemcpyAsync(g->dc, g->hc.data(), sizeof(DCand) * g->n_cand, cudaMemcpyHostToDevice, st));
    if (g->bwd_tiles) {
      std::vector<BwdTile> recs(g->bwd_tl.size());
      for (size_t i = 0; i < recs.size(); ++i) {
        const int4 t = g->bwd_tl[i];
        const DCand& d = g->hc[t.x];
        BwdTile& r = recs[i];
        r.moff = (long long)(d.m - d.p); r.voff = (long long)(d.v - d.p);
        r.goff = d.grad ? (long long)(d.grad - d.p) : 0;
        const int ht = g->bwd_small ? g->bwd_small : TC_BWD_HT;
        r.xsrc = nullptr; r.xkind = 0; r.xtap = 0; r.xcol = 0; r.xld = d.H; r.pad2 = r.pad3 = 0;
        if (t.y >= d.L) {                               // classifier tile (tensor-core head): W_c [C][H], columns t.z.., rows t.w..
          r.W = d.p + d.oWc + (long long)t.w * d.H + t.z;
          r.K = d.H; r.kw = d.H - t.z < TC_BWD_KT ? d.H - t.z : TC_BWD_KT;
          r.rows = g->bwd_small ? (d.C - t.w < ht ? d.C - t.w : ht) : d.C;
          r.xsrc = d.hid + (long long)(d.L - 1) * g->bmax * d.H + t.z;
          r.dz = d.dlog + t.w; r.dzld = TC_DLOG_LD; r.hw = TC_DLOG_LD - t.w < ht ? TC_DLOG_LD - t.w : ht;
        } else {
          const DLayer& ly = d.layer[t.y];
          r.W = d.p + ly.oW + (long long)t.w * ly.K + t.z;
          const int seg_end = tc_bwd_seg_end(ly.d_ske, ly.d_rgb, ly.K, t.z);
          r.K = ly.K; r.kw = seg_end - t.z < TC_BWD_KT ? seg_end - t.z : TC_BWD_KT;
          r.rows = d.H - t.w < ht ? d.H - t.w : ht;
          if (t.z < ly.d_ske) { r.xkind = 1; r.xtap = ly.ske_tap; r.xcol = t.z; }
          else if (t.z < ly.d_ske + ly.d_rgb) { r.xkind = 2; r.xtap = ly.rgb_tap; r.xcol = t.z - ly.d_ske; }
          else r.xsrc = d.hid + (long long)(t.y - 1) * g->bmax * d.H + (t.z - ly.d_ske - ly.d_rgb);
          r.dz = d.dzs + (long long)t.y * g->bmax * d.H + t.w; r.dzld = d.H; r.hw = r.rows;
        }
        r.cand = t.x; r.layer = t.y; r.kc0 = t.z; r.h0 = t.w; r.pad1 = 0;
        r.alpha = nullptr; r.gate = 0; r.slot = (int)i;
        if ((d.flags & MFAS_FLAG_ALPHAS) && t.y < d.L) {
          const DLayer& ly = d.layer[t.y];
          r.alpha = d.p + ly.oalpha;
          r.gate = t.z < ly.d_ske ? 1 : (t.z < ly.d_ske + ly.d_rgb ? 2 : 0);
        }
      }
      // pageable source: staged before the call returns
      CUDA_TRY(cudaMemcpyAsync(g->bwd_tiles, recs.data(), sizeof(BwdTile) * recs.size(), cudaMemcpyHostToDevice, st));
    }
    if (g->fwd_items) {
      // one item per (candidate, layer, 128-row tile, k-range); biggest first, so that the static round-robin
      // over the persistent CTAs ends level
      std::vector<FwdItem> its;
      const int Hp = ((g->Hmax + 127) / 128) * 128;
      for (int c = 0; c < g->n_cand; ++c) {
        const DCand& d = g->hc[c];
        int item0 = 0;
        for (int l = 0; l < d.L; ++l) {
          const DLayer& ly = d.layer[l];
          const bool gated = (d.flags & MFAS_FLAG_ALPHAS) != 0;          // items cut at the modality boundary (tc_fwd_items_g)
          const int ns = tc_fwd_items_g(ly.d_ske, ly.d_rgb, gated, g->kb_item);
          for (int sp = 0; sp < ns; ++sp)
            for (int m0 = 0; m0 < d.H; m0 += 128) {
              FwdItem it;
              tc_fwd_range_g(ly.d_ske, ly.d_rgb, sp, gated, it.kb0, it.kb1, g->kb_item);
              it.W = d.p + ly.oW + (long long)m0 * ly.K;
              it.part_off = (long long)c * g->part_stride + (long long)(item0 + sp) * Hp * g->npad + 4LL * m0;
              it.Hp = Hp; it.pad0 = it.pad1 = it.pad2 = 0;
              it.K = ly.K; it.fs_kb = ly.d_ske >> 5;
              it.ske_tap = ly.ske_tap; it.rgb_tap = ly.rgb_tap; it.cand = c;
              it.rows_valid = d.H - m0 < 128 ? d.H - m0 : 128;
              its.push_back(it);
            }
          item0 += ns;
        }
      }
      std::stable_sort(its.begin(), its.end(), [](const FwdItem& a, const FwdItem& b) { return a.kb1 - a.kb0 > b.kb1 - b.kb0; });
      if ((int)its.size() != g->n_fwd_items) return fail(MFAS_ERR_INVALID, "internal: forward item count changed");
      CUDA_TRY(cudaMemcpyAsync(g->fwd_items, its.data(), sizeof(FwdItem) * its.size(), cudaMemcpyHostToDevice, st));
      if (g->fwd_tma) {                                 // the W tile of item i: rows [m0, m0 + rows_valid) of its layer, all K columns
        std::vector<CUtensorMap> maps(its.size());
        for (size_t i = 0; i < its.size(); ++i)
          if (!encode_tile_map(&maps[i], its[i].W, its[i].K, its[i].rows_valid, its[i].K, 128))
            return fail(MFAS_ERR_CUDA, "cuTensorMapEncodeTiled failed for forward item %zu (K=%d, rows=%d)", i, its[i].K, its[i].rows_valid);
        CUDA_TRY(cudaMemcpyAsync(g->fwd_wmaps, maps.data(), sizeof(CUtensorMap) * maps.size(), cudaMemcpyHostToDevice, st));
      }
      if (g->fwd_items_ev) {                            // same items, partial sums laid out for 128 batch rows
        for (FwdItem& it : its) {
          const long long rel = it.part_off - (long long)it.cand * g->part_stride;     // (item index) * Hp * npad + 4 m0
          const long long idx = rel / ((long long)Hp * g->npad), m4 = rel % ((long long)Hp * g->npad);
          it.part_off = (long long)it.cand * g->part_stride_ev + idx * Hp * 128 + m4;
        }
        CUDA_TRY(cudaMemcpyAsync(g->fwd_items_ev, its.data(), sizeof(FwdItem) * its.size(), cudaMemcpyHostToDevice, st));
      }
    }
    g->dirty = false;
  }
  return MFAS_OK;
}

static int to_dcache(const mfas_group* g, const mfas_cache_desc* c, DCache* out) {
  if (!c) return fail(MFAS_ERR_INVALID, "null cache");
  if (c->n_rows <= 0) return fail(MFAS_ERR_INVALID, "cache needs n_rows>0");
  if (!g->multilabel && !c->labels) return fail(MFAS_ERR_INVALID, "cache needs labels");
  if (g->multilabel && (!c->targets || !c->pos_weight))
    return fail(MFAS_ERR_INVALID, "multi-label candidates need cache.targets [n_rows, C] and cache.pos_weight [C]");
  out->n_rows = c->n_rows;
  for (int t = 0; t < MFAS_NUM_TAPS; ++t) {
    if (!c->ske[t] || !c->rgb[t]) return fail(MFAS_ERR_INVALID, "cache tap %d is null", t);
    if (((uintptr_t)c->ske[t] | (uintptr_t)c->rgb[t]) & 15 || c->ske_ld[t] % 4 || c->rgb_ld[t] % 4)
      return fail(MFAS_ERR_INVALID, "cache taps must be 16-byte aligned with ld %% 4 == 0");
    out->ske[t] = c->ske[t]; out->rgb[t] = c->rgb[t];
    out->ske_ld[t] = c->ske_ld[t]; out->rgb_ld[t] = c->rgb_ld[t];
  }
  for (int k = 0; k < g->n_cand; ++k)
    for (int l = 0; l < g->lay[k].L; ++l)
      if (c->d_ske[g->lay[k].conf[l][0]] != g->lay[k].d_ske[l] || c->d_rgb[g->lay[k].conf[l][1]] != g->lay[k].d_rgb[l])
        return fail(MFAS_ERR_INVALID, "cache tap widths differ from the layout of candidate %d step %d", k, l);
  out->labels = (const long long*)c->labels;
  out->logit_rgb = c->logit_rgb; out->logit_ske = c->logit_ske;
  out->targets = c->targets; out->pos_weight = c->pos_weight;
  for (int k = 0; k < g->n_cand; ++k)
    if ((g->lay[k].flags & MFAS_FLAG_MULTITASK) && (!c->logit_rgb || !c->logit_ske))
      return fail(MFAS_ERR_INVALID, "multitask candidates need the cached backbone logits (cache.logit_rgb / logit_ske)");
  return MFAS_OK;
}

// Launch with (or, MFAS_PDL=0, without) programmatic stream serialization: see griddep_wait / griddep_launch in common.cuh.
// MFAS_PDL: bit 0 = the forward stream, bit 1 = the chain, bit 2 = the backward stream; 0 = plain launches.  Default 3:
// measured on B200 (profiles/r02i_pdl_ablation.txt) the forward stream and the chain as dependent launches take 3-5 % off
// the latency-bound steps (32 candidates: 98.7 -> 94 us; eval steps 64 -> 60 us) and nothing off the HBM-bound ones, while
// the backward stream as a dependent launch LOSES 12 % at 148 candidates (941 -> 1062 us): its CTAs then land on SMs in
// the order the chain's CTAs leave them instead of in blockIdx order, and neighbouring tiles of the list (which share
// their x columns through L2) end up on different dies.
static int pdl_mask() { static const int m = [] { const char* e = getenv("MFAS_PDL"); return e ? atoi(e) : 3; }(); return m; }
static thread_local int g_launch_cluster = 1;       // thread-block cluster size of the next launch_k (set and reset by the chain launch)
template <class... KArgs, class... Args>
static void launch_k(int which, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (pdl_mask() & which) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (g_launch_cluster > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = g_launch_cluster; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = at; cfg.numAttrs = n;
  cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);      // errors surface through cudaGetLastError (LAUNCH_CHECK)
}

#define LAUNCH_CHECK(g)                                                                       \
  do {                                                                                        \
    ++(g)->launches;                                                                          \
    cudaError_t e__ = cudaGetLastError();                                                     \
    if (e__ != cudaSuccess) return fail(MFAS_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

static int launch_bwd_stream(mfas_group* g, const DCache& cache, const BatchRef& batch, float step_size, float bc2_sqrt,
                             bool keep, bool head_tiles, cudaStream_t st);

// Tensor maps of the 8 feature taps of a cache (tile::gather4 source of the forward stream), cached per (pointers, strides, rows)
static const TapMaps* tap_maps_of(mfas_group* g, const DCache& c) {
  mfas_group::TapKey k;
  memset(&k, 0, sizeof(k));
  k.n_rows = c.n_rows;
  for (int t = 0; t < MFAS_NUM_TAPS; ++t) { k.ske[t] = c.ske[t]; k.rgb[t] = c.rgb[t]; k.ske_ld[t] = c.ske_ld[t]; k.rgb_ld[t] = c.rgb_ld[t]; }
  ++g->tap_clock;
  int slot = 0;
  for (int i = 0; i < 4; ++i) {
    if (g->tap_used[i] && !memcmp(&g->tap_key[i], &k, sizeof(k))) { g->tap_used[i] = g->tap_clock; return &g->tap_maps[i]; }
    if (g->tap_used[i] < g->tap_used[slot]) slot = i;
  }
  // widths: the widest layer that reads a tap decides nothing here -- a tap's tensor spans [n_rows][ld - offset] at most; its
  // true width is not in the descriptor, and columns past it belong to the next tap of the same matrix, which no item addresses
  for (int t = 0; t < MFAS_NUM_TAPS; ++t) {
    long long ws = 0, wr = 0;
    for (int cnd = 0; cnd < g->n_cand; ++cnd)
      for (int l = 0; l < g->lay[cnd].L; ++l) {
        if (g->lay[cnd].conf[l][0] == t) ws = g->lay[cnd].d_ske[l];
        if (g->lay[cnd].conf[l][1] == t) wr = g->lay[cnd].d_rgb[l];
      }
    if (ws == 0) ws = 32;                               // a tap no candidate reads: any valid extent
    if (wr == 0) wr = 32;
    if (!encode_tile_map(&g->tap_maps[slot].ske[t], c.ske[t], ws, c.n_rows, c.ske_ld[t], 1)) return nullptr;
    if (!encode_tile_map(&g->tap_maps[slot].rgb[t], c.rgb[t], wr, c.n_rows, c.rgb_ld[t], 1)) return nullptr;
  }
  g->tap_key[slot] = k;
  g->tap_used[slot] = g->tap_clock;
  return &g->tap_maps[slot];
}

// tc engine: one launch covers the feature columns of every layer, small per-layer kernels carry the chain
static int launch_step_tc(mfas_group* g, const DCache& cache, const BatchRef& batch, bool train, bool bn_train,
                          float step_size, float bc2_sqrt, uint32_t step, const HeadOut& ho, cudaStream_t st) {
  const TcErr terr{g->tc_err, g->timeline, g->l2_hints};
  const dim3 gf(g->items_fwd, (g->Hmax + 127) / 128, g->n_cand), gl((g->Hmax + TC_CB - 1) / TC_CB, g->n_cand);
  const bool prof = g->prof && train && g->chain == 2;
  if (prof) { g->prof_valid = false; cudaEventRecord(g->prof_ev[0], st); }
  const bool wide = g->ev128 && !train && !bn_train && batch.n_rows > 64;      // a 128-row dev / test step
  if (g->fwd_ws) {
    const int grid = g->n_fwd_items < g->n_sms ? g->n_fwd_items : g->n_sms;
    const FwdItem* items = wide ? g->fwd_items_ev : g->fwd_items;
    const TapMaps* tm = g->fwd_tma ? tap_maps_of(g, cache) : &g->tap_maps[0];
    if (!tm) return fail(MFAS_ERR_CUDA, "cuTensorMapEncodeTiled failed for the feature taps (pointer %p, ld %lld)", (const void*)cache.ske[0], cache.ske_ld[0]);
#define FW(N, X) launch_k(1, k_tc_fwd_ws<N, X>, dim3(grid), dim3(FwdWs<N, X>::THREADS), FwdWs<N, X>::SMEM, st, items, g->n_fwd_items, cache, batch, g->part, terr, (const CUtensorMap*)g->fwd_wmaps, *tm, g->fwd_tma)
#define FS(N, HN_) launch_k(1, k_tc_fwd_small<N, HN_>, dim3(grid), dim3(FwdSmall<N, HN_>::THREADS), FwdSmall<N, HN_>::SMEM, st, items, g->n_fwd_items, cache, batch, g->part, terr)
    if (g->fwd_small) {
      if (g->npad == 64 && !wide) { if (g->Hmax <= 16) FS(64, 16); else FS(64, 32); }
      else { if (g->Hmax <= 16) FS(128, 16); else FS(128, 32); }
    }
    else if (g->npad == 64 && !wide) { if (g->fwd_xr) FW(64, 1); else FW(64, 0); }
    else { if (g->fwd_xr) FW(128, 1); else FW(128, 0); }
#undef FS
#undef FW
  } else if (g->npad == 64) k_tc_fwd_all<64><<<gf, TC_THREADS, g->smem_tc_fwd, st>>>(g->dc, cache, batch, g->part, g->part_stride, terr);
  else k_tc_fwd_all<128><<<gf, TC_THREADS, g->smem_tc_fwd, st>>>(g->dc, cache, batch, g->part, g->part_stride, terr);
  LAUNCH_CHECK(g);
  const dim3 gc((g->Hmax + 127) / 128, g->n_cand);
  bool keep = false;                                  // the grad arena is a test facility: all candidates or none
  for (int c = 0; c < g->n_cand; ++c) keep = keep || g->hc[c].grad != nullptr;
  if (prof) cudaEventRecord(g->prof_ev[1], st);
  if (g->chain == 2 && train == bn_train) {           // the whole serial chain in one launch
    g_launch_cluster = g->chain_cl;
#define CA_(T, N, TH, ML, BM, PS) launch_k(2, k_chain_all<T, N, TH, ML>, dim3(g->n_cand * g->chain_cl), dim3(ChainCfg<N>::THREADS), g->smem_chain_all, st, (const DCand*)g->dc, cache, batch, (int)(BM), (const float*)g->part, (long long)(PS), g->hs_ld, g->lg_ld, g->adam, step_size, bc2_sqrt, g->drop_seed, g->drop_p, step, ho, terr, g->chain_cl)
#define CA(T, N, TH) do { if (g->multilabel) CA_(T, N, TH, true, g->bmax, g->part_stride); else CA_(T, N, TH, false, g->bmax, g->part_stride); } while (0)
#define CS_(T, N, HN_, ML, BM, PS) launch_k(2, k_chain_small<T, N, HN_, ML>, dim3(g->n_cand), dim3(ChainSmall<N, HN_>::THREADS), ChainSmall<N, HN_>::smem(g->Lmax, T, g->chain_small_items), st, (const DCand*)g->dc, cache, batch, (int)(BM), (const float*)g->part, (long long)(PS), g->adam, step_size, bc2_sqrt, g->drop_seed, g->drop_p, step, ho, terr, g->chain_small_items)
#define CS(T, N, BM, PS) do { if (g->chain_small == 16) { if (g->multilabel) CS_(T, N, 16, true, BM, PS); else CS_(T, N, 16, false, BM, PS); } else { if (g->multilabel) CS_(T, N, 32, true, BM, PS); else CS_(T, N, 32, false, BM, PS); } } while (0)
    if (g->chain_small) {
      g_launch_cluster = 1;
      if (wide) CS(false, 128, 128, g->part_stride_ev);
      else if (g->npad == 64) { if (train) CS(true, 64, g->bmax, g->part_stride); else CS(false, 64, g->bmax, g->part_stride); }
      else { if (train) CS(true, 128, g->bmax, g->part_stride); else CS(false, 128, g->bmax, g->part_stride); }
    }
    else if (wide) { if (g->multilabel) CA_(false, 128, true, true, 128, g->part_stride_ev); else CA_(false, 128, true, false, 128, g->part_stride_ev); }
    else if (g->npad == 64) { if (g->tchead) { if (train) CA(true, 64, true); else CA(false, 64, true); } else { if (train) CA(true, 64, false); else CA(false, 64, false); } }
    else { if (g->tchead) { if (train) CA(true, 128, true); else CA(false, 128, true); } else { if (train) CA(true, 128, false); else CA(false, 128, false); } }
#undef CA_
#undef CS
#undef CS_
    g_launch_cluster = 1;
#undef CA
    LAUNCH_CHECK(g);
    if (!train) return MFAS_OK;
    if (prof) cudaEventRecord(g->prof_ev[2], st);
    const int rc = launch_bwd_stream(g, cache, batch, step_size, bc2_sqrt, keep, g->tchead, st);
    if (prof) { cudaEventRecord(g->prof_ev[3], st); g->prof_valid = true; }
    return rc;
  }
  for (int l = 0; l < g->Lmax; ++l) {
#define FL(T, N) k_fwd_layer<T, N><<<gl, TC_CHAIN_THREADS, g->smem_fl, st>>>(g->dc, l, batch.n_rows, g->bmax, g->part, g->part_stride, g->drop_seed, g->drop_p, step)
#define CF(T, N) k_chain_fwd<T, N><<<gc, ChainCfg<N>::THREADS, g->smem_chain, st>>>(g->dc, l, batch.n_rows, g->bmax, g->part, g->part_stride, g->drop_seed, g->drop_p, step, terr)
    if (g->chain) {
      if (g->npad == 64) { if (bn_train) CF(true, 64); else CF(false, 64); }
      else { if (bn_train) CF(true, 128); else CF(false, 128); }
    } else {
      if (g->npad == 64) { if (bn_train) FL(true, 64); else FL(false, 64); }
      else { if (bn_train) FL(true, 128); else FL(false, 128); }
    }
#undef FL
#undef CF
    LAUNCH_CHECK(g);
  }
  if (g->multilabel) {
    if (train)
      k_head<true, true><<<g->n_cand, kHeadThreads, g->smem_head, st>>>(g->dc, cache, batch, g->bmax, g->hs_ld, g->lg_ld, g->adam, step_size, bc2_sqrt, ho);
    else
      k_head<false, true><<<g->n_cand, kHeadThreads, g->smem_head, st>>>(g->dc, cache, batch, g->bmax, g->hs_ld, g->lg_ld, g->adam, step_size, bc2_sqrt, ho);
  } else if (train)
    k_head<true><<<g->n_cand, kHeadThreads, g->smem_head, st>>>(g->dc, cache, batch, g->bmax, g->hs_ld, g->lg_ld, g->adam, step_size, bc2_sqrt, ho);
  else
    k_head<false><<<g->n_cand, kHeadThreads, g->smem_head, st>>>(g->dc, cache, batch, g->bmax, g->hs_ld, g->lg_ld, g->adam, step_size, bc2_sqrt, ho);
  LAUNCH_CHECK(g);
  if (!train) return MFAS_OK;
  for (int l = g->Lmax - 1; l >= 0; --l) {
    if (g->chain) {
      if (g->npad == 64) k_chain_bwd<64><<<gc, ChainCfg<64>::THREADS, g->smem_chain, st>>>(g->dc, l, batch.n_rows, g->bmax, g->adam, step_size, bc2_sqrt, g->drop_seed, g->drop_p, step, terr);
      else k_chain_bwd<128><<<gc, ChainCfg<128>::THREADS, g->smem_chain, st>>>(g->dc, l, batch.n_rows, g->bmax, g->adam, step_size, bc2_sqrt, g->drop_seed, g->drop_p, step, terr);
    } else {
      k_dzx<<<gl, kThreads, g->smem_dzx, st>>>(g->dc, l, batch.n_rows, g->bmax, g->adam, step_size, bc2_sqrt, g->drop_seed,
                                               g->drop_p, step);
    }
    LAUNCH_CHECK(g);
  }
  return launch_bwd_stream(g, cache, batch, step_size, bc2_sqrt, keep, false, st);
}

// the weight-gradient + Adam streaming kernel of a train step (all layers, all candidates)
static int launch_bwd_stream(mfas_group* g, const DCache& cache, const BatchRef& batch, float step_size, float bc2_sqrt,
                             bool keep, bool head_tiles, cudaStream_t st) {
  const TcErr terr{g->tc_err, g->timeline, g->l2_hints};
  const dim3 gb(g->items_bwd, (g->Hmax + TC_BWD_HT - 1) / TC_BWD_HT, g->n_cand);
  if (g->bwd_ws) {
    const int nt = head_tiles ? g->n_bwd_tiles : g->n_bwd_layer_tiles;
    const int grid = nt < g->n_sms ? nt : g->n_sms;
#define BSM(HN_, KG, AL) launch_k(4, k_tc_bwd_small<HN_, KG, AL>, dim3(grid), dim3(BwdSmall<HN_>::THREADS), BwdSmall<HN_>::SMEM, st, (const DCand*)g->dc, cache, batch, g->bmax, g->adam, step_size, bc2_sqrt, (const BwdTile*)g->bwd_tiles, nt, terr, g->dsp_tc)
#define BSM2(KG, AL) do { if (g->bwd_small == 16) BSM(16, KG, AL); else BSM(32, KG, AL); } while (0)
    if (g->bwd_small) {
      if (g->any_alphas) { if (keep) BSM2(true, true); else BSM2(false, true); }
      else { if (keep) BSM2(true, false); else BSM2(false, false); }
    } else
#define BWS(KG, AL) launch_k(4, k_tc_bwd_ws<KG, AL>, dim3(grid), dim3(TC_WS_THREADS), TC_WS_SMEM, st, (const DCand*)g->dc, cache, batch, g->bmax, g->adam, step_size, bc2_sqrt, (const BwdTile*)g->bwd_tiles, nt, terr, g->dsp_tc)
    if (g->any_alphas) { if (keep) BWS(true, true); else BWS(false, true); }
    else { if (keep) BWS(true, false); else BWS(false, false); }
#undef BWS
#undef BSM2
#undef BSM
    LAUNCH_CHECK(g);
    if (g->any_alphas) {                                // the gates' own gradient + Adam step, after every use of the old alphas
      const int per = 256 / MFAS_MAX_LAYERS;
      k_alpha_step_tc<<<(g->n_cand + per - 1) / per, 256, 0, st>>>(g->dc, g->n_cand, g->alpha_rng, g->dsp_tc, g->adam, step_size, bc2_sqrt);
      LAUNCH_CHECK(g);
    }
    return MFAS_OK;
  }
#define BW(BPV, KG) k_tc_bwd_all<BPV, KG><<<gb, TC_BWD_THREADS, g->smem_tc_bwd, st>>>(g->dc, cache, batch, g->bmax, g->adam, step_size, bc2_sqrt, terr, g->dbg)
  if (g->npad == 64) { if (keep) BW(64, true); else BW(64, false); }
  else { if (keep) BW(128, true); else BW(128, false); }
#undef BW
  LAUNCH_CHECK(g);
  return MFAS_OK;
}

// one forward (+ optional backward/Adam) of every candidate over one batch
static int launch_step(mfas_group* g, const DCache& cache, const BatchRef& batch, bool train, bool bn_train,
                       float step_size, float bc2_sqrt, uint32_t step, const HeadOut& ho, cudaStream_t st) {
  if (g->engine == 1) return launch_step_tc(g, cache, batch, train, bn_train, step_size, bc2_sqrt, step, ho, st);
  const dim3 fgrid((g->Hmax + FWD_HT - 1) / FWD_HT, g->n_cand);
  for (int l = 0; l < g->Lmax; ++l) {
    if (bn_train)
      k_fusion_fwd<true><<<fgrid, kThreads, 0, st>>>(g->dc, cache, batch, l, g->bmax, g->drop_seed, g->drop_p, step);
    else
      k_fusion_fwd<false><<<fgrid, kThreads, 0, st>>>(g->dc, cache, batch, l, g->bmax, g->drop_seed, g->drop_p, step);
    LAUNCH_CHECK(g);
  }
  if (g->multilabel) {
    if (train)
      k_head<true, true><<<g->n_cand, kHeadThreads, g->smem_head, st>>>(g->dc, cache, batch, g->bmax, g->hs_ld, g->lg_ld, g->adam, step_size, bc2_sqrt, ho);
    else
      k_head<false, true><<<g->n_cand, kHeadThreads, g->smem_head, st>>>(g->dc, cache, batch, g->bmax, g->hs_ld, g->lg_ld, g->adam, step_size, bc2_sqrt, ho);
  } else if (train)
    k_head<true><<<g->n_cand, kHeadThreads, g->smem_head, st>>>(g->dc, cache, batch, g->bmax, g->hs_ld, g->lg_ld, g->adam, step_size, bc2_sqrt, ho);
  else
    k_head<false><<<g->n_cand, kHeadThreads, g->smem_head, st>>>(g->dc, cache, batch, g->bmax, g->hs_ld, g->lg_ld, g->adam, step_size, bc2_sqrt, ho);
  LAUNCH_CHECK(g);
  if (!train) return MFAS_OK;
  for (int l = g->Lmax - 1; l >= 0; --l) {
    k_dz<<<dim3((g->Hmax + 31) / 32, g->n_cand), kThreads, 0, st>>>(g->dc, l, batch.n_rows, g->bmax, g->adam, step_size,
                                                                  bc2_sqrt, g->drop_seed, g->drop_p, step);
    LAUNCH_CHECK(g);
    k_fusion_bwd<<<dim3((g->Kmax[l] + BWD_KT - 1) / BWD_KT, g->n_cand), kThreads, g->smem_bwd, st>>>(
        g->dc, cache, batch, l, g->bmax, g->adam, step_size, bc2_sqrt);
    LAUNCH_CHECK(g);
  }
  if (g->any_alphas) {
    const int per = 256 / MFAS_MAX_LAYERS;
    k_alpha_step<<<(g->n_cand + per - 1) / per, 256, 0, st>>>(g->dc, g->n_cand, g->adam, step_size, bc2_sqrt);
    LAUNCH_CHECK(g);
  }
  return MFAS_OK;
}

static int check_batch(const mfas_group* g, int n_rows, bool train) {
  if (n_rows < 1 || n_rows > g->bmax) return fail(MFAS_ERR_INVALID, "n_rows=%d outside [1,%d]", n_rows, g->bmax);
  if (train && n_rows < 2)
    for (int c = 0; c < g->n_cand; ++c)
      if (g->lay[c].flags & MFAS_FLAG_BN)   // torch: "Expected more than 1 value per channel when training"
        return fail(MFAS_ERR_INVALID, "Expected more than 1 value per channel when training (batch of %d row)", n_rows);
  return MFAS_OK;
}

extern "C" int mfas_forward(mfas_group_t g, const mfas_cache_desc* cache, const int32_t* d_rows, int64_t rows_stride,
                            int32_t n_rows, int32_t train, int64_t step, float* d_logits, float* d_loss,
                            int32_t* d_correct, void* stream) {
  if (!g) return fail(MFAS_ERR_INVALID, "null group");
  DeviceGuard dg(g->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = check_batch(g, n_rows, train != 0);
  if (rc) return rc;
  DCache dc;
  if ((rc = to_dcache(g, cache, &dc))) return rc;
  if ((rc = sync_descriptors(g, st))) return rc;
  BatchRef b{d_rows, rows_stride, 0, n_rows};
  HeadOut ho{d_logits, d_loss, d_correct, nullptr, 0, 0};
  return launch_step(g, dc, b, false, train != 0, 0.f, 1.f, (uint32_t)step, ho, st);
}

extern "C" int mfas_train_step(mfas_group_t g, const mfas_cache_desc* cache, const int32_t* d_rows, int64_t rows_stride,
                               int32_t n_rows, float step_size, float bc2_sqrt, int64_t step, float* d_logits,
                               float* d_loss, int32_t* d_correct, void* stream) {
  if (!g) return fail(MFAS_ERR_INVALID, "null group");
  DeviceGuard dg(g->device);
  cudaStream_t st = (cudaStream_t)stream;
  int rc = check_batch(g, n_rows, true);
  if (rc) return rc;
  DCache dc;
  if ((rc = to_dcache(g, cache, &dc))) return rc;
  if ((rc = sync_descriptors(g, st))) return rc;
  BatchRef b{d_rows, rows_stride, 0, n_rows};
  HeadOut ho{d_logits, d_loss, d_correct, nullptr, 0, 0};
  return launch_step(g, dc, b, true, true, step_size, bc2_sqrt, (uint32_t)step, ho, st);
}

static int snapshot(mfas_group* g, int force, int dir, cudaStream_t st) {
  k_snapshot<<<dim3(32, g->n_cand), kThreads, 0, st>>>(g->dc, g->improved, force, dir);
  LAUNCH_CHECK(g);
  return MFAS_OK;
}

extern "C" int mfas_train_run(mfas_group_t g, const mfas_cache_desc* train, const mfas_cache_desc* dev,
                              const mfas_run_args* a, void* stream) {
  if (!g || !a) return fail(MFAS_ERR_INVALID, "null argument");
  DeviceGuard dg(g->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->n_epochs < 0 || a->batch < 1 || a->batch > g->bmax)
    return fail(MFAS_ERR_INVALID, "n_epochs=%d batch=%d (group batch_max=%d)", a->n_epochs, a->batch, g->bmax);
  if (!a->perm_train || !a->step_size || !a->bc2_sqrt || !a->stats || !a->best_acc || !a->best_epoch)
    return fail(MFAS_ERR_INVALID, "perm_train/step_size/bc2_sqrt/stats/best_acc/best_epoch must be non-null");
  DCache dtr, ddv;
  int rc;
  if ((rc = to_dcache(g, train, &dtr))) return rc;
  if ((rc = to_dcache(g, dev, &ddv))) return rc;
  const long long ntr = dtr.n_rows, ndv = ddv.n_rows;
  const int B = a->batch;
  const long long steps_tr = (ntr + B - 1) / B;
  const int last_tr = (int)(ntr - (steps_tr - 1) * B);
  if ((rc = check_batch(g, last_tr, true))) return rc;
  if ((rc = sync_descriptors(g, st))) return rc;

  const int E = a->n_epochs;
  CUDA_TRY(cudaMemsetAsync(a->stats, 0, sizeof(double) * 4 * (size_t)E * g->n_cand, st));
  if (a->best_acc_init)   // pageable source: staged before the call returns
    CUDA_TRY(cudaMemcpyAsync(a->best_acc, a->best_acc_init, sizeof(double) * g->n_cand, cudaMemcpyHostToDevice, st));
  else
    CUDA_TRY(cudaMemsetAsync(a->best_acc, 0, sizeof(double) * g->n_cand, st));
  CUDA_TRY(cudaMemsetAsync(a->best_epoch, 0xFF, sizeof(int32_t) * g->n_cand, st));
  if ((rc = snapshot(g, 1, 0, st))) return rc;            // best_model_sd = deepcopy(state_dict)  (ntu.py:17)
  const long long stat_stride = 4LL * E;
  long long t = 0;
  for (int e = 0; e < E; ++e) {
    for (long long s = 0; s < steps_tr; ++s, ++t) {        // phase 'train'
      const int n = (int)((s == steps_tr - 1) ? last_tr : B);
      BatchRef b{a->perm_train, (long long)E * ntr, (long long)e * ntr + s * B, n};
      HeadOut ho{nullptr, nullptr, nullptr, a->stats, stat_stride, 4LL * e};
      if ((rc = launch_step(g, dtr, b, true, true, a->step_size[t], a->bc2_sqrt[t], (uint32_t)(a->adam_t0 + t), ho, st)))
        return rc;
    }
    // phase 'dev': rows are independent in eval mode (running BatchNorm statistics, no dropout), so the pass may use
    // its own step width -- 128 rows when the wide eval path is on: the weights are streamed half as often.  Accuracy
    // counts are unchanged; the size-weighted loss sum differs by fp32 summation order only.
    const int Bd = g->ev128 ? 128 : B;
    const long long steps_dv_w = (ndv + Bd - 1) / Bd;
    for (long long s = 0; s < steps_dv_w; ++s) {
      const int n = (int)((s == steps_dv_w - 1) ? ndv - (steps_dv_w - 1) * Bd : Bd);
      BatchRef b{a->perm_dev, (long long)E * ndv, (a->perm_dev ? (long long)e * ndv : 0) + s * Bd, n};
      HeadOut ho{nullptr, nullptr, nullptr, a->stats, stat_stride, 4LL * e + 2};
      if ((rc = launch_step(g, ddv, b, false, false, 0.f, 1.f, 0u, ho, st))) return rc;
    }
    k_best_update<<<(g->n_cand + 127) / 128, 128, 0, st>>>(g->n_cand, a->stats, stat_stride, 4LL * e + 2, ndv, e,
                                                           a->best_acc, a->best_epoch, g->improved);
    LAUNCH_CHECK(g);
    if ((rc = snapshot(g, 0, 0, st))) return rc;           // deepcopy on improvement (ntu.py:82-84)
  }
  return snapshot(g, 1, 1, st);                            // model.load_state_dict(best_model_sd) (ntu.py:86)
}

extern "C" int mfas_eval_pass(mfas_group_t g, const mfas_cache_desc* cache, const int32_t* d_perm, int32_t batch,
                              double* d_out, void* stream) {
  if (!g || !d_out) return fail(MFAS_ERR_INVALID, "null argument");
  DeviceGuard dg(g->device);
  cudaStream_t st = (cudaStream_t)stream;
  if (batch < 1 || batch > g->bmax) return fail(MFAS_ERR_INVALID, "batch=%d (group batch_max=%d)", batch, g->bmax);
  DCache dc;
  int rc;
  if ((rc = to_dcache(g, cache, &dc))) return rc;
  if ((rc = sync_descriptors(g, st))) return rc;
  CUDA_TRY(cudaMemsetAsync(d_out, 0, sizeof(double) * 2 * g->n_cand, st));
  if (g->ev128) batch = 128;                               // (see mfas_train_run, phase 'dev')
  const long long n = dc.n_rows, steps = (n + batch - 1) / batch;
  for (long long s = 0; s < steps; ++s) {
    const int nr = (int)((s == steps - 1) ? n - (steps - 1) * batch : batch);
    BatchRef b{d_perm, n, s * batch, nr};
    HeadOut ho{nullptr, nullptr, nullptr, d_out, 2, 0};
    if ((rc = launch_step(g, dc, b, false, false, 0.f, 1.f, 0u, ho, st))) return rc;
  }
  return MFAS_OK;
}

// ---------------------------------------------------------------------------------------------
// feature-cache builder (the step before the path)
// ---------------------------------------------------------------------------------------------
extern "C" int mfas_global_pool(int32_t device, const float* d_in, int64_t B, int64_t C, int64_t S, float* d_out,
                                int64_t out_ld, void* stream) {
  if (!d_in || !d_out) return fail(MFAS_ERR_INVALID, "null argument");
  if (B < 1 || C < 1 || S < 1 || out_ld < C) return fail(MFAS_ERR_INVALID, "B=%lld C=%lld S=%lld out_ld=%lld", (long long)B, (long long)C, (long long)S, (long long)out_ld);
  if (((uintptr_t)d_in | (uintptr_t)d_out) & 3) return fail(MFAS_ERR_INVALID, "pointers must be 4-byte aligned");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(MFAS_ERR_INVALID, "device %d of %d", device, ndev);
  DeviceGuard dg(device);
  if (!dg.ok) return fail(MFAS_ERR_CUDA, "cudaSetDevice(%d) failed", device);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device);
  const bool short_rows = S < 2048;                      // four rows per warp (see kernels_pool.cuh)
  const long long rows = (long long)B * C, per = (kPoolThreads / 32) * (short_rows ? 4 : 1);
  const long long want = (rows + per - 1) / per, cap = (long long)nsm * 8;      // 8 CTAs of 256 threads per SM: full occupancy
  const int grid = (int)(want < cap ? want : cap);
  const int vec4 = (S % 4 == 0) && (((uintptr_t)d_in & 15) == 0);
  if (short_rows) k_global_pool<8><<<grid, kPoolThreads, 0, (cudaStream_t)stream>>>(d_in, rows, C, S, d_out, out_ld, vec4);
  else k_global_pool<32><<<grid, kPoolThreads, 0, (cudaStream_t)stream>>>(d_in, rows, C, S, d_out, out_ld, vec4);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MFAS_ERR_CUDA, "k_global_pool launch failed: %s", cudaGetErrorString(e));
  return MFAS_OK;
}
