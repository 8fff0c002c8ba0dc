// Host-side parameter initialisation: torch.nn.init.{kaiming_uniform_, uniform_} on the CPU draw one
// 32-bit mt19937 output per float from the global generator, strictly serially (~3 ns per element;
// 0.4 s for the 133 M parameters of 128 cfg2 candidates -- 40 % of an end-to-end search iteration).
// This restates that stream -- at::mt19937 (ATen/core/MT19937RNGEngine.h) and
// at::uniform_real_distribution<float> (ATen/core/TransformationHelper.h: (y & 0xFFFFFF) * 2^-24 *
// (to - from) + from) -- in bulk: the 624-word state is regenerated and tempered block by block with
// loops the compiler vectorises.  The caller reads torch.get_rng_state() (legacy 5056-byte layout),
// passes it here, and stores the advanced state back, so a given torch.manual_seed() yields the same
// weights as the reference constructor would.  mfas_b200/host_init.py cross-checks the first draws against
// torch itself and falls back to torch's initialisers on any mismatch.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/mfas_b200.h"

namespace {
constexpr int N = 624, M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfu, UMASK = 0x80000000u, LMASK = 0x7fffffffu;

struct Engine {
  uint32_t st[N + 1];
  int pos, rem;          // next word to hand out; words left in the current block
};

inline uint32_t twist(uint32_t u, uint32_t v) { return (((u & UMASK) | (v & LMASK)) >> 1) ^ ((0u - (v & 1u)) & MATRIX_A); }

void regen_scalar(uint32_t* p) {
  for (int i = 0; i < N - M; ++i) p[i] = p[i + M] ^ twist(p[i], p[i + 1]);
  for (int i = N - M; i < N - 1; ++i) p[i] = p[i + M - N] ^ twist(p[i], p[i + 1]);
  p[N - 1] = p[M - 1] ^ twist(p[N - 1], p[0]);
}

inline uint32_t temper(uint32_t y) {
  y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
  return y;
}

void emit_scalar(const uint32_t* s, float* out, int k, float from, float range, int fma) {
  for (int i = 0; i < k; ++i) {
    const float x = (float)(temper(s[i]) & 0xFFFFFFu) * 5.9604644775390625e-08f;
    if (fma) out[i] = __builtin_fmaf(x, range, from);
    else { volatile float prod = x * range; out[i] = prod + from; }       // two roundings, no contraction
  }
}

#if defined(__x86_64__)
// 8 lanes at a time.  Every vector step reads p[i+1 .. i+8] and p[i+M ..] before it writes p[i .. i+7], and the
// second loop's p[i+M-N] was written 227 elements earlier, so the order of reads and writes matches the scalar loop.
__attribute__((target("avx2,fma"))) inline __m256i twist8(__m256i u, __m256i v) {
  const __m256i y = _mm256_or_si256(_mm256_and_si256(u, _mm256_set1_epi32((int)UMASK)), _mm256_and_si256(v, _mm256_set1_epi32((int)LMASK)));
  const __m256i mag = _mm256_and_si256(_mm256_sub_epi32(_mm256_setzero_si256(), _mm256_and_si256(v, _mm256_set1_epi32(1))),
                                       _mm256_set1_epi32((int)MATRIX_A));
  return _mm256_xor_si256(_mm256_srli_epi32(y, 1), mag);
}
__attribute__((target("avx2,fma"))) void regen_avx2(uint32_t* p) {
  int i = 0;
  for (; i + 8 <= N - M; i += 8) {
    const __m256i u = _mm256_loadu_si256((const __m256i*)(p + i)), v = _mm256_loadu_si256((const __m256i*)(p + i + 1));
    const __m256i m = _mm256_loadu_si256((const __m256i*)(p + i + M));
    _mm256_storeu_si256((__m256i*)(p + i), _mm256_xor_si256(m, twist8(u, v)));
  }
  for (; i < N - M; ++i) p[i] = p[i + M] ^ twist(p[i], p[i + 1]);
  for (; i + 8 <= N - 1; i += 8) {
    const __m256i u = _mm256_loadu_si256((const __m256i*)(p + i)), v = _mm256_loadu_si256((const __m256i*)(p + i + 1));
    const __m256i m = _mm256_loadu_si256((const __m256i*)(p + i + M - N));
    _mm256_storeu_si256((__m256i*)(p + i), _mm256_xor_si256(m, twist8(u, v)));
  }
  for (; i < N - 1; ++i) p[i] = p[i + M - N] ^ twist(p[i], p[i + 1]);
  p[N - 1] = p[M - 1] ^ twist(p[N - 1], p[0]);
}
__attribute__((target("avx2,fma"))) void emit_avx2(const uint32_t* s, float* out, int k, float from, float range, int fma) {
  const __m256 vr = _mm256_set1_ps(range), vf = _mm256_set1_ps(from), sc = _mm256_set1_ps(5.9604644775390625e-08f);
  int i = 0;
  for (; i + 8 <= k; i += 8) {
    __m256i y = _mm256_loadu_si256((const __m256i*)(s + i));
    y = _mm256_xor_si256(y, _mm256_srli_epi32(y, 11));
    y = _mm256_xor_si256(y, _mm256_and_si256(_mm256_slli_epi32(y, 7), _mm256_set1_epi32((int)0x9d2c5680u)));
    y = _mm256_xor_si256(y, _mm256_and_si256(_mm256_slli_epi32(y, 15), _mm256_set1_epi32((int)0xefc60000u)));
    y = _mm256_xor_si256(y, _mm256_srli_epi32(y, 18));
    const __m256 x = _mm256_mul_ps(_mm256_cvtepi32_ps(_mm256_and_si256(y, _mm256_set1_epi32(0xFFFFFF))), sc);
    _mm256_storeu_ps(out + i, fma ? _mm256_fmadd_ps(x, vr, vf) : _mm256_add_ps(_mm256_mul_ps(x, vr), vf));
  }
  if (i < k) emit_scalar(s + i, out + i, k - i, from, range, fma);
}
#endif

// (Also measured in the build container and dropped, 154 M-word fill with 2 consumers at 64-69 ms: regenerating out of place straight
// into the ring instead of regen + memcpy: 64-68 ms, within noise -- the consumers' stores set the pace, not the producer;
// non-temporal stores in emit_avx2: 80 ms.)
// (A 16-lane AVX-512F twist was measured and dropped: on the build container's Xeon the 154 M-word fill went from 68 to 120 ms
// with 2 consumers and from 114 to 133 ms serial -- 512-bit unaligned loads and the licence down-clock cost more than the lanes gain.)

// tempered words as they are (count < 0 ops: the caller derives non-uniform draws from them, e.g. Box-Muller normals)
void emit_raw(const uint32_t* s, uint32_t* out, int k) {
  for (int i = 0; i < k; ++i) out[i] = temper(s[i]);
}

void regen(uint32_t* p) {
#if defined(__x86_64__)
  static const bool fast = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
  if (fast) return regen_avx2(p);
#endif
  regen_scalar(p);
}
void emit(const uint32_t* s, float* out, int k, float from, float range, int fma) {
#if defined(__x86_64__)
  static const bool fast = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma");
  if (fast) return emit_avx2(s, out, k, from, range, fma);
#endif
  emit_scalar(s, out, k, from, range, fma);
}
}  // namespace

// ---- pipelined variant for large fills -------------------------------------------------------------------
// The twist is a serial recurrence (one thread, ~0.3 ns/word); tempering, the float conversion and above all the
// stores into the pinned arena are not.  One producer thread regenerates state blocks into a ring of chunks,
// kConsumers threads temper/scale/store disjoint chunks of the word stream.  The word stream is cut into blocks:
// block 0 = what is left of the current state (rem0 words), block b >= 1 = one full regeneration (624 words);
// a chunk is kChunkBlocks consecutive blocks.  Result and final generator state are identical to the serial loop.
namespace {
constexpr int kChunkBlocks = 64, kRing = 24;

struct Stream {
  int n_ops;
  float* const* dst;
  const int64_t* count;
  const float* from;
  const float* to;
  std::vector<int64_t> start;      // first word of op i; start[n_ops] = total
  int fma;
};

// words [w0, w0 + k) of the stream come from src[0..k)
void emit_range(const Stream& S, int64_t w0, const uint32_t* src, int64_t k) {
  int lo = 0, hi = S.n_ops - 1;                      // op containing w0 (ops with count 0 are skipped by the search)
  while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (S.start[mid] <= w0) lo = mid; else hi = mid - 1; }
  int op = lo;
  while (k > 0) {
    while (S.start[op + 1] <= w0) ++op;
    const int64_t in_op = w0 - S.start[op], take = (S.start[op + 1] - w0 < k) ? S.start[op + 1] - w0 : k;
    if (S.count[op] < 0) emit_raw(src, reinterpret_cast<uint32_t*>(S.dst[op]) + in_op, (int)take);
    else emit(src, S.dst[op] + in_op, (int)take, S.from[op], S.to[op] - S.from[op], S.fma);
    src += take; w0 += take; k -= take;
  }
}

void fill_pipelined(Engine& e, const Stream& S, int n_consumers) {
  const int64_t total = S.start[S.n_ops];
  const int rem0 = e.rem;
  const int64_t n_blocks = 1 + (total > rem0 ? (total - rem0 + N - 1) / N : 0);       // block 0 + regenerated blocks
  const int64_t n_chunks = (n_blocks + kChunkBlocks - 1) / kChunkBlocks;
  std::vector<uint32_t> ring((size_t)kRing * kChunkBlocks * N);
  struct alignas(128) Flag { std::atomic<int64_t> v; };   // own cache line each: the spinning readers must not slow the writers next door
  std::vector<Flag> ready(kRing);                    // chunk index + 1 held by the slot (0: never filled); consumer stores -(chunk+1) when done
  for (auto& r : ready) r.v.store(0, std::memory_order_relaxed);
  auto block_words = [&](int64_t b, int64_t& w0) {   // words of block b that belong to the stream, and its first word
    if (b == 0) { w0 = 0; return (int64_t)(total < rem0 ? total : rem0); }
    w0 = rem0 + (b - 1) * N;
    const int64_t left = total - w0;
    return left < N ? (left < 0 ? 0 : left) : (int64_t)N;
  };
  std::vector<std::thread> consumers;
  for (int c = 0; c < n_consumers; ++c)
    consumers.emplace_back([&, c] {
      for (int64_t ch = c; ch < n_chunks; ch += n_consumers) {
        std::atomic<int64_t>& slot = ready[ch % kRing].v;
        while (slot.load(std::memory_order_acquire) != ch + 1) {
#if defined(__x86_64__)
          _mm_pause();
#endif
        }
        const uint32_t* base = ring.data() + (size_t)(ch % kRing) * kChunkBlocks * N;
        const int64_t b0 = ch * kChunkBlocks, b1 = b0 + kChunkBlocks < n_blocks ? b0 + kChunkBlocks : n_blocks;
        for (int64_t b = b0; b < b1; ++b) {
          int64_t w0;
          const int64_t k = block_words(b, w0);
          if (k > 0) emit_range(S, w0, base + (size_t)(b - b0) * N, k);
        }
        slot.store(-(ch + 1), std::memory_order_release);
      }
    });
  // producer (this thread)
  int64_t last_used = 0;                             // words of the stream taken from the last generated block
  for (int64_t ch = 0; ch < n_chunks; ++ch) {
    std::atomic<int64_t>& slot = ready[ch % kRing].v;
    if (ch >= kRing)
      while (slot.load(std::memory_order_acquire) != -(ch - kRing + 1)) {
#if defined(__x86_64__)
        _mm_pause();
#endif
      }
    uint32_t* base = ring.data() + (size_t)(ch % kRing) * kChunkBlocks * N;
    const int64_t b0 = ch * kChunkBlocks, b1 = b0 + kChunkBlocks < n_blocks ? b0 + kChunkBlocks : n_blocks;
    for (int64_t b = b0; b < b1; ++b) {
      uint32_t* o = base + (size_t)(b - b0) * N;
      if (b == 0) memcpy(o, e.st + e.pos, sizeof(uint32_t) * rem0);
      else { regen(e.st); memcpy(o, e.st, sizeof(uint32_t) * N); }
      int64_t w0;
      last_used = block_words(b, w0);
    }
    slot.store(ch + 1, std::memory_order_release);
  }
  for (auto& t : consumers) t.join();
  if (n_blocks == 1) { e.pos += (int)last_used; e.rem -= (int)last_used; }
  else { e.pos = (int)last_used; e.rem = N - (int)last_used; }
}
}  // namespace

// torch.get_rng_state() legacy layout: u64 seed | i32 left | i32 seeded | u64 next | u64 state[624] | ...
extern "C" int mfas_host_uniform_fill(uint8_t* torch_rng_state, int64_t state_bytes, int32_t n_ops,
                                      float* const* dst, const int64_t* count, const float* from, const float* to,
                                      int32_t use_fma) {
  if (!torch_rng_state || state_bytes < 24 + 8 * N || n_ops < 0 || (n_ops && (!dst || !count || !from || !to))) return MFAS_ERR_INVALID;
  int32_t left, seeded;
  uint64_t next;
  memcpy(&left, torch_rng_state + 8, 4);
  memcpy(&seeded, torch_rng_state + 12, 4);
  memcpy(&next, torch_rng_state + 16, 8);
  if (!seeded || left < 1 || left > N || next > (uint64_t)N) return MFAS_ERR_UNSUPPORTED;
  Engine e;
  for (int i = 0; i < N; ++i) { uint64_t w; memcpy(&w, torch_rng_state + 24 + 8 * i, 8); e.st[i] = (uint32_t)w; }
  e.pos = (int)next; e.rem = left - 1;
  int64_t total = 0;
  for (int op = 0; op < n_ops; ++op) total += count[op] < 0 ? -count[op] : count[op];     // count < 0: -count raw tempered words
  int n_consumers = 0;
  if (total >= (4 << 20)) {                          // MFAS_HOST_INIT_THREADS: consumer threads (0 = the serial loop)
    const char* te = getenv("MFAS_HOST_INIT_THREADS");
    const unsigned hw = std::thread::hardware_concurrency();
    n_consumers = te ? atoi(te) : (hw >= 4 ? 2 : 0);     // B200 box, 16 cores: 0 -> 112 ms, 2 -> 70 ms, 3 -> 85 ms, 5 -> 90 ms for 154 M words
    if (n_consumers > 16) n_consumers = 16;
  }
  if (n_consumers > 0) {
    Stream S{n_ops, dst, count, from, to, std::vector<int64_t>((size_t)n_ops + 1, 0), use_fma};
    for (int op = 0; op < n_ops; ++op) S.start[op + 1] = S.start[op] + (count[op] < 0 ? -count[op] : count[op]);
    fill_pipelined(e, S, n_consumers);
  } else {
    for (int op = 0; op < n_ops; ++op) {
      float* o = dst[op];
      const bool raw = count[op] < 0;
      int64_t n = raw ? -count[op] : count[op];
      const float range = to[op] - from[op];
      while (n > 0) {
        if (e.rem == 0) { regen(e.st); e.pos = 0; e.rem = N; }
        const int k = (int)(n < e.rem ? n : e.rem);
        if (raw) emit_raw(e.st + e.pos, reinterpret_cast<uint32_t*>(o), k);
        else emit(e.st + e.pos, o, k, from[op], range, use_fma);
        o += k; n -= k; e.pos += k; e.rem -= k;
      }
    }
  }
  left = e.rem + 1; next = (uint64_t)e.pos;
  memcpy(torch_rng_state + 8, &left, 4);
  memcpy(torch_rng_state + 16, &next, 8);
  for (int i = 0; i < N; ++i) { uint64_t w = e.st[i]; memcpy(torch_rng_state + 24 + 8 * i, &w, 8); }
  return MFAS_OK;
}
