// common.cuh -- device/host primitives shared by every dint_b200 kernel.
//
//  * fasthash64: the key -> slot hash of the reference (lock_2pl/udp/utils.h:20-57; ten identical
//    copies across the reference's udp/ directories), specialised for the two lengths the hot path
//    uses: len 4 (lock ids, lock_2pl/udp/server.cc:71) and len 8 (KV keys, store/udp/kvs.h:33-35).
//  * FastMod: exact u64 % u32 for a run-time divisor (table sizes are configuration, not constants)
//    by multiply-high -- a 64-bit hardware divide does not exist on the SM.
//  * unaligned wire access: the reference's datagrams are #pragma pack(1) structs of 6/9/23/53/55
//    bytes, so fields sit at arbitrary byte offsets inside a staged tile.
//  * TMA (cp.async.bulk) 1-D staging of a tile of wire records global<->shared.
#pragma once
#include <cstdint>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#endif

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

namespace dint {

#define DINT_HD __host__ __device__ __forceinline__
#define DINT_D __device__ __forceinline__

// ------------------------------------------------------------------------------------------------
// fasthash64 (seed 0xdeadbeef).  h0 = seed ^ (len * m).
// ------------------------------------------------------------------------------------------------
static constexpr uint64_t kFhM = 0x880355f21e6d1965ULL;
static constexpr uint64_t kFhSeed = 0xdeadbeefULL;

DINT_HD uint64_t fh_mix(uint64_t h) {
  h ^= h >> 23;
  h *= 0x2127599bf4325c37ULL;
  h ^= h >> 47;
  return h;
}
// len == 4: no 8-byte words, the tail switch folds the 4 bytes into v (utils.h:44-54).
DINT_HD uint64_t fasthash64_u32(uint32_t x) {
  uint64_t h = kFhSeed ^ (4ULL * kFhM);
  h ^= fh_mix((uint64_t)x);
  h *= kFhM;
  return fh_mix(h);
}
// len == 8: one pass of the word loop (utils.h:35-39), empty tail.
DINT_HD uint64_t fasthash64_u64(uint64_t x) {
  uint64_t h = kFhSeed ^ (8ULL * kFhM);
  h ^= fh_mix(x);
  h *= kFhM;
  return fh_mix(h);
}

// ------------------------------------------------------------------------------------------------
// Exact n % d, n any u64, 1 <= d < 2^32 (Granlund & Montgomery, "Division by invariant integers
// using multiplication", fig. 4.1 with N = 64): m' = floor(2^64 (2^l - d) / d) + 1, l = ceil(log2 d),
// t = mulhi(m', n), q = (t + ((n - t) >> 1)) >> (l - 1).
// ------------------------------------------------------------------------------------------------
struct FastMod {
  uint64_t magic;
  uint32_t d;
  uint32_t shift;   // l - 1, or log2(d) when pow2
  uint32_t pow2;
  uint32_t _pad;
};

inline FastMod make_fastmod(uint32_t d) {
  FastMod f{};
  f.d = d ? d : 1;
  d = f.d;
  if ((d & (d - 1)) == 0) {
    f.pow2 = 1;
    f.shift = 0;
    while ((1u << f.shift) < d) f.shift++;
    f.magic = 0;
    return f;
  }
  uint32_t l = 0;
  while ((1ULL << l) < d) l++;
  unsigned __int128 num = (unsigned __int128)((1ULL << l) - d) << 64;
  f.magic = (uint64_t)(num / d) + 1;
  f.shift = l - 1;
  f.pow2 = 0;
  return f;
}

DINT_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
DINT_HD uint64_t fast_div(uint64_t n, const FastMod& f) {
  if (f.pow2) return n >> f.shift;
  uint64_t t = mulhi64(f.magic, n);
  return (t + ((n - t) >> 1)) >> f.shift;
}
DINT_HD uint32_t fast_mod(uint64_t n, const FastMod& f) {
  if (f.pow2) return (uint32_t)(n & (uint64_t)(f.d - 1));
  return (uint32_t)(n - fast_div(n, f) * f.d);
}

#ifdef __CUDACC__
// ------------------------------------------------------------------------------------------------
// Unaligned little-endian field access through a generic pointer (shared tile or global memory).
// Built from ALIGNED 32-bit accesses + funnel shifts so that a thread touching a 40-byte value at
// an odd offset issues 11 word accesses instead of 40 byte accesses.
// ------------------------------------------------------------------------------------------------
DINT_D uint32_t ld_u32_unaligned(const uint8_t* p) {
  uintptr_t a = (uintptr_t)p;
  const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
  uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t lo = w[0];
  if (sh == 0) return lo;
  uint32_t hi = w[1];
  return __funnelshift_r(lo, hi, sh);
}
DINT_D uint64_t ld_u64_unaligned(const uint8_t* p) {
  uintptr_t a = (uintptr_t)p;
  const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
  uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t w0 = w[0], w1 = w[1];
  if (sh == 0) return ((uint64_t)w1 << 32) | w0;
  uint32_t w2 = w[2];
  return ((uint64_t)__funnelshift_r(w1, w2, sh) << 32) | __funnelshift_r(w0, w1, sh);
}
// Byte-granular store of one u32 at any alignment: touches only its own 4 bytes, so neighbouring
// records owned by other threads are never read-modify-written.
DINT_D void st_u32_unaligned(uint8_t* p, uint32_t v) {
  if (((uintptr_t)p & 3) == 0) {
    *(uint32_t*)p = v;
  } else {
    p[0] = (uint8_t)v;
    p[1] = (uint8_t)(v >> 8);
    p[2] = (uint8_t)(v >> 16);
    p[3] = (uint8_t)(v >> 24);
  }
}
// Load NW consecutive words starting at an arbitrary byte address.  Branch-free: lanes of a warp sit at
// all four alignments (record sizes are odd), so alignment-dependent branches would serialise.
template <int NW>
DINT_D void ld_words_unaligned(const uint8_t* p, uint32_t (&out)[NW]) {
  uintptr_t a = (uintptr_t)p;
  const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
  const uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t prev = w[0];
#pragma unroll
  for (int i = 0; i < NW; i++) {
    uint32_t nxt = (i + 1 < NW || sh) ? w[i + 1] : 0u;   // the word past the field is only touched when it holds field bytes
    out[i] = __funnelshift_r(prev, nxt, sh);
    prev = nxt;
  }
}
// Store NW consecutive words at an arbitrary byte address: aligned interior words, predicated byte
// stores for the (<=3 + <=3) edge bytes -- never a read-modify-write of bytes owned by a neighbour.
template <int NW>
DINT_D void st_words_unaligned(uint8_t* p, const uint32_t (&v)[NW]) {
  const uint32_t mis = (uint32_t)((uintptr_t)p & 3);
  uint32_t* w = (uint32_t*)(p - mis);                    // aligned word holding the first byte
  const uint32_t sh = mis * 8;
  if (mis == 0) w[0] = v[0];
#pragma unroll
  for (int b = 0; b < 3; b++)                            // head: bytes mis..3 of w[0]
    if (mis != 0 && mis + b < 4) p[b] = (uint8_t)(v[0] >> (8 * b));
#pragma unroll
  for (int i = 1; i < NW; i++) w[i] = __funnelshift_l(v[i - 1], v[i], sh);
  uint8_t* t = (uint8_t*)(w + NW);
#pragma unroll
  for (int b = 0; b < 3; b++)                            // tail: bytes 0..mis-1 of w[NW]
    if ((uint32_t)b < mis) t[b] = (uint8_t)(v[NW - 1] >> (32 - sh + 8 * b));
}

// Copy one packed wire record of MSG bytes between arbitrary byte addresses with word-wide accesses.
template <int MSG>
DINT_D void copy_record(uint8_t* dst, const uint8_t* src) {
  constexpr int NW = MSG / 4, TAIL = MSG % 4;
  if constexpr (NW > 0) {
    uint32_t w[NW];
    ld_words_unaligned<NW>(src, w);
    st_words_unaligned<NW>(dst, w);
  }
#pragma unroll
  for (int b = 0; b < TAIL; b++) dst[NW * 4 + b] = src[NW * 4 + b];
}

// ------------------------------------------------------------------------------------------------
// TMA 1-D bulk copies (cp.async.bulk; SASS UBLKCP).  Sizes and both addresses are 16-byte multiples.
// ------------------------------------------------------------------------------------------------
DINT_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DINT_D void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DINT_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
DINT_D void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
DINT_D void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
DINT_D void tma_store_1d(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_u32(smem_src)), "r"(bytes)
               : "memory");
}
DINT_D void tma_store_commit_and_wait() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
DINT_D void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#endif  // __CUDACC__

}  // namespace dint
