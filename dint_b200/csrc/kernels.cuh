// kernels.cuh -- the three launches of a chunk (see engine.cuh for the scheme).
#pragma once
#include <cooperative_groups.h>
#include "engine.cuh"

namespace dint {
namespace cg = cooperative_groups;

DINT_D uint32_t lane_id() { return threadIdx.x & 31; }
DINT_D uint32_t warp_id() { return threadIdx.x >> 5; }

// In-tile exclusive rank of `flag` in thread order; returns rank, writes the tile total to `total`.
// `scratch` is kThreads/32 words of shared memory.  Contains __syncthreads().
DINT_D uint32_t tile_rank(bool flag, uint32_t* scratch, uint32_t& total) {
  uint32_t bal = __ballot_sync(0xffffffffu, flag);
  uint32_t in_warp = __popc(bal & ((1u << lane_id()) - 1u));
  if (lane_id() == 0) scratch[warp_id()] = __popc(bal);
  __syncthreads();
  uint32_t off = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; w++) {
    uint32_t v = scratch[w];
    if (w < (int)warp_id()) off += v;
    tot += v;
  }
  __syncthreads();
  total = tot;
  return off + in_warp;
}

// ---------------------------------------------------------------------------------------------------
// K1 classify
// ---------------------------------------------------------------------------------------------------
template <int KIND, bool HAS_LOG>
__global__ void __launch_bounds__(kThreads) k_classify(const Ctx c) {
  using W = Wire<KIND>;
  __shared__ __align__(16) uint8_t tile[kTile * W::MSG + 16];
  __shared__ uint64_t bar;
  __shared__ uint32_t scratch[kThreads / 32];
  const uint32_t t = blockIdx.x, first = t * kTile;
  const uint32_t cnt = min((uint32_t)kTile, c.n - first);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  stage_in(tile, c.req + (size_t)first * W::MSG, cnt * W::MSG, &bar, 0);

  const bool valid = threadIdx.x < cnt;
  bool is_log = false;
  if (valid) {
    const uint8_t* rec = tile + threadIdx.x * W::MSG;
    TypeInfo ti = type_info<KIND>(rec);
    uint32_t g = kNoGroup;
    if (!ti.invalid && ti.mask) {
      if (!group_of<KIND>(c, rec, g)) g = kNoGroup;
      if (g != kNoGroup) {
        uint32_t* bm = c.bm;
        const uint32_t bw = c.bm_words;
        if (ti.mask & C_RA) bm_set(bm, g);                                        // R
        if (ti.mask & C_WA) { if (bm_fetch_set(bm + bw, g)) bm_set(bm + 2 * bw, g); }      // WA, WWA
        if (ti.mask & C_WL) { if (bm_fetch_set(bm + 3 * bw, g)) bm_set(bm + 4 * bw, g); }  // WL, WWL
      }
    }
    c.grp[first + threadIdx.x] = g;
    is_log = !ti.invalid && ti.is_log;
  }
  if (HAS_LOG) {
    uint32_t total;
    (void)tile_rank(is_log, scratch, total);
    if (threadIdx.x == 0) c.log_tilecnt[t] = total;
  }
}

// K1b: absolute append ordinal of every tile's first log append (single CTA).
__global__ void __launch_bounds__(kThreads) k_log_scan(const Ctx c) {
  __shared__ unsigned long long carry;
  __shared__ uint32_t wsum[kThreads / 32];
  if (threadIdx.x == 0) carry = c.log_total[0];
  __syncthreads();
  for (uint32_t base = 0; base < c.n_tiles; base += kThreads) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < c.n_tiles ? c.log_tilecnt[i] : 0;
    uint32_t x = v;                       // inclusive warp scan
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
      if ((int)lane_id() >= o) x += y;
    }
    if (lane_id() == 31) wsum[warp_id()] = x;
    __syncthreads();
    uint32_t woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; w++) {
      if (w < (int)warp_id()) woff += wsum[w];
      tot += wsum[w];
    }
    if (i < c.n_tiles) c.log_tilebase[i] = carry + woff + (x - v);
    __syncthreads();
    if (threadIdx.x == 0) carry += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    c.log_total[1] = carry;               // end ordinal of this chunk
    c.log_total[0] = carry;               // base of the next chunk
  }
}

// ---------------------------------------------------------------------------------------------------
// K2 apply
// ---------------------------------------------------------------------------------------------------
template <int KIND, bool HAS_LOG>
__global__ void __launch_bounds__(kThreads) k_apply(const Ctx c) {
  using W = Wire<KIND>;
  __shared__ __align__(16) uint8_t tile[kTile * W::MSG + 16];
  __shared__ uint64_t bar;
  __shared__ uint32_t scratch[kThreads / 32];
  const uint32_t t = blockIdx.x, first = t * kTile;
  const uint32_t cnt = min((uint32_t)kTile, c.n - first);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  stage_in(tile, c.req + (size_t)first * W::MSG, cnt * W::MSG, &bar, 0);

  const bool valid = threadIdx.x < cnt;
  uint8_t* rec = tile + threadIdx.x * W::MSG;
  TypeInfo ti{0, false, false};
  uint32_t g = kNoGroup;
  bool listed = false;
  if (valid) {
    ti = type_info<KIND>(rec);
    g = c.grp[first + threadIdx.x];
    if (!ti.invalid && ti.mask && g == kNoGroup) ti.invalid = true;   // not this shard's / bad table
    if (!ti.invalid && ti.mask) {
      const uint32_t* bm = c.bm;
      const uint32_t bw = c.bm_words;
      if ((ti.mask & C_RA) && bm_test(bm + bw, g)) listed = true;
      if ((ti.mask & C_WA) && (bm_test(bm, g) || bm_test(bm + 2 * bw, g))) listed = true;
      if ((ti.mask & C_WL) && bm_test(bm + 4 * bw, g)) listed = true;
    }
  }
  unsigned long long log_ord = 0;
  bool log_keep = false;
  if (HAS_LOG) {
    uint32_t total;
    uint32_t r = tile_rank(valid && !ti.invalid && ti.is_log, scratch, total);
    if (valid && !ti.invalid && ti.is_log) {
      log_ord = c.log_tilebase[t] + r;
      log_keep = log_ord + c.ring_n >= c.log_total[1];   // no later append of this chunk overwrites it
    }
  }
  {
    uint32_t total;
    uint32_t r = tile_rank(listed, scratch, total);
    if (listed) c.clist[(size_t)t * kTile + r] = first + threadIdx.x;
    if (threadIdx.x == 0) c.ccnt[t] = total;
  }
  if (valid) {
    if (ti.invalid) mark_invalid<KIND>(c, rec);
    else if (!listed) apply_one<KIND>(c, rec, g, log_ord, log_keep);
    // listed: the record leaves this kernel unchanged; K3 rewrites it in place in resp
  }
  stage_out(c.resp + (size_t)first * W::MSG, tile, cnt * W::MSG);
}

// ---------------------------------------------------------------------------------------------------
// K3 ordered replay (cooperative launch: the whole grid is co-resident and uses grid-wide barriers)
// ---------------------------------------------------------------------------------------------------
constexpr int kSortItems = 8;                         // items per thread per sort tile
constexpr int kSortTile = kThreads * kSortItems;      // 2048

template <int KIND>
DINT_D void replay_run(const Ctx& c, const uint64_t* sorted, uint32_t p, uint32_t nc) {
  using W = Wire<KIND>;
  const uint32_t g = (uint32_t)(sorted[p] >> 32);
  uint32_t len = 0;
  for (uint32_t q = p; q < nc; q++) {
    uint64_t e = sorted[q];
    if ((uint32_t)(e >> 32) != g) break;
    apply_one<KIND>(c, c.resp + (size_t)(uint32_t)e * W::MSG, g, 0, false);
    len++;
  }
  atomicMax(&c.counters[2], (unsigned long long)len);
}

template <int KIND>
__global__ void __launch_bounds__(kThreads) k_ordered(const Ctx c) {
  cg::grid_group grid = cg::this_grid();
  __shared__ uint64_t skeys[kSmallSort];              // 16 KB: small sort; reused as counters by the radix passes
  __shared__ uint32_t wsum[kThreads / 32];
  __shared__ uint32_t s_carry;
  const uint32_t tid = threadIdx.x;

  // ---- phase 0: exclusive prefix of the per-tile list lengths (CTA 0) ----------------------------
  if (blockIdx.x == 0) {
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < c.n_tiles; base += kThreads) {
      uint32_t i = base + tid;
      uint32_t v = i < c.n_tiles ? c.ccnt[i] : 0;
      uint32_t x = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if ((int)lane_id() >= o) x += y;
      }
      if (lane_id() == 31) wsum[warp_id()] = x;
      __syncthreads();
      uint32_t woff = 0, tot = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; w++) {
        if (w < (int)warp_id()) woff += wsum[w];
        tot += wsum[w];
      }
      if (i < c.n_tiles) c.cprefix[i] = s_carry + woff + (x - v);
      __syncthreads();
      if (tid == 0) s_carry += tot;
      __syncthreads();
    }
    if (tid == 0) c.cprefix[c.n_tiles] = s_carry;
  }
  grid.sync();
  const uint32_t nc = c.cprefix[c.n_tiles];

  if (nc != 0) {
    // ---- phase 1: densify the tile-segmented list into (group << 32 | index), index-ascending ----
    for (uint32_t t = blockIdx.x; t < c.n_tiles; t += gridDim.x) {
      uint32_t m = c.ccnt[t];
      if (tid < m) {
        uint32_t idx = c.clist[(size_t)t * kTile + tid];
        c.sortA[c.cprefix[t] + tid] = ((uint64_t)c.grp[idx] << 32) | idx;
      }
    }
    grid.sync();

    if (nc <= kSmallSort) {
      // ---- small path: bitonic sort of the full 64-bit keys in shared memory, one CTA --------------
      if (blockIdx.x == 0) {
        uint32_t npow = 1;
        while (npow < nc) npow <<= 1;
        for (uint32_t i = tid; i < npow; i += kThreads) skeys[i] = i < nc ? c.sortA[i] : ~0ULL;
        __syncthreads();
        for (uint32_t k = 2; k <= npow; k <<= 1) {
          for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < npow; i += kThreads) {
              uint32_t ixj = i ^ j;
              if (ixj > i) {
                uint64_t a = skeys[i], b = skeys[ixj];
                bool up = (i & k) == 0;
                if ((a > b) == up) { skeys[i] = b; skeys[ixj] = a; }
              }
            }
            __syncthreads();
          }
        }
        for (uint32_t p = tid; p < nc; p += kThreads)
          if (p == 0 || (uint32_t)(skeys[p - 1] >> 32) != (uint32_t)(skeys[p] >> 32))
            replay_run<KIND>(c, skeys, p, nc);
      }
    } else {
      // ---- general path: stable LSD radix sort by group id, 8 bits per pass -------------------------
      uint64_t* src = c.sortA;
      uint64_t* dst = c.sortB;
      const uint32_t n_st = (nc + kSortTile - 1) / kSortTile;
      uint32_t* s_hist = (uint32_t*)skeys;                       // [256]
      uint32_t* s_wcnt = (uint32_t*)skeys + 256;                 // [8 warps][256]
      for (uint32_t pass = 0; pass < c.sort_passes; pass++) {
        const uint32_t shift = 32 + 8 * pass;
        // (a) per-tile digit histograms
        for (uint32_t T = blockIdx.x; T < n_st; T += gridDim.x) {
          s_hist[tid] = 0;
          __syncthreads();
          uint32_t base = T * kSortTile;
#pragma unroll
          for (int k = 0; k < kSortItems; k++) {
            uint32_t i = base + k * kThreads + tid;
            if (i < nc) atomicAdd(&s_hist[(uint32_t)(src[i] >> shift) & 255u], 1u);
          }
          __syncthreads();
          c.ghist[(size_t)tid * n_st + T] = s_hist[tid];
          __syncthreads();
        }
        grid.sync();
        // (b) row totals
        for (uint32_t d = blockIdx.x; d < 256; d += gridDim.x) {
          uint32_t s = 0;
          for (uint32_t T = tid; T < n_st; T += kThreads) s += c.ghist[(size_t)d * n_st + T];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
          if (lane_id() == 0) wsum[warp_id()] = s;
          __syncthreads();
          if (tid == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < kThreads / 32; w++) tot += wsum[w];
            c.rowtot[d] = tot;
          }
          __syncthreads();
        }
        grid.sync();
        // (c) rows -> global exclusive offsets (digit-major, tile-minor)
        for (uint32_t d = blockIdx.x; d < 256; d += gridDim.x) {
          uint32_t part = (tid < d) ? c.rowtot[tid] : 0;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
          if (lane_id() == 0) wsum[warp_id()] = part;
          __syncthreads();
          if (tid == 0) {
            uint32_t b = 0;
            for (int w = 0; w < kThreads / 32; w++) b += wsum[w];
            s_carry = b;
          }
          __syncthreads();
          for (uint32_t base = 0; base < n_st; base += kThreads) {
            uint32_t T = base + tid;
            uint32_t v = T < n_st ? c.ghist[(size_t)d * n_st + T] : 0;
            uint32_t x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
              if ((int)lane_id() >= o) x += y;
            }
            if (lane_id() == 31) wsum[warp_id()] = x;
            __syncthreads();
            uint32_t woff = 0, tot = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; w++) {
              if (w < (int)warp_id()) woff += wsum[w];
              tot += wsum[w];
            }
            if (T < n_st) c.ghist[(size_t)d * n_st + T] = s_carry + woff + (x - v);
            __syncthreads();
            if (tid == 0) s_carry += tot;
            __syncthreads();
          }
        }
        grid.sync();
        // (d) stable scatter: warp w owns items [w*256, w*256+256) of the tile, in 8 rounds of 32
        for (uint32_t T = blockIdx.x; T < n_st; T += gridDim.x) {
          for (uint32_t i = tid; i < 8 * 256; i += kThreads) s_wcnt[i] = 0;
          __syncthreads();
          const uint32_t wbase = T * kSortTile + warp_id() * (kSortItems * 32);
          uint64_t key[kSortItems];
          uint32_t rank[kSortItems];
#pragma unroll
          for (int k = 0; k < kSortItems; k++) {
            uint32_t i = wbase + k * 32 + lane_id();
            bool ok = i < nc;
            key[k] = ok ? src[i] : 0;
            uint32_t dgt = ok ? ((uint32_t)(key[k] >> shift) & 255u) : 256u + lane_id();  // inactive lanes never match
            uint32_t peers = __match_any_sync(0xffffffffu, dgt);
            uint32_t before = __popc(peers & ((1u << lane_id()) - 1u));
            uint32_t basec = 0;
            if (ok) {
              uint32_t* cptr = &s_wcnt[warp_id() * 256 + dgt];
              if (before == 0) { basec = *cptr; *cptr = basec + __popc(peers); }
              basec = __shfl_sync(peers, basec, __ffs(peers) - 1);
            }
            rank[k] = basec + before;
            __syncwarp();
          }
          __syncthreads();
          // exclusive scan over warps for digit = tid
          {
            uint32_t run = c.ghist[(size_t)tid * n_st + T];
#pragma unroll
            for (int w = 0; w < kThreads / 32; w++) {
              uint32_t v = s_wcnt[w * 256 + tid];
              s_wcnt[w * 256 + tid] = run;
              run += v;
            }
          }
          __syncthreads();
#pragma unroll
          for (int k = 0; k < kSortItems; k++) {
            uint32_t i = wbase + k * 32 + lane_id();
            if (i < nc) {
              uint32_t dgt = (uint32_t)(key[k] >> shift) & 255u;
              dst[s_wcnt[warp_id() * 256 + dgt] + rank[k]] = key[k];
            }
          }
          __syncthreads();
        }
        grid.sync();
        uint64_t* tmp = src; src = dst; dst = tmp;
      }
      // ---- replay: one thread per same-group run ---------------------------------------------------
      for (uint32_t p = blockIdx.x * kThreads + tid; p < nc; p += gridDim.x * kThreads)
        if (p == 0 || (uint32_t)(src[p - 1] >> 32) != (uint32_t)(src[p] >> 32)) replay_run<KIND>(c, src, p, nc);
    }
    if (blockIdx.x == 0 && tid == 0) atomicAdd(&c.counters[1], (unsigned long long)nc);
  }

  // ---- clear the conflict bitmaps: only words this chunk touched -----------------------------------
  for (uint32_t i = blockIdx.x * kThreads + tid; i < c.n; i += gridDim.x * kThreads) {
    uint32_t g = c.grp[i];
    if (g != kNoGroup) {
      uint32_t w = g >> 5;
#pragma unroll
      for (int b = 0; b < 5; b++) c.bm[(size_t)b * c.bm_words + w] = 0;
    }
  }
}

}  // namespace dint
