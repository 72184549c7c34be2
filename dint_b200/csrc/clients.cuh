// clients.cuh -- the lock_fasst closed-loop clients ON the GPU.
//
// The reference's clients are Caladan uthreads on other machines (lock_fasst/caladan/client.cc:183-280): read the
// read set, lock the write set, validate by re-reading, abort or commit, start the next transaction.  workloads.cc
// restates them as round-based state machines on host cores (one request outstanding per client and round); this is
// the same state machine, decision for decision and draw for draw (same per-client xorshift64* stream), as ONE
// kernel per round: a thread absorbs its client's reply of round r and emits the request of round r + 1 straight
// into the engine's device request buffer.  With it "committed txn/s" and the abort rate are produced live, for as
// long as one likes, instead of replayed from a host-recorded trace.  Parity: tests/test_gpu_clients.py compares
// every round's request stream and the final counters with workloads.cc driving the oracle.
//
// State layout (structure of arrays; a round touches a header word, one key and one version per client):
//   hdr[c]    u64  {phase:8, pos:8, nr:8, nw:8, lim:8, wmask:16}    wmask bit i = rk[i] is also written
//   rng[c]    u64  xorshift64* state
//   rk[i][c]  u32  read set, ascending (i < nr)          rv[i][c]  u32  version read for rk[i]
#pragma once
#include "common.cuh"

namespace dint {

enum : uint32_t { CPH_READ = 0, CPH_ACQ, CPH_VALIDATE, CPH_ABORT, CPH_COMMIT };

struct ClientCtx {
  uint32_t n_clients, n_keys, read_pct, zipf_n;     // zipf_n != 0: keys are Zipf ranks, cdf[zipf_n]
  const double* cdf;
  unsigned long long* hdr;
  unsigned long long* rng;
  uint32_t* rk;                                     // [10][n_clients]
  uint32_t* rv;                                     // [10][n_clients]
  unsigned long long* stats;                        // [0] requests [1] committed [2] validation aborts [3] lock rejects [4] rounds
};

#ifdef __CUDACC__
struct CRng {                                       // workloads.cc Rng
  unsigned long long s;
  DINT_D unsigned long long next() {
    s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
    return s * 0x2545F4914F6CDD1DULL;
  }
  DINT_D uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (unsigned long long)n) >> 32); }
  DINT_D double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};
DINT_HD unsigned long long crng_seed(unsigned long long seed) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return (z ^ (z >> 31)) | 1;
}

struct CHdr { uint32_t phase, pos, nr, nw, lim, wmask; };
DINT_D CHdr chdr_unpack(unsigned long long h) {
  return CHdr{(uint32_t)(h & 255u), (uint32_t)(h >> 8) & 255u, (uint32_t)(h >> 16) & 255u, (uint32_t)(h >> 24) & 255u,
              (uint32_t)(h >> 32) & 255u, (uint32_t)(h >> 40) & 0xffffu};
}
DINT_D unsigned long long chdr_pack(const CHdr& h) {
  return (unsigned long long)h.phase | ((unsigned long long)h.pos << 8) | ((unsigned long long)h.nr << 16) |
         ((unsigned long long)h.nw << 24) | ((unsigned long long)h.lim << 32) | ((unsigned long long)h.wmask << 40);
}
// index into rk[] of the k-th written key
DINT_D uint32_t nth_set_bit(uint32_t mask, uint32_t k) {
  for (uint32_t i = 0; i < k; i++) mask &= mask - 1;
  return (uint32_t)__ffs((int)mask) - 1;
}

DINT_D uint32_t client_draw_key(const ClientCtx& c, CRng& r) {
  if (!c.zipf_n) return r.below(c.n_keys);
  const double u = r.unit();                        // std::lower_bound over the cdf
  uint32_t lo = 0, hi = c.zipf_n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (c.cdf[mid] < u) lo = mid + 1; else hi = mid;
  }
  return lo < c.zipf_n - 1 ? lo : c.zipf_n - 1;
}

// lock_fasst/caladan/trace_init.sh:12-24 through workloads.cc new_txn(): 5-10 distinct ids, sorted, each also written
// with probability 1 - read_pct
DINT_D void client_new_txn(const ClientCtx& c, uint32_t id, CRng& r, CHdr& h) {
  uint32_t want = 5 + r.below(6);
  if (want > c.n_keys) want = c.n_keys;
  uint32_t k[10];
  uint32_t n = 0;
  while (n < want) {
    const uint32_t x = client_draw_key(c, r);
    bool dup = false;
    for (uint32_t i = 0; i < n; i++) dup |= (k[i] == x);
    if (!dup) k[n++] = x;
  }
  for (uint32_t i = 1; i < n; i++) {                // insertion sort (n <= 10)
    const uint32_t x = k[i];
    uint32_t j = i;
    while (j > 0 && k[j - 1] > x) { k[j] = k[j - 1]; j--; }
    k[j] = x;
  }
  uint32_t wmask = 0, nw = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (r.below(100) >= c.read_pct) { wmask |= 1u << i; nw++; }
    c.rk[(size_t)i * c.n_clients + id] = k[i];
  }
  h.nr = n; h.nw = nw; h.wmask = wmask; h.pos = 0; h.phase = CPH_READ; h.lim = 0;
}

// first round: every client starts a transaction and emits its first request
__global__ void __launch_bounds__(256) k_clients_init(const ClientCtx c, unsigned long long seed, uint8_t* req) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= c.n_clients) return;
  CRng r{crng_seed(seed * 0x100000001B3ULL + id)};
  CHdr h{};
  client_new_txn(c, id, r, h);
  c.rng[id] = r.s;
  c.hdr[id] = chdr_pack(h);
  uint8_t* m = req + (size_t)id * 9;
  m[0] = 0;                                         // kRead of rk[0]
  st_u32_unaligned(m + 1, c.rk[id]);
  st_u32_unaligned(m + 5, 0u);
}

// one round: absorb the reply (workloads.cc absorb(), lock_fasst/caladan/client.cc:183-280), emit the next request
__global__ void __launch_bounds__(256) k_clients_step(const ClientCtx c, const uint8_t* resp, uint8_t* req) {
  const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t committed = 0, vaborts = 0, rejects = 0;
  __shared__ uint32_t s_list[256];
  __shared__ uint32_t s_cnt;
  if (threadIdx.x == 0) s_cnt = 0;
  __syncthreads();
  if (id < c.n_clients) {
    CHdr h = chdr_unpack(c.hdr[id]);
    const uint8_t* a = resp + (size_t)id * 9;
    const uint32_t type = a[0], ver = ld_u32_unaligned(a + 5);
    bool fresh = false;
    switch (h.phase) {
      case CPH_READ:
        c.rv[(size_t)h.pos * c.n_clients + id] = ver;
        if (++h.pos == h.nr) { h.pos = 0; h.phase = h.nw ? CPH_ACQ : CPH_VALIDATE; }
        break;
      case CPH_ACQ:
        if (type == 5) {                                      // kGrantLock
          if (++h.pos == h.nw) { h.pos = 0; h.phase = CPH_VALIDATE; }
        } else {                                              // kRejectLock: abort the locks taken so far, restart
          rejects = 1;
          if (h.pos) { h.lim = h.pos; h.pos = 0; h.phase = CPH_ABORT; }
          else { h.pos = 0; h.phase = CPH_READ; }
        }
        break;
      case CPH_VALIDATE:
        if (ver != c.rv[(size_t)h.pos * c.n_clients + id]) {  // client.cc:209-212 roll back
          vaborts = 1;
          if (h.nw) { h.lim = h.nw; h.pos = 0; h.phase = CPH_ABORT; }
          else { h.pos = 0; h.phase = CPH_READ; }
        } else if (++h.pos == h.nr) {
          if (h.nw) { h.pos = 0; h.phase = CPH_COMMIT; }
          else { committed = 1; fresh = true; }
        }
        break;
      case CPH_ABORT:
        if (++h.pos == h.lim) { h.pos = 0; h.phase = CPH_READ; }
        break;
      default:                                                // CPH_COMMIT
        if (++h.pos == h.nw) { committed = 1; fresh = true; }
        break;
    }
    if (fresh) {                                          // (new transactions are drawn below, by converged warps)
      s_list[atomicAdd(&s_cnt, 1u)] = id;
    } else {
      c.hdr[id] = chdr_pack(h);
      // ---- emit (workloads.cc emit()) ----
      uint32_t t, lid;
      if (h.phase == CPH_READ || h.phase == CPH_VALIDATE) { t = 0; lid = c.rk[(size_t)h.pos * c.n_clients + id]; }
      else {
        t = h.phase == CPH_ACQ ? 1u : h.phase == CPH_ABORT ? 2u : 3u;
        lid = c.rk[(size_t)nth_set_bit(h.wmask, h.pos) * c.n_clients + id];
      }
      uint8_t* m = req + (size_t)id * 9;
      m[0] = (uint8_t)t;
      st_u32_unaligned(m + 1, lid);
      st_u32_unaligned(m + 5, 0u);
    }
  }
  // ---- clients that finished a transaction (about 1 in 25 per round) start the next one: drawing 5-10 distinct keys,
  //      sorting them and choosing the write set is ~600 instructions, so the few clients of a CTA that need it are
  //      compacted into its first warps instead of dragging every warp through the divergent path ----
  __syncthreads();
  if (threadIdx.x < s_cnt) {
    const uint32_t id2 = s_list[threadIdx.x];
    CRng r{c.rng[id2]};
    CHdr h2{};
    client_new_txn(c, id2, r, h2);
    c.rng[id2] = r.s;
    c.hdr[id2] = chdr_pack(h2);
    uint8_t* m = req + (size_t)id2 * 9;
    m[0] = 0;                                              // kRead of the first key
    st_u32_unaligned(m + 1, c.rk[id2]);
    st_u32_unaligned(m + 5, 0u);
  }
  // counters: one atomic per warp and counter
  const uint32_t all = 0xffffffffu;
  const uint32_t nc = __popc(__ballot_sync(all, committed)), nv = __popc(__ballot_sync(all, vaborts)), nj = __popc(__ballot_sync(all, rejects));
  if ((threadIdx.x & 31) == 0) {
    if (nc) atomicAdd(&c.stats[1], (unsigned long long)nc);
    if (nv) atomicAdd(&c.stats[2], (unsigned long long)nv);
    if (nj) atomicAdd(&c.stats[3], (unsigned long long)nj);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&c.stats[0], (unsigned long long)c.n_clients); atomicAdd(&c.stats[4], 1ULL); }
}
#endif  // __CUDACC__

}  // namespace dint
