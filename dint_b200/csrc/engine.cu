// engine.cu -- host side of libdint_b200.so: state allocation in HBM, the per-chunk launch sequence,
// host<->device pipelining for dint_submit(), state inspection, and the extern "C" ABI of
// include/dint_b200.h.  No CPU implementation of the request path exists here: without a CUDA
// device every compute entry point returns DINT_ENODEV.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>

#include "../../include/dint_b200.h"
#include "kernels.cuh"
#include "route.cuh"
#include "kv.cuh"

using namespace dint;

static thread_local std::string g_last_error;
static int set_err(int code, const char* what, cudaError_t ce = cudaSuccess) {
  char buf[512];
  if (ce != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(ce));
  else snprintf(buf, sizeof buf, "%s", what);
  g_last_error = buf;
  return code;
}
#define CU(call)                                                              \
  do {                                                                        \
    cudaError_t _e = (call);                                                  \
    if (_e != cudaSuccess) return set_err(_e == cudaErrorMemoryAllocation ? DINT_ENOMEM : DINT_EIO, #call, _e); \
  } while (0)

static const uint32_t kMsgSize[DINT_NUM_KINDS] = {6, 9, 53, 53, 55, 23};
static const uint32_t kLogEntry[DINT_NUM_KINDS] = {0, 0, 56, 0, 64, 32};
static const uint32_t kValSize[DINT_NUM_KINDS] = {0, 0, 0, 40, 40, 8};

enum { KT_CLASSIFY = 0, KT_LOGSCAN, KT_APPLY, KT_ORDERED, KT_LOAD, KT_NUM };
static const char* kKernelNames[KT_NUM] = {"k_classify", "k_log_scan", "k_apply", "k_ordered", "k_kv_load"};

struct EvPair { cudaEvent_t a, b; int which; };
constexpr int kHostBufs = 4;      // device staging buffers of the host path (dint_submit)

struct dint_engine {
  int kind = 0;
  int device = 0;
  dint_cfg cfg{};
  uint32_t msg = 0;
  uint32_t chunk = 0;
  uint32_t max_tiles = 0;
  bool has_log = false;
  Ctx ctx{};                       // device pointers + constants; per-launch fields filled per chunk
  std::vector<void*> allocs;       // everything to cudaFree
  cudaStream_t stream = nullptr, s_in = nullptr, s_out = nullptr;
  uint8_t* d_req[kHostBufs] = {nullptr};
  uint8_t* d_resp[kHostBufs] = {nullptr};
  cudaEvent_t ev_in[kHostBufs]{}, ev_comp[kHostBufs]{}, ev_out[kHostBufs]{};
  uint32_t host_chunk = 0;                   // requests per host-path slice
  bool pdl = false;                          // DINT_PDL=1: programmatic dependent launches for K1 / K2 (experimental)
  bool plain_launches = false;               // inside the multi-GPU step: no cooperative launches (see GridBar)
  uint32_t host_min_slice = 0;               // smallest slice of the pyramid a host-path call is cut into
  bool host_ramp_up = true;
  unsigned long long* h_counters = nullptr;  // pinned mirror of ctx.counters (host path reads it without a blocking copy)
  int coop_grid = 0;
  int sms = 0;
  int grid_classify = 0, grid_apply = 0;     // persistent CTAs (SMs x resident CTAs per SM)
  uint32_t smem_stage = 0;                   // dynamic shared memory of K1/K2: kStages staged tiles
  uint64_t total_groups = 0;
  uint32_t* d_flags[2] = {nullptr, nullptr}; // flag-nibble sets, alternating per chunk
  uint32_t* d_grp[2] = {nullptr, nullptr};   // group ids of the current / previous chunk
  uint64_t chunk_seq = 0;
  uint32_t prev_n = 0;                       // requests of the previous chunk whose flags are still set
  bool ord_pending = false;                  // the previous chunk's listed requests await their replay
  uint8_t* ord_resp = nullptr;               //   ... and live in this reply array
  uint32_t smem_classify = 0;                // K1: max(stages, ordered-replay slices)
  uint32_t* d_nc = nullptr;                  // [2 chunks][2]: listed / overflow counters
  uint32_t* d_route = nullptr;               // multi-GPU dispatch scratch (per-tile per-shard counts)
  uint32_t route_tiles = 0;
  uint32_t* d_route2 = nullptr;              // dispatch scratch: [grid][8] per-CTA counts + the finished-CTA counter
  int grid_route = 0;                        // co-resident CTAs of k_route_dispatch
  // L2 persistence: the flag sets (+ lock_fasst lock bits) live in one arena that every launch maps
  // with a persisting access-policy window, so the streaming request/reply traffic cannot evict it
  uint8_t* hot_arena = nullptr;
  size_t hot_bytes = 0;
  bool use_window = false;
  cudaAccessPolicyWindow window{};
  // stats
  dint_stats stats{};
  // profiling
  uint32_t profiling = 0;          // bit k set: bracket launches of kernel k (KT_*) with CUDA events
  std::vector<EvPair> ev_pool;
  size_t ev_used = 0;
  double kt_ms[KT_NUM] = {0};
  uint64_t kt_n[KT_NUM] = {0};
  // KV host mirrors
  KvHost kv[kMaxTables];
};

template <typename T>
static int dalloc(dint_engine* e, T** p, size_t count, bool zero = true) {
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  void* q = nullptr;
  CU(cudaMalloc(&q, bytes));
  e->allocs.push_back(q);
  if (zero) CU(cudaMemsetAsync(q, 0, bytes, e->stream));
  *p = (T*)q;
  return DINT_OK;
}

// ---- profiling helpers ------------------------------------------------------------------------------
static int prof_flush(dint_engine* e) {
  for (size_t i = 0; i < e->ev_used; i++) {
    float ms = 0;
    CU(cudaEventSynchronize(e->ev_pool[i].b));
    CU(cudaEventElapsedTime(&ms, e->ev_pool[i].a, e->ev_pool[i].b));
    e->kt_ms[e->ev_pool[i].which] += ms;
    e->kt_n[e->ev_pool[i].which]++;
  }
  e->ev_used = 0;
  return DINT_OK;
}
struct ProfScope {
  dint_engine* e; cudaStream_t s; EvPair* p = nullptr;
  ProfScope(dint_engine* e_, cudaStream_t s_, int which) : e(e_), s(s_) {
    e->stats.kernel_launches++;
    if (!((e->profiling >> which) & 1u)) return;
    if (e->ev_used == e->ev_pool.size()) {
      if (e->ev_pool.size() >= 8192) { prof_flush(e); }
      else {
        EvPair np{}; cudaEventCreate(&np.a); cudaEventCreate(&np.b);
        e->ev_pool.push_back(np);
      }
    }
    p = &e->ev_pool[e->ev_used++];
    p->which = which;
    cudaEventRecord(p->a, s);
  }
  ~ProfScope() { if (p) cudaEventRecord(p->b, s); }
};

// ---- per-chunk launch sequence ------------------------------------------------------------------------
template <typename... Args>
static cudaError_t launch_ex(dint_engine* e, void (*kern)(Args...), int grid, int block, size_t smem, cudaStream_t s,
                             bool coop, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid < 1 ? 1 : grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[3];
  int na = 0;
  if (coop) { at[na].id = cudaLaunchAttributeCooperative; at[na].val.cooperative = 1; na++; }
  else if (e->pdl) {                                    // DINT_PDL=1: the launch may overlap its predecessor's tail
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    na++;
  }
  if (e->use_window) { at[na].id = cudaLaunchAttributeAccessPolicyWindow; at[na].val.accessPolicyWindow = e->window; na++; }
  cfg.attrs = at;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

// One chunk: K1 (classify this chunk + replay the previous chunk's listed requests), K2 (apply), and the
// fallback launch that only does work when one of THIS chunk's buckets overflowed.  c.n == 0 = flush: K1
// alone, replaying the last chunk's listed requests.
template <int KIND, bool HAS_LOG>
static int launch_chunk_t(dint_engine* e, const Ctx& c, cudaStream_t s) {
  {
    ProfScope ps(e, s, KT_CLASSIFY);
    int want = (int)c.n_tiles, clr = (int)((c.prev_n + 4 * kTile - 1) / (4 * kTile));
    if (clr > want) want = clr;
    int grid = (want < e->grid_classify && !c.ord_pending) ? want : e->grid_classify;
    CU(launch_ex(e, k_classify<KIND, HAS_LOG>, grid, kTile, e->smem_classify, s, false, c));
  }
  if (c.n == 0) { CU(cudaGetLastError()); return DINT_OK; }
  if (HAS_LOG) {
    ProfScope ps(e, s, KT_LOGSCAN);
    k_log_scan<<<1, kThreads, 0, s>>>(c);
  }
  {
    ProfScope ps(e, s, KT_APPLY);
    int grid = (int)c.n_tiles < e->grid_apply ? (int)c.n_tiles : e->grid_apply;
    CU(launch_ex(e, k_apply<KIND, HAS_LOG>, grid, kTile, e->smem_stage, s, false, c));
  }
  if (KIND != K_LOG) {   // the log server has no per-key state: nothing to order
    ProfScope ps(e, s, KT_ORDERED);
    Ctx f = c;           // this chunk's own counters / replies
    f.nc_ord = c.nc_cur;
    f.ord_resp = c.resp;
    f.coop_launch = e->plain_launches ? 0u : 1u;
    int g3 = e->coop_grid;
    if (e->plain_launches && e->sms > 0) {             // leave room for the one-warp flag-polling kernels of the other streams
      const int per = g3 / e->sms;
      g3 = (per > 1 ? per - 1 : 1) * e->sms;
    }
    CU(launch_ex(e, k_ordered<KIND>, g3, kThreads, 0, s, !e->plain_launches, f));
  }
  CU(cudaGetLastError());
  return DINT_OK;
}

static int launch_chunk(dint_engine* e, const Ctx& c, cudaStream_t s) {
  switch (e->kind) {
    case DINT_LOCK2PL: return launch_chunk_t<K_LOCK2PL, false>(e, c, s);
    case DINT_FASST: return launch_chunk_t<K_FASST, false>(e, c, s);
    case DINT_LOG: return launch_chunk_t<K_LOG, true>(e, c, s);
    case DINT_STORE: return launch_chunk_t<K_STORE, false>(e, c, s);
    case DINT_TATP: return launch_chunk_t<K_TATP, true>(e, c, s);
    case DINT_SMALLBANK: return launch_chunk_t<K_SMALLBANK, true>(e, c, s);
  }
  return DINT_EINVAL;
}

template <int KIND, bool HAS_LOG>
static int grids_for(dint_engine* e) {
  int per_sm = 0, sms = 0;
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device));
  e->smem_stage = Stage<Wire<KIND>::MSG>::N * Stage<Wire<KIND>::MSG>::BYTES;
  if (e->smem_stage > 48 * 1024) {
    CU(cudaFuncSetAttribute(k_classify<KIND, HAS_LOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_stage));
    CU(cudaFuncSetAttribute(k_apply<KIND, HAS_LOG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)e->smem_stage));
  }
  { const char* pad = getenv("DINT_APPLY_SMEM_PAD"); if (pad) e->smem_stage += (uint32_t)atoi(pad); }   // occupancy experiments
  e->smem_classify = e->smem_stage;
  if (KIND != K_LOG && (kTile / 32) * OrdSlice<KIND>::BYTES > e->smem_classify) e->smem_classify = (kTile / 32) * OrdSlice<KIND>::BYTES;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_classify<KIND, HAS_LOG>, kTile, e->smem_classify));
  if (per_sm < 1) return set_err(DINT_EIO, "k_classify cannot be resident");
  e->grid_classify = per_sm * sms;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_apply<KIND, HAS_LOG>, kTile, e->smem_stage));
  if (per_sm < 1) return set_err(DINT_EIO, "k_apply cannot be resident");
  e->grid_apply = per_sm * sms;
  if (KIND == K_LOG) { e->coop_grid = 1; return DINT_OK; }
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ordered<KIND>, kThreads, 0));
  if (per_sm < 1) return set_err(DINT_EIO, "k_ordered cannot be resident");
  if (per_sm > 4) per_sm = 4;
  e->coop_grid = per_sm * sms;
  e->sms = sms;
  return DINT_OK;
}

static void fill_chunk_ctx(dint_engine* e, Ctx& c) {
  const int cur = (int)(e->chunk_seq & 1);
  c.grp = e->d_grp[cur];
  c.grp_prev = e->d_grp[cur ^ 1];
  c.flags = e->d_flags[cur];
  c.flags_prev = e->d_flags[cur ^ 1];
  c.prev_n = e->prev_n;
  c.nc_cur = e->d_nc + 4 * cur;          // {listed, overflow, a writer exists, -}
  c.nc_ord = e->d_nc + 4 * (cur ^ 1);
  c.ord_pending = e->ord_pending ? 1u : 0u;
  c.ord_resp = e->ord_resp;
}

// one chunk (n <= e->chunk); leaves its listed requests pending until the next chunk or flush_ordered()
static int submit_chunk(dint_engine* e, const uint8_t* req, uint32_t n, uint8_t* resp, cudaStream_t s) {
  Ctx c = e->ctx;
  c.n = n;
  c.n_tiles = (n + kTile - 1) / kTile;
  c.req = req;
  c.resp = resp;
  fill_chunk_ctx(e, c);
  int rc = launch_chunk(e, c, s);
  if (rc) return rc;
  e->chunk_seq++;
  e->prev_n = n;
  e->ord_pending = e->kind != DINT_LOG;
  e->ord_resp = resp;
  e->stats.chunks++;
  e->stats.requests += n;
  return DINT_OK;
}

// replays the last chunk's listed requests (and retires its flags); after this every reply is final
static int flush_ordered(dint_engine* e, cudaStream_t s) {
  if (!e->ord_pending) return DINT_OK;
  Ctx c = e->ctx;
  c.n = 0;
  c.n_tiles = 0;
  c.req = nullptr;
  c.resp = nullptr;
  fill_chunk_ctx(e, c);
  int rc = launch_chunk(e, c, s);
  if (rc) return rc;
  e->ord_pending = false;
  e->prev_n = 0;                 // the flush also cleared that chunk's flags
  return DINT_OK;
}

static int run_device(dint_engine* e, const uint8_t* req, uint64_t n, uint8_t* resp, cudaStream_t s) {
  for (uint64_t off = 0; off < n; off += e->chunk) {
    uint32_t cn = (uint32_t)((n - off < e->chunk) ? (n - off) : e->chunk);
    int rc = submit_chunk(e, req + off * e->msg, cn, resp + off * e->msg, s);
    if (rc) return rc;
  }
  return flush_ordered(e, s);
}

static int pull_counters(dint_engine* e) {
  unsigned long long h[4];
  CU(cudaMemcpy(h, e->ctx.counters, sizeof h, cudaMemcpyDeviceToHost));
  e->stats.errors = h[0];
  e->stats.conflicted = h[1];
  e->stats.max_run = h[2];
  return DINT_OK;
}

template <int MSG>
static void route_scatter_t(const uint8_t* rq, const uint8_t* ow, uint32_t n, uint32_t world, const uint32_t* tb,
                            const uint32_t* totals, uint8_t* out, uint32_t* perm, uint32_t tiles, cudaStream_t s) {
  k_exact_scatter<MSG><<<tiles, kThreads, 0, s>>>(rq, ow, n, world, tb, totals, out, perm);
}
template <int MSG>
static void route_unpermute_t(const uint8_t* sorted, const uint32_t* perm, uint32_t n, uint8_t* out, cudaStream_t s) {
  k_exact_unpermute<MSG><<<(n + kThreads - 1) / kThreads, kThreads, 0, s>>>(sorted, perm, n, out);
}

// Host-path slice sizes for a call of n requests (see dint_submit): slices double from `mn` up to the plateau
// `mx`, the body moves in plateau slices (a remainder first), and the end halves back down to `mn`.
struct HostSlices {
  uint32_t lvl[32];          // pyramid levels below the plateau, smallest first
  uint32_t n_lvl = 0, i_up = 0, i_down = 0;
  uint64_t body = 0, plateau = 0, pending = 0;
  bool ramp_up;
  HostSlices(uint64_t n, uint32_t mn, uint32_t mx, bool ramp_up_) : ramp_up(ramp_up_) {
    if (mn > mx) mn = mx;
    uint64_t used = 0;
    for (uint64_t s = mn; s < mx && n_lvl < 32; s <<= 1) {
      const uint64_t cost = ramp_up ? 2 * s : s;
      if (used + cost > n) break;
      lvl[n_lvl++] = (uint32_t)s;
      used += cost;
    }
    body = n - used;
    plateau = n_lvl ? (uint64_t)lvl[n_lvl - 1] * 2 : mx;
    if (plateau > mx) plateau = mx;
    i_down = n_lvl;
    if (!ramp_up) i_up = n_lvl;
  }
  uint32_t next() {          // 0 = done
    if (i_up < n_lvl) return lvl[i_up++];
    if (body) {
      const uint64_t rem = body % plateau;
      uint64_t c = rem ? rem : plateau;
      if (pending) { c = pending; pending = 0; }
      else if (rem && body > rem) {                        // a remainder is shared with one plateau slice: two
        c = (plateau + rem + 1) / 2;                       // mid-size slices instead of a tiny one and a full one
        pending = plateau + rem - c;
      }
      body -= c;
      return (uint32_t)c;
    }
    if (i_down) return lvl[--i_down];
    return 0;
  }
};

// ---- fused dispatch / combine (route.cuh) ---------------------------------------------------------------
template <int KIND>
static int route_dispatch_t(dint_engine* e, const RouteArgs& a, cudaStream_t s) {
  using RT = RTile<Wire<KIND>::MSG>;
  if (!e->d_route2) {
    { const char* g = getenv("DINT_ROUTE_GRID"); e->grid_route = g ? atoi(g) : 148 * 4; if (e->grid_route < 1) e->grid_route = 1; }
    const size_t words = 16 + (size_t)(e->grid_route / 32 + 2) * kMaxShards + (size_t)e->grid_route * kMaxShards;
    CU(cudaMalloc(&e->d_route2, words * sizeof(uint32_t)));
    CU(cudaMemsetAsync(e->d_route2, 0, words * sizeof(uint32_t), s));
  }
  RouteArgs b = a;
  b.done = e->d_route2;
  b.grp_tot = e->d_route2 + 16;
  b.cta_tot = e->d_route2 + 16 + (size_t)(e->grid_route / 32 + 2) * kMaxShards;
  int grid = (int)b.n_tiles < e->grid_route ? (int)b.n_tiles : e->grid_route;
  if (grid < 1) grid = 1;
  k_route_count<KIND><<<grid, kThreads, 0, s>>>(e->ctx, b);
  k_route_scatter<KIND><<<grid, kThreads, RT::SMEM, s>>>(b);
  CU(cudaGetLastError());
  return DINT_OK;
}
template <int MSG>
static int route_combine_t(dint_engine* e, const RouteArgs& a, cudaStream_t s) {
  using RT = RTile<MSG>;
  int grid = (int)a.n_tiles;
  if (grid > 148 * 16) grid = 148 * 16;
  k_route_combine<MSG><<<grid, kThreads, RT::SMEM, s>>>(a);
  CU(cudaGetLastError());
  return DINT_OK;
}
static uint32_t route_tile_records(const dint_engine* e) { return e->msg <= 12 ? kThreads * 4u : (uint32_t)kThreads; }

// ======================================================================================================
extern "C" {

uint32_t dint_msg_size(int kind) { return (kind >= 0 && kind < DINT_NUM_KINDS) ? kMsgSize[kind] : 0; }
uint32_t dint_log_entry_size(int kind) { return (kind >= 0 && kind < DINT_NUM_KINDS) ? kLogEntry[kind] : 0; }
const char* dint_last_error(void) { return g_last_error.c_str(); }
uint64_t dint_test_fasthash64(uint64_t x, int len) { return len == 4 ? fasthash64_u32((uint32_t)x) : fasthash64_u64(x); }
uint32_t dint_test_fastmod(uint64_t n, uint32_t d) { FastMod f = make_fastmod(d); return fast_mod(n, f); }
uint32_t dint_test_host_slices(uint64_t n, uint32_t min_slice, uint32_t max_slice, int ramp_up, uint32_t* out, uint32_t cap) {
  HostSlices sched(n, min_slice, max_slice, ramp_up != 0);
  uint32_t k = 0;
  for (uint32_t cn; (cn = sched.next()) != 0; k++)
    if (k < cap) out[k] = cn;
  return k;
}

void dint_default_cfg(int kind, dint_cfg* cfg) {
  memset(cfg, 0, sizeof *cfg);
  cfg->lock_slots = 36000000u;
  cfg->log_ring = 1000000u;
  cfg->subs_sizing = (kind == DINT_TATP) ? 7000000u : 2000000u;
  cfg->subs_populate = cfg->subs_sizing;
  cfg->accts_sizing = 24000000u;
  cfg->accts_populate = cfg->accts_sizing;
  cfg->n_shards = 1;
  cfg->shard_id = 0;
  cfg->chunk = 1u << 20;
}

void* dint_host_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
void dint_host_free(void* p) { if (p) cudaFreeHost(p); }

void dint_destroy(dint_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (void* p : e->allocs) cudaFree(p);
  if (e->d_route) cudaFree(e->d_route);
  if (e->d_route2) cudaFree(e->d_route2);
  if (e->h_counters) cudaFreeHost(e->h_counters);
  for (auto& ep : e->ev_pool) { cudaEventDestroy(ep.a); cudaEventDestroy(ep.b); }
  for (int i = 0; i < kHostBufs; i++) {
    if (e->ev_in[i]) cudaEventDestroy(e->ev_in[i]);
    if (e->ev_comp[i]) cudaEventDestroy(e->ev_comp[i]);
    if (e->ev_out[i]) cudaEventDestroy(e->ev_out[i]);
  }
  if (e->stream) cudaStreamDestroy(e->stream);
  if (e->s_in) cudaStreamDestroy(e->s_in);
  if (e->s_out) cudaStreamDestroy(e->s_out);
  delete e;
}

static int create_impl(dint_engine* e) {
  const dint_cfg& cf = e->cfg;
  Ctx& c = e->ctx;
  CU(cudaSetDevice(e->device));
  CU(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
  for (int i = 0; i < kHostBufs; i++) {
    CU(cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&e->ev_comp[i], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming));
  }
  c.n_shards = cf.n_shards;
  c.shard_id = cf.shard_id;
  c.shard_div = make_fastmod(cf.n_shards);
  c.slot_mod = make_fastmod(cf.lock_slots);
  c.ring_n = cf.log_ring ? cf.log_ring : 1;

  // ---- per-kind state in HBM ----
  uint64_t groups = 0;
  auto local_groups = [&](uint64_t global) { return (global + cf.n_shards - 1) / cf.n_shards; };
  int rc;
  switch (e->kind) {
    case DINT_LOCK2PL:
      groups = local_groups(cf.lock_slots);
      if ((rc = dalloc(e, &c.cnt2, groups))) return rc;
      break;
    case DINT_FASST:
      groups = local_groups(cf.lock_slots);
#ifdef DINT_VER16
      if ((rc = dalloc(e, &c.ver16, groups))) return rc;
      if ((rc = dalloc(e, &c.ver_hi, groups))) return rc;
#else
      if ((rc = dalloc(e, &c.ver, groups))) return rc;     // lock bits: in the hot arena, below
#endif
      break;
    case DINT_LOG:
      groups = 0;
      break;
    default:
      if ((rc = kv_create_tables(e->kind, cf, c, e->kv, &groups,
                                 [&](void** p, size_t bytes) -> int {
                                   uint8_t* q = nullptr;
                                   int r = dalloc(e, &q, bytes);
                                   *p = q;
                                   return r;
                                 })))
        return rc == DINT_EINVAL ? set_err(rc, "bad KV configuration") : rc;
      break;
  }
  e->total_groups = groups;
  if (groups >= 0xffffffffULL) return set_err(DINT_EINVAL, "too many groups");
  {
    uint32_t fl = 25;                                  // 2^25 nibbles = 16 MB per set: L2-resident
    while (fl > 10 && (1ULL << (fl - 1)) >= groups * 2 + 2048) fl--;   // tiny group spaces need less
    c.flags_mask = (1u << fl) - 1;
    const size_t set_bytes = (size_t)4 << (fl - 3);
    const size_t lock_bytes = (e->kind == DINT_FASST) ? (((groups + 31) / 32) * 4 + 255) / 256 * 256 : 0;
    e->hot_bytes = 2 * set_bytes + lock_bytes;
    if ((rc = dalloc(e, &e->hot_arena, e->hot_bytes))) return rc;
    e->d_flags[0] = (uint32_t*)e->hot_arena;
    e->d_flags[1] = (uint32_t*)(e->hot_arena + set_bytes);
    if (lock_bytes) c.lockbits = (uint32_t*)(e->hot_arena + 2 * set_bytes);
    // reserve L2 for it
    int max_persist = 0, max_win = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, e->device);
    cudaDeviceGetAttribute(&max_win, cudaDevAttrMaxAccessPolicyWindowSize, e->device);
    const char* env = getenv("DINT_L2_PERSIST");
    if ((!env || atoi(env) != 0) && max_persist > 0 && max_win > 0) {
      size_t want = e->hot_bytes < (size_t)max_persist ? e->hot_bytes : (size_t)max_persist;
      if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
        e->window.base_ptr = e->hot_arena;
        e->window.num_bytes = e->hot_bytes < (size_t)max_win ? e->hot_bytes : (size_t)max_win;
        e->window.hitRatio = (float)((double)want / (double)e->window.num_bytes > 1.0 ? 1.0 : (double)want / (double)e->window.num_bytes);
        e->window.hitProp = cudaAccessPropertyPersisting;
        e->window.missProp = cudaAccessPropertyStreaming;
        e->use_window = true;
      } else cudaGetLastError();
    }
    const char* g = getenv("DINT_L2_FETCH");
    if (g) { cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(g)); cudaGetLastError(); }
  }
  uint32_t bits = 1;
  while ((1ULL << bits) < groups) bits++;
  c.sort_passes = (bits + 7) / 8;

  if (e->has_log) {
    if ((rc = dalloc(e, &c.ring, (size_t)c.ring_n * kLogEntry[e->kind]))) return rc;
  }
  // ---- chunk scratch ----
  const uint32_t ch = e->chunk;
  e->max_tiles = (ch + kTile - 1) / kTile;
  for (int i = 0; i < 2; i++)
    if ((rc = dalloc(e, &e->d_grp[i], ch))) return rc;
  if ((rc = dalloc(e, &c.clist, (size_t)e->max_tiles * kTile))) return rc;
  if ((rc = dalloc(e, &c.ccnt, e->max_tiles))) return rc;
  if ((rc = dalloc(e, &c.cprefix, e->max_tiles + 1))) return rc;
  if ((rc = dalloc(e, &e->d_nc, 8))) return rc;
  {
    uint32_t lg = 0;
    while (((uint64_t)kBucketFill << lg) < ch) lg++;
    c.bucket_log2 = lg;
    if ((rc = dalloc(e, &c.buckets, ((size_t)1 << lg) * kBucketCap, false))) return rc;
    if ((rc = dalloc(e, &c.bcnt, (size_t)1 << lg))) return rc;
  }
  if ((rc = dalloc(e, &c.sortA, ch))) return rc;
  if ((rc = dalloc(e, &c.sortB, ch))) return rc;
  if ((rc = dalloc(e, &c.ghist, (size_t)256 * ((ch + kSortTile - 1) / kSortTile)))) return rc;
  if ((rc = dalloc(e, &c.rowtot, 256))) return rc;
  if ((rc = dalloc(e, &c.log_tilecnt, e->max_tiles))) return rc;
  if ((rc = dalloc(e, &c.log_tilebase, e->max_tiles))) return rc;
  if ((rc = dalloc(e, &c.log_total, 2))) return rc;
  if ((rc = dalloc(e, &c.counters, 4))) return rc;
  if ((rc = dalloc(e, &c.gbar, 4))) return rc;
#ifdef DINT_TILE_TICKETS
  if ((rc = dalloc(e, &c.tickets, 4))) return rc;
#endif

  switch (e->kind) {
    case DINT_LOCK2PL: rc = grids_for<K_LOCK2PL, false>(e); break;
    case DINT_FASST: rc = grids_for<K_FASST, false>(e); break;
    case DINT_LOG: rc = grids_for<K_LOG, true>(e); break;
    case DINT_STORE: rc = grids_for<K_STORE, false>(e); break;
    case DINT_TATP: rc = grids_for<K_TATP, true>(e); break;
    default: rc = grids_for<K_SMALLBANK, true>(e); break;
  }
  if (rc) return rc;
  int coop = 0;
  CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, e->device));
  if (!coop) return set_err(DINT_ENODEV, "device lacks cooperative launch");
  CU(cudaStreamSynchronize(e->stream));
  return DINT_OK;
}

int dint_create(int kind, const dint_cfg* cfg, int device, dint_engine** out) {
  if (!out || kind < 0 || kind >= DINT_NUM_KINDS) return set_err(DINT_EINVAL, "bad kind/out");
  *out = nullptr;
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return set_err(DINT_ENODEV, "no CUDA device: dint_b200 has no CPU fallback");
  }
  if (device < 0 || device >= ndev) return set_err(DINT_EINVAL, "bad device ordinal");
  dint_engine* e = new dint_engine();
  e->kind = kind;
  e->device = device;
  if (cfg) e->cfg = *cfg; else dint_default_cfg(kind, &e->cfg);
  dint_cfg& cf = e->cfg;
  if (cf.n_shards == 0) cf.n_shards = 1;
  if (cf.shard_id >= cf.n_shards || cf.lock_slots == 0) { delete e; return set_err(DINT_EINVAL, "bad shard/lock_slots"); }
  if (cf.chunk == 0) cf.chunk = 1u << 20;
  e->chunk = (cf.chunk + kTile - 1) / kTile * kTile;
  {
    const char* hc = getenv("DINT_HOST_CHUNK");
    uint32_t want = hc ? (uint32_t)atoi(hc) : (1u << 18);
    if (want < (uint32_t)kTile) want = kTile;
    e->host_chunk = want < e->chunk ? (want + kTile - 1) / kTile * kTile : e->chunk;
    { const char* pd = getenv("DINT_PDL"); e->pdl = pd && atoi(pd) == 1; }
    const char* ms = getenv("DINT_HOST_MIN_SLICE");
    e->host_min_slice = ms ? (uint32_t)atoi(ms) : 131072u;
    const char* ru = getenv("DINT_HOST_RAMP_UP");
    e->host_ramp_up = ru ? atoi(ru) != 0 : true;
    if (e->host_min_slice < (uint32_t)kTile) e->host_min_slice = kTile;
  }
  e->msg = kMsgSize[kind];
  e->has_log = kLogEntry[kind] != 0;
  int rc = create_impl(e);
  if (rc) { dint_destroy(e); return rc; }
  *out = e;
  return DINT_OK;
}

int dint_route_owner(dint_engine* e, const void* req_dev, uint64_t n, uint8_t* owner_dev, void* cuda_stream) {
  if (!e || (n && (!req_dev || !owner_dev)) || n > 0xffffffffULL) return set_err(DINT_EINVAL, "bad argument");
  if (n == 0) return DINT_OK;
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const uint32_t blocks = (uint32_t)((n + kThreads - 1) / kThreads);
  const uint8_t* rq = (const uint8_t*)req_dev;
  e->stats.kernel_launches++;
  switch (e->kind) {
    case DINT_LOCK2PL: k_route_owner<K_LOCK2PL><<<blocks, kThreads, 0, s>>>(e->ctx, rq, (uint32_t)n, owner_dev); break;
    case DINT_FASST: k_route_owner<K_FASST><<<blocks, kThreads, 0, s>>>(e->ctx, rq, (uint32_t)n, owner_dev); break;
    case DINT_LOG: k_route_owner<K_LOG><<<blocks, kThreads, 0, s>>>(e->ctx, rq, (uint32_t)n, owner_dev); break;
    case DINT_STORE: k_route_owner<K_STORE><<<blocks, kThreads, 0, s>>>(e->ctx, rq, (uint32_t)n, owner_dev); break;
    case DINT_TATP: k_route_owner<K_TATP><<<blocks, kThreads, 0, s>>>(e->ctx, rq, (uint32_t)n, owner_dev); break;
    default: k_route_owner<K_SMALLBANK><<<blocks, kThreads, 0, s>>>(e->ctx, rq, (uint32_t)n, owner_dev); break;
  }
  CU(cudaGetLastError());
  return DINT_OK;
}

int dint_route_partition(dint_engine* e, const void* req_dev, const uint8_t* owner_dev, uint64_t n, uint32_t n_shards,
                         void* sorted_dev, uint32_t* perm_dev, uint32_t* counts_dev, void* cuda_stream) {
  if (!e || n_shards == 0 || n_shards > kMaxShards || n > 0xffffffffULL) return set_err(DINT_EINVAL, "bad argument");
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const uint32_t tiles = (uint32_t)((n + kThreads - 1) / kThreads);
  if (tiles > e->route_tiles) {                          // scratch: per-tile per-shard counts + totals
    if (e->d_route) { CU(cudaFree(e->d_route)); e->d_route = nullptr; }
    e->route_tiles = tiles + tiles / 2 + 64;
    CU(cudaMalloc(&e->d_route, ((size_t)e->route_tiles * kMaxShards + 3 * kMaxShards) * sizeof(uint32_t)));
  }
  uint32_t* totals = e->d_route;                         // [0..8) records per shard
  uint32_t* tilecnt = e->d_route + 3 * kMaxShards;
  if (n == 0) { CU(cudaMemsetAsync(counts_dev, 0, n_shards * sizeof(uint32_t), s)); return DINT_OK; }
  e->stats.kernel_launches += 3;
  k_exact_count<<<tiles, kThreads, 0, s>>>(owner_dev, (uint32_t)n, n_shards, tilecnt);
  k_exact_scan<<<n_shards, kThreads, 0, s>>>(tilecnt, tiles, totals);
  const uint8_t* rq = (const uint8_t*)req_dev;
  uint8_t* out = (uint8_t*)sorted_dev;
  switch (e->msg) {
    case 6: route_scatter_t<6>(rq, owner_dev, (uint32_t)n, n_shards, tilecnt, totals, out, perm_dev, tiles, s); break;
    case 9: route_scatter_t<9>(rq, owner_dev, (uint32_t)n, n_shards, tilecnt, totals, out, perm_dev, tiles, s); break;
    case 23: route_scatter_t<23>(rq, owner_dev, (uint32_t)n, n_shards, tilecnt, totals, out, perm_dev, tiles, s); break;
    case 53: route_scatter_t<53>(rq, owner_dev, (uint32_t)n, n_shards, tilecnt, totals, out, perm_dev, tiles, s); break;
    default: route_scatter_t<55>(rq, owner_dev, (uint32_t)n, n_shards, tilecnt, totals, out, perm_dev, tiles, s); break;
  }
  CU(cudaMemcpyAsync(counts_dev, totals, n_shards * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  CU(cudaGetLastError());
  return DINT_OK;
}

uint32_t dint_route_tile_records(dint_engine* e) { return e ? route_tile_records(e) : 0; }

int dint_route_dispatch(dint_engine* e, const void* req_dev, const uint8_t* owner_in_dev, uint64_t n, uint32_t n_shards, uint32_t rank,
                        uint32_t cap, const dint_peer_ptrs* slab_ptrs, const dint_peer_ptrs* sig_ptrs, uint32_t epoch,
                        uint8_t* owner_dev, uint32_t* tilebase_dev, uint32_t* flags_dev, void* cuda_stream) {
  if (!e || !slab_ptrs || !flags_dev || n_shards == 0 || n_shards > kMaxShards || rank >= n_shards || cap == 0 || n > 0xffffffffULL ||
      (n && (!req_dev || !owner_dev || !tilebase_dev)))
    return set_err(DINT_EINVAL, "bad argument");
  if ((uintptr_t)req_dev & 15) return set_err(DINT_EINVAL, "device buffers must be 16-byte aligned");
  if (!owner_in_dev && n_shards != e->ctx.n_shards) return set_err(DINT_EINVAL, "owner computation needs n_shards == cfg.n_shards");
  CU(cudaSetDevice(e->device));
  RouteArgs a{};
  a.req = (const uint8_t*)req_dev;
  a.owner_in = owner_in_dev;
  a.owner = owner_dev;
  a.tilebase = tilebase_dev;
  a.flags = flags_dev;
  a.n = (uint32_t)n;
  a.n_tiles = (uint32_t)((n + route_tile_records(e) - 1) / route_tile_records(e));
  a.world = n_shards;
  a.me = rank;
  a.cap = cap;
  a.epoch = epoch;
  for (uint32_t i = 0; i < kMaxShards; i++) { a.slab.p[i] = slab_ptrs->p[i]; a.sig.p[i] = sig_ptrs ? sig_ptrs->p[i] : 0; }
  cudaStream_t s = (cudaStream_t)cuda_stream;
  e->stats.kernel_launches += 2;
  switch (e->kind) {
    case DINT_LOCK2PL: return route_dispatch_t<K_LOCK2PL>(e, a, s);
    case DINT_FASST: return route_dispatch_t<K_FASST>(e, a, s);
    case DINT_LOG: return route_dispatch_t<K_LOG>(e, a, s);
    case DINT_STORE: return route_dispatch_t<K_STORE>(e, a, s);
    case DINT_TATP: return route_dispatch_t<K_TATP>(e, a, s);
    default: return route_dispatch_t<K_SMALLBANK>(e, a, s);
  }
}

int dint_route_combine(dint_engine* e, const dint_peer_ptrs* reply_slab_ptrs, const uint8_t* owner_dev, const uint32_t* tilebase_dev,
                       uint64_t n, uint32_t n_shards, uint32_t cap, void* out_dev, void* cuda_stream) {
  if (!e || !reply_slab_ptrs || n_shards == 0 || n_shards > kMaxShards || cap == 0 || n > 0xffffffffULL ||
      (n && (!owner_dev || !tilebase_dev || !out_dev)))
    return set_err(DINT_EINVAL, "bad argument");
  if ((uintptr_t)out_dev & 15) return set_err(DINT_EINVAL, "device buffers must be 16-byte aligned");
  if (n == 0) return DINT_OK;
  CU(cudaSetDevice(e->device));
  RouteArgs a{};
  a.owner = (uint8_t*)owner_dev;
  a.tilebase = (uint32_t*)tilebase_dev;
  a.out = (uint8_t*)out_dev;
  a.n = (uint32_t)n;
  a.n_tiles = (uint32_t)((n + route_tile_records(e) - 1) / route_tile_records(e));
  a.world = n_shards;
  a.cap = cap;
  for (uint32_t i = 0; i < kMaxShards; i++) a.slab.p[i] = reply_slab_ptrs->p[i];
  cudaStream_t s = (cudaStream_t)cuda_stream;
  e->stats.kernel_launches++;
  switch (e->msg) {
    case 6: return route_combine_t<6>(e, a, s);
    case 9: return route_combine_t<9>(e, a, s);
    case 23: return route_combine_t<23>(e, a, s);
    case 53: return route_combine_t<53>(e, a, s);
    default: return route_combine_t<55>(e, a, s);
  }
}

int dint_p2p_wait(dint_engine* e, const uint32_t* local_sig_dev, uint32_t n_shards, uint32_t epoch, uint32_t* flags_dev, void* cuda_stream) {
  if (!e || !local_sig_dev || n_shards == 0 || n_shards > kMaxShards) return set_err(DINT_EINVAL, "bad argument");
  CU(cudaSetDevice(e->device));
  e->stats.kernel_launches++;
  k_p2p_wait<<<1, 32, 0, (cudaStream_t)cuda_stream>>>(local_sig_dev, n_shards, epoch, flags_dev + 1);
  CU(cudaGetLastError());
  return DINT_OK;
}

int dint_p2p_signal(dint_engine* e, const dint_peer_ptrs* sig_ptrs, uint32_t n_shards, uint32_t rank, uint32_t epoch, void* cuda_stream) {
  if (!e || !sig_ptrs || n_shards == 0 || n_shards > kMaxShards || rank >= n_shards) return set_err(DINT_EINVAL, "bad argument");
  CU(cudaSetDevice(e->device));
  PeerPtrs sg{};
  for (uint32_t i = 0; i < kMaxShards; i++) sg.p[i] = sig_ptrs->p[i];
  e->stats.kernel_launches++;
  k_p2p_signal<<<1, 32, 0, (cudaStream_t)cuda_stream>>>(sg, n_shards, rank, epoch);
  CU(cudaGetLastError());
  return DINT_OK;
}

int dint_route_unpermute(dint_engine* e, const void* sorted_dev, const uint32_t* perm_dev, uint64_t n, void* out_dev, void* cuda_stream) {
  if (!e || n > 0xffffffffULL) return set_err(DINT_EINVAL, "bad argument");
  if (n == 0) return DINT_OK;
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)cuda_stream;
  e->stats.kernel_launches++;
  const uint8_t* in = (const uint8_t*)sorted_dev;
  uint8_t* out = (uint8_t*)out_dev;
  switch (e->msg) {
    case 6: route_unpermute_t<6>(in, perm_dev, (uint32_t)n, out, s); break;
    case 9: route_unpermute_t<9>(in, perm_dev, (uint32_t)n, out, s); break;
    case 23: route_unpermute_t<23>(in, perm_dev, (uint32_t)n, out, s); break;
    case 53: route_unpermute_t<53>(in, perm_dev, (uint32_t)n, out, s); break;
    default: route_unpermute_t<55>(in, perm_dev, (uint32_t)n, out, s); break;
  }
  CU(cudaGetLastError());
  return DINT_OK;
}

int dint_sync(dint_engine* e) {
  if (!e) return DINT_EINVAL;
  CU(cudaSetDevice(e->device));
  CU(cudaDeviceSynchronize());
  int rc = prof_flush(e);
  if (rc) return rc;
  unsigned long long before = e->stats.errors;
  if ((rc = pull_counters(e))) return rc;
  return e->stats.errors != before ? DINT_EPROTO : DINT_OK;
}

int dint_submit_device(dint_engine* e, const void* req_dev, uint64_t n, void* resp_dev, void* cuda_stream) {
  if (!e || (n && (!req_dev || !resp_dev))) return set_err(DINT_EINVAL, "null argument");
  if (((uintptr_t)req_dev | (uintptr_t)resp_dev) & 15) return set_err(DINT_EINVAL, "device buffers must be 16-byte aligned");
  CU(cudaSetDevice(e->device));
  cudaStream_t s = (cudaStream_t)cuda_stream;        // NULL = the legacy default stream, as in every CUDA API
  return run_device(e, (const uint8_t*)req_dev, n, (uint8_t*)resp_dev, s);
}

int dint_submit(dint_engine* e, const void* req, uint64_t n, void* resp) {
  if (!e || (n && (!req || !resp))) return set_err(DINT_EINVAL, "null argument");
  CU(cudaSetDevice(e->device));
  // the host path moves data in slices of `hchunk` requests through a ring of kHostBufs device buffers:
  // small slices keep the PCIe fill/drain bubbles short, the ring keeps both copy engines and the SMs busy
  const uint32_t hchunk = e->host_chunk;
  if (!e->d_req[0]) {
    for (int i = 0; i < kHostBufs; i++) {
      int rc;
      if ((rc = dalloc(e, &e->d_req[i], (size_t)hchunk * e->msg + 16, false))) return rc;
      if ((rc = dalloc(e, &e->d_resp[i], (size_t)hchunk * e->msg + 16, false))) return rc;
    }
  }
  static const bool trace = getenv("DINT_HOST_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  CU(cudaDeviceSynchronize());     // order after anything submitted on user streams
  const uint8_t* rq = (const uint8_t*)req;
  uint8_t* rs = (uint8_t*)resp;
  unsigned long long err_before = e->stats.errors;
  // DINT_HOST_ZEROCOPY=1 (experimental, not yet measured): no D2H stage -- the kernels store the replies straight
  // into `resp` when it is pinned, device-mapped host memory (dint_host_alloc): the tiles leave K2 as bulk stores
  // over PCIe while the next slice's H2D runs the other way, the replay patches its few records in place.
  static const bool zero_copy = getenv("DINT_HOST_ZEROCOPY") != nullptr && atoi(getenv("DINT_HOST_ZEROCOPY")) == 1;
  if (zero_copy && n && ((uintptr_t)resp & 15) == 0 && resp != req) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, resp) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
      uint8_t* rs_dev = (uint8_t*)at.devicePointer;
      uint64_t k = 0;
      for (uint64_t off = 0; off < n; off += hchunk, k++) {       // slices of hchunk (a multiple of 128 records): every
        const int b = (int)(k % kHostBufs);                        // slice of resp starts 16-byte aligned
        const uint64_t cn = (n - off < hchunk) ? (n - off) : hchunk;
        const size_t bytes = (size_t)cn * e->msg;
        if (k >= (uint64_t)kHostBufs) CU(cudaStreamWaitEvent(e->s_in, e->ev_comp[b], 0));      // slice k-kHostBufs replayed
        CU(cudaMemcpyAsync(e->d_req[b], rq + off * e->msg, bytes, cudaMemcpyHostToDevice, e->s_in));
        CU(cudaEventRecord(e->ev_in[b], e->s_in));
        CU(cudaStreamWaitEvent(e->stream, e->ev_in[b], 0));
        int rc = submit_chunk(e, e->d_req[b], (uint32_t)cn, rs_dev + off * e->msg, e->stream);
        if (rc) return rc;
        if (k >= 1) CU(cudaEventRecord(e->ev_comp[(int)((k - 1) % kHostBufs)], e->stream));     // slice k-1 is final now
        e->stats.h2d_bytes += bytes;
        e->stats.d2h_bytes += bytes;
      }
      int rc = flush_ordered(e, e->stream);
      if (rc) return rc;
      if (!e->h_counters) CU(cudaHostAlloc((void**)&e->h_counters, 4 * sizeof(unsigned long long), cudaHostAllocDefault));
      CU(cudaMemcpyAsync(e->h_counters, e->ctx.counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
      CU(cudaStreamSynchronize(e->stream));
      if ((rc = prof_flush(e))) return rc;
      e->stats.errors = e->h_counters[0];
      e->stats.conflicted = e->h_counters[1];
      e->stats.max_run = e->h_counters[2];
      return e->stats.errors != err_before ? DINT_EPROTO : DINT_OK;
    }
    cudaGetLastError();                                    // not device-mapped host memory: the copying path below
  }
  // three-stage pipeline: H2D (s_in) | kernels (stream) | D2H (s_out).  Slice k's replies are final only
  // after the launch that replays its listed requests -- K1 of slice k+1, or the flush after the last
  // slice -- so D2H(k) is ordered behind that.
  uint64_t k = 0;
  uint64_t prev_off = 0, prev_bytes = 0;
  auto copy_out_prev = [&](uint64_t kk) -> int {          // D2H of slice kk-1
    int pb = (int)((kk - 1) % kHostBufs);
    CU(cudaEventRecord(e->ev_comp[pb], e->stream));
    CU(cudaStreamWaitEvent(e->s_out, e->ev_comp[pb], 0));
    CU(cudaMemcpyAsync(rs + prev_off, e->d_resp[pb], prev_bytes, cudaMemcpyDeviceToHost, e->s_out));
    CU(cudaEventRecord(e->ev_out[pb], e->s_out));
    return DINT_OK;
  };
  // Slice schedule.  PCIe moves large copies better than small ones (B200 box, pinned, both directions busy:
  // 37 GB/s per direction at 2.4 MB, 46 at 9.4 MB, 50 at 38 MB: tools/pcie_probe.cu), but the first slice's
  // H2D and the last two slices' kernels + D2H overlap with nothing.  So a call is cut as a pyramid: slices
  // double from host_min_slice up to host_chunk, stay there, and halve back down at the end.  (Measured with
  // tools/e2e_probe.py: every schedule between 128K..1M slices lands within 5 % -- the copies, not the
  // kernels or the launches (11 us of CPU per slice), are the bound: 1M requests = 345 us vs 203 us for two
  // perfectly overlapped 9.4 MB copies.)
  HostSlices sched(n, e->host_min_slice, hchunk, e->host_ramp_up);
  uint64_t off = 0;
  for (uint64_t cn; (cn = sched.next()) != 0; k++) {
    int b = (int)(k % kHostBufs);
    size_t bytes = (size_t)cn * e->msg;
    if (k >= (uint64_t)kHostBufs) CU(cudaStreamWaitEvent(e->s_in, e->ev_comp[b], 0));   // slice k-kHostBufs no longer read
    CU(cudaMemcpyAsync(e->d_req[b], rq + off * e->msg, bytes, cudaMemcpyHostToDevice, e->s_in));
    CU(cudaEventRecord(e->ev_in[b], e->s_in));
    CU(cudaStreamWaitEvent(e->stream, e->ev_in[b], 0));
    if (k >= (uint64_t)kHostBufs) CU(cudaStreamWaitEvent(e->stream, e->ev_out[b], 0)); // its replies have left d_resp[b]
    int rc = submit_chunk(e, e->d_req[b], (uint32_t)cn, e->d_resp[b], e->stream);
    if (rc) return rc;
    if (k >= 1 && (rc = copy_out_prev(k))) return rc;                   // slice k-1 is final now
    prev_off = off * e->msg;
    prev_bytes = bytes;
    e->stats.h2d_bytes += bytes;
    e->stats.d2h_bytes += bytes;
    off += cn;
  }
  {
    int rc = flush_ordered(e, e->stream);
    if (rc) return rc;
    if (k >= 1 && (rc = copy_out_prev(k))) return rc;
  }
  if (!e->h_counters) CU(cudaHostAlloc((void**)&e->h_counters, 4 * sizeof(unsigned long long), cudaHostAllocDefault));
  CU(cudaMemcpyAsync(e->h_counters, e->ctx.counters, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
  const auto t_enq = std::chrono::steady_clock::now();
  CU(cudaStreamSynchronize(e->s_out));
  CU(cudaStreamSynchronize(e->stream));
  int rc = prof_flush(e);
  if (rc) return rc;
  if (trace) {
    const auto t_end = std::chrono::steady_clock::now();
    fprintf(stderr, "[dint_submit] n=%llu slices=%llu enqueue=%.1f us total=%.1f us\n", (unsigned long long)n, (unsigned long long)k,
            std::chrono::duration<double, std::micro>(t_enq - t_begin).count(),
            std::chrono::duration<double, std::micro>(t_end - t_begin).count());
  }
  e->stats.errors = e->h_counters[0];
  e->stats.conflicted = e->h_counters[1];
  e->stats.max_run = e->h_counters[2];
  return e->stats.errors != err_before ? DINT_EPROTO : DINT_OK;
}

int dint_lock_state(dint_engine* e, int table, uint32_t slot, uint32_t out[2]) {
  if (!e || !out) return DINT_EINVAL;
  CU(cudaSetDevice(e->device));
  CU(cudaDeviceSynchronize());
  out[0] = out[1] = 0;
  const Ctx& c = e->ctx;
  uint32_t g = slot;
  if (e->kind == DINT_TATP || e->kind == DINT_SMALLBANK) {
    if (table < 0 || table >= (int)c.n_tables) return DINT_EINVAL;
    if (slot % c.n_shards != c.shard_id) return DINT_EINVAL;
    g = c.tbl[table].grp_base + slot / c.n_shards;
  } else if (e->kind == DINT_LOCK2PL || e->kind == DINT_FASST) {
    if (slot % c.n_shards != c.shard_id) return DINT_EINVAL;
    g = slot / c.n_shards;
  } else return DINT_EINVAL;
  if (g >= e->total_groups) return DINT_EINVAL;
  if (e->kind == DINT_LOCK2PL || e->kind == DINT_SMALLBANK) {
    uint2 v;
    CU(cudaMemcpy(&v, c.cnt2 + g, sizeof v, cudaMemcpyDeviceToHost));
    out[0] = v.x; out[1] = v.y;
  } else {
    uint32_t w;
    CU(cudaMemcpy(&w, c.lockbits + (g >> 5), 4, cudaMemcpyDeviceToHost));
    out[0] = (w >> (g & 31)) & 1u;
#ifdef DINT_VER16
    if (e->kind == DINT_FASST) {
      uint16_t lo = 0;
      uint32_t hi = 0;
      CU(cudaMemcpy(&lo, c.ver16 + g, 2, cudaMemcpyDeviceToHost));
      if (lo & 0x8000u) CU(cudaMemcpy(&hi, c.ver_hi + g, 4, cudaMemcpyDeviceToHost));
      out[1] = (lo & 0x7fffu) | (hi << 15);
    }
#else
    if (e->kind == DINT_FASST) CU(cudaMemcpy(&out[1], c.ver + g, 4, cudaMemcpyDeviceToHost));
#endif
  }
  return DINT_OK;
}

uint32_t dint_lock_slot(dint_engine* e, int table, uint64_t k) {
  if (!e) return 0;
  const Ctx& c = e->ctx;
  if (e->kind == DINT_LOCK2PL || e->kind == DINT_FASST) return fast_mod(fasthash64_u32((uint32_t)k), c.slot_mod);
  if ((e->kind == DINT_TATP || e->kind == DINT_SMALLBANK) && table >= 0 && table < (int)c.n_tables)
    return fast_mod(fasthash64_u64(k), c.tbl[table].lock_mod);
  return 0;
}

int dint_dump_log(dint_engine* e, void* out, uint64_t* appended) {
  if (!e || !e->has_log) return DINT_EINVAL;
  CU(cudaSetDevice(e->device));
  CU(cudaDeviceSynchronize());
  if (out) CU(cudaMemcpy(out, e->ctx.ring, (size_t)e->ctx.ring_n * kLogEntry[e->kind], cudaMemcpyDeviceToHost));
  if (appended) {
    unsigned long long t[2];
    CU(cudaMemcpy(t, e->ctx.log_total, sizeof t, cudaMemcpyDeviceToHost));
    *appended = t[0];
  }
  return DINT_OK;
}

int dint_get_stats(dint_engine* e, dint_stats* s) {
  if (!e || !s) return DINT_EINVAL;
  CU(cudaSetDevice(e->device));
  CU(cudaDeviceSynchronize());
  int rc = pull_counters(e);
  if (rc) return rc;
  *s = e->stats;
  return DINT_OK;
}
void dint_reset_stats(dint_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  cudaMemset(e->ctx.counters, 0, 4 * sizeof(unsigned long long));
  e->stats = dint_stats{};
  for (int i = 0; i < KT_NUM; i++) { e->kt_ms[i] = 0; e->kt_n[i] = 0; }
}
int dint_profile(dint_engine* e, int enable) {
  if (!e) return DINT_EINVAL;
  CU(cudaSetDevice(e->device));
  CU(cudaDeviceSynchronize());
  int rc = prof_flush(e);
  e->profiling = enable < 0 ? 0u : (uint32_t)enable == 1u ? 0xffffffffu : (uint32_t)enable;   // 1 = all kernels; else a bit mask
  return rc;
}
int dint_kernel_times(dint_engine* e, dint_kernel_time* out, int max_entries) {
  if (!e || !out) return DINT_EINVAL;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  prof_flush(e);
  int k = 0;
  for (int i = 0; i < KT_NUM && k < max_entries; i++) {
    if (!e->kt_n[i]) continue;
    memset(&out[k], 0, sizeof out[k]);
    snprintf(out[k].name, sizeof out[k].name, "%s", kKernelNames[i]);
    out[k].launches = e->kt_n[i];
    out[k].total_ms = e->kt_ms[i];
    k++;
  }
  return k;
}

// ---- KV entry points (store / tatp / smallbank) -------------------------------------------------------
int dint_load(dint_engine* e, int table, const uint64_t* keys, const void* vals, uint64_t n) {
  if (!e || table < 0 || table >= (int)e->ctx.n_tables || (n && (!keys || !vals))) return set_err(DINT_EINVAL, "bad table/arguments");
  CU(cudaSetDevice(e->device));
  const uint32_t vs = kValSize[e->kind];
  const uint64_t batch = 1u << 20;
  uint64_t* dk = nullptr;
  uint8_t* dv = nullptr;
  CU(cudaMalloc(&dk, batch * 8));
  cudaError_t ce = cudaMalloc(&dv, batch * vs);
  if (ce != cudaSuccess) { cudaFree(dk); return set_err(DINT_ENOMEM, "cudaMalloc", ce); }
  int rc = DINT_OK;
  for (uint64_t off = 0; off < n && rc == DINT_OK; off += batch) {
    uint64_t m = (n - off < batch) ? (n - off) : batch;
    if (cudaMemcpyAsync(dk, keys + off, m * 8, cudaMemcpyHostToDevice, e->stream) != cudaSuccess ||
        cudaMemcpyAsync(dv, (const uint8_t*)vals + off * vs, m * vs, cudaMemcpyHostToDevice, e->stream) != cudaSuccess) {
      rc = set_err(DINT_EIO, "load copy", cudaGetLastError());
      break;
    }
    {
      ProfScope ps(e, e->stream, KT_LOAD);
      kv_launch_load(e->kind, e->ctx, table, dk, dv, (uint32_t)m, e->stream);
    }
    if (cudaStreamSynchronize(e->stream) != cudaSuccess) rc = set_err(DINT_EIO, "k_kv_load", cudaGetLastError());
  }
  cudaFree(dk);
  cudaFree(dv);
  if (rc == DINT_OK) {
    int r2 = pull_counters(e);
    if (r2) return r2;
    if (e->stats.errors) return set_err(DINT_ENOMEM, "KV table full during load");
  }
  return rc;
}

int dint_populate(dint_engine* e) {
  if (!e) return DINT_EINVAL;
  if (e->ctx.n_tables == 0) return DINT_OK;       // lock / log servers start from zeroed arrays
  return kv_populate(e->kind, e->cfg, [&](int table, const uint64_t* k, const void* v, uint64_t n) {
    return dint_load(e, table, k, v, n);
  });
}

int dint_kv_get(dint_engine* e, int table, uint64_t key, void* val, uint32_t* ver) {
  if (!e || table < 0 || table >= (int)e->ctx.n_tables) return DINT_EINVAL;
  CU(cudaSetDevice(e->device));
  CU(cudaDeviceSynchronize());
  return kv_host_get(e->ctx, table, key, kValSize[e->kind], val, ver);
}

int64_t dint_kv_count(dint_engine* e, int table) {
  if (!e || table < 0 || table >= (int)e->ctx.n_tables) return DINT_EINVAL;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  unsigned long long v = 0;
  if (cudaMemcpy(&v, e->ctx.tbl[table].live, 8, cudaMemcpyDeviceToHost) != cudaSuccess) return DINT_EIO;
  return (int64_t)v;
}

// ---- the whole sharded step over NVLink peer memory, driven from one host call -----------------------------
// Three streams per rank: `side` partitions batch j+1 into the owners' inboxes while the caller's stream runs
// the engine on batch j and `ret` pulls the replies of batch j-1 out of the owners' outboxes.  The only
// cross-GPU synchronisation is three arrays of epoch words per rank (requests written / replies written /
// replies read), n_sets buffer sets deep.
struct dint_shard_ctx {
  dint_engine* e = nullptr;
  uint32_t W = 0, me = 0, cap = 0, S = 0;
  uint64_t inbox[4][kMaxShards]{}, outbox[4][kMaxShards]{}, retbox[4][kMaxShards]{};
  bool push = false;                       // DINT_SHARD_PUSH=1 and return buffers given: owners push the replies
  cudaEvent_t ev_eng[4]{};
  PeerPtrs sigreq{}, sigrsp{}, sigdone{};
  uint32_t *my_req = nullptr, *my_rsp = nullptr, *my_done = nullptr;
  uint32_t epoch = 0;
  cudaStream_t side = nullptr, ret = nullptr;
  cudaEvent_t ev_disp[4]{}, ev_comb[4]{}, ev_fork = nullptr;
  uint8_t* owner[4]{};
  uint32_t* tilebase[4]{};
  uint32_t* flags = nullptr;
  uint64_t max_n = 0;
  bool three_streams = true;               // DINT_SHARD_STREAMS=1: everything on the caller's stream
  bool trace = false;                      // DINT_SHARD_TRACE: per-phase CUDA-event timing, printed at destroy
  std::vector<cudaEvent_t> tev;            // 9 events per batch
  double tsum[8]{};
  uint64_t tcount = 0;
};

int dint_shard_create(dint_engine* e, uint32_t n_shards, uint32_t rank, uint32_t cap, uint32_t n_sets, const dint_peer_ptrs* inbox_sets,
                      const dint_peer_ptrs* outbox_sets, const dint_peer_ptrs* retbox_sets, const dint_peer_ptrs* sig_blocks, uint64_t max_n,
                      dint_shard_ctx** out) {
  if (!e || !out || !inbox_sets || !outbox_sets || !sig_blocks || n_shards == 0 || n_shards > kMaxShards || rank >= n_shards || cap == 0 ||
      n_sets < 2 || n_sets > 4 || max_n == 0 || max_n > 0xffffffffULL)
    return set_err(DINT_EINVAL, "bad argument");
  CU(cudaSetDevice(e->device));
  dint_shard_ctx* c = new dint_shard_ctx();
  e->plain_launches = true;
  c->e = e; c->W = n_shards; c->me = rank; c->cap = cap; c->S = n_sets; c->max_n = max_n;
  for (uint32_t s = 0; s < n_sets; s++)
    for (uint32_t o = 0; o < n_shards; o++) {
      c->inbox[s][o] = inbox_sets[s].p[o];
      c->outbox[s][o] = outbox_sets[s].p[o];
      c->retbox[s][o] = retbox_sets ? retbox_sets[s].p[o] : 0;
    }
  for (uint32_t o = 0; o < n_shards; o++) {
    c->sigreq.p[o] = sig_blocks->p[o];
    c->sigrsp.p[o] = sig_blocks->p[o] + 64;
    c->sigdone.p[o] = sig_blocks->p[o] + 128;
  }
  c->my_req = (uint32_t*)sig_blocks->p[rank];
  c->my_rsp = (uint32_t*)(sig_blocks->p[rank] + 64);
  c->my_done = (uint32_t*)(sig_blocks->p[rank] + 128);
  CU(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&c->ret, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  const uint32_t tr = route_tile_records(e);
  const size_t tiles = (size_t)((max_n + tr - 1) / tr);
  for (uint32_t s = 0; s < n_sets; s++) {
    CU(cudaEventCreateWithFlags(&c->ev_disp[s], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_comb[s], cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_eng[s], cudaEventDisableTiming));
    CU(cudaMalloc(&c->owner[s], max_n + 16));
    CU(cudaMalloc(&c->tilebase[s], tiles * kMaxShards * sizeof(uint32_t)));
  }
  CU(cudaMalloc(&c->flags, 2 * sizeof(uint32_t)));
  CU(cudaMemset(c->flags, 0, 2 * sizeof(uint32_t)));
  c->trace = getenv("DINT_SHARD_TRACE") != nullptr;
  { const char* ps = getenv("DINT_SHARD_PUSH"); c->push = retbox_sets && ps && atoi(ps) == 1 && ((size_t)cap * e->msg) % 16 == 0; }
  { const char* ts = getenv("DINT_SHARD_STREAMS"); c->three_streams = !ts || atoi(ts) == 3; }
  *out = c;
  return DINT_OK;
}

void dint_shard_destroy(dint_shard_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->e->device);
  cudaDeviceSynchronize();
  if (c->trace && c->tcount) {
    static const char* names[8] = {"side: wait inbox free", "side: dispatch", "main: wait requests", "main: engine + signal",
                                   "ret: wait replies", "ret: combine + signal", "batch: dispatch start -> combine end", "main: total"};
    fprintf(stderr, "[dint_shard rank %u] %llu batches, us per batch:", c->me, (unsigned long long)c->tcount);
    for (int q = 0; q < 8; q++) fprintf(stderr, " | %s %.1f", names[q], c->tsum[q] * 1e3 / (double)c->tcount);
    fprintf(stderr, "\n");
  }
  for (cudaEvent_t ev : c->tev) cudaEventDestroy(ev);
  for (uint32_t s = 0; s < c->S; s++) {
    if (c->ev_disp[s]) cudaEventDestroy(c->ev_disp[s]);
    if (c->ev_comb[s]) cudaEventDestroy(c->ev_comb[s]);
    if (c->ev_eng[s]) cudaEventDestroy(c->ev_eng[s]);
    if (c->owner[s]) cudaFree(c->owner[s]);
    if (c->tilebase[s]) cudaFree(c->tilebase[s]);
  }
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->flags) cudaFree(c->flags);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->ret) cudaStreamDestroy(c->ret);
  delete c;
}

int dint_shard_flags(dint_shard_ctx* c, uint32_t out[2]) {
  if (!c || !out) return DINT_EINVAL;
  CU(cudaSetDevice(c->e->device));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(out, c->flags, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  CU(cudaMemset(c->flags, 0, 2 * sizeof(uint32_t)));
  return DINT_OK;
}

int dint_shard_submit_many(dint_shard_ctx* c, uint32_t k, const void* const* req_dev, const uint8_t* const* dst_dev, uint64_t n,
                           void* const* out_dev, void* cuda_stream) {
  if (!c || !req_dev || !out_dev || n == 0 || n > c->max_n) return set_err(DINT_EINVAL, "bad argument");
  dint_engine* e = c->e;
  CU(cudaSetDevice(e->device));
  cudaStream_t main = (cudaStream_t)cuda_stream;
  // Default: dispatch, engine and combine on three streams.  DINT_SHARD_STREAMS=1 issues everything on the
  // caller's stream in the order D(j+1) E(j) C(j-1) (each cross-GPU wait then has a whole batch of slack and the
  // SM-filling kernels never compete); measured on 2 GPUs: 177 us per 2^20-request batch against 148 us for the
  // three streams, although the kernels of the three streams slow each other down (engine 114 us vs 86).
  const bool one_stream = !c->three_streams;
  cudaStream_t side = one_stream ? main : c->side, ret = one_stream ? main : c->ret;
  const uint32_t lag = one_stream ? 1u : 0u;
  const uint32_t W = c->W, S = c->S;
  const size_t slab = (size_t)c->cap * e->msg;
  if (!one_stream) {
    CU(cudaEventRecord(c->ev_fork, main));
    CU(cudaStreamWaitEvent(side, c->ev_fork, 0));
    CU(cudaStreamWaitEvent(ret, c->ev_fork, 0));
  }
  if (c->trace && c->tev.size() < (size_t)9 * k) {
    const size_t old = c->tev.size();
    c->tev.resize((size_t)9 * k);
    for (size_t i = old; i < c->tev.size(); i++) CU(cudaEventCreate(&c->tev[i]));
  }
  auto mark = [&](uint32_t j, int which, cudaStream_t st) { if (c->trace) cudaEventRecord(c->tev[(size_t)9 * j + which], st); };
  auto dispatch = [&](uint32_t j, uint32_t ep) -> int {
    const uint32_t s = ep % S;
    mark(j, 0, side);
    if (ep > S && !one_stream) {                                                // (one stream: both hold by program order)
      k_p2p_wait<<<1, 32, 0, side>>>(c->my_rsp, W, ep - S, c->flags + 1);       // every owner has consumed inbox set s
      CU(cudaStreamWaitEvent(side, c->ev_comb[s], 0));                          // and my combine is done with its state
    }
    mark(j, 1, side);
    dint_peer_ptrs in{}, sg{};
    for (uint32_t o = 0; o < W; o++) { in.p[o] = c->inbox[s][o] + (uint64_t)c->me * slab; sg.p[o] = c->sigreq.p[o]; }
    int rc = dint_route_dispatch(e, req_dev[j], dst_dev ? dst_dev[j] : nullptr, n, W, c->me, c->cap, &in, &sg, ep, c->owner[s],
                                 c->tilebase[s], c->flags, side);
    if (rc) return rc;
    mark(j, 2, side);
    if (!one_stream) CU(cudaEventRecord(c->ev_disp[s], side));
    return DINT_OK;
  };
  auto combine = [&](uint32_t j, uint32_t ep) -> int {     // the replies of batch j come home
    const uint32_t s = ep % S;
    if (!one_stream) CU(cudaStreamWaitEvent(ret, c->ev_disp[s], 0));
    mark(j, 6, ret);
    if (c->push) {                                         // my replies to the other sources go out first, then the flag
      if (!one_stream) CU(cudaStreamWaitEvent(ret, c->ev_eng[s], 0));
      if (W > 1) {
        PeerPtrs rb{};
        for (uint32_t o = 0; o < W; o++) rb.p[o] = c->retbox[s][o];
        const uint32_t slab16 = (uint32_t)(slab / 16);
        k_push_slabs<<<148 * 2, kThreads, 0, ret>>>((const uint8_t*)c->outbox[s][c->me], rb, W, c->me, slab16);
      }
      k_p2p_signal<<<1, 32, 0, ret>>>(c->sigrsp, W, c->me, ep);
    }
    k_p2p_wait<<<1, 32, 0, ret>>>(c->my_rsp, W, ep, c->flags + 1);
    mark(j, 7, ret);
    dint_peer_ptrs ob{};
    for (uint32_t o = 0; o < W; o++)
      ob.p[o] = c->push ? (o == c->me ? c->outbox[s][c->me] + (uint64_t)c->me * slab : c->retbox[s][c->me] + (uint64_t)o * slab)
                        : c->outbox[s][o] + (uint64_t)c->me * slab;
    int rc = dint_route_combine(e, &ob, c->owner[s], c->tilebase[s], n, W, c->cap, out_dev[j], ret);
    if (rc) return rc;
    k_p2p_signal<<<1, 32, 0, ret>>>(c->sigdone, W, c->me, ep);
    mark(j, 8, ret);
    if (!one_stream) CU(cudaEventRecord(c->ev_comb[s], ret));
    return DINT_OK;
  };
  uint32_t ep0 = c->epoch;
  int rc = dispatch(0, ep0 + 1);
  if (rc) return rc;
  for (uint32_t j = 0; j < k; j++) {
    const uint32_t ep = ep0 + 1 + j, s = ep % S;
    if (j + 1 < k && (rc = dispatch(j + 1, ep + 1))) return rc;
    // the engine sees the batches in order
    mark(j, 3, main);
    k_p2p_wait<<<1, 32, 0, main>>>(c->my_req, W, ep, c->flags + 1);                    // every source's slab has arrived
    if (ep > S) k_p2p_wait<<<1, 32, 0, main>>>(c->my_done, W, ep - S, c->flags + 1);   // outbox set s has been read
    mark(j, 4, main);
    rc = run_device(e, (const uint8_t*)c->inbox[s][c->me], (uint64_t)W * c->cap, (uint8_t*)c->outbox[s][c->me], main);
    if (rc) return rc;
    if (!c->push) k_p2p_signal<<<1, 32, 0, main>>>(c->sigrsp, W, c->me, ep);
    else if (!one_stream) CU(cudaEventRecord(c->ev_eng[s], main));
    mark(j, 5, main);
    if (j >= lag && (rc = combine(j - lag, ep - lag))) return rc;
    e->stats.kernel_launches += 5;
  }
  for (uint32_t j = k - (lag < k ? lag : k); j < k; j++)
    if ((rc = combine(j, ep0 + 1 + j))) return rc;
  c->epoch = ep0 + k;
  if (!one_stream) {
    for (uint32_t s = 0; s < S; s++) CU(cudaStreamWaitEvent(main, c->ev_comb[s], 0));   // join
    CU(cudaStreamWaitEvent(main, c->ev_disp[(ep0 + k) % S], 0));
  }
  CU(cudaGetLastError());
  if (c->trace) {                                        // diagnostic mode: synchronises
    CU(cudaStreamSynchronize(main));
    static const int pairs[8][2] = {{0, 1}, {1, 2}, {3, 4}, {4, 5}, {6, 7}, {7, 8}, {0, 8}, {3, 5}};
    for (uint32_t j = 0; j < k; j++)
      for (int q = 0; q < 8; q++) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->tev[(size_t)9 * j + pairs[q][0]], c->tev[(size_t)9 * j + pairs[q][1]]) == cudaSuccess) c->tsum[q] += ms;
      }
    c->tcount += k;
  }
  return DINT_OK;
}

}  // extern "C"
