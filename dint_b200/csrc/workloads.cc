// workloads.cc -- libdint_wl.so: the reference's closed-loop clients restated as round-based request
// generators (CPU C++, no CUDA).  A "round" = every logical client has exactly one request outstanding;
// dint_wl_next() emits the round's requests in client order (that order IS the flat trace order), the
// caller hands them to a server (the GPU engine, the oracle, ...) and feeds the replies back with
// dint_wl_feed(), which advances every client's transaction state machine.
//
// What is restated ("REF" family):
//   lock_fasst : trace shape lock_fasst/caladan/trace_init.sh:9-27 (5-10 distinct ids, reads sorted, each
//                also written with p = 1 - r_prop); protocol lock_fasst/caladan/client.cc:183-280 (read ->
//                acquire -> validate by re-read -> abort | commit; a rejected lock aborts the locks taken
//                so far and restarts the transaction).
//   lock_2pl   : trace shape lock_2pl/caladan/trace_init.sh:9-24 (5-10 distinct ids ascending, exclusive
//                with p = 1 - r_prop, release in reverse); protocol lock_2pl/caladan/client.cc:164-236.
//   log_server : log_server/caladan/trace_init.sh:15-19 (key U[0,7009999], ver U[0,127], 40 random bytes).
//   store      : store/caladan/client_udp.cc:135-147,199-201 (LCG seed 0xdeadbeef + client gid; NURand s_id,
//                sf_type, start_time); "contention" = 50 % kSet (real kSet as client_ebpf.cc does -- the UDP
//                client's TxnSet sends kRead by mistake, client_udp.cc:180).
// The reference's Python generators are unseeded; here every client owns a splitmix/xorshift stream
// derived from (seed, client id).  "HOT" family = BASELINE.json's literal shape: n_keys (4800) keys,
// Zipf(theta) popularity.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.cuh"   // fasthash64 / exact fast modulo, shared with the device code

extern "C" {
typedef struct dint_wl_cfg {
  uint32_t kind;        // enum dint_kind (0 lock_2pl, 1 lock_fasst, 2 log, 3 store)
  uint32_t n_clients;
  uint64_t seed;
  uint32_t n_keys;      // lock ids are drawn from [0, n_keys) (reference: 24,000,000); store HOT: hot-set size
  double zipf_theta;    // 0 = uniform (reference); > 0: rank-frequency ~ 1 / rank^theta
  uint32_t read_pct;    // r_prop * 100 (reference 80): a key is read-only / shared with this probability
  uint32_t set_pct;     // store: percent of kSet (0 = "parallel", 50 = "contention")
  uint32_t store_subscribers;  // store: kSubscriberNum of the key generator (reference 2,000,000)
  uint32_t store_hot;   // store: 1 = HOT family (keys = first n_keys of the population, Zipf)
  uint32_t reserved[5];
} dint_wl_cfg;
typedef struct dint_wl dint_wl;
}

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed = 1) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL;          // splitmix64 to spread nearby seeds
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    s = (z ^ (z >> 31)) | 1;
  }
  uint64_t next() {                                       // xorshift64*
    s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
    return s * 0x2545F4914F6CDD1DULL;
  }
  uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
  double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct Zipf {                                             // inverse-CDF sampling over ranks 0..n-1
  std::vector<double> cdf;
  void init(uint32_t n, double theta) {
    cdf.resize(n);
    double acc = 0;
    for (uint32_t k = 0; k < n; k++) { acc += 1.0 / std::pow((double)(k + 1), theta); cdf[k] = acc; }
    for (auto& c : cdf) c /= acc;
  }
  uint32_t draw(Rng& r) const {
    double u = r.unit();
    return (uint32_t)std::min<size_t>(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin(), cdf.size() - 1);
  }
};

enum { K_LOCK2PL = 0, K_FASST = 1, K_LOG = 2, K_STORE = 3 };
static const uint32_t kMsg[4] = {6, 9, 53, 53};

struct LockClient {            // lock_fasst and lock_2pl share the transaction shape
  Rng rng;
  uint32_t rk[10], rv[10], wk[10];
  uint8_t wtype[10];           // lock_2pl: lock type per key (rk[] order)
  uint8_t nr = 0, nw = 0, phase = 0, pos = 0, lim = 0;
  uint64_t lcg = 0;            // store: the reference's LCG state
};

enum { PH_READ = 0, PH_ACQ, PH_VALIDATE, PH_ABORT, PH_COMMIT, PH_RELEASE, PH_REL_ABORT };

static inline void put32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }
static inline void put64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }
static inline uint32_t get32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

}  // namespace

struct dint_wl {
  dint_wl_cfg cfg;
  std::vector<LockClient> cl;
  Zipf zipf;
  uint64_t st_requests = 0, st_committed = 0, st_validation_aborts = 0, st_lock_rejects = 0, st_not_exist = 0,
           st_rounds = 0;

  uint32_t draw_key(Rng& r) { return cfg.zipf_theta > 0 ? zipf.draw(r) : r.below(cfg.n_keys); }

  void new_txn(LockClient& c) {             // trace_init.sh:12-24 of lock_fasst / lock_2pl
    uint32_t want = 5 + c.rng.below(6);
    if (want > cfg.n_keys) want = cfg.n_keys;
    uint32_t n = 0;
    while (n < want) {                      // random.sample: distinct keys
      uint32_t k = draw_key(c.rng);
      bool dup = false;
      for (uint32_t i = 0; i < n; i++) dup |= (c.rk[i] == k);
      if (!dup) c.rk[n++] = k;
    }
    std::sort(c.rk, c.rk + n);
    c.nr = (uint8_t)n;
    c.nw = 0;
    for (uint32_t i = 0; i < n; i++) {
      bool w = c.rng.below(100) >= cfg.read_pct;
      c.wtype[i] = w ? 1 : 0;
      if (w) c.wk[c.nw++] = c.rk[i];
    }
    c.pos = 0;
    c.phase = (cfg.kind == K_FASST) ? PH_READ : PH_ACQ;
  }

  // ---- store key generator: store/caladan/client_udp.cc:135-147 + store/udp/tatp.h:31-42 --------------
  static uint32_t fastrand(uint64_t* seed) { *seed = *seed * 1103515245ULL + 12345ULL; return (uint32_t)(*seed >> 32); }
  uint32_t nurand(uint64_t* seed) {
    const uint32_t S = cfg.store_subscribers;
    return ((fastrand(seed) % S) | (fastrand(seed) & 1048575u)) % S;
  }

  void emit(LockClient& c, uint8_t* m) {
    switch (cfg.kind) {
      case K_FASST: {                       // {type@0, lid@1, ver@5}
        uint8_t type; uint32_t lid;
        if (c.phase == PH_READ || c.phase == PH_VALIDATE) { type = 0; lid = c.rk[c.pos]; }
        else if (c.phase == PH_ACQ) { type = 1; lid = c.wk[c.pos]; }
        else if (c.phase == PH_ABORT) { type = 2; lid = c.wk[c.pos]; }
        else { type = 3; lid = c.wk[c.pos]; }
        m[0] = type; put32(m + 1, lid); put32(m + 5, 0);
        break;
      }
      case K_LOCK2PL: {                     // {action@0, lid@1, type@5}
        m[0] = (c.phase == PH_ACQ) ? 0 : 1;
        put32(m + 1, c.rk[c.pos]);
        m[5] = c.wtype[c.pos];
        break;
      }
      case K_LOG: {                         // {type@0, key@1, val@9, ver@49}
        m[0] = 0;
        put64(m + 1, c.rng.below(7010000));
        for (int i = 0; i < 40; i += 8) put64(m + 9 + i, c.rng.next());
        put32(m + 49, c.rng.below(128));
        break;
      }
      default: {                            // store
        memset(m, 0, 53);
        bool is_set;
        uint32_t s_id, sf, st, end_time = 0;
        if (cfg.store_hot) {
          is_set = c.rng.below(100) < cfg.set_pct;
          uint32_t r = draw_key(c.rng);
          s_id = r / 12; sf = (r % 12) / 3 + 1; st = (r % 3) * 8;
          end_time = c.rng.below(24);
        } else {
          is_set = (fastrand(&c.lcg) % 100) >= (100 - cfg.set_pct);   // workgen_arr: reads first, sets last
          s_id = nurand(&c.lcg);
          sf = (fastrand(&c.lcg) % 4) + 1;
          st = (fastrand(&c.lcg) % 3) * 8;
          if (is_set) end_time = fastrand(&c.lcg) % 24;
        }
        m[0] = is_set ? 1 : 0;
        put64(m + 1, (uint64_t)s_id | ((uint64_t)sf << 32) | ((uint64_t)st << 40));
        if (is_set) { m[9] = (uint8_t)end_time; m[10] = 0x5a; }
        break;
      }
    }
  }

  void absorb(LockClient& c, const uint8_t* m) {
    switch (cfg.kind) {
      case K_FASST: {
        const uint8_t type = m[0];
        switch (c.phase) {
          case PH_READ:
            c.rv[c.pos] = get32(m + 5);
            if (++c.pos == c.nr) { c.pos = 0; c.phase = c.nw ? PH_ACQ : PH_VALIDATE; }
            break;
          case PH_ACQ:
            if (type == 5) {                               // kGrantLock
              if (++c.pos == c.nw) { c.pos = 0; c.phase = PH_VALIDATE; }
            } else {                                       // kRejectLock: abort [0, pos), restart txn
              st_lock_rejects++;
              if (c.pos) { c.lim = c.pos; c.pos = 0; c.phase = PH_ABORT; }
              else { c.pos = 0; c.phase = PH_READ; }
            }
            break;
          case PH_VALIDATE:
            if (get32(m + 5) != c.rv[c.pos]) {             // client.cc:209-212 roll back
              st_validation_aborts++;
              if (c.nw) { c.lim = c.nw; c.pos = 0; c.phase = PH_ABORT; }
              else { c.pos = 0; c.phase = PH_READ; }
            } else if (++c.pos == c.nr) {
              if (c.nw) { c.pos = 0; c.phase = PH_COMMIT; }
              else { st_committed++; new_txn(c); }
            }
            break;
          case PH_ABORT:
            if (++c.pos == c.lim) { c.pos = 0; c.phase = PH_READ; }
            break;
          default:                                         // PH_COMMIT
            if (++c.pos == c.nw) { st_committed++; new_txn(c); }
            break;
        }
        break;
      }
      case K_LOCK2PL: {
        const uint8_t action = m[0];
        if (c.phase == PH_ACQ) {
          if (action == 2) {                               // kGrantLock
            if (++c.pos == c.nr) { c.pos = c.nr - 1; c.phase = PH_RELEASE; }
          } else {                                         // kRejectLock: release [0, pos) ascending, retry
            st_lock_rejects++;
            if (c.pos) { c.lim = c.pos; c.pos = 0; c.phase = PH_REL_ABORT; }
            else c.pos = 0;
          }
        } else if (c.phase == PH_RELEASE) {
          if (c.pos == 0) { st_committed++; new_txn(c); }
          else c.pos--;
        } else {                                           // PH_REL_ABORT
          if (++c.pos == c.lim) { c.pos = 0; c.phase = PH_ACQ; }
        }
        break;
      }
      case K_LOG:
        st_committed++;
        break;
      default:                                             // store: one request = one transaction
        if (m[0] == 7) st_not_exist++; else st_committed++;
        break;
    }
  }
};

extern "C" {

dint_wl* dint_wl_create(const dint_wl_cfg* cfg) {
  if (!cfg || cfg->kind > 3 || cfg->n_clients == 0) return nullptr;
  dint_wl* w = new dint_wl();
  w->cfg = *cfg;
  if (w->cfg.n_keys == 0) w->cfg.n_keys = 24000000u;
  if (w->cfg.store_subscribers == 0) w->cfg.store_subscribers = 2000000u;
  if (w->cfg.zipf_theta > 0) w->zipf.init(w->cfg.n_keys, w->cfg.zipf_theta);
  w->cl.resize(cfg->n_clients);
  for (uint32_t i = 0; i < cfg->n_clients; i++) {
    LockClient& c = w->cl[i];
    c.rng = Rng(cfg->seed * 0x100000001B3ULL + i);
    c.lcg = 0xdeadbeefULL + i;                             // client_udp.cc:201 tg_seed = 0xdeadbeef + wrkr_gid
    if (cfg->kind == K_FASST || cfg->kind == K_LOCK2PL) w->new_txn(c);
  }
  return w;
}
void dint_wl_destroy(dint_wl* w) { delete w; }
uint32_t dint_wl_msg_size(const dint_wl* w) { return kMsg[w->cfg.kind]; }

// one round: request of client i at req + i * msg.  Returns the number of requests (= n_clients).
uint64_t dint_wl_next(dint_wl* w, void* req) {
  uint8_t* out = (uint8_t*)req;
  const uint32_t msg = kMsg[w->cfg.kind];
  for (size_t i = 0; i < w->cl.size(); i++) w->emit(w->cl[i], out + i * msg);
  w->st_requests += w->cl.size();
  w->st_rounds++;
  return w->cl.size();
}
void dint_wl_feed(dint_wl* w, const void* resp) {
  const uint8_t* in = (const uint8_t*)resp;
  const uint32_t msg = kMsg[w->cfg.kind];
  for (size_t i = 0; i < w->cl.size(); i++) w->absorb(w->cl[i], in + i * msg);
}
// out: requests, committed txns, validation aborts, lock rejects, not-exist replies, rounds
void dint_wl_stats(const dint_wl* w, uint64_t out[6]) {
  out[0] = w->st_requests; out[1] = w->st_committed; out[2] = w->st_validation_aborts;
  out[3] = w->st_lock_rejects; out[4] = w->st_not_exist; out[5] = w->st_rounds;
}

// Host twin of dint_route_owner (used by the gloo / CPU tests of the sharded routing logic):
// owner[i] = (fasthash64(key) % mods[table]) % n_shards for requests that touch per-key state, else `self`.
// kind: enum dint_kind (0..5); mods: group modulus per table (lock kinds: mods[0] = lock_slots).
void dint_wl_owner(uint32_t kind, const uint32_t* mods, uint32_t n_shards, uint32_t self, const void* req, uint64_t n,
                   uint8_t* owner) {
  static const uint32_t msg[6] = {6, 9, 53, 53, 55, 23};
  const uint8_t* p = (const uint8_t*)req;
  for (uint64_t i = 0; i < n; i++, p += msg[kind]) {
    uint32_t o = self;
    bool keyed = false;
    uint32_t table = 0;
    uint64_t h = 0;
    switch (kind) {
      case 0: keyed = p[0] <= 1 && !(p[0] == 0 && p[5] > 1); if (keyed) { uint32_t k; memcpy(&k, p + 1, 4); h = dint::fasthash64_u32(k); } break;
      case 1: keyed = p[0] <= 3; if (keyed) { uint32_t k; memcpy(&k, p + 1, 4); h = dint::fasthash64_u32(k); } break;
      case 2: break;
      case 3: keyed = p[0] <= 2; if (keyed) { uint64_t k; memcpy(&k, p + 1, 8); h = dint::fasthash64_u64(k); } break;
      case 4: {
        const uint8_t t = p[1];
        table = p[2];
        keyed = table < 5 && (t <= 2 || t == 12 || t == 13 || t == 18 || t == 19 || t == 22 || t == 23);
        if (keyed) { uint64_t k; memcpy(&k, p + 3, 8); h = dint::fasthash64_u64(k); }
        break;
      }
      default: {
        table = p[2];
        keyed = table < 2 && p[1] <= 5;
        if (keyed) { uint64_t k; memcpy(&k, p + 3, 8); h = dint::fasthash64_u64(k); }
        break;
      }
    }
    if (keyed) o = (uint32_t)((h % mods[table]) % n_shards);
    owner[i] = (uint8_t)o;
  }
}

}  // extern "C"
