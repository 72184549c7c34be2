// kv.cuh -- the KV side of the hot path: the HBM-resident table that stands in for the reference's
// chained `kvs` (store/udp/kvs.h:13-136; tatp/udp/kvs.h and smallbank/udp/kvs.h are the same code with
// panics on missing keys), and the request semantics of the three servers built on it
// (store/udp/server.cc:75-97, tatp/udp/server_shard.cc:113-210, smallbank/udp/server_shard.cc:107-189).
//
// Layout.  The wire-visible behaviour of `kvs` is the map key -> (val, ver, exists); the reference's
// physical layout (bucket-head pointer array -> 224-byte 4-slot entries -> next pointers) costs two
// dependent DRAM misses per lookup.  Here a table is ONE open-addressing array of naturally aligned
// 64-byte entries (32-byte for smallbank's 8-byte values) at a load factor <= 0.5, so a GET that hits
// on its first probe is exactly one aligned 64-byte HBM access:
//     { u64 key; u32 ver; u32 meta; u8 val[VALSZ]; pad }
// meta: EMPTY (never used) / FULL / TOMB (deleted) / BUSY (insert in flight).  Entries never return to
// EMPTY inside a launch, so a probe sequence that reaches EMPTY proves absence.  Deletes leave tombstones
// (an insert reuses the first one on its probe path); the engine counts the entries that have left EMPTY
// (`live[1]`) and, between calls, rehashes a table into a fresh array once FULL + TOMB passes 70 % of its
// capacity (k_kv_rehash, engine.cu kv_maintain) -- the reference's chained kvs frees entries on delete
// (store/udp/kvs.h:124-133), so without this a long insert/delete churn would grow the probe chains without
// bound.  Requests on the same key are never
// concurrent (same group -> K3 replays them in one thread); requests on different keys only meet on
// the `meta` word, which is claimed with atomicCAS.
#pragma once
#include <functional>
#include "engine.cuh"

namespace dint {

template <int VALSZ> struct Ent {
  static constexpr int BYTES = (VALSZ == 40) ? 64 : 32;
  static constexpr int NW = VALSZ / 4;          // value words
  static constexpr int NV = BYTES / 16;         // 16-byte vectors per entry
};
constexpr uint64_t kKvMix = 0x9E3779B97F4A7C15ULL;

#ifdef __CUDACC__
DINT_D uint64_t kv_home(const KvTable& t, uint64_t h) { return (h * kKvMix) >> (64 - t.cap_log2); }

template <int VALSZ>
DINT_D void kv_load_entry(const uint8_t* e, uint4 (&v)[Ent<VALSZ>::NV]) {
#pragma unroll
  for (int k = 0; k < Ent<VALSZ>::NV; k++) v[k] = __ldcg((const uint4*)e + k);   // one aligned entry, L1 bypass
}

// Probe for `key`.  `v` arrives holding the HOME entry (fetched by prefetch()); on a hit returns the
// entry and leaves its 16-byte vectors in `v` (v[0] = key lo/hi, ver, meta; v[1..] = value).
template <int VALSZ>
DINT_D uint8_t* kv_find(const KvTable& t, uint64_t key, uint64_t h, uint4 (&v)[Ent<VALSZ>::NV]) {
  uint64_t i = kv_home(t, h);
  for (uint64_t probe = 0; probe <= t.cap_mask; probe++) {
    uint8_t* e = t.entries + (i << t.ent_shift);
    if (probe) kv_load_entry<VALSZ>(e, v);
    uint32_t meta = v[0].w;
    if (meta == ENT_EMPTY) return nullptr;
    if (meta == ENT_FULL && v[0].x == (uint32_t)key && v[0].y == (uint32_t)(key >> 32)) return e;
    i = (i + 1) & t.cap_mask;
  }
  return nullptr;
}

// kvs_get (kvs.h:37-55): on a hit copy val and ver into the wire record.
template <int VALSZ>
DINT_D bool kv_get_into(const KvTable& t, uint64_t key, uint64_t h, uint4 (&v)[Ent<VALSZ>::NV], uint8_t* wire_val,
                        uint8_t* wire_ver) {
  if (!kv_find<VALSZ>(t, key, h, v)) return false;
  uint32_t w[Ent<VALSZ>::NW];
#pragma unroll
  for (int k = 0; k < Ent<VALSZ>::NW; k++) {
    const uint4& q = v[1 + k / 4];
    w[k] = (k % 4 == 0) ? q.x : (k % 4 == 1) ? q.y : (k % 4 == 2) ? q.z : q.w;
  }
  st_words_unaligned<Ent<VALSZ>::NW>(wire_val, w);
  st_u32_unaligned(wire_ver, v[0].z);
  return true;
}

template <int VALSZ>
DINT_D void kv_write_val(uint8_t* e, const uint32_t (&w)[Ent<VALSZ>::NW]) {
  if constexpr (VALSZ == 40) {
    *((uint4*)(e + 16)) = make_uint4(w[0], w[1], w[2], w[3]);
    *((uint4*)(e + 32)) = make_uint4(w[4], w[5], w[6], w[7]);
    *((uint2*)(e + 48)) = make_uint2(w[8], w[9]);
  } else {
    *((uint2*)(e + 16)) = make_uint2(w[0], w[1]);
  }
}

// kvs_set (kvs.h:57-75): overwrite val, ver++.
template <int VALSZ>
DINT_D bool kv_set_from(const KvTable& t, uint64_t key, uint64_t h, uint4 (&v)[Ent<VALSZ>::NV], const uint8_t* wire_val) {
  uint8_t* e = kv_find<VALSZ>(t, key, h, v);
  if (!e) return false;
  uint32_t w[Ent<VALSZ>::NW];
  ld_words_unaligned<Ent<VALSZ>::NW>(wire_val, w);
  kv_write_val<VALSZ>(e, w);
  *((uint32_t*)(e + 8)) = v[0].z + 1;
  return true;
}

// kvs_insert (kvs.h:77-104): take the first free entry on the probe path, ver = 0.  Like the
// reference it does not look for an existing copy of the key.
template <int VALSZ>
DINT_D bool kv_insert_words(const KvTable& t, uint64_t key, uint64_t h, const uint32_t (&w)[Ent<VALSZ>::NW], uint32_t ver0 = 0) {
  uint64_t i = kv_home(t, h);
  for (uint64_t probe = 0; probe <= t.cap_mask; probe++) {
    uint8_t* e = t.entries + (i << t.ent_shift);
    uint32_t* meta = (uint32_t*)(e + 12);
    uint32_t m = __ldcg(meta);
    while (m == ENT_EMPTY || m == ENT_TOMB) {
      uint32_t old = atomicCAS(meta, m, (uint32_t)ENT_BUSY);
      if (old == m) {
        *((uint64_t*)e) = key;
        *((uint32_t*)(e + 8)) = ver0;
        kv_write_val<VALSZ>(e, w);
        __threadfence();
        *((volatile uint32_t*)meta) = ENT_FULL;
        atomicAdd(t.live, 1ULL);
        if (m == ENT_EMPTY) atomicAdd(t.live + 1, 1ULL);   // one more entry that can never prove absence again
        return true;
      }
      m = old;
    }
    i = (i + 1) & t.cap_mask;
  }
  return false;    // table full
}
template <int VALSZ>
DINT_D bool kv_insert_from(const KvTable& t, uint64_t key, uint64_t h, const uint8_t* wire_val) {
  uint32_t w[Ent<VALSZ>::NW];
  ld_words_unaligned<Ent<VALSZ>::NW>(wire_val, w);
  return kv_insert_words<VALSZ>(t, key, h, w);
}

// kvs_delete (kvs.h:106-136)
template <int VALSZ>
DINT_D bool kv_delete(const KvTable& t, uint64_t key, uint64_t h, uint4 (&v)[Ent<VALSZ>::NV]) {
  uint8_t* e = kv_find<VALSZ>(t, key, h, v);
  if (!e) return false;
  *((volatile uint32_t*)(e + 12)) = ENT_TOMB;
  atomicAdd(t.live, (unsigned long long)-1LL);
  return true;
}

// group of a KV key: the reference's bucket (store) or lock_hash (tatp.h:12-14, smallbank.h:12-14)
DINT_D uint32_t kv_group(const Ctx& c, uint32_t table, uint64_t h) {
  const KvTable& t = c.tbl[table];
  uint32_t gl;
  if (!to_local_group(c, fast_mod(h, t.lock_mod), gl)) return kNoGroup;
  return t.grp_base + gl;
}
// the home entry of (table, key): the one access a first-probe hit needs
template <int VALSZ>
DINT_D void kv_prefetch_home(const Ctx& c, uint32_t table, uint64_t h, uint4 (&v)[Ent<VALSZ>::NV]) {
  const KvTable& t = c.tbl[table];
  kv_load_entry<VALSZ>(t.entries + (kv_home(t, h) << t.ent_shift), v);
}

// Warp-cooperative fetch of every lane's home entry: NV adjacent lanes read one entry with a single
// coalesced request (64 B = 4 lanes x 16 B, 32 B = 2 lanes), then the 16-byte pieces are handed to the
// owning lane with shuffles.  Must be executed by all 32 lanes; `need` = this lane wants its entry.
template <int VALSZ>
DINT_D void kv_prefetch_home_coop(const uint8_t* entry, bool need, uint4 (&v)[Ent<VALSZ>::NV]) {
  constexpr int LPE = Ent<VALSZ>::NV;        // lanes per entry
  constexpr int EPR = 32 / LPE;              // entries fetched per round
  const uint32_t lane = threadIdx.x & 31;
  const unsigned long long a = (unsigned long long)entry;
#pragma unroll
  for (int r = 0; r < LPE; r++) {
    const int src = r * EPR + (int)(lane / LPE);
    const unsigned long long sa = __shfl_sync(0xffffffffu, a, src);
    const int sneed = __shfl_sync(0xffffffffu, need ? 1 : 0, src);
    uint4 piece = make_uint4(0, 0, 0, 0);
    if (sneed) piece = __ldcg((const uint4*)sa + (lane % LPE));
#pragma unroll
    for (int k = 0; k < LPE; k++) {
      const int from = (int)(lane % EPR) * LPE + k;
      uint4 got;
      got.x = __shfl_sync(0xffffffffu, piece.x, from);
      got.y = __shfl_sync(0xffffffffu, piece.y, from);
      got.z = __shfl_sync(0xffffffffu, piece.z, from);
      got.w = __shfl_sync(0xffffffffu, piece.w, from);
      if ((int)(lane / EPR) == r) v[k] = got;
    }
  }
}
template <int VALSZ>
DINT_D const uint8_t* kv_home_ptr(const Ctx& c, uint32_t table, uint64_t h) {
  const KvTable& t = c.tbl[table];
  return t.entries + (kv_home(t, h) << t.ent_shift);
}

// =================================== store ==========================================================
template <> DINT_D TypeInfo type_info<K_STORE>(const uint8_t* rec) {
  uint8_t t = rec[Wire<K_STORE>::TYPE];
  if (t == 0) return TypeInfo{C_RA, false, false};          // kRead  store/udp/server.cc:79-84
  if (t == 1 || t == 2) return TypeInfo{C_WA, false, false};  // kSet :86-91; kInsert: the eBPF server's
                                                            // population path (store/ebpf/store_user.c)
  return TypeInfo{0, true, false};                          // :93-94
}
template <> DINT_D KeyInfo key_info<K_STORE>(const Ctx& c, const uint8_t* rec) {
  KeyInfo k;
  k.key = ld_u64_unaligned(rec + Wire<K_STORE>::KEY);
  k.h = fasthash64_u64(k.key);                              // kvs.h:33-35
  k.grp = kv_group(c, 0, k.h);
  return k;
}
template <> struct Pre<K_STORE> { uint4 v[4]; };
template <> DINT_D Pre<K_STORE> prefetch<K_STORE>(const Ctx& c, const uint8_t* rec, const KeyInfo& ki, const TypeInfo&) {
  Pre<K_STORE> p;
  if (rec[Wire<K_STORE>::TYPE] != 2) kv_prefetch_home<40>(c, 0, ki.h, p.v);
  return p;
}
template <> DINT_D Pre<K_STORE> prefetch_coop<K_STORE>(const Ctx& c, const uint8_t* rec, const KeyInfo& ki, const TypeInfo&, bool active) {
  Pre<K_STORE> p;
  const bool need = active && rec[Wire<K_STORE>::TYPE] != 2;
  kv_prefetch_home_coop<40>(need ? kv_home_ptr<40>(c, 0, ki.h) : nullptr, need, p.v);
  return p;
}
template <>
DINT_D void apply_one<K_STORE>(const Ctx& c, uint8_t* rec, const KeyInfo& ki, const Pre<K_STORE>& pf, unsigned long long, bool) {
  using W = Wire<K_STORE>;
  const uint8_t t = rec[W::TYPE];
  uint4 v[4] = {pf.v[0], pf.v[1], pf.v[2], pf.v[3]};
  if (t == 0) {
    rec[W::TYPE] = kv_get_into<40>(c.tbl[0], ki.key, ki.h, v, rec + W::VAL, rec + W::VER) ? 3 : 7;   // kGrantRead / kNotExist
  } else if (t == 1) {
    rec[W::TYPE] = kv_set_from<40>(c.tbl[0], ki.key, ki.h, v, rec + W::VAL) ? 5 : 7;                 // kSetAck / kNotExist
  } else {
    if (kv_insert_from<40>(c.tbl[0], ki.key, ki.h, rec + W::VAL)) rec[W::TYPE] = 8;                  // kInsertAck
    else mark_invalid<K_STORE>(c, rec);
  }
}

// =================================== tatp ===========================================================
template <> DINT_D TypeInfo type_info<K_TATP>(const uint8_t* rec) {
  using W = Wire<K_TATP>;
  // kCommitLog / kDeleteLog never index tables[]: they log msg.table verbatim (tatp/udp/server_shard.cc:182-207)
  if (rec[W::TYPE] == 14 || rec[W::TYPE] == 24) return TypeInfo{0, false, true};
  if (rec[W::TABLE] >= 5) return TypeInfo{0, true, false};
  switch (rec[W::TYPE]) {
    case 0: return TypeInfo{C_RA, false, false};                    // kRead
    case 1: case 2: return TypeInfo{C_WL, false, false};            // kAcquireLock, kAbort
    case 12: case 18: case 22: return TypeInfo{C_WA | C_WL, false, false};  // kCommitPrim, kInsertPrim, kDeletePrim
    case 13: case 19: case 23: return TypeInfo{C_WA, false, false};         // kCommitBck, kInsertBck, kDeleteBck
    default: return TypeInfo{0, true, false};                       // tatp/udp/server_shard.cc:209
  }
}
template <> DINT_D KeyInfo key_info<K_TATP>(const Ctx& c, const uint8_t* rec) {
  using W = Wire<K_TATP>;
  KeyInfo k;
  k.key = ld_u64_unaligned(rec + W::KEY);
  k.h = fasthash64_u64(k.key);
  k.grp = kv_group(c, rec[W::TABLE], k.h);                          // lock_hash, tatp/udp/tatp.h:12-14
  return k;
}
template <> struct Pre<K_TATP> { uint4 v[4]; };
DINT_D bool tatp_touches_row(uint8_t type) {   // request types that look a row up (not insert / lock / log)
  return type == 0 || type == 12 || type == 13 || type == 22 || type == 23;
}
template <> DINT_D Pre<K_TATP> prefetch<K_TATP>(const Ctx& c, const uint8_t* rec, const KeyInfo& ki, const TypeInfo&) {
  using W = Wire<K_TATP>;
  Pre<K_TATP> p;
  if (tatp_touches_row(rec[W::TYPE])) kv_prefetch_home<40>(c, rec[W::TABLE], ki.h, p.v);
  return p;
}
template <> DINT_D Pre<K_TATP> prefetch_coop<K_TATP>(const Ctx& c, const uint8_t* rec, const KeyInfo& ki, const TypeInfo&, bool active) {
  using W = Wire<K_TATP>;
  Pre<K_TATP> p;
  const bool need = active && tatp_touches_row(rec[W::TYPE]);
  kv_prefetch_home_coop<40>(need ? kv_home_ptr<40>(c, rec[W::TABLE], ki.h) : nullptr, need, p.v);
  return p;
}
template <>
DINT_D void apply_one<K_TATP>(const Ctx& c, uint8_t* rec, const KeyInfo& ki, const Pre<K_TATP>& pf, unsigned long long ord,
                              bool keep) {
  using W = Wire<K_TATP>;
  const uint8_t type = rec[W::TYPE], table = rec[W::TABLE];
  if (type == 14 || type == 24) {          // kCommitLog :182-194 / kDeleteLog :196-207
    if (keep) {                            // log_entry {is_del@0 table@1 key@8 val@16 ver@56} tatp/udp/kvs.h:23-29
      uint8_t* e = c.ring + (size_t)(ord % c.ring_n) * W::LOGENT;
      uint32_t w[13];
      ld_words_unaligned<13>(rec + W::KEY, w);          // key, val, ver are contiguous on the wire
      e[0] = (type == 24);
      e[1] = table;
      *((uint2*)(e + 8)) = make_uint2(w[0], w[1]);
      if (type == 14) {
        *((uint4*)(e + 16)) = make_uint4(w[2], w[3], w[4], w[5]);
        *((uint4*)(e + 32)) = make_uint4(w[6], w[7], w[8], w[9]);
        *((uint2*)(e + 48)) = make_uint2(w[10], w[11]);
      }
      *((uint32_t*)(e + 56)) = w[12];
    }
    rec[W::TYPE] = (type == 14) ? 17 : 27;
    return;
  }
  const KvTable& t = c.tbl[table];
  const uint64_t key = ki.key, h = ki.h;
  const uint32_t g = ki.grp;
  uint4 v[4] = {pf.v[0], pf.v[1], pf.v[2], pf.v[3]};
  bool ok = true;
  switch (type) {
    case 0: rec[W::TYPE] = kv_get_into<40>(t, key, h, v, rec + W::VAL, rec + W::VER) ? 4 : 6; break;  // :116-121
    case 1: rec[W::TYPE] = bm_fetch_set(c.lockbits, g) ? 8 : 7; break;                                // :123-132
    case 2: bm_clear_bit(c.lockbits, g); rec[W::TYPE] = 9; break;                                     // :134-138
    // a would-panic request (kvs_set / kvs_delete on a missing key) applies NOTHING: the reference dies before the unlock
    case 12: ok = kv_set_from<40>(t, key, h, v, rec + W::VAL); if (ok) bm_clear_bit(c.lockbits, g); rec[W::TYPE] = 15; break;  // :140-146
    case 18: ok = kv_insert_from<40>(t, key, h, rec + W::VAL); if (ok) bm_clear_bit(c.lockbits, g); rec[W::TYPE] = 20; break;  // :148-154
    case 22: ok = kv_delete<40>(t, key, h, v); if (ok) bm_clear_bit(c.lockbits, g); rec[W::TYPE] = 25; break;                  // :156-162
    case 13: ok = kv_set_from<40>(t, key, h, v, rec + W::VAL); rec[W::TYPE] = 16; break;              // :164-168
    case 19: ok = kv_insert_from<40>(t, key, h, rec + W::VAL); rec[W::TYPE] = 21; break;              // :170-174
    default: ok = kv_delete<40>(t, key, h, v); rec[W::TYPE] = 26; break;                              // 23 :176-180
  }
  if (!ok) mark_invalid<K_TATP>(c, rec);   // kvs_set / kvs_delete on a missing key: tatp/udp/kvs.h:91,152 panic
}

// =================================== smallbank ======================================================
template <> DINT_D TypeInfo type_info<K_SMALLBANK>(const uint8_t* rec) {
  using W = Wire<K_SMALLBANK>;
  if (rec[W::TABLE] >= 2) return TypeInfo{0, true, false};
  switch (rec[W::TYPE]) {
    case 0: case 1: return TypeInfo{C_WL | C_RA, false, false};     // kAcquireShared / kAcquireExclusive (+kvs_get)
    case 2: case 3: return TypeInfo{C_WL, false, false};            // kReleaseShared / kReleaseExclusive
    case 4: case 5: return TypeInfo{C_WA, false, false};            // kCommitPrim / kCommitBck
    case 6: return TypeInfo{0, false, true};                        // kCommitLog
    default: return TypeInfo{0, true, false};                       // smallbank/udp/server_shard.cc:188
  }
}
template <> DINT_D KeyInfo key_info<K_SMALLBANK>(const Ctx& c, const uint8_t* rec) {
  using W = Wire<K_SMALLBANK>;
  KeyInfo k;
  k.key = ld_u64_unaligned(rec + W::KEY);
  k.h = fasthash64_u64(k.key);
  k.grp = kv_group(c, rec[W::TABLE], k.h);                          // :109 lock_hash
  return k;
}
template <> struct Pre<K_SMALLBANK> { uint4 v[2]; uint2 s; };
template <> DINT_D Pre<K_SMALLBANK> prefetch<K_SMALLBANK>(const Ctx& c, const uint8_t* rec, const KeyInfo& ki, const TypeInfo&) {
  using W = Wire<K_SMALLBANK>;
  Pre<K_SMALLBANK> p;
  const uint8_t type = rec[W::TYPE];
  if (type <= 3) p.s = __ldcg(&c.cnt2[ki.grp]);
  if (type != 2 && type != 3) kv_prefetch_home<8>(c, rec[W::TABLE], ki.h, p.v);
  return p;
}
template <> DINT_D Pre<K_SMALLBANK> prefetch_coop<K_SMALLBANK>(const Ctx& c, const uint8_t* rec, const KeyInfo& ki, const TypeInfo&, bool active) {
  using W = Wire<K_SMALLBANK>;
  Pre<K_SMALLBANK> p;
  const uint8_t type = active ? rec[W::TYPE] : 2;
  if (active && type <= 3) p.s = __ldcg(&c.cnt2[ki.grp]);
  const bool need = active && type != 2 && type != 3;
  kv_prefetch_home_coop<8>(need ? kv_home_ptr<8>(c, rec[W::TABLE], ki.h) : nullptr, need, p.v);
  return p;
}
template <>
DINT_D void apply_one<K_SMALLBANK>(const Ctx& c, uint8_t* rec, const KeyInfo& ki, const Pre<K_SMALLBANK>& pf,
                                   unsigned long long ord, bool keep) {
  using W = Wire<K_SMALLBANK>;
  const uint8_t type = rec[W::TYPE], table = rec[W::TABLE];
  if (type == 6) {                         // kCommitLog :175-186; log_entry {table@0 key@8 val@16 ver@24}
    if (keep) {
      uint8_t* e = c.ring + (size_t)(ord % c.ring_n) * W::LOGENT;
      uint32_t w[5];
      ld_words_unaligned<5>(rec + W::KEY, w);
      e[0] = table;
      *((uint2*)(e + 8)) = make_uint2(w[0], w[1]);
      *((uint2*)(e + 16)) = make_uint2(w[2], w[3]);
      *((uint32_t*)(e + 24)) = w[4];
    }
    rec[W::TYPE] = 15;
    return;
  }
  const KvTable& t = c.tbl[table];
  const uint64_t key = ki.key, h = ki.h;
  const uint32_t g = ki.grp;
  uint4 v[2] = {pf.v[0], pf.v[1]};
  bool ok = true;
  if (type <= 3) {
    uint2 s = pf.s;                        // x = num_ex, y = num_sh
    if (type == 0) {                       // :121-133
      if (s.x == 0) { s.y++; c.cnt2[g] = s; ok = kv_get_into<8>(t, key, h, v, rec + W::VAL, rec + W::VER); rec[W::TYPE] = 7; }
      else rec[W::TYPE] = 8;
    } else if (type == 1) {                // :135-147
      if (s.x == 0 && s.y == 0) { s.x++; c.cnt2[g] = s; ok = kv_get_into<8>(t, key, h, v, rec + W::VAL, rec + W::VER); rec[W::TYPE] = 9; }
      else rec[W::TYPE] = 10;
    } else if (type == 2) { s.y--; c.cnt2[g] = s; rec[W::TYPE] = 11; }   // :149-154
    else { s.x--; c.cnt2[g] = s; rec[W::TYPE] = 12; }                    // :156-161
  } else {                                 // kCommitPrim :163-167 / kCommitBck :169-173
    ok = kv_set_from<8>(t, key, h, v, rec + W::VAL);
    rec[W::TYPE] = (type == 4) ? 13 : 14;
  }
  if (!ok) mark_invalid<K_SMALLBANK>(c, rec);   // smallbank/udp/kvs.h:67,86 panic
}

// ---- bulk load (dint_load / dint_populate) and single-key inspection ------------------------------
template <int VALSZ>
__global__ void __launch_bounds__(256) k_kv_load(const Ctx c, int table, const uint64_t* keys, const uint8_t* vals, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const KvTable& t = c.tbl[table];
  uint64_t key = keys[i];
  uint64_t h = fasthash64_u64(key);
  if (kv_group(c, table, h) == kNoGroup) return;   // another shard's key
  uint32_t w[Ent<VALSZ>::NW];
  const uint32_t* src = (const uint32_t*)(vals + (size_t)i * VALSZ);
#pragma unroll
  for (int k = 0; k < Ent<VALSZ>::NW; k++) w[k] = src[k];
  if (!kv_insert_words<VALSZ>(t, key, h, w)) atomicAdd(&c.counters[0], 1ULL);
}

// Rehash of one table into a fresh (zeroed) array: every FULL entry moves with its version, tombstones vanish.
template <int VALSZ>
__global__ void __launch_bounds__(256) k_kv_rehash(const KvTable from, const KvTable to) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= from.cap_mask; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint8_t* e = from.entries + (i << from.ent_shift);
    uint4 v[Ent<VALSZ>::NV];
    kv_load_entry<VALSZ>(e, v);
    if (v[0].w != ENT_FULL) continue;
    const uint64_t key = ((uint64_t)v[0].y << 32) | v[0].x;
    uint32_t w[Ent<VALSZ>::NW];
    const uint32_t* flat = (const uint32_t*)v;
#pragma unroll
    for (int k = 0; k < Ent<VALSZ>::NW; k++) w[k] = flat[4 + k];
    kv_insert_words<VALSZ>(to, key, fasthash64_u64(key), w, v[0].z);   // (`to` is at least as large: cannot fail)
  }
}

template <int VALSZ>
__global__ void k_kv_get1(const Ctx c, int table, uint64_t key, uint32_t* out /* [0]=found [1]=ver [2..]=val */) {
  uint4 v[Ent<VALSZ>::NV];
  const uint64_t h = fasthash64_u64(key);
  kv_prefetch_home<VALSZ>(c, table, h, v);
  uint8_t* e = kv_find<VALSZ>(c.tbl[table], key, h, v);
  out[0] = e ? 1u : 0u;
  if (e) {
    out[1] = v[0].z;
    const uint32_t* flat = (const uint32_t*)v;
    for (int k = 0; k < Ent<VALSZ>::NW; k++) out[2 + k] = flat[4 + k];
  }
}
#endif  // __CUDACC__

// =================================== host side ======================================================
struct KvHost {
  uint64_t capacity = 0;
  uint32_t hash_size = 0;     // the reference's bucket count for this table
};

inline uint32_t next_log2(uint64_t x) {
  uint32_t l = 0;
  while ((1ULL << l) < x) l++;
  return l;
}

// Sizes follow the reference: store kvs_init(kSubscriberNum*18/4) (store/udp/server.cc:113); tatp
// kvs_init(S*3/2/4, S*3/2/4, S*15/4/4, S*15/4/4, S*45/8/4) (tatp/udp/server_shard.cc:75-79); smallbank
// kvs_init(A*3/2/4) x2 (smallbank/udp/server_shard.cc:72-73).  The group modulus is the bucket count
// for store and kKeysPerEntry*hash_size (lock_hash) for tatp / smallbank.
template <typename AllocFn>
int kv_create_tables(int kind, const dint_cfg& cf, Ctx& c, KvHost* kv, uint64_t* groups_out, AllocFn alloc) {
  const uint64_t S = cf.subs_sizing, Sp = cf.subs_populate, A = cf.accts_sizing, Ap = cf.accts_populate;
  uint32_t hs[kMaxTables] = {0};
  double expect[kMaxTables] = {0};
  uint32_t nt = 0, valsz = 40, lock_mul = 4;
  if (kind == DINT_STORE) {
    nt = 1; hs[0] = (uint32_t)(S * 18 / 4); expect[0] = 12.0 * Sp; lock_mul = 1;
  } else if (kind == DINT_TATP) {
    nt = 5;
    hs[0] = hs[1] = (uint32_t)(S * 3 / 2 / 4);
    hs[2] = hs[3] = (uint32_t)(S * 15 / 4 / 4);
    hs[4] = (uint32_t)(S * 45 / 8 / 4);
    expect[0] = expect[1] = 1.0 * Sp; expect[2] = expect[3] = 2.5 * Sp; expect[4] = 3.75 * Sp;
  } else if (kind == DINT_SMALLBANK) {
    nt = 2; valsz = 8;
    hs[0] = hs[1] = (uint32_t)(A * 3 / 2 / 4);
    expect[0] = expect[1] = 1.0 * Ap;
  } else return DINT_EINVAL;
  c.n_tables = nt;
  uint64_t base = 0;
  for (uint32_t t = 0; t < nt; t++) {
    if (hs[t] == 0) return DINT_EINVAL;
    KvTable& T = c.tbl[t];
    uint32_t lg = cf.kv_capacity_log2[t];
    if (lg == 0) {
      double need = 2.0 * expect[t] / cf.n_shards * 1.05 + 1024;
      lg = next_log2((uint64_t)need);
      if (lg < 10) lg = 10;
    }
    if (lg > 34) return DINT_EINVAL;
    T.cap_log2 = lg;
    T.cap_mask = (1ULL << lg) - 1;
    T.ent_shift = (valsz == 40) ? 6 : 5;
    uint64_t mod = (uint64_t)hs[t] * lock_mul;
    if (mod >= 0xffffffffULL) return DINT_EINVAL;
    T.lock_mod = make_fastmod((uint32_t)mod);
    T.grp_base = (uint32_t)base;
    T.n_groups = (uint32_t)((mod + cf.n_shards - 1) / cf.n_shards);
    base += T.n_groups;
    void* p = nullptr;
    int rc = alloc(&p, (size_t)(1ULL << lg) << T.ent_shift);
    if (rc) return rc;
    T.entries = (uint8_t*)p;
    if (t == 0) {                                    // {live, used} of every table side by side: one copy publishes them all
      rc = alloc(&p, 16 * kMaxTables);
      if (rc) return rc;
      c.tbl[0].live = (unsigned long long*)p;
    }
    T.live = c.tbl[0].live + 2 * t;
    kv[t].capacity = 1ULL << lg;
    kv[t].hash_size = hs[t];
  }
  // group state: tatp = one lock bit per group; smallbank = {num_ex, num_sh} per group
  void* p = nullptr;
  if (kind == DINT_TATP) {
    int rc = alloc(&p, ((base + 31) / 32) * 4);
    if (rc) return rc;
    c.lockbits = (uint32_t*)p;
  } else if (kind == DINT_SMALLBANK) {
    int rc = alloc(&p, base * sizeof(uint2));
    if (rc) return rc;
    c.cnt2 = (uint2*)p;
  }
  *groups_out = base;
  return DINT_OK;
}

#ifdef __CUDACC__
inline void kv_launch_load(int kind, const Ctx& c, int table, const uint64_t* dk, const uint8_t* dv, uint32_t n, cudaStream_t s) {
  if (n == 0) return;
  uint32_t blocks = (n + 255) / 256;
  if (kind == DINT_SMALLBANK) k_kv_load<8><<<blocks, 256, 0, s>>>(c, table, dk, dv, n);
  else k_kv_load<40><<<blocks, 256, 0, s>>>(c, table, dk, dv, n);
}

inline int kv_host_get(const Ctx& c, int table, uint64_t key, uint32_t valsz, void* val, uint32_t* ver) {
  uint32_t* d = nullptr;
  if (cudaMalloc(&d, 64) != cudaSuccess) return DINT_ENOMEM;
  if (valsz == 8) k_kv_get1<8><<<1, 1>>>(c, table, key, d);
  else k_kv_get1<40><<<1, 1>>>(c, table, key, d);
  uint32_t h[12] = {0};
  cudaError_t ce = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (ce != cudaSuccess) return DINT_EIO;
  if (!h[0]) return 1;
  if (ver) *ver = h[1];
  if (val) memcpy(val, &h[2], valsz);
  return 0;
}
#endif

// ---- deterministic population: the reference's populate_* generators, emitted as (key, val) batches --
inline uint32_t kv_fastrand(uint64_t* seed) {       // store/udp/tatp.h:31-34, tatp/udp/tatp.h:32-35
  *seed = *seed * 1103515245ULL + 12345ULL;
  return (uint32_t)(*seed >> 32);
}
inline uint64_t tatp_sub_nbr_of(uint32_t s_id) {    // tatp/udp/tatp.h:17-25,132-144: 3 x 12-bit BCD groups
  uint64_t r = 0;
  for (int g = 0; g < 3; g++) {
    uint32_t i = s_id % 1000;
    s_id /= 1000;
    r |= ((uint64_t)(((i / 100) % 10) << 8 | ((i / 10) % 10) << 4 | (i % 10))) << (12 * g);
  }
  return r;
}
inline int tatp_select_types(uint64_t* seed, uint8_t out[4]) {   // tatp/udp/tatp.h:254-282 with values {1,2,3,4}
  bool used[8] = {false};
  int want = (int)(kv_fastrand(seed) % 4) + 1, got = 0;
  while (got < want) {
    uint8_t v = (uint8_t)((kv_fastrand(seed) % 4) + 1);
    if (used[v]) continue;
    used[v] = true;
    out[got++] = v;
  }
  return got;
}

struct KvBatch {
  std::vector<uint64_t> keys;
  std::vector<uint8_t> vals;
  uint32_t valsz;
  int table;
  std::function<int(int, const uint64_t*, const void*, uint64_t)> sink;
  int rc = 0;
  KvBatch(int table_, uint32_t valsz_, std::function<int(int, const uint64_t*, const void*, uint64_t)> s, const dint_cfg& cf)
      : valsz(valsz_), table(table_), sink(std::move(s)), G(cf.txn_shards), gid(cf.txn_shard_id) {
    keys.reserve(1 << 20);
    vals.reserve((size_t)valsz << 20);
  }
  uint32_t G = 0, gid = 0;                           // replica filter (dint_cfg.txn_shards / txn_shard_id)
  std::vector<uint8_t> dummy;
  uint8_t* add(uint64_t key) {                       // returns the zeroed value slot
    if (G > 3) {                                     // not one of this key's three replica holders: drop it
      const uint32_t p = (uint32_t)(key % G), d = (gid + G - p) % G;
      if (d > 2) { dummy.assign(valsz, 0); return dummy.data(); }
    }
    if (keys.size() == (1u << 20)) flush();
    keys.push_back(key);
    vals.resize(vals.size() + valsz, 0);
    return vals.data() + vals.size() - valsz;
  }
  void flush() {
    if (!keys.empty() && rc == 0) rc = sink(table, keys.data(), vals.data(), keys.size());
    keys.clear();
    vals.clear();
  }
};

// Bytes the reference never assigns (stack structs copied whole: store_val_t.numberx[1..],
// tatp_sub_val_t.sub_nbr_unused, tatp_accinf_val_t.data2.., tatp_specfac_val_t.error_cntl/data_a/
// data_b[1..], tatp_callfwd_val_t.numberx[1..]) are zero here -- as they read back from oracle/_ref.
inline int kv_populate(int kind, const dint_cfg& cf, std::function<int(int, const uint64_t*, const void*, uint64_t)> sink) {
  if (kind == DINT_STORE) {                          // store/udp/tatp.h:45-66
    KvBatch b(0, 40, sink, cf);
    uint64_t seed = 0xdeadbeef;
    for (uint32_t s = 0; s < cf.subs_populate; s++)
      for (uint64_t sf = 1; sf <= 4; sf++)
        for (uint64_t st = 0; st <= 16; st += 8) {
          uint8_t* v = b.add((uint64_t)s | (sf << 32) | (st << 40));
          v[0] = (uint8_t)((kv_fastrand(&seed) % 24) + 1);   // end_time
          v[1] = 0x5a;                                       // numberx[0] = kValMagic
        }
    b.flush();
    return b.rc;
  }
  if (kind == DINT_SMALLBANK) {                      // smallbank/udp/smallbank.h:105-127
    KvBatch sav(0, 8, sink, cf), chk(1, 8, sink, cf);
    const float bal = 1000000000.0f;
    for (uint32_t a = 0; a < cf.accts_populate; a++) {
      uint8_t* v = sav.add(a);
      uint32_t magic = 97;
      memcpy(v, &magic, 4); memcpy(v + 4, &bal, 4);
      v = chk.add(a);
      magic = 98;
      memcpy(v, &magic, 4); memcpy(v + 4, &bal, 4);
    }
    sav.flush(); chk.flush();
    return sav.rc ? sav.rc : chk.rc;
  }
  if (kind != DINT_TATP) return DINT_OK;
  const uint32_t N = cf.subs_populate;
  {                                                  // tatp/udp/tatp.h:285-311 subscriber
    KvBatch b(0, 40, sink, cf);
    uint64_t seed = 0xdeadbeef;
    for (uint32_t s = 0; s < N; s++) {
      uint8_t* v = b.add(s);
      uint64_t nbr = tatp_sub_nbr_of(s);
      memcpy(v, &nbr, 8);
      for (int i = 0; i < 5; i++) v[15 + i] = (uint8_t)kv_fastrand(&seed);    // hex[5]
      for (int i = 0; i < 10; i++) v[20 + i] = (uint8_t)kv_fastrand(&seed);   // bytes[10]
      uint16_t bits = (uint16_t)kv_fastrand(&seed);
      memcpy(v + 30, &bits, 2);
      uint32_t msc = 97, vlr = kv_fastrand(&seed);
      memcpy(v + 32, &msc, 4); memcpy(v + 36, &vlr, 4);
    }
    b.flush();
    if (b.rc) return b.rc;
  }
  {                                                  // :314-329 secondary subscriber
    KvBatch b(1, 40, sink, cf);
    for (uint32_t s = 0; s < N; s++) {
      uint8_t* v = b.add(tatp_sub_nbr_of(s));
      memcpy(v, &s, 4);
      v[4] = 98;
    }
    b.flush();
    if (b.rc) return b.rc;
  }
  {                                                  // :332-357 access info
    KvBatch b(2, 40, sink, cf);
    uint64_t seed = 0xdeadbeef;
    for (uint32_t s = 0; s < N; s++) {
      uint8_t ty[4];
      int n = tatp_select_types(&seed, ty);
      for (int i = 0; i < n; i++) b.add((uint64_t)s | ((uint64_t)ty[i] << 32))[0] = 99;
    }
    b.flush();
    if (b.rc) return b.rc;
  }
  {                                                  // :360-412 special facility + call forwarding
    KvBatch sf(3, 40, sink, cf), cfw(4, 40, sink, cf);
    uint64_t seed = 0xdeadbeef;
    for (uint32_t s = 0; s < N; s++) {
      uint8_t ty[4];
      int n = tatp_select_types(&seed, ty);
      for (int i = 0; i < n; i++) {
        uint64_t t = ty[i];
        uint8_t* v = sf.add((uint64_t)s | (t << 32));
        v[3] = 100;
        v[0] = (kv_fastrand(&seed) % 100 < 85) ? 1 : 0;
        for (uint64_t st = 0; st <= 16; st += 8) {
          if (kv_fastrand(&seed) % 2 == 0) continue;
          uint8_t* w = cfw.add((uint64_t)s | (t << 32) | (st << 40));
          w[1] = 101;
          w[0] = (uint8_t)((kv_fastrand(&seed) % 24) + 1);
        }
      }
    }
    sf.flush(); cfw.flush();
    return sf.rc ? sf.rc : cfw.rc;
  }
}

}  // namespace dint
