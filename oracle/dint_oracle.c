/*
 * dint_oracle.c -- TEST INFRASTRUCTURE (see dint_oracle.h).  Plain-C, single-threaded restatement
 * of the six DINT reference UDP server handlers.  Every function cites the reference file:line it
 * follows (paths relative to /root/reference).  A 1-thread reference server never produces kRetry
 * (lock_2pl/udp/server.cc:75-80, smallbank/udp/server_shard.cc:111-119: the spin bit is only ever
 * seen set by ANOTHER thread), so the restatement has no spin bits.
 */
#include "dint_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ---- a1: fasthash64 / fasthash_mix -- lock_2pl/udp/utils.h:20-57 ------------------------------ */
static inline uint64_t fh_mix(uint64_t h) {            /* utils.h:20-25 */
  h ^= h >> 23;
  h *= 0x2127599bf4325c37ULL;
  h ^= h >> 47;
  return h;
}
uint64_t dint_oracle_fasthash64(const void *buf, uint64_t len, uint64_t seed) { /* utils.h:27-57 */
  const uint64_t m = 0x880355f21e6d1965ULL;
  const unsigned char *p = (const unsigned char *)buf;
  uint64_t h = seed ^ (len * m);
  uint64_t nwords = len / 8;
  for (uint64_t i = 0; i < nwords; i++) {              /* utils.h:35-39 */
    uint64_t v;
    memcpy(&v, p + 8 * i, 8);
    h ^= fh_mix(v);
    h *= m;
  }
  const unsigned char *t = p + 8 * nwords;
  uint64_t rem = len & 7;
  if (rem) {                                           /* utils.h:44-54 (fall-through switch) */
    uint64_t v = 0;
    for (uint64_t i = 0; i < rem; i++) v ^= (uint64_t)t[i] << (8 * i);
    h ^= fh_mix(v);
    h *= m;
  }
  return fh_mix(h);
}
static inline uint64_t hash_u32(uint32_t x) { return dint_oracle_fasthash64(&x, 4, 0xdeadbeef); }
static inline uint64_t hash_u64(uint64_t x) { return dint_oracle_fasthash64(&x, 8, 0xdeadbeef); }

/* ---- a5: kvs -- store/udp/kvs.h:13-136 (tatp/udp/kvs.h, smallbank/udp/kvs.h same shape) -------- */
#define KEYS_PER_ENTRY 4                               /* kvs.h:11 */
typedef struct kvs_entry {                             /* kvs.h:13-19 */
  uint64_t key[KEYS_PER_ENTRY];
  uint8_t *val;                                        /* KEYS_PER_ENTRY * val_size bytes */
  uint32_t ver[KEYS_PER_ENTRY];
  uint8_t valid[KEYS_PER_ENTRY];
  struct kvs_entry *next;
} kvs_entry;
typedef struct {                                       /* kvs.h:21-25 */
  uint32_t hash_size;
  uint32_t val_size;
  kvs_entry **bucket_heads;
  uint64_t live;
} kvs;

static void kvs_init(kvs *t, uint32_t hash_size, uint32_t val_size) {  /* kvs.h:27-31 */
  t->hash_size = hash_size;
  t->val_size = val_size;
  t->bucket_heads = (kvs_entry **)calloc(hash_size ? hash_size : 1, sizeof(kvs_entry *));
  t->live = 0;
}
static void kvs_free(kvs *t) {
  if (!t->bucket_heads) return;
  for (uint32_t b = 0; b < t->hash_size; b++) {
    kvs_entry *e = t->bucket_heads[b];
    while (e) { kvs_entry *n = e->next; free(e->val); free(e); e = n; }
  }
  free(t->bucket_heads);
  t->bucket_heads = NULL;
}
static inline uint32_t kvs_hash(const kvs *t, uint64_t key) {          /* kvs.h:33-35 */
  return (uint32_t)(hash_u64(key) % (uint64_t)t->hash_size);
}
static int kvs_get(kvs *t, uint64_t key, uint8_t *val, uint32_t *ver) { /* kvs.h:37-55 */
  for (kvs_entry *e = t->bucket_heads[kvs_hash(t, key)]; e; e = e->next)
    for (int i = 0; i < KEYS_PER_ENTRY; i++)
      if (e->key[i] == key && e->valid[i]) {
        memcpy(val, e->val + (size_t)i * t->val_size, t->val_size);
        *ver = e->ver[i];
        return 0;
      }
  return 1;
}
static int kvs_set(kvs *t, uint64_t key, const uint8_t *val) {          /* kvs.h:57-75 */
  for (kvs_entry *e = t->bucket_heads[kvs_hash(t, key)]; e; e = e->next)
    for (int i = 0; i < KEYS_PER_ENTRY; i++)
      if (e->key[i] == key && e->valid[i]) {
        memcpy(e->val + (size_t)i * t->val_size, val, t->val_size);
        e->ver[i]++;
        return 0;
      }
  return 1;
}
static void kvs_insert(kvs *t, uint64_t key, const uint8_t *val) {      /* kvs.h:77-104 */
  uint32_t b = kvs_hash(t, key);
  for (kvs_entry *e = t->bucket_heads[b]; e; e = e->next)
    for (int i = 0; i < KEYS_PER_ENTRY; i++)
      if (!e->valid[i]) {                               /* first free slot anywhere in the chain */
        e->key[i] = key;
        memcpy(e->val + (size_t)i * t->val_size, val, t->val_size);
        e->ver[i] = 0;
        e->valid[i] = 1;
        t->live++;
        return;
      }
  kvs_entry *e = (kvs_entry *)calloc(1, sizeof(kvs_entry));   /* new head entry, slot 0 */
  e->val = (uint8_t *)calloc(KEYS_PER_ENTRY, t->val_size);
  e->key[0] = key;
  memcpy(e->val, val, t->val_size);
  e->ver[0] = 0;
  e->valid[0] = 1;
  e->next = t->bucket_heads[b];
  t->bucket_heads[b] = e;
  t->live++;
}
static int kvs_delete(kvs *t, uint64_t key) {                           /* kvs.h:106-136 */
  uint32_t b = kvs_hash(t, key);
  kvs_entry *prev = NULL;
  for (kvs_entry *e = t->bucket_heads[b]; e; prev = e, e = e->next)
    for (int i = 0; i < KEYS_PER_ENTRY; i++)
      if (e->key[i] == key && e->valid[i]) {
        e->valid[i] = 0;
        t->live--;
        int all_invalid = 1;
        for (int j = 0; j < KEYS_PER_ENTRY; j++) if (e->valid[j]) { all_invalid = 0; break; }
        if (all_invalid) {
          if (prev) prev->next = e->next; else t->bucket_heads[b] = e->next;
          free(e->val);
          free(e);
        }
        return 0;
      }
  return 1;                                             /* reference: panic("kvs_delete: key not found") */
}

/* ---- oracle object ------------------------------------------------------------------------------ */
#define TATP_TABLES 5
#define SB_TABLES 2
struct dint_oracle {
  int kind;
  dint_oracle_cfg cfg;
  /* lock_2pl: num_ex/num_sh (server.cc:37-40); fasst: locks/ver_table (server.cc:35-38) */
  uint32_t *a0, *a1;
  /* tatp: txn_locks[table][slot] (server_shard.cc:57); smallbank: txn num_ex/num_sh (server_shard.cc:51-57) */
  uint32_t *tl[TATP_TABLES], *tex[SB_TABLES], *tsh[SB_TABLES];
  uint32_t lock_mod[TATP_TABLES];                      /* kKeysPerEntry * hash_size (tatp.h:12-14) */
  kvs tables[TATP_TABLES];
  int n_tables;
  /* log ring 0 */
  uint8_t *ring;
  uint32_t entry_size, log_cnt;
  uint64_t log_total;
};

uint32_t dint_oracle_msg_size(int kind) {
  static const uint32_t sz[6] = {6, 9, 53, 53, 55, 23};
  return (kind >= 0 && kind < 6) ? sz[kind] : 0;
}

void dint_oracle_default_cfg(int kind, dint_oracle_cfg *c) {
  memset(c, 0, sizeof(*c));
  c->lock_slots = 36000000u;            /* lock_2pl/udp/utils.h:16, lock_fasst/udp/utils.h:16 */
  c->log_ring = 1000000u;               /* log_server/udp/utils.h:16, tatp/udp/kvs.h, smallbank/udp/kvs.h */
  c->subs_sizing = (kind == ORA_TATP) ? 7000000u : 2000000u;  /* tatp/udp/tatp.h:28, store/udp/tatp.h:10 */
  c->subs_populate = c->subs_sizing;
  c->accts_sizing = 24000000u;          /* smallbank/udp/smallbank.h:17 */
  c->accts_populate = c->accts_sizing;
}

dint_oracle *dint_oracle_create(int kind, const dint_oracle_cfg *cfg) {
  dint_oracle *o = (dint_oracle *)calloc(1, sizeof(*o));
  o->kind = kind;
  if (cfg) o->cfg = *cfg; else dint_oracle_default_cfg(kind, &o->cfg);
  const dint_oracle_cfg *c = &o->cfg;
  uint64_t S = c->subs_sizing, A = c->accts_sizing;
  switch (kind) {
    case ORA_LOCK2PL:
    case ORA_FASST:
      o->a0 = (uint32_t *)calloc(c->lock_slots, 4);
      o->a1 = (uint32_t *)calloc(c->lock_slots, 4);
      break;
    case ORA_LOG:
      o->entry_size = 56;               /* log_server/udp/utils.h:19-23: {u64 key; u8 val[40]; u32 ver} */
      break;
    case ORA_STORE:
      o->n_tables = 1;                  /* store/udp/server.cc:113: kSubscriberNum*18/kKeysPerEntry */
      kvs_init(&o->tables[0], (uint32_t)(S * 18 / KEYS_PER_ENTRY), 40);
      break;
    case ORA_TATP: {
      o->n_tables = TATP_TABLES;        /* tatp/udp/server_shard.cc:75-79 */
      uint32_t hs[TATP_TABLES] = {(uint32_t)(S * 3 / 2 / KEYS_PER_ENTRY), (uint32_t)(S * 3 / 2 / KEYS_PER_ENTRY),
                                  (uint32_t)(S * 15 / 4 / KEYS_PER_ENTRY), (uint32_t)(S * 15 / 4 / KEYS_PER_ENTRY),
                                  (uint32_t)(S * 45 / 8 / KEYS_PER_ENTRY)};
      for (int t = 0; t < TATP_TABLES; t++) {
        kvs_init(&o->tables[t], hs[t], 40);
        o->lock_mod[t] = KEYS_PER_ENTRY * hs[t];
        o->tl[t] = (uint32_t *)calloc(o->lock_mod[t] ? o->lock_mod[t] : 1, 4);
      }
      o->entry_size = 64;               /* tatp/udp/kvs.h:23-29 {u8 is_del; u8 table; u64 key; u8 val[40]; u32 ver} */
      break;
    }
    case ORA_SMALLBANK: {
      o->n_tables = SB_TABLES;          /* smallbank/udp/server_shard.cc:72-73 */
      uint32_t hs = (uint32_t)(A * 3 / 2 / KEYS_PER_ENTRY);
      for (int t = 0; t < SB_TABLES; t++) {
        kvs_init(&o->tables[t], hs, 8);
        o->lock_mod[t] = KEYS_PER_ENTRY * hs;
        o->tex[t] = (uint32_t *)calloc(o->lock_mod[t] ? o->lock_mod[t] : 1, 4);
        o->tsh[t] = (uint32_t *)calloc(o->lock_mod[t] ? o->lock_mod[t] : 1, 4);
      }
      o->entry_size = 32;               /* smallbank/udp/kvs.h:20-25 {u8 table; u64 key; u8 val[8]; u32 ver} */
      break;
    }
    default:
      free(o);
      return NULL;
  }
  if (o->entry_size) o->ring = (uint8_t *)calloc(c->log_ring ? c->log_ring : 1, o->entry_size);
  return o;
}

void dint_oracle_destroy(dint_oracle *o) {
  if (!o) return;
  free(o->a0); free(o->a1); free(o->ring);
  for (int t = 0; t < TATP_TABLES; t++) { free(o->tl[t]); kvs_free(&o->tables[t]); }
  for (int t = 0; t < SB_TABLES; t++) { free(o->tex[t]); free(o->tsh[t]); }
  free(o);
}

/* ---- a7/a9/a11: table population ---------------------------------------------------------------- */
static inline uint32_t fastrand(uint64_t *seed) {      /* store/udp/tatp.h:31-34, tatp/udp/tatp.h:32-35 */
  *seed = *seed * 1103515245ULL + 12345ULL;
  return (uint32_t)(*seed >> 32);
}

/* store/udp/tatp.h:45-66.  The reference leaves val.numberx[1..38] uninitialised (a stack struct);
 * in the oracle/_ref build those bytes read back as zero (checked by tests/test_oracle_golden.py), so
 * the restatement zero-fills them. */
static void populate_store(dint_oracle *o) {
  uint64_t seed = 0xdeadbeef;
  uint8_t val[40];
  for (uint32_t s_id = 0; s_id < o->cfg.subs_populate; s_id++)
    for (uint32_t sf = 1; sf <= 4; sf++)
      for (uint32_t st = 0; st <= 16; st += 8) {
        uint64_t key = (uint64_t)s_id | ((uint64_t)sf << 32) | ((uint64_t)st << 40);  /* store_key_t, tatp.h:14-23 */
        memset(val, 0, sizeof(val));
        val[0] = (uint8_t)((fastrand(&seed) % 24) + 1);  /* end_time */
        val[1] = 0x5a;                                    /* numberx[0] = kValMagic, tatp.h:12 */
        kvs_insert(&o->tables[0], key, val);
      }
}

/* tatp/udp/tatp.h:17-25 (map_1000) and :132-144 (tatp_sid_to_sub_nbr) */
static uint64_t tatp_sub_nbr(uint32_t s_id) {
  uint64_t r = 0;
  for (int g = 0; g < 3; g++) {
    uint32_t i = s_id % 1000;
    s_id /= 1000;
    uint64_t m = ((uint64_t)((i / 100) % 10) << 8) | ((uint64_t)((i / 10) % 10) << 4) | (uint64_t)(i % 10);
    r |= m << (12 * g);
  }
  return r;                               /* dec_9_10_11 = 0, unused = 0 */
}

/* tatp/udp/tatp.h:254-282 select_between_n_and_m_from (values = {1,2,3,4}) */
static int select_1_to_4(uint64_t *seed, uint8_t out[4]) {
  int used[32] = {0};
  int to_select = (int)(fastrand(seed) % 4) + 1;
  int cnt = 0;
  for (int i = 0; i < to_select; i++) {
    uint8_t value = (uint8_t)((fastrand(seed) % 4) + 1);
    if (used[value]) { i--; continue; }
    used[value] = 1;
    out[cnt++] = value;
  }
  return cnt;
}

/* tatp/udp/tatp.h:285-412.  Fields the reference never assigns (sub_nbr_unused, accinf data2..,
 * specfac error_cntl/data_a/data_b[1..], callfwd numberx[1..]) are zero here; see populate_store. */
static void populate_tatp(dint_oracle *o) {
  uint32_t N = o->cfg.subs_populate;
  uint8_t val[40];
  uint64_t seed = 0xdeadbeef;
  for (uint32_t s = 0; s < N; s++) {                   /* :285-311 subscriber */
    memset(val, 0, 40);
    uint64_t nbr = tatp_sub_nbr(s);
    memcpy(val, &nbr, 8);                              /* sub_nbr @0; sub_nbr_unused[7] @8 */
    for (int i = 0; i < 5; i++) val[15 + i] = (uint8_t)fastrand(&seed);   /* hex[5] @15 */
    for (int i = 0; i < 10; i++) val[20 + i] = (uint8_t)fastrand(&seed);  /* bytes[10] @20 */
    uint16_t bits = (uint16_t)fastrand(&seed);                             /* short bits @30 */
    memcpy(val + 30, &bits, 2);
    uint32_t msc = 97;                                                     /* msc_location @32 */
    memcpy(val + 32, &msc, 4);
    uint32_t vlr = fastrand(&seed);                                        /* vlr_location @36 */
    memcpy(val + 36, &vlr, 4);
    kvs_insert(&o->tables[0], (uint64_t)s, val);
  }
  for (uint32_t s = 0; s < N; s++) {                   /* :314-329 second subscriber */
    memset(val, 0, 40);
    memcpy(val, &s, 4);
    val[4] = 98;
    kvs_insert(&o->tables[1], tatp_sub_nbr(s), val);
  }
  seed = 0xdeadbeef;
  for (uint32_t s = 0; s < N; s++) {                   /* :332-357 access info */
    uint8_t types[4];
    int n = select_1_to_4(&seed, types);
    for (int i = 0; i < n; i++) {
      memset(val, 0, 40);
      val[0] = 99;                                     /* data1 */
      kvs_insert(&o->tables[2], (uint64_t)s | ((uint64_t)types[i] << 32), val);
    }
  }
  seed = 0xdeadbeef;
  for (uint32_t s = 0; s < N; s++) {                   /* :360-412 special facility + call forwarding */
    uint8_t types[4];
    int n = select_1_to_4(&seed, types);
    for (int i = 0; i < n; i++) {
      uint64_t sf = types[i];
      memset(val, 0, 40);
      val[3] = 100;                                    /* data_b[0] @3 */
      val[0] = (fastrand(&seed) % 100 < 85) ? 1 : 0;   /* is_active @0 */
      kvs_insert(&o->tables[3], (uint64_t)s | (sf << 32), val);
      for (uint64_t st = 0; st <= 16; st += 8) {
        if (fastrand(&seed) % 2 == 0) continue;
        memset(val, 0, 40);
        val[1] = 101;                                  /* numberx[0] @1 */
        val[0] = (uint8_t)((fastrand(&seed) % 24) + 1);  /* end_time @0 */
        kvs_insert(&o->tables[4], (uint64_t)s | (sf << 32) | (st << 40), val);
      }
    }
  }
}

/* smallbank/udp/smallbank.h:105-127 */
static void populate_smallbank(dint_oracle *o) {
  for (uint32_t a = 0; a < o->cfg.accts_populate; a++) {
    uint8_t val[8];
    uint32_t magic = 97;
    float bal = 1000000000.0f;                         /* sav_val.bal = 1000000000ull */
    memcpy(val, &magic, 4);
    memcpy(val + 4, &bal, 4);
    kvs_insert(&o->tables[0], (uint64_t)a, val);
    magic = 98;
    memcpy(val, &magic, 4);
    kvs_insert(&o->tables[1], (uint64_t)a, val);
  }
}

void dint_oracle_populate(dint_oracle *o) {
  if (o->kind == ORA_STORE) populate_store(o);
  else if (o->kind == ORA_TATP) populate_tatp(o);
  else if (o->kind == ORA_SMALLBANK) populate_smallbank(o);
}

/* ---- handlers ----------------------------------------------------------------------------------- */
static inline uint32_t rd32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint64_t rd64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void wr32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }

/* a2: lock_2pl/udp/server.cc:70-122.  msg = {action@0, lid@1, type@5} */
static int step_lock2pl(dint_oracle *o, uint8_t *m) {
  uint32_t s = (uint32_t)(hash_u32(rd32(m + 1)) % (uint64_t)o->cfg.lock_slots);   /* :71-72 */
  uint32_t *num_ex = o->a0, *num_sh = o->a1;
  if (m[0] == 0) {                       /* kAcquireLock :82 */
    if (m[5] == 0) {                     /* kShared :83-93 */
      if (num_ex[s] == 0) { num_sh[s]++; m[0] = 2; } else m[0] = 3;
    } else if (m[5] == 1) {              /* kExclusive :96-107 */
      if (num_ex[s] == 0 && num_sh[s] == 0) { num_ex[s]++; m[0] = 2; } else m[0] = 3;
    } else return -1;                    /* panic("invalid lock type") :109 */
  } else if (m[0] == 1) {                /* kReleaseLock :112-119 */
    if (m[5] == 0) num_sh[s]--;
    else if (m[5] == 1) num_ex[s]--;
    m[0] = 5;                            /* kReleaseAck (any other type: no counter change) */
  } else return -1;                      /* panic("invalid action") :121 */
  return 0;
}

/* a3: lock_fasst/udp/server.cc:78-119.  msg = {type@0, lid@1, ver@5} */
static int step_fasst(dint_oracle *o, uint8_t *m) {
  uint32_t s = (uint32_t)(hash_u32(rd32(m + 1)) % (uint64_t)o->cfg.lock_slots);   /* :81-82 */
  uint32_t *locks = o->a0, *ver = o->a1;
  switch (m[0]) {
    case 0: m[0] = 4; wr32(m + 5, ver[s]); break;                       /* kRead :86-90 */
    case 1: if (locks[s] == 0) { locks[s] = 1; m[0] = 5; } else m[0] = 6; break;  /* kAcquireLock :92-101 */
    case 2: if (locks[s] == 1) locks[s] = 0; m[0] = 7; break;           /* kAbort :103-107 */
    case 3: ver[s]++; if (locks[s] == 1) locks[s] = 0; m[0] = 8; break;  /* kCommit :109-114 */
    default: return -1;                                                 /* :116-117 */
  }
  return 0;
}

/* a4: log_server/udp/server.cc:73-88.  msg = {type@0, key@1, val@9, ver@49} */
static int step_log(dint_oracle *o, uint8_t *m) {
  if (m[0] != 0) return -1;              /* :76-77 */
  uint8_t *e = o->ring + (size_t)o->log_cnt * 56;
  memcpy(e, m + 1, 8);                   /* key */
  memcpy(e + 8, m + 9, 40);              /* val */
  memcpy(e + 48, m + 49, 4);             /* ver */
  o->log_cnt = (o->log_cnt + 1) % o->cfg.log_ring;   /* :84 */
  o->log_total++;
  m[0] = 1;                              /* kAck :86 */
  return 0;
}

/* a6: store/udp/server.cc:75-97.  msg = {type@0, key@1, val@9, ver@49} */
static int step_store(dint_oracle *o, uint8_t *m) {
  uint64_t key = rd64(m + 1);
  if (m[0] == 0) {                       /* kRead :79-84 */
    uint32_t ver = rd32(m + 49);
    int rc = kvs_get(&o->tables[0], key, m + 9, &ver);
    wr32(m + 49, ver);
    m[0] = rc == 0 ? 3 : 7;              /* kGrantRead / kNotExist */
  } else if (m[0] == 1) {                /* kSet :86-91 */
    int rc = kvs_set(&o->tables[0], key, m + 9);
    m[0] = rc == 0 ? 5 : 7;              /* kSetAck / kNotExist */
  } else return -1;                      /* :93-94 */
  return 0;
}

static void log_append_tatp(dint_oracle *o, const uint8_t *m, int is_del) {
  /* tatp/udp/server_shard.cc:182-207; entry {is_del@0, table@1, key@8, val@16, ver@56} */
  uint8_t *e = o->ring + (size_t)o->log_cnt * 64;
  e[0] = (uint8_t)is_del;
  e[1] = m[2];
  memcpy(e + 8, m + 3, 8);
  if (!is_del) memcpy(e + 16, m + 11, 40);             /* kDeleteLog leaves val untouched :196-203 */
  memcpy(e + 56, m + 51, 4);
  o->log_cnt = (o->log_cnt + 1) % o->cfg.log_ring;
  o->log_total++;
}

/* a8: tatp/udp/server_shard.cc:113-210.  msg = {ord@0, type@1, table@2, key@3, val@11, ver@51} */
static int step_tatp(dint_oracle *o, uint8_t *m) {
  uint8_t table = m[2];
  uint8_t type = m[1];
  if (type > 27 || table >= TATP_TABLES) return -1;
  uint64_t key = rd64(m + 3);
  kvs *t = &o->tables[table];
  uint32_t ls = (uint32_t)(hash_u64(key) % (uint64_t)o->lock_mod[table]);   /* lock_hash, tatp.h:12-14 */
  uint32_t *lk = &o->tl[table][ls];
  switch (type) {
    case 0: {                            /* kRead :116-121 */
      uint32_t ver = rd32(m + 51);
      int rc = kvs_get(t, key, m + 11, &ver);
      wr32(m + 51, ver);
      m[1] = rc == 0 ? 4 : 6;            /* kGrantRead / kNotExist */
      break;
    }
    case 1: if (*lk == 0) { *lk = 1; m[1] = 7; } else m[1] = 8; break;      /* kAcquireLock :123-132 */
    case 2: if (*lk == 1) *lk = 0; m[1] = 9; break;                         /* kAbort :134-138 */
    case 12:                             /* kCommitPrim :140-146 */
      if (kvs_set(t, key, m + 11)) return -1;          /* tatp/udp/kvs.h:91 panic */
      if (*lk == 1) *lk = 0;
      m[1] = 15;
      break;
    case 18:                             /* kInsertPrim :148-154 */
      kvs_insert(t, key, m + 11);
      if (*lk == 1) *lk = 0;
      m[1] = 20;
      break;
    case 22:                             /* kDeletePrim :156-162 */
      if (kvs_delete(t, key)) return -1;               /* tatp/udp/kvs.h:152 panic */
      if (*lk == 1) *lk = 0;
      m[1] = 25;
      break;
    case 13: if (kvs_set(t, key, m + 11)) return -1; m[1] = 16; break;      /* kCommitBck :164-168 */
    case 19: kvs_insert(t, key, m + 11); m[1] = 21; break;                  /* kInsertBck :170-174 */
    case 23: if (kvs_delete(t, key)) return -1; m[1] = 26; break;           /* kDeleteBck :176-180 */
    case 14: log_append_tatp(o, m, 0); m[1] = 17; break;                    /* kCommitLog :182-194 */
    case 24: log_append_tatp(o, m, 1); m[1] = 27; break;                    /* kDeleteLog :196-207 */
    default: return -1;                  /* :209 */
  }
  return 0;
}

/* a10: smallbank/udp/server_shard.cc:107-189.  msg = {ord@0, type@1, table@2, key@3, val@11, ver@19} */
static int step_smallbank(dint_oracle *o, uint8_t *m) {
  uint8_t table = m[2];
  uint8_t type = m[1];
  if (table >= SB_TABLES) return -1;
  uint64_t key = rd64(m + 3);
  kvs *t = &o->tables[table];
  uint32_t lh = (uint32_t)(hash_u64(key) % (uint64_t)o->lock_mod[table]);   /* :109, smallbank.h:12-14 */
  uint32_t *ex = &o->tex[table][lh], *sh = &o->tsh[table][lh];
  switch (type) {
    case 0:                              /* kAcquireShared :121-133 */
      if (*ex == 0) {
        uint32_t ver = rd32(m + 19);
        (*sh)++;
        if (kvs_get(t, key, m + 11, &ver)) return -1;   /* smallbank/udp/kvs.h:67 panic */
        wr32(m + 19, ver);
        m[1] = 7;
      } else m[1] = 8;
      break;
    case 1:                              /* kAcquireExclusive :135-147 */
      if (*ex == 0 && *sh == 0) {
        uint32_t ver = rd32(m + 19);
        (*ex)++;
        if (kvs_get(t, key, m + 11, &ver)) return -1;
        wr32(m + 19, ver);
        m[1] = 9;
      } else m[1] = 10;
      break;
    case 2: (*sh)--; m[1] = 11; break;   /* kReleaseShared :149-154 */
    case 3: (*ex)--; m[1] = 12; break;   /* kReleaseExclusive :156-161 */
    case 4: if (kvs_set(t, key, m + 11)) return -1; m[1] = 13; break;   /* kCommitPrim :163-167 */
    case 5: if (kvs_set(t, key, m + 11)) return -1; m[1] = 14; break;   /* kCommitBck :169-173 */
    case 6: {                            /* kCommitLog :175-186; entry {table@0, key@8, val@16, ver@24} */
      uint8_t *e = o->ring + (size_t)o->log_cnt * 32;
      e[0] = table;
      memcpy(e + 8, m + 3, 8);
      memcpy(e + 16, m + 11, 8);
      memcpy(e + 24, m + 19, 4);
      o->log_cnt = (o->log_cnt + 1) % o->cfg.log_ring;
      o->log_total++;
      m[1] = 15;
      break;
    }
    default: return -1;                  /* :188 */
  }
  return 0;
}

int64_t dint_oracle_process(dint_oracle *o, const void *req, uint64_t n, void *resp) {
  uint32_t sz = dint_oracle_msg_size(o->kind);
  const uint8_t *in = (const uint8_t *)req;
  uint8_t *out = (uint8_t *)resp;
  for (uint64_t i = 0; i < n; i++) {
    uint8_t *m = out + i * sz;
    if (m != in + i * sz) memcpy(m, in + i * sz, sz);   /* the reply is the request buffer, mutated */
    int rc;
    switch (o->kind) {
      case ORA_LOCK2PL: rc = step_lock2pl(o, m); break;
      case ORA_FASST: rc = step_fasst(o, m); break;
      case ORA_LOG: rc = step_log(o, m); break;
      case ORA_STORE: rc = step_store(o, m); break;
      case ORA_TATP: rc = step_tatp(o, m); break;
      default: rc = step_smallbank(o, m); break;
    }
    if (rc) return -(int64_t)(i + 1);
  }
  return 0;
}

/* ---- state inspection --------------------------------------------------------------------------- */
int dint_oracle_kv_get(dint_oracle *o, int table, uint64_t key, uint8_t *val, uint32_t *ver) {
  if (table < 0 || table >= o->n_tables) return -1;
  return kvs_get(&o->tables[table], key, val, ver);
}
uint64_t dint_oracle_kv_count(dint_oracle *o, int table) {
  return (table >= 0 && table < o->n_tables) ? o->tables[table].live : 0;
}
uint32_t dint_oracle_lock_slot(dint_oracle *o, int table, uint64_t k) {
  if (o->kind == ORA_LOCK2PL || o->kind == ORA_FASST)
    return (uint32_t)(hash_u32((uint32_t)k) % (uint64_t)o->cfg.lock_slots);
  if ((o->kind == ORA_TATP && table < TATP_TABLES) || (o->kind == ORA_SMALLBANK && table < SB_TABLES))
    return (uint32_t)(hash_u64(k) % (uint64_t)o->lock_mod[table]);
  return 0;
}
void dint_oracle_lock_state(dint_oracle *o, int table, uint32_t slot, uint32_t out[2]) {
  out[0] = out[1] = 0;
  if (o->kind == ORA_LOCK2PL || o->kind == ORA_FASST) { out[0] = o->a0[slot]; out[1] = o->a1[slot]; }
  else if (o->kind == ORA_TATP) out[0] = o->tl[table][slot];
  else if (o->kind == ORA_SMALLBANK) { out[0] = o->tex[table][slot]; out[1] = o->tsh[table][slot]; }
}
uint64_t dint_oracle_log_appended(dint_oracle *o) { return o->log_total; }
uint32_t dint_oracle_log_entry_size(dint_oracle *o) { return o->entry_size; }
const void *dint_oracle_log_ring(dint_oracle *o) { return o->ring; }
