/* cluster_check.c -- TEST.  A C caller of the multi-GPU boundary (include/dint_b200.h: dint_cluster_*), checked
 * against the oracle (oracle/dint_oracle.h).  No Python, no torch: what a C/C++ transport front-end links.
 *   lock_fasst: 3 shards (slot % 3) must answer like ONE sequential server (lock_fasst/udp/server.cc:78-119)
 *   smallbank : 3 shard servers, the CLIENT names the shard (smallbank/caladan/client_udp_shard.cc:441-577)
 * Shards live on device 0 unless DINT_CHECK_DEVICES=a,b,c names one device per shard. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "dint_b200.h"
#include "dint_oracle.h"

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd(void) { rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng_state >> 33); }

static int fail(const char* what) { fprintf(stdout, "cluster_check: FAILED: %s (%s)\n", what, dint_last_error()); return 1; }

int main(void) {
  int devs[3] = {0, 0, 0};
  const char* dv = getenv("DINT_CHECK_DEVICES");
  if (dv) sscanf(dv, "%d,%d,%d", &devs[0], &devs[1], &devs[2]);
  /* ---- lock_fasst ---- */
  {
    const uint64_t n = 200000;
    uint8_t* req = (uint8_t*)dint_host_alloc(n * 9);
    uint8_t* got = (uint8_t*)dint_host_alloc(n * 9);
    uint8_t* want = (uint8_t*)malloc(n * 9);
    if (!req || !got || !want) return fail("alloc");
    dint_cluster* cl = NULL;
    if (dint_cluster_create(DINT_FASST, NULL, 3, devs, 16384, &cl) != DINT_OK) return fail("dint_cluster_create(lock_fasst)");
    dint_oracle_cfg oc;
    dint_oracle_default_cfg(ORA_FASST, &oc);
    dint_oracle* ora = dint_oracle_create(ORA_FASST, &oc);
    for (int pass = 0; pass < 3; pass++) {
      for (uint64_t i = 0; i < n; i++) {
        uint8_t* r = req + i * 9;
        const uint32_t lid = rnd() % 20000u, ver = rnd();
        r[0] = (uint8_t)(rnd() % 4u);
        memcpy(r + 1, &lid, 4);
        memcpy(r + 5, &ver, 4);
      }
      if (dint_oracle_process(ora, req, n, want) != 0) return fail("oracle");
      if (dint_cluster_submit(cl, req, n, NULL, got) != DINT_OK) return fail("dint_cluster_submit(lock_fasst)");
      if (memcmp(got, want, n * 9) != 0) return fail("lock_fasst replies differ from ONE sequential server");
    }
    dint_cluster_destroy(cl);
    dint_oracle_destroy(ora);
    dint_host_free(req); dint_host_free(got); free(want);
    printf("cluster_check: lock_fasst x3 shards == one server (3 x %llu requests)\n", (unsigned long long)n);
  }
  /* ---- smallbank, client-chosen shards ---- */
  {
    const uint64_t n = 60000;
    const uint32_t accts = 2000;
    uint8_t* req = (uint8_t*)dint_host_alloc(n * 23);
    uint8_t* got = (uint8_t*)dint_host_alloc(n * 23);
    uint8_t* dst = (uint8_t*)malloc(n);
    uint8_t* want = (uint8_t*)malloc(n * 23);
    uint8_t* part = (uint8_t*)malloc(n * 23);
    uint8_t* pres = (uint8_t*)malloc(n * 23);
    dint_cfg cfg;
    dint_default_cfg(DINT_SMALLBANK, &cfg);
    cfg.accts_populate = accts;
    dint_cluster* cl = NULL;
    if (dint_cluster_create(DINT_SMALLBANK, &cfg, 3, devs, 8192, &cl) != DINT_OK) return fail("dint_cluster_create(smallbank)");
    if (dint_cluster_populate(cl) != DINT_OK) return fail("dint_cluster_populate");
    dint_oracle* ora[3];
    for (int s = 0; s < 3; s++) {
      dint_oracle_cfg oc;
      dint_oracle_default_cfg(ORA_SMALLBANK, &oc);
      oc.accts_populate = accts;
      ora[s] = dint_oracle_create(ORA_SMALLBANK, &oc);
      dint_oracle_populate(ora[s]);
    }
    for (uint64_t i = 0; i < n; i++) {          /* acquire / release / commit / log traffic on existing accounts */
      uint8_t* r = req + i * 23;
      static const uint8_t types[7] = {0, 1, 2, 3, 4, 5, 6};
      const uint64_t key = rnd() % accts;
      memset(r, 0, 23);
      r[0] = (uint8_t)i;
      r[1] = types[rnd() % 7u];
      r[2] = (uint8_t)(rnd() % 2u);
      memcpy(r + 3, &key, 8);
      for (int b = 0; b < 8; b++) r[11 + b] = (uint8_t)rnd();
      dst[i] = (uint8_t)((key + rnd() % 3u) % 3u);   /* primary or one of its backups */
    }
    for (int s = 0; s < 3; s++) {               /* expectation: every shard server sees its records in index order */
      uint64_t m = 0;
      for (uint64_t i = 0; i < n; i++) if (dst[i] == s) memcpy(part + (m++) * 23, req + i * 23, 23);
      if (dint_oracle_process(ora[s], part, m, pres) != 0) return fail("oracle(smallbank)");
      m = 0;
      for (uint64_t i = 0; i < n; i++) if (dst[i] == s) memcpy(want + i * 23, pres + (m++) * 23, 23);
    }
    if (dint_cluster_submit(cl, req, n, dst, got) != DINT_OK) return fail("dint_cluster_submit(smallbank)");
    if (memcmp(got, want, n * 23) != 0) return fail("smallbank replies differ from three shard servers");
    dint_cluster_destroy(cl);
    for (int s = 0; s < 3; s++) dint_oracle_destroy(ora[s]);
    printf("cluster_check: smallbank x3 client-chosen shards == three shard servers (%llu requests)\n", (unsigned long long)n);
  }
  printf("cluster_check: OK\n");
  return 0;
}
