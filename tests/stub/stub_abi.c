/*
 * stub_abi.c -- TEST INFRASTRUCTURE.  A stand-in for libdint_b200.so that exports the few C-ABI entry points the
 * UDP front-end (dint_b200/csrc/udp_server.cc) calls and answers them with the CPU oracle.  Built into a scratch
 * directory by tests/test_udp_front_end_cpu.py and put in front of the real library with LD_LIBRARY_PATH, so that
 * the front-end's own logic -- recvmmsg batching, arrival order, reply addressing, datagram-size filter,
 * shutdown -- can be checked without a GPU.  Never shipped, never on the product path.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/dint_b200.h"
#include "../../oracle/dint_oracle.h"

struct dint_engine { dint_oracle *o; int kind; };
static const char *g_err = "";

uint32_t dint_msg_size(int kind) { return dint_oracle_msg_size(kind); }
const char *dint_last_error(void) { return g_err; }
void dint_default_cfg(int kind, dint_cfg *cfg) {
  memset(cfg, 0, sizeof *cfg);
  dint_oracle_cfg oc;
  dint_oracle_default_cfg(kind, &oc);
  cfg->lock_slots = oc.lock_slots; cfg->log_ring = oc.log_ring; cfg->subs_sizing = oc.subs_sizing;
  cfg->subs_populate = oc.subs_populate; cfg->accts_sizing = oc.accts_sizing; cfg->accts_populate = oc.accts_populate;
  cfg->n_shards = 1;
}
int dint_create(int kind, const dint_cfg *cfg, int device, dint_engine **out) {
  (void)device;
  dint_oracle_cfg oc;
  dint_oracle_default_cfg(kind, &oc);
  if (cfg) {
    oc.lock_slots = cfg->lock_slots; oc.log_ring = cfg->log_ring; oc.subs_sizing = cfg->subs_sizing;
    oc.subs_populate = cfg->subs_populate; oc.accts_sizing = cfg->accts_sizing; oc.accts_populate = cfg->accts_populate;
  }
  const char *small = getenv("DINT_STUB_SMALL");              /* tests: tiny populations */
  if (small) { oc.subs_populate = (uint32_t)atoi(small); oc.accts_populate = (uint32_t)atoi(small); }
  dint_engine *e = (dint_engine *)calloc(1, sizeof *e);
  e->o = dint_oracle_create(kind, &oc);
  e->kind = kind;
  if (!e->o) { free(e); g_err = "oracle create failed"; return DINT_ENOMEM; }
  *out = e;
  return DINT_OK;
}
void dint_destroy(dint_engine *e) { if (e) { dint_oracle_destroy(e->o); free(e); } }
int dint_populate(dint_engine *e) { dint_oracle_populate(e->o); return DINT_OK; }
int dint_submit(dint_engine *e, const void *req, uint64_t n, void *resp) {
  return dint_oracle_process(e->o, req, n, resp) == 0 ? DINT_OK : DINT_EPROTO;
}
void *dint_host_alloc(size_t bytes) { return malloc(bytes ? bytes : 1); }
void dint_host_free(void *p) { free(p); }

/* ---- dint_cluster_*: lock kinds / store = ONE sequential oracle (what the cluster must equal); tatp / smallbank =
 * n_gpus oracle shard servers, each fed its records in index order ---- */
struct dint_cluster { int kind, G; dint_oracle *o[8]; };
int dint_cluster_create(int kind, const dint_cfg *cfg, int n_gpus, const int *devices, uint64_t max_batch, dint_cluster **out) {
  (void)devices; (void)max_batch;
  if (n_gpus < 1 || n_gpus > 8) return DINT_EINVAL;
  dint_cluster *c = (dint_cluster *)calloc(1, sizeof *c);
  c->kind = kind;
  c->G = n_gpus;
  const int by_dst = kind == DINT_TATP || kind == DINT_SMALLBANK;
  for (int s = 0; s < (by_dst ? n_gpus : 1); s++) {
    dint_engine *e = NULL;
    if (dint_create(kind, cfg, 0, &e) != DINT_OK) return DINT_ENOMEM;
    c->o[s] = e->o;
    free(e);
  }
  *out = c;
  return DINT_OK;
}
int dint_cluster_populate(dint_cluster *c) {
  for (int s = 0; s < 8; s++) if (c->o[s]) dint_oracle_populate(c->o[s]);
  return DINT_OK;
}
int dint_cluster_submit(dint_cluster *c, const void *req, uint64_t n, const uint8_t *dst, void *resp) {
  const uint32_t msg = dint_oracle_msg_size(c->kind);
  if (!c->o[1]) return dint_oracle_process(c->o[0], req, n, resp) == 0 ? DINT_OK : DINT_EPROTO;
  int rc = DINT_OK;
  for (uint64_t i = 0; i < n; i++) {              /* one record at a time keeps every shard's order = index order */
    const uint8_t s = dst ? dst[i] : 0;
    if (s >= c->G || dint_oracle_process(c->o[s], (const uint8_t *)req + i * msg, 1, (uint8_t *)resp + i * msg) != 0) rc = DINT_EPROTO;
  }
  return rc;
}
void dint_cluster_destroy(dint_cluster *c) {
  if (!c) return;
  for (int s = 0; s < 8; s++) if (c->o[s]) dint_oracle_destroy(c->o[s]);
  free(c);
}
