/*
 * replay_shim.c -- TEST INFRASTRUCTURE (oracle side), never linked into the product.
 *
 * LD_PRELOAD interposer that lets the UNMODIFIED reference UDP servers
 * (/root/reference/<bench>/udp/server*.cc, compiled in place by oracle/Makefile
 * into oracle/_ref/) run without a network: the datagram a server thread would
 * have received with recvfrom() is taken from a flat binary trace of packed wire
 * structs, and the reply it hands to sendto() is appended to an output file.
 * With `server 1` the reference handler therefore processes request 0..n-1
 * strictly in order -- that run IS the bit-exact oracle.
 *
 * Interposed: socket, setsockopt, bind, recvfrom, sendto, sched_getcpu,
 *             pthread_setaffinity_np (only when DINT_SHIM_SPREAD=1).
 *
 * Why sched_getcpu: the log-writing servers pick their ring as (cpu-3)/2
 * (log_server/udp/server.cc:79-80, tatp/udp/server_shard.cc:183-184,
 * smallbank/udp/server_shard.cc:176-177) and the worker thread starts before
 * main() pins it (log_server/udp/server.cc:110-117), so an unpinned first request
 * could index ring -1.  We answer 2*tid+3, i.e. exactly what the reference's own
 * pinning (2*i+3) would have produced.
 *
 * Environment:
 *   DINT_TRACE    path of the request trace (n * msg_size bytes)            [required]
 *   DINT_OUT      path for the response stream (n * msg_size), optional
 *   DINT_STATS    path for a one-line JSON with timing, optional
 *   DINT_THREADS  number of server threads that will call recvfrom (default 1)
 *   DINT_REPEAT   replay the trace this many times (timing runs; default 1)
 *   DINT_SHIM_SPREAD=1  re-pin worker i to core i%ncores instead of (2i+3)%ncores
 *
 * Multi-thread mode (DINT_THREADS>1) is for the CPU-baseline timing only: threads
 * grab blocks of 1024 consecutive requests from a shared cursor, so ordering
 * between blocks is not deterministic and the output is not compared.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <fcntl.h>
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/types.h>
#include <netinet/in.h>
#include <time.h>
#include <unistd.h>

#define SHIM_PORT 20230
#define BLOCK 1024

static pthread_once_t g_once = PTHREAD_ONCE_INIT;
static const uint8_t *g_trace;
static size_t g_trace_bytes;
static uint8_t *g_out;
static int g_threads = 1;
static long g_repeat = 1;
static int g_spread = 0;
static atomic_long g_cursor;       /* next block start (in requests, across repeats) */
static atomic_int g_arrived;       /* threads that reached their first recvfrom */
static atomic_int g_done;          /* threads that hit EOF */
static atomic_int g_next_tid;
static atomic_int g_pin_seq;
static struct timespec g_t0;
static atomic_int g_t0_set;
static int g_fd_is_trace[4096];

static __thread int t_tid = -1;
static __thread long t_pos = 0, t_end = 0;   /* current block [pos,end) in requests */
static __thread long t_served = 0;

static void die(const char *m) { fprintf(stderr, "replay_shim: %s\n", m); _exit(97); }

static void shim_init(void) {
  const char *tp = getenv("DINT_TRACE");
  if (!tp) die("DINT_TRACE not set");
  int fd = open(tp, O_RDONLY);
  if (fd < 0) die("cannot open DINT_TRACE");
  struct stat st;
  fstat(fd, &st);
  g_trace_bytes = (size_t)st.st_size;
  if (g_trace_bytes) {
    g_trace = mmap(NULL, g_trace_bytes, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    if (g_trace == MAP_FAILED) die("mmap trace failed");
  }
  close(fd);
  const char *op = getenv("DINT_OUT");
  if (op && g_trace_bytes) {
    int ofd = open(op, O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (ofd < 0) die("cannot open DINT_OUT");
    if (ftruncate(ofd, (off_t)g_trace_bytes) != 0) die("ftruncate failed");
    g_out = mmap(NULL, g_trace_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, ofd, 0);
    if (g_out == MAP_FAILED) die("mmap out failed");
    close(ofd);
  } else if (op) {
    int ofd = open(op, O_RDWR | O_CREAT | O_TRUNC, 0644);
    if (ofd >= 0) close(ofd);
  }
  if (getenv("DINT_THREADS")) g_threads = atoi(getenv("DINT_THREADS"));
  if (g_threads < 1) g_threads = 1;
  if (getenv("DINT_REPEAT")) g_repeat = atol(getenv("DINT_REPEAT"));
  if (g_repeat < 1) g_repeat = 1;
  if (getenv("DINT_SHIM_SPREAD")) g_spread = atoi(getenv("DINT_SHIM_SPREAD"));
}

static void finish_and_exit(long total_reqs, size_t msg) {
  struct timespec t1;
  clock_gettime(CLOCK_MONOTONIC, &t1);
  double sec = (double)(t1.tv_sec - g_t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - g_t0.tv_nsec);
  if (g_out) msync(g_out, g_trace_bytes, MS_SYNC);
  const char *sp = getenv("DINT_STATS");
  if (sp) {
    FILE *f = fopen(sp, "w");
    if (f) {
      fprintf(f, "{\"requests\": %ld, \"seconds\": %.9f, \"req_per_s\": %.3f, \"threads\": %d, \"msg_size\": %zu}\n",
              total_reqs, sec, sec > 0 ? (double)total_reqs / sec : 0.0, g_threads, msg);
      fclose(f);
    }
  }
  fflush(NULL);
  _exit(0);
}

int socket(int domain, int type, int protocol) {
  (void)domain; (void)type; (void)protocol;
  pthread_once(&g_once, shim_init);
  int fd = open("/dev/null", O_RDWR);
  if (fd >= 0 && fd < 4096) g_fd_is_trace[fd] = 0;
  return fd;
}

int setsockopt(int fd, int level, int optname, const void *optval, socklen_t optlen) {
  (void)fd; (void)level; (void)optname; (void)optval; (void)optlen;
  return 0;
}

int bind(int fd, const struct sockaddr *addr, socklen_t len) {
  (void)len;
  const struct sockaddr_in *in = (const struct sockaddr_in *)addr;
  if (fd >= 0 && fd < 4096) g_fd_is_trace[fd] = (ntohs(in->sin_port) == SHIM_PORT);
  return 0;
}

ssize_t recvfrom(int fd, void *buf, size_t len, int flags, struct sockaddr *src, socklen_t *alen) {
  (void)flags;
  if (fd < 0 || fd >= 4096 || !g_fd_is_trace[fd]) {   /* e.g. the CPU-monitor socket :20231 */
    for (;;) pause();
  }
  size_t msg = len;
  long n_req = (long)(g_trace_bytes / msg);
  long total = n_req * g_repeat;
  if (t_tid < 0) {
    t_tid = atomic_fetch_add(&g_next_tid, 1);
    atomic_fetch_add(&g_arrived, 1);
    while (atomic_load(&g_arrived) < g_threads) sched_yield();
    if (g_threads > 1) usleep(20000);   /* let main() finish pinning the last worker */
    if (!atomic_exchange(&g_t0_set, 1)) clock_gettime(CLOCK_MONOTONIC, &g_t0);
  }
  if (t_pos == t_end) {
    long blk = (g_threads == 1) ? total : BLOCK;
    long start = atomic_fetch_add(&g_cursor, blk);
    if (start >= total) {
      int d = atomic_fetch_add(&g_done, 1) + 1;
      if (d == g_threads) finish_and_exit(total, msg);
      for (;;) pause();
    }
    t_pos = start;
    t_end = start + blk < total ? start + blk : total;
  }
  long i = t_pos % n_req;
  memcpy(buf, g_trace + (size_t)i * msg, msg);
  if (src && alen && *alen >= sizeof(struct sockaddr_in)) {
    memset(src, 0, sizeof(struct sockaddr_in));
    ((struct sockaddr_in *)src)->sin_family = AF_INET;
    *alen = sizeof(struct sockaddr_in);
  }
  t_served++;
  return (ssize_t)msg;
}

ssize_t sendto(int fd, const void *buf, size_t len, int flags, const struct sockaddr *dst, socklen_t alen) {
  (void)flags; (void)dst; (void)alen;
  if (fd >= 0 && fd < 4096 && g_fd_is_trace[fd]) {
    long n_req = (long)(g_trace_bytes / len);
    long i = t_pos % n_req;           /* reply to the request handed out by the last recvfrom */
    if (g_out) memcpy(g_out + (size_t)i * len, buf, len);
    t_pos++;
  }
  return (ssize_t)len;
}

int sched_getcpu(void) {
  int tid = t_tid < 0 ? 0 : t_tid;
  return 2 * (tid % 16) + 3;
}

int pthread_setaffinity_np(pthread_t th, size_t sz, const cpu_set_t *set) {
  static int (*real)(pthread_t, size_t, const cpu_set_t *);
  if (!real) real = dlsym(RTLD_NEXT, "pthread_setaffinity_np");
  if (!g_spread) { real(th, sz, set); return 0; }   /* a restricted cpuset may refuse: run unpinned */
  long nc = sysconf(_SC_NPROCESSORS_ONLN);
  int i = atomic_fetch_add(&g_pin_seq, 1);
  cpu_set_t s;
  CPU_ZERO(&s);
  CPU_SET((int)(i % nc), &s);
  real(th, sizeof(s), &s);
  return 0;
}
