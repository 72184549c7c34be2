/*
 * udp_blast.c -- TEST INFRASTRUCTURE (CPU-baseline side), never linked into the product.
 *
 * Multi-socket UDP replayer for the "reference UDP server as shipped" baseline.
 * T threads, each with its own connected socket (its own source port, so SO_REUSEPORT spreads the threads over
 * the server's sockets like the reference's client machines do), replay a slice of a flat trace of packed wire
 * structs against 127.0.0.1:<port> in windows of W datagrams: send W, receive W (100 ms timeout per datagram, a
 * lost one is counted and skipped).  Replies are counted, not compared: arrival order across sockets is not
 * deterministic (the bit-exact oracle is the replay shim, not this).
 *
 * usage: udp_blast <trace.bin> <msg_size> <port> <threads> <window> <seconds>   -> one JSON line on stdout
 */
#define _GNU_SOURCE
#include <arpa/inet.h>
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/socket.h>
#include <sys/stat.h>
#include <sys/time.h>
#include <time.h>
#include <unistd.h>

static const uint8_t *g_trace;
static size_t g_n, g_msg;
static int g_port, g_threads, g_window;
static double g_seconds;
static atomic_ullong g_ok, g_lost;
static atomic_int g_stop;

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + ts.tv_nsec * 1e-9;
}

static void *worker(void *arg) {
  const long id = (long)arg;
  int fd = socket(AF_INET, SOCK_DGRAM, 0);
  if (fd < 0) return NULL;
  struct sockaddr_in srv;
  memset(&srv, 0, sizeof srv);
  srv.sin_family = AF_INET;
  srv.sin_port = htons((unsigned short)g_port);
  srv.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
  if (connect(fd, (struct sockaddr *)&srv, sizeof srv) < 0) { close(fd); return NULL; }
  struct timeval tv = {0, 100000};
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
  const size_t per = g_n / (size_t)g_threads, lo = per * (size_t)id;
  uint8_t buf[256];
  unsigned long long ok = 0, lost = 0;
  size_t i = 0;
  while (!atomic_load_explicit(&g_stop, memory_order_relaxed)) {
    int sent = 0;
    for (int w = 0; w < g_window; w++, i = (i + 1) % per)
      if (send(fd, g_trace + (lo + i) * g_msg, g_msg, 0) == (ssize_t)g_msg) sent++;
    for (int w = 0; w < sent; w++) {
      ssize_t r = recv(fd, buf, sizeof buf, 0);
      if (r == (ssize_t)g_msg) ok++;
      else { lost += (unsigned long long)(sent - w); break; }       /* timeout: give up on this window */
    }
  }
  atomic_fetch_add(&g_ok, ok);
  atomic_fetch_add(&g_lost, lost);
  close(fd);
  return NULL;
}

int main(int argc, char **argv) {
  if (argc < 7) { fprintf(stderr, "usage: %s trace msg_size port threads window seconds\n", argv[0]); return 2; }
  g_msg = (size_t)atoi(argv[2]); g_port = atoi(argv[3]); g_threads = atoi(argv[4]); g_window = atoi(argv[5]); g_seconds = atof(argv[6]);
  int fd = open(argv[1], O_RDONLY);
  struct stat st;
  if (fd < 0 || fstat(fd, &st) < 0 || g_msg == 0 || g_msg > 255 || g_threads < 1 || g_window < 1) { perror("trace"); return 2; }
  g_trace = mmap(NULL, (size_t)st.st_size, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
  if (g_trace == MAP_FAILED) { perror("mmap"); return 2; }
  g_n = (size_t)st.st_size / g_msg;
  if (g_n / (size_t)g_threads == 0) { fprintf(stderr, "trace too short\n"); return 2; }
  pthread_t *th = calloc((size_t)g_threads, sizeof *th);
  const double t0 = now();
  for (long i = 0; i < g_threads; i++) pthread_create(&th[i], NULL, worker, (void *)i);
  while (now() - t0 < g_seconds) usleep(20000);
  atomic_store(&g_stop, 1);
  for (long i = 0; i < g_threads; i++) pthread_join(th[i], NULL);
  const double dt = now() - t0;
  printf("{\"replies\": %llu, \"lost\": %llu, \"seconds\": %.4f, \"req_per_s\": %.1f, \"client_threads\": %d, \"window\": %d}\n",
         (unsigned long long)g_ok, (unsigned long long)g_lost, dt, (double)g_ok / dt, g_threads, g_window);
  return 0;
}
