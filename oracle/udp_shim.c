/*
 * udp_shim.c -- TEST INFRASTRUCTURE (CPU-baseline side), never linked into the product.
 *
 * LD_PRELOAD interposer for the "reference UDP server as shipped" baseline: the
 * UNMODIFIED reference servers bind 10.10.1.<shard>:20230 (e.g. lock_fasst/udp/server.cc:45-50, 69-70), an
 * address this box does not have.  Only bind() is interposed: a 10.10.1.x address becomes 127.0.0.1 (port
 * DINT_UDP_PORT if set, so that concurrent runs do not collide).  Sockets, recvfrom and sendto stay real: every
 * request pays the two syscalls it pays in the reference deployment.
 */
#define _GNU_SOURCE
#include <arpa/inet.h>
#include <dlfcn.h>
#include <netinet/in.h>
#include <stdlib.h>
#include <string.h>
#include <sys/socket.h>

int bind(int fd, const struct sockaddr *addr, socklen_t len) {
  static int (*real_bind)(int, const struct sockaddr *, socklen_t);
  if (!real_bind) real_bind = (int (*)(int, const struct sockaddr *, socklen_t))dlsym(RTLD_NEXT, "bind");
  if (addr && addr->sa_family == AF_INET && len >= sizeof(struct sockaddr_in)) {
    struct sockaddr_in a;
    memcpy(&a, addr, sizeof a);
    if ((ntohl(a.sin_addr.s_addr) & 0xffffff00u) == 0x0a0a0100u) {        /* 10.10.1.x */
      a.sin_addr.s_addr = htonl(INADDR_LOOPBACK);
      const char *p = getenv("DINT_UDP_PORT");
      if (p) a.sin_port = htons((unsigned short)atoi(p));
      return real_bind(fd, (const struct sockaddr *)&a, sizeof a);
    }
  }
  return real_bind(fd, addr, len);
}
