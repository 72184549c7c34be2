// dint_udp_server -- the reference's UDP server, with the per-datagram handler replaced by the batched engine.
//
// What it replaces: `<bench>/udp/server*.cc` of the reference: N pinned threads, each
// `recvfrom -> fasthash64 -> switch(type) -> sendto` on its own SO_REUSEPORT socket (lock_fasst/udp/server.cc:40-119,
// tatp/udp/server_shard.cc:213-260).  Same wire protocol (one packed `struct message` per datagram, the reply is
// the request buffer mutated and sent back to the source address), same default port (20230), so the reference's
// clients (`client_udp*`) talk to it unchanged.
//
// Shape: ONE receive thread batches datagrams with recvmmsg into pinned memory; ONE submit thread hands each
// batch to dint_submit() (H2D, kernels, D2H inside) and answers with sendmmsg.  Two batch buffers ping-pong
// between the threads, so the network stack fills batch k+1 while the GPU serves batch k.  Requests are served in
// arrival order: request i of a batch sees the effects of every earlier request, exactly what ONE reference
// thread would have produced for that arrival order.
//
// Host C++ above the C ABI only (include/dint_b200.h): no CUDA here, no oracle, no CPU fallback -- without a
// GPU dint_create() fails and the server exits.
//
// usage: dint_udp_server <lock_2pl|lock_fasst|log_server|store|tatp|smallbank> [--port P] [--bind A.B.C.D]
//                        [--batch N] [--device D] [--shard-id I --shards G] [--linger-us U]
#include <arpa/inet.h>
#include <netinet/in.h>
#include <signal.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

extern "C" {
#include "../../include/dint_b200.h"
}

namespace {

std::atomic<bool> g_stop{false};
void on_signal(int) { g_stop.store(true); }

struct Batch {
  uint8_t* req = nullptr;               // pinned, n * msg
  uint8_t* resp = nullptr;              // pinned
  std::vector<mmsghdr> hdr;
  std::vector<iovec> iov;
  std::vector<sockaddr_in> peer;
  int n = 0;                            // datagrams in the batch
  enum { EMPTY, FULL } state = EMPTY;
};

struct Stats {
  std::atomic<uint64_t> datagrams{0}, batches{0}, dropped{0}, bad_records{0};
};

int kind_of(const std::string& s) {
  static const char* names[] = {"lock_2pl", "lock_fasst", "log_server", "store", "tatp", "smallbank"};
  for (int k = 0; k < 6; k++)
    if (s == names[k]) return k;
  return -1;
}

}  // namespace

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <lock_2pl|lock_fasst|log_server|store|tatp|smallbank> [--port P] [--bind ADDR] [--batch N] "
                    "[--device D] [--shards G --shard-id I] [--linger-us U]\n", argv[0]);
    return 2;
  }
  const int kind = kind_of(argv[1]);
  if (kind < 0) { fprintf(stderr, "unknown server kind '%s'\n", argv[1]); return 2; }
  int port = 20230, device = 0, linger_us = 50;
  unsigned batch_max = 65536, shards = 1, shard_id = 0;
  std::string bind_addr = "0.0.0.0";
  for (int i = 2; i + 1 < argc; i += 2) {
    const std::string a = argv[i];
    const char* v = argv[i + 1];
    if (a == "--port") port = atoi(v);
    else if (a == "--bind") bind_addr = v;
    else if (a == "--batch") batch_max = (unsigned)atoi(v);
    else if (a == "--device") device = atoi(v);
    else if (a == "--shards") shards = (unsigned)atoi(v);
    else if (a == "--shard-id") shard_id = (unsigned)atoi(v);
    else if (a == "--linger-us") linger_us = atoi(v);
    else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
  }
  if (batch_max < 1) batch_max = 1;
  const uint32_t msg = dint_msg_size(kind);

  // ---- engine: the state the reference keeps in its global arrays lives on the GPU ----
  dint_cfg cfg;
  dint_default_cfg(kind, &cfg);                       // kLockHashSize, table sizes, ring length of the reference
  if (kind == DINT_TATP || kind == DINT_SMALLBANK) {  // server_shard <id>: the CLIENT picks the shard; each holds its replicas
    cfg.txn_shards = shards;
    cfg.txn_shard_id = shard_id;
  }
  dint_engine* eng = nullptr;
  if (dint_create(kind, &cfg, device, &eng) != DINT_OK) {
    fprintf(stderr, "dint_udp_server: dint_create failed: %s\n", dint_last_error());
    return 1;
  }
  if (dint_populate(eng) != DINT_OK) {                // kvs_init + populate_* of the reference (no-op for lock / log)
    fprintf(stderr, "dint_udp_server: dint_populate failed: %s\n", dint_last_error());
    return 1;
  }

  // ---- socket: as the reference sets it up (SO_REUSEPORT kept so that several front-ends may share a port) ----
  const int fd = socket(AF_INET, SOCK_DGRAM, 0);
  if (fd < 0) { perror("socket"); return 1; }
  int one = 1, buf = 64 << 20;
  setsockopt(fd, SOL_SOCKET, SO_REUSEPORT, &one, sizeof one);
  setsockopt(fd, SOL_SOCKET, SO_RCVBUF, &buf, sizeof buf);
  setsockopt(fd, SOL_SOCKET, SO_SNDBUF, &buf, sizeof buf);
  timeval tv{0, 200000};                              // wake up 5x a second to notice a stop request
  setsockopt(fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
  sockaddr_in srv{};
  srv.sin_family = AF_INET;
  srv.sin_port = htons((uint16_t)port);
  if (inet_pton(AF_INET, bind_addr.c_str(), &srv.sin_addr) != 1) { fprintf(stderr, "bad --bind address\n"); return 2; }
  if (bind(fd, (sockaddr*)&srv, sizeof srv) < 0) { perror("bind"); return 1; }

  // ---- two batch buffers ping-pong between the receive and the submit thread ----
  Batch b[2];
  for (Batch& x : b) {
    x.req = (uint8_t*)dint_host_alloc((size_t)batch_max * msg);
    x.resp = (uint8_t*)dint_host_alloc((size_t)batch_max * msg);
    if (!x.req || !x.resp) { fprintf(stderr, "pinned allocation failed\n"); return 1; }
    x.hdr.resize(batch_max);
    x.iov.resize(batch_max);
    x.peer.resize(batch_max);
  }
  std::mutex mu;
  std::condition_variable cv;
  Stats st;
  signal(SIGINT, on_signal);
  signal(SIGTERM, on_signal);
  fprintf(stderr, "dint_udp_server: %s on %s:%d, device %d, batches of <= %u datagrams of %u bytes\n", argv[1], bind_addr.c_str(),
          port, device, batch_max, msg);

  std::thread submitter([&] {
    for (int k = 0;; k ^= 1) {
      Batch& x = b[k];
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return x.state == Batch::FULL || g_stop.load(); });
        if (x.state != Batch::FULL) return;
      }
      // replaces the switch(type) of the reference's server_loop for the whole batch, in arrival order
      const int rc = dint_submit(eng, x.req, (uint64_t)x.n, x.resp);
      if (rc != DINT_OK && rc != DINT_EPROTO) {       // DINT_EPROTO: malformed datagrams were answered with type 0xFF
        fprintf(stderr, "dint_udp_server: dint_submit failed: %s\n", dint_last_error());
        g_stop.store(true);
      } else {
        if (rc == DINT_EPROTO) st.bad_records++;
        for (int i = 0; i < x.n; i++) {
          x.iov[i].iov_base = x.resp + (size_t)i * msg;
          x.hdr[i].msg_hdr.msg_namelen = sizeof(sockaddr_in);
        }
        for (int sent = 0; sent < x.n;) {               // replaces net_send: the reply goes back to the source address
          const int r = sendmmsg(fd, x.hdr.data() + sent, (unsigned)(x.n - sent), 0);
          if (r <= 0) { st.dropped += (uint64_t)(x.n - sent); break; }
          sent += r;
        }
        st.datagrams += (uint64_t)x.n;
        st.batches++;
      }
      {
        std::lock_guard<std::mutex> lk(mu);
        x.state = Batch::EMPTY;
      }
      cv.notify_all();
    }
  });

  // ---- receive loop (replaces net_recv): block for the first datagram, then drain what has queued up ----
  for (int k = 0; !g_stop.load(); ) {
    Batch& x = b[k];
    {
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return x.state == Batch::EMPTY || g_stop.load(); });
      if (g_stop.load()) break;
    }
    for (unsigned i = 0; i < batch_max; i++) {
      x.iov[i] = {x.req + (size_t)i * msg, msg};
      x.hdr[i].msg_hdr = {&x.peer[i], sizeof(sockaddr_in), &x.iov[i], 1, nullptr, 0, 0};
      x.hdr[i].msg_len = 0;
    }
    int n = recvmmsg(fd, x.hdr.data(), batch_max, MSG_WAITFORONE, nullptr);
    if (n <= 0) continue;                             // timeout: look at the stop flag again
    const auto t0 = std::chrono::steady_clock::now();
    while ((unsigned)n < batch_max) {                 // keep draining for a short linger so that load builds batches
      const int m = recvmmsg(fd, x.hdr.data() + n, batch_max - (unsigned)n, MSG_DONTWAIT, nullptr);
      if (m > 0) { n += m; continue; }
      if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(linger_us)) break;
    }
    // a datagram of the wrong size cannot be a request of this server: compact it away
    int keep = 0;
    for (int i = 0; i < n; i++) {
      if (x.hdr[i].msg_len != msg) { st.dropped++; continue; }
      if (keep != i) {
        memcpy(x.req + (size_t)keep * msg, x.req + (size_t)i * msg, msg);
        x.peer[keep] = x.peer[i];
      }
      keep++;
    }
    if (keep == 0) continue;
    for (int i = 0; i < keep; i++) x.hdr[i].msg_hdr.msg_name = &x.peer[i];
    {
      std::lock_guard<std::mutex> lk(mu);
      x.n = keep;
      x.state = Batch::FULL;
    }
    cv.notify_all();
    k ^= 1;
  }
  g_stop.store(true);
  cv.notify_all();
  submitter.join();
  fprintf(stderr, "dint_udp_server: %llu datagrams in %llu batches (%.1f per batch), %llu dropped, %llu batches with malformed records\n",
          (unsigned long long)st.datagrams.load(), (unsigned long long)st.batches.load(),
          st.batches.load() ? (double)st.datagrams.load() / (double)st.batches.load() : 0.0, (unsigned long long)st.dropped.load(),
          (unsigned long long)st.bad_records.load());
  close(fd);
  for (Batch& x : b) { dint_host_free(x.req); dint_host_free(x.resp); }
  dint_destroy(eng);
  return 0;
}
