// dint_udp_server -- the reference's UDP server, with the per-datagram handler replaced by the batched engine.
//
// What it replaces: `<bench>/udp/server*.cc` of the reference: N pinned threads, each
// `recvfrom -> fasthash64 -> switch(type) -> sendto` on its own SO_REUSEPORT socket (lock_fasst/udp/server.cc:40-119,
// tatp/udp/server_shard.cc:213-260).  Same wire protocol (one packed `struct message` per datagram, the reply is
// the request buffer mutated and sent back to the source address), same default port (20230), so the reference's
// clients (`client_udp*`) talk to it unchanged.
//
// Shape: R socket threads (SO_REUSEPORT, like the reference's N threads) each alternate
//     recvmmsg (a batch) -> hand the batch to the engine thread -> sendmmsg (the replies),
// and ONE engine thread gathers whatever batches are ready, in socket order, into one pinned array, makes ONE
// dint_submit() call for all of them (H2D, kernels, D2H inside) and hands every socket thread its slice of the
// replies.  While the GPU serves one gathering, the other sockets keep receiving: that is the pipeline.
// Order: datagrams of one socket (one client 4-tuple always hashes to the same socket) keep their arrival
// order inside the array, and request i of a submit sees the effects of every earlier one -- what ONE reference
// thread would have produced for that arrival order.  (Across sockets the reference has no order either.)
//
// Host C++ above the C ABI only (include/dint_b200.h): no CUDA here, no oracle, no CPU fallback -- without a
// GPU dint_create() fails and the server exits.
//
// usage: dint_udp_server <lock_2pl|lock_fasst|log_server|store|tatp|smallbank> [--port P] [--bind A.B.C.D]
//                        [--sockets R] [--batch N] [--device D] [--shard-id I --shards G] [--linger-us U]
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

// one socket thread's batch: filled by recvmmsg, answered in place, sent back with sendmmsg
struct Worker {
  int fd = -1;
  std::vector<uint8_t> buf;             // n * msg: requests in, replies out
  std::vector<mmsghdr> hdr;
  std::vector<iovec> iov;
  std::vector<sockaddr_in> peer;
  int n = 0;
  enum State { FILLING, READY, IN_FLIGHT, DONE } state = FILLING;   // guarded by the shared mutex
  uint64_t datagrams = 0, dropped = 0;
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
    fprintf(stderr, "usage: %s <lock_2pl|lock_fasst|log_server|store|tatp|smallbank> [--port P] [--bind ADDR] [--sockets R] "
                    "[--batch N] [--device D] [--shards G --shard-id I] [--linger-us U]\n", argv[0]);
    return 2;
  }
  const int kind = kind_of(argv[1]);
  if (kind < 0) { fprintf(stderr, "unknown server kind '%s'\n", argv[1]); return 2; }
  int port = 20230, device = 0, linger_us = 50;
  unsigned batch_max = 16384, shards = 1, shard_id = 0;
  unsigned n_sock = std::thread::hardware_concurrency() / 2;
  if (n_sock < 1) n_sock = 1;
  if (n_sock > 8) n_sock = 8;                         // the reference runs `server 8` (exp/run_lock_fasst.sh)
  std::string bind_addr = "0.0.0.0";
  for (int i = 2; i + 1 < argc; i += 2) {
    const std::string a = argv[i];
    const char* v = argv[i + 1];
    if (a == "--port") port = atoi(v);
    else if (a == "--bind") bind_addr = v;
    else if (a == "--batch") batch_max = (unsigned)atoi(v);
    else if (a == "--sockets") n_sock = (unsigned)atoi(v);
    else if (a == "--device") device = atoi(v);
    else if (a == "--shards") shards = (unsigned)atoi(v);
    else if (a == "--shard-id") shard_id = (unsigned)atoi(v);
    else if (a == "--linger-us") linger_us = atoi(v);
    else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
  }
  if (batch_max < 1) batch_max = 1;
  if (n_sock < 1) n_sock = 1;
  if (n_sock > 64) n_sock = 64;
  sockaddr_in srv{};
  srv.sin_family = AF_INET;
  srv.sin_port = htons((uint16_t)port);
  if (inet_pton(AF_INET, bind_addr.c_str(), &srv.sin_addr) != 1) { fprintf(stderr, "bad --bind address\n"); return 2; }
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

  // ---- sockets: as the reference sets them up, one per thread, all bound to the same port ----
  std::vector<Worker> w(n_sock);
  for (Worker& x : w) {
    x.fd = socket(AF_INET, SOCK_DGRAM, 0);
    if (x.fd < 0) { perror("socket"); return 1; }
    int one = 1, buf = 32 << 20;
    setsockopt(x.fd, SOL_SOCKET, SO_REUSEPORT, &one, sizeof one);
    setsockopt(x.fd, SOL_SOCKET, SO_RCVBUF, &buf, sizeof buf);
    setsockopt(x.fd, SOL_SOCKET, SO_SNDBUF, &buf, sizeof buf);
    timeval tv{0, 200000};                            // wake up 5x a second to notice a stop request
    setsockopt(x.fd, SOL_SOCKET, SO_RCVTIMEO, &tv, sizeof tv);
    if (bind(x.fd, (sockaddr*)&srv, sizeof srv) < 0) { perror("bind"); return 1; }
    x.buf.resize((size_t)batch_max * msg);
    x.hdr.resize(batch_max);
    x.iov.resize(batch_max);
    x.peer.resize(batch_max);
  }
  // the array one dint_submit() sees: every ready batch, back to back (pinned: the engine copies from / to it)
  const size_t arr_cap = (size_t)batch_max * n_sock;
  uint8_t* req = (uint8_t*)dint_host_alloc(arr_cap * msg);
  uint8_t* resp = (uint8_t*)dint_host_alloc(arr_cap * msg);
  if (!req || !resp) { fprintf(stderr, "pinned allocation failed\n"); return 1; }

  std::mutex mu;
  std::condition_variable cv_engine, cv_workers;
  std::atomic<uint64_t> submits{0}, bad_batches{0};
  signal(SIGINT, on_signal);
  signal(SIGTERM, on_signal);
  fprintf(stderr, "dint_udp_server: %s on %s:%d, device %d, %u sockets, batches of <= %u datagrams of %u bytes\n", argv[1],
          bind_addr.c_str(), port, device, n_sock, batch_max, msg);

  auto socket_thread = [&](Worker& x) {
    while (!g_stop.load()) {
      // ---- receive (replaces net_recv): block for the first datagram, then drain what has queued up ----
      for (unsigned i = 0; i < batch_max; i++) {
        x.iov[i] = {x.buf.data() + (size_t)i * msg, msg};
        x.hdr[i].msg_hdr = {&x.peer[i], sizeof(sockaddr_in), &x.iov[i], 1, nullptr, 0, 0};
        x.hdr[i].msg_len = 0;
      }
      int n = recvmmsg(x.fd, x.hdr.data(), batch_max, MSG_WAITFORONE, nullptr);
      if (n <= 0) continue;                           // timeout: look at the stop flag again
      const auto t0 = std::chrono::steady_clock::now();
      while ((unsigned)n < batch_max) {               // keep draining for a short linger so that load builds batches
        const int m = recvmmsg(x.fd, x.hdr.data() + n, batch_max - (unsigned)n, MSG_DONTWAIT, nullptr);
        if (m > 0) { n += m; continue; }
        if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(linger_us)) break;
      }
      int keep = 0;                                   // a datagram of the wrong size cannot be a request of this server
      for (int i = 0; i < n; i++) {
        if (x.hdr[i].msg_len != msg) { x.dropped++; continue; }
        if (keep != i) {
          memcpy(x.buf.data() + (size_t)keep * msg, x.buf.data() + (size_t)i * msg, msg);
          x.peer[keep] = x.peer[i];
        }
        keep++;
      }
      if (keep == 0) continue;
      // ---- hand the batch to the engine thread and wait for the replies (they come back in x.buf) ----
      {
        std::unique_lock<std::mutex> lk(mu);
        x.n = keep;
        x.state = Worker::READY;
        cv_engine.notify_one();
        // (bounded waits: the signal handler only stores the flag, so a wakeup may be missed -- never a hang)
        while (!cv_workers.wait_for(lk, std::chrono::milliseconds(100), [&] { return x.state == Worker::DONE || g_stop.load(); })) {}
        if (x.state != Worker::DONE) return;
        x.state = Worker::FILLING;
      }
      // ---- send (replaces net_send): every reply goes back to the address its request came from ----
      for (int i = 0; i < keep; i++) {
        x.iov[i].iov_base = x.buf.data() + (size_t)i * msg;
        x.hdr[i].msg_hdr.msg_name = &x.peer[i];
        x.hdr[i].msg_hdr.msg_namelen = sizeof(sockaddr_in);
      }
      for (int sent = 0; sent < keep;) {
        const int r = sendmmsg(x.fd, x.hdr.data() + sent, (unsigned)(keep - sent), 0);
        if (r <= 0) { x.dropped += (uint64_t)(keep - sent); break; }
        sent += r;
      }
      x.datagrams += (uint64_t)keep;
    }
  };

  std::vector<std::thread> threads;
  for (Worker& x : w) threads.emplace_back(socket_thread, std::ref(x));

  // ---- engine thread (this one): gather the ready batches, one dint_submit, scatter the replies ----
  std::vector<int> taken;
  while (!g_stop.load()) {
    taken.clear();
    size_t n = 0;
    {
      std::unique_lock<std::mutex> lk(mu);
      cv_engine.wait_for(lk, std::chrono::milliseconds(200), [&] {
        for (const Worker& x : w)
          if (x.state == Worker::READY) return true;
        return g_stop.load();
      });
      for (unsigned i = 0; i < n_sock; i++)           // socket order: deterministic for a given set of ready batches
        if (w[i].state == Worker::READY) { w[i].state = Worker::IN_FLIGHT; taken.push_back((int)i); }
    }
    if (taken.empty()) continue;
    for (int i : taken) {                             // the batches back to back: ONE array for the engine
      memcpy(req + n * msg, w[i].buf.data(), (size_t)w[i].n * msg);
      n += (size_t)w[i].n;
    }
    // replaces the switch(type) of the reference's server_loop for every gathered datagram, in array order
    const int rc = dint_submit(eng, req, (uint64_t)n, resp);
    if (rc != DINT_OK && rc != DINT_EPROTO) {         // DINT_EPROTO: malformed records were answered with type 0xFF
      fprintf(stderr, "dint_udp_server: dint_submit failed: %s\n", dint_last_error());
      g_stop.store(true);
    }
    if (rc == DINT_EPROTO) bad_batches++;
    submits++;
    size_t off = 0;
    for (int i : taken) {
      memcpy(w[i].buf.data(), resp + off * msg, (size_t)w[i].n * msg);
      off += (size_t)w[i].n;
    }
    {
      std::lock_guard<std::mutex> lk(mu);
      for (int i : taken) w[i].state = Worker::DONE;
    }
    cv_workers.notify_all();
  }
  {
    std::lock_guard<std::mutex> lk(mu);               // under the mutex: no worker sits between its predicate and its wait
    g_stop.store(true);
  }
  cv_workers.notify_all();
  for (std::thread& t : threads) t.join();
  uint64_t total = 0, dropped = 0;
  for (Worker& x : w) { total += x.datagrams; dropped += x.dropped; close(x.fd); }
  fprintf(stderr, "dint_udp_server: %llu datagrams in %llu submits (%.1f per submit), %llu dropped, %llu submits with malformed records\n",
          (unsigned long long)total, (unsigned long long)submits.load(), submits.load() ? (double)total / (double)submits.load() : 0.0,
          (unsigned long long)dropped, (unsigned long long)bad_batches.load());
  dint_host_free(req);
  dint_host_free(resp);
  dint_destroy(eng);
  return 0;
}
