// PCIe probe: H2D / D2H / bidirectional bandwidth from pinned host memory bound to each NUMA node.
// build: nvcc -O2 -o pcie_probe pcie_probe.cu     run: ./pcie_probe
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#define MPOL_BIND 2
static long mbind_(void* a, unsigned long len, int mode, const unsigned long* mask, unsigned long maxnode, unsigned flags) {
  return syscall(SYS_mbind, a, len, mode, mask, maxnode, flags);
}
static double now() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
  int dev = 0; cudaSetDevice(dev);
  char pci[64]; cudaDeviceGetPCIBusId(pci, sizeof pci, dev);
  for (char* p = pci; *p; p++) *p = tolower(*p);
  char path[256]; snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/numa_node", pci);
  int gpu_node = -1; if (FILE* f = fopen(path, "r")) { if (fscanf(f, "%d", &gpu_node) != 1) gpu_node = -1; fclose(f); }
  int nodes = 0; for (;; nodes++) { snprintf(path, sizeof path, "/sys/devices/system/node/node%d", nodes); if (access(path, F_OK)) break; }
  printf("gpu %s numa_node=%d host nodes=%d cpus=%ld\n", pci, gpu_node, nodes, sysconf(_SC_NPROCESSORS_ONLN));
  const size_t B = 9437184 * 4;   // 4 x (1M x 9 bytes)
  void *d_in, *d_out; cudaMalloc(&d_in, B); cudaMalloc(&d_out, B);
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  for (int node = -1; node < nodes; node++) {
    void* h = mmap(nullptr, 2 * B, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (node >= 0) { unsigned long mask[16] = {0}; mask[node / 64] |= 1ul << (node % 64); if (mbind_(h, 2 * B, MPOL_BIND, mask, 1024, 0)) perror("mbind"); }
    memset(h, 1, 2 * B);
    if (cudaHostRegister(h, 2 * B, cudaHostRegisterDefault) != cudaSuccess) { printf("register failed\n"); return 1; }
    char* hi = (char*)h; char* ho = hi + B;
    for (size_t sz : {B / 16, B / 4, B}) {
      double best[3] = {1e30, 1e30, 1e30};
      for (int rep = 0; rep < 12; rep++) {
        cudaDeviceSynchronize(); double t0 = now();
        cudaMemcpyAsync(d_in, hi, sz, cudaMemcpyHostToDevice, s1); cudaStreamSynchronize(s1);
        double t1 = now();
        cudaMemcpyAsync(ho, d_out, sz, cudaMemcpyDeviceToHost, s2); cudaStreamSynchronize(s2);
        double t2 = now();
        cudaMemcpyAsync(d_in, hi, sz, cudaMemcpyHostToDevice, s1);
        cudaMemcpyAsync(ho, d_out, sz, cudaMemcpyDeviceToHost, s2);
        cudaStreamSynchronize(s1); cudaStreamSynchronize(s2);
        double t3 = now();
        if (t1 - t0 < best[0]) best[0] = t1 - t0;
        if (t2 - t1 < best[1]) best[1] = t2 - t1;
        if (t3 - t2 < best[2]) best[2] = t3 - t2;
      }
      printf("node %2d size %8.2f MB: H2D %6.1f GB/s  D2H %6.1f GB/s  both %6.1f GB/s per direction (%.0f us)\n", node, sz / 1e6,
             sz / best[0] / 1e3, sz / best[1] / 1e3, sz / best[2] / 1e3, best[2]);
    }
    cudaHostUnregister(h); munmap(h, 2 * B);
  }
  return 0;
}
