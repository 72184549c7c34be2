// ubench2.cu -- development microbenchmark: throughput of the random-access primitives K1/K2 are made of.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33; return x; }
enum { OP_LD4 = 0, OP_RED = 1, OP_ATOM = 2, OP_ST4 = 3 };
template <int OP>
__global__ void k(uint32_t* tbl, uint64_t mask, uint32_t* out, uint32_t per_thread) {
  uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint32_t acc = 0;
  for (uint32_t it = 0; it < per_thread; it++) {
    uint64_t e = mix(tid * 1315423911ULL + it * 2654435761ULL) & mask;
    if (OP == OP_LD4) acc += __ldcg(tbl + e);
    else if (OP == OP_RED) atomicOr(tbl + e, 1u << (tid & 31));
    else if (OP == OP_ATOM) acc += atomicOr(tbl + e, 1u << (tid & 31));
    else tbl[e] = (uint32_t)tid;
  }
  if (acc == 0x12345678) out[0] = acc;
}
template <int OP>
void run(const char* name, uint32_t* tbl, uint64_t words, int blocks, int threads, uint32_t per_thread, uint32_t* out) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<OP><<<blocks, threads>>>(tbl, words - 1, out, 2);
  cudaEventRecord(a);
  k<OP><<<blocks, threads>>>(tbl, words - 1, out, per_thread);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  double n = (double)blocks * threads * per_thread;
  printf("%-10s table=%6.0f MB  ops=%.1fM  %8.1f us  %7.1f G ops/s\n", name, words * 4 / 1048576.0, n / 1e6, ms * 1e3, n / ms / 1e6);
}
int main() {
  size_t bytes = 1ULL << 30;
  uint32_t* tbl; cudaMalloc(&tbl, bytes); cudaMemset(tbl, 0, bytes);
  uint32_t* out; cudaMalloc(&out, 4);
  for (size_t mb : {4, 16, 64, 144, 512}) {
    size_t words = 1; while (words * 4 < mb * 1048576) words <<= 1;   // pow2 >= mb
    for (int per : {1, 4}) {
      int blocks = (1 << 20) / 256 / per * 4;   // 4M ops total
      run<OP_LD4>("ld4", tbl, words, blocks, 256, per, out);
      run<OP_RED>("red", tbl, words, blocks, 256, per, out);
      run<OP_ATOM>("atom", tbl, words, blocks, 256, per, out);
      run<OP_ST4>("st4", tbl, words, blocks, 256, per, out);
    }
  }
  return 0;
}
