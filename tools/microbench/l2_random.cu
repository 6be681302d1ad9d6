// Latency of a batch of independent random 16-byte loads from an L2-resident table, at the occupancy of
// k_correspond (about 8-17 warps per SM).  Calibrates the kNN probe stage.
#include <cstdio>
#include <cuda_runtime.h>
template <int B> __global__ void probe(const uint4* tab, unsigned mask, int rounds, unsigned long long* cyc, unsigned* sink) {
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    uint4 e[B];
#pragma unroll
    for (int j = 0; j < B; ++j) { x = x * 1664525u + 1013904223u; e[j] = __ldg(&tab[(x >> 4) & mask]); }
#pragma unroll
    for (int j = 0; j < B; ++j) acc += e[j].x ^ e[j].w;
    x ^= acc;   // next round depends on this one
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) atomicAdd(cyc, (unsigned long long)(t1 - t0));
  if (acc == 0x12345) sink[0] = acc;
}
template <int B> void run(const uint4* tab, unsigned mask, int blocks, const char* what) {
  unsigned long long* c; unsigned* s; cudaMalloc(&c, 8); cudaMalloc(&s, 4); 
  probe<B><<<blocks, 128>>>(tab, mask, 8, c, s); cudaMemset(c, 0, 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a); probe<B><<<blocks, 128>>>(tab, mask, 8, c, s); cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, a, b);
  unsigned long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  printf("%-34s batch=%d blocks=%4d: %7.0f cycles per round (thread 0 avg), kernel %.1f us\n", what, B, blocks, (double)h / blocks / 8, ms * 1e3);
}
int main() {
  for (size_t mb : {22, 256}) {
    size_t n = mb * 1024 * 1024 / 16; unsigned mask = 1; while (mask * 2 <= n) mask *= 2; mask -= 1;
    uint4* tab; cudaMalloc(&tab, (size_t)(mask + 1) * 16); cudaMemset(tab, 1, (size_t)(mask + 1) * 16);
    char w[64]; snprintf(w, 64, "%zu MB table", (size_t)(mask + 1) * 16 >> 20);
    run<1>(tab, mask, 315, w); run<7>(tab, mask, 315, w); run<7>(tab, mask, 630, w); run<9>(tab, mask, 315, w); run<27>(tab, mask, 315, w);
    cudaFree(tab);
  }
  return 0;
}
