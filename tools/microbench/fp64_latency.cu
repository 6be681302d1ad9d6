// Dependent-chain latency of FP64 ops on one thread (B200 calibration for the solver / fit code).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
template <int OP> __global__ void chain(double* out, long long* cyc, double a, double b) {
  double x = a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = fma(x, b, a);
    if (OP == 1) x = x + b;
    if (OP == 2) x = 1.0 / (x + b);
    if (OP == 3) x = sqrt(x + b);
    if (OP == 4) x = rsqrt(x + b);
    if (OP == 5) { double s, c; sincos(x, &s, &c); x = s + c; }
    if (OP == 6) x = log(x + b);
    if (OP == 7) x = atan(x + b);
    if (OP == 8) { float f = (float)x; f = fmaf(f, (float)b, (float)a); x = f; }
    if (OP == 9) x = (x < b) ? x + a : x - a;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x; cyc[0] = t1 - t0;
}
template <int OP> void run(const char* name) {
  double* o; long long* c; cudaMalloc(&o, 8 * 64); cudaMalloc(&c, 8);
  chain<OP><<<1, 1>>>(o, c, 0.5, 1.0000001); chain<OP><<<1, 1>>>(o, c, 0.5, 1.0000001);
  long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
  chain<OP><<<1, 32>>>(o, c, 0.5, 1.0000001);
  long long h32; cudaMemcpy(&h32, c, 8, cudaMemcpyDeviceToHost);
  printf("%-28s %7.1f cycles/op (1 thread)  %7.1f (32 threads)\n", name, (double)h / N, (double)h32 / N);
}
int main() {
  run<0>("DFMA dependent"); run<1>("DADD dependent"); run<2>("1/(x+b) dependent"); run<3>("sqrt(x+b)");
  run<4>("rsqrt(x+b)"); run<5>("sincos"); run<6>("log"); run<7>("atan"); run<8>("cvt+FFMA+cvt"); run<9>("cmp+select+DADD");
  return 0;
}
