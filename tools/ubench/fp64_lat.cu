// Dependent-issue latency of FP64 operations on one warp (cycles per operation), B200.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(double* out, long long* cyc, double a, double b, int n) {
  double x = a + threadIdx.x * 1e-9;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (OP == 0) x = __dadd_rn(x, b);
    if (OP == 1) x = __dmul_rn(x, b);
    if (OP == 2) x = __fma_rn(x, b, a);
    if (OP == 3) x = __ddiv_rn(x, b);
    if (OP == 4) x = __dsqrt_rn(x) + a;
    if (OP == 5) x = log(x) + a;
    if (OP == 6) x = exp(x * 1e-3);
  }
  long long t1 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
  double* out; long long* cyc; cudaMalloc(&out, 8 << 20); cudaMallocManaged(&cyc, 8);
  const char* names[] = {"dadd", "dmul", "dfma", "ddiv_rn", "dsqrt_rn+add", "log+add", "exp(mul)"};
  const int n = 4096;
  for (int blocks : {1, 148 * 4}) for (int threads : {32, 256}) {
    printf("blocks %d threads %d:", blocks, threads);
#define RUN(OP) k<OP><<<blocks, threads>>>(out, cyc, 1.000001, 1.0000001, n); cudaDeviceSynchronize(); \
    k<OP><<<blocks, threads>>>(out, cyc, 1.000001, 1.0000001, n); cudaDeviceSynchronize(); printf(" %s %.1f", names[OP], (double)*cyc / n);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
    printf("\n");
  }
  return 0;
}
