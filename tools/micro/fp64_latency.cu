// fp64 latency probes on one warp: dependent DFMA chain, dependent division chain, dependent sqrt chain,
// independent divisions (ILP 4 / 8), SHFL of a double. Prints cycles per operation.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void probe(double *out, long long *cyc, double a, double b, int n) {
  double x = a + threadIdx.x * 1e-9, y = b, z0 = a, z1 = a + 1, z2 = a + 2, z3 = a + 3, z4 = a + 4, z5 = a + 5, z6 = a + 6, z7 = a + 7;
  long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (MODE == 0) x = fma(x, y, 1.0);
    if (MODE == 1) x = (x + 3.0) / y;
    if (MODE == 2) x = sqrt(x + 2.0);
    if (MODE == 3) { z0 = (z0 + 3.0) / y; z1 = (z1 + 3.0) / y; z2 = (z2 + 3.0) / y; z3 = (z3 + 3.0) / y; }
    if (MODE == 4) { z0 = (z0 + 3.0) / y; z1 = (z1 + 3.0) / y; z2 = (z2 + 3.0) / y; z3 = (z3 + 3.0) / y;
                     z4 = (z4 + 3.0) / y; z5 = (z5 + 3.0) / y; z6 = (z6 + 3.0) / y; z7 = (z7 + 3.0) / y; }
    if (MODE == 5) x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31);
    if (MODE == 6) x = x + y;
    if (MODE == 7) x = x * y + 0.5;  // fmad=false: DMUL + DADD
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  out[threadIdx.x] = x + z0 + z1 + z2 + z3 + z4 + z5 + z6 + z7;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 256 * 8); cudaMalloc(&cyc, 8);
  const int n = 2000;
  const char *names[] = {"dependent DFMA", "dependent DDIV (+DADD)", "dependent DSQRT (+DADD)", "4 independent DDIV per iter", "8 independent DDIV per iter", "SHFL double", "dependent DADD", "DMUL+DADD"};
  for (int threads = 32; threads <= 256; threads *= 8) {
    printf("-- %d threads per block, 1 block --\n", threads);
    for (int m = 0; m < 8; ++m) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (m) {
          case 0: probe<0><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, n); break;
          case 1: probe<1><<<1, threads>>>(out, cyc, 5.0, 1.7, n); break;
          case 2: probe<2><<<1, threads>>>(out, cyc, 5.0, 1.7, n); break;
          case 3: probe<3><<<1, threads>>>(out, cyc, 5.0, 1.7, n); break;
          case 4: probe<4><<<1, threads>>>(out, cyc, 5.0, 1.7, n); break;
          case 5: probe<5><<<1, threads>>>(out, cyc, 5.0, 1.7, n); break;
          case 6: probe<6><<<1, threads>>>(out, cyc, 5.0, 1e-9, n); break;
          case 7: probe<7><<<1, threads>>>(out, cyc, 1.0000001, 0.9999999, n); break;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("%-32s %8.1f cycles per iteration\n", names[m], (double)h / n);
    }
  }
  return 0;
}
