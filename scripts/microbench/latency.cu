// Dependent-issue latencies on sm_100a that bound the checkpoint walk (one chain per SM): fp64 add / fma, warp shuffle of a
// double, shared-memory store->load round trip, CTA barrier with 2 / 5 / 9 warps.  nvcc -arch=sm_100a -O3 -o latency latency.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, int n, double seed) {
  __shared__ double sm[1024];
  const int lane = threadIdx.x & 31;
  double a = seed + lane, b = seed * 0.5;
  long long t0, t1;
  // DADD chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) a = a + b;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) a = fma(a, 1.0000001, b);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) a = a * 1.0000001;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // shuffle + add chain (one reduction level)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) a += __shfl_xor_sync(0xffffffffu, a, 1);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // shared store -> syncwarp -> load of the neighbour's value
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) {
    sm[threadIdx.x] = a;
    __syncwarp();
    a = sm[threadIdx.x ^ 1] + 1e-9;
    __syncwarp();
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // CTA barrier
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) __syncthreads();
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // store -> barrier -> load (the row-sum exchange of the walk)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) {
    sm[threadIdx.x] = a;
    __syncthreads();
    a = sm[(threadIdx.x + 32) % blockDim.x] + 1e-9;
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  // 8 independent DFMA chains per thread (throughput of one warp)
  double c0 = a, c1 = a + 1, c2 = a + 2, c3 = a + 3, c4 = a + 4, c5 = a + 5, c6 = a + 6, c7 = a + 7;
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < n; ++i) {
    c0 = fma(c0, 1.0000001, b); c1 = fma(c1, 1.0000001, b); c2 = fma(c2, 1.0000001, b); c3 = fma(c3, 1.0000001, b);
    c4 = fma(c4, 1.0000001, b); c5 = fma(c5, 1.0000001, b); c6 = fma(c6, 1.0000001, b); c7 = fma(c7, 1.0000001, b);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[7] = t1 - t0;
  out[threadIdx.x] = a + c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 8192); cudaMalloc(&cyc, 64);
  const int n = 4096;
  const char* names[] = {"DADD dependent", "DFMA dependent", "DMUL dependent", "SHFL.xor(double)+DADD", "STS->syncwarp->LDS+DADD (x2 syncwarp)", "bar.sync", "STS->bar.sync->LDS+DADD", "8 independent DFMA (per 8 ops)"};
  for (int nt : {32, 64, 160, 288}) {
    k<<<1, nt>>>(out, cyc, n, 1.0);
    k<<<1, nt>>>(out, cyc, n, 1.0);
    long long h[8];
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    for (int i = 0; i < 8; ++i) printf("threads %3d  %-42s %7.1f cycles\n", nt, names[i], (double)h[i] / n);
  }
  return 0;
}
