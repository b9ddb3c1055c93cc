// fp64_latency.cu — dependent-issue latency of FP64 ops on one warp (clock64 around unrolled dependent chains), and the same
// chain with 2/4/8 independent accumulators per thread (ILP) and 1..8 warps per SM sub-partition (TLP).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void chain(double* out, long long* cyc, double a, double b, int iters) {
  double x[ILP];
  for (int k = 0; k < ILP; k++) x[k] = threadIdx.x * 1e-3 + k;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++)
#pragma unroll
      for (int k = 0; k < ILP; k++) x[k] = fma(x[k], a, b);
  }
  const long long t1 = clock64();
  double s = 0; for (int k = 0; k < ILP; k++) s += x[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void chainRcp(double* out, long long* cyc, int iters) {
  double x = 1.0 + threadIdx.x * 1e-3;
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 16; u++) { double y; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x)); x = y + 1.0; }
  }
  const long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void chainLds(double* out, long long* cyc, int iters) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i + 33) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < iters * 16; i++) p = nxt[p];
  const long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
void run(double* out, long long* cyc, int threads) {
  const int iters = 256;
  chain<ILP><<<1, threads>>>(out, cyc, 0.999999, 1e-9, iters);
  chain<ILP><<<1, threads>>>(out, cyc, 0.999999, 1e-9, iters);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("DFMA ilp %d warps/SM %2d (per sub-partition %.1f): %.2f cycles per dependent step, %.2f cycles per warp-instruction per sub-partition\n", ILP, threads / 32,
         threads / 128.0, (double)h / (iters * 16), (double)h / (iters * 16.0 * ILP * (threads / 32) / 4.0 > 0 ? iters * 16.0 * ILP * (threads < 128 ? 1 : threads / 128) : 1));
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1 << 12);
  run<1>(out, cyc, 32); run<2>(out, cyc, 32); run<4>(out, cyc, 32); run<8>(out, cyc, 32); run<16>(out, cyc, 32);
  run<1>(out, cyc, 128); run<1>(out, cyc, 256); run<1>(out, cyc, 512); run<1>(out, cyc, 1024);
  run<4>(out, cyc, 512); run<8>(out, cyc, 512);
  chainRcp<<<1, 32>>>(out, cyc, 256); cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("MUFU.RCP64H + DADD dependent step: %.2f cycles\n", (double)h / (256 * 16));
  chainLds<<<1, 32>>>(out, cyc, 256); cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("LDS pointer chase: %.2f cycles\n", (double)h / (256 * 16));
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
