// fp64_peak.cu — measures this B200's FP64 ceilings (vector DFMA, DMMA m8n8k4, both together) and a copy bandwidth.
// MEASURED_PEAKS.json has no FP64 figure; DESIGN.md / bench.py quote the numbers this prints.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

__global__ void dfmaKernel(double* out, int iters) {
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-3 + i;
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], b, c);
  }
  double s = 0; for (int i = 0; i < 8; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__global__ void dmmaKernel(double* out, int iters, int mix) {
  double c[8][2];
  for (int i = 0; i < 8; i++) { c[i][0] = 0; c[i][1] = 0; }
  double v[4] = {1, 2, 3, 4};
  const double a = 1e-3 * threadIdx.x, b = 1.0 + 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) dmma(c[i][0], c[i][1], a, b);
    if (mix) {
#pragma unroll
      for (int i = 0; i < 4; i++) v[i] = fma(v[i], b, a);
    }
  }
  double s = 0; for (int i = 0; i < 8; i++) s += c[i][0] + c[i][1];
  for (int i = 0; i < 4; i++) s += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void copyKernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = in[i];
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  double* out; cudaMalloc(&out, sizeof(double) * sms * 8 * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float ms;
  const int iters = 20000;
  for (int rep = 0; rep < 2; rep++) {
    dfmaKernel<<<sms * 8, 256>>>(out, iters);
  }
  cudaEventRecord(e0); dfmaKernel<<<sms * 8, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = 2.0 * 8 * iters * (double)sms * 8 * 256 / (ms * 1e-3) / 1e12;
  for (int rep = 0; rep < 2; rep++) dmmaKernel<<<sms * 8, 256>>>(out, iters, 0);
  cudaEventRecord(e0); dmmaKernel<<<sms * 8, 256>>>(out, iters, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  const double dm = 2.0 * 8 * 8 * 4 * 8 * iters * (double)sms * 8 * 8 / (ms * 1e-3) / 1e12;  // 256 FMA per warp-mma, 8 mma/iter, 8 warps/block
  cudaEventRecord(e0); dmmaKernel<<<sms * 8, 256>>>(out, iters, 1); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  const double mixT = (2.0 * 256 * 8 * iters * (double)sms * 8 * 8 + 2.0 * 4 * iters * (double)sms * 8 * 256) / (ms * 1e-3) / 1e12;
  const size_t n = (size_t)1 << 28;  // 4 GiB per buffer of double2
  double2 *a, *b; cudaMalloc(&a, n * sizeof(double2)); cudaMalloc(&b, n * sizeof(double2)); cudaMemset(a, 0, n * sizeof(double2));
  for (int rep = 0; rep < 2; rep++) copyKernel<<<sms * 16, 512>>>(a, b, n);
  cudaEventRecord(e0); copyKernel<<<sms * 16, 512>>>(a, b, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
  cudaEventElapsedTime(&ms, e0, e1);
  const double bw = 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.2f, \"dmma_tflops\": %.2f, \"dmma_plus_dfma_tflops\": %.2f, \"copy_gbs\": %.1f, \"err\": \"%s\"}\n",
         prop.name, sms, dfma, dm, mixT, bw, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
