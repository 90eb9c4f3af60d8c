// Microbenchmark: fp64 FMA (CUDA cores) vs fp64 mma.sync m8n8k4 (tensor cores) peak on this GPU.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peak tools/fp64_peak.cu && /tmp/fp64_peak
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dfma_kernel(double* out, int iters) {
  double a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 1e-3 + i;
  double x = 1.0000001, y = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
  double s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dmma_kernel(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = threadIdx.x * 1e-3 + i;
  double a = 1.0000001, b = 1e-3;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0;
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out;
  cudaMalloc(&out, sizeof(double) * sms * 8 * 1024);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {128, 256, 512, 1024}) {
    int blocks = sms * (2048 / threads);
    int iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dfma_kernel<<<blocks, threads>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 2.0 * 16 * iters * (double)blocks * threads;
    printf("DFMA  threads/block %4d: %.2f TFLOP/s\n", threads, fl / ms / 1e9);
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      dmma_kernel<<<blocks, threads>>>(out, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    cudaEventElapsedTime(&ms, e0, e1);
    fl = 2.0 * 256 * 8 * iters * (double)blocks * (threads / 32);
    printf("DMMA  threads/block %4d: %.2f TFLOP/s\n", threads, fl / ms / 1e9);
  }
  return 0;
}
