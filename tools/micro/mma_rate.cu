// micro-benchmark: warp-level mma.sync m16n8k8 tf32 issue rate per SM on sm_100a (is the legacy tensor path usable for tiny GEMMs?)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k(float* out, int iters) {
  float c[8][4];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  unsigned a0 = threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 * 11, b1 = a0 * 13;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0; for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) kf(float* out, int iters) {  // FFMA reference
  float c[32]; for (int i = 0; i < 32; ++i) c[i] = threadIdx.x + i;
  float a = threadIdx.x * 0.001f, b = 0.999f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = fmaf(c[i], b, a);
  }
  float s = 0; for (int i = 0; i < 32; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps = 4; warps <= 16; warps *= 2) {
    const int iters = 20000;
    k<<<148 * (warps / 4) / 2 * 2 / 2 * 1, 256>>>(out, 10);
    int blocks = 148 * warps / 8;
    k<<<blocks, 256>>>(out, iters); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<<<blocks, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double mma_per_sm = (double)iters * 8 * warps;  // per SM
    printf("warps/SM %2d: %.3f ms -> %.2f mma(m16n8k8 tf32)/us/SM = %.1f TFLOP/s chip (dense tf32)\n", warps, ms, mma_per_sm / (ms * 1e3),
           mma_per_sm * 148 * 2048.0 / (ms * 1e-3) / 1e12);
    cudaEventRecord(e0); kf<<<blocks, 256>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    printf("             FFMA: %.3f ms -> %.1f TFLOP/s chip\n", ms, (double)iters * 32 * 32 * warps * 148 * 2.0 / (ms * 1e-3) / 1e12);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
