// Micro-benchmark: do warp shuffles share the shared-memory data pipe?  (sm_100a)
// Throughput of 32-bit SHFL, of LDS.64 and of both interleaved, 32 warps per SM on every SM.
#include <cstdio>
#include <cuda_runtime.h>
#define IT 2048
__global__ void k_shfl(int* out, long long* cyc) {
  int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < IT; i++) {
    x0 = __shfl_sync(0xffffffffu, x0, (threadIdx.x + 1) & 31);
    x1 = __shfl_sync(0xffffffffu, x1, (threadIdx.x + 2) & 31);
    x2 = __shfl_sync(0xffffffffu, x2, (threadIdx.x + 3) & 31);
    x3 = __shfl_sync(0xffffffffu, x3, (threadIdx.x + 5) & 31);
  }
  __syncthreads();
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc) {
  __shared__ double s[1024 + 64];
  for (int i = threadIdx.x; i < 1024 + 64; i += blockDim.x) s[i] = i;
  __syncthreads();
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < IT; i++) {
    a0 += s[p]; a1 += s[p + 16]; a2 += s[p + 32]; a3 += s[p + 48];
    p = (p + 1) & 1023;
  }
  __syncthreads();
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_both(double* out, long long* cyc) {
  __shared__ double s[1024 + 64];
  for (int i = threadIdx.x; i < 1024 + 64; i += blockDim.x) s[i] = i;
  __syncthreads();
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3;
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < IT; i++) {
    a0 += s[p]; a1 += s[p + 16]; a2 += s[p + 32]; a3 += s[p + 48];
    x0 = __shfl_sync(0xffffffffu, x0, (threadIdx.x + 1) & 31);
    x1 = __shfl_sync(0xffffffffu, x1, (threadIdx.x + 2) & 31);
    x2 = __shfl_sync(0xffffffffu, x2, (threadIdx.x + 3) & 31);
    x3 = __shfl_sync(0xffffffffu, x3, (threadIdx.x + 5) & 31);
    p = (p + 1) & 1023;
  }
  __syncthreads();
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + x0 + x1 + x2 + x3;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  void* out; long long* cyc; long long h;
  cudaMalloc(&out, 8 * 148 * 1024); cudaMalloc(&cyc, 8);
  for (int rep = 0; rep < 2; rep++) {
    k_shfl<<<148, 1024>>>((int*)out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("4 SHFL.32 x 32 warps: %.2f cycles per iteration per SM (%.3f per warp-instr)\n", (double)h / IT, (double)h / IT / 128);
    k_lds<<<148, 1024>>>((double*)out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("4 LDS.64  x 32 warps: %.2f cycles per iteration per SM (%.3f per warp-instr)\n", (double)h / IT, (double)h / IT / 128);
    k_both<<<148, 1024>>>((double*)out, cyc); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("4 LDS.64 + 4 SHFL.32 x 32 warps: %.2f cycles per iteration per SM\n", (double)h / IT);
  }
  return 0;
}
