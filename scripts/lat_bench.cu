// Micro-benchmark: FP64 instruction latencies / issue intervals on one SM sub-partition (sm_100a).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/lat_bench scripts/lat_bench.cu
// The numbers size the latency model of the GP-fit and RK kernels in DESIGN.md.
#include <cstdio>
#include <cuda_runtime.h>

#define N_IT 4096

__global__ void k_dfma_dep(double* out, long long* cyc, double a, double b) {
  double x = threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) x = fma(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_dadd_dep(double* out, long long* cyc, double a, double b) {
  double x = threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) x = __dadd_rn(x, b);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ILP>
__global__ void k_dfma_ilp(double* out, long long* cyc, double a, double b) {
  double x[ILP];
#pragma unroll
  for (int j = 0; j < ILP; j++) x[j] = threadIdx.x * 1e-3 + j;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N_IT; i++) {
#pragma unroll
    for (int j = 0; j < ILP; j++) x[j] = fma(x[j], a, b);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int j = 0; j < ILP; j++) s += x[j];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_shfl_dep(double* out, long long* cyc) {
  double x = threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) x = __shfl_sync(0xffffffffu, x, (i * 7 + 3) & 31);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcp_dep(double* out, long long* cyc) {
  double x = 1.0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) {
    double y;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    x = y;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_rcpfull_dep(double* out, long long* cyc) {
  double x = 1.0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) {
    double y;
    asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    y = fma(y, e, y);
    e = fma(-x, y, 1.0);
    x = fma(y, e, y);
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_div_dep(double* out, long long* cyc, double a) {
  double x = 1.0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) x = a / x;
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_smem_dep(double* out, long long* cyc) {
  __shared__ double s[64];
  s[threadIdx.x] = threadIdx.x;
  s[threadIdx.x + 32] = threadIdx.x;
  __syncwarp();
  double x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N_IT; i++) {
    s[threadIdx.x] = x;
    __syncwarp();
    x = s[(threadIdx.x + 1) & 31];
    __syncwarp();
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_bar_dep(double* out, long long* cyc) {
  __shared__ double s[2][256];
  double x = threadIdx.x;
  long long t0 = clock64();
  for (int i = 0; i < N_IT; i++) {
    s[i & 1][threadIdx.x] = x;
    __syncthreads();
    x = s[i & 1][(threadIdx.x + 1) & 255] + 1.0;
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_exp_dep(double* out, long long* cyc) {
  double x = -1.0 - threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < N_IT; i++) x = -exp(x);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void k_log_dep(double* out, long long* cyc) {
  double x = 3.0 + threadIdx.x * 1e-3;
  long long t0 = clock64();
  for (int i = 0; i < N_IT; i++) x = 3.0 + log(x);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  double* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(double) * 148 * 64 * 1024);
  cudaMalloc(&cyc, sizeof(long long));
  long long h;
  auto rd = [&](const char* name, double per) {
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s %8.2f cycles\n", name, (double)h / N_IT / per);
  };
  for (int rep = 0; rep < 2; rep++) {
    k_dfma_dep<<<1, 32>>>(out, cyc, 0.999, 1e-3); rd("DFMA dependent latency", 1);
    k_dadd_dep<<<1, 32>>>(out, cyc, 0.999, 1e-3); rd("DADD dependent latency", 1);
    k_dfma_ilp<8><<<1, 32>>>(out, cyc, 0.999, 1e-3); rd("DFMA issue interval, 1 warp ILP8", 8);
    k_dfma_ilp<8><<<1, 128>>>(out, cyc, 0.999, 1e-3); rd("DFMA per-warp interval, 4 warps (1/SMSP)", 8);
    k_dfma_ilp<8><<<1, 256>>>(out, cyc, 0.999, 1e-3); rd("DFMA per-warp interval, 8 warps (2/SMSP)", 8);
    k_dfma_ilp<8><<<1, 512>>>(out, cyc, 0.999, 1e-3); rd("DFMA per-warp interval, 16 warps (4/SMSP)", 8);
    k_dfma_ilp<8><<<1, 1024>>>(out, cyc, 0.999, 1e-3); rd("DFMA per-warp interval, 32 warps (8/SMSP)", 8);
    k_shfl_dep<<<1, 32>>>(out, cyc); rd("64-bit SHFL dependent latency", 1);
    k_rcp_dep<<<1, 32>>>(out, cyc); rd("MUFU.RCP64H dependent latency", 1);
    k_rcpfull_dep<<<1, 32>>>(out, cyc); rd("rcp seed + 2 Newton (4 DFMA) latency", 1);
    k_div_dep<<<1, 32>>>(out, cyc, 1.7); rd("double division latency", 1);
    k_smem_dep<<<1, 32>>>(out, cyc); rd("st.shared+syncwarp+ld.shared+syncwarp", 1);
    k_bar_dep<<<1, 256>>>(out, cyc); rd("st.shared+__syncthreads(256)+ld+DADD", 1);
    k_exp_dep<<<1, 32>>>(out, cyc); rd("libdevice exp() dependent latency", 1);
    k_log_dep<<<1, 32>>>(out, cyc); rd("libdevice log() dependent latency", 1);
  }
  return 0;
}
