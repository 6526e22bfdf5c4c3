// Micro-benchmark: dependent-chain latency and throughput of fp64 add / mul / div on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain_add(double* out, double x, int n) {
  double acc = out[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { acc += x; acc += x; acc += x; acc += x; acc += x; acc += x; acc += x; acc += x; }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("DADD dependent chain: %.2f cycles/op (warps/block %d)\n", double(t1 - t0) / (8.0 * n), blockDim.x / 32);
}
__global__ void chain_lds_add(double* out, int n) {
  __shared__ double sm[8 * 32];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = 1e-9 * i;
  __syncthreads();
  double acc = out[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < n; i++) {
    const double v0 = sm[threadIdx.x % 32], v1 = sm[32 + threadIdx.x % 32], v2 = sm[64 + threadIdx.x % 32], v3 = sm[96 + threadIdx.x % 32];
    const double v4 = sm[128 + threadIdx.x % 32], v5 = sm[160 + threadIdx.x % 32], v6 = sm[192 + threadIdx.x % 32], v7 = sm[224 + threadIdx.x % 32];
    acc += v0; acc += v1; acc += v2; acc += v3; acc += v4; acc += v5; acc += v6; acc += v7;
  }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("LDS + DADD chain: %.2f cycles/op\n", double(t1 - t0) / (8.0 * n));
}
__global__ void thr_fma(double* out, double x, int n) {
  double a0 = out[threadIdx.x], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { a0 = a0 * x + x; a1 = a1 * x + x; a2 = a2 * x + x; a3 = a3 * x + x; a4 = a4 * x + x; a5 = a5 * x + x; a6 = a6 * x + x; a7 = a7 * x + x; }
  long long t1 = clock64();
  out[threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("DFMA throughput, %d warps/SM: %.2f cycles per warp-instruction per SM-subpartition-equivalent (total %.1f warp-instr/cycle/SM)\n", blockDim.x / 32, double(t1 - t0) / (8.0 * n), (blockDim.x / 32) * 8.0 * n / double(t1 - t0));
}
__global__ void chain_div(double* out, double x, int n) {
  double acc = out[threadIdx.x];
  long long t0 = clock64();
  for (int i = 0; i < n; i++) { acc = x / acc; acc = x / acc; acc = x / acc; acc = x / acc; }
  long long t1 = clock64();
  out[threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) printf("DDIV (IEEE) dependent chain: %.2f cycles/op\n", double(t1 - t0) / (4.0 * n));
}
int main() {
  double* d; cudaMalloc(&d, 1024 * 8); cudaMemset(d, 0, 1024 * 8);
  chain_add<<<1, 32>>>(d, 1e-9, 10000); cudaDeviceSynchronize();
  chain_add<<<1, 512>>>(d, 1e-9, 10000); cudaDeviceSynchronize();
  chain_lds_add<<<1, 32>>>(d, 10000); cudaDeviceSynchronize();
  thr_fma<<<1, 128>>>(d, 1.0000001, 10000); cudaDeviceSynchronize();
  thr_fma<<<1, 512>>>(d, 1.0000001, 10000); cudaDeviceSynchronize();
  thr_fma<<<1, 1024>>>(d, 1.0000001, 10000); cudaDeviceSynchronize();
  chain_div<<<1, 32>>>(d, 1.5, 10000); cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
