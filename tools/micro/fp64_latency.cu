// Microbenchmark: dependent-issue latency and throughput of fp64 / shared-memory instructions on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gpurun_out/fp64_latency tools/micro/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_chain(double *out, long long *cyc, double a, double b, int iters) {
  double x[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
#pragma unroll
      for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
  }
  const long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += x[c];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void lds_chain(int *out, long long *cyc, int iters) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i * 7 + 1) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) p = s[p];
  }
  const long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void sts_sync_lds_chain(double *out, long long *cyc, int iters) {
  __shared__ double s[64];
  double v = threadIdx.x;
  const int lane = threadIdx.x & 31;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      s[lane] = v;
      __syncwarp();
      v = s[(lane + 1) & 31] + 1.0;
      __syncwarp();
    }
  }
  const long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void shfl_chain(double *out, long long *cyc, int iters) {
  double v = threadIdx.x;
  const long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) v = __shfl_xor_sync(0xffffffffu, v, 1) + 1.0;
  }
  const long long t1 = clock64();
  out[threadIdx.x] = v;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  double *out; long long *cyc; long long h[8];
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024);
  const int iters = 1000;
#define RUN(name, kern, per_iter, ...) \
  kern<<<1, 32>>>(__VA_ARGS__); cudaDeviceSynchronize(); kern<<<1, 32>>>(__VA_ARGS__); cudaDeviceSynchronize(); \
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-40s %.2f cycles per step\n", name, (double)h[0] / (iters * per_iter));
  RUN("DFMA dependent chain (1 warp)", dfma_chain<1>, 16, out, cyc, 1.0000001, 1e-9, iters)
  RUN("DFMA 2 independent chains: per DFMA", dfma_chain<2>, 32, out, cyc, 1.0000001, 1e-9, iters)
  RUN("DFMA 4 independent chains: per DFMA", dfma_chain<4>, 64, out, cyc, 1.0000001, 1e-9, iters)
  RUN("DFMA 8 independent chains: per DFMA", dfma_chain<8>, 128, out, cyc, 1.0000001, 1e-9, iters)
  RUN("LDS dependent chain", lds_chain, 16, (int *)out, cyc, iters)
  RUN("STS + syncwarp + LDS + DADD + syncwarp", sts_sync_lds_chain, 8, out, cyc, iters)
  RUN("SHFL.64 + DADD", shfl_chain, 8, out, cyc, iters)
  // 4 warps on one SM sub-partition each? (128 threads = 1 warp per sub-partition)
  dfma_chain<8><<<1, 128>>>(out, cyc, 1.0000001, 1e-9, iters); cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-40s %.2f cycles per DFMA per warp\n", "8 chains, 4 warps (1 per SMSP)", (double)h[0] / (iters * 128));
  dfma_chain<8><<<1, 256>>>(out, cyc, 1.0000001, 1e-9, iters); cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-40s %.2f cycles per DFMA per warp\n", "8 chains, 8 warps (2 per SMSP)", (double)h[0] / (iters * 128));
  dfma_chain<1><<<1, 1024>>>(out, cyc, 1.0000001, 1e-9, iters); cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-40s %.2f cycles per DFMA per warp\n", "1 chain, 32 warps (8 per SMSP)", (double)h[0] / (iters * 16));
  return 0;
}
