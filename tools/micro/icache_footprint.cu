// Microbenchmark: instruction-cache capacity on sm_100a.  A loop whose body is K straight-line FFMAs
// (16 bytes each): cycles per instruction against the body's footprint, one warp per SM sub-partition and
// four, warps in step (same start) or staggered by a quarter of the body each.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/icache tools/micro/icache_footprint.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__global__ void body_kernel(float *out, long long *cyc, int iters, int stagger) {
  float a = threadIdx.x * 1e-3f, b = 1.0001f, c = 0.5f, d = 0.25f;
  const int warp = threadIdx.x >> 5;
  // staggered start: warp w first runs w/4 of a body's worth of a DIFFERENT loop so the warps drift apart
  if (stagger) {
    for (int i = 0; i < warp * 997; ++i) a = fmaf(a, b, c);
  }
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < K / 4; ++k) {
      a = fmaf(a, b, c);
      b = fmaf(b, 1.0000001f, d);
      c = fmaf(c, 0.9999999f, a);
      d = fmaf(d, 1.0000002f, b);
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int K>
void run(float *out, long long *cyc, int threads, int blocks, int stagger) {
  const int iters = (1 << 22) / K;
  body_kernel<K><<<blocks, threads>>>(out, cyc, iters, stagger);
  cudaDeviceSynchronize();
  body_kernel<K><<<blocks, threads>>>(out, cyc, iters, stagger);
  cudaDeviceSynchronize();
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("  body %4d KB: %.2f cycles per instruction per warp\n", K * 16 / 1024, (double)h / ((double)iters * K));
}

int main() {
  float *out; long long *cyc;
  cudaMalloc(&out, 1 << 24); cudaMalloc(&cyc, 1 << 16);
  struct { int threads, blocks, stagger; const char *what; } cfg[] = {
      {128, 1, 0, "1 CTA on the device, 4 warps (1 per sub-partition), in step"},
      {128, 148, 0, "148 CTAs (every SM), 4 warps each, in step"},
      {512, 148, 1, "148 CTAs (every SM), 16 warps each, staggered"},
  };
  for (auto &c : cfg) {
    printf("%s\n", c.what);
    run<256>(out, cyc, c.threads, c.blocks, c.stagger);
    run<1024>(out, cyc, c.threads, c.blocks, c.stagger);
    run<2048>(out, cyc, c.threads, c.blocks, c.stagger);
    run<2560>(out, cyc, c.threads, c.blocks, c.stagger);
    run<3072>(out, cyc, c.threads, c.blocks, c.stagger);
    run<3584>(out, cyc, c.threads, c.blocks, c.stagger);
    run<4096>(out, cyc, c.threads, c.blocks, c.stagger);
    run<6144>(out, cyc, c.threads, c.blocks, c.stagger);
    run<8192>(out, cyc, c.threads, c.blocks, c.stagger);
    run<16384>(out, cyc, c.threads, c.blocks, c.stagger);
  }
  return 0;
}
