// dev: dependent-chain latency and per-SM throughput of fp64 arithmetic on this GPU (the simulator kernels are bound by
// these, not by HBM -- DESIGN.md section 3).  nvcc -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void chain(double *out, long long *cyc, int iters, double a, double b) {
  double x = a + threadIdx.x * 1e-9, y = b;
  float xf = (float)x, yf = (float)b;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (OP == 0) x = fma(x, y, b);                                    // DFMA
      if (OP == 1) x = x + y;                                           // DADD
      if (OP == 2) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x)); x = r + b; }   // MUFU.RCP64H + DADD
      if (OP == 3) x = 1.0 / x + b;                                     // IEEE division + DADD
      if (OP == 4) xf = fmaf(xf, yf, yf);                               // FFMA (fp32 reference point)
      if (OP == 5) x = sqrt(x) + b;                                     // DSQRT + DADD
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = x + xf;
}

template <int OP>
void run(const char *name, int warps, int blocks) {
  double *out; long long *cyc, h;
  cudaMalloc(&out, sizeof(double) * blocks * warps * 32); cudaMalloc(&cyc, 8);
  const int iters = 2000;
  chain<OP><<<blocks, warps * 32>>>(out, cyc, iters, 0.999999, 1.0000001);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  chain<OP><<<blocks, warps * 32>>>(out, cyc, iters, 0.999999, 1.0000001);
  cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  const double ops = (double)iters * 16;
  printf("%-28s warps/CTA %2d CTAs %4d : %7.1f cycles per dependent op (warp 0), %8.2f Gop/s per thread-op aggregate\n", name, warps, blocks,
         h / ops, ops * warps * 32.0 * blocks / (ms * 1e-3) / 1e9);
  cudaFree(out); cudaFree(cyc);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s, %d SMs, %d MHz\n", p.name, p.multiProcessorCount, p.clockRate / 1000);
  run<0>("DFMA chain", 1, 1); run<1>("DADD chain", 1, 1); run<2>("RCP64H+DADD chain", 1, 1); run<3>("IEEE div+DADD chain", 1, 1);
  run<5>("DSQRT+DADD chain", 1, 1); run<4>("FFMA chain", 1, 1);
  for (int w : {4, 8, 16, 32}) run<0>("DFMA, one CTA per SM", w, p.multiProcessorCount);
  run<0>("DFMA, 2 CTAs per SM", 16, 2 * p.multiProcessorCount);
  run<4>("FFMA, 32 warps per SM", 32, p.multiProcessorCount);
  return 0;
}
