// Microbenchmark (diagnostic): MUFU throughput per SM for the activation candidates of the PACL epilogue.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_rate mufu_rate.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

template <int OP>
__global__ void k(float* out, long long* cyc, int iters) {
  float x[8];
  unsigned h[8];
  for (int i = 0; i < 8; ++i) { x[i] = 0.001f * (threadIdx.x + i); h[i] = 0x3c003800u + threadIdx.x + i; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      if (OP == 3) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 4) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 5) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 6) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(x[i]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += x[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int warps) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<OP><<<148, warps * 32>>>(out, cyc, iters);
  k<OP><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  double ops = (double)iters * 8 * warps * 32;
  printf("%-22s warps/SM %2d: %.2f instr-lanes/clk/SM  (%.1f cycles per warp-instruction per SMSP)\n", name, warps, ops / c,
         c / (iters * 8.0 * warps / 4.0));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int w : {4, 8}) {
    if (w == 4) { run<0>("tanh.approx.f32", 4); run<1>("ex2.approx.f32", 4); run<2>("rcp.approx.f32", 4); run<3>("tanh.approx.f16x2", 4);
                  run<4>("tanh.approx.bf16x2", 4); run<5>("ex2.approx.f16x2", 4); run<6>("fma.f32", 4); }
    else { run<0>("tanh.approx.f32", 8); run<1>("ex2.approx.f32", 8); run<2>("rcp.approx.f32", 8); run<3>("tanh.approx.f16x2", 8);
           run<4>("tanh.approx.bf16x2", 8); run<5>("ex2.approx.f16x2", 8); run<6>("fma.f32", 8); }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
