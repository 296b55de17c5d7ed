// MUFU throughput probe: is a packed half-precision ex2 (two results per instruction) issued at the rate of the fp32 one?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/lab/mufu_lab tools/lab/mufu_lab.cu && tools/lab/mufu_lab
// Prints results per second per variant (registers only, 8 independent chains per thread, all SMs busy).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

__device__ __forceinline__ float ex2_f32(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned ex2_bf16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ unsigned ex2_f16x2(unsigned x) { unsigned y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }

template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters) {
  float f[8];
  unsigned u[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { f[k] = -1e-3f * (threadIdx.x + k); u[k] = 0xbc00bc00u + threadIdx.x + k; }
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (MODE == 0) f[k] = ex2_f32(f[k]) - 1.0f;                  // one FADD keeps the value in range (separate pipe)
      if (MODE == 1) u[k] = ex2_bf16x2(u[k]) ^ 0x80008000u;        // sign flip on the ALU pipe keeps it in range
      if (MODE == 2) u[k] = ex2_f16x2(u[k]) ^ 0x80008000u;
    }
  }
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += f[k] + __uint_as_float(u[k]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, int results_per_instr) {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, iters = 20000;
  float* out; cudaMalloc(&out, sizeof(float) * blocks * 256);
  probe<MODE><<<blocks, 256>>>(out, 100);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  probe<MODE><<<blocks, 256>>>(out, iters);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  const double instr = (double)blocks * 256 * iters * 8;
  printf("%-28s %8.3f ms  %7.2f G instr/s  %7.2f G results/s  (%.1f results / clk / SM at 1.9 GHz)\n", name, ms, instr / ms / 1e6,
         instr * results_per_instr / ms / 1e6, instr * results_per_instr / ms / 1e6 / sms / 1.9);
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.ftz.bf16x2", 2);
  run<2>("ex2.approx.f16x2", 2);
  return 0;
}
