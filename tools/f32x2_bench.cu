// Microbenchmark: issue / pipe throughput of scalar FP32 vs packed f32x2 (FFMA2 / FADD2 / FMUL2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/f32x2_bench tools/f32x2_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r;}
__device__ __forceinline__ void upk(u64 v, float&a, float&b){ asm("mov.b64 {%0,%1}, %2;":"=f"(a),"=f"(b):"l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(r):"l"(a),"l"(b),"l"(c)); return r;}
__device__ __forceinline__ u64 add2(u64 a, u64 b){ u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(r):"l"(a),"l"(b)); return r;}
__device__ __forceinline__ float fma1(float a, float b, float c){ float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;":"=f"(r):"f"(a),"f"(b),"f"(c)); return r;}
__device__ __forceinline__ float add1(float a, float b){ float r; asm volatile("add.rn.f32 %0, %1, %2;":"=f"(r):"f"(a),"f"(b)); return r;}

constexpr int NCH = 8;      // independent chains per thread
constexpr int ITERS = 4096;

template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, float x, float y) {
  float s[NCH]; u64 p[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { s[i] = x + i + threadIdx.x; p[i] = pk(x + i, y + threadIdx.x); }
  const u64 cy = pk(y, y * 0.5f), cx = pk(x, -x);
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (MODE == 0) s[i] = fma1(s[i], y, x);                 // FFMA, 3 regs
      if (MODE == 1) p[i] = fma2(p[i], cy, cx);               // FFMA2
      if (MODE == 2) s[i] = add1(s[i], y);                    // FADD
      if (MODE == 3) p[i] = add2(p[i], cy);                   // FADD2
      if (MODE == 4) { s[i] = fma1(s[i], y, x); p[i] = add2(p[i], cy); }   // FFMA + FADD2 mix
      if (MODE == 5) { s[i] = fma1(s[i], y, x); s[i] = add1(s[i], y); }    // FFMA + FADD
      if (MODE == 6) { p[i] = fma2(p[i], cy, cx); p[i] = add2(p[i], cy); } // FFMA2 + FADD2
    }
  }
  float acc = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) { float a, b; upk(p[i], a, b); acc += s[i] + a + b; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> void run(const char* name, int instr_per_iter, int flops_per_instr_lane, int ctas_per_sm) {
  float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * ctas_per_sm;
  kern<MODE><<<grid, 256>>>(out, 1.0001f, 0.9999f);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  kern<MODE><<<grid, 256>>>(out, 1.0001f, 0.9999f);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double winstr = (double)grid * 8 * ITERS * NCH * instr_per_iter;   // warp instructions
  const double per_sm_per_ns = winstr / 148 / (ms * 1e6);
  printf("%-16s ctas/SM %d: %.3f ms  %.2f warp-instr/ns/SM (= %.2f per clk @1.965GHz)  %.1f TFLOP/s\n", name, ctas_per_sm, ms,
         per_sm_per_ns, per_sm_per_ns / 1.965, winstr * 32 * flops_per_instr_lane / (ms * 1e-3) / 1e12 / instr_per_iter * 1.0);
  cudaFree(out);
}

int main() {
  for (int c : {2, 4, 8}) {
    run<0>("FFMA", 1, 2, c);
    run<1>("FFMA2", 1, 4, c);
    run<2>("FADD", 1, 1, c);
    run<3>("FADD2", 1, 2, c);
    run<4>("FFMA+FADD2", 2, 3, c);   // flops column: (2 + 2) / 2 instr averaged below
    run<5>("FFMA+FADD", 2, 3, c);
    run<6>("FFMA2+FADD2", 2, 6, c);
  }
  return 0;
}
