// Microbenchmark: fp64 throughput of the DFMA pipe against the fp64 tensor instruction (mma.sync m8n8k4 f64, DMMA) on sm_100a,
// with the operand traffic a register-blocked GEMM would add left out (pure issue / pipe rate).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/f64_bench tools/f64_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int NCH = 16;      // independent chains per thread
constexpr int ITERS = 2048;

template <int MODE>
__global__ void __launch_bounds__(256) kern(double* out, double x, double y) {
  double s[NCH], t[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { s[i] = x + i + threadIdx.x; t[i] = y - i; }
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < NCH; ++i) {
      if (MODE == 0) s[i] = fma(s[i], y, x);
      if (MODE == 1) {   // D (2 regs) = A (1) * B (1) + C (2): 8x8x4 per warp = 256 FMA
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(s[i]), "+d"(t[i]) : "d"(x), "d"(y));
      }
    }
  }
  double acc = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) acc += s[i] + t[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE> void run(const char* name, double fma_per_warp_instr, int ctas_per_sm) {
  double* out; cudaMalloc(&out, 148 * 8 * 256 * 8);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int grid = 148 * ctas_per_sm;
  kern<MODE><<<grid, 256>>>(out, 1.0001, 0.9999);
  cudaDeviceSynchronize();
  cudaEventRecord(a);
  kern<MODE><<<grid, 256>>>(out, 1.0001, 0.9999);
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double winstr = (double)grid * 8 * ITERS * NCH;
  printf("%-6s ctas/SM %d: %.3f ms  %.3f warp-instr/clk/SM @1.965GHz  %.1f TFLOP/s\n", name, ctas_per_sm, ms,
         winstr / 148 / (ms * 1e6) / 1.965, winstr * fma_per_warp_instr * 2 / (ms * 1e-3) / 1e12);
  cudaFree(out);
}

int main() {
  for (int c : {1, 2, 4}) {
    run<0>("DFMA", 32, c);
    run<1>("DMMA", 256, c);
  }
  return 0;
}
