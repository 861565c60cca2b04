// Microbenchmark: tcgen05.mma kind::f16 issue/execute rate vs N and accumulator dependency.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../odin_b200/csrc/tc_ptx.cuh"
using namespace odin::ptx;

// mode: number of independent accumulators cycled (1 = fully dependent chain); ts: A from TMEM (1) or smem (0)
template <int CEV, int TS>
__global__ void __launch_bounds__(128, 1) k(int N, int nacc, int nmma, long long* out) {
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (raw & 1023u)) & 1023u;
  const uint32_t sbase = raw + pad;
  unsigned char* smem = smem_dyn + pad;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar = sbase + 65536 + 32768, tptr = bar + 32;
  for (int i = threadIdx.x; i < (65536 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1 << 20); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc512(tptr);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + 65536 + 32768 + 32);
  if (warp == 0) {
    const uint32_t idesc = idesc_f16(N);
    const uint64_t bdesc = desc_k_sw128(sbase), adesc = desc_k_sw128(sbase + 65536);
    long long t0 = 0, t1 = 0;
    for (int rep = 0; rep < 3; ++rep) {
      __syncwarp();
      t0 = clock64();
      if (elect_one()) {
        const uint32_t d0 = tmem + 128, d1 = tmem + 128 + (nacc > 1 ? N : 0), d2 = tmem + 128 + (nacc > 2 ? 2 * N : 0);
        const uint32_t dd[6] = {d0, d1, d2, d0, d1, d2};
        for (int i = 0; i < nmma; i += 12) {
#pragma unroll
          for (int j = 0; j < 12; ++j) {
            const uint32_t d = nacc == 1 ? d0 : (nacc == 2 ? ((j & 1) ? d1 : d0) : dd[j % 3]);
            const uint64_t o = (uint64_t)(((j & 3) * 32) >> 4);
            if (TS) mma_f16_ts(d, tmem + 8 * (j & 7), bdesc + o, idesc, 1u);
            else mma_f16_ss(d, adesc + o, bdesc + o, idesc, 1u);
            if (CEV > 0 && (j % CEV) == CEV - 1) tc_commit(bar + 8);
          }
        }
        tc_commit(bar);
      }
      __syncwarp();
      mbar_wait(bar, (uint32_t)(rep & 1));
      t1 = clock64();
    }
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc512(tmem); }
}

template <int CEV, int TS>
void run(long long* d, int smem) {
  cudaFuncSetAttribute(k<CEV, TS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int N : {64, 128, 256}) {
    const int nmma = 240, grid = 148;
    k<CEV, TS><<<grid, 128, smem>>>(N, 1, nmma, d);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%s N=%3d commit every %2d : %6.1f cycles/MMA (ideal %d) %s\n", TS ? "TS" : "SS", N, CEV,
           (double)mx / nmma, N / 2, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}
int main() {
  long long* d; cudaMalloc(&d, 148 * 8);
  const int smem = 65536 + 32768 + 64 + 1024;
  run<0, 1>(d, smem); run<12, 1>(d, smem); run<6, 1>(d, smem); run<3, 1>(d, smem); run<1, 1>(d, smem);
  run<0, 0>(d, smem); run<12, 0>(d, smem); run<3, 0>(d, smem);
  return 0;
}
