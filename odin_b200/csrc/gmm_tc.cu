// GMM-UBM Baum-Welch on the 5th-generation tensor cores (tcgen05 + TMEM), in
// 3xTF32 split precision.  sm_100a only.
//
// Reference arithmetic: odin/ml/gmm_tmat.py:1012-1041 (_fast_expectation) with
// the cached constants of :493-504.  Everything is evaluated in the log2 domain:
//
//   lp2[m,b] = sum_k W[m,k] * A[b,k]                       (GEMM 1, K = 128)
//       A[b,:] = [ x^2 (D) | x (D) | 0.. | 1 (k=120) | -lse2[b] (k=121) | 0.. ]
//       W[m,:] = [ log2(e) * (-0.5 prec | mu prec) | 0.. | log2(e) * cst | 1 | 0.. ]
//   pass 1 (k=121 column zero):   lse2[b] = log2 sum_m 2^lp2[m,b]
//   pass 2:   P[m,b] = 2^(lp2[m,b] - lse2[b])  (the subtraction rides in GEMM 1)
//             stat[j,m] += sum_b A[b,j] * P[m,b]            (GEMM 2, K = frames)
//             rows j < D -> S, D <= j < 2D -> F, j = 120 -> Z (= N in north_star)
//
// One CTA owns a chunk of 128 mixtures for its whole life and streams tiles of
// 32 frames through a warp-specialised mbarrier pipeline:
//
//   producers (4 warps)  X tile -> smem raw copy -> (a) K-major SWIZZLE_128B hi/lo
//                        operand tile for GEMM 1, (b) the transposed [j, frame]
//                        hi/lo operand of GEMM 2 written straight into TMEM
//   MMA issuer (1 lane)  GEMM 1:  D1[128 mix, 32 frames]  = Whi*Ahi + Whi*Alo + Wlo*Ahi
//                                 (Whi lives in TMEM for the whole kernel, Wlo in smem)
//                        GEMM 2:  D2[128 j, 128 mix]     += Thi*Phi + Thi*Plo + Tlo*Phi
//                                 (T from TMEM, P from smem)
//   epilogue (4 warps)   TMEM -> registers; pass 1: warp-level max / sum over the
//                        mixture lanes (CREDUX / REDUX) -> per-chunk partial LSE;
//                        pass 2: P = ex2(D1), hi/lo split, swizzled smem operand
//                        tile; every `flush_tiles` tiles D2 is drained into the
//                        caller's fp64 statistics with red.global.add.f64
//
// TMEM map (512 columns): [0,128) Whi | [128,256) D2 | [256,384) 4 x D1 | [384,512) 2 x (Thi|Tlo)
#include <math.h>
#include <stdlib.h>

#include <algorithm>

#include "gmm.cuh"

namespace odin {

namespace tc {

constexpr int K = 128;          // padded contraction length of GEMM 1 / rows of GEMM 2
constexpr int CM = 128;         // mixtures per CTA
constexpr int NF = 32;          // frames per tile
constexpr int K_ONE = 120;      // column holding the constant 1 (-> cst, Z)
constexpr int K_LSE = 121;      // column holding -lse2[b] in pass 2
constexpr int BT_STAGES = 2;
constexpr int D1_BUFS = 4;
constexpr int TT_BUFS = 2;
constexpr int P_BUFS = 2;
constexpr int MAX_D = 60;       // 2*D + 2 <= 122 and D % 4 == 0

constexpr uint32_t TM_WHI = 0, TM_D2 = 128, TM_D1 = 256, TM_TT = 384;

// shared memory map (after 1024-byte alignment)
constexpr uint32_t SM_WLO = 0;                          // 64 KB  [4 kblocks][128 rows][128 B]
constexpr uint32_t SM_BT = 65536;                       // 2 x (hi 16 KB | lo 16 KB)
constexpr uint32_t SM_P = SM_BT + BT_STAGES * 32768;    // 2 x (hi 16 KB | lo 16 KB)
constexpr uint32_t SM_XRAW = SM_P + P_BUFS * 32768;     // 2 x 32 x 60 floats
constexpr uint32_t SM_LSE = SM_XRAW + 2 * NF * MAX_D * 4;   // 2 x 32 floats
constexpr uint32_t SM_BAR = SM_LSE + 2 * NF * 4;        // mbarriers
constexpr uint32_t SM_TOTAL = SM_BAR + 256;
constexpr uint32_t SMEM_BYTES = SM_TOTAL + 1024;        // alignment slack

// mbarrier slots
constexpr int B_BT_FULL = 0, B_BT_EMPTY = 2, B_D1_FULL = 4, B_D1_EMPTY = 8, B_TT_FULL = 12, B_TT_EMPTY = 14,
              B_P_FULL = 16, B_P_EMPTY = 18, B_D2_FULL = 20, B_D2_EMPTY = 21, B_COUNT = 22, B_TMEM_PTR = 24;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Bounded wait: a protocol bug traps after ~2 s (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  uint64_t t0 = 0;
  for (uint32_t spin = 0; !ok; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!ok && spin >= 8u) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 2000000000ull) __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// one lane of a converged warp (keeps the surrounding control flow warp-uniform so
// descriptors stay in uniform registers; a `lane == 0` branch makes ptxas wrap every
// tcgen05.mma in an ELECT / R2UR waterfall loop and the issue rate collapses)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem desc]     (kind::tf32, cta_group::1)
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}

// K-major SWIZZLE_128B shared-memory matrix descriptor (8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t saddr) {
  uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);   // start address          bits [0,14)
  d |= (uint64_t)1 << 16;                              // leading byte offset (unused with swizzle)
  d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset     bits [32,46)
  d |= (uint64_t)1 << 46;                              // descriptor version 1 (sm_100)
  d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
  return d;
}
// instruction descriptor: fp32 accumulate, tf32 x tf32, A and B K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 32 / 16 consecutive TMEM columns of this thread's lane <-> registers
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(v[i]);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31]), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(v[i]);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
      ::"r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc512(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(512) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(512) : "memory");
}

// 16-byte async copy global -> shared; src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// x = hi + lo with hi the TF32 rounding of x (round half away, what cvt.rna.tf32.f32
// does, but 2 integer ops instead of the 4-instruction sequence ptxas emits for it);
// lo is exact in fp32 and is truncated to TF32 by the tensor core (error 2^-22 |x|).
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
  lo = x - hi;
}
__device__ __forceinline__ void split4(const float4& v, float4& h, float4& l) {
  split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
}
// byte offset of 16-byte chunk k4 (= k / 4) of row r inside a K-major SWIZZLE_128B
// tile whose 32-element column blocks are `rows` rows tall
__device__ __forceinline__ uint32_t sw128_off(int r, int k4, int rows) {
  return (uint32_t)(k4 >> 3) * (uint32_t)(rows * 128) + (uint32_t)r * 128u + (uint32_t)(((k4 & 7) ^ (r & 7)) << 4);
}

}  // namespace tc

// ---------------------------------------------------------------------------
// operand images of the model (refreshed with the cached constants)
// ---------------------------------------------------------------------------
// Whi       : [Mpad][128] plain rows (pass 2 copies them into TMEM)
// Whs / Wls : hi / lo parts, per chunk of 128 mixtures the exact shared-memory image of a
//             K-major SWIZZLE_128B tile: [kblock 4][row 128][16-byte chunk (q ^ (row & 7))][4]
__global__ void gmm_tc_prepare_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                                      const float* __restrict__ w, int D, int M, int Mpad,
                                      float* __restrict__ Whi, float* __restrict__ Whs, float* __restrict__ Wls) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mpad) return;
  const double LOG2E = 1.4426950408889634074;
  const int chunk = m / tc::CM, row = m % tc::CM;
  const size_t cbase = (size_t)chunk * tc::CM * tc::K;
  auto put = [&](int k, double v) {
    float h, l;
    tc::split_tf32((float)v, h, l);
    Whi[(size_t)m * tc::K + k] = h;
    const size_t o = cbase + (size_t)(k >> 5) * (tc::CM * 32) + row * 32 + ((((k & 31) >> 2) ^ (row & 7)) << 2) + (k & 3);
    Whs[o] = h;
    Wls[o] = l;
  };
  if (m >= M) {  // padding mixture: log-density -1e30 -> posterior exactly 0
    for (int k = 0; k < tc::K; ++k) put(k, k == tc::K_ONE ? -1e30 : 0.0);
    return;
  }
  double C = 0.0;
  for (int d = 0; d < D; ++d) {
    const double v = (double)var[(size_t)d * M + m] + ODIN_GMM_EPS;
    const double p = 1.0 / v;
    const double mu = (double)mean[(size_t)d * M + m];
    C += mu * mu * p + log(v);
    put(d, -0.5 * p * LOG2E);
    put(D + d, mu * p * LOG2E);
  }
  C -= 2.0 * log((double)w[m] + ODIN_GMM_EPS);
  for (int k = 2 * D; k < tc::K; ++k) put(k, 0.0);
  put(tc::K_ONE, -0.5 * (C + (double)D * 1.8378770664093454835606594728112) * LOG2E);
  put(tc::K_LSE, 1.0);
}

struct TcArgs {
  const float* X;
  const uint8_t* sad;
  int64_t N;
  int D, M;
  const float* Whi;
  const float* Whs;
  const float* Wls;
  const float* lse2;     // pass 2: per-frame log2-domain log-sum-exp
  float2* part;          // pass 1: [nchunks][part_stride] (max, sum) per frame
  int64_t part_stride;
  double* stats;         // pass 2
  int want_second;
  int flush_tiles;
};

// ---------------------------------------------------------------------------
// pass 1: per-chunk partial log-sum-exp.  Frames on the TMEM lanes.
//   D1[128 frames, 128 mixtures] = Ahi*Whi + Alo*Whi + Ahi*Wlo      (K = 128)
//   A (hi | lo) is written into TMEM by the producers one 32-column k-block at
//   a time (single buffer, one full/empty mbarrier pair per k-block, so block
//   kb of the next tile is refilled while the tensor core works on kb+1..3);
//   W (hi | lo) sits in shared memory for the whole kernel; the epilogue runs an
//   online max / sum over each thread's own 128-column row.
// TMEM: [0,128) Ahi | [128,256) Alo | [256,384) D1[0] | [384,512) D1[1]
// ---------------------------------------------------------------------------
namespace tcl {
constexpr int TF = 128;        // frames per tile
constexpr int THREADS = 288;   // warps 0-3 epilogue, 4 MMA issuer, 5-8 producers
constexpr uint32_t TM_AHI = 0, TM_ALO = 128, TM_D1 = 256;
constexpr uint32_t SM_WHS = 0, SM_WLS = 65536, SM_XRAW = 131072;      // 2 x 128 x 60 floats
constexpr uint32_t SM_BAR = SM_XRAW + 2 * TF * tc::MAX_D * 4;
constexpr uint32_t SMEM_BYTES = SM_BAR + 256 + 1024;
constexpr int B_A_FULL = 0, B_A_EMPTY = 4, B_D1_FULL = 8, B_D1_EMPTY = 10, B_TMEM_PTR = 16;
}  // namespace tcl

__global__ void __launch_bounds__(tcl::THREADS, 1) gmm_tc_lse_kernel(TcArgs a) {
  using namespace tc;
  using namespace tcl;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw_addr = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  unsigned char* smem = smem_dyn + pad;
  const uint32_t sbase = raw_addr + pad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x;
  const int D = a.D, d4 = a.D >> 2;
  const int64_t n_tiles = (a.N + TF - 1) / TF;
  const int64_t my_tiles = (n_tiles > (int64_t)blockIdx.y) ? (n_tiles - blockIdx.y + gridDim.y - 1) / gridDim.y : 0;
  auto bar = [&](int i) -> uint32_t { return sbase + tcl::SM_BAR + 8u * i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + tcl::SM_BAR + 8 * tcl::B_TMEM_PTR);

  if (warp == 4) {
    if (lane == 0) {
      for (int i = 0; i < 4; ++i) { mbar_init(bar(B_A_FULL + i), 128); mbar_init(bar(B_A_EMPTY + i), 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(bar(tcl::B_D1_FULL + i), 1); mbar_init(bar(tcl::B_D1_EMPTY + i), 128); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc512(bar(tcl::B_TMEM_PTR));
  }
  {  // W chunk images (hi | lo, already swizzled): global -> shared
    const uint4* s0 = reinterpret_cast<const uint4*>(a.Whs + (size_t)chunk * CM * K);
    const uint4* s1 = reinterpret_cast<const uint4*>(a.Wls + (size_t)chunk * CM * K);
    uint4* d0 = reinterpret_cast<uint4*>(smem + SM_WHS);
    uint4* d1 = reinterpret_cast<uint4*>(smem + SM_WLS);
    for (int i = tid; i < CM * K / 4; i += tcl::THREADS) { d0[i] = __ldg(s0 + i); d1[i] = __ldg(s1 + i); }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_s;

  if (warp >= 5) {
    // =========================================================== producers
    const int ptid = tid - 160;                 // 0..127
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;        // frame within the tile == TMEM lane
    const uint32_t lane_field = (uint32_t)(quarter * 32) << 16;
    const int items = TF * d4;                  // float4 items of one raw tile (<= 1920)
    const float4* X4 = reinterpret_cast<const float4*>(a.X);
    const int64_t total4 = a.N * d4;
    auto fetch = [&](int64_t it) {
      const int64_t tile = blockIdx.y + it * gridDim.y;
      const int64_t base4 = tile * items;
      const uint32_t dst = sbase + tcl::SM_XRAW + (uint32_t)(it & 1) * (TF * MAX_D * 4);
      for (int idx = ptid; idx < items; idx += 128) {
        const int64_t g = base4 + idx;
        const bool ok = g < total4;
        cp_async16(dst + idx * 16, X4 + (ok ? g : 0), ok ? 16u : 0u);
      }
      cp_async_commit();
    };
    if (my_tiles > 0) fetch(0);
    for (int64_t it = 0; it < my_tiles; ++it) {
      cp_async_wait_all();
      named_bar_sync(1, 128);   // tile `it` has landed; everyone is done with the other buffer
      if (it + 1 < my_tiles) fetch(it + 1);
      const float4* xr4 = reinterpret_cast<const float4*>(smem + tcl::SM_XRAW + (it & 1) * (TF * MAX_D * 4)) + row * d4;
      for (int kb = 0; kb < 4; ++kb) {
        float h[32], l[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int k4 = 8 * kb + c;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k4 < d4) {
            const float4 x = xr4[k4];
            v = make_float4(x.x * x.x, x.y * x.y, x.z * x.z, x.w * x.w);
          } else if (k4 < 2 * d4) {
            v = xr4[k4 - d4];
          } else if (k4 == K_ONE / 4) {
            v.x = 1.f;
          }
          float4 hh, ll;
          split4(v, hh, ll);
          h[4 * c] = hh.x; h[4 * c + 1] = hh.y; h[4 * c + 2] = hh.z; h[4 * c + 3] = hh.w;
          l[4 * c] = ll.x; l[4 * c + 1] = ll.y; l[4 * c + 2] = ll.z; l[4 * c + 3] = ll.w;
        }
        mbar_wait(bar(B_A_EMPTY + kb), (uint32_t)((it & 1) ^ 1));
        tc_fence_after();
        tmem_st32(tmem + TM_AHI + 32 * kb + lane_field, h);
        tmem_st32(tmem + TM_ALO + 32 * kb + lane_field, l);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar(B_A_FULL + kb));
      }
    }
  } else if (warp == 4) {
    // ========================================================== MMA issuer
    if (my_tiles > 0) {
      const uint32_t idesc = idesc_tf32(CM);
      const uint64_t whi0 = desc_k_sw128(sbase + SM_WHS), wlo0 = desc_k_sw128(sbase + SM_WLS);
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int buf = (int)(it & 1);
        mbar_wait(bar(tcl::B_D1_EMPTY + buf), (uint32_t)(((it >> 1) & 1) ^ 1));
        const uint32_t d = tmem + tcl::TM_D1 + 128 * buf;
        for (int kb = 0; kb < 4; ++kb) {
          mbar_wait(bar(B_A_FULL + kb), (uint32_t)(it & 1));
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int kk = 4 * kb + q;
              const uint32_t a_hi = tmem + TM_AHI + kk * 8, a_lo = tmem + TM_ALO + kk * 8;
              const uint64_t off = (uint64_t)((kb * 16384 + q * 32) >> 4);
              mma_ts(d, a_hi, whi0 + off, idesc, kk > 0 ? 1u : 0u);
              mma_ts(d, a_lo, whi0 + off, idesc, 1u);
              mma_ts(d, a_hi, wlo0 + off, idesc, 1u);
            }
            tc_commit(bar(B_A_EMPTY + kb));
            if (kb == 3) tc_commit(bar(tcl::B_D1_FULL + buf));
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ============================================================ epilogue
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int buf = (int)(it & 1);
      mbar_wait(bar(tcl::B_D1_FULL + buf), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float m = 0.f, s = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v[32];
        tmem_ld32(tmem + tcl::TM_D1 + 128 * buf + 32 * g + lane_field, v);
        float mx = v[0];
#pragma unroll
        for (int i = 1; i < 32; ++i) mx = fmaxf(mx, v[i]);
        float acc = 0.f;
        if (g == 0) {
          m = mx;
        } else {
          const float mn = fmaxf(m, mx);
          acc = s * ex2f(m - mn);
          m = mn;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) acc += ex2f(v[i] - m);
        s = acc;
      }
      tc_fence_before();
      mbar_arrive(bar(tcl::B_D1_EMPTY + buf));
      const int64_t tile = blockIdx.y + it * gridDim.y;
      const int64_t f = tile * TF + warp * 32 + lane;
      if (f < a.N) a.part[(size_t)chunk * a.part_stride + f] = make_float2(m, s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    tmem_dealloc512(tmem);
  }
}

// ---------------------------------------------------------------------------
// pass 2: posteriors and statistics.  Mixtures on the TMEM lanes.
//   D1[128 mix, 32 frames] = Whi*Bhi + Whi*Blo + Wlo*Bhi       (GEMM 1, K = 128;
//       Whi in TMEM for the whole kernel, Wlo and the frame tile B in smem)
//   P = ex2(D1)  (the -lse2 column of B makes D1 = lp2 - lse2)
//   D2[128 j, 128 mix] += Thi*Phi + Thi*Plo + Tlo*Phi           (GEMM 2, K = 32 frames;
//       T^T = [x^2 | x | 1] transposed, written into TMEM by the producers,
//       P from a K-major SWIZZLE_128B smem tile written by the epilogue)
//   every `flush_tiles` tiles D2 is drained into the fp64 statistics.
// 17 warps: 0-7 epilogue, 8 MMA issuer, 9-16 producers (two warps per TMEM lane
// quarter in each group: the roles are instruction-issue bound, not MMA bound,
// with one warp per scheduler).
// TMEM: [0,128) Whi | [128,256) D2 | [256,384) 4 x D1 | [384,512) 2 x (Thi | Tlo)
// ---------------------------------------------------------------------------
namespace tcs {
constexpr int THREADS = 544;
constexpr int NROLE = 256;     // threads per producer / epilogue group
}

__global__ void __launch_bounds__(tcs::THREADS, 1) gmm_tc_stats_kernel(TcArgs a) {
  using namespace tc;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw_addr = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  unsigned char* smem = smem_dyn + pad;
  const uint32_t sbase = raw_addr + pad;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x;
  const int D = a.D, d4 = a.D >> 2;
  const int64_t n_tiles = (a.N + NF - 1) / NF;
  const int64_t my_tiles = (n_tiles > (int64_t)blockIdx.y) ? (n_tiles - blockIdx.y + gridDim.y - 1) / gridDim.y : 0;
  auto bar = [&](int i) -> uint32_t { return sbase + SM_BAR + 8u * i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + SM_BAR + 8 * B_TMEM_PTR);

  // ------------------------------------------------------------- setup
  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < BT_STAGES; ++i) { mbar_init(bar(B_BT_FULL + i), tcs::NROLE); mbar_init(bar(B_BT_EMPTY + i), 1); }
      for (int i = 0; i < D1_BUFS; ++i) { mbar_init(bar(B_D1_FULL + i), 1); mbar_init(bar(B_D1_EMPTY + i), tcs::NROLE); }
      for (int i = 0; i < TT_BUFS; ++i) { mbar_init(bar(B_TT_FULL + i), tcs::NROLE); mbar_init(bar(B_TT_EMPTY + i), 1); }
      for (int i = 0; i < P_BUFS; ++i) { mbar_init(bar(B_P_FULL + i), tcs::NROLE); mbar_init(bar(B_P_EMPTY + i), 1); }
      mbar_init(bar(B_D2_FULL), 1);
      mbar_init(bar(B_D2_EMPTY), tcs::NROLE);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc512(bar(B_TMEM_PTR));
  }
  {  // Wlo chunk image: global -> shared (already swizzled)
    const uint4* src = reinterpret_cast<const uint4*>(a.Wls + (size_t)chunk * CM * K);
    uint4* dst = reinterpret_cast<uint4*>(smem + SM_WLO);
    for (int i = tid; i < CM * K / 4; i += tcs::THREADS) dst[i] = __ldg(src + i);
    // constant columns 2D..127 of both frame-tile stages: zeros and the 1 of column 120
    const int nconst = 32 - 2 * d4;
    for (int i = tid; i < BT_STAGES * 2 * NF * nconst; i += tcs::THREADS) {
      const int k4 = 2 * d4 + i % nconst;
      int t = i / nconst;
      const int r = t % NF; t /= NF;
      const int part = t & 1, s = t >> 1;   // part 0 = hi, 1 = lo
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (part == 0 && k4 == K_ONE / 4) v.x = 1.f;
      *reinterpret_cast<float4*>(smem + SM_BT + s * 32768 + part * 16384 + sw128_off(r, k4, NF)) = v;
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_s;
  if (warp < 8) {  // Whi rows -> TMEM columns [0,128): lane = mixture row, 64 columns per warp
    const int row = (warp & 3) * 32 + lane, half = warp >> 2;
    const float4* src = reinterpret_cast<const float4*>(a.Whi + ((size_t)chunk * CM + row) * K + 64 * half);
    for (int c = 0; c < 2; ++c) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = __ldg(src + c * 8 + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
      tmem_st32(tmem + TM_WHI + 64 * half + 32 * c + ((uint32_t)((warp & 3) * 32) << 16), v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp >= 9) {
    // =========================================================== producers
    const int ptid = tid - 288;               // 0..255
    const int quarter = warp & 3;             // TMEM lane quarter this warp may touch
    const int half = (warp - 9) >> 2;         // which 16 frames of the transposed operand
    const int j = quarter * 32 + lane;        // row of the transposed operand
    const int items = NF * d4;                // float4 items of one raw tile (<= 480)
    const float4* X4 = reinterpret_cast<const float4*>(a.X);
    const int64_t total4 = a.N * d4;
    // item -> swizzled offsets of its x^2 and x chunks (the same for every tile)
    uint32_t off_sq[2], off_x[2];
    bool valid[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int idx = ptid + 256 * u;
      valid[u] = idx < items;
      const int r = valid[u] ? idx / d4 : 0, i = valid[u] ? idx - r * d4 : 0;
      off_sq[u] = sw128_off(r, i, NF);
      off_x[u] = sw128_off(r, d4 + i, NF);
    }
    const uint32_t off_lse = sw128_off(ptid & 31, K_ONE / 4, NF);
    // transposed operand: which value this row carries
    const bool t_sq = j < D, t_x = j >= D && j < 2 * D;
    const int t_d = t_sq ? j : (t_x ? j - D : 0);
    const float t_const = (j == K_ONE) ? 1.f : 0.f;
    float4 pre[2];
    float lse_next = 0.f;
    auto prefetch = [&](int64_t it) {
      const int64_t tile = blockIdx.y + it * gridDim.y;
      const int64_t base4 = tile * items;
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int64_t g = base4 + ptid + 256 * u;
        pre[u] = (valid[u] && g < total4) ? __ldg(X4 + g) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (ptid < NF) {
        const int64_t f = tile * NF + ptid;
        const bool on = f < a.N && (a.sad == nullptr || a.sad[f] != 0);
        lse_next = on ? __ldg(a.lse2 + f) : 1e30f;   // 2^(lp - 1e30) = 0: masked / out of range
      }
    };
    if (my_tiles > 0) prefetch(0);
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int buf = (int)(it & 1);
      float4* xr4 = reinterpret_cast<float4*>(smem + SM_XRAW + buf * (NF * MAX_D * 4));
      const float* xr = reinterpret_cast<const float*>(xr4);
      const float4 cur0 = pre[0], cur1 = pre[1];
      const float lse_cur = lse_next;
      if (valid[0]) xr4[ptid] = cur0;
      if (valid[1]) xr4[ptid + 256] = cur1;
      if (it + 1 < my_tiles) prefetch(it + 1);
      named_bar_sync(1, tcs::NROLE);
      // ---- (a) GEMM-1 operand tile: [frame r][k] K-major, SWIZZLE_128B, hi | lo
      {
        const int s = (int)(it % BT_STAGES);
        mbar_wait(bar(B_BT_EMPTY + s), (uint32_t)(((it / BT_STAGES) & 1) ^ 1));
        unsigned char* hi = smem + SM_BT + s * 32768;
        unsigned char* lo = hi + 16384;
        auto put = [&](uint32_t off, const float4& v) {
          float4 h, l;
          split4(v, h, l);
          *reinterpret_cast<float4*>(hi + off) = h;
          *reinterpret_cast<float4*>(lo + off) = l;
        };
        if (valid[0]) {
          put(off_sq[0], make_float4(cur0.x * cur0.x, cur0.y * cur0.y, cur0.z * cur0.z, cur0.w * cur0.w));
          put(off_x[0], cur0);
        }
        if (valid[1]) {
          put(off_sq[1], make_float4(cur1.x * cur1.x, cur1.y * cur1.y, cur1.z * cur1.z, cur1.w * cur1.w));
          put(off_x[1], cur1);
        }
        if (ptid < NF) put(off_lse, make_float4(1.f, -lse_cur, 0.f, 0.f));
        fence_proxy_async();
        mbar_arrive(bar(B_BT_FULL + s));
      }
      // ---- (b) GEMM-2 operand: T^T[j][frame] hi | lo -> TMEM (lane = j), 16 frames per warp
      {
        const int ts = (int)(it % TT_BUFS);
        float h[16], l[16];
        if (t_sq || t_x) {
#pragma unroll
          for (int b = 0; b < 16; ++b) {
            const float x = xr[(16 * half + b) * D + t_d];
            split_tf32(t_sq ? x * x : x, h[b], l[b]);
          }
        } else {
#pragma unroll
          for (int b = 0; b < 16; ++b) { h[b] = t_const; l[b] = 0.f; }
        }
        mbar_wait(bar(B_TT_EMPTY + ts), (uint32_t)(((it / TT_BUFS) & 1) ^ 1));
        tc_fence_after();
        const uint32_t taddr = tmem + TM_TT + 64 * ts + 16 * half + ((uint32_t)(quarter * 32) << 16);
        tmem_st16(taddr, h);
        tmem_st16(taddr + 32, l);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar(B_TT_FULL + ts));
      }
    }
  } else if (warp == 8) {
    // ========================================================== MMA issuer
    // The whole warp walks the (warp-uniform) schedule; one elected lane issues.
    if (my_tiles > 0) {
      const uint32_t idesc1 = idesc_tf32(NF), idesc2 = idesc_tf32(CM);
      const uint64_t wlo_desc0 = desc_k_sw128(sbase + SM_WLO);
      bool acc = false, need_d2_empty = false;
      uint32_t d2_phase = 0;
      auto issue_g1 = [&](int64_t it) {
        const int s = (int)(it % BT_STAGES), b = (int)(it % D1_BUFS);
        mbar_wait(bar(B_BT_FULL + s), (uint32_t)((it / BT_STAGES) & 1));
        mbar_wait(bar(B_D1_EMPTY + b), (uint32_t)(((it / D1_BUFS) & 1) ^ 1));
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem + TM_D1 + 32 * b;
          // descriptors advance by adding the byte offset >> 4 to the start-address field
          const uint64_t bt_hi0 = desc_k_sw128(sbase + SM_BT + s * 32768), bt_lo0 = bt_hi0 + (16384 >> 4);
#pragma unroll
          for (int kk = 0; kk < K / 8; ++kk) {
            const uint32_t a_hi = tmem + TM_WHI + kk * 8;
            const uint64_t a_lo = wlo_desc0 + (uint64_t)(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4);
            const uint64_t b_hi = bt_hi0 + (uint64_t)(((kk >> 2) * 4096 + (kk & 3) * 32) >> 4);
            const uint64_t b_lo = bt_lo0 + (uint64_t)(((kk >> 2) * 4096 + (kk & 3) * 32) >> 4);
            mma_ts(d, a_hi, b_hi, idesc1, kk > 0 ? 1u : 0u);
            mma_ts(d, a_hi, b_lo, idesc1, 1u);
            mma_ss(d, a_lo, b_hi, idesc1, 1u);
          }
          tc_commit(bar(B_D1_FULL + b));
          tc_commit(bar(B_BT_EMPTY + s));
        }
        __syncwarp();
      };
      auto issue_g2 = [&](int64_t it) {
        const int pb = (int)(it % P_BUFS), ts = (int)(it % TT_BUFS);
        mbar_wait(bar(B_P_FULL + pb), (uint32_t)((it / P_BUFS) & 1));
        mbar_wait(bar(B_TT_FULL + ts), (uint32_t)((it / TT_BUFS) & 1));
        if (need_d2_empty) {
          mbar_wait(bar(B_D2_EMPTY), d2_phase);
          d2_phase ^= 1;
          need_d2_empty = false;
        }
        tc_fence_after();
        const bool flush = ((it + 1) % a.flush_tiles) == 0 || it + 1 == my_tiles;
        if (elect_one()) {
          const uint32_t d = tmem + TM_D2;
          const uint64_t p_hi0 = desc_k_sw128(sbase + SM_P + pb * 32768), p_lo0 = p_hi0 + (16384 >> 4);
#pragma unroll
          for (int ks = 0; ks < NF / 8; ++ks) {
            const uint32_t t_hi = tmem + TM_TT + 64 * ts + ks * 8, t_lo = t_hi + 32;
            const uint64_t b_hi = p_hi0 + (uint64_t)((ks * 32) >> 4), b_lo = p_lo0 + (uint64_t)((ks * 32) >> 4);
            mma_ts(d, t_hi, b_hi, idesc2, (acc || ks > 0) ? 1u : 0u);
            mma_ts(d, t_hi, b_lo, idesc2, 1u);
            mma_ts(d, t_lo, b_hi, idesc2, 1u);
          }
          tc_commit(bar(B_P_EMPTY + pb));
          tc_commit(bar(B_TT_EMPTY + ts));
          if (flush) tc_commit(bar(B_D2_FULL));
        }
        __syncwarp();
        acc = !flush;
        if (flush) need_d2_empty = true;
      };
      issue_g1(0);
      for (int64_t it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) issue_g1(it + 1);
        issue_g2(it);
      }
    }
  } else {
    // ============================================================ epilogue
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;                   // mixture within the chunk == TMEM lane
    const uint32_t lane_field = (uint32_t)(quarter * 32) << 16;
    uint32_t p_off[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) p_off[q] = sw128_off(row, 4 * half + q, CM);
    uint32_t d2_phase = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int b = (int)(it % D1_BUFS);
      mbar_wait(bar(B_D1_FULL + b), (uint32_t)((it / D1_BUFS) & 1));
      tc_fence_after();
      float v[16];
      tmem_ld16(tmem + TM_D1 + 32 * b + 16 * half + lane_field, v);
      tc_fence_before();
      mbar_arrive(bar(B_D1_EMPTY + b));
      float h[16], l[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) split_tf32(ex2f(v[i]), h[i], l[i]);
      const int pb = (int)(it % P_BUFS);
      mbar_wait(bar(B_P_EMPTY + pb), (uint32_t)(((it / P_BUFS) & 1) ^ 1));
      unsigned char* hi = smem + SM_P + pb * 32768;
      unsigned char* lo = hi + 16384;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        *reinterpret_cast<float4*>(hi + p_off[q]) = make_float4(h[4 * q], h[4 * q + 1], h[4 * q + 2], h[4 * q + 3]);
        *reinterpret_cast<float4*>(lo + p_off[q]) = make_float4(l[4 * q], l[4 * q + 1], l[4 * q + 2], l[4 * q + 3]);
      }
      fence_proxy_async();
      mbar_arrive(bar(B_P_FULL + pb));
      if (((it + 1) % a.flush_tiles) == 0 || it + 1 == my_tiles) {
        // drain D2[j = row, mixture column] into the fp64 statistics (64 columns per warp)
        mbar_wait(bar(B_D2_FULL), d2_phase);
        d2_phase ^= 1;
        tc_fence_after();
        double* dst = nullptr;
        if (row < D) { if (a.want_second) dst = a.stats + a.M + (size_t)D * a.M + (size_t)row * a.M; }
        else if (row < 2 * D) dst = a.stats + a.M + (size_t)(row - D) * a.M;
        else if (row == K_ONE) dst = a.stats;
        for (int c = 0; c < 2; ++c) {
          float s[32];
          tmem_ld32(tmem + TM_D2 + 64 * half + 32 * c + lane_field, s);
          if (dst != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int m = chunk * CM + 64 * half + 32 * c + i;
              if (m < a.M) atomicAdd(dst + m, (double)s[i]);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar(B_D2_EMPTY));
      }
    }
  }

  // ------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc512(tmem);
  }
}

// pass 1 tail: per-chunk partials -> lse2[b]; sum of log-likelihoods and frame count
__global__ void __launch_bounds__(256)
gmm_tc_combine_kernel(const float2* __restrict__ part, int nchunks, int64_t stride, int64_t n,
                      const uint8_t* __restrict__ sad, float* __restrict__ lse2, double* __restrict__ stat_L) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double lsum = 0.0, lcnt = 0.0;
  if (b < n) {
    float mx = -INFINITY;
    for (int c = 0; c < nchunks; ++c) mx = fmaxf(mx, part[(size_t)c * stride + b].x);
    float s = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const float2 p = part[(size_t)c * stride + b];
      s += p.y * exp2f(p.x - mx);
    }
    const float l2 = mx + log2f(s);
    lse2[b] = l2;
    if (sad == nullptr || sad[b] != 0) { lsum = (double)l2 * 0.69314718055994530942; lcnt = 1.0; }
  }
  if (stat_L == nullptr) return;
  __shared__ double red[8][2];
  lsum = warp_sum(lsum);
  lcnt = warp_sum(lcnt);
  if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5][0] = lsum; red[threadIdx.x >> 5][1] = lcnt; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s0 = 0, s1 = 0;
    for (int i = 0; i < 8; ++i) { s0 += red[i][0]; s1 += red[i][1]; }
    if (s1 > 0) { atomicAdd(stat_L, s0); atomicAdd(stat_L + 1, s1); }
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool gmm_tc_supported(const odin_gmm* g) {
  static const int min_m = [] { const char* e = getenv("ODIN_GMM_TC_MIN_M"); return e ? atoi(e) : 96; }();
  return g->D % 4 == 0 && g->D <= tc::MAX_D && g->D >= 4 && g->M >= min_m;
}

int gmm_tc_refresh(odin_gmm* g, cudaStream_t st) {
  gmm_tc_prepare_kernel<<<ceil_div(g->Mpad, 128), 128, 0, st>>>(g->d_mean, g->d_var, g->d_w, g->D, g->M, g->Mpad,
                                                                g->d_Whi, g->d_Whs, g->d_Wlo);
  ODIN_LAUNCH_CHECK("gmm_tc_prepare_kernel");
  return ODIN_OK;
}

// Tunables (read at every call so tests can vary them): tiles between drains of
// the fp32 TMEM accumulator into the fp64 statistics, frames per pass-1 launch.
static int tc_flush_tiles() {
  const char* e = getenv("ODIN_TC_FLUSH_TILES");
  int v = e ? atoi(e) : 512;
  return v < 1 ? 1 : v;
}
static int64_t tc_sub_batch() {
  const char* e = getenv("ODIN_TC_SUB_BATCH");
  int64_t v = e ? atoll(e) : (int64_t)1 << 20;  // bounds the partial-LSE workspace
  return v < tcl::TF ? tcl::TF : v;
}

static dim3 tc_grid(const odin_gmm* g, int64_t n_tiles) {
  const int nchunks = g->Mpad / tc::CM;
  int64_t splits = std::max<int64_t>(1, sm_count() / nchunks);
  splits = std::min<int64_t>(splits, n_tiles);
  return dim3(nchunks, (unsigned)splits);
}

int gmm_lse_tc(odin_gmm* g, const float* X, const uint8_t* sad, int64_t N, float* lse, double* stats,
               cudaStream_t st) {
  if (N <= 0) return ODIN_OK;
  const int nchunks = g->Mpad / tc::CM;
  const int64_t sub = std::min<int64_t>(N, tc_sub_batch());
  const int64_t stride = ceil_div<int64_t>(sub, tcl::TF) * tcl::TF;
  const int64_t need = stride * nchunks;
  if (need > g->part_cap) {
    if (g->d_part) ODIN_CUDA_CHECK(cudaFree(g->d_part));
    g->d_part = nullptr;
    g->part_cap = 0;
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_part, need * sizeof(float2)));
    g->part_cap = need;
  }
  ODIN_CUDA_CHECK(cudaFuncSetAttribute(gmm_tc_lse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tcl::SMEM_BYTES));
  double* statL = stats ? stats + (stats_size(g->D, g->M) - 2) : nullptr;
  for (int64_t s0 = 0; s0 < N; s0 += sub) {
    const int64_t n = std::min<int64_t>(sub, N - s0);
    TcArgs a{};
    a.X = X + s0 * g->D; a.N = n; a.D = g->D; a.M = g->M;
    a.Whi = g->d_Whi; a.Whs = g->d_Whs; a.Wls = g->d_Wlo;
    a.part = reinterpret_cast<float2*>(g->d_part); a.part_stride = stride;
    gmm_tc_lse_kernel<<<tc_grid(g, ceil_div<int64_t>(n, tcl::TF)), tcl::THREADS, tcl::SMEM_BYTES, st>>>(a);
    ODIN_LAUNCH_CHECK("gmm_tc_lse_kernel");
    gmm_tc_combine_kernel<<<(unsigned)ceil_div<int64_t>(n, 256), 256, 0, st>>>(
        reinterpret_cast<const float2*>(g->d_part), nchunks, stride, n, sad ? sad + s0 : nullptr, lse + s0, statL);
    ODIN_LAUNCH_CHECK("gmm_tc_combine_kernel");
  }
  return ODIN_OK;
}

int gmm_stats_tc(odin_gmm* g, const float* X, const uint8_t* sad, int64_t N, const float* lse, int want_second,
                 double* stats, cudaStream_t st) {
  if (N <= 0) return ODIN_OK;
  TcArgs a{};
  a.X = X; a.sad = sad; a.N = N; a.D = g->D; a.M = g->M;
  a.Whi = g->d_Whi; a.Whs = g->d_Whs; a.Wls = g->d_Wlo; a.lse2 = lse; a.stats = stats; a.want_second = want_second;
  a.flush_tiles = tc_flush_tiles();
  ODIN_CUDA_CHECK(cudaFuncSetAttribute(gmm_tc_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)tc::SMEM_BYTES));
  gmm_tc_stats_kernel<<<tc_grid(g, ceil_div<int64_t>(N, tc::NF)), tcs::THREADS, tc::SMEM_BYTES, st>>>(a);
  ODIN_LAUNCH_CHECK("gmm_tc_stats_kernel");
  return ODIN_OK;
}

}  // namespace odin
