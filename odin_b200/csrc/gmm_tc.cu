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
constexpr int THREADS = 288;    // warps 0-3 epilogue, 4 MMA issuer, 5-8 producers
constexpr int MAX_D = 60;       // 2*D + 2 <= 122 and D % 4 == 0

constexpr uint32_t TM_WHI = 0, TM_D2 = 128, TM_D1 = 256, TM_TT = 384;

// shared memory map (after 1024-byte alignment)
constexpr uint32_t SM_WLO = 0;                          // 64 KB  [4 kblocks][128 rows][128 B]
constexpr uint32_t SM_BT = 65536;                       // 2 x (hi 16 KB | lo 16 KB)
constexpr uint32_t SM_P = SM_BT + BT_STAGES * 32768;    // 2 x (hi 16 KB | lo 16 KB)
constexpr uint32_t SM_XRAW = SM_P + P_BUFS * 32768;     // 2 x 32 x 60 floats
constexpr uint32_t SM_LSE = SM_XRAW + 2 * NF * MAX_D * 4;   // 2 x 32 floats
constexpr uint32_t SM_PART = SM_LSE + 2 * NF * 4;       // 2 x 4 x 32 float2
constexpr uint32_t SM_BAR = SM_PART + 2 * 4 * NF * 8;   // mbarriers
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

__device__ __forceinline__ float tf32_rna(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#define ODIN_R32(v) \
  v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15], v[16], \
  v[17], v[18], v[19], v[20], v[21], v[22], v[23], v[24], v[25], v[26], v[27], v[28], v[29], v[30], v[31]

// 32 consecutive TMEM columns of this thread's lane -> registers
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
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace tc

// ---------------------------------------------------------------------------
// operand images of the model (refreshed with the cached constants)
// ---------------------------------------------------------------------------
// Whi : [Mpad][128] plain rows (copied to TMEM by the kernels)
// Wlo : per chunk of 128 mixtures the exact shared-memory image of a K-major
//       SWIZZLE_128B tile: [kblock 4][row 128][16-byte chunk (q ^ (row & 7))][4]
__global__ void gmm_tc_prepare_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                                      const float* __restrict__ w, int D, int M, int Mpad,
                                      float* __restrict__ Whi, float* __restrict__ Wlo) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mpad) return;
  const double LOG2E = 1.4426950408889634074;
  const int chunk = m / tc::CM, row = m % tc::CM;
  float* lo_base = Wlo + (size_t)chunk * tc::CM * tc::K;
  auto put = [&](int k, double v) {
    const float vf = (float)v;
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(vf));
    const float h = __uint_as_float(u);
    float l = vf - h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(l));
    l = __uint_as_float(u);
    Whi[(size_t)m * tc::K + k] = h;
    const int kb = k >> 5, q = (k & 31) >> 2, e = k & 3;
    lo_base[(size_t)kb * (tc::CM * 32) + row * 32 + ((q ^ (row & 7)) << 2) + e] = l;
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

// ---------------------------------------------------------------------------
// the tensor-core kernel (STATS = false: pass 1 / LSE partials, true: pass 2)
// ---------------------------------------------------------------------------
struct TcArgs {
  const float* X;
  const uint8_t* sad;
  int64_t N;
  int D, M;
  const float* Whi;
  const float* Wlo;
  const float* lse2;     // pass 2: per-frame log2-domain log-sum-exp
  float2* part;          // pass 1: [nchunks][part_stride] (max, sum) per frame
  int64_t part_stride;
  double* stats;         // pass 2
  int want_second;
  int flush_tiles;
};

template <bool STATS>
__global__ void __launch_bounds__(tc::THREADS, 1) gmm_tc_kernel(TcArgs a) {
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
  if (warp == 4) {
    if (lane == 0) {
      for (int i = 0; i < BT_STAGES; ++i) { mbar_init(bar(B_BT_FULL + i), 128); mbar_init(bar(B_BT_EMPTY + i), 1); }
      for (int i = 0; i < D1_BUFS; ++i) { mbar_init(bar(B_D1_FULL + i), 1); mbar_init(bar(B_D1_EMPTY + i), 128); }
      for (int i = 0; i < TT_BUFS; ++i) { mbar_init(bar(B_TT_FULL + i), 128); mbar_init(bar(B_TT_EMPTY + i), 1); }
      for (int i = 0; i < P_BUFS; ++i) { mbar_init(bar(B_P_FULL + i), 128); mbar_init(bar(B_P_EMPTY + i), 1); }
      mbar_init(bar(B_D2_FULL), 1);
      mbar_init(bar(B_D2_EMPTY), 128);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(bar(B_TMEM_PTR)), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // Wlo chunk image: global -> shared (already swizzled)
    const uint4* src = reinterpret_cast<const uint4*>(a.Wlo + (size_t)chunk * CM * K);
    uint4* dst = reinterpret_cast<uint4*>(smem + SM_WLO);
    for (int i = tid; i < CM * K / 4; i += THREADS) dst[i] = __ldg(src + i);
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_s;
  if (warp < 4) {  // Whi rows -> TMEM columns [0,128): lane = mixture row
    const int row = warp * 32 + lane;
    const float4* src = reinterpret_cast<const float4*>(a.Whi + ((size_t)chunk * CM + row) * K);
    for (int c = 0; c < 4; ++c) {
      float v[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 t = __ldg(src + c * 8 + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
      tmem_st32(tmem + TM_WHI + 32 * c + ((uint32_t)(warp * 32) << 16), v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp >= 5) {
    // =========================================================== producers
    const int ptid = tid - 160;              // 0..127
    const int quarter = warp & 3;            // TMEM lane quarter this warp may touch
    const int j = quarter * 32 + lane;       // row of the transposed operand
    const int items = NF * d4;               // float4 items of one raw tile
    const float4* X4 = reinterpret_cast<const float4*>(a.X);
    const int64_t total4 = a.N * d4;
    float4 pre[4];
    float lse_next = 0.f;
    auto prefetch = [&](int64_t it) {
      const int64_t tile = blockIdx.y + it * gridDim.y;
      const int64_t base4 = tile * items;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = ptid + 128 * u;
        const int64_t g = base4 + idx;
        pre[u] = (idx < items && g < total4) ? __ldg(X4 + g) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (STATS && ptid < NF) {
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
      float* lse_s = reinterpret_cast<float*>(smem + SM_LSE + buf * (NF * 4));
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = ptid + 128 * u;
        if (idx < items) xr4[idx] = pre[u];
      }
      if (STATS && ptid < NF) lse_s[ptid] = lse_next;
      if (it + 1 < my_tiles) prefetch(it + 1);
      named_bar_sync(1, 128);
      // ---- (a) GEMM-1 operand tile: [frame r][k] K-major, SWIZZLE_128B, hi | lo
      {
        const int s = (int)(it % BT_STAGES);
        mbar_wait(bar(B_BT_EMPTY + s), (uint32_t)(((it / BT_STAGES) & 1) ^ 1));
        unsigned char* hi = smem + SM_BT + s * 32768;
        unsigned char* lo = hi + 16384;
        auto put4 = [&](int r, int k4, float4 v) {   // k4 = k / 4
          float4 h, l;
          h.x = tf32_rna(v.x); h.y = tf32_rna(v.y); h.z = tf32_rna(v.z); h.w = tf32_rna(v.w);
          l.x = tf32_rna(v.x - h.x); l.y = tf32_rna(v.y - h.y); l.z = tf32_rna(v.z - h.z); l.w = tf32_rna(v.w - h.w);
          const uint32_t off = (uint32_t)(k4 >> 3) * 4096u + (uint32_t)r * 128u + (uint32_t)(((k4 & 7) ^ (r & 7)) << 4);
          *reinterpret_cast<float4*>(hi + off) = h;
          *reinterpret_cast<float4*>(lo + off) = l;
        };
        for (int idx = ptid; idx < items; idx += 128) {
          const int r = idx / d4, i = idx - r * d4;
          const float4 x = xr4[idx];
          put4(r, i, make_float4(x.x * x.x, x.y * x.y, x.z * x.z, x.w * x.w));
          put4(r, d4 + i, x);
        }
        // columns 2D .. 127: zeros, the constant 1 and -lse2
        const int k4_first = 2 * d4;
        for (int idx = ptid; idx < NF * (32 - k4_first); idx += 128) {
          const int r = idx / (32 - k4_first), k4 = k4_first + (idx - r * (32 - k4_first));
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (k4 == K_ONE / 4) { v.x = 1.f; v.y = STATS ? -lse_s[r] : 0.f; }
          put4(r, k4, v);
        }
        fence_proxy_async();
        mbar_arrive(bar(B_BT_FULL + s));
      }
      // ---- (b) GEMM-2 operand: T^T[j][frame] hi | lo -> TMEM (lane = j)
      if (STATS) {
        const int ts = (int)(it % TT_BUFS);
        mbar_wait(bar(B_TT_EMPTY + ts), (uint32_t)(((it / TT_BUFS) & 1) ^ 1));
        tc_fence_after();
        float h[32], l[32];
        if (j < 2 * D) {
          const bool sq = j < D;
          const int d = sq ? j : j - D;
#pragma unroll
          for (int b = 0; b < 32; ++b) {
            const float x = xr[b * D + d];
            const float v = sq ? x * x : x;
            h[b] = tf32_rna(v);
            l[b] = tf32_rna(v - h[b]);
          }
        } else {
          const float v = (j == K_ONE) ? 1.f : 0.f;
#pragma unroll
          for (int b = 0; b < 32; ++b) { h[b] = v; l[b] = 0.f; }
        }
        const uint32_t taddr = tmem + TM_TT + 64 * ts + ((uint32_t)(quarter * 32) << 16);
        tmem_st32(taddr, h);
        tmem_st32(taddr + 32, l);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(bar(B_TT_FULL + ts));
      }
    }
  } else if (warp == 4) {
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
        if (STATS) issue_g2(it);
      }
    }
  } else {
    // ============================================================ epilogue
    const int row = warp * 32 + lane;                      // mixture within the chunk == TMEM lane
    const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
    uint32_t d2_phase = 0;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int b = (int)(it % D1_BUFS);
      mbar_wait(bar(B_D1_FULL + b), (uint32_t)((it / D1_BUFS) & 1));
      tc_fence_after();
      float v[32];
      tmem_ld32(tmem + TM_D1 + 32 * b + lane_field, v);
      tc_fence_before();
      mbar_arrive(bar(B_D1_EMPTY + b));
      if (STATS) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = ex2f(v[i]);
        const int pb = (int)(it % P_BUFS);
        mbar_wait(bar(B_P_EMPTY + pb), (uint32_t)(((it / P_BUFS) & 1) ^ 1));
        unsigned char* hi = smem + SM_P + pb * 32768 + row * 128;
        unsigned char* lo = hi + 16384;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 h, l;
          h.x = tf32_rna(v[4 * q]); h.y = tf32_rna(v[4 * q + 1]); h.z = tf32_rna(v[4 * q + 2]); h.w = tf32_rna(v[4 * q + 3]);
          l.x = tf32_rna(v[4 * q] - h.x); l.y = tf32_rna(v[4 * q + 1] - h.y);
          l.z = tf32_rna(v[4 * q + 2] - h.z); l.w = tf32_rna(v[4 * q + 3] - h.w);
          const uint32_t off = (uint32_t)((q ^ (row & 7)) << 4);
          *reinterpret_cast<float4*>(hi + off) = h;
          *reinterpret_cast<float4*>(lo + off) = l;
        }
        fence_proxy_async();
        mbar_arrive(bar(B_P_FULL + pb));
        if (((it + 1) % a.flush_tiles) == 0 || it + 1 == my_tiles) {
          // drain D2[j = row, mixture column] into the fp64 statistics
          mbar_wait(bar(B_D2_FULL), d2_phase);
          d2_phase ^= 1;
          tc_fence_after();
          double* dst = nullptr;
          if (row < D) { if (a.want_second) dst = a.stats + a.M + (size_t)D * a.M + (size_t)row * a.M; }
          else if (row < 2 * D) dst = a.stats + a.M + (size_t)(row - D) * a.M;
          else if (row == K_ONE) dst = a.stats;
          for (int c = 0; c < 4; ++c) {
            float s[32];
            tmem_ld32(tmem + TM_D2 + 32 * c + lane_field, s);
            if (dst != nullptr) {
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int m = chunk * CM + 32 * c + i;
                if (m < a.M) atomicAdd(dst + m, (double)s[i]);
              }
            }
          }
          tc_fence_before();
          mbar_arrive(bar(B_D2_EMPTY));
        }
      } else {
        // pass 1: max / sum over the 128 mixture lanes for each of the 32 frames
        float pm = 0.f, ps = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float mx;
          asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(mx) : "f"(v[i]));
          const float e = ex2f(v[i] - mx);                       // in [0, 1]
          const uint32_t fx = __float2uint_rn(e * 67108864.f);   // 2^26 fixed point: warp sum < 2^31
          uint32_t sum;
          asm volatile("redux.sync.add.u32 %0, %1, 0xffffffff;" : "=r"(sum) : "r"(fx));
          if (lane == i) { pm = mx; ps = (float)sum * (1.f / 67108864.f); }
        }
        float2* part_s = reinterpret_cast<float2*>(smem + SM_PART) + (it & 1) * (4 * NF);
        part_s[warp * NF + lane] = make_float2(pm, ps);
        named_bar_sync(2, 128);
        if (warp == 0) {
          float2 p0 = part_s[lane], p1 = part_s[NF + lane], p2 = part_s[2 * NF + lane], p3 = part_s[3 * NF + lane];
          const float mx = fmaxf(fmaxf(p0.x, p1.x), fmaxf(p2.x, p3.x));
          const float sm = p0.y * ex2f(p0.x - mx) + p1.y * ex2f(p1.x - mx) + p2.y * ex2f(p2.x - mx) + p3.y * ex2f(p3.x - mx);
          const int64_t tile = blockIdx.y + it * gridDim.y;
          a.part[(size_t)chunk * a.part_stride + tile * NF + lane] = make_float2(mx, sm);
        }
      }
    }
  }

  // ------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
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
  return g->D % 4 == 0 && g->D <= tc::MAX_D && g->D >= 4 && g->M >= 96;
}

int gmm_tc_refresh(odin_gmm* g, cudaStream_t st) {
  gmm_tc_prepare_kernel<<<ceil_div(g->Mpad, 128), 128, 0, st>>>(g->d_mean, g->d_var, g->d_w, g->D, g->M, g->Mpad,
                                                                g->d_Whi, g->d_Wlo);
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
  return v < tc::NF ? tc::NF : v;
}

template <bool STATS>
static int tc_launch(odin_gmm* g, TcArgs& a, cudaStream_t st) {
  const int nchunks = g->Mpad / tc::CM;
  const int64_t n_tiles = ceil_div<int64_t>(a.N, tc::NF);
  int64_t splits = std::max<int64_t>(1, sm_count() / nchunks);
  splits = std::min<int64_t>(splits, n_tiles);
  auto k = gmm_tc_kernel<STATS>;
  ODIN_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES));
  k<<<dim3(nchunks, (unsigned)splits), tc::THREADS, tc::SMEM_BYTES, st>>>(a);
  ODIN_LAUNCH_CHECK(STATS ? "gmm_tc_kernel<stats>" : "gmm_tc_kernel<lse>");
  return ODIN_OK;
}

int gmm_lse_tc(odin_gmm* g, const float* X, const uint8_t* sad, int64_t N, float* lse, double* stats,
               cudaStream_t st) {
  if (N <= 0) return ODIN_OK;
  const int nchunks = g->Mpad / tc::CM;
  const int64_t sub = std::min<int64_t>(N, tc_sub_batch());
  const int64_t stride = ceil_div<int64_t>(sub, tc::NF) * tc::NF;
  const int64_t need = stride * nchunks;
  if (need > g->part_cap) {
    if (g->d_part) ODIN_CUDA_CHECK(cudaFree(g->d_part));
    g->d_part = nullptr;
    g->part_cap = 0;
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_part, need * sizeof(float2)));
    g->part_cap = need;
  }
  double* statL = stats ? stats + (stats_size(g->D, g->M) - 2) : nullptr;
  for (int64_t s0 = 0; s0 < N; s0 += sub) {
    const int64_t n = std::min<int64_t>(sub, N - s0);
    TcArgs a{};
    a.X = X + s0 * g->D; a.sad = nullptr; a.N = n; a.D = g->D; a.M = g->M;
    a.Whi = g->d_Whi; a.Wlo = g->d_Wlo; a.part = reinterpret_cast<float2*>(g->d_part); a.part_stride = stride;
    a.flush_tiles = 1 << 30;
    int rc = tc_launch<false>(g, a, st);
    if (rc) return rc;
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
  a.Whi = g->d_Whi; a.Wlo = g->d_Wlo; a.lse2 = lse; a.stats = stats; a.want_second = want_second;
  a.flush_tiles = tc_flush_tiles();
  return tc_launch<true>(g, a, st);
}

}  // namespace odin
