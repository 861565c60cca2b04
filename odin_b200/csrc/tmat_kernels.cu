// Total-variability model (i-vector extractor) on sm_100a, fp64 throughout.
//
// Reference arithmetic (trungnt13/odin-ai, odin/ml/gmm_tmat.py):
//   cached statistics        :1578-1589   T_invS = T / (Sigma + EPS); T_invS_Tt[m] = tril(T2_m T2_m^T),
//                                         T2 = T / (sqrt(Sigma) + EPS)
//   E-step                   :1694-1725   L1 = Z T_invS_Tt, B1 = F T_invS^T, per file: L = I + sym(L1),
//                                         Cxx = L^-1, Ex = Cxx B, llk, Exx = tril(Cxx + Ex Ex^T);
//                                         RU = Ex^T F, LU = Z^T Exx
//   M-step                   :1818-1865   per mixture solve(sym(LU_m), RU_m); minimum-divergence
//                                         T <- chol_upper(sym(sum LU / nframes)) T; orthogonalise T <- diag(s) V^T
//   i-vector                 :1898-1942
//
// Layout: Tm / T_invS / RU [tv, M*D] row-major (column m*D + d), T_invS_Tt / LU [M, t2] with
// t2 = tv (tv + 1) / 2 in np.tril_indices order (index a (a + 1) / 2 + b for a >= b), Z [n, M], F [n, M*D].
//
// Kernels
//   tmat_refresh_kernel   one CTA per mixture (T2 block in smem)
//   tmat_dgemm_kernel     fp64 GEMM with generic strides on the tensor instruction (mma.sync m8n8k4 f64 = DMMA), 128x128x16
//                         tiles (64-wide for skinny outputs), cp.async 3-stage, split-K (all products);
//                         tmat_gemm_kernel: the register-blocked DFMA kernel it replaced (ODIN_TMAT_GEMM_DFMA=1)
//   tmat_file_kernel      one CTA per file: the tv x tv system lives in ONE padded smem square (one-barrier-per-column
//                         Cholesky on unscaled columns, chol_lower) --
//                         Cholesky factor in the lower triangle, its inverse written transposed into the
//                         upper triangle, Cxx = G^-T G^-1 back into the lower triangle -- so tv = 128 fits
//   tmat_solve_kernel     one CTA per mixture: Cholesky + forward / backward substitution, one thread per column
//   tmat_mindiv_kernel    sum of LU over mixtures / nframes -> upper Cholesky factor
//   tmat_jacobi_kernel    one-sided (Hestenes) Jacobi on the ROWS of T: rotating rows until they are mutually
//                         orthogonal yields U^T T = diag(s) V^T directly, without forming T T^T or U; one launch
//                         per round of a round-robin tournament (tv / 2 disjoint pairs per round, a 4-CTA
//                         cluster per pair with the dot products reduced through distributed shared memory)
#include <dlfcn.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <new>
#include <vector>

#include <cooperative_groups.h>

#include "common.cuh"
#include "tmat.cuh"

namespace cg = cooperative_groups;

namespace odin {

constexpr double TM_EPS = 1e-6;  // gmm_tmat.py:27

__device__ __forceinline__ int tril_idx(int a, int b) { return a * (a + 1) / 2 + b; }  // a >= b

// ---------------------------------------------------------------------------
// cached statistics
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tmat_refresh_kernel(const double* __restrict__ Tm, const double* __restrict__ Sigma,
                                                           int tv, int D, int64_t MD, int M, double* __restrict__ T_invS,
                                                           double* __restrict__ T_invS_Tt, double* __restrict__ gws) {
  extern __shared__ double dyn_sm[];   // T2 block [tv][D + 1] (or in the global workspace for large tv)
  const int tid = threadIdx.x;
  const int P = D + 1;
  double* sm = gws ? gws + (size_t)blockIdx.x * tv * P : dyn_sm;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
  __syncthreads();
  for (int i = tid; i < tv * D; i += 256) {
    const int r = i / D, d = i - r * D;
    const int64_t g = (int64_t)r * MD + (int64_t)m * D + d;
    const double t = Tm[g], s = Sigma[(int64_t)m * D + d];
    T_invS[g] = t / (s + TM_EPS);
    sm[r * P + d] = t / (sqrt(s) + TM_EPS);
  }
  __syncthreads();
  const int t2 = tv * (tv + 1) / 2;
  for (int i = tid; i < tv * tv; i += 256) {
    const int a = i / tv, b = i - a * tv;
    if (b > a) continue;
    double acc = 0.0;
    for (int d = 0; d < D; ++d) acc = fma(sm[a * P + d], sm[b * P + d], acc);
    T_invS_Tt[(int64_t)m * t2 + tril_idx(a, b)] = acc;
  }
  }  // mixture loop
}

// ---------------------------------------------------------------------------
// C[M, N] = beta C + sum_k A(i, k) B(k, j);  A(i, k) = A[i sai + k sak],  B(k, j) = B[k sbk + j sbj]
// ---------------------------------------------------------------------------
constexpr int GBM = 128, GBN = 64, GK = 16;   // CTA tile 128 x 64, 256 threads, 8 x 4 outputs per thread

// gridDim.z > 1: split-K -- slice z of the K range writes its partial product (beta ignored) to C + z * M * ldc of a
// workspace; tmat_splitk_reduce_kernel adds the slices in a fixed order.
// Per k step a thread reads 8 + 4 operands (16-byte shared loads, the A operands of a warp are two broadcast rows)
// for 32 DFMAs: the first version's 4 x 4 tile spent as many shared-memory wavefronts as fp64 issue cycles.
__global__ void __launch_bounds__(256, 2) tmat_gemm_kernel(int M, int N, int K, const double* __restrict__ A, int64_t sai,
                                                           int64_t sak, const double* __restrict__ B, int64_t sbk,
                                                           int64_t sbj, double* __restrict__ C, int64_t ldc, double beta) {
  __shared__ __align__(16) double As[GK][GBM + 2];
  __shared__ __align__(16) double Bs[GK][GBN + 2];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;   // thread: rows 8 ty .. 8 ty + 7, cols 4 tx .. 4 tx + 3
  const int i0 = blockIdx.y * GBM, j0 = blockIdx.x * GBN;
  double acc[8][4];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  int kbeg = 0, kend = K;
  if (gridDim.z > 1) {
    const int per = ((K + (int)gridDim.z - 1) / (int)gridDim.z + GK - 1) / GK * GK;
    kbeg = (int)blockIdx.z * per;
    kend = min(K, kbeg + per);
    C += (int64_t)blockIdx.z * M * ldc;
    beta = 0.0;
  }
  for (int k0 = kbeg; k0 < kend; k0 += GK) {
    // tile loads: consecutive threads walk the unit-stride dimension of each operand
#pragma unroll
    for (int e = 0; e < (GBM * GK) / 256; ++e) {
      const int idx = tid + e * 256;
      int ii, kk;
      if (sak == 1) { kk = idx & (GK - 1); ii = idx >> 4; } else { ii = idx & (GBM - 1); kk = idx >> 7; }
      const int gi = i0 + ii, gk = k0 + kk;
      As[kk][ii] = (gi < M && gk < kend) ? A[(int64_t)gi * sai + (int64_t)gk * sak] : 0.0;
    }
#pragma unroll
    for (int e = 0; e < (GBN * GK) / 256; ++e) {
      const int idx = tid + e * 256;
      int jj, kb;
      if (sbk == 1) { kb = idx & (GK - 1); jj = idx >> 4; } else { jj = idx & (GBN - 1); kb = idx >> 6; }
      const int gj = j0 + jj, gkb = k0 + kb;
      Bs[kb][jj] = (gj < N && gkb < kend) ? B[(int64_t)gkb * sbk + (int64_t)gj * sbj] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      double av[8], bv[4];
      const double2* ap = reinterpret_cast<const double2*>(&As[kk][8 * ty]);
      const double2* bp = reinterpret_cast<const double2*>(&Bs[kk][4 * tx]);
#pragma unroll
      for (int a = 0; a < 4; ++a) { const double2 t = ap[a]; av[2 * a] = t.x; av[2 * a + 1] = t.y; }
#pragma unroll
      for (int b = 0; b < 2; ++b) { const double2 t = bp[b]; bv[2 * b] = t.x; bv[2 * b + 1] = t.y; }
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int gi = i0 + 8 * ty + a;
    if (gi >= M) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gj = j0 + 4 * tx + b;
      if (gj >= N) continue;
      double* c = C + (int64_t)gi * ldc + gj;
      *c = (beta == 0.0) ? acc[a][b] : fma(beta, *c, acc[a][b]);
    }
  }
}

// ---------------------------------------------------------------------------
// The same product on the fp64 TENSOR instruction (mma.sync m8n8k4 f64, SASS DMMA).  tools/f64_bench.cu on the box: DMMA
// sustains 37.0 TFLOP/s at 0.25 warp-instructions per clock and SM, DFMA 31.7 TFLOP/s at 1.70 -- the register-blocked
// kernel above spends its issue slots on the multiply-adds themselves (8.9 TFLOP/s over the four products of the
// E-step), here they are free for the operand traffic.  CTA tile BM x BN x 16, eight warps of (BM / WM) x (BN / WN),
// three cp.async stages (16-byte copies when the operand's unit-stride dimension allows it, 8-byte otherwise, zero fill
// at the edges).  An operand tile is stored the way it arrives -- unit stride along k: [rows][16 + 4], unit stride along
// the tile dimension: [16][rows + 4]; both row strides are 4 (mod 16) doubles, which makes the fragment reads of a half
// warp (4 k x 4 rows) fall on 16 distinct 8-byte banks.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_zfill(void* smem, const void* gmem, int bytes, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem);
  if (bytes == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(d), "l"(gmem), "r"(src_bytes) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(d), "l"(gmem), "r"(src_bytes) : "memory");
}

// one operand tile (ROWS x 16) of stage memory `dst`: element (r, k) of the tile is X[(r0 + r) sr + (k0 + k) sk]
template <int ROWS>
__device__ __forceinline__ void dgemm_load_tile(double* dst, const double* __restrict__ X, int64_t sr, int64_t sk, int r0, int nrows,
                                                int k0, int kend, bool vec, int tid) {
  constexpr int BK = 16;
  if (sk == 1) {            // unit stride along k: [ROWS][20]
    if (vec) {
      for (int idx = tid; idx < ROWS * (BK / 2); idx += 256) {
        const int r = idx >> 3, k = (idx & 7) * 2;
        const int left = (r0 + r < nrows) ? min(2, max(0, kend - (k0 + k))) : 0;
        const double* src = left > 0 ? X + (int64_t)(r0 + r) * sr + (k0 + k) : X;
        cp_async_zfill(dst + r * 20 + k, src, 16, 8 * left);
      }
    } else {
      for (int idx = tid; idx < ROWS * BK; idx += 256) {
        const int r = idx >> 4, k = idx & 15;
        const bool in = r0 + r < nrows && k0 + k < kend;
        cp_async_zfill(dst + r * 20 + k, in ? X + (int64_t)(r0 + r) * sr + (k0 + k) : X, 8, in ? 8 : 0);
      }
    }
  } else {                  // unit stride along the tile dimension: [16][ROWS + 4]
    if (vec) {
      for (int idx = tid; idx < BK * (ROWS / 2); idx += 256) {
        const int k = idx / (ROWS / 2), r = (idx - k * (ROWS / 2)) * 2;
        const int left = (k0 + k < kend) ? min(2, max(0, nrows - (r0 + r))) : 0;
        const double* src = left > 0 ? X + (int64_t)(k0 + k) * sk + (r0 + r) : X;
        cp_async_zfill(dst + k * (ROWS + 4) + r, src, 16, 8 * left);
      }
    } else {
      for (int idx = tid; idx < BK * ROWS; idx += 256) {
        const int k = idx / ROWS, r = idx - k * ROWS;
        const bool in = r0 + r < nrows && k0 + k < kend;
        cp_async_zfill(dst + k * (ROWS + 4) + r, in ? X + (int64_t)(k0 + k) * sk + (int64_t)(r0 + r) * sr : X, 8, in ? 8 : 0);
      }
    }
  }
}

template <int BM, int BN, int WM, int WN>
__global__ void __launch_bounds__(256, 1) tmat_dgemm_kernel(int M, int N, int K, const double* __restrict__ A, int64_t sai,
                                                            int64_t sak, const double* __restrict__ B, int64_t sbk, int64_t sbj,
                                                            double* __restrict__ C, int64_t ldc, double beta) {
  static_assert(WM * WN == 8, "eight warps");
  constexpr int BK = 16, STAGES = 3;
  constexpr int TM = BM / WM, TN = BN / WN, FM = TM / 8, FN = TN / 8;
  constexpr int SZA = BM * 20, SZB = BN * 20;   // doubles per stage (>= 16 * (rows + 4) for rows >= 16)
  extern __shared__ __align__(16) double dsm[];
  double* sA = dsm;
  double* sB = dsm + STAGES * SZA;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm0 = (warp / WN) * TM, wn0 = (warp % WN) * TN;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  int kbeg = 0, kend = K;
  if (gridDim.z > 1) {
    const int per = ((K + (int)gridDim.z - 1) / (int)gridDim.z + BK - 1) / BK * BK;
    kbeg = (int)blockIdx.z * per;
    kend = min(K, kbeg + per);
    C += (int64_t)blockIdx.z * M * ldc;
    beta = 0.0;
  }
  // 16-byte copies need even element offsets: the operand's non-unit stride even, its base 16-byte aligned (tile and
  // k origins are multiples of 16)
  const bool vecA = (sak == 1 || sai == 1) && (((sak == 1 ? sai : sak) & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool vecB = (sbk == 1 || sbj == 1) && (((sbk == 1 ? sbj : sbk) & 1) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  const int a_m = sak == 1 ? 20 : 1, a_k = sak == 1 ? 1 : BM + 4;   // smem strides of element (m, k)
  const int b_n = sbk == 1 ? 20 : 1, b_k = sbk == 1 ? 1 : BN + 4;
  const int nk = (kend - kbeg + BK - 1) / BK;
  auto load_stage = [&](int kt) {
    const int st = kt % STAGES, k0 = kbeg + kt * BK;
    dgemm_load_tile<BM>(sA + st * SZA, A, sai, sak, i0, M, k0, kend, vecA, tid);
    dgemm_load_tile<BN>(sB + st * SZB, B, sbj, sbk, j0, N, k0, kend, vecB, tid);
  };
  double acc[FM][FN][2];
#pragma unroll
  for (int a = 0; a < FM; ++a)
#pragma unroll
    for (int b = 0; b < FN; ++b) acc[a][b][0] = acc[a][b][1] = 0.0;
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) load_stage(s);
    asm volatile("cp.async.commit_group;" ::: "memory");
  }
  const int fr = lane >> 2, fk = lane & 3;   // fragment row (m or n) and k of this lane
  for (int kt = 0; kt < nk; ++kt) {
    asm volatile("cp.async.wait_group %0;" :: "n"(STAGES - 2) : "memory");
    __syncthreads();   // stage kt has landed for everybody; everybody is done with stage kt - 1, which is refilled now
    if (kt + STAGES - 1 < nk) load_stage(kt + STAGES - 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const double* pa = sA + (kt % STAGES) * SZA + (wm0 + fr) * a_m + fk * a_k;
    const double* pb = sB + (kt % STAGES) * SZB + (wn0 + fr) * b_n + fk * b_k;
#pragma unroll
    for (int ks = 0; ks < BK / 4; ++ks) {
      double av[FM], bv[FN];
#pragma unroll
      for (int a = 0; a < FM; ++a) av[a] = pa[8 * a * a_m + 4 * ks * a_k];
#pragma unroll
      for (int b = 0; b < FN; ++b) bv[b] = pb[8 * b * b_n + 4 * ks * b_k];
#pragma unroll
      for (int a = 0; a < FM; ++a)
#pragma unroll
        for (int b = 0; b < FN; ++b)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(acc[a][b][0]), "+d"(acc[a][b][1]) : "d"(av[a]), "d"(bv[b]));
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // C fragment: row lane / 4, columns 2 (lane % 4), + 1
#pragma unroll
  for (int a = 0; a < FM; ++a) {
    const int gi = i0 + wm0 + 8 * a + fr;
    if (gi >= M) continue;
#pragma unroll
    for (int b = 0; b < FN; ++b) {
      const int gj = j0 + wn0 + 8 * b + 2 * fk;
      double* c = C + (int64_t)gi * ldc + gj;
      if (gj < N) c[0] = (beta == 0.0) ? acc[a][b][0] : fma(beta, c[0], acc[a][b][0]);
      if (gj + 1 < N) c[1] = (beta == 0.0) ? acc[a][b][1] : fma(beta, c[1], acc[a][b][1]);
    }
  }
}

template <int BM, int BN, int WM, int WN>
static int launch_dgemm(dim3 grid, int M, int N, int K, const double* A, int64_t sai, int64_t sak, const double* B, int64_t sbk,
                        int64_t sbj, double* C, int64_t ldc, double beta, cudaStream_t st) {
  constexpr size_t smem = sizeof(double) * 3 * (BM * 20 + BN * 20);
  auto k = tmat_dgemm_kernel<BM, BN, WM, WN>;
  static bool attr_set = false;   // (per instantiation)
  if (!attr_set) {
    ODIN_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  k<<<grid, 256, smem, st>>>(M, N, K, A, sai, sak, B, sbk, sbj, C, ldc, beta);
  ODIN_LAUNCH_CHECK("tmat_dgemm_kernel");
  return ODIN_OK;
}

__global__ void __launch_bounds__(256) tmat_splitk_reduce_kernel(const double* __restrict__ ws, int splits, int64_t MN,
                                                                 int N, double* __restrict__ C, int64_t ldc, double beta) {
  for (int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x; e < MN; e += (int64_t)gridDim.x * 256) {
    double acc = 0.0;
    for (int z = 0; z < splits; ++z) acc += ws[(int64_t)z * MN + e];
    double* c = C + (e / N) * ldc + (e % N);
    *c = (beta == 0.0) ? acc : fma(beta, *c, acc);
  }
}

// ws / ws_cap: optional split-K workspace (doubles).  Split-K is used when the output has too few tiles to fill
// the GPU and K is long (B1 = F T_invS^T: 47 tiles, K = M D = 30 720 at the config-5 scale).
static int gemm(int M, int N, int K, const double* A, int64_t sai, int64_t sak, const double* B, int64_t sbk, int64_t sbj,
                double* C, int64_t ldc, double beta, cudaStream_t st, double* ws = nullptr, int64_t ws_cap = 0) {
  if (M <= 0 || N <= 0) return ODIN_OK;
  static const bool ffma = [] { const char* e = getenv("ODIN_TMAT_GEMM_DFMA"); return e && e[0] == '1'; }();   // A/B runs
  // tile shape by the output's aspect: 128 x 128, or a 64-wide side for the skinny products (tv = 64 rows / columns)
  const int bm = ffma ? GBM : (M <= 64 ? 64 : 128), bn = ffma ? GBN : (N <= 64 ? 64 : 128);
  dim3 grid((unsigned)ceil_div(N, bn), (unsigned)ceil_div(M, bm));
  const int64_t tiles = (int64_t)grid.x * grid.y;
  int splits = 1;
  if (ws != nullptr && K >= 1024) {
    if (ffma) {
      if (tiles < 2 * sm_count()) splits = (int)std::min<int64_t>(std::min<int64_t>(4 * sm_count() / tiles, K / 256), 32);
    } else if (tiles < 4 * sm_count()) {
      // one CTA per SM: the smallest split-K factor that fills >= 92 % of its last wave (slices of >= 512 columns of K)
      const int sms = sm_count();
      double best = 0.0;
      for (int sp = 1; sp <= 32 && K / sp >= 512; ++sp) {
        const int64_t ctas = tiles * sp;
        const double eff = (double)ctas / (double)(ceil_div<int64_t>(ctas, sms) * sms);
        if (eff > best + 1e-9) { best = eff; splits = sp; }
        if (eff >= 0.92) break;
      }
    }
    while (splits > 1 && (int64_t)splits * M * N > ws_cap) --splits;
  }
  grid.z = (unsigned)std::max(splits, 1);
  double* out = splits > 1 ? ws : C;
  const int64_t ldo = splits > 1 ? N : ldc;
  const double b0 = splits > 1 ? 0.0 : beta;
  if (ffma) {
    tmat_gemm_kernel<<<grid, 256, 0, st>>>(M, N, K, A, sai, sak, B, sbk, sbj, out, ldo, b0);
    ODIN_LAUNCH_CHECK("tmat_gemm_kernel");
  } else {
    int rc;
    if (bm == 128 && bn == 128) rc = launch_dgemm<128, 128, 4, 2>(grid, M, N, K, A, sai, sak, B, sbk, sbj, out, ldo, b0, st);
    else if (bm == 128) rc = launch_dgemm<128, 64, 4, 2>(grid, M, N, K, A, sai, sak, B, sbk, sbj, out, ldo, b0, st);
    else if (bn == 128) rc = launch_dgemm<64, 128, 2, 4>(grid, M, N, K, A, sai, sak, B, sbk, sbj, out, ldo, b0, st);
    else rc = launch_dgemm<64, 64, 2, 4>(grid, M, N, K, A, sai, sak, B, sbk, sbj, out, ldo, b0, st);
    if (rc) return rc;
  }
  if (splits <= 1) return ODIN_OK;
  const int64_t MN = (int64_t)M * N;
  tmat_splitk_reduce_kernel<<<(unsigned)std::min<int64_t>(ceil_div<int64_t>(MN, 256), sm_count() * 8), 256, 0, st>>>(
      ws, splits, MN, N, C, ldc, beta);
  ODIN_LAUNCH_CHECK("tmat_splitk_reduce_kernel");
  return ODIN_OK;
}

// ---------------------------------------------------------------------------
// in-place Cholesky of the lower triangle of S [n][P] (A = G G^T), P >= n + 1; returns false on a non-positive pivot.
//
// Right-looking on UNSCALED columns: step k subtracts (S[i][k] / S[k][k]) S[j][k] from the trailing triangle, so a step
// is ONE barrier (the pivot of step k + 1 is final when step k's updates are) with neither a square root nor a column
// scaling on the critical path; the columns are divided by the square roots of their pivots in one pass at the end
// (pivots parked in the pad column S[k][n]).  Threads form a 16 x (nt / 16) grid over (j, i): no index divisions.
// The first version (scale column, barrier, update through e -> (e / r, e % r), three barriers per column) took
// ~130 us of the ~200 us a 64 x 64 file spent in tmat_file_kernel.
// ---------------------------------------------------------------------------
__device__ bool chol_lower(double* S, int n, int P, int* s_flag) {
  const int tid = threadIdx.x, nt = blockDim.x;
  const int tx = tid & 15, ty = tid >> 4, ny = nt >> 4;
  for (int k = 0; k < n; ++k) {
    __syncthreads();
    const double akk = S[k * P + k];
    if (!(akk > 0.0)) {   // uniform: every thread reads the same value
      if (tid == 0) *s_flag = 1;
      return false;
    }
    const double inv = 1.0 / akk;
    for (int i = k + 1 + ty; i < n; i += ny) {
      const double lik = S[i * P + k] * inv;
      for (int j = k + 1 + tx; j <= i; j += 16) S[i * P + j] = fma(-lik, S[j * P + k], S[i * P + j]);
    }
  }
  __syncthreads();
  for (int k = tid; k < n; k += nt) S[k * P + n] = sqrt(S[k * P + k]);
  __syncthreads();
  for (int k = tx; k < n; k += 16) {
    const double g = S[k * P + n], rg = 1.0 / g;
    for (int i = k + ty; i < n; i += ny) S[i * P + k] = (i == k) ? g : S[i * P + k] * rg;
  }
  __syncthreads();
  return true;
}

// ---------------------------------------------------------------------------
// per-file posterior of the latent factor
// ---------------------------------------------------------------------------
struct FileArgs {
  int tv;
  int64_t n;
  double* L1;        // [n, t2] in: Z T_invS_Tt; out: Exx (when want_exx)
  const double* B1;  // [n, tv]
  double* Ex;        // [n, tv]
  double* llk;       // [n] nullable
  int want_exx;
  int* flag;
  double* gws;       // nullable: per-CTA squares in global memory (large tv)
};

__global__ void __launch_bounds__(256) tmat_file_kernel(FileArgs a) {
  extern __shared__ double dyn_sm[];
  const int n = a.tv, P = n + 1, tid = threadIdx.x;
  // the square lives in shared memory up to tv ~ 160 and in a per-CTA slab of global memory (L2) beyond
  double* S = a.gws ? a.gws + (size_t)blockIdx.x * ((size_t)n * P + 3 * n) : dyn_sm;   // [n][P]
  double* dg = S + n * P;      // [n] diagonal of the Cholesky factor, later diagonal of Cxx
  double* Bv = dg + n;         // [n]
  double* Ev = Bv + n;         // [n]
  __shared__ int s_flag;
  __shared__ double s_red[8];
  const int t2 = n * (n + 1) / 2;
  for (int64_t f = blockIdx.x; f < a.n; f += gridDim.x) {
    __syncthreads();
    if (tid == 0) s_flag = 0;
    double* row = a.L1 + f * t2;
    for (int e = tid; e < n * n; e += 256) {
      const int i = e / n, j = e - i * n;
      if (j > i) continue;
      S[i * P + j] = row[tril_idx(i, j)] + (i == j ? 1.0 : 0.0);   // L = I + sym(L1)
    }
    for (int i = tid; i < n; i += 256) Bv[i] = a.B1[f * n + i];
    if (!chol_lower(S, n, P, &s_flag)) {
      if (tid == 0) atomicExch(a.flag, 1);
      continue;
    }
    // G^-1 by forward substitution, one thread per column j; entry (i, j) is stored TRANSPOSED at S[j][i]
    // (strict upper triangle), the diagonal of G moves to dg[] and 1 / G[j][j] takes its place
    for (int i = tid; i < n; i += 256) dg[i] = S[i * P + i];
    __syncthreads();
    {
      // four lanes per column j: each forms a quarter of the dot product, the quarters meet by shuffle (fixed order),
      // lane 0 of the group writes; __syncwarp orders the group's write before its next reads
      const int q = tid & 3;
      const unsigned gmask = 0xFu << (threadIdx.x & 28);
      for (int j = tid >> 2; j < n; j += 64) {
        if (q == 0) S[j * P + j] = 1.0 / dg[j];
        __syncwarp(gmask);
        for (int i = j + 1; i < n; ++i) {
          double acc = 0.0;
          for (int k = j + q; k < i; k += 4) acc = fma(S[i * P + k], S[j * P + k], acc);   // G[i][k] * Ginv[k][j]
          acc += __shfl_xor_sync(gmask, acc, 1);
          acc += __shfl_xor_sync(gmask, acc, 2);
          if (q == 0) S[j * P + i] = -acc / dg[i];
          __syncwarp(gmask);
        }
      }
    }
    // careful: thread j reads G[i][k] for k in [j, i) -- strictly lower entries and never the diagonal slot of
    // another column (k < i), except k == j where S[j][j] holds Ginv[j][j]: that IS Ginv[k][j] for k == j, and the
    // G factor G[i][j] is read from S[i][j] (lower) -- distinct slots, so the loop above is hazard-free.
    __syncthreads();
    // Cxx = G^-T G^-1: Cxx[p][q] = sum_{i >= p} Ginv[i][p] Ginv[i][q] (p >= q) = row p . row q of the upper storage
    for (int e = tid; e < n * n; e += 256) {
      const int p = e / n, q = e - p * n;
      if (q > p) continue;
      double acc = 0.0;
      for (int i = p; i < n; ++i) acc = fma(S[p * P + i], S[q * P + i], acc);
      if (p == q) dg[p] = acc; else S[p * P + q] = acc;   // lower triangle is free again (G is consumed)
    }
    __syncthreads();
    // Ex = Cxx B
    for (int p = tid; p < n; p += 256) {
      double acc = 0.0;
      for (int q = 0; q < n; ++q) {
        const double c = (q == p) ? dg[p] : (q < p ? S[p * P + q] : S[q * P + p]);
        acc = fma(c, Bv[q], acc);
      }
      Ev[p] = acc;
      a.Ex[f * n + p] = acc;
    }
    __syncthreads();
    if (a.llk != nullptr) {   // -0.5 Ex^T (B - Ex) + Ex^T B   (gmm_tmat.py:1719)
      double part = 0.0;
      for (int p = tid; p < n; p += 256) part += -0.5 * Ev[p] * (Bv[p] - Ev[p]) + Ev[p] * Bv[p];
      part = warp_sum(part);
      if ((tid & 31) == 0) s_red[tid >> 5] = part;
      __syncthreads();
      if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_red[w];
        a.llk[f] = t;
      }
    }
    if (a.want_exx) {
      for (int e = tid; e < n * n; e += 256) {
        const int p = e / n, q = e - p * n;
        if (q > p) continue;
        const double c = (p == q) ? dg[p] : S[p * P + q];
        row[tril_idx(p, q)] = fma(Ev[p], Ev[q], c);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// M-step, per mixture: Tm[:, m-block] = sym(LU_m)^-1 RU[:, m-block]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tmat_solve_kernel(int tv, int D, int64_t MD, const double* __restrict__ LU,
                                                         const double* __restrict__ RU, double* __restrict__ Tm,
                                                         int* flag, int M, double* __restrict__ gws) {
  extern __shared__ double dyn_sm[];
  const int n = tv, P = n + 1, Q = D + 1, tid = threadIdx.x;
  double* S = gws ? gws + (size_t)blockIdx.x * ((size_t)n * P + (size_t)n * Q) : dyn_sm;   // [n][P]
  double* R = S + n * P;     // [n][Q]
  __shared__ int s_flag;
  const int t2 = n * (n + 1) / 2;
  for (int m = blockIdx.x; m < M; m += gridDim.x) {
  __syncthreads();
  if (tid == 0) s_flag = 0;
  for (int e = tid; e < n * n; e += 256) {
    const int i = e / n, j = e - i * n;
    if (j > i) continue;
    S[i * P + j] = LU[(int64_t)m * t2 + tril_idx(i, j)];
  }
  for (int e = tid; e < n * D; e += 256) {
    const int i = e / D, d = e - i * D;
    R[i * Q + d] = RU[(int64_t)i * MD + (int64_t)m * D + d];
  }
  if (!chol_lower(S, n, P, &s_flag)) {
    if (tid == 0) atomicExch(flag, 2);
    continue;
  }
  {
    // four lanes per right-hand side: quarter dot products joined by shuffle in a fixed order (one thread per column
    // left 60 of 256 threads walking 2 n^2 dependent multiply-adds)
    const int q = tid & 3;
    const unsigned gmask = 0xFu << (threadIdx.x & 28);
    for (int d = tid >> 2; d < D; d += 64) {
      for (int i = 0; i < n; ++i) {            // G y = r
        double acc = 0.0;
        for (int k = q; k < i; k += 4) acc = fma(S[i * P + k], R[k * Q + d], acc);
        acc += __shfl_xor_sync(gmask, acc, 1);
        acc += __shfl_xor_sync(gmask, acc, 2);
        if (q == 0) R[i * Q + d] = (R[i * Q + d] - acc) / S[i * P + i];
        __syncwarp(gmask);
      }
      for (int i = n - 1; i >= 0; --i) {       // G^T x = y
        double acc = 0.0;
        for (int k = i + 1 + q; k < n; k += 4) acc = fma(S[k * P + i], R[k * Q + d], acc);
        acc += __shfl_xor_sync(gmask, acc, 1);
        acc += __shfl_xor_sync(gmask, acc, 2);
        if (q == 0) R[i * Q + d] = (R[i * Q + d] - acc) / S[i * P + i];
        __syncwarp(gmask);
      }
    }
  }
  __syncthreads();
  for (int e = tid; e < n * D; e += 256) {
    const int i = e / D, d = e - i * D;
    Tm[(int64_t)i * MD + (int64_t)m * D + d] = R[i * Q + d];
  }
  }  // mixture loop
}

// minimum-divergence factor: U upper with sym(sum_m LU_m / nframes) = U^T U (scipy.linalg.cholesky default)
// LU.sum(0): one thread per column, mixtures in ascending order (a single CTA doing this inside the min-div kernel
// took 0.95 ms of a 9 ms EM iteration)
__global__ void __launch_bounds__(256) tmat_colsum_kernel(const double* __restrict__ LU, int nmix, int t2,
                                                          double* __restrict__ out) {
  const int j = blockIdx.x * 256 + threadIdx.x;
  if (j >= t2) return;
  double acc = 0.0;
  for (int m = 0; m < nmix; ++m) acc += LU[(int64_t)m * t2 + j];
  out[j] = acc;
}

__global__ void __launch_bounds__(1024) tmat_mindiv_kernel(int tv, int nmix, const double* __restrict__ LU,
                                                          const double* __restrict__ nframes, double* __restrict__ U,
                                                          int* flag, double* __restrict__ gws) {
  extern __shared__ double dyn_sm[];
  const int n = tv, P = n + 1, tid = threadIdx.x;
  double* S = gws ? gws : dyn_sm;
  __shared__ int s_flag;
  const int t2 = n * (n + 1) / 2;
  const double nf = *nframes;
  if (tid == 0) s_flag = 0;
  for (int e = tid; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    if (j > i) continue;
    S[i * P + j] = LU[tril_idx(i, j)] / nf;   // LU = LU.sum(0) here (tmat_colsum_kernel)
  }
  if (!chol_lower(S, n, P, &s_flag)) {
    if (tid == 0) atomicExch(flag, 3);
    return;
  }
  for (int e = tid; e < n * n; e += blockDim.x) {
    const int i = e / n, j = e - i * n;
    U[i * n + j] = (j >= i) ? S[j * P + i] : 0.0;   // U = G^T
  }
}

// ---------------------------------------------------------------------------
// one-sided Jacobi on the rows of W [tv, MD]: round `r` of a round-robin tournament over `np` players
// ---------------------------------------------------------------------------
constexpr int JT = 1024;   // threads per CTA: the three dot products are latency-bound, so use the whole SM
constexpr int JC = 4;      // CTAs (one cluster) per row pair: each takes a quarter of the columns

// One CLUSTER of JC CTAs per row pair: every CTA forms the three partial dot products over its slice of the
// columns, the partials meet through distributed shared memory (each CTA reads the JC slots in rank order, so
// all CTAs of the cluster hold bit-identical sums and take the same decision), then every CTA rotates its slice.
__global__ void __launch_bounds__(JT) tmat_jacobi_kernel(double* __restrict__ W, int tv, int64_t MD, int np, int r,
                                                         int* __restrict__ n_rot) {
  cg::cluster_group cl = cg::this_cluster();
  const int nc = (int)cl.num_blocks(), rank = (int)cl.block_rank();
  const int tid = threadIdx.x, i = blockIdx.x / nc;   // pair index 0 .. np/2 - 1
  int p, q;
  if (i == 0) { p = np - 1; q = r; }
  else { p = (r + i) % (np - 1); q = (r - i + (np - 1)) % (np - 1); }
  const bool bye = (p >= tv || q >= tv);   // the bye of an odd tournament (uniform over the cluster)
  if (p > q) { const int t = p; p = q; q = t; }
  const int64_t per = (MD + nc - 1) / nc;
  const int64_t j0 = per * rank, j1 = min(MD, j0 + per);
  double* wp = W + (int64_t)p * MD;
  double* wq = W + (int64_t)q * MD;
  double al = 0.0, be = 0.0, ga = 0.0;
  if (!bye)
    for (int64_t j = j0 + tid; j < j1; j += JT) {
      const double x = wp[j], y = wq[j];
      al = fma(x, x, al); be = fma(y, y, be); ga = fma(x, y, ga);
    }
  __shared__ double red[3][JT / 32];
  __shared__ double part[3];
  al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
  if ((tid & 31) == 0) { red[0][tid >> 5] = al; red[1][tid >> 5] = be; red[2][tid >> 5] = ga; }
  __syncthreads();
  if (tid < 3) {
    double t = 0.0;
    for (int w = 0; w < JT / 32; ++w) t += red[tid][w];
    part[tid] = t;
  }
  cl.sync();
  al = be = ga = 0.0;
  for (int c = 0; c < nc; ++c) {
    const double* rp = cl.map_shared_rank(part, c);
    al += rp[0]; be += rp[1]; ga += rp[2];
  }
  cl.sync();   // nobody leaves (or overwrites part) while a neighbour still reads it
  // already orthogonal: |cos| <= 1e-13.  (1e-15 sat at the rounding floor of the 30 720-long dot products -- ~1e-16 sqrt(MD)
  // -- so a sweep after the Gram pre-rotation still "rotated" pairs by angles of that size and a second sweep was needed
  // to see none: M-step 2.5 -> 2.2 ms at config-5 scale, 300 -> 246 ms at tv 400.  The tests ask for 1e-10.)
  if (bye || fabs(ga) <= 1e-13 * sqrt(al * be) || ga == 0.0) return;
  if (tid == 0 && rank == 0) atomicAdd(n_rot, 1);
  const double zeta = (be - al) / (2.0 * ga);
  const double t = (zeta >= 0.0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
  const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
  for (int64_t j = j0 + tid; j < j1; j += JT) {
    const double x = wp[j], y = wq[j];
    wp[j] = c * x - s * y;
    wq[j] = s * x + c * y;
  }
}

// Eigen-decomposition of the tv x tv Gram matrix G = T T^T = U S^2 U^T by two-sided cyclic Jacobi, one CTA, A and V
// in shared memory (tv <= TMAT_GRAM_MAX).  Used as a PRE-ROTATION: T <- U^T T makes the rows of T orthogonal up to
// eps * cond(T)^2, after which the one-sided sweeps on T itself (which do not square the condition number) converge in
// one or two passes instead of eight.  Vout [tv][tv]: column i = eigenvector of the i-th LARGEST eigenvalue.
constexpr int TMAT_GRAM_MAX = 96;

constexpr int ET = 1024;   // threads of the eigen-solver CTA (a round is ~n^2 / 2 independent updates between barriers)

__global__ void __launch_bounds__(ET) tmat_eig_kernel(const double* __restrict__ G, int n, double* __restrict__ Vout) {
  extern __shared__ double dyn_sm[];
  const int P = n + 1, tid = threadIdx.x;
  double* A = dyn_sm;            // [n][P]
  double* V = A + n * P;         // [n][P]
  double* cs = V + n * P;        // [n/2 + 1][2]
  __shared__ int s_rot;
  __shared__ short s_pq[TMAT_GRAM_MAX + 2];   // (p, q) of every pair of the current round, p < q, -1 = bye
  for (int e = tid; e < n * n; e += ET) {
    const int i = e / n, j = e - i * n;
    A[i * P + j] = G[e];
    V[i * P + j] = (i == j) ? 1.0 : 0.0;
  }
  const int np = n + (n & 1), half = np / 2;
  for (int sweep = 0; sweep < 40; ++sweep) {
    __syncthreads();
    if (tid == 0) s_rot = 0;
    for (int r = 0; r < np - 1; ++r) {
      __syncthreads();
      // rotation of every pair of this round
      if (tid < half) {
        const int i = tid;
        int p, q;
        if (i == 0) { p = np - 1; q = r; }
        else { p = (r + i) % (np - 1); q = (r - i + (np - 1)) % (np - 1); }
        double c = 1.0, s = 0.0;
        if (p < n && q < n) {
          if (p > q) { const int t = p; p = q; q = t; }
          const double apq = A[p * P + q], app = A[p * P + p], aqq = A[q * P + q];
          if (fabs(apq) > 1e-17 * sqrt(fabs(app * aqq)) && apq != 0.0) {
            const double tau = (aqq - app) / (2.0 * apq);
            const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + t * t);
            s = t * c;
            atomicAdd(&s_rot, 1);
          }
        } else {
          p = q = -1;
        }
        cs[2 * i] = c; cs[2 * i + 1] = s;
        s_pq[2 * i] = (short)p; s_pq[2 * i + 1] = (short)q;
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int e = tid; e < half * n; e += ET) {
        const int i = e / n, k = e - i * n;
        const int p = s_pq[2 * i], q = s_pq[2 * i + 1];
        const double c = cs[2 * i], s = cs[2 * i + 1];
        if (p < 0 || s == 0.0) continue;
        const double x = A[p * P + k], y = A[q * P + k];
        A[p * P + k] = c * x - s * y;
        A[q * P + k] = s * x + c * y;
      }
      __syncthreads();
      // columns: A <- A J, V <- V J
      for (int e = tid; e < half * n; e += ET) {
        const int i = e / n, k = e - i * n;
        const int p = s_pq[2 * i], q = s_pq[2 * i + 1];
        const double c = cs[2 * i], s = cs[2 * i + 1];
        if (p < 0 || s == 0.0) continue;
        double x = A[k * P + p], y = A[k * P + q];
        A[k * P + p] = c * x - s * y;
        A[k * P + q] = s * x + c * y;
        x = V[k * P + p]; y = V[k * P + q];
        V[k * P + p] = c * x - s * y;
        V[k * P + q] = s * x + c * y;
      }
    }
    __syncthreads();
    if (s_rot == 0) break;
  }
  __syncthreads();
  // eigenvalues descending (ties: lower index first) -> column order of Vout
  for (int q = tid; q < n; q += ET) {
    const double aq = A[q * P + q];
    int rank = 0;
    for (int p = 0; p < n; ++p) {
      const double ap = A[p * P + p];
      rank += (ap > aq) || (ap == aq && p < q);
    }
    for (int k = 0; k < n; ++k) Vout[k * n + rank] = V[k * P + q];
  }
}

// singular values: squared row norms, one CTA per row (fixed reduction order) ...
__global__ void __launch_bounds__(256) tmat_rownorm_kernel(const double* __restrict__ W, int64_t MD, double* __restrict__ nrm) {
  __shared__ double red[8];
  const int tid = threadIdx.x;
  const double* w = W + (int64_t)blockIdx.x * MD;
  double a = 0.0;
  for (int64_t j = tid; j < MD; j += 256) { const double x = w[j]; a = fma(x, x, a); }
  a = warp_sum(a);
  if ((tid & 31) == 0) red[tid >> 5] = a;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    nrm[blockIdx.x] = t;
  }
}

// ... -> descending order (stable)
__global__ void __launch_bounds__(256) tmat_order_kernel(const double* __restrict__ nrm, int tv, int* __restrict__ perm) {
  for (int q = threadIdx.x; q < tv; q += 256) {   // rank of row q = #rows with a larger norm (ties: lower index first)
    int rank = 0;
    for (int p = 0; p < tv; ++p) rank += (nrm[p] > nrm[q]) || (nrm[p] == nrm[q] && p < q);
    perm[rank] = q;
  }
}

__global__ void __launch_bounds__(256) tmat_gather_rows_kernel(const double* __restrict__ W, double* __restrict__ out,
                                                               const int* __restrict__ perm, int64_t MD) {
  const int r = blockIdx.y;
  const double* src = W + (int64_t)perm[r] * MD;
  for (int64_t j = blockIdx.x * 256 + threadIdx.x; j < MD; j += (int64_t)gridDim.x * 256) out[(int64_t)r * MD + j] = src[j];
}

// nframes and llk totals of a chunk, accumulated into the packed statistics.
// The reference takes nframes = ceil(sum Z) PER BATCH of its expectation() (gmm_tmat.py:1695, batches of
// 64 MiB / ((D M + M) itemsize) files, :1444-1449) and adds the batches up, so the ceil is applied per
// `rows_per_batch` files here as well: CTA b sums batch b (the integers it adds are exact in any order), the last CTA
// sums the likelihoods in a fixed order.  (One CTA for everything was 0.29 ms of a 5 ms E-step.)
__global__ void __launch_bounds__(1024) tmat_totals_kernel(const double* __restrict__ Z, int64_t n, int M,
                                                           int64_t rows_per_batch, const double* __restrict__ llk,
                                                           double* __restrict__ acc_llk, double* __restrict__ acc_nframes) {
  constexpr int NT = 1024;
  __shared__ double red[NT / 32];
  const int tid = threadIdx.x;
  auto block_sum = [&](double v) -> double {   // valid in thread 0
    v = warp_sum(v);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (tid == 0)
      for (int w = 0; w < NT / 32; ++w) t += red[w];
    return t;
  };
  if (blockIdx.x + 1 < gridDim.x) {
    const int64_t s = (int64_t)blockIdx.x * rows_per_batch;
    const int64_t e = min(n, s + rows_per_batch);
    const int64_t lo = s * M, hi = e * M;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;   // four independent chains per thread (the loads are the latency)
    int64_t i = lo + tid;
    for (; i + 3 * NT < hi; i += 4 * NT) { a0 += Z[i]; a1 += Z[i + NT]; a2 += Z[i + 2 * NT]; a3 += Z[i + 3 * NT]; }
    for (; i < hi; i += NT) a0 += Z[i];
    const double t = block_sum((a0 + a1) + (a2 + a3));
    if (tid == 0) atomicAdd(acc_nframes, ceil(t));
  } else {
    double b = 0.0;
    for (int64_t i = tid; i < n; i += NT) b += llk[i];
    b = block_sum(b);
    if (tid == 0) *acc_llk += b;
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static size_t square_smem(int tv, int extra_doubles) {
  return sizeof(double) * ((size_t)tv * (tv + 1) + extra_doubles);
}

constexpr size_t SMEM_LIMIT = 200 * 1024;   // beyond this a kernel's matrices move to a per-CTA slab of global memory

// per-CTA global slabs for the sizes whose systems do not fit shared memory (tv > ~150): `ctas` slabs of
// `doubles_per_cta`; the data stay L2-resident while a CTA works on them, but the factorisations are not blocked,
// so this route is much slower per FLOP than the shared-memory one
// pairs of rows still not orthogonal after the Gram pre-rotation: |G[p][q]| > 1e-13 sqrt(G[p][p] G[q][q]) (the criterion of
// tmat_jacobi_kernel), counted from ONE more Gram matrix -- when there is none, the verifying sweep (tv - 1 launches) is skipped
__global__ void __launch_bounds__(256) tmat_gramcheck_kernel(const double* __restrict__ G, int n, int* __restrict__ count) {
  const int e = blockIdx.x * 256 + threadIdx.x;
  if (e >= n * n) return;
  const int p = e / n, q = e - p * n;
  if (q >= p) return;
  const double g = G[e];
  if (fabs(g) > 1e-13 * sqrt(G[p * n + p] * G[q * n + q]) && g != 0.0) atomicAdd(count, 1);
}

// ---- symmetric eigen-solver for tv > TMAT_GRAM_MAX: cuSOLVER's Dsyevd, bound at run time ------------------------------
// The Gram pre-rotation needs the eigenvectors of ONE tv x tv matrix per M-step (tv 400-600 at the NIST-SRE recipe's
// scale).  That is a plain LAPACK call outside any hot loop, so beyond the size tmat_eig_kernel holds in shared memory it
// goes to the CUDA toolkit's library -- loaded with dlopen so that libodin_b200.so does not depend on it; when the
// library cannot be loaded the pre-rotation is skipped and the one-sided sweeps do all the work (slower, same result).
namespace {
struct Cusolver {
  bool tried = false;
  void* lib = nullptr;
  void* handle = nullptr;
  int (*create)(void**) = nullptr;
  int (*set_stream)(void*, cudaStream_t) = nullptr;
  int (*bufsize)(void*, int, int, int, const double*, int, const double*, int*) = nullptr;
  int (*syevd)(void*, int, int, int, double*, int, double*, double*, int, int*) = nullptr;
};
Cusolver g_cusolver;
std::mutex g_cusolver_mu;

bool cusolver_ready() {
  std::lock_guard<std::mutex> lock(g_cusolver_mu);
  Cusolver& c = g_cusolver;
  if (c.tried) return c.handle != nullptr;
  c.tried = true;
  const char* names[] = {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11",
                         "/usr/local/cuda/targets/x86_64-linux/lib/libcusolver.so.11"};
  for (const char* n : names)
    if ((c.lib = dlopen(n, RTLD_NOW | RTLD_LOCAL)) != nullptr) break;
  if (c.lib == nullptr) return false;
  c.create = reinterpret_cast<decltype(c.create)>(dlsym(c.lib, "cusolverDnCreate"));
  c.set_stream = reinterpret_cast<decltype(c.set_stream)>(dlsym(c.lib, "cusolverDnSetStream"));
  c.bufsize = reinterpret_cast<decltype(c.bufsize)>(dlsym(c.lib, "cusolverDnDsyevd_bufferSize"));
  c.syevd = reinterpret_cast<decltype(c.syevd)>(dlsym(c.lib, "cusolverDnDsyevd"));
  if (!c.create || !c.set_stream || !c.bufsize || !c.syevd) return false;
  if (c.create(&c.handle) != 0) c.handle = nullptr;
  return c.handle != nullptr;
}
}  // namespace

// Dsyevd leaves the eigenvectors in the columns of a column-major matrix, eigenvalues ascending; the pre-rotation wants
// Vout [tv][tv] row-major with column i = eigenvector of the i-th LARGEST eigenvalue.  A failed factorisation (info != 0)
// yields the identity: the one-sided sweeps then start from the unrotated T.
__global__ void tmat_eigperm_kernel(const double* __restrict__ A, const int* __restrict__ info, int n, double* __restrict__ Vout) {
  const bool ok = *info == 0;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * n; e += gridDim.x * blockDim.x) {
    const int r = e / n, i = e - r * n;
    Vout[e] = ok ? A[r + (size_t)(n - 1 - i) * n] : (r == i ? 1.0 : 0.0);
  }
}

static int reserve_gws(odin_tmat* t, size_t doubles_per_cta, int ctas) {
  const size_t need = doubles_per_cta * (size_t)ctas;
  if (need <= t->gws_cap) return ODIN_OK;
  cudaFree(t->d_gws);
  t->d_gws = nullptr; t->gws_cap = 0;
  cudaError_t e = cudaMalloc(&t->d_gws, sizeof(double) * need);
  if (e != cudaSuccess) return set_error(ODIN_ENOMEM, "T-matrix workspace (%zu MB): %s", need >> 17, cudaGetErrorString(e));
  t->gws_cap = need;
  return ODIN_OK;
}

int tmat_refresh(odin_tmat* t, cudaStream_t st) {
  const size_t per = (size_t)t->tv * (t->D + 1);
  const size_t smem = sizeof(double) * per;
  if (smem <= SMEM_LIMIT) {
    ODIN_CUDA_CHECK(cudaFuncSetAttribute(tmat_refresh_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tmat_refresh_kernel<<<t->M, 256, smem, st>>>(t->d_Tm, t->d_Sigma, t->tv, t->D, t->MD, t->M, t->d_TinvS, t->d_TinvSTt,
                                                 nullptr);
  } else {
    const int ctas = std::min(t->M, sm_count() * 4);
    int rc = reserve_gws(t, per, ctas);
    if (rc) return rc;
    tmat_refresh_kernel<<<ctas, 256, 0, st>>>(t->d_Tm, t->d_Sigma, t->tv, t->D, t->MD, t->M, t->d_TinvS, t->d_TinvSTt,
                                              t->d_gws);
  }
  ODIN_LAUNCH_CHECK("tmat_refresh_kernel");
  return ODIN_OK;
}

static int reserve_files(odin_tmat* t, int64_t n) {
  if (n <= t->cap_files) return ODIN_OK;
  cudaFree(t->d_L1); cudaFree(t->d_B1); cudaFree(t->d_Ex); cudaFree(t->d_llk); cudaFree(t->d_ws);
  t->d_L1 = t->d_B1 = t->d_Ex = t->d_llk = t->d_ws = nullptr;
  t->cap_files = 0;
  ODIN_CUDA_CHECK(cudaMalloc(&t->d_L1, sizeof(double) * (size_t)n * t->t2));
  ODIN_CUDA_CHECK(cudaMalloc(&t->d_B1, sizeof(double) * (size_t)n * t->tv));
  ODIN_CUDA_CHECK(cudaMalloc(&t->d_Ex, sizeof(double) * (size_t)n * t->tv));
  ODIN_CUDA_CHECK(cudaMalloc(&t->d_llk, sizeof(double) * (size_t)n));
  // split-K workspace: slices of B1 [n, tv] and of the accumulating products RU [tv, MD] / LU [M, t2] (at most 512 MiB)
  t->ws_cap = std::max<int64_t>((int64_t)16 * n * t->tv,
                                std::min<int64_t>((int64_t)4 * std::max<int64_t>((int64_t)t->tv * t->MD, (int64_t)t->M * t->t2),
                                                  (int64_t)64 << 20));
  ODIN_CUDA_CHECK(cudaMalloc(&t->d_ws, sizeof(double) * (size_t)t->ws_cap));
  t->cap_files = n;
  return ODIN_OK;
}

// posterior of a chunk of files: fills d_Ex (and d_L1 <- Exx, d_llk when training)
static int posterior_chunk(odin_tmat* t, const double* d_Z, const double* d_F, int64_t n, bool training, double* d_ex_out,
                           cudaStream_t st) {
  int rc;
  // L1 = Z T_invS_Tt  [n, t2]
  if ((rc = gemm((int)n, t->t2, t->M, d_Z, t->M, 1, t->d_TinvSTt, t->t2, 1, t->d_L1, t->t2, 0.0, st))) return rc;
  // B1 = F T_invS^T   [n, tv]
  if ((rc = gemm((int)n, t->tv, (int)t->MD, d_F, t->MD, 1, t->d_TinvS, 1, t->MD, t->d_B1, t->tv, 0.0, st, t->d_ws, t->ws_cap)))
    return rc;
  FileArgs a{};
  a.tv = t->tv; a.n = n; a.L1 = t->d_L1; a.B1 = t->d_B1; a.Ex = d_ex_out; a.llk = training ? t->d_llk : nullptr;
  a.want_exx = training ? 1 : 0; a.flag = t->d_flag;
  const size_t smem = square_smem(t->tv, 3 * t->tv);
  if (smem <= SMEM_LIMIT) {
    ODIN_CUDA_CHECK(cudaFuncSetAttribute(tmat_file_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (227 * 1024) / (smem + 1024)));
    const int64_t grid = std::min<int64_t>(n, (int64_t)sm_count() * per_sm);
    a.gws = nullptr;
    tmat_file_kernel<<<(unsigned)grid, 256, smem, st>>>(a);
  } else {
    const int64_t grid = std::min<int64_t>(n, (int64_t)sm_count() * 4);
    if ((rc = reserve_gws(t, smem / sizeof(double), (int)grid))) return rc;
    a.gws = t->d_gws;
    tmat_file_kernel<<<(unsigned)grid, 256, 0, st>>>(a);
  }
  ODIN_LAUNCH_CHECK("tmat_file_kernel");
  return ODIN_OK;
}

int tmat_estep(odin_tmat* t, const double* d_Z, const double* d_F, int64_t n_files, double* d_acc, cudaStream_t st) {
  double* d_LU = d_acc;
  double* d_RU = d_LU + (size_t)t->M * t->t2;
  double* d_llk = d_RU + (size_t)t->tv * t->MD;
  double* d_nframes = d_llk + 1;
  // files per chunk: a multiple of the reference's batch (see tmat_totals_kernel) within ~256 MiB of Exx rows
  const int64_t ref_batch = std::max<int64_t>(1, (int64_t)(64u << 20) / ((t->MD + t->M) * (int64_t)sizeof(double)));
  const int64_t cap = std::max<int64_t>(1, (int64_t)(256u << 20) / ((int64_t)sizeof(double) * t->t2));
  const int64_t chunk = std::max<int64_t>(1, cap / ref_batch) * ref_batch;
  int rc = reserve_files(t, std::min(chunk, n_files));
  if (rc) return rc;
  for (int64_t s = 0; s < n_files; s += chunk) {
    const int64_t n = std::min(chunk, n_files - s);
    const double* Z = d_Z + s * t->M;
    const double* F = d_F + s * t->MD;
    if ((rc = posterior_chunk(t, Z, F, n, true, t->d_Ex, st))) return rc;
    // RU += Ex^T F  [tv, MD];  LU += Z^T Exx  [M, t2]
    if ((rc = gemm(t->tv, (int)t->MD, (int)n, t->d_Ex, 1, t->tv, F, t->MD, 1, d_RU, t->MD, 1.0, st, t->d_ws, t->ws_cap))) return rc;
    if ((rc = gemm(t->M, t->t2, (int)n, Z, 1, t->M, t->d_L1, t->t2, 1, d_LU, t->t2, 1.0, st, t->d_ws, t->ws_cap))) return rc;
    tmat_totals_kernel<<<(unsigned)(ceil_div<int64_t>(n, ref_batch) + 1), 1024, 0, st>>>(Z, n, t->M, ref_batch, t->d_llk, d_llk,
                                                                                        d_nframes);
    ODIN_LAUNCH_CHECK("tmat_totals_kernel");
  }
  return ODIN_OK;
}

int tmat_ivector(odin_tmat* t, const double* d_Z, const double* d_F, int64_t n_files, double* d_out, cudaStream_t st) {
  const int64_t chunk = std::max<int64_t>(64, std::min<int64_t>(n_files, (int64_t)(256u << 20) / (sizeof(double) * t->t2)));
  int rc = reserve_files(t, std::min(chunk, n_files));
  if (rc) return rc;
  for (int64_t s = 0; s < n_files; s += chunk) {
    const int64_t n = std::min(chunk, n_files - s);
    if ((rc = posterior_chunk(t, d_Z + s * t->M, d_F + s * t->MD, n, false, d_out + s * t->tv, st))) return rc;
  }
  return ODIN_OK;
}

int tmat_mstep(odin_tmat* t, const double* d_acc, int min_div, int orthogonalize, int sweeps, cudaStream_t st) {
  const double* d_LU = d_acc;
  const double* d_RU = d_LU + (size_t)t->M * t->t2;
  const double* d_nframes = d_RU + (size_t)t->tv * t->MD + 1;
  int rc;
  {
    const size_t smem = square_smem(t->tv, t->tv * (t->D + 1));
    if (smem <= SMEM_LIMIT) {
      ODIN_CUDA_CHECK(cudaFuncSetAttribute(tmat_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tmat_solve_kernel<<<t->M, 256, smem, st>>>(t->tv, t->D, t->MD, d_LU, d_RU, t->d_Tm, t->d_flag, t->M, nullptr);
    } else {
      const int ctas = std::min(t->M, sm_count() * 4);
      if ((rc = reserve_gws(t, smem / sizeof(double), ctas))) return rc;
      tmat_solve_kernel<<<ctas, 256, 0, st>>>(t->tv, t->D, t->MD, d_LU, d_RU, t->d_Tm, t->d_flag, t->M, t->d_gws);
    }
    ODIN_LAUNCH_CHECK("tmat_solve_kernel");
  }
  if (min_div) {
    const size_t smem = square_smem(t->tv, 0);
    tmat_colsum_kernel<<<ceil_div(t->t2, 256), 256, 0, st>>>(d_LU, t->M, t->t2, t->d_small);
    ODIN_LAUNCH_CHECK("tmat_colsum_kernel");
    if (smem <= SMEM_LIMIT) {
      ODIN_CUDA_CHECK(cudaFuncSetAttribute(tmat_mindiv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      tmat_mindiv_kernel<<<1, 256, smem, st>>>(t->tv, t->M, t->d_small, d_nframes, t->d_U, t->d_flag, nullptr);
    } else {
      if ((rc = reserve_gws(t, smem / sizeof(double), 1))) return rc;
      tmat_mindiv_kernel<<<1, 1024, 0, st>>>(t->tv, t->M, t->d_small, d_nframes, t->d_U, t->d_flag, t->d_gws);
    }
    ODIN_LAUNCH_CHECK("tmat_mindiv_kernel");
    // Tm <- U Tm (through the T_invS buffer, which is rebuilt by the refresh below)
    if ((rc = gemm(t->tv, (int)t->MD, t->tv, t->d_U, t->tv, 1, t->d_Tm, t->MD, 1, t->d_TinvS, t->MD, 0.0, st))) return rc;
    ODIN_CUDA_CHECK(cudaMemcpyAsync(t->d_Tm, t->d_TinvS, sizeof(double) * (size_t)t->tv * t->MD, cudaMemcpyDeviceToDevice, st));
  }
  static const int eig_lib_min = [] { const char* e = getenv("ODIN_TMAT_EIG_LIB_MIN"); return e ? atoi(e) : 64; }();   // measured: own solver 0.8 / 2.2 / 5.9 ms M-step at tv 32 / 64 / 96, Dsyevd 2.5 / 2.0 / 4.3
  const bool eig_lib = t->tv >= eig_lib_min && getenv("ODIN_TMAT_NO_PREROT") == nullptr && cusolver_ready();
  bool prerot = false;
  if (orthogonalize && t->tv > 1 && t->tv <= TMAT_GRAM_MAX && !eig_lib) {
    prerot = true;
    // pre-rotation through the Gram matrix (three passes over T) -- see tmat_eig_kernel
    const int tv = t->tv;
    const size_t need = (size_t)33 * tv * tv;   // G | 32 split-K slices
    if ((rc = reserve_gws(t, need, 1))) return rc;
    double* G = t->d_gws;
    if ((rc = gemm(tv, tv, (int)t->MD, t->d_Tm, t->MD, 1, t->d_Tm, 1, t->MD, G, tv, 0.0, st, G + (size_t)tv * tv,
                   (int64_t)32 * tv * tv)))
      return rc;
    const size_t smem = sizeof(double) * (2 * (size_t)tv * (tv + 1) + tv + 4);
    ODIN_CUDA_CHECK(cudaFuncSetAttribute(tmat_eig_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tmat_eig_kernel<<<1, ET, smem, st>>>(G, tv, t->d_U);
    ODIN_LAUNCH_CHECK("tmat_eig_kernel");
    // T <- U^T T (through the T_invS buffer, rebuilt by the refresh at the end)
    if ((rc = gemm(tv, (int)t->MD, tv, t->d_U, 1, tv, t->d_Tm, t->MD, 1, t->d_TinvS, t->MD, 0.0, st))) return rc;
    ODIN_CUDA_CHECK(cudaMemcpyAsync(t->d_Tm, t->d_TinvS, sizeof(double) * (size_t)tv * t->MD, cudaMemcpyDeviceToDevice, st));
  }
  if (orthogonalize && t->tv > 1 && eig_lib) {
    // the same pre-rotation beyond the shared-memory eigen-solver: at tv 400, 2048 x 60 the one-sided sweeps alone took
    // 11 passes over the 393 MB matrix (4 389 launches, 0.97 s of a 1.24 s M-step; tools/tmat_scale.py)
    const int tv = t->tv;
    const size_t need = (size_t)33 * tv * tv;   // G | 32 split-K slices, reused as eigenvalues | info | Dsyevd workspace
    if ((rc = reserve_gws(t, need, 1))) return rc;
    double* G = t->d_gws;
    double* W = G + (size_t)tv * tv;
    int* info = reinterpret_cast<int*>(W + tv);
    double* work = W + tv + 2;
    int lwork = 0;
    Cusolver& cs = g_cusolver;
    const int CUSOLVER_EIG_MODE_VECTOR = 1, CUBLAS_FILL_MODE_LOWER = 0;
    if (cs.set_stream(cs.handle, st) == 0 &&
        cs.bufsize(cs.handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, tv, G, tv, W, &lwork) == 0 &&
        (size_t)lwork + tv + 2 <= (size_t)32 * tv * tv) {
      if ((rc = gemm(tv, tv, (int)t->MD, t->d_Tm, t->MD, 1, t->d_Tm, 1, t->MD, G, tv, 0.0, st, W, (int64_t)32 * tv * tv))) return rc;
      if (cs.syevd(cs.handle, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, tv, G, tv, W, work, lwork, info) != 0)
        return set_error(ODIN_ECUDA, "cusolverDnDsyevd failed (tv %d)", tv);
      tmat_eigperm_kernel<<<std::min(sm_count() * 4, ceil_div(tv * tv, 256)), 256, 0, st>>>(G, info, tv, t->d_U);
      ODIN_LAUNCH_CHECK("tmat_eigperm_kernel");
      if ((rc = gemm(tv, (int)t->MD, tv, t->d_U, 1, tv, t->d_Tm, t->MD, 1, t->d_TinvS, t->MD, 0.0, st))) return rc;
      ODIN_CUDA_CHECK(cudaMemcpyAsync(t->d_Tm, t->d_TinvS, sizeof(double) * (size_t)tv * t->MD, cudaMemcpyDeviceToDevice, st));
      prerot = true;
    }
  }
  bool rows_orthogonal = false;
  if (orthogonalize && prerot && getenv("ODIN_TMAT_NO_GRAMCHECK") == nullptr) {
    // one more Gram matrix (a ~60 us GEMM at tv = 64) instead of a verifying sweep of tv - 1 launches
    const int tv = t->tv;
    double* G = t->d_gws;   // (33 tv^2 doubles reserved by the pre-rotation)
    int* d_rot = t->d_flag + 1;
    if ((rc = gemm(tv, tv, (int)t->MD, t->d_Tm, t->MD, 1, t->d_Tm, 1, t->MD, G, tv, 0.0, st, G + (size_t)tv * tv,
                   (int64_t)32 * tv * tv)))
      return rc;
    ODIN_CUDA_CHECK(cudaMemsetAsync(d_rot, 0, sizeof(int), st));
    tmat_gramcheck_kernel<<<ceil_div(tv * tv, 256), 256, 0, st>>>(G, tv, d_rot);
    ODIN_LAUNCH_CHECK("tmat_gramcheck_kernel");
    int h_rot = 1;
    ODIN_CUDA_CHECK(cudaMemcpyAsync(&h_rot, d_rot, sizeof(int), cudaMemcpyDeviceToHost, st));
    ODIN_CUDA_CHECK(cudaStreamSynchronize(st));
    rows_orthogonal = h_rot == 0;
  }
  if (orthogonalize && t->tv > 1) {
    const int np = t->tv + (t->tv & 1);
    // sweeps until one of them rotates nothing (quadratic convergence: 6-9 sweeps in fp64), at most `sweeps`;
    // the rotation counter is read back once per sweep (the M-step runs once per EM iteration).  A sweep is
    // np - 1 dependent launches of a few microseconds each, i.e. launch-bound: it is captured ONCE into a CUDA
    // graph (the launches differ only in the round number) and replayed.
    int* d_rot = t->d_flag + 1;
    if (t->sweep_graph == nullptr) {
      cudaGraph_t graph = nullptr;
      cudaStream_t cs = nullptr;   // the caller's stream may be the legacy default stream, which cannot be captured
      ODIN_CUDA_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
      cudaError_t e_beg = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
      if (e_beg != cudaSuccess) {
        cudaStreamDestroy(cs);
        return set_error(ODIN_ECUDA, "cudaStreamBeginCapture: %s", cudaGetErrorString(e_beg));
      }
      cudaError_t e_cap = cudaMemsetAsync(d_rot, 0, sizeof(int), cs);
      for (int r = 0; r < np - 1 && e_cap == cudaSuccess; ++r) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(np / 2 * JC));
        cfg.blockDim = dim3(JT);
        cfg.stream = cs;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = JC; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        e_cap = cudaLaunchKernelEx(&cfg, tmat_jacobi_kernel, t->d_Tm, t->tv, t->MD, np, r, d_rot);
      }
      cudaError_t e_end = cudaStreamEndCapture(cs, &graph);
      cudaStreamDestroy(cs);
      if (e_cap != cudaSuccess || e_end != cudaSuccess || graph == nullptr) {
        if (graph) cudaGraphDestroy(graph);
        return set_error(ODIN_ECUDA, "capturing the Jacobi sweep: %s", cudaGetErrorString(e_cap != cudaSuccess ? e_cap : e_end));
      }
      cudaGraphExec_t exec = nullptr;
      cudaError_t e_inst = cudaGraphInstantiate(&exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e_inst != cudaSuccess) return set_error(ODIN_ECUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e_inst));
      t->sweep_graph = exec;
    }
    for (int sweep = 0; sweep < sweeps && !rows_orthogonal; ++sweep) {
      ODIN_CUDA_CHECK(cudaGraphLaunch((cudaGraphExec_t)t->sweep_graph, st));
      g_launches.fetch_add(np - 1, std::memory_order_relaxed);
      int h_rot = 0;
      ODIN_CUDA_CHECK(cudaMemcpyAsync(&h_rot, d_rot, sizeof(int), cudaMemcpyDeviceToHost, st));
      ODIN_CUDA_CHECK(cudaStreamSynchronize(st));
      if (h_rot == 0) break;
    }
    tmat_rownorm_kernel<<<t->tv, 256, 0, st>>>(t->d_Tm, t->MD, t->d_small);
    ODIN_LAUNCH_CHECK("tmat_rownorm_kernel");
    tmat_order_kernel<<<1, 256, 0, st>>>(t->d_small, t->tv, t->d_perm);
    ODIN_LAUNCH_CHECK("tmat_order_kernel");
    dim3 grid((unsigned)std::min<int64_t>(ceil_div<int64_t>(t->MD, 256), 64), (unsigned)t->tv);
    tmat_gather_rows_kernel<<<grid, 256, 0, st>>>(t->d_Tm, t->d_TinvS, t->d_perm, t->MD);
    ODIN_LAUNCH_CHECK("tmat_gather_rows_kernel");
    ODIN_CUDA_CHECK(cudaMemcpyAsync(t->d_Tm, t->d_TinvS, sizeof(double) * (size_t)t->tv * t->MD, cudaMemcpyDeviceToDevice, st));
  }
  return tmat_refresh(t, st);
}

}  // namespace odin
