// GMM-UBM Baum-Welch on the 5th-generation tensor cores, 3xFP16 split precision
// (tcgen05.mma kind::f16, fp32 accumulation in TMEM).  sm_100a only.  impl = 3.
//
// Reference arithmetic: odin/ml/gmm_tmat.py:1012-1041 (_fast_expectation) with the
// cached constants of :493-504; the same two-GEMM formulation as gmm_tc.cu:
//
//   lp2[m,b]  = sum_k W[m,k] A[b,k],  A[b,:] = [x^2 (D) | x (D) | 0.. | 1 (k=120) | 0..]
//                                     W[m,:] = log2(e) [-0.5 prec | mu prec | 0.. | cst | 0..]
//   pass 1:     lse2[b] = log2 sum_m 2^lp2[m,b]
//   pass 2:     P[m,b] = 2^(lp2[m,b] - lse2[b]);   stat[j,m] += sum_b A[b,j] P[m,b]
//
// Why fp16 and not tf32: both carry an 11-bit significand, so `x = hi + lo` with
// three MMAs per product (hi*hi + hi*lo + lo*hi) keeps ~22 bits either way, but
// kind::f16 issues K = 16 per instruction where kind::tf32 issues K = 8 -- twice the
// rate -- and its operands are half the bytes in shared memory / TMEM.  The 5-bit
// exponent is handled with EXACT power-of-two scalings:
//   * column k of A is scaled by 2^ea[k], ea chosen from max|A[:,k]| over the call's
//     frames (gmm_h_range_kernel) so that |A'| < 2^14; W'[m,k] = W[m,k] 2^-ea[k] 2^ew[m]
//     with a per-mixture 2^ew[m] normalising the row to < 2^14; the epilogue undoes
//     2^ew[m] with one multiply (folded into an FFMA in pass 2);
//   * P is scaled by 2^14 and the statistics are de-scaled by 2^(-14-ea[j]) when the
//     fp32 TMEM accumulator is drained into the fp64 statistics.
// Halves below 2^-14 go subnormal (absolute error 2^-25 on a 2^14 scale): harmless.
//
// The frame operands are built ONCE per sub-batch by gmm_h_image_kernel as the exact
// byte image of the K-major SWIZZLE_128B shared-memory tiles the MMAs read, so the
// hot kernels have no producer warps at all: one thread streams tiles with 1-D bulk
// copies (cp.async.bulk -> UBLKCP, the TMA engine) into mbarrier-guarded rings.
//
// pass 1  gmm_h_lse_kernel    frames on the TMEM lanes, 256 mixtures per CTA:
//           D1[128 frames, 256 mix] = Ahi*Whi + Alo*Whi + Ahi*Wlo   (SS mode, K = 128)
//           W' image resident in smem (128 KB), A ring 3 x 32 KB, D1 double-buffered
//           (2 x 256 TMEM columns); 8 epilogue warps run an online max / sum.
// pass 2  gmm_h_stats_kernel  mixtures on the lanes, 128 per CTA, 64-frame tiles:
//           D1[128 mix, 64 frames]  = Whi*Ahi + Whi*Alo + Wlo*Ahi   (W' hi|lo in TMEM)
//           P' = 2^(D1 dsc[m] + 14 - lse2[b]) -> fp16 hi|lo written back IN PLACE over
//           D1's TMEM columns (tcgen05.st), never touching shared memory
//           D2[128 mix, 128 j]     += Phi*Thi + Phi*Tlo + Plo*Thi   (P' from TMEM, T' =
//           transposed frame tile straight from the bulk-copied image)
//           TMEM: [0,64) Whi | [64,128) Wlo | [128,256) D2 | [256,512) 4 x (D1 / P')
//           smem: A ring 3 x 32 KB (freed by GEMM 1), T ring 3 x 32 KB (freed by GEMM 2)
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "gmm.cuh"
#include "tc_ptx.cuh"

namespace odin {

namespace hk {
constexpr int K = 128;        // padded contraction length of GEMM 1 / rows of GEMM 2
constexpr int K_ONE = 120;    // column holding the constant 1 (-> cst, Z)
constexpr int MAX_D = 60;     // 2*D <= 120 and D % 4 == 0
constexpr int CM1 = 256;      // pass 1: mixtures per CTA
constexpr int CM2 = 128;      // pass 2: mixtures per CTA
constexpr int TF1 = 128;      // pass 1: frames per tile = frames per image super-tile
constexpr int TF2 = 64;       // pass 2: frames per tile
constexpr int P_EXP = 14;     // P' = P * 2^14
constexpr int SUPER_A_BYTES = 65536;  // A image per 128 frames: [part 2][kblock 2][128 rows][128 B]
constexpr int TILE_T_BYTES = 32768;   // T image per 64 frames:  [part 2][128 rows][128 B]
constexpr int THREADS = 320;  // warps 0-7 epilogue, 8 MMA issuer, 9 bulk-copy producer

// ---- pass 1 shared memory map (after 1024-byte alignment)
constexpr uint32_t L_W = 0;                       // 128 KB W' image [part][kblock][256 rows][128 B]
constexpr uint32_t L_A = 131072;                  // 3 x (hi 16 KB | lo 16 KB) of one k-block
constexpr int L_ASLOTS = 3;
constexpr uint32_t L_DSC = L_A + L_ASLOTS * 32768;  // 256 floats
constexpr uint32_t L_BAR = L_DSC + 1024;
constexpr uint32_t L_SMEM = L_BAR + 256 + 1024;
constexpr int LB_W = 0, LB_AFULL = 1, LB_AEMPTY = 4, LB_D1FULL = 7, LB_D1EMPTY = 9, LB_TMEM = 12;

// ---- pass 2 shared memory map
constexpr uint32_t S_A = 0;                       // 3 x 32 KB: [hi kb0 | hi kb1 | lo kb0 | lo kb1] x 64 rows
constexpr int S_ASLOTS = 3;                       //   (freed by GEMM 1, so a slot turns around in ~5 MMA groups)
constexpr uint32_t S_T = S_A + S_ASLOTS * 32768;  // 3 x 32 KB: [Thi 16 KB | Tlo 16 KB] (freed by GEMM 2)
constexpr int S_TSLOTS = 3;
constexpr uint32_t S_DSJ = S_T + S_TSLOTS * 32768; // 128 doubles: 2^(-14 - ea[j])
constexpr uint32_t S_BAR = S_DSJ + 1024;
constexpr uint32_t S_SMEM = S_BAR + 256 + 1024;
constexpr int SB_AFULL = 0, SB_AEMPTY = 3, SB_TFULL = 6, SB_TEMPTY = 9, SB_D1FULL = 12, SB_PFULL = 16,
              SB_BUFEMPTY = 20, SB_D2FULL = 24, SB_D2EMPTY = 25, SB_TMEM = 28;
constexpr uint32_t TM_WHI = 0, TM_WLO = 64, TM_D2 = 128, TM_BUF = 256;
constexpr int NBUF = 4;
constexpr int LOOKAHEAD = 2;  // GEMM 1 of tile it+2 is issued before GEMM 2 of tile it
}  // namespace hk

struct HScale {          // device-resident, rebuilt per E-step call
  unsigned amax[64];     // float bits of max |x_d| over the call's frames
  float ascale[128];     // 2^ea[k]
  double dscale[128];    // 2^(-14 - ea[k])
  int ea[128];
};

// ---------------------------------------------------------------------------
// range of the data -> exact power-of-two column scales
// ---------------------------------------------------------------------------
__global__ void gmm_h_range_kernel(const float* __restrict__ X, int64_t N, int D, HScale* __restrict__ sc) {
  __shared__ unsigned smax[64];
  const int d4 = D >> 2;
  if (threadIdx.x < 64) smax[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t total4 = N * d4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;   // multiple of d4 (blockDim = 16 * d4)
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c = (int)(i0 % d4);
  float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
  const float4* X4 = reinterpret_cast<const float4*>(X);
  for (int64_t i = i0; i < total4; i += stride) {
    const float4 v = __ldg(X4 + i);
    m0 = fmaxf(m0, fabsf(v.x)); m1 = fmaxf(m1, fabsf(v.y)); m2 = fmaxf(m2, fabsf(v.z)); m3 = fmaxf(m3, fabsf(v.w));
  }
  atomicMax(&smax[4 * c + 0], __float_as_uint(m0));
  atomicMax(&smax[4 * c + 1], __float_as_uint(m1));
  atomicMax(&smax[4 * c + 2], __float_as_uint(m2));
  atomicMax(&smax[4 * c + 3], __float_as_uint(m3));
  __syncthreads();
  if (threadIdx.x < D) atomicMax(&sc->amax[threadIdx.x], smax[threadIdx.x]);
}

__global__ void gmm_h_scale_kernel(int D, HScale* __restrict__ sc) {
  const int k = threadIdx.x;
  if (k >= hk::K) return;
  float amax = 0.f;
  int ea = 0;
  if (k < D) { const float a = __uint_as_float(sc->amax[k]); amax = a * a; }
  else if (k < 2 * D) amax = __uint_as_float(sc->amax[k - D]);
  else if (k == hk::K_ONE) amax = 1.f;
  if (k == hk::K_ONE) {
    ea = 14;                       // the constant column carries exactly 2^14
  } else if (amax > 0.f && isfinite(amax)) {
    int e;
    frexpf(amax, &e);              // amax = f * 2^e, f in [0.5, 1)
    ea = 14 - e;                   // amax * 2^ea in [2^13, 2^14)
    ea = max(-100, min(100, ea));
  }
  sc->ea[k] = ea;
  sc->ascale[k] = exp2f((float)ea);
  sc->dscale[k] = exp2((double)(-hk::P_EXP - ea));
}

// ---------------------------------------------------------------------------
// model -> scaled fp16 hi|lo operand images
//   Wplain [Mpad][2][64] u32   : row m, part (hi, lo), K pairs        (pass 2, copied into TMEM)
//   Wimg   [Mpad/256][2][2][256 rows][128 B]  pre-swizzled smem image (pass 1)
//   dsc    [Mpad]              : 2^-ew[m] (2^90 for padding mixtures: lp2 -> -2^118)
// ---------------------------------------------------------------------------
__global__ void gmm_h_prepare_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                                     const float* __restrict__ w, int D, int M, int Mpad,
                                     const HScale* __restrict__ sc, uint32_t* __restrict__ Wplain,
                                     uint32_t* __restrict__ Wimg, float* __restrict__ dsc) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mpad) return;
  const double LOG2E = 1.4426950408889634074;
  float v[hk::K];
#pragma unroll 1
  for (int k = 0; k < hk::K; ++k) v[k] = 0.f;
  float scale_out = 1.f;
  if (m < M) {
    double C = 0.0;
    for (int d = 0; d < D; ++d) {
      const double vv = (double)var[(size_t)d * M + m] + ODIN_GMM_EPS;
      const double p = 1.0 / vv;
      const double mu = (double)mean[(size_t)d * M + m];
      C += mu * mu * p + log(vv);
      v[d] = (float)(-0.5 * p * LOG2E * exp2((double)-sc->ea[d]));
      v[D + d] = (float)(mu * p * LOG2E * exp2((double)-sc->ea[D + d]));
    }
    C -= 2.0 * log((double)w[m] + ODIN_GMM_EPS);
    v[hk::K_ONE] = (float)(-0.5 * (C + (double)D * 1.8378770664093454835606594728112) * LOG2E *
                           exp2((double)-sc->ea[hk::K_ONE]));
    float rmax = 0.f;
    for (int k = 0; k < hk::K; ++k) rmax = fmaxf(rmax, fabsf(v[k]));
    int ew = 0;
    if (rmax > 0.f && isfinite(rmax)) {
      int e;
      frexpf(rmax, &e);
      ew = max(-100, min(100, 14 - e));
    }
    const float up = exp2f((float)ew);
    for (int k = 0; k < hk::K; ++k) v[k] *= up;
    scale_out = exp2f((float)-ew);
  } else {
    v[hk::K_ONE] = -16384.f;   // * A' = 2^14 -> D1' = -2^28; dsc = 2^90 -> lp2 = -2^118: posterior exactly 0
    scale_out = exp2f(90.f);
  }
  dsc[m] = scale_out;
  const int chunk = m / hk::CM1, r = m % hk::CM1;
  for (int k = 0; k < hk::K; k += 2) {
    uint32_t hi, lo;
    ptx::split_h2(v[k], v[k + 1], hi, lo);
    Wplain[((size_t)m * 2 + 0) * 64 + (k >> 1)] = hi;
    Wplain[((size_t)m * 2 + 1) * 64 + (k >> 1)] = lo;
    const int kb = k >> 6, q = (k & 63) >> 3, e = k & 7;
    const size_t base = (size_t)chunk * 131072;
    const size_t off_hi = base + ((size_t)((0 * 2 + kb) * 256 + r)) * 128 + (size_t)(((q ^ (r & 7)) << 4) + e * 2);
    const size_t off_lo = base + ((size_t)((1 * 2 + kb) * 256 + r)) * 128 + (size_t)(((q ^ (r & 7)) << 4) + e * 2);
    Wimg[off_hi >> 2] = hi;
    Wimg[off_lo >> 2] = lo;
  }
}

// ---------------------------------------------------------------------------
// frames -> operand images (one CTA per super-tile of 128 frames)
//   imgA [super][part 2][kblock 2][128 frames][128 B]   rows = frames, K-major, SWIZZLE_128B
//   imgT [tile64][part 2][128 rows j][128 B]            rows = j, K = 64 frames
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gmm_h_image_kernel(const float* __restrict__ X, int64_t N, int D,
                                                          const HScale* __restrict__ sc,
                                                          unsigned char* __restrict__ imgA,
                                                          unsigned char* __restrict__ imgT) {
  __shared__ float xs[hk::TF1 * (hk::MAX_D + 1)];
  __shared__ float asc[hk::K];
  const int tid = threadIdx.x;
  const int DP = D + 1;
  const int64_t f0 = (int64_t)blockIdx.x * hk::TF1;
  if (tid < hk::K) asc[tid] = sc->ascale[tid];
  {
    const int d4 = D >> 2;
    const float4* X4 = reinterpret_cast<const float4*>(X);
    const int64_t total4 = N * d4;
    for (int i = tid; i < hk::TF1 * d4; i += 256) {
      const int64_t g = f0 * d4 + i;
      const float4 v = g < total4 ? __ldg(X4 + g) : make_float4(0.f, 0.f, 0.f, 0.f);
      const int r = i / d4, c = i - r * d4;
      float* dst = xs + r * DP + 4 * c;
      dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
    }
  }
  __syncthreads();
  auto value = [&](int f, int k) -> float {   // A'[f][k] = T'[k][f]
    float x = 0.f;
    if (k < D) { const float t = xs[f * DP + k]; x = t * t; }
    else if (k < 2 * D) x = xs[f * DP + k - D];
    else if (k == hk::K_ONE) x = 1.f;
    return x * asc[k];
  };
  // ---- A image: item = (frame row r, 16-byte chunk q of the 128-wide K axis)
  unsigned char* outA = imgA + (size_t)blockIdx.x * hk::SUPER_A_BYTES;
  for (int it = tid; it < hk::TF1 * 16; it += 256) {
    const int r = it >> 4, q = it & 15;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) ptx::split_h2(value(r, 8 * q + 2 * e), value(r, 8 * q + 2 * e + 1), hi[e], lo[e]);
    const int kb = q >> 3, qq = q & 7;
    const size_t o = ((size_t)(kb * 128 + r)) * 128 + (size_t)((qq ^ (r & 7)) << 4);
    *reinterpret_cast<uint4*>(outA + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(outA + 2 * 16384 + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
  // ---- T image: item = (half hh, row j, 16-byte chunk c = 8 frames)
  for (int it = tid; it < 2 * hk::K * 8; it += 256) {
    const int c = it & 7, j = (it >> 3) & 127, hh = it >> 10;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int f = 64 * hh + 8 * c + 2 * e;
      ptx::split_h2(value(f, j), value(f + 1, j), hi[e], lo[e]);
    }
    unsigned char* outT = imgT + ((size_t)blockIdx.x * 2 + hh) * hk::TILE_T_BYTES;
    const size_t o = (size_t)j * 128 + (size_t)((c ^ (j & 7)) << 4);
    *reinterpret_cast<uint4*>(outT + o) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4*>(outT + 16384 + o) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

struct HArgs {
  int64_t N;              // frames of this sub-batch
  int D, M;
  const unsigned char* imgA;
  const unsigned char* imgT;
  const uint32_t* Wplain;
  const unsigned char* Wimg;
  const float* dsc;       // [Mpad]
  const double* dscale;   // [128]
  const float* cb;        // pass 2: [tiles64 * 64]
  float2* part;           // pass 1: [nparts][part_stride]
  int64_t part_stride;
  double* stats;
  int want_second;
  int flush_tiles;
  // segmented mode (gmm_utt_stats_hseg): per-utterance outputs, the accumulator is flushed where the utterance changes
  const int* tile_utt;    // [tiles64 + 1] utterance of every 64-frame tile (-1: padding), one sentinel behind the last
  float* uZ;              // [n_utt, M]      zeroed by the caller
  float* uF;              // [n_utt, M * D]  raw first-order sums (centred afterwards)
};

// ---------------------------------------------------------------------------
// pass 1: per-(256-mixture chunk, column half) partial log-sum-exp
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(hk::THREADS, 1) gmm_h_lse_kernel(HArgs a) {
  using namespace hk;
  using namespace ptx;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw_addr = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  unsigned char* smem = smem_dyn + pad;
  const uint32_t sbase = raw_addr + pad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x;
  const int64_t n_super = (a.N + TF1 - 1) / TF1;
  const int64_t my_tiles = (n_super > (int64_t)blockIdx.y) ? (n_super - blockIdx.y + gridDim.y - 1) / gridDim.y : 0;
  auto bar = [&](int i) -> uint32_t { return sbase + L_BAR + 8u * i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + L_BAR + 8 * LB_TMEM);

  if (warp == 8) {
    if (lane == 0) {
      mbar_init(bar(LB_W), 1);
      for (int i = 0; i < L_ASLOTS; ++i) { mbar_init(bar(LB_AFULL + i), 1); mbar_init(bar(LB_AEMPTY + i), 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(bar(LB_D1FULL + i), 1); mbar_init(bar(LB_D1EMPTY + i), 256); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc512(bar(LB_TMEM));
  }
  if (tid < CM1) reinterpret_cast<float*>(smem + L_DSC)[tid] = a.dsc[chunk * CM1 + tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_s;

  if (warp == 9) {
    // ============================================== bulk-copy producer (one lane)
    if (lane == 0 && my_tiles > 0) {
      mbar_arrive_expect_tx(bar(LB_W), 131072u);
      for (int i = 0; i < 4; ++i)
        bulk_g2s(sbase + L_W + i * 32768, a.Wimg + (size_t)chunk * 131072 + (size_t)i * 32768, 32768u, bar(LB_W));
      int64_t pi = 0;
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int64_t st = blockIdx.y + it * gridDim.y;
        const unsigned char* src = a.imgA + (size_t)st * SUPER_A_BYTES;
        for (int kb = 0; kb < 2; ++kb, ++pi) {
          const int slot = (int)(pi % L_ASLOTS);
          mbar_wait(bar(LB_AEMPTY + slot), (uint32_t)(((pi / L_ASLOTS) & 1) ^ 1));
          mbar_arrive_expect_tx(bar(LB_AFULL + slot), 32768u);
          bulk_g2s(sbase + L_A + slot * 32768, src + (0 * 2 + kb) * 16384, 16384u, bar(LB_AFULL + slot));
          bulk_g2s(sbase + L_A + slot * 32768 + 16384, src + (1 * 2 + kb) * 16384, 16384u, bar(LB_AFULL + slot));
        }
      }
    }
  } else if (warp == 8) {
    // ========================================================== MMA issuer
    if (my_tiles > 0) {
      const uint32_t nt = (uint32_t)my_tiles;
      const uint32_t idesc = idesc_f16(CM1);
      const uint32_t bar0 = sbase + L_BAR;
      const uint64_t w0 = desc_k_sw128(sbase + L_W), descA = desc_k_sw128(sbase + L_A);
      mbar_wait(bar(LB_W), 0u);
      uint32_t slot = 0, spar = 0, buf = 0, bpar = 1;
      for (uint32_t it = 0; it < nt; ++it) {
        mbar_wait_fast(bar0 + 8u * (LB_D1EMPTY + buf), bpar);
        const uint32_t d = tmem + 256u * buf;
#pragma unroll
        for (int kb = 0; kb < 2; ++kb) {
          mbar_wait_fast(bar0 + 8u * (LB_AFULL + slot), spar);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t a_hi0 = descA + (uint64_t)(slot * (32768u >> 4)), a_lo0 = a_hi0 + (16384 >> 4);
            const uint64_t w_hi0 = w0 + (uint64_t)(((0 * 2 + kb) * 32768) >> 4);
            const uint64_t w_lo0 = w0 + (uint64_t)(((1 * 2 + kb) * 32768) >> 4);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t o = (uint64_t)((ks * 32) >> 4);
              mma_f16_ss(d, a_hi0 + o, w_hi0 + o, idesc, (kb | ks) ? 1u : 0u);
              mma_f16_ss(d, a_lo0 + o, w_hi0 + o, idesc, 1u);
              mma_f16_ss(d, a_hi0 + o, w_lo0 + o, idesc, 1u);
            }
            tc_commit(bar0 + 8u * (LB_AEMPTY + slot));
            if (kb == 1) tc_commit(bar0 + 8u * (LB_D1FULL + buf));
          }
          __syncwarp();
          if (++slot == L_ASLOTS) { slot = 0; spar ^= 1u; }
        }
        if (++buf == 2u) { buf = 0; bpar ^= 1u; }
      }
    }
  } else {
    // ============================================================ epilogue
    const int q = warp & 3, h = warp >> 2;
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const float4* dsc4 = reinterpret_cast<const float4*>(smem + L_DSC) + h * 32;
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int buf = (int)(it & 1);
      mbar_wait(bar(LB_D1FULL + buf), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      float m = 0.f, s = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        float v[32];
        tmem_ld32(tmem + 256u * buf + 128u * h + 32u * g + lane_field, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 sc4 = dsc4[g * 8 + i];
          v[4 * i] *= sc4.x; v[4 * i + 1] *= sc4.y; v[4 * i + 2] *= sc4.z; v[4 * i + 3] *= sc4.w;
        }
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
      mbar_arrive(bar(LB_D1EMPTY + buf));
      const int64_t st = blockIdx.y + it * gridDim.y;
      const int64_t f = st * TF1 + q * 32 + lane;
      if (f < a.N) a.part[(size_t)(chunk * 2 + h) * a.part_stride + f] = make_float2(m, s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc512(tmem);
  }
}

// pass 1 tail: partials -> cb[b] = 14 - lse2[b] (-1e30 for masked / padding frames);
// sum of log-likelihoods and frame count
__global__ void __launch_bounds__(256)
gmm_h_combine_kernel(const float2* __restrict__ part, int nparts, int64_t stride, int64_t n, int64_t npad,
                     const uint8_t* __restrict__ sad, float* __restrict__ cb, double* __restrict__ stat_L) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double lsum = 0.0, lcnt = 0.0;
  if (b < n) {
    float mx = -INFINITY;
    for (int c = 0; c < nparts; ++c) mx = fmaxf(mx, part[(size_t)c * stride + b].x);
    float s = 0.f;
    for (int c = 0; c < nparts; ++c) {
      const float2 p = part[(size_t)c * stride + b];
      s += p.y * exp2f(p.x - mx);
    }
    const float l2 = mx + log2f(s);
    const bool on = sad == nullptr || sad[b] != 0;
    cb[b] = on ? ((float)hk::P_EXP - l2) : -1e30f;
    if (on) { lsum = (double)l2 * 0.69314718055994530942; lcnt = 1.0; }
  } else if (b < npad) {
    cb[b] = -1e30f;
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
// pass 2: posteriors and statistics
// ---------------------------------------------------------------------------
// SEG: the tiles of a CTA are CONTIGUOUS, every tile belongs to one utterance (the caller pads utterances to whole
// tiles) and the accumulator is drained into that utterance's float32 rows whenever the next tile starts another one --
// the per-utterance statistics of thousands of short utterances in one launch (the i-vector extractor's
// input).  !SEG: tiles strided over the CTAs, periodic flush into the packed fp64 statistics of the whole call.
template <bool SEG>
__global__ void __launch_bounds__(hk::THREADS, 1) gmm_h_stats_kernel(HArgs a) {
  using namespace hk;
  using namespace ptx;
  extern __shared__ unsigned char smem_dyn[];
  const uint32_t raw_addr = smem_u32(smem_dyn);
  const uint32_t pad = (1024u - (raw_addr & 1023u)) & 1023u;
  unsigned char* smem = smem_dyn + pad;
  const uint32_t sbase = raw_addr + pad;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x;
  const int D = a.D;
  const int64_t n_tiles = 2 * ((a.N + TF1 - 1) / TF1);   // the images are padded to whole super-tiles
  const int64_t seg_per = (n_tiles + gridDim.y - 1) / gridDim.y;
  const int64_t seg_lo = seg_per * blockIdx.y;
  const int64_t seg_left = n_tiles - seg_lo;
  const int64_t my_tiles = SEG ? (seg_left <= 0 ? 0 : (seg_left < seg_per ? seg_left : seg_per))
                               : ((n_tiles > (int64_t)blockIdx.y) ? (n_tiles - blockIdx.y + gridDim.y - 1) / gridDim.y : 0);
  auto tile_of = [&](int64_t it) -> int64_t { return SEG ? seg_lo + it : (int64_t)blockIdx.y + it * gridDim.y; };
  auto bar = [&](int i) -> uint32_t { return sbase + S_BAR + 8u * i; };
  volatile uint32_t* tmem_ptr_s = reinterpret_cast<volatile uint32_t*>(smem + S_BAR + 8 * SB_TMEM);

  if (warp == 8) {
    if (lane == 0) {
      for (int i = 0; i < S_ASLOTS; ++i) { mbar_init(bar(SB_AFULL + i), 1); mbar_init(bar(SB_AEMPTY + i), 1); }
      for (int i = 0; i < S_TSLOTS; ++i) { mbar_init(bar(SB_TFULL + i), 1); mbar_init(bar(SB_TEMPTY + i), 1); }
      for (int i = 0; i < NBUF; ++i) {
        mbar_init(bar(SB_D1FULL + i), 1);
        mbar_init(bar(SB_PFULL + i), 256);
        mbar_init(bar(SB_BUFEMPTY + i), 1);
      }
      mbar_init(bar(SB_D2FULL), 1);
      mbar_init(bar(SB_D2EMPTY), 256);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc512(bar(SB_TMEM));
  }
  if (tid < K) reinterpret_cast<double*>(smem + S_DSJ)[tid] = a.dscale[tid];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_ptr_s;
  if (warp < 8) {  // W' rows -> TMEM: lane = mixture row; warps 0-3 the hi part, 4-7 the lo part
    const int q = warp & 3, part = warp >> 2;
    const int row = q * 32 + lane;
    const uint4* src = reinterpret_cast<const uint4*>(a.Wplain + ((size_t)(chunk * CM2 + row) * 2 + part) * 64);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 t = __ldg(src + c * 8 + i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
      }
      tmem_st32(tmem + (part ? TM_WLO : TM_WHI) + 32u * c + ((uint32_t)(q * 32) << 16), v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp == 9) {
    // ============================================== bulk-copy producer (one lane)
    if (lane == 0) {
      for (int64_t it = 0; it < my_tiles; ++it) {
        const int64_t tile = tile_of(it);
        const int64_t st = tile >> 1;
        const int half = (int)(tile & 1);
        {
          const int sa = (int)(it % S_ASLOTS);
          mbar_wait(bar(SB_AEMPTY + sa), (uint32_t)(((it / S_ASLOTS) & 1) ^ 1));
          mbar_arrive_expect_tx(bar(SB_AFULL + sa), 32768u);
          const unsigned char* src = a.imgA + (size_t)st * SUPER_A_BYTES + (size_t)half * 8192;
#pragma unroll
          for (int p = 0; p < 4; ++p)
            bulk_g2s(sbase + S_A + sa * 32768 + p * 8192, src + (size_t)p * 16384, 8192u, bar(SB_AFULL + sa));
        }
        {
          const int s4 = (int)(it % S_TSLOTS);
          mbar_wait(bar(SB_TEMPTY + s4), (uint32_t)(((it / S_TSLOTS) & 1) ^ 1));
          mbar_arrive_expect_tx(bar(SB_TFULL + s4), 32768u);
          bulk_g2s(sbase + S_T + s4 * 32768, a.imgT + (size_t)tile * TILE_T_BYTES, 32768u, bar(SB_TFULL + s4));
        }
      }
    }
  } else if (warp == 8) {
    // ========================================================== MMA issuer
    // Everything between two groups of MMAs is kept to a handful of 32-bit instructions (ring
    // cursors instead of div/mod, one try_wait per barrier): while this warp does bookkeeping the
    // tensor pipe lives off its instruction queue.
    if (my_tiles > 0) {
      const uint32_t nt = (uint32_t)my_tiles;
      const uint32_t idesc1 = idesc_f16(TF2), idesc2 = idesc_f16(K);
      const uint32_t bar0 = sbase + S_BAR;
      const uint64_t descA = desc_k_sw128(sbase + S_A), descT = desc_k_sw128(sbase + S_T);
      uint32_t g1_b = 0, g1_bpar = 1, g1_a = 0, g1_apar = 0;     // cursors of GEMM 1: D1/P' buffer, A slot
      uint32_t g2_b = 0, g2_bpar = 0, g2_t = 0, g2_tpar = 0;     // cursors of GEMM 2: P' buffer, T slot
      uint32_t flush_left = (uint32_t)a.flush_tiles, d2_phase = 0;
      bool acc = false, need_d2_empty = false;
      // SEG: utterance of the current tile and of the next two (read two tiles ahead: the load is off the critical path)
      int u_cur = 0, u_n1 = 0;
      uint32_t seg_it = 0;
      if (SEG) { u_cur = __ldg(a.tile_utt + seg_lo); u_n1 = nt > 1 ? __ldg(a.tile_utt + seg_lo + 1) : -2; }
      auto issue_g1 = [&]() {
        mbar_wait_fast(bar0 + 8u * (SB_BUFEMPTY + g1_b), g1_bpar);
        mbar_wait_fast(bar0 + 8u * (SB_AFULL + g1_a), g1_apar);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d = tmem + TM_BUF + 64u * g1_b;
          const uint64_t a0 = descA + (uint64_t)(g1_a * (32768u >> 4));
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const int kb = kk >> 2, ks = kk & 3;
            const uint64_t b_hi = a0 + (uint64_t)(((0 * 2 + kb) * 8192 + ks * 32) >> 4);
            const uint64_t b_lo = a0 + (uint64_t)(((1 * 2 + kb) * 8192 + ks * 32) >> 4);
            mma_f16_ts(d, tmem + TM_WHI + 8u * kk, b_hi, idesc1, kk > 0 ? 1u : 0u);
            mma_f16_ts(d, tmem + TM_WHI + 8u * kk, b_lo, idesc1, 1u);
            mma_f16_ts(d, tmem + TM_WLO + 8u * kk, b_hi, idesc1, 1u);
          }
          tc_commit(bar0 + 8u * (SB_AEMPTY + g1_a));
          tc_commit(bar0 + 8u * (SB_D1FULL + g1_b));
        }
        __syncwarp();
        if (++g1_b == NBUF) { g1_b = 0; g1_bpar ^= 1u; }
        if (++g1_a == S_ASLOTS) { g1_a = 0; g1_apar ^= 1u; }
      };
      auto issue_g2 = [&](bool last) {
        mbar_wait_fast(bar0 + 8u * (SB_PFULL + g2_b), g2_bpar);
        mbar_wait_fast(bar0 + 8u * (SB_TFULL + g2_t), g2_tpar);
        if (need_d2_empty) {
          mbar_wait_fast(bar0 + 8u * SB_D2EMPTY, d2_phase);
          d2_phase ^= 1u;
          need_d2_empty = false;
        }
        tc_fence_after();
        bool flush;
        if (SEG) {
          const int u_n2 = (seg_it + 2 < nt) ? __ldg(a.tile_utt + seg_lo + seg_it + 2) : -2;
          flush = last || u_n1 != u_cur || flush_left == 1u;   // ... and every flush_tiles tiles inside a long utterance
          --flush_left;
          u_cur = u_n1; u_n1 = u_n2; ++seg_it;
        } else {
          flush = (--flush_left == 0u) || last;
        }
        if (elect_one()) {
          const uint32_t d = tmem + TM_D2;
          const uint32_t p0 = tmem + TM_BUF + 64u * g2_b;
          const uint64_t t_hi0 = descT + (uint64_t)(g2_t * (32768u >> 4)), t_lo0 = t_hi0 + (16384 >> 4);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t p_hi = p0 + 32u * (ks >> 1) + 8u * (ks & 1), p_lo = p_hi + 16u;
            const uint64_t o = (uint64_t)((ks * 32) >> 4);
            mma_f16_ts(d, p_hi, t_hi0 + o, idesc2, (acc || ks > 0) ? 1u : 0u);
            mma_f16_ts(d, p_hi, t_lo0 + o, idesc2, 1u);
            mma_f16_ts(d, p_lo, t_hi0 + o, idesc2, 1u);
          }
          tc_commit(bar0 + 8u * (SB_TEMPTY + g2_t));
          tc_commit(bar0 + 8u * (SB_BUFEMPTY + g2_b));
          if (flush) tc_commit(bar0 + 8u * SB_D2FULL);
        }
        __syncwarp();
        acc = !flush;
        if (flush) { need_d2_empty = true; flush_left = (uint32_t)a.flush_tiles; }
        if (++g2_b == NBUF) { g2_b = 0; g2_bpar ^= 1u; }
        if (++g2_t == S_TSLOTS) { g2_t = 0; g2_tpar ^= 1u; }
      };
      uint32_t issued = 0;
      for (; issued < (uint32_t)LOOKAHEAD && issued < nt; ++issued) issue_g1();
      for (uint32_t it = 0; it < nt; ++it) {
        if (issued < nt) { issue_g1(); ++issued; }
        issue_g2(it + 1 == nt);
      }
    }
  } else {
    // ============================================================ epilogue
    const int q = warp & 3, h = warp >> 2;
    const int row = q * 32 + lane;                  // mixture within the chunk == TMEM lane
    const uint32_t lane_field = (uint32_t)(q * 32) << 16;
    const float dsc = a.dsc[chunk * CM2 + row];
    const double* dsj = reinterpret_cast<const double*>(smem + S_DSJ);
    uint32_t d2_phase = 0, flush_left = (uint32_t)a.flush_tiles;
    int u_cur = 0, u_n1 = 0;
    if (SEG && my_tiles > 0) { u_cur = __ldg(a.tile_utt + seg_lo); u_n1 = my_tiles > 1 ? __ldg(a.tile_utt + seg_lo + 1) : -2; }
    for (int64_t it = 0; it < my_tiles; ++it) {
      const int b = (int)(it % NBUF);
      int u_n2 = -2;
      if (SEG && it + 2 < my_tiles) u_n2 = __ldg(a.tile_utt + seg_lo + it + 2);
      // 14 - lse2[b] of this warp's 32 frames (warp-uniform addresses: one broadcast line per load),
      // issued before the wait so the latency hides behind GEMM 1
      float cbv[32];
      {
        const int64_t tile = tile_of(it);
        const float4* cb4 = reinterpret_cast<const float4*>(a.cb + (size_t)tile * TF2) + h * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t = __ldg(cb4 + i);
          cbv[4 * i] = t.x; cbv[4 * i + 1] = t.y; cbv[4 * i + 2] = t.z; cbv[4 * i + 3] = t.w;
        }
      }
      mbar_wait(bar(SB_D1FULL + b), (uint32_t)((it / NBUF) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem + TM_BUF + 64u * b + 32u * h + lane_field;
      uint32_t r[32];
      tmem_ld32_nowait(taddr, r);
      tmem_ld_wait();
      uint32_t hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float p0 = ex2f(fmaf(__uint_as_float(r[2 * i]), dsc, cbv[2 * i]));
        const float p1 = ex2f(fmaf(__uint_as_float(r[2 * i + 1]), dsc, cbv[2 * i + 1]));
        split_h2(p0, p1, hi[i], lo[i]);
      }
      tmem_st16(taddr, hi);          // P' hi over the first 16 of this warp's 32 columns,
      tmem_st16(taddr + 16u, lo);    // P' lo over the last 16 (in place over D1)
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(bar(SB_PFULL + b));
      const int u_here = u_cur;
      bool flush_now;
      if (SEG) { flush_now = it + 1 == my_tiles || u_n1 != u_cur || flush_left == 1u; --flush_left; u_cur = u_n1; u_n1 = u_n2; }
      else flush_now = (--flush_left == 0u) || it + 1 == my_tiles;
      if (flush_now) {
        flush_left = (uint32_t)a.flush_tiles;
        // drain D2[mixture = row, j] into the fp64 statistics (64 columns per warp)
        mbar_wait(bar(SB_D2FULL), d2_phase);
        d2_phase ^= 1;
        tc_fence_after();
        const int m = chunk * CM2 + row;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          const int j0 = 64 * h + 32 * c;
          if (SEG && (j0 + 32 <= D || (j0 >= 2 * D && !(j0 <= K_ONE && K_ONE < j0 + 32)))) continue;   // S columns / padding only
          float s[32];
          tmem_ld32(tmem + TM_D2 + 64u * h + 32u * c + lane_field, s);
          if (m < a.M && (!SEG || u_here >= 0)) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int j = j0 + i;
              if (SEG) {
                // float32 rows of utterance u_here (red.global.add.f32: an utterance that straddles two CTAs' tile ranges
                // receives two partial sums); 2^(-14 - ea[j]) is an exact scaling
                if (j >= D && j < 2 * D) atomicAdd(a.uF + ((size_t)u_here * a.M + m) * D + (j - D), s[i] * (float)dsj[j]);
                else if (j == K_ONE) atomicAdd(a.uZ + (size_t)u_here * a.M + m, s[i] * (float)dsj[j]);
              } else {
                double* dst = nullptr;
                if (j < D) { if (a.want_second) dst = a.stats + a.M + (size_t)D * a.M + (size_t)j * a.M; }
                else if (j < 2 * D) dst = a.stats + a.M + (size_t)(j - D) * a.M;
                else if (j == K_ONE) dst = a.stats;
                if (dst != nullptr) atomicAdd(dst + m, (double)s[i] * dsj[j]);
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(bar(SB_D2EMPTY));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tmem_dealloc512(tmem);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
bool gmm_h_supported(const odin_gmm* g) {
  // The kernels pad the mixtures to their 256 / 128-wide tiles, so they run at any M; measured on 1 M frames
  // (tools/gmm_small_m_sweep.py) the padded tensor-core E-step takes 0.70-0.79 ms for M = 1 .. 256 where the fp32
  // CUDA-core kernels take 1.7 ms (M <= 64), 3.1 ms (128), 6.0 ms (256) -- so the early stages of the split-and-train
  // schedule (M = 1, 2, 4, ...) use it too.  ODIN_GMM_H_MIN_M restores a threshold for A/B runs.
  static const int min_m = [] { const char* e = getenv("ODIN_GMM_H_MIN_M"); return e ? atoi(e) : 1; }();
  return g->D % 4 == 0 && g->D <= hk::MAX_D && g->D >= 4 && g->M >= min_m;
}

static int64_t h_sub_batch() {
  const char* e = getenv("ODIN_H_SUB_BATCH");
  int64_t v = e ? atoll(e) : (int64_t)1 << 20;
  v = std::max<int64_t>(v, hk::TF1);
  return (v + hk::TF1 - 1) / hk::TF1 * hk::TF1;
}
static int h_flush_tiles() {
  const char* e = getenv("ODIN_H_FLUSH_TILES");
  int v = e ? atoi(e) : 256;
  return v < 1 ? 1 : v;
}

static int h_reserve(odin_gmm* g, int64_t sub, bool need_images) {
  const int64_t mpad = ceil_div<int64_t>(g->max_nmix, hk::CM1) * hk::CM1;
  if (g->d_hscale == nullptr) {
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_hscale, sizeof(HScale)));
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_hWplain, (size_t)mpad * 2 * 64 * sizeof(uint32_t)));
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_hWimg, (size_t)(mpad / hk::CM1) * 131072));
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_hdsc, (size_t)mpad * sizeof(float)));
  }
  if (need_images && sub > g->h_cap) {
    if (g->d_himgA) ODIN_CUDA_CHECK(cudaFree(g->d_himgA));
    if (g->d_himgT) ODIN_CUDA_CHECK(cudaFree(g->d_himgT));
    g->d_himgA = g->d_himgT = nullptr; g->h_cap = 0;
    const int64_t nsuper = sub / hk::TF1;
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_himgA, (size_t)nsuper * hk::SUPER_A_BYTES));
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_himgT, (size_t)nsuper * 2 * hk::TILE_T_BYTES));
    g->h_cap = sub;
  }
  if (sub > g->h_cb_cap) {
    if (g->d_hcb) ODIN_CUDA_CHECK(cudaFree(g->d_hcb));
    g->d_hcb = nullptr; g->h_cb_cap = 0;
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_hcb, (size_t)sub * sizeof(float)));
    g->h_cb_cap = sub;
  }
  const int nparts = (int)(mpad / hk::CM1) * 2;
  const int64_t need = sub * nparts;
  if (need > g->part_cap) {
    if (g->d_part) ODIN_CUDA_CHECK(cudaFree(g->d_part));
    g->d_part = nullptr; g->part_cap = 0;
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_part, need * sizeof(float2)));
    g->part_cap = need;
  }
  return ODIN_OK;
}

static int h_launch_range(const float* X, int64_t N, int D, HScale* sc, cudaStream_t st) {
  ODIN_CUDA_CHECK(cudaMemsetAsync(sc, 0, sizeof(unsigned) * 64, st));
  const int threads = 16 * (D >> 2);
  const int64_t total4 = N * (D >> 2);
  const int grid = (int)std::min<int64_t>(ceil_div<int64_t>(total4, (int64_t)threads * 8), (int64_t)sm_count() * 8);
  gmm_h_range_kernel<<<std::max(grid, 1), threads, 0, st>>>(X, N, D, sc);
  ODIN_LAUNCH_CHECK("gmm_h_range_kernel");
  gmm_h_scale_kernel<<<1, 128, 0, st>>>(D, sc);
  ODIN_LAUNCH_CHECK("gmm_h_scale_kernel");
  return ODIN_OK;
}

static int h_launch_prepare(odin_gmm* g, const HScale* sc, cudaStream_t st) {
  const int mpad = ceil_div(g->M, hk::CM1) * hk::CM1;
  gmm_h_prepare_kernel<<<ceil_div(mpad, 64), 64, 0, st>>>(g->d_mean, g->d_var, g->d_w, g->D, g->M, mpad, sc,
                                                          reinterpret_cast<uint32_t*>(g->d_hWplain),
                                                          reinterpret_cast<uint32_t*>(g->d_hWimg), g->d_hdsc);
  ODIN_LAUNCH_CHECK("gmm_h_prepare_kernel");
  ODIN_CUDA_CHECK(cudaFuncSetAttribute(gmm_h_lse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hk::L_SMEM));
  ODIN_CUDA_CHECK(cudaFuncSetAttribute(gmm_h_stats_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hk::S_SMEM));
  ODIN_CUDA_CHECK(cudaFuncSetAttribute(gmm_h_stats_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hk::S_SMEM));
  return ODIN_OK;
}

// pass 1 + combine + pass 2 over n frames whose operand images start at imgA / imgT
static int h_launch_passes(odin_gmm* g, const HScale* sc, const unsigned char* imgA, const unsigned char* imgT,
                           int64_t n, int64_t part_stride, const uint8_t* sad, int want_second, double* stats,
                           bool record, cudaStream_t st, const int* tile_utt = nullptr, float* uZ = nullptr,
                           float* uF = nullptr) {
  const int D = g->D, M = g->M;
  const int mpad = ceil_div(M, hk::CM1) * hk::CM1;
  const int nch1 = mpad / hk::CM1, nch2 = mpad / hk::CM2;
  const int64_t nsuper = ceil_div<int64_t>(n, hk::TF1);
  double* statL = stats ? stats + (stats_size(D, M) - 2) : nullptr;
  HArgs a{};
  a.tile_utt = tile_utt; a.uZ = uZ; a.uF = uF;
  a.N = n; a.D = D; a.M = M;
  a.imgA = imgA;
  a.imgT = imgT;
  a.Wplain = reinterpret_cast<const uint32_t*>(g->d_hWplain);
  a.Wimg = reinterpret_cast<const unsigned char*>(g->d_hWimg);
  a.dsc = g->d_hdsc;
  a.dscale = sc->dscale;
  a.cb = g->d_hcb;
  a.part = reinterpret_cast<float2*>(g->d_part);
  a.part_stride = part_stride;
  a.stats = stats;
  a.want_second = want_second;
  // segmented mode sums into float32 rows: drain every 32 tiles (2 048 frames) so that the fp32 TMEM accumulation of a
  // long utterance stays ~1e-5 relative (256 tiles measured 3.7e-5 against the fp64-drained route's 3.6e-6)
  a.flush_tiles = tile_utt != nullptr ? std::min(h_flush_tiles(), 32) : h_flush_tiles();
  {
    const int64_t splits = std::max<int64_t>(1, std::min<int64_t>(sm_count() / nch1, nsuper));
    gmm_h_lse_kernel<<<dim3(nch1, (unsigned)splits), hk::THREADS, hk::L_SMEM, st>>>(a);
    ODIN_LAUNCH_CHECK("gmm_h_lse_kernel");
  }
  const int64_t npad = nsuper * hk::TF1;
  gmm_h_combine_kernel<<<(unsigned)ceil_div<int64_t>(npad, 256), 256, 0, st>>>(
      reinterpret_cast<const float2*>(g->d_part), nch1 * 2, part_stride, n, npad, sad, g->d_hcb, statL);
  ODIN_LAUNCH_CHECK("gmm_h_combine_kernel");
  if (record) ODIN_CUDA_CHECK(cudaEventRecord(g->ev[1], st));
  {
    const int64_t splits = std::max<int64_t>(1, std::min<int64_t>(sm_count() / nch2, nsuper * 2));
    if (tile_utt != nullptr) gmm_h_stats_kernel<true><<<dim3(nch2, (unsigned)splits), hk::THREADS, hk::S_SMEM, st>>>(a);
    else gmm_h_stats_kernel<false><<<dim3(nch2, (unsigned)splits), hk::THREADS, hk::S_SMEM, st>>>(a);
    ODIN_LAUNCH_CHECK("gmm_h_stats_kernel");
  }
  if (record) {
    ODIN_CUDA_CHECK(cudaEventRecord(g->ev[2], st));
    g->last_frames = n;
  }
  return ODIN_OK;
}

// Whole E-step (both passes) over N frames; events ev[0..2] bracket (image + pass 1 | pass 2)
// of the LAST sub-batch, g->last_frames = its size.
int gmm_estep_h(odin_gmm* g, const float* X, const uint8_t* sad, int64_t N, int want_second, double* stats,
                cudaStream_t st) {
  if (N <= 0) return ODIN_OK;
  const int64_t sub = std::min<int64_t>(h_sub_batch(), ceil_div<int64_t>(N, hk::TF1) * hk::TF1);
  int rc = h_reserve(g, sub, true);
  if (rc) return rc;
  const int D = g->D;
  HScale* sc = reinterpret_cast<HScale*>(g->d_hscale);
  if ((rc = h_launch_range(X, N, D, sc, st))) return rc;   // data range -> exact power-of-two scales
  if ((rc = h_launch_prepare(g, sc, st))) return rc;       // scaled model images
  for (int64_t s0 = 0; s0 < N; s0 += sub) {
    const int64_t n = std::min<int64_t>(sub, N - s0);
    const int64_t nsuper = ceil_div<int64_t>(n, hk::TF1);
    const bool last = s0 + sub >= N;
    if (last) ODIN_CUDA_CHECK(cudaEventRecord(g->ev[0], st));
    gmm_h_image_kernel<<<(unsigned)nsuper, 256, 0, st>>>(X + s0 * D, n, D, sc,
                                                          reinterpret_cast<unsigned char*>(g->d_himgA),
                                                          reinterpret_cast<unsigned char*>(g->d_himgT));
    ODIN_LAUNCH_CHECK("gmm_h_image_kernel");
    rc = h_launch_passes(g, sc, reinterpret_cast<const unsigned char*>(g->d_himgA),
                         reinterpret_cast<const unsigned char*>(g->d_himgT), n, sub, sad ? sad + s0 : nullptr,
                         want_second, stats, last, st);
    if (rc) return rc;
  }
  return ODIN_OK;
}

// ---- per-utterance statistics (gmm_tmat.py:708-767, 769-913) on the tensor-core E-step ----
// acc holds, per utterance of a group, the packed statistics the E-step accumulates (Z [M] | F [D, M] | ...);
// Z and the centred first-order statistics F-hat[m * D + d] = F[d, m] - mean[d, m] Z[m] go out as float32 rows.
__global__ void __launch_bounds__(256) gmm_h_centre_kernel(const double* __restrict__ acc, int64_t stride, int D, int M,
                                                           const float* __restrict__ mean, float* __restrict__ Z,
                                                           float* __restrict__ Fhat) {
  const double* a = acc + (int64_t)blockIdx.x * stride;
  float* z = Z + (int64_t)blockIdx.x * M;
  float* f = Fhat + (int64_t)blockIdx.x * M * D;
  for (int m = threadIdx.x; m < M; m += 256) z[m] = (float)a[m];
  for (int i = threadIdx.x; i < M * D; i += 256) {
    const int m = i / D, d = i - m * D;
    f[i] = (float)(a[M + (int64_t)d * M + m] - (double)mean[(int64_t)d * M + m] * a[m]);
  }
}

int gmm_utt_stats_h(odin_gmm* g, const float* X, const uint8_t* sad, const int64_t* h_off, int n_utt, float* d_Z,
                    float* d_Fhat, cudaStream_t st) {
  const int D = g->D, M = g->M;
  const int64_t base = h_off[0], N = h_off[n_utt] - base;
  int64_t maxlen = 0;
  for (int u = 0; u < n_utt; ++u) maxlen = std::max(maxlen, h_off[u + 1] - h_off[u]);
  if (N <= 0) {
    ODIN_CUDA_CHECK(cudaMemsetAsync(d_Z, 0, sizeof(float) * (size_t)n_utt * M, st));
    ODIN_CUDA_CHECK(cudaMemsetAsync(d_Fhat, 0, sizeof(float) * (size_t)n_utt * M * D, st));
    return ODIN_OK;
  }
  const int64_t sub = std::min<int64_t>(h_sub_batch(), ceil_div<int64_t>(maxlen, hk::TF1) * hk::TF1);
  int rc = h_reserve(g, sub, true);
  if (rc) return rc;
  HScale* sc = reinterpret_cast<HScale*>(g->d_hscale);
  // one data range (-> exact power-of-two scales) and one set of scaled model images for the whole batch
  if ((rc = h_launch_range(X, N, D, sc, st))) return rc;
  if ((rc = h_launch_prepare(g, sc, st))) return rc;
  const int64_t SD = stats_size(D, M);
  const int G = 32;
  if (g->utt_acc_cap < G * SD) {
    cudaFree(g->d_utt_acc);
    g->d_utt_acc = nullptr; g->utt_acc_cap = 0;
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_utt_acc, sizeof(double) * (size_t)(G * SD)));
    g->utt_acc_cap = G * SD;
  }
  for (int u0 = 0; u0 < n_utt; u0 += G) {
    const int cnt = std::min(G, n_utt - u0);
    ODIN_CUDA_CHECK(cudaMemsetAsync(g->d_utt_acc, 0, sizeof(double) * (size_t)(cnt * SD), st));
    for (int j = 0; j < cnt; ++j) {
      const int64_t lo = h_off[u0 + j] - base, n_u = h_off[u0 + j + 1] - h_off[u0 + j];
      for (int64_t s0 = 0; s0 < n_u; s0 += sub) {
        const int64_t n = std::min<int64_t>(sub, n_u - s0);
        const int64_t nsuper = ceil_div<int64_t>(n, hk::TF1);
        gmm_h_image_kernel<<<(unsigned)nsuper, 256, 0, st>>>(X + (lo + s0) * D, n, D, sc,
                                                              reinterpret_cast<unsigned char*>(g->d_himgA),
                                                              reinterpret_cast<unsigned char*>(g->d_himgT));
        ODIN_LAUNCH_CHECK("gmm_h_image_kernel");
        rc = h_launch_passes(g, sc, reinterpret_cast<const unsigned char*>(g->d_himgA),
                             reinterpret_cast<const unsigned char*>(g->d_himgT), n, sub, sad ? sad + lo + s0 : nullptr, 0,
                             g->d_utt_acc + (int64_t)j * SD, false, st);
        if (rc) return rc;
      }
    }
    gmm_h_centre_kernel<<<cnt, 256, 0, st>>>(g->d_utt_acc, SD, D, M, g->d_mean, d_Z + (int64_t)u0 * M,
                                             d_Fhat + (int64_t)u0 * M * D);
    ODIN_LAUNCH_CHECK("gmm_h_centre_kernel");
  }
  return ODIN_OK;
}

// ---- per-utterance statistics of MANY SHORT utterances in one tensor-core launch sequence -----------------------------
// The utterances of a group are laid out back to back, each padded to whole 64-frame tiles (zero rows, masked out of the
// posteriors through cb = -1e30 like SAD-rejected frames), so that a tile of pass 2 belongs to ONE utterance; pass 1 and the
// images do not care about the boundaries.  gmm_h_stats_kernel<true> then drains its accumulator at every utterance change.
__global__ void __launch_bounds__(256) gmm_h_segpad_kernel(const float* __restrict__ X, const uint8_t* __restrict__ sad, int D,
                                                           const int* __restrict__ tile_utt, const int64_t* __restrict__ off,
                                                           const int64_t* __restrict__ poff, int u0, int64_t p0,
                                                           float* __restrict__ Xp, uint8_t* __restrict__ mask) {
  const int64_t tile = blockIdx.x;
  const int u = tile_utt[tile];
  const int d4 = D >> 2;
  float4* dst = reinterpret_cast<float4*>(Xp + tile * hk::TF2 * D);
  int64_t src0 = 0, live = 0;
  if (u >= 0) {
    const int64_t li0 = p0 + tile * hk::TF2 - poff[u0 + u];   // index of the tile's first frame inside its utterance
    src0 = off[u0 + u] + li0;
    live = off[u0 + u + 1] - src0;                             // frames of the utterance from there on (may exceed the tile)
  }
  const float4* src = reinterpret_cast<const float4*>(X + src0 * D);
  for (int i = threadIdx.x; i < hk::TF2 * d4; i += 256) {
    const int r = i / d4;
    dst[i] = r < live ? __ldg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (threadIdx.x < hk::TF2) {
    const int r = threadIdx.x;
    mask[tile * hk::TF2 + r] = (r < live && (sad == nullptr || sad[src0 + r] != 0)) ? 1 : 0;
  }
}

// F-hat[u, m * D + d] = F[u, m, d] - mean[d, m] Z[u, m], in place
__global__ void __launch_bounds__(256) gmm_h_segcentre_kernel(const float* __restrict__ Z, const float* __restrict__ mean, int D,
                                                              int M, int64_t total, float* __restrict__ F) {
  const int64_t md = (int64_t)M * D;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const int64_t u = i / md, r = i - u * md;
    const int m = (int)(r / D), d = (int)(r - (int64_t)m * D);
    F[i] = F[i] - mean[(int64_t)d * M + m] * Z[u * M + m];
  }
}

int gmm_utt_stats_hseg(odin_gmm* g, const float* X, const uint8_t* sad, const int64_t* h_off, int n_utt, float* d_Z,
                       float* d_Fhat, cudaStream_t st) {
  const int D = g->D, M = g->M;
  const int64_t base = h_off[0], N = h_off[n_utt] - base;
  ODIN_CUDA_CHECK(cudaMemsetAsync(d_Z, 0, sizeof(float) * (size_t)n_utt * M, st));
  ODIN_CUDA_CHECK(cudaMemsetAsync(d_Fhat, 0, sizeof(float) * (size_t)n_utt * M * D, st));
  if (N <= 0) return ODIN_OK;
  // padded offsets and groups of whole utterances of at most `cap` padded frames
  std::vector<int64_t> off(n_utt + 1), poff(n_utt + 1);
  int64_t maxpad = 0;
  poff[0] = 0;
  for (int u = 0; u <= n_utt; ++u) off[u] = h_off[u] - base;
  for (int u = 0; u < n_utt; ++u) {
    const int64_t pl = ceil_div<int64_t>(off[u + 1] - off[u], hk::TF2) * hk::TF2;
    poff[u + 1] = poff[u] + pl;
    maxpad = std::max(maxpad, pl);
  }
  const int64_t cap = std::max<int64_t>(h_sub_batch(), ceil_div<int64_t>(maxpad, hk::TF1) * hk::TF1);
  const int64_t sub = std::min<int64_t>(cap, ceil_div<int64_t>(poff[n_utt], hk::TF1) * hk::TF1);
  int rc = h_reserve(g, sub, true);
  if (rc) return rc;
  const int64_t tiles_cap = sub / hk::TF2;
  if (g->seg_cap < sub || g->seg_utt_cap < n_utt + 1) {
    cudaFree(g->d_segX); cudaFree(g->d_segmask); cudaFree(g->d_segtile); cudaFree(g->d_segoff);
    g->d_segX = nullptr; g->d_segmask = nullptr; g->d_segtile = nullptr; g->d_segoff = nullptr;
    g->seg_cap = 0; g->seg_utt_cap = 0;
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_segX, sizeof(float) * (size_t)sub * D));
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_segmask, (size_t)sub));
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_segtile, sizeof(int) * (size_t)tiles_cap));
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_segoff, sizeof(int64_t) * 2 * (size_t)(n_utt + 1)));
    g->seg_cap = sub; g->seg_utt_cap = n_utt + 1;
  }
  int64_t* d_off = g->d_segoff;
  int64_t* d_poff = d_off + (n_utt + 1);
  ODIN_CUDA_CHECK(cudaMemcpyAsync(d_off, off.data(), sizeof(int64_t) * (n_utt + 1), cudaMemcpyHostToDevice, st));
  ODIN_CUDA_CHECK(cudaMemcpyAsync(d_poff, poff.data(), sizeof(int64_t) * (n_utt + 1), cudaMemcpyHostToDevice, st));
  HScale* sc = reinterpret_cast<HScale*>(g->d_hscale);
  // one data range (-> exact power-of-two scales) and one set of scaled model images for the whole batch
  if ((rc = h_launch_range(X, N, D, sc, st))) return rc;
  if ((rc = h_launch_prepare(g, sc, st))) return rc;
  std::vector<int> tile_utt;
  for (int u0 = 0; u0 < n_utt;) {
    int u1 = u0;
    while (u1 < n_utt && poff[u1 + 1] - poff[u0] <= sub) ++u1;   // (a single utterance always fits: sub >= maxpad)
    const int64_t np = poff[u1] - poff[u0];                        // padded frames of the group: a multiple of 64
    const int64_t nsuper = ceil_div<int64_t>(np, hk::TF1);
    const int64_t nt = nsuper * 2;
    tile_utt.assign((size_t)nt, -1);
    for (int u = u0; u < u1; ++u)
      for (int64_t t = (poff[u] - poff[u0]) / hk::TF2; t < (poff[u + 1] - poff[u0]) / hk::TF2; ++t) tile_utt[(size_t)t] = u - u0;
    ODIN_CUDA_CHECK(cudaMemcpyAsync(g->d_segtile, tile_utt.data(), sizeof(int) * (size_t)nt, cudaMemcpyHostToDevice, st));
    gmm_h_segpad_kernel<<<(unsigned)nt, 256, 0, st>>>(X, sad, D, g->d_segtile, d_off, d_poff, u0, poff[u0], g->d_segX,
                                                     g->d_segmask);
    ODIN_LAUNCH_CHECK("gmm_h_segpad_kernel");
    const int64_t n = nsuper * hk::TF1;
    gmm_h_image_kernel<<<(unsigned)nsuper, 256, 0, st>>>(g->d_segX, n, D, sc, reinterpret_cast<unsigned char*>(g->d_himgA),
                                                          reinterpret_cast<unsigned char*>(g->d_himgT));
    ODIN_LAUNCH_CHECK("gmm_h_image_kernel");
    rc = h_launch_passes(g, sc, reinterpret_cast<const unsigned char*>(g->d_himgA),
                         reinterpret_cast<const unsigned char*>(g->d_himgT), n, sub, g->d_segmask, 0, nullptr, false, st,
                         g->d_segtile, d_Z + (int64_t)u0 * M, d_Fhat + (int64_t)u0 * M * D);
    if (rc) return rc;
    u0 = u1;
  }
  const int64_t total = (int64_t)n_utt * M * D;
  gmm_h_segcentre_kernel<<<(unsigned)std::min<int64_t>(ceil_div<int64_t>(total, 256 * 8), (int64_t)sm_count() * 16), 256, 0, st>>>(
      d_Z, g->d_mean, D, M, total, d_Fhat);
  ODIN_LAUNCH_CHECK("gmm_h_segcentre_kernel");
  ODIN_CUDA_CHECK(cudaStreamSynchronize(st));   // the pageable staging vectors of this call go out of scope
  return ODIN_OK;
}

// ---- prepared frames: the operand images depend on the data only, so a resident frame matrix
// that is visited once per EM iteration gets them built ONCE (1 KB per frame)
struct GmmFrames {
  int D = 0, device = 0;
  int64_t N = 0;
  HScale* scale = nullptr;
  unsigned char* imgA = nullptr;
  unsigned char* imgT = nullptr;
};

void gmm_frames_destroy(void* p) {
  GmmFrames* f = reinterpret_cast<GmmFrames*>(p);
  if (!f) return;
  cudaFree(f->scale); cudaFree(f->imgA); cudaFree(f->imgT);
  delete f;
}

int gmm_frames_create(odin_gmm* g, const float* X, int64_t N, void** out, cudaStream_t st) {
  *out = nullptr;
  if (g->D % 4 != 0 || g->D > hk::MAX_D || g->D < 4)
    return set_error(ODIN_EINVAL, "prepared frames need D %% 4 == 0 and D <= %d", hk::MAX_D);
  GmmFrames* f = new (std::nothrow) GmmFrames();
  if (!f) return set_error(ODIN_ENOMEM, "out of host memory");
  f->D = g->D; f->N = N; f->device = g->device;
  const int64_t nsuper = ceil_div<int64_t>(N, hk::TF1);
  cudaError_t e;
  if ((e = cudaMalloc(&f->scale, sizeof(HScale))) != cudaSuccess ||
      (e = cudaMalloc(&f->imgA, (size_t)nsuper * hk::SUPER_A_BYTES)) != cudaSuccess ||
      (e = cudaMalloc(&f->imgT, (size_t)nsuper * 2 * hk::TILE_T_BYTES)) != cudaSuccess) {
    cudaGetLastError();
    gmm_frames_destroy(f);
    return set_error(ODIN_ENOMEM, "prepared frames: cudaMalloc of %lld MB failed: %s",
                     (long long)(nsuper * 131072 >> 20), cudaGetErrorString(e));
  }
  int rc = h_launch_range(X, N, g->D, f->scale, st);
  if (rc) { gmm_frames_destroy(f); return rc; }
  const int64_t per = (int64_t)1 << 22;   // super-tiles per launch (grid.x limit is not a concern; keeps launches short)
  for (int64_t s0 = 0; s0 < nsuper; s0 += per) {
    const int64_t ns = std::min<int64_t>(per, nsuper - s0);
    gmm_h_image_kernel<<<(unsigned)ns, 256, 0, st>>>(X + s0 * hk::TF1 * g->D, N - s0 * hk::TF1, g->D, f->scale,
                                                      f->imgA + (size_t)s0 * hk::SUPER_A_BYTES,
                                                      f->imgT + (size_t)s0 * 2 * hk::TILE_T_BYTES);
    ODIN_LAUNCH_CHECK("gmm_h_image_kernel");
  }
  *out = f;
  return ODIN_OK;
}

int gmm_estep_frames(odin_gmm* g, const void* pf, const uint8_t* sad, int want_second, double* stats,
                     cudaStream_t st) {
  const GmmFrames* f = reinterpret_cast<const GmmFrames*>(pf);
  if (f->D != g->D) return set_error(ODIN_EINVAL, "prepared frames have D=%d, model has D=%d", f->D, g->D);
  if (!gmm_h_supported(g)) return set_error(ODIN_EINVAL, "prepared frames need the 3xFP16 path (D a multiple of 4, M above ODIN_GMM_H_MIN_M)");
  const int64_t N = f->N;
  if (N <= 0) return ODIN_OK;
  const int64_t sub = std::min<int64_t>(h_sub_batch(), ceil_div<int64_t>(N, hk::TF1) * hk::TF1);
  int rc = h_reserve(g, sub, false);
  if (rc) return rc;
  if ((rc = h_launch_prepare(g, f->scale, st))) return rc;
  for (int64_t s0 = 0; s0 < N; s0 += sub) {
    const int64_t n = std::min<int64_t>(sub, N - s0);
    const bool last = s0 + sub >= N;
    if (last) ODIN_CUDA_CHECK(cudaEventRecord(g->ev[0], st));
    rc = h_launch_passes(g, f->scale, f->imgA + (size_t)(s0 / hk::TF1) * hk::SUPER_A_BYTES,
                         f->imgT + (size_t)(s0 / hk::TF2) * hk::TILE_T_BYTES, n, sub, sad ? sad + s0 : nullptr,
                         want_second, stats, last, st);
    if (rc) return rc;
  }
  return ODIN_OK;
}

void gmm_h_free(odin_gmm* g) {
  cudaFree(g->d_hscale); cudaFree(g->d_hWplain); cudaFree(g->d_hWimg); cudaFree(g->d_hdsc);
  cudaFree(g->d_himgA); cudaFree(g->d_himgT); cudaFree(g->d_hcb);
}

}  // namespace odin
