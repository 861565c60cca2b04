// GMM-UBM Baum-Welch: fp32 CUDA-core kernels (correctness template, generic
// shapes, and the fallback for shapes the tcgen05 path does not take).
//
// Reference arithmetic: odin/ml/gmm_tmat.py:1012-1041 (_fast_expectation, numpy
// branch), :493-504 (cached constants), :1233-1276 (maximization), :1308-1338
// (gmm_mixup), :708-767 (transform).
//
//   logprob[b,m] = sum_d x^2[b,d] * (-0.5 prec[d,m]) + x[b,d] * (mu prec)[d,m] + cst[m]
//   cst[m]       = -0.5 * (C[m] + D log 2pi)
//   lse[b]       = logsumexp_m logprob[b,m]              (kernel 1: gmm_lse_kernel)
//   post[b,m]    = exp(logprob[b,m] - lse[b])
//   [Z | F | S][m, j] = sum_b post[b,m] * [1, x, x^2][b,j]   (kernel 2: gmm_stats_kernel)
//
// Kernel 2 keeps the [64 mixtures x (2D+1)] accumulator of one mixture chunk in
// registers while it streams frames, recomputing logprob for its chunk only, so
// posteriors never touch HBM.  Per-CTA fp32 partial sums are flushed into the
// caller's fp64 statistics with red.global.add.f64 every kFlushTiles tiles.
#include <math.h>

#include "gmm.cuh"

namespace odin {

constexpr int TF = 64;        // frames per tile
constexpr int TM = 64;        // mixtures per chunk
constexpr int NT = 256;       // threads per CTA
constexpr int AS_LD = 68;     // padded leading dimension of As[k][f]
constexpr int kFlushTiles = 512;

// ---------------------------------------------------------------------------
// cached constants (gmm_tmat.py:493-504), computed in fp64 from the fp32 model
// ---------------------------------------------------------------------------
__global__ void gmm_prepare_kernel(const float* __restrict__ mean, const float* __restrict__ var,
                                   const float* __restrict__ w, int D, int M, int Mpad,
                                   float* __restrict__ Wk, float* __restrict__ cst) {
  int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= Mpad) return;
  if (m >= M) {
    for (int d = 0; d < 2 * D; ++d) Wk[(size_t)d * Mpad + m] = 0.f;
    cst[m] = -1e30f;
    return;
  }
  double C = 0.0;
  for (int d = 0; d < D; ++d) {
    double v = (double)var[(size_t)d * M + m] + ODIN_GMM_EPS;
    double p = 1.0 / v;
    double mu = (double)mean[(size_t)d * M + m];
    C += mu * mu * p + log(v);
    Wk[(size_t)d * Mpad + m] = (float)(-0.5 * p);
    Wk[(size_t)(D + d) * Mpad + m] = (float)(mu * p);
  }
  C -= 2.0 * log((double)w[m] + ODIN_GMM_EPS);
  cst[m] = (float)(-0.5 * (C + (double)D * 1.8378770664093454835606594728112));  // log(2 pi)
}

// ---------------------------------------------------------------------------
// tile helpers
// ---------------------------------------------------------------------------
// As[k][f]: k < D -> x^2, D <= k < 2D -> x.  Frames >= nvalid are zero-filled.
__device__ __forceinline__ void load_tile_As(const float* __restrict__ X, int64_t f0, int nvalid, int D,
                                             float* __restrict__ As) {
  const int total = TF * D;
  for (int idx = threadIdx.x; idx < total; idx += NT) {
    int f = idx / D, d = idx - f * D;
    float x = (f < nvalid) ? __ldg(X + (f0 + f) * (int64_t)D + d) : 0.f;
    As[d * AS_LD + f] = x * x;
    As[(D + d) * AS_LD + f] = x;
  }
}

__device__ __forceinline__ void load_chunk_Ws(const float* __restrict__ Wk, int Mpad, int c, int K2,
                                              float* __restrict__ Ws) {
  const int total = K2 * (TM / 4);
  for (int idx = threadIdx.x; idx < total; idx += NT) {
    int k = idx / (TM / 4), q = idx - k * (TM / 4);
    reinterpret_cast<float4*>(Ws)[k * (TM / 4) + q] =
        __ldg(reinterpret_cast<const float4*>(Wk + (size_t)k * Mpad + c * TM) + q);
  }
}

// acc[i][j] = sum_k As[k][4*ty+i] * Ws[k][4*tx+j]
__device__ __forceinline__ void lp_micro_tile(const float* __restrict__ As, const float* __restrict__ Ws,
                                              int K2, int ty, int tx, float (&acc)[4][4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 4
  for (int k = 0; k < K2; ++k) {
    float4 a = *reinterpret_cast<const float4*>(As + k * AS_LD + 4 * ty);
    float4 w = *reinterpret_cast<const float4*>(Ws + k * TM + 4 * tx);
    float av[4] = {a.x, a.y, a.z, a.w};
    float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
  }
}

// ---------------------------------------------------------------------------
// kernel 1: per-frame log-sum-exp (+ sum of LLK and frame count into stats)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(NT)
gmm_lse_kernel(const float* __restrict__ X, const uint8_t* __restrict__ sad, int64_t N, int D, int Mpad,
               int nchunks,
               const float* __restrict__ Wk, const float* __restrict__ cst, float* __restrict__ lse,
               double* __restrict__ stat_L /* [2]: sum LLK, nframes; nullable */) {
  extern __shared__ __align__(16) float smem[];
  const int K2 = 2 * D;
  float* As = smem;
  float* Ws = As + K2 * AS_LD;
  __shared__ double red[NT / 32][2];

  const int64_t f0 = (int64_t)blockIdx.x * TF;
  const int nvalid = (int)min((int64_t)TF, N - f0);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  load_tile_As(X, f0, nvalid, D, As);

  float m_run[4], s_run[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { m_run[i] = -1e30f; s_run[i] = 0.f; }

  for (int c = 0; c < nchunks; ++c) {
    __syncthreads();
    load_chunk_Ws(Wk, Mpad, c, K2, Ws);
    __syncthreads();
    float acc[4][4];
    lp_micro_tile(As, Ws, K2, ty, tx, acc);
    float4 c4 = __ldg(reinterpret_cast<const float4*>(cst + c * TM) + tx);
    float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v[4];
      float mx = -1e30f;
#pragma unroll
      for (int j = 0; j < 4; ++j) { v[j] = acc[i][j] + cv[j]; mx = fmaxf(mx, v[j]); }
      float mnew = fmaxf(m_run[i], mx);
      float s = s_run[i] * __expf(m_run[i] - mnew);
#pragma unroll
      for (int j = 0; j < 4; ++j) s += __expf(v[j] - mnew);
      s_run[i] = s;
      m_run[i] = mnew;
    }
  }
  // combine the 16 lanes (tx) that share a frame
  double lsum = 0.0, lcnt = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float mstar = m_run[i];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) mstar = fmaxf(mstar, __shfl_xor_sync(0xffffffffu, mstar, o));
    float s = s_run[i] * __expf(m_run[i] - mstar);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    float l = mstar + logf(s);
    int f = 4 * ty + i;
    if (tx == 0 && f < nvalid) {
      lse[f0 + f] = l;
      if (sad == nullptr || sad[f0 + f] != 0) { lsum += (double)l; lcnt += 1.0; }
    }
  }
  if (stat_L != nullptr) {
    lsum = warp_sum(lsum);
    lcnt = warp_sum(lcnt);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[warp][0] = lsum; red[warp][1] = lcnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double a = 0, b = 0;
      for (int i = 0; i < NT / 32; ++i) { a += red[i][0]; b += red[i][1]; }
      if (b > 0) { atomicAdd(stat_L, a); atomicAdd(stat_L + 1, b); }
    }
  }
}

// ---------------------------------------------------------------------------
// kernel 2: statistics.  grid.x = mixture chunk, grid.y = frame split (MODE 0)
// or utterance (MODE 1).
// ---------------------------------------------------------------------------
struct StatsArgs {
  const float* X;
  const uint8_t* sad;
  const float* lse;
  int64_t N;
  int D, M, Mpad;
  const float* Wk;
  const float* cst;
  // MODE 0
  double* stats;
  int want_second;
  int tiles_per_split;
  // MODE 1
  const int64_t* off;
  const float* mean;  // [D, M]
  float* Z;           // [n_utt, M]
  float* Fhat;        // [n_utt, M*D]
};

template <int JT, int MODE>
__global__ void __launch_bounds__(NT) gmm_stats_kernel(StatsArgs a) {
  extern __shared__ __align__(16) float smem[];
  constexpr int JP = 16 * JT;
  const int D = a.D, K2 = 2 * a.D;
  float* As = smem;                 // [K2][AS_LD]
  float* Ws = As + K2 * AS_LD;      // [K2][TM]
  float* At = Ws + K2 * TM;         // [TF][JP]  j: 0 -> 1, 1..D -> x, D+1..2D -> x^2
  float* Ps = At + TF * JP;         // [TF][TM]

  const int c = blockIdx.x;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  int64_t fb, fe;
  if (MODE == 0) {
    fb = (int64_t)blockIdx.y * a.tiles_per_split * TF;
    fe = min(a.N, fb + (int64_t)a.tiles_per_split * TF);
  } else {
    fb = a.off[blockIdx.y];
    fe = a.off[blockIdx.y + 1];
  }

  load_chunk_Ws(a.Wk, a.Mpad, c, K2, Ws);
  for (int idx = threadIdx.x; idx < TF * JP; idx += NT) At[idx] = 0.f;
  float4 c4 = __ldg(reinterpret_cast<const float4*>(a.cst + c * TM) + tx);
  const float cv[4] = {c4.x, c4.y, c4.z, c4.w};

  float acc2[4][JT];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < JT; ++j) acc2[i][j] = 0.f;

  auto flush_global = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int mi = c * TM + 4 * ty + i;
      if (mi >= a.M) continue;
#pragma unroll
      for (int jj = 0; jj < JT; ++jj) {
        int j = JT * tx + jj;
        float v = acc2[i][jj];
        if (j == 0) {
          atomicAdd(a.stats + mi, (double)v);
        } else if (j <= D) {
          atomicAdd(a.stats + a.M + (size_t)(j - 1) * a.M + mi, (double)v);
        } else if (j <= 2 * D && a.want_second) {
          atomicAdd(a.stats + a.M + (size_t)D * a.M + (size_t)(j - 1 - D) * a.M + mi, (double)v);
        }
        acc2[i][jj] = 0.f;
      }
    }
  };

  int since_flush = 0;
  for (int64_t t0 = fb; t0 < fe; t0 += TF) {
    const int nvalid = (int)min((int64_t)TF, fe - t0);
    __syncthreads();  // previous tile fully consumed (also orders the initial Ws/At fill)
    {
      const int total = TF * D;
      for (int idx = threadIdx.x; idx < total; idx += NT) {
        int f = idx / D, d = idx - f * D;
        float x = (f < nvalid) ? __ldg(a.X + (t0 + f) * (int64_t)D + d) : 0.f;
        float x2 = x * x;
        As[d * AS_LD + f] = x2;
        As[(D + d) * AS_LD + f] = x;
        At[f * JP + 1 + d] = x;
        At[f * JP + 1 + D + d] = x2;
      }
      for (int f = threadIdx.x; f < TF; f += NT) At[f * JP] = (f < nvalid) ? 1.f : 0.f;
    }
    __syncthreads();
    // step A: posteriors of this chunk for the tile
    {
      float acc[4][4];
      lp_micro_tile(As, Ws, K2, ty, tx, acc);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        int f = 4 * ty + i;
        bool on = f < nvalid;
        float l = 0.f;
        if (on) {
          l = __ldg(a.lse + t0 + f);
          if (a.sad != nullptr && a.sad[t0 + f] == 0) on = false;
        }
        float4 p;
        p.x = on ? __expf(acc[i][0] + cv[0] - l) : 0.f;
        p.y = on ? __expf(acc[i][1] + cv[1] - l) : 0.f;
        p.z = on ? __expf(acc[i][2] + cv[2] - l) : 0.f;
        p.w = on ? __expf(acc[i][3] + cv[3] - l) : 0.f;
        *reinterpret_cast<float4*>(Ps + f * TM + 4 * tx) = p;
      }
    }
    __syncthreads();
    // step B: acc2[m][j] += sum_f Ps[f][m] * At[f][j]   (ty -> mixtures, tx -> features)
#pragma unroll 2
    for (int f = 0; f < TF; ++f) {
      float4 p = *reinterpret_cast<const float4*>(Ps + f * TM + 4 * ty);
      const float pv[4] = {p.x, p.y, p.z, p.w};
      float av[JT];
#pragma unroll
      for (int q = 0; q < JT / 4; ++q) {
        float4 t = *reinterpret_cast<const float4*>(At + f * JP + JT * tx + 4 * q);
        av[4 * q] = t.x; av[4 * q + 1] = t.y; av[4 * q + 2] = t.z; av[4 * q + 3] = t.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int jj = 0; jj < JT; ++jj) acc2[i][jj] = fmaf(pv[i], av[jj], acc2[i][jj]);
    }
    if (MODE == 0 && ++since_flush == kFlushTiles) { flush_global(); since_flush = 0; }
  }

  if (MODE == 0) {
    flush_global();
  } else {
    // per-utterance centred statistics: Z[u, m], Fhat[u, m*D + d] = F[d,m] - mean[d,m] Z[m]
    __syncthreads();
    float* Zs = Ps;  // [TM]
    if (tx == 0) {
#pragma unroll
      for (int i = 0; i < 4; ++i) Zs[4 * ty + i] = acc2[i][0];
    }
    __syncthreads();
    const int u = blockIdx.y;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int mi = c * TM + 4 * ty + i;
      if (mi >= a.M) continue;
      float z = Zs[4 * ty + i];
#pragma unroll
      for (int jj = 0; jj < JT; ++jj) {
        int j = JT * tx + jj;
        if (j == 0) {
          a.Z[(size_t)u * a.M + mi] = z;
        } else if (j <= D) {
          int d = j - 1;
          a.Fhat[(size_t)u * a.M * D + (size_t)mi * D + d] =
              acc2[i][jj] - __ldg(a.mean + (size_t)d * a.M + mi) * z;
        }
      }
    }
  }
}

// posteriors to HBM (API parity with GMM.postprob; not on the training path)
__global__ void __launch_bounds__(NT)
gmm_post_kernel(const float* __restrict__ X, int64_t N, int D, int M, int Mpad, const float* __restrict__ Wk,
                const float* __restrict__ cst, const float* __restrict__ lse, float* __restrict__ post,
                float* __restrict__ logp) {
  extern __shared__ __align__(16) float smem[];
  const int K2 = 2 * D;
  float* As = smem;
  float* Ws = As + K2 * AS_LD;
  const int64_t f0 = (int64_t)blockIdx.x * TF;
  const int c = blockIdx.y;
  const int nvalid = (int)min((int64_t)TF, N - f0);
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
  load_tile_As(X, f0, nvalid, D, As);
  load_chunk_Ws(Wk, Mpad, c, K2, Ws);
  __syncthreads();
  float acc[4][4];
  lp_micro_tile(As, Ws, K2, ty, tx, acc);
  float4 c4 = __ldg(reinterpret_cast<const float4*>(cst + c * TM) + tx);
  const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int f = 4 * ty + i;
    if (f >= nvalid) continue;
    float l = lse[f0 + f];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m = c * TM + 4 * tx + j;
      if (m < M) {
        const float lp = acc[i][j] + cv[j];
        if (post != nullptr) post[(f0 + f) * (int64_t)M + m] = __expf(lp - l);
        if (logp != nullptr) logp[(f0 + f) * (int64_t)M + m] = lp;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// M-step (gmm_tmat.py:1233-1276) and mixture split (:1308-1338): one CTA, fp64
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
gmm_mstep_kernel(const double* __restrict__ stats, int D, int M, int allow_rollback, float* __restrict__ mean,
                 float* __restrict__ var, float* __restrict__ w, float* __restrict__ tmp /* [2*D*M + M] */,
                 int* __restrict__ rolled_back) {
  __shared__ double red[32];
  __shared__ int neg_flag;
  __shared__ double zsum;
  if (threadIdx.x == 0) neg_flag = 0;
  double s = 0.0;
  for (int m = threadIdx.x; m < M; m += blockDim.x) s += stats[m];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
    zsum = t;
  }
  __syncthreads();
  const double* Fs = stats + M;
  const double* Ss = stats + M + (size_t)D * M;
  float* tmean = tmp;
  float* tvar = tmp + (size_t)D * M;
  float* tw = tmp + 2 * (size_t)D * M;
  int neg = 0;
  for (int idx = threadIdx.x; idx < D * M; idx += blockDim.x) {
    int m = idx % M;
    double iN = 1.0 / (stats[m] + ODIN_GMM_EPS);
    double mu = Fs[idx] * iN;
    double v = Ss[idx] * iN - mu * mu;
    if (v < 0.0) neg = 1;
    tmean[idx] = (float)mu;
    tvar[idx] = (float)v;
  }
  for (int m = threadIdx.x; m < M; m += blockDim.x) tw[m] = (float)(stats[m] / zsum);
  if (neg) atomicOr(&neg_flag, 1);
  __syncthreads();
  const int flag = neg_flag;
  if (threadIdx.x == 0 && rolled_back != nullptr) *rolled_back = flag;
  if (flag && allow_rollback) return;  // keep the previous model (gmm_tmat.py:1262-1266)
  for (int idx = threadIdx.x; idx < D * M; idx += blockDim.x) {
    mean[idx] = tmean[idx];
    var[idx] = flag ? fmaxf(tvar[idx], 0.f) : tvar[idx];
  }
  for (int m = threadIdx.x; m < M; m += blockDim.x) w[m] = tw[m];
}

__global__ void gmm_mixup_kernel(const float* __restrict__ omean, const float* __restrict__ ovar,
                                 const float* __restrict__ ow, int D, int M, int newM,
                                 float* __restrict__ mean, float* __restrict__ var, float* __restrict__ w) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;  // new column
  if (j >= newM) return;
  int m = j < M ? j : j - M;
  float sgn = j < M ? -1.f : 1.f;
  float vmax = ovar[m];
  int arg = 0;
  for (int d = 1; d < D; ++d) {
    float v = ovar[(size_t)d * M + m];
    if (v > vmax) { vmax = v; arg = d; }  // first maximum, like numpy argmax
  }
  float p = __fmul_rn(0.55f, sqrtf(vmax));
  for (int d = 0; d < D; ++d) {
    float mu = omean[(size_t)d * M + m];
    mean[(size_t)d * newM + j] = (d == arg) ? mu + sgn * p : mu;
    var[(size_t)d * newM + j] = ovar[(size_t)d * M + m];
  }
  w[j] = 0.5f * ow[m];
}

// ---------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------
static size_t lse_smem(int D) { return (size_t)(2 * D) * (AS_LD + TM) * sizeof(float); }
static int pick_jt(int D) {
  int need = 2 * D + 1;
  for (int jt = 4; jt <= 16; jt += 4)
    if (16 * jt >= need) return jt;
  return -1;
}
static size_t stats_smem(int D, int jt) {
  return ((size_t)(2 * D) * (AS_LD + TM) + (size_t)TF * 16 * jt + (size_t)TF * TM) * sizeof(float);
}

int gmm_refresh_constants(odin_gmm* g, cudaStream_t st) {
  int threads = 128;
  gmm_prepare_kernel<<<ceil_div(g->Mpad, threads), threads, 0, st>>>(g->d_mean, g->d_var, g->d_w, g->D, g->M,
                                                                     g->Mpad, g->d_Wk, g->d_cst);
  ODIN_LAUNCH_CHECK("gmm_prepare_kernel");
  if (gmm_tc_supported(g)) return gmm_tc_refresh(g, st);
  return ODIN_OK;
}

int gmm_reserve_lse(odin_gmm* g, int64_t n) {
  if (n <= g->lse_cap) return ODIN_OK;
  if (g->d_lse) ODIN_CUDA_CHECK(cudaFree(g->d_lse));
  g->d_lse = nullptr;
  g->lse_cap = 0;
  int64_t cap = n + n / 8 + 1024;
  ODIN_CUDA_CHECK(cudaMalloc(&g->d_lse, cap * sizeof(float)));
  g->lse_cap = cap;
  return ODIN_OK;
}

int gmm_lse_ffma(odin_gmm* g, const float* X, const uint8_t* sad, int64_t N, float* lse, double* stats,
                 cudaStream_t st) {
  if (N <= 0) return ODIN_OK;
  size_t smem = lse_smem(g->D);
  ODIN_CUDA_CHECK(cudaFuncSetAttribute(gmm_lse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t tiles = ceil_div<int64_t>(N, TF);
  double* statL = stats ? stats + (stats_size(g->D, g->M) - 2) : nullptr;
  gmm_lse_kernel<<<(unsigned)tiles, NT, smem, st>>>(X, sad, N, g->D, g->Mpad, ceil_div(g->M, TM), g->d_Wk, g->d_cst,
                                                    lse, statL);
  ODIN_LAUNCH_CHECK("gmm_lse_kernel");
  return ODIN_OK;
}

template <int MODE>
static int launch_stats(odin_gmm* g, StatsArgs& a, dim3 grid, cudaStream_t st) {
  int jt = pick_jt(g->D);
  if (jt < 0) return set_error(ODIN_EINVAL, "feat_dim %d > 127 unsupported", g->D);
  size_t smem = stats_smem(g->D, jt);
  if (smem > 227 * 1024) return set_error(ODIN_EINVAL, "feat_dim %d needs %zu B smem", g->D, smem);
#define ODIN_STATS_CASE(J)                                                                               \
  case J:                                                                                                \
    ODIN_CUDA_CHECK(cudaFuncSetAttribute(gmm_stats_kernel<J, MODE>,                                      \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));       \
    gmm_stats_kernel<J, MODE><<<grid, NT, smem, st>>>(a);                                                \
    break;
  switch (jt) {
    ODIN_STATS_CASE(4)
    ODIN_STATS_CASE(8)
    ODIN_STATS_CASE(12)
    ODIN_STATS_CASE(16)
  }
#undef ODIN_STATS_CASE
  ODIN_LAUNCH_CHECK("gmm_stats_kernel");
  return ODIN_OK;
}

int gmm_stats_ffma(odin_gmm* g, const float* X, const uint8_t* sad, int64_t N, const float* lse,
                   int want_second, double* stats, cudaStream_t st) {
  if (N <= 0) return ODIN_OK;
  StatsArgs a{};
  a.X = X; a.sad = sad; a.lse = lse; a.N = N; a.D = g->D; a.M = g->M; a.Mpad = g->Mpad;
  a.Wk = g->d_Wk; a.cst = g->d_cst; a.stats = stats; a.want_second = want_second;
  int nchunks = ceil_div(g->M, TM);
  int64_t tiles = ceil_div<int64_t>(N, TF);
  int64_t want_splits = std::max<int64_t>(1, (int64_t)(2 * sm_count()) / nchunks);
  int64_t nsplit = std::min<int64_t>(tiles, want_splits);
  a.tiles_per_split = (int)ceil_div<int64_t>(tiles, nsplit);
  nsplit = ceil_div<int64_t>(tiles, a.tiles_per_split);
  return launch_stats<0>(g, a, dim3(nchunks, (unsigned)nsplit), st);
}

int gmm_utt_stats_ffma(odin_gmm* g, const float* X, const uint8_t* sad, const int64_t* d_off, int n_utt,
                       const float* lse, float* Z, float* Fhat, cudaStream_t st) {
  if (n_utt <= 0) return ODIN_OK;
  StatsArgs a{};
  a.X = X; a.sad = sad; a.lse = lse; a.N = 0; a.D = g->D; a.M = g->M; a.Mpad = g->Mpad;
  a.Wk = g->d_Wk; a.cst = g->d_cst; a.off = d_off; a.mean = g->d_mean; a.Z = Z; a.Fhat = Fhat;
  int nchunks = ceil_div(g->M, TM);
  for (int u0 = 0; u0 < n_utt; u0 += 65535) {  // grid.y limit
    int nu = std::min(65535, n_utt - u0);
    StatsArgs b = a;
    b.off = d_off + u0;
    b.Z = Z + (size_t)u0 * g->M;
    b.Fhat = Fhat + (size_t)u0 * g->M * g->D;
    int rc = launch_stats<1>(g, b, dim3(nchunks, nu), st);
    if (rc) return rc;
  }
  return ODIN_OK;
}

int gmm_post_ffma(odin_gmm* g, const float* X, int64_t N, const float* lse, float* post, float* logp,
                  cudaStream_t st) {
  if (N <= 0) return ODIN_OK;
  size_t smem = lse_smem(g->D);
  ODIN_CUDA_CHECK(cudaFuncSetAttribute(gmm_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int64_t tiles = ceil_div<int64_t>(N, TF);
  int nchunks = ceil_div(g->M, TM);
  for (int64_t t0 = 0; t0 < tiles; t0 += (1 << 30)) {
    int64_t nt = std::min<int64_t>(1 << 30, tiles - t0);
    gmm_post_kernel<<<dim3((unsigned)nt, nchunks), NT, smem, st>>>(X + t0 * TF * g->D, N - t0 * TF, g->D, g->M,
                                                                  g->Mpad, g->d_Wk, g->d_cst, lse + t0 * TF,
                                                                  post ? post + t0 * TF * (int64_t)g->M : nullptr,
                                                                  logp ? logp + t0 * TF * (int64_t)g->M : nullptr);
    ODIN_LAUNCH_CHECK("gmm_post_kernel");
  }
  return ODIN_OK;
}

int gmm_mstep_launch(odin_gmm* g, const double* stats, int allow_rollback, int* rolled_back, cudaStream_t st) {
  gmm_mstep_kernel<<<1, 1024, 0, st>>>(stats, g->D, g->M, allow_rollback, g->d_mean, g->d_var, g->d_w,
                                       g->d_prev, rolled_back);
  ODIN_LAUNCH_CHECK("gmm_mstep_kernel");
  return ODIN_OK;
}

int gmm_mixup_launch(odin_gmm* g, int newM, cudaStream_t st) {
  const int D = g->D, M = g->M;
  size_t dm = (size_t)D * M;
  float* omean = g->d_prev;
  float* ovar = g->d_prev + dm;
  float* ow = g->d_prev + 2 * dm;
  ODIN_CUDA_CHECK(cudaMemcpyAsync(omean, g->d_mean, dm * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ODIN_CUDA_CHECK(cudaMemcpyAsync(ovar, g->d_var, dm * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ODIN_CUDA_CHECK(cudaMemcpyAsync(ow, g->d_w, M * sizeof(float), cudaMemcpyDeviceToDevice, st));
  gmm_mixup_kernel<<<ceil_div(newM, 128), 128, 0, st>>>(omean, ovar, ow, D, M, newM, g->d_mean, g->d_var, g->d_w);
  ODIN_LAUNCH_CHECK("gmm_mixup_kernel");
  return ODIN_OK;
}

}  // namespace odin
