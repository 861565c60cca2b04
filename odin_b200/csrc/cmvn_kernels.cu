// CMVN: AcousticNorm (odin/preprocessing/speech.py:1536-1610) = signal.mvn (signal.py:853-876)
// followed by signal.wmvn(varnorm=False) (signal.py:878-924), over a ragged batch of utterances.
//
// The two reference stages collapse into one affine map per element,
//     y[t,d] = (x[t,d] - m[t,d]) * inv[d]
//   m[t,d]  = masked mean over the window of frame t  (windowed, nobs >= w)
//           = masked mean over the utterance           (otherwise)
//   inv[d]  = 1 / (masked utterance std + 1e-18)       (mean_var_norm && var_norm), else 1
// because mvn is affine per dimension: the window mean of (x - mu)/sigma is (winmean(x) - mu)/sigma.
// (The reference's second stage sees float64 round-off of the first; the difference is ~1e-16.)
// The window of frame t is [lo, lo + w) with lo = clamp(t - h, 0, nobs - w): the first / last h
// frames share the first / last full window exactly as signal.py:905-923 does.
// Statistics are accumulated in fp64; an empty selection gives NaN like numpy's mean of nothing.
//
// One CTA per utterance.  Phase 1: warps stride over frames, lanes over dimensions (coalesced rows),
// sum / sum of squares / count -> mean, inv in smem.  Phase 2: the utterance is cut into segments,
// each (segment, dimension) thread slides its window sum along its frames.
#include <math.h>

#include <algorithm>

#include "fe.cuh"

namespace odin {

constexpr int CM_THREADS = 256;
constexpr int CM_MAXD = 256;

struct CmvnArgs {
  const float* x;
  float* y;
  int dim;
  const int64_t* frame_off;
  int n_utt;
  const uint8_t* sad;   // nullable
  int mean_var_norm, var_norm, windowed, w;
};

__global__ void __launch_bounds__(CM_THREADS) fe_cmvn_kernel(CmvnArgs a) {
  __shared__ double s_sum[CM_MAXD], s_sq[CM_MAXD], s_mu[CM_MAXD], s_inv[CM_MAXD];
  __shared__ double s_cnt;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = a.dim;
  for (int u = blockIdx.x; u < a.n_utt; u += gridDim.x) {
    const int64_t base = a.frame_off[u];
    const int n = (int)(a.frame_off[u + 1] - base);
    if (n <= 0) continue;
    const float* x = a.x + base * D;
    float* y = a.y + base * D;
    const uint8_t* sad = a.sad ? a.sad + base : nullptr;
    const bool slide = a.windowed && n >= a.w;
    const bool need_global = a.mean_var_norm || (a.windowed && !slide);
    __syncthreads();
    for (int d = tid; d < D; d += CM_THREADS) { s_sum[d] = 0.0; s_sq[d] = 0.0; s_mu[d] = 0.0; s_inv[d] = 1.0; }
    if (tid == 0) s_cnt = 0.0;
    __syncthreads();
    if (need_global) {
      double cnt = 0.0;
      for (int d0 = 0; d0 < D; d0 += 32) {
        const int d = d0 + lane;
        double sm = 0.0, sq = 0.0;
        for (int t = warp; t < n; t += CM_THREADS / 32) {
          if (sad != nullptr && sad[t] == 0) continue;
          if (d0 == 0) cnt += 1.0;
          if (d < D) {
            const double v = (double)x[(int64_t)t * D + d];
            sm += v;
            sq = fma(v, v, sq);
          }
        }
        if (d < D) { atomicAdd(&s_sum[d], sm); atomicAdd(&s_sq[d], sq); }
      }
      if (lane == 0) atomicAdd(&s_cnt, cnt);
      __syncthreads();
      const double c = s_cnt;
      for (int d = tid; d < D; d += CM_THREADS) {
        const double mu = s_sum[d] / c;                       // 0/0 -> NaN like numpy
        double var = s_sq[d] / c - mu * mu;
        if (var < 0.0) var = 0.0;
        s_mu[d] = mu;
        if (a.mean_var_norm && a.var_norm) s_inv[d] = 1.0 / (sqrt(var) + 1e-18);
      }
      __syncthreads();
    }
    if (!slide) {
      // y = (x - mu) * inv  (if neither stage applies, mu = 0 and inv = 1: a copy)
      const int64_t total = (int64_t)n * D;
      for (int64_t i = tid; i < total; i += CM_THREADS) {
        const int d = (int)(i % D);
        y[i] = (float)(((double)x[i] - s_mu[d]) * s_inv[d]);
      }
      continue;
    }
    // ---- sliding window mean (w odd, n >= w)
    const int Dg = ((D + 31) / 32) * 32;
    const int nseg = max(1, CM_THREADS / Dg);
    const int w = a.w, h = (w - 1) / 2;
    for (int d0 = 0; d0 < D; d0 += CM_THREADS) {   // D <= 256: one trip
      const int seg = tid / Dg, d = d0 + tid % Dg;
      if (seg >= nseg || d >= D) continue;
      const int t_beg = (int)((int64_t)n * seg / nseg), t_end = (int)((int64_t)n * (seg + 1) / nseg);
      if (t_beg >= t_end) continue;
      int lo = min(max(t_beg - h, 0), n - w);
      double sm = 0.0, cnt = 0.0;
      for (int t = lo; t < lo + w; ++t)
        if (sad == nullptr || sad[t] != 0) { sm += (double)x[(int64_t)t * D + d]; cnt += 1.0; }
      const double inv = s_inv[d];
      for (int t = t_beg; t < t_end; ++t) {
        const int nlo = min(max(t - h, 0), n - w);
        if (nlo != lo) {   // the window advances by exactly one frame
          const int out = lo, in = lo + w;
          if (sad == nullptr || sad[out] != 0) { sm -= (double)x[(int64_t)out * D + d]; cnt -= 1.0; }
          if (sad == nullptr || sad[in] != 0) { sm += (double)x[(int64_t)in * D + d]; cnt += 1.0; }
          lo = nlo;
        }
        y[(int64_t)t * D + d] = (float)(((double)x[(int64_t)t * D + d] - sm / cnt) * inv);
      }
    }
  }
}

int fe_cmvn_launch(const float* d_x, float* d_y, int dim, const int64_t* d_frame_off, int n_utt,
                   const uint8_t* d_sad, int mean_var_norm, int var_norm, int windowed, int win_length,
                   cudaStream_t st) {
  CmvnArgs a{};
  a.x = d_x; a.y = d_y; a.dim = dim; a.frame_off = d_frame_off; a.n_utt = n_utt; a.sad = d_sad;
  a.mean_var_norm = mean_var_norm; a.var_norm = var_norm; a.windowed = windowed; a.w = win_length;
  const int grid = std::min(n_utt, sm_count() * 8);
  fe_cmvn_kernel<<<grid, CM_THREADS, 0, st>>>(a);
  ODIN_LAUNCH_CHECK("fe_cmvn_kernel");
  return ODIN_OK;
}

}  // namespace odin
