// The speech chain ONE STAGE AT A TIME (SURVEY 8a rows a2, a6, a8-a11 as stand-alone calls; 8b import surface
// `pp.signal.*` and the `_transform` of every single extractor).  The fused kernels (fe_kernels.cu, fe_frame5.cu)
// are the hot path; these entry points exist so that a recipe that calls `signal.pre_emphasis`, `signal.stft`,
// `signal.power_spectrogram`, `signal.mels_spectrogram`, `signal.ceps_spectrogram` or `signal.delta` directly -- or
// runs `STFTExtractor(...).transform(X)` on its own -- gets the same arithmetic from the device instead of a
// NotImplementedError.  All of them are HBM-bound passes over [T, width] matrices (algorithmic bytes = input +
// output rows); the arithmetic the reference does in float64 (filterbank / DCT dot products, lfilter) is done in
// fp64 here and cast to float32 on the way out.
//
//   pre_emphasis       signal.py:955-967        y[0] = s[0], y[t] = s[t] - c s[t-1] (float32, two roundings)
//   power_spectrogram  signal.py:1623-1648      |S| ** int(power)
//   mels_spectrogram   signal.py:1650-1691      mel_basis @ spec^T, power2db with the matrix-global top_db clip (:636-680)
//   ceps_spectrogram   signal.py:1693-1716      dct_basis @ mspec^T (optionally without row 0)
//   delta              signal.py:1002-1066      `order` causal lfilter passes over the edge-padded rows, every
//                                               result cut with the same window (the delay quirk)
#include <float.h>
#include <math.h>

#include <algorithm>
#include <vector>

#include "fe.cuh"

namespace odin {

__device__ __forceinline__ int sig_find(const int64_t* __restrict__ off, int n, int64_t v) {
  int lo = 0, hi = n;   // largest u with off[u] <= v
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// y[0] = s[0]; y[t] = s[t] - coeff * s[t-1] within every segment (1-D signals back to back).  rows2d: the 2-D form
// of the reference, s - c_[s[:, :1], s[:, :-1]] * coeff, whose first column is s[:, 0] * (1 - coeff) in two roundings.
__global__ void __launch_bounds__(256) sig_preemph_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                          const int64_t* __restrict__ off, int n_seg, float coeff, int rows2d) {
  const int64_t n = off[n_seg];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int u = sig_find(off, n_seg, i);
    const bool first = i == off[u];
    const float prev = first ? x[i] : x[i - 1];
    y[i] = (first && !rows2d) ? x[i] : __fsub_rn(x[i], __fmul_rn(prev, coeff));
  }
}

// |S| ** power: complex input -> hypot (np.abs), then the integer power by repeated multiplication
__global__ void __launch_bounds__(256) sig_power_kernel(const float* __restrict__ s, float* __restrict__ out, int64_t n,
                                                        int is_complex, int power) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double m;
    if (is_complex) {
      const float2 z = reinterpret_cast<const float2*>(s)[i];
      m = (power == 2) ? (double)z.x * z.x + (double)z.y * z.y : hypot((double)z.x, (double)z.y);
    } else {
      m = (double)s[i];
      if (power == 2) m = m * m;
    }
    if (power > 2) {
      double r = m;
      for (int p = 1; p < power; ++p) r *= m;
      m = r;
    }
    out[i] = (float)m;
  }
}

// one warp per frame: sparse triangles (CSR rows of the handle's filterbank), fp64 accumulation in bin order,
// 10 log10(max(1e-10, .)), per-utterance maximum for the clip pass
__global__ void __launch_bounds__(256) sig_mels_kernel(const float* __restrict__ spec, float* __restrict__ mspec, int64_t T,
                                                       int nbins, int n_mels, const int* __restrict__ mstart,
                                                       const int* __restrict__ mcnt, const int* __restrict__ moff,
                                                       const float* __restrict__ mw, const int64_t* __restrict__ off, int n_utt,
                                                       int* __restrict__ umax, int log_db) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t t = wid; t < T; t += nw) {
    const float* row = spec + t * nbins;
    float vmax = -FLT_MAX;
    for (int m = lane; m < n_mels; m += 32) {
      const int st = mstart[m], cn = mcnt[m];
      const float* w = mw + moff[m];
      double acc = 0.0;
      for (int i = 0; i < cn; ++i) acc = fma((double)w[i], (double)row[st + i], acc);
      const float v = log_db ? (float)(10.0 * log10(fmax(acc, 1e-10))) : (float)acc;
      mspec[t * n_mels + m] = v;
      vmax = fmaxf(vmax, v);
    }
    if (log_db) {
      vmax = warp_max(vmax);
      if (lane == 0) atomicMax(umax + sig_find(off, n_utt, t), float_to_ordered(vmax));
    }
  }
}

__global__ void __launch_bounds__(256) sig_clip_kernel(float* __restrict__ x, int64_t T, int width, const int64_t* __restrict__ off,
                                                       int n_utt, const int* __restrict__ umax, float top_db) {
  const int64_t n = T * width;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float floor_db = ordered_to_float(umax[sig_find(off, n_utt, i / width)]) - top_db;
    if (x[i] < floor_db) x[i] = floor_db;
  }
}

// out[t][c] = sum_m dct[c0 + c][m] * mspec[t][m]  (fp64, m ascending); dct is the handle's [n_rows, n_mels] basis
__global__ void __launch_bounds__(256) sig_ceps_kernel(const float* __restrict__ mspec, float* __restrict__ out, int64_t T,
                                                       int n_mels, const double* __restrict__ dct, int c0, int n_out) {
  extern __shared__ double sdct[];   // [n_out][n_mels]
  for (int i = threadIdx.x; i < n_out * n_mels; i += blockDim.x) sdct[i] = dct[(size_t)c0 * n_mels + i];
  __syncthreads();
  const int64_t n = T * n_out;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / n_out;
    const int c = (int)(i - t * n_out);
    const float* row = mspec + t * n_mels;
    const double* d = sdct + c * n_mels;
    double acc = 0.0;
    for (int m = 0; m < n_mels; ++m) acc = fma(d[m], (double)row[m], acc);
    out[i] = (float)acc;
  }
}

// delta of order 1 and 2 (signal.py:1002-1066).  xpad = rows edge-padded by W on both sides; y1[q] = sum_{i <= min(W-1, q)}
// win[i] xpad[q - i] (lfilter, zero initial state); d1[t] = y1[t + 2W - h], d2[t] = sum_i win[i] y1[t + 2W - h - i] with
// h = (W + 1) / 2 -- the second pass is delayed by W / 2 frames and sees the zero state at the utterance head.
__global__ void __launch_bounds__(256) sig_delta_kernel(const float* __restrict__ x, float* __restrict__ d1, float* __restrict__ d2,
                                                        int64_t T, int dim, const int64_t* __restrict__ off, int n_utt, int W) {
  const int h = (W + 1) / 2;
  double win[32];
  {
    double norm = 0.0;
    for (int m = -(W / 2); m <= W / 2; ++m) norm += (double)m * m;
    for (int i = 0; i < W; ++i) win[i] = (double)(W / 2 - i) / norm;
  }
  const int64_t n = T * dim;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / dim;
    const int c = (int)(i - t * dim);
    const int u = sig_find(off, n_utt, t);
    const int64_t base = off[u];
    const int Tu = (int)(off[u + 1] - base), tl = (int)(t - base);
    const float* col = x + base * dim + c;
    auto xpad = [&](int p) -> double { return (double)col[(int64_t)min(max(p - W, 0), Tu - 1) * dim]; };
    auto y1 = [&](int q) -> double {
      double acc = 0.0;
      const int top = min(W - 1, q);
      for (int k = 0; k <= top; ++k) acc = fma(win[k], xpad(q - k), acc);
      return acc;
    };
    const int j = tl + 2 * W - h;
    d1[i] = (float)y1(j);
    if (d2 != nullptr) {
      double acc = 0.0;
      for (int k = 0; k < W; ++k) acc = fma(win[k], y1(j - k), acc);
      d2[i] = (float)acc;
    }
  }
}

struct SigOffsets {   // host frame offsets -> stream-ordered device copy (+ optional per-utterance int scratch)
  int64_t* d_off = nullptr;
  int* d_umax = nullptr;
  cudaStream_t st;
  int init(const int64_t* h_off, int n_utt, bool want_umax, cudaStream_t s) {
    st = s;
    ODIN_CUDA_CHECK(cudaMallocAsync(&d_off, sizeof(int64_t) * (n_utt + 1), st));
    ODIN_CUDA_CHECK(cudaMemcpyAsync(d_off, h_off, sizeof(int64_t) * (n_utt + 1), cudaMemcpyHostToDevice, st));
    if (want_umax) {
      ODIN_CUDA_CHECK(cudaMallocAsync(&d_umax, sizeof(int) * n_utt, st));
      ODIN_CUDA_CHECK(cudaMemsetAsync(d_umax, 0x80, sizeof(int) * n_utt, st));   // ordered_to_float -> about -3.4e38
    }
    return ODIN_OK;
  }
  ~SigOffsets() {
    if (d_off) cudaFreeAsync(d_off, st);
    if (d_umax) cudaFreeAsync(d_umax, st);
  }
};

static unsigned sig_grid(int64_t n, int per_thread = 1) {
  return (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div<int64_t>(n, 256 * per_thread), (int64_t)sm_count() * 16));
}

static int check_offsets(const int64_t* h_off, int n_seg) {
  if (!h_off || n_seg < 0) return set_error(ODIN_EINVAL, "bad offsets");
  if (h_off[0] != 0) return set_error(ODIN_EINVAL, "offsets must start at 0");
  for (int u = 0; u < n_seg; ++u)
    if (h_off[u + 1] < h_off[u]) return set_error(ODIN_EINVAL, "offsets must be non-decreasing");
  return ODIN_OK;
}

}  // namespace odin

using namespace odin;

extern "C" {

int odin_sig_preemph(const float* d_x, float* d_y, const int64_t* h_offsets, int32_t n_seg, float coeff, int32_t rows2d,
                     void* stream) {
  if (!d_x || !d_y) return set_error(ODIN_EINVAL, "bad argument");
  int rc = check_offsets(h_offsets, n_seg);
  if (rc || (rc = require_device())) return rc;
  if (n_seg == 0 || h_offsets[n_seg] == 0) return ODIN_OK;
  SigOffsets o;
  if ((rc = o.init(h_offsets, n_seg, false, as_stream(stream)))) return rc;
  sig_preemph_kernel<<<sig_grid(h_offsets[n_seg]), 256, 0, o.st>>>(d_x, d_y, o.d_off, n_seg, coeff, rows2d);
  ODIN_LAUNCH_CHECK("sig_preemph_kernel");
  return ODIN_OK;
}

int odin_sig_power(const float* d_s, int32_t is_complex, int32_t power, float* d_out, int64_t n, void* stream) {
  if (!d_s || !d_out || n < 0 || power < 1 || power > 16) return set_error(ODIN_EINVAL, "bad argument (power must be 1..16)");
  int rc = require_device();
  if (rc) return rc;
  if (n == 0) return ODIN_OK;
  sig_power_kernel<<<sig_grid(n), 256, 0, as_stream(stream)>>>(d_s, d_out, n, is_complex, power);
  ODIN_LAUNCH_CHECK("sig_power_kernel");
  return ODIN_OK;
}

int odin_fe_mels(odin_fe_t* fe, const float* d_spec, const int64_t* h_frame_offsets, int32_t n_utt, float* d_mspec,
                 int32_t log_db, void* stream) {
  if (!fe || !d_spec || !d_mspec) return set_error(ODIN_EINVAL, "bad argument");
  int rc = check_offsets(h_frame_offsets, n_utt);
  if (rc) return rc;
  const int64_t T = h_frame_offsets[n_utt];
  if (n_utt == 0 || T == 0) return ODIN_OK;
  SigOffsets o;
  if ((rc = o.init(h_frame_offsets, n_utt, true, as_stream(stream)))) return rc;
  sig_mels_kernel<<<sig_grid(T * 32), 256, 0, o.st>>>(d_spec, d_mspec, T, fe->nbins, fe->n_mels, fe->d_mel_start, fe->d_mel_cnt,
                                                     fe->d_mel_off, fe->d_mel_w, o.d_off, n_utt, o.d_umax, log_db);
  ODIN_LAUNCH_CHECK("sig_mels_kernel");
  if (log_db && fe->cfg.top_db >= 0.f) {
    sig_clip_kernel<<<sig_grid(T * fe->n_mels), 256, 0, o.st>>>(d_mspec, T, fe->n_mels, o.d_off, n_utt, o.d_umax, fe->cfg.top_db);
    ODIN_LAUNCH_CHECK("sig_clip_kernel");
  }
  return ODIN_OK;
}

int odin_fe_ceps(odin_fe_t* fe, const float* d_mspec, int64_t n_frames, int32_t first_row, int32_t n_rows, float* d_out,
                 void* stream) {
  if (!fe || !d_mspec || !d_out || n_frames < 0) return set_error(ODIN_EINVAL, "bad argument");
  if (first_row < 0 || n_rows < 1 || first_row + n_rows > fe->n_c1)
    return set_error(ODIN_EINVAL, "DCT rows [%d, %d) outside the handle's basis (%d rows)", first_row, first_row + n_rows, fe->n_c1);
  if (n_frames == 0) return ODIN_OK;
  if (fe->d_dct64 == nullptr) {
    ODIN_CUDA_CHECK(cudaMalloc(&fe->d_dct64, sizeof(double) * fe->h_dct.size()));
    ODIN_CUDA_CHECK(cudaMemcpy(fe->d_dct64, fe->h_dct.data(), sizeof(double) * fe->h_dct.size(), cudaMemcpyHostToDevice));
  }
  const size_t smem = sizeof(double) * (size_t)n_rows * fe->n_mels;
  if (smem > 48 * 1024) ODIN_CUDA_CHECK(cudaFuncSetAttribute(sig_ceps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sig_ceps_kernel<<<sig_grid(n_frames * n_rows), 256, smem, as_stream(stream)>>>(d_mspec, d_out, n_frames, fe->n_mels, fe->d_dct64,
                                                                               first_row, n_rows);
  ODIN_LAUNCH_CHECK("sig_ceps_kernel");
  return ODIN_OK;
}

int odin_sig_delta(const float* d_x, int32_t dim, const int64_t* h_frame_offsets, int32_t n_utt, int32_t width, int32_t order,
                   float* d_delta1, float* d_delta2, void* stream) {
  if (!d_x || !d_delta1 || dim < 1) return set_error(ODIN_EINVAL, "bad argument");
  if (width < 3 || (width & 1) == 0 || width > 31) return set_error(ODIN_EINVAL, "width must be an odd integer in 3..31 (signal.py:1034-1035)");
  if (order < 1 || order > 2 || (order == 2 && !d_delta2)) return set_error(ODIN_EINVAL, "order must be 1 or 2 (with d_delta2 for 2)");
  int rc = check_offsets(h_frame_offsets, n_utt);
  if (rc || (rc = require_device())) return rc;
  const int64_t T = h_frame_offsets[n_utt];
  if (n_utt == 0 || T == 0) return ODIN_OK;
  SigOffsets o;
  if ((rc = o.init(h_frame_offsets, n_utt, false, as_stream(stream)))) return rc;
  sig_delta_kernel<<<sig_grid(T * dim), 256, 0, o.st>>>(d_x, d_delta1, order == 2 ? d_delta2 : nullptr, T, dim, o.d_off, n_utt, width);
  ODIN_LAUNCH_CHECK("sig_delta_kernel");
  return ODIN_OK;
}

}  // extern "C"
