// Feature-matrix extractors that sit beside the fused front-end, over ragged batches
// [T, dim] with host frame offsets:
//   StackFeatures   base.py:724-771 -> signal.stack_frames(keep_length=True)  signal.py:1225-1294
//   RASTAfilter     speech.py:1483-1533 -> signal.rastafilt (signal.py:926-953) + signal.shifted_deltas
//                   (signal.py:1068-1090, deltas of signal.py:1002-1066)
//   CalculateEnergy speech.py:623-649 -> signal.get_energy (signal.py:1421-1440)
// All are HBM-bound copies / short recurrences; the arithmetic that the reference does in float64 (RASTA,
// deltas, energy sums) is done in fp64 here and cast to float32 where the reference casts.
#include <cuda_fp16.h>
#include <algorithm>

#include "common.cuh"

namespace odin {

__device__ __forceinline__ int feat_find(const int64_t* __restrict__ off, int n, int64_t v) {
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (off[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

// Frames are walked in contiguous per-CTA ranges: the utterance of the first frame is found by one binary search,
// after that the index only moves forward (a per-frame search was 11 dependent loads in front of every row and
// held StackFeatures at 0.9 TB/s).
struct FrameWalk {
  int u;
  int64_t lo, hi;
  __device__ FrameWalk(const int64_t* __restrict__ off, int n_utt, int64_t t) {
    u = feat_find(off, n_utt, t);
    lo = off[u]; hi = off[u + 1];
  }
  __device__ __forceinline__ void to(const int64_t* __restrict__ off, int64_t t) {
    while (t >= hi) { ++u; lo = hi; hi = off[u + 1]; }
  }
};

// y[t] = [x[t-c], ..., x[t+c]] flattened, zeros outside the utterance
__global__ void __launch_bounds__(256) feat_stack_kernel(const float* __restrict__ x, float* __restrict__ y, int dim,
                                                         const int64_t* __restrict__ off, int n_utt, int c) {
  const int64_t T = off[n_utt];
  const int w = (2 * c + 1) * dim;
  const int64_t per = (T + gridDim.x - 1) / gridDim.x;
  const int64_t t_lo = per * blockIdx.x, t_hi = min(T, t_lo + per);
  if (t_lo >= t_hi) return;
  FrameWalk fw(off, n_utt, t_lo);
  const uint32_t mg = 0xFFFFFFFFu / (uint32_t)dim + 1u;   // j / dim by multiply-high (dim > 1; exact for j < 2^16)
  for (int64_t t = t_lo; t < t_hi; ++t) {
    fw.to(off, t);
    float* row = y + t * w;
    for (int j = threadIdx.x; j < w; j += 256) {
      const int k = (dim == 1) ? j : (int)__umulhi((uint32_t)j, mg), d = j - k * dim;
      const int64_t s = t + k - c;
      row[j] = (s >= fw.lo && s < fw.hi) ? x[s * dim + d] : 0.f;
    }
  }
}

// RASTA along time, one thread per (utterance, column): scipy's transposed direct-form II recurrence with the
// reference's warm-up (the first four outputs are zero; the FIR run over them only leaves the filter state).
// y64 [T, dim] keeps the fp64 result for the delta pass, y32 (row stride ld) receives the float32 cast.
// The recurrence is sequential in time, so the loads are issued eight rows ahead of the dependent chain.
__global__ void __launch_bounds__(128) feat_rasta_kernel(const float* __restrict__ x, double* __restrict__ y64,
                                                         float* __restrict__ y32, int ld, int dim,
                                                         const int64_t* __restrict__ off, int n_utt, int rasta) {
  const int64_t id = (int64_t)blockIdx.x * 128 + threadIdx.x;
  if (id >= (int64_t)n_utt * dim) return;
  const int u = (int)(id / dim), d = (int)(id - (int64_t)u * dim);
  const int64_t lo = off[u], hi = off[u + 1];
  const double b0 = 0.2, b1 = 0.1, b2 = -0.0, b3 = -0.1, b4 = -0.2;   // -arange(-2, 3) / 10
  double z0 = 0.0, z1 = 0.0, z2 = 0.0, z3 = 0.0;
  constexpr int PF = 8;
  for (int64_t t0 = lo; t0 < hi; t0 += PF) {
    float xv[PF];
#pragma unroll
    for (int q = 0; q < PF; ++q) xv[q] = (t0 + q < hi) ? x[(t0 + q) * dim + d] : 0.f;
#pragma unroll
    for (int q = 0; q < PF; ++q) {
      const int64_t t = t0 + q;
      if (t < hi) {
        const double xn = (double)xv[q];
        double yn;
        if (!rasta) {
          yn = xn;
        } else {
          // one formula for both phases: during the four warm-up frames the output is zero and the pole is off
          const bool warm = (t - lo) < 4;
          const double yr = __dadd_rn(__dmul_rn(b0, xn), z0);
          yn = warm ? 0.0 : yr;
          const double fb = warm ? 0.0 : __dmul_rn(0.94, yr);
          z0 = __dadd_rn(__dadd_rn(__dmul_rn(b1, xn), z1), fb);
          z1 = __dadd_rn(__dmul_rn(b2, xn), z2);
          z2 = __dadd_rn(__dmul_rn(b3, xn), z3);
          z3 = __dmul_rn(b4, xn);
        }
        y64[t * dim + d] = yn;
        y32[t * ld + d] = (float)yn;
      }
    }
  }
}

// first-order delta of every frame (width 2 sdc + 1, clamped edges, fp64, cast to float32 like signal.delta)
__global__ void __launch_bounds__(256) feat_delta1_kernel(const double* __restrict__ y64, float* __restrict__ dx, int dim,
                                                          const int64_t* __restrict__ off, int n_utt, int sdc) {
  const int64_t T = off[n_utt];
  const int h = sdc, W = 2 * sdc + 1;
  double norm = 0.0;
  for (int m = -h; m <= h; ++m) norm += (double)m * m;
  const int64_t per = (T + gridDim.x - 1) / gridDim.x;
  const int64_t t_lo = per * blockIdx.x, t_hi = min(T, t_lo + per);
  if (t_lo >= t_hi) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;   // one warp per frame, lanes over the columns
  if (t_lo + warp >= t_hi) return;
  FrameWalk fw(off, n_utt, t_lo + warp);
  for (int64_t t = t_lo + warp; t < t_hi; t += 8) {
    fw.to(off, t);
    const int64_t lo = fw.lo, hi = fw.hi;
    for (int d = lane; d < dim; d += 32) {
      double acc = 0.0;
      for (int k = 0; k < W; ++k) {                      // taps (h - k) / norm on frame t + h - k, k ascending
        int64_t s2 = t + h - k;
        s2 = s2 < lo ? lo : (s2 > hi - 1 ? hi - 1 : s2);
        acc = __dadd_rn(acc, __dmul_rn((double)(h - k) / norm, y64[s2 * dim + d]));
      }
      dx[t * dim + d] = (float)acc;
    }
  }
}

// shifted delta coefficients: block ix of row t = delta of frame min(t + 3 ix, T_u - 1), written behind the dim
// leading columns -- a pure gather of the float32 deltas (the fused version re-read three fp64 values per output)
__global__ void __launch_bounds__(256) feat_sdc_kernel(const float* __restrict__ dx, float* __restrict__ out, int ld,
                                                       int dim, const int64_t* __restrict__ off, int n_utt) {
  const int64_t T = off[n_utt];
  const int w = dim * dim;
  const int64_t per = (T + gridDim.x - 1) / gridDim.x;
  const int64_t t_lo = per * blockIdx.x, t_hi = min(T, t_lo + per);
  if (t_lo >= t_hi) return;
  FrameWalk fw(off, n_utt, t_lo);
  const uint32_t mg = 0xFFFFFFFFu / (uint32_t)dim + 1u;
  for (int64_t t = t_lo; t < t_hi; ++t) {
    fw.to(off, t);
    const int64_t hi = fw.hi;
    float* row = out + t * ld + dim;
    for (int j = threadIdx.x; j < w; j += 256) {
      const int ix = (dim == 1) ? j : (int)__umulhi((uint32_t)j, mg), d = j - ix * dim;
      int64_t f = t + 3 * (int64_t)ix;
      if (f > hi - 1) f = hi - 1;
      row[j] = dx[f * dim + d];
    }
  }
}

// log sum x^2 per row (fp64 accumulation), exact zero -> float32 eps (signal.py:1436)
__global__ void __launch_bounds__(256) feat_energy_kernel(const float* __restrict__ x, float* __restrict__ e, int64_t T,
                                                          int L, int take_log) {
  const int lane = threadIdx.x & 31;
  const int64_t row0 = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * 256) >> 5;
  for (int64_t t = row0; t < T; t += nw) {
    double acc = 0.0;
    for (int i = lane; i < L; i += 32) { const double v = (double)x[t * L + i]; acc = fma(v, v, acc); }
    acc = warp_sum(acc);
    if (lane == 0) {
      if (acc == 0.0) acc = 1.1920928955078125e-07;
      e[t] = (float)(take_log ? log(acc) : acc);
    }
  }
}

// signal.smooth(x, win, 'flat') (signal.py:969-1000) of a 0/1 vector: mirror extension 2 x[0] - x[win-1::-1] | x |
// 2 x[-1] - x[-1:-win:-1], np.convolve(w / win, s, 'same')[win : -win + 1].  Output t averages s[t + c .. t + c + win)
// with c = (win - 1) / 2 + 1; all terms are small integers, so the double result is exact.
__global__ void __launch_bounds__(256) feat_smooth_kernel(const uint8_t* __restrict__ x, double* __restrict__ y, int64_t n,
                                                          int win) {
  const int64_t c = (win - 1) / 2 + 1;
  for (int64_t t = (int64_t)blockIdx.x * 256 + threadIdx.x; t < n; t += (int64_t)gridDim.x * 256) {
    long long acc = 0;
    for (int j = 0; j < win; ++j) {
      const int64_t i = t + c + j;          // index into the extended vector s
      int v;
      if (i < win) v = 2 * (int)(x[0] != 0) - (int)(x[win - 1 - i] != 0);
      else if (i < win + n) v = (int)(x[i - win] != 0);
      else v = 2 * (int)(x[n - 1] != 0) - (int)(x[n - 1 - (i - win - n)] != 0);
      acc += v;
    }
    y[t] = (double)acc / (double)win;
  }
}

static int upload_offsets(const int64_t* h_off, int n_utt, int64_t** d_off, cudaStream_t st) {
  ODIN_CUDA_CHECK(cudaMallocAsync(d_off, sizeof(int64_t) * (n_utt + 1), st));
  cudaError_t e = cudaMemcpyAsync(*d_off, h_off, sizeof(int64_t) * (n_utt + 1), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { cudaFreeAsync(*d_off, st); return set_error(ODIN_ECUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e)); }
  return ODIN_OK;
}

// AsType (base.py:616-665) on the device: features stored as float16 (the recipes' AsType('float16') tail,
// examples/fsdd_ivec.py:105) cross PCIe at their stored width and are widened here, and feature rows leaving for
// a float16 store are narrowed (round to nearest even, like ndarray.astype) before the copy out.  HBM-bound:
// 16 bytes per thread and trip on the wider side.
template <typename S, typename D> __device__ __forceinline__ D cvt1(S v);
template <> __device__ __forceinline__ float cvt1<__half, float>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float cvt1<double, float>(double v) { return (float)v; }
template <> __device__ __forceinline__ __half cvt1<float, __half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ double cvt1<float, double>(float v) { return (double)v; }
template <> __device__ __forceinline__ float cvt1<float, float>(float v) { return v; }

template <typename S, typename D>
__global__ void __launch_bounds__(256) feat_convert_kernel(const S* __restrict__ src, D* __restrict__ dst, int64_t n) {
  constexpr int V = 4;
  const int64_t nv = n / V;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool aligned = ((reinterpret_cast<uintptr_t>(src) % (V * sizeof(S))) | (reinterpret_cast<uintptr_t>(dst) % (V * sizeof(D)))) == 0;
  struct alignas(V * sizeof(S)) SV { S e[V]; };
  struct alignas(V * sizeof(D)) DV { D e[V]; };
  if (aligned) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
      const SV a = reinterpret_cast<const SV*>(src)[i];
      DV b;
#pragma unroll
      for (int e = 0; e < V; ++e) b.e[e] = cvt1<S, D>(a.e[e]);
      reinterpret_cast<DV*>(dst)[i] = b;
    }
    for (int64_t i = nv * V + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = cvt1<S, D>(src[i]);
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = cvt1<S, D>(src[i]);
  }
}

template <typename S, typename D>
static int feat_convert_launch(const void* src, void* dst, int64_t n, cudaStream_t st) {
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(n, 4 * 256), (int64_t)sm_count() * 8);
  feat_convert_kernel<S, D><<<grid, 256, 0, st>>>(reinterpret_cast<const S*>(src), reinterpret_cast<D*>(dst), n);
  ODIN_LAUNCH_CHECK("feat_convert_kernel");
  return ODIN_OK;
}

}  // namespace odin

using namespace odin;

extern "C" {

int odin_feat_convert(const void* d_src, int32_t src_dtype, void* d_dst, int32_t dst_dtype, int64_t n, void* stream) {
  if (!d_src || !d_dst || n < 0) return set_error(ODIN_EINVAL, "bad argument");
  int rc = require_device();
  if (rc) return rc;
  if (n == 0) return ODIN_OK;
  cudaStream_t st = as_stream(stream);
  if (src_dtype == 0 && dst_dtype == 1) return feat_convert_launch<__half, float>(d_src, d_dst, n, st);
  if (src_dtype == 2 && dst_dtype == 1) return feat_convert_launch<double, float>(d_src, d_dst, n, st);
  if (src_dtype == 1 && dst_dtype == 0) return feat_convert_launch<float, __half>(d_src, d_dst, n, st);
  if (src_dtype == 1 && dst_dtype == 2) return feat_convert_launch<float, double>(d_src, d_dst, n, st);
  return set_error(ODIN_EINVAL, "odin_feat_convert: dtypes are 0 float16, 1 float32, 2 float64; one side must be float32");
}

int odin_feat_stack(const float* d_x, float* d_y, int32_t dim, const int64_t* h_frame_offsets, int32_t n_utt,
                    int32_t n_context, void* stream) {
  if (!d_x || !d_y || !h_frame_offsets || dim <= 0 || n_utt < 0 || n_context <= 0)
    return set_error(ODIN_EINVAL, "bad argument");
  int rc = require_device();
  if (rc) return rc;
  const int64_t T = n_utt ? h_frame_offsets[n_utt] : 0;
  if (T == 0) return ODIN_OK;
  cudaStream_t st = as_stream(stream);
  int64_t* d_off = nullptr;
  if ((rc = upload_offsets(h_frame_offsets, n_utt, &d_off, st))) return rc;
  const unsigned grid = (unsigned)std::min<int64_t>(T, (int64_t)sm_count() * 16);
  feat_stack_kernel<<<grid, 256, 0, st>>>(d_x, d_y, dim, d_off, n_utt, n_context);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(d_off, st);
  if (e != cudaSuccess) return set_error(ODIN_ECUDA, "launch feat_stack_kernel: %s", cudaGetErrorString(e));
  return ODIN_OK;
}

int odin_feat_rasta_sdc(const float* d_x, float* d_y, int32_t dim, const int64_t* h_frame_offsets, int32_t n_utt,
                        int32_t rasta, int32_t sdc, void* stream) {
  if (!d_x || !d_y || !h_frame_offsets || dim <= 0 || n_utt < 0 || sdc < 0) return set_error(ODIN_EINVAL, "bad argument");
  int rc = require_device();
  if (rc) return rc;
  const int64_t T = n_utt ? h_frame_offsets[n_utt] : 0;
  if (T == 0) return ODIN_OK;
  cudaStream_t st = as_stream(stream);
  int64_t* d_off = nullptr;
  if ((rc = upload_offsets(h_frame_offsets, n_utt, &d_off, st))) return rc;
  double* d_y64 = nullptr;
  cudaError_t e = cudaMallocAsync(&d_y64, sizeof(double) * (size_t)T * dim, st);
  if (e != cudaSuccess) { cudaFreeAsync(d_off, st); return set_error(ODIN_ENOMEM, "RASTA scratch: %s", cudaGetErrorString(e)); }
  const int ld = sdc >= 1 ? dim + dim * dim : dim;
  const int64_t nthr = (int64_t)n_utt * dim;
  feat_rasta_kernel<<<(unsigned)ceil_div<int64_t>(nthr, 128), 128, 0, st>>>(d_x, d_y64, d_y, ld, dim, d_off, n_utt, rasta);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  float* d_dx = nullptr;
  if (sdc >= 1) {
    e = cudaMallocAsync(&d_dx, sizeof(float) * (size_t)T * dim, st);
    if (e == cudaSuccess) {
      const unsigned grid = (unsigned)std::min<int64_t>(T, (int64_t)sm_count() * 16);
      feat_delta1_kernel<<<grid, 256, 0, st>>>(d_y64, d_dx, dim, d_off, n_utt, sdc);
      feat_sdc_kernel<<<grid, 256, 0, st>>>(d_dx, d_y, ld, dim, d_off, n_utt);
      g_launches.fetch_add(2, std::memory_order_relaxed);
    }
  }
  if (e == cudaSuccess) e = cudaGetLastError();
  if (d_dx) cudaFreeAsync(d_dx, st);
  cudaFreeAsync(d_y64, st);
  cudaFreeAsync(d_off, st);
  if (e != cudaSuccess) return set_error(ODIN_ECUDA, "launch feat_rasta/sdc kernel: %s", cudaGetErrorString(e));
  return ODIN_OK;
}

int odin_feat_energy(const float* d_frames, float* d_energy, int64_t n_frames, int32_t frame_len, int32_t take_log,
                     void* stream) {
  if (!d_frames || !d_energy || n_frames < 0 || frame_len <= 0) return set_error(ODIN_EINVAL, "bad argument");
  int rc = require_device();
  if (rc) return rc;
  if (n_frames == 0) return ODIN_OK;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(n_frames, 8), (int64_t)sm_count() * 16);
  feat_energy_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_frames, d_energy, n_frames, frame_len, take_log);
  ODIN_LAUNCH_CHECK("feat_energy_kernel");
  return ODIN_OK;
}

int odin_feat_smooth(const uint8_t* d_x, double* d_y, int64_t n, int32_t win, void* stream) {
  if (!d_x || !d_y || n < 0 || win < 3 || n < win) return set_error(ODIN_EINVAL, "bad argument (need win >= 3 and n >= win)");
  int rc = require_device();
  if (rc) return rc;
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(n, 256), (int64_t)sm_count() * 8);
  feat_smooth_kernel<<<grid, 256, 0, as_stream(stream)>>>(d_x, d_y, n, win);
  ODIN_LAUNCH_CHECK("feat_smooth_kernel");
  return ODIN_OK;
}

}  // extern "C"
