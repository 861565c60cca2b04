// extern "C" entry points of libodin_b200.so (see include/odin_b200.h) and the
// host-side table construction for the front-end.
#include <dlfcn.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "fe.cuh"
#include "fe_frame.cuh"
#include "tmat.cuh"
#include "fe_logic.cuh"
#include "gmm.cuh"

namespace odin {

static thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return set_error(ODIN_ENODEVICE, "no CUDA device available (%s); libodin_b200 has no CPU fallback",
                     e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  }
  // Stream-ordered scratch (cudaMallocAsync in the feature-matrix, CMVN and standalone SAD entry points) comes from
  // the device's default memory pool, which by default hands freed memory back to the driver at the next
  // synchronisation -- a ~1 GB scratch was then re-mapped on most calls (RASTA on 2 M frames: 1.1 ms or 6 ms from
  // one run to the next).  Keep up to 4 GiB cached per device instead.
  static thread_local int pool_dev = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev != pool_dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t keep = (uint64_t)4 << 30, cur = 0;
      if (cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur) == cudaSuccess && cur < keep)
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    cudaGetLastError();
    pool_dev = dev;
  }
  return ODIN_OK;
}

int sm_count() {
  static thread_local int cached_dev = -1, cached = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return cached;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) cached = v;
    cached_dev = dev;
  }
  return cached;
}

// ---------------------------------------------------------------------------
// front-end tables (fp64 on the host, as signal.py builds them)
// ---------------------------------------------------------------------------
static double hz2mel(double f) {  // signal.py:489-527
  const double f_sp = 200.0 / 3.0;
  if (f >= 1000.0) return 1000.0 / f_sp + log(f / 1000.0) / (log(6.4) / 27.0);
  return f / f_sp;
}
static double mel2hz(double m) {  // signal.py:529-568
  const double f_sp = 200.0 / 3.0, min_log_mel = 1000.0 / f_sp;
  if (m >= min_log_mel) return 1000.0 * exp((log(6.4) / 27.0) * (m - min_log_mel));
  return f_sp * m;
}
static std::vector<double> linspace(double a, double b, int n) {  // numpy: start + i*step, last = stop
  std::vector<double> v(n);
  if (n == 1) { v[0] = a; return v; }
  double step = (b - a) / (n - 1);
  for (int i = 0; i < n; ++i) v[i] = a + i * step;
  v[n - 1] = b;
  return v;
}

template <typename T>
static int upload(T** dst, const std::vector<T>& src) {
  size_t bytes = std::max<size_t>(1, src.size()) * sizeof(T);
  ODIN_CUDA_CHECK(cudaMalloc(dst, bytes));
  if (!src.empty()) ODIN_CUDA_CHECK(cudaMemcpy(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice));
  return ODIN_OK;
}

int fe_build_tables(odin_fe* fe) {
  const odin_fe_config& c = fe->cfg;
  const int L = fe->L, N = fe->N, nb = fe->nbins, nm = fe->n_mels;
  // window (scipy.signal.get_window(name, L, fftbins=True); signal.py:812-830)
  fe->h_win.resize(L);
  double wsum = 0.0;
  for (int k = 0; k < L; ++k) {
    double cs = cos(2.0 * M_PI * (double)k / (double)L);
    fe->h_win[k] = c.window == 0 ? 0.5 - 0.5 * cs : 0.54 - 0.46 * cs;
    wsum += fe->h_win[k];
  }
  double scale = sqrt(1.0 / (wsum * wsum));  // signal.py:1547
  fe->scale2 = (float)(scale * scale);
  std::vector<float> win32(L);
  for (int k = 0; k < L; ++k) win32[k] = (float)fe->h_win[k];
  std::vector<float2> tw(N);
  for (int k = 0; k < N; ++k) {
    double ang = -2.0 * M_PI * (double)k / (double)N;
    tw[k] = make_float2((float)cos(ang), (float)sin(ang));
  }
  // four-step FFT twiddles for fe_frame4_kernel: N = G * 32, tw4[k1 * G + l] = exp(-2 pi i l k1 / N)
  std::vector<float2> tw4(N);
  {
    const int G = N / 32;
    for (int k1 = 0; k1 < 32; ++k1)
      for (int l = 0; l < G; ++l) {
        double ang = -2.0 * M_PI * (double)l * (double)k1 / (double)N;
        tw4[(size_t)k1 * G + l] = make_float2((float)cos(ang), (float)sin(ang));
      }
  }
  // mel filterbank (signal.py:735-810), dense fp64 -> CSR fp32
  fe->h_mel.assign((size_t)nm * nb, 0.0);
  std::vector<double> binhz = linspace(0.0, (double)c.sr / 2.0, nb);
  std::vector<double> mels = linspace(hz2mel((double)c.fmin), hz2mel((double)c.fmax), nm + 2);
  std::vector<double> edges(nm + 2);
  for (int i = 0; i < nm + 2; ++i) edges[i] = mel2hz(mels[i]);
  std::vector<int> mstart(nm), mcnt(nm), moff(nm);
  std::vector<float> mw;
  for (int i = 0; i < nm; ++i) {
    double fd0 = edges[i + 1] - edges[i], fd1 = edges[i + 2] - edges[i + 1];
    double enorm = 2.0 / (edges[i + 2] - edges[i]);
    int first = -1, last = -1;
    for (int k = 0; k < nb; ++k) {
      double lower = -(edges[i] - binhz[k]) / fd0;
      double upper = (edges[i + 2] - binhz[k]) / fd1;
      double w = std::max(0.0, std::min(lower, upper)) * enorm;
      fe->h_mel[(size_t)i * nb + k] = w;
      if (w > 0.0) { if (first < 0) first = k; last = k; }
    }
    mstart[i] = first < 0 ? 0 : first;
    mcnt[i] = first < 0 ? 0 : last - first + 1;
    moff[i] = (int)mw.size();
    for (int k = 0; k < mcnt[i]; ++k) mw.push_back((float)fe->h_mel[(size_t)i * nb + mstart[i] + k]);
  }
  fe->mel_nnz = (int)mw.size();
  // Lane-balanced form for fe_frame4_kernel.  Filter widths grow with frequency (4 taps at the bottom, ~60 at
  // the top for n_fft 1024 / 80 mels), so "one filter per lane" makes every warp wait for its widest filter.
  // Here each filter is cut into chunks of at most ceil(nnz / 32) taps, the chunks are dealt to the 32 lanes
  // longest-first onto the least loaded lane, and a lane walks its own tap list; at the end of a chunk it
  // drops the partial sum into the chunk's slot, and filter m adds its slots [ps[m], ps[m+1]) in bin order.
  std::vector<float2> mtab;
  std::vector<int> mps(nm + 1, 0);
  {
    const int cap = std::max(4, (fe->mel_nnz + 31) / 32);
    struct Chunk { int m, start, cnt, slot; };
    std::vector<Chunk> chunks;
    for (int i = 0; i < nm; ++i) {
      mps[i] = (int)chunks.size();
      const int pieces = (mcnt[i] + cap - 1) / cap;
      for (int p = 0; p < pieces; ++p) {
        const int b0 = (int)((int64_t)mcnt[i] * p / pieces), b1 = (int)((int64_t)mcnt[i] * (p + 1) / pieces);
        chunks.push_back({i, mstart[i] + b0, b1 - b0, (int)chunks.size()});
      }
    }
    mps[nm] = (int)chunks.size();
    std::vector<Chunk> order(chunks);
    std::stable_sort(order.begin(), order.end(), [](const Chunk& x, const Chunk& y) { return x.cnt > y.cnt; });
    std::vector<std::vector<float2>> lanes(32);
    for (const Chunk& ch : order) {
      int best = 0;
      for (int l = 1; l < 32; ++l) if (lanes[l].size() < lanes[best].size()) best = l;
      for (int k = 0; k < ch.cnt; ++k) {
        const uint32_t meta = (uint32_t)(ch.start + k) | (k == ch.cnt - 1 ? 0x8000u : 0u) | ((uint32_t)ch.slot << 16);
        float mf;
        memcpy(&mf, &meta, 4);
        lanes[best].push_back(make_float2((float)fe->h_mel[(size_t)ch.m * nb + ch.start + k], mf));
      }
    }
    size_t trips = 0;
    for (int l = 0; l < 32; ++l) trips = std::max(trips, lanes[l].size());
    mtab.assign(std::max<size_t>(1, trips) * 32, make_float2(0.f, 0.f));   // padding: weight 0, bin 0, no flush
    for (int l = 0; l < 32; ++l)
      for (size_t t = 0; t < lanes[l].size(); ++t) mtab[t * 32 + l] = lanes[l][t];
    fe->mel_trips = (int)trips;
    fe->mel_chunks = (int)chunks.size();
    if (nb > 0x7fff || chunks.size() > 0xffff) return set_error(ODIN_EINVAL, "filterbank too large for the packed table");
  }
  // Lane-chunk form for fe_frame5_kernel (n_fft <= 1024).  With centres p_0 .. p_(nm+1), a bin in segment s =
  // [p_s, p_(s+1)) lies on the falling side of filter s-1 and on the rising side of filter s and nowhere else.
  // Lane l of a warp owns the NK = N/64 contiguous bins NK l .. NK l + NK-1 and keeps two accumulators, one for
  // the filter it is on the falling side of and one for the filter it is on the rising side of.  When its next
  // bin lies in the next segment the falling filter is complete for this lane: it is dropped into the next
  // partial-sum slot, the rising accumulator takes the falling role and a new rising one starts at zero; after
  // its last bin the lane drops both.  Filter m then adds its K or fewer slots (listed in bin order in `refs`).
  fe->win_c = (float)(0.5 * scale);
  std::vector<float2> m5w;
  std::vector<uint32_t> m5flags(32, 0u);
  std::vector<uint16_t> m5refs;
  fe->mel5_ok = false;
  fe->mel5_nslots = fe->mel5_k = 0;
  if (N <= 1024) {
    const int NK = N / 64, nhalf = N / 2;
    bool ok = true;
    std::vector<int> seg(nhalf, 0);
    m5w.assign((size_t)NK * 32, make_float2(0.f, 0.f));
    for (int k = 0; k < nhalf && ok; ++k) {
      int s = 0;
      for (int i = 1; i <= nm; ++i) s += edges[i] <= binhz[k];
      seg[k] = s;
      for (int m = 0; m < nm; ++m)
        if (fe->h_mel[(size_t)m * nb + k] != 0.0 && m != s && m != s - 1) ok = false;
      if (k > 0 && seg[k] - seg[k - 1] > 1) ok = false;   // a centre pair without a bin between: not this kernel's case
      const float wfall = s >= 1 ? (float)fe->h_mel[(size_t)(s - 1) * nb + k] : 0.f;
      const float wrise = s < nm ? (float)fe->h_mel[(size_t)s * nb + k] : 0.f;
      m5w[(size_t)(k % NK) * 32 + k / NK] = make_float2(wfall, wrise);
    }
    for (int m = 0; m < nm; ++m) ok = ok && fe->h_mel[(size_t)m * nb + nhalf] == 0.0;   // Nyquist bin: never weighted
    if (ok) {
      std::vector<int> slot_filter;   // filter fed by slot i (-1 / nm: none)
      for (int l = 0; l < 32; ++l) {
        uint32_t fl = 0;
        const uint32_t first = (uint32_t)slot_filter.size();
        for (int j = 0; j < NK; ++j) {
          const int k = NK * l + j;
          if (j == NK - 1) { slot_filter.push_back(seg[k] - 1); slot_filter.push_back(seg[k]); }
          else if (seg[k + 1] != seg[k]) { fl |= 1u << j; slot_filter.push_back(seg[k] - 1); }
        }
        m5flags[l] = fl | (first << 16);
      }
      const int nslots = (int)slot_filter.size();
      std::vector<std::vector<int>> of(nm);
      for (int i = 0; i < nslots; ++i)
        if (slot_filter[i] >= 0 && slot_filter[i] < nm) of[slot_filter[i]].push_back(i);
      size_t K = 1;
      for (int m = 0; m < nm; ++m) K = std::max(K, of[m].size());
      const int rounds = (nm + 31) / 32;
      K = (K + 3) & ~size_t(3);   // four 16-bit slot indices per 8-byte word: [rounds][K / 4][32][4]
      m5refs.assign((size_t)rounds * K * 32, (uint16_t)nslots);   // padding: the zero slot
      for (int m = 0; m < nm; ++m)
        for (size_t i = 0; i < of[m].size(); ++i)
          m5refs[((((size_t)(m / 32) * (K / 4) + i / 4) * 32) + m % 32) * 4 + i % 4] = (uint16_t)of[m][i];
      fe->mel5_nslots = nslots;
      fe->mel5_k = (int)K;
      // the slots (8 B each, + the zero slot) reuse a pair region of the kernel: 8 * f5_region<N>() bytes >= 8 * (N + 65)
      ok = nslots + 1 <= N + 65 && K <= 16;
    }
    fe->mel5_ok = ok;
  }
  // DCT-II orthonormal rows (signal.py:682-733)
  const int nc1 = fe->n_c1;
  fe->h_dct.assign((size_t)std::max(1, nc1) * nm, 0.0);
  std::vector<float> dct32((size_t)std::max(1, nc1) * nm, 0.f);
  for (int i = 0; i < nc1; ++i)
    for (int k = 0; k < nm; ++k) {
      double v = i == 0 ? 1.0 / sqrt((double)nm)
                        : cos((double)i * (double)(2 * k + 1) * M_PI / (2.0 * nm)) * sqrt(2.0 / nm);
      fe->h_dct[(size_t)i * nm + k] = v;
      dct32[(size_t)i * nm + k] = (float)v;
    }
  // delta taps (signal.py:1041-1045): (h, h-1, ..., -h) / sum m^2
  const int W = c.delta_width, h = W / 2;
  double norm = 0.0;
  for (int m = -h; m <= h; ++m) norm += (double)m * m;
  std::vector<float> taps(W);
  for (int k = 0; k < W; ++k) taps[k] = (float)((double)(h - k) / norm);

  int rc;
  if ((rc = upload(&fe->d_win32, win32))) return rc;
  if ((rc = upload(&fe->d_win64, fe->h_win))) return rc;
  if ((rc = upload(&fe->d_tw, tw))) return rc;
  if ((rc = upload(&fe->d_tw4, tw4))) return rc;
  if ((rc = upload(&fe->d_mel_start, mstart))) return rc;
  if ((rc = upload(&fe->d_mel_cnt, mcnt))) return rc;
  if ((rc = upload(&fe->d_mel_off, moff))) return rc;
  if ((rc = upload(&fe->d_mel_w, mw))) return rc;
  if ((rc = upload(&fe->d_mel_tab, mtab))) return rc;
  if ((rc = upload(&fe->d_mel_ps, mps))) return rc;
  if ((rc = upload(&fe->d_mel5_w, m5w))) return rc;
  if ((rc = upload(&fe->d_mel5_flags, m5flags))) return rc;
  if ((rc = upload(&fe->d_mel5_refs, m5refs))) return rc;
  if ((rc = upload(&fe->d_dct, dct32))) return rc;
  if ((rc = upload(&fe->d_taps, taps))) return rc;
  return ODIN_OK;
}

int fe_reserve(odin_fe* fe, int n_utt) {
  if (n_utt <= fe->cap_utt) return ODIN_OK;
  int cap = n_utt + n_utt / 4 + 16;
  auto freeall = [&]() {
    if (fe->h_stage) cudaFreeHost(fe->h_stage);
    cudaFree(fe->d_sample_off);   // one block: d_frame_off / d_tile_off / d_tile2_off / d_vad_order point into it
    cudaFree(fe->d_dcsum); cudaFree(fe->d_umax); cudaFree(fe->d_cnt);
    fe->h_stage = nullptr;
    fe->d_sample_off = fe->d_frame_off = fe->d_tile_off = fe->d_tile2_off = fe->d_vad_order = fe->d_cnt = nullptr;
    fe->d_dcsum = nullptr; fe->d_umax = nullptr; fe->cap_utt = 0;
  };
  freeall();
  size_t n1 = (size_t)cap + 1;
  ODIN_CUDA_CHECK(cudaMallocHost(&fe->h_stage, 5 * n1 * sizeof(int64_t)));
  ODIN_CUDA_CHECK(cudaMalloc(&fe->d_sample_off, 5 * n1 * sizeof(int64_t)));
  fe->d_frame_off = fe->d_sample_off + n1;
  fe->d_tile_off = fe->d_frame_off + n1;
  fe->d_tile2_off = fe->d_tile_off + n1;
  fe->d_vad_order = fe->d_tile2_off + n1;
  ODIN_CUDA_CHECK(cudaMalloc(&fe->d_dcsum, n1 * sizeof(double)));
  ODIN_CUDA_CHECK(cudaMalloc(&fe->d_umax, 2 * n1 * sizeof(int)));
  ODIN_CUDA_CHECK(cudaMalloc(&fe->d_cnt, n1 * sizeof(int64_t)));
  fe->cap_utt = cap;
  return ODIN_OK;
}

static int check_fe_config(const odin_fe_config& c) {
  if (c.sr <= 0 || c.frame_len <= 0 || c.hop <= 0) return set_error(ODIN_EINVAL, "sr/frame_len/hop must be > 0");
  if (c.n_fft != 256 && c.n_fft != 512 && c.n_fft != 1024 && c.n_fft != 2048)
    return set_error(ODIN_EINVAL, "n_fft must be 256, 512, 1024 or 2048 (got %d)", c.n_fft);
  if (c.n_fft < c.frame_len)
    return set_error(ODIN_EINVAL, "n_fft must be >= frame_len (signal.py:1523-1524)");
  if (c.window != 0 && c.window != 1) return set_error(ODIN_EINVAL, "window must be 0 (hann) or 1 (hamming)");
  if (c.n_mels <= 0 || c.n_mels > ODIN_FE_MAX_MELS) return set_error(ODIN_EINVAL, "n_mels must be in 1..128");
  if (!(c.fmin < c.fmax)) return set_error(ODIN_EINVAL, "fmin must < fmax (signal.py:1680-1682)");
  if (c.n_ceps < 0 || c.n_ceps + 1 > c.n_mels) return set_error(ODIN_EINVAL, "n_ceps must be in 0..n_mels-1");
  if (c.delta_order < 0 || c.delta_order > 2) return set_error(ODIN_EINVAL, "delta_order must be 0, 1 or 2");
  if (c.delta_width < 3 || (c.delta_width & 1) == 0 || c.delta_width > 31)
    return set_error(ODIN_EINVAL, "delta_width must be an odd integer in 3..31 (signal.py:1034-1035)");
  if (c.vad_kind < 0 || c.vad_kind > 2) return set_error(ODIN_EINVAL, "vad_kind must be 0, 1 or 2");
  if (c.vad_kind == 1 && (c.vad_nmix < 2 || c.vad_nmix > 4)) return set_error(ODIN_EINVAL, "vad_nmix must be 2..4");
  if (c.vad_kind == 2 && c.n_ceps <= 0) return set_error(ODIN_EINVAL, "SADthreshold runs on c0: needs n_ceps > 0");
  return ODIN_OK;
}

}  // namespace odin

using namespace odin;

extern "C" {

const char* odin_last_error(void) { return g_err.c_str(); }
int odin_version(void) { return 1000 * 0 + 1; }
int64_t odin_launch_count(void) { return g_launches.load(); }

// ------------------------------- front-end ---------------------------------
int odin_fe_create(const odin_fe_config* cfg, odin_fe_t** out) {
  if (!cfg || !out) return set_error(ODIN_EINVAL, "null argument");
  *out = nullptr;
  int rc = check_fe_config(*cfg);
  if (rc) return rc;
  if ((rc = require_device())) return rc;
  odin_fe* fe = new (std::nothrow) odin_fe();
  if (!fe) return set_error(ODIN_ENOMEM, "out of host memory");
  fe->cfg = *cfg;
  fe->L = cfg->frame_len; fe->hop = cfg->hop; fe->N = cfg->n_fft; fe->nbins = cfg->n_fft / 2 + 1;
  fe->pad = cfg->padding ? cfg->frame_len / 2 : 0;
  fe->n_mels = cfg->n_mels;
  fe->n_c1 = cfg->n_ceps > 0 ? cfg->n_ceps + 1 : 0;
  fe->feat_dim = cfg->n_ceps * (1 + cfg->delta_order);
  rc = fe_build_tables(fe);
  if (rc) { odin_fe_destroy(fe); return rc; }
  *out = fe;
  return ODIN_OK;
}

void odin_fe_destroy(odin_fe_t* fe) {
  if (!fe) return;
  cudaFree(fe->d_win32); cudaFree(fe->d_win64); cudaFree(fe->d_tw); cudaFree(fe->d_tw4); cudaFree(fe->d_mel_start);
  cudaFree(fe->d_mel_cnt); cudaFree(fe->d_mel_off); cudaFree(fe->d_mel_w); cudaFree(fe->d_dct);
  cudaFree(fe->d_mel_tab); cudaFree(fe->d_mel_ps);
  cudaFree(fe->d_dct64); cudaFree(fe->d_tile_ctr); cudaFree(fe->d_mel5_w); cudaFree(fe->d_mel5_flags); cudaFree(fe->d_mel5_refs);
  cudaFree(fe->d_taps); cudaFree(fe->d_sample_off); cudaFree(fe->d_dcsum); cudaFree(fe->d_umax);
  cudaFree(fe->d_cnt); cudaFree(fe->d_vad_scratch);
  if (fe->h_stage) cudaFreeHost(fe->h_stage);
  for (int i = 0; i < 7; ++i) if (fe->ev[i]) cudaEventDestroy(fe->ev[i]);
  if (fe->aux) cudaStreamDestroy(fe->aux);
  delete fe;
}

int odin_fe_feat_dim(const odin_fe_t* fe) { return fe ? fe->feat_dim : ODIN_EINVAL; }

static int frame_offsets_impl(int L, int hop, const int64_t* so, int n_utt, int64_t* fo, int pad = 0) {
  int rc = ODIN_OK;
  fo[0] = 0;
  for (int u = 0; u < n_utt; ++u) {
    int64_t n = so[u + 1] - so[u] + 2 * (int64_t)pad;   // signal.py:1529-1530
    int64_t T = n < L ? 0 : 1 + (n - L) / hop;  // signal.py:1532-1538
    if (n < L) rc = ODIN_ESHORT;
    fo[u + 1] = fo[u] + T;
  }
  return rc;
}

int odin_fe_frame_offsets(const odin_fe_t* fe, const int64_t* h_sample_offsets, int32_t n_utt,
                          int64_t* h_frame_offsets) {
  if (!fe || !h_sample_offsets || !h_frame_offsets || n_utt < 0) return set_error(ODIN_EINVAL, "bad argument");
  int rc = frame_offsets_impl(fe->L, fe->hop, h_sample_offsets, n_utt, h_frame_offsets, fe->pad);
  if (rc == ODIN_ESHORT) set_error(rc, "an utterance is shorter than one frame (%d samples)", fe->L);
  return rc;
}

int odin_host_frame_offsets(int32_t frame_len, int32_t hop, const int64_t* h_sample_offsets, int32_t n_utt,
                            int64_t* h_frame_offsets) {
  if (frame_len <= 0 || hop <= 0 || !h_sample_offsets || !h_frame_offsets || n_utt < 0)
    return set_error(ODIN_EINVAL, "bad argument");
  int rc = frame_offsets_impl(frame_len, hop, h_sample_offsets, n_utt, h_frame_offsets);
  if (rc == ODIN_ESHORT) set_error(rc, "an utterance is shorter than one frame (%d samples)", frame_len);
  return rc;
}

int odin_host_smooth(const uint8_t* x, int32_t n, int32_t win, int32_t wrap_u8, uint8_t* out) {
  if (!x || !out || n < 0) return set_error(ODIN_EINVAL, "bad argument");
  if (win < 3 || n < win) { for (int i = 0; i < n; ++i) out[i] = x[i]; return ODIN_OK; }
  auto get = [x](int i) -> int { return x[i] ? 1 : 0; };
  for (int t = 0; t < n; ++t) out[t] = (uint8_t)smooth_flat_ge_f(get, n, win, wrap_u8 != 0, t);
  return ODIN_OK;
}

// numpy-exact float32 mean/std (signal.py:305), exposed for the CPU tests
int odin_host_mean_std_f32(const float* e, int32_t n, float* mean, float* std_) {
  if (!e || n <= 0 || !mean || !std_) return set_error(ODIN_EINVAL, "bad argument");
  MeanStdF32 r = np_mean_std_f32(e, n);
  *mean = r.mean;
  *std_ = r.std;
  return ODIN_OK;
}

// dense fp64 tables as built for the device (tests compare them with the oracle)
int odin_fe_get_table(const odin_fe_t* fe, int32_t which, double* out, int64_t cap) {
  if (!fe || !out) return set_error(ODIN_EINVAL, "bad argument");
  const std::vector<double>* v = which == 0 ? &fe->h_win : which == 1 ? &fe->h_mel : which == 2 ? &fe->h_dct : nullptr;
  if (!v) return set_error(ODIN_EINVAL, "which must be 0 (window), 1 (mel), 2 (dct)");
  if ((int64_t)v->size() > cap) return set_error(ODIN_EINVAL, "buffer too small: need %zu", v->size());
  memcpy(out, v->data(), v->size() * sizeof(double));
  return (int)v->size();
}

int odin_fe_run(odin_fe_t* fe, const void* d_pcm, int32_t pcm_dtype, const int64_t* h_sample_offsets,
                int32_t n_utt, float* d_mspec, float* d_feat, float* d_energy, float* d_c0, uint8_t* d_sad,
                double* d_sad_thr, void* stream) {
  return odin_fe_run_spectra(fe, d_pcm, pcm_dtype, h_sample_offsets, n_utt, d_mspec, d_feat, d_energy, d_c0, d_sad,
                             d_sad_thr, nullptr, 0, stream);
}

// Host prelude shared by the entry points that walk a ragged batch of PCM: frame / tile offsets (integer exact,
// signal.py:1532-1538) and the SADgmm visiting order into the pinned staging block, one upload.
static int fe_prepare(odin_fe_t* fe, const int64_t* h_sample_offsets, int32_t n_utt, cudaStream_t st,
                      int64_t* total_frames, int64_t* n_tiles, int64_t* n_tiles2) {
  int rc = fe_reserve(fe, n_utt);
  if (rc) return rc;
  const size_t n1 = (size_t)fe->cap_utt + 1;
  int64_t* so = fe->h_stage;
  int64_t* fo = so + n1;
  int64_t* t1 = fo + n1;
  int64_t* t2 = t1 + n1;
  // the pinned staging block may still be in flight from the previous call on this stream
  ODIN_CUDA_CHECK(cudaStreamSynchronize(st));
  memcpy(so, h_sample_offsets, sizeof(int64_t) * (n_utt + 1));
  rc = frame_offsets_impl(fe->L, fe->hop, so, n_utt, fo, fe->pad);
  if (rc == ODIN_ESHORT) return set_error(rc, "an utterance is shorter than one frame (%d samples)", fe->L);
  t1[0] = t2[0] = 0;
  for (int u = 0; u < n_utt; ++u) {
    int64_t T = fo[u + 1] - fo[u];
    t1[u + 1] = t1[u] + ceil_div<int64_t>(T, ODIN_FE_TILE);
    t2[u + 1] = t2[u] + ceil_div<int64_t>(T, ODIN_FE_POST_TILE);
  }
  if (t1[n_utt] > 0x7fffffff) return set_error(ODIN_EINVAL, "batch too large (tiles)");
  {
    // Order in which the SADgmm kernel visits utterances (one cluster each): longest first, so the
    // hardware's in-order dispatch of clusters to freed SMs is longest-processing-time-first list scheduling
    // (ODIN_FE_VAD_LPT=0 restores the earlier static pairing of long with short utterances for A/B runs).
    int64_t* ord = t2 + n1;
    std::vector<int> idx(n_utt);
    for (int u = 0; u < n_utt; ++u) idx[u] = u;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return fo[a + 1] - fo[a] > fo[b + 1] - fo[b]; });
    const char* lpt = getenv("ODIN_FE_VAD_LPT");
    const int S = (lpt && lpt[0] == '0') ? std::min(n_utt, sm_count()) : 0;
    for (int i = 0; i < S; ++i) ord[i] = idx[S - 1 - i];
    for (int i = S; i < n_utt; ++i) ord[i] = idx[i];
  }
  // Only the used prefix of each of the five arrays travels: copying the whole staging block (capacity of the largest
  // batch this handle has seen: 560 KB after a 100 h batch) needs the DMA engine, where it queues behind the multi-MB
  // PCM copy of the NEXT chunk of a pipelined caller (run_host_packed) and stalls this chunk's kernels behind it --
  // the end-to-end front-end ran at 99 M instead of 158 M frames/s.  A few hundred bytes go inline.
  for (int k = 0; k < 5; ++k)
    ODIN_CUDA_CHECK(cudaMemcpyAsync(fe->d_sample_off + k * n1, fe->h_stage + k * n1, sizeof(int64_t) * (size_t)(n_utt + 1),
                                    cudaMemcpyHostToDevice, st));
  *total_frames = fo[n_utt]; *n_tiles = t1[n_utt]; *n_tiles2 = t2[n_utt];
  return ODIN_OK;
}

int odin_fe_run_spectra(odin_fe_t* fe, const void* d_pcm, int32_t pcm_dtype, const int64_t* h_sample_offsets,
                        int32_t n_utt, float* d_mspec, float* d_feat, float* d_energy, float* d_c0, uint8_t* d_sad,
                        double* d_sad_thr, float* d_spec, int32_t spec_log, void* stream) {
  if (!fe || !d_pcm || !h_sample_offsets || n_utt < 0) return set_error(ODIN_EINVAL, "bad argument");
  if (pcm_dtype != 0 && pcm_dtype != 1) return set_error(ODIN_EINVAL, "pcm_dtype must be 0 (int16) or 1 (float32)");
  if (!d_mspec) return set_error(ODIN_EINVAL, "d_mspec is required (scratch for the utterance pass)");
  if (n_utt == 0) return ODIN_OK;
  cudaStream_t st = as_stream(stream);
  int64_t T = 0, nt1 = 0, nt2 = 0;
  int rc = fe_prepare(fe, h_sample_offsets, n_utt, st, &T, &nt1, &nt2);
  if (rc) return rc;
  return fe_launch(fe, d_pcm, pcm_dtype, n_utt, T, nt1, nt2, d_mspec, d_feat, d_energy, d_c0, d_sad, d_sad_thr, d_spec,
                   spec_log, st);
}

int odin_fe_stft(odin_fe_t* fe, const void* d_pcm, int32_t pcm_dtype, const int64_t* h_sample_offsets, int32_t n_utt,
                 void* d_stft, float* d_energy, void* stream) {
  if (!fe || !d_pcm || !h_sample_offsets || n_utt < 0 || !d_stft) return set_error(ODIN_EINVAL, "bad argument");
  if (pcm_dtype != 0 && pcm_dtype != 1) return set_error(ODIN_EINVAL, "pcm_dtype must be 0 (int16) or 1 (float32)");
  if (n_utt == 0) return ODIN_OK;
  cudaStream_t st = as_stream(stream);
  int64_t T = 0, nt1 = 0, nt2 = 0;
  int rc = fe_prepare(fe, h_sample_offsets, n_utt, st, &T, &nt1, &nt2);
  if (rc) return rc;
  if (T == 0) return ODIN_OK;
  // the frame kernels always project onto the handle's filterbank: give them a scratch for the rows nobody asked for
  float* d_mspec = nullptr;
  ODIN_CUDA_CHECK(cudaMallocAsync(&d_mspec, sizeof(float) * (size_t)T * fe->n_mels, st));
  rc = fe_launch(fe, d_pcm, pcm_dtype, n_utt, T, nt1, nt2, d_mspec, nullptr, d_energy, nullptr, nullptr, nullptr, nullptr, 0, st,
                 reinterpret_cast<float2*>(d_stft));
  cudaFreeAsync(d_mspec, st);
  return rc;
}

int odin_fe_frames(odin_fe_t* fe, const void* d_pcm, int32_t pcm_dtype, const int64_t* h_sample_offsets, int32_t n_utt,
                   float* d_frames, float* d_energy, void* stream) {
  if (!fe || !d_pcm || !h_sample_offsets || n_utt < 0 || (!d_frames && !d_energy))
    return set_error(ODIN_EINVAL, "bad argument");
  if (pcm_dtype != 0 && pcm_dtype != 1) return set_error(ODIN_EINVAL, "pcm_dtype must be 0 (int16) or 1 (float32)");
  if (n_utt == 0) return ODIN_OK;
  cudaStream_t st = as_stream(stream);
  int64_t T = 0, nt1 = 0, nt2 = 0;
  int rc = fe_prepare(fe, h_sample_offsets, n_utt, st, &T, &nt1, &nt2);
  if (rc) return rc;
  return fe_frames_launch(fe, d_pcm, pcm_dtype, n_utt, T, nt1, d_frames, d_energy, st);
}

int odin_vad_gmm(const float* d_energy, const int64_t* h_frame_offsets, int32_t n_utt, int32_t nb_mixture,
                 int32_t nb_train_it, int32_t smooth_window, float mode, uint8_t* d_sad, double* d_threshold,
                 void* stream) {
  if (!d_energy || !h_frame_offsets || !d_sad || n_utt < 0) return set_error(ODIN_EINVAL, "bad argument");
  if (nb_mixture < 2 || nb_mixture > 4) return set_error(ODIN_EINVAL, "nb_mixture must be 2..4");
  int rc = require_device();
  if (rc) return rc;
  if (n_utt == 0) return ODIN_OK;
  return fe_vad_standalone(1, d_energy, h_frame_offsets, n_utt, nb_mixture, nb_train_it, smooth_window, (double)mode, 0, 0,
                           0, 0, d_sad, d_threshold, as_stream(stream));
}

int odin_vad_threshold(const float* d_energy, const int64_t* h_frame_offsets, int32_t n_utt, float energy_threshold,
                       float energy_mean_scale, int32_t frame_context, float proportion_threshold,
                       int32_t smooth_window, uint8_t* d_sad, double* d_threshold, void* stream) {
  if (!d_energy || !h_frame_offsets || !d_sad || n_utt < 0) return set_error(ODIN_EINVAL, "bad argument");
  int rc = require_device();
  if (rc) return rc;
  if (n_utt == 0) return ODIN_OK;
  return fe_vad_standalone(2, d_energy, h_frame_offsets, n_utt, 3, 25, smooth_window, 2.0, (double)energy_threshold,
                           (double)energy_mean_scale, (double)proportion_threshold, frame_context, d_sad, d_threshold,
                           as_stream(stream));
}

int odin_fe_compact(odin_fe_t* fe, const uint8_t* d_sad, const int64_t* h_frame_offsets, int32_t n_utt,
                    const float* d_feat, int32_t dim, int32_t keep_unvoiced, float* d_out, int64_t* d_out_offsets,
                    void* stream) {
  if (!fe || !d_sad || !h_frame_offsets || !d_feat || !d_out || !d_out_offsets || dim <= 0 || n_utt < 0)
    return set_error(ODIN_EINVAL, "bad argument");
  if (n_utt == 0) return ODIN_OK;
  int rc = fe_reserve(fe, n_utt);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const size_t n1 = (size_t)fe->cap_utt + 1;
  ODIN_CUDA_CHECK(cudaStreamSynchronize(st));
  memcpy(fe->h_stage + n1, h_frame_offsets, sizeof(int64_t) * (n_utt + 1));
  ODIN_CUDA_CHECK(cudaMemcpyAsync(fe->d_frame_off, fe->h_stage + n1, sizeof(int64_t) * (n_utt + 1),
                                  cudaMemcpyHostToDevice, st));
  return fe_compact_launch(fe, d_sad, n_utt, d_feat, dim, keep_unvoiced, d_out, d_out_offsets, st);
}

int odin_fe_cmvn(const float* d_x, float* d_y, int32_t dim, const int64_t* h_frame_offsets, int32_t n_utt,
                 const uint8_t* d_sad, int32_t mean_var_norm, int32_t var_norm, int32_t windowed, int32_t win_length,
                 void* stream) {
  if (!d_x || !d_y || !h_frame_offsets || dim <= 0 || dim > 256 || n_utt < 0)
    return set_error(ODIN_EINVAL, "bad argument (dim must be 1..256)");
  if (d_x == d_y && windowed) return set_error(ODIN_EINVAL, "windowed normalisation cannot run in place");
  if (windowed && (win_length < 3 || (win_length & 1) == 0))
    return set_error(ODIN_EINVAL, "Window length should be an odd integer >= 3");   // signal.py:896-897
  int rc = require_device();
  if (rc) return rc;
  if (n_utt == 0) return ODIN_OK;
  cudaStream_t st = as_stream(stream);
  int64_t* d_off = nullptr;
  ODIN_CUDA_CHECK(cudaMallocAsync(&d_off, sizeof(int64_t) * (n_utt + 1), st));
  cudaError_t e = cudaMemcpyAsync(d_off, h_frame_offsets, sizeof(int64_t) * (n_utt + 1), cudaMemcpyHostToDevice, st);
  if (e != cudaSuccess) { cudaFreeAsync(d_off, st); return set_error(ODIN_ECUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e)); }
  rc = fe_cmvn_launch(d_x, d_y, dim, d_off, n_utt, d_sad, mean_var_norm, var_norm, windowed, win_length, st);
  cudaFreeAsync(d_off, st);
  return rc;
}

// --------------------------------- GMM --------------------------------------
int odin_gmm_create(int32_t feat_dim, int32_t max_nmix, odin_gmm_t** out) {
  if (!out || feat_dim <= 0 || feat_dim > 127 || max_nmix <= 0 || max_nmix > 65536)
    return set_error(ODIN_EINVAL, "feat_dim must be 1..127 and max_nmix 1..65536");
  *out = nullptr;
  int rc = require_device();
  if (rc) return rc;
  odin_gmm* g = new (std::nothrow) odin_gmm();
  if (!g) return set_error(ODIN_ENOMEM, "out of host memory");
  g->D = feat_dim;
  g->max_nmix = max_nmix;
  cudaGetDevice(&g->device);
  const size_t maxpad = (size_t)ceil_div(max_nmix, 128) * 128;
  const size_t dm = (size_t)feat_dim * max_nmix;
  auto fail = [&](cudaError_t e) {
    odin_gmm_destroy(g);
    return set_error(ODIN_ECUDA, "cudaMalloc: %s", cudaGetErrorString(e));
  };
  cudaError_t e;
  if ((e = cudaMalloc(&g->d_mean, dm * sizeof(float))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&g->d_var, dm * sizeof(float))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&g->d_w, max_nmix * sizeof(float))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&g->d_Wk, 2 * (size_t)feat_dim * maxpad * sizeof(float))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&g->d_cst, maxpad * sizeof(float))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&g->d_Whi, maxpad * 128 * sizeof(float))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&g->d_Whs, maxpad * 128 * sizeof(float))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&g->d_Wlo, maxpad * 128 * sizeof(float))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&g->d_prev, (2 * dm + max_nmix) * sizeof(float))) != cudaSuccess) return fail(e);
  *out = g;
  return ODIN_OK;
}

void odin_gmm_destroy(odin_gmm_t* g) {
  if (!g) return;
  cudaFree(g->d_mean); cudaFree(g->d_var); cudaFree(g->d_w); cudaFree(g->d_Wk); cudaFree(g->d_cst);
  cudaFree(g->d_Whi); cudaFree(g->d_Whs); cudaFree(g->d_Wlo); cudaFree(g->d_part); cudaFree(g->d_utt_acc); cudaFree(g->d_segX); cudaFree(g->d_segmask); cudaFree(g->d_segtile); cudaFree(g->d_segoff); cudaFree(g->d_lse); cudaFree(g->d_prev); cudaFree(g->d_off);
  gmm_h_free(g);
  if (g->h_off) cudaFreeHost(g->h_off);
  for (int i = 0; i < 3; ++i) if (g->ev[i]) cudaEventDestroy(g->ev[i]);
  delete g;
}

int64_t odin_gmm_stats_size(const odin_gmm_t* g, int32_t nmix) {
  if (!g || nmix <= 0) return ODIN_EINVAL;
  return stats_size(g->D, nmix);
}

int odin_gmm_set_params(odin_gmm_t* g, int32_t nmix, const float* d_mean, const float* d_var, const float* d_w,
                        void* stream) {
  if (!g || !d_mean || !d_var || !d_w) return set_error(ODIN_EINVAL, "null argument");
  if (nmix <= 0 || nmix > g->max_nmix) return set_error(ODIN_EINVAL, "nmix %d outside 1..%d", nmix, g->max_nmix);
  cudaStream_t st = as_stream(stream);
  g->M = nmix;
  g->Mpad = ceil_div(nmix, 128) * 128;
  const size_t dm = (size_t)g->D * nmix;
  ODIN_CUDA_CHECK(cudaMemcpyAsync(g->d_mean, d_mean, dm * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ODIN_CUDA_CHECK(cudaMemcpyAsync(g->d_var, d_var, dm * sizeof(float), cudaMemcpyDeviceToDevice, st));
  ODIN_CUDA_CHECK(cudaMemcpyAsync(g->d_w, d_w, nmix * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return gmm_refresh_constants(g, st);
}

// impl: 0 auto, 1 fp32 CUDA cores, 2 3xTF32 tcgen05, 3 3xFP16 tcgen05
static int pick_impl(odin_gmm* g, int impl, int* use) {
  if (impl < 0 || impl > 3)
    return set_error(ODIN_EINVAL, "impl must be 0 (auto), 1 (fp32), 2 (tcgen05 3xTF32) or 3 (tcgen05 3xFP16)");
  const bool ok2 = gmm_tc_supported(g), ok3 = gmm_h_supported(g);
  if (impl == 2 && !ok2) return set_error(ODIN_EINVAL, "tcgen05 3xTF32 path unsupported for D=%d M=%d", g->D, g->M);
  if (impl == 3 && !ok3) return set_error(ODIN_EINVAL, "tcgen05 3xFP16 path unsupported for D=%d M=%d", g->D, g->M);
  *use = impl != 0 ? impl : (ok3 ? 3 : (ok2 ? 2 : 1));
  return ODIN_OK;
}

int odin_gmm_estep(odin_gmm_t* g, const float* d_X, const uint8_t* d_sad, int64_t n_frames, int32_t want_second,
                   double* d_stats, int32_t impl, void* stream) {
  if (!g || !d_X || !d_stats || n_frames < 0) return set_error(ODIN_EINVAL, "bad argument");
  if (g->M <= 0) return set_error(ODIN_EINVAL, "odin_gmm_set_params has not been called");
  if (n_frames == 0) return ODIN_OK;
  int use;
  int rc = pick_impl(g, impl, &use);
  if (rc) return rc;
  const bool tc = use == 2;
  cudaStream_t st = as_stream(stream);
  if (g->ev[0] == nullptr)
    for (int i = 0; i < 3; ++i) ODIN_CUDA_CHECK(cudaEventCreate(&g->ev[i]));
  if (use == 3) {
    if ((rc = gmm_estep_h(g, d_X, d_sad, n_frames, want_second, d_stats, st))) return rc;
    g->ev_valid = true;
    g->last_impl = 3;
    return ODIN_OK;
  }
  if ((rc = gmm_reserve_lse(g, n_frames))) return rc;
  ODIN_CUDA_CHECK(cudaEventRecord(g->ev[0], st));
  if (tc) rc = gmm_lse_tc(g, d_X, d_sad, n_frames, g->d_lse, d_stats, st);
  else rc = gmm_lse_ffma(g, d_X, d_sad, n_frames, g->d_lse, d_stats, st);
  if (rc) return rc;
  ODIN_CUDA_CHECK(cudaEventRecord(g->ev[1], st));
  if (tc) rc = gmm_stats_tc(g, d_X, d_sad, n_frames, g->d_lse, want_second, d_stats, st);
  else rc = gmm_stats_ffma(g, d_X, d_sad, n_frames, g->d_lse, want_second, d_stats, st);
  if (rc) return rc;
  ODIN_CUDA_CHECK(cudaEventRecord(g->ev[2], st));
  g->ev_valid = true;
  g->last_impl = tc ? 2 : 1;
  g->last_frames = n_frames;
  return ODIN_OK;
}

int odin_gmm_frames_create(odin_gmm_t* g, const float* d_X, int64_t n_frames, odin_gmm_frames_t** out, void* stream) {
  if (!g || !d_X || !out || n_frames <= 0) return set_error(ODIN_EINVAL, "bad argument");
  return gmm_frames_create(g, d_X, n_frames, reinterpret_cast<void**>(out), as_stream(stream));
}

void odin_gmm_frames_destroy(odin_gmm_frames_t* f) { gmm_frames_destroy(f); }

int odin_gmm_estep_frames(odin_gmm_t* g, const odin_gmm_frames_t* f, const uint8_t* d_sad, int32_t want_second,
                          double* d_stats, void* stream) {
  if (!g || !f || !d_stats) return set_error(ODIN_EINVAL, "bad argument");
  if (g->M <= 0) return set_error(ODIN_EINVAL, "odin_gmm_set_params has not been called");
  if (g->ev[0] == nullptr)
    for (int i = 0; i < 3; ++i) ODIN_CUDA_CHECK(cudaEventCreate(&g->ev[i]));
  int rc = gmm_estep_frames(g, f, d_sad, want_second, d_stats, as_stream(stream));
  if (rc) return rc;
  g->ev_valid = true;
  g->last_impl = 3;
  return ODIN_OK;
}

int64_t odin_gmm_last_estep_frames(const odin_gmm_t* g) { return g ? g->last_frames : ODIN_EINVAL; }

// The one exchange step of the multi-GPU path for a C / C++ binder that owns an NCCL communicator:
// ncclAllReduce(sum, double) of the packed statistics, in place and in stream.  NCCL is not linked: it is looked
// up at run time (the library the host application -- e.g. torch -- has already loaded), so the shared object has no
// NCCL dependency when the caller does its collectives elsewhere (the Python binding uses torch.distributed).
int odin_gmm_allreduce(odin_gmm_t* g, double* d_stats, void* nccl_comm, void* stream) {
  if (!g || !d_stats || !nccl_comm) return set_error(ODIN_EINVAL, "bad argument");
  if (g->M <= 0) return set_error(ODIN_EINVAL, "odin_gmm_set_params has not been called");
  typedef int (*allreduce_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
  typedef const char* (*errstr_fn)(int);
  static allreduce_fn fn = nullptr;
  static errstr_fn es = nullptr;
  if (fn == nullptr) {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(ODIN_ENODEVICE, "odin_gmm_allreduce: libnccl.so.2 cannot be loaded (%s)", dlerror());
    fn = reinterpret_cast<allreduce_fn>(dlsym(h, "ncclAllReduce"));
    es = reinterpret_cast<errstr_fn>(dlsym(h, "ncclGetErrorString"));
    if (!fn) return set_error(ODIN_ENODEVICE, "odin_gmm_allreduce: ncclAllReduce not found in libnccl");
  }
  const size_t n = (size_t)stats_size(g->D, g->M);
  const int rc = fn(d_stats, d_stats, n, /*ncclFloat64*/ 8, /*ncclSum*/ 0, nccl_comm, as_stream(stream));
  if (rc != 0) return set_error(ODIN_ECUDA, "ncclAllReduce: %s", es ? es(rc) : "error");
  return ODIN_OK;
}

int odin_gmm_last_estep_ms(odin_gmm_t* g, float* lse_ms, float* stats_ms, int32_t* impl_used) {
  if (!g || !lse_ms || !stats_ms) return set_error(ODIN_EINVAL, "null argument");
  if (!g->ev_valid) return set_error(ODIN_EINVAL, "no E-step has been recorded");
  ODIN_CUDA_CHECK(cudaEventSynchronize(g->ev[2]));
  ODIN_CUDA_CHECK(cudaEventElapsedTime(lse_ms, g->ev[0], g->ev[1]));
  ODIN_CUDA_CHECK(cudaEventElapsedTime(stats_ms, g->ev[1], g->ev[2]));
  if (impl_used) *impl_used = g->last_impl;
  return ODIN_OK;
}

int odin_fe_last_run_ms(odin_fe_t* fe, float* ms4) {
  if (!fe || !ms4) return set_error(ODIN_EINVAL, "null argument");
  if (!fe->ev_valid) return set_error(ODIN_EINVAL, "no run has been recorded");
  ODIN_CUDA_CHECK(cudaEventSynchronize(fe->ev[4]));
  for (int i = 0; i < 4; ++i) ODIN_CUDA_CHECK(cudaEventElapsedTime(ms4 + i, fe->ev[i], fe->ev[i + 1]));
  if (fe->vad_forked) ODIN_CUDA_CHECK(cudaEventElapsedTime(ms4 + 3, fe->ev[5], fe->ev[6]));
  return ODIN_OK;
}

int odin_gmm_mstep(odin_gmm_t* g, const double* d_stats, int32_t allow_rollback, float* d_mean, float* d_var,
                   float* d_w, int32_t* d_rolled_back, void* stream) {
  if (!g || !d_stats) return set_error(ODIN_EINVAL, "null argument");
  if (g->M <= 0) return set_error(ODIN_EINVAL, "odin_gmm_set_params has not been called");
  cudaStream_t st = as_stream(stream);
  int rc = gmm_mstep_launch(g, d_stats, allow_rollback, d_rolled_back, st);
  if (rc) return rc;
  const size_t dm = (size_t)g->D * g->M;
  if (d_mean) ODIN_CUDA_CHECK(cudaMemcpyAsync(d_mean, g->d_mean, dm * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (d_var) ODIN_CUDA_CHECK(cudaMemcpyAsync(d_var, g->d_var, dm * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (d_w) ODIN_CUDA_CHECK(cudaMemcpyAsync(d_w, g->d_w, g->M * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return gmm_refresh_constants(g, st);
}

int odin_gmm_mixup(odin_gmm_t* g, int32_t new_nmix, float* d_mean, float* d_var, float* d_w, void* stream) {
  if (!g) return set_error(ODIN_EINVAL, "null argument");
  if (g->M <= 0) return set_error(ODIN_EINVAL, "odin_gmm_set_params has not been called");
  if (new_nmix <= g->M || new_nmix > 2 * g->M || new_nmix > g->max_nmix)
    return set_error(ODIN_EINVAL, "new_nmix %d must be in (%d, min(%d, %d)]", new_nmix, g->M, 2 * g->M, g->max_nmix);
  cudaStream_t st = as_stream(stream);
  int rc = gmm_mixup_launch(g, new_nmix, st);
  if (rc) return rc;
  g->M = new_nmix;
  g->Mpad = ceil_div(new_nmix, 128) * 128;
  const size_t dm = (size_t)g->D * g->M;
  if (d_mean) ODIN_CUDA_CHECK(cudaMemcpyAsync(d_mean, g->d_mean, dm * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (d_var) ODIN_CUDA_CHECK(cudaMemcpyAsync(d_var, g->d_var, dm * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (d_w) ODIN_CUDA_CHECK(cudaMemcpyAsync(d_w, g->d_w, g->M * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return gmm_refresh_constants(g, st);
}

int odin_gmm_utt_stats(odin_gmm_t* g, const float* d_X, const uint8_t* d_sad, const int64_t* h_frame_offsets,
                       int32_t n_utt, float* d_Z, float* d_Fhat, int32_t impl, void* stream) {
  if (!g || !d_X || !h_frame_offsets || !d_Z || !d_Fhat || n_utt < 0) return set_error(ODIN_EINVAL, "bad argument");
  if (g->M <= 0) return set_error(ODIN_EINVAL, "odin_gmm_set_params has not been called");
  if (n_utt == 0) return ODIN_OK;
  cudaStream_t st = as_stream(stream);
  {
    // Tensor-core-sized models (D % 4 == 0, D <= 60) go through the tcgen05 3xFP16 kernels in SEGMENTED mode: the
    // utterances of the batch back to back, each padded to whole 64-frame tiles, one launch sequence for all of them, the
    // accumulator drained into the utterance's rows wherever the utterance changes (gmm_utt_stats_hseg).  Measured on
    // B200 against the fp32 CUDA-core kernels / the earlier one-launch-sequence-per-utterance route (gmm_utt_stats_h, kept
    // behind ODIN_GMM_UTT_SEG=0): 3 000 digits of 60-200 frames at M = 512: 6.1 -> 1.6 ms; 1 000 utterances of 500-6 000
    // frames at M = 2048: 153.7 / 56.5 -> 12.3 ms; 200 of 6 000-18 000 frames: 115.5 / 17.0 -> 8.5 ms.
    // Tiny batches stay on the fp32 kernels (the tensor route's fixed cost is a handful of extra launches).
    int use = 1;
    if (impl == 0 || impl == 3) {
      int rc = pick_impl(g, impl, &use);
      if (rc) return rc;
    }
    const int64_t total = h_frame_offsets[n_utt] - h_frame_offsets[0];
    static const int seg_env = [] { const char* e = getenv("ODIN_GMM_UTT_SEG"); return e ? atoi(e) : -1; }();
    if (use == 3 && seg_env == 0 && total >= (int64_t)1024 * n_utt) {
      ODIN_CUDA_CHECK(cudaStreamSynchronize(st));
      return gmm_utt_stats_h(g, d_X + h_frame_offsets[0] * g->D, d_sad ? d_sad + h_frame_offsets[0] : nullptr,
                             h_frame_offsets, n_utt, d_Z, d_Fhat, st);
    }
    if (use == 3 && seg_env != 0 && (impl == 3 || seg_env == 1 || total >= 16384)) {
      ODIN_CUDA_CHECK(cudaStreamSynchronize(st));
      return gmm_utt_stats_hseg(g, d_X + h_frame_offsets[0] * g->D, d_sad ? d_sad + h_frame_offsets[0] : nullptr, h_frame_offsets,
                                n_utt, d_Z, d_Fhat, st);
    }
  }
  if (g->off_cap < n_utt + 1) {
    if (g->h_off) cudaFreeHost(g->h_off);
    cudaFree(g->d_off);
    g->h_off = nullptr; g->d_off = nullptr; g->off_cap = 0;
    int64_t cap = n_utt + n_utt / 4 + 16;
    ODIN_CUDA_CHECK(cudaMallocHost(&g->h_off, cap * sizeof(int64_t)));
    ODIN_CUDA_CHECK(cudaMalloc(&g->d_off, cap * sizeof(int64_t)));
    g->off_cap = cap;
  }
  ODIN_CUDA_CHECK(cudaStreamSynchronize(st));
  const int64_t f_lo = h_frame_offsets[0], f_hi = h_frame_offsets[n_utt];
  for (int u = 0; u <= n_utt; ++u) g->h_off[u] = h_frame_offsets[u] - f_lo;
  ODIN_CUDA_CHECK(cudaMemcpyAsync(g->d_off, g->h_off, (n_utt + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
  const int64_t N = f_hi - f_lo;
  int rc = gmm_reserve_lse(g, N);
  if (rc) return rc;
  const float* X = d_X + f_lo * g->D;
  const uint8_t* sad = d_sad ? d_sad + f_lo : nullptr;
  if ((rc = gmm_lse_ffma(g, X, sad, N, g->d_lse, nullptr, st))) return rc;
  return gmm_utt_stats_ffma(g, X, sad, g->d_off, n_utt, g->d_lse, d_Z, d_Fhat, st);
}

int odin_gmm_score(odin_gmm_t* g, const float* d_X, int64_t n_frames, float* d_llk, float* d_post,
                   float* d_logprob, void* stream) {
  if (!g || !d_X || !d_llk || n_frames < 0) return set_error(ODIN_EINVAL, "bad argument");
  if (g->M <= 0) return set_error(ODIN_EINVAL, "odin_gmm_set_params has not been called");
  cudaStream_t st = as_stream(stream);
  int rc = gmm_lse_ffma(g, d_X, nullptr, n_frames, d_llk, nullptr, st);
  if (rc) return rc;
  if (d_post || d_logprob) return gmm_post_ffma(g, d_X, n_frames, d_llk, d_post, d_logprob, st);
  return ODIN_OK;
}

// --------------------------------- T-matrix ---------------------------------
int odin_tmat_create(int32_t tv_dim, int32_t nmix, int32_t feat_dim, odin_tmat_t** out) {
  if (!out || tv_dim <= 0 || nmix <= 0 || feat_dim <= 0) return set_error(ODIN_EINVAL, "bad argument");
  *out = nullptr;
  if (tv_dim > TMAT_MAX_TV || feat_dim > TMAT_MAX_D)
    return set_error(ODIN_EINVAL, "tv_dim must be <= %d and feat_dim <= %d", TMAT_MAX_TV, TMAT_MAX_D);
  int rc = require_device();
  if (rc) return rc;
  odin_tmat* t = new (std::nothrow) odin_tmat();
  if (!t) return set_error(ODIN_ENOMEM, "out of host memory");
  t->tv = tv_dim; t->M = nmix; t->D = feat_dim; t->t2 = tv_dim * (tv_dim + 1) / 2;
  t->MD = (int64_t)nmix * feat_dim;
  auto fail = [&](cudaError_t e) {
    odin_tmat_destroy(t);
    return set_error(ODIN_ENOMEM, "T-matrix buffers: %s", cudaGetErrorString(e));
  };
  cudaError_t e;
  const size_t big = sizeof(double) * (size_t)t->tv * t->MD;
  if ((e = cudaMalloc(&t->d_Tm, big)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&t->d_TinvS, big)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&t->d_Sigma, sizeof(double) * t->MD)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&t->d_TinvSTt, sizeof(double) * (size_t)t->M * t->t2)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&t->d_U, sizeof(double) * (size_t)t->tv * t->tv)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&t->d_perm, sizeof(int) * t->tv)) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&t->d_small, sizeof(double) * (size_t)std::max(t->t2, t->tv))) != cudaSuccess) return fail(e);
  if ((e = cudaMalloc(&t->d_flag, 2 * sizeof(int))) != cudaSuccess) return fail(e);   // status | Jacobi rotation counter
  if ((e = cudaMemset(t->d_flag, 0, 2 * sizeof(int))) != cudaSuccess) return fail(e);
  *out = t;
  return ODIN_OK;
}

void odin_tmat_destroy(odin_tmat_t* t) {
  if (!t) return;
  cudaFree(t->d_Tm); cudaFree(t->d_TinvS); cudaFree(t->d_Sigma); cudaFree(t->d_TinvSTt); cudaFree(t->d_U);
  cudaFree(t->d_perm); cudaFree(t->d_flag); cudaFree(t->d_L1); cudaFree(t->d_B1); cudaFree(t->d_Ex); cudaFree(t->d_llk);
  cudaFree(t->d_ws); cudaFree(t->d_gws); cudaFree(t->d_small);
  if (t->sweep_graph) cudaGraphExecDestroy((cudaGraphExec_t)t->sweep_graph);
  delete t;
}

int64_t odin_tmat_acc_size(const odin_tmat_t* t) {
  if (!t) return ODIN_EINVAL;
  return (int64_t)t->M * t->t2 + (int64_t)t->tv * t->MD + 2;
}

int odin_tmat_set_model(odin_tmat_t* t, const double* d_Tm, const double* d_Sigma, void* stream) {
  if (!t || !d_Tm) return set_error(ODIN_EINVAL, "bad argument");
  cudaStream_t st = as_stream(stream);
  ODIN_CUDA_CHECK(cudaMemcpyAsync(t->d_Tm, d_Tm, sizeof(double) * (size_t)t->tv * t->MD, cudaMemcpyDeviceToDevice, st));
  if (d_Sigma) ODIN_CUDA_CHECK(cudaMemcpyAsync(t->d_Sigma, d_Sigma, sizeof(double) * t->MD, cudaMemcpyDeviceToDevice, st));
  return tmat_refresh(t, st);
}

int odin_tmat_get_model(odin_tmat_t* t, double* d_Tm, double* d_T_invS, double* d_T_invS_Tt, void* stream) {
  if (!t) return set_error(ODIN_EINVAL, "bad argument");
  cudaStream_t st = as_stream(stream);
  const size_t big = sizeof(double) * (size_t)t->tv * t->MD;
  if (d_Tm) ODIN_CUDA_CHECK(cudaMemcpyAsync(d_Tm, t->d_Tm, big, cudaMemcpyDeviceToDevice, st));
  if (d_T_invS) ODIN_CUDA_CHECK(cudaMemcpyAsync(d_T_invS, t->d_TinvS, big, cudaMemcpyDeviceToDevice, st));
  if (d_T_invS_Tt)
    ODIN_CUDA_CHECK(cudaMemcpyAsync(d_T_invS_Tt, t->d_TinvSTt, sizeof(double) * (size_t)t->M * t->t2, cudaMemcpyDeviceToDevice, st));
  return ODIN_OK;
}

static int tmat_check_flag(odin_tmat* t, cudaStream_t st) {
  int flag = 0;
  ODIN_CUDA_CHECK(cudaMemcpyAsync(&flag, t->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  ODIN_CUDA_CHECK(cudaStreamSynchronize(st));
  if (flag != 0) {
    ODIN_CUDA_CHECK(cudaMemsetAsync(t->d_flag, 0, sizeof(int), st));
    return set_error(ODIN_ENUMERIC, "T-matrix: %s not positive definite",
                     flag == 1 ? "a posterior precision I + sum Z T' S^-1 T is" : flag == 2 ? "an M-step system sym(LU_m) is"
                                                                                            : "the minimum-divergence matrix is");
  }
  return ODIN_OK;
}

int odin_tmat_estep(odin_tmat_t* t, const double* d_Z, const double* d_F, int64_t n_files, double* d_acc, void* stream) {
  if (!t || !d_Z || !d_F || !d_acc || n_files < 0) return set_error(ODIN_EINVAL, "bad argument");
  if (n_files == 0) return ODIN_OK;
  return tmat_estep(t, d_Z, d_F, n_files, d_acc, as_stream(stream));
}

int odin_tmat_mstep(odin_tmat_t* t, const double* d_acc, int32_t min_div_est, int32_t orthogonalize, void* stream) {
  if (!t || !d_acc) return set_error(ODIN_EINVAL, "bad argument");
  cudaStream_t st = as_stream(stream);
  int rc = tmat_mstep(t, d_acc, min_div_est, orthogonalize, 30, st);
  if (rc) return rc;
  return tmat_check_flag(t, st);
}

int odin_tmat_ivector(odin_tmat_t* t, const double* d_Z, const double* d_F, int64_t n_files, double* d_out, void* stream) {
  if (!t || !d_Z || !d_F || !d_out || n_files < 0) return set_error(ODIN_EINVAL, "bad argument");
  if (n_files == 0) return ODIN_OK;
  cudaStream_t st = as_stream(stream);
  int rc = tmat_ivector(t, d_Z, d_F, n_files, d_out, st);
  if (rc) return rc;
  return tmat_check_flag(t, st);
}

}  // extern "C"
