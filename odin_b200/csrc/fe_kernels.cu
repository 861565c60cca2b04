// Speech front-end kernels for sm_100a.
//
// Reference arithmetic (trungnt13/odin-ai, file:line):
//   DC removal               odin/preprocessing/speech.py:453,472-473
//   pre-emphasis             odin/preprocessing/signal.py:955-967
//   framing/window/energy    signal.py:1442-1562, 1421-1440
//   |rfft|^2 * scale^2       signal.py:1555-1558, 1623-1648
//   mel filterbank + dB      signal.py:735-810, 1650-1691, 636-680
//   DCT -> MFCC, c0          signal.py:682-733, 1693-1716; speech.py:821-831
//   deltas                   signal.py:1002-1066; base.py:470-481
//   SADgmm                   signal.py:293-331; speech.py:1459-1477
//   SADthreshold             speech.py:1299-1324, 1415-1436
//   ApplyingSAD              speech.py:1732-1756
//
// Kernel plan (one ragged batch of utterances per call):
//   fe_dc_kernel     per-utterance sample sums (exact int64 for int16 PCM)
//   frame kernels    PCM -> unclipped log-mel rows, frame energies and an atomic per-utterance max; the two real frames
//                    of a pair are the re / im of one complex FFT, spectra never touch HBM:
//                      fe_frame5_kernel (fe_frame5.cu)  n_fft <= 1024: packed f32x2 four-step FFT in registers, every warp
//                                       on its own (cp.async staging), mel projection in registers -- the default;
//                      fe_frame4_kernel the scalar four-step kernel it replaced (spectrum / complex STFT outputs, unusual
//                                       filterbanks);  fe_frame_kernel  Stockham FFT in shared memory (n_fft = 2048)
//   fe_post9_kernel  utterance pass: utterance-global top_db clip, DCT, c0, delta / delta-delta with the reference's
//                    edge / latency quirks, on 16-byte conflict-free shared-memory accesses and packed pairs
//                    (9-tap deltas of order 2, widths divisible by 4); fe_post_kernel for every other shape
//   fe_vad_gmm_kernel / fe_vad_gmm_warp_kernel / fe_vad_thr_kernel
//                    SADgmm: a thread-block cluster (long utterances) or a warp (batches of short ones) per utterance:
//                    standardise (numpy-exact float32 mean / std), 1-D EM in fp64 with the component count as a template
//                    parameter, threshold, smoothing;  SADthreshold: one warp per utterance
//   fe_compact_*     ApplyingSAD row compaction.
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include <cooperative_groups.h>

#include "fe.cuh"
#include "fe_frame.cuh"
#include "fe_logic.cuh"

namespace cg = cooperative_groups;

namespace odin {

constexpr int PT = ODIN_FE_POST_TILE;

// ---------------------------------------------------------------------------
// small complex helpers
// ---------------------------------------------------------------------------
template <typename T> struct C2 { T x, y; };
template <typename T> __device__ __forceinline__ C2<T> cadd(C2<T> a, C2<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ C2<T> csub(C2<T> a, C2<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ C2<T> cmul(C2<T> a, C2<T> b) {
  return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T> __device__ __forceinline__ C2<T> cmul_negi(C2<T> a) { return {a.y, -a.x}; }  // a * (-i)

// W_32^q = exp(-2 pi i q / 32), q = 0..15
__device__ __forceinline__ constexpr double w32c(int q) {
  return q == 0 ? 1.0 : q == 1 ? 0.98078528040323044913 : q == 2 ? 0.92387953251128675613
       : q == 3 ? 0.83146961230254523708 : q == 4 ? 0.70710678118654752440 : q == 5 ? 0.55557023301960222474
       : q == 6 ? 0.38268343236508977173 : q == 7 ? 0.19509032201612826785 : q == 8 ? 0.0
       : q == 9 ? -0.19509032201612826785 : q == 10 ? -0.38268343236508977173 : q == 11 ? -0.55557023301960222474
       : q == 12 ? -0.70710678118654752440 : q == 13 ? -0.83146961230254523708 : q == 14 ? -0.92387953251128675613
       : -0.98078528040323044913;
}
__device__ __forceinline__ constexpr double w32s(int q) {  // -sin(2 pi q / 32)
  return q == 0 ? 0.0 : q == 1 ? -0.19509032201612826785 : q == 2 ? -0.38268343236508977173
       : q == 3 ? -0.55557023301960222474 : q == 4 ? -0.70710678118654752440 : q == 5 ? -0.83146961230254523708
       : q == 6 ? -0.92387953251128675613 : q == 7 ? -0.98078528040323044913 : q == 8 ? -1.0
       : q == 9 ? -0.98078528040323044913 : q == 10 ? -0.92387953251128675613 : q == 11 ? -0.83146961230254523708
       : q == 12 ? -0.70710678118654752440 : q == 13 ? -0.55557023301960222474 : q == 14 ? -0.38268343236508977173
       : -0.19509032201612826785;
}

// in-register forward DFT of size R (power of two <= 32), natural order in/out.
// NZ < R declares v[NZ..R) to be zero (a zero-padded frame: L <= n_fft / 2): the even / odd halves inherit
// the zero tail, and a sub-transform with a single live input is a broadcast, so the leaf butterflies vanish.
template <int R, typename T, int NZ = R> struct Dft {
  static __device__ __forceinline__ void run(C2<T> (&v)[R]) {
    if constexpr (NZ <= 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) v[q] = v[0];
    } else {
      C2<T> e[R / 2], o[R / 2];
#pragma unroll
      for (int q = 0; q < R / 2; ++q) { e[q] = v[2 * q]; o[q] = v[2 * q + 1]; }
      Dft<R / 2, T, (NZ + 1) / 2>::run(e);
      Dft<R / 2, T, NZ / 2>::run(o);
#pragma unroll
      for (int q = 0; q < R / 2; ++q) {
        C2<T> t;
        if (q == 0) t = o[q];
        else if (4 * q == R) t = cmul_negi(o[q]);
        else t = cmul(o[q], C2<T>{(T)w32c(q * (32 / R)), (T)w32s(q * (32 / R))});
        v[q] = cadd(e[q], t);
        v[q + R / 2] = csub(e[q], t);
      }
    }
  }
};
template <typename T, int NZ> struct Dft<1, T, NZ> {
  static __device__ __forceinline__ void run(C2<T> (&)[1]) {}
};

template <int N> struct FftPlan;
template <> struct FftPlan<256> { static constexpr int R0 = 8, R1 = 8, R2 = 4; };
template <> struct FftPlan<512> { static constexpr int R0 = 8, R1 = 8, R2 = 8; };
template <> struct FftPlan<1024> { static constexpr int R0 = 16, R1 = 8, R2 = 8; };
template <> struct FftPlan<2048> { static constexpr int R0 = 16, R1 = 16, R2 = 8; };

// one float2/double2 of padding every 16 elements keeps the strided Stockham
// stores spread over the banks
__device__ __forceinline__ int padi(int i) { return i + (i >> 4); }
template <int N> constexpr int padded_len() { return N + (N >> 4) + 1; }

// Stockham pass of radix R over a warp-private buffer, in place: every lane pulls
// all of its butterfly inputs into registers, the warp syncs, results are written
// back to the auto-sorted positions.  NS = product of the radices of earlier passes.
template <int N, int R, int NS, typename T>
__device__ __forceinline__ void fft_pass(C2<T>* buf, const C2<T>* __restrict__ tw, int lane) {
  constexpr int NB = N / (32 * R);
  C2<T> v[NB][R];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int j = lane + 32 * b;
#pragma unroll
    for (int r = 0; r < R; ++r) v[b][r] = buf[padi(j + r * (N / R))];
  }
  __syncwarp();
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int j = lane + 32 * b;
    const int k = j % NS;
    if (NS > 1) {
      constexpr int stride = N / (NS * R);
#pragma unroll
      for (int r = 1; r < R; ++r) v[b][r] = cmul(v[b][r], tw[r * k * stride]);
    }
    Dft<R, T>::run(v[b]);
    const int base = (j / NS) * (NS * R) + k;
#pragma unroll
    for (int r = 0; r < R; ++r) buf[padi(base + r * NS)] = v[b][r];
  }
  __syncwarp();
}

// ---------------------------------------------------------------------------
// fe_dc_kernel: per-utterance sums for the DC removal
// ---------------------------------------------------------------------------
constexpr int DC_CHUNK = 16384;

template <typename PCM>
__global__ void __launch_bounds__(256)
fe_dc_kernel(const PCM* __restrict__ pcm, const int64_t* __restrict__ sample_off, int n_utt, int max_chunks,
             double* __restrict__ dcsum) {
  const int u = blockIdx.x / max_chunks, c = blockIdx.x % max_chunks;
  if (u >= n_utt) return;
  const int64_t s0 = sample_off[u], n = sample_off[u + 1] - s0;
  const int64_t lo = (int64_t)c * DC_CHUNK;
  if (lo >= n) return;
  const int cnt = (int)min((int64_t)DC_CHUNK, n - lo);
  const PCM* p = pcm + s0 + lo;
  __shared__ double red[8];
  if (sizeof(PCM) == 2) {
    long long acc = 0;
    for (int i = threadIdx.x; i < cnt; i += 256) acc += (long long)p[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ long long redi[8];
    if ((threadIdx.x & 31) == 0) redi[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      long long t = 0;
      for (int i = 0; i < 8; ++i) t += redi[i];
      atomicAdd(reinterpret_cast<unsigned long long*>(dcsum) + u, (unsigned long long)t);
    }
  } else {
    double acc = 0;
    for (int i = threadIdx.x; i < cnt; i += 256) acc += (double)p[i];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int i = 0; i < 8; ++i) t += red[i];
      atomicAdd(dcsum + u, t);
    }
  }
}

// ---------------------------------------------------------------------------
// fe_frame_kernel
// ---------------------------------------------------------------------------
template <int N, typename T, typename PCM>
__global__ void __launch_bounds__(FE_THREADS) fe_frame_kernel(FrameArgs a) {
  using P = FftPlan<N>;
  constexpr int PL = padded_len<N>();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: win64 [L] | tw [N] | warp bufs [FE_WARPS][PL] | win32 [L] | sbuf [(FT-1)*hop + L + 16]
  double* win64 = reinterpret_cast<double*>(smem_raw);
  C2<T>* tw = reinterpret_cast<C2<T>*>(win64 + a.L + (a.L & 1));
  C2<T>* bufs = tw + N;
  float* win32 = reinterpret_cast<float*>(bufs + FE_WARPS * PL);
  float* sbuf = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(win32 + a.L) + 15) & ~uintptr_t(15));   // 16-byte stores
  __shared__ int cta_max, cta_max_spec;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.L, hop = a.hop;
  for (int i = tid; i < L; i += FE_THREADS) { win64[i] = a.win64[i]; win32[i] = a.win32[i]; }
  for (int i = tid; i < N; i += FE_THREADS) tw[i] = reinterpret_cast<const C2<T>*>(a.tw)[i];
  C2<T>* buf = bufs + warp * PL;
  const PCM* __restrict__ pcm = reinterpret_cast<const PCM*>(a.pcm);
  const float coef = a.preemph;

  for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int u = find_segment(a.tile_off, a.n_utt, tile);
    const int64_t s0 = a.sample_off[u];
    const int64_t n_u = a.sample_off[u + 1] - s0;
    const int64_t fbase = a.frame_off[u];
    const int T_u = (int)(a.frame_off[u + 1] - fbase);
    const int t0 = (int)(tile - a.tile_off[u]) * FT;
    const int nf = min(FT, T_u - t0);
    float mean = 0.f;
    if (a.remove_dc) {
      double s = (sizeof(PCM) == 2) ? (double)reinterpret_cast<const long long*>(a.dcsum)[u] : a.dcsum[u];
      mean = (float)(s / (double)n_u);
    }
    __syncthreads();  // previous tile done with sbuf / cta_max (and the table fill on the first trip)
    if (tid == 0) { cta_max = float_to_ordered(-FLT_MAX); cta_max_spec = float_to_ordered(-FLT_MAX); }
    const float* stile = stage_pcm<PCM>(sbuf, pcm + s0, n_u, (int64_t)t0 * hop, (nf - 1) * hop + L, mean, coef, a.pad,
                                        tid, FE_THREADS);
    __syncthreads();

    float wmax = -FLT_MAX, smax = -FLT_MAX;
    for (int pair = warp; 2 * pair < nf; pair += FE_WARPS) {
      const int fA = 2 * pair, fB = fA + 1;
      const bool hasB = fB < nf;
      const float* sA = stile + fA * hop;
      const float* sB = stile + (hasB ? fB : fA) * hop;
      // ---- pass 0 fused with windowing and the fp64 frame energy ----
      {
        constexpr int R = P::R0;
        constexpr int NB = N / (32 * R);
        double eA = 0.0, eB = 0.0;
        C2<T> v[NB][R];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int j = lane + 32 * b;
#pragma unroll
          for (int r = 0; r < R; ++r) {
            const int i = j + r * (N / R);
            C2<T> z = {(T)0, (T)0};
            if (i < L) {
              const float xa = sA[i], xb = hasB ? sB[i] : 0.f;
              const double w = win64[i];
              const double wa = w * (double)xa, wb = w * (double)xb;
              eA = fma(wa, wa, eA);
              eB = fma(wb, wb, eB);
              if (sizeof(T) == 8) { z.x = (T)wa; z.y = (T)wb; }
              else { const float wf = win32[i]; z.x = (T)(wf * xa); z.y = (T)(wf * xb); }
            }
            v[b][r] = z;
          }
        }
        __syncwarp();  // previous pair's mel stage has finished reading buf
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int j = lane + 32 * b;
          Dft<R, T>::run(v[b]);
#pragma unroll
          for (int r = 0; r < R; ++r) buf[padi(j * R + r)] = v[b][r];
        }
        __syncwarp();
        if (a.energy != nullptr) {
          eA = warp_sum(eA);
          eB = warp_sum(eB);
          if (lane == 0) {
            if (eA == 0.0) eA = (double)FLT_EPSILON;  // signal.py:1436
            a.energy[fbase + t0 + fA] = (float)log(eA);
            if (hasB) {
              if (eB == 0.0) eB = (double)FLT_EPSILON;
              a.energy[fbase + t0 + fB] = (float)log(eB);
            }
          }
        }
      }
      fft_pass<N, P::R1, P::R0, T>(buf, tw, lane);
      fft_pass<N, P::R2, P::R0 * P::R1, T>(buf, tw, lane);
      // ---- split the packed spectra: XA = (Z[k] + conj Z[N-k])/2, XB = (Z[k] - conj Z[N-k])/(2i) ----
      constexpr int NK = (N / 2) / 32;  // bins per lane below N/2
      T pa[NK + 1], pb[NK + 1];
      const T q = (T)0.25 * (T)a.scale2;
#pragma unroll
      for (int i = 0; i < NK; ++i) {
        const int k = lane + 32 * i;
        const C2<T> z1 = buf[padi(k)], z2 = buf[padi((N - k) & (N - 1))];
        const T ar = z1.x + z2.x, ai = z1.y - z2.y;
        const T br = z1.y + z2.y, bi = z2.x - z1.x;
        pa[i] = (ar * ar + ai * ai) * q;
        pb[i] = (br * br + bi * bi) * q;
        if (a.cspec != nullptr) {   // signal.stft: the complex spectra themselves
          const float hs = 0.5f * a.scale1;
          float2* cA = a.cspec + (fbase + t0 + fA) * (N / 2 + 1);
          cA[k] = make_float2((float)ar * hs, (float)ai * hs);
          if (hasB) cA[(N / 2 + 1) + k] = make_float2((float)br * hs, (float)bi * hs);
        }
      }
      {
        const C2<T> zn = buf[padi(N / 2)];
        pa[NK] = (zn.x * zn.x) * (T)a.scale2;
        pb[NK] = (zn.y * zn.y) * (T)a.scale2;
        if (a.cspec != nullptr && lane == 0) {
          float2* cA = a.cspec + (fbase + t0 + fA) * (N / 2 + 1);
          cA[N / 2] = make_float2((float)zn.x * a.scale1, 0.f);
          if (hasB) cA[(N / 2 + 1) + N / 2] = make_float2((float)zn.y * a.scale1, 0.f);
        }
      }
      if (a.spec != nullptr) {   // SpectraExtractor: the power spectrum itself is an output
        constexpr int NBIN = N / 2 + 1;
        float* rowA = a.spec + (fbase + t0 + fA) * NBIN;
        float* rowB = rowA + NBIN;
#pragma unroll
        for (int i = 0; i <= NK; ++i) {
          if (i == NK && lane != 0) break;
          const int k = (i == NK) ? N / 2 : lane + 32 * i;
          const float va = a.spec_log ? (float)db10<T>(pa[i]) : (float)pa[i];
          rowA[k] = va;
          smax = fmaxf(smax, va);
          if (hasB) {
            const float vb = a.spec_log ? (float)db10<T>(pb[i]) : (float)pb[i];
            rowB[k] = vb;
            smax = fmaxf(smax, vb);
          }
        }
      }
      __syncwarp();
      T* PA = reinterpret_cast<T*>(buf);
      T* PB = PA + (N / 2 + 1);
#pragma unroll
      for (int i = 0; i < NK; ++i) { PA[lane + 32 * i] = pa[i]; PB[lane + 32 * i] = pb[i]; }
      if (lane == 0) { PA[N / 2] = pa[NK]; PB[N / 2] = pb[NK]; }
      __syncwarp();
      // ---- sparse mel triangles, one filter per lane ----
      for (int m = lane; m < a.n_mels; m += 32) {
        const int st = a.mel_start[m], cn = a.mel_cnt[m];
        const float* __restrict__ w = a.mel_w + a.mel_off[m];
        T accA = 0, accB = 0;
        for (int i = 0; i < cn; ++i) {
          const T wi = (T)__ldg(w + i);
          accA += wi * PA[st + i];
          accB += wi * PB[st + i];
        }
        const float dA = (float)db10<T>(accA);
        a.mspec[(fbase + t0 + fA) * a.n_mels + m] = dA;
        wmax = fmaxf(wmax, dA);
        if (hasB) {
          const float dB = (float)db10<T>(accB);
          a.mspec[(fbase + t0 + fB) * a.n_mels + m] = dB;
          wmax = fmaxf(wmax, dB);
        }
      }
    }
    wmax = warp_max(wmax);
    if (lane == 0) atomicMax(&cta_max, float_to_ordered(wmax));
    if (a.spec != nullptr && a.spec_log) {
      smax = warp_max(smax);
      if (lane == 0) atomicMax(&cta_max_spec, float_to_ordered(smax));
    }
    __syncthreads();
    if (tid == 0) {
      atomicMax(a.umax + u, cta_max);
      if (a.spec != nullptr && a.spec_log) atomicMax(a.umax_spec + u, cta_max_spec);
    }
  }
}

// ---------------------------------------------------------------------------
// fe_frame4_kernel: same contract as fe_frame_kernel, four-step FFT.
//
// N = G * 32.  G lanes share one complex FFT (two real frames as re / im) and every
// lane keeps 32 elements in registers:
//   step A   lane l loads z[l + G r], r = 0..31 (window multiply and the fp64 frame
//            energy fused), runs a 32-point DFT over r in registers
//   twiddle  Y[l][k1] *= w_N^(l k1)                       (table tw4[k1][l])
//   exchange through a warp-private padded smem tile [k1][G + 1] -- the ONLY round
//            trip of the spectrum through shared memory (the Stockham kernel made three)
//   step B   lane j runs the G-point DFT over l for its 32 / G values of k1 and writes
//            X[k1 + 32 k2] in natural order for the split / mel stage
// A warp therefore handles 32 / G frame pairs at once (1 at n_fft 1024, 2 at 512, 4 at
// 256) in <= 128 registers, so two CTAs (16 warps) fit per SM instead of one.
// ---------------------------------------------------------------------------
template <int N> constexpr int f4_region() { return 32 * (N / 32 + 1); }   // float2 per pair

template <int N, typename PCM, bool ZH>   // ZH: L <= N / 2, rows r >= 16 of step A are zero padding
__global__ void __launch_bounds__(FE_THREADS, 2) fe_frame4_kernel(FrameArgs a) {
  using T = float;
  constexpr int G = N / 32, NP = 32 / G, RS = G + 1, REG = f4_region<N>();
  static_assert(REG >= N, "pair region must hold the natural-order spectrum");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: win64 [L] | tw4 [N] | warp bufs [FE_WARPS][NP * REG] | mel table [trips][32] float2 | win32 [L] |
  //         mel slots [n_mels + 1] | sbuf [(FT-1)*hop + L + 16]
  double* win64 = reinterpret_cast<double*>(smem_raw);
  C2<T>* tw4 = reinterpret_cast<C2<T>*>(win64 + a.L + (a.L & 1));
  C2<T>* bufs = tw4 + N;
  float2* mtab = reinterpret_cast<float2*>(bufs + FE_WARPS * NP * REG);
  float* win32 = reinterpret_cast<float*>(mtab + a.mel_trips * 32);
  int* mps = reinterpret_cast<int*>(win32 + a.L);
  float* sbuf = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(mps + a.n_mels + 1) + 15) & ~uintptr_t(15));   // 16-byte stores
  __shared__ int cta_max, cta_max_spec;
  __shared__ double s_en[FT];   // frame energies of the tile; their logs are taken by one warp at the end

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / G, l = lane % G;
  const int L = a.L, hop = a.hop;
  for (int i = tid; i < L; i += FE_THREADS) { win64[i] = a.win64[i]; win32[i] = a.win32[i]; }
  for (int i = tid; i < N; i += FE_THREADS) tw4[i] = reinterpret_cast<const C2<T>*>(a.tw)[i];
  for (int i = tid; i < a.mel_trips * 32; i += FE_THREADS) mtab[i] = a.mel_tab[i];
  for (int i = tid; i <= a.n_mels; i += FE_THREADS) mps[i] = a.mel_ps[i];
  C2<T>* wbuf = bufs + warp * (NP * REG);
  const PCM* __restrict__ pcm = reinterpret_cast<const PCM*>(a.pcm);
  const float coef = a.preemph;

  // contiguous range of tiles per CTA: one binary search, then the utterance index walks forward
  const int64_t per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t tile_lo = per * blockIdx.x, tile_hi = min(a.n_tiles, tile_lo + per);
  if (tile_lo >= tile_hi) return;
  int u = find_segment(a.tile_off, a.n_utt, tile_lo);
  int64_t u_end = a.tile_off[u + 1];
  for (int64_t tile = tile_lo; tile < tile_hi; ++tile) {
    while (tile >= u_end) { ++u; u_end = a.tile_off[u + 1]; }
    const int64_t s0 = a.sample_off[u];
    const int64_t n_u = a.sample_off[u + 1] - s0;
    const int64_t fbase = a.frame_off[u];
    const int T_u = (int)(a.frame_off[u + 1] - fbase);
    const int t0 = (int)(tile - a.tile_off[u]) * FT;
    const int nf = min(FT, T_u - t0);
    float mean = 0.f;
    if (a.remove_dc) {
      double s = (sizeof(PCM) == 2) ? (double)reinterpret_cast<const long long*>(a.dcsum)[u] : a.dcsum[u];
      mean = (float)(s / (double)n_u);
    }
    __syncthreads();  // previous tile done with sbuf / cta_max (and the table fill on the first trip)
    if (tid == 0) { cta_max = float_to_ordered(-FLT_MAX); cta_max_spec = float_to_ordered(-FLT_MAX); }
    const float* stile = stage_pcm<PCM>(sbuf, pcm + s0, n_u, (int64_t)t0 * hop, (nf - 1) * hop + L, mean, coef, a.pad,
                                        tid, FE_THREADS);
    __syncthreads();

    float wmax = -FLT_MAX, smax = -FLT_MAX;
    const int npairs = (nf + 1) >> 1;
    for (int base = warp * NP; base < npairs; base += FE_WARPS * NP) {
      C2<T>* reg = wbuf + g * REG;
      // ---------------- step A: load + window + energy, 32-point DFT over r, twiddle
      {
        const int pr = base + g;
        const int fA = 2 * pr, fB = fA + 1;
        const bool hasA = fA < nf, hasB = fB < nf;
        const float* sA = stile + (hasA ? fA : 0) * hop;
        const float* sB = stile + (hasB ? fB : 0) * hop;
        double eA = 0.0, eB = 0.0;
        C2<T> v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const int i = l + G * r;
          C2<T> z = {0.f, 0.f};
          if ((!ZH || r < 16) && i < L && hasA) {
            const float xa = sA[i], xb = hasB ? sB[i] : 0.f;
            const double w = win64[i];
            const double wa = w * (double)xa, wb = w * (double)xb;
            eA = fma(wa, wa, eA);
            eB = fma(wb, wb, eB);
            const float wf = win32[i];
            z.x = wf * xa; z.y = wf * xb;
          }
          v[r] = z;
        }
        if (a.energy != nullptr) {
#pragma unroll
          for (int o = G / 2; o > 0; o >>= 1) {
            eA += __shfl_xor_sync(0xffffffffu, eA, o);
            eB += __shfl_xor_sync(0xffffffffu, eB, o);
          }
          if (l == 0 && hasA) {
            s_en[fA] = eA;
            if (hasB) s_en[fB] = eB;
          }
        }
        Dft<32, T, ZH ? 16 : 32>::run(v);
#pragma unroll
        for (int k1 = 1; k1 < 32; ++k1) v[k1] = cmul(v[k1], tw4[k1 * G + l]);
        __syncwarp();  // the previous pass' mel stage has finished reading the regions
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) reg[k1 * RS + l] = v[k1];
      }
      __syncwarp();
      // ---------------- step B: G-point DFT over l for k1 = l + G t, natural-order store
      {
        C2<T> x[NP][G];
#pragma unroll
        for (int t = 0; t < NP; ++t) {
          const int k1 = l + G * t;
#pragma unroll
          for (int i = 0; i < G; ++i) x[t][i] = reg[k1 * RS + i];
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < NP; ++t) {
          Dft<G, T>::run(x[t]);
          const int k1 = l + G * t;
#pragma unroll
          for (int k2 = 0; k2 < G; ++k2) reg[k1 + 32 * k2] = x[t][k2];
        }
      }
      __syncwarp();
      // ---------------- split + |.|^2 + mel + dB, the whole warp on one pair at a time
#pragma unroll 1
      for (int pp = 0; pp < NP; ++pp) {
        const int fA = 2 * (base + pp), fB = fA + 1;
        if (fA >= nf) break;
        const bool hasB = fB < nf;
        C2<T>* buf = wbuf + pp * REG;
        constexpr int NK = (N / 2) / 32;  // bins per lane below N/2
        T pa[NK + 1], pb[NK + 1];
        const T q = (T)0.25 * (T)a.scale2;
#pragma unroll
        for (int i = 0; i < NK; ++i) {
          const int k = lane + 32 * i;
          const C2<T> z1 = buf[k], z2 = buf[(N - k) & (N - 1)];
          const T ar = z1.x + z2.x, ai = z1.y - z2.y;
          const T br = z1.y + z2.y, bi = z2.x - z1.x;
          pa[i] = (ar * ar + ai * ai) * q;
          pb[i] = (br * br + bi * bi) * q;
          if (a.cspec != nullptr) {   // signal.stft: the complex spectra themselves
            const float hs = 0.5f * a.scale1;
            float2* cA = a.cspec + (fbase + t0 + fA) * (N / 2 + 1);
            cA[k] = make_float2(ar * hs, ai * hs);
            if (hasB) cA[(N / 2 + 1) + k] = make_float2(br * hs, bi * hs);
          }
        }
        {
          const C2<T> zn = buf[N / 2];
          pa[NK] = (zn.x * zn.x) * (T)a.scale2;
          pb[NK] = (zn.y * zn.y) * (T)a.scale2;
          if (a.cspec != nullptr && lane == 0) {
            float2* cA = a.cspec + (fbase + t0 + fA) * (N / 2 + 1);
            cA[N / 2] = make_float2(zn.x * a.scale1, 0.f);
            if (hasB) cA[(N / 2 + 1) + N / 2] = make_float2(zn.y * a.scale1, 0.f);
          }
        }
        if (a.spec != nullptr) {   // SpectraExtractor: the power spectrum itself is an output
          constexpr int NBIN = N / 2 + 1;
          float* rowA = a.spec + (fbase + t0 + fA) * NBIN;
          float* rowB = rowA + NBIN;
#pragma unroll
          for (int i = 0; i <= NK; ++i) {
            if (i == NK && lane != 0) break;
            const int k = (i == NK) ? N / 2 : lane + 32 * i;
            const float va = a.spec_log ? (float)db10<T>(pa[i]) : (float)pa[i];
            rowA[k] = va;
            smax = fmaxf(smax, va);
            if (hasB) {
              const float vb = a.spec_log ? (float)db10<T>(pb[i]) : (float)pb[i];
              rowB[k] = vb;
              smax = fmaxf(smax, vb);
            }
          }
        }
        __syncwarp();
        T* PA = reinterpret_cast<T*>(buf);
        T* PB = PA + (N / 2 + 1);
#pragma unroll
        for (int i = 0; i < NK; ++i) { PA[lane + 32 * i] = pa[i]; PB[lane + 32 * i] = pb[i]; }
        if (lane == 0) { PA[N / 2] = pa[NK]; PB[N / 2] = pb[NK]; }
        __syncwarp();
        // lane-balanced sparse projection: every lane walks its own list of (weight, bin) taps and drops a
        // partial sum into the chunk's slot at the end of each chunk; slots live behind PA | PB in the region
        T* QA = PB + (N / 2 + 1);
        T* QB = QA + a.mel_chunks;
        {
          T accA = 0, accB = 0;
          for (int i = 0; i < a.mel_trips; ++i) {
            const float2 e = mtab[i * 32 + lane];
            const uint32_t meta = __float_as_uint(e.y);
            const int bin = meta & 0x7fff;
            accA = fmaf(e.x, PA[bin], accA);
            accB = fmaf(e.x, PB[bin], accB);
            if (meta & 0x8000u) {
              QA[meta >> 16] = accA; QB[meta >> 16] = accB;
              accA = 0; accB = 0;
            }
          }
        }
        __syncwarp();
        for (int m = lane; m < a.n_mels; m += 32) {
          const int s0 = mps[m], s1 = mps[m + 1];
          T accA = 0, accB = 0;
          for (int i = s0; i < s1; ++i) { accA += QA[i]; accB += QB[i]; }
          const float dA = db10<T>(accA);
          a.mspec[(fbase + t0 + fA) * a.n_mels + m] = dA;
          wmax = fmaxf(wmax, dA);
          if (hasB) {
            const float dB = db10<T>(accB);
            a.mspec[(fbase + t0 + fB) * a.n_mels + m] = dB;
            wmax = fmaxf(wmax, dB);
          }
        }
      }
    }
    wmax = warp_max(wmax);
    if (lane == 0) atomicMax(&cta_max, float_to_ordered(wmax));
    if (a.spec != nullptr && a.spec_log) {
      smax = warp_max(smax);
      if (lane == 0) atomicMax(&cta_max_spec, float_to_ordered(smax));
    }
    __syncthreads();
    if (tid == 0) {
      atomicMax(a.umax + u, cta_max);
      if (a.spec != nullptr && a.spec_log) atomicMax(a.umax_spec + u, cta_max_spec);
    }
    if (a.energy != nullptr && tid < nf) {
      double e = s_en[tid];
      if (e == 0.0) e = (double)FLT_EPSILON;  // signal.py:1436
      a.energy[fbase + t0 + tid] = (float)log(e);
    }
  }
}

// ---------------------------------------------------------------------------
// fe_post_kernel: clip, DCT, deltas
// ---------------------------------------------------------------------------
struct PostArgs {
  const int64_t* frame_off;
  const int64_t* tile2_off;
  int n_utt;
  const int* umax;
  float top_db;
  int64_t n_tiles;
  int n_mels, n_c1, n_ceps, W, order;
  const float* dct;   // [n_c1, n_mels]
  const float* taps;  // [W]
  float* mspec;       // in/out
  int write_mspec;
  float* feat;        // [T, n_ceps*(order+1)] nullable
  float* c0;          // [T] nullable
};

// Division of small non-negative ints by a runtime constant: q = umulhi(n, M), M = floor((2^32-1)/d) + 1
// (exact while n*d < 2^32; d == 1 wraps M to 0, which is the pass-through flag).
__device__ __forceinline__ uint32_t fdiv_magic(int d) { return 0xFFFFFFFFu / (uint32_t)d + 1u; }
__device__ __forceinline__ int fdiv(int n, uint32_t M) { return M ? (int)__umulhi((uint32_t)n, M) : n; }

// SpectraExtractor's dB spectrum: clip at (utterance max - top_db) (signal.py:676-679), in place, tiles of PT frames
__global__ void __launch_bounds__(256) fe_spec_clip_kernel(float* __restrict__ spec, const int64_t* __restrict__ frame_off,
                                                           const int64_t* __restrict__ tile2_off, int n_utt,
                                                           int64_t n_tiles, const int* __restrict__ umax_spec,
                                                           float top_db, int nb) {
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int u = find_segment(tile2_off, n_utt, tile);
    const int64_t base = frame_off[u];
    const int T = (int)(frame_off[u + 1] - base);
    const int t0 = (int)(tile - tile2_off[u]) * PT;
    const int nf = min(PT, T - t0);
    const float floor_db = ordered_to_float(umax_spec[u]) - top_db;
    float* p = spec + (base + t0) * nb;
    for (int i = threadIdx.x; i < nf * nb; i += 256)
      if (p[i] < floor_db) p[i] = floor_db;
  }
}

// Four consecutive outputs of a W-tap causal FIR down a strided column: out[j] = sum_k taps[k] x[(j + W-1 - k) * stride],
// k ascending from a zero accumulator (the order of the one-row loops), the W + 3 inputs read once.
template <int W>
__device__ __forceinline__ void fir4(const float* __restrict__ x, int stride, const float* __restrict__ taps,
                                     float (&out)[4]) {
  float v[W + 3], tp[W];
#pragma unroll
  for (int q = 0; q < W + 3; ++q) v[q] = x[q * stride];
#pragma unroll
  for (int k = 0; k < W; ++k) tp[k] = taps[k];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < W; ++k) acc = fmaf(tp[k], v[j + W - 1 - k], acc);
    out[j] = acc;
  }
}

// The same for TWO adjacent columns at once on packed fp32 pairs: out[j] = (column c, column c + 1) of row j.  Each
// lane of an FFMA2 is the IEEE fma of the scalar loop, so the results are bit-identical to fir4's.
template <int W>
__device__ __forceinline__ void fir4x2(const float* __restrict__ x, int stride, const float (&tp)[W], u64 (&out)[4]) {
  u64 v[W + 3];
#pragma unroll
  for (int q = 0; q < W + 3; ++q) v[q] = pk2(x[q * stride], x[q * stride + 1]);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    u64 acc = 0ull;
#pragma unroll
    for (int k = 0; k < W; ++k) acc = fma2(pk2(tp[k], tp[k]), v[j + W - 1 - k], acc);
    out[j] = acc;
  }
}

// Utterance pass.  Issue-bound (ncu: 72 % of issue slots at 728 warp-instructions per frame in its first
// version), so the work is arranged to cost few instructions: runtime divisors go through fdiv, the DCT is
// register-blocked 2 rows x 8 coefficients with the basis transposed and zero-padded to a multiple of 8
// (two broadcast 16-byte loads per mel), rows of a thread are half a tile apart so lanes walk distinct banks.
__global__ void __launch_bounds__(256, 3) fe_post_kernel(PostArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int h = a.W / 2;
  const int HL = (a.order >= 2) ? (a.W - 1 + h) : (a.order == 1 ? h : 0);
  const int HR = (a.order >= 1) ? h : 0;
  const int NR = PT + HL + HR;             // rows of mel / cepstra held
  const int MS = a.n_mels | 1;              // odd row stride: lanes walking rows hit distinct banks
  const int C8 = (a.n_c1 + 7) & ~7;         // coefficients padded to whole groups of 8
  float* sdct = sm;                         // [n_mels][C8]  transposed basis, 16-byte aligned rows
  float* staps = sdct + a.n_mels * C8;      // [W] (+ pad to 4)
  float* scep = staps + ((a.W + 3) & ~3);   // [NR][n_c1]
  float* smel = scep + ((NR * a.n_c1 + 3) & ~3);   // [NR][MS]
  float* sd1 = smel;                        // [ND][n_ceps] first-order deltas: the mel rows are dead by then
  const uint32_t mg_c8 = fdiv_magic(C8), mg_ceps = fdiv_magic(max(a.n_ceps, 1));

  // persistent CTA over a contiguous range of tiles: the tables are filled once and the utterance of a
  // tile is found by walking forward from the previous one (one binary search per CTA)
  for (int i = tid; i < a.n_mels * C8; i += 256) {
    const int m = fdiv(i, mg_c8), c = i - m * C8;
    sdct[i] = (c < a.n_c1) ? a.dct[c * a.n_mels + m] : 0.f;
  }
  for (int i = tid; i < a.W; i += 256) staps[i] = a.taps[i];
  float tp9[9];   // the taps of the usual 9-wide delta window, in registers
#pragma unroll
  for (int k = 0; k < 9; ++k) tp9[k] = a.taps[min(k, a.W - 1)];
  const int64_t per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t tile_lo = per * blockIdx.x, tile_hi = min(a.n_tiles, tile_lo + per);
  if (tile_lo >= tile_hi) return;
  int u = find_segment(a.tile2_off, a.n_utt, tile_lo);
  int64_t u_end = a.tile2_off[u + 1];
  for (int64_t tile = tile_lo; tile < tile_hi; ++tile) {
  while (tile >= u_end) { ++u; u_end = a.tile2_off[u + 1]; }   // utterances shorter than a frame own no tile
  const int64_t base = a.frame_off[u];
  const int T = (int)(a.frame_off[u + 1] - base);
  const int t0 = (int)(tile - a.tile2_off[u]) * PT;
  const int nf = min(PT, T - t0);
  const int lo = max(0, t0 - HL), hi = min(T, t0 + nf + HR);
  const int nrows = hi - lo;
  const float floor_db = (a.top_db >= 0.f) ? ordered_to_float(a.umax[u]) - a.top_db : -FLT_MAX;
  __syncthreads();   // the previous tile is done with smel / scep / sd1 (and the table fill on the first trip)
  {
    float* g0 = a.mspec + (base + lo) * a.n_mels;
    if ((a.n_mels & 3) == 0 && (reinterpret_cast<uintptr_t>(a.mspec) & 15) == 0) {
      const int nq = a.n_mels >> 2;
      const uint32_t mg_nq = fdiv_magic(nq);
      float4* g4 = reinterpret_cast<float4*>(g0);
      // the rows are read and (clipped) written back through the same pointer, so the loads of a batch
      // are issued before its stores by hand: 4 x 16 B in flight per thread
      const int n4 = nrows * nq;
      for (int i0 = tid; i0 < n4; i0 += 4 * 256) {
        float4 vv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k * 256;
          vv[k] = (i < n4) ? g4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k * 256;
          if (i < n4) {
            const int r = fdiv(i, mg_nq), q = i - r * nq;
            float4 v = vv[k];
            const bool clip = fminf(fminf(v.x, v.y), fminf(v.z, v.w)) < floor_db;
            v.x = fmaxf(v.x, floor_db); v.y = fmaxf(v.y, floor_db);
            v.z = fmaxf(v.z, floor_db); v.w = fmaxf(v.w, floor_db);
            float* d = smel + r * MS + 4 * q;
            d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            const int t = lo + r;
            if (clip && a.write_mspec && t >= t0 && t < t0 + nf) g4[i] = v;   // unclipped rows are already final
          }
        }
      }
    } else {
      const uint32_t mg_m = fdiv_magic(a.n_mels);
      for (int i = tid; i < nrows * a.n_mels; i += 256) {
        const int r = fdiv(i, mg_m), m = i - r * a.n_mels;
        const float v = fmaxf(g0[i], floor_db);
        smel[r * MS + m] = v;
        const int t = lo + r;
        if (a.write_mspec && t >= t0 && t < t0 + nf) g0[i] = v;
      }
    }
  }
  __syncthreads();
  if (a.n_ceps <= 0) continue;
  // DCT (signal.py:1711): cep[r][c] = sum_m dct[c][m] * mel[r][m], m ascending from a zero accumulator.
  {
    const int half = (nrows + 1) >> 1;
    const int ncg = C8 >> 3;
    const uint32_t mg_half = fdiv_magic(half);
    for (int i = tid; i < half * ncg; i += 256) {
      const int g = fdiv(i, mg_half), r0 = i - g * half;
      const bool two = r0 + half < nrows;
      const int r1 = two ? r0 + half : r0;
      const float* m0 = smel + r0 * MS;
      const float* m1 = smel + r1 * MS;
      const float4* dq = reinterpret_cast<const float4*>(sdct + 8 * g);
      const int dstride = C8 >> 2;
      // (coefficient 2p, coefficient 2p + 1) per 64-bit accumulator: 8 FFMA2 per mel for the 2 x 8 block; each lane
      // of an FFMA2 is the fmaf of the scalar loop (m ascending from a zero accumulator), so the bits are unchanged
      u64 pa0[4] = {0ull, 0ull, 0ull, 0ull}, pa1[4] = {0ull, 0ull, 0ull, 0ull};
      const ulonglong2* dq2 = reinterpret_cast<const ulonglong2*>(dq);
#pragma unroll 4
      for (int m = 0; m < a.n_mels; ++m) {
        const float x0 = m0[m], x1 = m1[m];
        const ulonglong2 da = dq2[m * dstride], db = dq2[m * dstride + 1];
        const u64 xx0 = pk2(x0, x0), xx1 = pk2(x1, x1);
        pa0[0] = fma2(da.x, xx0, pa0[0]); pa1[0] = fma2(da.x, xx1, pa1[0]);
        pa0[1] = fma2(da.y, xx0, pa0[1]); pa1[1] = fma2(da.y, xx1, pa1[1]);
        pa0[2] = fma2(db.x, xx0, pa0[2]); pa1[2] = fma2(db.x, xx1, pa1[2]);
        pa0[3] = fma2(db.y, xx0, pa0[3]); pa1[3] = fma2(db.y, xx1, pa1[3]);
      }
      float acc0[8], acc1[8];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        acc0[2 * q] = lo32(pa0[q]); acc0[2 * q + 1] = hi32(pa0[q]);
        acc1[2 * q] = lo32(pa1[q]); acc1[2 * q + 1] = hi32(pa1[q]);
      }
      const int c0i = 8 * g;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0i + j < a.n_c1) {
          scep[r0 * a.n_c1 + c0i + j] = acc0[j];
          if (two) scep[r1 * a.n_c1 + c0i + j] = acc1[j];
        }
      if (g == 0 && a.c0 != nullptr) {
        const int ta = lo + r0, tb = lo + r1;
        if (ta >= t0 && ta < t0 + nf) a.c0[base + ta] = acc0[0];
        if (two && tb >= t0 && tb < t0 + nf) a.c0[base + tb] = acc1[0];
      }
    }
  }
  __syncthreads();   // scep complete; smel may now be overwritten by sd1
  if (a.feat == nullptr) continue;
  const int fd = a.n_ceps * (a.order + 1);
  // first-order deltas D(u) for u in [ulo, t0+nf); a thread takes four consecutive rows of one coefficient
  // so that the W + 3 inputs are read once (fir4, W == 9); rows that touch an utterance edge go one by one
  const int ulo = t0 - ((a.order >= 2) ? a.W - 1 : 0);
  const int nd = t0 + nf - ulo;
  float* sd2 = sd1 + nd * a.n_ceps;   // [nf][n_ceps] second-order deltas
  const bool pairs = a.W == 9 && (a.n_ceps & 1) == 0;   // packed route: a thread takes 4 rows x 2 adjacent coefficients
  if (a.order >= 1) {
    const int ng = (nd + 3) >> 2;
    const int ncol = pairs ? a.n_ceps >> 1 : a.n_ceps, cw = pairs ? 2 : 1;
    const uint32_t mg_col = fdiv_magic(ncol);
    for (int i = tid; i < ng * ncol; i += 256) {
      const int rg = fdiv(i, mg_col), c = (i - rg * ncol) * cw;
      const int r0 = 4 * rg, uu0 = ulo + r0;
      if (a.W == 9 && r0 + 3 < nd && uu0 - h >= 0 && uu0 + 3 + h <= T - 1) {
        if (pairs) {
          u64 o4[4];
          fir4x2<9>(scep + (uu0 - h - lo) * a.n_c1 + 1 + c, a.n_c1, tp9, o4);
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<u64*>(sd1 + (r0 + j) * a.n_ceps + c) = o4[j];
        } else {
          float o4[4];
          fir4<9>(scep + (uu0 - h - lo) * a.n_c1 + 1 + c, a.n_c1, staps, o4);
#pragma unroll
          for (int j = 0; j < 4; ++j) sd1[(r0 + j) * a.n_ceps + c] = o4[j];
        }
        continue;
      }
      for (int cc = c; cc < c + cw; ++cc)
      for (int r = r0; r < min(r0 + 4, nd); ++r) {
        const int uu = ulo + r;
        float acc = 0.f;
        if (uu - h >= 0 && uu + h <= T - 1) {   // interior: no clamping
          const float* p = scep + (uu + h - lo) * a.n_c1 + 1 + cc;
          for (int k = 0; k < a.W; ++k) acc = fmaf(staps[k], p[-k * a.n_c1], acc);
        } else if (uu >= -(h + 1)) {
          for (int k = 0; k < a.W; ++k) {
            int t = min(max(uu + h - k, 0), T - 1);
            acc = fmaf(staps[k], scep[(t - lo) * a.n_c1 + 1 + cc], acc);
          }
        } else {  // zero initial state of the causal filter
          const int j = uu + 2 * a.W - h - 1;
          float ts = 0.f;
          for (int k = 0; k <= j; ++k) ts += staps[k];
          acc = ts * scep[(0 - lo) * a.n_c1 + 1 + cc];
        }
        sd1[r * a.n_ceps + cc] = acc;
      }
    }
  }
  if (a.order >= 2) {
    __syncthreads();
    // second-order deltas: DD(t) = sum_k taps[k] D(t - k), every input row is held (no edges)
    const int ng = (nf + 3) >> 2;
    const int ncol = pairs ? a.n_ceps >> 1 : a.n_ceps, cw = pairs ? 2 : 1;
    const uint32_t mg_col = fdiv_magic(ncol);
    for (int i = tid; i < ng * ncol; i += 256) {
      const int rg = fdiv(i, mg_col), c = (i - rg * ncol) * cw;
      const int r0 = 4 * rg;   // row t = t0 + r0, its D row index is t - ulo = r0 + W - 1
      if (a.W == 9 && r0 + 3 < nf) {
        if (pairs) {
          u64 o4[4];
          fir4x2<9>(sd1 + r0 * a.n_ceps + c, a.n_ceps, tp9, o4);
#pragma unroll
          for (int j = 0; j < 4; ++j) *reinterpret_cast<u64*>(sd2 + (r0 + j) * a.n_ceps + c) = o4[j];
        } else {
          float o4[4];
          fir4<9>(sd1 + r0 * a.n_ceps + c, a.n_ceps, staps, o4);
#pragma unroll
          for (int j = 0; j < 4; ++j) sd2[(r0 + j) * a.n_ceps + c] = o4[j];
        }
        continue;
      }
      for (int cc = c; cc < c + cw; ++cc)
      for (int r = r0; r < min(r0 + 4, nf); ++r) {
        float acc = 0.f;
        const float* p = sd1 + (r + a.W - 1) * a.n_ceps + cc;
        for (int k = 0; k < a.W; ++k) acc = fmaf(staps[k], p[-k * a.n_ceps], acc);
        sd2[r * a.n_ceps + cc] = acc;
      }
    }
  }
  __syncthreads();
  float* fout = a.feat + (base + t0) * fd;
  if ((a.n_ceps & 3) == 0 && (reinterpret_cast<uintptr_t>(a.feat) & 15) == 0) {
    // 16 bytes per store: a quad never straddles the static / delta / delta-delta blocks of a row
    const int qrow = fd >> 2, qblk = a.n_ceps >> 2;
    const uint32_t mg_qrow = fdiv_magic(qrow), mg_qblk = fdiv_magic(qblk);
    float4* fout4 = reinterpret_cast<float4*>(fout);
    for (int i = tid; i < nf * qrow; i += 256) {
      const int r = fdiv(i, mg_qrow), q = i - r * qrow;
      const int o = fdiv(q, mg_qblk), c = 4 * (q - o * qblk);
      const float* src = (o == 0) ? scep + (t0 + r - lo) * a.n_c1 + 1 + c
                                  : (o == 1 ? sd1 + (t0 + r - ulo) * a.n_ceps + c : sd2 + r * a.n_ceps + c);
      fout4[i] = make_float4(src[0], src[1], src[2], src[3]);
    }
  } else {
    const uint32_t mg_fd = fdiv_magic(fd);
    for (int i = tid; i < nf * fd; i += 256) {
      const int r = fdiv(i, mg_fd), j = i - r * fd;
      const int o = fdiv(j, mg_ceps), c = j - o * a.n_ceps;
      const float* src = (o == 0) ? scep + (t0 + r - lo) * a.n_c1 + 1 + c
                                  : (o == 1 ? sd1 + (t0 + r - ulo) * a.n_ceps + c : sd2 + r * a.n_ceps + c);
      fout[i] = *src;
    }
  }
  }  // tile loop
}

// ---------------------------------------------------------------------------
// fe_post9_kernel: the utterance pass for the usual shape -- 9-tap deltas of order 2, n_mels and n_ceps multiples of 4,
// 16-byte aligned buffers.  Same arithmetic, in the same order, as fe_post_kernel (outputs are bit-identical); what
// changes is how the data moves.  ncu of fe_post_kernel on 25 h of config 3: shared-memory pipe 67 % busy with 38 % of
// its wavefronts bank conflicts (scalar stores of the mel rows 4-way, delta inputs at an odd row stride, scalar reads of
// the output phase 4-way), 51 % of the issue slots.  Here every shared-memory access of the hot phases is a 16- or
// 8-byte access at a row stride of 4 (mod 8) words, which is conflict-free per quarter warp:
//   * mel rows [NR][MS] stored and read as float4 (lanes walk rows in the DCT);
//   * cepstra [NR][CS] without c0 (the DCT basis is permuted: slots 0..n_ceps-1 the cepstra, slot n_ceps c0, which goes
//     straight to global memory), so a thread stores its 8 coefficients as two float4 and the deltas read (c, c + 1) as one
//     64-bit word;
//   * deltas as 6-row x 2-coefficient FIR blocks on packed pairs (14 reads for 12 outputs; 230 / 220 work items fill the
//     256 threads in one round);
//   * the output phase moves one float4 per load and store.
// ---------------------------------------------------------------------------
constexpr int P9_R = 6;
__host__ __device__ __forceinline__ int pad4odd(int n) { return ((n >> 2) & 1) ? n : n + 4; }

template <int R>
__device__ __forceinline__ void fir9x2(const float* __restrict__ x, int stride, const float (&tp)[9], float* __restrict__ y) {
  u64 v[R + 8];
#pragma unroll
  for (int q = 0; q < R + 8; ++q) v[q] = *reinterpret_cast<const u64*>(x + q * stride);
#pragma unroll
  for (int j = 0; j < R; ++j) {
    u64 acc = 0ull;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc = fma2(pk2(tp[k], tp[k]), v[j + 8 - k], acc);
    *reinterpret_cast<u64*>(y + j * stride) = acc;
  }
}

__global__ void __launch_bounds__(256, 3) fe_post9_kernel(PostArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  constexpr int W = 9, h = 4, HL = W - 1 + h, HR = h, NR = PT + HL + HR;
  const int MS = pad4odd(a.n_mels), CS = pad4odd(a.n_ceps);
  const int C8 = (a.n_c1 + 7) & ~7;
  float* sdct = sm;                        // [n_mels][C8] transposed, permuted basis
  float* scep = sdct + a.n_mels * C8;      // [NR][CS]
  float* smel = scep + NR * CS;            // [NR][MS]
  float* sd1 = smel;                       // [nd][CS] first-order deltas: the mel rows are dead by then
  const int nq = a.n_mels >> 2, qblk = a.n_ceps >> 2, qrow = 3 * qblk, ncol = a.n_ceps >> 1;
  const uint32_t mg_c8 = fdiv_magic(C8), mg_nq = fdiv_magic(nq), mg_col = fdiv_magic(ncol);
  const uint32_t mg_qrow = fdiv_magic(qrow), mg_qblk = fdiv_magic(qblk);
  for (int i = tid; i < a.n_mels * C8; i += 256) {
    const int m = fdiv(i, mg_c8), sl = i - m * C8;
    sdct[i] = (sl < a.n_ceps) ? a.dct[(sl + 1) * a.n_mels + m] : (sl == a.n_ceps ? a.dct[m] : 0.f);
  }
  float tp[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) tp[k] = a.taps[k];
  const int64_t per = (a.n_tiles + gridDim.x - 1) / gridDim.x;
  const int64_t tile_lo = per * blockIdx.x, tile_hi = min(a.n_tiles, tile_lo + per);
  if (tile_lo >= tile_hi) return;
  int u = find_segment(a.tile2_off, a.n_utt, tile_lo);
  int64_t u_end = a.tile2_off[u + 1];
  for (int64_t tile = tile_lo; tile < tile_hi; ++tile) {
    while (tile >= u_end) { ++u; u_end = a.tile2_off[u + 1]; }
    const int64_t base = a.frame_off[u];
    const int T = (int)(a.frame_off[u + 1] - base);
    const int t0 = (int)(tile - a.tile2_off[u]) * PT;
    const int nf = min(PT, T - t0);
    const int lo = max(0, t0 - HL), hi = min(T, t0 + nf + HR);
    const int nrows = hi - lo;
    const float floor_db = (a.top_db >= 0.f) ? ordered_to_float(a.umax[u]) - a.top_db : -FLT_MAX;
    __syncthreads();   // the previous tile is done with the buffers (and the table fill on the first trip)
    {
      float4* g4 = reinterpret_cast<float4*>(a.mspec + (base + lo) * a.n_mels);
      const int n4 = nrows * nq;
      for (int i0 = tid; i0 < n4; i0 += 4 * 256) {
        float4 vv[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k * 256;
          vv[k] = (i < n4) ? g4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int i = i0 + k * 256;
          if (i < n4) {
            const int r = fdiv(i, mg_nq), q = i - r * nq;
            float4 v = vv[k];
            const bool clip = fminf(fminf(v.x, v.y), fminf(v.z, v.w)) < floor_db;
            v.x = fmaxf(v.x, floor_db); v.y = fmaxf(v.y, floor_db);
            v.z = fmaxf(v.z, floor_db); v.w = fmaxf(v.w, floor_db);
            *reinterpret_cast<float4*>(smel + r * MS + 4 * q) = v;
            const int t = lo + r;
            if (clip && a.write_mspec && t >= t0 && t < t0 + nf) g4[i] = v;   // unclipped rows are already final
          }
        }
      }
    }
    __syncthreads();
    // DCT (signal.py:1711), 2 rows x 8 slots per thread, m ascending from a zero accumulator
    {
      const int half = (nrows + 1) >> 1;
      const int ncg = C8 >> 3, dstride = C8 >> 2;
      const uint32_t mg_half = fdiv_magic(half);
      for (int i = tid; i < half * ncg; i += 256) {
        const int g = fdiv(i, mg_half), r0 = i - g * half;
        const bool two = r0 + half < nrows;
        const int r1 = two ? r0 + half : r0;
        const float4* m0 = reinterpret_cast<const float4*>(smel + r0 * MS);
        const float4* m1 = reinterpret_cast<const float4*>(smel + r1 * MS);
        const ulonglong2* dq = reinterpret_cast<const ulonglong2*>(sdct + 8 * g);
        u64 pa0[4] = {0ull, 0ull, 0ull, 0ull}, pa1[4] = {0ull, 0ull, 0ull, 0ull};
#pragma unroll 2
        for (int m4 = 0; m4 < nq; ++m4) {
          const float4 xa = m0[m4], xb = m1[m4];
          const float x0[4] = {xa.x, xa.y, xa.z, xa.w}, x1[4] = {xb.x, xb.y, xb.z, xb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const ulonglong2 da = dq[(4 * m4 + e) * dstride], db = dq[(4 * m4 + e) * dstride + 1];
            const u64 xx0 = pk2(x0[e], x0[e]), xx1 = pk2(x1[e], x1[e]);
            pa0[0] = fma2(da.x, xx0, pa0[0]); pa1[0] = fma2(da.x, xx1, pa1[0]);
            pa0[1] = fma2(da.y, xx0, pa0[1]); pa1[1] = fma2(da.y, xx1, pa1[1]);
            pa0[2] = fma2(db.x, xx0, pa0[2]); pa1[2] = fma2(db.x, xx1, pa1[2]);
            pa0[3] = fma2(db.y, xx0, pa0[3]); pa1[3] = fma2(db.y, xx1, pa1[3]);
          }
        }
        const int s0 = 8 * g;
        if (s0 + 3 < a.n_ceps) {
          *reinterpret_cast<ulonglong2*>(scep + r0 * CS + s0) = make_ulonglong2(pa0[0], pa0[1]);
          if (two) *reinterpret_cast<ulonglong2*>(scep + r1 * CS + s0) = make_ulonglong2(pa1[0], pa1[1]);
        }
        if (s0 + 7 < a.n_ceps) {
          *reinterpret_cast<ulonglong2*>(scep + r0 * CS + s0 + 4) = make_ulonglong2(pa0[2], pa0[3]);
          if (two) *reinterpret_cast<ulonglong2*>(scep + r1 * CS + s0 + 4) = make_ulonglong2(pa1[2], pa1[3]);
        }
        if (a.c0 != nullptr && (s0 == a.n_ceps || s0 + 4 == a.n_ceps)) {
          const bool first = s0 == a.n_ceps;
          const int ta = lo + r0, tb = lo + r1;
          if (ta >= t0 && ta < t0 + nf) a.c0[base + ta] = lo32(first ? pa0[0] : pa0[2]);
          if (two && tb >= t0 && tb < t0 + nf) a.c0[base + tb] = lo32(first ? pa1[0] : pa1[2]);
        }
      }
    }
    __syncthreads();   // scep complete; smel may now be overwritten by sd1
    // first-order deltas D(uu) for uu in [ulo, t0 + nf): blocks of P9_R rows x 2 coefficients; rows that touch an
    // utterance edge go one by one
    const int ulo = t0 - (W - 1);
    const int nd = t0 + nf - ulo;
    float* sd2 = sd1 + nd * CS;   // [nf][CS] second-order deltas
    {
      const int ng = (nd + P9_R - 1) / P9_R;
      for (int i = tid; i < ng * ncol; i += 256) {
        const int rg = fdiv(i, mg_col), c = (i - rg * ncol) * 2;
        const int r0 = P9_R * rg, uu0 = ulo + r0;
        if (r0 + P9_R - 1 < nd && uu0 - h >= 0 && uu0 + P9_R - 1 + h <= T - 1) {
          fir9x2<P9_R>(scep + (uu0 - h - lo) * CS + c, CS, tp, sd1 + r0 * CS + c);
          continue;
        }
        for (int cc = c; cc < c + 2; ++cc)
          for (int r = r0; r < min(r0 + P9_R, nd); ++r) {
            const int uu = ulo + r;
            float acc = 0.f;
            if (uu - h >= 0 && uu + h <= T - 1) {   // interior: no clamping
              const float* p = scep + (uu + h - lo) * CS + cc;
#pragma unroll
              for (int k = 0; k < W; ++k) acc = fmaf(tp[k], p[-k * CS], acc);
            } else if (uu >= -(h + 1)) {
#pragma unroll
              for (int k = 0; k < W; ++k) {
                const int t = min(max(uu + h - k, 0), T - 1);
                acc = fmaf(tp[k], scep[(t - lo) * CS + cc], acc);
              }
            } else {  // zero initial state of the causal filter
              const int j = uu + 2 * W - h - 1;
              float ts = 0.f;
#pragma unroll
              for (int k = 0; k < W; ++k) if (k <= j) ts += tp[k];
              acc = ts * scep[(0 - lo) * CS + cc];
            }
            sd1[r * CS + cc] = acc;
          }
      }
    }
    __syncthreads();
    // second-order deltas: DD(t) = sum_k taps[k] D(t - k), every input row is held (no edges)
    {
      const int ng = (nf + P9_R - 1) / P9_R;
      for (int i = tid; i < ng * ncol; i += 256) {
        const int rg = fdiv(i, mg_col), c = (i - rg * ncol) * 2;
        const int r0 = P9_R * rg;   // row t = t0 + r0, its D row index is t - ulo = r0 + W - 1
        if (r0 + P9_R - 1 < nf) {
          fir9x2<P9_R>(sd1 + r0 * CS + c, CS, tp, sd2 + r0 * CS + c);
          continue;
        }
        for (int cc = c; cc < c + 2; ++cc)
          for (int r = r0; r < min(r0 + P9_R, nf); ++r) {
            float acc = 0.f;
            const float* p = sd1 + (r + W - 1) * CS + cc;
#pragma unroll
            for (int k = 0; k < W; ++k) acc = fmaf(tp[k], p[-k * CS], acc);
            sd2[r * CS + cc] = acc;
          }
      }
    }
    __syncthreads();
    // rows [static | delta | delta-delta], one float4 per load and store
    float4* fout4 = reinterpret_cast<float4*>(a.feat + (base + t0) * (3 * a.n_ceps));
    for (int i = tid; i < nf * qrow; i += 256) {
      const int r = fdiv(i, mg_qrow), q = i - r * qrow;
      const int o = fdiv(q, mg_qblk), c = 4 * (q - o * qblk);
      const float* src = (o == 0) ? scep + (t0 + r - lo) * CS + c
                                  : (o == 1 ? sd1 + (t0 + r - ulo) * CS + c : sd2 + r * CS + c);
      fout4[i] = *reinterpret_cast<const float4*>(src);
    }
  }  // tile loop
}

// ---------------------------------------------------------------------------
// VAD: one warp per utterance
// ---------------------------------------------------------------------------
struct VadArgs {
  const int64_t* order;   // SADgmm: utterance visited by cluster i (nullable = identity)
  const int64_t* frame_off;
  int n_utt;
  const float* x;   // energy [T] (SADgmm) or c0 [T] (SADthreshold)
  uint8_t* sad;     // [T]
  double* thr_out;  // [n_utt] nullable
  float* scratch;   // [T] standardised / normalised values
  int nmix, iters, smooth;
  double mode;
  double thr_energy, thr_mean_scale, thr_proportion;
  int thr_context;
};

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }

// SADgmm: one thread-block CLUSTER of 1..8 CTAs per utterance (sized per launch so that the
// clusters of a batch roughly fill the GPU: few long utterances get 8 CTAs each, thousands get 1).
//
// The 1-D EM of sklearn's GaussianMixture is ~350 fp64 instructions per frame and iteration, so a
// minute-long utterance on one SM is bound by that SM's fp64 pipe (the batch's longest utterance set
// the kernel time).  The frames of an utterance are therefore spread over the CTAs of a cluster; the
// thirteen sufficient statistics are reduced lanes -> warps -> CTA -> cluster (distributed shared
// memory) in a FIXED order, and every thread of every CTA then runs the tiny M-step redundantly on
// bit-identical numbers, so control flow (convergence test, component-drop retry) stays uniform.
constexpr int VAD_THREADS = 256;
constexpr int VAD_WARPS = VAD_THREADS / 32;
constexpr int VAD_CL_MAX = 8;               // portable cluster size limit
constexpr int VAD_KMAX = 4;
constexpr int VAD_NRED = 3 * VAD_KMAX + 2;   // lsum, nk[4], sx[4], sxx[4], #non-finite
constexpr int VAD_XR = 16;                  // frames per thread staged in shared memory by the EM
constexpr int VAD_WARP_MAX_T = 512;          // utterances up to this many frames: one warp each (fe_vad_gmm_warp_kernel)
constexpr int VAD_MAXLEAF = 1024;            // leaves of numpy's pairwise sum held in smem (n <~ 58 000 frames)

struct VadShared {
  double red[VAD_WARPS][VAD_NRED];
  double part[2][VAD_CL_MAX][VAD_NRED];   // [parity][source CTA]: every CTA pushes its partials to every CTA
  double tot[VAD_NRED];
  double par[VAD_KMAX][8];                // per component: prec, a0, b0, ld, lw, mu, pch, w
  int bad[VAD_KMAX];                      // M-step outcome per component: variance collapsed
  int parity;
  float xs[VAD_XR * VAD_THREADS];         // [j][tid]: the j-th frame of thread tid (standardised energies)
  float ms[2];
  int nleaf;
  int leaf_lo[VAD_MAXLEAF];
  int leaf_cnt[VAD_MAXLEAF];
  float leaf_sum[VAD_MAXLEAF];
  int stk_a[40], stk_b[40];                 // thread 0's explicit recursion stack (kept out of local memory)
  float stk_f[40];
};

// cluster-wide sums of v[0..VAD_NRED) -> sh.tot (same bits in every CTA; visible after the caller's next
// __syncthreads).  Every CTA pushes its partials into every CTA's shared memory (DSMEM stores), so ONE
// cluster barrier per reduction is enough; the slots alternate with the call parity, which keeps a fast CTA's
// next push away from the slots a slow CTA is still reading (it cannot be two reductions ahead).
__device__ __forceinline__ void vad_cluster_reduce(double (&v)[VAD_NRED], VadShared& sh, cg::cluster_group& cl) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int k = 0; k < VAD_NRED; ++k) v[k] = warp_sum(v[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < VAD_NRED; ++k) sh.red[warp][k] = v[k];
  }
  __syncthreads();
  const int par = sh.parity;
  const unsigned nb = cl.num_blocks(), me = cl.block_rank();
  if (tid < VAD_NRED) {
    double t = 0.0;
    for (int wv = 0; wv < VAD_WARPS; ++wv) t += sh.red[wv][tid];
    for (unsigned r = 0; r < nb; ++r) *cl.map_shared_rank(&sh.part[par][me][tid], r) = t;
  }
  cl.sync();
  if (tid < VAD_NRED) {
    double t = 0.0;
    for (unsigned r = 0; r < nb; ++r) t += sh.part[par][r][tid];
    sh.tot[tid] = t;
  }
  if (tid == 32) sh.parity = par ^ 1;   // read by everybody before the barrier above, next read after the caller's
}

// numpy's pairwise float32 sum of f(0..n) (see fe_logic.cuh), block-cooperative and bit-identical:
// thread 0 lists the <= 128-element leaves of the recursion, 8-lane groups sum one leaf each with the
// eight strided accumulators of numpy's unrolled loop, thread 0 folds the leaf sums along the same tree.
template <class F>
__device__ float block_np_pairwise_sum_f32(F f, int n, VadShared& sh, bool reuse_leaves = false) {
  const int tid = threadIdx.x;
  if (tid == 0 && !reuse_leaves) {   // (the table depends on n only)
    int *s_lo = sh.stk_a, *s_n = sh.stk_b, sp = 0, nl = 0;
    s_lo[0] = 0; s_n[0] = n; sp = 1;
    while (sp > 0 && nl >= 0) {
      const int lo = s_lo[sp - 1], cnt = s_n[sp - 1];
      --sp;
      if (cnt <= 128) {
        if (nl >= VAD_MAXLEAF) { nl = -1; break; }
        sh.leaf_lo[nl] = lo; sh.leaf_cnt[nl] = cnt; ++nl;
      } else {
        int n2 = cnt / 2;
        n2 -= n2 % 8;
        s_lo[sp] = lo + n2; s_n[sp] = cnt - n2; ++sp;   // right half is visited after ...
        s_lo[sp] = lo; s_n[sp] = n2; ++sp;              // ... the left half
      }
    }
    sh.nleaf = nl;
  }
  __syncthreads();
  const int nleaf = sh.nleaf;
  if (nleaf < 0) {  // too long for the leaf table: warp 0 walks the tree on its own
    if (tid < 32) {
      const float r = warp_np_pairwise_sum_f32(f, n, tid);
      if (tid == 0) sh.leaf_sum[0] = r;
    }
    __syncthreads();
    const float r = sh.leaf_sum[0];
    __syncthreads();
    return r;
  }
  const int gid = tid >> 3, l8 = tid & 7;
  for (int p0 = 0; p0 < nleaf; p0 += VAD_THREADS / 8) {
    const int leaf = p0 + gid;
    const bool active = leaf < nleaf;
    const int lo = active ? sh.leaf_lo[leaf] : 0, cnt = active ? sh.leaf_cnt[leaf] : 0;
    float r = 0.f;
    const int body = cnt - (cnt % 8);
    if (cnt >= 8) {
      r = f(lo + l8);
      for (int i = 8; i < body; i += 8) r = __fadd_rn(r, f(lo + i + l8));
    }
    __syncwarp();
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    if (active && l8 == 0) {
      float res = r;
      if (cnt < 8) {
        res = 0.f;
        for (int i = 0; i < cnt; ++i) res = __fadd_rn(res, f(lo + i));
      } else {
        for (int i = body; i < cnt; ++i) res = __fadd_rn(res, f(lo + i));
      }
      sh.leaf_sum[leaf] = res;
    }
  }
  __syncthreads();
  if (tid == 0) {
    // fold: post-order evaluation of the same recursion, leaves consumed left to right
    int *s_n = sh.stk_a, *s_stage = sh.stk_b, sp = 1, next = 0;
    float *s_left = sh.stk_f, ret = 0.f;
    s_n[0] = n; s_stage[0] = 0; s_left[0] = 0.f;
    while (sp > 0) {
      const int t = sp - 1, cnt = s_n[t];
      if (cnt <= 128) { ret = sh.leaf_sum[next++]; --sp; continue; }
      int n2 = cnt / 2;
      n2 -= n2 % 8;
      if (s_stage[t] == 0) { s_stage[t] = 1; s_n[sp] = n2; s_stage[sp] = 0; ++sp; }
      else if (s_stage[t] == 1) { s_left[t] = ret; s_stage[t] = 2; s_n[sp] = cnt - n2; s_stage[sp] = 0; ++sp; }
      else { ret = __fadd_rn(s_left[t], ret); --sp; }
    }
    sh.leaf_sum[0] = ret;
  }
  __syncthreads();
  const float r = sh.leaf_sum[0];
  __syncthreads();
  return r;
}

// 1-D EM of sklearn GaussianMixture with fixed inits (see oracle/frontend.py _em_1d).
// Returns false where sklearn would raise ValueError.
//
// Who computes what: the per-frame E-step is spread over all threads of the cluster (a thread's frames
// are staged in shared memory across iterations when there are <= VAD_XR of them); the sufficient statistics are
// reduced in a fixed order (vad_cluster_reduce); the M-step and the derived constants of component k
// (two fp64 logs, divisions, a square root: ~1 000 instructions when every thread replayed them) are
// computed by thread k of every CTA and broadcast through shared memory.  All CTAs see identical bits, so
// the convergence test and the component-drop retry stay uniform across the cluster.
enum { VP_PREC = 0, VP_A0, VP_B0, VP_LD, VP_LW, VP_MU, VP_PCH, VP_W };

__device__ __forceinline__ void vad_store_params(VadShared& sh, int k, double mu, double pch, double w) {
  const double prec = dmul(pch, pch);
  sh.par[k][VP_PREC] = prec;
  sh.par[k][VP_A0] = dmul(dmul(mu, mu), prec);
  sh.par[k][VP_B0] = dmul(mu, prec);
  sh.par[k][VP_LD] = log(pch);
  sh.par[k][VP_LW] = log(w);
  sh.par[k][VP_MU] = mu;
  sh.par[k][VP_PCH] = pch;
  sh.par[k][VP_W] = w;
}

// E-step of one frame (sklearn _estimate_log_prob_resp, 1-D) accumulated into acc[] = lsum | nk | sx | sxx | #non-finite.
// The component count is a template parameter: with a run-time count the parameter and accumulator arrays were
// indexed dynamically and lived in local memory (314 LDL / 285 STL in the kernel, a 1 376-byte stack frame).
template <int NC>
__device__ __forceinline__ void vad_frame(const float xf, const double (&prec)[NC], const double (&a0)[NC],
                                          const double (&b0)[NC], const double (&ld)[NC], const double (&lw)[NC],
                                          double (&acc)[VAD_NRED], double& sprod, int& nprod) {
  constexpr int KMAX = VAD_KMAX;
  const double LOG2PI = 1.8378770664093454835606594728112;
  if (!isfinite(xf)) acc[VAD_NRED - 1] += 1.0;
  const double xd = (double)xf, x2 = (double)__fmul_rn(xf, xf);  // x*x is float32 in sklearn
  double wl[NC], mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    double lp = dadd(dadd(a0[k], -dmul(2.0, dmul(xd, b0[k]))), dmul(x2, prec[k]));
    lp = dadd(dmul(-0.5, dadd(LOG2PI, lp)), ld[k]);
    wl[k] = dadd(lp, lw[k]);
    mx = fmax(mx, wl[k]);
  }
  double ek[NC], s = 0.0;
  if constexpr (NC == 3) {
    const bool am = wl[0] == mx, bm = wl[1] == mx, cm = wl[2] == mx;
    const double dA = wl[0] - mx, dB = wl[1] - mx, dC = wl[2] - mx;
    const double e1 = exp(am ? dB : dA), e2 = exp(cm ? dB : dC);
    ek[0] = am ? 1.0 : e1;
    ek[2] = cm ? 1.0 : e2;
    ek[1] = bm ? 1.0 : (am ? e1 : e2);
    s = (ek[0] + ek[1]) + ek[2];
  } else {
#pragma unroll
    for (int k = 0; k < NC; ++k) { ek[k] = exp(wl[k] - mx); s += ek[k]; }
  }
  const double inv = 1.0 / s;
  acc[0] += mx;
  sprod *= s;
  if (++nprod == 256) { acc[0] += log(sprod); sprod = 1.0; nprod = 0; }   // nc^256 stays finite
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const double r = ek[k] * inv;
    acc[1 + k] += r;
    acc[1 + KMAX + k] = fma(r, xd, acc[1 + KMAX + k]);
    acc[1 + 2 * KMAX + k] = fma(r, x2, acc[1 + 2 * KMAX + k]);
  }
}

// The E-step of one EM iteration over the frames get(0), get(1), ... of this thread; par[k] as vad_store_params left it.
template <int NC, class Get>
__device__ __forceinline__ void vad_estep(const double (*par)[8], int count, Get get, double (&acc)[VAD_NRED]) {
  double prec[NC], a0[NC], b0[NC], ld[NC], lw[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    prec[k] = par[k][VP_PREC]; a0[k] = par[k][VP_A0]; b0[k] = par[k][VP_B0];
    ld[k] = par[k][VP_LD]; lw[k] = par[k][VP_LW];
  }
#pragma unroll
  for (int k = 0; k < VAD_NRED; ++k) acc[k] = 0.0;
  // The per-frame work is ~all fp64 transcendental code, so it is trimmed to what the sums need:
  //  * exp(wl[k] - max) of the maximal component is exactly 1: with three components the two other
  //    differences are picked with selects and only TWO exponentials are evaluated;
  //  * log(s) only feeds the running sum of log-likelihoods: the s of a thread's frames are
  //    multiplied (s in [1, nc], <= a few dozen frames per thread and iteration) and ONE log is
  //    taken at the end;
  //  * the responsibilities exp(wl - norm) are formed as exp(wl - max) / s.
  // Each step differs from sklearn's expression by <= 1 ulp of a double.
  double sprod = 1.0;
  int nprod = 0;
#pragma unroll 2
  for (int j = 0; j < count; ++j) vad_frame<NC>(get(j), prec, a0, b0, ld, lw, acc, sprod, nprod);
  acc[0] += log(sprod);
}

__device__ bool vad_em(const float* __restrict__ x, int n, int nc, int max_iter, VadShared& sh, cg::cluster_group& cl,
                       double* mu_out, double* prec_out) {
  constexpr int KMAX = VAD_KMAX;
  const int tid = threadIdx.x;
  const int gt = (int)cl.block_rank() * VAD_THREADS + tid;
  const int gstride = (int)cl.num_blocks() * VAD_THREADS;
  if (n < max(nc, 2)) return false;
  const bool staged = n <= VAD_XR * gstride;   // this thread's frames fit its column of sh.xs
  if (staged) {
    for (int j = 0, i = gt; i < n; ++j, i += gstride) sh.xs[j * VAD_THREADS + tid] = x[i];
  }
  if (tid < nc) {
    vad_store_params(sh, tid, -2.0 + 4.0 * tid / (double)(nc - 1), 1.0, 1.0 / nc);
    sh.bad[tid] = 0;
  }
  __syncthreads();
  double lower = -INFINITY;
  for (int it = 0; it < max_iter; ++it) {
    double acc[VAD_NRED];
    {
      const int count = gt < n ? (n - gt + gstride - 1) / gstride : 0;   // frames gt, gt + gstride, ... of this thread
      const float* xcol = sh.xs + tid;
      const float* xg = x + gt;
      auto run = [&](auto nc_tag) {
        constexpr int NC = decltype(nc_tag)::value;
        if (staged) vad_estep<NC>(sh.par, count, [xcol](int j) { return xcol[j * VAD_THREADS]; }, acc);
        else vad_estep<NC>(sh.par, count, [xg, gstride](int j) { return xg[(size_t)j * gstride]; }, acc);
      };
      if (nc == 3) run(std::integral_constant<int, 3>());
      else if (nc == 2) run(std::integral_constant<int, 2>());
      else run(std::integral_constant<int, 4>());
    }
    vad_cluster_reduce(acc, sh, cl);
    // M-step (sklearn _estimate_gaussian_parameters, 1-D): component k by lane k of warp 0 -- the same warp
    // that wrote sh.tot, so a warp barrier orders the two
    if (tid < 32) {
      __syncwarp();
      if (tid < nc) {
        double nkk = 0.0, nksum = 0.0;   // (static indexing: no local-memory arrays)
#pragma unroll
        for (int k = 0; k < KMAX; ++k)
          if (k < nc) {
            const double v = sh.tot[1 + k] + 10.0 * DBL_EPSILON;
            nksum += v;
            if (k == tid) nkk = v;
          }
        const int k = tid;
        const double mu = sh.tot[1 + KMAX + k] / nkk;
        const double var = dadd(dadd(sh.tot[1 + 2 * KMAX + k] / nkk, -dmul(mu, mu)), 1e-6);
        sh.bad[k] = !(var > 0.0);
        vad_store_params(sh, k, mu, 1.0 / sqrt(var), nkk / nksum);
      }
    }
    __syncthreads();
    if (sh.tot[VAD_NRED - 1] != 0.0) return false;   // sklearn rejects non-finite input
    bool collapsed = false;
    for (int k = 0; k < nc; ++k) collapsed |= sh.bad[k] != 0;
    if (collapsed) return false;
    const double new_lower = sh.tot[0] / (double)n;
    const double change = new_lower - lower;
    lower = new_lower;
    if (fabs(change) < 1e-3) break;
  }
  // the component with the largest mean (first one on ties, like np.argmax)
  int kb = 0;
  for (int k = 1; k < nc; ++k) if (sh.par[k][VP_MU] > sh.par[kb][VP_MU]) kb = k;
  *mu_out = sh.par[kb][VP_MU];
  *prec_out = dmul(sh.par[kb][VP_PCH], sh.par[kb][VP_PCH]);
  return true;
}

__global__ void __launch_bounds__(VAD_THREADS, 2) fe_vad_gmm_kernel(VadArgs a) {
  __shared__ VadShared sh;
  cg::cluster_group cl = cg::this_cluster();
  const int tid = threadIdx.x;
  const int gt = (int)cl.block_rank() * VAD_THREADS + tid;
  const int ncta = (int)cl.num_blocks();       // cluster size, chosen per launch (1, 2, 4 or 8)
  const int gstride = ncta * VAD_THREADS;
  const int n_clusters = gridDim.x / ncta;
  if (tid == 0) sh.parity = 0;
  __syncthreads();
  for (int ci = blockIdx.x / ncta; ci < a.n_utt; ci += n_clusters) {
    const int u = a.order ? (int)a.order[ci] : ci;
    const int64_t base = a.frame_off[u];
    const int n = (int)(a.frame_off[u + 1] - base);
    if (n <= 0) continue;
    const float* e = a.x + base;
    float* xs = a.scratch + base;
    uint8_t* out = a.sad + base;
    int nc = a.nmix;
    bool ok = false;
    double thr = 0.0;
    const float* src = e;
    while (true) {
      // standardise in float32 exactly as numpy does (signal.py:305); the retry path of the
      // reference re-standardises the already standardised vector.  Every CTA of the cluster
      // evaluates mean / std redundantly (same bits), then writes its strided share of xs.
      const float ssum = block_np_pairwise_sum_f32([src](int i) { return src[i]; }, n, sh);
      const float mean = (float)((double)ssum / (double)n);
      const float ss = block_np_pairwise_sum_f32(
          [src, mean](int i) { const float d = __fadd_rn(src[i], -mean); return __fmul_rn(d, d); }, n, sh, true);
      const float sd = sqrtf((float)((double)ss / (double)n));
      cl.sync();  // every CTA is done reading src (== xs on the retry path) before it is overwritten
      for (int i = gt; i < n; i += gstride) xs[i] = __fdiv_rn(__fsub_rn(src[i], mean), sd);
      __threadfence();
      cl.sync();
      src = xs;
      double mu_b, prec_b;
      if (vad_em(xs, n, nc, a.iters, sh, cl, &mu_b, &prec_b)) {
        thr = dadd(mu_b, -dmul(a.mode, sqrt(1.0 / prec_b)));
        ok = true;
        break;
      }
      if (nc - 1 >= 2) { --nc; continue; }
      break;
    }
    if (a.thr_out != nullptr && gt == 0) a.thr_out[u] = ok ? thr : 0.0;
    if (!ok) {
      for (int i = gt; i < n; i += gstride) out[i] = 0;
    } else {
      auto raw = [xs, thr](int i) -> int { return ((double)xs[i] > thr) ? 1 : 0; };
      const bool do_smooth = a.smooth >= 3 && n >= a.smooth;
      for (int i = gt; i < n; i += gstride)
        out[i] = do_smooth ? (uint8_t)smooth_flat_ge_f(raw, n, a.smooth, false, i) : (uint8_t)raw(i);
    }
    cl.sync();  // xs / sh are reused by the cluster's next utterance
  }
}

// EM of one utterance by one warp (fe_vad_gmm_warp_kernel): returns false where sklearn would raise; mu / precision
// Cholesky factor of the component with the largest mean.
template <int NC>
__device__ bool vad_warp_em(const float* __restrict__ xs, int n, int max_iter, int lane, double& mu_b, double& pch_b) {
  constexpr int KMAX = VAD_KMAX;
  if (n < (NC > 2 ? NC : 2)) return false;
  double w[NC], mu[NC], pch[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) { w[k] = 1.0 / NC; mu[k] = -2.0 + 4.0 * k / (double)(NC - 1); pch[k] = 1.0; }
  double lower = -INFINITY;
  const int count = lane < n ? (n - lane + 31) / 32 : 0;
  const float* xl = xs + lane;
  for (int it = 0; it < max_iter; ++it) {
    double par[NC][8];
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      const double prec = dmul(pch[k], pch[k]);
      par[k][VP_PREC] = prec;
      par[k][VP_A0] = dmul(dmul(mu[k], mu[k]), prec);
      par[k][VP_B0] = dmul(mu[k], prec);
      par[k][VP_LD] = log(pch[k]);
      par[k][VP_LW] = log(w[k]);
    }
    double acc[VAD_NRED];
    vad_estep<NC>(par, count, [xl](int j) { return xl[32 * j]; }, acc);
#pragma unroll
    for (int k = 0; k < VAD_NRED; ++k) acc[k] = warp_sum(acc[k]);
    if (acc[VAD_NRED - 1] != 0.0) return false;   // sklearn rejects non-finite input
    double nk[NC], nksum = 0.0;
    bool collapsed = false;
#pragma unroll
    for (int k = 0; k < NC; ++k) { nk[k] = acc[1 + k] + 10.0 * DBL_EPSILON; nksum += nk[k]; }
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      mu[k] = acc[1 + KMAX + k] / nk[k];
      const double var = dadd(dadd(acc[1 + 2 * KMAX + k] / nk[k], -dmul(mu[k], mu[k])), 1e-6);
      if (!(var > 0.0)) collapsed = true;
      pch[k] = 1.0 / sqrt(var);
      w[k] = nk[k] / nksum;
    }
    if (collapsed) return false;
    const double new_lower = acc[0] / (double)n;
    const double change = new_lower - lower;
    lower = new_lower;
    if (fabs(change) < 1e-3) break;
  }
  mu_b = mu[0]; pch_b = pch[0];
#pragma unroll
  for (int k = 1; k < NC; ++k) if (mu[k] > mu_b) { mu_b = mu[k]; pch_b = pch[k]; }
  return true;
}

// SADgmm for SHORT utterances: one WARP per utterance, eight independent utterances per CTA, no block or cluster
// barrier anywhere.  A 3 s utterance (298 frames) leaves a 256-thread CTA with one frame per thread and ~25 EM
// iterations of pure barrier / M-step latency (ncu on 2 000 such utterances: 22 % of the instructions were the
// warp shuffles of the 14-value reduction, most stall samples sat on barriers); here every lane takes ~10 frames,
// the statistics are reduced with one xor-butterfly (bit-identical on all lanes, so every lane replays the tiny
// M-step on the same numbers) and the eight warps of a CTA hide each other's latency.
// Utterances [first, n_utt) of the visiting order (longest first) are handled here.
__global__ void __launch_bounds__(VAD_THREADS, 2) fe_vad_gmm_warp_kernel(VadArgs a, int first) {
  constexpr int KMAX = VAD_KMAX;
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * VAD_THREADS + threadIdx.x) >> 5;
  const int nw = (gridDim.x * VAD_THREADS) >> 5;
  for (int ci = first + wid; ci < a.n_utt; ci += nw) {
    const int u = a.order ? (int)a.order[ci] : ci;
    const int64_t base = a.frame_off[u];
    const int n = (int)(a.frame_off[u + 1] - base);
    if (n <= 0) continue;
    const float* e = a.x + base;
    float* xs = a.scratch + base;
    uint8_t* out = a.sad + base;
    int nc = a.nmix;
    bool ok = false;
    double thr = 0.0;
    const float* src = e;
    while (true) {
      // standardise in float32 exactly as numpy does (signal.py:305)
      const float ssum = warp_np_pairwise_sum_f32([src](int i) { return src[i]; }, n, lane);
      const float mean = (float)((double)ssum / (double)n);
      const float ss = warp_np_pairwise_sum_f32(
          [src, mean](int i) { const float d = __fadd_rn(src[i], -mean); return __fmul_rn(d, d); }, n, lane);
      const float sd = sqrtf((float)((double)ss / (double)n));
      __syncwarp();
      for (int i = lane; i < n; i += 32) xs[i] = __fdiv_rn(__fsub_rn(src[i], mean), sd);
      __syncwarp();
      src = xs;
      // ---- EM (same arithmetic as vad_em; reductions by butterfly) ----
      double mu_b = 0.0, pch_b = 1.0;
      const bool fitted = nc == 3 ? vad_warp_em<3>(xs, n, a.iters, lane, mu_b, pch_b)
                        : nc == 2 ? vad_warp_em<2>(xs, n, a.iters, lane, mu_b, pch_b)
                                  : vad_warp_em<4>(xs, n, a.iters, lane, mu_b, pch_b);
      if (fitted) {
        thr = dadd(mu_b, -dmul(a.mode, sqrt(1.0 / dmul(pch_b, pch_b))));
        ok = true;
        break;
      }
      if (nc - 1 >= 2) { --nc; continue; }
      break;
    }
    if (a.thr_out != nullptr && lane == 0) a.thr_out[u] = ok ? thr : 0.0;
    if (!ok) {
      for (int i = lane; i < n; i += 32) out[i] = 0;
    } else {
      auto raw = [xs, thr](int i) -> int { return ((double)xs[i] > thr) ? 1 : 0; };
      const bool do_smooth = a.smooth >= 3 && n >= a.smooth;
      for (int i = lane; i < n; i += 32)
        out[i] = do_smooth ? (uint8_t)smooth_flat_ge_f(raw, n, a.smooth, false, i) : (uint8_t)raw(i);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) fe_vad_thr_kernel(VadArgs a) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int u = wid; u < a.n_utt; u += nw) {
    const int64_t base = a.frame_off[u];
    const int n = (int)(a.frame_off[u + 1] - base);
    if (n <= 0) continue;
    const float* e = a.x + base;
    float* en = a.scratch + base;
    uint8_t* out = a.sad + base;
    float lo = FLT_MAX, hi = -FLT_MAX;
    for (int i = lane; i < n; i += 32) { lo = fminf(lo, e[i]); hi = fmaxf(hi, e[i]); }
    lo = warp_min(lo);
    hi = warp_max(hi);
    const float range = __fsub_rn(hi, lo);
    for (int i = lane; i < n; i += 32) en[i] = __fdiv_rn(__fsub_rn(e[i], lo), range);  // speech.py:1305-1307
    __syncwarp();
    double thr = a.thr_energy;
    if (a.thr_mean_scale != 0.0) {
      float s = 0.f;  // numba: sequential float32 accumulation (speech.py:1310,1328-1332)
      if (lane == 0)
        for (int i = 0; i < n; ++i) s = __fadd_rn(s, en[i]);
      s = __shfl_sync(0xffffffffu, s, 0);
      thr = dadd(thr, dmul(a.thr_mean_scale, (double)s) / (double)n);
    }
    if (a.thr_out != nullptr && lane == 0) a.thr_out[u] = thr;
    // context vote (speech.py:1313-1324) -> uint8, then uint8-wrapping smooth (speech.py:1426-1431)
    const int ctx = a.thr_context;
    const double prop = a.thr_proportion;
    auto vote = [en, n, thr, ctx, prop](int t) -> int {
      int num = 0, den = 0;
      for (int t2 = t - ctx; t2 <= t + ctx; ++t2)
        if (t2 >= 0 && t2 < n) { ++den; if ((double)en[t2] > thr) ++num; }
      return ((double)num >= (double)den * prop) ? 1 : 0;
    };
    const bool do_smooth = a.smooth >= 3 && n >= a.smooth;
    for (int i = lane; i < n; i += 32)
      out[i] = do_smooth ? (uint8_t)smooth_flat_ge_f(vote, n, a.smooth, true, i) : (uint8_t)vote(i);
  }
}

// ---------------------------------------------------------------------------
// ApplyingSAD compaction
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
fe_count_kernel(const uint8_t* __restrict__ sad, const int64_t* __restrict__ frame_off, int n_utt,
                int keep_unvoiced, int64_t* __restrict__ cnt) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int u = wid; u < n_utt; u += nw) {
    const int64_t b = frame_off[u], n = frame_off[u + 1] - b;
    int c = 0;
    for (int64_t i = lane; i < n; i += 32) c += sad[b + i] != 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (lane == 0) cnt[u] = (c == 0 && keep_unvoiced) ? -n : c;  // negative = keep all n frames
  }
}

__global__ void fe_scan_kernel(const int64_t* __restrict__ cnt, int n_utt, int64_t* __restrict__ out_off) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;  // n_utt is small; a serial scan is microseconds
  int64_t acc = 0;
  for (int u = 0; u < n_utt; ++u) {
    out_off[u] = acc;
    int64_t c = cnt[u];
    acc += c < 0 ? -c : c;
  }
  out_off[n_utt] = acc;
}

__global__ void __launch_bounds__(128)
fe_compact_kernel(const uint8_t* __restrict__ sad, const int64_t* __restrict__ frame_off,
                  const int64_t* __restrict__ cnt, const int64_t* __restrict__ out_off, int n_utt,
                  const float* __restrict__ feat, int dim, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int u = wid; u < n_utt; u += nw) {
    const int64_t b = frame_off[u], n = frame_off[u + 1] - b;
    const bool all = cnt[u] < 0;
    int64_t w = out_off[u];
    for (int64_t i0 = 0; i0 < n; i0 += 32) {
      const int64_t i = i0 + lane;
      const bool keep = (i < n) && (all || sad[b + i] != 0);
      const unsigned mask = __ballot_sync(0xffffffffu, keep);
      // rows of this group that survive, in order
      for (unsigned mm = mask; mm != 0; mm &= mm - 1) {
        const int src_lane = __ffs(mm) - 1;
        const int rank = __popc(mask & ((1u << src_lane) - 1));
        const float* s = feat + (b + i0 + src_lane) * dim;
        float* d = out + (w + rank) * dim;
        for (int j = lane; j < dim; j += 32) d[j] = s[j];
      }
      w += __popc(mask);
    }
  }
}

// ---------------------------------------------------------------------------
// host launcher
// ---------------------------------------------------------------------------
template <int N, typename T>
static size_t frame_smem(int L, int hop) {
  size_t b = (size_t)(L + (L & 1)) * sizeof(double);
  b += (size_t)N * sizeof(C2<T>);
  b += (size_t)FE_WARPS * padded_len<N>() * sizeof(C2<T>);
  b += (size_t)L * sizeof(float);
  b += (size_t)((FT - 1) * hop + L + 16 + 4) * sizeof(float);   // + misalignment slots and the 16-byte round-up (stage_pcm)
  return b;
}

template <int N, typename T, typename PCM>
static int launch_frame(const FrameArgs& a, cudaStream_t st) {
  size_t smem = frame_smem<N, T>(a.L, a.hop);
  if (smem > 227 * 1024) return set_error(ODIN_EINVAL, "frame kernel needs %zu B smem (hop too large)", smem);
  auto k = fe_frame_kernel<N, T, PCM>;
  ODIN_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(4, (227 * 1024) / (smem + 1024)));
  int64_t grid = std::min<int64_t>(a.n_tiles, (int64_t)sm_count() * per_sm);
  k<<<(unsigned)grid, FE_THREADS, smem, st>>>(a);
  ODIN_LAUNCH_CHECK("fe_frame_kernel");
  return ODIN_OK;
}

template <typename T, typename PCM>
static int dispatch_frame(int N, const FrameArgs& a, cudaStream_t st) {
  switch (N) {
    case 256: return launch_frame<256, T, PCM>(a, st);
    case 512: return launch_frame<512, T, PCM>(a, st);
    case 1024: return launch_frame<1024, T, PCM>(a, st);
    case 2048: return launch_frame<2048, T, PCM>(a, st);
  }
  return set_error(ODIN_EINVAL, "n_fft %d unsupported (256/512/1024/2048)", N);
}

template <int N, typename PCM>
static int launch_frame4(const FrameArgs& a, cudaStream_t st) {
  size_t smem = (size_t)(a.L + (a.L & 1)) * sizeof(double) + (size_t)N * sizeof(float2) +
                (size_t)FE_WARPS * (32 / (N / 32)) * f4_region<N>() * sizeof(float2) + (size_t)a.L * sizeof(float) +
                (size_t)a.mel_trips * 32 * sizeof(float2) + (size_t)(a.n_mels + 1) * sizeof(int) +
                (size_t)((FT - 1) * a.hop + a.L + 16 + 4) * sizeof(float);   // + stage_pcm's misalignment slots / round-up
  if (smem > 227 * 1024) return set_error(ODIN_EINVAL, "frame kernel needs %zu B smem (hop too large)", smem);
  auto k = (a.L <= N / 2) ? fe_frame4_kernel<N, PCM, true> : fe_frame4_kernel<N, PCM, false>;
  ODIN_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = (int)std::max<size_t>(1, std::min<size_t>(2, (227 * 1024) / (smem + 1024)));
  int64_t grid = std::min<int64_t>(a.n_tiles, (int64_t)sm_count() * per_sm);
  k<<<(unsigned)grid, FE_THREADS, smem, st>>>(a);
  ODIN_LAUNCH_CHECK("fe_frame4_kernel");
  return ODIN_OK;
}

template <typename PCM>
static int dispatch_frame4(int N, const FrameArgs& a, cudaStream_t st) {
  switch (N) {
    case 256: return launch_frame4<256, PCM>(a, st);
    case 512: return launch_frame4<512, PCM>(a, st);
    case 1024: return launch_frame4<1024, PCM>(a, st);
  }
  return set_error(ODIN_EINVAL, "n_fft %d unsupported by the four-step kernel", N);
}

// Launches the SAD kernels for a ragged batch: kind 1 = SADgmm (cluster kernel for long utterances, warp kernel for
// many short ones), otherwise SADthreshold.  `fo` = HOST frame offsets [n_utt + 1]; v.order must list the
// utterances longest first (or be null when the legacy order is forced).
static int vad_dispatch(const VadArgs& v, int kind, const int64_t* fo, int n_utt, cudaStream_t vst) {
    {
      int warps_per_cta = 4;
      int grid = (int)std::min<int64_t>(ceil_div(n_utt, warps_per_cta), (int64_t)sm_count() * 8);
      if (kind == 1) {
        // Utterances of up to VAD_WARP_MAX_T frames go to the warp-per-utterance kernel, longer ones to the
        // cluster kernel.  The visiting order is longest first, so the long ones are its first n_long entries.
        const char* lpt = getenv("ODIN_FE_VAD_LPT");
        const char* nowarp = getenv("ODIN_FE_VAD_NOWARP");   // A/B runs
        int n_long = n_utt;
        int64_t max_T = 0;
        for (int u = 0; u < n_utt; ++u) max_T = std::max(max_T, fo[u + 1] - fo[u]);
        if (!(lpt && lpt[0] == '0') && !(nowarp && nowarp[0] == '1')) {
          n_long = 0;
          for (int u = 0; u < n_utt; ++u) n_long += (fo[u + 1] - fo[u]) > VAD_WARP_MAX_T;
          // a warp takes ~2x longer over ONE utterance than a CTA does, so the warp kernel only pays once there are
          // enough short utterances to fill the SMs with warps (measured: 100 x 3 s 0.13 ms on CTAs vs 0.24 on warps;
          // 2 000 x 3 s 0.84 vs 0.29)
          if (n_utt - n_long < 4 * sm_count()) n_long = n_utt;
        }
        if (n_long > 0) {
          // cluster size: the largest power of two that keeps the batch within ~3 CTAs per SM (2 are
          // resident at once; measured on the config-3 shard, 216 utterances: 1 -> 0.82 ms, 2 -> 0.61,
          // 4 -> 0.66, 8 -> 1.17); clusters start longest utterance first and the hardware hands the
          // next one to whichever SMs free up
          int ncta = 1;
          while (ncta < VAD_CL_MAX && (int64_t)n_long * (ncta * 2) <= (int64_t)sm_count() * 3) ncta *= 2;
          // ... but never more CTAs than the longest utterance can feed (>= 4 frames per thread)
          while (ncta > 1 && max_T < (int64_t)4 * VAD_THREADS * ncta) ncta /= 2;
          if (const char* ev = getenv("ODIN_FE_VAD_NCTA")) {   // A/B runs
            const int v2 = atoi(ev);
            if (v2 == 1 || v2 == 2 || v2 == 4 || v2 == 8) ncta = v2;
          }
          VadArgs vl = v;
          vl.n_utt = n_long;
          const int64_t n_cl = std::min<int64_t>(n_long, (int64_t)sm_count() * 4);
          cudaLaunchConfig_t cfg = {};
          cfg.gridDim = dim3((unsigned)(n_cl * ncta));
          cfg.blockDim = dim3(VAD_THREADS);
          cfg.dynamicSmemBytes = 0;
          cfg.stream = vst;
          cudaLaunchAttribute attr[1];
          attr[0].id = cudaLaunchAttributeClusterDimension;
          attr[0].val.clusterDim.x = (unsigned)ncta;
          attr[0].val.clusterDim.y = 1;
          attr[0].val.clusterDim.z = 1;
          cfg.attrs = attr;
          cfg.numAttrs = 1;
          ODIN_CUDA_CHECK(cudaLaunchKernelEx(&cfg, fe_vad_gmm_kernel, vl));
          ODIN_LAUNCH_CHECK("fe_vad_gmm_kernel");
        }
        if (n_long < n_utt) {
          const int n_short = n_utt - n_long;
          const int wgrid = (int)std::min<int64_t>(ceil_div(n_short, VAD_WARPS), (int64_t)sm_count() * 8);
          fe_vad_gmm_warp_kernel<<<wgrid, VAD_THREADS, 0, vst>>>(v, n_long);
          ODIN_LAUNCH_CHECK("fe_vad_gmm_warp_kernel");
        }
      } else {
        fe_vad_thr_kernel<<<grid, 128, 0, vst>>>(v);
        ODIN_LAUNCH_CHECK("fe_vad_thr_kernel");
      }
    }
  return ODIN_OK;
}

// SAD on given per-frame energies, without a front-end handle (SADgmm / SADthreshold applied to any energy
// feature of a pipeline, signal.vad_split_audio): temporaries are stream-ordered allocations.
int fe_vad_standalone(int kind, const float* d_x, const int64_t* h_fo, int n_utt, int nmix, int iters, int smooth,
                      double mode, double thr_energy, double thr_mean_scale, double thr_proportion, int thr_context,
                      uint8_t* d_sad, double* d_thr, cudaStream_t st) {
  const int64_t T = h_fo[n_utt] - h_fo[0];
  if (T <= 0) return ODIN_OK;
  std::vector<int64_t> host(2 * (size_t)n_utt + 1);
  for (int u = 0; u <= n_utt; ++u) host[u] = h_fo[u] - h_fo[0];
  {
    std::vector<int> idx(n_utt);
    for (int u = 0; u < n_utt; ++u) idx[u] = u;
    std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return host[a + 1] - host[a] > host[b + 1] - host[b]; });
    for (int i = 0; i < n_utt; ++i) host[n_utt + 1 + i] = idx[i];
  }
  int64_t* d_meta = nullptr;
  float* d_scratch = nullptr;
  ODIN_CUDA_CHECK(cudaMallocAsync(&d_meta, sizeof(int64_t) * host.size(), st));
  cudaError_t e = cudaMallocAsync(&d_scratch, sizeof(float) * (size_t)(T + T / 8 + 256), st);
  if (e != cudaSuccess) { cudaFreeAsync(d_meta, st); return set_error(ODIN_ENOMEM, "SAD scratch: %s", cudaGetErrorString(e)); }
  // pageable source: the copy is staged before the call returns, so `host` may go out of scope
  e = cudaMemcpyAsync(d_meta, host.data(), sizeof(int64_t) * host.size(), cudaMemcpyHostToDevice, st);
  int rc = ODIN_OK;
  if (e != cudaSuccess) rc = set_error(ODIN_ECUDA, "cudaMemcpyAsync: %s", cudaGetErrorString(e));
  if (rc == ODIN_OK) {
    VadArgs v{};
    v.frame_off = d_meta; v.order = d_meta + n_utt + 1; v.n_utt = n_utt;
    v.x = d_x + h_fo[0]; v.sad = d_sad + h_fo[0]; v.thr_out = d_thr; v.scratch = d_scratch;
    v.nmix = nmix; v.iters = iters; v.smooth = smooth; v.mode = mode;
    v.thr_energy = thr_energy; v.thr_mean_scale = thr_mean_scale; v.thr_proportion = thr_proportion;
    v.thr_context = thr_context;
    rc = vad_dispatch(v, kind, host.data(), n_utt, st);
  }
  cudaFreeAsync(d_scratch, st);
  cudaFreeAsync(d_meta, st);
  return rc;
}

int fe_launch(odin_fe* fe, const void* d_pcm, int pcm_dtype, int n_utt, int64_t total_frames, int64_t n_tiles,
              int64_t n_tiles2, float* d_mspec, float* d_feat, float* d_energy, float* d_c0, uint8_t* d_sad,
              double* d_sad_thr, float* d_spec, int spec_log, cudaStream_t st, float2* d_cspec) {
  const odin_fe_config& c = fe->cfg;
  if (total_frames <= 0) return ODIN_OK;
  if (fe->ev[0] == nullptr) {
    for (int i = 0; i < 7; ++i) ODIN_CUDA_CHECK(cudaEventCreate(&fe->ev[i]));
    ODIN_CUDA_CHECK(cudaStreamCreateWithFlags(&fe->aux, cudaStreamNonBlocking));
  }
  ODIN_CUDA_CHECK(cudaEventRecord(fe->ev[0], st));
  // 1. DC sums
  if (c.remove_dc) {
    ODIN_CUDA_CHECK(cudaMemsetAsync(fe->d_dcsum, 0, sizeof(double) * n_utt, st));
    int64_t maxlen = 0;
    for (int u = 0; u < n_utt; ++u) maxlen = std::max(maxlen, fe->h_stage[u + 1] - fe->h_stage[u]);
    int max_chunks = (int)ceil_div<int64_t>(maxlen, DC_CHUNK);
    int64_t grid = (int64_t)n_utt * max_chunks;
    if (grid > 0x7fffffff) return set_error(ODIN_EINVAL, "batch too large for fe_dc_kernel");
    if (pcm_dtype == 0)
      fe_dc_kernel<int16_t><<<(unsigned)grid, 256, 0, st>>>((const int16_t*)d_pcm, fe->d_sample_off, n_utt,
                                                            max_chunks, fe->d_dcsum);
    else
      fe_dc_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const float*)d_pcm, fe->d_sample_off, n_utt,
                                                          max_chunks, fe->d_dcsum);
    ODIN_LAUNCH_CHECK("fe_dc_kernel");
  }
  ODIN_CUDA_CHECK(cudaEventRecord(fe->ev[1], st));
  // 2. frame kernel
  {
    // 0x80808080 decodes (ordered_to_float) to about -3.4e38: below any log-mel value
    ODIN_CUDA_CHECK(cudaMemsetAsync(fe->d_umax, 0x80, sizeof(int) * 2 * ((size_t)fe->cap_utt + 1), st));
    FrameArgs a{};
    a.pcm = d_pcm; a.sample_off = fe->d_sample_off; a.frame_off = fe->d_frame_off; a.tile_off = fe->d_tile_off;
    a.n_utt = n_utt; a.n_tiles = n_tiles; a.dcsum = fe->d_dcsum; a.L = fe->L; a.hop = fe->hop;
    a.remove_dc = c.remove_dc; a.preemph = c.preemph; a.win32 = fe->d_win32; a.win64 = fe->d_win64;
    a.tw = fe->d_tw; a.mel_start = fe->d_mel_start; a.mel_cnt = fe->d_mel_cnt; a.mel_off = fe->d_mel_off;
    a.mel_w = fe->d_mel_w; a.n_mels = fe->n_mels; a.scale2 = fe->scale2; a.mspec = d_mspec;
    a.energy = d_energy; a.umax = fe->d_umax;
    a.pad = fe->pad; a.spec = d_spec; a.spec_log = spec_log; a.umax_spec = fe->d_umax + fe->cap_utt + 1;
    a.cspec = d_cspec; a.scale1 = sqrtf(fe->scale2);
    a.mel_nnz = fe->mel_nnz;
    a.mel_tab = fe->d_mel_tab; a.mel_ps = fe->d_mel_ps; a.mel_trips = fe->mel_trips; a.mel_chunks = fe->mel_chunks;
    // four-step FFT kernel up to n_fft = 1024; the Stockham kernel keeps n_fft = 2048
    // (ODIN_FE_STOCKHAM=1 forces it everywhere, for A/B runs)
    const char* force = getenv("ODIN_FE_STOCKHAM");
    // (the chunk slots of the balanced filterbank sit behind the two power spectra in a pair region)
    const bool slots_fit = fe->N + 2 + 2 * fe->mel_chunks <= 2 * 32 * (fe->N / 32 + 1);
    const bool four = fe->N <= 1024 && slots_fit && !(force && force[0] == '1');
    // packed-f32x2 kernel (fe_frame5.cu) whenever the filterbank has the plain triangular structure and the
    // power spectrum itself is not an output (ODIN_FE_FRAME4=1 keeps the scalar four-step kernel, for A/B runs)
    const char* keep4 = getenv("ODIN_FE_FRAME4");
    const bool five = four && fe->mel5_ok && d_spec == nullptr && d_cspec == nullptr && !(keep4 && keep4[0] == '1');
    int rc;
    if (five) {
      a.tw = fe->d_tw4;
      a.win_c = fe->win_c; a.mel5_w = fe->d_mel5_w; a.mel5_flags = fe->d_mel5_flags;
      a.mel5_refs = fe->d_mel5_refs; a.mel5_nslots = fe->mel5_nslots; a.mel5_k = fe->mel5_k;
      a.n_samples = fe->h_stage[n_utt];
      if (fe->d_tile_ctr == nullptr) ODIN_CUDA_CHECK(cudaMalloc(&fe->d_tile_ctr, sizeof(int)));
      ODIN_CUDA_CHECK(cudaMemsetAsync(fe->d_tile_ctr, 0, sizeof(int), st));
      a.tile_ctr = fe->d_tile_ctr;
      rc = fe_frame5_launch(fe->N, pcm_dtype, a, st);
    } else if (four) {
      a.tw = fe->d_tw4;
      rc = (pcm_dtype == 0) ? dispatch_frame4<int16_t>(fe->N, a, st) : dispatch_frame4<float>(fe->N, a, st);
    } else {
      rc = (pcm_dtype == 0) ? dispatch_frame<float, int16_t>(fe->N, a, st)
                            : dispatch_frame<float, float>(fe->N, a, st);
    }
    if (rc) return rc;
  }
  if (d_spec != nullptr && spec_log && c.top_db >= 0.f) {
    const int64_t grid = std::min<int64_t>(n_tiles2, (int64_t)sm_count() * 8);
    fe_spec_clip_kernel<<<(unsigned)grid, 256, 0, st>>>(d_spec, fe->d_frame_off, fe->d_tile2_off, n_utt, n_tiles2,
                                                        fe->d_umax + fe->cap_utt + 1, c.top_db, fe->nbins);
    ODIN_LAUNCH_CHECK("fe_spec_clip_kernel");
  }
  ODIN_CUDA_CHECK(cudaEventRecord(fe->ev[2], st));
  // SADgmm reads only the frame energies, so it is forked onto the auxiliary stream and runs beside the
  // utterance pass (its tail is a handful of long utterances); SADthreshold needs c0 and stays in line.
  const bool has_vad = c.vad_kind != 0 && d_sad != nullptr;
  // (measured again with the current kernels on the config-3 shard: forked 2.59 ms per batch, in line 2.45 -- the
  // SADgmm CTAs hold every register of an SM, so the utterance pass cannot fill in beside them; off unless
  // ODIN_FE_VAD_FORK=1)
  const char* wantfork = getenv("ODIN_FE_VAD_FORK");
  const bool fork = has_vad && c.vad_kind == 1 && (wantfork && wantfork[0] == '1');
  fe->vad_forked = fork;
  auto launch_vad = [&](cudaStream_t vst) -> int {
    {
      VadArgs v{};
      v.frame_off = fe->d_frame_off; v.n_utt = n_utt; v.sad = d_sad; v.thr_out = d_sad_thr;
      v.order = fe->d_vad_order;
      v.nmix = c.vad_nmix; v.iters = c.vad_iters; v.smooth = c.vad_smooth; v.mode = (double)c.vad_mode;
      v.thr_energy = (double)c.thr_energy; v.thr_mean_scale = (double)c.thr_mean_scale;
      v.thr_proportion = (double)c.thr_proportion; v.thr_context = c.thr_context;
      // scratch: reuse the per-frame part of the handle
      if (fe->vad_scratch_cap < total_frames) {
        if (fe->d_vad_scratch) ODIN_CUDA_CHECK(cudaFree(fe->d_vad_scratch));
        fe->d_vad_scratch = nullptr; fe->vad_scratch_cap = 0;
        ODIN_CUDA_CHECK(cudaMalloc(&fe->d_vad_scratch, sizeof(float) * (total_frames + total_frames / 8 + 256)));
        fe->vad_scratch_cap = total_frames + total_frames / 8 + 256;
      }
      v.scratch = fe->d_vad_scratch;
      if (c.vad_kind == 1 && d_energy == nullptr) return set_error(ODIN_EINVAL, "SADgmm needs d_energy");
      if (c.vad_kind != 1 && d_c0 == nullptr) return set_error(ODIN_EINVAL, "SADthreshold needs d_c0");
      v.x = (c.vad_kind == 1) ? d_energy : d_c0;
      int rc_v = vad_dispatch(v, c.vad_kind, fe->h_stage + ((size_t)fe->cap_utt + 1), n_utt, vst);
      if (rc_v) return rc_v;
    }
    return ODIN_OK;
  };
  if (fork) {
    ODIN_CUDA_CHECK(cudaStreamWaitEvent(fe->aux, fe->ev[2], 0));
    ODIN_CUDA_CHECK(cudaEventRecord(fe->ev[5], fe->aux));
    int rc = launch_vad(fe->aux);
    if (rc) return rc;
    ODIN_CUDA_CHECK(cudaEventRecord(fe->ev[6], fe->aux));
  }
  // 3. utterance pass
  {
    PostArgs p{};
    p.frame_off = fe->d_frame_off; p.tile2_off = fe->d_tile2_off; p.n_utt = n_utt; p.umax = fe->d_umax;
    p.top_db = c.top_db; p.n_mels = fe->n_mels; p.n_c1 = fe->n_c1; p.n_ceps = c.n_ceps; p.W = c.delta_width;
    p.order = c.delta_order; p.dct = fe->d_dct; p.taps = fe->d_taps; p.mspec = d_mspec; p.write_mspec = 1;
    p.feat = d_feat; p.c0 = d_c0;
    const int h = c.delta_width / 2;
    const int HL = (c.delta_order >= 2) ? (c.delta_width - 1 + h) : (c.delta_order == 1 ? h : 0);
    const int HR = (c.delta_order >= 1) ? h : 0;
    const int NR = PT + HL + HR, ND = PT + ((c.delta_order >= 2) ? c.delta_width - 1 : 0);
    const bool post9 = c.delta_width == 9 && c.delta_order == 2 && c.n_ceps > 0 && (c.n_ceps & 3) == 0 && (fe->n_mels & 3) == 0 &&
                       d_feat != nullptr && ((reinterpret_cast<uintptr_t>(d_feat) | reinterpret_cast<uintptr_t>(d_mspec)) & 15) == 0 &&
                       getenv("ODIN_FE_POST_GENERIC") == nullptr;
    // sd1 ([ND][n_ceps]) and sd2 ([PT][n_ceps]) alias the mel rows
    const size_t mel_or_d1 = std::max((size_t)NR * (fe->n_mels | 1), (size_t)(ND + PT) * c.n_ceps);
    size_t smem = sizeof(float) * (mel_or_d1 + (((size_t)NR * fe->n_c1 + 3) & ~size_t(3)) + (size_t)((fe->n_c1 + 7) & ~7) * fe->n_mels +
                                   ((c.delta_width + 3) & ~3) + 4);
    if (post9) {
      const size_t CS = pad4odd(c.n_ceps), MS = pad4odd(fe->n_mels);
      smem = sizeof(float) * ((size_t)((fe->n_c1 + 7) & ~7) * fe->n_mels + (size_t)NR * CS + std::max((size_t)NR * MS, (size_t)(ND + PT) * CS));
    }
    if (smem > 227 * 1024) return set_error(ODIN_EINVAL, "post kernel needs %zu B smem", smem);
    auto kpost = post9 ? fe_post9_kernel : fe_post_kernel;
    ODIN_CUDA_CHECK(cudaFuncSetAttribute(kpost, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    p.n_tiles = n_tiles2;
    const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(3, (227 * 1024) / (smem + 1024)));
    const int64_t pgrid = std::min<int64_t>(n_tiles2, (int64_t)sm_count() * per_sm);
    kpost<<<(unsigned)pgrid, 256, smem, st>>>(p);
    ODIN_LAUNCH_CHECK("fe_post_kernel");
  }
  ODIN_CUDA_CHECK(cudaEventRecord(fe->ev[3], st));
  if (fork) {
    ODIN_CUDA_CHECK(cudaStreamWaitEvent(st, fe->ev[6], 0));
  } else if (has_vad) {
    int rc = launch_vad(st);
    if (rc) return rc;
  }
  ODIN_CUDA_CHECK(cudaEventRecord(fe->ev[4], st));
  fe->ev_valid = true;
  return ODIN_OK;
}

// ---------------------------------------------------------------------------
// Framing (speech.py:569-620): windowed frames [T, L] as float32 and their log energy (get_energy on the
// windowed frame, signal.py:1421-1440); same tiles and staging as the frame kernels
// ---------------------------------------------------------------------------
template <typename PCM>
__global__ void __launch_bounds__(FE_THREADS) fe_frames_kernel(FrameArgs a, float* __restrict__ frames) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* win64 = reinterpret_cast<double*>(smem_raw);
  float* sbuf = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(win64 + a.L) + 15) & ~uintptr_t(15));
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int L = a.L, hop = a.hop;
  for (int i = tid; i < L; i += FE_THREADS) win64[i] = a.win64[i];
  const PCM* __restrict__ pcm = reinterpret_cast<const PCM*>(a.pcm);
  for (int64_t tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int u = find_segment(a.tile_off, a.n_utt, tile);
    const int64_t s0 = a.sample_off[u];
    const int64_t n_u = a.sample_off[u + 1] - s0;
    const int64_t fbase = a.frame_off[u];
    const int T_u = (int)(a.frame_off[u + 1] - fbase);
    const int t0 = (int)(tile - a.tile_off[u]) * FT;
    const int nf = min(FT, T_u - t0);
    float mean = 0.f;
    if (a.remove_dc) {
      double s = (sizeof(PCM) == 2) ? (double)reinterpret_cast<const long long*>(a.dcsum)[u] : a.dcsum[u];
      mean = (float)(s / (double)n_u);
    }
    __syncthreads();
    const float* stile = stage_pcm<PCM>(sbuf, pcm + s0, n_u, (int64_t)t0 * hop, (nf - 1) * hop + L, mean, a.preemph, a.pad,
                                        tid, FE_THREADS);
    __syncthreads();
    for (int f = warp; f < nf; f += FE_WARPS) {
      const float* sf = stile + f * hop;
      double e = 0.0;
      for (int i = lane; i < L; i += 32) {
        const double v = win64[i] * (double)sf[i];
        e = fma(v, v, e);
        if (frames != nullptr) frames[(fbase + t0 + f) * L + i] = (float)v;
      }
      e = warp_sum(e);
      if (lane == 0 && a.energy != nullptr) {
        if (e == 0.0) e = (double)FLT_EPSILON;
        a.energy[fbase + t0 + f] = (float)log(e);
      }
    }
  }
}

int fe_frames_launch(odin_fe* fe, const void* d_pcm, int pcm_dtype, int n_utt, int64_t total_frames, int64_t n_tiles,
                     float* d_frames, float* d_energy, cudaStream_t st) {
  const odin_fe_config& c = fe->cfg;
  if (total_frames <= 0) return ODIN_OK;
  if (c.remove_dc) {
    ODIN_CUDA_CHECK(cudaMemsetAsync(fe->d_dcsum, 0, sizeof(double) * n_utt, st));
    int64_t maxlen = 0;
    for (int u = 0; u < n_utt; ++u) maxlen = std::max(maxlen, fe->h_stage[u + 1] - fe->h_stage[u]);
    int max_chunks = (int)ceil_div<int64_t>(maxlen, DC_CHUNK);
    int64_t grid = (int64_t)n_utt * max_chunks;
    if (grid > 0x7fffffff) return set_error(ODIN_EINVAL, "batch too large for fe_dc_kernel");
    if (pcm_dtype == 0)
      fe_dc_kernel<int16_t><<<(unsigned)grid, 256, 0, st>>>((const int16_t*)d_pcm, fe->d_sample_off, n_utt, max_chunks, fe->d_dcsum);
    else
      fe_dc_kernel<float><<<(unsigned)grid, 256, 0, st>>>((const float*)d_pcm, fe->d_sample_off, n_utt, max_chunks, fe->d_dcsum);
    ODIN_LAUNCH_CHECK("fe_dc_kernel");
  }
  FrameArgs a{};
  a.pcm = d_pcm; a.sample_off = fe->d_sample_off; a.frame_off = fe->d_frame_off; a.tile_off = fe->d_tile_off;
  a.n_utt = n_utt; a.n_tiles = n_tiles; a.dcsum = fe->d_dcsum; a.L = fe->L; a.hop = fe->hop;
  a.remove_dc = c.remove_dc; a.preemph = c.preemph; a.win64 = fe->d_win64; a.energy = d_energy; a.pad = fe->pad;
  const size_t smem = sizeof(double) * fe->L + sizeof(float) * ((size_t)(FT - 1) * fe->hop + fe->L + 16 + 4) + 16;
  if (smem > 227 * 1024) return set_error(ODIN_EINVAL, "framing kernel needs %zu B smem (hop too large)", smem);
  const int64_t grid = std::min<int64_t>(n_tiles, (int64_t)sm_count() * 4);
  if (pcm_dtype == 0) {
    ODIN_CUDA_CHECK(cudaFuncSetAttribute(fe_frames_kernel<int16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fe_frames_kernel<int16_t><<<(unsigned)grid, FE_THREADS, smem, st>>>(a, d_frames);
  } else {
    ODIN_CUDA_CHECK(cudaFuncSetAttribute(fe_frames_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fe_frames_kernel<float><<<(unsigned)grid, FE_THREADS, smem, st>>>(a, d_frames);
  }
  ODIN_LAUNCH_CHECK("fe_frames_kernel");
  return ODIN_OK;
}

int fe_compact_launch(odin_fe* fe, const uint8_t* d_sad, int n_utt, const float* d_feat, int dim,
                      int keep_unvoiced, float* d_out, int64_t* d_out_offsets, cudaStream_t st) {
  int grid = (int)std::min<int64_t>(ceil_div(n_utt, 4), (int64_t)sm_count() * 8);
  fe_count_kernel<<<grid, 128, 0, st>>>(d_sad, fe->d_frame_off, n_utt, keep_unvoiced, fe->d_cnt);
  ODIN_LAUNCH_CHECK("fe_count_kernel");
  fe_scan_kernel<<<1, 32, 0, st>>>(fe->d_cnt, n_utt, d_out_offsets);
  ODIN_LAUNCH_CHECK("fe_scan_kernel");
  fe_compact_kernel<<<grid, 128, 0, st>>>(d_sad, fe->d_frame_off, fe->d_cnt, d_out_offsets, n_utt, d_feat, dim,
                                          d_out);
  ODIN_LAUNCH_CHECK("fe_compact_kernel");
  return ODIN_OK;
}

}  // namespace odin
