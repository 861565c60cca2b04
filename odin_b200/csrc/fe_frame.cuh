// Pieces shared by the frame kernels (fe_kernels.cu: fe_frame_kernel / fe_frame4_kernel / fe_frames_kernel,
// fe_frame5.cu: fe_frame5_kernel): launch arguments, the PCM staging copy, the tile -> utterance search, dB.
#pragma once
#include <float.h>
#include <math.h>

#include "fe.cuh"

namespace odin {

constexpr int FT = ODIN_FE_TILE;
constexpr int FE_WARPS = 8;
constexpr int FE_THREADS = FE_WARPS * 32;

struct FrameArgs {
  const void* pcm;
  const int64_t* sample_off;
  const int64_t* frame_off;
  const int64_t* tile_off;
  int n_utt;
  int64_t n_tiles;
  const double* dcsum;
  int L, hop, remove_dc;
  float preemph;
  const float* win32;
  const double* win64;
  const void* tw;  // C2<T>[N]: Stockham table exp(-2 pi i k / N), or the four-step table for fe_frame4_kernel
  int mel_nnz;
  const int* mel_start;
  const int* mel_cnt;
  const int* mel_off;
  const float* mel_w;
  const float2* mel_tab;   // lane-balanced filterbank (fe_frame4_kernel)
  const int* mel_ps;
  int mel_trips, mel_chunks;
  int n_mels;
  float scale2;
  float* mspec;   // [T, n_mels] unclipped dB
  float* energy;  // [T] nullable
  int* umax;      // [n_utt]
  int pad;        // zeros virtually prepended / appended to every utterance (stft(padding=True), signal.py:1529-1530)
  float* spec;    // [T, N/2+1] nullable: power spectrum (SpectraExtractor, signal.py:1718-1832), dB when spec_log
  int spec_log;
  int* umax_spec; // [n_utt] ordered-int max of the dB spectrum
  float2* cspec;  // [T, N/2+1] nullable: the complex STFT itself (signal.stft, signal.py:1442-1562), scaled by 1 / sum(w)
  float scale1;   // 1 / sum(w)
  // fe_frame5_kernel (fe_frame5.cu): 1/2 * 1/sum(w) (folded into the fp32 samples), lane-chunk form of the filterbank
  float win_c;
  const float2* mel5_w;      // [N/64][32] {falling-side weight, rising-side weight} of bin (N/64) * lane + j
  const uint32_t* mel5_flags;// [32] bit j: a mel centre lies between bins j and j + 1 of the lane's chunk | first slot << 16
  const uint16_t* mel5_refs; // [rounds][K][32] partial-sum slots of filter 32 round + lane (padded with the zero slot)
  int mel5_nslots, mel5_k;   // slots written per pair (the zero slot is index nslots); refs per filter
  int64_t n_samples;         // length of the whole PCM buffer (bounds of the 16-byte async copies)
  int* tile_ctr;             // zeroed per launch: the next tile to hand out (fe_frame5_kernel)
};

// PCM tile -> shared memory with DC removal and pre-emphasis fused (speech.py:472-473, signal.py:955-967);
// virtual sample v of the (padded) utterance is real sample v - pad, zeros outside (signal.py:1529-1530:
// the padding is applied to the processed signal, so padded samples are exact zeros).
//
// ncu put 31 % of the frame kernel's stall samples on the scalar 2-byte loads of the first version, so the
// tile is read as 16-byte vectors aligned in GLOBAL memory (8 int16 / 4 float samples per load) and written as
// 16-byte shared-memory stores: element i of the tile lands at sbase[i + mis], where mis (0..7) is the
// misalignment of the tile's first sample; the function returns sbase + mis, the tile origin for the readers.
// sbase must be 32-byte aligned and hold cnt + 16 floats.
template <typename PCM>
__device__ __forceinline__ float* stage_pcm(float* __restrict__ sbase, const PCM* __restrict__ pu, int64_t n_u,
                                            int64_t v0, int cnt, float mean, float coef, int pad, int tid, int nthr) {
  constexpr int V = 16 / (int)sizeof(PCM);
  const int64_t gstart = v0 - pad;   // utterance index of tile element 0 (negative inside the left padding)
  const int mis = (int)((reinterpret_cast<uintptr_t>(pu + gstart) & 15) / sizeof(PCM));
  const int n_chunks = (cnt + mis + V - 1) / V;
  for (int c = tid; c < n_chunks; c += nthr) {
    const int64_t g0 = gstart + (int64_t)c * V - mis;   // utterance index of the chunk's first sample
    float x[V], prev0 = 0.f;
    if (g0 >= 0 && g0 + V <= n_u) {
      union { uint4 u; PCM e[V]; } raw;
      raw.u = *reinterpret_cast<const uint4*>(pu + g0);
#pragma unroll
      for (int e = 0; e < V; ++e) x[e] = __fsub_rn((float)raw.e[e], mean);
      if (g0 > 0) prev0 = __fsub_rn((float)pu[g0 - 1], mean);
    } else {
#pragma unroll
      for (int e = 0; e < V; ++e) {
        const int64_t g = g0 + e;
        x[e] = (g >= 0 && g < n_u) ? __fsub_rn((float)pu[g], mean) : 0.f;
      }
      if (g0 > 0 && g0 - 1 < n_u) prev0 = __fsub_rn((float)pu[g0 - 1], mean);
    }
    float y[V];
#pragma unroll
    for (int e = 0; e < V; ++e) {
      const int64_t g = g0 + e;
      float cur = x[e];
      if (coef != 0.f && g > 0) cur = __fsub_rn(cur, __fmul_rn(coef, e == 0 ? prev0 : x[e - 1]));  // two roundings, like numpy (signal.py:965)
      y[e] = (g >= 0 && g < n_u) ? cur : 0.f;
    }
    float4* dst = reinterpret_cast<float4*>(sbase + c * V);
#pragma unroll
    for (int q = 0; q < V / 4; ++q) dst[q] = make_float4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
  }
  return sbase + mis;
}

__device__ __forceinline__ int find_segment(const int64_t* __restrict__ off, int n, int64_t v) {
  // largest u in [0, n) with off[u] <= v   (off is non-decreasing, off[0] = 0)
  int lo = 0, hi = n;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (off[mid] <= v) lo = mid; else hi = mid;
  }
  return lo;
}

template <typename T> __device__ __forceinline__ T db10(T v);
// 10 log10(max(v, 1e-10)) (signal.py:636-680) through the SFU: MUFU.LG2 is within 2 ulp of log2 (absolute error
// < 1e-5 dB over the range of a log-mel value, against a 1e-4 x 80 dB tolerance) and replaces ~20 instructions
// of log10f by two -- it was 4 % of the frame kernel's instructions.
template <> __device__ __forceinline__ float db10<float>(float v) {
  return 3.0102999566398120f * __log2f(fmaxf(v, 1e-10f));
}
template <> __device__ __forceinline__ double db10<double>(double v) { return 10.0 * log10(fmax(v, 1e-10)); }


// ---- packed fp32 pairs (sm_100a FADD2 / FMUL2 / FFMA2): one 64-bit register holds (lo, hi) ----
typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo32(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi32(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// (re, im) * (wr + i wi) = (re wr, im wr) + (im, re) * (-wi, wi)
__device__ __forceinline__ u64 cmulw(u64 a, float wr, float wi) {
  return fma2(pk2(hi32(a), lo32(a)), pk2(-wi, wi), mul2(a, pk2(wr, wr)));
}

// fe_frame5.cu: packed-f32x2 four-step kernel (n_fft <= 1024, conforming filterbank, no spectrum output)
int fe_frame5_launch(int N, int pcm_dtype, const FrameArgs& a, cudaStream_t st);

}  // namespace odin
