// fe_frame5_kernel: the frame kernel of the speech front-end on packed fp32 pairs (sm_100a FADD2 / FMUL2 / FFMA2).
//
// Same contract as fe_frame4_kernel (fe_kernels.cu; reference arithmetic: odin/preprocessing/signal.py:1442-1562
// framing / window / rfft / scale, :1421-1440 frame energy, :1623-1648 power spectrum, :1650-1691 mel projection,
// :636-680 power2db) and the same four-step decomposition N = G * 32 with two real frames packed as re / im of one
// complex transform, but written for the instruction-issue roofline that bounds it (ncu of fe_frame4_kernel:
// 1 451 warp-instructions per frame, 59 % of the issue slots, FMA pipe 27 %):
//
//   * a complex value lives in one 64-bit register pair and every butterfly is add.f32x2 / sub.f32x2 /
//     mul.f32x2 / fma.f32x2: a complex add is ONE issue slot instead of two, a complex multiply two instead of
//     four (ptxas folds the re <-> im swap and the alternating sign into the operand modifiers of FFMA2, and
//     the multiplication by -i of a radix-4 step into those of FADD2).  tools/f32x2_bench.cu: FFMA2 / FADD2 keep
//     the FP32 lanes as busy as scalar code (72.8 vs 67.4 TFLOP/s) at half the issue slots.
//   * conjugate-symmetric split in 4 packed instructions per bin: with s = Z[k] + Z[N-k], d = Z[k] - Z[N-k] the
//     two power spectra are (|A|^2, |B|^2) = s*s + swap(d)*swap(d); the 1/4 and the 1/sum(w)^2 are folded
//     into the fp32 window.
//   * mel projection on the power values while they are in registers.  A lane owns N/64 CONTIGUOUS bins; a
//     bin between the centres p_s and p_(s+1) ("segment" s) only feeds the falling side of filter s-1 and the
//     rising side of filter s, so a lane keeps two packed accumulators (frames A | B) -- the filter it is on
//     the falling side of and the one it is on the rising side of -- drops the first into a slot when its bins
//     cross a centre and hands the second over.  Filter m adds its <= K slots (K = 2..4).  This replaces the
//     table walk of fe_frame4_kernel (random shared-memory reads: 25 % of its instructions and all of its bank
//     conflicts).
//   * every warp runs on its own (no block barrier after the table fill): the raw int16 samples of its next pass arrive by
//     cp.async while it computes, are converted (DC removal, pre-emphasis) into its pair region, and tiles are drawn
//     from a global counter; one CTA of 20 warps per SM stores the tables once.
//   * the rows of step A that are zero padding (L <= n_fft / 2 and the tail of the last live row) are template
//     parameters: no predicated-off instructions are issued for them.
//
// Shared memory per warp is ONE region per frame pair, used four times: float32 samples of the pass -> exchange tile [k1][l] of the four-step
// FFT -> natural-order spectrum (one pad element per lane chunk: conflict-free for the contiguous reads) ->
// partial-sum slots of the mel projection (the spectrum is pulled into registers first).
#include <float.h>
#include <math.h>
#include <stdlib.h>

#include <type_traits>

#include "fe.cuh"
#include "fe_frame.cuh"

namespace odin {

// cos / -sin of 2 pi q / 32
__device__ __forceinline__ constexpr float c32(int q) {
  constexpr float t[9] = {1.0f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                          0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                          0.19509032201612826785f, 0.0f};
  return q <= 8 ? t[q] : -t[16 - q];
}
__device__ __forceinline__ constexpr float s32(int q) { return q <= 8 ? -c32(8 - q) : -c32(q - 8); }

// In-register forward DFT of size R (power of two <= 32) over packed complex values, natural order in / out.
// NZ < R declares v[NZ..R) to be zero: the even / odd halves inherit the zero tail and a sub-transform with a
// single live input is a broadcast.
template <int R, int NZ = R> struct Dft5 {
  static __device__ __forceinline__ void run(u64 (&v)[R]) {
    if constexpr (NZ <= 1) {
#pragma unroll
      for (int q = 1; q < R; ++q) v[q] = v[0];
    } else {
      u64 e[R / 2], o[R / 2];
#pragma unroll
      for (int q = 0; q < R / 2; ++q) { e[q] = v[2 * q]; o[q] = v[2 * q + 1]; }
      Dft5<R / 2, (NZ + 1) / 2>::run(e);
      Dft5<R / 2, NZ / 2>::run(o);
#pragma unroll
      for (int q = 0; q < R / 2; ++q) {
        u64 t;
        if (q == 0) t = o[q];
        else if (4 * q == R) t = pk2(hi32(o[q]), -lo32(o[q]));   // * (-i)
        else t = cmulw(o[q], c32(q * (32 / R)), s32(q * (32 / R)));
        v[q] = add2(e[q], t);
        v[q + R / 2] = sub2(e[q], t);
      }
    }
  }
};
template <int NZ> struct Dft5<1, NZ> {
  static __device__ __forceinline__ void run(u64 (&)[1]) {}
};

// float2 elements of one pair region: the exchange tile (32 rows of G + 1), then the natural-order spectrum with
// one pad element behind every chunk of N / 64 bins (+ the copy of X[0] that stands in for X[N]); even, and
// = 8 mod 16 at G = 8 so that the two pair regions of a half-warp fall on different banks.
template <int N> constexpr int f5_region() {
  constexpr int tile = 32 * (N / 32 + 1), nat = N + 64 + 1;
  constexpr int r = (tile > nat ? tile : nat);
  return N == 256 ? ((r + 15) / 16) * 16 + 8 : (r + 1) & ~1;
}

// ---- per-warp staging of one PASS (the 2 * NP frames a warp transforms together) --------------------------------
// What a lane needs to know about a tile of FT frames (a tile never crosses an utterance).
struct F5Tile {
  int u;
  int64_t s0, n_u, fbase;
  int t0, nf;
  float mean;
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem)
               : "memory");
}

// int16 PCM: the raw samples of a pass -- one sample before its first frame (pre-emphasis) up to the end of its last
// frame -- are fetched by the warp with 16-byte asynchronous copies while it still computes the PREVIOUS pass;
// chunks that stick out of the PCM buffer are filled sample by sample.  raw[j] = buffer sample (s0 + g0 - 1 - mis + j),
// g0 = (t0 + f0) * hop - pad the utterance index of the pass' first sample, mis = 0..7 the 16-byte misalignment.
__device__ __forceinline__ int f5_raw_mis(const int16_t* pcm, const F5Tile& t, int f0, int hop, int pad) {
  return (int)((reinterpret_cast<uintptr_t>(pcm + t.s0 + ((int64_t)(t.t0 + f0) * hop - pad - 1)) >> 1) & 7);
}
__device__ __forceinline__ void f5_issue_raw(int16_t* raw, const int16_t* __restrict__ pcm, int64_t n_samples,
                                             const F5Tile& t, int f0, int fl, int L, int hop, int pad, int lane) {
  const int mis = f5_raw_mis(pcm, t, f0, hop, pad);
  const int64_t e0 = t.s0 + ((int64_t)(t.t0 + f0) * hop - pad - 1) - mis;   // buffer index of raw[0]: 16-byte aligned address
  const int nch = ((fl - 1) * hop + L + 1 + mis + 7) >> 3;
  for (int c = lane; c < nch; c += 32) {
    const int64_t ea = e0 + 8 * c;
    if (ea >= 0 && ea + 8 <= n_samples) {
      cp_async16(raw + 8 * c, pcm + ea);
    } else {
      for (int e = 0; e < 8; ++e) raw[8 * c + e] = (ea + e >= 0 && ea + e < n_samples) ? pcm[ea + e] : (int16_t)0;
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
// raw -> dst (float32, DC removed, pre-emphasised; speech.py:472-473, signal.py:955-967: two roundings per stage;
// zeros in the virtual padding).  Four outputs per lane and trip, one 16-byte store; a pass inside its utterance
// takes the path without bounds tests.
__device__ __forceinline__ void f5_convert_raw(float* __restrict__ dst, const int16_t* __restrict__ raw, int mis,
                                               const F5Tile& t, int f0, int fl, int L, int hop, int pad, float coef, int lane) {
  const int cnt = (fl - 1) * hop + L;
  const int64_t g0 = (int64_t)(t.t0 + f0) * hop - pad;   // utterance index of output 0
  const int16_t* __restrict__ r0 = raw + mis;             // r0[o] = sample g0 + o - 1
  const float mean = t.mean;
  const bool interior = g0 >= 1 && g0 + cnt <= t.n_u;
  for (int c = lane; 4 * c < cnt; c += 32) {
    float y[4];
    if (interior) {
      float x[5];
#pragma unroll
      for (int e = 0; e < 5; ++e) x[e] = __fsub_rn((float)r0[4 * c + e], mean);
#pragma unroll
      for (int e = 0; e < 4; ++e) y[e] = (coef != 0.f) ? __fsub_rn(x[e + 1], __fmul_rn(coef, x[e])) : x[e + 1];
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int64_t g = g0 + 4 * c + e;
        const bool in = g >= 0 && g < t.n_u;
        float cur = in ? __fsub_rn((float)r0[4 * c + e + 1], mean) : 0.f;
        if (in && coef != 0.f && g > 0) cur = __fsub_rn(cur, __fmul_rn(coef, __fsub_rn((float)r0[4 * c + e], mean)));
        y[e] = cur;
      }
    }
    *reinterpret_cast<float4*>(dst + 4 * c) = make_float4(y[0], y[1], y[2], y[3]);
  }
}
// float32 PCM: straight from global memory (no look-ahead; the rare input type)
__device__ __forceinline__ void f5_convert_f32(float* __restrict__ dst, const float* __restrict__ pu, const F5Tile& t, int f0,
                                               int fl, int L, int hop, int pad, float coef, int lane) {
  const int cnt = (fl - 1) * hop + L;
  const int64_t g0 = (int64_t)(t.t0 + f0) * hop - pad;
  const float mean = t.mean;
  for (int o = lane; o < cnt; o += 32) {
    const int64_t g = g0 + o;
    const bool in = g >= 0 && g < t.n_u;
    float cur = in ? __fsub_rn(pu[in ? g : 0], mean) : 0.f;
    if (in && coef != 0.f && g > 0) cur = __fsub_rn(cur, __fmul_rn(coef, __fsub_rn(pu[g - 1], mean)));
    dst[o] = cur;
  }
}

// NZ live rows in step A (rows r >= NZ are zero padding); EXACT: (NZ - 1) * G <= L, so only the last live row
// needs the i < L test.
//
// Every WARP runs on its own: it owns a contiguous range of 32-frame tiles, walks their passes (2 * NP frames
// each), and there is no block-level barrier after the table fill.  Per pass: the raw int16 samples were fetched
// into a per-warp buffer by cp.async during the previous pass; they are converted (DC, pre-emphasis) into the
// warp's pair region as float32, the next pass' samples are requested, and the region is then reused -- as before --
// as exchange tile, natural-order spectrum and partial-sum slots.  (The block-wide staging of the first version
// cost two barriers per tile and left all eight warps waiting for HBM together: 17 % of the stall samples.)
constexpr int F5_MAX_WARPS = 20;          // one CTA per SM: 5 warps per scheduler at 96 registers (a sixth would leave 80 and spill)
constexpr int F5_MAX_THREADS = 32 * F5_MAX_WARPS;

template <int N, typename PCM, int NZ, bool EXACT>
__global__ void __launch_bounds__(F5_MAX_THREADS, 1) fe_frame5_kernel(FrameArgs a) {
  constexpr int G = N / 32, NP = 32 / G, RS = G + 1, REG = f5_region<N>(), NK = N / 64;
  constexpr bool ASYNC = sizeof(PCM) == 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: win64 [L] | pair regions [warps][NP * REG] float2 | tw4 [N] float2 | mel weights [NK][32] float2 |
  //         refs [rounds][K/4][32] 4 x u16 | raw int16 [warps][rawlen] (int16 PCM)
  double* win64 = reinterpret_cast<double*>(smem_raw);
  u64* bufs = reinterpret_cast<u64*>(win64 + a.L + (a.L & 1));
  const int nthr = blockDim.x, nwarps = nthr >> 5;
  u64* tw4 = bufs + nwarps * NP * REG;
  float2* melw = reinterpret_cast<float2*>(tw4 + N);
  uint16_t* refs = reinterpret_cast<uint16_t*>(melw + NK * 32);
  const int rounds = (a.n_mels + 31) >> 5, K = a.mel5_k;
  const int L = a.L, hop = a.hop;
  const int rawlen = ((2 * NP - 1) * hop + L + 1 + 8 + 7) & ~7;   // int16 per buffer (multiple of 8: 16-byte slots)
  int16_t* raws = reinterpret_cast<int16_t*>((reinterpret_cast<uintptr_t>(refs + rounds * K * 32) + 15) & ~uintptr_t(15));

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane / G, l = lane % G;
  for (int i = tid; i < L; i += nthr) win64[i] = a.win64[i];
  for (int i = tid; i < N; i += nthr) tw4[i] = reinterpret_cast<const u64*>(a.tw)[i];
  for (int i = tid; i < NK * 32; i += nthr) melw[i] = a.mel5_w[i];
  for (int i = tid; i < rounds * K * 32; i += nthr) refs[i] = a.mel5_refs[i];
  const uint32_t mflags = a.mel5_flags[lane];   // bits 0..NK-2: a centre lies behind bin j; bits 16..: the lane's first slot
  const int zero_slot = a.mel5_nslots;
  const float win_c = a.win_c;
  u64* wbuf = bufs + warp * (NP * REG);
  float* stile = reinterpret_cast<float*>(wbuf);           // float32 samples of the pass (before the region is reused)
  int16_t* raw = raws + warp * rawlen;
  const PCM* __restrict__ pcm = reinterpret_cast<const PCM*>(a.pcm);
  const float coef = a.preemph;
  __syncthreads();   // tables are in place: from here on the warps never meet again

  // Tiles are handed out one at a time from a global counter (lane 0 draws, the warp follows): with ~9 tiles per warp
  // a static split leaves 8 % of the warps idle through the last tile.  A warp always holds the tile it works on and
  // the one after it, so that the next tile's first samples can be requested a pass ahead.
  auto draw = [&]() -> int64_t {
    int t = 0;
    if (lane == 0) t = atomicAdd(a.tile_ctr, 1);
    return (int64_t)__shfl_sync(0xffffffffu, t, 0);
  };
  auto tile_info = [&](int64_t tile) -> F5Tile {
    const int u = find_segment(a.tile_off, a.n_utt, tile);
    F5Tile t;
    t.u = u;
    t.s0 = a.sample_off[u];
    t.n_u = a.sample_off[u + 1] - t.s0;
    t.fbase = a.frame_off[u];
    t.t0 = (int)(tile - a.tile_off[u]) * FT;
    t.nf = min(FT, (int)(a.frame_off[u + 1] - t.fbase) - t.t0);
    t.mean = 0.f;
    if (a.remove_dc) {
      double s = (sizeof(PCM) == 2) ? (double)reinterpret_cast<const long long*>(a.dcsum)[u] : a.dcsum[u];
      t.mean = (float)(s / (double)t.n_u);
    }
    return t;
  };
  int64_t tile = draw();
  if (tile >= a.n_tiles) return;
  int64_t tile_next = draw();
  F5Tile cur = tile_info(tile);
  int f0 = 0;                                // first frame of the pass inside its tile
  double my_e = 0.0;                         // energy of frame `lane` of the current tile
  if constexpr (ASYNC)
    f5_issue_raw(raw, reinterpret_cast<const int16_t*>(pcm), a.n_samples, cur, 0, min(2 * NP, cur.nf), L, hop, a.pad, lane);
  while (true) {
    const int nf = cur.nf;
    const int fl = min(2 * NP, nf - f0);     // frames of this pass
    // the pass after this one: same tile, or the first of the next tile
    F5Tile nxt = cur;
    int nf0 = f0 + 2 * NP;
    bool has_next = true;
    if (nf0 >= nf) {
      nf0 = 0;
      if (tile_next < a.n_tiles) nxt = tile_info(tile_next); else has_next = false;
    }
    if constexpr (ASYNC) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();   // every lane's copies of this pass have landed; the previous pass is done with the region
      f5_convert_raw(stile, raw, f5_raw_mis(reinterpret_cast<const int16_t*>(pcm), cur, f0, hop, a.pad), cur, f0, fl, L, hop,
                     a.pad, coef, lane);
      __syncwarp();   // the float32 samples are in place and the raw buffer is free: it takes the next pass' samples
      if (has_next)
        f5_issue_raw(raw, reinterpret_cast<const int16_t*>(pcm), a.n_samples, nxt, nf0, min(2 * NP, nxt.nf - nf0), L, hop,
                     a.pad, lane);
    } else {
      __syncwarp();
      f5_convert_f32(stile, reinterpret_cast<const float*>(pcm) + cur.s0, cur, f0, fl, L, hop, a.pad, coef, lane);
      __syncwarp();
    }

    float wmax = -FLT_MAX;
    {
      u64* reg = wbuf + g * REG;
      // ---------------- step A: load + window + energy, 32-point DFT over r, twiddle
      {
        const int fA = f0 + 2 * g, fB = fA + 1;
        // frames past the end of the tile are computed on the pass' last frame and never stored
        const float* sA = stile + min(2 * g, fl - 1) * hop + l;
        const float* sB = stile + min(2 * g + 1, fl - 1) * hop + l;
        const double* wp = win64 + l;
        double eA = 0.0, eB = 0.0;
        u64 v[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          u64 z = 0ull;
          if (r < NZ) {
            const bool live = (EXACT && r < NZ - 1) ? true : (l + G * r < L);
            if (live) {
              // the window multiply in fp64 (signal.py:1543-1545) serves the frame energy; its rounding to fp32,
              // times 1/2 * 1/sum(w), is the FFT input -- one 8-byte table read per sample pair
              const double w = wp[G * r];
              const double wa = w * (double)sA[G * r], wb = w * (double)sB[G * r];
              eA = fma(wa, wa, eA);
              eB = fma(wb, wb, eB);
              z = mul2(pk2((float)wa, (float)wb), pk2(win_c, win_c));
            }
          }
          v[r] = z;
        }
        if (a.energy != nullptr) {
#pragma unroll
          for (int o = G / 2; o > 0; o >>= 1) {
            eA += __shfl_xor_sync(0xffffffffu, eA, o);
            eB += __shfl_xor_sync(0xffffffffu, eB, o);
          }
          // lane j of the warp keeps the energy of frame j of the tile until the tile's last pass (fp64 log there):
          // frame f0 + i of this pass sits in every lane of group i / 2 as eA (i even) or eB (i odd)
          const int i = lane - f0;
          const int src = ((i >> 1) * G) & 31;
          const double ea = __shfl_sync(0xffffffffu, eA, src), eb = __shfl_sync(0xffffffffu, eB, src);
          if (i >= 0 && i < 2 * NP) my_e = (i & 1) ? eb : ea;
        }
        Dft5<32, NZ>::run(v);
        // twiddles W_N^(k1 l), k1 = 8 a + b: ten table reads (b = 1..7, 8 a = 8, 16, 24) and 21 complex products instead
        // of 31 reads -- a packed complex multiply is two issue slots, a 64-bit shared-memory read two wavefronts of
        // the crossbar that bounds this kernel
        {
          u64 wa[4];
#pragma unroll
          for (int q = 1; q < 4; ++q) {
            wa[q] = tw4[8 * q * G + l];
            v[8 * q] = cmulw(v[8 * q], lo32(wa[q]), hi32(wa[q]));
          }
#pragma unroll
          for (int b = 1; b < 8; ++b) {
            const u64 wb = tw4[b * G + l];
            v[b] = cmulw(v[b], lo32(wb), hi32(wb));
#pragma unroll
            for (int q = 1; q < 4; ++q) {
              const u64 w = cmulw(wa[q], lo32(wb), hi32(wb));
              v[8 * q + b] = cmulw(v[8 * q + b], lo32(w), hi32(w));
            }
          }
        }
        __syncwarp();  // every lane has its samples in registers: the region becomes the exchange tile
#pragma unroll
        for (int k1 = 0; k1 < 32; ++k1) reg[k1 * RS + l] = v[k1];
      }
      __syncwarp();
      // ---------------- step B: G-point DFT over l for k1 = l + G t, natural-order store (chunk-padded)
      {
        u64 x[NP][G];
#pragma unroll
        for (int t = 0; t < NP; ++t) {
          const int k1 = l + G * t;
#pragma unroll
          for (int i = 0; i < G; ++i) x[t][i] = reg[k1 * RS + i];
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < NP; ++t) {
          Dft5<G>::run(x[t]);
          // k = k1 + 32 k2 -> k + k / NK; k1 = l + G t < 32, so k / NK = (32 / NK) k2 + k1 / NK
          u64* dst = reg + (l + G * t) + (l + G * t) / NK;
#pragma unroll
          for (int k2 = 0; k2 < G; ++k2) dst[32 * k2 + (32 / NK) * k2] = x[t][k2];
        }
        if (l == 0) reg[N + 64] = x[0][0];   // X[N] = X[0]: the mirror of the DC bin
      }
      __syncwarp();
      // ---------------- split + |.|^2 + mel + dB, the whole warp on one pair at a time
#pragma unroll 1
      for (int pp = 0; pp < NP; ++pp) {
        const int fA = f0 + 2 * pp, fB = fA + 1;
        if (fA >= nf) break;
        const bool hasB = fB < nf;
        u64* buf = wbuf + pp * REG;
        // lane owns bins k = NK lane + j; mirror N - k sits in chunk 63 - lane at offset NK - j (j >= 1), or at
        // the head of chunk 64 - lane (j = 0)
        const u64* p1 = buf + (NK + 1) * lane;
        const u64* p2 = buf + (NK + 1) * (63 - lane) + NK;
        u64 z1[NK], z2[NK];
#pragma unroll
        for (int j = 0; j < NK; ++j) { z1[j] = p1[j]; z2[j] = (j == 0) ? p2[1] : p2[-j]; }
        __syncwarp();   // the spectrum is in registers: the region now takes the partial-sum slots
        u64* slot = buf + (mflags >> 16);
        if (lane == 0) buf[zero_slot] = 0ull;
        u64 accF = 0ull, accR = 0ull;   // filter on whose falling / rising side the current bin lies: (frame A, frame B)
#pragma unroll
        for (int j = 0; j < NK; ++j) {
          const u64 s = add2(z1[j], z2[j]), d = sub2(z1[j], z2[j]);
          const u64 ds = pk2(hi32(d), lo32(d));
          const u64 P = fma2(ds, ds, mul2(s, s));   // (|A_k|^2, |B_k|^2) / sum(w)^2
          const float2 w = melw[j * 32 + lane];
          accF = fma2(P, pk2(w.x, w.x), accF);
          accR = fma2(P, pk2(w.y, w.y), accR);
          if (j < NK - 1) {
            if ((mflags >> j) & 1u) { *slot++ = accF; accF = accR; accR = 0ull; }
          } else {
            slot[0] = accF; slot[1] = accR;
          }
        }
        __syncwarp();
        float* rowA = a.mspec + (cur.fbase + cur.t0 + fA) * a.n_mels;
        for (int r = 0; r < rounds; ++r) {
          // the filter's slots in bin order, four 16-bit indices per word: four independent loads, then a tree
          const u64* rp = reinterpret_cast<const u64*>(refs) + r * (K >> 2) * 32 + lane;
          u64 acc = 0ull;
          for (int i = 0; i < (K >> 2); ++i) {
            const u64 ix = rp[i * 32];
            const u64 q0 = buf[ix & 0xffffu], q1 = buf[(ix >> 16) & 0xffffu];
            const u64 q2 = buf[(ix >> 32) & 0xffffu], q3 = buf[ix >> 48];
            acc = add2(acc, add2(add2(q0, q1), add2(q2, q3)));
          }
          const int m = 32 * r + lane;
          if (m < a.n_mels) {
            const float dA = db10<float>(lo32(acc));
            rowA[m] = dA;
            wmax = fmaxf(wmax, dA);
            if (hasB) {
              const float dB = db10<float>(hi32(acc));
              rowA[a.n_mels + m] = dB;
              wmax = fmaxf(wmax, dB);
            }
          }
        }
      }
      wmax = warp_max(wmax);
      if (lane == 0) atomicMax(a.umax + cur.u, float_to_ordered(wmax));
    }
    if (f0 + 2 * NP >= nf && a.energy != nullptr) {   // last pass of the tile: the logs of its frame energies
      if (lane < nf) {
        double e = my_e;
        if (e == 0.0) e = (double)FLT_EPSILON;  // signal.py:1436
        a.energy[cur.fbase + cur.t0 + lane] = (float)log(e);
      }
    }
    if (!has_next) break;
    if (nf0 == 0) { tile = tile_next; tile_next = draw(); }
    cur = nxt;
    f0 = nf0;
  }
}

template <int N, typename PCM, int NZ, bool EXACT>
static int launch5(const FrameArgs& a, cudaStream_t st) {
  constexpr int NP = 32 / (N / 32), NK = N / 64;
  const size_t rawlen = ((size_t)(2 * NP - 1) * a.hop + a.L + 1 + 8 + 7) & ~size_t(7);
  const size_t fixed = (size_t)(a.L + (a.L & 1)) * sizeof(double) + (size_t)N * sizeof(float2) + (size_t)NK * 32 * sizeof(float2) +
                       (size_t)((a.n_mels + 31) / 32) * a.mel5_k * 32 * sizeof(uint16_t) + 16;
  const size_t per_warp = (size_t)NP * f5_region<N>() * sizeof(float2) + (sizeof(PCM) == 2 ? rawlen * sizeof(int16_t) : 0);
  // the float32 samples of a pass are staged in the warp's pair region before it becomes the exchange tile
  if (((size_t)(2 * NP - 1) * a.hop + a.L + 4) * sizeof(float) > (size_t)NP * f5_region<N>() * sizeof(float2))
    return set_error(ODIN_EINVAL, "frame kernel: hop %d / frame %d too long for the pass staging", a.hop, a.L);
  // ONE CTA per SM with as many independent warps as its shared memory holds (the tables are stored once)
  const int warps = (int)std::min<size_t>(F5_MAX_WARPS, (227 * 1024 - fixed) / per_warp);
  if (warps < 1) return set_error(ODIN_EINVAL, "frame kernel: tables of %zu B leave no room for a warp", fixed);
  const size_t smem = fixed + warps * per_warp;
  auto k = fe_frame5_kernel<N, PCM, NZ, EXACT>;
  ODIN_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t grid = std::min<int64_t>(ceil_div<int64_t>(a.n_tiles, warps), (int64_t)sm_count());
  k<<<(unsigned)grid, 32 * warps, smem, st>>>(a);
  ODIN_LAUNCH_CHECK("fe_frame5_kernel");
  return ODIN_OK;
}

template <int N, typename PCM>
static int dispatch5(const FrameArgs& a, cudaStream_t st) {
  constexpr int G = N / 32;
  const int rows = (a.L + G - 1) / G;   // live rows of step A
  if constexpr (N == 1024) {
    if (rows <= 13 && 12 * G <= a.L) return launch5<N, PCM, 13, true>(a, st);
  } else {
    if (rows <= 25 && 24 * G <= a.L) return launch5<N, PCM, 25, true>(a, st);
  }
  if (rows <= 16) return launch5<N, PCM, 16, false>(a, st);
  return launch5<N, PCM, 32, false>(a, st);
}

int fe_frame5_launch(int N, int pcm_dtype, const FrameArgs& a, cudaStream_t st) {
  switch (N) {
    case 256: return pcm_dtype == 0 ? dispatch5<256, int16_t>(a, st) : dispatch5<256, float>(a, st);
    case 512: return pcm_dtype == 0 ? dispatch5<512, int16_t>(a, st) : dispatch5<512, float>(a, st);
    case 1024: return pcm_dtype == 0 ? dispatch5<1024, int16_t>(a, st) : dispatch5<1024, float>(a, st);
  }
  return set_error(ODIN_EINVAL, "n_fft %d unsupported by fe_frame5_kernel", N);
}

}  // namespace odin
