// __host__ __device__ restatements of the reference's small integer / float32
// logic, shared by the kernels and by the CPU-side tests (through the
// odin_host_* entry points).
#pragma once
#include <math.h>
#include <stdint.h>

#ifndef __CUDACC__
#define __host__
#define __device__
#endif

namespace odin {

// numpy's pairwise float32 summation (umath pairwise_sum, PW_BLOCKSIZE = 128):
// what np.mean / np.std run on the float32 log-energy in signal.py:305.
// Separate roundings are forced so the device version cannot contract.
__host__ __device__ inline float f32_add(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  volatile float r = a + b;
  return r;
#endif
}

__host__ __device__ inline float np_pairwise_sum_f32_impl(const float* a, int n) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = f32_add(res, a[i]);
    return res;
  } else if (n <= 128) {
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = a[j];
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] = f32_add(r[j], a[i + j]);
    float res = f32_add(f32_add(f32_add(r[0], r[1]), f32_add(r[2], r[3])),
                        f32_add(f32_add(r[4], r[5]), f32_add(r[6], r[7])));
    for (; i < n; ++i) res = f32_add(res, a[i]);
    return res;
  } else {
    int n2 = n / 2;
    n2 -= n2 % 8;
    return f32_add(np_pairwise_sum_f32_impl(a, n2), np_pairwise_sum_f32_impl(a + n2, n - n2));
  }
}

// Same, over a generated sequence f(i) (used for sum((e - mean)^2) without a
// temporary array).  F: int -> float.
template <class F>
__host__ __device__ inline float np_pairwise_sum_f32_gen(F f, int lo, int n) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = f32_add(res, f(lo + i));
    return res;
  } else if (n <= 128) {
    float r[8];
    for (int j = 0; j < 8; ++j) r[j] = f(lo + j);
    int i;
    for (i = 8; i < n - (n % 8); i += 8)
      for (int j = 0; j < 8; ++j) r[j] = f32_add(r[j], f(lo + i + j));
    float res = f32_add(f32_add(f32_add(r[0], r[1]), f32_add(r[2], r[3])),
                        f32_add(f32_add(r[4], r[5]), f32_add(r[6], r[7])));
    for (; i < n; ++i) res = f32_add(res, f(lo + i));
    return res;
  } else {
    int n2 = n / 2;
    n2 -= n2 % 8;
    return f32_add(np_pairwise_sum_f32_gen(f, lo, n2), np_pairwise_sum_f32_gen(f, lo + n2, n - n2));
  }
}

// np.mean / np.std of a float32 vector as numpy evaluates them
// (numpy/_core/_methods.py _mean/_var: float32 pairwise sums, the division by
// the count is done in float64 and cast back).
struct MeanStdF32 {
  float mean, std;
};
__host__ __device__ inline MeanStdF32 np_mean_std_f32(const float* e, int n) {
  MeanStdF32 r;
  float s = np_pairwise_sum_f32_impl(e, n);
  r.mean = (float)((double)s / (double)n);
  const float mean = r.mean;
  float ss = np_pairwise_sum_f32_gen(
      [e, mean](int i) {
        float d = f32_add(e[i], -mean);
#ifdef __CUDA_ARCH__
        return __fmul_rn(d, d);
#else
        volatile float p = d * d;
        return (float)p;
#endif
      },
      0, n);
  float var = (float)((double)ss / (double)n);
  r.std = sqrtf(var);
  return r;
}

#ifdef __CUDACC__
// Warp-cooperative, recursion-free twin of np_pairwise_sum_f32_gen (bit-identical
// result, returned to every lane).  The recursion of numpy's pairwise_sum is
// unrolled onto an explicit stack (depth <= log2(n/128)+2); inside a <=128
// element leaf lanes 0..7 carry the eight strided accumulators and the fixed
// combine tree ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) is an xor-butterfly (float
// addition is commutative, so the butterfly reproduces the tree exactly).
// Device recursion is avoided on purpose: its stack use is invisible to ptxas
// and overflowed the 1 KB default stack for utterances beyond ~15 s.
template <class F>
__device__ inline float warp_np_pairwise_sum_f32(F f, int n, int lane) {
  auto leaf = [&](int lo, int cnt) -> float {
    if (cnt < 8) {
      float res = 0.f;
      for (int i = 0; i < cnt; ++i) res = __fadd_rn(res, f(lo + i));
      return res;
    }
    const int body = cnt - (cnt % 8);
    float r = 0.f;
    if (lane < 8) {
      r = f(lo + lane);
      for (int i = 8; i < body; i += 8) r = __fadd_rn(r, f(lo + i + lane));
    }
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 1));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 2));
    r = __fadd_rn(r, __shfl_xor_sync(0xffffffffu, r, 4));
    float res = __shfl_sync(0xffffffffu, r, 0);
    for (int i = body; i < cnt; ++i) res = __fadd_rn(res, f(lo + i));
    return res;
  };
  constexpr int kDepth = 32;
  int s_lo[kDepth], s_n[kDepth], s_stage[kDepth];
  float s_left[kDepth];
  int sp = 0;
  s_lo[0] = 0; s_n[0] = n; s_stage[0] = 0; s_left[0] = 0.f;
  sp = 1;
  float ret = 0.f;
  while (sp > 0) {
    const int t = sp - 1;
    const int lo = s_lo[t], cnt = s_n[t];
    if (cnt <= 128) {
      ret = leaf(lo, cnt);
      --sp;
      continue;
    }
    int n2 = cnt / 2;
    n2 -= n2 % 8;
    if (s_stage[t] == 0) {
      s_stage[t] = 1;
      s_lo[sp] = lo; s_n[sp] = n2; s_stage[sp] = 0; ++sp;
    } else if (s_stage[t] == 1) {
      s_left[t] = ret;
      s_stage[t] = 2;
      s_lo[sp] = lo + n2; s_n[sp] = cnt - n2; s_stage[sp] = 0; ++sp;
    } else {
      ret = __fadd_rn(s_left[t], ret);
      --sp;
    }
  }
  return ret;
}

// np.mean / np.std (float32) of e[0..n), warp-cooperative; same value on every lane.
__device__ inline MeanStdF32 warp_np_mean_std_f32(const float* e, int n, int lane) {
  MeanStdF32 r;
  const float s = warp_np_pairwise_sum_f32([e](int i) { return e[i]; }, n, lane);
  r.mean = (float)((double)s / (double)n);
  const float mean = r.mean;
  const float ss = warp_np_pairwise_sum_f32(
      [e, mean](int i) {
        const float d = __fadd_rn(e[i], -mean);
        return __fmul_rn(d, d);
      },
      n, lane);
  r.std = sqrtf((float)((double)ss / (double)n));
  return r;
}
#endif  // __CUDACC__

// signal.py:969-1000 smooth(x, win, 'flat') followed by `>= 2/win`, evaluated
// for output element t.  x(i) -> 0/1 for i in [0, n).  wrap_u8 selects the
// uint8 route of SADthreshold (speech.py:1426-1431) where 2*x[0]-x[k] wraps to
// 255; otherwise the bool->int64 route of SADgmm (speech.py:1465-1473).
// np.convolve(w/w.sum(), s, 'same') accumulates s[i]*(1/win) sequentially in
// float64 with increasing i; that order is kept so the `>=` decision is
// identical.  Requires n >= win (the reference's slice arithmetic breaks below).
template <class F>
__host__ __device__ inline bool smooth_flat_ge_f(F x, int n, int win, bool wrap_u8, int t) {
  const double k = 1.0 / (double)win;
  const int x0 = x(0), xl = x(n - 1);
  double acc = 0.0;
  const int base = t + win - win / 2;
  for (int j = 0; j < win; ++j) {
    int i = base + j;  // index into s = [2x0 - x[win-1::-1], x, 2x[-1] - x[-1:-win:-1]]
    int v;
    if (i < win) {
      v = 2 * x0 - x(win - 1 - i);
      if (wrap_u8) v &= 0xff;
    } else if (i < win + n) {
      v = x(i - win);
    } else {
      v = 2 * xl - x(n - 1 - (i - win - n));
      if (wrap_u8) v &= 0xff;
    }
    double p = (double)v * k;
#ifdef __CUDA_ARCH__
    acc = __dadd_rn(acc, p);
#else
    volatile double tmp = acc + p;
    acc = tmp;
#endif
  }
  return acc >= 2.0 / (double)win;
}

}  // namespace odin
