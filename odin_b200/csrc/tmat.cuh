// Internal declarations for the total-variability (T-matrix / i-vector) path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int TMAT_MAX_TV = 1024;  // tv <= ~150: the tv x tv systems live in one shared-memory square (fp64); larger ones
                                   // in per-CTA slabs of global memory (unblocked, slow: DESIGN.md 4.3)
constexpr int TMAT_MAX_D = 256;

struct odin_tmat {
  int tv = 0, M = 0, D = 0, t2 = 0;
  int64_t MD = 0;
  double* d_Tm = nullptr;      // [tv, MD]
  double* d_Sigma = nullptr;   // [MD]
  double* d_TinvS = nullptr;   // [tv, MD]
  double* d_TinvSTt = nullptr; // [M, t2]
  double* d_U = nullptr;       // [tv, tv] minimum-divergence factor
  int* d_perm = nullptr;       // [tv]
  double* d_small = nullptr;   // [t2]: LU.sum(0) for the min-div step, later the squared row norms
  void* sweep_graph = nullptr; // cudaGraphExec_t: one Jacobi sweep (tv - 1 rounds), captured on first use
  int* d_flag = nullptr;       // 0 ok; 1 E-step system / 2 M-step system / 3 min-div matrix not positive definite
  // per-call scratch (capacity in files)
  int64_t cap_files = 0;
  double* d_L1 = nullptr;      // [cap, t2]  Z T_invS_Tt, overwritten by Exx
  double* d_B1 = nullptr;      // [cap, tv]
  double* d_Ex = nullptr;      // [cap, tv]
  double* d_llk = nullptr;     // [cap]
  double* d_ws = nullptr;      // split-K workspace of the skinny B1 product
  int64_t ws_cap = 0;
  double* d_gws = nullptr;     // per-CTA squares in global memory for sizes beyond shared memory
  size_t gws_cap = 0;
};

namespace odin {
int tmat_refresh(odin_tmat* t, cudaStream_t st);
int tmat_estep(odin_tmat* t, const double* d_Z, const double* d_F, int64_t n_files, double* d_acc, cudaStream_t st);
int tmat_mstep(odin_tmat* t, const double* d_acc, int min_div, int orthogonalize, int sweeps, cudaStream_t st);
int tmat_ivector(odin_tmat* t, const double* d_Z, const double* d_F, int64_t n_files, double* d_out, cudaStream_t st);
}  // namespace odin
