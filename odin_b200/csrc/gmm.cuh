// GMM-UBM handle shared by the fp32 CUDA-core kernels (gmm_kernels.cu) and the
// 3xTF32 tcgen05 kernels (gmm_tc.cu).
#pragma once
#include "common.cuh"

#define ODIN_GMM_EPS 1e-6  // gmm_tmat.py:27

struct odin_gmm {
  int D = 0;        // feature dimension
  int max_nmix = 0;
  int M = 0;        // current number of mixtures
  int device = 0;
  // current model, reference layout [D, M] row-major (fp32)
  float* d_mean = nullptr;
  float* d_var = nullptr;
  float* d_w = nullptr;
  // cached posterior constants (gmm_tmat.py:493-504), refreshed by set_params / mstep / mixup
  //   Wk  [2D, Mpad]: rows 0..D-1 = -0.5*precision, rows D..2D-1 = mean*precision
  //   cst [Mpad]    : -0.5*(C + D log 2pi); padding columns hold -1e30
  float* d_Wk = nullptr;
  float* d_cst = nullptr;
  int Mpad = 0;  // M rounded up to 128
  // tcgen05 operand images (gmm_tc.cu): hi and lo TF32 splits of the [Mpad, 128]
  // log2-domain weight rows; Whi plain (copied into TMEM), Whs (hi) / Wlo (lo)
  // pre-swizzled per 128-mixture chunk as K-major SWIZZLE_128B shared-memory tiles.
  float* d_Whi = nullptr;
  float* d_Whs = nullptr;
  float* d_Wlo = nullptr;
  // pass-1 workspace of the tcgen05 path: per-chunk partial (max, sum) per frame
  void* d_part = nullptr;
  int64_t part_cap = 0;  // float2 elements
  double* d_utt_acc = nullptr;   // per-utterance statistics of a group of utterances (gmm_utt_stats_h)
  int64_t utt_acc_cap = 0;
  // segmented route (gmm_utt_stats_hseg): utterances padded to whole 64-frame tiles
  float* d_segX = nullptr;       // [seg_cap, D]
  uint8_t* d_segmask = nullptr;  // [seg_cap]
  int* d_segtile = nullptr;      // [seg_cap / 64] utterance of every tile
  int64_t* d_segoff = nullptr;   // [2][seg_utt_cap] frame offsets | padded offsets
  int64_t seg_cap = 0, seg_utt_cap = 0;
  // 3xFP16 tcgen05 path (gmm_h.cu): column scales, scaled model images, per-sub-batch frame
  // operand images (A: rows = frames, T: rows = [x^2|x|1] transposed) and 14 - lse2 per frame
  void* d_hscale = nullptr;
  void* d_hWplain = nullptr;
  void* d_hWimg = nullptr;
  float* d_hdsc = nullptr;
  void* d_himgA = nullptr;
  void* d_himgT = nullptr;
  float* d_hcb = nullptr;
  int64_t h_cap = 0;       // frames the images hold
  int64_t h_cb_cap = 0;    // frames d_hcb holds
  int64_t last_frames = 0; // frames covered by the events of the most recent E-step
  // per-frame log-sum-exp workspace (grows on demand)
  float* d_lse = nullptr;
  int64_t lse_cap = 0;
  // scratch for M-step rollback
  float* d_prev = nullptr;  // [3][D*max_nmix] (mean, var) + w
  // pinned staging + device copy of frame offsets for utt_stats
  int64_t* h_off = nullptr;
  int64_t* d_off = nullptr;
  int64_t off_cap = 0;
  // events bracketing the two kernels of the most recent E-step (odin_gmm_last_estep_ms)
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
  bool ev_valid = false;
  int last_impl = 0;
};

namespace odin {

int gmm_refresh_constants(odin_gmm* g, cudaStream_t st);
int gmm_reserve_lse(odin_gmm* g, int64_t n);

// fp32 CUDA-core path
int gmm_lse_ffma(odin_gmm* g, const float* X, const uint8_t* sad, int64_t N, float* lse,
                 double* stats /*nullable: L, nframes accumulated*/, cudaStream_t st);
int gmm_stats_ffma(odin_gmm* g, const float* X, const uint8_t* sad, int64_t N, const float* lse,
                   int want_second, double* stats, cudaStream_t st);
int gmm_utt_stats_ffma(odin_gmm* g, const float* X, const uint8_t* sad, const int64_t* d_off,
                       int n_utt, const float* lse, float* Z, float* Fhat, cudaStream_t st);
int gmm_post_ffma(odin_gmm* g, const float* X, int64_t N, const float* lse, float* post, float* logp,
                  cudaStream_t st);

// 3xTF32 tcgen05 path (gmm_tc.cu); returns ODIN_EINVAL when the shape is unsupported
bool gmm_tc_supported(const odin_gmm* g);
int gmm_tc_refresh(odin_gmm* g, cudaStream_t st);
int gmm_lse_tc(odin_gmm* g, const float* X, const uint8_t* sad, int64_t N, float* lse, double* stats,
               cudaStream_t st);
int gmm_stats_tc(odin_gmm* g, const float* X, const uint8_t* sad, int64_t N, const float* lse,
                 int want_second, double* stats, cudaStream_t st);

// 3xFP16 tcgen05 path (gmm_h.cu): both passes, records ev[0..2] itself
bool gmm_h_supported(const odin_gmm* g);
int gmm_estep_h(odin_gmm* g, const float* X, const uint8_t* sad, int64_t N, int want_second, double* stats,
                cudaStream_t st);
void gmm_h_free(odin_gmm* g);
// prepared frames (operand images of a resident frame matrix, built once)
int gmm_frames_create(odin_gmm* g, const float* X, int64_t N, void** out, cudaStream_t st);
void gmm_frames_destroy(void* f);
int gmm_utt_stats_h(odin_gmm* g, const float* X, const uint8_t* sad, const int64_t* h_off, int n_utt, float* d_Z,
                    float* d_Fhat, cudaStream_t st);
int gmm_utt_stats_hseg(odin_gmm* g, const float* X, const uint8_t* sad, const int64_t* h_off, int n_utt, float* d_Z,
                       float* d_Fhat, cudaStream_t st);
int gmm_estep_frames(odin_gmm* g, const void* f, const uint8_t* sad, int want_second, double* stats,
                     cudaStream_t st);

int gmm_mstep_launch(odin_gmm* g, const double* stats, int allow_rollback, int* rolled_back, cudaStream_t st);
int gmm_mixup_launch(odin_gmm* g, int newM, cudaStream_t st);

inline int64_t stats_size(int D, int M) { return (int64_t)M * (2 * D + 1) + 2; }

}  // namespace odin
