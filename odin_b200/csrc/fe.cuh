// Front-end handle: tables + per-run scratch.  Kernels live in fe_kernels.cu.
#pragma once
#include <vector>

#include "common.cuh"

constexpr int ODIN_FE_TILE = 32;       // frames per CTA tile in the frame kernel
constexpr int ODIN_FE_POST_TILE = 128; // frames per CTA tile in the utterance-pass kernel
constexpr int ODIN_FE_MAX_MELS = 128;

struct odin_fe {
  odin_fe_config cfg;
  int L = 0, hop = 0, N = 0, nbins = 0;
  int pad = 0;   // zeros on both sides of every utterance (cfg.padding ? L / 2 : 0)
  int n_mels = 0, n_c1 = 0 /* n_ceps+1 rows of the DCT */, feat_dim = 0;
  float scale2 = 0.f;  // (1/sum w)^2  (signal.py:1547,1557-1558 applied to |S|^2)
  // device tables
  float* d_win32 = nullptr;    // [L]
  double* d_win64 = nullptr;   // [L]
  float2* d_tw = nullptr;      // [N] exp(-2 pi i k / N)
  float2* d_tw4 = nullptr;     // [32][N/32] four-step twiddles exp(-2 pi i l k1 / N)
  int* d_mel_start = nullptr;  // [n_mels] first bin with non-zero weight
  int* d_mel_cnt = nullptr;    // [n_mels]
  int* d_mel_off = nullptr;    // [n_mels] offset into d_mel_w
  float* d_mel_w = nullptr;    // [nnz]
  float* d_dct = nullptr;      // [n_c1, n_mels]
  double* d_dct64 = nullptr;   // the same in fp64, uploaded on first use by odin_fe_ceps (sig_kernels.cu)
  float* d_taps = nullptr;     // [delta_width]
  int mel_nnz = 0;
  // lane-balanced form of the same filterbank for fe_frame4_kernel (see capi.cu: fe_build_tables)
  float2* d_mel_tab = nullptr; // [mel_trips][32] {weight, bin | flush << 15 | slot << 16}
  int* d_mel_ps = nullptr;     // [n_mels + 1] partial-sum slots of filter m: [ps[m], ps[m+1])
  int mel_trips = 0, mel_chunks = 0;
  // segment form of the filterbank + scaled window for fe_frame5_kernel (fe_frame5.cu); mel5_ok = the bank has
  // the triangular structure that kernel assumes (every bin feeds at most the two filters around it)
  float win_c = 0.f;               // 1/2 * 1/sum(w)
  float2* d_mel5_w = nullptr;      // [N/64][32]
  uint32_t* d_mel5_flags = nullptr;// [32]
  uint16_t* d_mel5_refs = nullptr; // [ceil(n_mels/32)][mel5_k][32]
  int mel5_nslots = 0, mel5_k = 0;
  bool mel5_ok = false;
  int* d_tile_ctr = nullptr;       // work counter of fe_frame5_kernel
  // per-run scratch (capacity in utterances)
  int cap_utt = 0;
  int64_t* h_stage = nullptr;  // pinned [5*(cap+1)]: sample_off, frame_off, tile_off, tile2_off, vad order
  int64_t* d_sample_off = nullptr;
  int64_t* d_frame_off = nullptr;
  int64_t* d_tile_off = nullptr;
  int64_t* d_tile2_off = nullptr;
  int64_t* d_vad_order = nullptr;  // [cap] utterance visiting order of the SADgmm kernel
  double* d_dcsum = nullptr;   // [cap] (int64 bit pattern for int16 input)
  int* d_umax = nullptr;       // [2 cap] ordered-int encoded utterance max of log-mel dB | of the dB spectrum
  int64_t* d_cnt = nullptr;    // [cap+1] compaction counts / offsets
  float* d_vad_scratch = nullptr;  // [frames] standardised energies
  int64_t vad_scratch_cap = 0;
  // events bracketing dc | frame | post | vad kernels of the most recent run (odin_fe_last_run_ms)
  cudaEvent_t ev[7] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool ev_valid = false;
  bool vad_forked = false;      // last run had the SADgmm kernel on the auxiliary stream (ev[5] -> ev[6])
  cudaStream_t aux = nullptr;   // SADgmm only needs the frame energies: it runs beside the utterance pass
  // host copies of the tables (tests, debugging)
  std::vector<double> h_win;
  std::vector<double> h_mel;   // dense [n_mels, nbins]
  std::vector<double> h_dct;
};

namespace odin {

int fe_build_tables(odin_fe* fe);           // host fp64 -> device
int fe_reserve(odin_fe* fe, int n_utt);
int fe_launch(odin_fe* fe, const void* d_pcm, int pcm_dtype, int n_utt, int64_t total_frames,
              int64_t n_tiles, int64_t n_tiles2, float* d_mspec, float* d_feat, float* d_energy, float* d_c0,
              uint8_t* d_sad, double* d_sad_thr, float* d_spec, int spec_log, cudaStream_t st, float2* d_cspec = nullptr);
int fe_frames_launch(odin_fe* fe, const void* d_pcm, int pcm_dtype, int n_utt, int64_t total_frames, int64_t n_tiles,
                     float* d_frames, float* d_energy, cudaStream_t st);
int fe_vad_standalone(int kind, const float* d_x, const int64_t* h_fo, int n_utt, int nmix, int iters, int smooth,
                      double mode, double thr_energy, double thr_mean_scale, double thr_proportion, int thr_context,
                      uint8_t* d_sad, double* d_thr, cudaStream_t st);
int fe_cmvn_launch(const float* d_x, float* d_y, int dim, const int64_t* d_frame_off, int n_utt,
                   const uint8_t* d_sad, int mean_var_norm, int var_norm, int windowed, int win_length,
                   cudaStream_t st);
int fe_compact_launch(odin_fe* fe, const uint8_t* d_sad, int n_utt, const float* d_feat, int dim,
                      int keep_unvoiced, float* d_out, int64_t* d_out_offsets, cudaStream_t st);

}  // namespace odin
