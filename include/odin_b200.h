/* odin_b200.h -- C-ABI of libodin_b200.so: the B200 (sm_100a) replacement for the
 * arithmetic of odin-ai's speech front-end and GMM-UBM Baum-Welch hot path.
 *
 * The reference (trungnt13/odin-ai) is pure Python and has no FFI layer of its
 * own; each entry point below names the reference function(s) whose arithmetic
 * it replaces (file:line relative to the reference root).  The Python classes in
 * odin_b200/preprocessing and odin_b200/ml keep the reference's Extractor /
 * GMM surface and call these functions through ctypes (INTEGRATION.md shows the
 * stub a reference maintainer would add).
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / CUDA-runtime types in signatures
 *    (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream).
 *  - pointers prefixed d_ are DEVICE pointers owned by the caller, h_ are HOST
 *    pointers.  The library allocates only handle-internal tables / workspace.
 *  - every function returns 0 on success or a negative ODIN_E* code and never
 *    throws; odin_last_error() returns the message of the calling thread's last
 *    failure.
 *  - all kernels are launched on the caller's stream; functions are asynchronous
 *    unless stated.  Handles are not thread-safe: one per GPU / stream.
 *  - there is NO CPU fallback: without a CUDA device every compute entry point
 *    fails with ODIN_ENODEVICE.
 */
#ifndef ODIN_B200_H_
#define ODIN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODIN_OK 0
#define ODIN_EINVAL (-1)    /* bad argument / unsupported configuration */
#define ODIN_ENODEVICE (-2) /* no usable CUDA device */
#define ODIN_ECUDA (-3)     /* CUDA runtime error (see odin_last_error) */
#define ODIN_ENOMEM (-4)
#define ODIN_ESHORT (-5)    /* an utterance is shorter than one frame (signal.py:1532-1538 raises) */
#define ODIN_ENUMERIC (-6)  /* a linear system was not positive definite (T-matrix path) */

const char* odin_last_error(void);
int odin_version(void); /* 1000*major + minor */

/* ------------------------------------------------------------------------- */
/* Front-end: AudioReader DC removal -> PreEmphasis -> STFT(+energy) ->        */
/* PowerSpec -> MelsSpec(+power2db) -> MFCCs -> Delta -> SADgmm / SADthreshold */
/* ------------------------------------------------------------------------- */

typedef struct odin_fe odin_fe_t;

typedef struct {
  int32_t sr;
  int32_t frame_len;      /* samples; speech.py:207-220 resolves seconds -> samples */
  int32_t hop;            /* samples */
  int32_t n_fft;          /* power of two, 256..2048, >= frame_len (signal.py:1523-1524) */
  int32_t window;         /* 0 = hann, 1 = hamming (periodic; signal.py:812-830) */
  int32_t remove_dc;      /* speech.py:472-473 */
  float preemph;          /* 0 disables; signal.py:955-967 */
  int32_t n_mels;         /* <= 128 */
  float fmin, fmax;       /* Hz, already int-truncated as signal.py:1672-1681 does */
  float top_db;           /* < 0 disables the clip; signal.py:676-679 */
  int32_t n_ceps;         /* cepstra kept AFTER dropping c0 (speech.py:821-831); 0 = no MFCC */
  int32_t delta_width;    /* odd >= 3; signal.py:1002-1066 */
  int32_t delta_order;    /* 0, 1 or 2: output = [mfcc, d1, d2][: order+1] concatenated */
  int32_t vad_kind;       /* 0 none, 1 SADgmm on stft energy, 2 SADthreshold on c0 */
  int32_t vad_nmix;       /* SADgmm nb_mixture (speech.py:1447), 2..4 */
  int32_t vad_iters;      /* SADgmm nb_train_it */
  int32_t vad_smooth;     /* smooth_window; < 3 disables (signal.py:986-987) */
  float vad_mode;         /* signal.py:275-278, 2.0 = standard */
  float thr_energy;       /* SADthreshold energy_threshold (speech.py:1392-1399) */
  float thr_mean_scale;
  float thr_proportion;
  int32_t thr_context;
  int32_t padding;        /* stft(padding=True): frame_len/2 zeros on both sides (signal.py:1529-1530) */
} odin_fe_config;

/* Builds the window / twiddle / sparse-mel / DCT tables (computed in fp64 on the
 * host exactly as signal.py:682-810 does, stored fp32 + fp64 window) and uploads
 * them.  Synchronous. */
int odin_fe_create(const odin_fe_config* cfg, odin_fe_t** out);
void odin_fe_destroy(odin_fe_t* fe);

/* Output row width of `feat`: n_ceps * (1 + delta_order). */
int odin_fe_feat_dim(const odin_fe_t* fe);

/* HOST, integer-exact: frame_offsets[u+1]-frame_offsets[u] = 1+(n_u-L)//hop
 * (signal.py:1532-1538; n_u counts the padding when cfg.padding is set).  Returns ODIN_ESHORT if any utterance has n_u < L
 * (offsets are still written, with 0 frames for that utterance). */
int odin_fe_frame_offsets(const odin_fe_t* fe, const int64_t* h_sample_offsets, int32_t n_utt,
                          int64_t* h_frame_offsets);

/* HOST mirrors of the device integer logic, for CPU-side tests.
 * odin_host_smooth: signal.py:969-1000 with window='flat' followed by
 * `>= 2/win`; wrap_u8 = 1 reproduces the uint8 route of SADthreshold
 * (speech.py:1426-1431), 0 the bool route of SADgmm (speech.py:1465-1473). */
int odin_host_smooth(const uint8_t* x, int32_t n, int32_t win, int32_t wrap_u8, uint8_t* out);
/* np.mean / np.std of a float32 vector exactly as numpy evaluates them (pairwise
 * float32 sums; signal.py:305), the standardisation the SADgmm kernel applies. */
int odin_host_mean_std_f32(const float* e, int32_t n, float* mean, float* std_);
/* Handle-free twin of odin_fe_frame_offsets. */
int odin_host_frame_offsets(int32_t frame_len, int32_t hop, const int64_t* h_sample_offsets, int32_t n_utt,
                            int64_t* h_frame_offsets);
/* Dense fp64 tables as built for the device: which = 0 window [frame_len],
 * 1 mel filterbank [n_mels, n_fft/2+1] (signal.py:735-810), 2 DCT [n_ceps+1, n_mels]
 * (signal.py:682-733).  Returns the number of doubles written or a negative code. */
int odin_fe_get_table(const odin_fe_t* fe, int32_t which, double* out, int64_t cap);

/* The fused front-end over a ragged batch of utterances.
 *   d_pcm            concatenated samples, int16 (pcm_dtype 0) or float32 (1); int16 is NOT
 *                    rescaled (speech.py:453)
 *   h_sample_offsets [n_utt+1] HOST, utterance u = samples [off[u], off[u+1])
 * Outputs (device, any may be NULL except where noted), T = total frames:
 *   d_mspec  [T, n_mels]  log-mel dB after the utterance-global top_db clip (signal.py:1650-1691)
 *   d_feat   [T, feat_dim] MFCC (+deltas) (signal.py:1693-1716, 1002-1066; base.py:470-481)
 *   d_energy [T]          log frame energy on the windowed frame (signal.py:1421-1440)
 *   d_c0     [T]          DCT row 0 ("mfcc_energy", speech.py:829-830)
 *   d_sad    [T] uint8    VAD mask; d_sad_thr [n_utt] double threshold (speech.py:1433-1477)
 * d_mspec is required scratch whenever n_ceps > 0 or any output depends on it.
 * Asynchronous on `stream` apart from a small pinned->device upload of the offsets. */
int odin_fe_run(odin_fe_t* fe, const void* d_pcm, int32_t pcm_dtype, const int64_t* h_sample_offsets,
                int32_t n_utt, float* d_mspec, float* d_feat, float* d_energy, float* d_c0,
                uint8_t* d_sad, double* d_sad_thr, void* stream);

/* odin_fe_run plus the power spectrum itself as an output -- the all-in-one SpectraExtractor
 * (speech.py:849-929 -> signal.spectra, signal.py:1718-1832):
 *   d_spec [T, n_fft/2+1]  |rfft|^2 / (sum w)^2; spec_log != 0 converts it with power2db and clips it at
 *                          (utterance max - top_db) like signal.py:636-680 (spectra() passes top_db = 80).
 * d_spec == NULL makes this call identical to odin_fe_run. */
int odin_fe_run_spectra(odin_fe_t* fe, const void* d_pcm, int32_t pcm_dtype, const int64_t* h_sample_offsets,
                        int32_t n_utt, float* d_mspec, float* d_feat, float* d_energy, float* d_c0,
                        uint8_t* d_sad, double* d_sad_thr, float* d_spec, int32_t spec_log, void* stream);

/* The chain one stage at a time (sig_kernels.cu) -- what `pp.signal.*` and the `_transform` of a single extractor call.
 *   odin_fe_stft      signal.stft (signal.py:1442-1562) / STFTExtractor (speech.py:655-745): d_stft [T, n_fft/2+1] complex64
 *                     (interleaved re, im), scaled by 1 / sum(window); d_energy [T] log frame energy or NULL.  DC removal,
 *                     pre-emphasis and padding follow the handle, like odin_fe_run (use a handle with preemph = 0,
 *                     remove_dc = 0 for the bare function).
 *   odin_sig_preemph  signal.pre_emphasis (signal.py:955-967) over 1-D signals back to back (h_offsets [n_seg+1], HOST,
 *                     starting at 0); rows2d != 0 selects the reference's 2-D form (first column s[:,0] * (1 - coeff)).
 *   odin_sig_power    signal.power_spectrogram (signal.py:1623-1648): |S| ** power; is_complex: d_s holds n (re, im) pairs.
 *   odin_fe_mels      signal.mels_spectrogram (signal.py:1650-1691) / MelsSpecExtractor (speech.py:766-802): d_spec
 *                     [T, n_fft/2+1] power spectrum -> d_mspec [T, n_mels] through the handle's filterbank; log_db != 0
 *                     applies power2db with the handle's top_db against the maximum of each utterance's matrix.
 *   odin_fe_ceps      signal.ceps_spectrogram (signal.py:1693-1716) / MFCCsExtractor (speech.py:805-831): rows
 *                     [first_row, first_row + n_rows) of the handle's DCT basis (n_ceps + 1 rows) applied to d_mspec.
 *   odin_sig_delta    signal.delta (signal.py:1002-1066) along time for every utterance of a ragged [T, dim] batch:
 *                     order 1 -> d_delta1, order 2 -> d_delta1 and d_delta2 (with the lfilter delay of SURVEY 8.1-Q1). */
int odin_fe_stft(odin_fe_t* fe, const void* d_pcm, int32_t pcm_dtype, const int64_t* h_sample_offsets, int32_t n_utt,
                 void* d_stft, float* d_energy, void* stream);
int odin_sig_preemph(const float* d_x, float* d_y, const int64_t* h_offsets, int32_t n_seg, float coeff, int32_t rows2d,
                     void* stream);
int odin_sig_power(const float* d_s, int32_t is_complex, int32_t power, float* d_out, int64_t n, void* stream);
int odin_fe_mels(odin_fe_t* fe, const float* d_spec, const int64_t* h_frame_offsets, int32_t n_utt, float* d_mspec,
                 int32_t log_db, void* stream);
int odin_fe_ceps(odin_fe_t* fe, const float* d_mspec, int64_t n_frames, int32_t first_row, int32_t n_rows, float* d_out,
                 void* stream);
int odin_sig_delta(const float* d_x, int32_t dim, const int64_t* h_frame_offsets, int32_t n_utt, int32_t width, int32_t order,
                   float* d_delta1, float* d_delta2, void* stream);

/* Framing (speech.py:569-620) [+ CalculateEnergy, speech.py:623-649]: d_frames [T, frame_len] = window * signal
 * (float32; the reference keeps float64), d_energy [T] = log sum (windowed frame)^2; either may be NULL.  DC removal,
 * pre-emphasis and padding follow the handle's configuration, like odin_fe_run. */
int odin_fe_frames(odin_fe_t* fe, const void* d_pcm, int32_t pcm_dtype, const int64_t* h_sample_offsets,
                   int32_t n_utt, float* d_frames, float* d_energy, void* stream);

/* Feature-matrix extractors over a ragged batch ([T, dim] float32, h_frame_offsets [n_utt+1] HOST):
 *   odin_feat_stack      StackFeatures (base.py:724-771 = signal.stack_frames(keep_length=True), signal.py:1225-1294):
 *                        d_y [T, (2 n_context + 1) dim], row t = rows t-c .. t+c, zeros outside the utterance
 *   odin_feat_rasta_sdc  RASTAfilter (speech.py:1483-1533): rasta != 0 applies signal.rastafilt (signal.py:926-953) along
 *                        time; sdc >= 1 appends signal.shifted_deltas(N = k = dim, d = sdc, P = 3) (signal.py:1068-1090):
 *                        d_y [T, dim] or [T, dim + dim*dim]
 *   odin_feat_energy     CalculateEnergy on explicit frames (signal.py:1421-1440): d_energy [n_frames] */
int odin_feat_stack(const float* d_x, float* d_y, int32_t dim, const int64_t* h_frame_offsets, int32_t n_utt,
                    int32_t n_context, void* stream);
int odin_feat_rasta_sdc(const float* d_x, float* d_y, int32_t dim, const int64_t* h_frame_offsets, int32_t n_utt,
                        int32_t rasta, int32_t sdc, void* stream);
int odin_feat_energy(const float* d_frames, float* d_energy, int64_t n_frames, int32_t frame_len, int32_t take_log,
                     void* stream);
/* AsType (base.py:616-665) on device buffers: n elements, dtype codes 0 float16, 1 float32, 2 float64 (one side must be
 * float32; narrowing rounds to nearest even like ndarray.astype).  Lets float16 feature stores (the recipes'
 * AsType('float16') tail, examples/fsdd_ivec.py:105) cross PCIe at their stored width in both directions. */
int odin_feat_convert(const void* d_src, int32_t src_dtype, void* d_dst, int32_t dst_dtype, int64_t n, void* stream);
/* signal.smooth(x, win, window='flat') of a 0/1 vector (signal.py:969-1000), float64 out; win >= 3, n >= win. */
int odin_feat_smooth(const uint8_t* d_x, double* d_y, int64_t n, int32_t win, void* stream);

/* The two SAD extractors on ANY per-frame energy feature of a ragged batch (d_energy [T] float32, h_frame_offsets
 * [n_utt+1] HOST), without a front-end handle -- the same kernels odin_fe_run uses on the STFT energy / c0:
 *   odin_vad_gmm        SADgmm._transform = signal.vad_energy + smooth (speech.py:1439-1477, signal.py:293-331, 969-1000);
 *                       mode = 2.0 is VAD_MODE_STANDARD (signal.py:275-278)
 *   odin_vad_threshold  SADthreshold._transform (speech.py:1299-1332, 1415-1436)
 * d_sad [T] uint8, d_threshold [n_utt] double (nullable). */
int odin_vad_gmm(const float* d_energy, const int64_t* h_frame_offsets, int32_t n_utt, int32_t nb_mixture,
                 int32_t nb_train_it, int32_t smooth_window, float mode, uint8_t* d_sad, double* d_threshold,
                 void* stream);
int odin_vad_threshold(const float* d_energy, const int64_t* h_frame_offsets, int32_t n_utt, float energy_threshold,
                       float energy_mean_scale, int32_t frame_context, float proportion_threshold,
                       int32_t smooth_window, uint8_t* d_sad, double* d_threshold, void* stream);

/* ApplyingSAD (speech.py:1732-1756): order-preserving row compaction of `d_feat`
 * [T, dim] by the mask.  d_out_offsets [n_utt+1] receives the compacted utterance
 * boundaries (d_out_offsets[n_utt] = number of rows written).  keep_unvoiced = 1
 * keeps every frame of an utterance whose mask is all zero. */
int odin_fe_compact(odin_fe_t* fe, const uint8_t* d_sad, const int64_t* h_frame_offsets, int32_t n_utt,
                    const float* d_feat, int32_t dim, int32_t keep_unvoiced, float* d_out,
                    int64_t* d_out_offsets, void* stream);

/* AcousticNorm (speech.py:1536-1610) over a ragged batch: mean_var_norm -> signal.mvn(varnorm =
 * var_norm) (signal.py:853-876), then windowed -> signal.wmvn(w = win_length, varnorm = False)
 * (signal.py:878-924), statistics restricted to frames with d_sad != 0 when d_sad is given
 * (an empty selection yields NaN, as numpy's mean of an empty slice does).
 * d_x, d_y: [T, dim] float32 (dim <= 256); h_frame_offsets [n_utt+1] HOST; d_y may alias d_x only
 * when windowed == 0.  The host copy of the offsets is consumed before the call returns. */
int odin_fe_cmvn(const float* d_x, float* d_y, int32_t dim, const int64_t* h_frame_offsets, int32_t n_utt,
                 const uint8_t* d_sad, int32_t mean_var_norm, int32_t var_norm, int32_t windowed,
                 int32_t win_length, void* stream);

/* ------------------------------------------------------------------------- */
/* GMM-UBM: odin/ml/gmm_tmat.py                                               */
/* ------------------------------------------------------------------------- */

typedef struct odin_gmm odin_gmm_t;

/* D = feature dimension, max_nmix = largest mixture count the handle will see. */
int odin_gmm_create(int32_t feat_dim, int32_t max_nmix, odin_gmm_t** out);
void odin_gmm_destroy(odin_gmm_t* g);

/* Number of doubles in a packed statistics buffer for `nmix` mixtures:
 * Z[M] | F[D,M] | S[D,M] | L (sum of per-frame log-likelihood) | nframes. */
int64_t odin_gmm_stats_size(const odin_gmm_t* g, int32_t nmix);

/* Uploads the model and refreshes the cached posterior constants
 * (gmm_tmat.py:493-504): precision = 1/(var+1e-6), mu*precision,
 * C = sum mu^2 prec + sum log(var+1e-6) - 2 log(w+1e-6).
 * d_mean, d_var: [D, M] row-major (the reference layout); d_w: [M]. */
int odin_gmm_set_params(odin_gmm_t* g, int32_t nmix, const float* d_mean, const float* d_var,
                        const float* d_w, void* stream);

/* E-step over N frames (gmm_tmat.py:1012-1041): ACCUMULATES into the packed
 * d_stats (caller zeroes it at the start of an EM iteration; several calls /
 * batches may accumulate into the same buffer, which replaces the host-side sum
 * of gmm_tmat.py:1148-1156,249-265).  d_sad (uint8 [N]) may be NULL; frames with
 * sad == 0 are skipped (gmm_tmat.py:162-164).  want_second = 0 skips S.
 * impl: 0 = auto, 1 = fp32 CUDA-core kernels, 2 = 3xTF32 tcgen05 kernels (M >= 96),
 * 3 = 3xFP16 tcgen05 kernels (M >= 256; D % 4 == 0 and D <= 60 for both tensor paths). */
int odin_gmm_estep(odin_gmm_t* g, const float* d_X, const uint8_t* d_sad, int64_t n_frames,
                   int32_t want_second, double* d_stats, int32_t impl, void* stream);

/* Prepared frames (3xFP16 tensor path only).  The kernel-ready operand images of a frame matrix
 * depend on the data alone, so for frames that stay resident across EM iterations
 * (gmm_tmat.py:1278-1306 visits the same X every iteration) they can be built once: 1 KB of HBM
 * per frame.  odin_gmm_estep_frames == odin_gmm_estep(impl 3) on the same frames, minus the image
 * build.  ODIN_ENOMEM if the images do not fit (fall back to odin_gmm_estep). */
typedef struct odin_gmm_frames odin_gmm_frames_t;
int odin_gmm_frames_create(odin_gmm_t* g, const float* d_X, int64_t n_frames, odin_gmm_frames_t** out,
                           void* stream);
void odin_gmm_frames_destroy(odin_gmm_frames_t* f);
int odin_gmm_estep_frames(odin_gmm_t* g, const odin_gmm_frames_t* f, const uint8_t* d_sad, int32_t want_second,
                          double* d_stats, void* stream);

/* The exchange step of SURVEY 8e for a binder that owns an NCCL communicator (replaces _ExpectationResults.update +
 * the multiprocessing queue, gmm_tmat.py:249-265, 1199-1220): in-place, in-stream ncclAllReduce(sum, double) of the packed
 * statistics of the current model size.  `nccl_comm` is an ncclComm_t; libnccl is resolved at run time (no link-time
 * dependency).  The Python binding performs the same all-reduce through torch.distributed (odin_b200/sharding.py). */
int odin_gmm_allreduce(odin_gmm_t* g, double* d_stats, void* nccl_comm, void* stream);

/* M-step (gmm_tmat.py:1233-1276) from packed stats, in fp64 on device; writes the
 * new model into the handle AND to d_mean/d_var/d_w (fp32, reference layout).
 * If any variance < 0: allow_rollback = 1 keeps the previous model, 0 clips at 0;
 * *d_rolled_back (int32 on device) is set to 1 in either case. */
int odin_gmm_mstep(odin_gmm_t* g, const double* d_stats, int32_t allow_rollback, float* d_mean,
                   float* d_var, float* d_w, int32_t* d_rolled_back, void* stream);

/* Mixture split (gmm_tmat.py:1308-1338): M -> min(2M, new_nmix); reads and
 * rewrites the caller's arrays, which must have room for new_nmix columns laid
 * out as [D, new_nmix] AFTER the call (input is [D, M] packed). */
int odin_gmm_mixup(odin_gmm_t* g, int32_t new_nmix, float* d_mean, float* d_var, float* d_w,
                   void* stream);

/* Per-utterance centred statistics (gmm_tmat.py:708-767, 769-913):
 *   d_Z    [n_utt, M]     zeroth order
 *   d_Fhat [n_utt, M*D]   (F - mean*Z) flattened column-major: index m*D + d
 * h_frame_offsets [n_utt+1] HOST. */
int odin_gmm_utt_stats(odin_gmm_t* g, const float* d_X, const uint8_t* d_sad,
                       const int64_t* h_frame_offsets, int32_t n_utt, float* d_Z, float* d_Fhat,
                       int32_t impl, void* stream);

/* Per-frame quantities (gmm_tmat.py:916-995): d_llk [N] = logsumexp_m logprob;
 * d_post [N, M] posteriors and d_logprob [N, M] component log-densities (either may be NULL). */
int odin_gmm_score(odin_gmm_t* g, const float* d_X, int64_t n_frames, float* d_llk, float* d_post,
                   float* d_logprob, void* stream);

/* Device time of the kernels of the most recent call, measured with CUDA events
 * recorded on the caller's stream around each kernel (blocks until they complete).
 * odin_gmm_last_estep_ms: log-sum-exp kernel(s) and statistics kernel of the last
 * odin_gmm_estep; *impl_used = 1 (fp32), 2 (tcgen05 3xTF32) or 3 (tcgen05 3xFP16).
 * odin_fe_last_run_ms: ms4 = {DC sums, frame kernel, utterance pass, VAD}. */
int odin_gmm_last_estep_ms(odin_gmm_t* g, float* lse_ms, float* stats_ms, int32_t* impl_used);
/* Frames covered by those events (impl 3 times the last sub-batch only; others the whole call). */
int64_t odin_gmm_last_estep_frames(const odin_gmm_t* g);
int odin_fe_last_run_ms(odin_fe_t* fe, float* ms4);

/* Number of kernels this library has launched since load (bench.py's gpu_launches). */
int64_t odin_launch_count(void);

/* ------------------------------------------------------------------------- */
/* Total-variability model / i-vectors: odin/ml/gmm_tmat.py class Tmatrix      */
/* (fp64 like the reference's default dtype, gmm_tmat.py:1412)    */
/* ------------------------------------------------------------------------- */
typedef struct odin_tmat odin_tmat_t;

/* tv_dim <= 1024, feat_dim <= 256.  Up to tv_dim ~ 150 the tv x tv systems are factorised in shared memory; larger
 * ones run from per-CTA slabs of global memory with the same (unblocked) algorithms, i.e. correct but slow. */
int odin_tmat_create(int32_t tv_dim, int32_t nmix, int32_t feat_dim, odin_tmat_t** out);
void odin_tmat_destroy(odin_tmat_t* t);

/* Number of doubles in the packed E-step statistics:
 *   LU [nmix, tv(tv+1)/2] | RU [tv, nmix*feat_dim] | llk | nframes        (gmm_tmat.py:1694-1725)
 * One buffer so that ranks holding different files can all-reduce it in one collective. */
int64_t odin_tmat_acc_size(const odin_tmat_t* t);

/* d_Tm [tv, nmix*feat_dim] (column m*feat_dim + d), d_Sigma [nmix*feat_dim] = GMM variances mixture-major
 * (gmm_tmat.py:1466-1468).  Rebuilds T_invS and T_invS_Tt (gmm_tmat.py:1578-1589).  d_Sigma may be NULL to
 * keep the current one. */
int odin_tmat_set_model(odin_tmat_t* t, const double* d_Tm, const double* d_Sigma, void* stream);
/* Copies out any of Tm [tv, MD], T_invS [tv, MD], T_invS_Tt [nmix, t2] (NULL = skip). */
int odin_tmat_get_model(odin_tmat_t* t, double* d_Tm, double* d_T_invS, double* d_T_invS_Tt, void* stream);

/* E-step over n_files utterances: d_Z [n, nmix], d_F [n, nmix*feat_dim] (centred first-order statistics
 * of odin_gmm_utt_stats, as doubles); ACCUMULATES into d_acc (zero it before the first call of an iteration). */
int odin_tmat_estep(odin_tmat_t* t, const double* d_Z, const double* d_F, int64_t n_files, double* d_acc,
                    void* stream);

/* M-step from the (all-reduced) statistics: per-mixture solve, minimum-divergence re-estimation and
 * orthogonalisation (gmm_tmat.py:1818-1865), then the cached statistics are rebuilt.  The orthogonalised
 * T equals the reference's diag(s) V^T up to the sign of each row (the SVD's sign convention).
 * Synchronises the stream; returns ODIN_ENUMERIC when a system was not positive definite. */
int odin_tmat_mstep(odin_tmat_t* t, const double* d_acc, int32_t min_div_est, int32_t orthogonalize,
                    void* stream);

/* i-vectors (gmm_tmat.py:1898-1942): d_out [n_files, tv]. */
int odin_tmat_ivector(odin_tmat_t* t, const double* d_Z, const double* d_F, int64_t n_files, double* d_out,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ODIN_B200_H_ */
