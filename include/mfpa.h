/*
 * mfpa.h — C ABI of the B200-native musicFPaugment query path (libmfpa.so).
 *
 * The reference (deezer/musicFPaugment) is pure Python and has no FFI layer;
 * its boundary for this path is the Python call surface listed in SURVEY.md
 * §8(b).  Each entry point below names the reference function(s) it replaces
 * (paths relative to the reference root).  The Python mirror in
 * musicfpaugment_b200/ binds these with ctypes.
 *
 * Conventions
 *   - every function returns 0 on success, a negative MFPA_E* code otherwise;
 *     mfpa_last_error() gives the message of the calling thread's last failure.
 *   - "_dev" pointers are device pointers on the context's GPU, caller-owned.
 *     "_host" pointers are ordinary host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *     Device entry points are asynchronous on that stream; *_host entry points
 *     synchronise before returning.
 *   - a context owns scratch buffers that grow on demand; one context must not
 *     be used from two threads at once.
 *   - geometry is the reference's fixed analysis setting
 *     (testing/parameters.py:17-26): n_fft 512, hop 256, 257 bins, 256 kept.
 *
 * "item" = one (query, shift) pair: item = query * shifts + shift.  A shifted
 * item analyses x[off:], off = int(shift / shifts * 256)
 * (afp/audfprint/peak_extractor.py:411-413).
 */
#ifndef MFPA_H
#define MFPA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MFPA_ABI_VERSION 1

#define MFPA_N_FFT 512
#define MFPA_HOP 256
#define MFPA_BINS 257      /* rfft bins */
#define MFPA_ROWS 256      /* bins kept by the peak picker (Nyquist dropped) */
#define MFPA_MAG_PITCH 264 /* elements per frame in the internal frame-major magnitude layout */
#define MFPA_MAX_PKS 5     /* hard upper bound on pks-per-frame */
#define MFPA_MAX_SHIFTS 8
#define MFPA_HASHES_PER_FRAME 15 /* MFPA_MAX_PKS * fanout(3): capacity per frame of a hash list */

#define MFPA_OK 0
#define MFPA_EINVAL -1 /* bad argument */
#define MFPA_ECUDA -2  /* CUDA runtime error */
#define MFPA_ENOMEM -3
#define MFPA_ECAP -4   /* an output capacity was too small */

typedef struct mfpa_ctx mfpa_ctx;

/* Analysis parameters of Audfprint_peaks.__init__ (peak_extractor.py:86-113)
 * and find_peaks (:295).  mfpa_afp_defaults() fills the reference's
 * afp_settings["audfprint"] values. */
typedef struct mfpa_afp_params {
  double a_dec;    /* 1 - 0.01*(density*sqrt(n_hop/352.8)/35), computed by the caller in float64 */
  double f_sd;     /* "freq-sd": Gaussian spreading width (30) */
  int32_t maxpks;  /* "pks-per-frame" (5), <= MFPA_MAX_PKS */
  int32_t mindt;   /* 2  */
  int32_t targetdt;/* 63 */
  int32_t targetdf;/* 31 */
  int32_t fanout;  /* maxpairsperpeak 3 */
  int32_t reserved;
} mfpa_afp_params;

/* ---- context ---------------------------------------------------------- */
int mfpa_abi_version(void);
const char* mfpa_last_error(void);
int mfpa_create(mfpa_ctx** out, int device);
void mfpa_destroy(mfpa_ctx* ctx);
void mfpa_afp_defaults(mfpa_afp_params* p);
/* Upload the 513-entry spreading table exp(-0.5*((k/f_sd)^2)), k=-256..256,
 * computed by the caller with numpy so it is bit-identical to the reference's
 * cached __sp_vals (peak_extractor.py:159-165).  Without this call the library
 * fills the table with the host C library's exp(). */
int mfpa_set_spread_table(mfpa_ctx* ctx, const double* table513_host);

/* Options.  MFPA_OPT_PEAKS_F64 (default 0): 1 makes mfpa_audfprint_peaks / mfpa_fingerprint* run the
 * float64 picker (the one behind mfpa_audfprint_peaks_from_spec) on the float32 magnitudes instead of
 * the float32 picker; slower, same hashes on every input tested. */
#define MFPA_OPT_PEAKS_F64 1
/* MFPA_OPT_MATCH_PACKED (default 0): 1 makes mfpa_match_counts write, and mfpa_match_select read, the
 * per-(query, track) raw counts as pairs of 16-bit counters: counts_dev is then uint32
 * [B][(n_tracks + 1) / 2], word j = count[2j] | count[2j+1] << 16.  Summing such rows as uint32 (the
 * sharded path's NCCL reduction) is exact as long as no total reaches 65536; half the bytes. */
#define MFPA_OPT_MATCH_PACKED 2
/* MFPA_OPT_MATCH_UNFUSED (default 0): mfpa_match normally runs counts + select + collect as one kernel
 * whose histogram never leaves shared memory (indexes of up to 110000 tracks); 1 forces the four-step
 * path (the one the sharded matcher drives with collectives in between) for comparison; 2 forces it with int32
 * counters in global memory (exact mfpa_match_counts for queries whose packed 16-bit counters overflowed, nrows -6). */
#define MFPA_OPT_MATCH_UNFUSED 3
/* MFPA_OPT_PART_BUDGET_MB (default 2048): device scratch, in MiB, that the partitioned long-filter path of
 * mfpa_augment may use for block spectra; queries with long filters are processed in groups that fit. */
#define MFPA_OPT_PART_BUDGET_MB 4
/* MFPA_OPT_CONV_OCC (default 3): resident blocks per SM the FFT-convolution kernels of mfpa_augment are compiled
 * for - 3 (80 registers per thread, 24 warps per SM) or 2 (up to 128 registers); a tuning knob, same results. */
#define MFPA_OPT_CONV_OCC 5
/* MFPA_OPT_STAGE_TIMES (default 0): 1 makes mfpa_augment_fingerprint and mfpa_fingerprint stamp a CUDA event at the
 * start of every stage on the caller's stream (setting the option also clears the record); mfpa_stage_times then
 * returns the mean duration of each stage over the last (up to 16) calls - how bench.py times the dominant kernel
 * live inside its timed region. */
#define MFPA_OPT_STAGE_TIMES 6
/* MFPA_OPT_CLIP_POOLED (default 0): 1 gives Clipping the reference's BATCH semantics in mfpa_augment*: torch.quantile
 * with a vector of q and no dim flattens its input, so every row that has MFPA_AUG_CLIP set is clipped to the
 * quantiles of ITS percentile over the samples of ALL such rows of the call (clipping.py:67-101; identical to the
 * per-row rule when one row is selected, the only case the reference's drivers produce).  At most 2^24 pooled
 * samples, torch.quantile's own limit.  The drop-in's batch_augment sets it. */
#define MFPA_OPT_CLIP_POOLED 7
int mfpa_set_option(mfpa_ctx* ctx, int option, int value);
#define MFPA_STAGE_HPF1_FILTER 0  /* filter_spectrum_kernel (loudspeaker high-pass) */
#define MFPA_STAGE_HPF1_CONV 1    /* fftconv_kernel */
#define MFPA_STAGE_IR_FILTER 2
#define MFPA_STAGE_IR_CONV 3
#define MFPA_STAGE_MIX 4          /* clip_sample + mix + clip_finish */
#define MFPA_STAGE_CLIP_LPF 5
#define MFPA_STAGE_HPF3_FILTER 6
#define MFPA_STAGE_HPF3_CONV 7    /* (+ final normalisation in mfpa_augment) */
#define MFPA_STAGE_STFT 8
#define MFPA_STAGE_PEAKS 9
#define MFPA_STAGE_LANDMARKS 10   /* landmarks + hashes (+ shift merge) */
#define MFPA_N_STAGES 11
/* ms_out[MFPA_N_STAGES]: mean milliseconds per stage (0 for stages that did not run); returns the number of calls
 * averaged, or a negative error.  Synchronises with the recorded events. */
int mfpa_stage_times(mfpa_ctx* ctx, float* ms_out);

/* ---- geometry --------------------------------------------------------- */
int mfpa_num_frames(int n_samples);           /* afp/audfprint/stft.py:50-53 */
int mfpa_shift_offset(int shift, int shifts); /* peak_extractor.py:412 */

/* ---- S2: magnitude STFT  (afp/audfprint/stft.py:15-62 + abs, peak_extractor.py:257-261)
 * x_dev: B rows of T float32 samples, row stride `x_stride` elements.
 * mag_dev: [B*shifts][mfpa_num_frames(T)][MFPA_MAG_PITCH] float32, frame-major.
 *          Elements 0..256 of each frame row are |rfft|.  Elements 258/259 of every EVEN frame row
 *          carry picker statistics of the frame pair (2k, 2k+1): the sum of log2 and the minimum
 *          of their magnitudes (mfpa_audfprint_peaks uses them to skip a pass over the data);
 *          the remaining elements are padding.
 * qmax_dev: [B*shifts] float32, max of each item's magnitudes (the divisor of
 *          `sgram /= np.max(sgram)`, peak_extractor.py:263). */
int mfpa_stft_mag(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, int shifts,
                  float* mag_dev, float* qmax_dev, void* stream);

/* The module's public function: stft(signal, n_fft, hop_length, window) (afp/audfprint/stft.py:15-62), any
 * n_fft / hop / window, computed in float64 like the reference (np.pad reflect by n_fft/2, frames of
 * len(window) samples every `hop`, times the window, rfft of n_fft points).
 * x_dev: T float64 samples; window_dev: win_len float64 values;
 * out_dev: complex128 [n_fft/2 + 1][n_frames] (interleaved re, im), n_frames = 1 + (T + 2 (n_fft/2) - win_len) / hop
 * as returned by mfpa_stft_num_frames.  Not the throughput path (that is mfpa_stft_mag). */
int mfpa_stft_num_frames(int n_samples, int n_fft, int hop, int win_len);
int mfpa_stft_complex(mfpa_ctx* ctx, const double* x_dev, int T, int n_fft, int hop, const double* window_dev,
                      int win_len, double* out_dev, void* stream);

/* Normalised spectrogram in the reference's layout: spec[item][257][n_frames_max]
 * float64 = mag/qmax — the `spec` returned by find_peaks (peak_extractor.py:271,311). */
int mfpa_spec_from_mag(mfpa_ctx* ctx, const float* mag_dev, const float* qmax_dev, int B, int T,
                       int shifts, double* spec_dev, void* stream);

/* ---- S3: audfprint peak picking  (peak_extractor.py:263-311: /max, log, -mean,
 * lfilter, _decaying_threshold_fwd_prune :173-204, _bwd_prune_peaks :206-234)
 * Output: one 64-bit record per (item, frame): byte0 = number of peaks (<=5),
 * bytes 1..5 = their bins in ascending order.  rec_dev: [B*shifts][n_frames_max].
 * npeaks_dev: [B*shifts] int32 total per item (may be NULL).
 * With qmax_dev != NULL, mag_dev must have been written by mfpa_stft_mag (statistics in the row
 * padding, see above); with qmax_dev == NULL the float64 picker runs on elements 0..256 only. */
int mfpa_audfprint_peaks(mfpa_ctx* ctx, const float* mag_dev, const float* qmax_dev, int B, int T,
                         int shifts, const mfpa_afp_params* p, uint64_t* rec_dev,
                         int32_t* npeaks_dev, void* stream);

/* Same picker fed a float64 spectrogram in the REFERENCE layout [items][rows][n_frames]
 * (parity entry: "bit-exact when fed the reference spectrogram").
 * stage 0: spec is find_peaks' magnitude spectrogram, rows = 257 (normalised or not);
 * stage 1: spec is the already filtered sgram of :286-290, rows = 256. */
int mfpa_audfprint_peaks_from_spec(mfpa_ctx* ctx, const double* spec_dev, int items, int n_frames,
                                   int stage, const mfpa_afp_params* p, uint64_t* rec_dev,
                                   int32_t* npeaks_dev, void* stream);

/* records -> (col, bin) int32 rows in find_peaks' pklist order (:303-309);
 * peaks_dev: [items][cap][2]. */
int mfpa_peaks_list(mfpa_ctx* ctx, const uint64_t* rec_dev, int items, int n_frames,
                    int32_t* peaks_dev, int cap, int32_t* npeaks_dev, void* stream);
/* records -> peaks_mask float32 [items][256][n_frames] (:303). */
int mfpa_peaks_mask(mfpa_ctx* ctx, const uint64_t* rec_dev, int items, int n_frames,
                    float* mask_dev, void* stream);

/* ---- S4: landmarks + 20-bit hashes  (peaks2landmarks :313-346, landmarks2hashes :40-58)
 * hashes_dev: [items][cap][2] int32 rows (time, hash); nh_dev: [items].
 * sorted = 0: rows in the reference's landmark order;
 * sorted = 1: rows ordered by (time, hash) — for shifts == 1 this is
 *             wavfile2hashes' final unique/sorted array (:448-460).
 * n_frames_item_dev (nullable): per-item frame counts when they differ. */
int mfpa_landmark_hashes(mfpa_ctx* ctx, const uint64_t* rec_dev, int items, int n_frames,
                         const mfpa_afp_params* p, int sorted, int32_t* hashes_dev, int cap,
                         int32_t* nh_dev, void* stream);

/* Concatenate the `shifts` lists of each query, unique + sort on (time, hash)
 * (wavfile2hashes :437-460).  in: [B*shifts][cap_in][2] (each list sorted by
 * (time,hash)); out: [B][cap_out][2]. */
int mfpa_merge_shifts(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B,
                      int shifts, int cap_in, int n_frames, int32_t* out_dev, int cap_out,
                      int32_t* nout_dev, void* stream);

/* ---- fused S2-S4: Audfprint_peaks.wavfile2hashes without the file read (:426-460)
 * x -> unique sorted (time, hash) rows.  hashes_dev: [B][cap][2], nh_dev: [B]. */
int mfpa_fingerprint(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, int shifts,
                     const mfpa_afp_params* p, int32_t* hashes_dev, int cap, int32_t* nh_dev,
                     void* stream);
/* Host-buffer form of mfpa_fingerprint (the call a reference user makes once per batch
 * instead of looping wavfile2hashes over files).  x_host: [B][T] float32 (pinned memory
 * lets the copies overlap the kernels; pageable memory works, slower).  The batch is
 * processed in chunks: host->device copy of chunk i+1 overlaps the kernels of chunk i.
 * Output is CSR: rows of query q are rows_host[offsets_host[q] .. offsets_host[q+1]) as
 * (time, hash) int32 pairs; offsets_host has B+1 entries.  Returns MFPA_ECAP (offsets
 * still filled in) when more than rows_cap rows were produced. */
int mfpa_fingerprint_host(mfpa_ctx* ctx, const float* x_host, int B, int T, int shifts,
                          const mfpa_afp_params* p, int32_t* rows_host, int64_t rows_cap,
                          int64_t* offsets_host);

/* Same, for 16-bit PCM input (what a decoded WAV/MP3 query holds before the reference's loaders turn
 * it into float32, afp/audfprint/peak_extractor.py:348-389): samples are x / 32768, exactly the
 * float32 values a host conversion would produce, converted on the device after a copy half the size.
 * Results are identical to mfpa_fingerprint_host on (float)x / 32768. */
int mfpa_fingerprint_host_pcm16(mfpa_ctx* ctx, const int16_t* x_host, int B, int T, int shifts,
                                const mfpa_afp_params* p, int32_t* rows_host, int64_t rows_cap,
                                int64_t* offsets_host);

/* Ragged [items][cap][2] rows + counts -> CSR on the device: offsets_dev [items+1] int64
 * (exclusive scan of min(n, cap)), rows_dev [>= total][2]. */
int mfpa_compact_rows(mfpa_ctx* ctx, const int32_t* rows_in_dev, const int32_t* n_dev, int items, int cap,
                      int64_t* offsets_dev, int32_t* rows_dev, int64_t rows_cap, void* stream);

/* ---- S1: AugmentFP degradation chain  (augmentation/__init__.py:46-97) -------------
 * One record per query with the parameters the reference samples in
 * randomize_parameters() (SURVEY.md App. C dump schema).  `apply` bit i set = the i-th
 * transform's Bernoulli gate (transform.py:101-105) fired for this query. */
#define MFPA_AUG_HPF1 1u  /* HighPassFilter  "cutoff_freq1" (loudspeaker)  pass_filters.py:144-155 */
#define MFPA_AUG_IR 2u    /* ApplyImpulseResponse                           impulse_response.py:73-116 */
#define MFPA_AUG_NOISE 4u /* AddBackgroundNoise                             background_noise.py:183-213 */
#define MFPA_AUG_GAIN 8u  /* Gain                                           gain.py:62-70 */
#define MFPA_AUG_CLIP 16u /* Clipping                                       clipping.py:67-101 */
#define MFPA_AUG_LPF 32u  /* LowPassFilter   "cutoff_freq2"                 pass_filters.py:84-115 */
#define MFPA_AUG_HPF3 64u /* HighPassFilter  "cutoff_freq3" (microphone) */
#define MFPA_AUG_NORM 128u /* final PeakNormalization (p = 1 in AugmentFP)  peak_normalization.py:38-67 */

typedef struct mfpa_aug_params {
  uint32_t apply;     /* MFPA_AUG_* bits */
  float fc1_hz;       /* transform_parameters["cutoff_freq"] of the three pass filters, in Hz */
  float fc2_hz;
  float fc3_hz;
  float snr_db;       /* transform_parameters["snr_in_db"] */
  float gain_factor;  /* transform_parameters["gain_factors"] = 10^(dB/20) */
  float clip_p;       /* transform_parameters["percentile_threshold"] */
  int32_t ir_len;     /* valid samples of this query's impulse response (<= ir_stride) */
} mfpa_aug_params;

/* Longest FIR the CUDA path accepts (taps = 2*int(4*sr/fc)+1 <= this: cut-offs down to 0.061 Hz at
 * 8 kHz).  Filters of up to 8193 taps (fc >= 7.8 Hz at 8 kHz) take one overlap-save block per 8192+
 * outputs; longer ones - the reference draws the loudspeaker cut-off mel-uniformly from [0, 150] Hz, so
 * about 6 % of its default draws - run as uniformly partitioned overlap-save. */
#define MFPA_AUG_MAX_TAPS 1048577
/* Longest impulse response, in samples (32 s at 8 kHz); above 8192 samples: partitioned as well. */
#define MFPA_AUG_MAX_IR 262144

/* AugmentFP.__call__ / batch_augment on dumped parameters.  Clipping's quantiles are per query (what the reference
 * computes at B = 1, the only batch size its drivers use) unless MFPA_OPT_CLIP_POOLED asks for the pooled rule of
 * larger batches; the final PeakNormalization applies to the queries that have MFPA_AUG_NORM set.
 * x_dev [B][T] (row stride x_stride), ir_dev [B][ir_stride] (may be NULL when no query has
 * MFPA_AUG_IR), noise_dev [B][T] contiguous, RMS-normalised like random_background()
 * leaves it (may be NULL when no query has MFPA_AUG_NOISE), params_host [B],
 * out_dev [B][T] contiguous.  Cut-offs that the reference rejects (<= 0 or > sr/2,
 * pass_filters.py:103-110) and filters longer than MFPA_AUG_MAX_TAPS return MFPA_EINVAL. */
int mfpa_augment(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, int sample_rate,
                 const mfpa_aug_params* params_host, const float* ir_dev, int ir_stride,
                 const float* noise_dev, float* out_dev, void* stream);

/* One filter of a julius.LowPassFilters bank per row (julius 0.2.7 lowpass.py; the reference's BandPassFilter /
 * BandStopFilter call julius.bandpass_filter = bank([low, high]) and subtract, band_filters.py:126-156,195):
 * taps = hann(2 half + 1) * 2c * sinc(2c pi t) normalised to sum 1 with c = cutoff_host[row] and
 * half = int(8 / width_host[row] / 2) - the bank gives all its filters the window of its lowest cut-off, so a band
 * filter passes the low cut-off as `width` of both rows; replicate padding; any length up to MFPA_AUG_MAX_TAPS.
 * Both arrays are fractions of the sample rate in (0, 0.5] (else MFPA_EINVAL, where julius raises ValueError).
 * x_dev [B][T] (row stride x_stride), out_dev [B][T] contiguous.  Runs the low-pass stage of mfpa_augment. */
int mfpa_lowpass_filters(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, const double* cutoff_host,
                         const double* width_host, float* out_dev, void* stream);

/* ---- noise preparation ahead of the chain
 * AddBackgroundNoise.random_background (augmentation/transformations/background_noise.py:64-141):
 * a query's noise row is the concatenation of pieces cut from randomly chosen background files
 * (a mix-up pair contributes samplePairing = (a + b) / 2, :11-12), each piece RMS-normalised
 * (x / (rms + 1e-8), augmentation/utils.py:189-205), the concatenation RMS-normalised again.
 * The random choices stay on the host (the drop-in draws them in the reference's RNG order);
 * the arithmetic runs here on a device-resident bank of decoded background audio.
 * bank_dev: float32 [bank_len], every background file back to back at the target sample rate.
 * pieces_host [n_pieces]: the pieces of all B rows; every row must be covered exactly once
 * (pieces of a row tile [0, T)).  out_dev [B][T] contiguous. */
typedef struct mfpa_noise_piece {
  int64_t src_a;    /* first sample of the piece in bank_dev */
  int64_t src_b;    /* second source of a mix-up pair, or -1 */
  int32_t query;    /* row of out_dev */
  int32_t dst;      /* first sample in the row */
  int32_t len;      /* samples (>= 1) */
  int32_t reserved;
} mfpa_noise_piece;
int mfpa_noise_assemble(mfpa_ctx* ctx, const float* bank_dev, int64_t bank_len, const mfpa_noise_piece* pieces_host,
                        int n_pieces, int B, int T, float* out_dev, void* stream);

/* Fused S1-S4 (BASELINE.json config 3): augment, then fingerprint the degraded queries
 * without leaving the device.  The final peak normalisation is skipped on this path
 * (the spectrogram is divided by its own maximum, peak_extractor.py:263). */
int mfpa_augment_fingerprint(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, int sample_rate,
                             const mfpa_aug_params* params_host, const float* ir_dev, int ir_stride,
                             const float* noise_dev, int shifts, const mfpa_afp_params* p,
                             int32_t* hashes_dev, int cap, int32_t* nh_dev, void* stream);

/* Host-buffer form of mfpa_augment_fingerprint: the call a user of testing/generate_queries.py +
 * audfprint_exps.py makes once per batch instead of looping AugmentFP.__call__ (augmentation/__init__.py:95-97)
 * and wavfile2hashes (peak_extractor.py:426-460) over files.  Only the QUERIES stream from the host (pinned
 * memory lets the copies overlap the kernels), chunked and double-buffered like mfpa_fingerprint_host; the
 * degradation sources are device-resident:
 *   impulse responses  ir_dev + (ir_offsets_host ? ir_offsets_host[q] : q * ir_stride), params_host[q].ir_len samples
 *                      (ir_bank_len = samples in ir_dev when offsets are given)
 *   noise              noise_dev [B][T] rows as random_background() leaves them, OR a bank of decoded background
 *                      audio + the pieces of every row (mfpa_noise_assemble, assembled chunk by chunk; pieces
 *                      grouped by query, queries ascending).
 * Output is CSR like mfpa_fingerprint_host. */
typedef struct mfpa_chain_inputs {
  const float* x_host;              /* [B][T] float32 ... */
  const int16_t* x_pcm16_host;      /* ... or [B][T] 16-bit PCM, samples x / 32768 (exactly one of the two) */
  const float* ir_dev;              /* NULL when no query has MFPA_AUG_IR */
  int64_t ir_bank_len;
  const int64_t* ir_offsets_host;   /* [B] or NULL */
  int32_t ir_stride;
  int32_t n_pieces;
  const float* noise_dev;           /* [B][T] or NULL */
  const float* noise_bank_dev;      /* with pieces_host */
  int64_t noise_bank_len;
  const mfpa_noise_piece* pieces_host;
} mfpa_chain_inputs;
int mfpa_augment_fingerprint_host(mfpa_ctx* ctx, const mfpa_chain_inputs* in, int B, int T, int sample_rate,
                                  const mfpa_aug_params* params_host, int shifts, const mfpa_afp_params* p,
                                  int32_t* rows_host, int64_t rows_cap, int64_t* offsets_host);

/* ---- S5: landmark-hash matching  (HashTable.get_hits hash_table.py:220-246,
 * Matcher._best_count_ids / _approx_match_counts / match_hashes audfprint_match.py:102-129,235-349)
 *
 * mfpa_index_load uploads one hash-range shard of a reference HashTable: the rows
 * [hash_lo, hash_lo + n_buckets) of `table` (uint32 [n_buckets][depth], entry =
 * ((id+1) << maxtimebits) + time, hash_table.py:53-68,100) and of `counts`, plus the full
 * `hashesperid` [n_tracks].  hashbits is the table's (20).  A single-GPU user passes
 * hash_lo = 0, n_buckets = 1 << hashbits. */
int mfpa_index_load(mfpa_ctx* ctx, const uint32_t* table_host, const int32_t* counts_host, int hash_lo,
                    int n_buckets, int depth, int hashbits, int maxtimebits,
                    const uint32_t* hashesperid_host, int n_tracks);

typedef struct mfpa_match_params {  /* Matcher.__init__ defaults, audfprint_match.py:76-100 */
  int32_t window;        /* 2   */
  int32_t threshcount;   /* 5   */
  int32_t search_depth;  /* 100 (<= 128) */
  int32_t max_alignments_per_id; /* 100 */
} mfpa_match_params;
void mfpa_match_defaults(mfpa_match_params* p);

/* get_hits for ONE query: hashes_dev [n][2] (time, hash) -> hits_dev [<= hits_cap][4] int32 rows
 * [id, t_ref - t_q, hash, t_q] in the reference's order; *nhits_dev (int64, device) = rows produced
 * by this shard (may exceed hits_cap; only hits_cap rows are written). */
int mfpa_get_hits(mfpa_ctx* ctx, const int32_t* hashes_dev, int n, int32_t* hits_dev, int64_t hits_cap,
                  int64_t* nhits_dev, void* stream);

/* The four steps of match_hashes for a batch of B queries (hashes_dev [B][cap][2], nh_dev [B]).
 * Multi-GPU (index sharded by hash range): sum counts_dev over the shards between steps 1 and 2 (all-reduce,
 * or reduce-scatter to the rank that owns the query), bring every shard's list_dev/nlist_dev of a query
 * together between steps 3 and 4 (all-gather / all-to-all); single GPU: call mfpa_match.
 *  1 counts : counts_dev [B][n_tracks] int32 raw hit counts from this shard's buckets
 *  2 select : cand_dev [B][search_depth][2] (id, raw) in _best_count_ids order, ncand_dev [B]
 *  3 collect: list_dev [B][list_cap] uint32 = (candidate index << 16) | (t_ref - t_q + 16384),
 *             nlist_dev [B] (may exceed list_cap -> step 4 reports -1 for that query)
 *  4 align  : results_dev [B][max_rows][7] int32 rows [id, filtered count, time skew, raw count,
 *             candidate rank, 0, 0] ordered by filtered count descending; nrows_dev [B]
 *             (-1 / -2 = a capacity was exceeded for that query). */
int mfpa_match_counts(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B, int cap,
                      int32_t* counts_dev, void* stream);
int mfpa_match_select(mfpa_ctx* ctx, const int32_t* counts_dev, int B, const mfpa_match_params* p,
                      int32_t* cand_dev, int32_t* ncand_dev, void* stream);
int mfpa_match_collect(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B, int cap,
                       const int32_t* cand_dev, const int32_t* ncand_dev, const mfpa_match_params* p,
                       uint32_t* list_dev, int list_cap, int32_t* nlist_dev, void* stream);
int mfpa_match_align(mfpa_ctx* ctx, const uint32_t* lists_dev, const int32_t* nlists_dev, int n_lists, int B,
                     int list_cap, const int32_t* cand_dev, const int32_t* ncand_dev,
                     const mfpa_match_params* p, int32_t* results_dev, int32_t* nrows_dev, int max_rows,
                     void* stream);
/* Sparse exchange for an index sharded by hash range: instead of summing dense per-track
 * histograms across the shards, every shard lists the hits of ITS buckets as 32-bit words
 * (track << 15) | (t_ref - t_q + 16384) - a few thousand per query and shard - and ONE all-to-all keyed by the
 * query's owner rank moves them; the owner then runs the whole of match_hashes on the words of all shards.
 *  emit : words_dev [B][words_cap], nwords_dev [B] hits of this shard per query (may exceed words_cap: the owner
 *         step then reports -1 for the query; -5 = a query time outside [0, 16384))
 *  owner: words_dev [n_shards][B][words_cap] + nwords_dev [n_shards][B] as the all-to-all leaves them ->
 *         results / nrows like step 4 above.  Indexes of up to 110000 tracks. */
int mfpa_match_emit(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B, int cap, uint32_t* words_dev,
                    int words_cap, int32_t* nwords_dev, void* stream);
int mfpa_match_owner(mfpa_ctx* ctx, const uint32_t* words_dev, const int32_t* nwords_dev, int n_shards, int B, int words_cap,
                     const mfpa_match_params* p, int32_t* results_dev, int32_t* nrows_dev, int max_rows, void* stream);
/* ---- the exchange fused into the sweep: peer memory over NVLink (one process per GPU, same node)
 * Replaces the NCCL all-to-all between mfpa_match_emit and mfpa_match_owner: every rank allocates its receive
 * buffers with mfpa_peer_alloc, passes the 64-byte handle to the other ranks (any side channel - the Python host
 * uses torch.distributed's object all-gather) and maps theirs with mfpa_peer_open.  mfpa_match_emit_peer then
 * writes the words of query q of a sub-batch of B queries (B a multiple of world) directly into the memory of its
 * owner rank q / (B / world), at words[owner] + ((rank * own + q % own) * words_cap) and nwords[owner][rank * own +
 * q % own] - i.e. the owner's buffer ends up as the [n_shards = world][own][words_cap] array mfpa_match_owner reads.
 * mfpa_peer_barrier is the cross-rank barrier between the two (a one-block kernel on `stream`: writes `epoch` into
 * slot `rank` of every rank's flags row - MFPA_MAX_PEERS words, zero-initialised by mfpa_peer_alloc - and waits for
 * every slot of its own row; epochs must increase by one per barrier on every rank; a rank that never arrives makes
 * the waiting kernels trap after 20 s).  Pointers of the set are device pointers valid in THIS process. */
#define MFPA_MAX_PEERS 16
typedef struct {
  uint32_t* words[MFPA_MAX_PEERS];
  int32_t* nwords[MFPA_MAX_PEERS];
  uint32_t* flags[MFPA_MAX_PEERS];
  int32_t world, rank;
} mfpa_peer_set;
int mfpa_peer_alloc(mfpa_ctx* ctx, uint64_t bytes, void** dev_ptr, void* handle64);   /* cudaMalloc + zero + IPC handle */
int mfpa_peer_open(mfpa_ctx* ctx, const void* handle64, void** dev_ptr);              /* map another rank's buffer */
int mfpa_peer_close(mfpa_ctx* ctx, void* dev_ptr);
int mfpa_peer_free(mfpa_ctx* ctx, void* dev_ptr);
int mfpa_match_emit_peer(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B, int cap,
                         const mfpa_peer_set* peers, int words_cap, void* stream);
int mfpa_peer_barrier(mfpa_ctx* ctx, const mfpa_peer_set* peers, uint32_t epoch, void* stream);
/* Single-shard convenience: steps 1-4 with internal scratch, queries processed in sub-batches.  nrows_dev[q] < 0 flags a
 * query that could not be matched: -1 / -2 a capacity was exceeded, -5 a query time outside [0, 16384) (the hit
 * keys hold t_ref - t_q next to the table's 14-bit reference times; the reference has no such limit), -6 one track
 * collected >= 65536 hits and the packed 16-bit counters of the one-kernel matcher overflowed (detected by a
 * checksum of the histogram, never silent; such a query is also past the 8192 candidate hits of step 4, and
 * MFPA_OPT_MATCH_UNFUSED = 2 gives its exact int32 raw counts through mfpa_match_counts). */
int mfpa_match(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B, int cap,
               const mfpa_match_params* p, int32_t* results_dev, int32_t* nrows_dev, int max_rows, void* stream);

/* ---- Dejavu 2-D peak finder  (afp/dejavu/fingerprint.py:94-171; PEAK_NEIGHBORHOOD_SIZE 10,
 * amp_min 50: afp/dejavu/variables.py:19, testing/parameters.py:27-34)
 * arr_dev: [B][F][N] float64 (is_f64 = 1, the reference's dtype) or float32, the log spectrogram
 * that fingerprint() passes to get_2D_peaks.  mask_dev: [B][F][N] uint8 peak mask.
 * peaks_dev (nullable): [B][cap][2] int32 (freq, time) rows in np.where order; npeaks_dev [B]. */
int mfpa_dejavu_peaks(mfpa_ctx* ctx, const void* arr_dev, int is_f64, int B, int F, int N, int neighborhood,
                      double amp_min, uint8_t* mask_dev, int32_t* peaks_dev, int cap, int32_t* npeaks_dev,
                      void* stream);

/* Front end of Dejavu's fingerprint() (afp/dejavu/fingerprint.py:60-79):
 *   mfpa_dejavu_psd: matplotlib.mlab.specgram(x, NFFT=512, Fs, window_hanning, noverlap=256)[0] divided by its
 *     maximum (:60-68) -> psd_dev float32 [B][257][mfpa_dejavu_num_frames(T)] ((T - 256) / 256 segments, no
 *     padding; the /Fs and /sum(w^2) factors cancel in the normalisation);
 *   mfpa_dejavu_log: arr = 10 ln(max(p, max(p) / 1e6)) - mean(...) per item (:77-79), p = psd, or psd^2 when
 *     `square` (the UNet path squares the network output, :74) -> arr_dev float32 [B][F][N], the array
 *     mfpa_dejavu_peaks takes. */
int mfpa_dejavu_num_frames(int n_samples);
int mfpa_dejavu_psd(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, float* psd_dev, void* stream);
int mfpa_dejavu_log(mfpa_ctx* ctx, const float* psd_dev, int B, int F, int N, int square, float* arr_dev, void* stream);

/* ---- in-memory Dejavu index and offset vote  (replaces the Postgres tables behind Dejavu.find_matches /
 * align_matches: afp/dejavu/postgres_database.py:180-229, afp/dejavu/dejavu.py:295-378)
 * A fingerprint row is (hash, song, offset); a Dejavu hash is 20 hex digits of a SHA-1 (80 bits) = key (leading
 * 16 digits as uint64) + tail (last 4 digits as uint16).  Rows must be sorted by key. */
typedef struct mfpa_dejavu_index mfpa_dejavu_index;
int mfpa_dejavu_index_create(mfpa_ctx* ctx, const uint64_t* keys_host, const uint16_t* tails_host, const int32_t* songs_host,
                             const int32_t* offsets_host, int64_t n, int n_songs, mfpa_dejavu_index** out);
void mfpa_dejavu_index_destroy(mfpa_dejavu_index* index);
/* return_matches for one query: the DISTINCT query hashes (qkeys/qtails [n_hashes]) with the offsets each was seen at
 * (CSR: qoffsets_dev[qstart_dev[i] .. qstart_dev[i+1])).  For every index row with an equal hash and every such query
 * offset: one (song, row offset - query offset) pair -> pairs_dev [pairs_cap][2] (unordered; *n_pairs_dev may exceed
 * pairs_cap, then call again with room); dedup_dev [n_songs] = matching rows per song, each distinct hash counted once
 * (the reference's dedup_hashes). */
int mfpa_dejavu_return_matches(mfpa_dejavu_index* index, const uint64_t* qkeys_dev, const uint16_t* qtails_dev,
                               const int32_t* qstart_dev, const int32_t* qoffsets_dev, int n_hashes, int32_t* pairs_dev,
                               int64_t pairs_cap, int64_t* n_pairs_dev, int32_t* dedup_dev, void* stream);
/* align_matches, top match: best_dev[3] = (song, offset difference, number of pairs voting for it) by count descending,
 * ties to the smaller song id, then the smaller offset (what the reference's stable sorts return first); song -1 when
 * there is no pair. */
int mfpa_dejavu_align(mfpa_dejavu_index* index, const int32_t* pairs_dev, int64_t n_pairs, int32_t* best_dev, void* stream);

/* ---- evaluation after the path  (testing/metrics.py:7-192, used by compute_peaks_metrics, testing/audfprint_exps.py:86-157)
 * mfpa_mask_metrics: the sums behind Recall / Precision / F1score of two peak masks [B][H][W]:
 *   out4_dev = { #{gt != 0}, sum of predicted looked at from {gt != 0}, #{predicted != 0}, sum of gt looked at from
 *   {predicted != 0} }.  metrics.py's 3x3 kernel is zero off its centre, so a peak at (i, j) looks at one position of the
 *   other mask: (i, j), except (i + 1, j) for i == 0 and (i, j + 1) for j == 0 (the reference slices window and kernel
 *   from opposite ends there, :44-47, :72-75);
 *   recall = out[1] / out[0], precision = out[3] / out[2] (0 when the denominator is 0, :33-34, :112-113).
 * mfpa_psnr_stats: out3_dev = { sum (pred - target)^2, min(target), max(target) } in float64 - what torchmetrics'
 *   PeakSignalNoiseRatio accumulates. */
int mfpa_mask_metrics(mfpa_ctx* ctx, const float* predicted_dev, const float* gt_dev, int B, int H, int W, double* out4_dev,
                      void* stream);
int mfpa_psnr_stats(mfpa_ctx* ctx, const double* pred_dev, const double* target_dev, int64_t n, double* out3_dev, void* stream);

/* ---- optional UNet magnitude-spectrogram denoiser  (training/unet.py:75-108: UNet(1, 1, bilinear=False),
 * eval mode; inserted between `sgram /= max` and the log at afp/audfprint/peak_extractor.py:265-269 and
 * afp/dejavu/fingerprint.py:70-75).  bf16 activations/weights, fp32 accumulation, BatchNorm folded;
 * every 3x3 / transposed convolution is a tcgen05 implicit GEMM (csrc/unet.cu). */
typedef struct mfpa_unet mfpa_unet;
/* Number of float32 values mfpa_unet_load expects: the reference model's state_dict() in its own
 * order, every floating tensor flattened and concatenated, `num_batches_tracked` entries skipped. */
int64_t mfpa_unet_num_params(void);
int mfpa_unet_create(mfpa_ctx* ctx, mfpa_unet** out);
void mfpa_unet_destroy(mfpa_unet* unet);
int mfpa_unet_load(mfpa_unet* unet, const float* params_host, int64_t n_floats);
/* Images per pass through the network (activation arena = ~45 MB per 257x251 image); default 37 (a
 * quarter of the 148 SMs: every layer's tile count is then a whole number of waves). */
int mfpa_unet_set_max_chunk(mfpa_unet* unet, int images);
/* out[b] = unet(in[b] / div[b]) for B single-channel H x W images.  in/out are float32 device arrays
 * addressed as base + b*stride_n + h*stride_h + w*stride_w (elements), so both the reference layout
 * [B][H][W] and the frame-major magnitude layout of mfpa_stft_mag (h = bin: stride 1, w = frame: stride
 * MFPA_MAG_PITCH) work; in == out is allowed.  div_dev may be NULL. */
int mfpa_unet_forward(mfpa_ctx* ctx, mfpa_unet* unet, const float* in_dev, int64_t in_stride_n, int64_t in_stride_h,
                      int64_t in_stride_w, const float* div_dev, int B, int H, int W, float* out_dev,
                      int64_t out_stride_n, int64_t out_stride_h, int64_t out_stride_w, void* stream);
/* One convolution of the network on caller-owned NHWC bf16 tensors (the unit the parity tests drive):
 * out[..., coff:coff+cout] = act(conv(in, w) * scale + shift); w_dev [cout][taps][cin] bf16, taps 9 (3x3,
 * padding 1) or 1; cin, cout multiples of 64.  bn / mt / stages = 0 pick the tile configuration.
 * halo_wh > 0 selects the halo-reuse 3x3 kernel with that halo pitch (tile width halo_wh - 2); wres = 1
 * additionally keeps the weights resident in shared memory (cout == 64 only). */
int mfpa_conv_bf16(mfpa_ctx* ctx, const void* in_dev, int N, int H, int W, int cin, const void* w_dev, int cout,
                   int taps, const float* scale_dev, const float* shift_dev, int relu, void* out_dev, int ldc,
                   int coff, int bn, int mt, int stages, int halo_wh, int wres, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MFPA_H */
