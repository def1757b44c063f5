// Shared declarations for libmfpa (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/mfpa.h"

#ifndef __CUDA_ARCH__
#define MFPA_HOST_ONLY
#endif

namespace mfpa {

constexpr int kNfft = MFPA_N_FFT;
constexpr int kHop = MFPA_HOP;
constexpr int kBins = MFPA_BINS;
constexpr int kRows = MFPA_ROWS;
constexpr int kPitch = MFPA_MAG_PITCH;
constexpr int kMaxPks = MFPA_MAX_PKS;
constexpr int kSpreadLen = 2 * MFPA_ROWS + 1;  // 513
constexpr int kNumSMs = 148;                   // B200

void set_error(const char* fmt, ...);

#define MFPA_CUDA(call)                                                              \
  do {                                                                               \
    cudaError_t _e = (call);                                                         \
    if (_e != cudaSuccess) {                                                         \
      mfpa::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return MFPA_ECUDA;                                                             \
    }                                                                                \
  } while (0)

#define MFPA_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      mfpa::set_error(__VA_ARGS__);    \
      return MFPA_EINVAL;              \
    }                                  \
  } while (0)

// A grow-only device buffer owned by the context.
struct Scratch {
  void* ptr = nullptr;
  size_t bytes = 0;
  int reserve(size_t need);
  void release();
};

inline int num_frames(int n_samples) { return 1 + n_samples / kHop; }
inline int shift_offset(int shift, int shifts) {
  // int(shift / shifts * n_hop) in Python float arithmetic (peak_extractor.py:412)
  if (shifts < 2) return 0;
  return (int)((double)shift / (double)shifts * (double)kHop);
}

}  // namespace mfpa

struct mfpa_ctx {
  int device = 0;
  int num_sms = mfpa::kNumSMs;
  int opt_peaks_f64 = 0;            // MFPA_OPT_PEAKS_F64
  int opt_match_packed = 0;         // MFPA_OPT_MATCH_PACKED
  int opt_match_unfused = 0;        // MFPA_OPT_MATCH_UNFUSED
  int opt_part_budget_mb = 2048;    // MFPA_OPT_PART_BUDGET_MB
  int opt_conv_occ = 3;             // MFPA_OPT_CONV_OCC
  int opt_stage_times = 0;          // MFPA_OPT_STAGE_TIMES
  int opt_clip_pooled = 0;          // MFPA_OPT_CLIP_POOLED
  // per-stage CUDA events of the last (up to 16) mfpa_augment_fingerprint / mfpa_fingerprint calls
  cudaEvent_t stage_ev[16][MFPA_N_STAGES + 1] = {};
  int stage_calls = 0, stage_slot = 0;
  unsigned stage_marked[16] = {};   // bit k: stage k (or the end mark MFPA_N_STAGES) was stamped in this slot
  bool in_host_pipeline = false;    // inside mfpa_augment_fingerprint_host: small uploads go through launch_pull
  bool stage_chained = false;       // the next mfpa_fingerprint continues the record an augment call opened
  double* spread_dev = nullptr;     // [513] Gaussian table
  float2* tw_dev = nullptr;         // FFT twiddles (stft.cu layout)
  float* win_dev = nullptr;         // [512] analysis window
  float* win_dejavu_dev = nullptr;  // [512] np.hanning(512) x 1/2 (mlab.window_hanning, afp/dejavu/fingerprint.py:64)
  mfpa::Scratch mag, qmax, rec, fwd, hashes, nh, misc, spec64, xin, out_h, out_n;
  // augmentation / matching state is appended by their translation units
  mfpa::Scratch aug_a, aug_b, aug_c, aug_d, aug_small, aug_lists, aug_long, aug_part, aug_noise, aug_pool;
  float2* aug_tw_dev = nullptr;     // planar twiddle tables of the 8192-point FFT (augment.cu, fftconv_core.cuh)
  void* aug_pinned = nullptr;
  void* aug_copy_done = nullptr;    // cudaEvent_t: the last H2D copy out of aug_pinned
  size_t aug_pinned_bytes = 0;
  // index shard (match.cu)
  uint32_t* index_table = nullptr;
  int32_t* index_counts = nullptr;
  uint32_t* index_hashesperid = nullptr;
  uint32_t index_hp_min = 1;        // smallest non-zero hashesperid (lets the matcher skip tracks by raw count alone)
  int index_hash_lo = 0, index_hash_hi = 0, index_depth = 0, index_ntracks = 0, index_maxtimebits = 14;
  int index_hashmask = (1 << 20) - 1;
  mfpa::Scratch match_a, match_b, match_c;
  void* pinned = nullptr;
  size_t pinned_bytes = 0;
  // host-path pipeline (capi.cu): copy-in, compute and copy-out streams, double-buffered chunks
  cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_run[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  mfpa::Scratch h_x[2], h_x16[2], h_rows[2], h_csr[2], h_n[2], h_off[2], h_noise[2];
  void* pieces_pinned[2] = {nullptr, nullptr};   // per-slot staging of a chunk's noise pieces (mfpa_augment_fingerprint_host)
  size_t pieces_pinned_bytes[2] = {0, 0};
};

// ---- kernel launchers implemented in the stage translation units ----------
namespace mfpa {

int launch_stft_mag(mfpa_ctx* ctx, const float* x, int B, int T, int64_t stride, int shifts,
                    float* mag, float* qmax, cudaStream_t st, const float* window = nullptr);
int launch_stft_complex(const double* x, int T, int n_fft, int hop, const double* win, int win_len, int n_frames,
                        double* out, cudaStream_t st);
int dejavu_num_frames(int T);
int launch_dejavu_psd(mfpa_ctx* ctx, const float* x, int B, int T, int64_t stride, float* psd, cudaStream_t st);
int launch_dejavu_log(const float* psd, int B, int n, int square, float* arr, cudaStream_t st);
int launch_spec_from_mag(const float* mag, const float* qmax, int B, int T, int shifts,
                         double* spec, cudaStream_t st);
int launch_peaks_f32(mfpa_ctx* ctx, const float* mag, const float* qmax, int B, int T, int shifts,
                     const mfpa_afp_params& p, uint64_t* rec, int32_t* npeaks, cudaStream_t st);
int launch_peaks_from_spec(mfpa_ctx* ctx, const double* spec, int items, int n_frames, int stage,
                           const mfpa_afp_params& p, uint64_t* rec, int32_t* npeaks, cudaStream_t st);
int launch_peaks_list(const uint64_t* rec, int items, int n_frames, int32_t* peaks, int cap,
                      int32_t* npeaks, cudaStream_t st);
int launch_peaks_mask(const uint64_t* rec, int items, int n_frames, float* mask, cudaStream_t st);
int launch_landmark_hashes(const uint64_t* rec, int items, int n_frames, const mfpa_afp_params& p,
                           int sorted, int32_t* hashes, int cap, int32_t* nh, cudaStream_t st);
int launch_merge_shifts(const int32_t* hashes, const int32_t* nh, int B, int shifts, int cap_in,
                        int n_frames, int32_t* out, int cap_out, int32_t* nout, cudaStream_t st);
int launch_compact_rows(const int32_t* rows_in, const int32_t* n, int items, int cap, int64_t* offsets,
                        int32_t* rows, int64_t rows_cap, cudaStream_t st);
int launch_augment(mfpa_ctx* ctx, const float* x, int B, int T, int64_t x_stride, int sample_rate,
                   const mfpa_aug_params* params_host, const float* ir, int ir_stride, const float* noise,
                   float* out, bool final_norm, cudaStream_t st, const int64_t* ir_offsets = nullptr, int64_t ir_bank_len = 0,
                   const struct LpfShape* lpf = nullptr);
// Overrides the low-pass stage's filter per query (mfpa_lowpass_filters): cut-off and window width as fractions of the
// sample rate - julius.LowPassFilters gives every filter of a bank the window of the LOWEST cut-off.
struct LpfShape { const double* cutoff; const double* width; };
int launch_noise_assemble(mfpa_ctx* ctx, const float* bank, int64_t bank_len, const mfpa_noise_piece* pieces, int n_pieces,
                          int B, int T, float* out, cudaStream_t st, bool pieces_pinned = false);
int launch_match_counts(mfpa_ctx* ctx, const int32_t* hashes, const int32_t* nh, int B, int cap, int32_t* counts,
                        cudaStream_t st);
bool match_fused_ok(const mfpa_ctx* ctx);
int launch_match_fused(mfpa_ctx* ctx, const int32_t* hashes, const int32_t* nh, int B, int cap, int threshcount,
                       int search_depth, int32_t* cand, int32_t* ncand, uint32_t* list, int list_cap, int32_t* nlist,
                       cudaStream_t st);
bool match_sparse_ok(const mfpa_ctx* ctx);
int launch_match_emit_peer(mfpa_ctx* ctx, const int32_t* hashes, const int32_t* nh, int B, int cap, const mfpa_peer_set* peers,
                           int words_cap, cudaStream_t st);
int launch_peer_barrier(const mfpa_peer_set* peers, uint32_t epoch, cudaStream_t st);
int launch_match_emit(mfpa_ctx* ctx, const int32_t* hashes, const int32_t* nh, int B, int cap, uint32_t* words, int words_cap,
                      int32_t* nwords, cudaStream_t st);
int launch_match_owner(mfpa_ctx* ctx, const uint32_t* words, const int32_t* nwords, int n_shards, int B, int words_cap,
                       int threshcount, int search_depth, int32_t* cand, int32_t* ncand, uint32_t* list, int list_cap,
                       int32_t* nlist, cudaStream_t st);
int launch_match_select(mfpa_ctx* ctx, const int32_t* counts, int B, int threshcount, int search_depth, int32_t* cand,
                        int32_t* ncand, cudaStream_t st);
int launch_match_collect(mfpa_ctx* ctx, const int32_t* hashes, const int32_t* nh, int B, int cap, const int32_t* cand,
                         const int32_t* ncand, int search_depth, uint32_t* list, int list_cap, int32_t* nlist,
                         cudaStream_t st);
int launch_match_align(const uint32_t* lists, const int32_t* nlists, int n_lists, int B, int list_cap, const int32_t* cand,
                       const int32_t* ncand, int search_depth, int window, int threshcount, int max_align,
                       int32_t* results, int32_t* nrows, int max_rows, cudaStream_t st);
int launch_get_hits(mfpa_ctx* ctx, const int32_t* hashes, int n, int32_t* hits, int64_t hits_cap, int64_t* nhits,
                    cudaStream_t st);
int launch_dejavu_peaks(const void* arr, int is_f64, int B, int F, int N, int r, double amp_min, uint8_t* mask,
                        int32_t* peaks, int cap, int32_t* npeaks, cudaStream_t st);
int launch_mask_metrics(const float* pred, const float* gt, int B, int H, int W, double* out4, cudaStream_t st);
int launch_psnr_stats(const double* pred, const double* target, int64_t n, double* out3, cudaStream_t st);
int stft_init_tables(mfpa_ctx* ctx);
// MFPA_OPT_STAGE_TIMES: stage_begin opens the slot of one chain / fingerprint call, stage_mark(k) stamps the START of
// stage k (MFPA_STAGE_*) on the stream the kernels run on; a no-op unless the option is set
// Small host->device uploads from PINNED host memory done by a kernel (the SMs read the mapped host buffer) instead of
// cudaMemcpyAsync: a DMA copy queues on the one host->device copy engine BEHIND the next chunk's 640 MB query copy of
// the host pipelines, which delayed every chunk's kernels by a whole chunk copy (measured: 65 -> 51 ms per 10 k queries).
int launch_pull(mfpa_ctx* ctx, void* dst_dev, const void* src_pinned, size_t bytes, cudaStream_t st);
int stage_begin(mfpa_ctx* ctx);
void stage_mark(mfpa_ctx* ctx, int stage, cudaStream_t st);

}  // namespace mfpa
