// extern "C" surface of libmfpa.so (see include/mfpa.h for the contract).
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "common.cuh"

namespace mfpa {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int Scratch::reserve(size_t need) {
  if (need <= bytes) return 0;
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  bytes = 0;
  // grow with 25 % head-room so alternating batch sizes do not thrash
  size_t want = need + need / 4;
  cudaError_t e = cudaMalloc(&ptr, want);
  if (e != cudaSuccess) {
    e = cudaMalloc(&ptr, need);
    want = need;
  }
  if (e != cudaSuccess) {
    ptr = nullptr;
    set_error("cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
    return MFPA_ENOMEM;
  }
  bytes = want;
  return 0;
}

void Scratch::release() {
  if (ptr) cudaFree(ptr);
  ptr = nullptr;
  bytes = 0;
}

static int check_afp(const mfpa_afp_params* p) {
  MFPA_REQUIRE(p != nullptr, "afp params are NULL");
  MFPA_REQUIRE(p->maxpks >= 1 && p->maxpks <= MFPA_MAX_PKS, "pks-per-frame %d not in 1..%d", p->maxpks, MFPA_MAX_PKS);
  MFPA_REQUIRE(p->a_dec > 0.0 && p->a_dec <= 1.0, "a_dec %g not in (0,1]", p->a_dec);
  MFPA_REQUIRE(p->mindt >= 0 && p->targetdt > p->mindt && p->targetdt <= 64, "bad mindt/targetdt %d/%d", p->mindt, p->targetdt);
  MFPA_REQUIRE(p->targetdf >= 1 && p->targetdf <= 32, "targetdf %d not in 1..32", p->targetdf);
  return MFPA_OK;
}

static int check_batch(int B, int T, int shifts) {
  MFPA_REQUIRE(B >= 1, "batch %d < 1", B);
  MFPA_REQUIRE(T >= 1, "n_samples %d < 1 (find_peaks returns an empty list for empty input; handle on the host)", T);
  MFPA_REQUIRE(shifts >= 1 && shifts <= MFPA_MAX_SHIFTS, "shifts %d not in 1..%d", shifts, MFPA_MAX_SHIFTS);
  MFPA_REQUIRE(T > shift_offset(shifts - 1, shifts), "n_samples %d shorter than the largest shift", T);
  return MFPA_OK;
}

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

}  // namespace mfpa

namespace mfpa {

template <typename W>
__global__ void __launch_bounds__(256) pull_kernel(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t bytes) {
  const size_t nw = bytes / sizeof(W);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < nw; i += (size_t)gridDim.x * 256)
    reinterpret_cast<W*>(dst)[i] = reinterpret_cast<const W*>(src)[i];
  for (size_t i = nw * sizeof(W) + (size_t)blockIdx.x * 256 + threadIdx.x; i < bytes; i += (size_t)gridDim.x * 256) dst[i] = src[i];
}

int launch_pull(mfpa_ctx* ctx, void* dst_dev, const void* src_pinned, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return MFPA_OK;
  // Only the chunked host pipelines need it (their big query copies occupy the copy engine); everywhere else the plain
  // DMA copy is used.  MFPA_NO_PULL forces the DMA copy there too: Nsight Compute hangs on kernels that read mapped
  // host memory, so profiling runs of the host path set it.
  static const bool no_pull = getenv("MFPA_NO_PULL") != nullptr;
  if (no_pull || !ctx->in_host_pipeline) {
    MFPA_CUDA(cudaMemcpyAsync(dst_dev, src_pinned, bytes, cudaMemcpyHostToDevice, st));
    return MFPA_OK;
  }
  const uintptr_t al = (uintptr_t)dst_dev | (uintptr_t)src_pinned;
  const unsigned blocks = (unsigned)((bytes / 16 + 255) / 256 < 64 ? (bytes / 16 + 255) / 256 : 64) + 1;
  if ((al & 15) == 0) pull_kernel<uint4><<<blocks, 256, 0, st>>>((unsigned char*)dst_dev, (const unsigned char*)src_pinned, bytes);
  else if ((al & 3) == 0) pull_kernel<uint32_t><<<blocks, 256, 0, st>>>((unsigned char*)dst_dev, (const unsigned char*)src_pinned, bytes);
  else pull_kernel<unsigned char><<<blocks, 256, 0, st>>>((unsigned char*)dst_dev, (const unsigned char*)src_pinned, bytes);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int stage_begin(mfpa_ctx* ctx) {
  if (!ctx->opt_stage_times) return MFPA_OK;
  ctx->stage_slot = ctx->stage_calls % 16;
  ++ctx->stage_calls;
  ctx->stage_marked[ctx->stage_slot] = 0;
  for (int k = 0; k <= MFPA_N_STAGES; ++k) {
    cudaEvent_t& e = ctx->stage_ev[ctx->stage_slot][k];
    if (!e) MFPA_CUDA(cudaEventCreate(&e));
  }
  return MFPA_OK;
}

void stage_mark(mfpa_ctx* ctx, int stage, cudaStream_t st) {
  if (ctx->opt_stage_times && ctx->stage_calls > 0) {
    cudaEventRecord(ctx->stage_ev[ctx->stage_slot][stage], st);
    ctx->stage_marked[ctx->stage_slot] |= 1u << stage;
  }
}

}  // namespace mfpa

using namespace mfpa;

extern "C" {

int mfpa_stage_times(mfpa_ctx* ctx, float* ms_out) {
  MFPA_REQUIRE(ctx && ms_out, "stage_times: NULL argument");
  DeviceGuard guard(ctx->device);
  for (int k = 0; k < MFPA_N_STAGES; ++k) ms_out[k] = 0.f;
  const int n = ctx->stage_calls < 16 ? ctx->stage_calls : 16;
  for (int s = 0; s < n; ++s) {
    cudaEvent_t* e = ctx->stage_ev[s];
    const unsigned marked = ctx->stage_marked[s];
    if (!(marked >> MFPA_N_STAGES & 1u)) continue;   // the call did not reach its end mark
    MFPA_CUDA(cudaEventSynchronize(e[MFPA_N_STAGES]));
    for (int k = 0; k < MFPA_N_STAGES; ++k) {
      if (!(marked >> k & 1u)) continue;             // stage k did not run in this call
      int j = k + 1;
      while (!(marked >> j & 1u)) ++j;               // next stamped stage (the end mark at the latest)
      float ms = 0.f;
      MFPA_CUDA(cudaEventElapsedTime(&ms, e[k], e[j]));
      ms_out[k] += ms / n;
    }
  }
  return n;
}

int mfpa_abi_version(void) { return MFPA_ABI_VERSION; }
const char* mfpa_last_error(void) { return g_err; }
int mfpa_num_frames(int n_samples) { return num_frames(n_samples); }
int mfpa_shift_offset(int shift, int shifts) { return shift_offset(shift, shifts); }

void mfpa_afp_defaults(mfpa_afp_params* p) {
  // testing/parameters.py:17-26 + Audfprint_peaks.__init__ (peak_extractor.py:99-108)
  p->a_dec = 1.0 - 0.01 * (20.0 * sqrt(256.0 / 352.8) / 35.0);
  p->f_sd = 30.0;
  p->maxpks = 5;
  p->mindt = 2;
  p->targetdt = 63;
  p->targetdf = 31;
  p->fanout = 3;
  p->reserved = 0;
}

int mfpa_create(mfpa_ctx** out, int device) {
  MFPA_REQUIRE(out != nullptr, "mfpa_create: out is NULL");
  *out = nullptr;
  int n = 0;
  MFPA_CUDA(cudaGetDeviceCount(&n));
  MFPA_REQUIRE(device >= 0 && device < n, "mfpa_create: device %d not in [0,%d)", device, n);
  DeviceGuard guard(device);
  cudaDeviceProp prop;
  MFPA_CUDA(cudaGetDeviceProperties(&prop, device));
  MFPA_REQUIRE(prop.major == 10, "libmfpa is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  mfpa_ctx* ctx = new (std::nothrow) mfpa_ctx();
  if (!ctx) { set_error("out of host memory"); return MFPA_ENOMEM; }
  ctx->device = device;
  ctx->num_sms = prop.multiProcessorCount;
  if (const char* e = getenv("MFPA_CONV_OCC")) { if (e[0] == '2') ctx->opt_conv_occ = 2; }
  MFPA_CUDA(cudaMalloc(&ctx->spread_dev, sizeof(double) * kSpreadLen));
  double tab[kSpreadLen];
  for (int k = -kRows; k <= kRows; ++k) { const double u = (double)k / 30.0; tab[k + kRows] = exp(-0.5 * (u * u)); }
  MFPA_CUDA(cudaMemcpy(ctx->spread_dev, tab, sizeof(tab), cudaMemcpyHostToDevice));
  if (int e = stft_init_tables(ctx)) { mfpa_destroy(ctx); return e; }
  MFPA_CUDA(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
  MFPA_CUDA(cudaStreamCreateWithFlags(&ctx->s_run, cudaStreamNonBlocking));
  MFPA_CUDA(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    MFPA_CUDA(cudaEventCreateWithFlags(&ctx->ev_in[b], cudaEventDisableTiming));
    MFPA_CUDA(cudaEventCreateWithFlags(&ctx->ev_run[b], cudaEventDisableTiming));
    MFPA_CUDA(cudaEventCreateWithFlags(&ctx->ev_out[b], cudaEventDisableTiming));
  }
  *out = ctx;
  return MFPA_OK;
}

void mfpa_destroy(mfpa_ctx* ctx) {
  if (!ctx) return;
  DeviceGuard guard(ctx->device);
  cudaDeviceSynchronize();
  Scratch* all[] = {&ctx->mag, &ctx->qmax, &ctx->rec, &ctx->fwd, &ctx->hashes, &ctx->nh, &ctx->misc, &ctx->spec64,
                    &ctx->xin, &ctx->out_h, &ctx->out_n, &ctx->aug_a, &ctx->aug_b, &ctx->aug_c, &ctx->aug_d,
                    &ctx->aug_small, &ctx->aug_lists, &ctx->aug_long, &ctx->aug_part, &ctx->aug_noise, &ctx->aug_pool, &ctx->match_a, &ctx->match_b, &ctx->match_c};
  for (Scratch* s : all) s->release();
  for (int b = 0; b < 2; ++b) {
    ctx->h_x[b].release(); ctx->h_x16[b].release(); ctx->h_rows[b].release(); ctx->h_csr[b].release(); ctx->h_n[b].release(); ctx->h_off[b].release(); ctx->h_noise[b].release();
    if (ctx->pieces_pinned[b]) cudaFreeHost(ctx->pieces_pinned[b]);
    if (ctx->ev_in[b]) cudaEventDestroy(ctx->ev_in[b]);
    if (ctx->ev_run[b]) cudaEventDestroy(ctx->ev_run[b]);
    if (ctx->ev_out[b]) cudaEventDestroy(ctx->ev_out[b]);
  }
  for (int sl = 0; sl < 16; ++sl)
    for (int k = 0; k <= MFPA_N_STAGES; ++k)
      if (ctx->stage_ev[sl][k]) cudaEventDestroy(ctx->stage_ev[sl][k]);
  if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
  if (ctx->s_run) cudaStreamDestroy(ctx->s_run);
  if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
  if (ctx->spread_dev) cudaFree(ctx->spread_dev);
  if (ctx->tw_dev) cudaFree(ctx->tw_dev);
  if (ctx->win_dev) cudaFree(ctx->win_dev);
  if (ctx->win_dejavu_dev) cudaFree(ctx->win_dejavu_dev);
  if (ctx->index_table) cudaFree(ctx->index_table);
  if (ctx->index_counts) cudaFree(ctx->index_counts);
  if (ctx->index_hashesperid) cudaFree(ctx->index_hashesperid);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->aug_pinned) cudaFreeHost(ctx->aug_pinned);
  if (ctx->aug_copy_done) cudaEventDestroy((cudaEvent_t)ctx->aug_copy_done);
  if (ctx->aug_tw_dev) cudaFree(ctx->aug_tw_dev);
  delete ctx;
}

int mfpa_set_spread_table(mfpa_ctx* ctx, const double* table513_host) {
  MFPA_REQUIRE(ctx && table513_host, "set_spread_table: NULL argument");
  DeviceGuard guard(ctx->device);
  MFPA_CUDA(cudaMemcpy(ctx->spread_dev, table513_host, sizeof(double) * kSpreadLen, cudaMemcpyHostToDevice));
  return MFPA_OK;
}

int mfpa_set_option(mfpa_ctx* ctx, int option, int value) {
  MFPA_REQUIRE(ctx != nullptr, "set_option: ctx is NULL");
  switch (option) {
    case MFPA_OPT_PEAKS_F64: ctx->opt_peaks_f64 = value != 0; return MFPA_OK;
    case MFPA_OPT_MATCH_PACKED: ctx->opt_match_packed = value != 0; return MFPA_OK;
    case MFPA_OPT_MATCH_UNFUSED: ctx->opt_match_unfused = value == 2 ? 2 : (value != 0); return MFPA_OK;
    case MFPA_OPT_PART_BUDGET_MB:
      MFPA_REQUIRE(value >= 1, "set_option: MFPA_OPT_PART_BUDGET_MB %d < 1", value);
      ctx->opt_part_budget_mb = value;
      return MFPA_OK;
    case MFPA_OPT_CONV_OCC:
      MFPA_REQUIRE(value == 2 || value == 3, "set_option: MFPA_OPT_CONV_OCC %d not in {2, 3}", value);
      ctx->opt_conv_occ = value;
      return MFPA_OK;
    case MFPA_OPT_CLIP_POOLED: ctx->opt_clip_pooled = value != 0; return MFPA_OK;
    case MFPA_OPT_STAGE_TIMES:
      ctx->opt_stage_times = value != 0;
      ctx->stage_calls = 0;
      return MFPA_OK;
    default: break;
  }
  set_error("set_option: unknown option %d", option);
  return MFPA_EINVAL;
}

int mfpa_stft_mag(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, int shifts,
                  float* mag_dev, float* qmax_dev, void* stream) {
  MFPA_REQUIRE(ctx && x_dev && mag_dev && qmax_dev, "stft_mag: NULL argument");
  if (int e = check_batch(B, T, shifts)) return e;
  MFPA_REQUIRE(x_stride >= T, "stft_mag: row stride %lld < n_samples %d", (long long)x_stride, T);
  DeviceGuard guard(ctx->device);
  return launch_stft_mag(ctx, x_dev, B, T, x_stride, shifts, mag_dev, qmax_dev, (cudaStream_t)stream);
}

int mfpa_stft_num_frames(int n_samples, int n_fft, int hop, int win_len) {
  if (n_samples < 1 || n_fft < 2 || hop < 1 || win_len < 1) return 0;
  const int64_t padded = (int64_t)n_samples + 2 * (n_fft / 2);
  if (padded < win_len) return 0;
  return (int)(1 + (padded - win_len) / hop);
}

int mfpa_stft_complex(mfpa_ctx* ctx, const double* x_dev, int T, int n_fft, int hop, const double* window_dev,
                      int win_len, double* out_dev, void* stream) {
  MFPA_REQUIRE(ctx && x_dev && window_dev && out_dev, "stft_complex: NULL argument");
  MFPA_REQUIRE(T >= 1 && n_fft >= 2 && n_fft <= 8192 && hop >= 1 && win_len >= 1, "stft_complex: T %d, n_fft %d, hop %d, window %d",
               T, n_fft, hop, win_len);
  const int n_frames = mfpa_stft_num_frames(T, n_fft, hop, win_len);
  MFPA_REQUIRE(n_frames >= 1, "stft_complex: signal of %d samples is shorter than one window of %d", T, win_len);
  DeviceGuard guard(ctx->device);
  return launch_stft_complex(x_dev, T, n_fft, hop, window_dev, win_len, n_frames, out_dev, (cudaStream_t)stream);
}

int mfpa_spec_from_mag(mfpa_ctx* ctx, const float* mag_dev, const float* qmax_dev, int B, int T,
                       int shifts, double* spec_dev, void* stream) {
  MFPA_REQUIRE(ctx && mag_dev && qmax_dev && spec_dev, "spec_from_mag: NULL argument");
  if (int e = check_batch(B, T, shifts)) return e;
  DeviceGuard guard(ctx->device);
  return launch_spec_from_mag(mag_dev, qmax_dev, B, T, shifts, spec_dev, (cudaStream_t)stream);
}

int mfpa_audfprint_peaks(mfpa_ctx* ctx, const float* mag_dev, const float* qmax_dev, int B, int T,
                         int shifts, const mfpa_afp_params* p, uint64_t* rec_dev, int32_t* npeaks_dev,
                         void* stream) {
  MFPA_REQUIRE(ctx && mag_dev && rec_dev, "audfprint_peaks: NULL argument");
  if (int e = check_batch(B, T, shifts)) return e;
  if (int e = check_afp(p)) return e;
  MFPA_REQUIRE(((uintptr_t)mag_dev & 15) == 0, "audfprint_peaks: mag must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  return launch_peaks_f32(ctx, mag_dev, qmax_dev, B, T, shifts, *p, rec_dev, npeaks_dev, (cudaStream_t)stream);
}

int mfpa_audfprint_peaks_from_spec(mfpa_ctx* ctx, const double* spec_dev, int items, int n_frames,
                                   int stage, const mfpa_afp_params* p, uint64_t* rec_dev,
                                   int32_t* npeaks_dev, void* stream) {
  MFPA_REQUIRE(ctx && spec_dev && rec_dev, "peaks_from_spec: NULL argument");
  MFPA_REQUIRE(items >= 1 && n_frames >= 1, "peaks_from_spec: items %d, n_frames %d", items, n_frames);
  MFPA_REQUIRE(stage == 0 || stage == 1, "peaks_from_spec: stage %d not in {0,1}", stage);
  if (int e = check_afp(p)) return e;
  DeviceGuard guard(ctx->device);
  return launch_peaks_from_spec(ctx, spec_dev, items, n_frames, stage, *p, rec_dev, npeaks_dev, (cudaStream_t)stream);
}

int mfpa_peaks_list(mfpa_ctx* ctx, const uint64_t* rec_dev, int items, int n_frames, int32_t* peaks_dev,
                    int cap, int32_t* npeaks_dev, void* stream) {
  MFPA_REQUIRE(ctx && rec_dev && peaks_dev, "peaks_list: NULL argument");
  MFPA_REQUIRE(items >= 1 && n_frames >= 1 && cap >= 1, "peaks_list: bad sizes");
  DeviceGuard guard(ctx->device);
  return launch_peaks_list(rec_dev, items, n_frames, peaks_dev, cap, npeaks_dev, (cudaStream_t)stream);
}

int mfpa_peaks_mask(mfpa_ctx* ctx, const uint64_t* rec_dev, int items, int n_frames, float* mask_dev,
                    void* stream) {
  MFPA_REQUIRE(ctx && rec_dev && mask_dev, "peaks_mask: NULL argument");
  MFPA_REQUIRE(items >= 1 && n_frames >= 1, "peaks_mask: bad sizes");
  DeviceGuard guard(ctx->device);
  return launch_peaks_mask(rec_dev, items, n_frames, mask_dev, (cudaStream_t)stream);
}

int mfpa_landmark_hashes(mfpa_ctx* ctx, const uint64_t* rec_dev, int items, int n_frames,
                         const mfpa_afp_params* p, int sorted, int32_t* hashes_dev, int cap,
                         int32_t* nh_dev, void* stream) {
  MFPA_REQUIRE(ctx && rec_dev && hashes_dev && nh_dev, "landmark_hashes: NULL argument");
  MFPA_REQUIRE(items >= 1 && n_frames >= 1 && cap >= 1, "landmark_hashes: bad sizes");
  if (int e = check_afp(p)) return e;
  DeviceGuard guard(ctx->device);
  return launch_landmark_hashes(rec_dev, items, n_frames, *p, sorted, hashes_dev, cap, nh_dev, (cudaStream_t)stream);
}

int mfpa_merge_shifts(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B, int shifts,
                      int cap_in, int n_frames, int32_t* out_dev, int cap_out, int32_t* nout_dev,
                      void* stream) {
  MFPA_REQUIRE(ctx && hashes_dev && nh_dev && out_dev && nout_dev, "merge_shifts: NULL argument");
  MFPA_REQUIRE(B >= 1 && cap_in >= 1 && cap_out >= 1 && n_frames >= 1, "merge_shifts: bad sizes");
  DeviceGuard guard(ctx->device);
  return launch_merge_shifts(hashes_dev, nh_dev, B, shifts, cap_in, n_frames, out_dev, cap_out, nout_dev,
                             (cudaStream_t)stream);
}

int mfpa_fingerprint(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, int shifts,
                     const mfpa_afp_params* p, int32_t* hashes_dev, int cap, int32_t* nh_dev, void* stream) {
  MFPA_REQUIRE(ctx && x_dev && hashes_dev && nh_dev, "fingerprint: NULL argument");
  if (int e = check_batch(B, T, shifts)) return e;
  if (int e = check_afp(p)) return e;
  MFPA_REQUIRE(x_stride >= T, "fingerprint: row stride %lld < n_samples %d", (long long)x_stride, T);
  MFPA_REQUIRE(cap >= 1, "fingerprint: cap %d < 1", cap);
  DeviceGuard guard(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int items = B * shifts;
  const int n_max = num_frames(T);
  if (ctx->mag.reserve(sizeof(float) * (size_t)items * n_max * kPitch)) return MFPA_ENOMEM;
  if (ctx->qmax.reserve(sizeof(float) * items)) return MFPA_ENOMEM;
  if (ctx->rec.reserve(sizeof(uint64_t) * (size_t)items * n_max)) return MFPA_ENOMEM;
  float* mag = (float*)ctx->mag.ptr;
  float* qmax = (float*)ctx->qmax.ptr;
  uint64_t* rec = (uint64_t*)ctx->rec.ptr;
  if (!ctx->stage_chained) { if (int e = stage_begin(ctx)) return e; }
  ctx->stage_chained = false;
  stage_mark(ctx, MFPA_STAGE_STFT, st);
  if (int e = launch_stft_mag(ctx, x_dev, B, T, x_stride, shifts, mag, qmax, st)) return e;
  stage_mark(ctx, MFPA_STAGE_PEAKS, st);
  if (int e = launch_peaks_f32(ctx, mag, qmax, B, T, shifts, *p, rec, nullptr, st)) return e;
  stage_mark(ctx, MFPA_STAGE_LANDMARKS, st);
  if (shifts == 1) {
    const int e = launch_landmark_hashes(rec, items, n_max, *p, 1, hashes_dev, cap, nh_dev, st);
    stage_mark(ctx, MFPA_N_STAGES, st);
    return e;
  }
  const int cap_in = MFPA_HASHES_PER_FRAME * n_max;
  if (ctx->hashes.reserve(sizeof(int32_t) * 2 * (size_t)items * cap_in)) return MFPA_ENOMEM;
  if (ctx->nh.reserve(sizeof(int32_t) * items)) return MFPA_ENOMEM;
  if (int e = launch_landmark_hashes(rec, items, n_max, *p, 1, (int32_t*)ctx->hashes.ptr, cap_in,
                                     (int32_t*)ctx->nh.ptr, st)) return e;
  const int e = launch_merge_shifts((int32_t*)ctx->hashes.ptr, (int32_t*)ctx->nh.ptr, B, shifts, cap_in, n_max,
                                    hashes_dev, cap, nh_dev, st);
  stage_mark(ctx, MFPA_N_STAGES, st);
  return e;
}

int mfpa_compact_rows(mfpa_ctx* ctx, const int32_t* rows_in_dev, const int32_t* n_dev, int items, int cap,
                      int64_t* offsets_dev, int32_t* rows_dev, int64_t rows_cap, void* stream) {
  MFPA_REQUIRE(ctx && rows_in_dev && n_dev && offsets_dev && rows_dev, "compact_rows: NULL argument");
  MFPA_REQUIRE(items >= 1 && cap >= 1 && rows_cap >= 0, "compact_rows: bad sizes");
  DeviceGuard guard(ctx->device);
  return launch_compact_rows(rows_in_dev, n_dev, items, cap, offsets_dev, rows_dev, rows_cap, (cudaStream_t)stream);
}

// Chunked, double-buffered host path: H2D of chunk i+1 (s_in) overlaps the kernels of
// chunk i (s_run); compacted rows leave on s_out.  Scratch used by the kernels is only
// touched from s_run, so chunks never race on it.
__global__ void __launch_bounds__(256) pcm16_to_f32_kernel(const int16_t* __restrict__ in, float* __restrict__ out, int64_t n) {
  // 8 samples per thread: one 16-byte load, two 16-byte stores; x / 32768 is exact in float32
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const uint4 v = *reinterpret_cast<const uint4*>(in + i);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    float f[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[2 * k] = (float)(int16_t)(w[k] & 0xffffu) * (1.0f / 32768.0f);
      f[2 * k + 1] = (float)(int16_t)(w[k] >> 16) * (1.0f / 32768.0f);
    }
    *reinterpret_cast<float4*>(out + i) = make_float4(f[0], f[1], f[2], f[3]);
    *reinterpret_cast<float4*>(out + i + 4) = make_float4(f[4], f[5], f[6], f[7]);
  } else {
    for (int64_t k = i; k < n; ++k) out[k] = (float)in[k] * (1.0f / 32768.0f);
  }
}

static int fingerprint_host_impl(mfpa_ctx* ctx, const void* x_host_v, bool pcm16, int B, int T, int shifts,
                                 const mfpa_afp_params* p, int32_t* rows_host, int64_t rows_cap,
                                 int64_t* offsets_host) {
  const float* x_host = (const float*)x_host_v;
  const int16_t* x_host16 = (const int16_t*)x_host_v;
  MFPA_REQUIRE(ctx && x_host_v && rows_host && offsets_host, "fingerprint_host: NULL argument");
  if (int e = check_batch(B, T, shifts)) return e;
  if (int e = check_afp(p)) return e;
  MFPA_REQUIRE(rows_cap >= 0, "fingerprint_host: rows_cap < 0");
  DeviceGuard guard(ctx->device);
  const int n_max = num_frames(T);
  const int cap = MFPA_HASHES_PER_FRAME * n_max * shifts;
  int chunk = (int)((int64_t)(256 << 20) / ((int64_t)T * (int64_t)sizeof(float)));  // ~256 MiB of samples
  if (chunk < 1) chunk = 1;
  if (chunk > B) chunk = B;
  for (int b = 0; b < 2; ++b) {
    if (ctx->h_x[b].reserve(sizeof(float) * (size_t)chunk * T)) return MFPA_ENOMEM;
    if (pcm16 && ctx->h_x16[b].reserve(sizeof(int16_t) * (size_t)chunk * T + 16)) return MFPA_ENOMEM;
    if (ctx->h_rows[b].reserve(sizeof(int32_t) * 2 * (size_t)chunk * cap)) return MFPA_ENOMEM;
    if (ctx->h_csr[b].reserve(sizeof(int32_t) * 2 * (size_t)chunk * cap)) return MFPA_ENOMEM;
    if (ctx->h_n[b].reserve(sizeof(int32_t) * chunk)) return MFPA_ENOMEM;
    if (ctx->h_off[b].reserve(sizeof(int64_t) * ((size_t)chunk + 1))) return MFPA_ENOMEM;
  }
  if (ctx->pinned_bytes < 2 * sizeof(int64_t) * ((size_t)chunk + 1)) {
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 2 * sizeof(int64_t) * ((size_t)chunk + 1);
    MFPA_CUDA(cudaMallocHost(&ctx->pinned, ctx->pinned_bytes));
  }
  int64_t* off_pinned[2] = {(int64_t*)ctx->pinned, (int64_t*)ctx->pinned + chunk + 1};
  const int n_chunks = (B + chunk - 1) / chunk;
  int64_t total = 0;
  bool overflow = false;
  offsets_host[0] = 0;
  auto issue = [&](int ci) -> int {
    const int b = ci & 1, q0 = ci * chunk, nq = (B - q0 < chunk) ? (B - q0) : chunk;
    // chunk ci-2 used these buffers: its kernels and its copy-out must be done
    if (ci >= 2) {
      MFPA_CUDA(cudaStreamWaitEvent(ctx->s_in, ctx->ev_run[b], 0));
    }
    if (pcm16) {
      const int64_t n = (int64_t)nq * T;
      MFPA_CUDA(cudaMemcpyAsync(ctx->h_x16[b].ptr, x_host16 + (size_t)q0 * T, sizeof(int16_t) * (size_t)n,
                                cudaMemcpyHostToDevice, ctx->s_in));
      pcm16_to_f32_kernel<<<(unsigned)((n / 8 + 256) / 256), 256, 0, ctx->s_in>>>((const int16_t*)ctx->h_x16[b].ptr,
                                                                                 (float*)ctx->h_x[b].ptr, n);
      MFPA_CUDA(cudaGetLastError());
    } else {
      MFPA_CUDA(cudaMemcpyAsync(ctx->h_x[b].ptr, x_host + (size_t)q0 * T, sizeof(float) * (size_t)nq * T,
                                cudaMemcpyHostToDevice, ctx->s_in));
    }
    MFPA_CUDA(cudaEventRecord(ctx->ev_in[b], ctx->s_in));
    MFPA_CUDA(cudaStreamWaitEvent(ctx->s_run, ctx->ev_in[b], 0));
    if (ci >= 2) MFPA_CUDA(cudaStreamWaitEvent(ctx->s_run, ctx->ev_out[b], 0));
    if (int e = mfpa_fingerprint(ctx, (const float*)ctx->h_x[b].ptr, nq, T, T, shifts, p, (int32_t*)ctx->h_rows[b].ptr,
                                 cap, (int32_t*)ctx->h_n[b].ptr, ctx->s_run)) return e;
    if (int e = launch_compact_rows((int32_t*)ctx->h_rows[b].ptr, (int32_t*)ctx->h_n[b].ptr, nq, cap,
                                    (int64_t*)ctx->h_off[b].ptr, (int32_t*)ctx->h_csr[b].ptr, (int64_t)nq * cap,
                                    ctx->s_run)) return e;
    MFPA_CUDA(cudaMemcpyAsync(off_pinned[b], ctx->h_off[b].ptr, sizeof(int64_t) * (nq + 1), cudaMemcpyDeviceToHost,
                              ctx->s_run));
    MFPA_CUDA(cudaEventRecord(ctx->ev_run[b], ctx->s_run));
    return MFPA_OK;
  };
  if (int e = issue(0)) return e;
  for (int ci = 0; ci < n_chunks; ++ci) {
    const int b = ci & 1, q0 = ci * chunk, nq = (B - q0 < chunk) ? (B - q0) : chunk;
    if (ci + 1 < n_chunks)
      if (int e = issue(ci + 1)) return e;
    MFPA_CUDA(cudaEventSynchronize(ctx->ev_run[b]));
    const int64_t n_rows = off_pinned[b][nq];
    for (int i = 1; i <= nq; ++i) offsets_host[q0 + i] = total + off_pinned[b][i];
    if (total + n_rows > rows_cap) overflow = true;
    if (!overflow && n_rows > 0) {
      MFPA_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->ev_run[b], 0));
      MFPA_CUDA(cudaMemcpyAsync(rows_host + 2 * total, ctx->h_csr[b].ptr, sizeof(int32_t) * 2 * (size_t)n_rows,
                                cudaMemcpyDeviceToHost, ctx->s_out));
    }
    MFPA_CUDA(cudaEventRecord(ctx->ev_out[b], ctx->s_out));
    total += n_rows;
  }
  MFPA_CUDA(cudaStreamSynchronize(ctx->s_out));
  MFPA_CUDA(cudaStreamSynchronize(ctx->s_run));
  if (overflow) {
    set_error("fingerprint_host: %lld rows produced, capacity %lld", (long long)total, (long long)rows_cap);
    return MFPA_ECAP;
  }
  return MFPA_OK;
}

int mfpa_fingerprint_host(mfpa_ctx* ctx, const float* x_host, int B, int T, int shifts,
                          const mfpa_afp_params* p, int32_t* rows_host, int64_t rows_cap,
                          int64_t* offsets_host) {
  return fingerprint_host_impl(ctx, x_host, false, B, T, shifts, p, rows_host, rows_cap, offsets_host);
}

int mfpa_fingerprint_host_pcm16(mfpa_ctx* ctx, const int16_t* x_host, int B, int T, int shifts,
                                const mfpa_afp_params* p, int32_t* rows_host, int64_t rows_cap,
                                int64_t* offsets_host) {
  return fingerprint_host_impl(ctx, x_host, true, B, T, shifts, p, rows_host, rows_cap, offsets_host);
}

static int check_augment(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, int sample_rate,
                         const mfpa_aug_params* params_host) {
  MFPA_REQUIRE(ctx && x_dev && params_host, "augment: NULL argument");
  MFPA_REQUIRE(B >= 1 && T >= 2, "augment: batch %d, n_samples %d", B, T);
  MFPA_REQUIRE(B <= 65535, "augment: batch %d > 65535 (one grid row per query); split the batch", B);
  MFPA_REQUIRE(x_stride >= T, "augment: row stride %lld < n_samples %d", (long long)x_stride, T);
  MFPA_REQUIRE(sample_rate > 0, "augment: sample_rate %d", sample_rate);
  return MFPA_OK;
}

int mfpa_augment(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, int sample_rate,
                 const mfpa_aug_params* params_host, const float* ir_dev, int ir_stride,
                 const float* noise_dev, float* out_dev, void* stream) {
  if (int e = check_augment(ctx, x_dev, B, T, x_stride, sample_rate, params_host)) return e;
  MFPA_REQUIRE(out_dev != nullptr, "augment: out_dev is NULL");
  DeviceGuard guard(ctx->device);
  return launch_augment(ctx, x_dev, B, T, x_stride, sample_rate, params_host, ir_dev, ir_stride, noise_dev, out_dev,
                        true, (cudaStream_t)stream);
}

int mfpa_lowpass_filters(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, const double* cutoff_host,
                         const double* width_host, float* out_dev, void* stream) {
  MFPA_REQUIRE(ctx && x_dev && cutoff_host && width_host && out_dev && B >= 1 && T >= 1 && x_stride >= T,
               "lowpass_filters: bad argument");
  DeviceGuard guard(ctx->device);
  std::vector<mfpa_aug_params> pp((size_t)B);
  for (int i = 0; i < B; ++i) {
    pp[i] = mfpa_aug_params{};
    pp[i].apply = MFPA_AUG_LPF;
    pp[i].fc2_hz = (float)cutoff_host[i];   // for messages only: the shape below is what the stage uses
    pp[i].gain_factor = 1.0f;
  }
  const LpfShape shape{cutoff_host, width_host};
  return launch_augment(ctx, x_dev, B, T, x_stride, 1, pp.data(), nullptr, 0, nullptr, out_dev, false, (cudaStream_t)stream,
                        nullptr, 0, &shape);
}

int mfpa_noise_assemble(mfpa_ctx* ctx, const float* bank_dev, int64_t bank_len, const mfpa_noise_piece* pieces_host,
                        int n_pieces, int B, int T, float* out_dev, void* stream) {
  MFPA_REQUIRE(ctx && bank_dev && pieces_host && out_dev, "noise_assemble: NULL argument");
  MFPA_REQUIRE(bank_len >= 1 && n_pieces >= 1 && B >= 1 && T >= 1, "noise_assemble: bad sizes");
  DeviceGuard guard(ctx->device);
  return launch_noise_assemble(ctx, bank_dev, bank_len, pieces_host, n_pieces, B, T, out_dev, (cudaStream_t)stream);
}

int mfpa_augment_fingerprint(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, int sample_rate,
                             const mfpa_aug_params* params_host, const float* ir_dev, int ir_stride,
                             const float* noise_dev, int shifts, const mfpa_afp_params* p,
                             int32_t* hashes_dev, int cap, int32_t* nh_dev, void* stream) {
  if (int e = check_augment(ctx, x_dev, B, T, x_stride, sample_rate, params_host)) return e;
  DeviceGuard guard(ctx->device);
  // PeakNormalization (stage 7) is skipped: the picker divides the magnitudes by their maximum and
  // works on log differences, so a positive scale of the waveform cancels; the fingerprint stages read
  // stage 6's output where launch_augment leaves it (ctx->aug_a) and no copy of the waveform is made.
  if (int e = stage_begin(ctx)) return e;
  if (int e = launch_augment(ctx, x_dev, B, T, x_stride, sample_rate, params_host, ir_dev, ir_stride, noise_dev, nullptr,
                             false, (cudaStream_t)stream)) return e;
  ctx->stage_chained = true;   // the analysis continues this call's record
  return mfpa_fingerprint(ctx, (const float*)ctx->aug_a.ptr, B, T, T, shifts, p, hashes_dev, cap, nh_dev, stream);
}

// Host-buffer form of mfpa_augment_fingerprint: the same chunked pipeline as fingerprint_host_impl, with the
// degradation chain ahead of the analysis.  Only the queries stream from the host; impulse responses and noise
// are device-resident (rows, or a bank plus per-query offsets / random_background pieces assembled per chunk).
int mfpa_augment_fingerprint_host(mfpa_ctx* ctx, const mfpa_chain_inputs* in, int B, int T, int sample_rate,
                                  const mfpa_aug_params* params_host, int shifts, const mfpa_afp_params* p,
                                  int32_t* rows_host, int64_t rows_cap, int64_t* offsets_host) {
  MFPA_REQUIRE(ctx && in && params_host && rows_host && offsets_host, "augment_fingerprint_host: NULL argument");
  MFPA_REQUIRE((in->x_host != nullptr) != (in->x_pcm16_host != nullptr), "augment_fingerprint_host: give x_host or x_pcm16_host");
  MFPA_REQUIRE(!(in->noise_dev && in->pieces_host), "augment_fingerprint_host: give noise rows or noise pieces, not both");
  MFPA_REQUIRE(!in->pieces_host || (in->noise_bank_dev && in->noise_bank_len >= 1 && in->n_pieces >= B),
               "augment_fingerprint_host: noise pieces need a bank and at least one piece per query");
  if (int e = check_batch(B, T, shifts)) return e;
  if (int e = check_afp(p)) return e;
  MFPA_REQUIRE(T >= 2 && sample_rate > 0 && rows_cap >= 0, "augment_fingerprint_host: n_samples %d, sample_rate %d", T, sample_rate);
  const bool pcm16 = in->x_pcm16_host != nullptr;
  DeviceGuard guard(ctx->device);
  const int n_max = num_frames(T);
  const int cap = MFPA_HASHES_PER_FRAME * n_max * shifts;
  // Chunks of up to ~256 MiB of samples (MFPA_CHUNK_MB overrides), the batch split evenly: a chunk's kernels (~2 us per
  // query) hide behind the next chunk's copy (~4.7 us per float32 query over PCIe 5 x16), so only the first copy and the
  // last chunk's kernels are exposed
  int64_t chunk_bytes = (int64_t)256 << 20;
  if (const char* e = getenv("MFPA_CHUNK_MB")) { const long v = atol(e); if (v > 0) chunk_bytes = (int64_t)v << 20; }
  int chunk = (int)(chunk_bytes / ((int64_t)T * (int64_t)sizeof(float)));
  if (chunk < 1) chunk = 1;
  if (chunk > B) chunk = B;
  chunk = (B + (B + chunk - 1) / chunk - 1) / ((B + chunk - 1) / chunk);
  // pieces are consumed chunk by chunk: they must come grouped by query, queries ascending
  std::vector<int> piece_begin;
  if (in->pieces_host) {
    piece_begin.assign((size_t)B + 1, 0);
    int prev = 0;
    for (int i = 0; i < in->n_pieces; ++i) {
      const int q = in->pieces_host[i].query;
      MFPA_REQUIRE(q >= prev && q < B, "augment_fingerprint_host: noise piece %d: query %d out of order or range", i, q);
      prev = q;
      piece_begin[(size_t)q + 1] = i + 1;
    }
    for (int q = 1; q <= B; ++q) if (piece_begin[q] < piece_begin[q - 1]) piece_begin[q] = piece_begin[q - 1];
  }
  for (int b = 0; b < 2; ++b) {
    if (ctx->h_x[b].reserve(sizeof(float) * (size_t)chunk * T)) return MFPA_ENOMEM;
    if (pcm16 && ctx->h_x16[b].reserve(sizeof(int16_t) * (size_t)chunk * T + 16)) return MFPA_ENOMEM;
    if (in->pieces_host && ctx->h_noise[b].reserve(sizeof(float) * (size_t)chunk * T)) return MFPA_ENOMEM;
    if (ctx->h_rows[b].reserve(sizeof(int32_t) * 2 * (size_t)chunk * cap)) return MFPA_ENOMEM;
    if (ctx->h_csr[b].reserve(sizeof(int32_t) * 2 * (size_t)chunk * cap)) return MFPA_ENOMEM;
    if (ctx->h_n[b].reserve(sizeof(int32_t) * chunk)) return MFPA_ENOMEM;
    if (ctx->h_off[b].reserve(sizeof(int64_t) * ((size_t)chunk + 1))) return MFPA_ENOMEM;
  }
  if (ctx->pinned_bytes < 2 * sizeof(int64_t) * ((size_t)chunk + 1)) {
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    ctx->pinned = nullptr;
    ctx->pinned_bytes = 2 * sizeof(int64_t) * ((size_t)chunk + 1);
    MFPA_CUDA(cudaMallocHost(&ctx->pinned, ctx->pinned_bytes));
  }
  int64_t* off_pinned[2] = {(int64_t*)ctx->pinned, (int64_t*)ctx->pinned + chunk + 1};
  const int n_chunks = (B + chunk - 1) / chunk;
  int64_t total = 0;
  bool overflow = false;
  offsets_host[0] = 0;
  // MFPA_DEBUG_PIPE=1: device timeline of the pipeline (copy-in and kernel spans per chunk) on stderr
  const bool dbg = getenv("MFPA_DEBUG_PIPE") != nullptr;
  std::vector<cudaEvent_t> dev;
  if (dbg) { dev.resize(4 * (size_t)n_chunks + 1); for (auto& e : dev) cudaEventCreate(&e); cudaEventRecord(dev[4 * n_chunks], ctx->s_in); }
  auto issue = [&](int ci) -> int {
    const int b = ci & 1, q0 = ci * chunk, nq = (B - q0 < chunk) ? (B - q0) : chunk;
    if (ci >= 2) MFPA_CUDA(cudaStreamWaitEvent(ctx->s_in, ctx->ev_run[b], 0));   // chunk ci-2 used these buffers
    if (dbg) cudaEventRecord(dev[4 * ci], ctx->s_in);
    const int64_t n = (int64_t)nq * T;
    if (pcm16) {
      MFPA_CUDA(cudaMemcpyAsync(ctx->h_x16[b].ptr, in->x_pcm16_host + (size_t)q0 * T, sizeof(int16_t) * (size_t)n,
                                cudaMemcpyHostToDevice, ctx->s_in));
      pcm16_to_f32_kernel<<<(unsigned)((n / 8 + 256) / 256), 256, 0, ctx->s_in>>>((const int16_t*)ctx->h_x16[b].ptr,
                                                                                 (float*)ctx->h_x[b].ptr, n);
      MFPA_CUDA(cudaGetLastError());
    } else {
      MFPA_CUDA(cudaMemcpyAsync(ctx->h_x[b].ptr, in->x_host + (size_t)q0 * T, sizeof(float) * (size_t)n,
                                cudaMemcpyHostToDevice, ctx->s_in));
    }
    MFPA_CUDA(cudaEventRecord(ctx->ev_in[b], ctx->s_in));
    if (dbg) cudaEventRecord(dev[4 * ci + 1], ctx->s_in);
    MFPA_CUDA(cudaStreamWaitEvent(ctx->s_run, ctx->ev_in[b], 0));
    if (ci >= 2) MFPA_CUDA(cudaStreamWaitEvent(ctx->s_run, ctx->ev_out[b], 0));
    if (dbg) cudaEventRecord(dev[4 * ci + 2], ctx->s_run);
    const float* noise = in->noise_dev ? in->noise_dev + (size_t)q0 * T : nullptr;
    if (in->pieces_host) {
      // this chunk's pieces, rows re-based to the chunk, staged in pinned memory of slot b (free again: the copy
      // that read it was ordered before ev_run[b] of chunk ci-2, which the caller has waited for)
      const int p0 = piece_begin[q0], p1 = piece_begin[(size_t)q0 + nq];
      const size_t bytes = sizeof(mfpa_noise_piece) * (size_t)(p1 - p0);
      if (ctx->pieces_pinned_bytes[b] < bytes) {
        if (ctx->pieces_pinned[b]) cudaFreeHost(ctx->pieces_pinned[b]);
        ctx->pieces_pinned[b] = nullptr;
        ctx->pieces_pinned_bytes[b] = bytes + bytes / 4 + 64;
        MFPA_CUDA(cudaMallocHost(&ctx->pieces_pinned[b], ctx->pieces_pinned_bytes[b]));
      }
      mfpa_noise_piece* pp = (mfpa_noise_piece*)ctx->pieces_pinned[b];
      for (int i = p0; i < p1; ++i) { pp[i - p0] = in->pieces_host[i]; pp[i - p0].query -= q0; }
      if (int e = launch_noise_assemble(ctx, in->noise_bank_dev, in->noise_bank_len, pp, p1 - p0, nq, T,
                                        (float*)ctx->h_noise[b].ptr, ctx->s_run, true)) return e;
      noise = (const float*)ctx->h_noise[b].ptr;
    }
    const float* ir = in->ir_dev;
    if (ir && !in->ir_offsets_host) ir += (size_t)q0 * in->ir_stride;
    if (int e = stage_begin(ctx)) return e;
    ctx->stage_chained = true;
    if (int e = launch_augment(ctx, (const float*)ctx->h_x[b].ptr, nq, T, T, sample_rate, params_host + q0, ir, in->ir_stride,
                               noise, nullptr, false, ctx->s_run, in->ir_offsets_host ? in->ir_offsets_host + q0 : nullptr,
                               in->ir_bank_len)) return e;
    if (int e = mfpa_fingerprint(ctx, (const float*)ctx->aug_a.ptr, nq, T, T, shifts, p, (int32_t*)ctx->h_rows[b].ptr, cap,
                                 (int32_t*)ctx->h_n[b].ptr, ctx->s_run)) return e;
    if (int e = launch_compact_rows((int32_t*)ctx->h_rows[b].ptr, (int32_t*)ctx->h_n[b].ptr, nq, cap,
                                    (int64_t*)ctx->h_off[b].ptr, (int32_t*)ctx->h_csr[b].ptr, (int64_t)nq * cap,
                                    ctx->s_run)) return e;
    MFPA_CUDA(cudaMemcpyAsync(off_pinned[b], ctx->h_off[b].ptr, sizeof(int64_t) * (nq + 1), cudaMemcpyDeviceToHost,
                              ctx->s_run));
    MFPA_CUDA(cudaEventRecord(ctx->ev_run[b], ctx->s_run));
    if (dbg) cudaEventRecord(dev[4 * ci + 3], ctx->s_run);
    return MFPA_OK;
  };
  struct PipelineFlag {   // small uploads inside the pipeline bypass the copy engine (launch_pull)
    mfpa_ctx* c;
    explicit PipelineFlag(mfpa_ctx* ctx) : c(ctx) { c->in_host_pipeline = true; }
    ~PipelineFlag() { c->in_host_pipeline = false; }
  } pipeline_flag(ctx);
  if (int e = issue(0)) return e;
  for (int ci = 0; ci < n_chunks; ++ci) {
    const int b = ci & 1, q0 = ci * chunk, nq = (B - q0 < chunk) ? (B - q0) : chunk;
    if (ci + 1 < n_chunks)
      if (int e = issue(ci + 1)) return e;
    MFPA_CUDA(cudaEventSynchronize(ctx->ev_run[b]));
    const int64_t n_rows = off_pinned[b][nq];
    for (int i = 1; i <= nq; ++i) offsets_host[q0 + i] = total + off_pinned[b][i];
    if (total + n_rows > rows_cap) overflow = true;
    if (!overflow && n_rows > 0) {
      MFPA_CUDA(cudaStreamWaitEvent(ctx->s_out, ctx->ev_run[b], 0));
      MFPA_CUDA(cudaMemcpyAsync(rows_host + 2 * total, ctx->h_csr[b].ptr, sizeof(int32_t) * 2 * (size_t)n_rows,
                                cudaMemcpyDeviceToHost, ctx->s_out));
    }
    MFPA_CUDA(cudaEventRecord(ctx->ev_out[b], ctx->s_out));
    total += n_rows;
  }
  MFPA_CUDA(cudaStreamSynchronize(ctx->s_out));
  MFPA_CUDA(cudaStreamSynchronize(ctx->s_run));
  if (dbg) {
    for (int ci = 0; ci < n_chunks; ++ci) {
      float t[4];
      for (int k = 0; k < 4; ++k) cudaEventElapsedTime(&t[k], dev[4 * n_chunks], dev[4 * ci + k]);
      fprintf(stderr, "pipe chunk %2d: copy %7.2f .. %7.2f ms   kernels %7.2f .. %7.2f ms\n", ci, t[0], t[1], t[2], t[3]);
    }
    for (auto& e : dev) cudaEventDestroy(e);
  }
  if (overflow) {
    set_error("augment_fingerprint_host: %lld rows produced, capacity %lld", (long long)total, (long long)rows_cap);
    return MFPA_ECAP;
  }
  return MFPA_OK;
}

void mfpa_match_defaults(mfpa_match_params* p) {
  p->window = 2; p->threshcount = 5; p->search_depth = 100; p->max_alignments_per_id = 100;
}

static int check_match(mfpa_ctx* ctx, const mfpa_match_params* p) {
  MFPA_REQUIRE(ctx != nullptr, "match: ctx is NULL");
  MFPA_REQUIRE(ctx->index_table != nullptr, "match: no index loaded (mfpa_index_load)");
  if (p) {
    MFPA_REQUIRE(p->search_depth >= 1 && p->search_depth <= 128, "match: search_depth %d not in 1..128", p->search_depth);
    MFPA_REQUIRE(p->window >= 0 && p->threshcount >= 0 && p->max_alignments_per_id >= 0, "match: negative parameter");
  }
  return MFPA_OK;
}

int mfpa_index_load(mfpa_ctx* ctx, const uint32_t* table_host, const int32_t* counts_host, int hash_lo, int n_buckets,
                    int depth, int hashbits, int maxtimebits, const uint32_t* hashesperid_host, int n_tracks) {
  MFPA_REQUIRE(ctx && table_host && counts_host && hashesperid_host, "index_load: NULL argument");
  MFPA_REQUIRE(hashbits >= 1 && hashbits <= 30 && maxtimebits >= 1 && maxtimebits <= 15, "index_load: hashbits %d / maxtimebits %d", hashbits, maxtimebits);
  MFPA_REQUIRE(hash_lo >= 0 && n_buckets >= 1 && hash_lo + (int64_t)n_buckets <= (1ll << hashbits), "index_load: bad hash range");
  MFPA_REQUIRE(depth >= 1 && n_tracks >= 1 && (int64_t)n_tracks + 1 < (1ll << (32 - maxtimebits)), "index_load: depth %d / n_tracks %d", depth, n_tracks);
  DeviceGuard guard(ctx->device);
  if (ctx->index_table) cudaFree(ctx->index_table);
  if (ctx->index_counts) cudaFree(ctx->index_counts);
  if (ctx->index_hashesperid) cudaFree(ctx->index_hashesperid);
  ctx->index_table = nullptr; ctx->index_counts = nullptr; ctx->index_hashesperid = nullptr;
  const size_t tb = sizeof(uint32_t) * (size_t)n_buckets * depth;
  MFPA_CUDA(cudaMalloc(&ctx->index_table, tb));
  MFPA_CUDA(cudaMalloc(&ctx->index_counts, sizeof(int32_t) * (size_t)n_buckets));
  MFPA_CUDA(cudaMalloc(&ctx->index_hashesperid, sizeof(uint32_t) * (size_t)n_tracks));
  MFPA_CUDA(cudaMemcpy(ctx->index_table, table_host, tb, cudaMemcpyHostToDevice));
  MFPA_CUDA(cudaMemcpy(ctx->index_counts, counts_host, sizeof(int32_t) * (size_t)n_buckets, cudaMemcpyHostToDevice));
  MFPA_CUDA(cudaMemcpy(ctx->index_hashesperid, hashesperid_host, sizeof(uint32_t) * (size_t)n_tracks, cudaMemcpyHostToDevice));
  ctx->index_hash_lo = hash_lo; ctx->index_hash_hi = hash_lo + n_buckets; ctx->index_depth = depth;
  ctx->index_ntracks = n_tracks; ctx->index_maxtimebits = maxtimebits; ctx->index_hashmask = (1 << hashbits) - 1;
  uint32_t hp_min = 0xffffffffu;
  for (int i = 0; i < n_tracks; ++i)
    if (hashesperid_host[i] != 0 && hashesperid_host[i] < hp_min) hp_min = hashesperid_host[i];
  ctx->index_hp_min = hp_min == 0xffffffffu ? 1u : hp_min;
  return MFPA_OK;
}

int mfpa_get_hits(mfpa_ctx* ctx, const int32_t* hashes_dev, int n, int32_t* hits_dev, int64_t hits_cap,
                  int64_t* nhits_dev, void* stream) {
  if (int e = check_match(ctx, nullptr)) return e;
  MFPA_REQUIRE(hashes_dev && hits_dev && nhits_dev && n >= 0 && hits_cap >= 0, "get_hits: bad argument");
  MFPA_REQUIRE(((uintptr_t)hits_dev & 15) == 0, "get_hits: hits_dev must be 16-byte aligned");
  DeviceGuard guard(ctx->device);
  return launch_get_hits(ctx, hashes_dev, n, hits_dev, hits_cap, nhits_dev, (cudaStream_t)stream);
}

int mfpa_match_counts(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B, int cap,
                      int32_t* counts_dev, void* stream) {
  if (int e = check_match(ctx, nullptr)) return e;
  MFPA_REQUIRE(hashes_dev && nh_dev && counts_dev && B >= 1 && cap >= 1, "match_counts: bad argument");
  DeviceGuard guard(ctx->device);
  return launch_match_counts(ctx, hashes_dev, nh_dev, B, cap, counts_dev, (cudaStream_t)stream);
}

int mfpa_match_select(mfpa_ctx* ctx, const int32_t* counts_dev, int B, const mfpa_match_params* p,
                      int32_t* cand_dev, int32_t* ncand_dev, void* stream) {
  if (int e = check_match(ctx, p)) return e;
  MFPA_REQUIRE(p && counts_dev && cand_dev && ncand_dev && B >= 1, "match_select: bad argument");
  DeviceGuard guard(ctx->device);
  return launch_match_select(ctx, counts_dev, B, p->threshcount, p->search_depth, cand_dev, ncand_dev, (cudaStream_t)stream);
}

int mfpa_match_collect(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B, int cap,
                       const int32_t* cand_dev, const int32_t* ncand_dev, const mfpa_match_params* p,
                       uint32_t* list_dev, int list_cap, int32_t* nlist_dev, void* stream) {
  if (int e = check_match(ctx, p)) return e;
  MFPA_REQUIRE(p && hashes_dev && nh_dev && cand_dev && ncand_dev && list_dev && nlist_dev && B >= 1 && cap >= 1 && list_cap >= 1,
               "match_collect: bad argument");
  DeviceGuard guard(ctx->device);
  return launch_match_collect(ctx, hashes_dev, nh_dev, B, cap, cand_dev, ncand_dev, p->search_depth, list_dev, list_cap,
                              nlist_dev, (cudaStream_t)stream);
}

int mfpa_match_align(mfpa_ctx* ctx, const uint32_t* lists_dev, const int32_t* nlists_dev, int n_lists, int B,
                     int list_cap, const int32_t* cand_dev, const int32_t* ncand_dev, const mfpa_match_params* p,
                     int32_t* results_dev, int32_t* nrows_dev, int max_rows, void* stream) {
  MFPA_REQUIRE(ctx && p && lists_dev && nlists_dev && cand_dev && ncand_dev && results_dev && nrows_dev, "match_align: NULL argument");
  MFPA_REQUIRE(n_lists >= 1 && B >= 1 && list_cap >= 1 && max_rows >= 1, "match_align: bad sizes");
  MFPA_REQUIRE(p->search_depth >= 1 && p->search_depth <= 128, "match_align: search_depth %d", p->search_depth);
  DeviceGuard guard(ctx->device);
  return launch_match_align(lists_dev, nlists_dev, n_lists, B, list_cap, cand_dev, ncand_dev, p->search_depth, p->window,
                            p->threshcount, p->max_alignments_per_id, results_dev, nrows_dev, max_rows, (cudaStream_t)stream);
}

int mfpa_match_emit(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B, int cap, uint32_t* words_dev,
                    int words_cap, int32_t* nwords_dev, void* stream) {
  if (int e = check_match(ctx, nullptr)) return e;
  MFPA_REQUIRE(hashes_dev && nh_dev && words_dev && nwords_dev && B >= 1 && cap >= 1 && words_cap >= 1, "match_emit: bad argument");
  MFPA_REQUIRE(match_sparse_ok(ctx), "match_emit: the (track, time skew) hit words hold indexes of up to %d tracks", 110000);
  DeviceGuard guard(ctx->device);
  return launch_match_emit(ctx, hashes_dev, nh_dev, B, cap, words_dev, words_cap, nwords_dev, (cudaStream_t)stream);
}

int mfpa_peer_alloc(mfpa_ctx* ctx, uint64_t bytes, void** dev_ptr, void* handle64) {
  MFPA_REQUIRE(ctx && dev_ptr && handle64 && bytes >= 1, "peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  DeviceGuard guard(ctx->device);
  void* p = nullptr;
  MFPA_CUDA(cudaMalloc(&p, (size_t)bytes));
  cudaIpcMemHandle_t h;
  if (cudaError_t e = cudaMemset(p, 0, (size_t)bytes); e != cudaSuccess || (e = cudaIpcGetMemHandle(&h, p)) != cudaSuccess) {
    cudaFree(p);
    MFPA_REQUIRE(false, "peer_alloc: %s", cudaGetErrorString(e));
  }
  MFPA_CUDA(cudaDeviceSynchronize());
  memcpy(handle64, &h, sizeof(h));
  *dev_ptr = p;
  return MFPA_OK;
}

int mfpa_peer_open(mfpa_ctx* ctx, const void* handle64, void** dev_ptr) {
  MFPA_REQUIRE(ctx && dev_ptr && handle64, "peer_open: bad argument");
  DeviceGuard guard(ctx->device);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  MFPA_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return MFPA_OK;
}

int mfpa_peer_close(mfpa_ctx* ctx, void* dev_ptr) {
  MFPA_REQUIRE(ctx && dev_ptr, "peer_close: bad argument");
  DeviceGuard guard(ctx->device);
  MFPA_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  return MFPA_OK;
}

int mfpa_peer_free(mfpa_ctx* ctx, void* dev_ptr) {
  MFPA_REQUIRE(ctx && dev_ptr, "peer_free: bad argument");
  DeviceGuard guard(ctx->device);
  MFPA_CUDA(cudaFree(dev_ptr));
  return MFPA_OK;
}

static int check_peers(mfpa_ctx* ctx, const mfpa_peer_set* peers, bool lists) {
  MFPA_REQUIRE(ctx && peers && peers->world >= 1 && peers->world <= MFPA_MAX_PEERS && peers->rank >= 0 && peers->rank < peers->world,
               "peer set: world must be in [1, %d] and rank in [0, world)", MFPA_MAX_PEERS);
  for (int r = 0; r < peers->world; ++r)
    MFPA_REQUIRE(peers->flags[r] && (!lists || (peers->words[r] && peers->nwords[r])), "peer set: rank %d has a NULL buffer", r);
  return MFPA_OK;
}

int mfpa_match_emit_peer(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B, int cap,
                         const mfpa_peer_set* peers, int words_cap, void* stream) {
  if (int e = check_match(ctx, nullptr)) return e;
  if (int e = check_peers(ctx, peers, true)) return e;
  MFPA_REQUIRE(hashes_dev && nh_dev && B >= 1 && cap >= 1 && words_cap >= 1 && B % peers->world == 0,
               "match_emit_peer: bad argument (the sub-batch must split evenly over the %d ranks)", peers->world);
  MFPA_REQUIRE(match_sparse_ok(ctx), "match_emit_peer: the (track, time skew) hit words hold indexes of up to %d tracks", 110000);
  DeviceGuard guard(ctx->device);
  return launch_match_emit_peer(ctx, hashes_dev, nh_dev, B, cap, peers, words_cap, (cudaStream_t)stream);
}

int mfpa_peer_barrier(mfpa_ctx* ctx, const mfpa_peer_set* peers, uint32_t epoch, void* stream) {
  if (int e = check_peers(ctx, peers, false)) return e;
  DeviceGuard guard(ctx->device);
  return launch_peer_barrier(peers, epoch, (cudaStream_t)stream);
}

int mfpa_match_owner(mfpa_ctx* ctx, const uint32_t* words_dev, const int32_t* nwords_dev, int n_shards, int B, int words_cap,
                     const mfpa_match_params* p, int32_t* results_dev, int32_t* nrows_dev, int max_rows, void* stream) {
  if (int e = check_match(ctx, p)) return e;
  MFPA_REQUIRE(p && words_dev && nwords_dev && results_dev && nrows_dev && n_shards >= 1 && B >= 1 && words_cap >= 1 && max_rows >= 1,
               "match_owner: bad argument");
  MFPA_REQUIRE(match_sparse_ok(ctx), "match_owner: the (track, time skew) hit words hold indexes of up to %d tracks", 110000);
  DeviceGuard guard(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int list_cap = 8192;
  if (ctx->match_b.reserve(sizeof(int32_t) * (size_t)B * (2 * p->search_depth + 2))) return MFPA_ENOMEM;
  if (ctx->match_c.reserve(sizeof(uint32_t) * (size_t)B * list_cap)) return MFPA_ENOMEM;
  int32_t* cand = (int32_t*)ctx->match_b.ptr;
  int32_t* ncand = cand + (size_t)B * 2 * p->search_depth;
  int32_t* nlist = ncand + B;
  uint32_t* list = (uint32_t*)ctx->match_c.ptr;
  if (int e = launch_match_owner(ctx, words_dev, nwords_dev, n_shards, B, words_cap, p->threshcount, p->search_depth, cand, ncand,
                                 list, list_cap, nlist, st)) return e;
  return launch_match_align(list, nlist, 1, B, list_cap, cand, ncand, p->search_depth, p->window, p->threshcount,
                            p->max_alignments_per_id, results_dev, nrows_dev, max_rows, st);
}

int mfpa_match(mfpa_ctx* ctx, const int32_t* hashes_dev, const int32_t* nh_dev, int B, int cap,
               const mfpa_match_params* p, int32_t* results_dev, int32_t* nrows_dev, int max_rows, void* stream) {
  if (int e = check_match(ctx, p)) return e;
  MFPA_REQUIRE(p && hashes_dev && nh_dev && results_dev && nrows_dev && B >= 1 && cap >= 1 && max_rows >= 1, "match: bad argument");
  DeviceGuard guard(ctx->device);
  cudaStream_t st = (cudaStream_t)stream;
  const int nt = ctx->index_ntracks;
  if (match_fused_ok(ctx) && !ctx->opt_match_unfused) {
    // single shard, histogram fits shared memory: counts + select + collect never leave the block
    const int list_cap = 8192;
    const int sub = B < 4096 ? B : 4096;
    if (ctx->match_b.reserve(sizeof(int32_t) * (size_t)sub * (2 * p->search_depth + 2))) return MFPA_ENOMEM;
    if (ctx->match_c.reserve(sizeof(uint32_t) * (size_t)sub * list_cap)) return MFPA_ENOMEM;
    int32_t* cand = (int32_t*)ctx->match_b.ptr;
    int32_t* ncand = cand + (size_t)sub * 2 * p->search_depth;
    int32_t* nlist = ncand + sub;
    uint32_t* list = (uint32_t*)ctx->match_c.ptr;
    for (int q0 = 0; q0 < B; q0 += sub) {
      const int nq = B - q0 < sub ? B - q0 : sub;
      const int32_t* hq = hashes_dev + (size_t)q0 * cap * 2;
      if (int e = launch_match_fused(ctx, hq, nh_dev + q0, nq, cap, p->threshcount, p->search_depth, cand, ncand, list,
                                     list_cap, nlist, st)) return e;
      if (int e = launch_match_align(list, nlist, 1, nq, list_cap, cand, ncand, p->search_depth, p->window, p->threshcount,
                                     p->max_alignments_per_id, results_dev + (size_t)q0 * max_rows * 7, nrows_dev + q0,
                                     max_rows, st)) return e;
    }
    return MFPA_OK;
  }
  int sub = (int)((int64_t)(256 << 20) / ((int64_t)nt * 4));  // ~256 MiB of dense counts per sub-batch
  if (sub < 1) sub = 1;
  if (sub > B) sub = B;
  const int list_cap = 8192;
  if (ctx->match_a.reserve(sizeof(int32_t) * (size_t)sub * nt)) return MFPA_ENOMEM;
  if (ctx->match_b.reserve(sizeof(int32_t) * (size_t)sub * (2 * p->search_depth + 2))) return MFPA_ENOMEM;
  if (ctx->match_c.reserve(sizeof(uint32_t) * (size_t)sub * list_cap)) return MFPA_ENOMEM;
  int32_t* counts = (int32_t*)ctx->match_a.ptr;
  int32_t* cand = (int32_t*)ctx->match_b.ptr;
  int32_t* ncand = cand + (size_t)sub * 2 * p->search_depth;
  int32_t* nlist = ncand + sub;
  uint32_t* list = (uint32_t*)ctx->match_c.ptr;
  for (int q0 = 0; q0 < B; q0 += sub) {
    const int nq = B - q0 < sub ? B - q0 : sub;
    const int32_t* hq = hashes_dev + (size_t)q0 * cap * 2;
    if (int e = launch_match_counts(ctx, hq, nh_dev + q0, nq, cap, counts, st)) return e;
    if (int e = launch_match_select(ctx, counts, nq, p->threshcount, p->search_depth, cand, ncand, st)) return e;
    if (int e = launch_match_collect(ctx, hq, nh_dev + q0, nq, cap, cand, ncand, p->search_depth, list, list_cap, nlist, st)) return e;
    if (int e = launch_match_align(list, nlist, 1, nq, list_cap, cand, ncand, p->search_depth, p->window, p->threshcount,
                                   p->max_alignments_per_id, results_dev + (size_t)q0 * max_rows * 7, nrows_dev + q0,
                                   max_rows, st)) return e;
  }
  return MFPA_OK;
}

int mfpa_dejavu_peaks(mfpa_ctx* ctx, const void* arr_dev, int is_f64, int B, int F, int N, int neighborhood,
                      double amp_min, uint8_t* mask_dev, int32_t* peaks_dev, int cap, int32_t* npeaks_dev, void* stream) {
  MFPA_REQUIRE(ctx && arr_dev && mask_dev, "dejavu_peaks: NULL argument");
  MFPA_REQUIRE(B >= 1 && F >= 1 && N >= 1, "dejavu_peaks: bad sizes %d x %d x %d", B, F, N);
  MFPA_REQUIRE(B <= 65535, "dejavu_peaks: batch %d > 65535 (one grid layer per spectrogram); split the batch", B);
  MFPA_REQUIRE((peaks_dev == nullptr) == (npeaks_dev == nullptr) && (peaks_dev == nullptr || cap >= 1), "dejavu_peaks: peaks/npeaks/cap mismatch");
  DeviceGuard guard(ctx->device);
  return launch_dejavu_peaks(arr_dev, is_f64, B, F, N, neighborhood, amp_min, mask_dev, peaks_dev, cap, npeaks_dev,
                             (cudaStream_t)stream);
}

int mfpa_mask_metrics(mfpa_ctx* ctx, const float* predicted_dev, const float* gt_dev, int B, int H, int W, double* out4_dev,
                      void* stream) {
  MFPA_REQUIRE(ctx && predicted_dev && gt_dev && out4_dev && B >= 1 && H >= 1 && W >= 1, "mask_metrics: bad argument");
  DeviceGuard guard(ctx->device);
  return launch_mask_metrics(predicted_dev, gt_dev, B, H, W, out4_dev, (cudaStream_t)stream);
}

int mfpa_psnr_stats(mfpa_ctx* ctx, const double* pred_dev, const double* target_dev, int64_t n, double* out3_dev, void* stream) {
  MFPA_REQUIRE(ctx && pred_dev && target_dev && out3_dev && n >= 1, "psnr_stats: bad argument");
  DeviceGuard guard(ctx->device);
  return launch_psnr_stats(pred_dev, target_dev, n, out3_dev, (cudaStream_t)stream);
}

int mfpa_dejavu_num_frames(int n_samples) { return dejavu_num_frames(n_samples); }

int mfpa_dejavu_psd(mfpa_ctx* ctx, const float* x_dev, int B, int T, int64_t x_stride, float* psd_dev, void* stream) {
  MFPA_REQUIRE(ctx && x_dev && psd_dev, "dejavu_psd: NULL argument");
  MFPA_REQUIRE(B >= 1 && T >= 1 && x_stride >= T, "dejavu_psd: batch %d, n_samples %d, stride %lld", B, T, (long long)x_stride);
  DeviceGuard guard(ctx->device);
  return launch_dejavu_psd(ctx, x_dev, B, T, x_stride, psd_dev, (cudaStream_t)stream);
}

int mfpa_dejavu_log(mfpa_ctx* ctx, const float* psd_dev, int B, int F, int N, int square, float* arr_dev, void* stream) {
  MFPA_REQUIRE(ctx && psd_dev && arr_dev, "dejavu_log: NULL argument");
  MFPA_REQUIRE(B >= 1 && F >= 1 && N >= 1, "dejavu_log: bad sizes %d x %d x %d", B, F, N);
  DeviceGuard guard(ctx->device);
  return launch_dejavu_log(psd_dev, B, F * N, square, arr_dev, (cudaStream_t)stream);
}

}  // extern "C"
