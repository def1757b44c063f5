// S3 — audfprint peak picking: one warp per (query, shift) item, 8 frequency
// bins per lane, time frames scanned sequentially.
// Replaces Audfprint_peaks.find_peaks after the STFT (afp/audfprint/peak_extractor.py:263-311):
//   sgram /= max; log(max(., max/1e6)); -= mean            :263-276
//   lfilter([1,-1],[1,-0.98]) per bin, Nyquist row dropped   :286-290
//   _decaying_threshold_fwd_prune                            :173-204
//   _decaying_threshold_bwd_prune_peaks                      :206-234
// All arithmetic is float64 with explicitly un-fused mul/add so every comparison
// sees the value numpy/scipy would have produced; the only inexact pieces are
// the device log() (<= 1 ulp) and the summation order of the mean.
//
// Data flow per item:
//   phase 0 (only when no qmax is supplied) max over the spectrogram
//   phase 1 mean of log: product of clamped values with exponent bookkeeping,
//           one log per lane (more accurate than summing 64 k rounded logs)
//   phase 2 forward scan: IIR state, threshold and current frame live in
//           registers; kept peaks (<= 5/frame, sorted by (value, bin) desc) go
//           to a small per-item scratch list
//   phase 3 backward scan over the scratch list only,
//           writing one packed 64-bit record per frame.
#include "common.cuh"

namespace mfpa {

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr unsigned kFull = 0xffffffffu;
constexpr int kGPad = kSpreadLen + (kSpreadLen >> 3) + 1;  // padded Gaussian table in smem

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(kFull, v, src); }

template <typename T> struct Loader;
template <> struct Loader<float> {
  static __device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
};

// log() of a positive normal float as float64, accurate to ~1e-15 absolute:
// 128-entry table on the top mantissa bits (1/c, log c), log1p polynomial on the
// remainder, exponent * ln2 from a second table.  Used on the fused float32-STFT
// path only; the float64 parity path calls the CUDA math library log().
struct LogTables {
  double2 c[128];   // {1/c_i, log(c_i)}, c_i = 1 + (i + 0.5)/128
  double e[256];    // (biased exponent - 127) * ln2
};
__device__ __forceinline__ double log_fast(float v, const LogTables& t) {
  const unsigned u = __float_as_uint(v);
  const unsigned mant = u & 0x7fffffu;
  const double m = __hiloint2double((int)(0x3ff00000u | (mant >> 3)), (int)(mant << 29));
  const double2 ci = t.c[mant >> 16];
  const double r = fma(m, ci.x, -1.0);  // |r| <= 2^-8
  double q = fma(r, 0.2, -0.25);
  q = fma(q, r, 1.0 / 3.0);
  q = fma(q, r, -0.5);
  q = fma(q, r, 1.0);
  return t.e[u >> 23] + fma(r, q, ci.y);
}
template <> struct Loader<double> {
  static __device__ __forceinline__ void load8(const double* p, double (&v)[8]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double2 a = __ldg(reinterpret_cast<const double2*>(p) + i);
      v[2 * i] = a.x; v[2 * i + 1] = a.y;
    }
  }
};

struct PeakArgs {
  const void* mag;        // [items][n_max][264]
  const float* qmax;      // nullable
  int items, T, shifts, n_max, n_fixed;  // n_fixed > 0: every item has n_fixed frames
  double a_dec;
  int maxpks;
  const double* spread;   // [513]
  uint64_t* rec;          // [items][n_max]
  int32_t* npeaks;        // nullable
  uint64_t* fwd_pack;     // [items][n_max]   bins (bytes 0..4) + count (byte 7)
  double* fwd_val;        // [items][n_max][5]
};

// sth[j] = max(sth[j], val * G[256 + bin - pos])      (peak_extractor.py:168-170, :198-201)
__device__ __forceinline__ void spread_one(double (&sth)[8], const double* g_s, int lane, int pos, double val) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = kRows + 8 * lane + j - pos;
    sth[j] = fmax(sth[j], __dmul_rn(val, g_s[i + (i >> 3)]));
  }
}

// locmax (peak_extractor.py:61-73) on a 256-vector held 8 bins per lane; returns an 8-bit mask.
__device__ __forceinline__ unsigned locmax8(const double (&y)[8], int lane) {
  const double left = __shfl_up_sync(kFull, y[7], 1);
  const double right = __shfl_down_sync(kFull, y[0], 1);
  unsigned m = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const bool ge = (j == 0) ? (lane == 0 ? true : (y[0] >= left)) : (y[j] >= y[j - 1]);
    const bool ge_next = (j == 7) ? (lane == 31 ? false : (right >= y[7])) : (y[j + 1] >= y[j]);
    if (ge && !ge_next) m |= 1u << j;
  }
  return m;
}

// spreadpeaksinvector (peak_extractor.py:115-125): threshold = max over local maxima of val*G, from zeros.
__device__ __forceinline__ void spread_locmax(const double (&v)[8], double (&sth)[8], const double* g_s, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j) sth[j] = 0.0;
  const unsigned lm = locmax8(v, lane);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    unsigned bal = __ballot_sync(kFull, (lm >> j) & 1u);
    while (bal) {
      const int src = __ffs(bal) - 1;
      bal &= bal - 1;
      spread_one(sth, g_s, lane, 8 * src + j, shfl_d(v[j], src));
    }
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// ascending sort of up to 5 bins packed one per byte (dead entries = 0xFF), returns packed record.
__device__ __forceinline__ uint64_t finalize_record(uint64_t pack, int n, unsigned alive) {
  int b[5];
  int cnt = 0;
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    const bool ok = i < n && ((alive >> i) & 1u);
    b[i] = ok ? (int)((pack >> (8 * i)) & 0xff) : 0x100;
    cnt += ok;
  }
#define MFPA_CSWAP(i, j) { const int lo = min(b[i], b[j]), hi = max(b[i], b[j]); b[i] = lo; b[j] = hi; }
  MFPA_CSWAP(0, 1) MFPA_CSWAP(3, 4) MFPA_CSWAP(2, 4) MFPA_CSWAP(2, 3) MFPA_CSWAP(0, 3)
  MFPA_CSWAP(0, 2) MFPA_CSWAP(1, 4) MFPA_CSWAP(1, 3) MFPA_CSWAP(1, 2)
#undef MFPA_CSWAP
  uint64_t r = (uint64_t)cnt;
#pragma unroll
  for (int i = 0; i < 5; ++i)
    if (i < cnt) r |= (uint64_t)(b[i] & 0xff) << (8 * (i + 1));
  return r;
}

template <typename MagT, bool kPreFiltered>
__global__ void __launch_bounds__(kWarpsPerBlock * 32) peaks_kernel(const PeakArgs a) {
  constexpr bool kFast = sizeof(MagT) == 4 && !kPreFiltered;
  __shared__ double g_s[kGPad];
  __shared__ LogTables lt;
  for (int i = threadIdx.x; i < kSpreadLen; i += blockDim.x) g_s[i + (i >> 3)] = a.spread[i];
  if (kFast) {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
      const double c = 1.0 + ((double)i + 0.5) / 128.0;
      lt.c[i] = make_double2(1.0 / c, log(c));
    }
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lt.e[i] = (double)(i - 127) * 0.69314718055994530942;
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (item >= a.items) return;

  int n_frames = a.n_fixed;
  if (n_frames <= 0) {
    const int sh = item % a.shifts;
    const int off = a.shifts < 2 ? 0 : (int)((double)sh / (double)a.shifts * (double)kHop);
    n_frames = 1 + (a.T - off) / kHop;
  }
  const MagT* base = reinterpret_cast<const MagT*>(a.mag) + (int64_t)item * a.n_max * kPitch;
  const MagT* col0 = base + 8 * lane;
  uint64_t* rec = a.rec + (int64_t)item * a.n_max;
  uint64_t* fwd_pack = a.fwd_pack + (int64_t)item * a.n_max;
  double* fwd_val = a.fwd_val + (int64_t)item * a.n_max * kMaxPks;

  double M = 1.0, mean = 0.0;
  float floor_f = 0.f;   // kFast: clamp level 1e-6 * M in float32
  double c0 = 0.0;       // kFast: log(M) + mean
  bool silent = false, fast = false;
  if (!kPreFiltered) {
    // ---- phase 0: divisor of `sgram /= np.max(sgram)` (:263) ----
    if (a.qmax) {
      M = (double)a.qmax[item];
    } else {
      double m = 0.0;
      for (int c = 0; c < n_frames; ++c) {
        MagT v[8];
        Loader<MagT>::load8(col0 + (int64_t)c * kPitch, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) m = fmax(m, (double)v[j]);
        if (lane == 0) m = fmax(m, (double)base[(int64_t)c * kPitch + kRows]);
      }
      M = warp_max(m);
    }
    // All-zero (or NaN) input: the reference divides 0/0, skips the log and ends with no peaks (:272-280).
    silent = !(M > 0.0);
    fast = kFast && M > 1e-30 && M < 1e30;  // else (absurd scale) take the exact float64 route
    if (!silent) {
      // ---- phase 1: mean over all 257 x N of log(max(v/M, 1e-6)) (:275-276) ----
      // Sum of logs = log of product: multiply the clamped values, peel the exponent off
      // after every few factors, take one log per lane at the end.
      int esum = 0;
      double lane_log;
      if (kFast && fast) {
        const float rMf = 1.0f / (float)M;
        float P = 1.0f;
        for (int c = 0; c < n_frames; ++c) {
          float v[8];
          Loader<float>::load8(reinterpret_cast<const float*>(col0) + (int64_t)c * kPitch, v);
          const float ny = lane == 0 ? (float)base[(int64_t)c * kPitch + kRows] * rMf : 1.0f;
          const float p0 = fmaxf(v[0] * rMf, 1e-6f) * fmaxf(v[1] * rMf, 1e-6f) * fmaxf(v[2] * rMf, 1e-6f) * fmaxf(v[3] * rMf, 1e-6f);
          const float p1 = fmaxf(v[4] * rMf, 1e-6f) * fmaxf(v[5] * rMf, 1e-6f) * fmaxf(v[6] * rMf, 1e-6f) * fmaxf(v[7] * rMf, 1e-6f);
          P *= p0;  // [1,2) * [1e-24, ~1]
          unsigned u = __float_as_uint(P);
          esum += (int)(u >> 23) - 127;
          P = __uint_as_float((u & 0x007fffffu) | 0x3f800000u) * (p1 * fmaxf(ny, 1e-6f));  // >= 1e-30
          u = __float_as_uint(P);
          esum += (int)(u >> 23) - 127;
          P = __uint_as_float((u & 0x007fffffu) | 0x3f800000u);
        }
        lane_log = log((double)P);
      } else {
        const double rM = 1.0 / M;
        double P = 1.0;
        for (int c = 0; c < n_frames; ++c) {
          MagT v[8];
          Loader<MagT>::load8(col0 + (int64_t)c * kPitch, v);
          double pr = 1.0;
#pragma unroll
          for (int j = 0; j < 8; ++j) pr *= fmax((double)v[j] * rM, 1e-6);
          if (lane == 0) pr *= fmax((double)base[(int64_t)c * kPitch + kRows] * rM, 1e-6);
          P *= pr;  // P in [1,2) * [1e-54, 1]: no underflow
          const int hi = __double2hiint(P);
          esum += ((hi >> 20) & 0x7ff) - 1023;
          P = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(P));
        }
        lane_log = log(P);
      }
      const double total = warp_sum((double)esum * 0.69314718055994530942 + lane_log);
      mean = total / ((double)kBins * (double)n_frames);
      if (kFast && fast) {
        floor_f = 1e-6f * (float)M;
        c0 = log(M) + mean;
      }
    }
  }

  int total_peaks = 0;
  if (!silent) {
    double y[8], z[8], sth[8];
    // x -> y for one frame
    auto advance = [&](const MagT (&v)[8]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (kPreFiltered) {
          y[j] = (double)v[j];
        } else {
          double x;
          if (kFast && fast) {
            // log(max(v/M, 1e-6)) - mean = log(max(v, 1e-6 M)) - (log M + mean)
            x = __dsub_rn(log_fast(fmaxf((float)v[j], floor_f), lt), c0);
          } else {
            const double d = fmax((double)v[j] / M, 1e-6);   // :263, :275
            x = __dsub_rn(log(d), mean);                     // :276
          }
          y[j] = __dadd_rn(x, z[j]);                         // lfilter DF-II transposed (:286-288)
          z[j] = __dadd_rn(-x, __dmul_rn(0.98, y[j]));
        }
      }
    };
    // ---- initial threshold from the first min(10, N) frames (:180-182) ----
    {
      double m[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = 0.0;
      const int n0 = n_frames < 10 ? n_frames : 10;
      for (int c = 0; c < n0; ++c) {
        MagT v[8];
        Loader<MagT>::load8(col0 + (int64_t)c * kPitch, v);
        advance(v);
#pragma unroll
        for (int j = 0; j < 8; ++j) m[j] = c == 0 ? y[j] : fmax(m[j], y[j]);
      }
      spread_locmax(m, sth, g_s, lane);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = 0.0;
    }
    // ---- phase 2: forward prune (:190-203) ----
    MagT vn[8];
    Loader<MagT>::load8(col0, vn);
    for (int c = 0; c < n_frames; ++c) {
      MagT v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = vn[j];
      if (c + 1 < n_frames) Loader<MagT>::load8(col0 + (int64_t)(c + 1) * kPitch, vn);
      advance(v);
      unsigned cand = locmax8(y, lane);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (!(y[j] > sth[j])) cand &= ~(1u << j);
      uint64_t pack = 0;
      int nk = 0;
      if (__ballot_sync(kFull, cand != 0)) {
        // take candidates in (value, bin) descending order, at most maxpks (:196-197)
        while (nk < a.maxpks) {
          double bv = 0.0;
          int bb = -1;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (((cand >> j) & 1u) && (bb < 0 || y[j] >= bv)) { bv = y[j]; bb = 8 * lane + j; }
#pragma unroll
          for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(kFull, bv, o);
            const int ob = __shfl_xor_sync(kFull, bb, o);
            if (ob >= 0 && (bb < 0 || ov > bv || (ov == bv && ob > bb))) { bv = ov; bb = ob; }
          }
          if (bb < 0) break;
          if ((bb >> 3) == lane) cand &= ~(1u << (bb & 7));
          spread_one(sth, g_s, lane, bb, bv);
          if (lane == 0) fwd_val[(int64_t)c * kMaxPks + nk] = bv;
          pack |= (uint64_t)bb << (8 * nk);
          ++nk;
        }
      }
      if (lane == 0) fwd_pack[c] = pack | ((uint64_t)nk << 56);
#pragma unroll
      for (int j = 0; j < 8; ++j) sth[j] = __dmul_rn(sth[j], a.a_dec);  // :203
    }
    __syncwarp();
    // ---- phase 3: backward prune (:217-233); y[] still holds the last frame ----
    spread_locmax(y, sth, g_s, lane);
    uint64_t prev_pack = 0;
    int prev_n = 0;
    unsigned prev_alive = 0;
    for (int c = n_frames - 1; c >= 0; --c) {
      const uint64_t fp = __ldcg(fwd_pack + c);
      const int n = (int)(fp >> 56);
      uint64_t cur_pack = 0;
      int cur_n = 0;
      for (int i = 0; i < n; ++i) {
        const int b = (int)((fp >> (8 * i)) & 0xff);
        const double val = __ldcg(fwd_val + (int64_t)c * kMaxPks + i);
        double mine = sth[0];
#pragma unroll
        for (int j = 1; j < 8; ++j) mine = ((b & 7) == j) ? sth[j] : mine;
        const double t = shfl_d(mine, b >> 3);
        if (val >= t) {
          spread_one(sth, g_s, lane, b, val);
          cur_pack |= (uint64_t)b << (8 * cur_n);
          ++cur_n;
#pragma unroll
          for (int k = 0; k < 5; ++k)  // delete a following peak in the same bin (:228-229)
            if (k < prev_n && (int)((prev_pack >> (8 * k)) & 0xff) == b) prev_alive &= ~(1u << k);
        }
      }
      if (c + 1 < n_frames) {
        const uint64_t r = finalize_record(prev_pack, prev_n, prev_alive);
        total_peaks += (int)(r & 0xff);
        if (lane == 0) rec[c + 1] = r;
      }
      prev_pack = cur_pack; prev_n = cur_n; prev_alive = 0x1f;
#pragma unroll
      for (int j = 0; j < 8; ++j) sth[j] = __dmul_rn(a.a_dec, sth[j]);  // :233
    }
    const uint64_t r = finalize_record(prev_pack, prev_n, prev_alive);
    total_peaks += (int)(r & 0xff);
    if (lane == 0) rec[0] = r;
  }
  for (int c = (silent ? 0 : n_frames) + lane; c < a.n_max; c += 32) rec[c] = 0;
  if (lane == 0 && a.npeaks) a.npeaks[item] = total_peaks;
}

// ---------------------------------------------------------------------------------------
// Fast float32 picker for the fused path (mag produced by stft_mag_kernel, itself float32).
// The float64 kernel above stays the parity entry ("bit-exact when fed the reference
// spectrogram"); end to end the float32 STFT already differs from the reference's float64
// one by ~1e-7 relative, and a float32 picker changes no hash over 516 seeded queries
// ("precision of the fused path"), so the hot path does not pay for float64.
//
// One warp per item.  Lane l owns bins l, l+32, ..., l+224 (interleaved: the Gaussian
// table reads of a spread are conflict-free with immediate offsets, loads are 128-byte
// coalesced).  Everything is in log2 units (the picker is invariant to the log base).
//   pass 1   mean of log2(max(v, 1e-6 M))
//   pass 2   forward scan; kept peaks of frame c (<= 5, sorted by (value, bin) desc by a
//            lane-parallel rank) go to a per-warp shared-memory list
//   pass 3   backward scan over that list only; peaks that pass are appended to P[c]
//   pass 4   lane-parallel finalise: record[f] = sorted(P[f] \ P[f-1])   (:227-229)
constexpr int kFW = 4;          // warps (items) per block
constexpr int kRawCap = 128;    // local maxima per frame <= 128
constexpr int kGF = 516;        // float Gaussian table, padded to a multiple of 4

struct FastArgs {
  const float* mag;     // [items][n_max][264]
  const float* qmax;    // [items]
  int items, T, shifts, n_max;
  float a_dec;
  int maxpks;
  const double* spread;  // [513]
  uint64_t* rec;         // [items][n_max]
  int32_t* npeaks;       // nullable
};

__device__ __forceinline__ float lg2_ftz(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void load_frame(const float* p, float (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __ldg(p + 32 * j);
}

// sth[j] = max(sth[j], val * G[256 + bin - pos]) for this lane's bins (bin = lane + 32 j)
__device__ __forceinline__ void spread_f(float (&sth)[8], const float* g_s, int lane, int pos, float val) {
  const float* gp = g_s + (kRows + lane - pos);
#pragma unroll
  for (int j = 0; j < 8; ++j) sth[j] = fmaxf(sth[j], val * gp[32 * j]);
}

// locmax (:61-73) over the interleaved layout; bit j of the result = bin lane + 32 j.
__device__ __forceinline__ unsigned locmax_f(const float (&y)[8], int lane) {
  unsigned ge = 0;
  float prev_rot = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float rot = __shfl_sync(kFull, y[j], (lane + 31) & 31);  // bin - 1 for lanes > 0
    const float left = lane == 0 ? prev_rot : rot;                 // lane 0: bin 32 j - 1 = lane 31 of j - 1
    const bool g = (j == 0 && lane == 0) ? true : (y[j] >= left);
    ge |= (g ? 1u : 0u) << j;
    prev_rot = rot;
  }
  const unsigned nb = __shfl_sync(kFull, ge, (lane + 1) & 31);     // ge of bin + 1
  const unsigned ge_next = lane == 31 ? (nb >> 1) : nb;            // bin 255 has no right neighbour
  return ge & ~ge_next & 0xffu;
}

__device__ __forceinline__ void spread_locmax_f(const float (&v)[8], float (&sth)[8], const float* g_s, int lane) {
#pragma unroll
  for (int j = 0; j < 8; ++j) sth[j] = 0.f;
  const unsigned lm = locmax_f(v, lane);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    unsigned bal = __ballot_sync(kFull, ((lm >> j) & 1u) && v[j] > 0.f);  // <= 0 cannot raise a zero base
    while (bal) {
      const int src = __ffs(bal) - 1;
      bal &= bal - 1;
      spread_f(sth, g_s, lane, src + 32 * j, __shfl_sync(kFull, v[j], src));
    }
  }
}

template <bool kPre>
__device__ __forceinline__ void fast_item(const FastArgs& a, int item, int lane, int n_frames, float M,
                                          const float* g_s, float* ybuf, float* pv, uint8_t* pb, uint8_t* pc,
                                          uint8_t* raw) {
  const float* base = a.mag + (int64_t)item * a.n_max * kPitch;
  const float* col = base + lane;
  // kPre: absurdly small scale; bring the data into the normal range first (exact power of two)
  const float pre = kPre ? 1.8446744e19f : 1.0f;  // 2^64
  const float floor_v = 1e-6f * (M * pre);
  // ---- pass 1: mean over 257 x N of log2(max(v, floor)) (:275-276) ----
  // stft_mag_kernel leaves, in elements 258/259 of every even frame row, the sum of log2 and the
  // minimum of the magnitudes of that frame pair.  Where nothing lies below the floor the clamp is
  // the identity and the pair's stored sum is its contribution.
  double acc = 0.0;
  // exact contribution of one frame: this lane's eight bins (+ the Nyquist bin on lane 0), clamped at the floor
  auto frame_sum = [&](int c) -> float {
    float v[8];
    load_frame(col + (int64_t)c * kPitch, v);
    float l[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) l[j] = lg2_ftz(fmaxf(kPre ? v[j] * pre : v[j], floor_v));
    float s = ((l[0] + l[1]) + (l[2] + l[3])) + ((l[4] + l[5]) + (l[6] + l[7]));
    if (lane == 0) {
      const float ny = __ldg(base + (int64_t)c * kPitch + kRows);
      s += lg2_ftz(fmaxf(kPre ? ny * pre : ny, floor_v));
    }
    return s;
  };
  if (!kPre) {
    // Frame pairs with nothing below the floor (their minimum says so) contribute their stored sum; the others -
    // a handful per degraded query: the DC bin of a few frames after the high-pass filters, digital silence - are
    // summed exactly, by the whole warp, two frames at a time.  (Scanning the whole item instead whenever one bin
    // was clamped read every degraded query's spectrogram twice: 4.9 GB instead of 2.7 GB per 10 k queries.)
    for (int cb = 0; cb < n_frames; cb += 64) {
      const int c = cb + 2 * lane;
      bool redo = false;
      if (c < n_frames) {
        const float2 st = __ldg(reinterpret_cast<const float2*>(base + (int64_t)c * kPitch + kBins + 1));
        redo = !(st.y >= floor_v) || !(fabsf(st.x) < 1e30f);   // something is clamped (or -inf / NaN)
        if (!redo) acc += (double)st.x;
      }
      unsigned m = __ballot_sync(kFull, redo);
      while (m) {
        const int cc = cb + 2 * (__ffs(m) - 1);
        m &= m - 1;
        acc += (double)frame_sum(cc);
        if (cc + 1 < n_frames) acc += (double)frame_sum(cc + 1);
      }
    }
  } else {
    for (int c = 0; c < n_frames; ++c) acc += (double)frame_sum(c);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  const float c0 = (float)(acc / ((double)kBins * (double)n_frames));

  float y[8], z[8], sth[8];
  auto advance = [&](const float (&v)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float x = lg2_ftz(fmaxf(kPre ? v[j] * pre : v[j], floor_v)) - c0;
      y[j] = x + z[j];                 // lfilter([1,-1],[1,-0.98]) DF-II transposed (:286-288)
      z[j] = fmaf(0.98f, y[j], -x);
    }
  };
  // ---- initial threshold from the first min(10, N) frames (:180-182) ----
  {
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = 0.f;
    const int n0 = n_frames < 10 ? n_frames : 10;
    for (int c = 0; c < n0; ++c) {
      float v[8];
      load_frame(col + (int64_t)c * kPitch, v);
      advance(v);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = c == 0 ? y[j] : fmaxf(m[j], y[j]);
    }
    spread_locmax_f(m, sth, g_s, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = 0.f;
  }
  // ---- pass 2: forward prune (:190-203) ----
  const unsigned lt_mask = (1u << lane) - 1u;
  float v1[8], v2[8];
  load_frame(col, v1);
  load_frame(col + (int64_t)(n_frames > 1 ? 1 : 0) * kPitch, v2);
  for (int c = 0; c < n_frames; ++c) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = v1[j]; v1[j] = v2[j]; }
    if (c + 2 < n_frames) load_frame(col + (int64_t)(c + 2) * kPitch, v2);
    advance(v);
    unsigned cand = locmax_f(y, lane);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(y[j] > sth[j])) cand &= ~(1u << j);
    const int cnt = __popc(cand);
    const unsigned has = __ballot_sync(kFull, cnt != 0);
    int nk = 0;
    if (has) {
      int excl, total;
      if (!__any_sync(kFull, cnt > 1)) {
        excl = __popc(has & lt_mask);
        total = __popc(has);
      } else {
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(kFull, incl, o);
          if (lane >= o) incl += t;
        }
        excl = incl - cnt;
        total = __shfl_sync(kFull, incl, 31);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) ybuf[lane + 32 * j] = y[j];
      for (unsigned m = cand; m; m &= m - 1) raw[excl++] = (uint8_t)(lane + 32 * (__ffs(m) - 1));
      __syncwarp();
      // rank by (value, bin) descending; the first maxpks ranks are kept, already in order (:196-197)
      for (int e = lane; e < total; e += 32) {
        const int b = raw[e];
        const float val = ybuf[b];
        int rank = 0;
        for (int k = 0; k < total; ++k) {
          const int bk = raw[k];
          const float vk = ybuf[bk];
          rank += (vk > val || (vk == val && bk > b)) ? 1 : 0;
        }
        if (rank < a.maxpks) { pv[c * kMaxPks + rank] = val; pb[c * kMaxPks + rank] = (uint8_t)b; }
      }
      nk = total < a.maxpks ? total : a.maxpks;
      __syncwarp();
      for (int i = 0; i < nk; ++i) spread_f(sth, g_s, lane, pb[c * kMaxPks + i], pv[c * kMaxPks + i]);
    }
    if (lane == 0) pc[c] = (uint8_t)nk;
#pragma unroll
    for (int j = 0; j < 8; ++j) sth[j] *= a.a_dec;  // :203
  }
  __syncwarp();
  // ---- pass 3: backward prune (:217-233); y[] still holds the last frame ----
  spread_locmax_f(y, sth, g_s, lane);
  for (int c = n_frames - 1; c >= 0; --c) {
    const int n = pc[c];
    int kept = 0;
    for (int i = 0; i < n; ++i) {
      const int b = pb[c * kMaxPks + i];
      const float val = pv[c * kMaxPks + i];
      const int jj = b >> 5;
      float mine = sth[0];
#pragma unroll
      for (int j = 1; j < 8; ++j) mine = (jj == j) ? sth[j] : mine;
      const float t = __shfl_sync(kFull, mine, b & 31);
      if (val >= t) {
        spread_f(sth, g_s, lane, b, val);
        if (lane == 0) pb[c * kMaxPks + kept] = (uint8_t)b;  // P[c], compacted in place (kept <= i)
        ++kept;
      }
    }
    if (lane == 0) pc[c] = (uint8_t)kept;
#pragma unroll
    for (int j = 0; j < 8; ++j) sth[j] *= a.a_dec;  // :233
  }
  __syncwarp();
  // ---- pass 4: record[f] = sorted(P[f] \ P[f-1]), one frame per lane ----
  uint64_t* rec = a.rec + (int64_t)item * a.n_max;
  int total_peaks = 0;
  for (int f0 = 0; f0 < a.n_max; f0 += 32) {
    const int f = f0 + lane;
    uint64_t r = 0;
    if (f < n_frames) {
      const int n = pc[f];
      const int np = f > 0 ? pc[f - 1] : 0;
      int b[5];
      int cntb = 0;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        int bi = 0x100;
        if (i < n) {
          bi = pb[f * kMaxPks + i];
          for (int k = 0; k < np; ++k)
            if (pb[(f - 1) * kMaxPks + k] == bi) bi = 0x100;  // cleared by the same bin one frame earlier
        }
        b[i] = bi;
        cntb += bi < 0x100;
      }
#define MFPA_CSWAP(i, j) { const int lo = min(b[i], b[j]), hi = max(b[i], b[j]); b[i] = lo; b[j] = hi; }
      MFPA_CSWAP(0, 1) MFPA_CSWAP(3, 4) MFPA_CSWAP(2, 4) MFPA_CSWAP(2, 3) MFPA_CSWAP(0, 3)
      MFPA_CSWAP(0, 2) MFPA_CSWAP(1, 4) MFPA_CSWAP(1, 3) MFPA_CSWAP(1, 2)
#undef MFPA_CSWAP
      r = (uint64_t)cntb;
#pragma unroll
      for (int i = 0; i < 5; ++i)
        if (i < cntb) r |= (uint64_t)(b[i] & 0xff) << (8 * (i + 1));
      total_peaks += cntb;
    }
    if (f < a.n_max) rec[f] = r;
  }
  if (a.npeaks) {
#pragma unroll
    for (int o = 16; o; o >>= 1) total_peaks += __shfl_xor_sync(kFull, total_peaks, o);
    if (lane == 0) a.npeaks[item] = total_peaks;
  }
}

__global__ void __launch_bounds__(kFW * 32) peaks_fast_kernel(const FastArgs a) {
  extern __shared__ __align__(16) unsigned char fast_smem[];
  float* g_s = reinterpret_cast<float*>(fast_smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nm = a.n_max;
  float* ybuf = g_s + kGF + warp * kRows;
  float* pv = g_s + kGF + kFW * kRows + (size_t)warp * nm * kMaxPks;
  uint8_t* bytes = reinterpret_cast<uint8_t*>(g_s + kGF + kFW * kRows + (size_t)kFW * nm * kMaxPks);
  uint8_t* pb = bytes + (size_t)warp * nm * kMaxPks;
  uint8_t* pc = bytes + (size_t)kFW * nm * kMaxPks + (size_t)warp * nm;
  uint8_t* raw = bytes + (size_t)kFW * nm * (kMaxPks + 1) + warp * kRawCap;
  for (int i = threadIdx.x; i < kSpreadLen; i += blockDim.x) g_s[i] = (float)a.spread[i];
  __syncthreads();
  const int item = blockIdx.x * kFW + warp;
  if (item >= a.items) return;
  const int sh = item % a.shifts;
  const int off = a.shifts < 2 ? 0 : (int)((double)sh / (double)a.shifts * (double)kHop);
  const int n_frames = 1 + (a.T - off) / kHop;
  const float M = a.qmax[item];
  if (!(M > 0.f)) {  // all-zero (or NaN) input: no peaks (:272-280)
    uint64_t* rec = a.rec + (int64_t)item * a.n_max;
    for (int c = lane; c < a.n_max; c += 32) rec[c] = 0;
    if (lane == 0 && a.npeaks) a.npeaks[item] = 0;
    return;
  }
  if (M >= 1e-30f)
    fast_item<false>(a, item, lane, n_frames, M, g_s, ybuf, pv, pb, pc, raw);
  else
    fast_item<true>(a, item, lane, n_frames, M, g_s, ybuf, pv, pb, pc, raw);
}

static size_t fast_smem_bytes(int n_max) {
  return sizeof(float) * (kGF + kFW * kRows + (size_t)kFW * n_max * kMaxPks) +
         (size_t)kFW * n_max * (kMaxPks + 1) + kFW * kRawCap;
}

// [items][rows][n] float64 (reference layout) -> [items][n][264] float64 (frame-major)
__global__ void spec_to_frames_kernel(const double* __restrict__ spec, int rows, int n, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int item = blockIdx.z;
  const int c0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += 8) {
    const int b = b0 + i, c = c0 + tx;
    tile[i][tx] = (b < rows && c < n) ? spec[((int64_t)item * rows + b) * n + c] : 0.0;
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, b = b0 + tx;
    if (c < n && b < kPitch) out[((int64_t)item * n + c) * kPitch + b] = (b < rows) ? tile[tx][i] : 0.0;
  }
}

}  // namespace

static int fwd_scratch(mfpa_ctx* ctx, int items, int n_max, PeakArgs& a) {
  const size_t per_item = (size_t)n_max * (8 + 8 * kMaxPks);
  if (ctx->fwd.reserve(per_item * items)) return MFPA_ENOMEM;
  a.fwd_pack = reinterpret_cast<uint64_t*>(ctx->fwd.ptr);
  a.fwd_val = reinterpret_cast<double*>(a.fwd_pack + (size_t)items * n_max);
  return MFPA_OK;
}

int launch_peaks_f32(mfpa_ctx* ctx, const float* mag, const float* qmax, int B, int T, int shifts,
                     const mfpa_afp_params& p, uint64_t* rec, int32_t* npeaks, cudaStream_t st) {
  const int n_max_f = num_frames(T);
  const size_t fast_bytes = fast_smem_bytes(n_max_f);
  if (!ctx->opt_peaks_f64 && qmax != nullptr && fast_bytes <= 100 * 1024) {
    // float32 picker: pick lists live in shared memory, no global scratch
    FastArgs f{};
    f.mag = mag; f.qmax = qmax; f.items = B * shifts; f.T = T; f.shifts = shifts; f.n_max = n_max_f;
    f.a_dec = (float)p.a_dec; f.maxpks = p.maxpks; f.spread = ctx->spread_dev; f.rec = rec; f.npeaks = npeaks;
    MFPA_CUDA(cudaFuncSetAttribute(peaks_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_bytes));
    peaks_fast_kernel<<<(f.items + kFW - 1) / kFW, kFW * 32, fast_bytes, st>>>(f);
    MFPA_CUDA(cudaGetLastError());
    return MFPA_OK;
  }
  PeakArgs a{};
  a.mag = mag; a.qmax = qmax; a.items = B * shifts; a.T = T; a.shifts = shifts;
  a.n_max = num_frames(T); a.n_fixed = 0; a.a_dec = p.a_dec; a.maxpks = p.maxpks;
  a.spread = ctx->spread_dev; a.rec = rec; a.npeaks = npeaks;
  if (int e = fwd_scratch(ctx, a.items, a.n_max, a)) return e;
  const int blocks = (a.items + kWarpsPerBlock - 1) / kWarpsPerBlock;
  peaks_kernel<float, false><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_peaks_from_spec(mfpa_ctx* ctx, const double* spec, int items, int n_frames, int stage,
                           const mfpa_afp_params& p, uint64_t* rec, int32_t* npeaks, cudaStream_t st) {
  const int rows = stage == 0 ? kBins : kRows;
  if (ctx->spec64.reserve(sizeof(double) * (size_t)items * n_frames * kPitch)) return MFPA_ENOMEM;
  double* frames = reinterpret_cast<double*>(ctx->spec64.ptr);
  for (int i0 = 0; i0 < items; i0 += 32768) {
    const int n = items - i0 < 32768 ? items - i0 : 32768;
    dim3 grid((n_frames + 31) / 32, (kPitch + 31) / 32, n);
    spec_to_frames_kernel<<<grid, dim3(32, 8), 0, st>>>(spec + (int64_t)i0 * rows * n_frames, rows, n_frames,
                                                       frames + (int64_t)i0 * n_frames * kPitch);
    MFPA_CUDA(cudaGetLastError());
  }
  PeakArgs a{};
  a.mag = frames; a.qmax = nullptr; a.items = items; a.T = 0; a.shifts = 1;
  a.n_max = n_frames; a.n_fixed = n_frames; a.a_dec = p.a_dec; a.maxpks = p.maxpks;
  a.spread = ctx->spread_dev; a.rec = rec; a.npeaks = npeaks;
  if (int e = fwd_scratch(ctx, items, n_frames, a)) return e;
  const int blocks = (items + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (stage == 0)
    peaks_kernel<double, false><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
  else
    peaks_kernel<double, true><<<blocks, kWarpsPerBlock * 32, 0, st>>>(a);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

}  // namespace mfpa
