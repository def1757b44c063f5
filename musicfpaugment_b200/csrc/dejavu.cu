// Dejavu 2-D peak finder: get_2D_peaks (afp/dejavu/fingerprint.py:94-171, constants
// afp/dejavu/variables.py:18-19).
//   local_max  = maximum_filter(arr, 21x21 square, mode='reflect') == arr          (:128)
//   eroded_bg  = binary_erosion(arr == 0, 21x21, border_value=1)                   (:131-134)
//   detected   = local_max XOR eroded_bg ; keep arr > amp_min                      (:137-148)
// With a full square footprint both filters are separable, and scipy's 'reflect' boundary
// only mirrors samples that the clipped window already contains, so each is a running
// reduction over the in-image part of the window.  One block produces a 32 x 32 output
// tile from a (32+2r)^2 halo tile held in shared memory; comparisons are done in the
// input's own type so `== arr` is exact.
#include "common.cuh"

namespace mfpa {

namespace {

constexpr int kTile = 32;
constexpr int kMaxR = 10;  // PEAK_NEIGHBORHOOD_SIZE of the reference; bounds the shared-memory halo
constexpr int kHalo = kTile + 2 * kMaxR;

template <typename T>
__global__ void __launch_bounds__(256) dejavu_peaks_kernel(const T* __restrict__ arr, int F, int N, int r, T amp_min,
                                                           uint8_t* __restrict__ mask) {
  __shared__ T in[kHalo][kHalo + 1];
  __shared__ T rowmax[kHalo][kTile + 1];
  __shared__ uint8_t rowany[kHalo][kTile + 4];
  const int b = blockIdx.z, f0 = blockIdx.y * kTile, t0 = blockIdx.x * kTile;
  const T* a = arr + (int64_t)b * F * N;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int span = kTile + 2 * r;
  // halo tile; out-of-image cells are marked and skipped by the reductions
  for (int i = tid; i < span * span; i += 256) {
    const int y = i / span, x = i - y * span;
    const int f = f0 - r + y, t = t0 - r + x;
    in[y][x] = (f >= 0 && f < F && t >= 0 && t < N) ? a[(int64_t)f * N + t] : (T)0;
  }
  __syncthreads();
  // pass 1: along time
  for (int i = tid; i < span * kTile; i += 256) {
    const int y = i / kTile, x = i - y * kTile;
    const int f = f0 - r + y;
    bool have = false, any = false;
    T m = (T)0;
    if (f >= 0 && f < F) {
      for (int k = 0; k <= 2 * r; ++k) {
        const int t = t0 + x - r + k;
        if (t < 0 || t >= N) continue;
        const T v = in[y][x + k];
        m = have ? (v > m ? v : m) : v;
        have = true;
        any |= (v != (T)0);
      }
    }
    rowmax[y][x] = m;
    rowany[y][x] = (have ? 2 : 0) | (any ? 1 : 0);
  }
  __syncthreads();
  // pass 2: along frequency, then the decision
  for (int i = tid; i < kTile * kTile; i += 256) {
    const int y = i / kTile, x = i - y * kTile;
    const int f = f0 + y, t = t0 + x;
    if (f >= F || t >= N) continue;
    bool have = false, any = false;
    T m = (T)0;
    for (int k = 0; k <= 2 * r; ++k) {
      const uint8_t fl = rowany[y + k][x];
      if (!(fl & 2)) continue;
      const T v = rowmax[y + k][x];
      m = have ? (v > m ? v : m) : v;
      have = true;
      any |= (fl & 1);
    }
    const T c = in[y + r][x + r];
    const bool local_max = (m == c);
    const bool eroded_bg = !any;
    mask[((int64_t)b * F + f) * N + t] = ((local_max != eroded_bg) && c > amp_min) ? 1 : 0;
  }
}

// mask -> (freq, time) rows in row-major order (np.where order, fingerprint.py:141-152)
__global__ void __launch_bounds__(1024) mask_to_list_kernel(const uint8_t* __restrict__ mask, int F, int N,
                                                            int32_t* __restrict__ peaks, int cap, int32_t* __restrict__ npeaks) {
  __shared__ int warp_tot[32];
  __shared__ int carry;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint8_t* m = mask + (int64_t)b * F * N;
  int2* out = reinterpret_cast<int2*>(peaks) + (int64_t)b * cap;
  if (tid == 0) carry = 0;
  __syncthreads();
  const int total = F * N;
  for (int base = 0; base < total; base += 1024) {
    const int i = base + tid;
    const int v = (i < total && m[i]) ? 1 : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
      int w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
      warp_tot[lane] = wi - w;
    }
    __syncthreads();
    if (v) {
      const int pos = carry + warp_tot[warp] + __popc(bal & ((1u << lane) - 1u));
      if (pos < cap) out[pos] = make_int2(i / N, i % N);
    }
    __syncthreads();
    if (tid == 1023) carry += warp_tot[31] + __popc(bal);
    __syncthreads();
  }
  if (tid == 0) npeaks[b] = carry;
}

}  // namespace

int launch_dejavu_peaks(const void* arr, int is_f64, int B, int F, int N, int r, double amp_min, uint8_t* mask,
                        int32_t* peaks, int cap, int32_t* npeaks, cudaStream_t st) {
  MFPA_REQUIRE(r >= 1 && r <= kMaxR, "dejavu: neighbourhood %d not in 1..%d", r, kMaxR);
  dim3 grid((N + kTile - 1) / kTile, (F + kTile - 1) / kTile, B);
  MFPA_REQUIRE(B <= 65535, "dejavu: batch %d > 65535", B);
  if (is_f64)
    dejavu_peaks_kernel<double><<<grid, dim3(32, 8), 0, st>>>((const double*)arr, F, N, r, amp_min, mask);
  else
    dejavu_peaks_kernel<float><<<grid, dim3(32, 8), 0, st>>>((const float*)arr, F, N, r, (float)amp_min, mask);
  MFPA_CUDA(cudaGetLastError());
  if (peaks && npeaks) {
    mask_to_list_kernel<<<B, 1024, 0, st>>>(mask, F, N, peaks, cap, npeaks);
    MFPA_CUDA(cudaGetLastError());
  }
  return MFPA_OK;
}


// ---- fingerprint() front end (afp/dejavu/fingerprint.py:60-79) --------------------------------
// mlab.specgram's segments (512 samples every 256, no padding, np.hanning(512)) are exactly the
// INTERIOR frames 1..n-2 of the audfprint STFT (frame f of that transform covers samples
// [256 (f-1), 256 (f+1)), reflect padding only touches frames 0 and n-1), so the PSD reuses
// stft_mag_kernel with the symmetric window table and squares the magnitudes:
//   psd[k][j] = |X_{j+1}[k]|^2 * (2 if 0 < k < 256 else 1)      (/Fs and /sum(w^2) cancel in "/= max")
namespace {

__device__ __forceinline__ float block_reduce_max(float v, float* red) {
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[0];
  for (int w = 1; w < (int)blockDim.x / 32; ++w) r = fmaxf(r, red[w]);
  __syncthreads();
  return r;
}
__device__ __forceinline__ double block_reduce_sum(double v, double* red) {
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double r = 0.0;
  for (int w = 0; w < (int)blockDim.x / 32; ++w) r += red[w];
  __syncthreads();
  return r;
}

// mag [item][n_stft][264] (frame-major, Dejavu window) -> psd [item][257][nd] normalised by its maximum.
// One block per item; 32 x 32 tiles go through shared memory so both sides are coalesced.
__global__ void __launch_bounds__(256) dejavu_psd_kernel(const float* __restrict__ mag, int n_stft, int nd,
                                                         float* __restrict__ psd) {
  __shared__ float tile[32][33];
  __shared__ float red[8];
  const int item = blockIdx.x, tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* m = mag + (int64_t)item * n_stft * kPitch;
  float vmax = 0.f;
  for (int j = ty; j < nd; j += 8)
    for (int k = tx; k < kBins; k += 32) {
      const float a = m[(int64_t)(j + 1) * kPitch + k];
      vmax = fmaxf(vmax, a * a * ((k > 0 && k < kBins - 1) ? 2.f : 1.f));
    }
  vmax = block_reduce_max(vmax, red);
  float* out = psd + (int64_t)item * kBins * nd;
  for (int k0 = 0; k0 < kBins; k0 += 32)
    for (int j0 = 0; j0 < nd; j0 += 32) {
      for (int i = ty; i < 32; i += 8) {
        const int j = j0 + i, k = k0 + tx;
        float v = 0.f;
        if (j < nd && k < kBins) {
          const float a = m[(int64_t)(j + 1) * kPitch + k];
          v = a * a * ((k > 0 && k < kBins - 1) ? 2.f : 1.f) / vmax;   // arr2D /= arr2D.max()
        }
        tile[i][tx] = v;
      }
      __syncthreads();
      for (int i = ty; i < 32; i += 8) {
        const int k = k0 + i, j = j0 + tx;
        if (k < kBins && j < nd) out[(int64_t)k * nd + j] = tile[tx][i];
      }
      __syncthreads();
    }
}

// arr = 10 ln(max(p, max(p) / 1e6)) - mean, p = psd or psd^2 (after the UNet, fingerprint.py:74-79)
__global__ void __launch_bounds__(256) dejavu_log_kernel(const float* __restrict__ psd, int n, int square,
                                                         float* __restrict__ arr) {
  __shared__ float redf[8];
  __shared__ double redd[8];
  const int item = blockIdx.x;
  const float* p = psd + (int64_t)item * n;
  float* o = arr + (int64_t)item * n;
  float vmax = -INFINITY;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float v = square ? p[i] * p[i] : p[i];
    vmax = fmaxf(vmax, v);
  }
  vmax = block_reduce_max(vmax, redf);
  const float floor_v = vmax / 1e6f;
  double sum = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float v = square ? p[i] * p[i] : p[i];
    const float l = 10.f * logf(fmaxf(v, floor_v));
    o[i] = l;
    sum += (double)l;
  }
  const float mean = (float)(block_reduce_sum(sum, redd) / (double)n);
  for (int i = threadIdx.x; i < n; i += 256) o[i] -= mean;
}

}  // namespace

int dejavu_num_frames(int T) { return T >= kNfft ? (T - kHop) / kHop : 0; }

int launch_dejavu_psd(mfpa_ctx* ctx, const float* x, int B, int T, int64_t stride, float* psd, cudaStream_t st) {
  const int n_stft = num_frames(T), nd = dejavu_num_frames(T);
  MFPA_REQUIRE(nd >= 1, "dejavu_psd: %d samples are fewer than one 512-sample segment", T);
  if (ctx->mag.reserve(sizeof(float) * (size_t)B * n_stft * kPitch) || ctx->qmax.reserve(sizeof(float) * (size_t)B))
    return MFPA_ENOMEM;
  if (int e = launch_stft_mag(ctx, x, B, T, stride, 1, (float*)ctx->mag.ptr, (float*)ctx->qmax.ptr, st, ctx->win_dejavu_dev))
    return e;
  dejavu_psd_kernel<<<B, 256, 0, st>>>((const float*)ctx->mag.ptr, n_stft, nd, psd);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_dejavu_log(const float* psd, int B, int n, int square, float* arr, cudaStream_t st) {
  dejavu_log_kernel<<<B, 256, 0, st>>>(psd, n, square, arr);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

}  // namespace mfpa
