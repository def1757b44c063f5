// S2 — batched magnitude STFT (n_fft 512, hop 256, reflect padding, Hann(514)[1:-1]).
// Replaces afp/audfprint/stft.py:15-62 + np.abs (peak_extractor.py:257-261).
//
// A half-warp transforms one frame: the 512 windowed real samples are packed into a
// 256-point complex FFT (z[n] = x[2n] + i x[2n+1]) computed as 16 x 16 — a radix-16
// butterfly entirely in registers (two layers of radix-4 with constant twiddles), one
// transpose through a conflict-free padded shared-memory tile, a second radix-16 — and
// un-packed to the 257 real-FFT bins pair-wise (bins k and 256-k share their loads and
// even/odd parts).  A block stages 16 consecutive frames (17 * 256 samples, read once,
// coalesced, cp.async double-buffered) so the 50 % frame overlap never re-reads HBM.
// Twiddles come from float64-computed tables; the window is stored pre-multiplied by 1/2
// (the real-FFT un-packing's factor).
// Output is frame-major [item][frame][264] so both this kernel's stores and the
// peak picker's per-frame loads are fully coalesced.
#include <math.h>

#include "common.cuh"

namespace mfpa {

namespace {

constexpr int kWarps = 8;
constexpr int kFramesPerWarp = 2;               // one per half-warp
constexpr int kTile = kWarps * kFramesPerWarp;  // 16 frames per block
constexpr int kExF2 = 272;                      // float2 per frame exchange tile (16 rows x 17)
constexpr size_t kStftSmem =
    sizeof(float) * (2 * (kTile + 1) * kHop + kNfft) + sizeof(float2) * (256 + kWarps * 2 * kExF2);

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// complex add / subtract of interleaved (re, im) pairs: one packed f32x2 instruction each (same roundings)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }

// forward 4-point DFT in place: (a, b, c, d) = inputs n = 0..3 -> outputs k = 0..3
__device__ __forceinline__ void dft4(float2& a, float2& b, float2& c, float2& d) {
  const float2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = csub(b, d);
  a = cadd(s0, s2);
  c = csub(s0, s2);
  b = make_float2(s1.x + s3.y, s1.y - s3.x);  // s1 - i s3
  d = make_float2(s1.x - s3.y, s1.y + s3.x);  // s1 + i s3
}

// forward 16-point DFT in registers.  Input v[n]; output X[k] is left at v[4 (k & 3) + (k >> 2)].
#define FFT16_OUT(v, k) v[4 * ((k) & 3) + ((k) >> 2)]
__device__ __forceinline__ void fft16(float2 (&v)[16]) {
  const float h = 0.70710678118654752440f, c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);  // v[4 k1 + n2] = a[k1][n2]
  // a[k1][n2] *= W16^(n2 k1),  W16^m = cos(pi m / 8) - i sin(pi m / 8)
  v[5] = make_float2(v[5].x * c1 + v[5].y * s1, v[5].y * c1 - v[5].x * s1);        // m = 1
  v[6] = make_float2(h * (v[6].x + v[6].y), h * (v[6].y - v[6].x));                // m = 2
  v[7] = make_float2(v[7].x * s1 + v[7].y * c1, v[7].y * s1 - v[7].x * c1);        // m = 3
  v[9] = make_float2(h * (v[9].x + v[9].y), h * (v[9].y - v[9].x));                // m = 2
  v[10] = make_float2(v[10].y, -v[10].x);                                          // m = 4: * (-i)
  v[11] = make_float2(h * (v[11].y - v[11].x), -h * (v[11].x + v[11].y));          // m = 6
  v[13] = make_float2(v[13].x * s1 + v[13].y * c1, v[13].y * s1 - v[13].x * c1);   // m = 3
  v[14] = make_float2(h * (v[14].y - v[14].x), -h * (v[14].x + v[14].y));          // m = 6
  v[15] = make_float2(-v[15].x * c1 - v[15].y * s1, v[15].x * s1 - v[15].y * c1);  // m = 9: (-c1, +s1)
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);  // -> X[k1 + 4 k2]
}

__device__ __forceinline__ int reflect_index(int s, int len) {
  // numpy "reflect" padding (no edge repeat), valid for any pad width.
  if (len == 1) return 0;
  const int period = 2 * len - 2;
  s %= period;
  if (s < 0) s += period;
  return s < len ? s : period - s;
}

__device__ __forceinline__ float sqrt_approx(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));  // MUFU.SQRT, <= 1 ulp-ish; tolerance is 1e-4
  return r;
}

__device__ __forceinline__ float lg2_ftz(float x) {
  float r;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

struct TileInfo {
  const float* xq;  // first sample of the item (after its shift offset)
  int len, n_frames, f0, item;
};

struct ShiftOffsets { int off[MFPA_MAX_SHIFTS]; };  // int(shift / shifts * 256), computed on the host

__device__ __forceinline__ TileInfo tile_info(unsigned tile, unsigned tiles, const float* x, int T, int64_t x_stride,
                                              unsigned shifts, const ShiftOffsets& so) {
  TileInfo t;
  const unsigned item = tile / tiles;
  const unsigned q = item / shifts, sh = item - q * shifts;
  const int off = so.off[sh];
  t.item = (int)item;
  t.len = T - off;
  t.n_frames = 1 + t.len / kHop;
  t.f0 = (int)(tile - item * tiles) * kTile;
  t.xq = x + (int64_t)q * x_stride + off;
  return t;
}

// Stage samples [256*(f0-1), 256*(f0+kTile)) of one item into xs.  Every 16-byte chunk that lies inside
// the signal goes through cp.async (overlapping the previous tile's FFTs); only the chunks that reach past
// either end of the signal (numpy's reflect padding: the first and last tile of an item) use plain loads,
// and only as far as the tile's existing frames read.
__device__ __forceinline__ void stage_tile(const TileInfo& t, float* xs, int tid) {
  if (t.f0 >= t.n_frames) return;
  const int s0 = kHop * (t.f0 - 1);
  const int need = (min(t.n_frames - t.f0, kTile) + 1) * kHop;
  const bool aligned = (((uintptr_t)(t.xq + s0)) & 15) == 0;   // s0 is a multiple of 256
  for (int i = tid * 4; i < need; i += kWarps * 32 * 4) {
    const int s = s0 + i;
    if (aligned && s >= 0 && s + 4 <= t.len) {
      cp_async16(xs + i, t.xq + s);
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        int ss = s + e;
        if (ss < 0 || ss >= t.len) ss = reflect_index(ss, t.len);
        xs[i + e] = __ldg(t.xq + ss);
      }
    }
  }
}

__global__ void __launch_bounds__(kWarps * 32, 3)
stft_mag_kernel(const float* __restrict__ x, int T, int64_t x_stride, int shifts, int n_max, unsigned total_tiles,
                const __grid_constant__ ShiftOffsets so, const float2* __restrict__ tw, const float* __restrict__ win,
                float* __restrict__ mag, float* __restrict__ qmax) {
  extern __shared__ __align__(16) float smem[];
  float (*xs)[(kTile + 1) * kHop] = reinterpret_cast<float (*)[(kTile + 1) * kHop]>(smem);
  float* win_s = smem + 2 * (kTile + 1) * kHop;
  float2* tw_s = reinterpret_cast<float2*>(win_s + kNfft);  // [k1][n2] = W256^(n2 k1)
  float2* ex_all = tw_s + 256;

  const unsigned tiles = (n_max + kTile - 1) / kTile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int half = lane >> 4, l16 = lane & 15;
  for (int i = tid; i < kNfft; i += kWarps * 32) win_s[i] = win[i];
  for (int i = tid; i < 256; i += kWarps * 32) tw_s[i] = __ldg(tw + (((i >> 4) * (i & 15)) & 255));

  // un-packing twiddles W512^k for this lane's pairs k = lane + 32 j
  float2 tw3[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) tw3[j] = __ldg(tw + 256 + lane + 32 * j);

  float2* exw = ex_all + warp * 2 * kExF2;  // this warp's two tiles
  float2* ex = exw + half * kExF2;          // this half-warp's tile

  unsigned tile = blockIdx.x;
  int buf = 0;
  if (tile < total_tiles) stage_tile(tile_info(tile, tiles, x, T, x_stride, shifts, so), xs[0], tid);
  cp_async_commit();
  for (; tile < total_tiles; tile += gridDim.x, buf ^= 1) {
    const TileInfo ti = tile_info(tile, tiles, x, T, x_stride, shifts, so);
    // one barrier per tile: it publishes this tile's samples (staged one iteration ago) and proves every
    // warp is done with the other buffer, which the next tile's copy may then refill under this tile's FFTs
    cp_async_wait<0>();
    __syncthreads();
    const unsigned next = tile + gridDim.x;
    if (next < total_tiles) stage_tile(tile_info(next, tiles, x, T, x_stride, shifts, so), xs[buf ^ 1], tid);
    cp_async_commit();
    float vmax = 0.f;
    if (ti.f0 + 2 * warp < ti.n_frames) {  // warp-uniform: at least this warp's first frame exists
      const float* xf = xs[buf] + (2 * warp + half) * kHop + 2 * l16;
      const float* wf = win_s + 2 * l16;
      float2 v[16];
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {  // z[16 n1 + l16]
        const float2 s = *reinterpret_cast<const float2*>(xf + 32 * n1);
        const float2 w = *reinterpret_cast<const float2*>(wf + 32 * n1);
        v[n1] = make_float2(s.x * w.x, s.y * w.y);
      }
      fft16(v);  // over n1 -> k1
      ex[l16] = FFT16_OUT(v, 0);
#pragma unroll
      for (int k1 = 1; k1 < 16; ++k1) ex[k1 * 17 + l16] = cmul(FFT16_OUT(v, k1), tw_s[k1 * 16 + l16]);
      __syncwarp();
#pragma unroll
      for (int n2 = 0; n2 < 16; ++n2) v[n2] = ex[l16 * 17 + n2];  // lane = k1 now
      __syncwarp();
      fft16(v);  // over n2 -> k2: Z[l16 + 16 k2]
#pragma unroll
      for (int k2 = 0; k2 < 16; ++k2) ex[l16 + 16 * k2] = FFT16_OUT(v, k2);
      __syncwarp();
      // real-FFT un-packing, whole warp per frame: X[k] = E + W^k O, X[256-k] = conj(E - W^k O)
      float lsum = 0.f, lmin = 3.4e38f;  // picker statistics of this warp's frame pair (see mfpa.h, mag layout)
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        const int f = ti.f0 + 2 * warp + hh;
        if (f < ti.n_frames) {  // warp-uniform
          const float2* zs = exw + hh * kExF2;
          float* out = mag + ((int64_t)ti.item * n_max + f) * kPitch;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = lane + 32 * j;
            const float2 a = zs[k], c = zs[(256 - k) & 255];
            const float er = a.x + c.x, ei = a.y - c.y, o_r = a.y + c.y, o_i = c.x - a.x;
            const float tr = tw3[j].x * o_r - tw3[j].y * o_i, ti_ = tw3[j].x * o_i + tw3[j].y * o_r;
            const float x1r = er + tr, x1i = ei + ti_, x2r = er - tr, x2i = ei - ti_;
            const float m1 = sqrt_approx(x1r * x1r + x1i * x1i), m2 = sqrt_approx(x2r * x2r + x2i * x2i);
            out[k] = m1;
            out[256 - k] = m2;
            vmax = fmaxf(vmax, fmaxf(m1, m2));
            lsum += lg2_ftz(m1) + lg2_ftz(m2);
            lmin = fminf(lmin, fminf(m1, m2));
          }
          if (lane == 0) {  // bin 128 pairs with itself: |X[128]| = |Z[128]| (Z was halved by the window)
            const float2 a = zs[128];
            const float m = 2.0f * sqrt_approx(a.x * a.x + a.y * a.y);
            out[128] = m;
            vmax = fmaxf(vmax, m);
            lsum += lg2_ftz(m);
            lmin = fminf(lmin, m);
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        lsum += __shfl_xor_sync(0xffffffffu, lsum, o);
        lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, o));
      }
      if (lane == 0) {
        float* row = mag + ((int64_t)ti.item * n_max + ti.f0 + 2 * warp) * kPitch;
        *reinterpret_cast<float2*>(row + kBins + 1) = make_float2(lsum, lmin);  // elements 258, 259
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0 && ti.f0 + 2 * warp < ti.n_frames) atomicMax(reinterpret_cast<int*>(qmax + ti.item), __float_as_int(vmax));
  }
  cp_async_wait<0>();
}

// mag [item][frame][264] f32 -> spec [item][257][n_max] f64, divided by qmax.
__global__ void spec_from_mag_kernel(const float* __restrict__ mag, const float* __restrict__ qmax,
                                     int T, int shifts, int n_max, double* __restrict__ spec,
                                     int item_base) {
  __shared__ float tile[32][33];
  const int item = blockIdx.z;  // relative to the chunk the pointers were advanced to
  const int sh = (item_base + item) % shifts;
  const int off = shifts < 2 ? 0 : (int)((double)sh / (double)shifts * (double)kHop);
  const int n_frames = 1 + (T - off) / kHop;
  const int c0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, b = b0 + tx;
    tile[i][tx] = (c < n_frames && b < kBins) ? mag[((int64_t)item * n_max + c) * kPitch + b] : 0.f;
  }
  __syncthreads();
  const double m = (double)qmax[item];
  for (int i = ty; i < 32; i += 8) {
    const int b = b0 + i, c = c0 + tx;
    if (b < kBins && c < n_max)
      spec[((int64_t)item * kBins + b) * n_max + c] = c < n_frames ? (double)tile[tx][i] / m : 0.0;
  }
}


// ---- complex STFT in float64 (afp/audfprint/stft.py:15-62: the module's public function) -------------
// One block per frame, one thread per bin: a direct DFT of the windowed frame with an exact-to-rounding
// float64 twiddle table.  Not on the throughput path (find_peaks uses stft_mag_kernel); it exists so that
// stft.stft() keeps its meaning - complex128 [n_fft/2 + 1][frames] - for callers of the module.
__global__ void __launch_bounds__(256) stft_complex_kernel(const double* __restrict__ x, int T, int n_fft, int hop,
                                                           const double* __restrict__ win, int L, int n_frames,
                                                           double2* __restrict__ out) {
  extern __shared__ __align__(16) double sm_d[];
  double2* tw = reinterpret_cast<double2*>(sm_d);   // [n_fft]
  double* fr = sm_d + 2 * n_fft;                    // [L] windowed frame
  const int f = blockIdx.x, pad = n_fft / 2;
  for (int i = threadIdx.x; i < L; i += blockDim.x) fr[i] = x[reflect_index(f * hop + i - pad, T)] * win[i];
  for (int m = threadIdx.x; m < n_fft; m += blockDim.x) {
    double sn, cs;
    sincospi(2.0 * (double)m / (double)n_fft, &sn, &cs);
    tw[m] = make_double2(cs, -sn);
  }
  __syncthreads();
  for (int k = threadIdx.x; k <= n_fft / 2; k += blockDim.x) {
    double re = 0.0, im = 0.0;
    int idx = 0;
    for (int n = 0; n < L; ++n) {
      const double2 w = tw[idx];
      re += fr[n] * w.x;
      im += fr[n] * w.y;
      idx += k;
      if (idx >= n_fft) idx -= n_fft;
    }
    out[(int64_t)k * n_frames + f] = make_double2(re, im);
  }
}

}  // namespace

int stft_init_tables(mfpa_ctx* ctx) {
  float2 tw[256 + 257];
  float win[kNfft];
  const double pi = 3.14159265358979323846;
  for (int k = 0; k < 256; ++k) tw[k] = make_float2((float)cos(2 * pi * k / 256), (float)-sin(2 * pi * k / 256));
  for (int k = 0; k <= 256; ++k) tw[256 + k] = make_float2((float)cos(2 * pi * k / 512), (float)-sin(2 * pi * k / 512));
  // np.hanning(514)[1:-1]: 0.5 - 0.5*cos(2*pi*n/513), n = 1..512
  for (int n = 0; n < kNfft; ++n) win[n] = (float)(0.5 * (0.5 - 0.5 * cos(2 * pi * (n + 1) / (kNfft + 1))));  // x 1/2: real-FFT un-packing factor
  MFPA_CUDA(cudaMalloc(&ctx->tw_dev, sizeof(tw)));
  MFPA_CUDA(cudaMalloc(&ctx->win_dev, sizeof(win)));
  MFPA_CUDA(cudaMemcpy(ctx->tw_dev, tw, sizeof(tw), cudaMemcpyHostToDevice));
  MFPA_CUDA(cudaMemcpy(ctx->win_dev, win, sizeof(win), cudaMemcpyHostToDevice));
  // np.hanning(512): 0.5 - 0.5*cos(2*pi*n/511) (symmetric; matplotlib.mlab.window_hanning)
  for (int n = 0; n < kNfft; ++n) win[n] = (float)(0.5 * (0.5 - 0.5 * cos(2 * pi * n / (kNfft - 1))));
  MFPA_CUDA(cudaMalloc(&ctx->win_dejavu_dev, sizeof(win)));
  MFPA_CUDA(cudaMemcpy(ctx->win_dejavu_dev, win, sizeof(win), cudaMemcpyHostToDevice));
  return MFPA_OK;
}

int launch_stft_mag(mfpa_ctx* ctx, const float* x, int B, int T, int64_t stride, int shifts,
                    float* mag, float* qmax, cudaStream_t st, const float* window) {
  const int items = B * shifts;
  const int n_max = num_frames(T);
  MFPA_CUDA(cudaMemsetAsync(qmax, 0, sizeof(float) * items, st));
  const int64_t total_tiles = (int64_t)items * ((n_max + kTile - 1) / kTile);
  MFPA_REQUIRE(total_tiles < (1ll << 31) - 4096, "stft_mag: batch too large (%lld tiles); split it", (long long)total_tiles);
  const int64_t resident = (int64_t)ctx->num_sms * 3;  // persistent blocks, 3 per SM (__launch_bounds__)
  const unsigned blocks = (unsigned)(total_tiles < resident ? total_tiles : resident);
  ShiftOffsets so{};
  for (int sft = 0; sft < shifts; ++sft) so.off[sft] = shift_offset(sft, shifts);
  MFPA_CUDA(cudaFuncSetAttribute(stft_mag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStftSmem));
  stft_mag_kernel<<<blocks, kWarps * 32, kStftSmem, st>>>(x, T, stride, shifts, n_max, (unsigned)total_tiles, so, ctx->tw_dev,
                                                          window ? window : ctx->win_dev, mag, qmax);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_spec_from_mag(const float* mag, const float* qmax, int B, int T, int shifts, double* spec,
                         cudaStream_t st) {
  const int n_max = num_frames(T);
  const int items = B * shifts;
  for (int i0 = 0; i0 < items; i0 += 32768) {  // 32768 is a multiple of every shifts <= 8 power of two... keep item%shifts intact
    const int n = items - i0 < 32768 ? items - i0 : 32768;
    dim3 grid((n_max + 31) / 32, (kBins + 31) / 32, n);
    spec_from_mag_kernel<<<grid, dim3(32, 8), 0, st>>>(mag + (int64_t)i0 * n_max * kPitch, qmax + i0, T, shifts,
                                                      n_max, spec + (int64_t)i0 * kBins * n_max, i0);
    MFPA_CUDA(cudaGetLastError());
  }
  return MFPA_OK;
}

int launch_stft_complex(const double* x, int T, int n_fft, int hop, const double* win, int win_len, int n_frames,
                        double* out, cudaStream_t st) {
  const int L = win_len < n_fft ? win_len : n_fft;   // np.fft.rfft(frames, n_fft) crops / zero-pads the frame
  const size_t smem = sizeof(double) * (2 * (size_t)n_fft + L);
  MFPA_REQUIRE(smem <= 200 * 1024, "stft_complex: n_fft %d too large for one block's shared memory", n_fft);
  MFPA_CUDA(cudaFuncSetAttribute(stft_complex_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  stft_complex_kernel<<<n_frames, 256, smem, st>>>(x, T, n_fft, hop, win, L, n_frames, reinterpret_cast<double2*>(out));
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

}  // namespace mfpa
