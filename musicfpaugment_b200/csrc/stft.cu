// S2 — batched magnitude STFT (n_fft 512, hop 256, reflect padding, Hann(514)[1:-1]).
// Replaces afp/audfprint/stft.py:15-62 + np.abs (peak_extractor.py:257-261).
//
// One warp transforms one frame: the 512 real samples are packed into a
// 256-point complex FFT (z[n] = x[2n] + i x[2n+1]) computed as radix 8 x 8 x 4 —
// two in-register radix-8 passes exchanged through a conflict-free padded
// shared-memory tile, one radix-4 pass across lane quads with shuffles — and
// un-packed to the 257 real-FFT bins.  A block stages 16 consecutive frames
// (17 * 256 samples, read once, coalesced) so the 50 % frame overlap never
// re-reads HBM.  Twiddles come from a float64-computed table.
// Output is frame-major [item][frame][264] so both this kernel's stores and the
// peak picker's per-frame loads are fully coalesced.
#include <math.h>

#include "common.cuh"

namespace mfpa {

namespace {

constexpr int kWarps = 8;
constexpr int kFramesPerWarp = 2;
constexpr int kTile = kWarps * kFramesPerWarp;  // 16 frames per block
constexpr int kExStride = 36;                   // padded row stride of the exchange tile
constexpr int kExSize = 8 * kExStride;          // 288 floats per component
constexpr size_t kStftSmem = sizeof(float) * (2 * (kTile + 1) * kHop + kNfft + 2 * kWarps * kExSize);

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }  // a * (-i)

// In-place forward 8-point DFT, natural order in and out.
__device__ __forceinline__ void fft8(float2 (&v)[8]) {
  const float h = 0.70710678118654752440f;
  float2 a0 = cadd(v[0], v[4]), a1 = cadd(v[1], v[5]), a2 = cadd(v[2], v[6]), a3 = cadd(v[3], v[7]);
  float2 b0 = csub(v[0], v[4]), b1 = csub(v[1], v[5]), b2 = csub(v[2], v[6]), b3 = csub(v[3], v[7]);
  b1 = make_float2(h * (b1.x + b1.y), h * (b1.y - b1.x));   // * (1 - i)/sqrt2
  b2 = mul_mi(b2);                                          // * (-i)
  b3 = make_float2(h * (b3.y - b3.x), -h * (b3.x + b3.y));  // * (-1 - i)/sqrt2
  float2 c0 = cadd(a0, a2), c1 = cadd(a1, a3), d0 = csub(a0, a2), d1 = mul_mi(csub(a1, a3));
  v[0] = cadd(c0, c1); v[4] = csub(c0, c1); v[2] = cadd(d0, d1); v[6] = csub(d0, d1);
  c0 = cadd(b0, b2); c1 = cadd(b1, b3); d0 = csub(b0, b2); d1 = mul_mi(csub(b1, b3));
  v[1] = cadd(c0, c1); v[5] = csub(c0, c1); v[3] = cadd(d0, d1); v[7] = csub(d0, d1);
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

__device__ __forceinline__ TileInfo tile_info(int64_t tile, int tiles, const float* x, int T, int64_t x_stride, int shifts) {
  TileInfo t;
  t.item = (int)(tile / tiles);
  const int q = t.item / shifts, sh = t.item - q * shifts;
  const int off = shifts < 2 ? 0 : (int)((double)sh / (double)shifts * (double)kHop);
  t.len = T - off;
  t.n_frames = 1 + t.len / kHop;
  t.f0 = (int)(tile - (int64_t)t.item * tiles) * kTile;
  t.xq = x + (int64_t)q * x_stride + off;
  return t;
}

// Stage samples [256*(f0-1), 256*(f0+kTile)) of one item into xs.  Interior tiles whose
// source is 16-byte aligned go through cp.async (overlapping the previous tile's FFTs);
// tiles touching either end of the signal apply numpy's reflect padding with plain loads.
__device__ __forceinline__ void stage_tile(const TileInfo& t, float* xs, int tid) {
  if (t.f0 >= t.n_frames) return;
  const int s0 = kHop * (t.f0 - 1);
  const bool interior = s0 >= 0 && s0 + (kTile + 1) * kHop <= t.len && ((((uintptr_t)(t.xq + s0)) & 15) == 0);
  if (interior) {
    for (int i = tid * 4; i < (kTile + 1) * kHop; i += kWarps * 32 * 4) cp_async16(xs + i, t.xq + s0 + i);
  } else {
    for (int i = tid; i < (kTile + 1) * kHop; i += kWarps * 32) {
      int s = s0 + i;
      if (s < 0 || s >= t.len) s = reflect_index(s, t.len);
      xs[i] = __ldg(t.xq + s);
    }
  }
}

__global__ void __launch_bounds__(kWarps * 32, 3)
stft_mag_kernel(const float* __restrict__ x, int T, int64_t x_stride, int shifts, int n_max, int64_t total_tiles,
                const float2* __restrict__ tw, const float* __restrict__ win,
                float* __restrict__ mag, float* __restrict__ qmax) {
  extern __shared__ __align__(16) float smem[];
  float (*xs)[(kTile + 1) * kHop] = reinterpret_cast<float (*)[(kTile + 1) * kHop]>(smem);
  float* win_s = smem + 2 * (kTile + 1) * kHop;
  float (*ex_re)[kExSize] = reinterpret_cast<float (*)[kExSize]>(win_s + kNfft);
  float (*ex_im)[kExSize] = reinterpret_cast<float (*)[kExSize]>(win_s + kNfft + kWarps * kExSize);

  const int tiles = (n_max + kTile - 1) / kTile;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < kNfft; i += kWarps * 32) win_s[i] = win[i];

  // per-lane twiddles, loaded once per (persistent) block
  float2 tw1[8], tw2[8], tw3[8];
  const int g = lane >> 2, p = lane & 3;
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    tw1[k] = __ldg(tw + lane * k);   // W256^(n2*k1)
    tw2[k] = __ldg(tw + 8 * p * k);  // W32^(p*q)
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) tw3[j] = __ldg(tw + 256 + lane + 32 * j);  // W512^k

  float* er = ex_re[warp];
  float* ei = ex_im[warp];

  int64_t tile = blockIdx.x;
  int buf = 0;
  if (tile < total_tiles) stage_tile(tile_info(tile, tiles, x, T, x_stride, shifts), xs[0], tid);
  cp_async_commit();
  for (; tile < total_tiles; tile += gridDim.x, buf ^= 1) {
    const TileInfo ti = tile_info(tile, tiles, x, T, x_stride, shifts);
    const int64_t next = tile + gridDim.x;
    if (next < total_tiles) stage_tile(tile_info(next, tiles, x, T, x_stride, shifts), xs[buf ^ 1], tid);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    float vmax = 0.f;
    for (int fi = 0; fi < kFramesPerWarp; ++fi) {
      const int fl = warp + kWarps * fi;
      const int f = ti.f0 + fl;
      if (f >= ti.n_frames) break;  // warp-uniform
      const float* xf = xs[buf] + fl * kHop;
      float2 v[8];
#pragma unroll
      for (int n1 = 0; n1 < 8; ++n1) {
        const int n = 2 * (32 * n1 + lane);
        const float2 s = *reinterpret_cast<const float2*>(xf + n);
        const float2 w = *reinterpret_cast<const float2*>(win_s + n);
        v[n1] = make_float2(s.x * w.x, s.y * w.y);
      }
      fft8(v);  // over n1 -> k1
#pragma unroll
      for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], tw1[k]);
#pragma unroll
      for (int k = 0; k < 8; ++k) { er[k * kExStride + lane] = v[k].x; ei[k * kExStride + lane] = v[k].y; }
      __syncwarp();
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const int idx = g * kExStride + 4 * m + p;
        v[m] = make_float2(er[idx], ei[idx]);
      }
      __syncwarp();
      fft8(v);  // over m -> q
#pragma unroll
      for (int k = 1; k < 8; ++k) v[k] = cmul(v[k], tw2[k]);
      // radix-4 across the lane quad (p); lane p ends with output r = bitrev2(p)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float2 t = make_float2(__shfl_xor_sync(0xffffffffu, v[k].x, 2), __shfl_xor_sync(0xffffffffu, v[k].y, 2));
        float2 c = (p & 2) ? csub(t, v[k]) : cadd(v[k], t);
        if (p == 3) c = mul_mi(c);
        t = make_float2(__shfl_xor_sync(0xffffffffu, c.x, 1), __shfl_xor_sync(0xffffffffu, c.y, 1));
        v[k] = (p & 1) ? csub(t, c) : cadd(c, t);
      }
      const int r = ((p & 1) << 1) | (p >> 1);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int kk = g + 8 * k + 72 * r;  // natural index g + 8q + 64r, padded by 8 per 64
        er[kk] = v[k].x; ei[kk] = v[k].y;
      }
      __syncwarp();
      float* out = mag + ((int64_t)ti.item * n_max + f) * kPitch;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = lane + 32 * j;
        const int kc = (256 - k) & 255;
        const int ka = k + 8 * (k >> 6), kb = kc + 8 * (kc >> 6);
        const float ar = er[ka], ai = ei[ka], cr = er[kb], ci = ei[kb];
        const float e_r = 0.5f * (ar + cr), e_i = 0.5f * (ai - ci);
        const float o_r = 0.5f * (ai + ci), o_i = -0.5f * (ar - cr);
        const float xr = e_r + tw3[j].x * o_r - tw3[j].y * o_i;
        const float xi = e_i + tw3[j].x * o_i + tw3[j].y * o_r;
        const float m = sqrt_approx(xr * xr + xi * xi);
        out[k] = m;
        vmax = fmaxf(vmax, m);
      }
      if (lane == 0) {  // Nyquist bin: Re(Z0) - Im(Z0)
        const float m = fabsf(er[0] - ei[0]);
        out[256] = m;
        vmax = fmaxf(vmax, m);
      }
      __syncwarp();
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (lane == 0 && ti.f0 + warp < ti.n_frames) atomicMax(reinterpret_cast<int*>(qmax + ti.item), __float_as_int(vmax));
    __syncthreads();  // everyone is done with xs[buf] before it is refilled two iterations later
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

}  // namespace

int stft_init_tables(mfpa_ctx* ctx) {
  float2 tw[256 + 257];
  float win[kNfft];
  const double pi = 3.14159265358979323846;
  for (int k = 0; k < 256; ++k) tw[k] = make_float2((float)cos(2 * pi * k / 256), (float)-sin(2 * pi * k / 256));
  for (int k = 0; k <= 256; ++k) tw[256 + k] = make_float2((float)cos(2 * pi * k / 512), (float)-sin(2 * pi * k / 512));
  // np.hanning(514)[1:-1]: 0.5 - 0.5*cos(2*pi*n/513), n = 1..512
  for (int n = 0; n < kNfft; ++n) win[n] = (float)(0.5 - 0.5 * cos(2 * pi * (n + 1) / (kNfft + 1)));
  MFPA_CUDA(cudaMalloc(&ctx->tw_dev, sizeof(tw)));
  MFPA_CUDA(cudaMalloc(&ctx->win_dev, sizeof(win)));
  MFPA_CUDA(cudaMemcpy(ctx->tw_dev, tw, sizeof(tw), cudaMemcpyHostToDevice));
  MFPA_CUDA(cudaMemcpy(ctx->win_dev, win, sizeof(win), cudaMemcpyHostToDevice));
  return MFPA_OK;
}

int launch_stft_mag(mfpa_ctx* ctx, const float* x, int B, int T, int64_t stride, int shifts,
                    float* mag, float* qmax, cudaStream_t st) {
  const int items = B * shifts;
  const int n_max = num_frames(T);
  MFPA_CUDA(cudaMemsetAsync(qmax, 0, sizeof(float) * items, st));
  const int64_t total_tiles = (int64_t)items * ((n_max + kTile - 1) / kTile);
  const int64_t resident = (int64_t)ctx->num_sms * 3;  // persistent blocks, 3 per SM (__launch_bounds__)
  const unsigned blocks = (unsigned)(total_tiles < resident ? total_tiles : resident);
  MFPA_CUDA(cudaFuncSetAttribute(stft_mag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStftSmem));
  stft_mag_kernel<<<blocks, kWarps * 32, kStftSmem, st>>>(x, T, stride, shifts, n_max, total_tiles, ctx->tw_dev, ctx->win_dev,
                                                  mag, qmax);
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

}  // namespace mfpa
