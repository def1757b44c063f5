// S1 — AugmentFP degradation chain on dumped parameters (augmentation/__init__.py:46-97).
//
//   stage 1  HighPassFilter(fc1)        x - lowpass(x)                     pass_filters.py:144-155
//   stage 2  ApplyImpulseResponse       full conv, / peak of FULL result   impulse_response.py:73-116
//   stage 3  AddBackgroundNoise         + rms/10^(snr/20) * noise, / peak  background_noise.py:183-213
//   stage 4  Gain, Clipping             * g, clamp to p/2, 1-p/2 quantiles gain.py:62-70, clipping.py:67-101
//   stage 5  LowPassFilter(fc2)         windowed-sinc FIR (julius)         pass_filters.py:84-115
//   stage 6  HighPassFilter(fc3)
//   stage 7  PeakNormalization          / peak if > 0                      peak_normalization.py:38-67
//
// Long convolutions (stages 1, 2, 6) are overlap-save blocks of 16384 REAL samples transformed
// as one 8192-point complex FFT held in shared memory (even samples in the real part, odd samples
// in the imaginary part; fftconv_core.cuh: planar layout, two butterflies per thread in packed
// FADD2/FMUL2/FFMA2 arithmetic).  A per-query kernel turns the filter (FIR taps generated on the
// fly, or the impulse response) into its "double pairs" once; every signal block then does
// forward FFT -> one fused pass (split into the real-signal spectrum, multiply by the filter,
// re-pack) -> inverse FFT.  Forward is decimation-in-frequency and the inverse undoes it pass by
// pass, so no bit-reversal pass is ever needed; the filter is stored in the order the fused pass
// reads it (coalesced 16-byte loads).  Block starts and filter delays are kept multiples of four
// samples so the waveform is read and written with 16-byte accesses.  The per-query reductions that gate the next stage (peak of the full convolution,
// RMS, peak after the mix, clip quantiles, final peak) are produced by the kernel that writes the data.
// Filters longer than one block takes (FIRs above 8193 taps, responses above 8192 samples) run as uniformly
// partitioned overlap-save on the same transform (part_* kernels below).  The clip quantiles cost no pass
// over the signal: a sampled pair of tail thresholds, the tails listed by the mix kernel, a shared-memory
// radix select (clip_sample / mix / clip_finish).  The file also holds the noise-row assembly that runs
// ahead of the chain (mfpa_noise_assemble).
#include <math.h>

#include <vector>

#include "common.cuh"
#include "fftconv_core.cuh"

namespace mfpa {

namespace {

using fc::FN;
using fc::FM;
using fc::FT;
using fc::FPADF;
using fc::c2;
constexpr int kTwFloats = 2 * fc::kTwA + 2 * fc::kTwB;   // pass-A and pass-B twiddle tables, planar
constexpr size_t kConvSmem = sizeof(float) * (fc::kPlaneFloats + 64);
constexpr float kPi = 3.14159265358979323846f;

struct AugQ {            // per-query derived parameters (device copy)
  uint32_t apply;
  int half1, half2, half3, ir_len;
  float c1x2, c2x2, c3x2;         // float(2*cutoff)
  float arg1, arg2, arg3;         // float(2*cutoff*pi)
  float snr_div;                  // 10^(snr_db/20)
  float gain, q_lo;
  uint32_t long_mask;             // stages whose filter is longer than one overlap-save block takes (kLong* bits)
  uint32_t pad_;
  int64_t ir_off;                 // first sample of this query's impulse response in the ir buffer
};                                // sizeof % 8 == 0: the AugS array (doubles) follows the AugQ array in one allocation
static_assert(sizeof(AugQ) % 8 == 0, "AugS follows AugQ[B] and holds doubles");

struct AugS {            // per-query running statistics (zeroed per call)
  double ss_a, ss_b;     // sum of squares of the stage-1 / stage-2 output (first T samples)
  float max_a, max_b;    // peak of stage-1 output / of the FULL stage-2 convolution
  float max_z, max_v;    // peak after the noise mix / after stage 6
  float lo, hi;          // clip thresholds in the gained domain
  unsigned t_lo_key, t_hi_key;   // tail thresholds placed by clip_sample_kernel (order keys)
  unsigned cnt_lo, cnt_hi;       // samples beyond them, counted (and listed) by mix_kernel
  float pad[2];
};

// shared-memory carve-up of the convolution kernels
struct ConvSmem {
  float *re, *im, *twa_re, *twa_im, *twb_re, *twb_im, *red;
  __device__ explicit ConvSmem(float* s)
      : re(s), im(s + FPADF), twa_re(s + 2 * FPADF), twa_im(twa_re + fc::kTwA), twb_re(twa_im + fc::kTwA),
        twb_im(twb_re + fc::kTwB), red(twb_im + fc::kTwB) {}
  __device__ void load_tables(const float* __restrict__ tw_g, int tid) const {
    for (int i = tid; i < kTwFloats; i += FT) twa_re[i] = __ldg(tw_g + i);
  }
};

// forward: natural order in, digit-reversed out.  inverse: digit-reversed in, natural out (x FM).  Both end
// with a barrier; the caller provides the one that publishes the input.
__device__ __forceinline__ void fft_forward(const ConvSmem& s, int tid) {
  fc::pass_a<false>(s.re, s.im, s.twa_re, s.twa_im, tid); __syncthreads();
  fc::pass_b<false>(s.re, s.im, s.twb_re, s.twb_im, tid); __syncthreads();
  fc::pass_last_fwd(s.re, s.im, tid);                     __syncthreads();
}
__device__ __forceinline__ void fft_inverse(const ConvSmem& s, int tid) {
  fc::pass_last_inv(s.re, s.im, tid);                     __syncthreads();
  fc::pass_b<true>(s.re, s.im, s.twb_re, s.twb_im, tid);  __syncthreads();
  fc::pass_a<true>(s.re, s.im, s.twa_re, s.twa_im, tid);  __syncthreads();
}

__device__ __forceinline__ float block_sum(float v, float* red, int tid) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < FT / 32; ++w) t += red[w];
  __syncthreads();
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red, int tid) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < FT / 32; ++w) t = fmaxf(t, red[w]);
  __syncthreads();
  return t;
}
__device__ __forceinline__ void atomic_max_pos(float* addr, float v) { atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v)); }

// julius low-pass tap i of a filter with half-width `half` (unnormalised): 2c * hann * sinc
__device__ __forceinline__ float fir_tap(int i, int half, float c2, float argscale) {
  if (half == 0) return c2;
  const float win = 0.5f - 0.5f * cosf(2.0f * kPi * (float)i / (float)(2 * half));
  const float arg = (float)(i - half) * argscale;
  const float sinc = arg == 0.f ? 1.f : sinf(arg) / arg;
  return c2 * win * sinc;
}

enum { kModeHP = 0, kModeIR = 1, kModeLP = 2 };

struct ConvArgs {
  const float* in; int64_t in_stride;   // [B][T]
  float* out;                            // [B][T] contiguous
  const float* ir; int ir_stride;        // kModeIR
  const AugQ* q; AugS* st;
  int T; uint32_t bit; int which;        // which FIR (1, 2 or 3) / which stats slot
  float4* hspec;
  int highpass;                          // centred FIR: 1 = the filter is delta - lowpass (x - lowpass(x), pass_filters.py:149-155)                         // [B][2][FM/4] filter double pairs (fc::dp_index): plane 0 = S, plane 1 = R
};

__device__ __forceinline__ void filter_params(const AugQ& q, int which, int& half, float& c2, float& argscale) {
  half = which == 1 ? q.half1 : (which == 2 ? q.half2 : q.half3);
  c2 = which == 1 ? q.c1x2 : (which == 2 ? q.c2x2 : q.c3x2);
  argscale = which == 1 ? q.arg1 : (which == 2 ? q.arg2 : q.arg3);
}

enum { kLongHP1 = 1u, kLongIR = 2u, kLongLP = 4u, kLongHP3 = 8u };
template <int MODE> __device__ __forceinline__ uint32_t long_bit(int which) {
  return MODE == kModeIR ? kLongIR : (which == 1 ? kLongHP1 : (which == 2 ? kLongLP : kLongHP3));
}

using fc::ConvGeom;
using fc::conv_geom;

__device__ __forceinline__ void store_filter_pairs(const ConvSmem& s, float scale, float4* __restrict__ hs,
                                                   float4* __restrict__ hr, int tid) {
  fc::filter_pairs_store(s.re, s.im, scale, hs, hr, tid);
}
__device__ __forceinline__ void apply_filter_pairs(const ConvSmem& s, const float4* __restrict__ hs,
                                                   const float4* __restrict__ hr, int tid) {
  fc::filter_pairs_apply(s.re, s.im, hs, hr, tid);
}

// Packs FN real samples v(i), i = 0 .. FN-1, into the planes (even samples -> re, odd -> im): thread tid
// provides float4 number tid + FT u through `quad(i4)`.
template <int BATCH, typename F> __device__ __forceinline__ void fill_planes(const ConvSmem& s, int tid, F quad) {
  // BATCH global loads are issued together ahead of their shared-memory stores;
  // padi(2 (tid + FT u)) = padi(2 tid) + 544 u
  const int o0 = fc::padi(2 * tid);
#pragma unroll 1
  for (int u0 = 0; u0 < FN / 4 / FT; u0 += BATCH) {
    float4 v[BATCH];
#pragma unroll
    for (int u = 0; u < BATCH; ++u) v[u] = quad(tid + FT * (u0 + u));
#pragma unroll
    for (int u = 0; u < BATCH; ++u) {
      const int o = o0 + 544 * (u0 + u);
      *reinterpret_cast<float2*>(s.re + o) = make_float2(v[u].x, v[u].z);
      *reinterpret_cast<float2*>(s.im + o) = make_float2(v[u].y, v[u].w);
    }
  }
}

// One block per query: double pairs of the query's filter (zero-padded to FN real samples), scaled by
// 1/FM (the unscaled inverse transform) and, for the FIRs, by 1/sum(taps) (julius normalises to DC gain 1).
template <int MODE>
__global__ void __launch_bounds__(FT, 3) filter_spectrum_kernel(const ConvArgs a, const float* __restrict__ tw_g) {
  extern __shared__ __align__(16) float smem_f[];
  const ConvSmem s(smem_f);
  const int tid = threadIdx.x, qi = blockIdx.x;
  const AugQ q = a.q[qi];
  if (!(q.apply & a.bit) || (q.long_mask & long_bit<MODE>(a.which))) return;
  int half = 0;
  float c2 = 0.f, argscale = 0.f;
  if (MODE != kModeIR) filter_params(q, a.which, half, c2, argscale);
  const ConvGeom g = conv_geom(MODE == kModeIR, half, q.ir_len, a.T);
  s.load_tables(tw_g, tid);
  const float* ir = MODE == kModeIR ? a.ir + q.ir_off : nullptr;
  float hsum = 0.f;
  fill_planes<1>(s, tid, [&](int i4) {
    float h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = 4 * i4 + e;
      h[e] = 0.f;
      if (i < g.K) {
        if (MODE == kModeIR) h[e] = ir[i];
        else if (i >= g.zeros) h[e] = fir_tap(i - g.zeros, half, c2, argscale);
      }
      hsum += h[e];
      if (MODE != kModeIR && a.highpass) h[e] = -h[e];
    }
    return make_float4(h[0], h[1], h[2], h[3]);
  });
  float scale = 1.0f / (float)FM;
  if (MODE != kModeIR) {
    const float sum = block_sum(hsum, s.red, tid);   // ends with a barrier: the planes are complete
    scale /= sum;
    // x - lowpass(x) = (delta - lowpass) * x: the unit tap sits at the filter's delay g.lead (a multiple of 4),
    // worth `sum` before the normalisation by 1 / sum
    if (a.highpass && tid == 0) s.re[fc::padi(g.lead >> 1)] += sum;
  }
  __syncthreads();
  fft_forward(s, tid);
  float4* hs = a.hspec + (size_t)qi * (FM / 2);
  store_filter_pairs(s, scale, hs, hs + FM / 4, tid);
}

// OCC = blocks per SM the register allocation aims at: 2 (up to 128 registers) or 3 (80 registers, 24 warps)
template <int MODE, int OCC>
__global__ void __launch_bounds__(FT, OCC) fftconv_kernel(const ConvArgs a, const float* __restrict__ tw_g) {
  extern __shared__ __align__(16) float smem_f[];
  const ConvSmem s(smem_f);
  const int tid = threadIdx.x, qi = blockIdx.y, blk = blockIdx.x;
  const AugQ q = a.q[qi];
  const float* __restrict__ in = a.in + (int64_t)qi * a.in_stride;
  float* __restrict__ out = a.out + (int64_t)qi * a.T;
  const int T = a.T;
  float vmax = 0.f, vss = 0.f;
  // statistics the later stages read: peak and energy of the response stage's output (noise mix), peak of the last
  // stage's (final normalisation); the first stage's output needs neither
  const bool want_max = MODE == kModeIR || a.which == 3, want_ss = MODE == kModeIR;
  if ((q.apply & a.bit) && (q.long_mask & long_bit<MODE>(a.which))) return;   // part_conv_kernel's query
  // 16-byte accesses need aligned rows (the in-block offsets are multiples of 4 samples by construction)
  const bool vec_in = ((reinterpret_cast<uintptr_t>(in) & 15) == 0), vec_out = ((reinterpret_cast<uintptr_t>(out) & 15) == 0);

  if (!(q.apply & a.bit)) {
    // transform not applied to this query: pass the samples through, still report statistics
    for (int n = blk * FN + tid; n < min(T, (blk + 1) * FN); n += FT) {
      const float v = in[n];
      out[n] = v;
      vmax = fmaxf(vmax, fabsf(v));
      vss += v * v;
    }
  } else {
    const int half = MODE == kModeIR ? 0 : (a.which == 1 ? q.half1 : (a.which == 2 ? q.half2 : q.half3));
    const ConvGeom g = conv_geom(MODE == kModeIR, half, q.ir_len, T);
    const int n0 = blk * g.V;
    if (n0 >= g.n_total) return;  // block-uniform
    const float4* hs = a.hspec + (size_t)qi * (FM / 2);
    // the filter's double pairs are read between the two transforms: pull this thread's 16 into L2 now
#pragma unroll
    for (int d2 = 0; d2 < 8; ++d2) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(hs + fc::dp_index(tid, d2)));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(hs + FM / 4 + fc::dp_index(tid, d2)));
    }
    s.load_tables(tw_g, tid);
    __syncthreads();   // the twiddle tables
    const int s0 = n0 - g.lead;
    // forward pass A straight from global memory: thread tid's 16 quads are the float4 numbers tid + 256 t of the block
    if (vec_in && s0 >= 0 && s0 + FN <= T) {   // interior block: sixteen 16-byte loads in flight per thread
      const float4* src = reinterpret_cast<const float4*>(in + s0) + tid;
      fc::pass_a_fwd_quads(s.re, s.im, s.twa_re, s.twa_im, tid, [&](int t) { return __ldg(src + FT * t); });
    } else {
      fc::pass_a_fwd_quads(s.re, s.im, s.twa_re, s.twa_im, tid, [&](int t) {
        const int n = s0 + 4 * (tid + FT * t);
        if (vec_in && n >= 0 && n + 3 < T) return __ldg(reinterpret_cast<const float4*>(in + n));
        float x[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int ne = n + e;
          if (MODE == kModeIR) x[e] = (ne >= 0 && ne < T) ? __ldg(in + ne) : 0.f;   // zero extension
          else x[e] = __ldg(in + min(max(ne, 0), T - 1));                           // replicate padding (julius)
        }
        return make_float4(x[0], x[1], x[2], x[3]);
      });
    }
    __syncthreads();
    fc::pass_b<false>(s.re, s.im, s.twb_re, s.twb_im, tid); __syncthreads();
    fc::pass_last_fwd(s.re, s.im, tid);                     __syncthreads();
    apply_filter_pairs(s, hs, hs + FM / 4, tid);
    __syncthreads();
    fc::pass_last_inv(s.re, s.im, tid);                     __syncthreads();
    fc::pass_b<true>(s.re, s.im, s.twb_re, s.twb_im, tid);  __syncthreads();
    // the inverse pass A hands over float4s of four consecutive samples c[4 tid + 1024 k ..]: y[n0 + i] = c[off + i], off a
    // multiple of 4, so they go to global memory as they are.  A high-pass needs no second look at x: its filter is
    // delta - lowpass (filter_spectrum_kernel)
    const int n_end = min(g.n_total, n0 + g.V);
    fc::pass_a_inv_quads(s.re, s.im, s.twa_re, s.twa_im, tid, [&](int k, float4 v4) {
      const int i = 4 * tid + 1024 * k - g.off;
      const int n = n0 + i;
      if (i < 0 || n >= n_end) return;
      const float v[4] = {v4.x, v4.y, v4.z, v4.w};
      if (n + 3 < n_end && n + 3 < T) {
        if (want_max) vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(v[0]), fabsf(v[1]))), fmaxf(fabsf(v[2]), fabsf(v[3])));
        if (want_ss) vss = fmaf(v[0], v[0], fmaf(v[1], v[1], fmaf(v[2], v[2], fmaf(v[3], v[3], vss))));
        if (vec_out) *reinterpret_cast<float4*>(out + n) = v4;
        else { out[n] = v[0]; out[n + 1] = v[1]; out[n + 2] = v[2]; out[n + 3] = v[3]; }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (n + e < n_end) {
            vmax = fmaxf(vmax, fabsf(v[e]));
            if (n + e < T) { out[n + e] = v[e]; vss += v[e] * v[e]; }
          }
      }
    });
  }
  if (!want_max) return;   // block-uniform
  vmax = block_max(vmax, s.red, tid);
  if (want_ss) vss = block_sum(vss, s.red, tid);
  if (tid == 0) {
    AugS* st = a.st + qi;
    if (MODE == kModeIR) { atomic_max_pos(&st->max_b, vmax); atomicAdd(&st->ss_b, (double)vss); }
    else atomic_max_pos(&st->max_v, vmax);
  }
}

// ---- stage 4: clip thresholds by exact radix select (torch.quantile, linear interpolation) ----
__device__ __forceinline__ unsigned order_key(float v) {
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_value(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

constexpr int kSelThreads = 256;
constexpr int kSelSample = 512;    // strided sample that places the tail thresholds (each sample costs a DRAM sector)
constexpr int kSelCap = 4096;      // capacity of each tail list
constexpr int kMixStage = 256;     // per-block staging of the tails in mix_kernel

// General exact radix select (any rank): sorted[r] for two ranks at once, 8 bits per pass, then one
// more pass for the successors sorted[r+1].  Fallback of clip_finish_kernel.
__device__ void clip_select_general(const float* __restrict__ z, int T, const int (&r0)[2], unsigned (&hist)[2][256],
                                    unsigned* sh, unsigned (&kout)[2], unsigned (&ksucc)[2]) {
  unsigned* sel_prefix = sh;      // [2]
  unsigned* sel_rank = sh + 2;    // [2]
  unsigned* cnt_le = sh + 4;      // [2]
  unsigned* min_gt = sh + 6;      // [2]
  const int tid = threadIdx.x;
  if (tid < 2) { sel_prefix[tid] = 0; sel_rank[tid] = (unsigned)r0[tid]; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 512; i += kSelThreads) hist[i >> 8][i & 255] = 0;
    __syncthreads();
    const unsigned pre0 = sel_prefix[0], pre1 = sel_prefix[1];
    const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int n = tid; n < T; n += kSelThreads) {
      const unsigned k = order_key(z[n]);
      const unsigned d = (k >> shift) & 255u;
      const bool m0 = (k & himask) == pre0, m1 = (k & himask) == pre1;
      // warp-aggregated shared-memory atomics: audio samples share their leading bits
      const unsigned active = __activemask();
      const unsigned code = (m0 ? 0x100u : 0u) | (m1 ? 0x200u : 0u) | d;
      const unsigned peers = __match_any_sync(active, code);
      if ((int)(__ffs(peers) - 1) == (tid & 31)) {
        const unsigned c = __popc(peers);
        if (m0) atomicAdd(&hist[0][d], c);
        if (m1) atomicAdd(&hist[1][d], c);
      }
    }
    __syncthreads();
    if (tid < 2) {
      unsigned r = sel_rank[tid], acc = 0;
      int d = 0;
      for (; d < 256; ++d) {
        const unsigned c = hist[tid][d];
        if (acc + c > r) break;
        acc += c;
      }
      sel_rank[tid] = r - acc;
      sel_prefix[tid] |= (unsigned)d << shift;
    }
    __syncthreads();
  }
  const unsigned k0 = sel_prefix[0], k1 = sel_prefix[1];
  if (tid < 2) { cnt_le[tid] = 0; min_gt[tid] = 0xffffffffu; }
  __syncthreads();
  unsigned c0 = 0, c1 = 0, g0 = 0xffffffffu, g1 = 0xffffffffu;
  for (int n = tid; n < T; n += kSelThreads) {
    const unsigned k = order_key(z[n]);
    if (k <= k0) ++c0; else g0 = min(g0, k);
    if (k <= k1) ++c1; else g1 = min(g1, k);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
    g0 = min(g0, __shfl_xor_sync(0xffffffffu, g0, o)); g1 = min(g1, __shfl_xor_sync(0xffffffffu, g1, o));
  }
  if ((tid & 31) == 0) {
    atomicAdd(&cnt_le[0], c0); atomicAdd(&cnt_le[1], c1);
    atomicMin(&min_gt[0], g0); atomicMin(&min_gt[1], g1);
  }
  __syncthreads();
  kout[0] = k0; kout[1] = k1;
  for (int i = 0; i < 2; ++i) {
    const bool same = cnt_le[i] >= (unsigned)r0[i] + 2u || r0[i] + 1 > T - 1;
    ksucc[i] = same ? kout[i] : min_gt[i];
  }
}

// ---- filters longer than one block: uniformly partitioned overlap-save ------------------------------
// A low cut-off makes the windowed-sinc FIR arbitrarily long (taps = 2 int(4 sr / fc) + 1: 16 001 taps at
// 4 Hz, more than the signal below 1 Hz) and real room responses run to several seconds.  Such a filter is
// cut into P partitions of S = FN/2 taps; with c[n] = sum_k h[k] xe[n + D - k] (xe = padded input, D = half
// for the centred FIRs, 0 for the causal response),
//     c[n] = sum_p sum_{k<S} h_p[k] xe[n + D - pS - k],
// so output block b (n in [bS, (b+1)S)) is ONE inverse transform of  sum_p H_p * X_{b-p},  where X_j is
// the spectrum of the FN input samples starting at jS + D - S + 1.  Three kernels over the list of long
// queries: partition double pairs (S, R of fftconv_core.cuh), input-block transforms (each computed once, kept
// in HBM as two planes in the transform's digit-reversed order), and the accumulate + inverse + epilogue kernel.
constexpr int kPartS = FN / 2;

struct PartArgs {
  ConvArgs c;
  const int* list;      // [n_long] query indices
  float4* hparts;       // [n_long][p_cap][2][FM/4] double pairs of every partition
  float* xspec;         // [n_long][nb_in_cap][2][FM] planes of every input-block transform
  int p_cap, nb_in_cap;
};

template <int MODE>
__device__ __forceinline__ void part_geometry(const AugQ& q, int which, int T, int& K, int& half, int& D, int& P,
                                              int& n_total) {
  if (MODE == kModeIR) { K = q.ir_len; half = 0; D = 0; n_total = T + K - 1; }
  else { half = which == 1 ? q.half1 : (which == 2 ? q.half2 : q.half3); K = 2 * half + 1; D = half; n_total = T; }
  P = (K + kPartS - 1) / kPartS;
}

// grid (p_cap, n_long): double pairs of partition p (taps [pS, (p+1)S) zero-padded to FN), scaled like filter_spectrum_kernel
template <int MODE>
__global__ void __launch_bounds__(FT, 2) part_filter_kernel(const PartArgs a, const float* __restrict__ tw_g) {
  extern __shared__ __align__(16) float smem_f[];
  const ConvSmem s(smem_f);
  const int tid = threadIdx.x, li = blockIdx.y, qi = a.list[li], p = blockIdx.x;
  const AugQ q = a.c.q[qi];
  int K, half, D, P, n_total;
  part_geometry<MODE>(q, a.c.which, a.c.T, K, half, D, P, n_total);
  if (p >= P) return;
  float c2 = 0.f, argscale = 0.f;
  int hh;
  if (MODE != kModeIR) filter_params(q, a.c.which, hh, c2, argscale);
  s.load_tables(tw_g, tid);
  const float* ir = MODE == kModeIR ? a.c.ir + q.ir_off : nullptr;
  float scale = 1.0f / (float)FM;
  if (MODE != kModeIR) {   // julius normalises by the sum of ALL taps
    float hs = 0.f;
    for (int i = tid; i < K; i += FT) hs += fir_tap(i, half, c2, argscale);
    scale /= block_sum(hs, s.red, tid);
  }
  fill_planes<1>(s, tid, [&](int i4) {
    float h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 4 * i4 + e, i = p * kPartS + j;
      h[e] = 0.f;
      if (j < kPartS && i < K) h[e] = MODE == kModeIR ? ir[i] : fir_tap(i, half, c2, argscale);
    }
    return make_float4(h[0], h[1], h[2], h[3]);
  });
  __syncthreads();
  fft_forward(s, tid);
  float4* hs = a.hparts + ((size_t)li * a.p_cap + p) * (FM / 2);
  store_filter_pairs(s, scale, hs, hs + FM / 4, tid);
}

// grid (nb_in_cap, n_long): transform of the packed input block j - (P-1)
template <int MODE>
__global__ void __launch_bounds__(FT, 2) part_forward_kernel(const PartArgs a, const float* __restrict__ tw_g) {
  extern __shared__ __align__(16) float smem_f[];
  const ConvSmem s(smem_f);
  const int tid = threadIdx.x, li = blockIdx.y, qi = a.list[li], jb = blockIdx.x;
  const AugQ q = a.c.q[qi];
  int K, half, D, P, n_total;
  part_geometry<MODE>(q, a.c.which, a.c.T, K, half, D, P, n_total);
  const int nb_out = (n_total + kPartS - 1) / kPartS;
  if (jb >= nb_out + P - 1) return;
  const int T = a.c.T;
  const float* __restrict__ in = a.c.in + (int64_t)qi * a.c.in_stride;
  const int64_t s0 = (int64_t)(jb - (P - 1)) * kPartS + D - kPartS + 1;
  s.load_tables(tw_g, tid);
  fill_planes<1>(s, tid, [&](int i4) {
    float x[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t n = s0 + 4 * i4 + e;
      if (MODE == kModeIR) x[e] = (n >= 0 && n < T) ? __ldg(in + n) : 0.f;                  // zero extension
      else x[e] = __ldg(in + (n < 0 ? 0 : (n > T - 1 ? T - 1 : n)));                         // replicate padding (julius)
    }
    return make_float4(x[0], x[1], x[2], x[3]);
  });
  __syncthreads();
  fft_forward(s, tid);
  float* zg = a.xspec + ((size_t)li * a.nb_in_cap + jb) * (2 * FM);
  for (int i = tid; i < FM / 2; i += FT) {
    const int o = fc::padi(2 * i);
    *reinterpret_cast<float2*>(zg + 2 * i) = *reinterpret_cast<const float2*>(s.re + o);
    *reinterpret_cast<float2*>(zg + FM + 2 * i) = *reinterpret_cast<const float2*>(s.im + o);
  }
}

// grid (nb_out_max, n_long): sum_p (filter pairs of p) x (transform of block b - p) -> inverse transform -> the
// same epilogue as fftconv_kernel
template <int MODE>
__global__ void __launch_bounds__(FT, 2) part_conv_kernel(const PartArgs a, const float* __restrict__ tw_g) {
  extern __shared__ __align__(16) float smem_f[];
  const ConvSmem s(smem_f);
  const int tid = threadIdx.x, li = blockIdx.y, qi = a.list[li], b = blockIdx.x;
  const AugQ q = a.c.q[qi];
  int K, half, D, P, n_total;
  part_geometry<MODE>(q, a.c.which, a.c.T, K, half, D, P, n_total);
  const int nb_out = (n_total + kPartS - 1) / kPartS;
  if (b >= nb_out) return;
  const int T = a.c.T;
  const float* __restrict__ in = a.c.in + (int64_t)qi * a.c.in_stride;
  float* __restrict__ out = a.c.out + (int64_t)qi * T;
  s.load_tables(tw_g, tid);
  const float4* __restrict__ hq = a.hparts + (size_t)li * a.p_cap * (FM / 2);
  const float* __restrict__ zq = a.xspec + (size_t)li * a.nb_in_cap * (2 * FM);
  {
    const int m = tid, mbar = fc::partner_block(m);
    float sw, cw;
    sincospif((float)fc::block_c(m) * (1.0f / (float)FM), &sw, &cw);
    c2 wv = fc::dp_twiddle0(cw, sw);
#pragma unroll 1
    for (int d2 = 0; d2 < 8; ++d2) {
      if (d2) wv = fc::dp_twiddle_next(wv);
      if (m == 0 && d2 == 0) {
        fc::SpecAcc acc = fc::spec_acc_zero();
        for (int p = 0; p < P; ++p) {
          const float* z = zq + (size_t)(b - p + P - 1) * (2 * FM);
          const float4 vs = __ldg(hq + (size_t)p * (FM / 2)), vr = __ldg(hq + (size_t)p * (FM / 2) + FM / 4);
          fc::Specials sp;
          sp.a0 = vs.x; sp.a1 = vs.y; sp.bre = vs.z; sp.bim = vs.w;
          sp.sre = vr.x; sp.sim = vr.y; sp.rre = vr.z; sp.rim = vr.w;
          const float z8[8] = {z[0], z[FM], z[1], z[FM + 1], z[16], z[FM + 16], z[17], z[FM + 17]};
          fc::specials_accumulate(acc, z8, sp);
        }
        fc::specials_finish(s.re, s.im, acc);
        continue;
      }
      // unpadded plane offsets of the two slots, and their padded shared-memory twins
      const int ua = 32 * m + 2 * d2, ub = 32 * mbar + 2 * ((m == 0 ? 16 : 15) - d2);
      const int i = fc::dp_index(m, d2);
      const float2 zero = make_float2(0.f, 0.f);
      c2 ey = fc::mk(zero, zero), dy = fc::mk(zero, zero);
      for (int p = 0; p < P; ++p) {
        const float* z = zq + (size_t)(b - p + P - 1) * (2 * FM);
        const float4 vs = __ldg(hq + (size_t)p * (FM / 2) + i), vr = __ldg(hq + (size_t)p * (FM / 2) + FM / 4 + i);
        c2 ep, f;
        fc::dp_split(fc::cld(z, z + FM, ua), fc::cld(z, z + FM, ub), wv, ep, f);
        const c2 S = fc::mk(fc::f2(vs.x, vs.y), fc::f2(vs.z, vs.w)), R = fc::mk(fc::f2(vr.x, vr.y), fc::f2(vr.z, vr.w));
        ey = fc::cmac(R, f, fc::cmac(S, ep, ey));
        dy = fc::cmac(S, f, fc::cmac(R, ep, dy));
      }
      c2 za, zb;
      fc::dp_repack(ey, dy, wv, za, zb);
      fc::cst(s.re, s.im, fc::padi(ua), za);
      fc::cst(s.re, s.im, fc::padi(ub), zb);
    }
  }
  __syncthreads();
  fft_inverse(s, tid);
  float vmax = 0.f, vss = 0.f;
  const int n_end = min(n_total, (b + 1) * kPartS);
  for (int n = b * kPartS + tid; n < n_end; n += FT) {
    const int jj = n - b * kPartS + kPartS - 1;
    const int o = fc::padi(jj >> 1);
    const float c = (jj & 1) ? s.im[o] : s.re[o];
    const float v = MODE == kModeHP ? in[n] - c : c;
    vmax = fmaxf(vmax, fabsf(v));
    if (n < T) { out[n] = v; vss += v * v; }
  }
  vmax = block_max(vmax, s.red, tid);
  vss = block_sum(vss, s.red, tid);
  if (tid == 0) {
    AugS* st = a.c.st + qi;
    if (MODE == kModeHP && a.c.which == 1) { atomic_max_pos(&st->max_a, vmax); atomicAdd(&st->ss_a, (double)vss); }
    else if (MODE == kModeIR) { atomic_max_pos(&st->max_b, vmax); atomicAdd(&st->ss_b, (double)vss); }
    else if (MODE == kModeHP && a.c.which == 3) atomic_max_pos(&st->max_v, vmax);
  }
}

// Exact selection in a shared-memory array: k = sorted(a)[r] and its successor sorted(a)[min(r+1, n-1)],
// by four 8-bit radix passes (256-bin histogram, digit found by one warp) and one counting pass.  All
// threads of the block call it; `hist` is 256 words, `sh` 4 words of shared scratch.
__device__ void smem_select(const unsigned* a, int n, int r, unsigned* hist, unsigned* sh, int tid, unsigned& k_out,
                            unsigned& k_succ) {
  unsigned prefix = 0, rank = (unsigned)r;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    __syncthreads();
    for (int i = tid; i < 256; i += kSelThreads) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kSelThreads) {
      const unsigned k = a[i];
      if ((k & himask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid < 32) {   // digit d with cum(d-1) <= rank < cum(d); lane owns bins 8*lane .. 8*lane+7
      unsigned c[8], tot = 0;
#pragma unroll
      for (int b = 0; b < 8; ++b) { c[b] = hist[8 * tid + b]; tot += c[b]; }
      unsigned incl = tot;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned v = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= o) incl += v; }
      unsigned acc = incl - tot;
      if (acc <= rank && rank < incl) {   // exactly one lane
        int d = 0;
        for (; d < 8; ++d) { if (acc + c[d] > rank) break; acc += c[d]; }
        sh[0] = prefix | ((unsigned)(8 * tid + d) << shift);
        sh[1] = rank - acc;
      }
    }
    __syncthreads();
    prefix = sh[0];
    rank = sh[1];
  }
  __syncthreads();
  if (tid == 0) { sh[2] = 0u; sh[3] = 0xffffffffu; }
  __syncthreads();
  unsigned c_le = 0, g = 0xffffffffu;
  for (int i = tid; i < n; i += kSelThreads) {
    const unsigned k = a[i];
    if (k <= prefix) ++c_le; else g = min(g, k);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    c_le += __shfl_xor_sync(0xffffffffu, c_le, o);
    g = min(g, __shfl_xor_sync(0xffffffffu, g, o));
  }
  if ((tid & 31) == 0) { atomicAdd(&sh[2], c_le); atomicMin(&sh[3], g); }
  __syncthreads();
  k_out = prefix;
  k_succ = (sh[2] >= (unsigned)r + 2u || r + 1 > n - 1) ? prefix : sh[3];
  __syncthreads();
}

// Clip thresholds = the order statistics sorted[r], sorted[r+1] of torch.quantile's two ranks over z.
// Clipping percentiles are small (p <= 0.01 in AugmentFP), so both ranks sit in the tails and no pass over
// z is spent on them:
//   clip_sample_kernel  (before the mix) evaluates z at a strided sample, sorts it and places two tail
//                       thresholds that bracket the wanted ranks with a > 4 sigma margin;
//   mix_kernel          lists the few hundred z beyond the thresholds while it writes z, and counts them;
//   clip_finish_kernel  checks the counts (the proof that the lists hold the wanted ranks), sorts the lists
//                       and interpolates.  When the check fails (ranks not in the tails, pathological data)
//                       the general radix select runs over z instead: always exact.
struct ClipPlan {
  float pos[2];
  int r0[2], need_lo, need_hi, s_lo, s_hi;
  bool fast;
};
__device__ __forceinline__ ClipPlan clip_plan(const AugQ& q, int T) {
  ClipPlan p;
  // torch.quantile: rank = q * (n - 1) in float32, lo = floor(rank)
  const float q_hi = 1.0f - q.q_lo;
  p.pos[0] = q.q_lo * (float)(T - 1);
  p.pos[1] = q_hi * (float)(T - 1);
  p.r0[0] = min((int)floorf(p.pos[0]), T - 1);
  p.r0[1] = min((int)floorf(p.pos[1]), T - 1);
  p.need_lo = min(p.r0[0] + 2, T);   // smallest elements needed: sorted[0 .. r0[0]+1]
  p.need_hi = T - p.r0[1];           // largest elements needed: sorted[r0[1] .. T-1]
  p.s_lo = p.s_hi = 0;
  p.fast = T >= 4 * kSelSample;
  if (p.fast) {
    // sample order statistic whose full-data count exceeds `need` with > 4 sigma margin
    const float f = (float)kSelSample / (float)T;
    const float m_lo = p.need_lo * f, m_hi = p.need_hi * f;
    p.s_lo = (int)ceilf(m_lo + 4.f * sqrtf(m_lo) + 6.f);
    p.s_hi = (int)ceilf(m_hi + 4.f * sqrtf(m_hi) + 6.f);
    p.fast = p.s_lo < kSelSample / 8 && p.s_hi < kSelSample / 8;   // expected list length stays well below kSelCap
  }
  return p;
}

// the scale mix_kernel applies to the noise: calculate_rms of the (peak-normalised) stage-2 output over 10^(snr/20)
__device__ __forceinline__ float mix_noise_scale(const AugQ& q, const AugS& s, int T, float peak) {
  const float rms = sqrtf((float)(s.ss_b / (double)T)) / peak;
  return (q.apply & MFPA_AUG_NOISE) ? rms / q.snr_div : 0.f;
}

__global__ void __launch_bounds__(kSelThreads) clip_sample_kernel(const float* __restrict__ in, const float* __restrict__ noise,
                                                                  const AugQ* __restrict__ qs, AugS* st, int T) {
  __shared__ unsigned sample[kSelSample];
  const int qi = blockIdx.x, tid = threadIdx.x;
  const AugQ q = qs[qi];
  AugS* s = st + qi;
  if (!(q.apply & MFPA_AUG_CLIP)) return;
  const ClipPlan p = clip_plan(q, T);
  if (!p.fast) {   // nothing is listed; clip_finish_kernel takes the general path
    if (tid == 0) { s->t_lo_key = 0u; s->t_hi_key = 0xffffffffu; }
    return;
  }
  const bool ir_on = q.apply & MFPA_AUG_IR, nz_on = q.apply & MFPA_AUG_NOISE;
  const float peak = ir_on ? s->max_b : 1.f;
  const float ns = mix_noise_scale(q, *s, T, peak);
  const int64_t row = (int64_t)qi * T;
  for (int i = tid; i < kSelSample; i += kSelThreads) {
    const int n = (int)(((int64_t)i * T) / kSelSample);
    float v = in[row + n];
    if (ir_on) v *= 1.0f / peak;
    if (nz_on) v = fmaf(ns, noise[row + n], v);
    sample[i] = order_key(v);
  }
  __shared__ unsigned hist[256], sh[4];
  unsigned k_lo, k_hi, unused;
  smem_select(sample, kSelSample, p.s_lo, hist, sh, tid, k_lo, unused);
  smem_select(sample, kSelSample, kSelSample - 1 - p.s_hi, hist, sh, tid, k_hi, unused);
  if (tid == 0) { s->t_lo_key = k_lo; s->t_hi_key = k_hi; }
}

__global__ void __launch_bounds__(kSelThreads) clip_finish_kernel(const float* __restrict__ z_all, const unsigned* __restrict__ lists,
                                                                  const AugQ* __restrict__ qs, AugS* st, int T) {
  __shared__ unsigned lo_list[kSelCap], hi_list[kSelCap];
  __shared__ unsigned hist[2][256];   // general path only
  __shared__ unsigned sh[8];
  const int qi = blockIdx.x, tid = threadIdx.x;
  const AugQ q = qs[qi];
  AugS* s = st + qi;
  if (!(q.apply & MFPA_AUG_CLIP)) {
    if (tid == 0) { s->lo = -INFINITY; s->hi = INFINITY; }
    return;
  }
  const ClipPlan p = clip_plan(q, T);
  const int c_lo = (int)s->cnt_lo, c_hi = (int)s->cnt_hi;
  const bool fast = p.fast && c_lo >= p.need_lo && c_lo <= kSelCap && c_hi >= p.need_hi && c_hi <= kSelCap;
  unsigned kout[2], ksucc[2];
  if (fast) {
    const unsigned* gl = lists + (size_t)qi * 2 * kSelCap;
    for (int i = tid; i < c_lo; i += kSelThreads) lo_list[i] = gl[i];
    for (int i = tid; i < c_hi; i += kSelThreads) hi_list[i] = gl[kSelCap + i];
    // sorted[r0[0]] is rank r0[0] of the low list; sorted[r0[1]] is rank c_hi - (T - r0[1]) of the high list
    smem_select(lo_list, c_lo, p.r0[0], hist[0], sh, tid, kout[0], ksucc[0]);
    smem_select(hi_list, c_hi, c_hi - (T - p.r0[1]), hist[0], sh, tid, kout[1], ksucc[1]);
  } else {
    clip_select_general(z_all + (int64_t)qi * T, T, p.r0, hist, sh, kout, ksucc);
  }
  if (tid == 0) {
    // elementwise map into the gained domain, monotone so the order statistics carry over
    const bool nz_on = q.apply & MFPA_AUG_NOISE;
    const float peak = nz_on ? s->max_z : 1.f;
    const float g = (q.apply & MFPA_AUG_GAIN) ? q.gain : 1.f;
    float thr[2];
    for (int i = 0; i < 2; ++i) {
      const float va = key_value(kout[i]), vb = key_value(ksucc[i]);
      const float fa = (nz_on ? va / peak : va) * g, fb = (nz_on ? vb / peak : vb) * g;
      const float w = p.pos[i] - (float)p.r0[i];
      thr[i] = fa + (fb - fa) * w;
    }
    s->lo = thr[0];
    s->hi = thr[1];
  }
}

// ---- Clipping over a pooled sub-batch (clipping.py:67-101 with B > 1) ---------------------------------------
// torch.quantile(samples[:, 0, :], q) with a VECTOR q and no dim flattens its input: every selected row gets the
// quantiles of ITS q over the samples of ALL selected rows.  Reproduced when
// MFPA_OPT_CLIP_POOLED is set (the drop-in's batch_augment): the gained samples of the selected rows are written as
// order keys into one array, sorted (bitonic, global memory: this is the reference's rarely used batch path, not the
// throughput path), and every row interpolates its two thresholds in it.
__device__ __forceinline__ float clip_pre_scale(const AugQ& q, const AugS& s) {   // what clip_lpf_kernel multiplies z by
  const float g = (q.apply & MFPA_AUG_GAIN) ? q.gain : 1.f;
  return (q.apply & MFPA_AUG_NOISE) ? g / s.max_z : g;
}
__global__ void __launch_bounds__(256) pool_fill_kernel(const float* __restrict__ z_all, const int* __restrict__ sel, int T,
                                                        const AugQ* __restrict__ qs, const AugS* __restrict__ st,
                                                        unsigned* __restrict__ keys) {
  const int r = blockIdx.y, qi = sel[r];
  const float pre = clip_pre_scale(qs[qi], st[qi]);
  const float* z = z_all + (int64_t)qi * T;
  for (int n = blockIdx.x * 256 + threadIdx.x; n < T; n += gridDim.x * 256) keys[(int64_t)r * T + n] = order_key(z[n] * pre);
}
__global__ void __launch_bounds__(256) pool_pad_kernel(unsigned* __restrict__ keys, int64_t n, int64_t n_pad) {
  for (int64_t i = n + (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * 256) keys[i] = 0xffffffffu;
}
__global__ void __launch_bounds__(256) bitonic_step_kernel(unsigned* __restrict__ keys, int64_t n_pad, int64_t k, int64_t j) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n_pad; i += (int64_t)gridDim.x * 256) {
    const int64_t ixj = i ^ j;
    if (ixj > i) {
      const unsigned a = keys[i], b = keys[ixj];
      if ((a > b) == ((i & k) == 0)) { keys[i] = b; keys[ixj] = a; }
    }
  }
}
__global__ void pool_thresholds_kernel(const unsigned* __restrict__ sorted, int64_t n, const int* __restrict__ sel, int n_sel,
                                       const AugQ* __restrict__ qs, AugS* st) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_sel) return;
  const int qi = sel[r];
  const float q_lo = qs[qi].q_lo, q_hi = 1.0f - q_lo;
  float thr[2];
#pragma unroll
  for (int i = 0; i < 2; ++i) {   // torch.quantile: rank = q (n - 1) in float32, lerp between floor and ceil
    const float pos = (i ? q_hi : q_lo) * (float)(n - 1);
    const int64_t below = min((int64_t)floorf(pos), n - 1), above = min((int64_t)ceilf(pos), n - 1);
    const float fa = key_value(sorted[below]), fb = key_value(sorted[above]);
    thr[i] = fa + (fb - fa) * (pos - (float)below);
  }
  st[qi].lo = thr[0];
  st[qi].hi = thr[1];
}

// stage 3: z = in/peak_b + (rms/10^(snr/20)) * noise ; statistics: peak of z, and the tails of z beyond
// the thresholds clip_sample_kernel placed (lists[query][0 = low, 1 = high][kSelCap] order keys + counts)
__global__ void __launch_bounds__(256) mix_kernel(const float* __restrict__ in, const float* __restrict__ noise,
                                                  float* __restrict__ z, const AugQ* __restrict__ qs, AugS* st, int T,
                                                  unsigned* __restrict__ lists) {
  __shared__ float red[8];
  const int qi = blockIdx.y, tid = threadIdx.x;
  const AugQ q = qs[qi];
  AugS* s = st + qi;
  const bool ir_on = q.apply & MFPA_AUG_IR, nz_on = q.apply & MFPA_AUG_NOISE;
  const float peak = ir_on ? s->max_b : 1.f;
  const float ns = mix_noise_scale(q, *s, T, peak);
  const bool clip_on = q.apply & MFPA_AUG_CLIP;
  const unsigned t_lo = clip_on ? s->t_lo_key : 0u, t_hi = clip_on ? s->t_hi_key : 0xffffffffu;
  unsigned* ql = lists + (size_t)qi * 2 * kSelCap;
  // tails are staged per block (shared-memory atomics) and appended to the query's lists with one global
  // atomic per block and side; a block that outgrows its stage appends the excess directly
  __shared__ unsigned stage[2][kMixStage];
  __shared__ unsigned stage_n[2], stage_base[2];
  if (tid < 2) stage_n[tid] = 0;
  __syncthreads();
  auto tail_side = [&](unsigned k, int side) {
    const unsigned slot = atomicAdd(&stage_n[side], 1u);
    if (slot < kMixStage) { stage[side][slot] = k; return; }
    const unsigned g = atomicAdd(side ? &s->cnt_hi : &s->cnt_lo, 1u);
    if (g < kSelCap) ql[side * kSelCap + g] = k;
  };
  // the order keys are monotone in the value, so the common case is two float compares
  const float f_lo = clip_on ? key_value(t_lo) : -INFINITY, f_hi = clip_on ? key_value(t_hi) : INFINITY;
  auto tail = [&](float v) {
    if (v <= f_lo || v >= f_hi) {   // ~1 % of the samples
      if (!clip_on) return;
      const unsigned k = order_key(v);
      if (k <= t_lo) tail_side(k, 0);
      if (k >= t_hi) tail_side(k, 1);
    }
  };
  float vmax = 0.f;
  const int64_t row = (int64_t)qi * T;
  // x / peak as x * (1 / peak): one rounding more than the reference's division, four IEEE divisions less per
  // 16 bytes (this kernel is issue-bound, not bandwidth-bound, with them)
  const float rpeak = ir_on ? 1.0f / peak : 1.f;
  for (int n = blockIdx.x * blockDim.x * 4 + tid * 4; n < T; n += gridDim.x * blockDim.x * 4) {
    if (n + 3 < T && ((row + n) & 3) == 0) {
      float4 v = *reinterpret_cast<const float4*>(in + row + n);
      if (ir_on) { v.x *= rpeak; v.y *= rpeak; v.z *= rpeak; v.w *= rpeak; }
      if (nz_on) {
        const float4 b = *reinterpret_cast<const float4*>(noise + row + n);
        v.x = fmaf(ns, b.x, v.x); v.y = fmaf(ns, b.y, v.y); v.z = fmaf(ns, b.z, v.z); v.w = fmaf(ns, b.w, v.w);
      }
      *reinterpret_cast<float4*>(z + row + n) = v;
      vmax = fmaxf(fmaxf(vmax, fmaxf(fabsf(v.x), fabsf(v.y))), fmaxf(fabsf(v.z), fabsf(v.w)));
      tail(v.x); tail(v.y); tail(v.z); tail(v.w);
    } else {
      for (int m = n; m < min(n + 4, T); ++m) {
        float v = in[row + m];
        if (ir_on) v *= rpeak;
        if (nz_on) v = fmaf(ns, noise[row + m], v);
        z[row + m] = v;
        vmax = fmaxf(vmax, fabsf(v));
        tail(v);
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  if ((tid & 31) == 0) red[tid >> 5] = vmax;
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < 8; ++w) vmax = fmaxf(vmax, red[w]);
    atomic_max_pos(&s->max_z, vmax);
  }
  if (tid < 2) {
    const unsigned n = min(stage_n[tid], (unsigned)kMixStage);
    stage_base[tid] = n ? atomicAdd(tid ? &s->cnt_hi : &s->cnt_lo, n) : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const unsigned n = min(stage_n[side], (unsigned)kMixStage), base = stage_base[side];
    for (unsigned i = tid; i < n; i += 256)
      if (base + i < kSelCap) ql[side * kSelCap + base + i] = stage[side][i];
  }
}

// stage 3b-5: w = clamp((z/peak_z)*gain, lo, hi); u = lowpass(w) with a short FIR (direct form).
// The tile starts round_up(half, 4) samples before its first output so that 16-byte global loads,
// shared-memory float4 reads and 16-byte stores all stay aligned; the taps are stored shifted by the
// same amount and zero-padded to whole float4 groups.  Each thread produces four consecutive outputs
// from a sliding pair of float4 tile reads: 16 FMAs per two 16-byte shared loads.
constexpr int kLpMaxHalf = 64;
constexpr int kLpTile = 1024;
constexpr int kLpTapsPad = 2 * kLpMaxHalf + 8;
constexpr int kLpTileLen = kLpTile + kLpTapsPad + 8;
__global__ void __launch_bounds__(256) clip_lpf_kernel(const float* __restrict__ z_all, float* __restrict__ u_all,
                                                       const AugQ* __restrict__ qs, const AugS* __restrict__ st, int T,
                                                       int materialise_only) {
  __shared__ __align__(16) float taps[kLpTapsPad];
  __shared__ __align__(16) float tile[kLpTileLen];
  __shared__ float red[8];
  const int qi = blockIdx.y, tid = threadIdx.x;
  const AugQ q = qs[qi];
  const AugS s = st[qi];
  const bool nz_on = q.apply & MFPA_AUG_NOISE;
  const float peak = nz_on ? s.max_z : 1.f;
  const float g = (q.apply & MFPA_AUG_GAIN) ? q.gain : 1.f;
  const bool lp_on = (q.apply & MFPA_AUG_LPF) && !materialise_only;
  const int half = lp_on ? q.half2 : 0;
  const int half_al = (half + 3) & ~3, shift = half_al - half;
  const int groups = lp_on ? (2 * half + 1 + shift + 3) >> 2 : 0;   // float4 groups of shifted taps
  const int64_t row = (int64_t)qi * T;
  const float* z = z_all + row;
  float* u = u_all + row;
  const bool vec_ok = (row & 3) == 0 && (reinterpret_cast<uintptr_t>(z_all) & 15) == 0 &&
                      (reinterpret_cast<uintptr_t>(u_all) & 15) == 0;
  float hsum = 1.f;
  if (lp_on) {
    float hs = 0.f;
    for (int i = tid; i < kLpTapsPad; i += 256) {
      const int k = i - shift;
      const float h = (k >= 0 && k <= 2 * half) ? fir_tap(k, half, q.c2x2, q.arg2) : 0.f;
      taps[i] = h;
      hs += h;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) hs += __shfl_xor_sync(0xffffffffu, hs, o);
    if ((tid & 31) == 0) red[tid >> 5] = hs;
    __syncthreads();
    hsum = 0.f;
    for (int w = 0; w < 8; ++w) hsum += red[w];
  }
  const float inv = 1.0f / hsum;
  const float pre = nz_on ? g / peak : g;   // (z / peak) * gain in one multiply
  auto shape = [&](float v) { return fminf(fmaxf(v * pre, s.lo), s.hi); };
  const int fill4 = (kLpTile + 4 * groups + 4) >> 2;   // float4 slots of the tile that are read
  const float4* tile4 = reinterpret_cast<const float4*>(tile);
  const float4* taps4 = reinterpret_cast<const float4*>(taps);
  for (int n0 = blockIdx.x * kLpTile; n0 < T; n0 += gridDim.x * kLpTile) {
    __syncthreads();
    for (int i4 = tid; i4 < fill4; i4 += 256) {
      const int n = n0 - half_al + 4 * i4;
      float4 v;
      if (vec_ok && n >= 0 && n + 3 < T) {
        v = *reinterpret_cast<const float4*>(z + n);
      } else {   // replicate padding (julius pads with the edge sample)
        v.x = z[min(max(n, 0), T - 1)];
        v.y = z[min(max(n + 1, 0), T - 1)];
        v.z = z[min(max(n + 2, 0), T - 1)];
        v.w = z[min(max(n + 3, 0), T - 1)];
      }
      v.x = shape(v.x); v.y = shape(v.y); v.z = shape(v.z); v.w = shape(v.w);
      reinterpret_cast<float4*>(tile)[i4] = v;
    }
    __syncthreads();
    const int n = n0 + 4 * tid;
    if (n < T) {
      float4 c = tile4[tid], r;
      if (lp_on) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int m = 0; m < groups; ++m) {
          const float4 d = tile4[tid + m + 1], h = taps4[m];
          a0 = fmaf(h.x, c.x, a0); a0 = fmaf(h.y, c.y, a0); a0 = fmaf(h.z, c.z, a0); a0 = fmaf(h.w, c.w, a0);
          a1 = fmaf(h.x, c.y, a1); a1 = fmaf(h.y, c.z, a1); a1 = fmaf(h.z, c.w, a1); a1 = fmaf(h.w, d.x, a1);
          a2 = fmaf(h.x, c.z, a2); a2 = fmaf(h.y, c.w, a2); a2 = fmaf(h.z, d.x, a2); a2 = fmaf(h.w, d.y, a2);
          a3 = fmaf(h.x, c.w, a3); a3 = fmaf(h.y, d.x, a3); a3 = fmaf(h.z, d.y, a3); a3 = fmaf(h.w, d.z, a3);
          c = d;
        }
        r = make_float4(a0 * inv, a1 * inv, a2 * inv, a3 * inv);
      } else {
        r = c;
      }
      if (vec_ok && n + 3 < T) {
        *reinterpret_cast<float4*>(u + n) = r;
      } else {
        u[n] = r.x;
        if (n + 1 < T) u[n + 1] = r.y;
        if (n + 2 < T) u[n + 2] = r.z;
        if (n + 3 < T) u[n + 3] = r.w;
      }
    }
  }
}

// stage 7 (or a plain copy when final_norm is off / the peak is 0)
__global__ void __launch_bounds__(256) norm_kernel(const float* __restrict__ v, float* __restrict__ out, const AugQ* __restrict__ qs,
                                                   const AugS* __restrict__ st, int T, int normalise) {
  const int qi = blockIdx.y;
  const float peak = st[qi].max_v;
  const bool on = normalise && (qs[qi].apply & MFPA_AUG_NORM) && peak > 0.f;
  const int64_t row = (int64_t)qi * T;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < T; n += gridDim.x * blockDim.x) {
    const float x = v[row + n];
    out[row + n] = on ? x / peak : x;
  }
}

// ---- noise rows from a device-resident bank (random_background, background_noise.py:64-141) ----
struct NoiseStat { double ss; };   // per piece: sum of squares of the (paired) piece

// grid (n_pieces, chunks): copy the piece (or the mean of a mix-up pair) into its row, sum of squares per piece
__global__ void __launch_bounds__(256) noise_gather_kernel(const float* __restrict__ bank, const mfpa_noise_piece* __restrict__ pieces,
                                                           int T, float* __restrict__ out, double* __restrict__ piece_ss) {
  __shared__ float red[8];
  const mfpa_noise_piece p = pieces[blockIdx.x];
  const float* a = bank + p.src_a;
  const float* b = p.src_b >= 0 ? bank + p.src_b : nullptr;
  float* o = out + (int64_t)p.query * T + p.dst;
  float ss = 0.f;
  for (int i = blockIdx.y * 256 + threadIdx.x; i < p.len; i += gridDim.y * 256) {
    float v = __ldg(a + i);
    if (b) v = (v + __ldg(b + i)) / 2.0f;   // samplePairing (:11-12)
    o[i] = v;
    ss += v * v;
  }
#pragma unroll
  for (int k = 16; k; k >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, k);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(piece_ss + blockIdx.x, (double)t);
  }
}

// one thread per piece: sum of squares of the row after every piece was RMS-normalised
__global__ void noise_total_kernel(const mfpa_noise_piece* __restrict__ pieces, int n_pieces, const double* __restrict__ piece_ss,
                                   double* __restrict__ row_ss) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pieces) return;
  const mfpa_noise_piece p = pieces[i];
  const float rms = sqrtf((float)(piece_ss[i] / (double)p.len));
  const float s = 1.0f / (rms + 1e-8f);
  atomicAdd(row_ss + p.query, piece_ss[i] * (double)s * (double)s);
}

// grid (n_pieces, chunks): x / (rms_piece + 1e-8) / (rms_row + 1e-8)   (rms_normalize twice, utils.py:189-205)
__global__ void __launch_bounds__(256) noise_scale_kernel(const mfpa_noise_piece* __restrict__ pieces, int T,
                                                          const double* __restrict__ piece_ss, const double* __restrict__ row_ss,
                                                          float* __restrict__ out) {
  const mfpa_noise_piece p = pieces[blockIdx.x];
  const float rms_p = sqrtf((float)(piece_ss[blockIdx.x] / (double)p.len));
  const float rms_r = sqrtf((float)(row_ss[p.query] / (double)T));
  const float dp = rms_p + 1e-8f, dr = rms_r + 1e-8f;
  float* o = out + (int64_t)p.query * T + p.dst;
  for (int i = blockIdx.y * 256 + threadIdx.x; i < p.len; i += gridDim.y * 256) o[i] = (o[i] / dp) / dr;
}

}  // namespace

int launch_noise_assemble(mfpa_ctx* ctx, const float* bank, int64_t bank_len, const mfpa_noise_piece* pieces, int n_pieces,
                          int B, int T, float* out, cudaStream_t st, bool pieces_pinned) {
  // validate on the host: sources inside the bank, every row tiled exactly once
  std::vector<int64_t> covered((size_t)B, 0);
  int max_len = 1;
  for (int i = 0; i < n_pieces; ++i) {
    const mfpa_noise_piece& p = pieces[i];
    MFPA_REQUIRE(p.query >= 0 && p.query < B && p.len >= 1 && p.dst >= 0 && (int64_t)p.dst + p.len <= T,
                 "noise_assemble: piece %d: row %d, dst %d, len %d outside [0, %d) x [0, %d)", i, p.query, p.dst, p.len, B, T);
    MFPA_REQUIRE(p.src_a >= 0 && p.src_a + p.len <= bank_len && (p.src_b < 0 || p.src_b + p.len <= bank_len),
                 "noise_assemble: piece %d reads outside the bank (%lld samples)", i, (long long)bank_len);
    covered[p.query] += p.len;
    max_len = p.len > max_len ? p.len : max_len;
  }
  for (int q = 0; q < B; ++q)
    MFPA_REQUIRE(covered[q] == T, "noise_assemble: row %d is covered by %lld samples, not %d", q, (long long)covered[q], T);
  const size_t bytes = sizeof(mfpa_noise_piece) * (size_t)n_pieces + sizeof(double) * ((size_t)n_pieces + B);
  if (ctx->aug_noise.reserve(bytes)) return MFPA_ENOMEM;
  double* piece_ss = (double*)ctx->aug_noise.ptr;
  double* row_ss = piece_ss + n_pieces;
  mfpa_noise_piece* dp = (mfpa_noise_piece*)(row_ss + B);
  if (pieces_pinned) {   // the host pipeline's staging buffer: pulled by a kernel, off the copy engine's queue
    if (int e = launch_pull(ctx, dp, pieces, sizeof(mfpa_noise_piece) * (size_t)n_pieces, st)) return e;
  } else {
    MFPA_CUDA(cudaMemcpyAsync(dp, pieces, sizeof(mfpa_noise_piece) * (size_t)n_pieces, cudaMemcpyHostToDevice, st));
    MFPA_CUDA(cudaStreamSynchronize(st));   // pieces_host is caller-owned pageable memory
  }
  MFPA_CUDA(cudaMemsetAsync(piece_ss, 0, sizeof(double) * ((size_t)n_pieces + B), st));
  const dim3 grid((unsigned)n_pieces, (unsigned)((max_len + 4095) / 4096));   // pieces on x: no 65535 limit
  noise_gather_kernel<<<grid, 256, 0, st>>>(bank, dp, T, out, piece_ss);
  noise_total_kernel<<<(n_pieces + 255) / 256, 256, 0, st>>>(dp, n_pieces, piece_ss, row_ss);
  noise_scale_kernel<<<grid, 256, 0, st>>>(dp, T, piece_ss, row_ss, out);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

static int aug_init_tables(mfpa_ctx* ctx) {
  if ((const float*)ctx->aug_tw_dev) return MFPA_OK;
  // planar tables of fftconv_core.cuh: exp(-2 pi i e / FM), e < 512, then exp(-2 pi i e / 512), e < 32
  static float tw[kTwFloats];
  const double pi = 3.14159265358979323846;
  for (int e = 0; e < fc::kTwA; ++e) {
    tw[e] = (float)cos(2 * pi * e / FM);
    tw[fc::kTwA + e] = (float)-sin(2 * pi * e / FM);
  }
  for (int e = 0; e < fc::kTwB; ++e) {
    tw[2 * fc::kTwA + e] = (float)cos(2 * pi * e / 512);
    tw[2 * fc::kTwA + fc::kTwB + e] = (float)-sin(2 * pi * e / 512);
  }
  MFPA_CUDA(cudaMalloc(&ctx->aug_tw_dev, sizeof(tw)));
  MFPA_CUDA(cudaMemcpy(ctx->aug_tw_dev, tw, sizeof(tw), cudaMemcpyHostToDevice));
  MFPA_CUDA(cudaFuncSetAttribute(fftconv_kernel<kModeHP, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(fftconv_kernel<kModeIR, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(fftconv_kernel<kModeHP, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(fftconv_kernel<kModeIR, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(filter_spectrum_kernel<kModeHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(filter_spectrum_kernel<kModeIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(part_filter_kernel<kModeHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(part_filter_kernel<kModeIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(part_filter_kernel<kModeLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(part_forward_kernel<kModeHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(part_forward_kernel<kModeIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(part_forward_kernel<kModeLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(part_conv_kernel<kModeHP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(part_conv_kernel<kModeIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  MFPA_CUDA(cudaFuncSetAttribute(part_conv_kernel<kModeLP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kConvSmem));
  return MFPA_OK;
}

// julius: half = int(zeros / cutoff / 2) with zeros = 8, cutoff = float32(fc) / float32(sr) as a Python float
static int fir_half(float fc_hz, int sample_rate, const char* name, int qi, double* cutoff_out) {
  const double c = (double)(fc_hz / (float)sample_rate);
  if (!(c > 0.0) || c > 0.5) {
    set_error("augment: query %d: %s cut-off %g Hz is outside (0, sr/2] - the reference raises ValueError "
              "(pass_filters.py:103-110)", qi, name, (double)fc_hz);
    return -1;
  }
  *cutoff_out = c;
  const double h = 8.0 / c / 2.0;
  if (h > 1e9) return 1 << 30;
  return (int)h;
}

int launch_augment(mfpa_ctx* ctx, const float* x, int B, int T, int64_t x_stride, int sample_rate,
                   const mfpa_aug_params* pp, const float* ir, int ir_stride, const float* noise, float* out,
                   bool final_norm, cudaStream_t st, const int64_t* ir_offsets, int64_t ir_bank_len, const LpfShape* lpf) {
  if (int e = aug_init_tables(ctx)) return e;
  // ---- derive per-query parameters on the host
  const size_t need = (sizeof(AugQ) + 5 * sizeof(int)) * (size_t)B;
  if (ctx->aug_copy_done) MFPA_CUDA(cudaEventSynchronize((cudaEvent_t)ctx->aug_copy_done));   // previous call's H2D of this buffer
  else { cudaEvent_t ev; MFPA_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)); ctx->aug_copy_done = ev; }
  if (ctx->aug_pinned_bytes < need) {
    if (ctx->aug_pinned) cudaFreeHost(ctx->aug_pinned);
    ctx->aug_pinned = nullptr;
    ctx->aug_pinned_bytes = need + need / 4;
    MFPA_CUDA(cudaMallocHost(&ctx->aug_pinned, ctx->aug_pinned_bytes));
  }
  AugQ* hq = (AugQ*)ctx->aug_pinned;
  int* hlist[4];   // queries whose stage-1 / IR / stage-5 / stage-6 filter needs the partitioned path
  int nlong[4] = {0, 0, 0, 0}, max_k[4] = {0, 0, 0, 0};
  for (int l = 0; l < 4; ++l) hlist[l] = reinterpret_cast<int*>(hq + B) + (size_t)l * B;
  int* hclip = reinterpret_cast<int*>(hq + B) + (size_t)4 * B;   // rows with Clipping applied (MFPA_OPT_CLIP_POOLED)
  int n_clip = 0;
  constexpr int kFastTaps = FN / 2 + 1;   // one overlap-save block keeps >= FN/2 valid outputs
  // grid width of each convolution stage: the most blocks any query needs (pass-through queries: T / FN)
  const int nb_pass = (T + FN - 1) / FN;
  int nb1 = nb_pass, nb2 = nb_pass, nb3 = nb_pass, nbir = nb_pass;
  auto blocks_of = [&](bool causal, int half, int ir_len) {
    const ConvGeom g = conv_geom(causal, half, ir_len, T);
    return (g.n_total + g.V - 1) / g.V;
  };
  auto up = [](int& a, int b) { if (b > a) a = b; };
  bool any_long_lp = false;
  auto note_long = [&](int l, uint32_t bit, int K, AugQ& q, int i) {
    q.long_mask |= bit; hlist[l][nlong[l]++] = i; max_k[l] = K > max_k[l] ? K : max_k[l];
  };
  for (int i = 0; i < B; ++i) {
    const mfpa_aug_params& p = pp[i];
    AugQ q{};
    q.apply = p.apply;
    double c;
    if (p.apply & MFPA_AUG_HPF1) {
      q.half1 = fir_half(p.fc1_hz, sample_rate, "loudspeaker high-pass", i, &c);
      if (q.half1 < 0) return MFPA_EINVAL;
      MFPA_REQUIRE(2 * (int64_t)q.half1 + 1 <= MFPA_AUG_MAX_TAPS, "augment: query %d: high-pass cut-off %g Hz needs %lld taps "
                   "(limit %d)", i, (double)p.fc1_hz, 2ll * q.half1 + 1, MFPA_AUG_MAX_TAPS);
      q.c1x2 = (float)(2.0 * c); q.arg1 = (float)(2.0 * c * 3.14159265358979323846);
      if (2 * q.half1 + 1 > kFastTaps) note_long(0, kLongHP1, 2 * q.half1 + 1, q, i);
      else up(nb1, blocks_of(false, q.half1, 0));
    }
    if (p.apply & MFPA_AUG_LPF) {
      if (lpf) {   // cut-off and window width given separately, as fractions of the sample rate
        c = lpf->cutoff[i];
        const double w = lpf->width[i];
        MFPA_REQUIRE(c > 0.0 && c <= 0.5 && w > 0.0 && w <= 0.5, "lowpass_filters: row %d: cut-off %g / window cut-off %g outside "
                     "(0, 0.5] of the sample rate (julius raises ValueError)", i, c, w);
        const double h = 8.0 / w / 2.0;
        q.half2 = h > 1e9 ? 1 << 30 : (int)h;
      } else {
        q.half2 = fir_half(p.fc2_hz, sample_rate, "low-pass", i, &c);
      }
      if (q.half2 < 0) return MFPA_EINVAL;
      MFPA_REQUIRE(2 * (int64_t)q.half2 + 1 <= MFPA_AUG_MAX_TAPS, "augment: query %d: low-pass cut-off %g Hz needs %lld taps "
                   "(limit %d)", i, (double)p.fc2_hz, 2ll * q.half2 + 1, MFPA_AUG_MAX_TAPS);
      q.c2x2 = (float)(2.0 * c); q.arg2 = (float)(2.0 * c * 3.14159265358979323846);
      if (q.half2 > kLpMaxHalf) any_long_lp = true;
      if (2 * q.half2 + 1 > kFastTaps) note_long(2, kLongLP, 2 * q.half2 + 1, q, i);
      else up(nb2, blocks_of(false, q.half2, 0));
    }
    if (p.apply & MFPA_AUG_HPF3) {
      q.half3 = fir_half(p.fc3_hz, sample_rate, "microphone high-pass", i, &c);
      if (q.half3 < 0) return MFPA_EINVAL;
      MFPA_REQUIRE(2 * (int64_t)q.half3 + 1 <= MFPA_AUG_MAX_TAPS, "augment: query %d: high-pass cut-off %g Hz needs %lld taps "
                   "(limit %d)", i, (double)p.fc3_hz, 2ll * q.half3 + 1, MFPA_AUG_MAX_TAPS);
      q.c3x2 = (float)(2.0 * c); q.arg3 = (float)(2.0 * c * 3.14159265358979323846);
      if (2 * q.half3 + 1 > kFastTaps) note_long(3, kLongHP3, 2 * q.half3 + 1, q, i);
      else up(nb3, blocks_of(false, q.half3, 0));
    }
    if (p.apply & MFPA_AUG_IR) {
      MFPA_REQUIRE(ir != nullptr, "augment: query %d applies an impulse response but ir_dev is NULL", i);
      if (ir_offsets) {   // responses addressed inside a device-resident bank
        MFPA_REQUIRE(p.ir_len >= 1 && p.ir_len <= MFPA_AUG_MAX_IR && ir_offsets[i] >= 0 && ir_offsets[i] + p.ir_len <= ir_bank_len,
                     "augment: query %d: impulse response [%lld, +%d) is outside the bank of %lld samples (or longer than %d)", i,
                     (long long)ir_offsets[i], p.ir_len, (long long)ir_bank_len, MFPA_AUG_MAX_IR);
        q.ir_off = ir_offsets[i];
      } else {
        MFPA_REQUIRE(p.ir_len >= 1 && p.ir_len <= ir_stride && p.ir_len <= MFPA_AUG_MAX_IR,
                     "augment: query %d: ir_len %d not in [1, min(ir_stride %d, %d)]", i, p.ir_len, ir_stride, MFPA_AUG_MAX_IR);
        q.ir_off = (int64_t)i * ir_stride;
      }
      q.ir_len = p.ir_len;
      if (p.ir_len > FN / 2) note_long(1, kLongIR, p.ir_len, q, i);
      else up(nbir, blocks_of(true, 0, p.ir_len));
    }
    if (p.apply & MFPA_AUG_NOISE) MFPA_REQUIRE(noise != nullptr, "augment: query %d mixes noise but noise_dev is NULL", i);
    if (p.apply & MFPA_AUG_CLIP) hclip[n_clip++] = i;
    q.snr_div = powf(10.0f, p.snr_db / 20.0f);
    q.gain = p.gain_factor;
    q.q_lo = p.clip_p / 2.0f;
    hq[i] = q;
  }
  MFPA_REQUIRE(noise == nullptr || ((uintptr_t)noise & 15) == 0, "augment: noise_dev must be 16-byte aligned");
  const size_t rows = sizeof(float) * (size_t)B * T;
  if (ctx->aug_a.reserve(rows) || ctx->aug_b.reserve(rows)) return MFPA_ENOMEM;
  if (ctx->aug_small.reserve((sizeof(AugQ) + sizeof(AugS)) * (size_t)B)) return MFPA_ENOMEM;
  if (ctx->aug_d.reserve(sizeof(float2) * (size_t)B * FM)) return MFPA_ENOMEM;
  float4* hspec = (float4*)ctx->aug_d.ptr;
  float* bufA = (float*)ctx->aug_a.ptr;
  float* bufB = (float*)ctx->aug_b.ptr;
  AugQ* dq = (AugQ*)ctx->aug_small.ptr;
  AugS* ds = (AugS*)(dq + B);
  if (int e = launch_pull(ctx, dq, hq, sizeof(AugQ) * (size_t)B, st)) return e;   // hq is pinned (ctx->aug_pinned)
  MFPA_CUDA(cudaMemsetAsync(ds, 0, sizeof(AugS) * (size_t)B, st));
  int* dlist = nullptr;
  if (nlong[0] + nlong[1] + nlong[2] + nlong[3]) {
    if (ctx->aug_long.reserve(sizeof(int) * 4 * (size_t)B)) return MFPA_ENOMEM;
    dlist = (int*)ctx->aug_long.ptr;
    if (int e = launch_pull(ctx, dlist, hlist[0], sizeof(int) * 4 * (size_t)B, st)) return e;
  }
  MFPA_CUDA(cudaEventRecord((cudaEvent_t)ctx->aug_copy_done, st));
  // partitioned overlap-save for the queries of list l (filters longer than one block takes)
  auto run_long = [&](int l, int mode, ConvArgs c) -> int {
    if (!nlong[l]) return MFPA_OK;
    const int P = (max_k[l] + kPartS - 1) / kPartS;
    const int n_total = mode == kModeIR ? T + max_k[l] - 1 : T;
    const int nb_out = (n_total + kPartS - 1) / kPartS, nb_in = nb_out + P - 1;
    // spectra scratch is (P + nb_in) x 64 KB per query: run the list in groups that keep it under the budget
    // (MFPA_OPT_PART_BUDGET_MB, 2 GiB by default)
    const size_t per_query = sizeof(float2) * FM * (size_t)(P + nb_in);
    int group = (int)(((size_t)ctx->opt_part_budget_mb << 20) / per_query);
    group = group < 1 ? 1 : (group > nlong[l] ? nlong[l] : group);
    if (ctx->aug_part.reserve(per_query * (size_t)group)) return MFPA_ENOMEM;
    for (int g0 = 0; g0 < nlong[l]; g0 += group) {
    const int ng = nlong[l] - g0 < group ? nlong[l] - g0 : group;
    PartArgs a{c, dlist + (size_t)l * B + g0, (float4*)ctx->aug_part.ptr,
               (float*)((float2*)ctx->aug_part.ptr + (size_t)ng * P * FM), P, nb_in};
    const dim3 gf(P, ng), gi(nb_in, ng), go(nb_out, ng);
    if (mode == kModeHP) {
      part_filter_kernel<kModeHP><<<gf, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
      part_forward_kernel<kModeHP><<<gi, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
      part_conv_kernel<kModeHP><<<go, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
    } else if (mode == kModeIR) {
      part_filter_kernel<kModeIR><<<gf, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
      part_forward_kernel<kModeIR><<<gi, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
      part_conv_kernel<kModeIR><<<go, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
    } else {
      part_filter_kernel<kModeLP><<<gf, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
      part_forward_kernel<kModeLP><<<gi, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
      part_conv_kernel<kModeLP><<<go, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
    }
    MFPA_CUDA(cudaGetLastError());
    }
    return MFPA_OK;
  };

  // one overlap-save block per grid cell; a centred FIR's kernel is the same for low- and high-pass (the
  // filter carries the difference), blocks per SM by MFPA_OPT_CONV_OCC
  const int occ = ctx->opt_conv_occ;
  auto launch_conv = [&](bool causal, dim3 grid, const ConvArgs& a) {
    const float* tw = (const float*)ctx->aug_tw_dev;
    if (causal) {
      if (occ == 2) fftconv_kernel<kModeIR, 2><<<grid, FT, kConvSmem, st>>>(a, tw);
      else fftconv_kernel<kModeIR, 3><<<grid, FT, kConvSmem, st>>>(a, tw);
    } else {
      if (occ == 2) fftconv_kernel<kModeHP, 2><<<grid, FT, kConvSmem, st>>>(a, tw);
      else fftconv_kernel<kModeHP, 3><<<grid, FT, kConvSmem, st>>>(a, tw);
    }
  };
  // stage 1: x -> A
  {
    ConvArgs a{x, x_stride, bufA, nullptr, 0, dq, ds, T, MFPA_AUG_HPF1, 1, hspec, 1};
    stage_mark(ctx, MFPA_STAGE_HPF1_FILTER, st);
    filter_spectrum_kernel<kModeHP><<<B, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
    stage_mark(ctx, MFPA_STAGE_HPF1_CONV, st);
    launch_conv(false, dim3((unsigned)nb1, B), a);
    MFPA_CUDA(cudaGetLastError());
    if (int e = run_long(0, kModeHP, a)) return e;
  }
  // stage 2: A -> B
  {
    ConvArgs a{bufA, T, bufB, ir, ir_stride, dq, ds, T, MFPA_AUG_IR, 0, hspec, 0};
    stage_mark(ctx, MFPA_STAGE_IR_FILTER, st);
    filter_spectrum_kernel<kModeIR><<<B, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
    stage_mark(ctx, MFPA_STAGE_IR_CONV, st);
    launch_conv(true, dim3((unsigned)nbir, B), a);
    MFPA_CUDA(cudaGetLastError());
    if (int e = run_long(1, kModeIR, a)) return e;
  }
  // stage 3: B -> A (z)
  {
    const unsigned gx = (unsigned)((T + 8191) / 8192);
    if (ctx->aug_lists.reserve(sizeof(unsigned) * 2 * kSelCap * (size_t)B)) return MFPA_ENOMEM;
    unsigned* lists = (unsigned*)ctx->aug_lists.ptr;
    stage_mark(ctx, MFPA_STAGE_MIX, st);
    clip_sample_kernel<<<B, kSelThreads, 0, st>>>(bufB, noise, dq, ds, T);
    mix_kernel<<<dim3(gx, B), 256, 0, st>>>(bufB, noise, bufA, dq, ds, T, lists);
    // stage 4: clip thresholds from the listed tails of z
    clip_finish_kernel<<<B, kSelThreads, 0, st>>>(bufA, lists, dq, ds, T);
    MFPA_CUDA(cudaGetLastError());
    if (ctx->opt_clip_pooled && n_clip > 1) {
      // the reference's batch semantics: thresholds from the pooled samples of the selected rows
      const int64_t n = (int64_t)n_clip * T;
      MFPA_REQUIRE(n <= (int64_t)1 << 24, "augment: pooled Clipping over %lld samples - torch.quantile refuses inputs above 16 777 216 "
                   "elements (clipping.py:76-90)", (long long)n);
      int64_t n_pad = 1;
      while (n_pad < n) n_pad <<= 1;
      if (ctx->aug_pool.reserve(sizeof(unsigned) * (size_t)n_pad + sizeof(int) * (size_t)B)) return MFPA_ENOMEM;
      unsigned* keys = (unsigned*)ctx->aug_pool.ptr;
      int* dsel = (int*)(keys + n_pad);
      if (int e = launch_pull(ctx, dsel, hclip, sizeof(int) * (size_t)n_clip, st)) return e;
      pool_fill_kernel<<<dim3((unsigned)((T + 4095) / 4096), (unsigned)n_clip), 256, 0, st>>>(bufA, dsel, T, dq, ds, keys);
      if (n_pad > n) pool_pad_kernel<<<1024, 256, 0, st>>>(keys, n, n_pad);
      const unsigned sort_blocks = (unsigned)((n_pad + 255) / 256 < 4736 ? (n_pad + 255) / 256 : 4736);
      for (int64_t k = 2; k <= n_pad; k <<= 1)
        for (int64_t j = k >> 1; j > 0; j >>= 1) bitonic_step_kernel<<<sort_blocks, 256, 0, st>>>(keys, n_pad, k, j);
      pool_thresholds_kernel<<<(n_clip + 127) / 128, 128, 0, st>>>(keys, n, dsel, n_clip, dq, ds);
      MFPA_CUDA(cudaGetLastError());
    }
  }
  // stage 5: A -> B (u)
  stage_mark(ctx, MFPA_STAGE_CLIP_LPF, st);
  {
    // every block first builds the query's taps (sinf / cosf per tap, a block reduction): give a block as many
    // 1024-sample tiles as the batch allows while ~4 blocks per SM-slot stay available (1.50 -> 1.38 ms per 10 k
    // queries; 16 blocks per query spent 8 % of the kernel on the taps)
    const unsigned tiles = (unsigned)((T + kLpTile - 1) / kLpTile);
    unsigned gx = (unsigned)((148 * 8 * 4 + B - 1) / B);
    gx = gx < 1u ? 1u : (gx > tiles ? tiles : gx);
    if (!any_long_lp) {
      clip_lpf_kernel<<<dim3(gx, B), 256, 0, st>>>(bufA, bufB, dq, ds, T, 0);
      MFPA_CUDA(cudaGetLastError());
    } else {
      // a low cut-off makes the "low-pass" FIR long: materialise the clipped signal, then FFT-convolve
      if (ctx->aug_c.reserve(rows)) return MFPA_ENOMEM;
      float* bufC = (float*)ctx->aug_c.ptr;
      clip_lpf_kernel<<<dim3(gx, B), 256, 0, st>>>(bufA, bufC, dq, ds, T, 1);
      MFPA_CUDA(cudaGetLastError());
      ConvArgs a{bufC, T, bufB, nullptr, 0, dq, ds, T, MFPA_AUG_LPF, 2, hspec, 0};
      filter_spectrum_kernel<kModeHP><<<B, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
      launch_conv(false, dim3((unsigned)nb2, B), a);
      MFPA_CUDA(cudaGetLastError());
      if (int e = run_long(2, kModeLP, a)) return e;
    }
  }
  // stage 6: B -> A (v) ; stage 7: A -> out
  {
    ConvArgs a{bufB, T, bufA, nullptr, 0, dq, ds, T, MFPA_AUG_HPF3, 3, hspec, 1};
    stage_mark(ctx, MFPA_STAGE_HPF3_FILTER, st);
    filter_spectrum_kernel<kModeHP><<<B, FT, kConvSmem, st>>>(a, (const float*)ctx->aug_tw_dev);
    stage_mark(ctx, MFPA_STAGE_HPF3_CONV, st);
    launch_conv(false, dim3((unsigned)nb3, B), a);
    MFPA_CUDA(cudaGetLastError());
    if (int e = run_long(3, kModeHP, a)) return e;
    const unsigned gx = (unsigned)((T + 4095) / 4096);
    if (out) {   // out == nullptr: the caller consumes stage 6's output in place (ctx->aug_a)
      norm_kernel<<<dim3(gx, B), 256, 0, st>>>(bufA, out, dq, ds, T, final_norm ? 1 : 0);
      MFPA_CUDA(cudaGetLastError());
    }
  }
  return MFPA_OK;
}

}  // namespace mfpa
