// 8192-point complex FFT + real-signal split/multiply pass for the overlap-save convolutions of
// augment.cu (AugmentFP's FIR filters and impulse response, augmentation/transformations/
// pass_filters.py:84-155, impulse_response.py:119-164).
//
// Layout and arithmetic are chosen for sm_100's packed FP32 pipe (FADD2 / FMUL2 / FFMA2, two
// float32 operations per issue slot):
//   * the 8192 complex points live in shared memory as two PLANES (re[], im[]), padded by two
//     floats per 32 so that every pass below is bank-conflict free;
//   * every thread works on TWO butterflies at once, the pair held as float2 = (butterfly 0,
//     butterfly 1) per plane.  A complex multiply of the pair is then exactly four packed
//     instructions and every butterfly add one, with no swaps or half negations;
//   * 8192 = 16 * 16 * 16 * 2: two radix-16 passes through shared memory (twiddle powers from a
//     depth-4 product tree of one table entry), and a last pass that does the third radix-16 AND
//     the radix-2 on the 32 contiguous points a thread owns (constant twiddles, no table).
// Forward is decimation in frequency (natural order in, digit-reversed out), the inverse undoes it
// pass by pass (digit-reversed in, natural out, unscaled), so no reordering pass exists.
//
// Everything here is __host__ __device__: tests/host_fftconv_check.cu runs the very same passes
// on the CPU (thread loop per pass) against a float64 direct convolution, because the build box
// has no GPU.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace mfpa {
namespace fc {

constexpr int FN = 16384;          // real samples per overlap-save block
constexpr int LOGM = 13;
constexpr int FM = 1 << LOGM;      // complex FFT length (FN / 2)
constexpr int FT = 256;            // threads per block: one double butterfly per thread and pass
constexpr int FPADF = FM + 2 * (FM / 32);   // floats per padded plane
constexpr int kTwA = 512;          // pass-A table: exp(-2 pi i e / FM), e < 512
constexpr int kTwB = 32;           // pass-B table: exp(-2 pi i e / 512), e < 32
constexpr int kPlaneFloats = 2 * FPADF + 2 * kTwA + 2 * kTwB;
constexpr int kNumDp = FM / 4;     // double pairs handled by the split pass

#if defined(__CUDACC__)
#define FC_HD __host__ __device__ __forceinline__
#else
#define FC_HD inline
#endif

FC_HD int padi(int i) { return i + ((i >> 5) << 1); }

// ---- packed float2 arithmetic (device: one instruction each; host: the same roundings) ----
FC_HD float2 f2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
FC_HD float2 add2(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  return __fadd2_rn(a, b);
#else
  return f2(a.x + b.x, a.y + b.y);
#endif
}
FC_HD float2 sub2(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  return __fadd2_rn(a, f2(-b.x, -b.y));
#else
  return f2(a.x - b.x, a.y - b.y);
#endif
}
FC_HD float2 mul2(float2 a, float2 b) {
#ifdef __CUDA_ARCH__
  return __fmul2_rn(a, b);
#else
  return f2(a.x * b.x, a.y * b.y);
#endif
}
FC_HD float2 fma2(float2 a, float2 b, float2 c) {   // a * b + c
#ifdef __CUDA_ARCH__
  return __ffma2_rn(a, b, c);
#else
  return f2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}
FC_HD float2 fnma2(float2 a, float2 b, float2 c) {  // c - a * b
#ifdef __CUDA_ARCH__
  return __ffma2_rn(f2(-a.x, -a.y), b, c);
#else
  return f2(fmaf(-a.x, b.x, c.x), fmaf(-a.y, b.y, c.y));
#endif
}
FC_HD float2 bc(float a) { return f2(a, a); }
FC_HD float2 swp(float2 a) { return f2(a.y, a.x); }
// (x + y, x - y) of one register pair.  Two scalar adds: ptxas follows the one-instruction packed form
// (FADD2 R, R.HI_LO.NP, R.LO_HI) with two register copies whenever its result feeds a vector store.
FC_HD float2 sumdiff(float2 a) { return f2(a.x + a.y, a.x - a.y); }

// two complex numbers, planar: re = (re of #0, re of #1), im likewise
struct c2 { float2 re, im; };
FC_HD c2 mk(float2 re, float2 im) { c2 r; r.re = re; r.im = im; return r; }
FC_HD c2 cadd(c2 a, c2 b) { return mk(add2(a.re, b.re), add2(a.im, b.im)); }
FC_HD c2 csub(c2 a, c2 b) { return mk(sub2(a.re, b.re), sub2(a.im, b.im)); }
FC_HD c2 cmul(c2 a, c2 b) {
  return mk(fnma2(a.im, b.im, mul2(a.re, b.re)), fma2(a.im, b.re, mul2(a.re, b.im)));
}
FC_HD c2 cmulc(c2 a, c2 b) {   // a * conj(b)
  return mk(fma2(a.im, b.im, mul2(a.re, b.re)), fnma2(a.re, b.im, mul2(a.im, b.re)));
}
FC_HD c2 cmac(c2 a, c2 b, c2 acc) {   // acc + a * b
  return mk(fnma2(a.im, b.im, fma2(a.re, b.re, acc.re)), fma2(a.im, b.re, fma2(a.re, b.im, acc.im)));
}
FC_HD c2 cmulk(c2 a, float cr, float ci) {   // times the constant cr + i ci (immediate operands)
  return mk(fnma2(a.im, bc(ci), mul2(a.re, bc(cr))), fma2(a.re, bc(ci), mul2(a.im, bc(cr))));
}
FC_HD c2 cld(const float* re, const float* im, int off) {
  return mk(*reinterpret_cast<const float2*>(re + off), *reinterpret_cast<const float2*>(im + off));
}
FC_HD void cst(float* re, float* im, int off, c2 v) {
  *reinterpret_cast<float2*>(re + off) = v.re;
  *reinterpret_cast<float2*>(im + off) = v.im;
}

// 4-point DFT with kernel e^{S i 2 pi nk/4}; S = -1 forward, +1 inverse (unscaled)
template <int S> FC_HD void dft4(c2& a, c2& b, c2& c, c2& d) {
  const c2 s0 = cadd(a, c), s1 = csub(a, c), s2 = cadd(b, d), s3 = csub(b, d);
  a = cadd(s0, s2);
  c = csub(s0, s2);
  if (S < 0) {   // b = s1 - i s3, d = s1 + i s3
    b = mk(add2(s1.re, s3.im), sub2(s1.im, s3.re));
    d = mk(sub2(s1.re, s3.im), add2(s1.im, s3.re));
  } else {
    b = mk(sub2(s1.re, s3.im), add2(s1.im, s3.re));
    d = mk(add2(s1.re, s3.im), sub2(s1.im, s3.re));
  }
}

// 16-point DFT, natural order in; output X[k] lands in v[4 * (k & 3) + (k >> 2)]
#define FC_OUT16(k) (4 * ((k) & 3) + ((k) >> 2))
template <int S> FC_HD void dft16(c2 (&v)[16]) {
  constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f, h = 0.70710678118654752440f;
  constexpr float sg = S < 0 ? -1.f : 1.f;
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4<S>(v[n2], v[4 + n2], v[8 + n2], v[12 + n2]);
  // v[4 k1 + n2] *= W16^(n2 k1), W16 = exp(S 2 pi i / 16)
  v[5] = cmulk(v[5], c1, sg * s1);
  v[6] = cmulk(v[6], h, sg * h);
  v[7] = cmulk(v[7], s1, sg * c1);
  v[9] = cmulk(v[9], h, sg * h);
  v[10] = S < 0 ? mk(v[10].im, f2(-v[10].re.x, -v[10].re.y)) : mk(f2(-v[10].im.x, -v[10].im.y), v[10].re);
  v[11] = cmulk(v[11], -h, sg * h);
  v[13] = cmulk(v[13], s1, sg * c1);
  v[14] = cmulk(v[14], -h, sg * h);
  v[15] = cmulk(v[15], -c1, sg * -s1);
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) dft4<S>(v[4 * k1], v[4 * k1 + 1], v[4 * k1 + 2], v[4 * k1 + 3]);
}

// powers w^1 .. w^15 of a twiddle pair with multiplication depth <= 4
FC_HD void twiddle_powers(c2 w1, c2 (&w)[16]) {
  w[1] = w1; w[2] = cmul(w1, w1); w[3] = cmul(w[2], w1); w[4] = cmul(w[2], w[2]);
  w[5] = cmul(w[4], w1); w[6] = cmul(w[3], w[3]); w[7] = cmul(w[4], w[3]); w[8] = cmul(w[4], w[4]);
  w[9] = cmul(w[8], w1); w[10] = cmul(w[5], w[5]); w[11] = cmul(w[8], w[3]); w[12] = cmul(w[6], w[6]);
  w[13] = cmul(w[8], w[5]); w[14] = cmul(w[7], w[7]); w[15] = cmul(w[8], w[7]);
}

// One radix-16 pass on the double butterfly whose 16 points sit at off0 + t * stride (padded float
// offsets).  w1 = twiddle pair of the two butterflies; the inverse multiplies by the conjugate powers.
template <bool INV> FC_HD void pass16(float* re, float* im, int off0, int stride, c2 w1) {
  c2 v[16], w[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) v[t] = cld(re, im, off0 + t * stride);
  twiddle_powers(w1, w);
  if (INV) {
#pragma unroll
    for (int t = 1; t < 16; ++t) v[t] = cmulc(v[t], w[t]);
    dft16<1>(v);
#pragma unroll
    for (int k = 0; k < 16; ++k) cst(re, im, off0 + k * stride, v[FC_OUT16(k)]);
  } else {
    dft16<-1>(v);
#pragma unroll
    for (int k = 0; k < 16; ++k) cst(re, im, off0 + k * stride, k ? cmul(v[FC_OUT16(k)], w[k]) : v[FC_OUT16(k)]);
  }
}

// cos / sin of 2 pi k / 32, k < 16
#define FC_C32 {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f, 0.70710678118654752440f, \
                0.55557023301960222474f, 0.38268343236508977173f, 0.19509032201612826785f, 0.f, -0.19509032201612826785f,  \
                -0.38268343236508977173f, -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,   \
                -0.92387953251128675613f, -0.98078528040323044913f}
#define FC_S32 {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f, 0.70710678118654752440f, \
                0.83146961230254523708f, 0.92387953251128675613f, 0.98078528040323044913f, 1.f, 0.98078528040323044913f,  \
                0.92387953251128675613f, 0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,   \
                0.38268343236508977173f, 0.19509032201612826785f}

// Last forward pass on the 32 contiguous points of block m: radix-16 of the even-indexed and the
// odd-indexed points (= the two butterflies of a float2), twiddle exp(-2 pi i k / 32) on the odd
// one, radix-2 across the pair.
FC_HD void pass_last_fwd(float* re, float* im, int m) {
  constexpr float c32[16] = FC_C32, s32[16] = FC_S32;
  const int base = 34 * m;
  c2 v[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) v[t] = cld(re, im, base + 2 * t);
  dft16<-1>(v);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    c2 x = v[FC_OUT16(k)];
    if (k) {   // (x.#0, x.#1) *= (1, c - i s): constant pairs, no lane shuffling
      const float2 cc = f2(1.f, c32[k]), ss = f2(0.f, s32[k]);
      x = mk(fma2(x.im, ss, mul2(x.re, cc)), fnma2(x.re, ss, mul2(x.im, cc)));
    }
    cst(re, im, base + 2 * k, mk(sumdiff(x.re), sumdiff(x.im)));
  }
}
// ... and its inverse: radix-2, conjugate twiddle, inverse radix-16
FC_HD void pass_last_inv(float* re, float* im, int m) {
  constexpr float c32[16] = FC_C32, s32[16] = FC_S32;
  const int base = 34 * m;
  c2 v[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const c2 a = cld(re, im, base + 2 * k);
    c2 x = mk(sumdiff(a.re), sumdiff(a.im));
    if (k) {   // (x.#0, x.#1) *= (1, c + i s)
      const float2 cc = f2(1.f, c32[k]), ss = f2(0.f, s32[k]);
      x = mk(fnma2(x.im, ss, mul2(x.re, cc)), fma2(x.re, ss, mul2(x.im, cc)));
    }
    v[k] = x;
  }
  dft16<1>(v);
#pragma unroll
  for (int t = 0; t < 16; ++t) cst(re, im, base + 2 * t, v[FC_OUT16(t)]);
}

// The three passes, thread p of FT (a barrier separates consecutive passes).
// pass A: sub-transform length 8192, stride 512, butterflies j = 2p, 2p+1
template <bool INV> FC_HD void pass_a(float* re, float* im, const float* twa_re, const float* twa_im, int p) {
  const c2 w1 = cld(twa_re, twa_im, 2 * p);
  pass16<INV>(re, im, padi(2 * p), 544, w1);   // padi(j + 512 t) = padi(j) + 544 t
}
// pass B: sub-transform length 512, stride 32, group p >> 4, butterflies j = 2 (p & 15), +1
template <bool INV> FC_HD void pass_b(float* re, float* im, const float* twb_re, const float* twb_im, int p) {
  const int j = 2 * (p & 15);
  const c2 w1 = cld(twb_re, twb_im, j);
  pass16<INV>(re, im, padi(512 * (p >> 4) + j), 34, w1);   // padi(base + 32 t) = padi(base) + 34 t
}

// Pass A fused with the block's global-memory traffic.  The two butterflies of thread p (j = 2p, 2p + 1) take points
// j + 512 t, which in the packed real block (even sample -> re, odd -> im) are the FOUR consecutive samples
// 4p + 1024 t .. + 3: one 16-byte load per t, coalesced over p.  So the forward pass A reads its input straight from
// global memory (no staging pass through shared memory, one barrier less) and the inverse pass A - the last pass of
// the inverse transform - hands its output to the caller as float4 of four consecutive time samples.
//   quad(t)     -> float4 = samples 4p + 1024 t .. + 3 of the block (t < 16)
//   store(k, v) <-  v = samples 4p + 1024 k .. + 3 of the circular convolution (unscaled inverse, k < 16)
template <typename Q>
FC_HD void pass_a_fwd_quads(float* re, float* im, const float* twa_re, const float* twa_im, int p, Q quad) {
  c2 v[16], w[16];
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const float4 x = quad(t);
    v[t] = mk(f2(x.x, x.z), f2(x.y, x.w));
  }
  twiddle_powers(cld(twa_re, twa_im, 2 * p), w);
  dft16<-1>(v);
  const int off0 = padi(2 * p);
#pragma unroll
  for (int k = 0; k < 16; ++k) cst(re, im, off0 + k * 544, k ? cmul(v[FC_OUT16(k)], w[k]) : v[FC_OUT16(k)]);
}
template <typename S>
FC_HD void pass_a_inv_quads(const float* re, const float* im, const float* twa_re, const float* twa_im, int p, S store) {
  c2 v[16], w[16];
  const int off0 = padi(2 * p);
#pragma unroll
  for (int t = 0; t < 16; ++t) v[t] = cld(re, im, off0 + t * 544);
  twiddle_powers(cld(twa_re, twa_im, 2 * p), w);
#pragma unroll
  for (int t = 1; t < 16; ++t) v[t] = cmulc(v[t], w[t]);
  dft16<1>(v);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const c2 x = v[FC_OUT16(k)];
    float4 o;
    o.x = x.re.x; o.y = x.im.x; o.z = x.re.y; o.w = x.im.y;
    store(k, o);
  }
}

// position of frequency k after the forward transform, and the frequency held at position r
FC_HD int rev_pos(int k) { return ((k & 15) << 9) | (((k >> 4) & 15) << 5) | (((k >> 8) & 15) << 1) | (k >> 12); }
FC_HD int pos_freq(int r) { return (r >> 9) | (((r >> 5) & 15) << 4) | (((r >> 1) & 15) << 8) | ((r & 1) << 12); }

// ---- real-signal split / multiply / re-pack --------------------------------------------------------
// Z = FFT_FM(even samples + i odd samples).  Frequencies k and FM - k of the real signal's spectrum come
// from the pair Z[k], Z[FM - k].  With  E' = Z[k] + conj Z[FM-k],  D' = Z[k] - conj Z[FM-k],
// w' = -i exp(-i pi k / FM),  F = w' D':
//     X[k] = (E' + F) / 2,   conj X[FM-k] = (E' - F) / 2.
// The filter is kept as  S = s/4 (E'_h) , R = s/4 (F_h)  of ITS packed transform (s = output scale), so
//     Ey = S E' + R F,  Dy = R E' + S F,  V = conj(w') Dy,   Zy[k] = Ey + V,  Zy[FM-k] = conj(Ey - V)
// is the packed transform of the circular convolution.  Block m of the digit-reversed layout (positions
// 32 m .. 32 m + 31) holds frequencies c + 256 d2 + 4096 p (c = c(m), d2 < 16, p < 2) at 2 d2 + p; its
// partners are in block mbar at 2 (15 - d2) + (1 - p)  [block 0: slot 16 - d2].  One "double pair"
// (m, d2 < 8) = the float2 at slot d2 of block m against the swapped float2 at the partner slot: lane 0
// is the pair of k0 = c + 256 d2, lane 1 the pair of k0 + 4096 (twiddle times -i).
struct DpGeom { int off_a, off_b; float cw, sw; };   // cos / sin of pi c / FM (block twiddle)

FC_HD int block_c(int m) { return ((m >> 4) & 15) + 16 * (m & 15); }
FC_HD int partner_block(int m) {
  if (m == 0) return 0;
  const int cb = 256 - block_c(m);
  return ((cb & 15) << 4) | (cb >> 4);
}
// float offsets of the two slots of double pair (m, d2)
FC_HD void dp_offsets(int m, int mbar, int d2, int& off_a, int& off_b) {
  off_a = 34 * m + 2 * d2;
  off_b = 34 * mbar + 2 * ((m == 0 ? 16 : 15) - d2);
}
// twiddle pair w' of double pair (m, d2): lane 0 = -i W^{k0} = (-sin, -cos)(pi k0 / FM), lane 1 = -W^{k0},
// k0 = c(m) + 256 d2.  d2 -> d2 + 1 turns both lanes by exp(-i pi / 32): one constant complex multiply.
FC_HD c2 dp_twiddle0(float cw, float sw) { return mk(f2(-sw, -cw), f2(-cw, sw)); }
FC_HD c2 dp_twiddle_next(c2 wv) { return cmulk(wv, 0.99518472667219688624f, -0.09801714032956060199f); }
FC_HD void dp_split(c2 z1, c2 z2raw, c2 wv, c2& ep, c2& f) {
  const float2 z2re = swp(z2raw.re), z2im = swp(z2raw.im);
  ep = mk(add2(z1.re, z2re), sub2(z1.im, z2im));
  const c2 dp = mk(sub2(z1.re, z2re), add2(z1.im, z2im));
  f = cmul(wv, dp);
}
// (Ey, Dy) -> the two float2 to store at slot a and slot b
FC_HD void dp_repack(c2 ey, c2 dy, c2 wv, c2& za, c2& zb) {
  const c2 v = cmulc(dy, wv);
  za = cadd(ey, v);
  // conj(Ey - V), lanes swapped back into the partner slot's order
  zb = mk(sub2(swp(ey.re), swp(v.re)), sub2(swp(v.im), swp(ey.im)));
}

// filter-side float4 pair of a double pair (two planes of FM/4 float4 per query):
//   hs[dpi] = (S.re.#0, S.re.#1, S.im.#0, S.im.#1), hr[dpi] likewise for R;  dpi = d2 * 256 + m
FC_HD int dp_index(int m, int d2) { return d2 * 256 + m; }

// Scalar version of one pair for the three self-paired specials of block 0 (thread 0 only):
//   position 0: k = 0 and k = FM (both real)    position 1: k = FM/2 (pairs with itself)
//   positions 16 / 17: k = 2048 and k = 6144 (pair with each other)
struct Specials { float a0, a1, bre, bim; float sre, sim, rre, rim; };   // = hs[0], hr[0]
constexpr float kCosQ = 0.70710678118654752440f;   // cos(pi 2048 / 8192)

// filter side: Zh = packed transform of the taps (in place in re/im), s = scale
FC_HD Specials specials_from_filter(const float* re, const float* im, float s) {
  Specials sp;
  const float zx = re[0], zy = im[0];
  sp.a0 = 0.5f * s * (zx + zy);
  sp.a1 = 0.5f * s * (zx - zy);
  sp.bre = s * re[1];
  sp.bim = s * im[1];
  // pair (2048, 6144): w' = -i exp(-i pi / 4) = (-sin, -cos)(pi/4)
  const float z1r = re[16], z1i = im[16], z2r = re[17], z2i = im[17];
  const float er = z1r + z2r, ei = z1i - z2i, dr = z1r - z2r, di = z1i + z2i;
  const float wr = -kCosQ, wi = -kCosQ;
  const float fr = wr * dr - wi * di, fi = wr * di + wi * dr;
  sp.sre = 0.25f * s * er; sp.sim = 0.25f * s * ei;
  sp.rre = 0.25f * s * fr; sp.rim = 0.25f * s * fi;
  return sp;
}
// signal side: accumulate one (block transform, filter) product of the specials, then write them back.
// z8 = {re[0], im[0], re[1], im[1], re[16], im[16], re[17], im[17]} of the block transform.
struct SpecAcc { float y0, ym, br, bi, eyr, eyi, dyr, dyi; };
FC_HD SpecAcc spec_acc_zero() { SpecAcc a; a.y0 = a.ym = a.br = a.bi = a.eyr = a.eyi = a.dyr = a.dyi = 0.f; return a; }
FC_HD void specials_accumulate(SpecAcc& a, const float (&z)[8], const Specials& sp) {
  a.y0 += (z[0] + z[1]) * sp.a0;
  a.ym += (z[0] - z[1]) * sp.a1;
  a.br += z[2] * sp.bre - z[3] * sp.bim;
  a.bi += z[2] * sp.bim + z[3] * sp.bre;
  const float er = z[4] + z[6], ei = z[5] - z[7], dr = z[4] - z[6], di = z[5] + z[7];
  const float wr = -kCosQ, wi = -kCosQ;
  const float fr = wr * dr - wi * di, fi = wr * di + wi * dr;
  a.eyr += sp.sre * er - sp.sim * ei + sp.rre * fr - sp.rim * fi;
  a.eyi += sp.sre * ei + sp.sim * er + sp.rre * fi + sp.rim * fr;
  a.dyr += sp.rre * er - sp.rim * ei + sp.sre * fr - sp.sim * fi;
  a.dyi += sp.rre * ei + sp.rim * er + sp.sre * fi + sp.sim * fr;
}
FC_HD void specials_finish(float* re, float* im, const SpecAcc& a) {
  re[0] = a.y0 + a.ym; im[0] = a.y0 - a.ym;
  re[1] = a.br; im[1] = a.bi;
  const float wr = -kCosQ, wi = -kCosQ;
  const float vr = a.dyr * wr + a.dyi * wi, vi = a.dyi * wr - a.dyr * wi;   // Dy conj(w')
  re[16] = a.eyr + vr; im[16] = a.eyi + vi;
  re[17] = a.eyr - vr; im[17] = vi - a.eyi;
}
FC_HD void specials_apply(float* re, float* im, const Specials& sp) {
  const float z[8] = {re[0], im[0], re[1], im[1], re[16], im[16], re[17], im[17]};
  SpecAcc a = spec_acc_zero();
  specials_accumulate(a, z, sp);
  specials_finish(re, im, a);
}

// ---- the split pass as thread m of FT runs it (shared by the kernels and the host check) -------------
FC_HD void block_twiddle(int m, float& cw, float& sw) {   // cos / sin of pi c(m) / FM
#ifdef __CUDA_ARCH__
  sincospif((float)block_c(m) * (1.0f / (float)FM), &sw, &cw);
#else
  const double a = 3.14159265358979323846 * (double)block_c(m) / (double)FM;
  cw = (float)cos(a); sw = (float)sin(a);
#endif
}
FC_HD float4 ldg4(const float4* p) {
#ifdef __CUDA_ARCH__
  return __ldg(p);
#else
  return *p;
#endif
}
FC_HD float4 f4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
FC_HD Specials specials_unpack(float4 vs, float4 vr) {
  Specials sp;
  sp.a0 = vs.x; sp.a1 = vs.y; sp.bre = vs.z; sp.bim = vs.w;
  sp.sre = vr.x; sp.sim = vr.y; sp.rre = vr.z; sp.rim = vr.w;
  return sp;
}
FC_HD c2 c2_unpack(float4 v) { return mk(f2(v.x, v.y), f2(v.z, v.w)); }

// Filter side: the packed transform of the taps (in the planes) -> S / R double pairs of thread m.
FC_HD void filter_pairs_store(const float* re, const float* im, float scale, float4* hs, float4* hr, int m) {
  const int mbar = partner_block(m);
  float sw, cw;
  block_twiddle(m, cw, sw);
  const float q = 0.25f * scale;
  c2 wv = dp_twiddle0(cw, sw);
#pragma unroll
  for (int d2 = 0; d2 < 8; ++d2) {
    if (d2) wv = dp_twiddle_next(wv);
    if (m == 0 && d2 == 0) {
      const Specials sp = specials_from_filter(re, im, scale);
      hs[0] = f4(sp.a0, sp.a1, sp.bre, sp.bim);
      hr[0] = f4(sp.sre, sp.sim, sp.rre, sp.rim);
      continue;
    }
    int oa, ob;
    dp_offsets(m, mbar, d2, oa, ob);
    c2 ep, f;
    dp_split(cld(re, im, oa), cld(re, im, ob), wv, ep, f);
    const int i = dp_index(m, d2);
    hs[i] = f4(q * ep.re.x, q * ep.re.y, q * ep.im.x, q * ep.im.y);
    hr[i] = f4(q * f.re.x, q * f.re.y, q * f.im.x, q * f.im.y);
  }
}

// Signal side: packed transform of the block (in the planes) -> packed transform of block (*) filter.
FC_HD void filter_pairs_apply(float* re, float* im, const float4* hs, const float4* hr, int m) {
  const int mbar = partner_block(m);
  float sw, cw;
  block_twiddle(m, cw, sw);
  c2 wv = dp_twiddle0(cw, sw);
#pragma unroll
  for (int h4 = 0; h4 < 2; ++h4) {
    float4 vs[4], vr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {   // filter values of four double pairs, fetched ahead of the math
      const int i = dp_index(m, 4 * h4 + u);
      vs[u] = ldg4(hs + i);
      vr[u] = ldg4(hr + i);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int d2 = 4 * h4 + u;
      if (d2) wv = dp_twiddle_next(wv);
      if (m == 0 && d2 == 0) {
        specials_apply(re, im, specials_unpack(vs[u], vr[u]));
        continue;
      }
      int oa, ob;
      dp_offsets(m, mbar, d2, oa, ob);
      c2 ep, f, za, zb;
      dp_split(cld(re, im, oa), cld(re, im, ob), wv, ep, f);
      const c2 S = c2_unpack(vs[u]), R = c2_unpack(vr[u]);
      const c2 ey = cmac(R, f, cmul(S, ep)), dy = cmac(S, f, cmul(R, ep));
      dp_repack(ey, dy, wv, za, zb);
      cst(re, im, oa, za);
      cst(re, im, ob, zb);
    }
  }
}

// ---- block geometry of the one-block-per-V-outputs path ---------------------------------------------
// Everything is kept a multiple of 4 samples so the signal is read and written with 16-byte accesses: a
// centred FIR (half-width h) gets d = (-h) mod 4 zero taps prepended (delay h + d = 0 mod 4); the causal
// response starts its block round_up4(K - 1) samples early.
//   c[i] = circular convolution of the block that starts at sample n0 - lead;  y[n] = c[n - n0 + off]
struct ConvGeom { int K, zeros, lead, off, V, n_total; };
FC_HD ConvGeom conv_geom(bool causal, int half, int ir_len, int T) {
  ConvGeom g;
  if (causal) {
    g.K = ir_len; g.zeros = 0;
    g.lead = (ir_len - 1 + 3) & ~3;
    g.off = g.lead;
    g.n_total = T + ir_len - 1;
  } else {
    g.zeros = (4 - (half & 3)) & 3;
    g.K = 2 * half + 1 + g.zeros;
    g.lead = half + g.zeros;
    g.off = 2 * g.lead;
    g.n_total = T;
  }
  g.V = FN - g.off;
  return g;
}

}  // namespace fc
}  // namespace mfpa
