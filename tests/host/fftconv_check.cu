// CPU check of csrc/fftconv_core.cuh (test infrastructure; compiled and run by tests/test_fftconv_host.py).
// The build container has no GPU, so the block-level algorithm of augment.cu's convolution kernels -
// packing, the three FFT passes, the split/multiply pass, the inverse, the block geometry - is executed
// here exactly as a 256-thread block would (one loop over the thread index per pass = one barrier
// interval), through the same __host__ __device__ functions the kernels call, and compared with a float64
// direct convolution with the reference's boundary rules (julius replicate padding for the centred FIRs,
// pass_filters.py:84-155; zero extension and full-length output for the impulse response,
// impulse_response.py:119-164).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../../musicfpaugment_b200/csrc/fftconv_core.cuh"

using namespace mfpa::fc;

struct Block {
  std::vector<float> mem;
  float *re, *im, *twa_re, *twa_im, *twb_re, *twb_im;
  Block() : mem(kPlaneFloats, 0.f) {
    re = mem.data(); im = re + FPADF; twa_re = im + FPADF; twa_im = twa_re + kTwA; twb_re = twa_im + kTwA; twb_im = twb_re + kTwB;
    const double pi = 3.14159265358979323846;
    for (int e = 0; e < kTwA; ++e) { twa_re[e] = (float)cos(2 * pi * e / FM); twa_im[e] = (float)-sin(2 * pi * e / FM); }
    for (int e = 0; e < kTwB; ++e) { twb_re[e] = (float)cos(2 * pi * e / 512); twb_im[e] = (float)-sin(2 * pi * e / 512); }
  }
  template <typename F> void fill(F sample) {   // sample(i) = real sample i of the block, i < FN
    for (int i4 = 0; i4 < FN / 4; ++i4) {
      const int o = padi(2 * i4);
      re[o] = sample(4 * i4); im[o] = sample(4 * i4 + 1); re[o + 1] = sample(4 * i4 + 2); im[o + 1] = sample(4 * i4 + 3);
    }
  }
  void forward() {
    for (int t = 0; t < FT; ++t) pass_a<false>(re, im, twa_re, twa_im, t);
    for (int t = 0; t < FT; ++t) pass_b<false>(re, im, twb_re, twb_im, t);
    for (int t = 0; t < FT; ++t) pass_last_fwd(re, im, t);
  }
  void inverse() {
    for (int t = 0; t < FT; ++t) pass_last_inv(re, im, t);
    for (int t = 0; t < FT; ++t) pass_b<true>(re, im, twb_re, twb_im, t);
    for (int t = 0; t < FT; ++t) pass_a<true>(re, im, twa_re, twa_im, t);
  }
  float out(int ci) const { const int o = padi(ci >> 1); return (ci & 1) ? im[o] : re[o]; }
  // the kernels' fused forms: pass A reads the block from "global memory", the inverse pass A writes the result there
  template <typename F> void forward_from(F sample) {
    for (int t = 0; t < FT; ++t)
      pass_a_fwd_quads(re, im, twa_re, twa_im, t, [&](int u) {
        float4 v; const int i = 4 * t + 1024 * u;
        v.x = sample(i); v.y = sample(i + 1); v.z = sample(i + 2); v.w = sample(i + 3);
        return v; });
    for (int t = 0; t < FT; ++t) pass_b<false>(re, im, twb_re, twb_im, t);
    for (int t = 0; t < FT; ++t) pass_last_fwd(re, im, t);
  }
  void inverse_to(float* c) {   // c[FN]
    for (int t = 0; t < FT; ++t) pass_last_inv(re, im, t);
    for (int t = 0; t < FT; ++t) pass_b<true>(re, im, twb_re, twb_im, t);
    for (int t = 0; t < FT; ++t)
      pass_a_inv_quads(re, im, twa_re, twa_im, t, [&](int k, float4 v) {
        const int i = 4 * t + 1024 * k;
        c[i] = v.x; c[i + 1] = v.y; c[i + 2] = v.z; c[i + 3] = v.w; });
  }
};

static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

// 1. the transform itself: forward output at rev_pos(k) equals the DFT, inverse(forward(x)) = FM x
static int check_fft() {
  Block b;
  std::vector<double> xr(FM), xi(FM);
  for (int i = 0; i < FM; ++i) { xr[i] = frand(); xi[i] = frand(); }
  for (int i = 0; i < FM; ++i) { b.re[padi(i)] = (float)xr[i]; b.im[padi(i)] = (float)xi[i]; }
  b.forward();
  double worst = 0;
  const double pi = 3.14159265358979323846;
  for (int kk = 0; kk < 64; ++kk) {
    const int k = (kk * 997 + 13) % FM;
    double sr = 0, si = 0;
    for (int n = 0; n < FM; ++n) {
      const double a = -2 * pi * (double)((long long)n * k % FM) / FM;
      sr += xr[n] * cos(a) - xi[n] * sin(a);
      si += xr[n] * sin(a) + xi[n] * cos(a);
    }
    const int r = rev_pos(k);
    if (pos_freq(r) != k) { printf("pos_freq(rev_pos(%d)) = %d\n", k, pos_freq(r)); return 1; }
    const double er = fabs(b.re[padi(r)] - sr) + fabs(b.im[padi(r)] - si);
    if (er > worst) worst = er;
  }
  b.inverse();
  double worst_inv = 0;
  for (int i = 0; i < FM; ++i) {
    const double e = fabs(b.re[padi(i)] / FM - xr[i]) + fabs(b.im[padi(i)] / FM - xi[i]);
    if (e > worst_inv) worst_inv = e;
  }
  printf("fft: max |X - DFT| = %.3g (values ~ %.0f), max |ifft(fft(x)) - x| = %.3g\n", worst, sqrt((double)FM), worst_inv);
  return (worst < 2e-3 && worst_inv < 1e-5) ? 0 : 1;
}

// 2. block convolution against a float64 direct convolution
static int check_conv(bool causal, int half, int ir_len, int T) {
  const ConvGeom g = conv_geom(causal, half, ir_len, T);
  const int Kh = causal ? ir_len : 2 * half + 1;
  std::vector<double> h(Kh), x(T);
  for (int i = 0; i < T; ++i) x[i] = frand();
  double hsum = 0;
  for (int i = 0; i < Kh; ++i) {
    if (causal) h[i] = frand() * exp(-i / (0.15 * 8000.0));
    else {
      const double t = i - half, c = 4.0 / (half > 0 ? half : 1);   // any smooth low-pass shape will do
      h[i] = (0.5 - 0.5 * cos(2 * 3.14159265358979323846 * i / (2.0 * half + (half == 0)))) * (t == 0 ? 1.0 : sin(c * t) / (c * t));
      if (half == 0) h[i] = 1.0;
    }
    hsum += h[i];
  }
  // filter side
  Block fb;
  fb.fill([&](int i) { return (i >= g.zeros && i - g.zeros < Kh) ? (float)h[i - g.zeros] : 0.f; });
  fb.forward();
  const float scale = (float)(1.0 / FM / (causal ? 1.0 : hsum));
  std::vector<float4> hs(FM / 4), hr(FM / 4);
  for (int t = 0; t < FT; ++t) filter_pairs_store(fb.re, fb.im, scale, hs.data(), hr.data(), t);
  // signal side, block by block
  const int n_blocks = (g.n_total + g.V - 1) / g.V;
  std::vector<float> y(g.n_total, 0.f);
  Block sb;
  for (int blk = 0; blk < n_blocks; ++blk) {
    const int n0 = blk * g.V, s0 = n0 - g.lead;
    auto sample = [&](int i) {
      const int n = s0 + i;
      if (causal) return (n >= 0 && n < T) ? (float)x[n] : 0.f;
      return (float)x[n < 0 ? 0 : (n > T - 1 ? T - 1 : n)];
    };
    if (blk & 1) {   // alternate between the staged form and the form fused with the global-memory traffic
      sb.fill(sample);
      sb.forward();
      for (int t = 0; t < FT; ++t) filter_pairs_apply(sb.re, sb.im, hs.data(), hr.data(), t);
      sb.inverse();
      for (int n = n0; n < g.n_total && n < n0 + g.V; ++n) y[n] = sb.out(n - n0 + g.off);
    } else {
      std::vector<float> c(FN);
      sb.forward_from(sample);
      for (int t = 0; t < FT; ++t) filter_pairs_apply(sb.re, sb.im, hs.data(), hr.data(), t);
      sb.inverse_to(c.data());
      for (int n = n0; n < g.n_total && n < n0 + g.V; ++n) y[n] = c[n - n0 + g.off];
    }
  }
  // reference
  double worst = 0, peak = 0;
  const int step = g.n_total > 4000 ? 7 : 1;
  for (int n = 0; n < g.n_total; n += step) {
    double acc = 0;
    for (int k = 0; k < Kh; ++k) {
      const int m = causal ? n - k : n + half - k;
      double xv;
      if (causal) xv = (m >= 0 && m < T) ? x[m] : 0.0;
      else xv = x[m < 0 ? 0 : (m > T - 1 ? T - 1 : m)];
      acc += h[k] * (double)(float)xv;
    }
    if (!causal) acc /= hsum;
    if (fabs(acc) > peak) peak = fabs(acc);
    const double e = fabs(acc - y[n]);
    if (e > worst) worst = e;
  }
  printf("conv %s half %d ir_len %d T %d: %d blocks, V %d, max err %.3g (peak %.3g)\n", causal ? "causal" : "centred", half, ir_len, T,
         n_blocks, g.V, worst, peak);
  return worst <= 2e-5 * (peak > 1 ? peak : 1) ? 0 : 1;
}

int main() {
  srand(12345);
  int bad = check_fft();
  bad += check_conv(false, 0, 0, 5000);
  bad += check_conv(false, 1, 0, 5000);
  bad += check_conv(false, 8, 0, 20000);        // a 17-tap low-pass
  bad += check_conv(false, 213, 0, 64000);      // 150 Hz high-pass: 427 taps
  bad += check_conv(false, 1066, 0, 40000);     // 30 Hz: 2133 taps
  bad += check_conv(false, 1599, 0, 30001);     // half = 3 mod 4, odd T
  bad += check_conv(false, 4095, 0, 20000);     // the longest one-block filter but one
  bad += check_conv(false, 4096, 0, 20000);     // the longest one-block filter
  bad += check_conv(true, 0, 1, 3000);
  bad += check_conv(true, 0, 8000, 64000);      // 1 s response on an 8 s query (BASELINE configs[2])
  bad += check_conv(true, 0, 8192, 17000);
  bad += check_conv(true, 0, 5, 70000);
  printf(bad ? "FAILED (%d)\n" : "all ok\n", bad);
  return bad ? 1 : 0;
}
