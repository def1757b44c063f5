// Evaluation step after the path: peak-mask precision / recall and PSNR sums.
// Replaces the per-peak Python loops of testing/metrics.py:10-162 (Recall, Precision) and the sums behind
// torchmetrics' PeakSignalNoiseRatio as testing/metrics.py:7 configures it.  The 3x3 kernel of the reference is
// zero except for its centre, so a peak of one mask looks at ONE position of the other: its own (i, j) in the
// interior and at the high edges, (i + 1, j) / (i, j + 1) at i == 0 / j == 0, where the reference cuts the window
// on the low side but the kernel on the high side (metrics.py:44-47, :72-75) - reproduced as it is.
#include "common.cuh"

namespace mfpa {

namespace {

// masks [B][H][W]: out[0] = #{gt != 0}, out[1] = sum over gt != 0 of predicted at the looked-at position,
// out[2] = #{predicted != 0}, out[3] = sum over predicted != 0 of gt there   (float64 sums: exact for 0/1 masks)
__global__ void __launch_bounds__(256) mask_metrics_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                           int64_t n, int H, int W, double* __restrict__ out) {
  __shared__ double red[4][8];
  double a[4] = {0.0, 0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p = pred[i], g = gt[i];
    if (p == 0.f && g == 0.f) continue;
    const int col = (int)(i % W), row = (int)((i / W) % H);
    const int64_t look = i + ((row == 0 && H > 1) ? W : 0) + ((col == 0 && W > 1) ? 1 : 0);
    if (g != 0.f) { a[0] += 1.0; a[1] += (double)pred[look]; }
    if (p != 0.f) { a[2] += 1.0; a[3] += (double)gt[look]; }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
#pragma unroll
    for (int o = 16; o; o >>= 1) a[k] += __shfl_xor_sync(0xffffffffu, a[k], o);
    if ((threadIdx.x & 31) == 0) red[k][threadIdx.x >> 5] = a[k];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
    atomicAdd(out + threadIdx.x, t);
  }
}

__device__ __forceinline__ void atomic_min_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) > v) {
    const unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}
__device__ __forceinline__ void atomic_max_double(double* addr, double v) {
  unsigned long long* a = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *a;
  while (__longlong_as_double((long long)old) < v) {
    const unsigned long long prev = atomicCAS(a, old, (unsigned long long)__double_as_longlong(v));
    if (prev == old) break;
    old = prev;
  }
}

// out[0] = sum (pred - target)^2, out[1] = min(target), out[2] = max(target)   (out[1], out[2] pre-set by the launcher)
__global__ void __launch_bounds__(256) psnr_stats_kernel(const double* __restrict__ pred, const double* __restrict__ target,
                                                         int64_t n, double* __restrict__ out) {
  __shared__ double red[3][8];
  double sse = 0.0, lo = INFINITY, hi = -INFINITY;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const double t = target[i], d = pred[i] - t;
    sse += d * d;
    lo = fmin(lo, t);
    hi = fmax(hi, t);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    sse += __shfl_xor_sync(0xffffffffu, sse, o);
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sse; red[1][threadIdx.x >> 5] = lo; red[2][threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) { sse += red[0][w]; lo = fmin(lo, red[1][w]); hi = fmax(hi, red[2][w]); }
    atomicAdd(out, sse);
    atomic_min_double(out + 1, lo);
    atomic_max_double(out + 2, hi);
  }
}

}  // namespace

int launch_mask_metrics(const float* pred, const float* gt, int B, int H, int W, double* out4, cudaStream_t st) {
  MFPA_CUDA(cudaMemsetAsync(out4, 0, sizeof(double) * 4, st));
  const int64_t n = (int64_t)B * H * W;
  const unsigned blocks = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  mask_metrics_kernel<<<blocks, 256, 0, st>>>(pred, gt, n, H, W, out4);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_psnr_stats(const double* pred, const double* target, int64_t n, double* out3, cudaStream_t st) {
  const double init[3] = {0.0, INFINITY, -INFINITY};
  MFPA_CUDA(cudaMemcpyAsync(out3, init, sizeof(init), cudaMemcpyHostToDevice, st));
  MFPA_CUDA(cudaStreamSynchronize(st));   // `init` lives on this stack frame
  const unsigned blocks = (unsigned)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  psnr_stats_kernel<<<blocks, 256, 0, st>>>(pred, target, n, out3);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

}  // namespace mfpa
