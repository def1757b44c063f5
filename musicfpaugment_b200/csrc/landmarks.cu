// S4 — target-zone landmark pairing, 20-bit hash packing, shift merging, and
// the record -> list / mask conversions.
// Replaces Audfprint_peaks.peaks2landmarks (afp/audfprint/peak_extractor.py:313-346),
// landmarks2hashes (:40-58) and the concatenate/unique/sort tail of wavfile2hashes (:437-460).
//
// Input is the packed per-frame record array produced by peaks.cu
// (byte 0 = count, bytes 1..5 = bins ascending).  One warp handles one item.
// landmark_list_kernel (items of up to 1024 frames): the warp compacts the records into a
// (frame, bin)-ordered peak list in shared memory and every lane pairs one PEAK by walking the
// list - the reference's scan order without the empty frames.  landmark_kernel (longer items):
// one lane per start frame; the pairing window (frames c+2 .. c+62) is read as consecutive
// 8-byte records so every step of the window is one coalesced load.
#include "common.cuh"

namespace mfpa {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kWarpsPerBlock = 4;
constexpr int kFanMax = 3;

__device__ __forceinline__ int rec_bin(uint64_t r, int i) { return (int)((r >> (8 * (i + 1))) & 0xff); }

__global__ void __launch_bounds__(kWarpsPerBlock * 32)
landmark_kernel(const uint64_t* __restrict__ rec_all, int items, int n_max, int mindt, int targetdt,
                int targetdf, int fanout, int sorted, int32_t* __restrict__ hashes, int cap,
                int32_t* __restrict__ nh) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (item >= items) return;
  const uint64_t* rec = rec_all + (int64_t)item * n_max;
  int2* out = reinterpret_cast<int2*>(hashes) + (int64_t)item * cap;

  // scols = column of the final peak + 1 (:325)
  int scols = 0;
  for (int base = 0; base < n_max; base += 32) {
    const int c = base + lane;
    const unsigned bal = __ballot_sync(kFull, c < n_max && (rec[c] & 0xff) != 0);
    if (bal) scols = base + 32 - __clz(bal);
  }
  int written = 0;
  for (int base = 0; base < scols; base += 32) {
    const int c = base + lane;
    const uint64_t r = c < scols ? rec[c] : 0;
    const int n = (int)(r & 0xff);
    uint32_t hb[kMaxPks * kFanMax];
    int cnt[kMaxPks];
#pragma unroll
    for (int i = 0; i < kMaxPks; ++i) cnt[i] = 0;
    if (n) {
      const int c_end = min(scols, c + targetdt);
      int open = n;  // peaks that can still take pairs
      for (int c2 = c + mindt; c2 < c_end && open; ++c2) {
        const uint64_t r2 = rec[c2];
        const int n2 = (int)(r2 & 0xff);
        if (!n2) continue;
#pragma unroll
        for (int i = 0; i < kMaxPks; ++i) {
          if (i < n && cnt[i] < fanout) {
            const int b = rec_bin(r, i);
            for (int k = 0; k < n2; ++k) {
              const int b2 = rec_bin(r2, k);
              if (abs(b2 - b) < targetdf && cnt[i] < fanout) {
                hb[i * kFanMax + cnt[i]] = ((uint32_t)(b & 255) << 12) | ((uint32_t)((b2 - b) & 63) << 6) |
                                           (uint32_t)((c2 - c) & 63);
                if (++cnt[i] == fanout) --open;
              }
            }
          }
        }
      }
    }
    int mine = 0;
#pragma unroll
    for (int i = 0; i < kMaxPks; ++i) {
      if (sorted && cnt[i] > 1) {  // per-peak hashes ascending; peaks are already in bin order
        uint32_t* h = hb + i * kFanMax;
        if (h[0] > h[1]) { const uint32_t t = h[0]; h[0] = h[1]; h[1] = t; }
        if (cnt[i] > 2) {
          if (h[1] > h[2]) { const uint32_t t = h[1]; h[1] = h[2]; h[2] = t; }
          if (h[0] > h[1]) { const uint32_t t = h[0]; h[0] = h[1]; h[1] = t; }
        }
      }
      mine += cnt[i];
    }
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    int pos = written + incl - mine;
#pragma unroll
    for (int i = 0; i < kMaxPks; ++i)
      for (int k = 0; k < cnt[i]; ++k) {
        if (pos < cap) out[pos] = make_int2(c, (int)hb[i * kFanMax + k]);
        ++pos;
      }
    written += __shfl_sync(kFull, incl, 31);
  }
  if (lane == 0) nh[item] = written;  // may exceed cap: caller checks
}

// Same pairing, one lane per PEAK instead of per frame (items of up to kListFrames frames): the warp first
// compacts the item's records into a (frame, bin)-ordered peak list in shared memory with the list offset of
// every frame, then each lane walks the list from the first peak of frame c + mindt - which is exactly the
// reference's scan order (frames ascending, bins ascending) without visiting empty frames - until it has
// `fanout` pairs or leaves the target zone.  Frames hold ~0.5 peaks on music, so all lanes stay busy and the
// walk is ~30 list entries instead of 61 records x 5 slots.
constexpr int kListFrames = 1024;
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
landmark_list_kernel(const uint64_t* __restrict__ rec_all, int items, int n_max, int mindt, int targetdt,
                     int targetdf, int fanout, int sorted, int32_t* __restrict__ hashes, int cap,
                     int32_t* __restrict__ nh) {
  extern __shared__ unsigned lm_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * kWarpsPerBlock + warp;
  if (item >= items) return;
  const int per_warp = kMaxPks * n_max + n_max + 1;
  unsigned* list = lm_smem + warp * per_warp;                   // (frame << 8) | bin, frame-major
  int* off = reinterpret_cast<int*>(list + kMaxPks * n_max);    // off[c] = list index of frame c's first peak
  const uint64_t* rec = rec_all + (int64_t)item * n_max;
  int2* out = reinterpret_cast<int2*>(hashes) + (int64_t)item * cap;

  int total = 0, scols = 0;
  for (int base = 0; base < n_max; base += 32) {
    const int c = base + lane;
    const uint64_t r = c < n_max ? rec[c] : 0;
    const int n = min((int)(r & 0xff), kMaxPks);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    const int first = total + incl - n;
    if (c < n_max) off[c] = first;
    for (int i = 0; i < n; ++i) list[first + i] = ((unsigned)c << 8) | (unsigned)rec_bin(r, i);
    const unsigned bal = __ballot_sync(kFull, n != 0);
    if (bal) scols = base + 32 - __clz(bal);   // column of the final peak + 1 (:325)
    total += __shfl_sync(kFull, incl, 31);
  }
  if (lane == 0) off[n_max] = total;
  __syncwarp();

  int written = 0;
  for (int base = 0; base < total; base += 32) {
    const int i = base + lane;
    uint32_t hb[kFanMax];
    int cnt = 0, c = 0;
    if (i < total) {
      const unsigned p = list[i];
      c = (int)(p >> 8);
      const int b = (int)(p & 255u);
      const int c_end = min(scols, c + targetdt);
      for (int j = off[min(max(c + mindt, 0), n_max)]; j < total && cnt < fanout; ++j) {
        const unsigned p2 = list[j];
        const int c2 = (int)(p2 >> 8);
        if (c2 >= c_end) break;
        const int b2 = (int)(p2 & 255u);
        if (abs(b2 - b) < targetdf)
          hb[cnt++] = ((uint32_t)(b & 255) << 12) | ((uint32_t)((b2 - b) & 63) << 6) | (uint32_t)((c2 - c) & 63);
      }
      if (sorted && cnt > 1) {  // per-peak hashes ascending; peaks are already in bin order
        if (hb[0] > hb[1]) { const uint32_t t = hb[0]; hb[0] = hb[1]; hb[1] = t; }
        if (cnt > 2) {
          if (hb[1] > hb[2]) { const uint32_t t = hb[1]; hb[1] = hb[2]; hb[2] = t; }
          if (hb[0] > hb[1]) { const uint32_t t = hb[0]; hb[0] = hb[1]; hb[1] = t; }
        }
      }
    }
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    int pos = written + incl - cnt;
#pragma unroll
    for (int k = 0; k < kFanMax; ++k)
      if (k < cnt) {
        if (pos < cap) out[pos] = make_int2(c, (int)hb[k]);
        ++pos;
      }
    written += __shfl_sync(kFull, incl, 31);
  }
  if (lane == 0) nh[item] = written;  // may exceed cap: caller checks
}

// One block per query: merge `shifts` (time,hash)-sorted lists, drop duplicates.
// Thread t owns time value t: gathers the <= 15 rows per list with that time,
// insertion-merges them, then a block scan places each time's run.
__global__ void __launch_bounds__(256)
merge_shifts_kernel(const int32_t* __restrict__ hashes, const int32_t* __restrict__ nh, int shifts,
                    int cap_in, int n_frames, int32_t* __restrict__ out_all, int cap_out,
                    int32_t* __restrict__ nout) {
  __shared__ int scan[256];
  __shared__ int carry;
  const int q = blockIdx.x, tid = threadIdx.x;
  int2* out = reinterpret_cast<int2*>(out_all) + (int64_t)q * cap_out;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int t0 = 0; t0 < n_frames; t0 += 256) {
    const int t = t0 + tid;
    uint32_t buf[MFPA_HASHES_PER_FRAME * MFPA_MAX_SHIFTS];
    int n = 0;
    if (t < n_frames) {
      for (int s = 0; s < shifts; ++s) {
        const int item = q * shifts + s;
        const int2* lst = reinterpret_cast<const int2*>(hashes) + (int64_t)item * cap_in;
        const int len = min(nh[item], cap_in);
        int lo = 0, hi = len;  // first row with time >= t
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (lst[mid].x < t) lo = mid + 1; else hi = mid;
        }
        for (; lo < len; ++lo) {
          const int2 row = lst[lo];
          if (row.x != t) break;
          const uint32_t h = (uint32_t)row.y;
          int k = n;
          while (k > 0 && buf[k - 1] > h) --k;
          if (k > 0 && buf[k - 1] == h) continue;  // duplicate (time, hash)
          if (n < MFPA_HASHES_PER_FRAME * MFPA_MAX_SHIFTS) {
            for (int m = n; m > k; --m) buf[m] = buf[m - 1];
            buf[k] = h;
            ++n;
          }
        }
      }
    }
    scan[tid] = n;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
      const int v = tid >= o ? scan[tid - o] : 0;
      __syncthreads();
      scan[tid] += v;
      __syncthreads();
    }
    int pos = carry + scan[tid] - n;
    for (int k = 0; k < n; ++k, ++pos)
      if (pos < cap_out) out[pos] = make_int2(t, (int)buf[k]);
    __syncthreads();
    if (tid == 255) carry += scan[255];
    __syncthreads();
  }
  if (tid == 0) nout[q] = carry;
}

// records -> (col, bin) rows, pklist order (:303-309)
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
peaks_list_kernel(const uint64_t* __restrict__ rec_all, int items, int n_max, int32_t* __restrict__ peaks,
                  int cap, int32_t* __restrict__ npeaks) {
  const int lane = threadIdx.x & 31;
  const int item = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (item >= items) return;
  const uint64_t* rec = rec_all + (int64_t)item * n_max;
  int2* out = reinterpret_cast<int2*>(peaks) + (int64_t)item * cap;
  int written = 0;
  for (int base = 0; base < n_max; base += 32) {
    const int c = base + lane;
    const uint64_t r = c < n_max ? rec[c] : 0;
    const int n = (int)(r & 0xff);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    int pos = written + incl - n;
    for (int i = 0; i < n; ++i, ++pos)
      if (pos < cap) out[pos] = make_int2(c, rec_bin(r, i));
    written += __shfl_sync(kFull, incl, 31);
  }
  if (lane == 0 && npeaks) npeaks[item] = written;
}

// records -> float mask [items][256][n_max] (already zeroed)
__global__ void peaks_mask_kernel(const uint64_t* __restrict__ rec, int64_t total, int n_max, float* __restrict__ mask) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint64_t r = rec[i];
  const int n = (int)(r & 0xff);
  const int64_t item = i / n_max;
  const int c = (int)(i - item * n_max);
  for (int k = 0; k < n; ++k) mask[(item * kRows + rec_bin(r, k)) * n_max + c] = 1.0f;
}

// exclusive scan of min(n[i], cap) -> offsets[items+1] (one block; items is O(10^4))
__global__ void __launch_bounds__(1024) offsets_scan_kernel(const int32_t* __restrict__ n, int items, int cap,
                                                            int64_t* __restrict__ offsets) {
  __shared__ int64_t warp_tot[32];
  __shared__ int64_t carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < items; base += 1024) {
    const int i = base + tid;
    const int64_t v = i < items ? (int64_t)min(n[i], cap) : 0;
    int64_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t t = __shfl_up_sync(kFull, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int64_t w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t t = __shfl_up_sync(kFull, wi, o);
        if (lane >= o) wi += t;
      }
      warp_tot[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    const int64_t excl = carry + warp_tot[warp] + incl - v;
    if (i < items) offsets[i] = excl;
    __syncthreads();
    if (tid == 1023) carry = excl + v;
    __syncthreads();
  }
  if (tid == 0) offsets[items] = carry;
}

__global__ void __launch_bounds__(128) compact_rows_kernel(const int32_t* __restrict__ rows_in, const int32_t* __restrict__ n,
                                                           int cap, const int64_t* __restrict__ offsets,
                                                           int32_t* __restrict__ rows, int64_t rows_cap) {
  const int item = blockIdx.x;
  const int cnt = min(n[item], cap);
  const int64_t off = offsets[item];
  const int2* src = reinterpret_cast<const int2*>(rows_in) + (int64_t)item * cap;
  int2* dst = reinterpret_cast<int2*>(rows);
  for (int k = threadIdx.x; k < cnt; k += blockDim.x)
    if (off + k < rows_cap) dst[off + k] = src[k];
}

}  // namespace

int launch_compact_rows(const int32_t* rows_in, const int32_t* n, int items, int cap, int64_t* offsets,
                        int32_t* rows, int64_t rows_cap, cudaStream_t st) {
  offsets_scan_kernel<<<1, 1024, 0, st>>>(n, items, cap, offsets);
  MFPA_CUDA(cudaGetLastError());
  compact_rows_kernel<<<items, 128, 0, st>>>(rows_in, n, cap, offsets, rows, rows_cap);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_landmark_hashes(const uint64_t* rec, int items, int n_frames, const mfpa_afp_params& p,
                           int sorted, int32_t* hashes, int cap, int32_t* nh, cudaStream_t st) {
  MFPA_REQUIRE(p.fanout >= 1 && p.fanout <= kFanMax, "landmarks: fanout %d not in 1..%d", p.fanout, kFanMax);
  const int blocks = (items + kWarpsPerBlock - 1) / kWarpsPerBlock;
  if (n_frames <= kListFrames) {
    const size_t smem = sizeof(unsigned) * kWarpsPerBlock * ((size_t)kMaxPks * n_frames + n_frames + 1);
    MFPA_CUDA(cudaFuncSetAttribute(landmark_list_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    landmark_list_kernel<<<blocks, kWarpsPerBlock * 32, smem, st>>>(rec, items, n_frames, p.mindt, p.targetdt, p.targetdf,
                                                                    p.fanout, sorted, hashes, cap, nh);
  } else {
    landmark_kernel<<<blocks, kWarpsPerBlock * 32, 0, st>>>(rec, items, n_frames, p.mindt, p.targetdt, p.targetdf,
                                                            p.fanout, sorted, hashes, cap, nh);
  }
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_merge_shifts(const int32_t* hashes, const int32_t* nh, int B, int shifts, int cap_in,
                        int n_frames, int32_t* out, int cap_out, int32_t* nout, cudaStream_t st) {
  MFPA_REQUIRE(shifts >= 1 && shifts <= MFPA_MAX_SHIFTS, "merge: shifts %d not in 1..%d", shifts, MFPA_MAX_SHIFTS);
  merge_shifts_kernel<<<B, 256, 0, st>>>(hashes, nh, shifts, cap_in, n_frames, out, cap_out, nout);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_peaks_list(const uint64_t* rec, int items, int n_frames, int32_t* peaks, int cap,
                      int32_t* npeaks, cudaStream_t st) {
  const int blocks = (items + kWarpsPerBlock - 1) / kWarpsPerBlock;
  peaks_list_kernel<<<blocks, kWarpsPerBlock * 32, 0, st>>>(rec, items, n_frames, peaks, cap, npeaks);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_peaks_mask(const uint64_t* rec, int items, int n_frames, float* mask, cudaStream_t st) {
  const int64_t total = (int64_t)items * n_frames;
  MFPA_CUDA(cudaMemsetAsync(mask, 0, sizeof(float) * total * kRows, st));
  peaks_mask_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(rec, total, n_frames, mask);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

}  // namespace mfpa
