// UNet magnitude-spectrogram denoiser forward (reference: training/unet.py:8-108, called at
// afp/audfprint/peak_extractor.py:265-269 and afp/dejavu/fingerprint.py:70-75).
//
// The 18 3x3 convolutions and the 4 stride-2 transposed convolutions (99.9 % of the 93.4 GFLOP per
// 257x251 spectrogram) run as bf16 implicit GEMMs on the tcgen05 tensor cores:
//
//   D[pixel, cout] = sum over (tap, cin) A[pixel shifted by tap, cin] * W[cout, tap, cin]
//
//   * activations live in HBM as NHWC bf16, so one (tap, 64-channel) slice of an 8x16 / 16x16 pixel
//     tile is a 4-D TMA box {64 ch, 16 w, 8|16 h, 1 n} at the shifted coordinates; the box lands in
//     shared memory as a 128-byte-swizzled K-major [pixels][64] operand and TMA's out-of-bounds zero
//     fill IS the convolution's zero padding (no im2col buffer, no border code);
//   * weights are [cout][tap][cin] bf16, a 2-D TMA box {64, BN};
//   * one elected thread issues tcgen05.mma (M = 128 pixels, N = BN, K = 16) into TMEM accumulators,
//     a ring of mbarrier-guarded stages decouples the TMA warp from the MMA warp, and four epilogue
//     warps read TMEM with tcgen05.ld, apply the folded BatchNorm scale/shift + ReLU and store bf16
//     NHWC (optionally into a channel slice of a skip-concatenation buffer, as the 2x2 scatter of a
//     transposed convolution, or fused with the final 1x1 convolution).
//
// The 1 -> 64 input convolution, the 2x2 max pools and nothing else are plain CUDA-core kernels.
#include <cuda.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace {

using bf16 = __nv_bfloat16;

constexpr int kTileW = 16;        // pixels per tile row
constexpr int kBlockK = 64;       // channels per pipeline stage (64 bf16 = one 128-byte swizzle row)
constexpr int kATile = 128 * 128; // bytes of one 128-pixel x 64-channel operand tile
constexpr int kThreads = 192;     // warps 0-3 epilogue, warp 4 TMA producer, warp 5 MMA issuer

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart
// (SmemDescriptor fields: start address >> 4 | LBO (unused, 1) << 16 | SBO (1024 >> 4) << 32 |
//  version 1 << 46 | layout SWIZZLE_128B (2) << 61).
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
// kind::f16 instruction descriptor: D fp32 (bit 4), A and B bf16 (bits 7, 10), both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24.
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// ------------------------------------------------------------------------------------ the GEMM kernel
enum EpiMode : int { kEpiStore = 0, kEpiUpscatter = 1, kEpiOutc = 2 };

struct ConvArgs {
  int H, W;          // input (= output for 3x3) spatial size
  int cin;           // input channels (multiple of 64)
  int taps;          // 9 (3x3, pad 1) or 1 (1x1 / transposed-convolution GEMM)
  int stages;        // pipeline depth
  int mode, relu;
  const float* scale;  // [cout] folded BatchNorm scale (or ones)
  const float* shift;  // [cout] folded BatchNorm shift / bias
  bf16* out;           // NHWC bf16
  int out_h, out_w;    // spatial size of `out` (mode 1: the skip tensor's size)
  int ldc, coff;       // channels per pixel of `out`, first channel written
  int cout;            // output channels (mode 1: channels per (a, b) phase; GEMM N = 4 * cout)
  int n_total;         // GEMM N (cout, or 4 * cout for the transposed convolution)
  int N;               // images
  // mode 2: fused 1x1 output convolution (training/unet.py:66-72)
  const float* w_out;  // [64]
  float b_out;
  float* out_f32;
  long long of_n, of_h, of_w;  // element strides of out_f32
};

// Persistent: gridDim.x CTAs walk the (image, pixel tile, channel tile) list with stride gridDim.x.
// The smem ring runs across tile boundaries (the TMA warp prefetches the next tile while the epilogue
// drains this one) and, when 2 * MT * BN <= 512 TMEM columns, accumulators are double-buffered so the
// MMAs of tile i+1 overlap the epilogue of tile i.
template <int BN, int MT>
__global__ void __launch_bounds__(kThreads, 1)
conv_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const ConvArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  constexpr int kStageBytes = MT * kATile + BN * 128;
  constexpr int kAcc = (2 * MT * BN <= 512) ? 2 : 1;   // TMEM accumulator buffers
  constexpr int kTmemCols = kAcc * MT * BN;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + (size_t)a.stages * kStageBytes);  // full[S], empty[S], tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * a.stages + 4);
  float* s_scale = (float*)(tmem_slot + 2);
  float* s_shift = s_scale + a.cout;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_w = (a.W + kTileW - 1) / kTileW;
  const int tiles_sp = tiles_w * ((a.H + 8 * MT - 1) / (8 * MT));
  const int tiles_n = a.n_total / BN;
  const int total = tiles_n * tiles_sp * a.N;
  const int nkb = a.taps * (a.cin / kBlockK);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (a.stages + s); };
  auto tmem_full_bar = [&](int b) { return bar0 + 8u * (2 * a.stages + b); };
  auto tmem_empty_bar = [&](int b) { return bar0 + 8u * (2 * a.stages + 2 + b); };
  // tile index -> (channel tile fastest, so CTAs running side by side share the activation tile in L2)
  auto decode = [&](int t, int& n0, int& w0, int& h0, int& img) {
    n0 = (t % tiles_n) * BN;
    t /= tiles_n;
    const int sp = t % tiles_sp;
    img = t / tiles_sp;
    w0 = (sp % tiles_w) * kTileW;
    h0 = (sp / tiles_w) * (8 * MT);
  };

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < a.stages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < 4) {
    for (int i = threadIdx.x; i < a.cout; i += 128) {
      s_scale[i] = a.scale[i];
      s_shift[i] = a.shift[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      // ---- TMA producer: one (tap, 64-channel) slice of the pixel tile + the matching weight slice per stage
      const int chunks = a.cin / kBlockK;
      int it = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int n0, w0, h0, img;
        decode(t, n0, w0, h0, img);
        for (int tap = 0; tap < a.taps; ++tap) {
          const int dh = a.taps == 9 ? tap / 3 - 1 : 0, dw = a.taps == 9 ? tap % 3 - 1 : 0;
          for (int cc = 0; cc < chunks; ++cc, ++it) {
            const int s = it % a.stages;
            const uint32_t ph = (it / a.stages) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            mbar_expect_tx(full_bar(s), kStageBytes);
            const uint32_t dst = smem_u32(smem + (size_t)s * kStageBytes);
            tma_load_4d(dst, &tmA, full_bar(s), cc * kBlockK, w0 + dw, h0 + dh, img);
            tma_load_2d(dst + MT * kATile, &tmB, full_bar(s), tap * a.cin + cc * kBlockK, n0);
          }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // ---- MMA issuer
      constexpr uint32_t idesc = umma_idesc(128, BN);
      int it = 0, i = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
        const int acc = i % kAcc;
        const uint32_t acc_ph = (i / kAcc) & 1;
        mbar_wait(tmem_empty_bar(acc), acc_ph ^ 1);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * (MT * BN);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % a.stages;
          const uint32_t ph = (it / a.stages) & 1;
          mbar_wait(full_bar(s), ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)s * kStageBytes);
          const uint32_t sb = sa + MT * kATile;
#pragma unroll
          for (int j = 0; j < MT; ++j) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k) {
              tc_mma(d0 + j * BN, umma_desc(sa + j * kATile + k * 32), umma_desc(sb + k * 32), idesc,
                     (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          tc_commit(empty_bar(s));  // frees the stage when the MMAs above have read it
        }
        tc_commit(tmem_full_bar(acc));
      }
    }
  } else {
    // ---- epilogue: warp w owns TMEM lanes 32w..32w+31 = pixels 32w..32w+31 of each 128-pixel sub-tile
    const int m = warp * 32 + lane;
    int i = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
      const int acc = i % kAcc;
      const uint32_t acc_ph = (i / kAcc) & 1;
      int n0, w0, h0, img;
      decode(t, n0, w0, h0, img);
      mbar_wait(tmem_full_bar(acc), acc_ph);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < MT; ++j) {
        const int h = h0 + 8 * j + (m >> 4), w = w0 + (m & 15);
        const bool inside = h < a.H && w < a.W;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * (MT * BN) + j * BN;
        if (a.mode == kEpiOutc) {
          float dot = 0.f;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tc_ld32(taddr + c0, v);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              float y = fmaf(__uint_as_float(v[q]), s_scale[n0 + c0 + q], s_shift[n0 + c0 + q]);
              y = fmaxf(y, 0.f);
              dot = fmaf(y, __ldg(a.w_out + n0 + c0 + q), dot);
            }
          }
          if (inside) a.out_f32[img * a.of_n + h * a.of_h + w * a.of_w] = dot + a.b_out;
        } else {
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tc_ld32(taddr + c0, v);
            // output channel / pixel of this 32-column group
            int co = n0 + c0;
            size_t pix;
            if (a.mode == kEpiUpscatter) {
              const int phase = co / a.cout;  // (a, b) of ConvTranspose2d(k=2, s=2): out[2h+a][2w+b]
              co -= phase * a.cout;
              pix = ((size_t)img * a.out_h + (2 * h + (phase >> 1))) * a.out_w + (2 * w + (phase & 1));
            } else {
              pix = ((size_t)img * a.out_h + h) * a.out_w + w;
            }
            uint32_t packed[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              float y0 = fmaf(__uint_as_float(v[2 * q]), s_scale[co + 2 * q], s_shift[co + 2 * q]);
              float y1 = fmaf(__uint_as_float(v[2 * q + 1]), s_scale[co + 2 * q + 1], s_shift[co + 2 * q + 1]);
              if (a.relu) {
                y0 = fmaxf(y0, 0.f);
                y1 = fmaxf(y1, 0.f);
              }
              __nv_bfloat162 pk = __floats2bfloat162_rn(y0, y1);
              packed[q] = *reinterpret_cast<uint32_t*>(&pk);
            }
            if (inside) {
              uint4* d4 = reinterpret_cast<uint4*>(a.out + pix * a.ldc + a.coff + co);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                d4[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------ halo-reuse 3x3 kernel
// The per-tap kernel above fetches every activation tile nine times from L2, and L2 -> SM bandwidth
// (~42 B/clk/SM) is what bounds it.  This variant loads the (th+2) x Wh halo of a tile ONCE per 64-channel
// chunk and feeds all nine taps from it: with the halo stored one 128-byte swizzled row per pixel, pitch
// Wh pixels, the operand of tap (kh, kw) is simply the same buffer starting (kh*Wh + kw) rows later - M
// runs over flat halo positions p = h*Wh + w, rows with w >= Wh-2 are computed and thrown away.  Such a
// start is 128-byte but not 1024-byte aligned; measured on B200, the tensor core applies the 128-byte
// swizzle to absolute shared-memory address bits, so the descriptor's base-offset field stays 0 and the
// pattern TMA wrote (also a function of the absolute address) is read back correctly.
// WRES keeps the whole weight tensor of a Cout = BN layer resident in shared memory (the 64-channel layers
// at full resolution, whose weights would otherwise be re-read from L2 for every tile).
struct HaloArgs {
  ConvArgs c;
  int wh;          // halo pitch in pixels (tile width = wh - 2)
  int th;          // output rows per tile
  int a_bytes;     // bytes of one halo stage (multiple of 1024)
  int a_stages, b_stages;
};

template <int BN, int MT, bool WRES>
__global__ void __launch_bounds__(kThreads, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const HaloArgs ha) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const ConvArgs& a = ha.c;
  constexpr int kBBytes = BN * 128;
  constexpr int kAcc = (2 * MT * BN <= 512) ? 2 : 1;
  constexpr int kTmemCols = kAcc * MT * BN;
  const int chunks = a.cin / kBlockK;
  const int nb = WRES ? 9 * chunks : ha.b_stages;  // weight tiles held in shared memory
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_b = smem + (size_t)ha.a_stages * ha.a_bytes;
  uint64_t* bars = (uint64_t*)(smem_b + (size_t)nb * kBBytes);
  // a_full[SA], a_empty[SA], b_full[SB], b_empty[SB], tmem_full[2], tmem_empty[2], wres_full
  const int SA = ha.a_stages, SB = ha.b_stages;
  uint32_t* tmem_slot = (uint32_t*)(bars + 2 * SA + 2 * SB + 5);
  float* s_scale = (float*)(tmem_slot + 2);
  float* s_shift = s_scale + a.cout;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tw = ha.wh - 2;
  const int tiles_w = (a.W + tw - 1) / tw;
  const int tiles_sp = tiles_w * ((a.H + ha.th - 1) / ha.th);
  const int tiles_n = a.n_total / BN;
  const int total = tiles_n * tiles_sp * a.N;
  const uint32_t bar0 = smem_u32(bars);
  auto a_full = [&](int s) { return bar0 + 8u * s; };
  auto a_empty = [&](int s) { return bar0 + 8u * (SA + s); };
  auto b_full = [&](int s) { return bar0 + 8u * (2 * SA + s); };
  auto b_empty = [&](int s) { return bar0 + 8u * (2 * SA + SB + s); };
  auto tmem_full_bar = [&](int b) { return bar0 + 8u * (2 * SA + 2 * SB + b); };
  auto tmem_empty_bar = [&](int b) { return bar0 + 8u * (2 * SA + 2 * SB + 2 + b); };
  const uint32_t wres_bar = bar0 + 8u * (2 * SA + 2 * SB + 4);
  auto decode = [&](int t, int& n0, int& w0, int& h0, int& img) {
    n0 = (t % tiles_n) * BN;
    t /= tiles_n;
    const int sp = t % tiles_sp;
    img = t / tiles_sp;
    w0 = (sp % tiles_w) * tw;
    h0 = (sp / tiles_w) * ha.th;
  };

  if (warp == 4 && lane == 0) {
    for (int s = 0; s < SA; ++s) {
      mbar_init(a_full(s), 1);
      mbar_init(a_empty(s), 1);
    }
    for (int s = 0; s < SB; ++s) {
      mbar_init(b_full(s), 1);
      mbar_init(b_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), 4);
    }
    mbar_init(wres_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"((uint32_t)kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp < 4) {
    for (int i = threadIdx.x; i < a.cout; i += 128) {
      s_scale[i] = a.scale[i];
      s_shift[i] = a.shift[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t halo_box_bytes = (uint32_t)(ha.th + 2) * ha.wh * 128u;

  if (warp == 4) {
    if (lane == 0) {
      // ---- TMA producer
      if (WRES) {
        mbar_expect_tx(wres_bar, (uint32_t)(9 * chunks) * kBBytes);
        for (int tap = 0; tap < 9; ++tap)
          for (int cc = 0; cc < chunks; ++cc)
            tma_load_2d(smem_u32(smem_b + (size_t)(tap * chunks + cc) * kBBytes), &tmB, wres_bar, tap * a.cin + cc * kBlockK, 0);
      }
      int ia = 0, ib = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        int n0, w0, h0, img;
        decode(t, n0, w0, h0, img);
        for (int cc = 0; cc < chunks; ++cc, ++ia) {
          const int sa = ia % SA;
          mbar_wait(a_empty(sa), ((ia / SA) & 1) ^ 1);
          mbar_expect_tx(a_full(sa), halo_box_bytes);
          tma_load_4d(smem_u32(smem + (size_t)sa * ha.a_bytes), &tmA, a_full(sa), cc * kBlockK, w0 - 1, h0 - 1, img);
          if (!WRES) {
            for (int tap = 0; tap < 9; ++tap, ++ib) {
              const int sb = ib % SB;
              mbar_wait(b_empty(sb), ((ib / SB) & 1) ^ 1);
              mbar_expect_tx(b_full(sb), kBBytes);
              tma_load_2d(smem_u32(smem_b + (size_t)sb * kBBytes), &tmB, b_full(sb), tap * a.cin + cc * kBlockK, n0);
            }
          }
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      // ---- MMA issuer
      constexpr uint32_t idesc = umma_idesc(128, BN);
      if (WRES) {
        mbar_wait(wres_bar, 0);
        tc_fence_after();
      }
      int ia = 0, ib = 0, i = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
        const int acc = i % kAcc;
        mbar_wait(tmem_empty_bar(acc), ((i / kAcc) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * (MT * BN);
        for (int cc = 0; cc < chunks; ++cc, ++ia) {
          const int sa = ia % SA;
          mbar_wait(a_full(sa), (ia / SA) & 1);
          tc_fence_after();
          const uint32_t abase = smem_u32(smem + (size_t)sa * ha.a_bytes);
          for (int tap = 0; tap < 9; ++tap) {
            uint32_t sbaddr;
            int sb = 0;
            if (WRES) {
              sbaddr = smem_u32(smem_b + (size_t)(tap * chunks + cc) * kBBytes);
            } else {
              sb = ib % SB;
              mbar_wait(b_full(sb), (ib / SB) & 1);
              tc_fence_after();
              sbaddr = smem_u32(smem_b + (size_t)sb * kBBytes);
              ++ib;
            }
            const uint32_t arow = abase + (uint32_t)((tap / 3) * ha.wh + tap % 3) * 128u;
#pragma unroll
            for (int j = 0; j < MT; ++j) {
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                tc_mma(d0 + j * BN, umma_desc(arow + j * kATile + k * 32), umma_desc(sbaddr + k * 32), idesc,
                       (cc > 0 || tap > 0 || k > 0) ? 1u : 0u);
              }
            }
            if (!WRES) tc_commit(b_empty(sb));
          }
          tc_commit(a_empty(sa));
        }
        tc_commit(tmem_full_bar(acc));
      }
    }
  } else {
    // ---- epilogue: TMEM lane m of sub-tile j = flat halo position j*128 + m
    int i = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++i) {
      const int acc = i % kAcc;
      int n0, w0, h0, img;
      decode(t, n0, w0, h0, img);
      mbar_wait(tmem_full_bar(acc), (i / kAcc) & 1);
      tc_fence_after();
#pragma unroll 1
      for (int j = 0; j < MT; ++j) {
        const int p = j * 128 + warp * 32 + lane;
        const int hl = p / ha.wh, wl = p - hl * ha.wh;
        const int h = h0 + hl, w = w0 + wl;
        const bool inside = wl < tw && hl < ha.th && h < a.H && w < a.W;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * (MT * BN) + j * BN;
        if (a.mode == kEpiOutc) {
          float dot = 0.f;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tc_ld32(taddr + c0, v);
#pragma unroll
            for (int q = 0; q < 32; ++q) {
              float y = fmaf(__uint_as_float(v[q]), s_scale[n0 + c0 + q], s_shift[n0 + c0 + q]);
              y = fmaxf(y, 0.f);
              dot = fmaf(y, __ldg(a.w_out + n0 + c0 + q), dot);
            }
          }
          if (inside) a.out_f32[img * a.of_n + h * a.of_h + w * a.of_w] = dot + a.b_out;
        } else {
          const size_t pix = ((size_t)img * a.out_h + h) * a.out_w + w;
#pragma unroll 1
          for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tc_ld32(taddr + c0, v);
            const int co = n0 + c0;
            uint32_t packed[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
              float y0 = fmaf(__uint_as_float(v[2 * q]), s_scale[co + 2 * q], s_shift[co + 2 * q]);
              float y1 = fmaf(__uint_as_float(v[2 * q + 1]), s_scale[co + 2 * q + 1], s_shift[co + 2 * q + 1]);
              if (a.relu) {
                y0 = fmaxf(y0, 0.f);
                y1 = fmaxf(y1, 0.f);
              }
              __nv_bfloat162 pk = __floats2bfloat162_rn(y0, y1);
              packed[q] = *reinterpret_cast<uint32_t*>(&pk);
            }
            if (inside) {
              uint4* d4 = reinterpret_cast<uint4*>(a.out + pix * a.ldc + a.coff + co);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                d4[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------ CUDA-core kernels
// inc.double_conv.0: Conv2d(1, 64, 3, padding=1, bias=False) + BatchNorm + ReLU (training/unet.py:15-18).
// in: float32, element strides (in_n, in_h, in_w), optionally divided by div[n] (sgram /= max,
// peak_extractor.py:263); out: NHWC bf16 [N][H][W][64].  A block stages the 10 x 34 input halo of an
// 8 x 32 pixel tile in shared memory (read along whichever input dimension has unit stride), then every
// thread produces 8 channels of a pixel so that a warp stores 4 pixels x 128 contiguous bytes.
constexpr int kInTh = 32, kInTw = 32;   // 32 rows per block: the 576 weights each thread keeps in registers are loaded once per 32 pixel rows
__global__ void __launch_bounds__(256)
conv_in_kernel(const float* __restrict__ in, long long in_n, long long in_h, long long in_w, const float* __restrict__ div,
               int H, int W, const float* __restrict__ wgt /* [9][64] */, const float* __restrict__ scale,
               const float* __restrict__ shift, bf16* __restrict__ out) {
  __shared__ float s_w[9 * 64], s_sc[64], s_sh[64];
  __shared__ float s_x[kInTh + 2][kInTw + 2 + 1];
  for (int i = threadIdx.x; i < 9 * 64; i += blockDim.x) s_w[i] = wgt[i];
  if (threadIdx.x < 64) {
    s_sc[threadIdx.x] = scale[threadIdx.x];
    s_sh[threadIdx.x] = shift[threadIdx.x];
  }
  const int n = blockIdx.z, h0 = blockIdx.y * kInTh, w0 = blockIdx.x * kInTw;
  const float dv = div ? div[n] : 1.0f;
  const float* base = in + n * in_n;
  constexpr int kHalo = (kInTh + 2) * (kInTw + 2);
  for (int i = threadIdx.x; i < kHalo; i += blockDim.x) {
    int r, c;
    if (in_h == 1) {  // frame-major magnitudes: consecutive threads walk the bins
      r = i % (kInTh + 2);
      c = i / (kInTh + 2);
    } else {
      r = i / (kInTw + 2);
      c = i % (kInTw + 2);
    }
    const int hh = h0 + r - 1, ww = w0 + c - 1;
    s_x[r][c] = (hh >= 0 && hh < H && ww >= 0 && ww < W) ? base[hh * in_h + ww * in_w] / dv : 0.f;
  }
  __syncthreads();
  const int cg = threadIdx.x & 7;  // channels 8*cg .. 8*cg+7
  float wr[9][8], sc[8], sh[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    sc[q] = s_sc[8 * cg + q];
    sh[q] = s_sh[8 * cg + q];
#pragma unroll
    for (int t = 0; t < 9; ++t) wr[t][q] = s_w[t * 64 + 8 * cg + q];
  }
#pragma unroll 1
  for (int it = 0; it < kInTh; ++it) {
    const int p = it * 32 + (threadIdx.x >> 3);  // 32 pixels (one tile row) per pass
    const int r = p / kInTw, c = p % kInTw;
    const int h = h0 + r, w = w0 + c;
    if (h >= H || w >= W) continue;
    float x[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) x[t] = s_x[r + t / 3][c + t % 3];
    uint32_t packed[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        a0 = fmaf(x[t], wr[t][2 * q], a0);
        a1 = fmaf(x[t], wr[t][2 * q + 1], a1);
      }
      a0 = fmaxf(fmaf(a0, sc[2 * q], sh[2 * q]), 0.f);
      a1 = fmaxf(fmaf(a1, sc[2 * q + 1], sh[2 * q + 1]), 0.f);
      __nv_bfloat162 pk = __floats2bfloat162_rn(a0, a1);
      packed[q] = *reinterpret_cast<uint32_t*>(&pk);
    }
    *reinterpret_cast<uint4*>(out + (((size_t)n * H + h) * W + w) * 64 + 8 * cg) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  }
}

// nn.MaxPool2d(2) (training/unet.py:33): in NHWC with ld_in channels per pixel (first C used),
// out NHWC [N][H/2][W/2][C].  One thread per 8 channels of one output pixel.
__global__ void __launch_bounds__(256)
maxpool_kernel(const bf16* __restrict__ in, int H, int W, int ld_in, int C, bf16* __restrict__ out, long long total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = C / 8;
  const int Ho = H / 2, Wo = W / 2;
  const int cg = (int)(i % c8);
  long long p = i / c8;
  const int wo = (int)(p % Wo);
  p /= Wo;
  const int ho = (int)(p % Ho);
  const int n = (int)(p / Ho);
  const bf16* src = in + (((size_t)n * H + 2 * ho) * W + 2 * wo) * ld_in + cg * 8;
  uint4 q[4];
  q[0] = *reinterpret_cast<const uint4*>(src);
  q[1] = *reinterpret_cast<const uint4*>(src + ld_in);
  q[2] = *reinterpret_cast<const uint4*>(src + (size_t)W * ld_in);
  q[3] = *reinterpret_cast<const uint4*>(src + (size_t)W * ld_in + ld_in);
  uint4 r;
  uint32_t* rr = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    __nv_bfloat162 m = *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const uint32_t*>(&q[0]) + k);
#pragma unroll
    for (int t = 1; t < 4; ++t)
      m = __hmax2(m, *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const uint32_t*>(&q[t]) + k));
    rr[k] = *reinterpret_cast<uint32_t*>(&m);
  }
  *reinterpret_cast<uint4*>(out + (((size_t)n * Ho + ho) * Wo + wo) * C + cg * 8) = r;
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// NHWC bf16 activation [N][H][W][C] -> 4-D map, box {64, 16, box_h, 1}, 128-byte swizzle, zero OOB fill
int make_act_map(CUtensorMap* m, const bf16* ptr, int N, int H, int W, int C, int box_h, int box_w = kTileW) {
  EncodeTiledFn fn = encode_fn();
  MFPA_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MFPA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation %dx%dx%dx%d) failed: %d", N, H, W, C, (int)r);
  return MFPA_OK;
}
// weights [rows][K] bf16 -> 2-D map, box {64, box_rows}
int make_wgt_map(CUtensorMap* m, const bf16* ptr, int rows, int K, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  MFPA_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MFPA_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights %dx%d) failed: %d", rows, K, (int)r);
  return MFPA_OK;
}

struct GemmLaunch {
  CUtensorMap tmA, tmB;
  ConvArgs args;
  int bn, mt, n_total, N;
  int ctas = mfpa::kNumSMs;  // persistent grid size
  // halo-reuse variant (wh > 0)
  int wh = 0, th = 0, wres = 0, a_stages = 0, b_stages = 0, a_bytes = 0;
};

// geometry of the halo kernel for pitch wh: rows per tile and bytes of one halo stage
void halo_geometry(int wh, int mt, int* th, int* a_bytes) {
  const int tw = wh - 2;
  *th = (128 * mt - tw) / wh + 1;
  int rows = (*th + 2) * wh;
  const int reach = 128 * mt + 2 * wh + 2;  // last row any tap's operand touches
  if (rows < reach) rows = reach;
  *a_bytes = (rows * 128 + 1023) / 1024 * 1024;
}
size_t halo_smem_bytes(const GemmLaunch& g, int cin, int cout) {
  const int nb = g.wres ? 9 * (cin / kBlockK) : g.b_stages;
  return 1024 + (size_t)g.a_stages * g.a_bytes + (size_t)nb * g.bn * 128 + 8 * (2 * g.a_stages + 2 * g.b_stages + 5) + 8 +
         2 * (size_t)cout * sizeof(float);
}

template <int BN, int MT, bool WRES>
int launch_halo_t(const GemmLaunch& g, cudaStream_t st) {
  MFPA_CUDA(cudaFuncSetAttribute(conv_halo_kernel<BN, MT, WRES>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  HaloArgs ha{};
  ha.c = g.args;
  ha.c.N = g.N;
  ha.c.n_total = g.n_total;
  ha.wh = g.wh;
  ha.th = g.th;
  ha.a_bytes = g.a_bytes;
  ha.a_stages = g.a_stages;
  ha.b_stages = g.b_stages;
  const int tw = g.wh - 2;
  const long long tiles = (long long)((ha.c.W + tw - 1) / tw) * ((ha.c.H + g.th - 1) / g.th) * (g.n_total / BN) * g.N;
  MFPA_REQUIRE(tiles < (1ll << 31), "conv halo: too many tiles");
  const int grid = (int)(tiles < g.ctas ? tiles : g.ctas);
  conv_halo_kernel<BN, MT, WRES><<<grid, kThreads, halo_smem_bytes(g, ha.c.cin, ha.c.cout), st>>>(g.tmA, g.tmB, ha);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_halo(const GemmLaunch& g, cudaStream_t st) {
  switch (g.bn * 100 + g.mt * 10 + g.wres) {
    case 6410: return launch_halo_t<64, 1, false>(g, st);
    case 6411: return launch_halo_t<64, 1, true>(g, st);
    case 6420: return launch_halo_t<64, 2, false>(g, st);
    case 6421: return launch_halo_t<64, 2, true>(g, st);
    case 12810: return launch_halo_t<128, 1, false>(g, st);
    case 12820: return launch_halo_t<128, 2, false>(g, st);
    case 25610: return launch_halo_t<256, 1, false>(g, st);
    case 25620: return launch_halo_t<256, 2, false>(g, st);
  }
  mfpa::set_error("conv halo: unsupported configuration BN=%d MT=%d WRES=%d", g.bn, g.mt, g.wres);
  return MFPA_EINVAL;
}

size_t gemm_smem_bytes(int mt, int bn, int stages, int cout) {
  return 1024 + (size_t)stages * (mt * kATile + bn * 128) + 8 * (2 * stages + 4) + 8 + 2 * (size_t)cout * sizeof(float);
}

template <int BN, int MT>
int launch_gemm_t(const GemmLaunch& g, cudaStream_t st) {
  MFPA_CUDA(cudaFuncSetAttribute(conv_gemm_kernel<BN, MT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  ConvArgs a = g.args;
  a.N = g.N;
  a.n_total = g.n_total;
  const long long tiles = (long long)((a.W + kTileW - 1) / kTileW) * ((a.H + 8 * MT - 1) / (8 * MT)) * (g.n_total / BN) * g.N;
  MFPA_REQUIRE(tiles < (1ll << 31), "conv gemm: too many tiles");
  const int grid = (int)(tiles < g.ctas ? tiles : g.ctas);
  conv_gemm_kernel<BN, MT><<<grid, kThreads, gemm_smem_bytes(MT, BN, a.stages, a.cout), st>>>(g.tmA, g.tmB, a);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_gemm(const GemmLaunch& g, cudaStream_t st) {
  if (g.wh > 0) return launch_halo(g, st);
  switch (g.bn * 10 + g.mt) {
    case 641: return launch_gemm_t<64, 1>(g, st);
    case 642: return launch_gemm_t<64, 2>(g, st);
    case 1281: return launch_gemm_t<128, 1>(g, st);
    case 1282: return launch_gemm_t<128, 2>(g, st);
    case 2561: return launch_gemm_t<256, 1>(g, st);
    case 2562: return launch_gemm_t<256, 2>(g, st);
  }
  mfpa::set_error("conv gemm: unsupported tile BN=%d MT=%d", g.bn, g.mt);
  return MFPA_EINVAL;
}

int pick_stages(int bn, int mt) {
  const int stage = mt * kATile + bn * 128;
  int s = (int)((208 * 1024) / stage);
  if (s > 8) s = 8;
  return s < 1 ? 1 : s;
}

int pick_bn(int cout) { return cout >= 256 ? 256 : cout >= 128 ? 128 : 64; }

}  // namespace

// ====================================================================================== the network
struct ConvBN {
  bf16* w = nullptr;       // [cout][9*cin]
  float* scale = nullptr;  // [cout]
  float* shift = nullptr;
  int cin = 0, cout = 0;
};
struct UpConv {
  bf16* w = nullptr;  // [4*cout][cin], row = (a*2+b)*cout + co
  float* ones = nullptr;
  float* bias = nullptr;
  int cin = 0, cout = 0;
};

struct mfpa_unet {
  int device = 0;
  // parameters
  float* in_w = nullptr;  // [9][64] of inc.double_conv.0
  float *in_scale = nullptr, *in_shift = nullptr;
  ConvBN conv[17];  // every 3x3 conv + BN after the first: inc.3, down1.0/.3 ... down4, up1.conv.0/.3 ... up4
  UpConv up[4];
  float* out_w = nullptr;  // [64]
  float out_b = 0.f;
  bool loaded = false;
  std::vector<void*> owned;
  // activation arena for one chunk of images at one geometry
  int geo_n = 0, geo_h = 0, geo_w = 0;
  int H[5] = {0}, W[5] = {0};
  bf16* cat[4] = {nullptr};   // level l: [n][H_l][W_l][2*C_l]: skip in channels [0, C_l), upsampled in [C_l, 2 C_l)
  bf16* mid[5] = {nullptr};   // DoubleConv middle activations, level l: [n][H_l][W_l][C_l]
  bf16* pool[5] = {nullptr};  // pooled input of level l (l >= 1): [n][H_l][W_l][C_{l-1}]
  bf16* x5 = nullptr;         // [n][H_4][W_4][1024]
  bf16* dec[4] = {nullptr};   // decoder outputs u1..u3 at levels 3..1 (dec[l]), C_l channels
  std::vector<void*> arena;
  std::vector<GemmLaunch> plan;  // in execution order
  // images per pass: 37 = 148 / 4, and every level of the 257x251 geometry has a multiple of 4 tiles per
  // image, so each persistent launch fills a whole number of waves of the 148 SMs
  int max_chunk = 37;
};

namespace {

constexpr int kC[5] = {64, 128, 256, 512, 1024};

int dev_alloc(std::vector<void*>& keep, void** p, size_t bytes, bool zero) {
  MFPA_CUDA(cudaMalloc(p, bytes));
  keep.push_back(*p);
  if (zero) MFPA_CUDA(cudaMemset(*p, 0, bytes));
  return MFPA_OK;
}

int upload_f32(mfpa_unet* u, float** dst, const float* src, size_t n) {
  int rc = dev_alloc(u->owned, (void**)dst, n * sizeof(float), false);
  if (rc) return rc;
  MFPA_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
  return MFPA_OK;
}
int upload_bf16(mfpa_unet* u, bf16** dst, const std::vector<float>& src) {
  std::vector<bf16> h(src.size());
  for (size_t i = 0; i < src.size(); ++i) h[i] = __float2bfloat16(src[i]);
  int rc = dev_alloc(u->owned, (void**)dst, h.size() * sizeof(bf16), false);
  if (rc) return rc;
  MFPA_CUDA(cudaMemcpy(*dst, h.data(), h.size() * sizeof(bf16), cudaMemcpyHostToDevice));
  return MFPA_OK;
}

// BatchNorm2d in eval mode folded to y = x * scale + shift (eps 1e-5, torch default)
void fold_bn(const float* g, const float* b, const float* mean, const float* var, int c, std::vector<float>& scale,
             std::vector<float>& shift) {
  scale.resize(c);
  shift.resize(c);
  for (int i = 0; i < c; ++i) {
    const float s = g[i] / sqrtf(var[i] + 1e-5f);
    scale[i] = s;
    shift[i] = b[i] - mean[i] * s;
  }
}

void free_list(std::vector<void*>& v) {
  for (void* p : v) cudaFree(p);
  v.clear();
}

// 3x3 conv as a GEMM launch: in [n][h][w][cin] -> out channel slice
int plan_conv(mfpa_unet* u, const ConvBN& c, const bf16* in, int n, int h, int w, bf16* out, int ldc, int coff, int mode) {
  GemmLaunch g{};
  g.bn = pick_bn(c.cout);
  g.mt = h >= 16 ? 2 : 1;
  if (mode == kEpiOutc) g.bn = 64;
  g.n_total = c.cout;
  g.N = n;
  // Kernel choice (measured on B200, scratch/tune_conv.py): the 64-output-channel layers are bound by
  // L2 -> SM operand traffic in the per-tap kernel, so they use the halo-reuse kernel with the weights
  // resident in shared memory; wider layers are weight-traffic bound either way and keep
  // the per-tap kernel, with 128-pixel tiles where that evens out the last wave.
  if (c.cout <= 64 && w >= 30) {
    g.mt = 2;
    g.wres = c.cout == 64;
    g.a_stages = 2;
    g.b_stages = 4;
    if (g.wres && c.cin > 64) {
      g.wh = 18;
    } else {
      int k = 1;
      while ((w + k - 1) / k + 2 > 130) ++k;
      g.wh = (w + k - 1) / k + 2;
    }
    halo_geometry(g.wh, g.mt, &g.th, &g.a_bytes);
    if (halo_smem_bytes(g, c.cin, c.cout) > 227 * 1024) g.wh = 0;  // does not fit: per-tap kernel
  } else if (c.cout >= 256 && c.cin <= 128 && h * w >= 2048) {
    g.mt = 1;
  }
  int rc = g.wh > 0 ? make_act_map(&g.tmA, in, n, h, w, c.cin, g.th + 2, g.wh) : make_act_map(&g.tmA, in, n, h, w, c.cin, 8 * g.mt);
  if (rc) return rc;
  rc = make_wgt_map(&g.tmB, c.w, c.cout, 9 * c.cin, g.bn);
  if (rc) return rc;
  ConvArgs& a = g.args;
  a.H = h;
  a.W = w;
  a.cin = c.cin;
  a.taps = 9;
  a.stages = pick_stages(g.bn, g.mt);
  a.mode = mode;
  a.relu = 1;
  a.scale = c.scale;
  a.shift = c.shift;
  a.out = out;
  a.out_h = h;
  a.out_w = w;
  a.ldc = ldc;
  a.coff = coff;
  a.cout = c.cout;
  u->plan.push_back(g);
  return MFPA_OK;
}

// ConvTranspose2d(cin, cout, 2, stride 2) as a GEMM with N = 4*cout and a 2x2 scatter into the skip buffer
int plan_up(mfpa_unet* u, const UpConv& c, const bf16* in, int n, int h, int w, bf16* out, int out_h, int out_w, int ldc,
            int coff) {
  GemmLaunch g{};
  g.bn = 256;  // all four (a, b) phases of 64 channels, or a slice of one phase, per tile
  g.mt = 1;    // 2 x 256 TMEM columns: the (store-heavy) epilogue of one tile overlaps the MMAs of the next
  g.n_total = 4 * c.cout;
  g.N = n;
  int rc = make_act_map(&g.tmA, in, n, h, w, c.cin, 8 * g.mt);
  if (rc) return rc;
  rc = make_wgt_map(&g.tmB, c.w, 4 * c.cout, c.cin, g.bn);
  if (rc) return rc;
  ConvArgs& a = g.args;
  a.H = h;
  a.W = w;
  a.cin = c.cin;
  a.taps = 1;
  a.stages = pick_stages(g.bn, g.mt);
  a.mode = kEpiUpscatter;
  a.relu = 0;
  a.scale = c.ones;
  a.shift = c.bias;
  a.out = out;
  a.out_h = out_h;
  a.out_w = out_w;
  a.ldc = ldc;
  a.coff = coff;
  a.cout = c.cout;
  u->plan.push_back(g);
  return MFPA_OK;
}

int build_geometry(mfpa_unet* u, int n, int h, int w) {
  if (u->geo_n == n && u->geo_h == h && u->geo_w == w) return MFPA_OK;
  free_list(u->arena);
  u->plan.clear();
  u->geo_n = 0;
  MFPA_REQUIRE(h >= 16 && w >= 16, "UNet input %dx%d is too small for four 2x2 poolings", h, w);
  u->H[0] = h;
  u->W[0] = w;
  for (int l = 1; l < 5; ++l) {
    u->H[l] = u->H[l - 1] / 2;
    u->W[l] = u->W[l - 1] / 2;
  }
  int rc;
  auto alloc = [&](bf16** p, int l, int c, bool zero) {
    return dev_alloc(u->arena, (void**)p, (size_t)n * u->H[l] * u->W[l] * c * sizeof(bf16), zero);
  };
  for (int l = 0; l < 4; ++l) {
    // zeroed once: the F.pad border of the upsampled half (training/unet.py:59-62) is never written
    if ((rc = alloc(&u->cat[l], l, 2 * kC[l], true))) return rc;
    if (l >= 1 && (rc = alloc(&u->dec[l], l, kC[l], false))) return rc;  // level 0 ends in the fused 1x1 output conv
  }
  for (int l = 0; l < 5; ++l) {
    if ((rc = alloc(&u->mid[l], l, kC[l], false))) return rc;
    if (l >= 1 && (rc = alloc(&u->pool[l], l, kC[l - 1], false))) return rc;
  }
  if ((rc = alloc(&u->x5, 4, kC[4], false))) return rc;

  // plan (GEMM launches only; conv_in and the pools are issued by forward())
  // encoder: conv index 0 = inc.3; 1,2 = down1; 3,4 = down2; 5,6 = down3; 7,8 = down4
  if ((rc = plan_conv(u, u->conv[0], u->mid[0], n, u->H[0], u->W[0], u->cat[0], 2 * kC[0], 0, kEpiStore))) return rc;
  for (int l = 1; l < 5; ++l) {
    if ((rc = plan_conv(u, u->conv[2 * l - 1], u->pool[l], n, u->H[l], u->W[l], u->mid[l], kC[l], 0, kEpiStore))) return rc;
    bf16* out = l < 4 ? u->cat[l] : u->x5;
    const int ldc = l < 4 ? 2 * kC[l] : kC[4];
    if ((rc = plan_conv(u, u->conv[2 * l], u->mid[l], n, u->H[l], u->W[l], out, ldc, 0, kEpiStore))) return rc;
  }
  // decoder: up[i] lifts level 4-i to level 3-i; conv 9+2i, 10+2i
  const bf16* cur = u->x5;
  for (int i = 0; i < 4; ++i) {
    const int l = 3 - i;  // target level
    if ((rc = plan_up(u, u->up[i], cur, n, u->H[l + 1], u->W[l + 1], u->cat[l], u->H[l], u->W[l], 2 * kC[l], kC[l]))) return rc;
    if ((rc = plan_conv(u, u->conv[9 + 2 * i], u->cat[l], n, u->H[l], u->W[l], u->mid[l], kC[l], 0, kEpiStore))) return rc;
    const int mode = l == 0 ? kEpiOutc : kEpiStore;
    if ((rc = plan_conv(u, u->conv[10 + 2 * i], u->mid[l], n, u->H[l], u->W[l], u->dec[l], kC[l], 0, mode))) return rc;
    cur = u->dec[l];
  }
  u->geo_n = n;
  u->geo_h = h;
  u->geo_w = w;
  return MFPA_OK;
}

}  // namespace

// ====================================================================================== C ABI
extern "C" {

int64_t mfpa_unet_num_params(void) {
  int64_t n = 64 * 9 + 4 * 64;
  auto conv = [&](int64_t cin, int64_t cout) { n += cout * cin * 9 + 4 * cout; };
  conv(64, 64);
  for (int l = 1; l < 5; ++l) {
    conv(kC[l - 1], kC[l]);
    conv(kC[l], kC[l]);
  }
  for (int i = 0; i < 4; ++i) {
    const int64_t cin = kC[4 - i], cout = kC[3 - i];
    n += cin * cout * 4 + cout;
    conv(cin, cout);
    conv(cout, cout);
  }
  return n + 64 + 1;
}

int mfpa_unet_create(mfpa_ctx* ctx, mfpa_unet** out) {
  MFPA_REQUIRE(ctx && out, "mfpa_unet_create: null argument");
  MFPA_CUDA(cudaSetDevice(ctx->device));
  mfpa_unet* u = new mfpa_unet();
  u->device = ctx->device;
  *out = u;
  return MFPA_OK;
}

void mfpa_unet_destroy(mfpa_unet* u) {
  if (!u) return;
  cudaSetDevice(u->device);
  free_list(u->arena);
  free_list(u->owned);
  delete u;
}

int mfpa_unet_set_max_chunk(mfpa_unet* u, int images) {
  MFPA_REQUIRE(u && images >= 1 && images <= 4096, "mfpa_unet_set_max_chunk: bad argument");
  u->max_chunk = images;
  return MFPA_OK;
}

int mfpa_unet_load(mfpa_unet* u, const float* p, int64_t n) {
  MFPA_REQUIRE(u && p, "mfpa_unet_load: null argument");
  MFPA_REQUIRE(n == mfpa_unet_num_params(), "mfpa_unet_load: expected %lld floats, got %lld",
               (long long)mfpa_unet_num_params(), (long long)n);
  MFPA_CUDA(cudaSetDevice(u->device));
  free_list(u->owned);
  free_list(u->arena);
  u->plan.clear();
  u->geo_n = 0;
  u->loaded = false;
  int rc;
  std::vector<float> scale, shift, tmp;
  // inc.double_conv.0.weight [64][1][3][3] -> [tap][64]
  tmp.resize(9 * 64);
  for (int co = 0; co < 64; ++co)
    for (int t = 0; t < 9; ++t) tmp[t * 64 + co] = p[co * 9 + t];
  if ((rc = upload_f32(u, &u->in_w, tmp.data(), tmp.size()))) return rc;
  p += 64 * 9;
  fold_bn(p, p + 64, p + 128, p + 192, 64, scale, shift);
  p += 256;
  if ((rc = upload_f32(u, &u->in_scale, scale.data(), 64))) return rc;
  if ((rc = upload_f32(u, &u->in_shift, shift.data(), 64))) return rc;

  auto load_conv = [&](ConvBN& c, int cin, int cout) -> int {
    c.cin = cin;
    c.cout = cout;
    // torch [cout][cin][3][3] -> [cout][tap][cin]
    tmp.resize((size_t)cout * 9 * cin);
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int t = 0; t < 9; ++t) tmp[((size_t)co * 9 + t) * cin + ci] = p[((size_t)co * cin + ci) * 9 + t];
    int r = upload_bf16(u, &c.w, tmp);
    if (r) return r;
    p += (size_t)cout * cin * 9;
    fold_bn(p, p + cout, p + 2 * cout, p + 3 * cout, cout, scale, shift);
    p += 4 * cout;
    if ((r = upload_f32(u, &c.scale, scale.data(), cout))) return r;
    return upload_f32(u, &c.shift, shift.data(), cout);
  };
  if ((rc = load_conv(u->conv[0], 64, 64))) return rc;
  for (int l = 1; l < 5; ++l) {
    if ((rc = load_conv(u->conv[2 * l - 1], kC[l - 1], kC[l]))) return rc;
    if ((rc = load_conv(u->conv[2 * l], kC[l], kC[l]))) return rc;
  }
  for (int i = 0; i < 4; ++i) {
    const int cin = kC[4 - i], cout = kC[3 - i];
    UpConv& c = u->up[i];
    c.cin = cin;
    c.cout = cout;
    // torch ConvTranspose2d weight [cin][cout][2][2] -> [(a*2+b)*cout + co][cin]
    tmp.resize((size_t)4 * cout * cin);
    for (int ci = 0; ci < cin; ++ci)
      for (int co = 0; co < cout; ++co)
        for (int ab = 0; ab < 4; ++ab) tmp[((size_t)ab * cout + co) * cin + ci] = p[((size_t)ci * cout + co) * 4 + ab];
    if ((rc = upload_bf16(u, &c.w, tmp))) return rc;
    p += (size_t)cin * cout * 4;
    if ((rc = upload_f32(u, &c.bias, p, cout))) return rc;
    p += cout;
    std::vector<float> ones(cout, 1.0f);
    if ((rc = upload_f32(u, &c.ones, ones.data(), cout))) return rc;
    if ((rc = load_conv(u->conv[9 + 2 * i], cin, cout))) return rc;
    if ((rc = load_conv(u->conv[10 + 2 * i], cout, cout))) return rc;
  }
  if ((rc = upload_f32(u, &u->out_w, p, 64))) return rc;
  u->out_b = p[64];
  u->loaded = true;
  return MFPA_OK;
}

int mfpa_unet_forward(mfpa_ctx* ctx, mfpa_unet* u, const float* in_dev, int64_t in_n, int64_t in_h, int64_t in_w,
                      const float* div_dev, int B, int H, int W, float* out_dev, int64_t out_n, int64_t out_h,
                      int64_t out_w, void* stream) {
  MFPA_REQUIRE(ctx && u && in_dev && out_dev, "mfpa_unet_forward: null argument");
  MFPA_REQUIRE(u->loaded, "mfpa_unet_forward: no weights loaded (mfpa_unet_load)");
  MFPA_REQUIRE(B >= 0, "mfpa_unet_forward: negative batch");
  if (B == 0) return MFPA_OK;
  MFPA_CUDA(cudaSetDevice(u->device));
  cudaStream_t st = (cudaStream_t)stream;
  const int chunk = B < u->max_chunk ? B : u->max_chunk;
  if (u->geo_n != chunk || u->geo_h != H || u->geo_w != W) {
    MFPA_CUDA(cudaStreamSynchronize(st));  // the arena may still be in use
    int rc = build_geometry(u, chunk, H, W);
    if (rc) return rc;
  }
  for (int b0 = 0; b0 < B; b0 += chunk) {
    const int n = (B - b0) < chunk ? (B - b0) : chunk;
    {
      dim3 grid((W + kInTw - 1) / kInTw, (H + kInTh - 1) / kInTh, n);
      conv_in_kernel<<<grid, 256, 0, st>>>(in_dev + b0 * in_n, in_n, in_h, in_w, div_dev ? div_dev + b0 : nullptr, H, W,
                                           u->in_w, u->in_scale, u->in_shift, u->mid[0]);
      MFPA_CUDA(cudaGetLastError());
    }
    size_t k = 0;
    auto run = [&](bool last) -> int {
      GemmLaunch g = u->plan[k++];
      g.N = n;
      if (last) {
        g.args.w_out = u->out_w;
        g.args.b_out = u->out_b;
        g.args.out_f32 = out_dev + b0 * out_n;
        g.args.of_n = out_n;
        g.args.of_h = out_h;
        g.args.of_w = out_w;
      }
      return launch_gemm(g, st);
    };
    int rc;
    if ((rc = run(false))) return rc;  // inc.3 -> cat[0][:, :64]
    for (int l = 1; l < 5; ++l) {
      const bf16* src = u->cat[l - 1];
      const long long total = (long long)n * u->H[l] * u->W[l] * (kC[l - 1] / 8);
      maxpool_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, u->H[l - 1], u->W[l - 1], 2 * kC[l - 1], kC[l - 1],
                                                                      u->pool[l], total);
      MFPA_CUDA(cudaGetLastError());
      if ((rc = run(false))) return rc;
      if ((rc = run(false))) return rc;
    }
    for (int i = 0; i < 4; ++i) {
      if ((rc = run(false))) return rc;
      if ((rc = run(false))) return rc;
      if ((rc = run(i == 3))) return rc;
    }
  }
  return MFPA_OK;
}

// One 3x3 convolution + scale/shift (+ReLU) on caller-owned NHWC bf16 tensors: the unit under test
// of the implicit-GEMM kernel.  w_dev: [cout][9][cin] bf16; out: channel slice [coff, coff+cout) of
// [N][H][W][ldc].  taps = 1 runs a 1x1 convolution (w_dev [cout][cin]).
int mfpa_conv_bf16(mfpa_ctx* ctx, const void* in_dev, int N, int H, int W, int cin, const void* w_dev, int cout, int taps,
                   const float* scale_dev, const float* shift_dev, int relu, void* out_dev, int ldc, int coff, int bn,
                   int mt, int stages, int halo_wh, int wres, void* stream) {
  MFPA_REQUIRE(ctx && in_dev && w_dev && out_dev && scale_dev && shift_dev, "mfpa_conv_bf16: null argument");
  MFPA_REQUIRE(cin % 64 == 0 && cout % 64 == 0 && (taps == 9 || taps == 1), "mfpa_conv_bf16: cin/cout must be multiples of 64, taps 1 or 9");
  MFPA_REQUIRE(ldc % 8 == 0 && coff % 8 == 0, "mfpa_conv_bf16: ldc and coff must be multiples of 8");
  GemmLaunch g{};
  g.bn = bn > 0 ? bn : pick_bn(cout);
  g.mt = mt > 0 ? mt : (H >= 16 ? 2 : 1);
  MFPA_REQUIRE(cout % g.bn == 0, "mfpa_conv_bf16: cout %d is not a multiple of the N tile %d", cout, g.bn);
  g.n_total = cout;
  g.N = N;
  int rc;
  ConvArgs& a = g.args;
  if (halo_wh > 0) {
    MFPA_REQUIRE(taps == 9 && halo_wh >= 3 && halo_wh <= 256, "mfpa_conv_bf16: the halo kernel needs taps = 9 and 3 <= wh <= 256");
    MFPA_REQUIRE(!wres || (g.bn == cout && g.bn == 64), "mfpa_conv_bf16: resident weights need cout == bn == 64");
    g.wh = halo_wh;
    g.wres = wres ? 1 : 0;
    halo_geometry(g.wh, g.mt, &g.th, &g.a_bytes);
    g.a_stages = stages > 0 ? stages : 2;
    g.b_stages = 4;
    MFPA_REQUIRE(g.th + 2 <= 256, "mfpa_conv_bf16: halo box too tall");
    MFPA_REQUIRE(halo_smem_bytes(g, cin, cout) <= 227 * 1024, "mfpa_conv_bf16: halo configuration needs %zu bytes of shared memory",
                 halo_smem_bytes(g, cin, cout));
    rc = make_act_map(&g.tmA, (const bf16*)in_dev, N, H, W, cin, g.th + 2, g.wh);
  } else {
    rc = make_act_map(&g.tmA, (const bf16*)in_dev, N, H, W, cin, 8 * g.mt);
  }
  if (rc) return rc;
  rc = make_wgt_map(&g.tmB, (const bf16*)w_dev, cout, taps * cin, g.bn);
  if (rc) return rc;
  a.H = H;
  a.W = W;
  a.cin = cin;
  a.taps = taps;
  a.stages = stages > 0 ? stages : pick_stages(g.bn, g.mt);
  MFPA_REQUIRE(halo_wh > 0 || gemm_smem_bytes(g.mt, g.bn, a.stages, cout) <= 227 * 1024, "mfpa_conv_bf16: %d stages do not fit shared memory", a.stages);
  a.mode = kEpiStore;
  a.relu = relu;
  a.scale = scale_dev;
  a.shift = shift_dev;
  a.out = (bf16*)out_dev;
  a.out_h = H;
  a.out_w = W;
  a.ldc = ldc;
  a.coff = coff;
  a.cout = cout;
  return launch_gemm(g, (cudaStream_t)stream);
}

}  // extern "C"
