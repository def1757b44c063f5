// S5 — landmark-hash matching against a hash-range shard of the reference index.
// Replaces HashTable.get_hits (afp/audfprint/hash_table.py:220-246) and
// Matcher.match_hashes (afp/audfprint/audfprint_match.py:102-129, 235-349).
//
// Index layout is the reference's own (hash_table.py:53-68): table[bucket][depth] uint32,
// entry = ((id + 1) << maxtimebits) + time, counts[bucket] int32 (may exceed depth).
// A context holds the buckets [hash_lo, hash_lo + n_buckets) only.  With the whole index in one context
// mfpa_match runs counts + select + collect as ONE kernel whose per-track histogram never leaves shared
// memory (match_fused_kernel) followed by the align kernel.  For an index sharded by hash range the same
// four steps are separate kernels, so that the two reductions the algorithm needs can cross GPUs with
// plain NCCL collectives between them (musicfpaugment_b200/sharded.py):
//   counts   per-(query, track) raw hit counts from the local buckets (shared-memory histogram, dense
//            int32 or 16-bit-packed row out)            -> reduce-scatter(sum) to the query's owner rank
//   select   candidates: top min(#{raw > 5}, 100) by raw / hashesperid (:110-129) -> all-gather
//   collect  (candidate, delta-t) of every local hit of a candidate -> all-to-all to the owners
//   align    per candidate delta-t histogram, local-max modes > 5, +-window sum (:266-316), rows
//            ordered by filtered count descending (:341)  -> all-gather of the rows
#include "common.cuh"

namespace mfpa {

namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr int kCountThreads = 512;
constexpr int kMaxTracksSmem = 110000;  // packed 16-bit counters: 220 KB of shared memory
constexpr int kAlignCap = 8192;         // hits of all candidates of one query that align_kernel can sort
constexpr int kAlignSmall = 1024;       // ... in its first pass (most queries: a few hundred hits)
constexpr int kAlignPending = -3;       // nrows marker: query left for the full-capacity pass
constexpr int kDtOff = 16384;

struct IndexView {
  const uint32_t* table;
  const int32_t* counts;
  const uint32_t* hashesperid;
  int hash_lo, n_buckets, depth, maxtimebits, n_tracks, hashmask;
  uint32_t hp_min;
};

// ---- counts ------------------------------------------------------------------------------
// (the shared-memory histogram version is match_counts_sweep_kernel, next to the sweep it shares with the
// one-kernel matcher)
// Fallback for indexes with more tracks than the shared-memory histogram holds: atomics on the dense row.
__global__ void __launch_bounds__(kCountThreads) match_counts_global_kernel(const IndexView ix, const int32_t* __restrict__ hashes,
                                                                           const int32_t* __restrict__ nh, int cap,
                                                                           int32_t* __restrict__ out) {
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int32_t* o = out + (int64_t)q * ix.n_tracks;
  const int2* rows = reinterpret_cast<const int2*>(hashes) + (int64_t)q * cap;
  const int n = min(nh[q], cap);
  for (int r = warp; r < n; r += kCountThreads / 32) {
    const int h = (rows[r].y & ix.hashmask) - ix.hash_lo;
    if (h < 0 || h >= ix.n_buckets) continue;
    const int cnt = min(ix.depth, ix.counts[h]);
    const uint32_t* bucket = ix.table + (int64_t)h * ix.depth;
    for (int s = lane; s < cnt; s += 32) {
      const int id = (int)(bucket[s] >> ix.maxtimebits) - 1;
      if (id >= 0 && id < ix.n_tracks) atomicAdd(&o[id], 1);
    }
  }
}

// ---- select ------------------------------------------------------------------------------
// Order of _best_count_ids: weighted count raw/hashesperid descending; ties -> higher id first
// (what argsort(...)[::-1] gives for a stable sort; numpy's is unstable, ties are arbitrary there).
struct Cand { long long raw; long long hp; int id; };
__device__ __forceinline__ bool before(const Cand& a, const Cand& b) {  // a ranks ahead of b
  const long long l = a.raw * b.hp, r = b.raw * a.hp;
  return l > r || (l == r && a.id > b.id);
}

// Exact ranking of a short list of contender tracks (shared memory): the first `depth` by `before`, written
// as (id, raw) pairs to cand_q.  All NT threads of the block call it.
template <int NT, typename CountOf>
__device__ __forceinline__ void rank_list(const int* list, int n_list, int depth, CountOf count_of,
                                          const uint32_t* __restrict__ hpid, int32_t* __restrict__ cand_q,
                                          Cand* s_best, Cand* s_prev, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  __syncthreads();
  if (tid == 0) *s_prev = Cand{1, 0, 0x7fffffff};  // +infinity sentinel (hp = 0)
  __syncthreads();
  for (int k = 0; k < depth; ++k) {
    const Cand prev = *s_prev;
    Cand best{-1, 1, -1};
    for (int c = tid; c < n_list; c += NT) {
      const int i = list[c];
      const Cand x{(long long)count_of(i), (long long)hpid[i], i};
      const bool after_prev = prev.hp == 0 || before(prev, x);
      if (after_prev && (best.id < 0 || before(x, best))) best = x;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      Cand y;
      y.raw = __shfl_xor_sync(kFull, best.raw, o); y.hp = __shfl_xor_sync(kFull, best.hp, o); y.id = __shfl_xor_sync(kFull, best.id, o);
      if (y.id >= 0 && (best.id < 0 || before(y, best))) best = y;
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (tid == 0) {
      Cand b = s_best[0];
      for (int w = 1; w < NT / 32; ++w) if (s_best[w].id >= 0 && (b.id < 0 || before(s_best[w], b))) b = s_best[w];
      *s_prev = b;
      cand_q[2 * k] = b.id;
      cand_q[2 * k + 1] = (int)b.raw;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(512) match_select_kernel(const int32_t* __restrict__ counts, const uint32_t* __restrict__ hpid,
                                                           int n_tracks, int threshcount, int search_depth,
                                                           int32_t* __restrict__ cand, int32_t* __restrict__ ncand, int packed,
                                                           uint32_t hp_min) {
  constexpr int kListCap = 4096;
  __shared__ int s_int[16];
  __shared__ float s_flt[16];
  __shared__ Cand s_best[16];
  __shared__ Cand s_prev;
  __shared__ int s_list[kListCap];
  __shared__ int s_n;
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int32_t* c = counts + (int64_t)q * (packed ? (n_tracks + 1) >> 1 : n_tracks);
  auto count_of = [&](int i) -> int {
    return packed ? (int)((reinterpret_cast<const unsigned*>(c)[i >> 1] >> ((i & 1) * 16)) & 0xffffu) : c[i];
  };
  // the tracks wanted all have quotient raw / hashesperid >= m = min over the tracks with raw > threshcount
  // (see match_fused_kernel): list those contenders in one more pass over the row and rank the short list
  int gt = 0;
  float fmin = INFINITY;
  for (int i = tid; i < n_tracks; i += 512) {
    const int raw = count_of(i);
    if (raw > threshcount) { ++gt; fmin = fminf(fmin, __fdividef((float)raw, (float)__ldg(hpid + i))); }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    gt += __shfl_xor_sync(kFull, gt, o);
    fmin = fminf(fmin, __shfl_xor_sync(kFull, fmin, o));
  }
  if (lane == 0) { s_int[warp] = gt; s_flt[warp] = fmin; }
  if (tid == 0) s_n = 0;
  __syncthreads();
  gt = 0;
  for (int w = 0; w < 16; ++w) { gt += s_int[w]; fmin = fminf(fmin, s_flt[w]); }
  const int depth = min(gt, search_depth);
  if (tid == 0) ncand[q] = depth;
  if (depth == 0) return;
  const float cut = fmin * 0.99999f;
  const int raw_min = max(1, (int)floorf(cut * (float)hp_min * 0.99999f));   // raw < raw_min => quotient < cut
  for (int i = tid; i < n_tracks; i += 512) {
    const int raw = count_of(i);
    if (raw < raw_min) continue;
    if (__fdividef((float)raw, (float)__ldg(hpid + i)) < cut) continue;
    const int slot = atomicAdd(&s_n, 1);
    if (slot < kListCap) s_list[slot] = i;
  }
  __syncthreads();
  if (s_n <= kListCap) {
    rank_list<512>(s_list, s_n, depth, count_of, hpid, cand + (int64_t)q * search_depth * 2, s_best, &s_prev, tid);
    return;
  }
  // more contenders than the list holds: rank by repeated scans of the whole row
  if (tid == 0) s_prev = Cand{1, 0, 0x7fffffff};  // +infinity sentinel (hp = 0)
  __syncthreads();
  for (int k = 0; k < depth; ++k) {
    const Cand prev = s_prev;
    Cand best{-1, 1, -1};
    for (int i = tid; i < n_tracks; i += 512) {
      const int raw = count_of(i);
      if (raw <= 0) continue;
      const Cand x{raw, (long long)hpid[i], i};
      const bool after_prev = prev.hp == 0 || before(prev, x);
      if (after_prev && (best.id < 0 || before(x, best))) best = x;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      Cand y;
      y.raw = __shfl_xor_sync(kFull, best.raw, o); y.hp = __shfl_xor_sync(kFull, best.hp, o); y.id = __shfl_xor_sync(kFull, best.id, o);
      if (y.id >= 0 && (best.id < 0 || before(y, best))) best = y;
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (tid == 0) {
      Cand b = s_best[0];
      for (int w = 1; w < 16; ++w) if (s_best[w].id >= 0 && (b.id < 0 || before(s_best[w], b))) b = s_best[w];
      s_prev = b;
      cand[((int64_t)q * search_depth + k) * 2] = b.id;
      cand[((int64_t)q * search_depth + k) * 2 + 1] = (int)b.raw;
    }
    __syncthreads();
  }
}

// ---- counts + select + collect in one block per query (single shard: no exchange between the steps) ----
// The packed 16-bit histogram never leaves shared memory: the block counts the query's hits, ranks the
// candidates straight from the histogram, then re-uses the same words as a "candidate index + 1" map for
// the second sweep over the (L2-hot) bucket rows that emits the candidates' (index, delta-t) hits.
// The sweeps are latency-bound gathers, so each warp keeps four bucket rows (16 loads per lane) in flight;
// the per-row (bucket, count, query time) triples are staged through a small shared-memory cache first.
constexpr int kFusedRows = 768;
constexpr int kFusedThreads = 1024;
struct RowCache { int hb[kFusedRows]; int cnt[kFusedRows]; int t[kFusedRows]; int n; };   // n: rows kept (non-empty buckets of this shard)

enum { kSweepCount = 0, kSweepCollect = 1, kSweepEmit = 2 };
constexpr int kHitDtBits = 15;          // hit word of the sparse exchange: (track << 15) | (t_ref - t_q + 16384)
constexpr int kHitMaxTracks = 1 << 17;  // ... which leaves 17 bits for the track

// MODE kSweepCount: raw counts into the packed shared-memory histogram; kSweepCollect: (candidate, delta-t) hits of
// the tracks marked in it; kSweepEmit: every hit of this shard as a (track, delta-t) word (sparse exchange).
// A query time outside [0, 2^14) cannot be packed next to the table's 14-bit reference times: *s_bad is raised.
// The packed histogram occupies (words + 3) & ~3 shared-memory words; cleared (padding included, the vector scans of
// fused_select read it) with 16-byte stores.
__device__ __forceinline__ void zero_hist(unsigned* hist, int words, int tid) {
  uint4* h4 = reinterpret_cast<uint4*>(hist);
  for (int i = tid; i < ((words + 3) >> 2); i += kFusedThreads) h4[i] = make_uint4(0u, 0u, 0u, 0u);
}

template <int MODE>
__device__ __forceinline__ void fused_sweep(const IndexView& ix, const int2* __restrict__ rows, int n, RowCache* rc,
                                            unsigned* hist, uint32_t* __restrict__ out, int list_cap, int* s_n, int tid,
                                            int* s_bad = nullptr, unsigned* s_hits = nullptr) {
  constexpr bool COLLECT = MODE == kSweepCollect;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t tmask = (1u << ix.maxtimebits) - 1u;
  const unsigned short* mark = reinterpret_cast<const unsigned short*>(hist);
  for (int c0 = 0; c0 < n; c0 += kFusedRows) {
    const int n_in = min(kFusedRows, n - c0);
    if (tid == 0) rc->n = 0;
    __syncthreads();
    // keep the rows whose bucket lies in this shard and is not empty (1 in `world` on a hash-range shard): the
    // warps then sweep only those, spread evenly (warp w takes kept rows w, w + 32, w + 64, w + 96 of a round)
    for (int i = tid; i < n_in; i += kFusedThreads) {
      const int2 row = rows[c0 + i];
      const int hb = (row.y & ix.hashmask) - ix.hash_lo;
      const int cnt = (hb >= 0 && hb < ix.n_buckets) ? min(ix.depth, ix.counts[hb]) : 0;
      if (s_bad && (row.x < 0 || row.x >= kDtOff)) *s_bad = 1;
      if (cnt > 0) {
        const int slot = atomicAdd(&rc->n, 1);
        rc->hb[slot] = hb;
        rc->cnt[slot] = cnt;
        rc->t[slot] = row.x;
        if (s_hits) atomicAdd(s_hits, (unsigned)cnt);
      }
    }
    __syncthreads();
    const int nc = rc->n;
    for (int s0 = 0; s0 < ix.depth; s0 += 128) {   // one trip for depth <= 128 (the reference's is 100)
      for (int r = warp; r < nc; r += 128) {   // rounds of 4 x 32 rows
        uint32_t v[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int rr = r + 32 * a;
          const int cnt = rr < nc ? rc->cnt[rr] : 0;
          const uint32_t* bucket = ix.table + (int64_t)(rr < nc ? rc->hb[rr] : 0) * ix.depth;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int s = s0 + lane + 32 * i;
            v[a][i] = s < cnt ? __ldg(bucket + s) : 0u;   // 0 = empty slot (stored ids are id + 1 >= 1)
          }
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const unsigned id = (v[a][i] >> ix.maxtimebits) - 1u;   // empty slot -> 0xffffffff
            if (MODE == kSweepEmit) {
              // one shared-memory atomic per warp and 32 slots; the warp's words land next to each other, so the
              // stores (to the local list or straight into the owner rank's memory over NVLink) are coalesced
              const bool ok = id < (unsigned)ix.n_tracks;
              const unsigned m = __ballot_sync(kFull, ok);
              if (m == 0u) continue;   // warp-uniform
              int base = 0;
              if (lane == 0) base = atomicAdd(s_n, __popc(m));
              base = __shfl_sync(kFull, base, 0);
              const int pos = base + __popc(m & ((1u << lane) - 1u));
              if (ok && pos < list_cap) {
                const int dt = (int)(v[a][i] & tmask) - rc->t[r + 32 * a];
                out[pos] = (id << kHitDtBits) | (uint32_t)(dt + kDtOff);
              }
              continue;
            }
            if (id >= (unsigned)ix.n_tracks) continue;
            if (!COLLECT) {
              atomicAdd(&hist[id >> 1], 1u << ((id & 1u) << 4));
            } else {
              const unsigned k = mark[id];
              if (k) {
                const int dt = (int)(v[a][i] & tmask) - rc->t[r + 32 * a];
                const int pos = atomicAdd(s_n, 1);
                if (pos < list_cap) out[pos] = ((k - 1u) << 16) | (uint32_t)(dt + kDtOff);
              }
            }
          }
        }
      }
    }
  }
  __syncthreads();
}

// counts step of the sharded path: the same first sweep, histogram written out as a dense row (int32, or the
// 16-bit pairs as they are under MFPA_OPT_MATCH_PACKED: half the bytes to reduce-scatter).  One block per
// query; every query hash that falls in this shard contributes the first min(depth, counts[h]) entries of
// its bucket (hash_table.py:235-236).
__global__ void __launch_bounds__(kFusedThreads, 1)
match_counts_sweep_kernel(const IndexView ix, const int32_t* __restrict__ hashes, const int32_t* __restrict__ nh, int cap,
                          int32_t* __restrict__ out, int packed) {
  extern __shared__ __align__(16) unsigned fused_smem[];
  const int words = (ix.n_tracks + 1) >> 1;
  unsigned* hist = fused_smem;
  RowCache* rc = reinterpret_cast<RowCache*>(fused_smem + ((words + 3) & ~3));
  const int q = blockIdx.x, tid = threadIdx.x;
  const int2* rows = reinterpret_cast<const int2*>(hashes) + (int64_t)q * cap;
  const int n = min(nh[q], cap);
  zero_hist(hist, words, tid);
  fused_sweep<kSweepCount>(ix, rows, n, rc, hist, nullptr, 0, nullptr, tid);
  if (packed) {
    unsigned* o = reinterpret_cast<unsigned*>(out) + (int64_t)q * words;
    for (int i = tid; i < words; i += kFusedThreads) o[i] = hist[i];
    return;
  }
  int32_t* o = out + (int64_t)q * ix.n_tracks;
  for (int i = tid; i < ix.n_tracks; i += kFusedThreads) o[i] = (int)((hist[i >> 1] >> ((i & 1) * 16)) & 0xffffu);
}

// collect step of the sharded path with the same machinery: the shared-memory words hold the track ->
// (candidate index + 1) map, the second sweep emits the candidates' (index, delta-t) hits of this shard.
__global__ void __launch_bounds__(kFusedThreads, 1)
match_collect_sweep_kernel(const IndexView ix, const int32_t* __restrict__ hashes, const int32_t* __restrict__ nh, int cap,
                           const int32_t* __restrict__ cand, const int32_t* __restrict__ ncand, int search_depth,
                           uint32_t* __restrict__ list, int list_cap, int32_t* __restrict__ nlist) {
  extern __shared__ __align__(16) unsigned fused_smem[];
  __shared__ int s_n;
  const int words = (ix.n_tracks + 1) >> 1;
  unsigned* hist = fused_smem;
  RowCache* rc = reinterpret_cast<RowCache*>(fused_smem + ((words + 3) & ~3));
  const int q = blockIdx.x, tid = threadIdx.x;
  const int nc = min(ncand[q], min(search_depth, 128));
  if (nc <= 0) {   // block-uniform
    if (tid == 0) nlist[q] = 0;
    return;
  }
  zero_hist(hist, words, tid);
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (int k = tid; k < nc; k += kFusedThreads) {
    const int id = cand[((int64_t)q * search_depth + k) * 2];
    if (id >= 0 && id < ix.n_tracks) reinterpret_cast<unsigned short*>(hist)[id] = (unsigned short)(k + 1);
  }
  const int2* rows = reinterpret_cast<const int2*>(hashes) + (int64_t)q * cap;
  fused_sweep<kSweepCollect>(ix, rows, min(nh[q], cap), rc, hist, list + (int64_t)q * list_cap, list_cap, &s_n, tid);
  if (tid == 0) nlist[q] = s_n;
}

// select (_best_count_ids, audfprint_match.py:102-129) on the packed shared-memory histogram of one query; all
// kFusedThreads threads of the block call it.  Returns depth = min(#{raw > threshcount}, search_depth), writes the
// (id, raw) pairs of the ranked candidates to cand_q and depth to *ncand_q.  `scratch` = kContCap ints.
// depth tracks are wanted, ranked by raw / hashesperid over ALL tracks.  Every track with raw > threshcount has a
// quotient >= m = min of theirs, and there are >= depth of them, so the wanted tracks all have quotient >= m: one
// cheap pass finds the count and m, a second lists the contenders {quotient >= m} (tracks whose raw count is below
// m * min(hashesperid) are skipped without touching hashesperid), and the exact ranking runs over that short list.
// float32 quotients carry a relative error < 1e-6; the list is cut at m * (1 - 1e-5) and ranked with exact integer
// cross products.
constexpr int kFusedWarps = kFusedThreads / 32;
constexpr int kContCap = (int)(sizeof(RowCache) / sizeof(int));
struct SelectSmem { int s_int[kFusedWarps]; float s_flt[kFusedWarps]; Cand s_best[kFusedWarps]; Cand s_prev; int s_n; };

__device__ __forceinline__ int fused_select(const IndexView& ix, const unsigned* hist, int words, int threshcount, int search_depth,
                                            int* contenders, SelectSmem* sm, int32_t* __restrict__ cand_q,
                                            int32_t* __restrict__ ncand_q, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  auto count_of = [&](int i) -> int { return (int)((hist[i >> 1] >> ((i & 1) * 16)) & 0xffffu); };
  int gt = 0;
  float fmin = INFINITY;
  // uniform trip count so the (rare) tracks above the count threshold can be handled behind a warp vote: the
  // common iteration is one 16-byte shared-memory load (eight counters), three packed maxima and a compare
  const uint4* hist4 = reinterpret_cast<const uint4*>(hist);
  const int words4 = (words + 3) >> 2;
  auto max8 = [](const uint4& h) -> int {
    const unsigned m = __vmaxu2(__vmaxu2(h.x, h.y), __vmaxu2(h.z, h.w));
    return (int)max(m & 0xffffu, m >> 16);
  };
  for (int w0 = 0; w0 < words4; w0 += kFusedThreads) {
    const int w4 = w0 + tid;
    const uint4 h = w4 < words4 ? hist4[w4] : make_uint4(0u, 0u, 0u, 0u);
    if (__any_sync(kFull, max8(h) > threshcount)) {
      const unsigned hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r0 = (int)(hw[j] & 0xffffu), r1 = (int)(hw[j] >> 16), w = 4 * w4 + j;
        if (r0 > threshcount) { ++gt; fmin = fminf(fmin, __fdividef((float)r0, (float)__ldg(ix.hashesperid + 2 * w))); }
        if (r1 > threshcount) { ++gt; fmin = fminf(fmin, __fdividef((float)r1, (float)__ldg(ix.hashesperid + 2 * w + 1))); }
      }
    }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    gt += __shfl_xor_sync(kFull, gt, o);
    fmin = fminf(fmin, __shfl_xor_sync(kFull, fmin, o));
  }
  if (lane == 0) { sm->s_int[warp] = gt; sm->s_flt[warp] = fmin; }
  if (tid == 0) sm->s_n = 0;
  __syncthreads();
  gt = 0;
  for (int w = 0; w < kFusedWarps; ++w) { gt += sm->s_int[w]; fmin = fminf(fmin, sm->s_flt[w]); }
  const int depth = min(gt, search_depth);
  const float cut = fmin * 0.99999f;
  const int raw_min = max(1, (int)floorf(cut * (float)ix.hp_min * 0.99999f));   // raw < raw_min => quotient < cut
  if (tid == 0) *ncand_q = depth;
  bool listed = depth > 0;
  if (listed) {
    for (int w0 = 0; w0 < words4; w0 += kFusedThreads) {
      const int w4 = w0 + tid;
      const uint4 h = w4 < words4 ? hist4[w4] : make_uint4(0u, 0u, 0u, 0u);
      if (__any_sync(kFull, max8(h) >= raw_min)) {
        const unsigned hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int raw = (int)((hw[j] >> (16 * e)) & 0xffffu), i = 2 * (4 * w4 + j) + e;
            if (raw < raw_min) continue;
            if (__fdividef((float)raw, (float)__ldg(ix.hashesperid + i)) < cut) continue;
            const int slot = atomicAdd(&sm->s_n, 1);
            if (slot < kContCap) contenders[slot] = i;
          }
        }
      }
    }
    __syncthreads();
    listed = sm->s_n <= kContCap;
  }
  if (depth > 0 && listed) {
    const int nc_list = sm->s_n;
    rank_list<kFusedThreads>(contenders, nc_list, depth, count_of, ix.hashesperid, cand_q, sm->s_best, &sm->s_prev, tid);
    __syncthreads();
  } else if (depth > 0) {
    // more contenders than the list holds: rank by repeated scans of the whole histogram
    __syncthreads();
    if (tid == 0) sm->s_prev = Cand{1, 0, 0x7fffffff};
    __syncthreads();
    for (int k = 0; k < depth; ++k) {
      const Cand prev = sm->s_prev;
      Cand best{-1, 1, -1};
      const float f_hi = prev.hp == 0 ? INFINITY : __fdividef((float)prev.raw, (float)prev.hp) * 1.00001f;
      float f_lo = cut;
      for (int w = tid; w < words; w += kFusedThreads) {
        const unsigned h2 = hist[w];
        if (!h2) continue;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int raw = (int)((h2 >> (16 * e)) & 0xffffu), i = 2 * w + e;
          if (raw < raw_min || i >= ix.n_tracks) continue;
          const unsigned hp = __ldg(ix.hashesperid + i);
          const float fx = __fdividef((float)raw, (float)hp);
          if (fx > f_hi || fx < f_lo) continue;
          const Cand x{raw, (long long)hp, i};
          const bool after_prev = prev.hp == 0 || before(prev, x);
          if (after_prev && (best.id < 0 || before(x, best))) { best = x; f_lo = fmaxf(cut, fx * 0.99999f); }
        }
      }
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        Cand y;
        y.raw = __shfl_xor_sync(kFull, best.raw, o); y.hp = __shfl_xor_sync(kFull, best.hp, o); y.id = __shfl_xor_sync(kFull, best.id, o);
        if (y.id >= 0 && (best.id < 0 || before(y, best))) best = y;
      }
      if (lane == 0) sm->s_best[warp] = best;
      __syncthreads();
      if (tid == 0) {
        Cand b = sm->s_best[0];
        for (int w = 1; w < kFusedWarps; ++w) if (sm->s_best[w].id >= 0 && (b.id < 0 || before(sm->s_best[w], b))) b = sm->s_best[w];
        sm->s_prev = b;
        cand_q[2 * k] = b.id;
        cand_q[2 * k + 1] = (int)b.raw;
      }
      __syncthreads();
    }
  }
  return depth;
}

constexpr int kBadQueryTime = -5;   // nrows marker: a query time outside [0, 2^14)
constexpr int kWideCounts = -6;     // nrows marker: a track collected >= 65536 hits, the packed 16-bit counters overflowed

// The packed histogram adds 1 << 16 or 1 to a 32-bit word per hit; a track with >= 65536 hits carries into its
// neighbour.  Impossible below 65536 hits in total; above, the halves must still add up to the hit count (a carry
// changes the sum by 65535, a wrap of the upper half by 65536).  All threads call it; returns true when consistent.
__device__ __forceinline__ bool hist_consistent(const unsigned* hist, int words, unsigned total_hits, unsigned* s_sum, int tid) {
  if (total_hits < 65536u) return true;   // block-uniform
  if (tid == 0) *s_sum = 0u;
  __syncthreads();
  unsigned acc = 0;
  for (int w = tid; w < words; w += kFusedThreads) { const unsigned h2 = hist[w]; acc += (h2 & 0xffffu) + (h2 >> 16); }
#pragma unroll
  for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  if ((tid & 31) == 0) atomicAdd(s_sum, acc);
  __syncthreads();
  return *s_sum == total_hits;
}

__global__ void __launch_bounds__(kFusedThreads, 1)
match_fused_kernel(const IndexView ix, const int32_t* __restrict__ hashes, const int32_t* __restrict__ nh, int cap,
                   int threshcount, int search_depth, int32_t* __restrict__ cand, int32_t* __restrict__ ncand,
                   uint32_t* __restrict__ list, int list_cap, int32_t* __restrict__ nlist) {
  extern __shared__ __align__(16) unsigned fused_smem[];
  const int words = (ix.n_tracks + 1) >> 1;
  unsigned* hist = fused_smem;
  RowCache* rc = reinterpret_cast<RowCache*>(fused_smem + ((words + 3) & ~3));
  __shared__ SelectSmem sel;
  __shared__ int s_n, s_bad;
  __shared__ unsigned s_hits, s_sum;
  const int q = blockIdx.x, tid = threadIdx.x;
  const int2* rows = reinterpret_cast<const int2*>(hashes) + (int64_t)q * cap;
  const int n = min(nh[q], cap);
  zero_hist(hist, words, tid);
  if (tid == 0) { s_n = 0; s_bad = 0; s_hits = 0u; }
  fused_sweep<kSweepCount>(ix, rows, n, rc, hist, nullptr, 0, nullptr, tid, &s_bad, &s_hits);
  if (s_bad) {   // block-uniform after the sweep's closing barrier
    if (tid == 0) { ncand[q] = 0; nlist[q] = kBadQueryTime; }
    return;
  }
  if (!hist_consistent(hist, words, s_hits, &s_sum, tid)) {
    if (tid == 0) { ncand[q] = 0; nlist[q] = kWideCounts; }
    return;
  }
  // the row cache is idle during the select: it holds the contender list
  const int depth = fused_select(ix, hist, words, threshcount, search_depth, reinterpret_cast<int*>(rc), &sel,
                                 cand + (int64_t)q * search_depth * 2, ncand + q, tid);
  if (depth == 0) {
    if (tid == 0) nlist[q] = 0;
    return;
  }
  // the histogram becomes the candidate map: 16-bit slot of track id = candidate index + 1
  zero_hist(hist, words, tid);
  __syncthreads();
  const int nc = min(depth, 128);   // align_kernel handles at most 128 candidates (check_match bounds search_depth)
  for (int k = tid; k < nc; k += kFusedThreads)
    reinterpret_cast<unsigned short*>(hist)[cand[((int64_t)q * search_depth + k) * 2]] = (unsigned short)(k + 1);
  fused_sweep<kSweepCollect>(ix, rows, n, rc, hist, list + (int64_t)q * list_cap, list_cap, &s_n, tid);
  if (tid == 0) nlist[q] = s_n;
}

// ---- sparse exchange for an index sharded by hash range ("all-to-all the compact hit lists") -------
// emit: every rank sweeps ITS buckets once per query and lists the hits as (track, delta-t) words - a few thousand
//       per query and shard, where the dense per-track histogram is 200 KB; one all-to-all keyed by the query's
//       owner moves them (musicfpaugment_b200/sharded.py);
// owner: the owner rebuilds the shared-memory histogram from the words of all shards, selects the candidates and
//       collects their (candidate, delta-t) hits from the same words - no second exchange, no histogram in HBM.
__global__ void __launch_bounds__(kFusedThreads, 2)
match_emit_kernel(const IndexView ix, const int32_t* __restrict__ hashes, const int32_t* __restrict__ nh, int cap,
                  uint32_t* __restrict__ words_out, int words_cap, int32_t* __restrict__ nwords) {
  __shared__ RowCache rc;
  __shared__ int s_n, s_bad;
  const int q = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) { s_n = 0; s_bad = 0; }
  const int2* rows = reinterpret_cast<const int2*>(hashes) + (int64_t)q * cap;
  fused_sweep<kSweepEmit>(ix, rows, min(nh[q], cap), &rc, nullptr, words_out + (int64_t)q * words_cap, words_cap, &s_n, tid, &s_bad);
  if (tid == 0) nwords[q] = s_bad ? kBadQueryTime : s_n;
}

// The same sweep with the exchange fused in: query q of the sub-batch belongs to rank q / own, and its words go
// straight into that rank's receive buffer (peer memory mapped through CUDA IPC, NVLink stores) at the slot
// [this rank][q % own] - the layout match_owner_kernel reads.  No staging list in local HBM, no collective call.
// Consecutive blocks take queries of different owners (block b: owner b % world, its query b / world), so the stores
// to every peer - and the local ones - are spread over the whole kernel instead of one link at a time.
struct PeerDst { uint32_t* words[MFPA_MAX_PEERS]; int32_t* nwords[MFPA_MAX_PEERS]; int own, rank, world, fence; };
__global__ void __launch_bounds__(kFusedThreads, 2)
match_emit_peer_kernel(const IndexView ix, const int32_t* __restrict__ hashes, const int32_t* __restrict__ nh, int cap,
                       const PeerDst d, int words_cap) {
  __shared__ RowCache rc;
  __shared__ int s_n, s_bad;
  const int tid = threadIdx.x;
  const int owner = (int)(blockIdx.x % (unsigned)d.world), j = (int)(blockIdx.x / (unsigned)d.world);
  const int q = owner * d.own + j;
  const int64_t slot = (int64_t)d.rank * d.own + j;
  if (tid == 0) { s_n = 0; s_bad = 0; }
  const int2* rows = reinterpret_cast<const int2*>(hashes) + (int64_t)q * cap;
  fused_sweep<kSweepEmit>(ix, rows, min(nh[q], cap), &rc, nullptr, d.words[owner] + slot * words_cap, words_cap, &s_n, tid, &s_bad);
  // One system-scope fence per block, by the thread that publishes the count, after the block barrier that ends
  // the sweep: it is cumulative over the words the other threads stored before that barrier.  (A fence in every
  // thread cost as much as the sweep itself; none at all leaves the ordering to the kernel boundary alone.)
  if (tid == 0) {
    d.nwords[owner][slot] = s_bad ? kBadQueryTime : s_n;
    if (d.fence) __threadfence_system();
  }
}

// Barrier between the ranks' streams, in peer memory: rank r writes the epoch into slot r of every rank's flag
// row, then waits until every slot of its own row has reached it.  Everything a rank stored to peer memory before
// the barrier (kernels earlier in its stream) is visible to the peers' kernels after it.  A peer that never
// arrives traps after 20 s instead of hanging the GPU.
struct PeerFlags { uint32_t* flags[MFPA_MAX_PEERS]; int world, rank; };
__global__ void peer_barrier_kernel(const PeerFlags f, uint32_t epoch) {
  const int r = threadIdx.x;
  if (r >= f.world) return;
  __threadfence_system();
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f.flags[r] + f.rank), "r"(epoch) : "memory");
  const uint32_t* mine = f.flags[f.rank] + r;
  unsigned long long t0, t1;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((int32_t)(v - epoch) >= 0) break;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 20000000000ull) __trap();
    __nanosleep(200);
  }
  __threadfence_system();
}

// Visit the n words of one shard's list with `gsize` threads (t = this thread's index among them).  The visits are
// shared-memory atomics / look-ups that wait on the word, so the loads are batched ahead of them: four 16-byte
// loads (16 words) per thread in flight when the list is 16-byte aligned (it is for the library's own buffers),
// one word at a time otherwise.
template <typename F>
__device__ __forceinline__ void for_each_word(const uint32_t* __restrict__ w, int n, bool vec, int t, int gsize, F f) {
  if (!vec) {
    for (int i = t; i < n; i += gsize) f(__ldg(w + i));
    return;
  }
  const uint4* w4 = reinterpret_cast<const uint4*>(w);
  const int n4 = n >> 2;
  int i = t;
  for (; i + 3 * gsize < n4; i += 4 * gsize) {
    const uint4 a = __ldg(w4 + i), b = __ldg(w4 + i + gsize), c = __ldg(w4 + i + 2 * gsize), d = __ldg(w4 + i + 3 * gsize);
    f(a.x); f(a.y); f(a.z); f(a.w); f(b.x); f(b.y); f(b.z); f(b.w);
    f(c.x); f(c.y); f(c.z); f(c.w); f(d.x); f(d.y); f(d.z); f(d.w);
  }
  for (; i < n4; i += gsize) {
    const uint4 a = __ldg(w4 + i);
    f(a.x); f(a.y); f(a.z); f(a.w);
  }
  for (int j = (n4 << 2) + t; j < n; j += gsize) f(__ldg(w + j));
}

// The shards' lists of one query, visited by the block split into `groups` thread groups (a power of two <= the
// number of lists): with eight short lists, 128 threads per list keep as many loads in flight as 1024 threads on
// one long list do.
template <typename F>
__device__ __forceinline__ void for_each_list_word(const uint32_t* __restrict__ words_in, const int32_t* __restrict__ nwords,
                                                   int n_shards, int B, int q, int words_cap, bool vec, int groups, int tid,
                                                   F f) {
  const int gsize = kFusedThreads / groups, g = tid / gsize, t = tid - g * gsize;
  for (int l = g; l < n_shards; l += groups) {
    const int n = nwords[(int64_t)l * B + q];
    if (n < 0 || n > words_cap) continue;   // flagged by the caller
    for_each_word(words_in + ((int64_t)l * B + q) * words_cap, n, vec, t, gsize, f);
  }
}

// words: [n_shards][B][words_cap], nwords: [n_shards][B] (what the exchange leaves on the owner)
__global__ void __launch_bounds__(kFusedThreads, 1)
match_owner_kernel(const IndexView ix, const uint32_t* __restrict__ words_in, const int32_t* __restrict__ nwords, int n_shards,
                   int B, int words_cap, int threshcount, int search_depth, int32_t* __restrict__ cand,
                   int32_t* __restrict__ ncand, uint32_t* __restrict__ list, int list_cap, int32_t* __restrict__ nlist) {
  extern __shared__ __align__(16) unsigned fused_smem[];
  const int words = (ix.n_tracks + 1) >> 1;
  unsigned* hist = fused_smem;
  int* contenders = reinterpret_cast<int*>(fused_smem + ((words + 3) & ~3));   // kContCap ints
  __shared__ SelectSmem sel;
  __shared__ int s_n, s_err;
  __shared__ unsigned s_sum;
  const int q = blockIdx.x, tid = threadIdx.x;
  const bool vec = (reinterpret_cast<uintptr_t>(words_in) & 15) == 0 && (words_cap & 3) == 0;
  zero_hist(hist, words, tid);
  if (tid == 0) { s_n = 0; s_err = 0; }
  __syncthreads();
  unsigned total_hits = 0;
  for (int l = 0; l < n_shards; ++l) {
    const int n = nwords[(int64_t)l * B + q];
    if (n < 0 || n > words_cap) { if (tid == 0) s_err = n < 0 ? n : -1; continue; }
    total_hits += (unsigned)n;
  }
  int groups = 1;
  while (groups * 2 <= n_shards && groups < 32) groups *= 2;
  for_each_list_word(words_in, nwords, n_shards, B, q, words_cap, vec, groups, tid, [&](uint32_t v) {
    const unsigned id = v >> kHitDtBits;
    atomicAdd(&hist[id >> 1], 1u << ((id & 1u) << 4));
  });
  __syncthreads();
  if (s_err) {   // a shard's list overflowed (-1) or a query time was out of range: the host wrapper raises
    if (tid == 0) { ncand[q] = 0; nlist[q] = s_err; }
    return;
  }
  if (!hist_consistent(hist, words, total_hits, &s_sum, tid)) {
    if (tid == 0) { ncand[q] = 0; nlist[q] = kWideCounts; }
    return;
  }
  const int depth = fused_select(ix, hist, words, threshcount, search_depth, contenders, &sel,
                                 cand + (int64_t)q * search_depth * 2, ncand + q, tid);
  if (depth == 0) {
    if (tid == 0) nlist[q] = 0;
    return;
  }
  zero_hist(hist, words, tid);
  __syncthreads();
  const int nc = min(depth, 128);
  for (int k = tid; k < nc; k += kFusedThreads)
    reinterpret_cast<unsigned short*>(hist)[cand[((int64_t)q * search_depth + k) * 2]] = (unsigned short)(k + 1);
  __syncthreads();
  const unsigned short* mark = reinterpret_cast<const unsigned short*>(hist);
  uint32_t* out = list + (int64_t)q * list_cap;
  for_each_list_word(words_in, nwords, n_shards, B, q, words_cap, vec, groups, tid, [&](uint32_t v) {
    const unsigned k = mark[v >> kHitDtBits];
    if (k) {
      const int pos = atomicAdd(&s_n, 1);
      if (pos < list_cap) out[pos] = ((k - 1u) << 16) | (v & ((1u << kHitDtBits) - 1u));
    }
  });
  __syncthreads();
  if (tid == 0) nlist[q] = s_n;
}

// ---- collect -----------------------------------------------------------------------------
// (candidate index << 16) | (t_ref - t_q + 16384) for every local hit of a candidate track.
__global__ void __launch_bounds__(256) match_collect_kernel(const IndexView ix, const int32_t* __restrict__ hashes,
                                                            const int32_t* __restrict__ nh, int cap,
                                                            const int32_t* __restrict__ cand, const int32_t* __restrict__ ncand,
                                                            int search_depth, uint32_t* __restrict__ list, int list_cap,
                                                            int32_t* __restrict__ nlist) {
  __shared__ int s_ids[128];
  __shared__ int s_n;
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nc = min(ncand[q], min(search_depth, 128));
  for (int i = tid; i < nc; i += 256) s_ids[i] = cand[((int64_t)q * search_depth + i) * 2];
  if (tid == 0) s_n = 0;
  __syncthreads();
  if (nc > 0) {
    const int2* rows = reinterpret_cast<const int2*>(hashes) + (int64_t)q * cap;
    uint32_t* out = list + (int64_t)q * list_cap;
    const int n = min(nh[q], cap);
    const uint32_t tmask = (1u << ix.maxtimebits) - 1u;
    for (int r = warp; r < n; r += 8) {
      const int2 row = rows[r];
      const int h = (row.y & ix.hashmask) - ix.hash_lo;
      if (h < 0 || h >= ix.n_buckets) continue;
      const int cnt = min(ix.depth, ix.counts[h]);
      const uint32_t* bucket = ix.table + (int64_t)h * ix.depth;
      for (int s = lane; s < cnt; s += 32) {
        const uint32_t v = bucket[s];
        const int id = (int)(v >> ix.maxtimebits) - 1;
        for (int k = 0; k < nc; ++k)
          if (s_ids[k] == id) {
            const int dt = (int)(v & tmask) - row.x;
            const int pos = atomicAdd(&s_n, 1);
            if (pos < list_cap) out[pos] = ((uint32_t)k << 16) | (uint32_t)(dt + kDtOff);
            break;
          }
      }
    }
  }
  __syncthreads();
  if (tid == 0) nlist[q] = s_n;
}

// ---- align -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) match_align_kernel(const uint32_t* __restrict__ lists, const int32_t* __restrict__ nlists,
                                                          int n_lists, int B, int list_cap, const int32_t* __restrict__ cand,
                                                          const int32_t* __restrict__ ncand, int search_depth, int window,
                                                          int threshcount, int max_align, int32_t* __restrict__ results,
                                                          int32_t* __restrict__ nrows, int max_rows, int acap, int pass) {
  // Two passes over the batch: pass 0 runs every query with room for acap = kAlignSmall hits (12 KB of shared
  // memory: eight blocks per SM) and marks the few queries that need more as pending; pass 1 gives those the
  // full kAlignCap (96 KB) and leaves the others alone.
  extern __shared__ uint32_t sm[];
  uint32_t* keys = sm;                    // [acap] sorted hits
  uint32_t* bin_key = sm + acap;          // [acap] run starts: key of each distinct (cand, dt)
  int* bin_cnt = reinterpret_cast<int*>(sm + 2 * acap);  // [acap]
  __shared__ int s_total, s_rows, s_seg[130];
  const int q = blockIdx.x, tid = threadIdx.x;
  if (pass == 1 && nrows[q] != kAlignPending) return;
  const int nc = min(ncand[q], min(search_depth, 128));
  if (tid == 0) { s_total = 0; s_rows = 0; }
  __syncthreads();
  // gather the per-shard lists
  bool overflow = false;
  for (int l = 0; l < n_lists; ++l) {
    const int n = nlists[(int64_t)l * B + q];
    if (n < 0) {   // the producer flagged this query (bad query time, overflowed exchange list)
      if (tid == 0) nrows[q] = n;
      return;
    }
    if (n > list_cap) overflow = true;
    const uint32_t* src = lists + ((int64_t)l * B + q) * list_cap;
    __shared__ int s_base;
    if (tid == 0) { s_base = s_total; s_total += min(n, list_cap); }
    __syncthreads();
    for (int i = tid; i < min(n, list_cap); i += 256)
      if (s_base + i < acap) keys[s_base + i] = src[i];
    __syncthreads();
  }
  const int total = s_total;
  if (overflow || total > acap) {
    // pass 0: try again with the full capacity; pass 1 (or a shard list that overflowed): the host wrapper raises
    if (tid == 0) nrows[q] = (!overflow && pass == 0 && total <= kAlignCap) ? kAlignPending : -1;
    return;
  }
  if (nc == 0 || total == 0) {
    if (tid == 0) nrows[q] = 0;
    return;
  }
  // bitonic sort of the keys (padded with 0xffffffff)
  int np2 = 1;
  while (np2 < total) np2 <<= 1;
  for (int i = total + tid; i < np2; i += 256) keys[i] = 0xffffffffu;
  __syncthreads();
  for (int k = 2; k <= np2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < np2; i += 256) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const uint32_t a = keys[i], b = keys[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { keys[i] = b; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  // run-length encode (sequential over <= 8192 sorted keys; candidates are few)
  if (tid == 0) {
    int nb = 0;
    for (int c = 0; c <= nc; ++c) s_seg[c] = 0;
    for (int i = 0; i < total; ++i) {
      if (nb == 0 || bin_key[nb - 1] != keys[i]) { bin_key[nb] = keys[i]; bin_cnt[nb] = 0; ++nb; }
      ++bin_cnt[nb - 1];
    }
    // segment boundaries per candidate
    int b = 0;
    for (int c = 0; c < nc; ++c) {
      s_seg[c] = b;
      while (b < nb && (int)(bin_key[b] >> 16) == c) ++b;
    }
    s_seg[nc] = b;
  }
  __syncthreads();
  // one thread per candidate: modes of its delta-t histogram (audfprint_match.py:276-315)
  int32_t* res = results + (int64_t)q * max_rows * 7;
  if (tid < nc) {
    const int b0 = s_seg[tid], b1 = s_seg[tid + 1];
    const int id = cand[((int64_t)q * search_depth + tid) * 2], raw = cand[((int64_t)q * search_depth + tid) * 2 + 1];
    // "filtered" flag kept in the sign bit of a scratch copy: reuse keys[] region of this segment
    // (keys are no longer needed); keys[b] = filtered count
    for (int b = b0; b < b1; ++b) {
      const int dt = (int)(bin_key[b] & 0xffffu), c = bin_cnt[b];
      const int left = (b > b0 && (int)(bin_key[b - 1] & 0xffffu) == dt - 1) ? bin_cnt[b - 1] : 0;
      const int right = (b + 1 < b1 && (int)(bin_key[b + 1] & 0xffffu) == dt + 1) ? bin_cnt[b + 1] : 0;
      keys[b] = (c >= left && right < c) ? (uint32_t)c : 0u;
    }
    int found = 0;
    while (true) {
      int best = -1, bv = 0;
      for (int b = b0; b < b1; ++b)
        if ((int)keys[b] > bv) { bv = (int)keys[b]; best = b; }  // first maximum = smallest delta-t
      if (best < 0 || bv <= threshcount) break;
      const int mode = (int)(bin_key[best] & 0xffffu);
      int sum = 0;
      for (int b = b0; b < b1; ++b) {
        const int dt = (int)(bin_key[b] & 0xffffu);
        if (dt >= mode - window && dt <= mode + window) { sum += bin_cnt[b]; keys[b] = 0; }
      }
      const int row = atomicAdd(&s_rows, 1);
      if (row < max_rows) {
        int32_t* r = res + row * 7;
        r[0] = id; r[1] = sum; r[2] = mode - kDtOff; r[3] = raw; r[4] = tid; r[5] = 0; r[6] = 0;
      }
      if (++found > max_align) break;
    }
  }
  __syncthreads();
  // order rows by filtered count descending, then by candidate rank (:341; ties are arbitrary in numpy)
  const int nr = min(s_rows, max_rows);
  if (tid == 0) {
    for (int i = 1; i < nr; ++i) {
      int32_t t[7];
      for (int k = 0; k < 7; ++k) t[k] = res[i * 7 + k];
      int j = i - 1;
      while (j >= 0 && (res[j * 7 + 1] < t[1] || (res[j * 7 + 1] == t[1] && (res[j * 7 + 4] > t[4] ||
             (res[j * 7 + 4] == t[4] && res[j * 7 + 2] > t[2]))))) {
        for (int k = 0; k < 7; ++k) res[(j + 1) * 7 + k] = res[j * 7 + k];
        --j;
      }
      for (int k = 0; k < 7; ++k) res[(j + 1) * 7 + k] = t[k];
    }
    nrows[q] = s_rows > max_rows ? -2 : nr;
  }
}

// HashTable.get_hits for one query (hash_table.py:220-246): rows [id, t_ref - t_q, hash, t_q] in query order.
__global__ void __launch_bounds__(1024) get_hits_kernel(const IndexView ix, const int32_t* __restrict__ hashes, int n,
                                                        int32_t* __restrict__ hits, long long hits_cap,
                                                        long long* __restrict__ nhits) {
  __shared__ long long warp_tot[32];
  __shared__ long long carry;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int2* rows = reinterpret_cast<const int2*>(hashes);
  const uint32_t tmask = (1u << ix.maxtimebits) - 1u;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024) {
    const int r = base + tid;
    int cnt = 0, h = 0, t = 0, hb = -1;
    if (r < n) {
      t = rows[r].x; h = rows[r].y & ix.hashmask; hb = h - ix.hash_lo;
      if (hb >= 0 && hb < ix.n_buckets) cnt = min(ix.depth, ix.counts[hb]);
    }
    long long incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(kFull, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      long long w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(kFull, wi, o); if (lane >= o) wi += v; }
      warp_tot[lane] = wi - w;
    }
    __syncthreads();
    long long pos = carry + warp_tot[warp] + incl - cnt;
    if (cnt > 0) {
      const uint32_t* bucket = ix.table + (int64_t)hb * ix.depth;
      for (int s = 0; s < cnt; ++s, ++pos) {
        if (pos >= hits_cap) break;
        const uint32_t v = bucket[s];
        int4* o = reinterpret_cast<int4*>(hits) + pos;
        *o = make_int4((int)(v >> ix.maxtimebits) - 1, (int)(v & tmask) - t, h, t);
      }
    }
    __syncthreads();
    if (tid == 1023) carry = carry + warp_tot[31] + incl;
    __syncthreads();
  }
  if (tid == 0) *nhits = carry;
}

IndexView view(const mfpa_ctx* ctx) {
  IndexView v;
  v.table = ctx->index_table; v.counts = ctx->index_counts; v.hashesperid = ctx->index_hashesperid;
  v.hash_lo = ctx->index_hash_lo; v.n_buckets = ctx->index_hash_hi - ctx->index_hash_lo; v.depth = ctx->index_depth;
  v.maxtimebits = ctx->index_maxtimebits; v.n_tracks = ctx->index_ntracks; v.hashmask = ctx->index_hashmask;
  v.hp_min = ctx->index_hp_min;
  return v;
}

}  // namespace

int launch_match_counts(mfpa_ctx* ctx, const int32_t* hashes, const int32_t* nh, int B, int cap, int32_t* counts,
                        cudaStream_t st) {
  const IndexView ix = view(ctx);
  if (ix.n_tracks <= kMaxTracksSmem && ctx->opt_match_unfused != 2) {   // 2 = wide: int32 counters in global memory
    const size_t words = (size_t)((ix.n_tracks + 1) / 2);
    const size_t smem = sizeof(unsigned) * ((words + 3) & ~(size_t)3) + sizeof(RowCache);
    MFPA_CUDA(cudaFuncSetAttribute(match_counts_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_counts_sweep_kernel<<<B, kFusedThreads, smem, st>>>(ix, hashes, nh, cap, counts, ctx->opt_match_packed);
  } else {
    MFPA_REQUIRE(!ctx->opt_match_packed, "match_counts: packed counts need n_tracks <= %d", kMaxTracksSmem);
    MFPA_CUDA(cudaMemsetAsync(counts, 0, sizeof(int32_t) * (size_t)B * ix.n_tracks, st));
    match_counts_global_kernel<<<B, kCountThreads, 0, st>>>(ix, hashes, nh, cap, counts);
  }
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

bool match_fused_ok(const mfpa_ctx* ctx) { return ctx->index_ntracks <= kMaxTracksSmem; }

int launch_match_fused(mfpa_ctx* ctx, const int32_t* hashes, const int32_t* nh, int B, int cap, int threshcount,
                       int search_depth, int32_t* cand, int32_t* ncand, uint32_t* list, int list_cap, int32_t* nlist,
                       cudaStream_t st) {
  const IndexView ix = view(ctx);
  const size_t words = (size_t)((ix.n_tracks + 1) / 2);
  const size_t smem = sizeof(unsigned) * ((words + 3) & ~(size_t)3) + sizeof(RowCache);
  MFPA_CUDA(cudaFuncSetAttribute(match_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  match_fused_kernel<<<B, kFusedThreads, smem, st>>>(ix, hashes, nh, cap, threshcount, search_depth, cand, ncand, list,
                                                     list_cap, nlist);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

bool match_sparse_ok(const mfpa_ctx* ctx) { return ctx->index_ntracks <= kMaxTracksSmem && ctx->index_ntracks <= kHitMaxTracks; }

int launch_match_emit(mfpa_ctx* ctx, const int32_t* hashes, const int32_t* nh, int B, int cap, uint32_t* words, int words_cap,
                      int32_t* nwords, cudaStream_t st) {
  match_emit_kernel<<<B, kFusedThreads, 0, st>>>(view(ctx), hashes, nh, cap, words, words_cap, nwords);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_match_emit_peer(mfpa_ctx* ctx, const int32_t* hashes, const int32_t* nh, int B, int cap, const mfpa_peer_set* peers,
                           int words_cap, cudaStream_t st) {
  PeerDst d;
  for (int r = 0; r < MFPA_MAX_PEERS; ++r) { d.words[r] = peers->words[r]; d.nwords[r] = peers->nwords[r]; }
  d.own = B / peers->world;
  d.rank = peers->rank;
  d.world = peers->world;
  static const bool no_fence = getenv("MFPA_PEER_NO_FENCE") != nullptr;   // measurement knob
  d.fence = no_fence ? 0 : 1;
  match_emit_peer_kernel<<<B, kFusedThreads, 0, st>>>(view(ctx), hashes, nh, cap, d, words_cap);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_peer_barrier(const mfpa_peer_set* peers, uint32_t epoch, cudaStream_t st) {
  PeerFlags f;
  for (int r = 0; r < MFPA_MAX_PEERS; ++r) f.flags[r] = peers->flags[r];
  f.world = peers->world;
  f.rank = peers->rank;
  peer_barrier_kernel<<<1, 32, 0, st>>>(f, epoch);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_match_owner(mfpa_ctx* ctx, const uint32_t* words, const int32_t* nwords, int n_shards, int B, int words_cap,
                       int threshcount, int search_depth, int32_t* cand, int32_t* ncand, uint32_t* list, int list_cap,
                       int32_t* nlist, cudaStream_t st) {
  const IndexView ix = view(ctx);
  const size_t hw = (size_t)((ix.n_tracks + 1) / 2);
  const size_t smem = sizeof(unsigned) * ((hw + 3) & ~(size_t)3) + sizeof(int) * kContCap;
  MFPA_CUDA(cudaFuncSetAttribute(match_owner_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  match_owner_kernel<<<B, kFusedThreads, smem, st>>>(ix, words, nwords, n_shards, B, words_cap, threshcount, search_depth, cand,
                                                     ncand, list, list_cap, nlist);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_match_select(mfpa_ctx* ctx, const int32_t* counts, int B, int threshcount, int search_depth, int32_t* cand,
                        int32_t* ncand, cudaStream_t st) {
  match_select_kernel<<<B, 512, 0, st>>>(counts, ctx->index_hashesperid, ctx->index_ntracks, threshcount, search_depth,
                                         cand, ncand, ctx->opt_match_packed, ctx->index_hp_min);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_match_collect(mfpa_ctx* ctx, const int32_t* hashes, const int32_t* nh, int B, int cap, const int32_t* cand,
                         const int32_t* ncand, int search_depth, uint32_t* list, int list_cap, int32_t* nlist,
                         cudaStream_t st) {
  const IndexView ix = view(ctx);
  if (ix.n_tracks <= kMaxTracksSmem) {
    const size_t words = (size_t)((ix.n_tracks + 1) / 2);
    const size_t smem = sizeof(unsigned) * ((words + 3) & ~(size_t)3) + sizeof(RowCache);
    MFPA_CUDA(cudaFuncSetAttribute(match_collect_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    match_collect_sweep_kernel<<<B, kFusedThreads, smem, st>>>(ix, hashes, nh, cap, cand, ncand, search_depth, list, list_cap, nlist);
  } else {
    match_collect_kernel<<<B, 256, 0, st>>>(ix, hashes, nh, cap, cand, ncand, search_depth, list, list_cap, nlist);
  }
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_match_align(const uint32_t* lists, const int32_t* nlists, int n_lists, int B, int list_cap, const int32_t* cand,
                       const int32_t* ncand, int search_depth, int window, int threshcount, int max_align,
                       int32_t* results, int32_t* nrows, int max_rows, cudaStream_t st) {
  MFPA_CUDA(cudaFuncSetAttribute(match_align_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(uint32_t) * 3 * kAlignCap)));
  match_align_kernel<<<B, 256, sizeof(uint32_t) * 3 * kAlignSmall, st>>>(lists, nlists, n_lists, B, list_cap, cand, ncand,
                                                                         search_depth, window, threshcount, max_align,
                                                                         results, nrows, max_rows, kAlignSmall, 0);
  match_align_kernel<<<B, 256, sizeof(uint32_t) * 3 * kAlignCap, st>>>(lists, nlists, n_lists, B, list_cap, cand, ncand,
                                                                       search_depth, window, threshcount, max_align,
                                                                       results, nrows, max_rows, kAlignCap, 1);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

int launch_get_hits(mfpa_ctx* ctx, const int32_t* hashes, int n, int32_t* hits, int64_t hits_cap, int64_t* nhits,
                    cudaStream_t st) {
  get_hits_kernel<<<1, 1024, 0, st>>>(view(ctx), hashes, n, hits, (long long)hits_cap, (long long*)nhits);
  MFPA_CUDA(cudaGetLastError());
  return MFPA_OK;
}

}  // namespace mfpa
