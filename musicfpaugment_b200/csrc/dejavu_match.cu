// In-memory Dejavu index and offset vote: replaces the Postgres round trips of
// CommonDatabase.return_matches (afp/dejavu/postgres_database.py:180-229: one SELECT per query hash) and the
// sort / groupby vote of Dejavu.align_matches (afp/dejavu/dejavu.py:312-378) for the top match.
//
// Index: the fingerprints table as four parallel arrays sorted by hash.  A Dejavu hash is the first 20 hex
// digits of a SHA-1 (FINGERPRINT_REDUCTION, variables.py:22) = 80 bits: `key` holds the leading 64, `tail` the
// other 16, so equality is exact.  Lookup = binary search of the key range, one thread per distinct query hash;
// every matching row yields one (song, db offset - query offset) pair per query offset of that hash and counts
// once towards the song's matched-hash total (dedup_hashes).  Vote = multiplicity of every distinct pair in an
// open-addressing table in global memory, then the best entry by (count desc, song asc, offset asc) - the order
// align_matches' stable sorts produce.
#include "common.cuh"

struct mfpa_dejavu_index {
  mfpa_ctx* ctx = nullptr;
  uint64_t* keys = nullptr;
  uint16_t* tails = nullptr;
  int32_t* songs = nullptr;
  int32_t* offsets = nullptr;
  int64_t n = 0;
  int n_songs = 0;
  mfpa::Scratch table;   // vote table: u64 keys then u32 counts
};

namespace mfpa {

namespace {

constexpr unsigned long long kEmpty = ~0ull;

__global__ void __launch_bounds__(256) dejavu_lookup_kernel(const uint64_t* __restrict__ keys, const uint16_t* __restrict__ tails,
                                                            const int32_t* __restrict__ songs, const int32_t* __restrict__ offs,
                                                            long long n, const uint64_t* __restrict__ qkeys,
                                                            const uint16_t* __restrict__ qtails, const int32_t* __restrict__ qstart,
                                                            const int32_t* __restrict__ qoffs, int nq, int n_songs,
                                                            int32_t* __restrict__ pairs, long long pairs_cap,
                                                            unsigned long long* __restrict__ n_pairs, int32_t* __restrict__ dedup) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const uint64_t k = qkeys[i];
  const uint16_t t = qtails[i];
  long long lo = 0, hi = n;
  while (lo < hi) {   // first row with key >= k
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] < k) lo = mid + 1; else hi = mid;
  }
  const int o0 = qstart[i], o1 = qstart[i + 1];
  for (long long r = lo; r < n && keys[r] == k; ++r) {
    if (tails[r] != t) continue;
    const int sid = songs[r];
    if (sid >= 0 && sid < n_songs) atomicAdd(dedup + sid, 1);
    const unsigned long long pos = atomicAdd(n_pairs, (unsigned long long)(o1 - o0));
    for (int o = o0; o < o1; ++o) {
      const long long p = (long long)pos + (o - o0);
      if (p < pairs_cap) { pairs[2 * p] = sid; pairs[2 * p + 1] = offs[r] - qoffs[o]; }
    }
  }
}

__device__ __forceinline__ unsigned long long pair_key(int sid, int diff) {
  return ((unsigned long long)(unsigned)sid << 32) | (unsigned long long)((unsigned)diff ^ 0x80000000u);   // diff order-preserving
}
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

__global__ void __launch_bounds__(256) dejavu_vote_insert_kernel(const int32_t* __restrict__ pairs, long long n_pairs,
                                                                 unsigned long long* __restrict__ tkeys, unsigned* __restrict__ tcnt,
                                                                 unsigned long long mask) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += (long long)gridDim.x * blockDim.x) {
    const unsigned long long k = pair_key(pairs[2 * i], pairs[2 * i + 1]);
    unsigned long long slot = mix64(k) & mask;
    while (true) {
      const unsigned long long prev = atomicCAS(tkeys + slot, kEmpty, k);
      if (prev == kEmpty || prev == k) { atomicAdd(tcnt + slot, 1u); break; }
      slot = (slot + 1) & mask;
    }
  }
}

// one block: best entry by (count desc, key asc) - key asc = song asc, then offset difference asc
__global__ void __launch_bounds__(1024) dejavu_vote_best_kernel(const unsigned long long* __restrict__ tkeys,
                                                                const unsigned* __restrict__ tcnt, unsigned long long size,
                                                                int32_t* __restrict__ best) {
  __shared__ unsigned s_cnt[32];
  __shared__ unsigned long long s_key[32];
  unsigned bc = 0;
  unsigned long long bk = kEmpty;
  for (unsigned long long i = threadIdx.x; i < size; i += 1024) {
    const unsigned long long k = tkeys[i];
    if (k == kEmpty) continue;
    const unsigned c = tcnt[i];
    if (c > bc || (c == bc && k < bk)) { bc = c; bk = k; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const unsigned c = __shfl_xor_sync(0xffffffffu, bc, o);
    const unsigned long long k = __shfl_xor_sync(0xffffffffu, bk, o);
    if (c > bc || (c == bc && k < bk)) { bc = c; bk = k; }
  }
  if ((threadIdx.x & 31) == 0) { s_cnt[threadIdx.x >> 5] = bc; s_key[threadIdx.x >> 5] = bk; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w)
      if (s_cnt[w] > bc || (s_cnt[w] == bc && s_key[w] < bk)) { bc = s_cnt[w]; bk = s_key[w]; }
    best[0] = bc ? (int32_t)(bk >> 32) : -1;
    best[1] = bc ? (int32_t)((unsigned)(bk & 0xffffffffu) ^ 0x80000000u) : 0;
    best[2] = (int32_t)bc;
  }
}

}  // namespace

}  // namespace mfpa

using namespace mfpa;

extern "C" {

int mfpa_dejavu_index_create(mfpa_ctx* ctx, const uint64_t* keys_host, const uint16_t* tails_host, const int32_t* songs_host,
                             const int32_t* offsets_host, int64_t n, int n_songs, mfpa_dejavu_index** out) {
  MFPA_REQUIRE(ctx && out && n >= 0 && n_songs >= 0 && (n == 0 || (keys_host && tails_host && songs_host && offsets_host)),
               "dejavu_index_create: bad argument");
  for (int64_t i = 1; i < n; ++i)
    MFPA_REQUIRE(keys_host[i - 1] <= keys_host[i], "dejavu_index_create: keys are not sorted (row %lld)", (long long)i);
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(ctx->device);
  mfpa_dejavu_index* ix = new mfpa_dejavu_index();
  ix->ctx = ctx; ix->n = n; ix->n_songs = n_songs;
  const size_t m = (size_t)(n > 0 ? n : 1);
  cudaError_t e = cudaMalloc(&ix->keys, sizeof(uint64_t) * m);
  if (e == cudaSuccess) e = cudaMalloc(&ix->tails, sizeof(uint16_t) * m);
  if (e == cudaSuccess) e = cudaMalloc(&ix->songs, sizeof(int32_t) * m);
  if (e == cudaSuccess) e = cudaMalloc(&ix->offsets, sizeof(int32_t) * m);
  if (e == cudaSuccess && n > 0) {
    cudaMemcpy(ix->keys, keys_host, sizeof(uint64_t) * n, cudaMemcpyHostToDevice);
    cudaMemcpy(ix->tails, tails_host, sizeof(uint16_t) * n, cudaMemcpyHostToDevice);
    cudaMemcpy(ix->songs, songs_host, sizeof(int32_t) * n, cudaMemcpyHostToDevice);
    e = cudaMemcpy(ix->offsets, offsets_host, sizeof(int32_t) * n, cudaMemcpyHostToDevice);
  }
  if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
  if (e != cudaSuccess) {
    set_error("dejavu_index_create: %s", cudaGetErrorString(e));
    mfpa_dejavu_index_destroy(ix);
    return MFPA_ECUDA;
  }
  *out = ix;
  return MFPA_OK;
}

void mfpa_dejavu_index_destroy(mfpa_dejavu_index* ix) {
  if (!ix) return;
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(ix->ctx->device);
  if (ix->keys) cudaFree(ix->keys);
  if (ix->tails) cudaFree(ix->tails);
  if (ix->songs) cudaFree(ix->songs);
  if (ix->offsets) cudaFree(ix->offsets);
  ix->table.release();
  if (prev >= 0 && prev != ix->ctx->device) cudaSetDevice(prev);
  delete ix;
}

int mfpa_dejavu_return_matches(mfpa_dejavu_index* ix, const uint64_t* qkeys_dev, const uint16_t* qtails_dev,
                               const int32_t* qstart_dev, const int32_t* qoffsets_dev, int n_hashes, int32_t* pairs_dev,
                               int64_t pairs_cap, int64_t* n_pairs_dev, int32_t* dedup_dev, void* stream) {
  MFPA_REQUIRE(ix && n_pairs_dev && dedup_dev && n_hashes >= 0 && pairs_cap >= 0, "dejavu_return_matches: bad argument");
  MFPA_REQUIRE(n_hashes == 0 || (qkeys_dev && qtails_dev && qstart_dev && qoffsets_dev), "dejavu_return_matches: NULL query arrays");
  MFPA_REQUIRE(pairs_cap == 0 || pairs_dev, "dejavu_return_matches: pairs_dev is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != ix->ctx->device) cudaSetDevice(ix->ctx->device);
  cudaMemsetAsync(n_pairs_dev, 0, sizeof(int64_t), st);
  if (ix->n_songs > 0) cudaMemsetAsync(dedup_dev, 0, sizeof(int32_t) * (size_t)ix->n_songs, st);
  if (n_hashes > 0 && ix->n > 0)
    dejavu_lookup_kernel<<<(n_hashes + 255) / 256, 256, 0, st>>>(ix->keys, ix->tails, ix->songs, ix->offsets, (long long)ix->n, qkeys_dev,
                                                                qtails_dev, qstart_dev, qoffsets_dev, n_hashes, ix->n_songs, pairs_dev,
                                                                (long long)pairs_cap, (unsigned long long*)n_pairs_dev, dedup_dev);
  const cudaError_t e = cudaGetLastError();
  if (prev >= 0 && prev != ix->ctx->device) cudaSetDevice(prev);
  if (e != cudaSuccess) { set_error("dejavu_return_matches: %s", cudaGetErrorString(e)); return MFPA_ECUDA; }
  return MFPA_OK;
}

int mfpa_dejavu_align(mfpa_dejavu_index* ix, const int32_t* pairs_dev, int64_t n_pairs, int32_t* best_dev, void* stream) {
  MFPA_REQUIRE(ix && best_dev && n_pairs >= 0 && (n_pairs == 0 || pairs_dev), "dejavu_align: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  int prev = -1;
  cudaGetDevice(&prev);
  if (prev != ix->ctx->device) cudaSetDevice(ix->ctx->device);
  unsigned long long size = 1024;
  while (size < 2ull * (unsigned long long)n_pairs) size <<= 1;
  int rc = MFPA_OK;
  if (ix->table.reserve((sizeof(unsigned long long) + sizeof(unsigned)) * size)) rc = MFPA_ENOMEM;
  if (rc == MFPA_OK) {
    unsigned long long* tkeys = (unsigned long long*)ix->table.ptr;
    unsigned* tcnt = (unsigned*)(tkeys + size);
    cudaMemsetAsync(tkeys, 0xff, sizeof(unsigned long long) * size, st);
    cudaMemsetAsync(tcnt, 0, sizeof(unsigned) * size, st);
    if (n_pairs > 0) {
      const unsigned blocks = (unsigned)((n_pairs + 255) / 256 < 2368 ? (n_pairs + 255) / 256 : 2368);
      dejavu_vote_insert_kernel<<<blocks, 256, 0, st>>>(pairs_dev, (long long)n_pairs, tkeys, tcnt, size - 1);
    }
    dejavu_vote_best_kernel<<<1, 1024, 0, st>>>(tkeys, tcnt, size, best_dev);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("dejavu_align: %s", cudaGetErrorString(e)); rc = MFPA_ECUDA; }
  }
  if (prev >= 0 && prev != ix->ctx->device) cudaSetDevice(prev);
  return rc;
}

}  // extern "C"
