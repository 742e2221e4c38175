// Index build of `ghostm db` on the device: the stable sort behind the CSR k-mer table.
//
// The reference builds positions[] with a counting sort over all 32^w keys on one host thread
// (db_creator.cpp:167-241): histogram, prefix sums, then a scatter in ascending db offset, so the
// offsets of one key stay ascending.  Here the same permutation is produced by a hand-written
// least-significant-digit counting sort in passes of 8 bits over (key, offset) pairs - every pass
// is itself a counting sort (per-block digit histogram, one global exclusive scan in digit-major
// order, stable scatter), and stability of every pass makes the whole sort stable, i.e. equal keys
// keep their ascending offsets exactly like the reference's scatter loop.
//
// A block owns a contiguous tile of 4096 pairs, a warp 512 of them in 16 rounds of 32 consecutive
// pairs; the rank of a pair inside its round comes from __match_any_sync (peers with the same digit
// below my lane), across rounds from a running per-warp base in shared memory, across warps and
// blocks from the scanned histograms.  Non-indexable offsets carry the key 0xFFFFFFFF and end up
// behind every real key as long as the passes cover one bit more than the key width.
#include "gm_common.cuh"

namespace gm {

namespace {

constexpr int kRxWarps = 8;
constexpr int kRxThreads = kRxWarps * 32;
constexpr int kRxRounds = 16;
constexpr uint32_t kRxTile = kRxThreads * kRxRounds;      // pairs per block
constexpr uint32_t kFull = 0xFFFFFFFFu;

// blockhist[d * n_blocks + block] = pairs of the block's tile whose digit is d
__global__ void __launch_bounds__(kRxThreads) rx_hist_kernel(const uint32_t *__restrict__ keys, uint32_t n,
                                                             uint32_t shift, uint32_t *__restrict__ blockhist,
                                                             uint32_t n_blocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kRxTile;
#pragma unroll 4
  for (uint32_t k = 0; k < (uint32_t)kRxRounds; ++k) {
    const uint32_t i = base + k * kRxThreads + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  blockhist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

// ---- exclusive scan of a u32 array in three steps (tiles of 4096, totals, add) ----------------
constexpr uint32_t kScanTile = 4096;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_sums, uint32_t *total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, incl, o);
    if (lane >= (uint32_t)o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t s = lane < blockDim.x / 32 ? warp_sums[lane] : 0u, si = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, si, o);
      if (lane >= (uint32_t)o) si += t;
    }
    warp_sums[lane] = si - s;
    if (lane == 31) *total = si;
  }
  __syncthreads();
  return incl - v + warp_sums[warp];
}

// every thread scans 16 consecutive entries; tile = 256 threads x 16
__global__ void __launch_bounds__(256) scan_tiles_kernel(uint32_t *data, size_t n, uint32_t *tile_sums) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t total;
  const size_t base = (size_t)blockIdx.x * kScanTile + (size_t)threadIdx.x * 16;
  uint32_t v[16], sum = 0;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    v[k] = base + k < n ? data[base + k] : 0u;
    sum += v[k];
  }
  uint32_t run = block_exclusive_scan(sum, warp_sums, &total);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    if (base + k < n) data[base + k] = run;
    run += v[k];
  }
  if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// one block: exclusive scan of the tile totals (any number, chunk by chunk with a carry)
__global__ void __launch_bounds__(1024) scan_totals_kernel(uint32_t *tile_sums, uint32_t n_tiles) {
  __shared__ uint32_t warp_sums[32];
  __shared__ uint32_t total;
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n_tiles; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n_tiles ? tile_sums[i] : 0u;
    const uint32_t ex = block_exclusive_scan(v, warp_sums, &total);
    if (i < n_tiles) tile_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) scan_add_kernel(uint32_t *data, size_t n, const uint32_t *tile_sums) {
  const uint32_t add = tile_sums[blockIdx.x];
  const size_t base = (size_t)blockIdx.x * kScanTile;
#pragma unroll 4
  for (uint32_t k = 0; k < 16; ++k) {
    const size_t i = base + k * 256 + threadIdx.x;
    if (i < n) data[i] += add;
  }
}

// stable scatter of one pass: offs = exclusive scan of blockhist (digit-major)
__global__ void __launch_bounds__(kRxThreads) rx_scatter_kernel(const uint32_t *__restrict__ keys_in,
                                                                const uint32_t *__restrict__ vals_in,
                                                                uint32_t *__restrict__ keys_out,
                                                                uint32_t *__restrict__ vals_out, uint32_t n,
                                                                uint32_t shift, const uint32_t *__restrict__ offs,
                                                                uint32_t n_blocks) {
  __shared__ uint32_t wh[kRxWarps][256];     // per-warp digit counts, then running output bases
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int w = 0; w < kRxWarps; ++w) wh[w][threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kRxTile + warp * (32u * kRxRounds);
  uint32_t k[kRxRounds];
#pragma unroll
  for (int r = 0; r < kRxRounds; ++r) {
    const uint32_t i = base + r * 32 + lane;
    k[r] = i < n ? keys_in[i] : 0u;
    if (i < n) atomicAdd(&wh[warp][(k[r] >> shift) & 255u], 1u);
  }
  __syncthreads();
  {   // digit = threadIdx.x: bases of the warps in warp order behind the block's global offset
    uint32_t run = offs[(size_t)threadIdx.x * n_blocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kRxWarps; ++w) {
      const uint32_t c = wh[w][threadIdx.x];
      wh[w][threadIdx.x] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kRxRounds; ++r) {
    const uint32_t i = base + r * 32 + lane;
    const bool valid = i < n;
    const uint32_t d = (k[r] >> shift) & 255u;
    const uint32_t peers = __match_any_sync(kFull, valid ? d : 256u + lane);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    uint32_t dest = 0;
    if (valid) dest = wh[warp][d] + rank;
    __syncwarp();
    if (valid && rank + 1 == (uint32_t)__popc(peers)) wh[warp][d] = dest + 1;   // the last peer moves the base
    __syncwarp();
    if (valid) {
      keys_out[dest] = k[r];
      vals_out[dest] = vals_in[i];
    }
  }
}

}  // namespace

size_t index_sort_scratch_words(uint32_t n) {
  const uint32_t n_blocks = (n + kRxTile - 1) / kRxTile;
  const size_t hist = (size_t)256 * n_blocks;
  return hist + (hist + kScanTile - 1) / kScanTile + 16;
}

// Stable sort of n (key, value) pairs by the low `bits` bits of the key (rounded up to whole
// 8-bit passes).  keys / vals are destroyed; the sorted keys end in *sorted_keys (one of keys,
// keys_tmp), the sorted values in vals_final.  scratch: index_sort_scratch_words(n) words.
cudaError_t index_sort_pairs(uint32_t *keys, uint32_t *vals, uint32_t *keys_tmp, uint32_t *vals_tmp,
                             uint32_t *vals_final, uint32_t n, uint32_t bits, uint32_t *scratch,
                             uint32_t **sorted_keys, uint32_t *launches, cudaStream_t stream) {
  const uint32_t n_blocks = (n + kRxTile - 1) / kRxTile;
  const size_t hist_len = (size_t)256 * n_blocks;
  const uint32_t n_tiles = (uint32_t)((hist_len + kScanTile - 1) / kScanTile);
  uint32_t *blockhist = scratch, *tile_sums = scratch + hist_len;
  const uint32_t passes = (bits + 7) / 8;
  uint32_t *kin = keys, *vin = vals, *kout = keys_tmp, *vout = vals_tmp;
  for (uint32_t pass = 0; pass < passes; ++pass) {
    if (pass + 1 == passes) vout = vals_final;
    rx_hist_kernel<<<n_blocks, kRxThreads, 0, stream>>>(kin, n, pass * 8, blockhist, n_blocks);
    scan_tiles_kernel<<<n_tiles, 256, 0, stream>>>(blockhist, hist_len, tile_sums);
    scan_totals_kernel<<<1, 1024, 0, stream>>>(tile_sums, n_tiles);
    scan_add_kernel<<<n_tiles, 256, 0, stream>>>(blockhist, hist_len, tile_sums);
    rx_scatter_kernel<<<n_blocks, kRxThreads, 0, stream>>>(kin, vin, kout, vout, n, pass * 8, blockhist, n_blocks);
    if (launches) *launches += 5;
    uint32_t *t = kin; kin = kout; kout = t;
    t = vin; vin = (vout == vals_final ? vals_final : vout); vout = t;
  }
  *sorted_keys = kin;
  return cudaGetLastError();
}

}  // namespace gm
