// Device-wide stable LSD radix sort of (uint32 key, uint32 value) pairs, 4 bits per pass, and the 30-bit
// Morton code that orders the builder's primitives along a space-filling curve before the binned-SAH
// sweep (gathers of neighbouring slots then hit neighbouring memory, and the stable partitions of the
// sweep keep that order inside every node).
#pragma once
#include "common.cuh"
#include "scan.cuh"

namespace nrt {

constexpr int kSortBlock = 256;
constexpr int kSortItems = 4;  // consecutive keys per thread (blocked arrangement keeps the sort stable)
constexpr int kSortTile = kSortBlock * kSortItems;
constexpr int kSortDigits = 16;

__device__ __forceinline__ uint32_t spread_bits_10(uint32_t v) {  // 10 bits -> every third bit
  v &= 0x3FFu;
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

// centroid = (plo.w, phi.w, pcz); scene box as ordered-uint keys decoded by the caller
static __global__ void morton_kernel(const float4 *__restrict__ plo, const float4 *__restrict__ phi,
                                     const float *__restrict__ pcz, uint32_t n, float3 smin, float3 sinv,
                                     uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x = (plo[i].w - smin.x) * sinv.x, y = (phi[i].w - smin.y) * sinv.y, z = (pcz[i] - smin.z) * sinv.z;
  int xi = min(1023, max(0, (int)x)), yi = min(1023, max(0, (int)y)), zi = min(1023, max(0, (int)z));
  keys[i] = (spread_bits_10((uint32_t)xi) << 2) | (spread_bits_10((uint32_t)yi) << 1) | spread_bits_10((uint32_t)zi);
  vals[i] = i;
}

// table[d * n_tiles + tile] = number of keys of `tile` whose digit is d
static __global__ void __launch_bounds__(kSortBlock)
    radix_hist_kernel(const uint32_t *__restrict__ keys, uint32_t n, int shift, uint32_t n_tiles,
                      uint32_t *__restrict__ table) {
  __shared__ uint32_t cnt[kSortDigits];
  if (threadIdx.x < kSortDigits) cnt[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kSortTile + threadIdx.x * kSortItems;
#pragma unroll
  for (int k = 0; k < kSortItems; k++) {
    if (base + k < n) atomicAdd(&cnt[(keys[base + k] >> shift) & 15u], 1u);
  }
  __syncthreads();
  if (threadIdx.x < kSortDigits) table[threadIdx.x * n_tiles + blockIdx.x] = cnt[threadIdx.x];
}

// offsets = exclusive scan of `table` (digit-major), so offsets[d*n_tiles+tile] is the first output slot
// of (digit d, tile).  Inside the tile, keys keep their order (thread-major, item-minor).
static __global__ void __launch_bounds__(kSortBlock)
    radix_scatter_kernel(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint32_t n, int shift,
                         uint32_t n_tiles, const uint32_t *__restrict__ offsets, uint32_t *__restrict__ keys_out,
                         uint32_t *__restrict__ vals_out) {
  __shared__ uint32_t cnt[kSortDigits * kSortBlock];  // [digit][thread], later its exclusive scan
  __shared__ uint32_t warp_sums[kSortBlock / 32];
  __shared__ uint32_t digit_start[kSortDigits];
  const int t = threadIdx.x;
  const uint32_t base = blockIdx.x * kSortTile + t * kSortItems;
  uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];
#pragma unroll
  for (int d = 0; d < kSortDigits; d++) cnt[d * kSortBlock + t] = 0;
#pragma unroll
  for (int k = 0; k < kSortItems; k++) {
    if (base + k < n) {
      key[k] = keys[base + k];
      val[k] = vals[base + k];
      const uint32_t d = (key[k] >> shift) & 15u;
      rank[k] = cnt[d * kSortBlock + t]++;  // own column: no conflicts, order of items kept
    }
  }
  __syncthreads();
  // exclusive scan of the 16*256 counters in digit-major order; thread t owns entries [16t, 16t+16)
  uint32_t loc[kSortDigits], sum = 0;
#pragma unroll
  for (int j = 0; j < kSortDigits; j++) {
    loc[j] = cnt[t * kSortDigits + j];
    sum += loc[j];
  }
  uint32_t inc = sum;
  const int lane = t & 31, wid = t >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t x = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += x;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < wid; w++) woff += warp_sums[w];
  uint32_t run = woff + inc - sum;
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kSortDigits; j++) {
    cnt[t * kSortDigits + j] = run;
    run += loc[j];
  }
  __syncthreads();
  if (t < kSortDigits) digit_start[t] = cnt[t * kSortBlock];
  __syncthreads();
#pragma unroll
  for (int k = 0; k < kSortItems; k++) {
    if (base + k < n) {
      const uint32_t d = (key[k] >> shift) & 15u;
      const uint32_t pos = offsets[d * n_tiles + blockIdx.x] + (cnt[d * kSortBlock + t] - digit_start[d]) + rank[k];
      keys_out[pos] = key[k];
      vals_out[pos] = val[k];
    }
  }
}

// Sorts in place logically: on return (keys, vals) hold the sorted pairs (the *_tmp arrays are scratch).
// `table` needs 16 * n_tiles + 1 words, `scratch` scan_scratch_words(16 * n_tiles) words.
// Only key bits [first_bit, key_bits) take part (a stable sort on the high bits alone: equal prefixes keep their input
// order).
static int radix_sort_pairs(uint32_t *&keys, uint32_t *&vals, uint32_t *&keys_tmp, uint32_t *&vals_tmp, uint32_t n,
                            int first_bit, int key_bits, uint32_t *table, uint32_t *scratch, cudaStream_t s) {
  const uint32_t n_tiles = (n + kSortTile - 1) / kSortTile;
  for (int shift = first_bit; shift < key_bits; shift += 4) {
    radix_hist_kernel<<<n_tiles, kSortBlock, 0, s>>>(keys, n, shift, n_tiles, table);
    NRT_CUDA(cudaGetLastError());
    int rc = exclusive_scan_u32_async(table, table, kSortDigits * n_tiles, scratch, s);
    if (rc != NRT_OK) return rc;
    radix_scatter_kernel<<<n_tiles, kSortBlock, 0, s>>>(keys, vals, n, shift, n_tiles, table, keys_tmp, vals_tmp);
    NRT_CUDA(cudaGetLastError());
    uint32_t *t = keys;
    keys = keys_tmp;
    keys_tmp = t;
    t = vals;
    vals = vals_tmp;
    vals_tmp = t;
  }
  return NRT_OK;
}

static __global__ void gather_prims_kernel(const uint32_t *__restrict__ order, const float4 *__restrict__ plo_u,
                                           const float4 *__restrict__ phi_u, const float *__restrict__ pcz_u,
                                           uint32_t n, float4 *__restrict__ plo, float4 *__restrict__ phi,
                                           float *__restrict__ pcz) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t p = order[i];
  plo[i] = plo_u[p];
  phi[i] = phi_u[p];
  pcz[i] = pcz_u[p];
}

static __global__ void map_indices_kernel(const uint32_t *__restrict__ slots, const uint32_t *__restrict__ order,
                                          uint32_t n, uint32_t *__restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = order[slots[i]];
}

}  // namespace nrt
