// Device-wide exclusive prefix sum over uint32 (reduce-then-scan, 2048 items per CTA).
#pragma once
#include "common.cuh"

namespace nrt {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanBlock * kScanItems;

// Scans each tile independently (exclusive) and writes the tile total.
static __global__ void scan_tiles_kernel(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t n,
                                         uint32_t *__restrict__ tile_sums) {
  __shared__ uint32_t warp_sums[kScanBlock / 32];
  const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    uint32_t i = base + k;
    v[k] = i < n ? in[i] : 0u;
    sum += v[k];
  }
  // inclusive warp scan of the per-thread sums
  uint32_t inc = sum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    uint32_t w = lane < kScanBlock / 32 ? warp_sums[lane] : 0u;
    uint32_t wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xFFFFFFFFu, wi, o);
      if (lane >= o) wi += t;
    }
    if (lane < kScanBlock / 32) warp_sums[lane] = wi - w;  // exclusive warp offsets
    if (lane == kScanBlock / 32 - 1 && tile_sums) tile_sums[blockIdx.x] = wi;
  }
  __syncthreads();
  uint32_t run = warp_sums[wid] + inc - sum;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    uint32_t i = base + k;
    if (i < n) out[i] = run;
    run += v[k];
  }
}

static __global__ void add_tile_offsets_kernel(uint32_t *__restrict__ out, uint32_t n,
                                               const uint32_t *__restrict__ tile_offsets) {
  const uint32_t off = tile_offsets[blockIdx.x];
  const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
#pragma unroll
  for (int k = 0; k < kScanItems; k++) {
    uint32_t i = base + k;
    if (i < n) out[i] += off;
  }
}

// out[i] = sum(in[0..i)).  `total` (optional, host) receives sum(in[0..n)) -- that read synchronises
// the stream.  in == out is allowed.
static int exclusive_scan_u32(const uint32_t *d_in, uint32_t *d_out, uint32_t n, uint32_t *total, cudaStream_t s) {
  if (n == 0) {
    if (total) *total = 0;
    return NRT_OK;
  }
  uint32_t tiles = (n + kScanTile - 1) / kScanTile;
  uint32_t *d_sums = nullptr;
  NRT_CUDA(cudaMalloc(&d_sums, sizeof(uint32_t) * (size_t)(tiles + 1)));
  scan_tiles_kernel<<<tiles, kScanBlock, 0, s>>>(d_in, d_out, n, d_sums);
  NRT_CUDA(cudaGetLastError());
  int rc = NRT_OK;
  if (tiles > 1) {
    uint32_t t2 = 0;
    rc = exclusive_scan_u32(d_sums, d_sums, tiles, total ? &t2 : nullptr, s);
    if (rc == NRT_OK) {
      add_tile_offsets_kernel<<<tiles, kScanBlock, 0, s>>>(d_out, n, d_sums);
      if (cudaGetLastError() != cudaSuccess) rc = NRT_ERR_CUDA;
    }
    if (total) *total = t2;
  } else if (total) {
    if (cudaMemcpyAsync(total, d_sums, sizeof(uint32_t), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess)
      rc = NRT_ERR_CUDA;
  }
  if (cudaStreamSynchronize(s) != cudaSuccess) rc = NRT_ERR_CUDA;
  cudaFree(d_sums);
  return rc;
}

// Scratch words needed by exclusive_scan_u32_async for n items.
static inline size_t scan_scratch_words(uint32_t n) {
  size_t total = 0;
  while (n > 1) {
    uint32_t tiles = (n + kScanTile - 1) / kScanTile;
    total += tiles + 1;
    if (tiles <= 1) break;
    n = tiles;
  }
  return total + 4;
}

// Same scan with caller-provided scratch; nothing is allocated, nothing synchronises.
static int exclusive_scan_u32_async(const uint32_t *d_in, uint32_t *d_out, uint32_t n, uint32_t *d_scratch,
                                    cudaStream_t s) {
  if (n == 0) return NRT_OK;
  uint32_t tiles = (n + kScanTile - 1) / kScanTile;
  scan_tiles_kernel<<<tiles, kScanBlock, 0, s>>>(d_in, d_out, n, d_scratch);
  NRT_CUDA(cudaGetLastError());
  if (tiles > 1) {
    int rc = exclusive_scan_u32_async(d_scratch, d_scratch, tiles, d_scratch + tiles + 1, s);
    if (rc != NRT_OK) return rc;
    add_tile_offsets_kernel<<<tiles, kScanBlock, 0, s>>>(d_out, n, d_scratch);
    NRT_CUDA(cudaGetLastError());
  }
  return NRT_OK;
}

}  // namespace nrt
