// Measured roofs for bench.py's roofline block: streaming-read bandwidth of a buffer that fits the L2 (126 MB) and of
// one that does not (HBM), and the pinned host <-> device copy rates -- measured in the same process, on the same
// device and clocks as the traversal they are compared with.  Not part of the traversal path.
#include "common.cuh"

namespace nrt {
namespace {

__global__ void __launch_bounds__(256) read_sum_kernel(const float4 *__restrict__ p, size_t n4, float *sink) {
  float acc = 0.0f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(p + i);
    acc += (v.x + v.y) + (v.z + v.w);
  }
  if (acc == 1.2345e-30f) *sink = acc;  // keeps the loads alive, never true in practice
}

}  // namespace
}  // namespace nrt

using namespace nrt;

// Reads `bytes` of device memory `iters` times with 16-byte loads from every SM; *gb_per_s = the best iteration.
// bytes <= ~64 MB stays L2-resident after the first pass (L2 read roof), bytes >= 1 GB measures HBM.
extern "C" int nrt_probe_read_gbs(size_t bytes, int iters, double *gb_per_s) {
  if (!gb_per_s || bytes < 4096 || iters < 1) {
    set_error("nrt_probe_read_gbs: bad arguments");
    return NRT_ERR_INVALID;
  }
  int device = 0;
  DeviceGuard dg_caller;
  int rc = select_device(&device);
  if (rc != NRT_OK) return rc;
  float4 *d = nullptr;
  float *sink = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const size_t n4 = bytes / sizeof(float4);
  cudaError_t e = cudaMalloc(&d, n4 * sizeof(float4));
  if (e == cudaSuccess) e = cudaMalloc(&sink, sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(d, 0, n4 * sizeof(float4));
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  double best = 0.0;
  if (e == cudaSuccess) {
    const unsigned grid = (unsigned)device_sm_count(device) * 8u;
    read_sum_kernel<<<grid, 256>>>(d, n4, sink);  // warm-up (and L2 fill for small buffers)
    for (int it = 0; it < iters && e == cudaSuccess; it++) {
      cudaEventRecord(e0);
      read_sum_kernel<<<grid, 256>>>(d, n4, sink);
      cudaEventRecord(e1);
      e = cudaEventSynchronize(e1);
      float ms = 0.0f;
      if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
      if (e == cudaSuccess && ms > 0.0f) best = std::max(best, (double)(n4 * sizeof(float4)) / (ms * 1e-3) / 1e9);
    }
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(d);
  cudaFree(sink);
  if (e != cudaSuccess) return cuda_fail(e, "nrt_probe_read_gbs", __FILE__, __LINE__);
  *gb_per_s = best;
  return NRT_OK;
}

// Pinned host <-> device copy rate (GB/s, best of iters) over `bytes`: the ceiling of the host-buffer entry point
// nrt_traverse (36 B up + 17 B down per ray).  direction 0 = host -> device, 1 = device -> host.
extern "C" int nrt_probe_copy_gbs(size_t bytes, int iters, int direction, double *gb_per_s) {
  if (!gb_per_s || bytes < 4096 || iters < 1 || direction < 0 || direction > 1) {
    set_error("nrt_probe_copy_gbs: bad arguments");
    return NRT_ERR_INVALID;
  }
  DeviceGuard dg_caller;
  int rc = select_device(nullptr);
  if (rc != NRT_OK) return rc;
  void *h = nullptr, *d = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t e = cudaMallocHost(&h, bytes);
  if (e == cudaSuccess) e = cudaMalloc(&d, bytes);
  if (e == cudaSuccess) memset(h, 1, bytes);
  if (e == cudaSuccess) e = cudaEventCreate(&e0);
  if (e == cudaSuccess) e = cudaEventCreate(&e1);
  double best = 0.0;
  for (int it = 0; it <= iters && e == cudaSuccess; it++) {  // iteration 0 warms up
    cudaEventRecord(e0);
    e = direction == 0 ? cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice) : cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost);
    cudaEventRecord(e1);
    if (e == cudaSuccess) e = cudaEventSynchronize(e1);
    float ms = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    if (e == cudaSuccess && it > 0 && ms > 0.0f) best = std::max(best, (double)bytes / (ms * 1e-3) / 1e9);
  }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  cudaFree(d);
  if (h) cudaFreeHost(h);
  if (e != cudaSuccess) return cuda_fail(e, "nrt_probe_copy_gbs", __FILE__, __LINE__);
  *gb_per_s = best;
  return NRT_OK;
}
