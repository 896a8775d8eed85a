// BVHAccel<double>: the fp64 instantiation of Build / Traverse (sm_100a, --fmad=false).
//
// Replaces (file:line under /root/reference):
//   BVHAccel<double>::Build              nanort.h:1892-2149   nrt_build_f64
//   BVHAccel<double>::Traverse           nanort.h:2487-2556   traverse_f64_kernel
//   IntersectRayAABB<double>             nanort.h:2327-2370   slab_d
//   TriangleIntersector<double>::Intersect / PrepareTraversal  nanort.h:1054-1201   tri_test_d / setup_ray_d
//   vsafe_inverse<double>                nanort.h:414-465     safe_inverse_d
//
// Build: fp64 adds nothing to the SHAPE of a good tree, so the topology comes from the production builder run over the
// float-rounded vertices; what must be double is every box the traversal tests, and those are refitted exactly from the
// double vertices (leaf: min/max of its triangles; branch: union of its children, bottom-up with one atomic arrival
// counter per branch).  The result is a BVHNode<double> array in the reference's layout and order conventions.
// Traverse: one thread per ray in the reference's visiting order, all arithmetic in double with every operation
// individually rounded, so t / u / v equal the reference's bits for the reported primitive.
#include <float.h>
#include <math_constants.h>

#include <algorithm>
#include <condition_variable>
#include <cstring>
#include <new>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace nrt {

int build_reference_tree_f64_on_device(const double *d_verts, const uint32_t *d_faces, uint32_t n, uint32_t bin_size,
                                       uint32_t min_leaf_primitives, uint32_t max_tree_depth, uint32_t shallow_depth,
                                       uint32_t min_primitives_for_parallel_build, bool cpp11_order, void **d_nodes_out,
                                       uint32_t **d_indices_out, size_t *n_nodes_out, BuildStats16 *stats_out,
                                       double root_bmin[3], double root_bmax[3], cudaStream_t s);

namespace {

struct Node64 {
  double bmin[3], bmax[3];
  int32_t flag, axis;
  uint32_t data[2];
};
static_assert(sizeof(Node64) == 64, "BVHNode<double> layout");
struct Ray72 {
  double org[3], dir[3], min_t, max_t;
  uint32_t type, pad;
};
static_assert(sizeof(Ray72) == 72, "Ray<double> layout");
struct Hit32 {
  double u, v, t;
  uint32_t prim_id, pad;
};
static_assert(sizeof(Hit32) == 32, "TriangleIntersection<double> layout");
struct BuildOptions32 {
  double cost_t_aabb;
  uint32_t min_leaf_primitives, max_tree_depth, bin_size, shallow_depth, min_primitives_for_parallel_build;
  uint8_t cache_bbox, pad[3];
};
static_assert(sizeof(BuildOptions32) == 32, "BVHBuildOptions<double> layout");

struct AccelF64 {
  int device = 0;
  uint32_t n_prims = 0;
  size_t n_nodes = 0, n_verts = 0;
  Node64 *d_nodes = nullptr;
  uint32_t *d_indices = nullptr, *d_faces = nullptr;
  double *d_verts = nullptr;  // packed xyz
  BuildStats16 stats;
  double root_bmin[3], root_bmax[3];
  std::vector<Node64> h_nodes;
  std::vector<uint32_t> h_indices;
  bool mirrors_valid = false;
  cudaStream_t stream = nullptr;
  // Traverse: three staging slots, each with its own stream (copy up / traverse / copy down overlap across chunks)
  cudaStream_t tstream[3] = {nullptr, nullptr, nullptr};
  void *d_rays[3] = {nullptr, nullptr, nullptr}, *d_hits[3] = {nullptr, nullptr, nullptr}, *d_mask[3] = {nullptr, nullptr, nullptr};
  size_t stage = 0;
  // fast-path layout (f64_fast.cuh), derived on the first fast Traverse
  void *d_pair = nullptr, *d_tris_fast = nullptr;
  unsigned long long *d_cursor = nullptr;  // ring of 8 ray-pool cursors
  unsigned cursor_next = 0;
  bool fast_ready = false;
  std::mutex mu;
  // small reference-order calls (the facade's one-ray Traverse): zero-copy slots as in Accel::SmallSlot
  static constexpr int kSmallSlots = 8;
  static constexpr size_t kSmallRays = 64;
  struct SmallSlot {
    void *h = nullptr;  // rays (kSmallRays x 72 B) | records (x 32 B) | flags (x 1 B)
    cudaStream_t s = nullptr;
    bool busy = false;
  };
  SmallSlot small[kSmallSlots];
  std::mutex small_mu;
  std::condition_variable small_cv;
};

void destroy_f64(AccelF64 *a) {
  if (!a) return;
  DeviceGuard dg(a->device);
  cudaFree(a->d_nodes);
  cudaFree(a->d_indices);
  cudaFree(a->d_faces);
  cudaFree(a->d_verts);
  for (int i = 0; i < 3; i++) {
    cudaFree(a->d_rays[i]);
    cudaFree(a->d_hits[i]);
    cudaFree(a->d_mask[i]);
    if (a->tstream[i]) cudaStreamDestroy(a->tstream[i]);
  }
  for (int i = 0; i < AccelF64::kSmallSlots; i++) {
    if (a->small[i].h) cudaFreeHost(a->small[i].h);
    if (a->small[i].s) cudaStreamDestroy(a->small[i].s);
  }
  cudaFree(a->d_pair);
  cudaFree(a->d_tris_fast);
  cudaFree(a->d_cursor);
  if (a->stream) cudaStreamDestroy(a->stream);
  delete a;
}

// ---- build ---------------------------------------------------------------------------------------------------
__global__ void f64_round_kernel(const double *__restrict__ v, size_t n, float *__restrict__ out) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (float)v[i];
}

constexpr uint32_t kNoParent = 0xFFFFFFFFu;

__global__ void f64_parent_kernel(const Node40 *__restrict__ nodes, uint32_t n, uint32_t *__restrict__ parent) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i == 0) parent[0] = kNoParent;
  const Node40 nd = nodes[i];
  if (nd.flag == 0) {
    parent[nd.data[0]] = i;
    parent[nd.data[1]] = i;
  }
}

// one thread per node; leaves compute their exact double box and climb: the second child to arrive at a branch
// merges both child boxes (the first one leaves), so every branch is written once, after both children
__global__ void f64_refit_kernel(const Node40 *__restrict__ nodes, uint32_t n, const uint32_t *__restrict__ parent,
                                 const uint32_t *__restrict__ indices, const uint32_t *__restrict__ faces,
                                 const double *__restrict__ verts, uint32_t *__restrict__ arrived,
                                 Node64 *__restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Node40 nd = nodes[i];
  Node64 o;
  o.flag = nd.flag;
  o.axis = nd.axis;
  o.data[0] = nd.data[0];
  o.data[1] = nd.data[1];
  if (nd.flag == 0) {  // topology fields now, box by whichever child arrives second
    volatile int32_t *w = reinterpret_cast<volatile int32_t *>(&out[i].flag);
    w[0] = o.flag;
    w[1] = o.axis;
    w[2] = (int32_t)o.data[0];
    w[3] = (int32_t)o.data[1];
    return;
  }
  double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (uint32_t k = 0; k < nd.data[0]; k++) {
    const uint32_t prim = indices[nd.data[1] + k];
    for (int c = 0; c < 3; c++) {
      const double *p = verts + 3 * (size_t)faces[3 * (size_t)prim + c];
      for (int a = 0; a < 3; a++) {
        lo[a] = fmin(lo[a], p[a]);
        hi[a] = fmax(hi[a], p[a]);
      }
    }
  }
  for (int a = 0; a < 3; a++) {
    o.bmin[a] = lo[a];
    o.bmax[a] = hi[a];
  }
  out[i] = o;
  uint32_t cur = i;
  for (;;) {
    const uint32_t p = parent[cur];
    if (p == kNoParent) break;
    __threadfence();
    if (atomicAdd(arrived + p, 1u) == 0u) break;  // the sibling subtree is not finished yet
    __threadfence();
    const Node40 pn = nodes[p];
    const volatile double *a = reinterpret_cast<const volatile double *>(out + pn.data[0]);
    const volatile double *b = reinterpret_cast<const volatile double *>(out + pn.data[1]);
    volatile double *d = reinterpret_cast<volatile double *>(out + p);
    for (int k = 0; k < 3; k++) {
      d[k] = fmin(a[k], b[k]);
      d[3 + k] = fmax(a[3 + k], b[3 + k]);
    }
    cur = p;
  }
}

// ---- traversal -------------------------------------------------------------------------------------------------
struct RayCtxD {
  double ox, oy, oz, ix, iy, iz, Sx, Sy, Sz, t_min;
  int sx, sy, sz, kx, ky, kz;
};

__device__ __forceinline__ double safe_inverse_d(double d, bool cpp03) {
  if (fabs(d) < DBL_EPSILON) {
    const bool neg = cpp03 ? (d < 0.0) : ((unsigned long long)__double_as_longlong(d) >> 63) != 0ull;
    return neg ? -CUDART_INF : CUDART_INF;
  }
  return 1.0 / d;
}

__device__ __forceinline__ double sel3d(int k, double x, double y, double z) { return k == 0 ? x : (k == 1 ? y : z); }

__device__ __forceinline__ void setup_ray_d(RayCtxD &c, const Ray72 &r, bool cpp03) {
  const double dx = r.dir[0], dy = r.dir[1], dz = r.dir[2];
  c.ox = r.org[0], c.oy = r.org[1], c.oz = r.org[2];
  c.sx = dx < 0.0, c.sy = dy < 0.0, c.sz = dz < 0.0;
  c.ix = safe_inverse_d(dx, cpp03), c.iy = safe_inverse_d(dy, cpp03), c.iz = safe_inverse_d(dz, cpp03);
  int kz = 0;
  double m = fabs(dx);
  if (m < fabs(dy)) {
    kz = 1;
    m = fabs(dy);
  }
  if (m < fabs(dz)) kz = 2;
  int kx = (kz == 2) ? 0 : kz + 1;
  int ky = (kx == 2) ? 0 : kx + 1;
  const double dkz = sel3d(kz, dx, dy, dz);
  if (dkz < 0.0) {
    const int t = kx;
    kx = ky;
    ky = t;
  }
  c.kx = kx, c.ky = ky, c.kz = kz;
  c.Sx = sel3d(kx, dx, dy, dz) / dkz;
  c.Sy = sel3d(ky, dx, dy, dz) / dkz;
  c.Sz = 1.0 / dkz;
  c.t_min = r.min_t;
}

// safemax / safemin of the reference: the SECOND operand survives a NaN in the first, a NaN second operand wins
__device__ __forceinline__ double smax_d(double a, double b) { return (a > b) ? a : b; }
__device__ __forceinline__ double smin_d(double a, double b) { return (a < b) ? a : b; }

__device__ __forceinline__ bool slab_d(const RayCtxD &c, const Node64 *nd, double min_t, double max_t) {
  const double *f = reinterpret_cast<const double *>(nd);
  const double lox = __ldg(f + 0), loy = __ldg(f + 1), loz = __ldg(f + 2);
  const double hix = __ldg(f + 3), hiy = __ldg(f + 4), hiz = __ldg(f + 5);
  const double tnx = ((c.sx ? hix : lox) - c.ox) * c.ix;
  const double tny = ((c.sy ? hiy : loy) - c.oy) * c.iy;
  const double tnz = ((c.sz ? hiz : loz) - c.oz) * c.iz;
  const double tfx = (((c.sx ? lox : hix) - c.ox) * c.ix) * 1.0000000000000004;
  const double tfy = (((c.sy ? loy : hiy) - c.oy) * c.iy) * 1.0000000000000004;
  const double tfz = (((c.sz ? loz : hiz) - c.oz) * c.iz) * 1.0000000000000004;
  const double tmin = smax_d(tnz, smax_d(tny, smax_d(tnx, min_t)));
  const double tmax = smin_d(tfz, smin_d(tfy, smin_d(tfx, max_t)));
  return tmin <= tmax;
}

struct BestD {
  double t, u, v;
  uint32_t prim;
};

__device__ __forceinline__ bool tri_test_d(const RayCtxD &c, const TraceOptions16 &opt, const double *__restrict__ verts,
                                           const uint32_t *__restrict__ faces, uint32_t prim, BestD &best) {
  if (prim < opt.prim_ids_range[0] || prim >= opt.prim_ids_range[1]) return false;
  if (prim == opt.skip_prim_id) return false;
  const double *p0 = verts + 3 * (size_t)__ldg(faces + 3 * (size_t)prim);
  const double *p1 = verts + 3 * (size_t)__ldg(faces + 3 * (size_t)prim + 1);
  const double *p2 = verts + 3 * (size_t)__ldg(faces + 3 * (size_t)prim + 2);
  const double A0 = __ldg(p0) - c.ox, A1 = __ldg(p0 + 1) - c.oy, A2 = __ldg(p0 + 2) - c.oz;
  const double B0 = __ldg(p1) - c.ox, B1 = __ldg(p1 + 1) - c.oy, B2 = __ldg(p1 + 2) - c.oz;
  const double C0 = __ldg(p2) - c.ox, C1 = __ldg(p2 + 1) - c.oy, C2 = __ldg(p2 + 2) - c.oz;
  const double Akz = sel3d(c.kz, A0, A1, A2), Bkz = sel3d(c.kz, B0, B1, B2), Ckz = sel3d(c.kz, C0, C1, C2);
  const double Ax = sel3d(c.kx, A0, A1, A2) - c.Sx * Akz, Ay = sel3d(c.ky, A0, A1, A2) - c.Sy * Akz;
  const double Bx = sel3d(c.kx, B0, B1, B2) - c.Sx * Bkz, By = sel3d(c.ky, B0, B1, B2) - c.Sy * Bkz;
  const double Cx = sel3d(c.kx, C0, C1, C2) - c.Sx * Ckz, Cy = sel3d(c.ky, C0, C1, C2) - c.Sy * Ckz;
  // the reference's "double precision fallback" for U, V or W == 0 recomputes the very same double expressions
  const double U = Cx * By - Cy * Bx;
  const double V = Ax * Cy - Ay * Cx;
  const double W = Bx * Ay - By * Ax;
  if (U < 0.0 || V < 0.0 || W < 0.0) {
    if (opt.cull_back_face || U > 0.0 || V > 0.0 || W > 0.0) return false;
  }
  const double det = (U + V) + W;
  if (det == 0.0) return false;
  const double Az = c.Sz * Akz, Bz = c.Sz * Bkz, Cz = c.Sz * Ckz;
  const double D = (U * Az + V * Bz) + W * Cz;
  const double rcp = 1.0 / det;
  const double tt = D * rcp;
  if (tt > best.t) return false;
  if (tt < c.t_min) return false;
  best.t = tt;
  best.u = V * rcp;
  best.v = W * rcp;
  best.prim = prim;
  return true;
}

constexpr int kStackD = 512;  // kNANORT_MAX_STACK_DEPTH

__global__ void __launch_bounds__(128)
    traverse_f64_kernel(const Node64 *__restrict__ nodes, const uint32_t *__restrict__ indices,
                        const uint32_t *__restrict__ faces, const double *__restrict__ verts,
                        const Ray72 *__restrict__ rays, size_t n, Hit32 *__restrict__ hits, uint8_t *__restrict__ mask,
                        TraceOptions16 opt, uint32_t flags) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Ray72 r = rays[i];
  RayCtxD c;
  setup_ray_d(c, r, (flags & NRT_TRAVERSE_CPP03_INVERSE) != 0);
  BestD best;
  best.t = r.max_t;
  best.u = 0.0;
  best.v = 0.0;
  best.prim = 0xFFFFFFFFu;
  double hit_t = r.max_t;
  uint32_t stack[kStackD];
  int sp = 0;
  stack[0] = 0;
  while (sp >= 0) {
    const Node64 *nd = nodes + stack[sp];
    sp--;
    if (!slab_d(c, nd, r.min_t, hit_t)) continue;
    const uint32_t d0 = __ldg(&nd->data[0]), d1 = __ldg(&nd->data[1]);
    if (__ldg(&nd->flag) == 0) {
      const int axis = __ldg(&nd->axis);
      const int sgn = axis == 0 ? c.sx : (axis == 1 ? c.sy : c.sz);
      if (sp + 2 < kStackD) {
        stack[++sp] = sgn ? d0 : d1;
        stack[++sp] = sgn ? d1 : d0;
      }
    } else {
      bool any = false;
      for (uint32_t k = 0; k < d0; k++)
        if (tri_test_d(c, opt, verts, faces, __ldg(indices + d1 + k), best)) any = true;
      if (any) hit_t = best.t;
    }
  }
  const bool hit = best.t < r.max_t;
  Hit32 h;
  h.u = hit ? best.u : 0.0;
  h.v = hit ? best.v : 0.0;
  h.t = hit ? best.t : r.max_t;
  h.prim_id = hit ? best.prim : 0xFFFFFFFFu;
  h.pad = 0;
  hits[i] = h;
  if (mask) mask[i] = hit ? 1 : 0;
}

#include "f64_fast.cuh"

// PairNodeD / TriD arrays of the fast kernel from the BVHNode<double> array, indices_, faces and double vertices the
// accel already holds on the device.  Called under a->mu.
int derive_fast_layout_f64(AccelF64 *a) {
  if (a->fast_ready) return NRT_OK;
  const uint32_t nn = (uint32_t)a->n_nodes;
  uint32_t *d_flags = nullptr, *d_widx = nullptr;
  uint32_t n_branch = 0;
  int rc = NRT_OK;
  cudaStream_t s = a->stream;
  cudaError_t e = cudaMalloc(&d_flags, sizeof(uint32_t) * (size_t)nn);
  if (e == cudaSuccess) e = cudaMalloc(&d_widx, sizeof(uint32_t) * (size_t)nn);
  if (e == cudaSuccess) {
    f64_branch_flags_kernel<<<(nn + 255) / 256, 256, 0, s>>>(a->d_nodes, nn, d_flags);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) rc = exclusive_scan_u32(d_flags, d_widx, nn, &n_branch, s);
  const size_t n_pair = n_branch > 0 ? n_branch : 1;
  if (e == cudaSuccess && rc == NRT_OK) e = cudaMalloc(&a->d_pair, sizeof(PairNodeD) * n_pair);
  if (e == cudaSuccess && rc == NRT_OK) e = cudaMalloc(&a->d_tris_fast, sizeof(TriD) * (size_t)a->n_prims);
  if (e == cudaSuccess && rc == NRT_OK && !a->d_cursor) e = cudaMalloc(&a->d_cursor, sizeof(unsigned long long) * 8);
  if (e == cudaSuccess && rc == NRT_OK) {
    f64_tris_kernel<<<(a->n_prims + 255) / 256, 256, 0, s>>>(a->d_indices, a->d_faces, a->d_verts, a->n_prims,
                                                            static_cast<TriD *>(a->d_tris_fast));
    f64_pair_kernel<<<(nn + 255) / 256, 256, 0, s>>>(a->d_nodes, nn, d_widx, static_cast<PairNodeD *>(a->d_pair),
                                                    static_cast<TriD *>(a->d_tris_fast));
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d_flags);
  cudaFree(d_widx);
  if (e != cudaSuccess || rc != NRT_OK) {
    cudaFree(a->d_pair);
    cudaFree(a->d_tris_fast);
    a->d_pair = a->d_tris_fast = nullptr;
    return e != cudaSuccess ? cuda_fail(e, "derive_fast_layout_f64", __FILE__, __LINE__) : rc;
  }
  a->fast_ready = true;
  return NRT_OK;
}

template <int DEPTH>
cudaError_t launch_fast_f64(AccelF64 *a, const Ray72 *d_rays, size_t m, Hit32 *d_hits, uint8_t *d_mask,
                            const TraceOptions16 &opt, uint32_t flags, cudaStream_t s) {
  unsigned long long *cursor = a->d_cursor + (a->cursor_next++ & 7u);
  cudaError_t e = cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), s);
  if (e != cudaSuccess) return e;
  size_t grid = (size_t)device_sm_count(a->device) * kFastBlocksPerSmD;  // persistent: every SM holds its complement
  const size_t need = ((m + 31) / 32 + kFastBlockD / 32 - 1) / (kFastBlockD / 32);
  if (grid > need) grid = need;
  if (grid == 0) grid = 1;
  traverse_fast_f64_kernel<DEPTH><<<(unsigned)grid, kFastBlockD, 0, s>>>(
      static_cast<const PairNodeD *>(a->d_pair), static_cast<const TriD *>(a->d_tris_fast), d_rays, m, d_hits, d_mask, opt,
      flags, cursor);
  return cudaGetLastError();
}

}  // namespace
}  // namespace nrt

using namespace nrt;

#define F64_CUDA(expr)                                           \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) {                                     \
      rc = cuda_fail(_e, #expr, __FILE__, __LINE__);             \
      goto fail;                                                 \
    }                                                            \
  } while (0)

extern "C" {

int nrt_build_f64(const double *verts, size_t stride_bytes, size_t n_verts, const uint32_t *faces, uint32_t n_prims,
                  const void *build_opts_32B, nrt_accel_f64 **out) {
  return nrt_build_f64_ex(verts, stride_bytes, n_verts, faces, n_prims, build_opts_32B, NRT_BUILD_FAST, out);
}

int nrt_build_f64_ex(const double *verts, size_t stride_bytes, size_t n_verts, const uint32_t *faces, uint32_t n_prims,
                     const void *build_opts_32B, uint32_t flags, nrt_accel_f64 **out) {
  if (!out) {
    set_error("nrt_build_f64: out is NULL");
    return NRT_ERR_INVALID;
  }
  *out = nullptr;
  if (n_prims == 0) {  // Build returns false (nanort.h:1907-1909)
    set_error("nrt_build_f64: num_primitives == 0");
    return NRT_ERR_INVALID;
  }
  if (!verts || !faces || stride_bytes < 24) {
    set_error("nrt_build_f64: bad geometry pointers / stride");
    return NRT_ERR_INVALID;
  }
  BuildOptions32 o64;
  {
    const BuildOptions28 d = default_build_options();
    o64.cost_t_aabb = d.cost_t_aabb;
    o64.min_leaf_primitives = d.min_leaf_primitives;
    o64.max_tree_depth = d.max_tree_depth;
    o64.bin_size = d.bin_size;
    o64.shallow_depth = d.shallow_depth;
    o64.min_primitives_for_parallel_build = d.min_primitives_for_parallel_build;
    o64.cache_bbox = 0;
    o64.pad[0] = o64.pad[1] = o64.pad[2] = 0;
  }
  if (build_opts_32B) memcpy(&o64, build_opts_32B, sizeof(o64));
  if (o64.bin_size < 2 || o64.max_tree_depth > 500) {
    set_error("nrt_build_f64: bin_size must be > 1 and max_tree_depth <= 500");
    return NRT_ERR_INVALID;
  }
  int device = 0;
  DeviceGuard dg_caller;  // select_device makes the chosen device current; the caller gets its own back
  int rc = select_device(&device);
  if (rc != NRT_OK) return rc;
  if (n_verts == 0) {
    uint32_t m = 0;
    for (size_t k = 0; k < (size_t)n_prims * 3; k++) m = std::max(m, faces[k]);
    n_verts = (size_t)m + 1;
  }
  AccelF64 *a = new (std::nothrow) AccelF64();
  Accel *t = new (std::nothrow) Accel();  // float topology, discarded after the refit
  uint32_t *d_parent = nullptr, *d_arrived = nullptr;
  if (!a || !t) {
    delete a;
    delete t;
    return NRT_ERR_NOMEM;
  }
  a->device = t->device = device;
  a->n_prims = t->n_prims = n_prims;
  a->n_verts = t->n_verts = n_verts;
  {
    std::vector<double> packed(3 * n_verts);
    for (size_t i = 0; i < n_verts; i++) {
      const double *p = reinterpret_cast<const double *>(reinterpret_cast<const char *>(verts) + i * stride_bytes);
      packed[3 * i] = p[0], packed[3 * i + 1] = p[1], packed[3 * i + 2] = p[2];
    }
    F64_CUDA(cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking));
    F64_CUDA(cudaMalloc(&a->d_verts, sizeof(double) * 3 * n_verts));
    F64_CUDA(cudaMalloc(&a->d_faces, sizeof(uint32_t) * 3 * (size_t)n_prims));
    F64_CUDA(cudaMalloc(&t->d_verts, sizeof(float) * 3 * n_verts));
    // stream-ordered with the kernels that consume them (a->stream is non-blocking, see api.cu:upload_geometry)
    F64_CUDA(cudaMemcpyAsync(a->d_verts, packed.data(), sizeof(double) * 3 * n_verts, cudaMemcpyHostToDevice, a->stream));
    F64_CUDA(cudaMemcpyAsync(a->d_faces, faces, sizeof(uint32_t) * 3 * (size_t)n_prims, cudaMemcpyHostToDevice, a->stream));
    F64_CUDA(cudaStreamSynchronize(a->stream));
  }
  if (flags & NRT_BUILD_REFERENCE_TREE) {
    // conformance build: the reference's own BVHNode<double> array and indices_, bit for bit (build_ref64.cu)
    void *d_nodes = nullptr;
    cudaFree(t->d_verts);
    t->d_verts = nullptr;
    rc = build_reference_tree_f64_on_device(a->d_verts, a->d_faces, n_prims, o64.bin_size, o64.min_leaf_primitives,
                                            o64.max_tree_depth, o64.shallow_depth, o64.min_primitives_for_parallel_build,
                                            (flags & NRT_BUILD_REFERENCE_CPP03_ORDER) == 0, &d_nodes, &a->d_indices,
                                            &a->n_nodes, &a->stats, a->root_bmin, a->root_bmax, a->stream);
    if (rc != NRT_OK) goto fail;
    a->d_nodes = static_cast<Node64 *>(d_nodes);
    delete t;
    *out = reinterpret_cast<nrt_accel_f64 *>(a);
    return NRT_OK;
  }
  t->d_faces = a->d_faces;  // shared, freed with `a`
  t->options = default_build_options();
  t->options.cost_t_aabb = (float)o64.cost_t_aabb;
  t->options.min_leaf_primitives = o64.min_leaf_primitives;
  t->options.max_tree_depth = o64.max_tree_depth;
  t->options.bin_size = o64.bin_size;
  t->options.shallow_depth = o64.shallow_depth;
  t->options.min_primitives_for_parallel_build = o64.min_primitives_for_parallel_build;
  f64_round_kernel<<<(unsigned)((3 * n_verts + 255) / 256), 256, 0, a->stream>>>(a->d_verts, 3 * n_verts, t->d_verts);
  F64_CUDA(cudaGetLastError());
  rc = build_on_device(t, a->stream);
  if (rc != NRT_OK) goto fail;
  {
    const uint32_t nn = (uint32_t)t->n_nodes;
    a->n_nodes = nn;
    F64_CUDA(cudaMalloc(&a->d_nodes, sizeof(Node64) * (size_t)nn));
    F64_CUDA(cudaMalloc(&d_parent, sizeof(uint32_t) * (size_t)nn));
    F64_CUDA(cudaMalloc(&d_arrived, sizeof(uint32_t) * (size_t)nn));
    F64_CUDA(cudaMemsetAsync(d_arrived, 0, sizeof(uint32_t) * (size_t)nn, a->stream));
    f64_parent_kernel<<<(nn + 255) / 256, 256, 0, a->stream>>>(t->d_nodes, nn, d_parent);
    f64_refit_kernel<<<(nn + 127) / 128, 128, 0, a->stream>>>(t->d_nodes, nn, d_parent, t->d_indices, a->d_faces,
                                                              a->d_verts, d_arrived, a->d_nodes);
    F64_CUDA(cudaGetLastError());
    Node64 root;
    F64_CUDA(cudaMemcpyAsync(&root, a->d_nodes, sizeof(Node64), cudaMemcpyDeviceToHost, a->stream));
    F64_CUDA(cudaStreamSynchronize(a->stream));
    for (int k = 0; k < 3; k++) {
      a->root_bmin[k] = root.bmin[k];
      a->root_bmax[k] = root.bmax[k];
    }
  }
  a->d_indices = t->d_indices;
  t->d_indices = nullptr;
  a->stats = t->stats;
  cudaFree(d_parent);
  cudaFree(d_arrived);
  cudaFree(t->d_nodes);
  cudaFree(t->d_verts);
  delete t;
  *out = reinterpret_cast<nrt_accel_f64 *>(a);
  return NRT_OK;
fail:
  cudaFree(d_parent);
  cudaFree(d_arrived);
  cudaFree(t->d_nodes);
  cudaFree(t->d_indices);
  cudaFree(t->d_verts);
  delete t;
  destroy_f64(a);
  return rc;
}

int nrt_adopt_f64(const void *nodes_64B, size_t n_nodes, const uint32_t *indices, size_t n_indices, const double *verts,
                  size_t stride_bytes, size_t n_verts, const uint32_t *faces, uint32_t n_prims, nrt_accel_f64 **out) {
  if (!out) {
    set_error("nrt_adopt_f64: out is NULL");
    return NRT_ERR_INVALID;
  }
  *out = nullptr;
  if (!nodes_64B || !indices || !verts || !faces || n_nodes == 0 || n_prims == 0 || n_indices != n_prims ||
      stride_bytes < 24) {
    set_error("nrt_adopt_f64: bad arguments");
    return NRT_ERR_INVALID;
  }
  const Node64 *hn = static_cast<const Node64 *>(nodes_64B);
  BuildStats16 st = {0, 0, 0, 0.0f};
  {  // an adopted tree is foreign data: the same structure check as nrt_adopt (api.cu:validate_foreign_tree)
    std::string why;
    if (!validate_foreign_tree64(nodes_64B, n_nodes, indices, n_indices, n_prims, &st, &why)) {
      set_error("nrt_adopt_f64: " + why);
      return NRT_ERR_INVALID;
    }
  }
  int device = 0;
  DeviceGuard dg_caller;  // select_device makes the chosen device current; the caller gets its own back
  int rc = select_device(&device);
  if (rc != NRT_OK) return rc;
  if (n_verts == 0) {
    uint32_t m = 0;
    for (size_t k = 0; k < (size_t)n_prims * 3; k++) m = std::max(m, faces[k]);
    n_verts = (size_t)m + 1;
  }
  AccelF64 *a = new (std::nothrow) AccelF64();
  if (!a) return NRT_ERR_NOMEM;
  a->device = device;
  a->n_prims = n_prims;
  a->n_verts = n_verts;
  a->n_nodes = n_nodes;
  a->stats = st;
  std::vector<double> packed(3 * n_verts);
  for (size_t i = 0; i < n_verts; i++) {
    const double *p = reinterpret_cast<const double *>(reinterpret_cast<const char *>(verts) + i * stride_bytes);
    packed[3 * i] = p[0], packed[3 * i + 1] = p[1], packed[3 * i + 2] = p[2];
  }
  cudaError_t e = cudaStreamCreateWithFlags(&a->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&a->d_verts, sizeof(double) * 3 * n_verts);
  if (e == cudaSuccess) e = cudaMalloc(&a->d_faces, sizeof(uint32_t) * 3 * (size_t)n_prims);
  if (e == cudaSuccess) e = cudaMalloc(&a->d_nodes, sizeof(Node64) * n_nodes);
  if (e == cudaSuccess) e = cudaMalloc(&a->d_indices, sizeof(uint32_t) * n_indices);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(a->d_verts, packed.data(), sizeof(double) * 3 * n_verts, cudaMemcpyHostToDevice, a->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(a->d_faces, faces, sizeof(uint32_t) * 3 * (size_t)n_prims, cudaMemcpyHostToDevice, a->stream);
  if (e == cudaSuccess) e = cudaMemcpyAsync(a->d_nodes, hn, sizeof(Node64) * n_nodes, cudaMemcpyHostToDevice, a->stream);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(a->d_indices, indices, sizeof(uint32_t) * n_indices, cudaMemcpyHostToDevice, a->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(a->stream);
  if (e != cudaSuccess) {
    rc = cuda_fail(e, "nrt_adopt_f64 upload", __FILE__, __LINE__);
    destroy_f64(a);
    return rc;
  }
  for (int k = 0; k < 3; k++) {
    a->root_bmin[k] = hn[0].bmin[k];
    a->root_bmax[k] = hn[0].bmax[k];
  }
  a->h_nodes.assign(hn, hn + n_nodes);
  a->h_indices.assign(indices, indices + n_indices);
  a->mirrors_valid = true;
  *out = reinterpret_cast<nrt_accel_f64 *>(a);
  return NRT_OK;
}

void nrt_free_f64(nrt_accel_f64 *a) { destroy_f64(reinterpret_cast<AccelF64 *>(a)); }

int nrt_stats_f64(const nrt_accel_f64 *h, void *stats_16B) {
  if (!h || !stats_16B) {
    set_error("nrt_stats_f64: NULL argument");
    return NRT_ERR_INVALID;
  }
  memcpy(stats_16B, &reinterpret_cast<const AccelF64 *>(h)->stats, sizeof(BuildStats16));
  return NRT_OK;
}

int nrt_bounding_box_f64(const nrt_accel_f64 *h, double bmin[3], double bmax[3]) {
  if (!h || !bmin || !bmax) {
    set_error("nrt_bounding_box_f64: NULL argument");
    return NRT_ERR_INVALID;
  }
  const AccelF64 *a = reinterpret_cast<const AccelF64 *>(h);
  for (int k = 0; k < 3; k++) {
    bmin[k] = a->root_bmin[k];
    bmax[k] = a->root_bmax[k];
  }
  return NRT_OK;
}

int nrt_nodes_f64(nrt_accel_f64 *h, const void **nodes_64B, size_t *n_nodes, const uint32_t **indices,
                  size_t *n_indices) {
  if (!h) {
    set_error("nrt_nodes_f64: NULL accel");
    return NRT_ERR_INVALID;
  }
  AccelF64 *a = reinterpret_cast<AccelF64 *>(h);
  std::lock_guard<std::mutex> lock(a->mu);
  if (!a->mirrors_valid) {
    NRT_DEVICE(a->device);
    a->h_nodes.resize(a->n_nodes);
    a->h_indices.resize(a->n_prims);
    NRT_CUDA(cudaMemcpy(a->h_nodes.data(), a->d_nodes, sizeof(Node64) * a->n_nodes, cudaMemcpyDeviceToHost));
    NRT_CUDA(cudaMemcpy(a->h_indices.data(), a->d_indices, sizeof(uint32_t) * a->n_prims, cudaMemcpyDeviceToHost));
    a->mirrors_valid = true;
  }
  if (nodes_64B) *nodes_64B = a->h_nodes.data();
  if (n_nodes) *n_nodes = a->h_nodes.size();
  if (indices) *indices = a->h_indices.data();
  if (n_indices) *n_indices = a->h_indices.size();
  return NRT_OK;
}

int nrt_traverse_f64(const nrt_accel_f64 *h, const void *rays_72B, size_t n_rays, void *hits_32B, uint8_t *hit_mask,
                     const void *trace_opts_16B, uint32_t flags) {
  if (!h || (n_rays && (!rays_72B || !hits_32B))) {
    set_error("nrt_traverse_f64: NULL argument");
    return NRT_ERR_INVALID;
  }
  if (n_rays == 0) return NRT_OK;
  AccelF64 *a = const_cast<AccelF64 *>(reinterpret_cast<const AccelF64 *>(h));
  TraceOptions16 opt = default_trace_options();
  if (trace_opts_16B) memcpy(&opt, trace_opts_16B, sizeof(opt));
  if (n_rays <= AccelF64::kSmallRays && (flags & NRT_TRAVERSE_CONFORMANCE)) {
    // low-latency path of the facade's per-ray Traverse: the kernel reads the rays from and writes the records to a
    // pinned host slot (zero copy), one synchronisation, host threads side by side
    NRT_DEVICE(a->device);
    const size_t off_hits = AccelF64::kSmallRays * sizeof(Ray72), off_mask = off_hits + AccelF64::kSmallRays * sizeof(Hit32);
    int idx = -1;
    {
      std::unique_lock<std::mutex> lk(a->small_mu);
      for (;;) {
        for (int i = 0; i < AccelF64::kSmallSlots && idx < 0; i++)
          if (!a->small[i].busy) idx = i;
        if (idx >= 0) break;
        a->small_cv.wait(lk);
      }
      a->small[idx].busy = true;
    }
    AccelF64::SmallSlot &sl = a->small[idx];
    cudaError_t e = cudaSuccess;
    if (!sl.h) e = cudaHostAlloc(&sl.h, off_mask + AccelF64::kSmallRays, cudaHostAllocPortable | cudaHostAllocMapped);
    if (e == cudaSuccess && !sl.s) e = cudaStreamCreateWithFlags(&sl.s, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
      char *hb = static_cast<char *>(sl.h);
      memcpy(hb, rays_72B, n_rays * sizeof(Ray72));
      traverse_f64_kernel<<<1, 64, 0, sl.s>>>(a->d_nodes, a->d_indices, a->d_faces, a->d_verts,
                                             reinterpret_cast<const Ray72 *>(hb), n_rays, reinterpret_cast<Hit32 *>(hb + off_hits),
                                             hit_mask ? reinterpret_cast<uint8_t *>(hb + off_mask) : nullptr, opt, flags);
      e = cudaGetLastError();
      const cudaError_t es = cudaStreamSynchronize(sl.s);
      if (e == cudaSuccess) e = es;
      if (e == cudaSuccess) {
        memcpy(hits_32B, hb + off_hits, n_rays * sizeof(Hit32));
        if (hit_mask) memcpy(hit_mask, hb + off_mask, n_rays);
      }
    }
    {
      std::lock_guard<std::mutex> lk(a->small_mu);
      sl.busy = false;
    }
    a->small_cv.notify_one();
    NRT_CUDA(e);
    return NRT_OK;
  }
  std::lock_guard<std::mutex> lock(a->mu);  // Traverse is const and thread-safe in the reference; staging is shared
  NRT_DEVICE(a->device);
  const bool fast = (flags & NRT_TRAVERSE_CONFORMANCE) == 0;
  if (fast) {
    const int rc = derive_fast_layout_f64(a);
    if (rc != NRT_OK) return rc;
  }
  const size_t chunk = std::min(n_rays, (size_t)1 << 20);  // 1 Mi rays = 72 MiB up, 33 MiB down per chunk
  if (a->stage < chunk) {
    for (int i = 0; i < 3; i++) {
      cudaFree(a->d_rays[i]);
      cudaFree(a->d_hits[i]);
      cudaFree(a->d_mask[i]);
      a->d_rays[i] = a->d_hits[i] = a->d_mask[i] = nullptr;
    }
    a->stage = 0;
    for (int i = 0; i < 3; i++) {
      if (!a->tstream[i]) NRT_CUDA(cudaStreamCreateWithFlags(&a->tstream[i], cudaStreamNonBlocking));
      NRT_CUDA(cudaMalloc(&a->d_rays[i], chunk * sizeof(Ray72)));
      NRT_CUDA(cudaMalloc(&a->d_hits[i], chunk * sizeof(Hit32)));
      NRT_CUDA(cudaMalloc(&a->d_mask[i], chunk));
    }
    a->stage = chunk;
  }
  const char *src = static_cast<const char *>(rays_72B);
  char *dst = static_cast<char *>(hits_32B);
  const bool deep = a->stats.max_tree_depth + 2 > 64u;  // adopted reference trees reach depth 256
  cudaError_t e = cudaSuccess;
  int slot = 0;
  for (size_t done = 0; done < n_rays && e == cudaSuccess; done += chunk) {
    const size_t m = std::min(chunk, n_rays - done);
    cudaStream_t s = a->tstream[slot];
    e = cudaStreamSynchronize(s);  // the slot's previous chunk (3 iterations ago) has drained
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(a->d_rays[slot], src + done * sizeof(Ray72), m * sizeof(Ray72), cudaMemcpyHostToDevice, s);
    if (e != cudaSuccess) break;
    const Ray72 *d_r = static_cast<const Ray72 *>(a->d_rays[slot]);
    Hit32 *d_h = static_cast<Hit32 *>(a->d_hits[slot]);
    uint8_t *d_m = static_cast<uint8_t *>(a->d_mask[slot]);
    if (fast) {
      e = deep ? launch_fast_f64<512>(a, d_r, m, d_h, d_m, opt, flags, s) : launch_fast_f64<64>(a, d_r, m, d_h, d_m, opt, flags, s);
    } else {
      traverse_f64_kernel<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(a->d_nodes, a->d_indices, a->d_faces, a->d_verts, d_r,
                                                                      m, d_h, d_m, opt, flags);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(dst + done * sizeof(Hit32), d_h, m * sizeof(Hit32), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess && hit_mask) e = cudaMemcpyAsync(hit_mask + done, d_m, m, cudaMemcpyDeviceToHost, s);
    slot = (slot + 1) % 3;
  }
  // success or not, nothing may still be writing into the caller's buffers when this call returns
  for (int i = 0; i < 3; i++) {
    const cudaError_t es = a->tstream[i] ? cudaStreamSynchronize(a->tstream[i]) : cudaSuccess;
    if (e == cudaSuccess) e = es;
  }
  NRT_CUDA(e);
  return NRT_OK;
}

int nrt_traverse_f64_device(const nrt_accel_f64 *h, const void *d_rays_72B, size_t n_rays, void *d_hits_32B,
                            uint8_t *d_hit_mask, const void *trace_opts_16B, uint32_t flags, void *stream) {
  if (!h || (n_rays && (!d_rays_72B || !d_hits_32B))) {
    set_error("nrt_traverse_f64_device: NULL argument");
    return NRT_ERR_INVALID;
  }
  if (n_rays == 0) return NRT_OK;
  AccelF64 *a = const_cast<AccelF64 *>(reinterpret_cast<const AccelF64 *>(h));
  TraceOptions16 opt = default_trace_options();
  if (trace_opts_16B) memcpy(&opt, trace_opts_16B, sizeof(opt));
  std::lock_guard<std::mutex> lock(a->mu);  // lazy layout + the cursor ring
  NRT_DEVICE(a->device);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const Ray72 *d_r = static_cast<const Ray72 *>(d_rays_72B);
  Hit32 *d_h = static_cast<Hit32 *>(d_hits_32B);
  cudaError_t e;
  if (flags & NRT_TRAVERSE_CONFORMANCE) {
    traverse_f64_kernel<<<(unsigned)((n_rays + 127) / 128), 128, 0, s>>>(a->d_nodes, a->d_indices, a->d_faces, a->d_verts, d_r,
                                                                         n_rays, d_h, d_hit_mask, opt, flags);
    e = cudaGetLastError();
  } else {
    const int rc = derive_fast_layout_f64(a);
    if (rc != NRT_OK) return rc;
    e = a->stats.max_tree_depth + 2 > 64u ? launch_fast_f64<512>(a, d_r, n_rays, d_h, d_hit_mask, opt, flags, s)
                                          : launch_fast_f64<64>(a, d_r, n_rays, d_h, d_hit_mask, opt, flags, s);
  }
  NRT_CUDA(e);
  return NRT_OK;
}

}  // extern "C"
