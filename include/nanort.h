// nanort.h -- drop-in facade of lighttransport/nanort's header API whose BVH build and ray traversal run
// on an NVIDIA B200 through the C-ABI of include/nanort_b200.h (libnanort_b200.so).
//
// This file is NOT the reference header and shares no code with it.  It re-declares, with the same
// names, member order and defaults, the part of the API surface that the hot path needs (SURVEY.md
// section 8b; file:line below are /root/reference/nanort.h):
//
//   nanort::Ray<float>                       :474-496   36 bytes
//   nanort::BVHNode<float>                   :498-550   40 bytes
//   nanort::BVHBuildOptions<float>           :559-583   28 bytes
//   nanort::BVHBuildStatistics               :586-599
//   nanort::BVHTraceOptions                  :604-624   16 bytes
//   nanort::TriangleSAHPred<float>           :863-919
//   nanort::TriangleMesh<float>              :922-991
//   nanort::TriangleIntersection<float>      :996-1005  16 bytes
//   nanort::TriangleIntersector<float, H>    :1014-1229 (constructors + geometry accessors)
//   nanort::BVHAccel<float>                  :698-860   Build / Traverse / GetStatistics / GetNodes /
//                                                        GetIndices / BoundingBox / IsValid / Dump / Load
//
// so that a caller written against nanort -- e.g. the loop of examples/path_tracer/main.cc:742-763 and
// :839-854 -- compiles unchanged and links against -lnanort_b200.  Extension: BVHAccel::TraverseBatch, the
// batched form of Traverse that a wavefront renderer should use (one call per bounce, not per ray).
//
// float and double with the built-in triangle classes are supported (custom Prim/Pred/Intersector models would
// need device-side user code); BVHAccel<double> is at the end of this file.  There is no CPU fallback: without a CUDA device Build
// returns false and Traverse reports a miss after printing nrt_last_error() to stderr.
#ifndef NANORT_H_
#define NANORT_H_

#include <stddef.h>
#include <stdint.h>

// the reference header's own include set (nanort.h:36-49, 67-73): callers such as examples/path_tracer/main.cc rely on
// it transitively (std::atomic / std::thread under NANORT_USE_CPP11_FEATURE, assert, std::string ...)
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <type_traits>
#include <vector>

// constants of the reference (nanort.h:62-65)
#define kNANORT_MAX_STACK_DEPTH (512)
#define kNANORT_MIN_PRIMITIVES_FOR_PARALLEL_BUILD (1024 * 8)
#define kNANORT_SHALLOW_DEPTH (4)

#ifdef NANORT_USE_CPP11_FEATURE
#include <atomic>
#include <thread>
#define kNANORT_MAX_THREADS (256)
#ifndef NANORT_ENABLE_PARALLEL_BUILD
#define NANORT_ENABLE_PARALLEL_BUILD
#endif
#endif

#include "nanort_b200.h"

// The reference's own switch keeps its meaning (nanort.h:51-82, 418-462): with NANORT_USE_CPP11_FEATURE
// vsafe_inverse treats -0.0f as negative (copysign) and, in conformance mode, large scenes get the parallel build's
// node order; without it the C++03 conventions apply.
#ifdef NANORT_USE_CPP11_FEATURE
#define NANORT_B200_INVERSE_FLAG 0u
#define NANORT_B200_ORDER_FLAG 0u
#else
#define NANORT_B200_INVERSE_FLAG NRT_TRAVERSE_CPP03_INVERSE
#define NANORT_B200_ORDER_FLAG NRT_BUILD_REFERENCE_CPP03_ORDER
#endif
// -DNANORT_B200_CONFORMANCE: Build writes exactly the arrays CPU nanort would (conformance builder) and Traverse
// walks them in the reference's visiting order -- bit-identical results, ties included.  Default: the fast path.
#ifdef NANORT_B200_CONFORMANCE
#define NANORT_B200_BUILD_FLAGS (NRT_BUILD_REFERENCE_TREE | NANORT_B200_ORDER_FLAG)
#define NANORT_B200_TRAVERSE_FLAGS (NRT_TRAVERSE_CONFORMANCE | NANORT_B200_INVERSE_FLAG)
#else
#define NANORT_B200_BUILD_FLAGS NRT_BUILD_FAST
#define NANORT_B200_TRAVERSE_FLAGS (NRT_TRAVERSE_FAST | NANORT_B200_INVERSE_FLAG)
#endif

namespace nanort {

// ---- ray type flags (carried in Ray::type, never read by the kernel) -- :86-94
typedef enum {
  RAY_TYPE_NONE = 0x0,
  RAY_TYPE_PRIMARY = 0x1,
  RAY_TYPE_SECONDARY = 0x2,
  RAY_TYPE_DIFFUSE = 0x4,
  RAY_TYPE_REFLECTION = 0x8,
  RAY_TYPE_REFRACTION = 0x10
} RayType;

// ---- small 3-vector + helpers of the public namespace -- :321-412
template <typename T = float>
class real3 {
 public:
  real3() { v[0] = v[1] = v[2] = T(0); }
  real3(T s) { v[0] = v[1] = v[2] = s; }  // implicit, as nanort.h:325
  real3(T x, T y, T z) {
    v[0] = x;
    v[1] = y;
    v[2] = z;
  }
  explicit real3(const T *p) {
    v[0] = p[0];
    v[1] = p[1];
    v[2] = p[2];
  }
  T x() const { return v[0]; }
  T y() const { return v[1]; }
  T z() const { return v[2]; }
  T operator[](int i) const { return v[i]; }
  T &operator[](int i) { return v[i]; }
  real3 operator+(const real3 &o) const { return real3(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  real3 operator-(const real3 &o) const { return real3(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
  real3 operator*(const real3 &o) const { return real3(v[0] * o.v[0], v[1] * o.v[1], v[2] * o.v[2]); }
  real3 operator/(const real3 &o) const { return real3(v[0] / o.v[0], v[1] / o.v[1], v[2] / o.v[2]); }
  real3 operator*(T s) const { return real3(v[0] * s, v[1] * s, v[2] * s); }
  real3 operator-() const { return real3(-v[0], -v[1], -v[2]); }
  real3 &operator+=(const real3 &o) {
    v[0] += o.v[0];
    v[1] += o.v[1];
    v[2] += o.v[2];
    return *this;
  }
  T v[3];
};
template <typename T>
inline real3<T> operator*(T s, const real3<T> &a) {
  return a * s;
}
template <typename T>
inline T vdot(const real3<T> &a, const real3<T> &b) {
  return a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
}
template <typename T>
inline real3<T> vcross(const real3<T> &a, const real3<T> &b) {
  return real3<T>(a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]);
}
template <typename T>
inline T vlength(const real3<T> &a) {
  return std::sqrt(vdot(a, a));
}
template <typename T>
inline real3<T> vneg(const real3<T> &a) {  // nanort.h:377-380
  return real3<T>(-a[0], -a[1], -a[2]);
}
template <typename T>
inline real3<T> vnormalize(const real3<T> &a) {  // nanort.h:387-398: unchanged below epsilon, else scaled by 1 / length
  real3<T> r = a;
  const T l = vlength(a);
  if (std::fabs(l) > std::numeric_limits<T>::epsilon()) {
    const T inv = T(1) / l;
    r[0] *= inv;
    r[1] *= inv;
    r[2] *= inv;
  }
  return r;
}
// nanort.h:1235-1243: the reference's public min / max helpers (its slab test's NaN rule: a NaN FIRST operand loses);
// callers such as examples/nanosg/nanosg.h:624-628 use them in their own intersectors
template <class T>
inline const T &safemin(const T &a, const T &b) {
  return (a < b) ? a : b;
}
template <class T>
inline const T &safemax(const T &a, const T &b) {
  return (a > b) ? a : b;
}
// nanort.h:414-465: 1 / v with +-inf for |v| < epsilon; the sign convention follows NANORT_USE_CPP11_FEATURE
template <typename T>
inline real3<T> vsafe_inverse(const real3<T> v) {
  real3<T> r;
  for (int k = 0; k < 3; k++) {
    if (std::fabs(v[k]) < std::numeric_limits<T>::epsilon()) {
#ifdef NANORT_USE_CPP11_FEATURE
      r[k] = std::numeric_limits<T>::infinity() * std::copysign(T(1), v[k]);
#else
      r[k] = std::numeric_limits<T>::infinity() * (v[k] < T(0) ? T(-1) : T(1));
#endif
    } else {
      r[k] = T(1) / v[k];
    }
  }
  return r;
}
template <typename real>
inline const real *get_vertex_addr(const real *p, const size_t idx, const size_t stride_bytes) {  // nanort.h:467-472
  return reinterpret_cast<const real *>(reinterpret_cast<const unsigned char *>(p) + idx * stride_bytes);
}

// ---- value types ----------------------------------------------------------------------------
template <typename T = float>
class Ray {
 public:
  Ray() : min_t(T(0)), max_t(std::numeric_limits<T>::max()), type(RAY_TYPE_NONE) {
    org[0] = org[1] = org[2] = T(0);
    dir[0] = dir[1] = T(0);
    dir[2] = T(-1);
  }
  T org[3];
  T dir[3];
  T min_t;
  T max_t;
  unsigned int type;
};

template <typename T = float>
class BVHNode {
 public:
  T bmin[3];
  T bmax[3];
  int flag;  // 1 = leaf, 0 = branch
  int axis;
  unsigned int data[2];  // leaf {count, first index}; branch {left, right}
};

template <typename T = float>
struct BVHBuildOptions {
  T cost_t_aabb;
  unsigned int min_leaf_primitives;
  unsigned int max_tree_depth;
  unsigned int bin_size;
  unsigned int shallow_depth;
  unsigned int min_primitives_for_parallel_build;
  bool cache_bbox;
  unsigned char pad[3];
  BVHBuildOptions()
      : cost_t_aabb(T(0.2)),
        min_leaf_primitives(4),
        max_tree_depth(256),
        bin_size(64),
        shallow_depth(4),
        min_primitives_for_parallel_build(1024 * 8),
        cache_bbox(false) {
    pad[0] = pad[1] = pad[2] = 0;
  }
};

class BVHBuildStatistics {
 public:
  unsigned int max_tree_depth;
  unsigned int num_leaf_nodes;
  unsigned int num_branch_nodes;
  float build_secs;  // device time of the build kernels (the reference never fills this field)
  BVHBuildStatistics() : max_tree_depth(0), num_leaf_nodes(0), num_branch_nodes(0), build_secs(0.0f) {}
};

class BVHTraceOptions {
 public:
  unsigned int prim_ids_range[2];
  unsigned int skip_prim_id;
  bool cull_back_face;
  unsigned char pad[3];
  BVHTraceOptions() {
    prim_ids_range[0] = 0;
    prim_ids_range[1] = 0x7FFFFFFF;
    skip_prim_id = static_cast<unsigned int>(-1);
    cull_back_face = false;
    pad[0] = pad[1] = pad[2] = 0;
  }
};

static_assert(sizeof(Ray<float>) == 36, "nanort::Ray<float> layout");
static_assert(sizeof(BVHNode<float>) == 40, "nanort::BVHNode<float> layout");
static_assert(sizeof(BVHBuildOptions<float>) == 28, "nanort::BVHBuildOptions<float> layout");
static_assert(sizeof(BVHBuildStatistics) == 16, "nanort::BVHBuildStatistics layout");
static_assert(sizeof(BVHTraceOptions) == 16, "nanort::BVHTraceOptions layout");

// ---- node-traversal records of the two-level API -- :629-694
template <typename T>
class BBox {
 public:
  real3<T> bmin;
  real3<T> bmax;
  BBox() {
    bmin[0] = bmin[1] = bmin[2] = std::numeric_limits<T>::max();
    bmax[0] = bmax[1] = bmax[2] = -std::numeric_limits<T>::max();
  }
};

template <typename T>
class NodeHit {
 public:
  NodeHit()
      : t_min(std::numeric_limits<T>::max()), t_max(-std::numeric_limits<T>::max()),
        node_id(static_cast<unsigned int>(-1)) {}
  T t_min;
  T t_max;
  unsigned int node_id;
};

template <typename T>
class NodeHitComparator {
 public:
  inline bool operator()(const NodeHit<T> &a, const NodeHit<T> &b) { return a.t_min < b.t_min; }
};

template <class H>
class IntersectComparator {  // :552-557
 public:
  bool operator()(const H &a, const H &b) const { return a.t < b.t; }
};

// nanort.h:286-319: the reference's fixed-capacity vector (a std::vector over a stack arena).  Same surface
// (operator->, operator[], container()); storage is a std::vector reserved to the capacity.
template <typename T, size_t stack_capacity>
class StackVector {
 public:
  typedef std::vector<T> ContainerType;
  StackVector() { container_.reserve(stack_capacity); }
  ContainerType &container() { return container_; }
  const ContainerType &container() const { return container_; }
  ContainerType *operator->() { return &container_; }
  const ContainerType *operator->() const { return &container_; }
  T &operator[](size_t i) { return container_[i]; }
  const T &operator[](size_t i) const { return container_[i]; }

 private:
  ContainerType container_;
};

// ---- built-in triangle classes ----------------------------------------------------------------
namespace detail {
// a mutex that does not make its owner non-copyable (copies of a BVHAccel share the device tree, not the lock)
struct CopyableMutex {
  std::mutex m;
  CopyableMutex() {}
  CopyableMutex(const CopyableMutex &) {}
  CopyableMutex &operator=(const CopyableMutex &) { return *this; }
};
template <typename T>
inline const T *vertex_at(const T *base, size_t i, size_t stride_bytes) {
  return reinterpret_cast<const T *>(reinterpret_cast<const unsigned char *>(base) + i * stride_bytes);
}
}  // namespace detail

template <typename T = float>
class TriangleSAHPred {
 public:
  TriangleSAHPred(const T *vertices, const unsigned int *faces, size_t vertex_stride_bytes)
      : axis_(0), pos_(T(0)), vertices_(vertices), faces_(faces), vertex_stride_bytes_(vertex_stride_bytes) {}
  void Set(int axis, T pos) const {
    axis_ = axis;
    pos_ = pos;
  }
  // left iff the centroid (times three) lies below the plane
  bool operator()(unsigned int i) const {
    const T *a = detail::vertex_at(vertices_, faces_[3 * i + 0], vertex_stride_bytes_);
    const T *b = detail::vertex_at(vertices_, faces_[3 * i + 1], vertex_stride_bytes_);
    const T *c = detail::vertex_at(vertices_, faces_[3 * i + 2], vertex_stride_bytes_);
    return (a[axis_] + b[axis_] + c[axis_]) < pos_ * T(3);
  }
  const T *GetVertices() const { return vertices_; }
  const unsigned int *GetFaces() const { return faces_; }
  size_t GetVertexStrideBytes() const { return vertex_stride_bytes_; }

 private:
  mutable int axis_;
  mutable T pos_;
  const T *vertices_;
  const unsigned int *faces_;
  size_t vertex_stride_bytes_;
};

template <typename T = float>
class TriangleMesh {
 public:
  TriangleMesh(const T *vertices, const unsigned int *faces, const size_t vertex_stride_bytes)
      : vertices_(vertices), faces_(faces), vertex_stride_bytes_(vertex_stride_bytes) {}

  void BoundingBox(real3<T> *bmin, real3<T> *bmax, unsigned int prim_index) const {
    for (int c = 0; c < 3; c++) {
      const T *p = detail::vertex_at(vertices_, faces_[3 * prim_index + c], vertex_stride_bytes_);
      for (int k = 0; k < 3; k++) {
        if (c == 0 || p[k] < (*bmin)[k]) (*bmin)[k] = p[k];
        if (c == 0 || p[k] > (*bmax)[k]) (*bmax)[k] = p[k];
      }
    }
  }
  void BoundingBoxAndCenter(real3<T> *bmin, real3<T> *bmax, real3<T> *center, unsigned int prim_index) const {
    BoundingBox(bmin, bmax, prim_index);
    const T *a = detail::vertex_at(vertices_, faces_[3 * prim_index + 0], vertex_stride_bytes_);
    const T *b = detail::vertex_at(vertices_, faces_[3 * prim_index + 1], vertex_stride_bytes_);
    const T *c = detail::vertex_at(vertices_, faces_[3 * prim_index + 2], vertex_stride_bytes_);
    for (int k = 0; k < 3; k++) (*center)[k] = ((a[k] + b[k]) + c[k]) * (T(1) / T(3));
  }

  const T *vertices_;
  const unsigned int *faces_;
  const size_t vertex_stride_bytes_;

  const T *GetVertices() const { return vertices_; }
  const unsigned int *GetFaces() const { return faces_; }
  size_t GetVertexStrideBytes() const { return vertex_stride_bytes_; }
};

template <typename T = float>
class TriangleIntersection {
 public:
  T u;
  T v;
  T t;
  unsigned int prim_id;
};
static_assert(sizeof(TriangleIntersection<float>) == 16, "nanort::TriangleIntersection<float> layout");

// The watertight ray/triangle test itself runs on the device (csrc/traverse.cu:tri_test); this class only
// carries the geometry pointers, exactly as the reference's constructors take them.
template <typename T = float, class H = TriangleIntersection<T> >
class TriangleIntersector {
 public:
  template <class M>
  explicit TriangleIntersector(const M &m)
      : vertices_(m.GetVertices()), faces_(m.GetFaces()), vertex_stride_bytes_(m.GetVertexStrideBytes()) {}
  template <class M>
  explicit TriangleIntersector(const M *m)
      : vertices_(m->GetVertices()), faces_(m->GetFaces()), vertex_stride_bytes_(m->GetVertexStrideBytes()) {}
  TriangleIntersector(const T *vertices, const unsigned int *faces, const size_t vertex_stride_bytes)
      : vertices_(vertices), faces_(faces), vertex_stride_bytes_(vertex_stride_bytes) {}

  const T *GetVertices() const { return vertices_; }
  const unsigned int *GetFaces() const { return faces_; }
  size_t GetVertexStrideBytes() const { return vertex_stride_bytes_; }

  /// Closest distance found by the last Traverse made with this object (nanort.h:1159; callers such as
  /// examples/par_msquare/main.cc:501 read it instead of the hit record).
  T GetT() const { return t_; }
  /// BVHAccel::Traverse stores the device's result here (what Update + PostTraversal do in the reference, :1153-1213)
  void SetResult(T t, T u, T v, unsigned int prim_id) const {
    t_ = t;
    u_ = u;
    v_ = v;
    prim_id_ = prim_id;
  }

 private:
  const T *vertices_;
  const unsigned int *faces_;
  const size_t vertex_stride_bytes_;
  mutable T t_ = T(0), u_ = T(0), v_ = T(0);
  mutable unsigned int prim_id_ = static_cast<unsigned int>(-1);
};

namespace detail {
// writes the result into intersectors that can take it (the built-in one); user types without SetResult are left alone
template <class I, class T>
inline auto set_result(const I &isec, T t, T u, T v, unsigned int prim, int) -> decltype(isec.SetResult(t, u, v, prim), void()) {
  isec.SetResult(t, u, v, prim);
}
template <class I, class T>
inline void set_result(const I &, T, T, T, unsigned int, long) {}
}  // namespace detail

// ---- user-defined primitives ------------------------------------------------------------------------
// The reference lets a caller hand ANY Prim / Pred / Intersector classes to BVHAccel (nanort.h:698-860).  Host functors
// cannot run inside a CUDA kernel, so on the device the concept maps to primitive KINDS (include/nanort_b200.h,
// nrt_build_prims).  Build() picks the kind from the Prim class:
//   * TriangleMesh<T>                                  -> triangles (the hot path)
//   * a class with members `vertices_` and `radiuss_`  -> spheres: the SphereGeometry of the reference's
//     examples/particle_primitive/main.cc:105-145 -- that example compiles unmodified; other sphere-like classes opt in
//     by specialising nanort::DeviceSpheres<Prim> below
//   * anything else                                    -> its bounding boxes: p.BoundingBox() is evaluated on the HOST for
//     every primitive and the tree is built over those boxes; such an accel answers ListNodeIntersections (the
//     node-level query of the two-level API, nanort.h:2607-2692) -- Traverse() needs a kind the device knows.
template <class Prim, class Enable = void>
struct DeviceSpheres {
  static const bool value = false;
};
template <class Prim>
struct DeviceSpheres<Prim, decltype((void)std::declval<const Prim &>().vertices_, (void)std::declval<const Prim &>().radiuss_, void())> {
  static const bool value = true;
  static const float *centers(const Prim &p) { return p.vertices_; }
  static const float *radii(const Prim &p) { return p.radiuss_; }
  static size_t center_stride_bytes(const Prim &) { return 3 * sizeof(float); }
};

namespace detail {
struct triangle_tag {};
struct sphere_tag {};
struct boxes_tag {};
template <class Prim>
struct prim_tag {
  typedef typename std::conditional<std::is_same<Prim, TriangleMesh<float> >::value, triangle_tag,
                                    typename std::conditional<DeviceSpheres<Prim>::value, sphere_tag, boxes_tag>::type>::type type;
};
// examples/cylinder_primitive/main.cc names its members exactly like the sphere example (`vertices_`, `radiuss_`) but
// stores TWO end points and radii per primitive and tests caps: its intersector is recognised by `test_cap_` and refused
// (cylinders are not a device kind) instead of being walked as spheres
template <class I, class Enable = void>
struct is_cylinder_like : std::false_type {};
template <class I>
struct is_cylinder_like<I, decltype((void)std::declval<const I &>().test_cap_, void())> : std::true_type {};
template <class I>
struct is_triangle_intersector : std::false_type {};
template <class H>
struct is_triangle_intersector<TriangleIntersector<float, H> > : std::true_type {};
}  // namespace detail

// ---- BVHAccel ------------------------------------------------------------------------------------
template <typename T>
class BVHAccel {
  static_assert(std::is_same<T, float>::value || std::is_same<T, double>::value,
                "nanort_b200: BVHAccel runs on the GPU for float and double");
};

template <>
class BVHAccel<float> {
 public:
  BVHAccel() : n_prims_(0) {}
  ~BVHAccel() {}

  /// Builds the BVH on the device.  Returns false for num_primitives == 0 (like the reference) and when no
  /// CUDA device is usable.
  template <class Prim, class Pred>
  bool Build(const unsigned int num_primitives, const Prim &p, const Pred &pred,
             const BVHBuildOptions<float> &options = BVHBuildOptions<float>()) {
    (void)pred;  // the SAH partition runs on the device over the primitives' boxes / centroids
    handle_.reset();
    nodes_.clear();
    indices_.clear();
    mirrors_ = false;
    stats_ = BVHBuildStatistics();
    options_ = options;
    n_prims_ = 0;
    kind_ = 0;
    if (num_primitives == 0) return false;
    nrt_accel *h = NULL;
    const int rc = BuildKind(num_primitives, p, options, &h, typename detail::prim_tag<Prim>::type());
    if (rc != NRT_OK) {
      fprintf(stderr, "nanort_b200: Build failed: %s\n", nrt_last_error());
      return false;
    }
    handle_ = std::shared_ptr<nrt_accel>(h, nrt_free);
    n_prims_ = num_primitives;
    nrt_stats(h, &stats_);
    return true;
  }

  BVHBuildStatistics GetStatistics() const { return stats_; }

  /// One ray, synchronously -- source compatible with the reference, but every call is a full host<->device
  /// round trip.  Renderers should batch with TraverseBatch.
  template <class I, class H>
  bool Traverse(const Ray<float> &ray, const I &intersector, H *isect,
                const BVHTraceOptions &options = BVHTraceOptions()) const {
    if (!ReadyFor(intersector, detail::is_triangle_intersector<I>())) return false;
    TriangleIntersection<float> rec;
    unsigned char hit = 0;
    if (nrt_traverse(handle_.get(), &ray, 1, &rec, &hit, &options, NANORT_B200_TRAVERSE_FLAGS) != NRT_OK) {
      fprintf(stderr, "nanort_b200: Traverse failed: %s\n", nrt_last_error());
      return false;
    }
    detail::set_result(intersector, rec.t, rec.u, rec.v, rec.prim_id, 0);
    if (hit && isect) {  // *isect stays untouched on a miss, as in the reference
      isect->t = rec.t;
      isect->u = rec.u;
      isect->v = rec.v;
      isect->prim_id = rec.prim_id;
    }
    return hit != 0;
  }

  /// Extension: n rays at once.  hits[i] is valid where hit_mask[i] != 0 (miss records are
  /// {0, 0, ray.max_t, 0xFFFFFFFF}).  Returns the number of hits, or (size_t)-1 on error.
  /// `flags`: NRT_TRAVERSE_FAST or NRT_TRAVERSE_CONFORMANCE (reference visiting order).
  template <class I>
  size_t TraverseBatch(const Ray<float> *rays, size_t n, const I &intersector, TriangleIntersection<float> *hits,
                       unsigned char *hit_mask, const BVHTraceOptions &options = BVHTraceOptions(),
                       unsigned int flags = NANORT_B200_TRAVERSE_FLAGS) const {
    if (!Ready(intersector)) return static_cast<size_t>(-1);
    std::vector<unsigned char> tmp;
    if (!hit_mask) {
      tmp.resize(n);
      hit_mask = tmp.data();
    }
    if (nrt_traverse(handle_.get(), rays, n, hits, hit_mask, &options, flags) != NRT_OK) {
      fprintf(stderr, "nanort_b200: TraverseBatch failed: %s\n", nrt_last_error());
      return static_cast<size_t>(-1);
    }
    size_t c = 0;
    for (size_t i = 0; i < n; i++) c += hit_mask[i] ? 1 : 0;
    return c;
  }

  /// nanort.h:2607-2692: the (at most max_intersections) nearest primitive BOXES the ray pierces, nearest first.  The
  /// accel must have been built from a box-like Prim (see "user-defined primitives" above); the box test is the one of
  /// the reference's NodeBBoxIntersector (examples/nanosg/nanosg.h:562-640) -- `intersector` only names the kind.
  template <class I>
  bool ListNodeIntersections(const Ray<float> &ray, int max_intersections, const I &intersector,
                             StackVector<NodeHit<float>, 128> *hits) const {
    (void)intersector;
    (*hits)->clear();
    if (!handle_ || kind_ != (int)NRT_PRIM_BOXES) {
      fprintf(stderr, "nanort_b200: ListNodeIntersections needs an accel built over boxes\n");
      return false;
    }
    if (max_intersections > 64) max_intersections = 64;
    struct Rec {
      float t_min, t_max;
      unsigned int node_id;
    } recs[64];
    uint32_t count = 0;
    if (nrt_list_node_intersections(handle_.get(), &ray, 1, max_intersections, recs, &count, NANORT_B200_INVERSE_FLAG) != NRT_OK) {
      fprintf(stderr, "nanort_b200: ListNodeIntersections failed: %s\n", nrt_last_error());
      return false;
    }
    (*hits)->resize(count);
    for (uint32_t k = 0; k < count; k++) {
      (*hits)[k].t_min = recs[k].t_min;
      (*hits)[k].t_max = recs[k].t_max;
      (*hits)[k].node_id = recs[k].node_id;
    }
    return count > 0;
  }

  const std::vector<BVHNode<float> > &GetNodes() const {
    Mirror();
    return nodes_;
  }
  const std::vector<unsigned int> &GetIndices() const {
    Mirror();
    return indices_;
  }

  void BoundingBox(float bmin[3], float bmax[3]) const {
    if (!IsValid()) {
      bmin[0] = bmin[1] = bmin[2] = std::numeric_limits<float>::max();
      bmax[0] = bmax[1] = bmax[2] = -std::numeric_limits<float>::max();
      return;
    }
    if (handle_) {
      nrt_bounding_box(handle_.get(), bmin, bmax);
    } else {
      for (int k = 0; k < 3; k++) {
        bmin[k] = nodes_[0].bmin[k];
        bmax[k] = nodes_[0].bmax[k];
      }
    }
  }

  bool IsValid() const { return handle_ || !nodes_.empty(); }

  /// Raw dump in the reference's format (size_t count; nodes; size_t count; indices).
  bool Dump(FILE *fp) const {
    Mirror();
    size_t n = nodes_.size(), m = indices_.size();
    if (fwrite(&n, sizeof(size_t), 1, fp) != 1) return false;
    if (n && fwrite(nodes_.data(), sizeof(BVHNode<float>), n, fp) != n) return false;
    if (fwrite(&m, sizeof(size_t), 1, fp) != 1) return false;
    if (m && fwrite(indices_.data(), sizeof(unsigned int), m, fp) != m) return false;
    return true;
  }
  bool Dump(const char *filename) const {
    FILE *fp = fopen(filename, "wb");
    if (!fp) return false;
    bool ok = Dump(fp);
    fclose(fp);
    return ok;
  }
  /// Loads a dumped tree (e.g. one built by CPU nanort).  The device copy is created by the first Traverse,
  /// which brings the geometry pointers, as in the reference where Load restores nodes/indices only.
  bool Load(FILE *fp) {
    handle_.reset();
    nodes_.clear();
    indices_.clear();
    mirrors_ = false;
    size_t n = 0, m = 0;
    if (fread(&n, sizeof(size_t), 1, fp) != 1 || n == 0) return false;
    nodes_.resize(n);
    if (fread(nodes_.data(), sizeof(BVHNode<float>), n, fp) != n) return false;
    if (fread(&m, sizeof(size_t), 1, fp) != 1) return false;
    indices_.resize(m);
    if (m && fread(indices_.data(), sizeof(unsigned int), m, fp) != m) return false;
    mirrors_ = true;
    n_prims_ = static_cast<unsigned int>(m);
    return true;
  }
  bool Load(const char *filename) {
    FILE *fp = fopen(filename, "rb");
    if (!fp) return false;
    bool ok = Load(fp);
    fclose(fp);
    return ok;
  }

  /// Escape hatch for code that wants the device-pointer entry points of nanort_b200.h.
  const nrt_accel *NativeHandle() const { return handle_.get(); }

 private:
  template <class Prim>
  int BuildKind(unsigned int n, const Prim &p, const BVHBuildOptions<float> &options, nrt_accel **h, detail::triangle_tag) {
    kind_ = 0;
    return nrt_build_ex(p.GetVertices(), p.GetVertexStrideBytes(), 0, p.GetFaces(), n, &options, NANORT_B200_BUILD_FLAGS, h);
  }
  template <class Prim>
  int BuildKind(unsigned int n, const Prim &p, const BVHBuildOptions<float> &options, nrt_accel **h, detail::sphere_tag) {
    kind_ = (int)NRT_PRIM_SPHERES;
    return nrt_build_prims(NRT_PRIM_SPHERES, DeviceSpheres<Prim>::centers(p), DeviceSpheres<Prim>::center_stride_bytes(p),
                           DeviceSpheres<Prim>::radii(p), n, &options, h);
  }
  template <class Prim>
  int BuildKind(unsigned int n, const Prim &p, const BVHBuildOptions<float> &options, nrt_accel **h, detail::boxes_tag) {
    kind_ = (int)NRT_PRIM_BOXES;
    std::vector<float> boxes(6 * static_cast<size_t>(n));  // the user's Prim::BoundingBox, evaluated on the host
    for (unsigned int i = 0; i < n; i++) {
      real3<float> bmin, bmax;
      p.BoundingBox(&bmin, &bmax, i);
      for (int k = 0; k < 3; k++) {
        boxes[6 * static_cast<size_t>(i) + k] = bmin[k];
        boxes[6 * static_cast<size_t>(i) + 3 + k] = bmax[k];
      }
    }
    return nrt_build_prims(NRT_PRIM_BOXES, boxes.data(), 24, NULL, n, &options, h);
  }
  template <class I>
  bool ReadyFor(const I &isec, std::true_type) const {
    if (kind_ != 0) {
      fprintf(stderr, "nanort_b200: a TriangleIntersector was handed to an accel built over another primitive kind\n");
      return false;
    }
    return Ready(isec);
  }
  template <class I>
  bool ReadyFor(const I &, std::false_type) const {
    if (detail::is_cylinder_like<I>::value) {
      fprintf(stderr, "nanort_b200: this intersector (a `test_cap_` member: the cylinder model of "
                      "examples/cylinder_primitive) is not a primitive kind the device knows; spheres, boxes and "
                      "triangles are\n");
      return false;
    }
    if (handle_ && kind_ == (int)NRT_PRIM_SPHERES) return true;
    fprintf(stderr, "nanort_b200: Traverse with a user-defined intersector needs an accel of a kind the device knows "
                    "(spheres: see DeviceSpheres in nanort.h); this accel has kind %d\n", kind_);
    return false;
  }
  // The reference's Traverse may run on many threads at once (examples/path_tracer/main.cc:787-799): the lazy adopt of
  // a Load()ed tree and the lazy host mirror are serialised.
  template <class I>
  bool Ready(const I &isec) const {
    std::lock_guard<std::mutex> lock(mu_.m);
    if (handle_) return true;
    if (nodes_.empty()) return false;
    // a Load()ed tree: adopt it now that the intersector supplies the geometry
    unsigned int max_prim = 0;
    for (size_t i = 0; i < indices_.size(); i++) max_prim = indices_[i] > max_prim ? indices_[i] : max_prim;
    nrt_accel *h = NULL;
    int rc = nrt_adopt(nodes_.data(), nodes_.size(), indices_.data(), indices_.size(), isec.GetVertices(),
                       isec.GetVertexStrideBytes(), 0, isec.GetFaces(), static_cast<uint32_t>(indices_.size()), &h);
    (void)max_prim;
    if (rc != NRT_OK) {
      fprintf(stderr, "nanort_b200: adopting the loaded tree failed: %s\n", nrt_last_error());
      return false;
    }
    handle_ = std::shared_ptr<nrt_accel>(h, nrt_free);
    nrt_stats(h, &stats_);
    return true;
  }
  void Mirror() const {
    std::lock_guard<std::mutex> lock(mu_.m);
    if (mirrors_ || !handle_) return;
    const void *pn = NULL;
    const uint32_t *pi = NULL;
    size_t nn = 0, ni = 0;
    if (nrt_nodes(handle_.get(), &pn, &nn, &pi, &ni) != NRT_OK) return;
    nodes_.resize(nn);
    if (nn) memcpy(nodes_.data(), pn, nn * sizeof(BVHNode<float>));
    indices_.assign(pi, pi + ni);
    mirrors_ = true;
  }

  mutable std::shared_ptr<nrt_accel> handle_;  // copies of a BVHAccel share the (immutable) device tree
  mutable std::vector<BVHNode<float> > nodes_;
  mutable std::vector<unsigned int> indices_;
  mutable bool mirrors_ = false;
  mutable detail::CopyableMutex mu_;
  BVHBuildOptions<float> options_;
  mutable BVHBuildStatistics stats_;
  unsigned int n_prims_;
  int kind_ = 0;  // 0 triangles, NRT_PRIM_SPHERES, NRT_PRIM_BOXES
};

// ---- BVHAccel<double> ----------------------------------------------------------------------------
// The fp64 instantiation (nrt_build_f64 / nrt_adopt_f64 / nrt_traverse_f64): same members as above.  The tree's
// topology comes from the production builder, every node box is exact in double, Traverse computes in double --
// t / u / v carry the reference's bits for the reported primitive.  Traverse (one ray) walks in the reference's visiting
// order; TraverseBatch uses the persistent-warp fast kernel (csrc/f64_fast.cuh) unless NANORT_B200_CONFORMANCE is defined.
static_assert(sizeof(Ray<double>) == 72, "nanort::Ray<double> layout");
static_assert(sizeof(BVHNode<double>) == 64, "nanort::BVHNode<double> layout");
static_assert(sizeof(BVHBuildOptions<double>) == 32, "nanort::BVHBuildOptions<double> layout");
static_assert(sizeof(TriangleIntersection<double>) == 32, "nanort::TriangleIntersection<double> layout");

template <>
class BVHAccel<double> {
 public:
  BVHAccel() : mirrors_(false) {}

  template <class Prim, class Pred>
  bool Build(const unsigned int num_primitives, const Prim &p, const Pred &pred,
             const BVHBuildOptions<double> &options = BVHBuildOptions<double>()) {
    static_assert(std::is_same<Prim, TriangleMesh<double> >::value && std::is_same<Pred, TriangleSAHPred<double> >::value,
                  "nanort_b200: Build runs on the GPU for TriangleMesh<double> + TriangleSAHPred<double> only");
    (void)pred;
    handle_.reset();
    nodes_.clear();
    indices_.clear();
    mirrors_ = false;
    stats_ = BVHBuildStatistics();
    if (num_primitives == 0) return false;
    nrt_accel_f64 *h = NULL;
    // -DNANORT_B200_CONFORMANCE: the reference's own BVHNode<double> array and indices_ (csrc/build_ref64.cu)
    if (nrt_build_f64_ex(p.GetVertices(), p.GetVertexStrideBytes(), 0, p.GetFaces(), num_primitives, &options,
                         NANORT_B200_BUILD_FLAGS, &h) != NRT_OK) {
      fprintf(stderr, "nanort_b200: Build<double> failed: %s\n", nrt_last_error());
      return false;
    }
    handle_ = std::shared_ptr<nrt_accel_f64>(h, nrt_free_f64);
    nrt_stats_f64(h, &stats_);
    return true;
  }

  BVHBuildStatistics GetStatistics() const { return stats_; }
  bool IsValid() const { return handle_ != NULL || !nodes_.empty(); }

  /// Raw dump / load in the reference's format (nanort.h:2164-2276); a loaded tree reaches the device at the first
  /// Traverse, which brings the geometry pointers.
  bool Dump(FILE *fp) const {
    Mirror();
    size_t n = nodes_.size(), m = indices_.size();
    if (fwrite(&n, sizeof(size_t), 1, fp) != 1) return false;
    if (n && fwrite(nodes_.data(), sizeof(BVHNode<double>), n, fp) != n) return false;
    if (fwrite(&m, sizeof(size_t), 1, fp) != 1) return false;
    if (m && fwrite(indices_.data(), sizeof(unsigned int), m, fp) != m) return false;
    return true;
  }
  bool Load(FILE *fp) {
    handle_.reset();
    nodes_.clear();
    indices_.clear();
    mirrors_ = false;
    size_t n = 0, m = 0;
    if (fread(&n, sizeof(size_t), 1, fp) != 1 || n == 0) return false;
    nodes_.resize(n);
    if (fread(nodes_.data(), sizeof(BVHNode<double>), n, fp) != n) return false;
    if (fread(&m, sizeof(size_t), 1, fp) != 1) return false;
    indices_.resize(m);
    if (m && fread(indices_.data(), sizeof(unsigned int), m, fp) != m) return false;
    mirrors_ = true;
    return true;
  }

  template <class I, class H>
  bool Traverse(const Ray<double> &ray, const I &intersector, H *isect,
                const BVHTraceOptions &options = BVHTraceOptions()) const {
    if (!Ready(intersector)) return false;
    TriangleIntersection<double> rec;
    unsigned char hit = 0;
    // one ray: the reference-order kernel (a persistent-warp launch has nothing to amortise over)
    if (nrt_traverse_f64(handle_.get(), &ray, 1, &rec, &hit, &options, NRT_TRAVERSE_CONFORMANCE | NANORT_B200_INVERSE_FLAG) != NRT_OK) {
      fprintf(stderr, "nanort_b200: Traverse<double> failed: %s\n", nrt_last_error());
      return false;
    }
    detail::set_result(intersector, rec.t, rec.u, rec.v, rec.prim_id, 0);
    if (hit && isect) {
      isect->t = rec.t;
      isect->u = rec.u;
      isect->v = rec.v;
      isect->prim_id = rec.prim_id;
    }
    return hit != 0;
  }

  /// Extension: n rays at once (see BVHAccel<float>::TraverseBatch).
  template <class I>
  size_t TraverseBatch(const Ray<double> *rays, size_t n, const I &intersector, TriangleIntersection<double> *hits,
                       unsigned char *hit_mask, const BVHTraceOptions &options = BVHTraceOptions()) const {
    if (!Ready(intersector)) return static_cast<size_t>(-1);
    std::vector<unsigned char> tmp;
    if (!hit_mask) {
      tmp.resize(n);
      hit_mask = tmp.data();
    }
    if (nrt_traverse_f64(handle_.get(), rays, n, hits, hit_mask, &options, NANORT_B200_TRAVERSE_FLAGS) != NRT_OK) {
      fprintf(stderr, "nanort_b200: TraverseBatch<double> failed: %s\n", nrt_last_error());
      return static_cast<size_t>(-1);
    }
    size_t c = 0;
    for (size_t i = 0; i < n; i++) c += hit_mask[i] ? 1 : 0;
    return c;
  }

  const std::vector<BVHNode<double> > &GetNodes() const {
    Mirror();
    return nodes_;
  }
  const std::vector<unsigned int> &GetIndices() const {
    Mirror();
    return indices_;
  }

  void BoundingBox(double bmin[3], double bmax[3]) const {
    if (handle_) {
      nrt_bounding_box_f64(handle_.get(), bmin, bmax);
    } else if (!nodes_.empty()) {
      for (int k = 0; k < 3; k++) {
        bmin[k] = nodes_[0].bmin[k];
        bmax[k] = nodes_[0].bmax[k];
      }
    } else {
      bmin[0] = bmin[1] = bmin[2] = std::numeric_limits<double>::max();
      bmax[0] = bmax[1] = bmax[2] = -std::numeric_limits<double>::max();
    }
  }

 private:
  template <class I>
  bool Ready(const I &isec) const {
    std::lock_guard<std::mutex> lock(mu_.m);
    if (handle_) return true;
    if (nodes_.empty()) return false;
    nrt_accel_f64 *h = NULL;  // a Load()ed tree: adopt it now that the intersector supplies the geometry
    if (nrt_adopt_f64(nodes_.data(), nodes_.size(), indices_.data(), indices_.size(), isec.GetVertices(),
                      isec.GetVertexStrideBytes(), 0, isec.GetFaces(), static_cast<uint32_t>(indices_.size()), &h) !=
        NRT_OK) {
      fprintf(stderr, "nanort_b200: adopting the loaded fp64 tree failed: %s\n", nrt_last_error());
      return false;
    }
    handle_ = std::shared_ptr<nrt_accel_f64>(h, nrt_free_f64);
    nrt_stats_f64(h, &stats_);
    return true;
  }
  void Mirror() const {
    std::lock_guard<std::mutex> lock(mu_.m);
    if (mirrors_ || !handle_) return;
    const void *pn = NULL;
    const uint32_t *pi = NULL;
    size_t nn = 0, ni = 0;
    if (nrt_nodes_f64(handle_.get(), &pn, &nn, &pi, &ni) != NRT_OK) return;
    nodes_.resize(nn);
    if (nn) memcpy(nodes_.data(), pn, nn * sizeof(BVHNode<double>));
    indices_.assign(pi, pi + ni);
    mirrors_ = true;
  }
  mutable std::shared_ptr<nrt_accel_f64> handle_;
  mutable std::vector<BVHNode<double> > nodes_;
  mutable std::vector<unsigned int> indices_;
  mutable bool mirrors_;
  mutable detail::CopyableMutex mu_;
  mutable BVHBuildStatistics stats_;
};

}  // namespace nanort

#endif  // NANORT_H_
