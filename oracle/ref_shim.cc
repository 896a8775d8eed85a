// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// Thin extern "C" shim around the UNMODIFIED reference header (nanort.h found
// with -I$(NANORT_REF), normally /root/reference).  It is compiled by
// oracle/Makefile into oracle/_ref/libnanort_ref.so (C++11/threads mode, the
// mode examples/path_tracer/Makefile:2 uses) and oracle/_ref/libnanort_ref03.so
// (C++03 serial mode).  No reference source is copied into this repository:
// the header is read in place at compile time and only the resulting .so lives
// under the git-ignored oracle/_ref/.
//
// Uses of the reference API (file:line in /root/reference):
//   BVHAccel<float>::Build            nanort.h:1892-2149
//   BVHAccel<float>::Traverse         nanort.h:2487-2556
//   BVHAccel<float>::Dump/Load(FILE*) nanort.h:2164-2276 (node-array injection)
//   TriangleMesh / TriangleSAHPred / TriangleIntersector  nanort.h:863-1229
#define NANORT_ENABLE_SERIALIZATION
#include "nanort.h"

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

namespace {

struct RefAccel {
  nanort::BVHAccel<float> accel;
  const float *verts;
  const unsigned int *faces;
  size_t stride;
};

}  // namespace

extern "C" {

// 1 when built with NANORT_USE_CPP11_FEATURE (copysign vsafe_inverse, threaded
// shallow-tree build above 8192 prims), 0 for the C++03 serial mode.
int ref_mode_cpp11(void) {
#ifdef NANORT_USE_CPP11_FEATURE
  return 1;
#else
  return 0;
#endif
}

void ref_sizes(uint32_t out[5]) {
  out[0] = sizeof(nanort::BVHNode<float>);
  out[1] = sizeof(nanort::Ray<float>);
  out[2] = sizeof(nanort::TriangleIntersection<float>);
  out[3] = sizeof(nanort::BVHBuildOptions<float>);
  out[4] = sizeof(nanort::BVHTraceOptions);
}

void ref_default_build_options(void *out28) {
  nanort::BVHBuildOptions<float> o;
  std::memcpy(out28, &o, sizeof(o));
}

void ref_default_trace_options(void *out16) {
  nanort::BVHTraceOptions o;
  std::memcpy(out16, &o, sizeof(o));
}

// Geometry pointers are borrowed (as in the reference API) and must outlive
// the handle.
void *ref_build(const float *verts, size_t stride, const uint32_t *faces,
                uint32_t n_prims, const void *build_opts28) {
  RefAccel *r = new RefAccel();
  r->verts = verts;
  r->faces = faces;
  r->stride = stride;
  nanort::BVHBuildOptions<float> o;
  if (build_opts28) std::memcpy(&o, build_opts28, sizeof(o));
  nanort::TriangleMesh<float> mesh(verts, faces, stride);
  nanort::TriangleSAHPred<float> pred(verts, faces, stride);
  bool ok = r->accel.Build(n_prims, mesh, pred, o);
  if (!ok) {
    delete r;
    return NULL;
  }
  return r;
}

// Injects an arbitrary (nodes, indices) pair -- e.g. a GPU-built tree -- into
// a reference BVHAccel through its own Dump format: size_t n; nodes; size_t
// m; indices.
void *ref_adopt(const void *nodes40, size_t n_nodes, const uint32_t *indices,
                size_t n_indices, const float *verts, size_t stride,
                const uint32_t *faces) {
  size_t bytes = 2 * sizeof(size_t) + n_nodes * sizeof(nanort::BVHNode<float>) +
                 n_indices * sizeof(uint32_t);
  std::vector<unsigned char> buf(bytes);
  unsigned char *p = buf.data();
  std::memcpy(p, &n_nodes, sizeof(size_t));
  p += sizeof(size_t);
  std::memcpy(p, nodes40, n_nodes * sizeof(nanort::BVHNode<float>));
  p += n_nodes * sizeof(nanort::BVHNode<float>);
  std::memcpy(p, &n_indices, sizeof(size_t));
  p += sizeof(size_t);
  std::memcpy(p, indices, n_indices * sizeof(uint32_t));
  FILE *fp = fmemopen(buf.data(), bytes, "rb");
  if (!fp) return NULL;
  RefAccel *r = new RefAccel();
  r->verts = verts;
  r->faces = faces;
  r->stride = stride;
  bool ok = r->accel.Load(fp);
  fclose(fp);
  if (!ok) {
    delete r;
    return NULL;
  }
  return r;
}

void ref_free(void *h) { delete static_cast<RefAccel *>(h); }

void ref_stats(const void *h, uint32_t out[3]) {
  nanort::BVHBuildStatistics s = static_cast<const RefAccel *>(h)->accel.GetStatistics();
  out[0] = s.max_tree_depth;
  out[1] = s.num_leaf_nodes;
  out[2] = s.num_branch_nodes;
}

void ref_bounding_box(const void *h, float bmin[3], float bmax[3]) {
  static_cast<const RefAccel *>(h)->accel.BoundingBox(bmin, bmax);
}

size_t ref_num_nodes(const void *h) {
  return static_cast<const RefAccel *>(h)->accel.GetNodes().size();
}
size_t ref_num_indices(const void *h) {
  return static_cast<const RefAccel *>(h)->accel.GetIndices().size();
}
void ref_copy_nodes(const void *h, void *out40) {
  const std::vector<nanort::BVHNode<float> > &n =
      static_cast<const RefAccel *>(h)->accel.GetNodes();
  std::memcpy(out40, n.data(), n.size() * sizeof(nanort::BVHNode<float>));
}
void ref_copy_indices(const void *h, uint32_t *out) {
  const std::vector<unsigned int> &n = static_cast<const RefAccel *>(h)->accel.GetIndices();
  std::memcpy(out, n.data(), n.size() * sizeof(uint32_t));
}

// One Traverse per ray, exactly as examples/path_tracer/main.cc:851-854 does
// (a fresh intersector per ray).  hits[i] is written only on hit (reference
// semantics); mask[i] = 1/0.  Rays are handed out in chunks through an atomic
// counter like the row loop of examples/path_tracer/main.cc:787-799.
// Returns the number of hits.
size_t ref_traverse_batch(const void *h, const void *rays36, size_t n_rays, void *hits16,
                          uint8_t *mask, const void *trace_opts16, int n_threads) {
  const RefAccel *r = static_cast<const RefAccel *>(h);
  const nanort::Ray<float> *rays = static_cast<const nanort::Ray<float> *>(rays36);
  nanort::TriangleIntersection<float> *hits =
      static_cast<nanort::TriangleIntersection<float> *>(hits16);
  nanort::BVHTraceOptions topt;
  if (trace_opts16) std::memcpy(&topt, trace_opts16, sizeof(topt));
  if (n_threads < 1) n_threads = 1;
  std::atomic<size_t> next(0);
  std::atomic<size_t> total(0);
  const size_t chunk = 1024;
  auto work = [&]() {
    size_t local = 0;
    for (;;) {
      size_t b = next.fetch_add(chunk);
      if (b >= n_rays) break;
      size_t e = b + chunk < n_rays ? b + chunk : n_rays;
      for (size_t i = b; i < e; i++) {
        nanort::TriangleIntersector<float> isector(r->verts, r->faces, r->stride);
        nanort::TriangleIntersection<float> isect;
        bool hit = r->accel.Traverse(rays[i], isector, &isect, topt);
        if (hit) {
          hits[i] = isect;
          local++;
        }
        if (mask) mask[i] = hit ? 1 : 0;
      }
    }
    total += local;
  };
  if (n_threads == 1) {
    work();
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work);
    for (auto &t : th) t.join();
  }
  return total.load();
}

// fp64 single-ray path, used only to replay the reference's one regression
// program (test/regression/possible-accuracy-problem-30/main.cc:24-76).
// out = {t, u, v}; returns 1 on hit.
int ref_traverse_one_f64(const double *verts, const uint32_t *faces, uint32_t n_prims,
                         const double org[3], const double dir[3], double min_t, double max_t,
                         double out_tuv[3], uint32_t *out_prim) {
  nanort::BVHAccel<double> accel;
  nanort::TriangleMesh<double> mesh(verts, faces, sizeof(double) * 3);
  nanort::TriangleSAHPred<double> pred(verts, faces, sizeof(double) * 3);
  if (!accel.Build(n_prims, mesh, pred, nanort::BVHBuildOptions<double>())) return -1;
  nanort::Ray<double> ray;
  for (int k = 0; k < 3; k++) {
    ray.org[k] = org[k];
    ray.dir[k] = dir[k];
  }
  ray.min_t = min_t;
  ray.max_t = max_t;
  nanort::TriangleIntersector<double, nanort::TriangleIntersection<double> > isector(
      verts, faces, sizeof(double) * 3);
  nanort::TriangleIntersection<double> isect;
  bool hit = accel.Traverse(ray, isector, &isect);
  if (hit) {
    out_tuv[0] = isect.t;
    out_tuv[1] = isect.u;
    out_tuv[2] = isect.v;
    *out_prim = isect.prim_id;
  }
  return hit ? 1 : 0;
}

// ---- BVHAccel<double>: the same entry points for the fp64 instantiation of the reference -------------------
struct RefAccelD {
  nanort::BVHAccel<double> accel;
  const double *verts;
  const unsigned int *faces;
  size_t stride;
};

void ref64_sizes(uint32_t out[5]) {
  out[0] = sizeof(nanort::BVHNode<double>);
  out[1] = sizeof(nanort::Ray<double>);
  out[2] = sizeof(nanort::TriangleIntersection<double>);
  out[3] = sizeof(nanort::BVHBuildOptions<double>);
  out[4] = sizeof(nanort::BVHTraceOptions);
}

void *ref64_build(const double *verts, size_t stride, const uint32_t *faces, uint32_t n_prims,
                  const void *build_opts32) {
  RefAccelD *r = new RefAccelD();
  r->verts = verts;
  r->faces = faces;
  r->stride = stride;
  nanort::BVHBuildOptions<double> o;
  if (build_opts32) std::memcpy(&o, build_opts32, sizeof(o));
  nanort::TriangleMesh<double> mesh(verts, faces, stride);
  nanort::TriangleSAHPred<double> pred(verts, faces, stride);
  if (!r->accel.Build(n_prims, mesh, pred, o)) {
    delete r;
    return NULL;
  }
  return r;
}

// a (nodes, indices) pair from elsewhere -- e.g. the GPU's fp64 tree -- through the reference's own Load
void *ref64_adopt(const void *nodes64, size_t n_nodes, const uint32_t *indices, size_t n_indices,
                  const double *verts, size_t stride, const uint32_t *faces) {
  size_t bytes = 2 * sizeof(size_t) + n_nodes * sizeof(nanort::BVHNode<double>) + n_indices * sizeof(uint32_t);
  std::vector<unsigned char> buf(bytes);
  unsigned char *p = buf.data();
  std::memcpy(p, &n_nodes, sizeof(size_t));
  p += sizeof(size_t);
  std::memcpy(p, nodes64, n_nodes * sizeof(nanort::BVHNode<double>));
  p += n_nodes * sizeof(nanort::BVHNode<double>);
  std::memcpy(p, &n_indices, sizeof(size_t));
  p += sizeof(size_t);
  std::memcpy(p, indices, n_indices * sizeof(uint32_t));
  FILE *fp = fmemopen(buf.data(), bytes, "rb");
  if (!fp) return NULL;
  RefAccelD *r = new RefAccelD();
  r->verts = verts;
  r->faces = faces;
  r->stride = stride;
  bool ok = r->accel.Load(fp);
  fclose(fp);
  if (!ok) {
    delete r;
    return NULL;
  }
  return r;
}

void ref64_free(void *h) { delete static_cast<RefAccelD *>(h); }
size_t ref64_num_nodes(const void *h) { return static_cast<const RefAccelD *>(h)->accel.GetNodes().size(); }
void ref64_copy_nodes(const void *h, void *out64) {
  const std::vector<nanort::BVHNode<double> > &n = static_cast<const RefAccelD *>(h)->accel.GetNodes();
  std::memcpy(out64, n.data(), n.size() * sizeof(nanort::BVHNode<double>));
}
void ref64_copy_indices(const void *h, uint32_t *out) {
  const std::vector<unsigned int> &n = static_cast<const RefAccelD *>(h)->accel.GetIndices();
  std::memcpy(out, n.data(), n.size() * sizeof(uint32_t));
}
void ref64_bounding_box(const void *h, double bmin[3], double bmax[3]) {
  static_cast<const RefAccelD *>(h)->accel.BoundingBox(bmin, bmax);
}

size_t ref64_traverse_batch(const void *h, const void *rays72, size_t n_rays, void *hits32, uint8_t *mask,
                            const void *trace_opts16, int n_threads) {
  const RefAccelD *r = static_cast<const RefAccelD *>(h);
  const nanort::Ray<double> *rays = static_cast<const nanort::Ray<double> *>(rays72);
  nanort::TriangleIntersection<double> *hits = static_cast<nanort::TriangleIntersection<double> *>(hits32);
  nanort::BVHTraceOptions topt;
  if (trace_opts16) std::memcpy(&topt, trace_opts16, sizeof(topt));
  if (n_threads < 1) n_threads = 1;
  std::atomic<size_t> next(0), total(0);
  auto work = [&]() {
    size_t local = 0;
    for (;;) {
      size_t b = next.fetch_add(1024);
      if (b >= n_rays) break;
      size_t e = b + 1024 < n_rays ? b + 1024 : n_rays;
      for (size_t i = b; i < e; i++) {
        nanort::TriangleIntersector<double, nanort::TriangleIntersection<double> > isector(r->verts, r->faces,
                                                                                           r->stride);
        nanort::TriangleIntersection<double> isect;
        bool hit = r->accel.Traverse(rays[i], isector, &isect, topt);
        if (hit) {
          hits[i] = isect;
          local++;
        }
        if (mask) mask[i] = hit ? 1 : 0;
      }
    }
    total += local;
  };
  if (n_threads == 1) {
    work();
  } else {
    std::vector<std::thread> th;
    for (int t = 0; t < n_threads; t++) th.emplace_back(work);
    for (auto &t : th) t.join();
  }
  return total.load();
}

}  // extern "C"
