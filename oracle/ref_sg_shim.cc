// TEST INFRASTRUCTURE ONLY -- never linked into or called by the product path.
//
// extern "C" shim around the UNMODIFIED two-level scene of the reference
// (examples/nanosg/nanosg.h + nanort.h, both read in place with -I; nothing is
// copied into this repository).  oracle/Makefile compiles it into
// oracle/_ref/libnanosg_ref.so (C++11 mode) and libnanosg_ref03.so.
//
// Uses of the reference API (file:line in /root/reference):
//   nanosg::Node<float,M>::SetLocalXform / Update   examples/nanosg/nanosg.h:400-451
//   nanosg::Scene<float,M>::AddNode / Commit        examples/nanosg/nanosg.h:673-755
//   nanosg::Scene<float,M>::Traverse                examples/nanosg/nanosg.h:779-875
//   BVHAccel<float>::ListNodeIntersections          nanort.h:2607-2692
// The only liberty taken is `#define private public` around the nanosg include
// so that the tests can read the top-level BVH (Scene::toplevel_accel_).
#include <algorithm>
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <iostream>
#include <limits>
#include <queue>
#include <string>
#include <thread>
#include <vector>

#include "nanort.h"

#define private public
#include "nanosg.h"
#undef private

namespace {

// the mesh interface nanosg::Node and TriangleIntersector(const M*) need
// (examples/nanosg/mesh.h has the same members); normals are not part of the
// comparison, GetNormal reports zeros.
struct ShimMesh {
  std::vector<float> vertices;
  std::vector<unsigned int> faces;
  size_t stride;
  const float *GetVertices() const { return vertices.data(); }
  const unsigned int *GetFaces() const { return faces.data(); }
  size_t GetVertexStrideBytes() const { return stride; }
  void GetNormal(float Ng[3], float Ns[3], unsigned int, float, float) const {
    Ng[0] = Ng[1] = Ng[2] = 0.0f;
    Ns[0] = Ns[1] = Ns[2] = 0.0f;
  }
};

typedef nanosg::Node<float, ShimMesh> SgNode;
typedef nanosg::Scene<float, ShimMesh> SgScene;
typedef nanosg::Intersection<float> SgHit;

struct RefScene {
  std::deque<ShimMesh> meshes;  // stable addresses
  SgScene scene;
};

// {u, v, t, prim_id, node_id, P[3]} -- the record the C-ABI returns
struct Hit32 {
  float u, v, t;
  uint32_t prim_id, node_id;
  float P[3];
};

}  // namespace

extern "C" {

void *refsg_create(void) { return new RefScene(); }
void refsg_free(void *h) { delete static_cast<RefScene *>(h); }

// add one instance: its own copy of the triangles plus a local transform
// (float[4][4], the layout Node::SetLocalXform takes)
int refsg_add_node(void *h, const float *verts, size_t n_verts, const uint32_t *faces, size_t n_prims,
                   const float xform[16]) {
  RefScene *r = static_cast<RefScene *>(h);
  r->meshes.emplace_back();
  ShimMesh &m = r->meshes.back();
  m.vertices.assign(verts, verts + 3 * n_verts);
  m.faces.assign(faces, faces + 3 * n_prims);
  m.stride = sizeof(float) * 3;
  SgNode node(&m);
  float x[4][4];
  std::memcpy(x, xform, sizeof(x));
  node.SetLocalXform(x);
  return r->scene.AddNode(node) ? 0 : -1;
}

int refsg_commit(void *h) { return static_cast<RefScene *>(h)->scene.Commit() ? 0 : -1; }

void refsg_bounding_box(const void *h, float bmin[3], float bmax[3]) {
  static_cast<const RefScene *>(h)->scene.GetBoundingBox(bmin, bmax);
}

// per-instance derived state after Commit: out = xform[16] inv[16] inv33[16] invT33[16] lbmin[3] lbmax[3]
// xbmin[3] xbmax[3]  (76 floats)
void refsg_node_state(const void *h, size_t i, float out[76]) {
  const SgNode &n = static_cast<const RefScene *>(h)->scene.GetNodes()[i];
  std::memcpy(out, n.xform_, 64);
  std::memcpy(out + 16, n.inv_xform_, 64);
  std::memcpy(out + 32, n.inv_xform33_, 64);
  std::memcpy(out + 48, n.inv_transpose_xform33_, 64);
  float a[3], b[3];
  n.GetLocalBoundingBox(a, b);
  std::memcpy(out + 64, a, 12);
  std::memcpy(out + 67, b, 12);
  n.GetWorldBoundingBox(a, b);
  std::memcpy(out + 70, a, 12);
  std::memcpy(out + 73, b, 12);
}

size_t refsg_top_num_nodes(const void *h) {
  return static_cast<const RefScene *>(h)->scene.toplevel_accel_.GetNodes().size();
}
void refsg_top_copy(const void *h, void *nodes40, uint32_t *indices) {
  const nanort::BVHAccel<float> &a = static_cast<const RefScene *>(h)->scene.toplevel_accel_;
  std::memcpy(nodes40, a.GetNodes().data(), a.GetNodes().size() * sizeof(nanort::BVHNode<float>));
  std::memcpy(indices, a.GetIndices().data(), a.GetIndices().size() * sizeof(unsigned int));
}
// the instance's own bottom-level tree
size_t refsg_node_num_nodes(const void *h, size_t i) {
  return static_cast<const RefScene *>(h)->scene.GetNodes()[i].GetAccel().GetNodes().size();
}
void refsg_node_copy(const void *h, size_t i, void *nodes40, uint32_t *indices) {
  const nanort::BVHAccel<float> &a = static_cast<const RefScene *>(h)->scene.GetNodes()[i].GetAccel();
  std::memcpy(nodes40, a.GetNodes().data(), a.GetNodes().size() * sizeof(nanort::BVHNode<float>));
  std::memcpy(indices, a.GetIndices().data(), a.GetIndices().size() * sizeof(unsigned int));
}

// first stage alone: the sorted (t_min, t_max, node_id) list of one ray; returns the count (<= max_hits <= 128)
int refsg_list_node_intersections(const void *h, const void *ray36, int max_hits, float *tmin, float *tmax,
                                  uint32_t *ids) {
  const RefScene *r = static_cast<const RefScene *>(h);
  nanosg::NodeBBoxIntersector<float, ShimMesh> isector(&r->scene.GetNodes());
  nanort::StackVector<nanort::NodeHit<float>, 128> node_hits;
  const nanort::Ray<float> &ray = *static_cast<const nanort::Ray<float> *>(ray36);
  if (!r->scene.toplevel_accel_.ListNodeIntersections(ray, max_hits, isector, &node_hits)) return 0;
  for (size_t i = 0; i < node_hits->size(); i++) {
    tmin[i] = node_hits[i].t_min;
    tmax[i] = node_hits[i].t_max;
    ids[i] = node_hits[i].node_id;
  }
  return static_cast<int>(node_hits->size());
}

// Scene::Traverse over a batch; hits32[i] is written only where the ray hit
size_t refsg_traverse_batch(const void *h, const void *rays36, size_t n_rays, void *hits32, uint8_t *mask,
                            int n_threads) {
  const RefScene *r = static_cast<const RefScene *>(h);
  const nanort::Ray<float> *rays = static_cast<const nanort::Ray<float> *>(rays36);
  Hit32 *hits = static_cast<Hit32 *>(hits32);
  if (n_threads < 1) n_threads = 1;
  std::atomic<size_t> next(0), total(0);
  auto work = [&]() {
    size_t local = 0;
    for (;;) {
      size_t b = next.fetch_add(256);
      if (b >= n_rays) break;
      size_t e = std::min(b + 256, n_rays);
      for (size_t i = b; i < e; i++) {
        nanort::Ray<float> ray = rays[i];
        SgHit isect;
        bool hit = r->scene.Traverse<SgHit, nanort::TriangleIntersector<float, SgHit> >(ray, &isect, false);
        if (hit) {
          hits[i].u = isect.u;
          hits[i].v = isect.v;
          hits[i].t = isect.t;
          hits[i].prim_id = isect.prim_id;
          hits[i].node_id = isect.node_id;
          hits[i].P[0] = isect.P[0];
          hits[i].P[1] = isect.P[1];
          hits[i].P[2] = isect.P[2];
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
