// nanosg.h -- drop-in facade of the reference's scene-graph example header (examples/nanosg/nanosg.h) whose
// Commit and Traverse run on an NVIDIA B200 through the nrt_scene_* entry points of include/nanort_b200.h.
//
// This file is NOT the reference header and shares no code with it.  It re-declares, with the same names and
// public members, what a renderer written against nanosg uses (file:line below are
// /root/reference/examples/nanosg/nanosg.h):
//
//   nanosg::Intersection<T>            :303-316   t, prim_id, u, v, node_id, P, Ns, Ng
//   nanosg::Node<T, M>                 :322-506   SetLocalXform / GetLocalXformPtr / GetXformPtr / GetMesh /
//                                                 SetName / GetName / AddChild / GetChildren /
//                                                 GetWorldBoundingBox / GetLocalBoundingBox,
//                                                 public xform_ / inv_xform_ / inv_xform33_ / inv_transpose_xform33_
//   nanosg::Scene<T, M>                :664-905   AddNode / GetNodes / FindNode / Commit / GetBoundingBox / Traverse
//
// M is the caller's mesh class with the members the reference reads: `vertices` (std::vector<float>, xyz),
// `faces` (std::vector<unsigned int>, 3 per triangle), `stride` (bytes per vertex) and
// `GetNormal(Ng, Ns, prim_id, u, v)`.  Nodes that share a mesh object share one device BVH.  As in the reference,
// only the scene's root nodes are intersected (children are kept and updated, never traversed, :779-875).
// Extension: Scene::TraverseBatch, the form a renderer should use (one call for a whole ray array).
#ifndef NANOSG_H_
#define NANOSG_H_

#include <stdio.h>
#include <string.h>

#include <limits>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "nanort.h"

namespace nanosg {

template <typename T>
struct Intersection {
  T t;
  unsigned int prim_id;
  T u;
  T v;
  unsigned int node_id;
  nanort::real3<T> P;
  nanort::real3<T> Ns;
  nanort::real3<T> Ng;
};

namespace detail {
// dst = v . m  (row vector, translation in row 3), the reference's Matrix::MultV
template <typename T>
inline void MultV(T dst[3], const T m[4][4], const T v[3]) {
  T t[3];
  for (int k = 0; k < 3; k++) t[k] = ((m[0][k] * v[0] + m[1][k] * v[1]) + m[2][k] * v[2]) + m[3][k];
  dst[0] = t[0];
  dst[1] = t[1];
  dst[2] = t[2];
}
template <typename T>
inline void Identity(T m[4][4]) {
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) m[i][j] = (i == j) ? T(1) : T(0);
}
}  // namespace detail

template <typename T, class M>
class Scene;

template <typename T, class M>
class Node {
  static_assert(std::is_same<T, float>::value, "nanort_b200: only Node<float, M> runs on the GPU");

 public:
  typedef Node<T, M> type;

  explicit Node(const M *mesh) : mesh_(mesh) {
    for (int k = 0; k < 3; k++) {
      xbmin_[k] = lbmin_[k] = std::numeric_limits<T>::max();
      xbmax_[k] = lbmax_[k] = -std::numeric_limits<T>::max();
    }
    detail::Identity(local_xform_);
    detail::Identity(xform_);
    detail::Identity(inv_xform_);
    detail::Identity(inv_xform33_);
    inv_xform33_[3][3] = T(0);
    detail::Identity(inv_transpose_xform33_);
    inv_transpose_xform33_[3][3] = T(0);
  }

  void SetName(const std::string &name) { name_ = name; }
  const std::string &GetName() const { return name_; }
  void AddChild(const type &child) { children_.push_back(child); }
  const std::vector<type> &GetChildren() const { return children_; }
  std::vector<type> &GetChildren() { return children_; }

  void SetLocalXform(const T xform[4][4]) { memcpy(local_xform_, xform, sizeof(T) * 16); }
  const T *GetLocalXformPtr() const { return &local_xform_[0][0]; }
  const T *GetXformPtr() const { return &xform_[0][0]; }
  const M *GetMesh() const { return mesh_; }

  inline void GetWorldBoundingBox(T bmin[3], T bmax[3]) const {
    for (int k = 0; k < 3; k++) {
      bmin[k] = xbmin_[k];
      bmax[k] = xbmax_[k];
    }
  }
  inline void GetLocalBoundingBox(T bmin[3], T bmax[3]) const {
    for (int k = 0; k < 3; k++) {
      bmin[k] = lbmin_[k];
      bmax[k] = lbmax_[k];
    }
  }

  T local_xform_[4][4];
  T xform_[4][4];
  T inv_xform_[4][4];
  T inv_xform33_[4][4];
  T inv_transpose_xform33_[4][4];

 private:
  friend class Scene<T, M>;
  T lbmin_[3], lbmax_[3], xbmin_[3], xbmax_[3];
  std::string name_;
  const M *mesh_;
  std::vector<type> children_;
};

template <typename T, class M>
class Scene {
  static_assert(std::is_same<T, float>::value, "nanort_b200: only Scene<float, M> runs on the GPU");

 public:
  Scene() {
    bmin_[0] = bmin_[1] = bmin_[2] = std::numeric_limits<T>::max();
    bmax_[0] = bmax_[1] = bmax_[2] = -std::numeric_limits<T>::max();
  }

  bool AddNode(const Node<T, M> &node) {
    nodes_.push_back(node);
    return true;
  }
  const std::vector<Node<T, M> > &GetNodes() const { return nodes_; }

  bool FindNode(const std::string &name, Node<T, M> **found_node) {
    if (!found_node || name.empty()) return false;
    for (size_t i = 0; i < nodes_.size(); i++)
      if (FindRecursive(name, &nodes_[i], found_node)) return true;
    return false;
  }

  /// Builds one device BVH per distinct mesh, the per-node state and the top-level tree.
  bool Commit() {
    scene_.reset();
    if (nodes_.empty()) {
      fprintf(stderr, "You are attempting to commit an empty scene!\n");
      return false;
    }
    std::vector<nrt_instance> inst(nodes_.size());
    for (size_t i = 0; i < nodes_.size(); i++) {
      const M *mesh = nodes_[i].mesh_;
      if (!mesh || mesh->vertices.size() <= 3 || mesh->faces.size() < 3) {
        fprintf(stderr, "nanort_b200: scene node %zu has no triangle mesh\n", i);
        return false;
      }
      std::shared_ptr<nanort::BVHAccel<T> > &acc = accels_[mesh];
      if (!acc) {
        acc.reset(new nanort::BVHAccel<T>());
        nanort::TriangleMesh<T> tm(mesh->vertices.data(), mesh->faces.data(), mesh->stride);
        nanort::TriangleSAHPred<T> pred(mesh->vertices.data(), mesh->faces.data(), mesh->stride);
        if (!acc->Build(static_cast<unsigned int>(mesh->faces.size()) / 3, tm, pred)) return false;
      }
      inst[i].accel = acc->NativeHandle();
      memcpy(inst[i].xform, nodes_[i].local_xform_, sizeof(float) * 16);
    }
    nrt_scene *s = NULL;
    if (nrt_scene_commit(inst.data(), static_cast<uint32_t>(inst.size()), NANORT_B200_BUILD_FLAGS, &s) != NRT_OK) {
      fprintf(stderr, "nanort_b200: Commit failed: %s\n", nrt_last_error());
      return false;
    }
    scene_ = std::shared_ptr<nrt_scene>(s, nrt_scene_free);
    for (size_t i = 0; i < nodes_.size(); i++) {  // Node::Update's results, for callers that read them
      float st[76];
      nrt_scene_instance_state(s, static_cast<uint32_t>(i), st);
      Node<T, M> &n = nodes_[i];
      memcpy(n.xform_, st, 64);
      memcpy(n.inv_xform_, st + 16, 64);
      memcpy(n.inv_xform33_, st + 32, 64);
      memcpy(n.inv_transpose_xform33_, st + 48, 64);
      memcpy(n.lbmin_, st + 64, 12);
      memcpy(n.lbmax_, st + 67, 12);
      memcpy(n.xbmin_, st + 70, 12);
      memcpy(n.xbmax_, st + 73, 12);
    }
    nrt_scene_bounding_box(s, bmin_, bmax_);
    return true;
  }

  void GetBoundingBox(T bmin[3], T bmax[3]) const {
    for (int k = 0; k < 3; k++) {
      bmin[k] = bmin_[k];
      bmax[k] = bmax_[k];
    }
  }

  /// One ray, synchronously (a full host<->device round trip per call; batch with TraverseBatch).  `cull_back_face`
  /// is accepted and, exactly like in the reference (:800-829), has no effect.
  template <class H, class I>
  bool Traverse(nanort::Ray<T> &ray, H *isect, const bool cull_back_face = false) const {
    (void)cull_back_face;
    if (!scene_) return false;
    nrt_scene_hit rec;
    unsigned char hit = 0;
    if (nrt_scene_traverse(scene_.get(), &ray, 1, &rec, &hit, NANORT_B200_TRAVERSE_FLAGS) != NRT_OK) {
      fprintf(stderr, "nanort_b200: Scene::Traverse failed: %s\n", nrt_last_error());
      return false;
    }
    if (hit) Fill(rec, isect);
    return hit != 0;
  }

  /// Extension: n rays at once; hits[i] is valid where hit_mask[i] != 0.  Returns the number of hits or (size_t)-1.
  size_t TraverseBatch(const nanort::Ray<T> *rays, size_t n, nrt_scene_hit *hits, unsigned char *hit_mask,
                       unsigned int flags = NANORT_B200_TRAVERSE_FLAGS) const {
    if (!scene_) return static_cast<size_t>(-1);
    std::vector<unsigned char> tmp;
    if (!hit_mask) {
      tmp.resize(n);
      hit_mask = tmp.data();
    }
    if (nrt_scene_traverse(scene_.get(), rays, n, hits, hit_mask, flags) != NRT_OK) {
      fprintf(stderr, "nanort_b200: Scene::TraverseBatch failed: %s\n", nrt_last_error());
      return static_cast<size_t>(-1);
    }
    size_t c = 0;
    for (size_t i = 0; i < n; i++) c += hit_mask[i] ? 1 : 0;
    return c;
  }

  /// The rest of Scene::Traverse for one record: normals from the node's mesh, moved to world space (:857-867).
  template <class H>
  void Fill(const nrt_scene_hit &rec, H *isect) const {
    const Node<T, M> &node = nodes_[rec.node_id];
    isect->t = rec.t;
    isect->prim_id = rec.prim_id;
    isect->u = rec.u;
    isect->v = rec.v;
    isect->node_id = rec.node_id;
    T Ng[3], Ns[3], w[3];
    node.GetMesh()->GetNormal(Ng, Ns, rec.prim_id, rec.u, rec.v);
    for (int k = 0; k < 3; k++) isect->P[k] = rec.P[k];
    detail::MultV(w, node.inv_transpose_xform33_, Ng);
    for (int k = 0; k < 3; k++) isect->Ng[k] = w[k];
    detail::MultV(w, node.inv_transpose_xform33_, Ns);
    for (int k = 0; k < 3; k++) isect->Ns[k] = w[k];
  }

 private:
  bool FindRecursive(const std::string &name, Node<T, M> *root, Node<T, M> **found) {
    if (root->GetName().compare(name) == 0) {
      *found = root;
      return true;
    }
    for (size_t i = 0; i < root->GetChildren().size(); i++)
      if (FindRecursive(name, &root->GetChildren()[i], found)) return true;
    return false;
  }

  T bmin_[3], bmax_[3];
  std::vector<Node<T, M> > nodes_;
  std::map<const M *, std::shared_ptr<nanort::BVHAccel<T> > > accels_;
  std::shared_ptr<nrt_scene> scene_;
};

}  // namespace nanosg

#endif  // NANOSG_H_
