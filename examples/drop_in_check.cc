// Source-level drop-in check: this file uses ONLY the API that lighttransport/nanort's header and
// include/nanort.h (this repository) have in common -- the calls of examples/path_tracer/main.cc:742-763
// (Build) and :839-854 (per-ray Traverse).  examples/Makefile compiles it twice:
//   bin/drop_in_check_b200  against include/nanort.h      (GPU, libnanort_b200.so)
//   bin/drop_in_check_ref   against /root/reference/nanort.h (CPU reference; only where that tree exists)
// Both print one line per ray; tests/test_gpu_dropin.py diffs the outputs.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

#include "nanort.h"

static unsigned int lcg(unsigned int &s) {
  s = s * 1664525u + 1013904223u;
  return s;
}
static float frand(unsigned int &s) { return float(lcg(s) >> 8) * (1.0f / 16777216.0f); }

int main(int argc, char **argv) {
  const int grid = argc > 1 ? atoi(argv[1]) : 24;   // grid x grid bumpy quads
  const int n_rays = argc > 2 ? atoi(argv[2]) : 400;
  std::vector<float> verts;
  std::vector<unsigned int> faces;
  unsigned int seed = 12345u;
  for (int z = 0; z <= grid; z++)
    for (int x = 0; x <= grid; x++) {
      verts.push_back(float(x) / grid * 10.0f - 5.0f);
      verts.push_back(0.6f * frand(seed));
      verts.push_back(float(z) / grid * 10.0f - 5.0f);
    }
  for (int z = 0; z < grid; z++)
    for (int x = 0; x < grid; x++) {
      unsigned int a = z * (grid + 1) + x, b = a + 1, c = a + grid + 2, d = a + grid + 1;
      faces.push_back(a); faces.push_back(c); faces.push_back(b);
      faces.push_back(a); faces.push_back(d); faces.push_back(c);
    }
  const unsigned int n_faces = (unsigned int)(faces.size() / 3);

  nanort::BVHBuildOptions<float> build_options;  // defaults
  build_options.cache_bbox = false;
  nanort::TriangleMesh<float> triangle_mesh(verts.data(), faces.data(), sizeof(float) * 3);
  nanort::TriangleSAHPred<float> triangle_pred(verts.data(), faces.data(), sizeof(float) * 3);
  nanort::BVHAccel<float> accel;
  bool ret = accel.Build(n_faces, triangle_mesh, triangle_pred, build_options);
  if (!ret) {
    fprintf(stderr, "Build failed\n");
    return 1;
  }
  nanort::BVHBuildStatistics stats = accel.GetStatistics();
  float bmin[3], bmax[3];
  accel.BoundingBox(bmin, bmax);
  printf("faces %u leaves-branches %d valid %d bbox %.9g %.9g %.9g %.9g %.9g %.9g\n", n_faces,
         int(stats.num_leaf_nodes) - int(stats.num_branch_nodes), accel.IsValid() ? 1 : 0, bmin[0], bmin[1], bmin[2],
         bmax[0], bmax[1], bmax[2]);

#ifdef PRINT_TREE
  // conformance builds must agree on the whole tree, not only on the hits
  {
    const std::vector<nanort::BVHNode<float> > &nodes = accel.GetNodes();
    const std::vector<unsigned int> &indices = accel.GetIndices();
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < nodes.size(); i++) {
      const nanort::BVHNode<float> &nd = nodes[i];
      unsigned int w[10];
      memcpy(w, nd.bmin, 12);
      memcpy(w + 3, nd.bmax, 12);
      w[6] = (unsigned int)nd.flag;
      w[7] = nd.flag ? 0u : (unsigned int)nd.axis;  // leaf.axis is uninitialised in the reference
      w[8] = nd.data[0];
      w[9] = nd.data[1];
      for (int k = 0; k < 10; k++) h = (h ^ w[k]) * 1099511628211ull;
    }
    for (size_t i = 0; i < indices.size(); i++) h = (h ^ indices[i]) * 1099511628211ull;
    printf("tree nodes %zu leaves %u branches %u depth %u hash %016llx\n", nodes.size(), stats.num_leaf_nodes,
           stats.num_branch_nodes, stats.max_tree_depth, h);
  }
#endif
  int hits = 0;
  for (int i = 0; i < n_rays; i++) {
    nanort::Ray<float> ray;
    ray.org[0] = (frand(seed) - 0.5f) * 8.0f;
    ray.org[1] = 3.0f + frand(seed);
    ray.org[2] = (frand(seed) - 0.5f) * 8.0f;
    float d[3] = {frand(seed) - 0.5f, -1.0f, frand(seed) - 0.5f};
    float l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    ray.dir[0] = d[0] / l;
    ray.dir[1] = d[1] / l;
    ray.dir[2] = d[2] / l;
    ray.min_t = 0.001f;
    ray.max_t = (i % 5 == 0) ? 3.2f : 1.0e+30f;
    nanort::TriangleIntersector<> triangle_intersector(verts.data(), faces.data(), sizeof(float) * 3);
    nanort::TriangleIntersection<> isect;
    isect.t = -1.0f;
    isect.u = isect.v = -1.0f;
    isect.prim_id = 7777777u;  // must stay untouched on a miss
    nanort::BVHTraceOptions trace_options;
    if (i % 7 == 0) trace_options.cull_back_face = true;
    bool hit = accel.Traverse(ray, triangle_intersector, &isect, trace_options);
    hits += hit ? 1 : 0;
    printf("%d %d %u %.9g %.9g %.9g\n", i, hit ? 1 : 0, isect.prim_id, isect.t, isect.u, isect.v);
  }
  printf("hits %d\n", hits);
  return 0;
}
