// BVHAccel::Dump / Load interchange (the reference's raw format, nanort.h:2164-2276): compiled against the
// reference header (bin/dump_load_ref, CPU) and against include/nanort.h (bin/dump_load_b200, GPU).
//   dump_load_X dump FILE      build the test mesh, Dump the tree to FILE
//   dump_load_X load FILE      Load the tree from FILE, traverse the test rays, print the hits
// tests/test_gpu_dropin.py crosses them: a GPU-built tree traversed by CPU nanort and a CPU-built tree
// traversed on the GPU must both print what CPU nanort prints for its own tree.
#define NANORT_ENABLE_SERIALIZATION
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "nanort.h"

static unsigned int lcg(unsigned int &s) {
  s = s * 1664525u + 1013904223u;
  return s;
}
static float frand(unsigned int &s) { return float(lcg(s) >> 8) * (1.0f / 16777216.0f); }

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  const int grid = 40;
  std::vector<float> verts;
  std::vector<unsigned int> faces;
  unsigned int seed = 777u;
  for (int z = 0; z <= grid; z++)
    for (int x = 0; x <= grid; x++) {
      verts.push_back(float(x) / grid * 8.0f - 4.0f);
      verts.push_back(0.8f * frand(seed));
      verts.push_back(float(z) / grid * 8.0f - 4.0f);
    }
  for (int z = 0; z < grid; z++)
    for (int x = 0; x < grid; x++) {
      unsigned int a = z * (grid + 1) + x, b = a + 1, c = a + grid + 2, d = a + grid + 1;
      faces.push_back(a); faces.push_back(c); faces.push_back(b);
      faces.push_back(a); faces.push_back(d); faces.push_back(c);
    }
  nanort::BVHAccel<float> accel;
  if (!strcmp(argv[1], "dump")) {
    nanort::TriangleMesh<float> mesh(verts.data(), faces.data(), sizeof(float) * 3);
    nanort::TriangleSAHPred<float> pred(verts.data(), faces.data(), sizeof(float) * 3);
    if (!accel.Build((unsigned int)(faces.size() / 3), mesh, pred)) return 1;
    if (!accel.Dump(argv[2])) return 1;
    printf("dumped %zu nodes %zu indices\n", accel.GetNodes().size(), accel.GetIndices().size());
    return 0;
  }
  if (!accel.Load(argv[2])) {
    fprintf(stderr, "Load failed\n");
    return 1;
  }
  nanort::TriangleIntersector<> isector(verts.data(), faces.data(), sizeof(float) * 3);
  int hits = 0;
  for (int i = 0; i < 300; i++) {
    nanort::Ray<float> ray;
    ray.org[0] = (frand(seed) - 0.5f) * 7.0f;
    ray.org[1] = 2.5f + frand(seed);
    ray.org[2] = (frand(seed) - 0.5f) * 7.0f;
    float d[3] = {frand(seed) - 0.5f, -1.0f, frand(seed) - 0.5f};
    float l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    for (int k = 0; k < 3; k++) ray.dir[k] = d[k] / l;
    ray.min_t = 0.001f;
    ray.max_t = 1.0e+30f;
    nanort::TriangleIntersection<> isect;
    isect.t = -1.0f; isect.u = isect.v = -1.0f; isect.prim_id = 4242u;
    bool hit = accel.Traverse(ray, isector, &isect);
    hits += hit;
    printf("%d %d %u %.9g %.9g %.9g\n", i, hit ? 1 : 0, isect.prim_id, isect.t, isect.u, isect.v);
  }
  printf("hits %d\n", hits);
  return 0;
}
