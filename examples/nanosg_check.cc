// Drop-in check for the two-level scene: this file compiles UNCHANGED against the reference's
// examples/nanosg/nanosg.h (+ nanort.h) and against include/nanosg.h of this repository, and must print the same
// lines.  Scene: three procedural meshes, 40 nodes with translation / rotation / non-uniform scale, a 96 x 64 grid
// of unit-length camera rays plus rays from inside the scene.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "nanort.h"
#include "nanosg.h"

struct Mesh {
  std::vector<float> vertices;
  std::vector<unsigned int> faces;
  size_t stride;
  // TriangleIntersector(const M*) of the reference reads these three
  const float *GetVertices() const { return vertices.data(); }
  const unsigned int *GetFaces() const { return faces.data(); }
  size_t GetVertexStrideBytes() const { return stride; }
  void GetNormal(float Ng[3], float Ns[3], unsigned int f, float, float) const {
    const float *a = &vertices[3 * faces[3 * f]], *b = &vertices[3 * faces[3 * f + 1]], *c = &vertices[3 * faces[3 * f + 2]];
    const float e0[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e1[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
    Ng[0] = e0[1] * e1[2] - e0[2] * e1[1];
    Ng[1] = e0[2] * e1[0] - e0[0] * e1[2];
    Ng[2] = e0[0] * e1[1] - e0[1] * e1[0];
    Ns[0] = Ng[0];
    Ns[1] = Ng[1];
    Ns[2] = Ng[2];
  }
};

static Mesh Sphere(int n_lon, int n_lat, float r) {
  Mesh m;
  m.stride = sizeof(float) * 3;
  for (int j = 0; j <= n_lat; j++)
    for (int i = 0; i < n_lon; i++) {
      const float th = 3.14159265f * float(j) / float(n_lat), ph = 6.2831853f * float(i) / float(n_lon);
      m.vertices.push_back(r * std::sin(th) * std::cos(ph));
      m.vertices.push_back(r * std::cos(th));
      m.vertices.push_back(r * std::sin(th) * std::sin(ph));
    }
  for (int j = 0; j < n_lat; j++)
    for (int i = 0; i < n_lon; i++) {
      const unsigned a = j * n_lon + i, b = j * n_lon + (i + 1) % n_lon, c = a + n_lon, d = b + n_lon;
      if (j > 0) {
        m.faces.push_back(a);
        m.faces.push_back(b);
        m.faces.push_back(c);
      }
      if (j < n_lat - 1) {
        m.faces.push_back(b);
        m.faces.push_back(d);
        m.faces.push_back(c);
      }
    }
  return m;
}

static unsigned Bits(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}

static unsigned lcg(unsigned &s) {
  s = s * 1664525u + 1013904223u;
  return s >> 8;
}
static float rnd(unsigned &s) { return float(lcg(s)) / 16777216.0f; }

int main() {
  Mesh meshes[3] = {Sphere(12, 8, 1.0f), Sphere(20, 14, 0.7f), Sphere(5, 4, 1.3f)};
  nanosg::Scene<float, Mesh> scene;
  unsigned seed = 7;
  for (int k = 0; k < 40; k++) {
    nanosg::Node<float, Mesh> node(&meshes[k % 3]);
    const float yaw = 6.2831853f * rnd(seed), sx = 0.5f + rnd(seed), sy = 0.5f + rnd(seed), sz = 0.5f + rnd(seed);
    float x[4][4] = {{sx * std::cos(yaw), 0.0f, -sx * std::sin(yaw), 0.0f},
                     {0.0f, sy, 0.0f, 0.0f},
                     {sz * std::sin(yaw), 0.0f, sz * std::cos(yaw), 0.0f},
                     {12.0f * (rnd(seed) - 0.5f), 4.0f * (rnd(seed) - 0.5f), 12.0f * (rnd(seed) - 0.5f), 1.0f}};
    node.SetLocalXform(x);
    node.SetName("node" + std::to_string(k));
    scene.AddNode(node);
  }
  if (!scene.Commit()) {
    printf("commit failed\n");
    return 1;
  }
  float bmin[3], bmax[3];
  scene.GetBoundingBox(bmin, bmax);
  printf("bbox %08x %08x %08x %08x %08x %08x\n", Bits(bmin[0]), Bits(bmin[1]), Bits(bmin[2]), Bits(bmax[0]),
         Bits(bmax[1]), Bits(bmax[2]));
  const nanosg::Node<float, Mesh> &n7 = scene.GetNodes()[7];
  printf("node7 inv %08x %08x %08x world box %08x %08x\n", Bits(n7.inv_xform_[0][0]), Bits(n7.inv_xform_[3][1]),
         Bits(n7.inv_transpose_xform33_[2][0]), Bits(n7.GetXformPtr()[12]), Bits(n7.inv_xform33_[1][1]));
  size_t n_hit = 0;
  for (int pass = 0; pass < 2; pass++) {
    const int W = pass == 0 ? 96 : 48, H = pass == 0 ? 64 : 32;
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        nanort::Ray<float> ray;
        float d[3];
        if (pass == 0) {
          ray.org[0] = 0.0f, ray.org[1] = 3.0f, ray.org[2] = 16.0f;
          d[0] = (float(x) + 0.5f) / float(W) - 0.5f, d[1] = (float(y) + 0.5f) / float(H) - 0.6f, d[2] = -1.0f;
        } else {
          ray.org[0] = 10.0f * (rnd(seed) - 0.5f), ray.org[1] = 3.0f * (rnd(seed) - 0.5f), ray.org[2] = 10.0f * (rnd(seed) - 0.5f);
          d[0] = rnd(seed) - 0.5f, d[1] = rnd(seed) - 0.5f, d[2] = rnd(seed) - 0.5f;
        }
        const float len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        for (int k = 0; k < 3; k++) ray.dir[k] = d[k] / len;
        ray.min_t = 0.0f;
        ray.max_t = 1.0e+30f;
        nanosg::Intersection<float> isect;
        const bool hit = scene.Traverse<nanosg::Intersection<float>,
                                        nanort::TriangleIntersector<float, nanosg::Intersection<float> > >(ray, &isect);
        if (hit) {
          n_hit++;
          printf("%d %d %d: node %u prim %u t %08x u %08x v %08x P %08x %08x %08x Ng %08x %08x %08x\n", pass, y, x,
                 isect.node_id, isect.prim_id, Bits(isect.t), Bits(isect.u), Bits(isect.v), Bits(isect.P[0]),
                 Bits(isect.P[1]), Bits(isect.P[2]), Bits(isect.Ng[0]), Bits(isect.Ng[1]), Bits(isect.Ng[2]));
        }
      }
  }
  printf("hits %zu\n", n_hit);
  return 0;
}
