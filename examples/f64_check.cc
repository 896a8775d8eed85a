// Drop-in check for BVHAccel<double>: compiles unchanged against the reference's nanort.h and against
// include/nanort.h.  Scenario of the reference's regression program (one triangle, a ray whose x direction is
// 0 or -5.3e-17; expected u = 0.68, v = 0.131201) plus a fan of 2,000 double-precision rays over a small mesh.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <vector>

#include "nanort.h"

static unsigned long long Bits(double d) {
  unsigned long long u;
  memcpy(&u, &d, 8);
  return u;
}

int main() {
  typedef nanort::TriangleIntersection<double> Hit;
  {
    const double v[9] = {1, 2, -3, -1, 2, -3, 1, 2, 3};
    const unsigned int f[3] = {0, 1, 2};
    nanort::TriangleMesh<double> mesh(v, f, sizeof(double) * 3);
    nanort::TriangleSAHPred<double> pred(v, f, sizeof(double) * 3);
    nanort::BVHAccel<double> accel;
    if (!accel.Build(1, mesh, pred)) return 1;
    const double dxs[2] = {0.0, -5.30287619e-17};
    for (int k = 0; k < 2; k++) {
      nanort::Ray<double> ray;
      ray.org[0] = -0.36, ray.org[1] = 7.93890843, ray.org[2] = 1.2160368;
      double d[3] = {dxs[k], -8.66025404e-01, -0.5};
      const double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      for (int c = 0; c < 3; c++) ray.dir[c] = d[c] / len;
      ray.min_t = 0.0;
      ray.max_t = 1.0e30;
      nanort::TriangleIntersector<double, Hit> isector(v, f, sizeof(double) * 3);
      Hit h;
      const bool hit = accel.Traverse(ray, isector, &h);
      printf("regression %d: hit %d t %016llx u %016llx v %016llx (u %.6f v %.6f)\n", k, (int)hit, Bits(h.t), Bits(h.u),
             Bits(h.v), h.u, h.v);
    }
  }
  {
    // a bumpy 24 x 24 height field with coordinates that need double
    const int N = 24;
    std::vector<double> v;
    std::vector<unsigned int> f;
    for (int j = 0; j <= N; j++)
      for (int i = 0; i <= N; i++) {
        v.push_back(i * (1.0 / 3.0) + 1e-9 * j);
        v.push_back(0.25 * std::sin(0.7 * i) * std::cos(0.9 * j) + 1e-10 * i);
        v.push_back(j * (1.0 / 7.0) * 2.0);
      }
    for (int j = 0; j < N; j++)
      for (int i = 0; i < N; i++) {
        const unsigned a = j * (N + 1) + i, b = a + 1, c = a + N + 1, d = c + 1;
        const unsigned t[6] = {a, c, b, b, c, d};
        f.insert(f.end(), t, t + 6);
      }
    nanort::TriangleMesh<double> mesh(v.data(), f.data(), sizeof(double) * 3);
    nanort::TriangleSAHPred<double> pred(v.data(), f.data(), sizeof(double) * 3);
    nanort::BVHAccel<double> accel;
    if (!accel.Build((unsigned)f.size() / 3, mesh, pred)) return 1;
    double bmin[3], bmax[3];
    accel.BoundingBox(bmin, bmax);
    printf("bbox %016llx %016llx %016llx %016llx %016llx %016llx\n", Bits(bmin[0]), Bits(bmin[1]), Bits(bmin[2]),
           Bits(bmax[0]), Bits(bmax[1]), Bits(bmax[2]));
    unsigned hits = 0;
    for (int k = 0; k < 2000; k++) {
      nanort::Ray<double> ray;
      ray.org[0] = 4.0 + 3.5 * std::sin(0.37 * k), ray.org[1] = 3.0, ray.org[2] = 3.4 + 3.0 * std::cos(0.11 * k);
      double d[3] = {0.3 * std::sin(1.3 * k), -1.0, 0.3 * std::cos(0.7 * k)};
      const double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      for (int c = 0; c < 3; c++) ray.dir[c] = d[c] / len;
      ray.min_t = 0.0;
      ray.max_t = 1.0e30;
      nanort::TriangleIntersector<double, Hit> isector(v.data(), f.data(), sizeof(double) * 3);
      Hit h;
      if (accel.Traverse(ray, isector, &h)) {
        hits++;
        printf("%d: t %016llx u %016llx v %016llx\n", k, Bits(h.t), Bits(h.u), Bits(h.v));
      }
    }
    printf("hits %u\n", hits);
  }
  return 0;
}
