// The reference's Traverse is called concurrently from worker threads, each with its own intersector
// (examples/path_tracer/main.cc:787-799, 851-854).  The facade must allow exactly that: 8 threads pull rows from
// an atomic counter and call accel.Traverse per ray; the result must equal one batched call.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <thread>
#include <vector>

#include "nanort.h"

int main() {
  const int grid = 60, W = 96, H = 64;
  std::vector<float> verts;
  std::vector<unsigned int> faces;
  for (int z = 0; z <= grid; z++)
    for (int x = 0; x <= grid; x++) {
      float fx = float(x) / grid * 10.0f - 5.0f, fz = float(z) / grid * 10.0f - 5.0f;
      verts.push_back(fx);
      verts.push_back(0.6f * std::sin(fx * 1.3f) * std::cos(fz * 0.9f));
      verts.push_back(fz);
    }
  for (int z = 0; z < grid; z++)
    for (int x = 0; x < grid; x++) {
      unsigned int a = z * (grid + 1) + x, b = a + 1, c = a + grid + 2, d = a + grid + 1;
      faces.push_back(a); faces.push_back(c); faces.push_back(b);
      faces.push_back(a); faces.push_back(d); faces.push_back(c);
    }
  nanort::TriangleMesh<float> mesh(verts.data(), faces.data(), sizeof(float) * 3);
  nanort::TriangleSAHPred<float> pred(verts.data(), faces.data(), sizeof(float) * 3);
  nanort::BVHAccel<float> accel;
  if (!accel.Build((unsigned int)(faces.size() / 3), mesh, pred)) return 1;

  std::vector<nanort::Ray<float> > rays(size_t(W) * H);
  for (int y = 0; y < H; y++)
    for (int x = 0; x < W; x++) {
      nanort::Ray<float> &r = rays[size_t(y) * W + x];
      r.org[0] = 0.0f; r.org[1] = 4.0f; r.org[2] = 9.0f;
      float d[3] = {(x + 0.5f) / W - 0.5f, 0.1f - (y + 0.5f) / H, -1.0f};
      float l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
      for (int k = 0; k < 3; k++) r.dir[k] = d[k] / l;
      r.min_t = 1e-3f; r.max_t = 1e30f;
    }
  nanort::TriangleIntersector<> isector(verts.data(), faces.data(), sizeof(float) * 3);
  std::vector<nanort::TriangleIntersection<float> > batch(rays.size());
  std::vector<unsigned char> bmask(rays.size());
  accel.TraverseBatch(rays.data(), rays.size(), isector, batch.data(), bmask.data());

  std::vector<nanort::TriangleIntersection<float> > per(rays.size());
  std::vector<unsigned char> pmask(rays.size(), 0);
  std::atomic<int> row(0);
  std::vector<std::thread> workers;
  for (int t = 0; t < 8; t++)
    workers.emplace_back([&]() {
      int y;
      while ((y = row++) < H) {
        for (int x = 0; x < W; x++) {
          size_t i = size_t(y) * W + x;
          nanort::TriangleIntersector<> local(verts.data(), faces.data(), sizeof(float) * 3);  // one per ray, as main.cc:851
          nanort::TriangleIntersection<> isect;
          pmask[i] = accel.Traverse(rays[i], local, &isect) ? 1 : 0;
          if (pmask[i]) per[i] = isect;
        }
      }
    });
  for (auto &w : workers) w.join();
  size_t bad = 0, hits = 0;
  for (size_t i = 0; i < rays.size(); i++) {
    hits += bmask[i];
    if (bmask[i] != pmask[i]) bad++;
    else if (bmask[i] && (batch[i].t != per[i].t || batch[i].u != per[i].u || batch[i].v != per[i].v ||
                          batch[i].prim_id != per[i].prim_id)) bad++;
  }
  printf("rays %zu hits %zu mismatches %zu\n", rays.size(), hits, bad);
  return bad == 0 && hits > 0 ? 0 : 1;
}
