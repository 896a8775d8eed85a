// Wavefront use of the facade: the pixel -> sample loop of examples/path_tracer/main.cc:804-854 re-expressed
// bounce by bounce, so that each bounce is ONE BVHAccel::TraverseBatch call instead of one Traverse per ray.
// Renders primary + 1-bounce ambient occlusion of a bumpy heightfield to ao.ppm.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "nanort.h"

static unsigned int hash_u32(unsigned int x) {
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return x;
}
static float rnd(unsigned int pix, unsigned int smp, unsigned int dim) {
  unsigned int h = hash_u32(hash_u32(pix + 0x9E3779B1u) + smp * 0x85EBCA77u + dim * 0xC2B2AE3Du);
  return float(h >> 8) * (1.0f / 16777216.0f);
}

int main(int argc, char **argv) {
  const int W = argc > 1 ? atoi(argv[1]) : 512, H = argc > 2 ? atoi(argv[2]) : 512, SPP = argc > 3 ? atoi(argv[3]) : 8;
  const int grid = 200;
  std::vector<float> verts;
  std::vector<unsigned int> faces;
  for (int z = 0; z <= grid; z++)
    for (int x = 0; x <= grid; x++) {
      float fx = float(x) / grid * 10.0f - 5.0f, fz = float(z) / grid * 10.0f - 5.0f;
      verts.push_back(fx);
      verts.push_back(0.5f * std::sin(fx * 1.7f) * std::cos(fz * 1.3f) + 0.15f * std::sin(fx * 7.0f + fz * 5.0f));
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
  nanort::BVHBuildStatistics st = accel.GetStatistics();
  printf("built %zu triangles: %u leaves, depth %u, %.3f ms on the device\n", faces.size() / 3, st.num_leaf_nodes,
         st.max_tree_depth, st.build_secs * 1e3f);
  nanort::TriangleIntersector<> isector(verts.data(), faces.data(), sizeof(float) * 3);

  std::vector<float> image(size_t(W) * H, 0.0f);
  std::vector<nanort::Ray<float> > rays(size_t(W) * H), ao;
  std::vector<nanort::TriangleIntersection<float> > hits(rays.size()), ao_hits;
  std::vector<unsigned char> mask(rays.size()), ao_mask;
  std::vector<unsigned int> ao_pix;
  size_t total_rays = 0;
  for (int s = 0; s < SPP; s++) {
    // bounce 0: camera rays (examples/path_tracer/main.cc:809-817)
    for (int y = 0; y < H; y++)
      for (int x = 0; x < W; x++) {
        unsigned int pix = y * W + x;
        nanort::Ray<float> &r = rays[pix];
        r.org[0] = 0.0f; r.org[1] = 5.0f; r.org[2] = 9.0f;
        float d[3] = {(x + rnd(pix, s, 0)) / W - 0.5f, 0.5f - (y + rnd(pix, s, 1)) / H - 0.45f, -1.0f};
        float l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        r.dir[0] = d[0] / l; r.dir[1] = d[1] / l; r.dir[2] = d[2] / l;
        r.min_t = 1e-3f; r.max_t = 1e30f;
      }
    accel.TraverseBatch(rays.data(), rays.size(), isector, hits.data(), mask.data());
    // bounce 1: one cosine-weighted AO ray per hit (main.cc:860, 306-312, 216-250, 675-701)
    ao.clear(); ao_pix.clear();
    for (size_t i = 0; i < rays.size(); i++) {
      if (!mask[i]) { image[i] += 1.0f; continue; }
      const nanort::TriangleIntersection<float> &h = hits[i];
      const float *p0 = &verts[3 * faces[3 * h.prim_id]], *p1 = &verts[3 * faces[3 * h.prim_id + 1]],
                  *p2 = &verts[3 * faces[3 * h.prim_id + 2]];
      float e1[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]}, e2[3] = {p2[0] - p0[0], p2[1] - p0[1], p2[2] - p0[2]};
      float n[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
      float ln = std::sqrt(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
      for (int k = 0; k < 3; k++) n[k] /= ln;
      if (n[0] * rays[i].dir[0] + n[1] * rays[i].dir[1] + n[2] * rays[i].dir[2] > 0) for (int k = 0; k < 3; k++) n[k] = -n[k];
      float sg = n[2] >= 0 ? 1.0f : -1.0f, a = -1.0f / (sg + n[2]), b = n[0] * n[1] * a;
      float t1[3] = {1 + sg * n[0] * n[0] * a, sg * b, -sg * n[0]}, t2[3] = {b, sg + n[1] * n[1] * a, -n[1]};
      float u1 = rnd((unsigned int)i, s, 2), u2 = rnd((unsigned int)i, s, 3), rr = std::sqrt(u1), ph = 6.2831853f * u2;
      float lx = rr * std::cos(ph), ly = rr * std::sin(ph), lz = std::sqrt(1 - u1);
      nanort::Ray<float> r;
      for (int k = 0; k < 3; k++) {
        r.org[k] = rays[i].org[k] + rays[i].dir[k] * h.t;
        r.dir[k] = t1[k] * lx + t2[k] * ly + n[k] * lz;
      }
      r.min_t = 1e-3f; r.max_t = 2.5f;
      ao.push_back(r); ao_pix.push_back((unsigned int)i);
    }
    ao_hits.resize(ao.size()); ao_mask.resize(ao.size());
    if (!ao.empty()) accel.TraverseBatch(ao.data(), ao.size(), isector, ao_hits.data(), ao_mask.data());
    for (size_t j = 0; j < ao.size(); j++) if (!ao_mask[j]) image[ao_pix[j]] += 1.0f;
    total_rays += rays.size() + ao.size();
  }
  FILE *fp = fopen("ao.ppm", "wb");
  if (fp) {
    fprintf(fp, "P5\n%d %d\n255\n", W, H);
    for (size_t i = 0; i < image.size(); i++) fputc((int)(255.0f * std::pow(image[i] / SPP, 1.0f / 2.2f)), fp);
    fclose(fp);
  }
  printf("traced %zu rays in %d batched Traverse calls -> ao.ppm\n", total_rays, 2 * SPP);
  return 0;
}
