// TEST INFRASTRUCTURE ONLY -- the reference's own custom-primitive model behind a C interface.
//
// Compiles the UNMODIFIED /root/reference/examples/particle_primitive/main.cc (its main() renamed away; nothing of it is
// copied into the repository) together with the unmodified nanort.h, and drives the example's SphereGeometry, SpherePred
// and SphereIntersector<SphereIntersection> (main.cc:80-291) through BVHAccel<float>::Build / Traverse exactly as the
// example's main() does (main.cc:338-401).  This is the oracle of the NRT_PRIM_SPHERES kind (csrc/prims.cu).
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

#define main nanort_reference_particle_primitive_main
#include "main.cc"  // found through -I/root/reference/examples/particle_primitive (oracle/Makefile)
#undef main

struct RefSpheres {
  std::vector<float> centers, radii;
  nanort::BVHAccel<float> accel;
};

extern "C" {

void *refsph_build(const float *centers, const float *radii, size_t n) {
  RefSpheres *s = new RefSpheres();
  s->centers.assign(centers, centers + 3 * n);
  s->radii.assign(radii, radii + n);
  nanort::BVHBuildOptions<float> options;  // default options, as the example
  SphereGeometry sphere_geom(&s->centers.at(0), &s->radii.at(0));
  SpherePred sphere_pred(&s->centers.at(0));
  if (!s->accel.Build(static_cast<unsigned int>(n), sphere_geom, sphere_pred, options)) {
    delete s;
    return NULL;
  }
  return s;
}

void refsph_free(void *h) { delete static_cast<RefSpheres *>(h); }

void refsph_bounding_box(const void *h, float bmin[3], float bmax[3]) {
  static_cast<const RefSpheres *>(h)->accel.BoundingBox(bmin, bmax);
}

// hits: {u, v, t, prim_id} per ray, untouched (zeroed by the caller) on a miss; trace options: prim id range only
void refsph_traverse(const void *h, const void *rays36, size_t n_rays, void *hits16, uint8_t *mask, uint32_t prim_lo,
                     uint32_t prim_hi, int threads) {
  const RefSpheres *s = static_cast<const RefSpheres *>(h);
  const nanort::Ray<float> *rays = static_cast<const nanort::Ray<float> *>(rays36);
  SphereIntersection *out = static_cast<SphereIntersection *>(hits16);
  static_assert(sizeof(SphereIntersection) == 16, "SphereIntersection is {u, v, t, prim_id}");
  nanort::BVHTraceOptions opt;
  opt.prim_ids_range[0] = prim_lo;
  opt.prim_ids_range[1] = prim_hi;
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([=]() {
      for (size_t i = (size_t)t; i < n_rays; i += (size_t)threads) {
        SphereIntersector<SphereIntersection> isecter(&s->centers.at(0), &s->radii.at(0));
        SphereIntersection isect;
        memset(&isect, 0, sizeof(isect));
        const bool hit = s->accel.Traverse(rays[i], isecter, &isect, opt);
        mask[i] = hit ? 1 : 0;
        if (hit) out[i] = isect;
      }
    });
  }
  for (auto &th : pool) th.join();
}

}  // extern "C"
