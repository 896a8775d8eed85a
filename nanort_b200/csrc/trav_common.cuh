// Per-ray device helpers shared by the traversal kernels (traverse.cu) and the two-level scene kernels
// (scene.cu): ray constants, the slab test and the watertight triangle test, all in the reference's
// arithmetic order (compiled with --fmad=false).
#pragma once
#include <float.h>
#include <math_constants.h>

#include "common.cuh"

namespace nrt {

#define FULL_MASK 0xFFFFFFFFu

// ------------------------------------------------------------------ per-ray constants
struct RayCtx {
  float ox, oy, oz;
  float ix, iy, iz;  // vsafe_inverse(dir)
  float Sx, Sy, Sz;  // watertight shear constants
  float t_min;
  int sx, sy, sz;    // dir < 0
  int kx, ky, kz;
};

__device__ __forceinline__ float safe_inverse(float d, bool cpp03) {
  if (fabsf(d) < FLT_EPSILON) {
    // C++11 mode: copysign(1, d) -> -0.0f gives -inf; C++03 mode: (d < 0) ? -1 : 1 -> -0.0f gives +inf
    bool neg = cpp03 ? (d < 0.0f) : (__float_as_uint(d) >> 31) != 0u;
    return neg ? -CUDART_INF_F : CUDART_INF_F;
  }
  return 1.0f / d;
}

__device__ __forceinline__ float sel3(int k, float x, float y, float z) {
  return k == 0 ? x : (k == 1 ? y : z);
}

__device__ __forceinline__ void setup_ray(RayCtx &c, float ox, float oy, float oz, float dx, float dy,
                                          float dz, float min_t, bool cpp03) {
  c.ox = ox;
  c.oy = oy;
  c.oz = oz;
  c.sx = dx < 0.0f;
  c.sy = dy < 0.0f;
  c.sz = dz < 0.0f;
  c.ix = safe_inverse(dx, cpp03);
  c.iy = safe_inverse(dy, cpp03);
  c.iz = safe_inverse(dz, cpp03);
  int kz = 0;
  float m = fabsf(dx);
  if (m < fabsf(dy)) {
    kz = 1;
    m = fabsf(dy);
  }
  if (m < fabsf(dz)) kz = 2;
  int kx = (kz == 2) ? 0 : kz + 1;
  int ky = (kx == 2) ? 0 : kx + 1;
  float dkz = sel3(kz, dx, dy, dz);
  if (dkz < 0.0f) {
    int t = kx;
    kx = ky;
    ky = t;
  }
  c.kx = kx;
  c.ky = ky;
  c.kz = kz;
  c.Sx = sel3(kx, dx, dy, dz) / dkz;
  c.Sy = sel3(ky, dx, dy, dz) / dkz;
  c.Sz = 1.0f / dkz;
  c.t_min = min_t;
}

// A NaN in min_t or max_t poisons the reference's safemax / safemin chains (the range value sits in the slot whose
// NaN is NOT dropped, nanort.h:2316-2321): every slab test fails, the ray misses everything.  The kernels' fmaxf /
// fminf would drop that NaN, so such rays are retired before they start.
__device__ __forceinline__ bool range_has_nan(float min_t, float max_t) { return (min_t != min_t) | (max_t != max_t); }

// Slab test of one box (nanort.h:2284-2325).  fmaxf/fminf drop a NaN operand
// exactly like the reference's safemax/safemin do for the per-axis value in
// the first slot; the running value is never NaN (range_has_nan() rays never get here).
__device__ __forceinline__ bool slab(const RayCtx &c, float lox, float loy, float loz, float hix,
                                     float hiy, float hiz, float min_t, float max_t, float &tnear) {
  float nx = c.sx ? hix : lox, fx = c.sx ? lox : hix;
  float ny = c.sy ? hiy : loy, fy = c.sy ? loy : hiy;
  float nz = c.sz ? hiz : loz, fz = c.sz ? loz : hiz;
  float tnx = (nx - c.ox) * c.ix;
  float tny = (ny - c.oy) * c.iy;
  float tnz = (nz - c.oz) * c.iz;
  float tfx = ((fx - c.ox) * c.ix) * 1.00000024f;
  float tfy = ((fy - c.oy) * c.iy) * 1.00000024f;
  float tfz = ((fz - c.oz) * c.iz) * 1.00000024f;
  float tmin = fmaxf(tnz, fmaxf(tny, fmaxf(tnx, min_t)));
  float tmax = fminf(tfz, fminf(tfy, fminf(tfx, max_t)));
  tnear = tmin;
  return tmin <= tmax;
}

struct Best {
  float t, u, v;
  uint32_t prim;
};

// Watertight ray/triangle test, arithmetic order of nanort.h:1073-1147.
// Accepts t_min <= tt <= best.t (ties replace, like the reference).
__device__ __forceinline__ bool tri_test(const RayCtx &c, const TraceOptions16 &opt, float4 a, float4 b,
                                         float4 cc, Best &best) {
  uint32_t prim = __float_as_uint(a.w);
  if (prim < opt.prim_ids_range[0] || prim >= opt.prim_ids_range[1]) return false;
  if (prim == opt.skip_prim_id) return false;
  float A0 = a.x - c.ox, A1 = a.y - c.oy, A2 = a.z - c.oz;
  float B0 = b.x - c.ox, B1 = b.y - c.oy, B2 = b.z - c.oz;
  float C0 = cc.x - c.ox, C1 = cc.y - c.oy, C2 = cc.z - c.oz;
  float Akz = sel3(c.kz, A0, A1, A2), Bkz = sel3(c.kz, B0, B1, B2), Ckz = sel3(c.kz, C0, C1, C2);
  float Ax = sel3(c.kx, A0, A1, A2) - c.Sx * Akz;
  float Ay = sel3(c.ky, A0, A1, A2) - c.Sy * Akz;
  float Bx = sel3(c.kx, B0, B1, B2) - c.Sx * Bkz;
  float By = sel3(c.ky, B0, B1, B2) - c.Sy * Bkz;
  float Cx = sel3(c.kx, C0, C1, C2) - c.Sx * Ckz;
  float Cy = sel3(c.ky, C0, C1, C2) - c.Sy * Ckz;
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {
    // exact products in binary64, one rounding in the subtraction, one in the narrowing
    U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
    V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
    W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
  }
  if (U < 0.0f || V < 0.0f || W < 0.0f) {
    if (opt.cull_back_face || U > 0.0f || V > 0.0f || W > 0.0f) return false;
  }
  float det = (U + V) + W;
  if (det == 0.0f) return false;
  float Az = c.Sz * Akz, Bz = c.Sz * Bkz, Cz = c.Sz * Ckz;
  float D = (U * Az + V * Bz) + W * Cz;
  float rcp = 1.0f / det;
  float tt = D * rcp;
  if (tt > best.t) return false;
  if (tt < c.t_min) return false;
  best.t = tt;
  best.u = V * rcp;
  best.v = W * rcp;
  best.prim = prim;
  return true;
}

__device__ __forceinline__ void write_result(Hit16 *hits, uint8_t *mask, size_t i, const Best &best,
                                             float max_t) {
  bool hit = best.t < max_t;  // a hit exactly at max_t is a miss (nanort.h:2552)
  float4 r;
  if (hit) {
    r = make_float4(best.u, best.v, best.t, __uint_as_float(best.prim));
  } else {
    r = make_float4(0.0f, 0.0f, max_t, __uint_as_float(0xFFFFFFFFu));
  }
  reinterpret_cast<float4 *>(hits)[i] = r;
  if (mask) mask[i] = hit ? 1 : 0;
}

// Branch-free form of tri_test for the fast kernels (same arithmetic, same acceptance rule).
__device__ __forceinline__ void tri_test2(const RayCtx &c, const TraceOptions16 &opt, float4 a, float4 b, float4 cc,
                                          Best &best) {
  const uint32_t prim = __float_as_uint(a.w);
  bool rej = (prim < opt.prim_ids_range[0]) | (prim >= opt.prim_ids_range[1]) | (prim == opt.skip_prim_id);
  const float A0 = a.x - c.ox, A1 = a.y - c.oy, A2 = a.z - c.oz;
  const float B0 = b.x - c.ox, B1 = b.y - c.oy, B2 = b.z - c.oz;
  const float C0 = cc.x - c.ox, C1 = cc.y - c.oy, C2 = cc.z - c.oz;
  const float Akz = sel3(c.kz, A0, A1, A2), Bkz = sel3(c.kz, B0, B1, B2), Ckz = sel3(c.kz, C0, C1, C2);
  const float Ax = sel3(c.kx, A0, A1, A2) - c.Sx * Akz;
  const float Ay = sel3(c.ky, A0, A1, A2) - c.Sy * Akz;
  const float Bx = sel3(c.kx, B0, B1, B2) - c.Sx * Bkz;
  const float By = sel3(c.ky, B0, B1, B2) - c.Sy * Bkz;
  const float Cx = sel3(c.kx, C0, C1, C2) - c.Sx * Ckz;
  const float Cy = sel3(c.ky, C0, C1, C2) - c.Sy * Ckz;
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {  // rare: exact edge / vertex hits
    U = (float)((double)Cx * (double)By - (double)Cy * (double)Bx);
    V = (float)((double)Ax * (double)Cy - (double)Ay * (double)Cx);
    W = (float)((double)Bx * (double)Ay - (double)By * (double)Ax);
  }
  const bool neg = (U < 0.0f) | (V < 0.0f) | (W < 0.0f);
  const bool pos = (U > 0.0f) | (V > 0.0f) | (W > 0.0f);
  rej |= neg & ((opt.cull_back_face != 0) | pos);
  const float det = (U + V) + W;
  rej |= (det == 0.0f);
  const float Az = c.Sz * Akz, Bz = c.Sz * Bkz, Cz = c.Sz * Ckz;
  const float D = (U * Az + V * Bz) + W * Cz;
  const float rcp = 1.0f / det;
  const float tt = D * rcp;
  rej |= (tt > best.t) | (tt < c.t_min);
  if (!rej) {
    best.t = tt;
    best.u = V * rcp;
    best.v = W * rcp;
    best.prim = prim;
  }
}

}  // namespace nrt
