// Wavefront pieces shared by render.cu (stand-alone stage kernels) and traverse.cu (the same stages fused
// into the traversal kernel's retire step): ray queues, counter-based RNG, slot -> pixel mapping, AO ray
// construction.  Reference pieces restated here (file:line under /root/reference/examples/path_tracer):
//   hit point main.cc:860, geometric normal main.cc:306-312 flipped to the viewer main.cc:878-881,
//   orthonormal basis + cosine direction main.cc:216-250, occlusion query main.cc:675-701.
#pragma once
#include "common.cuh"

namespace nrt {

struct Wave {
  float4 *org_tmin;  // primary queue (SoA): org.xyz, min_t
  float4 *dir_tmax;  //                      dir.xyz, max_t
  Hit16 *hits;
  uint32_t *pix;        // pixel of a primary slot (0xFFFFFFFF = slot outside the image)
  float4 *ao_org_tmin;  // compacted AO queue
  float4 *ao_dir_tmax;
  uint32_t *ao_pix;
  Hit16 *ao_hits;
};

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {  // lowbias32, same as scenes.py:hash_u32
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}

// scenes.py:rand_ps
__device__ __forceinline__ float rand_ps(uint32_t pix, uint32_t smp, uint32_t dim, uint32_t seed) {
  uint32_t h = hash_u32(pix + seed * 0x9E3779B1u);
  h = hash_u32(h + smp * 0x85EBCA77u + dim * 0xC2B2AE3Du);
  return (float)(h >> 8) * (1.0f / 16777216.0f);
}

// The part of nrt_ao_params / nrt_path_params that maps ray slots to pixels.
struct TileMap {
  uint32_t width, height, spp, sample0, tile_w, tile_h, shard, n_shards;
  uint32_t packed;  // accumulate into a tile-major buffer of this shard's tiles (NRT_AO_PACKED_TILES) instead of the image
};
__host__ __device__ __forceinline__ TileMap tile_map(const nrt_ao_params &p) {
  return TileMap{p.width, p.height, p.spp, p.sample0, p.tile_w, p.tile_h, p.shard, p.n_shards,
                 (p.flags & NRT_AO_PACKED_TILES) ? 1u : 0u};
}
__host__ __device__ __forceinline__ TileMap tile_map(const nrt_path_params &p) {
  return TileMap{p.width, p.height, p.spp, p.sample0, p.tile_w, p.tile_h, p.shard, p.n_shards, 0u};
}

// Division of a 32-bit value by a run-time constant without a divide (Granlund-Montgomery round-up method, exact for
// every uint32 x): the slot -> pixel mapping of a camera ray needs four of them, and an integer division costs more
// instructions than the slab test of a node pair.
struct FastDiv {
  uint32_t d, mul, shift;  // shift == 0xFFFFFFFF: d is 1
  __host__ __device__ FastDiv() : d(1), mul(0), shift(0xFFFFFFFFu) {}
  __host__ explicit FastDiv(uint32_t dd) : d(dd ? dd : 1), mul(0), shift(0xFFFFFFFFu) {
    if (d > 1) {
      uint32_t l = 0;
      while ((1ull << l) < d) l++;  // ceil(log2 d) >= 1
      mul = (uint32_t)(((1ull << 32) * ((1ull << l) - d)) / d + 1ull);
      shift = l - 1;
    }
  }
  __device__ __forceinline__ uint32_t div(uint32_t x) const {
    if (shift == 0xFFFFFFFFu) return x;
    const uint32_t t = __umulhi(mul, x);
    return (t + ((x - t) >> 1)) >> shift;
  }
  __device__ __forceinline__ void divmod(uint32_t x, uint32_t &q, uint32_t &r) const {
    q = div(x);
    r = x - q * d;
  }
};

// Slot -> (pixel, sample).  Slots enumerate this shard's tiles; inside a tile the order is sample-major over
// 8x4 pixel blocks, so the 32 lanes of a warp start as one coherent 8x4 packet.
// `acc` = where the pixel's samples are accumulated: the pixel itself, or -- tile-major packing for the multi-GPU
// gather (comm.cu) -- k * tile_pixels + row-major offset inside the tile.
__device__ __forceinline__ bool slot_to_pixel(const TileMap &p, unsigned long long slot, uint32_t &pix,
                                              uint32_t &smp, uint32_t &acc) {
  const uint32_t tile_pix = p.tile_w * p.tile_h;
  const unsigned long long per_tile = (unsigned long long)tile_pix * p.spp;
  const uint32_t k = (uint32_t)(slot / per_tile);  // k-th tile of this shard
  const uint32_t rem = (uint32_t)(slot % per_tile);
  smp = rem / tile_pix;
  const uint32_t q = rem % tile_pix;
  const uint32_t bw = p.tile_w / 8;  // 8x4 blocks per tile row
  const uint32_t blk = q / 32, in = q % 32;
  const uint32_t bx = blk % bw, by = blk / bw;
  const uint32_t lx = bx * 8 + (in & 7), ly = by * 4 + (in >> 3);
  const uint32_t tiles_x = (p.width + p.tile_w - 1) / p.tile_w;
  const uint32_t tile = k * p.n_shards + p.shard;
  const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
  const uint32_t x = tx * p.tile_w + lx, y = ty * p.tile_h + ly;
  if (x >= p.width || y >= p.height) return false;
  pix = y * p.width + x;
  acc = p.packed ? k * tile_pix + ly * p.tile_w + lx : pix;
  return true;
}

__device__ __forceinline__ bool slot_to_pixel(const TileMap &p, unsigned long long slot, uint32_t &pix,
                                              uint32_t &smp) {
  uint32_t acc;
  return slot_to_pixel(p, slot, pix, smp, acc);
}

__device__ __forceinline__ bool slot_to_pixel(const nrt_ao_params &p, unsigned long long slot, uint32_t &pix,
                                              uint32_t &smp) {
  return slot_to_pixel(tile_map(p), slot, pix, smp);
}

__device__ __forceinline__ uint32_t slot_sample(const nrt_ao_params &p, unsigned long long slot) {
  const uint32_t tile_pix = p.tile_w * p.tile_h;
  return p.sample0 + (uint32_t)((slot % ((unsigned long long)tile_pix * p.spp)) / tile_pix);
}

// Jittered pinhole camera ray of (pixel, sample) -- examples/path_tracer/main.cc:809-817.  One definition for the
// stand-alone generator kernel, the in-kernel generator (CameraRays) and the AO epilogue, so that all of them
// produce bit-identical rays.
__device__ __forceinline__ void camera_ray(const float *cam, uint32_t width, uint32_t height, uint32_t seed,
                                           uint32_t pix, uint32_t smp, float &dx, float &dy, float &dz) {
  const float jx = rand_ps(pix, smp, 0, seed), jy = rand_ps(pix, smp, 1, seed);
  const float px = (float)(pix % width), py = (float)(pix / width);
  const float sx = (px + jx) / (float)width - 0.5f;
  const float sy = 0.5f - (py + jy) / (float)height;
  dx = cam[3] * sx + cam[6] * sy + cam[9];
  dy = cam[4] * sx + cam[7] * sy + cam[10];
  dz = cam[5] * sx + cam[8] * sy + cam[11];
  const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
  dx *= inv;
  dy *= inv;
  dz *= inv;
}

// Ray "loader" that generates the camera ray of slot (slot0 + i) instead of reading a queue: the primary
// traversal then needs no generator kernel and no primary ray queue at all.  The mapping is slot_to_pixel() with its
// divisions replaced by multiplications (slot0 is a multiple of the per-tile slot count: waves are whole tiles), and
// what the retire step needs again -- direction, pixel, sample, accumulation index -- travels as the ray's payload
// (6 words the kernel parks in thread-local memory) instead of being recomputed.
struct CameraRays {
  static constexpr int kPayloadWords = 6;
  nrt_ao_params p;
  uint32_t k0;  // slot0 / per_tile: ordinal (within the shard) of the wave's first tile
  FastDiv per_tile, tile_pix, bw, tiles_x;
  uint32_t packed;
  __host__ CameraRays(const nrt_ao_params &pp, unsigned long long slot0) : p(pp) {
    const uint32_t tp = pp.tile_w * pp.tile_h;
    per_tile = FastDiv(tp * pp.spp);
    tile_pix = FastDiv(tp);
    bw = FastDiv(pp.tile_w / 8);
    tiles_x = FastDiv((pp.width + pp.tile_w - 1) / pp.tile_w);
    k0 = (uint32_t)(slot0 / ((unsigned long long)tp * pp.spp));
    packed = (pp.flags & NRT_AO_PACKED_TILES) ? 1u : 0u;
  }
  __device__ __forceinline__ void load(size_t i, float &ox, float &oy, float &oz, float &dx, float &dy, float &dz,
                                       float &tmin, float &tmax, uint32_t *payload) const {
    uint32_t k, rem, smp, q, blk, bx, by, ty, tx;
    per_tile.divmod((uint32_t)i, k, rem);
    k += k0;
    tile_pix.divmod(rem, smp, q);
    blk = q >> 5;
    const uint32_t in = q & 31u;
    bw.divmod(blk, by, bx);
    const uint32_t lx = bx * 8 + (in & 7), ly = by * 4 + (in >> 3);
    tiles_x.divmod(k * p.n_shards + p.shard, ty, tx);
    const uint32_t x = tx * p.tile_w + lx, y = ty * p.tile_h + ly;
    ox = p.cam[0];
    oy = p.cam[1];
    oz = p.cam[2];
    if (x < p.width && y < p.height) {
      const uint32_t pix = y * p.width + x;
      smp += p.sample0;
      // camera_ray() with the pixel coordinates at hand (same arithmetic, same rays bit for bit)
      const float jx = rand_ps(pix, smp, 0, p.seed), jy = rand_ps(pix, smp, 1, p.seed);
      const float sx = ((float)x + jx) / (float)p.width - 0.5f;
      const float sy = 0.5f - ((float)y + jy) / (float)p.height;
      dx = p.cam[3] * sx + p.cam[6] * sy + p.cam[9];
      dy = p.cam[4] * sx + p.cam[7] * sy + p.cam[10];
      dz = p.cam[5] * sx + p.cam[8] * sy + p.cam[11];
      const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
      dx *= inv;
      dy *= inv;
      dz *= inv;
      tmin = p.ray_min_t;
      tmax = p.ray_max_t;
      payload[0] = __float_as_uint(dx);
      payload[1] = __float_as_uint(dy);
      payload[2] = __float_as_uint(dz);
      payload[3] = pix;
      payload[4] = smp;
      payload[5] = packed ? k * tile_pix.d + ly * p.tile_w + lx : pix;
    } else {  // slot outside the image: retires at the root as a miss
      dx = 0.0f;
      dy = 0.0f;
      dz = -1.0f;
      tmin = 0.0f;
      tmax = -1.0f;
      payload[3] = 0xFFFFFFFFu;
    }
  }
};

// One cosine-hemisphere AO ray from a primary hit.
__device__ __forceinline__ void make_ao_ray(const nrt_ao_params &p, uint32_t pix, uint32_t smp, float4 o, float4 d,
                                            float t, uint32_t prim, const float *__restrict__ verts,
                                            const uint32_t *__restrict__ faces, float4 &o4, float4 &d4) {
  const float Px = o.x + d.x * t, Py = o.y + d.y * t, Pz = o.z + d.z * t;
  const uint32_t f0 = faces[3 * (size_t)prim], f1 = faces[3 * (size_t)prim + 1], f2 = faces[3 * (size_t)prim + 2];
  const float *p0 = verts + 3 * (size_t)f0, *p1 = verts + 3 * (size_t)f1, *p2 = verts + 3 * (size_t)f2;
  const float e1x = p1[0] - p0[0], e1y = p1[1] - p0[1], e1z = p1[2] - p0[2];
  const float e2x = p2[0] - p0[0], e2y = p2[1] - p0[1], e2z = p2[2] - p0[2];
  float nx = e1y * e2z - e1z * e2y, ny = e1z * e2x - e1x * e2z, nz = e1x * e2y - e1y * e2x;
  float ln = sqrtf(nx * nx + ny * ny + nz * nz);
  ln = ln > 0.0f ? 1.0f / ln : 0.0f;
  nx *= ln;
  ny *= ln;
  nz *= ln;
  if (nx * d.x + ny * d.y + nz * d.z > 0.0f) {
    nx = -nx;
    ny = -ny;
    nz = -nz;
  }
  // branch-free orthonormal basis around n
  const float sg = nz >= 0.0f ? 1.0f : -1.0f;
  const float a = -1.0f / (sg + nz), b = nx * ny * a;
  const float t1x = 1.0f + sg * nx * nx * a, t1y = sg * b, t1z = -sg * nx;
  const float t2x = b, t2y = sg + ny * ny * a, t2z = -ny;
  const float u1 = rand_ps(pix, smp, 2, p.seed), u2 = rand_ps(pix, smp, 3, p.seed);
  const float r = sqrtf(u1), ph = 6.28318530718f * u2;
  float sn, cs;
  sincosf(ph, &sn, &cs);
  const float lx = r * cs, ly = r * sn, lz = sqrtf(fmaxf(0.0f, 1.0f - u1));
  const float wx = t1x * lx + t2x * ly + nx * lz, wy = t1y * lx + t2y * ly + ny * lz, wz = t1z * lx + t2z * ly + nz * lz;
  const float il = 1.0f / sqrtf(wx * wx + wy * wy + wz * wz);
  o4 = make_float4(Px, Py, Pz, p.ao_min_t);
  d4 = make_float4(wx * il, wy * il, wz * il, p.ao_max_t);
}

// ---- retire-step functors of traverse_fast2_kernel.  Called by ALL 32 lanes of a warp (`retiring` says
// whether this lane's ray just finished), so they may use full-mask warp votes.
struct StoreHitsEpilogue {
  static constexpr bool kAnyHit = false;  // true: the kernel retires a ray at its first hit inside [min_t, max_t)
  Hit16 *hits;
  uint8_t *mask;
  __device__ __forceinline__ void operator()(bool retiring, size_t ray_idx, float t, float u, float v, uint32_t prim,
                                             float max_t, const uint32_t *) const {
    if (retiring && hits) {
      const bool hit = t < max_t;  // a hit exactly at max_t is a miss (nanort.h:2552)
      float4 r = hit ? make_float4(u, v, t, __uint_as_float(prim)) : make_float4(0.0f, 0.0f, max_t, __uint_as_float(0xFFFFFFFFu));
      __stcs(reinterpret_cast<float4 *>(hits) + ray_idx, r);  // written once, never re-read by this kernel
      if (mask) mask[ray_idx] = hit ? 1 : 0;
    }
  }
};

// primary rays: a hit spawns its AO ray straight into the compacted AO queue, a miss adds 1 to its pixel.
// GEN: the primary rays were generated in the kernel (CameraRays) -- pixel and ray are recomputed from the slot.
template <bool GEN>
struct PrimaryToAoEpilogue {
  static constexpr bool kAnyHit = false;
  nrt_ao_params p;
  unsigned long long slot0;
  Wave w;
  const float *verts;
  const uint32_t *faces;
  float *accum;
  unsigned long long *counters;  // [0] AO rays of this wave
  __device__ __forceinline__ void operator()(bool retiring, size_t ray_idx, float t, float u, float v, uint32_t prim,
                                             float max_t, const uint32_t *payload) const {
    (void)u;
    (void)v;
    bool make = false;
    float4 o4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = o4;
    uint32_t pix = 0xFFFFFFFFu, acc = 0;
    if (retiring) {
      uint32_t smp = 0;
      float4 ro, rd;
      if (GEN) {  // the camera ray's payload (CameraRays::load): direction, pixel, sample, accumulation index
        pix = payload[3];
        if (pix != 0xFFFFFFFFu) {
          smp = payload[4];
          acc = payload[5];
          ro = make_float4(p.cam[0], p.cam[1], p.cam[2], p.ray_min_t);
          rd = make_float4(__uint_as_float(payload[0]), __uint_as_float(payload[1]), __uint_as_float(payload[2]),
                           p.ray_max_t);
        }
      } else {
        pix = w.pix[ray_idx];
        acc = pix;  // the unfused path does not pack (run_ao_pass rejects the combination)
        if (pix != 0xFFFFFFFFu) {
          smp = slot_sample(p, slot0 + ray_idx);
          ro = w.org_tmin[ray_idx];
          rd = w.dir_tmax[ray_idx];
        }
      }
      if (pix != 0xFFFFFFFFu) {
        if (t < max_t) {
          make_ao_ray(p, pix, smp, ro, rd, t, prim, verts, faces, o4, d4);
          make = true;
        } else {
          atomicAdd(accum + acc, 1.0f);
        }
      }
    }
    const unsigned m = __ballot_sync(0xFFFFFFFFu, make);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counters, (unsigned long long)__popc(m));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (make) {
      const unsigned long long j = base + __popc(m & ((1u << lane) - 1u));
      w.ao_org_tmin[j] = o4;
      w.ao_dir_tmax[j] = d4;
      w.ao_pix[j] = acc;  // the AO retire step only needs the accumulation index
    }
  }
};

// AO rays: an unoccluded ray adds 1 to its pixel; nothing else is written
struct AoAccumulateEpilogue {
  static constexpr bool kAnyHit = false;
  const uint32_t *ao_pix;
  float *accum;
  unsigned long long *totals;  // [1] occluded AO rays
  __device__ __forceinline__ void operator()(bool retiring, size_t ray_idx, float t, float u, float v, uint32_t prim,
                                             float max_t, const uint32_t *) const {
    (void)u;
    (void)v;
    (void)prim;
    const bool occluded = retiring && (t < max_t);
    if (retiring && !occluded) atomicAdd(accum + ao_pix[ray_idx], 1.0f);
    const unsigned m = __ballot_sync(0xFFFFFFFFu, occluded);
    if (m != 0u && (int)(threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(totals + 1, (unsigned long long)__popc(m));
  }
};

// ------------------------------------------------------------------ path tracing wavefront
struct PathQueues {
  // radiance queue in / out (ping-pong), SoA rays + the path (= primary slot of the wave) each ray belongs to
  float4 *org_tmin[2];
  float4 *dir_tmax[2];
  uint32_t *path_id[2];
  // shadow queue: ray + (contribution.rgb, pixel)
  float4 *sh_org_tmin;
  float4 *sh_dir_tmax;
  float4 *sh_contrib_pix;
  float4 *weight;  // per path: throughput rgb
};

__device__ __forceinline__ void geometric_normal(const float *__restrict__ verts, const uint32_t *__restrict__ faces,
                                                 uint32_t prim, float &nx, float &ny, float &nz, float &area2) {
  const uint32_t f0 = faces[3 * (size_t)prim], f1 = faces[3 * (size_t)prim + 1], f2 = faces[3 * (size_t)prim + 2];
  const float *p0 = verts + 3 * (size_t)f0, *p1 = verts + 3 * (size_t)f1, *p2 = verts + 3 * (size_t)f2;
  const float e1x = p1[0] - p0[0], e1y = p1[1] - p0[1], e1z = p1[2] - p0[2];
  const float e2x = p2[0] - p0[0], e2y = p2[1] - p0[1], e2z = p2[2] - p0[2];
  nx = e1y * e2z - e1z * e2y;
  ny = e1z * e2x - e1x * e2z;
  nz = e1x * e2y - e1y * e2x;
  area2 = sqrtf(nx * nx + ny * ny + nz * nz);
  const float il = area2 > 0.0f ? 1.0f / area2 : 0.0f;
  nx *= il;
  ny *= il;
  nz *= il;
}

// 16 floats per material, the tinyobj fields the reference reads (main.cc:884-892)
struct PathMaterial {
  float diffuse[3];
  float specular[3];
  float transmittance[3];
  float emission[3];
  float ior;
  float dissolve;
  float pad[2];
};

// Radiance rays of bounce `bounce`: the retire step is the reference's per-hit shading block
// (examples/path_tracer/main.cc:856-976): normal, material, Fresnel, lobe probabilities, lobe choice.
struct PathShadeEpilogue {
  static constexpr bool kAnyHit = false;
  nrt_path_params p;
  unsigned long long slot0;
  int in;  // which radiance queue is being traversed; (in ^ 1) receives the continuation rays
  uint32_t bounce;
  PathQueues q;
  const float *verts;
  const uint32_t *faces;
  float *accum;                  // rgb
  unsigned long long *counters;  // [0] continuation rays, [1] shadow rays of this bounce
  __device__ __forceinline__ void operator()(bool retiring, size_t ray_idx, float t, float u, float v, uint32_t prim,
                                             float max_t, const uint32_t *) const {
    bool cont = false, shadow = false;
    float4 co = make_float4(0, 0, 0, 0), cd = co, so = co, sd = co, sc = co;
    uint32_t pid = 0;
    if (retiring && t < max_t) {
      pid = q.path_id[in][ray_idx];
      uint32_t pix, smp;
      if (slot_to_pixel(tile_map(p), slot0 + pid, pix, smp)) {
        smp += p.sample0;
        const float4 o = q.org_tmin[in][ray_idx], d = q.dir_tmax[in][ray_idx];
        float4 w = q.weight[pid];  // throughput rgb, w.w = do_emission (no light sampling at the previous event)
        const PathMaterial *mats = reinterpret_cast<const PathMaterial *>(p.d_materials);
        const uint32_t *mat_ids = reinterpret_cast<const uint32_t *>(p.d_material_ids);
        const uint32_t *emissive = reinterpret_cast<const uint32_t *>(p.d_emissive_faces);
        const float *fv_normals = reinterpret_cast<const float *>(p.d_facevarying_normals);
        // ---- normal: interpolated face-varying normals when given (main.cc:862-875), else geometric
        float nx, ny, nz, a2;
        if (fv_normals) {
          const float *n0 = fv_normals + 9 * (size_t)prim;
          const float b0 = 1.0f - u - v;
          nx = b0 * n0[0] + u * n0[3] + v * n0[6];
          ny = b0 * n0[1] + u * n0[4] + v * n0[7];
          nz = b0 * n0[2] + u * n0[5] + v * n0[8];
          const float l = sqrtf(nx * nx + ny * ny + nz * nz);
          if (fabsf(l) > 1.0e-6f) {
            const float il = 1.0f / l;
            nx *= il;
            ny *= il;
            nz *= il;
          }
        } else {
          // no normals given: the flat normal the example's loader stores for such a mesh, calcNormal's
          // cross(v2 - v0, v1 - v0) (main.cc:306-312, 566-601) -- the opposite of cross(e1, e2)
          geometric_normal(verts, faces, prim, nx, ny, nz, a2);
          nx = -nx;
          ny = -ny;
          nz = -nz;
        }
        const float onx = nx, ony = ny, onz = nz;  // originalNorm
        const float ndotd = nx * d.x + ny * d.y + nz * d.z;
        if (ndotd > 0.0f) {  // flip towards the incoming ray (main.cc:878-881)
          nx = -nx;
          ny = -ny;
          nz = -nz;
        }
        const PathMaterial m = mats[mat_ids ? mat_ids[prim] : 0u];
        // ---- Fresnel and lobe probabilities (main.cc:894-929)
        const float inside = ndotd < 0.0f ? -1.0f : 1.0f;  // sign(dot(rayDir, originalNorm))
        const float n1 = inside < 0.0f ? 1.0f / m.ior : m.ior;
        const float n2 = 1.0f / n1;
        const float r0s = (n1 - n2) / (n1 + n2);
        const float r0 = r0s * r0s;
        const float hdn = 1.0f - (-(d.x * nx + d.y * ny + d.z * nz));
        const float fresnel = r0 + (1.0f - r0) * (hdn * hdn * hdn * hdn * hdn);
        const float third = 1.0f / 3.0f;
        float rhoS = (third * m.specular[0] + third * m.specular[1] + third * m.specular[2]) * fresnel;
        float rhoD = (third * m.diffuse[0] + third * m.diffuse[1] + third * m.diffuse[2]) * (1.0f - fresnel) *
                     (1.0f - m.dissolve);
        float rhoR = (third * m.transmittance[0] + third * m.transmittance[1] + third * m.transmittance[2]) *
                     (1.0f - fresnel) * m.dissolve;
        float rhoE = third * m.emission[0] + third * m.emission[1] + third * m.emission[2];
        const float total = rhoS + rhoD + rhoR + rhoE;
        if (!(total < 0.0001f)) {
          rhoS /= total;
          rhoD /= total;
          rhoR /= total;
          const uint32_t dim = 8u + 8u * bounce;
          const float pick = rand_ps(pix, smp, dim + 5, p.seed);
          const float Px = o.x + d.x * t, Py = o.y + d.y * t, Pz = o.z + d.z * t;
          float ox = 0.f, oy = 0.f, oz = 0.f;  // outDir
          bool scatter = true;
          if (pick < rhoS) {  // glossy reflection
            const float k = 2.0f * (d.x * nx + d.y * ny + d.z * nz);
            ox = d.x - k * nx;
            oy = d.y - k * ny;
            oz = d.z - k * nz;
            w.x *= m.specular[0];
            w.y *= m.specular[1];
            w.z *= m.specular[2];
            w.w = 1.0f;
          } else if (pick < rhoS + rhoD) {  // diffuse + next-event estimation
            if (p.n_emissive > 0) {  // MeshLight::sampleDirect (main.cc:337-392)
              float xi1 = rand_ps(pix, smp, dim + 0, p.seed);
              const float xi2 = rand_ps(pix, smp, dim + 1, p.seed);
              const float nf = (float)p.n_emissive;
              const uint32_t face = min((uint32_t)floorf(xi1 * nf), p.n_emissive - 1u);
              xi1 = xi1 * nf - (float)face;
              const uint32_t fid = emissive[face];
              const PathMaterial lm = mats[mat_ids ? mat_ids[fid] : 0u];
              const uint32_t f0 = faces[3 * (size_t)fid], f1 = faces[3 * (size_t)fid + 1], f2 = faces[3 * (size_t)fid + 2];
              const float *v0 = verts + 3 * (size_t)f0, *v1 = verts + 3 * (size_t)f1, *v2 = verts + 3 * (size_t)f2;
              const float s1 = sqrtf(xi1), c0 = 1.0f - s1, c1 = s1 * (1.0f - xi2), c2 = s1 * xi2;
              float lnx, lny, lnz, la2;
              geometric_normal(verts, faces, fid, lnx, lny, lnz, la2);
              const float area = 0.5f * la2;
              float lx = c0 * v0[0] + c1 * v1[0] + c2 * v2[0] - Px, ly = c0 * v0[1] + c1 * v1[1] + c2 * v2[1] - Py,
                    lz = c0 * v0[2] + c1 * v1[2] + c2 * v2[2] - Pz;
              const float dist = sqrtf(lx * lx + ly * ly + lz * lz);
              if (dist > 0.000001f) {
                const float id = 1.0f / dist;
                lx *= id;
                ly *= id;
                lz *= id;
                // The reference traces the shadow ray whenever the solid-angle pdf is positive -- also when the light
                // faces away (cosAtLight = 0 makes PdfAtoW infinite and the contribution exactly 0, main.cc:381-390,
                // 943-950): those Traverse calls are part of its loop, so they are part of ours.
                const float cos_l = fmaxf(-(lx * lnx + ly * lny + lz * lnz), 0.0f);
                const float pdf = (1.0f / nf) * (1.0f / area) * (dist * dist) / fabsf(cos_l);  // PdfAtoW
                if (pdf > 0.0f) {
                  const float cos_t = fabsf(lx * nx + ly * ny + lz * nz);
                  const float k = (1.0f / 3.14159265358979f) * cos_l * cos_t / pdf;  // brdf * cosine EDF * cos / pdf
                  so = make_float4(Px, Py, Pz, 0.00001f);
                  sd = make_float4(lx, ly, lz, dist - 0.00001f);
                  sc = make_float4(k * m.diffuse[0] * lm.emission[0] * w.x, k * m.diffuse[1] * lm.emission[1] * w.y,
                                   k * m.diffuse[2] * lm.emission[2] * w.z, __uint_as_float(pix));
                  shadow = true;
                }
              }
            }
            // cosine-weighted direction about the flipped normal (main.cc:216-250)
            const float sg = nz >= 0.0f ? 1.0f : -1.0f;
            const float a = -1.0f / (sg + nz), b = nx * ny * a;
            const float t1x = 1.0f + sg * nx * nx * a, t1y = sg * b, t1z = -sg * nx;
            const float t2x = b, t2y = sg + ny * ny * a, t2z = -ny;
            const float u1 = rand_ps(pix, smp, dim + 2, p.seed), u2 = rand_ps(pix, smp, dim + 3, p.seed);
            const float r = sqrtf(u1);
            float sn, cs;
            sincosf(6.28318530718f * u2, &sn, &cs);
            const float hx = r * cs, hy = r * sn, hz = sqrtf(fmaxf(0.0f, 1.0f - u1));
            ox = t1x * hx + t2x * hy + nx * hz;
            oy = t1y * hx + t2y * hy + ny * hz;
            oz = t1z * hx + t2z * hy + nz * hz;
            w.x *= m.diffuse[0];
            w.y *= m.diffuse[1];
            w.z *= m.diffuse[2];
            w.w = 0.0f;
          } else if (pick < rhoD + rhoS + rhoR) {  // refraction: refract(rayDir, -inside * originalNorm, n1)
            const float rnx = -inside * onx, rny = -inside * ony, rnz = -inside * onz;
            const float ndi = rnx * d.x + rny * d.y + rnz * d.z;
            const float k = 1.0f - n1 * n1 * (1.0f - ndi * ndi);
            if (k < 0.0f) {
              ox = oy = oz = 0.0f;  // the reference continues with a zero direction (a ray that hits nothing)
            } else {
              const float c = n1 * ndi + sqrtf(k);
              ox = n1 * d.x - c * rnx;
              oy = n1 * d.y - c * rny;
              oz = n1 * d.z - c * rnz;
            }
            w.x *= m.transmittance[0];
            w.y *= m.transmittance[1];
            w.z *= m.transmittance[2];
            w.w = 1.0f;
          } else {  // emission (cosine EDF), only if the previous event did not sample the lights
            if (w.w != 0.0f) {
              const float c = fmaxf(-(onx * d.x + ony * d.y + onz * d.z), 0.0f);
              atomicAdd(accum + 3 * (size_t)pix + 0, c * m.emission[0] * w.x);
              atomicAdd(accum + 3 * (size_t)pix + 1, c * m.emission[1] * w.y);
              atomicAdd(accum + 3 * (size_t)pix + 2, c * m.emission[2] * w.z);
            }
            scatter = false;
          }
          // ---- continuation + Russian roulette of the NEXT bounce (main.cc:828-837)
          if (scatter && bounce + 1 < p.max_bounces) {
            bool alive = true;
            if (bounce + 1 > 3) {
              alive = rand_ps(pix, smp, dim + 4, p.seed) >= 0.2f;
              const float inv = 1.0f / 0.8f;
              w.x *= inv;
              w.y *= inv;
              w.z *= inv;
            }
            if (alive) {
              co = make_float4(Px, Py, Pz, p.ray_min_t);
              cd = make_float4(ox, oy, oz, p.ray_max_t);
              q.weight[pid] = w;
              cont = true;
            }
          }
        }
      }
    }
    const int lane = threadIdx.x & 31;
    const unsigned mc = __ballot_sync(0xFFFFFFFFu, cont), ms = __ballot_sync(0xFFFFFFFFu, shadow);
    if ((mc | ms) == 0u) return;
    unsigned long long bc = 0, bs = 0;
    if (lane == 0) {
      if (mc) bc = atomicAdd(counters + 0, (unsigned long long)__popc(mc));
      if (ms) bs = atomicAdd(counters + 1, (unsigned long long)__popc(ms));
    }
    bc = __shfl_sync(0xFFFFFFFFu, bc, 0);
    bs = __shfl_sync(0xFFFFFFFFu, bs, 0);
    const unsigned lt = (1u << lane) - 1u;
    if (cont) {
      const unsigned long long j = bc + __popc(mc & lt);
      q.org_tmin[in ^ 1][j] = co;
      q.dir_tmax[in ^ 1][j] = cd;
      q.path_id[in ^ 1][j] = pid;
    }
    if (shadow) {
      const unsigned long long j = bs + __popc(ms & lt);
      q.sh_org_tmin[j] = so;
      q.sh_dir_tmax[j] = sd;
      q.sh_contrib_pix[j] = sc;
    }
  }
};

// Shadow rays: an unoccluded light sample adds its contribution (CheckForOccluder returned false)
struct ShadowAccumulateEpilogue {
  static constexpr bool kAnyHit = false;
  const float4 *contrib_pix;
  float *accum;
  __device__ __forceinline__ void operator()(bool retiring, size_t ray_idx, float t, float u, float v, uint32_t prim,
                                             float max_t, const uint32_t *) const {
    (void)u;
    (void)v;
    (void)prim;
    if (retiring && !(t < max_t)) {
      const float4 c = contrib_pix[ray_idx];
      const size_t pix = __float_as_uint(c.w);
      atomicAdd(accum + 3 * pix + 0, c.x);
      atomicAdd(accum + 3 * pix + 1, c.y);
      atomicAdd(accum + 3 * pix + 2, c.z);
    }
  }
};

// NRT_TRAVERSE_ANY_HIT: the same retire steps on rays the kernel stops at their first hit
template <class E>
struct AnyHit : E {
  static constexpr bool kAnyHit = true;
  __host__ __device__ explicit AnyHit(const E &e) : E(e) {}
};

}  // namespace nrt
