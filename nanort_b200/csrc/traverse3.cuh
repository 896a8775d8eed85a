// Round-2 production traversal kernel (sm_100a).  Same algorithm and the same arithmetic as traverse_fast2_kernel
// (persistent warps, ray pool refilled by warp ballot, while-while with one postponed leaf per lane, per-lane stack
// of (ref, entry distance)), re-shaped by what the round-1 ncu captures showed (profiles/r02_*):
//
//   * 12 of the 54 instructions of a child-pair slab test were FSELs picking bmin/bmax by ray_dir_sign
//     (nanort.h:2291-2302) and they sit on the half-rate ALU pipe (60 % busy)  -> PairNode: the selection is an
//     address computed once per ray, the planes arrive as {near0 near1 far0 far1}.
//   * the (kx, ky, kz) permutation of the watertight test (nanort.h:1073-1081) compiled to eight small branches per
//     triangle (48 of 169 instructions)                                      -> TriCM: component-major triangles,
//     the permutation is the load address, the ray carries its origin already permuted.
//   * the far-plane widening `* 1.00000024f` (nanort.h:2303-2305) was applied per axis; rounding is monotonic, so
//     min(a*w, b*w, c*w) == min(a, b, c)*w bit for bit (w > 0): one multiply per box instead of three.
//   * three inlined copies of the pop loop ran at 3-16 active lanes           -> one pop site per iteration.
//
// Replaces (file:line under /root/reference): BVHAccel<float>::Traverse nanort.h:2487-2556, TestLeafNode :2372-2407,
// IntersectRayAABB<float> :2284-2325, TriangleIntersector::Intersect :1054-1150, PrepareTraversal :1163-1201.
#pragma once
#include "common.cuh"
#include "trav_common.cuh"

namespace nrt {

struct RayCtx3 {
  float ox, oy, oz;     // origin
  float ix, iy, iz;     // vsafe_inverse(dir)
  float Sx, Sy, Sz;     // watertight shear constants
  float okx, oky, okz;  // origin permuted by (kx, ky, kz)
  float t_min;
  // which float4 (16-byte unit) of a PairNode holds {near0 near1 far0 far1} of each axis: 0/1, 2/3, 4/5
  uint32_t nx, ny, nz;
  uint32_t tx, ty, tz;  // which float4 of a TriCM holds component kx / ky / kz: 0..2
};

__device__ __forceinline__ void setup_ray3(RayCtx3 &c, float ox, float oy, float oz, float dx, float dy, float dz,
                                           float min_t, bool cpp03) {
  RayCtx r;
  setup_ray(r, ox, oy, oz, dx, dy, dz, min_t, cpp03);  // the one definition of the reference's per-ray constants
  c.ox = ox;
  c.oy = oy;
  c.oz = oz;
  c.ix = r.ix;
  c.iy = r.iy;
  c.iz = r.iz;
  c.Sx = r.Sx;
  c.Sy = r.Sy;
  c.Sz = r.Sz;
  c.okx = sel3(r.kx, ox, oy, oz);
  c.oky = sel3(r.ky, ox, oy, oz);
  c.okz = sel3(r.kz, ox, oy, oz);
  c.t_min = min_t;
  c.nx = r.sx ? 1u : 0u;
  c.ny = r.sy ? 3u : 2u;
  c.nz = r.sz ? 5u : 4u;
  c.tx = (uint32_t)r.kx;
  c.ty = (uint32_t)r.ky;
  c.tz = (uint32_t)r.kz;
}

// Both child boxes of a PairNode (nanort.h:2284-2325, twice).  X = {near0 near1 far0 far1} of the x planes, etc.
// NaN handling as in slab(): fmaxf / fminf drop a NaN plane value, the range values are never NaN.
__device__ __forceinline__ void slab_pair(const RayCtx3 &c, const float4 X, const float4 Y, const float4 Z, float min_t,
                                          float best_t, bool &h0, bool &h1, float &t0, float &t1) {
  const float n0x = (X.x - c.ox) * c.ix, n1x = (X.y - c.ox) * c.ix;
  const float f0x = (X.z - c.ox) * c.ix, f1x = (X.w - c.ox) * c.ix;
  const float n0y = (Y.x - c.oy) * c.iy, n1y = (Y.y - c.oy) * c.iy;
  const float f0y = (Y.z - c.oy) * c.iy, f1y = (Y.w - c.oy) * c.iy;
  const float n0z = (Z.x - c.oz) * c.iz, n1z = (Z.y - c.oz) * c.iz;
  const float f0z = (Z.z - c.oz) * c.iz, f1z = (Z.w - c.oz) * c.iz;
  t0 = fmaxf(fmaxf(fmaxf(n0x, n0y), n0z), min_t);
  t1 = fmaxf(fmaxf(fmaxf(n1x, n1y), n1z), min_t);
  // min over the axes first, widen once: identical to widening each axis (monotonic rounding, factor > 0)
  const float e0 = fminf(fminf(fminf(f0x, f0y), f0z) * 1.00000024f, best_t);
  const float e1 = fminf(fminf(fminf(f1x, f1y), f1z) * 1.00000024f, best_t);
  h0 = t0 <= e0;
  h1 = t1 <= e1;
}

// Same for the 64-byte WideNode (q0 = c0.lo.xyz, c0.hi.x | q1 = c0.hi.yz, c1.lo.xy | q2 = c1.lo.z, c1.hi.xyz): the
// planes are picked by the ray's direction signs with selects.
__device__ __forceinline__ void slab_pair_sel(const RayCtx3 &c, const float4 q0, const float4 q1, const float4 q2,
                                              float min_t, float best_t, bool &h0, bool &h1, float &t0, float &t1) {
  const bool sx = (c.nx & 1u) != 0u, sy = (c.ny & 1u) != 0u, sz = (c.nz & 1u) != 0u;
  const float4 X = make_float4(sx ? q0.w : q0.x, sx ? q2.y : q1.z, sx ? q0.x : q0.w, sx ? q1.z : q2.y);
  const float4 Y = make_float4(sy ? q1.x : q0.y, sy ? q2.z : q1.w, sy ? q0.y : q1.x, sy ? q1.w : q2.z);
  const float4 Z = make_float4(sz ? q1.y : q0.z, sz ? q2.w : q2.x, sz ? q0.z : q1.y, sz ? q2.x : q2.w);
  slab_pair(c, X, Y, Z, min_t, best_t, h0, h1, t0, t1);
}

// Watertight test on a component-major triangle: VX = {A[kx] B[kx] C[kx] w} etc.  Arithmetic order of
// nanort.h:1073-1147 (tri_test2 with the permutation already applied by the loads).
__device__ __forceinline__ void tri_test3(const RayCtx3 &c, const TraceOptions16 &opt, bool filtered, const float4 VX,
                                          const float4 VY, const float4 VZ, Best &best) {
  const uint32_t prim = __float_as_uint(VX.w) & 0x7FFFFFFFu;
  bool rej = false;
  if (filtered)  // warp-uniform: the options are kernel parameters
    rej = (prim < opt.prim_ids_range[0]) | (prim >= opt.prim_ids_range[1]) | (prim == opt.skip_prim_id);
  const float Akx = VX.x - c.okx, Bkx = VX.y - c.okx, Ckx = VX.z - c.okx;
  const float Aky = VY.x - c.oky, Bky = VY.y - c.oky, Cky = VY.z - c.oky;
  const float Akz = VZ.x - c.okz, Bkz = VZ.y - c.okz, Ckz = VZ.z - c.okz;
  const float Ax = Akx - c.Sx * Akz, Ay = Aky - c.Sy * Akz;
  const float Bx = Bkx - c.Sx * Bkz, By = Bky - c.Sy * Bkz;
  const float Cx = Ckx - c.Sx * Ckz, Cy = Cky - c.Sy * Ckz;
  float U = Cx * By - Cy * Bx;
  float V = Ax * Cy - Ay * Cx;
  float W = Bx * Ay - By * Ax;
  if (U == 0.0f || V == 0.0f || W == 0.0f) {  // rare: exact edge / vertex hits (nanort.h:1095-1107)
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

// 256-bit global load (sm_100: LDG.E.256), read-only path.  One instruction -- one L1 wavefront per lane -- for half of
// a 64-byte WideNode; p must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const void *p, float4 &a, float4 &b) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
               : "l"(p));
}

constexpr int kNone3 = kEmptyLeaf;  // "no node": finished, or (with a non-empty stack) waiting for a pop

template <int BLOCK_, int MINB_, int REFILL_MIN_, int NODE_EXIT_, bool PAIR128_ = true, bool LOAD256_ = false,
          int LEAF_AGAIN_MIN_ = 1, int LEAF_SLOTS_ = 1, bool DEFER_RETIRE_ = false, int NODE_UNROLL_ = 1>
struct Policy3 {
  static constexpr int kBlock = BLOCK_;
  static constexpr int kMinBlocks = MINB_;
  static constexpr int kRefillMin = REFILL_MIN_;  // lanes that must have retired before the warp fetches new rays
  static constexpr int kNodeExit = NODE_EXIT_;    // leave the node phase when fewer lanes than this want node work
  // true: 128-byte PairNode (sign-addressed planes, no selects); false: the 64-byte WideNode with 12 selects per pair
  // (half the cache footprint per node) -- A/B knob, see profiles/r02_variant_sweep.md
  static constexpr bool kPair128 = PAIR128_;
  // 64-byte node only: fetch it with two 256-bit loads instead of four (the L1 wavefront count of incoherent rays)
  static constexpr bool kLoad256 = LOAD256_;
  // after the first leaf round of an outer iteration, another one runs only while at least this many lanes hold a
  // leaf (1: until none is left; 33: one round) -- a lane's second leaf (found while the first was postponed) is
  // otherwise tested in a round of its own with the few lanes that have one; carried over, it joins the next phase's
  static constexpr int kLeafAgainMin = LEAF_AGAIN_MIN_;
  // postponed leaves a lane may hold before it parks (1 or 2)
  static constexpr int kLeafSlots = LEAF_SLOTS_;
  // true: finished rays wait for the retire step until retired + empty lanes reach kRefillMin (or nothing else is
  // left to do), so that the epilogue and the refill that follows run with more lanes
  static constexpr bool kDeferRetire = DEFER_RETIRE_;
  // node steps per evaluation of the node phase's exit conditions (two ballots + a population count per check)
  static constexpr int kNodeUnroll = NODE_UNROLL_;
};

// DEPTH: capacity of the per-lane stack (entries); chosen by the launcher from the tree depth, so a push can
// never overflow (a child pair pushes one entry and descends one level).
template <class Rays, int DEPTH, bool COUNT, class P, class Epi>
__global__ void __launch_bounds__(P::kBlock, P::kMinBlocks)
    traverse_fast3_kernel(const void *__restrict__ pair, const TriCM *__restrict__ tris, Rays rays, size_t n, Epi epi,
                          TraceOptions16 opt, uint32_t flags, unsigned long long *cursor, unsigned long long *counts,
                          const unsigned long long *n_ptr) {
  if (n_ptr) n = (size_t)*n_ptr;
  const int lane = threadIdx.x & 31;
  const unsigned lt_mask = (1u << lane) - 1u;
  const bool cpp03 = (flags & NRT_TRAVERSE_CPP03_INVERSE) != 0;
  const bool filtered = opt.prim_ids_range[0] != 0u || opt.prim_ids_range[1] < 0x7FFFFFFFu ||
                        opt.skip_prim_id < 0x7FFFFFFFu;
  const float4 *const pair4 = reinterpret_cast<const float4 *>(pair);  // 8 float4 per PairNode
  const float4 *const tris4 = reinterpret_cast<const float4 *>(tris);  // 3 float4 per TriCM

  // per-lane stack; the two extra entries keep what only the retire step reads (ray index, max_t) out of the
  // register file (thread-local memory, L1-resident)
  constexpr int PW = Rays::kPayloadWords;  // words a ray loader hands to the retire step (parked next to the stack)
  uint2 lstk[DEPTH + 2 + (PW + 1) / 2];
  int sp = 0;
  RayCtx3 c;
  Best best;
  float min_t = 0.0f;
  bool alive = false;
  int cur = kNone3, leaf = kNone3;
  int leaf2 = kNone3;  // second postponed leaf (P::kLeafSlots == 2 only; leaf2 != kNone3 implies leaf != kNone3)
  bool exhausted = false;
  unsigned long long n_boxes = 0, n_prims = 0;
  // COUNT only: lane-state histogram of the warp's iterations (nrt_traverse_lane_stats_device; lane 0 accumulates)
  unsigned long long st[14] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};

  for (;;) {
    // ---- replace retired rays (warp-ballot compaction of the ray pool)
    const unsigned dead = __ballot_sync(FULL_MASK, !alive);
    if (dead != 0u && !exhausted && (dead == FULL_MASK || __popc(dead) >= P::kRefillMin)) {
      const int cnt = __popc(dead);
      const int leader = __ffs(dead) - 1;
      if (COUNT) {
        st[0] += 1;    // refill events
        st[1] += cnt;  // lanes refilled (incl. lanes past the end of the ray set)
      }
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(cursor, (unsigned long long)cnt);
      base = __shfl_sync(FULL_MASK, base, leader);
      if (base + (unsigned long long)cnt >= (unsigned long long)n) exhausted = true;
      if (!alive) {
        const unsigned long long mine = base + (unsigned long long)__popc(dead & lt_mask);
        if (mine < (unsigned long long)n) {
          float ox, oy, oz, dx, dy, dz, max_t;
          uint32_t payload[PW > 0 ? PW : 1];
          rays.load((size_t)mine, ox, oy, oz, dx, dy, dz, min_t, max_t, payload);
#pragma unroll
          for (int w = 0; w + 1 < PW; w += 2) lstk[DEPTH + 2 + w / 2] = make_uint2(payload[w], payload[w + 1]);
          if (PW & 1) lstk[DEPTH + 2 + PW / 2].x = payload[PW - 1];
          setup_ray3(c, ox, oy, oz, dx, dy, dz, min_t, cpp03);
          best.t = max_t;
          best.u = 0.0f;
          best.v = 0.0f;
          best.prim = 0xFFFFFFFFu;
          alive = true;
          lstk[DEPTH] = make_uint2((uint32_t)mine, (uint32_t)(mine >> 32));
          lstk[DEPTH + 1].x = __float_as_uint(max_t);
          sp = 0;
          cur = range_has_nan(min_t, max_t) ? kNone3 : 0;
          leaf = kNone3;
          leaf2 = kNone3;
          if (COUNT) n_boxes += 1;
        }
      }
    }
    if (__all_sync(FULL_MASK, !alive)) {
      if (exhausted) break;
      continue;
    }

    // ---- inner nodes.  A lane wants node work when it stands on a node, or stands nowhere but has stack entries.
    for (;;) {
      bool want = cur >= 0 || (cur == kNone3 && sp > 0);
      const unsigned desc = __ballot_sync(FULL_MASK, want);
      if (desc == 0u) break;
      if (P::kNodeExit > 1 && __popc(desc) < P::kNodeExit && __any_sync(FULL_MASK, leaf != kNone3)) break;
#pragma unroll
      for (int rep = 0; rep < P::kNodeUnroll; ++rep) {
        if (rep > 0) want = cur >= 0 || (cur == kNone3 && sp > 0);
        if (want) {
          // The one pop site.  Pops until the lane stands on a node again: entries that start behind the current best
          // are dropped without touching memory (same visit set as the reference's re-test at pop, nanort.h:2532), a
          // popped leaf goes into the free leaf slot and the popping continues -- so that every lane that still has
          // inner nodes to visit takes a node step in THIS iteration (a lane that only pops is a wasted warp step).
          while (cur == kNone3 && sp > 0) {
            --sp;
            const uint2 e = lstk[sp];
            if (__uint_as_float(e.y) <= best.t) {
              cur = (int)e.x;
              if (cur < 0 && leaf == kNone3) {
                leaf = cur;
                cur = kNone3;
              } else if (P::kLeafSlots == 2 && cur < 0 && leaf2 == kNone3) {
                leaf2 = cur;
                cur = kNone3;
              }
            }
          }
        }
        if (COUNT) {  // who does what in this warp step
          const bool fin = alive && cur == kNone3 && leaf == kNone3 && sp == 0;
          st[2] += 1;                                                  // node-phase warp steps
          st[3] += __popc(__ballot_sync(FULL_MASK, cur >= 0));          // lanes testing a child pair
          st[4] += __popc(__ballot_sync(FULL_MASK, !alive));            // lanes without a ray
          st[5] += __popc(__ballot_sync(FULL_MASK, fin));               // lanes whose ray is finished (waits for the retire step)
          st[6] += __popc(__ballot_sync(FULL_MASK, alive && !fin && cur < 0));  // lanes parked on leaves
        }
        if (want) {
          if (cur >= 0) {
            bool h0, h1;
            float t0, t1;
            int2 R;
            if (P::kPair128) {
              const uint32_t n8 = (uint32_t)cur * 8u;
              const float4 X = __ldg(pair4 + (n8 + c.nx));
              const float4 Y = __ldg(pair4 + (n8 + c.ny));
              const float4 Z = __ldg(pair4 + (n8 + c.nz));
              R = __ldg(reinterpret_cast<const int2 *>(pair4 + (n8 + 6u)));
              slab_pair(c, X, Y, Z, min_t, best.t, h0, h1, t0, t1);
            } else if (P::kLoad256) {
              const float4 *q = pair4 + (uint32_t)cur * 4u;
              float4 q0, q1, q2, q3;
              ldg256(q, q0, q1);
              ldg256(q + 2, q2, q3);
              R = make_int2(__float_as_int(q3.x), __float_as_int(q3.y));
              slab_pair_sel(c, q0, q1, q2, min_t, best.t, h0, h1, t0, t1);
            } else {
              const float4 *q = pair4 + (uint32_t)cur * 4u;
              const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
              R = __ldg(reinterpret_cast<const int2 *>(q + 3));
              slab_pair_sel(c, q0, q1, q2, min_t, best.t, h0, h1, t0, t1);
            }
            if (COUNT) n_boxes += 2;
            const bool swap = t1 < t0;
            const bool both = h0 & h1;
            if (both) {
              lstk[sp] = make_uint2((uint32_t)(swap ? R.x : R.y), __float_as_uint(swap ? t0 : t1));
              sp++;
            }
            cur = both ? (swap ? R.y : R.x) : (h0 ? R.x : (h1 ? R.y : kNone3));
            if (cur < 0 && cur != kNone3 && leaf == kNone3) {  // postpone the first leaf, keep descending
              leaf = cur;
              cur = kNone3;
            } else if (P::kLeafSlots == 2 && cur < 0 && cur != kNone3 && leaf2 == kNone3) {
              leaf2 = cur;
              cur = kNone3;
            }
          }
        }
      }
    }

    // ---- leaves
    for (bool first_round = true;; first_round = false) {
      const unsigned with_leaf = __ballot_sync(FULL_MASK, leaf != kNone3);
      if (with_leaf == 0u) break;
      if (P::kLeafAgainMin > 1 && !first_round && __popc(with_leaf) < P::kLeafAgainMin) break;  // carried over
      if (COUNT) {
        st[7] += 1;                                                // leaf-phase rounds
        st[8] += __popc(with_leaf);  // lanes that enter a round with a leaf
      }
      if (leaf != kNone3) {
        uint32_t slot = (uint32_t)(~leaf);
        for (;;) {
          if (COUNT && lane == __ffs(__activemask()) - 1) st[9] += 1;  // triangle-test warp steps
          const uint32_t s3 = slot * 3u;
          const float4 VX = __ldg(tris4 + (s3 + c.tx));
          const float4 VY = __ldg(tris4 + (s3 + c.ty));
          const float4 VZ = __ldg(tris4 + (s3 + c.tz));
          if (COUNT) n_prims++;
          tri_test3(c, opt, filtered, VX, VY, VZ, best);
          if ((int)__float_as_uint(VX.w) < 0) break;  // last triangle of the leaf
          slot++;
        }
        leaf = kNone3;
        if (P::kLeafSlots == 2) {
          leaf = leaf2;
          leaf2 = kNone3;
        }
        if (cur < 0 && cur != kNone3) {  // a leaf was waiting in cur (the lane was parked)
          if (leaf == kNone3) {
            leaf = cur;
            cur = kNone3;
          } else if (P::kLeafSlots == 2) {  // leaf2 was just vacated
            leaf2 = cur;
            cur = kNone3;
          }
        }
        // occlusion query (NRT_TRAVERSE_ANY_HIT): a hit inside [min_t, max_t) ends the ray.  A record accepted AT
        // max_t is a miss (nanort.h:2552) and does not.
        if (Epi::kAnyHit && best.prim != 0xFFFFFFFFu && best.t < __uint_as_float(lstk[DEPTH + 1].x)) {
          sp = 0;
          cur = kNone3;
          leaf = kNone3;
          leaf2 = kNone3;
        }
      }
    }

    // ---- retire: the epilogue (store the hit / spawn the AO ray / accumulate / shade) runs warp-wide
    const bool retiring = alive && cur == kNone3 && leaf == kNone3 && sp == 0;
    bool retire_now;
    if (P::kDeferRetire) {
      // wait until the retire step and the refill behind it have enough lanes -- unless the ray set is exhausted
      // (no refill will come) or no lane has anything else to do
      const unsigned rm = __ballot_sync(FULL_MASK, retiring);
      const unsigned idle = rm | __ballot_sync(FULL_MASK, !alive);
      retire_now = rm != 0u && (exhausted || idle == FULL_MASK || __popc(idle) >= P::kRefillMin);
    } else {
      retire_now = __any_sync(FULL_MASK, retiring);
    }
    if (COUNT) {
      st[12] += 1;  // outer iterations
      if (retire_now) {
        st[10] += 1;                                              // retire events
        st[11] += __popc(__ballot_sync(FULL_MASK, retiring));  // lanes retiring
      }
    }
    if (retire_now) {
      size_t ray_idx = 0;
      float max_t = 0.0f;
      uint32_t payload[PW > 0 ? PW : 1];
      if (retiring) {
        const uint2 id = lstk[DEPTH];
        ray_idx = (size_t)id.x | ((size_t)id.y << 32);
        max_t = __uint_as_float(lstk[DEPTH + 1].x);
#pragma unroll
        for (int w = 0; w + 1 < PW; w += 2) {
          const uint2 e = lstk[DEPTH + 2 + w / 2];
          payload[w] = e.x;
          payload[w + 1] = e.y;
        }
        if (PW & 1) payload[PW - 1] = lstk[DEPTH + 2 + PW / 2].x;
      }
      epi(retiring, ray_idx, best.t, best.u, best.v, best.prim, max_t, payload);
    }
    if (retiring && retire_now) alive = false;
  }

  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      n_boxes += __shfl_down_sync(FULL_MASK, n_boxes, o);
      n_prims += __shfl_down_sync(FULL_MASK, n_prims, o);
    }
    // warp-uniform statistics were added by every lane alike: lane 0's copy is the warp's; st[9] is per lane
    for (int o = 16; o > 0; o >>= 1) st[9] += __shfl_down_sync(FULL_MASK, st[9], o);
    if (lane == 0) {
      atomicAdd(counts + 0, n_boxes);
      atomicAdd(counts + 1, n_prims);
#pragma unroll
      for (int k = 0; k < 14; ++k) atomicAdd(counts + 2 + k, st[k]);
    }
  }
}

}  // namespace nrt
