// Ray traversal kernels (sm_100a).  Compiled with --fmad=false: every float
// operation below is individually rounded, exactly like the reference's x86-64
// SSE2 scalar code, so sign decisions on U/V/W and
// the reported t/u/v are bit-identical to CPU nanort for the same triangle.
//
// Replaces (file:line under /root/reference):
//   BVHAccel<float>::Traverse            nanort.h:2487-2556
//   BVHAccel<float>::TestLeafNode        nanort.h:2372-2407
//   IntersectRayAABB<float>              nanort.h:2284-2325
//   TriangleIntersector::Intersect       nanort.h:1054-1150
//   TriangleIntersector::PrepareTraversal nanort.h:1163-1201
//   vsafe_inverse                        nanort.h:414-465
#include <float.h>
#include <math_constants.h>

#include <stdlib.h>

#include "common.cuh"
#include "trav_common.cuh"
#include "wavefront.cuh"
#include "traverse3.cuh"

namespace nrt {

// ------------------------------------------------------------------ ray loaders
struct AosRays {
  static constexpr int kPayloadWords = 0;
  const Ray36 *rays;
  __device__ __forceinline__ void load(size_t i, float &ox, float &oy, float &oz, float &dx, float &dy,
                                       float &dz, float &tmin, float &tmax, uint32_t * = nullptr) const {
    const float *p = reinterpret_cast<const float *>(rays + i);
    ox = __ldcs(p + 0);
    oy = __ldcs(p + 1);
    oz = __ldcs(p + 2);
    dx = __ldcs(p + 3);
    dy = __ldcs(p + 4);
    dz = __ldcs(p + 5);
    tmin = __ldcs(p + 6);
    tmax = __ldcs(p + 7);
  }
};

// 32-byte ray records (NRT_TRAVERSE_RAY32): {org.xyz, dir.x} {dir.yz, min_t, max_t}, two 128-bit loads
struct Aos32Rays {
  static constexpr int kPayloadWords = 0;
  const float4 *rays;
  __device__ __forceinline__ void load(size_t i, float &ox, float &oy, float &oz, float &dx, float &dy,
                                       float &dz, float &tmin, float &tmax, uint32_t * = nullptr) const {
    const float4 a = __ldcs(rays + 2 * i), b = __ldcs(rays + 2 * i + 1);
    ox = a.x;
    oy = a.y;
    oz = a.z;
    dx = a.w;
    dy = b.x;
    dz = b.y;
    tmin = b.z;
    tmax = b.w;
  }
};

struct SoaRays {
  static constexpr int kPayloadWords = 0;
  const float4 *org_tmin;
  const float4 *dir_tmax;
  __device__ __forceinline__ void load(size_t i, float &ox, float &oy, float &oz, float &dx, float &dy,
                                       float &dz, float &tmin, float &tmax, uint32_t * = nullptr) const {
    float4 o = __ldcs(org_tmin + i);  // read once: evict-first, keep L1/L2 for the tree
    float4 d = __ldcs(dir_tmax + i);
    ox = o.x;
    oy = o.y;
    oz = o.z;
    tmin = o.w;
    dx = d.x;
    dy = d.y;
    dz = d.z;
    tmax = d.w;
  }
};

// ------------------------------------------------------------------ conformance walk
// One thread per ray over the nanort 40-byte node array, in the reference's
// exact order: pop, slab-test against the current best, near child by
// dir_sign[node.axis], leaf prims in indices_ order (nanort.h:2526-2547).
constexpr int kConfStack = 512;  // kNANORT_MAX_STACK_DEPTH

template <class Rays, bool COUNT>
__global__ void __launch_bounds__(128)
    traverse_conformance_kernel(const Node40 *__restrict__ nodes, const PackedTri *__restrict__ tris,
                                Rays rays, size_t n, Hit16 *__restrict__ hits, uint8_t *__restrict__ mask,
                                TraceOptions16 opt, uint32_t flags, unsigned long long *counts,
                                const unsigned long long *n_ptr) {
  if (n_ptr) n = (size_t)*n_ptr;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long n_boxes = 0, n_prims = 0;
  if (i < n) {
    float ox, oy, oz, dx, dy, dz, min_t, max_t;
    rays.load(i, ox, oy, oz, dx, dy, dz, min_t, max_t);
    RayCtx c;
    setup_ray(c, ox, oy, oz, dx, dy, dz, min_t, (flags & NRT_TRAVERSE_CPP03_INVERSE) != 0);
    Best best;
    best.t = max_t;
    best.u = 0.0f;
    best.v = 0.0f;
    best.prim = 0xFFFFFFFFu;
    float hit_t = max_t;
    uint32_t stack[kConfStack];
    int sp = range_has_nan(min_t, max_t) ? -1 : 0;
    stack[0] = 0;
    while (sp >= 0) {
      const Node40 *nd = nodes + stack[sp];
      sp--;
      const float *f = reinterpret_cast<const float *>(nd);
      float lox = __ldg(f + 0), loy = __ldg(f + 1), loz = __ldg(f + 2);
      float hix = __ldg(f + 3), hiy = __ldg(f + 4), hiz = __ldg(f + 5);
      if (COUNT) n_boxes++;
      float tn;
      if (!slab(c, lox, loy, loz, hix, hiy, hiz, min_t, hit_t, tn)) continue;
      int flag = __ldg(&nd->flag);
      uint32_t d0 = __ldg(&nd->data[0]), d1 = __ldg(&nd->data[1]);
      if (flag == 0) {
        int axis = __ldg(&nd->axis);
        int sgn = axis == 0 ? c.sx : (axis == 1 ? c.sy : c.sz);
        uint32_t nearc = sgn ? d1 : d0;
        uint32_t farc = sgn ? d0 : d1;
        if (sp + 2 < kConfStack) {  // the reference only asserts here (nanort.h:2550)
          stack[++sp] = farc;
          stack[++sp] = nearc;
        }
      } else {
        bool any = false;
        for (uint32_t k = 0; k < d0; k++) {
          const float4 *t = reinterpret_cast<const float4 *>(tris + (size_t)d1 + k);
          float4 a = __ldg(t), b = __ldg(t + 1), cc = __ldg(t + 2);
          if (COUNT) n_prims++;
          if (tri_test(c, opt, a, b, cc, best)) any = true;
        }
        if (any) hit_t = best.t;
      }
    }
    if (hits) write_result(hits, mask, i, best, max_t);
  }
  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      n_boxes += __shfl_down_sync(FULL_MASK, n_boxes, o);
      n_prims += __shfl_down_sync(FULL_MASK, n_prims, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd(counts + 0, n_boxes);
      atomicAdd(counts + 1, n_prims);
    }
  }
}

// ------------------------------------------------------------------ fast path
// Persistent warps pull rays from a global cursor and replace finished rays with new ones once enough lanes of the
// warp have retired (warp-ballot compaction of the ray pool).  Traversal is while-while with one postponed leaf per
// lane over the 64-byte child-pair nodes; the per-lane stack keeps (ref, entry distance) so that a popped subtree
// that now lies behind the current best is skipped without touching memory -- the same visit set the reference
// obtains by re-testing the box when it is popped (nanort.h:2532).
//
// Shaped by the round-1 ncu captures (profiles/r01_*): the first version of this kernel was issue bound with ~29 % of
// the issued instructions being control flow and 18 (primary) / 12 (AO) of 32 lanes active.  Hence:
//   * triangle test without early returns (one predicate at the end; only the fp64 fallback branches)
//   * empty children carry an inverted box, so no reference checks in the node step
//   * child selection by selects, one predicated push
//   * policy knobs: lanes that must have retired before a refill, lanes that must still be descending for the node
//     phase to continue, CTA size / minimum CTAs per SM, stack entries kept in shared memory, TMA-staged top treelet
constexpr int kNone = kEmptyLeaf;
constexpr int kFastBlock = 128;
constexpr int kStackSmem = 16;  // upper bound of the per-lane stack entries a policy may keep in shared memory

template <int BLOCK_, int MINB_, int REFILL_MIN_, int NODE_EXIT_, int TREELET_ = 0, int STACK_SMEM_ = 16>
struct FastPolicy {
  static constexpr int kStackEntries = STACK_SMEM_;  // stack entries per lane kept in shared memory (<= 16)
  static constexpr int kBlock = BLOCK_;
  static constexpr int kMinBlocks = MINB_;
  static constexpr int kRefillMin = REFILL_MIN_;
  static constexpr int kNodeExit = NODE_EXIT_;  // leave the node phase when fewer lanes than this descend
  // > 0: the first kTreelet WideNodes (the BFS-ordered top of the tree, layout.cu) are staged into shared
  // memory by one TMA bulk copy (cp.async.bulk + mbarrier) when the CTA starts
  static constexpr int kTreelet = TREELET_;
};

// ---- TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP + SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                             unsigned long long *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}


template <class Rays, int LOCAL_DEPTH, bool COUNT, class P, class Epi>
__global__ void __launch_bounds__(P::kBlock, P::kMinBlocks)
    traverse_fast2_kernel(const WideNode *__restrict__ wide, const PackedTri *__restrict__ tris, Rays rays, size_t n,
                          Epi epi, TraceOptions16 opt, uint32_t flags, unsigned long long *cursor,
                          unsigned long long *counts, const unsigned long long *n_ptr, int n_top_avail) {
  constexpr int BLOCK = P::kBlock;
  constexpr int TREELET = P::kTreelet;
  constexpr int SSM = P::kStackEntries;  // the local part grows by what shared memory gives up
  __shared__ uint2 stk[(SSM > 0 ? SSM : 1) * BLOCK];
  __shared__ __align__(128) float4 top[TREELET > 0 ? TREELET * 4 : 1];
  __shared__ __align__(8) unsigned long long top_bar;
  if (n_ptr) n = (size_t)*n_ptr;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  int n_top = 0;
  if (TREELET > 0) {
    n_top = min(TREELET, n_top_avail);
    if (tid == 0) mbar_init(&top_bar, 1);
    __syncthreads();
    if (tid == 0) {
      mbar_expect_tx(&top_bar, (uint32_t)n_top * 64u);
      tma_bulk_g2s(top, wide, (uint32_t)n_top * 64u, &top_bar);
    }
    mbar_wait(&top_bar, 0);
  }
  const unsigned lt_mask = (1u << lane) - 1u;
  const bool cpp03 = (flags & NRT_TRAVERSE_CPP03_INVERSE) != 0;

  uint2 lstk[LOCAL_DEPTH + (kStackSmem - SSM)];
  int sp = 0;
  RayCtx c;
  Best best;
  float max_t = 0.0f, min_t = 0.0f;
  long long ray_idx = -1;
  int cur = kNone, leaf = kNone;
  bool exhausted = false;
  unsigned long long n_boxes = 0, n_prims = 0;

  auto push = [&](int ref, float t) {
    const uint2 e = make_uint2((uint32_t)ref, __float_as_uint(t));
    if (sp < SSM)
      stk[sp * BLOCK + tid] = e;
    else if (sp - SSM < LOCAL_DEPTH + (kStackSmem - SSM))
      lstk[sp - SSM] = e;
    sp++;
  };
  // next stack entry that does not start behind the current best, or kNone
  auto pop = [&]() -> int {
    while (sp > 0) {
      --sp;
      uint2 e;
      if (sp < SSM)
        e = stk[sp * BLOCK + tid];
      else if (sp - SSM < LOCAL_DEPTH + (kStackSmem - SSM))
        e = lstk[sp - SSM];
      else
        continue;
      if (__uint_as_float(e.y) <= best.t) return (int)e.x;
    }
    return kNone;
  };

  for (;;) {
    // ---- replace retired rays
    const unsigned dead = __ballot_sync(FULL_MASK, ray_idx < 0);
    if (dead != 0u && !exhausted && (dead == FULL_MASK || __popc(dead) >= P::kRefillMin)) {
      const int cnt = __popc(dead);
      const int leader = __ffs(dead) - 1;
      unsigned long long base = 0;
      if (lane == leader) base = atomicAdd(cursor, (unsigned long long)cnt);
      base = __shfl_sync(FULL_MASK, base, leader);
      if (base + (unsigned long long)cnt >= (unsigned long long)n) exhausted = true;
      if (ray_idx < 0) {
        const unsigned long long mine = base + (unsigned long long)__popc(dead & lt_mask);
        if (mine < (unsigned long long)n) {
          float ox, oy, oz, dx, dy, dz;
          rays.load((size_t)mine, ox, oy, oz, dx, dy, dz, min_t, max_t);
          setup_ray(c, ox, oy, oz, dx, dy, dz, min_t, cpp03);
          best.t = max_t;
          best.u = 0.0f;
          best.v = 0.0f;
          best.prim = 0xFFFFFFFFu;
          ray_idx = (long long)mine;
          sp = 0;
          cur = range_has_nan(min_t, max_t) ? kNone : 0;
          leaf = kNone;
          if (COUNT) n_boxes += 1;
        }
      }
    }
    if (__all_sync(FULL_MASK, ray_idx < 0)) {
      if (exhausted) break;
      continue;
    }

    // ---- inner nodes
    for (;;) {
      const unsigned desc = __ballot_sync(FULL_MASK, cur >= 0);
      if (desc == 0u) break;
      if (P::kNodeExit > 1 && __popc(desc) < P::kNodeExit &&
          __any_sync(FULL_MASK, leaf != kNone))  // few lanes still descend while others wait with leaves
        break;
      if (cur >= 0) {
        float4 q0, q1, q2;
        int4 q3;
        if (TREELET > 0 && cur < n_top) {
          const float4 *sp4 = top + cur * 4;
          q0 = sp4[0];
          q1 = sp4[1];
          q2 = sp4[2];
          const float4 r3 = sp4[3];
          q3 = make_int4(__float_as_int(r3.x), __float_as_int(r3.y), __float_as_int(r3.z), __float_as_int(r3.w));
        } else {
          const float4 *p = reinterpret_cast<const float4 *>(wide + cur);
          q0 = __ldg(p);
          q1 = __ldg(p + 1);
          q2 = __ldg(p + 2);
          q3 = __ldg(reinterpret_cast<const int4 *>(p + 3));
        }
        float t0, t1;
        const bool h0 = slab(c, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, min_t, best.t, t0);
        const bool h1 = slab(c, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, min_t, best.t, t1);
        if (COUNT) n_boxes += 2;
        const bool both = h0 & h1;
        const bool swap = t1 < t0;
        const int nearr = swap ? q3.y : q3.x;
        const int farr = swap ? q3.x : q3.y;
        if (both) push(farr, swap ? t0 : t1);
        int next = both ? nearr : (h0 ? q3.x : q3.y);
        if (!(h0 | h1)) next = pop();
        if (next < 0 && next != kNone && leaf == kNone) {  // postpone the first leaf, keep descending
          leaf = next;
          next = pop();
        }
        cur = next;
      }
    }

    // ---- leaves
    for (;;) {
      if (!__any_sync(FULL_MASK, leaf != kNone)) break;
      if (leaf != kNone) {
        const float4 *t = reinterpret_cast<const float4 *>(tris + (size_t)(~leaf));
        for (;;) {
          const float4 a = __ldg(t), b = __ldg(t + 1), cc = __ldg(t + 2);
          if (COUNT) n_prims++;
          tri_test2(c, opt, a, b, cc, best);
          if (__float_as_uint(b.w) != 0u) break;
          t += 3;
        }
        leaf = kNone;
        if (cur < 0 && cur != kNone) {
          leaf = cur;
          cur = pop();
        }
      }
    }

    // ---- retire: the epilogue (store the hit / spawn the AO ray / accumulate) runs warp-wide
    const bool retiring = ray_idx >= 0 && cur == kNone && leaf == kNone;
    if (__any_sync(FULL_MASK, retiring)) epi(retiring, (size_t)ray_idx, best.t, best.u, best.v, best.prim, max_t, nullptr);
    if (retiring) ray_idx = -1;
  }

  if (COUNT) {
    for (int o = 16; o > 0; o >>= 1) {
      n_boxes += __shfl_down_sync(FULL_MASK, n_boxes, o);
      n_prims += __shfl_down_sync(FULL_MASK, n_prims, o);
    }
    if (lane == 0) {
      atomicAdd(counts + 0, n_boxes);
      atomicAdd(counts + 1, n_prims);
    }
  }
}

// ------------------------------------------------------------------ launchers
static int g_sm_count[64] = {0};
int device_sm_count(int device) {
  if (device < 0 || device >= 64) return 148;
  if (g_sm_count[device] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || v <= 0) v = 148;
    g_sm_count[device] = v;
  }
  return g_sm_count[device];
}

template <class Rays, int LOCAL_DEPTH, bool COUNT, class P, class Epi>
static cudaError_t launch_fast2(const Accel *a, Rays rays, size_t n, Epi epi, const TraceOptions16 &opt,
                                uint32_t flags, unsigned long long *cursor, unsigned long long *d_counts,
                                const unsigned long long *n_ptr, cudaStream_t s) {
  const int sms = device_sm_count(a->device);
  const size_t warps_per_block = P::kBlock / 32;
  size_t grid = (size_t)sms * P::kMinBlocks;
  const size_t need_blocks = ((n + 31) / 32 + warps_per_block - 1) / warps_per_block;
  if (grid > need_blocks) grid = need_blocks;
  if (grid == 0) grid = 1;
  traverse_fast2_kernel<Rays, LOCAL_DEPTH, COUNT, P, Epi><<<(unsigned)grid, P::kBlock, 0, s>>>(
      a->d_wide, a->d_tris, rays, n, epi, opt, flags, cursor, d_counts, n_ptr, (int)a->n_top);
  return cudaGetLastError();
}

// Round-1 policy, kept as experiment variants 100+ (profiles/r01_variant_sweep.md): 128-thread CTAs, 10 per SM
// (48 registers), refill when 16 lanes retired, node phase ends below 8 descending lanes, no TMA treelet, whole
// per-lane stack in thread-local memory.
typedef FastPolicy<128, 10, 16, 8, 0, 0> OldDefaultPolicy;
// Round-2 defaults (profiles/r02_variant_sweep.md): 128-thread CTAs, 10 per SM (48 registers), refill when 16 lanes
// retired, node phase ends below 8 descending lanes.  Coherent launches (camera rays, caller-supplied rays) read the
// 128-byte PairNode (no selects: they are issue bound); incoherent launches (AO, shadow and bounce rays) and trees
// whose PairNode array would not stay L2-resident read the 64-byte WideNode (they are bound by the L1 data pipe and
// by cache capacity, and the ALU pipe has room for the 12 selects).
// Leaf batching (profiles/r02_leaf_batching_sweep.md): after the first leaf round of an outer iteration another one
// runs only while >= 8 (coherent) / >= 12 (incoherent) lanes hold a leaf -- a lane's second leaf otherwise costs a
// round of its own at ~5 active lanes; the camera-ray launch, whose retire step spawns the AO ray (630 instructions),
// also waits with the retire step until retired + empty lanes reach the refill threshold.
// Two node steps per evaluation of the node phase's exit conditions (+1.8 ... +2.8 %, same sweep file).
// Refill threshold (same file, "refill" section): with deferred retire the camera launch does best when a warp takes a
// whole new 8x4-pixel packet only after ALL its lanes have finished (32) -- its rays have similar lengths, and the AO rays
// it spawns then reach the AO queue in packets of neighbouring pixels, which speeds the AO launch up too; caller-supplied
// rays 24; incoherent launches 20 (28 and more lose 10-50 % there).
typedef Policy3<128, 10, 24, 8, true, false, 8, 1, false, 2> DefaultPolicy;
// The camera launch runs 9 CTAs per SM (56 registers: its AO-spawn retire step spills at 48) and, like the incoherent
// launches, three node steps per exit check.
typedef Policy3<128, 9, 32, 8, true, false, 12, 1, true, 3> CameraPolicy;
typedef Policy3<128, 10, 20, 8, false, false, 12, 1, false, 3> IncoherentPolicy;
typedef Policy3<128, 9, 32, 8, false, false, 12, 1, true, 3> IncoherentCameraPolicy;
// PairNode arrays above this size are not used (126 MB L2; the triangles want their share)
constexpr size_t kPair128MaxBytes = (size_t)96 << 20;

static unsigned long long *next_cursor(const Accel *a, cudaStream_t s, cudaError_t *e) {
  // ring of 32 cursors: launches in flight on different streams never share one.  More than 32 traversal launches
  // of ONE accel in flight at once would alias; every entry point of this library keeps at most 3.
  unsigned long long *cursor =
      reinterpret_cast<unsigned long long *>(a->d_counters) + 16 + (a->cursor_ring.fetch_add(1) & 31u);
  *e = cudaMemsetAsync(cursor, 0, sizeof(unsigned long long), s);
  return cursor;
}

template <class Rays, int DEPTH, bool COUNT, class P, class Epi>
static cudaError_t launch_fast3(const Accel *a, Rays rays, size_t n, Epi epi, const TraceOptions16 &opt,
                                uint32_t flags, unsigned long long *cursor, unsigned long long *d_counts,
                                const unsigned long long *n_ptr, cudaStream_t s) {
  const int sms = device_sm_count(a->device);
  const size_t warps_per_block = P::kBlock / 32;
  size_t grid = (size_t)sms * P::kMinBlocks;  // persistent: every SM holds its full complement of CTAs
  const size_t need_blocks = ((n + 31) / 32 + warps_per_block - 1) / warps_per_block;
  if (grid > need_blocks) grid = need_blocks;
  if (grid == 0) grid = 1;
  const void *nodes = P::kPair128 ? static_cast<const void *>(a->d_pair) : static_cast<const void *>(a->d_wide);
  // experiment knob (tools/trav_sweep.py): unused dynamic shared memory per CTA, i.e. that much less L1 per SM --
  // measures how much a shared-memory staging scheme would cost before it saves anything
  static const size_t pad = getenv("NRT_SMEM_PAD") ? (size_t)atoi(getenv("NRT_SMEM_PAD")) : 0;
  if (pad > 48 * 1024)
    cudaFuncSetAttribute(traverse_fast3_kernel<Rays, DEPTH, COUNT, P, Epi>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
  traverse_fast3_kernel<Rays, DEPTH, COUNT, P, Epi><<<(unsigned)grid, P::kBlock, pad, s>>>(
      nodes, a->d_tris_cm, rays, n, epi, opt, flags, cursor, d_counts, n_ptr);
  return cudaGetLastError();
}

// a child pair pushes at most one entry and descends one level: the stack never holds more entries than the tree
// has levels.  Two instantiations: 64 entries (every tree the production builder emits) and 512 (the reference's
// kNANORT_MAX_STACK_DEPTH, for adopted trees; nrt_build / nrt_adopt reject deeper ones).
static bool needs_deep_stack(const Accel *a) { return a->stats.max_tree_depth + 2 > 64u; }

template <class Rays, bool COUNT, class P, class Epi>
static int launch_fast3_any(const Accel *a, Rays rays, size_t n, Epi epi, const TraceOptions16 &opt, uint32_t flags,
                            unsigned long long *d_counts, const unsigned long long *n_ptr, cudaStream_t s);

// the default for coherent launches, with the size cut-off
template <class Rays, bool COUNT, class Epi>
static int launch_fast3_coherent(const Accel *a, Rays rays, size_t n, Epi epi, const TraceOptions16 &opt, uint32_t flags,
                                 unsigned long long *d_counts, const unsigned long long *n_ptr, cudaStream_t s) {
  if (a->n_wide * sizeof(PairNode) > kPair128MaxBytes)
    return launch_fast3_any<Rays, COUNT, IncoherentPolicy>(a, rays, n, epi, opt, flags, d_counts, n_ptr, s);
  return launch_fast3_any<Rays, COUNT, DefaultPolicy>(a, rays, n, epi, opt, flags, d_counts, n_ptr, s);
}

template <class Rays, bool COUNT, class P, class Epi>
static int launch_fast3_any(const Accel *a, Rays rays, size_t n, Epi epi, const TraceOptions16 &opt, uint32_t flags,
                            unsigned long long *d_counts, const unsigned long long *n_ptr, cudaStream_t s) {
  if (!a->d_pair || !a->d_tris_cm) {
    set_error("traverse: this accel has no triangle traversal layout");
    return NRT_ERR_INVALID;
  }
  cudaError_t e;
  unsigned long long *cursor = next_cursor(a, s, &e);
  NRT_CUDA(e);
  if (needs_deep_stack(a))
    e = launch_fast3<Rays, 512, COUNT, P>(a, rays, n, epi, opt, flags, cursor, d_counts, n_ptr, s);
  else
    e = launch_fast3<Rays, 64, COUNT, P>(a, rays, n, epi, opt, flags, cursor, d_counts, n_ptr, s);
  NRT_CUDA(e);
  return NRT_OK;
}

template <class Rays, bool COUNT>
static int launch_fast(const Accel *a, Rays rays, size_t n, Hit16 *d_hits, uint8_t *d_mask,
                       const TraceOptions16 &opt, uint32_t flags, unsigned long long *d_counts, cudaStream_t s,
                       const unsigned long long *n_ptr = nullptr) {
  const uint32_t variant = (flags >> 8) & 0xFFu;  // experiment selector (tools/trav_sweep.py); 0 = default
  const StoreHitsEpilogue epi{d_hits, d_mask};
  if (variant == 0 || COUNT) return launch_fast3_coherent<Rays, COUNT>(a, rays, n, epi, opt, flags, d_counts, n_ptr, s);
  if (variant < 100) {
#define NRT_VARIANT3(id, ...)                                                                                          \
  case id:                                                                                                            \
    return launch_fast3_any<Rays, false, Policy3<__VA_ARGS__> >(a, rays, n, epi, opt, flags, d_counts, n_ptr, s);
    switch (variant) {
      NRT_VARIANT3(1, 128, 8, 16, 8)
      NRT_VARIANT3(2, 128, 9, 16, 8)
      NRT_VARIANT3(3, 128, 12, 16, 8)
      NRT_VARIANT3(4, 128, 10, 8, 8)
      NRT_VARIANT3(5, 128, 10, 24, 8)
      NRT_VARIANT3(6, 128, 10, 16, 4)
      NRT_VARIANT3(7, 128, 10, 16, 12)
      NRT_VARIANT3(8, 128, 10, 16, 0)
      NRT_VARIANT3(9, 256, 5, 16, 8)
      NRT_VARIANT3(10, 64, 20, 16, 8)
      NRT_VARIANT3(11, 128, 7, 16, 8)
      NRT_VARIANT3(12, 128, 6, 16, 8)
      NRT_VARIANT3(20, 128, 10, 16, 8, false)
      NRT_VARIANT3(21, 128, 9, 16, 8, false)
      NRT_VARIANT3(22, 128, 8, 16, 8, false)
      NRT_VARIANT3(23, 128, 10, 16, 4, false)
      NRT_VARIANT3(19, 128, 10, 16, 8, true)  // PairNode regardless of the size cut-off
      NRT_VARIANT3(30, 128, 10, 16, 8, false, true)  // 2 x LDG.256 per node: measured 5-7 % slower than 4 loads
      // leaf batching / retire batching knobs <.., leaf-again-min, leaf slots, deferred retire>; 4x/6x = PairNode, 5x/7x = WideNode
      NRT_VARIANT3(40, 128, 10, 16, 8, true, false, 33, 1, false)
      NRT_VARIANT3(41, 128, 10, 16, 8, true, false, 33, 2, false)
      NRT_VARIANT3(42, 128, 10, 16, 8, true, false, 1, 1, true)
      NRT_VARIANT3(43, 128, 10, 16, 8, true, false, 33, 1, true)
      NRT_VARIANT3(44, 128, 10, 16, 8, true, false, 33, 2, true)
      NRT_VARIANT3(45, 128, 10, 16, 8, true, false, 1, 2, false)
      NRT_VARIANT3(46, 128, 10, 16, 12, true, false, 33, 1, false)
      NRT_VARIANT3(47, 128, 10, 16, 16, true, false, 33, 1, false)
      NRT_VARIANT3(48, 128, 10, 16, 12, true, false, 33, 2, false)
      NRT_VARIANT3(49, 128, 10, 16, 16, true, false, 33, 2, false)
      NRT_VARIANT3(50, 128, 10, 16, 8, false, false, 33, 1, false)
      NRT_VARIANT3(51, 128, 10, 16, 8, false, false, 33, 2, false)
      NRT_VARIANT3(52, 128, 10, 16, 8, false, false, 1, 1, true)
      NRT_VARIANT3(53, 128, 10, 16, 8, false, false, 33, 1, true)
      NRT_VARIANT3(54, 128, 10, 16, 8, false, false, 33, 2, true)
      NRT_VARIANT3(55, 128, 10, 16, 8, false, false, 1, 2, false)
      NRT_VARIANT3(56, 128, 10, 16, 12, false, false, 33, 1, false)
      NRT_VARIANT3(57, 128, 10, 16, 16, false, false, 33, 1, false)
      NRT_VARIANT3(58, 128, 10, 16, 12, false, false, 33, 2, false)
      NRT_VARIANT3(59, 128, 10, 16, 16, false, false, 33, 2, false)
      NRT_VARIANT3(60, 128, 10, 16, 8, true, false, 8, 1, false)
      NRT_VARIANT3(61, 128, 10, 16, 8, true, false, 12, 1, false)
      NRT_VARIANT3(62, 128, 10, 16, 8, true, false, 16, 1, false)
      NRT_VARIANT3(63, 128, 10, 16, 8, true, false, 8, 2, false)
      NRT_VARIANT3(64, 128, 10, 16, 8, true, false, 12, 2, false)
      NRT_VARIANT3(65, 128, 10, 16, 8, true, false, 16, 2, false)
      NRT_VARIANT3(66, 128, 10, 16, 8, true, false, 24, 2, false)
      NRT_VARIANT3(67, 128, 10, 16, 8, true, false, 12, 2, true)
      NRT_VARIANT3(68, 128, 10, 16, 12, true, false, 12, 2, false)
      NRT_VARIANT3(69, 128, 10, 16, 4, true, false, 33, 2, false)
      NRT_VARIANT3(70, 128, 10, 16, 8, false, false, 8, 1, false)
      NRT_VARIANT3(71, 128, 10, 16, 8, false, false, 12, 1, false)
      NRT_VARIANT3(72, 128, 10, 16, 8, false, false, 16, 1, false)
      NRT_VARIANT3(73, 128, 10, 16, 8, false, false, 8, 2, false)
      NRT_VARIANT3(74, 128, 10, 16, 8, false, false, 12, 2, false)
      NRT_VARIANT3(75, 128, 10, 16, 8, false, false, 16, 2, false)
      NRT_VARIANT3(76, 128, 10, 16, 8, false, false, 24, 2, false)
      NRT_VARIANT3(77, 128, 10, 16, 8, false, false, 12, 2, true)
      NRT_VARIANT3(78, 128, 10, 16, 12, false, false, 12, 2, false)
      NRT_VARIANT3(79, 128, 10, 16, 4, false, false, 33, 2, false)
      // node steps per exit check (8x = PairNode, 9x = WideNode)
      NRT_VARIANT3(80, 128, 10, 16, 8, true, false, 8, 1, false, 2)
      NRT_VARIANT3(81, 128, 10, 16, 8, true, false, 8, 1, false, 3)
      NRT_VARIANT3(82, 128, 10, 16, 8, true, false, 8, 1, false, 4)
      NRT_VARIANT3(83, 128, 10, 24, 8, true, false, 8, 1, false, 2)
      NRT_VARIANT3(84, 128, 10, 32, 8, true, false, 8, 1, false, 2)
      NRT_VARIANT3(85, 128, 10, 32, 8, true, false, 8, 1, true, 2)
      NRT_VARIANT3(93, 128, 10, 24, 8, false, false, 12, 1, false, 2)
      NRT_VARIANT3(94, 128, 10, 32, 8, false, false, 12, 1, false, 2)
      NRT_VARIANT3(90, 128, 10, 16, 8, false, false, 12, 1, false, 2)
      NRT_VARIANT3(91, 128, 10, 16, 8, false, false, 12, 1, false, 3)
      NRT_VARIANT3(92, 128, 10, 16, 8, false, false, 12, 1, false, 4)
      default:
        set_error("nrt_traverse: unknown kernel variant in flags");
        return NRT_ERR_INVALID;
    }
#undef NRT_VARIANT3
  }
  // ---- round-1 kernel (64-byte WideNode + vertex-major PackedTri), kept for A/B runs
  cudaError_t e;
  unsigned long long *cursor = next_cursor(a, s, &e);
  NRT_CUDA(e);
  const bool deep = a->stats.max_tree_depth + 2 > (uint32_t)(kStackSmem + 48);
  if (deep) {
    e = launch_fast2<Rays, 512 - kStackSmem, false, OldDefaultPolicy>(a, rays, n, epi, opt, flags, cursor, d_counts, n_ptr, s);
  } else {
#define NRT_VARIANT(id, ...)                                                                                        \
  case id:                                                                                                          \
    e = launch_fast2<Rays, 48, false, FastPolicy<__VA_ARGS__> >(a, rays, n, epi, opt, flags, cursor, d_counts,      \
                                                                n_ptr, s);                                          \
    break;
    switch (variant) {
      NRT_VARIANT(100, 128, 10, 16, 8, 0, 0)
      NRT_VARIANT(130, 128, 10, 16, 8, 64)
      NRT_VARIANT(131, 128, 9, 16, 8, 128)
      NRT_VARIANT(140, 128, 10, 16, 8, 0, 8)
      NRT_VARIANT(142, 128, 10, 16, 8, 0, 4)
      default:
        set_error("nrt_traverse: unknown kernel variant in flags");
        return NRT_ERR_INVALID;
    }
#undef NRT_VARIANT
  }
  NRT_CUDA(e);
  return NRT_OK;
}

template <class Rays, bool COUNT>
static int launch_conf(const Accel *a, Rays rays, size_t n, Hit16 *d_hits, uint8_t *d_mask,
                       const TraceOptions16 &opt, uint32_t flags, unsigned long long *d_counts, cudaStream_t s,
                       const unsigned long long *n_ptr = nullptr) {
  size_t grid = (n + 127) / 128;
  if (grid == 0) return NRT_OK;
  traverse_conformance_kernel<Rays, COUNT><<<(unsigned)grid, 128, 0, s>>>(a->d_nodes, a->d_tris, rays, n, d_hits,
                                                                       d_mask, opt, flags, d_counts, n_ptr);
  NRT_CUDA(cudaGetLastError());
  return NRT_OK;
}

int launch_traverse(const Accel *a, const Ray36 *d_rays, size_t n, Hit16 *d_hits, uint8_t *d_mask,
                    const TraceOptions16 &opt, uint32_t flags, cudaStream_t s) {
  if (n == 0) return NRT_OK;
  if (flags & NRT_TRAVERSE_RAY32) {  // compact records: default policy only (the experiment variants read Ray36)
    if (a->prim_kind != 0 || (reinterpret_cast<uintptr_t>(d_rays) & 15u) != 0) {
      set_error("nrt_traverse: NRT_TRAVERSE_RAY32 needs a triangle accel and 16-byte aligned rays");
      return NRT_ERR_INVALID;
    }
    Aos32Rays r32{reinterpret_cast<const float4 *>(d_rays)};
    if (flags & NRT_TRAVERSE_CONFORMANCE)
      return launch_conf<Aos32Rays, false>(a, r32, n, d_hits, d_mask, opt, flags, nullptr, s);
    const StoreHitsEpilogue epi{d_hits, d_mask};
    if (flags & NRT_TRAVERSE_ANY_HIT)
      return launch_fast3_coherent<Aos32Rays, false>(a, r32, n, AnyHit<StoreHitsEpilogue>(epi), opt, flags, nullptr, nullptr, s);
    return launch_fast3_coherent<Aos32Rays, false>(a, r32, n, epi, opt, flags, nullptr, nullptr, s);
  }
  if (a->prim_kind != 0) return launch_traverse_prims(a, d_rays, n, d_hits, d_mask, opt, flags, s);  // spheres ...
  AosRays r{d_rays};
  if (flags & NRT_TRAVERSE_CONFORMANCE) return launch_conf<AosRays, false>(a, r, n, d_hits, d_mask, opt, flags, nullptr, s);
  if (flags & NRT_TRAVERSE_ANY_HIT)  // occlusion query: default policy only
    return launch_fast3_coherent<AosRays, false>(a, r, n, AnyHit<StoreHitsEpilogue>(StoreHitsEpilogue{d_hits, d_mask}), opt,
                                                 flags, nullptr, nullptr, s);
  return launch_fast<AosRays, false>(a, r, n, d_hits, d_mask, opt, flags, nullptr, s);
}

int launch_traverse_soa(const Accel *a, const float4 *d_org_tmin, const float4 *d_dir_tmax, size_t n,
                        Hit16 *d_hits, const TraceOptions16 &opt, uint32_t flags, cudaStream_t s) {
  if (n == 0) return NRT_OK;
  SoaRays r{d_org_tmin, d_dir_tmax};
  if (flags & NRT_TRAVERSE_CONFORMANCE) return launch_conf<SoaRays, false>(a, r, n, d_hits, nullptr, opt, flags, nullptr, s);
  return launch_fast<SoaRays, false>(a, r, n, d_hits, nullptr, opt, flags, nullptr, s);
}

// `capacity` bounds the count stored at d_count (grid sizing only).
int launch_traverse_soa_devcount(const Accel *a, const float4 *d_org_tmin, const float4 *d_dir_tmax,
                                 const unsigned long long *d_count, size_t capacity, Hit16 *d_hits,
                                 const TraceOptions16 &opt, uint32_t flags, cudaStream_t s) {
  if (capacity == 0) return NRT_OK;
  SoaRays r{d_org_tmin, d_dir_tmax};
  if (flags & NRT_TRAVERSE_CONFORMANCE)
    return launch_conf<SoaRays, false>(a, r, capacity, d_hits, nullptr, opt, flags, nullptr, s, d_count);
  return launch_fast<SoaRays, false>(a, r, capacity, d_hits, nullptr, opt, flags, nullptr, s, d_count);
}

// Experiment selector for the fused launches of the AO / path passes (tools/ao_exp_sweep.py): NRT_AO_EXP="<p><a><r><s>",
// one digit each for the camera-ray launch, the AO launch, the path tracer's radiance launch and its shadow launch;
// 0 = the default policy.  Read at every launch (cheap) so that one
// process can sweep.
//   variations of the launch's default policy:  1 / 2: node-phase exit below 6 / 12   3 / 4: leaf-again-min 8 / 16
//   5: three node steps per exit check   6 / 7: one / two CTAs per SM less   8: refill at 16   9: (5) + (6)
static int ao_exp(int which) {
  const char *e = getenv("NRT_AO_EXP");
  if (!e || !e[0]) return 0;
  for (int k = 0; k < which; ++k)
    if (!e[k + 1]) return 0;
  const int d = e[which] - '0';
  return (d < 0 || d > 9) ? 0 : d;
}
#define NRT_AO_EXP_SWITCH(which, MINB, REFILL, PAIR, AGAIN, DEFER, CALL)                                         \
  switch (ao_exp(which)) {                                                                                     \
    case 1: { typedef Policy3<128, MINB, REFILL, 6, PAIR, false, AGAIN, 1, DEFER, 2> PX; return CALL; }          \
    case 2: { typedef Policy3<128, MINB, REFILL, 12, PAIR, false, AGAIN, 1, DEFER, 2> PX; return CALL; }         \
    case 3: { typedef Policy3<128, MINB, REFILL, 8, PAIR, false, 8, 1, DEFER, 2> PX; return CALL; }              \
    case 4: { typedef Policy3<128, MINB, REFILL, 8, PAIR, false, 16, 1, DEFER, 2> PX; return CALL; }             \
    case 5: { typedef Policy3<128, MINB, REFILL, 8, PAIR, false, AGAIN, 1, DEFER, 3> PX; return CALL; }          \
    case 6: { typedef Policy3<128, MINB - 1, REFILL, 8, PAIR, false, AGAIN, 1, DEFER, 2> PX; return CALL; }      \
    case 7: { typedef Policy3<128, MINB - 2, REFILL, 8, PAIR, false, AGAIN, 1, DEFER, 2> PX; return CALL; }      \
    case 8: { typedef Policy3<128, MINB, 16, 8, PAIR, false, AGAIN, 1, DEFER, 2> PX; return CALL; }              \
    case 9: { typedef Policy3<128, MINB - 1, REFILL, 8, PAIR, false, AGAIN, 1, DEFER, 3> PX; return CALL; }      \
    default: break;                                                                                            \
  }

// Fused wavefront launches (render.cu): the retire step spawns the AO ray / accumulates visibility.
template <class Epi, class P = DefaultPolicy, class Rays = SoaRays>
static int launch_fused(const Accel *a, Rays rays, size_t n, const unsigned long long *n_ptr, Epi epi,
                        const TraceOptions16 &opt, uint32_t flags, cudaStream_t s) {
  if (n == 0) return NRT_OK;
  return launch_fast3_any<Rays, false, P>(a, rays, n, epi, opt, flags, nullptr, n_ptr, s);
}

int launch_traverse_primary_fused(const Accel *a, const Wave &w, const nrt_ao_params &p, unsigned long long slot0,
                                  size_t count, float *d_accum, unsigned long long *d_wave_counters,
                                  const TraceOptions16 &opt, uint32_t flags, cudaStream_t s) {
  PrimaryToAoEpilogue<false> epi{p, slot0, w, a->d_verts, a->d_faces, d_accum, d_wave_counters};
  return launch_fused(a, SoaRays{w.org_tmin, w.dir_tmax}, count, nullptr, epi, opt, flags, s);
}

// Same, with the camera rays generated inside the kernel (no generator kernel, no primary queue)
int launch_traverse_camera_fused(const Accel *a, const Wave &w, const nrt_ao_params &p, unsigned long long slot0,
                                 size_t count, float *d_accum, unsigned long long *d_wave_counters,
                                 const TraceOptions16 &opt, uint32_t flags, cudaStream_t s) {
  PrimaryToAoEpilogue<true> epi{p, slot0, w, a->d_verts, a->d_faces, d_accum, d_wave_counters};
  if (count == 0) return NRT_OK;
  if (a->n_wide * sizeof(PairNode) <= kPair128MaxBytes) {
    NRT_AO_EXP_SWITCH(0, 9, 32, true, 12, true, (launch_fast3_any<CameraRays, false, PX>(a, CameraRays(p, slot0), count, epi, opt, flags,
                                                                        nullptr, nullptr, s)))
  }
  if (a->n_wide * sizeof(PairNode) > kPair128MaxBytes)
    return launch_fast3_any<CameraRays, false, IncoherentCameraPolicy>(a, CameraRays(p, slot0), count, epi, opt, flags, nullptr,
                                                                       nullptr, s);
  return launch_fast3_any<CameraRays, false, CameraPolicy>(a, CameraRays(p, slot0), count, epi, opt, flags, nullptr, nullptr, s);
}

int launch_traverse_ao_fused(const Accel *a, const Wave &w, const unsigned long long *d_count, size_t capacity,
                             float *d_accum, unsigned long long *d_totals, const TraceOptions16 &opt, uint32_t flags,
                             cudaStream_t s) {
  AoAccumulateEpilogue epi{w.ao_pix, d_accum, d_totals};
  if (flags & NRT_TRAVERSE_ANY_HIT)
    return launch_fused<AnyHit<AoAccumulateEpilogue>, IncoherentPolicy>(a, SoaRays{w.ao_org_tmin, w.ao_dir_tmax}, capacity, d_count,
                                                                        AnyHit<AoAccumulateEpilogue>(epi), opt, flags, s);
  NRT_AO_EXP_SWITCH(1, 10, 20, false, 12, false, (launch_fused<AoAccumulateEpilogue, PX>(a, SoaRays{w.ao_org_tmin, w.ao_dir_tmax}, capacity,
                                                                      d_count, epi, opt, flags, s)))
  return launch_fused<AoAccumulateEpilogue, IncoherentPolicy>(a, SoaRays{w.ao_org_tmin, w.ao_dir_tmax}, capacity, d_count, epi,
                                                              opt, flags, s);
}

int launch_traverse_path_radiance(const Accel *a, const PathShadeEpilogue &epi, const unsigned long long *d_count,
                                  size_t capacity, const TraceOptions16 &opt, uint32_t flags, cudaStream_t s) {
  // the shading block needs more registers than the plain traversal: 8 CTAs/SM (64 registers) instead of 10
  NRT_AO_EXP_SWITCH(2, 8, 24, false, 8, true, (launch_fused<PathShadeEpilogue, PX>(a, SoaRays{epi.q.org_tmin[epi.in], epi.q.dir_tmax[epi.in]},
                                                                      capacity, d_count, epi, opt, flags, s)))
  // its retire step IS the shading block: deferred retire (it runs with more lanes), profiles/r02_leaf_batching_sweep.md
  return launch_fused<PathShadeEpilogue, Policy3<128, 8, 24, 8, false, false, 8, 1, true, 2> >(
      a, SoaRays{epi.q.org_tmin[epi.in], epi.q.dir_tmax[epi.in]}, capacity, d_count, epi, opt, flags, s);
}

int launch_traverse_path_shadow(const Accel *a, const PathQueues &q, const unsigned long long *d_count,
                                size_t capacity, float *d_accum, const TraceOptions16 &opt, uint32_t flags,
                                cudaStream_t s) {
  ShadowAccumulateEpilogue epi{q.sh_contrib_pix, d_accum};
  if (flags & NRT_TRAVERSE_ANY_HIT)
    return launch_fused<AnyHit<ShadowAccumulateEpilogue>, IncoherentPolicy>(a, SoaRays{q.sh_org_tmin, q.sh_dir_tmax}, capacity,
                                                                            d_count, AnyHit<ShadowAccumulateEpilogue>(epi), opt,
                                                                            flags, s);
  NRT_AO_EXP_SWITCH(3, 10, 20, false, 12, false, (launch_fused<ShadowAccumulateEpilogue, PX>(a, SoaRays{q.sh_org_tmin, q.sh_dir_tmax}, capacity,
                                                                           d_count, epi, opt, flags, s)))
  return launch_fused<ShadowAccumulateEpilogue, IncoherentPolicy>(a, SoaRays{q.sh_org_tmin, q.sh_dir_tmax}, capacity, d_count,
                                                                  epi, opt, flags, s);
}

int launch_traverse_count(const Accel *a, const Ray36 *d_rays, size_t n, const TraceOptions16 &opt,
                          uint32_t flags, uint64_t *d_counts2, cudaStream_t s) {
  NRT_CUDA(cudaMemsetAsync(d_counts2, 0, 16 * sizeof(uint64_t), s));  // [0] boxes, [1] prims, [2..15] lane statistics
  if (n == 0) return NRT_OK;
  AosRays r{d_rays};
  unsigned long long *cnt = reinterpret_cast<unsigned long long *>(d_counts2);
  if (flags & NRT_TRAVERSE_CONFORMANCE) return launch_conf<AosRays, true>(a, r, n, nullptr, nullptr, opt, flags, cnt, s);
  return launch_fast<AosRays, true>(a, r, n, nullptr, nullptr, opt, flags, cnt, s);
}

}  // namespace nrt
