// Reference-exact builder for BVHAccel<double> ("conformance build", T = double): produces, on the device, the very
// BVHNode<double> array and indices_ that CPU nanort's BVHAccel<double>::Build writes at the pinned commit -- bit for bit.
// It is build_ref.cu restated for double (same level-synchronous structure, same quirks: x-only binning from the guard
// at nanort.h:1357, try-next-axis / object-median fallback nanort.h:1827-1857, TriangleSAHPred + libstdc++'s
// std::partition element order, serial or C++11-joined node order); what changes with T is every piece of arithmetic
// that decides a split: bounding boxes, bin indices, bin boxes, SAH costs, cut positions and the predicate are
// computed in double (ContributeBinBuffer nanort.h:1314-1367, FindCutFromBinBuffer :1381-1430, CalculateSurfaceArea
// :1278-1283, BoundingBoxAndCenter :958-971, TriangleSAHPred :897-911), min/max boxes travel as order-preserving 64-bit
// keys through atomicMin / atomicMax.  Checked bit for bit against the test suite's fp64 checker (itself pinned to the
// unmodified reference's BVHAccel<double>) by tests/test_gpu_f64.py.  Speed is not a goal.
#include <algorithm>
#include <vector>

#include "common.cuh"
#include "scan.cuh"

namespace nrt {
namespace {

constexpr uint32_t kInactive = 0xFFFFFFFFu;
constexpr int kBinWords = 8;  // count, min xyz, max xyz, pad -- 64-bit words here

struct alignas(32) D4 {
  double x, y, z, w;
};
struct alignas(16) D2 {
  double x, y;
};
__device__ __forceinline__ D4 make_d4(double x, double y, double z, double w) {
  D4 r;
  r.x = x, r.y = y, r.z = z, r.w = w;
  return r;
}
__device__ __forceinline__ D2 make_d2(double x, double y) {
  D2 r;
  r.x = x, r.y = y;
  return r;
}

struct Node64 {  // BVHNode<double> (nanort.h:527-569 with T = double)
  double bmin[3], bmax[3];
  int32_t flag, axis;
  uint32_t data[2];
};
static_assert(sizeof(Node64) == 64, "BVHNode<double> layout");

struct BNodeD {  // build node, double boxes; integer fields as in build_common.cuh:BNode
  double bmin[3];
  double bmax[3];
  uint32_t l, r;
  uint32_t left;   // pool index of the left child (right = left + 1); kInactive for a leaf
  uint32_t depth;
  uint32_t rturns;  // right turns on the root path
  uint32_t axis;
  uint32_t split_bin;
  uint32_t nleft;
  uint32_t slot;
  uint32_t pad;
  uint32_t pad2[2];
};
static_assert(sizeof(BNodeD) == 96, "BNodeD");

typedef unsigned long long key64;
// order-preserving double <-> uint64 key for atomicMin / atomicMax
__device__ __forceinline__ key64 dkey(double f) {
  const key64 u = (key64)__double_as_longlong(f);
  return (u & 0x8000000000000000ull) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double dunkey(key64 k) {
  return __longlong_as_double((long long)((k & 0x8000000000000000ull) ? (k ^ 0x8000000000000000ull) : ~k));
}
constexpr key64 kKeyMax = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ double box_area_d(double lx, double ly, double lz, double hx, double hy, double hz) {
  const double dx = hx - lx, dy = hy - ly, dz = hz - lz;
  return 2.0 * ((dx * dy + dy * dz) + dz * dx);  // CalculateSurfaceArea, nanort.h:1278-1283
}
__device__ __forceinline__ int bin_of_d(double c, double nmin, double inv, int B) {
  const double q = (c - nmin) * inv;
  int qi = (int)q;  // truncation, as the reference's int(quantized_center[j])
  qi = qi < 0 ? 0 : qi;
  return qi > B - 1 ? B - 1 : qi;
}
__device__ __forceinline__ double inv_extent_d(double lo, double hi, int B) {
  const double sz = hi - lo;
  return sz > 0.0 ? (double)B / sz : 0.0;
}

// FindCutFromBinBuffer for one axis (nanort.h:1381-1430), sequential as the reference writes it -- one lane per node is
// enough here (bin_size <= 256 iterations twice): right-to-left sweep stores count * area per bin, left-to-right sweep
// adds the left side; strict `<`, first minimum wins; a side without primitives makes the cost NaN (0 * inf) and never
// wins.  Returns minBin (1 when no boundary separates the centroids).
__device__ int find_cut_d(const key64 *bins, int B, double *cost_scratch) {
  const double kMax = 1.7976931348623157e308;
  double lo[3] = {kMax, kMax, kMax}, hi[3] = {-kMax, -kMax, -kMax};
  size_t count = 0;
  for (int i = B - 1; i > 0; --i) {
    const key64 *w = bins + (size_t)i * kBinWords;
    if (w[0] != 0ull) {
      for (int k = 0; k < 3; k++) {
        lo[k] = fmin(dunkey(w[1 + k]), lo[k]);
        hi[k] = fmax(dunkey(w[4 + k]), hi[k]);
      }
    }
    count += (size_t)w[0];
    cost_scratch[i] = (double)count * box_area_d(lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]);
  }
  count = 0;
  for (int k = 0; k < 3; k++) {
    lo[k] = kMax;
    hi[k] = -kMax;
  }
  double min_cost = kMax;
  int min_bin = 1;
  for (int i = 0; i < B - 1; i++) {
    const key64 *w = bins + (size_t)i * kBinWords;
    if (w[0] != 0ull) {
      for (int k = 0; k < 3; k++) {
        lo[k] = fmin(dunkey(w[1 + k]), lo[k]);
        hi[k] = fmax(dunkey(w[4 + k]), hi[k]);
      }
    }
    count += (size_t)w[0];
    const double cost = (double)count * box_area_d(lo[0], lo[1], lo[2], hi[0], hi[1], hi[2]) + cost_scratch[i + 1];
    if (cost < min_cost) {
      min_cost = cost;
      min_bin = i + 1;
    }
  }
  return min_bin;
}

// per primitive: A = (bmin.xyz, center.x), B = (bmax.xyz, sum.x), C = (sum.y, sum.z); sum = (p0+p1)+p2,
// center = sum * (T(1)/T(3))  (TriangleMesh<double>::BoundingBoxAndCenter, nanort.h:958-971)
__global__ void ref_prim_kernel(const double *__restrict__ verts, const uint32_t *__restrict__ faces, uint32_t n,
                                D4 *__restrict__ A, D4 *__restrict__ B, D2 *__restrict__ C) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t f0 = faces[3 * (size_t)i], f1 = faces[3 * (size_t)i + 1], f2 = faces[3 * (size_t)i + 2];
  const double *p0 = verts + 3 * (size_t)f0, *p1 = verts + 3 * (size_t)f1, *p2 = verts + 3 * (size_t)f2;
  double lo[3], hi[3], s[3];
  for (int k = 0; k < 3; k++) {
    lo[k] = fmin(p0[k], fmin(p1[k], p2[k]));
    hi[k] = fmax(p0[k], fmax(p1[k], p2[k]));
    s[k] = (p0[k] + p1[k]) + p2[k];
  }
  A[i] = make_d4(lo[0], lo[1], lo[2], s[0] * (1.0 / 3.0));
  B[i] = make_d4(hi[0], hi[1], hi[2], s[0]);
  C[i] = make_d2(s[1], s[2]);
}

struct RefCounters {
  uint32_t pool;
  uint32_t n_fresh[2];   // nodes created by the previous / this level (ping-pong)
  uint32_t n_active[2];  // the ones of them that will be split
  uint32_t pad[3];
};

// BNodeD fields as used here: split_bin = level in which the node was created, slot = its index in that level's
// fresh list, pad = index in the level's active list (kInactive for leaves), nleft = mid - l once split.

__global__ void ref_init_kernel(BNodeD *pool, RefCounters *ctr, uint32_t n, uint32_t min_leaf, uint32_t max_depth,
                                uint32_t *fresh0, uint32_t *active0) {
  BNodeD r;
  for (int k = 0; k < 3; k++) r.bmin[k] = r.bmax[k] = 0.0;
  r.l = 0;
  r.r = n;
  r.left = kInactive;
  r.depth = 0;
  r.rturns = 0;
  r.axis = 0;
  r.split_bin = 0;
  r.nleft = 0;
  r.slot = 0;
  r.pad = kInactive;
  r.pad2[0] = r.pad2[1] = 0;
  ctr->pool = 1;
  ctr->n_fresh[0] = 1;
  ctr->n_fresh[1] = 0;
  ctr->n_active[0] = ctr->n_active[1] = 0;
  fresh0[0] = 0;
  if (!(n <= min_leaf || 0 >= max_depth)) {
    active0[0] = 0;
    r.pad = 0;
    ctr->n_active[0] = 1;
  }
  pool[0] = r;
}

__global__ void ref_iota_kernel(uint32_t *idx, uint32_t *node_of, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    idx[i] = i;
    node_of[i] = 0;
  }
}

__global__ void ref_keys_init_kernel(key64 *keys, uint32_t n_fresh) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_fresh * 6) keys[i] = (i % 6) < 3 ? kKeyMax : 0ull;
}

// exact box of every fresh node: min/max over the boxes of its primitives (ComputeBoundingBox, nanort.h:1545-1567)
__global__ void __launch_bounds__(256)
    ref_bbox_kernel(const BNodeD *__restrict__ pool, const uint32_t *__restrict__ node_of,
                    const uint32_t *__restrict__ idx, const D4 *__restrict__ A, const D4 *__restrict__ B,
                    uint32_t n, uint32_t level, key64 *__restrict__ keys) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const BNodeD nd = pool[node_of[p]];
  if (nd.split_bin != level) return;
  const uint32_t s = idx[p];
  const D4 a = A[s], b = B[s];
  key64 *k = keys + (size_t)nd.slot * 6;
  atomicMin(k + 0, dkey(a.x));
  atomicMin(k + 1, dkey(a.y));
  atomicMin(k + 2, dkey(a.z));
  atomicMax(k + 3, dkey(b.x));
  atomicMax(k + 4, dkey(b.y));
  atomicMax(k + 5, dkey(b.z));
}

__global__ void ref_bbox_store_kernel(BNodeD *pool, const uint32_t *__restrict__ fresh, uint32_t n_fresh,
                                      const key64 *__restrict__ keys) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_fresh) return;
  BNodeD *nd = pool + fresh[i];
  for (int k = 0; k < 3; k++) {
    nd->bmin[k] = dunkey(keys[(size_t)i * 6 + k]);
    nd->bmax[k] = dunkey(keys[(size_t)i * 6 + 3 + k]);
  }
}

__global__ void ref_bins_clear_kernel(key64 *bins, size_t words) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < words) {
    int w = (int)(i & (kBinWords - 1));
    bins[i] = (w >= 1 && w <= 3) ? kKeyMax : 0ull;
  }
}

// x-axis bins only (the pinned commit's guard, nanort.h:1357): count + exact box per bin
__global__ void __launch_bounds__(256)
    ref_bins_kernel(const BNodeD *__restrict__ pool, const uint32_t *__restrict__ node_of,
                    const uint32_t *__restrict__ idx, const D4 *__restrict__ A, const D4 *__restrict__ B,
                    uint32_t n, uint32_t level, int nbins, key64 *__restrict__ bins) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const BNodeD nd = pool[node_of[p]];
  if (nd.split_bin != level || nd.pad == kInactive) return;
  const uint32_t s = idx[p];
  const D4 a = A[s], b = B[s];
  const int bx = bin_of_d(a.w, nd.bmin[0], inv_extent_d(nd.bmin[0], nd.bmax[0], nbins), nbins);
  key64 *w = bins + ((size_t)nd.pad * nbins + bx) * kBinWords;
  atomicAdd(w, 1ull);
  atomicMin(w + 1, dkey(a.x));
  atomicMin(w + 2, dkey(a.y));
  atomicMin(w + 3, dkey(a.z));
  atomicMax(w + 4, dkey(b.x));
  atomicMax(w + 5, dkey(b.y));
  atomicMax(w + 6, dkey(b.z));
}

// one THREAD per active node: the three candidate planes (nanort.h:1423) and the per-node round state
__global__ void __launch_bounds__(64)
    ref_cut_kernel(const BNodeD *__restrict__ pool, const uint32_t *__restrict__ active, uint32_t n_active,
                   const key64 *__restrict__ bins, int nbins, double *__restrict__ cost_scratch, double *__restrict__ cut3,
                   uint32_t *__restrict__ cnt, uint32_t *__restrict__ state) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_active) return;
  const BNodeD nd = pool[active[a]];
  const int min_bin_x = find_cut_d(bins + (size_t)a * nbins * kBinWords, nbins, cost_scratch + (size_t)a * nbins);
  const double fB = (double)nbins;
  cut3[(size_t)a * 3 + 0] = (double)min_bin_x * ((nd.bmax[0] - nd.bmin[0]) / fB) + nd.bmin[0];
  cut3[(size_t)a * 3 + 1] = 1.0 * ((nd.bmax[1] - nd.bmin[1]) / fB) + nd.bmin[1];
  cut3[(size_t)a * 3 + 2] = 1.0 * ((nd.bmax[2] - nd.bmin[2]) / fB) + nd.bmin[2];
  cnt[a] = 0;
  state[a] = 0;  // bit 31 = decided, bit 30 = partition needed, low bits = axis
}

// (p0+p1)+p2 < pos*3 (TriangleSAHPred<double>, nanort.h:897-911)
__device__ __forceinline__ bool ref_pred(const D4 &b, const D2 &c, int axis, double pos) {
  const double s = axis == 0 ? b.w : (axis == 1 ? c.x : c.y);
  return s < pos * 3.0;
}

// round k: how many primitives of every undecided node satisfy the predicate on axis k
__global__ void __launch_bounds__(256)
    ref_count_kernel(const BNodeD *__restrict__ pool, const uint32_t *__restrict__ node_of,
                     const uint32_t *__restrict__ idx, const D4 *__restrict__ B, const D2 *__restrict__ C,
                     uint32_t n, uint32_t level, int axis, const double *__restrict__ cut3,
                     const uint32_t *__restrict__ state, uint32_t *__restrict__ cnt) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t slot = kInactive;
  bool t = false;
  if (p < n) {
    const BNodeD nd = pool[node_of[p]];
    if (nd.split_bin == level && nd.pad != kInactive && !(state[nd.pad] >> 31)) {
      slot = nd.pad;
      const uint32_t s = idx[p];
      t = ref_pred(B[s], C[s], axis, cut3[(size_t)slot * 3 + axis]);
    }
  }
  const unsigned same = __match_any_sync(0xFFFFFFFFu, slot);
  if (same == 0xFFFFFFFFu) {
    const unsigned m = __ballot_sync(0xFFFFFFFFu, t);
    if (slot != kInactive && (threadIdx.x & 31) == 0 && m) atomicAdd(cnt + slot, (uint32_t)__popc(m));
  } else if (t) {
    atomicAdd(cnt + slot, 1u);
  }
}

__global__ void ref_decide_kernel(BNodeD *pool, const uint32_t *__restrict__ active, uint32_t n_active, int axis,
                                  uint32_t *__restrict__ cnt, uint32_t *__restrict__ state) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_active || (state[a] >> 31)) return;
  BNodeD *nd = pool + active[a];
  const uint32_t n = nd->r - nd->l, c = cnt[a];
  if (c > 0 && c < n) {
    state[a] = 0xC0000000u | (uint32_t)axis;  // decided, partition on this axis
    nd->axis = (uint32_t)axis;
    nd->nleft = c;
  } else if (axis == 2) {
    state[a] = 0x80000000u | 2u;  // all three attempts failed: object median, order untouched, label = last axis
    nd->axis = 2u;
    nd->nleft = n >> 1;
  } else {
    cnt[a] = 0;
  }
}

// flags of the elements std::partition will move: falses in the left part, trues in the right part
__global__ void __launch_bounds__(256)
    ref_flags_kernel(const BNodeD *__restrict__ pool, const uint32_t *__restrict__ node_of,
                     const uint32_t *__restrict__ idx, const D4 *__restrict__ B, const D2 *__restrict__ C,
                     uint32_t n, uint32_t level, const double *__restrict__ cut3, const uint32_t *__restrict__ state,
                     uint32_t *__restrict__ mf, uint32_t *__restrict__ mt) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n) return;
  uint32_t f = 0, t = 0;
  if (p < n) {
    const BNodeD nd = pool[node_of[p]];
    if (nd.split_bin == level && nd.pad != kInactive && (state[nd.pad] & 0x40000000u)) {
      const int axis = (int)(state[nd.pad] & 3u);
      const uint32_t s = idx[p];
      const bool pr = ref_pred(B[s], C[s], axis, cut3[(size_t)nd.pad * 3 + axis]);
      const bool left_part = p < nd.l + nd.nleft;
      f = (left_part && !pr) ? 1u : 0u;
      t = (!left_part && pr) ? 1u : 0u;
    }
  }
  mf[p] = f;  // entry n stays 0: the scans then hold totals at index n
  mt[p] = t;
}

__global__ void ref_compact_kernel(const uint32_t *__restrict__ mf, const uint32_t *__restrict__ mt,
                                   const uint32_t *__restrict__ smf, const uint32_t *__restrict__ smt, uint32_t n,
                                   uint32_t *__restrict__ mf_list, uint32_t *__restrict__ mt_list) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  if (mf[p]) mf_list[smf[p]] = p;
  if (mt[p]) mt_list[smt[p]] = p;
}

// the k-th misplaced false from the left trades places with the k-th misplaced true from the right
__global__ void __launch_bounds__(256)
    ref_permute_kernel(const BNodeD *__restrict__ pool, const uint32_t *__restrict__ node_of,
                       const uint32_t *__restrict__ idx, const uint32_t *__restrict__ mf,
                       const uint32_t *__restrict__ mt, const uint32_t *__restrict__ smf,
                       const uint32_t *__restrict__ smt, const uint32_t *__restrict__ mf_list,
                       const uint32_t *__restrict__ mt_list, uint32_t n, uint32_t *__restrict__ out) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint32_t v = idx[p];
  if (mf[p] || mt[p]) {
    const BNodeD nd = pool[node_of[p]];
    const uint32_t mid = nd.l + nd.nleft;
    const uint32_t m = smf[mid] - smf[nd.l];  // misplaced pairs of this node
    if (mf[p]) {
      const uint32_t k = smf[p] - smf[nd.l];
      v = idx[mt_list[smt[mid] + (m - 1u - k)]];
    } else {
      const uint32_t k_from_right = (m - 1u) - (smt[p] - smt[mid]);
      v = idx[mf_list[smf[nd.l] + k_from_right]];
    }
  }
  out[p] = v;
}

__global__ void ref_children_kernel(BNodeD *pool, RefCounters *ctr, const uint32_t *__restrict__ active,
                                    uint32_t n_active, int cur, uint32_t level, uint32_t min_leaf, uint32_t max_depth,
                                    uint32_t *__restrict__ fresh_next, uint32_t *__restrict__ active_next) {
  const uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_active) return;
  BNodeD *nd = pool + active[a];
  const uint32_t left = atomicAdd(&ctr->pool, 2u);
  const uint32_t fs = atomicAdd(&ctr->n_fresh[cur ^ 1], 2u);
  nd->left = left;
  for (int side = 0; side < 2; side++) {
    BNodeD c;
    for (int k = 0; k < 3; k++) c.bmin[k] = c.bmax[k] = 0.0;
    c.l = side ? nd->l + nd->nleft : nd->l;
    c.r = side ? nd->r : nd->l + nd->nleft;
    c.left = kInactive;
    c.depth = nd->depth + 1;
    c.rturns = nd->rturns + (uint32_t)side;
    c.axis = 0;
    c.split_bin = level + 1;
    c.nleft = 0;
    c.slot = fs + side;
    c.pad = kInactive;
    c.pad2[0] = c.pad2[1] = 0;
    fresh_next[fs + side] = left + side;
    const uint32_t cn = c.r - c.l;
    if (!(cn <= min_leaf || c.depth >= max_depth)) {
      const uint32_t as = atomicAdd(&ctr->n_active[cur ^ 1], 1u);
      active_next[as] = left + side;
      c.pad = as;
    }
    pool[left + side] = c;
  }
}

__global__ void ref_nodeof_kernel(const BNodeD *__restrict__ pool, uint32_t *__restrict__ node_of, uint32_t n,
                                  uint32_t level) {
  const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const uint32_t nid = node_of[p];
  const BNodeD nd = pool[nid];
  if (nd.split_bin == level && nd.pad != kInactive) node_of[p] = nd.left + (p < nd.l + nd.nleft ? 0u : 1u);
}

__global__ void ref_reset_kernel(RefCounters *ctr, int which) {
  ctr->n_fresh[which] = 0;
  ctr->n_active[which] = 0;
}

// ---- emission
__global__ void ref_mark_kernel(const BNodeD *__restrict__ pool, uint32_t n_nodes, uint32_t *__restrict__ leaf_start,
                                uint32_t *stats /* [0] max depth, [1] leaves */) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const BNodeD nd = pool[i];
  atomicMax(stats + 0, nd.depth);
  if (nd.left == kInactive) {
    leaf_start[nd.l] = 1u;
    atomicAdd(stats + 1, 1u);
  }
}

// deferred sub-tree roots of the C++11 build: branch nodes at depth == shallow_depth (nanort.h:1656-1670)
__global__ void ref_collect_deferred_kernel(const BNodeD *__restrict__ pool, uint32_t n_nodes, uint32_t shallow,
                                            const uint32_t *__restrict__ leaves_before, uint2 *__restrict__ table,
                                            uint32_t *count, uint32_t cap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const BNodeD nd = pool[i];
  if (nd.depth != shallow || nd.left == kInactive) return;
  const uint32_t pre = 2u * leaves_before[nd.l] - nd.rturns + nd.depth;
  const uint32_t size = 2u * (leaves_before[nd.r] - leaves_before[nd.l]) - 1u;
  const uint32_t k = atomicAdd(count, 1u);
  if (k < cap) table[k] = make_uint2(pre, size);
}

// serial pre-order index -> index in the array the C++11 parallel build leaves behind.
// tab[j] = (pre of the j-th deferred root, nodes appended before its sub-array), sorted by pre.
__device__ __forceinline__ uint32_t ref_map_index(uint32_t pre, uint32_t depth, uint32_t shallow, const uint4 *tab,
                                                  uint32_t n_tab, uint32_t n_shallow) {
  if (n_tab == 0) return pre;
  // j = number of deferred roots with pre(R) < pre  (for a deep node: its own root is the last of them or equal)
  uint32_t lo = 0, hi = n_tab;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (tab[mid].x < pre)
      lo = mid + 1;
    else
      hi = mid;
  }
  if (depth <= shallow) {
    // shallow node: drop the deep nodes of the deferred sub-trees that precede it
    const uint32_t skipped = lo == 0 ? 0u : tab[lo - 1].z;  // sum of (size - 1) over roots before it
    return pre - skipped;
  }
  const uint32_t j = lo - 1;  // deep node: root j is the last root with pre(R) < pre
  return n_shallow + tab[j].y + (pre - tab[j].x - 1u);
}

__global__ void ref_emit_kernel(const BNodeD *__restrict__ pool, uint32_t n_nodes,
                                const uint32_t *__restrict__ leaves_before, uint32_t shallow, const uint4 *__restrict__ tab,
                                uint32_t n_tab, uint32_t n_shallow, Node64 *__restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_nodes) return;
  const BNodeD nd = pool[i];
  const uint32_t lb = leaves_before[nd.l];
  const uint32_t pre = 2u * lb - nd.rturns + nd.depth;
  Node64 o;
  for (int k = 0; k < 3; k++) {
    o.bmin[k] = nd.bmin[k];
    o.bmax[k] = nd.bmax[k];
  }
  if (nd.left == kInactive) {
    o.flag = 1;
    o.axis = 0;  // the reference leaves leaf.axis uninitialised
    o.data[0] = nd.r - nd.l;
    o.data[1] = nd.l;
  } else {
    o.flag = 0;
    o.axis = (int32_t)nd.axis;
    const uint32_t mid = nd.l + nd.nleft;
    o.data[0] = ref_map_index(pre + 1u, nd.depth + 1u, shallow, tab, n_tab, n_shallow);
    o.data[1] = ref_map_index(pre + 2u * (leaves_before[mid] - lb), nd.depth + 1u, shallow, tab, n_tab, n_shallow);
  }
  out[ref_map_index(pre, nd.depth, shallow, tab, n_tab, n_shallow)] = o;
}

__global__ void ref_count_shallow_kernel(const BNodeD *__restrict__ pool, uint32_t n_nodes, uint32_t shallow,
                                         uint32_t *count) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_nodes && pool[i].depth <= shallow) atomicAdd(count, 1u);
}

}  // namespace

#define RB_CUDA(expr)                                          \
  do {                                                         \
    cudaError_t _e = (expr);                                   \
    if (_e != cudaSuccess) {                                   \
      rc = cuda_fail(_e, #expr, __FILE__, __LINE__);           \
      goto done;                                               \
    }                                                          \
  } while (0)
#define RB_CHECK(expr)           \
  do {                           \
    rc = (expr);                 \
    if (rc != NRT_OK) goto done; \
  } while (0)

// cpp11_order: emit the node order of the C++11 parallel build when n > min_primitives_for_parallel_build.
// Inputs: packed double vertices + faces on the device.  Outputs (device arrays owned by the caller on success):
// *d_nodes_out (BVHNode<double>[*n_nodes_out], 2n+2 capacity), *d_indices_out (uint32[n]).
int build_reference_tree_f64_on_device(const double *d_verts, const uint32_t *d_faces, uint32_t n, uint32_t bin_size,
                                       uint32_t min_leaf_primitives, uint32_t max_tree_depth, uint32_t shallow_depth,
                                       uint32_t min_primitives_for_parallel_build, bool cpp11_order, void **d_nodes_out,
                                       uint32_t **d_indices_out, size_t *n_nodes_out, BuildStats16 *stats_out,
                                       double root_bmin[3], double root_bmax[3], cudaStream_t s) {
  struct {
    uint32_t bin_size, min_leaf_primitives, max_tree_depth, shallow_depth, min_primitives_for_parallel_build;
  } opt = {bin_size, min_leaf_primitives, max_tree_depth, shallow_depth, min_primitives_for_parallel_build};
  const int nbins = (int)opt.bin_size;
  const uint32_t min_leaf = opt.min_leaf_primitives < 1 ? 1u : opt.min_leaf_primitives;
  if (nbins > 256) {
    set_error("nrt_build: bin_size > 256 is not supported by the device builders");
    return NRT_ERR_INVALID;
  }
  const bool joined = cpp11_order && n > opt.min_primitives_for_parallel_build;
  if (joined && opt.shallow_depth > 12) {
    set_error("nrt_build: shallow_depth > 12 is not supported by the reference-order emission");
    return NRT_ERR_INVALID;
  }
  int rc = NRT_OK;
  D4 *dA = nullptr, *dB = nullptr;
  D2 *dC = nullptr;
  Node64 *d_nodes = nullptr;
  uint32_t *d_indices = nullptr;
  double *d_cost = nullptr;
  uint32_t *d_idx[2] = {nullptr, nullptr}, *d_nodeof = nullptr, *d_fresh[2] = {nullptr, nullptr},
           *d_active[2] = {nullptr, nullptr};
  key64 *d_keys = nullptr, *d_bins = nullptr;
  uint32_t *d_cnt = nullptr, *d_state = nullptr, *d_mf = nullptr, *d_mt = nullptr,
           *d_smf = nullptr, *d_smt = nullptr, *d_mfl = nullptr, *d_mtl = nullptr, *d_scratch = nullptr, *d_small = nullptr;
  double *d_cut = nullptr;
  BNodeD *d_pool = nullptr;
  RefCounters *d_ctr = nullptr, hc;
  uint2 *d_tab2 = nullptr;
  uint4 *d_tab4 = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  const size_t max_nodes = 2 * (size_t)n + 2;
  const size_t max_active = (size_t)n / ((size_t)min_leaf + 1) + 2;
  const uint32_t grid_n = (n + 255) / 256, grid_n1 = (n + 1 + 255) / 256;
  int cur = 0, which = 0;
  uint32_t level = 0, n_nodes = 0, n_tab = 0, n_shallow = 0;
  uint32_t hstats[2] = {0, 0};
  std::vector<uint2> tab2;
  std::vector<uint4> tab4;

  RB_CUDA(cudaEventCreate(&ev0));
  RB_CUDA(cudaEventCreate(&ev1));
  RB_CUDA(cudaMalloc(&dA, sizeof(D4) * (size_t)n));
  RB_CUDA(cudaMalloc(&dB, sizeof(D4) * (size_t)n));
  RB_CUDA(cudaMalloc(&dC, sizeof(D2) * (size_t)n));
  for (int i = 0; i < 2; i++) {
    RB_CUDA(cudaMalloc(&d_idx[i], sizeof(uint32_t) * (size_t)n));
    RB_CUDA(cudaMalloc(&d_fresh[i], sizeof(uint32_t) * (2 * max_active + 2)));
    RB_CUDA(cudaMalloc(&d_active[i], sizeof(uint32_t) * max_active));
  }
  RB_CUDA(cudaMalloc(&d_nodeof, sizeof(uint32_t) * (size_t)n));
  RB_CUDA(cudaMalloc(&d_keys, sizeof(key64) * 6 * (2 * max_active + 2)));
  RB_CUDA(cudaMalloc(&d_bins, sizeof(key64) * max_active * (size_t)nbins * kBinWords));
  RB_CUDA(cudaMalloc(&d_cost, sizeof(double) * max_active * (size_t)nbins));
  RB_CUDA(cudaMalloc(&d_cut, sizeof(double) * 3 * max_active));
  RB_CUDA(cudaMalloc(&d_cnt, sizeof(uint32_t) * max_active));
  RB_CUDA(cudaMalloc(&d_state, sizeof(uint32_t) * max_active));
  RB_CUDA(cudaMalloc(&d_mf, sizeof(uint32_t) * ((size_t)n + 1)));
  RB_CUDA(cudaMalloc(&d_mt, sizeof(uint32_t) * ((size_t)n + 1)));
  RB_CUDA(cudaMalloc(&d_smf, sizeof(uint32_t) * ((size_t)n + 1)));
  RB_CUDA(cudaMalloc(&d_smt, sizeof(uint32_t) * ((size_t)n + 1)));
  RB_CUDA(cudaMalloc(&d_mfl, sizeof(uint32_t) * ((size_t)n + 1)));
  RB_CUDA(cudaMalloc(&d_mtl, sizeof(uint32_t) * ((size_t)n + 1)));
  RB_CUDA(cudaMalloc(&d_scratch, sizeof(uint32_t) * scan_scratch_words(n + 1)));
  RB_CUDA(cudaMalloc(&d_small, sizeof(uint32_t) * 8));
  RB_CUDA(cudaMalloc(&d_pool, sizeof(BNodeD) * max_nodes));
  RB_CUDA(cudaMalloc(&d_ctr, sizeof(RefCounters)));
  RB_CUDA(cudaMemsetAsync(d_ctr, 0, sizeof(RefCounters), s));  // incl. the padding the host reads back
  RB_CUDA(cudaMalloc(&d_nodes, sizeof(Node64) * max_nodes));
  RB_CUDA(cudaMalloc(&d_indices, sizeof(uint32_t) * (size_t)n));

  RB_CUDA(cudaEventRecord(ev0, s));
  ref_prim_kernel<<<grid_n, 256, 0, s>>>(d_verts, d_faces, n, dA, dB, dC);
  ref_iota_kernel<<<grid_n, 256, 0, s>>>(d_idx[0], d_nodeof, n);
  ref_init_kernel<<<1, 1, 0, s>>>(d_pool, d_ctr, n, min_leaf, opt.max_tree_depth, d_fresh[0], d_active[0]);
  RB_CUDA(cudaGetLastError());

  for (;;) {
    RB_CUDA(cudaMemcpyAsync(&hc, d_ctr, sizeof(hc), cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaStreamSynchronize(s));
    const uint32_t n_fresh = hc.n_fresh[cur], n_active = hc.n_active[cur];
    if (n_fresh == 0) break;
    // ---- boxes of the nodes created by the previous level
    ref_keys_init_kernel<<<(n_fresh * 6 + 255) / 256, 256, 0, s>>>(d_keys, n_fresh);
    ref_bbox_kernel<<<grid_n, 256, 0, s>>>(d_pool, d_nodeof, d_idx[which], dA, dB, n, level, d_keys);
    ref_bbox_store_kernel<<<(n_fresh + 255) / 256, 256, 0, s>>>(d_pool, d_fresh[cur], n_fresh, d_keys);
    ref_reset_kernel<<<1, 1, 0, s>>>(d_ctr, cur ^ 1);
    RB_CUDA(cudaGetLastError());
    if (n_active > 0) {
      // ---- candidate planes
      const size_t words = (size_t)n_active * nbins * kBinWords;
      ref_bins_clear_kernel<<<(unsigned)((words + 255) / 256), 256, 0, s>>>(d_bins, words);
      ref_bins_kernel<<<grid_n, 256, 0, s>>>(d_pool, d_nodeof, d_idx[which], dA, dB, n, level, nbins, d_bins);
      ref_cut_kernel<<<(n_active + 63) / 64, 64, 0, s>>>(d_pool, d_active[cur], n_active, d_bins, nbins, d_cost, d_cut,
                                                          d_cnt, d_state);
      // ---- up to three partition attempts (nanort.h:1827-1857)
      for (int axis = 0; axis < 3; axis++) {
        ref_count_kernel<<<grid_n, 256, 0, s>>>(d_pool, d_nodeof, d_idx[which], dB, dC, n, level, axis, d_cut,
                                                d_state, d_cnt);
        ref_decide_kernel<<<(n_active + 255) / 256, 256, 0, s>>>(d_pool, d_active[cur], n_active, axis, d_cnt, d_state);
      }
      RB_CUDA(cudaGetLastError());
      // ---- std::partition's element order
      ref_flags_kernel<<<grid_n1, 256, 0, s>>>(d_pool, d_nodeof, d_idx[which], dB, dC, n, level, d_cut, d_state, d_mf,
                                               d_mt);
      RB_CHECK(exclusive_scan_u32_async(d_mf, d_smf, n + 1, d_scratch, s));
      RB_CHECK(exclusive_scan_u32_async(d_mt, d_smt, n + 1, d_scratch, s));
      ref_compact_kernel<<<grid_n, 256, 0, s>>>(d_mf, d_mt, d_smf, d_smt, n, d_mfl, d_mtl);
      ref_permute_kernel<<<grid_n, 256, 0, s>>>(d_pool, d_nodeof, d_idx[which], d_mf, d_mt, d_smf, d_smt, d_mfl, d_mtl,
                                                n, d_idx[which ^ 1]);
      which ^= 1;
      // ---- children
      ref_children_kernel<<<(n_active + 255) / 256, 256, 0, s>>>(d_pool, d_ctr, d_active[cur], n_active, cur, level,
                                                                min_leaf, opt.max_tree_depth, d_fresh[cur ^ 1],
                                                                d_active[cur ^ 1]);
      ref_nodeof_kernel<<<grid_n, 256, 0, s>>>(d_pool, d_nodeof, n, level);
      RB_CUDA(cudaGetLastError());
    }
    cur ^= 1;
    level++;
    if (level > opt.max_tree_depth + 2u) {
      set_error("reference-exact build: level loop did not terminate");
      rc = NRT_ERR_INVALID;
      goto done;
    }
  }
  n_nodes = hc.pool;

  // ---- emission
  RB_CUDA(cudaMemsetAsync(d_mf, 0, sizeof(uint32_t) * ((size_t)n + 1), s));
  RB_CUDA(cudaMemsetAsync(d_small, 0, sizeof(uint32_t) * 8, s));
  ref_mark_kernel<<<(n_nodes + 255) / 256, 256, 0, s>>>(d_pool, n_nodes, d_mf, d_small);
  RB_CHECK(exclusive_scan_u32_async(d_mf, d_smf, n + 1, d_scratch, s));  // leaves starting before a position
  if (joined) {
    const uint32_t cap = 1u << opt.shallow_depth;
    RB_CUDA(cudaMalloc(&d_tab2, sizeof(uint2) * cap));
    RB_CUDA(cudaMalloc(&d_tab4, sizeof(uint4) * cap));
    ref_collect_deferred_kernel<<<(n_nodes + 255) / 256, 256, 0, s>>>(d_pool, n_nodes, opt.shallow_depth, d_smf, d_tab2,
                                                                      d_small + 2, cap);
    ref_count_shallow_kernel<<<(n_nodes + 255) / 256, 256, 0, s>>>(d_pool, n_nodes, opt.shallow_depth, d_small + 3);
    uint32_t hs[4];
    RB_CUDA(cudaMemcpyAsync(hs, d_small, sizeof(hs), cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaStreamSynchronize(s));
    n_tab = std::min(hs[2], cap);
    n_shallow = hs[3];
    tab2.resize(n_tab);
    if (n_tab) RB_CUDA(cudaMemcpy(tab2.data(), d_tab2, sizeof(uint2) * n_tab, cudaMemcpyDeviceToHost));
    std::sort(tab2.begin(), tab2.end(), [](const uint2 &x, const uint2 &y) { return x.x < y.x; });
    tab4.resize(n_tab);
    uint32_t before = 0;  // nodes appended before root j's sub-array = sum over i<j of (size_i - 1)
    for (uint32_t j = 0; j < n_tab; j++) {
      tab4[j].x = tab2[j].x;
      tab4[j].y = before;
      before += tab2[j].y - 1u;
      tab4[j].z = before;  // deep nodes of roots 0..j: what a later shallow node has to skip
      tab4[j].w = 0;
    }
    if (n_tab) RB_CUDA(cudaMemcpyAsync(d_tab4, tab4.data(), sizeof(uint4) * n_tab, cudaMemcpyHostToDevice, s));
  }
  ref_emit_kernel<<<(n_nodes + 255) / 256, 256, 0, s>>>(d_pool, n_nodes, d_smf, opt.shallow_depth, d_tab4, n_tab,
                                                        n_shallow, d_nodes);
  RB_CUDA(cudaGetLastError());
  RB_CUDA(cudaMemcpyAsync(d_indices, d_idx[which], sizeof(uint32_t) * (size_t)n, cudaMemcpyDeviceToDevice, s));
  RB_CUDA(cudaEventRecord(ev1, s));
  RB_CUDA(cudaMemcpyAsync(hstats, d_small, sizeof(hstats), cudaMemcpyDeviceToHost, s));
  {
    BNodeD root;
    RB_CUDA(cudaMemcpyAsync(&root, d_pool, sizeof(BNodeD), cudaMemcpyDeviceToHost, s));
    RB_CUDA(cudaStreamSynchronize(s));
    for (int k = 0; k < 3; k++) {
      root_bmin[k] = root.bmin[k];
      root_bmax[k] = root.bmax[k];
    }
    float ms = 0.0f;
    RB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    *n_nodes_out = n_nodes;
    stats_out->max_tree_depth = hstats[0];
    stats_out->num_leaf_nodes = hstats[1];
    stats_out->num_branch_nodes = n_nodes - hstats[1];
    stats_out->build_secs = ms * 1e-3f;
    *d_nodes_out = d_nodes;
    *d_indices_out = d_indices;
    d_nodes = nullptr;
    d_indices = nullptr;
  }

done:
  cudaFree(d_nodes);  // only on failure: success hands them to the caller
  cudaFree(d_indices);
  cudaFree(d_cost);
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dC);
  for (int i = 0; i < 2; i++) {
    cudaFree(d_idx[i]);
    cudaFree(d_fresh[i]);
    cudaFree(d_active[i]);
  }
  cudaFree(d_nodeof);
  cudaFree(d_keys);
  cudaFree(d_bins);
  cudaFree(d_cut);
  cudaFree(d_cnt);
  cudaFree(d_state);
  cudaFree(d_mf);
  cudaFree(d_mt);
  cudaFree(d_smf);
  cudaFree(d_smt);
  cudaFree(d_mfl);
  cudaFree(d_mtl);
  cudaFree(d_scratch);
  cudaFree(d_small);
  cudaFree(d_pool);
  cudaFree(d_ctr);
  cudaFree(d_tab2);
  cudaFree(d_tab4);
  if (ev0) cudaEventDestroy(ev0);
  if (ev1) cudaEventDestroy(ev1);
  return rc;
}

}  // namespace nrt
