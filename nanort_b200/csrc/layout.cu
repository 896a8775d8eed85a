// Derives the private traversal layout (64-byte child-pair nodes + 48-byte packed
// triangles in leaf order) from the API-visible nanort arrays that Build or
// nrt_adopt left on the device:
//   nodes   BVHNode<float>[n_nodes]   /root/reference/nanort.h:498-550
//   indices uint32[n_prims]           (BVHAccel::indices_, nanort.h:855)
//   faces / verts as TriangleMesh holds them (nanort.h:925-930)
// The packed triangles remove the reference's two levels of indirection at
// intersection time (indices_ -> faces -> vertices, nanort.h:2394 + 1065-1071).
#include "common.cuh"
#include "scan.cuh"

namespace nrt {

__global__ void pack_tris_kernel(const uint32_t *__restrict__ indices, const uint32_t *__restrict__ faces,
                                 const float *__restrict__ verts, uint32_t n, PackedTri *__restrict__ out) {
  uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  uint32_t prim = indices[slot];
  uint32_t f0 = faces[3 * (size_t)prim + 0], f1 = faces[3 * (size_t)prim + 1], f2 = faces[3 * (size_t)prim + 2];
  const float *p0 = verts + 3 * (size_t)f0, *p1 = verts + 3 * (size_t)f1, *p2 = verts + 3 * (size_t)f2;
  PackedTri t;
  t.a = make_float4(p0[0], p0[1], p0[2], __uint_as_float(prim));
  t.b = make_float4(p1[0], p1[1], p1[2], __uint_as_float(0u));
  t.c = make_float4(p2[0], p2[1], p2[2], 0.0f);
  out[slot] = t;
}

// Box primitives (top-level tree of a two-level scene): the same 48-byte slot carries the instance's world box,
//   bmin.xyz, instance id | bmax.xyz, last_in_leaf flag | unused
__global__ void pack_boxes_kernel(const uint32_t *__restrict__ indices, const float *__restrict__ boxes6, uint32_t n,
                                  PackedTri *__restrict__ out) {
  uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n) return;
  const uint32_t prim = indices[slot];
  const float *b = boxes6 + 6 * (size_t)prim;
  PackedTri t;
  t.a = make_float4(b[0], b[1], b[2], __uint_as_float(prim));
  t.b = make_float4(b[3], b[4], b[5], __uint_as_float(0u));
  t.c = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  out[slot] = t;
}

__global__ void branch_flags_kernel(const Node40 *__restrict__ nodes, uint32_t n, uint32_t *__restrict__ flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = nodes[i].flag == 0 ? 1u : 0u;
}

__device__ __forceinline__ int child_ref(const Node40 &c, uint32_t cidx, const uint32_t *widx) {
  if (c.flag == 0) return (int)widx[cidx];
  if (c.data[0] == 0) return kEmptyLeaf;
  return ~(int)c.data[1];
}

__device__ __forceinline__ void invert_box(Node40 &c) {
  for (int k = 0; k < 3; k++) {
    c.bmin[k] = 3.402823466e38f;
    c.bmax[k] = -3.402823466e38f;
  }
}

__global__ void wide_nodes_kernel(const Node40 *__restrict__ nodes, uint32_t n, const uint32_t *__restrict__ widx,
                                  WideNode *__restrict__ wide, PackedTri *__restrict__ tris) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Node40 nd = nodes[i];
  if (nd.flag != 0) {
    // leaf: mark its last triangle so that the traversal needs no count
    if (nd.data[0] > 0) {
      float *w = reinterpret_cast<float *>(&tris[(size_t)nd.data[1] + nd.data[0] - 1].b) + 3;
      *w = __uint_as_float(1u);
    }
    if (i == 0) {
      // the whole tree is one leaf: a pair whose second child is empty
      WideNode w;
      w.q0 = make_float4(nd.bmin[0], nd.bmin[1], nd.bmin[2], nd.bmax[0]);
      // the empty second child carries an inverted box: it can never pass the slab test
      w.q1 = make_float4(nd.bmax[1], nd.bmax[2], 3.402823466e38f, 3.402823466e38f);
      w.q2 = make_float4(3.402823466e38f, -3.402823466e38f, -3.402823466e38f, -3.402823466e38f);
      w.q3 = make_int4(nd.data[0] ? ~(int)nd.data[1] : kEmptyLeaf, kEmptyLeaf, 0, 0);
      wide[0] = w;
    }
    return;
  }
  Node40 c0 = nodes[nd.data[0]], c1 = nodes[nd.data[1]];
  // a child leaf without primitives (reference trees built with min_leaf_primitives == 0 contain them) carries an
  // inverted box: it can never pass the slab test, so no kernel ever has to follow an empty reference
  if (c0.flag != 0 && c0.data[0] == 0) invert_box(c0);
  if (c1.flag != 0 && c1.data[0] == 0) invert_box(c1);
  WideNode w;
  w.q0 = make_float4(c0.bmin[0], c0.bmin[1], c0.bmin[2], c0.bmax[0]);
  w.q1 = make_float4(c0.bmax[1], c0.bmax[2], c1.bmin[0], c1.bmin[1]);
  w.q2 = make_float4(c1.bmin[2], c1.bmax[0], c1.bmax[1], c1.bmax[2]);
  w.q3 = make_int4(child_ref(c0, nd.data[0], widx), child_ref(c1, nd.data[1], widx), nd.axis, 0);
  wide[widx[i]] = w;
}

// ---- top treelet: the first kTopNodes WideNodes in breadth-first order go to the front of the array, so that
// the traversal kernel can stage them into shared memory with a single contiguous TMA bulk copy and address
// them with the same index (ref < n_top -> shared memory copy).  The rest keeps its depth-first order.
constexpr uint32_t kTopNodes = 256;

__global__ void bfs_top_kernel(const WideNode *__restrict__ wide, uint32_t n_wide, uint32_t *__restrict__ new_idx,
                               uint32_t *__restrict__ not_top, uint32_t *n_top_out) {
  __shared__ uint32_t queue[kTopNodes];
  if (threadIdx.x != 0) return;
  uint32_t head = 0, tail = 1;
  queue[0] = 0;
  while (head < tail) {
    const uint32_t i = queue[head];
    new_idx[i] = head;
    not_top[i] = 0u;
    head++;
    const int4 q3 = wide[i].q3;
    if (q3.x >= 0 && tail < kTopNodes && (uint32_t)q3.x < n_wide) queue[tail++] = (uint32_t)q3.x;
    if (q3.y >= 0 && tail < kTopNodes && (uint32_t)q3.y < n_wide) queue[tail++] = (uint32_t)q3.y;
  }
  *n_top_out = tail;
}

__global__ void fill_u32_kernel(uint32_t *p, uint32_t n, uint32_t v) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void finish_new_idx_kernel(const uint32_t *__restrict__ not_top, const uint32_t *__restrict__ rank,
                                      const uint32_t *n_top, uint32_t n, uint32_t *__restrict__ new_idx) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && not_top[i]) new_idx[i] = *n_top + rank[i];
}

__global__ void remap_wide_kernel(const WideNode *__restrict__ in, const uint32_t *__restrict__ new_idx, uint32_t n,
                                  WideNode *__restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  WideNode w = in[i];
  if (w.q3.x >= 0) w.q3.x = (int)new_idx[w.q3.x];
  if (w.q3.y >= 0) w.q3.y = (int)new_idx[w.q3.y];
  out[new_idx[i]] = w;
}

// ---- round-2 layout: PairNode (sign-addressed planes) and TriCM (component-major triangles), derived from the
// final WideNode / PackedTri arrays so that indices, refs and slots are shared by every kernel
__global__ void pair_from_wide_kernel(const WideNode *__restrict__ wide, uint32_t n, PairNode *__restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const WideNode w = wide[i];
  // WideNode: q0 = c0.lo.xyz, c0.hi.x | q1 = c0.hi.yz, c1.lo.xy | q2 = c1.lo.z, c1.hi.xyz
  const float lo0x = w.q0.x, lo0y = w.q0.y, lo0z = w.q0.z, hi0x = w.q0.w, hi0y = w.q1.x, hi0z = w.q1.y;
  const float lo1x = w.q1.z, lo1y = w.q1.w, lo1z = w.q2.x, hi1x = w.q2.y, hi1y = w.q2.z, hi1z = w.q2.w;
  PairNode p;
  p.x[0] = make_float4(lo0x, lo1x, hi0x, hi1x);
  p.x[1] = make_float4(hi0x, hi1x, lo0x, lo1x);
  p.y[0] = make_float4(lo0y, lo1y, hi0y, hi1y);
  p.y[1] = make_float4(hi0y, hi1y, lo0y, lo1y);
  p.z[0] = make_float4(lo0z, lo1z, hi0z, hi1z);
  p.z[1] = make_float4(hi0z, hi1z, lo0z, lo1z);
  p.r = w.q3;
  p.pad = make_int4(0, 0, 0, 0);
  out[i] = p;
}

__global__ void tris_cm_kernel(const PackedTri *__restrict__ in, uint32_t n, TriCM *__restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const PackedTri t = in[i];
  const uint32_t w = (__float_as_uint(t.a.w) & 0x7FFFFFFFu) | (__float_as_uint(t.b.w) != 0u ? 0x80000000u : 0u);
  const float wf = __uint_as_float(w);
  TriCM o;
  o.X = make_float4(t.a.x, t.b.x, t.c.x, wf);
  o.Y = make_float4(t.a.y, t.b.y, t.c.y, wf);
  o.Z = make_float4(t.a.z, t.b.z, t.c.z, wf);
  out[i] = o;
}

static int reorder_top_treelet(Accel *a, cudaStream_t s) {
  const uint32_t n = (uint32_t)a->n_wide;
  uint32_t *d_new = nullptr, *d_not = nullptr, *d_rank = nullptr, *d_ntop = nullptr;
  WideNode *d_out = nullptr;
  cudaError_t ea = cudaMalloc(&d_new, sizeof(uint32_t) * (size_t)n);
  if (ea == cudaSuccess) ea = cudaMalloc(&d_not, sizeof(uint32_t) * (size_t)n);
  if (ea == cudaSuccess) ea = cudaMalloc(&d_rank, sizeof(uint32_t) * (size_t)n);
  if (ea == cudaSuccess) ea = cudaMalloc(&d_ntop, sizeof(uint32_t));
  if (ea == cudaSuccess) ea = cudaMalloc(&d_out, sizeof(WideNode) * (size_t)n);
  if (ea == cudaSuccess) {
    fill_u32_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_not, n, 1u);
    bfs_top_kernel<<<1, 32, 0, s>>>(a->d_wide, n, d_new, d_not, d_ntop);
    ea = cudaGetLastError();
  }
  uint32_t n_rest = 0;
  int rc = ea == cudaSuccess ? exclusive_scan_u32(d_not, d_rank, n, &n_rest, s)
                             : cuda_fail(ea, "reorder_top_treelet", __FILE__, __LINE__);
  if (rc == NRT_OK) {
    finish_new_idx_kernel<<<(n + 255) / 256, 256, 0, s>>>(d_not, d_rank, d_ntop, n, d_new);
    remap_wide_kernel<<<(n + 255) / 256, 256, 0, s>>>(a->d_wide, d_new, n, d_out);
    uint32_t h_ntop = 0;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(&h_ntop, d_ntop, sizeof(uint32_t), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
      rc = cuda_fail(e, "reorder_top_treelet", __FILE__, __LINE__);
    } else {
      cudaFree(a->d_wide);
      a->d_wide = d_out;
      d_out = nullptr;
      a->n_top = h_ntop;
    }
  }
  cudaFree(d_new);
  cudaFree(d_not);
  cudaFree(d_rank);
  cudaFree(d_ntop);
  cudaFree(d_out);
  return rc;
}

int derive_private_layout(Accel *a, cudaStream_t s) {
  const uint32_t n_nodes = (uint32_t)a->n_nodes;
  const uint32_t n_prims = a->n_prims;
  cudaFree(a->d_tris);
  cudaFree(a->d_wide);
  cudaFree(a->d_pair);
  cudaFree(a->d_tris_cm);
  a->d_tris = nullptr;
  a->d_wide = nullptr;
  a->d_pair = nullptr;
  a->d_tris_cm = nullptr;
  NRT_CUDA(cudaMalloc(&a->d_tris, sizeof(PackedTri) * (size_t)n_prims));
  if (a->d_prim_boxes)
    pack_boxes_kernel<<<(n_prims + 255) / 256, 256, 0, s>>>(a->d_indices, a->d_prim_boxes, n_prims, a->d_tris);
  else
    pack_tris_kernel<<<(n_prims + 255) / 256, 256, 0, s>>>(a->d_indices, a->d_faces, a->d_verts, n_prims, a->d_tris);
  NRT_CUDA(cudaGetLastError());

  uint32_t *d_flags = nullptr, *d_widx = nullptr;
  uint32_t n_branch = 0;
  int rc = NRT_OK;
  cudaError_t e = cudaMalloc(&d_flags, sizeof(uint32_t) * (size_t)n_nodes);
  if (e == cudaSuccess) e = cudaMalloc(&d_widx, sizeof(uint32_t) * (size_t)n_nodes);
  if (e == cudaSuccess) {
    branch_flags_kernel<<<(n_nodes + 255) / 256, 256, 0, s>>>(a->d_nodes, n_nodes, d_flags);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) rc = exclusive_scan_u32(d_flags, d_widx, n_nodes, &n_branch, s);
  if (e == cudaSuccess && rc == NRT_OK) {
    a->n_wide = n_branch > 0 ? n_branch : 1;
    a->root_is_leaf = (n_branch == 0);
    e = cudaMalloc(&a->d_wide, sizeof(WideNode) * a->n_wide);
  }
  if (e == cudaSuccess && rc == NRT_OK) {
    wide_nodes_kernel<<<(n_nodes + 255) / 256, 256, 0, s>>>(a->d_nodes, n_nodes, d_widx, a->d_wide, a->d_tris);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(d_flags);  // every path, error or not
  cudaFree(d_widx);
  if (e != cudaSuccess) return cuda_fail(e, "derive_private_layout", __FILE__, __LINE__);
  if (rc != NRT_OK) return rc;
  rc = reorder_top_treelet(a, s);
  if (rc != NRT_OK || a->d_prim_boxes) return rc;  // box accels (top level of a scene) are walked by scene.cu only
  NRT_CUDA(cudaMalloc(&a->d_pair, sizeof(PairNode) * a->n_wide));
  NRT_CUDA(cudaMalloc(&a->d_tris_cm, sizeof(TriCM) * (size_t)n_prims));
  pair_from_wide_kernel<<<((uint32_t)a->n_wide + 255) / 256, 256, 0, s>>>(a->d_wide, (uint32_t)a->n_wide, a->d_pair);
  tris_cm_kernel<<<(n_prims + 255) / 256, 256, 0, s>>>(a->d_tris, n_prims, a->d_tris_cm);
  NRT_CUDA(cudaGetLastError());
  NRT_CUDA(cudaStreamSynchronize(s));
  return NRT_OK;
}

}  // namespace nrt
