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

__global__ void branch_flags_kernel(const Node40 *__restrict__ nodes, uint32_t n, uint32_t *__restrict__ flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) flags[i] = nodes[i].flag == 0 ? 1u : 0u;
}

__device__ __forceinline__ int child_ref(const Node40 &c, uint32_t cidx, const uint32_t *widx) {
  if (c.flag == 0) return (int)widx[cidx];
  if (c.data[0] == 0) return kEmptyLeaf;
  return ~(int)c.data[1];
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
  WideNode w;
  w.q0 = make_float4(c0.bmin[0], c0.bmin[1], c0.bmin[2], c0.bmax[0]);
  w.q1 = make_float4(c0.bmax[1], c0.bmax[2], c1.bmin[0], c1.bmin[1]);
  w.q2 = make_float4(c1.bmin[2], c1.bmax[0], c1.bmax[1], c1.bmax[2]);
  w.q3 = make_int4(child_ref(c0, nd.data[0], widx), child_ref(c1, nd.data[1], widx), nd.axis, 0);
  wide[widx[i]] = w;
}

int derive_private_layout(Accel *a, cudaStream_t s) {
  const uint32_t n_nodes = (uint32_t)a->n_nodes;
  const uint32_t n_prims = a->n_prims;
  if (a->d_tris) cudaFree(a->d_tris);
  if (a->d_wide) cudaFree(a->d_wide);
  a->d_tris = nullptr;
  a->d_wide = nullptr;
  NRT_CUDA(cudaMalloc(&a->d_tris, sizeof(PackedTri) * (size_t)n_prims));
  pack_tris_kernel<<<(n_prims + 255) / 256, 256, 0, s>>>(a->d_indices, a->d_faces, a->d_verts, n_prims, a->d_tris);
  NRT_CUDA(cudaGetLastError());

  uint32_t *d_flags = nullptr, *d_widx = nullptr;
  NRT_CUDA(cudaMalloc(&d_flags, sizeof(uint32_t) * (size_t)n_nodes));
  NRT_CUDA(cudaMalloc(&d_widx, sizeof(uint32_t) * (size_t)n_nodes));
  branch_flags_kernel<<<(n_nodes + 255) / 256, 256, 0, s>>>(a->d_nodes, n_nodes, d_flags);
  NRT_CUDA(cudaGetLastError());
  uint32_t n_branch = 0;
  int rc = exclusive_scan_u32(d_flags, d_widx, n_nodes, &n_branch, s);
  if (rc != NRT_OK) {
    cudaFree(d_flags);
    cudaFree(d_widx);
    return rc;
  }
  a->n_wide = n_branch > 0 ? n_branch : 1;
  a->root_is_leaf = (n_branch == 0);
  NRT_CUDA(cudaMalloc(&a->d_wide, sizeof(WideNode) * a->n_wide));
  wide_nodes_kernel<<<(n_nodes + 255) / 256, 256, 0, s>>>(a->d_nodes, n_nodes, d_widx, a->d_wide, a->d_tris);
  NRT_CUDA(cudaGetLastError());
  NRT_CUDA(cudaStreamSynchronize(s));
  cudaFree(d_flags);
  cudaFree(d_widx);
  return NRT_OK;
}

}  // namespace nrt
