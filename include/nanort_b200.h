/*
 * nanort_b200.h -- the C-ABI boundary of the B200-native nanort hot path.
 *
 * nanort itself has no binary interface: its "plugin API" is the set of C++
 * templates in nanort.h.  This header is the extern "C"
 * surface that a drop-in nanort.h facade (include/nanort.h in this repository)
 * binds instead of running the templates on the CPU.  Plain pointers and sizes
 * only; every record keeps nanort's byte layout:
 *
 *   ray   36 B  nanort::Ray<float>                  /root/reference/nanort.h:474-496
 *   hit   16 B  nanort::TriangleIntersection<float> nanort.h:996-1005  {u, v, t, prim_id}
 *   node  40 B  nanort::BVHNode<float>              nanort.h:498-550
 *   build options 28 B  nanort::BVHBuildOptions<float>  nanort.h:559-583
 *   build stats   16 B  nanort::BVHBuildStatistics      nanort.h:586-599
 *   trace options 16 B  nanort::BVHTraceOptions         nanort.h:604-624
 *
 * All functions return NRT_OK (0) or a negative error code; nrt_last_error()
 * gives the message of the calling thread's last failure.  There is no CPU
 * fallback: without a usable CUDA device every entry point fails with
 * NRT_ERR_CUDA.
 */
#ifndef NANORT_B200_H_
#define NANORT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NRT_OK 0
#define NRT_ERR_INVALID -1  /* bad argument (null pointer, n_prims == 0 -> Build returns false, nanort.h:1907) */
#define NRT_ERR_CUDA -2     /* CUDA runtime / launch failure, or no device */
#define NRT_ERR_NOMEM -3

/* nrt_traverse* flags */
#define NRT_TRAVERSE_FAST 0u         /* private 64-B child-pair layout, near-child-by-distance order      */
#define NRT_TRAVERSE_CONFORMANCE 1u  /* walk the nanort 40-B node array in the reference's exact order
                                        (nanort.h:2526-2547): identical winners even for exact-t ties   */
#define NRT_TRAVERSE_CPP03_INVERSE 2u /* vsafe_inverse sign convention of the C++03 build (nanort.h:440-462:
                                        -0.0f -> +inf).  Default is the C++11 one (copysign, :418-439)   */

/* nrt_traverse / nrt_traverse_device: the ray array holds 32-byte records {org[3], dir[3], min_t, max_t} -- nanort::Ray
 * (nanort.h:474-496) without its `type` word, which Traverse never reads -- 16-byte aligned.  Opt-in for callers bound by
 * the host link: 32 B up (+ 16 B down with hit_mask == NULL: a miss is prim_id == 0xFFFFFFFF) instead of 36 + 17 per ray.
 * Same hits bit for bit.  Triangle accels only. */
#define NRT_TRAVERSE_RAY32 4u

/* Occlusion queries (opt-in; nanort itself has no any-hit, examples/path_tracer/main.cc:675-701 asks Traverse for the
 * closest hit and looks at the bool): the ray stops at the first primitive it hits inside [min_t, max_t).
 *   nrt_ao_params.flags / nrt_path_params.flags: the AO / shadow launches of the wavefront passes -- the framebuffer
 *     is the same bit for bit (only "occluded or not" reaches it); camera and bounce rays stay closest-hit.
 *   nrt_traverse / nrt_traverse_device: hit_mask (and prim_id != 0xFFFFFFFF) is the same; the RECORD of a hit ray is
 *     that of some primitive on the ray, not necessarily the closest.
 * Fast path only (ignored with NRT_TRAVERSE_CONFORMANCE).  Triangle accels only. */
#define NRT_TRAVERSE_ANY_HIT 8u

/* nrt_ao_params.flags only: run the AO stages as stand-alone kernels instead of fused into the traversal
 * kernel's retire step (A/B and equivalence tests; results are identical) */
#define NRT_AO_UNFUSED 0x10000u
/* nrt_ao_params.flags: d_accum is a tile-major buffer of THIS shard's tiles (tile k of the shard at float offset
 * k * tile_w * tile_h, rows of tile_w floats) instead of the width x height image -- what nrt_render_ao_sharded
 * accumulates into so that the framebuffer all-gather needs no pack step */
#define NRT_AO_PACKED_TILES 0x20000u

typedef struct nrt_accel nrt_accel; /* opaque: device-resident BVH + host mirrors */

const char *nrt_last_error(void);
int nrt_device_count(void);
/* Device used by subsequently created accels of this thread (default 0). */
int nrt_set_device(int device);

/*
 * Replaces BVHAccel<float>::Build<TriangleMesh<float>, TriangleSAHPred<float>>
 * (nanort.h:1892-2149; caller examples/path_tracer/main.cc:752-763).
 * verts/faces are HOST pointers laid out as TriangleMesh takes them
 * (nanort.h:925-930): vertex i at (char*)verts + i*stride_bytes, 3 uint32 per
 * face.  They are copied to the device; unlike the reference the pointers need
 * not outlive the call.  n_verts == 0 derives max(face index)+1.
 * build_opts_28B == NULL uses the defaults of nanort.h:574-582.
 * Returns NRT_ERR_INVALID for n_prims == 0 (reference: Build returns false).
 */
int nrt_build(const float *verts, size_t stride_bytes, size_t n_verts, const uint32_t *faces,
              uint32_t n_prims, const void *build_opts_28B, nrt_accel **out);

/*
 * nrt_build with a choice of builder.  flags:
 *   NRT_BUILD_FAST (0)              the production binned-SAH builder (what nrt_build runs)
 *   NRT_BUILD_REFERENCE_TREE        conformance build: reproduce, on the device, exactly the node array and
 *                                   indices_ CPU nanort writes at the pinned commit (x-only binning of
 *                                   nanort.h:1357, axis retry / median fallback :1827-1857, TriangleSAHPred +
 *                                   libstdc++ std::partition element order, BuildTree / parallel-join node order)
 *   NRT_BUILD_REFERENCE_CPP03_ORDER with the flag above: the serial (C++03 / small scene) node order even above
 *                                   min_primitives_for_parallel_build (default: the C++11 build's joined order)
 */
#define NRT_BUILD_FAST 0u
#define NRT_BUILD_REFERENCE_TREE 1u
#define NRT_BUILD_REFERENCE_CPP03_ORDER 2u
int nrt_build_ex(const float *verts, size_t stride_bytes, size_t n_verts, const uint32_t *faces, uint32_t n_prims,
                 const void *build_opts_28B, uint32_t flags, nrt_accel **out);

/*
 * Conformance entry: adopt an existing nanort-layout tree (e.g. one built by
 * the CPU reference; BVHAccel::GetNodes()/GetIndices(), nanort.h:786-787, or
 * the Dump format nanort.h:2164-2220) so the GPU traverses exactly that tree.
 */
int nrt_adopt(const void *nodes_40B, size_t n_nodes, const uint32_t *indices, size_t n_indices,
              const float *verts, size_t stride_bytes, size_t n_verts, const uint32_t *faces,
              uint32_t n_prims, nrt_accel **out);

void nrt_free(nrt_accel *a);

/* BVHAccel::GetStatistics (nanort.h:725): {max_tree_depth, num_leaf_nodes, num_branch_nodes, build_secs}.
 * build_secs is the device time of the build kernels (the reference never fills it, nanort.h:591). */
int nrt_stats(const nrt_accel *a, void *stats_16B);
/* BVHAccel::BoundingBox (nanort.h:792-804). */
int nrt_bounding_box(const nrt_accel *a, float bmin[3], float bmax[3]);
/* ------------------------------------------------------------------ non-triangle primitives
 * nanort's Prim / Pred / Intersector concept (nanort.h:698-860, 1014-1229) on the device: host functors cannot run in a
 * kernel, so the hook is a set of primitive KINDS, each the device restatement of one of the reference's own models:
 *   NRT_PRIM_SPHERES  examples/particle_primitive/main.cc:80-291 (SpherePred, SphereGeometry, SphereIntersector):
 *                     data = centers (float3, stride_bytes apart), aux = radii.  The accel then works with
 *                     nrt_traverse / nrt_traverse_device / nrt_nodes / nrt_stats / nrt_bounding_box; hit records are
 *                     {u, v, t, prim_id} as SphereIntersector::PostTraversal fills them (NOTE the model's own
 *                     behaviour, kept: no ray.min_t test inside Intersect, only prim_ids_range of the trace options).
 *   NRT_PRIM_BOXES    the node-level primitive of the two-level API (NodeBBoxGeometry / NodeBBoxIntersector,
 *                     examples/nanosg/nanosg.h:447-640): data = {bmin.xyz, bmax.xyz} per box, stride 24, aux NULL;
 *                     query with nrt_list_node_intersections.
 * The tree is built by the production builder over the primitives' boxes (their Prim::BoundingBox). */
#define NRT_PRIM_SPHERES 1u
#define NRT_PRIM_BOXES 2u
int nrt_build_prims(uint32_t kind, const float *data, size_t stride_bytes, const float *aux, uint32_t n_prims,
                    const void *build_opts_28B, nrt_accel **out);
/* BVHAccel::ListNodeIntersections (nanort.h:2607-2692) for n rays over a NRT_PRIM_BOXES accel: per ray the (at most
 * max_intersections <= 64) nearest boxes the ray pierces, nearest first, as records {float t_min, float t_max,
 * uint32 node_id} at hits_12B[ray * max_intersections + k], k < counts[ray].  Host pointers.
 * flags: NRT_TRAVERSE_CPP03_INVERSE. */
int nrt_list_node_intersections(const nrt_accel *a, const void *rays_36B, size_t n_rays, int max_intersections,
                                void *hits_12B, uint32_t *counts, uint32_t flags);

/* BVHAccel::GetNodes / GetIndices (nanort.h:786-787): host mirror in nanort layout, downloaded on
 * first use and owned by the accel. */
int nrt_nodes(nrt_accel *a, const void **nodes_40B, size_t *n_nodes, const uint32_t **indices,
              size_t *n_indices);

/*
 * Replaces a loop of BVHAccel<float>::Traverse<TriangleIntersector<float>,
 * TriangleIntersection<float>> calls (nanort.h:2487-2556; callers
 * examples/path_tracer/main.cc:854 and :696).  HOST buffers; the call copies
 * rays up, traverses and copies results back, pipelined in chunks, and returns
 * when hits/hit_mask are complete.  hit_mask[i] = 1/0 is Traverse's return
 * value.  hits[i] = {u,v,t,prim_id} for hits; for misses the record is
 * {0, 0, ray.max_t, 0xFFFFFFFF} (the reference leaves *isect untouched on a
 * miss, nanort.h:1205-1213 -- the nanort.h facade restores that per ray).
 * trace_opts_16B == NULL -> defaults of nanort.h:617-623.  hit_mask may be NULL.
 */
int nrt_traverse(const nrt_accel *a, const void *rays_36B, size_t n_rays, void *hits_16B,
                 uint8_t *hit_mask, const void *trace_opts_16B, uint32_t flags);

/* Same with DEVICE pointers on `stream` (a cudaStream_t, NULL = default stream); asynchronous. */
int nrt_traverse_device(const nrt_accel *a, const void *d_rays_36B, size_t n_rays, void *d_hits_16B,
                        uint8_t *d_hit_mask, const void *trace_opts_16B, uint32_t flags, void *stream);

/*
 * Counting variant (not timed anywhere): walks the same rays and returns the totals the roofline
 * arithmetic needs: boxes tested (one per 40-B nanort node the reference's
 * Traverse would pop on this tree, nanort.h:2527) and triangles tested (Intersect calls, :2397).
 */
int nrt_traverse_count_device(const nrt_accel *a, const void *d_rays_36B, size_t n_rays,
                              const void *trace_opts_16B, uint32_t flags, uint64_t *boxes_tested,
                              uint64_t *prims_tested, void *stream);

/*
 * Profiling aid (not timed anywhere): the same counting walk, returning how the 32 lanes of the persistent warps
 * spent their steps.  stats16 (host): [0] boxes tested, [1] triangles tested, [2] refill events, [3] lanes refilled,
 * [4] node-phase warp steps, and summed over those steps the lanes [5] testing a child pair, [6] without a ray,
 * [7] whose ray has finished and waits for the retire step, [8] parked on postponed leaves; [9] leaf-phase rounds,
 * [10] lanes entering a round with a leaf, [11] triangle-test warp steps, [12] retire events, [13] lanes retired,
 * [14] outer iterations, [15] reserved.  Explains smsp__thread_inst_executed_per_inst_executed in the ncu captures.
 */
int nrt_traverse_lane_stats_device(const nrt_accel *a, const void *d_rays_36B, size_t n_rays,
                                   const void *trace_opts_16B, uint32_t flags, uint64_t *stats16, void *stream);

/* Measured roofs for benchmark reports (bench.py's roofline block), same process / device / clocks as the traversal:
 * streaming read bandwidth over `bytes` of device memory with 16-byte loads (<= ~64 MB: L2-resident after the warm-up
 * pass -> L2 roof; >= 1 GB -> HBM roof), and the pinned host <-> device copy rate (direction 0 = H2D, 1 = D2H).
 * GB/s, best of `iters`.  Not part of the traversal path. */
int nrt_probe_read_gbs(size_t bytes, int iters, double *gb_per_s);
int nrt_probe_copy_gbs(size_t bytes, int iters, int direction, double *gb_per_s);

/* Pinned host memory for ray / hit buffers handed to nrt_traverse (plain memory works too, slower). */
void *nrt_host_alloc(size_t bytes);
void nrt_host_free(void *p);

/*
 * Device-resident wavefront pass for the headline metric: jittered pinhole primary rays (the camera of
 * examples/path_tracer/main.cc:809-817, 839-849) -> Traverse -> one cosine-hemisphere AO ray per hit
 * (hit point main.cc:860, geometric normal :306-312 flipped to the viewer :878-881, ONB + cosine
 * direction :216-250, closest-hit occlusion query as CheckForOccluder :675-701) -> Traverse ->
 * accumulate visibility.  Pixels are taken from tiles of tile_w x tile_h pixels; tile k belongs to
 * shard (k % n_shards) -- this is how the work is split across GPUs.
 */
typedef struct nrt_ao_params {
  float cam[12];        /* org, right*sx, up*sy, forward (nanort_b200/scenes.py:look_at) */
  uint32_t width, height;
  uint32_t spp;         /* samples per pixel in this call */
  uint32_t sample0;     /* index of the first sample (for progressive calls) */
  uint32_t seed;
  uint32_t tile_w, tile_h;
  uint32_t shard, n_shards;
  float ray_min_t, ray_max_t;
  float ao_min_t, ao_max_t;
  uint32_t flags;       /* NRT_TRAVERSE_* */
} nrt_ao_params;

typedef struct nrt_ao_result {
  uint64_t primary_rays;
  uint64_t ao_rays;     /* == primary hits */
  uint64_t ao_hits;     /* occluded AO rays */
  float traverse_ms;    /* device time spent inside the traversal kernels (CUDA events) */
  float total_ms;       /* device time of the whole pass */
  uint32_t launches;    /* kernels launched by this call */
  uint32_t traverse_launches;
  float primary_traverse_ms; /* traverse_ms split by launch kind: camera-ray launches ... */
  float ao_traverse_ms;      /* ... and AO-ray launches */
} nrt_ao_result;

/* d_accum: DEVICE float[width*height] accumulating sum of visibility (1 = unoccluded, 0.0 for primary
 * misses counted as 1); only this shard's pixels are touched.  Asynchronous w.r.t. the host except for
 * the final read-back of the counters in *res (res may be NULL to skip that sync). */
int nrt_render_ao_device(const nrt_accel *a, const nrt_ao_params *p, float *d_accum, nrt_ao_result *res,
                         void *stream);

/* Workload export for benchmarks and parity tests: runs the same pass and additionally writes the two ray
 * queues as 36-byte nanort::Ray records (DEVICE buffers sized for the shard's slot count): primary ray of
 * slot i at d_primary_rays_36B[i], AO rays appended in queue order.  The exported arrays are what
 * bench.py feeds to nrt_traverse (host-buffer arm) and to the CPU reference, so that all arms trace the
 * very same rays. */
int nrt_ao_workload_device(const nrt_accel *a, const nrt_ao_params *p, float *d_accum, void *d_primary_rays_36B,
                           void *d_ao_rays_36B, uint64_t *n_primary, uint64_t *n_ao, void *stream);

/* ------------------------------------------------------------------ multi-GPU
 * One process (or host thread) per GPU.  Rays shard by image tile: tile t belongs to rank t % world; the BVH is
 * replicated (every rank calls nrt_build on the same arrays: the builder is deterministic); there is no exchange
 * during traversal and ONE collective per frame, the framebuffer all-gather (ncclAllGather over NVLink, in place, on
 * the pass's stream).  NCCL is bound at run time (dlopen), so nothing here is needed for single-GPU use.
 *
 *   rank 0:            nrt_comm_unique_id(id)           -> ship the 128 bytes to the other ranks (pipe, file, MPI ...)
 *   every rank:        nrt_set_device(local_gpu); nrt_build(...); nrt_comm_init(id, rank, world, &comm)
 *   every frame, all:  nrt_render_ao_sharded(accel, comm, &params, d_frame_full, &res, stream)
 *
 * nrt_render_ao_sharded = nrt_render_ao_device over this rank's tiles (params.shard / n_shards are overwritten by
 * the communicator's rank / size) + all-gather + unpack: on return (stream order) d_frame_full, a DEVICE
 * float[width*height] on EVERY rank, holds the whole frame's visibility sums.  *res counts this rank's rays.
 * examples/multi_gpu_ao.cc drives it with fork(), without MPI or torchrun. */
typedef struct nrt_comm nrt_comm;
int nrt_comm_unique_id(void *id_128B);
int nrt_comm_init(const void *id_128B, int rank, int world, nrt_comm **out);
void nrt_comm_free(nrt_comm *c);
int nrt_comm_rank(const nrt_comm *c, int *rank, int *world);
int nrt_render_ao_sharded(const nrt_accel *a, nrt_comm *c, const nrt_ao_params *p, float *d_frame_full,
                          nrt_ao_result *res, void *stream);

/*
 * Device-resident wavefront form of the reference path tracer's pixel -> sample -> bounce loop
 * (examples/path_tracer/main.cc:804-991) with its material model (tinyobj materials: diffuse, specular,
 * transmittance, emission, ior, dissolve; main.cc:884-973):
 *   camera ray :809-817, Russian roulette after bounce 3 with p = 0.2 :828-837, radiance Traverse :839-854,
 *   interpolated face-varying normal :862-875 (geometric normal when none is given) flipped to the viewer
 *   :878-881, Schlick Fresnel :894-900, lobe probabilities rhoS/rhoD/rhoR/rhoE :902-929, glossy reflection
 *   :931-935, diffuse lobe with next-event estimation (MeshLight::sampleDirect :337-392 + the CheckForOccluder
 *   shadow Traverse :675-701) and cosine-weighted continuation :937-957 / :216-250, refraction :958-962,
 *   emission (only when the previous event did no light sampling) :963-971, at most max_bounces.
 * Every bounce is two traversal launches (radiance rays, shadow rays) whose retire steps do the shading.
 * All pointers are DEVICE pointers.
 */
typedef struct nrt_path_params {
  float cam[12];
  uint32_t width, height;
  uint32_t spp, sample0, seed;
  uint32_t tile_w, tile_h, shard, n_shards;
  uint32_t max_bounces;   /* uMaxBounces, main.cc:33 (10) */
  float ray_min_t, ray_max_t; /* 1e-3, 1e30 (main.cc:840-842) */
  uint32_t n_materials, n_emissive;
  const void *d_materials;            /* n_materials x 16 floats: diffuse[3] specular[3] transmittance[3] emission[3]
                                         ior dissolve pad pad */
  const void *d_material_ids;         /* uint32 per face, or NULL = material 0 everywhere (Mesh::material_ids) */
  const void *d_emissive_faces;       /* uint32[n_emissive]: faces whose material emits (MeshLight, main.cc:323-335) */
  const void *d_facevarying_normals;  /* float[9 * n_faces] or NULL (Mesh::facevarying_normals) */
  uint32_t flags;         /* NRT_TRAVERSE_* */
  uint32_t pad;
} nrt_path_params;

typedef struct nrt_path_result {
  uint64_t camera_rays;
  uint64_t radiance_rays; /* all radiance Traverse calls, camera rays included */
  uint64_t shadow_rays;
  float traverse_ms, total_ms;
  uint32_t launches, traverse_launches;
} nrt_path_result;

/* d_accum_rgb: DEVICE float[3*width*height], sum over samples of the path radiance (divide by spp). */
int nrt_render_path_device(const nrt_accel *a, const nrt_path_params *p, float *d_accum_rgb, nrt_path_result *res,
                           void *stream);


/* One bounce of that loop on caller-owned DEVICE queues (what nrt_render_path_device repeats max_bounces times): the
 * n_rays radiance rays {org.xyz,min_t | dir.xyz,max_t} with their path ids (= primary slot of the path under the tile
 * map of *p, which gives pixel and sample) are traversed and shaded as bounce `bounce` (main.cc:856-976); continuation
 * rays go to d_out_*, shadow rays {ray, contribution.rgb | pixel} to d_sh_*, d_weight[path id] = {throughput.rgb,
 * do_emission} is read and updated, emission goes to d_accum_rgb; unless skip_shadow_pass the shadow rays are then
 * traversed and the unoccluded contributions added to d_accum_rgb (main.cc:940-947, 675-701).  Output queues need
 * room for n_rays entries.  Synchronises the stream to return the two counts. */
int nrt_path_bounce_device(const nrt_accel *a, const nrt_path_params *p, uint32_t bounce, uint64_t n_rays,
                           const void *d_org_tmin, const void *d_dir_tmax, const uint32_t *d_path_id, void *d_weight,
                           void *d_out_org_tmin, void *d_out_dir_tmax, uint32_t *d_out_path_id, void *d_sh_org_tmin,
                           void *d_sh_dir_tmax, void *d_sh_contrib_pix, float *d_accum_rgb, uint64_t *n_continue,
                           uint64_t *n_shadow, int skip_shadow_pass, void *stream);


/* ------------------------------------------------------------------ two-level scene (instancing)
 * Replaces the reference's scene-graph example, which is how nanort traverses transformed instances:
 *   nanosg::Scene<float,M>::AddNode / Commit   examples/nanosg/nanosg.h:673-755   -> nrt_scene_commit
 *   nanosg::Scene<float,M>::Traverse           examples/nanosg/nanosg.h:779-875   -> nrt_scene_traverse[_device]
 *   BVHAccel<float>::ListNodeIntersections     nanort.h:2607-2692                 (inside the kernels)
 * An instance is a bottom-level accel (borrowed: it must outlive the scene and live on the same device) plus the
 * node's local transform, float[4][4] in the reference's row-vector convention (p' = p . M, translation in row 3,
 * nanosg.h:214-222).  Many instances may share one accel.  The hit record carries what Scene::Traverse returns
 * short of the mesh-dependent normals: {u, v, t (world distance), prim_id, node_id, P (world hit point)}.
 * Reference behaviours kept as they are: the world ray's min_t / max_t gate only the top-level walk; the local ray
 * is {0, FLT_MAX}; at most the 64 nearest instance boxes are examined per ray; cull_back_face is inert. */
typedef struct nrt_scene nrt_scene;

typedef struct nrt_instance {
  const nrt_accel *accel;
  float xform[16];
} nrt_instance;

typedef struct nrt_scene_hit {
  float u, v, t;
  uint32_t prim_id, node_id;
  float P[3];
} nrt_scene_hit; /* 32 bytes */

/* flags: NRT_BUILD_FAST (good top-level tree) or NRT_BUILD_REFERENCE_TREE [| NRT_BUILD_REFERENCE_CPP03_ORDER]
 * (top-level node array bit-identical to the reference's Commit).  n_instances == 0 fails like Commit does. */
int nrt_scene_commit(const nrt_instance *instances, uint32_t n_instances, uint32_t flags, nrt_scene **out);
void nrt_scene_free(nrt_scene *s);
/* Scene::GetBoundingBox (nanosg.h:761-769) */
int nrt_scene_bounding_box(const nrt_scene *s, float bmin[3], float bmax[3]);
/* host mirror of the top-level tree in nanort's layout (indices = instance ids) */
int nrt_scene_nodes(nrt_scene *s, const void **nodes_40B, size_t *n_nodes, const uint32_t **indices,
                    size_t *n_indices);
/* per-instance derived state as Node::Update computes it (nanosg.h:400-445), 76 floats:
 * xform[16] inverse[16] inverse33[16] inverse_transpose33[16] local_bmin[3] local_bmax[3] world_bmin[3] world_bmax[3] */
int nrt_scene_instance_state(const nrt_scene *s, uint32_t instance, float out76[76]);
/* n x Scene::Traverse with HOST pointers; hits[i] = {0,0,max_t,~0,~0,0,0,0} and hit_mask[i] = 0 on a miss.
 * flags: NRT_TRAVERSE_FAST / NRT_TRAVERSE_CONFORMANCE (the reference's exact list-then-visit order) /
 * NRT_TRAVERSE_CPP03_INVERSE */
int nrt_scene_traverse(const nrt_scene *s, const void *rays_36B, size_t n_rays, void *hits_32B, uint8_t *hit_mask,
                       uint32_t flags);
/* Primary + 1-bounce AO over the scene: nrt_render_ao_device's pass (same nrt_ao_params, same tile / shard slot order,
 * same camera rays and cosine directions, d_accum as there) with Scene::Traverse as its traversal step, as stand-alone
 * stage kernels.  Two differences follow from nanosg's semantics (nanosg.h:831-836: an instance is walked with the local
 * range {0, FLT_MAX}): the AO ray starts ao_min_t above the surface along the viewer-facing geometric normal (computed
 * from the hit triangle moved to world space by the instance's matrix), and it counts as occluded iff the reported world
 * distance is below ao_max_t.  NRT_AO_PACKED_TILES is not supported.  Triangle instances only.  Synchronous per wave
 * (the AO ray count is read on the host); res may be NULL. */
int nrt_scene_render_ao_device(const nrt_scene *s, const nrt_ao_params *params, float *d_accum, nrt_ao_result *res,
                               void *stream);
/* the same with DEVICE pointers, asynchronous on `stream` */
int nrt_scene_traverse_device(const nrt_scene *s, const void *d_rays_36B, size_t n_rays, void *d_hits_32B,
                              uint8_t *d_hit_mask, uint32_t flags, void *stream);


/* ------------------------------------------------------------------ BVHAccel<double>
 * The fp64 instantiation of the same path (the reference's regression program
 * test/regression/possible-accuracy-problem-30 is written against it).  Records are the reference's double layouts:
 * Ray<double> 72 B, BVHNode<double> 64 B, TriangleIntersection<double> 32 B {u, v, t, prim_id, pad},
 * BVHBuildOptions<double> 32 B; BVHTraceOptions is the same 16 B.
 *   BVHAccel<double>::Build     nanort.h:1892-2149  -> nrt_build_f64: topology from the production builder over the
 *                               float-rounded geometry, then every node box refitted EXACTLY in double from the
 *                               double vertices (leaf = min/max of its triangles, branch = union of its children)
 *   BVHAccel<double>::Traverse  nanort.h:2487-2556  -> nrt_traverse_f64, two kernels like the float path, both with the
 *                               double specialisations (IntersectRayAABB<double> nanort.h:2327-2370, DBL_EPSILON in
 *                               vsafe_inverse) in double arithmetic, t / u / v bit-identical to the reference's for
 *                               the reported primitive:
 *                                 NRT_TRAVERSE_FAST (default)  persistent warps over a private 256-B child-pair /
 *                                   96-B component-major triangle layout derived on first use (csrc/f64_fast.cuh);
 *                                   near child first by entry distance, so WHICH of two primitives hit at exactly the
 *                                   same t is reported may differ from the reference
 *                                 NRT_TRAVERSE_CONFORMANCE     one thread per ray over the BVHNode<double> array in the
 *                                   reference's visiting order: identical winners even for exact-t ties
 *                               Host buffers; rays go up and records come back in chunks through three stream slots. */
typedef struct nrt_accel_f64 nrt_accel_f64;
int nrt_build_f64(const double *verts, size_t stride_bytes, size_t n_verts_or_0, const uint32_t *faces,
                  uint32_t n_prims, const void *build_opts_32B, nrt_accel_f64 **out);
/* BVHAccel<double>::Load (nanort.h:2252-2275) + the geometry the first Traverse brings: an existing BVHNode<double>
 * array, e.g. one dumped by CPU nanort, validated and walked in the reference's order */
/* Same with the flags of nrt_build_ex: NRT_BUILD_REFERENCE_TREE makes the device write the very BVHNode<double> array and
 * indices_ the reference's BVHAccel<double>::Build writes (x-only binning, axis retry / median fallback, std::partition's
 * element order, all split decisions in double; C++11 joined node order unless NRT_BUILD_REFERENCE_CPP03_ORDER), so that
 * GetNodes() / GetIndices() / Dump() equal CPU nanort's for T = double too (csrc/build_ref64.cu). */
int nrt_build_f64_ex(const double *verts, size_t stride_bytes, size_t n_verts_or_0, const uint32_t *faces,
                     uint32_t n_prims, const void *build_opts_32B, uint32_t flags, nrt_accel_f64 **out);
int nrt_adopt_f64(const void *nodes_64B, size_t n_nodes, const uint32_t *indices, size_t n_indices, const double *verts,
                  size_t stride_bytes, size_t n_verts_or_0, const uint32_t *faces, uint32_t n_prims,
                  nrt_accel_f64 **out);
void nrt_free_f64(nrt_accel_f64 *a);
int nrt_stats_f64(const nrt_accel_f64 *a, void *stats_16B);
int nrt_bounding_box_f64(const nrt_accel_f64 *a, double bmin[3], double bmax[3]);
int nrt_nodes_f64(nrt_accel_f64 *a, const void **nodes_64B, size_t *n_nodes, const uint32_t **indices,
                  size_t *n_indices);
/* host pointers; hits[i] = {0, 0, max_t, ~0} and hit_mask[i] = 0 on a miss;
 * flags: NRT_TRAVERSE_FAST | NRT_TRAVERSE_CONFORMANCE, NRT_TRAVERSE_CPP03_INVERSE */
/* (below) the same with DEVICE pointers, asynchronous on `stream` (at most 8 launches of one accel in flight) */
int nrt_traverse_f64_device(const nrt_accel_f64 *a, const void *d_rays_72B, size_t n_rays, void *d_hits_32B,
                            uint8_t *d_hit_mask, const void *trace_opts_16B, uint32_t flags, void *stream);
int nrt_traverse_f64(const nrt_accel_f64 *a, const void *rays_72B, size_t n_rays, void *hits_32B, uint8_t *hit_mask,
                     const void *trace_opts_16B, uint32_t flags);

#ifdef __cplusplus
}
#endif
#endif /* NANORT_B200_H_ */
