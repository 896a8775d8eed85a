// TEST INFRASTRUCTURE ONLY -- the reference path tracer's own shading code behind a C interface.
//
// Compiles the UNMODIFIED /root/reference/examples/path_tracer/main.cc into this translation unit (its main() is
// renamed away, nothing of it is copied into the repository) so that the checker of tests/test_gpu_path.py calls the
// reference's OWN functions for everything the example factors out:
//     float3 / vdot / vcross / normalize            main.cc:111-208
//     uniformFloat (rand() based)                   main.cc:210-212
//     revisedONB + directionCosTheta                main.cc:216-250
//     PdfAtoW                                       main.cc:252-255
//     calcNormal                                    main.cc:306-312
//     MeshLight (constructor + sampleDirect)        main.cc:321-399
//     sign / reflect / refract / pow5 / fresnel_schlick   main.cc:644-663
// The per-hit block of the bounce loop (main.cc:826-976) is written inline in the example's main(); pt_ref_shade below
// walks through it statement by statement, calling the functions above, for ONE bounce of n given rays with their hit
// records -- the unit the device's retire step (csrc/wavefront.cuh: PathShadeEpilogue) implements.
//
// Random numbers: the example draws from libc rand() through uniformFloat.  This file defines its own rand(), which
// serves numbers queued by the caller, so that the reference's draws are exactly the counter-hash values the device
// uses (tests generate them with nanort_b200/scenes.py:rand_ps): a queued u in [0,1) with 24 significant bits comes
// back from uniformFloat(0, 1) bit for bit (float(r) / RAND_MAX with r = u * 2^31 and float(RAND_MAX) == 2^31).
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

static float g_queue[8];
static int g_q_head = 0, g_q_tail = 0;
static long g_unexpected_draws = 0;
extern "C" int rand(void) noexcept {  // replaces libc's for this shared object
  if (g_q_head >= g_q_tail) {
    g_unexpected_draws++;
    return 0;
  }
  const float u = g_queue[g_q_head++];
  return (int)(u * 2147483648.0f);  // exact: u = k * 2^-24, k < 2^24
}
static void queue_draws(float a, float b) {
  g_queue[0] = a;
  g_queue[1] = b;
  g_q_head = 0;
  g_q_tail = 2;
}

#define main nanort_reference_path_tracer_main
#include "main.cc"  // found through -I/root/reference/examples/path_tracer (oracle/Makefile)
#undef main

extern "C" {

long pt_ref_unexpected_draws(void) { return g_unexpected_draws; }

// 16 floats per material, the tinyobj fields main.cc reads: diffuse[3] specular[3] transmittance[3] emission[3] ior
// dissolve pad pad  (= csrc/wavefront.cuh: PathMaterial)
struct PtRefScene {
  Mesh mesh;
  std::vector<tinyobj::material_t> materials;
  MeshLight *lights;
};

void *pt_ref_scene(const float *verts, size_t n_verts, const unsigned int *faces, size_t n_faces,
                   const unsigned int *material_ids, const float *facevarying_normals, const float *materials16,
                   size_t n_materials) {
  PtRefScene *s = new PtRefScene();
  memset(&s->mesh, 0, sizeof(Mesh));
  s->mesh.num_vertices = n_verts;
  s->mesh.num_faces = n_faces;
  s->mesh.vertices = const_cast<float *>(verts);  // borrowed: the caller keeps the arrays alive
  s->mesh.faces = const_cast<unsigned int *>(faces);
  s->mesh.material_ids = const_cast<unsigned int *>(material_ids);
  s->mesh.facevarying_normals = const_cast<float *>(facevarying_normals);
  s->materials.resize(n_materials);
  for (size_t i = 0; i < n_materials; i++) {
    tinyobj::material_t &m = s->materials[i];
    for (int k = 0; k < 3; k++) m.ambient[k] = 0.0f;
    m.shininess = 1.0f;
    m.illum = 0;
    const float *p = materials16 + 16 * i;
    for (int k = 0; k < 3; k++) {
      m.diffuse[k] = p[k];
      m.specular[k] = p[3 + k];
      m.transmittance[k] = p[6 + k];
      m.emission[k] = p[9 + k];
    }
    m.ior = p[12];
    m.dissolve = p[13];
  }
  s->lights = new MeshLight(s->mesh, s->materials);  // the example's own emissive-face list (main.cc:323-335)
  return s;
}

void pt_ref_scene_free(void *h) {
  PtRefScene *s = static_cast<PtRefScene *>(h);
  if (!s) return;
  delete s->lights;
  delete s;
}

size_t pt_ref_emissive_faces(void *h, unsigned int *out, size_t cap) {
  PtRefScene *s = static_cast<PtRefScene *>(h);
  const size_t n = s->lights->emissive_faces_.size();
  for (size_t i = 0; i < n && i < cap; i++) out[i] = s->lights->emissive_faces_[i].face_;
  return n;
}

// The loader's flat face normal (LoadObj uses calcNormal when the OBJ has none, main.cc:566-601): 9 floats per face.
void pt_ref_face_normals(const float *verts, const unsigned int *faces, size_t n_faces, float *out9) {
  for (size_t f = 0; f < n_faces; f++) {
    float3 v[3];
    for (int k = 0; k < 3; k++) v[k] = float3(verts + 3 * faces[3 * f + k]);
    float3 N;
    calcNormal(N, v[0], v[1], v[2]);
    for (int k = 0; k < 3; k++) {
      out9[9 * f + 3 * k + 0] = N[0];
      out9[9 * f + 3 * k + 1] = N[1];
      out9[9 * f + 3 * k + 2] = N[2];
    }
  }
}

// One bounce (index b, 0-based) of main.cc's loop body for n rays that HIT (records u, v, t, prim).  Per ray in:
//   org[3], dir[3]            the ray that was traversed (rayOrg, rayDir before `rayOrg += rayDir * isect.t`)
//   hit {u, v, t, prim}       TriangleIntersection of that Traverse
//   weight_in[4]              {weight.rgb, do_emmition} BEFORE this bounce's Russian-roulette factor... see below
//   draws[6]                  {Xi1, Xi2 (sampleDirect), u1, phi01 (directionCosTheta), rr_next, pick}
// The device folds the roulette of bounce b+1 into the end of bounce b (it decides there whether a continuation ray is
// queued); the same split is used here: this function applies main.cc:856-976 for bounce b and then main.cc:828-837
// for bounce b+1 (and the b+1 < max_bounces loop condition), so that outputs line up with the device's queues.
// Per ray out:
//   flags          bit0 continuation ray queued, bit1 shadow ray queued, bit2 emission added
//   next_org[3], next_dir[3], weight_out[4]
//   shadow_org[3], shadow_dir[3], shadow_max_t, shadow_contrib[3]   (contribution if the light sample is visible)
//   emission[3]    radiance added to the pixel by the EMIT branch
void pt_ref_shade(void *h, size_t n, unsigned int b, unsigned int max_bounces, const float *org, const float *dir,
                  const float *hit_uvt, const unsigned int *hit_prim, const float *weight_in, const float *draws,
                  unsigned int *flags, float *next_org, float *next_dir, float *weight_out, float *shadow_org,
                  float *shadow_dir, float *shadow_max_t, float *shadow_contrib, float *emission) {
  PtRefScene *s = static_cast<PtRefScene *>(h);
  const Mesh &mesh = s->mesh;
  const std::vector<tinyobj::material_t> &materials = s->materials;
  const MeshLight &lights = *s->lights;
  for (size_t i = 0; i < n; i++) {
    float3 rayOrg(org + 3 * i), rayDir(dir + 3 * i);
    float3 weight(weight_in + 4 * i);
    bool do_emmition = weight_in[4 * i + 3] != 0.0f;
    const float *dr = draws + 6 * i;
    unsigned int fl = 0;
    float3 emit(0, 0, 0), sh_o(0, 0, 0), sh_d(0, 0, 0), sh_c(0, 0, 0), outDir(0, 0, 0);
    float sh_t = 0.0f;
    bool terminated = false;

    // ---- main.cc:856 onwards, `hit` is true
    const float t = hit_uvt[3 * i + 2];
    rayOrg += rayDir * t;
    unsigned int fid = hit_prim[i];
    float3 norm(0, 0, 0);
    if (mesh.facevarying_normals) {
      float3 normals[3];
      for (int vId = 0; vId < 3; vId++) {
        normals[vId][0] = mesh.facevarying_normals[9 * fid + 3 * vId + 0];
        normals[vId][1] = mesh.facevarying_normals[9 * fid + 3 * vId + 1];
        normals[vId][2] = mesh.facevarying_normals[9 * fid + 3 * vId + 2];
      }
      float u = hit_uvt[3 * i + 0];
      float v = hit_uvt[3 * i + 1];
      norm = (1.0 - u - v) * normals[0] + u * normals[1] + v * normals[2];
      norm.normalize();
    }
    float3 originalNorm = norm;
    if (vdot(norm, rayDir) > 0) {
      norm *= -1;
    }
    unsigned int matId = mesh.material_ids[fid];
    tinyobj::material_t mat = materials[matId];
    float3 diffuseColor(mat.diffuse);
    float3 emissiveColor(mat.emission);
    float3 specularColor(mat.specular);
    float3 refractionColor(mat.transmittance);
    float ior = mat.ior;
    float inside = sign(vdot(rayDir, originalNorm));
    float n1 = inside < 0 ? 1.0 / ior : ior;
    float n2 = 1.0 / n1;
    float fresnel = fresnel_schlick(-rayDir, norm, (n1 - n2) / (n1 + n2));
    float rhoS = vdot(float3(1, 1, 1) / 3.0f, specularColor) * fresnel;
    float rhoD = vdot(float3(1, 1, 1) / 3.0f, diffuseColor) * (1.0 - fresnel) * (1.0 - mat.dissolve);
    float rhoR = vdot(float3(1, 1, 1) / 3.0f, refractionColor) * (1.0 - fresnel) * mat.dissolve;
    float rhoE = vdot(float3(1, 1, 1) / 3.0f, emissiveColor);
    float totalrho = rhoS + rhoD + rhoR + rhoE;
    if (totalrho < 0.0001) {
      terminated = true;
    } else {
      rhoS /= totalrho;
      rhoD /= totalrho;
      rhoR /= totalrho;
      rhoE /= totalrho;
      float rand = dr[5];  // `float rand = uniformFloat(0, 1);`
      if (rand < rhoS) {
        outDir = reflect(rayDir, norm);
        weight *= specularColor;
        do_emmition = true;
      } else if (rand < rhoS + rhoD) {
        float3 brdfEval = (1.0f / M_PI) * diffuseColor;
        float3 ldir, ll;
        float lpdf, ldist;
        lights.sampleDirect(rayOrg, dr[0], dr[1], ldir, ldist, lpdf, ll);  // the reference's function; the two numbers
                                                                           // are its uniformFloat(0,1) arguments
        if (lpdf > 0.0f) {
          float cosTheta = std::abs(vdot(ldir, norm));
          float3 directLight = (brdfEval * ll * cosTheta) / lpdf;
          // CheckForOccluder(rayOrg, rayOrg + ldir * ldist) (main.cc:675-701): the shadow ray it builds
          {
            static const float ray_eps = 0.00001f;
            float3 p1 = rayOrg, p2 = rayOrg + ldir * ldist;
            float3 sdir = p2 - p1;
            float dist = sdir.length();
            sdir.normalize();
            sh_o = p1;
            sh_d = sdir;
            sh_t = dist - ray_eps;
          }
          sh_c = directLight * weight;  // `color += directLight * visible * weight`
          fl |= 2u;
        }
        queue_draws(dr[2], dr[3]);  // directionCosTheta draws u1 then phi through uniformFloat -> rand()
        outDir = directionCosTheta(norm);
        weight *= diffuseColor;
        do_emmition = false;
      } else if (rand < rhoD + rhoS + rhoR) {
        outDir = refract(rayDir, -inside * originalNorm, n1);
        weight *= refractionColor;
        do_emmition = true;
      } else {
        if (do_emmition) {
          emit = std::max(vdot(originalNorm, -rayDir), 0.0f) * emissiveColor * weight;
          fl |= 4u;
        }
        terminated = true;
      }
    }
    // ---- top of the next iteration (main.cc:827-837): loop condition and Russian roulette of bounce b + 1
    if (!terminated && b + 1 < max_bounces) {
      float rr_fac = 1.0f;
      bool alive = true;
      if (b + 1 > 3) {
        float rr_rand = dr[4];
        float termination_probability = 0.2f;
        if (rr_rand < termination_probability) alive = false;
        rr_fac = 1.0 - termination_probability;
      }
      if (alive) {
        weight *= 1.0 / rr_fac;
        fl |= 1u;
      }
    }
    flags[i] = fl;
    for (int k = 0; k < 3; k++) {
      next_org[3 * i + k] = rayOrg[k];
      next_dir[3 * i + k] = outDir[k];
      weight_out[4 * i + k] = weight[k];
      shadow_org[3 * i + k] = sh_o[k];
      shadow_dir[3 * i + k] = sh_d[k];
      shadow_contrib[3 * i + k] = sh_c[k];
      emission[3 * i + k] = emit[k];
    }
    weight_out[4 * i + 3] = do_emmition ? 1.0f : 0.0f;
    shadow_max_t[i] = sh_t;
  }
}

}  // extern "C"
