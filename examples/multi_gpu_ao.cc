// Multi-GPU primary + AO frame from plain C++: one PROCESS per GPU created with fork(), no MPI, no torchrun, no
// Python.  Every rank builds the same BVH (the builder is deterministic), traces the tiles t with t % world == rank
// and takes part in ONE collective per frame, the framebuffer all-gather inside nrt_render_ao_sharded
// (include/nanort_b200.h, "multi-GPU"; NCCL over NVLink).  Rank 0 writes ao_multi.ppm and prints the frame's rays/s.
//
//   multi_gpu_ao [n_gpus (default: all)] [width height spp frames]
//
// The scene is the benchmark's procedural sphere grid (nanort_b200/scenes.py:sphere_grid, restated below so that the
// example has no dependencies).
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <chrono>
#include <vector>

#include <cuda_runtime_api.h>

#include "nanort_b200.h"

namespace {

void sphere_grid(int nx, int nz, std::vector<float> *v, std::vector<unsigned> *f) {
  const int n_lon = 25, n_lat = 21;
  const float radius = 0.4f, pi = 3.14159265358979f;
  for (int gz = 0; gz < nz; gz++) {
    for (int gx = 0; gx < nx; gx++) {
      const unsigned base = (unsigned)(v->size() / 3);
      const float cx = gx - 0.5f * (nx - 1), cz = gz - 0.5f * (nz - 1), cy = radius;
      for (int la = 0; la <= n_lat; la++) {
        const float th = pi * la / n_lat;
        for (int lo = 0; lo < n_lon; lo++) {
          const float ph = 2.0f * pi * lo / n_lon;
          v->push_back(cx + radius * sinf(th) * cosf(ph));
          v->push_back(cy + radius * cosf(th));
          v->push_back(cz + radius * sinf(th) * sinf(ph));
        }
      }
      for (int la = 0; la < n_lat; la++) {
        for (int lo = 0; lo < n_lon; lo++) {
          const unsigned a = base + la * n_lon + lo, b = base + la * n_lon + (lo + 1) % n_lon;
          const unsigned c = a + n_lon, d = b + n_lon;
          if (la > 0) {
            f->push_back(a), f->push_back(b), f->push_back(d);
          }
          if (la < n_lat - 1) {
            f->push_back(a), f->push_back(d), f->push_back(c);
          }
        }
      }
    }
  }
  const unsigned base = (unsigned)(v->size() / 3);  // floor quad
  const float e = 0.5f * nx + 1.0f, g = 0.5f * nz + 1.0f;
  const float q[12] = {-e, 0, -g, e, 0, -g, e, 0, g, -e, 0, g};
  v->insert(v->end(), q, q + 12);
  const unsigned fl[6] = {base, base + 1, base + 2, base, base + 2, base + 3};
  f->insert(f->end(), fl, fl + 6);
}

void look_at(const float org[3], const float tgt[3], float fov_deg, float aspect, float cam[12]) {
  float fw[3] = {tgt[0] - org[0], tgt[1] - org[1], tgt[2] - org[2]};
  float l = sqrtf(fw[0] * fw[0] + fw[1] * fw[1] + fw[2] * fw[2]);
  for (int k = 0; k < 3; k++) fw[k] /= l;
  const float up[3] = {0, 1, 0};
  float r[3] = {fw[1] * up[2] - fw[2] * up[1], fw[2] * up[0] - fw[0] * up[2], fw[0] * up[1] - fw[1] * up[0]};
  l = sqrtf(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  for (int k = 0; k < 3; k++) r[k] /= l;
  const float u[3] = {r[1] * fw[2] - r[2] * fw[1], r[2] * fw[0] - r[0] * fw[2], r[0] * fw[1] - r[1] * fw[0]};
  const float sy = 2.0f * tanf(0.5f * fov_deg * 3.14159265358979f / 180.0f), sx = sy * aspect;
  for (int k = 0; k < 3; k++) {
    cam[k] = org[k];
    cam[3 + k] = r[k] * sx;
    cam[6 + k] = u[k] * sy;
    cam[9 + k] = fw[k];
  }
}

#define CHECK(call)                                                               \
  do {                                                                            \
    if ((call) != NRT_OK) {                                                       \
      fprintf(stderr, "rank %d: %s failed: %s\n", rank, #call, nrt_last_error()); \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

int run_rank(int rank, int world, const unsigned char *id, int width, int height, int spp, int frames) {
  std::vector<float> v;
  std::vector<unsigned> f;
  sphere_grid(10, 10, &v, &f);
  CHECK(nrt_set_device(rank));
  nrt_accel *accel = NULL;
  CHECK(nrt_build(v.data(), 12, v.size() / 3, f.data(), (uint32_t)(f.size() / 3), NULL, &accel));
  nrt_comm *comm = NULL;
  CHECK(nrt_comm_init(id, rank, world, &comm));

  nrt_ao_params p;
  memset(&p, 0, sizeof(p));
  const float org[3] = {0.0f, 6.0f, 11.0f}, tgt[3] = {0.0f, 0.3f, 0.0f};
  look_at(org, tgt, 45.0f, (float)width / height, p.cam);
  float bmin[3], bmax[3];
  CHECK(nrt_bounding_box(accel, bmin, bmax));
  const float diag = sqrtf((bmax[0] - bmin[0]) * (bmax[0] - bmin[0]) + (bmax[1] - bmin[1]) * (bmax[1] - bmin[1]) +
                           (bmax[2] - bmin[2]) * (bmax[2] - bmin[2]));
  p.width = width, p.height = height, p.spp = spp, p.sample0 = 0, p.seed = 1;
  p.tile_w = 64, p.tile_h = 8;
  p.ray_min_t = 1e-3f, p.ray_max_t = 1e30f, p.ao_min_t = 1e-3f, p.ao_max_t = 0.25f * diag;

  cudaSetDevice(rank);
  float *d_frame = NULL;
  if (cudaMalloc(reinterpret_cast<void **>(&d_frame), sizeof(float) * width * height) != cudaSuccess) return 1;
  nrt_ao_result res;
  CHECK(nrt_render_ao_sharded(accel, comm, &p, d_frame, &res, NULL));  // warm-up (allocations, NCCL channels)
  cudaDeviceSynchronize();
  const auto t0 = std::chrono::steady_clock::now();
  for (int i = 0; i < frames; i++) CHECK(nrt_render_ao_sharded(accel, comm, &p, d_frame, i + 1 == frames ? &res : NULL, NULL));
  cudaDeviceSynchronize();
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  std::vector<float> frame((size_t)width * height);
  cudaMemcpy(frame.data(), d_frame, sizeof(float) * frame.size(), cudaMemcpyDeviceToHost);
  double sum = 0.0;
  for (float x : frame) sum += x;
  // every rank holds the WHOLE frame: its sum is the same on all ranks and equals (primary misses + unoccluded AO
  // rays) summed over the ranks; each rank prints its share so the caller can check the identity
  printf("rank %d/%d: %llu primary + %llu AO rays (%llu occluded) per frame, frame sum %.0f, %.2f ms/frame\n", rank, world,
         (unsigned long long)res.primary_rays, (unsigned long long)res.ao_rays, (unsigned long long)res.ao_hits, sum,
         1e3 * dt / frames);
  if (rank == 0) {
    const double rays = (double)width * height * spp * (1.0 + (double)res.ao_rays / (double)res.primary_rays);
    printf("frame %dx%dx%d spp on %d GPU(s): ~%.1f Mrays/s\n", width, height, spp, world, rays * frames / dt / 1e6);
    FILE *fp = fopen("ao_multi.ppm", "wb");
    if (fp) {
      fprintf(fp, "P6\n%d %d\n255\n", width, height);
      for (size_t i = 0; i < frame.size(); i++) {
        const unsigned char c = (unsigned char)(255.0f * fminf(1.0f, frame[i] / spp));
        fputc(c, fp), fputc(c, fp), fputc(c, fp);
      }
      fclose(fp);
    }
  }
  cudaFree(d_frame);
  nrt_comm_free(comm);
  nrt_free(accel);
  fflush(stdout);  // the caller leaves through _exit(), which does not flush stdio
  return 0;
}

}  // namespace

// CUDA contexts do not survive fork(): the parent never touches CUDA, even the device count is asked in a child
int device_count_in_child() {
  const pid_t pid = fork();
  if (pid == 0) _exit(nrt_device_count() & 0xFF);
  int st = 0;
  waitpid(pid, &st, 0);
  return WIFEXITED(st) ? WEXITSTATUS(st) : 0;
}

int main(int argc, char **argv) {
  int world = argc > 1 ? atoi(argv[1]) : device_count_in_child();
  const int width = argc > 2 ? atoi(argv[2]) : 1920, height = argc > 3 ? atoi(argv[3]) : 1080;
  const int spp = argc > 4 ? atoi(argv[4]) : 64, frames = argc > 5 ? atoi(argv[5]) : 5;
  if (world < 1) {
    fprintf(stderr, "no CUDA device: %s\n", nrt_last_error());
    return 1;
  }
  // the NCCL unique id travels through a page shared with the children (created BEFORE any CUDA call in this process:
  // CUDA contexts do not survive fork(), so the parent only forks and waits)
  unsigned char *shared = (unsigned char *)mmap(NULL, 4096, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
  if (shared == MAP_FAILED) return 1;
  volatile unsigned char *ready = shared + 128;
  std::vector<pid_t> kids;
  for (int rank = 0; rank < world; rank++) {
    pid_t pid = fork();
    if (pid == 0) {
      if (rank == 0) {
        if (nrt_comm_unique_id(shared) != NRT_OK) {
          fprintf(stderr, "nrt_comm_unique_id: %s\n", nrt_last_error());
          *ready = 2;
          _exit(1);
        }
        __sync_synchronize();
        *ready = 1;
      } else {
        while (*ready == 0) usleep(1000);
        if (*ready != 1) _exit(1);
      }
      _exit(run_rank(rank, world, shared, width, height, spp, frames));
    }
    kids.push_back(pid);
  }
  int bad = 0;
  for (pid_t k : kids) {
    int st = 0;
    waitpid(k, &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) bad++;
  }
  return bad ? 1 : 0;
}
