#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric: Mrays/s, primary + 1-bounce AO rays.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): procedural 100,002-triangle sphere grid, 1920x1080, 16 spp of
jittered pinhole primary rays + one cosine-hemisphere AO ray per primary hit (closest-hit queries, as the
reference's CheckForOccluder does).  One "step" = one pass over all samples of the image.

  value  : device-resident wavefront pass (nrt_render_ao_device), scene + BVH already in HBM.
  e2e    : the same rays through the reference-facing batch call nrt_traverse with HOST buffers (pinned):
           H2D of 36-byte rays and D2H of 16-byte hits + 1-byte flags are inside the timed region.
  N > 1  : one process per GPU (torchrun); WEAK scaling -- the image gets 16*N spp, tiles are dealt
           round-robin to the ranks (each traces as many rays as the single-GPU job), the BVH is rebuilt
           identically on every rank, and the only collective is the framebuffer all_gather (NCCL) at
           the end of each step, inside the timed region.
  --impl reference : the UNMODIFIED reference (oracle/_ref, built from /root/reference/nanort.h) -- or the
           oracle port when that is absent -- traces a bounded sample of the very same ray arrays on the
           host cores; rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, SPP = 1920, 1080, 16
TILE_W, TILE_H = 64, 8
SCENE = "sphere_grid"
METRIC = "Mrays/sec (primary+1-bounce AO)"
UNIT = "Mrays/s"


# ----------------------------------------------------------------------------------- helpers
def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms while the timed region runs."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock and throttle reasons of the samples taken inside [t_begin, t_end] (time.time());
        the sampler is started before the warm-up so that nvidia-smi is already streaming when the timed
        region begins."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        import datetime

        sm, mx, reasons, sm_all = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, cmax = float(f[1]), float(f[2])
            except ValueError:
                continue
            sm_all.append(clk)
            if t_begin is not None and not (t_begin - 0.05 <= ts <= t_end + 0.05):
                continue
            sm.append(clk)
            mx.append(cmax)
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else (float(np.median(sm_all)) if sm_all else None),
                "sm_max_mhz": max(mx) if mx else None, "samples_in_timed_region": len(sm),
                "samples_total": len(sm_all), "reasons": sorted(reasons)}


def ao_params(api, S, cam, diag, n_shards, shard):
    p = api.AoParams()
    for i in range(12):
        p.cam[i] = float(cam[i])
    p.width, p.height, p.spp, p.sample0, p.seed = WIDTH, HEIGHT, SPP * n_shards, 0, 1
    p.tile_w, p.tile_h, p.shard, p.n_shards = TILE_W, TILE_H, shard, n_shards
    p.ray_min_t, p.ray_max_t, p.ao_min_t, p.ao_max_t = 1e-3, 1e30, 1e-3, 0.25 * diag
    p.flags = 0
    return p


def config_dict(n_gpus):
    return {
        "workload": f"sphere_grid 100,002 triangles (BASELINE.json configs[1]), {WIDTH}x{HEIGHT}, {SPP}*N spp "
                    f"(N={n_gpus}: {SPP * n_gpus} spp), primary + 1 cosine AO ray per hit, closest-hit; "
                    f"tiles {TILE_W}x{TILE_H} round-robin over ranks; BVH replicated; framebuffer all_gather",
        "rays_per_gpu_per_step": "33,177,600 primary + ~18.2 M AO, in 2 waves of 16 Mi camera rays",
        "l2_policy": "inputs larger than L2: camera rays are generated inside the traversal kernel (no input stream); the "
                     "only stream between the two launches of a wave is the compacted AO queue, ~9 M rays x 36 B = 330 MB per "
                     "16 Mi-ray wave = 2.6 x the 126 MB L2; scene + BVH (7 MB) stay cache-resident by design",
        "parallelism": f"ray-tile sharding x{n_gpus}",
    }


# ----------------------------------------------------------------------------------- reference arm
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class CpuReference:
    """The reference's own CPU implementation of the path: oracle/_ref (unmodified nanort.h) when it was
    built, else the oracle port.  This is the one place outside tests/ that executes oracle/."""

    def __init__(self, verts, faces):
        from oracle import orc

        self.threads = host_threads()
        t0 = time.time()
        if orc.Reference.available(True):
            self.kind = "reference"
            self.acc = orc.Reference(True).build(verts, faces)
            self.trav = lambda rays: self.acc.traverse(rays, threads=self.threads)
        else:
            self.kind = "port"
            port = orc.Port()
            nodes, idx, _ = port.build(verts, faces, mode=orc.MODE_CPP11)
            self.trav = lambda rays: port.traverse(nodes, idx, verts, faces, rays, threads=self.threads)
        self.build_s = time.time() - t0

    def calibrate(self, primary, ao, target_s=12.0):
        """Picks a strided sample of the two exported ray arrays that takes about target_s seconds."""
        probe_n = 20000
        for _ in range(4):  # grow the probe until it runs long enough to give a stable rate
            sp = primary[:: max(1, len(primary) // probe_n)]
            sa = ao[:: max(1, len(ao) // probe_n)]
            t0 = time.time()
            self.trav(sp)
            self.trav(sa)
            dt = max(time.time() - t0, 1e-3)
            if dt > 1.0 or probe_n >= len(primary):
                break
            probe_n *= 6
        rate = (len(sp) + len(sa)) / dt
        want = int(rate * target_s)
        frac = min(1.0, want / float(len(primary) + len(ao)))
        kp = max(1, int(round(1.0 / frac)))
        self.sample_primary = np.ascontiguousarray(primary[::kp])
        self.sample_ao = np.ascontiguousarray(ao[::kp])
        self.sample_desc = (f"every {kp}-th ray of the step's {len(primary)} primary + {len(ao)} AO rays "
                            f"({len(self.sample_primary)} + {len(self.sample_ao)} rays)")

    def calibrate_standalone(self, S, verts, faces, cam, ao_max_t, target_s=12.0):
        """Reference arm: no GPU code anywhere.  The sample's rays come from the numpy generators
        (same camera / jitter hash as the device pass): every k-th pixel, all SPP samples, AO rays from the
        reference's own primary hits."""

        def make(stride):
            pixels = np.arange(0, WIDTH * HEIGHT, stride, dtype=np.int64)
            prim = S.primary_rays(cam, WIDTH, HEIGHT, spp=SPP, seed=1, pixels=pixels, min_t=1e-3, max_t=1e30)
            hits, mask = self.trav(prim)
            ao, _ = S.ao_rays(verts, faces, prim, hits, mask, seed=2, min_t=1e-3, max_t=ao_max_t)
            return prim, ao

        stride = max(1, (WIDTH * HEIGHT * SPP) // 20000)
        for _ in range(4):  # grow the probe until it runs long enough to give a stable rate
            prim, ao = make(stride)
            t0 = time.time()
            self.trav(prim)
            self.trav(ao)
            dt = max(time.time() - t0, 1e-3)
            if dt > 1.0 or stride == 1:
                break
            stride = max(1, stride // 6)
        rate = (len(prim) + len(ao)) / dt
        want = rate * target_s
        full = WIDTH * HEIGHT * SPP * (1.0 + len(ao) / max(1, len(prim)))
        stride = max(1, int(round(full / want)))
        self.sample_primary, self.sample_ao = make(stride)
        self.sample_desc = (f"every {stride}-th pixel of {WIDTH}x{HEIGHT} at {SPP} spp, numpy-generated: "
                            f"{len(self.sample_primary)} primary + {len(self.sample_ao)} AO rays")

    def step(self):
        t0 = time.time()
        self.trav(self.sample_primary)
        self.trav(self.sample_ao)
        return time.time() - t0, len(self.sample_primary) + len(self.sample_ao)


# ----------------------------------------------------------------------------------- main
def bind_to_gpu_numa_node(torch, local_rank):
    """Multi-rank runs: keep this rank's threads -- and with them its pinned host buffers (first touch) -- on the CPUs
    NVML reports as local to the rank's GPU, so that the host-buffer arm of 8 ranks does not cross the socket
    interconnect.  Best effort; any failure leaves the default affinity."""
    try:
        import pynvml

        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world

    if args.impl == "reference":
        if rank != 0:
            return 0  # rank 0 alone runs and prints the reference arm
        from nanort_b200 import scenes as S

        verts, faces = S.make_scene(SCENE)
        cam = S.scene_camera(SCENE, WIDTH, HEIGHT)
        diag = float(np.linalg.norm(verts.max(axis=0) - verts.min(axis=0)))
        ref = CpuReference(verts, faces)
        # every step traces the same bounded sample; the whole run (W + K steps) is kept to about 2.5 minutes
        per_step = max(2.0, min(12.0, 150.0 / max(1, args.steps + args.warmup)))
        if os.environ.get("NRT_BENCH_REF_STEP_S"):  # tests shrink the sample (tests/test_bench_contract.py)
            per_step = float(os.environ["NRT_BENCH_REF_STEP_S"])
        ref.calibrate_standalone(S, verts, faces, cam, 0.25 * diag, target_s=per_step)
        for _ in range(args.warmup):
            ref.step()
        tot_t, tot_n = 0.0, 0
        for _ in range(args.steps):
            dt, n = ref.step()
            tot_t += dt
            tot_n += n
        val = tot_n / tot_t / 1e6
        line = {
            "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_t / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": ref.threads, "kind": ref.kind,
                             "sample": ref.sample_desc, "build_s": ref.build_s},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line), flush=True)
        return 0

    import torch

    from nanort_b200 import api, dist as nd, scenes as S

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        bind_to_gpu_numa_node(torch, local_rank)
    distributed = world > 1
    if distributed:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    verts, faces = S.make_scene(SCENE)
    cam = S.scene_camera(SCENE, WIDTH, HEIGHT)
    n_shards, shard = world, rank

    acc = api.BVHAccel(device=local_rank)
    t0 = time.time()
    acc.Build(len(faces), verts, faces)
    build_wall_ms = (time.time() - t0) * 1e3
    stats = acc.GetStatistics()
    bmin, bmax = acc.BoundingBox()
    diag = float(np.linalg.norm(bmax - bmin))
    p = ao_params(api, S, cam, diag, n_shards, shard)
    accum = torch.zeros(WIDTH * HEIGHT, dtype=torch.float32, device=dev)

    # ---- the step's ray arrays, exported once (untimed) for the host-buffer arm, the CPU baseline and the counters
    slots = nd.shard_ray_count(WIDTH, HEIGHT, TILE_W, TILE_H, shard, n_shards, p.spp)
    d_primary = torch.empty(slots * 36, dtype=torch.uint8, device=dev)
    d_ao = torch.empty(slots * 36, dtype=torch.uint8, device=dev)
    n_primary, n_ao = acc.ExportAOWorkload(p, accum.data_ptr(), d_primary.data_ptr(), d_ao.data_ptr())
    assert n_primary == slots, (n_primary, slots)
    accum.zero_()

    gather = nd.FramebufferGather(WIDTH, HEIGHT, TILE_W, TILE_H, world, rank, dev) if distributed else None

    def device_step():
        accum.zero_()
        r = acc.RenderAO(p, accum.data_ptr(), want_result=False)
        if gather is not None:
            gather.gather(accum)
        return r

    def sync_all():
        torch.cuda.synchronize(dev)
        if distributed:
            dist.barrier()
            torch.cuda.synchronize(dev)

    # ---- value: device-resident pass
    sampler = ClockSampler(local_rank)
    sampler.start()
    for _ in range(args.warmup):
        device_step()
    sync_all()
    time.sleep(0.3)  # let nvidia-smi reach its streaming state before the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    t_begin = time.time()
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    sync_all()
    t_end = time.time()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t_begin, t_end)
    # one more instrumented pass (outside the timed region) for counts, launch counts and the in-kernel time
    accum.zero_()
    r = acc.RenderAO(p, accum.data_ptr(), want_result=True)
    rays_step = int(r.primary_rays + r.ao_rays)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    tot = torch.tensor([rays_step], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max, rays_all = float(t.item()), float(tot.item())
    value = rays_all * args.steps / (ms_max * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (traverse_fast2_kernel, both instantiations of a wave): algorithmic bytes / in-kernel time
    boxes_p, prims_p = acc.CountDevice(d_primary.data_ptr(), n_primary)
    boxes_a, prims_a = acc.CountDevice(d_ao.data_ptr(), n_ao)
    alg_bytes = 52.0 * (n_primary + n_ao) + 40.0 * (boxes_p + boxes_a) + 52.0 * (prims_p + prims_a)
    trav_ms = float(r.traverse_ms)
    peak, peak_src = measured_peak_gbs()
    achieved = alg_bytes / (trav_ms * 1e-3) / 1e9
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traverse_traffic.json")))
        # the capture's launches are smaller than a bench wave: scale its DRAM bytes per ray to this run's launch size
        if tj.get("dram_bytes_per_ray"):
            traffic = tj["dram_bytes_per_ray"] * (n_primary + n_ao) / max(1, r.traverse_launches)
        else:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {
        "bound": "hbm", "kernel": "traverse_fast2_kernel<CameraRays, ..., PrimaryToAoEpilogue> + traverse_fast2_kernel<SoaRays, ..., AoAccumulateEpilogue>",
        "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "peak_source": peak_src, "traffic": traffic,
        "launches_per_step": int(r.traverse_launches), "avg_launch_ms": trav_ms / max(1, r.traverse_launches),
        "alg_bytes_per_launch": alg_bytes / max(1, r.traverse_launches),
        "bytes_per_ray": alg_bytes / (n_primary + n_ao),
        "boxes_per_ray": (boxes_p + boxes_a) / (n_primary + n_ao), "prims_per_ray": (prims_p + prims_a) / (n_primary + n_ao),
        "traverse_share_of_step": trav_ms / float(r.total_ms),
    }

    # ---- e2e: host buffers through nrt_traverse (H2D rays, D2H hits + flags inside the timed region)
    e2e = None
    if not args.no_e2e:
        hp = api.PinnedArray(n_primary, S.RAY_DTYPE)
        ha = api.PinnedArray(max(n_ao, 1), S.RAY_DTYPE)
        hp.array[:] = d_primary.cpu().numpy().view(S.RAY_DTYPE)
        ha.array[:n_ao] = d_ao[: n_ao * 36].cpu().numpy().view(S.RAY_DTYPE)
        hits_p, mask_p = api.PinnedArray(n_primary, S.HIT_DTYPE), api.PinnedArray(n_primary, np.uint8)
        hits_a, mask_a = api.PinnedArray(max(n_ao, 1), S.HIT_DTYPE), api.PinnedArray(max(n_ao, 1), np.uint8)

        def host_step():
            acc.Traverse(hp.array, hits=hits_p.array, mask=mask_p.array)
            acc.Traverse(ha.array[:n_ao], hits=hits_a.array[:n_ao], mask=mask_a.array[:n_ao])

        for _ in range(2):
            host_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            host_step()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        # the host-buffer arm must report what the device pass found
        assert int(mask_p.array.sum()) == n_ao, "primary hits of the host-buffer arm != AO ray count"
        e2e = {"value": rays_all * args.steps / float(tt.item()) / 1e6, "unit": UNIT,
               "h2d_bytes_per_step": int(36 * (n_primary + n_ao)), "d2h_bytes_per_step": int(17 * (n_primary + n_ao)),
               "api": "nrt_traverse (host rays -> host hits), pinned buffers, 2 calls per step"}
        # for comparison, the wavefront entry point end to end: camera parameters in (host struct), framebuffer out to
        # pinned host memory every step -- what a renderer pays when it hands the whole pass to the library
        fb_host = torch.empty(WIDTH * HEIGHT, dtype=torch.float32).pin_memory()

        def render_step():
            accum.zero_()
            acc.RenderAO(p, accum.data_ptr(), want_result=False)
            fb_host.copy_(accum, non_blocking=True)
            torch.cuda.synchronize(dev)

        for _ in range(2):
            render_step()
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            render_step()
        dt = time.perf_counter() - t0
        tr = torch.tensor([dt], dtype=torch.float64, device=dev)
        if distributed:
            dist.all_reduce(tr, op=dist.ReduceOp.MAX)
        e2e["render_api"] = {"value": rays_all * args.steps / float(tr.item()) / 1e6, "unit": UNIT,
                             "api": "nrt_render_ao_device + framebuffer D2H per step",
                             "h2d_bytes_per_step": C.sizeof(api.AoParams), "d2h_bytes_per_step": WIDTH * HEIGHT * 4}

    # ---- CPU baseline beside it (rank 0, N = 1 only): bounded sample of the same rays
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        h_primary = d_primary.cpu().numpy().view(S.RAY_DTYPE)
        h_ao = d_ao[: n_ao * 36].cpu().numpy().view(S.RAY_DTYPE)
        ref = CpuReference(verts, faces)
        ref.calibrate(h_primary, h_ao, target_s=15.0)
        dt, n = ref.step()
        cpu_baseline = {"value": n / dt / 1e6, "unit": UNIT, "cores": ref.threads, "kind": ref.kind,
                        "sample": ref.sample_desc, "build_s": ref.build_s}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config_dict(args.gpus),
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(r.launches) * args.steps,
            "roofline": roofline, "cpu_baseline": cpu_baseline,
            "build": {"device_ms": stats["build_secs"] * 1e3, "wall_ms_incl_upload": build_wall_ms,
                      "nodes": stats["num_leaf_nodes"] + stats["num_branch_nodes"], "depth": stats["max_tree_depth"]},
            "rays_per_step": rays_all, "ao_occluded_fraction": float(r.ao_hits) / max(1, r.ao_rays),
        }
        print(json.dumps(line), flush=True)
    if distributed:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
