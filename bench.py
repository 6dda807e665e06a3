#!/usr/bin/env python
"""bench.py — the hot path of BASELINE.json measured on B200(s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mesh soup|grid]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the batched closest-hit traversal over 10M coherent primary rays (3163 x 3163
pinhole camera, reference test/benchmark.cpp:340-358) against a 1M-triangle synthetic mesh
(BASELINE.json configs[1]); with N GPUs every rank traces its own 10M-ray shard of an N x 10M batch
(BVH replicated, weak scaling) and the 16-byte hit records are all-gathered over NCCL chunk by chunk,
overlapped with the traversal.  `value` is whole-job Mrays/s with rays resident in HBM; `build` reports
the LBVH build rate of the same mesh (Mtris/s) timed in the same run; `e2e` is the same metric through
the C ABI with pinned HOST buffers (H2D of the rays and D2H of the hits inside the timed region).

`--impl reference` times the reference's own CPU implementation (oracle/_ref = the unmodified
reference compiled in place; the plain-C oracle port when that is absent) on a bounded sample of the
same workload with every host thread, and prints the same JSON line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from bvh_b200 import scenes  # noqa: E402

TRIS = 1_000_000
IMG = 3163                       # 3163^2 = 10 004 569 rays
CPU_SAMPLE_ROWS = 320            # ~1M rays for the CPU arm (a band of image rows through the centre)
FALLBACK_HBM_GBS = 6650.0        # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--mesh", default="soup", choices=["soup", "grid"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--sah-treelets", action="store_true",
                    help="EXPERIMENTAL: build with the SAH treelet pass (treelet_sah.cuh; not the default, not yet validated on hardware)")
    ap.add_argument("--kernel", default="persistent", choices=["persistent", "simple"])
    ap.add_argument("--chunks", type=int, default=4, help="NCCL gather chunks per step when N > 1")
    ap.add_argument("--gather", default="auto", choices=["auto", "multicast", "peer", "nccl"],
                    help="N > 1: fused peer-memory gather (multicast / peer stores) or NCCL all-gather")
    return ap.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons during the timed region (NVML; nvidia-smi as a fallback)."""

    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag = index, threading.Event()
        self.sm, self.sm_max, self.reasons = [], None, set()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def sample(self):
        if self.nvml is not None:
            n = self.nvml
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
            try:
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
            except Exception:
                mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
            for name, bit in self.REASONS:
                if mask & bit:
                    self.reasons.add(name)
            return
        out = subprocess.run(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        f = [x.strip() for x in out.split(",")]
        self.sm.append(float(f[0]))
        self.sm_max = float(f[1])
        for (name, _), v in zip(self.REASONS, f[2:6]):
            if v.lower().startswith("active"):
                self.reasons.add(name)

    def run(self):
        while not self.stop_flag.is_set():
            try:
                self.sample()
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": ["no samples"], "samples": 0}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(sm)}


def camera_rays(kind, rank=0, world=1, rows=None):
    """This rank's shard: the same 3163^2 camera with a sub-pixel offset of rank/world, so that the
    global batch is world x 10M distinct primary rays of one distribution."""
    kw = dict(pixel_offset=rank / world)
    if rows is not None:
        kw.update(y_begin=rows[0], y_end=rows[1])
    return scenes.make_primary(kind, IMG, IMG, **kw)


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU path on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_arm(kind, steps, warmup, quick=False):
    """Times DefaultBuilder (Quality::High, the library default) and Bvh::intersect over a bounded sample
    with all host threads.  Only this function (and tests/, smoke()) touches oracle/."""
    from oracle.pyoracle import TIE_LOWEST_ID, Oracle, Ref, ref_available
    tris = scenes.make_mesh(kind, TRIS)
    r0 = (IMG - CPU_SAMPLE_ROWS) // 2
    rays = camera_rays(kind, rows=(r0, r0 + CPU_SAMPLE_ROWS))
    sample = f"{rays.shape[0]} rays = image rows {r0}..{r0 + CPU_SAMPLE_ROWS} of the {IMG}x{IMG} camera, {tris.shape[0]} triangles"
    if ref_available():
        ref = Ref()
        cores = ref.thread_count(0)
        bb, cc = ref.tri_bboxes_centers(tris)
        t0 = time.perf_counter()
        tree = ref.build(bb, cc, quality="high", threads=0)
        build_s = time.perf_counter() - t0
        build_low_s = ref.time_build(bb, cc, quality="low", threads=0)
        ref.set_triangles(tree, tris)
        times = []
        for i in range(warmup + steps):
            ref.trace(tree, rays, flags=TIE_LOWEST_ID, threads=0, outputs=False)
            if i >= warmup:
                times.append(ref.last_trace_seconds)
        kind_ = "reference"
        ref.trace(tree, rays[:100000], flags=TIE_LOWEST_ID, threads=-1, outputs=False)
        single = 100000 / ref.last_trace_seconds / 1e6
    else:
        orc = Oracle()
        cores = 1
        bb, cc = orc.tri_bboxes_centers(tris)
        t0 = time.perf_counter()
        tree = orc.build(bb, cc, quality="low")
        build_s = build_low_s = time.perf_counter() - t0
        orc.set_triangles(tree, tris)
        rays = rays[: rays.shape[0] // 8]
        sample = f"{rays.shape[0]} rays (scalar port), {tris.shape[0]} triangles"
        times = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            orc.trace(tree, rays, flags=TIE_LOWEST_ID)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        kind_ = "port"
        single = rays.shape[0] / np.median(times) / 1e6
    mean_s = float(np.mean(times))
    return {"value": rays.shape[0] / mean_s / 1e6, "unit": "Mrays/s", "cores": cores, "kind": kind_, "sample": sample,
            "ms_per_step": mean_s * 1e3, "single_thread_mrays": single,
            "build_high_mtris": tris.shape[0] / build_s / 1e6, "build_low_mtris": tris.shape[0] / build_low_s / 1e6}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    kind = args.mesh
    if args.sah_treelets:
        os.environ["BVH_B200_SAH_TREELETS"] = "1"
    workload = f"{kind}-1M triangles, {IMG}x{IMG} = {IMG * IMG} coherent primary rays per GPU, closest-hit"

    if args.impl == "reference":
        if rank != 0:
            return
        res = cpu_arm(kind, args.steps, args.warmup)
        line = {"impl": "reference", "metric": "primary closest-hit rays per second", "value": res["value"], "unit": "Mrays/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "bvh": "reference DefaultBuilder Quality::High", "sample": res["sample"]},
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "build": {"high_mtris_per_s": res["build_high_mtris"], "low_mtris_per_s": res["build_low_mtris"]},
                "e2e": {"value": res["value"], "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    import torch
    import torch.distributed as dist
    import bvh_b200.api as api
    from bvh_b200.multi_gpu import ShardedTracer

    if not torch.cuda.is_available() or api.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    api.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    stream = torch.cuda.Stream(device=device)
    torch.cuda.set_stream(stream)                 # everything below runs on this (non-default) stream,
    api.set_stream(stream.cuda_stream)            # library kernels included: CUDA events see them

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- inputs -----------------------------------------------------------------------------------
    tris_np = scenes.make_mesh(kind, TRIS)
    n_tris = tris_np.shape[0]
    verts = torch.from_numpy(tris_np).to(device)
    rays_np = camera_rays(kind, rank, world)
    n_rays = rays_np.shape[0]
    rays_pinned = torch.from_numpy(rays_np).pin_memory()
    rays = rays_pinned.to(device, non_blocking=True)
    torch.cuda.synchronize()
    peak_gbs, peak_src = hbm_peak()
    kflag = api.KERNEL_SIMPLE if args.kernel == "simple" else 0

    # ---- build: K timed LBVH builds from device-resident vertices ----------------------------------
    def build():
        return api.Bvh.build_triangles(verts.data_ptr(), count=n_tris, dtype=np.float32, flags=api.DEVICE_POINTERS)

    for _ in range(args.warmup):
        build().destroy()
    barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for _ in range(args.steps):
        build().destroy()
    b1.record()
    barrier()
    build_ms = max_over_ranks(b0.elapsed_time(b1) / args.steps)
    bvh = build()
    t0 = time.perf_counter()
    for _ in range(3):
        api.Bvh.build_triangles(tris_np).destroy()
    build_e2e_ms = (time.perf_counter() - t0) / 3 * 1e3
    build_bytes = 300.0                            # 32-bit Morton pipeline, bytes per triangle
    build_info = {"value": n_tris / build_ms / 1e3, "unit": "Mtris/s", "ms": build_ms, "scope": "per GPU (BVH replicated)",
                  "e2e_ms_host_vertices": build_e2e_ms, "depth": bvh.depth,
                  "roofline": {"bound": "hbm", "achieved": build_bytes * n_tris / (build_ms * 1e-3) / 1e9, "peak": peak_gbs,
                               "unit": "GB/s", "frac": build_bytes * n_tris / (build_ms * 1e-3) / 1e9 / peak_gbs,
                               "traffic": None, "bytes_per_tri": build_bytes}}

    # ---- algorithmic bytes of the traversal, measured on the tree actually traversed ---------------
    stats_sample = rays[: min(n_rays, 2_000_000)].contiguous()
    st = torch.empty((stats_sample.shape[0], 3), dtype=torch.int32, device=device)
    scratch_hits = torch.empty((stats_sample.shape[0], 4), dtype=torch.int32, device=device)
    if api.lib().bvh3f_intersect_rays_stats(bvh.handle, stats_sample.data_ptr(), stats_sample.shape[0], scratch_hits.data_ptr(),
                                            st.data_ptr(), api.DEVICE_POINTERS):
        raise SystemExit(api.last_error())
    torch.cuda.synchronize()
    s_inner, s_leaves, s_tri = (float(x) for x in st.double().mean(dim=0).tolist())
    hit_frac = float((scratch_hits[:, 0] != -1).double().mean().item())
    bytes_per_ray = 32 + 16 + 64 * s_inner + 48 * s_tri
    del st, scratch_hits, stats_sample

    # ---- traversal: W warm-up + K timed steps -------------------------------------------------------
    def trace(b, e, out):
        if api.lib().bvh3f_intersect_rays(bvh.handle, rays.data_ptr() + 32 * b, e - b, out.data_ptr(), api.DEVICE_POINTERS | kflag):
            raise SystemExit(api.last_error())

    tracer, gather_desc = None, None
    if world > 1 and args.gather != "nccl":
        try:
            from bvh_b200.multi_gpu import FusedGatherTracer
            tracer = FusedGatherTracer(bvh, rays, 4, flags=api.DEVICE_POINTERS | kflag, mode=args.gather)
            gather_desc = (f"fused in the traversal kernel: every hit record stored into all {world} ranks' symmetric-memory "
                           f"buffers ({'one multimem store via the NVSwitch multicast address' if tracer.mode == 'multicast' else 'one NVLink peer store per rank'}), "
                           "symmetric-memory barrier per step")
        except Exception as exc:                 # no symmetric memory on this box: NCCL all-gather instead
            if args.gather != "auto":
                raise
            if rank == 0:
                print(f"[bench] fused gather unavailable ({type(exc).__name__}: {exc}); using NCCL", file=sys.stderr)
            tracer = None
    if tracer is None:
        tracer = ShardedTracer(n_rays, 4, torch.int32, device, trace, chunks=args.chunks if world > 1 else 1)
        if world > 1:
            gather_desc = f"NCCL all_gather_into_tensor of hit records, {args.chunks} chunks overlapped with traversal"
    for _ in range(max(3, args.warmup)):
        tracer.step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    starts = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ends = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(args.steps):
        starts[k].record()
        tracer.step()
        ends[k].record()
    e1.record()
    barrier()
    clocks = sampler.summary()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = total_ms / args.steps
    kernel_ms = float(np.mean([s.elapsed_time(e) for s, e in zip(starts, ends)])) if world == 1 else None
    value = world * n_rays / ms_per_step / 1e3                     # Mrays/s, whole job

    hits_np = tracer.local.cpu().numpy().view(api.HIT3F).reshape(-1)
    checksum = int(hits_np["prim_id"].astype(np.uint64).sum())
    if world > 1:
        g = tracer.global_hits()
        assert g.shape[0] == world * n_rays
        assert torch.equal(g[rank * n_rays:(rank + 1) * n_rays], tracer.local), "gathered hits do not match the local shard"
        # every rank must hold every other rank's shard: compare per-shard checksums across ranks
        sums = g.view(world, n_rays, 4)[:, :, 0].to(torch.int64).sum(dim=1)
        ref_sums = sums.clone()
        dist.broadcast(ref_sums, src=0)
        assert torch.equal(sums, ref_sums), "ranks disagree on the gathered hit records"

    # ---- e2e through the C ABI with pinned host buffers ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        hits_pinned = torch.empty((n_rays, 4), dtype=torch.int32).pin_memory()
        def call():
            if api.lib().bvh3f_intersect_rays(bvh.handle, rays_pinned.data_ptr(), n_rays, hits_pinned.data_ptr(), kflag):
                raise SystemExit(api.last_error())
        for _ in range(2):
            call()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            call()
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) / args.steps * 1e3)
        assert np.array_equal(hits_pinned.numpy().view(api.HIT3F).reshape(-1), hits_np), "e2e hits differ from device-resident hits"
        e2e = {"value": world * n_rays / e2e_ms / 1e3, "unit": "Mrays/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(n_rays * 32), "d2h_bytes_per_step": int(n_rays * 16),
               "api": "bvh3f_intersect_rays(host rays, host hits) with pinned buffers"}

    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        res = cpu_arm(kind, 3, 1)
        cpu = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
        cpu.update(single_thread_mrays=res["single_thread_mrays"], build_high_mtris=res["build_high_mtris"],
                   build_low_mtris=res["build_low_mtris"])

    if rank == 0:
        dur_ms = kernel_ms if kernel_ms is not None else ms_per_step
        achieved = bytes_per_ray * n_rays / (dur_ms * 1e-3) / 1e9
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"trace_{args.kernel}_{kind}")
        except Exception:
            pass
        line = {
            "metric": "primary closest-hit rays per second", "value": value, "unit": "Mrays/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "bvh": "LBVH (30-bit Morton, SAH leaf collapse, max_leaf_size 8)" + (" + experimental SAH treelet pass" if args.sah_treelets else "")
                              + " built on the GPU, replicated per rank",
                       "kernel": args.kernel, "tie_break": "lowest original id (canonical)",
                       "l2": "inputs larger than L2: 320 MB of rays + 160 MB of hits streamed per step, no flush needed",
                       "hit_fraction": hit_frac, "inner_steps_per_ray": s_inner, "leaves_per_ray": s_leaves, "tri_tests_per_ray": s_tri,
                       "gather": gather_desc,
                       "hits_checksum": checksum},
            "roofline": {"bound": "hbm", "kernel": f"trace_{args.kernel}_kernel<float>", "achieved": achieved, "peak": peak_gbs,
                         "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": traffic,
                         "bytes_per_ray": bytes_per_ray, "launch_ms": dur_ms},
            "build": build_info,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": args.steps * len(tracer.bounds),
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
