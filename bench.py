#!/usr/bin/env python
"""bench.py -- ray-bounce segments/s of the batched acoustic hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c5] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one frame of the hot path: AudioRaytracerJobBatched + AudioPermeationJobBatched +
ProcessAudioDataJob over one synthetic batch (BASELINE.json configs). The default workload is C3
(64 sources, 1M rays x 12 bounces, 4,096 colliders) -- the configuration BASELINE.json quotes "at 1/2/4/8 B200";
with N > 1 the SAME batch is ray-sharded over the ranks (strong scaling), no data-path collective, and only the
per-source partial results (a few KB) are all-gathered over NCCL.

Prints ONE JSON line (rank 0). `value` = segments/s with scene and rays resident in HBM, device-timed (CUDA events
on the library's stream, max over ranks). `e2e` = the same metric through the C ABI with HOST buffers: scene + ray
upload and all per-ray outputs copied back inside the timed region (wall clock around schedule -> complete).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic flops per primitive test; index order sphere, AABB, OBB
FLOPS_TRACE = (38, 35, 98)      # trace / echo / muffle tests (RT)
FLOPS_PERM_FIRST = (38, 35, 98)  # PM first-hit tests
FLOPS_PERM_LOSS = (29, 38, 101)  # PM loss tests
METRIC = "ray-bounce segments/sec (device-timed)"
UNIT = "segments/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=["c1", "c1_1src", "c2", "c3", "c4", "c5"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=None, help="override the ray count (debug only; marks the line)")
    ap.add_argument("--chunk", type=int, default=0, help="rays per interleaved shard chunk (N > 1); 0 = the library's choice, 8 chunks per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, device_index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.period = period_s
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def run(self):
        if self.h is None:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._halt.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        self.join(timeout=2)
        med = int(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def flops_of(counters: dict):
    """Algorithmic flops of a frame from its oracle-equivalent test counters."""
    g = lambda k: counters.get(k) or counters.get({"traceTests": "trace_tests", "echoTests": "echo_tests",
                                                   "muffleTests": "muffle_tests", "permFirstTests": "perm_first_tests",
                                                   "permLossTests": "perm_loss_tests"}[k])
    trace = sum(f * (a + b + c) for f, a, b, c in zip(FLOPS_TRACE, g("traceTests"), g("echoTests"), g("muffleTests")))
    perm = sum(f * a for f, a in zip(FLOPS_PERM_FIRST, g("permFirstTests"))) + \
        sum(f * a for f, a in zip(FLOPS_PERM_LOSS, g("permLossTests")))
    return trace, perm


# rays per window of the CPU sample (4 windows at 10/35/60/85 % of the Fibonacci index), fixed per workload so that the
# `cpu_baseline` leg of this arm and `--impl reference` time exactly the same work (about 2-4 s on 16 host threads)
CPU_SAMPLE_WINDOW = {"c1": None, "c1_1src": None, "c2": None, "c3": 512, "c4": 256, "c5": 384}


def cpu_sample_windows(workload: str, scene):
    N = scene.n_rays
    win = CPU_SAMPLE_WINDOW.get(workload)
    if win is None or 4 * win >= N:
        return [(0, N)], f"all {N} rays, RT+PM jobs, full scene and targets"
    starts = [int(f * (N - win)) for f in (0.1, 0.35, 0.6, 0.85)]
    return [(s0, win) for s0 in starts], (f"{4 * win} of {N} rays (4 windows of {win} at 10/35/60/85% of the Fibonacci index), "
                                           "RT+PM jobs, full scene and targets")


def cpu_reference_sample(scene, workload: str, threads: int, shrink: int = 1):
    """The reference algorithm on the host cores (oracle port, RT + PM) over the workload's fixed ray sample (`shrink`:
    only every shrink-th part of each window, for the single-thread figure). Returns (segments/s, description, seconds)."""
    from oracle import oracle as orc
    windows, desc = cpu_sample_windows(workload, scene)
    seg, t = 0, 0.0
    for s0, cnt in windows:
        cnt = max(1, cnt // shrink)
        t0 = time.perf_counter()
        c = orc.trace_range(scene, s0, cnt, threads=threads, with_outputs=False)
        orc.permeation_range(scene, s0, cnt, threads=threads)
        t += time.perf_counter() - t0
        seg += c["segments"]
    if shrink > 1:
        desc = f"the first 1/{shrink} of each window of: " + desc
    return seg / t, desc, t


def run_reference(args, scene, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    vals = []
    desc = ""
    for i in range(args.warmup + args.steps):
        v, desc, t = cpu_reference_sample(scene, args.workload, threads)
        if i >= args.warmup:
            vals.append((v, t))
    value = float(np.mean([v for v, _ in vals]))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean([t for _, t in vals]) * 1e3), "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, scene, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc,
                         "values_per_step": [v for v, _ in vals]},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is C#/Unity and cannot run here or on the GPU box (no dotnet / mono: profiles/r02_dotnet_probe.txt); "
                "this is the C restatement oracle/audiort_oracle.c of its Execute() bodies (brute-force scans, as the reference "
                "loops do) on all host threads, one step = the workload's fixed ray sample",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, scene, world):
    return {"workload": f"{args.workload}: {scene.n_targets} sources, {scene.n_rays} rays x {scene.max_hits_per_ray} hits, "
                        f"{len(scene.aabbs)} AABB + {len(scene.obbs)} OBB + {len(scene.spheres)} spheres",
            "rays": scene.n_rays, "max_hits_per_ray": scene.max_hits_per_ray, "targets": scene.n_targets,
            "colliders": scene.n_colliders, "batch_count": scene.batch_count,
            "sharding": "none" if world == 1 else f"rays interleaved in chunks of {args.chunk or 'N / (8 x ranks)'} over {world} ranks; per-source partials "
                                                  "all-gathered on the device inside the library (ncclAllGather), merged and finalised on every rank",
            "l2": "flushed between steps (256 MiB write)", "reverb_reduction": "exact integer",
            "path": "default: bounce-only tracer on the uniform grid, then one query kernel over all hit points using per-frame target "
                    "fans (direction-binned collider lists around the listener and every source; fan build inside the timed region "
                    "every step), permeation loss lines sorted by (source, direction bin) on the device every frame; "
                    "bit-identical to the brute-force scans (ART_FRAME_BRUTE_FORCE timed beside it)",
            "rays_override": args.rays is not None}


def main():
    args = parse_args()
    rank, world, local = dist_env()
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner to fd 1) write to stderr instead
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    from audio_raytracer_b200 import scenes
    scene = scenes.make_config(args.workload, batch_count=max(1, args.gpus), n_rays=args.rays)

    if args.impl == "reference":
        run_reference(args, scene, rank, world)
        return

    import torch
    import torch.distributed as dist
    from audio_raytracer_b200 import build, native
    # one rank builds (a stale library would otherwise be rebuilt by N processes into the same files), the others wait
    lock = os.path.join(ROOT, "audio-raytracer_b200", ".build.lock")
    import fcntl
    with open(lock, "w") as lf:
        fcntl.flock(lf, fcntl.LOCK_EX)
        try:
            build.build()
        finally:
            fcntl.flock(lf, fcntl.LOCK_UN)
    torch.cuda.set_device(local)
    # keep stdout to the one JSON line: NCCL prints its version banner to stdout at WARN and above
    if os.environ.get("ART_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = os.environ["ART_NCCL_DEBUG"]
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

    ctx = native.Context(device=local)
    if world > 1:
        # the library exchanges the per-source partial blobs itself (device-resident ncclAllGather on its own stream) and
        # finalises on every rank; torch.distributed only carries the 128-byte communicator id and the timing reductions
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid = torch.frombuffer(bytearray(native.comm_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(uid, 0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()), rank, world, args.chunk)
    native.upload(ctx, scene)
    n_local = ctx.local_ray_count()
    H, Na, T = scene.max_hits_per_ray, scene.n_targets, scene.batch_count
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    blob_bytes = int(native.load_library().art_partials_size(Na, T))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def device_step(flags=0):
        flush.zero_()
        torch.cuda.synchronize()
        r = ctx.run_frame(scene, flags=flags | native.FRAME_NO_HOST_OUTPUTS, want=())
        return r.counters

    # ---- untimed passes: (1) oracle-equivalent test counts of the reference's full scans -> reference-formulation flops
    #      of a step; (2) the brute-force kernels without counters -> their own roofline; (3) the tests the default
    #      (uniform-grid) kernels actually execute
    barrier()
    cnt = device_step(native.FRAME_COUNTERS)
    trace_flops_local, perm_flops_local = flops_of(cnt)
    device_step(native.FRAME_BRUTE_FORCE)
    bf = device_step(native.FRAME_BRUTE_FORCE)
    gst = device_step(native.FRAME_GRID_STATS)
    grid_used = int(gst.get("gridUsed", 0))
    exec_trace_flops = sum(f * n for f, n in zip(FLOPS_TRACE, gst["gridTraceTests"]))
    exec_perm_flops = sum(f * n for f, n in zip(FLOPS_PERM_FIRST, gst["gridPermFirstTests"])) + \
        sum(f * n for f, n in zip(FLOPS_PERM_LOSS, gst["gridPermLossTests"]))

    for _ in range(args.warmup):
        device_step()

    # ---- timed: device-resident inputs, CUDA events on the library stream
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    wall0 = time.perf_counter()
    dev_ms, trace_ms, perm_ms, reduce_ms, ex_ms_tot, launches, segs = [], [], [], [], 0.0, 0, 0
    fan_ms, bounce_ms, query_ms = [], [], []
    timed_grid_used = 0
    for _ in range(args.steps):
        c = device_step()
        dev_ms.append(c["deviceMs"]); trace_ms.append(c["traceMs"]); perm_ms.append(c["permeationMs"]); reduce_ms.append(c["reduceMs"])
        fan_ms.append(c["fanBuildMs"]); bounce_ms.append(c["bounceMs"]); query_ms.append(c["queryMs"])
        ex_ms_tot += c["exchangeMs"]
        launches += c["kernelLaunches"]
        segs += c["segments"]
        timed_grid_used = int(c.get("gridUsed", 0))
    barrier()
    wall_dev = time.perf_counter() - wall0
    clocks = sampler.stop()

    # ---- e2e: host buffers through the C ABI, uploads and read-backs inside the timed region
    out = native.FrameResult()
    for _ in range(max(1, min(args.warmup, 2))):
        native.upload(ctx, scene)
        ctx.run_frame(scene, result=out, want=("echo", "hit_points", "hit_counts"))
    barrier()
    t0 = time.perf_counter()
    e2e_segs = 0
    for _ in range(args.steps):
        native.upload(ctx, scene)                                  # H2D: scene structs + ray directions (+ targets)
        r = ctx.run_frame(scene, result=out, want=("echo", "hit_points", "hit_counts"))   # (world > 1: merged + finalised in the library)
        e2e_segs += r.counters["segments"]
    barrier()
    e2e_s = time.perf_counter() - t0
    h2d = n_local * 6 + len(scene.aabbs) * 20 + len(scene.obbs) * 26 + len(scene.spheres) * 16 + Na * 12
    d2h = n_local * H * (2 + 6) + n_local + blob_bytes

    # ---- reduce over ranks
    def allmax(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        if world == 1:
            return float(x)
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    step_ms = allmax(sum(dev_ms) + ex_ms_tot) / args.steps
    total_segs = allsum(segs)
    value = total_segs / args.steps / (step_ms * 1e-3)
    e2e_value = allsum(e2e_segs) / allmax(e2e_s)
    trace_ms_avg = float(np.mean(trace_ms))
    trace_flops_all = allsum(trace_flops_local)
    perm_flops_all = allsum(perm_flops_local)
    launches_all = int(allsum(launches))
    wall_dev_max = allmax(wall_dev)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        n_sms = 148
        peak_tflops = n_sms * 128 * sm_max * 1e6 / 1e12          # FP32 instruction issue, no FMA
        perm_ms_avg = float(np.mean(perm_ms))
        tfl = lambda flops, ms: (flops / (ms * 1e-3) / 1e12) if ms and ms > 0 else None
        fan_ms_avg, bounce_ms_avg, query_ms_avg = float(np.mean(fan_ms)), float(np.mean(bounce_ms)), float(np.mean(query_ms))
        q_tests = [int(x) for x in gst.get("gridQueryTests", [0, 0, 0])]
        b_tests = [int(t) - q for t, q in zip(gst["gridTraceTests"], q_tests)]
        q_flops = sum(f * n for f, n in zip(FLOPS_TRACE, q_tests))
        b_flops = sum(f * n for f, n in zip(FLOPS_TRACE, b_tests))
        split = query_ms_avg > 0
        if grid_used & 1 and split:
            # the dominant kernel of the step: the echo / muffle queries of all hit points (k1_query_fan.cu)
            achieved = tfl(q_flops, query_ms_avg)
            kernel_name = "query_fan_kernel (K1b: echo + muffle queries of every hit point against the target fans)"
            accounting = ("collider tests the kernel actually executed (ART_FRAME_GRID_STATS) x reference flops per test; the "
                          "acceleration structures change WHICH tests run (about 1.2 per occlusion query instead of the reference's "
                          "scan: three quarters of the queries are decided blocked by the covering-depth compare of their direction bin, "
                          "without a collider test), so the full-scan count is reported beside it (SURVEY 8f-4). The fraction is low by "
                          "design: a query costs a direction-bin lookup, and a surviving one an exact sqrt + four exact reciprocals, "
                          "before its first test; "
                          "profiles/ holds issue-slot and lane utilisation. 'kernels' lists the bounce tracer and K2 as well.")
        elif grid_used & 1:
            achieved = tfl(exec_trace_flops, trace_ms_avg)
            kernel_name = "trace_grid_kernel (K1, grid walk: bounce rays and occlusion queries in one kernel)"
            accounting = "collider tests the kernel actually executed (ART_FRAME_GRID_STATS) x reference flops per test"
        else:
            achieved = tfl(trace_flops_local, trace_ms_avg)
            kernel_name = "trace_kernel (K1, brute force)"
            accounting = "tests of the reference's full scans (early exits honoured) x reference flops per test"
        # dram bytes per launch and issue-slot / lane utilisation come from ONE committed ncu --set full capture; they are only
        # reported when that capture was taken from exactly these kernel sources
        traffic, ncu_util, traffic_note = None, None, None
        try:
            import hashlib
            h = hashlib.sha1()
            cs = os.path.join(ROOT, "audio-raytracer_b200", "csrc")
            for f in sorted(os.listdir(cs)):
                h.update(open(os.path.join(cs, f), "rb").read())
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            ent = tr.get(f"{args.workload}:{scene.n_rays}:{world}")
            if ent and ent.get("csrc_sha1") == h.hexdigest():
                ncu_util = ent.get("ncu")
                by_kernel = ent.get("traffic") or {}
                traffic = by_kernel.get("query_fan_kernel" if split else "trace_grid_kernel")     # the dominant kernel's, bytes per launch
                traffic_note = f"dram bytes per launch by kernel: {by_kernel} ({ent.get('source')})"
            elif ent:
                traffic_note = f"profiles/traffic.json was captured from other kernel sources ({ent.get('csrc_sha1', '?')[:10]} != {h.hexdigest()[:10]}): not reported"
        except Exception:
            pass
        micro = {}
        try:
            micro = {"fadd_fmul_tops": ctx.microbench(0) / 1e3, "fmnmx_tops": ctx.microbench(1) / 1e3,
                     "ffma_tops": ctx.microbench(2) / 1e3}
        except Exception as ex:  # pragma: no cover
            micro = {"error": str(ex)}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            cpu_reference_sample(scene, args.workload, threads, shrink=8)          # warm-up (page in the scene, spin up the cores)
            v, desc, t = cpu_reference_sample(scene, args.workload, threads)
            v1, desc1, t1 = cpu_reference_sample(scene, args.workload, 1, shrink=8)
            cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc, "seconds": t,
                   "single_thread": {"value": v1, "sample": desc1, "seconds": t1},
                   "gpu_brute_force_over_cpu": (bf["segments"] / (bf["deviceMs"] * 1e-3) / v) if bf["deviceMs"] else None,
                   "gpu_default_over_cpu": value / v,
                   "note": "C port of the reference's brute-force loops. gpu_brute_force_over_cpu compares like with like (same "
                           "tests executed); the rest of gpu_default_over_cpu is algorithmic (grid + target fans)."}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(args, scene, world),
            "segments_per_step": total_segs / args.steps,
            "kernel_ms": {"trace": trace_ms_avg, "trace_fan_build": fan_ms_avg, "trace_bounce": bounce_ms_avg, "trace_queries": query_ms_avg,
                          "permeation": float(np.mean(perm_ms)), "reduce": float(np.mean(reduce_ms)),
                          "partials_allgather": ex_ms_tot / args.steps,
                          "note": "trace = per-frame fan build (fan_order_kernel + fan_project_kernel + fan_match_kernel) + bounce tracer (bounce_kernel) "
                                  "+ query_fan_kernel; partials_allgather = ncclAllGather of the per-source blobs inside the library"},
            "wall_ms_per_step_device_mode": wall_dev_max / args.steps * 1e3,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": allmax(e2e_s) / args.steps * 1e3 if world == 1 else None},
            "gpu_launches": launches_all,
            "roofline": {"bound": "fp32", "kernel": kernel_name, "achieved": achieved, "peak": peak_tflops,
                         "unit": "TFLOP/s", "frac": (achieved / peak_tflops) if achieved else None, "traffic": traffic,
                         "traffic_note": traffic_note,
                         "accounting": accounting,
                         "kernels": {
                             "query_fan_kernel": {"ms": query_ms_avg, "tests_executed": q_tests, "executed_flops": q_flops,
                                                  "achieved": tfl(q_flops, query_ms_avg), "frac": (tfl(q_flops, query_ms_avg) or 0) / peak_tflops},
                             "bounce_kernel": {"ms": bounce_ms_avg, "tests_executed": b_tests, "executed_flops": b_flops,
                                                               "achieved": tfl(b_flops, bounce_ms_avg), "frac": (tfl(b_flops, bounce_ms_avg) or 0) / peak_tflops},
                             "permeation (K2)": {"ms": perm_ms_avg, "executed_flops": exec_perm_flops,
                                                 "achieved": tfl(exec_perm_flops, perm_ms_avg), "frac": (tfl(exec_perm_flops, perm_ms_avg) or 0) / peak_tflops}},
                         "peak_source": f"148 SM x 128 FP32 lanes x {sm_max:.0f} MHz, un-fused issue rate (no FP32 figure in "
                                        "MEASURED_PEAKS.json; parity forbids FMA contraction, SURVEY 8d); FMA peak = 2x; "
                                        "on-box microbenchmarks in 'measured_issue_rates'",
                         "peak_fma": 2 * peak_tflops,
                         "ncu_utilisation": ncu_util,
                         "executed_flops_per_launch": (q_flops if split else exec_trace_flops) if grid_used & 1 else trace_flops_local,
                         "flops_per_test": {"sphere": 38, "aabb": 35, "obb": 98},
                         "measured_issue_rates": micro,
                         "reference_formulation": {
                             "what": "the reference's own full linear scans (early exits honoured), counted by the brute-force kernels' COUNT variant",
                             "algorithmic_flops_per_launch": trace_flops_local,
                             "equivalent_tflops_of_default_path": tfl(trace_flops_local, trace_ms_avg),
                             "tests_full_scan": [int(a + b + c) for a, b, c in zip(cnt["traceTests"], cnt["echoTests"], cnt["muffleTests"])],
                             "tests_executed": [int(x) for x in gst["gridTraceTests"]],
                             "cells_visited": int(gst["gridTraceCells"])},
                         "brute_force_kernel": {
                             "kernel": "trace_kernel (ART_FRAME_BRUTE_FORCE: every collider for every query, as the reference loops do)",
                             "ms": bf["traceMs"], "achieved": tfl(trace_flops_local, bf["traceMs"]),
                             "frac": (tfl(trace_flops_local, bf["traceMs"]) or 0) / peak_tflops,
                             "step_ms": bf["deviceMs"], "segments_per_s": bf["segments"] / (bf["deviceMs"] * 1e-3) if bf["deviceMs"] else None},
                         "permeation_kernel": {
                             "kernel": ("perm_loss_binned_kernel + permeation_grid_kernel (K2: first hits, then the loss lines counting-"
                                        "sorted by (target, direction bin) and evaluated 32 of one bin at a time)" if timed_grid_used & 32
                                        else "permeation_grid_kernel (K2)" if grid_used & 2 else "permeation_kernel (K2, brute force)"),
                             "achieved": tfl(exec_perm_flops if grid_used & 2 else perm_flops_local, perm_ms_avg),
                             "executed_flops_per_launch": exec_perm_flops if grid_used & 2 else perm_flops_local,
                             "algorithmic_flops_per_launch": perm_flops_local,
                             "equivalent_tflops_of_default_path": tfl(perm_flops_local, perm_ms_avg),
                             "brute_force_ms": bf["permeationMs"], "brute_force_achieved": tfl(perm_flops_local, bf["permeationMs"])}},
            "cpu_baseline": cpu,
            "counters_first_step": {k: cnt[k] for k in ("segments", "segmentHits", "traceTests", "echoTests", "muffleTests",
                                                        "echoQueries", "muffleQueries", "permHitRays")},
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
