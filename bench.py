#!/usr/bin/env python
"""bench.py -- device-timed PLOC+CWBVH build (Mtris/s) and CWBVH closest-hit traversal (Mrays/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload s3|kitchen|soup|terrain|bounce|demoscene|dynamic]

Default workload `s3` = BASELINE.json configs[3] with configs[1] beside it in the same JSON line:
  * top level: the synthetic 10 M-triangle scene demoscene(2237, 0), `fast_build` on ONE GPU, the finished CWBVH broadcast with NCCL
    (obvhs_cuda_cwbvh_broadcast), 100 320 000 incoherent diffuse-bounce rays (examples/demoscene.rs:126-178, generated on the
    device) sharded over the N GPUs in contiguous ranges -- STRONG scaling: the ray set is fixed, per-GPU work shrinks with N;
  * "kitchen": kitchen.obj, `fast_build`, 1920x1080 primary rays (examples/obj_cwbvh.rs), sharded the same way.
One "step" = one pass of the hot path over one batch: build_cwbvh_from_tris (PLOC -> reinsertion -> CWBVH collapse) on rank 0,
broadcast, closest-hit traversal of this rank's rays, all through the C ABI with inputs resident in HBM. `value` is traversal
Mrays/s (all ranks' rays / max-over-ranks device time), `build` carries the Mtris/s half of the metric, `e2e` the same call with
pinned HOST buffers. `parity` compares the run's own results with the CPU oracle on a sample and across ranks.
`--impl reference` times the CPU restatement of the reference (oracle/, OpenMP on all host cores; the rustc/rayon binary cannot
be produced in this image) on a bounded sample of the same workloads.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "CWBVH closest-hit traversal Mrays/s (PLOC+CWBVH build Mtris/s in `build`)"


def make_workload(name: str, n_tris: int, seed: int = 0):
    """-> (tris (n,12) f32, rays (m,16) f32, description, preset)"""
    from obvhs_b200 import camera, test_util as tu

    if name == "kitchen":
        # BASELINE.json configs[1]: assets/kitchen.obj, fast_build (examples/obj_cwbvh.rs:47), 1920x1080 primary rays
        tris = tu.kitchen()
        rays = camera.primary_rays(camera.kitchen_camera(1920))
        return tris, rays, "kitchen.obj 56939 tris, fast_build, 1920x1080 primary rays (examples/obj_cwbvh.rs camera)", "fast_build"
    if name == "terrain":
        res = int(round((n_tris / 2) ** 0.5))
        tris = cached_demoscene(res)
        cam = camera.demoscene_camera(1920)
        rays = camera.demoscene_primary(cam, 0)
        return tris, rays, f"demoscene({res},0) {tris.shape[0]} tris, fast_build, {cam.width}x{cam.height} jittered primary rays", "fast_build"
    if name == "soup":
        tris = tu.triangle_soup(n_tris, seed)
        rng = np.random.default_rng(seed)
        m = 1920 * 1080
        o = rng.random((m, 3), dtype=np.float32) * np.float32(1.2) - np.float32(0.1)
        t = rng.random((m, 3), dtype=np.float32)
        d = t - o
        d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
        from obvhs_b200.types import make_rays

        rays = make_rays(o, d.astype(np.float32), 0.0, np.inf)
        return tris, rays, f"hashed triangle soup {n_tris} tris, fast_build, {m} incoherent rays", "fast_build"
    if name == "bounce":
        # BASELINE.json configs[3] / SURVEY.md 8(d) S3: 10M-triangle terrain, demoscene camera 1280x475, `samples` AA samples;
        # the timed set is the cosine-hemisphere bounce ray leaving every primary hit (examples/demoscene.rs:126-178).
        # rays=None: they are generated after the build, by tracing the primary rays with the implementation under test.
        res = int(round((n_tris / 2) ** 0.5))
        tris = cached_demoscene(res)
        return tris, None, f"demoscene({res},0) {tris.shape[0]} tris, fast_build, diffuse bounce rays of 1280x475 px", "fast_build"
    raise SystemExit(f"unknown workload {name}")


def bounce_samples(total_samples: int, rank: int, world: int):
    """AA samples traced by `rank` (strong scaling: the sample set is fixed, ranks take contiguous slices)."""
    from obvhs_b200.sharding import shard_range

    lo, hi = shard_range(total_samples, rank, world)
    return range(lo, hi)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line), through NVML in a
    thread of this process (the same calls `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons...` makes).
    OBVHS_CLOCK_SAMPLER=smi uses an nvidia-smi -lms 100 subprocess instead, =off disables sampling."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_s: float = 0.1):
        self.gpu, self.period = gpu_index, period_s
        self.mode = os.environ.get("OBVHS_CLOCK_SAMPLER", "nvml")
        self.sm, self.mx, self.reasons = [], [], set()
        self.proc = self.thread = None
        self.stop_flag = threading.Event()

    def start(self):
        try:
            if self.mode == "off":
                return
            if self.mode == "smi":
                self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._pump_smi, daemon=True)
            else:
                import pynvml

                pynvml.nvmlInit()
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[self.gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.gpu
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
                self.nv = pynvml
                self.thread = threading.Thread(target=self._pump_nvml, daemon=True)
            self.thread.start()
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
            self.thread = None

    def _pump_nvml(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mx.append(float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self.stop_flag.wait(self.period)

    def _pump_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.proc.stdout:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                self.sm.append(float(f[1]))
                self.mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    self.reasons.add(nm)

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampler unavailable: " + getattr(self, "err", self.mode)]}
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:  # noqa: BLE001
                self.proc.kill()
        else:
            if not self.sm:
                time.sleep(0.02)
            self.stop_flag.set()
        self.thread.join(timeout=2)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "how": self.mode}


def bind_to_gpu_numa_node(local_rank: int, world: int):
    """N > 1 only: run this rank (and so allocate its pinned staging buffers, first touch) on the CPUs NVML reports as local to its
    GPU. Eight ranks that float across both sockets pull half of their host batches over the inter-socket link. Returns a note for
    the JSON line. (At N = 1 the process keeps every core: the cpu_baseline leg runs there.)"""
    if world <= 1 or os.environ.get("OBVHS_BENCH_NUMA", "1") == "0":
        return None
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local_rank
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {w * 64 + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1} & os.sched_getaffinity(0)
        if not cpus:
            return "no GPU-local CPUs reported"
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} GPU-local CPUs"
    except Exception as e:  # noqa: BLE001  (a box without NVML affinity data just keeps the default placement)
        return f"unbound ({type(e).__name__})"


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_reference_leg(tris, rays, preset, steps, warmup):
    """The reference's CPU algorithm (C++ restatement in oracle/, OpenMP over all host cores) on the same workload."""
    import oracle_bind as ob

    # all host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the oracle takes an explicit thread count)
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    build_s, trav_s = [], []
    hits = None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        c = ob.build_cwbvh_from_tris(tris, preset, threads=threads)
        bt = c.bvh_tris(tris)
        t1 = time.perf_counter()
        hits = c.ray_traverse(bt, rays, threads=threads, use_simd=True)
        t2 = time.perf_counter()
        if it >= warmup:
            build_s.append(c.core_build_seconds)
            trav_s.append(t2 - t1)
    return {"threads": threads, "build_s": float(np.mean(build_s)), "trav_s": float(np.mean(trav_s)), "hits": hits}


def cpu_bounce_rays(tris, preset, n_samples):
    """Bounce set for the CPU legs: primary hits come from the CPU restatement (bit-identical to the GPU path's)."""
    import oracle_bind as ob
    from obvhs_b200 import camera

    threads = len(os.sched_getaffinity(0))
    c = ob.build_cwbvh_from_tris(tris, preset, threads=threads)
    bt = c.bvh_tris(tris)
    rays, _ = camera.demoscene_bounce_set(camera.demoscene_camera(1280), range(n_samples), bt,
                                          lambda r: c.ray_traverse(bt, r, threads=threads, use_simd=True))
    return rays


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tris, rays, desc, preset = make_workload(args.workload, args.tris)
    if rays is None:  # bounce set of a bounded number of AA samples, primary hits from the CPU path itself
        rays = cpu_bounce_rays(tris, preset, max(1, args.ref_rays // (1280 * 475)))
    # bounded sample so `--steps K --warmup W` ends within a few minutes on the host cores
    max_rays = args.ref_rays
    sample = rays[:: max(1, rays.shape[0] // max_rays)][:max_rays] if rays.shape[0] > max_rays else rays
    r = cpu_reference_leg(tris, sample, preset, args.steps, min(args.warmup, 1))
    mrays = sample.shape[0] / r["trav_s"] / 1e6
    mtris = tris.shape[0] / r["build_s"] / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": (r["trav_s"] + r["build_s"]) * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic" if args.workload != "kitchen" else "kitchen.obj fixture (reference asset)",
        "config": {"workload": desc, "preset": preset},
        "build": {"value": mtris, "unit": "Mtris/s", "ms": r["build_s"] * 1e3},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": r["threads"], "kind": "port",
                         "sample": f"{sample.shape[0]} of {rays.shape[0]} rays (strided), full build; C++ restatement of the obvhs CPU path, not the rustc/rayon binary",
                         "build_mtris_per_s": mtris},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# =====================================================================================================================
# Headline: S3 (10 M triangles, 100 M sharded bounce rays) + kitchen
# =====================================================================================================================
S3_RES = 2237             # demoscene(2237, 0) -> 10 008 338 triangles
S3_RAYS = 100_320_000     # 1280 x 475 px x 165 samples; here: the first 100 320 000 BOUNCE rays of the recipe


def cached_demoscene(res: int):
    """tu.demoscene(res, 0) costs ~25 s of numpy at 10 M triangles and the driver runs bench.py eight times on one box
    (N = 1, 2, 4, 8, both arms): keep the generated INPUT triangles in /tmp for the following runs."""
    from obvhs_b200 import test_util as tu

    path = os.path.join(os.environ.get("OBVHS_CACHE_DIR", "/tmp/obvhs_bench_cache"), f"demoscene_{res}_0.npy")
    try:
        a = np.load(path)
        if a.shape == (2 * res * res, 12) and a.dtype == np.float32:
            return a
    except Exception:  # noqa: BLE001
        pass
    a = tu.demoscene(res, 0)
    try:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = f"{path}.{os.getpid()}.tmp.npy"
        np.save(tmp, a)
        os.replace(tmp, path)
    except Exception:  # noqa: BLE001
        pass
    return a


def headline_config(args):
    """The `config` object: identical for both arms (the driver compares them)."""
    res = S3_RES if args.tris == 10_000_000 else int(round((args.tris / 2) ** 0.5))
    return {"workload": f"S3: demoscene({res},0) {2 * res * res} tris, fast_build on one GPU, {args.rays} incoherent diffuse-bounce rays "
                        f"(examples/demoscene.rs:126-178) sharded over the GPUs; kitchen.obj 56939 tris, fast_build, 1920x1080 primary rays in `kitchen`",
            "preset": "fast_build", "tris": 2 * res * res, "rays_total": args.rays,
            "l2": "S3 inputs (3.2 GB of rays) exceed L2; L2 also flushed between steps (256 MB write)",
            "multi_gpu": "build on rank 0, NCCL broadcast below the C ABI, rays sharded in contiguous ranges (strong scaling)"}, res


def hits_hash_np(hits) -> int:
    """Order-independent 64-bit checksum of a RayHit array: sum over rays of (t_bits << 32 | primitive_id) mod 2^64."""
    h = np.asarray(hits).view(np.uint32).reshape(-1, 4)
    with np.errstate(over="ignore"):
        return int(((h[:, 3].astype(np.uint64) << np.uint64(32)) | h[:, 0].astype(np.uint64)).sum(dtype=np.uint64))


def hits_hash_dev(d_hits) -> int:
    import torch

    v = (d_hits[:, 3].to(torch.int64) & 0xFFFFFFFF) << 32 | (d_hits[:, 0].to(torch.int64) & 0xFFFFFFFF)
    return int(v.sum().item()) & 0xFFFFFFFFFFFFFFFF


def bytes_hash_dev(ptr: int, nbytes: int, device: int) -> int:
    """64-bit word sum of a device buffer (nbytes rounded down to 8)."""
    import torch

    from obvhs_b200.sharding import device_bytes_tensor

    if not ptr or nbytes < 8:
        return 0
    t = device_bytes_tensor(ptr, nbytes - nbytes % 8, device).view(torch.int64)
    return int(t.sum().item()) & 0xFFFFFFFFFFFFFFFF


class Env:
    pass


def setup_env(args):
    import torch

    from obvhs_b200 import api

    e = Env()
    e.world = int(os.environ.get("WORLD_SIZE", "1"))
    e.rank = int(os.environ.get("RANK", "0"))
    e.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    e.numa_note = bind_to_gpu_numa_node(e.local_rank, e.world)
    torch.cuda.set_device(e.local_rank)
    e.dist = None
    if e.world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{e.local_rank}"))
        e.dist = dist
    e.dev = torch.device(f"cuda:{e.local_rank}")
    e.stream = torch.cuda.Stream(device=e.dev)
    e.ctx = api.Context(e.local_rank, stream=e.stream.cuda_stream)
    if os.environ.get("OBVHS_BENCH_HOST_SLICE"):  # tuning sweeps only
        e.ctx.set_option("host_slice", os.environ["OBVHS_BENCH_HOST_SLICE"])
    if e.dist:  # the library's own communicator (NCCL below the C ABI); torch.distributed only ships the 128-byte id
        box = [api.nccl_unique_id() if e.rank == 0 else None]
        e.dist.broadcast_object_list(box, src=0)
        e.ctx.comm_init(box[0], e.rank, e.world)
    with torch.cuda.stream(e.stream):
        e.flush = torch.empty(256 << 20, dtype=torch.uint8, device=e.dev)  # > 126 MB L2
    return e


def all_max(e, values):
    import torch

    t = torch.tensor(values, dtype=torch.float64, device=e.dev)
    if e.dist:
        e.dist.all_reduce(t, op=e.dist.ReduceOp.MAX)
    return t.tolist()


def all_same_u64(e, value: int) -> bool:
    """True when every rank holds the same 64-bit value."""
    import torch

    v = value - (1 << 64) if value >= (1 << 63) else value
    lo = torch.tensor([v], dtype=torch.int64, device=e.dev)
    hi = lo.clone()
    if e.dist:
        e.dist.all_reduce(lo, op=e.dist.ReduceOp.MIN)
        e.dist.all_reduce(hi, op=e.dist.ReduceOp.MAX)
    return int(lo.item()) == int(hi.item())


def host_link_probe(e, h_in, d_in, d_out, h_out):
    """Plain pinned copies of this rank's e2e buffers, both directions at once: the link's share of the e2e time."""
    import torch

    s_in, s_out = torch.cuda.Stream(device=e.dev), torch.cuda.Stream(device=e.dev)
    best = None
    for _ in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s_in):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"concurrent_ms": best * 1e3, "h2d_gbs": h_in.numel() * h_in.element_size() / best / 1e9,
            "d2h_gbs": h_out.numel() * h_out.element_size() / best / 1e9}


def measure(e, args, name, tris, n_tris, preset, d_args, n_total, kernel, bound, oracle_tree=None):
    """One workload through the timed loop. tris: host triangles on rank 0 (None elsewhere); d_args: this rank's (n, 8) device ray
    records. Returns the result object (rank 0) -- every rank must call it (collectives inside)."""
    import torch

    from obvhs_b200 import api
    from obvhs_b200.types import RAY_HIT, make_rays

    ctx, stream, dev, rank, world, dist = e.ctx, e.stream, e.dev, e.rank, e.world, e.dist
    params = api.BvhBuildParams.preset(preset)
    n_rays = int(d_args.shape[0])
    with torch.cuda.stream(stream):
        d_tris = torch.from_numpy(tris).to(dev) if rank == 0 else None
        d_hits = torch.empty((n_rays, 4), dtype=torch.int32, device=dev)
        d_counters = torch.zeros(2, dtype=torch.int64, device=dev)
    stream.synchronize()

    def ev():
        return torch.cuda.Event(enable_timing=True)

    state = {"bvh": None}

    def one_step():
        with torch.cuda.stream(stream):
            e.flush.zero_()  # L2 flush between iterations, outside the timed events
            e0, e1, e2, e3 = ev(), ev(), ev(), ev()
            e0.record(stream)
            bvh = api.build_cwbvh_from_tris(d_tris, params, ctx=ctx) if rank == 0 else state["bvh"]
            e1.record(stream)
            if world > 1:
                bvh = api.CwBvh.broadcast(bvh, ctx, 0)
            e2.record(stream)
            bvh.ray_traverse(d_args, out=d_hits)
            e3.record(stream)
        stream.synchronize()
        # the previous tree stays alive while the next one is built (the context's result cache then holds both buffer sets)
        state["prev"], state["bvh"] = state["bvh"], bvh
        # the broadcast is timed on the building rank: elsewhere e0..e2 mostly waits for rank 0's build
        return e0.elapsed_time(e1) if rank == 0 else 0.0, e1.elapsed_time(e2) if rank == 0 else 0.0, e2.elapsed_time(e3), e0.elapsed_time(e3)

    for _ in range(args.warmup):
        one_step()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(e.local_rank, period_s=0.02)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    sums = [0.0, 0.0, 0.0, 0.0]
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        torch.cuda.nvtx.range_push("obvhs_timed_step")  # (ncu --nvtx --nvtx-include "obvhs_timed_step/" lists exactly one step's launches)
        for k, v in enumerate(one_step()):
            sums[k] += v
        torch.cuda.nvtx.range_pop()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    build_ms, bcast_ms, trav_ms, step_ms = [v / args.steps for v in all_max(e, sums)]
    my_trav_ms = sums[2] / args.steps
    value = n_total / (trav_ms * 1e-3) / 1e6  # all ranks' rays / max-over-ranks time
    bvh = state["bvh"]

    # ---- roofline of the traversal kernel on THIS GPU: algorithmic bytes from the per-launch counters -------------------
    with torch.cuda.stream(stream):
        d_counters.zero_()
        bvh.ray_traverse(d_args, out=d_hits, counters=d_counters)
    stream.synchronize()
    nodes_visited, tris_tested = [int(x) for x in d_counters.tolist()]
    alg_bytes = n_rays * (32 + 16) + 80 * nodes_visited + 48 * tris_tested  # SURVEY.md section 8(d) B_trav
    peak, peak_src = measured_peak_hbm()
    achieved = alg_bytes / (my_trav_ms * 1e-3) / 1e9
    traffic = ncu_traffic(name, n_rays)
    hit_count = int((d_hits[:, 3].view(torch.float32) < 3.0e38).sum().item())

    # ---- parity: replicas identical across ranks, this run's hits equal to the CPU oracle's on a sample -----------------
    nodes_p, prims_p, tris_p = bvh.device_ptrs()
    tree_hash = (bytes_hash_dev(nodes_p, bvh.node_count * 80, e.local_rank) ^ bytes_hash_dev(prims_p, bvh.prim_count * 4, e.local_rank)
                 ^ bytes_hash_dev(tris_p, bvh.prim_count * 64, e.local_rank))
    n_probe = min(1 << 20, n_rays)
    with torch.cuda.stream(stream):  # (torch's collectives order themselves against the CURRENT stream: keep everything on ours)
        probe_n = torch.tensor([n_probe], dtype=torch.int64, device=dev)
        if dist:
            dist.broadcast(probe_n, src=0)
        n_probe = int(probe_n.item())
        probe = d_args[:n_probe].clone() if rank == 0 else torch.empty((n_probe, 8), dtype=torch.float32, device=dev)
        if dist:
            dist.broadcast(probe, src=0)
        p_hits = torch.empty((n_probe, 4), dtype=torch.int32, device=dev)
        bvh.ray_traverse(probe, out=p_hits)
    stream.synchronize()
    probe_hash = hits_hash_dev(p_hits)
    parity = {"replicas_identical_across_ranks": all_same_u64(e, tree_hash), "probe_hits_identical_across_ranks": all_same_u64(e, probe_hash),
              "probe_rays": n_probe, "tree_hash": f"{tree_hash:016x}", "probe_hits_hash": f"{probe_hash:016x}"}
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        import oracle_bind as ob

        threads = len(os.sched_getaffinity(0))
        t0 = time.perf_counter()
        c = oracle_tree if oracle_tree is not None else ob.build_cwbvh_from_tris(tris, preset, threads=threads)
        cpu_build_wall = time.perf_counter() - t0
        bt = c.bvh_tris(tris)
        gn, gp, _ = bvh.download()
        wn, wp, _ = c.get()
        n_cmp = min(args.parity_rays, n_rays)
        idx = torch.arange(0, n_rays, max(1, n_rays // n_cmp), device=dev)[:n_cmp]
        a = d_args[idx].cpu().numpy()
        sample = make_rays(a[:, 0:3], a[:, 4:7], a[:, 3], a[:, 7])
        got = d_hits[idx].cpu().numpy().view(RAY_HIT).reshape(-1)
        want = c.ray_traverse(bt, sample, threads=threads, use_simd=True)
        parity.update({"oracle_sample_rays": int(sample.shape[0]),
                       "cwbvh_nodes_bit_exact": bool(gn.shape == wn.shape and gn.tobytes() == wn.tobytes()),
                       "primitive_indices_equal": bool(np.array_equal(gp, wp)),
                       "hit_ids_mismatch": int((got["primitive_id"] != want["primitive_id"]).sum()),
                       "hit_t_bits_mismatch": int((got["t"].view(np.uint32) != want["t"].view(np.uint32)).sum()),
                       "hits_hash_gpu": f"{hits_hash_np(got):016x}", "hits_hash_oracle": f"{hits_hash_np(want):016x}"})
        parity["ok"] = bool(parity["cwbvh_nodes_bit_exact"] and parity["primitive_indices_equal"] and parity["hit_ids_mismatch"] == 0
                            and parity["hit_t_bits_mismatch"] == 0)
        if world == 1:  # the CPU baseline proper: rank 0, N = 1 only
            m = min(args.ref_rays, n_rays)
            idx = torch.arange(0, n_rays, max(1, n_rays // m), device=dev)[:m]
            a = d_args[idx].cpu().numpy()
            sample = make_rays(a[:, 0:3], a[:, 4:7], a[:, 3], a[:, 7])
            c.ray_traverse(bt, sample[:4096], threads=threads, use_simd=True)
            t0 = time.perf_counter()
            c.ray_traverse(bt, sample, threads=threads, use_simd=True)
            dt = time.perf_counter() - t0
            cpu = {"value": sample.shape[0] / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                   "sample": f"{sample.shape[0]} of {n_rays} rays (strided), one full build; C++ restatement of the obvhs CPU path (OpenMP), "
                             "not the rustc/rayon binary",
                   "build_mtris_per_s": n_tris / c.core_build_seconds / 1e6, "build_wall_s": cpu_build_wall}
        del c, bt
    ok_flag = torch.tensor([1 if parity.get("ok", True) and parity["replicas_identical_across_ranks"] and parity["probe_hits_identical_across_ranks"] else 0],
                           dtype=torch.int64, device=dev)
    if dist:
        dist.broadcast(ok_flag, src=0)

    # ---- e2e: the same calls with pinned HOST buffers, copies inside the timed region; every rank its own slice -----------
    h_args = torch.empty((n_rays, 8), dtype=torch.float32).pin_memory()
    h_args.copy_(d_args)
    h_hits = torch.empty((n_rays, 4), dtype=torch.int32).pin_memory()
    hits_np = h_hits.numpy().view(RAY_HIT).reshape(-1)
    e_t, e_b, e_s, e_n = [], [], [], []
    if dist:
        dist.barrier()
    for it in range(2 + max(2, min(args.steps, 5))):
        t1 = time.perf_counter()
        bvh.ray_traverse(h_args.numpy(), out=hits_np)
        t2 = time.perf_counter()
        if it >= 2:
            e_n.append(t2 - t1)
    e2e_hit_count = int((hits_np["t"] < 3.0e38).sum())
    assert e2e_hit_count == hit_count, (e2e_hit_count, hit_count)
    # Every ray of these workloads is Ray::new_inf(origin, direction) / Ray::new(o, d, 0.0, f32::MAX) (examples/demoscene.rs:152,178;
    # obj_cwbvh.rs:104): one (tmin, tmax) for the batch, so the call a user makes ships origin and direction only (24 B per ray)
    with torch.cuda.stream(stream):
        t_lo, t_hi = d_args[:, 3].min().item(), d_args[:, 3].max().item()
        m_lo, m_hi = d_args[:, 7].min().item(), d_args[:, 7].max().item()
    uniform = bool(t_lo == t_hi and m_lo == m_hi)
    if uniform:
        h_od = torch.empty((n_rays, 6), dtype=torch.float32).pin_memory()
        with torch.cuda.stream(stream):
            d_od = d_args[:, [0, 1, 2, 4, 5, 6]].contiguous()
            h_od.copy_(d_od)
            ref_hash = hits_hash_np(hits_np)
        stream.synchronize()
        h_hits8 = torch.empty((n_rays, 2), dtype=torch.int32).pin_memory()  # ObvhsRayHit8 {primitive_id, t}: all this path writes into a RayHit
        hits8_np = h_hits8.numpy()
        for it in range(2 + max(2, min(args.steps, 5))):
            t1 = time.perf_counter()
            bvh.ray_od_traverse(h_od.numpy(), t_lo, m_lo, out=hits8_np, hit8=True)
            t2 = time.perf_counter()
            if it >= 2:
                e_t.append(t2 - t1)
        h8 = hits8_np.view(np.uint32)
        with np.errstate(over="ignore"):
            hash8 = int(((h8[:, 1].astype(np.uint64) << np.uint64(32)) | h8[:, 0].astype(np.uint64)).sum(dtype=np.uint64))
        assert hash8 == ref_hash, "24-byte ray records / 8-byte hit records gave different hits than the Ray::new records"
        link = host_link_probe(e, h_od, d_od, d_hits.view(-1)[: 2 * n_rays].view(n_rays, 2), h_hits8)
        del d_od, h_od, h_hits8, hits8_np, h8
    else:
        e_t = list(e_n)
        link = host_link_probe(e, h_args, d_args, d_hits, h_hits)
    # the drop-in call over the reference's own 64-byte Ray structs, on a bounded prefix of the slice
    n_struct = min(n_rays, 1 << 25)
    h_rays = torch.empty((n_struct, 16), dtype=torch.float32).pin_memory()
    with torch.cuda.stream(stream):
        d_rays = torch.empty((n_struct, 16), dtype=torch.float32, device=dev)
        api.ray_new(d_args[:n_struct], out=d_rays, ctx=ctx)
        h_rays.copy_(d_rays)
    stream.synchronize()
    del d_rays
    for it in range(2 + max(2, min(args.steps, 5))):
        t1 = time.perf_counter()
        bvh.ray_traverse(h_rays.numpy(), out=hits_np[:n_struct])
        t2 = time.perf_counter()
        if it >= 2:
            e_s.append(t2 - t1)
    if rank == 0:  # host-to-host build
        h_tris = torch.from_numpy(tris).pin_memory()
        for it in range(4):
            t0 = time.perf_counter()
            eb = api.build_cwbvh_from_tris(h_tris.numpy(), params, ctx=ctx)
            if it >= 1:
                e_b.append(time.perf_counter() - t0)
        del eb, h_tris
    e2e_trav_s, e2e_struct_s, e2e_build_s, e2e_new_s = all_max(e, [float(np.mean(e_t)), float(np.mean(e_s)) / n_struct * n_rays,
                                                                    float(np.mean(e_b)) if e_b else 0.0, float(np.mean(e_n))])
    link_all = all_max(e, [link["concurrent_ms"], -link["h2d_gbs"], -link["d2h_gbs"]])
    res = None
    if rank == 0:
        res = {
            "value": value, "unit": "Mrays/s", "ms_per_step": step_ms, "traverse_ms": trav_ms, "broadcast_ms": bcast_ms if world > 1 else 0.0,
            "rays_total": n_total, "rays_this_gpu": n_rays, "tris": n_tris, "hits_this_gpu": hit_count,
            "build": {"value": n_tris / (build_ms * 1e-3) / 1e6 if build_ms > 0 else None, "unit": "Mtris/s", "ms": build_ms, "cwbvh_nodes": bvh.node_count,
                      "floor_bytes": 48 * n_tris + 4 * n_tris + 80 * bvh.node_count,
                      "frac_of_hbm_at_floor": (48 * n_tris + 4 * n_tris + 80 * bvh.node_count) / (build_ms * 1e-3) / 1e9 / peak if build_ms > 0 else None},
            "broadcast": ({"ms": bcast_ms, "bytes": 80 * bvh.node_count + 68 * bvh.prim_count,
                           "gbs": (80 * bvh.node_count + 68 * bvh.prim_count) / (bcast_ms * 1e-3) / 1e9 if bcast_ms > 0 else None,
                           "how": "obvhs_cuda_cwbvh_broadcast on the building rank's stream: 64-byte header + one grouped NCCL launch (nodes, indices, RtTriangles)"}
                          if world > 1 else None),
            "roofline": {"bound": bound, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "frac_algorithmic": achieved / peak,
                         "frac_dram": (traffic / (my_trav_ms * 1e-3) / 1e9 / peak) if traffic else None, "traffic": traffic,
                         "kernel": kernel, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": my_trav_ms,
                         "nodes_visited": nodes_visited, "tris_tested": tris_tested,
                         "note": "per GPU (rank 0's launch). B_trav = rays*(32+16) + 80*nodes_visited + 48*tris_tested (SURVEY.md 8d); traffic = DRAM bytes "
                                 "of the same launch from the committed ncu capture (profiles/traffic.json), null when the launch size differs from it"},
            "cpu_baseline": cpu, "parity": parity,
            "e2e": {"value": n_total / e2e_trav_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": (24 if uniform else 32) * n_rays, "d2h_bytes_per_step": (8 if uniform else 16) * n_rays,
                    "how": ("obvhs_cuda_cwbvh_ray_od_traverse_hit8_batch: pinned HOST (origin, direction) records (24 B/ray, one tmin/tmax per batch = Ray::new_inf) in, "
                            "pinned HOST {primitive_id, t} records (8 B/ray) out, "
                            if uniform else "obvhs_cuda_cwbvh_ray_new_traverse_batch: pinned HOST Ray::new records (32 B/ray) in, pinned HOST RayHits out, ") +
                           "every rank its own slice, max over ranks; the kernels run Ray::new as they fetch a ray",
                    "ray_new": {"value": n_total / e2e_new_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 32 * n_rays,
                                "d2h_bytes_per_step": 16 * n_rays,
                                "how": "obvhs_cuda_cwbvh_ray_new_traverse_batch over 32-byte Ray::new records (per-ray tmin / tmax), 16-byte RayHits out"},
                    "ray_struct": {"value": n_total / e2e_struct_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 64 * n_rays, "measured_on_rays": n_struct,
                                   "how": "obvhs_cuda_cwbvh_ray_traverse_batch over the reference's 64-byte Ray structs (the drop-in call)"},
                    "build": {"mtris_per_s": n_tris / e2e_build_s / 1e6 if e2e_build_s > 0 else None, "h2d_bytes": 48 * n_tris,
                              "how": "obvhs_cuda_build_cwbvh_from_tris from pinned HOST triangles, tree left on the device (rank 0)"},
                    "host_link": {"slowest_rank_concurrent_copy_ms": link_all[0], "slowest_rank_h2d_gbs": -link_all[1], "slowest_rank_d2h_gbs": -link_all[2],
                                  "how": "plain pinned cudaMemcpyAsync of this rank's ray records up and hits down at the same time (what bounds e2e)"}},
            "gpu_launches": launches, "clocks": clocks, "wall_s_timed_region": wall,
        }
    state.clear()
    return res, bool(ok_flag.item())


def run_headline(args):
    import torch

    from obvhs_b200 import api, camera, device_rays, test_util as tu
    from obvhs_b200.sharding import device_bytes_tensor, shard_range
    from obvhs_b200.types import ray_args_of

    e = setup_env(args)
    config, res_terrain = headline_config(args)
    # ---- S3 ----------------------------------------------------------------------------------------------------------
    n_tris = 2 * res_terrain * res_terrain
    tris = cached_demoscene(res_terrain) if e.rank == 0 else None
    params = api.BvhBuildParams.fast_build()
    with torch.cuda.stream(e.stream):
        bvh0 = api.build_cwbvh_from_tris(torch.from_numpy(tris).to(e.dev), params, ctx=e.ctx) if e.rank == 0 else None
        if e.world > 1:
            bvh0 = api.CwBvh.broadcast(bvh0, e.ctx, 0)
        _, _, tris_p = bvh0.device_ptrs()
        rt_tris = device_bytes_tensor(tris_p, bvh0.prim_count * 64, e.local_rank).view(torch.float32).view(-1, 16)
        lo, hi = shard_range(args.rays, e.rank, e.world)
        cam = device_rays.DeviceCamera(camera.demoscene_camera(1280), e.dev)
        d_args, aa_samples, n_primary = device_rays.bounce_set(cam, bvh0, rt_tris, lo, hi, args.rays)
    e.stream.synchronize()
    del rt_tris, bvh0
    s3, ok_s3 = measure(e, args, "s3", tris, n_tris, "fast_build", d_args, args.rays, "traverse_persistent_kernel<CwTree, closest>", "issue/L1-L2")
    del d_args, tris
    torch.cuda.empty_cache()
    # ---- kitchen -------------------------------------------------------------------------------------------------------
    ktris = tu.kitchen()
    krays = ray_args_of(camera.primary_rays(camera.kitchen_camera(1920)))
    klo, khi = shard_range(krays.shape[0], e.rank, e.world)
    with torch.cuda.stream(e.stream):
        d_kargs = torch.from_numpy(np.ascontiguousarray(krays[klo:khi])).to(e.dev)
    kitchen, ok_k = measure(e, args, "kitchen", ktris if e.rank == 0 else None, ktris.shape[0], "fast_build", d_kargs, krays.shape[0],
                            "traverse_kernel<CwTree, closest>", "issue/L2")
    if e.rank == 0:
        s3["roofline"]["note"] += ("; on this ray set (7.6 nodes and 0.3 triangles per ray) the upper tree levels are served by L1 (76 % hits) and L2 (57 %): "
                                   "the kernel is issue-bound (ncu, profiles/r2a_traverse_kernel_s3_full.md: issue slots 77 % busy, ALU pipe 68 %, 23.4 of 32 "
                                   "lanes per instruction); frac_algorithmic is cache-served reuse, frac_dram the share of HBM bandwidth actually used")
        kitchen["roofline"]["note"] += ("; the kitchen's tree (0.5 MB) and triangles (3.6 MB) live in L1/L2: frac_algorithmic measures cache-served reuse, "
                                       "the kernel is issue-bound (ncu: profiles/), frac_dram is the share of HBM bandwidth it actually uses")
        line = {"metric": METRIC, "value": s3["value"], "unit": "Mrays/s", "n_gpus": e.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": s3["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic (demoscene terrain, device-generated rays); kitchen.obj fixture (reference asset) in `kitchen`", "config": config,
                "detail": {"aa_samples_walked": aa_samples, "primary_rays_traced_for_the_bounce_set": n_primary, "host_numa": e.numa_note,
                           "rays_per_gpu": hi - lo}}
        for k in ("build", "traverse_ms", "broadcast_ms", "broadcast", "roofline", "cpu_baseline", "parity", "e2e", "gpu_launches", "clocks",
                  "wall_s_timed_region"):
            line[k] = s3[k]
        line["gpu_launches"] = s3["gpu_launches"] + kitchen["gpu_launches"]
        line["kitchen"] = kitchen
        line["parity_ok"] = bool(ok_s3 and ok_k)
        print(json.dumps(line))
    if e.dist:
        e.dist.barrier()
        e.dist.destroy_process_group()
    if not (ok_s3 and ok_k):
        raise SystemExit("bench.py: parity check failed (see `parity` in the JSON line)")


def run_cornell(args):
    """--workload cornell: BASELINE.json configs[0] / SURVEY.md 8(d) S0 (examples/cornell_box_cwbvh.rs:22-126): 34 triangles,
    medium_build, 1280x720 primary rays, eye (0,1,2.1) -> (0,1,0), fov 90. Sharded over the GPUs like the headline."""
    import torch

    from obvhs_b200 import camera, test_util as tu
    from obvhs_b200.sharding import shard_range
    from obvhs_b200.types import ray_args_of

    tris = tu.cornell_box()
    rays = camera.primary_rays(camera.cornell_camera(1280, 720))
    config = {"workload": f"S0: cornell box {tris.shape[0]} tris (examples/cornell_box_cwbvh.rs), medium_build, 1280x720 primary rays", "preset": "medium_build",
              "tris": int(tris.shape[0]), "rays_total": int(rays.shape[0]), "l2": "flushed between steps (256 MB write)"}
    if args.impl == "reference":
        if int(os.environ.get("RANK", "0")) != 0:
            return
        r = cpu_reference_leg(tris, rays, "medium_build", max(1, min(args.steps, 5)), 1)
        v = rays.shape[0] / r["trav_s"] / 1e6
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": (r["trav_s"] + r["build_s"]) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                          "data": "cornell box (the example's own 34 triangles)", "config": config,
                          "build": {"value": tris.shape[0] / r["build_s"] / 1e6, "unit": "Mtris/s", "ms": r["build_s"] * 1e3},
                          "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": r["threads"], "kind": "port", "sample": "all rays, full build per step"},
                          "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    e = setup_env(args)
    args_all = ray_args_of(rays)
    lo, hi = shard_range(args_all.shape[0], e.rank, e.world)
    with torch.cuda.stream(e.stream):
        d_args = torch.from_numpy(np.ascontiguousarray(args_all[lo:hi])).to(e.dev)
    res, ok = measure(e, args, "cornell", tris if e.rank == 0 else None, tris.shape[0], "medium_build", d_args, args_all.shape[0],
                      "traverse_kernel<CwTree, closest>", "issue/L1")
    if e.rank == 0:
        line = {"metric": METRIC, "n_gpus": e.world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "cornell box (the example's own 34 triangles), generated rays", "config": config}
        line.update(res)
        line["parity_ok"] = ok
        print(json.dumps(line))
    if e.dist:
        e.dist.barrier()
        e.dist.destroy_process_group()
    if not ok:
        raise SystemExit("bench.py: parity check failed (see `parity` in the JSON line)")


def run_passes(args):
    """--workload passes: the two layout passes next to the hot path, each timed on the device with the oracle's
    sequential loop timed beside it and byte-compared: `Bvh2::reorder_in_stack_traversal_order` (src/bvh2/mod.rs:462-500) and
    `CwBvh::order_children` as a separate pass (src/cwbvh/mod.rs:520-735) on a tree converted WITHOUT ordering. One GPU."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import torch

    import oracle_bind as ob  # the checker and the CPU leg only
    from obvhs_b200 import api, camera, test_util as tu
    from obvhs_b200.types import ray_args_of

    e = setup_env(args) if int(os.environ.get("WORLD_SIZE", "1")) == 1 else None
    if e is None:
        raise SystemExit("--workload passes is a one-GPU measurement (run it without torchrun)")
    res = S3_RES if args.tris == 10_000_000 else int(round((args.tris / 2) ** 0.5))
    scenes = [("kitchen", tu.kitchen(), camera.kitchen_camera(1920)), (f"demoscene({res},0)", cached_demoscene(res), camera.demoscene_camera(1280))]
    params = api.BvhBuildParams.fast_build()
    single = dataclasses.replace(params, max_prims_per_leaf=1)
    out = {}

    def timed_ms(fn):
        with torch.cuda.stream(e.stream):
            e.flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(e.stream)
            r = fn()
            b.record(e.stream)
        e.stream.synchronize()
        return a.elapsed_time(b), r

    for name, tris, cam in scenes:
        n = tris.shape[0]
        with torch.cuda.stream(e.stream):
            d_tris = torch.from_numpy(tris).to(e.dev)
            v = d_tris.view(n, 3, 4)[:, :, :3]
            d_aabbs = torch.zeros((n, 8), dtype=torch.float32, device=e.dev)
            d_aabbs[:, 0:3] = v.amin(1)
            d_aabbs[:, 4:7] = v.amax(1)
            d_rays = torch.from_numpy(ray_args_of(camera.primary_rays(cam))).to(e.dev)
            n_rays = d_rays.shape[0]
            d_hits = torch.empty((n_rays, 4), dtype=torch.int32, device=e.dev)
        e.stream.synchronize()
        aabbs = d_aabbs.cpu().numpy()
        steps = max(1, min(args.steps, 5))

        def t_hash():
            return int((d_hits[:, 3].to(torch.int64) & 0xFFFFFFFF).sum().item())

        # ---- Bvh2::reorder_in_stack_traversal_order
        ms, trav_before, trav_after = [], [], []
        for it in range(steps):
            with torch.cuda.stream(e.stream):
                b2 = api.build_bvh2_from_tris(d_tris, params, ctx=e.ctx)
            if it == 0:
                before_nodes, before_prims = b2.download()
                depth = b2.max_depth
            trav_before.append(timed_ms(lambda: b2.ray_traverse(d_rays, out=d_hits))[0])
            h0 = t_hash()
            ms.append(timed_ms(b2.reorder_in_stack_traversal_order)[0])
            trav_after.append(timed_ms(lambda: b2.ray_traverse(d_rays, out=d_hits))[0])
            h1 = t_hash()
        got_nodes, got_prims = b2.download()
        w = ob.bvh2_from(before_nodes, before_prims, depth)
        t0 = time.perf_counter()
        w.reorder_in_stack_traversal_order()
        cpu_ms = (time.perf_counter() - t0) * 1e3
        wn, wp = w.get()
        nodes2 = int(got_nodes.shape[0])
        reorder = {"ms": float(np.mean(ms)), "nodes": nodes2, "Mnodes_per_s": nodes2 / float(np.mean(ms)) / 1e3,
                   "algorithmic_bytes": nodes2 * (48 + 48 + 4 + 4 + 4), "gbs": nodes2 * 108 / float(np.mean(ms)) / 1e6,
                   "cpu_oracle_ms": cpu_ms, "bit_exact_vs_oracle": bool(got_nodes.tobytes() == wn.tobytes() and np.array_equal(got_prims, wp)),
                   "bvh2_traverse_ms_before": float(np.mean(trav_before)), "bvh2_traverse_ms_after": float(np.mean(trav_after)),
                   "closest_t_unchanged": bool(h0 == h1), "rays": n_rays}
        del b2, w, wn, wp, got_nodes, before_nodes
        # ---- CwBvh::order_children on a tree converted without the converter's ordering
        ms, trav_unordered, trav_ordered = [], [], []
        for it in range(steps):
            with torch.cuda.stream(e.stream):
                b2 = api.build_bvh2_from_tris(d_tris, single, ctx=e.ctx)  # the converter wants one primitive per Bvh2 leaf (bvh2_to_cwbvh.rs:201)
                cw = api.bvh2_to_cwbvh(b2, 3, False, False)
                cw.set_triangles(d_tris)
            if it == 0:
                u_nodes, u_prims, _ = cw.download()
                total = cw.total_aabb()
            trav_unordered.append(timed_ms(lambda: cw.ray_traverse(d_rays, out=d_hits))[0])
            h0 = t_hash()
            ms.append(timed_ms(lambda: cw.order_children(d_aabbs, False))[0])
            trav_ordered.append(timed_ms(lambda: cw.ray_traverse(d_rays, out=d_hits))[0])
            h1 = t_hash()
        g_nodes, g_prims, _ = cw.download()
        w = ob.cwbvh_from(u_nodes, u_prims, total)
        t0 = time.perf_counter()
        w.order_children(aabbs, False)
        cpu_ms = (time.perf_counter() - t0) * 1e3
        wn, wp, _ = w.get()
        nn = int(g_nodes.shape[0])
        order = {"ms": float(np.mean(ms)), "nodes": nn, "Mnodes_per_s": nn / float(np.mean(ms)) / 1e3,
                 "cpu_oracle_ms": cpu_ms, "bit_exact_vs_oracle": bool(g_nodes.tobytes() == wn.tobytes() and np.array_equal(g_prims, wp)),
                 "cwbvh_traverse_ms_unordered": float(np.mean(trav_unordered)), "cwbvh_traverse_ms_ordered": float(np.mean(trav_ordered)),
                 "closest_t_unchanged": bool(h0 == h1), "rays": n_rays}
        out[name] = {"tris": int(n), "reorder_in_stack_traversal_order": reorder, "order_children": order}
        del b2, cw, w, d_tris, d_aabbs, d_rays, d_hits
        torch.cuda.empty_cache()
    ok = all(s[k]["bit_exact_vs_oracle"] and s[k]["closest_t_unchanged"] for s in out.values() for k in ("reorder_in_stack_traversal_order", "order_children"))
    print(json.dumps({"metric": "layout_pass_ms", "unit": "ms", "higher_is_better": False, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
                      "dtype": "u8/f32", "data": "kitchen.obj fixture; synthetic demoscene terrain", "preset": "fast_build",
                      "timing": "CUDA events on the context's stream, L2 flushed before every timed call, mean of the steps; the CPU figure is the "
                                "oracle's single-threaded sequential loop on this box (the reference's passes are single-threaded)",
                      "scenes": out, "parity_ok": ok}))
    if not ok:
        raise SystemExit("bench.py: layout pass parity failed")


def run_headline_reference(args):
    """--impl reference for the headline: the CPU restatement on a bounded sample of both workloads, same `config`."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import oracle_bind as ob
    from obvhs_b200 import camera, test_util as tu

    config, res_terrain = headline_config(args)
    threads = len(os.sched_getaffinity(0))
    tris = cached_demoscene(res_terrain)
    n_builds = max(1, min(args.steps, 2))
    build_s = []
    c = None
    for _ in range(n_builds):
        c = ob.build_cwbvh_from_tris(tris, "fast_build", threads=threads)
        build_s.append(c.core_build_seconds)
    bt = c.bvh_tris(tris)
    n_samples = max(1, int(round(args.ref_rays / (1280 * 475 * 0.6))))
    rays, _ = camera.demoscene_bounce_set(camera.demoscene_camera(1280), range(n_samples), bt,
                                          lambda r: c.ray_traverse(bt, r, threads=threads, use_simd=True))
    rays = rays[: args.ref_rays]
    ts = []
    for it in range(min(args.warmup, 1) + args.steps):
        t0 = time.perf_counter()
        c.ray_traverse(bt, rays, threads=threads, use_simd=True)
        if it >= min(args.warmup, 1):
            ts.append(time.perf_counter() - t0)
    trav_s, b_s = float(np.mean(ts)), float(np.mean(build_s))
    mrays = rays.shape[0] / trav_s / 1e6
    ktris = tu.kitchen()
    krays = camera.primary_rays(camera.kitchen_camera(1920))
    kr = cpu_reference_leg(ktris, krays, "fast_build", max(1, min(args.steps, 5)), 1)
    kmrays = krays.shape[0] / kr["trav_s"] / 1e6
    sample = (f"S3: {rays.shape[0]} bounce rays of AA samples 0..{n_samples - 1} (CPU-generated with the same recipe) per step, {n_builds} full builds of "
              f"{tris.shape[0]} tris; kitchen: all {krays.shape[0]} rays and a full build per step; C++ restatement of the obvhs CPU path (OpenMP), not the "
              "rustc/rayon binary")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": mrays, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": (trav_s + b_s) * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic (demoscene terrain); kitchen.obj fixture (reference asset) in `kitchen`", "config": config,
        "build": {"value": tris.shape[0] / b_s / 1e6, "unit": "Mtris/s", "ms": b_s * 1e3},
        "cpu_baseline": {"value": mrays, "unit": "Mrays/s", "cores": threads, "kind": "port", "sample": sample, "build_mtris_per_s": tris.shape[0] / b_s / 1e6},
        "e2e": {"value": mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "kitchen": {"value": kmrays, "unit": "Mrays/s", "build": {"value": ktris.shape[0] / kr["build_s"] / 1e6, "unit": "Mtris/s", "ms": kr["build_s"] * 1e3},
                    "e2e": {"value": kmrays, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}},
        "gpu_launches": 0}))


def run_ours(args):
    import torch

    from obvhs_b200 import api, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    numa_note = bind_to_gpu_numa_node(local_rank, world)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")
    stream = torch.cuda.Stream(device=dev)
    ctx = api.Context(local_rank, stream=stream.cuda_stream)
    if os.environ.get("OBVHS_BENCH_HOST_SLICE"):  # tuning sweeps only
        ctx.set_option("host_slice", os.environ["OBVHS_BENCH_HOST_SLICE"])

    tris, rays, desc, preset = make_workload(args.workload, args.tris)
    params = api.BvhBuildParams.preset(preset)
    n_tris = tris.shape[0]
    strong = rays is None
    with torch.cuda.stream(stream):
        d_tris = torch.from_numpy(tris).to(dev)
    if strong:
        # rank 0 builds once, every rank gets a replica and generates the bounce rays of ITS slice of the AA samples
        from obvhs_b200 import camera
        from obvhs_b200.types import RAY_HIT

        with torch.cuda.stream(stream):
            bvh0 = api.build_cwbvh_from_tris(d_tris, params, ctx=ctx) if rank == 0 else None
            if world > 1:
                bvh0 = sharding.broadcast_cwbvh(bvh0, ctx, src=0)
            _, prim_idx, _ = bvh0.download()
            bvh_tris = tris[prim_idx]
            rays, n_primary = camera.demoscene_bounce_set(camera.demoscene_camera(1280), bounce_samples(args.samples, rank, world), bvh_tris,
                                                          lambda r: bvh0.ray_traverse(r))
        del bvh0, bvh_tris
        desc += f", {args.samples} AA samples"
    n_rays = rays.shape[0]
    with torch.cuda.stream(stream):
        d_rays = torch.from_numpy(rays).to(dev)
        d_hits = torch.empty((n_rays, 4), dtype=torch.int32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
        d_counters = torch.zeros(2, dtype=torch.int64, device=dev)
    stream.synchronize()
    n_rays_all = n_rays
    if dist:
        t = torch.tensor([n_rays], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        n_rays_all = int(t.item())

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def one_step(timed):
        """build (rank 0) -> [broadcast] -> traverse. Returns (build_ms, bcast_ms, trav_ms) events' readings."""
        with torch.cuda.stream(stream):
            flush.zero_()  # L2 flush between iterations, outside the timed events
            e0, e1, e2, e3 = ev(), ev(), ev(), ev()
            e0.record(stream)
            bvh = api.build_cwbvh_from_tris(d_tris, params, ctx=ctx) if rank == 0 else None
            e1.record(stream)
            if world > 1:
                bvh = sharding.broadcast_cwbvh(bvh, ctx, src=0)
            e2.record(stream)
            bvh.ray_traverse(d_rays, out=d_hits)
            e3.record(stream)
        stream.synchronize()
        # the broadcast is timed on the building rank only: on the others e1..e2 mostly waits for rank 0's build
        return e0.elapsed_time(e1), (e1.elapsed_time(e2) if rank == 0 else 0.0), e2.elapsed_time(e3), e0.elapsed_time(e3), bvh

    bvh = None
    for _ in range(args.warmup):
        # keep the previous tree alive while the next one is built, exactly like the timed loop, so the context's result
        # cache holds both buffer sets before timing starts (a cold cudaMalloc of a 10M-triangle tree costs tens of ms)
        _, _, _, _, bvh = one_step(False)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count
    b_ms, c_ms, t_ms, s_ms = [], [], [], []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        b, c, t, whole, bvh = one_step(True)
        b_ms.append(b)
        c_ms.append(c)
        t_ms.append(t)
        s_ms.append(whole)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    wall = time.perf_counter() - wall0
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None

    tot = torch.tensor([sum(b_ms), sum(c_ms), sum(t_ms), sum(s_ms)], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)  # max over ranks
    build_ms, bcast_ms, trav_ms, step_ms = (tot / args.steps).tolist()
    value = n_rays_all / (trav_ms * 1e-3) / 1e6  # all ranks' rays / max-over-ranks time
    build_mtris = n_tris / (build_ms * 1e-3) / 1e6 if build_ms > 0 else None

    # ---- roofline of the dominant kernel (traverse_kernel): algorithmic bytes from the per-launch counters ----------
    with torch.cuda.stream(stream):
        d_counters.zero_()
        bvh.ray_traverse(d_rays, out=d_hits, counters=d_counters)
    stream.synchronize()
    nodes_visited, tris_tested = [int(x) for x in d_counters.tolist()]
    alg_bytes = n_rays * (32 + 16) + 80 * nodes_visited + 48 * tris_tested  # SURVEY.md section 8(d) B_trav
    peak, peak_src = measured_peak_hbm()
    achieved = alg_bytes / (trav_ms * 1e-3) / 1e9
    hit_count = int((d_hits[:, 3].view(torch.float32) < 3.0e38).sum().item())

    # ---- e2e: same step through the C ABI with HOST buffers (pinned), copies inside the timed region; every rank runs it
    h_tris = torch.from_numpy(tris).pin_memory()
    h_rays = torch.from_numpy(rays).pin_memory()
    h_hits = torch.empty((n_rays, 4), dtype=torch.int32).pin_memory()
    from obvhs_b200.types import RAY_HIT

    hits_np = h_hits.numpy().view(RAY_HIT).reshape(-1)
    # the same rays as Ray::new arguments (32 B: origin, tmin, direction, tmax): the constructor then runs on the device
    from obvhs_b200.types import ray_args_of

    h_args = torch.from_numpy(ray_args_of(rays)).pin_memory()
    e_b, e_t, e_s = [], [], []
    if dist:
        dist.barrier()
    for it in range(2 + max(2, min(args.steps, 5))):
        t0 = time.perf_counter()
        eb = api.build_cwbvh_from_tris(h_tris.numpy(), params, ctx=ctx)
        t1 = time.perf_counter()
        eb.ray_traverse(h_args.numpy(), out=hits_np)
        t2 = time.perf_counter()
        if it >= 2:
            e_b.append(t1 - t0)
            e_t.append(t2 - t1)
    assert int((hits_np["t"] < 3.0e38).sum()) == hit_count
    hits_np["t"] = 0
    for it in range(2 + max(2, min(args.steps, 5))):
        t1 = time.perf_counter()
        eb.ray_traverse(h_rays.numpy(), out=hits_np)
        t2 = time.perf_counter()
        if it >= 2:
            e_s.append(t2 - t1)
    assert int((hits_np["t"] < 3.0e38).sum()) == hit_count
    e2e_t = torch.tensor([float(np.mean(e_t)), float(np.mean(e_b)), float(np.mean(e_s))], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_trav_s, e2e_build_s, e2e_struct_s = e2e_t.tolist()
    e2e_mrays = n_rays_all / e2e_trav_s / 1e6
    e2e_struct_mrays = n_rays_all / e2e_struct_s / 1e6
    e2e_mtris = n_tris / e2e_build_s / 1e6

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            max_rays = args.ref_rays
            sample = rays[:: max(1, n_rays // max_rays)][:max_rays] if n_rays > max_rays else rays
            r = cpu_reference_leg(tris, sample, preset, 1 if n_tris > 2_000_000 else 2, 1 if n_tris <= 2_000_000 else 0)
            cpu = {"value": sample.shape[0] / r["trav_s"] / 1e6, "unit": "Mrays/s", "cores": r["threads"], "kind": "port",
                   "sample": f"{sample.shape[0]} of {n_rays} rays (strided) + full build; C++ restatement of the obvhs CPU path "
                             "(OpenMP), not the rustc/rayon binary",
                   "build_mtris_per_s": n_tris / r["build_s"] / 1e6}
        line = {
            "metric": METRIC, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic" if args.workload != "kitchen" else "kitchen.obj fixture (reference asset), generated rays",
            "config": {"workload": desc, "preset": preset, "rays_per_gpu": n_rays, "rays_total": n_rays_all, "tris": n_tris, "l2": "flushed between steps (256 MB write)",
                       "multi_gpu": "build on rank 0, NCCL broadcast, rays sharded (one batch per GPU)" if world > 1 else "single GPU",
                       "host_numa": numa_note},
            "build": {"value": build_mtris, "unit": "Mtris/s", "ms": build_ms, "cwbvh_nodes": bvh.node_count},
            "traverse_ms": trav_ms, "broadcast_ms": bcast_ms,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.workload),
                         "kernel": ("traverse_persistent_kernel<CwTree, closest>" if n_tris > 262144 or args.workload in ("soup", "bounce")
                                    else "traverse_kernel<CwTree, closest>"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "nodes_visited": nodes_visited, "tris_tested": tris_tested,
                         "note": "B_trav = rays*(32+16) + 80*nodes_visited + 48*tris_tested (SURVEY.md 8d); kitchen tree+tris fit in L2, "
                                 "so a fraction near or above 1 means L2-served reuse, not DRAM streaming"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 32 * n_rays + 48 * n_tris, "d2h_bytes_per_step": 16 * n_rays,
                    "build_mtris_per_s": e2e_mtris,
                    "how": "obvhs_cuda_build_cwbvh_from_tris + obvhs_cuda_cwbvh_ray_new_traverse_batch (Ray::new arguments, 32 B per ray, constructor on "
                           "the device) with pinned HOST buffers, every rank its own batch, max over ranks",
                    "ray_struct": {"value": e2e_struct_mrays, "unit": "Mrays/s", "h2d_bytes_per_step": 64 * n_rays,
                                   "how": "obvhs_cuda_cwbvh_ray_traverse_batch over the 64-byte Ray array (inv_direction filled on the host)"}},
            "gpu_launches": launches, "hits": hit_count, "clocks": clocks, "wall_s_timed_region": wall,
        }
        print(json.dumps(line))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def ncu_traffic(workload, n_rays=None):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/traffic.json), or None. With n_rays:
    only when the capture was taken on a launch of that many rays (the headline's N = 1 configuration)."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "traffic.json")) as f:
            ent = json.load(f).get(workload, {})
        if n_rays is not None and ent.get("rays") not in (None, n_rays):
            return None
        return ent.get("bytes")
    except (OSError, ValueError):
        return None


def dynamic_frames(tris, n_frames):
    """BASELINE config 5 / SURVEY.md 8(d) S4: per frame every vertex is displaced by 0.01*(hash_noise-0.5); returns the
    per-primitive AABBs of each frame, (n_frames, n, 8) f32."""
    from obvhs_b200 import test_util as tu

    t0 = tris.reshape(-1, 3, 4)
    k = np.arange(t0.shape[0] * 9, dtype=np.uint32)
    out = np.zeros((n_frames, t0.shape[0], 8), dtype=np.float32)
    for f in range(n_frames):
        noise = tu.hash_noise(k, np.uint32(f), np.uint32(17)).reshape(-1, 3, 3)
        v = t0[:, :, 0:3] + (noise - np.float32(0.5)) * np.float32(0.01)
        out[f, :, 0:3] = v.min(axis=1)
        out[f, :, 4:7] = v.max(axis=1)
    return out


def dynamic_frames_device(d_tris, n_frames):
    """The same frames generated on the device (torch), for the 100-frame run: (n_frames, n, 8) f32 on d_tris' device."""
    import torch

    from obvhs_b200 import device_rays

    n = d_tris.shape[0]
    t0 = d_tris.view(n, 3, 4)[:, :, 0:3]
    k = torch.arange(n * 9, dtype=torch.int64, device=d_tris.device)
    out = torch.zeros((n_frames, n, 8), dtype=torch.float32, device=d_tris.device)
    for f in range(n_frames):
        noise = device_rays.hash_noise(k, torch.full_like(k, f), 17).view(n, 3, 3)
        v = t0 + (noise - 0.5) * 0.01
        out[f, :, 0:3] = v.amin(dim=1)
        out[f, :, 4:7] = v.amax(dim=1)
    return out


def run_dynamic(args):
    """--workload dynamic: dynamic Bvh2 maintenance (examples/physics.rs update loop): a step = one frame = rewrite the leaf
    AABBs of 1M moving triangles, refit_all (bvh2/mod.rs:527-569), ReinsertionOptimizer::run(0.01) (reinsertion.rs:40-57)."""
    from obvhs_b200 import test_util as tu

    rank = int(os.environ.get("RANK", "0"))
    res = int(round((args.tris / 2) ** 0.5)) if args.tris != 10_000_000 else 708
    tris = tu.demoscene(res, 0)
    n = tris.shape[0]
    n_frames = max(1, args.frames)  # the GPU arm runs this many DISTINCT frames (100), generated on the device
    n_cpu_frames = min(n_frames, 8)  # the CPU legs use the first few of the same sequence (numpy generator, bounded)
    frames = dynamic_frames(tris, n_cpu_frames)
    desc = f"demoscene({res},0) {n} moving tris: per frame set leaf AABBs + refit_all + reinsertion(0.01); {n_frames} distinct frames"
    if args.impl == "reference":
        if rank != 0:
            return
        import oracle_bind as ob

        threads = len(os.sched_getaffinity(0))
        b = ob.ploc_build(ob.tri_aabbs(tris), None, 6, 64, 2, threads=threads)
        b.reinsertion_run(0.02, threads=threads)
        ts = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            b.set_leaf_aabbs(frames[it % n_cpu_frames])
            b.refit_all()
            b.reinsertion_run(0.01, threads=threads)
            if it >= args.warmup:
                ts.append(time.perf_counter() - t0)
        ms = float(np.mean(ts)) * 1e3
        v = n / ms / 1e3
        print(json.dumps({"impl": "reference", "metric": "dynamic Bvh2 refit + reinsertion Mtris/s", "value": v, "unit": "Mtris/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic", "config": {"workload": desc},
                          "cpu_baseline": {"value": v, "unit": "Mtris/s", "cores": threads, "kind": "port", "sample": f"{args.steps} full frames"},
                          "e2e": {"value": v, "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    import torch

    from obvhs_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:  # replicas only: the maintenance loop is globally ordered, every rank runs its own copy
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")
    stream = torch.cuda.Stream(device=dev)
    ctx = api.Context(local_rank, stream=stream.cuda_stream)
    with torch.cuda.stream(stream):
        d_tris = torch.from_numpy(tris).to(dev)
        d_frames = dynamic_frames_device(d_tris, n_frames)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        bvh = api.PlocBuilder(ctx).build_tris(api.PlocSearchDistance.Low, d_tris, api.SortPrecision.U64, 2)
        opt = api.ReinsertionOptimizer()
        opt.run(bvh, 0.02)
    stream.synchronize()

    def frame(it, src):
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(stream)
            bvh.set_leaf_aabbs(src[it % n_frames])
            e1.record(stream)
            opt.run(bvh, 0.01)
            e2.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1), e1.elapsed_time(e2)

    for it in range(args.warmup):
        frame(it, d_frames)
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    refit_ms, reins_ms = [], []
    n_timed = max(args.steps, n_frames)  # every distinct frame at least once
    for it in range(n_timed):
        a, b = frame(args.warmup + it, d_frames)
        refit_ms.append(a)
        reins_ms.append(b)
    torch.cuda.synchronize()
    launches = ctx.launch_count - l0
    clocks = sampler.stop() if rank == 0 else None
    tot = torch.tensor([sum(refit_ms), sum(reins_ms)], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(tot, op=dist.ReduceOp.MAX)
    refit, reins = (tot / n_timed).tolist()
    value = world * n / ((refit + reins) * 1e-3) / 1e6
    # e2e: per-frame AABBs come from pinned HOST memory
    n_e2e = min(n_frames, 8)
    h_frames = torch.empty((n_e2e, n, 8), dtype=torch.float32).pin_memory()
    h_frames.copy_(d_frames[:n_e2e])
    hs = [h_frames[f].numpy() for f in range(n_e2e)]
    e_t = []
    for it in range(2 + max(2, min(args.steps, 5))):
        t0 = time.perf_counter()
        bvh.set_leaf_aabbs(hs[it % n_e2e])
        opt.run(bvh, 0.01)
        ctx.synchronize()
        if it >= 2:
            e_t.append(time.perf_counter() - t0)
    e2e = torch.tensor([float(np.mean(e_t))], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(e2e, op=dist.ReduceOp.MAX)
    peak, peak_src = measured_peak_hbm()
    refit_bytes = 32 * n + 64 * (2 * n - 1) + 4 * (2 * n - 1)  # leaf AABBs in; every node read by its parent + written; parents
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle_bind as ob

            threads = len(os.sched_getaffinity(0))
            b = ob.ploc_build(ob.tri_aabbs(tris), None, 6, 64, 2, threads=threads)
            b.reinsertion_run(0.02, threads=threads)
            ts = []
            for it in range(4):
                t0 = time.perf_counter()
                b.set_leaf_aabbs(frames[it % n_cpu_frames])
                b.refit_all()
                b.reinsertion_run(0.01, threads=threads)
                ts.append(time.perf_counter() - t0)
            cpu = {"value": n / float(np.mean(ts[1:])) / 1e6, "unit": "Mtris/s", "cores": threads, "kind": "port", "sample": "3 full frames"}
        print(json.dumps({
            "metric": "dynamic Bvh2 refit + reinsertion Mtris/s", "value": value, "unit": "Mtris/s", "n_gpus": world, "steps": n_timed,
            "warmup": args.warmup, "ms_per_step": refit + reins, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": desc, "tris": n, "l2": "flushed between steps (256 MB write)",
                                            "multi_gpu": "replicas only" if world > 1 else "single GPU"},
            "refit_ms": refit, "reinsertion_ms": reins, "reinsertions_applied_last_frame": opt.applied,
            "roofline": {"bound": "hbm", "achieved": refit_bytes / (refit * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": refit_bytes / (refit * 1e-3) / 1e9 / peak, "traffic": None, "kernel": "set_leaf_aabbs + refit_bottom_up_kernel",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": refit_bytes},
            "cpu_baseline": cpu,
            "e2e": {"value": world * n / e2e.item() / 1e6, "unit": "Mtris/s", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 0,
                    "how": "obvhs_cuda_bvh2_set_leaf_aabbs (pinned HOST AABBs) + obvhs_cuda_reinsertion_run, synchronised per frame"},
            "gpu_launches": launches, "clocks": clocks}))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def run_demoscene(args):
    """--workload demoscene: BASELINE.json configs[2] (examples/demoscene.rs): demoscene(1280, 570) = 3,276,800 triangles,
    medium_build, through BOTH tree types: build_bvh2_from_tris + Bvh2 traversal (what the example renders with) and
    build_cwbvh_from_tris + CwBvh traversal. Rays: the jittered / depth-of-field primary set of AA sample 0 (closest hit) and the
    sun shadow ray leaving every primary hit (`ray_traverse_miss`, demoscene.rs:185-193). `value` = CwBvh closest-hit Mrays/s."""
    import torch

    from obvhs_b200 import api, camera, test_util as tu
    from obvhs_b200.types import make_rays

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    width = 1280 if args.tris == 10_000_000 else int(round((args.tris / 2) ** 0.5))
    tris = tu.demoscene(width, 570)
    n_tris = tris.shape[0]
    cam = camera.demoscene_camera(1280)
    rays_all = camera.demoscene_primary(cam, 0)
    from obvhs_b200.sharding import shard_range

    lo, hi = shard_range(rays_all.shape[0], rank, world)
    rays = np.ascontiguousarray(rays_all[lo:hi])
    desc = f"demoscene({width},570) {n_tris} tris, medium_build, {cam.width}x{cam.height} primary rays + sun shadow rays, Bvh2 and CwBvh"
    sun = np.array([0.35, -0.1, 0.19], np.float32)
    sun = sun / np.float32(np.sqrt(np.float32((sun * sun).sum())))
    if args.impl == "reference":
        if rank != 0:
            return
        import oracle_bind as ob

        threads = len(os.sched_getaffinity(0))
        t0 = time.perf_counter()
        c = ob.build_cwbvh_from_tris(tris, "medium_build", threads=threads)
        build_s = time.perf_counter() - t0
        bt = c.bvh_tris(tris)
        sample = rays_all[:: max(1, rays_all.shape[0] // args.ref_rays)][: args.ref_rays]
        ts = []
        for it in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            c.ray_traverse(bt, sample, threads=threads)
            if it >= args.warmup:
                ts.append(time.perf_counter() - t0)
        v = sample.shape[0] / float(np.mean(ts)) / 1e6
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": float(np.mean(ts)) * 1e3, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": desc, "preset": "medium_build"},
                          "build": {"value": n_tris / build_s / 1e6, "unit": "Mtris/s", "ms": build_s * 1e3},
                          "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": threads, "kind": "port",
                                           "sample": f"{sample.shape[0]} of {rays_all.shape[0]} primary rays (strided), full CwBvh build"},
                          "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")
    stream = torch.cuda.Stream(device=dev)
    ctx = api.Context(local_rank, stream=stream.cuda_stream)
    params = api.BvhBuildParams.medium_build()
    n_rays = rays.shape[0]
    with torch.cuda.stream(stream):
        d_tris = torch.from_numpy(tris).to(dev)
        d_rays = torch.from_numpy(rays).to(dev)
        d_hits = torch.empty((n_rays, 4), dtype=torch.int32, device=dev)
        d_miss = torch.empty(n_rays, dtype=torch.uint8, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        d_counters = torch.zeros(2, dtype=torch.int64, device=dev)
        # shadow rays from the CwBvh primary hits (replicated build on every rank: no collective on this path)
        cw = api.build_cwbvh_from_tris(d_tris, params, ctx=ctx)
        cw.ray_traverse(d_rays, out=d_hits)
    stream.synchronize()
    t = d_hits[:, 3].view(torch.float32).cpu().numpy()
    hit = t < np.float32(3.0e38)
    o, d = rays[hit, 0:3], rays[hit, 4:7]
    hit_p = o + d * t[hit, None] - d * np.float32(0.01)
    shadow = make_rays(hit_p.astype(np.float32), np.tile(-sun, (hit_p.shape[0], 1)).astype(np.float32), 0.0, np.inf)
    n_shadow = shadow.shape[0]
    with torch.cuda.stream(stream):
        d_shadow = torch.from_numpy(shadow).to(dev)
    stream.synchronize()

    def timed(fn):
        with torch.cuda.stream(stream):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            r = fn()
            e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1), r

    res = {}
    launches0 = None
    sampler = ClockSampler(local_rank)
    for tree in ("cwbvh", "bvh2"):
        build = (lambda: api.build_cwbvh_from_tris(d_tris, params, ctx=ctx)) if tree == "cwbvh" else (lambda: api.build_bvh2_from_tris(d_tris, params, ctx=ctx))
        b_ms, p_ms, s_ms = [], [], []
        keep = None
        for it in range(args.warmup + args.steps):
            if it == args.warmup and tree == "cwbvh":
                if dist:
                    dist.barrier()
                torch.cuda.synchronize()
                if rank == 0:
                    sampler.start()
                launches0 = ctx.launch_count
            tb, bvh = timed(build)
            tp, _ = timed(lambda: bvh.ray_traverse(d_rays, out=d_hits))
            tsd, _ = timed(lambda: bvh.ray_traverse_miss(d_shadow, out=d_miss[:n_shadow]))
            keep = bvh  # the previous tree stays alive while the next one is built (result cache holds both sets)
            if it >= args.warmup:
                b_ms.append(tb)
                p_ms.append(tp)
                s_ms.append(tsd)
        tot = torch.tensor([sum(b_ms), sum(p_ms), sum(s_ms)], dtype=torch.float64, device=dev)
        if dist:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        bm, pm, sm_ = (tot / args.steps).tolist()
        cnt = torch.tensor([n_rays, n_shadow], dtype=torch.int64, device=dev)
        if dist:
            dist.all_reduce(cnt)
        nr, ns = cnt.tolist()
        res[tree] = {"build_ms": bm, "build_mtris_per_s": n_tris / (bm * 1e-3) / 1e6, "primary_ms": pm, "primary_mrays_per_s": nr / (pm * 1e-3) / 1e6,
                     "shadow_ms": sm_, "shadow_mrays_per_s": ns / (sm_ * 1e-3) / 1e6, "unoccluded": int(d_miss[:n_shadow].sum().item())}
        if tree == "cwbvh":
            with torch.cuda.stream(stream):
                d_counters.zero_()
                keep.ray_traverse(d_rays, out=d_hits, counters=d_counters)
            stream.synchronize()
            nodes_visited, tris_tested = [int(x) for x in d_counters.tolist()]
            cw_keep = keep
    torch.cuda.synchronize()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    # e2e: CwBvh primary rays from pinned host memory, hits back to the host
    from obvhs_b200.types import RAY_HIT

    from obvhs_b200.types import ray_args_of

    h_rays = torch.from_numpy(rays).pin_memory()
    h_args = torch.from_numpy(ray_args_of(rays)).pin_memory()  # Ray::new arguments, 32 B per ray: constructor on the device
    h_hits = torch.empty((n_rays, 4), dtype=torch.int32).pin_memory()
    hits_np = h_hits.numpy().view(RAY_HIT).reshape(-1)
    e_t, e_s = [], []
    for src, acc in ((h_args, e_t), (h_rays, e_s)):
        for it in range(5):
            t0 = time.perf_counter()
            cw_keep.ray_traverse(src.numpy(), out=hits_np)
            if it >= 2:
                acc.append(time.perf_counter() - t0)
    e2e_both = torch.tensor([float(np.mean(e_t)), float(np.mean(e_s))], dtype=torch.float64, device=dev)
    if dist:
        dist.all_reduce(e2e_both, op=dist.ReduceOp.MAX)
    e2e, e2e_struct = e2e_both[0], e2e_both[1]
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        alg_bytes = n_rays * 48 + 80 * nodes_visited + 48 * tris_tested
        achieved = alg_bytes / (res["cwbvh"]["primary_ms"] * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            import oracle_bind as ob

            threads = len(os.sched_getaffinity(0))
            sample = rays[:: max(1, n_rays // 200_000)][:200_000]
            c = ob.build_cwbvh_from_tris(tris, "medium_build", threads=threads)
            bt = c.bvh_tris(tris)
            c.ray_traverse(bt, sample[:1000], threads=threads)
            t0 = time.perf_counter()
            c.ray_traverse(bt, sample, threads=threads)
            dt = time.perf_counter() - t0
            cpu = {"value": sample.shape[0] / dt / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "port",
                   "sample": f"{sample.shape[0]} of {n_rays} primary rays (strided), CwBvh; C++ restatement of the obvhs CPU path (OpenMP)",
                   "build_mtris_per_s": n_tris / c.core_build_seconds / 1e6}
        nr_all = res["cwbvh"]["primary_mrays_per_s"] * res["cwbvh"]["primary_ms"] * 1e3
        print(json.dumps({
            "metric": METRIC, "value": res["cwbvh"]["primary_mrays_per_s"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sum(res[t][k] for t in res for k in ("build_ms", "primary_ms", "shadow_ms")), "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "preset": "medium_build", "rays_per_gpu": n_rays, "shadow_rays_per_gpu": n_shadow, "tris": n_tris,
                       "l2": "flushed between timed calls (256 MB write)", "multi_gpu": "replicated build, rays sharded" if world > 1 else "single GPU"},
            "build": {"value": res["cwbvh"]["build_mtris_per_s"], "unit": "Mtris/s", "ms": res["cwbvh"]["build_ms"]},
            "cwbvh": res["cwbvh"], "bvh2": res["bvh2"],
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "traverse_persistent_kernel<CwTree, closest>", "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "nodes_visited": nodes_visited, "tris_tested": tris_tested},
            "cpu_baseline": cpu,
            "e2e": {"value": nr_all / e2e.item() / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 32 * n_rays, "d2h_bytes_per_step": 16 * n_rays,
                    "how": "obvhs_cuda_cwbvh_ray_new_traverse_batch (Ray::new arguments, 32 B per ray) with pinned HOST args / hits",
                    "ray_struct": {"value": nr_all / e2e_struct.item() / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": 64 * n_rays,
                                   "how": "obvhs_cuda_cwbvh_ray_traverse_batch over the 64-byte Ray array"}},
            "gpu_launches": launches, "clocks": clocks}))
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="s3", choices=["s3", "cornell", "kitchen", "soup", "terrain", "bounce", "dynamic", "demoscene", "passes"])
    ap.add_argument("--frames", type=int, default=100, help="distinct frames of the dynamic workload (SURVEY.md 8d S4: 100)")
    ap.add_argument("--rays", type=int, default=S3_RAYS, help="size of the headline's global bounce-ray set")
    ap.add_argument("--parity-rays", type=int, default=262_144, help="rays of the in-run oracle comparison")
    ap.add_argument("--samples", type=int, default=24, help="AA samples of the bounce workload (165 = the 100M-ray set of SURVEY.md 8d)")
    ap.add_argument("--tris", type=int, default=10_000_000)
    ap.add_argument("--ref-rays", type=int, default=2_073_600, help="ray sample bound for the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.workload == "s3":
        run_headline_reference(args) if args.impl == "reference" else run_headline(args)
    elif args.workload == "cornell":
        run_cornell(args)
    elif args.workload == "dynamic":
        run_dynamic(args)
    elif args.workload == "demoscene":
        run_demoscene(args)
    elif args.workload == "passes":
        run_passes(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
