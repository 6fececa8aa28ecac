#!/usr/bin/env python
"""Benchmark of the hot path, one BASELINE.json config per run (default c2: igl::signed_distance with the
FAST_WINDING_NUMBER sign, 256^3 grid queries against a 1 M-face torus-knot mesh), as queries/second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c1|c2|c3|c4|c5] [--impl reference]

A "step" is one pass of the hot path over ONE FIXED query set (strong scaling: the set does not grow with N).
  value : device-resident throughput.  The query set and the outputs live in rank 0's HBM.  Per step: rank 0
          Morton-sorts once; every rank (one process per GPU, torchrun) takes super-chunks of sorted queries off a
          counter in rank 0's memory (atomics over NVLink), reads the queries from and writes S/I/C/W to rank 0's
          memory directly (CUDA-IPC-mapped peer memory: scatter and gather are the kernels' own loads and stores).
          Two tiny NCCL all-reduces per step order the streams.  CUDA events on the launching streams, max over ranks.
  e2e   : the same metric through the host-pointer C ABI call a libigl user makes, H2D + kernels + D2H inside the
          timed region, page-locked host arrays.  N > 1: the ranks call iglb200_signed_distance_coop on ONE query set in
          shared host memory, dealing chunks off one shared counter; each GPU DMA-writes its chunks into the shared
          S/I/C.  Extra keys: `pageable` (plain malloc'ed arrays, what Eigen::MatrixXd gives) and `single_process`
          (ONE process driving all N GPUs through iglb200_init(N) — the call a C++ user of the drop-in makes).
  roofline : the dominant kernel against the FP32 ceiling (executed work; both the theoretical and the
          micro-benchmarked peak), `roofline_hbm` the same kernel against the measured HBM peak (not the bound).
  cpu_baseline : the oracle (genuine reference FWN core + restated f64 AABB) under igl::parallel_for semantics on
          this box's host cores, bounded sample (N = 1 only).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from libigl_b200 import synth  # noqa: E402

UNIT = "queries/s"
# algorithmic HBM bytes per query (fp32 SoA accounting) and flops of the REFERENCE algorithm per unit of work
BYTES_PER_QUERY = {"sd": 36, "cp": 32, "cloud": 16}
FLOP_NODE_FWN, FLOP_TRI_FWN, FLOP_NODE_CP, FLOP_TRI_CP = 372, 100, 50, 85
# flops our search kernel EXECUTES per lane (DESIGN.md "Kernels"): a node visit = 4 child bounds, a face test = tri_dist_f32
OURS_FLOP_CP_NODE_LANE, OURS_FLOP_CP_FACE_LANE = 68, 76

CONFIGS = {
    "c1": dict(kind="sd", metric="signed_distance_fwn_queries_per_sec", mesh=("icosphere", 4), grid=64, scale=1.5,
               name="c1: igl::signed_distance(FAST_WINDING_NUMBER) 64^3 grid vs 4x-upsampled icosahedron sphere"),
    "c2": dict(kind="sd", metric="signed_distance_fwn_queries_per_sec", mesh=("knot", 1_000_000), grid=256, scale=1.1,
               name="c2: igl::signed_distance(FAST_WINDING_NUMBER) 256^3 grid vs 1M-face torus-knot mesh"),
    "c3": dict(kind="cp", metric="aabb_squared_distance_queries_per_sec", mesh=("noisy", 9), nq=10_000_000,
               name="c3: igl::AABB::squared_distance 10M random points vs 5M-face noisy sphere"),
    "c4": dict(kind="cloud", metric="cloud_fast_winding_number_queries_per_sec", npts=2_000_000, nq=20_000_000,
               name="c4: point-cloud fast_winding_number (order 2, beta 2), 2M oriented points, 20M queries"),
    "c5": dict(kind="sd", metric="signed_distance_fwn_queries_per_sec", mesh=("knot", 10_000_000), grid=512, scale=1.1,
               name="c5: igl::signed_distance(FAST_WINDING_NUMBER) 512^3 grid vs 10M-face torus-knot mesh, query-sharded"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback", 1965.0


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop_flag, self.t = index, [], False, None

    def _run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().splitlines():
                    self.rows.append([x.strip() for x in line.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def start(self):
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def stop(self):
        self.stop_flag = True
        if self.t:
            self.t.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workloads
def make_mesh(cfg, args):
    kind, arg = cfg["mesh"]
    if kind == "icosphere":
        return synth.icosphere(arg)
    if kind == "knot":
        V, F = synth.torus_knot(*synth.knot_dims(args.faces or arg))
        if os.environ.get("IGLB200_BENCH_SHUFFLE_FACES"):   # tuning aid: same mesh, rows of F in random order
            F = np.ascontiguousarray(F[np.random.default_rng(0).permutation(F.shape[0])])
        return V, F
    if kind == "noisy":
        return synth.noisy_sphere(arg if not args.faces else max(2, int(round(np.log(args.faces / 20.0) / np.log(4.0)))), 0.01, 0)
    raise ValueError(kind)


def make_queries(cfg, args, V):
    if cfg["kind"] == "sd":
        g = args.grid or cfg["grid"]
        lo, hi = synth.bbox(V)
        return synth.grid_points(lo, hi, g, cfg["scale"]), g
    return synth.random_points(args.queries or cfg["nq"], -1.5, 1.5, 1), 0


def make_cloud(cfg, args):
    V, F = synth.icosphere(7)
    Pc, N, A = synth.oriented_cloud(V, F, args.faces or cfg["npts"], 2)
    Q = synth.random_points(args.queries or cfg["nq"], -1.5, 1.5, 2)
    return Pc, N, A, Q


def config_dict(cfg, args, n_units, nq, grid, world):
    d = {"workload": cfg["name"], "config": args.config, "queries": int(nq), "n_gpus": world,
         "sharding": "one fixed query set; super-chunks of Morton-sorted queries dealt dynamically to the GPUs" if world > 1 else "single GPU",
         "l2": "inputs+outputs per step exceed the 126 MB L2 (c2: 1.1 GB); no explicit flush", "sort_queries": int(args.sort)}
    if cfg["kind"] == "cloud":
        d["points"] = int(n_units)
    else:
        d["faces"] = int(n_units)
    if grid:
        d["grid"] = int(grid)
    return d


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_sample(P, grid, target_queries):
    """A compact sample of the workload: whole z-slabs from the middle of a grid, a leading block of random queries."""
    if grid:
        per = grid * grid
        nsl = max(1, min(grid, int(round(target_queries / per))))
        z0 = (grid - nsl) // 2
        return P[z0 * per:(z0 + nsl) * per], f"{nsl} central z-slabs of the {grid}^3 grid = {nsl * per} queries"
    k = int(max(1000, min(P.shape[0], target_queries)))
    return P[:k], f"the first {k} of the {P.shape[0]} random queries"


def run_cpu(cfg, args, budget_s, steps=1, warmup=0):
    """kind sd: genuine reference FWN (oracle/_ref) + restated f64 AABB (port), one parallel_for pass per structure as
    signed_distance.cpp:90-138 does the work; cp: the restated AABB alone; cloud: the restated octree evaluation."""
    import oracle
    kind = cfg["kind"]
    build_s = {}
    if kind == "cloud":
        Pc, N, A, Q = make_cloud(cfg, args)
        t0 = time.perf_counter()
        cl = oracle.PortCloud(Pc, N, A, 2)
        build_s["octree+precompute"] = time.perf_counter() - t0
        fn = lambda X: cl.eval(X, 2.0)
        P, grid, n_units = Q, 0, Pc.shape[0]
        what = "restated igl::octree + fast_winding_number(P,N,A,...) (oracle port)"
    else:
        V, F = make_mesh(cfg, args)
        P, grid = make_queries(cfg, args, V)
        n_units = F.shape[0]
        t0 = time.perf_counter()
        tree = oracle.PortAABB(V, F)
        build_s["aabb"] = time.perf_counter() - t0
        if kind == "sd":
            t0 = time.perf_counter()
            fwn = oracle.RefFwn(V, F, 2)
            build_s["fwn"] = time.perf_counter() - t0

            def fn(X):
                sqrD, I, C = tree.squared_distance(X)
                w = fwn.winding(X).astype(np.float32).astype(np.float64)
                return (1.0 - 2.0 * np.abs(w)) * np.sqrt(sqrD)
            what = "genuine UT_SolidAngle (oracle/_ref) + restated f64 igl::AABB (oracle port)"
        else:
            fn = lambda X: tree.squared_distance(X)
            what = "restated f64 igl::AABB (oracle port)"
    probe, _ = cpu_sample(P, grid, max(grid * grid, 20000))
    t0 = time.perf_counter()
    fn(probe)
    rate = probe.shape[0] / (time.perf_counter() - t0)
    sample, desc = cpu_sample(P, grid, rate * budget_s)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        fn(sample)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    return dict(value=sample.shape[0] / t, ms_per_step=t * 1e3, sample=desc + "; " + what, n=sample.shape[0],
                cores=oracle.num_threads(), build_s=build_s, n_units=n_units, nq=P.shape[0], grid=grid)


def reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cfg = CONFIGS[args.config]
    r = run_cpu(cfg, args, budget_s=args.cpu_seconds, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": cfg["metric"], "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64 distance + f32 winding number", "data": "synthetic",
        "config": config_dict(cfg, args, r["n_units"], r["nq"], r["grid"], args.gpus),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                         "build_s": r["build_s"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm
class Job:
    """torch.distributed plumbing of one run (NCCL for the device side, gloo for host-side barriers)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.host_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            self.host_group = dist.new_group(backend="gloo")
        self.tok = torch.zeros(1, dtype=torch.int32, device=self.dev)

    def stream_fence(self):
        """orders every rank's current stream after every other rank's (a 4-byte NCCL all-reduce on the stream)"""
        if self.world > 1:
            self.dist.all_reduce(self.tok)

    def host_barrier(self):
        if self.world > 1:
            self.dist.barrier(group=self.host_group)

    def bcast_obj(self, obj):
        if self.world == 1:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0, group=self.host_group)
        return box[0]

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather_floats(self, xs):
        if self.world == 1:
            return [list(xs)]
        out = [None] * self.world
        self.dist.all_gather_object(out, list(xs), group=self.host_group)
        return out

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def timed_steps(job, stream, step, warmup, steps):
    torch = job.torch
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            step()
    torch.cuda.synchronize()
    job.host_barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    with torch.cuda.stream(stream):
        job.stream_fence()
        e0.record(stream)
        for _ in range(steps):
            step()
        job.stream_fence()
        e1.record(stream)
    torch.cuda.synchronize()
    job.host_barrier()
    return job.max_over_ranks(e0.elapsed_time(e1))


def ours_mesh(args, cfg, job, light=False):
    """configs c1, c2, c3, c5: queries against a triangle mesh (signed distance or squared distance).
    light: the device-resident leg only (no e2e, no CPU arm, no rooflines).  Returns the result on rank 0."""
    import libigl_b200
    torch = job.torch
    steps = min(args.steps, 3) if light else args.steps
    warmup = min(args.warmup, 2) if light else args.warmup
    rank, world, dev = job.rank, job.world, job.dev
    api = libigl_b200.lib()
    api.set_device(job.local)
    L = api.lib
    vp = ctypes.c_void_p
    mode = 0 if cfg["kind"] == "sd" else 1
    sign = 4
    lower, upper = float(np.finfo(np.float64).tiny), float(np.finfo(np.float64).max)

    # --- the workload exists on rank 0 only: ONE query set
    V = F = P = None
    grid = 0
    if rank == 0:
        V, F = make_mesh(cfg, args)
        P, grid = make_queries(cfg, args, V)
    nq, nF, grid = job.bcast_obj((P.shape[0], F.shape[0], grid) if rank == 0 else None)

    # --- trees: built once on rank 0, replicated with one NCCL broadcast of the flat arena
    bcast_ms, keep = 0.0, None
    mesh = api.Mesh(V, F, 2) if rank == 0 else None
    if world > 1:
        from libigl_b200.dist import broadcast_mesh
        torch.cuda.synchronize()
        job.host_barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        mesh, keep = broadcast_mesh(api, mesh, src=0, device=dev)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    info = mesh.info()

    # --- rank 0 owns the device-resident query set and the outputs: ONE allocation [scratch | P | S | I | C | W] that the
    #     other ranks map with CUDA IPC (their kernels read P and write S/I/C/W in rank 0's HBM over NVLink)
    al = lambda x: (x + 255) & ~255
    sbytes = al(int(L.iglb200_shard_scratch_bytes(nq)))
    offP = sbytes
    offS = offP + al(24 * nq)
    offI = offS + al(8 * nq)
    offC = offI + al(4 * nq)
    offW = offC + al(24 * nq)
    total = offW + al(8 * nq)
    base = vp()
    handle = ctypes.create_string_buffer(64)
    if rank == 0:
        api.check(L.iglb200_ipc_alloc(total, ctypes.byref(base), handle))
    hb = job.bcast_obj(bytes(handle.raw) if rank == 0 else None)
    if rank != 0:
        api.check(L.iglb200_ipc_open(hb, ctypes.byref(base)))
    b0 = base.value
    scratch, pP, pS, pI, pC, pW = vp(b0), vp(b0 + offP), vp(b0 + offS), vp(b0 + offI), vp(b0 + offC), vp(b0 + offW)
    super_rows = int(args.super_rows or L.iglb200_shard_super_rows(nq, world))
    dP = dS = dI = dC = dW = None
    hP = None
    if rank == 0:
        from libigl_b200.dist import _arena_tensor
        whole = _arena_tensor(api, b0, total, dev)                 # zero-copy torch view of the IPC allocation
        hP = torch.from_numpy(np.ascontiguousarray(P.T)).pin_memory()      # [3, n] == column-major n x 3
        dP = whole[offP:offP + 24 * nq].view(torch.float64).view(3, nq)
        dP.copy_(hP, non_blocking=True)
        dS = whole[offS:offS + 8 * nq].view(torch.float64)
        dI = whole[offI:offI + 4 * nq].view(torch.int32)
        dC = whole[offC:offC + 24 * nq].view(torch.float64).view(3, nq)
        dW = whole[offW:offW + 8 * nq].view(torch.float64)
    stream = torch.cuda.Stream(device=dev)
    sp = lambda: vp(stream.cuda_stream)
    launches = [0]

    def step():
        n_l = 0
        if rank == 0:
            api.check(L.iglb200_shard_begin_dev(mesh.h, pP, nq, scratch, int(args.sort), super_rows, sp()))
            n_l += api.last_launch_count()
        job.stream_fence()
        api.check(L.iglb200_shard_run_dev(mesh.h, pP, nq, scratch, mode, sign, lower, upper, 2.0, int(args.sort), pS, pI, pC, pW,
                                          super_rows, 0, rank, world, sp()))
        n_l += api.last_launch_count()
        job.stream_fence()
        if rank == 0:
            api.check(L.iglb200_shard_finish_dev(mesh.h, nq, scratch, mode, sign, lower, upper, int(args.sort), pS, pI, pC, pW, sp()))
            n_l += api.last_launch_count()
        launches[0] = n_l

    # warm-up, then the timed region
    timed_steps(job, stream, step, warmup, 0)
    sampler = ClockSampler(job.local)
    if rank == 0:
        sampler.start()
    api.profile_enable(True)
    ms = timed_steps(job, stream, step, 0, steps)
    api.check(L.iglb200_poll_device_error())
    prof = api.profile_collect()
    api.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    value = nq * steps / (ms * 1e-3)
    kernel_ms = {k: v / steps for k, v in prof.items()}
    per_rank = job.gather_floats([kernel_ms.get("closest", 0.0), kernel_ms.get("winding", 0.0)])

    # N = 1: the fused single-GPU entry point on the same data must give the same bits (and its time is reported)
    fused_ms = None
    if world == 1 and not light:
        qs = torch.empty(int(L.iglb200_query_scratch_bytes(nq)), dtype=torch.uint8, device=dev)
        S2, I2, C2, W2 = torch.empty_like(dS), torch.empty_like(dI), torch.empty_like(dC), torch.empty_like(dW)

        def fused():
            if mode == 0:
                api.check(L.iglb200_signed_distance_dev(mesh.h, dP.data_ptr(), nq, sign, lower, upper, S2.data_ptr(), I2.data_ptr(),
                                                        C2.data_ptr(), W2.data_ptr(), None, qs.data_ptr(), int(args.sort), sp()))
            else:
                api.check(L.iglb200_squared_distance_dev(mesh.h, dP.data_ptr(), nq, S2.data_ptr(), I2.data_ptr(), C2.data_ptr(),
                                                         qs.data_ptr(), int(args.sort), sp()))
        fused_ms = timed_steps(job, stream, fused, 1, max(2, args.steps // 2)) / max(2, args.steps // 2)
        same = torch.equal(dI, I2) and torch.equal(dC, C2) and bool(((dS == S2) | (torch.isnan(dS) & torch.isnan(S2))).all())
        assert same, "sharded and fused device paths differ"
        del qs, S2, I2, C2, W2

    # --- end to end through the host-pointer C ABI
    e2e = None if light else run_e2e_mesh(job, api, mesh, cfg, args, hP, nq, mode, sign, lower, upper, dS, dI)

    line = None
    if rank == 0:
        line = {
            "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 distance + f32 winding number" if mode == 0 else "f64", "data": "synthetic",
            "config": config_dict(cfg, args, nF, nq, grid, world),
            "e2e": e2e,
            "gpu_launches": int(launches[0] * steps),
            "kernel_ms_per_step": kernel_ms,
            "per_rank_kernel_ms": {"closest": [r[0] for r in per_rank], "winding": [r[1] for r in per_rank]},
            "super_rows": super_rows,
            "clocks": clocks,
            "tree": {"build_ms": info["build_ms"], "nodes": info["n_nodes"], "levels": info["n_levels"],
                     "device_bytes": info["device_bytes"], "broadcast_ms": bcast_ms},
        }
        if fused_ms is not None:
            line["fused_single_call_ms_per_step"] = fused_ms
        if not light:
            add_rooflines(line, api, mesh, cfg, args, kernel_ms, info, nq, value, P, grid, world)
        if world == 1 and not args.no_cpu and not light:
            r = run_cpu(cfg, args, budget_s=args.cpu_seconds)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                    "build_s": r["build_s"]}
    torch.cuda.synchronize()
    job.host_barrier()
    if rank != 0:
        L.iglb200_ipc_close(base)
    job.host_barrier()
    if rank == 0:
        del dP, dS, dI, dC, dW, whole
        L.iglb200_ipc_free(base)
    mesh.close()
    del keep
    torch.cuda.empty_cache()
    return line


def run_e2e_mesh(job, api, mesh, cfg, args, hP, nq, mode, sign, lower, upper, dS, dI):
    """Host-pointer C ABI, H2D + kernels + D2H inside the timed region (wall clock between host barriers, max over ranks)."""
    torch = job.torch
    L = api.lib
    vp = ctypes.c_void_p
    rank, world = job.rank, job.world
    steps = args.e2e_steps or args.steps
    out = {"unit": UNIT, "h2d_bytes_per_step": int(24 * nq), "d2h_bytes_per_step": int(36 * nq), "steps": steps}

    def call(Pp, Sp, Ip, Cp, coop=None):
        # exactly the outputs of igl::signed_distance(P,V,F,type,S,I,C,N): S, I, C (N is not written for this sign type)
        if mode == 0 and coop is None:
            api.check(L.iglb200_signed_distance(mesh.h, Pp, 1, nq, 1, nq, sign, lower, upper, Sp, 1, Ip, Cp, 1, 1, nq,
                                                None, 1, None, 1, 1, nq))
        elif mode == 0:
            api.check(L.iglb200_signed_distance_coop(mesh.h, Pp, 1, nq, 1, nq, sign, lower, upper, Sp, 1, Ip, Cp, 1, 1, nq,
                                                     None, 1, None, 1, 1, nq, coop[0], coop[1]))
        else:
            api.check(L.iglb200_squared_distance(mesh.h, Pp, 1, nq, 1, nq, Sp, 1, Ip, Cp, 1, 1, nq))

    def timed(fn, k):
        fn(0)
        job.host_barrier()
        t0 = time.perf_counter()
        for i in range(k):
            fn(i + 1)
        torch.cuda.synchronize()
        job.host_barrier()
        return job.max_over_ranks(time.perf_counter() - t0)

    if world == 1:
        hS = torch.empty(nq, dtype=torch.float64).pin_memory()
        hI = torch.empty(nq, dtype=torch.int32).pin_memory()
        hC = torch.empty((3, nq), dtype=torch.float64).pin_memory()
        t = timed(lambda i: call(hP.data_ptr(), hS.data_ptr(), hI.data_ptr(), hC.data_ptr()), steps)
        out["value"] = nq * steps / t
        out["host_memory"] = "page-locked"
        # device and host paths must agree bit for bit
        assert torch.equal(hI, dI.cpu()), "device-resident and host-pointer paths differ (I)"
        a, b = hS, dS.cpu()
        assert bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all()), "device-resident and host-pointer paths differ (S)"
        # the same call from plain pageable arrays (what Eigen::MatrixXd hands over): staged through the pinned ring
        pP = np.array(hP.numpy(), copy=True)
        pS, pI, pC = np.empty(nq), np.empty(nq, np.int32), np.empty((3, nq))
        t = timed(lambda i: call(pP.ctypes.data, pS.ctypes.data, pI.ctypes.data, pC.ctypes.data), steps)
        out["pageable"] = {"value": nq * steps / t, "unit": UNIT}
        assert np.array_equal(pI, hI.numpy()), "pageable and pinned host paths differ"
        return out

    if mode != 0:
        out["value"] = None
        out["note"] = "cooperative multi-process host calls exist for signed_distance only"
        return out
    # N > 1: ONE query set in shared host memory, page-locked by every rank; chunks dealt off one shared counter
    from libigl_b200.dist import SharedArrays
    need = 64 * nq + (1 << 20)
    st = os.statvfs("/dev/shm")
    if st.f_bavail * st.f_frsize < need * 1.1 and rank == 0:
        out["value"] = None
        out["note"] = f"/dev/shm too small for the shared query set ({need >> 20} MiB)"
    ok = job.bcast_obj("value" not in out if rank == 0 else None)
    if ok:
        name = f"iglb200_bench_{os.environ.get('MASTER_PORT', '0')}"
        sh = None
        if rank == 0:
            sh = SharedArrays(name, nq, create=True)
            sh.P[:] = hP.numpy()
            sh.counter[:] = 0
        job.host_barrier()
        if rank != 0:
            sh = SharedArrays(name, nq, create=False)
        sh.register(api)
        # every rank runs its own workers: enough chunks that all of them draw several
        chunk = int(args.chunk_rows or max(1 << 16, min(1 << 20, -(-nq // (world * 12)))))
        base = sh.counter.ctypes.data

        def coop_step(i):
            # ping-pong counters: rank 0 zeroes the one the NEXT step will use; one host barrier per step
            if rank == 0:
                sh.counter[(i + 1) & 1] = 0
            call(sh.P.ctypes.data, sh.S.ctypes.data, sh.I.ctypes.data, sh.C.ctypes.data, coop=(vp(base + 8 * (i & 1)), chunk))
            job.host_barrier()
        t = timed(coop_step, steps)
        out["value"] = nq * steps / t
        out["host_memory"] = "POSIX shared memory, page-locked by every rank (cudaHostRegister)"
        out["chunk_rows"] = chunk
        if rank == 0:
            assert np.array_equal(sh.I, dI.cpu().numpy()), "cooperative host path and device path differ"
        job.host_barrier()
        sh.close(unlink=(rank == 0))
    # ONE process driving all GPUs (iglb200_init(N)): measured by rank 0 while the other ranks wait on the host
    if not args.no_single_process:
        job.torch.cuda.synchronize()
        job.host_barrier()
        if rank == 0:
            try:
                import libigl_b200
                V, F = make_mesh(cfg, args)
                api.init(world)
                api.set_device(0)
                m2 = api.Mesh(V, F, 2)
                gi = api.group_info()
                hS = torch.empty(nq, dtype=torch.float64).pin_memory()
                hI = torch.empty(nq, dtype=torch.int32).pin_memory()
                hC = torch.empty((3, nq), dtype=torch.float64).pin_memory()

                def one(i):
                    api.check(L.iglb200_signed_distance(m2.h, hP.data_ptr(), 1, nq, 1, nq, sign, lower, upper, hS.data_ptr(), 1,
                                                        hI.data_ptr(), hC.data_ptr(), 1, 1, nq, None, 1, None, 1, 1, nq))
                one(0)
                t0 = time.perf_counter()
                for i in range(steps):
                    one(i)
                t = time.perf_counter() - t0
                assert torch.equal(hI, dI.cpu()), "single-process multi-GPU path and device path differ"
                out["single_process"] = {"value": nq * steps / t, "unit": UNIT, "devices": gi["devices"],
                                         "replication": gi["transport"], "replicate_ms": gi["replicate_ms"]}
                m2.close()
                api.init(1)
                api.set_device(0)
            except Exception as ex:   # reported, not fatal: the multi-process numbers above stand on their own
                out["single_process"] = {"error": str(ex)[:300]}
        job.host_barrier()
    return out


def add_rooflines(line, api, mesh, cfg, args, kernel_ms, info, nq, value, P, grid, world):
    hbm, which, sm_max = peaks()
    dom = max(kernel_ms, key=kernel_ms.get) if kernel_ms else None
    if not dom:
        return
    t_dom = kernel_ms[dom] * 1e-3
    # HBM: algorithmic bytes of one launch on THIS rank = its share of the query I/O + one read of the tree
    alg_bytes = BYTES_PER_QUERY[cfg["kind"]] * nq / world + info["device_bytes"]
    ach = alg_bytes / t_dom / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic_r2.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        if tj.get("kernel") == dom and tj.get("queries") == nq and tj.get("faces") == int(info["n_faces"]) and world == 1:
            traffic = tj["dram_bytes_per_launch"]   # ncu --set full capture of this kernel on this workload
    line["roofline_hbm"] = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                            "traffic": traffic, "peak_source": which, "algorithmic_bytes_per_launch": int(alg_bytes),
                            "launch_ms": kernel_ms[dom],
                            "note": "not the bound: arithmetic intensity ~1e3 flop/B (SURVEY §8d); reported for completeness"}
    # FP32: work the dominant kernel EXECUTES (its own node/face counters x flops per lane and unit)
    try:
        samp, _ = cpu_sample(P, grid, 65536)
        vc = mesh.visit_counts(samp[:65536])
        theo = 2 * 128 * 148 * sm_max * 1e6
        measured = float(api.lib.iglb200_debug_fp32_peak_tflops()) * 1e12
        if dom == "closest":
            flops_launch = (vc["cp_nodes_per_warp"] * OURS_FLOP_CP_NODE_LANE + vc["cp_tris_per_warp"] * OURS_FLOP_CP_FACE_LANE) * 32 * (nq / 32.0) / world
            src = "kernel's own visit counters (65 536-query sample) x flops per lane: node visit 68, face test 76"
        else:
            flops_launch = (vc["fwn_nodes"] * FLOP_NODE_FWN + vc["fwn_tris"] * FLOP_TRI_FWN) * nq / world
            src = "visit counters x 372 flop per node visit, 100 per exact triangle (the reference's own walk)"
        fj = os.path.join(ROOT, "profiles", "flops_r2.json")
        ncu_flops = ncu_note = None
        if os.path.exists(fj):
            j = json.load(open(fj))
            if j.get("kernel") == dom and j.get("queries") == nq and j.get("faces") == int(info["n_faces"]):
                # smsp__sass_thread_inst_executed_op_{ffma x2, fadd, fmul}_pred_on of the whole query set; a rank's launch does 1/N of it
                ncu_flops = j["fp32_flops_per_launch"] / world
                ncu_note = j.get("note")
        used = ncu_flops if ncu_flops else flops_launch
        ach_t = used / t_dom
        line["roofline"] = {"bound": "fp32", "kernel": dom, "achieved": ach_t / 1e12, "peak": theo / 1e12, "unit": "TFLOP/s",
                            "frac": ach_t / theo, "traffic": traffic,
                            "peak_source": "2 x 128 lanes x 148 SMs x sm_max_mhz", "peak_microbenchmark": measured / 1e12,
                            "frac_of_microbenchmark": ach_t / measured if measured > 0 else None,
                            "executed_flops_per_launch": used,
                            "numerator": "ncu sass op counters (profiles/flops_r2.json)" if ncu_flops else src,
                            "numerator_note": ncu_note if ncu_flops else None,
                            "launch_ms": kernel_ms[dom],
                            "counts": {"cp_nodes_per_packet": vc["cp_nodes_per_warp"], "cp_faces_per_packet": vc["cp_tris_per_warp"],
                                       "fwn_nodes_per_query": vc["fwn_nodes"], "fwn_tris_per_query": vc["fwn_tris"]}}
    except Exception as ex:
        line["roofline"] = {"bound": "fp32", "error": str(ex)[:300]}


def ours_cloud(args, cfg):
    """config c4: oriented point cloud.  Every rank builds the (deterministic) octree; N > 1: static contiguous slices of
    the Morton-uniform random queries, gathered to rank 0 with NCCL (the evaluation cost per query is uniform)."""
    import libigl_b200
    job = Job()
    torch, dist = job.torch, job.dist
    rank, world, dev = job.rank, job.world, job.dev
    api = libigl_b200.lib()
    api.set_device(job.local)
    L = api.lib
    Pc, N, A, Q = make_cloud(cfg, args)
    nq = Q.shape[0]
    t0 = time.perf_counter()
    cl = api.Cloud(Pc, N, A, 2)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    from libigl_b200.dist import shard_range
    lo, hi = shard_range(nq, rank, world)
    hQ = torch.from_numpy(np.ascontiguousarray(Q.T)).pin_memory()
    dQ = torch.from_numpy(np.ascontiguousarray(Q[lo:hi].T)).to(dev)
    dW = torch.empty(hi - lo, dtype=torch.float64, device=dev)
    sizes = [shard_range(nq, r, world) for r in range(world)]
    parts = [torch.empty(b - a, dtype=torch.float64, device=dev) for a, b in sizes] if (world > 1 and rank == 0) else None
    stream = torch.cuda.Stream(device=dev)

    def step():
        api.check(L.iglb200_cloud_eval_dev(cl.h, dQ.data_ptr(), hi - lo, 2.0, dW.data_ptr(), ctypes.c_void_p(stream.cuda_stream)))
        if world > 1:
            dist.gather(dW, parts, dst=0)
    timed_steps(job, stream, step, args.warmup, 0)
    sampler = ClockSampler(job.local)
    if rank == 0:
        sampler.start()
    api.profile_enable(True)
    ms = timed_steps(job, stream, step, 0, args.steps)
    prof = api.profile_collect()
    api.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    value = nq * args.steps / (ms * 1e-3)
    kernel_ms = {k: v / args.steps for k, v in prof.items()}
    e2e = None
    if world == 1:
        hW = torch.empty(nq, dtype=torch.float64).pin_memory()
        steps = args.e2e_steps or args.steps
        f = lambda: api.check(L.iglb200_cloud_eval(cl.h, hQ.data_ptr(), 1, nq, 1, nq, 2.0, hW.data_ptr(), 1))
        f()
        t0 = time.perf_counter()
        for _ in range(steps):
            f()
        t = time.perf_counter() - t0
        e2e = {"value": nq * steps / t, "unit": UNIT, "h2d_bytes_per_step": int(24 * nq), "d2h_bytes_per_step": int(8 * nq), "steps": steps}
        assert torch.equal(hW, dW.cpu())
    if rank == 0:
        info = cl.info()
        hbm, which, _ = peaks()
        dom = "cloud_eval"
        t_dom = kernel_ms.get(dom, ms / args.steps) * 1e-3
        alg = BYTES_PER_QUERY["cloud"] * nq / world + info["device_bytes"]
        line = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic", "config": config_dict(cfg, args, Pc.shape[0], nq, 0, world), "e2e": e2e,
                "gpu_launches": int(api.last_launch_count() * args.steps), "kernel_ms_per_step": kernel_ms, "clocks": clocks,
                "roofline_hbm": {"bound": "hbm", "kernel": dom, "achieved": alg / t_dom / 1e9, "peak": hbm, "unit": "GB/s",
                                 "frac": alg / t_dom / 1e9 / hbm, "traffic": None, "peak_source": which},
                "tree": {"build_s": build_s, "cells": info["n_cells"], "levels": info["n_levels"], "device_bytes": info["device_bytes"]}}
        if world == 1 and not args.no_cpu:
            r = run_cpu(cfg, args, budget_s=args.cpu_seconds)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                    "build_s": r["build_s"]}
        emit(line)
    job.host_barrier()
    cl.close()
    job.close()


_RESULT_FD = None


def keep_stdout_clean():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to fd 1 when the
    box sets NCCL_DEBUG), so fd 1 is pointed at stderr for the run and the result line goes to the real stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    text = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(text.decode()); sys.stdout.flush()
    else:
        os.write(_RESULT_FD, text)


def main():
    keep_stdout_clean()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--faces", type=int, default=0, help="override the mesh size of the config (tuning / smoke runs)")
    ap.add_argument("--grid", type=int, default=0, help="override the grid resolution of the config")
    ap.add_argument("--queries", type=int, default=0, help="override the number of random queries (c3, c4)")
    ap.add_argument("--sort", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=0, help="default: --steps")
    ap.add_argument("--super-rows", type=int, default=0)
    ap.add_argument("--chunk-rows", type=int, default=0)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-single-process", action="store_true")
    ap.add_argument("--no-c5", action="store_true", help="default run: skip the config-5 leg that rides along with c2")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    cfg = CONFIGS[args.config]
    if args.impl == "reference":
        reference_arm(args)
    elif cfg["kind"] == "cloud":
        ours_cloud(args, cfg)
    else:
        job = Job()
        line = ours_mesh(args, cfg, job)
        if args.config == "c2" and not args.no_c5 and not args.faces and not args.grid:
            # BASELINE config 5 rides along in the same JSON line: 512^3 vs 10 M faces, query-sharded over the same N GPUs
            # (device-resident leg only; `--config c5` runs it in full)
            try:
                c5 = ours_mesh(args, CONFIGS["c5"], job, light=True)
                if line is not None and c5 is not None:
                    line["c5"] = {k: c5[k] for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "config",
                                                     "kernel_ms_per_step", "per_rank_kernel_ms", "super_rows", "tree")}
            except Exception as ex:   # the headline line stands on its own
                if line is not None:
                    line["c5"] = {"error": str(ex)[:300]}
        if line is not None:
            emit(line)
        job.close()


if __name__ == "__main__":
    main()
