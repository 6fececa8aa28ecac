#!/usr/bin/env python
"""Benchmark of the hot path: igl::signed_distance (FAST_WINDING_NUMBER sign)
on BASELINE.json config c2 — 256^3 grid queries against a 1 M-face torus-knot
mesh — as queries/second.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over the whole batch of queries.
  value : device-resident throughput (queries already in HBM), CUDA events on
          the launching stream, max over ranks.
  e2e   : the same metric through the host-pointer C ABI call a libigl user
          makes (pinned host buffers, H2D + kernels + D2H inside the timed region).
  roofline : the dominant kernel against the measured HBM peak (the path is
          latency/FP32 bound, not HBM bound — the fp32 fraction is reported too).
  cpu_baseline : the oracle (genuine reference FWN core + restated f64 AABB)
          under igl::parallel_for semantics on this box's host cores, bounded sample.
Under torchrun (N>1) rank 0 builds the trees, the arena is broadcast with NCCL,
every rank answers its own 256^3 grid (weak scaling), rank 0 prints one JSON line.
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

from libigl_b200 import synth  # noqa: E402

METRIC = "signed_distance_fwn_queries_per_sec"
UNIT = "queries/s"
BYTES_PER_QUERY = 36          # P 12 B in; S 4 + I 4 + C 12 + W 4 out (fp32 SoA accounting)
FLOP_NODE_FWN, FLOP_TRI_FWN, FLOP_NODE_CP, FLOP_TRI_CP = 372, 100, 50, 85


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


def workload(n_faces, grid, rank=0):
    V, F = synth.torus_knot(*synth.knot_dims(n_faces))
    if os.environ.get("IGLB200_BENCH_SHUFFLE_FACES"):   # tuning aid: same mesh, rows of F in random order
        F = np.ascontiguousarray(F[np.random.default_rng(0).permutation(F.shape[0])])
    lo, hi = synth.bbox(V)
    # every rank answers a different grid of the same size (weak scaling): scale 1.1, 1.11, ...
    P = synth.grid_points(lo, hi, grid, 1.1 + 0.01 * rank)
    return V, F, P


def cpu_sample(P, grid, target_queries):
    """A Morton-compact sample of the workload: whole z-slabs from the middle of the grid."""
    per = grid * grid
    nsl = max(1, min(grid, int(round(target_queries / per))))
    z0 = (grid - nsl) // 2
    return P[z0 * per:(z0 + nsl) * per], f"{nsl} central z-slabs of the {grid}^3 grid = {nsl * per} queries"


def run_cpu(V, F, P, grid, budget_s, steps=1, warmup=0):
    """oracle = genuine reference FWN (oracle/_ref) + restated f64 AABB (port), timed like
    signed_distance.cpp:90-138 does the work: one parallel_for pass per structure."""
    import oracle
    t0 = time.perf_counter()
    tree = oracle.PortAABB(V, F)
    t_tree = time.perf_counter() - t0
    t0 = time.perf_counter()
    fwn = oracle.RefFwn(V, F, 2)
    t_fwn = time.perf_counter() - t0
    # calibrate on ~1/64 of a slab budget
    probe, _ = cpu_sample(P, grid, grid * grid)
    t0 = time.perf_counter()
    oracle.signed_distance_fwn(probe, V, F, tree=tree, fwn=fwn)
    rate = probe.shape[0] / (time.perf_counter() - t0)
    sample, desc = cpu_sample(P, grid, rate * budget_s)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        _, _, _, _, cnt = cpu_pass(oracle, sample, V, F, tree, fwn)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    return dict(value=sample.shape[0] / t, ms_per_step=t * 1e3, sample=desc, n=sample.shape[0], cores=oracle.num_threads(),
                build_s=dict(aabb=t_tree, fwn=t_fwn), counters=cnt)


def cpu_pass(oracle, sample, V, F, tree, fwn):
    sqrD, I, C = tree.squared_distance(sample)
    w = fwn.winding(sample).astype(np.float32).astype(np.float64)
    S = (1.0 - 2.0 * np.abs(w)) * np.sqrt(sqrD)
    return S, I, C, w, None


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    V, F, P = workload(args.faces, args.grid, 0)
    r = run_cpu(V, F, P, args.grid, budget_s=args.cpu_seconds, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
        "config": config_dict(args, F.shape[0], P.shape[0]),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": r["sample"] + "; genuine UT_SolidAngle (oracle/_ref) + restated f64 igl::AABB (oracle port)",
                         "build_s": r["build_s"]},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def config_dict(args, nF, nQ):
    return {"workload": f"c2: igl::signed_distance(FAST_WINDING_NUMBER) {args.grid}^3 grid vs {nF}-face torus-knot mesh",
            "faces": int(nF), "queries_per_gpu": int(nQ), "grid": args.grid,
            "l2": "inputs+outputs per step (1.1 GB) exceed the 126 MB L2; no explicit flush",
            "sort_queries": int(args.sort)}


def ours(args):
    import torch
    import torch.distributed as dist

    import libigl_b200

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    api = libigl_b200.lib()
    api.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    V, F, P = workload(args.faces, args.grid, rank)
    nq = P.shape[0]

    # --- trees: built once on rank 0, replicated with one NCCL broadcast of the flat arena
    bcast_ms = 0.0
    keep = None
    if world == 1:
        mesh = api.Mesh(V, F, 2)
    else:
        from libigl_b200.dist import broadcast_mesh
        mesh = api.Mesh(V, F, 2) if rank == 0 else None
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        mesh, keep = broadcast_mesh(api, mesh, src=0, device=dev)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    info = mesh.info()

    # --- device-resident buffers (canonical layout: column-major doubles)
    Pcm = np.ascontiguousarray(P.T)                      # [3, n] == column-major n x 3
    hP = torch.from_numpy(Pcm).pin_memory()
    dP = hP.to(dev, non_blocking=True)
    dS = torch.empty(nq, dtype=torch.float64, device=dev)
    dI = torch.empty(nq, dtype=torch.int32, device=dev)
    dC = torch.empty((3, nq), dtype=torch.float64, device=dev)
    dW = torch.empty(nq, dtype=torch.float64, device=dev)
    scratch = torch.empty(int(api.lib.iglb200_query_scratch_bytes(nq)), dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(device=dev)
    lower, upper = float(np.finfo(np.float64).tiny), float(np.finfo(np.float64).max)

    def step():
        api.check(api.lib.iglb200_signed_distance_dev(mesh.h, dP.data_ptr(), nq, 4, lower, upper, dS.data_ptr(),
                                                      dI.data_ptr(), dC.data_ptr(), dW.data_ptr(), None, scratch.data_ptr(),
                                                      int(args.sort), stream.cuda_stream))

    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        for _ in range(args.warmup):
            step()
    torch.cuda.synchronize()
    launches_per_step = api.last_launch_count()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    api.profile_enable(True)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(args.steps):
            step()
        e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    api.check(api.lib.iglb200_poll_device_error())
    prof = api.profile_collect()
    api.profile_enable(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * nq * args.steps / (ms * 1e-3)

    # --- end to end through the host-pointer C ABI (what igl::signed_distance's caller sees)
    hS = torch.empty(nq, dtype=torch.float64).pin_memory()
    hI = torch.empty(nq, dtype=torch.int32).pin_memory()
    hC = torch.empty((3, nq), dtype=torch.float64).pin_memory()

    def e2e_step():
        # exactly the outputs of igl::signed_distance(P,V,F,type,S,I,C,N): S, I, C (N is not written for this sign type;
        # the winding number itself is an extension output and is not asked for)
        api.check(api.lib.iglb200_signed_distance(mesh.h, hP.data_ptr(), 1, nq, 1, nq, 4, lower, upper,
                                                  hS.data_ptr(), 1, hI.data_ptr(), hC.data_ptr(), 1, 1, nq,
                                                  None, 1, None, 1, 1, nq))

    e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = world * nq * args.e2e_steps / t_e2e
    # device and host paths must agree bit for bit
    assert torch.equal(hS, dS.cpu()) or bool(torch.isnan(hS).any()), "device-resident and host-pointer paths differ"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # --- roofline of the dominant kernel
    hbm, which, sm_max = peaks()
    dom = max(prof, key=prof.get) if prof else None
    kernel_ms = {k: v / args.steps for k, v in prof.items()}
    roof = None
    if dom:
        t_dom = kernel_ms[dom] * 1e-3
        alg_bytes = BYTES_PER_QUERY * nq + info["device_bytes"]   # per launch: query I/O + one read of the tree
        ach = alg_bytes / t_dom / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_r1.json")
        if dom == "closest" and os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("queries") == nq and tj.get("faces") == int(F.shape[0]):
                traffic = tj["dram_bytes_per_launch"]   # ncu --set full capture of this kernel on this workload
        roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
                "traffic": traffic, "peak_source": which, "algorithmic_bytes_per_launch": int(alg_bytes),
                "launch_ms": kernel_ms[dom],
                "note": "not HBM bound: arithmetic intensity ~1e3 flop/B; see fp32 below and DESIGN.md"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 distance + f32 winding number", "data": "synthetic",
        "config": config_dict(args, F.shape[0], nq),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(24 * nq), "d2h_bytes_per_step": int(36 * nq),
                "steps": args.e2e_steps},
        "gpu_launches": int(launches_per_step * args.steps),
        "kernel_ms_per_step": kernel_ms,
        "roofline": roof,
        "clocks": clocks,
        "tree": {"build_ms": info["build_ms"], "nodes": info["n_nodes"], "levels": info["n_levels"],
                 "device_bytes": info["device_bytes"], "broadcast_ms": bcast_ms},
    }
    if world == 1 and not args.no_cpu:
        r = run_cpu(V, F, P, args.grid, budget_s=args.cpu_seconds)
        line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": r["sample"] + "; genuine UT_SolidAngle (oracle/_ref) + restated f64 igl::AABB (oracle port)",
                                "build_s": r["build_s"]}
        # algorithmic flops from the instrumented oracle on a small sample
        try:
            import oracle
            samp, _ = cpu_sample(P, args.grid, 200000)
            tree = oracle.PortAABB(V, F)
            _, _, _, cnt = tree.squared_distance(samp, count=True)
            n_cp_node, n_cp_tri = cnt[0] / samp.shape[0], cnt[1] / samp.shape[0]
            vc = mesh.visit_counts(samp[:65536])   # the winding walk is the reference's own: same nodes, same faces
            n_fwn_node, n_fwn_tri = vc["fwn_nodes"], vc["fwn_tris"]
            flops_q = FLOP_NODE_FWN * n_fwn_node + FLOP_TRI_FWN * n_fwn_tri + FLOP_NODE_CP * n_cp_node + FLOP_TRI_CP * n_cp_tri
            measured = float(api.lib.iglb200_debug_fp32_peak_tflops())
            fp32_peak = measured * 1e12 if measured > 0 else 2 * 128 * 148 * sm_max * 1e6
            line["fp32"] = {"flops_per_query": flops_q, "achieved_tflops": flops_q * value / 1e12,
                            "peak_tflops": fp32_peak / 1e12, "frac": flops_q * value / fp32_peak,
                            "counts": {"cp_nodes_per_query_reference_tree": n_cp_node, "cp_tris_per_query": n_cp_tri,
                                       "fwn_nodes_per_query": n_fwn_node, "fwn_tris_per_query": n_fwn_tri,
                                       "ours_cp_nodes_per_warp": vc["cp_nodes_per_warp"], "ours_cp_faces_per_warp": vc["cp_tris_per_warp"]},
                            "peak_source": "FMA micro-benchmark on this GPU" if measured > 0 else "2*128*148*sm_max",
                            "note": "work of the reference algorithm (binary AABB, 1 tri/leaf) per SURVEY §8(d)"}
        except Exception as ex:  # the flop estimate is informative only
            line["fp32"] = {"error": str(ex)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--faces", type=int, default=1_000_000)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--sort", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
