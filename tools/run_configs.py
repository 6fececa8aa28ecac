#!/usr/bin/env python
"""Runs every BASELINE.json config once on one GPU (device-resident timing, CUDA events) and
checks a sample of each against the oracle.  Not the driver's bench (that is bench.py = c2)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import libigl_b200  # noqa: E402
import oracle  # noqa: E402
import parity  # noqa: E402
from libigl_b200 import synth  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def dev_cols(P):
    return torch.from_numpy(np.ascontiguousarray(P.T)).cuda()


def run_sd(api, name, V, F, P, sample_stride, check=True):
    t0 = time.perf_counter()
    m = api.Mesh(V, F, 2)
    info = m.info()
    n = P.shape[0]
    dP = dev_cols(P)
    dS = torch.empty(n, dtype=torch.float64, device="cuda")
    dI = torch.empty(n, dtype=torch.int32, device="cuda")
    dC = torch.empty((3, n), dtype=torch.float64, device="cuda")
    dW = torch.empty(n, dtype=torch.float64, device="cuda")
    scr = torch.empty(int(api.lib.iglb200_query_scratch_bytes(n)), dtype=torch.uint8, device="cuda")
    lo, up = float(np.finfo(np.float64).tiny), float(np.finfo(np.float64).max)

    def step():
        api.check(api.lib.iglb200_signed_distance_dev(m.h, dP.data_ptr(), n, 4, lo, up, dS.data_ptr(), dI.data_ptr(),
                                                      dC.data_ptr(), dW.data_ptr(), None, scr.data_ptr(), 1, None))
    ms = timed(step)
    api.check(api.lib.iglb200_poll_device_error())
    out = dict(config=name, faces=int(F.shape[0]), queries=int(n), ms=ms, qps=n / ms * 1e3, build_ms=info["build_ms"],
               nodes=info["n_nodes"], levels=info["n_levels"], arena_mb=info["device_bytes"] / 1e6)
    if check:
        k = np.arange(0, n, sample_stride)
        S, I, C, W = dS.cpu().numpy()[k], dI.cpu().numpy()[k], dC.cpu().numpy().T[k], dW.cpu().numpy()[k]
        r = parity.check_signed_distance(V, F, P[k], S, I, C, W, what=name)
        out["parity_sample"] = r
    m.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c1,c2,c3,c4,c5")
    ap.add_argument("--small", action="store_true", help="reduced sizes (smoke)")
    a = ap.parse_args()
    api = libigl_b200.lib()
    res = []
    for c in a.configs.split(","):
        if c == "c1":
            V, F = synth.icosphere(4)
            res.append(run_sd(api, "c1 icosphere 5k faces, 64^3", V, F, synth.grid_points(*synth.bbox(V), 64, 1.5), 7))
        elif c == "c2":
            nf, g = (100_000, 96) if a.small else (1_000_000, 256)
            V, F = synth.torus_knot(*synth.knot_dims(nf))
            res.append(run_sd(api, f"c2 knot {nf} faces, {g}^3", V, F, synth.grid_points(*synth.bbox(V), g, 1.1), 211))
        elif c == "c3":
            lvl, nq = (6, 500_000) if a.small else (9, 10_000_000)
            V, F = synth.noisy_sphere(lvl, 0.01, 0)
            P = synth.random_points(nq, -1.5, 1.5, 1)
            m = api.Mesh(V, F, 0)
            n = P.shape[0]
            dP = dev_cols(P)
            dD = torch.empty(n, dtype=torch.float64, device="cuda")
            dI = torch.empty(n, dtype=torch.int32, device="cuda")
            dC = torch.empty((3, n), dtype=torch.float64, device="cuda")
            scr = torch.empty(int(api.lib.iglb200_query_scratch_bytes(n)), dtype=torch.uint8, device="cuda")
            ms = timed(lambda: api.check(api.lib.iglb200_squared_distance_dev(m.h, dP.data_ptr(), n, dD.data_ptr(), dI.data_ptr(),
                                                                             dC.data_ptr(), scr.data_ptr(), 1, None)))
            k = np.arange(0, n, 50)
            Do, Io, Co = oracle.PortAABB(V, F).squared_distance(P[k])
            ties = parity.check_closest(V, F, P[k], dD.cpu().numpy()[k], dI.cpu().numpy()[k], dC.cpu().numpy().T[k], Do, Io, Co, "c3")
            info = m.info()
            res.append(dict(config=f"c3 noisy sphere {F.shape[0]} faces, {n} random (squared_distance)", faces=int(F.shape[0]), queries=n,
                            ms=ms, qps=n / ms * 1e3, build_ms=info["build_ms"], nodes=info["n_nodes"], levels=info["n_levels"],
                            parity_sample=dict(n=len(k), ties=ties)))
            m.close()
        elif c == "c4":
            npts, nq = (100_000, 500_000) if a.small else (2_000_000, 20_000_000)
            V, F = synth.icosphere(7)
            Pc, N, A = synth.oriented_cloud(V, F, npts, 2)
            Q = synth.random_points(nq, -1.5, 1.5, 2)
            t0 = time.perf_counter()
            cl = api.Cloud(Pc, N, A, 2)
            torch.cuda.synchronize()
            tb = time.perf_counter() - t0
            dQ = dev_cols(Q)
            dW = torch.empty(nq, dtype=torch.float64, device="cuda")
            api.profile_enable(True)
            ms = timed(lambda: api.check(api.lib.iglb200_cloud_eval_dev(cl.h, dQ.data_ptr(), nq, 2.0, dW.data_ptr(), None)), reps=2)
            spans = {k: v / 3 for k, v in api.profile_collect().items()}
            api.profile_enable(False)
            k = np.arange(0, nq, max(1, nq // 20000))
            Wo = oracle.PortCloud(Pc, N, A, 2).eval(Q[k], 2.0)
            d = np.abs(dW.cpu().numpy()[k] - Wo).max()
            assert d < 1e-10, d
            res.append(dict(config=f"c4 cloud {npts} points, {nq} queries", points=npts, queries=nq, ms=ms, qps=nq / ms * 1e3,
                            kernel_ms=spans, build_s=tb, cells=cl.info()["n_cells"], levels=cl.info()["n_levels"], parity_sample=dict(n=len(k), dW=float(d))))
            cl.close()
        elif c == "c5":
            nf, g = (400_000, 128) if a.small else (10_000_000, 512)
            V, F = synth.torus_knot(*synth.knot_dims(nf))
            res.append(run_sd(api, f"c5 knot {nf} faces, {g}^3", V, F, synth.grid_points(*synth.bbox(V), g, 1.1), 4099))
        print(json.dumps(res[-1]), flush=True)


if __name__ == "__main__":
    main()
