#!/usr/bin/env python
"""Tuning aid (one GPU): what does ONE slice of the sharded c2 step cost when a GPU runs it alone?

The c2 query set is cut into K slices as the sharded path does (iglb200_shard_run_dev, shard k of K, capped at the size of
its slice); each slice is run and timed on its own.  sum(slices) - (time of the whole set) is what K small launches lose to
ramp-up and tails with no second GPU involved; the difference between that and the per-rank times of a real K-GPU run is
what NVLink traffic and remote atomics cost.

    python tools/slice_probe.py [K=8] [faces=1000000] [grid=256]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import libigl_b200  # noqa: E402
from libigl_b200 import synth  # noqa: E402


def main():
    K = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    faces = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
    grid = int(sys.argv[3]) if len(sys.argv) > 3 else 256
    api = libigl_b200.lib()
    L = api.lib
    V, F = synth.torus_knot(*synth.knot_dims(faces))
    lo, hi = synth.bbox(V)
    P = synth.grid_points(lo, hi, grid, 1.1)
    n = P.shape[0]
    dev = torch.device("cuda", 0)
    dP = torch.from_numpy(np.ascontiguousarray(P.T)).to(dev)
    m = api.Mesh(V, F, 2)
    vp = C.c_void_p
    scratch = torch.empty(int(L.iglb200_shard_scratch_bytes(n)), dtype=torch.uint8, device=dev)
    out = [torch.empty(n, dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
           torch.empty((3, n), dtype=torch.float64, device=dev), torch.empty(n, dtype=torch.float64, device=dev)]
    po = [vp(t.data_ptr()) for t in out]
    tiny, huge = float(np.finfo(float).tiny), float(np.finfo(float).max)
    sup = int(L.iglb200_shard_super_rows(n, K))
    nsup = (n + sup - 1) // sup

    def run(mode, k, nk, cap, reps=5):
        ts = []
        for r in range(reps + 2):
            api.check(L.iglb200_shard_begin_dev(m.h, vp(dP.data_ptr()), n, vp(scratch.data_ptr()), 1, sup, None))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            api.check(L.iglb200_shard_run_dev(m.h, vp(dP.data_ptr()), n, vp(scratch.data_ptr()), mode, 4, tiny, huge, 2.0, 1, *po, sup, cap, k, nk, None))
            e1.record()
            torch.cuda.synchronize()
            if r >= 2:
                ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    if hasattr(L, "iglb200_debug_timeline"):   # a -DCP_TIMELINE build: when do the warps of the search kernel run dry?
        nb = 148 * 32
        tl = torch.zeros(8 * nb + 8, dtype=torch.int64, device=dev)
        L.iglb200_debug_timeline.argtypes = [vp]
        for k, nk, cap in ((0, 1, 0), (0, K, nsup // K), (K // 2, K, nsup // K)):
            api.check(L.iglb200_shard_begin_dev(m.h, vp(dP.data_ptr()), n, vp(scratch.data_ptr()), 1, sup, None))
            tl.zero_()
            assert L.iglb200_debug_timeline(vp(tl.data_ptr())) == 0
            torch.cuda.synchronize()
            api.check(L.iglb200_shard_run_dev(m.h, vp(dP.data_ptr()), n, vp(scratch.data_ptr()), 1, 4, tiny, huge, 2.0, 1, *po, sup, cap, k, nk, None))
            torch.cuda.synchronize()
            extra = tl.cpu().numpy()[8 * 2960:8 * 2960 + 5]
            t = tl.cpu().numpy()[:8 * 2960].reshape(2960, 8).astype(np.float64)
            t = t[t[:, 0] > 0]
            t0 = t[:, 0].min()
            first, last, exit_ = (t[:, 0] - t0) / 1e6, (t[:, 1] - t0) / 1e6, (t[:, 2] - t0) / 1e6
            end = exit_.max()
            busy = (exit_ - first).sum() / (len(t) * end)
            pct = lambda x: " ".join("%.2f" % v for v in np.percentile(x, [0, 10, 50, 90, 99, 100]))
            print(f"timeline shard {k}/{nk}: warps {len(t)}, kernel {end:.2f} ms, warp-time in use {100 * busy:.1f}%\n"
                  f"   first request at (ms, pct 0/10/50/90/99/100): {pct(first)}\n"
                  f"   last work taken at: {pct(last)}\n   exit at: {pct(exit_)}\n"
                  f"   duration of the last piece of work (exit - last taken): {pct(exit_ - last)}\n"
                  f"   longest walk of each warp: visits {pct(t[:, 3])}, ms {pct(t[:, 4] / 1e6)}; donations {int(t[:, 5].sum())}, tasks run {int(t[:, 6].sum())}\n"
                  f"   checks: {extra[2]}, with idle warps and stack < 2: {extra[0]}, with idle warps and stack >= 2: {extra[1]}, stack max {extra[3]} mean {extra[4] / max(extra[2], 1):.1f}\n"
                  f"   slowest walks (ms, visits, packet): " + " ".join("(%.2f, %d, %d)" % (t[j, 4] / 1e6, t[j, 3], t[j, 7]) for j in np.argsort(-t[:, 4])[:8]), flush=True)
        L.iglb200_debug_timeline(vp(0))
    for mode, name in ((1, "search"), (0, "search+winding")):
        whole = run(mode, 0, 1, 0)
        parts = [run(mode, k, K, nsup * (k + 1) // K - nsup * k // K) for k in range(K)]
        print(f"{name}: whole {whole:.2f} ms; {K} slices alone: {' '.join('%.2f' % t for t in parts)}; sum {sum(parts):.2f} "
              f"(+{100 * (sum(parts) / whole - 1):.0f}%), max {max(parts):.2f}, whole/{K} {whole / K:.2f}", flush=True)
    m.close()


if __name__ == "__main__":
    main()
