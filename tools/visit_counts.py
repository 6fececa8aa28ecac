"""Print the traversal work counters of the c2 workload (debug aid for kernel tuning)."""
import sys, numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import libigl_b200 as ig
from libigl_b200 import synth

def main(nf=1_000_000, grid=256, sample=65536):
    V, F = synth.torus_knot(*synth.knot_dims(nf))
    lo, hi = synth.bbox(V)
    from libigl_b200 import api as A
    mesh = A.Mesh(ig.lib(), V, F, 2)
    # a Morton-coherent block of the grid: 32 x 32 x 64 cells from the middle
    P = synth.grid_points(lo, hi, grid)
    P = P.reshape(grid, grid, grid, 3)
    for name, blk in (("middle", P[96:128, 96:128, 96:160]), ("corner", P[:32, :32, :64])):
        print(name, mesh.visit_counts(np.ascontiguousarray(blk.reshape(-1, 3))[:sample]))

if __name__ == "__main__":
    main()
