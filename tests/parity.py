"""Parity gates of SURVEY.md §8(d), shared by the CPU (emulated-kernel) and GPU tests.

The checker is the oracle: genuine reference FWN (oracle/_ref) + restated f64
AABB/Ericson (oracle/oracle_port.cpp).  Tolerances are north_star's:
  * I and sign(S) identical, except on exact ties (the two faces have bit-equal
    squared distance to the query) or where |S| < 1e-6 * bbox diagonal;
  * S, C, W within 1e-5 relative (W: absolute floor 1, S: floor 1e-6*diag);
  * NaN / sentinel rows identical.
"""
import numpy as np

import oracle

REL = 1e-5


def check_closest(V, F, P, D, I, C, Do, Io, Co, what=""):
    V = np.asarray(V, np.float64)
    diag = float(np.linalg.norm(V.max(0) - V.min(0)))
    mism = I != Io
    if mism.any():
        k = np.nonzero(mism)[0]
        d_ours = oracle.face_distance(V, F, P[k], I[k])
        d_ref = oracle.face_distance(V, F, P[k], Io[k])
        nontie = d_ours != d_ref
        assert not nontie.any(), f"{what}: {nontie.sum()} closest-face mismatches that are not exact ties"
    # the device evaluates Ericson in un-fused f64 in the oracle's order: distances are bit-identical
    assert np.array_equal(D, Do), f"{what}: squared distances differ, max |d| {np.abs(D - Do).max():.3e}"
    assert np.abs(C - Co).max(initial=0.0) <= REL * diag, f"{what}: closest points differ"
    same = ~mism
    assert np.array_equal(C[same], Co[same]), f"{what}: closest points not bit-identical on agreeing faces"
    return int(mism.sum())


def check_signed_distance(V, F, P, S, I, C, W, lower=None, upper=None, what="", fwn=None, tree=None):
    V = np.asarray(V, np.float64)
    F = np.asarray(F, np.int32)
    P = np.ascontiguousarray(P, np.float64)
    diag = float(np.linalg.norm(V.max(0) - V.min(0)))
    So, Io, Co, Wo = oracle.signed_distance_fwn(P, V, F, lower, upper, tree=tree, fwn=fwn)
    nan_o = np.isnan(So)
    assert np.array_equal(np.isnan(S), nan_o), f"{what}: NaN rows differ"
    assert np.array_equal(I[nan_o], Io[nan_o]), f"{what}: sentinel I differs"
    assert np.array_equal(C[nan_o], Co[nan_o]), f"{what}: sentinel C differs"
    ok = ~nan_o
    S, I, C, W, So, Io, Co, Wo, Pk = S[ok], I[ok], C[ok], W[ok], So[ok], Io[ok], Co[ok], Wo[ok], P[ok]
    # W
    dW = np.abs(W - Wo)
    assert (dW <= REL * np.maximum(1.0, np.abs(Wo))).all(), f"{what}: W off by {dW.max():.3e}"
    # S
    dS = np.abs(S - So)
    assert (dS <= REL * np.maximum(np.abs(So), 1e-6 * diag)).all(), f"{what}: S off by {dS.max():.3e}"
    far = np.abs(So) >= 1e-6 * diag
    assert np.array_equal(np.sign(S[far]), np.sign(So[far])), f"{what}: sign(S) differs"
    # I, C
    mism = I != Io
    if mism.any():
        k = np.nonzero(mism)[0]
        d_ours = oracle.face_distance(V, F, Pk[k], I[k])
        d_ref = oracle.face_distance(V, F, Pk[k], Io[k])
        bad = (d_ours != d_ref) & far[k]
        assert not bad.any(), f"{what}: {bad.sum()} face mismatches that are not exact ties"
    assert (np.linalg.norm(C - Co, axis=1) <= REL * diag).all(), f"{what}: C off"
    return dict(n=len(So), nan=int(nan_o.sum()), ties=int(mism.sum()), dW=float(dW.max(initial=0)),
                dS=float(dS.max(initial=0)))


def match_nodes(ca, cb):
    """Pair node ids of two 4-ary trees with identical lane structure; None if they differ."""
    pairs = [(0, 0)]
    out = []
    while pairs:
        a, b = pairs.pop()
        out.append((a, b))
        for x, y in zip(ca[a], cb[b]):
            x, y = int(x), int(y)
            xi = bool(x & 0x80000000) and x != 0xFFFFFFFF
            yi = bool(y & 0x80000000) and y != 0xFFFFFFFF
            if xi != yi:
                return None
            if xi:
                pairs.append((x & 0x7FFFFFFF, y & 0x7FFFFFFF))
            elif x != y:
                return None
    return np.array(out)


def check_tree_against_reference(mesh, V, F, ref=None, what=""):
    """Topology (cluster membership, lane order) and all 23 moment rows vs the genuine reference."""
    a = mesh.download()
    ref = ref or oracle.RefFwn(V, F)
    rn, rd = ref.nodes(), ref.boxdata()
    assert a["fwn_child"].shape == rn.shape, f"{what}: node count {a['fwn_child'].shape[0]} vs reference {rn.shape[0]}"
    pr = match_nodes(a["fwn_child"], rn)
    assert pr is not None, f"{what}: tree topology differs from the reference"
    assert len(pr) == rn.shape[0]
    da, db = a["fwn_data"][pr[:, 0]], rd[pr[:, 1]]
    # P and R^2 steer the descend test: must be bit-identical
    assert np.array_equal(da[:, :4].view(np.uint32), db[:, :4].view(np.uint32)), f"{what}: cluster centres/radii differ"
    fin = np.isfinite(db)
    scale = np.abs(db[fin]).max(initial=1.0)
    assert np.abs(da[fin] - db[fin]).max(initial=0.0) <= 1e-6 * scale, f"{what}: moments differ"
    assert np.array_equal(a["cp_child"], a["fwn_child"])
    return a
