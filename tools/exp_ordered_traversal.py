#!/usr/bin/env python
"""Offline study (no GPU): how many BVH records would a front-to-back closest-hit walk visit, against the reference's fixed-order
skip-link walk, on the SAME flattened trees and the same rays?  (tests/test_bvh_topology_invariance.py shows
that reordering the walk changes no image in single-emissive scenes, so an opt-in image-exact mode may reorder.)

  fixed   = the oracle's step counter (hko_trace_steps): TLAS records + BLAS records + triangle tests of traverse_top / traverse_bottom
  ordered = this script: at every inner node both child boxes are tested, the nearer child is entered first, the farther one is
            pushed with its entry distance and dropped when popped if that distance is not below the best hit so far; the same
            two-level structure (TLAS leaf -> transform the ray -> BLAS), the same slab and Moeller-Trumbore tests (float64 here)

usage: tools/exp_ordered_traversal.py [scene=city] [width=64] [height=36] [rays=1500]
Rays: the bounce rays of the benchmark path — cosine-distributed directions from the G-buffer points of frame 1."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bevy_hikari_b200 import layout as L   # noqa: E402
from tests.conftest import Bench            # noqa: E402

LEAF = 0x80000000


def slab(o, inv, mn, mx):
    t0, t1 = (mn - o) * inv, (mx - o) * inv
    tmin, tmax = np.minimum(t0, t1).max(), np.maximum(t0, t1).min()
    return tmin if (tmax >= tmin and tmax >= 0) else np.inf


def tri(o, d, p):
    ab, ac = p[1] - p[0], p[2] - p[0]
    u_vec = np.cross(d, ac); det = ab @ u_vec
    if abs(det) < 1.1920929e-7:
        return np.inf
    inv = 1.0 / det; ao = o - p[0]
    u = (ao @ u_vec) * inv
    if u < 0 or u > 1:
        return np.inf
    v_vec = np.cross(ao, ab); v = (d @ v_vec) * inv
    if v < 0 or u + v > 1:
        return np.inf
    t = (ac @ v_vec) * inv
    return t if t > 1.1920929e-7 else np.inf


def ordered_walk(nodes, lo, hi, o, d, best, leaf_fn, counter):
    """closest hit below `best` in the record range [lo, hi): front to back with a stack of (entry distance, first, end)"""
    inv = 1.0 / np.where(d == 0, 1e-30, d)
    stack = [(0.0, lo, hi)]
    while stack:
        t_in, first, end = stack.pop()
        if t_in >= best:
            continue
        n = nodes[first]
        if int(n["entry_index"]) >= LEAF:                       # the subtree is one leaf record
            counter[0] += 1
            best = min(best, leaf_fn(int(n["entry_index"]) - LEAF, best))
            continue
        # the range holds navigator records of the children: [first] and [its exit] ... (binary here: exactly two)
        kids = []
        i = first
        while i < end:
            r = nodes[i]
            counter[0] += 1
            kids.append((slab(o, inv, r["min"].astype(np.float64), r["max"].astype(np.float64)), int(r["entry_index"]), int(r["exit_index"])))
            i = int(r["exit_index"])
        for t_k, a, b_ in sorted(kids, reverse=True):           # farther first onto the stack, nearer popped first
            if t_k < best:
                stack.append((t_k, a, b_))
    return best


def main():
    args = dict(a.split("=") for a in sys.argv[1:])
    scene, W, H, n_rays = args.get("scene", "city"), int(args.get("width", 64)), int(args.get("height", 36)), int(args.get("rays", 1500))
    b = Bench(scene, W, H, config="city_4k" if scene == "city" else "cornell_1080p")
    orc = b.oracle()
    orc.prepass(b.inputs(1))
    pos = orc.readback(L.OUT_GBUFFER_POSITION).reshape(-1, 4)
    nrm = np.maximum(orc.readback(L.OUT_GBUFFER_NORMAL).astype(np.float32) / 127.0, -1.0).reshape(-1, 4)[:, :3]
    covered = np.nonzero(pos[:, 3] > 0)[0]
    rng = np.random.default_rng(1)
    pick = rng.choice(covered, min(n_rays, len(covered)), replace=False)
    N = nrm[pick] / np.linalg.norm(nrm[pick], axis=1, keepdims=True)
    r1, r2 = rng.random(len(pick)), rng.random(len(pick))
    a = 2 * np.pi * r2; rad = np.sqrt(r1)
    local = np.stack([rad * np.cos(a), rad * np.sin(a), np.sqrt(1 - r1)], 1)
    t = np.cross(N, np.where(np.abs(N[:, :1]) < 0.9, [[1.0, 0, 0]], [[0, 1.0, 0]])); t /= np.linalg.norm(t, axis=1, keepdims=True)
    bt = np.cross(N, t)
    D = (t * local[:, :1] + bt * local[:, 1:2] + N * local[:, 2:3]).astype(np.float32)
    O = (pos[pick, :3] + N * 0.02).astype(np.float32)
    rays = np.zeros(len(pick), L.RAY)
    rays["origin"], rays["direction"] = O, D
    rays["max_distance"], rays["early_distance"], rays["exclude_instance"] = np.float32(3.4e38), 0.0, 0xFFFFFFFF
    fixed = orc.trace_steps(rays).astype(np.int64)
    hits = orc.trace_rays(rays)

    bufs = b.world.buffers()
    inst, inodes, anodes, prims = bufs["instances"], bufs["instance_nodes"], bufs["asset_nodes"], bufs["primitives"]
    ordered = np.zeros((len(pick), 2), np.int64)
    agree = 0
    for k in range(len(pick)):
        o, d = O[k].astype(np.float64), D[k].astype(np.float64)
        cnt_t, cnt_b, cnt_tri = [0], [0], [0]

        def instance_leaf(i, best):
            one = inst[i]
            inv_model = one["inverse_transpose_model"].reshape(4, 4).astype(np.float64)     # [col][row] of the inverse TRANSPOSE
            m = inv_model                                                                    # transpose(inverse_transpose) applied as row-vector product
            ol = np.append(o, 1.0) @ m.T; ol = ol[:3] / ol[3]
            dl = np.append(d, 0.0) @ m.T; dl = dl[:3]
            mesh = one["mesh"]
            base, count, p0 = int(mesh["node_offset"]), int(mesh["node_count"]), int(mesh["primitive"])

            def triangle_leaf(j, best_):
                cnt_tri[0] += 1
                return tri(ol, dl, prims[p0 + j]["vertices"]["position"].astype(np.float64))
            sub = anodes[base:base + count].copy()
            return ordered_walk(sub, 0, count, ol, dl, best, triangle_leaf, cnt_b)
        best = ordered_walk(inodes, 0, len(inodes), o, d, np.inf, instance_leaf, cnt_t)
        ordered[k] = (cnt_t[0] + cnt_b[0], cnt_tri[0])
        ref_t = hits["distance"][k] if hits["instance_index"][k] != 0xFFFFFFFF else np.inf
        agree += (np.isinf(best) and np.isinf(ref_t)) or (np.isfinite(best) and abs(best - ref_t) <= 1e-3 * max(1.0, ref_t))
    f_rec, f_tri = fixed[:, 0] + fixed[:, 1], fixed[:, 2]
    print(f"{scene} {W}x{H}: {len(pick)} bounce rays, closest hit agrees for {agree} of them")
    print(f"  records visited per ray   fixed order {f_rec.mean():8.1f}   front-to-back {ordered[:, 0].mean():8.1f}   ratio {f_rec.mean() / ordered[:, 0].mean():.2f}")
    print(f"  triangle tests per ray    fixed order {f_tri.mean():8.1f}   front-to-back {ordered[:, 1].mean():8.1f}   ratio {f_tri.mean() / max(ordered[:, 1].mean(), 1e-9):.2f}")
    print(f"  max records (fixed / ordered) {f_rec.max()} / {ordered[:, 0].max()};  SIMT-relevant spread: std {f_rec.std():.1f} / {ordered[:, 0].std():.1f}")


if __name__ == "__main__":
    main()
