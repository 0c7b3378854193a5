#!/usr/bin/env python
"""Offline study (no GPU): what would a 4-wide BVH with an ordered stack walk buy over the reference's skip-link walk on the bounce
rays of the benchmark path?  (VERDICT r1 "next round" item 8; companion of tools/exp_ordered_traversal.py, which priced ordering alone.)

The binary trees are the ones the host builds (bvh 0.7.1 restated); each is collapsed to 4 children per node by repeatedly opening the
child with the largest surface area.  Rays are the cosine-distributed bounce rays of whole 8x4 pixel tiles (= the warps of the light
kernels), so that per-warp maxima can be priced next to per-ray means:

  fixed   skip-link walk in the reference's order (this script's own walker; records / instance entries / triangle tests)
  wide    4-wide nodes, children sorted by entry distance, farther ones pushed with their distance and dropped when stale

Cost model (SASS instruction counts of the shipped walk, tools/sass_lines.py): skip-link record 37, instance entry 150 (matrix, three
IEEE reciprocals), triangle 60; 4-wide node 8 loads + 4 slab tests + sort + pushes ~ 125.

usage: tools/exp_wide_traversal.py [scene=city] [width=480] [height=270] [tiles=40]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bevy_hikari_b200 import layout as L   # noqa: E402
from tests.conftest import Bench            # noqa: E402

LEAF = 0x80000000
C_REC, C_INST, C_TRI, C_NODE = 37, 150, 60, 125


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


class Tree:
    """binary tree parsed back from a flat skip-link array: node = ('leaf', shape, box) | ('inner', [children], box)"""
    def __init__(self, nodes):
        self.mn = nodes["min"].astype(np.float64); self.mx = nodes["max"].astype(np.float64)
        self.entry = nodes["entry_index"].astype(np.int64); self.exit = nodes["exit_index"].astype(np.int64)
        self.n = len(nodes)
        if self.n == 1:       # a single-shape BVH is one leaf record without a navigator
            self.root_children = [("leaf", int(self.entry[0]) - LEAF, (np.full(3, -np.inf), np.full(3, np.inf)))]
        else:
            self.root_children = self.parse_range(0, self.n)

    def parse_range(self, i, end):
        kids = []
        while i < end:
            box = (self.mn[i], self.mx[i])
            e = int(self.exit[i])
            if self.entry[i + 1] >= LEAF if i + 1 < self.n else False:
                kids.append(("leaf", int(self.entry[i + 1]) - LEAF, box))
            else:
                kids.append(("inner", self.parse_range(i + 1, e), box))
            i = e
        return kids


def area(box):
    s = np.maximum(box[1] - box[0], 0)
    return s[0] * s[1] + s[0] * s[2] + s[1] * s[2]


def collapse(children, width=4):
    """list of binary children -> wide node: ('node', [(box, child)]) with child = ('leaf', shape) | wide node"""
    kids = list(children)
    while len(kids) < width:
        cand = [(area(k[2]), j) for j, k in enumerate(kids) if k[0] == "inner"]
        if not cand:
            break
        _, j = max(cand)
        k = kids.pop(j)
        kids[j:j] = k[1]
    out = []
    for k in kids:
        out.append((k[2], ("leaf", k[1]) if k[0] == "leaf" else collapse(k[1], width)))
    return ("node", out)


def fixed_walk(t, o, d, best, leaf_fn, cnt):
    inv = 1.0 / np.where(d == 0, 1e-30, d)
    i = 0
    while i < t.n:
        cnt[0] += 1
        if t.entry[i] >= LEAF:
            best = min(best, leaf_fn(int(t.entry[i]) - LEAF, best))
            i = int(t.exit[i])
        else:
            i = int(t.entry[i]) if slab(o, inv, t.mn[i], t.mx[i]) < best else int(t.exit[i])
    return best


def wide_walk(root, o, d, best, leaf_fn, cnt):
    inv = 1.0 / np.where(d == 0, 1e-30, d)
    stack = [(0.0, root)]
    while stack:
        t_in, node = stack.pop()
        if t_in >= best:
            continue
        if node[0] == "leaf":
            best = min(best, leaf_fn(node[1], best))
            continue
        cnt[0] += 1
        hits = []
        for box, child in node[1]:
            tk = slab(o, inv, box[0], box[1])
            if tk < best:
                hits.append((tk, id(child), child))
        for tk, _, child in sorted(hits, reverse=True):
            stack.append((tk, child))
    return best


def main():
    args = dict(a.split("=") for a in sys.argv[1:])
    scene, W, H, n_tiles = args.get("scene", "city"), int(args.get("width", 480)), int(args.get("height", 270)), int(args.get("tiles", 40))
    width = int(args.get("wide", 4))
    b = Bench(scene, W, H, config={"city": "city_4k", "town": "scene_1080p"}.get(scene, "cornell_1080p"))
    orc = b.oracle()
    orc.prepass(b.inputs(1))
    pos = orc.readback(L.OUT_GBUFFER_POSITION).reshape(H, W, 4)
    nrm = np.maximum(orc.readback(L.OUT_GBUFFER_NORMAL).astype(np.float32) / 127.0, -1.0).reshape(H, W, 4)[..., :3]
    rng = np.random.default_rng(1)
    tiles = []
    while len(tiles) < n_tiles:
        tx, ty = int(rng.integers(0, W // 8)) * 8, int(rng.integers(0, H // 4)) * 4
        if (pos[ty:ty + 4, tx:tx + 8, 3] > 0).all():
            tiles.append((tx, ty))
    bufs = b.world.buffers()
    inst, prims = bufs["instances"], bufs["primitives"]
    tlas = Tree(bufs["instance_nodes"])
    tlas_wide = collapse(tlas.root_children, width)
    blas, blas_wide = {}, {}

    def mesh_tree(one):
        mesh = one["mesh"]
        key = (int(mesh["node_offset"]), int(mesh["node_count"]))
        if key not in blas:
            blas[key] = Tree(bufs["asset_nodes"][key[0]:key[0] + key[1]])
            blas_wide[key] = collapse(blas[key].root_children, width) if key[1] > 1 else ("leaf", 0)
        return key

    per_ray = []          # (fixed cost, wide cost, fixed steps, wide steps)
    agree = 0
    for tx, ty in tiles:
        for py in range(ty, ty + 4):
            for px in range(tx, tx + 8):
                N = nrm[py, px].astype(np.float64); N /= np.linalg.norm(N)
                r1, r2 = rng.random(), rng.random()
                a = 2 * np.pi * r2; rad = np.sqrt(r1)
                t = np.cross(N, [1.0, 0, 0] if abs(N[0]) < 0.9 else [0, 1.0, 0]); t /= np.linalg.norm(t)
                bt = np.cross(N, t)
                d = t * rad * np.cos(a) + bt * rad * np.sin(a) + N * np.sqrt(1 - r1)
                o = pos[py, px, :3].astype(np.float64) + N * 0.02
                res = []
                for mode in ("fixed", "wide"):
                    c_t, c_b, c_i, c_tri = [0], [0], [0], [0]

                    def instance_leaf(i, best):
                        c_i[0] += 1
                        one = inst[i]
                        m = one["inverse_transpose_model"].reshape(4, 4).astype(np.float64)
                        ol = np.append(o, 1.0) @ m.T; ol = ol[:3] / ol[3]
                        dl = (np.append(d, 0.0) @ m.T)[:3]
                        key = mesh_tree(one)
                        p0 = int(one["mesh"]["primitive"])

                        def triangle_leaf(j, best_):
                            c_tri[0] += 1
                            return tri(ol, dl, prims[p0 + j]["vertices"]["position"].astype(np.float64))
                        if mode == "fixed":
                            return fixed_walk(blas[key], ol, dl, best, triangle_leaf, c_b)
                        return wide_walk(blas_wide[key], ol, dl, best, triangle_leaf, c_b)
                    if mode == "fixed":
                        best = fixed_walk(tlas, o, d, np.inf, instance_leaf, c_t)
                        cost = (c_t[0] + c_b[0] - c_i[0] - c_tri[0]) * C_REC + c_i[0] * C_INST + c_tri[0] * C_TRI
                        steps = c_t[0] + c_b[0]
                    else:
                        best = wide_walk(tlas_wide, o, d, np.inf, instance_leaf, c_t)
                        cost = (c_t[0] + c_b[0]) * C_NODE + c_i[0] * C_INST + c_tri[0] * C_TRI
                        steps = c_t[0] + c_b[0] + c_i[0] + c_tri[0]
                    res.append((best, cost, steps, c_i[0], c_tri[0]))
                agree += (np.isinf(res[0][0]) and np.isinf(res[1][0])) or abs(res[0][0] - res[1][0]) <= 1e-9 * max(1.0, res[0][0])
                per_ray.append((res[0][1], res[1][1], res[0][2], res[1][2], res[0][3], res[1][3], res[0][4], res[1][4]))
    a = np.array(per_ray, np.float64).reshape(len(tiles), 32, 8)
    n = a.shape[0] * 32
    print(f"{scene} {W}x{H}, {len(tiles)} warps of 8x4 pixels ({n} bounce rays), {width}-wide; closest hit agrees on {agree} rays")
    for name, c, s, ci, ct in (("skip-link, fixed order", 0, 2, 4, 6), (f"{width}-wide, ordered", 1, 3, 5, 7)):
        cost, steps = a[..., c], a[..., s]
        print(f"  {name:24s} dependent steps/ray {steps.mean():7.1f} (warp max {steps.max(axis=1).mean():7.1f})  instance entries {a[..., ci].mean():5.2f}  "
              f"triangles {a[..., ct].mean():5.2f}  instructions/ray {cost.mean():8.0f}  per-warp max {cost.max(axis=1).mean():8.0f}  "
              f"SIMT efficiency (mean/max) {cost.mean() / cost.max(axis=1).mean():.2f}")
    print(f"  ratio fixed / wide: instructions {a[..., 0].mean() / a[..., 1].mean():.2f}   per-warp max {a[..., 0].max(axis=1).mean() / a[..., 1].max(axis=1).mean():.2f}   "
          f"dependent steps (warp max) {a[..., 2].max(axis=1).mean() / a[..., 3].max(axis=1).mean():.2f}")


if __name__ == "__main__":
    main()
