"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement, in numpy float32 arithmetic, of the reference's host-side scene preparation:

  GpuMesh::try_from            src/mesh_material/mod.rs:379-467   triangle list -> primitives, BVH build + flatten
  GpuNode::pack                src/mesh_material/mod.rs:185-201
  build_alias_table            src/mesh_material/mod.rs:330-376
  transformed_primitive_areas  src/mesh_material/mod.rs:318-328
  prepare_mesh_assets          src/mesh_material/mesh.rs:106-166  concatenation + GpuMeshIndex offsets
  prepare_instances            src/mesh_material/instance.rs:245-444  instance AABB, TLAS, emissives, alias tables
  prepare_material_assets      src/mesh_material/material.rs:139-203

Third-party arithmetic NOT present under /root/reference, restated from the published sources of the pinned
dependencies (Cargo.toml:20 `bvh = "= 0.7.1"`, Cargo.toml:27-30 `bevy = "0.9.1"` -> glam 0.22):
  bvh 0.7.1  BVH::build (recursive SAH, 6 buckets on the longest centroid axis, one shape per leaf),
             BVH::flatten_custom (depth-first navigator / leaf records),
             AABB::{empty,grow,join,center,size,surface_area,largest_axis}, EPSILON = 0.00001
  glam       Mat4::inverse (cofactor expansion), Mat4::transform_point3 / transform_vector3a
Parity unpinned: the reference holds no golden vectors for any of this; anchored on its call sites
(mod.rs:458-459, instance.rs:368-369,425-426) only.
"""
import sys

import numpy as np

from bevy_hikari_b200 import layout as L

F = np.float32
U32_MAX = 0xFFFFFFFF
EPSILON = F(0.00001)
NUM_BUCKETS = 6
SWAP_CHILDREN = False      # True: flatten the right child before the left one (never what bvh 0.7.1 does; a test hook)
INF = F(np.inf)


# ----------------------------------------------------------------------------------------------- bvh 0.7.1
def _surface_area(mn, mx):
    s = mx - mn
    return F(2.0) * (s[0] * s[1] + s[0] * s[2] + s[1] * s[2])


def _largest_axis(mn, mx):
    s = mx - mn
    if s[0] > s[1] and s[0] > s[2]:
        return 0
    if s[1] > s[2]:
        return 1
    return 2


class _Bvh:
    """nodes: list of ('leaf', shape) | ['node', l_min, l_max, l_index, r_min, r_max, r_index]"""

    def __init__(self, aabb_min, aabb_max):
        self.mn = np.ascontiguousarray(aabb_min, F)
        self.mx = np.ascontiguousarray(aabb_max, F)
        # AABB::center() = min + size / 2
        self.center = self.mn + (self.mx - self.mn) / F(2.0)
        self.nodes = []
        self.shape_node_index = np.zeros(len(self.mn), np.uint32)
        sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
        self._build(np.arange(len(self.mn)))

    def _joint(self, idx):
        return self.mn[idx].min(axis=0), self.mx[idx].max(axis=0)

    def _build(self, indices):
        nodes = self.nodes
        if len(indices) == 1:
            node_index = len(nodes)
            nodes.append(("leaf", int(indices[0])))
            self.shape_node_index[indices[0]] = node_index
            return node_index
        aabb_mn, aabb_mx = self._joint(indices)
        c = self.center[indices]
        cb_mn, cb_mx = c.min(axis=0), c.max(axis=0)
        node_index = len(nodes)
        nodes.append(None)
        axis = _largest_axis(cb_mn, cb_mx)
        split_axis_size = cb_mx[axis] - cb_mn[axis]
        if split_axis_size < EPSILON:
            half = len(indices) // 2
            li, ri = indices[:half], indices[half:]
            l_mn, l_mx = self._joint(li)
            r_mn, r_mx = self._joint(ri)
        else:
            rel = (c[:, axis] - cb_mn[axis]) / split_axis_size
            bucket = (rel * (F(NUM_BUCKETS) - F(0.01))).astype(np.int64)  # `as usize` truncation
            b_size = np.zeros(NUM_BUCKETS, np.int64)
            b_mn = np.full((NUM_BUCKETS, 3), INF, F)
            b_mx = np.full((NUM_BUCKETS, 3), -INF, F)
            for b in range(NUM_BUCKETS):
                sel = indices[bucket == b]
                b_size[b] = len(sel)
                if len(sel):
                    b_mn[b], b_mx[b] = self._joint(sel)
            min_bucket, min_cost = 0, INF
            l_mn = l_mx = r_mn = r_mx = None
            parent_area = _surface_area(aabb_mn, aabb_mx)
            with np.errstate(invalid="ignore", over="ignore"):
                for i in range(NUM_BUCKETS - 1):
                    cl_mn, cl_mx = b_mn[:i + 1].min(axis=0), b_mx[:i + 1].max(axis=0)
                    cr_mn, cr_mx = b_mn[i + 1:].min(axis=0), b_mx[i + 1:].max(axis=0)
                    nl, nr = F(b_size[:i + 1].sum()), F(b_size[i + 1:].sum())
                    cost = (nl * _surface_area(cl_mn, cl_mx) + nr * _surface_area(cr_mn, cr_mx)) / parent_area
                    if cost < min_cost:
                        min_bucket, min_cost = i, cost
                        l_mn, l_mx, r_mn, r_mx = cl_mn, cl_mx, cr_mn, cr_mx
            if l_mn is None:  # every cost NaN: Rust keeps AABB::empty() children and min_bucket 0
                l_mn = r_mn = np.full(3, INF, F)
                l_mx = r_mx = np.full(3, -INF, F)
            # concatenate bucket assignment vectors in bucket order (stable within a bucket)
            order = np.argsort(bucket, kind="stable")
            sorted_idx, sorted_b = indices[order], bucket[order]
            li, ri = sorted_idx[sorted_b <= min_bucket], sorted_idx[sorted_b > min_bucket]
        l_index = self._build(li)
        r_index = self._build(ri)
        nodes[node_index] = ("node", l_mn, l_mx, l_index, r_mn, r_mx, r_index)
        return node_index

    # flatten_custom(&GpuNode::pack)
    def flatten(self):
        out = []

        def pack(mn, mx, entry, exit_, shape):
            if entry == U32_MAX:
                entry = shape | 0x80000000
            out_rec = np.zeros((), L.NODE)
            out_rec["min"], out_rec["max"] = mn, mx
            out_rec["entry_index"], out_rec["exit_index"] = entry, exit_
            return out_rec

        empty_mn, empty_mx = np.full(3, INF, F), np.full(3, -INF, F)

        def flat(ni, next_free):
            n = self.nodes[ni]
            if n[0] == "leaf":
                next_shape = next_free + 1
                out.append(pack(empty_mn, empty_mx, U32_MAX, next_shape, n[1]))
                return next_shape
            _, l_mn, l_mx, l_index, r_mn, r_mx, r_index = n
            if SWAP_CHILDREN:      # test hook (tests/test_bvh_topology_invariance.py): a different, equally valid visit order
                l_mn, l_mx, l_index, r_mn, r_mx, r_index = r_mn, r_mx, r_index, l_mn, l_mx, l_index
            after_l = branch(l_mn, l_mx, l_index, next_free)
            return branch(r_mn, r_mx, r_index, after_l)

        def branch(mn, mx, index, next_free):
            out.append(None)
            assert len(out) - 1 == next_free
            after = flat(index, next_free + 1)
            out[next_free] = pack(mn, mx, next_free + 1, after, U32_MAX)
            return after

        flat(0, 0)
        return np.array(out, L.NODE) if out else np.zeros(0, L.NODE)


def build_bvh(aabb_min, aabb_max):
    """-> (flattened hk_node records, per-shape BVH node index)"""
    if len(aabb_min) == 0:
        return np.zeros(0, L.NODE), np.zeros(0, np.uint32)
    b = _Bvh(aabb_min, aabb_max)
    return b.flatten(), b.shape_node_index


# ------------------------------------------------------------------------------------------------- glam
def mat4_cols(m16):
    return np.asarray(m16, F).reshape(4, 4)  # [col][row]


def mat4_inverse(m16):
    m = mat4_cols(m16)
    (m00, m01, m02, m03), (m10, m11, m12, m13), (m20, m21, m22, m23), (m30, m31, m32, m33) = m
    coef00 = m22 * m33 - m32 * m23
    coef02 = m12 * m33 - m32 * m13
    coef03 = m12 * m23 - m22 * m13
    coef04 = m21 * m33 - m31 * m23
    coef06 = m11 * m33 - m31 * m13
    coef07 = m11 * m23 - m21 * m13
    coef08 = m21 * m32 - m31 * m22
    coef10 = m11 * m32 - m31 * m12
    coef11 = m11 * m22 - m21 * m12
    coef12 = m20 * m33 - m30 * m23
    coef14 = m10 * m33 - m30 * m13
    coef15 = m10 * m23 - m20 * m13
    coef16 = m20 * m32 - m30 * m22
    coef18 = m10 * m32 - m30 * m12
    coef19 = m10 * m22 - m20 * m12
    coef20 = m20 * m31 - m30 * m21
    coef22 = m10 * m31 - m30 * m11
    coef23 = m10 * m21 - m20 * m11
    V = lambda *a: np.array(a, F)
    fac0, fac1, fac2 = V(coef00, coef00, coef02, coef03), V(coef04, coef04, coef06, coef07), V(coef08, coef08, coef10, coef11)
    fac3, fac4, fac5 = V(coef12, coef12, coef14, coef15), V(coef16, coef16, coef18, coef19), V(coef20, coef20, coef22, coef23)
    vec0, vec1, vec2, vec3 = V(m10, m00, m00, m00), V(m11, m01, m01, m01), V(m12, m02, m02, m02), V(m13, m03, m03, m03)
    inv0 = vec1 * fac0 - vec2 * fac1 + vec3 * fac2
    inv1 = vec0 * fac0 - vec2 * fac3 + vec3 * fac4
    inv2 = vec0 * fac1 - vec1 * fac3 + vec3 * fac5
    inv3 = vec0 * fac2 - vec1 * fac4 + vec2 * fac5
    sign_a, sign_b = V(1, -1, 1, -1), V(-1, 1, -1, 1)
    inverse = np.stack([inv0 * sign_a, inv1 * sign_b, inv2 * sign_a, inv3 * sign_b])
    col0 = V(inverse[0][0], inverse[1][0], inverse[2][0], inverse[3][0])
    dot0 = m[0] * col0
    dot1 = F(F(F(dot0[0] + dot0[1]) + dot0[2]) + dot0[3])
    rcp_det = F(1.0) / dot1
    return (inverse * rcp_det).astype(F)


def transform_point3(m16, p):
    m = mat4_cols(m16)
    r = m[0] * p[0]
    r = r + m[1] * p[1]
    r = r + m[2] * p[2]
    r = r + m[3]
    return r[:3].astype(F)


def transform_vector3(m16, p):
    m = mat4_cols(m16)
    r = m[0] * p[0]
    r = r + m[1] * p[1]
    r = r + m[2] * p[2]
    return r[:3].astype(F)


def _cross(a, b):
    return np.array([a[1] * b[2] - b[1] * a[2], a[2] * b[0] - b[2] * a[0], a[0] * b[1] - b[0] * a[1]], F)


def _length(a):
    return F(np.sqrt(F(F(F(a[0] * a[0]) + F(a[1] * a[1])) + F(a[2] * a[2]))))


# ---------------------------------------------------------------------------------------- mesh -> GpuMesh
def gpu_mesh(pos, nrm, uv, idx):
    """GpuMesh::try_from (mod.rs:379-467), TriangleList only."""
    pos, nrm, uv = np.asarray(pos, F), np.asarray(nrm, F), np.asarray(uv, F)
    idx = np.asarray(idx, np.uint32).reshape(-1, 3)
    if len(idx) == 0:
        raise ValueError("NoPrimitive")
    verts = np.zeros(len(pos), L.VERTEX)
    verts["position"], verts["normal"], verts["u"], verts["v"] = pos, nrm, uv[:, 0], uv[:, 1]
    prims = np.zeros(len(idx), L.PRIMITIVE)
    tri = pos[idx]  # (n,3,3)
    prims["vertices"]["position"] = tri
    prims["vertices"]["index"] = idx
    nodes, _ = build_bvh(tri.min(axis=1), tri.max(axis=1))
    return {"vertices": verts, "primitives": prims, "nodes": nodes, "tri_index": idx}


def transformed_primitive_areas(mesh, transform):
    out = np.zeros(len(mesh["primitives"]), F)
    pos = mesh["vertices"]["position"]
    for i, (a, b, c) in enumerate(mesh["tri_index"]):
        v0, v1, v2 = (transform_point3(transform, pos[k]) for k in (a, b, c))
        out[i] = F(0.5) * abs(_length(_cross(v1 - v0, v2 - v0)))
    return out


def build_alias_table(mesh, transform):
    areas = transformed_primitive_areas(mesh, transform)
    n = len(areas)
    table = np.zeros(n, L.ALIAS_ENTRY)
    if n == 0:
        return table
    surface_area = F(0)
    for a in areas:  # iter().sum() : sequential f32
        surface_area = F(surface_area + a)
    mean_area = surface_area / F(n)
    probs = [(i, F(a / mean_area)) for i, a in enumerate(areas)]
    over = [p for p in probs if p[1] > 1.0]
    under = [p for p in probs if p[1] < 1.0]
    table["index"] = np.arange(n, dtype=np.uint32)
    while under and over:
        oi, op = over.pop()
        ui, up = under.pop()
        delta = F(F(1.0) - up)
        op = F(op - delta)
        assert op >= 0.0
        if op > 1.0:
            over.append((oi, op))
        elif op < 1.0:
            under.append((oi, op))
        table[ui] = (delta, oi)
    return table


# ------------------------------------------------------------------------------------------- whole scene
def build_scene(meshes, inst_mesh, inst_material, inst_transform, materials):
    """meshes: list of (pos,nrm,uv,idx); materials: structured array L.MATERIAL; -> dict of the 9 buffers."""
    gm = [gpu_mesh(*m) for m in meshes]
    # prepare_mesh_assets: concatenate in asset order
    mesh_index = []
    v_off = p_off = n_off = 0
    for g in gm:
        mesh_index.append((v_off, p_off, n_off, len(g["nodes"])))
        v_off += len(g["vertices"]); p_off += len(g["primitives"]); n_off += len(g["nodes"])
    vertices = np.concatenate([g["vertices"] for g in gm])
    primitives = np.concatenate([g["primitives"] for g in gm])
    asset_nodes = np.concatenate([g["nodes"] for g in gm])

    n_inst = len(inst_mesh)
    instances = np.zeros(n_inst, L.INSTANCE)
    for i in range(n_inst):
        g = gm[inst_mesh[i]]
        pos = g["vertices"]["position"]
        mn, mx = pos.min(axis=0), pos.max(axis=0)           # bevy Aabb::from_min_max of the mesh
        center, half = F(0.5) * (mx + mn), F(0.5) * (mx - mn)
        xf = np.asarray(inst_transform[i], F)
        c = transform_point3(xf, center)
        lo, hi = np.zeros(3, F), np.zeros(3, F)             # instance.rs:298-305 (starts from ZERO)
        for k in range(8):
            sgn = np.array([2 * (k & 1) - 1, 2 * ((k >> 1) & 1) - 1, 2 * ((k >> 2) & 1) - 1], F)
            v = transform_vector3(xf, half * sgn)
            lo, hi = np.minimum(lo, v), np.maximum(hi, v)
        instances[i]["min"], instances[i]["max"] = lo + c, hi + c
        instances[i]["model"] = xf
        inv = mat4_inverse(xf)
        instances[i]["inverse_transpose_model"] = inv.T.reshape(16)  # .inverse().transpose()
        instances[i]["mesh"] = mesh_index[inst_mesh[i]]
        instances[i]["material"] = inst_material[i]
    instance_nodes, node_idx = build_bvh(instances["min"], instances["max"])
    instances["node_index"] = node_idx

    emissives, alias = [], []
    for i in range(n_inst):
        mat = materials[inst_material[i]]
        e = mat["emissive"]
        intensity = F(F(255.0) * e[3]) * _length(e[:3])
        if intensity > 0.0:
            g = gm[inst_mesh[i]]
            xf = instances[i]["model"]
            table = build_alias_table(g, xf)
            off = sum(len(t) for t in alias)
            alias.append(table)
            areas = transformed_primitive_areas(g, xf)
            sa = F(0)
            for a in areas:
                sa = F(sa + a)
            em = np.zeros((), L.EMISSIVE)
            em["emissive"] = e
            em["position"] = F(0.5) * (instances[i]["max"] + instances[i]["min"])
            em["radius"] = F(F(0.5) * _length(instances[i]["max"] - instances[i]["min"])) + F(np.sqrt(intensity))
            em["instance"] = i
            em["alias_table_offset"], em["alias_table_count"] = off, len(table)
            em["surface_area"] = sa
            emissives.append(em)
    emissives = np.array(emissives, L.EMISSIVE) if emissives else np.zeros(0, L.EMISSIVE)
    if len(emissives):
        r = emissives["radius"][:, None]
        emissive_nodes, enode = build_bvh(emissives["position"] - r, emissives["position"] + r)
        emissives["node_index"] = enode
    else:
        emissive_nodes = np.zeros(0, L.NODE)
    alias_table = np.concatenate(alias) if alias else np.zeros(0, L.ALIAS_ENTRY)
    return {"vertices": vertices, "primitives": primitives, "asset_nodes": asset_nodes, "alias_table": alias_table,
            "instances": instances, "instance_nodes": instance_nodes, "materials": np.asarray(materials, L.MATERIAL),
            "emissive_nodes": emissive_nodes, "emissives": emissives}
