#!/usr/bin/env python
"""Experiment (GPU): throughput of the stand-alone traversal kernel (k_trace_rays, 10 KB of code) on the ray population
k_indirect traces, in the same warp layout (8x4 pixel tiles) — to separate the cost of traversal itself from the cost of
running it inside the 126 KB light kernel (instruction starvation, ncu: 7 warps stalled 'no instruction' per issue)."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bevy_hikari_b200 import layout as L
from tests.conftest import Bench

cfg = sys.argv[1] if len(sys.argv) > 1 else "cornell_1080p"
from bevy_hikari_b200 import scenes
c = scenes.CONFIGS[cfg]
b = Bench(c["scene"], c["width"], c["height"], config=cfg)
dev = b.device()
dev.set_profiling(False, True)
for f in range(1, 4):
    dev.render_frame(b.inputs(f))
st = dev.stats()
print("frame kernel ms:", {n: round(st.ms_kernel[i], 3) for i, n in enumerate(L.KERNEL_NAMES) if st.ms_kernel[i] > 0})
pos = dev.readback(L.OUT_GBUFFER_POSITION)
nrm = dev.readback(L.OUT_GBUFFER_NORMAL).astype(np.float32) / 127.0
H, W = pos.shape[:2]
# 8x4 tile order (one warp = one tile), like tile_pixel()
ys, xs = np.mgrid[0:H, 0:W]
tile_key = (ys // 8) * 10_000_000 + (xs // 16) * 1000 + ((ys % 8) // 4 * 2 + (xs % 16) // 8) * 100 + (ys % 4) * 8 + (xs % 8)
order = np.argsort(tile_key.reshape(-1), kind="stable")
P = pos.reshape(-1, 4)[order]; N = nrm.reshape(-1, 4)[order][:, :3]
hit = P[:, 3] > 0
rng = np.random.default_rng(1)
n = len(P)
N = N / np.maximum(np.linalg.norm(N, axis=1, keepdims=True), 1e-6)
r1, r2 = rng.random(n), rng.random(n)
phi = 2 * np.pi * r1; ct = np.sqrt(1 - r2); stt = np.sqrt(r2)
local = np.stack([np.cos(phi) * stt, np.sin(phi) * stt, ct], axis=1)
up = np.where(np.abs(N[:, 1:2]) < 0.99, np.array([[0, 1.0, 0]]), np.array([[1.0, 0, 0]]))
T = np.cross(up, N); T /= np.linalg.norm(T, axis=1, keepdims=True); B = np.cross(N, T)
D = (T * local[:, :1] + B * local[:, 1:2] + N * local[:, 2:3]).astype(np.float32)
rays = np.zeros(n, L.RAY)
rays["origin"] = P[:, :3] + N.astype(np.float32) * np.float32(0.01)
rays["direction"] = D
rays["max_distance"] = np.float32(3.0e38)
rays["early_distance"] = 0.0
rays["exclude_instance"] = 0xFFFFFFFF
# background pixels trace nothing in k_indirect: give them a ray that misses the TLAS root immediately
rays["origin"][~hit] = (1e6, 1e6, 1e6); rays["direction"][~hit] = (0, 1, 0)
for rep in range(3):
    hits = dev.trace_rays(rays)
    ms = dev.stats().ms_kernel[15]
print(f"closest-hit bounce rays: {hit.sum()} rays (+{(~hit).sum()} idle lanes), k_trace_rays {ms:.3f} ms -> {hit.sum() / ms / 1e3:.0f} Mrays/s")
# shadow-like rays from the hit points towards the light (any-hit: early_distance = max_distance)
ok = hit & (hits["instance_index"] != 0xFFFFFFFF)
hp = rays["origin"] + rays["direction"] * hits["distance"][:, None]
light = np.array([0.0, 1.98, 0.03], np.float32) + (rng.random((n, 3)).astype(np.float32) - 0.5) * np.array([0.46, 0.0, 0.38], np.float32)
d2 = light - hp; dist = np.linalg.norm(d2, axis=1); d2 = (d2 / np.maximum(dist[:, None], 1e-6)).astype(np.float32)
rays2 = rays.copy()
rays2["origin"][ok] = (hp - rays["direction"] * np.float32(0.01))[ok]
rays2["direction"][ok] = d2[ok]
rays2["max_distance"][ok] = dist[ok].astype(np.float32)
rays2["early_distance"][ok] = dist[ok].astype(np.float32)
rays2["origin"][~ok] = (1e6, 1e6, 1e6); rays2["direction"][~ok] = (0, 1, 0)
for rep in range(3):
    dev.trace_rays(rays2)
    ms2 = dev.stats().ms_kernel[15]
print(f"any-hit shadow rays: {ok.sum()} rays, k_trace_rays {ms2:.3f} ms -> {ok.sum() / ms2 / 1e3:.0f} Mrays/s")
# the same rays in random order (what a global ray queue without sorting would see)
perm = rng.permutation(n)
for rep in range(2):
    dev.trace_rays(rays[perm]); ms3 = dev.stats().ms_kernel[15]
print(f"closest-hit, shuffled across the frame: {ms3:.3f} ms")
# compacted (no idle lanes): only the real rays
rc = rays[hit]
for rep in range(2):
    dev.trace_rays(rc); ms4 = dev.stats().ms_kernel[15]
print(f"closest-hit, compacted to {len(rc)} rays: {ms4:.3f} ms")
