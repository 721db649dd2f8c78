"""Seeded synthetic scenes (SURVEY.md section 8(d) [DECISION]); host-side NumPy so the CPU
oracle and the GPU path see bit-identical inputs.  No dataset is available offline.

  (U) uniform volume:  x,y ~ U[-4,4], z ~ U[0,3]          -- sparse balls, little early exit
  (S) surface:         floor + 4 walls of a 6x6x3 m room + 12 random axis-aligned boxes
                       (edge U[0.3,2.0]), uniform on the surfaces, + N(0, 5 mm) noise
Input feature = height above the lowest point, shape (B,1,N).  Scene `i` uses
default_rng(1234 + i)."""
from __future__ import annotations

import numpy as np


def _surface_scene(rng, N):
    # surfaces: (origin, edge_u, edge_v) rectangles
    rects = []
    L, H = 6.0, 3.0
    o = -L / 2
    rects.append((np.array([o, o, 0.0]), np.array([L, 0, 0]), np.array([0, L, 0])))          # floor
    rects.append((np.array([o, o, 0.0]), np.array([L, 0, 0]), np.array([0, 0, H])))          # walls
    rects.append((np.array([o, -o, 0.0]), np.array([L, 0, 0]), np.array([0, 0, H])))
    rects.append((np.array([o, o, 0.0]), np.array([0, L, 0]), np.array([0, 0, H])))
    rects.append((np.array([-o, o, 0.0]), np.array([0, L, 0]), np.array([0, 0, H])))
    for _ in range(12):
        e = rng.uniform(0.3, 2.0, 3)
        c = np.array([rng.uniform(o, -o - e[0]), rng.uniform(o, -o - e[1]), 0.0])
        ex, ey, ez = np.array([e[0], 0, 0]), np.array([0, e[1], 0]), np.array([0, 0, min(e[2], H)])
        rects += [(c + ez, ex, ey),                                  # top
                  (c, ex, ez), (c + ey, ex, ez), (c, ey, ez), (c + ex, ey, ez)]
    area = np.array([np.linalg.norm(np.cross(u, v)) for (_, u, v) in rects])
    which = rng.choice(len(rects), size=N, p=area / area.sum())
    a = rng.random(N)
    b = rng.random(N)
    org = np.stack([r[0] for r in rects])[which]
    eu = np.stack([r[1] for r in rects])[which]
    ev = np.stack([r[2] for r in rects])[which]
    p = org + a[:, None] * eu + b[:, None] * ev + rng.normal(0.0, 0.005, (N, 3))
    return p.astype(np.float32)


def make_scenes(B: int, N: int, kind: str = "surface", first_scene: int = 0):
    """-> xyz (B,N,3) f32, features (B,1,N) f32 (height)."""
    xyz = np.empty((B, N, 3), dtype=np.float32)
    for i in range(B):
        rng = np.random.default_rng(1234 + first_scene + i)
        if kind == "uniform":
            p = rng.random((N, 3)) * np.array([8.0, 8.0, 3.0]) - np.array([4.0, 4.0, 0.0])
            xyz[i] = p.astype(np.float32)
        elif kind == "surface":
            xyz[i] = _surface_scene(rng, N)
        else:
            raise ValueError(f"unknown scene kind {kind!r}")
    feat = (xyz[:, :, 2] - xyz[:, :, 2].min(axis=1, keepdims=True))[:, None, :].astype(np.float32)
    return np.ascontiguousarray(xyz), np.ascontiguousarray(feat)


def make_sizes(B: int, K: int, first_scene: int = 0):
    """Predicted box sizes per cluster, size ~ U[0.2, 2.0]^3 (BASELINE config 3) -> (B,K,3) f32."""
    out = np.empty((B, K, 3), dtype=np.float32)
    for i in range(B):
        rng = np.random.default_rng(99991 + first_scene + i)
        out[i] = rng.uniform(0.2, 2.0, (K, 3)).astype(np.float32)
    return out
