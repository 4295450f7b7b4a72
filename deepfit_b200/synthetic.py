"""Seeded synthetic analytic-surface clouds of the sizes BASELINE.json names.
Pure numpy on the host: these are inputs, not part of the hot path."""
from __future__ import annotations

import numpy as np


def _finish(p, normals, noise, rng):
    p = np.asarray(p, dtype=np.float64)
    if noise > 0:
        bbdiag = np.linalg.norm(p.max(0) - p.min(0))
        p = p + normals * rng.normal(scale=noise * bbdiag, size=(p.shape[0], 1))
    return p.astype(np.float32)


def sphere_cloud(n: int, seed: int = 1234, radius: float = 1.0, noise: float = 0.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return _finish(v * radius, v, noise, rng)


def torus_cloud(n: int, seed: int = 1234, R: float = 1.0, r: float = 0.3, noise: float = 0.0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    u = rng.random(n) * 2 * np.pi
    v = rng.random(n) * 2 * np.pi
    cu, su, cv, sv = np.cos(u), np.sin(u), np.cos(v), np.sin(v)
    p = np.stack([(R + r * cv) * cu, (R + r * cv) * su, r * sv], 1)
    nrm = np.stack([cv * cu, cv * su, sv], 1)
    return _finish(p, nrm, noise, rng)


def multi_surface_cloud(n: int, seed: int = 4321, noise: float = 0.0) -> np.ndarray:
    """Union of sphere, torus, saddle sheet and cylinder in a common box (config 4)."""
    rng = np.random.default_rng(seed)
    m = n // 4
    parts = [sphere_cloud(m, seed + 1, 0.6, noise) + np.float32([1.5, 0, 0]),
             torus_cloud(m, seed + 2, 0.8, 0.25, noise) + np.float32([-1.5, 0, 0])]
    xy = rng.random((m, 2)) * 2 - 1
    z = 0.4 * (xy[:, 0] ** 2 - xy[:, 1] ** 2)
    parts.append(np.stack([xy[:, 0], xy[:, 1] + 2.0, z], 1).astype(np.float32))
    m2 = n - 3 * m
    t = rng.random(m2) * 2 * np.pi
    h = rng.random(m2) * 2 - 1
    parts.append(np.stack([0.5 * np.cos(t), 0.5 * np.sin(t) - 2.0, h], 1).astype(np.float32))
    p = np.concatenate(parts)
    return p[rng.permutation(len(p))]


def variable_density_cloud(n: int, seed: int = 9876) -> np.ndarray:
    """Lidar-like 1/r^2 density falloff on a wavy ground sheet plus a few objects (config 5)."""
    rng = np.random.default_rng(seed)
    r = 1.0 / (rng.random(n) * 0.98 + 0.02)          # heavy near the sensor, out to 50 units
    th = rng.random(n) * 2 * np.pi
    x, y = r * np.cos(th), r * np.sin(th)
    z = 0.3 * np.sin(0.4 * x) * np.cos(0.3 * y) + 0.02 * rng.normal(size=n)
    return np.stack([x, y, z], 1).astype(np.float32)
