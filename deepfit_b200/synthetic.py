"""Seeded synthetic analytic-surface clouds of the sizes BASELINE.json names.
Pure numpy on the host: these are inputs, not part of the hot path.  ``synthetic_state_dict`` is the seeded
stand-in for the missing ``trained_models/DeepFit/DeepFit.pth`` used by bench.py and tools/ (the tests use the
oracle's own generator and the golden fixtures made with it).  Nothing here imports ``oracle/``."""
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


def write_pcpnet_like(root: str, n_shapes: int = 20, n_points: int = 100000, n_sparse: int = 5000,
                      noise_levels=(0.0, 0.0025, 0.012), list_name: str = "testset_synthetic.txt", seed0: int = 2000):
    """BASELINE config 3: a PCPNet-shaped dataset -- per shape ``.xyz``, analytic ``.normals``, ``.pidx``
    (sparse patch centres) and a list file; Gaussian noise of the named fractions of the bounding-box
    diagonal along the normal; raw (un-normalised) coordinates like PCPNet's own files."""
    import os
    os.makedirs(root, exist_ok=True)
    names = []
    kinds = ["sphere", "torus", "ellipsoid", "sheet", "cylinder"]
    for i in range(n_shapes):
        rng = np.random.default_rng(seed0 + i)
        kind = kinds[i % len(kinds)]
        noise = noise_levels[i % len(noise_levels)]
        if kind == "sphere":
            v = rng.normal(size=(n_points, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
            p, nrm = 0.8 * v, v
        elif kind == "torus":
            u, w = rng.random(n_points) * 2 * np.pi, rng.random(n_points) * 2 * np.pi
            p = np.stack([(1 + 0.3 * np.cos(w)) * np.cos(u), (1 + 0.3 * np.cos(w)) * np.sin(u), 0.3 * np.sin(w)], 1)
            nrm = np.stack([np.cos(w) * np.cos(u), np.cos(w) * np.sin(u), np.sin(w)], 1)
        elif kind == "ellipsoid":
            v = rng.normal(size=(n_points, 3)); v /= np.linalg.norm(v, axis=1, keepdims=True)
            ax = np.array([1.0, 0.7, 0.5])
            p = v * ax
            nrm = v / ax; nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        elif kind == "sheet":
            xy = rng.random((n_points, 2)) * 2 - 1
            z = 0.3 * np.sin(2 * xy[:, 0]) * np.cos(1.5 * xy[:, 1])
            gx = 0.6 * np.cos(2 * xy[:, 0]) * np.cos(1.5 * xy[:, 1]); gy = -0.45 * np.sin(2 * xy[:, 0]) * np.sin(1.5 * xy[:, 1])
            p = np.stack([xy[:, 0], xy[:, 1], z], 1)
            nrm = np.stack([-gx, -gy, np.ones(n_points)], 1); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
        else:
            t, h = rng.random(n_points) * 2 * np.pi, rng.random(n_points) * 2 - 1
            p = np.stack([0.5 * np.cos(t), 0.5 * np.sin(t), h], 1)
            nrm = np.stack([np.cos(t), np.sin(t), np.zeros(n_points)], 1)
        pts = _finish(p, nrm, noise, rng)
        name = "%s%dk_%02d_noise_%.4f" % (kind, n_points // 1000, i, noise)
        np.savetxt(os.path.join(root, name + ".xyz"), pts, fmt="%.6f")
        np.savetxt(os.path.join(root, name + ".normals"), nrm, fmt="%.6f")
        np.savetxt(os.path.join(root, name + ".pidx"), np.sort(rng.choice(n_points, min(n_sparse, n_points), replace=False)), fmt="%d")
        names.append(name)
    with open(os.path.join(root, list_name), "w") as fh:
        fh.write("\n".join(names) + "\n")
    return names


def synthetic_state_dict(seed: int = 0, num_points: int = 256, jet_order: int = 3, head_gain: float = 2.0,
                         device="cuda"):
    """A reproducible random ``state_dict`` with the reference's 125 keys (models/DeepFit.py constructors):
    He-uniform conv/fc weights, randomised BatchNorm affine + running statistics, transform heads shrunk
    so the predicted rotations stay near identity, and the last layer rescaled so the predicted weights
    spread over (0.01, 1.01).  The rescaling runs the network itself -- on the GPU, through the CUDA
    library (there is no CPU path) -- on eight synthetic paraboloid patches and reads the logits back from
    the sigmoid weights."""
    import math
    from collections import OrderedDict

    import torch

    from .DeepFit import DeepFit
    g = torch.Generator().manual_seed(int(seed))

    def U(shape, bound):
        return (torch.rand(tuple(shape), generator=g) * 2 - 1) * bound

    model = DeepFit(k=1, num_points=num_points, use_point_stn=True, use_feat_stn=True, point_tuple=1, sym_op="max",
                    arch="simple", jet_order=jet_order, weight_mode="sigmoid")
    sd = OrderedDict()
    for name, t in model.state_dict().items():
        leaf = name.rsplit(".", 1)[1]
        is_bn = ".bn" in name or name.startswith("bn")
        if leaf == "num_batches_tracked":
            sd[name] = torch.tensor(100, dtype=torch.long)
        elif is_bn:
            sd[name] = {"weight": lambda: 1.0 + U(t.shape, 0.25), "bias": lambda: U(t.shape, 0.1),
                        "running_mean": lambda: U(t.shape, 0.1), "running_var": lambda: 1.0 + U(t.shape, 0.5)}[leaf]()
        elif leaf == "weight":
            sd[name] = U(t.shape, math.sqrt(6.0 / t.shape[1]))
        else:
            sd[name] = U(t.shape, 0.1)
    for p, shrink in (("feat.pointfeat.stn1.fc3", 0.02), ("feat.pointfeat.stn2.fc3", 0.01)):
        sd[p + ".weight"] *= shrink
        sd[p + ".bias"] *= shrink
    sd["conv1.weight"][:, 1024:, :] *= 4.0           # let the per-point branch matter next to the 1024 global channels
    sd["conv4.weight"] *= 0.05                       # keep the uncalibrated logits well inside the sigmoid's linear range
    sd["conv4.bias"] *= 0.0
    k = num_points
    xy = torch.rand((8, 2, k), generator=g) * 1.4 - 0.7
    c = torch.rand((8, 3, 1), generator=g) - 0.5
    zz = 0.3 * (c[:, 0:1] * xy[:, 0:1] ** 2 + c[:, 1:2] * xy[:, 1:2] ** 2 + c[:, 2:3] * xy[:, 0:1] * xy[:, 1:2])
    zz = zz + 0.02 * (torch.rand((8, 1, k), generator=g) - 0.5)
    model.load_state_dict(sd)
    model.to(device).eval()
    with torch.no_grad():
        w = model(torch.cat([xy, zz], 1).to(device))[2].double().cpu() - 0.01
    w = w.clamp(1e-6, 1 - 1e-6)
    a = torch.log(w / (1 - w))
    scale = head_gain / float(a.std())
    sd["conv4.bias"] = ((sd["conv4.bias"].double() - a.mean()) * scale).float()
    sd["conv4.weight"] = (sd["conv4.weight"].double() * scale).float()
    return sd


# ---------------------------------------------------------------------------------------------------------------
# Device-side generators for the large named workloads (10 M / 100 M points): same surfaces as above, sampled with a
# seeded torch.Generator on the GPU so that no rank spends minutes in numpy or ships gigabytes over PCIe.  Inputs of
# the benchmark, not part of the hot path.  (The Philox stream differs from numpy's, so these clouds are NOT
# point-for-point the numpy ones; the tests at small sizes use the numpy generators.)
def device_cloud(kind: str, n: int, seed: int, device="cuda", noise: float = 0.0):
    """float32 [n,3] on ``device``, normalised to the unit sphere like tutorial_utils.py:14-15 (mean in fp64).
    kind: 'torus' (config 1/2 shape), 'sphere', 'multi' (config 4: sphere + torus + saddle sheet + cylinder in a common
    box), 'density' (config 5: lidar-like 1/r^2 density falloff on a wavy ground sheet)."""
    import math

    import torch
    g = torch.Generator(device=device).manual_seed(int(seed))
    f64 = dict(dtype=torch.float64, device=device, generator=g)

    def rnd(m):
        return torch.rand(m, **f64)

    def torus(m, R, r):
        u, v = rnd(m) * (2 * math.pi), rnd(m) * (2 * math.pi)
        cu, su, cv, sv = torch.cos(u), torch.sin(u), torch.cos(v), torch.sin(v)
        return torch.stack([(R + r * cv) * cu, (R + r * cv) * su, r * sv], 1), torch.stack([cv * cu, cv * su, sv], 1)

    def sphere(m, rad):
        v = torch.randn((m, 3), **f64)
        v = v / v.norm(dim=1, keepdim=True)
        return v * rad, v

    if kind == "torus":
        p, nrm = torus(n, 1.0, 0.3)
    elif kind == "sphere":
        p, nrm = sphere(n, 1.0)
    elif kind == "multi":
        m = n // 4
        a, na = sphere(m, 0.6)
        a[:, 0] += 1.5
        b, nb = torus(m, 0.8, 0.25)
        b[:, 0] -= 1.5
        xy = rnd((m, 2)) * 2 - 1
        c = torch.stack([xy[:, 0], xy[:, 1] + 2.0, 0.4 * (xy[:, 0] ** 2 - xy[:, 1] ** 2)], 1)
        gx, gy = 0.8 * xy[:, 0], -0.8 * xy[:, 1]
        nc = torch.stack([-gx, -gy, torch.ones_like(gx)], 1)
        nc = nc / nc.norm(dim=1, keepdim=True)
        m2 = n - 3 * m
        t, h = rnd(m2) * (2 * math.pi), rnd(m2) * 2 - 1
        d = torch.stack([0.5 * torch.cos(t), 0.5 * torch.sin(t) - 2.0, h], 1)
        nd = torch.stack([torch.cos(t), torch.sin(t), torch.zeros_like(t)], 1)
        p, nrm = torch.cat([a, b, c, d]), torch.cat([na, nb, nc, nd])
        perm = torch.randperm(n, device=device, generator=g)      # file order must not follow the surfaces
        p, nrm = p[perm], nrm[perm]
    elif kind == "density":
        r = 1.0 / (rnd(n) * 0.98 + 0.02)                          # heavy near the sensor, out to 50 units
        th = rnd(n) * (2 * math.pi)
        x, y = r * torch.cos(th), r * torch.sin(th)
        z = 0.3 * torch.sin(0.4 * x) * torch.cos(0.3 * y) + 0.02 * torch.randn(n, **f64)
        p, nrm = torch.stack([x, y, z], 1), None
    else:
        raise ValueError("unknown cloud kind %r" % kind)
    if noise > 0 and nrm is not None:
        bb = (p.max(0)[0] - p.min(0)[0]).norm()
        p = p + nrm * (torch.randn((n, 1), **f64) * (noise * bb))
    p32 = p.float()
    bbdiag = (p32.max(0)[0] - p32.min(0)[0]).double().norm()
    mean = p32.double().mean(0)
    return ((p32.double() - mean) / (0.5 * bbdiag)).float().contiguous()
