#!/usr/bin/env python
"""bench.py -- DeepFit hot path throughput on B200 (normals + curvatures per second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload torus|boxy|config2|config4|config5] [--scaling weak|strong]

Default workload (BASELINE.json metric "normals/sec (DeepFit jet-3, k=256)", the shape of configs[0] on the config-2
synthetic surface): a seeded noisy torus of ``--points`` points PER GPU (weak scaling: N GPUs index an N-times larger,
replicated cloud and each rank owns a contiguous point-index range), DeepFit mode, jet order 3, k=256, seeded synthetic
state_dict (the pretrained blob is missing from the reference checkout), parity-grade precision.  The other workloads
are the named BASELINE configs at their named sizes:

    boxy     configs[0] itself: the tutorial cloud Boxy_smooth100k (100 000 points, from tests/golden), DeepFit o3 k=256
    config2  classic (unweighted) jet, --order 1..4, k=256, 1 M-point noisy torus / sphere (--cloud)
    config4  DeepFit order 4, k=512, 10 M-point multi-surface scan; --scaling strong shards the FIXED cloud over N GPUs
    config5  DeepFit order 3, k=256, 100 M-point variable-density scan, fixed cloud sharded over N GPUs + all-gather

One step = grid build + kNN + centre/scale/PCA + weight network + weighted jet fit + normal/curvature for every point
of the rank's range through ONE ``dfb_pipeline_run`` call, written in place into the all-gather buffers, + the NCCL
all-gather of the outputs when N > 1.

Prints ONE JSON line (see the driver contract).  ``value`` is measured with the cloud resident in HBM; ``e2e`` is the
same number of steps with the pinned-host -> device copy of the cloud and the device -> host read of the (gathered)
outputs inside the timed region.  ``--impl reference`` times the CPU restatement of the reference path (oracle port:
scipy cKDTree + torch CPU, all host threads) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

UNIT = "points/s"

# algorithmic tensor-pipe work of the weight network per patch (split head)
MLP_FLOP_PER_PATCH = {256: 326.9e6, 512: 648.0e6}


def chain_flop_per_patch(k: int) -> float:
    """The three fused chain kernels of one forward: three 128->1024 layers + three 64->128 + four 64->64 (incl. x*T2)
    + three K=3 layers, all over k points."""
    return 2.0 * k * (3 * 128 * 1024 + 3 * 64 * 128 + 4 * 64 * 64 + 3 * 3 * 64)


WORKLOADS = {
    #            cloud      points (per GPU when weak)  k    order  mode       default scaling
    "torus":   ("torus",    262144,                     256, 3,     "DeepFit", "weak"),
    "boxy":    ("boxy",     100000,                     256, 3,     "DeepFit", "strong"),
    "config2": ("torus",    1_000_000,                  256, 3,     "classic", "strong"),
    "config4": ("multi",    10_000_000,                 512, 4,     "DeepFit", "strong"),
    "config5": ("density",  100_000_000,                256, 3,     "DeepFit", "strong"),
}


def metric_name(mode: str, order: int, k: int) -> str:
    return "normals_per_sec_%s_jet%d_k%d" % ("deepfit" if mode == "DeepFit" else "classic", order, k)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            d = json.load(fh)
        return d.get("hbm_gbs", 6650.0), d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_cloud_host(kind: str, total_points: int) -> np.ndarray:
    """Host-side (numpy) cloud for the small workloads and the CPU arm: normalised float32 [n,3]."""
    from deepfit_b200 import synthetic
    from deepfit_b200.tutorial_utils import normalize_cloud
    if kind == "boxy":
        raw = np.load(os.path.join(ROOT, "tests", "golden", "tutorial_clouds.npz"))["Boxy_smooth100k"]
    elif kind == "torus":
        raw = synthetic.torus_cloud(total_points, seed=1234, noise=0.005)
    elif kind == "sphere":
        raw = synthetic.sphere_cloud(total_points, seed=1234, noise=0.005)
    elif kind == "multi":
        raw = synthetic.multi_surface_cloud(total_points, seed=4321)
    else:
        raw = synthetic.variable_density_cloud(total_points, seed=9876)
    return np.ascontiguousarray(normalize_cloud(raw)[0], dtype=np.float32)


def cpu_reference_rate(points: np.ndarray, sample: int, k: int, order: int, mode: str, seed: int = 0, sd=None):
    """The reference's CPU path (oracle port) on `sample` query points of the same cloud, all host threads."""
    from oracle import deepfit_oracle as O
    if sd is None and mode == "DeepFit":
        sd = O.seeded_state_dict(0)
    rng = np.random.default_rng(seed)
    q = np.sort(rng.choice(points.shape[0], sample, replace=False))
    import scipy.spatial as spatial
    t0 = time.perf_counter()
    tree = spatial.cKDTree(points, 10)                      # tutorial_utils.py:16 (amortised over the cloud in reality)
    t_tree = time.perf_counter() - t0
    t0 = time.perf_counter()
    with torch.no_grad():
        dist, idx = tree.query(points[q], k=k, workers=-1)
        for s in range(0, sample, 256):
            items = [O.dataset_item(points, int(c), k, idx[s + j], dist[s + j]) for j, c in enumerate(q[s:s + 256])]
            P = torch.stack([it[0] for it in items])
            U = torch.stack([it[1] for it in items])
            if mode == "DeepFit":
                n, beta, w, trans, t2, _ = O.deepfit_forward(sd, P, order)
                O.world_normals(n, U, trans)
            else:
                beta, n = O.fit_wjet(P, torch.ones_like(P[:, 0]), order)
                O.world_normals(n, U)
            if order > 1:
                O.principal_curvatures(beta)
    dt = time.perf_counter() - t0
    return sample / dt, dt, t_tree


def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every host core."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    torch.set_num_threads(max(1, n))
    return torch.get_num_threads()


def ncu_traffic(kernel: str, patches_per_launch: int) -> dict:
    """`roofline.traffic` = DRAM bytes per launch of `kernel` from the committed ncu capture, or null when there is none."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "ncu_traffic.json")) as fh:
            t = json.load(fh)[kernel]
        return {"traffic": t["dram_bytes_per_launch"] * patches_per_launch / t["patches_per_launch"],
                "traffic_source": "%s: mean dram__bytes_read.sum + dram__bytes_write.sum of %d launches of %d patches under "
                                  "`ncu --set full`, scaled to %d patches per launch; not measured in this run"
                                  % (t["source"], t["launches_captured"], t["patches_per_launch"], patches_per_launch)}
    except (OSError, KeyError, ValueError):
        return {"traffic": None, "traffic_source": "no committed ncu capture (profiles/ncu_traffic.json)"}


def resolve_workload(args, world):
    cloud, points, k, order, mode, scaling = WORKLOADS[args.workload]
    if args.points:
        points = args.points
    if args.cloud:
        cloud = args.cloud
    if args.order and mode == "classic":
        order = args.order
    if args.scaling:
        scaling = args.scaling
    if cloud == "boxy":
        scaling, points = "strong", 100000
    n_total = points * world if scaling == "weak" else points
    return cloud, n_total, k, order, mode, scaling


def workload_text(args, cloud, n_total, k, order, mode, scaling, world):
    what = {"torus": "synthetic noisy torus (seed 1234, sigma 0.005 bbdiag)", "sphere": "synthetic noisy sphere (seed 1234)",
            "boxy": "tutorial cloud Boxy_smooth100k (BASELINE configs[0])", "multi": "synthetic multi-surface scan (sphere + torus + saddle sheet + cylinder, seed 4321)",
            "density": "synthetic variable-density scan (1/r^2 falloff, seed 9876)"}[cloud]
    return "%s: %s jet-%d k=%d on %s, %d points total (%s scaling over %d GPU%s), %s, all five stages incl. grid build" % (
        args.workload, mode, order, k, what, n_total, scaling, world, "" if world == 1 else "s",
        "seeded state_dict" if mode == "DeepFit" else "no network")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    use_all_host_threads()
    cloud, n_total, k, order, mode, scaling = resolve_workload(args, args.gpus)
    n_host = min(n_total, 4_000_000)      # the CPU arm indexes a bounded prefix-sized cloud of the same generator
    pts = make_cloud_host(cloud, n_host)
    sample = args.ref_sample
    for _ in range(args.warmup):
        cpu_reference_rate(pts, min(sample, 256), k, order, mode)
    t_total = 0.0
    for i in range(args.steps):
        r, dt, _ = cpu_reference_rate(pts, sample, k, order, mode, seed=i)
        t_total += dt
    value = sample * args.steps / t_total
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": metric_name(mode, order, k), "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / args.steps, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f32", "data": "synthetic" if cloud != "boxy" else "tutorial cloud",
        "config": {"workload": workload_text(args, cloud, n_total, k, order, mode, scaling, args.gpus),
                   "sample": "%d query points per step" % sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d random query points per step of a %d-point cloud of the workload's generator; cKDTree(workers=-1) + "
                                   "torch CPU (%d threads), eval-mode forward, batch 256; tree build excluded" % (sample, pts.shape[0], cores)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------ checks & baselines

def knn_bruteforce_device(pts: torch.Tensor, queries: torch.Tensor, k: int):
    """Exact kNN by brute force in torch fp64 ON THE DEVICE (library ops as the checker, not our kernel): rank key
    (d2 accumulated x->y->z without FMA contraction, index), like oracle.knn_bruteforce.  Returns idx i64 [Q,k]."""
    P = pts.double()
    out = torch.empty((queries.numel(), k), dtype=torch.int64, device=pts.device)
    step = max(1, min(64, (1 << 28) // max(1, pts.shape[0])))
    for s in range(0, queries.numel(), step):
        q = P[queries[s:s + step].long()]
        dx = q[:, None, 0] - P[None, :, 0]
        d2 = dx * dx
        dy = q[:, None, 1] - P[None, :, 1]
        d2 = d2 + dy * dy
        dz = q[:, None, 2] - P[None, :, 2]
        d2 = d2 + dz * dz
        del dx, dy, dz
        extra = min(64, pts.shape[0] - k)
        vals, idx = torch.topk(d2, k + extra, dim=1, largest=False)
        # canonical order among candidates: (value, index); ties beyond the k + extra window fall back to a full sort
        for r in range(vals.shape[0]):
            v, i = vals[r], idx[r]
            if extra and v[k - 1] == v[-1]:
                cand = torch.nonzero(d2[r] <= v[k - 1]).flatten()
                v, i = d2[r][cand], cand
            order = torch.argsort(i, stable=True)
            v, i = v[order], i[order]
            order = torch.argsort(v, stable=True)
            out[s + r] = i[order][:k]
        del d2
    return out


def spot_check(est, pts, begin, normals, curv, k, order, mode, sd_cpu, n_check=256):
    """After the timed region: `n_check` of this rank's outputs against the checkers -- kNN vs an on-device fp64 brute
    force, normals / curvatures vs the oracle (CPU) on our patches -- plus isfinite over everything that was produced."""
    from deepfit_b200 import ops
    from oracle import deepfit_oracle as O
    n_check = min(n_check, normals.shape[0])
    idx, rad = est.grid.query(k, q_begin=begin, q_count=n_check)
    q = torch.arange(begin, begin + n_check, device=pts.device)
    knn_ok = bool(torch.equal(idx.long(), knn_bruteforce_device(pts, q, k)))
    patch, U = ops.patch_pca(pts, idx, rad, q_begin=begin)
    with torch.no_grad():
        if mode == "DeepFit":
            n, beta, w, trans, t2, _ = O.deepfit_forward(sd_cpu, patch.cpu(), order)
            ref_n = O.world_normals(n, U.cpu(), trans).numpy()
        else:
            beta, n = O.fit_wjet(patch.cpu(), torch.ones(n_check, k), order)
            ref_n = O.world_normals(n, U.cpu()).numpy()
        ang = O.unoriented_angle_deg(normals[:n_check].cpu().numpy(), ref_n)
        res = {"points": n_check, "knn_exact_vs_fp64_bruteforce": knn_ok, "max_angle_deg": float(ang.max()),
               "median_angle_deg": float(np.median(ang))}
        # the same reference algebra evaluated in float64: how far the REFERENCE's fp32 result is from the exact one on
        # each of these rows (its noise floor), and the per-row gates of the parity tests, max(gate, 2 x floor)
        p64, U64 = patch.cpu().double(), U.cpu().double()
        if mode == "DeepFit":
            sd64 = {k_: (v.double() if v.is_floating_point() else v) for k_, v in sd_cpu.items()}
            n64, beta64, _, trans64, _, _ = O.deepfit_forward(sd64, p64, order)
            ex_n = O.world_normals(n64, U64, trans64).numpy()
        else:
            beta64, n64 = O.fit_wjet(p64, torch.ones(n_check, k, dtype=torch.float64), order)
            ex_n = O.world_normals(n64, U64).numpy()
        gate_a = np.maximum(0.01, 2 * O.unoriented_angle_deg(ref_n, ex_n))
        res["worst_angle_over_row_gate"] = float((ang / gate_a).max())
        if order > 1:
            ref_c = (O.principal_curvatures(beta).double() / rad.cpu()[:, None]).numpy()
            ex_c = (O.principal_curvatures(beta64).double() / rad.cpu()[:, None]).numpy()
            ours_c = curv[:n_check].cpu().numpy()
            err = O.rectified_rel_err(ours_c, ref_c).max(1)
            floor_c = O.rectified_rel_err(ref_c, ex_c).max(1)
            res["max_curv_err"] = float(err.max())
            res["worst_curv_over_row_gate"] = float((err / np.maximum(1e-3, 2 * floor_c)).max())
            res["max_curv_err_reference_vs_fp64"] = float(floor_c.max())
            res["max_curv_err_ours_vs_fp64"] = float(O.rectified_rel_err(ours_c, ex_c).max())
    res["all_finite"] = bool(torch.isfinite(normals).all()) and (curv is None or bool(torch.isfinite(curv).all()))
    return res


def library_baseline(sd, patch, order, iters=3):
    """SURVEY.md section 2b / BASELINE.md: the reference's own PyTorch-eager arithmetic (cuDNN / cuBLAS conv1d + linear
    + batch_norm as the reference module writes them, incl. the 1088-channel concat, then fit_Wjet with a batched
    Cholesky) on the SAME B200, fp32 and TF32 -- the 'library kernel' bar for stages 3-5.  Patches/s for one chunk."""
    import torch.nn.functional as F
    dev = patch.device
    sd = {k_: v.to(dev) for k_, v in sd.items()}

    def bn(p, x):
        return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)

    def conv(p, x):
        return F.conv1d(x, sd[p + ".weight"], sd[p + ".bias"])

    def fc(p, x):
        return F.linear(x, sd[p + ".weight"], sd[p + ".bias"])

    def trunk(p, x):
        x = F.relu(bn(p + ".bn1", conv(p + ".conv1", x)))
        x = F.relu(bn(p + ".bn2", conv(p + ".conv2", x)))
        x = F.relu(bn(p + ".bn3", conv(p + ".conv3", x)))
        g = x.max(dim=2)[0]
        g = F.relu(bn(p + ".bn4", fc(p + ".fc1", g)))
        g = F.relu(bn(p + ".bn5", fc(p + ".fc2", g)))
        return fc(p + ".fc3", g)

    def forward(points):
        B, _, k = points.shape
        q = trunk("feat.pointfeat.stn1", points)
        q = q + q.new_tensor([1.0, 0.0, 0.0, 0.0])
        s = 2.0 / (q * q).sum(1)
        h = q.unsqueeze(2) * q.unsqueeze(1)
        R = torch.stack([1 - (h[:, 2, 2] + h[:, 3, 3]) * s, (h[:, 1, 2] - h[:, 3, 0]) * s, (h[:, 1, 3] + h[:, 2, 0]) * s,
                         (h[:, 1, 2] + h[:, 3, 0]) * s, 1 - (h[:, 1, 1] + h[:, 3, 3]) * s, (h[:, 2, 3] - h[:, 1, 0]) * s,
                         (h[:, 1, 3] - h[:, 2, 0]) * s, (h[:, 2, 3] + h[:, 1, 0]) * s, 1 - (h[:, 1, 1] + h[:, 2, 2]) * s], 1).view(B, 3, 3)
        pts = torch.bmm(points.transpose(2, 1), R).transpose(2, 1).contiguous()
        x = F.relu(bn("feat.pointfeat.bn1", conv("feat.pointfeat.conv1", pts)))
        x = F.relu(bn("feat.pointfeat.bn2", conv("feat.pointfeat.conv2", x)))
        t2 = trunk("feat.pointfeat.stn2", x)
        t2 = (t2 + torch.eye(64, device=dev).reshape(1, 4096)).reshape(B, 64, 64)
        pf = torch.bmm(x.transpose(2, 1), t2).transpose(2, 1)
        y = F.relu(bn("feat.bn2", conv("feat.conv2", pf)))
        y = bn("feat.bn3", conv("feat.conv3", y))
        z = torch.cat([y.max(dim=2, keepdim=True)[0].expand(-1, -1, k), pf], 1)
        z = F.relu(bn("bn1", conv("conv1", z)))
        z = F.relu(bn("bn2", conv("conv2", z)))
        z = F.relu(bn("bn3", conv("conv3", z)))
        w = 0.01 + torch.sigmoid(conv("conv4", z).squeeze(1))
        # fit_Wjet (models/DeepFit.py:30-88), order 3 column order
        xx, yy, zz = pts[:, 0], pts[:, 1], pts[:, 2]
        hh = (xx.abs().mean(1) + yy.abs().mean(1)) / 2
        xx, yy = xx / hh[:, None], yy / hh[:, None]
        cols = [xx, yy, xx * xx, yy * yy, xx * yy]
        if order >= 3:
            cols += [xx ** 3, yy ** 3, xx * xx * yy, yy * yy * xx]
        if order >= 4:
            cols += [xx ** 4, yy ** 4, xx ** 3 * yy, yy ** 3 * xx, yy * yy * xx * xx]
        A = torch.stack(cols + [torch.ones_like(xx)], 2)
        XtX = A.transpose(1, 2) @ (w[:, :, None] * A)
        XtY = A.transpose(1, 2) @ (w * zz)[:, :, None]
        L, _ = torch.linalg.cholesky_ex(XtX)
        return torch.cholesky_solve(XtY, L)

    res = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        with torch.no_grad():
            for name, tf32 in (("fp32", False), ("tf32", True)):
                torch.backends.cuda.matmul.allow_tf32 = tf32
                torch.backends.cudnn.allow_tf32 = tf32
                for _ in range(2):
                    forward(patch)
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                torch.cuda.synchronize()
                ev[0].record()
                for _ in range(iters):
                    forward(patch)
                ev[1].record()
                torch.cuda.synchronize()
                res[name + "_patches_per_s"] = patch.shape[0] * iters / (ev[0].elapsed_time(ev[1]) * 1e-3)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    res["what"] = ("the reference's PyTorch-eager arithmetic for stages 3-5 (conv1d / linear / batch_norm as written, 1088-channel "
                   "concat, batched Cholesky fit) on this B200, batch %d, k=%d; kNN and PCA (CPU in the reference) excluded" %
                   (patch.shape[0], patch.shape[2]))
    return res


def mix_layers(precision):
    """The chain layers a precision runs with the hi*hi product only (f16mix), by name; None for the other modes."""
    name, _, layers = precision.partition(":")
    if name != "f16mix":
        return None
    return sorted(layers.split("+")) if layers else ["qstn_max", "stn_max"]


def chain_product_units(precision):
    """Tensor-core MMAs executed per algorithmic product, averaged over the MACs of the three chain kernels (per point:
    128 -> 1024 = 131 072 in each chain; small tensor-core layers 8 192 QSTN, 12 288 feature STN, 16 384 encoder)."""
    name = precision.partition(":")[0]
    if name != "f16mix":
        return {"f16x3": 3, "bf16x3": 3, "f16f8": 2, "bf16": 1}[name]
    macs = {"qstn_small": 8192, "qstn_max": 131072, "stn_small": 12288, "stn_max": 131072, "enc_small": 16384, "enc_max": 131072}
    single = set(mix_layers(precision))
    return sum(m * (1 if l in single else 3) for l, m in macs.items()) / sum(macs.values())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", type=str, default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", type=str, default="torus", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", type=str, default="", choices=["", "weak", "strong"])
    ap.add_argument("--points", type=int, default=0, help="points per GPU (weak) or in total (strong); 0 = the workload's size")
    ap.add_argument("--cloud", type=str, default="", help="override the workload's surface: torus | sphere | multi | density")
    ap.add_argument("--order", type=int, default=0, help="jet order of the classic workload (config2)")
    ap.add_argument("--precision", type=str, default="f16x3",
                    help="f16x3 (split FP16, 3 products, 22-bit operands) | f16mix[:layer+layer] (f16x3 operands, the named chain "
                         "layers with the hi*hi product only; default layers: qstn_max+stn_max) | bf16x3 | f16f8 | bf16 (approximations)")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--ref-sample", type=int, default=2048)
    ap.add_argument("--e2e-steps", type=int, default=-1, help="steps of the end-to-end region (default: same as --steps)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--allow-dbg", action="store_true", help="accept a non-empty DFB_DBG (kernel experiments; recorded in the line)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    dbg = os.environ.get("DFB_DBG", "")
    if dbg and not args.allow_dbg:
        sys.exit("bench.py: DFB_DBG=%r switches kernel paths off; refusing to measure (pass --allow-dbg for experiments)" % dbg)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    from deepfit_b200 import _lib, ops, synthetic
    from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
    from deepfit_b200.sharding import GatherBuffers, shard_range
    for kv in dbg.split(","):
        if ":" in kv:
            _lib.load().dfb_dbg_set(int(kv.split(":")[0]), int(kv.split(":")[1]))

    cloud, n_total, K, ORDER, mode, scaling = resolve_workload(args, world)
    chunk = args.chunk or (16384 if mode == "DeepFit" else 262144)
    # the cloud: small workloads come from the host generators (numpy); the 10 M / 100 M ones are sampled on the device
    if cloud == "boxy" or n_total <= 4_000_000:
        pts_host = torch.from_numpy(make_cloud_host(cloud, n_total)).pin_memory()
        pts = pts_host.to(dev, non_blocking=True)
    else:
        seed = {"multi": 4321, "density": 9876}.get(cloud, 1234)
        pts = synthetic.device_cloud(cloud, n_total, seed, dev, noise=0.005 if cloud in ("torus", "sphere") else 0.0)
        if world > 1:        # one bit-identical cloud on every rank, whatever the generators did
            dist.broadcast(pts, 0)
        pts_host = torch.empty(pts.shape, dtype=pts.dtype, pin_memory=True)
        pts_host.copy_(pts)
    torch.cuda.synchronize()
    model = None
    state_dict = None
    if mode == "DeepFit":
        state_dict = synthetic.synthetic_state_dict(0, K, ORDER, device=dev)   # DeepFit.pth is not shipped with the reference
        model = model_from_state_dict(state_dict, K, ORDER, "sigmoid", args.precision, dev)
    est = NormalEstimator(pts, K, ORDER, mode, model, chunk=chunk)
    begin, end = shard_range(n_total, rank, world)
    bufs = GatherBuffers(n_total, world, dev, want_curv=ORDER > 1)
    out_n, out_c = bufs.local_views(rank)
    host_n = torch.empty(bufs.normals.shape, dtype=torch.float32, pin_memory=True) if rank == 0 else None
    host_c = torch.empty(bufs.curv.shape, dtype=torch.float64, pin_memory=True) if (rank == 0 and bufs.curv is not None) else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > L2 (126 MB)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(e2e: bool):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        flush.zero_()
        barrier()
        ev[0].record()
        if e2e:
            est.points = pts_host.to(dev, non_blocking=True)
        est.rebuild_grid()
        est.run(begin, end, out_normals=out_n, out_curv=out_c, check=False)
        ev[1].record()
        bufs.all_gather(rank)
        ev[2].record()
        if e2e and rank == 0:      # the writer rank reads the whole gathered result
            host_n.copy_(bufs.normals, non_blocking=True)
            if host_c is not None:
                host_c.copy_(bufs.curv, non_blocking=True)
        ev[3].record()
        torch.cuda.synchronize()
        return ev[0].elapsed_time(ev[3]), ev[1].elapsed_time(ev[2])

    def timed_region(e2e: bool, steps: int):
        per_step, gather = [], []
        for _ in range(steps):
            t, g = step(e2e)
            per_step.append(t)
            gather.append(g)
        mine = torch.tensor([sum(per_step), sum(gather)], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(allr, mine)
        else:
            allr = [mine]
        tot = [float(t[0]) for t in allr]
        return max(tot), tot, max(float(t[1]) for t in allr)

    for _ in range(args.warmup):
        step(False)
    if est.net is not None:
        est.net.status()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ops.launch_count(reset=True)
    if est.net is not None:
        est.net.profile(True)
    ms_total, ms_ranks, ms_gather = timed_region(False, args.steps)
    prof = est.net.profile_read() if est.net is not None else {}
    if est.net is not None:
        est.net.profile(False)
        est.net.status()
    launches = ops.launch_count(reset=True)
    clocks = sampler.stop() if rank == 0 else None
    value = n_total * args.steps / (ms_total * 1e-3)
    timed_n, timed_c = out_n.clone(), (out_c.clone() if out_c is not None else None)   # what the timed region produced

    # per-stage split (one extra, untimed-for-the-headline pass with events between stages, first chunk of the range)
    def staged_pass():
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        torch.cuda.synchronize()
        evs[0].record()
        est.rebuild_grid()
        evs[1].record()
        c = min(chunk, end - begin)
        # the pipeline visits the queries in the grid's Morton order; so does this split
        qlist = est.grid.morton_order(begin, end)[:c].contiguous()
        torch.cuda.synchronize()
        evs[1].record()
        idx, rad = est.grid.query(K, query=qlist)
        evs[2].record()
        patch, U = ops.patch_pca(pts, idx, rad, query=qlist)
        evs[3].record()
        w = trans = None
        if est.net is not None:
            w, trans, _, _ = est.net.forward(patch, want_trans2=False)
        evs[4].record()
        ops.wls_fit(patch, w, ORDER, trans=trans, U=U, rad=rad)
        evs[5].record()
        torch.cuda.synchronize()
        names = ["grid_build_ms", "knn_ms", "pca_ms", "network_ms", "wls_ms"]
        d = {n_: evs[i].elapsed_time(evs[i + 1]) for i, n_ in enumerate(names)}
        d["chunk_points"] = c
        return d, patch
    est.points = pts
    staged_pass()                       # first pass pays the allocator's cudaMallocs for these shapes
    stages, patch0 = staged_pass()

    e2e_steps = args.steps if args.e2e_steps < 0 else args.e2e_steps
    e2e_value = None
    if e2e_steps > 0:
        e2e_ms, _, _ = timed_region(True, e2e_steps)
        e2e_value = n_total * e2e_steps / (e2e_ms * 1e-3)
    est.points = pts

    if rank == 0:
        hbm, tf_burst, tf_sus, which = measured_peaks()
        n_patches_timed = (end - begin) * args.steps
        line = {
            "metric": metric_name(mode, ORDER, K), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": {"f16x3": "f16x3 (split-FP16 tensor-core MMA, three products per k-step, fp32 accumulate; fp32-grade: the parity mode)",
                      "f16f8": "f16f8 (fp16 tensor-core MMA + one e4m3 correction MMA per k-step, fp32 accumulate; weights 3e-4 off)",
                      "bf16x3": "bf16x3 (split-bf16 tensor-core MMA, three products per k-step, fp32 accumulate; weights 1e-4 off)",
                      "f16mix": "f16mix (f16x3 operands and products everywhere except the chain layers of config.precision_mix, which issue "
                                "only the FP16 hi*hi product; fp32 accumulate)",
                      "bf16": "bf16"}[args.precision.partition(":")[0]]
            if mode == "DeepFit" else "f64 moments / f32 solve",
            "data": "synthetic" if cloud != "boxy" else "tutorial cloud (tests/golden), seeded weights",
            "config": {"workload": workload_text(args, cloud, n_total, K, ORDER, mode, scaling, world),
                       "points_total": n_total, "points_this_rank": end - begin, "k": K, "jet_order": ORDER,
                       "precision": args.precision if mode == "DeepFit" else None,
                       "precision_mix": mix_layers(args.precision) if mode == "DeepFit" else None,
                       "chunk": chunk, "l2": "256 MiB flush buffer written before every step",
                       "sharding": "contiguous point-index ranges, replicated cloud+grid, outputs written in place into the all-gather "
                                   "buffers, one in-place all-gather per dtype (f32 normals 12 B/pt, f64 curvatures 16 B/pt)",
                       "dfb_dbg": dbg or None},
            "e2e": {"value": e2e_value, "unit": UNIT, "steps": e2e_steps, "h2d_bytes_per_step": int(n_total * 12),
                    "d2h_bytes_per_step": int(bufs.normals.numel() * 4 + (bufs.curv.numel() * 8 if bufs.curv is not None else 0)),
                    "note": "every rank copies the replicated cloud host->device; rank 0 reads the whole gathered result back"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "rank_ms_per_step": {"min": min(ms_ranks) / args.steps, "max": max(ms_ranks) / args.steps},
            "gather_ms_per_step": ms_gather / args.steps,
            "gather_bytes_per_rank": bufs.bytes_per_rank() if world > 1 else 0,
            "stages_one_chunk": stages,
        }
        if mode == "DeepFit":
            gm_ms, gm_n = prof["gemm_max_128x1024"]
            achieved = (n_patches_timed * chain_flop_per_patch(K)) / (gm_ms * 1e-3) / 1e12 if gm_ms > 0 else None
            net_ms = sum(v[0] for v in prof.values())
            units = chain_product_units(args.precision)
            line["roofline"] = {
                "bound": "tensor", "kernel": "chain_max_kernel<k/128> (points -> 64 -> [64 -> 64 ->] 128 -> 1024 + max-pool, persistent, "
                                             "3 launches per forward)",
                "achieved": achieved, "peak": tf_sus, "unit": "TFLOP/s", "frac": (achieved / tf_sus) if achieved else None,
                # dram__bytes_read.sum + dram__bytes_write.sum per launch: the counters need ncu, so this run cannot measure
                # them; the committed `ncu --set full` capture of the same kernels is quoted (profiles/ncu_traffic.json,
                # written by tools/ncu_traffic.py), scaled to this run's patches per launch, with its source
                **ncu_traffic("chain_max_kernel", min(chunk, end - begin)),
                "peak_source": "%s MEASURED_PEAKS.json bf16_tflops_sustained" % which,
                "note": "achieved = algorithmic FLOPs of the three chain kernels (%.1f MFLOP/patch at k=%d) / their CUDA-event time. "
                        "Precision %s executes %.3g tensor-core MMA(s) per algorithmic product of these kernels (22-bit operands = split "
                        "FP16, hi*hi + hi*lo + lo*hi; f16mix: one product in the layers of config.precision_mix), so frac tops out near %.3f at the sustained clock the peak "
                        "was measured at; executed_frac = frac x that count" % (chain_flop_per_patch(K) / 1e6, K, args.precision, units, 1.0 / units),
                "executed_frac": (achieved * units / tf_sus) if achieved else None,
                "product_units": units,
                "launches": int(gm_n), "avg_launch_ms": gm_ms / gm_n if gm_n else None,
                "share_of_network": gm_ms / net_ms if net_ms else None,
                "network_tflops_algorithmic": (n_patches_timed * MLP_FLOP_PER_PATCH.get(K, 0)) / (net_ms * 1e-3) / 1e12 if net_ms else None}
            line["kernel_ms"] = {k_: round(v[0], 3) for k_, v in prof.items()}
        else:
            b = 3180 if K == 256 else 12 * K + 108
            line["roofline"] = {"bound": "hbm", "kernel": "wls_kernel<order> (fused jet fit + normal + curvature)",
                                "achieved": stages["chunk_points"] * b / (stages["wls_ms"] * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                                **ncu_traffic("wls_kernel", stages["chunk_points"]),
                                "peak_source": "%s MEASURED_PEAKS.json hbm_gbs" % which,
                                "note": "classic path: %d algorithmic bytes per patch (SURVEY 8d); kNN (instruction-bound) dominates "
                                        "this path's time, see stages_one_chunk" % b}
            line["roofline"]["frac"] = line["roofline"]["achieved"] / hbm
        # HBM-bound stages against SURVEY.md section 8(d)'s algorithmic bytes per patch: kNN + gather/PCA unfused convention
        # (12 + 12k read, 4k + 12k + 44 written), jet fit + epilogue (12k + 4k + 80 read, 4M + 24 written)
        b12 = 12 + 12 * K + 4 * K + 12 * K + 44
        b45 = 12 * K + (4 * K if mode == "DeepFit" else 0) + 80 + 4 * ops.JET_M[ORDER] + 24
        line["stage_rooflines"] = {
            "knn_pca": {"bound": "hbm", "bytes_per_patch": b12, "unit": "GB/s", "peak": hbm,
                        "achieved": stages["chunk_points"] * b12 / ((stages["knn_ms"] + stages["pca_ms"]) * 1e-3) / 1e9,
                        "queries_per_s": stages["chunk_points"] / (stages["knn_ms"] * 1e-3)},
            "wls": {"bound": "hbm", "bytes_per_patch": b45, "unit": "GB/s", "peak": hbm,
                    "achieved": stages["chunk_points"] * b45 / (stages["wls_ms"] * 1e-3) / 1e9},
        }
        for v in line["stage_rooflines"].values():
            v["frac"] = v["achieved"] / v["peak"]
        sd_cpu = {k_: v.cpu() for k_, v in state_dict.items()} if state_dict is not None else None
        if not args.no_check:
            line["check"] = spot_check(est, pts, begin, timed_n, timed_c, K, ORDER, mode, sd_cpu)
        if mode == "DeepFit" and world == 1 and not args.no_library_baseline and ORDER in (3, 4):
            lb = library_baseline(state_dict, patch0[:1024].contiguous(), ORDER)
            own = stages["chunk_points"] / ((stages["network_ms"] + stages["wls_ms"]) * 1e-3)
            lb["ours_stage345_patches_per_s"] = own
            lb["speedup_vs_fp32"] = own / lb["fp32_patches_per_s"]
            lb["speedup_vs_tf32"] = own / lb["tf32_patches_per_s"]
            line["library_baseline"] = lb
        if not args.no_cpu_baseline and world == 1:   # reported on rank 0 at N=1 only
            use_all_host_threads()
            host_pts = pts_host.numpy() if n_total <= 4_000_000 else make_cloud_host(cloud, 4_000_000)
            r, dt, t_tree = cpu_reference_rate(host_pts, args.ref_sample, K, ORDER, mode, sd=sd_cpu)
            line["cpu_baseline"] = {"value": r, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "%d random query points of a %d-point cloud of the same generator in %.1f s; "
                                              "cKDTree(workers=-1) + torch CPU, eval-mode forward, batch 256; tree build (%.2f s) excluded"
                                              % (args.ref_sample, host_pts.shape[0], dt, t_tree)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
