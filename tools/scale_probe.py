"""Stage timings on the larger synthetic clouds BASELINE.json names (configs 2, 4, 5 shapes), one GPU.

    python tools/scale_probe.py --points 10000000 --cloud multi --k 256 --order 3 --queries 1000000

Classic (network-free) path: grid build, kNN, centre/scale/PCA, jet fit -- CUDA-event timings per stage and
size-independent kNN properties on the measured queries (sorted distances, self first, no duplicates)."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--points", type=int, default=10_000_000)
    ap.add_argument("--cloud", type=str, default="multi", choices=["multi", "torus", "sphere", "density"])
    ap.add_argument("--k", type=int, default=256)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--queries", type=int, default=1_000_000)
    ap.add_argument("--chunk", type=int, default=65536)
    ap.add_argument("--morton", action="store_true", help="visit the queries in the grid's Morton order (explicit query lists)")
    args = ap.parse_args()
    from deepfit_b200 import ops, synthetic
    from deepfit_b200.tutorial_utils import normalize_cloud
    gen = {"multi": synthetic.multi_surface_cloud, "torus": synthetic.torus_cloud, "sphere": synthetic.sphere_cloud,
           "density": synthetic.variable_density_cloud}[args.cloud]
    pts_np = np.ascontiguousarray(normalize_cloud(gen(args.points))[0])
    pts = torch.from_numpy(pts_np).cuda()
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    e0, e1 = ev(), ev()
    grid = ops.Grid(pts)
    torch.cuda.synchronize()
    e0.record()
    grid.update(pts)
    e1.record()
    torch.cuda.synchronize()
    t_grid = e0.elapsed_time(e1)
    q = min(args.queries, args.points)
    # untimed warm-up chunk: the caching allocator's first cudaMallocs would otherwise land inside the events
    c0 = min(args.chunk, q)
    idx, rad, dist = grid.query(args.k, q_begin=0, q_count=c0, return_dist=True)
    patch, U = ops.patch_pca(pts, idx, rad, q_begin=0)
    ops.wls_fit(patch, None, args.order, U=U, rad=rad)
    del idx, rad, dist, patch, U
    torch.cuda.synchronize()
    order = grid.morton_order(0, q) if args.morton else None
    stage = {"knn": 0.0, "pca": 0.0, "wls": 0.0}
    worst_knn_chunk = 0.0
    ok = True
    for s in range(0, q, args.chunk):
        c = min(args.chunk, q - s)
        a, b, d, f = ev(), ev(), ev(), ev()
        a.record()
        qlist = order[s:s + c] if args.morton else None
        idx, rad, dist = grid.query(args.k, q_begin=s, q_count=c, query=qlist, return_dist=True)
        b.record()
        patch, U = ops.patch_pca(pts, idx, rad, q_begin=s, query=qlist)
        d.record()
        out = ops.wls_fit(patch, None, args.order, U=U, rad=rad)
        f.record()
        torch.cuda.synchronize()
        stage["knn"] += a.elapsed_time(b)
        worst_knn_chunk = max(worst_knn_chunk, a.elapsed_time(b))
        stage["pca"] += b.elapsed_time(d)
        stage["wls"] += d.elapsed_time(f)
        if s == 0:
            ok = ok and bool((dist[:, 1:] >= dist[:, :-1]).all()) and bool((dist[:, 0] == 0).all())
            srt = torch.sort(idx.long(), dim=1)[0]
            ok = ok and bool((srt[:, 1:] != srt[:, :-1]).all())
            ok = ok and bool(torch.isfinite(out["n_world"]).all())
    total = sum(stage.values())
    print(json.dumps({"cloud": args.cloud, "points": args.points, "k": args.k, "order": args.order, "queries": q, "morton": bool(args.morton),
                      "grid_build_ms": round(t_grid, 3), "stage_ms": {k_: round(v, 2) for k_, v in stage.items()},
                      "queries_per_s_classic": q / (total * 1e-3), "knn_queries_per_s": q / (stage["knn"] * 1e-3),
                      "worst_knn_chunk_ms": round(worst_knn_chunk, 2), "properties_ok": ok}))


if __name__ == "__main__":
    main()
