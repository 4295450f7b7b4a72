"""BASELINE config 3 at its named size: DeepFit jet order 3, k = 256, full test_c_est.py outputs (normals, curvatures,
weights, beta, trans) on 20 synthetic PCPNet-shaped 100 000-point clouds with Gaussian noise 0 / 0.0025 / 0.012 of the
bounding-box diagonal, through ``deepfit_b200.pcpnet.estimate_shapes`` (reference loop: test_c_est.py:104-218).

    python tools/config3_run.py [--shapes 20] [--points 100000] [--root /tmp/dfb_config3] [--no-write]

Prints one JSON line: points/s of the compute path alone (kNN ... curvature, device-synchronised), and the text I/O
reported separately (loading the .xyz files; writing the five export files per shape: 1.3 GB of '%.18e' text per dense
100 k-point shape, removed again shape by shape).  Seeded stand-in weights (the pretrained blob is missing from the
reference checkout).  Parity of the same path on the same kind of data: tests/test_gpu_config3.py."""
import argparse
import json
import os
import shutil
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfit_b200 import pcpnet, synthetic  # noqa: E402
from deepfit_b200.pipeline import model_from_state_dict  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", type=int, default=20)
    ap.add_argument("--points", type=int, default=100000)
    ap.add_argument("--root", type=str, default="/tmp/dfb_config3")
    ap.add_argument("--no-write", action="store_true", help="skip the text export (compute + load only)")
    ap.add_argument("--precision", type=str, default="f16x3")
    args = ap.parse_args()
    dev = torch.device("cuda")
    shutil.rmtree(args.root, ignore_errors=True)
    t0 = time.perf_counter()
    names = synthetic.write_pcpnet_like(os.path.join(args.root, "pclouds"), n_shapes=args.shapes, n_points=args.points)
    gen_s = time.perf_counter() - t0
    sd = synthetic.synthetic_state_dict(0, 256, 3, device=dev)
    model = model_from_state_dict(sd, 256, 3, "sigmoid", args.precision, dev)
    ds = pcpnet.PointcloudPatchDataset(os.path.join(args.root, "pclouds"), "testset_synthetic.txt", points_per_patch=256,
                                       use_pca=True, sparse_patches=False, neighbor_search_method='k', device=dev)
    out_dir = os.path.join(args.root, "results")
    # warm-up on the first shape (kernel attributes, workspace, file cache), not timed
    one = pcpnet.PointcloudPatchDataset(os.path.join(args.root, "pclouds"), "testset_synthetic.txt", points_per_patch=256,
                                        use_pca=True, sparse_patches=False, neighbor_search_method='k', device=dev)
    one.shape_names = one.shape_names[:1]
    one.shape_patch_count = one.shape_patch_count[:1]
    pcpnet.estimate_shapes(one, model, out_dir, write=False, keep_results=False)
    timings = {}
    state = {"exported": 0, "finite": True, "other_s": 0.0}

    def on_shape(name, arrays):
        # not part of the path: the run's own sanity check, and the 0.7 GB of exports per shape removed as we go
        t = time.perf_counter()
        state["finite"] = state["finite"] and all(bool(np.isfinite(a).all()) for a in arrays.values())
        if not args.no_write:
            for ext in ("normals", "curv", "weights", "beta", "trans"):
                f = os.path.join(out_dir, name + "." + ext)
                state["exported"] += os.path.getsize(f)
                os.remove(f)
        state["other_s"] += time.perf_counter() - t

    t0 = time.perf_counter()
    pcpnet.estimate_shapes(ds, model, out_dir, write=not args.no_write, timings=timings, keep_results=False, on_shape=on_shape)
    exported, finite = state["exported"], state["finite"]
    wall = time.perf_counter() - t0 - state["other_s"]
    model._net.status()
    n_total = args.shapes * args.points
    print(json.dumps({
        "workload": "config3: DeepFit jet-3 k=256, full test_c_est outputs, %d synthetic PCPNet-shaped clouds of %d points, "
                    "noise 0 / 0.0025 / 0.012, dense (every point a patch centre), seeded state_dict" % (args.shapes, args.points),
        "points_total": n_total, "precision": args.precision,
        "compute_s": round(timings.get("compute", 0.0), 3), "compute_points_per_s": round(n_total / timings["compute"]),
        "load_text_s": round(timings.get("load", 0.0), 3), "export_text_s": round(timings.get("export", 0.0), 3),
        "export_bytes": exported, "export_GBps": round(exported / max(timings.get("export", 0.0), 1e-9) / 1e9, 3),
        "wall_s": round(wall, 3), "points_per_s_incl_text_io": round(n_total / wall), "dataset_generation_s": round(gen_s, 1),
        "all_finite": bool(finite)}))
    shutil.rmtree(args.root, ignore_errors=True)


if __name__ == "__main__":
    main()
