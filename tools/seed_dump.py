"""Dump the CUDA path's outputs for tests/test_gpu_pipeline.py::test_deepfit_pipeline_other_weight_sets (lidar cloud,
given seeds) -> gpurun_out/seed_dump.npz, so that the per-row fp32 noise floors of the reference can be studied offline
with any number of neighbour permutations (oracle/studies/floor_estimate.py)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from deepfit_b200 import ops
    from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
    from oracle import deepfit_oracle as O
    clouds = np.load(os.path.join(ROOT, "tests", "golden", "tutorial_clouds.npz"))
    out = {}
    for name in ("0000000000", "Boxy_smooth100k"):
        pts = torch.from_numpy(O.normalize_cloud(clouds[name])[0]).cuda()
        for seed in (0, 1, 2, 3, 4):
            sd = O.seeded_state_dict(seed)
            begin, count = 40000 + 1000 * seed, 384
            for precision in ("f16x3", "bf16x3"):
                model = model_from_state_dict(sd, 256, 3, "sigmoid", precision)
                est = NormalEstimator(pts, 256, mode="DeepFit", model=model, chunk=256)
                extras = {}
                normals, curv = est.run(begin, begin + count, extras=extras, want_extras=("beta", "weights", "trans"))
                key = "%s/seed%d/%s/" % (name, seed, precision)
                out[key + "normals"] = normals.cpu().numpy()
                out[key + "curv"] = curv.cpu().numpy()
                out[key + "weights"] = extras["weights"].cpu().numpy()
                out[key + "beta"] = extras["beta"].cpu().numpy()
            idx, rad = est.grid.query(256, q_begin=begin, q_count=count)
            patch, U = ops.patch_pca(pts, idx, rad, q_begin=begin)
            key = "%s/seed%d/" % (name, seed)
            out[key + "patch"] = patch.cpu().numpy()
            out[key + "U"] = U.cpu().numpy()
            out[key + "rad"] = rad.cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "seed_dump.npz"), **out)
    print("wrote seed_dump.npz:", len(out), "arrays")


if __name__ == "__main__":
    main()
