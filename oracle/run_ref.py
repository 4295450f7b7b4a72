"""TEST INFRASTRUCTURE ONLY -- time the UNMODIFIED reference script as shipped.

    python oracle/run_ref.py classic [cloud.xyz]          # tutorial/compute_normals.py --mode classic, whole cloud
    python oracle/run_ref.py DeepFit [cloud.xyz] [points] # --mode DeepFit on a prefix of the cloud (default 4096 points)

Applies the shims of oracle/shims.py (outside the reference), changes into /root/reference/tutorial and `runpy`s
tutorial/compute_normals.py with --gpu_idx -1: the script's own DataLoader (batch 128, 8 worker processes,
compute_normals.py:52), its train-mode network (it never calls model.eval()), its np.savetxt exports -- everything as
shipped.  The pretrained blob is missing from the checkout, so DeepFit mode is pointed at a temporary model directory
holding the reference's own DeepFit_params.pth and a DeepFit.pth made from the seeded state_dict; speed does not depend
on the weight values.  Prints one JSON line: points/s end to end (text load, tree build, loop, text export).

The reference tree exists only in the authoring container, so this cannot run on the GPU box: bench.py's reference arm
there is the oracle PORT of the same arithmetic with a batched cKDTree query (a FASTER baseline than this script, so the
reported GPU/CPU ratio is conservative); this file is how the figure for the script itself was obtained.
"""
import json
import os
import runpy
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shims  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "classic"
    cloud = sys.argv[2] if len(sys.argv) > 2 else os.path.join(shims.REFERENCE_ROOT, "tutorial", "Boxy_smooth100k.xyz")
    if not shims.reference_available():
        sys.exit("the reference tree is not present at %s" % shims.REFERENCE_ROOT)
    shims.apply_shims()
    import torch
    tmp = tempfile.mkdtemp(prefix="run_ref_")
    try:
        argv = ["compute_normals.py", "--output_path", os.path.join(tmp, "out"), "--gpu_idx", "-1", "--mode", mode,
                "--k_neighbors", "256", "--jet_order", "3", "--compute_curvatures", "True"]
        if mode == "DeepFit":
            from oracle import deepfit_oracle as O
            n_pts = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
            pts = np.loadtxt(cloud)[:n_pts]
            cloud = os.path.join(tmp, "prefix%d.xyz" % n_pts)
            np.savetxt(cloud, pts)
            model_dir = os.path.join(tmp, "model")
            os.makedirs(model_dir)
            shutil.copy(os.path.join(shims.REFERENCE_ROOT, "trained_models", "DeepFit", "DeepFit_params.pth"), model_dir)
            torch.save(O.seeded_state_dict(0), os.path.join(model_dir, "DeepFit.pth"))
            argv += ["--trained_model_path", model_dir]
        n = sum(1 for _ in open(cloud))
        argv += ["--input", cloud]
        tut = os.path.join(shims.REFERENCE_ROOT, "tutorial")
        os.chdir(tut)
        sys.path.insert(0, tut)
        sys.argv = argv
        t0 = time.time()
        runpy.run_path(os.path.join(tut, "compute_normals.py"), run_name="__main__")
        dt = time.time() - t0
        outs = sorted(os.listdir(os.path.join(tmp, "out")))
        print(json.dumps({"impl": "reference script, unmodified (tutorial/compute_normals.py --gpu_idx -1)", "mode": mode,
                          "points": n, "seconds": round(dt, 2), "points_per_s": round(n / dt, 1), "cpu_count": os.cpu_count(),
                          "torch_threads": torch.get_num_threads(), "dataloader_workers": 8, "outputs": outs}))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
