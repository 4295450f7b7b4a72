"""Dump what the CUDA path produces on the per-row gate patches (tests/golden) so that a failing row can be analysed
offline: weights / trans / rotated points / beta / normal per precision -> gpurun_out/gate_outputs.npz."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from deepfit_b200 import DeepFit as D
    from deepfit_b200 import ops
    from oracle import deepfit_oracle as O
    golden = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
    out = {}
    layers = D.fold_batchnorm(O.seeded_state_dict(0))
    for name in ("Boxy_smooth100k", "0000000000"):
        P = torch.from_numpy(golden[name + "/gate_patch"]).cuda()
        for precision in ("f16x3", "bf16x3", "f16f8"):
            net = ops.WeightNetwork(layers, 256, "sigmoid", precision, max_batch=512)
            w, trans, _, pts = net.forward(P, want_trans2=False, want_points=True)
            net.status()
            res = ops.wls_fit(pts, w, 3)
            key = "%s/%s/" % (name, precision)
            out[key + "weights"] = w.cpu().numpy()
            out[key + "trans"] = trans.cpu().numpy()
            out[key + "pts"] = pts.cpu().numpy()
            out[key + "beta"] = res["beta"].cpu().numpy()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    np.savez_compressed(os.path.join(ROOT, "gpurun_out", "gate_outputs.npz"), **out)
    print("wrote gate_outputs.npz", {k_: v.shape for k_, v in out.items()})


if __name__ == "__main__":
    main()
