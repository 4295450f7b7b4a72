"""Per-layer precision ladder ON THE GPU: for each precision string (f16x3, f16mix:<layers>, ...) run the per-row gate of
tests/test_gpu_network.py::test_deepfit_forward_per_row_gate on the 512 golden patches of both tutorial clouds and print
the achieved maxima without asserting, so that one GPU call ranks every candidate layer set.

    python tools/mix_probe.py f16x3 f16mix:qstn_max f16mix:qstn_max+stn_max ...

gate_i = max(gate, 2 x the reference's own fp32 noise on row i) (oracle/make_golden.py); a mode is a parity mode when
every row of both clouds is inside its gate for the normal (0.01 deg) and the curvature (1e-3 rectified-relative).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from deepfit_b200 import DeepFit as D
    from deepfit_b200.pipeline import model_from_state_dict
    from oracle import deepfit_oracle as O
    golden = np.load(os.path.join(ROOT, "tests", "golden", "reference_outputs.npz"))
    sd = O.seeded_state_dict(0)
    for kv in filter(None, os.environ.get("DFB_DBG", "").split(",")):     # kernel A/B switches, e.g. DFB_DBG=11:0 (bench.py's format)
        from deepfit_b200 import _lib
        key, val = kv.split(":")
        _lib.load().dfb_dbg_set(int(key), int(val))
        print("dfb_dbg_set(%s, %s)" % (key, val))
    for precision in sys.argv[1:]:
        model = model_from_state_dict(sd, 256, 3, "sigmoid", precision)
        verdict = True
        for name in ("Boxy_smooth100k", "0000000000"):
            P = torch.from_numpy(golden[name + "/gate_patch"]).cuda()
            U = torch.from_numpy(golden[name + "/gate_U"])
            rad = golden[name + "/gate_rad"]
            normal, beta, weights, trans, trans2, _ = model(P)
            dw = (weights.cpu() - torch.from_numpy(golden[name + "/gate_weights"])).abs().max().item()
            dt = (trans.cpu() - torch.from_numpy(golden[name + "/gate_trans"])).abs().max().item()
            nw = O.world_normals(normal.cpu(), U, trans.cpu()).numpy()
            ang = O.unoriented_angle_deg(nw, golden[name + "/gate_normal_world"])
            c, _ = D.compute_principal_curvatures(beta)
            cerr = O.rectified_rel_err(c.cpu().numpy().astype(np.float64) / rad[:, None], golden[name + "/gate_curv_scaled"]).max(1)
            gate_a = np.maximum(0.01, 2 * golden[name + "/gate_floor_angle_deg"])
            gate_c = np.maximum(1e-3, 2 * golden[name + "/gate_floor_curv"])
            ra, rc = ang / gate_a, cerr / gate_c
            ok = bool((ra <= 1).all() and (rc <= 1).all())
            verdict = verdict and ok
            print("%-34s %-16s |dw| %.1e |dtrans| %.1e | angle/gate max %.3f p99 %.3f (%d rows > 0.5, %d > 1) | curv/gate max %.3f "
                  "p99 %.3f (%d rows > 0.5, %d > 1) | angle max %.2e deg, curv max %.2e | %s"
                  % (precision, name, dw, dt, ra.max(), np.percentile(ra, 99), int((ra > 0.5).sum()), int((ra > 1).sum()), rc.max(),
                     np.percentile(rc, 99), int((rc > 0.5).sum()), int((rc > 1).sum()), ang.max(), cerr.max(), "ok" if ok else "OVER"))
        print("%-34s => %s" % (precision, "every gate holds" if verdict else "NOT a parity mode"), flush=True)
        del model


if __name__ == "__main__":
    main()
