"""How well do a few neighbour permutations estimate the reference's per-row fp32 noise floor?

Test infrastructure (CPU).  Input: gpurun_out/seed_dump.npz (tools/seed_dump.py: the CUDA path's normals / curvatures on
384 points of each tutorial cloud under the seeded state_dicts 0-4, in f16x3 and bf16x3, with the patches it saw).
For every (cloud, seed) the reference algebra (oracle, fp32) is evaluated on the same patches, on N_PERM neighbour
permutations of them and in fp64; floor_i(n) = max deviation from the fp64 result among the fp32 forward and the first n
permutations; gate_i(n) = max(gate, 2 floor_i(n)).  Printed: the CUDA path's worst error / gate_i(n) for n = 3, 8, 16, 32,
and the same statistic for a FRESH fp32 evaluation of the reference itself (permutation n+1 .. ) -- the reference's own
chance of "failing" a gate estimated from n samples.

    python oracle/studies/floor_estimate.py [n_perm=32]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import deepfit_oracle as O  # noqa: E402


def main():
    n_perm = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    d = np.load(os.path.join(ROOT, "gpurun_out", "seed_dump.npz"))
    torch.set_num_threads(os.cpu_count())
    for name in ("Boxy_smooth100k", "0000000000"):
        for seed in (0, 1, 2, 3, 4):
            key = "%s/seed%d/" % (name, seed)
            P = torch.from_numpy(d[key + "patch"])
            U = torch.from_numpy(d[key + "U"])
            rad = torch.from_numpy(d[key + "rad"])
            sd = O.seeded_state_dict(seed)
            sd64 = {k_: (v.double() if v.is_floating_point() else v) for k_, v in sd.items()}

            def fwd(Pp, s=sd, dt=torch.float32):
                with torch.no_grad():
                    n, beta, w, trans, t2, _ = O.deepfit_forward(s, Pp.to(dt), 3)
                return (O.world_normals(n, U.to(dt), trans).numpy(),
                        (O.principal_curvatures(beta).double() / rad[:, None]).numpy())
            ref_n, ref_c = fwd(P)
            ex_n, ex_c = fwd(P, sd64, torch.float64)
            g = torch.Generator().manual_seed(5)
            dev_a = [O.unoriented_angle_deg(ref_n, ex_n)]
            dev_c = [O.rectified_rel_err(ref_c, ex_c).max(1)]
            perm_n, perm_c = [], []
            for _ in range(n_perm):
                pn, pc = fwd(P[:, :, torch.randperm(P.shape[2], generator=g)].contiguous())
                perm_n.append(pn); perm_c.append(pc)
                dev_a.append(O.unoriented_angle_deg(pn, ex_n))
                dev_c.append(O.rectified_rel_err(pc, ex_c).max(1))
            dev_a, dev_c = np.stack(dev_a), np.stack(dev_c)        # [1 + n_perm, rows]
            line = "%-16s seed %d:" % (name, seed)
            for prec in ("f16x3", "bf16x3"):
                ang = O.unoriented_angle_deg(d[key + prec + "/normals"], ref_n)
                err = O.rectified_rel_err(d[key + prec + "/curv"], ref_c).max(1)
                parts = []
                for n in (3, 8, 16, n_perm):
                    ga = np.maximum(0.01, 2 * dev_a[:1 + n].max(0))
                    gc = np.maximum(1e-3, 2 * dev_c[:1 + n].max(0))
                    parts.append("n=%d %.2f/%.2f" % (n, (ang / ga).max(), (err / gc).max()))
                line += "  %s angle/curv worst ratio: %s;" % (prec, ", ".join(parts))
            # the reference against itself: evaluation j (a permutation not used for the floor) vs the unpermuted forward
            parts = []
            for n in (3, 8, 16):
                ga = np.maximum(0.01, 2 * dev_a[:1 + n].max(0))
                gc = np.maximum(1e-3, 2 * dev_c[:1 + n].max(0))
                ra = max((O.unoriented_angle_deg(perm_n[j], ref_n) / ga).max() for j in range(n, n_perm))
                rc = max((O.rectified_rel_err(perm_c[j], ref_c).max(1) / gc).max() for j in range(n, n_perm))
                parts.append("n=%d %.2f/%.2f" % (n, ra, rc))
            line += "  reference (fresh permutations) vs itself: %s" % ", ".join(parts)
            print(line, flush=True)


if __name__ == "__main__":
    main()
