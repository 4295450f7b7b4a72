"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):

    python oracle/make_golden.py

Everything written here comes from the reference's own code (through oracle/shims.py):
``tutorial_utils.SinglePointCloudDataset`` (cKDTree + torch.svd), ``DeepFit.fit_Wjet``,
``tutorial_utils.compute_principal_curvatures`` and ``DeepFit.DeepFit.forward`` (eval mode, seeded
state_dict from ``oracle.deepfit_oracle.seeded_state_dict`` because DeepFit.pth is missing).  The
two tutorial clouds are stored as float32 arrays (what ``np.loadtxt(..).astype('float32')`` yields,
tutorial_utils.py:13) because /root/reference does not exist on the GPU box.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import deepfit_oracle as O  # noqa: E402
from oracle.shims import REFERENCE_ROOT, import_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
K = 256
N_QUERY = 64
N_MODEL = 16          # rows with every network output stored (incl. the 64 x 64 feature transform)
N_GATE = 512          # rows of the per-row parity gate: patches, reference outputs, noise floor
N_ENSEMBLE = 8        # neighbour permutations of the reference forward that estimate its own fp32 noise per row


def main():
    RD, RT = import_reference()
    torch.set_grad_enabled(False)
    os.makedirs(GOLDEN, exist_ok=True)
    clouds = {}
    out = {}
    for name in ("Boxy_smooth100k", "0000000000"):
        path = os.path.join(REFERENCE_ROOT, "tutorial", name + ".xyz")
        raw = np.loadtxt(path).astype("float32")
        clouds[name] = raw
        ds = RT.SinglePointCloudDataset(path, K)
        rng = np.random.default_rng(1234)
        q = np.sort(rng.choice(len(ds), N_QUERY, replace=False)).astype(np.int64)
        dist, idx = ds.kdtree.query(ds.points[q], k=K)
        dist, idx = O.canonicalize_knn(dist, idx)
        items = [ds[int(i)] for i in q]
        patch = torch.stack([it[0] for it in items])
        U = torch.stack([it[1] for it in items])
        rad = np.array([float(it[2]) for it in items], dtype=np.float64)
        key = name
        out[key + "/query"] = q
        out[key + "/knn_idx"] = idx.astype(np.int32)
        out[key + "/knn_dist"] = dist
        out[key + "/patch"] = patch.numpy()
        out[key + "/U"] = U.numpy()
        out[key + "/rad"] = rad
        out[key + "/bbdiag"] = np.float64(ds.bbdiag)
        # also k=512 radius for a handful of queries (config 4 shape)
        d512, i512 = ds.kdtree.query(ds.points[q[:8]], k=512)
        d512, i512 = O.canonicalize_knn(d512, i512)
        out[key + "/knn512_idx"] = i512.astype(np.int32)
        out[key + "/knn512_dist"] = d512
        # classic jet fit, orders 1..4, on 32 patches (a multiple of the reference's sub-batch of 16)
        P32 = patch[:32]
        for order in (1, 2, 3, 4):
            beta, n_est, _ = RD.fit_Wjet(P32, torch.ones_like(P32[:, 0]), order=order)
            out["%s/fit%d_beta" % (key, order)] = beta.numpy()
            out["%s/fit%d_n" % (key, order)] = n_est.numpy()
            if order > 1:
                curv, dirs = RT.compute_principal_curvatures(beta)
                out["%s/fit%d_curv" % (key, order)] = curv.numpy()
        # per-neighbour normals (compute_neighbor_normals=True, DeepFit.py:90-115) on 16 patches (a whole sub-batch:
        # fewer would fall into the reference's B mod 16 hole and come back as beta = 0)
        for order in (1, 2, 3, 4):
            b8, _, nn8 = RD.fit_Wjet(P32[:16], torch.ones_like(P32[:16, 0]), order=order, compute_neighbor_normals=True)
            out["%s/nn%d" % (key, order)] = nn8.numpy()
        # weighted fit with explicit random weights
        g = torch.Generator().manual_seed(7)
        wts = torch.rand(P32.shape[0], K, generator=g) * 0.99 + 0.01
        beta, n_est, _ = RD.fit_Wjet(P32, wts, order=3)
        out[key + "/wfit3_w"] = wts.numpy()
        out[key + "/wfit3_beta"] = beta.numpy()
        out[key + "/wfit3_n"] = n_est.numpy()
        # network: the reference module in eval mode on the reference's own patches
        sd = O.seeded_state_dict(0)
        model = RD.DeepFit(k=1, num_points=K, use_point_stn=1, use_feat_stn=True, point_tuple=1, sym_op="max",
                           arch="simple", n_gaussians=1, jet_order=3, weight_mode="sigmoid", use_consistency=False)
        model.load_state_dict(sd, strict=True)
        model.eval()
        PM = patch[:N_MODEL]
        normal, beta, weights, trans, trans2, _ = model(PM)
        out[key + "/net_normal"] = normal.numpy()
        out[key + "/net_beta"] = beta.numpy()
        out[key + "/net_weights"] = weights.numpy()
        out[key + "/net_trans"] = trans.numpy()
        out[key + "/net_trans2"] = trans2.numpy()
        # end-to-end world normals / curvatures the way test_n_est.py:154-161 post-processes them
        nw = torch.bmm(normal.unsqueeze(1), trans.transpose(2, 1)).squeeze(1)
        nw = torch.bmm(nw.unsqueeze(1), U[:N_MODEL].transpose(2, 1)).squeeze(1)
        out[key + "/net_normal_world"] = nw.numpy()
        curv, dirs = RT.compute_principal_curvatures(beta)
        out[key + "/net_curv_scaled"] = curv.numpy().astype(np.float64) / rad[:N_MODEL, None]
        # principal directions (tutorial_utils.py:223-224) and their world-space form (test_c_est.py:162-168)
        out[key + "/net_dirs"] = dirs.numpy()
        dw = torch.matmul(dirs, trans.transpose(2, 1))
        dw = torch.matmul(dw, U[:N_MODEL].transpose(2, 1))
        out[key + "/net_dirs_world"] = dw.numpy()
        cfit, dfit = RT.compute_principal_curvatures(torch.from_numpy(out[key + "/fit3_beta"]))
        out[key + "/fit3_dirs"] = dfit.numpy()

        # ---- per-row parity gate set: N_GATE random patches of this cloud through the reference, plus the reference's
        # own noise floor on each row: the deviation from the fp64-exact result (oracle evaluated in float64) of
        # N_ENSEMBLE + 1 mathematically equivalent fp32 evaluations of the reference (its forward on the patch and on
        # N_ENSEMBLE neighbour permutations of it -- the network max-pools over the points and the jet fit sums over
        # them, so only the fp32 summation order changes)
        rng2 = np.random.default_rng(4321)
        qg = np.sort(rng2.choice(len(ds), N_GATE, replace=False)).astype(np.int64)
        itg = [ds[int(i)] for i in qg]
        Pg = torch.stack([it[0] for it in itg])
        Ug = torch.stack([it[1] for it in itg])
        radg = np.array([float(it[2]) for it in itg], dtype=np.float64)

        def ref_forward(P_):
            outs = [model(P_[s:s + 64]) for s in range(0, P_.shape[0], 64)]
            n_ = torch.cat([o[0] for o in outs]); b_ = torch.cat([o[1] for o in outs])
            w_ = torch.cat([o[2] for o in outs]); t_ = torch.cat([o[3] for o in outs])
            nw_ = torch.bmm(n_.unsqueeze(1), t_.transpose(2, 1)).squeeze(1)
            nw_ = torch.bmm(nw_.unsqueeze(1), Ug.transpose(2, 1)).squeeze(1)
            c_, _ = RT.compute_principal_curvatures(b_)
            return nw_.numpy(), c_.numpy().astype(np.float64) / radg[:, None], w_, t_, b_
        nw_ref, c_ref, w_ref, t_ref, b_ref = ref_forward(Pg)
        sd64 = {k_: (v.double() if v.is_floating_point() else v) for k_, v in sd.items()}
        w64, R64, _, pts64 = O.weight_network(sd64, Pg.double())
        beta64, n64 = O.fit_wjet(pts64, w64, 3)
        nw64 = O.world_normals(n64, Ug.double(), R64).numpy()
        c64 = O.principal_curvatures(beta64).numpy() / radg[:, None]
        floor_a = O.unoriented_angle_deg(nw_ref, nw64)
        floor_c = O.rectified_rel_err(c_ref, c64).max(1)
        gperm = torch.Generator().manual_seed(99)
        for j in range(N_ENSEMBLE):
            perm = torch.randperm(K, generator=gperm)
            nw_p, c_p, _, _, _ = ref_forward(Pg[:, :, perm].contiguous())
            floor_a = np.maximum(floor_a, O.unoriented_angle_deg(nw_p, nw64))
            floor_c = np.maximum(floor_c, O.rectified_rel_err(c_p, c64).max(1))
        out[key + "/gate_query"] = qg
        out[key + "/gate_patch"] = Pg.numpy()
        out[key + "/gate_U"] = Ug.numpy()
        out[key + "/gate_rad"] = radg
        out[key + "/gate_weights"] = w_ref.numpy()
        out[key + "/gate_trans"] = t_ref.numpy()
        out[key + "/gate_normal_world"] = nw_ref
        out[key + "/gate_curv_scaled"] = c_ref
        out[key + "/gate_floor_angle_deg"] = floor_a
        out[key + "/gate_floor_curv"] = floor_c
        print(name, "gate rows: floor angle max %.3e p99 %.3e, rows above 0.005 deg: %d; floor curv max %.3e, rows above 5e-4: %d"
              % (floor_a.max(), np.quantile(floor_a, 0.99), int((floor_a > 0.005).sum()), floor_c.max(), int((floor_c > 5e-4).sum())))
    np.savez_compressed(os.path.join(GOLDEN, "reference_outputs.npz"), **out)
    np.savez_compressed(os.path.join(GOLDEN, "tutorial_clouds.npz"), **clouds)
    for f in ("reference_outputs.npz", "tutorial_clouds.npz"):
        print(f, os.path.getsize(os.path.join(GOLDEN, f)) / 1e6, "MB")


if __name__ == "__main__":
    main()
