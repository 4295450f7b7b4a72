"""CPU: pin the oracle.  (a) against the reference itself, imported through oracle/shims.py, where
/root/reference exists (the authoring container); (b) everywhere, against the committed golden
vectors that the reference produced (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import CLOUD_NAMES
from oracle import deepfit_oracle as O
from oracle import shims

torch.set_grad_enabled(False)


@pytest.mark.parametrize("name", CLOUD_NAMES)
def test_oracle_knn_and_patches_match_golden(golden, clouds, name):
    pts, bb = O.normalize_cloud(clouds[name])
    assert bb == float(golden[name + "/bbdiag"])
    q = golden[name + "/query"][:12]
    idx, dist = O.knn_bruteforce(pts, q, 256)
    ref_idx, ref_dist = golden[name + "/knn_idx"][:12], golden[name + "/knn_dist"][:12]
    assert np.array_equal(dist, ref_dist)
    for r in range(len(q)):        # identical except inside a tie group cut by the k-th place
        bad = idx[r] != ref_idx[r]
        assert np.all(dist[r][bad] == dist[r][-1])
    for j in range(4):
        patch, U, rad = O.dataset_item(pts, int(q[j]), 256, ref_idx[j].astype(np.int64), ref_dist[j])
        assert rad == golden[name + "/rad"][j]
        # same LAPACK call as the reference; allow for a different CPU's SVD rounding/sign
        Uref = torch.from_numpy(golden[name + "/U"][j])
        s = torch.sign((U * Uref).sum(0))
        assert (U * s[None, :] - Uref).abs().max() < 1e-4
        assert (patch * s[:, None] - torch.from_numpy(golden[name + "/patch"][j])).abs().max() < 1e-4


@pytest.mark.parametrize("name", CLOUD_NAMES)
def test_oracle_fit_and_curvature_match_golden(golden, name):
    P = torch.from_numpy(golden[name + "/patch"][:32])
    for order in (1, 2, 3, 4):
        beta, n = O.fit_wjet(P, torch.ones(32, 256), order)
        ref_b = golden["%s/fit%d_beta" % (name, order)]
        assert np.abs(beta.numpy() - ref_b).max() < (1e-4 if order < 4 else 2e-2) * max(1.0, np.abs(ref_b).max())
        assert O.unoriented_angle_deg(n.numpy(), golden["%s/fit%d_n" % (name, order)]).max() < (1e-3 if order < 4 else 1e-2)
        if order > 1:
            ref_c = golden["%s/fit%d_curv" % (name, order)]
            c = O.principal_curvatures(torch.from_numpy(ref_b))
            assert O.rectified_rel_err(c.numpy(), ref_c).max() < 1e-5
            c2 = O.principal_curvatures_closed_form(torch.from_numpy(ref_b))
            assert O.rectified_rel_err(c2.numpy(), ref_c).max() < 1e-5
    beta, n = O.fit_wjet(P, torch.from_numpy(golden[name + "/wfit3_w"]), 3)
    assert O.unoriented_angle_deg(n.numpy(), golden[name + "/wfit3_n"]).max() < 1e-3


@pytest.mark.parametrize("name", CLOUD_NAMES)
def test_oracle_network_matches_golden(golden, name):
    sd = O.seeded_state_dict(0)
    assert len(sd) == 125
    P = torch.from_numpy(golden[name + "/patch"][:16])
    n, beta, w, trans, t2, _ = O.deepfit_forward(sd, P, 3)
    assert (w - torch.from_numpy(golden[name + "/net_weights"])).abs().max() < 1e-4
    assert (trans - torch.from_numpy(golden[name + "/net_trans"])).abs().max() < 1e-5
    assert (t2 - torch.from_numpy(golden[name + "/net_trans2"])).abs().max() < 1e-5
    nw = O.world_normals(n, torch.from_numpy(golden[name + "/U"][:16]), trans)
    assert O.unoriented_angle_deg(nw.numpy(), golden[name + "/net_normal_world"]).max() < 5e-3
    assert float(w.min()) < 0.1 and float(w.max()) > 0.9      # the seeded weights really spread


@pytest.mark.parametrize("name", CLOUD_NAMES)
def test_oracle_neighbor_normals_match_golden(golden, name):
    """fit_Wjet(..., compute_neighbor_normals=True): gradient at x/h, y/h with the un-preconditioned beta."""
    P = torch.from_numpy(golden[name + "/patch"][:16])
    for order in (1, 2, 3, 4):
        beta = torch.from_numpy(golden["%s/fit%d_beta" % (name, order)][:16])
        nn_ = O.neighbor_normals(P, beta, order)
        ref = torch.from_numpy(golden["%s/nn%d" % (name, order)])
        assert tuple(nn_.shape) == (16, 256, 3)
        assert (nn_ - ref).abs().max() < 2e-5


def test_quaternion_and_metrics():
    R = O.quat_to_rotmat(torch.tensor([[1.0, 0, 0, 0], [0.5, 0.5, 0.5, 0.5]]))
    assert torch.allclose(R[0], torch.eye(3))
    assert torch.allclose(R[1] @ R[1].t(), torch.eye(3), atol=1e-6)
    a = np.array([[0.0, 0, 1.0]])
    assert O.unoriented_angle_deg(a, -a)[0] == 0.0
    assert abs(O.unoriented_angle_deg(a, np.array([[0.0, 1.0, 1.0]]))[0] - 45.0) < 1e-9


@pytest.mark.skipif(not shims.reference_available(), reason="/root/reference is only present in the authoring container")
def test_oracle_equals_unmodified_reference():
    import os
    RD, RT = shims.import_reference()
    path = os.path.join(shims.REFERENCE_ROOT, "tutorial", "netsuke100k.xyz")   # a cloud NOT in the fixtures
    ds = RT.SinglePointCloudDataset(path, 256)
    raw = np.loadtxt(path).astype("float32")
    pts, bb = O.normalize_cloud(raw)
    assert np.array_equal(pts, ds.points) and bb == ds.bbdiag
    q = np.arange(7, 100000, 6250)[:16]
    items = [ds[int(i)] for i in q]
    idx, dist = O.knn_bruteforce(pts, q, 256)
    d_ref, i_ref = O.canonicalize_knn(*ds.kdtree.query(pts[q], k=256))
    assert np.array_equal(dist, d_ref) and np.array_equal(idx, i_ref)
    P = torch.stack([O.dataset_item(pts, int(c), 256, idx[j], dist[j])[0] for j, c in enumerate(q)])
    Pref = torch.stack([it[0] for it in items])
    assert torch.equal(P, Pref)
    for order in (1, 2, 3, 4):
        b_ref, n_ref, _ = RD.fit_Wjet(Pref, torch.ones_like(Pref[:, 0]), order=order)
        b, n = O.fit_wjet(Pref, torch.ones(16, 256), order)
        assert (b - b_ref).abs().max() < 1e-4 * max(1.0, float(b_ref.abs().max()))
        assert O.unoriented_angle_deg(n.numpy(), n_ref.numpy()).max() < 1e-4
        if order > 1:
            c_ref, _ = RT.compute_principal_curvatures(b_ref)
            assert torch.equal(O.principal_curvatures(b_ref), c_ref)
    sd = O.seeded_state_dict(3)
    m = RD.DeepFit(k=1, num_points=256, use_point_stn=1, use_feat_stn=True, point_tuple=1, sym_op="max", arch="simple",
                   n_gaussians=1, jet_order=3, weight_mode="sigmoid", use_consistency=False)
    m.load_state_dict(sd, strict=True)
    m.eval()
    n_ref, b_ref, w_ref, tr_ref, t2_ref, _ = m(Pref)
    n, b, w, tr, t2, _ = O.deepfit_forward(sd, Pref, 3)
    assert torch.equal(w, w_ref) and torch.equal(tr, tr_ref) and torch.equal(t2, t2_ref)
    assert O.unoriented_angle_deg(n.numpy(), n_ref.numpy()).max() < 1e-4


@pytest.mark.skipif(not shims.reference_available(), reason="/root/reference is only present in the authoring container")
def test_oracle_equals_reference_pcpnet_dataset(tmp_path):
    """dataset.py:268-417 (k-mode, use_pca, sparse .pidx centres, raw points) == oracle.dataset_item on the raw cloud."""
    import sys
    from deepfit_b200 import synthetic
    shims.apply_shims()
    root = str(tmp_path)
    names = synthetic.write_pcpnet_like(root, n_shapes=2, n_points=3000, n_sparse=40)
    sys.path.insert(0, shims.REFERENCE_ROOT)
    try:
        import importlib
        ref_dataset = importlib.import_module("dataset")
    finally:
        sys.path.remove(shims.REFERENCE_ROOT)
    ds = ref_dataset.PointcloudPatchDataset(root=root, shape_list_filename="testset_synthetic.txt", patch_radius=[0.05],
                                            points_per_patch=256, patch_features=[], seed=1, use_pca=True, center="point",
                                            point_tuple=1, cache_capacity=2, sparse_patches=True, neighbor_search_method="k")
    assert ds.shape_patch_count == [40, 40]
    for gi in (0, 7, 39, 40, 79):
        patch, trans, scale = ds[gi]
        si, pi = divmod(gi, 40)
        pts = np.loadtxt(tmp_path / (names[si] + ".xyz")).astype("float32")
        pidx = np.loadtxt(tmp_path / (names[si] + ".pidx")).astype("int64")
        p, U, rad = O.dataset_item(pts, int(pidx[pi]), 256)
        assert rad == float(scale)
        assert torch.equal(p.t(), patch) and torch.equal(U, trans)


@pytest.mark.skipif(not shims.reference_available(), reason="/root/reference is only present in the authoring container")
def test_oracle_equals_reference_pcpnet_patch_features(tmp_path):
    """dataset.py:338-372: ground-truth normal / neighbour normals / curvatures returned with a patch."""
    import sys
    from deepfit_b200 import synthetic
    shims.apply_shims()
    root = str(tmp_path)
    names = synthetic.write_pcpnet_like(root, n_shapes=1, n_points=3000, n_sparse=20)
    rng = np.random.default_rng(0)
    np.savetxt(tmp_path / (names[0] + ".curv"), rng.normal(size=(3000, 2)), fmt="%.6f")
    sys.path.insert(0, shims.REFERENCE_ROOT)
    try:
        import importlib
        ref_dataset = importlib.import_module("dataset")
    finally:
        sys.path.remove(shims.REFERENCE_ROOT)
    feats = ['normal', 'neighbor_normals', 'max_curvature', 'min_curvature']
    ds = ref_dataset.PointcloudPatchDataset(root=root, shape_list_filename="testset_synthetic.txt", patch_radius=[0.05],
                                            points_per_patch=256, patch_features=feats, seed=1, use_pca=True, center="point",
                                            point_tuple=1, cache_capacity=2, sparse_patches=True, neighbor_search_method="k")
    pts = np.loadtxt(tmp_path / (names[0] + ".xyz")).astype("float32")
    nrm = np.loadtxt(tmp_path / (names[0] + ".normals")).astype("float32")
    cur = np.loadtxt(tmp_path / (names[0] + ".curv")).astype("float32")
    pidx = np.loadtxt(tmp_path / (names[0] + ".pidx")).astype("int64")
    rad_abs = float(np.linalg.norm(pts.max(0) - pts.min(0), 2)) * 0.05
    for gi in (0, 5, 19):
        patch, normal, nn, cmax, cmin, trans, scale = ds[gi]
        idx, dist = O.knn_ckdtree(pts, [int(pidx[gi])], 256)
        f = O.pcpnet_patch_features(nrm, cur, int(pidx[gi]), idx[0], trans, rad_abs)
        assert torch.equal(f["normal"], normal) and torch.equal(f["max_curvature"], cmax) and torch.equal(f["min_curvature"], cmin)
        # the neighbour order of the tree query is arbitrary among ties only; compare as sets of rows
        assert torch.equal(torch.sort(f["neighbor_normals"], dim=0)[0], torch.sort(nn, dim=0)[0])
