"""BASELINE config 3: PCPNet-shaped multi-shape inference with full test_c_est.py outputs on 100 000-point
clouds at the three named noise levels, against the oracle on >= 512 patch centres per noise level.  Reference loop:
test_c_est.py:104-218 over dataset.py:268-417 (raw un-normalised points, k-nearest-neighbour patches, PCA)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_POINTS = 100000
N_SPARSE = 4096
N_PARITY = 512


@pytest.fixture(scope="module")
def dataset(tmp_path_factory):
    from deepfit_b200 import synthetic
    root = str(tmp_path_factory.mktemp("config3") / "pclouds")
    # three shapes = the three noise levels 0 / 0.0025 / 0.012 (sphere, torus, ellipsoid)
    names = synthetic.write_pcpnet_like(root, n_shapes=3, n_points=N_POINTS, n_sparse=N_SPARSE)
    return root, names


def test_config3_full_outputs_parity_at_size(dataset, tmp_path):
    from deepfit_b200 import pcpnet
    from deepfit_b200.pipeline import model_from_state_dict
    from oracle import deepfit_oracle as O
    root, names = dataset
    assert [n.split("_noise_")[1] for n in names] == ["0.0000", "0.0025", "0.0120"]
    sd = O.seeded_state_dict(0)
    sd64 = {k_: (v.double() if v.is_floating_point() else v) for k_, v in sd.items()}
    model = model_from_state_dict(sd, 256, 3, "sigmoid", "f16x3")
    ds = pcpnet.PointcloudPatchDataset(root, "testset_synthetic.txt", points_per_patch=256, use_pca=True,
                                       sparse_patches=True, neighbor_search_method='k')
    res = pcpnet.estimate_shapes(ds, model, str(tmp_path / "results"), batch=2048)
    raw_ds = pcpnet.PointcloudPatchDataset(root, "testset_synthetic.txt", points_per_patch=256, use_pca=False,
                                           sparse_patches=True, neighbor_search_method='k')
    for si, name in enumerate(names):
        for ext, cols in (("normals", 3), ("curv", 2), ("weights", 256), ("beta", 10), ("trans", 9)):
            arr = np.loadtxt(tmp_path / "results" / (name + "." + ext))
            assert arr.shape == (N_SPARSE, cols)
            assert np.array_equal(arr.astype(np.float32), res[name][ext])       # the files hold what was computed
        pts = np.loadtxt(root + "/" + name + ".xyz").astype("float32")           # raw, NOT normalised
        pidx = np.loadtxt(root + "/" + name + ".pidx").astype("int64")
        c = pidx[:N_PARITY]
        # stage 1: the reference's own search (cKDTree), canonical tie order
        ref_idx, ref_dist = O.knn_ckdtree(pts, c, 256, workers=-1)
        ref_dist, ref_idx = O.canonicalize_knn(ref_dist, ref_idx)
        patch, U, rad = ds.patches(si, 0, N_PARITY)
        assert np.array_equal(rad.cpu().numpy(), ref_dist[:, -1])
        # stage 2 on raw coordinates (dataset.py:292-334): bit-equal patches; PCA frames modulo column sign
        raw_patch, _, _ = raw_ds.patches(si, 0, N_PARITY)
        for j in range(0, N_PARITY, 8):
            ref_p = O.extract_patch(pts, int(c[j]), ref_idx[j], ref_dist[j, -1])
            assert torch.equal(raw_patch[j].cpu(), ref_p.t())
            _, ref_U = O.pca_align(ref_p)
            assert np.abs(np.abs((ref_U.t() @ U[j].cpu()).numpy()) - np.eye(3)).max() < 5e-3
        # stages 3-5 on the same patches: per-row gates anchored on the fp64 evaluation of the reference
        P = patch.cpu()
        with torch.no_grad():
            n, beta, w, trans, t2, _ = O.deepfit_forward(sd, P, 3)
            ref_n = O.world_normals(n, U.cpu(), trans).numpy()
            ref_c = (O.principal_curvatures(beta).double() / rad.cpu()[:, None]).numpy()
            n64, beta64, _, trans64, _, _ = O.deepfit_forward(sd64, P.double(), 3)
            ex_n = O.world_normals(n64, U.cpu().double(), trans64).numpy()
            ex_c = (O.principal_curvatures(beta64).double() / rad.cpu()[:, None]).numpy()
        gate_a = np.maximum(0.01, 2 * O.unoriented_angle_deg(ref_n, ex_n))
        gate_c = np.maximum(1e-3, 2 * O.rectified_rel_err(ref_c, ex_c).max(1))
        ang = O.unoriented_angle_deg(res[name]["normals"][:N_PARITY], ref_n)
        cerr = O.rectified_rel_err(res[name]["curv"][:N_PARITY], ref_c).max(1)
        dw = np.abs(res[name]["weights"][:N_PARITY] - w.numpy()).max()
        dt = np.abs(res[name]["trans"][:N_PARITY] - trans.reshape(-1, 9).numpy()).max()
        db = np.abs(res[name]["beta"][:N_PARITY] - beta.numpy()).max()
        print("%s: %d centres; angle max %.2e deg (rows with a relaxed gate: %d), worst angle/gate %.2f; curvature max %.2e "
              "(relaxed rows: %d), worst err/gate %.2f; |dw| %.1e |dtrans| %.1e |dbeta| %.1e"
              % (name, N_PARITY, ang.max(), int((gate_a > 0.01).sum()), (ang / gate_a).max(), cerr.max(),
                 int((gate_c > 1e-3).sum()), (cerr / gate_c).max(), dw, dt, db))
        assert (ang <= gate_a).all() and (cerr <= gate_c).all()
        assert dw < 5e-5 and dt < 1e-4


def test_config3_dense_pipeline_equals_stage_loop(tmp_path):
    """Dense shapes take the C pipeline (one dfb_pipeline_run per shape); it must give the bits of the stage-by-stage loop."""
    from deepfit_b200 import pcpnet, synthetic
    from deepfit_b200.pipeline import model_from_state_dict
    from oracle import deepfit_oracle as O
    root = str(tmp_path / "pclouds")
    names = synthetic.write_pcpnet_like(root, n_shapes=2, n_points=20000, n_sparse=100)
    model = model_from_state_dict(O.seeded_state_dict(0), 256, 3, "sigmoid", "f16x3")
    out = {}
    for use_pipeline in (True, False):
        ds = pcpnet.PointcloudPatchDataset(root, "testset_synthetic.txt", points_per_patch=256, use_pca=True,
                                           sparse_patches=False, neighbor_search_method='k')
        timings = {}
        out[use_pipeline] = pcpnet.estimate_shapes(ds, model, str(tmp_path / "results"), batch=4096, write=False,
                                                   timings=timings, use_pipeline=use_pipeline)
        assert set(timings) == {"load", "compute", "export"}
    for name in names:
        for key in ("normals", "curv", "weights", "beta", "trans"):
            assert out[True][name][key].shape[0] == 20000
            assert np.array_equal(out[True][name][key], out[False][name][key]), (name, key)
    # classic mode (no network): weights of ones, identity trans
    ds = pcpnet.PointcloudPatchDataset(root, "testset_synthetic.txt", points_per_patch=256, use_pca=True,
                                       sparse_patches=False, neighbor_search_method='k')
    a = pcpnet.estimate_shapes(ds, None, str(tmp_path / "results"), jet_order=2, write=False)
    b = pcpnet.estimate_shapes(ds, None, str(tmp_path / "results"), jet_order=2, write=False, use_pipeline=False)
    for name in names:
        for key in ("normals", "curv", "weights", "beta", "trans"):
            assert np.array_equal(a[name][key], b[name][key]), (name, key)


@pytest.mark.parametrize("use_pca", [True, False])
def test_pcpnet_items_with_ground_truth_features(tmp_path, use_pca):
    """items() = the reference's batched __getitem__ (dataset.py:268-417): ground-truth normal, neighbour normals and
    curvatures next to the patch, rotated into the patch's PCA frame, against the oracle restatement (which is pinned
    on the unmodified reference by tests/test_oracle.py::test_oracle_equals_reference_pcpnet_patch_features)."""
    from deepfit_b200 import pcpnet, synthetic
    from oracle import deepfit_oracle as O
    root = str(tmp_path / "pclouds")
    names = synthetic.write_pcpnet_like(root, n_shapes=2, n_points=5000, n_sparse=64)
    rng = np.random.default_rng(0)
    for name in names:
        np.savetxt(root + "/" + name + ".curv", rng.normal(size=(5000, 2)), fmt="%.6f")
    feats = ['normal', 'max_curvature', 'neighbor_normals', 'min_curvature']
    ds = pcpnet.PointcloudPatchDataset(root, "testset_synthetic.txt", patch_radius=[0.05], points_per_patch=256,
                                       patch_features=feats, use_pca=use_pca, sparse_patches=True, neighbor_search_method='k')
    with pytest.raises(ValueError):
        pcpnet.PointcloudPatchDataset(root, "testset_synthetic.txt", patch_features=['colour'])
    for si, name in enumerate(names):
        pts = np.loadtxt(root + "/" + name + ".xyz").astype("float32")
        nrm = np.loadtxt(root + "/" + name + ".normals").astype("float32")
        cur = np.loadtxt(root + "/" + name + ".curv").astype("float32")
        pidx = np.loadtxt(root + "/" + name + ".pidx").astype("int64")
        rad_abs = float(np.linalg.norm(pts.max(0) - pts.min(0), 2)) * 0.05
        patch, normal, cmax, nn, cmin, trans, scale = ds.items(si, 0, 64)
        assert patch.shape == (64, 256, 3) and nn.shape == (64, 256, 3) and cmax.shape == (64, 1) and trans.shape == (64, 3, 3)
        p2, t2, s2, idx, valid = ds.patches(si, 0, 64, return_index=True)
        assert torch.equal(patch, p2.transpose(1, 2)) and torch.equal(trans, t2) and bool((valid == 256).all())
        if not use_pca:
            assert torch.equal(trans, torch.eye(3, device="cuda").expand(64, 3, 3))
        ref_idx, ref_dist = O.knn_bruteforce(pts, pidx[:64], 256)
        assert np.array_equal(idx.cpu().numpy(), ref_idx.astype(np.int32))
        for j in range(64):
            f = O.pcpnet_patch_features(nrm, cur, int(pidx[j]), ref_idx[j], trans[j].cpu() if use_pca else None, rad_abs)
            assert torch.allclose(normal[j].cpu(), f["normal"], atol=1e-6, rtol=0)
            assert torch.allclose(nn[j].cpu(), f["neighbor_normals"], atol=1e-6, rtol=0)
            assert torch.equal(cmax[j].cpu(), f["max_curvature"]) and torch.equal(cmin[j].cpu(), f["min_curvature"])
