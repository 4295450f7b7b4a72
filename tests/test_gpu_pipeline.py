"""End-to-end (gpu): the whole hot path through the drop-in API against the oracle, sharding
equivalence, and the drop-in CLI's output files."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_classic_pipeline_vs_oracle_on_tutorial_cloud(clouds, tmp_path):
    from deepfit_b200 import tutorial_utils as tu
    from deepfit_b200.pipeline import NormalEstimator
    from oracle import deepfit_oracle as O
    raw = clouds["Boxy_smooth100k"]
    ds = tu.SinglePointCloudDataset(raw, 256)
    ref_pts, ref_bb = O.normalize_cloud(raw)
    assert np.array_equal(ds.points, ref_pts) and ds.bbdiag == ref_bb
    est = NormalEstimator(ds.points_gpu, 256, jet_order=3, mode="classic")
    normals, curv = est.run()
    torch.cuda.synchronize()
    assert tuple(normals.shape) == (100000, 3) and tuple(curv.shape) == (100000, 2) and curv.dtype == torch.float64
    rng = np.random.default_rng(2)
    q = np.sort(rng.choice(100000, 512, replace=False))
    ref = O.compute_normals(ref_pts, q, 256, 3, mode="classic")
    ang = O.unoriented_angle_deg(normals.cpu().numpy()[q], ref["normals"])
    assert ang.max() <= 0.01, "max angular deviation %g deg" % ang.max()
    # the PCA z-axis sign is library-defined in the reference: (k1,k2) -> (-k2,-k1) when it flips
    c = curv.cpu().numpy()[q]
    flipped = np.stack([-c[:, 1], -c[:, 0]], 1)
    e1 = O.rectified_rel_err(c, ref["curv"]).max(1)
    e2 = O.rectified_rel_err(flipped, ref["curv"]).max(1)
    assert np.minimum(e1, e2).max() <= 1e-3, "max curvature error %g" % np.minimum(e1, e2).max()
    # sharding: two half ranges computed separately are bit-identical to the single run
    n1, c1 = est.run(0, 50000)
    n2, c2 = est.run(50000, 100000)
    assert torch.equal(torch.cat([n1, n2]), normals) and torch.equal(torch.cat([c1, c2]), curv)
    # dataset API contract (tutorial_utils.py:19-30)
    patch, trans, rad = ds[int(q[0])]
    assert tuple(patch.shape) == (3, 256) and tuple(trans.shape) == (3, 3) and isinstance(rad, float)
    assert rad == ref["rad"][0] and len(ds) == 100000


_ORACLE_CACHE = {}


def _oracle_on_our_patches(name, sd, patch, U, rad, n_perm=3, n_fresh=0):
    """Reference algebra (oracle, fp32) on OUR patches + its per-row fp32 noise floor: the largest deviation from the
    fp64-exact result among the fp32 forward and n_perm neighbour permutations of it (same rule as the golden gate set).
    n_fresh > 0 adds a fifth element: the reference measured against ITSELF under those gates -- the worst (angle, curvature)
    error / gate of n_fresh further permutations (not used for the floors) against the unpermuted forward."""
    from oracle import deepfit_oracle as O
    if name in _ORACLE_CACHE:
        return _ORACLE_CACHE[name]
    with torch.no_grad():
        P, Uc, radc = patch.cpu(), U.cpu(), rad.cpu()

        def fwd(Pp, s=sd, dt=torch.float32):
            n, beta, w, trans, t2, _ = O.deepfit_forward(s, Pp.to(dt), 3)
            return (O.world_normals(n, Uc.to(dt), trans).numpy(),
                    (O.principal_curvatures(beta).double() / radc[:, None]).numpy())
        ref_n, ref_c = fwd(P)
        sd64 = {k_: (v.double() if v.is_floating_point() else v) for k_, v in sd.items()}
        ex_n, ex_c = fwd(P, sd64, torch.float64)
        fa = O.unoriented_angle_deg(ref_n, ex_n)
        fc = O.rectified_rel_err(ref_c, ex_c).max(1)
        g = torch.Generator().manual_seed(5)
        for _ in range(n_perm):
            pn, pc = fwd(P[:, :, torch.randperm(P.shape[2], generator=g)].contiguous())
            fa = np.maximum(fa, O.unoriented_angle_deg(pn, ex_n))
            fc = np.maximum(fc, O.rectified_rel_err(pc, ex_c).max(1))
        out = (ref_n, ref_c, fa, fc)
        if n_fresh:
            gate_a, gate_c = np.maximum(0.01, 2 * fa), np.maximum(1e-3, 2 * fc)
            ra = rc = 0.0
            for _ in range(n_fresh):
                pn, pc = fwd(P[:, :, torch.randperm(P.shape[2], generator=g)].contiguous())
                ra = max(ra, float((O.unoriented_angle_deg(pn, ref_n) / gate_a).max()))
                rc = max(rc, float((O.rectified_rel_err(pc, ref_c).max(1) / gate_c).max()))
            out = out + ((ra, rc),)
    _ORACLE_CACHE[name] = out
    return _ORACLE_CACHE[name]


@pytest.mark.parametrize("name,precision", [(n, p) for n in ("Boxy_smooth100k", "0000000000") for p in ("f16x3", "bf16x3", "f16f8")]
                         + [("Boxy_smooth100k", "f16mix")])
def test_deepfit_pipeline_vs_oracle(clouds, name, precision):
    """DeepFit mode end to end on both tutorial clouds (smooth CAD surface and noisy lidar scan).  The network is fed
    OUR PCA frame, so the oracle runs the reference algebra on our patches (the sign-alignment protocol of SURVEY.md
    section 8c).  Gates per row: max(0.01 deg, 2 x the reference's own fp32 noise on that row) and the same for 1e-3
    on the curvatures; no row of the smooth cloud is relaxed.  f16x3 is the parity mode (every gate on every row);
    the approximate modes bf16x3 / f16f8 get 2x / 4x the curvature gate on the lidar cloud (see test_gpu_network.py);
    f16mix (single-product max-pool layers in the two STN chains) holds every gate of the smooth cloud and is not
    claimed on the lidar one (3 deg off on its worst row of this range)."""
    from deepfit_b200 import ops
    from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
    from oracle import deepfit_oracle as O
    raw = clouds[name]
    pts_np = O.normalize_cloud(raw)[0]
    pts = torch.from_numpy(pts_np).cuda()
    sd = O.seeded_state_dict(0)
    model = model_from_state_dict(sd, 256, 3, "sigmoid", precision)
    est = NormalEstimator(pts, 256, mode="DeepFit", model=model, chunk=256)
    begin, count = 1000, 640                      # 2.5 chunks: ragged tail, both halves of the double buffer
    normals, curv = est.run(begin, begin + count)
    idx, rad = est.grid.query(256, q_begin=begin, q_count=count)
    patch, U = ops.patch_pca(pts, idx, rad, q_begin=begin)
    ref_n, ref_c, fa, fc = _oracle_on_our_patches(name, sd, patch, U, rad)
    ang = O.unoriented_angle_deg(normals.cpu().numpy(), ref_n)
    err = O.rectified_rel_err(curv.cpu().numpy(), ref_c).max(1)
    gate_a, gate_c = np.maximum(0.01, 2 * fa), np.maximum(1e-3, 2 * fc)
    print("%s %s: angle max %.3e deg (worst angle/gate %.2f, %d of %d rows relaxed); curvature max %.3e (worst err/gate %.2f, "
          "%d rows relaxed)" % (name, precision, ang.max(), (ang / gate_a).max(), int((gate_a > 0.01).sum()), count, err.max(),
                                (err / gate_c).max(), int((gate_c > 1e-3).sum())))
    assert (ang <= gate_a).all(), "max angular deviation %g deg (p99 %g)" % (ang.max(), np.quantile(ang, 0.99))
    cf = {"f16x3": 1.0, "bf16x3": 2.0, "f16f8": 4.0}[precision] if name == "0000000000" else 1.0
    assert (err <= cf * gate_c).all(), "max curvature error / gate %g" % (err / gate_c).max()
    if name == "Boxy_smooth100k":
        assert (gate_a <= 0.01).all() and (gate_c <= 1e-3).all()


@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("name", ["Boxy_smooth100k", "0000000000"])
def test_deepfit_pipeline_other_weight_sets(clouds, name, seed):
    """The pretrained blob is missing from the reference checkout, so every other parity test rests on ONE seeded
    state_dict (seed 0, the one the golden vectors were generated with).  Same end-to-end check as
    test_deepfit_pipeline_vs_oracle in the parity mode under two more weight sets, the oracle (reference algebra,
    fp32) and the per-row fp32 noise floors evaluated on the fly.

    Every row must hold its normal gate.  For the curvature the statement is: every row holds its gate, OR the reference
    itself does not -- on the lidar cloud two fp32 evaluations of the REFERENCE that differ only in the order of the
    neighbours disagree by 1.1-2.7x the per-row gate on their worst row (floors from 3 permutations; 0.8-1.1x with 16:
    oracle/studies/floor_estimate.py, profiles/r2_floor_estimate_seeds.txt), so a row of ours may exceed its gate by no
    more than the reference exceeds it against itself, measured here on 3 fresh permutations.  Measured: smooth cloud
    0.04 / 0.07 of the gate; lidar cloud 0.44 (seed 1) and 1.00 (seed 2: one strict row of 384 whose network weights are
    below 0.02 for 133 of its 256 neighbours; 1.28 before the FC layers' corrections were kept out of the main
    accumulator, csrc/mlp.cu gemm_kernel), the reference against itself 1.3-2.7 there."""
    from deepfit_b200 import ops
    from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
    from oracle import deepfit_oracle as O
    pts = torch.from_numpy(O.normalize_cloud(clouds[name])[0]).cuda()
    sd = O.seeded_state_dict(seed)
    model = model_from_state_dict(sd, 256, 3, "sigmoid", "f16x3")
    est = NormalEstimator(pts, 256, mode="DeepFit", model=model, chunk=256)
    begin, count = 40000 + 1000 * seed, 384
    normals, curv = est.run(begin, begin + count)
    idx, rad = est.grid.query(256, q_begin=begin, q_count=count)
    patch, U = ops.patch_pca(pts, idx, rad, q_begin=begin)
    ref_n, ref_c, fa, fc, (self_a, self_c) = _oracle_on_our_patches("%s/seed%d" % (name, seed), sd, patch, U, rad, n_fresh=3)
    ang = O.unoriented_angle_deg(normals.cpu().numpy(), ref_n)
    err = O.rectified_rel_err(curv.cpu().numpy(), ref_c).max(1)
    gate_a, gate_c = np.maximum(0.01, 2 * fa), np.maximum(1e-3, 2 * fc)
    print("%s seed %d: angle max %.3e deg (worst angle/gate %.2f, %d of %d rows relaxed); curvature max %.3e (worst err/gate "
          "%.2f, %d rows relaxed, %d over); the reference against itself: %.2f / %.2f"
          % (name, seed, ang.max(), (ang / gate_a).max(), int((gate_a > 0.01).sum()), count, err.max(), (err / gate_c).max(),
             int((gate_c > 1e-3).sum()), int((err > gate_c).sum()), self_a, self_c))
    assert (ang <= gate_a).all(), "max angular deviation %g deg, worst angle/gate %g" % (ang.max(), (ang / gate_a).max())
    assert (err <= max(1.0, self_c) * gate_c).all() and (err > gate_c).sum() <= 1, \
        "max curvature error / gate %g (the reference against itself: %g)" % ((err / gate_c).max(), self_c)
    if name == "Boxy_smooth100k":
        assert (gate_a <= 0.01).all() and (gate_c <= 1e-3).all() and (err <= gate_c).all()


@pytest.mark.parametrize("file_order", [False, True], ids=["morton_order", "file_order"])
@pytest.mark.parametrize("mode", ["classic", "DeepFit"])
def test_pipeline_call_equals_stage_calls(clouds, mode, file_order):
    """dfb_pipeline_run (chunk loop + two-stream schedule inside the library) is bit-identical to calling the five
    stages one by one in file order, for every optional output, over ragged chunking -- whether the pipeline visits the
    queries in the grid's Morton order (default; results scattered back to their file-order rows) or in file order;
    DFB_PIPE_KEEP_QSTN reproduces the tutorial script's omission (compute_normals.py:67)."""
    from deepfit_b200 import ops
    from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
    from oracle import deepfit_oracle as O
    pts = torch.from_numpy(O.normalize_cloud(clouds["0000000000"])[0]).cuda()
    k, order = 256, 3
    model = model_from_state_dict(O.seeded_state_dict(1), k, order, "sigmoid", "f16x3") if mode == "DeepFit" else None
    est = NormalEstimator(pts, k, order, mode, model, chunk=300, file_order=file_order)
    begin, end = 777, 777 + 1000                  # 3 chunks of 300 + a tail of 100
    extras = {}
    normals, curv = est.run(begin, end, extras=extras)
    torch.cuda.synchronize()
    idx, rad = est.grid.query(k, q_begin=begin, q_count=end - begin)
    patch, U = ops.patch_pca(pts, idx, rad, q_begin=begin)
    if mode == "DeepFit":
        w, trans, t2, _ = est.net.forward(patch, want_trans2=True)
        out = ops.wls_fit(patch, w, order, trans=trans, U=U, rad=rad, want_dirs=True)
        assert torch.equal(extras["weights"], w) and torch.equal(extras["trans"], trans) and torch.equal(extras["trans2"], t2)
    else:
        out = ops.wls_fit(patch, None, order, U=U, rad=rad, want_dirs=True)
        assert "weights" not in extras
    assert torch.equal(normals, out["n_world"]) and torch.equal(curv, out["curv_scaled"])
    assert torch.equal(extras["beta"], out["beta"]) and torch.equal(extras["info"], out["info"])
    assert torch.equal(extras["U"], U) and torch.equal(extras["rad"], rad)
    assert torch.equal(extras["dirs_world"], out["dirs_world"])
    # a second run re-uses the workspace and the events
    n2, c2 = est.run(begin, end)
    assert torch.equal(n2, normals) and torch.equal(c2, curv)
    if mode == "DeepFit":
        est2 = NormalEstimator(pts, k, order, mode, model, chunk=300, cancel_qstn=False, grid=est.grid)
        n3, _ = est2.run(begin, end)
        expect = torch.bmm(out["n_local"].unsqueeze(1), U.transpose(2, 1)).squeeze(1)
        assert (n3 - expect).abs().max().item() < 1e-6


def test_cli_writes_reference_format(clouds, tmp_path):
    from deepfit_b200 import compute_normals
    raw = clouds["Boxy_smooth100k"][:20000]
    inp = tmp_path / "mini.xyz"
    np.savetxt(inp, raw, fmt="%.6f")
    compute_normals.main(["--input", str(inp), "--output_path", str(tmp_path / "out"), "--mode", "classic",
                          "--k_neighbors", "64", "--jet_order", "2"])
    nrm = np.loadtxt(tmp_path / "out" / "mini.normals")
    crv = np.loadtxt(tmp_path / "out" / "mini.curv")
    assert nrm.shape == (20000, 3) and crv.shape == (20000, 2)
    assert np.allclose(np.linalg.norm(nrm, axis=1), 1.0, atol=1e-5)
    assert (crv[:, 0] <= crv[:, 1]).all()
    with pytest.raises(ValueError):   # the reference dies the same way for order 1 (tutorial_utils.py:204-205)
        compute_normals.main(["--input", str(inp), "--output_path", str(tmp_path / "o1"), "--mode", "classic",
                              "--k_neighbors", "64", "--jet_order", "1"])


def test_deepfit_order4_k512_stress_shape(clouds):
    """BASELINE config 4 shape (jet order 4, k=512) through the whole path vs the oracle on our patches."""
    from deepfit_b200 import ops
    from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
    from oracle import deepfit_oracle as O
    pts_np = O.normalize_cloud(clouds["Boxy_smooth100k"])[0]
    pts = torch.from_numpy(pts_np).cuda()
    sd = O.seeded_state_dict(2)
    model = model_from_state_dict(sd, 512, 4, "sigmoid", "f16x3")
    est = NormalEstimator(pts, 512, mode="DeepFit", model=model, chunk=128)
    begin, count = 40000, 192
    normals, curv = est.run(begin, begin + count)
    est.net.status()
    idx, rad, dist = est.grid.query(512, q_begin=begin, q_count=count, return_dist=True)
    ref_idx, ref_dist = O.knn_bruteforce(pts_np, np.arange(begin, begin + 32), 512)
    assert np.array_equal(idx[:32].cpu().numpy(), ref_idx.astype(np.int32))
    assert np.array_equal(dist[:32].cpu().numpy(), ref_dist)
    patch, U = ops.patch_pca(pts, idx, rad, q_begin=begin)
    with torch.no_grad():
        n, beta, w, trans, t2, _ = O.deepfit_forward(sd, patch.cpu(), 4)
        ref_n = O.world_normals(n, U.cpu(), trans).numpy()
        # the exact (fp64) solution of the same weighted system bounds the reference's own fp32 error
        w64, tr64 = w.double(), trans.double()
        prot = torch.bmm(patch.cpu().double().transpose(1, 2), tr64).transpose(1, 2).contiguous()
        _, n64 = O.fit_wjet(prot, w64, 4)
        ex_n = O.world_normals(n64, U.cpu().double(), tr64).numpy()
    ref_err = O.unoriented_angle_deg(ref_n, ex_n)
    ang = O.unoriented_angle_deg(normals.cpu().numpy(), ref_n)
    assert (ang <= np.maximum(0.01, 2 * ref_err)).all(), "max angular deviation %g deg" % ang.max()
    assert tuple(curv.shape) == (count, 2) and bool(torch.isfinite(curv).all())


def test_classic_orders_on_synthetic_sphere_and_torus():
    """BASELINE config 2 shape at a reduced size: classic jet orders 1-4 on noisy analytic surfaces, every
    point of the cloud; normals against the analytic ones, and bit-identical across chunk sizes."""
    from deepfit_b200 import synthetic
    from deepfit_b200.pipeline import NormalEstimator
    raw = synthetic.torus_cloud(60000, seed=1234, noise=0.0)
    pts = torch.from_numpy(raw).cuda()
    u = np.arctan2(raw[:, 1], raw[:, 0])
    centre = np.stack([np.cos(u), np.sin(u), np.zeros_like(u)], 1)
    true_n = raw - centre
    true_n /= np.linalg.norm(true_n, axis=1, keepdims=True)
    for order in (1, 2, 3, 4):
        est = NormalEstimator(pts, 64, jet_order=order, mode="classic", chunk=16384)
        normals, curv = est.run()
        cosang = np.abs((normals.cpu().numpy() * true_n).sum(1))
        assert np.degrees(np.arccos(np.clip(cosang, 0, 1))).max() < (6.0 if order == 1 else 3.0)
        assert (curv is None) == (order == 1)
        est2 = NormalEstimator(pts, 64, jet_order=order, mode="classic", chunk=5000)
        n2, c2 = est2.run()
        assert torch.equal(normals, n2) and (curv is None or torch.equal(curv, c2))


@pytest.mark.parametrize("kind", ["torus", "sphere"])
def test_config2_classic_orders_at_named_size(kind):
    """BASELINE config 2 at its named size: classic (unweighted) jet orders 1-4, k=256, on the 1 M-point noisy
    torus / sphere (seed 1234, sigma 0.005 bbdiag) through the C pipeline; 512 random centres against the oracle
    (cKDTree neighbours, reference fit) with the per-row gate max(gate, 2 x |reference fp32 - fp64-exact|)."""
    from deepfit_b200 import synthetic
    from deepfit_b200.pipeline import NormalEstimator
    from deepfit_b200.tutorial_utils import normalize_cloud
    from oracle import deepfit_oracle as O
    gen = synthetic.torus_cloud if kind == "torus" else synthetic.sphere_cloud
    pts = np.ascontiguousarray(normalize_cloud(gen(1_000_000, seed=1234, noise=0.005))[0], dtype=np.float32)
    dev = torch.from_numpy(pts).cuda()
    q = np.sort(np.random.default_rng(7).choice(len(pts), 512, replace=False))
    idx, dist = O.knn_ckdtree(pts, q, 256, workers=-1)
    dist, idx = O.canonicalize_knn(dist, idx)
    items = [O.dataset_item(pts, int(c), 256, idx[j], dist[j]) for j, c in enumerate(q)]
    P = torch.stack([it[0] for it in items])
    U = torch.stack([it[1] for it in items])
    rad = torch.tensor([it[2] for it in items], dtype=torch.float64)
    for order in (1, 2, 3, 4):
        est = NormalEstimator(dev, 256, jet_order=order, mode="classic")
        normals, curv = est.run()
        torch.cuda.synchronize()
        assert bool(torch.isfinite(normals).all()) and (curv is None) == (order == 1)
        with torch.no_grad():
            beta, n = O.fit_wjet(P, torch.ones_like(P[:, 0]), order)
            beta64, n64 = O.fit_wjet(P.double(), torch.ones_like(P[:, 0]).double(), order)
            ref_n, ex_n = O.world_normals(n, U).numpy(), O.world_normals(n64, U.double()).numpy()
        floor = O.unoriented_angle_deg(ref_n, ex_n)
        ang = O.unoriented_angle_deg(normals.cpu().numpy()[q], ref_n)
        assert (ang <= np.maximum(0.01, 2 * floor)).all(), "order %d: max angular deviation %g deg" % (order, ang.max())
        if order > 1:
            assert bool(torch.isfinite(curv).all())
            ref_c = (O.principal_curvatures(beta).double() / rad[:, None]).numpy()
            ex_c = (O.principal_curvatures(beta64).double() / rad[:, None]).numpy()
            cfloor = O.rectified_rel_err(ref_c, ex_c).max(1)
            c = curv.cpu().numpy()[q]
            # the PCA z-axis sign is library-defined in the reference: (k1,k2) -> (-k2,-k1) when it flips
            flipped = np.stack([-c[:, 1], -c[:, 0]], 1)
            err = np.minimum(O.rectified_rel_err(c, ref_c).max(1), O.rectified_rel_err(flipped, ref_c).max(1))
            assert (err <= np.maximum(1e-3, 2 * cfloor)).all(), "order %d: max curvature error %g" % (order, err.max())
        print("config2 %s order %d: max angle %.2e deg" % (kind, order, ang.max()))


def test_pcpnet_style_full_outputs(tmp_path):
    """BASELINE config 3 at a reduced size: multi-shape dataset with sparse .pidx centres on RAW points, full
    test_c_est.py exports, against the oracle on the same neighbourhoods."""
    from deepfit_b200 import pcpnet, synthetic
    from deepfit_b200.pipeline import model_from_state_dict
    from oracle import deepfit_oracle as O
    root = str(tmp_path / "pclouds")
    names = synthetic.write_pcpnet_like(root, n_shapes=3, n_points=6000, n_sparse=200)
    sd = O.seeded_state_dict(0)
    model = model_from_state_dict(sd, 256, 3, "sigmoid", "f16x3")
    ds = pcpnet.PointcloudPatchDataset(root, "testset_synthetic.txt", points_per_patch=256, use_pca=True,
                                       sparse_patches=True, neighbor_search_method='k')
    assert ds.shape_names == names and ds.shape_patch_count == [200, 200, 200]
    res = pcpnet.estimate_shapes(ds, model, str(tmp_path / "results"), batch=128)
    model._net.status()
    for si, name in enumerate(names):
        for ext, cols in (("normals", 3), ("curv", 2), ("weights", 256), ("beta", 10), ("trans", 9)):
            arr = np.loadtxt(tmp_path / "results" / (name + "." + ext))
            assert arr.shape == (200, cols)
            assert np.array_equal(arr.astype(np.float32), res[name][ext])
        pts = np.loadtxt(tmp_path / "pclouds" / (name + ".xyz")).astype("float32")    # raw, NOT normalised
        pidx = np.loadtxt(tmp_path / "pclouds" / (name + ".pidx")).astype("int64")
        ref_idx, ref_dist = O.knn_bruteforce(pts, pidx[:48], 256)
        patch, U, rad = ds.patches(si, 0, 48)
        assert np.array_equal(rad.cpu().numpy(), ref_dist[:, -1])
        with torch.no_grad():
            n, beta, w, trans, t2, _ = O.deepfit_forward(sd, patch.cpu(), 3)
            ref_n = O.world_normals(n, U.cpu(), trans).numpy()
            ref_c = (O.principal_curvatures(beta).double() / rad.cpu()[:, None]).numpy()
        assert O.unoriented_angle_deg(res[name]["normals"][:48], ref_n).max() <= 0.01
        assert O.rectified_rel_err(res[name]["curv"][:48], ref_c).max() <= 1e-3
        assert np.abs(res[name]["weights"][:48] - w.numpy()).max() < 2e-3
        assert np.abs(res[name]["trans"][:48] - trans.reshape(-1, 9).numpy()).max() < 1e-4
        # exact patch extraction on raw coordinates (dataset.py:292-334)
        raw_patch, _, _ = pcpnet.PointcloudPatchDataset(root, "testset_synthetic.txt", points_per_patch=256, use_pca=False,
                                                         sparse_patches=True).patches(si, 0, 4)
        for j in range(4):
            assert torch.equal(raw_patch[j].cpu(), O.extract_patch(pts, int(pidx[j]), ref_idx[j], ref_dist[j, -1]).t())


def test_deepfit_k128_single_tile_patches(clouds):
    """k=128: one 128-row tile per patch (the NT=1 instantiation of every fused kernel) vs the oracle."""
    from deepfit_b200 import ops
    from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
    from oracle import deepfit_oracle as O
    pts_np = O.normalize_cloud(clouds["Boxy_smooth100k"])[0]
    pts = torch.from_numpy(pts_np).cuda()
    sd = O.seeded_state_dict(3)
    model = model_from_state_dict(sd, 128, 2, "tanh", "f16x3")
    est = NormalEstimator(pts, 128, mode="DeepFit", model=model, chunk=256)
    begin, count = 7000, 301          # odd tile count in the last chunk: exercises the non-cluster head kernel too
    normals, curv = est.run(begin, begin + count)
    est.net.status()
    idx, rad = est.grid.query(128, q_begin=begin, q_count=count)
    patch, U = ops.patch_pca(pts, idx, rad, q_begin=begin)
    with torch.no_grad():
        n, beta, w, trans, t2, _ = O.deepfit_forward(sd, patch.cpu(), 2, weight_mode="tanh")
        ref_n = O.world_normals(n, U.cpu(), trans).numpy()
        ref_c = (O.principal_curvatures(beta).double() / rad.cpu()[:, None]).numpy()
    ang = O.unoriented_angle_deg(normals.cpu().numpy(), ref_n)
    assert ang.max() <= 0.01, "max angular deviation %g deg" % ang.max()
    assert O.rectified_rel_err(curv.cpu().numpy(), ref_c).max() <= 1e-3


def test_cli_deepfit_mode_with_saved_model(clouds, tmp_path):
    """The drop-in CLI in DeepFit mode: reads <trained_model_path>/DeepFit_params.pth (pickled Namespace) and
    DeepFit.pth (state_dict) exactly like tutorial/compute_normals.py:34-42, writes .normals / .curv."""
    import argparse
    from deepfit_b200 import compute_normals
    from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
    from deepfit_b200 import tutorial_utils as tu
    from oracle import deepfit_oracle as O
    mdir = tmp_path / "trained_models" / "DeepFit"
    mdir.mkdir(parents=True)
    sd = O.seeded_state_dict(0)
    torch.save(sd, mdir / "DeepFit.pth")
    params = argparse.Namespace(jet_order=3, use_point_stn=1, use_feat_stn=True, point_tuple=1, sym_op="max", arch="simple",
                                n_gaussians=1, weight_mode="sigmoid", points_per_patch=256)
    torch.save(params, mdir / "DeepFit_params.pth")
    raw = clouds["Boxy_smooth100k"][:6000]
    inp = tmp_path / "part.xyz"
    np.savetxt(inp, raw, fmt="%.6f")
    compute_normals.main(["--input", str(inp), "--output_path", str(tmp_path / "out"), "--mode", "DeepFit",
                          "--trained_model_path", str(mdir), "--k_neighbors", "256"])
    nrm = np.loadtxt(tmp_path / "out" / "part.normals")
    crv = np.loadtxt(tmp_path / "out" / "part.curv")
    assert nrm.shape == (6000, 3) and crv.shape == (6000, 2)
    ds = tu.SinglePointCloudDataset(str(inp), 256)
    model = model_from_state_dict(sd, 256, 3, "sigmoid", "f16x3")
    n2, c2 = NormalEstimator(ds.points_gpu, 256, mode="DeepFit", model=model).run()
    assert np.array_equal(nrm.astype(np.float32), n2.cpu().numpy())       # '%.18e' round-trips exactly
    assert np.array_equal(crv, c2.cpu().numpy())
    # --ref-compat reproduces the shipped script's omission of the QSTN cancellation (compute_normals.py:67)
    compute_normals.main(["--input", str(inp), "--output_path", str(tmp_path / "out2"), "--mode", "DeepFit",
                          "--trained_model_path", str(mdir), "--k_neighbors", "256", "--ref-compat"])
    nrm2 = np.loadtxt(tmp_path / "out2" / "part.normals")
    assert O.unoriented_angle_deg(nrm, nrm2).max() > 0.1


@pytest.mark.parametrize("use_pca", [False, True])
def test_pcpnet_ball_query_patches(tmp_path, use_pca):
    """neighbor_search_method='r' (reference dataset.py:289-366) on raw PCPNet-shaped clouds: the ball of
    patch_radius x bounding-box diagonal, seeded RandomState subsample when it is larger than points_per_patch, zero
    padding when smaller, PCA over the real points only.  The restatement below is the reference's code path with the
    ball in ascending index order (the documented deviation: scipy's traversal order is not reproducible)."""
    import scipy.spatial as spatial
    from deepfit_b200 import pcpnet, synthetic
    root = str(tmp_path / "pclouds")
    names = synthetic.write_pcpnet_like(root, n_shapes=2, n_points=20000, n_sparse=150)
    k, seed = 256, 3627473
    for radius in (0.012, 0.06):          # ~60 points per ball (padding) and ~1500 (subsampling)
        ds = pcpnet.PointcloudPatchDataset(root, "testset_synthetic.txt", patch_radius=[radius], points_per_patch=k,
                                           seed=seed, use_pca=use_pca, sparse_patches=True, neighbor_search_method='r')
        rng = np.random.RandomState(seed)
        saw_pad = saw_sub = False
        for si, name in enumerate(names):
            patch, trans, rad = ds.patches(si, 0, 150)
            patch, trans = patch.cpu(), trans.cpu()
            pts = np.loadtxt(os.path.join(root, name + ".xyz")).astype("float32")
            pidx = np.loadtxt(os.path.join(root, name + ".pidx")).astype("int")
            tree = spatial.cKDTree(pts, 10)
            bbdiag = float(np.linalg.norm(pts.max(0) - pts.min(0), 2))
            r_abs = bbdiag * radius
            assert float(rad[0]) == r_abs
            for i, c in enumerate(pidx):
                inds = np.sort(np.array(tree.query_ball_point(pts[c, :], r_abs)))
                count = int(min(k, len(inds)))
                if count < len(inds):
                    inds = inds[rng.choice(len(inds), count, replace=False)]
                    saw_sub = True
                saw_pad = saw_pad or count < k
                ref = torch.zeros(k, 3)
                ref[:count] = torch.from_numpy(pts[inds, :]) - torch.from_numpy(pts[c, :])
                ref[:count] = ref[:count] / r_abs
                ours = patch[i].t()                                   # [k,3]
                if not use_pca:
                    assert torch.equal(ours, ref), (name, i, count)
                    assert torch.equal(trans[i], torch.eye(3))
                else:
                    assert bool((ours[count:] == 0).all())            # padding stays exactly zero
                    U = trans[i].double()
                    assert (U.t() @ U - torch.eye(3, dtype=torch.float64)).abs().max() < 1e-6
                    assert (ref.double() @ U - ours.double()).abs().max() < 2e-6
                    cen = ref[:count].double() - ref[:count].double().mean(0)
                    D = U.t() @ (cen.t() @ cen) @ U                  # frame of the REAL points only
                    off = D - torch.diag(torch.diagonal(D))
                    assert off.abs().max() < 1e-5 * (cen.t() @ cen).abs().max()
        assert saw_pad if radius < 0.03 else saw_sub
    # and the export loop runs on ball-query patches (classic fit; curvature scaled by the ball radius)
    res = pcpnet.estimate_shapes(ds, None, str(tmp_path / "out"), jet_order=3, batch=64, write=False)
    for name in names:
        assert res[name]["normals"].shape == (150, 3) and np.isfinite(res[name]["normals"]).all()
        assert np.isfinite(res[name]["curv"]).all()
