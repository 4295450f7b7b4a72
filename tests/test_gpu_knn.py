"""Stage 1 parity (gpu): dfb_knn vs the reference's cKDTree output (golden, canonicalised) and vs the
oracle's exact (fp64 d^2, index) brute force.  Bit-exact: indices and fp64 distances must be equal."""
import numpy as np
import pytest
import torch

from conftest import CLOUD_NAMES

pytestmark = pytest.mark.gpu


def _normalize(raw):
    from oracle import deepfit_oracle as O
    return O.normalize_cloud(raw)[0]


def _query(points_np, q_np, k):
    from deepfit_b200 import ops
    pts = torch.from_numpy(np.ascontiguousarray(points_np)).cuda()
    grid = ops.Grid(pts)
    idx, rad, dist = grid.query(k, query=torch.from_numpy(q_np.astype(np.int32)).cuda(), return_dist=True)
    torch.cuda.synchronize()
    return idx.cpu().numpy(), rad.cpu().numpy(), dist.cpu().numpy()


def _assert_equal_mod_boundary_ties(idx, dist, ref_idx, ref_dist):
    """cKDTree picks an arbitrary member of a tie group that straddles the k-th place, so the SET may
    differ there; everything else, and every distance, must be identical."""
    assert np.array_equal(dist, ref_dist), "distances differ: max |d| = %g" % np.abs(dist - ref_dist).max()
    for r in range(idx.shape[0]):
        if np.array_equal(idx[r], ref_idx[r]):
            continue
        bad = np.nonzero(idx[r] != ref_idx[r])[0]
        assert np.all(dist[r][bad] == dist[r][-1]), "row %d differs away from the boundary tie group" % r


@pytest.mark.parametrize("name", CLOUD_NAMES)
def test_knn_matches_reference_ckdtree_golden(golden, clouds, name):
    pts = _normalize(clouds[name])
    q = golden[name + "/query"]
    idx, rad, dist = _query(pts, q, 256)
    _assert_equal_mod_boundary_ties(idx, dist, golden[name + "/knn_idx"], golden[name + "/knn_dist"])
    assert np.array_equal(rad, golden[name + "/rad"])
    idx, rad, dist = _query(pts, q[:8], 512)
    _assert_equal_mod_boundary_ties(idx, dist, golden[name + "/knn512_idx"], golden[name + "/knn512_dist"])


@pytest.mark.parametrize("name", CLOUD_NAMES)
def test_knn_matches_oracle_bruteforce_exactly(clouds, name):
    from oracle import deepfit_oracle as O
    pts = _normalize(clouds[name])
    rng = np.random.default_rng(5)
    q = rng.choice(len(pts), 96, replace=False)
    ref_idx, ref_dist = O.knn_bruteforce(pts, q, 256)
    idx, rad, dist = _query(pts, q, 256)
    assert np.array_equal(idx, ref_idx.astype(np.int32))
    assert np.array_equal(dist, ref_dist)


def _adversarial_clouds():
    rng = np.random.default_rng(11)
    g = np.stack(np.meshgrid(np.arange(16), np.arange(16), np.arange(16), indexing="ij"), -1).reshape(-1, 3)
    yield "lattice_exact_ties", g.astype(np.float32) * 0.125, 27
    base = rng.normal(size=(1500, 3)).astype(np.float32)
    yield "triplicated_points", np.concatenate([base, base, base])[rng.permutation(4500)], 64
    yield "uniform_volume_k17", rng.random((50000, 3)).astype(np.float32), 17
    dense = rng.normal(scale=0.003, size=(30000, 3))
    sparse = rng.normal(scale=1.0, size=(3000, 3))
    yield "variable_density", np.concatenate([dense, sparse]).astype(np.float32), 256
    yield "n_equals_k", rng.random((300, 3)).astype(np.float32), 300
    yield "k1", rng.random((1000, 3)).astype(np.float32), 1
    yield "planar_sheet", np.concatenate([rng.random((20000, 2)), np.zeros((20000, 1))], 1).astype(np.float32), 100
    yield "collinear", np.stack([np.linspace(0, 1, 5000), np.zeros(5000), np.zeros(5000)], 1).astype(np.float32), 33
    yield "k1024", rng.normal(size=(20000, 3)).astype(np.float32), 1024
    # sub-grids of heavy cells (csrc/knn.cu): blobs three and six orders of magnitude denser than their surroundings,
    # queries on both sides of the heavy / light boundary; exact ties and duplicates inside a heavy cell; k = 512
    nested = np.concatenate([rng.normal(scale=1.0, size=(20000, 3)), rng.normal(scale=0.05, size=(20000, 3)) + 0.3,
                             rng.normal(scale=0.002, size=(20000, 3)) + 0.3, rng.normal(scale=0.0005, size=(6000, 3)) - 0.7])
    yield "nested_density", nested.astype(np.float32), 256
    yield "nested_density_k512", nested.astype(np.float32), 512
    lat = (g.astype(np.float32) * 1e-4 + 0.5)
    yield "heavy_cell_lattice_ties", np.concatenate([lat, lat[:1000], rng.random((8000, 3)).astype(np.float32)])[rng.permutation(13096)], 40
    yield "heavy_cell_identical_points", np.concatenate([np.full((2000, 3), 0.25, np.float32), rng.random((6000, 3)).astype(np.float32)]), 300
    sheet = np.concatenate([rng.random((40000, 2)) * 0.01 + 0.4, rng.normal(scale=1e-5, size=(40000, 1))], 1)
    yield "dense_thin_sheet_in_sparse_box", np.concatenate([sheet, rng.random((4000, 3))]).astype(np.float32), 128


@pytest.mark.parametrize("case", list(_adversarial_clouds()), ids=lambda c: c[0])
def test_knn_adversarial_exact(case):
    from oracle import deepfit_oracle as O
    _, pts, k = case
    rng = np.random.default_rng(3)
    q = rng.choice(len(pts), min(64, len(pts)), replace=False)
    ref_idx, ref_dist = O.knn_bruteforce(pts, q, k)
    idx, rad, dist = _query(pts, q, k)
    assert np.array_equal(idx, ref_idx.astype(np.int32))
    assert np.array_equal(dist, ref_dist)
    assert np.array_equal(rad, ref_dist[:, -1])


def test_knn_full_size_properties():
    """1M-point synthetic torus (BASELINE config 2 shape): every query, size-independent properties, plus
    256 random queries against an fp64 brute force evaluated with torch on the device."""
    from deepfit_b200 import ops, synthetic
    pts_np = synthetic.torus_cloud(1_000_000, seed=1234)
    pts = torch.from_numpy(pts_np).cuda()
    grid = ops.Grid(pts)
    k = 256
    n = pts.shape[0]
    rng = np.random.default_rng(0)
    sample = torch.from_numpy(rng.choice(n, 256, replace=False).astype(np.int64)).cuda()
    ar = torch.arange(n, device="cuda")
    for s in range(0, n, 250_000):
        idx, rad, dist = grid.query(k, q_begin=s, q_count=250_000, return_dist=True)
        assert bool((dist[:, 1:] >= dist[:, :-1]).all())                 # sorted
        assert bool((dist[:, 0] == 0).all())                             # self (or a duplicate) first
        assert bool((rad == dist[:, -1]).all())
        assert int(idx.min()) >= 0 and int(idx.max()) < n
        srt = torch.sort(idx.long(), dim=1)[0]
        assert bool((srt[:, 1:] != srt[:, :-1]).all())                   # no repeated neighbour
        # recompute the distances of the reported neighbours exactly
        qp = pts[s:s + 250_000].double()
        nb = pts[idx.long().reshape(-1)].double().reshape(-1, k, 3)
        dx = nb[:, :, 0] - qp[:, None, 0]
        d2 = dx * dx
        dy = nb[:, :, 1] - qp[:, None, 1]
        d2 = d2 + dy * dy
        dz = nb[:, :, 2] - qp[:, None, 2]
        d2 = d2 + dz * dz
        assert bool((torch.sqrt(d2) == dist).all())
        del nb, dx, dy, dz, d2
    pd = pts.double()
    idx, rad, dist = grid.query(k, query=sample.int(), return_dist=True)
    for j in range(sample.numel()):
        qv = pd[sample[j]]
        dx = pd[:, 0] - qv[0]
        d2 = dx * dx
        dy = pd[:, 1] - qv[1]
        d2 = d2 + dy * dy
        dz = pd[:, 2] - qv[2]
        d2 = d2 + dz * dz
        val, order = torch.sort(d2, stable=True)                         # stable => ties by ascending index
        assert bool((order[:k] == idx[j].long()).all()), "query %d" % j
        assert bool((torch.sqrt(val[:k]) == dist[j]).all())


def _bruteforce_rows(pts, queries, k):
    """Exact kNN of a few queries over a LARGE cloud by brute force in torch fp64 on the device (library ops as the
    checker): key (d2 accumulated x -> y -> z without contraction, index).  Returns (idx i64 [Q,k], dist f64 [Q,k])."""
    px, py, pz = pts[:, 0].double(), pts[:, 1].double(), pts[:, 2].double()
    out_i = torch.empty((queries.numel(), k), dtype=torch.int64, device=pts.device)
    out_d = torch.empty((queries.numel(), k), dtype=torch.float64, device=pts.device)
    for j in range(queries.numel()):
        c = int(queries[j])
        dx = px - px[c]
        d2 = dx * dx
        dy = py - py[c]
        d2 = d2 + dy * dy
        dz = pz - pz[c]
        d2 = d2 + dz * dz
        kth = torch.topk(d2, k, largest=False)[0].max()
        cand = torch.nonzero(d2 <= kth).flatten()            # every member of the tie group at the k-th place included
        v = d2[cand]                                          # cand ascending => a stable sort by value breaks ties by index
        order = torch.argsort(v, stable=True)[:k]
        out_i[j] = cand[order]
        out_d[j] = torch.sqrt(v[order])
    return out_i, out_d


@pytest.mark.parametrize("kind,n,seed,ks", [("multi", 10_000_000, 4321, (256, 512)), ("density", 32_000_000, 9876, (256,))],
                         ids=["config4_multi_10M", "config5_density_32M"])
def test_knn_exact_at_named_sizes(kind, n, seed, ks):
    """BASELINE configs 4 / 5 at size: the 10 M-point multi-surface scan (k = 256 and 512) and a
    32 M-point variable-density scan (1/r^2 falloff: cell occupancies from a handful to tens of thousands), 256 random
    queries each plus the 64 queries in the densest cell region, bit-equal -- indices and fp64 distances -- to an fp64
    brute force over the whole cloud; size-independent properties on a contiguous slab of queries."""
    from deepfit_b200 import ops, synthetic
    pts = synthetic.device_cloud(kind, n, seed, "cuda")
    grid = ops.Grid(pts)
    rng = np.random.default_rng(5)
    sample = torch.from_numpy(rng.choice(n, 256, replace=False).astype(np.int64)).cuda()
    # the points nearest to the cloud's densest spot (the sensor ring of the density scan / anywhere for the multi scan)
    centre = pts[torch.randint(0, n, (1,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))]
    if kind == "density":
        r2 = (pts[:, 0].double() ** 2 + pts[:, 1].double() ** 2)
        centre = pts[torch.argmin(r2)][None]
    dense = torch.topk(((pts - centre) ** 2).sum(1), 64, largest=False)[1]
    sample = torch.cat([sample, dense])
    for k in ks:
        idx, rad, dist = grid.query(k, query=sample.int(), return_dist=True)
        ref_i, ref_d = _bruteforce_rows(pts, sample, k)
        assert bool((idx.long() == ref_i).all()), "k=%d: rows %s differ" % (k, torch.nonzero((idx.long() != ref_i).any(1)).flatten()[:8].tolist())
        assert bool((dist == ref_d).all()) and bool((rad == ref_d[:, -1]).all())
        # properties on a slab of consecutive queries through the range form of the call (what the pipeline uses)
        s = n // 3
        idx, rad, dist = grid.query(k, q_begin=s, q_count=100_000, return_dist=True)
        assert bool((dist[:, 1:] >= dist[:, :-1]).all()) and bool((dist[:, 0] == 0).all()) and bool((rad == dist[:, -1]).all())
        srt = torch.sort(idx.long(), dim=1)[0]
        assert bool((srt[:, 1:] != srt[:, :-1]).all())
        sub = torch.arange(s, s + 100_000, 3125, device="cuda")
        ref_i, ref_d = _bruteforce_rows(pts, sub, k)
        assert bool((idx[sub - s].long() == ref_i).all()) and bool((dist[sub - s] == ref_d).all())


def _ball_sets(pts_np, q, radius):
    from deepfit_b200 import ops
    pts = torch.from_numpy(np.ascontiguousarray(pts_np)).cuda()
    grid = ops.Grid(pts)
    qt = torch.from_numpy(np.asarray(q, dtype=np.int32)).cuda()
    rad = torch.from_numpy(np.asarray(radius, dtype=np.float64)).cuda() if np.ndim(radius) else float(radius)
    off, idx = grid.ball_query(rad, query=qt)
    off, idx = off.cpu().numpy(), idx.cpu().numpy()
    return [idx[off[i]:off[i + 1]] for i in range(len(q))]


def test_morton_order_is_a_permutation_of_the_range(clouds):
    """dfb_grid_morton_order: every index of [begin, end) exactly once, in the order of the grid's sorted copy (so
    consecutive entries are spatial neighbours); and a query list in that order returns the same rows as file order."""
    from deepfit_b200 import ops
    pts_np = _normalize(clouds["0000000000"])
    pts = torch.from_numpy(pts_np).cuda()
    grid = ops.Grid(pts)
    n = len(pts_np)
    whole = grid.morton_order()
    assert torch.equal(torch.sort(whole.long())[0], torch.arange(n, device="cuda"))
    for b, e in ((0, 1), (12345, 12345), (777, 40000), (n - 5000, n)):
        part = grid.morton_order(b, e)
        assert part.numel() == e - b
        keep = whole[(whole >= b) & (whole < e)]
        assert torch.equal(part, keep)              # the sub-range keeps the whole cloud's relative order
    # spatial coherence: consecutive entries are far closer than random pairs
    p = pts[whole.long()]
    step = (p[1:] - p[:-1]).norm(dim=1).median().item()
    rnd = (pts[torch.randperm(n, device="cuda")[: n - 1]] - pts[:-1]).norm(dim=1).median().item()
    assert step < 0.05 * rnd
    part = grid.morton_order(777, 3000)
    idx_m, rad_m = grid.query(256, query=part)
    idx_f, rad_f = grid.query(256, q_begin=777, q_count=3000 - 777)
    inv = (part.long() - 777)
    assert torch.equal(idx_m, idx_f[inv]) and torch.equal(rad_m, rad_f[inv])


@pytest.mark.parametrize("name", CLOUD_NAMES)
def test_ball_query_matches_ckdtree_sets(clouds, name):
    """neighbor_search_method='r' (reference dataset.py:289-291): the SET cKDTree.query_ball_point returns, for the
    radii PCPNet uses (fractions of the bounding-box diagonal of the raw cloud), bit for bit."""
    import scipy.spatial as spatial
    pts = clouds[name]
    tree = spatial.cKDTree(pts, 10)
    bbdiag = float(np.linalg.norm(pts.max(0) - pts.min(0), 2))
    rng = np.random.default_rng(17)
    q = rng.choice(len(pts), 64, replace=False)
    for frac in (0.01, 0.03, 0.05):
        rad = bbdiag * frac
        ours = _ball_sets(pts, q, rad)
        for i, c in enumerate(q):
            ref = np.sort(np.asarray(tree.query_ball_point(pts[c, :], rad), dtype=np.int32))
            assert np.array_equal(ours[i], ref), (name, frac, int(c), len(ours[i]), len(ref))
    # one radius per query
    rads = bbdiag * rng.uniform(0.005, 0.06, size=len(q))
    ours = _ball_sets(pts, q, rads)
    for i, c in enumerate(q):
        ref = np.sort(np.asarray(tree.query_ball_point(pts[c, :], float(rads[i])), dtype=np.int32))
        assert np.array_equal(ours[i], ref)


def test_ball_query_edge_cases():
    import scipy.spatial as spatial
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(12), indexing="ij"), -1).reshape(-1, 3)
    lattice = (g.astype(np.float32) * 0.125)
    rng = np.random.default_rng(23)
    base = rng.normal(size=(2000, 3)).astype(np.float32)
    dup = np.concatenate([base, base])[rng.permutation(4000)]
    for pts, radii in ((lattice, (0.0, 0.125, 0.125 * np.sqrt(2.0), 0.3, 10.0)), (dup, (0.0, 1e-9, 0.2, 100.0))):
        tree = spatial.cKDTree(pts, 10)
        q = rng.choice(len(pts), 48, replace=False)
        for r in radii:
            ours = _ball_sets(pts, q, float(r))
            for i, c in enumerate(q):
                ref = np.sort(np.asarray(tree.query_ball_point(pts[c, :], float(r)), dtype=np.int32))
                assert np.array_equal(ours[i], ref), (float(r), int(c), len(ours[i]), len(ref))
