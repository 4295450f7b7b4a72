"""Stage 3 parity (gpu): the tcgen05 GEMM building block against torch, and the whole weight network
against the reference module's outputs (golden, seeded state_dict) and the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import CLOUD_NAMES

pytestmark = pytest.mark.gpu


def _dbg_gemm(A, Bm, bias, mode, nprod, relu, k_pts):
    from deepfit_b200 import _lib
    lib = _lib.load()
    M, K = A.shape
    N = Bm.shape[0]
    rows_out = M if mode == 0 else M // k_pts
    out = torch.full((rows_out, N), float("nan"), device="cuda")
    out_t = torch.full((rows_out, N), float("nan"), device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    rc = lib.dfb_dbg_gemm(C.c_void_p(A.data_ptr()), C.c_void_p(Bm.data_ptr()),
                          C.c_void_p(bias.data_ptr()) if bias is not None else None,
                          M, N, K, mode, nprod, relu, k_pts, C.c_void_p(out.data_ptr()), C.c_void_p(out_t.data_ptr()), st)
    _lib.check(rc)
    torch.cuda.synchronize()
    return out, out_t


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


# nprod = tensor-core product units per k-step: 3 = split bf16 (hi*hi + hi*lo + lo*hi), 1 = single bf16,
# 2 = f16f8 (fp16 main product + one e4m3 MMA carrying both first-order corrections), 4 = f16x3 (split FP16, three
# products, 22-bit operands -- the parity mode)
GEMM_TOL = {4: 2 ** -18, 3: 2 ** -13, 2: 2 ** -12, 1: 2 ** -9}          # relative to the largest sum of |products|
RESPLIT_TOL = {4: 2 ** -20, 3: 2 ** -15, 2: 2 ** -14, 1: 2 ** -7}       # the re-split tiled output, relative to the largest value
NPRODS = [4, 3, 2, 1]


def _ref(A, Bm, bias, relu, nprod):
    if nprod == 1:
        A, Bm = _bf16(A), _bf16(Bm)
    r = A.double() @ Bm.double().t()
    if bias is not None:
        r = r + bias.double()
    if relu:
        r = torch.relu(r)
    return r


GEMM_SHAPES = [(256, 128, 64), (384, 256, 128), (256, 512, 512), (512, 64, 64), (128, 128, 1024), (1024, 256, 256)]


@pytest.mark.parametrize("nprod", NPRODS)
@pytest.mark.parametrize("shape", GEMM_SHAPES, ids=lambda s: "M%dN%dK%d" % s)
def test_umma_gemm_points_as_rows(shape, nprod):
    M, N, K = shape
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = torch.randn(M, K, device="cuda", generator=g)
    Bm = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    out, out_t = _dbg_gemm(A, Bm, bias, 0, nprod, 1, 128)
    ref = _ref(A, Bm, bias, True, nprod)
    # error model: split-bf16 keeps 16 mantissa bits per operand (2^-16 relative per product, plus the
    # dropped lo*lo term), plain bf16 keeps 8, f16f8 11 bits + corrections good to 2^-4 of 2^-12; all accumulate in fp32
    mag = (A.abs().double() @ Bm.abs().double().t()).max().item()
    tol = mag * GEMM_TOL[nprod]
    err = (out.double() - ref).abs().max().item()
    assert err < tol, "GEMM mismatch: max err %g (tol %g); sample out %s ref %s" % (
        err, tol, out[0, :4].tolist(), ref[0, :4].tolist())
    # the re-split tiled output (next layer's operand) carries the same values to 2^-16 relative
    tol_t = tol + ref.abs().max().item() * RESPLIT_TOL[nprod]
    assert (out_t.double() - ref).abs().max().item() < tol_t


def test_tensor_core_accumulate_truncates_and_corrections_apart(dbg_switch):
    """What bounds the parity mode's accuracy: tcgen05.mma TRUNCATES its fp32 accumulator after
    every MMA.  Operands exact in split FP16, C against the float64 sum, in units of 2^-24 of sum|products|:
    positive operands (a growing accumulator) show a negative bias that grows with the number of MMAs (K = 1024, three
    products per k-step: -121 measured; a round-to-nearest accumulate would be unbiased within +-2); keeping the
    hi*lo + lo*hi corrections out of the main accumulator (dfb_dbg_set key 11, the product default for this kernel) takes
    their truncations away: -88, and on signed operands the largest error drops from 8.0 to 3.3."""
    g = torch.Generator(device="cuda").manual_seed(1)
    M, N, K = 256, 128, 1024
    A = (torch.rand(M, K, device="cuda", generator=g) + 0.5) / 32
    Bm = (torch.rand(N, K, device="cuda", generator=g) + 0.5) / 32
    sign = (torch.randint(0, 2, (M, K), device="cuda", generator=g) * 2 - 1).float()
    stats = {}
    for name, a in (("positive", A), ("signed", A * sign)):
        ref = a.double() @ Bm.double().t()
        scale = a.abs().double() @ Bm.abs().double().t()
        for apart in (0, 1):
            dbg_switch(11, apart, 1)
            out, _ = _dbg_gemm(a.contiguous(), Bm.contiguous(), None, 0, 4, 0, 128)
            rel = (out.double() - ref) / scale * 2 ** 24
            stats[name, apart] = (rel.mean().item(), rel.abs().max().item())
    print({"%s, corrections %s" % (k_[0], "apart" if k_[1] else "in the accumulator"): "mean %+.1f max %.1f" % v for k_, v in stats.items()})
    assert stats["positive", 0][0] < -60 and stats["positive", 1][0] < -40            # truncation, not rounding
    assert stats["positive", 1][0] > 0.85 * stats["positive", 0][0]                    # ... fewer of them
    assert stats["signed", 1][1] < 0.6 * stats["signed", 0][1]


@pytest.mark.parametrize("nprod", NPRODS)
@pytest.mark.parametrize("N", [256, 1024])
@pytest.mark.parametrize("k_pts", [128, 256, 512])
def test_umma_gemm_channels_as_rows_with_maxpool(k_pts, N, nprod):
    n_patch, K = 5, 128
    g = torch.Generator(device="cuda").manual_seed(k_pts)
    A = torch.randn(n_patch * k_pts, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    out, out_t = _dbg_gemm(A, W, bias, 1, nprod, 0, k_pts)
    full = _ref(A, W, None, False, nprod).reshape(n_patch, k_pts, N)
    ref = full.max(1)[0] + bias.double()
    mag = (A.abs().double() @ W.abs().double().t()).max().item()
    tol = mag * GEMM_TOL[nprod]
    assert (out.double() - ref).abs().max().item() < tol
    assert (out_t.double() - ref).abs().max().item() < tol + ref.abs().max().item() * RESPLIT_TOL[nprod]


def _torch_network(layers, P, mode):
    """The same algebra as the CUDA pipeline (folded BN, split head) in plain torch on the device; used to
    localise a failure stage by stage.  mode 'fp32' or 'bf16'."""
    from oracle import deepfit_oracle as O
    r = (lambda x: x) if mode == "fp32" else _bf16
    lin = lambda x, i, cols=None: r(x) @ r(layers[i][0].cuda() if cols is None else layers[i][0].cuda()[:, cols]).t()  # noqa: E731
    b = lambda i: layers[i][1].cuda()  # noqa: E731
    X = P.transpose(1, 2)
    out = {}
    x = torch.relu(X @ layers[0][0].cuda().t() + b(0))
    x = torch.relu(lin(x, 1) + b(1))
    g1 = torch.relu(lin(x, 2) + b(2)).max(1)[0]
    f = torch.relu(lin(g1, 3) + b(3))
    f = torch.relu(lin(f, 4) + b(4))
    q = f @ layers[5][0].cuda().t() + b(5) + torch.tensor([1.0, 0, 0, 0], device="cuda")
    R = O.quat_to_rotmat(q)
    out["trans"] = R
    Xr = X @ R
    out["x64a"] = x1 = torch.relu(Xr @ layers[6][0].cuda().t() + b(6))
    out["x64b"] = x = torch.relu(lin(x1, 7) + b(7))
    s = torch.relu(lin(x, 8) + b(8))
    s = torch.relu(lin(s, 9) + b(9))
    g2 = torch.relu(lin(s, 10) + b(10)).max(1)[0]
    out["f512"] = f = torch.relu(lin(g2, 11) + b(11))
    out["f256"] = f = torch.relu(lin(f, 12) + b(12))
    t2 = (lin(f, 13) + b(13) + torch.eye(64, device="cuda").reshape(1, 4096)).reshape(-1, 64, 64)
    out["trans2"] = t2
    out["x64c"] = pf = torch.bmm(r(x), r(t2))
    out["x128"] = y = torch.relu(lin(pf, 14) + b(14))
    out["gfeat"] = g3 = (lin(y, 15) + b(15)).max(1)[0]
    hg = lin(g3, 16, slice(0, 1024)) + b(16)
    out["h512"] = h = torch.relu(lin(pf, 16, slice(1024, 1088)) + hg[:, None, :])
    out["h256"] = h = torch.relu(lin(h, 17) + b(17))
    h = torch.relu(lin(h, 18) + b(18))
    a = (h @ layers[19][0].cuda().t() + b(19)).squeeze(-1)
    out["weights"] = 0.01 + torch.sigmoid(a)
    return out


def _dump(net, which, rows, cols):
    from deepfit_b200 import _lib
    o = torch.empty((rows, cols), device="cuda")
    _lib.check(_lib.load().dfb_dbg_model_dump(net._h, which, rows, cols, C.c_void_p(o.data_ptr()),
                                              C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return o


@pytest.fixture
def dbg_switch():
    """Flip a library debug switch (dfb_dbg_set) for one test and restore the default afterwards."""
    from deepfit_b200 import _lib
    lib = _lib.load()
    changed = []

    def _set(key, value, default):
        lib.dfb_dbg_set(key, value)
        changed.append((key, default))
    yield _set
    for key, default in changed:
        lib.dfb_dbg_set(key, default)


@pytest.mark.parametrize("variant", ["fused", "fused_l12_only", "unfused"])
@pytest.mark.parametrize("precision", ["f16x3", "bf16x3", "f16f8", "f16mix", "bf16"])
def test_network_stage_by_stage_vs_torch(golden, precision, variant, dbg_switch):
    from deepfit_b200 import DeepFit as D
    from deepfit_b200 import ops
    from oracle import deepfit_oracle as O
    fused = 0 if variant == "unfused" else 1
    l3 = 1 if variant == "fused" else 0
    dbg_switch(2, fused, 1)                                    # fused head (64->512->256) vs two launches
    dbg_switch(4, l3, 1)                                       # 256->128->1 + weight map fused behind it
    chain = 0 if variant == "unfused" else 1
    dbg_switch(5, chain, 1)                                    # points -> ... -> 1024+max chains in one kernel each
    name = CLOUD_NAMES[0]
    P = torch.from_numpy(golden[name + "/patch"][:16]).cuda()
    layers = D.fold_batchnorm(O.seeded_state_dict(0))
    net = ops.WeightNetwork(layers, 256, "sigmoid", precision, max_batch=128)
    w, trans, trans2, pts = net.forward(P, want_trans2=True, want_points=True)
    net.status()
    ref = _torch_network(layers, P, "bf16" if precision == "bf16" else "fp32")
    B, k = 16, 256
    report = {}
    for which, nm, rows, cols in [(0, "x64a", B * k, 64), (1, "x64b", B * k, 64), (2, "x64c", B * k, 64),
                                  (3, "x128", B * k, 128), (4, "h512", B * k, 512), (5, "h256", B * k, 256),
                                  (6, "gfeat", B, 1024), (7, "f512", B, 512), (8, "f256", B, 256)]:
        if (nm == "h512" and fused) or (nm == "h256" and fused and l3):
            continue                      # never materialised by the fused head kernel
        if chain and nm in ("x64a", "x64b", "x128"):
            continue                      # stay in shared memory inside the fused chain kernels
        got = _dump(net, which, rows, cols)
        r = ref[nm].reshape(rows, cols)
        report[nm] = ((got - r).abs().max() / r.abs().max().clamp_min(1e-6)).item()
    report["trans"] = (trans - ref["trans"]).abs().max().item()
    report["trans2"] = (trans2 - ref["trans2"]).abs().max().item()
    report["weights"] = (w - ref["weights"]).abs().max().item()
    report["pts"] = (pts - torch.bmm(P.transpose(1, 2), ref["trans"]).transpose(1, 2)).abs().max().item()
    tol = {"f16x3": 2e-5, "bf16x3": 3e-4, "f16f8": 6e-4, "f16mix": 6e-4, "bf16": 6e-2}[precision]
    print(precision, variant, {k_: "%.1e" % v for k_, v in report.items()})
    bad = {k_: v for k_, v in report.items() if not v < tol}
    assert not bad, "stage errors (relative to stage max): %s" % report


# f16x3 (split FP16, 22-bit operands) is THE parity mode: every gate of BASELINE.json holds on every row.  bf16x3 and
# f16f8 (16- / 15-bit operands) are faster approximations: their weights are 1e-4 / 3e-4 off the reference's, which the
# normals tolerate (gate holds) but the curvatures of ill-conditioned lidar patches do not (measured up to 1.2x / 2.9x the
# gate); their tests document that bound instead of claiming parity.
#
# f16mix is the per-layer precision ladder VERDICT r1 asked for, measured on the GPU (tools/mix_probe.py,
# profiles/r2_mix_probe.log): f16x3 operands everywhere, but the 128 -> 1024 max-pool layers of the QSTN and feature-STN
# chains (42 % of the network's MACs) issue the FP16 hi*hi product only.  +21 % throughput; every gate of the smooth cloud
# holds (normals 0.17, curvatures 0.53 of their gates), but NO single-product layer survives the lidar cloud: its
# ill-conditioned rows sit at 0.62 of the curvature gate already in f16x3 (the reference's own fp32 noise), and one
# 11-bit layer moves 7-18 of the 512 rows over it (QSTN max layer alone 6.4x, feature-STN max layer alone 3.1x, both
# 7.1x; the encoder's 46x; normals 1.04x on one row).  An approximation for CAD-like clouds, not a parity mode.
PARITY_PRECISIONS = ["f16x3"]
ALL_PRECISIONS = ["f16x3", "bf16x3", "f16f8"]
GATE_PRECISIONS = ALL_PRECISIONS + ["f16mix"]
WEIGHT_TOL = {"f16x3": 5e-5, "bf16x3": 3e-4, "f16f8": 6e-4, "f16mix": 6e-4}
LIDAR_CURV_FACTOR = {"f16x3": 1.0, "bf16x3": 3.0, "f16f8": 4.0, "f16mix": 10.0}     # allowed multiple of the per-row curvature gate
#                    measured:        0.62           1.2 - 2.0      2.9           7.1        (bf16x3 moves with the accumulation order)
LIDAR_ANGLE_FACTOR = {"f16mix": 1.5}                                                # ... of the per-row normal gate (default 1)


@pytest.mark.parametrize("precision", GATE_PRECISIONS)
@pytest.mark.parametrize("name", CLOUD_NAMES)
def test_deepfit_forward_matches_reference_golden(golden, name, precision):
    """DeepFit.forward drop-in vs the reference module (eval mode) on the reference's own patches: every output."""
    from deepfit_b200.pipeline import model_from_state_dict
    from oracle import deepfit_oracle as O
    P = torch.from_numpy(golden[name + "/patch"][:16]).cuda()
    model = model_from_state_dict(O.seeded_state_dict(0), 256, 3, "sigmoid", precision)
    normal, beta, weights, trans, trans2, nn_ = model(P)
    assert nn_ is None
    assert tuple(normal.shape) == (16, 3) and tuple(beta.shape) == (16, 10) and tuple(weights.shape) == (16, 256)
    assert tuple(trans.shape) == (16, 3, 3) and tuple(trans2.shape) == (16, 64, 64)
    dw = (weights.cpu() - torch.from_numpy(golden[name + "/net_weights"])).abs().max().item()
    dt = (trans.cpu() - torch.from_numpy(golden[name + "/net_trans"])).abs().max().item()
    dt2 = (trans2.cpu() - torch.from_numpy(golden[name + "/net_trans2"])).abs().max().item()
    print("%s %s: max |dw| %.2e, |dtrans| %.2e, |dtrans2| %.2e" % (name, precision, dw, dt, dt2))
    assert dw < WEIGHT_TOL[precision] and dt < 1e-4 and dt2 < 1e-4


@pytest.mark.parametrize("precision", GATE_PRECISIONS)
@pytest.mark.parametrize("name", CLOUD_NAMES)
def test_deepfit_forward_per_row_gate(golden, name, precision):
    """BASELINE.json's gates (0.01 deg unoriented, 1e-3 rectified-relative curvature) on 512 patches per cloud against
    the unmodified reference's outputs, PER ROW: gate_i = max(gate, 2 x floor_i), where floor_i is the reference's own
    fp32 noise on that row -- the largest deviation from the fp64-exact result among 9 mathematically equivalent fp32
    evaluations of the reference (its forward on the patch and on 8 neighbour permutations of it; generated by
    oracle/make_golden.py).  On the smooth cloud no row is relaxed; on the lidar cloud (scan lines -> ill-conditioned
    fits) a few percent are.  The achieved maxima are printed.  The parity mode (f16x3) must hold every gate; the
    approximate modes must hold the normal gate everywhere, the curvature gate on the smooth cloud and stay inside
    LIDAR_CURV_FACTOR x the curvature gate on the lidar cloud."""
    from deepfit_b200.pipeline import model_from_state_dict
    from oracle import deepfit_oracle as O
    P = torch.from_numpy(golden[name + "/gate_patch"]).cuda()
    U = torch.from_numpy(golden[name + "/gate_U"])
    rad = golden[name + "/gate_rad"]
    model = model_from_state_dict(O.seeded_state_dict(0), 256, 3, "sigmoid", precision)
    normal, beta, weights, trans, trans2, _ = model(P)
    dw = (weights.cpu() - torch.from_numpy(golden[name + "/gate_weights"])).abs().max().item()
    dt = (trans.cpu() - torch.from_numpy(golden[name + "/gate_trans"])).abs().max().item()
    nw = O.world_normals(normal.cpu(), U, trans.cpu()).numpy()
    ang = O.unoriented_angle_deg(nw, golden[name + "/gate_normal_world"])
    cerr = O.rectified_rel_err(model_curv(beta, rad), golden[name + "/gate_curv_scaled"]).max(1)
    gate_a = np.maximum(0.01, 2 * golden[name + "/gate_floor_angle_deg"])
    gate_c = np.maximum(1e-3, 2 * golden[name + "/gate_floor_curv"])
    strict_a, strict_c = gate_a <= 0.01, gate_c <= 1e-3
    print("%s %s: %d rows; max |dw| %.2e |dtrans| %.2e; angle max %.3e deg on the %d strict rows (gate 0.01), worst "
          "angle/gate %.2f over all rows; curvature max %.3e on the %d strict rows (gate 1e-3), worst err/gate %.2f"
          % (name, precision, len(ang), dw, dt, ang[strict_a].max(), int(strict_a.sum()), (ang / gate_a).max(),
             cerr[strict_c].max(), int(strict_c.sum()), (cerr / gate_c).max()))
    assert dw < WEIGHT_TOL[precision], dw
    assert dt < 1e-4, dt
    af = LIDAR_ANGLE_FACTOR.get(precision, 1.0) if name == "0000000000" else 1.0
    assert (ang <= af * gate_a).all(), "rows over their gate: %s" % np.nonzero(ang > af * gate_a)[0][:10]
    cf = LIDAR_CURV_FACTOR[precision] if name == "0000000000" else 1.0
    assert (cerr <= cf * gate_c).all(), "rows over their gate: %s" % np.nonzero(cerr > cf * gate_c)[0][:10]
    if name == "Boxy_smooth100k":
        assert strict_a.all() and strict_c.all()     # the smooth cloud needs no relaxation at all


def test_f16mix_layer_masks():
    """'f16mix' == its default layer set spelled out, bit for bit; 'f16mix' with a layer the ladder does not run
    single-product by default differs from it; an empty set cannot be spelled (mask 0 means the default)."""
    from deepfit_b200 import DeepFit as D
    from deepfit_b200 import ops
    from oracle import deepfit_oracle as O
    g = torch.Generator().manual_seed(5)
    patch = (torch.rand((200, 3, 256), generator=g) - 0.5).cuda()
    layers = D.fold_batchnorm(O.seeded_state_dict(0))
    out = {}
    for p in ("f16mix", "f16mix:qstn_max+stn_max", "f16mix:qstn_max", "f16x3"):
        net = ops.WeightNetwork(layers, 256, "sigmoid", p, max_batch=256)
        w, trans, t2, _ = net.forward(patch, want_trans2=True)
        net.status()
        out[p] = (w.clone(), trans.clone(), t2.clone())
        net.close()
    for a, b in zip(out["f16mix"], out["f16mix:qstn_max+stn_max"]):
        assert torch.equal(a, b)
    assert not torch.equal(out["f16mix"][0], out["f16mix:qstn_max"][0])
    assert not torch.equal(out["f16mix:qstn_max"][1], out["f16x3"][1])          # the QSTN rotation feels its max layer
    assert torch.equal(out["f16mix:qstn_max"][1], out["f16mix"][1])            # ... and nothing of the feature STN's
    assert (out["f16mix"][0] - out["f16x3"][0]).abs().max() < 2e-3


def model_curv(beta, rad):
    from deepfit_b200 import DeepFit as D
    c, _ = D.compute_principal_curvatures(beta)
    return c.cpu().numpy().astype(np.float64) / rad[:, None]


def test_forward_is_batch_independent_and_chunked():
    """Eval-mode semantics: results do not depend on batch composition; n > max_batch is chunked."""
    from deepfit_b200 import DeepFit as D
    from deepfit_b200 import ops
    from oracle import deepfit_oracle as O
    g = torch.Generator().manual_seed(0)
    P = torch.randn(300, 3, 256, generator=g).cuda() * 0.3
    layers = D.fold_batchnorm(O.seeded_state_dict(1))
    net = ops.WeightNetwork(layers, 256, "sigmoid", "f16x3", max_batch=128)
    w_all, t_all, _, _ = net.forward(P, want_trans2=False)
    w_one, t_one, _, _ = net.forward(P[137:138].contiguous(), want_trans2=False)
    net.status()
    assert torch.equal(w_all[137:138], w_one) and torch.equal(t_all[137:138], t_one)
    with pytest.raises(ValueError):
        net.forward(torch.zeros(2, 3, 128, device="cuda"))


@pytest.mark.parametrize("precision", ALL_PRECISIONS)
@pytest.mark.parametrize("B,k", [(3000, 256), (1, 256), (149, 256), (333, 128), (600, 512)])
def test_persistent_kernels_match_one_cta_per_tile(dbg_switch, B, k, precision):
    """The fused kernels run as persistent CTAs (one per SM walking patches / tiles); with more work items than SMs
    every barrier parity is exercised across iterations (3000 patches = ~20 patches and ~40 head tiles per SM; the small
    and odd sizes leave some CTAs with one item fewer or none; k = 128 has one row tile per patch, k = 512 two work items
    per patch).  Results must be bit-identical to one CTA per item."""
    from deepfit_b200 import DeepFit as D
    from deepfit_b200 import ops
    from oracle import deepfit_oracle as O
    g = torch.Generator().manual_seed(3)
    xy = torch.rand((B, 2, k), generator=g) * 1.4 - 0.7
    c = torch.rand((B, 3, 1), generator=g) - 0.5
    zz = 0.3 * (c[:, 0:1] * xy[:, 0:1] ** 2 + c[:, 1:2] * xy[:, 1:2] ** 2 + c[:, 2:3] * xy[:, 0:1] * xy[:, 1:2])
    patch = torch.cat([xy, zz], 1).cuda().contiguous()
    net = ops.WeightNetwork(D.fold_batchnorm(O.seeded_state_dict(0)), k, "sigmoid", precision, max_batch=max(B, 128))
    out = {}
    for name, head_p, chain_p in (("per_item", 0, 0), ("persistent", 1, 1)):
        dbg_switch(8, head_p, 1)
        dbg_switch(10, chain_p, 1)
        w, trans, t2, _ = net.forward(patch, want_trans2=True)
        net.status()
        out[name] = (w.clone(), trans.clone(), t2.clone())
    for a, b in zip(out["per_item"], out["persistent"]):
        assert torch.equal(a, b)
    # and against the fp32 torch restatement on a sample of patches from every wave
    sel = torch.arange(0, B, 149)
    wr, tr, t2r, _ = O.weight_network(O.seeded_state_dict(0), patch[sel].cpu())
    assert (out["persistent"][0][sel].cpu() - wr).abs().max() < 5e-4
    assert (out["persistent"][1][sel].cpu() - tr).abs().max() < 2e-4


@pytest.mark.parametrize("precision", ALL_PRECISIONS)
@pytest.mark.parametrize("k", [128, 384, 512])
def test_network_other_neighbourhood_sizes(k, precision):
    """k = 128 and 512 run the fused chain kernels (one tile / two 256-point work items per patch), k = 384 the generic
    kernels; all against the fp32 torch restatement."""
    from deepfit_b200 import DeepFit as D
    from deepfit_b200 import ops
    from oracle import deepfit_oracle as O
    g = torch.Generator().manual_seed(11)
    B = 40
    xy = torch.rand((B, 2, k), generator=g) * 1.4 - 0.7
    c = torch.rand((B, 3, 1), generator=g) - 0.5
    zz = 0.3 * (c[:, 0:1] * xy[:, 0:1] ** 2 + c[:, 1:2] * xy[:, 1:2] ** 2 + c[:, 2:3] * xy[:, 0:1] * xy[:, 1:2])
    patch = torch.cat([xy, zz], 1).contiguous()
    sd = O.seeded_state_dict(0)
    net = ops.WeightNetwork(D.fold_batchnorm(sd), k, "sigmoid", precision, max_batch=128)
    w, trans, t2, pts = net.forward(patch.cuda(), want_trans2=True, want_points=True)
    net.status()
    wr, tr, t2r, pr = O.weight_network(sd, patch)
    assert (trans.cpu() - tr).abs().max() < 2e-4
    assert (t2.cpu() - t2r).abs().max() < 2e-3
    assert (pts.cpu() - pr).abs().max() < 2e-4
    assert (w.cpu() - wr).abs().max() < 5e-4
