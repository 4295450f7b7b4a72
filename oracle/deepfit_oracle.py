"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy / torch-CPU fp32) of DeepFit's inference hot path.

This module is the *checker* for the CUDA path in ``deepfit_b200/``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import it.  Product code never does; it fails loudly when its CUDA library is missing.

Pinning status
--------------
The reference ships no tests and no golden vectors.  The oracle is pinned by
running the *unmodified reference* in the authoring container (through ``oracle/shims.py``)
and (a) asserting equality in ``tests/test_oracle_vs_reference.py`` (skipped where
/root/reference is absent) and (b) committing the reference's own outputs as fixtures under
``tests/golden/`` (generator: ``oracle/make_golden.py``).  The pretrained blob
``trained_models/DeepFit/DeepFit.pth`` is missing from the reference checkout, so stage-3
parity is pinned on a *seeded synthetic state_dict* with the exact key set
(``seeded_state_dict``): parity on the real pretrained weights is UNPINNED until that blob exists.

Third-party arithmetic reached by the reference and restated here:
  * scipy ``cKDTree`` (requirements.txt:7 pins scipy==1.3.2): Euclidean kNN on float32
    coordinates promoted to fp64; restated as brute force over ``(fp64 d^2, index)``.
  * LAPACK ``gesdd`` via ``torch.svd`` (tutorial_utils.py:49): restated with the same call;
    its column signs are library-defined, parity is modulo per-column sign.
  * LAPACK ``potrf/potrs`` via ``torch.cholesky`` (models/DeepFit.py:136-138): same call
    through ``torch.linalg``.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------
# Stage 1a: cloud normalisation -- tutorial/tutorial_utils.py:13-15
# ----------------------------------------------------------------------------------------

def normalize_cloud(points32: np.ndarray):
    """float32 [N,3] -> unit-sphere-normalised float32 [N,3] and bbdiag (python float).

    tutorial_utils.py:14  bbdiag = float(||max - min||_2)           (float32 norm -> python float)
    tutorial_utils.py:15  p = (p - mean0(p)) / (0.5 * bbdiag)       (stays float32)
    """
    p = np.asarray(points32, dtype=np.float32)
    bbdiag = float(np.linalg.norm(p.max(0) - p.min(0), 2))
    return ((p - p.mean(0)) / (0.5 * bbdiag)).astype(np.float32), bbdiag


# ----------------------------------------------------------------------------------------
# Stage 1b: kNN -- tutorial/tutorial_utils.py:21 (scipy cKDTree.query), dataset.py:293
# ----------------------------------------------------------------------------------------

def knn_bruteforce(points32: np.ndarray, query_idx, k: int, chunk: int = 256):
    """Exact kNN, rank key (fp64 d^2 accumulated x->y->z, index).  Returns idx int64 [Q,k] and
    dist fp64 [Q,k] (sqrt of d^2), both ascending -- the canonicalised form of what
    ``cKDTree(points,10).query(points[i], k)`` returns (tutorial_utils.py:16,21)."""
    p = np.asarray(points32, dtype=np.float32).astype(np.float64)
    query_idx = np.asarray(query_idx, dtype=np.int64)
    n = p.shape[0]
    out_i = np.empty((len(query_idx), k), dtype=np.int64)
    out_d = np.empty((len(query_idx), k), dtype=np.float64)
    ar = np.arange(n, dtype=np.int64)
    for s in range(0, len(query_idx), chunk):
        q = p[query_idx[s:s + chunk]]
        dx = q[:, None, 0] - p[None, :, 0]
        d2 = dx * dx
        dy = q[:, None, 1] - p[None, :, 1]
        d2 = d2 + dy * dy
        dz = q[:, None, 2] - p[None, :, 2]
        d2 = d2 + dz * dz
        for r in range(q.shape[0]):
            row = d2[r]
            # candidates: everything <= the k-th smallest value, then exact lexicographic order
            kth = np.partition(row, k - 1)[k - 1]
            cand = ar[row <= kth]
            order = np.lexsort((cand, row[cand]))[:k]
            sel = cand[order]
            out_i[s + r] = sel
            out_d[s + r] = np.sqrt(row[sel])
    return out_i, out_d


def canonicalize_knn(dist: np.ndarray, idx: np.ndarray):
    """Re-order one query's (dist, idx) rows by (dist, idx): cKDTree's tie order is unspecified."""
    dist = np.asarray(dist)
    idx = np.asarray(idx)
    if dist.ndim == 1:
        o = np.lexsort((idx, dist))
        return dist[o], idx[o]
    d2 = np.empty_like(dist)
    i2 = np.empty_like(idx)
    for r in range(dist.shape[0]):
        o = np.lexsort((idx[r], dist[r]))
        d2[r] = dist[r][o]
        i2[r] = idx[r][o]
    return d2, i2


def knn_ckdtree(points32: np.ndarray, query_idx, k: int, workers: int = 1):
    """The reference's actual neighbour search: scipy cKDTree(points, leafsize=10).query(., k)
    (tutorial_utils.py:16,21).  Used as the CPU baseline and to pin ``knn_bruteforce``."""
    import scipy.spatial as spatial
    tree = spatial.cKDTree(points32, 10)
    d, i = tree.query(points32[np.asarray(query_idx)], k=k, workers=workers)
    return i, d


# ----------------------------------------------------------------------------------------
# Stage 2: patch extraction + PCA -- tutorial/tutorial_utils.py:22-30, 35-57
# ----------------------------------------------------------------------------------------

def extract_patch(points32: np.ndarray, center: int, idx: np.ndarray, rad: float) -> torch.Tensor:
    """tutorial_utils.py:23-27: (p[idx] - p[center]) / rad in float32 (rad enters as a python scalar,
    torch divides fp32 tensor by the scalar rounded to fp32)."""
    pts = torch.from_numpy(points32[np.asarray(idx, dtype=np.int64), :])
    pts = pts - torch.from_numpy(points32[center, :])
    return pts / float(rad)


def pca_align(patch: torch.Tensor):
    """tutorial_utils.py:46-57.  patch f32 [k,3] -> (aligned [k,3], U [3,3]).
    U = left singular vectors of (P-m)^T; aligned = (P-m)U - (-m U)  (== P U up to rounding)."""
    m = patch.mean(0)
    pc = patch - m
    U, _, _ = torch.linalg.svd(pc.t(), full_matrices=False)
    aligned = pc @ U
    aligned = aligned - ((-m) @ U)
    return aligned, U


def dataset_item(points32: np.ndarray, center: int, k: int, idx=None, dist=None):
    """SinglePointCloudDataset.__getitem__ (tutorial_utils.py:19-30): returns
    (patch f32 [3,k], U f32 [3,3], rad fp64)."""
    if idx is None:
        idx, dist = knn_bruteforce(points32, [center], k)
        idx, dist = idx[0], dist[0]
    rad = float(np.max(dist))
    aligned, U = pca_align(extract_patch(points32, center, idx, rad))
    return aligned.t().contiguous(), U, rad


def pcpnet_patch_features(normals32, curv32, center: int, idx, U: torch.Tensor | None, patch_radius_abs: float):
    """Ground-truth features PointcloudPatchDataset.__getitem__ returns next to a patch (dataset.py:338-349, 368-372):
    normal f32 [3] and neighbor_normals f32 [k,3] = the shape's normals of the centre / of the patch points, times the
    PCA frame ``U`` when use_pca; (max_curvature, min_curvature) = curv[center] * absolute patch radius."""
    out = {}
    if normals32 is not None:
        n = torch.from_numpy(normals32[center, :])
        nn = torch.from_numpy(normals32[np.asarray(idx, dtype=np.int64), :])
        out["normal"] = torch.matmul(n, U) if U is not None else n
        out["neighbor_normals"] = torch.matmul(nn, U) if U is not None else nn
    if curv32 is not None:
        c = torch.from_numpy(curv32[center, :]) * patch_radius_abs
        out["max_curvature"], out["min_curvature"] = c[0:1], c[1:2]
    return out


# ----------------------------------------------------------------------------------------
# Stage 3: the weight network -- models/DeepFit.py:157-441 (eval-mode BatchNorm)
# ----------------------------------------------------------------------------------------

def model_layout(jet_order: int = 3):
    """(prefix, kind, in, out) for every parametrised layer; kinds: conv (Conv1d k=1 [out,in,1]),
    fc (Linear [out,in]), bn (BatchNorm1d).  Mirrors the constructors at models/DeepFit.py:
    PointNetFeatures :160-177, PointNetEncoder :211-226, DeepFit :271-298, STN :324-350,
    QSTN :383-408, for arch='simple', use_point_stn=1, use_feat_stn=True, point_tuple=1."""
    L = []

    def stn(prefix, dim):
        L.extend([
            (prefix + ".conv1", "conv", dim, 64), (prefix + ".conv2", "conv", 64, 128),
            (prefix + ".conv3", "conv", 128, 1024),
            (prefix + ".fc1", "fc", 1024, 512), (prefix + ".fc2", "fc", 512, 256),
            (prefix + ".fc3", "fc", 256, 4 if dim == 3 else dim * dim),
            (prefix + ".bn1", "bn", 64, 64), (prefix + ".bn2", "bn", 128, 128),
            (prefix + ".bn3", "bn", 1024, 1024), (prefix + ".bn4", "bn", 512, 512),
            (prefix + ".bn5", "bn", 256, 256)])

    L.extend([("feat.pointfeat.conv1", "conv", 3, 64), ("feat.pointfeat.conv2", "conv", 64, 64),
              ("feat.pointfeat.bn1", "bn", 64, 64), ("feat.pointfeat.bn2", "bn", 64, 64)])
    stn("feat.pointfeat.stn1", 3)
    stn("feat.pointfeat.stn2", 64)
    L.extend([("feat.conv2", "conv", 64, 128), ("feat.conv3", "conv", 128, 1024),
              ("feat.bn2", "bn", 128, 128), ("feat.bn3", "bn", 1024, 1024)])
    L.extend([("conv1", "conv", 1088, 512), ("conv2", "conv", 512, 256), ("conv3", "conv", 256, 128),
              ("conv4", "conv", 128, 1),
              ("bn1", "bn", 512, 512), ("bn2", "bn", 256, 256), ("bn3", "bn", 128, 128)])
    return L


def seeded_state_dict(seed: int = 0, head_gain: float = 2.0) -> "OrderedDict[str, torch.Tensor]":
    """A reproducible stand-in for the missing DeepFit.pth: same 125 keys/shapes.  Conv/FC weights
    are He-uniform (bound sqrt(6/in)) so per-point signal survives the ReLU stack, biases
    U(-0.1, 0.1); BN affine and running statistics are randomised so eval-mode BN is non-trivial;
    conv4 is calibrated (see below) so the predicted weights spread over (0.01, 1.01) instead of
    sitting at 0.51, and the last STN layers are shrunk so the predicted transforms stay within
    O(0.1) of identity, as trained ones do."""
    g = torch.Generator().manual_seed(int(seed))
    sd = OrderedDict()

    def U(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    for prefix, kind, cin, cout in model_layout():
        if kind in ("conv", "fc"):
            b = math.sqrt(6.0 / cin)
            sd[prefix + ".weight"] = U((cout, cin, 1) if kind == "conv" else (cout, cin), b)
            sd[prefix + ".bias"] = U((cout,), 0.1)
        else:
            sd[prefix + ".weight"] = 1.0 + U((cout,), 0.25)
            sd[prefix + ".bias"] = U((cout,), 0.1)
            sd[prefix + ".running_mean"] = U((cout,), 0.1)
            sd[prefix + ".running_var"] = 1.0 + U((cout,), 0.5)
            sd[prefix + ".num_batches_tracked"] = torch.tensor(100, dtype=torch.long)
    for p, shrink in (("feat.pointfeat.stn1.fc3", 0.02), ("feat.pointfeat.stn2.fc3", 0.01)):
        sd[p + ".weight"] = sd[p + ".weight"] * shrink
        sd[p + ".bias"] = sd[p + ".bias"] * shrink
    # let the per-point branch of the 1088-wide head input matter next to the 1024 global channels
    sd["conv1.weight"][:, 1024:, :] *= 4.0
    # calibrate the last layer in float64 on a deterministic batch of synthetic paraboloid patches so
    # that logits have mean 0 / std ``head_gain`` (weights then cover most of (0.01, 1.01))
    k = 256
    xy = torch.rand((8, 2, k), generator=g, dtype=torch.float64) * 1.4 - 0.7
    c = torch.rand((8, 3, 1), generator=g, dtype=torch.float64) - 0.5
    zz = 0.3 * (c[:, 0:1] * xy[:, 0:1] ** 2 + c[:, 1:2] * xy[:, 1:2] ** 2 + c[:, 2:3] * xy[:, 0:1] * xy[:, 1:2])
    zz = zz + 0.02 * (torch.rand((8, 1, k), generator=g, dtype=torch.float64) - 0.5)
    sd64 = OrderedDict((n, t.double() if t.is_floating_point() else t) for n, t in sd.items())
    logits = {}
    weight_network(sd64, torch.cat([xy, zz], 1), stage_hook=lambda n, t: logits.__setitem__(n, t))
    a = logits["logit"]
    scale = head_gain / float(a.std())
    sd["conv4.weight"] = (sd64["conv4.weight"] * scale).float()
    sd["conv4.bias"] = ((sd64["conv4.bias"] - a.mean()) * scale).float()
    return sd


def _bn(sd, prefix, x):
    """Eval-mode BatchNorm1d, default eps=1e-5 (constructors models/DeepFit.py:169-170 etc.)."""
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], training=False, eps=1e-5)


def _conv(sd, prefix, x):
    return F.conv1d(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def _fc(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def quat_to_rotmat(q: torch.Tensor) -> torch.Tensor:
    """utils/normal_estimation_utils.py:365-390: s = 2/|q|^2, h = q q^T."""
    s = 2.0 / (q * q).sum(1)
    h = q.unsqueeze(2) * q.unsqueeze(1)
    R = q.new_empty(q.shape[0], 3, 3)
    R[:, 0, 0] = 1 - (h[:, 2, 2] + h[:, 3, 3]) * s
    R[:, 0, 1] = (h[:, 1, 2] - h[:, 3, 0]) * s
    R[:, 0, 2] = (h[:, 1, 3] + h[:, 2, 0]) * s
    R[:, 1, 0] = (h[:, 1, 2] + h[:, 3, 0]) * s
    R[:, 1, 1] = 1 - (h[:, 1, 1] + h[:, 3, 3]) * s
    R[:, 1, 2] = (h[:, 2, 3] - h[:, 1, 0]) * s
    R[:, 2, 0] = (h[:, 1, 3] - h[:, 2, 0]) * s
    R[:, 2, 1] = (h[:, 2, 3] + h[:, 1, 0]) * s
    R[:, 2, 2] = 1 - (h[:, 1, 1] + h[:, 2, 2]) * s
    return R


def _stn_trunk(sd, prefix, x):
    """Shared by QSTN.forward (models/DeepFit.py:410-432) and STN.forward (:353-375):
    three conv+BN+ReLU, max over all k points, two fc+BN+ReLU, final fc."""
    x = F.relu(_bn(sd, prefix + ".bn1", _conv(sd, prefix + ".conv1", x)))
    x = F.relu(_bn(sd, prefix + ".bn2", _conv(sd, prefix + ".conv2", x)))
    x = F.relu(_bn(sd, prefix + ".bn3", _conv(sd, prefix + ".conv3", x)))
    g = x.max(dim=2)[0]                          # MaxPool1d(num_points) with num_points == k
    g = F.relu(_bn(sd, prefix + ".bn4", _fc(sd, prefix + ".fc1", g)))
    g = F.relu(_bn(sd, prefix + ".bn5", _fc(sd, prefix + ".fc2", g)))
    return _fc(sd, prefix + ".fc3", g)


def weight_network(sd, points: torch.Tensor, weight_mode: str = "sigmoid", stage_hook=None):
    """DeepFit.forward up to the weights (models/DeepFit.py:301-316) for the pretrained config
    (use_point_stn=1, use_feat_stn=True, arch='simple', point_tuple=1), eval-mode BN.

    points f32 [B,3,k] -> (weights [B,k], trans [B,3,3], trans2 [B,64,64], rotated points [B,3,k]).
    ``stage_hook(name, tensor)`` if given receives intermediate activations."""
    hook = stage_hook or (lambda n, t: None)
    B = points.shape[0]
    # QSTN (:410-441): quaternion + identity -> rotation
    q = _stn_trunk(sd, "feat.pointfeat.stn1", points)
    q = q + q.new_tensor([1.0, 0.0, 0.0, 0.0])
    trans = quat_to_rotmat(q)
    hook("trans", trans)
    # :188-190 row-vector rotation  P' = P R
    pts = torch.bmm(points.transpose(2, 1), trans).transpose(2, 1).contiguous()
    # :196-197
    x = F.relu(_bn(sd, "feat.pointfeat.bn1", _conv(sd, "feat.pointfeat.conv1", pts)))
    x = F.relu(_bn(sd, "feat.pointfeat.bn2", _conv(sd, "feat.pointfeat.conv2", x)))
    hook("x64", x)
    # feature STN (:201-204, :353-380)
    t2 = _stn_trunk(sd, "feat.pointfeat.stn2", x)
    t2 = (t2 + torch.eye(64, dtype=t2.dtype).reshape(1, 4096)).reshape(B, 64, 64)
    hook("trans2", t2)
    pointfeat = torch.bmm(x.transpose(2, 1), t2).transpose(2, 1)
    hook("pointfeat", pointfeat)
    # PointNetEncoder.forward (:229-237): conv3+BN has NO ReLU before the max
    y = F.relu(_bn(sd, "feat.bn2", _conv(sd, "feat.conv2", pointfeat)))
    y = _bn(sd, "feat.bn3", _conv(sd, "feat.conv3", y))
    gfeat = y.max(dim=2, keepdim=True)[0]
    hook("global", gfeat)
    z = torch.cat([gfeat.expand(-1, -1, points.shape[2]), pointfeat], 1)
    # DeepFit.forward head (:304-316)
    z = F.relu(_bn(sd, "bn1", _conv(sd, "conv1", z)))
    z = F.relu(_bn(sd, "bn2", _conv(sd, "conv2", z)))
    z = F.relu(_bn(sd, "bn3", _conv(sd, "conv3", z)))
    a = _conv(sd, "conv4", z).squeeze(1)
    hook("logit", a)
    if weight_mode == "sigmoid":
        w = 0.01 + torch.sigmoid(a)                       # :316
    elif weight_mode == "tanh":
        w = (0.01 + 1.0 + torch.tanh(a)) / 2.0            # :313-314
    else:
        raise NotImplementedError("weight_mode %r (softmax is batch-axis softmax in the reference, :310)" % weight_mode)
    return w, trans, t2, pts


# ----------------------------------------------------------------------------------------
# Stage 4: weighted jet fit -- models/DeepFit.py:11-154
# ----------------------------------------------------------------------------------------

# column monomials (power of x, power of y) in the reference's column order, DeepFit.py:51-76
JET_COLUMNS = {
    1: [(1, 0), (0, 1), (0, 0)],
    2: [(1, 0), (0, 1), (2, 0), (0, 2), (1, 1), (0, 0)],
    3: [(1, 0), (0, 1), (2, 0), (0, 2), (1, 1), (3, 0), (0, 3), (2, 1), (1, 2), (0, 0)],
    4: [(1, 0), (0, 1), (2, 0), (0, 2), (1, 1), (3, 0), (0, 3), (2, 1), (1, 2),
        (4, 0), (0, 4), (3, 1), (1, 3), (2, 2), (0, 0)],
}


def jet_design_matrix(x: torch.Tensor, y: torch.Tensor, order: int) -> torch.Tensor:
    cols = [x.pow(a) * y.pow(b) if (a or b) else torch.ones_like(x) for a, b in JET_COLUMNS[order]]
    return torch.stack(cols, dim=-1)


def fit_wjet(points: torch.Tensor, weights: torch.Tensor, order: int = 2, return_system: bool = False):
    """fit_Wjet (models/DeepFit.py:11-117) with every row solved (the reference leaves rows
    >= 16*floor(B/16) at zero, :129-133 -- a documented deviation).

    points [B,3,k], weights [B,k] -> beta [B,M], n_est [B,3] (+ XtX, XtY, ok if return_system).
    Rows whose fp32 Cholesky fails are returned with ok=False and beta from a ridge retry
    (the reference retries with *random* noise, :139-146, so those rows are unpinned)."""
    if order not in JET_COLUMNS:
        raise ValueError("Polynomial order unsupported, please use 1 or 2 ")
    B, _, k = points.shape
    x, y, z = points[:, 0, :], points[:, 1, :], points[:, 2, :]
    # :37-39 zero-weight guard
    valid = (weights > 1e-3).sum(dim=1, keepdim=True)
    w = torch.where(valid > 18, weights, torch.ones_like(weights))
    if order > 1:
        # :41-49 CGAL-style preconditioning
        h = (x.abs().mean(1) + y.abs().mean(1)) / 2
        h = torch.where(h.abs() < 1e-4, torch.full_like(h, 0.1), h)
        x = x / h[:, None]
        y = y / h[:, None]
    A = jet_design_matrix(x, y, order)                          # [B,k,M]
    At = A.transpose(1, 2)
    XtX = At @ (w[:, :, None] * A)                              # :80
    XtY = At @ (w * z)[:, :, None]                              # :81
    L, info = torch.linalg.cholesky_ex(XtX)
    ok = info == 0
    beta = torch.zeros_like(XtY)
    if ok.any():
        beta[ok] = torch.cholesky_solve(XtY[ok], L[ok])
    if (~ok).any():
        bad = ~ok
        ridge = XtX[bad] + 0.01 * torch.diag_embed(torch.diagonal(XtX[bad], dim1=1, dim2=2))
        beta[bad] = torch.linalg.solve(ridge, XtY[bad])
    beta = beta.squeeze(-1)
    if order > 1:
        # :85-86 D_inv
        deg = torch.tensor([a + b for a, b in JET_COLUMNS[order]], dtype=beta.dtype)
        beta = beta / h[:, None].pow(deg[None, :])
    n = F.normalize(torch.cat([-beta[:, 0:2], torch.ones(B, 1, dtype=beta.dtype)], dim=1), p=2, dim=1)   # :88
    if return_system:
        return beta, n, XtX, XtY.squeeze(-1), ok
    return beta, n


def neighbor_normals(points: torch.Tensor, beta: torch.Tensor, order: int) -> torch.Tensor:
    """fit_Wjet(..., compute_neighbor_normals=True), models/DeepFit.py:90-115: normals of the jet at every
    neighbour.  The reference evaluates the gradient at the PRECONDITIONED x/h, y/h (:48-49) with the
    un-preconditioned beta (:85-86); restated as is.  points [B,3,k], beta [B,M] -> [B,k,3]."""
    x, y = points[:, 0, :], points[:, 1, :]
    if order > 1:
        h = (x.abs().mean(1) + y.abs().mean(1)) / 2
        h = torch.where(h.abs() < 1e-4, torch.full_like(h, 0.1), h)
        x, y = x / h[:, None], y / h[:, None]
    b = [beta[:, i:i + 1] for i in range(beta.shape[1])]
    gx, gy = b[0].expand_as(x).clone(), b[1].expand_as(x).clone()
    if order >= 2:
        gx = gx + 2 * b[2] * x + b[4] * y
        gy = gy + 2 * b[3] * y + b[4] * x
    if order >= 3:
        gx = gx + 3 * b[5] * x * x + 2 * b[7] * x * y + b[8] * y * y
        gy = gy + 3 * b[6] * y * y + b[7] * x * x + 2 * b[8] * x * y
    if order >= 4:
        gx = gx + 4 * b[9] * x ** 3 + 3 * b[11] * x * x * y + b[12] * y ** 3 + 2 * b[13] * y * y * x
        gy = gy + 4 * b[10] * y ** 3 + b[11] * x ** 3 + 3 * b[12] * x * y * y + 2 * b[13] * y * x * x
    return F.normalize(torch.stack([-gx, -gy, torch.ones_like(gx)], -1), p=2, dim=-1)


# ----------------------------------------------------------------------------------------
# Stage 5: principal curvatures -- tutorial/tutorial_utils.py:196-226 (= models/DeepFit.py:444-474)
# ----------------------------------------------------------------------------------------

def principal_curvatures(beta: torch.Tensor) -> torch.Tensor:
    """Eigenvalues (ascending) of the symmetric matrix built from the UPPER triangle of
    M = -I^{-1} II (torch.symeig default upper=True, tutorial_utils.py:221-223).  [B,2]."""
    if beta.shape[1] < 5:
        raise ValueError("Can't compute curvatures for jet with order less than 2")
    b0, b1, b2, b3, b4 = (beta[:, i] for i in range(5))
    E = 1 + b0 * b0
    G = 1 + b1 * b1
    Fm = b1 * b0
    nrm = torch.sqrt(b0 * b0 + b1 * b1 + 1)
    e = 2 * b2 / nrm
    f = b4 / nrm
    g = 2 * b3 / nrm
    I = torch.stack([torch.stack([E, Fm], -1), torch.stack([Fm, G], -1)], -2)
    II = torch.stack([torch.stack([e, f], -1), torch.stack([f, g], -1)], -2)
    M = -torch.bmm(torch.inverse(I), II)
    return torch.linalg.eigvalsh(M, UPLO="U")


def principal_curvatures_closed_form(beta: torch.Tensor) -> torch.Tensor:
    """Same quantity by the 2x2 closed form on (M00, M01, M11) -- what the CUDA epilogue evaluates."""
    b = beta.double()
    b0, b1, b2, b3, b4 = (b[:, i] for i in range(5))
    E = 1 + b0 * b0
    G = 1 + b1 * b1
    Fm = b1 * b0
    nrm = torch.sqrt(b0 * b0 + b1 * b1 + 1)
    e, f, g = 2 * b2 / nrm, b4 / nrm, 2 * b3 / nrm
    det = E * G - Fm * Fm
    m00 = -(G * e - Fm * f) / det
    m01 = -(G * f - Fm * g) / det
    m11 = -(-Fm * f + E * g) / det
    mid = 0.5 * (m00 + m11)
    rad = torch.sqrt(0.25 * (m00 - m11) ** 2 + m01 * m01)
    return torch.stack([mid - rad, mid + rad], -1).to(beta.dtype)


# ----------------------------------------------------------------------------------------
# End to end -- tutorial/compute_normals.py:56-81 with the test_n_est.py:154-161 semantics
# ----------------------------------------------------------------------------------------

def deepfit_forward(sd, points: torch.Tensor, order: int = 3, weight_mode: str = "sigmoid"):
    """DeepFit.forward (models/DeepFit.py:301-321): (normal, beta, weights, trans, trans2, None)."""
    w, trans, t2, pts = weight_network(sd, points, weight_mode)
    beta, n = fit_wjet(pts, w, order)
    return n, beta, w, trans, t2, None


def world_normals(n_local: torch.Tensor, U: torch.Tensor, trans: torch.Tensor | None = None):
    """Row-vector back-rotation: undo QSTN (test_n_est.py:154-157) then PCA
    (compute_normals.py:67 / test_n_est.py:161):  n_world = (n R^T) U^T."""
    n = n_local
    if trans is not None:
        n = torch.bmm(n.unsqueeze(1), trans.transpose(2, 1)).squeeze(1)
    return torch.bmm(n.unsqueeze(1), U.transpose(2, 1)).squeeze(1)


def compute_normals(points32_normalized: np.ndarray, query_idx, k: int, order: int, mode: str = "classic",
                    sd=None, weight_mode: str = "sigmoid", use_ckdtree: bool = True, batch: int = 256):
    """The whole hot path on CPU for a list of query indices of an already-normalised cloud.
    Returns dict(normals [Q,3] f32, curv [Q,2] f64 (None for order 1), rad [Q] f64, idx [Q,k])."""
    query_idx = np.asarray(query_idx, dtype=np.int64)
    if use_ckdtree:
        idx, dist = knn_ckdtree(points32_normalized, query_idx, k)
        dist, idx = canonicalize_knn(dist, idx)
    else:
        idx, dist = knn_bruteforce(points32_normalized, query_idx, k)
    normals, curvs, rads = [], [], []
    for s in range(0, len(query_idx), batch):
        items = [dataset_item(points32_normalized, int(c), k, idx[s + j], dist[s + j])
                 for j, c in enumerate(query_idx[s:s + batch])]
        P = torch.stack([it[0] for it in items])
        U = torch.stack([it[1] for it in items])
        rad = torch.tensor([it[2] for it in items], dtype=torch.float64)
        with torch.no_grad():
            if mode == "DeepFit":
                n, beta, w, trans, t2, _ = deepfit_forward(sd, P, order, weight_mode)
                nw = world_normals(n, U, trans)
            else:
                beta, n = fit_wjet(P, torch.ones_like(P[:, 0]), order)
                nw = world_normals(n, U)
            normals.append(nw)
            if order > 1:
                curvs.append(principal_curvatures(beta).double() / rad[:, None])   # compute_normals.py:79
        rads.append(rad)
    return dict(normals=torch.cat(normals).numpy(),
                curv=torch.cat(curvs).numpy() if curvs else None,
                rad=torch.cat(rads).numpy(), idx=idx)


# ----------------------------------------------------------------------------------------
# Parity metrics (tolerances from BASELINE.json north_star; formulas from evaluate.py:132-157,
# evaluate_curvatures.py:142)
# ----------------------------------------------------------------------------------------

def unoriented_angle_deg(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    cr = np.linalg.norm(np.cross(a, b), axis=-1)
    dt = np.abs(np.sum(a * b, axis=-1))
    return np.degrees(np.arctan2(cr, dt))


def rectified_rel_err(x: np.ndarray, ref: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return np.abs(x - ref) / np.maximum(np.abs(ref), 1.0)
