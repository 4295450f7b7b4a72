"""Drop-in for the reference's ``models/DeepFit.py`` public surface, backed by the sm_100a kernels.

Same names, argument meaning and error behaviour as the reference (sitzikbs/DeepFit):

* ``fit_Wjet(points, weights, order=2, compute_neighbor_normals=False)`` -> ``(beta, n_est, None)``
  (reference models/DeepFit.py:11-117)
* ``compute_principal_curvatures(beta)`` -> ``(curvatures, dirs)`` (models/DeepFit.py:444-474)
* ``DeepFit(k, num_points, use_point_stn, use_feat_stn, point_tuple, sym_op, arch, n_gaussians,
  jet_order, weight_mode, use_consistency)`` with ``load_state_dict`` accepting the reference's
  125-entry ``state_dict`` and ``forward(points[B,3,k])`` ->
  ``(normal, beta, weights, trans, trans2, neighbor_normals)`` (models/DeepFit.py:270-321)

Documented deviations: BatchNorm always runs in eval mode (folded into the
preceding layer); all B rows are solved (the reference leaves the last ``B mod 16`` rows at zero);
a failed Cholesky is retried with a deterministic ridge and flagged instead of random noise;
``weight_mode='softmax'``, ``arch='3dmfv'`` and ``point_tuple>1`` raise.
All tensors must live on a CUDA device: there is no CPU fallback.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib, ops

__all__ = ["fit_Wjet", "compute_principal_curvatures", "DeepFit", "fold_batchnorm"]


def fit_Wjet(points, weights, order=2, compute_neighbor_normals=False):
    """Fit an n-jet to weighted, PCA-aligned patches.  points [B,3,k], weights [B,k] (ones for the
    classic fit).  Returns (beta [B,M], n_est [B,3], neighbor_normals [B,k,3] or None) like the reference."""
    if order not in ops.JET_M:
        raise ValueError("Polynomial order unsupported, please use 1 or 2 ")
    squeeze = points.dim() == 2
    if squeeze:
        points = points.unsqueeze(0)
        weights = weights.reshape(1, -1)
    out = ops.wls_fit(points, weights, order, want_curv=False)
    beta, n_est = out["beta"], out["n_local"]
    neighbor_normals = ops.neighbor_normals(points, beta, order) if compute_neighbor_normals else None
    if squeeze:
        beta = beta.squeeze(0)
    return beta, n_est, neighbor_normals


def compute_principal_curvatures(beta):
    """Principal curvatures (ascending) and directions from jet coefficients, with the reference's
    upper-triangle ``symeig`` semantics.  Returns (curvatures [B,2], dirs [B,2,3]): ``dirs`` is symeig's eigenvector
    matrix (eigenvectors in columns) with a zero column appended, as in the reference (models/DeepFit.py:469-472);
    the sign of each eigenvector is library-defined there, here its largest-magnitude component is positive."""
    if beta.shape[1] < 5:
        raise ValueError("Can't compute curvatures for jet with order less than 2")
    return ops.principal_curvatures(beta.float(), want_dirs=True)


# (module path, kind, in, out) -- the parameter tree of the reference for arch='simple'
def _stn_spec(prefix, dim):
    return [(prefix + ".conv1", "conv", dim, 64), (prefix + ".conv2", "conv", 64, 128),
            (prefix + ".conv3", "conv", 128, 1024), (prefix + ".fc1", "fc", 1024, 512),
            (prefix + ".fc2", "fc", 512, 256), (prefix + ".fc3", "fc", 256, 4 if dim == 3 else dim * dim),
            (prefix + ".bn1", "bn", 0, 64), (prefix + ".bn2", "bn", 0, 128), (prefix + ".bn3", "bn", 0, 1024),
            (prefix + ".bn4", "bn", 0, 512), (prefix + ".bn5", "bn", 0, 256)]


def _param_spec(use_point_stn, use_feat_stn, k_out):
    spec = [("feat.pointfeat.conv1", "conv", 3, 64), ("feat.pointfeat.conv2", "conv", 64, 64),
            ("feat.pointfeat.bn1", "bn", 0, 64), ("feat.pointfeat.bn2", "bn", 0, 64)]
    if use_point_stn:
        spec += _stn_spec("feat.pointfeat.stn1", 3)
    if use_feat_stn:
        spec += _stn_spec("feat.pointfeat.stn2", 64)
    spec += [("feat.conv2", "conv", 64, 128), ("feat.conv3", "conv", 128, 1024),
             ("feat.bn2", "bn", 0, 128), ("feat.bn3", "bn", 0, 1024),
             ("conv1", "conv", 1088, 512), ("conv2", "conv", 512, 256), ("conv3", "conv", 256, 128),
             ("conv4", "conv", 128, k_out), ("bn1", "bn", 0, 512), ("bn2", "bn", 0, 256), ("bn3", "bn", 0, 128)]
    return spec


# C-ABI layer order (include/deepfit_b200.h enum dfb_layer_id) -> (conv/fc module path, bn path or None)
_LAYER_SOURCES = [
    ("feat.pointfeat.stn1.conv1", "feat.pointfeat.stn1.bn1"), ("feat.pointfeat.stn1.conv2", "feat.pointfeat.stn1.bn2"),
    ("feat.pointfeat.stn1.conv3", "feat.pointfeat.stn1.bn3"), ("feat.pointfeat.stn1.fc1", "feat.pointfeat.stn1.bn4"),
    ("feat.pointfeat.stn1.fc2", "feat.pointfeat.stn1.bn5"), ("feat.pointfeat.stn1.fc3", None),
    ("feat.pointfeat.conv1", "feat.pointfeat.bn1"), ("feat.pointfeat.conv2", "feat.pointfeat.bn2"),
    ("feat.pointfeat.stn2.conv1", "feat.pointfeat.stn2.bn1"), ("feat.pointfeat.stn2.conv2", "feat.pointfeat.stn2.bn2"),
    ("feat.pointfeat.stn2.conv3", "feat.pointfeat.stn2.bn3"), ("feat.pointfeat.stn2.fc1", "feat.pointfeat.stn2.bn4"),
    ("feat.pointfeat.stn2.fc2", "feat.pointfeat.stn2.bn5"), ("feat.pointfeat.stn2.fc3", None),
    ("feat.conv2", "feat.bn2"), ("feat.conv3", "feat.bn3"),
    ("conv1", "bn1"), ("conv2", "bn2"), ("conv3", "bn3"), ("conv4", None),
]


def fold_batchnorm(state_dict):
    """Fold every eval-mode BatchNorm1d (y = g (x - mu) / sqrt(var + 1e-5) + b) into the preceding
    Conv1d(kernel=1) / Linear.  Returns the list of (weight [out,in], bias [out]) fp32 CPU tensors
    in C-ABI layer order."""
    layers = []
    for src, bn in _LAYER_SOURCES:
        w = state_dict[src + ".weight"].detach().to("cpu", torch.float64)
        if w.dim() == 3:
            w = w.squeeze(-1)
        b = state_dict[src + ".bias"].detach().to("cpu", torch.float64)
        if bn is not None:
            s = state_dict[bn + ".weight"].detach().to("cpu", torch.float64) / torch.sqrt(
                state_dict[bn + ".running_var"].detach().to("cpu", torch.float64) + 1e-5)
            w = w * s[:, None]
            b = (b - state_dict[bn + ".running_mean"].detach().to("cpu", torch.float64)) * s \
                + state_dict[bn + ".bias"].detach().to("cpu", torch.float64)
        layers.append((w.float().contiguous(), b.float().contiguous()))
    return layers


class _Node(nn.Module):
    """Plain container so parameter names match the reference's nested modules."""


class DeepFit(nn.Module):
    def __init__(self, k=1, num_points=500, use_point_stn=False, use_feat_stn=False, point_tuple=1,
                 sym_op='max', arch=None, n_gaussians=5, jet_order=2, weight_mode="tanh",
                 use_consistency=False, precision="f16x3"):
        super().__init__()
        if arch == '3dmfv':
            raise NotImplementedError("arch='3dmfv' (Fisher-vector encoder) is outside the hot path; the "
                                      "pretrained model is arch='simple'")
        if point_tuple != 1:
            raise NotImplementedError("point_tuple > 1 is a training-time variant that is not supported")
        if sym_op != 'max':
            raise NotImplementedError("sym_op=%r: only 'max' is implemented (the reference's STNs hard-code "
                                      "MaxPool1d too)" % (sym_op,))
        if k != 1:
            raise NotImplementedError("k (weights per point) must be 1")
        if not (use_point_stn and use_feat_stn):
            raise NotImplementedError("only the pretrained topology (use_point_stn=1, use_feat_stn=True) is built")
        if weight_mode not in ("sigmoid", "tanh"):
            raise NotImplementedError(
                "weight_mode=%r: the reference's softmax branch normalises over the batch axis "
                "(F.softmax without dim, models/DeepFit.py:310) and is not reproduced" % (weight_mode,))
        if jet_order not in ops.JET_M:
            raise ValueError("Polynomial order unsupported, please use 1 or 2 ")
        self.k = k
        self.num_points = num_points
        self.point_tuple = point_tuple
        self.jet_order = jet_order
        self.weight_mode = weight_mode
        self.compute_neighbor_normals = use_consistency
        self.precision = precision
        for path, kind, cin, cout in _param_spec(use_point_stn, use_feat_stn, k):
            parent = self
            parts = path.split(".")
            for name in parts[:-1]:
                if not hasattr(parent, name):
                    parent.add_module(name, _Node())
                parent = getattr(parent, name)
            if kind == "conv":
                leaf = nn.Conv1d(cin, cout, 1)
            elif kind == "fc":
                leaf = nn.Linear(cin, cout)
            else:
                leaf = nn.BatchNorm1d(cout)
            parent.add_module(parts[-1], leaf)
        self._net = None
        self._net_key = None

    # -- packed-network cache ------------------------------------------------------------------
    def _state_key(self, device):
        return (str(device), tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers())))

    def _network(self, device, max_batch=0):
        """The packed device network for the current parameters (rebuilt when they change or when a
        larger internal chunk is requested)."""
        key = self._state_key(device)
        have = getattr(self, "_net_max_batch", 0)
        if self._net is None or self._net_key != key or int(max_batch) > have:
            if self._net is not None:
                self._net.close()
            want = max(int(max_batch), have if self._net_key == key else 0) or 1024
            self._net = ops.WeightNetwork(fold_batchnorm(self.state_dict()), self.num_points, self.weight_mode,
                                          self.precision, max_batch=want, device=device)
            self._net_key = key
            self._net_max_batch = want
        return self._net

    def forward(self, points):
        """points f32 [B,3,k] (PCA-aligned patches) ->
        (normal [B,3], beta [B,M], weights [B,k], trans [B,3,3], trans2 [B,64,64], None)."""
        if not points.is_cuda:
            raise ValueError("deepfit_b200.DeepFit runs on a CUDA device only (no CPU fallback); move the points "
                             "and the model to cuda")
        if points.shape[2] != self.num_points:
            raise ValueError("points per patch (%d) must equal num_points (%d): the reference pools with "
                             "MaxPool1d(num_points), models/DeepFit.py:336,395" % (points.shape[2], self.num_points))
        net = self._network(points.device)
        weights, trans, trans2, _ = net.forward(points.float(), want_trans2=True)
        out = ops.wls_fit(points.float(), weights, self.jet_order, trans=trans, want_curv=False)
        neighbor_normals = None
        if self.compute_neighbor_normals:   # use_consistency=True: normals of the jet at every (rotated) neighbour
            neighbor_normals = ops.neighbor_normals(points.float(), out["beta"], self.jet_order, trans=trans)
        net.status()           # raises on a pipeline time-out instead of returning garbage
        return out["n_local"], out["beta"], weights, trans, trans2, neighbor_normals
