"""Whole-cloud hot path: kNN -> centre/scale/PCA -> (weight network) -> jet fit -> normal + curvature.

This is the loop of the reference's ``tutorial/compute_normals.py:56-81`` without the DataLoader: the
cloud and the grid stay resident in HBM, queries are processed in chunks of point indices, and no
intermediate leaves the device.  The multi-GPU contract (contiguous point-index shards of the replicated
cloud, outputs written in place into the all-gather buffers) lives in ``deepfit_b200.sharding``.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .DeepFit import DeepFit, fold_batchnorm
from .sharding import GatherBuffers, gather_outputs, shard_capacity, shard_range  # noqa: F401  (re-exported)


class NormalEstimator:
    """Device-resident estimator for one cloud.

    points : CUDA float32 [N,3] (already normalised the way the caller wants -- the tutorial dataset
             normalises to the unit sphere, the PCPNet loader does not)
    mode   : 'classic' (unweighted jet, compute_normals.py:72) or 'DeepFit'
    model  : a ``deepfit_b200.DeepFit.DeepFit`` (DeepFit mode)
    cancel_qstn : undo the QSTN rotation on the normals like test_n_est.py:154-157 (True) or
                  reproduce the tutorial script's omission (False, ``--ref-compat``)
    grid   : an existing ``ops.Grid`` over ``points`` to reuse (e.g. ``SinglePointCloudDataset.kdtree``)
             instead of building a second copy of the cloud and its table
    """

    def __init__(self, points: torch.Tensor, k: int, jet_order: int = 3, mode: str = "classic",
                 model: Optional[DeepFit] = None, chunk: int = 0, cancel_qstn: bool = True,
                 grid: Optional[ops.Grid] = None, file_order: bool = False):
        if mode not in ("classic", "DeepFit"):
            raise ValueError("mode must be 'classic' or 'DeepFit'")
        if mode == "DeepFit" and model is None:
            raise ValueError("DeepFit mode needs a model")
        self.points = points.contiguous()
        self.n = int(points.shape[0])
        self.k = int(k)
        self.order = int(model.jet_order if mode == "DeepFit" else jet_order)
        self.mode = mode
        self.model = model
        self.cancel_qstn = cancel_qstn
        self.file_order = file_order    # visit queries in file order (the reference's) instead of Morton order; same results
        if grid is not None and grid.n != self.n:
            raise ValueError("grid was built over %d points, the cloud has %d" % (grid.n, self.n))
        self.grid = grid if grid is not None else ops.Grid(self.points)
        self.chunk = int(chunk) if chunk else (16384 if mode == "DeepFit" else 262144)

    @property
    def net(self):
        """The packed network for the model's CURRENT parameters (resolved on every use: the model rebuilds it after
        ``load_state_dict`` / ``.to()`` or when another user asked for a larger internal chunk)."""
        if self.mode != "DeepFit":
            return None
        return self.model._network(self.points.device, max_batch=self.chunk)

    def rebuild_grid(self):
        """Stage-1 build for the current ``self.points`` (same size): in place, asynchronous."""
        self.grid.update(self.points)

    def _pipeline(self) -> ops.Pipeline:
        """The C pipeline for the model's current packed network (rebuilt when the network handle changes)."""
        net = self.net
        key = (id(net), net._h.value if net is not None else 0)
        if getattr(self, "_pipe", None) is None or self._pipe_key != key:
            if getattr(self, "_pipe", None) is not None:
                self._pipe.close()
            self._pipe = ops.Pipeline(self.grid, net, self.k, self.order, self.chunk, keep_qstn=not self.cancel_qstn,
                                      file_order=self.file_order)
            self._pipe_key = key
        return self._pipe

    _EXTRA_NAMES = ("beta", "info", "rad", "U", "weights", "trans", "trans2", "dirs_world")

    def run(self, begin: int = 0, end: Optional[int] = None, out_normals: Optional[torch.Tensor] = None,
            out_curv: Optional[torch.Tensor] = None, extras: Optional[dict] = None, check: bool = True,
            want_extras: Optional[tuple] = None):
        """Normals f32 [Q,3] and curvatures f64 [Q,2] (k1<=k2, divided by the patch radius) for point
        indices [begin, end), through ``dfb_pipeline_run`` (one C call: chunk loop, workspace and the two-stream
        schedule live in the library).  ``extras`` (a dict) receives whole-range tensors ``beta / info / rad / U``
        (+ ``weights / trans / trans2`` in DeepFit mode, ``dirs_world`` for order >= 2) when given (config-3 style
        full outputs; ``want_extras`` narrows the list -- trans2 alone is 16 KB per point); a key already present with a
        tensor is used as the output buffer.  ``check`` synchronises
        once at the end and raises if any tensor-core pipeline of the network reported a time-out
        (``dfb_model_status``)."""
        end = self.n if end is None else int(end)
        pipe = self._pipeline()
        want = ["normals"] + (["curv"] if self.order > 1 else [])
        out = {}
        if out_normals is not None:
            out["normals"] = out_normals
        if out_curv is not None:
            out["curv"] = out_curv
        if extras is not None:
            for name in (self._EXTRA_NAMES if want_extras is None else want_extras):
                if name not in self._EXTRA_NAMES:
                    raise ValueError("unknown output %r" % (name,))
                if name in ("weights", "trans", "trans2") and self.mode != "DeepFit":
                    continue
                if name == "dirs_world" and self.order < 2:
                    continue
                want.append(name)
                if isinstance(extras.get(name), torch.Tensor):
                    out[name] = extras[name]
        res = pipe.run(begin, end, want, out)
        if extras is not None:
            for name in want[1:]:
                if name != "curv":
                    extras[name] = res[name]
        if self.net is not None and check:
            self.net.status()       # a tensor-core pipeline time-out must never end as silently wrong files
        return res["normals"], res.get("curv")


def model_from_state_dict(state_dict, num_points: int, jet_order: int = 3, weight_mode: str = "sigmoid",
                          precision: str = "f16x3", device="cuda") -> DeepFit:
    m = DeepFit(k=1, num_points=num_points, use_point_stn=True, use_feat_stn=True, point_tuple=1, sym_op="max",
                arch="simple", jet_order=jet_order, weight_mode=weight_mode, precision=precision)
    m.load_state_dict(state_dict)
    m.to(device)
    m.eval()
    return m


__all__ = ["NormalEstimator", "GatherBuffers", "shard_range", "shard_capacity", "gather_outputs", "model_from_state_dict",
           "fold_batchnorm"]
