"""Torch-facing wrappers of the C ABI: one function per stage, device tensors in, device tensors out.

PyTorch owns the memory and the stream; every computation happens in ``libdeepfit_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

JET_M = {1: 3, 2: 6, 3: 10, 4: 15}


def _p(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError("%s must be a CUDA tensor: deepfit_b200 has no CPU path" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


class Grid:
    """Search structure over a cloud (replaces ``scipy.spatial.cKDTree(points, 10)``,
    reference tutorial/tutorial_utils.py:16)."""

    def __init__(self, points: torch.Tensor):
        points = _chk(points, torch.float32, "points")
        if points.dim() != 2 or points.shape[1] != 3:
            raise ValueError("points must be [N,3]")
        self.n = int(points.shape[0])
        self.device = points.device
        self._h = C.c_void_p()
        lib = _lib.load()
        with torch.cuda.device(self.device):
            _lib.check(lib.dfb_grid_create(_p(points), self.n, _stream(self.device), C.byref(self._h)))

    def update(self, points: torch.Tensor):
        """Rebuild in place for another cloud of the same size (asynchronous, no allocation)."""
        points = _chk(points, torch.float32, "points")
        if tuple(points.shape) != (self.n, 3):
            raise ValueError("Grid.update needs a cloud of the same size [%d,3]" % self.n)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().dfb_grid_update(self._h, _p(points), _stream(self.device)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().dfb_grid_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001 - interpreter shutdown
            pass

    def morton_order(self, q_begin: int = 0, q_end: Optional[int] = None) -> torch.Tensor:
        """The point indices ``q_begin .. q_end`` in the grid's spatial (Morton) order, int32 [q_end - q_begin]: a query
        list in which consecutive queries share their neighbourhood cells."""
        q_end = self.n if q_end is None else int(q_end)
        out = torch.empty((max(q_end - q_begin, 0),), dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().dfb_grid_morton_order(self._h, q_begin, q_end, _p(out), _stream(self.device)))
        return out

    def query(self, k: int, q_begin: int = 0, q_count: Optional[int] = None, query: Optional[torch.Tensor] = None,
              return_dist: bool = False):
        """kNN of points ``q_begin .. q_begin+q_count`` (or of the int32 index tensor ``query``).
        Returns (idx i32 [Q,k], rad f64 [Q][, dist f64 [Q,k]])."""
        if query is not None:
            query = _chk(query, torch.int32, "query")
            q_count = int(query.numel())
        elif q_count is None:
            q_count = self.n - q_begin
        idx = torch.empty((q_count, k), dtype=torch.int32, device=self.device)
        rad = torch.empty((q_count,), dtype=torch.float64, device=self.device)
        dist = torch.empty((q_count, k), dtype=torch.float64, device=self.device) if return_dist else None
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().dfb_knn(self._h, _p(query), q_begin, q_count, k, _p(idx), _p(dist), _p(rad),
                                           _stream(self.device)))
        return (idx, rad, dist) if return_dist else (idx, rad)

    def ball_query(self, radius, q_begin: int = 0, q_count: Optional[int] = None, query: Optional[torch.Tensor] = None,
                   sort: bool = True):
        """Every point within ``radius`` (a float, or an f64 tensor with one radius per query) of points
        ``q_begin .. q_begin+q_count`` (or of the int32 index tensor ``query``): the set
        ``cKDTree.query_ball_point`` returns (reference dataset.py:289-291).  Returns
        (offsets i64 [Q+1], idx i32 [offsets[-1]]): the neighbours of query q are ``idx[offsets[q]:offsets[q+1]]``,
        ascending by point index when ``sort`` (the reference's order is the tree's traversal order)."""
        if query is not None:
            query = _chk(query, torch.int32, "query")
            q_count = int(query.numel())
        elif q_count is None:
            q_count = self.n - q_begin
        rad_t = None
        r = 0.0
        if isinstance(radius, torch.Tensor):
            rad_t = _chk(radius, torch.float64, "radius")
            if rad_t.numel() != q_count:
                raise ValueError("radius tensor must hold one radius per query")
        else:
            r = float(radius)
        lib = _lib.load()
        count = torch.empty((q_count,), dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.dfb_ball_count(self._h, _p(query), q_begin, q_count, _p(rad_t), r, _p(count), _stream(self.device)))
            offsets = torch.zeros((q_count + 1,), dtype=torch.int64, device=self.device)
            torch.cumsum(count, 0, out=offsets[1:])
            total = int(offsets[-1].item()) if q_count else 0
            idx = torch.empty((total,), dtype=torch.int32, device=self.device)
            if total:
                _lib.check(lib.dfb_ball_fill(self._h, _p(query), q_begin, q_count, _p(rad_t), r, _p(offsets), _p(idx),
                                             _stream(self.device)))
                if sort:   # canonical order: (query, index); not the hot path, so the library sort does it
                    seg = torch.repeat_interleave(torch.arange(q_count, device=self.device), count)
                    key = seg * int(self.n) + idx.long()
                    idx = (torch.sort(key)[0] - seg * int(self.n)).to(torch.int32)
        return offsets, idx


def patch_pca(points: torch.Tensor, idx: torch.Tensor, rad: Optional[torch.Tensor], q_begin: int = 0,
              query: Optional[torch.Tensor] = None, scale: bool = True, pca: bool = True,
              valid: Optional[torch.Tensor] = None):
    """(points[idx] - points[centre]) / rad, PCA-aligned.  Returns (patch f32 [Q,3,k], U f32 [Q,3,3]).
    ``valid`` (i32 [Q]): only the first ``valid[q]`` neighbours of a row are real; the other rows of the patch are
    zero padding outside the PCA (ball-query patches with fewer than k points)."""
    points = _chk(points, torch.float32, "points")
    idx = _chk(idx, torch.int32, "idx")
    q, k = int(idx.shape[0]), int(idx.shape[1])
    if rad is not None:
        rad = _chk(rad, torch.float64, "rad")
    if query is not None:
        query = _chk(query, torch.int32, "query")
    if valid is not None:
        valid = _chk(valid, torch.int32, "valid")
        if valid.numel() != q:
            raise ValueError("valid must hold one count per patch")
    patch = torch.empty((q, 3, k), dtype=torch.float32, device=points.device)
    U = torch.empty((q, 3, 3), dtype=torch.float32, device=points.device)
    flags = (1 if scale else 0) | (0 if pca else 2)
    with torch.cuda.device(points.device):
        _lib.check(_lib.load().dfb_patch_pca_ragged(_p(points), int(points.shape[0]), _p(idx), _p(valid), _p(query), q_begin, q, k,
                                                    _p(rad), flags, _p(patch), _p(U), _stream(points.device)))
    return patch, U


def wls_fit(patch: torch.Tensor, weights: Optional[torch.Tensor], order: int, trans: Optional[torch.Tensor] = None,
            U: Optional[torch.Tensor] = None, rad: Optional[torch.Tensor] = None, want_curv: bool = True,
            want_dirs: bool = False):
    """Fused weighted jet fit + normal + curvature.  Returns a dict of device tensors:
    beta [n,M], n_local [n,3], n_world [n,3] (if U or trans given), curv [n,2] f32 and
    curv_scaled [n,2] f64 (order >= 2), info [n] i32; with ``want_dirs`` also dirs [n,2,3] (the second return
    value of compute_principal_curvatures) and dirs_world [n,2,3] (rotated back by trans^T and U^T,
    test_c_est.py:162-168)."""
    patch = _chk(patch, torch.float32, "points")
    if patch.dim() != 3 or patch.shape[1] != 3:
        raise ValueError("points must be [B,3,k]")
    if order not in JET_M:
        raise ValueError("Polynomial order unsupported, please use 1 or 2 ")
    n, k = int(patch.shape[0]), int(patch.shape[2])
    dev = patch.device
    if weights is not None:
        weights = _chk(weights, torch.float32, "weights").reshape(n, k)
    if trans is not None:
        trans = _chk(trans, torch.float32, "trans")
    if U is not None:
        U = _chk(U, torch.float32, "U")
    if rad is not None:
        rad = _chk(rad, torch.float64, "rad")
    out = {
        "beta": torch.empty((n, JET_M[order]), dtype=torch.float32, device=dev),
        "n_local": torch.empty((n, 3), dtype=torch.float32, device=dev),
        "info": torch.empty((n,), dtype=torch.int32, device=dev),
    }
    if U is not None or trans is not None:
        out["n_world"] = torch.empty((n, 3), dtype=torch.float32, device=dev)
    if order > 1 and want_curv:
        out["curv"] = torch.empty((n, 2), dtype=torch.float32, device=dev)
        out["curv_scaled"] = torch.empty((n, 2), dtype=torch.float64, device=dev)
    if order > 1 and want_dirs:
        out["dirs"] = torch.empty((n, 2, 3), dtype=torch.float32, device=dev)
        out["dirs_world"] = torch.empty((n, 2, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().dfb_wls_fit(_p(patch), _p(weights), _p(trans), _p(U), _p(rad), n, k, order,
                                           _p(out["beta"]), _p(out["n_local"]), _p(out.get("n_world")),
                                           _p(out.get("curv")), _p(out.get("curv_scaled")), _p(out.get("dirs")),
                                           _p(out.get("dirs_world")), _p(out["info"]), _stream(dev)))
    return out


def neighbor_normals(patch: torch.Tensor, beta: torch.Tensor, order: int, trans: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Per-neighbour analytic normals of the fitted jet, [n,k,3] (reference models/DeepFit.py:90-115)."""
    patch = _chk(patch, torch.float32, "points")
    beta = _chk(beta, torch.float32, "beta")
    if trans is not None:
        trans = _chk(trans, torch.float32, "trans")
    n, k = int(patch.shape[0]), int(patch.shape[2])
    out = torch.empty((n, k, 3), dtype=torch.float32, device=patch.device)
    with torch.cuda.device(patch.device):
        _lib.check(_lib.load().dfb_neighbor_normals(_p(patch), _p(trans), _p(beta), n, k, order, _p(out), _stream(patch.device)))
    return out


def launch_count(reset: bool = False) -> int:
    """Number of CUDA kernels the library has launched in this process."""
    return int(_lib.load().dfb_launch_count(1 if reset else 0))


def principal_curvatures(beta: torch.Tensor, want_dirs: bool = False):
    """Curvatures f32 [n,2] (ascending); with ``want_dirs`` also the directions f32 [n,2,3] in the layout of the
    reference's compute_principal_curvatures (tutorial_utils.py:223-224)."""
    beta = _chk(beta, torch.float32, "beta")
    if beta.dim() != 2 or beta.shape[1] < 5:
        raise ValueError("Can't compute curvatures for jet with order less than 2")
    curv = torch.empty((beta.shape[0], 2), dtype=torch.float32, device=beta.device)
    dirs = torch.empty((beta.shape[0], 2, 3), dtype=torch.float32, device=beta.device) if want_dirs else None
    with torch.cuda.device(beta.device):
        _lib.check(_lib.load().dfb_principal_curvatures(_p(beta), int(beta.shape[0]), int(beta.shape[1]), _p(curv), _p(dirs),
                                                        _stream(beta.device)))
    return (curv, dirs) if want_dirs else curv


class Pipeline:
    """The fused path 1 -> 5 behind one C call (``dfb_pipeline_run``): the batch loop of the reference's
    tutorial/compute_normals.py:56-81.  ``net`` is a ``WeightNetwork`` (DeepFit mode) or None (classic)."""

    OUTPUTS = {"normals": (torch.float32, lambda k, m: (3,)), "curv": (torch.float64, lambda k, m: (2,)),
               "beta": (torch.float32, lambda k, m: (m,)), "weights": (torch.float32, lambda k, m: (k,)),
               "trans": (torch.float32, lambda k, m: (3, 3)), "trans2": (torch.float32, lambda k, m: (64, 64)),
               "U": (torch.float32, lambda k, m: (3, 3)), "rad": (torch.float64, lambda k, m: ()),
               "dirs_world": (torch.float32, lambda k, m: (2, 3)), "info": (torch.int32, lambda k, m: ())}

    def __init__(self, grid: Grid, net: Optional["WeightNetwork"], k: int, order: int, chunk: int = 0, keep_qstn: bool = False,
                 file_order: bool = False):
        """``file_order``: visit the queries in file order like the reference instead of the grid's Morton order (same
        results; the neighbour search is slower)."""
        if order not in JET_M:
            raise ValueError("Polynomial order unsupported, please use 1 or 2 ")
        self.grid, self.net = grid, net        # keep both alive: the handle borrows them
        self.k, self.order = int(k), int(order)
        self.device = grid.device
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().dfb_pipeline_create(grid._h, net._h if net is not None else None, self.k, self.order, int(chunk),
                                                       (_lib.DFB_PIPE_KEEP_QSTN if keep_qstn else 0) |
                                                       (_lib.DFB_PIPE_FILE_ORDER if file_order else 0), C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().dfb_pipeline_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def run(self, begin: int, end: int, want=("normals", "curv"), out: Optional[dict] = None) -> dict:
        """Asynchronous on the current stream.  ``want`` names the outputs to produce (see OUTPUTS); ``out`` may hold
        preallocated tensors for some of them (e.g. views into an all-gather buffer)."""
        q = int(end) - int(begin)
        res = {}
        o = _lib.DfbPipelineOutputs()
        for name in want:
            if name == "curv" and self.order < 2:
                continue
            dtype, shp = self.OUTPUTS[name]
            shape = (q,) + tuple(shp(self.k, JET_M[self.order]))
            t = out.get(name) if out else None
            if t is None:
                t = torch.empty(shape, dtype=dtype, device=self.device)
            else:
                t = _chk(t, dtype, name)
                if tuple(t.shape) != shape or not t.is_contiguous():
                    raise ValueError("%s must be a contiguous %s tensor of shape %s" % (name, dtype, shape))
            res[name] = t
            setattr(o, "d_" + name, t.data_ptr())
        if "normals" not in res:
            raise ValueError("the pipeline always produces normals")
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().dfb_pipeline_run(self._h, int(begin), int(end), C.byref(o), _stream(self.device)))
        return res


class WeightNetwork:
    """Device-resident, pre-packed weight network (stage 3).  ``layers`` is a list of
    (weight fp32 [out,in], bias fp32 [out]) CPU tensors in ``_lib.LAYER_NAMES`` order, BN folded."""

    def __init__(self, layers, k: int, weight_mode: str = "sigmoid", precision: str = "f16x3", max_batch: int = 0,
                 device=None):
        if weight_mode not in _lib.WEIGHT_MODE:
            raise NotImplementedError(
                "weight_mode %r: the reference's softmax branch is a batch-axis softmax (models/DeepFit.py:310) "
                "and is not reproduced" % weight_mode)
        precision_code = _lib.precision_code(precision)
        self.device = torch.device(device if device is not None else "cuda")
        self.k = int(k)
        desc = _lib.DfbModelDesc()
        desc.k = self.k
        desc.weight_mode = _lib.WEIGHT_MODE[weight_mode]
        desc.precision = precision_code
        desc.max_batch = int(max_batch)
        keep = []
        assert len(layers) == _lib.DFB_NUM_LAYERS
        for i, (w, b) in enumerate(layers):
            w = w.detach().to("cpu", torch.float32).contiguous()
            b = b.detach().to("cpu", torch.float32).contiguous()
            keep.append((w, b))
            desc.layers[i].h_weight = w.data_ptr()
            desc.layers[i].h_bias = b.data_ptr()
            desc.layers[i].out_features = int(w.shape[0])
            desc.layers[i].in_features = int(w.shape[1])
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().dfb_model_create(C.byref(desc), _stream(self.device), C.byref(self._h)))
        del keep

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            _lib.load().dfb_model_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    PROFILE_CLASSES = ["pointwise", "gemm_small", "gemm_max_128x1024", "fc", "head", "quat"]

    def profile(self, enable: bool):
        """Record per-kernel-class CUDA-event timings of subsequent forwards."""
        _lib.check(_lib.load().dfb_model_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        """Synchronise; returns {class: (milliseconds, launches)} accumulated since the last read."""
        n = len(self.PROFILE_CLASSES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().dfb_model_profile_read(self._h, ms, cnt, n, _stream(self.device)))
        return {name: (ms[i], cnt[i]) for i, name in enumerate(self.PROFILE_CLASSES)}

    def status(self):
        """Synchronise and raise if any tensor-core pipeline of this network timed out."""
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().dfb_model_status(self._h, _stream(self.device)))

    def forward(self, patch: torch.Tensor, want_trans2: bool = True, want_points: bool = False):
        """patch f32 [n,3,k] -> (weights [n,k], trans [n,3,3], trans2 [n,64,64] | None, points_rot | None)."""
        patch = _chk(patch, torch.float32, "points")
        n, k = int(patch.shape[0]), int(patch.shape[2])
        if k != self.k:
            raise ValueError("this network was built for %d points per patch (MaxPool1d(num_points), reference "
                             "models/DeepFit.py:336,395), got %d" % (self.k, k))
        dev = patch.device
        weights = torch.empty((n, k), dtype=torch.float32, device=dev)
        trans = torch.empty((n, 3, 3), dtype=torch.float32, device=dev)
        trans2 = torch.empty((n, 64, 64), dtype=torch.float32, device=dev) if want_trans2 else None
        pts = torch.empty((n, 3, k), dtype=torch.float32, device=dev) if want_points else None
        with torch.cuda.device(dev):
            _lib.check(_lib.load().dfb_model_forward(self._h, _p(patch), n, _p(weights), _p(trans), _p(trans2), _p(pts),
                                                     _stream(dev)))
        return weights, trans, trans2, pts
