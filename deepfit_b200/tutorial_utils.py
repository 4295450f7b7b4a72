"""Drop-in for the hot-path part of the reference's ``tutorial/tutorial_utils.py``.

``SinglePointCloudDataset(point_filename, points_per_patch)`` keeps the reference's attributes
(``points``, ``bbdiag``, ``kdtree``, ``points_per_patch``) and ``__len__`` / ``__getitem__`` contract
(tutorial/tutorial_utils.py:10-57) while the neighbour search and the centre/scale/PCA step run as
CUDA kernels; ``batch()`` is the device-resident bulk form the drop-in CLI uses instead of a
multi-process ``DataLoader``.  ``compute_principal_curvatures`` mirrors :196-226.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ops
from .DeepFit import compute_principal_curvatures  # noqa: F401  (same function, as in the reference)

__all__ = ["SinglePointCloudDataset", "compute_principal_curvatures", "load_xyz", "loadtxt_f32", "savetxt", "normalize_cloud"]


def loadtxt_f32(filename) -> np.ndarray:
    """``np.loadtxt(filename).astype('float32')`` (reference tutorial_utils.py:13) with all host threads."""
    import ctypes as C

    from . import _lib
    lib = _lib.load()
    rows, cols = C.c_int64(0), C.c_int32(0)
    path = os.fsencode(str(filename))
    _lib.check(lib.dfb_loadtxt_f32(path, None, 0, C.byref(rows), C.byref(cols)))
    out = np.empty((rows.value, cols.value), dtype=np.float32)
    _lib.check(lib.dfb_loadtxt_f32(path, out.ctypes.data, out.size, C.byref(rows), C.byref(cols)))
    return out


def savetxt(filename, array) -> None:
    """``np.savetxt(filename, array)`` with numpy's default format, byte for byte, with all host threads."""
    from . import _lib
    lib = _lib.load()
    a = np.ascontiguousarray(array)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    if a.dtype == np.float32:
        fn = lib.dfb_savetxt_f32
    else:
        a = np.ascontiguousarray(a, dtype=np.float64)
        fn = lib.dfb_savetxt_f64
    _lib.check(fn(os.fsencode(str(filename)), a.ctypes.data, a.shape[0], a.shape[1]))


def load_xyz(point_filename) -> np.ndarray:
    """Text ``.xyz`` (N x >=3 columns) -> float32 [N,3]; ``.npy`` is accepted as a binary side channel."""
    if str(point_filename).endswith(".npy"):
        return np.load(point_filename).astype("float32")[:, :3]
    return np.ascontiguousarray(loadtxt_f32(point_filename)[:, :3])


def normalize_cloud(points: np.ndarray):
    """Shrink the shape to the unit sphere exactly as the reference does (float32 numpy, same
    expression order, tutorial/tutorial_utils.py:14-15) so kNN inputs are bit-identical."""
    points = np.asarray(points, dtype=np.float32)
    bbdiag = float(np.linalg.norm(points.max(0) - points.min(0), 2))
    return (points - points.mean(0)) / (0.5 * bbdiag), bbdiag


class SinglePointCloudDataset:
    def __init__(self, point_filename, points_per_patch, device="cuda", normalize=True):
        self.points_per_patch = int(points_per_patch)
        if isinstance(point_filename, np.ndarray):
            pts = np.asarray(point_filename, dtype=np.float32)
        else:
            pts = load_xyz(point_filename)
        if normalize:
            self.points, self.bbdiag = normalize_cloud(pts)
        else:
            self.points = pts
            self.bbdiag = float(np.linalg.norm(pts.max(0) - pts.min(0), 2))
        self.points = np.ascontiguousarray(self.points, dtype=np.float32)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise ValueError("SinglePointCloudDataset needs a CUDA device: the neighbour search is a CUDA kernel")
        self.points_gpu = torch.from_numpy(self.points).to(self.device)
        self.kdtree = ops.Grid(self.points_gpu)  # plays the role of cKDTree(points, 10)
        self.pc_plot = None

    def __len__(self):
        return self.points.shape[0]

    def batch(self, start: int, count: int):
        """Patches of points ``start .. start+count`` as device tensors:
        (patch f32 [B,3,k], trans f32 [B,3,3], rad f64 [B])."""
        idx, rad = self.kdtree.query(self.points_per_patch, q_begin=start, q_count=count)
        patch, trans = ops.patch_pca(self.points_gpu, idx, rad, q_begin=start)
        return patch, trans, rad

    def batch_indices(self, indices: torch.Tensor):
        indices = indices.to(self.device, torch.int32)
        idx, rad = self.kdtree.query(self.points_per_patch, query=indices)
        patch, trans = ops.patch_pca(self.points_gpu, idx, rad, query=indices)
        return patch, trans, rad

    def __getitem__(self, index):
        patch, trans, rad = self.batch(int(index), 1)
        return patch[0].cpu(), trans[0].cpu(), float(rad[0].item())

    def batches(self, batch_size: int):
        n = len(self)
        for s in range(0, n, batch_size):
            yield self.batch(s, min(batch_size, n - s))
