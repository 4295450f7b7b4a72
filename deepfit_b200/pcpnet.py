"""PCPNet-style multi-shape inference with full outputs (BASELINE config 3).

Mirrors the k-nearest-neighbour path of the reference's ``PointcloudPatchDataset`` (dataset.py:173-442:
raw un-normalised points, ``center='point'``, patch scaled by the k-th neighbour distance, optional PCA,
optional sparse ``.pidx`` centres, sequential order) and the export loop of ``test_c_est.py:104-218``
(per shape ``.normals [P,3] .curv [P,2] .weights [P,k] .beta [P,M] .trans [P,9]`` via ``np.savetxt``).
``neighbor_search_method='r'`` (dataset.py:289-311, fixed-radius balls of ``patch_radius`` x bounding-box
diagonal, random subsample when the ball holds more than ``points_per_patch`` points, zero padding when fewer) is
supported for one scale; the ball is the exact set ``cKDTree.query_ball_point`` returns, but its elements are
taken in ascending index order where the reference takes the tree's traversal order, so the seeded
``RandomState`` subsample picks different members of the same ball (documented deviation: the traversal
order of scipy's tree cannot be reproduced).  Ground-truth patch features ('normal', 'neighbor_normals',
'max_curvature', 'min_curvature', rotated into the PCA frame like dataset.py:368-372) come with ``items()``, the
batched ``__getitem__``.  Point tuples and density jitter are training-time options and raise.  Unlike the reference nothing is written next to the input data (it drops
``.npy`` copies into the dataset directory, dataset.py:227-264).
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch

from . import ops
from .DeepFit import DeepFit
from .tutorial_utils import load_xyz, savetxt, loadtxt_f32


class PointcloudPatchDataset:
    def __init__(self, root, shape_list_filename, patch_radius=None, points_per_patch=256, patch_features=(),
                 seed=None, identical_epochs=False, use_pca=False, center='point', point_tuple=1, cache_capacity=1,
                 point_count_std=0.0, sparse_patches=False, neighbor_search_method='k', device="cuda"):
        if neighbor_search_method not in ('k', 'r'):
            raise ValueError("neighbor_search_method must be 'k' or 'r'")
        if neighbor_search_method == 'r':
            if patch_radius is None or len(patch_radius) != 1:
                raise NotImplementedError("ball-query patches are built for exactly one patch_radius (one scale)")
        self.neighbor_search_method = neighbor_search_method
        self.patch_radius = None if patch_radius is None else [float(r) for r in patch_radius]
        self.identical_epochs = bool(identical_epochs)
        # rng for picking points in a patch, dataset.py:217-220
        self.seed = int(np.random.randint(0, 2**32 - 1)) if seed is None else int(seed)
        self.rng = np.random.RandomState(self.seed)
        # ground-truth features returned next to the patch, dataset.py:197-205
        self.patch_features = tuple(patch_features)
        self.include_normals = self.include_curvatures = self.include_neighbor_normals = False
        for pfeat in self.patch_features:
            if pfeat == 'normal':
                self.include_normals = True
            elif pfeat in ('max_curvature', 'min_curvature'):
                self.include_curvatures = True
            elif pfeat == 'neighbor_normals':
                self.include_neighbor_normals = True
            else:
                raise ValueError('Unknown patch feature: %s' % (pfeat))
        if center != 'point':
            raise NotImplementedError("only center='point' is built (the pretrained configuration)")
        if point_tuple != 1 or point_count_std != 0.0:
            raise NotImplementedError("point tuples / density jitter are training-time options")
        self.root = root
        self.points_per_patch = int(points_per_patch)
        self.use_pca = bool(use_pca)
        self.sparse_patches = bool(sparse_patches)
        self.device = torch.device(device)
        with open(os.path.join(root, shape_list_filename)) as f:
            self.shape_names: List[str] = [x.strip() for x in f.readlines() if x.strip()]
        self.shape_patch_count: List[int] = []
        self._pidx = []
        for name in self.shape_names:
            if self.sparse_patches:
                pidx = np.loadtxt(os.path.join(root, name + '.pidx')).astype('int64').reshape(-1)
                self._pidx.append(pidx)
                self.shape_patch_count.append(len(pidx))
            else:
                self._pidx.append(None)
                self.shape_patch_count.append(None)   # filled when the shape is loaded
        self._loaded = (-1, None, None)
        self._bbdiag = {}
        self._gt = {}

    def __len__(self):
        return sum(self._count(i) for i in range(len(self.shape_names)))

    def _count(self, shape_ind):
        if self.shape_patch_count[shape_ind] is None:
            self.load_shape(shape_ind)
        return self.shape_patch_count[shape_ind]

    def load_shape(self, shape_ind):
        """Points (raw, float32) on the device + their search grid; one shape resident at a time."""
        if self._loaded[0] != shape_ind:
            pts = load_xyz(os.path.join(self.root, self.shape_names[shape_ind] + '.xyz'))
            gpu = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.float32)).to(self.device)
            grid = self._loaded[2]
            if grid is not None and grid.n == int(gpu.shape[0]):
                grid.update(gpu)            # same size (PCPNet's 100 k-point shapes): rebuilt in place, nothing allocated
            else:
                if grid is not None:
                    grid.close()
                grid = ops.Grid(gpu)
            self._loaded = (shape_ind, gpu, grid)
            self._bbdiag[shape_ind] = float(np.linalg.norm(pts.max(0) - pts.min(0), 2))   # dataset.py:262
            # ground-truth normals / curvatures of the shape (dataset.py:13-34, 236-247), only when a feature asks for them
            self._gt = {}
            base = os.path.join(self.root, self.shape_names[shape_ind])
            if self.include_normals or self.include_neighbor_normals:
                self._gt["normals"] = torch.from_numpy(np.ascontiguousarray(load_xyz(base + '.normals'))).to(self.device)
            if self.include_curvatures:
                self._gt["curv"] = torch.from_numpy(np.ascontiguousarray(loadtxt_f32(base + '.curv')).reshape(-1, 2)).to(self.device)
            if self.shape_patch_count[shape_ind] is None:
                self.shape_patch_count[shape_ind] = int(gpu.shape[0])
        return self._loaded[1], self._loaded[2]

    def centres(self, shape_ind) -> Optional[torch.Tensor]:
        p = self._pidx[shape_ind]
        return None if p is None else torch.from_numpy(p.astype(np.int32)).to(self.device)

    def patches(self, shape_ind, start, count, return_index=False):
        """Patches ``start .. start+count`` of a shape (in ``.pidx`` order when sparse):
        (patch f32 [B,3,k], trans f32 [B,3,3], patch_scale f64 [B]) -- the reference returns [k,3] per item
        and its callers transpose (test_c_est.py:131).  ``return_index`` appends the neighbour indices i32 [B,k] and
        the number of real (non-padding) rows per patch i32 [B]."""
        pts, grid = self.load_shape(shape_ind)
        centres = self.centres(shape_ind)
        k = self.points_per_patch
        if self.neighbor_search_method == 'r':
            out = self._ball_patches(shape_ind, pts, grid, centres, start, count)
            return out if return_index else out[:3]
        if centres is None:
            idx, rad = grid.query(k, q_begin=start, q_count=count)
            patch, trans = ops.patch_pca(pts, idx, rad, q_begin=start, pca=self.use_pca)
        else:
            q = centres[start:start + count].contiguous()
            idx, rad = grid.query(k, query=q)
            patch, trans = ops.patch_pca(pts, idx, rad, query=q, pca=self.use_pca)
        if return_index:
            return patch, trans, rad, idx, torch.full((count,), k, dtype=torch.int32, device=self.device)
        return patch, trans, rad

    def items(self, shape_ind, start, count):
        """The batched ``__getitem__`` of the reference (dataset.py:268-417) for patches ``start .. start+count`` of a
        shape: ``(patch_pts f32 [B,k,3],) + features in patch_features order + (trans f32 [B,3,3], patch_scale f64 [B])``.
        Features: 'normal' f32 [B,3] and 'neighbor_normals' f32 [B,k,3] are the shape's ground-truth normals of the centre /
        of the patch points, rotated into the PCA frame when ``use_pca`` (:368-372; padding rows of ball patches stay zero);
        'max_curvature' / 'min_curvature' f32 [B,1] are the ground-truth curvatures times the absolute patch radius (:343-349)."""
        patch, trans, scale, idx, valid = self.patches(shape_ind, start, count, return_index=True)
        centres = self.centres(shape_ind)
        c = (torch.arange(start, start + count, device=self.device) if centres is None else centres[start:start + count]).long()
        feats = {}
        if self.include_normals:
            n = self._gt["normals"][c]
            feats['normal'] = torch.bmm(n[:, None, :], trans)[:, 0] if self.use_pca else n
        if self.include_neighbor_normals:
            nn = self._gt["normals"][idx.long()]
            nn = nn * (torch.arange(self.points_per_patch, device=self.device)[None, :] < valid[:, None])[..., None]
            feats['neighbor_normals'] = torch.bmm(nn, trans) if self.use_pca else nn
        if self.include_curvatures:
            if self.patch_radius is None:
                raise ValueError("curvature features are scaled by patch_radius x bounding-box diagonal: pass patch_radius")
            cv = self._gt["curv"][c] * float(self._bbdiag[shape_ind] * self.patch_radius[0])
            feats['max_curvature'], feats['min_curvature'] = cv[:, 0:1], cv[:, 1:2]
        return (patch.transpose(1, 2).contiguous(),) + tuple(feats[f] for f in self.patch_features) + (trans, scale)


    def _ball_patches(self, shape_ind, pts, grid, centres, start, count):
        """dataset.py:289-334 for neighbor_search_method='r': the ball of radius patch_radius * bbdiag, a seeded random
        subset when it holds more than points_per_patch points, zero rows when fewer; scale = the radius."""
        k = self.points_per_patch
        rad_abs = self._bbdiag[shape_ind] * self.patch_radius[0]
        if centres is None:
            q_idx = np.arange(start, start + count, dtype=np.int64)
            offsets, ball = grid.ball_query(rad_abs, q_begin=start, q_count=count)
        else:
            q = centres[start:start + count].contiguous()
            q_idx = q.cpu().numpy().astype(np.int64)
            offsets, ball = grid.ball_query(rad_abs, query=q)
        off = offsets.cpu().numpy()
        ball_h = ball.cpu().numpy()
        sel = np.empty((count, k), dtype=np.int32)
        valid = np.empty((count,), dtype=np.int32)
        base = sum(self._count(i) for i in range(shape_ind))   # global patch index, for identical_epochs
        for i in range(count):
            inds = ball_h[off[i]:off[i + 1]]
            if self.identical_epochs:
                self.rng.seed((self.seed + base + start + i) % (2**32))
            point_count = int(min(k, len(inds)))
            if point_count < len(inds):
                inds = inds[self.rng.choice(len(inds), point_count, replace=False)]
            sel[i, :point_count] = inds
            sel[i, point_count:] = q_idx[i]     # padding: the centre itself, i.e. a zero row after centring
            valid[i] = point_count
        idx = torch.from_numpy(sel).to(self.device)
        valid_t = torch.from_numpy(valid).to(self.device)
        rad = torch.full((count,), rad_abs, dtype=torch.float64, device=self.device)
        if centres is None:
            patch, trans = ops.patch_pca(pts, idx, rad, q_begin=start, pca=self.use_pca, valid=valid_t)
        else:
            patch, trans = ops.patch_pca(pts, idx, rad, query=q, pca=self.use_pca, valid=valid_t)
        return patch, trans, rad, idx, valid_t


def estimate_shapes(dataset: PointcloudPatchDataset, model: Optional[DeepFit], output_dir: str, jet_order: int = 3,
                    batch: int = 4096, use_point_stn: bool = True, write: bool = True, timings: Optional[dict] = None,
                    use_pipeline: bool = True, keep_results: bool = True, on_shape=None):
    """Run every shape of the dataset and export the five per-shape files of test_c_est.py:201-210.
    ``model=None`` runs the classic unweighted fit (weights file of ones, identity trans).
    Returns {shape_name: dict of numpy arrays} (empty when ``keep_results`` is False: 20 dense 100 k-point shapes
    hold 2 GB of weights).

    Dense k-nearest-neighbour shapes with PCA (the pretrained configuration) go through the C pipeline
    (``dfb_pipeline_run`` with every optional output: one call per shape, 16 384-query chunks, two streams);
    sparse ``.pidx`` centres, ball-query patches and ``use_pca=False`` take the stage-by-stage loop below.  Both give the
    same bits (tests/test_gpu_config3.py).  ``timings`` (a dict) accumulates seconds spent in ``load`` (text -> device +
    grid), ``compute`` (kNN ... curvature, device-synchronised) and ``export`` (device -> host + text files).
    ``on_shape(name, arrays)`` is called after each shape's files are written (outside the timings)."""
    import time
    if write:
        os.makedirs(output_dir, exist_ok=True)
    results = {}
    k = dataset.points_per_patch
    order = model.jet_order if model is not None else jet_order
    dev = dataset.device

    def tick(name, t0):
        if timings is not None:
            torch.cuda.synchronize(dev)
            timings[name] = timings.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    est = None          # the C pipeline and its workspace are kept while consecutive shapes share their grid object
    host = {}           # pinned staging buffers of the five outputs, grown on demand

    def to_host(key, t):
        t = t.contiguous()
        buf = host.get(key)
        if buf is None or buf.numel() < t.numel() or buf.dtype != t.dtype:
            buf = host[key] = torch.empty(t.numel(), dtype=t.dtype, pin_memory=True)
        view = buf[:t.numel()].view(t.shape)
        view.copy_(t, non_blocking=True)
        return view

    for si, name in enumerate(dataset.shape_names):
        t0 = time.perf_counter()
        pts, grid = dataset.load_shape(si)
        n_patch = dataset._count(si)
        t0 = tick("load", t0)
        dense_knn = dataset.neighbor_search_method == 'k' and not dataset.sparse_patches and dataset.use_pca
        if use_pipeline and dense_knn and (model is None or use_point_stn):
            from .pipeline import NormalEstimator
            if est is None or est.grid is not grid:
                if est is not None and getattr(est, "_pipe", None) is not None:
                    est._pipe.close()
                est = NormalEstimator(pts, k, jet_order=order, mode="classic" if model is None else "DeepFit", model=model,
                                      grid=grid)
            est.points = pts
            extras = {}
            normal_prop, curv64 = est.run(extras=extras, want_extras=("beta", "weights", "trans"))
            beta_prop = extras["beta"]
            curv_prop = curv64.float() if curv64 is not None else torch.zeros((n_patch, 2), dtype=torch.float32, device=dev)
            if model is not None:
                weights_prop, trans_prop = extras["weights"], extras["trans"].reshape(n_patch, 9)
            else:
                weights_prop = torch.ones((n_patch, k), dtype=torch.float32, device=dev)
                trans_prop = torch.eye(3, device=dev).reshape(1, 9).repeat(n_patch, 1)
        else:
            net = model._network(dev, max_batch=batch) if model is not None else None
            normal_prop = torch.zeros((n_patch, 3), dtype=torch.float32, device=dev)
            curv_prop = torch.zeros((n_patch, 2), dtype=torch.float32, device=dev)
            weights_prop = torch.ones((n_patch, k), dtype=torch.float32, device=dev)
            beta_prop = torch.zeros((n_patch, ops.JET_M[order]), dtype=torch.float32, device=dev)
            trans_prop = torch.eye(3, device=dev).reshape(1, 9).repeat(n_patch, 1)
            for s in range(0, n_patch, batch):
                c = min(batch, n_patch - s)
                patch, data_trans, rad = dataset.patches(si, s, c)
                if net is not None:
                    weights, trans, _, _ = net.forward(patch, want_trans2=False)
                    out = ops.wls_fit(patch, weights, order, trans=trans if use_point_stn else None,
                                      U=data_trans if dataset.use_pca else None, rad=rad)
                    weights_prop[s:s + c] = weights
                    trans_prop[s:s + c] = trans.reshape(c, 9)
                else:
                    out = ops.wls_fit(patch, None, order, U=data_trans if dataset.use_pca else None, rad=rad)
                normal_prop[s:s + c] = out["n_world"] if "n_world" in out else out["n_local"]
                beta_prop[s:s + c] = out["beta"]
                if order > 1:
                    curv_prop[s:s + c] = out["curv_scaled"].float()     # the reference stores into a float32 buffer
            if net is not None:
                net.status()       # before anything is written: a pipeline time-out raises
        t0 = tick("compute", t0)
        res = {"normals": to_host("normals", normal_prop), "curv": to_host("curv", curv_prop), "weights": to_host("weights", weights_prop),
               "beta": to_host("beta", beta_prop), "trans": to_host("trans", trans_prop)}
        torch.cuda.synchronize(dev)
        res = {key: v.numpy() for key, v in res.items()}      # views of the pinned buffers, valid until the next shape
        if keep_results:
            results[name] = {key: v.copy() for key, v in res.items()}
        if write:
            for ext, arr in res.items():
                savetxt(os.path.join(output_dir, name + '.' + ext), arr)
            print('saved normals for ' + name)
        tick("export", t0)
    return results


def main(argv=None):
    import argparse

    from .pipeline import model_from_state_dict
    ap = argparse.ArgumentParser(description="PCPNet-style inference with full outputs (reference test_c_est.py)")
    ap.add_argument('--indir', type=str, default='./pclouds', help='input folder (point clouds)')
    ap.add_argument('--testset', type=str, default='testset_no_noise.txt', help='shape set file name')
    ap.add_argument('--outdir', type=str, default='./results', help='output folder')
    ap.add_argument('--trained_model_path', type=str, default='../trained_models/DeepFit')
    ap.add_argument('--sparse_patches', type=int, default=False, help='evaluate on the .pidx subset')
    ap.add_argument('--batchSize', type=int, default=4096)
    ap.add_argument('--gpu_idx', type=int, default=0)
    ap.add_argument('--precision', type=str, default='f16x3')
    args = ap.parse_args(argv)
    dev = torch.device("cuda:%d" % args.gpu_idx)
    torch.cuda.set_device(dev)
    params = torch.load(os.path.join(args.trained_model_path, 'DeepFit_params.pth'), weights_only=False)
    sd = torch.load(os.path.join(args.trained_model_path, 'DeepFit.pth'), map_location='cpu')
    model = model_from_state_dict(sd, params.points_per_patch, params.jet_order, params.weight_mode, args.precision, dev)
    ds = PointcloudPatchDataset(args.indir, args.testset, patch_radius=getattr(params, 'patch_radius', None),
                                points_per_patch=params.points_per_patch, seed=getattr(params, 'seed', None),
                                identical_epochs=bool(getattr(params, 'identical_epochs', False)),
                                use_pca=params.use_pca, sparse_patches=bool(args.sparse_patches),
                                neighbor_search_method=params.neighbor_search, device=dev)
    estimate_shapes(ds, model, args.outdir, batch=args.batchSize, use_point_stn=bool(params.use_point_stn))


if __name__ == '__main__':
    main()
