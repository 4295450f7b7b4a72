"""The C ABI used from compiled code: tests/c_client/abi_client.cu (CUDA runtime + include/deepfit_b200.h, no Python, no
torch) runs kNN -> patch/PCA -> classic jet fit on a small cloud; its output must equal the Python binding's bit for bit."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_client(tmp_path):
    from deepfit_b200 import _lib
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("no nvcc on this box")
    _lib.load()
    lib = _lib.lib_path()
    exe = str(tmp_path / "abi_client")
    r = subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", os.path.join(ROOT, "include"),
                        os.path.join(ROOT, "tests", "c_client", "abi_client.cu"), "-o", exe, lib,
                        "-Xlinker", "-rpath," + os.path.dirname(lib)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["f16x3", "bf16x3", "f16f8"])
def test_c_client_runs_deepfit_mode_through_the_pipeline(tmp_path, precision):
    """dfb_model_create + dfb_pipeline_create + one dfb_pipeline_run from compiled code == NormalEstimator.run, bit for bit."""
    from deepfit_b200 import DeepFit as D
    from deepfit_b200 import _lib, synthetic
    from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
    from oracle import deepfit_oracle as O
    exe = _build_client(tmp_path)
    n, k, order, nq = 20000, 256, 3, 1500
    pts = synthetic.torus_cloud(n, seed=7, noise=0.002)
    raw = str(tmp_path / "cloud.f32")
    pts.astype(np.float32).tofile(raw)
    sd = O.seeded_state_dict(0)
    wfile = str(tmp_path / "weights.f32")
    with open(wfile, "wb") as fh:
        for w, b in D.fold_batchnorm(sd):
            fh.write(w.numpy().astype(np.float32).tobytes())
            fh.write(b.numpy().astype(np.float32).tobytes())
    r = subprocess.run([exe, raw, str(n), str(k), str(order), str(nq), wfile, str(_lib.PRECISION[precision])],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.array([[float(x) for x in ln.split()] for ln in r.stdout.strip().splitlines()])
    assert got.shape == (nq, 6)
    dev = torch.from_numpy(pts).cuda()
    model = model_from_state_dict(sd, k, order, "sigmoid", precision, "cuda")
    est = NormalEstimator(dev, k, order, "DeepFit", model, chunk=1024)
    normals, curv = est.run(0, nq)
    assert np.array_equal(got[:, 1:4].astype(np.float32), normals.cpu().numpy())
    assert np.array_equal(got[:, 4:6], curv.cpu().numpy())


@pytest.mark.gpu
def test_c_client_matches_python_binding(tmp_path):
    from deepfit_b200 import ops, synthetic
    exe = _build_client(tmp_path)
    n, k, order, nq = 20000, 256, 3, 500
    pts = synthetic.torus_cloud(n, seed=7, noise=0.002)
    raw = str(tmp_path / "cloud.f32")
    pts.astype(np.float32).tofile(raw)
    r = subprocess.run([exe, raw, str(n), str(k), str(order), str(nq)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    got = np.array([[float(x) for x in ln.split()] for ln in r.stdout.strip().splitlines()])
    assert got.shape == (nq, 6)
    dev = torch.from_numpy(pts).cuda()
    grid = ops.Grid(dev)
    idx, rad = grid.query(k, q_begin=0, q_count=nq)
    patch, U = ops.patch_pca(dev, idx, rad, q_begin=0)
    out = ops.wls_fit(patch, None, order, U=U, rad=rad)
    assert np.array_equal(got[:, 0].astype(np.int64), idx[:, -1].cpu().numpy().astype(np.int64))
    assert np.array_equal(got[:, 1:4].astype(np.float32), out["n_world"].cpu().numpy())
    assert np.array_equal(got[:, 4:6], out["curv_scaled"].cpu().numpy())
