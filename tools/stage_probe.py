"""Per-stage CUDA-event times of one DeepFit-mode chunk: python tools/stage_probe.py K ORDER POINTS CHUNK"""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfit_b200 import ops, synthetic
from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
from deepfit_b200.tutorial_utils import normalize_cloud
k, order, n, chunk = (int(a) for a in sys.argv[1:5])
pts = torch.from_numpy(np.ascontiguousarray(normalize_cloud(synthetic.multi_surface_cloud(n))[0])).cuda()
model = model_from_state_dict(synthetic.synthetic_state_dict(0, k, order), k, order, "sigmoid", "f16x3")
est = NormalEstimator(pts, k, order, "DeepFit", model, chunk=chunk)
for rep in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    torch.cuda.synchronize()
    ev[0].record(); idx, rad = est.grid.query(k, q_begin=0, q_count=chunk)
    ev[1].record(); patch, U = ops.patch_pca(pts, idx, rad, q_begin=0)
    ev[2].record(); w, trans, _, _ = est.net.forward(patch, want_trans2=False)
    ev[3].record(); out = ops.wls_fit(patch, w, order, trans=trans, U=U, rad=rad)
    ev[4].record(); nrm, crv = est.run(0, chunk)
    ev[5].record(); torch.cuda.synchronize()
names = ["knn", "pca", "network", "wls", "est.run(same chunk)"]
print(json.dumps({"k": k, "chunk": chunk, **{nm: round(ev[i].elapsed_time(ev[i + 1]), 3) for i, nm in enumerate(names)}}))
