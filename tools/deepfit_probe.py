"""DeepFit-mode throughput for a given (k, order) on a multi-surface cloud: python tools/deepfit_probe.py K ORDER POINTS [CHUNK]"""
import sys, os, time, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfit_b200 import synthetic
from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
from deepfit_b200.tutorial_utils import normalize_cloud
k, order, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
pts = torch.from_numpy(np.ascontiguousarray(normalize_cloud(synthetic.multi_surface_cloud(n))[0])).cuda()
sd = synthetic.synthetic_state_dict(0, k, order)
model = model_from_state_dict(sd, k, order, "sigmoid", "f16x3")
est = NormalEstimator(pts, k, order, "DeepFit", model, chunk=int(sys.argv[4]) if len(sys.argv) > 4 else 0)
q = min(n, 131072)
est.net.profile(True)   # creates its event pool now, outside the timed region
for _ in range(2):      # warm-up at full chunk size: allocator, lazily built buffers
    est.run(0, min(q, est.chunk))
torch.cuda.synchronize()
est.net.profile_read()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); nrm, crv = est.run(0, q); e1.record(); torch.cuda.synchronize()
est.net.status()
ms = e0.elapsed_time(e1)
print(json.dumps({"k": k, "order": order, "points": n, "queries": q, "ms": ms, "points_per_s": q / ms * 1e3, "prof": {k_: round(v[0], 2) for k_, v in est.net.profile_read().items()}, "finite": bool(torch.isfinite(nrm).all())}))
