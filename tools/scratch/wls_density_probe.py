import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepfit_b200 import ops, synthetic
from deepfit_b200.tutorial_utils import normalize_cloud
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30_000_000
pts = torch.from_numpy(np.ascontiguousarray(normalize_cloud(synthetic.variable_density_cloud(n))[0])).cuda()
grid = ops.Grid(pts)
ev = lambda: torch.cuda.Event(enable_timing=True)
for s in range(0, 65536 * 6, 65536):
    idx, rad, dist = grid.query(256, q_begin=s, q_count=65536, return_dist=True)
    patch, U = ops.patch_pca(pts, idx, rad, q_begin=s)
    torch.cuda.synchronize()
    ts = []
    for rep in range(3):
        a, b = ev(), ev()
        a.record(); out = ops.wls_fit(patch, None, 3, U=U, rad=rad); b.record()
        torch.cuda.synchronize(); ts.append(round(a.elapsed_time(b), 3))
    a, b = ev(), ev()
    a.record(); p2, U2 = ops.patch_pca(pts, idx, rad, q_begin=s); b.record(); torch.cuda.synchronize()
    info = out["info"]
    print(json.dumps({"chunk": s, "wls_ms": ts, "pca_ms": round(a.elapsed_time(b), 3),
                      "info_hist": torch.bincount(info.flatten().long(), minlength=3).tolist(),
                      "rad_min": float(rad.min()), "rad_med": float(rad.median()), "rad_max": float(rad.max()),
                      "patch_absmax": float(patch.abs().max()), "nan": int(torch.isnan(patch).sum()),
                      "keys": list(out.keys())}))
