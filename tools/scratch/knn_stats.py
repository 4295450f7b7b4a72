import sys, os, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from deepfit_b200 import ops, synthetic, _lib
from deepfit_b200.tutorial_utils import normalize_cloud
lib = _lib.load()
for name, gen, n in (("torus", synthetic.torus_cloud, 1_000_000), ("density", synthetic.variable_density_cloud, 30_000_000)):
    pts = torch.from_numpy(np.ascontiguousarray(normalize_cloud(gen(n))[0])).cuda()
    grid = ops.Grid(pts)
    out = (C.c_ulonglong * 8)()
    lib.dfb_dbg_knn_stats(out)
    q = 65536
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); idx, rad = grid.query(256, q_begin=0, q_count=q); e1.record(); torch.cuda.synchronize()
    lib.dfb_dbg_knn_stats(out)
    o = [x / q for x in out]
    print(name, "ms %.2f per query: overflow selects %.2f, fallback sorts %.2f, early selects %.2f, candidates %.0f, mean count after overflow select %.0f, levels %.2f" % (e0.elapsed_time(e1), o[0], o[1], o[2], o[3], (out[4] / max(out[0], 1)), o[5]))
    del grid, pts
