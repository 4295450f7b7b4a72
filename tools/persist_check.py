"""Persistent kernels vs one CTA per item, bit-equality on 700 patches (dfb_dbg_set keys 8 / 10): python tools/persist_check.py"""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfit_b200 import _lib, ops
from deepfit_b200.pipeline import model_from_state_dict
from oracle import deepfit_oracle as O
lib = _lib.load()
sd = O.seeded_state_dict(0)
g = torch.Generator().manual_seed(3)
B = 700
xy = torch.rand((B, 2, 256), generator=g) * 1.4 - 0.7
c = torch.rand((B, 3, 1), generator=g) - 0.5
zz = 0.3 * (c[:, 0:1] * xy[:, 0:1] ** 2 + c[:, 1:2] * xy[:, 1:2] ** 2 + c[:, 2:3] * xy[:, 0:1] * xy[:, 1:2])
patch = torch.cat([xy, zz], 1).cuda().contiguous()
model = model_from_state_dict(sd, 256, 3, "sigmoid", "f16x3")
net = model._network(patch.device, 1024)
res = {}
for name, k8, k10 in (("both off", 0, 0), ("head persistent", 1, 0), ("chain persistent", 0, 1), ("both", 1, 1)):
    lib.dfb_dbg_set(8, k8); lib.dfb_dbg_set(10, k10)
    w, trans, t2, _ = net.forward(patch, want_trans2=True)
    net.status()
    torch.cuda.synchronize()
    res[name] = (w.clone(), trans.clone(), t2.clone())
ref = res["both off"]
for name, r in res.items():
    d = [float((a - b).abs().max()) for a, b in zip(r, ref)]
    bad = (r[0] - ref[0]).abs().amax(1) > 1e-5
    print(name, "max|dw| %.3g max|dtrans| %.3g max|dt2| %.3g; bad patches: %s" % (d[0], d[1], d[2], torch.nonzero(bad).flatten()[:20].tolist()), int(bad.sum()))
