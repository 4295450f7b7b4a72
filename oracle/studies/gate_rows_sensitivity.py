"""TEST INFRASTRUCTURE ONLY -- which layer group's product precision limits the curvature gate on the lidar rows?

    python oracle/studies/gate_rows_sensitivity.py [cloud] [gpu_dump.npz]

Works on the 512 per-row-gate patches of tests/golden (the rows the GPU test checks).  (1) With the GPU's own outputs
(tools/dump_gate_outputs.py): swaps our weights / our rotated points one at a time into the reference fit to see which
one carries the curvature error.  (2) CPU emulation (same product formats as precision_ladder.py): every group in b3
except one in fp32, and every group in fp32 except one in b3 / hf (fp16 main + e4m3 correction).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import deepfit_oracle as O  # noqa: E402

torch.set_grad_enabled(False)
name = sys.argv[1] if len(sys.argv) > 1 else "0000000000"
dump = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "gpurun_out", "gate_outputs.npz")
G = np.load(os.path.join(ROOT, "tests/golden/reference_outputs.npz"))
P = torch.from_numpy(G[name + "/gate_patch"])
Us = torch.from_numpy(G[name + "/gate_U"])
rad = G[name + "/gate_rad"]
c_gold = G[name + "/gate_curv_scaled"]
gate_c = np.maximum(1e-3, 2 * G[name + "/gate_floor_curv"])
gate_a = np.maximum(0.01, 2 * G[name + "/gate_floor_angle_deg"])
nw_gold = G[name + "/gate_normal_world"]
sd = O.seeded_state_dict(0)

w_ref, R_ref, t2_ref, pts_ref = O.weight_network(sd, P)


def curv_err(pts, w, tag, R=None):
    beta, n = O.fit_wjet(pts, w, 3)
    c = O.principal_curvatures(beta).numpy().astype(np.float64) / rad[:, None]
    ce = O.rectified_rel_err(c, c_gold).max(1)
    msg = "%-40s curv max %.2e  max/gate %.2f  rows over gate %d" % (tag, ce.max(), (ce / gate_c).max(), int((ce > gate_c).sum()))
    if R is not None:
        nw = O.world_normals(n, Us, R).numpy()
        ang = O.unoriented_angle_deg(nw, nw_gold)
        msg += " | angle max/gate %.2f" % (ang / gate_a).max()
    print(msg)
    sys.stdout.flush()
    return ce


ce0 = curv_err(pts_ref, w_ref, "oracle fp32 (w, pts)", R_ref)
# fp64 fit on fp32 network outputs: how much of the golden-vs-anything difference is the fp32 FIT itself?
beta64, _ = O.fit_wjet(pts_ref.double(), w_ref.double(), 3)
c64 = O.principal_curvatures(beta64).numpy() / rad[:, None]
ce = O.rectified_rel_err(c64, c_gold).max(1)
print("%-40s curv max %.2e  max/gate %.2f  rows over gate %d" % ("fp64 FIT on the fp32 network outputs", ce.max(), (ce / gate_c).max(), int((ce > gate_c).sum())))

if os.path.isfile(dump):
    D = np.load(dump)
    for prec in ("bf16x3", "f16f8"):
        key = "%s/%s/" % (name, prec)
        w_g, pts_g = torch.from_numpy(D[key + "weights"]), torch.from_numpy(D[key + "pts"])
        b_g = torch.from_numpy(D[key + "beta"])
        c = O.principal_curvatures(b_g).numpy().astype(np.float64) / rad[:, None]
        ce = O.rectified_rel_err(c, c_gold).max(1)
        bad = np.nonzero(ce > gate_c)[0]
        print("GPU %s: curv max/gate %.2f rows over gate %s  (|dw| %.2e, |dpts| %.2e)" % (prec, (ce / gate_c).max(), bad.tolist(),
              (w_g - w_ref).abs().max(), (pts_g - pts_ref).abs().max()))
        curv_err(pts_g, w_g, "  oracle fp32 fit on GPU w + GPU pts")
        curv_err(pts_ref, w_g, "  oracle fp32 fit on GPU w + ref pts")
        curv_err(pts_g, w_ref, "  oracle fp32 fit on ref w + GPU pts")
        b64, _ = O.fit_wjet(pts_g.double(), w_g.double(), 3)
        c64 = O.principal_curvatures(b64).numpy() / rad[:, None]
        ce64 = O.rectified_rel_err(c64, c_gold).max(1)
        print("  %-38s curv max/gate %.2f rows over gate %d" % ("fp64 fit on GPU w + GPU pts", (ce64 / gate_c).max(), int((ce64 > gate_c).sum())))
        for r in bad[:6]:
            dwr = (w_g[r] - w_ref[r]).abs()
            print("   row %d: gate %.2e err %.2e  |dw| max %.2e mean %.2e  curv gold %s  floor %.2e" % (r, gate_c[r], ce[r], dwr.max(), dwr.mean(), c_gold[r], G[name + "/gate_floor_curv"][r]))

if "--emulate" not in sys.argv:
    sys.exit(0)

# ------------------------------------------------------------------ CPU emulation per group
import importlib.util  # noqa: E402


def fold(conv, bn):
    W = sd[conv + ".weight"]
    W = W.squeeze(-1) if W.dim() == 3 else W
    b = sd[conv + ".bias"]
    if bn is None:
        return W.clone(), b.clone()
    s = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + 1e-5)
    return W * s[:, None], (b - sd[bn + ".running_mean"]) * s + sd[bn + ".bias"]


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def fp16(x):
    return x.to(torch.float16).to(torch.float32)


def e4m3(x):
    return x.to(torch.float8_e4m3fn).to(torch.float32)


def mm(x, W, mode, bmm=False):
    f = torch.bmm if bmm else (lambda a, b: a @ b)
    Wt = W if bmm else W.t()
    if mode == "fp32":
        return f(x, Wt)
    if mode == "b3":
        xh = bf16(x); xl = bf16(x - xh); wh = bf16(Wt); wl = bf16(Wt - wh)
        return f(xh, wh) + f(xh, wl) + f(xl, wh)
    if mode == "h3":       # split fp16, three products (22 bits); lo planes assumed scaled into the normal range
        xh = fp16(x); xl = x - xh; wh = fp16(Wt); wl = Wt - wh
        xl = fp16(xl * 2048.0) / 2048.0; wl = fp16(wl * 2048.0) / 2048.0
        return f(xh, wh) + f(xh, wl) + f(xl, wh)
    if mode == "hf":       # the kernels' f16f8: fp16 main + e4m3 corrections
        xh = fp16(x * 128.0) / 128.0; xl = x - xh; wh = fp16(Wt * 256.0) / 256.0; wl = Wt - wh
        return f(xh, wh) + f(e4m3(xh), e4m3(wl * 32768.0)) / 32768.0 + f(e4m3(xl * 4096.0), e4m3(wh * 8.0)) / 32768.0
    if mode == "hh":       # fp16 main + fp16 corrections with e4m3-free lo planes but 1 product for both (idealised)
        xh = fp16(x * 128.0) / 128.0; xl = x - xh; wh = fp16(Wt * 256.0) / 256.0; wl = Wt - wh
        return f(xh, wh) + f(xh, wl) + f(xl, wh)
    raise ValueError(mode)


def net(P, md):
    B, _, k = P.shape
    X = P.transpose(1, 2)

    def trunk(pref, x, first_fp32, gm, fm):
        W, b = fold(pref + ".conv1", pref + ".bn1")
        x = torch.relu((x @ W.t() if first_fp32 else mm(x, W, gm)) + b)
        W, b = fold(pref + ".conv2", pref + ".bn2")
        x = torch.relu(mm(x, W, gm) + b)
        W, b = fold(pref + ".conv3", pref + ".bn3")
        x = torch.relu(mm(x, W, gm) + b)
        g = x.max(dim=1)[0]
        W, b = fold(pref + ".fc1", pref + ".bn4")
        g = torch.relu(mm(g, W, fm) + b)
        W, b = fold(pref + ".fc2", pref + ".bn5")
        g = torch.relu(mm(g, W, fm) + b)
        W, b = fold(pref + ".fc3", None)
        return mm(g, W, fm) + b

    qv = trunk("feat.pointfeat.stn1", X, True, md["qstn"], md["fc"])
    qv = qv + torch.tensor([1.0, 0, 0, 0])
    R = O.quat_to_rotmat(qv)
    Xr = X @ R
    W, b = fold("feat.pointfeat.conv1", "feat.pointfeat.bn1")
    x1 = torch.relu(Xr @ W.t() + b)
    W, b = fold("feat.pointfeat.conv2", "feat.pointfeat.bn2")
    xs = torch.relu(mm(x1, W, md["stn"]) + b)
    x = torch.relu(mm(x1, W, md["pf"]) + b)
    t2 = trunk("feat.pointfeat.stn2", xs, False, md["stn"], md["fc"])
    t2 = (t2 + torch.eye(64).reshape(1, 4096)).reshape(B, 64, 64)
    pf = mm(x, t2, md["pf"], bmm=True)
    W, b = fold("feat.conv2", "feat.bn2")
    y = torch.relu(mm(pf, W, md["enc"]) + b)
    W, b = fold("feat.conv3", "feat.bn3")
    y = mm(y, W, md["enc"]) + b
    g = y.max(dim=1)[0]
    W, b = fold("conv1", "bn1")
    hg = mm(g, W[:, :1024], md["fc"]) + b
    z = torch.relu(mm(pf, W[:, 1024:], md["h1"]) + hg[:, None, :])
    W, b = fold("conv2", "bn2")
    z = torch.relu(mm(z, W, md["h2"]) + b)
    W, b = fold("conv3", "bn3")
    z = torch.relu(mm(z, W, md["h3"]) + b)
    W, b = fold("conv4", None)
    a = (z @ W.t() + b).squeeze(-1)
    return 0.01 + torch.sigmoid(a), R, Xr.transpose(1, 2).contiguous()


GROUPS = ["qstn", "stn", "enc", "fc", "pf", "h1", "h2", "h3"]


def run(tag, md):
    w, R, pts = net(P, md)
    print("  |dw| %.2e |dR| %.2e" % ((w - w_ref).abs().max(), (R - R_ref).abs().max()), end=" ")
    curv_err(pts, w, tag, R)


for base in ("b3", "hf"):
    run("all " + base, {g_: base for g_ in GROUPS})
    for g_ in GROUPS:
        md = {x: base for x in GROUPS}
        md[g_] = "fp32"
        run("all %s, %s=fp32" % (base, g_), md)
    for g_ in GROUPS:
        md = {x: "fp32" for x in GROUPS}
        md[g_] = base
        run("all fp32, %s=%s" % (g_, base), md)
run("all h3", {g_: "h3" for g_ in GROUPS})
run("all hh", {g_: "hh" for g_ in GROUPS})
