"""TEST INFRASTRUCTURE ONLY -- CPU emulation of a per-layer-group precision ladder for stage 3.

    python oracle/studies/precision_ladder.py <cloud> [seed] [n_patches]

Question: which layer groups of the weight network can run with ONE tensor-core
product per k-step instead of the three of split-bf16 (bf16x3), and in which operand format, with the
end-to-end parity gate still holding?  Product formats emulated (fp32 accumulate in every case):

    b3  split bf16, hi*hi + hi*lo + lo*hi   (16 mantissa bits per operand; the round-1 parity mode)
    b1  single bf16 product                  (8 bits)
    h1  single fp16 product                  (11 bits; same tcgen05 kind::f16 rate as bf16)
    t1  single tf32 product                  (11 bits; half rate -- listed for reference only)
    h3  split fp16, three products           (22 bits)

Gate: per row, max(0.01 deg, 2 x |reference fp32 - fp64-exact|) for the normal, and
max(1e-3, 2 x the same difference) for the rectified curvature error -- the rule the WLS tests use.
"fp64-exact" is the oracle network + fit evaluated in float64 on the same patches.
Output: one line per configuration: worst angle, worst angle/gate ratio, worst curvature/gate ratio.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import deepfit_oracle as O  # noqa: E402

torch.set_grad_enabled(False)
name = sys.argv[1]
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
NP = int(sys.argv[3]) if len(sys.argv) > 3 else 512
raw = np.load(os.path.join(ROOT, "tests/golden/tutorial_clouds.npz"))[name]
p, bb = O.normalize_cloud(raw)
rng = np.random.default_rng(0)
q = rng.choice(len(p), NP, replace=False)
idx, dist = O.knn_ckdtree(p, q, 256, workers=-1)
dist, idx = O.canonicalize_knn(dist, idx)
items = [O.dataset_item(p, int(c), 256, idx[j], dist[j]) for j, c in enumerate(q)]
P = torch.stack([it[0] for it in items])
Us = torch.stack([it[1] for it in items])
sd = O.seeded_state_dict(seed)


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


def tf32(x):
    u = x.contiguous().view(torch.int32)
    u = (u + 0x1000) & ~0x1FFF           # round-to-nearest (ties away) to 10 explicit mantissa bits
    return u.view(torch.float32)


def split(x, r):
    hi = r(x)
    return hi, r(x - hi)


def mm(x, W, mode, bmm=False):
    f = torch.bmm if bmm else (lambda a, b: a @ b)
    Wt = W if bmm else W.t()
    if mode == "fp32":
        return f(x, Wt)
    if mode in ("b1", "h1", "t1"):
        r = {"b1": bf16, "h1": fp16, "t1": tf32}[mode]
        return f(r(x), r(Wt))
    r = bf16 if mode == "b3" else fp16
    xh, xl = split(x, r)
    wh, wl = split(Wt, r)
    return f(xh, wh) + f(xh, wl) + f(xl, wh)


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
    # the 64->64 layer runs twice in the kernels: inside the STN chain (its precision) and inside the encoder chain
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


# reference (fp32 oracle) and fp64-exact evaluation of the same patches
w_ref, R_ref, t2_ref, pts_ref = O.weight_network(sd, P)
beta_ref, n_ref = O.fit_wjet(pts_ref, w_ref, 3)
nw_ref = O.world_normals(n_ref, Us, R_ref).numpy()
c_ref = O.principal_curvatures(beta_ref).numpy()
sd64 = {k_: (v.double() if v.is_floating_point() else v) for k_, v in sd.items()}
w64, R64, _, pts64 = O.weight_network(sd64, P.double())
beta64, n64 = O.fit_wjet(pts64, w64, 3)
nw64 = O.world_normals(n64, Us.double(), R64).numpy()
c64 = O.principal_curvatures(beta64).numpy()
ref_ang_err = O.unoriented_angle_deg(nw_ref, nw64)
ref_c_err = O.rectified_rel_err(c_ref, c64).max(1)
gate_a = np.maximum(0.01, 2 * ref_ang_err)
gate_c = np.maximum(1e-3, 2 * ref_c_err)
print("# %s seed %d, %d patches: reference fp32 vs fp64: angle max %.2e p99 %.2e med %.2e | curv max %.2e ; rows with a relaxed gate: %d / %d"
      % (name, seed, NP, ref_ang_err.max(), np.quantile(ref_ang_err, 0.99), np.median(ref_ang_err), ref_c_err.max(),
         int((gate_a > 0.01).sum()), int((gate_c > 1e-3).sum())))

GROUPS = ["qstn", "stn", "enc", "fc", "pf", "h1", "h2", "h3"]
base = {g_: "b3" for g_ in GROUPS}


def run(tag, **kw):
    md = dict(base)
    md.update(kw)
    w, R, pts = net(P, md)
    beta, n = O.fit_wjet(pts, w, 3)
    nw = O.world_normals(n, Us, R).numpy()
    ang = O.unoriented_angle_deg(nw, nw_ref)
    ce = O.rectified_rel_err(O.principal_curvatures(beta).numpy(), c_ref).max(1)
    print("%-34s w %.1e R %.1e | angle max %.2e p99 %.2e  max/gate %.2f | curv max %.2e max/gate %.2f  %s"
          % (tag, (w - w_ref).abs().max(), (R - R_ref).abs().max(), ang.max(), np.quantile(ang, 0.99), (ang / gate_a).max(),
             ce.max(), (ce / gate_c).max(), "PASS" if (ang / gate_a).max() <= 1 and (ce / gate_c).max() <= 1 else "fail"))
    sys.stdout.flush()


run("all b3")
run("all h3", **{g_: "h3" for g_ in GROUPS})
for key in GROUPS:
    run(key + "=h1", **{key: "h1"})
for key in ["qstn", "stn", "h1"]:
    run(key + "=b1", **{key: "b1"})
    run(key + "=t1", **{key: "t1"})
run("qstn,stn=h1", qstn="h1", stn="h1")
run("qstn,stn,h1=h1", qstn="h1", stn="h1", h1="h1")
run("qstn,stn,h1,h3=h1", qstn="h1", stn="h1", h1="h1", h3="h1")
run("qstn,stn,h1,h2,h3=h1", qstn="h1", stn="h1", h1="h1", h2="h1", h3="h1")
run("all h1", **{g_: "h1" for g_ in GROUPS})
