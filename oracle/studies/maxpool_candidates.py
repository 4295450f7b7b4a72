"""TEST INFRASTRUCTURE ONLY -- how many (point, channel) pairs can still be the maximum of a 128 -> 1024 + max-pool
layer when the products are evaluated ONCE in FP16 (11-bit operands) instead of fp32-grade (three split products)?

    python oracle/studies/maxpool_candidates.py [cloud]

For y[k, c] = w_c . x_k + b_c and its FP16-operand evaluation y~ the rigorous bound is
    |y~ - y| <= (2^-10 + 2^-22) sum_i |x_ki w_ci| <= t_kc = 2^-10 (1 + 2^-12) |x_k|_2 |w_c|_2   (Cauchy-Schwarz),
so only points with y~[k, c] + t_kc >= max_k' (y~[k', c] - t_k'c) can hold the exact maximum; those are the pairs an exact
(fp32 FMA) recomputation has to visit.  Counts per channel are printed for the three max-pool layers (QSTN, STN,
encoder) on the 512 gate patches of tests/golden.
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
G = np.load(os.path.join(ROOT, "tests/golden/reference_outputs.npz"))
P = torch.from_numpy(G[name + "/gate_patch"])[:256]
sd = O.seeded_state_dict(0)


def fold(conv, bn):
    W = sd[conv + ".weight"]
    W = W.squeeze(-1) if W.dim() == 3 else W
    b = sd[conv + ".bias"]
    if bn is None:
        return W.clone(), b.clone()
    s = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + 1e-5)
    return W * s[:, None], (b - sd[bn + ".running_mean"]) * s + sd[bn + ".bias"]


def fp16(x):
    return x.to(torch.float16).to(torch.float32)


def report(tag, x, W, b, relu):
    """x [B, k, 128] fp32 inputs of the layer, W [1024, 128], b [1024]"""
    y = (x.double() @ W.double().t() + b.double())                    # exact
    ya = (fp16(x * 128.0) / 128.0) @ (fp16(W * 256.0) / 256.0).t() + b   # one FP16 product, fp32 accumulate
    err = (ya.double() - y).abs()
    xn = x.norm(dim=2)                                                # [B, k]
    wn = W.norm(dim=1)                                                # [1024]
    t = (2.0 ** -10) * (1 + 2.0 ** -12) * xn[:, :, None] * wn[None, None, :]
    assert (err <= t.double() + 1e-30).all(), "bound violated"
    lo = (ya - t).max(dim=1, keepdim=True)[0]                         # best lower bound of the maximum
    cand = (ya + t) >= lo
    if relu:                                                          # relu(max) == 0 whenever no upper bound is positive
        live = ((ya + t) > 0).any(dim=1, keepdim=True)
        cand = cand & live
    n = cand.sum(dim=1).float()                                       # [B, 1024] candidates per (patch, channel)
    # a cheaper test the kernel could use: one threshold per (patch, channel) from the patch's largest |x_k|
    tc = (2.0 ** -10) * (1 + 2.0 ** -12) * xn.max(dim=1)[0][:, None, None] * wn[None, None, :]
    cand2 = ya >= (ya.max(dim=1, keepdim=True)[0] - 2 * tc)
    if relu:
        cand2 = cand2 & ((ya.max(dim=1, keepdim=True)[0] + tc) > 0)
    n2 = cand2.sum(dim=1).float()
    pts_any = cand.any(dim=2).float().sum(dim=1)                      # points that are a candidate of some channel
    print("%-8s err/bound max %.3f mean %.4f | candidates per channel (per-point bound): mean %.2f p99 %.0f max %.0f ; "
          "(per-patch bound): mean %.2f p99 %.0f max %.0f | points that are candidates somewhere: mean %.0f of %d"
          % (tag, (err / t.double()).max(), (err / t.double()).mean(), n.mean(), n.quantile(0.99), n.max(), n2.mean(),
             n2.quantile(0.99), n2.max(), pts_any.mean(), x.shape[1]))
    # and with a 2-unit evaluation (15-bit operands, f16f8): window 16x narrower
    t8 = t / 16
    c8 = (ya + t8) >= (ya - t8).max(dim=1, keepdim=True)[0]
    print("%-8s   with a 16x narrower window (15-bit operands): mean %.2f p99 %.0f" % ("", c8.sum(dim=1).float().mean(), c8.sum(dim=1).float().quantile(0.99)))
    sys.stdout.flush()


B, _, k = P.shape
X = P.transpose(1, 2)


def trunk_inputs(pref, x):
    W, b = fold(pref + ".conv1", pref + ".bn1")
    x = torch.relu(x @ W.t() + b)
    W, b = fold(pref + ".conv2", pref + ".bn2")
    x = torch.relu(x @ W.t() + b)
    return x


def trunk_rest(pref, x):
    W, b = fold(pref + ".conv3", pref + ".bn3")
    g = torch.relu(x @ W.t() + b).max(dim=1)[0]
    W, b = fold(pref + ".fc1", pref + ".bn4")
    g = torch.relu(g @ W.t() + b)
    W, b = fold(pref + ".fc2", pref + ".bn5")
    g = torch.relu(g @ W.t() + b)
    W, b = fold(pref + ".fc3", None)
    return g @ W.t() + b


xq = trunk_inputs("feat.pointfeat.stn1", X)
report("qstn", xq, *fold("feat.pointfeat.stn1.conv3", "feat.pointfeat.stn1.bn3"), relu=True)
qv = trunk_rest("feat.pointfeat.stn1", xq) + torch.tensor([1.0, 0, 0, 0])
R = O.quat_to_rotmat(qv)
Xr = X @ R
W, b = fold("feat.pointfeat.conv1", "feat.pointfeat.bn1")
x1 = torch.relu(Xr @ W.t() + b)
W, b = fold("feat.pointfeat.conv2", "feat.pointfeat.bn2")
x2 = torch.relu(x1 @ W.t() + b)
xs = trunk_inputs("feat.pointfeat.stn2", x2)
report("stn", xs, *fold("feat.pointfeat.stn2.conv3", "feat.pointfeat.stn2.bn3"), relu=True)
t2 = (trunk_rest("feat.pointfeat.stn2", xs) + torch.eye(64).reshape(1, 4096)).reshape(B, 64, 64)
pf = torch.bmm(x2, t2)
W, b = fold("feat.conv2", "feat.bn2")
xe = torch.relu(pf @ W.t() + b)
report("encoder", xe, *fold("feat.conv3", "feat.bn3"), relu=False)
