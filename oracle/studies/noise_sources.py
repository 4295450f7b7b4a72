"""TEST INFRASTRUCTURE ONLY -- where does the f16x3 path's weight noise come from?

    python oracle/studies/noise_sources.py <cloud> [seed] [n_patches]

The CUDA path's point weights are 2-3e-5 off the fp64-exact network (tools/seed_dump.py), the reference's fp32 network
2e-6.  Candidates: (a) the 22-bit operands and the dropped lo*lo term of the split-FP16 products, (b) the tensor core's
fp32 accumulation (products are aligned to the largest exponent of a k-group and truncated, not rounded).  This script
emulates (a) alone -- scaled split exactly as csrc/mlp.cu does it (x * 2^7 = xh + xl, w * 2^8 = wh + wl, all FP16),
three products, accumulated EXACTLY (float64) or in float32 round-to-nearest -- and prints max / median |dw| against the
float64 network.  What is left of the measured 2-3e-5 is (b).
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
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
NP = int(sys.argv[3]) if len(sys.argv) > 3 else 64
golden = np.load(os.path.join(ROOT, "tests/golden/reference_outputs.npz"))
P = torch.from_numpy(golden[name + "/gate_patch"][:NP])
sd = O.seeded_state_dict(seed)


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


def mm(x, W, mode, bmm=False):
    f = torch.bmm if bmm else (lambda a, b: a @ b)
    Wt = W if bmm else W.t()
    if mode == "fp32":
        return f(x, Wt)
    if mode == "fp64":
        return f(x.double(), Wt.double()).float()
    xs, ws = x * 128.0, Wt * 256.0
    xh = fp16(xs); xl = fp16(xs - xh)
    wh = fp16(ws); wl = fp16(ws - wh)
    if mode == "h3_exact_acc":
        d = f(xh.double(), wh.double()) + f(xh.double(), wl.double()) + f(xl.double(), wh.double())
        return (d / 32768.0).float()
    if mode == "h4_exact_acc":
        d = f(xh.double(), wh.double()) + f(xh.double(), wl.double()) + f(xl.double(), wh.double()) + f(xl.double(), wl.double())
        return (d / 32768.0).float()
    if mode == "h3_fp32_acc":
        return (f(xh, wh) + f(xh, wl) + f(xl, wh)) / 32768.0
    if mode.startswith("tc"):
        return tc_mm(xh, xl, wh, wl, mode, bmm) / 32768.0
    raise ValueError(mode)


def rz32(x64):
    """float64 -> float32 rounded TOWARD ZERO (what tools/tc_accum_probe.py measures for tcgen05.mma's fp32 accumulate:
    a bias of -0.5 ulp per MMA on a growing accumulator)."""
    y = x64.float()
    over = y.double().abs() > x64.abs()
    return torch.where(over, torch.nextafter(y, torch.zeros_like(y)), y)


def tc_mm(xh, xl, wh, wl, mode, bmm):
    """One MMA = 16 channels: D = RZ_fp32(C + the exact sum of the 16 products).  Orders of the three products:
    tc          hi*hi, hi*lo, lo*hi per k-step (csrc/mlp.cu mma_kstep)
    tc_cf       corrections first: hi*lo and lo*hi of EVERY k-step, then hi*hi of every k-step
    tc_cf64     corrections first inside each 64-channel block (what a streamed operand ring allows)
    tc_1        hi*hi only, exact operands do not matter: isolates the accumulate (uses hi planes only)"""
    f = torch.bmm if bmm else (lambda a, b: a @ b)
    K = xh.shape[-1]
    C = None

    def step(a, b, ks, C):
        sl = slice(ks * 16, ks * 16 + 16)
        p = f(a[..., sl].double(), b[..., sl, :].double())
        return rz32(p if C is None else C.double() + p)
    nks = K // 16
    if mode == "tc":
        for ks in range(nks):
            for a, b in ((xh, wh), (xh, wl), (xl, wh)):
                C = step(a, b, ks, C)
    elif mode == "tc_cf":
        for ks in range(nks):
            C = step(xh, wl, ks, C)
            C = step(xl, wh, ks, C)
        for ks in range(nks):
            C = step(xh, wh, ks, C)
    elif mode == "tc_cf64":
        for kb in range(0, nks, 4):
            for ks in range(kb, min(kb + 4, nks)):
                C = step(xh, wl, ks, C)
                C = step(xl, wh, ks, C)
            for ks in range(kb, min(kb + 4, nks)):
                C = step(xh, wh, ks, C)
    elif mode == "tc_1":
        for ks in range(nks):
            C = step(xh, wh, ks, C)
    else:
        raise ValueError(mode)
    return C


GROUPS = ["small", "max", "fc", "h1", "h2", "h3"]   # 64-wide chain layers | the three 128 -> 1024 | per-patch FCs | head layers


def net(P, m):
    B, _, k = P.shape
    X = P.transpose(1, 2)
    if isinstance(m, str):
        m = {g_: m for g_ in GROUPS}

    def trunk(pref, x, first_fp32):
        W, b = fold(pref + ".conv1", pref + ".bn1")
        x = torch.relu((x @ W.t() if first_fp32 else mm(x, W, m["small"])) + b)
        W, b = fold(pref + ".conv2", pref + ".bn2")
        x = torch.relu(mm(x, W, m["small"]) + b)
        W, b = fold(pref + ".conv3", pref + ".bn3")
        x = torch.relu(mm(x, W, m["max"]) + b)
        g = x.max(dim=1)[0]
        W, b = fold(pref + ".fc1", pref + ".bn4")
        g = torch.relu(mm(g, W, m["fc"]) + b)
        W, b = fold(pref + ".fc2", pref + ".bn5")
        g = torch.relu(mm(g, W, m["fc"]) + b)
        W, b = fold(pref + ".fc3", None)
        return mm(g, W, m["fc"]) + b

    qv = trunk("feat.pointfeat.stn1", X, True) + torch.tensor([1.0, 0, 0, 0])
    R = O.quat_to_rotmat(qv)
    Xr = X @ R
    W, b = fold("feat.pointfeat.conv1", "feat.pointfeat.bn1")
    x1 = torch.relu(Xr @ W.t() + b)
    W, b = fold("feat.pointfeat.conv2", "feat.pointfeat.bn2")
    x = torch.relu(mm(x1, W, m["small"]) + b)
    t2 = trunk("feat.pointfeat.stn2", x, False)
    t2 = (t2 + torch.eye(64).reshape(1, 4096)).reshape(B, 64, 64)
    pf = mm(x, t2, m["small"], bmm=True)
    W, b = fold("feat.conv2", "feat.bn2")
    y = torch.relu(mm(pf, W, m["small"]) + b)
    W, b = fold("feat.conv3", "feat.bn3")
    y = mm(y, W, m["max"]) + b
    g = y.max(dim=1)[0]
    W, b = fold("conv1", "bn1")
    hg = mm(g, W[:, :1024], m["fc"]) + b
    z = torch.relu(mm(pf, W[:, 1024:], m["h1"]) + hg[:, None, :])
    W, b = fold("conv2", "bn2")
    z = torch.relu(mm(z, W, m["h2"]) + b)
    W, b = fold("conv3", "bn3")
    z = torch.relu(mm(z, W, m["h3"]) + b)
    W, b = fold("conv4", None)
    a = (z @ W.t() + b).squeeze(-1)
    return 0.01 + torch.sigmoid(a), R


sd64 = {k_: (v.double() if v.is_floating_point() else v) for k_, v in sd.items()}
w64, R64, _, _ = O.weight_network(sd64, P.double())
w32, R32, _, _ = O.weight_network(sd, P)
print("# %s seed %d, %d patches; max / median over patches of max_j |w - w_fp64|, and max |R - R_fp64|" % (name, seed, NP))
print("%-44s %.2e / %.2e   R %.2e" % ("reference fp32 network (oracle)", (w32 - w64).abs().max(), (w32 - w64).abs().max(1)[0].median(), (R32 - R64).abs().max()))
for mode, what in (("fp32", "this script's layer order, fp32 matmuls"), ("fp64", "fp64 matmuls, fp32 activations between layers"),
                   ("h3_exact_acc", "scaled split FP16, 3 products, exact accumulate"), ("h4_exact_acc", "... 4 products (lo*lo kept)"),
                   ("h3_fp32_acc", "scaled split FP16, 3 products, fp32 RN accumulate"),
                   ("tc", "tensor-core model: RZ fp32 accumulate per MMA, order hh hl lh per k-step (the kernels)"),
                   ("tc_cf64", "... corrections first inside each 64-channel block"),
                   ("tc_cf", "... corrections of the whole K first, then every hi*hi")):
    w, R = net(P, mode)
    print("%-44s %.2e / %.2e   R %.2e   (%s)" % (mode, (w - w64).abs().max(), (w - w64).abs().max(1)[0].median(), (R - R64).abs().max(), what))
    sys.stdout.flush()

# attribution: which layer groups carry the tensor-core accumulate noise?  One group at a time with an EXACT accumulate
# (the others as the kernels run them), then the sets a two-pass "corrections first" order could cover.
base = {g_: "tc" for g_ in GROUPS}
for tag, kw in ([("%s exact, rest tc" % g_, {g_: "h3_exact_acc"}) for g_ in GROUPS]
                + [("%s corrections-first, rest tc" % g_, {g_: "tc_cf"}) for g_ in ("max", "fc", "h2", "h3")]
                + [("max+h3 corrections-first", {"max": "tc_cf", "h3": "tc_cf"}),
                   ("max+h3+fc corrections-first", {"max": "tc_cf", "h3": "tc_cf", "fc": "tc_cf"}),
                   ("max+h3+fc+h2 corrections-first", {"max": "tc_cf", "h3": "tc_cf", "fc": "tc_cf", "h2": "tc_cf"})]):
    md = dict(base); md.update(kw)
    w, R = net(P, md)
    print("%-44s %.2e / %.2e   R %.2e" % (tag, (w - w64).abs().max(), (w - w64).abs().max(1)[0].median(), (R - R64).abs().max()))
    sys.stdout.flush()
