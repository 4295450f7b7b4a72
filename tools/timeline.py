"""Per-CTA phase timeline of the two fused network kernels (profiling aid; clock64 stamps, see DFB_TL in mlp.cu).

    python tools/timeline.py [--cta 2000] [--batch 4096]

Prints, for one traced CTA of each chain kernel (QSTN / STN / encoder) and of the head kernel, the cycle offsets
of the phase boundaries relative to the CTA's first instruction."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHAIN = {0: "entry", 1: "setup done", 2: "P done t0", 3: "P done t1", 22: "M start (X ready)", 23: "first W3 stage landed",
         44: "  L0 t0 epi: tmem_ld done", 45: "  L0 t0 epi: stores issued", 46: "  L0 t0 epi: proxy fence done",
         40: "epi grp0 end", 41: "epi grp1 end", 42: "exit", 43: "TMA: W3 stage 0 issued", 47: "TMA: W3 stage 2 issued", 48: "mma: buffer 0 free for ct0"}
for l in range(3):
    for t in range(2):
        b = 4 + (l * 2 + t) * 3
        CHAIN[b] = "L%d t%d mma issue" % (l, t)
        CHAIN[b + 1] = "L%d t%d acc ready" % (l, t)
        CHAIN[b + 2] = "L%d t%d epi done" % (l, t)
for ct in range(8):
    CHAIN[24 + ct] = "M ct%d issued" % ct
    CHAIN[32 + ct] = "M ct%d acc ready" % ct
HEAD = {0: "tile start", 2: "pf in TMEM", 52: "D2 ready", 53: "h2 written", 54: "mma: h2 ready", 55: "mma: L3 issued",
        56: "D3 ready", 57: "epi end", 58: "exit", 59: "TMA: first W3 issue"}
for c in range(8):
    b = 4 + c * 6
    for i, nm in enumerate(["mma1 issue", "D1 ready", "-", "h1 written", "mma2 start", "mma2 issued"]):
        HEAD[b + i] = "c%d %s" % (c, nm)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cta", type=int, default=2000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--precision", type=str, default="f16x3")
    args = ap.parse_args()
    from deepfit_b200 import _lib
    from deepfit_b200.pipeline import model_from_state_dict
    from deepfit_b200.synthetic import synthetic_state_dict
    lib = _lib.load()
    model = model_from_state_dict(synthetic_state_dict(0), 256, jet_order=3, precision=args.precision)
    g = torch.Generator(device="cuda").manual_seed(1)
    patch = torch.randn(args.batch, 3, 256, device="cuda", generator=g) * 0.3
    net = model._network(patch.device, args.batch)
    for _ in range(3):
        net.forward(patch)
    torch.cuda.synchronize()
    if os.environ.get("DFB_HEAD_DBG"):
        lib.dfb_dbg_set(6, int(os.environ["DFB_HEAD_DBG"]))   # head experiment bits (results are garbage, timing only)
    lib.dfb_dbg_set(7, args.cta)
    lib.dfb_dbg_set(9, 1)
    net.forward(patch)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 256)()
    lib.dfb_dbg_timeline.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    rc = lib.dfb_dbg_timeline(buf, 256)
    lib.dfb_dbg_set(7, -1)
    lib.dfb_dbg_set(9, 0)
    assert rc == 0
    n_cta = min(16384, args.batch * 2)
    tr = (C.c_longlong * (3 * n_cta))()
    lib.dfb_dbg_cta_trace.argtypes = [C.POINTER(C.c_longlong), C.c_int]
    assert lib.dfb_dbg_cta_trace(tr, n_cta) == 0
    tr = np.array(tr[:], dtype=np.int64).reshape(n_cta, 3)
    dur, gap = [], []
    for sm in np.unique(tr[:, 0]):
        rows = tr[tr[:, 0] == sm]
        rows = rows[np.argsort(rows[:, 1])]
        dur.extend((rows[:, 2] - rows[:, 1]).tolist())
        gap.extend((rows[1:, 1] - rows[:-1, 2]).tolist())
    print("== head kernel, all %d tiles: resident %.2f us median (p10 %.2f, p90 %.2f); gap between consecutive tiles on one SM "
          "%.2f us median (p10 %.2f, p90 %.2f); kernel span %.1f us"
          % (n_cta, np.median(dur) / 1e3, np.percentile(dur, 10) / 1e3, np.percentile(dur, 90) / 1e3, np.median(gap) / 1e3,
             np.percentile(gap, 10) / 1e3, np.percentile(gap, 90) / 1e3, (tr[:, 2].max() - tr[:, 1].min()) / 1e3))
    tl = np.array(buf[:], dtype=np.int64).reshape(4, 64)
    for region, name, names in ((0, "QSTN chain", CHAIN), (1, "STN chain", CHAIN), (2, "encoder chain", CHAIN), (3, "head", HEAD)):
        t0 = tl[region, 0]
        print("== %s, CTA %d" % (name, args.cta))
        rows = sorted(((int(tl[region, s] - t0), nm) for s, nm in names.items() if tl[region, s] != 0))
        prev = 0
        for t, nm in rows:
            print("%8d  (+%6d)  %s" % (t, t - prev, nm))
            prev = t


if __name__ == "__main__":
    main()
