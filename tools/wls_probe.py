"""Stage 4+5 alone: CUDA-event time of dfb_wls_fit on resident random patches (L2 flushed between launches).

    python tools/wls_probe.py [--k 256] [--n 65536 262144] [--orders 1 2 3 4]

Reports patches/s and the HBM roofline fraction on the algorithmic bytes of SURVEY 8d (4 240 B per weighted patch at
k = 256, order 3; classic = no weights, no R)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepfit_b200 import ops  # noqa: E402

M_OF = {1: 3, 2: 6, 3: 10, 4: 15}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--k", type=int, default=256)
    ap.add_argument("--n", type=int, nargs="+", default=[16384, 65536, 262144])
    ap.add_argument("--orders", type=int, nargs="+", default=[1, 2, 3, 4])
    ap.add_argument("--peak", type=float, default=6556.8)
    args = ap.parse_args()
    dev = torch.device("cuda")
    g = torch.Generator(device=dev).manual_seed(0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for n in args.n:
        patch = (torch.rand((n, 3, args.k), generator=g, device=dev) - 0.5)
        patch[:, 2] *= 0.1
        w = 0.01 + torch.rand((n, args.k), generator=g, device=dev)
        rad = torch.rand((n,), generator=g, device=dev, dtype=torch.float64) + 0.5
        U = torch.eye(3, device=dev).repeat(n, 1, 1).contiguous()
        R = U.clone()
        for order in args.orders:
            for mode in ("classic", "weighted"):
                kw = dict(U=U, rad=rad) if mode == "classic" else dict(U=U, rad=rad, trans=R)
                ww = None if mode == "classic" else w
                ops.wls_fit(patch, ww, order, **kw)
                ts = []
                for _ in range(5):
                    flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ops.wls_fit(patch, ww, order, **kw)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                ms = sorted(ts)[len(ts) // 2]
                M = M_OF[order]
                by = 12 * args.k + (4 * args.k + 36 if mode == "weighted" else 0) + 8 + 36 + 4 * M + 12 + 8 + 4
                print(json.dumps({"n": n, "k": args.k, "order": order, "mode": mode, "ms": round(ms, 4),
                                  "patches_per_s": round(n / (ms * 1e-3)), "bytes_per_patch": by,
                                  "GBps": round(n * by / (ms * 1e-3) / 1e9, 1), "hbm_frac": round(n * by / (ms * 1e-3) / 1e9 / args.peak, 3)}))


if __name__ == "__main__":
    main()
