"""Multi-GPU parity check (run under torchrun on N GPUs of one box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_check.py

Every rank computes its contiguous point-index shard, one NCCL all-gather assembles file order, and each
rank checks the gathered result bit-for-bit against the same cloud computed whole on its own GPU
(SURVEY.md section 8e / tests item 7)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from deepfit_b200 import synthetic
    from deepfit_b200.pipeline import NormalEstimator, model_from_state_dict
    from deepfit_b200.sharding import GatherBuffers, shard_range
    from deepfit_b200.tutorial_utils import normalize_cloud
    from deepfit_b200.synthetic import synthetic_state_dict
    n = 30011  # deliberately not divisible by the world size
    pts = torch.from_numpy(np.ascontiguousarray(normalize_cloud(synthetic.torus_cloud(n, seed=5, noise=0.003))[0])).cuda()
    ok = True
    for mode in ("classic", "DeepFit"):
        model = model_from_state_dict(synthetic_state_dict(0), 256, 3, "sigmoid", "f16x3", "cuda") if mode == "DeepFit" else None
        est = NormalEstimator(pts, 256, 3, mode, model, chunk=2048)
        b, e = shard_range(n, rank, world)
        bufs = GatherBuffers(n, world, pts.device)
        vn, vc = bufs.local_views(rank)
        est.run(b, e, out_normals=vn, out_curv=vc)      # written in place into the rank's slot of the gather buffers
        bufs.all_gather(rank)
        ng, cg = bufs.file_order()
        nf, cf = est.run(0, n)
        same = torch.equal(ng, nf) and torch.equal(cg, cf)
        print("rank %d/%d %s: shard [%d,%d) gathered == whole-cloud: %s" % (rank, world, mode, b, e, same), flush=True)
        ok = ok and same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    if int(flag.item()) != 1:
        sys.exit(1)


if __name__ == "__main__":
    main()
