"""CPU, world_size 2 over gloo: the multi-GPU contract (contiguous point-index shards, one all-gather,
outputs in file order) exercised on the host-side logic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deepfit_b200.pipeline import gather_outputs, shard_range
    b, e = shard_range(n, rank, world)
    g = torch.Generator().manual_seed(123)
    normals = torch.randn(n, 3, generator=g)
    curv = torch.randn(n, 2, generator=g, dtype=torch.float64)
    nn, cc = gather_outputs(normals[b:e].clone(), curv[b:e].clone(), n, rank, world)
    ok = torch.equal(nn, normals) and torch.equal(cc, curv)
    nn2, cc2 = gather_outputs(normals[b:e].clone(), None, n, rank, world)
    ok = ok and torch.equal(nn2, normals) and cc2 is None
    # the hot-path form: each rank writes its range straight into its slot, the gather is in place, nothing is widened
    from deepfit_b200.sharding import GatherBuffers
    bufs = GatherBuffers(n, world, "cpu")
    vn, vc = bufs.local_views(rank)
    ok = ok and vn.shape[0] == e - b and vn.dtype == torch.float32 and vc.dtype == torch.float64
    ok = ok and vn.data_ptr() == bufs.normals[rank * bufs.cap:].data_ptr()
    vn.copy_(normals[b:e])
    vc.copy_(curv[b:e])
    bufs.all_gather(rank)
    fn, fc = bufs.file_order()
    ok = ok and torch.equal(fn, normals) and torch.equal(fc, curv)
    ok = ok and bufs.bytes_per_rank() == bufs.cap * 28
    with open(os.path.join(tmp, "ok%d" % rank), "w") as fh:
        fh.write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 4096, 3])
def test_two_rank_gather_is_file_order(tmp_path, n):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert open(tmp_path / ("ok%d" % r)).read() == "1"


def test_shard_ranges_partition_the_cloud():
    from deepfit_b200.pipeline import shard_capacity, shard_range
    for n in (1, 7, 100000, 120498, 10_000_000):
        for world in (1, 2, 4, 8):
            edges = [shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            for a, b in zip(edges[:-1], edges[1:]):
                assert a[1] == b[0]
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) <= shard_capacity(n, world)
