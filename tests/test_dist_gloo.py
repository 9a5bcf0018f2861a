"""world_size-2 gloo test (CPU) of the N>1 host path: scatter the batch from rank 0, process each rank's
chunk independently, gather on rank 0 -- result must be bit-identical to processing the whole batch in one
process (items are independent; no collective inside the path).  The per-chunk processor is the CPU oracle
standing in for the CUDA kernels (no GPU here); the sharding/collective code is the product's."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from dasp_pytorch_b200 import dist as ddist

SR = 44100


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        x = torch.rand(batch, 2, 300) * 2 - 1
        drive = torch.rand(batch * 2) * 24
        gdb = torch.rand(batch) * 12
        full_x = x if rank == 0 else None
        # round trip
        mine = ddist.scatter_batch(full_x, batch, (2, 300), torch.float32, "cpu")
        lo, hi = ddist.shard_bounds(batch, world, rank)
        assert torch.equal(mine, x[lo:hi])
        back = ddist.gather_batch(mine, batch)
        if rank == 0:
            assert torch.equal(back, x)
        # sharded processing == unsharded processing, per item, bit for bit
        y = ddist.process_sharded(lambda xc, d: oracle.distortion(xc, SR, d), full_x, [drive] if rank == 0 else None,
                                  batch, (2, 300), 1, "cpu", rows_per_item=2)
        z = ddist.process_sharded(lambda xc, g: oracle.gain(xc, SR, g), full_x, [gdb] if rank == 0 else None, batch,
                                  (2, 300), 1, "cpu")
        if rank == 0:
            assert torch.equal(y, oracle.distortion(x, SR, drive))
            assert torch.equal(z, oracle.gain(x, SR, gdb))
        else:
            assert y is None and z is None
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [6, 5])
def test_scatter_process_gather_world2(batch):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, batch, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
