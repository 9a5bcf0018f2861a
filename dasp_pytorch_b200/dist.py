"""Batch sharding over the GPUs of one box.

Every processor on the hot path is independent per batch item -- parameters are per item, the compressor's
side chain, the EQ's filter and the reverb's impulse response never mix items, and neither do the
gradients -- so the path shards by contiguous batch chunks with NO collective inside it.  A collective is
only needed at the edges, when one rank holds the whole batch: ``scatter_batch`` hands every rank its chunk
(one grouped send/recv; NCCL over NVLink on the GPU box, gloo in the CPU tests) and
``gather_batch`` collects the processed chunks back the same way.  One process per GPU; launch with
``torch.distributed.run``.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the contiguous item range owned by ``rank``: sizes differ by at most one, ranks with
    lower index get the larger chunks (``torch.tensor_split`` convention)."""
    if batch < 0 or world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad shard request batch={batch} world_size={world_size} rank={rank}")
    base, extra = divmod(batch, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(batch: int, world_size: int) -> List[int]:
    return [hi - lo for lo, hi in (shard_bounds(batch, world_size, r) for r in range(world_size))]


def shard_tensors(tensors: Sequence[torch.Tensor], world_size: int, rank: int, rows_per_item: int = 1):
    """slice dim 0 of every tensor to this rank's items (``rows_per_item`` = 2 for per-(item, channel)
    parameters such as the stereo distortion drive)."""
    out = []
    for t in tensors:
        items = t.shape[0] // rows_per_item
        lo, hi = shard_bounds(items, world_size, rank)
        out.append(t[lo * rows_per_item: hi * rows_per_item])
    return out


def scatter_batch(full: Optional[torch.Tensor], batch: int, tail_shape: Sequence[int], dtype, device,
                  src: int = 0, group=None) -> torch.Tensor:
    """rank ``src`` holds ``full`` of shape ``(batch, *tail_shape)``; every rank returns its chunk.

    One grouped send/recv (``batch_isend_irecv``: a single NCCL group launch on the GPU box): every receiver gets
    exactly its rows straight out of ``full`` -- uneven shard sizes need no padding and the root makes no staging
    copies (its own chunk is a view-copy on the device)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(batch, world)
    mine = torch.empty((sizes[rank], *tail_shape), dtype=dtype, device=device)
    if world == 1:
        mine.copy_(full)
        return mine
    ops = []
    if rank == src:
        off = 0
        for r, sz in enumerate(sizes):
            piece = full[off: off + sz]
            off += sz
            if r == src:
                mine.copy_(piece)
            elif sz > 0:
                ops.append(dist.P2POp(dist.isend, piece.contiguous(), r, group))
    elif sizes[rank] > 0:
        ops.append(dist.P2POp(dist.irecv, mine, src, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return mine


def gather_batch(chunk: torch.Tensor, batch: int, dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """inverse of ``scatter_batch``: rank ``dst`` returns the ``(batch, ...)`` tensor, the others ``None``.
    The receivers write straight into the rows of the result (no padding, no concatenation)."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(batch, world)
    if chunk.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {chunk.shape[0]} items, its shard of {batch} has {sizes[rank]}")
    if world == 1:
        return chunk.clone()
    ops = []
    out = None
    if rank == dst:
        out = torch.empty((batch, *chunk.shape[1:]), dtype=chunk.dtype, device=chunk.device)
        off = 0
        for r, sz in enumerate(sizes):
            if r == dst:
                out[off: off + sz].copy_(chunk)
            elif sz > 0:
                ops.append(dist.P2POp(dist.irecv, out[off: off + sz], r, group))
            off += sz
    elif sizes[rank] > 0:
        ops.append(dist.P2POp(dist.isend, chunk.contiguous(), dst, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


def process_sharded(fn, x_full: Optional[torch.Tensor], params_full: Optional[Sequence[torch.Tensor]], batch: int,
                    x_tail: Sequence[int], n_params: int, device, rows_per_item: int = 1, src: int = 0, group=None):
    """scatter ``x`` and the per-item parameters from ``src``, run ``fn(x_chunk, *param_chunks)`` on every
    rank's own chunk, gather the result on ``src`` (``None`` elsewhere).  ``fn`` is one of the functional
    processors closed over its sample rate; ``params_full[i]`` holds ``batch * rows_per_item`` elements."""
    x = scatter_batch(x_full, batch, x_tail, torch.float32, device, src, group)
    ps = []
    for i in range(n_params):
        full = params_full[i].reshape(batch, rows_per_item).to(torch.float32) if params_full is not None else None
        ps.append(scatter_batch(full, batch, (rows_per_item,), torch.float32, device, src, group).reshape(-1))
    y = fn(x, *ps)
    return gather_batch(y, batch, src, group)
