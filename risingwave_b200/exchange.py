"""Multi-GPU hash shuffle: replaces the Dispatch -> Exchange -> Merge executors for the hash-shuffle
path only (src/stream/src/executor/dispatch.rs:897-1080, exchange/, merge.rs).

One process per GPU.  Per device batch:
  1. `rwgpu_shuffle_partition_device`: vnode = crc32(dist key) % 256 (vnode.rs:45-50), destination =
     vnode_to_dest[vnode] (contiguous equal vnode ranges, the shape of the reference's test mapping
     dispatch.rs:1566-1573), STABLE partition of ops + every column into per-destination regions;
  2. all-to-all of the per-destination row counts, then one all-to-all-v per column buffer over
     NCCL / NVLink (torch.distributed.all_to_all_single with split sizes);
  3. the join / agg kernels consume the received buffers directly (they are ordinary device chunks).
Row order per (source, key) is preserved, and a receiver concatenates sources in rank order, so
per-key order is deterministic.  Barriers are a host-side `dist.barrier()` (they are broadcast in
the reference too, dispatch.rs:940-947).  `U-/U+` pairs whose distribution key changes must be
rewritten to `-/+` before partitioning (`rwgpu_dispatch_rewrite_ops`, dispatch.rs:1001-1019).

The split-size logic is backend-agnostic (`plan_splits`) and is covered on CPU with gloo,
world_size 2 (tests/test_exchange_gloo.py).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def vnode_to_dest_table(world: int, vnode_count: int = 256) -> torch.Tensor:
    """contiguous equal vnode ranges -> destination rank."""
    return (torch.arange(vnode_count, dtype=torch.int64) * world // vnode_count).to(torch.int32)


def plan_splits(send_counts: torch.Tensor, group=None) -> Tuple[List[int], List[int]]:
    """Exchange per-destination row counts; returns (input_split_sizes, output_split_sizes)."""
    world = dist.get_world_size(group)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts, group=group)
    ins = [int(x) for x in send_counts.cpu().tolist()]
    outs = [int(x) for x in recv_counts.cpu().tolist()]
    assert len(ins) == world and len(outs) == world
    return ins, outs


def all_to_all_columns(ops: torch.Tensor, cols: Sequence[torch.Tensor], ins: List[int], outs: List[int], group=None):
    """all-to-all-v of the ops bytes and of every column buffer (rows already grouped by destination)."""
    n_out = sum(outs)
    n_in = sum(ins)
    r_ops = torch.empty(n_out, dtype=ops.dtype, device=ops.device)
    dist.all_to_all_single(r_ops, ops[:n_in], outs, ins, group=group)
    r_cols = []
    for c in cols:
        r = torch.empty(n_out, dtype=c.dtype, device=c.device)
        dist.all_to_all_single(r, c[:n_in], outs, ins, group=group)
        r_cols.append(r)
    return r_ops, r_cols


class ShufflePlan:
    def __init__(self, world: int, rank: int, key_indices: Sequence[int], types: Sequence[int], vnode_count: int = 256):
        self.world, self.rank = world, rank
        self.keys = list(key_indices)
        self.types = list(types)
        self.vnode_count = vnode_count
        self.v2d = vnode_to_dest_table(world, vnode_count).cuda()

    def exchange(self, chunk, stream=None):
        """DeviceChunk -> (ops, cols) holding the rows this rank owns, from all ranks."""
        from . import device
        ops, cols, counts, offsets = device.shuffle_partition(chunk, self.keys, self.v2d, self.world, self.vnode_count, stream)
        ins, outs = plan_splits(counts)
        return all_to_all_columns(ops, cols, ins, outs)
