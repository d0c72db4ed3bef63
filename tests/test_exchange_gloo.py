"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU hash shuffle (split planning and
all-to-all-v of column buffers), with the vnode partition restated in numpy (CRC32 via zlib)."""
import os
import socket
import zlib

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


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from risingwave_b200 import exchange
    rng = np.random.default_rng(100 + rank)
    n = 1000 + 37 * rank
    key = rng.integers(0, 500, n).astype(np.int64)
    pay = (np.arange(n) + rank * 1_000_000).astype(np.int64)
    ops = rng.integers(1, 3, n).astype(np.uint8)
    v2d = exchange.vnode_to_dest_table(world).numpy()
    vnode = np.array([zlib.crc32(int(k).to_bytes(8, "little", signed=True)) % 256 for k in key])
    dest = v2d[vnode]
    order = np.argsort(dest, kind="stable")  # what rwgpu_shuffle_partition_device produces
    counts = torch.from_numpy(np.bincount(dest, minlength=world).astype(np.int64))
    ins, outs = exchange.plan_splits(counts)
    r_ops, (r_key, r_pay) = exchange.all_to_all_columns(torch.from_numpy(ops[order]), [torch.from_numpy(key[order]), torch.from_numpy(pay[order])], ins, outs)
    # every received key belongs to this rank; sources arrive in rank order with their row order kept
    rv = np.array([zlib.crc32(int(k).to_bytes(8, "little", signed=True)) % 256 for k in r_key.numpy()])
    ok = bool((v2d[rv] == rank).all())
    src = r_pay.numpy() // 1_000_000
    ok = ok and bool((np.diff(src) >= 0).all())
    for s in range(world):
        seg = r_pay.numpy()[src == s]
        ok = ok and bool((np.diff(seg) > 0).all())
    q.put((rank, ok, int(sum(outs)), int(sum(ins)), int(r_ops.numel())))
    dist.barrier()
    dist.destroy_process_group()


def test_all_to_all_v_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in ps]
    assert all(r[1] for r in res), res
    assert sum(r[2] for r in res) == sum(r[3] for r in res) == 1000 + 1037
    assert all(r[2] == r[4] for r in res)
